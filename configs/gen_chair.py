# Chair generation, B200 build.  Same keys / values as the sampling-relevant part of the reference's
# configs/gen_chair.py (model.encoder :6-46 - generation path only -, model.diffusion :48-85, num_timesteps / npoints
# :88-89); the ShapeNet dataset is outside this build, so synthetic part-segmented clouds stand in for it (--task val).
# The reference's own configs (/root/reference/configs/*.py) load unmodified through difffacto_b200.config.
denoiser = dict(
    type='TransformerNet',
    in_channels=3, out_channels=3,
    n_heads=8, d_head=16, depth=5,
    context_dim=256 + 6,          # part style code + (mean, variance) of the part
    n_class=4, class_cond=True,
    cat_params_to_x=True, cat_class_to_x=True,
    use_linear=True, single_attn=True, use_checkpoint=False,
    dropout=0.2,
)

model = dict(
    type='AnchorDiffAE',
    encoder=dict(                 # reference configs/gen_chair.py:6-46; only its generation path (sample_latents) is built here
        type='PartEncoderForTransformerDecoder',
        encoder=dict(type='PointNetV2', zdim=256, point_dim=3, per_part_mlp=True),
        part_aligner=dict(
            type="PartAlignerTransformer", in_channels=256, out_channels=6, n_class=4, d_head=32, depth=5, n_heads=8, dropout=0.,
            use_checkpoint=False, use_linear=True, class_cond=True, single_attn=True, add_class_cond=True, cimle=True,
            noise_scale=100, cond_noise_type=0),
        n_class=4, kl_weight=0, fit_loss_type=4, fit_loss_weight=1.0, use_flow=True, latent_flow_depth=14,
        latent_flow_hidden_dim=256, include_z=False, include_part_code=True, include_params=True, use_gt_params=False,
        kl_weight_annealing=False, gen=True, prior_var=1.0,
    ),
    diffusion=dict(
        type='AnchoredDiffusion',
        net=denoiser,
        mode='linear', beta_1=1e-4, beta_T=.02, k=1.0,
        model_mean_type="epsilon", loss_type='mse',
        res=False, include_anchors=False, learn_variance=True,
        use_beta=False, rescale_timesteps=False,
        guidance=False, classifier_weight=1.,
        ddim_sampling=False, ddim_nsteps=25, ddim_discretize='quad', ddim_eta=1.,
    ),
    sampler=dict(type='Uniform'),
    num_anchors=4,
    num_timesteps=100,            # shipped value; the BASELINE metric runs the same path with 1000
    npoints=2048,
    gen=True, ret_traj=True, ret_interval=10,
)

dataset = dict(
    val=dict(type="SyntheticPartSeg", batch_size=32, npoints=2048, n_parts=4, num_batches=1, seed=0),
)

logger = dict(type="RunLogger")
precision = "bf16"                # "bf16": tcgen05 tensor cores, "fp32": CUDA-core reference numerics
