# Stage-1 training on the B200 path: the encoder / diffusion blocks and training hyper-parameters of the reference's
# configs/train_chair_stage1.py (:4-70, :113-139) with the synthetic part-segmented dataset standing in for ShapeNetSegPart.
_base_ = 'gen_chair.py'
model = dict(
    num_timesteps=200, ret_traj=False,
    encoder=dict(                 # reference configs/train_chair_stage1.py:4-28: PointNetV2 + latent-flow prior, ground-truth part parameters
        _cover_=True,
        type='PartEncoderForTransformerDecoder',
        encoder=dict(type='PointNetV2', zdim=256, point_dim=3, per_part_mlp=True),
        n_class=4, kl_weight=5e-4, fit_loss_type=4, fit_loss_weight=1.0, use_flow=True, latent_flow_depth=14,
        latent_flow_hidden_dim=256, include_z=False, include_part_code=True, include_params=True, use_gt_params=True,
        kl_weight_annealing=False, min_kl_weight=1e-7, kl_weight_annealing_end_epoch=4000, gen=True, prior_var=1.0,
    ),
)
dataset = dict(
    train=dict(type="SyntheticPartSeg", batch_size=128, npoints=2048, n_parts=4, num_batches=8, seed=1),
    val=dict(type="SyntheticPartSeg", batch_size=32, npoints=2048, n_parts=4, num_batches=1, seed=0),
)
optimizer = dict(type='Adam', lr=0.002, weight_decay=0.)
max_epoch = 8000
checkpoint_interval = 500
log_interval = 50
max_norm = 10
precision = "fp32"   # the differentiable path runs the fp32 primitives; sampling with the trained weights may use bf16
