# Denoiser training on the B200 path: the `diffusion=dict(...)` block and training hyper-parameters of the reference's
# configs/train_chair_stage1.py (:29-70, :113-139) with the synthetic part-segmented dataset standing in for
# ShapeNetSegPart + the part encoder (out of scope).
_base_ = 'gen_chair.py'
model = dict(num_timesteps=200, ret_traj=False)
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
