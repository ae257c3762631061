#!/usr/bin/env python
"""Entry point with the reference's CLI (tools/run_net.py:8-124): --config-file, --task, --prefix, --launcher, --seed ...
`--task val` runs reverse-diffusion sampling on the B200 path, `--task train` trains the diffusion denoiser; the other tasks of
the reference are outside this build."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser(description="DiffFacto sampling on B200")
    ap.add_argument("--config-file", default="", metavar="FILE", type=str)
    ap.add_argument("--task", default="val", type=str, help="train,val,val_gen,interpolation (val and train are built here)")
    ap.add_argument("--prefix", default="", type=str)
    ap.add_argument("--launcher", choices=["none", "pytorch"], default="none")
    ap.add_argument("--local_rank", type=int, default=0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no_cuda", action="store_true")
    ap.add_argument("--sync_bn", action="store_true")
    ap.add_argument("--deterministic", action="store_true")
    ap.add_argument("--gen_num", type=int, default=32)
    ap.add_argument("--param_sample_num", type=int, default=1)
    ap.add_argument("--num_timesteps", type=int, default=None, help="override model.num_timesteps (e.g. 1000)")
    args = ap.parse_args()
    if args.no_cuda:
        raise SystemExit("difffacto_b200 has no CPU path (--no_cuda is not supported)")
    import torch
    import torch.distributed as dist
    from difffacto_b200.config import get_cfg, init_cfg
    import difffacto_b200  # noqa: F401  registers NETS / DIFFUSIONS / METRICS / DATASETS
    import difffacto_b200.datasets  # noqa: F401
    from difffacto_b200.runner import Runner
    distributed = args.launcher == "pytorch" or int(os.environ.get("WORLD_SIZE", "1")) > 1
    local = int(os.environ.get("LOCAL_RANK", args.local_rank))
    torch.cuda.set_device(local)
    if distributed:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.config_file:
        init_cfg(args.config_file)
    cfg = get_cfg()
    if args.prefix:
        cfg.name = f"{args.prefix}_{cfg.name}"
        cfg.work_dir = f"work_dirs/{cfg.name}"
    if args.num_timesteps:
        cfg.model.num_timesteps = args.num_timesteps
    runner = Runner(f"cuda:{local}", args)
    if args.task == "val":
        runner.val()
    elif args.task == "train":
        runner.run()
    elif args.task == "val_gen":
        runner.generate_samples(args.gen_num, args.param_sample_num)
    else:
        raise SystemExit(f"--task {args.task} is outside the B200 build (val: sampling, val_gen: generation from the prior, train: denoiser training)")
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
