"""GPU box: the `train` block of the bench line on its own (1 GPU, or under torchrun)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from tools import bench_blocks

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "0"  # torchrun exports 1
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
out = bench_blocks.train_block(torch, dist, bench.build_model, bench.synthetic_batch, world, rank, local, bench.FLOP_PER_POINT_STEP)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
