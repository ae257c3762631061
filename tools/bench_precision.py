"""GPU box: the `precision_modes` block of the bench line on its own (shapes/s on a bounded sample + max |eps - oracle| per mode).
    python tools/bench_precision.py [modes...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tools import bench_blocks

modes = tuple(sys.argv[1:]) or ("bf16", "tf32", "fp32")
torch.cuda.set_device(0)
out = bench_blocks.precision_block(torch, bench.build_model, bench.synthetic_batch, 32, 2048, 1000, bench.FLOP_PER_POINT_STEP, modes=modes)
print(json.dumps(out))
