"""GPU box: why is the generator block (decode() loop over p_sample_loop_progressive) sometimes host-bound?  Device allocator counters
and wall time per decode(), in the order of bench.py's blocks and after torch.cuda.empty_cache()."""
import gc, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from tools import bench_blocks as BB
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
B, N, T = bench.B_PER_GPU, bench.NPTS, bench.T_STEPS
diff = bench.build_model(T, "bf16").to(dev).eval()
res = {k: v.to(dev) for k, v in bench.synthetic_batch(100, B, N).items()}

def decode():
    final = dict()
    for t, sample in diff.p_sample_loop_progressive([B, 3, N], anchors=res["anchors"], variance=res["variance"], ctx=[res["code"], res["params"]],
                                                    noise=None, anchor_assignment=res["assign"], valid_id=res["valid"], device="cuda", progress=False):
        if t == 0:
            final["pred"] = sample["sample"].transpose(2, 1)
        elif t % 10 == 0:
            final[t] = sample["sample"].transpose(2, 1)
    return final

def run(tag, n=3):
    final = None
    for i in range(n):
        s0 = torch.cuda.memory_stats()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        final = decode()
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        s1 = torch.cuda.memory_stats()
        print(f"{tag} decode {i}: {dt * 1e3:.1f} ms, cudaMalloc calls {s1['num_device_alloc'] - s0['num_device_alloc']}, cudaFree calls "
              f"{s1['num_device_free'] - s0['num_device_free']}, reserved {torch.cuda.memory_reserved() >> 20} MB", flush=True)

run("fresh process:")
for s in range(3):  # the headline loop of bench.py
    diff.p_sample_loop([B, 3, N], res["anchors"], ctx=[res["code"], res["params"]], variance=res["variance"], anchor_assignment=res["assign"],
                       valid_id=res["valid"], rng="philox", seed=s)
out = BB.train_block(torch, dist, bench.build_model, bench.synthetic_batch, 1, 0, 0, bench.FLOP_PER_POINT_STEP)
run("after headline loop + train block:")
gc.collect(); torch.cuda.empty_cache()
run("after empty_cache:")
