"""Op-level benchmark (GPU box): PointNet++ ops, Chamfer and EMD on the SA/FP shapes of SURVEY.md section 8d, against the
reference's own kernels compiled for sm_100a (oracle/_ref) when present.  Prints one JSON line per case:
time (CUDA events, median of 20 after 5 warm-ups, L2 flushed between iterations), algorithmic bytes, GB/s, fraction of the
measured HBM peak, and speed-up over the reference kernel."""
import importlib.util
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

from gpu_util import cu, part_cloud
from difffacto_b200.pointnet2_ops import pointnet2_utils as pu
from difffacto_b200.metrics.chamfer import chamfer_forward
from difffacto_b200.metrics import emdFunction


def load_ref():
    out = {}
    for name in ("ref_pointnet2_ext", "ref_chamfer", "ref_emd"):
        p = os.path.join(ROOT, "oracle", "_ref", name + ".so")
        if os.path.exists(p):
            spec = importlib.util.spec_from_file_location(name, p)
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            out[name] = m
    return out


REF = load_ref()
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]  # optional op-name filters, e.g. `bench_ops.py group ball`


def want(name):
    return not ONLY or any(o in name for o in ONLY)


def timeit(fn, iters=20, warm=5, flush=True, name=None):
    if name is not None and not want(name):
        return None
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            FLUSH.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))


def report(name, shape, ours_us, ref_us, alg_bytes, note=""):
    if ours_us is None:
        return
    gbs = alg_bytes / ours_us / 1e3
    print(json.dumps({"op": name, "shape": shape, "ours_us": round(ours_us, 2), "ref_us": None if ref_us is None else round(ref_us, 2),
                      "speedup_vs_ref_kernel": None if ref_us is None else round(ref_us / ours_us, 2), "alg_bytes": alg_bytes,
                      "alg_GBps": round(gbs, 1), "frac_hbm_peak": round(gbs / PEAK, 4), "note": note}), flush=True)


def main():
    rng = np.random.default_rng(0)
    E = REF.get("ref_pointnet2_ext")
    for B in (32, 256):
        xyz = cu(part_cloud(rng, B, 2048))
        # FPS 2048 -> 512 (SA1), 512 -> 128 (SA2)
        for (n, m) in ((2048, 512), (512, 128)):
            x = xyz[:, :n].contiguous()
            o = timeit(lambda: pu.furthest_point_sample(x, m), name="furthest_point_sampling")
            r = timeit(lambda: E.furthest_point_sampling(x, m), name="furthest_point_sampling") if E else None
            report("furthest_point_sampling", f"B={B} n={n} m={m}", o, r, B * (12 * n + 4 * m), "latency bound: m-1 dependent rounds")
        sel = pu.furthest_point_sample(xyz, 512)
        xyz_t = xyz.transpose(1, 2).contiguous()
        new_xyz = pu.gather_operation(xyz_t, sel).transpose(1, 2).contiguous()
        o = timeit(lambda: pu.gather_operation(xyz_t, sel), name="gather_points")
        r = timeit(lambda: E.gather_points(xyz_t, sel), name="gather_points") if E else None
        report("gather_points", f"B={B} C=3 n=2048 m=512", o, r, B * 4 * (3 * 2048 + 512 + 3 * 512))
        for (rad, ns) in ((0.2, 64), (0.1, 16), (0.4, 128)):
            o = timeit(lambda: pu.ball_query(rad, ns, xyz, new_xyz), name="ball_query")
            r = timeit(lambda: E.ball_query(new_xyz, xyz, rad, ns), name="ball_query") if E else None
            report("ball_query", f"B={B} n=2048 m=512 r={rad} ns={ns}", o, r, B * (12 * 2048 + 12 * 512 + 4 * 512 * ns),
                   "O(M*N) FP32 pair tests with ordered early exit; HBM floor quoted")
        idx = pu.ball_query(0.2, 64, xyz, new_xyz)
        for C in (7, 131):
            feats = torch.randn(B, C, 2048, device="cuda")
            o = timeit(lambda: pu.grouping_operation(feats, idx), name="group_points")
            r = timeit(lambda: E.group_points(feats, idx), name="group_points") if E else None
            report("group_points", f"B={B} C={C} n=2048 np=512 ns=64", o, r, B * 4 * (C * 2048 + 512 * 64 + C * 512 * 64))
        feats2 = torch.randn(B, 320, 512, device="cuda")
        idx2 = pu.ball_query(0.4, 64, new_xyz, new_xyz[:, :128].contiguous())
        o = timeit(lambda: pu.grouping_operation(feats2, idx2), name="group_points")
        r = timeit(lambda: E.group_points(feats2, idx2), name="group_points") if E else None
        report("group_points", f"B={B} C=320 n=512 np=128 ns=64", o, r, B * 4 * (320 * 512 + 128 * 64 + 320 * 128 * 64))
        known = new_xyz
        o = timeit(lambda: pu.three_nn(xyz, known), name="three_nn")
        r = timeit(lambda: E.three_nn(xyz, known), name="three_nn") if E else None
        report("three_nn", f"B={B} n=2048 m=512", o, r, B * (12 * 2048 + 12 * 512 + 24 * 2048), "O(n*m) FP32 pair tests; HBM floor quoted")
        dist, i3 = pu.three_nn(xyz, known)
        w = torch.softmax(-dist, -1).contiguous()
        f3 = torch.randn(B, 256, 512, device="cuda")
        o = timeit(lambda: pu.three_interpolate(f3, i3, w), name="three_interpolate")
        r = timeit(lambda: E.three_interpolate(f3, i3, w), name="three_interpolate") if E else None
        report("three_interpolate", f"B={B} c=256 m=512 n=2048", o, r, B * 4 * (256 * 512 + 6 * 2048 + 256 * 2048))
    C = REF.get("ref_chamfer")
    for (B, n) in ((32, 2048), (256, 2048), (1024, 2048), (16, 8192), (1024, 512)):
        a, b = torch.rand(B, n, 3, device="cuda"), torch.rand(B, n, 3, device="cuda")
        o = timeit(lambda: chamfer_forward(a, b), name="chamfer_forward")
        r = timeit(lambda: C.forward(a, b), name="chamfer_forward") if C else None
        report("chamfer_forward", f"B={B} n=m={n}", o, r, B * (24 * n + 16 * n), f"{2 * B * n * n / o / 1e6:.2f} Tpair/s (FP32 ALU bound)" if o else "")
    Em = REF.get("ref_emd")
    for (B, n, eps, iters) in ((32, 2048, 0.005, 50), (32, 2048, 0.002, 10000), (4, 8192, 0.005, 50)):
        a, b = torch.rand(B, n, 3, device="cuda"), torch.rand(B, n, 3, device="cuda")
        o = timeit(lambda: emdFunction.apply(a, b, eps, iters), iters=3, warm=1, flush=False, name="emd_forward")
        r = None
        if Em:
            z = lambda *s, dt=torch.float32: torch.zeros(*s, device="cuda", dtype=dt)

            def ref():
                Em.forward(a, b, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n, dt=torch.int32),
                           z(B, n), z(B, n), z(B * n, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32),
                           z(512, dt=torch.int32), z(B * n, dt=torch.int32), eps, iters)
            r = timeit(ref, iters=3, warm=1, flush=False, name="emd_forward")
        report("emd_forward", f"B={B} n={n} eps={eps} iters={iters}", o, r, B * 24 * n,
               "one persistent kernel vs 7 launches per auction round in the reference")


if __name__ == "__main__":
    main()
