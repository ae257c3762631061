import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, importlib.util
from difffacto_b200.metrics import emdFunction
from oracle import pointnet2_oracle as O
spec = importlib.util.spec_from_file_location("ref_emd", "oracle/_ref/ref_emd.so"); E = importlib.util.module_from_spec(spec); spec.loader.exec_module(E)
cu = lambda a: torch.from_numpy(a).cuda()
for n, eps, iters in [(1024, 0.005, 50), (2048, 0.005, 50), (1024, 0.002, 10000), (4096, 0.005, 50)]:
    rng = np.random.default_rng(n + iters)
    B = 4 if n < 4096 else 2
    a = rng.random((B, n, 3)).astype(np.float32); b = rng.random((B, n, 3)).astype(np.float32)
    got = [np.sqrt(emdFunction.apply(cu(a), cu(b), eps, iters)[0].cpu().numpy()).mean(1) for _ in range(2)]
    odist, _, _ = O.emd_forward(a, b, eps, iters); exp = np.sqrt(odist).mean(1)
    refs = []
    for _ in range(3):
        z = lambda *s, dt=torch.float32: torch.zeros(*s, device="cuda", dtype=dt)
        rdist, rass = z(B, n), z(B, n, dt=torch.int32) - 1
        E.forward(cu(a), cu(b), rdist, rass, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n, dt=torch.int32), z(B, n), z(B, n),
                  z(B * n, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32), z(B * n, dt=torch.int32), eps, iters)
        refs.append(torch.sqrt(rdist).mean(1).cpu().numpy())
    rel = lambda x, y: np.abs(x - y).max() / np.abs(y).max()
    print(f"n={n} eps={eps} iters={iters}: ours run-to-run {rel(got[0], got[1]):.2e}  ours-vs-oracle {rel(got[0], exp):.2e}  ours-vs-ref {rel(got[0], refs[0]):.2e}  "
          f"ref run-to-run {max(rel(refs[0], refs[1]), rel(refs[0], refs[2])):.2e}  ref-vs-oracle {rel(refs[0], exp):.2e}")
