"""Import-name shim: the reference does `import emd` (python/difffacto/metrics/emd/emd_module.py:26) and calls
`emd.forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments, unass_idx, unass_cnt,
unass_cnt_sum, cnt_tmp, max_idx, eps, iters)` / `emd.backward(xyz1, xyz2, gradxyz, graddist, idx)` -- the two functions of its
compiled extension (emd.cpp:25-28), all buffers caller-allocated.  With this repo on PYTHONPATH the name resolves to the B200
auction kernel (difffacto_b200/csrc/metrics.cu)."""
from difffacto_b200.metrics.emd import emd_backward as backward  # noqa: F401
from difffacto_b200.metrics.emd import emd_forward as forward  # noqa: F401

__version__ = "1.0.0+b200"
