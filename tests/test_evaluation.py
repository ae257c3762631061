"""Metric pipeline (difffacto_b200/metrics/evaluation.py) against the reference's evaluation_utils.py: the pure-torch
statistics (knn, lgan_mmd_cov, lgan_mmd_cov_match) on CPU against golden outputs minted from the reference, and - on the
GPU - the all-pairs CD/EMD matrices against the per-pair kernels and the C oracle."""
import os

import numpy as np
import pytest
import torch

from difffacto_b200.metrics import evaluation as E

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def eg():
    return np.load(os.path.join(HERE, "golden", "eval_golden.npz"))


@pytest.mark.parametrize("tag", ["sq", "rect"])
def test_knn_and_mmd_cov_match_reference(eg, tag):
    Mrs, Mrr, Mss = (torch.from_numpy(eg[f"{tag}_{k}"]) for k in ("Mrs", "Mrr", "Mss"))
    nn = E.knn(Mrr, Mrs, Mss, 1, sqrt=False)
    nn1 = E.knn(Mrr, Mrs, Mss, 1, sqrt=True, one_way=True)
    mc = E.lgan_mmd_cov(Mrs.t())
    mm, midx = E.lgan_mmd_cov_match(Mrs.t())
    for name, res in (("knn", nn), ("knn1way", nn1), ("mmdcov", mc), ("match", mm)):
        for k, v in res.items():
            assert float(v) == pytest.approx(float(eg[f"{tag}_{name}_{k}"]), rel=1e-6, abs=1e-7), (name, k)
    assert np.array_equal(midx.numpy(), eg[f"{tag}_match_idx"])


@pytest.mark.gpu
def test_pairwise_matrices_equal_per_pair_kernels():
    """The flattened 512-pair blocks give exactly the reference's per-(sample, ref-batch) results, masks included."""
    from difffacto_b200.metrics import EMD
    from oracle import pointnet2_oracle as O
    g = torch.Generator().manual_seed(3)
    S, R, n = 5, 7, 1024
    smp = torch.rand(S, n, 3, generator=g).cuda()
    ref = torch.rand(R, n, 3, generator=g).cuda()
    ms = (torch.rand(S, n, generator=g) > 0.3).float().cuda()
    mr = (torch.rand(R, n, generator=g) > 0.3).float().cuda()
    cd, emd = E._pairwise_EMD_CD_(smp, ref, 3, verbose=False, mask_sample=ms, mask_ref=mr)
    assert cd.shape == emd.shape == (S, R)
    emd_mod = EMD(0.002, 10000, True)
    for i in (0, 4):  # the reference's loop body for one sample: expand, one call per reference batch
        exp = smp[i].view(1, -1, 3).expand(R, -1, -1).contiguous()
        dl, dr = E.distChamferCUDA(exp, ref)
        want = (dl * ms[i].unsqueeze(0)).sum(1) / ms[i].sum() + (dr * mr).sum(1) / mr.sum(1)
        assert torch.allclose(cd[i], want, rtol=1e-6, atol=1e-7)
        assert torch.allclose(emd[i], emd_mod(exp, ref), rtol=2e-3, atol=1e-4)  # the auction is order-dependent only through ties
    # Chamfer against the CPU oracle (bit-exact distances)
    cd_plain, _ = E._pairwise_EMD_CD_(smp[:2], ref[:2], 32, verbose=False)
    for i in range(2):
        for j in range(2):
            d1, d2, _, _ = O.chamfer_forward(smp[i:i + 1].cpu().numpy(), ref[j:j + 1].cpu().numpy())
            assert abs(cd_plain[i, j].item() - (d1.mean() + d2.mean())) < 1e-6


@pytest.mark.gpu
def test_compute_all_metrics_keys_and_sanity():
    g = torch.Generator().manual_seed(9)
    ref = torch.rand(6, 1024, 3, generator=g)
    smp = ref[torch.tensor([0, 1, 2, 3])] + 0.001 * torch.randn(4, 1024, 3, generator=g)  # samples = jittered references
    res = E.compute_all_metrics(smp.clamp(0, 1), ref, 32)
    for k in ("lgan_mmd-CD", "lgan_cov-CD", "lgan_mmd_smp-CD", "lgan_mmd-EMD", "lgan_cov-EMD", "lgan_mmd_smp-EMD",
              "1-NN-CD-acc", "1-NN-CD-acc_t", "1-NN-CD-acc_f", "1-NN-EMD-acc", "1-NN-EMD-acc_t", "1-NN-EMD-acc_f"):
        assert k in res and torch.isfinite(res[k]).all(), k
    assert res["lgan_mmd_smp-CD"].item() < 1e-4          # every sample sits on a reference shape
    assert res["lgan_cov-CD"].item() == pytest.approx(4 / 6)  # distinct samples that are some reference's nearest / N_ref (reference :270)
    r = E.EMD_CD(smp.clamp(0, 1), ref[:4], 32)
    assert r["MMD-CD"].item() < 1e-4 and r["MMD-EMD"].item() < 0.01
