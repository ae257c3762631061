"""Generation-metric pipeline on the B200 Chamfer / EMD kernels, behind the reference's API
(python/difffacto/datasets/evaluation_utils.py: distChamferCUDA :17-19, emd_approx :84-89, EMD_CD :106-139,
_pairwise_EMD_CD_ :143-200, knn :205-244, lgan_mmd_cov :247-276, lgan_mmd_cov_match :279-291,
compute_all_metrics :500-541).  Same function names, arguments and result keys.

What changes is the shape of the work.  The reference walks a Python double loop - for every sample cloud, for every
`batch_size` reference clouds: expand, one Chamfer launch, one EMD call (70 000 kernel launches at the evaluation
setting) - so an S x R evaluation issues S*ceil(R/batch_size) small calls that each fill a fraction of the GPU.  Every
(sample, reference) pair is independent, so here the S*R pairs are flattened and cut into blocks of up to
`PAIR_BLOCK` = 512 pairs (the EMD kernel's batch contract): one Chamfer launch (both directions) and one persistent
EMD launch per 512 pairs, every SM busy.  Results are identical pair by pair (the reference's `batch_size` only
controls its launch granularity); masks follow the reference's broadcasting exactly.
"""
import torch

from .chamfer import ChamferDistanceL2_split
from .emd import EMD

PAIR_BLOCK = 512  # pairs per kernel call (dfb200_emd_forward accepts at most 512 cloud pairs)


def distChamferCUDA(x, y):
    """reference :17-19"""
    return ChamferDistanceL2_split(reduce=False)(x, y)


def emd_approx(sample, ref):
    """reference :84-89: auction EMD at the evaluation setting (eps 0.002, 10000 rounds), sqrt(dist).mean(1)"""
    assert sample.size(1) == ref.size(1), "Not sure what would EMD do in this case"
    return EMD(0.002, 10000, True)(sample, ref)


def EMD_CD(sample_pcs, ref_pcs, batch_size, accelerated_cd=True, reduced=True):
    """reference :106-139: matched pairs (sample i vs reference i)."""
    N_sample, N_ref = sample_pcs.shape[0], ref_pcs.shape[0]
    assert N_sample == N_ref, "REF:%d SMP:%d" % (N_ref, N_sample)
    cd_lst, emd_lst = [], []
    for b_start in range(0, N_sample, PAIR_BLOCK):
        s, r = sample_pcs[b_start:b_start + PAIR_BLOCK].cuda(), ref_pcs[b_start:b_start + PAIR_BLOCK].cuda()
        dl, dr = distChamferCUDA(s, r)
        cd_lst.append(dl.mean(dim=1) + dr.mean(dim=1))
        emd_lst.append(emd_approx(s, r))
    cd, emd = torch.cat(cd_lst), torch.cat(emd_lst)
    if reduced:
        cd, emd = cd.mean(), emd.mean()
    return {'MMD-CD': cd, 'MMD-EMD': emd}


def _pairwise_EMD_CD_(sample_pcs, ref_pcs, batch_size, accelerated_cd=True, verbose=True, mask_sample=None, mask_ref=None):
    """All-pairs CD and EMD matrices (N_sample, N_ref), reference :143-200.  `batch_size`, `accelerated_cd` and `verbose`
    are accepted for interface parity; the pair blocking is PAIR_BLOCK (see module docstring)."""
    N_sample, N_ref = sample_pcs.shape[0], ref_pcs.shape[0]
    dev = torch.device("cuda", torch.cuda.current_device())
    sample_pcs, ref_pcs = sample_pcs.to(dev).float(), ref_pcs.to(dev).float()
    if mask_sample is not None:
        mask_sample = mask_sample.to(dev).float()
    if mask_ref is not None:
        mask_ref = mask_ref.to(dev).float()
    total = N_sample * N_ref
    all_cd = torch.empty(total, device=dev)
    all_emd = torch.empty(total, device=dev)
    flat = torch.arange(total, device=dev)
    for p0 in range(0, total, PAIR_BLOCK):
        pid = flat[p0:p0 + PAIR_BLOCK]
        si, ri = pid // N_ref, pid % N_ref
        s, r = sample_pcs.index_select(0, si), ref_pcs.index_select(0, ri)
        dl, dr = distChamferCUDA(s, r)
        if mask_sample is not None:
            ms = mask_sample.index_select(0, si)
            dl_mean = (dl * ms).sum(1) / ms.sum(1)
        else:
            dl_mean = dl.mean(1)
        if mask_ref is not None:
            mr = mask_ref.index_select(0, ri)
            dr_mean = (dr * mr).sum(1) / mr.sum(1)
        else:
            dr_mean = dr.mean(1)
        all_cd[p0:p0 + PAIR_BLOCK] = dl_mean + dr_mean
        all_emd[p0:p0 + PAIR_BLOCK] = emd_approx(s, r)
    return all_cd.view(N_sample, N_ref), all_emd.view(N_sample, N_ref)


def knn(Mxx, Mxy, Myy, k, sqrt=False, one_way=False):
    """1-NN two-sample test, reference :205-244 (adapted from xuqiantong/GAN-Metrics)."""
    n0, n1 = Mxx.size(0), Myy.size(0)
    label = torch.cat((torch.ones(n0), torch.zeros(n1))).to(Mxx)
    M = torch.cat([torch.cat((Mxx, Mxy), 1), torch.cat((Mxy.transpose(0, 1), Myy), 1)], 0)
    if sqrt:
        M = M.abs().sqrt()
    INFINITY = float('inf')
    val, idx = (M + torch.diag(INFINITY * torch.ones(n0 + n1).to(Mxx))).topk(k, 0, False)
    count = torch.zeros(n0 + n1).to(Mxx)
    for i in range(0, k):
        count = count + label.index_select(0, idx[i])
    pred = torch.ge(count, (float(k) / 2) * torch.ones(n0 + n1).to(Mxx)).float()
    if one_way:
        pred = pred[:n0]
        label = pred[:n0]  # (sic) the reference compares the prediction with itself in one-way mode
    s = {
        'tp': (pred * label).sum(),
        'fp': (pred * (1 - label)).sum(),
        'fn': ((1 - pred) * label).sum(),
        'tn': ((1 - pred) * (1 - label)).sum(),
    }
    s.update({
        'precision': s['tp'] / (s['tp'] + s['fp'] + 1e-10),
        'recall': s['tp'] / (s['tp'] + s['fn'] + 1e-10),
        'acc_t': s['tp'] / (s['tp'] + s['fn'] + 1e-10),
        'acc_f': s['tn'] / (s['tn'] + s['fp'] + 1e-10),
        'acc': torch.eq(label, pred).float().mean(),
    })
    return s


def lgan_mmd_cov(all_dist, thresh=1000):
    """MMD / coverage of latent_3d_points with DiffFacto's outlier rule, reference :247-276."""
    N_sample, N_ref = all_dist.size(0), all_dist.size(1)
    min_val_fromsmp, min_idx = torch.min(all_dist, dim=1)
    min_val, idx = torch.min(all_dist, dim=0)
    min_val, idxx = torch.sort(min_val)
    min_val_fromsmp, _ = torch.sort(min_val_fromsmp)
    sorted_idx = idx[idxx]
    outlier_mask = min_val > thresh
    if torch.any(outlier_mask):
        sorted_idx[outlier_mask] = sorted_idx[0]
    mmd = min_val.mean()
    mmd_smp = min_val_fromsmp.mean()
    cov = float(sorted_idx.unique().view(-1).size(0)) / float(N_ref)
    cov = torch.tensor(cov).to(all_dist)
    return {'lgan_mmd': mmd, 'lgan_cov': cov, 'lgan_mmd_smp': mmd_smp}


def lgan_mmd_cov_match(all_dist):
    """reference :279-291"""
    N_sample, N_ref = all_dist.size(0), all_dist.size(1)
    min_val_fromsmp, min_idx = torch.min(all_dist, dim=1)
    min_val, _ = torch.min(all_dist, dim=0)
    mmd = min_val.mean()
    mmd_smp = min_val_fromsmp.mean()
    cov = float(min_idx.unique().view(-1).size(0)) / float(N_ref)
    cov = torch.tensor(cov).to(all_dist)
    return {'lgan_mmd': mmd, 'lgan_cov': cov, 'lgan_mmd_smp': mmd_smp}, min_idx.view(-1)


def compute_all_metrics(sample_pcs, ref_pcs, batch_size, accelerated_cd=True, one_way=False, mask=None):
    """MMD / COV / 1-NN accuracy under CD and EMD, reference :500-541 (same result keys)."""
    results = {}
    M_rs_cd, M_rs_emd = _pairwise_EMD_CD_(ref_pcs, sample_pcs, batch_size, accelerated_cd=accelerated_cd, mask_ref=mask)
    results.update({"%s-CD" % k: v for k, v in lgan_mmd_cov(M_rs_cd.t()).items()})
    results.update({"%s-EMD" % k: v for k, v in lgan_mmd_cov(M_rs_emd.t()).items()})
    M_rr_cd, M_rr_emd = _pairwise_EMD_CD_(ref_pcs, ref_pcs, batch_size, accelerated_cd=accelerated_cd)
    if not one_way:
        M_ss_cd, M_ss_emd = _pairwise_EMD_CD_(sample_pcs, sample_pcs, batch_size, accelerated_cd=accelerated_cd, mask_ref=mask,
                                              mask_sample=mask)
    else:
        INFINITY = float('inf')
        M_ss_cd = torch.zeros(M_rs_cd.shape[1], M_rs_cd.shape[1]).to(M_rs_cd) + INFINITY
        M_ss_emd = torch.zeros(M_rs_cd.shape[1], M_rs_cd.shape[1]).to(M_rs_cd) + INFINITY
    one_nn_cd_res = knn(M_rr_cd, M_rs_cd, M_ss_cd, 1, sqrt=False, one_way=one_way)
    results.update({"1-NN-CD-%s" % k: v for k, v in one_nn_cd_res.items() if 'acc' in k})
    one_nn_emd_res = knn(M_rr_emd, M_rs_emd, M_ss_emd, 1, sqrt=False, one_way=one_way)
    results.update({"1-NN-EMD-%s" % k: v for k, v in one_nn_emd_res.items() if 'acc' in k})
    return results
