"""-m gpu: the sm_100a PointNet++ ops (through the C ABI / pointnet2_ops API) against
 (1) the C oracle restatement and (2) the reference's own kernels compiled for sm_100a (oracle/_ref).
Index outputs and forward values must be BIT-EXACT; atomics-based grads within 1e-5."""
import numpy as np
import pytest
import torch

from gpu_util import cu, part_cloud
from oracle import pointnet2_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pu():
    from difffacto_b200.pointnet2_ops import pointnet2_utils
    return pointnet2_utils


FPS_SHAPES = [(3, 2048, 512), (2, 512, 128), (2, 128, 32), (1, 1000, 77), (2, 8192, 256), (1, 33, 33), (1, 1, 1),
              (1, 5000, 64), (1, 12000, 40), (2, 4096, 300), (1, 1024, 1024), (2, 600, 100), (1, 300, 64), (1, 70, 70),
              (1, 511, 50), (1, 513, 50)]


@pytest.mark.parametrize("B,N,M", FPS_SHAPES)
def test_fps_bit_exact(pu, ref_ext, B, N, M):
    rng = np.random.default_rng(N * 7 + M)
    xyz = part_cloud(rng, B, N)
    got = pu.furthest_point_sample(cu(xyz), M).cpu().numpy()
    assert got.dtype == np.int32 and got.shape == (B, M)
    assert np.array_equal(got, O.furthest_point_sampling(xyz, M))
    if "ref_pointnet2_ext" in ref_ext:
        ref = ref_ext["ref_pointnet2_ext"].furthest_point_sampling(cu(xyz), M).cpu().numpy()
        assert np.array_equal(got, ref)


def test_fps_tie_rule_and_degenerate(pu, ref_ext):
    xyz = np.tile(np.array([[1.0, 1.0, 1.0]], np.float32), (512, 1))
    xyz[128] = xyz[256] = [3.0, 1.0, 1.0]
    assert pu.furthest_point_sample(cu(xyz[None]), 2)[0, 1].item() == 256
    # heavy ties: integer lattice, many equal distances
    rng = np.random.default_rng(5)
    lat = rng.integers(1, 4, (2, 2048, 3)).astype(np.float32)
    got = pu.furthest_point_sample(cu(lat), 64).cpu().numpy()
    assert np.array_equal(got, O.furthest_point_sampling(lat, 64))
    if "ref_pointnet2_ext" in ref_ext:
        assert np.array_equal(got, ref_ext["ref_pointnet2_ext"].furthest_point_sampling(cu(lat), 64).cpu().numpy())
    z = np.zeros((1, 64, 3), np.float32)  # every point skipped by the |p|^2 <= 1e-3 rule
    assert pu.furthest_point_sample(cu(z), 8).cpu().tolist() == [[0] * 8]


def test_fps_temp_output_matches_reference_state():
    from difffacto_b200 import _lib
    rng = np.random.default_rng(3)
    xyz = part_cloud(rng, 2, 700)
    x = cu(xyz)
    idx = torch.empty(2, 50, dtype=torch.int32, device="cuda")
    temp = torch.empty(2, 700, device="cuda")
    _lib.check(_lib.load().dfb200_furthest_point_sampling(2, 700, 50, _lib.ptr(x), _lib.ptr(temp), _lib.ptr(idx), _lib.stream()))
    oidx, otemp = O.furthest_point_sampling(xyz, 50, return_temp=True)
    assert np.array_equal(idx.cpu().numpy(), oidx) and np.array_equal(temp.cpu().numpy(), otemp)


BQ = [(0.2, 64, 2048, 512), (0.4, 64, 512, 128), (0.1, 16, 2048, 512), (0.4, 128, 2048, 512), (0.8, 128, 512, 128),
      (0.05, 7, 300, 50), (10.0, 5, 100, 3), (0.3, 33, 20000, 64),
      # grid path (1024 <= n <= 4096): ragged m, dense fall-back (0.8 / 2.0), tiny radius, n not a multiple of 32
      (0.1, 16, 1024, 300), (0.2, 32, 4096, 777), (0.05, 8, 3001, 100), (0.8, 128, 2048, 512), (2.0, 64, 2048, 100),
      (0.013, 4, 2048, 33), (0.2, 200, 2048, 512), (0.3, 1, 1500, 257), (0.15, 40, 8192, 700), (0.1, 16, 2048, 31)]


@pytest.mark.parametrize("r,ns,N,M", BQ)
def test_ball_query_bit_exact(pu, ref_ext, r, ns, N, M):
    rng = np.random.default_rng(int(r * 100) + ns + N)
    B = 3
    xyz = part_cloud(rng, B, N)
    new_xyz = xyz[:, rng.permutation(N)[:M]].copy()
    new_xyz[:, 0] = 50.0  # an empty ball
    got = pu.ball_query(r, ns, cu(xyz), cu(new_xyz)).cpu().numpy()
    assert got.dtype == np.int32 and np.array_equal(got, O.ball_query(new_xyz, xyz, r, ns))
    assert (got[:, 0] == 0).all()
    if "ref_pointnet2_ext" in ref_ext:
        ref = ref_ext["ref_pointnet2_ext"].ball_query(cu(new_xyz), cu(xyz), r, ns).cpu().numpy()
        assert np.array_equal(got, ref)


def test_ball_query_grid_path_edge_cases(pu, ref_ext):
    """Inputs that stress the binned path: centres outside / on the faces of the cloud's bounding box, a degenerate
    (single-point) cloud, planar clouds (one grid axis collapses), non-finite coordinates and r = 0 (ordered-scan
    fall-back inside the same kernel), many clouds (256-thread variant)."""
    rng = np.random.default_rng(11)
    N = 2048
    cases = []
    xyz = part_cloud(rng, 2, N)
    far = np.concatenate([xyz[:, :300] + 0.15 * rng.standard_normal((2, 300, 3)).astype(np.float32),
                          3.0 * rng.standard_normal((2, 100, 3)).astype(np.float32),
                          np.stack([xyz.min(1), xyz.max(1)], 1),
                          np.stack([xyz.min(1) - 0.05, xyz.max(1) + 0.05], 1)], 1).astype(np.float32)
    cases += [(xyz, far, 0.1, 16), (xyz, far, 0.25, 48)]
    same = np.tile(np.array([[[0.3, -0.2, 0.1]]], np.float32), (1, N, 1))
    cases.append((same, same[:, :40] + np.float32(0.01), 0.1, 8))
    plane = part_cloud(rng, 1, N)
    plane[..., 2] = 0.25
    cases.append((plane, plane[:, :128].copy(), 0.1, 32))
    bad = part_cloud(rng, 3, N)
    bad[0, 5] = np.nan
    bad[1, 7, 1] = np.inf
    bad[2, 9, 2] = -np.inf
    cases.append((bad, bad[:, 100:200].copy(), 0.2, 16))
    cases.append((xyz, xyz[:, :64].copy(), 0.0, 4))
    many = part_cloud(rng, 300, 1024)
    cases.append((many, many[:, rng.permutation(1024)[:256]].copy(), 0.15, 24))
    for (p, c, r, ns) in cases:
        got = pu.ball_query(r, ns, cu(p), cu(c)).cpu().numpy()
        assert np.array_equal(got, O.ball_query(c, p, r, ns)), (p.shape, c.shape, r, ns)
        if "ref_pointnet2_ext" in ref_ext:
            assert np.array_equal(got, ref_ext["ref_pointnet2_ext"].ball_query(cu(c), cu(p), r, ns).cpu().numpy())


def test_ball_query_grid_path_large_batch_sorted_centres(pu, ref_ext):
    """Batch large enough that every persistent CTA runs >= 6 passes of a cloud: the grid kernel then sorts the
    centres by cell (bitonic sort) and the slices of different CTAs must still cover every centre exactly once."""
    rng = np.random.default_rng(21)
    xyz = part_cloud(rng, 200, 2048)
    sel = np.stack([rng.permutation(2048)[:500] for _ in range(200)])
    new_xyz = np.take_along_axis(xyz, sel[..., None].repeat(3, -1), 1).copy()
    for (r, ns) in ((0.2, 32), (0.25, 20)):
        got = pu.ball_query(r, ns, cu(xyz), cu(new_xyz)).cpu().numpy()
        assert np.array_equal(got, O.ball_query(new_xyz, xyz, r, ns))
        if "ref_pointnet2_ext" in ref_ext:
            assert np.array_equal(got, ref_ext["ref_pointnet2_ext"].ball_query(cu(new_xyz), cu(xyz), r, ns).cpu().numpy())


def test_ball_query_thread_per_centre_path(pu, ref_ext):
    """Shapes the default policy sends to ball_query_tpc_kernel (nsample >= 48, batch x centres >= one CTA per SM): ragged n
    (not a multiple of 32) and m (partial warps / CTAs), dense balls that stop early, empty balls (centres far from the cloud),
    non-finite points, r = 0 - bit-exact against the C oracle and the reference kernel."""
    rng = np.random.default_rng(33)
    cases = []
    xyz = part_cloud(rng, 90, 777)
    sel = np.stack([rng.permutation(777)[:433] for _ in range(90)])
    ctr = np.take_along_axis(xyz, sel[..., None].repeat(3, -1), 1).copy()
    ctr[:, :7] += np.float32(5.0)          # empty balls
    cases += [(xyz, ctr, 0.2, 64), (xyz, ctr, 0.45, 48), (xyz, ctr, 0.0, 50), (xyz, ctr, 0.05, 129)]
    big = part_cloud(rng, 150, 2048)
    big[3, 11] = np.nan
    big[4, 12, 0] = np.inf
    cases.append((big, big[:, 1000:1300].copy(), 0.3, 100))
    for (p, c, r, ns) in cases:
        got = pu.ball_query(r, ns, cu(p), cu(c)).cpu().numpy()
        assert np.array_equal(got, O.ball_query(c, p, r, ns)), (p.shape, c.shape, r, ns)
        if "ref_pointnet2_ext" in ref_ext:
            assert np.array_equal(got, ref_ext["ref_pointnet2_ext"].ball_query(cu(c), cu(p), r, ns).cpu().numpy())


@pytest.mark.parametrize("B,C,N,NP,NS", [(4, 7, 2048, 512, 64), (2, 131, 512, 128, 64), (2, 320, 512, 128, 32), (1, 3, 50, 7, 5), (2, 4, 100, 9, 1)])
def test_group_and_gather_bit_exact(pu, ref_ext, B, C, N, NP, NS):
    rng = np.random.default_rng(C + NP)
    feats = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, NP, NS)).astype(np.int32)
    f = cu(feats).requires_grad_(True)
    out = pu.grouping_operation(f, cu(idx))
    assert np.array_equal(out.detach().cpu().numpy(), O.group_points(feats, idx))
    go = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(cu(go))
    assert np.allclose(f.grad.cpu().numpy(), O.group_points_grad(go, idx, N), rtol=1e-5, atol=1e-5)
    gidx = idx[:, :, 0].copy()
    f2 = cu(feats).requires_grad_(True)
    g = pu.gather_operation(f2, cu(gidx))
    assert np.array_equal(g.detach().cpu().numpy(), O.gather_points(feats, gidx))
    gg = rng.standard_normal(g.shape).astype(np.float32)
    g.backward(cu(gg))
    assert np.allclose(f2.grad.cpu().numpy(), O.gather_points_grad(gg, gidx, N), rtol=1e-5, atol=1e-5)
    if "ref_pointnet2_ext" in ref_ext:
        E = ref_ext["ref_pointnet2_ext"]
        assert torch.equal(out.detach(), E.group_points(cu(feats), cu(idx)))
        assert torch.equal(g.detach(), E.gather_points(cu(feats), cu(gidx)))


def test_gather_all_shapes_of_the_sampling_path(pu):
    # PartEncoder.gather_all: (B,3,4)/(B,1,4) part params broadcast to 2048 points (part_encoders.py:417-428)
    rng = np.random.default_rng(0)
    mean = rng.standard_normal((32, 3, 4)).astype(np.float32)
    seg = np.repeat(np.arange(4, dtype=np.int32)[None], 32, 0).repeat(512, 1)
    out = pu.gather_operation(cu(mean), cu(seg)).cpu().numpy()
    assert np.array_equal(out, O.gather_points(mean, seg))


@pytest.mark.parametrize("B,n,m", [(3, 2048, 512), (2, 512, 128), (1, 100, 5000), (2, 7, 2), (1, 5, 3),
                                   # two-unknowns-per-thread kernel (batch x n large enough): 64- and 256-thread variants,
                                   # odd n, several known tiles, fewer than 3 known points
                                   (40, 2048, 512), (300, 1001, 77), (80, 600, 2500), (200, 300, 2)])
def test_three_nn_bit_exact(pu, ref_ext, B, n, m):
    rng = np.random.default_rng(n + m)
    unknown = part_cloud(rng, B, n, origin_frac=0)
    known = part_cloud(rng, B, m, origin_frac=0)
    if m >= 4:
        known[:, 3] = known[:, 1]  # duplicates -> ties
    dist, idx = pu.three_nn(cu(unknown), cu(known))
    od2, oidx = O.three_nn(unknown, known)
    assert np.array_equal(idx.cpu().numpy(), oidx)
    assert np.array_equal(dist.cpu().numpy(), np.sqrt(od2))  # the Python API returns sqrt(dist2)
    if "ref_pointnet2_ext" in ref_ext:
        rd2, ridx = ref_ext["ref_pointnet2_ext"].three_nn(cu(unknown), cu(known))
        assert torch.equal(idx, ridx) and torch.equal(dist, torch.sqrt(rd2))


@pytest.mark.parametrize("B,c,m,n", [(2, 256, 512, 2048), (2, 5, 30, 77), (3, 10, 64, 1000), (2, 7, 100, 1022), (40, 6, 128, 2052)])
def test_three_interpolate_bit_exact(pu, ref_ext, B, c, m, n):
    rng = np.random.default_rng(c)
    feats = rng.standard_normal((B, c, m)).astype(np.float32)
    idx = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    w = rng.random((B, n, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    f = cu(feats).requires_grad_(True)
    out = pu.three_interpolate(f, cu(idx), cu(w))
    assert np.array_equal(out.detach().cpu().numpy(), O.three_interpolate(feats, idx, w))
    go = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(cu(go))
    assert np.allclose(f.grad.cpu().numpy(), O.three_interpolate_grad(go, idx, w, m), rtol=1e-4, atol=1e-5)
    if "ref_pointnet2_ext" in ref_ext:
        assert torch.equal(out.detach(), ref_ext["ref_pointnet2_ext"].three_interpolate(cu(feats), cu(idx), cu(w)))


def test_sa_module_end_to_end_matches_reference_ops(pu, ref_ext):
    """PointnetSAModule forward (FPS -> gather -> ball_query -> group -> MLP -> maxpool) with our ops vs
    the same module graph evaluated with the reference extension's ops."""
    if "ref_pointnet2_ext" not in ref_ext:
        pytest.skip("oracle/_ref not built")
    from difffacto_b200.pointnet2_ops.pointnet2_modules import PointnetSAModule
    E = ref_ext["ref_pointnet2_ext"]
    torch.manual_seed(0)
    sa = PointnetSAModule(npoint=512, radius=0.2, nsample=64, mlp=[4, 64, 64, 128], use_xyz=True).cuda().eval()
    rng = np.random.default_rng(0)
    xyz = cu(part_cloud(rng, 4, 2048))
    feats = torch.randn(4, 4, 2048, device="cuda")
    with torch.no_grad():
        new_xyz, new_f = sa(xyz, feats)
        sel = E.furthest_point_sampling(xyz, 512)
        ref_xyz = E.gather_points(xyz.transpose(1, 2).contiguous(), sel).transpose(1, 2).contiguous()
        idx = E.ball_query(ref_xyz, xyz, 0.2, 64)
        gx = E.group_points(xyz.transpose(1, 2).contiguous(), idx) - ref_xyz.transpose(1, 2).unsqueeze(-1)
        g = torch.cat([gx, E.group_points(feats, idx)], dim=1)
        ref_f = torch.nn.functional.max_pool2d(sa.mlps[0](g), kernel_size=[1, 64]).squeeze(-1)
    assert torch.equal(new_xyz, ref_xyz) and torch.equal(new_f, ref_f)


def test_empty_inputs(pu):
    assert pu.furthest_point_sample(torch.empty(0, 16, 3, device="cuda"), 4).shape == (0, 4)
    assert pu.grouping_operation(torch.rand(2, 3, 8, device="cuda"), torch.empty(2, 0, 4, dtype=torch.int32, device="cuda")).shape == (2, 3, 0, 4)
    assert pu.ball_query(0.1, 4, torch.rand(2, 8, 3, device="cuda"), torch.empty(2, 0, 3, device="cuda")).shape == (2, 0, 4)
