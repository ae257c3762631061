"""-m gpu: the tcgen05 building blocks (dfb200_selftest_umma, diagnostic build of the library) against a bf16-operand
fp32-accumulate reference."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("variant", [0, 2, 4, 6])
@pytest.mark.parametrize("N,K", [(128, 128), (64, 128), (128, 32), (32, 16)])
def test_umma_selftest(variant, N, K):
    from tools.diag_umma import ref
    from difffacto_b200 import _lib
    torch.manual_seed(N + K + variant)
    A = torch.randn(128, K, device="cuda")
    W = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    Cin = torch.randn(128, N, device="cuda")
    D = torch.empty(128, N, device="cuda")
    scratch = torch.zeros(N * K * 2 + 256, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load_diag().dfb200_selftest_umma(variant, N, K, _lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(Cin),
                                                _lib.ptr(D), _lib.ptr(scratch), _lib.stream()))
    torch.cuda.synchronize()
    assert (D - ref(A, W, bias, Cin)).abs().max().item() < 2e-4 * K ** 0.5 + 1e-4


@pytest.mark.parametrize("variant", [0, 4, 8, 12])
@pytest.mark.parametrize("N,K", [(128, 128), (128, 64), (64, 128), (32, 32)])
def test_umma_cta_pair_selftest(variant, N, K):
    """cta_group::2 (M = 256 across a 2-CTA cluster): split-B tiles, remote mbarrier arrives, multicast commit."""
    from tools.diag_umma import ref
    from difffacto_b200 import _lib
    torch.manual_seed(N + K + variant)
    A = torch.randn(256, K, device="cuda")
    W = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda")
    Cin = torch.randn(256, N, device="cuda")
    D = torch.empty(256, N, device="cuda")
    scratch = torch.zeros(N * K * 2 + 256, dtype=torch.uint8, device="cuda")
    _lib.check(_lib.load_diag().dfb200_selftest_umma2(variant, N, K, _lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(Cin),
                                                 _lib.ptr(D), _lib.ptr(scratch), _lib.stream()))
    torch.cuda.synchronize()
    assert (D - ref(A, W, bias, Cin)).abs().max().item() < 2e-4 * K ** 0.5 + 1e-4
