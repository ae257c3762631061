"""CPU guard for the GELU forms of the CUDA kernels.  The reference's GEGLU is `x * F.gelu(gate)` with the exact erf GELU
(python/difffacto/models/diffusions/nets/attention.py:50-57).  The kernels use three cheaper forms whose deviation from the erf GELU is
part of each precision mode's stated tolerance:
  bf16 mode  (csrc/denoiser_tc.cu)    0.5 g (1 + tanh(g (c0 + c1 g^2))), (c0, c1) refit        max |dev| 2.7e-4
  tf32 mode  (csrc/denoiser_tf32.cu)  g / (1 + 2^(g (k0 + k1 g^2 + k2 g^4 + k3 g^6)))           max |dev| 2.7e-5
  training   (csrc/geglu_math.cuh)    Abramowitz-Stegun 7.1.26 erf from one exponential          |cdf dev| 3e-7 (fp32)
This test reads the constants OUT OF THE SOURCES, restates each form in numpy float32 and checks the documented bound on a dense grid,
so an edit of a constant that breaks the bound fails here, without a GPU."""
import math
import os
import re

import numpy as np

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "difffacto_b200", "csrc")
X = np.concatenate([np.linspace(-12.0, 12.0, 480001), np.array([-40.0, -20.0, 20.0, 40.0, 0.0])]).astype(np.float32)
_erf = np.vectorize(math.erf)


def _gelu_exact(x):
    x = x.astype(np.float64)
    return x * 0.5 * (1.0 + _erf(x / math.sqrt(2.0)))


def _floats(text, pattern, n):
    m = re.search(pattern, text, re.S)
    assert m, f"pattern not found: {pattern}"
    vals = [float(v.rstrip("f")) for v in m.groups()]
    assert len(vals) == n
    return vals


def test_bf16_mode_tanh_form_constants():
    src = open(os.path.join(CSRC, "denoiser_tc.cu")).read()
    c1, c0 = _floats(src, r"__ffma2_rn\(g2, f2s\(([0-9.]+)f\), f2s\(([0-9.]+)f\)\)", 2)  # in = g * (g^2 c1 + c0)
    g = X.astype(np.float32)
    inner = g * (np.float32(c0) + np.float32(c1) * g * g)
    gelu = 0.5 * g.astype(np.float64) * (1.0 + np.tanh(inner.astype(np.float64)))
    dev = np.abs(gelu - _gelu_exact(X)).max()
    assert dev < 3.0e-4, dev          # documented: 2.7e-4, below the bf16 rounding of the activations (tanh.approx adds ~5e-4 relative)
    assert abs(c0 - math.sqrt(2.0 / math.pi)) < 0.01 and 0.03 < c1 < 0.04  # a refit of the textbook (0.79788, 0.035677), not something else


def test_tf32_mode_sigmoid_polynomial_constants():
    src = open(os.path.join(CSRC, "denoiser_tf32.cu")).read()
    K, = _floats(src, r"constexpr float K = ([-0-9.e]+f?);", 1)
    k = _floats(src, r"k0 = K \* ([-0-9.e]+f?), k1 = K \* ([-0-9.e]+f?), k2 = K \* ([-0-9.e]+f?),\s*k3 = K \* ([-0-9.e]+f?);", 4)
    assert abs(K + 2.0 / math.log(2.0)) < 1e-6                      # exp(-2 q) as 2^(K q)
    g = X.astype(np.float32)
    k32 = [np.float32(K) * np.float32(v) for v in k]
    s = g * g
    p = k32[2] + k32[3] * s
    p = k32[1] + p * s
    p = k32[0] + p * s
    with np.errstate(over="ignore"):  # 2^arg -> inf for very negative g: the kernel's rcp(inf) = 0, here g / inf = -0
        e = np.exp2((g * p).astype(np.float64))
        gelu = g.astype(np.float64) / (1.0 + e)
    dev = np.abs(gelu - _gelu_exact(X)).max()
    assert dev < 4.0e-5, dev          # documented: 2.7e-5, a tenth of the tf32 rounding of the result
    # the inner polynomial keeps the sign of g for every input (the form is sign-safe: Phi -> 0 / 1 at -inf / +inf)
    inner = np.float64(k[0]) + np.float64(k[1]) * 1600.0 + np.float64(k[2]) * 1600.0 ** 2 + np.float64(k[3]) * 1600.0 ** 3
    assert inner > 0.5 and gelu[X == -40.0][0] == 0.0 and gelu[X == 40.0][0] == 40.0


def test_training_path_abramowitz_stegun_cdf_pdf():
    src = open(os.path.join(CSRC, "geglu_math.cuh")).read()
    pz, = _floats(src, r"fmaf\(([0-9.]+)f, z, 1\.f\)", 1)
    a5, a4 = _floats(src, r"float p = fmaf\(([-0-9.]+)f, t, ([-0-9.]+)f\);", 2)
    a3, = _floats(src, r"p = fmaf\(p, t, (1\.[0-9]+)f\);", 1)
    a2, = _floats(src, r"p = fmaf\(p, t, (-0\.2[0-9]+)f\);", 1)
    a1, = _floats(src, r"p = fmaf\(p, t, (0\.2[0-9]+)f\);", 1)
    f = np.float32
    g = X.astype(np.float32)
    z = np.abs(g) * f(0.70710678118654752440)
    t = f(1.0) / (f(pz) * z + f(1.0))
    e = np.exp(-(z * z)).astype(np.float32)
    p = f(a5) * t + f(a4)
    p = p * t + f(a3)
    p = p * t + f(a2)
    p = p * t + f(a1)
    erf_abs = f(1.0) - p * t * e
    cdf = f(0.5) * (f(1.0) + np.copysign(erf_abs, g))
    pdf = f(0.39894228040143267794) * e
    x64 = X.astype(np.float64)
    cdf_exact = 0.5 * (1.0 + _erf(x64 / math.sqrt(2.0)))
    pdf_exact = np.exp(-0.5 * x64 * x64) / math.sqrt(2.0 * math.pi)
    assert np.abs(cdf - cdf_exact).max() < 3.5e-7      # A-S 7.1.26: |erf error| <= 1.5e-7, halved for the CDF, plus fp32 rounding
    assert np.abs(pdf - pdf_exact).max() < 1.0e-7
    # GEGLU forward g * cdf and its derivative cdf + g * pdf, as the kernels combine them
    assert np.abs(g * cdf - _gelu_exact(X)).max() < 2.0e-6
    assert np.abs((cdf + g * pdf) - (cdf_exact + x64 * pdf_exact)).max() < 2.0e-6
