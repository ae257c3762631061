"""ctypes binding of the C ABI in include/difffacto_b200.h.

The product path fails loudly if the CUDA library is missing: there is no eager / CPU fallback.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libdifffacto_b200.so")
DIAG_LIB_PATH = os.path.join(_PKG, "lib", "libdifffacto_b200_diag.so")  # -DDFB200_DIAGNOSTICS build (tests / tools only)

c_int, c_float, c_void_p, c_size_t, c_u64 = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64


class DenoiserCfg(ctypes.Structure):
    """struct dfb200_denoiser_cfg"""
    _fields_ = [("in_channels", c_int), ("out_channels", c_int), ("n_heads", c_int), ("d_head", c_int),
                ("depth", c_int), ("context_dim", c_int), ("n_class", c_int), ("flags", c_int)]


class SampleOpts(ctypes.Structure):
    """struct dfb200_sample_opts"""
    _fields_ = [("timesteps", c_void_p), ("timesteps_host", ctypes.POINTER(c_int)), ("n_timesteps", c_int),
                ("first_step", c_int), ("num_steps", c_int), ("tables_ready", c_int), ("ddim", c_int), ("ddim_eta", c_float),
                ("alphas_cumprod_prev", c_void_p), ("xt_dir_coeff", c_void_p), ("guidance", c_int),
                ("classifier_weight", c_float), ("step_sample", c_void_p), ("step_xstart", c_void_p),
                ("step_sample_list", ctypes.POINTER(c_void_p)), ("step_xstart_list", ctypes.POINTER(c_void_p))]


NET_CLASS_COND, NET_CAT_PARAMS_TO_X, NET_CAT_CLASS_TO_X, NET_MASK_UNREFERENCED, NET_INCLUDE_STD = 1, 2, 4, 8, 16
MODE_FP32, MODE_BF16, MODE_TF32 = 0, 1, 2
SCHED_ROWS = 8
P = c_void_p
_CFG = ctypes.POINTER(DenoiserCfg)

# name -> (restype, argtypes); must list every symbol declared in include/difffacto_b200.h
SIGNATURES = {
    "dfb200_abi_version": (c_int, []),
    "dfb200_last_error": (ctypes.c_char_p, []),
    "dfb200_launch_count": (ctypes.c_ulonglong, []),
    "dfb200_gather_points": (c_int, [c_int] * 4 + [P, P, P, P]),
    "dfb200_gather_points_grad": (c_int, [c_int] * 4 + [P, P, P, P]),
    "dfb200_furthest_point_sampling": (c_int, [c_int] * 3 + [P, P, P, P]),
    "dfb200_query_ball_point": (c_int, [c_int] * 3 + [c_float, c_int, P, P, P, P]),
    "dfb200_group_points": (c_int, [c_int] * 5 + [P, P, P, P]),
    "dfb200_group_points_grad": (c_int, [c_int] * 5 + [P, P, P, P]),
    "dfb200_three_nn": (c_int, [c_int] * 3 + [P, P, P, P, P]),
    "dfb200_three_interpolate": (c_int, [c_int] * 4 + [P, P, P, P, P]),
    "dfb200_three_interpolate_grad": (c_int, [c_int] * 4 + [P, P, P, P, P]),
    "dfb200_chamfer_forward": (c_int, [c_int, c_int, P, c_int, P, P, P, P, P, P]),
    "dfb200_chamfer_backward": (c_int, [c_int, c_int, P, c_int, P, P, P, P, P, P, P, P]),
    "dfb200_emd_forward": (c_int, [c_int, c_int] + [P] * 14 + [c_float, c_int, P]),
    "dfb200_emd_backward": (c_int, [c_int, c_int, P, P, P, P, P, P]),
    "dfb200_denoiser_num_params": (c_int, [_CFG]),
    "dfb200_denoiser_packed_bytes": (c_size_t, [_CFG]),
    "dfb200_denoiser_pack": (c_int, [_CFG, ctypes.POINTER(P), c_int, P, P]),
    "dfb200_denoiser_workspace_bytes": (c_size_t, [_CFG, c_int, c_int, c_int]),
    "dfb200_denoiser_forward": (c_int, [_CFG, P, c_int, c_int, c_int] + [P] * 8 + [P, c_size_t, P]),
    "dfb200_ddpm_step": (c_int, [c_int] * 3 + [P] * 10),
    "dfb200_q_sample": (c_int, [c_int] * 3 + [P] * 8),
    "dfb200_sgemm": (c_int, [c_int] * 5 + [P, c_int, P, c_int, P, c_int, P, c_int, c_int, P]),
    "dfb200_gemm_bf16": (c_int, [c_int] * 5 + [P, c_int, P, c_int, P, c_int, P, c_int, c_int, P]),
    "dfb200_colsum_accumulate": (c_int, [ctypes.c_longlong, c_int, P, c_int, P, P]),
    "dfb200_layernorm128_forward": (c_int, [ctypes.c_longlong] + [P] * 7),
    "dfb200_layernorm128_backward": (c_int, [ctypes.c_longlong] + [P] * 9),
    "dfb200_layernorm128_backward_residual": (c_int, [ctypes.c_longlong] + [P] * 10),
    "dfb200_geglu_forward": (c_int, [ctypes.c_longlong, c_int, P, P, P]),
    "dfb200_geglu_backward": (c_int, [ctypes.c_longlong, c_int, P, P, P, P]),
    "dfb200_part_attention_forward": (c_int, [c_int, c_int] + [P] * 7),
    "dfb200_part_attention_backward": (c_int, [c_int, c_int] + [P] * 10),
    "dfb200_timestep_embedding": (c_int, [c_int, P, P, P, P]),
    "dfb200_geglu_dropout_forward": (c_int, [ctypes.c_longlong, c_int, c_float, c_u64, c_u64, P, P, P, P]),
    "dfb200_geglu_dropout_backward": (c_int, [ctypes.c_longlong, c_int, c_float, c_u64, c_u64, P, P, P, P, P, P]),
    "dfb200_dropout": (c_int, [c_size_t, c_float, c_u64, c_u64, P, P, P, P]),
    "dfb200_dropout_stepped": (c_int, [c_size_t, c_float, c_u64, c_u64, P, P, P, P, P]),
    "dfb200_ff_in_forward": (c_int, [ctypes.c_longlong, c_int, c_int, P, c_int, P, c_int, P, c_float, c_u64, c_u64, P, P, P, P]),
    "dfb200_adam_step": (c_int, [c_int, P, P] + [c_float] * 5 + [P, P]),
    "dfb200_q_sample_backward": (c_int, [c_int] * 3 + [P] * 9),
    "dfb200_layernorm_forward": (c_int, [ctypes.c_longlong, c_int, P, P, P, P, P]),
    "dfb200_relu": (c_int, [c_size_t, P, P]),
    "dfb200_scale": (c_int, [c_size_t, c_float, P, P, P]),
    "dfb200_exp_shift": (c_int, [c_size_t, c_float, P, P, P]),
    "dfb200_coupling_reverse": (c_int, [c_int, c_int, P, P, c_int, P]),
    "dfb200_token_attention": (c_int, [c_int] * 4 + [P] * 6),
    "dfb200_batchnorm_forward": (c_int, [ctypes.c_longlong, c_int, c_int] + [P] * 8 + [c_float, P, P]),
    "dfb200_batchnorm_apply": (c_int, [ctypes.c_longlong, c_int, c_int] + [P] * 7),
    "dfb200_batchnorm_backward": (c_int, [ctypes.c_longlong, c_int, c_int] + [P] * 10),
    "dfb200_relu_backward": (c_int, [c_size_t, P, P, P, P]),
    "dfb200_weighted_maxpool_forward": (c_int, [c_int] * 4 + [c_float, P, P, P, P, P]),
    "dfb200_weighted_maxpool_backward": (c_int, [c_int] * 4 + [c_float, P, P, P, P, P]),
    "dfb200_coupling_forward": (c_int, [c_int, c_int, P, P, c_int, P, c_int, P, P]),
    "dfb200_coupling_backward": (c_int, [c_int, c_int, P, P, c_int, P, c_int, P, P, P, c_int, P]),
    "dfb200_ddim_step": (c_int, [c_int] * 3 + [P] * 9 + [c_float, P, P, P]),
    "dfb200_guidance_mix": (c_int, [c_size_t, c_float, P, P, P, P]),
    "dfb200_philox_normal": (c_int, [P, c_size_t, c_u64, c_u64, P]),
    "dfb200_ddpm_sample_loop_workspace_bytes": (c_size_t, [_CFG, c_int, c_int, c_int, c_int]),
    "dfb200_ddpm_sample_loop": (c_int, [_CFG, P, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P, P, P, P, c_u64, P,
                                         c_int, P, c_size_t, P]),
    "dfb200_sample_loop_chunk": (c_int, [_CFG, c_int, c_int, c_int, c_int]),
    "dfb200_sample_loop": (c_int, [_CFG, P, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P, P, P, P, c_u64, P,
                                    c_int, ctypes.POINTER(SampleOpts), P, c_size_t, P]),
}

# include/difffacto_b200_diag.h: exported only by the diagnostic build
DIAG_SIGNATURES = {
    "dfb200_bench_umma": (c_int, [c_int, c_int, c_int, c_int, P, P]),
    "dfb200_bench_umma2": (c_int, [c_int, c_int, c_int, c_int, P, P]),
    "dfb200_debug_tc_timeline": (c_int, [P, c_int]),
    "dfb200_selftest_umma": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "dfb200_selftest_umma2": (c_int, [c_int, c_int, c_int, P, P, P, P, P, P, P]),
}

_lib = None
_diag = None


def _open(path, signatures, how):
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `{how}` "
                          "(difffacto_b200 has no CPU or eager-PyTorch fallback)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in signatures.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def load():
    """dlopen the in-tree library and type every entry point (no CUDA call is made here).
    DFB200_DIAGNOSTICS=1 in the environment makes the whole package run on the diagnostic build (tools/diag_*.py)."""
    global _lib
    if _lib is None:
        # DFB200_LIB_PATH: A/B measurements of two builds of THIS library on one box (tools/); never a fallback
        _lib = load_diag() if os.environ.get("DFB200_DIAGNOSTICS") == "1" else \
            _open(os.environ.get("DFB200_LIB_PATH") or LIB_PATH, SIGNATURES, "python -m difffacto_b200.build")
    return _lib


def load_diag():
    """The diagnostic build (all product entry points + include/difffacto_b200_diag.h); tests and tools only."""
    global _diag
    if _diag is None:
        _diag = _open(DIAG_LIB_PATH, {**SIGNATURES, **DIAG_SIGNATURES}, "python -m difffacto_b200.build --diag")
    return _diag


class DFB200Error(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        msg = load().dfb200_last_error().decode(errors="replace")
        raise DFB200Error(f"difffacto_b200 status {rc}: {msg}")


def ptr(t):
    """Raw device pointer of a tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


_RAW_STREAM = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_GET_DEVICE = getattr(torch._C, "_cuda_getDevice", None)


def stream():
    """Current CUDA stream of the current device as a raw handle.  torch.cuda.current_stream() builds a Stream object
    through several Python layers (about 10 % of the host time of a training step, which is launch-bound); the C binding
    returns the same handle directly."""
    if _RAW_STREAM is not None and _GET_DEVICE is not None:
        return c_void_p(_RAW_STREAM(_GET_DEVICE()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)  # public API, if a torch build lacks the bindings


class on:
    """`with on(tensor.device):` - device guard that is free when the tensor already lives on the current device (the
    common case; torch.cuda.device() costs a few microseconds per launch otherwise)."""
    __slots__ = ("guard",)

    def __init__(self, device):
        if isinstance(device, str):  # 'cuda' / 'cuda:1', as AnchorDiffAE.decode passes by default
            device = torch.device(device)
        idx = device.index if isinstance(device, torch.device) else int(device)
        if idx is None:  # 'cuda' without an index = the current device; CPU: nothing to guard (callers reject CPU tensors)
            self.guard = None
        else:
            cur = _GET_DEVICE() if _GET_DEVICE is not None else torch.cuda.current_device()
            self.guard = None if idx == cur else torch.cuda.device(idx)

    def __enter__(self):
        if self.guard is not None:
            self.guard.__enter__()
        return self

    def __exit__(self, *a):
        if self.guard is not None:
            return self.guard.__exit__(*a)
        return False


def require_cuda(*tensors):
    # same contract as the reference's CHECK_CUDA / "CPU not supported" (ball_query.cpp:28)
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("CPU not supported: difffacto_b200 ops take CUDA tensors")


def require(t, dtype, name):
    # reference: CHECK_CONTIGUOUS / CHECK_IS_FLOAT / CHECK_IS_INT (include/utils.h:5-25)
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be a {'float' if dtype == torch.float32 else 'int'} tensor")


def launch_count():
    return int(load().dfb200_launch_count())
