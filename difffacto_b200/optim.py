"""Adam for the training hot path: every tensor of a parameter group in ONE kernel launch (csrc/train_ops.cu adam_kernel).

The reference builds `torch.optim.Adam` from `cfg.optimizer` (python/difffacto/runner/runner.py:60-66).  Same update rule here
(torch.optim.Adam, no amsgrad, L2 weight decay, bias corrections from the step count), same `state_dict()` layout per parameter
({'step', 'exp_avg', 'exp_avg_sq'}), so checkpoints move both ways; the difference is the launch count: torch's fused multi-tensor
Adam takes 8 + 3 launches (~0.25 ms) for the ~130 small tensors of the denoiser, this one takes 2 (step-count increment + update).
The step count is a device scalar shared by the group, so `step()` is capturable in a CUDA graph (train_graph.GraphedTrainStep);
`grad_scale` (a device scalar, e.g. the clipping coefficient) is folded into the update instead of a pass over the gradients.
"""
import ctypes

import torch

from . import _lib
from ._lib import check


class _AdamTensor(ctypes.Structure):  # include/difffacto_b200.h dfb200_adam_tensor_t
    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("exp_avg", ctypes.c_void_p), ("exp_avg_sq", ctypes.c_void_p),
                ("count", ctypes.c_longlong)]


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0., amsgrad=False, capturable=True):
        if amsgrad:
            raise NotImplementedError("FusedAdam: amsgrad is not implemented (use torch.optim.Adam)")
        if not 0. <= lr or not 0. <= eps or not (0. <= betas[0] < 1. and 0. <= betas[1] < 1.) or not 0. <= weight_decay:
            raise ValueError(f"FusedAdam: invalid hyper-parameters lr={lr} betas={betas} eps={eps} weight_decay={weight_decay}")
        # 'capturable' is always true here (the step count lives on the device); the key is what GraphedTrainStep checks
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, capturable=True))

    def _group_step(self, group, tensors):
        """ONE int64 device scalar per group, referenced from every parameter's state['step'] (state_dict keeps torch's layout)."""
        step = None
        for p in tensors:
            s = self.state[p].get("step")
            if s is not None:
                step = s
                break
        if step is None:
            step = torch.zeros((), dtype=torch.int64, device=tensors[0].device)
        elif not (torch.is_tensor(step) and step.dtype == torch.int64 and step.device == tensors[0].device and step.dim() == 0):
            step = torch.as_tensor(step).to(device=tensors[0].device, dtype=torch.int64).reshape(())  # loaded from a torch.optim.Adam checkpoint
        for p in tensors:
            st = self.state[p]
            st["step"] = step
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return step

    @torch.no_grad()
    def step(self, closure=None, grad_scale=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for group in self.param_groups:
            tensors = [p for p in group["params"] if p.grad is not None]
            if not tensors:
                continue
            for p in tensors:
                if p.dtype != torch.float32 or not p.is_cuda or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("FusedAdam: parameters and gradients must be dense fp32 CUDA tensors")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdam: parameters and gradients must be contiguous")
            step = self._group_step(group, tensors)
            step.add_(1)
            table = (_AdamTensor * len(tensors))()
            for e, p in zip(table, tensors):
                st = self.state[p]
                e.param, e.grad, e.exp_avg, e.exp_avg_sq, e.count = p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()
            b1, b2 = group["betas"]
            gs = None
            if grad_scale is not None:
                if grad_scale.dtype != torch.float32 or grad_scale.device != tensors[0].device:
                    raise RuntimeError("FusedAdam: grad_scale must be an fp32 scalar on the parameters' device")
                gs = grad_scale.data_ptr()
            with torch.cuda.device(tensors[0].device):
                check(lib.dfb200_adam_step(len(tensors), ctypes.cast(table, ctypes.c_void_p), step.data_ptr(), float(group["lr"]), float(b1), float(b2),
                                           float(group["eps"]), float(group["weight_decay"]), gs, _lib.stream()))
        return loss
