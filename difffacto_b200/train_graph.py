"""One training step (forward + backward + optimizer) of the diffusion denoiser as ONE CUDA graph.

The training path launches ~260 small kernels per step from Python (autograd Functions over the C ABI); at 16 shapes x 2048
points per GPU the host needs ~6 ms to issue what the device runs in less, so the step is launch-bound.  `GraphedTrainStep`
captures the whole step once (torch.cuda.CUDAGraph, whole-network capture as in the PyTorch CUDA-graphs notes) and replays it:
inputs are copied into static buffers, the loss comes back in a static tensor.

What makes the step capturable:
  * every kernel goes to torch's CURRENT stream (`_lib.stream()`), which is the capture stream inside `torch.cuda.graph`;
  * dropout masks are keyed by a DEVICE-resident step counter (`train_ops.set_step_counter`, dfb200_dropout_stepped) that the
    graph increments once per replay - no host-side seed draw per call;
  * the optimizer runs in capturable mode (torch.optim.Adam(capturable=True): step count and learning rate live on the device);
  * `DistributedDataParallel` models are captured with their bucketed NCCL all-reduces inside the graph (NCCL >= 2.9.6); the
    warm-up runs 11 eager iterations first, as DDP needs before a capture.
Reference loop this replaces: python/difffacto/runner/runner.py:299-349 (zero_grad / forward / backward / clip / step).
"""
import torch

from . import train_ops as T


class GraphedTrainStep:
    def __init__(self, loss_fn, params, optimizer, example_inputs, max_norm=None, warmup=3, seed=0):
        """loss_fn(**inputs) -> scalar loss tensor (builds the autograd graph); params: the tensors `optimizer` updates;
        example_inputs: dict of CUDA tensors fixing shapes / dtypes of every later call; max_norm: optional gradient clipping
        (torch.nn.utils.clip_grad_norm_, reference runner.py:335-336)."""
        dev = next(iter(example_inputs.values())).device
        self.static = {k: v.clone() for k, v in example_inputs.items()}
        self.params = list(params)
        self.optimizer = optimizer
        self.counter = torch.zeros(1, dtype=torch.int64, device=dev)
        self._loss_fn, self._max_norm, self._seed = loss_fn, max_norm, int(seed)
        for g in optimizer.param_groups:
            if "capturable" in g and not g["capturable"]:
                raise ValueError("GraphedTrainStep needs an optimizer in capturable mode (e.g. torch.optim.Adam(..., capturable=True))")
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        T.set_step_counter(self.counter, self._seed)
        try:
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    self._eager_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(self.graph):
                self.loss = self._eager_step(zero=False)
        finally:
            T.set_step_counter(None)

    def _eager_step(self, zero=True):
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self._loss_fn(**self.static)
        loss.backward()
        if self._max_norm:
            torch.nn.utils.clip_grad_norm_(self.params, self._max_norm)
        self.optimizer.step()
        self.counter.add_(1)
        return loss.detach()

    def __call__(self, **inputs):
        """Copies `inputs` into the static buffers, replays the step, returns the (static) loss tensor of this step."""
        for k, v in inputs.items():
            self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.loss
