"""Whole-step CUDA-graph training harness with data-parallel gradient exchange.

Replaces the hot loop of the reference trainer, resnet/train.py:387-409 (`output = model(images); loss = criterion(...);
optimizer.zero_grad(); loss.backward(); optimizer.step()`) and its `DistributedDataParallel` wrap (train.py:172-174) for
static-shape training: a resnet50_mrlal step is ~800 kernel launches and the Python / launch path alone costs more than
the GPU time of the step on a B200, so the step is captured once and replayed.

    step = GraphedStep(model, optimizer, loss_fn, example_images, example_labels)     # two lines in train.py
    loss = step(images, labels)                                                       # instead of the five above

What is captured (graph A): zero the flat fp32 gradient buffer -> forward under autocast -> loss -> backward into views of
that buffer -> (with a process group) NCCL all-reduce (AVG) of the flat buffer, captured in the same graph so there is
no host round trip between backward and the exchange.  `overlap=True` instead cuts the buffer into buckets in reverse
parameter order and launches each bucket's all-reduce from a post-accumulate-grad hook on a side stream as soon as its last
gradient exists, overlapping the exchange with the rest of backward.  On B200 that is NOT the default: the tail sweeps are
persistent kernels with one CTA per SM, and a concurrent NCCL kernel takes SMs away from them — the sweep then needs a
second wave and the step gets slower than with the 0.3 ms serial all-reduce (see profiles/ for the 2-GPU A/B).
Graph B is the optimizer step.  `capture=False` runs the identical sequence eagerly (tests compare the
two).  Parameters and buffers are broadcast from rank 0 at construction, like DDP does.
"""
from __future__ import annotations

from typing import Callable, List, Optional

import torch
import torch.distributed as dist

__all__ = ["GraphedStep"]


def _grad_view(p: torch.Tensor, seg: torch.Tensor) -> torch.Tensor:
    """A view of the flat segment with the parameter's shape AND strides (channels_last conv weights keep theirs, so the
    backward kernels write into the buffer directly instead of through a layout-converting copy)."""
    if p.dim() == 4 and not p.is_contiguous() and p.is_contiguous(memory_format=torch.channels_last):
        k, c, r, s = p.shape
        return seg.view(k, r, s, c).permute(0, 3, 1, 2)
    return seg.view(p.shape)


class GraphedStep:
    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, loss_fn: Callable,
                 example_inputs: torch.Tensor, example_targets: torch.Tensor, *,
                 autocast_dtype: Optional[torch.dtype] = torch.bfloat16, process_group=None, bucket_mb: float = 25.0,
                 warmup: int = 3, capture: bool = True, broadcast: bool = True, overlap: bool = False,
                 capture_collective: bool = False):
        if not example_inputs.is_cuda:
            raise RuntimeError("GraphedStep: inputs must live on a CUDA device")
        self.model, self.opt, self.loss_fn = model, optimizer, loss_fn
        self.autocast_dtype = autocast_dtype
        self.dev = example_inputs.device
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or
                                                            (dist.is_available() and dist.is_initialized())) else 1
        if self.world > 1 and self.pg is None:
            self.pg = dist.group.WORLD
        self.inputs = example_inputs.clone()
        self.targets = example_targets.clone()
        self.params: List[torch.nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.dev)
        for p in self.params:
            if p.dtype != torch.float32:
                raise RuntimeError("GraphedStep keeps fp32 master weights / gradients (use autocast for bf16 compute)")
        # gradient views + buckets (reverse registration order ~ order in which backward produces them)
        off = 0
        self._range = {}
        for p in self.params:
            p.grad = _grad_view(p, self.flat[off:off + p.numel()])
            self._range[p] = (off, off + p.numel())
            off += p.numel()
        self.buckets = []          # (start, end, [params])
        self.overlap = bool(overlap) and self.world > 1
        # the flat all-reduce either sits between the two graphs (default, one eager NCCL call per step) or inside graph A
        self.capture_collective = bool(capture_collective) or self.overlap
        if self.world > 1 and not self.overlap:
            self.buckets = [(0, n, list(self.params))]
            if broadcast:
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, 0, group=self.pg)
        if self.overlap:
            cap = int(bucket_mb * (1 << 20) / 4)
            cur, cur_n = [], 0
            for p in reversed(self.params):
                cur.append(p)
                cur_n += p.numel()
                if cur_n >= cap:
                    self.buckets.append(cur)
                    cur, cur_n = [], 0
            if cur:
                self.buckets.append(cur)
            self.buckets = [(min(self._range[p][0] for p in b), max(self._range[p][1] for p in b), b) for b in self.buckets]
            self._bucket_of = {p: i for i, (_, _, b) in enumerate(self.buckets) for p in b}
            self._pending = [0] * len(self.buckets)
            self.comm_stream = torch.cuda.Stream(device=self.dev)
            self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
            if broadcast:
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, 0, group=self.pg)
        self.collective_launches = 0
        self.graph_a = self.graph_b = None
        self.loss = None
        self._armed = False
        self._capturing = False
        if capture:
            self._capture(warmup)

    # ------------------------------------------------------------------------------------------ pieces of a step
    def _arm(self):
        if self.overlap:
            self._pending = [len(b) for (_, _, b) in self.buckets]
            self._armed = True

    def _on_grad(self, p):
        """post-accumulate-grad hook: the bucket's last gradient is in the flat buffer -> all-reduce it on the side stream."""
        if not self._armed:
            return
        i = self._bucket_of[p]
        self._pending[i] -= 1
        if self._pending[i] == 0:
            s, e, _ = self.buckets[i]
            main = torch.cuda.current_stream(self.dev)
            self.comm_stream.wait_stream(main)
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.AVG, group=self.pg)
            self.collective_launches += 1

    def _fwd_bwd(self):
        self.flat.zero_()
        self._arm()
        if self.autocast_dtype is not None:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                out = self.model(self.inputs)
        else:
            out = self.model(self.inputs)
        loss = self.loss_fn(out.float(), self.targets)
        loss.backward()
        if self.world > 1 and not self.overlap and (self.capture_collective or self.graph_a is None and not self._capturing):
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.pg)
            self.collective_launches += 1
        if self.overlap:
            self._armed = False
            if any(n_ != 0 for n_ in self._pending):   # a parameter received no gradient this step: reduce what is left
                for i, n_ in enumerate(self._pending):
                    if n_ != 0:
                        s, e, _ = self.buckets[i]
                        self.comm_stream.wait_stream(torch.cuda.current_stream(self.dev))
                        with torch.cuda.stream(self.comm_stream):
                            dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.AVG, group=self.pg)
            torch.cuda.current_stream(self.dev).wait_stream(self.comm_stream)   # join: gradients are averaged
        return loss

    def _capture(self, warmup: int):
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):   # the exact captured sequence, on the capture stream
                self._fwd_bwd()
                self.opt.step()
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        self.collective_launches = 0
        self.graph_a, self.graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        # with NCCL in the process, its watchdog thread polls CUDA events while this thread captures: thread-local capture
        # mode keeps those (legal) calls from invalidating the capture
        mode = "thread_local" if self.world > 1 else "global"
        self._capturing = True
        with torch.cuda.graph(self.graph_a, capture_error_mode=mode):
            self.loss = self._fwd_bwd()
        self._capturing = False
        with torch.cuda.graph(self.graph_b, pool=self.graph_a.pool(), capture_error_mode=mode):
            self.opt.step()
        torch.cuda.synchronize(self.dev)

    # ------------------------------------------------------------------------------------------ public
    def load(self, inputs: Optional[torch.Tensor] = None, targets: Optional[torch.Tensor] = None):
        """Copy a batch into the static buffers the graphs read (device-to-device or pinned-host-to-device)."""
        if inputs is not None:
            self.inputs.copy_(inputs, non_blocking=True)
        if targets is not None:
            self.targets.copy_(targets, non_blocking=True)

    def __call__(self, inputs: Optional[torch.Tensor] = None, targets: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One training step; returns the (device) loss tensor of this step."""
        self.load(inputs, targets)
        if self.graph_a is not None:
            self.graph_a.replay()
            if self.world > 1 and not self.capture_collective:
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.pg)
            self.graph_b.replay()
            return self.loss
        loss = self._fwd_bwd()
        self.opt.step()
        return loss

    step = __call__

    def close(self):
        """Drop the hooks and the captured graphs (release them before `destroy_process_group()`: a live graph keeps a
        reference to the NCCL communicator it was captured on)."""
        for h in getattr(self, "_hooks", []):
            h.remove()
        self._hooks = []
        torch.cuda.synchronize(self.dev)
        for g in (self.graph_a, self.graph_b):
            if g is not None:
                g.reset()
        self.graph_a = self.graph_b = None
