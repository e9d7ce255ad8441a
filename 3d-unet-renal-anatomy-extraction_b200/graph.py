"""Whole-step CUDA graph: zero_grad + forward + loss + backward [+ gradient all-reduce] + optimizer step captured once
and replayed, so the ~1 600 kernel launches of a training step cost no host time (Python, ctypes, TMA descriptor
encoding all happen at capture time only).

The reference's step loop (trainer.py:473-509) is eager; this is an optional accelerator for the same loop:
``Trainer(..., cuda_graph=True)`` or directly

    step = GraphedTrainStep(model, loss_fn, optimizer)
    loss = step(image, label)          # device tensors or pinned host tensors of a fixed shape

The first ``warmup`` calls run eagerly (they are real training steps; they also let the caching allocator and the
plans settle), the next call captures, every later call copies the batch into the static buffers and replays.
A batch of a different shape falls back to the eager path.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import ops
from . import parallel


def _make_capturable(optimizer):
    for group in optimizer.param_groups:
        if "capturable" in group:
            group["capturable"] = True


class GraphedTrainStep:
    def __init__(self, model, loss_fn, optimizer, warmup: int = 3, allreduce: bool = True, capture_collectives: bool = True,
                 overlap: bool = True):
        """capture_collectives (data parallel only): capture the NCCL gradient all-reduce INSIDE the step's graph (one
        graph per step, no host round trip between backward and optimizer); with `overlap` the first ~85 % of the
        gradient bytes (head, decoder, bottom level) are all-reduced on NCCL's stream while the encoder half of the
        backward pass is still running, and 1 / world is folded into the gradient unpack.  False = round-1 behaviour:
        two graphs with an eager all-reduce between them."""
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.warmup, self.allreduce = warmup, allreduce
        self.capture_collectives, self.overlap = capture_collectives, overlap
        self.mode = None                # "single" | "one-graph-dp" | "two-graph-dp" once captured
        self.calls = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_image = self.static_label = self.static_loss = self.static_logits = None
        self.key = None
        self._side = None
        self.graph_opt = None
        self.launches_per_replay = 0
        self.replayed = False           # whether the most recent call replayed the graph (False: it ran eagerly)
        self.last_label = None          # the label tensor the most recent call computed its loss against
        self._static_grads, self._static_flats = [], []
        _make_capturable(optimizer)

    def _fwd_bwd(self, image, label):
        self.model.train()
        self.optimizer.zero_grad(set_to_none=True)
        logits = self.model(image)
        loss = self.loss_fn(logits, label)
        loss.backward()
        return loss, logits

    def _eager(self, image, label):
        loss, logits = self._fwd_bwd(image, label)
        if self.allreduce:
            parallel.all_reduce_gradients(self.model)
        self.optimizer.step()
        return loss, logits

    def __call__(self, image: torch.Tensor, label: torch.Tensor):
        """Returns (loss, logits) -- device tensors; under replay they are the graph's static outputs.

        Single process: ONE graph holds the whole step.  Data parallel (world > 1 and allreduce): TWO graphs --
        zero_grad + forward + loss + backward, then the optimizer step -- with the NCCL gradient all-reduce issued
        eagerly between them on the static gradient buffers (no collective is captured)."""
        dev = next(self.model.parameters()).device
        key = (tuple(image.shape), tuple(label.shape), image.dtype, label.dtype, self.model.training)
        self.calls += 1
        world = parallel.rank_world()[1]
        split = world > 1 and self.allreduce
        if split and self.calls == 2 and self.capture_collectives:
            # engines exist after the first step: average inside the unpack, send the first chunk during backward
            parallel.prescale_gradients(self.model)
            if self.overlap:
                parallel.overlap_gradient_all_reduce(self.model)
        self.replayed = False
        if self.key is not None and key != self.key:
            # another batch shape (a short last batch) or mode (first train batch after a validation pass): eager, and
            # the caller must compare the logits with THIS label tensor, not with the graph's static one
            self.last_label = label.to(dev, non_blocking=True)
            return self._eager(image.to(dev, non_blocking=True), self.last_label)
        if self.calls <= self.warmup:
            # warm-up steps run on a side stream: autograd's AccumulateGrad nodes are then not tied to the legacy
            # default stream, which a capturing stream may not synchronise with (cudaErrorStreamCaptureImplicit)
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            self._side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(self._side):
                self.last_label = label.to(dev, non_blocking=True)
                out = self._eager(image.to(dev, non_blocking=True), self.last_label)
            torch.cuda.current_stream(dev).wait_stream(self._side)
            return out
        if self.graph is None:
            self.key = key
            self.static_image = torch.empty(image.shape, dtype=image.dtype, device=dev)
            self.static_label = torch.empty(label.shape, dtype=label.dtype, device=dev)
            self.static_image.copy_(image, non_blocking=True)
            self.static_label.copy_(label, non_blocking=True)
            torch.cuda.synchronize()
            ops.check_device_errors()
            launches0 = ops.LAUNCHES
            g = torch.cuda.CUDAGraph()
            self.optimizer.zero_grad(set_to_none=True)
            one_graph = split and self.capture_collectives
            if one_graph:
                try:
                    with torch.cuda.graph(g):
                        self.static_loss, self.static_logits = self._eager(self.static_image, self.static_label)
                    self.mode = "one-graph-dp"
                except Exception as e:          # NCCL capture refused: fall back to two graphs around an eager all-reduce
                    import warnings
                    warnings.warn(f"GraphedTrainStep: capturing the gradient all-reduce failed ({e!r}); using two graphs")
                    torch.cuda.synchronize()
                    one_graph = False
                    parallel.overlap_gradient_all_reduce(self.model, enable=False)
                    g = torch.cuda.CUDAGraph()
                    self.optimizer.zero_grad(set_to_none=True)
            if one_graph:
                split = False
            else:
                with torch.cuda.graph(g):
                    if split:
                        self.static_loss, self.static_logits = self._fwd_bwd(self.static_image, self.static_label)
                        self.mode = "two-graph-dp"
                    else:
                        self.static_loss, self.static_logits = self._eager(self.static_image, self.static_label)
                        self.mode = "single"
            self.graph = g
            if split:
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=g.pool()):
                    self.optimizer.step()
                self.graph_opt = g2
            self.launches_per_replay = ops.LAUNCHES - launches0      # this library's kernels inside the graph(s)
            # the gradient tensors the graph writes (and the engine's flat buffers behind them): re-attached before
            # every replay, so that an eager step in between cannot leave .grad / the all-reduce pointing elsewhere
            self._static_grads = [(p, p.grad) for p in self.model.parameters()]
            self._static_flats = [(eng, eng._last_gflat) for eng in parallel._engines(self.model)]
        else:
            self.static_image.copy_(image, non_blocking=True)
            self.static_label.copy_(label, non_blocking=True)
        for p, g in self._static_grads:
            p.grad = g
        for eng, flats in self._static_flats:
            eng._last_gflat = flats
        self.replayed, self.last_label = True, self.static_label
        self.graph.replay()
        if self.graph_opt is not None:
            parallel.all_reduce_gradients(self.model)
            self.graph_opt.replay()
        return self.static_loss, self.static_logits
