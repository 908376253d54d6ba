"""CUDA-graph capture of the training step (forward + backward of the whole hot path).

One SPAIR step at the default config is ~2,000 kernel launches (31 wavefronts x (cuBLAS GEMMs + head /
glimpse kernels), forward and backward) of a few microseconds each, so an eagerly launched step is bound by
the host's launch rate, not by the GPU.  ``GraphedTrainStep`` captures zero-grad + forward + backward once
(after warm-up on a side stream) and replays it: per step the host only refreshes the two step-dependent
schedules (``SPAIR.prepare_step``), copies the batch into the static input and launches one graph.  The
gradient allreduce (dp.py) and the optimiser step stay outside the graph.

Requirements: gradients live in a ``dp.GradientBucket`` (static memory, zeroed inside the graph); the batch
shape is fixed; noise is drawn on the device inside the graph (PyTorch's graph-safe Philox generator
advances per replay); the writer is not called from a captured step.
"""
from __future__ import annotations

import gc

import torch

from .dp import GradientBucket, trainable_parameters


class _PendingLoss:
    def __init__(self, host, event):
        self._host, self._event = host, event

    def value(self) -> float:
        self._event.synchronize()
        return float(self._host)


class GraphedTrainStep:
    def __init__(self, net, example_x: torch.Tensor, bucket: GradientBucket = None, global_step: int = 1000, warmup: int = 3):
        if not example_x.is_cuda:
            raise ValueError("GraphedTrainStep needs a CUDA batch")
        self.net = net
        self._copy_stream, self._staged, self._staged_ready, self._staged_free, self._has_staged = None, None, None, None, False
        self.bucket = bucket if bucket is not None else GradientBucket(trainable_parameters(net))
        self.static_x = example_x.detach().clone().float().contiguous()
        dev = example_x.device
        # drop every reference to earlier autograd graphs: their AccumulateGrad nodes are bound to the stream they
        # were created on and would drag the legacy stream into the capture
        net._latents, net.dist_param, net.dist = None, {}, {}
        gc.collect()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                      # builds the plan, cuBLAS/cuDNN handles, func attributes
                self.bucket.zero()
                net(self.static_x, global_step)[0].backward()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        # the warm-up graphs go too: the capture then creates its own AccumulateGrad nodes on the capture stream (the same
        # side stream), so the captured backward needs no cross-stream hand-over for the parameter gradients
        net._latents, net.dist_param, net.dist = None, {}, {}
        gc.collect()
        net.prepare_step(global_step, dev)
        self.graph = torch.cuda.CUDAGraph()
        net._static_step = True
        try:
            with torch.cuda.graph(self.graph, stream=side):
                self.bucket.zero()
                self.out = net(self.static_x, global_step)
                self.out[0].backward()
        finally:
            net._static_step = False
        self.bucket.check_attached()

    def prefetch(self, x_host: torch.Tensor) -> None:
        """Starts the host->device copy of the NEXT step's batch on a copy stream, so that it overlaps the step that is
        running (input double buffering: the captured graph reads ``static_x``, the copy lands in a second buffer).
        The next ``__call__(None, step)`` consumes it.  ``x_host`` should be pinned."""
        dev = self.static_x.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged = torch.empty_like(self.static_x)
            self._staged_ready = torch.cuda.Event()
            self._staged_free = torch.cuda.Event()
            self._staged_free.record(torch.cuda.current_stream(dev))
        self._copy_stream.wait_event(self._staged_free)          # the previous consumer of the staging buffer is done
        with torch.cuda.stream(self._copy_stream):
            self._staged.copy_(x_host, non_blocking=True)
            self._staged_ready.record(self._copy_stream)
        self._has_staged = True

    def loss_async(self):
        """Non-blocking read of the step that was just launched: snapshots the (static) loss output, starts its D2H copy
        into pinned memory and returns a handle whose ``value()`` waits for that copy only.  Calling ``value()`` one step
        later (after the next step has been launched) keeps the host one step ahead of the GPU, so the per-step launch work
        of the host (schedule refresh, graph launch, optimiser launch) overlaps the previous step's kernels."""
        dev = self.static_x.device
        if not hasattr(self, "_loss_ring"):
            self._loss_ring = [(torch.empty((), device=dev), torch.empty((), pin_memory=True), torch.cuda.Event()) for _ in range(4)]
            self._loss_slot = 0
        snap, host, ev = self._loss_ring[self._loss_slot]
        self._loss_slot = (self._loss_slot + 1) % len(self._loss_ring)
        snap.copy_(self.out[0].detach(), non_blocking=True)
        host.copy_(snap, non_blocking=True)
        ev.record(torch.cuda.current_stream(dev))
        return _PendingLoss(host, ev)

    def __call__(self, x, global_step: int):
        """Runs one captured step on ``x`` (device tensor, or pinned host tensor — copied asynchronously; ``None`` = the
        batch staged by ``prefetch``).  Returns the static output tuple (loss, recon_x, z_where, z_pres); gradients are
        in the bucket."""
        dev = self.static_x.device
        self.net.prepare_step(global_step, dev)
        if x is None:
            if not self._has_staged:
                raise RuntimeError("no batch staged: call prefetch(x_host) first")
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(self._staged_ready)
            self.static_x.copy_(self._staged, non_blocking=True)
            self._staged_free.record(cur)
            self._has_staged = False
        else:
            self.static_x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.out
