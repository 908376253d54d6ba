"""Data-parallel training over images: one process per GPU, NCCL allreduce of the gradients only.

The reference has no distributed code at all (``train.py:27-30`` picks one device).
The path shards naturally by image — every per-cell computation, the KL scan and the renderer are
independent across the batch index — so rank r of R takes B/R images and the only exchange per
step is ONE ``all_reduce(SUM)`` over a flat fp32 gradient bucket (~5.8 MB at the default config)
over NVLink 5 / NVSwitch.  Forward activations are never exchanged.

Objective equivalence: the reference loss is ``BCE(sum over the batch) +
beta * sum_names mean_b(KL)`` (models.py:547,553,558) — a batch SUM plus a batch MEAN.  With R equal
shards, ``loss_r = recon_sum_r + beta * KL_mean_r / R`` summed over ranks equals the single-process
loss on the whole batch, so gradients are all-reduced with SUM (not averaged) and the local KL term
is scaled by 1/R (``SPAIR.kl_scale``).

``attn.*`` parameters never receive a gradient (the reference discards the attention output,
models.py:120): they are left out of the bucket and keep ``grad is None``.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment.  Returns (rank, world, local_rank);
    a plain ``python`` launch (no RANK in the environment) is world size 1 without a process group."""
    if "RANK" not in os.environ:
        return 0, 1, 0
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def trainable_parameters(module: torch.nn.Module) -> List[Tuple[str, torch.nn.Parameter]]:
    """Parameters that take part in training: everything except the dead ``attn.*`` branch."""
    return [(n, p) for n, p in module.named_parameters() if p.requires_grad and not n.startswith("attn.")]


class GradientBucket:
    """One flat fp32 buffer holding every trainable gradient; ``p.grad`` are views into it, so the
    allreduce needs no packing copies and ``zero()`` is a single memset."""

    ALIGN = 64      # floats: every tensor starts on a 256-byte boundary (vector loads, cuBLAS / cuDNN kernel selection)

    def __init__(self, named_params: Iterable[Tuple[str, torch.nn.Parameter]]):
        self.params = list(named_params)
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0][1].device
        self.offsets, off = [], 0
        for _, p in self.params:
            self.offsets.append(off)
            off += -(-p.numel() // self.ALIGN) * self.ALIGN
        # the padding between tensors stays zero: it adds nothing to the allreduce sum and an optimiser leaves it at 0
        self.flat = torch.zeros(off, device=dev, dtype=torch.float32)
        for (_, p), off in zip(self.params, self.offsets):
            # same strides as the parameter (e.g. channels_last conv weights): fused optimisers require matching layouts
            p.grad = self.flat[off:off + p.numel()].as_strided(p.shape, p.stride())

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def flatten_parameters(self) -> torch.nn.Parameter:
        """Moves the trainable parameters into ONE flat buffer (each ``p.data`` becomes a view of it, values kept) and
        returns it as a Parameter whose ``.grad`` is the gradient bucket.  An elementwise optimiser (Adam, SGD without
        per-tensor terms) over ``[flat_parameter]`` then updates every parameter in one kernel over 1.46 M elements
        instead of one multi-tensor launch over 46 small tensors (0.125 ms -> 0.02 ms per step at the default config).
        Call before a CUDA graph is captured (the graph holds the parameter addresses)."""
        if getattr(self, "flat_param", None) is not None:
            return self.flat_param
        flat = torch.zeros_like(self.flat)
        with torch.no_grad():
            for (name, p), off in zip(self.params, self.offsets):
                if not p.is_contiguous():
                    raise RuntimeError("parameter %s is not contiguous; cannot alias it into the flat buffer" % name)
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        self.flat_param = torch.nn.Parameter(flat, requires_grad=True)
        self.flat_param.grad = self.flat
        return self.flat_param

    def zero(self) -> None:
        self.flat.zero_()

    def check_attached(self) -> None:
        """Optimisers that set ``p.grad = None`` (``zero_grad(set_to_none=True)``) detach the views."""
        for name, p in self.params:
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or \
                    p.grad.data_ptr() >= self.flat.data_ptr() + self.nbytes:
                raise RuntimeError("gradient of %s is no longer a view of the bucket; use bucket.zero() "
                                   "instead of optimizer.zero_grad(set_to_none=True)" % name)

    def all_reduce(self, group=None, async_op: bool = False):
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class DataParallelSPAIR:
    """Wraps a ``SPAIR`` replica.  ``step(x_local, global_step)`` = zero grads, forward, backward,
    gradient allreduce; the optimiser step stays with the caller (identical on every rank because
    the reduced gradients are identical)."""

    def __init__(self, net, world_size: Optional[int] = None, group=None):
        self.net = net
        self.group = group
        self.world = world_size if world_size is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.bucket = GradientBucket(trainable_parameters(net))
        net.kl_scale = 1.0 / self.world

    def broadcast_parameters(self, src: int = 0) -> None:
        """Make every replica start from rank ``src``'s parameters (one flat broadcast)."""
        if self.world == 1:
            return
        tensors = [p.data for p in self.net.parameters()] + [b.data for b in self.net.buffers()]
        flat = torch.cat([t.reshape(-1).float() for t in tensors])
        dist.broadcast(flat, src=src, group=self.group)
        off = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n

    def step(self, x_local: torch.Tensor, global_step: int):
        """Returns (local loss, recon_x, z_where, z_pres) after gradients have been summed over ranks.
        The sum of the local losses over ranks is the single-process loss of the whole batch."""
        self.bucket.zero()
        out = self.net(x_local, global_step)
        out[0].backward()
        if self.world > 1:
            self.bucket.all_reduce(self.group)
        return out

    def global_loss(self, local_loss: torch.Tensor) -> torch.Tensor:
        total = local_loss.detach().clone()
        if self.world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return total


def shard_batch(x: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Contiguous shard of the batch for this rank (B must divide evenly so the KL means agree)."""
    if x.shape[0] % world:
        raise ValueError("batch %d is not divisible by world size %d" % (x.shape[0], world))
    per = x.shape[0] // world
    return x[rank * per:(rank + 1) * per]
