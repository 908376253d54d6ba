"""Data-parallel host logic on CPU: world_size 2, gloo backend, rendezvous on 127.0.0.1.

(1) GradientBucket + all_reduce(SUM) on a small stand-in module: bucket layout, exclusion of the dead
    ``attn.*`` branch, summed gradients identical on both ranks.
(2) The full SPAIR model (kernel binding replaced by the CPU test double, tests/cpu_kernel_mock.py) trained
    data-parallel on two half batches with the loss rule of SURVEY.md §8(e) reproduces the single-process
    gradients and loss of the whole batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


class _Patch:
    def setattr(self, obj, name, value):
        setattr(obj, name, value)


def _init(rank, world, port):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from spair_pytorch_b200 import dp
    assert dp.init_distributed("gloo") == (rank, world, rank)
    return dp


def _bucket_worker(rank, world, port, q):
    dp = _init(rank, world, port)

    class Toy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.a = torch.nn.Linear(5, 3)
            self.attn = torch.nn.Linear(2, 2)       # never used -> must stay out of the bucket
            self.b = torch.nn.Parameter(torch.ones(4))

    torch.manual_seed(0)
    m = Toy()
    bucket = dp.GradientBucket(dp.trainable_parameters(m))
    assert bucket.flat.numel() == 3 * dp.GradientBucket.ALIGN and bucket.offsets == [0, 64, 128]   # 4, 15, 3 elements, padded
    assert m.attn.weight.grad is None
    x = torch.arange(10.0).view(2, 5) + rank
    (m.a(x).sum() * (rank + 1) + (m.b * (rank + 2)).sum()).backward()
    assert [n for n, _ in bucket.params] == ["b", "a.weight", "a.bias"]
    assert m.b.grad.data_ptr() == bucket.flat.data_ptr()             # gradients accumulated in place in the bucket
    bucket.all_reduce()
    bucket.check_attached()
    q.put((rank, bucket.flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def _spair_worker(rank, world, port, q):
    dp = _init(rank, world, port)
    from oracle import spair_oracle as so
    from tests import cpu_kernel_mock, helpers
    cpu_kernel_mock.install(_Patch())
    net = helpers.build_model("tiny")
    ddp = dp.DataParallelSPAIR(net)
    ddp.broadcast_parameters()
    B = 4
    x = so.scattered_sprites(B, (1, 40, 40), seed=3, sprite_px=(6, 14))
    noise = so.random_noise(torch.Generator().manual_seed(9), B, (5, 5), 50)
    sl = slice(rank * B // world, (rank + 1) * B // world)
    net.set_noise(noise.eps_where[sl], noise.eps_attr[sl], noise.eps_depth[sl], noise.u_pres[sl])
    loss = ddp.step(dp.shard_batch(x, rank, world), 1001)[0]
    total = ddp.global_loss(loss)
    q.put((rank, float(total), ddp.bucket.flat.clone()))
    dist.barrier()
    dist.destroy_process_group()


def _spawn(fn, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=fn, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out, key=lambda t: t[0])


def test_gradient_bucket_allreduce_world2():
    (r0, g0), (r1, g1) = _spawn(_bucket_worker)
    assert torch.equal(g0, g1)
    # d/dW of sum(Linear(x)) * k summed over ranks: rows equal to sum_r (r+1) * sum_b x_r[b]
    x0, x1 = torch.arange(10.0).view(2, 5), torch.arange(10.0).view(2, 5) + 1
    want_row = 1 * x0.sum(0) + 2 * x1.sum(0)
    # tensors start on 64-float boundaries (GradientBucket.ALIGN); the padding stays zero through the allreduce
    assert torch.allclose(g0[:4], torch.full((4,), 2.0 + 3.0))
    assert torch.allclose(g0[64:79].view(3, 5), want_row.expand(3, 5))
    assert torch.allclose(g0[128:131], torch.full((3,), 2.0 * 1 + 2.0 * 2))
    assert not g0[4:64].any() and not g0[79:128].any() and not g0[131:].any()


def test_spair_data_parallel_equals_single_process(monkeypatch):
    from oracle import spair_oracle as so
    from spair_pytorch_b200 import dp
    from tests import cpu_kernel_mock, helpers
    (_, loss0, g0), (_, loss1, g1) = _spawn(_spair_worker)
    assert torch.equal(g0, g1) and loss0 == loss1
    cpu_kernel_mock.install(monkeypatch)
    net = helpers.build_model("tiny")
    bucket = dp.GradientBucket(dp.trainable_parameters(net))
    B = 4
    x = so.scattered_sprites(B, (1, 40, 40), seed=3, sprite_px=(6, 14))
    noise = so.random_noise(torch.Generator().manual_seed(9), B, (5, 5), 50)
    net.set_noise(noise.eps_where, noise.eps_attr, noise.eps_depth, noise.u_pres)
    loss = net(x, 1001)[0]
    loss.backward()
    assert abs(float(loss) - loss0) <= 1e-4 * abs(float(loss))
    scale = float(bucket.flat.norm()) / bucket.flat.numel() ** 0.5
    assert float((bucket.flat - g0).abs().max()) <= 1e-5 + 2e-4 * scale
