#!/usr/bin/env python
"""Data-parallel SPAIR training on 1..8 B200 — the launcher that sits beside the reference's train.py.

The reference trains on ONE device (train.py:27-30,41) although its README promises "all available GPUs"
(README.md:21).  This script is that missing piece, built on the drop-in `spair` package of this repo:

    python train_dp.py --steps 200                                        # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
           --master-addr 127.0.0.1 train_dp.py --steps 2000 --ckpt-dir runs/a --ckpt-every 500

* same model / seed / optimiser as train.py:39-44 (torch.manual_seed(3), Adam lr 1e-4);
* every rank trains on its own shard of the global batch (cfg.BATCH_SIZE per GPU by default); gradients are
  summed with one NCCL all_reduce over a flat bucket, the KL term is scaled by 1/world (dp.py) so the
  objective equals the single-process one on the global batch;
* the step (zero-grad + forward + backward) is replayed from one CUDA graph unless --eager;
* data: the HDF5 file of the reference (train.py:38, needs h5py) or, by default, procedurally generated
  scattered-sprite scenes with the same item schema;
* checkpoints hold model + optimiser + step (the reference saves the model only and cannot resume,
  train.py:85-90); `--resume` continues from the newest one.  The model part is the reference's own
  state_dict layout, so reference checkpoints load as well.
"""
from __future__ import annotations

import argparse
import glob
import os
import time

import torch
from torch.utils import data as torch_data

from spair import config as cfg
from spair import metric
from spair.dataloader import SimpleScatteredMNISTDataset, scattered_sprites_gpu
from spair.models import SPAIR
from spair_pytorch_b200 import dp
from spair_pytorch_b200.graphed import GraphedTrainStep


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass

    add_image = add_figure = add_histogram = add_scalar


def newest_checkpoint(ckpt_dir):
    files = sorted(glob.glob(os.path.join(ckpt_dir, "step_*.pt")), key=lambda f: int(os.path.basename(f)[5:-3]))
    return files[-1] if files else None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--batch", type=int, default=cfg.BATCH_SIZE, help="images per GPU")
    ap.add_argument("--hdf5", default=None, help="reference dataset (train/full/{image,bbox,digit_count}); default: procedural")
    ap.add_argument("--lr", type=float, default=1e-4)
    ap.add_argument("--start-step", type=int, default=0)
    ap.add_argument("--ckpt-dir", default=None)
    ap.add_argument("--ckpt-every", type=int, default=1000)
    ap.add_argument("--resume", action="store_true")
    ap.add_argument("--eager", action="store_true", help="do not capture the step into a CUDA graph")
    ap.add_argument("--log-every", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1000, help="noise / scene streams are a function of (seed, rank, step)")
    ap.add_argument("--save-final", default=None, help="write the model state_dict here after the last step (rank 0)")
    args = ap.parse_args()

    rank, world, local_rank = dp.init_distributed()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg.BATCH_SIZE = args.batch                                   # metric.py reads it
    writer = _NullWriter()
    if rank == 0:
        try:
            from tensorboardX import SummaryWriter                # train.py:9,21
            writer = SummaryWriter("logs_v2/dp")
        except ImportError:
            pass

    # fixed shapes: let cuDNN pick its fastest fp32 conv algorithms — unless bitwise reproducibility was asked for
    # (SPAIR_DETERMINISTIC=1: the autotuner may pick different algorithms in different processes)
    torch.backends.cudnn.benchmark = "SPAIR_DETERMINISTIC" not in os.environ
    torch.manual_seed(3)                                          # train.py:39
    net = SPAIR(cfg.INPUT_IMAGE_SHAPE, writer if args.eager else _NullWriter(), dev).to(dev)
    ddp = dp.DataParallelSPAIR(net, world_size=world)
    ddp.broadcast_parameters()
    opt = torch.optim.Adam([ddp.bucket.flatten_parameters()], lr=args.lr, fused=True)   # train.py:44, over the flat buffer
    step = args.start_step
    if args.resume and args.ckpt_dir and newest_checkpoint(args.ckpt_dir):
        ck = torch.load(newest_checkpoint(args.ckpt_dir), map_location=dev)
        net.load_state_dict(ck["model"])
        opt.load_state_dict(ck["optim"])
        step = ck["step"] + 1
        if rank == 0:
            print("resumed from step", ck["step"])

    loader, sampler, epoch = None, None, 0
    if args.hdf5:                                                 # the reference's dataset, one DataLoader per rank
        dataset = SimpleScatteredMNISTDataset(args.hdf5)
        sampler = torch_data.distributed.DistributedSampler(dataset, world, rank) if world > 1 else None
        loader = torch_data.DataLoader(dataset, batch_size=args.batch, pin_memory=True, num_workers=2, drop_last=True,
                                       sampler=sampler)
        it = iter(loader)
    data_gen = torch.Generator(device=dev)                        # procedural scenes made on the device
    gstep, t0, seen = None, time.time(), 0
    last = step + args.steps
    while step < last:
        if loader is not None:
            try:
                x_image, y_bbox, y_count = next(it)
            except StopIteration:
                epoch += 1
                if sampler is not None:
                    sampler.set_epoch(epoch)      # a new shuffle every epoch (DistributedSampler repeats its order otherwise)
                it = iter(loader)
                continue
            x_image = x_image.float()
        else:
            data_gen.manual_seed(1234 + 7919 * rank + 104729 * step)
            x_image, y_bbox, y_count = scattered_sprites_gpu(args.batch, cfg.INPUT_IMAGE_SHAPE, dev, data_gen)
        if gstep is None and not args.eager:
            gstep = GraphedTrainStep(net, x_image.to(dev), bucket=ddp.bucket, global_step=step)
        # the latent noise of a step is a function of (seed, rank, step), not of how many draws came before it: a captured
        # graph reads the generator's seed at replay, so re-seeding here makes `--resume` continue the SAME trajectory
        torch.cuda.manual_seed(args.seed + rank + 1000003 * step)
        if args.eager:
            out = ddp.step(x_image.to(dev, non_blocking=True), step)
        else:
            out = gstep(x_image, step)                            # async H2D + one graph replay
            if world > 1:
                ddp.bucket.all_reduce()
        opt.step()
        seen += args.batch * world
        if step % args.log_every == 0:
            loss = ddp.global_loss(out[0])                        # device->host only when logging
            if rank == 0:
                dt = time.time() - t0
                print("step %6d  loss/img %.3f  %.0f img/s" % (step, float(loss) / (args.batch * world), seen / max(dt, 1e-9)))
                writer.add_scalar("losses/total", float(loss), step)
                if step > 1000 and step % (5 * args.log_every) == 0:                # train.py:76-82
                    z_where, z_pres = out[2].detach().clone(), out[3].detach().clone()
                    ap_ = metric.mAP(z_where, z_pres, y_bbox.to(dev).float(), y_count.to(dev).float())
                    print("           bbox AP %.4f  count err %.3f"
                          % (float(ap_), float(metric.object_count_accuracy(z_pres, y_count.to(dev).float()))))
        if args.ckpt_dir and rank == 0 and step > 0 and step % args.ckpt_every == 0:
            os.makedirs(args.ckpt_dir, exist_ok=True)
            torch.save({"model": net.state_dict(), "optim": opt.state_dict(), "step": step},
                       os.path.join(args.ckpt_dir, "step_%d.pt" % step))
        step += 1
    torch.cuda.synchronize()
    if args.save_final and rank == 0:
        torch.save(net.state_dict(), args.save_final)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
