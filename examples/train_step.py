#!/usr/bin/env python
"""BASELINE config 4 in miniature: encoder -> tamper/splice -> K-way attack bank -> localiser, one
optimisation step, DDP batch-sharded (one process per GPU, NCCL only in the gradient all-reduce).

The reference's step is models/IRNcrop_model.py:325-416 (per-frame attack loop :357-370) and
models/IRNp_model.py:609-686 (8-way attack, straight-through, Quantization).  The encoder/localiser
networks themselves are OUT OF SCOPE of this repo (dense conv nets served by
cuDNN); small stand-ins with the same interface are used here so the step runs anywhere.  What IS
in scope and exercised: the attack layer (wmattack.*), the splice prologue, the fused
clamp+straight-through+Quantization epilogue writing into the K-way batch, frames folded into the
batch axis, bf16 autocast around the networks with the attack layer in fp32.

    python examples/train_step.py --steps 5                       # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_step.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-watermarking-forgery-detection_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import wmattack  # noqa: E402
from wmattack.sharding import frame_shard  # noqa: E402


def conv_stack(cin, cout, width=32):
    return nn.Sequential(nn.Conv2d(cin, width, 3, padding=1), nn.ELU(), nn.Conv2d(width, width, 3, padding=1), nn.ELU(),
                         nn.Conv2d(width, cout, 3, padding=1))


class Step(nn.Module):
    def __init__(self, h, w):
        super().__init__()
        self.encoder = conv_stack(3, 3)           # stand-in for Inveritible_Decolorization_PAMI
        self.localiser = conv_stack(3, 1)         # stand-in for UNet(3, 1, 32)
        self.splice = wmattack.Splice()
        # differentiable JPEG family with true gradients + the straight-through bank of the trainers
        self.diffjpeg = wmattack.DiffJPEG(True, h, w, quality=75)
        self.bank = wmattack.AttackBank([
            wmattack.Resize(), wmattack.Combined([wmattack.JpegMask(70), wmattack.Jpeg(70), wmattack.JpegSS(70)]),
            wmattack.MiddleBlur(3), wmattack.GaussianBlur(), wmattack.Gaussian(), wmattack.Identity()])

    def forward(self, frames, previous, mask):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            marked = torch.clamp(frames + 0.1 * torch.tanh(self.encoder(frames)), 0, 1)
        marked = marked.float()                                   # attack layer runs in fp32
        tampered = self.splice(marked, previous, mask)            # forward*(1-mask) + previous*mask
        attacked = self.bank(tampered)                            # [K*B,3,H,W], clamp + STE + 8-bit quantise
        attacked = torch.cat([attacked, self.diffjpeg(tampered)], 0)   # a branch with TRUE attack gradients
        k = attacked.shape[0] // frames.shape[0]
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = self.localiser(attacked)
        target = mask.repeat(k, 1, 1, 1)
        loss_loc = nn.functional.binary_cross_entropy_with_logits(logits.float(), target)
        loss_img = nn.functional.mse_loss(marked, frames)
        return loss_loc + loss_img, loss_loc.detach(), loss_img.detach()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--clips", type=int, default=8, help="global number of clips [B,3,T,H,W]")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--size", type=int, default=128)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    torch.manual_seed(10); np.random.seed(10)                     # train.py:317-329: every rank shares the seed
    model = Step(args.size, args.size).to(dev)
    net = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
    a, b = frame_shard(args.clips, rank, world)                   # contiguous clip shard of this rank
    gen = torch.Generator(dev).manual_seed(1234)
    for step in range(args.steps):
        clip = torch.rand(args.clips, 3, args.frames, args.size, args.size, device=dev, generator=gen)[a:b]
        mask5 = (torch.rand(args.clips, 1, args.frames, args.size, args.size, device=dev, generator=gen)[a:b] > 0.88).float()
        # frames folded into the batch axis (the reference loops over t in python, IRNcrop_model.py:357)
        frames = clip.permute(0, 2, 1, 3, 4).reshape(-1, 3, args.size, args.size)
        mask = mask5.permute(0, 2, 1, 3, 4).reshape(-1, 1, args.size, args.size)
        previous = frames.roll(1, 0)
        loss, l_loc, l_img = net(frames, previous, mask)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        if rank == 0:
            print(f"step {step}: loss {float(loss):.4f} (localise {float(l_loc):.4f}, image {float(l_img):.5f}) "
                  f"attacks {model.bank.names}", flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()
    return float(loss)


if __name__ == "__main__":
    main()
