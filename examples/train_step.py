#!/usr/bin/env python
"""BASELINE config 4: the IRN video-watermark train step — encoder -> tamper/splice -> hybrid attack ->
localiser, one optimisation step, bf16 autocast around the networks, attack layer in fp32, DDP
batch-sharded (one process per GPU; NCCL only in the gradient all-reduce).

The reference's step is models/IRNcrop_model.py:325-416: netG(real_H) -> clamp -> Quantization ->
`forward*(1-mask) + previous*mask` (:348) -> per-frame 5-way attack loop (:357-370: Resize,
combined_jpeg_strong, combined_jpeg_weak, MiddleBlur(3), GaussianBlur, mixed with softmax weights) ->
clamp -> Quantization -> generator (UNet) -> BCEWithLogits losses (:378-404) -> AdamW.
Networks: the reference's own `Inveritible_Decolorization_PAMI(block_num=[1,1,1], ResBlock)` and
`UNet(3,1,32)` when baseline/_ref carries them (vendored, unmodified, git-ignored), else small conv
stand-ins with the same interface (stated in the record).  The networks are OUT OF SCOPE of this repo
(dense convs served by cuDNN); in scope and exercised: the attack layer (`wmattack.*`, or — for the
comparison arm — the reference's `noise_layers` modules on the same GPU), the splice prologue and the
Quantization epilogue.  Frames are folded into the batch axis (the reference loops over t in python).

    python examples/train_step.py --steps 5                       # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/train_step.py
"""
import argparse
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-watermarking-forgery-detection_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import wmattack  # noqa: E402
from wmattack.sharding import frame_shard  # noqa: E402

JPEG_QS = (50, 80, 90, 70, 60)


def conv_stack(cin, cout, width=32):
    return nn.Sequential(nn.Conv2d(cin, width, 3, padding=1), nn.ELU(), nn.Conv2d(width, width, 3, padding=1), nn.ELU(),
                         nn.Conv2d(width, cout, 3, padding=1))


def build_nets(kind, h, w):
    """("reference" | "standin") -> encoder, localiser, description."""
    if kind == "reference":
        from baseline import ref_harness as RH
        inv, unet = RH.networks()
        enc = inv.Inveritible_Decolorization_PAMI(dims_in=[[3, h, w]], block_num=[1, 1, 1], subnet_constructor=inv.ResBlock)
        loc = unet.UNet(in_channels=3, out_channels=1, init_features=32)
        return enc, loc, "reference Inveritible_Decolorization_PAMI([1,1,1], ResBlock) + UNet(3,1,32) from baseline/_ref"
    return conv_stack(3, 3), conv_stack(3, 1), "stand-in 3-layer conv stacks (baseline/_ref absent)"


def build_attacks(kind, dev):
    """The five attacks of models/IRNcrop_model.py:96-104, ours or the reference's own modules."""
    if kind == "identity":
        return [wmattack.Identity() for _ in range(5)]
    if kind == "ours":
        A = wmattack
    else:
        from baseline import ref_harness as RH
        A = RH.layers("cuda")
    def jpegs():
        members = []
        for q in JPEG_QS:
            members += [A.JpegMask(q), A.Jpeg(q)]
        members += [A.JpegSS(q) for q in (50, 60, 70, 80, 90)]
        return A.Combined(members)
    layers = [A.Resize(), jpegs(), jpegs(), A.MiddleBlur(3), A.GaussianBlur()]
    if kind != "ours":
        for m in layers:
            m.to(dev)
        # upstream Combined reads selected.name (combined.py:19); MiddleBlur has none -> only the jpeg Combined is used
    return layers


class Step(nn.Module):
    def __init__(self, h, w, nets="standin", attack="ours", dev="cuda"):
        super().__init__()
        self.encoder, self.localiser, self.nets_desc = build_nets(nets, h, w)
        self.attack_kind = attack
        self.attacks = build_attacks(attack, dev)          # plain list: attack layers hold no parameters
        self.splice = wmattack.Splice()
        self.quant = wmattack.Quantization()
        self.mix = wmattack.AttackMix() if attack == "ours" else None      # the reference arm keeps the trainer's torch ops
        self.is_invertible = nets == "reference"
        self.attack_ms = None

    def forward(self, frames, previous, mask):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            if self.is_invertible:
                marked = self.encoder(frames)
            else:
                marked = frames + 0.1 * torch.tanh(self.encoder(frames))
        marked = marked.float()
        marked = marked + (torch.clamp(marked, 0, 1) - marked).detach()     # clamp_with_grad (IRNcrop_model.py:344)
        marked = self.quant(marked)                                          # :345
        tampered = self.splice(marked, previous, mask)                       # :348  forward*(1-mask) + previous*mask
        # hybrid attack (:357-370): softmax-weighted mix of the five attacked versions, per frame
        alpha = torch.softmax(torch.randn(frames.shape[0], 5, device=frames.device), dim=1)
        if self.mix is not None:             # our arm: mix + clamp_with_grad (:373) + Quantization (:374) in one pass
            attacked = self.mix([layer(tampered) for layer in self.attacks], alpha)
        else:
            attacked = None
            for k, layer in enumerate(self.attacks):
                y = layer(tampered)
                y = y[0] if isinstance(y, tuple) else y
                term = alpha[:, k].view(-1, 1, 1, 1) * y
                attacked = term if attacked is None else attacked + term
            attacked = attacked + (torch.clamp(attacked, 0, 1) - attacked).detach()   # :373
            attacked = self.quant(attacked)                                           # :374
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = self.localiser(attacked)
        bce = nn.functional.binary_cross_entropy_with_logits
        loss_loc = bce(logits.float(), mask)                                      # l_backward (:392)
        loss_img = bce(marked, frames)                                            # l_forward (:385-388)
        return loss_loc + loss_img, loss_loc.detach(), loss_img.detach()


def make_batch(clips, frames, size, rank, world, dev, seed=1234):
    a, b = frame_shard(clips, rank, world)                        # contiguous clip shard of this rank
    gen = torch.Generator(dev).manual_seed(seed + rank)
    n = (b - a) * frames
    x = torch.rand(n, 3, size, size, device=dev, generator=gen)
    mask = (torch.rand(n, 1, size, size, device=dev, generator=gen) > 0.88).float()     # ~12 % ones (DAVIS mask rate)
    return x, x.roll(1, 0), mask


def run(attack, nets, rank, world, dev, clips=64, frames=8, size=256, steps=5, warmup=2, verbose=False):
    """ms per optimisation step (CUDA events, max over ranks) of the config-4 step with the given attack arm."""
    torch.manual_seed(10); np.random.seed(10); random.seed(10)    # train.py:317-329: every rank shares the seed
    model = Step(size, size, nets=nets, attack=attack, dev=dev).to(dev)
    net = nn.parallel.DistributedDataParallel(model, device_ids=[dev.index]) if world > 1 else model
    opt = torch.optim.AdamW(model.parameters(), lr=2e-4)
    x, prev, mask = make_batch(clips, frames, size, rank, world, dev)

    def one():
        loss, l_loc, l_img = net(x, prev, mask)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss
    for _ in range(warmup):
        loss = one()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(steps):
        loss = one()
        if verbose and rank == 0:
            print(f"step {s}: loss {float(loss.detach()):.4f}", flush=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    desc, lossv = model.nets_desc, float(loss.detach())
    del model, net, opt, x, prev, mask
    torch.cuda.empty_cache()
    return ms, desc, lossv


def bench(rank, world, dev, clips=64, frames=8, size=256, steps=5, warmup=2):
    """bench.py's `train_step` record: steps/s and frames/s of the config-4 step at this world size with
    (a) our attack layer, (b) the reference's attack modules on the same GPUs, (c) no attack (nets only)."""
    from baseline import ref_harness as RH
    nets = "reference" if RH.available() else "standin"
    out = {"workload": f"configs[3]: IRN train step, {clips} clips x {frames} frames x 3x{size}x{size} global "
                       f"({clips // world} clips per rank), bf16 autocast nets + fp32 attack layer, AdamW, "
                       f"DDP x{world}" + ("" if world > 1 else " (single process)")}
    n_frames = clips * frames
    res = {}
    arms = ["ours", "identity"] + (["reference"] if RH.available() else [])
    for arm in arms:
        try:
            ms, desc, lossv = run(arm, nets, rank, world, dev, clips, frames, size, steps, warmup)
            res[arm] = ms
            out["nets"] = desc
            out[f"{arm}_attack"] = {"ms_per_step": round(ms, 2), "steps_per_s": round(1e3 / ms, 3),
                                    "frames_per_s": round(n_frames * 1e3 / ms, 1), "loss": round(lossv, 4)}
        except Exception as e:
            out[f"{arm}_attack"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.empty_cache()
    if "ours" in res and "identity" in res:
        out["attack_layer_share_ours"] = round(max(0.0, 1 - res["identity"] / res["ours"]), 4)
    if "reference" in res and "identity" in res:
        out["attack_layer_share_reference"] = round(max(0.0, 1 - res["identity"] / res["reference"]), 4)
    if "ours" in res and "reference" in res:
        out["step_speedup_vs_reference_attack"] = round(res["reference"] / res["ours"], 3)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--clips", type=int, default=8, help="global number of clips [B,3,T,H,W]")
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--attack", default="ours", choices=("ours", "reference", "identity"))
    ap.add_argument("--nets", default="auto", choices=("auto", "reference", "standin"))
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=dev)
    nets = args.nets
    if nets == "auto":
        from baseline import ref_harness as RH
        nets = "reference" if RH.available() else "standin"
    ms, desc, lossv = run(args.attack, nets, rank, world, dev, args.clips, args.frames, args.size, args.steps, 1, verbose=True)
    if rank == 0:
        print(f"{ms:.2f} ms/step, {args.clips * args.frames * 1e3 / ms:.0f} frames/s, attack={args.attack}, nets: {desc}")
    if world > 1:
        torch.distributed.destroy_process_group()
    return lossv


if __name__ == "__main__":
    main()
