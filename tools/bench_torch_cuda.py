#!/usr/bin/env python
"""PyTorch-on-the-same-GPU baseline for the headline workload (NOT part of the product, NOT run by bench.py):
the torch restatement of the reference in oracle/paella_oracle.py executed on `cuda` — fp32 (TF32 off / on) and
fp16 GEMM operands (what torch.autocast gives the reference in the notebook).  The reference tree itself cannot travel
to the GPU box, so this port is the closest stand-in for "the reference's own PyTorch-CUDA path" of BASELINE.json.

    python tools/bench_torch_cuda.py --batch 64 --iters 2        # prints one JSON line per mode
    python tools/bench_torch_cuda.py --device cpu --tiny         # plumbing check without a GPU
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import paella_oracle as po  # noqa: E402
from paella_b200.modules import Paella  # noqa: E402
from paella_b200.synth import rerandomize_, synthetic_conditioning  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--batch", type=int, default=bench.BATCH)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--tiny", action="store_true", help="small config / latent for a plumbing check")
    args = ap.parse_args()
    dev = torch.device(args.device)
    if args.tiny:
        kw = dict(c_in=16, c_out=16, num_labels=64, c_r=64, c_cond=32, c_hidden=[32, 64, 64], nhead=[-1, 4, 4], blocks=[1, 2, 1],
                  clip_embd=24, byt5_embd=40)
        latent, steps, L = 8, 2, 8
    else:
        kw, latent, steps, L = dict(byt5_embd=2560), bench.LATENT, bench.SAMPLE_STEPS, bench.BYT5_LEN
    torch.manual_seed(0)
    m = Paella(**kw).eval()
    rerandomize_(m.state_dict(), seed=0)
    sd = {k: v.to(dev) for k, v in m.state_dict().items()}
    cfg = po.PaellaConfig(**kw)
    B = args.batch
    cond, uncond = synthetic_conditioning(B, L, byt5_embd=kw["byt5_embd"], clip_embd=kw.get("clip_embd", 1024), device=dev)
    K = cfg.num_labels
    g = torch.Generator(device=dev).manual_seed(1)

    def one(mm):
        draws = {"init": torch.randint(0, K, (B, latent, latent), generator=g, device=dev),
                 "q": [torch.empty(B * latent * latent, K, device=dev).exponential_(1, generator=g) for _ in range(steps)],
                 "u": [torch.rand(B, latent, latent, generator=g, device=dev) for _ in range(steps - 1)]}
        with torch.inference_mode():
            return po.sample(sd, cfg, cond, (B, latent, latent), uncond, steps=steps, renoise_steps=steps - 1,
                             temperature=(1.0, 0.2), cfg_scale=8.0, draws=draws, mm=mm)

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize()

    def mm_half(a, w):          # a real half-precision tensor-core GEMM (what autocast dispatches), fp32 result
        return (a.half() @ w.half().t()).float()
    modes = [("fp32", po.mm_fp32, False), ("fp32+tf32", po.mm_fp32, True), ("fp16 gemm", mm_half, False)]
    for name, mm, tf32 in modes:
        torch.backends.cuda.matmul.allow_tf32 = tf32
        one(mm)
        sync()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            one(mm)
        sync()
        dt = (time.perf_counter() - t0) / args.iters
        print(json.dumps({"impl": "torch restatement on " + str(dev), "mode": name, "batch": B, "latent": latent, "steps": steps,
                          "ms_per_call": dt * 1e3, "images_per_s": B / dt}), flush=True)


if __name__ == "__main__":
    main()
