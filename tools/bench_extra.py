#!/usr/bin/env python
"""Secondary workloads of SURVEY.md §8(d) on one GPU (not the bench.py headline):
   cfg4  sample() 12-step CFG, 64x64 latents, bs=16, +clip_image
   cfg5  VQGAN encode -> indices -> decode_indices, 256x256 images
   cfg2+ the headline sample() followed by decode_indices (images/s including the decode)
Prints one JSON object per workload."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from paella_b200 import utils as U  # noqa: E402
from paella_b200 import _lib  # noqa: E402
from paella_b200.synth import rerandomize_, synthetic_conditioning  # noqa: E402
from paella_b200.vqgan import VQModel  # noqa: E402


def timed(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="cfg4,cfg5,cfg2dec")
    ap.add_argument("--vq-batch", type=int, default=64)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    which = args.which.split(",")
    model = None
    if "cfg4" in which or "cfg2dec" in which:
        model = bench.build_model(dev)
        model.pack_weights()
    if "cfg4" in which:
        B, H, steps = 16, 64, 12
        cond, uncond = synthetic_conditioning(B, bench.BYT5_LEN, with_clip_image=True, device=dev)
        fn = lambda: U.sample(model, cond, (B, H, H), uncond, steps=steps, renoise_steps=steps - 1)
        ms = timed(fn, 2, 3)
        print(json.dumps({"workload": "cfg4: sample() 12-step CFG, 64x64 latents, bs=16, L=128+clip+clip_image", "ms_per_call": ms,
                          "images_per_s": B / ms * 1e3, "algorithmic_tflops": 11.5 * B / ms * 1e3}))
    vq = None
    if "cfg5" in which or "cfg2dec" in which:
        vq = VQModel().eval()
        rerandomize_(vq.state_dict(), seed=1)
        vq = vq.to(dev)
        vq.pack_weights()
    if "cfg5" in which:
        B = args.vq_batch
        g = torch.Generator(device=dev).manual_seed(5)
        img = torch.rand(B, 3, 256, 256, device=dev, generator=g)
        idx = vq.encode(img)[2]
        ms_e = timed(lambda: vq.encode(img), 2, 5)
        ms_d = timed(lambda: vq.decode_indices(idx), 2, 5)
        print(json.dumps({"workload": f"cfg5: VQGAN f4 256x256 images, bs={B}", "encode_ms": ms_e, "decode_ms": ms_d,
                          "encode_images_per_s": B / ms_e * 1e3, "decode_images_per_s": B / ms_d * 1e3,
                          "encode_tflops": 0.029 * B / ms_e * 1e3, "decode_tflops": 0.136 * B / ms_d * 1e3}))
    if "cfg2dec" in which:
        B = bench.BATCH
        cond, uncond = synthetic_conditioning(B, bench.BYT5_LEN, device=dev)
        shape = (B, bench.LATENT, bench.LATENT)

        def fn():
            tok = U.sample(model, cond, shape, uncond, steps=bench.SAMPLE_STEPS, renoise_steps=bench.RENOISE)
            return vq.decode_indices(tok).clamp_(0, 1)
        ms = timed(fn, 2, 4)
        print(json.dumps({"workload": "cfg2 + decode_indices (32x32 tokens -> 128x128 px), bs=64", "ms_per_call": ms,
                          "images_per_s": B / ms * 1e3}))


if __name__ == "__main__":
    main()
