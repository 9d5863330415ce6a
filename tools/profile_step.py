#!/usr/bin/env python
"""One sample() call of the bench workload bracketed by cudaProfilerStart/Stop, for ncu:
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_step.py
   ncu --profile-from-start off --set full --import-source on -k regex:gemm_f16_kernel -c 3 -o Y python tools/profile_step.py
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from paella_b200 import utils as U  # noqa: E402
from paella_b200.synth import synthetic_conditioning  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--sample-steps", type=int, default=8)
args = ap.parse_args()
dev = torch.device("cuda", 0)
model = bench.build_model(dev)
model.pack_weights()
cond, uncond = synthetic_conditioning(args.batch, bench.BYT5_LEN, device=dev)
shape = (args.batch, bench.LATENT, bench.LATENT)
torch.manual_seed(0)
U.sample(model, cond, shape, uncond, steps=2, renoise_steps=1)          # warm-up (lazy init, tensor maps)
torch.cuda.synchronize()
torch.cuda.profiler.start()
U.sample(model, cond, shape, uncond, steps=args.sample_steps, renoise_steps=args.sample_steps - 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one sample() call, batch", args.batch)
