#!/usr/bin/env python
"""Benchmark of the hot path: images/s of the 8-step CFG sample() at 32x32 latents, bs=64 per GPU.

  python bench.py --gpus N --steps K --warmup W                    (our CUDA path; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K --warmup W   (the reference algorithm on the host CPU cores)

One "step" = one full sample() call over one batch (BASELINE.json configs[1]: bs=64, 32x32 latents,
8 denoising steps with classifier-free guidance = 16 denoiser sample-forwards + 8 resamples per image), the
reference-default 1.008 B denoiser, synthetic ByT5/CLIP embeddings (L=128), re-randomised weights
(paella_b200/synth.py).  Prints ONE JSON line (rank 0).

  value   images/s with the conditioning tensors already resident in HBM
  e2e     images/s through the public API with HOST (pinned) conditioning: H2D of byt5/clip (cond+uncond) and
          D2H of the sampled tokens inside the timed region, every step
  roofline  tcgen05 GEMM family: algorithmic FLOPs of one step / summed CUDA-event durations of its launches,
          measured live by one extra profiled step after the timed region (events on the launching stream)
  cpu_baseline  oracle port (torch CPU fp32) on a bounded sample of the same workload, rank 0, N=1 only
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec @256x256-class (32x32 latent, 8192 codes), 8-step CFG sample, bs=64 per GPU"
LATENT, SAMPLE_STEPS, RENOISE, BYT5_LEN, BATCH = 32, 8, 7, 128, 64
WORKLOAD = ("sample() 8-step CFG, 32x32 latents, 8192 codes, 1.008B denoiser (reference default; readme says 573M), "
            "L_byt5=128+clip, synthetic embeddings, re-randomised weights")


def log(msg):
    print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def host_threads():
    """Threads torch will actually use on this host (respects the container's CPU limit better than os.cpu_count())."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        n = os.cpu_count() or 1
    n = max(1, min(n, torch.get_num_threads() if torch.get_num_threads() > 0 else n))
    torch.set_num_threads(n)
    return n


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def ncu_traffic():
    """DRAM bytes per launch of the dominant (GEMM) kernels from the committed `ncu --set full` capture: the
    launch-weighted mean over the main-path shapes of one forward (bench.py cannot run ncu itself)."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r01_ncu_traffic.json")
    try:
        ks = json.load(open(path))["kernels"]
        n = sum(k["launches_per_forward"] for k in ks)
        mean = sum(k["launches_per_forward"] * k["dram_bytes"] for k in ks) / n
        alg = sum(k["launches_per_forward"] * k["algorithmic_bytes"] for k in ks) / n
        return mean, f"profiles/r01_ncu_traffic.json (launch-weighted mean of {len(ks)} shapes; algorithmic {alg / 1e6:.0f} MB)"
    except (OSError, KeyError, ValueError, ZeroDivisionError):
        return None, None


def build_model(device):
    from paella_b200.modules import Paella
    from paella_b200.synth import rerandomize_
    torch.manual_seed(0)
    m = Paella(byt5_embd=2560).eval()
    rerandomize_(m.state_dict(), seed=0)
    return m.to(device)


def reference_arm(args, rank, world):
    """The reference's algorithm on the box's host cores (oracle port: torch CPU fp32, all threads).
    A step = one bounded sample of the same workload: the full 8-step CFG sample() at batch `ref_batch`."""
    if rank != 0:
        return
    from oracle import paella_oracle as po
    from paella_b200.modules import Paella
    from paella_b200.synth import rerandomize_, synthetic_conditioning
    cores = host_threads()
    log(f"reference arm: oracle port on {cores} host threads")
    torch.manual_seed(0)
    m = Paella(byt5_embd=2560).eval()
    rerandomize_(m.state_dict(), seed=0)
    sd = {k: v for k, v in m.state_dict().items()}
    oc = po.PaellaConfig(byt5_embd=2560)
    B = args.ref_batch
    cond, uncond = synthetic_conditioning(B, BYT5_LEN)
    g = torch.Generator().manual_seed(1)

    def one():
        draws = {"init": torch.randint(0, 8192, (B, LATENT, LATENT), generator=g),
                 "q": [torch.empty(B * LATENT * LATENT, 8192).exponential_(1, generator=g) for _ in range(SAMPLE_STEPS)],
                 "u": [torch.rand(B, LATENT, LATENT, generator=g) for _ in range(RENOISE)]}
        with torch.inference_mode():
            return po.sample(sd, oc, cond, (B, LATENT, LATENT), uncond, steps=SAMPLE_STEPS, renoise_steps=RENOISE,
                             temperature=(1.0, 0.2), cfg_scale=8.0, draws=draws)
    for _ in range(args.warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t0
    val = B * args.steps / dt
    sample = f"full 8-step CFG sample() at bs={B} (of the bs=64 workload), torch CPU fp32, {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": B},
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--ref-batch", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-baseline-only", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cpu_baseline_only:
        print(json.dumps(cpu_baseline()), flush=True)
        return

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import torch.distributed as dist
    from paella_b200 import _lib
    from paella_b200 import utils as U
    from paella_b200.synth import synthetic_conditioning
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    model = build_model(dev)
    model.pack_weights(broadcast_src=0 if world > 1 else None)       # the one collective: weight blob broadcast
    B = args.batch
    cond_h, uncond_h = synthetic_conditioning(B, BYT5_LEN, seed=1234 + rank, pin=True)
    cond_d = {k: v.to(dev) for k, v in cond_h.items()}
    uncond_d = {k: v.to(dev) for k, v in uncond_h.items()}
    shape = (B, LATENT, LATENT)
    torch.manual_seed(1234 + rank)

    def step_resident():
        return U.sample(model, cond_d, shape, uncond_d, steps=SAMPLE_STEPS, renoise_steps=RENOISE, temperature=(1.0, 0.2), cfg=8.0)

    host_tokens = torch.empty(shape, dtype=torch.int64).pin_memory()

    def step_e2e():
        c = {k: v.to(dev, non_blocking=True) for k, v in cond_h.items()}
        u = {k: v.to(dev, non_blocking=True) for k, v in uncond_h.items()}
        toks = U.sample(model, c, shape, u, steps=SAMPLE_STEPS, renoise_steps=RENOISE, temperature=(1.0, 0.2), cfg=8.0)
        host_tokens.copy_(toks, non_blocking=True)
        return toks

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms)

    log(f"rank {rank}: model packed; warm-up x{max(args.warmup, 3)}")
    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    log(f"timing {args.steps} steps (HBM-resident inputs)")
    launches0 = L.pb200_launch_count()
    with ClockSampler(local_rank) as clk:
        ms = timed(step_resident, args.steps)
    launches = L.pb200_launch_count() - launches0
    log(f"  {ms / args.steps:.1f} ms/step; timing {args.steps} steps end to end (host inputs)")
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    log(f"  {ms_e2e / args.steps:.1f} ms/step; profiled step")
    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in list(cond_h.values()) + list(uncond_h.values()))
    d2h = host_tokens.numel() * host_tokens.element_size()

    # one extra profiled step: CUDA events around every launch, by kernel family
    L.pb200_profile_enable(1)
    step_resident()
    buf = (b"\0" * 65536)
    import ctypes
    cbuf = ctypes.create_string_buffer(65536)
    _lib.check(L.pb200_profile_report(cbuf, 65536), "profile_report")
    prof = json.loads(cbuf.value.decode())
    L.pb200_profile_enable(0)
    pk = peaks()
    gemm_ms = sum(v["ms"] for k, v in prof.items() if k.startswith("gemm") or k == "fused_sampler")
    gemm_fl = sum(v["work"] for k, v in prof.items() if k.startswith("gemm") or k == "fused_sampler")
    gemm_n = sum(v["launches"] for k, v in prof.items() if k.startswith("gemm") or k == "fused_sampler")
    total_ms = sum(v["ms"] for v in prof.values())
    achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic, traffic_src = ncu_traffic()
    roofline = {"bound": "tensor", "kernel": "gemm_f16_kernel (tcgen05 GEMM family incl. fused sampler)", "achieved": achieved,
                "peak": pk["tflops"], "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": traffic,
                "traffic_source": traffic_src,
                "peak_source": pk["src"] + " bf16 sustained", "launches_per_step": gemm_n,
                "avg_launch_ms": gemm_ms / max(gemm_n, 1), "share_of_step": gemm_ms / total_ms if total_ms else None,
                "families": {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                 "rate": (v["work"] / (v["ms"] / 1e3) / 1e12 if v["ms"] > 0 else None)} for k, v in prof.items()}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    out = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
           "data": "synthetic",
           "config": {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": B * world, "parallelism": f"batch-shard x{world}, 1 weight broadcast",
                      "l2": "inputs larger than L2 (2.0 GB fp16 weights + activations per step)"},
           "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": ms_e2e / args.steps},
           "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        # in a child process with a hard limit, so the GPU line is printed whatever the host CPU does
        log("cpu baseline (oracle port on the host cores, bounded to one image) ...")
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-baseline-only"], capture_output=True, text=True,
                               timeout=240)
            out["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": f"not measured: {type(e).__name__} (240 s limit)"}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline():
    """Oracle port on the host cores: ONE sample-forward pair... bounded to ~10-30 s: the full 8-step CFG sample() at bs=1."""
    from oracle import paella_oracle as po
    from paella_b200.modules import Paella
    from paella_b200.synth import rerandomize_, synthetic_conditioning
    cores = host_threads()
    torch.manual_seed(0)
    m = Paella(byt5_embd=2560).eval()
    rerandomize_(m.state_dict(), seed=0)
    sd = {k: v for k, v in m.state_dict().items()}
    oc = po.PaellaConfig(byt5_embd=2560)
    cond, uncond = synthetic_conditioning(1, BYT5_LEN)
    g = torch.Generator().manual_seed(1)
    draws = {"init": torch.randint(0, 8192, (1, LATENT, LATENT), generator=g),
             "q": [torch.empty(LATENT * LATENT, 8192).exponential_(1, generator=g) for _ in range(SAMPLE_STEPS)],
             "u": [torch.rand(1, LATENT, LATENT, generator=g) for _ in range(RENOISE)]}
    with torch.inference_mode():
        po.paella_forward(sd, oc, draws["init"], torch.ones(1), **cond)          # warm-up
        t0 = time.perf_counter()
        po.sample(sd, oc, cond, (1, LATENT, LATENT), uncond, steps=SAMPLE_STEPS, renoise_steps=RENOISE, temperature=(1.0, 0.2),
                  cfg_scale=8.0, draws=draws)
        dt = time.perf_counter() - t0
    return {"value": 1.0 / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"one image: full 8-step CFG sample() at bs=1 (16 forwards + 8 resamples), torch CPU fp32, {cores} threads, {dt:.1f} s"}


if __name__ == "__main__":
    main()
