"""Synthetic weights and inputs for parity tests and benchmarks.

The reference's own initialisation makes a random-init model degenerate
(SURVEY.md F3): ``clf[1].weight`` is zero (logits identically 0,
ref/src/modules.py:195), every ``TimestepBlock.mapper.weight`` is zero (:203-204),
GRN gamma/beta are zero (:34-35), the VQGAN ``ResBlock.gammas`` are zero
(ref/src/vqgan.py:23) and BatchNorm running stats are (0,1) (:66).  A broken
kernel would pass against that.  ``rerandomize_`` replaces every such tensor
with seeded noise; the recipe is part of the benchmark definition:

  * any floating-point tensor that is entirely zero            -> N(0, 0.02)
  * VQGAN ``*.gammas``                                          -> N(0, 0.5)
  * BatchNorm ``running_mean`` -> N(0,1);  ``running_var``      -> U(0.5, 1.5)
  * ``vquantizer.codebook.weight``                              -> N(0, 1)

All draws come from one CPU ``torch.Generator`` walked over the state-dict in
key order, so the result depends only on (key order, shapes, seed).
"""
from __future__ import annotations

from typing import Dict

import torch


def rerandomize_(state_dict: Dict[str, torch.Tensor], seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device="cpu").manual_seed(seed)
    for name, t in state_dict.items():
        if not torch.is_floating_point(t):
            continue
        new = None
        if name.endswith("gammas"):
            new = torch.randn(t.shape, generator=g) * 0.5
        elif name.endswith("running_mean"):
            new = torch.randn(t.shape, generator=g)
        elif name.endswith("running_var"):
            new = torch.rand(t.shape, generator=g) + 0.5
        elif name.endswith("codebook.weight"):
            new = torch.randn(t.shape, generator=g)
        elif t.numel() > 0 and bool((t == 0).all()):
            new = torch.randn(t.shape, generator=g) * 0.02
        if new is not None:
            with torch.no_grad():
                t.copy_(new.to(device=t.device, dtype=t.dtype))
    return state_dict


def synthetic_conditioning(batch: int, byt5_len: int = 128, byt5_embd: int = 2560, clip_embd: int = 1024,
                           with_clip: bool = True, with_clip_image: bool = False, seed: int = 1234,
                           device="cpu", pin: bool = False):
    """byt5 ~ N(0,1) [B,L,E], clip ~ N(0,1) [B,1024]; uncond = zeros of the
    same shapes (and ``clip_image=None`` on the unconditional side, as in the notebook)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    cond = {"byt5": torch.randn(batch, byt5_len, byt5_embd, generator=g)}
    if with_clip:
        cond["clip"] = torch.randn(batch, clip_embd, generator=g)
    if with_clip_image:
        cond["clip_image"] = torch.randn(batch, clip_embd, generator=g)
    uncond = {"byt5": torch.zeros_like(cond["byt5"])}
    if with_clip:
        uncond["clip"] = torch.zeros_like(cond["clip"])
    if pin:
        cond = {k: v.pin_memory() for k, v in cond.items()}
        uncond = {k: v.pin_memory() for k, v in uncond.items()}
    if str(device) != "cpu":
        cond = {k: v.to(device, non_blocking=True) for k, v in cond.items()}
        uncond = {k: v.to(device, non_blocking=True) for k, v in uncond.items()}
    return cond, uncond
