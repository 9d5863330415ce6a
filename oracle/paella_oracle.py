"""CPU oracle for the Paella denoiser forward and the sample() loop.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the
product (``paella_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline legs may use it.

This is a *restatement* of the reference algorithm as plain functions over a
state-dict (same key names as the reference modules), written channels-last
with explicit matrix products, so that every step of the CUDA path has a CPU
counterpart with the same intermediate tensors.  It is pinned against the real
reference (``/root/reference/src/modules.py``, imported in the dev container)
by ``tests/golden/make_golden.py`` -> ``tests/golden/*.npz`` and by
``tests/test_oracle_vs_reference.py`` (skipped where the reference tree is
absent, i.e. on the GPU box).

Reference lines restated here (``ref`` = dome272/Paella @ e1ab72b):
  gen_r_embedding      ref/src/modules.py:212-221
  gen_c_embeddings     ref/src/modules.py:223-232 (+ list clip_image, ref/utils/modules.py:228-235)
  in_mapper/embedding  ref/src/modules.py:126-134,271
  ResBlock             ref/src/modules.py:43-62   (LayerNorm2d :22-27, GlobalResponseNorm :30-40)
  TimestepBlock        ref/src/modules.py:99-106
  AttnBlock            ref/src/modules.py:65-79   (Attention2D :7-19; attn_weights: ref/utils/alter_attention.py:19-36)
  FeedForwardBlock     ref/src/modules.py:82-96
  level resamplers     ref/src/modules.py:153-156,172-175
  clf / out_mapper     ref/src/modules.py:179-187
  forward              ref/src/modules.py:263-275
  add_noise            ref/src/modules.py:277-283
  sample               ref/src/utils.py:35-55, ref/src_distributed/utils.py:97-126, notebook cell 3
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


@dataclass
class PaellaConfig:
    """Constructor arguments of the reference ``Paella`` (ref/src/modules.py:110-112)."""
    c_in: int = 256
    c_out: int = 256
    num_labels: int = 8192
    c_r: int = 64
    patch_size: int = 2
    c_cond: int = 1024
    c_hidden: Sequence[int] = (640, 1280, 1280)
    nhead: Sequence[int] = (-1, 16, 16)
    blocks: Sequence[int] = (6, 16, 6)
    level_config: Sequence[str] = ("CT", "CTA", "CTA")
    clip_embd: int = 1024
    byt5_embd: int = 1536
    clip_seq_len: int = 4
    kernel_size: int = 3
    self_attn: bool = True


# --------------------------------------------------------------------------
# matrix product hook: fp32, or fp16-rounded operands with fp32 accumulation
# (predicts the error budget of the tensor-core path on CPU).
# --------------------------------------------------------------------------
def mm_fp32(a: Tensor, w: Tensor) -> Tensor:
    """a [..., K] @ w[N, K]^T in fp32."""
    return a @ w.t()


def mm_f16_operands(a: Tensor, w: Tensor) -> Tensor:
    return a.half().float() @ w.half().float().t()


def ln(x: Tensor, eps: float = 1e-6) -> Tensor:
    """LayerNorm over the last dim, no affine (biased variance), fp32."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def gelu_erf(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def r_embedding(r: Tensor, c_r: int, max_positions: int = 10000) -> Tensor:
    r = r * max_positions
    half = c_r // 2
    k = math.log(max_positions) / (half - 1)
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=r.device) * (-k))
    ang = r[:, None] * freq[None, :]
    emb = torch.cat([ang.sin(), ang.cos()], dim=1)
    if c_r % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


def c_embeddings(sd, cfg: PaellaConfig, byt5, clip, clip_image, mm=mm_fp32) -> Tensor:
    seq = mm(byt5, sd["byt5_mapper.weight"]) + sd["byt5_mapper.bias"]
    B = byt5.shape[0]
    if clip is not None:
        c = mm(clip, sd["clip_mapper.weight"]) + sd["clip_mapper.bias"]
        seq = torch.cat([seq, c.view(B, -1, cfg.c_cond)], dim=1)
    if clip_image is not None:
        cis = clip_image if isinstance(clip_image, (list, tuple)) else [clip_image]
        for ci in cis:
            c = mm(ci, sd["clip_image_mapper.weight"]) + sd["clip_image_mapper.bias"]
            seq = torch.cat([seq, c.view(B, -1, cfg.c_cond)], dim=1)
    return ln(seq)


def dwconv3x3_nhwc(x: Tensor, w: Tensor, b: Tensor, x_skip: Optional[Tensor] = None) -> Tensor:
    """Depthwise k×k conv, zero padding k//2, channels-last.

    ``w`` is the reference weight [c, 1|2, k, k].  With a skip tensor the
    reference concatenates [x, skip] on channels and uses groups=c, so output
    channel g reads concatenated channels 2g and 2g+1 (ref/src/modules.py:46,59).
    """
    B, H, W, c = x.shape
    k = w.shape[-1]
    p = k // 2
    src = x if x_skip is None else torch.cat([x, x_skip], dim=-1)
    srcp = F.pad(src, (0, 0, p, p, p, p))
    out = torch.zeros(B, H, W, c, dtype=torch.float32, device=x.device) + b
    per = w.shape[1]
    for j in range(per):
        chan = torch.arange(c, device=x.device) * per + j
        sj = srcp[..., chan]
        for ky in range(k):
            for kx in range(k):
                out = out + sj[:, ky:ky + H, kx:kx + W, :] * w[:, j, ky, kx]
    return out


def grn(h: Tensor, gamma: Tensor, beta: Tensor) -> Tensor:
    """GlobalResponseNorm on [B,H,W,C] (ref/src/modules.py:37-40)."""
    gx = torch.sqrt((h * h).sum(dim=(1, 2), keepdim=True))
    nx = gx / (gx.mean(dim=-1, keepdim=True) + 1e-6)
    return gamma.view(1, 1, 1, -1) * (h * nx) + beta.view(1, 1, 1, -1) + h


def mlp_grn(xn: Tensor, sd, pre: str, mm) -> Tensor:
    h = mm(xn, sd[pre + "channelwise.0.weight"]) + sd[pre + "channelwise.0.bias"]
    h = gelu_erf(h)
    h = grn(h, sd[pre + "channelwise.2.gamma"], sd[pre + "channelwise.2.beta"])
    return mm(h, sd[pre + "channelwise.4.weight"]) + sd[pre + "channelwise.4.bias"]


def resblock(x: Tensor, sd, pre: str, x_skip=None, mm=mm_fp32) -> Tensor:
    d = dwconv3x3_nhwc(x, sd[pre + "depthwise.weight"], sd[pre + "depthwise.bias"], x_skip)
    return x + mlp_grn(ln(d), sd, pre, mm)


def feedforward_block(x: Tensor, sd, pre: str, mm=mm_fp32) -> Tensor:
    return x + mlp_grn(ln(x), sd, pre, mm)


def timestep_block(x: Tensor, r_embed: Tensor, sd, pre: str) -> Tensor:
    ab = r_embed @ sd[pre + "mapper.weight"].t() + sd[pre + "mapper.bias"]
    c = x.shape[-1]
    a, b = ab[:, :c], ab[:, c:]
    return x * (1 + a[:, None, None, :]) + b[:, None, None, :]


def attention_core(q, k, v, nhead: int, attn_weights: Optional[Tensor] = None) -> Tensor:
    """q [B,Nq,E], k/v [B,Nk,E] -> [B,Nq,E]; softmax(qk^T/sqrt(hd)) v per head.

    ``attn_weights`` (1-D, len n) scales the softmax output of the LAST n key
    columns, after the softmax, without renormalising (ref/utils/alter_attention.py:23-34).
    """
    B, Nq, E = q.shape
    Nk = k.shape[1]
    hd = E // nhead
    qh = q.view(B, Nq, nhead, hd).permute(0, 2, 1, 3)
    kh = k.view(B, Nk, nhead, hd).permute(0, 2, 1, 3)
    vh = v.view(B, Nk, nhead, hd).permute(0, 2, 1, 3)
    s = (qh @ kh.transpose(-2, -1)) / (hd ** 0.5)
    p = torch.softmax(s, dim=-1)
    if attn_weights is not None:
        w = torch.ones(Nk, device=q.device)
        w[-attn_weights.shape[0]:] = attn_weights.float()
        p = p * w
    o = p @ vh
    return o.permute(0, 2, 1, 3).reshape(B, Nq, E)


def attn_block(x: Tensor, c_embed: Tensor, sd, pre: str, nhead: int, self_attn: bool,
               attn_weights=None, mm=mm_fp32) -> Tensor:
    B, H, W, c = x.shape
    kv = mm(F.silu(c_embed), sd[pre + "kv_mapper.1.weight"]) + sd[pre + "kv_mapper.1.bias"]
    xn = ln(x).reshape(B, H * W, c)
    kvs = torch.cat([xn, kv], dim=1) if self_attn else kv
    w_in = sd[pre + "attention.attn.in_proj_weight"]
    b_in = sd[pre + "attention.attn.in_proj_bias"]
    q = mm(xn, w_in[:c]) + b_in[:c]
    k = mm(kvs, w_in[c:2 * c]) + b_in[c:2 * c]
    v = mm(kvs, w_in[2 * c:]) + b_in[2 * c:]
    o = attention_core(q, k, v, nhead, attn_weights)
    o = mm(o, sd[pre + "attention.attn.out_proj.weight"]) + sd[pre + "attention.attn.out_proj.bias"]
    return x + o.view(B, H, W, c)


def patchify2(x: Tensor) -> Tensor:
    """[B,H,W,C] -> [B,H/2,W/2,(dy,dx,C)]."""
    B, H, W, C = x.shape
    x = x.view(B, H // 2, 2, W // 2, 2, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, H // 2, W // 2, 4 * C)


def unpatchify2(y: Tensor, cout: int) -> Tensor:
    """[B,h,w,(dy,dx,cout)] -> [B,2h,2w,cout]."""
    B, h, w, _ = y.shape
    y = y.view(B, h, w, 2, 2, cout).permute(0, 1, 3, 2, 4, 5)
    return y.reshape(B, 2 * h, 2 * w, cout)


def down_resample(x: Tensor, w: Tensor, b: Tensor, mm=mm_fp32) -> Tensor:
    """LN2d + Conv2d(k=2,s=2) as a patchify GEMM.  w: [Cout,Cin,2,2]."""
    cout = w.shape[0]
    wm = w.permute(0, 2, 3, 1).reshape(cout, -1)      # [(Cout), (dy,dx,Cin)]
    return mm(patchify2(ln(x)), wm) + b


def up_resample(x: Tensor, w: Tensor, b: Tensor, mm=mm_fp32) -> Tensor:
    """LN2d + ConvTranspose2d(k=2,s=2) as an un-patchify GEMM.  w: [Cin,Cout,2,2]."""
    cin, cout = w.shape[0], w.shape[1]
    wm = w.permute(2, 3, 1, 0).reshape(4 * cout, cin)  # [(dy,dx,Cout), Cin]
    y = mm(ln(x), wm) + b.repeat(4)
    return unpatchify2(y, cout)


def embed_tokens(tokens: Tensor, sd, cfg: PaellaConfig, mm=mm_fp32) -> Tensor:
    """in_mapper + embedding: [B,H,W] i64 -> [B,H/2,W/2,c0] (ref/src/modules.py:126-134,271)."""
    e = ln(sd["in_mapper.0.weight"][tokens])                      # [B,H,W,c_in]
    B, H, W, C = e.shape
    ps = cfg.patch_size
    # PixelUnshuffle: out channel = c*ps*ps + dy*ps + dx
    e = e.view(B, H // ps, ps, W // ps, ps, C).permute(0, 1, 3, 5, 2, 4).reshape(B, H // ps, W // ps, C * ps * ps)
    w = sd["embedding.1.weight"].reshape(cfg.c_hidden[0], -1)
    return ln(mm(e, w) + sd["embedding.1.bias"])


def classifier_features(x: Tensor, sd, cfg: PaellaConfig, mm=mm_fp32) -> Tensor:
    """clf + out_mapper's LayerNorm: [B,h,w,c0] -> [B,H,W,c_out] (ref/src/modules.py:179-185)."""
    B, h, w, _ = x.shape
    ps = cfg.patch_size
    y = mm(ln(x), sd["clf.1.weight"].reshape(cfg.c_out * ps * ps, -1)) + sd["clf.1.bias"]
    # PixelShuffle: in channel = c*ps*ps + dy*ps + dx
    y = y.view(B, h, w, cfg.c_out, ps, ps).permute(0, 1, 4, 2, 5, 3).reshape(B, h * ps, w * ps, cfg.c_out)
    return ln(y)


def iter_blocks(cfg: PaellaConfig):
    """Yield (prefix, kind, level, c, c_skip) in execution order — mirrors the
    ModuleList construction at ref/src/modules.py:149-176."""
    L = len(cfg.c_hidden)
    for i in range(L):
        j = 0
        if i > 0:
            yield (f"down_blocks.{i}.{j}.", "down", i, cfg.c_hidden[i], 0)
            j += 1
        for _ in range(cfg.blocks[i]):
            for bt in cfg.level_config[i]:
                yield (f"down_blocks.{i}.{j}.", bt, i, cfg.c_hidden[i], 0)
                j += 1
        yield (None, "save", i, cfg.c_hidden[i], 0)
    for ui, i in enumerate(reversed(range(L))):
        j = 0
        for jj in range(cfg.blocks[i]):
            for k, bt in enumerate(cfg.level_config[i]):
                skip = cfg.c_hidden[i] if (i < L - 1 and jj == 0 and k == 0) else 0
                yield (f"up_blocks.{ui}.{j}.", bt, i, cfg.c_hidden[i], skip)
                j += 1
        if i > 0:
            yield (f"up_blocks.{ui}.{j}.", "up", i, cfg.c_hidden[i], 0)


def paella_features(sd: Dict[str, Tensor], cfg: PaellaConfig, x: Tensor, r: Tensor, byt5: Tensor,
                    clip=None, clip_image=None, attn_weights=None, mm: Callable = mm_fp32,
                    taps: Optional[dict] = None) -> Tensor:
    """Everything up to (and including) out_mapper's LayerNorm: -> [B,H,W,c_out]."""
    r_embed = r_embedding(r.float(), cfg.c_r)
    c_embed = c_embeddings(sd, cfg, byt5, clip, clip_image, mm)
    h = embed_tokens(x, sd, cfg, mm)
    if taps is not None:
        taps["r_embed"], taps["c_embed"], taps["embed"] = r_embed, c_embed, h
    saved: List[Tensor] = []
    L = len(cfg.c_hidden)
    for pre, kind, lvl, c, c_skip in iter_blocks(cfg):
        if kind == "save":
            saved.insert(0, h)
            continue
        if kind == "down":
            h = down_resample(h, sd[pre + "1.weight"], sd[pre + "1.bias"], mm)
        elif kind == "up":
            h = up_resample(h, sd[pre + "1.weight"], sd[pre + "1.bias"], mm)
        elif kind == "C":
            skip = None
            if c_skip:
                skip = saved[L - 1 - lvl]
            h = resblock(h, sd, pre, skip, mm)
        elif kind == "T":
            h = timestep_block(h, r_embed, sd, pre)
        elif kind == "A":
            h = attn_block(h, c_embed, sd, pre, cfg.nhead[lvl], cfg.self_attn, attn_weights, mm)
        elif kind == "F":
            h = feedforward_block(h, sd, pre, mm)
        else:
            raise ValueError(kind)
        if taps is not None and pre is not None:
            taps[pre] = h
    # NB: the first up level starts from level_outputs[0] == the deepest saved tensor == h already.
    return classifier_features(h, sd, cfg, mm)


def paella_forward(sd, cfg: PaellaConfig, x, r, byt5, clip=None, clip_image=None, attn_weights=None,
                   mm: Callable = mm_fp32, taps=None) -> Tensor:
    """Reference-shaped output: logits [B,num_labels,H,W] fp32 (ref/src/modules.py:263-275)."""
    a = paella_features(sd, cfg, x, r, byt5, clip, clip_image, attn_weights, mm, taps)
    w = sd["out_mapper.1.weight"].reshape(cfg.num_labels, cfg.c_out)
    return mm(a, w).permute(0, 3, 1, 2).contiguous()


def add_noise(x: Tensor, t: Tensor, random_x: Tensor, u: Tensor):
    """ref/src/modules.py:277-283 with the uniform draw ``u`` passed in."""
    mask = (u <= t[:, None, None]).long()
    return x * (1 - mask) + random_x * mask, mask


def sample(sd, cfg: PaellaConfig, model_inputs: dict, latent_shape, unconditional_inputs: dict,
           steps=12, renoise_steps=11, temperature=(1.0, 0.2), cfg_scale=8.0, t_start=1.0, t_end=0.0,
           draws: Optional[dict] = None, mm=mm_fp32):
    """ref/src/utils.py:35-55 with every random draw supplied by ``draws``:
    draws['init'] [B,H,W] i64; draws['q'][i] [B*H*W,K] Exp(1); draws['u'][i] [B,H,W] U[0,1).
    multinomial(p,1) == argmax(p/q) (verified against torch in tests)."""
    B = latent_shape[0]
    sampled = draws["init"].clone()
    init_noise = sampled.clone()
    t_list = torch.linspace(t_start, t_end, steps + 1)
    temps = torch.linspace(temperature[0], temperature[1], steps)
    for i in range(steps):
        t = torch.ones(B, device=sampled.device) * t_list[i]
        logits = paella_forward(sd, cfg, sampled, t, mm=mm, **model_inputs)
        if cfg_scale:
            logits = logits * cfg_scale + paella_forward(sd, cfg, sampled, t, mm=mm, **unconditional_inputs) * (1 - cfg_scale)
        scores = logits.div(temps[i]).softmax(dim=1)
        p = scores.permute(0, 2, 3, 1).reshape(-1, logits.size(1))
        sampled = torch.argmax(p / draws["q"][i], dim=-1).view(logits.size(0), *logits.shape[2:])
        if i < renoise_steps:
            t_next = torch.ones(B, device=sampled.device) * t_list[i + 1]
            sampled, _ = add_noise(sampled, t_next, init_noise, draws["u"][i])
    return sampled
