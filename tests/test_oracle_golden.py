"""Pin the CPU oracle (oracle/) to the golden vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import torch

from helpers import load_golden, oracle_cfg, t
from oracle import paella_oracle as po
from oracle import vqgan_oracle as vo

TOL = dict(rtol=2e-4, atol=2e-5)


def test_r_and_c_embeddings():
    cfg, sd, g = load_golden("paella_tiny.npz")
    oc = oracle_cfg(cfg)
    torch.testing.assert_close(po.r_embedding(t(g["r"]), oc.c_r), t(g["r_embed"]), rtol=1e-5, atol=1e-6)
    ce = po.c_embeddings(sd, oc, t(g["byt5"]), t(g["clip"]), t(g["clip_image"]))
    torch.testing.assert_close(ce, t(g["c_embed"]), **TOL)


def test_forward_logits_and_taps():
    cfg, sd, g = load_golden("paella_tiny.npz")
    oc = oracle_cfg(cfg)
    taps = {}
    lg = po.paella_forward(sd, oc, t(g["x"]), t(g["r"]), t(g["byt5"]), t(g["clip"]), t(g["clip_image"]), taps=taps)
    torch.testing.assert_close(lg, t(g["logits"]), **TOL)
    n = 0
    for k, v in g.items():
        if k.startswith("tap:"):
            got = taps[k[4:]].permute(0, 3, 1, 2)
            torch.testing.assert_close(got, t(v), **TOL)
            n += 1
    assert n >= 10


def test_forward_conditioning_variants():
    cfg, sd, g = load_golden("paella_tiny.npz")
    oc = oracle_cfg(cfg)
    a = (t(g["x"]), t(g["r"]), t(g["byt5"]))
    torch.testing.assert_close(po.paella_forward(sd, oc, *a, t(g["clip"])), t(g["logits_noimg"]), **TOL)
    torch.testing.assert_close(po.paella_forward(sd, oc, *a), t(g["logits_byt5only"]), **TOL)


def test_attn_weights_variant():
    cfg, sd, g = load_golden("paella_tiny.npz")
    oc = oracle_cfg(cfg)
    a = (t(g["x"]), t(g["r"]), t(g["byt5"]), t(g["clip"]), t(g["clip_image"]))
    torch.testing.assert_close(po.paella_forward(sd, oc, *a, attn_weights=t(g["attn_weights"])), t(g["logits_attnw"]), **TOL)
    torch.testing.assert_close(po.paella_forward(sd, oc, *a), t(g["logits_custom_mha"]), **TOL)


def test_add_noise():
    cfg, sd, g = load_golden("paella_tiny.npz")
    out, mask = po.add_noise(t(g["x"]), t(g["an_t"]), t(g["an_random_x"]), t(g["an_u"]))
    assert torch.equal(out, t(g["an_out"])) and torch.equal(mask, t(g["an_mask"]))


def test_sample_loop_tokens():
    cfg, sd, _ = load_golden("paella_tiny.npz")
    _, _, s = load_golden("sample_tiny.npz")
    oc = oracle_cfg(cfg)
    byt5, clip = t(s["byt5"]), t(s["clip"])
    draws = {"init": t(s["init"]), "q": [t(q) for q in s["q"]], "u": [t(u) for u in s["u"]]}
    toks = po.sample(sd, oc, {"byt5": byt5, "clip": clip}, tuple(s["init"].shape),
                     {"byt5": torch.zeros_like(byt5), "clip": torch.zeros_like(clip)},
                     steps=int(s["steps"]), renoise_steps=int(s["renoise_steps"]), temperature=(1.0, 0.2),
                     cfg_scale=8.0, draws=draws)
    assert torch.equal(toks, t(s["tokens"]))


def test_vqgan_conv_stacks():
    cfg, sd, g = load_golden("vqgan_tiny.npz")
    lat = vo.encode_latents(sd, t(g["img"]))
    torch.testing.assert_close(lat.permute(0, 3, 1, 2), t(g["latents"]), **TOL)
    nb = cfg["bottleneck_blocks"]
    dec = vo.decode_latents(sd, sd["vquantizer.codebook.weight"][t(g["idx_rand"])], n_bottleneck=nb)
    torch.testing.assert_close(dec, t(g["dec_rand"]), **TOL)
    dec2 = vo.decode_latents(sd, (t(g["qe"]) * 0.3764).permute(0, 2, 3, 1), n_bottleneck=nb)
    torch.testing.assert_close(dec2, t(g["dec"]), **TOL)


def test_vq_nearest_definitions_agree():
    """C oracle (exact fmaf) == fp64 emulation == brute force argmin on well-separated data."""
    g = torch.Generator().manual_seed(0)
    cb = torch.randn(512, 4, generator=g)
    x = torch.randn(4096, 4, generator=g)
    a = vo.vq_nearest(x, cb)
    d = ((x[:, None, :].double() - cb[None].double()) ** 2).sum(-1)
    top2 = torch.topk(d, 2, dim=1, largest=False)
    safe = (top2.values[:, 1] - top2.values[:, 0]) > 1e-5
    assert torch.equal(a[safe], top2.indices[safe, 0])
    assert safe.float().mean() > 0.99
    # tie-break: duplicate codes -> lowest index wins
    cb2 = torch.cat([cb[:8], cb[:8]])
    assert int(vo.vq_nearest(cb[:8].clone(), cb2).max()) < 8
    lib = vo._c_oracle()
    if lib:
        saved, vo._C_LIB = vo._C_LIB, False
        try:
            b = vo.vq_nearest(x, cb)
        finally:
            vo._C_LIB = saved
        assert torch.equal(a, b)
