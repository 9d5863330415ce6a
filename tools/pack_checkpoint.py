#!/usr/bin/env python
"""Convert the reference's checkpoints into the library's packed form.  Needs a GPU: packing runs the
library's own repack kernels.

    python tools/pack_checkpoint.py paella  models/paella_v3.pt  paella_v3.pb200   [--byt5-embd 2560]
    python tools/pack_checkpoint.py vqgan   models/vqgan_f4.pt   vqgan_f4.pb200

Accepts a bare state dict (paella_v3.pt, nb:178-180) or ``{'state_dict': ...}`` (vqgan_f4.pt, ref/src/utils.py:26); extra
``vquantizer.*`` buffers of the unpinned torchtools quantiser (EMA statistics) are ignored; any other unexpected / missing
key is an error (strict load).  Then:  ``Paella.from_packed(path)`` / ``VQModel.from_packed(path)``.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def load_state_dict_file(path):
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(ckpt, dict) and "state_dict" in ckpt and isinstance(ckpt["state_dict"], dict):
        ckpt = ckpt["state_dict"]
    return {k[len("module."):] if k.startswith("module.") else k: v for k, v in ckpt.items()}       # DDP-saved checkpoints


def pack(kind, src, dst, device="cuda", **ctor):
    from paella_b200.modules import Paella
    from paella_b200.vqgan import VQModel
    sd = load_state_dict_file(src)
    if kind == "paella":
        if "byt5_embd" not in ctor and "byt5_mapper.weight" in sd:
            ctor["byt5_embd"] = sd["byt5_mapper.weight"].shape[1]        # 2560 for paella_v3 (nb:177), 1536 in the ctor default
        m = Paella(**ctor)
    else:
        m = VQModel(**ctor)
    m.load_state_dict(sd, strict=True)
    m = m.eval().requires_grad_(False).to(device)
    m.pack_weights()
    m.save_packed(dst)
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kind", choices=["paella", "vqgan"])
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--byt5-embd", type=int, default=None)
    a = ap.parse_args()
    ctor = {"byt5_embd": a.byt5_embd} if (a.kind == "paella" and a.byt5_embd) else {}
    m = pack(a.kind, a.src, a.dst, **ctor)
    print(f"{a.dst}: {m._blob.numel() / 1e6:.1f} MB packed ({a.kind})")


if __name__ == "__main__":
    main()
