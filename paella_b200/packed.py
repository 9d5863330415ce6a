"""On-disk form of the library's packed weight blobs: ``tools/pack_checkpoint.py`` converts the reference's
checkpoints (``paella_v3.pt`` -- a bare state dict, nb:178-180; ``vqgan_f4.pt`` -- ``{'state_dict': ...}``, ref/src/utils.py:26,
nb:157) once; ``Paella.from_packed`` / ``VQModel.from_packed`` then start from one H2D copy.

File = ``torch.save`` of {format, kind, abi, config, nbytes, checksum, blob(uint8, CPU)}.  The blob layout is private to one
library ABI version (include/paella_b200.h PB200_ABI_VERSION) and plan; both are checked at load time.
"""
from __future__ import annotations

import torch

from ._lib import PaellaB200Error, lib

FORMAT = "paella_b200.packed/1"


def _checksum(blob_cpu: torch.Tensor) -> int:
    n8 = blob_cpu.numel() // 8 * 8
    return int(blob_cpu[:n8].view(torch.int64).sum()) + int(blob_cpu[n8:].to(torch.int64).sum())


def save_blob(path: str, kind: str, config: dict, blob: torch.Tensor) -> None:
    b = blob.detach().to("cpu").contiguous()
    torch.save({"format": FORMAT, "kind": kind, "abi": int(lib().pb200_abi_version()), "config": config, "nbytes": b.numel(),
                "checksum": _checksum(b), "blob": b}, path)


def load_blob(path: str, kind: str, device):
    d = torch.load(path, map_location="cpu", weights_only=False)
    if not isinstance(d, dict) or d.get("format") != FORMAT:
        raise PaellaB200Error(f"{path}: not a {FORMAT} file (pack the checkpoint with tools/pack_checkpoint.py)")
    if d["kind"] != kind:
        raise PaellaB200Error(f"{path}: holds a packed '{d['kind']}', expected '{kind}'")
    if d["abi"] != int(lib().pb200_abi_version()):
        raise PaellaB200Error(f"{path}: packed for library ABI {d['abi']}, this build is ABI {lib().pb200_abi_version()}: re-pack")
    b = d["blob"]
    if b.numel() != d["nbytes"] or _checksum(b) != d["checksum"]:
        raise PaellaB200Error(f"{path}: blob is truncated or corrupt (size/checksum mismatch)")
    dev = torch.device(device)
    if dev.type != "cuda":
        raise PaellaB200Error("packed models load onto CUDA devices only (no CPU fallback)")
    return d["config"], b.to(dev)
