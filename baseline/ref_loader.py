"""Import the unmodified reference from baseline/_ref (see install_ref.py) for the baseline arms of bench.py.

`torchtools` (third-party, unpinned, absent offline — SURVEY.md F4) is stubbed so that ref/src/vqgan.py imports; the
denoiser + sample() path that the baseline arms time never touches the stub.  The stub's quantiser is the published
torchtools form (expanded squared-L2 via addmm, first minimum) in plain torch ops.
"""
import importlib.util
import os
import sys
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "src", "modules.py"))


def _install_torchtools_stub():
    if "torchtools.nn" in sys.modules:
        return

    class VectorQuantize(nn.Module):
        def __init__(self, embedding_size, k, ema_decay=0.99, ema_loss=False):
            super().__init__()
            self.codebook = nn.Embedding(k, embedding_size)
            self.codebook.weight.data.uniform_(-1. / k, 1. / k)

        def forward(self, x, get_losses=True, dim=-1):
            if dim != -1:
                x = x.movedim(dim, -1)
            flat = x.reshape(-1, x.shape[-1])
            cb = self.codebook.weight
            d = torch.addmm(cb.pow(2).sum(1) + flat.pow(2).sum(1, keepdim=True), flat, cb.t(), alpha=-2.0, beta=1.0)
            idx = d.min(dim=1)[1]
            zq = cb.index_select(0, idx).view(x.shape)
            loss = (zq - x).pow(2).mean()
            if dim != -1:
                zq = zq.movedim(-1, dim)
            return zq, (loss, loss), idx.view(x.shape[:-1])

        def idx2vq(self, idx, dim=-1):
            q = self.codebook(idx)
            return q.movedim(-1, dim) if dim != -1 else q

    tt, ttnn = types.ModuleType("torchtools"), types.ModuleType("torchtools.nn")
    ttnn.VectorQuantize = VectorQuantize
    tt.nn = ttnn
    sys.modules["torchtools"], sys.modules["torchtools.nn"] = tt, ttnn


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load():
    """-> namespace(modules=ref/src/modules.py, nb_modules=ref/utils/modules.py, alter=..., vqgan=..., sample=ref sample())."""
    if not available():
        raise FileNotFoundError("baseline/_ref is empty: run `python baseline/install_ref.py` where /root/reference exists")
    _install_torchtools_stub()
    mods = _load("ref_src_modules", "src/modules.py")
    nb = _load("ref_utils_modules", "utils/modules.py")
    alter = _load("ref_alter_attention", "utils/alter_attention.py")
    vq = _load("vqgan", "src/vqgan.py")            # ref/src/utils.py does `from vqgan import VQModel`
    try:
        sample = _load("ref_src_utils", "src/utils.py").sample
    except Exception:      # noqa: BLE001 — torchvision / transformers import trouble: exec the function source alone
        src = open(os.path.join(REF, "src", "utils.py")).read()
        ns = {"torch": torch}
        exec(compile(src[src.index("def sample("):], "ref_src_utils_sample", "exec"), ns)
        sample = ns["sample"]
    return types.SimpleNamespace(modules=mods, nb_modules=nb, alter=alter, vqgan=vq, sample=sample)
