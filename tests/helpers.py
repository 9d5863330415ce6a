"""Shared test helpers: golden fixture loading, config parsing."""
import ast
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    data = {k: z[k] for k in z.files}
    sd = {k[3:]: torch.from_numpy(v.copy()) for k, v in data.items() if k.startswith("sd:")}
    rest = {k: v for k, v in data.items() if not k.startswith("sd:")}
    cfg = ast.literal_eval(str(rest.pop("cfg_json"))) if "cfg_json" in rest else None
    return cfg, sd, rest


def t(a):
    return torch.from_numpy(np.asarray(a).copy())


def oracle_cfg(cfg_dict):
    from oracle.paella_oracle import PaellaConfig
    d = {k: v for k, v in cfg_dict.items() if k != "dropout"}
    return PaellaConfig(**d)
