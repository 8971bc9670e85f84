"""Network weights for the product path: seeded random state_dicts (same shapes / names / init law as the
reference modules, torch.nn default U(-1/sqrt(H), 1/sqrt(H))) for benchmarks when no trained checkpoint is at
hand, and a loader for the reference's checkpoint format (torch.save dict with `model_state_dict`,
pepper/modules/python/models/ModelHander.py:88-110)."""
from __future__ import annotations

import numpy as np


def _u(rng, shape, bound):
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


def random_variant_state(seed: int = 0) -> dict:
    rng = np.random.default_rng(seed)
    s = {}
    for layer, inp in (("encoder", 26), ("decoder", 512)):
        b = 1.0 / np.sqrt(256)
        for suf in ("", "_reverse"):
            s[f"{layer}.weight_ih_l0{suf}"] = _u(rng, (1024, inp), b)
            s[f"{layer}.weight_hh_l0{suf}"] = _u(rng, (1024, 256), b)
            s[f"{layer}.bias_ih_l0{suf}"] = _u(rng, (1024,), b)
            s[f"{layer}.bias_hh_l0{suf}"] = _u(rng, (1024,), b)
    dims = [(512, 16896), (512, 512), (512, 512), (512, 512), (512, 512)]
    for i, (o, k) in enumerate(dims, start=1):
        s[f"linear_{i}.weight"] = _u(rng, (o, k), 1.5 / np.sqrt(k))
        s[f"linear_{i}.bias"] = _u(rng, (o,), 1.0 / np.sqrt(k))
    s["output_layer_type.weight"] = _u(rng, (3, 512), 24.0 / np.sqrt(512))
    s["output_layer_type.bias"] = _u(rng, (3,), 1.0 / np.sqrt(512))
    return s


def random_polish_state(seed: int = 0) -> dict:
    rng = np.random.default_rng(seed)
    s = {}
    for layer, inp in (("gru_encoder", 10), ("gru_decoder", 256)):
        b = 2.0 / np.sqrt(128)
        for suf in ("", "_reverse"):
            s[f"{layer}.weight_ih_l0{suf}"] = _u(rng, (384, inp), b)
            s[f"{layer}.weight_hh_l0{suf}"] = _u(rng, (384, 128), b)
            s[f"{layer}.bias_ih_l0{suf}"] = _u(rng, (384,), b / 2)
            s[f"{layer}.bias_hh_l0{suf}"] = _u(rng, (384,), b / 2)
    s["dense1.weight"] = _u(rng, (5, 256), 12.0 / np.sqrt(256))
    s["dense1.bias"] = _u(rng, (5,), 1.0 / np.sqrt(256))
    return s


def load_checkpoint(path: str) -> dict:
    """Reference `.pkl` checkpoint -> state_dict with the DataParallel 'module.' prefix stripped."""
    import torch
    ck = torch.load(path, map_location="cpu")
    sd = ck["model_state_dict"] if "model_state_dict" in ck else ck
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}
