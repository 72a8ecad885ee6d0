"""Shared helpers for the tests: golden-fixture loading and error metrics."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def golden_problem(name, dtype=torch.float32):
    """Rebuilds the synth-style problem dict from a render_* fixture (+ the shared net weights)."""
    g = load_npz(name)
    nets = load_npz("nets_seed0.npz")
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
    prob = {"coarse": {}, "nerf_skin": {}, "rays": {}, "num_bones": 25}
    for k, v in nets.items():
        net, key = k.split(".", 1)
        prob[net][key] = t(v)
    for k, v in g.items():
        if k.startswith("in.rays."):
            prob["rays"][k[len("in.rays."):]] = t(v)
        elif k.startswith("in."):
            prob[k[3:]] = t(v)
    return prob, g


def max_abs(a, b):
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64) if not torch.is_tensor(a) else a.detach().double().cpu()
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64) if not torch.is_tensor(b) else b.detach().double().cpu()
    return float((a - b).abs().max())


def rel_err(a, b):
    """max |a-b| / max |b| -- error relative to the tensor's own scale."""
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64) if not torch.is_tensor(a) else a.detach().double().cpu()
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64) if not torch.is_tensor(b) else b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
