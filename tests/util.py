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


class ReplayRng:
    """Feeds the random draws recorded in a golden fixture (``rng.<i>``, in the reference's call order) to the code
    under test: torch.rand / randn / rand_like / randn_like return the recorded tensors, moved to ``device``, and the
    shapes are checked so that a missing or re-ordered draw fails loudly."""
    NAMES = ("rand", "randn", "rand_like", "randn_like")

    def __init__(self, g, device):
        self.draws = []
        i = 0
        while "rng.%d" % i in g:
            self.draws.append(torch.from_numpy(np.asarray(g["rng.%d" % i])))
            i += 1
        self.device, self.i = device, 0

    def _next(self, shape, dtype=None):
        assert self.i < len(self.draws), "the code under test draws more random tensors than the reference did"
        t = self.draws[self.i]
        self.i += 1
        if shape is not None:
            assert tuple(t.shape) == tuple(shape), "draw %d: reference shape %s, requested %s" % (self.i - 1, tuple(t.shape), tuple(shape))
        return t.to(self.device).to(dtype or torch.float32)

    def __enter__(self):
        self.orig = {n: getattr(torch, n) for n in self.NAMES}

        def sized(*a, **k):
            shape = a[0] if len(a) == 1 and isinstance(a[0], (tuple, list, torch.Size)) else a
            return self._next(shape, k.get("dtype"))

        def like(x, **k):
            return self._next(x.shape, x.dtype)
        torch.rand, torch.randn, torch.rand_like, torch.randn_like = sized, sized, like, like
        return self

    def __exit__(self, *exc):
        for n in self.NAMES:
            setattr(torch, n, self.orig[n])
        return False


def dump_table(name, lines):
    """Per-tensor error tables of the parity tests end up under $MODA_PARITY_DIR (default gpurun_out/parity) so that a
    GPU session can bring them back for profiles/."""
    d = os.environ.get("MODA_PARITY_DIR", os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out", "parity"))
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name + ".txt"), "w") as fh:
            fh.write("\n".join(lines) + "\n")
    except OSError:
        pass


def fixture_problem(name, nets=("coarse", "nerf_skin"), dtype=torch.float32):
    """Problem dict from a round-2 fixture that carries its own network weights (``net.<name>.<key>``)."""
    g = load_npz(name)
    t = lambda a: torch.from_numpy(np.asarray(a)).to(dtype)
    prob = {n: {} for n in nets}
    prob["rays"], prob["num_bones"] = {}, 25
    for k, v in g.items():
        if k.startswith("net."):
            _, net, key = k.split(".", 2)
            if net in prob:
                prob[net][key] = t(v)
        elif k.startswith("in.rays."):
            prob["rays"][k[len("in.rays."):]] = t(v)
        elif k.startswith("in."):
            prob[k[3:]] = t(v)
    return prob, g
