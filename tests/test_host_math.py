"""CPU check of the hand-derived adjoints in moda_b200/csrc/moda_math.h (the per-sample math the CUDA
kernels execute) against the oracle's autograd.  The header is compiled for the host with g++ into a
test-only library (tests/host_math_harness.cpp); the product never loads it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from moda_b200 import synth
from oracle import restated as O

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(HERE, "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libhostmath.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "moda_b200", "csrc"),
                           os.path.join(HERE, "host_math_harness.cpp"), "-o", so])
    return ctypes.CDLL(so)


def fp(t):
    if t is None:
        return None
    assert t.dtype == torch.float32 and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


def _problem(R=6, S=16, seed=0):
    sp = synth.make_skin_problem(R, S, seed=seed)
    gen = torch.Generator().manual_seed(seed + 100)
    dskin = 0.5 * torch.randn(R, S, 25, generator=gen)
    gy = torch.randn(R, S, 3, generator=gen)
    return sp, dskin, gy, gen


def _run_host(lib, pts, bones, rts, aux, dskin, skin_in, gy, gskin, per_ray, deform, invert, want_y=True):
    R, S, _ = pts.shape
    B = bones.shape[-2]
    y = torch.zeros(R, S, 3) if want_y else None
    skin = torch.zeros(R, S, B) if skin_in is None else None
    lib.h_skin_warp_fwd(fp(pts), fp(bones), fp(rts), fp(aux), fp(dskin), fp(skin_in), fp(y), fp(skin), R, S, B,
                        per_ray, deform, invert)
    g = dict(pts=torch.zeros(R, S, 3), dskin=torch.zeros(R, S, B) if dskin is not None else None,
             skin_in=torch.zeros(R, S, B) if skin_in is not None else None,
             rts=torch.zeros(R, B, 8) if rts is not None else None, bones=torch.zeros_like(bones), aux=torch.zeros(2))
    lib.h_skin_warp_bwd(fp(pts), fp(bones), fp(rts), fp(aux), fp(dskin), fp(skin_in), fp(gy), fp(gskin),
                        fp(g["pts"]), fp(g["dskin"]), fp(g["skin_in"]), fp(g["rts"]), fp(g["bones"]), fp(g["aux"]),
                        R, S, B, per_ray, deform, invert)
    return y, skin, g


def _close(a, b, tol, name):
    a, b = a.double(), b.double()
    err = float((a - b).abs().max())
    scale = max(1.0, float(b.abs().max()))
    assert err <= tol * scale, "%s: err %.3e scale %.3e" % (name, err, scale)


@pytest.mark.parametrize("mode", ["backward_warp", "forward_warp"])
def test_fused_warp_against_oracle_autograd(lib, mode):
    sp, dskin, gy, _ = _problem()
    bw = mode == "backward_warp"
    pts, bones, aux = sp["xyz"].contiguous(), sp["bones_rst"].contiguous(), sp["skin_aux"].contiguous()
    rts = sp["bone_rts"].view(-1, 25, 8).contiguous()
    y, _, g = _run_host(lib, pts, bones, rts, aux, dskin, None, gy, None, 0, int(bw), int(bw))
    # oracle in float64
    P = [t.double().requires_grad_(True) for t in (pts, bones, rts, aux, dskin)]
    p64, b64, r64, a64, d64 = P
    if bw:
        bdfm = O.bone_transform(b64, r64)
        w = O.skinning(bdfm, p64, d64, a64)
    else:
        w = O.skinning(b64, p64, d64, a64)
    yo, _ = O.neu_dbs(b64, r64, w, p64, backward=bw)
    (yo * gy.double()).sum().backward()
    _close(y, yo.detach(), 1e-5, "y")  # fp32 noise through the peaked softmax: the fp32 oracle itself is 4e-6 from fp64
    _close(g["pts"], p64.grad, 2e-4, "gpts")
    _close(g["dskin"], d64.grad, 2e-4, "gdskin")
    _close(g["rts"], r64.grad, 2e-4, "grts")
    _close(g["bones"], b64.grad, 2e-4, "gbones")
    _close(g["aux"][:1], a64.grad[:1], 2e-4, "gaux")


def test_skinning_only_and_blend_only(lib):
    sp, dskin, gy, gen = _problem(seed=2)
    pts, aux = sp["xyz"].contiguous(), sp["skin_aux"].contiguous()
    rts = sp["bone_rts"].view(-1, 25, 8).contiguous()
    bones_dfm = O.bone_transform(sp["bones_rst"], rts).contiguous()
    gskin = torch.randn(6, 16, 25, generator=gen)
    # skinning(bones (R,B,10), pts, dskin) -> skin, gradient arriving on the weights
    _, skin, g = _run_host(lib, pts, bones_dfm, None, aux, dskin, None, None, gskin, 1, 0, 0, want_y=False)
    P = [t.double().requires_grad_(True) for t in (pts, bones_dfm, aux, dskin)]
    w = O.skinning(P[1], P[0], P[3], P[2])
    (w * gskin.double()).sum().backward()
    _close(skin, w.detach(), 3e-5, "skin")
    _close(g["pts"], P[0].grad, 2e-4, "gpts")
    _close(g["bones"], P[1].grad, 2e-4, "gbones")
    _close(g["aux"][:1], P[2].grad[:1], 2e-4, "gaux")
    _close(g["dskin"], P[3].grad, 2e-4, "gdskin")
    # dqs_blend_skinning(dq, skin, pts) with given weights
    skin_in = w.detach().float().contiguous()
    y, _, g = _run_host(lib, pts, sp["bones_rst"].contiguous(), rts, aux, None, skin_in, gy, None, 0, 0, 0)
    Q = [t.double().requires_grad_(True) for t in (pts, rts, skin_in)]
    yo = O.dqs_blend_skinning(Q[1], Q[2], Q[0])
    (yo * gy.double()).sum().backward()
    _close(y, yo.detach(), 1e-5, "blend y")
    _close(g["pts"], Q[0].grad, 2e-4, "blend gpts")
    _close(g["rts"], Q[1].grad, 2e-4, "blend gdq")
    _close(g["skin_in"], Q[2].grad, 2e-4, "blend gskin")


def test_bone_transform_and_density(lib):
    sp, _, _, gen = _problem(seed=3)
    bones = sp["bones_rst"].contiguous()
    rts = (sp["bone_rts"].view(-1, 25, 8) * 1.3).contiguous()  # non-unit real parts on purpose
    R, B = rts.shape[0], 25
    out = torch.zeros(R, B, 10)
    lib.h_bone_transform_fwd(fp(bones), fp(rts), fp(out), R, B)
    b64, r64 = bones.double().requires_grad_(True), rts.double().requires_grad_(True)
    oo = O.bone_transform(b64, r64)
    go = torch.randn(R, B, 10, generator=gen)
    (oo * go.double()).sum().backward()
    gb, gr = torch.zeros(B, 10), torch.zeros(R, B, 8)
    lib.h_bone_transform_bwd(fp(bones), fp(rts), fp(go), fp(gb), fp(gr), R, B)
    _close(out, oo.detach(), 2e-6, "bone_transform")
    _close(gb, b64.grad, 2e-5, "gbones")
    _close(gr, r64.grad, 2e-5, "grts")
    # density -> alpha and its partials (rendering.py:199-207)
    n = 64
    sig = torch.randn(n, generator=gen) * 0.2
    sig[0] = 0.0
    delta = torch.rand(n, generator=gen) * 0.01 + 1e-4
    beta = torch.tensor([0.1], dtype=torch.float64, requires_grad=True)
    s64, d64 = sig.double().requires_grad_(True), delta.double().requires_grad_(True)
    al = O.density_to_alpha(s64, d64, beta)
    al.sum().backward()
    a, ds, dib, dd = (torch.zeros(n) for _ in range(4))
    ib = 1.0 / (0.1 + 1e-9)
    ctypes.CDLL  # keep import used
    lib.h_density_alpha(fp(sig), fp(delta), ctypes.c_float(ib), fp(a), fp(ds), fp(dib), fp(dd), n)
    _close(a, al.detach(), 2e-6, "alpha")
    _close(ds, s64.grad, 2e-5, "da/dsigma")
    _close(dd, d64.grad, 2e-5, "da/ddelta")
    # d/d beta = sum(da/dib) * d ib/d beta
    gbeta = float(dib.double().sum()) * (-ib * ib)
    assert abs(gbeta - float(beta.grad)) <= 2e-4 * max(1.0, abs(float(beta.grad)))
