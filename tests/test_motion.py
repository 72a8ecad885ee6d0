"""The alternative motion models of SURVEY.md 8(f) rank 5 -- LBS (opts.lbs) and the free-form flow fields (flowbw / flowfw
as Transhead / SE3head) -- against fixtures produced by executing the real reference (oracle/make_golden_motion.py).
CPU: the oracle restatement is pinned against the fixtures.  GPU: render_rays of this package against the same
fixtures, outputs and every gradient, in the exact (fp32 SIMT) mode and in the default fp16 tensor-core mode."""
import numpy as np
import pytest
import torch

from moda_b200 import synth
from oracle import restated as O
from tests.util import dump_table, load_npz, max_abs, rel_err

OUT_KEYS = ("img_coarse", "depth_rnd", "sil_coarse", "frame_cyc_dis", "xyz_camera_vis", "xyz_canonical_vis")
FIXTURES = {"lbs": ("render_lbs_n16_fp32.npz", 5), "trans": ("render_flow_trans_n16_fp32.npz", 6),
            "se3": ("render_flow_se3_n16_fp32.npz", 6)}


def _problem(kind):
    """The problem is re-created from its seed; the fixture carries the rays and a checksum of every net."""
    name, seed = FIXTURES[kind]
    g = load_npz(name)
    prob = synth.make_motion_problem(16, kind, seed=seed)
    for k, v in g.items():
        if k.startswith("in.rays."):
            assert np.array_equal(prob["rays"][k[8:]].numpy(), v), "synth.make_motion_problem drifted from the fixture: " + k
        if k.startswith("in.checksum."):
            s = sum(float(t.double().abs().sum()) for t in prob[k[12:]].values())
            assert abs(s - float(v)) <= 1e-9 * abs(float(v)), "net weights drifted from the fixture: " + k
    return prob, g


@pytest.mark.parametrize("kind", ["lbs", "trans", "se3"])
def test_oracle_motion_models_against_reference(kind):
    prob, g = _problem(kind)
    prob = O.to_dtype(prob, torch.float32)
    leaves = O.require_grads(prob)
    res = O.render_rays(prob, n_samples=128, perturb=0.0)
    O.parity_loss(res).backward()
    for k in OUT_KEYS:
        assert max_abs(res[k], g["out." + k]) < 2e-5, k
    worst = 0.0
    for k, v in leaves.items():
        if "grad." + k not in g:
            continue
        got = v.grad if v.grad is not None else torch.zeros_like(v)
        e = rel_err(got, g["grad." + k]) if np.abs(g["grad." + k]).max() > 0 else max_abs(got, g["grad." + k])
        worst = max(worst, e)
        assert e < 2e-3, "%s: %.2e" % (k, e)
    assert any(("grad." + k) in g for k in leaves), "no gradient compared"


def test_lbs_helper_functions_against_reference():
    prob, g = _problem("lbs")
    B = prob["num_bones"]
    bt = O.bone_transform_rigid(prob["bones_rst"], prob["rays"]["bone_rts"])
    assert max_abs(bt, g["fn.bone_transform"]) < 2e-6
    v = prob["rays"]["bone_rts"].reshape(16, B, 12)
    rts = torch.cat([v[..., :9].reshape(16, B, 3, 3), v[..., 9:, None]], -1)
    assert max_abs(O.rts_invert(rts), g["fn.rts_invert"]) < 1e-6


def _gpu_run(kind, precision):
    from moda_b200 import config, models as MM
    from moda_b200.rendering import render_rays
    prob, g = _problem(kind)
    old = config.precision
    config.set_precision(precision)
    try:
        models, emb, rays = MM.build_motion_models(prob, "cuda")
        for m in models.values():
            if isinstance(m, torch.nn.Module):
                m.train()
        opts = synth.default_opts()
        if kind == "lbs":
            opts.lbs, opts.neudbs = True, False
        res = render_rays(models, emb, rays, N_samples=128, perturb=0, noise_std=0, chunk=32768, img_size=512, opts=opts)
        loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()
        loss.backward()
        torch.cuda.synchronize()
    finally:
        config.set_precision(old)
    grads = {}
    for net in ("coarse", "nerf_skin", "flowbw", "flowfw"):
        if net in models:
            for k, p in models[net].named_parameters():
                grads["%s.%s" % (net, k)] = p.grad if p.grad is not None else torch.zeros_like(p)
    for k in ("bones_rst", "skin_aux"):
        if k in models:
            grads[k] = models[k].grad
    if "rest_pose_code" in models:
        grads["rest_pose_code"] = models["rest_pose_code"].weight.grad
    for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
        if k in rays and rays[k].grad is not None:
            grads["rays." + k] = rays[k].grad
    return res, grads, g


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["lbs", "trans", "se3"])
@pytest.mark.parametrize("precision", ["fp32", "fp16"])
def test_motion_models_on_gpu_against_reference(kind, precision):
    """fp32 mode: outputs 2e-5 abs, gradients 5e-3 relative to each tensor's scale (two fp32 evaluations of these
    gradients differ by that much: the fp32 reference itself sits 4e-3 from its fp64 run).  fp16
    mode (trunk and nerf_skin on the tensor-core chains; the flow nets stay on the fp32 kernels): outputs 1e-3 abs
    (north_star), gradients 2e-2."""
    res, grads, g = _gpu_run(kind, precision)
    out_bar, grad_bar = (2e-5, 5e-3) if precision == "fp32" else (1e-3, 2e-2)
    table = ["%-44s %10s %10s" % ("tensor", "err", "|ref|max")]
    bad = []
    for k in OUT_KEYS:
        e = max_abs(res[k], g["out." + k])
        table.append("%-44s %10.2e %10.2e  (abs)" % ("out." + k, e, float(np.abs(g["out." + k]).max())))
        if not e < out_bar:
            bad.append(table[-1])
    n = 0
    for k, v in sorted(grads.items()):
        if "grad." + k not in g:
            continue
        ref = g["grad." + k]
        if np.abs(ref).max() == 0:
            e = max_abs(v, ref)
        else:
            e = rel_err(v, ref)
        n += 1
        table.append("%-44s %10.2e %10.2e" % (k, e, float(np.abs(ref).max())))
        if not e < grad_bar:
            bad.append(table[-1])
    dump_table("r02_motion_%s_%s" % (kind, precision), table)
    assert n >= 20, "too few gradients compared (%d)" % n
    assert not bad, "; ".join(bad)


def test_lbs_and_flow_host_algebra_matches_the_oracle_on_cpu():
    """The O(rays x bones) algebra of the alternative motion models is plain tensor code in the product (no kernel): on
    the CPU it must agree with the oracle restatement, values and gradients -- bone_transform (rigid), matrix_to_quaternion
    (incl. rotations whose largest quaternion component is not the real part), rts_invert, blend_skinning / lbs, the SO(3)
    exponential map and the SE3head output stage."""
    from moda_b200 import geom_utils as G
    from moda_b200.nerf import so3_exp_map
    gen = torch.Generator().manual_seed(2)
    B, N, S = 7, 5, 9
    # rotations by large angles about random axes: every pivot of matrix_to_quaternion gets used
    w = torch.randn(N * B, 3, generator=gen) * 2.0
    R = O.so3_exp(w)
    assert max_abs(so3_exp_map(w), R) < 1e-6
    q = G.matrix_to_quaternion(R)
    assert max_abs(q, O.matrix_to_quat(R)) < 1e-6
    assert len(set(O.matrix_to_quat(R).abs().argmax(-1).tolist())) >= 3, "fixture does not exercise several pivots"
    T = 0.1 * torch.randn(N, B, 3, generator=gen)
    bones = torch.cat([0.1 * torch.randn(B, 3, generator=gen), torch.nn.functional.normalize(torch.randn(B, 4, generator=gen), dim=-1),
                       0.1 * torch.randn(B, 3, generator=gen)], -1)
    skin = torch.softmax(torch.randn(N, S, B, generator=gen), -1)
    xyz = torch.randn(N, S, 3, generator=gen)
    up = torch.randn(N, S, 3, generator=gen)
    ub = torch.randn(N, B, 10, generator=gen)
    for backward in (True, False):
        grads = []
        for mod in ("product", "oracle"):
            rts = torch.cat([R.reshape(N, B, 9), T], -1).reshape(N, B * 12).clone().requires_grad_(True)
            b = bones.clone().requires_grad_(True)
            sk = skin.clone().requires_grad_(True)
            if mod == "product":
                x, bd = G.lbs(b, rts, sk, xyz, backward=backward)
            else:
                x = O.lbs(b, rts, sk, xyz, backward=backward)
                bd = O.bone_transform_rigid(b, rts)
            ((x * up).sum() + (bd * ub).sum()).backward()
            grads.append((x.detach(), bd.detach(), rts.grad, b.grad, sk.grad))
        for a, c in zip(*grads):
            assert rel_err(a, c) < 1e-5
    rts3 = torch.cat([R.reshape(N, B, 3, 3), T[..., None]], -1)
    assert max_abs(G.rts_invert(rts3), O.rts_invert(rts3)) < 1e-6
    # SE3head output stage = the oracle's flow_field tail
    from moda_b200.nerf import SE3head
    raw = torch.randn(N, S, 9, generator=gen)
    flow = SE3head.post(None, raw, xyz)
    r = raw.reshape(-1, 9)
    p = xyz.reshape(-1, 3) + 0.1 * r[:, 3:6]
    ref = (O.so3_exp(r[:, :3]).matmul(p[..., None])[..., 0] - 0.1 * r[:, 3:6] + 0.1 * r[:, 6:9]).reshape(xyz.shape) - xyz
    assert max_abs(flow, ref) < 1e-6
