"""GPU parity tests (run on the B200 box): the CUDA path, called through the reference-named public API,
against (a) the golden vectors produced by executing the real reference and (b) the CPU oracle on fresh
seeded inputs.  Tolerances follow BASELINE.json's north_star: skinned points 1e-5 relative (fp32),
rendered rgb/sil/depth 1e-3 absolute, gradients 1e-3 relative to each tensor's scale."""
import numpy as np
import pytest
import torch

from tests.util import golden_problem, load_npz, max_abs, rel_err

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(autouse=True)
def _exact_mode():
    """Parity is pinned in the exact fp32 mode; the tensor-core (fp16-operand) mode has its own tests below."""
    from moda_b200 import config
    old = config.precision
    config.set_precision("fp32")
    yield
    config.set_precision(old)


def cu(a):
    return torch.from_numpy(np.asarray(a)).float().to(DEV)


def test_library_loads_on_device():
    from moda_b200 import _lib
    L = _lib.lib()
    assert L.moda_device_check() == 0, L.moda_last_error()
    assert b"sm_100a" in L.moda_version()


def test_no_cpu_fallback():
    from moda_b200.nerf import Embedding
    with pytest.raises(RuntimeError):
        Embedding(3, 10)(torch.zeros(4, 3))  # CPU tensor must be refused, not silently computed


def test_geometry_golden():
    from moda_b200 import geom_utils as G, dual_quat as DQ
    from moda_b200.nerf import Embedding
    g = load_npz("geometry_fp32.npz")
    x = cu(g["embed.x"])
    for a in ("10", "6.4", "2.0"):
        assert max_abs(Embedding(3, 10, alpha=float(a))(x), g["embed.xyz.alpha" + a]) < 2e-6
    assert max_abs(Embedding(3, 4, alpha=10)(x), g["embed.dir"]) < 2e-6
    bones, rts, aux, xyz, dsk = (cu(g[k]) for k in ("geom.bones", "geom.rts", "geom.skin_aux", "geom.xyz", "geom.dskin"))
    bd = G.bone_transform(bones, rts, True, is_vec=True)
    assert max_abs(bd, g["geom.bones_dfm"]) < 2e-6
    sbw = G.skinning(bd, xyz, dsk, skin_aux=aux)
    assert max_abs(sbw, g["geom.skin_bw"]) < 5e-5
    assert max_abs(G.skinning(bones, xyz, None, skin_aux=aux), g["geom.skin_rest"]) < 5e-5
    # warps are compared with the reference's own weights as input, isolating the blend
    xc, bd2, _ = G.neu_dbs(bones, rts, cu(g["geom.skin_bw"]), xyz, backward=True)
    assert rel_err(xc, g["geom.xyz_can"]) < 1e-5
    assert max_abs(bd2, g["geom.bones_dfm"]) < 2e-6
    xf, _, _ = G.neu_dbs(bones, rts, cu(g["geom.skin_rest"]), xyz, backward=False)
    assert rel_err(xf, g["geom.xyz_fw"]) < 1e-5
    assert rel_err(G.dqs_blend_skinning(rts.view(-1, 25, 8), cu(g["geom.skin_bw"]), xyz), g["geom.blend"]) < 1e-5
    # fused weights+warp against the reference's two-step result
    assert rel_err(G.warp_points(xyz, bones, rts, aux, dsk, backward=True), g["geom.xyz_can"]) < 2e-5
    a, b = cu(g["dq.a"]), cu(g["dq.b"])
    assert max_abs(DQ.dq_mul(a, b), g["dq.mul"]) < 2e-6
    assert max_abs(DQ.dq_normalize(a), g["dq.normalize"]) < 2e-6
    assert rel_err(DQ.dq_inverse(a), g["dq.inverse"]) < 1e-5
    assert max_abs(DQ.dq_quaternion_conjugate(a), g["dq.qconj"]) == 0
    assert max_abs(DQ.dq_combined_conjugate(a), g["dq.cconj"]) == 0
    assert max_abs(DQ.q_mul(a[..., :4].reshape(-1, 4), b[..., :4].reshape(-1, 4)), g["dq.qmul"]) < 2e-6
    assert max_abs(DQ.q_normalize(a[..., :4].reshape(-1, 4)), g["dq.qnormalize"]) < 2e-6


def _coarse_from_golden():
    from moda_b200.nerf import NeRF
    nets = load_npz("nets_seed0.npz")
    sd = {k[len("coarse."):]: torch.from_numpy(v) for k, v in nets.items() if k.startswith("coarse.")}
    m = NeRF(in_channels_xyz=63, in_channels_dir=27 + 64, init_beta=0.1)
    m.load_state_dict(sd)
    return m.to(DEV)


def test_mlp_composite_pdf_grid_golden():
    from moda_b200 import rendering as Rn, geom_utils as G
    from moda_b200.nerf import Embedding
    from moda_b200.extract import density_grid
    g = load_npz("geometry_fp32.npz")
    coarse = _coarse_from_golden()
    emb = Embedding(3, 10, alpha=10)
    pts, z, d, env = cu(g["comp.pts"]), cu(g["comp.z"]), cu(g["comp.rays_d"]), cu(g["comp.env_code"])
    de = cu(g["comp.dir_embedded"])
    R, S = z.shape
    raw = G.evaluate_mlp(coarse, pts, embed_xyz=emb, dir_embedded=de[:, None].repeat(1, S, 1), code=env, chunk=4096)
    assert max_abs(raw, g["comp.raw"]) < 5e-6
    # the modular NeRF.forward on a materialised input must agree with the fused assembly
    x = torch.cat([emb(pts), de[:, None].repeat(1, S, 1), env[:, None].repeat(1, S, 1)], -1)
    assert max_abs(coarse(x), g["comp.raw"]) < 5e-6
    models = {"coarse": coarse}
    torch.manual_seed(0)
    rgb, feat, depth, w, vis, sil = Rn.inference(models, emb, pts, d, de, z, R, S, 32768, 0.0, env_code=env)
    for got, key in ((rgb, "comp.rgb"), (depth, "comp.depth"), (sil, "comp.sil"), (w, "comp.weights"), (vis, "comp.vis")):
        assert max_abs(got, g[key]) < 5e-6, key
    bins, wts = cu(g["pdf.bins"]), cu(g["pdf.weights"])
    assert max_abs(Rn.sample_pdf(bins, wts, 32, det=True), g["pdf.det"]) < 5e-6
    assert max_abs(Rn.sample_pdf(bins, wts, 32, det=False, u=cu(g["pdf.u"])), g["pdf.rand"]) < 5e-6
    vol = density_grid(coarse, 12, (0.3, 0.3, 0.3), embedding_xyz=emb)
    assert max_abs(vol, g["grid.sigma"]) < 5e-6


def _render_and_grads(prob, n_samples=128, **kw):
    from moda_b200 import models as MM, synth
    from moda_b200.rendering import render_rays
    models, emb, rays = MM.build_models(prob, DEV)
    models["coarse"].train()
    models["nerf_skin"].train()
    res = render_rays(models, emb, rays, N_samples=n_samples, perturb=0, noise_std=0, chunk=32768, img_size=512,
                      opts=synth.default_opts(), **kw)
    loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()
    loss.backward()
    grads = {}
    for k, p in models["coarse"].named_parameters():
        grads["coarse." + k] = p.grad
    for k, p in models["nerf_skin"].named_parameters():
        grads["nerf_skin." + k] = p.grad if p.grad is not None else torch.zeros_like(p)
    grads["bones_rst"] = models["bones_rst"].grad
    grads["skin_aux"] = models["skin_aux"].grad
    grads["rest_pose_code"] = models["rest_pose_code"].weight.grad
    for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
        grads["rays." + k] = rays[k].grad
    return res, loss, grads


def _report(pairs, tol, what):
    bad = []
    for name, got, ref in pairs:
        e = rel_err(got, ref) if what == "rel" else max_abs(got, ref)
        if not e <= tol:
            bad.append("%s: %.3e" % (name, e))
    assert not bad, "; ".join(bad)


def _report_vs_truth(triples, base_tol, slack=3.0):
    """triples: (name, got, truth_fp64, ref_fp32).  Several gradients of this path are ill-conditioned in
    fp32 (the skinning softmax has logits of -1000 e^s d^2, geom_utils.py:265-266: the reference's OWN fp32
    gradients differ from its fp64 gradients by up to 1e-1 relative on bone_rts / rays_o / bones_rst, see
    DESIGN.md "precision").  So the bar is: relative error against the fp64 truth <= base_tol, or no worse than
    ``slack`` x the error the fp32 reference itself makes on that tensor."""
    bad, table = [], []
    for name, got, truth, ref32 in triples:
        e = rel_err(got, truth)
        e_ref = rel_err(ref32, truth) if ref32 is not None else 0.0
        table.append("%-40s ours %.2e  fp32-ref %.2e" % (name, e, e_ref))
        if not e <= max(base_tol, slack * e_ref):
            bad.append("%s: ours %.3e vs fp32 reference %.3e" % (name, e, e_ref))
    print("\n".join(table))
    assert not bad, "; ".join(bad)


@pytest.mark.parametrize("name,fine", [("render_n32_fp32.npz", False), ("render_fine_n16_fp32.npz", True)])
def test_render_rays_against_reference_golden(name, fine):
    prob, g = golden_problem(name)
    g64 = load_npz("render_n32_fp64.npz") if not fine else None
    res, loss, grads = _render_and_grads(prob, use_fine=fine)
    # skinned points: 1e-5 relative (fp32); rendered values: 1e-3 absolute
    if g64 is not None:
        _report_vs_truth([(k, res[k], g64["out." + k], g["out." + k]) for k in ("xyz_camera_vis", "xyz_canonical_vis")], 1e-5)
    else:
        _report([(k, res[k], g["out." + k]) for k in ("xyz_camera_vis", "xyz_canonical_vis")], 3e-5, "rel")
    _report([(k, res[k], g["out." + k]) for k in ("img_coarse", "depth_rnd", "sil_coarse", "frame_cyc_dis")], 1e-3, "abs")
    assert abs(float(loss.detach()) - float(g["out.loss"])) < 1e-3
    keys = [k for k in g if k.startswith("grad.")]
    assert len(keys) > 20
    if g64 is not None:
        # fp64 truth where the fixture has it (all small tensors, nerf_skin, a subset of the trunk) ...
        _report_vs_truth([(k[5:], grads[k[5:]], g64[k], g[k]) for k in keys if k in g64], 1e-3)
        # ... and the fp32 reference for the remaining (well-conditioned) trunk layers
        _report([(k[5:], grads[k[5:]], g[k]) for k in keys if k not in g64], 1e-3, "rel")
    else:
        # importance-sampled pass: only an fp32 fixture exists; ill-conditioned tensors get the slack the
        # fp32 reference needs against its own fp64 run on the coarse fixture (<= 1.5e-1 relative)
        loose = ("nerf_skin.", "skin_aux", "bones_rst", "rays.")
        _report([(k[5:], grads[k[5:]], g[k]) for k in keys if not k[5:].startswith(loose)], 1e-3, "rel")
        _report([(k[5:], grads[k[5:]], g[k]) for k in keys if k[5:].startswith(loose)], 1.5e-1, "rel")


def test_render_rays_against_oracle_fresh_seed_with_jitter():
    from moda_b200 import synth, models as MM
    from moda_b200.rendering import render_rays
    from oracle import restated as O
    N, S = 96, 128
    prob = synth.make_problem(N, seed=5)
    torch.manual_seed(9)
    jitter = torch.rand(N, S)
    runs = {}
    for dt in (torch.float64, torch.float32):
        p = O.to_dtype(prob, dt)
        leaves = O.require_grads(p)
        r = O.render_rays(p, n_samples=S, perturb=1.0, perturb_rand=jitter.to(dt))
        O.parity_loss(r).backward()
        runs[dt] = (r, leaves)
    (res_o, leaves), (res_32, leaves32) = runs[torch.float64], runs[torch.float32]
    models, emb, rays = MM.build_models(prob, DEV)
    # feed the same jitter: render_rays draws torch.rand on the device (same call as rendering.py:82)
    orig = torch.rand
    try:
        torch.rand = lambda *a, **k: jitter.to(DEV)
        res = render_rays(models, emb, rays, N_samples=S, perturb=1.0, noise_std=0, opts=synth.default_opts(), img_size=512)
    finally:
        torch.rand = orig
    loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()
    loss.backward()
    _report_vs_truth([(k, res[k], res_o[k], res_32[k]) for k in ("xyz_camera_vis", "xyz_canonical_vis")], 1e-5)
    _report([(k, res[k], res_o[k]) for k in ("img_coarse", "depth_rnd", "sil_coarse", "frame_cyc_dis")], 1e-3, "abs")
    gz = lambda d, k: d[k].grad if d[k].grad is not None else torch.zeros_like(d[k])
    tr = [("coarse." + k, p.grad, gz(leaves, "coarse." + k), gz(leaves32, "coarse." + k))
          for k, p in models["coarse"].named_parameters()]
    tr += [("nerf_skin." + k, p.grad, gz(leaves, "nerf_skin." + k), gz(leaves32, "nerf_skin." + k))
           for k, p in models["nerf_skin"].named_parameters() if p.grad is not None]
    tr += [("bones_rst", models["bones_rst"].grad, gz(leaves, "bones_rst"), gz(leaves32, "bones_rst")),
           ("skin_aux", models["skin_aux"].grad[:1], gz(leaves, "skin_aux")[:1], gz(leaves32, "skin_aux")[:1]),
           ("rest_pose_code", models["rest_pose_code"].weight.grad, gz(leaves, "rest_pose_code"), gz(leaves32, "rest_pose_code"))]
    tr += [("rays." + k, rays[k].grad, gz(leaves, "rays." + k), gz(leaves32, "rays." + k))
           for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d")]
    _report_vs_truth(tr, 1e-3)


def test_edge_cases():
    from moda_b200 import geom_utils as G, synth
    from moda_b200.nerf import Embedding
    # empty batch
    assert Embedding(3, 10)(torch.zeros(0, 3, device=DEV)).shape == (0, 63)
    sp = synth.make_skin_problem(3, 5, seed=1)  # ragged: 5 samples per ray, not a multiple of anything
    xyz, bones, aux = sp["xyz"].to(DEV), sp["bones_rst"].to(DEV), sp["skin_aux"].to(DEV)
    w = G.skinning(bones, xyz, None, skin_aux=aux)
    assert w.shape == (3, 5, 25)
    assert max_abs(w.sum(-1), torch.ones(3, 5)) < 1e-5
    # a single bone: weights are exactly one and the warp is a rigid transform; bw o fw = identity
    rts = sp["bone_rts"].view(3, 25, 8)[:, :1].contiguous().to(DEV)
    b1 = bones[:1].contiguous()
    y = G.warp_points(xyz, b1, rts, aux, None, backward=True)
    x2 = G.warp_points(y, b1, rts, aux, None, backward=False)
    assert rel_err(x2, xyz) < 1e-5


def test_full_size_properties():
    """BASELINE config sizes, checked through size-independent properties (no oracle at this size)."""
    from moda_b200 import geom_utils as G, synth, models as MM
    from moda_b200.rendering import render_rays
    N, S = 8192, 128
    prob = synth.make_problem(N, seed=2)
    models, emb, rays = MM.build_models(prob, DEV, requires_grad=False)
    with torch.no_grad():
        res = render_rays(models, emb, rays, N_samples=S, perturb=0, noise_std=0, opts=synth.default_opts(), img_size=512)
    for k in ("img_coarse", "depth_rnd", "sil_coarse", "frame_cyc_dis"):
        assert torch.isfinite(res[k]).all(), k
    assert float(res["sil_coarse"].min()) >= -1e-6 and float(res["sil_coarse"].max()) <= 1 + 1e-5
    assert float(res["img_coarse"].min()) >= -1e-6 and float(res["img_coarse"].max()) <= 1 + 1e-5
    near, far = 0.1, 0.5
    assert float(res["depth_rnd"].max()) <= far + 1e-5
    # rays are independent: any sub-batch renders to the same values (sharding invariance)
    sub = {k: v[1000:1100] for k, v in rays.items()}
    with torch.no_grad():
        res2 = render_rays(models, emb, sub, N_samples=S, perturb=0, noise_std=0, opts=synth.default_opts(), img_size=512)
    for k in ("img_coarse", "depth_rnd", "sil_coarse", "frame_cyc_dis"):
        assert max_abs(res[k][1000:1100], res2[k]) < 1e-6, k
    # skinning weights form a partition of unity; with identity bone transforms both warps are the identity
    xyz = res["xyz_camera_vis"]
    w = G.skinning(models["bones_rst"], xyz[:2048], None, skin_aux=models["skin_aux"])
    assert max_abs(w.sum(-1), torch.ones(2048, S)) < 1e-5
    ident = torch.zeros(N, 25, 8, device=DEV)
    ident[..., 0] = 1
    y = G.warp_points(xyz, models["bones_rst"], ident, models["skin_aux"], None, backward=True)
    assert rel_err(y, xyz) < 1e-6


# (the fp16-mode-vs-fp64-oracle test moved to tests/test_gpu_parity2.py with a per-tensor relative gradient bar)


def test_tensor_core_trunk_matches_simt_trunk():
    """Same inputs through both implementations of nerf_coarse (values and gradients, fp16-level agreement)."""
    from moda_b200 import config, geom_utils as G
    from moda_b200.nerf import Embedding
    coarse = _coarse_from_golden()
    emb = Embedding(3, 10, alpha=10)
    gen = torch.Generator().manual_seed(3)
    R, S = 300, 128
    pts = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(DEV)
    de = torch.randn(R, 27, generator=gen).to(DEV)
    env = (0.1 * torch.randn(R, 64, generator=gen)).to(DEV)
    gout = torch.randn(R, S, 4, generator=gen).to(DEV) * 1e-4
    outs = {}
    for mode in ("fp32", "fp16"):
        config.set_precision(mode)
        coarse.zero_grad()
        p, d, e = pts.clone().requires_grad_(True), de.clone().requires_grad_(True), env.clone().requires_grad_(True)
        raw = G.evaluate_mlp(coarse, p, embed_xyz=emb, dir_embedded=d, code=e)
        (raw * gout).sum().backward()
        outs[mode] = (raw.detach(), p.grad, d.grad, e.grad, {k: v.grad.clone() for k, v in coarse.named_parameters() if v.grad is not None})
    a, b = outs["fp32"], outs["fp16"]
    assert max_abs(b[0], a[0]) < 5e-3
    # norm-wise agreement at fp16 level (gxyz is amplified by 2^9 through the highest PE band)
    nrel = lambda x, y: float((x.double() - y.double()).norm() / (y.double().norm() + 1e-30))
    for i, nm, tol in ((1, "gxyz", 6e-2), (2, "gdir", 2e-2), (3, "genv", 2e-2)):
        assert nrel(b[i], a[i]) < tol, (nm, nrel(b[i], a[i]))
    # gout is random-sign noise, so every gradient here is a random walk over samples and a ReLU unit whose
    # pre-activation changes sign under fp16 operand rounding (a fraction f ~ 1e-3 of them) perturbs it by
    # sqrt(f) ~ 3-4 % norm-wise; with a real loss (test_tensor_core_mode_against_fp64_oracle) the contributions
    # are coherent and the same flips stay below the 1e-3 bar.
    worst = max((nrel(b[4][k], a[4][k]), k) for k in a[4])
    assert worst[0] < 6e-2, worst


def test_split_precision_skin_mlp_matches_fp32_simt():
    """nerf_skin on tensor cores with (hi, lo) fp16 operand pairs must be fp32-class in the FORWARD direction: the
    delta logits feed a softmax over O(100) Gaussian logits, so fp16/TF32 operand rounding (1e-3 relative) would
    not do.  Its adjoint runs on plain fp16 operands (fp32 accumulation): gradients agree at fp16 level."""
    from moda_b200 import config, geom_utils as G
    from moda_b200.nerf import Embedding, NeRF
    nets = load_npz("nets_seed0.npz")
    sd = {k[len("nerf_skin."):]: torch.from_numpy(v) for k, v in nets.items() if k.startswith("nerf_skin.")}
    skin = NeRF(in_channels_xyz=63 + 128, D=5, W=64, in_channels_dir=0, out_channels=25, raw_feat=True, in_channels_code=128)
    skin.load_state_dict(sd)
    skin = skin.to(DEV)
    emb = Embedding(3, 10, alpha=10)
    gen = torch.Generator().manual_seed(5)
    R, S = 200, 128
    pts = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(DEV)
    gout = torch.randn(R, S, 25, generator=gen).to(DEV) * 1e-3
    for code in ((0.1 * torch.randn(R, 128, generator=gen)).to(DEV), torch.randn(1, 128, generator=gen).to(DEV)):
        outs = {}
        for mode in ("fp32", "fp16"):
            config.set_precision(mode)
            skin.zero_grad()
            p, c = pts.clone().requires_grad_(True), code.clone().requires_grad_(True)
            out = G.evaluate_mlp(skin, p, embed_xyz=emb, code=c)
            assert out.shape == (R, S, 25)
            (out * gout).sum().backward()
            outs[mode] = (out.detach(), p.grad, c.grad, {k: v.grad.clone() for k, v in skin.named_parameters() if v.grad is not None})
        a, b = outs["fp32"], outs["fp16"]
        nrel = lambda x, y: float((x.double() - y.double()).norm() / (y.double().norm() + 1e-30))
        assert max_abs(b[0], a[0]) < 2e-5, max_abs(b[0], a[0])
        # the PE adjoint amplifies by 2^9 with cancellation: both fp32-class paths carry ~1e-3 noise there
        assert nrel(b[1], a[1]) < 5e-3, ("gpts", nrel(b[1], a[1]))
        # a code row shared by all 25 600 samples: its gradient is a sum of signed terms that cancels by ~sqrt(N), which
        # amplifies the fp16 rounding of the adjoint chain relative to the result (per-ray codes: ~1e-3)
        assert nrel(b[2], a[2]) < (1e-2 if code.shape[0] == 1 else 3e-3), ("gcode", nrel(b[2], a[2]))
        assert set(a[3]) == set(b[3])
        # bias gradients are signed sums over all samples too (same cancellation as the shared code row); against the
        # fp64 oracle the folded and the unfolded chains sit at the same distance (profiles/r02_grad_table_n8192_*)
        # (with a shared code row the code columns of xyz_encoding_1 / _5 .weight are such sums as well)
        wbar = 1e-2 if code.shape[0] == 1 else 3e-3
        worst = max((nrel(b[3][k], a[3][k]) / (1e-2 if k.endswith("bias") else wbar), k) for k in a[3])
        assert worst[0] < 1.0, worst


def test_fused_chains_match_layer_by_layer_kernels():
    """csrc/chain.cu (one persistent kernel per MLP pass) against csrc/tc_gemm.cu (one kernel per layer): same
    precision policy, so values and every gradient agree far below the fp16-vs-fp32 gap.  Sizes cover a partial
    last tile (R*S not a multiple of 128 x SMs) and the single-row pose code of the forward warp."""
    from moda_b200 import config, geom_utils as G, synth, models as MM
    config.set_precision("fp16")
    prob = synth.make_problem(8, seed=0)
    models, emb, _ = MM.build_models(prob, DEV)
    nrel = lambda x, y: float((x.double() - y.double()).norm() / (y.double().norm() + 1e-30))

    def run(model, kind, R, S, fused, single):
        gen = torch.Generator().manual_seed(7)
        pts = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(DEV).requires_grad_(True)
        model.zero_grad()
        config.fused = fused
        if kind == "trunk":
            de = torch.randn(R, 27, generator=gen).to(DEV).requires_grad_(True)
            env = (0.1 * torch.randn(R, 64, generator=gen)).to(DEV).requires_grad_(True)
            gout = torch.randn(R, S, 4, generator=gen).to(DEV) * 1e-3
            out = G.evaluate_mlp(model, pts, embed_xyz=emb["xyz"], dir_embedded=de, code=env)
            (out * gout).sum().backward()
            grads = {"pts": pts.grad, "dir": de.grad, "env": env.grad}
        else:
            code = (0.1 * torch.randn(1 if single else R, 128, generator=gen)).to(DEV).requires_grad_(True)
            gout = torch.randn(R, S, 25, generator=gen).to(DEV) * 1e-3
            out = G.evaluate_mlp(model, pts, embed_xyz=emb["xyz"], code=code)
            (out * gout).sum().backward()
            grads = {"pts": pts.grad, "code": code.grad}
        grads.update({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
        return out.detach().clone(), grads

    try:
        for kind, model in (("trunk", models["coarse"]), ("skin", models["nerf_skin"])):
            for R, S, single in ((3, 128, False), (301, 128, False), (37, 64, True)):
                if kind == "trunk" and single:
                    continue
                a = run(model, kind, R, S, False, single)
                b = run(model, kind, R, S, True, single)
                assert max_abs(b[0], a[0]) < (2e-3 if kind == "trunk" else 2e-5), (kind, R, max_abs(b[0], a[0]))
                assert set(a[1]) == set(b[1])
                worst = max((nrel(b[1][k], a[1][k]), k) for k in a[1])
                assert worst[0] < 2e-2, (kind, R, worst)
    finally:
        config.fused = True


def test_trunk_chain_launch_modes_are_bit_identical():
    """The launch modes of the 256-wide chains (csrc/chain.cu) perform the same arithmetic in the same order per tile:
    single-CTA kernels, CTA pairs (tcgen05 cta_group::2, PAIR = 1) and CTA pairs with TWO tiles in flight per CTA
    (SLOTS = 2).  Outputs and the data gradients are bit-identical, the weight gradients (atomic accumulation order)
    agree to rounding.  Sizes: an odd tile count (the pair's dead tile), a ragged last tile, more tile pairs than CTAs,
    and enough tiles that every CTA runs several tiles per slot with an odd tile count on some CTAs."""
    from moda_b200 import config, geom_utils as G, synth, models as MM
    from moda_b200.extract import density_grid
    config.set_precision("fp16")
    prob = synth.make_problem(8, seed=0)
    models, emb, _ = MM.build_models(prob, DEV)
    model = models["coarse"]
    nrel = lambda x, y: float((x.double() - y.double()).norm() / (y.double().norm() + 1e-30))

    def run(R, S):
        gen = torch.Generator().manual_seed(11)
        pts = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(DEV).requires_grad_(True)
        de = torch.randn(R, 27, generator=gen).to(DEV).requires_grad_(True)
        env = (0.1 * torch.randn(R, 64, generator=gen)).to(DEV).requires_grad_(True)
        gout = torch.randn(R, S, 4, generator=gen).to(DEV) * 1e-3
        model.zero_grad()
        out = G.evaluate_mlp(model, pts, embed_xyz=emb["xyz"], dir_embedded=de, code=env)
        (out * gout).sum().backward()
        data = {"pts": pts.grad.clone(), "dir": de.grad.clone(), "env": env.grad.clone()}
        par = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        with torch.no_grad():
            vol = density_grid(model, 24 if R < 1000 else 48, (0.3, 0.3, 0.3), emb["xyz"])
        return out.detach().clone(), data, par, vol

    was = (config.trunk_pair, config.trunk_slots)
    try:
        for R, S in ((3, 128), (37, 64), (700, 128), (2401, 128)):
            config.set_trunk_pair(False)
            a = run(R, S)
            for slots in (1, 2):
                config.set_trunk_pair(True)
                config.set_trunk_slots(slots)
                b = run(R, S)
                tag = (R, S, "slots=%d" % slots)
                assert torch.equal(a[0], b[0]), (tag, max_abs(a[0], b[0]))
                assert torch.equal(a[3], b[3]), (tag, "density grid")
                for k in ("pts",):
                    assert torch.equal(a[1][k], b[1][k]), (tag, k, max_abs(a[1][k], b[1][k]))
                for k in ("dir", "env"):
                    assert nrel(b[1][k], a[1][k]) < 1e-5, (tag, k)
                worst = max((nrel(b[2][k], a[2][k]), k) for k in a[2])
                assert worst[0] < 1e-4, (tag, worst)
    finally:
        config.set_trunk_pair(was[0])
        config.set_trunk_slots(was[1])


def test_training_step_on_flat_parameter_buffer():
    """bench.py / the data-parallel path re-home every parameter into one flat buffer (moda_b200.parallel.FlatParams):
    the kernels must accept those views (128-bit loads need the 16-byte alignment FlatParams guarantees) and give
    exactly the result they give on separately allocated parameters."""
    from moda_b200 import synth, models as MM, config
    from moda_b200.parallel import FlatParams
    from moda_b200.rendering import render_rays
    config.set_precision("fp16")
    N, S = 160, 128
    prob = synth.make_problem(N, seed=2)
    out = []
    for flat in (False, True):
        models, emb, rays = MM.build_models(prob, DEV)
        fp = FlatParams(MM.parameters_of(models)) if flat else None
        res = render_rays(models, emb, rays, N_samples=S, perturb=0, noise_std=0, opts=synth.default_opts(), img_size=512)
        loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()
        loss.backward()
        out.append((res["img_coarse"].detach().clone(), [None if p.grad is None else p.grad.clone() for p in MM.parameters_of(models)]))
        if flat:
            assert all(p.data_ptr() % 16 == 0 for p in MM.parameters_of(models))
            assert float(fp.grad.abs().sum()) > 0
    assert torch.equal(out[0][0], out[1][0])
    assert len(out[0][1]) == len(out[1][1])
    n_checked = 0
    for a, b in zip(out[0][1], out[1][1]):
        if a is None:   # no gradient reaches it (nerf_skin's discarded sigma head): its flat .grad view stays zero
            assert float(b.abs().sum()) == 0
            continue
        n_checked += 1
        assert rel_err(b, a) < 1e-5   # atomics in the per-bone reductions reorder sums run to run
    assert n_checked >= 40
