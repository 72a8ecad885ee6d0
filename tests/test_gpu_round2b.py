"""GPU tests of the second round-2 session: the multi-job weight-gradient kernel against an fp64 product, the folded
final layer against the unfolded chain programs, and the deferred weight-gradient join against the in-Function join."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _nrel(x, y):
    return float((x.double() - y.double()).norm() / (y.double().norm() + 1e-30))


@pytest.mark.parametrize("N,K", [(256, 256), (256, 64), (64, 64), (128, 128)])
def test_wgrad_multi_against_fp64_product(N, K):
    """dW[j] (first n_valid rows / k_valid columns, at a column offset of a wider matrix) += oscale dY[j]^T X[j] and
    dbias[j] += oscale colsum(dY[j]) for several jobs of one shape in one launch; M not a multiple of the 64-row chunk,
    fewer chunks than CTAs for the last job count."""
    from moda_b200 import chain_tc
    gen = torch.Generator().manual_seed(N * 1000 + K)
    for M, njobs in ((4096 + 37, 3), (200, 9), (64 * 148 * 2 + 5, 2)):
        osc = torch.tensor([0.25], device=DEV)
        jobs, refs = [], []
        for j in range(njobs):
            dY = (torch.randn(M, N, generator=gen) * 0.5).to(DEV).half()
            X = torch.randn(M, K, generator=gen).to(DEV).half()
            n_valid = N if j % 2 == 0 else N - 7
            k_valid = K if j % 3 == 0 else K - 1
            col0 = 0 if j % 2 == 0 else 5
            dW = torch.randn(N, K + 9, generator=gen).to(DEV)
            db = torch.randn(N, generator=gen).to(DEV) if j % 2 == 0 else None
            ref_w = dW.double().clone()
            ref_w[:n_valid, col0:col0 + k_valid] += 0.25 * (dY.double().t() @ X.double())[:n_valid, :k_valid]
            ref_b = None if db is None else db.double() + 0.25 * torch.cat([dY.double().sum(0)[:n_valid],
                                                                              torch.zeros(N - n_valid, device=DEV, dtype=torch.float64)])
            jobs.append((dY, X, dW, col0, n_valid, k_valid, db))
            refs.append((ref_w, ref_b))
        chain_tc._wgrad_multi(jobs, N, K, M, osc)
        torch.cuda.synchronize()
        for j, ((dY, X, dW, col0, n_valid, k_valid, db), (ref_w, ref_b)) in enumerate(zip(jobs, refs)):
            scale = float((ref_w - dW.double()).abs().max() / ref_w.abs().max())
            assert scale < 2e-6, (M, njobs, j, "dW", scale)
            if db is not None:
                e = float((ref_b - db.double()).abs().max() / ref_b.abs().max())
                assert e < 2e-6, (M, njobs, j, "dbias", e)


def _mlp_pass(model, emb, pts, code, dirs, w):
    from moda_b200 import geom_utils as G
    model.zero_grad()
    p = pts.clone().requires_grad_(True)
    c = code.clone().requires_grad_(True)
    if dirs is not None:
        d = dirs.clone().requires_grad_(True)
        out = G.evaluate_mlp(model, p, embed_xyz=emb, dir_embedded=d, code=c)
    else:
        out = G.evaluate_mlp(model, p, embed_xyz=emb, code=c)
    (out * w).sum().backward()
    return out.detach(), p.grad, c.grad, {k: v.grad.clone() for k, v in model.named_parameters() if v.grad is not None}


def test_folded_final_layer_matches_unfolded_chain_programs():
    """xyz_encoding_final folded into dir_encoding (config.fold_final) is the same mathematics: both
    chain programs agree to fp16-operand rounding on outputs and on every gradient, including those of the two layers that
    were folded (their gradients are mapped back from dW')."""
    from moda_b200 import config, synth, models as MM
    config.set_precision("fp16")
    prob = synth.make_problem(8, seed=0)
    models, emb, _ = MM.build_models(prob, DEV)
    gen = torch.Generator().manual_seed(11)
    R, S = 300, 128   # 300 tiles: two waves of the CTA-pair kernels with a ragged tail
    pts = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(DEV)
    was = config.fold_final
    try:
        for name, code, dirs, wshape in (("coarse", (0.1 * torch.randn(R, 64, generator=gen)).to(DEV),
                                          torch.randn(R, 27, generator=gen).to(DEV), 4),
                                         ("nerf_skin", (0.1 * torch.randn(R, 128, generator=gen)).to(DEV), None, 25)):
            w = (torch.randn(R, S, wshape, generator=gen) * 1e-3).to(DEV)
            res = {}
            for fold in (False, True):
                config.fold_final = fold
                res[fold] = _mlp_pass(models[name], emb["xyz"], pts, code, dirs, w)
            a, b = res[False], res[True]
            out_bar = 2e-3 if name == "coarse" else 2e-5
            assert _nrel(b[0], a[0]) < out_bar, (name, "out", _nrel(b[0], a[0]))
            assert _nrel(b[1], a[1]) < 5e-3, (name, "gpts", _nrel(b[1], a[1]))
            # per-ray code gradients: a direction-layer ReLU unit within fp16 rounding of its kink flips between the two
            # programs for a few rays (same effect as against the fp64 oracle, profiles/r02_diag_env_code.txt)
            assert _nrel(b[2], a[2]) < 1e-2, (name, "gcode", _nrel(b[2], a[2]))
            assert set(a[3]) == set(b[3])
            for k in a[3]:
                bar = 1e-2 if k.endswith("bias") else 5e-3
                assert _nrel(b[3][k], a[3][k]) < bar, (name, k, _nrel(b[3][k], a[3][k]))
            for k in ("xyz_encoding_final.weight", "xyz_encoding_final.bias", "dir_encoding.0.weight"):
                assert float(b[3][k].abs().max()) > 0, (name, k, "no gradient mapped back")
    finally:
        config.fold_final = was


def test_deferred_weight_gradient_join_gives_the_same_step():
    """With FlatParams the weight-gradient kernels run on side streams joined at the end of backward()
    (chain_tc._SideQueue).  Same kernels, same inputs: the flat gradient equals the one of the in-Function join up to the
    order of the fp32 atomics, and a second step right behind the first (operands kept alive until the join, then freed)
    reproduces it."""
    from moda_b200 import config, synth, models as MM
    from moda_b200.parallel import FlatParams
    from moda_b200.rendering import render_rays
    config.set_precision("fp16")
    prob = synth.make_problem(512, seed=4)
    was = config.defer_wgrad
    grads = {}
    try:
        for defer in (False, True, True):
            config.defer_wgrad = defer
            models, emb, rays = MM.build_models(prob, DEV)
            flat = FlatParams(MM.parameters_of(models))
            for rep in range(2):
                flat.zero_grad()
                res = render_rays(models, emb, rays, N_samples=128, perturb=0, noise_std=0, opts=synth.default_opts(), img_size=512)
                loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()
                loss.backward()
                g = flat.grad.clone()
                grads.setdefault(defer, []).append(g)
                # churn the allocator: anything freed too early would be overwritten here
                junk = [torch.full((1 << 20,), 7.0, device=DEV) for _ in range(8)]
                del junk
        torch.cuda.synchronize()
    finally:
        config.defer_wgrad = was
    ref = grads[False][0]
    assert float(ref.abs().max()) > 0
    for g in grads[False][1:] + grads[True]:
        assert float((g - ref).abs().max() / ref.abs().max()) < 1e-5


def test_feat_chain_matches_layer_by_layer_kernels():
    """nerf_feat (5 x 128) as one chain kernel per pass against the layer-by-layer tensor-core kernels (generic_tc) and the
    fp32 SIMT kernels: outputs and every gradient, at a size with several waves and a ragged last tile, in inference mode
    too (nothing saved)."""
    from moda_b200 import config, synth, models as MM, geom_utils as G
    prob = synth.make_full_problem(8, seed=1)
    models, emb, _ = MM.build_full_models(prob, DEV)
    feat = models["nerf_feat"]
    gen = torch.Generator().manual_seed(21)
    was = (config.precision, config.feat_chain)
    try:
        for R, S in ((331, 128), (16, 128), (63, 127), (1, 50)):
            pts = (torch.rand(R, S, 3, generator=gen) * 0.4 - 0.2).to(DEV)
            w = (torch.randn(R, S, 16, generator=gen) * 1e-3).to(DEV)
            res = {}
            for tag, prec, chain in (("simt", "fp32", False), ("layered", "fp16", False), ("chain", "fp16", True)):
                config.set_precision(prec)
                config.feat_chain = chain
                feat.zero_grad()
                p = pts.clone().requires_grad_(True)
                out = G.evaluate_mlp(feat, p, embed_xyz=emb["xyz"])
                assert out.shape == (R, S, 16)
                (out * w).sum().backward()
                res[tag] = (out.detach(), p.grad, {k: v.grad.clone() for k, v in feat.named_parameters() if v.grad is not None})
                if tag == "chain":
                    with torch.no_grad():
                        out2 = G.evaluate_mlp(feat, pts, embed_xyz=emb["xyz"])
                    assert torch.equal(out2, out.detach())
            # outputs: both fp16 paths against the fp32 kernels; gradients: this random upstream gradient is a stress case
            # (signed sums over all samples cancel by ~sqrt(N) and the PE adjoint multiplies the fp16 noise of d_pe by
            # up to 2^9), so the bar for the chain is what the layer-by-layer fp16 kernels reach on the same input
            sim, lay, cha = res["simt"], res["layered"], res["chain"]
            assert _nrel(cha[0], sim[0]) < 3e-3, (R, "out", _nrel(cha[0], sim[0]))
            assert _nrel(cha[0], lay[0]) < 3e-3, (R, "out vs layered", _nrel(cha[0], lay[0]))
            assert set(sim[2]) == set(cha[2]) == set(lay[2])
            pairs = [("gpts", lay[1], cha[1], sim[1])] + [(k, lay[2][k], cha[2][k], sim[2][k]) for k in sim[2]]
            for k, l_, c_, s_ in pairs:
                el, ec = _nrel(l_, s_), _nrel(c_, s_)
                print("%4d x %3d  %-28s layered %.3e  chain %.3e" % (R, S, k, el, ec))
                assert ec < max(1.3 * el, 5e-3), (R, k, "chain %.3e layered %.3e" % (ec, el))
    finally:
        config.set_precision(was[0])
        config.feat_chain = was[1]


def test_flat_adamw_step_matches_torch_adamw():
    """FlatParams.adamw_step (moda_adamw_flat) against torch.optim.AdamW on the same flat buffer: five steps with fresh
    gradients, defaults and a non-default configuration."""
    from moda_b200.parallel import FlatParams
    gen = torch.Generator().manual_seed(3)
    for kw in (dict(lr=1e-4), dict(lr=3e-3, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.1)):
        ts = [torch.nn.Parameter(torch.randn(s, generator=gen).to(DEV)) for s in ((256, 63), (256,), (3, 128), (1,), (25, 10))]
        flat = FlatParams(ts)
        ref = flat.flat.detach().clone().requires_grad_(True)
        opt = torch.optim.AdamW([ref], **kw)
        for step in range(5):
            g = torch.randn(flat.numel, generator=gen).to(DEV) * (0.1 + step)
            flat.grad.copy_(g)
            ref.grad = g.clone()
            flat.adamw_step(**kw)
            opt.step()
            err = float((flat.flat.detach() - ref.detach()).abs().max() / ref.detach().abs().max())
            assert err < 2e-6, (kw, step, err)
        assert torch.equal(ts[0].data.reshape(-1), flat.flat.detach()[:256 * 63]), "parameters are views of the flat buffer"


def test_edge_cases_of_the_new_entry_points():
    """Empty inputs (P = 0) and a single sample through the chain Functions added this round (nerf_vis on the nerf_skin
    chain, nerf_feat chain), an empty job list of the multi-job weight gradient, and an LBS warp with ONE bone (weights
    exactly one: the blend is that bone's rigid transform and backward o forward is the identity)."""
    from moda_b200 import chain_tc, config, synth, models as MM, geom_utils as G
    config.set_precision("fp16")
    prob = synth.make_full_problem(4, seed=2)
    models, emb, _ = MM.build_full_models(prob, DEV)
    for name, oc in (("nerf_vis", 1), ("nerf_feat", 16)):
        net = models[name]
        net.zero_grad()
        out0 = G.evaluate_mlp(net, torch.zeros(0, 7, 3, device=DEV), embed_xyz=emb["xyz"])
        assert out0.shape == (0, 7, oc)
        p = (torch.rand(1, 1, 3, device=DEV) * 0.2).requires_grad_(True)
        out1 = G.evaluate_mlp(net, p, embed_xyz=emb["xyz"])
        assert out1.shape == (1, 1, oc) and torch.isfinite(out1).all()
        out1.sum().backward()
        assert p.grad is not None and torch.isfinite(p.grad).all()
        assert all(torch.isfinite(q.grad).all() for q in net.parameters() if q.grad is not None)
    chain_tc._wgrad_multi([], 64, 64, 0, torch.ones(1, device=DEV))   # nothing to do: no launch, no error
    # LBS with one bone
    gen = torch.Generator().manual_seed(5)
    R = synth._small_rotation(gen, 3, 0.3).reshape(3, 1, 9)
    T = 0.05 * torch.randn(3, 1, 3, generator=gen)
    rts = torch.cat([R, T], -1).reshape(3, 12).to(DEV)
    bones = prob["bones_rst"][:1].to(DEV)
    xyz = (torch.rand(3, 5, 3, generator=gen) - 0.5).to(DEV)
    skin = torch.ones(3, 5, 1, device=DEV)
    y, bd = G.lbs(bones, rts, skin, xyz, backward=True)
    x2, _ = G.lbs(bones, rts, skin, y, backward=False)
    assert float((x2 - xyz).abs().max()) < 1e-5
    assert bd.shape == (3, 1, 10)


def test_sinkhorn_pass_against_matrix_vector_products():
    """moda_sinkhorn_pass (one pass over K for y = K x, z = g(y), w = K^T z) against torch in fp64: both modes, a row
    count that is not a multiple of the rows per trip, the 8000-column lattice size and a small one, row sums only."""
    from moda_b200._lib import call, ptr, stream
    gen = torch.Generator().manual_seed(9)
    for n, m in ((8192 + 3, 8000), (37, 64), (1, 8192)):
        K = torch.rand(n, m, generator=gen).to(DEV) * 0.01
        x = torch.rand(m, generator=gen).to(DEV)
        u = torch.rand(n, generator=gen).to(DEV) + 0.5
        v = torch.rand(n, generator=gen).to(DEV) + 0.5
        Kd, xd = K.double(), x.double()
        for mode in (0, 1):
            y, z = torch.empty(n, device=DEV), torch.empty(n, device=DEV)
            w = torch.zeros(m, device=DEV)
            call("moda_sinkhorn_pass", ptr(K), n, m, ptr(x), ptr(y), ptr(z), ptr(w), mode, 1.0 / n, 1e-8, ptr(u), ptr(v),
                 0, None, None, None, 0.0, None, stream())
            yd = Kd @ xd
            zd = (1.0 / n) / (yd + 1e-8) if mode == 0 else -yd * u.double() / (v.double() + 1e-8)
            wd = Kd.t() @ zd
            for name, got, ref in (("y", y, yd), ("z", z, zd), ("w", w, wd)):
                e = float((got.double() - ref).abs().max() / ref.abs().max())
                assert e < 5e-6, (n, m, mode, name, e)
        y2 = torch.empty(n, device=DEV)
        call("moda_sinkhorn_pass", ptr(K), n, m, ptr(x), ptr(y2), None, None, 0, 1.0, 0.0, None, None,
             0, None, None, None, 0.0, None, stream())
        assert float((y2.double() - Kd @ xd).abs().max() / (Kd @ xd).abs().max()) < 5e-6
        # input vector formed while staged: xmode 1 (b = p2 / (c + delta)) and xmode 2 (gc = -gb b / (c + delta))
        c = torch.rand(m, generator=gen).to(DEV) + 0.1
        gbv = torch.randn(m, generator=gen).to(DEV)
        for xmode, xref in ((1, (1.0 / m) / (c.double() + 1e-8)), (2, -gbv.double() * x.double() / (c.double() + 1e-8))):
            y3, xo = torch.empty(n, device=DEV), torch.empty(m, device=DEV)
            call("moda_sinkhorn_pass", ptr(K), n, m, None, ptr(y3), None, None, 0, 1.0, 0.0, None, None,
                 xmode, ptr(c), ptr(gbv), ptr(x), 1.0 / m, ptr(xo), stream())
            assert float((xo.double() - xref).abs().max() / xref.abs().max()) < 1e-6, (n, m, xmode)
            assert float((y3.double() - Kd @ xref).abs().max() / (Kd @ xref).abs().max()) < 5e-6, (n, m, xmode)


def test_sinkhorn_matrix_rows_cols_gcost_against_fp64():
    """The other one-pass kernels of the feature matching (csrc/sinkhorn.cu) against torch in fp64: the matrix exp((F V^T -
    1) / eps) with its first column sums, the four-vector row and column products, and gF / gV from the factored cost
    gradient; ragged sizes (rows not a multiple of the row blocks, columns not a multiple of 256)."""
    from moda_b200._lib import call, ptr, stream
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(10)
    for n, m in ((1000 + 3, 8000), (37, 64), (1, 260)):
        f = F.normalize(torch.randn(n, 16, generator=gen), 2, -1).to(DEV)
        v = F.normalize(torch.randn(m, 16, generator=gen), 2, -1).to(DEV)
        K = torch.empty(n, m, device=DEV)
        c = torch.zeros(m, device=DEV)
        call("moda_sinkhorn_matrix", ptr(f), ptr(v), n, m, 16, 0.03, 1.0 / n, ptr(K), ptr(c), stream())
        Kd = torch.exp((f.double() @ v.double().t() - 1.0) / 0.03)
        # the exponent is (f.v - 1) / 0.03: one fp32 rounding of f.v (6e-8) is a relative 2e-6 in K
        assert float(((K.double() - Kd).abs() / Kd).max()) < 2e-5, (n, m)
        assert float((c.double() - Kd.sum(0) / n).abs().max() / (Kd.sum(0) / n).abs().max()) < 1e-5
        X = torch.randn(m, 4, generator=gen).to(DEV)
        out = torch.empty(n, 4, device=DEV)
        call("moda_sinkhorn_rows4", ptr(K), n, m, ptr(X), ptr(out), stream())
        ref = K.double() @ X.double()
        assert float((out.double() - ref).abs().max() / ref.abs().max()) < 5e-6, (n, m)
        Wt = torch.randn(n, 4, generator=gen).to(DEV)
        oc = torch.zeros(m, 4, device=DEV)
        call("moda_sinkhorn_cols4", ptr(K), n, m, ptr(Wt), ptr(oc), stream())
        ref = K.double().t() @ Wt.double()
        assert float((oc.double() - ref).abs().max() / ref.abs().max()) < 5e-6, (n, m)
        L = torch.randn(n, 44, generator=gen).to(DEV)
        Rm = torch.randn(44, m, generator=gen).to(DEV)
        gF, gV = torch.zeros(n, 16, device=DEV), torch.zeros(m, 16, device=DEV)
        call("moda_sinkhorn_gcost", ptr(K), n, m, ptr(L), ptr(Rm), 44, ptr(f), ptr(v), 16, 0.03, ptr(gF), ptr(gV), stream())
        gc = K.double() * (L.double() @ Rm.double()) / 0.03
        for name, got, ref in (("gF", gF, gc @ v.double()), ("gV", gV, gc.t() @ f.double())):
            assert float((got.double() - ref).abs().max() / ref.abs().max()) < 1e-5, (n, m, name)


def test_sinkhorn_match_fused_against_torch_ops_and_fp64():
    """SinkhornMatchFn with its five primitives on the one-pass kernels against the same algebra with tensor-op primitives
    (the form the CPU oracle test pins against autograd through the reference's loop) in fp64 and fp32: matched points and
    the gradients of both feature sets."""
    from moda_b200 import loss_utils as LU
    import torch.nn.functional as F
    gen = torch.Generator().manual_seed(11)
    n, m = 300, 512
    f0 = torch.randn(n, 16, generator=gen)
    v0 = torch.randn(m, 16, generator=gen)
    q = (torch.rand(m, 3, generator=gen) - 0.5).to(DEV)
    gout = torch.randn(n, 3, generator=gen).to(DEV)

    def run(dtype, fused):
        f = f0.to(DEV, dtype).requires_grad_(True)
        v = v0.to(DEV, dtype).requires_grad_(True)
        LU._USE_KERNELS = fused
        try:
            pts = LU.SinkhornMatchFn.apply(F.normalize(f, 2, -1), F.normalize(v, 2, -1), q.to(dtype))
            (pts * gout.to(dtype)).sum().backward()
        finally:
            LU._USE_KERNELS = True
        return pts.detach().double(), f.grad.double(), v.grad.double()

    ref = run(torch.float64, False)
    plain = run(torch.float32, False)
    fused = run(torch.float32, True)
    for name, r, p, g in zip(("pts", "d feats", "d lattice feats"), ref, plain, fused):
        scale = float(r.abs().max())
        e_plain, e_fused = float((p - r).abs().max()) / scale, float((g - r).abs().max()) / scale
        assert e_fused < max(2e-5, 3 * e_plain), (name, e_fused, e_plain)


def test_fused_flow_rendering_against_tensor_ops_fp64():
    """project_render_flo on the fused kernels (csrc/flow.cu) against the tensor-op composition obj_to_cam -> pinhole_cam ->
    vrender_flo in fp64: flow, validity flags and the gradients of points, weights and the camera vector; samples behind
    the camera and far outside the image are in the batch (they must drop out of the sums and get no gradient); ragged
    sample count and a ray count that is not a multiple of the warps per CTA."""
    from moda_b200 import geom_utils as G
    gen = torch.Generator().manual_seed(12)
    N, S, img = 37, 45, 64.0
    xyz0 = torch.randn(N, S, 3, generator=gen) * 0.3
    xyz0[:, :, 2] += 3.0
    xyz0[3, 5, 2] = -4.0            # behind the camera
    xyz0[7, :, 0] += 40.0           # a whole ray far outside the image
    xyz0[9, 11, 1] = 25.0
    w0 = torch.rand(N, S, generator=gen)
    A = torch.linalg.qr(torch.randn(N, 3, 3, generator=gen))[0]
    T = torch.randn(N, 3, generator=gen) * 0.1
    Kinv = torch.zeros(N, 3, 3)
    fx = 60.0 + torch.rand(N, generator=gen) * 5
    fy = 62.0 + torch.rand(N, generator=gen) * 5
    Kinv[:, 0, 0], Kinv[:, 1, 1], Kinv[:, 2, 2] = 1 / fx, 1 / fy, 1.0
    Kinv[:, 0, 2], Kinv[:, 1, 2] = -32.0 / fx, -30.0 / fy
    rtk0 = torch.cat([A.reshape(N, 9), T, Kinv.reshape(N, 9)], -1)
    xys = (torch.rand(N, 2, generator=gen) * img)
    gout = torch.randn(N, 2, generator=gen)

    def run(dtype, dev):
        xyz = xyz0.to(dev, dtype).requires_grad_(True)
        w = w0.to(dev, dtype).requires_grad_(True)
        rtk = rtk0.to(dev, dtype).requires_grad_(True)
        flo, valid = G.project_render_flo(w, xyz, rtk, xys.to(dev, dtype), img, N)
        (flo * gout.to(dev, dtype)).sum().backward()
        return [t.detach().double().cpu() for t in (flo, valid, xyz.grad, w.grad, rtk.grad)]

    ref = run(torch.float64, "cpu")      # tensor-op composition
    got = run(torch.float32, DEV)        # fused kernels
    assert torch.equal(ref[1].reshape(-1), got[1].reshape(-1))
    assert float(ref[1].sum()) < N       # the batch does contain rays with dropped samples
    for name, r, g in zip(("flo", "valid", "d xyz", "d weights", "d rtk_vec"), ref, got):
        scale = float(r.abs().max()) + 1e-30
        assert float((g.reshape(r.shape) - r).abs().max()) / scale < 2e-5, name
    assert float(got[2][3, 5].abs().max()) == 0.0 and float(got[3][7].abs().max()) == 0.0


def test_dense_target_flow_branch_equals_target_branch_on_the_same_frame():
    """rendering.py:448-459, 491-499: the `dentrg` pair (bone_rts_dentrg / rtk_vec_dentrg -> fdp_coarse, fdp_valid) runs the
    same third warp + projection + flow rendering as the target pair; fed the target frame's pose and camera it must
    reproduce flo_coarse / flo_valid, and the result carries both key pairs."""
    from moda_b200 import config, synth, models as MM
    from moda_b200.rendering import render_rays
    config.set_precision("fp16")
    prob = synth.make_full_problem(16, seed=4)
    models, emb, rays = MM.build_full_models(prob, DEV)
    rays = dict(rays)
    rays["bone_rts_dentrg"] = rays["bone_rts_target"].detach().clone()
    rays["rtk_vec_dentrg"] = rays["rtk_vec_target"].detach().clone()
    for m in ("coarse", "nerf_skin", "nerf_vis", "nerf_feat"):
        models[m].train()
    res = render_rays(models, emb, rays, N_samples=128, perturb=0, noise_std=0, chunk=32768,
                      obj_bound=prob["obj_bound"].numpy(), img_size=prob["img_size"], opts=synth.full_opts())
    assert res["fdp_coarse"].shape == res["flo_coarse"].shape and res["fdp_valid"].shape == res["flo_valid"].shape
    assert torch.equal(res["fdp_valid"], res["flo_valid"])
    d = (res["fdp_coarse"] - res["flo_coarse"]).detach().abs().max()
    assert float(d) <= 1e-6 * float(res["flo_coarse"].detach().abs().max() + 1.0)
    res["fdp_coarse"].sum().backward()      # the branch is differentiable down to the bones
    assert models["bones_rst"].grad is not None and torch.isfinite(models["bones_rst"].grad).all()
