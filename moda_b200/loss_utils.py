"""Loss terms that ``inference_deform`` evaluates inside the renderer at MoDA's default flags (nnutils/loss_utils.py):
feature matching against the canonical feature volume, key-point reprojection, and the visibility-field loss.

Per-sample work (the MLPs on P points, the skinning warps) goes through the CUDA kernels of this package
(``geom_utils.evaluate_mlp`` / ``warp_points``).  The (rays x 8000) soft-argmax / Sinkhorn algebra of ``feat_match``
and the per-ray camera algebra are expressed with device tensor ops: they are O(rays), not O(samples), and are listed
in DESIGN.md as the next candidates for fused kernels.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import geom_utils as G
from ._lib import call, ptr, stream


def visibility_loss(mlp, embed, xyz_pos, w_pos, bound, chunk):
    """loss_utils.py:125-149.  w_pos: (rays, samples) transmittance returned by the compositor; negatives are drawn
    uniformly in the object bound (same draw as the reference: torch.rand(1, nsample, 3) on the host)."""
    device = next(mlp.parameters()).device
    xyz_pos = xyz_pos.detach()
    w_pos = w_pos.detach()
    nsample = w_pos.shape[0] * w_pos.shape[1]
    # the reference draws on the host and uploads 12 B per sample (loss_utils.py:138); same distribution, drawn on the
    # device instead
    bound_t = torch.as_tensor(np.asarray(bound), dtype=torch.float32, device=device)[None, None]
    xyz_neg = torch.rand(1, nsample, 3, device=device) * 2 * bound_t - bound_t
    vis_neg_pred = G.evaluate_mlp(mlp, xyz_neg, embed_xyz=embed, chunk=chunk)[..., 0]
    vis_loss_neg = -F.logsigmoid(-vis_neg_pred).sum() * 0.1 / nsample
    vis_pos_pred = G.evaluate_mlp(mlp, xyz_pos, embed_xyz=embed, chunk=chunk)[..., 0]
    vis_loss_pos = -(F.logsigmoid(vis_pos_pred) * w_pos).sum() / nsample
    return vis_loss_pos + vis_loss_neg


def compute_pts_exp(pts_prob, pts):
    """loss_utils.py:162-172: expected 3-D point of a ray under its (renormalised) compositing weights."""
    ndepth = pts_prob.shape[-1]
    p = pts_prob.reshape(-1, ndepth, 1)
    p = p / (1e-9 + p.sum(1)[:, None])
    return (pts * p).sum(1)


def feat_match_loss(nerf_feat, embedding_xyz, feats, pts, pts_prob, bound, use_corr=True, use_ot=False,
                    is_training=True):
    """loss_utils.py:174-209.  feats (...,F) pixel features, pts (...,S,3), pts_prob (...,S)."""
    base_shape = feats.shape[:-1]
    nfeat = feats.shape[-1]
    ndepth = pts_prob.shape[-1]
    feats = feats.reshape(-1, nfeat)
    pts = pts.reshape(-1, ndepth, 3)
    pts_exp = compute_pts_exp(pts_prob, pts)
    pts_pred, corr_err = feat_match(nerf_feat, embedding_xyz, feats, bound, grid_size=20, use_corr=use_corr,
                                    use_ot=use_ot, is_training=is_training)
    feat_err = (pts_pred - pts_exp).norm(2, -1)
    pts_pred = pts_pred.reshape(base_shape + (3,))
    pts_exp = pts_exp.reshape(base_shape + (3,))
    feat_err = feat_err.reshape(base_shape + (1,))
    if use_corr:
        corr_err = corr_err.reshape(base_shape + (1,))
    return pts_pred, pts_exp, feat_err, corr_err


def _fused_ok(feats, vol_feat):
    """The one-pass Sinkhorn kernels (csrc/sinkhorn.cu) take 16-channel fp32 CUDA features and a lattice of m % 4 == 0, m <=
    8192 points; anything else (the CPU oracle tests, other feature widths) runs the same algebra as torch ops."""
    m = vol_feat.shape[0]
    return (feats.is_cuda and feats.dtype == torch.float32 and vol_feat.dtype == torch.float32 and feats.shape[1] == 16 and
            vol_feat.shape[1] == 16 and m % 4 == 0 and m <= 8192 and feats.shape[0] > 0)


class SinkhornMatchFn(torch.autograd.Function):
    """pts_pred (N,3) = rownorm(a K b^T) Q with K = exp(-(1 - F V^T) / 0.03) and (a, b) from 20 Sinkhorn iterations on
    uniform marginals (loss_utils.py:347-386) -- with the ANALYTIC adjoint of the unrolled iterations instead of an
    autograd tape over 40 (N x M) matrix-vector products (each of which would keep and differentiate a 262 MB matrix at
    8192 rays x 8000 lattice points).  Notes: `a` cancels in the row normalisation, so only b = b_20 enters the output;
    with c_i = K^T a_{i-1}, b_i = p2 / (c_i + d), d_i = K b_i, a_i = p1 / (d_i + d) the reverse sweep is
        gc_i = -gb_i b_i / (c_i + d);  gK += a_{i-1} (x) gc_i;  ga_{i-1} = K gc_i;
        gd_{i-1} = -ga_{i-1} a_{i-1} / (d_{i-1} + d);  gK += gd_{i-1} (x) b_{i-1};  gb_{i-1} = K^T gd_{i-1}
    so gK is a rank-40 sum plus the direct term, formed once.  Verified against autograd through the reference's loop
    to 1e-15 in fp64 (tests/test_oracle.py)."""
    EPS, DELTA, ITERS = 0.03, 1e-8, 20

    @staticmethod
    def forward(ctx, feats, vol_feat, query):
        if _fused_ok(feats, vol_feat):
            return SinkhornMatchFn._forward_fused(ctx, feats, vol_feat, query)
        K = torch.exp((feats.matmul(vol_feat.t()) - 1.0) / SinkhornMatchFn.EPS)
        n, m = K.shape
        a = torch.full((n,), 1.0 / n, device=K.device, dtype=K.dtype)
        As, Bs, Cs, Ds = [a], [], [], []
        dl, it = SinkhornMatchFn.DELTA, SinkhornMatchFn.ITERS
        for _ in range(it):
            c = torch.mv(K.t(), a)
            b = (1.0 / m) / (c + dl)
            d = torch.mv(K, b)
            a = (1.0 / n) / (d + dl)
            As.append(a), Bs.append(b), Cs.append(c), Ds.append(d)
        # T = a K b^T row-normalised: with the final a this is K b^T / rowsum; the reference normalises a K b^T whose row
        # sums are a_20 (K b_20): identical up to rounding
        s = Ds[-1]
        pts = (K * Bs[-1][None]).matmul(query) / s[:, None]
        ctx.fused = False
        ctx.save_for_backward(feats, vol_feat, query, K, pts, s)
        ctx.hist = (As, Bs, Cs, Ds)
        return pts

    @staticmethod
    def _forward_fused(ctx, feats, vol_feat, query):
        """The same iteration with every (n x m)-sized operation as one pass over K (csrc/sinkhorn.cu): the matrix is written
        once together with the first column sums (moda_sinkhorn_matrix); each iteration is ONE pass (moda_sinkhorn_pass: b_i
        from c_i while it is staged, d_i = K b_i, a_i = p1 / (d_i + delta) and, with the rows still on chip, c_{i+1} = K^T
        a_i); the soft-argmax is one pass with four vectors (moda_sinkhorn_rows4)."""
        feats, vol_feat, query = feats.contiguous(), vol_feat.contiguous(), query.contiguous()
        n, m, dev = feats.shape[0], vol_feat.shape[0], feats.device
        dl, it = SinkhornMatchFn.DELTA, SinkhornMatchFn.ITERS
        new = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
        K = new(n, m)
        acc = torch.zeros(it, m, device=dev, dtype=torch.float32)     # the column sums c_1 .. c_20: one fill
        call("moda_sinkhorn_matrix", ptr(feats), ptr(vol_feat), n, m, feats.shape[1], SinkhornMatchFn.EPS, 1.0 / n, ptr(K),
             ptr(acc[0]), stream())
        As, Bs, Cs, Ds = [torch.full((n,), 1.0 / n, device=dev, dtype=torch.float32)], [], [], []
        hist = new(it, m)                                                  # b_1 .. b_20
        rows = new(2 * it, n)                                              # d_i, a_i
        for i in range(it):
            c, b, d, a = acc[i], hist[i], rows[2 * i], rows[2 * i + 1]
            c_next = acc[i + 1] if i + 1 < it else None
            call("moda_sinkhorn_pass", ptr(K), n, m, None, ptr(d), ptr(a), ptr(c_next), 0, 1.0 / n, dl, None, None,
                 1, ptr(c), None, None, 1.0 / m, ptr(b), stream())
            As.append(a), Bs.append(b), Cs.append(c), Ds.append(d)
        # pts = (K b) Q / rowsum: X = b (x) [Qx Qy Qz 1]; the fourth product is the row sum d_20 again
        X = torch.cat([query, torch.ones_like(query[:, :1])], 1) * Bs[-1][:, None]
        out4 = new(n, 4)
        call("moda_sinkhorn_rows4", ptr(K), n, m, ptr(X), ptr(out4), stream())
        s = Ds[-1]
        pts = out4[:, :3] / s[:, None]
        ctx.fused = True
        ctx.save_for_backward(feats, vol_feat, query, K, pts, s)
        ctx.hist = (As, Bs, Cs, Ds)
        return pts

    @staticmethod
    def backward(ctx, g):
        if ctx.fused:
            return SinkhornMatchFn._backward_fused(ctx, g)
        feats, vol_feat, query, K, pts, s = ctx.saved_tensors
        As, Bs, Cs, Ds = ctx.hist
        dl, it = SinkhornMatchFn.DELTA, SinkhornMatchFn.ITERS
        b20 = Bs[-1]
        U = (g.matmul(query.t()) - (g * pts).sum(-1, keepdim=True)) / s[:, None]   # dL / d(K_nj b_j)
        gb = (U * K).sum(0)
        gK = U * b20[None]
        left, right = [], []            # gK += sum_k left_k (x) right_k
        for i in range(it, 0, -1):
            gc = -gb * Bs[i - 1] / (Cs[i - 1] + dl)
            left.append(As[i - 1]), right.append(gc)
            if i - 1 < 1:
                break
            ga = torch.mv(K, gc)
            gd = -ga * As[i - 1] / (Ds[i - 2] + dl)
            gb = torch.mv(K.t(), gd)
            left.append(gd), right.append(Bs[i - 2])
        gK.addmm_(torch.stack(left, 1), torch.stack(right, 0))
        gcost = gK.mul_(K).mul_(1.0 / SinkhornMatchFn.EPS)
        gq = None
        if ctx.needs_input_grad[2]:
            gq = (K * b20[None] / s[:, None]).t().matmul(g)
        return gcost.matmul(vol_feat), gcost.t().matmul(feats), gq

    @staticmethod
    def _backward_fused(ctx, g):
        """The reverse sweep with one pass over K per iteration, and dLoss/dK kept in factored form: the direct term
        U (.) b_20 with U[r][j] = (g_r . Q_j - g_r . pts_r) / s_r has rank 4 and the sweep adds 39 outer products, so
        gcost = K / eps * (L Rm) with L (n, 44), Rm (44, m) is consumed tile by tile (moda_sinkhorn_gcost) and never
        written."""
        feats, vol_feat, query, K, pts, s = ctx.saved_tensors
        As, Bs, Cs, Ds = ctx.hist
        dl, it = SinkhornMatchFn.DELTA, SinkhornMatchFn.ITERS
        n, m, dev = K.shape[0], K.shape[1], K.device
        new = lambda *shape: torch.empty(*shape, device=dev, dtype=torch.float32)
        g = g.contiguous().float()
        b20 = Bs[-1]
        RANK = 44                       # 4 (direct term) + 20 (a_{i-1} (x) gc_i) + 19 (gd_{i-1} (x) b_{i-1}) + 1 zero
        L, Rm = torch.zeros(n, RANK, device=dev, dtype=torch.float32), torch.zeros(RANK, m, device=dev, dtype=torch.float32)
        # direct term: U = alpha . Q^T - beta with alpha = g / s, beta = (g . pts) / s
        L[:, 0:3] = g / s[:, None]
        L[:, 3] = -(g * pts).sum(-1) / s
        Rm[0:3] = query.t() * b20[None]
        Rm[3] = b20
        # gb_20 = sum_r K[r][j] U[r][j] = Q_j . (K^T alpha)_j - (K^T beta)_j: one pass with four row-weight vectors
        cols = torch.zeros(m, 4, device=dev, dtype=torch.float32)
        call("moda_sinkhorn_cols4", ptr(K), n, m, ptr(L[:, 0:4].contiguous()), ptr(cols), stream())
        gb = ((cols[:, 0:3] * query).sum(-1) + cols[:, 3]).contiguous()
        acc = torch.zeros(it, m, device=dev, dtype=torch.float32)       # gb_{i-1}: column sums of the passes
        gds = new(it - 1, n)
        # factor columns 4, 6, .. 42: a_{i-1} (x) gc_i for i = 20 .. 1; columns 5, 7, .. 41: gd_{i-1} (x) b_{i-1} for i = 20 .. 2
        for q, i in enumerate(range(it, 1, -1)):
            # gc_i = -gb_i b_i / (c_i + delta) formed while staged (and written to its row of Rm), ga = K gc_i,
            # gd = -ga a_{i-1} / (d_{i-1} + delta), gb_{i-1} = K^T gd: one pass over K
            gb_next = acc[i - 1]
            call("moda_sinkhorn_pass", ptr(K), n, m, None, None, ptr(gds[q]), ptr(gb_next), 1, 0.0, dl, ptr(As[i - 1]),
                 ptr(Ds[i - 2]), 2, ptr(Cs[i - 1]), ptr(gb), ptr(Bs[i - 1]), 0.0, ptr(Rm[4 + 2 * q]), stream())
            gb = gb_next
        Rm[4 + 2 * (it - 1)] = -gb * Bs[0] / (Cs[0] + dl)
        L[:, 4:4 + 2 * it:2] = torch.stack([As[i - 1] for i in range(it, 0, -1)], 1)
        L[:, 5:5 + 2 * (it - 1):2] = gds.t()
        Rm[5:5 + 2 * (it - 1):2] = torch.stack([Bs[i - 2] for i in range(it, 1, -1)], 0)
        gF = torch.zeros_like(feats)
        gV = torch.zeros_like(vol_feat)
        call("moda_sinkhorn_gcost", ptr(K), n, m, ptr(L), ptr(Rm), RANK, ptr(feats), ptr(vol_feat), feats.shape[1],
             SinkhornMatchFn.EPS, ptr(gF), ptr(gV), stream())
        gq = None
        if ctx.needs_input_grad[2]:
            gq = (K * b20[None] / s[:, None]).t().matmul(g)
        return gF, gV, gq


def feat_match(nerf_feat, embedding_xyz, feats, bound, grid_size=20, use_corr=True, use_ot=False, is_training=True,
               init_pts=None, rt_entropy=False):
    """loss_utils.py:273-405: soft-argmax of pixel features over the canonical feature volume sampled on a
    grid_size^3 lattice (softmax with temperature |beta|, or 20 Sinkhorn iterations at eps = 0.03 when ``use_ot``)."""
    if init_pts is not None:
        raise NotImplementedError("feat_match(init_pts=...) is only used by the reference's offline pose refinement")
    device = feats.device
    feats = F.normalize(feats, 2, -1)
    bound = np.asarray(bound, dtype=np.float32)
    pxd = np.linspace(-bound[0], bound[0], grid_size).astype(np.float32)
    pyd = np.linspace(-bound[1], bound[1], grid_size).astype(np.float32)
    pzd = np.linspace(-bound[2], bound[2], grid_size).astype(np.float32)
    query_yxz = torch.from_numpy(np.stack(np.meshgrid(pyd, pxd, pzd), -1)).to(device).reshape(-1, 3)
    query_xyz = torch.cat([query_yxz[:, 1:2], query_yxz[:, 0:1], query_yxz[:, 2:3]], -1)[None]
    if is_training:   # jitter the lattice (loss_utils.py:305-308)
        bound_t = torch.from_numpy(bound)[None, None].to(device)
        query_xyz = query_xyz + torch.randn_like(query_xyz) * bound_t * 0.05
    # canonical features on the lattice: one MLP evaluation over all grid_size^3 points (the reference chunks by 8192)
    # on the fp32 kernels in every precision mode: 8000 points cost nothing, and these features enter
    # exp((f.v - 1) / 0.03) -- an fp16-operand error of 3e-3 in v moves the matched points by ~2e-5, which the
    # positional encoding of the reprojection path (kp_reproj evaluates nerf_skin AT the matched points) multiplies by 2^9
    from . import config
    with config.exact():
        vol_feat = G.evaluate_mlp(nerf_feat, query_xyz[0][:, None], embed_xyz=embedding_xyz)[:, 0]
    vol_feat = F.normalize(vol_feat, 2, -1)
    if use_ot and not use_corr and not rt_entropy:
        return SinkhornMatchFn.apply(feats, vol_feat, query_xyz[0]), 0
    cost_vol = feats.matmul(vol_feat.t())
    if not use_ot:
        cost_vol = cost_vol * (nerf_feat.beta.abs() + 1e-9)
    if use_ot:
        K = torch.exp(-(1.0 - cost_vol[None]) / 0.03)
        n1, n2 = K.shape[1], K.shape[2]
        a = torch.ones(1, n1, 1, device=device, dtype=K.dtype) / n1
        prob1 = torch.ones(1, n1, 1, device=device, dtype=K.dtype) / n1
        prob2 = torch.ones(1, n2, 1, device=device, dtype=K.dtype) / n2
        for _ in range(20):
            KTa = torch.bmm(K.transpose(1, 2), a)
            b = prob2 / (KTa + 1e-8)
            Kb = torch.bmm(K, b)
            a = prob1 / (Kb + 1e-8)
        T_m = a * K * b.transpose(1, 2)
        prob_vol = (T_m / T_m.sum(2, keepdim=True))[0]
    else:
        prob_vol = cost_vol.softmax(-1)
    if use_corr:
        T_T = prob_vol.matmul(prob_vol.t())
        corr_err = (T_T - torch.eye(prob_vol.shape[0], device=device)).norm(2, -1)
    else:
        corr_err = 0
    pts_pred = (prob_vol[..., None] * query_xyz).sum(1)
    if rt_entropy:
        match_unc = (-prob_vol * prob_vol.clamp(1e-9, 1 - 1e-9).log()).sum(1)[:, None] / np.log(grid_size ** 3)
        return pts_pred, match_unc, corr_err
    return pts_pred, corr_err


def kp_reproj(pts_pred, models, embedding_xyz, rays, to_target=False, neudbs=True):
    """loss_utils.py:224-270: canonical points -> (forward warp of their frame) -> camera -> pixels.  (...,3) -> (N,1,2)."""
    N = pts_pred.reshape(-1, 3).shape[0]
    xyz = pts_pred.reshape(-1, 1, 3)
    rtk_vec = (rays["rtk_vec_target"] if to_target else rays["rtk_vec"]).reshape(N, -1)
    if "bones" in models:
        bone_rts_fw = (rays["bone_rts_target"] if to_target else rays["bone_rts"]).reshape(N, -1)
        bones = models["bones_rst"]
        rest_pose_code = models["rest_pose_code"](torch.zeros(1, dtype=torch.long, device=bones.device))
        if neudbs:
            dskin = G.mlp_skinning(models.get("nerf_skin"), rest_pose_code, xyz, embed_xyz=embedding_xyz, _pitched=True)
            xyz = G.warp_points(xyz, bones, bone_rts_fw, models["skin_aux"], dskin, backward=False)
        else:   # the LBS motion model (loss_utils.py:255-257)
            skin_fw = G.gauss_mlp_skinning(xyz, embedding_xyz, bones, rest_pose_code, models.get("nerf_skin"),
                                           skin_aux=models["skin_aux"])
            xyz, _ = G.lbs(bones, bone_rts_fw, skin_fw, xyz, backward=False)
    Rmat = rtk_vec[:, 0:9].reshape(N, 1, 3, 3)
    Tmat = rtk_vec[:, 9:12].reshape(N, 1, 3)
    Kinv = rtk_vec[:, 12:21].reshape(N, 1, 3, 3)
    K = G.mat2K(G.Kmatinv(Kinv))
    xyz = G.obj_to_cam(xyz, Rmat, Tmat)
    xyz = G.pinhole_cam(xyz, K)
    return xyz[..., :2]


def kp_reproj_loss(pts_pred, xys, models, embedding_xyz, rays, neudbs=True):
    """loss_utils.py:211-222."""
    xys = xys.reshape(-1, 1, 2)
    xy_reproj = kp_reproj(pts_pred, models, embedding_xyz, rays, neudbs=neudbs)
    proj_err = (xys - xy_reproj[..., :2]).norm(2, -1)
    return proj_err.reshape(pts_pred.shape[:-1] + (1,))
