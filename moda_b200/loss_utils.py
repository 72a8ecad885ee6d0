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


# ---------------------------------------------------------------------------------------- Sinkhorn feature matching
# Every (rays x lattice)-sized operation of the matching is one of five primitives.  On 16-channel fp32 CUDA features they
# are the one-pass kernels of csrc/sinkhorn.cu; on anything else (CPU tensors of the oracle tests, fp64, other feature
# widths) the same primitive is a tensor op, so the algebra of SinkhornMatchFn below exists once and is pinned on the CPU
# against autograd through the reference's loop (tests/test_oracle.py) and on the GPU against itself in fp64.
_USE_KERNELS = True   # tests switch the kernels off to compare them with the tensor-op primitives on the same device


def _sk_use_kernels(feats, vol_feat):
    m = vol_feat.shape[0]
    return (_USE_KERNELS and feats.is_cuda and feats.dtype == torch.float32 and vol_feat.dtype == torch.float32 and
            feats.shape[1] == 16 and vol_feat.shape[1] == 16 and m % 4 == 0 and m <= 8192 and feats.shape[0] > 0)


def _sk_matrix(feats, vol_feat, eps, use, c_out=None):
    """K = exp((F V^T - 1) / eps) and the first column sums c_1 = K^T a_0 with a_0 = 1/n (moda_sinkhorn_matrix)."""
    n, m = feats.shape[0], vol_feat.shape[0]
    if use:
        K = torch.empty(n, m, device=feats.device, dtype=torch.float32)
        c = c_out if c_out is not None else torch.zeros(m, device=feats.device, dtype=torch.float32)
        call("moda_sinkhorn_matrix", ptr(feats), ptr(vol_feat), n, m, feats.shape[1], eps, 1.0 / n, ptr(K), ptr(c), stream())
        return K, c
    K = torch.exp((feats.matmul(vol_feat.t()) - 1.0) / eps)
    return K, K.sum(0) / n


def _sk_pass(K, use, mode, p, delta, u=None, v=None, xmode=0, x=None, xc=None, xg=None, xb=None, xp=0.0, want_y=True,
             want_w=True, w_out=None):
    """One pass over K (moda_sinkhorn_pass): the input vector x (given, or xp / (xc + delta), or -xg xb / (xc + delta)),
    y = K x, z = p / (y + delta) (mode 0) or -y u / (v + delta) (mode 1), w = K^T z.  Returns (x, y, z, w)."""
    n, m = K.shape
    if use:
        new = lambda k: torch.empty(k, device=K.device, dtype=torch.float32)
        xo = x if xmode == 0 else new(m)
        y = new(n) if want_y else None
        z = new(n)
        w = None
        if want_w:
            w = w_out if w_out is not None else torch.zeros(m, device=K.device, dtype=torch.float32)
        call("moda_sinkhorn_pass", ptr(K), n, m, ptr(x) if xmode == 0 else None, ptr(y), ptr(z), ptr(w), mode, p, delta,
             ptr(u), ptr(v), xmode, ptr(xc), ptr(xg), ptr(xb), xp, ptr(xo) if xmode else None, stream())
        return xo, y, z, w
    if xmode == 1:
        x = xp / (xc + delta)
    elif xmode == 2:
        x = -xg * xb / (xc + delta)
    y = torch.mv(K, x)
    z = p / (y + delta) if mode == 0 else -y * u / (v + delta)
    return x, y, z, (torch.mv(K.t(), z) if want_w else None)


def _sk_rows4(K, X, use):
    """K X for X (m, 4) (moda_sinkhorn_rows4)."""
    if use:
        out = torch.empty(K.shape[0], 4, device=K.device, dtype=torch.float32)
        call("moda_sinkhorn_rows4", ptr(K), K.shape[0], K.shape[1], ptr(X.contiguous()), ptr(out), stream())
        return out
    return K.matmul(X)


def _sk_cols4(K, Wt, use):
    """K^T Wt for Wt (n, 4) (moda_sinkhorn_cols4)."""
    if use:
        out = torch.zeros(K.shape[1], 4, device=K.device, dtype=torch.float32)
        call("moda_sinkhorn_cols4", ptr(K), K.shape[0], K.shape[1], ptr(Wt.contiguous()), ptr(out), stream())
        return out
    return K.t().matmul(Wt)


def _sk_gcost(K, L, Rm, feats, vol_feat, eps, use):
    """gF = gcost V and gV = gcost^T F for gcost = K / eps * (L Rm), never materialised on the kernel path
    (moda_sinkhorn_gcost, factor rank 44)."""
    if use and L.shape[1] == 44:
        gF, gV = torch.zeros_like(feats), torch.zeros_like(vol_feat)
        call("moda_sinkhorn_gcost", ptr(K), K.shape[0], K.shape[1], ptr(L.contiguous()), ptr(Rm.contiguous()), 44, ptr(feats),
             ptr(vol_feat), feats.shape[1], eps, ptr(gF), ptr(gV), stream())
        return gF, gV
    gcost = K * L.matmul(Rm) / eps
    return gcost.matmul(vol_feat), gcost.t().matmul(feats)


class SinkhornMatchFn(torch.autograd.Function):
    """pts_pred (N,3) = rownorm(a K b^T) Q with K = exp(-(1 - F V^T) / 0.03) and (a, b) from 20 Sinkhorn iterations on
    uniform marginals (loss_utils.py:347-386) -- with the ANALYTIC adjoint of the unrolled iterations instead of an
    autograd tape over 40 (N x M) matrix-vector products (each of which would keep and differentiate a 262 MB matrix at
    8192 rays x 8000 lattice points).  Notes: `a` cancels in the row normalisation, so only b = b_20 enters the output;
    with c_i = K^T a_{i-1}, b_i = p2 / (c_i + d), d_i = K b_i, a_i = p1 / (d_i + d) the reverse sweep is
        gc_i = -gb_i b_i / (c_i + d);  gK += a_{i-1} (x) gc_i;  ga_{i-1} = K gc_i;
        gd_{i-1} = -ga_{i-1} a_{i-1} / (d_{i-1} + d);  gK += gd_{i-1} (x) b_{i-1};  gb_{i-1} = K^T gd_{i-1}
    so dLoss/dK is a sum of 39 outer products plus the direct term U (.) b_20, U[r][j] = (g_r . Q_j - g_r . pts_r) / s_r, which
    has rank 4: it is kept as factors L (N, 44), Rm (44, M) and consumed as gcost = K / eps * (L Rm) by one kernel that never
    writes it.  One pass over K per iteration (b_i from c_i while staged, d_i, a_i and c_{i+1} with the rows still on chip),
    one for the matrix itself, one for the soft-argmax, one for the direct term of the adjoint.  Verified against autograd
    through the reference's loop to 1e-12 in fp64 (tests/test_oracle.py)."""
    EPS, DELTA, ITERS = 0.03, 1e-8, 20

    @staticmethod
    def forward(ctx, feats, vol_feat, query):
        feats, vol_feat, query = feats.contiguous(), vol_feat.contiguous(), query.contiguous()
        use = _sk_use_kernels(feats, vol_feat)
        n, m, dev = feats.shape[0], vol_feat.shape[0], feats.device
        dl, it = SinkhornMatchFn.DELTA, SinkhornMatchFn.ITERS
        acc = torch.zeros(it, m, device=dev, dtype=torch.float32) if use else None    # c_1 .. c_20: one fill
        K, c = _sk_matrix(feats, vol_feat, SinkhornMatchFn.EPS, use, acc[0] if use else None)
        As, Bs, Cs, Ds = [torch.full((n,), 1.0 / n, device=dev, dtype=K.dtype)], [], [], []
        for i in range(it):
            last = i + 1 == it
            b, d, a, c_next = _sk_pass(K, use, 0, 1.0 / n, dl, xmode=1, xc=c, xp=1.0 / m, want_w=not last,
                                       w_out=acc[i + 1] if use and not last else None)
            As.append(a), Bs.append(b), Cs.append(c), Ds.append(d)
            c = c_next
        # T = a K b^T row-normalised: with the final a this is K b^T / rowsum; the reference normalises a K b^T whose row
        # sums are a_20 (K b_20): identical up to rounding.  X = b (x) [Qx Qy Qz 1]: the fourth product is the row sum again
        s = Ds[-1]
        X = torch.cat([query, torch.ones_like(query[:, :1])], 1) * Bs[-1][:, None]
        pts = _sk_rows4(K, X, use)[:, :3] / s[:, None]
        ctx.use = use
        ctx.save_for_backward(feats, vol_feat, query, K, pts, s)
        ctx.hist = (As, Bs, Cs, Ds)
        return pts

    @staticmethod
    def backward(ctx, g):
        feats, vol_feat, query, K, pts, s = ctx.saved_tensors
        As, Bs, Cs, Ds = ctx.hist
        use = ctx.use
        dl, it = SinkhornMatchFn.DELTA, SinkhornMatchFn.ITERS
        n, m, dev = K.shape[0], K.shape[1], K.device
        g = g.contiguous().to(K.dtype)
        b20 = Bs[-1]
        # direct term: dL / d(K_rj b_j) = U[r][j] = alpha_r . Q_j - beta_r with alpha = g / s, beta = (g . pts) / s
        Wt = torch.cat([g / s[:, None], (-(g * pts).sum(-1) / s)[:, None]], 1)          # (n, 4) = [alpha | -beta]
        # gb_20 = sum_r K[r][j] U[r][j] = Q_j . (K^T alpha)_j - (K^T beta)_j: one pass with four row-weight vectors
        cols = _sk_cols4(K, Wt, use)
        gb = ((cols[:, 0:3] * query).sum(-1) + cols[:, 3]).contiguous()
        acc = torch.zeros(it, m, device=dev, dtype=torch.float32) if use else None      # gb_{i-1}: column sums of the passes
        left, right = [], []            # dLoss/dK = U (.) b_20 + sum_k left_k (x) right_k
        for i in range(it, 1, -1):
            # gc_i = -gb_i b_i / (c_i + delta) formed while staged, ga = K gc_i, gd = -ga a_{i-1} / (d_{i-1} + delta),
            # gb_{i-1} = K^T gd: one pass over K
            gc, _, gd, gb = _sk_pass(K, use, 1, 0.0, dl, u=As[i - 1], v=Ds[i - 2], xmode=2, xc=Cs[i - 1], xg=gb, xb=Bs[i - 1],
                                     want_y=False, w_out=acc[i - 1] if use else None)
            left += [As[i - 1], gd]
            right += [gc, Bs[i - 2]]
        left.append(As[0])
        right.append(-gb * Bs[0] / (Cs[0] + dl))
        rank = 4 + len(left)
        pad = (-rank) % 4
        L = torch.cat([Wt, torch.stack(left, 1), torch.zeros(n, pad, device=dev, dtype=K.dtype)], 1)
        Rm = torch.cat([query.t() * b20[None], b20[None], torch.stack(right, 0), torch.zeros(pad, m, device=dev, dtype=K.dtype)], 0)
        gF, gV = _sk_gcost(K, L, Rm, feats, vol_feat, SinkhornMatchFn.EPS, use)
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
