"""TEST INFRASTRUCTURE ONLY -- CPU restatement (the oracle) of MoDA's articulated volume renderer.

This file is the checker for the CUDA path in ``moda_b200/``; it is never imported by the product
(only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs).  It needs only torch (CPU) and works in fp32 or fp64 (dtype follows the inputs).
Gradients come from torch autograd on this restatement.

Parity status: PINNED BY EXECUTING THE REFERENCE.  The reference has no tests or golden vectors of
its own for this path; ``oracle/make_golden.py`` runs the real reference code
(imported from /root/reference in the build container) on the seeded inputs of
``moda_b200/synth.py`` and commits its outputs and gradients under ``tests/golden/``;
``tests/test_oracle.py`` checks this restatement against those files (and against the live
reference when /root/reference is present).

Each function cites the reference lines it restates (paths relative to /root/reference).  The
functional style (state dicts instead of nn.Modules) is this repo's own; op order within a formula
follows the reference so that fp32 results agree to rounding.  Chunk sizes are the reference's
(4096 / 8192 rays) so that timing this file is a fair stand-in for timing the reference on a CPU.
"""
import math

import torch
import torch.nn.functional as F

# ------------------------------------------------------------------------------------------------
# positional encoding -- nnutils/nerf.py:35-75


def pe_window(n_freqs, alpha, dtype=torch.float32, device=None):
    """nerf.py:63-66: w_k = 0.5 * (1 + cos(pi * clamp(alpha - k, 0, 1) + pi))."""
    k = torch.arange(n_freqs, dtype=dtype, device=device)
    return 0.5 * (1 + torch.cos(math.pi * torch.clamp(alpha - k, 0.0, 1.0) + math.pi))


def embed(x, n_freqs, alpha=None):
    """nerf.py:47-75.  Layout: [x | w0 sin(x) | w0 cos(x) | w1 sin(2x) | w1 cos(2x) | ...]."""
    if n_freqs <= 0:
        return x
    if alpha is None:
        alpha = n_freqs
    c = x.shape[-1]
    flat = x.reshape(-1, c)
    win = pe_window(n_freqs, float(alpha), dtype=x.dtype, device=x.device)
    parts = [flat]
    for k in range(n_freqs):
        f = float(2 ** k)
        parts.append(win[k] * torch.sin(f * flat))
        parts.append(win[k] * torch.cos(f * flat))
    return torch.cat(parts, -1).reshape(x.shape[:-1] + (c * (1 + 2 * n_freqs),))


# ------------------------------------------------------------------------------------------------
# the D x W MLP -- nnutils/nerf.py:147-198


class NerfSpec:
    """Shape description of one ``NeRF`` instance (nerf.py:84-105)."""

    def __init__(self, D=8, W=256, in_channels_xyz=63, in_channels_dir=27, out_channels=3,
                 skips=(4,), raw_feat=False):
        self.D, self.W = D, W
        self.in_channels_xyz, self.in_channels_dir = in_channels_xyz, in_channels_dir
        self.out_channels, self.skips, self.raw_feat = out_channels, tuple(skips), raw_feat


COARSE_SPEC = NerfSpec(8, 256, 63, 27 + 64, 3, (4,), False)
SKIN_SPEC = NerfSpec(5, 64, 63 + 128, 0, 25, (4,), True)
FEAT_SPEC = NerfSpec(5, 128, 63, 0, 16, (4,), True)   # nerf_feat, moda.py:447-449
VIS_SPEC = NerfSpec(5, 64, 63, 0, 1, (4,), True)      # nerf_vis, moda.py:344-348


def nerf_forward(sd, spec, x, sigma_only=False):
    """nerf.py:163-198.  ``sd`` uses the reference's state-dict names."""
    cx = spec.in_channels_xyz
    inp = x[..., :cx]
    h = inp
    for i in range(spec.D):
        if i in spec.skips:
            h = torch.cat([inp, h], -1)
        h = torch.relu(F.linear(h, sd["xyz_encoding_%d.0.weight" % (i + 1)],
                                sd["xyz_encoding_%d.0.bias" % (i + 1)]))
    sigma = F.linear(h, sd["sigma.weight"], sd["sigma.bias"])
    if sigma_only:
        return sigma
    fin = F.linear(h, sd["xyz_encoding_final.weight"], sd["xyz_encoding_final.bias"])
    d_in = torch.cat([fin, x[..., cx:cx + spec.in_channels_dir]], -1)
    dfe = torch.relu(F.linear(d_in, sd["dir_encoding.0.weight"], sd["dir_encoding.0.bias"]))
    rgb = F.linear(dfe, sd["rgb.0.weight"], sd["rgb.0.bias"])
    if spec.raw_feat:
        return rgb
    return torch.cat([torch.sigmoid(rgb), sigma], -1)


def evaluate_mlp(sd, spec, pts, n_freqs=None, alpha=None, dir_embedded=None, code=None,
                 chunk=32 * 1024, sigma_only=False):
    """geom_utils.py:19-57: ray-chunked [PE | dir | code] assembly then the MLP.

    pts (R,S,k); dir_embedded (R,S,c) or None; code (R,c) / (R,1,c) / (1,c) or None.
    """
    R, S, _ = pts.shape
    if code is not None and code.dim() == 2 and code.shape[0] != R:
        code = code.repeat(R, 1)
    outs = []
    for i in range(0, R, chunk):
        e = pts[i:i + chunk]
        if n_freqs is not None:
            e = embed(e, n_freqs, alpha)
        if dir_embedded is not None:
            e = torch.cat([e, dir_embedded[i:i + chunk]], -1)
        if code is not None:
            cc = code[i:i + chunk]
            if cc.dim() == 2:
                cc = cc[:, None]
            e = torch.cat([e, cc.repeat(1, S, 1)], -1)
        outs.append(nerf_forward(sd, spec, e, sigma_only=sigma_only))
    return torch.cat(outs, 0)


# ------------------------------------------------------------------------------------------------
# quaternion helpers -- third_party/pytorch3d/pytorch3d/transforms/rotation_conversions.py


def quat_to_matrix(q):
    """rotation_conversions.py:41-69 (two_s = 2/|q|^2: no normalisation required)."""
    r, i, j, k = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    m = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return m.reshape(q.shape[:-1] + (3, 3))


def quat_raw_mul(a, b):
    """rotation_conversions.py:374-392 (Hamilton product, real part first)."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw), -1)


def quat_mul_std(a, b):
    """rotation_conversions.py:359-409: product flipped so that the real part is >= 0."""
    ab = quat_raw_mul(a, b)
    return torch.where(ab[..., 0:1] < 0, -ab, ab)


# ------------------------------------------------------------------------------------------------
# dual quaternions -- nnutils/dual_quat.py


def q_normalize(q):
    """dual_quat.py:4-12."""
    return q / torch.sqrt((q * q).sum(-1))[..., None]


def q_mul(q1, q2):
    """dual_quat.py:14-31 (== q1 (x) q2, verified against the outer-product form)."""
    return quat_raw_mul(q1, q2)


def dq_mul(a, b):
    """dual_quat.py:33-49: (r1 r2, r1 d2 + d1 r2)."""
    ar, ad = a[..., :4], a[..., 4:]
    br, bd = b[..., :4], b[..., 4:]
    return torch.cat([q_mul(ar, br), q_mul(ar, bd) + q_mul(ad, br)], -1)


def dq_normalize(dq):
    """dual_quat.py:51-62: all 8 components divided by the norm of the real quaternion."""
    return dq / torch.norm(dq[..., :4], dim=-1, keepdim=True)


_CONJ_Q = (1.0, -1.0, -1.0, -1.0, 1.0, -1.0, -1.0, -1.0)
_CONJ_C = (1.0, -1.0, -1.0, -1.0, -1.0, 1.0, 1.0, 1.0)


def dq_quaternion_conjugate(dq):
    """dual_quat.py:65-74."""
    return dq * torch.tensor(_CONJ_Q, dtype=dq.dtype, device=dq.device)


def dq_combined_conjugate(dq):
    """dual_quat.py:76-85."""
    return dq * torch.tensor(_CONJ_C, dtype=dq.dtype, device=dq.device)


def dq_inverse(dq):
    """dual_quat.py:87-93: conj_q(dq) / |r|^2."""
    return dq_quaternion_conjugate(dq) / (dq[..., :4] ** 2).sum(-1, keepdim=True)


# ------------------------------------------------------------------------------------------------
# Gaussian bones, skinning weights, DQ blending -- nnutils/geom_utils.py


def bone_transform(bones, rts):
    """geom_utils.py:59-111, neudbs branch 73-86.  bones (...,B,10), rts (R, B*8) -> (R,B,10)."""
    B = bones.shape[-2]
    b = bones.reshape(-1, B, 10)
    rts = rts.reshape(-1, B, 8)
    R = rts.shape[0]
    qr, qd = rts[..., :4], rts[..., 4:]
    rot = quat_to_matrix(qr)
    conj = qr * torch.tensor((1.0, -1.0, -1.0, -1.0), dtype=qr.dtype, device=qr.device)
    trn = (2 * quat_raw_mul(qd, conj))[..., 1:]
    center = rot.matmul(b[:, :, :3, None])[..., 0] + trn
    orient = quat_mul_std(qr, b[:, :, 3:7])
    scale = b[:, :, 7:10].repeat(R // b.shape[0], 1, 1) if b.shape[0] != R else b[:, :, 7:10]
    return torch.cat([center, orient, scale], -1)


def vec_to_sim3(vec):
    """geom_utils.py:187-199."""
    center = vec[..., :3]
    orient = quat_to_matrix(F.normalize(vec[..., 3:7], 2, -1))
    scale = vec[..., 7:10].exp()
    return center, orient, scale


def _skinning_chunk(bones, pts, dskin, skin_aux):
    """geom_utils.py:237-277."""
    R, S, _ = pts.shape
    B = bones.shape[-2]
    center, orient, scale = vec_to_sim3(bones)
    orient_t = orient.permute(0, 1, 3, 2)
    diff = center.view(R, 1, B, 3) - pts.view(R, S, 1, 3)
    # axis_rotate (geom_utils.py:231-235): rows of R^T dotted with diff
    m = (orient_t.view(R, 1, B, 3, 3) * diff.view(R, S, B, 1, 3)).sum(4)
    m = scale.view(R, 1, B, 3) * m.pow(2)
    m = m * 100 * skin_aux[0].exp()
    logit = -10 * m.sum(3)
    if dskin is not None:
        logit = logit + dskin
    return logit.softmax(2)


def skinning(bones, pts, dskin=None, skin_aux=None, chunk=4096):
    """geom_utils.py:280-302.  bones (B,10) or (R,B,10); pts (R,S,3) -> (R,S,B)."""
    R = pts.shape[0]
    B = bones.shape[-2]
    if bones.dim() == 2:
        bones = bones[None].repeat(R, 1, 1)
    bones = bones.reshape(-1, B, 10)
    out = []
    for i in range(0, R, chunk):
        out.append(_skinning_chunk(bones[i:i + chunk], pts[i:i + chunk],
                                   None if dskin is None else dskin[i:i + chunk], skin_aux))
    return torch.cat(out, 0)


def gauss_mlp_skinning(xyz, n_freqs, alpha, bones, pose_code, skin_sd, skin_spec, skin_aux):
    """geom_utils.py:202-229: Delta logits from nerf_skin(PE(xyz) | pose_code), then ``skinning``."""
    R = xyz.shape[0]
    if pose_code.dim() == 2 and pose_code.shape[0] != R:
        pose_code = pose_code[None].repeat(R, 1, 1)
    dskin = None
    if skin_sd is not None:
        dskin = evaluate_mlp(skin_sd, skin_spec, embed(xyz, n_freqs, alpha), code=pose_code,
                             chunk=8 * 1024)
    return skinning(bones, xyz, dskin, skin_aux)


def _dqs_blend_chunk(dq, skin, pts):
    """geom_utils.py:457-493."""
    b = (skin[..., None] * dq[:, None]).sum(2)
    c = dq_normalize(b)
    a0, d0 = c[..., 0:1], c[..., 1:4]
    ae, de = c[..., 4:5], c[..., 5:8]
    trans = 2 * (a0 * de - ae * d0 + torch.cross(d0, de, dim=-1))
    rot = pts + 2 * torch.cross(d0, torch.cross(d0, pts, dim=-1) + a0 * pts, dim=-1)
    return rot + trans


def dqs_blend_skinning(dq, skin, pts, chunk=4096):
    """geom_utils.py:495-517."""
    B = dq.shape[-2]
    S = pts.shape[-2]
    pts = pts.reshape(-1, S, 3)
    dq = dq.reshape(-1, B, 8)
    out = []
    for i in range(0, pts.shape[0], chunk):
        out.append(_dqs_blend_chunk(dq[i:i + chunk], skin[i:i + chunk], pts[i:i + chunk]))
    return torch.cat(out, 0)


def neu_dbs(bones, rts_fw, skin, xyz_in, backward=True):
    """geom_utils.py:372-456 without the optional ``nerf_dis`` residual."""
    B = bones.shape[-2]
    rts = rts_fw.reshape(-1, B, 8)
    bones_dfm = bone_transform(bones.reshape(-1, B, 10), rts)
    dq = dq_inverse(rts) if backward else rts
    return dqs_blend_skinning(dq, skin, xyz_in), bones_dfm


# ------------------------------------------------------------------------------------------------
# volume rendering -- nnutils/rendering.py


def sample_depths(near, far, n_samples, use_disp=False, perturb=0.0, perturb_rand=None):
    """rendering.py:68-83.  ``perturb_rand`` is the U[0,1) tensor the reference draws at :82."""
    s = torch.linspace(0, 1, n_samples, device=near.device, dtype=near.dtype)
    if not use_disp:
        z = near * (1 - s) + far * s
    else:
        z = 1 / (1 / near * (1 - s) + 1 / far * s)
    z = z.expand(near.shape[0], n_samples)
    if perturb > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat([mid, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mid], -1)
        if perturb_rand is None:
            perturb_rand = torch.rand(z.shape, device=z.device, dtype=z.dtype)
        z = lower + (upper - lower) * (perturb * perturb_rand)
    return z


def density_to_alpha(sigma_raw, deltas, beta):
    """rendering.py:199-207 (VolSDF-style Laplace CDF of the signed distance -sigma)."""
    ibeta = 1 / (beta.abs() + 1e-9)
    sdf = -sigma_raw
    dens = (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() * ibeta)) * ibeta
    return 1 - torch.exp(-deltas * dens)


def composite(rgbs, sigma_raw, z_vals, rays_d, beta, noise=None, alpha_mask=None):
    """rendering.py:183-235.  Returns rgb (R,3), depth (R,), sil (R,), weights (R,S), visibility."""
    deltas = z_vals[:, 1:] - z_vals[:, :-1]
    deltas = torch.cat([deltas, 1e10 * torch.ones_like(deltas[:, :1])], -1)
    deltas = deltas * torch.norm(rays_d.unsqueeze(1), dim=-1)
    if noise is not None:
        sigma_raw = sigma_raw + noise
    alphas = density_to_alpha(sigma_raw, deltas, beta)
    if alpha_mask is not None:  # rendering.py:210-215 (out-of-bound / invisible samples)
        alphas = torch.where(alpha_mask, torch.zeros_like(alphas), alphas)
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1 - alphas + 1e-10], -1)
    trans = torch.cumprod(shifted, -1)[:, :-1]
    weights = alphas * trans
    rgb = (weights.unsqueeze(-1) * rgbs).sum(-2)
    depth = (weights * z_vals).sum(-1)
    sil = weights[:, :-1].sum(-1)
    return rgb, depth, sil, weights, trans.detach()


def sample_pdf(bins, weights, n_importance, det=False, eps=1e-5, u=None):
    """rendering.py:582-623."""
    R, n = weights.shape
    weights = weights + eps
    pdf = weights / weights.sum(-1, keepdim=True)
    cdf = torch.cumsum(pdf, -1)
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    if u is None:
        if det:
            u = torch.linspace(0, 1, n_importance, device=bins.device, dtype=bins.dtype).expand(R, n_importance)
        else:
            u = torch.rand(R, n_importance, device=bins.device, dtype=bins.dtype)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, n)
    ig = torch.stack([below, above], -1).view(R, 2 * n_importance)
    cdf_g = torch.gather(cdf, 1, ig).view(R, n_importance, 2)
    bins_g = torch.gather(bins, 1, ig).view(R, n_importance, 2)
    denom = cdf_g[..., 1] - cdf_g[..., 0]
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    return bins_g[..., 0] + (u - cdf_g[..., 0]) / denom * (bins_g[..., 1] - bins_g[..., 0])


# ------------------------------------------------------------------------------------------------
# alternative motion models: LBS and the free-form flow fields
def matrix_to_quat(m):
    """pytorch3d matrix_to_quaternion (rotation_conversions.py:101-152): candidate with the largest pivot."""
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.reshape(m.shape[:-2] + (9,)).unbind(-1)
    t = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1)
    qa = torch.zeros_like(t)
    qa[t > 0] = torch.sqrt(t[t > 0])
    rows = [torch.stack([qa[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
            torch.stack([m21 - m12, qa[..., 1] ** 2, m10 + m01, m02 + m20], -1),
            torch.stack([m02 - m20, m10 + m01, qa[..., 2] ** 2, m12 + m21], -1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, qa[..., 3] ** 2], -1)]
    cands = torch.stack(rows, -2) / (2.0 * qa[..., None].clamp_min(0.1))
    pick = F.one_hot(qa.argmax(-1), 4) > 0.5
    return cands[pick, :].reshape(m.shape[:-2] + (4,))


def bone_transform_rigid(bones, rts):
    """geom_utils.py:87-107 (``neudbs=False``, is_vec): bones (B,10), rts (bs,B,12) [R | T] -> (bs,B,10)."""
    B = bones.shape[-2]
    rts = rts.reshape(-1, B, 12)
    Rm, Tm = rts[..., :9].reshape(-1, B, 3, 3), rts[..., 9:]
    b = bones.reshape(1, B, 10)
    center = Rm.matmul(b[..., :3, None].expand(Rm.shape[0], B, 3, 1))[..., 0] + Tm
    orient = quat_mul_std(matrix_to_quat(Rm), b[..., 3:7].expand(Rm.shape[0], B, 4))
    return torch.cat([center, orient, b[..., 7:].expand(Rm.shape[0], B, 3)], -1)


def rts_invert(rts):
    """geom_utils.py:142-153."""
    Ri = rts[..., :3].transpose(-1, -2)
    return torch.cat([Ri, -Ri.matmul(rts[..., 3:])], -1)


def blend_skinning(rts, skin, pts):
    """geom_utils.py:304-326: rts (bs,B,3,4), skin (bs,N,B), pts (bs,N,3)."""
    Rw = (skin[..., None, None] * rts[:, None, :, :, :3]).sum(2)
    Tw = (skin[..., None] * rts[:, None, :, :, 3]).sum(2)
    return Rw.matmul(pts[..., None])[..., 0] + Tw


def lbs(bones, rts_fw, skin, xyz, backward=True):
    """geom_utils.py:906-931."""
    B = bones.shape[-2]
    v = rts_fw.reshape(-1, B, 12)
    rts = torch.cat([v[..., :9].reshape(-1, B, 3, 3), v[..., 9:, None]], -1)
    return blend_skinning(rts_invert(rts) if backward else rts, skin, xyz)


def so3_exp(w, eps=1e-4):
    """pytorch3d so3_exponential_map (so3.py:148-176)."""
    ang = (w * w).sum(-1).clamp(eps).sqrt()
    K = torch.zeros(w.shape[0], 3, 3, dtype=w.dtype)
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -w[:, 2], w[:, 1], w[:, 2], -w[:, 0], -w[:, 1], w[:, 0]
    return (ang.sin() / ang)[:, None, None] * K + ((1 - ang.cos()) / ang ** 2)[:, None, None] * K.bmm(K) + torch.eye(3, dtype=w.dtype)


FLOW_SPEC = {"trans": NerfSpec(5, 128, 63 + 128, 0, 3, (4,), True), "se3": NerfSpec(5, 128, 63 + 128, 0, 9, (4,), True)}


def flow_field(sd, kind, xyz, code, n_freqs, alpha):
    """Transhead / SE3head on [PE(xyz) | code] (nerf.py:200-237; evaluate_mlp call sites rendering.py:262-283)."""
    raw = evaluate_mlp(sd, FLOW_SPEC[kind], xyz, n_freqs, alpha, code=code, chunk=32768)
    if kind == "trans":
        return raw * 0.1
    r = raw.reshape(-1, 9)
    pivot, trans = r[:, 3:6] * 0.1, r[:, 6:9] * 0.1
    p = xyz.reshape(-1, 3) + pivot
    w = so3_exp(r[:, :3]).matmul(p[..., None])[..., 0] - pivot + trans
    return w.reshape(xyz.shape) - xyz


def _deform_and_render(prob, xyz, z_vals, dir_embedded, n_freqs, alpha, fine_iter, noise=None,
                       vis_sd=None, vis_spec=None, obj_bound=None):
    """rendering.py:239-579 restricted to the core path, plus the LBS / flow-field motion
    models when ``prob['motion']`` names one (synth.make_motion_problem)."""
    rays = prob["rays"]
    R, S, _ = xyz.shape
    res = {}
    xyz_frame = xyz
    motion = prob.get("motion")
    if motion in ("trans", "se3"):   # rendering.py:257-274
        time_emb = rays["time_embedded"][:, None]
        flow_bw = flow_field(prob["flowbw"], motion, xyz, time_emb, n_freqs, alpha)
        xyz = xyz + flow_bw
        if fine_iter:
            cyc = (flow_bw + flow_field(prob["flowfw"], motion, xyz, time_emb, n_freqs, alpha)).norm(2, -1)
    elif motion == "lbs":   # rendering.py:303-341 with opts.lbs
        bones_rst, rts_fw, skin_aux = prob["bones_rst"], rays["bone_rts"], prob["skin_aux"]
        spec = prob.get("skin_spec", SKIN_SPEC)
        bones_dfm = bone_transform_rigid(bones_rst, rts_fw)
        skin_bw = gauss_mlp_skinning(xyz, n_freqs, alpha, bones_dfm, rays["time_embedded"][:, None], prob["nerf_skin"],
                                     spec, skin_aux)
        xyz = lbs(bones_rst, rts_fw, skin_bw, xyz, backward=True)
        if fine_iter:
            skin_fw = gauss_mlp_skinning(xyz, n_freqs, alpha, bones_rst, prob["rest_pose_code"][0:1], prob["nerf_skin"],
                                         spec, skin_aux)
            cyc = (xyz_frame - lbs(bones_rst, rts_fw, skin_fw, xyz, backward=False)).norm(2, -1)
    elif prob.get("bones_rst") is not None:
        bones_rst = prob["bones_rst"]
        rts_fw = rays["bone_rts"]
        skin_aux = prob["skin_aux"]
        rest_code = prob["rest_pose_code"][0:1]  # rest_pose_code(LongTensor([0])), rendering.py:293-294
        skin_sd = prob.get("nerf_skin")
        spec = prob.get("skin_spec", SKIN_SPEC)
        time_emb = rays["time_embedded"][:, None]
        bones_dfm = bone_transform(bones_rst, rts_fw)
        skin_bw = gauss_mlp_skinning(xyz, n_freqs, alpha, bones_dfm, time_emb, skin_sd, spec, skin_aux)
        xyz, _ = neu_dbs(bones_rst, rts_fw, skin_bw, xyz, backward=True)
        if fine_iter:
            skin_fw = gauss_mlp_skinning(xyz, n_freqs, alpha, bones_rst, rest_code, skin_sd, spec, skin_aux)
            xyz_cyc, _ = neu_dbs(bones_rst, rts_fw, skin_fw, xyz, backward=False)
            cyc = (xyz_frame - xyz_cyc).norm(2, -1)
    alpha_mask = None
    if vis_sd is not None:  # render_vis branch, rendering.py:373-379, 210-215
        vis_pred = evaluate_mlp(vis_sd, vis_spec, embed(xyz, n_freqs, alpha), chunk=32768)[..., 0].sigmoid()
        cb = torch.as_tensor(obj_bound, dtype=xyz.dtype, device=xyz.device)[None, None]
        alpha_mask = ((xyz.abs() > cb).sum(-1) > 0) | (vis_pred < 0.5)
    dir_rep = dir_embedded[:, None].expand(R, S, dir_embedded.shape[-1])
    out = evaluate_mlp(prob["coarse"], prob.get("coarse_spec", COARSE_SPEC), xyz, n_freqs, alpha,
                       dir_embedded=dir_rep, code=rays.get("env_code"), chunk=4096)
    rgbs, sig = out[..., :3], out[..., 3]
    rgb, depth, sil, weights, vis = composite(rgbs, sig, z_vals, rays["rays_d"], prob["coarse"]["beta"],
                                              noise=noise, alpha_mask=alpha_mask)
    res["img_coarse"], res["depth_rnd"], res["sil_coarse"] = rgb, depth, sil
    if vis_sd is not None:
        res["vis_pred"] = (vis_pred * weights).sum(-1)
    if fine_iter:
        res["xyz_camera_vis"] = xyz_frame
        if prob.get("bones_rst") is not None or motion in ("trans", "se3"):
            res["xyz_canonical_vis"] = xyz
            res["frame_cyc_dis"] = (cyc * weights.detach()).sum(-1)
    return res, weights


def render_rays(prob, n_samples=128, use_disp=False, perturb=0.0, perturb_rand=None, noise=None,
                use_fine=False, xyz_freqs=10, dir_freqs=4, alpha=10, vis_sd=None, vis_spec=None,
                obj_bound=None, pdf_u=None):
    """rendering.py:19-122.  ``prob`` is the dict of ``moda_b200.synth.make_problem``."""
    rays = prob["rays"]
    if use_fine:
        n_samples = n_samples // 2
    d = rays["rays_d"]
    dir_embedded = embed(d / d.norm(2, -1)[:, None], dir_freqs, alpha)
    z = sample_depths(rays["near"], rays["far"], n_samples, use_disp, perturb, perturb_rand)
    xyz = rays["rays_o"].unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(2)
    if use_fine:
        with torch.no_grad():
            _, w = _deform_and_render(prob, xyz, z, dir_embedded, xyz_freqs, alpha, False, noise=None,
                                      vis_sd=None)
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        z_new = sample_pdf(mid, w[:, 1:-1], n_samples, det=(perturb == 0), u=pdf_u).detach()
        z, _ = torch.sort(torch.cat([z, z_new], -1), -1)
        xyz = rays["rays_o"].unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(2)
    res, _ = _deform_and_render(prob, xyz, z, dir_embedded, xyz_freqs, alpha, True, noise=noise,
                                vis_sd=vis_sd, vis_spec=vis_spec, obj_bound=obj_bound)
    return res


def parity_loss(res):
    """The scalar used for gradient parity."""
    return ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() \
        + res["frame_cyc_dis"].mean()


def skin_warp_roundtrip(bones_rst, bone_rts, skin_aux, xyz):
    """BASELINE config 4: Gaussian weights + backward warp, then weights + forward warp, no Delta MLP."""
    bones_dfm = bone_transform(bones_rst, bone_rts)
    w_bw = skinning(bones_dfm, xyz, None, skin_aux)
    xyz_can, _ = neu_dbs(bones_rst, bone_rts, w_bw, xyz, backward=True)
    w_fw = skinning(bones_rst, xyz_can, None, skin_aux)
    xyz_cyc, _ = neu_dbs(bones_rst, bone_rts, w_fw, xyz_can, backward=False)
    return xyz_can, xyz_cyc


def density_grid(coarse_sd, grid_size, bound, spec=COARSE_SPEC, n_freqs=10, alpha=10, chunk=32768,
                 x_range=None, vis_sd=None, symm_shape=False):
    """train_utils.py:1377-1425: sigma on a G^3 lattice, (x,y,z) C-order, 32768-point chunks; with ``vis_sd`` the
    visibility pass (:1407-1425: density -1 where sigmoid(nerf_vis) < 0.5), with ``symm_shape`` evaluation at |x| (:1398)."""
    G = grid_size
    dt = coarse_sd["sigma.weight"].dtype
    ax = [torch.linspace(-float(b), float(b), G, dtype=dt) for b in bound]
    xs = ax[0] if x_range is None else ax[0][x_range[0]:x_range[1]]
    pts = torch.stack(torch.meshgrid(xs, ax[1], ax[2], indexing="ij"), -1).reshape(-1, 3)
    out = []
    for i in range(0, pts.shape[0], chunk):
        q = pts[i:i + chunk]
        if symm_shape:
            q = torch.cat([q[:, :1].abs(), q[:, 1:]], -1)
        sig = nerf_forward(coarse_sd, spec, embed(q, n_freqs, alpha), sigma_only=True)
        if vis_sd is not None:
            vis = nerf_forward(vis_sd, VIS_SPEC, embed(pts[i:i + chunk], n_freqs, alpha))[..., 0].sigmoid()
            sig = torch.where(vis[:, None] < 0.5, torch.full_like(sig, -1.0), sig)
        out.append(sig)
    return torch.cat(out, 0).reshape(len(xs), G, G)


def to_dtype(prob, dtype):
    """Deep-copies a synth problem to another dtype (fp64 oracle runs) with grads enabled on leaves."""
    def conv(v):
        if isinstance(v, dict):
            return {k: conv(x) for k, x in v.items()}
        if torch.is_tensor(v) and v.is_floating_point():
            return v.detach().clone().to(dtype)
        return v
    return conv(prob)


GRAD_PARAMS = ("bones_rst", "skin_aux", "rest_pose_code")
GRAD_RAYS = ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d")


def require_grads(prob):
    leaves = {}
    for k in GRAD_PARAMS:
        if k in prob:
            prob[k].requires_grad_(True)
            leaves[k] = prob[k]
    for net in ("coarse", "nerf_skin", "flowbw", "flowfw"):
        for k, v in prob.get(net, {}).items():
            v.requires_grad_(True)
            leaves[net + "." + k] = v
    for k in GRAD_RAYS:
        if k in prob["rays"]:
            prob["rays"][k].requires_grad_(True)
            leaves["rays." + k] = prob["rays"][k]
    return leaves


# ------------------------------------------------------------------------------------------------
# round 2: the default-flag training step -- nerf_feat / nerf_vis, feature matching,
# key-point reprojection, the third warp with flow rendering, per-ray loss terms

def obj_to_cam(verts, Rmat, Tmat):
    """geom_utils.py:567-581.  verts (N,S,3), Rmat (N,3,3), Tmat (N,3)."""
    return verts.matmul(Rmat.transpose(1, 2)) + Tmat[:, None]


def rtk_split(rtk_vec):
    """rendering.py:437-441 / loss_utils.py:259-262: (N,21) -> R (N,3,3), T (N,3), K (N,4) = (fx, fy, px, py) recovered
    from Kinv via Kmatinv + mat2K (geom_utils.py:612-652)."""
    N = rtk_vec.shape[0]
    Rmat = rtk_vec[:, 0:9].reshape(N, 3, 3)
    Tmat = rtk_vec[:, 9:12]
    Kinv = rtk_vec[:, 12:21].reshape(N, 3, 3)
    k = torch.stack([Kinv[:, 0, 0], Kinv[:, 1, 1], Kinv[:, 0, 2], Kinv[:, 1, 2]], -1)      # mat2K(Kinv)
    K = torch.stack([1.0 / k[:, 0], 1.0 / k[:, 1], -k[:, 2] / k[:, 0], -k[:, 3] / k[:, 1]], -1)  # mat2K(K2inv(k))
    return Rmat, Tmat, K


def pinhole_cam(verts, K):
    """geom_utils.py:654-672.  verts (N,S,3), K (N,4)."""
    x = verts[..., 0] * K[:, None, 0] + verts[..., 2] * K[:, None, 2]
    y = verts[..., 1] * K[:, None, 1] + verts[..., 2] * K[:, None, 3]
    z = verts[..., 2]
    return torch.stack([x / (1e-6 + z), y / (1e-6 + z), z], -1)


def project(xyz, rtk_vec):
    Rmat, Tmat, K = rtk_split(rtk_vec)
    return pinhole_cam(obj_to_cam(xyz, Rmat, Tmat), K)


def vrender_flo(weights, xyz_target, xys, img_size):
    """geom_utils.py:1704-1743."""
    xy = xyz_target[..., :2]
    invalid = (xyz_target[..., 2] < 1e-5) | (xy.norm(2, -1) > 2 * img_size)
    w = torch.where(invalid, torch.zeros_like(weights), weights)
    xy = torch.where(invalid[..., None], torch.zeros_like(xy), xy)
    w = w / (1e-9 + w.sum(-1, keepdim=True))
    flo = (w[..., None] * (xy - xys[:, None])).sum(-2) / img_size * 2
    return flo, (invalid.sum(-1) == 0).to(weights.dtype)[..., None]


def compute_pts_exp(pts_prob, pts):
    """loss_utils.py:162-172."""
    p = pts_prob / (1e-9 + pts_prob.sum(-1, keepdim=True))
    return (pts * p[..., None]).sum(1)


def feat_match(feat_sd, feats, bound, beta, n_freqs=10, alpha=10, grid_size=20, use_ot=True, noise=None,
               spec=FEAT_SPEC):
    """loss_utils.py:273-405 (use_corr off).  noise: the N(0,1) draw of :308 (1, grid^3, 3) or None (eval mode)."""
    dt = feats.dtype
    feats = F.normalize(feats, 2, -1)
    ax = [torch.linspace(-float(b), float(b), grid_size, dtype=torch.float32).to(dt) for b in bound]
    # np.meshgrid(pyd, pxd, pzd) (default 'xy' indexing) -> (y,x,z) columns, then reordered to (x,y,z): :296-298
    gy, gx, gz = torch.meshgrid(ax[1], ax[0], ax[2], indexing="xy")
    q = torch.stack([gx, gy, gz], -1).reshape(-1, 3)
    if noise is not None:
        q = q + noise.reshape(-1, 3).to(dt) * torch.as_tensor([float(b) for b in bound], dtype=dt)[None] * 0.05
    vol = nerf_forward(feat_sd, spec, embed(q, n_freqs, alpha))
    vol = F.normalize(vol, 2, -1)
    cost = feats.matmul(vol.t())
    if use_ot:
        K = torch.exp(-(1.0 - cost) / 0.03)
        n1, n2 = K.shape
        a = torch.full((n1, 1), 1.0 / n1, dtype=dt)
        for _ in range(20):
            b = (1.0 / n2) / (K.t().matmul(a) + 1e-8)
            a = (1.0 / n1) / (K.matmul(b) + 1e-8)
        T = a * K * b.t()
        prob = T / T.sum(1, keepdim=True)
    else:
        prob = (cost * (beta.abs() + 1e-9)).softmax(-1)
    return prob.matmul(q)


def visibility_loss(vis_sd, xyz_pos, w_pos, bound, neg_u, n_freqs=10, alpha=10, spec=VIS_SPEC):
    """loss_utils.py:125-149.  neg_u: the U[0,1) draw of :138, (1, P, 3)."""
    dt = xyz_pos.dtype
    n = w_pos.numel()
    b = torch.as_tensor([float(x) for x in bound], dtype=dt)[None, None]
    xyz_neg = neg_u.to(dt) * 2 * b - b
    neg = nerf_forward(vis_sd, spec, embed(xyz_neg, n_freqs, alpha))[..., 0]
    pos = nerf_forward(vis_sd, spec, embed(xyz_pos.detach(), n_freqs, alpha))[..., 0]
    return -(F.logsigmoid(pos) * w_pos.detach()).sum() / n + -F.logsigmoid(-neg).sum() * 0.1 / n


def render_rays_full(prob, n_samples=128, noise=None, fm_noise=None, vis_u=None, xyz_freqs=10, dir_freqs=4, alpha=10,
                     use_ot=True, is_training=True):
    """rendering.py:19-579 at MoDA's default flags (dist_corresp, use_corresp, use_ot on; use_corr off), perturb = 0:
    ``prob`` from moda_b200.synth.make_full_problem.  Random inputs of the reference's in-call draws are arguments."""
    rays = prob["rays"]
    img_size, bound = prob["img_size"], [float(b) for b in prob["obj_bound"]]
    R = rays["rays_d"].shape[0]
    d = rays["rays_d"]
    dir_embedded = embed(d / d.norm(2, -1)[:, None], dir_freqs, alpha)
    z = sample_depths(rays["near"], rays["far"], n_samples)
    xyz_frame = rays["rays_o"].unsqueeze(1) + d.unsqueeze(1) * z.unsqueeze(2)
    S = n_samples
    bones_rst, skin_aux = prob["bones_rst"], prob["skin_aux"]
    rest_code = prob["rest_pose_code"][0:1]
    skin_sd = prob["nerf_skin"]
    bones_dfm = bone_transform(bones_rst, rays["bone_rts"])
    skin_bw = gauss_mlp_skinning(xyz_frame, xyz_freqs, alpha, bones_dfm, rays["time_embedded"][:, None], skin_sd, SKIN_SPEC,
                                 skin_aux)
    xyz, _ = neu_dbs(bones_rst, rays["bone_rts"], skin_bw, xyz_frame, backward=True)
    skin_fw = gauss_mlp_skinning(xyz, xyz_freqs, alpha, bones_rst, rest_code, skin_sd, SKIN_SPEC, skin_aux)
    xyz_cyc, _ = neu_dbs(bones_rst, rays["bone_rts"], skin_fw, xyz, backward=False)
    cyc = (xyz_frame - xyz_cyc).norm(2, -1)
    xyz_target, _ = neu_dbs(bones_rst, rays["bone_rts_target"], skin_fw, xyz, backward=False)
    dir_rep = dir_embedded[:, None].expand(R, S, dir_embedded.shape[-1])
    out = evaluate_mlp(prob["coarse"], COARSE_SPEC, xyz, xyz_freqs, alpha, dir_embedded=dir_rep, code=rays["env_code"],
                       chunk=4096)
    feat = evaluate_mlp(prob["nerf_feat"], FEAT_SPEC, xyz, xyz_freqs, alpha, chunk=4096)
    rgb, depth, sil, weights, vis = composite(out[..., :3], out[..., 3], z, d, prob["coarse"]["beta"], noise=noise)
    feat_rnd = (weights.unsqueeze(-1) * feat).sum(-2)
    res = {"img_coarse": rgb, "depth_rnd": depth, "sil_coarse": sil}
    # feature matching + reprojection (rendering.py:413-432)
    pts_exp = compute_pts_exp(weights, xyz)
    pts_pred = feat_match(prob["nerf_feat"], rays["feats_at_samp"], bound, prob["nerf_feat"]["beta"], xyz_freqs, alpha,
                          use_ot=use_ot, noise=fm_noise if is_training else None)
    res["pts_pred"], res["pts_exp"] = pts_pred, pts_exp
    res["feat_err"] = (pts_pred - pts_exp).norm(2, -1)[:, None]
    pp = pts_pred[:, None]
    skin_kp = gauss_mlp_skinning(pp, xyz_freqs, alpha, bones_rst, rest_code, skin_sd, SKIN_SPEC, skin_aux)
    kp, _ = neu_dbs(bones_rst, rays["bone_rts"], skin_kp, pp, backward=False)
    xy_reproj = project(kp, rays["rtk_vec"])[..., :2]
    res["proj_err"] = (rays["xys"][:, None] - xy_reproj).norm(2, -1) / img_size * 2
    xyz_target = project(xyz_target, rays["rtk_vec_target"])
    res["xyz_camera_vis"], res["xyz_canonical_vis"] = xyz_frame, xyz
    res["pts_exp_vis"], res["pts_pred_vis"] = pts_exp, pts_pred
    res["frame_cyc_dis"] = (cyc * weights.detach()).sum(-1)
    if is_training:
        res["vis_loss"] = visibility_loss(prob["nerf_vis"], xyz, vis, bound, vis_u, xyz_freqs, alpha)
    flo, flo_valid = vrender_flo(weights, xyz_target, rays["xys"], img_size)
    res["flo_coarse"], res["flo_valid"] = flo, flo_valid
    # per-ray losses (rendering.py:516-578)
    img_s, sil_s, vis_s = rays["img_at_samp"], rays["sil_at_samp"], rays["vis_at_samp"]
    img_loss = (rgb - img_s).pow(2).mean(-1)[..., None]
    wt = 1
    if is_training and sil_s.sum() > 0 and (1 - sil_s).sum() > 0:
        pos_wt = vis_s.sum() / sil_s[vis_s > 0].sum()
        neg_wt = vis_s.sum() / (1 - sil_s[vis_s > 0]).sum()
        wt = 0.5 * pos_wt * sil_s + 0.5 * neg_wt * (1 - sil_s)
    sil_loss = (sil[..., None] - sil_s).pow(2) * wt * vis_s
    flo_loss = (flo - rays["flo_at_samp"]).pow(2).sum(-1)
    cfd = rays["cfd_at_samp"]
    sel = (sil_s > 0) & (flo_valid == 1) & (cfd != 0)
    if sel.sum() > 0:
        cfd = cfd / cfd[sel].mean()
    res.update(img_at_samp=img_s, sil_at_samp=sil_s, vis_at_samp=vis_s, sil_at_samp_flo=sel, flo_at_samp=rays["flo_at_samp"])
    res["img_loss_samp"] = img_loss * sil_s
    res["sil_loss_samp"] = sil_loss
    res["flo_loss_samp"] = flo_loss[..., None] * cfd * sil_s
    res["frnd_loss_samp"] = (F.normalize(feat_rnd, 2, -1) - rays["feats_at_samp"]).pow(2).mean(-1) * sil_s[..., 0]
    return res


FULL_LOSS_KEYS = ("img_loss_samp", "sil_loss_samp", "flo_loss_samp", "feat_err", "proj_err", "frnd_loss_samp", "frame_cyc_dis")


def full_loss(res):
    """The scalar of oracle/make_golden2.py:full_loss."""
    loss = res["vis_loss"]
    for k in FULL_LOSS_KEYS:
        loss = loss + res[k].mean()
    return loss
