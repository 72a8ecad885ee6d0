"""Skinning / warp functions with the reference's names and signatures (nnutils/geom_utils.py), CUDA-backed.

Only the dual-quaternion ("neudbs") motion model of MoDA is implemented; the LBS alternative
(geom_utils.py:304-348) is off by default in the reference (moda.py:72-80) and listed as "next" in
SURVEY.md section 8(f).
"""
import torch

from . import chain_tc, config, generic_tc, skin_tc, trunk_tc
from .ops import (BoneTransformFn, SkinWarpFn, SEG_DENSE, SEG_BCAST, SEG_PE)


def _split_segments(segs, cx):
    """Splits a column-segment list at column ``cx`` into (xyz part, dir part)."""
    xyz, dirs, k = [], [], 0
    for (kind, idx, width, aux, col) in segs:
        if k + width <= cx:
            xyz.append((kind, idx, width, aux, col))
        elif k >= cx:
            dirs.append((kind, idx, width, aux, col))
        else:
            if kind == SEG_PE:
                raise RuntimeError("a positional-encoding segment straddles in_channels_xyz")
            a = cx - k
            xyz.append((kind, idx, a, aux, col))
            dirs.append((kind, idx, width - a, aux, col + a))
        k += width
    return xyz, dirs


def evaluate_mlp(model, xyz_embedded, embed_xyz=None, dir_embedded=None, chunk=32 * 1024, xyz=None, code=None,
                 appearance_code=None, sigma_only=False, use_semantic=False, _pitched=False):
    """geom_utils.py:19-57 plus the output stage of the flow-field heads (Transhead / SE3head override forward() in the
    reference, nerf.py:200-237): a model with a ``post(raw, xyz)`` method gets it applied to the raw MLP output, with
    ``xyz`` the sample points (the ``xyz=`` argument, or the points themselves when they are embedded here)."""
    out = _evaluate_mlp(model, xyz_embedded, embed_xyz, dir_embedded, chunk, xyz, code, appearance_code, sigma_only,
                        use_semantic, _pitched)
    post = getattr(model, "post", None)
    if post is not None and not sigma_only:
        pts = xyz if xyz is not None else (xyz_embedded if embed_xyz is not None else None)
        out = post(out, pts)
    return out


def _evaluate_mlp(model, xyz_embedded, embed_xyz=None, dir_embedded=None, chunk=32 * 1024, xyz=None, code=None,
                  appearance_code=None, sigma_only=False, use_semantic=False, _pitched=False):
    """geom_utils.py:19-57.  xyz_embedded: (B,nbins,k) points (if ``embed_xyz`` is given) or features.

    The reference concatenates [PE | dir | code | appearance] per ray-chunk and calls the MLP; here the
    concatenation is virtual (column segments assembled inside the GEMM tile loaders), so ``chunk`` only
    exists for signature compatibility.  Returns (B, nbins, out).
    """
    Bn, nbins, k = xyz_embedded.shape
    M = Bn * nbins
    inputs, segs, win = [], [], None
    pts2 = xyz_embedded.reshape(M, k)
    inputs.append(pts2)
    if embed_xyz is not None and embed_xyz.N_freqs > 0:
        segs.append((SEG_PE, 0, embed_xyz.out_channels, k, 0))
        win = embed_xyz.window()
    else:
        segs.append((SEG_DENSE, 0, k, 1, 0))
    if dir_embedded is not None:
        d = dir_embedded
        if d.dim() == 3 and d.shape[1] == nbins:  # per-sample (the reference's repeat_interleave'd form)
            if d.stride(1) == 0 and nbins > 1:  # an expand()ed per-ray tensor: keep it per-ray
                inputs.append(d[:, 0])
                segs.append((SEG_BCAST, len(inputs) - 1, d.shape[-1], nbins, 0))
            else:
                inputs.append(d.reshape(M, d.shape[-1]))
                segs.append((SEG_DENSE, len(inputs) - 1, d.shape[-1], 1, 0))
        else:  # per-ray (B,c) or (B,1,c)
            inputs.append(d.reshape(Bn, d.shape[-1]))
            segs.append((SEG_BCAST, len(inputs) - 1, d.shape[-1], nbins, 0))
    for c in (code, appearance_code):
        if c is None:
            continue
        if c.dim() == 3:
            c = c.reshape(c.shape[0], c.shape[-1]) if c.shape[1] == 1 else c
        if c.dim() == 3:  # already per-sample
            inputs.append(c.reshape(M, c.shape[-1]))
            segs.append((SEG_DENSE, len(inputs) - 1, c.shape[-1], 1, 0))
        elif c.shape[0] != Bn:  # shared by every ray (code.repeat(B,1), geom_utils.py:38-39)
            inputs.append(c.reshape(1, c.shape[-1]))
            segs.append((SEG_BCAST, len(inputs) - 1, c.shape[-1], M, 0))
        else:
            inputs.append(c)
            segs.append((SEG_BCAST, len(inputs) - 1, c.shape[-1], nbins, 0))
    cx = model.in_channels_xyz
    # tensor-core fast path for the reference's nerf_coarse on [PE(xyz) | per-ray dir | per-ray env]
    if (config.precision == "fp16" and not sigma_only and len(segs) in (2, 3) and segs[0][0] == SEG_PE
            and all(sg[0] == SEG_BCAST and sg[3] == nbins for sg in segs[1:])
            and trunk_tc.supported(model, sum(sg[2] for sg in segs[1:])) and embed_xyz.N_freqs == 10 and k == 3):
        env = inputs[2] if len(segs) == 3 else None
        if config.fused:
            out = chain_tc.TrunkChainFn.apply(pts2, inputs[1], env, nbins, win, torch.is_grad_enabled(), *model.param_list())
        else:
            out = trunk_tc.TrunkTcFn.apply(pts2, inputs[1], env, nbins, win, *model.param_list())
        return out.reshape(Bn, nbins, 4)
    # tensor-core (split-precision) path for the reference's nerf_skin on [PE(xyz) | pose code]
    if (config.precision == "fp16" and not sigma_only and len(segs) == 2 and segs[0][0] == SEG_PE
            and segs[1][0] == SEG_BCAST and segs[1][3] in (nbins, M) and k == 3 and embed_xyz.N_freqs == 10
            and skin_tc.supported(model, segs[1][2])):
        if config.fused:
            out = chain_tc.SkinChainFn.apply(pts2, inputs[1], nbins, win, torch.is_grad_enabled(), *model.param_list())
        else:
            out = skin_tc.SkinMlpTcFn.apply(pts2, inputs[1], nbins, win, *model.param_list())
        out = out.reshape(Bn, nbins, 32)
        return out if _pitched else out[..., :model.out_channels]
    # density-only grid query (mesh extraction): the sigma program of the chain kernel, inference only
    if (config.precision == "fp16" and config.fused and sigma_only and len(segs) == 1 and segs[0][0] == SEG_PE and k == 3
            and embed_xyz.N_freqs == 10 and trunk_tc.supported(model, model.in_channels_dir)
            and not (torch.is_grad_enabled() and (pts2.requires_grad or any(p.requires_grad for p in model.parameters())))):
        return chain_tc.trunk_sigma(pts2, win, model.param_list()).reshape(Bn, nbins, 1)
    # tensor-core path for the auxiliary raw-feature MLPs on a plain PE(xyz) input (nerf_feat, nerf_vis)
    if (config.precision == "fp16" and not sigma_only and len(segs) == 1 and segs[0][0] == SEG_PE
            and generic_tc.supported(model, embed_xyz, k)):
        if config.fused and skin_tc.supported(model, 0):
            # a 5 x 64 raw-feature net without code columns (nerf_vis) is nerf_skin's architecture: same chain kernels
            out = chain_tc.SkinChainFn.apply(pts2, None, nbins, win, torch.is_grad_enabled(), *model.param_list())
            return out.reshape(Bn, nbins, 32)[..., :model.out_channels]
        if config.fused and config.feat_chain and chain_tc.feat_supported(model, embed_xyz, k):
            # nerf_feat (5 x 128): one chain kernel per pass
            out = chain_tc.FeatChainFn.apply(pts2, win, torch.is_grad_enabled(), *model.param_list())
            return out.reshape(Bn, nbins, 32)[..., :model.out_channels]
        skip = model.skips[0] if model.skips else None
        out = generic_tc.GenericTcFn.apply(pts2, win, torch.is_grad_enabled(), model.D, model.W, skip, *model.param_list())
        return out.reshape(Bn, nbins, 32)[..., :model.out_channels]
    xyz_segs, dir_segs = _split_segments(segs, cx)
    if len(xyz_segs) > 2 or len(dir_segs) > 2:
        raise NotImplementedError("more than two column segments per input group")
    out = model.run(M, inputs, xyz_segs, dir_segs, win, sigma_only)
    return out.reshape(Bn, nbins, out.shape[-1])


def bone_transform(bones_in, rts, neudbs=True, is_vec=False):
    """geom_utils.py:59-111.  neudbs: bones (...,B,10) moved by dual quaternions rts (...,B*8) -> (bs,B,10) (CUDA kernel).
    Otherwise (the LBS motion model, :87-107): rts are rigid transforms, (..,B,12) [R row-major | T] when ``is_vec`` else
    (..,B,3,4); centre' = R c + T, orient' = standardize(quat(R) (x) orient).  O(rays x bones) device tensor algebra."""
    if neudbs:
        return BoneTransformFn.apply(bones_in, rts)
    B = bones_in.shape[-2]
    bones = bones_in.reshape(-1, B, 10)
    if is_vec:
        rts = rts.reshape(-1, B, 12)
        Rmat, Tmat = rts[..., :9].reshape(-1, B, 3, 3), rts[..., 9:12]
    else:
        rts = rts.reshape(-1, B, 3, 4)
        Rmat, Tmat = rts[..., :3], rts[..., 3]
    center = (Rmat * bones[:, :, None, :3]).sum(-1) + Tmat
    orient = quaternion_multiply(matrix_to_quaternion(Rmat), bones[:, :, 3:7].expand(Rmat.shape[0], B, 4))
    scale = bones[:, :, 7:10].expand(Rmat.shape[0], B, 3)
    return torch.cat([center, orient, scale], -1)


def matrix_to_quaternion(matrix):
    """pytorch3d 0.6.1 ``matrix_to_quaternion`` (third_party/pytorch3d/.../rotation_conversions.py:101-152): of the four
    algebraically equal candidates (one per quaternion component used as the pivot) the one with the largest pivot is
    returned; gradients flow through that candidate only, and a non-positive pivot radicand has zero subgradient."""
    m = matrix.reshape(matrix.shape[:-2] + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = m.unbind(-1)
    rad = torch.stack([1 + m00 + m11 + m22, 1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22], -1)
    pos = rad > 0
    q_abs = torch.where(pos, torch.sqrt(torch.where(pos, rad, torch.ones_like(rad))), torch.zeros_like(rad))
    sq = q_abs * q_abs
    a, b, c = m21 - m12, m02 - m20, m10 - m01          # 4 r i, 4 r j, 4 r k
    d, e, f = m10 + m01, m02 + m20, m12 + m21          # 4 i j, 4 i k, 4 j k
    cand = torch.stack([torch.stack([sq[..., 0], a, b, c], -1), torch.stack([a, sq[..., 1], d, e], -1),
                        torch.stack([b, d, sq[..., 2], f], -1), torch.stack([c, e, f, sq[..., 3]], -1)], -2)
    cand = cand / (2.0 * q_abs.clamp_min(0.1))[..., None]
    best = q_abs.argmax(-1)
    return torch.gather(cand, -2, best[..., None, None].expand(best.shape + (1, 4))).squeeze(-2)


def quaternion_multiply(a, b):
    """pytorch3d ``quaternion_multiply`` (rotation_conversions.py:359-409): Hamilton product, real part made >= 0."""
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    ab = torch.stack((aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                      aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw), -1)
    return torch.where(ab[..., :1] < 0, -ab, ab)


def rts_invert(rts_in):
    """geom_utils.py:142-153: inverse of rigid transforms (...,3,4): [R^T | -R^T T]."""
    rts = rts_in.reshape(-1, 3, 4)
    Ri = rts[:, :, :3].transpose(1, 2)
    Ti = -(Ri * rts[:, None, :, 3]).sum(-1)
    return torch.cat([Ri, Ti[..., None]], -1).reshape(rts_in.shape)


def blend_skinning(rts, skin, pts):
    """geom_utils.py:304-348: linear blend skinning, x' = (sum_b w_b R_b) x + sum_b w_b T_b.  rts (bs,B,3,4), skin
    (bs,N,B), pts (bs,N,3).  The blended 3x4 transform of every point is ONE batched (N,B) x (B,12) product per ray (the
    reference materialises the (bs,N,B,3,3) broadcast product, 472 MB per 4096-ray chunk)."""
    B = rts.shape[-3]
    N = pts.shape[-2]
    pts = pts.reshape(-1, N, 3)
    rts = rts.reshape(-1, B, 12)
    G = torch.bmm(skin.reshape(-1, N, B), rts).reshape(-1, N, 3, 4)
    return (G[..., :3] * pts[:, :, None, :]).sum(-1) + G[..., 3]


def lbs(bones, rts_fw, skin, xyz_in, backward=True):
    """geom_utils.py:906-931: the LBS motion model (``--lbs``): rts_fw (bs,B*12) rigid transforms of the rest bones,
    backward = blend of their inverses.  Returns (xyz, bones_dfm)."""
    B = bones.shape[-2]
    N = xyz_in.shape[-2]
    bs = rts_fw.shape[0]
    xyz_in = xyz_in.reshape(-1, N, 3)
    v = rts_fw.reshape(-1, B, 12)
    rts = torch.cat([v[..., :9].reshape(bs, B, 3, 3), v[..., 9:12, None]], -1)
    bones_dfm = bone_transform(bones.reshape(-1, B, 10), rts, neudbs=False)
    xyz = blend_skinning(rts_invert(rts) if backward else rts, skin, xyz_in)
    return xyz, bones_dfm


def quaternion_to_matrix(q):
    """pytorch3d rotation_conversions.py:41-69 (host-side helper; not on the per-sample path)."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def vec_to_sim3(vec):
    """geom_utils.py:187-199 (per-bone helper; the per-sample use is fused into the skinning kernel)."""
    center = vec[..., :3]
    orient = torch.nn.functional.normalize(vec[..., 3:7], 2, -1)
    orient = quaternion_to_matrix(orient)
    scale = vec[..., 7:10].exp()
    return center, orient, scale


def skinning(bones, pts, dskin=None, skin_aux=None):
    """geom_utils.py:280-302: Gaussian-bone skinning weights, (bs,N,B)."""
    _, skin = SkinWarpFn.apply(pts, bones, None, skin_aux, dskin, None, False, False, False, True)
    return skin


def mlp_skinning(mlp, code, pts_embed, embed_xyz=None, _pitched=False):
    """geom_utils.py:219-229: delta skinning logits from nerf_skin.  With ``_pitched`` (internal callers that
    hand the result straight to the skinning kernels) the tensor-core path returns its zero-padded 32-column
    rows instead of a sliced view."""
    if mlp is None:
        return None
    return evaluate_mlp(mlp, pts_embed, embed_xyz=embed_xyz, code=code, chunk=8 * 1024, _pitched=_pitched)


def gauss_mlp_skinning(xyz, embedding_xyz, bones, pose_code, nerf_skin, skin_aux=None):
    """geom_utils.py:202-217.  The positional encoding is computed inside the first layer's tile loader
    instead of being materialised (``embedding_xyz(xyz)`` at :214)."""
    N_rays = xyz.shape[0]
    if pose_code.dim() == 2 and pose_code.shape[0] != N_rays:
        pose_code = pose_code.reshape(1, -1)
    dskin = mlp_skinning(nerf_skin, pose_code, xyz, embed_xyz=embedding_xyz, _pitched=True)
    return skinning(bones, xyz, dskin, skin_aux=skin_aux)


def _identity_bones(B, device):
    b = torch.zeros(B, 10, device=device)
    b[:, 3] = 1.0
    return b


def dqs_blend_skinning(dq, skin, pts):
    """geom_utils.py:495-517: blend per-bone dual quaternions with ``skin`` and transform ``pts``."""
    B = dq.shape[-2]
    N = pts.shape[-2]
    pts = pts.reshape(-1, N, 3)
    y, _ = SkinWarpFn.apply(pts, _identity_bones(B, pts.device), dq.reshape(-1, B, 8), None, None, skin,
                            False, False, True, False)
    return y


def neu_dbs(bones, rts_fw, skin, xyz_in, nerf_dis=None, embedding_xyz=None, code=None, backward=True):
    """geom_utils.py:372-456.  Returns (xyz, bones_dfm, xyz_dis | 0)."""
    B = bones.shape[-2]
    N = xyz_in.shape[-2]
    bones = bones.reshape(-1, B, 10)
    xyz_in = xyz_in.reshape(-1, N, 3)
    rts_fw = rts_fw.reshape(-1, B, 8)
    xyz_dis = 0
    if nerf_dis is not None and not backward:
        xyz_dis = evaluate_mlp(nerf_dis, xyz_in, embedding_xyz, code=code, chunk=xyz_in.shape[0])
        xyz_in = xyz_in + xyz_dis
    xyz, _ = SkinWarpFn.apply(xyz_in, _identity_bones(B, xyz_in.device), rts_fw, None, None, skin, False,
                              bool(backward), True, False)
    if nerf_dis is not None and backward:
        xyz_dis = evaluate_mlp(nerf_dis, xyz_in, embedding_xyz, code=code, chunk=xyz_in.shape[0])
        xyz = xyz - xyz_dis
    bones_dfm = bone_transform(bones, rts_fw, neudbs=True)
    return xyz, bones_dfm, xyz_dis


def warp_points(xyz, bones_rst, rts_fw, skin_aux, dskin, backward=True):
    """Fused gauss-skinning + DQ warp (what inference_deform does at rendering.py:303-341 through
    gauss_mlp_skinning + neu_dbs) without materialising the (N,S,B) weights."""
    B = bones_rst.shape[-2]
    y, _ = SkinWarpFn.apply(xyz, bones_rst.reshape(B, 10), rts_fw.reshape(-1, B, 8), skin_aux, dskin, None,
                            bool(backward), bool(backward), True, False)
    return y


# ------------------------------------------------------------------------------------------------ cameras, rays, flow
# Per-RAY (not per-sample) bookkeeping of the caller side and the small projective pieces of the
# flow / reprojection terms (rank 1).  These run on whatever device their inputs live on; the per-sample work they feed
# (warps, MLPs, compositing) goes through the CUDA kernels above.

def obj_to_cam(in_verts, Rmat, Tmat):
    """geom_utils.py:567-581: verts (...,N,3) -> R verts + T with per-batch R (...,3,3), T (...,3)."""
    verts = in_verts
    if verts.dim() == 2:
        verts = verts[None]
    verts = verts.reshape(-1, verts.shape[1], 3)
    Rt = Rmat.reshape(-1, 3, 3).permute(0, 2, 1)
    verts = verts.matmul(Rt) + Tmat.reshape(-1, 1, 3)
    return verts.reshape(in_verts.shape)


def K2mat(K):
    """geom_utils.py:596-610: (fx, fy, px, py) -> 3x3 intrinsics."""
    K = K.reshape(-1, 4)
    Kmat = torch.zeros(K.shape[0], 3, 3, device=K.device, dtype=K.dtype)
    Kmat[:, 0, 0], Kmat[:, 1, 1], Kmat[:, 0, 2], Kmat[:, 1, 2], Kmat[:, 2, 2] = K[:, 0], K[:, 1], K[:, 2], K[:, 3], 1
    return Kmat


def mat2K(Kmat):
    """geom_utils.py:612-627."""
    shape = Kmat.shape[:-2]
    Kmat = Kmat.reshape(-1, 3, 3)
    K = torch.stack([Kmat[:, 0, 0], Kmat[:, 1, 1], Kmat[:, 0, 2], Kmat[:, 1, 2]], -1)
    return K.reshape(shape + (4,))


def K2inv(K):
    """geom_utils.py:638-652: inverse intrinsics matrix from (fx, fy, px, py)."""
    K = K.reshape(-1, 4)
    Kmat = torch.zeros(K.shape[0], 3, 3, device=K.device, dtype=K.dtype)
    Kmat[:, 0, 0], Kmat[:, 1, 1] = 1.0 / K[:, 0], 1.0 / K[:, 1]
    Kmat[:, 0, 2], Kmat[:, 1, 2], Kmat[:, 2, 2] = -K[:, 2] / K[:, 0], -K[:, 3] / K[:, 1], 1
    return Kmat


def Kmatinv(Kmat):
    """geom_utils.py:629-636."""
    return K2inv(mat2K(Kmat)).reshape(Kmat.shape)


def pinhole_cam(in_verts, K):
    """geom_utils.py:654-672: verts (...,N,3), K (...,4) -> (x / z, y / z, z) with z guarded by 1e-6."""
    verts = in_verts.reshape(-1, in_verts.shape[1], 3)
    Kt = K2mat(K.reshape(-1, 4)).permute(0, 2, 1)
    verts = verts.matmul(Kt)
    z = verts[:, :, 2:3]
    xy = verts[:, :, :2] / (1e-6 + z)
    return torch.cat([xy, z], -1).reshape(in_verts.shape)


def vrender_flo(weights_coarse, xyz_coarse_target, xys, img_size):
    """geom_utils.py:1704-1743: expected 2-D motion of a ray's samples projected into the target view.
    weights (...,S), xyz_target (...,S,3) already projected (x, y, depth), xys (...,2) -> flo (...,2), valid (...,1)."""
    wshape = weights_coarse.shape
    xyz_t = xyz_coarse_target.reshape(wshape + (3,))
    xy_t = xyz_t[..., :2]
    invalid = torch.logical_or(xyz_t[..., -1] < 1e-5, xy_t.norm(2, -1).abs() > 2 * img_size)
    w = torch.where(invalid, torch.zeros_like(weights_coarse), weights_coarse)
    xy_t = torch.where(invalid[..., None], torch.zeros_like(xy_t), xy_t)
    w = w / (1e-9 + w.sum(-1)[..., None])
    flo = (w[..., None] * (xy_t - xys.reshape(wshape[:-1] + (1, 2)))).sum(-2)
    flo = flo / img_size * 2
    valid = (invalid.sum(-1) == 0).float()[..., None]
    return flo, valid


class FlowRenderFn(torch.autograd.Function):
    """``vrender_flo(weights, pinhole_cam(obj_to_cam(xyz, R, T), K), xys, img_size)`` as one kernel each way
    (csrc/flow.cu): xyz (N,S,3) root-frame samples of a paired frame, R (N,9), T (N,3), K (N,4), weights (N,S), xys (N,2)."""

    @staticmethod
    def forward(ctx, xyz, R, T, K, weights, xys, img_size):
        from ._lib import call, ptr, stream, f32
        xyz, R, T, K, weights, xys = (f32(t).contiguous() for t in (xyz, R, T, K, weights, xys))
        N, S = weights.shape
        flo = torch.empty(N, 2, device=xyz.device, dtype=torch.float32)
        valid = torch.empty(N, 1, device=xyz.device, dtype=torch.float32)
        call("moda_flow_render_fwd", ptr(xyz), ptr(R), ptr(T), ptr(K), ptr(weights), ptr(xys), N, S, float(img_size), ptr(flo),
             ptr(valid), stream())
        ctx.save_for_backward(xyz, R, T, K, weights, xys)
        ctx.img_size = float(img_size)
        ctx.mark_non_differentiable(valid)
        return flo, valid

    @staticmethod
    def backward(ctx, gflo, _gvalid):
        from ._lib import call, ptr, stream, f32
        xyz, R, T, K, weights, xys = ctx.saved_tensors
        N, S = weights.shape
        new = lambda *shape: torch.empty(*shape, device=xyz.device, dtype=torch.float32)
        gxyz, gR, gT, gK, gw = new(N, S, 3), new(N, 9), new(N, 3), new(N, 4), new(N, S)
        call("moda_flow_render_bwd", ptr(xyz), ptr(R), ptr(T), ptr(K), ptr(weights), ptr(xys), ptr(f32(gflo).contiguous()), N, S,
             ctx.img_size, ptr(gxyz), ptr(gR), ptr(gT), ptr(gK), ptr(gw), stream())
        return gxyz, gR, gT, gK, gw, None, None


def project_render_flo(weights_coarse, xyz_root, rtk_vec, xys, img_size, N_rays):
    """rendering.py:434-459 + 480-499 for one paired frame: project the warped samples (N,S,3) with the frame's camera
    (rtk_vec (N,21) = R | T | K^-1) and render their expected 2-D motion.  On CUDA tensors this is one fused kernel each
    way; the per-ray camera algebra (K from K^-1) stays as tensor ops so that its gradient reaches rtk_vec."""
    Rmat = rtk_vec[:, 0:9].reshape(N_rays, 1, 3, 3)
    Tmat = rtk_vec[:, 9:12].reshape(N_rays, 1, 3)
    K = mat2K(Kmatinv(rtk_vec[:, 12:21].reshape(N_rays, 1, 3, 3)))
    S = weights_coarse.shape[-1]
    if xyz_root.is_cuda and weights_coarse.dim() == 2 and weights_coarse.shape[0] == N_rays:
        return FlowRenderFn.apply(xyz_root.reshape(N_rays, S, 3), Rmat.reshape(N_rays, 9), Tmat.reshape(N_rays, 3),
                                  K.reshape(N_rays, 4), weights_coarse, xys.reshape(N_rays, 2), img_size)
    return vrender_flo(weights_coarse, pinhole_cam(obj_to_cam(xyz_root, Rmat, Tmat), K), xys, img_size)


def diff_flo(pts_target, xys, img_size):
    """geom_utils.py:1745-1757."""
    return (pts_target.reshape(xys.shape) - xys) / img_size * 2


def raycast(xys, Rmat, Tmat, Kinv, near_far):
    """geom_utils.py:746-794: pixel coordinates -> rays in the root frame.  Returns the reference's ``rays`` dict
    (rays_o, rays_d, near, far, rtk_vec, xys, nsample, bs), every per-ray tensor (bs, nsample, ·)."""
    Rmat = Rmat.reshape(-1, 3, 3).float()
    Tmat = Tmat.reshape(-1, 1, 3).float()
    Kinv = Kinv.reshape(-1, 3, 3).float()
    xys = xys.float()
    bs, nsample, _ = xys.shape
    xy1s = torch.cat([xys, torch.ones_like(xys[:, :, :1])], 2)
    xyz3d = xy1s.matmul(Kinv.permute(0, 2, 1))
    ray_directions = xyz3d.matmul(Rmat)
    ray_origins = -Tmat.matmul(Rmat)
    if near_far is not None:
        znear = torch.ones(bs, nsample, 1, device=xys.device) * near_far[:, 0, None, None]
        zfar = torch.ones(bs, nsample, 1, device=xys.device) * near_far[:, 1, None, None]
    else:
        lbound, ubound = -1.5, 1.5
        znear = (Tmat[:, :, -1:].repeat(1, nsample, 1) + lbound).clamp_min(1e-5)
        zfar = Tmat[:, :, -1:].repeat(1, nsample, 1) + ubound
    rtk_vec = torch.cat([Rmat.reshape(-1, 1, 9), Tmat.reshape(-1, 1, 3), Kinv.reshape(-1, 1, 9)], -1)
    return {"rays_o": ray_origins.repeat(1, nsample, 1), "rays_d": ray_directions, "near": znear, "far": zfar,
            "rtk_vec": rtk_vec.repeat(1, nsample, 1), "xys": xys, "nsample": nsample, "bs": bs}


def sample_xy(img_size, bs, nsample, device, return_all=False, lineid=None):
    """geom_utils.py:796-827: pixel samples (rand_inds (bs,ns) long, xys (bs,ns,2)); same draws (torch.multinomial)."""
    ax = torch.arange(img_size, device=device, dtype=torch.float32)
    gy, gx = torch.meshgrid(ax, ax, indexing="ij")
    xygrid = torch.stack([gx, gy], -1).reshape(1, -1, 2)
    if return_all:
        xys = xygrid.repeat(bs, 1, 1)
        nsample = xys.shape[1]
        rand_inds = torch.arange(nsample, dtype=torch.float32)[None].repeat(bs, 1)
    else:
        if lineid is None:
            probs = torch.ones(img_size ** 2, device=device)
            rand_inds = torch.multinomial(probs, bs * nsample, replacement=False).view(bs, nsample)
            xys = xygrid[0][rand_inds]
        else:
            probs = torch.ones(img_size, device=device)
            rand_inds = torch.multinomial(probs, bs * nsample, replacement=True).view(bs, nsample)
            xys = xygrid[0][rand_inds].clone()
            xys[..., 1] = xys[..., 1] + lineid[:, None]
    return rand_inds.long(), xys


def chunk_rays(rays, start, delta):
    """geom_utils.py:829-838."""
    return {k: v.reshape(-1, v.shape[-1])[start:start + delta] for k, v in rays.items() if torch.is_tensor(v)}


# ------------------------------------------------------------------------------------------------ rest-pose correction
def correct_bones(model, bones_rst, inverse=False, neudbs=True):
    """geom_utils.py:933-949: rest bones moved by the rest-pose transform Jb* = rts_head(rest_pose_code)."""
    if not neudbs:
        raise NotImplementedError("only the dual-quaternion (neudbs) motion model is implemented")
    from . import dual_quat as DQ
    dev = bones_rst.device
    rest_pose_code = model.rest_pose_code(torch.zeros(1, dtype=torch.long, device=dev))
    bone_rts_rst = model.nerf_body_rts[1](rest_pose_code)[0]
    if inverse:
        shape = bone_rts_rst.shape
        bone_rts_rst = DQ.dq_inverse(bone_rts_rst.reshape(-1, bones_rst.shape[-2], 8)).reshape(shape)
    bones_rst = bone_transform(bones_rst, bone_rts_rst, neudbs, is_vec=True)[0]
    return bones_rst, bone_rts_rst


def correct_rest_pose(opts, bone_rts_fw, bone_rts_rst, neudbs):
    """geom_utils.py:951-972: delta transform Jb (Jb*)^-1 in dual-quaternion algebra."""
    if not neudbs:
        raise NotImplementedError("only the dual-quaternion (neudbs) motion model is implemented")
    from . import dual_quat as DQ
    rts_shape = bone_rts_fw.shape
    B = opts.num_bones
    inv = DQ.dq_inverse(bone_rts_rst.reshape(-1, B, 8))
    fw = bone_rts_fw.reshape(-1, B, 8)
    inv = inv.expand(fw.shape[0], B, 8).contiguous()
    return DQ.dq_mul(inv, fw.contiguous()).reshape(rts_shape)


# ------------------------------------------------------------------------------------------------ mesh-extraction warps
def _frame_inputs(opts, model, npts, embedid):
    """Shared by warp_bw / warp_fw (geom_utils.py:974-1073): per-point frame id -> pose code, corrected bones and the
    delta bone transforms of that frame."""
    query_time = torch.ones(npts, 1, device=model.device).long() * embedid
    bone_rts_fw = model.nerf_body_rts(query_time)
    bones_rst, bone_rts_rst = correct_bones(model, model.bones, neudbs=opts.neudbs)
    bone_rts_fw = correct_rest_pose(opts, bone_rts_fw, bone_rts_rst, opts.neudbs)
    return query_time, bones_rst, bone_rts_fw


def warp_bw(opts, model, rt_dict, query_xyz_chunk, embedid):
    """geom_utils.py:974-1027 (articulated branch): view-space points of frame ``embedid`` -> canonical space."""
    if getattr(opts, "flowbw", False) or getattr(opts, "lbs", False):
        raise NotImplementedError("warp_bw: only the neudbs motion model is implemented")
    n = query_xyz_chunk.shape[0]
    query_time, bones_rst, bone_rts_fw = _frame_inputs(opts, model, n, embedid)
    pts = query_xyz_chunk[:, None]
    nerf_skin = model.nerf_skin if opts.nerf_skin else None
    time_embedded = model.pose_code(query_time)
    nerf_dis = model.nerf_dis if getattr(opts, "nerf_dis", False) else None
    if nerf_dis is None:   # fused weights + warp
        dskin = mlp_skinning(nerf_skin, time_embedded, pts, embed_xyz=model.embedding_xyz, _pitched=True)
        out = warp_points(pts, bones_rst, bone_rts_fw, model.skin_aux, dskin, backward=True)
        bones_dfm = bone_transform(bones_rst, bone_rts_fw, opts.neudbs, is_vec=True)
    else:                  # the residual field is evaluated at the un-warped points: two-step path as in the reference
        bones_dfm = bone_transform(bones_rst, bone_rts_fw, opts.neudbs, is_vec=True)
        skin = gauss_mlp_skinning(pts, model.embedding_xyz, bones_dfm, time_embedded, nerf_skin, skin_aux=model.skin_aux)
        out, bones_dfm, _ = neu_dbs(bones_rst, bone_rts_fw, skin, pts, nerf_dis, model.embedding_xyz, time_embedded)
    rt_dict["bones"] = bones_dfm
    return out[:, 0], rt_dict


def warp_fw(opts, model, rt_dict, vertices, embedid):
    """geom_utils.py:1029-1073 (articulated branch): canonical mesh vertices -> frame ``embedid``.  Returns a numpy
    array like the reference (the mesh is assembled on the host)."""
    if getattr(opts, "flowbw", False) or getattr(opts, "lbs", False):
        raise NotImplementedError("warp_fw: only the neudbs motion model is implemented")
    pts_can = torch.as_tensor(vertices, dtype=torch.float32).to(model.device)[:, None]
    n = pts_can.shape[0]
    _, bones_rst, bone_rts_fw = _frame_inputs(opts, model, n, embedid)
    nerf_skin = model.nerf_skin if opts.nerf_skin else None
    rest_pose_code = model.rest_pose_code(torch.zeros(1, dtype=torch.long, device=bones_rst.device))
    nerf_dis = model.nerf_dis if getattr(opts, "nerf_dis", False) else None
    if nerf_dis is None:
        dskin = mlp_skinning(nerf_skin, rest_pose_code, pts_can, embed_xyz=model.embedding_xyz, _pitched=True)
        pts_dfm = warp_points(pts_can, bones_rst, bone_rts_fw, model.skin_aux, dskin, backward=False)
        bones_dfm = bone_transform(bones_rst, bone_rts_fw, opts.neudbs, is_vec=True)
    else:   # weights from the undisplaced points, blend applied to the displaced ones (geom_utils.py:423-429)
        skin = gauss_mlp_skinning(pts_can, model.embedding_xyz, bones_rst, rest_pose_code, nerf_skin, skin_aux=model.skin_aux)
        pts_dfm, bones_dfm, _ = neu_dbs(bones_rst, bone_rts_fw, skin, pts_can, nerf_dis, model.embedding_xyz, rest_pose_code,
                                        backward=False)
    rt_dict["bones"] = bones_dfm
    return pts_dfm[:, 0].detach().cpu().numpy(), rt_dict
