"""Skinning / warp functions with the reference's names and signatures (nnutils/geom_utils.py), CUDA-backed.

Only the dual-quaternion ("neudbs") motion model of MoDA is implemented; the LBS alternative
(geom_utils.py:304-348) is off by default in the reference (moda.py:72-80) and listed as "next" in
SURVEY.md section 8(f).
"""
import torch

from . import chain_tc, config, skin_tc, trunk_tc
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
    xyz_segs, dir_segs = _split_segments(segs, cx)
    if len(xyz_segs) > 2 or len(dir_segs) > 2:
        raise NotImplementedError("more than two column segments per input group")
    out = model.run(M, inputs, xyz_segs, dir_segs, win, sigma_only)
    return out.reshape(Bn, nbins, out.shape[-1])


def bone_transform(bones_in, rts, neudbs=True, is_vec=False):
    """geom_utils.py:59-111 (dual-quaternion branch): bones (...,B,10) moved by rts (...,B*8) -> (bs,B,10)."""
    if not neudbs:
        raise NotImplementedError("only the dual-quaternion (neudbs) motion model is implemented")
    return BoneTransformFn.apply(bones_in, rts)


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
