"""Seeded synthetic inputs for the articulated volume-rendering path.

Everything is produced as plain CPU fp32 tensors (state dicts + a flat ``rays`` dict) so that the same
numbers can be loaded into this package's CUDA-backed modules, into the CPU oracle
(``oracle/restated.py``) and into the real reference (``oracle/ref_loader.py``).  The recipe mirrors
what ``nnutils/moda.py:271-329`` constructs (8x256 ``nerf_coarse`` with 63+27+64 inputs, 5x64
``nerf_skin`` with 63+128 inputs and one logit per bone, 25 Gaussian bones, a 128-d rest pose code)
and what ``moda.py:1281-1327`` / ``geom_utils.py:785-794`` put into ``rays``.
"""
import math

import torch

NUM_BONES = 25
T_EMBED = 128
ENV_DIM = 64
PE_XYZ = 63
PE_DIR = 27
F32 = torch.float32


def _linear_init(gen, out_f, in_f):
    """nn.Linear's default init: U(-1/sqrt(in), 1/sqrt(in)) for weight and bias."""
    bound = 1.0 / math.sqrt(in_f)
    w = (torch.rand(out_f, in_f, generator=gen, dtype=F32) * 2 - 1) * bound
    b = (torch.rand(out_f, generator=gen, dtype=F32) * 2 - 1) * bound
    return w, b


def nerf_state(gen, D=8, W=256, in_channels_xyz=63, in_channels_dir=27, out_channels=3,
               skips=(4,), init_beta=0.1):
    """State dict with the reference's parameter names (``nnutils/nerf.py:107-136``)."""
    sd = {}
    for i in range(D):
        if i == 0:
            k = in_channels_xyz
        elif i in skips:
            k = W + in_channels_xyz
        else:
            k = W
        w, b = _linear_init(gen, W, k)
        sd["xyz_encoding_%d.0.weight" % (i + 1)] = w
        sd["xyz_encoding_%d.0.bias" % (i + 1)] = b
    sd["xyz_encoding_final.weight"], sd["xyz_encoding_final.bias"] = _linear_init(gen, W, W)
    sd["dir_encoding.0.weight"], sd["dir_encoding.0.bias"] = _linear_init(gen, W // 2, W + in_channels_dir)
    sd["sigma.weight"], sd["sigma.bias"] = _linear_init(gen, 1, W)
    sd["rgb.0.weight"], sd["rgb.0.bias"] = _linear_init(gen, out_channels, W // 2)
    sd["beta"] = torch.tensor([init_beta], dtype=torch.float32)
    return sd


def q_mul_np(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack((aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw), -1)


def make_problem(n_rays, seed=0, num_bones=NUM_BONES, semi_transparent=True):
    """Returns ``dict(coarse=sd, nerf_skin=sd, bones_rst, skin_aux, rest_pose_code, rays={...})``.

    ``semi_transparent`` rescales the sigma head (weight*6, bias=-0.35) so that rays neither saturate
    nor vanish (sil mean ~0.36 instead of 0.997), which keeps every gradient path alive.
    """
    gen = torch.Generator().manual_seed(seed)
    coarse = nerf_state(gen, D=8, W=256, in_channels_xyz=PE_XYZ, in_channels_dir=PE_DIR + ENV_DIM,
                        out_channels=3, init_beta=0.1)
    if semi_transparent:
        coarse["sigma.weight"] = coarse["sigma.weight"] * 6.0
        coarse["sigma.bias"] = torch.full_like(coarse["sigma.bias"], -0.35)
    skin = nerf_state(gen, D=5, W=64, in_channels_xyz=PE_XYZ + T_EMBED, in_channels_dir=0,
                      out_channels=num_bones, init_beta=0.01)
    rest_pose_code = torch.randn(1, T_EMBED, generator=gen, dtype=F32)  # nn.Embedding default init N(0,1)

    center = 0.1 * torch.randn(num_bones, 3, generator=gen, dtype=F32)
    orient = torch.tensor([1.0, 0, 0, 0], dtype=F32) + 0.1 * torch.randn(num_bones, 4, generator=gen, dtype=F32)
    orient = orient / orient.norm(dim=-1, keepdim=True)
    lscale = 0.1 * torch.randn(num_bones, 3, generator=gen, dtype=F32)
    bones_rst = torch.cat([center, orient, lscale], -1)
    skin_aux = torch.tensor([0.0, 10.0], dtype=F32)

    N = n_rays
    rays_o = 0.05 * torch.randn(N, 3, generator=gen, dtype=F32) - torch.tensor([0.0, 0.0, 0.3], dtype=F32)
    rays_d = torch.randn(N, 3, generator=gen, dtype=F32)
    rays_d[:, 2] = rays_d[:, 2].abs() + 1.0
    near = torch.full((N, 1), 0.1, dtype=F32)
    far = torch.full((N, 1), 0.5, dtype=F32)
    xys = torch.rand(N, 2, generator=gen, dtype=F32) * 512
    time_embedded = 0.1 * torch.randn(N, T_EMBED, generator=gen, dtype=F32)
    env_code = 0.1 * torch.randn(N, ENV_DIM, generator=gen, dtype=F32)
    r = torch.tensor([1.0, 0, 0, 0], dtype=F32) + 0.2 * torch.randn(N, num_bones, 4, generator=gen, dtype=F32)
    r = r / r.norm(dim=-1, keepdim=True)
    t = 0.05 * torch.randn(N, num_bones, 3, generator=gen, dtype=F32)
    tq = torch.cat([torch.zeros(N, num_bones, 1, dtype=F32), t], -1)
    d = 0.5 * q_mul_np(tq, r)
    bone_rts = torch.cat([r, d], -1).reshape(N, num_bones * 8)
    rays = dict(rays_o=rays_o, rays_d=rays_d, near=near, far=far, xys=xys,
                time_embedded=time_embedded, env_code=env_code, bone_rts=bone_rts)
    return dict(coarse=coarse, nerf_skin=skin, rest_pose_code=rest_pose_code, bones_rst=bones_rst,
                skin_aux=skin_aux, rays=rays, num_bones=num_bones)


def make_skin_problem(n_rays, n_samples=128, seed=0, num_bones=NUM_BONES):
    """Inputs of the DQ-skinning microbench (BASELINE config 4): points along synthetic rays."""
    p = make_problem(n_rays, seed=seed, num_bones=num_bones)
    rays = p["rays"]
    s = torch.linspace(0, 1, n_samples, dtype=F32)
    z = rays["near"] * (1 - s) + rays["far"] * s
    xyz = rays["rays_o"][:, None] + rays["rays_d"][:, None] * z[..., None]
    return dict(xyz=xyz, bones_rst=p["bones_rst"], skin_aux=p["skin_aux"], bone_rts=rays["bone_rts"],
                num_bones=num_bones)


def default_opts():
    """The ``opts`` fields the rendering path reads (SURVEY.md section 5, config/flags row)."""
    import types
    return types.SimpleNamespace(neudbs=True, lbs=False, dist_corresp=False, use_corresp=False,
                                 use_corr=False, use_ot=False, symm_shape=False, scale_rgb=1.3,
                                 rgb_filter=False, s3im_loss=False)


# ------------------------------------------------------------------------------------------------ round 2
# the full default-flag step of MoDA: nerf_feat / nerf_vis, a paired target frame with its
# own camera and bone transforms, per-ray observations.  Same conventions as above: plain CPU fp32 tensors.
IMG_SIZE = 512
NUM_FEAT = 16
OBJ_BOUND = (0.18, 0.2, 0.22)


def _small_rotation(gen, n, scale):
    """(n,3,3) rotation matrices from a small random quaternion perturbation of the identity."""
    q = torch.tensor([1.0, 0, 0, 0], dtype=F32) + scale * torch.randn(n, 4, generator=gen, dtype=F32)
    q = q / q.norm(dim=-1, keepdim=True)
    r, i, j, k = q.unbind(-1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(n, 3, 3)


def _random_dq(gen, n, num_bones, rs=0.2, ts=0.05):
    r = torch.tensor([1.0, 0, 0, 0], dtype=F32) + rs * torch.randn(n, num_bones, 4, generator=gen, dtype=F32)
    r = r / r.norm(dim=-1, keepdim=True)
    t = ts * torch.randn(n, num_bones, 3, generator=gen, dtype=F32)
    tq = torch.cat([torch.zeros(n, num_bones, 1, dtype=F32), t], -1)
    return torch.cat([r, 0.5 * q_mul_np(tq, r)], -1).reshape(n, num_bones * 8)


def make_full_problem(n_rays, seed=0, num_bones=NUM_BONES):
    """``make_problem`` + everything the default-flag training step reads (nnutils/moda.py:344-348, 447-449,
    1215-1327; geom_utils.py:746-794): ``nerf_feat`` (5x128 -> 16), ``nerf_vis`` (5x64 -> 1), rays cast from two
    cameras (the first half of the rays belongs to frame 0 and targets frame 1, the second half the reverse),
    ``rtk_vec`` / ``rtk_vec_target`` (N,21) = [R | T | Kinv], ``bone_rts_target`` and the per-ray observations."""
    p = make_problem(n_rays, seed=seed, num_bones=num_bones)
    gen = torch.Generator().manual_seed(seed + 7919)
    p["nerf_feat"] = nerf_state(gen, D=5, W=128, in_channels_xyz=PE_XYZ, in_channels_dir=0, out_channels=NUM_FEAT,
                                init_beta=1.0)
    p["nerf_vis"] = nerf_state(gen, D=5, W=64, in_channels_xyz=PE_XYZ, in_channels_dir=0, out_channels=1,
                               init_beta=0.01)
    # keep the visibility logits off the 0.5 decision threshold of rendering.py:215 (a random-init net outputs ~0)
    p["nerf_vis"]["rgb.0.weight"] = p["nerf_vis"]["rgb.0.weight"] * 8.0
    N = n_rays
    h = N // 2
    Rm = _small_rotation(gen, 2, 0.05)
    Tm = torch.tensor([[0.0, 0.0, 0.3]], dtype=F32) + 0.02 * torch.randn(2, 3, generator=gen, dtype=F32)
    K = torch.tensor([[600.0, 600.0, 256.0, 256.0], [620.0, 610.0, 250.0, 260.0]], dtype=F32)
    Kinv = torch.zeros(2, 3, 3, dtype=F32)
    Kinv[:, 0, 0] = 1 / K[:, 0]
    Kinv[:, 1, 1] = 1 / K[:, 1]
    Kinv[:, 0, 2] = -K[:, 2] / K[:, 0]
    Kinv[:, 1, 2] = -K[:, 3] / K[:, 1]
    Kinv[:, 2, 2] = 1
    frame = torch.cat([torch.zeros(h, dtype=torch.long), torch.ones(N - h, dtype=torch.long)])
    xys = torch.rand(N, 2, generator=gen, dtype=F32) * (IMG_SIZE - 1)
    xy1 = torch.cat([xys, torch.ones(N, 1, dtype=F32)], -1)
    xyz3d = torch.einsum("nj,nij->ni", xy1, Kinv[frame])            # xy1s.matmul(Kinv^T), geom_utils.py:764
    rays_d = torch.einsum("ni,nij->nj", xyz3d, Rm[frame])           # .matmul(Rmat), :765
    rays_o = -torch.einsum("ni,nij->nj", Tm[frame], Rm[frame])      # :766
    rtk = torch.cat([Rm.reshape(2, 9), Tm, Kinv.reshape(2, 9)], -1)
    rays = p["rays"]
    rays["rays_o"], rays["rays_d"], rays["xys"] = rays_o, rays_d, xys
    rays["rtk_vec"] = rtk[frame]
    rays["rtk_vec_target"] = rtk[1 - frame]
    rays["bone_rts_target"] = _random_dq(gen, N, num_bones)
    rays["feats_at_samp"] = torch.randn(N, NUM_FEAT, generator=gen, dtype=F32)
    rays["img_at_samp"] = torch.rand(N, 3, generator=gen, dtype=F32)
    rays["sil_at_samp"] = (torch.rand(N, 1, generator=gen, dtype=F32) < 0.6).float()
    rays["vis_at_samp"] = (torch.rand(N, 1, generator=gen, dtype=F32) < 0.9).float()
    rays["flo_at_samp"] = 0.05 * torch.randn(N, 2, generator=gen, dtype=F32)
    cfd = torch.rand(N, 1, generator=gen, dtype=F32)
    rays["cfd_at_samp"] = torch.where(cfd < 0.2, torch.zeros_like(cfd), cfd)
    p["obj_bound"] = torch.tensor(OBJ_BOUND, dtype=F32)
    p["img_size"] = IMG_SIZE
    return p


def full_opts(**kw):
    """``default_opts`` with the reference's default flag values for the full step (nnutils/moda.py:149-170)."""
    o = default_opts()
    o.dist_corresp, o.use_corresp, o.use_ot, o.use_corr = True, True, True, False
    for k, v in kw.items():
        setattr(o, k, v)
    return o


# ------------------------------------------------------------------------------------------------ motion models
def make_motion_problem(n_rays, kind, seed=0, num_bones=NUM_BONES):
    """The alternative motion models of SURVEY.md 8(f) rank 5 on the core problem.  kind = "lbs": ``rays['bone_rts']``
    holds rigid transforms (N, B*12) = [R row-major | T] per bone (moda.py:301-311, opts.lbs); kind = "trans" / "se3": no
    bones, free-form flow fields ``flowbw`` / ``flowfw`` (5x128 MLPs on [PE(xyz) | time code], 3 or 9 outputs:
    Transhead / SE3head, moda.py:285-299)."""
    p = make_problem(n_rays, seed=seed, num_bones=num_bones)
    gen = torch.Generator().manual_seed(seed + 1000)
    p["motion"] = kind
    if kind == "lbs":
        R = _small_rotation(gen, n_rays * num_bones, 0.2).reshape(n_rays, num_bones, 9)
        T = 0.05 * torch.randn(n_rays, num_bones, 3, generator=gen, dtype=F32)
        p["rays"]["bone_rts"] = torch.cat([R, T], -1).reshape(n_rays, num_bones * 12)
        # softer Gaussians (skin_aux[0] = log scale of the sharpness, geom_utils.py:245-262): with the core problem's
        # value the far samples get an exactly one-hot weight, their LBS cycle x -> R^-1 -> R returns x up to rounding,
        # and d|x - x_cyc| there is a unit vector in the direction of that rounding noise -- in the reference as much as
        # here, so no two implementations agree on it.  With several bones blending, |x - x_cyc| is far from zero.
        p["skin_aux"] = torch.tensor([-4.0, 10.0], dtype=F32)
        return p
    if kind not in ("trans", "se3"):
        raise ValueError(kind)
    oc = 3 if kind == "trans" else 9
    for k in ("flowbw", "flowfw"):
        sd = nerf_state(gen, D=5, W=128, in_channels_xyz=PE_XYZ + T_EMBED, in_channels_dir=0, out_channels=oc, init_beta=0.01)
        # default-init heads give flows of ~1e-3: scale the last layer so that the warp visibly moves the samples
        sd["rgb.0.weight"] = sd["rgb.0.weight"] * 8.0
        p[k] = sd
    for k in ("bones_rst", "skin_aux", "nerf_skin", "rest_pose_code"):
        p.pop(k)
    p["rays"].pop("bone_rts")
    return p
