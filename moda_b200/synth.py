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
