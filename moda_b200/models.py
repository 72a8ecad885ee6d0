"""Builds the ``models`` / ``embeddings`` dicts that ``render_rays`` consumes, the way
``nnutils/moda.py:271-329`` does, from a synthetic problem (``moda_b200.synth``) or from scratch."""
import torch
from torch import nn

from .nerf import Embedding, NeRF


def build_models(prob, device="cuda", requires_grad=True):
    """prob: dict from synth.make_problem.  Returns (models, embeddings, rays) on ``device``."""
    nb = prob["num_bones"]
    coarse = NeRF(in_channels_xyz=63, in_channels_dir=27 + 64, init_beta=0.1)
    skin = NeRF(in_channels_xyz=63 + 128, D=5, W=64, in_channels_dir=0, out_channels=nb, raw_feat=True,
                in_channels_code=128)
    coarse.load_state_dict(prob["coarse"])
    skin.load_state_dict(prob["nerf_skin"])
    rest = nn.Embedding(1, 128)
    rest.weight.data.copy_(prob["rest_pose_code"])
    coarse, skin, rest = coarse.to(device), skin.to(device), rest.to(device)
    bones_rst = prob["bones_rst"].clone().to(device).requires_grad_(requires_grad)
    skin_aux = prob["skin_aux"].clone().to(device).requires_grad_(requires_grad)
    models = {"coarse": coarse, "bones": bones_rst, "bones_rst": bones_rst, "skin_aux": skin_aux,
              "nerf_skin": skin, "rest_pose_code": rest}
    embeddings = {"xyz": Embedding(3, 10, alpha=10), "dir": Embedding(3, 4, alpha=10)}
    rays = {k: v.clone().to(device) for k, v in prob["rays"].items()}
    if requires_grad:
        for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
            rays[k].requires_grad_(True)
    return models, embeddings, rays


def build_full_models(prob, device="cuda", requires_grad=True):
    """``build_models`` + nerf_feat (5x128 -> 16 features, init_beta 1) and nerf_vis (5x64 -> 1) as nnutils/moda.py:344-348
    and 447-449 construct them; prob from synth.make_full_problem."""
    models, embeddings, rays = build_models(prob, device, requires_grad)
    feat = NeRF(in_channels_xyz=63, D=5, W=128, out_channels=16, in_channels_dir=0, raw_feat=True, init_beta=1.)
    vis = NeRF(in_channels_xyz=63, D=5, W=64, out_channels=1, in_channels_dir=0, raw_feat=True)
    feat.load_state_dict(prob["nerf_feat"])
    vis.load_state_dict(prob["nerf_vis"])
    models["nerf_feat"], models["nerf_vis"] = feat.to(device), vis.to(device)
    if requires_grad and "bone_rts_target" in rays:
        rays["bone_rts_target"].requires_grad_(True)
    return models, embeddings, rays


def parameters_of(models):
    """Every leaf tensor of ``models`` that receives a gradient from ``render_rays``."""
    ps = []
    for k in ("coarse", "nerf_skin", "rest_pose_code", "nerf_feat", "nerf_vis"):
        if k in models:
            ps += [p for p in models[k].parameters()]
    for k in ("bones_rst", "skin_aux"):
        if k in models and models[k].requires_grad:
            ps.append(models[k])
    return ps


def build_motion_models(prob, device="cuda", requires_grad=True):
    """models / embeddings / rays for ``synth.make_motion_problem`` (LBS: the core dict with rigid ``bone_rts``; flow
    fields: ``coarse`` + ``flowbw`` / ``flowfw`` as nnutils/moda.py:285-299 constructs them)."""
    from .nerf import SE3head, Transhead
    if prob["motion"] == "lbs":
        return build_models(prob, device, requires_grad)
    coarse = NeRF(in_channels_xyz=63, in_channels_dir=27 + 64, init_beta=0.1)
    coarse.load_state_dict(prob["coarse"])
    arch, oc = (Transhead, 3) if prob["motion"] == "trans" else (SE3head, 9)
    models = {"coarse": coarse.to(device)}
    for k in ("flowbw", "flowfw"):
        m = arch(in_channels_xyz=63 + 128, D=5, W=128, out_channels=oc, in_channels_dir=0, raw_feat=True)
        m.load_state_dict(prob[k])
        models[k] = m.to(device)
    embeddings = {"xyz": Embedding(3, 10, alpha=10), "dir": Embedding(3, 4, alpha=10)}
    rays = {k: v.clone().to(device) for k, v in prob["rays"].items()}
    if requires_grad:
        for k in ("time_embedded", "env_code", "rays_o", "rays_d"):
            rays[k].requires_grad_(True)
    return models, embeddings, rays
