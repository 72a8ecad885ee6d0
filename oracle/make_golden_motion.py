"""TEST INFRASTRUCTURE ONLY -- fixtures for the alternative motion models, produced by
executing the REAL reference.

Run in the build container (needs /root/reference):  python oracle/make_golden_motion.py
Writes (tests/golden/):
  render_lbs_n16_fp32.npz     render_rays with opts.lbs (rigid transforms per bone + linear blend skinning:
                              rendering.py:303-341, geom_utils.py:87-107, 142-153, 304-348, 906-931)
  render_flow_trans_n16_fp32.npz / render_flow_se3_n16_fp32.npz
                              render_rays with flowbw / flowfw = Transhead / SE3head (rendering.py:257-286,
                              nerf.py:200-237)
each with every result key of the path and all parameter / input gradients, training mode, perturb = 0, noise_std = 0.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moda_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import build_reference_models, flat  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
KEYS = ("img_coarse", "depth_rnd", "sil_coarse", "frame_cyc_dis", "xyz_camera_vis", "xyz_canonical_vis")


def _loss(res):
    return ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()


def _inputs(prob, nets):
    """The rays travel with the fixture; the nets are re-created from the seed by synth.make_motion_problem and pinned
    here by a checksum (sum of |w| per net), so that a change of the generator cannot go unnoticed."""
    out = flat("in.rays.", prob["rays"])
    for net in nets:
        out["in.checksum." + net] = np.float64(sum(float(v.double().abs().sum()) for v in prob[net].values()))
    return out


def golden_lbs(ref, name, n=16, seed=5):
    prob = synth.make_motion_problem(n, "lbs", seed=seed)
    models, emb = build_reference_models(ref, prob)
    rays = {k: v.clone() for k, v in prob["rays"].items()}
    for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
        rays[k].requires_grad_(True)
    for m in (models["coarse"], models["nerf_skin"]):
        m.train()
    opts = synth.default_opts()
    opts.lbs, opts.neudbs = True, False
    res = ref.render_rays(models, emb, rays, N_samples=128, perturb=0, noise_std=0, chunk=32768, img_size=512, opts=opts)
    loss = _loss(res)
    loss.backward()
    out = {"out." + k: res[k].detach().numpy() for k in KEYS}
    out["out.loss"] = loss.detach().numpy()
    out.update(_inputs(prob, ("coarse", "nerf_skin")))
    for net in ("coarse", "nerf_skin"):
        for k, p in models[net].named_parameters():
            out["grad.%s.%s" % (net, k)] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    out["grad.bones_rst"] = models["bones_rst"].grad.numpy()
    out["grad.skin_aux"] = models["skin_aux"].grad.numpy()
    out["grad.rest_pose_code"] = models["rest_pose_code"].weight.grad.numpy()
    for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
        out["grad.rays." + k] = rays[k].grad.numpy()
    # per-function vectors of the LBS helpers
    G = ref.geom_utils
    B = prob["num_bones"]
    v = prob["rays"]["bone_rts"].reshape(n, B, 12)
    rts = torch.cat([v[..., :9].reshape(n, B, 3, 3), v[..., 9:, None]], -1)
    out["fn.bone_transform"] = G.bone_transform(prob["bones_rst"][None], prob["rays"]["bone_rts"], False, is_vec=True).numpy()
    out["fn.rts_invert"] = G.rts_invert(rts).numpy()
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "loss", float(loss), "cyc", float(res["frame_cyc_dis"].mean()))


def golden_flow(ref, kind, name, n=16, seed=6):
    prob = synth.make_motion_problem(n, kind, seed=seed)
    coarse = ref.NeRF(in_channels_xyz=63, in_channels_dir=27 + 64, init_beta=0.1)
    coarse.load_state_dict(prob["coarse"])
    arch, oc = (ref.nerf.Transhead, 3) if kind == "trans" else (ref.nerf.SE3head, 9)
    models = {"coarse": coarse}
    for k in ("flowbw", "flowfw"):
        m = arch(in_channels_xyz=63 + 128, D=5, W=128, out_channels=oc, in_channels_dir=0, raw_feat=True)
        m.load_state_dict(prob[k])
        models[k] = m
    emb = {"xyz": ref.Embedding(3, 10, alpha=10), "dir": ref.Embedding(3, 4, alpha=10)}
    rays = {k: v.clone() for k, v in prob["rays"].items()}
    for k in ("time_embedded", "env_code", "rays_o", "rays_d"):
        rays[k].requires_grad_(True)
    for m in models.values():
        m.train()
    res = ref.render_rays(models, emb, rays, N_samples=128, perturb=0, noise_std=0, chunk=32768, img_size=512,
                          opts=synth.default_opts())
    loss = _loss(res)
    loss.backward()
    out = {"out." + k: res[k].detach().numpy() for k in KEYS}
    out["out.loss"] = loss.detach().numpy()
    out.update(_inputs(prob, ("coarse", "flowbw", "flowfw")))
    for net in ("coarse", "flowbw", "flowfw"):
        for k, p in models[net].named_parameters():
            out["grad.%s.%s" % (net, k)] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    for k in ("time_embedded", "env_code", "rays_o", "rays_d"):
        out["grad.rays." + k] = rays[k].grad.numpy()
    np.savez_compressed(os.path.join(OUT, name), **out)
    d = (res["xyz_camera_vis"] - res["xyz_canonical_vis"]).norm(2, -1)
    print(name, "loss", float(loss), "cyc", float(res["frame_cyc_dis"].mean()), "mean |flow_bw|", float(d.mean()))


if __name__ == "__main__":
    ref = ref_loader.load()
    golden_lbs(ref, "render_lbs_n16_fp32.npz")
    golden_flow(ref, "trans", "render_flow_trans_n16_fp32.npz")
    golden_flow(ref, "se3", "render_flow_se3_n16_fp32.npz")
