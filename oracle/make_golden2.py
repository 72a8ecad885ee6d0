"""TEST INFRASTRUCTURE ONLY -- round-2 fixtures, again produced by executing the REAL reference.

Run in the build container (needs /root/reference):  python oracle/make_golden2.py
Writes (tests/golden/):
  render_opts_n16_fp32.npz   render_rays with use_disp=True, Embedding.alpha=6.4, noise_std=0.3 (rendering.py:72,
                             nerf.py:63-69, rendering.py:193-196), outputs + all gradients
  render_vis_n16_fp32.npz    render_rays(render_vis=True, obj_bound=...) in eval mode (rendering.py:210-215, 373-379)
  render_full_n16_fp32.npz   the default-flag training step: nerf_feat + nerf_vis + paired target frame
  render_full_n16_fp64.npz   (rendering.py:174-178, 233, 345-352, 405-449, 475-489, 516-579; loss_utils.py:125-149,
                             162-405; geom_utils.py:567-672, 1704-1743), every result key + all gradients
  callers_fp32.npz           raycast / sample_xy / DQ_RTHead / correct_bones / correct_rest_pose / lbs / nerf_dis /
                             warp_fw / warp_bw / symm_shape / flow fields (geom_utils.py:746-1073, nerf.py:200-279)
Every random draw the reference makes inside the call (torch.rand / randn / rand_like / randn_like) is recorded
in call order as rng.<i> so that the CUDA path can be fed the same numbers.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moda_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import build_reference_models, flat  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


class RngTape:
    """Records (or replays) the reference's in-call random draws."""
    NAMES = ("rand", "randn", "rand_like", "randn_like")

    def __init__(self):
        self.draws = []

    def __enter__(self):
        self.orig = {n: getattr(torch, n) for n in self.NAMES}
        for n in self.NAMES:
            setattr(torch, n, self._wrap(self.orig[n]))
        return self

    def _wrap(self, fn):
        def f(*a, **k):
            t = fn(*a, **k)
            self.draws.append(t.detach().cpu().clone())
            return t
        return f

    def __exit__(self, *exc):
        for n in self.NAMES:
            setattr(torch, n, self.orig[n])
        return False

    def dump(self, out):
        for i, t in enumerate(self.draws):
            out["rng.%d" % i] = t.numpy()


def _grads_into(out, models, rays, nets=("coarse", "nerf_skin")):
    for net in nets:
        for k, p in models[net].named_parameters():
            out["grad.%s.%s" % (net, k)] = (p.grad if p.grad is not None else torch.zeros_like(p)).detach().numpy()
    out["grad.bones_rst"] = models["bones_rst"].grad.numpy()
    out["grad.skin_aux"] = models["skin_aux"].grad.numpy()
    out["grad.rest_pose_code"] = models["rest_pose_code"].weight.grad.numpy()
    for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d", "bone_rts_target"):
        if k in rays and rays[k].grad is not None:
            out["grad.rays." + k] = rays[k].grad.numpy()


def golden_opts(ref, name):
    prob = synth.make_problem(16, seed=3)
    models, emb = build_reference_models(ref, prob)
    emb = {"xyz": ref.Embedding(3, 10, alpha=6.4), "dir": ref.Embedding(3, 4, alpha=6.4)}
    rays = {k: v.clone() for k, v in prob["rays"].items()}
    for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
        rays[k].requires_grad_(True)
    for m in (models["coarse"], models["nerf_skin"]):
        m.train()
    torch.manual_seed(77)
    with RngTape() as tape:
        res = ref.render_rays(models, emb, rays, N_samples=128, use_disp=True, perturb=0, noise_std=0.3, chunk=32768,
                              img_size=512, opts=synth.default_opts())
    loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()
    loss.backward()
    out = {}
    out.update(flat("in.rays.", prob["rays"]))
    for k in ("bones_rst", "skin_aux", "rest_pose_code"):
        out["in." + k] = prob[k].numpy()
    out.update(flat("net.coarse.", prob["coarse"]))
    out.update(flat("net.nerf_skin.", prob["nerf_skin"]))
    for k in ("img_coarse", "depth_rnd", "sil_coarse", "frame_cyc_dis", "xyz_camera_vis", "xyz_canonical_vis"):
        out["out." + k] = res[k].detach().numpy()
    out["out.loss"] = loss.detach().numpy()
    _grads_into(out, models, rays)
    tape.dump(out)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "loss", float(loss), "draws", [tuple(t.shape) for t in tape.draws])


def _nerf(ref, sd, **kw):
    m = ref.NeRF(**kw)
    m.load_state_dict(sd)
    return m


def build_full_models(ref, prob, dtype=torch.float32):
    models, emb = build_reference_models(ref, prob, dtype)
    models["nerf_feat"] = _nerf(ref, prob["nerf_feat"], in_channels_xyz=63, D=5, W=128, out_channels=16, in_channels_dir=0,
                                raw_feat=True, init_beta=1.).to(dtype)
    models["nerf_vis"] = _nerf(ref, prob["nerf_vis"], in_channels_xyz=63, D=5, W=64, out_channels=1, in_channels_dir=0,
                               raw_feat=True).to(dtype)
    return models, emb


def golden_vis(ref, name, seed=4):
    prob = synth.make_full_problem(16, seed=seed)
    # a useful fixture masks a good share of the samples both ways: tighter bound, visibility logits centred on 0
    prob["obj_bound"] = prob["obj_bound"] * 0.5
    prob["nerf_vis"]["rgb.0.bias"] = prob["nerf_vis"]["rgb.0.bias"] - 0.12
    models, emb = build_full_models(ref, prob)
    for m in ("coarse", "nerf_skin", "nerf_vis", "nerf_feat"):
        models[m].eval()
    models.pop("nerf_feat")
    rays = {k: prob["rays"][k].clone() for k in ("rays_o", "rays_d", "near", "far", "xys", "time_embedded", "env_code",
                                                  "bone_rts")}
    bound = prob["obj_bound"].numpy()
    with torch.no_grad(), RngTape() as tape:
        res = ref.render_rays(models, emb, rays, N_samples=128, perturb=0, noise_std=0, chunk=32768, obj_bound=bound,
                              img_size=512, opts=synth.default_opts(), render_vis=True)
        # decision margins of the two masks (rendering.py:212, 215): the fixture is only useful if no sample sits
        # on a threshold
        xyz = res["xyz_canonical_vis"]
        vp = ref.geom_utils.evaluate_mlp(models["nerf_vis"], emb["xyz"](xyz), chunk=32768)[..., 0].sigmoid()
    margin_v = float((vp - 0.5).abs().min())
    margin_b = float((xyz.abs() - torch.tensor(bound)[None, None]).abs().min())
    if not (margin_v > 1e-4 and margin_b > 1e-5):
        print("seed", seed, "sits on a mask threshold", margin_v, margin_b, "-> next seed")
        return golden_vis(ref, name, seed + 1)
    out = {}
    out.update(flat("in.rays.", rays))
    for k in ("bones_rst", "skin_aux", "rest_pose_code", "obj_bound"):
        out["in." + k] = prob[k].numpy()
    for net in ("coarse", "nerf_skin", "nerf_vis"):
        out.update(flat("net.%s." % net, prob[net]))
    for k, v in res.items():
        out["out." + k] = v.detach().numpy()
    out["masked_frac"] = np.float32(float(((vp < 0.5) | ((xyz.abs() > torch.tensor(bound)[None, None]).sum(-1) > 0)).float().mean()))
    out["seed"] = np.int64(seed)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "keys", sorted(res), "masked fraction", float(out["masked_frac"]), "margins", margin_v, margin_b)


FULL_LOSS_KEYS = ("img_loss_samp", "sil_loss_samp", "flo_loss_samp", "feat_err", "proj_err", "frnd_loss_samp",
                  "frame_cyc_dis")


def full_loss(res):
    """One scalar touching every differentiable output of the default-flag step (weights as nnutils/moda.py:540-705
    orders of magnitude; the exact mix does not matter for parity, it only has to be the same on both sides)."""
    loss = res["vis_loss"]
    for k in FULL_LOSS_KEYS:
        loss = loss + res[k].mean()
    return loss


def golden_full(ref, name, dtype=torch.float32, replay=None):
    prob = synth.make_full_problem(16, seed=5)
    models, emb = build_full_models(ref, prob, dtype)
    for m in ("coarse", "nerf_skin", "nerf_vis", "nerf_feat"):
        models[m].train()
    rays = {k: v.clone().to(dtype) for k, v in prob["rays"].items()}
    for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d", "bone_rts_target"):
        rays[k].requires_grad_(True)
    bound = prob["obj_bound"].numpy()
    torch.manual_seed(4321)
    if replay is None:
        tape = RngTape()
    else:
        tape = ReplayTape(replay, dtype)
    with tape:
        res = ref.render_rays(models, emb, rays, N_samples=128, perturb=0, noise_std=0, chunk=32768, obj_bound=bound,
                              img_size=prob["img_size"], opts=synth.full_opts())
    loss = full_loss(res)
    loss.backward()
    out = {}
    if replay is None:   # inputs, weights and draws live in the fp32 file only (the fp64 run replays them)
        out.update(flat("in.rays.", prob["rays"]))
        for k in ("bones_rst", "skin_aux", "rest_pose_code", "obj_bound"):
            out["in." + k] = prob[k].numpy()
        for net in ("coarse", "nerf_skin", "nerf_vis", "nerf_feat"):
            out.update(flat("net.%s." % net, prob[net]))
    for k, v in res.items():
        out["out." + k] = v.detach().numpy()
    out["out.loss"] = loss.detach().numpy()
    _grads_into(out, models, rays, nets=("coarse", "nerf_skin", "nerf_vis", "nerf_feat"))
    if replay is None:
        tape.dump(out)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "loss", float(loss), "keys", sorted(res))
    return tape.draws if replay is None else None


class ReplayTape(RngTape):
    """Feeds recorded draws back (used to run the fp64 reference on the fp32 run's random numbers)."""

    def __init__(self, draws, dtype):
        super().__init__()
        self.src, self.dtype, self.i = draws, dtype, 0

    def _wrap(self, fn):
        def f(*a, **k):
            t = self.src[self.i].to(self.dtype)
            self.i += 1
            return t
        return f


def main():
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["opts", "vis", "full", "callers"]
    if "opts" in which:
        golden_opts(ref, "render_opts_n16_fp32.npz")
    if "vis" in which:
        golden_vis(ref, "render_vis_n16_fp32.npz")
    if "full" in which:
        draws = golden_full(ref, "render_full_n16_fp32.npz")
        torch.set_default_dtype(torch.float64)
        try:
            golden_full(ref, "render_full_n16_fp64.npz", torch.float64, replay=draws)
        finally:
            torch.set_default_dtype(torch.float32)
    if "callers" in which:
        from oracle.make_golden_callers import golden_callers
        golden_callers(ref, "callers_fp32.npz")


if __name__ == "__main__":
    main()
