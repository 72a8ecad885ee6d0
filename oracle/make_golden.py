"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by executing the REAL reference.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
The fixtures carry both the inputs (so nothing depends on RNG reproducibility across machines) and
the reference's outputs / gradients.  Sizes are kept small (<= 64 rays) so the files stay a few MB.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from moda_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def build_reference_models(ref, prob, dtype=torch.float32):
    """models / embeddings dicts exactly as nnutils/moda.py:271-329 builds them."""
    nb = prob["num_bones"]
    coarse = ref.NeRF(in_channels_xyz=63, in_channels_dir=27 + 64, init_beta=0.1)
    skin = ref.NeRF(in_channels_xyz=63 + 128, D=5, W=64, in_channels_dir=0, out_channels=nb,
                    raw_feat=True, in_channels_code=128)
    coarse.load_state_dict(prob["coarse"])
    skin.load_state_dict(prob["nerf_skin"])
    rest = torch.nn.Embedding(1, 128)
    rest.weight.data.copy_(prob["rest_pose_code"])
    coarse, skin, rest = coarse.to(dtype), skin.to(dtype), rest.to(dtype)
    bones_rst = prob["bones_rst"].clone().to(dtype).requires_grad_(True)
    skin_aux = prob["skin_aux"].clone().to(dtype).requires_grad_(True)
    models = {"coarse": coarse, "bones": bones_rst, "bones_rst": bones_rst, "skin_aux": skin_aux,
              "nerf_skin": skin, "rest_pose_code": rest}
    emb = {"xyz": ref.Embedding(3, 10, alpha=10), "dir": ref.Embedding(3, 4, alpha=10)}
    return models, emb


def flat(prefix, d):
    return {prefix + k: v.detach().cpu().numpy() for k, v in d.items()}


def golden_render(ref, n_rays, seed, name, dtype=torch.float32, use_fine=False, perturb=0.0, full_grads=True):
    prob = synth.make_problem(n_rays, seed=seed)
    models, emb = build_reference_models(ref, prob, dtype)
    rays = {k: v.clone().to(dtype) for k, v in prob["rays"].items()}
    for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
        rays[k].requires_grad_(True)
    for m in (models["coarse"], models["nerf_skin"]):
        m.train()
    opts = synth.default_opts()
    torch.manual_seed(1234)
    res = ref.render_rays(models, emb, rays, N_samples=128, perturb=perturb, noise_std=0, chunk=32768,
                          use_fine=use_fine, img_size=512, opts=opts)
    loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() \
        + res["frame_cyc_dis"].mean()
    loss.backward()
    out = {}
    out.update(flat("in.rays.", prob["rays"]))
    for k in ("bones_rst", "skin_aux", "rest_pose_code"):
        out["in." + k] = prob[k].numpy()
    for k in ("img_coarse", "depth_rnd", "sil_coarse", "frame_cyc_dis", "xyz_camera_vis",
              "xyz_canonical_vis"):
        out["out." + k] = res[k].detach().numpy()
    out["out.loss"] = loss.detach().numpy()
    keep = None if full_grads else ("xyz_encoding_1.0", "xyz_encoding_5.0", "xyz_encoding_8.0.bias", "sigma",
                                    "rgb.0", "dir_encoding.0", "beta", "xyz_encoding_final.bias")
    for k, p in models["coarse"].named_parameters():
        if keep is None or any(k.startswith(x) for x in keep):
            out["grad.coarse." + k] = p.grad.numpy()
    for k, p in models["nerf_skin"].named_parameters():
        # the sigma head of nerf_skin is computed and discarded (nerf.py:178): no gradient
        out["grad.nerf_skin." + k] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
    out["grad.bones_rst"] = models["bones_rst"].grad.numpy()
    out["grad.skin_aux"] = models["skin_aux"].grad.numpy()
    out["grad.rest_pose_code"] = models["rest_pose_code"].weight.grad.numpy()
    for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d"):
        out["grad.rays." + k] = rays[k].grad.numpy()
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "loss", float(loss), "sil mean", float(res["sil_coarse"].mean()))


def golden_geometry(ref, name):
    """Per-function vectors: Embedding, bone_transform, skinning, dqs blend, dq algebra, compositing,
    sample_pdf, density grid."""
    G = ref.geom_utils
    DQ = ref.dual_quat
    prob = synth.make_problem(8, seed=0)
    sp = synth.make_skin_problem(8, 16, seed=0)
    gen = torch.Generator().manual_seed(7)
    out = {}
    x = torch.randn(5, 7, 3, generator=gen) * 0.3
    out["embed.x"] = x.numpy()
    for a in (10, 6.4, 2.0):
        e = ref.Embedding(3, 10, alpha=a)
        out["embed.xyz.alpha%s" % a] = e(x).numpy()
    out["embed.dir"] = ref.Embedding(3, 4, alpha=10)(x).numpy()

    bones, rts, aux, xyz = sp["bones_rst"], sp["bone_rts"], sp["skin_aux"], sp["xyz"]
    bd = G.bone_transform(bones, rts, True, is_vec=True)
    dsk = 0.5 * torch.randn(8, 16, 25, generator=gen)
    skin_bw = G.skinning(bd, xyz, dsk, skin_aux=aux)
    skin_rest = G.skinning(bones, xyz, None, skin_aux=aux)
    xyz_can, bd2, _ = G.neu_dbs(bones, rts, skin_bw, xyz, backward=True)
    xyz_fw, _, _ = G.neu_dbs(bones, rts, skin_rest, xyz, backward=False)
    blend = G.dqs_blend_skinning(rts.view(-1, 25, 8), skin_bw, xyz)
    out.update({"geom.bones": bones.numpy(), "geom.rts": rts.numpy(), "geom.skin_aux": aux.numpy(),
                "geom.xyz": xyz.numpy(), "geom.dskin": dsk.numpy(), "geom.bones_dfm": bd.numpy(),
                "geom.skin_bw": skin_bw.numpy(), "geom.skin_rest": skin_rest.numpy(),
                "geom.xyz_can": xyz_can.numpy(), "geom.xyz_fw": xyz_fw.numpy(),
                "geom.blend": blend.numpy()})

    a = torch.randn(6, 25, 8, generator=gen)
    b = torch.randn(6, 25, 8, generator=gen)
    out.update({"dq.a": a.numpy(), "dq.b": b.numpy(), "dq.mul": DQ.dq_mul(a, b).numpy(),
                "dq.normalize": DQ.dq_normalize(a).numpy(), "dq.inverse": DQ.dq_inverse(a).numpy(),
                "dq.qconj": DQ.dq_quaternion_conjugate(a).numpy(),
                "dq.cconj": DQ.dq_combined_conjugate(a).numpy(),
                "dq.qmul": DQ.q_mul(a[..., :4].reshape(-1, 4), b[..., :4].reshape(-1, 4)).numpy(),
                "dq.qnormalize": DQ.q_normalize(a[..., :4].reshape(-1, 4)).numpy()})

    # compositing through the reference's `inference` with a fixed MLP (rendering.py:124-237)
    models, emb = build_reference_models(ref, prob)
    rays = prob["rays"]
    R, S = 8, 32
    z = ref.rendering.torch.linspace(0, 1, S)
    z = rays["near"] * (1 - z) + rays["far"] * z
    pts = rays["rays_o"][:, None] + rays["rays_d"][:, None] * z[..., None]
    dn = rays["rays_d"] / rays["rays_d"].norm(2, -1)[:, None]
    de = emb["dir"](dn)
    with torch.no_grad():
        rgb, feat, depth, w, vis, sil = ref.rendering.inference(
            models, emb["xyz"], pts, rays["rays_d"], de, z, R, S, 32768, 0.0,
            env_code=rays["env_code"])
        raw = G.evaluate_mlp(models["coarse"], pts, embed_xyz=emb["xyz"],
                             dir_embedded=de[:, None].repeat(1, S, 1), code=rays["env_code"], chunk=4096)
    out.update({"comp.pts": pts.numpy(), "comp.z": z.numpy(), "comp.rays_d": rays["rays_d"].numpy(),
                "comp.env_code": rays["env_code"].numpy(), "comp.raw": raw.numpy(),
                "comp.rgb": rgb.numpy(), "comp.depth": depth.numpy(), "comp.sil": sil.numpy(),
                "comp.weights": w.numpy(), "comp.vis": vis.numpy(), "comp.dir_embedded": de.numpy()})

    wts = torch.rand(8, 30, generator=gen)
    bins = torch.sort(torch.rand(8, 31, generator=gen), -1)[0]
    out.update({"pdf.weights": wts.numpy(), "pdf.bins": bins.numpy(),
                "pdf.det": ref.rendering.sample_pdf(bins, wts, 32, det=True).numpy()})
    u = torch.rand(8, 32, generator=gen)
    torch.manual_seed(5)
    u_ref = torch.rand(8, 32)
    torch.manual_seed(5)
    out.update({"pdf.u": u_ref.numpy(), "pdf.rand": ref.rendering.sample_pdf(bins, wts, 32, det=False).numpy()})

    # density grid (train_utils.py:1377-1404) at G=12, bound 0.3
    Gs = 12
    ax = np.linspace(-0.3, 0.3, Gs).astype(np.float32)
    qx, qy, qz = np.meshgrid(ax, ax, ax, indexing="ij")
    q = torch.Tensor(np.stack([qx.reshape(-1), qy.reshape(-1), qz.reshape(-1)], -1))
    with torch.no_grad():
        vol = models["coarse"](emb["xyz"](q), sigma_only=True).view(Gs, Gs, Gs)
    out["grid.sigma"] = vol.numpy()
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "written")


def main():
    ref = ref_loader.load()
    os.makedirs(OUT, exist_ok=True)
    # network weights depend only on the seed (they are drawn first): stored once, shared by all files
    prob = synth.make_problem(1, seed=0)
    nets = {}
    nets.update(flat("coarse.", prob["coarse"]))
    nets.update(flat("nerf_skin.", prob["nerf_skin"]))
    np.savez_compressed(os.path.join(OUT, "nets_seed0.npz"), **nets)
    golden_render(ref, 32, 0, "render_n32_fp32.npz", torch.float32)
    torch.set_default_dtype(torch.float64)
    try:
        golden_render(ref, 32, 0, "render_n32_fp64.npz", torch.float64, full_grads=False)
    finally:
        torch.set_default_dtype(torch.float32)
    golden_render(ref, 16, 0, "render_fine_n16_fp32.npz", torch.float32, use_fine=True, full_grads=False)
    golden_geometry(ref, "geometry_fp32.npz")


if __name__ == "__main__":
    main()
