"""Diagnostic: per-ray env_code gradient of the fp16 chain mode against the fp32 SIMT mode at 8192 rays."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moda_b200 import synth, models as MM, config
from moda_b200.rendering import render_rays

N, S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 128
prob = synth.make_problem(N, seed=11)
out = {}
for mode in ("fp32", "fp16"):
    config.set_precision(mode)
    models, emb, rays = MM.build_models(prob, "cuda")
    res = render_rays(models, emb, rays, N_samples=S, perturb=0, noise_std=0, opts=synth.default_opts(), img_size=512)
    loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()
    loss.backward()
    out[mode] = {k: rays[k].grad.clone() for k in ("env_code", "time_embedded", "rays_d")}
for k in out["fp32"]:
    a, b = out["fp32"][k].double(), out["fp16"][k].double()
    d = (a - b).abs()
    print(k, "max|g|", float(a.abs().max()), "max abs diff", float(d.max()), "rms diff", float(d.pow(2).mean().sqrt()),
          "rms g", float(a.pow(2).mean().sqrt()))
    per_ray = d.max(-1).values
    gmax_ray = a.abs().max(-1).values
    top = per_ray.topk(8).indices
    for i in top.tolist():
        print("   ray %5d  max diff %.3e  ray |g|max %.3e  rel-to-ray %.3f" % (i, float(per_ray[i]), float(gmax_ray[i]), float(per_ray[i] / gmax_ray[i])))
    rel = (per_ray / (gmax_ray + 1e-30))
    print("   per-ray relative error quantiles 50/90/99/max: %.2e %.2e %.2e %.2e" % tuple(float(rel.quantile(q)) for q in (0.5, 0.9, 0.99, 1.0)))
