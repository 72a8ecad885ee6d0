"""Warm per-kernel device times of one training step from torch.profiler (CUPTI), as opposed to ncu's cold-cache serialised
launches: python tools/kineto_step.py [rays] [full]   ("full" = the default-flag step of bench.py --workload full)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from moda_b200 import synth, models as MM
from moda_b200.parallel import FlatParams
from moda_b200.rendering import render_rays

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
FULL = len(sys.argv) > 2 and sys.argv[2] == "full"
if FULL:
    prob = synth.make_full_problem(R, seed=0)
    models, emb, rays = MM.build_full_models(prob, dev)
    for k in ("coarse", "nerf_skin", "nerf_feat", "nerf_vis"):
        models[k].train()
    opts = synth.full_opts()
    bound = prob["obj_bound"].numpy()
    KEYS = ("img_loss_samp", "sil_loss_samp", "flo_loss_samp", "feat_err", "proj_err", "frnd_loss_samp", "frame_cyc_dis")
else:
    prob = synth.make_problem(R, seed=0)
    models, emb, rays = MM.build_models(prob, dev)
    models["coarse"].train(); models["nerf_skin"].train()
    opts = synth.default_opts()
flat = FlatParams(MM.parameters_of(models))


def step():
    flat.zero_grad()
    if FULL:
        res = render_rays(models, emb, rays, N_samples=128, perturb=1.0, noise_std=0.0, chunk=32768, obj_bound=bound,
                          img_size=prob["img_size"], opts=opts)
        loss = res["vis_loss"]
        for k in KEYS:
            loss = loss + res[k].mean()
    else:
        res = render_rays(models, emb, rays, N_samples=128, perturb=1.0, noise_std=0.0, chunk=32768, img_size=512, opts=opts)
        loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()
    loss.backward()
    flat.adamw_step(lr=1e-4)


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 5
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(N):
    step()
t1.record(); torch.cuda.synchronize()
print("# eager step: %.3f ms" % (t0.elapsed_time(t1) / N))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
agg = {}
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        k = e.name[:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
print("# %d rays: %.3f ms of kernel time per step, %d launches per step" % (R, tot / N / 1e3, sum(v[0] for v in agg.values()) / N))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print("%-70s n=%5.1f %8.1f us/step  %5.1f us each" % (k, n / N, t / N, t / n))
