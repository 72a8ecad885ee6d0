"""Warm per-kernel device times of one training step from torch.profiler (CUPTI), as opposed to ncu's cold-cache serialised
launches: python tools/kineto_step.py [rays]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from moda_b200 import synth, models as MM
from moda_b200.parallel import FlatParams
from moda_b200.rendering import render_rays

R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
prob = synth.make_problem(R, seed=0)
models, emb, rays = MM.build_models(prob, dev)
models["coarse"].train(); models["nerf_skin"].train()
opts = synth.default_opts()
flat = FlatParams(MM.parameters_of(models))


def step():
    flat.zero_grad()
    res = render_rays(models, emb, rays, N_samples=128, perturb=1.0, noise_std=0.0, chunk=32768, img_size=512, opts=opts)
    loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()
    loss.backward()
    flat.adamw_step(lr=1e-4)


for _ in range(3):
    step()
torch.cuda.synchronize()
N = 5
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
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%-70s n=%5.1f %8.1f us/step  %5.1f us each" % (k, n / N, t / N, t / n))
