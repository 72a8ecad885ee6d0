"""Dev helper: per-entry-point device time of one full-size training step (not a bench number)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moda_b200 import synth, models as MM, _lib
from moda_b200.rendering import render_rays

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
prob = synth.make_problem(N, seed=0)
models, emb, rays = MM.build_models(prob, "cuda")
opts = synth.default_opts()

def step():
    res = render_rays(models, emb, rays, N_samples=128, perturb=1.0, noise_std=0, opts=opts, img_size=512)
    loss = ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()
    loss.backward()
    return loss

for _ in range(2):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    step()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print("step %.2f ms -> %.1f rays/s ; peak mem %.1f GB" % (dt * 1e3, N / dt, torch.cuda.max_memory_allocated() / 1e9))
_lib.PROFILE = {}
step()
summ = _lib.profile_summary()
tot = sum(v[1] for v in summ.values())
for k, (n, ms) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
    print("%-28s calls %4d  %9.3f ms  %5.1f%%" % (k, n, ms, 100 * ms / tot))
print("sum of kernel time %.2f ms" % tot)
