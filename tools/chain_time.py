"""Per-entry device times of the fused chains at the training size, with checksums to compare library variants:
    [MODA_B200_LIB=gpurun_variants/lib_x.so] python tools/chain_time.py [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moda_b200 import config, geom_utils as G, synth, models as MM, _lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = "cuda"
prob = synth.make_problem(8, seed=0)
models, emb, rays = MM.build_models(prob, dev)
R, S = 8192, 128
gen = torch.Generator().manual_seed(1)
pts = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(dev).requires_grad_(True)
de = torch.randn(R, 27, generator=gen).to(dev).requires_grad_(True)
env = (0.1 * torch.randn(R, 64, generator=gen)).to(dev).requires_grad_(True)
code = (0.1 * torch.randn(R, 128, generator=gen)).to(dev).requires_grad_(True)
wt = torch.randn(R * S, 4, generator=gen).to(dev) * 1e-3
ws = torch.randn(R * S, 25, generator=gen).to(dev) * 1e-3
config.fused = True


def once():
    o1 = G.evaluate_mlp(models["coarse"], pts, embed_xyz=emb["xyz"], dir_embedded=de, code=env)
    o2 = G.evaluate_mlp(models["nerf_skin"], pts, embed_xyz=emb["xyz"], code=code)
    ((o1.reshape(-1, 4) * wt).sum() + (o2.reshape(-1, o2.shape[-1])[:, :25] * ws).sum()).backward()
    return o1, o2


def sums():
    g = [float(pts.grad.double().abs().sum()), float(code.grad.double().abs().sum()), float(env.grad.double().abs().sum())]
    for m in ("coarse", "nerf_skin"):
        g.append(sum(float(p.grad.double().abs().sum()) for p in models[m].parameters() if p.grad is not None))
    return g


for _ in range(2):
    once()
for t in (pts, de, env, code):
    t.grad = None
for m in ("coarse", "nerf_skin"):
    models[m].zero_grad(set_to_none=True)
o1, o2 = once()
torch.cuda.synchronize()
print("lib", _lib.LIB_PATH)
print("checksums out %.6e %.6e grads %s" % (float(o1.double().abs().sum()), float(o2.double().abs().sum()),
                                           " ".join("%.6e" % v for v in sums())))
_lib.PROFILE = {}
for _ in range(reps):
    once()
summ = _lib.profile_summary()
_lib.PROFILE = None
tot = 0.0
for k, (n, ms) in sorted(summ.items(), key=lambda kv: -kv[1][1]):
    print("%-24s n/rep=%5.1f  %8.3f ms/rep" % (k, n / reps, ms / reps))
    tot += ms / reps
print("total %.3f ms/rep" % tot)
