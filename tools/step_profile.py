"""Every C-ABI call of one training step in launch order with its device time:
    python tools/step_profile.py [rays]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moda_b200 import synth, models as MM, _lib, config
config.side_stream = False   # one stream: per-call intervals do not overlap
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


for _ in range(3):
    step()
torch.cuda.synchronize()
order = []
orig = _lib.call


def call(name, *a):
    order.append(name)
    return orig(name, *a)


_lib.PROFILE = {}
# the wrappers bound `call` at import time: patch the modules that use it
import moda_b200.ops as ops, moda_b200.chain_tc as ct
mods = [m for m in (ops, ct) if hasattr(m, "call")]
for m in mods:
    m.call = call
step()
torch.cuda.synchronize()
idx = {}
tot = 0.0
for name in order:
    i = idx.get(name, 0)
    idx[name] = i + 1
    e0, e1 = _lib.PROFILE[name][i]
    ms = e0.elapsed_time(e1)
    tot += ms
    print("%-26s %8.3f ms" % (name, ms))
print("sum %.3f ms" % tot)
