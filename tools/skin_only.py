import sys, os
sys.path.insert(0, "/root/repo")
import torch
from moda_b200 import config, geom_utils as G, synth, models as MM
dev = "cuda"
prob = synth.make_problem(8, seed=0)
models, emb, rays = MM.build_models(prob, dev)
skin = models["nerf_skin"]
R, S = 3, 128
gen = torch.Generator().manual_seed(7)
pts = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(dev).requires_grad_(True)
code = (0.1 * torch.randn(R, 128, generator=gen)).to(dev).requires_grad_(True)
config.fused = True
out = G.evaluate_mlp(skin, pts, embed_xyz=emb["xyz"], code=code)
torch.cuda.synchronize()
print("fwd ok", float(out.abs().max()))
(out * 1e-3).sum().backward()
torch.cuda.synchronize()
print("bwd ok")
