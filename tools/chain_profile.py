"""Runs the fused chains once at the training size (for ncu captures): python tools/chain_profile.py [trunk|skin]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moda_b200 import config, geom_utils as G, synth, models as MM

which = sys.argv[1] if len(sys.argv) > 1 else "trunk"
dev = "cuda"
prob = synth.make_problem(8, seed=0)
models, emb, rays = MM.build_models(prob, dev)
R, S = 8192, 128
gen = torch.Generator().manual_seed(1)
pts = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(dev).requires_grad_(True)
config.fused = True
for _ in range(2):
    if which == "trunk":
        de = torch.randn(R, 27, generator=gen).to(dev).requires_grad_(True)
        env = (0.1 * torch.randn(R, 64, generator=gen)).to(dev).requires_grad_(True)
        out = G.evaluate_mlp(models["coarse"], pts, embed_xyz=emb["xyz"], dir_embedded=de, code=env)
    else:
        code = (0.1 * torch.randn(R, 128, generator=gen)).to(dev).requires_grad_(True)
        out = G.evaluate_mlp(models["nerf_skin"], pts, embed_xyz=emb["xyz"], code=code)
    (out * 1e-3).sum().backward()
    torch.cuda.synchronize()
print("done")
