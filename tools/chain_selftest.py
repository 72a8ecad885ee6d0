"""GPU self-test of the fused chain kernels against the layer-by-layer tensor-core modules (same math, same
precision policy): values, every gradient, and timings at the training size."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moda_b200 import config, geom_utils as G, synth, models as MM, _lib
from moda_b200.nerf import Embedding

dev = "cuda"
torch.manual_seed(0)
prob = synth.make_problem(8, seed=0)
models, emb, rays = MM.build_models(prob, dev)
coarse, skin = models["coarse"], models["nerf_skin"]
embx = emb["xyz"]
ok = True


def nrel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def run(model, kind, R, S, fused, gen_seed, single_code=False):
    gen = torch.Generator().manual_seed(gen_seed)
    pts = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(dev).requires_grad_(True)
    model.zero_grad()
    config.fused = fused
    if kind == "trunk":
        de = torch.randn(R, 27, generator=gen).to(dev).requires_grad_(True)
        env = (0.1 * torch.randn(R, 64, generator=gen)).to(dev).requires_grad_(True)
        gout = torch.randn(R, S, 4, generator=gen).to(dev) * 1e-3
        out = G.evaluate_mlp(model, pts, embed_xyz=embx, dir_embedded=de, code=env)
        (out * gout).sum().backward()
        grads = {"pts": pts.grad, "dir": de.grad, "env": env.grad}
    else:
        code = (0.1 * torch.randn(1 if single_code else R, 128, generator=gen)).to(dev).requires_grad_(True)
        gout = torch.randn(R, S, 25, generator=gen).to(dev) * 1e-3
        out = G.evaluate_mlp(model, pts, embed_xyz=embx, code=code)
        (out * gout).sum().backward()
        grads = {"pts": pts.grad, "code": code.grad}
    for k, p in model.named_parameters():
        if p.grad is not None:
            grads[k] = p.grad.clone()
    return out.detach().clone(), grads


for kind, model in (("trunk", coarse), ("skin", skin)):
    for (R, S, single) in ((3, 128, False), (300, 128, False), (37, 128, True)):
        if kind == "trunk" and single:
            continue
        a = run(model, kind, R, S, False, 7, single)
        b = run(model, kind, R, S, True, 7, single)
        e = float((a[0] - b[0]).abs().max())
        tol = 2e-3 if kind == "trunk" else 2e-5
        flag = e < tol
        ok &= flag
        print("%s R=%d single=%d  out max|diff| %.3e %s" % (kind, R, single, e, "ok" if flag else "FAIL"))
        worst = max((nrel(b[1][k], a[1][k]), k) for k in a[1])
        flag = worst[0] < 2e-2 and set(a[1]) == set(b[1])
        ok &= flag
        print("   grads worst nrel %.3e (%s) %s" % (worst[0], worst[1], "ok" if flag else "FAIL"))
        if not flag:
            for k in a[1]:
                print("      %-28s %.3e" % (k, nrel(b[1][k], a[1][k])))
print("ALL OK" if ok else "SOME FAILED")

# timings at the training size
R, S = 8192, 128
for kind, model in (("trunk", coarse), ("skin", skin)):
    for fused in (False, True):
        run(model, kind, R, S, fused, 3)
        torch.cuda.synchronize()
        _lib.PROFILE = {}
        run(model, kind, R, S, fused, 3)
        summ = _lib.profile_summary()
        _lib.PROFILE = None
        tot = sum(ms for _, ms in summ.values())
        print("%s fused=%d: %.3f ms in C-ABI kernels" % (kind, fused, tot))
        for k, (n, ms) in sorted(summ.items(), key=lambda kv: -kv[1][1])[:8]:
            print("      %-24s n=%3d %8.3f ms" % (k, n, ms))
