"""Event timeline of one tile inside the fused chain kernel (block 0, third tile): python tools/chain_trace.py [trunk|skin] [fwd|bwd]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from moda_b200 import config, geom_utils as G, synth, models as MM, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "trunk"
phase = sys.argv[2] if len(sys.argv) > 2 else "fwd"
dev = "cuda"
prob = synth.make_problem(8, seed=0)
models, emb, rays = MM.build_models(prob, dev)
R, S = 8192, 128
gen = torch.Generator().manual_seed(1)
pts = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(dev).requires_grad_(True)
config.fused = True
buf = torch.zeros(32 + 4 * 1000 * 32, dtype=torch.int64, device=dev)


def run(trace_fwd, trace_bwd):
    if which == "trunk":
        de = torch.randn(R, 27, generator=gen).to(dev).requires_grad_(True)
        env = (0.1 * torch.randn(R, 64, generator=gen)).to(dev).requires_grad_(True)
        _lib.lib().moda_chain_set_trace(buf.data_ptr() if trace_fwd else None)
        out = G.evaluate_mlp(models["coarse"], pts, embed_xyz=emb["xyz"], dir_embedded=de, code=env)
    else:
        code = (0.1 * torch.randn(R, 128, generator=gen)).to(dev).requires_grad_(True)
        _lib.lib().moda_chain_set_trace(buf.data_ptr() if trace_fwd else None)
        out = G.evaluate_mlp(models["nerf_skin"], pts, embed_xyz=emb["xyz"], code=code)
    _lib.lib().moda_chain_set_trace(buf.data_ptr() if trace_bwd else None)
    (out * 1e-3).sum().backward()
    _lib.lib().moda_chain_set_trace(None)
    torch.cuda.synchronize()


run(False, False)
run(phase == "fwd", phase == "bwd")
b = buf.cpu()
rec = []
for region in range(32):
    n = min(int(b[region]), 1000)
    rec += b[32 + 4000 * region:32 + 4000 * region + 4 * n].reshape(n, 4).tolist()
rec.sort(key=lambda r: r[3])
t0 = rec[0][3]
names = {0: "mma  step begin (wait acc_free)", 1: "mma  acc_free seen", 2: "mma  all MMAs issued", 3: "mma  W stage full", 4: "mma  step committed", 5: "mma  A chunk ready", 6: "mma  chunk issued", 7: "mma  stores-done seen",
         10: "epi0 acc_full seen", 11: "epi0 chunk written", 12: "epi0 store-read waited", 13: "epi0 barrier passed",
         20: "stor chunk ready seen", 21: "stor smem read done",
         30: "pe   computed", 31: "pe   chunk free"}
for ev, a, bb, t in rec:
    if 200 <= ev < 210:
        nm = names.get(ev - 200, str(ev)).replace("mma ", "mma1")
    elif ev >= 40:
        nm = "epi%d.%d %s" % ((ev - 40) // 80, ((ev - 40) % 80) // 10, {0: "acc_full seen", 1: "chunk written", 2: "step done (acc_free arrive)"}[ev % 10])
    else:
        nm = names.get(ev, str(ev))
    print("%8d  %-28s %3d %3d" % (t - t0, nm, a, bb))
