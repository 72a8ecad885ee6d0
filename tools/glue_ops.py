"""Which python lines launch the ATen glue kernels of a step (a TorchDispatchMode logging op, source line and output size):
python tools/glue_ops.py [rays] [full]"""
import sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from moda_b200 import synth, models as MM
from moda_b200.parallel import FlatParams
from moda_b200.rendering import render_rays
R = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
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
for _ in range(3): step()
torch.cuda.synchronize()
import traceback
from torch.utils._python_dispatch import TorchDispatchMode
SKIP = ("view", "as_strided", "reshape", "slice", "select", "expand", "t.default", "transpose", "permute", "detach", "unsqueeze",
        "squeeze", "narrow", "alias", "unbind", "split", "empty", "_local_scalar", "lift_fresh", "unflatten", "flatten")
agg = collections.Counter()
size = collections.Counter()
class Log(TorchDispatchMode):
    def __torch_dispatch__(self, func, types, args=(), kwargs=None):
        name = str(func)
        if not any(k in name for k in SKIP):
            fr = [f for f in traceback.extract_stack() if "/moda_b200/" in f.filename]
            where = "%s:%d %s" % (os.path.basename(fr[-1].filename), fr[-1].lineno, fr[-1].name) if fr else "(autograd / tool)"
            agg[(where, name)] += 1
        out = func(*args, **(kwargs or {}))
        if not any(k in name for k in SKIP) and isinstance(out, torch.Tensor):
            size[(where, name)] += out.numel()
        return out
with Log():
    step()
torch.cuda.synchronize()
for (w, n), c in sorted(agg.items()):
    print("%3d %-40s %-44s %10.2f M elements out" % (c, n, w, size[(w, n)] / 1e6))
print("total", sum(agg.values()))
