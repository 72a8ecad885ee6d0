"""Two forward + backward calls of the Sinkhorn feature matching at the default-flag step's size (8192 rays x 8000 lattice
points, 16 channels) for ncu:   ncu --set full -k regex:sinkhorn --launch-skip 44 -c 44 python tools/sinkhorn_profile.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from moda_b200.loss_utils import SinkhornMatchFn

dev = torch.device("cuda", 0)
gen = torch.Generator().manual_seed(0)
n, m = 8192, 8000
f = F.normalize(torch.randn(n, 16, generator=gen), 2, -1).to(dev).requires_grad_(True)
v = F.normalize(torch.randn(m, 16, generator=gen), 2, -1).to(dev).requires_grad_(True)
q = (torch.rand(m, 3, generator=gen) - 0.5).to(dev)
for _ in range(2):
    pts = SinkhornMatchFn.apply(f, v, q)
    pts.square().sum().backward()
torch.cuda.synchronize()
print("ok", float(pts.abs().max()))
