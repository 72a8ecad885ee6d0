"""Density query on a regular grid -- the GPU part of ``v2s_trainer.extract_mesh``
(nnutils/train_utils.py:1377-1404): lattice coordinates -> PE -> nerf_coarse(sigma_only) -> (G,G,G) volume.
The grid is sharded in x-slabs across ranks; marching cubes stays downstream."""
import torch

from . import geom_utils as G


def density_grid(nerf_coarse, grid_size, bound, embedding_xyz, chunk=1 << 21, x_range=None):
    """Returns sigma on the [x_range) slab of a grid_size^3 lattice spanning [-bound, bound]^3 (C-order x,y,z)."""
    dev = nerf_coarse.sigma.weight.device
    Gs = grid_size
    ax = [torch.linspace(-float(b), float(b), Gs, device=dev) for b in bound]
    x0, x1 = (0, Gs) if x_range is None else x_range
    out = torch.empty(x1 - x0, Gs, Gs, device=dev)
    rows = max(1, chunk // (Gs * Gs))
    with torch.no_grad():
        for i in range(x0, x1, rows):
            j = min(x1, i + rows)
            pts = torch.stack(torch.meshgrid(ax[0][i:j], ax[1], ax[2], indexing="ij"), -1).reshape(1, -1, 3)
            sig = G.evaluate_mlp(nerf_coarse, pts, embed_xyz=embedding_xyz, sigma_only=True)
            out[i - x0:j - x0] = sig.reshape(j - i, Gs, Gs)
    return out
