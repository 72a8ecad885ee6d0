"""Density query on a regular grid -- the GPU part of ``v2s_trainer.extract_mesh``
(nnutils/train_utils.py:1364-1441): lattice coordinates [-> backward warp of frame ``embedid``] [-> |x| for symmetric
shapes] -> PE -> nerf_coarse(sigma_only) -> (G,G,G) volume, then the visibility pass that sets the density of
never-observed lattice points to -1 (:1407-1425).  The grid is sharded in x-slabs across ranks (SURVEY.md section
8(e)); marching cubes (PyMCubes on the host in the reference, :1441) stays downstream and consumes the volume."""
import numpy as np
import torch

from . import geom_utils as G


def _lattice(grid_size, bound, dev, i, j):
    ax = [torch.linspace(-float(b), float(b), grid_size, device=dev) for b in bound]
    return torch.stack(torch.meshgrid(ax[0][i:j], ax[1], ax[2], indexing="ij"), -1).reshape(-1, 3)


def density_grid(nerf_coarse, grid_size, bound, embedding_xyz, chunk=1 << 21, x_range=None, nerf_vis=None,
                 symm_shape=False, warp=None):
    """sigma on the [x_range) slab of a grid_size^3 lattice spanning [-bound, bound]^3 (C-order x,y,z).

    nerf_vis:   visibility field; lattice points it calls unobserved (sigmoid < 0.5) get density -1
                (train_utils.py:1407-1425, evaluated at the UN-warped lattice points like the reference)
    symm_shape: evaluate at |x| (:1398-1400)
    warp:       callable (n,3) -> (n,3) applied to each chunk first (the ``warp_bw`` of :1394-1396)"""
    dev = nerf_coarse.sigma.weight.device
    Gs = grid_size
    x0, x1 = (0, Gs) if x_range is None else x_range
    out = torch.empty(x1 - x0, Gs, Gs, device=dev)
    rows = max(1, chunk // (Gs * Gs))
    with torch.no_grad():
        for i in range(x0, x1, rows):
            j = min(x1, i + rows)
            pts = _lattice(Gs, bound, dev, i, j)
            q = pts
            if warp is not None:
                q = warp(q)
            if symm_shape:
                q = torch.cat([q[:, :1].abs(), q[:, 1:]], -1)
            sig = G.evaluate_mlp(nerf_coarse, q.reshape(1, -1, 3), embed_xyz=embedding_xyz, sigma_only=True).reshape(-1)
            if nerf_vis is not None:
                vis = G.evaluate_mlp(nerf_vis, pts.reshape(1, -1, 3), embed_xyz=embedding_xyz)[..., 0].sigmoid().reshape(-1)
                sig = torch.where(vis < 0.5, torch.full_like(sig, -1.0), sig)
            out[i - x0:j - x0] = sig.reshape(j - i, Gs, Gs)
    return out


def extract_volume(model, chunk, grid_size, embedid=None):
    """The volume that ``extract_mesh`` hands to marching cubes (train_utils.py:1365-1425), from a model object with the
    reference's attributes (opts, latest_vars['obj_bound' | 'idk'], nerf_coarse, nerf_vis, embedding_xyz, near_far)."""
    opts = model.opts
    bound = model.latest_vars["obj_bound"] if model.near_far is not None else 1.5 * np.asarray([1, 1, 1])
    warp = None
    if embedid is not None and not getattr(opts, "queryfw", True):
        warp = lambda q: G.warp_bw(opts, model, {}, q, embedid)[0]
    use_vis = (not getattr(opts, "full_mesh", False)) and model.latest_vars["idk"].sum() > 0
    if use_vis and not opts.nerf_vis:
        raise NotImplementedError("compute_point_visibility (train_utils.py:1419-1421) is deprecated in the reference")
    return density_grid(model.nerf_coarse, grid_size, bound, model.embedding_xyz, chunk=max(chunk, grid_size * grid_size),
                        nerf_vis=model.nerf_vis if use_vis else None, symm_shape=getattr(opts, "symm_shape", False), warp=warp)


def grid_to_object(vertices, grid_size, bound):
    """train_utils.py:1442: marching-cubes vertex indices -> object coordinates."""
    return (np.asarray(vertices) - grid_size / 2) / grid_size * 2 * np.asarray(bound)[None, :]
