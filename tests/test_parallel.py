"""CPU tests (gloo, world_size 2) of the data-parallel host logic: ray sharding + the single flat-buffer
all-reduce reproduce the single-process gradient."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from moda_b200.parallel import FlatParams, shard_rays


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _toy_loss(params, rays, n_total):
    w, b = params
    pred = rays["x"] @ w + b
    return ((pred - rays["y"]) ** 2).sum() / n_total  # each rank scales by its share of the global batch


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    rays = {"x": torch.randn(10, 4, generator=g), "y": torch.randn(10, 3, generator=g), "meta": 7}
    w = torch.randn(4, 3, generator=g).requires_grad_(True)
    b = torch.randn(3, generator=g).requires_grad_(True)
    fp = FlatParams([w, b])
    mine = shard_rays(rays, rank, world)
    assert mine["meta"] == 7 and mine["x"].shape[0] == 5
    fp.zero_grad()
    _toy_loss([w, b], mine, 10).backward()
    fp.allreduce()
    ret[rank] = torch.cat([w.grad.reshape(-1), b.grad.reshape(-1)]).clone()
    dist.destroy_process_group()


def test_flat_allreduce_matches_single_process():
    world = 2
    port = _free_port()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")  # spawned workers import this module
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    g = torch.Generator().manual_seed(0)
    rays = {"x": torch.randn(10, 4, generator=g), "y": torch.randn(10, 3, generator=g)}
    w = torch.randn(4, 3, generator=g).requires_grad_(True)
    b = torch.randn(3, generator=g).requires_grad_(True)
    _toy_loss([w, b], rays, 10).backward()
    ref = torch.cat([w.grad.reshape(-1), b.grad.reshape(-1)])
    for r in range(world):
        assert torch.allclose(ret[r], ref, atol=1e-6), r


def test_flat_params_views():
    a = torch.randn(3, 2).requires_grad_(True)
    b = torch.randn(5).requires_grad_(True)
    a0, b0 = a.detach().clone(), b.detach().clone()
    fp = FlatParams([a, b])
    assert torch.equal(a.detach(), a0) and torch.equal(b.detach(), b0)
    (a.sum() * 2 + (b ** 2).sum()).backward()
    o = fp.offsets
    assert o[0] == 0 and o[1] % FlatParams.ALIGN == 0
    assert torch.allclose(fp.grad[:6], torch.full((6,), 2.0))
    assert torch.allclose(fp.grad[o[1]:o[1] + 5], 2 * b0)
    assert float(fp.grad[6:o[1]].abs().sum()) == 0  # the alignment padding stays inert
    with torch.no_grad():
        fp.flat.add_(1.0)
    assert torch.allclose(a.detach(), a0 + 1)


def test_shard_rays_ragged():
    rays = {"x": torch.arange(10), "k": "v"}
    parts = [shard_rays(rays, r, 4)["x"] for r in range(4)]
    assert torch.equal(torch.cat(parts), rays["x"])
    assert [p.numel() for p in parts] == [3, 3, 3, 1]


def test_flat_params_refuses_detached_gradient_views():
    """the fused backward accumulates straight into the .grad views of the flat buffer; dropping those views
    (optimizer.zero_grad(set_to_none=True), module.zero_grad()) must raise instead of silently freezing training."""
    import pytest
    from moda_b200.parallel import FlatParams
    a, b = torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(7))
    fp = FlatParams([a, b])
    fp.check()
    opt = torch.optim.SGD([fp.flat], lr=0.1)
    fp.zero_grad()
    fp.check()
    opt.zero_grad()          # set_to_none=True by default: flat.grad becomes None
    with pytest.raises(RuntimeError):
        fp.allreduce()
    fp2 = FlatParams([torch.nn.Parameter(torch.randn(4))])
    fp2.tensors[0].grad = None
    with pytest.raises(RuntimeError):
        fp2.check()
