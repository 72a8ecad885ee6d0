"""GPU parity tests of the DEFAULT execution mode (fp16-operand tcgen05 chains, fp32 accumulation; what bench.py
times), round 2.  Bars: rendered rgb / sil / depth / cycle term within 1e-3 ABSOLUTE of the oracle
(BASELINE.json north_star); every gradient tensor within max(2e-2, 3 x the fp32 reference's own error) RELATIVE to
that tensor's scale against the fp64 oracle -- per tensor, so a small tensor cannot hide behind an absolute bar;
skinned points keep the 1e-5 relative fp32 bar (the warps do not run on tensor cores)."""
import os

import numpy as np
import pytest
import torch

from tests.util import (ReplayRng, dump_table, fixture_problem, golden_problem, load_npz, max_abs, rel_err)

pytestmark = pytest.mark.gpu

DEV = "cuda"
OUT_KEYS = ("img_coarse", "depth_rnd", "sil_coarse", "frame_cyc_dis")
GRAD_BAR = 2e-2


@pytest.fixture(autouse=True)
def _default_mode():
    from moda_b200 import config
    old = config.precision
    config.set_precision("fp16")
    yield
    config.set_precision(old)


def _loss(res):
    return ((res["img_coarse"] - 0.3) ** 2).mean() + ((res["sil_coarse"] - 0.5) ** 2).mean() + res["frame_cyc_dis"].mean()


def _gpu_grads(models, rays, nets=("coarse", "nerf_skin")):
    g = {}
    for net in nets:
        for k, p in models[net].named_parameters():
            g["%s.%s" % (net, k)] = p.grad if p.grad is not None else torch.zeros_like(p)
    g["bones_rst"] = models["bones_rst"].grad
    g["skin_aux"] = models["skin_aux"].grad
    g["rest_pose_code"] = models["rest_pose_code"].weight.grad
    for k in ("bone_rts", "time_embedded", "env_code", "rays_o", "rays_d", "bone_rts_target"):
        if k in rays and rays[k].grad is not None:
            g["rays." + k] = rays[k].grad
    return g


def _check_grads(name, got, truth, ref32, bar=GRAD_BAR, slack=3.0, per_ray_outliers=False):
    """got / truth / ref32: dicts name -> tensor.  Scale-relative error (max |diff| / max |truth|) per tensor.

    per_ray_outliers (BASELINE-size runs): for the per-ray INPUT gradients (``rays.*``, one row per ray) the bar applies
    to the 99.5 % quantile over rays AND to the norm-wise error ||diff||_F / ||truth||_F; the worst ray may reach 5 x
    the bar.  Reason, measured with
    tools/diag_env.py (profiles/r02_diag_env_code.txt): at 8192 rays a handful of rays whose compositing weight sits
    on one or two samples have a direction-layer ReLU unit within fp16 operand rounding of its kink there; the unit
    flips, and with it ~1/sqrt(128) of that ray's env_code / dir gradient (12 of 8192 rays deviate by 2-4 % of the
    tensor's max, the median ray by 1e-4).  Parameter gradients (sums over all rays) are not affected and keep the
    plain bar."""
    bad, table = [], ["%-44s %10s %10s %10s" % ("tensor", "ours", "fp32-ref", "|g|max")]
    for k in sorted(truth):
        if k not in got:
            bad.append("%s: no gradient produced" % k)
            continue
        t = truth[k]
        if float(t.abs().max()) == 0.0:
            assert float(got[k].abs().max()) == 0.0, k + " must be exactly zero"
            continue
        e = rel_err(got[k], t)
        e_ref = rel_err(ref32[k], t) if ref32 is not None and k in ref32 else 0.0
        table.append("%-44s %10.2e %10.2e %10.2e" % (k, e, e_ref, float(t.abs().max())))
        lim = max(bar, slack * e_ref)
        if per_ray_outliers and k.startswith("rays.") and t.dim() == 2 and t.shape[0] >= 1024:
            d = (got[k].detach().double().cpu() - t.double()).abs().max(-1).values / float(t.abs().max())
            q = float(d.quantile(0.995))
            nrm = float((got[k].detach().double().cpu() - t.double()).norm() / t.double().norm())
            table[-1] += "   per-ray 99.5%% quantile %.2e, norm-wise %.2e, rays above the bar: %d of %d" % (
                q, nrm, int((d > lim).sum()), d.numel())
            if not (q <= lim and nrm <= lim and e <= 5 * lim):
                bad.append(table[-1])
        elif not e <= lim:
            bad.append(table[-1])
    print("\n".join(table))
    dump_table(name, table)
    assert not bad, "; ".join(bad)


def _oracle_runs(prob, S, chunk=None, **kw):
    """fp64 and fp32 oracle runs (outputs + leaf gradients).  With ``chunk`` the rays are processed in slices and the
    loss is assembled as the same global mean, so memory stays bounded at BASELINE sizes."""
    from oracle import restated as O
    runs = {}
    N = prob["rays"]["rays_d"].shape[0]
    for dt in (torch.float64, torch.float32):
        p = O.to_dtype(prob, dt)
        leaves = O.require_grads(p)
        step = chunk or N
        outs = []
        for i in range(0, N, step):
            pc = dict(p)
            pc["rays"] = {k: v[i:i + step] for k, v in p["rays"].items()}
            kw_c = {k: (v[i:i + step].to(dt) if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == N else v) for k, v in kw.items()}
            r = O.render_rays(pc, n_samples=S, **kw_c)
            loss = ((r["img_coarse"] - 0.3) ** 2).sum() / (3 * N) + ((r["sil_coarse"] - 0.5) ** 2).sum() / N \
                + r["frame_cyc_dis"].sum() / N
            loss.backward()
            outs.append({k: v.detach() for k, v in r.items()})
        res = {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}
        grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
        runs[dt] = (res, grads)
    return runs[torch.float64], runs[torch.float32]


def test_tensor_core_mode_gradients_scale_relative_n128():
    """the fp16 mode's gradients against the fp64 oracle with a PER-TENSOR relative bar (the
    earlier 1e-3 absolute bar was vacuous for the 24 tensors whose |g|max is below 1e-3), skin_aux included."""
    from moda_b200 import synth, models as MM
    from moda_b200.rendering import render_rays
    N, S = 128, 128
    prob = synth.make_problem(N, seed=4)
    (res_o, g64), (res_32, g32) = _oracle_runs(prob, S, perturb=0.0)
    models, emb, rays = MM.build_models(prob, DEV)
    res = render_rays(models, emb, rays, N_samples=S, perturb=0, noise_std=0, opts=synth.default_opts(), img_size=512)
    _loss(res).backward()
    for k in OUT_KEYS:
        assert max_abs(res[k], res_o[k]) < 1e-3, (k, max_abs(res[k], res_o[k]))
    for k in ("xyz_camera_vis", "xyz_canonical_vis"):
        assert rel_err(res[k], res_o[k]) <= max(1e-5, 3 * rel_err(res_32[k], res_o[k])), k
    got = _gpu_grads(models, rays)
    got["skin_aux"], g64["skin_aux"], g32["skin_aux"] = got["skin_aux"][:1], g64["skin_aux"][:1], g32["skin_aux"][:1]
    _check_grads("r02_grad_table_n128_fp16_vs_fp64", got, g64, g32)


def test_full_size_8192x128_against_oracle():
    """BASELINE configs[1] size (8192 rays x 128 samples, 25 bones) through the default fp16 path against the CPU
    oracle in fp64 (truth) and fp32 (the reference's own arithmetic): 57 tile waves on 148 SMs, ragged nothing,
    two-stream weight gradients -- everything the 128-ray tests cannot exercise.  ~40 s of CPU."""
    from moda_b200 import synth, models as MM
    from moda_b200.rendering import render_rays
    N, S = 8192, 128
    prob = synth.make_problem(N, seed=11)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    (res_o, g64), (res_32, g32) = _oracle_runs(prob, S, chunk=1024, perturb=0.0)
    models, emb, rays = MM.build_models(prob, DEV)
    res = render_rays(models, emb, rays, N_samples=S, perturb=0, noise_std=0, opts=synth.default_opts(), img_size=512)
    _loss(res).backward()
    torch.cuda.synchronize()
    tab = []
    for k in OUT_KEYS:
        e = max_abs(res[k], res_o[k])
        tab.append("%-20s max abs err %.2e (fp32 oracle %.2e)" % (k, e, max_abs(res_32[k], res_o[k])))
        assert e < 1e-3, tab[-1]
    for k in ("xyz_camera_vis", "xyz_canonical_vis"):
        e, e32 = rel_err(res[k], res_o[k]), rel_err(res_32[k], res_o[k])
        tab.append("%-20s rel err %.2e (fp32 oracle %.2e)" % (k, e, e32))
        assert e <= max(1e-5, 3 * e32), tab[-1]
    print("\n".join(tab))
    dump_table("r02_outputs_n8192_fp16_vs_fp64", tab)
    got = _gpu_grads(models, rays)
    got["skin_aux"], g64["skin_aux"], g32["skin_aux"] = got["skin_aux"][:1], g64["skin_aux"][:1], g32["skin_aux"][:1]
    _check_grads("r02_grad_table_n8192_fp16_vs_fp64", got, g64, g32, per_ray_outliers=True)


def test_density_grid_fp16_chain_against_golden_and_oracle():
    """the sigma-only chain program that `bench.py --workload grid` times, against (i) the real
    reference's grid (tests/golden, G=12) and (ii) the fp64 oracle at G=64 (262 144 points = 2048 tiles, 14 waves).
    fp16 operands through 8 layers and a x6 sigma head: stated tolerance 1e-2 absolute on raw sigma (|sigma| ~ 3);
    the rendered-output bar (1e-3) is checked on composited values in the render tests."""
    from moda_b200.extract import density_grid
    from moda_b200.nerf import Embedding, NeRF
    from oracle import restated as O
    g = load_npz("geometry_fp32.npz")
    nets = load_npz("nets_seed0.npz")
    sd = {k[len("coarse."):]: torch.from_numpy(v) for k, v in nets.items() if k.startswith("coarse.")}
    coarse = NeRF(in_channels_xyz=63, in_channels_dir=27 + 64, init_beta=0.1)
    coarse.load_state_dict(sd)
    coarse = coarse.to(DEV)
    emb = Embedding(3, 10, alpha=10)
    from moda_b200 import _lib
    n0 = _lib.LAUNCHES
    vol = density_grid(coarse, 12, (0.3, 0.3, 0.3), embedding_xyz=emb)
    assert _lib.LAUNCHES > n0
    e12 = max_abs(vol, g["grid.sigma"])
    sd64 = {k: v.double() for k, v in sd.items()}
    ref = O.density_grid(sd64, 64, (0.3, 0.3, 0.3))
    ref32 = O.density_grid(sd, 64, (0.3, 0.3, 0.3))
    vol64 = density_grid(coarse, 64, (0.3, 0.3, 0.3), embedding_xyz=emb)
    e64 = max_abs(vol64, ref)
    tab = ["G=12 vs reference golden: max abs %.3e (|sigma|max %.2f)" % (e12, float(np.abs(g["grid.sigma"]).max())),
           "G=64 vs fp64 oracle:      max abs %.3e, rms %.3e (fp32 oracle: %.3e; |sigma|max %.2f)"
           % (e64, float((vol64.double().cpu() - ref).pow(2).mean().sqrt()), max_abs(ref32, ref), float(ref.abs().max()))]
    print("\n".join(tab))
    dump_table("r02_density_grid_fp16_vs_oracle", tab)
    assert e12 < 1e-2 and e64 < 1e-2, tab


def _models_from(prob, extra=()):
    from moda_b200 import models as MM
    from moda_b200.nerf import NeRF
    models, emb, rays = MM.build_models(prob, DEV)
    for name in extra:
        if name == "nerf_vis":
            m = NeRF(in_channels_xyz=63, D=5, W=64, out_channels=1, in_channels_dir=0, raw_feat=True)
        else:
            m = NeRF(in_channels_xyz=63, D=5, W=128, out_channels=16, in_channels_dir=0, raw_feat=True, init_beta=1.)
        m.load_state_dict(prob[name])
        models[name] = m.to(DEV)
    return models, emb, rays


@pytest.mark.parametrize("mode", ["fp16", "fp32"])
def test_use_disp_pe_window_and_noise_against_reference_golden(mode):
    """rendering.py:72 (disparity sampling), nerf.py:63-69 with alpha = 6.4 (annealed PE window, inside the chain
    kernel's PE producer in fp16 mode) and rendering.py:193-196 (noise_std > 0, the reference's own draw replayed)."""
    from moda_b200 import config, synth
    from moda_b200.nerf import Embedding
    from moda_b200.rendering import render_rays
    config.set_precision(mode)
    prob, g = fixture_problem("render_opts_n16_fp32.npz")
    models, _, rays = _models_from(prob)
    emb = {"xyz": Embedding(3, 10, alpha=6.4), "dir": Embedding(3, 4, alpha=6.4)}
    with ReplayRng(g, DEV) as tape:
        res = render_rays(models, emb, rays, N_samples=128, use_disp=True, perturb=0, noise_std=0.3, chunk=32768,
                          img_size=512, opts=synth.default_opts())
    assert tape.i == len(tape.draws) == 1
    _loss(res).backward()
    for k in OUT_KEYS:
        assert max_abs(res[k], g["out." + k]) < 1e-3, (k, max_abs(res[k], g["out." + k]))
    for k in ("xyz_camera_vis", "xyz_canonical_vis"):
        assert rel_err(res[k], g["out." + k]) < 3e-5, k
    got = _gpu_grads(models, rays)
    truth = {k[5:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("grad.")}
    # only an fp32 fixture here: the ill-conditioned tensors get the slack the fp32 reference itself needs against
    # its fp64 run (measured on render_n32: <= 1.5e-1), the others the mode's bar
    loose = ("nerf_skin.", "skin_aux", "bones_rst", "rays.", "rest_pose_code")
    well = {k: v for k, v in truth.items() if not k.startswith(loose)}
    ill = {k: v for k, v in truth.items() if k.startswith(loose)}
    ill["skin_aux"], got["skin_aux"] = ill["skin_aux"][:1], got["skin_aux"][:1]
    _check_grads("r02_grad_table_opts_%s_well" % mode, got, well, None, bar=GRAD_BAR if mode == "fp16" else 1e-3)
    _check_grads("r02_grad_table_opts_%s_ill" % mode, got, ill, None, bar=1.5e-1)


@pytest.mark.parametrize("mode", ["fp16", "fp32"])
def test_render_vis_and_obj_bound_masks_against_reference_golden(mode):
    """rendering.py:210-215, 373-379: alphas of out-of-bound samples and of samples nerf_vis calls invisible are
    zeroed; result['vis_pred'] = sum_s vis_pred w (:404).  55 % of the fixture's samples are masked."""
    from moda_b200 import config, synth
    from moda_b200.rendering import render_rays
    config.set_precision(mode)
    prob, g = fixture_problem("render_vis_n16_fp32.npz", nets=("coarse", "nerf_skin", "nerf_vis"))
    models, emb, rays = _models_from(prob, extra=("nerf_vis",))
    for m in ("coarse", "nerf_skin", "nerf_vis"):
        models[m].eval()
    with torch.no_grad():
        res = render_rays(models, emb, rays, N_samples=128, perturb=0, noise_std=0, chunk=32768,
                          obj_bound=prob["obj_bound"].numpy(), img_size=512, opts=synth.default_opts(), render_vis=True)
    assert float(g["masked_frac"]) > 0.3
    assert set(res) == {k[4:] for k in g if k.startswith("out.")}
    for k in OUT_KEYS + ("vis_pred",):
        assert max_abs(res[k], g["out." + k]) < 1e-3, (k, max_abs(res[k], g["out." + k]))


def _two_rank_worker(rank, world, port, prob, q):
    import torch.distributed as dist
    from moda_b200 import synth, models as MM
    from moda_b200.parallel import FlatParams, shard_rays
    from moda_b200.rendering import render_rays
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    dev = "cuda:%d" % rank
    models, emb, rays = MM.build_models(prob, dev)
    flat = FlatParams(MM.parameters_of(models))
    N = rays["rays_d"].shape[0]
    mine = shard_rays(rays, rank, world)
    res = render_rays(models, emb, mine, N_samples=128, perturb=0, noise_std=0, opts=synth.default_opts(), img_size=512)
    n = mine["rays_d"].shape[0]
    (_loss(res) * (n / N)).backward()
    flat.allreduce()
    torch.cuda.synchronize()
    if rank == 0:
        q.put(flat.grad.cpu())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_nccl_gradients_equal_single_rank():
    """SURVEY 8(e) on hardware: the ray batch sharded over 2 GPUs + ONE ncclAllReduce of the flat gradient buffer
    gives the single-GPU gradient (each rank scales its loss by its share)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from moda_b200 import synth, models as MM
    from moda_b200.parallel import FlatParams
    from moda_b200.rendering import render_rays
    N = 512
    prob = synth.make_problem(N, seed=6)
    models, emb, rays = MM.build_models(prob, DEV)
    flat = FlatParams(MM.parameters_of(models))
    res = render_rays(models, emb, rays, N_samples=128, perturb=0, noise_std=0, opts=synth.default_opts(), img_size=512)
    _loss(res).backward()
    single = flat.grad.cpu()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_two_rank_worker, args=(r, 2, port, prob, q)) for r in range(2)]
    for p in procs:
        p.start()
    both = q.get(timeout=600)
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    # same kernels, different summation order across the two halves (atomics + the all-reduce): rounding only
    e = float((both - single).abs().max() / single.abs().max())
    print("2-rank vs 1-rank flat gradient: max rel diff %.2e over %d values" % (e, single.numel()))
    dump_table("r02_two_rank_vs_single_rank", ["max |g2 - g1| / max |g1| = %.3e over %d values" % (e, single.numel())])
    assert e < 1e-3


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_default_flag_training_step_against_reference_golden(mode):
    """SURVEY 8(f) rank 1 / VERDICT r1 missing #1-#2: a ``models`` dict built as nnutils/moda.py:271-348, 447-449 builds
    it (coarse, bones, skin_aux, nerf_skin, rest_pose_code, nerf_vis, nerf_feat) and the ``rays`` of a paired-frame
    training batch (rtk_vec[_target], bone_rts_target, feats_at_samp, *_at_samp) run through render_rays at the
    reference's default flags (dist_corresp, use_corresp, use_ot).  Every result key and every gradient against the
    real reference (fp64 run as truth, fp32 run for the slack), in-call random draws replayed."""
    from moda_b200 import config, synth, models as MM
    from moda_b200.rendering import render_rays
    from oracle import restated as O
    config.set_precision(mode)
    nets = ("coarse", "nerf_skin", "nerf_vis", "nerf_feat")
    prob, g = fixture_problem("render_full_n16_fp32.npz", nets=nets)
    g64 = load_npz("render_full_n16_fp64.npz")
    models, emb, rays = MM.build_full_models(prob, DEV)
    for m in nets:
        models[m].train()
    with ReplayRng(g, DEV) as tape:
        res = render_rays(models, emb, rays, N_samples=128, perturb=0, noise_std=0, chunk=32768,
                          obj_bound=prob["obj_bound"].numpy(), img_size=512, opts=synth.full_opts())
    assert tape.i == len(tape.draws) == 3, "same random draws, in the reference's order"
    assert set(res) == {k[4:] for k in g if k.startswith("out.")} - {"loss"}
    loss = O.full_loss(res)
    loss.backward()
    tab = []
    for k in sorted(res):
        ref = np.asarray(g64["out." + k], dtype=np.float64).reshape(tuple(res[k].shape))
        e = max_abs(res[k].double() if res[k].dtype != torch.bool else res[k].double(), ref)
        scale = float(np.abs(ref).max())
        tab.append("%-20s max abs err %.2e (|ref|max %.2e)" % (k, e, scale))
        # rendered values 1e-3 absolute; pixel-unit quantities (flow / reprojection in normalised image units of
        # O(1), feature-matched points) relative to their scale
        assert e <= 1e-3 * max(1.0, scale), tab[-1]
    print("\n".join(tab))
    dump_table("r02_full_step_outputs_%s" % mode, tab)
    assert abs(float(loss.detach()) - float(g64["out.loss"])) < 2e-3
    got = _gpu_grads(models, rays, nets=nets)
    truth = {k[5:]: torch.from_numpy(g64[k]) for k in g64 if k.startswith("grad.")}
    ref32 = {k[5:]: torch.from_numpy(g[k]) for k in g if k.startswith("grad.")}
    got["skin_aux"], truth["skin_aux"], ref32["skin_aux"] = got["skin_aux"][:1], truth["skin_aux"][:1], ref32["skin_aux"][:1]
    _check_grads("r02_full_step_grads_%s" % mode, got, truth, ref32, bar=GRAD_BAR if mode == "fp16" else 1e-3)


def test_generic_tensor_core_mlp_matches_fp32_simt():
    """moda_b200/generic_tc.py (nerf_feat 5x128 -> 16, nerf_vis 5x64 -> 1 on tcgen05, layer by layer) against the exact
    fp32 SIMT path: values and every gradient at fp16-operand level.  Sizes include a ragged last tile."""
    from moda_b200 import config, geom_utils as G, synth
    from moda_b200.nerf import Embedding, NeRF
    prob = synth.make_full_problem(4, seed=1)
    emb = Embedding(3, 10, alpha=10)
    nrel = lambda x, y: float((x.double() - y.double()).norm() / (y.double().norm() + 1e-30))
    for name, kw in (("nerf_feat", dict(D=5, W=128, out_channels=16, init_beta=1.)), ("nerf_vis", dict(D=5, W=64, out_channels=1))):
        model = NeRF(in_channels_xyz=63, in_channels_dir=0, raw_feat=True, **kw)
        model.load_state_dict(prob[name])
        model = model.to(DEV)
        oc = kw["out_channels"]
        for R, S in ((5, 128), (301, 100)):
            gen = torch.Generator().manual_seed(3)
            pts0 = (torch.rand(R, S, 3, generator=gen) * 0.6 - 0.3).to(DEV)
            gout = torch.randn(R, S, oc, generator=gen).to(DEV) * 1e-3
            outs = {}
            for mode in ("fp32", "fp16"):
                config.set_precision(mode)
                model.zero_grad()
                p = pts0.clone().requires_grad_(True)
                out = G.evaluate_mlp(model, p, embed_xyz=emb)
                assert out.shape == (R, S, oc)
                (out * gout).sum().backward()
                outs[mode] = (out.detach(), p.grad, {k: v.grad.clone() for k, v in model.named_parameters() if v.grad is not None})
                with torch.no_grad():
                    assert torch.equal(G.evaluate_mlp(model, pts0, embed_xyz=emb), out.detach()) or mode == "fp32"
            a, b = outs["fp32"], outs["fp16"]
            scale = float(a[0].abs().max())
            assert max_abs(b[0], a[0]) < 5e-3 * max(scale, 1.0), (name, R, max_abs(b[0], a[0]), scale)
            assert nrel(b[1], a[1]) < 6e-2, (name, "gpts", nrel(b[1], a[1]))
            assert set(a[2]) == set(b[2]), (name, set(a[2]) ^ set(b[2]))
            worst = max((nrel(b[2][k], a[2][k]), k) for k in a[2])
            assert worst[0] < 6e-2, (name, R, worst)
