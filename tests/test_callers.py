"""Caller side of the renderer against fixtures produced by the real reference
(oracle/make_golden_callers.py).  The per-ray camera / ray algebra is device-agnostic tensor code and is checked here
on the CPU; everything that touches the CUDA kernels (pose head, rest-pose correction, mesh-extraction warps with and
without the nerf_dis residual field, symm_shape) is in the ``gpu`` tests below."""
import types

import numpy as np
import pytest
import torch
from torch import nn

from tests.util import ReplayRng, load_npz, max_abs, rel_err

T = lambda a: torch.from_numpy(np.asarray(a))


def test_raycast_sample_xy_and_camera_algebra_against_reference():
    from moda_b200 import geom_utils as G
    g = load_npz("callers_fp32.npz")
    rays = G.raycast(T(g["raycast.xys"]), T(g["raycast.Rmat"]), T(g["raycast.Tmat"]), T(g["raycast.Kinv"]), T(g["raycast.near_far"]))
    for k in ("rays_o", "rays_d", "near", "far", "rtk_vec", "xys"):
        assert max_abs(rays[k], g["raycast.out." + k]) < 1e-6, k
    assert rays["bs"] == 2 and rays["nsample"] == 48
    nonf = G.raycast(T(g["raycast.xys"]), T(g["raycast.Rmat"]), T(g["raycast.Tmat"]), T(g["raycast.Kinv"]), None)
    assert max_abs(nonf["near"], g["raycast.nonf.near"]) < 1e-6 and max_abs(nonf["far"], g["raycast.nonf.far"]) < 1e-6
    ri, xy = G.sample_xy(6, 2, 0, "cpu", return_all=True)
    assert torch.equal(ri, T(g["sample_xy.all.rand_inds"])) and max_abs(xy, g["sample_xy.all.xys"]) == 0
    ri, xy = G.sample_xy(16, 3, 20, "cpu")    # random draw: a permutation sample of distinct pixels inside the image
    assert ri.shape == (3, 20) and len(set(ri.reshape(-1).tolist())) == 60 and float(xy.max()) <= 15
    assert torch.equal(xy[..., 0].long() + 16 * xy[..., 1].long(), ri)
    chunk = G.chunk_rays({"a": torch.arange(24.).reshape(2, 4, 3), "n": 4}, 2, 3)
    assert list(chunk) == ["a"] and chunk["a"].shape == (3, 3) and float(chunk["a"][0, 0]) == 6.0
    pts, Rm, Tm, K = T(g["cam.pts"]), T(g["cam.Rmat"]), T(g["cam.Tmat"]), T(g["cam.K"])
    assert max_abs(G.obj_to_cam(pts, Rm.view(5, 1, 3, 3), Tm.view(5, 1, 3)), g["cam.obj_to_cam"]) < 1e-6
    pix = G.pinhole_cam(T(g["cam.cam_in"]), K.view(5, 1, 4))
    assert rel_err(pix, g["cam.pinhole"]) < 1e-6
    assert rel_err(G.Kmatinv(G.K2mat(K)), g["cam.Kmatinv"]) < 1e-6 and max_abs(G.mat2K(G.K2mat(K)), g["cam.mat2K"]) == 0
    flo, valid = G.vrender_flo(T(g["flo.w"]), T(g["cam.pinhole"]), T(g["flo.xys"]), 512)
    assert max_abs(flo, g["flo.out"]) < 1e-5 and max_abs(valid, g["flo.valid"]) == 0 and float(valid.min()) == 0.0
    assert max_abs(G.diff_flo(T(g["flo.xys"])[:, None] + 3.0, T(g["flo.xys"]), 512), np.full((5, 2), 3.0 / 512 * 2)) < 1e-6


def _stub_model(g, dev, with_dis):
    from moda_b200.nerf import DQ_RTHead, Embedding, NeRF
    B = 25
    sd = lambda pre: {k[len(pre):]: T(v) for k, v in g.items() if k.startswith(pre) and k[len(pre):] not in ("in", "out")}
    m = types.SimpleNamespace()
    m.device = torch.device(dev)
    m.opts = types.SimpleNamespace(num_bones=B, flowbw=False, lbs=False, neudbs=True, nerf_skin=True, nerf_dis=with_dis)
    m.bones, m.skin_aux = T(g["warp.bones"]).to(dev), T(g["warp.skin_aux"]).to(dev)
    m.embedding_xyz = Embedding(3, 10, alpha=10)
    m.nerf_skin = NeRF(in_channels_xyz=63 + 128, D=5, W=64, in_channels_dir=0, out_channels=B, raw_feat=True, in_channels_code=128)
    m.nerf_skin.load_state_dict(sd("warp.nerf_skin."))
    m.rest_pose_code, m.pose_code = nn.Embedding(1, 128), nn.Embedding(6, 128)
    m.rest_pose_code.load_state_dict(sd("warp.rest_pose_code."))
    m.pose_code.load_state_dict(sd("warp.pose_code."))
    head = DQ_RTHead(use_quat=True, in_channels_xyz=128, in_channels_dir=0, out_channels=7 * B, raw_feat=True)
    head.load_state_dict(sd("warp.head."))
    m.nerf_body_rts = nn.Sequential(m.pose_code, head).to(dev)
    m.nerf_skin, m.rest_pose_code = m.nerf_skin.to(dev), m.rest_pose_code.to(dev)
    if with_dis:
        m.nerf_dis = NeRF(in_channels_xyz=63 + 128, D=5, W=128, in_channels_dir=0, out_channels=3, raw_feat=True, in_channels_code=128)
        m.nerf_dis.load_state_dict(sd("warpdis.nerf_dis."))
        m.nerf_dis = m.nerf_dis.to(dev)
    return m


def test_project_render_flo_tensor_path_against_oracle():
    """project_render_flo (rendering.py:434-459 + 480-499 for one paired frame) on CPU tensors -- the tensor-op path the
    fused CUDA kernels are tested against on the GPU -- against the oracle's restatement of obj_to_cam / pinhole_cam /
    vrender_flo, values and gradients, with samples behind the camera and outside the image in the batch."""
    from moda_b200 import geom_utils as G
    from oracle import restated as O
    gen = torch.Generator().manual_seed(21)
    N, S, img = 9, 17, 64.0
    xyz0 = torch.randn(N, S, 3, generator=gen, dtype=torch.float64) * 0.3
    xyz0[:, :, 2] += 3.0
    xyz0[2, 4, 2] = -2.0
    xyz0[5, :, 1] += 30.0
    w0 = torch.rand(N, S, generator=gen, dtype=torch.float64)
    Rm = torch.linalg.qr(torch.randn(N, 3, 3, generator=gen, dtype=torch.float64))[0]
    Tm = torch.randn(N, 3, generator=gen, dtype=torch.float64) * 0.1
    Kinv = torch.zeros(N, 3, 3, dtype=torch.float64)
    Kinv[:, 0, 0], Kinv[:, 1, 1], Kinv[:, 2, 2], Kinv[:, 0, 2], Kinv[:, 1, 2] = 1 / 60.0, 1 / 62.0, 1.0, -32 / 60.0, -30 / 62.0
    rtk0 = torch.cat([Rm.reshape(N, 9), Tm, Kinv.reshape(N, 9)], -1)
    xys = torch.rand(N, 2, generator=gen, dtype=torch.float64) * img
    gout = torch.randn(N, 2, generator=gen, dtype=torch.float64)
    outs = []
    for fn in (lambda w, x, r: G.project_render_flo(w, x, r, xys, img, N),
               lambda w, x, r: O.vrender_flo(w, O.project(x, r), xys, img)):
        xyz, w, rtk = (t.clone().requires_grad_(True) for t in (xyz0, w0, rtk0))
        flo, valid = fn(w, xyz, rtk)
        (flo * gout).sum().backward()
        outs.append((flo.detach(), valid.detach().reshape(-1), xyz.grad, w.grad, rtk.grad))
    assert float(outs[0][1].sum()) < N
    for a, b in zip(*outs):
        assert max_abs(a, b.numpy()) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_pose_head_rest_pose_correction_and_mesh_warps_against_reference(mode):
    """DQ_RTHead (nerf.py:239-279), FrameCode (:346-380), correct_bones / correct_rest_pose (geom_utils.py:933-972),
    warp_fw / warp_bw (:974-1073) with the reference's signatures, with and without nerf_dis (:350-456)."""
    from moda_b200 import config, geom_utils as G
    from moda_b200.nerf import FrameCode
    dev = "cuda"
    old = config.precision
    config.set_precision(mode)
    try:
        g = load_npz("callers_fp32.npz")
        fc = FrameCode(10, 32, np.asarray([0, 10, 25]))
        fc.load_state_dict({k[len("framecode."):]: T(v) for k, v in g.items() if k.startswith("framecode.") and
                            k not in ("framecode.fid", "framecode.out")})
        assert max_abs(fc.to(dev)(T(g["framecode.fid"]).to(dev)), g["framecode.out"]) < 5e-5   # PE arguments reach 2^9 rad: fp32 sin differs by ~1e-5 between libm and the device
        for tag, with_dis in (("warp", False), ("warpdis", True)):
            m = _stub_model(g, dev, with_dis)
            with torch.no_grad():
                out = m.nerf_body_rts[1](T(g[tag + ".head.in"]).to(dev))
                assert out.shape == (6, 1, 200) and max_abs(out, g[tag + ".head.out"]) < 2e-5
                bones_rst, rts_rst = G.correct_bones(m, m.bones, neudbs=True)
                assert max_abs(bones_rst, g[tag + ".correct_bones.bones"]) < 2e-5
                assert max_abs(rts_rst, g[tag + ".correct_bones.rts"]) < 2e-5
                bi, ri = G.correct_bones(m, m.bones, inverse=True, neudbs=True)
                assert max_abs(bi, g[tag + ".correct_bones.inv.bones"]) < 2e-5 and max_abs(ri, g[tag + ".correct_bones.inv.rts"]) < 2e-5
                delta = G.correct_rest_pose(m.opts, T(g[tag + ".correct_rest_pose.in"]).to(dev), rts_rst, True)
                assert max_abs(delta, g[tag + ".correct_rest_pose.out"]) < 2e-5
                vf, rt = G.warp_fw(m.opts, m, {}, g[tag + ".verts"], 3)
                assert isinstance(vf, np.ndarray) and rel_err(vf, g[tag + ".fw"]) < 2e-5, rel_err(vf, g[tag + ".fw"])
                assert max_abs(rt["bones"], g[tag + ".fw.bones"]) < 2e-5
                vb, rt = G.warp_bw(m.opts, m, {}, T(g[tag + ".fw"]).to(dev), 3)
                assert rel_err(vb, g[tag + ".bw"]) < 2e-5, rel_err(vb, g[tag + ".bw"])
                assert max_abs(rt["bones"], g[tag + ".bw.bones"]) < 2e-5
    finally:
        config.set_precision(old)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_symm_shape_against_reference(mode):
    """rendering.py:385-391: half of the samples (the reference's own draw, replayed) are evaluated at their x-mirror."""
    from moda_b200 import config, synth, models as MM
    from moda_b200.rendering import render_rays
    dev = "cuda"
    old = config.precision
    config.set_precision(mode)
    try:
        g = load_npz("callers_fp32.npz")
        prob = {"coarse": {}, "nerf_skin": {}, "rays": {}, "num_bones": 25}
        for k, v in g.items():
            if k.startswith("symm.net."):
                _, _, net, key = k.split(".", 3)
                prob[net][key] = T(v)
            elif k.startswith("symm.in.rays."):
                prob["rays"][k[len("symm.in.rays."):]] = T(v)
            elif k.startswith("symm.in."):
                prob[k[len("symm.in."):]] = T(v)
        models, emb, rays = MM.build_models(prob, dev, requires_grad=False)
        for mm in ("coarse", "nerf_skin"):
            models[mm].eval()
        opts = synth.default_opts()
        opts.symm_shape = True
        sub = {k[len("symm."):]: v for k, v in g.items() if k.startswith("symm.rng.")}
        with torch.no_grad(), ReplayRng(sub, dev) as tape:
            res = render_rays(models, emb, rays, N_samples=64, perturb=0, noise_std=0, chunk=32768, img_size=512, opts=opts)
        assert tape.i == 2
        for k in ("img_coarse", "sil_coarse", "depth_rnd", "frame_cyc_dis"):
            assert max_abs(res[k], g["symm.out." + k]) < 1e-3, (k, max_abs(res[k], g["symm.out." + k]))
    finally:
        config.set_precision(old)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_mesh_extraction_volume_with_visibility_pass_and_symmetry(mode):
    """train_utils.py:1398-1425: the volume handed to marching cubes -- |x| evaluation for symmetric shapes and density
    -1 at lattice points nerf_vis calls unobserved -- against the oracle, incl. the x-slab sharding."""
    from moda_b200 import config, synth, models as MM
    from moda_b200.extract import density_grid, grid_to_object
    from oracle import restated as O
    old = config.precision
    config.set_precision(mode)
    try:
        prob = synth.make_full_problem(4, seed=3)
        models, emb, _ = MM.build_full_models(prob, "cuda", requires_grad=False)
        bound, Gs = (0.2, 0.25, 0.3), 20
        ref = O.density_grid(prob["coarse"], Gs, bound, vis_sd=prob["nerf_vis"], symm_shape=True)
        vol = density_grid(models["coarse"], Gs, bound, emb["xyz"], nerf_vis=models["nerf_vis"], symm_shape=True, chunk=3 * Gs * Gs)
        masked = ref == -1
        assert 0.05 < float(masked.float().mean()) < 0.95
        # a lattice point whose visibility sits on the 0.5 threshold may flip: compare where both agree on the mask
        agree = (vol.cpu() == -1) == masked
        assert float(agree.float().mean()) > 0.999
        tol = 5e-6 if mode == "fp32" else 1e-2
        assert max_abs(torch.where(agree, vol.cpu(), ref), ref) < tol
        slabs = torch.cat([density_grid(models["coarse"], Gs, bound, emb["xyz"], nerf_vis=models["nerf_vis"], symm_shape=True,
                                        x_range=(a, b)) for a, b in ((0, 7), (7, 20))], 0)
        assert torch.equal(slabs, vol)
        v = grid_to_object(np.asarray([[0.0, 10.0, 20.0]]), Gs, bound)
        assert np.allclose(v, [[-0.2, 0.0, 0.3]])
    finally:
        config.set_precision(old)
