"""TEST INFRASTRUCTURE ONLY -- fixtures for the caller side of the renderer, produced
by executing the REAL reference (see oracle/make_golden2.py, which calls golden_callers):
  raycast / sample_xy                     geom_utils.py:746-827
  DQ_RTHead / FrameCode                   nerf.py:239-279, 346-380
  correct_bones / correct_rest_pose       geom_utils.py:933-972
  warp_fw / warp_bw (with and without the nerf_dis residual field)   geom_utils.py:974-1073, 350-456
  obj_to_cam / pinhole_cam / vrender_flo  geom_utils.py:567-672, 1704-1743
  render_rays with opts.symm_shape        rendering.py:385-391
"""
import os
import types

import numpy as np
import torch
from torch import nn

from moda_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
B = synth.NUM_BONES


def _sd(prefix, module):
    return {prefix + k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def build_stub_model(ref, gen_seed=21, with_dis=False):
    """The attributes of ``nnutils.moda.moda`` that warp_fw / warp_bw / correct_bones read, built with the reference's
    own classes (moda.py:271-348)."""
    torch.manual_seed(gen_seed)
    m = types.SimpleNamespace()
    m.device = torch.device("cpu")
    m.opts = types.SimpleNamespace(num_bones=B, flowbw=False, lbs=False, neudbs=True, nerf_skin=True, nerf_dis=with_dis)
    prob = synth.make_problem(4, seed=gen_seed)
    m.bones = prob["bones_rst"].clone()
    m.skin_aux = prob["skin_aux"].clone()
    m.embedding_xyz = ref.Embedding(3, 10, alpha=10)
    m.nerf_skin = ref.NeRF(in_channels_xyz=63 + 128, D=5, W=64, in_channels_dir=0, out_channels=B, raw_feat=True,
                           in_channels_code=128)
    m.nerf_skin.load_state_dict(prob["nerf_skin"])
    m.rest_pose_code = nn.Embedding(1, 128)
    m.pose_code = nn.Embedding(6, 128)
    with torch.no_grad():
        m.pose_code.weight.mul_(0.1)
        m.rest_pose_code.weight.mul_(0.1)
    head = ref.nerf.DQ_RTHead(use_quat=True, in_channels_xyz=128, in_channels_dir=0, out_channels=7 * B, raw_feat=True)
    with torch.no_grad():   # the zero-initialised biases give the identity rotation nowhere: break the symmetry
        head.rgb[0].bias.copy_(0.3 * torch.randn(7 * B))
        head.rgb[0].bias[3::7] += 1.0
    m.nerf_body_rts = nn.Sequential(m.pose_code, head)
    if with_dis:
        m.nerf_dis = ref.NeRF(in_channels_xyz=63 + 128, D=5, W=128, in_channels_dir=0, out_channels=3, raw_feat=True,
                              in_channels_code=128)
        with torch.no_grad():
            m.nerf_dis.rgb[0].weight.mul_(0.05)
            m.nerf_dis.rgb[0].bias.mul_(0.05)
    return m


def golden_callers(ref, name):
    G = ref.geom_utils
    out = {}
    gen = torch.Generator().manual_seed(99)
    # ---- raycast / sample_xy
    bs, ns = 2, 48
    xys = torch.rand(bs, ns, 2, generator=gen) * 511
    Rm = synth._small_rotation(gen, bs, 0.1)
    Tm = torch.tensor([[0.0, 0.0, 0.3]]) + 0.02 * torch.randn(bs, 3, generator=gen)
    K = torch.tensor([[600.0, 600.0, 256.0, 256.0], [620.0, 610.0, 250.0, 260.0]])
    Kinv = G.K2inv(K)
    near_far = torch.tensor([[0.1, 0.5], [0.15, 0.6]])
    rays = G.raycast(xys, Rm, Tm, Kinv, near_far)
    out.update({"raycast.xys": xys.numpy(), "raycast.Rmat": Rm.numpy(), "raycast.Tmat": Tm.numpy(), "raycast.Kinv": Kinv.numpy(),
                "raycast.near_far": near_far.numpy()})
    for k in ("rays_o", "rays_d", "near", "far", "rtk_vec", "xys"):
        out["raycast.out." + k] = rays[k].numpy()
    rays_nf = G.raycast(xys, Rm, Tm, Kinv, None)
    out["raycast.nonf.near"], out["raycast.nonf.far"] = rays_nf["near"].numpy(), rays_nf["far"].numpy()
    ri, xy_all = G.sample_xy(6, 2, 0, "cpu", return_all=True)
    out["sample_xy.all.rand_inds"], out["sample_xy.all.xys"] = ri.numpy(), xy_all.numpy()
    # ---- camera algebra + flow rendering
    pts = torch.randn(5, 7, 3, generator=gen) * 0.1
    Rm5 = synth._small_rotation(gen, 5, 0.2)
    Tm5 = torch.tensor([[0.0, 0.0, 0.3]]) + 0.05 * torch.randn(5, 3, generator=gen)
    K5 = torch.tensor([[600.0, 610.0, 256.0, 250.0]]).repeat(5, 1) + torch.randn(5, 4, generator=gen)
    cam = G.obj_to_cam(pts, Rm5.view(5, 1, 3, 3), Tm5.view(5, 1, 3))
    cam[0, 0, 2] = -0.2      # a point behind the camera and a far-off one: the invalid branch of vrender_flo
    pix = G.pinhole_cam(cam, K5.view(5, 1, 4))
    w = torch.rand(5, 7, generator=gen)
    xys5 = torch.rand(5, 2, generator=gen) * 511
    flo, valid = G.vrender_flo(w, pix, xys5, 512)
    out.update({"cam.pts": pts.numpy(), "cam.Rmat": Rm5.numpy(), "cam.Tmat": Tm5.numpy(), "cam.K": K5.numpy(),
                "cam.obj_to_cam": G.obj_to_cam(pts, Rm5.view(5, 1, 3, 3), Tm5.view(5, 1, 3)).numpy(), "cam.cam_in": cam.numpy(),
                "cam.pinhole": pix.numpy(), "cam.Kmatinv": G.Kmatinv(G.K2mat(K5)).numpy(), "cam.mat2K": G.mat2K(G.K2mat(K5)).numpy(),
                "flo.w": w.numpy(), "flo.xys": xys5.numpy(), "flo.out": flo.numpy(), "flo.valid": valid.numpy()})
    # ---- FrameCode
    vid_offset = np.asarray([0, 10, 25])
    fc = ref.nerf.FrameCode(10, 32, vid_offset)
    fid = torch.tensor([0, 3, 9, 10, 17, 24])
    out.update(_sd("framecode.", fc))
    out["framecode.fid"], out["framecode.out"] = fid.numpy(), fc(fid).detach().numpy()
    # ---- DQ_RTHead, correct_bones, correct_rest_pose, warp_fw / warp_bw
    for tag, with_dis in (("warp", False), ("warpdis", True)):
        m = build_stub_model(ref, with_dis=with_dis)
        head = m.nerf_body_rts[1]
        if not with_dis:   # same seed -> same head / codes / nerf_skin / bones in both variants: stored once
            out.update(_sd("warp.head.", head))
            out.update(_sd("warp.pose_code.", m.pose_code))
            out.update(_sd("warp.rest_pose_code.", m.rest_pose_code))
            out.update(_sd("warp.nerf_skin.", m.nerf_skin))
            out["warp.bones"], out["warp.skin_aux"] = m.bones.numpy(), m.skin_aux.numpy()
        else:
            out.update(_sd(tag + ".nerf_dis.", m.nerf_dis))
        with torch.no_grad():
            codes = m.pose_code(torch.arange(6))
            out[tag + ".head.in"], out[tag + ".head.out"] = codes.numpy(), head(codes).numpy()
            bones_rst, rts_rst = G.correct_bones(m, m.bones, neudbs=True)
            bones_rst_inv, rts_rst_inv = G.correct_bones(m, m.bones, inverse=True, neudbs=True)
            rts_fw = m.nerf_body_rts(torch.tensor([[2], [4], [5]]))
            delta = G.correct_rest_pose(m.opts, rts_fw, rts_rst, True)
            out.update({tag + ".correct_bones.bones": bones_rst.numpy(), tag + ".correct_bones.rts": rts_rst.numpy(),
                        tag + ".correct_bones.inv.bones": bones_rst_inv.numpy(), tag + ".correct_bones.inv.rts": rts_rst_inv.numpy(),
                        tag + ".correct_rest_pose.in": rts_fw.numpy(), tag + ".correct_rest_pose.out": delta.numpy()})
            verts = (torch.rand(300, 3, generator=gen) * 0.4 - 0.2)
            vf, rt = G.warp_fw(m.opts, m, {}, verts.numpy(), 3)
            out[tag + ".verts"], out[tag + ".fw"], out[tag + ".fw.bones"] = verts.numpy(), vf, rt["bones"].numpy()
            vb, rt = G.warp_bw(m.opts, m, {}, torch.Tensor(vf).clone(), 3)
            out[tag + ".bw"], out[tag + ".bw.bones"] = vb.numpy(), rt["bones"].numpy()
    # ---- symm_shape through render_rays (rendering.py:385-391)
    from oracle.make_golden import build_reference_models
    from oracle.make_golden2 import RngTape
    prob = synth.make_problem(8, seed=9)
    models, emb = build_reference_models(ref, prob)
    for mm in (models["coarse"], models["nerf_skin"]):
        mm.eval()
    opts = synth.default_opts()
    opts.symm_shape = True
    rays_s = {k: v.clone() for k, v in prob["rays"].items()}
    with torch.no_grad(), RngTape() as tape:
        res = ref.render_rays(models, emb, rays_s, N_samples=64, perturb=0, noise_std=0, chunk=32768, img_size=512, opts=opts)
    for k, v in prob["rays"].items():
        out["symm.in.rays." + k] = v.numpy()
    for k in ("bones_rst", "skin_aux", "rest_pose_code"):
        out["symm.in." + k] = prob[k].numpy()
    for net in ("coarse", "nerf_skin"):
        for k, v in prob[net].items():
            out["symm.net.%s.%s" % (net, k)] = v.numpy()
    for k in ("img_coarse", "sil_coarse", "depth_rnd", "frame_cyc_dis"):
        out["symm.out." + k] = res[k].numpy()
    for i, t in enumerate(tape.draws):
        out["symm.rng.%d" % i] = t.numpy()
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, "written;", len(out), "arrays; symm draws", [tuple(t.shape) for t in tape.draws])
