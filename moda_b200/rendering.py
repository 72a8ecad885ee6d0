"""``render_rays`` with the reference's signature and result dict (nnutils/rendering.py:19-122), executed by
the CUDA kernels of this package: sample generation, nerf_skin delta logits, fused Gaussian-skinning +
dual-quaternion backward / forward warps, the 8x256 trunk, and the compositor with the cycle term.

Implemented: SURVEY.md section 8(a)'s core path -- ``models`` keys {coarse, bones, bones_rst, skin_aux, nerf_skin,
rest_pose_code}, ``rays`` keys {rays_o, rays_d, near, far, xys, time_embedded, bone_rts, env_code} -- and, since round
2, what MoDA's DEFAULT flags add to a training step (section 8(f) rank 1): ``nerf_feat`` feature rendering +
``feat_match_loss`` + key-point reprojection, the third warp to the paired frame with flow rendering
(``rtk_vec_target`` / ``bone_rts_target`` / ``*_dentrg``), ``nerf_vis`` (``render_vis`` masking and ``vis_loss``),
``nerf_dis`` residual fields, ``symm_shape`` and the per-ray loss terms; plus the two alternative motion models of
section 8(f) rank 5: LBS (``opts.lbs``: rigid transforms per bone, linear blend) and the free-form flow fields
(``flowbw`` / ``flowfw``: Transhead or SE3head).  Branches that remain out of scope (nerf_unc, appearance codes, s3im)
raise NotImplementedError instead of silently doing something else.
"""
import torch
import torch.nn.functional as F

from . import geom_utils as G
from . import loss_utils as L
from .ops import CompositeFn, PointsFromDepthsFn, SampleRaysFn, SamplePdfFn, WeightedSumFn

_UNSUPPORTED_MODELS = ("nerf_unc",)
_UNSUPPORTED_RAYS = ("appearance_code",)


def render_rays(models, embeddings, rays, N_samples=64, use_disp=False, perturb=0, noise_std=1, chunk=1024 * 32,
                obj_bound=None, use_fine=False, img_size=None, progress=None, opts=None, render_vis=False):
    """rendering.py:19-122.  Same inputs, same result keys; see the module docstring for the covered flags."""
    for k in _UNSUPPORTED_MODELS:
        if k in models:
            raise NotImplementedError("models['%s'] is outside the accelerated core path (SURVEY.md 8(f))" % k)
    for k in _UNSUPPORTED_RAYS:
        if k in rays:
            raise NotImplementedError("rays['%s'] is outside the accelerated core path (SURVEY.md 8(f))" % k)
    if (opts is not None and "bones" in models and "flowbw" not in models
            and not (getattr(opts, "lbs", False) or getattr(opts, "neudbs", True))):
        raise ValueError("models['bones'] needs a motion model: opts.neudbs or opts.lbs (rendering.py:312-322)")
    if opts is not None and getattr(opts, "s3im_loss", False):
        raise NotImplementedError("opts.s3im_loss is off in every MoDA script (moda.py:170) and not implemented")
    if use_fine:
        N_samples = N_samples // 2
    embedding_xyz, embedding_dir = embeddings["xyz"], embeddings["dir"]
    rays_o, rays_d = rays["rays_o"], rays["rays_d"]
    near, far = rays["near"], rays["far"]
    N_rays = rays_d.shape[0]
    dev = rays_d.device

    jitter = None
    if perturb > 0:  # same draw, same shape as rendering.py:82
        jitter = torch.rand(N_rays, N_samples, device=dev)
    z_vals, xyz_sampled, rays_d_norm = SampleRaysFn.apply(rays_o, rays_d, near, far, jitter, float(perturb),
                                                          bool(use_disp), N_samples)
    dir_embedded = embedding_dir(rays_d_norm)

    if use_fine:
        with torch.no_grad():
            _, weights_coarse = inference_deform(xyz_sampled, rays, models, chunk, N_samples, N_rays, embedding_xyz,
                                                 rays_d, noise_std, obj_bound, dir_embedded, z_vals, img_size,
                                                 progress, opts, fine_iter=False)
        N_importance = N_samples
        u = None if perturb == 0 else torch.rand(N_rays, N_importance, device=dev)
        z_vals = sample_pdf_merge(z_vals, weights_coarse, N_importance, det=(perturb == 0), u=u)
        xyz_sampled = PointsFromDepthsFn.apply(rays_o, rays_d, z_vals)
        N_samples = N_samples + N_importance

    result, _ = inference_deform(xyz_sampled, rays, models, chunk, N_samples, N_rays, embedding_xyz, rays_d,
                                 noise_std, obj_bound, dir_embedded, z_vals, img_size, progress, opts,
                                 render_vis=render_vis)
    return result


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5, u=None):
    """rendering.py:582-623: inverse-CDF sampling of ``N_importance`` depths per ray."""
    if u is None and not det:
        u = torch.rand(weights.shape[0], N_importance, device=weights.device)
    return SamplePdfFn.apply(bins, weights, N_importance, bool(det), float(eps), u, None)


def sample_pdf_merge(z_vals, weights_coarse, N_importance, det, u=None, eps=1e-5):
    """rendering.py:103-110: mid-point bins, sample_pdf on weights[:,1:-1], then the sorted union with z_vals."""
    if u is None and not det:
        u = torch.rand(z_vals.shape[0], N_importance, device=z_vals.device)
    return SamplePdfFn.apply(None, weights_coarse, N_importance, bool(det), float(eps), u, z_vals)


def inference(models, embedding_xyz, xyz_, dir_, dir_embedded, z_vals, N_rays, N_samples, chunk, noise_std,
              env_code=None, appearance_code=None, weights_only=False, clip_bound=None, vis_pred=None,
              scale_rgb=1.3, rgb_filter=False, cyc_pair=None):
    """rendering.py:124-237.  Returns (rgb, feat, depth, weights, visibility, sil[, cyc])."""
    if rgb_filter:
        raise NotImplementedError("rgb_filter is off in every MoDA script")
    if appearance_code is not None:
        raise NotImplementedError("appearance_code is outside the accelerated core path")
    nerf_sdf = models["coarse"]
    xyz_input = xyz_.reshape(N_rays, -1, 3)
    S = xyz_input.shape[1]
    if dir_embedded.dim() == 2 and dir_embedded.shape[0] == N_rays * S and S > 1:
        dir_embedded = dir_embedded.reshape(N_rays, S, -1)  # the reference's repeat_interleave'd layout
    out = G.evaluate_mlp(nerf_sdf, xyz_input, embed_xyz=embedding_xyz, dir_embedded=dir_embedded, code=env_code,
                         chunk=4096, sigma_only=weights_only)
    if weights_only:
        raw = torch.cat([torch.zeros(N_rays, S, 3, device=out.device), out], -1)
    else:
        raw = out
    feat = None
    if "nerf_feat" in models:   # rendering.py:174-178
        feat = G.evaluate_mlp(models["nerf_feat"], xyz_input, embed_xyz=embedding_xyz, chunk=4096)
    # the reference draws the noise unconditionally (rendering.py:193): keep the generator in step
    noise = torch.randn(N_rays, S, device=raw.device)
    noise = noise * noise_std if noise_std != 0 else None
    mask = None
    if clip_bound is not None:
        cb = torch.as_tensor(clip_bound, dtype=torch.float32, device=raw.device)[None, None]
        mask = (xyz_input.abs() > cb).sum(-1) > 0
    if vis_pred is not None:
        m2 = vis_pred < 0.5
        mask = m2 if mask is None else (mask | m2)
    xa, xb = cyc_pair if cyc_pair is not None else (None, None)
    rgb, depth, sil, weights, vis, cyc = CompositeFn.apply(raw, z_vals, dir_, nerf_sdf.beta, noise, mask, xa, xb)
    if feat is not None:
        feat = WeightedSumFn.apply(weights, feat)   # rendering.py:233
    else:
        feat = torch.zeros_like(rgb)
    if cyc_pair is not None:
        return rgb, feat, depth, weights, vis, sil, cyc
    return rgb, feat, depth, weights, vis, sil


def inference_deform(xyz_coarse_sampled, rays, models, chunk, N_samples, N_rays, embedding_xyz, rays_d, noise_std,
                     obj_bound, dir_embedded, z_vals, img_size, progress, opts, fine_iter=True, render_vis=False):
    """rendering.py:239-579: backward warp, cycle forward warp (+ the warps to the paired frames), trunk + feature +
    compositing, and the per-ray terms of the training losses."""
    is_training = models["coarse"].training
    dist_corresp = bool(getattr(opts, "dist_corresp", False))
    use_corresp = bool(getattr(opts, "use_corresp", False))
    xys = rays.get("xys")
    xyz_coarse_frame = xyz_coarse_sampled
    xyz_coarse_target = xyz_coarse_dentrg = xyz_coarse_sampled
    result = {}
    cyc_pair = None
    has_flow = "flowbw" in models
    has_bones = "bones" in models and not has_flow   # rendering.py:257 / 289: the flow fields take precedence
    use_lbs = has_bones and bool(getattr(opts, "lbs", False))
    if has_flow:
        # free-form deformation (rendering.py:257-286): backward flow to the canonical space, forward flow back
        # (cycle term ||flow_bw + flow_fw||) and to the paired frames; nets: Transhead / SE3head (nerf.py:200-237)
        code_of = lambda k: rays[k][:, None]
        flow_bw = G.evaluate_mlp(models["flowbw"], xyz_coarse_sampled, embed_xyz=embedding_xyz, code=code_of("time_embedded"),
                                 chunk=chunk // N_samples)
        xyz_coarse_sampled = xyz_coarse_sampled + flow_bw
        if fine_iter:
            fw = lambda k: G.evaluate_mlp(models["flowfw"], xyz_coarse_sampled, embed_xyz=embedding_xyz, code=code_of(k),
                                          chunk=chunk // N_samples)
            cyc_pair = (flow_bw, -fw("time_embedded"))
            if "time_embedded_target" in rays:
                xyz_coarse_target = xyz_coarse_sampled + fw("time_embedded_target")
            if "time_embedded_dentrg" in rays:
                xyz_coarse_dentrg = xyz_coarse_sampled + fw("time_embedded_dentrg")
    elif use_lbs:
        # linear blend skinning (rendering.py:303-360 with opts.lbs; geom_utils.py:304-348, 906-931): Gaussian-bone
        # weights from the skinning kernel, the blend as one batched (S,B) x (B,12) product per ray
        bones_rst = models["bones_rst"]
        bone_rts_fw = rays["bone_rts"]
        skin_aux = models["skin_aux"]
        rest_pose_code = models["rest_pose_code"](torch.zeros(1, dtype=torch.long, device=bones_rst.device))
        nerf_skin = models.get("nerf_skin")
        if "nerf_dis" in models:
            raise NotImplementedError("nerf_dis with opts.lbs reads an undefined variable in the reference (rendering.py:324)")
        bones_dfm = G.bone_transform(bones_rst, bone_rts_fw, False, is_vec=True)
        skin_bw = G.gauss_mlp_skinning(xyz_coarse_sampled, embedding_xyz, bones_dfm, rays["time_embedded"], nerf_skin,
                                       skin_aux=skin_aux)
        xyz_coarse_sampled, _ = G.lbs(bones_rst, bone_rts_fw, skin_bw, xyz_coarse_sampled)
        if fine_iter:
            skin_fw = G.gauss_mlp_skinning(xyz_coarse_sampled, embedding_xyz, bones_rst, rest_pose_code, nerf_skin,
                                           skin_aux=skin_aux)
            fw = lambda rts: G.lbs(bones_rst, rts, skin_fw, xyz_coarse_sampled, backward=False)[0]
            cyc_pair = (xyz_coarse_frame, fw(bone_rts_fw))
            if dist_corresp and "bone_rts_target" in rays:
                xyz_coarse_target = fw(rays["bone_rts_target"])
            if dist_corresp and "bone_rts_dentrg" in rays:
                xyz_coarse_dentrg = fw(rays["bone_rts_dentrg"])
    elif has_bones:
        bones_rst = models["bones_rst"]
        bone_rts_fw = rays["bone_rts"]
        skin_aux = models["skin_aux"]
        rest_pose_code = models["rest_pose_code"]
        rest_pose_code = rest_pose_code(torch.zeros(1, dtype=torch.long, device=bones_rst.device))
        nerf_skin = models.get("nerf_skin")
        nerf_dis = models.get("nerf_dis")
        time_embedded = rays["time_embedded"]
        # backward warp (rendering.py:303-322): delta logits, then skinning + DQ blend fused in one kernel
        dskin_bw = G.mlp_skinning(nerf_skin, time_embedded, xyz_coarse_sampled, embed_xyz=embedding_xyz, _pitched=True)
        xyz_in = xyz_coarse_sampled
        xyz_coarse_sampled = G.warp_points(xyz_in, bones_rst, bone_rts_fw, skin_aux, dskin_bw, backward=True)
        if nerf_dis is not None:   # geom_utils.py:416-418: the residual is evaluated at the un-warped points
            xyz_dis = G.evaluate_mlp(nerf_dis, xyz_in, embedding_xyz, code=time_embedded[:, None], chunk=xyz_in.shape[0])
            xyz_coarse_sampled = xyz_coarse_sampled - xyz_dis
            result["dis_reg"] = torch.norm(xyz_dis, dim=2, keepdim=False)
        if fine_iter:
            # cycle forward warp (rendering.py:330-341) and the warps of the same canonical points to the paired frames
            # (:345-360): one set of forward skinning weights (delta logits from the rest pose code), three transforms
            dskin_fw = G.mlp_skinning(nerf_skin, rest_pose_code, xyz_coarse_sampled, embed_xyz=embedding_xyz, _pitched=True)
            if nerf_dis is None:
                fw = lambda rts: G.warp_points(xyz_coarse_sampled, bones_rst, rts, skin_aux, dskin_fw, backward=False)
            else:
                # geom_utils.py:423-429: the weights come from the undisplaced canonical points, the blend acts on the
                # displaced ones -> explicit weights, then the blend kernel
                skin_fw = G.skinning(bones_rst, xyz_coarse_sampled, dskin_fw, skin_aux=skin_aux)
                xyz_dis_fw = G.evaluate_mlp(nerf_dis, xyz_coarse_sampled, embedding_xyz, code=rest_pose_code,
                                            chunk=xyz_coarse_sampled.shape[0])
                result["dis_reg_forward"] = torch.norm(xyz_dis_fw, dim=2, keepdim=False)
                xyz_disp = xyz_coarse_sampled + xyz_dis_fw
                fw = lambda rts: G.neu_dbs(bones_rst, rts, skin_fw, xyz_disp, backward=False)[0]
            xyz_coarse_frame_cyc = fw(bone_rts_fw)
            cyc_pair = (xyz_coarse_frame, xyz_coarse_frame_cyc)
            if dist_corresp and "bone_rts_target" in rays:
                xyz_coarse_target = fw(rays["bone_rts_target"])
            if dist_corresp and "bone_rts_dentrg" in rays:
                xyz_coarse_dentrg = fw(rays["bone_rts_dentrg"])
    env_code = rays.get("env_code")
    if render_vis:
        clip_bound = obj_bound
        vis_pred = G.evaluate_mlp(models["nerf_vis"], xyz_coarse_sampled, embed_xyz=embedding_xyz,
                                  chunk=chunk)[..., 0].sigmoid()
    else:
        clip_bound, vis_pred = None, None
    if opts is not None and getattr(opts, "symm_shape", False):
        # rendering.py:385-391: half of the samples (a fresh draw per call) are evaluated at their x-mirror image
        xyz_x = xyz_coarse_sampled[..., :1]
        symm_mask = torch.rand_like(xyz_x) < 0.5
        xyz_input = torch.cat([torch.where(symm_mask, -xyz_x, xyz_x), xyz_coarse_sampled[..., 1:3]], -1)
    else:
        xyz_input = xyz_coarse_sampled
    out = inference(models, embedding_xyz, xyz_input, rays_d, dir_embedded, z_vals, N_rays, N_samples,
                    chunk, noise_std, env_code=env_code, clip_bound=clip_bound, vis_pred=vis_pred,
                    scale_rgb=getattr(opts, "scale_rgb", 1.3), rgb_filter=getattr(opts, "rgb_filter", False),
                    cyc_pair=cyc_pair)
    rgb_coarse, feat_rnd, depth_rnd, weights_coarse, vis_coarse, sil_coarse = out[:6]
    result["img_coarse"] = rgb_coarse
    result["depth_rnd"] = depth_rnd
    result["sil_coarse"] = sil_coarse
    if render_vis:
        result["vis_pred"] = (vis_pred * weights_coarse).sum(-1)
    if not fine_iter:
        return result, weights_coarse

    pts_target = None
    if use_corresp and not dist_corresp and "rtk_vec_target" in rays:   # rendering.py:407-411
        pts_exp = L.compute_pts_exp(weights_coarse, xyz_coarse_sampled)
        pts_target = L.kp_reproj(pts_exp, models, embedding_xyz, rays, to_target=True, neudbs=opts.neudbs)
    if "feats_at_samp" in rays:   # rendering.py:413-432: feature matching + 3d-2d reprojection
        pts_pred, pts_exp, feat_err, corr_err = L.feat_match_loss(
            models["nerf_feat"], embedding_xyz, rays["feats_at_samp"], xyz_coarse_sampled, weights_coarse, obj_bound,
            getattr(opts, "use_corr", False), getattr(opts, "use_ot", False), is_training=is_training)
        proj_err = L.kp_reproj_loss(pts_pred, xys, models, embedding_xyz, rays, neudbs=opts.neudbs)
        result["pts_pred"], result["pts_exp"] = pts_pred, pts_exp
        result["feat_err"] = feat_err
        if getattr(opts, "use_corr", False):
            result["corr_err"] = corr_err
        result["proj_err"] = proj_err / img_size * 2
    # (the projections of the warped samples into the paired views, :434-459, happen together with the flow rendering below)

    result["xyz_camera_vis"] = xyz_coarse_frame
    if has_bones or has_flow:
        result["xyz_canonical_vis"] = xyz_coarse_sampled
    if "feats_at_samp" in rays:
        result["pts_exp_vis"] = pts_exp
        result["pts_pred_vis"] = pts_pred
    if has_bones or has_flow:
        result["frame_cyc_dis"] = out[6]
    if is_training and "nerf_vis" in models:   # :475-477
        result["vis_loss"] = L.visibility_loss(models["nerf_vis"], embedding_xyz, xyz_coarse_sampled, vis_coarse,
                                               obj_bound, chunk)
    flo_coarse = flo_valid = None
    if "rtk_vec_target" in rays:   # :480-489
        if dist_corresp:
            flo_coarse, flo_valid = G.project_render_flo(weights_coarse, xyz_coarse_target, rays["rtk_vec_target"], xys,
                                                         img_size, N_rays)
        else:
            if pts_target is None:
                raise RuntimeError("rtk_vec_target without dist_corresp needs opts.use_corresp (rendering.py:407-411)")
            flo_coarse = G.diff_flo(pts_target, xys, img_size)
            flo_valid = torch.ones_like(flo_coarse[..., :1])
        result["flo_coarse"], result["flo_valid"] = flo_coarse, flo_valid
    if "rtk_vec_dentrg" in rays:   # :491-499
        if not dist_corresp:
            raise NotImplementedError("rtk_vec_dentrg without dist_corresp reads an undefined variable in the reference "
                                      "(rendering.py:496)")
        result["fdp_coarse"], result["fdp_valid"] = G.project_render_flo(weights_coarse, xyz_coarse_dentrg,
                                                                         rays["rtk_vec_dentrg"], xys, img_size, N_rays)
    if "img_at_samp" in rays:
        _per_ray_losses(result, rays, rgb_coarse, sil_coarse, is_training, flo_coarse, flo_valid)
    if "feats_at_samp" in rays:   # :572-578
        feat_n = F.normalize(feat_rnd, 2, -1)
        frnd_loss_samp = (feat_n - rays["feats_at_samp"]).pow(2).mean(-1)
        result["frnd_loss_samp"] = frnd_loss_samp * rays["sil_at_samp"][..., 0]
    return result, weights_coarse


def _per_ray_losses(result, rays, rgb_coarse, sil_coarse, is_training, flo_coarse=None, flo_valid=None):
    """rendering.py:516-570: O(N_rays) bookkeeping on already-rendered values."""
    img_at_samp, sil_at_samp, vis_at_samp = rays["img_at_samp"], rays["sil_at_samp"], rays["vis_at_samp"]
    img_loss_samp = (rgb_coarse - img_at_samp).pow(2).mean(-1)[..., None]
    sil_balance_wt = 1
    if is_training:
        # the reference tests `sil_at_samp.sum()>0` on the host (a device sync, rendering.py:535); the same
        # weights are formed on the device and selected with where()
        vis_on = (vis_at_samp > 0).to(sil_at_samp.dtype)
        pos = (sil_at_samp * vis_on).sum()
        tot = vis_at_samp.sum()
        neg = ((1 - sil_at_samp) * vis_on).sum()
        wt = 0.5 * (tot / pos) * sil_at_samp + 0.5 * (tot / neg) * (1 - sil_at_samp)
        ok = (sil_at_samp.sum() > 0) & ((1 - sil_at_samp).sum() > 0)
        sil_balance_wt = torch.where(ok, wt, torch.ones_like(wt))
    sil_loss_samp = (sil_coarse[..., None] - sil_at_samp).pow(2) * sil_balance_wt * vis_at_samp
    result["img_at_samp"], result["sil_at_samp"], result["vis_at_samp"] = img_at_samp, sil_at_samp, vis_at_samp
    if flo_coarse is not None and "flo_at_samp" in rays:   # :545-553
        flo_at_samp, cfd_at_samp = rays["flo_at_samp"], rays["cfd_at_samp"]
        flo_loss_samp = (flo_coarse - flo_at_samp).pow(2).sum(-1)
        sil_at_samp_flo = (sil_at_samp > 0) & (flo_valid == 1) & ~(cfd_at_samp == 0)
        sel = sil_at_samp_flo.to(cfd_at_samp.dtype)
        cnt = sel.sum()
        mean_cfd = (cfd_at_samp * sel).sum() / cnt.clamp_min(1)
        cfd = torch.where(cnt > 0, cfd_at_samp / mean_cfd, cfd_at_samp)
        result["sil_at_samp_flo"] = sil_at_samp_flo
        result["flo_at_samp"] = flo_at_samp
        result["flo_loss_samp"] = flo_loss_samp[..., None] * cfd * sil_at_samp
    result["img_loss_samp"] = img_loss_samp * sil_at_samp
    result["sil_loss_samp"] = sil_loss_samp
