"""``render_rays`` with the reference's signature and result dict (nnutils/rendering.py:19-122), executed by
the CUDA kernels of this package: sample generation, nerf_skin delta logits, fused Gaussian-skinning +
dual-quaternion backward / forward warps, the 8x256 trunk, and the compositor with the cycle term.

What is implemented is SURVEY.md section 8(a)'s core path: ``models`` keys {coarse, bones, bones_rst,
skin_aux, nerf_skin, rest_pose_code} (+ nerf_vis for ``render_vis``), ``rays`` keys {rays_o, rays_d, near,
far, xys, time_embedded, bone_rts, env_code} (+ the ``*_at_samp`` per-ray loss inputs).  Branches that
SURVEY.md section 8(f) lists as "next" (flow fields, LBS, feature matching, flow rendering, nerf_unc) raise
NotImplementedError instead of silently doing something else.
"""
import torch

from . import geom_utils as G
from .ops import CompositeFn, PointsFromDepthsFn, SampleRaysFn, SamplePdfFn

_UNSUPPORTED_MODELS = ("flowbw", "flowfw", "nerf_feat", "nerf_unc", "nerf_dis")
_UNSUPPORTED_RAYS = ("rtk_vec_target", "rtk_vec_dentrg", "feats_at_samp", "bone_rts_target", "bone_rts_dentrg",
                     "appearance_code")


def render_rays(models, embeddings, rays, N_samples=64, use_disp=False, perturb=0, noise_std=1, chunk=1024 * 32,
                obj_bound=None, use_fine=False, img_size=None, progress=None, opts=None, render_vis=False):
    """rendering.py:19-122.  Same inputs, same result keys; see the module docstring for the covered flags."""
    for k in _UNSUPPORTED_MODELS:
        if k in models:
            raise NotImplementedError("models['%s'] is outside the accelerated core path (SURVEY.md 8(f))" % k)
    for k in _UNSUPPORTED_RAYS:
        if k in rays:
            raise NotImplementedError("rays['%s'] is outside the accelerated core path (SURVEY.md 8(f))" % k)
    if opts is not None and (getattr(opts, "lbs", False) or not getattr(opts, "neudbs", True)) and "bones" in models:
        raise NotImplementedError("only the dual-quaternion (neudbs) motion model is implemented")
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
    feat = torch.zeros_like(rgb)
    if cyc_pair is not None:
        return rgb, feat, depth, weights, vis, sil, cyc
    return rgb, feat, depth, weights, vis, sil


def inference_deform(xyz_coarse_sampled, rays, models, chunk, N_samples, N_rays, embedding_xyz, rays_d, noise_std,
                     obj_bound, dir_embedded, z_vals, img_size, progress, opts, fine_iter=True, render_vis=False):
    """rendering.py:239-579, core path: backward warp, cycle forward warp, trunk + compositing, per-ray terms."""
    is_training = models["coarse"].training
    xyz_coarse_frame = xyz_coarse_sampled
    result = {}
    cyc_pair = None
    has_bones = "bones" in models
    if has_bones:
        bones_rst = models["bones_rst"]
        bone_rts_fw = rays["bone_rts"]
        skin_aux = models["skin_aux"]
        rest_pose_code = models["rest_pose_code"]
        rest_pose_code = rest_pose_code(torch.zeros(1, dtype=torch.long, device=bones_rst.device))
        nerf_skin = models.get("nerf_skin")
        time_embedded = rays["time_embedded"]
        # backward warp (rendering.py:303-322): delta logits, then skinning + DQ blend fused in one kernel
        dskin_bw = G.mlp_skinning(nerf_skin, time_embedded, xyz_coarse_sampled, embed_xyz=embedding_xyz, _pitched=True)
        xyz_coarse_sampled = G.warp_points(xyz_coarse_sampled, bones_rst, bone_rts_fw, skin_aux, dskin_bw,
                                           backward=True)
        if fine_iter:
            # cycle forward warp (rendering.py:330-341)
            dskin_fw = G.mlp_skinning(nerf_skin, rest_pose_code, xyz_coarse_sampled, embed_xyz=embedding_xyz, _pitched=True)
            xyz_coarse_frame_cyc = G.warp_points(xyz_coarse_sampled, bones_rst, bone_rts_fw, skin_aux, dskin_fw,
                                                 backward=False)
            cyc_pair = (xyz_coarse_frame, xyz_coarse_frame_cyc)
    env_code = rays.get("env_code")
    if render_vis:
        clip_bound = obj_bound
        vis_pred = G.evaluate_mlp(models["nerf_vis"], xyz_coarse_sampled, embed_xyz=embedding_xyz,
                                  chunk=chunk)[..., 0].sigmoid()
    else:
        clip_bound, vis_pred = None, None
    if opts is not None and getattr(opts, "symm_shape", False):
        raise NotImplementedError("symm_shape is outside the accelerated core path")
    out = inference(models, embedding_xyz, xyz_coarse_sampled, rays_d, dir_embedded, z_vals, N_rays, N_samples,
                    chunk, noise_std, env_code=env_code, clip_bound=clip_bound, vis_pred=vis_pred,
                    scale_rgb=getattr(opts, "scale_rgb", 1.3), rgb_filter=getattr(opts, "rgb_filter", False),
                    cyc_pair=cyc_pair)
    rgb_coarse, _, depth_rnd, weights_coarse, vis_coarse, sil_coarse = out[:6]
    result["img_coarse"] = rgb_coarse
    result["depth_rnd"] = depth_rnd
    result["sil_coarse"] = sil_coarse
    if render_vis:
        result["vis_pred"] = (vis_pred * weights_coarse).sum(-1)
    if fine_iter:
        result["xyz_camera_vis"] = xyz_coarse_frame
        if has_bones:
            result["xyz_canonical_vis"] = xyz_coarse_sampled
            result["frame_cyc_dis"] = out[6]
        if "img_at_samp" in rays:
            _per_ray_losses(result, rays, rgb_coarse, sil_coarse, is_training)
    return result, weights_coarse


def _per_ray_losses(result, rays, rgb_coarse, sil_coarse, is_training):
    """rendering.py:516-566 without the flow term: O(N_rays) bookkeeping on already-rendered values."""
    img_at_samp, sil_at_samp, vis_at_samp = rays["img_at_samp"], rays["sil_at_samp"], rays["vis_at_samp"]
    img_loss_samp = (rgb_coarse - img_at_samp).pow(2).mean(-1)[..., None]
    sil_balance_wt = 1
    if is_training:
        # the reference tests `sil_at_samp.sum()>0` on the host (a device sync, rendering.py:535); the same
        # weights are formed on the device and selected with where()
        pos = sil_at_samp[vis_at_samp > 0].sum() if False else (sil_at_samp * (vis_at_samp > 0)).sum()
        tot = vis_at_samp.sum()
        neg = ((1 - sil_at_samp) * (vis_at_samp > 0)).sum()
        wt = 0.5 * (tot / pos) * sil_at_samp + 0.5 * (tot / neg) * (1 - sil_at_samp)
        ok = (sil_at_samp.sum() > 0) & ((1 - sil_at_samp).sum() > 0)
        sil_balance_wt = torch.where(ok, wt, torch.ones_like(wt))
    sil_loss_samp = (sil_coarse[..., None] - sil_at_samp).pow(2) * sil_balance_wt * vis_at_samp
    result["img_at_samp"], result["sil_at_samp"], result["vis_at_samp"] = img_at_samp, sil_at_samp, vis_at_samp
    result["img_loss_samp"] = img_loss_samp * sil_at_samp
    result["sil_loss_samp"] = sil_loss_samp
