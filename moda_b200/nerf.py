"""``Embedding`` and ``NeRF`` with the reference's constructor arguments, forward signatures and
state-dict names (nnutils/nerf.py:13-198), executed by the CUDA linear-layer kernels.

State-dict compatibility matters because checkpoints chain across training stages
(SURVEY.md section 5): ``xyz_encoding_{i}.0.weight``, ``xyz_encoding_final``, ``dir_encoding.0``,
``sigma``, ``rgb.0``, ``beta``.
"""
import torch
from torch import nn

from .ops import EmbedFn, MlpFn, MlpSpec, pe_window, SEG_DENSE


class Embedding(nn.Module):
    """(x, w_k sin(2^k x), w_k cos(2^k x))_k with the annealing window (nerf.py:13-75)."""

    def __init__(self, in_channels, N_freqs, logscale=True, alpha=None):
        super().__init__()
        if not logscale:
            raise NotImplementedError("only logscale=True frequency bands (2^k) are used by MoDA")
        self.N_freqs = N_freqs
        self.in_channels = in_channels
        self.nfuncs = 2
        self.out_channels = in_channels * (2 * N_freqs + 1)
        self.alpha = self.N_freqs if alpha is None else alpha
        self.freq_bands = 2 ** torch.linspace(0, N_freqs - 1, N_freqs) if N_freqs > 0 else torch.zeros(0)

    def window(self):
        return pe_window(self.N_freqs, self.alpha)

    def forward(self, x):
        if self.N_freqs <= 0:
            return x
        return EmbedFn.apply(x, self.N_freqs, self.window())


class NeRF(nn.Module):
    """D x W MLP with skip connection, sigma / rgb heads and learnable ``beta`` (nerf.py:83-198)."""

    def __init__(self, D=8, W=256, in_channels_xyz=63, in_channels_dir=27, out_channels=3, skips=[4],
                 raw_feat=False, init_beta=1. / 100, activation=nn.ReLU(True), in_channels_code=0,
                 enable_semantic=False):
        super().__init__()
        if not isinstance(activation, nn.ReLU):
            raise NotImplementedError("the CUDA MLP path implements ReLU, the activation MoDA uses")
        if enable_semantic:
            raise NotImplementedError("enable_semantic is disabled everywhere in the reference (moda.py:273)")
        self.D, self.W = D, W
        self.in_channels_xyz, self.in_channels_dir = in_channels_xyz, in_channels_dir
        self.in_channels_code = in_channels_code
        self.skips = list(skips)
        self.use_xyz = False
        self.enable_semantic = enable_semantic
        self.out_channels = out_channels
        self.weights_reg = []
        for i in range(D):
            if i == 0:
                layer = nn.Linear(in_channels_xyz, W)
                self.weights_reg.append("xyz_encoding_%d" % (i + 1))
            elif i in skips:
                layer = nn.Linear(W + in_channels_xyz, W)
                self.weights_reg.append("xyz_encoding_%d" % (i + 1))
            else:
                layer = nn.Linear(W, W)
            setattr(self, "xyz_encoding_%d" % (i + 1), nn.Sequential(layer, activation))
        self.xyz_encoding_final = nn.Linear(W, W)
        self.dir_encoding = nn.Sequential(nn.Linear(W + in_channels_dir, W // 2), activation)
        self.sigma = nn.Linear(W, 1)
        self.rgb = nn.Sequential(nn.Linear(W // 2, out_channels))
        self.raw_feat = raw_feat
        self.beta = nn.Parameter(torch.Tensor([init_beta]))

    def param_list(self):
        """[W1,b1,...,WD,bD, Wf,bf, Wd,bd, Ws,bs, Wr,br] in the order MlpFn expects."""
        ps = []
        for i in range(self.D):
            lin = getattr(self, "xyz_encoding_%d" % (i + 1))[0]
            ps += [lin.weight, lin.bias]
        ps += [self.xyz_encoding_final.weight, self.xyz_encoding_final.bias,
               self.dir_encoding[0].weight, self.dir_encoding[0].bias,
               self.sigma.weight, self.sigma.bias, self.rgb[0].weight, self.rgb[0].bias]
        return ps

    def run(self, M, inputs, xyz_segs, dir_segs, win, sigma_only=False):
        """Evaluates the MLP on a virtual (M, in_xyz + in_dir) input assembled from ``inputs``."""
        assert sum(s[2] for s in xyz_segs) == self.in_channels_xyz, "xyz segments do not add up to in_channels_xyz"
        if not sigma_only:
            assert sum(s[2] for s in dir_segs) == self.in_channels_dir, "dir segments do not add up"
        spec = MlpSpec(self.D, self.W, self.out_channels, self.skips, self.raw_feat, sigma_only, xyz_segs,
                       [] if sigma_only else dir_segs, win, len(inputs))
        return MlpFn.apply(spec, M, *inputs, *self.param_list())

    def forward(self, x, xyz=None, sigma_only=False):
        """x: (..., in_channels_xyz [+ in_channels_dir]) -> (..., 4) [rgb|sigma] / (..., 1) / raw features."""
        lead = x.shape[:-1]
        x2 = x.reshape(-1, x.shape[-1])
        cx = self.in_channels_xyz
        xyz_segs = [(SEG_DENSE, 0, cx, 1, 0)]
        dir_segs = [(SEG_DENSE, 0, self.in_channels_dir, 1, cx)] if self.in_channels_dir > 0 else []
        out = self.run(x2.shape[0], [x2], xyz_segs, dir_segs, None, sigma_only)
        return out.reshape(lead + (out.shape[-1],))


class DQ_RTHead(NeRF):
    """Pose MLP -> one dual quaternion per bone (nnutils/nerf.py:239-279): out (bs, 1, B*8) with
    dq = [normalize(r), 0.5 * (0, 0.1 t) (x) normalize(r)].  Evaluated once per FRAME (moda.py:1304), not per sample."""

    def __init__(self, use_quat, **kwargs):
        super().__init__(**kwargs)
        if not use_quat:
            raise NotImplementedError("DQ_RTHead is only used with use_quat=True (moda.py:313-317)")
        self.use_quat = use_quat
        self.num_output = 7
        for m in self.modules():
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()

    def forward(self, x):
        from . import dual_quat as DQ
        x = super().forward(x)
        bs = x.shape[0]
        rts = x.reshape(-1, self.num_output)
        tmat = rts[:, 0:3] * 0.1
        rquat = DQ.q_normalize(rts[:, 3:7].contiguous())
        tquat = torch.cat((torch.zeros_like(tmat[:, :1]), tmat), -1)
        dq_d = 0.5 * DQ.q_mul(tquat.contiguous(), rquat)
        return torch.cat((rquat, dq_d), -1).reshape(bs, 1, -1)


class Transhead(NeRF):
    """Translation flow field (nnutils/nerf.py:200-210): 0.1 * MLP(x).  ``post`` is the output stage that
    geom_utils.evaluate_mlp applies after running the MLP on the CUDA kernels."""

    def post(self, raw, xyz=None):
        return raw * 0.1

    def forward(self, x, xyz=None, sigma_only=False):
        out = super().forward(x, sigma_only=sigma_only)
        return out if sigma_only else self.post(out, xyz)


def so3_exp_map(log_rot, eps=1e-4):
    """pytorch3d ``so3_exponential_map`` (third_party/pytorch3d/.../so3.py:110-176): Rodrigues' formula with the squared
    angle clamped at eps.  (n,3) -> (n,3,3)."""
    x, y, z = log_rot.unbind(-1)
    ang = (log_rot * log_rot).sum(-1).clamp_min(eps).sqrt()
    f1 = ang.sin() / ang
    f2 = (1.0 - ang.cos()) / (ang * ang)
    o = torch.zeros_like(x)
    K = torch.stack([o, -z, y, z, o, -x, -y, x, o], -1).reshape(-1, 3, 3)
    eye = torch.eye(3, dtype=log_rot.dtype, device=log_rot.device)
    return f1[:, None, None] * K + f2[:, None, None] * torch.bmm(K, K) + eye


class SE3head(NeRF):
    """Per-point rigid flow field (nnutils/nerf.py:212-237, after Nerfies): the MLP emits (rotation, pivot, translation);
    flow = R (x + 0.1 pivot) - 0.1 pivot + 0.1 translation - x with R = exp(rotation)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)
        self.use_xyz = True

    def post(self, raw, xyz):
        shape = xyz.shape
        r = raw.reshape(-1, 9)
        pivot, trans = r[:, 3:6] * 0.1, r[:, 6:9] * 0.1
        p = xyz.reshape(-1, 3)
        w = (so3_exp_map(r[:, 0:3]) * (p + pivot)[:, None, :]).sum(-1) - pivot + trans
        return w.reshape(shape) - xyz

    def forward(self, x, xyz=None, sigma_only=False):
        out = super().forward(x, sigma_only=sigma_only)
        return out if sigma_only else self.post(out, xyz)


class RTHead(NeRF):
    """Pose MLP of the LBS motion model (nnutils/nerf.py:307-344): per bone [R (9, row-major) | 0.1 t]; R from a
    quaternion (use_quat) or the exponential map of a delta rotation.  Evaluated once per frame (moda.py:307-311)."""

    def __init__(self, use_quat, **kwargs):
        super().__init__(**kwargs)
        self.use_quat = use_quat
        self.num_output = 7 if use_quat else 6
        for m in self.modules():
            if isinstance(m, nn.Linear) and m.bias is not None:
                m.bias.data.zero_()

    def forward(self, x):
        from .geom_utils import quaternion_to_matrix
        x = super().forward(x)
        bs = x.shape[0]
        rts = x.reshape(-1, self.num_output)
        tmat = rts[:, 0:3] * 0.1
        if self.use_quat:
            rmat = quaternion_to_matrix(torch.nn.functional.normalize(rts[:, 3:7], 2, -1))
        else:
            rmat = so3_exp_map(rts[:, 3:6])
        return torch.cat([rmat.reshape(-1, 9), tmat], -1).reshape(bs, 1, -1)


def fid_reindex(fid, num_vids, vid_offset):
    """geom_utils.py:1759-1777: absolute frame id -> (video id, frame id relative to the middle of its video,
    normalised by the longest video)."""
    tid = torch.zeros_like(fid).float()
    vid = torch.zeros_like(fid)
    max_ts = float((vid_offset[1:] - vid_offset[:-1]).max())
    for i in range(num_vids):
        assign = torch.logical_and(fid >= int(vid_offset[i]), fid < int(vid_offset[i + 1]))
        vid = torch.where(assign, torch.full_like(vid, i), vid)
        doffset = float(vid_offset[i])
        dnmax = float(vid_offset[i + 1]) - doffset
        tid = torch.where(assign, (fid.float() - doffset - dnmax / 2) / max_ts * 2, tid)
    return vid, tid


class FrameCode(nn.Module):
    """Frame index -> code through a per-video Fourier basis (nnutils/nerf.py:346-380): pose / environment codes."""

    def __init__(self, num_freq, embedding_dim, vid_offset, scale=1):
        super().__init__()
        import numpy as np
        self.vid_offset = np.asarray(vid_offset)
        self.num_vids = len(vid_offset) - 1
        max_ts = (self.vid_offset[1:] - self.vid_offset[:-1]).max()
        self.num_freq = 2 * int(np.log2(max_ts)) - 2
        self.fourier_embed = Embedding(1, num_freq, alpha=num_freq)
        self.basis_mlp = nn.Linear(self.num_vids * self.fourier_embed.out_channels, embedding_dim)
        self.scale = scale

    def forward(self, fid):
        bs = fid.shape[0]
        vid, tid = fid_reindex(fid, self.num_vids, self.vid_offset)
        tid = (tid * self.scale).reshape(bs, 1)
        vid = vid.reshape(bs, 1)
        coeff = self.fourier_embed(tid)
        onehot = torch.nn.functional.one_hot(vid, num_classes=self.num_vids)
        coeff = (coeff[..., None] * onehot).reshape(bs, -1)
        return torch.nn.functional.linear(coeff, self.basis_mlp.weight, self.basis_mlp.bias)
