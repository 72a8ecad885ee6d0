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
