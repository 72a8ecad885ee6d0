"""Tensor-core execution of the small auxiliary MLPs of a default-flag MoDA step -- ``nerf_feat`` (5x128 -> 16
canonical features, nnutils/moda.py:447-449) and ``nerf_vis`` (5x64 -> 1 visibility logit, :344-348), and any other
``NeRF(raw_feat=True, in_channels_dir=0)`` on a plain PE(xyz) input with W in {64, 128} -- layer by layer on the
tcgen05 kernels of csrc/tc_gemm.cu (the same building blocks as ``trunk_tc``): fp16 operands, fp32 accumulation,
activations kept in fp16 between layers, weight gradients on ``moda_tc_wgrad`` straight from the saved fp16
activations, the backward chain under a power-of-two loss scale.

NeRF.forward with raw_feat (nerf.py:163-198): D x (Linear + ReLU) with the input re-concatenated at the skip layer,
xyz_encoding_final (no activation), dir_encoding (W -> W/2, ReLU), rgb (W/2 -> out); the sigma head is computed and
discarded by the reference and is skipped here.  The fp32 SIMT path (ops.MlpFn) remains the exact-mode reference.
"""
import torch

from ._lib import call, ptr, stream, f32
from .ops import _win_array
from .trunk_tc import _pack, _tcl, _wgrad as _wgrad_full

HALF = torch.float16


def _pad64(n):
    return (n + 63) // 64 * 64


def supported(model, embed_xyz, k):
    return (embed_xyz is not None and embed_xyz.N_freqs == 10 and k == 3 and model.raw_feat and model.in_channels_dir == 0
            and model.in_channels_xyz == 63 and model.W in (64, 128) and model.out_channels <= 32
            and len(model.skips) <= 1 and all(0 < s < model.D for s in model.skips) and model.D >= 2)


def _wgrad(dY, N, X, K, M, dW, col0, n_valid, k_valid, oscale, dbias=None):
    call("moda_tc_wgrad", ptr(dY), dY.stride(0), N, ptr(X), X.stride(0), K, M, ptr(dW) + 4 * col0, dW.stride(0), n_valid,
         k_valid, ptr(oscale), ptr(dbias), stream())


class _Packed:
    """fp16 copies of the weights in the layouts the kernels read: forward B[N = out, K = in] (K-major rows),
    backward B[N = in, K = out]; every dimension zero-padded to a multiple of 64."""

    def __init__(self, params, D, W, skip, Wh, oc, need_backward):
        dev = params[0].device
        z = lambda r, c: torch.zeros(r, c, device=dev, dtype=HALF)
        Wl = [params[2 * i] for i in range(D)]
        Wf, Wd, Wr = params[2 * D], params[2 * D + 2], params[2 * D + 6]
        WhP, ocP = _pad64(Wh), 32
        self.fwd = []
        for i in range(D):
            if i == 0:
                t = z(W, 64)
                _pack(Wl[0], 63, 0, t, 0, W, 64, False)
            elif i == skip:
                t = z(W, 64 + W)
                _pack(Wl[i], 63, 0, t, 0, W, 64, False)
                _pack(Wl[i], W, 63, t, 64, W, W, False)
            else:
                t = z(W, W)
                _pack(Wl[i], W, 0, t, 0, W, W, False)
            self.fwd.append(t)
        self.Wf = z(W, W)
        _pack(Wf, W, 0, self.Wf, 0, W, W, False)
        self.Wd = z(WhP, W)
        _pack(Wd, W, 0, self.Wd, 0, WhP, W, False)
        self.Wr = z(ocP, WhP)
        _pack(Wr, Wh, 0, self.Wr, 0, ocP, WhP, False)
        if need_backward:
            self.T = [None] * D
            for i in range(1, D):
                t = z(W, W)
                _pack(Wl[i], W, 63 if i == skip else 0, t, 0, W, W, True)
                self.T[i] = t
            self.T_pe1 = z(64, W)
            _pack(Wl[0], 63, 0, self.T_pe1, 0, 64, W, True)
            if skip is not None:
                self.T_pes = z(64, W)
                _pack(Wl[skip], 63, 0, self.T_pes, 0, 64, W, True)
            self.WfT = z(W, W)
            _pack(Wf, W, 0, self.WfT, 0, W, W, True)
            self.WdT = z(W, WhP)
            _pack(Wd, W, 0, self.WdT, 0, W, WhP, True)
            self.WrT = z(WhP, 64)
            _pack(Wr, Wh, 0, self.WrT, 0, WhP, 64, True)


class GenericTcFn(torch.autograd.Function):
    """apply(xyz (P,3), win, save, D, W, skip | None, *params) -> (P, 32) fp32, columns >= out_channels are zero.
    params = [W1,b1,...,WD,bD, Wf,bf, Wd,bd, Ws,bs, Wr,br] (NeRF.param_list())."""

    @staticmethod
    def forward(ctx, xyz, win, save, D, W, skip, *params):
        xyz_shape = xyz.shape
        xyz = f32(xyz).reshape(-1, 3)
        P, dev = xyz.shape[0], xyz.device
        ctx.param_refs = params
        params = [f32(p) for p in params]
        need_bw = bool(save) and any(ctx.needs_input_grad)
        Wf, bf, Wd, bd, _, _, Wr, br = params[2 * D:2 * D + 8]
        Wh, oc = Wd.shape[0], Wr.shape[0]
        WhP = _pad64(Wh)
        pk = _Packed(params, D, W, skip, Wh, oc, need_bw)
        b = [params[2 * i + 1] for i in range(D)]
        h16 = lambda n: torch.empty(P, n, device=dev, dtype=HALF)
        wa, _ = _win_array(win)
        A0 = h16(64)
        call("moda_pe16_fwd", ptr(xyz), ptr(A0), None, 64, P, len(win), wa, stream())
        H = []
        for i in range(D):
            y = h16(W)
            if i == 0:
                _tcl(A0, 64, None, 0, pk.fwd[0], P, W, bias=b[0], relu=1, y16=y)
            elif i == skip:
                _tcl(A0, 64, H[i - 1], W, pk.fwd[i], P, W, bias=b[i], relu=1, y16=y)
            else:
                _tcl(H[i - 1], W, None, 0, pk.fwd[i], P, W, bias=b[i], relu=1, y16=y)
            H.append(y)
            if not need_bw and i >= 1:
                H[i - 1] = None   # inference: a layer's input is dead once its output exists
        fin = h16(W)
        _tcl(H[D - 1], W, None, 0, pk.Wf, P, W, bias=bf, y16=fin)
        bdp = torch.zeros(WhP, device=dev, dtype=torch.float32)
        bdp[:Wh] = bd
        dfe = h16(WhP)
        _tcl(fin, W, None, 0, pk.Wd, P, WhP, bias=bdp, relu=1, y16=dfe)
        brp = torch.zeros(32, device=dev, dtype=torch.float32)
        brp[:oc] = br
        out = torch.empty(P, 32, device=dev, dtype=torch.float32)
        _tcl(dfe, WhP, None, 0, pk.Wr, P, 32, bias=brp, y32=out)
        if need_bw:
            ctx.save_for_backward(xyz, *params)
            ctx.act = (A0, H, fin, dfe, pk)
            ctx.meta = (win, xyz_shape, D, W, skip, Wh, oc)
        return out

    @staticmethod
    def backward(ctx, gout):
        from .chain_tc import _grad_targets, _loss_scale
        xyz = ctx.saved_tensors[0]
        params = list(ctx.saved_tensors[1:])
        A0, H, fin, dfe, pk = ctx.act
        win, xyz_shape, D, W, skip, Wh, oc = ctx.meta
        P, dev = xyz.shape[0], xyz.device
        WhP = _pad64(Wh)
        g, gret = _grad_targets(ctx.param_refs, params)
        gout = f32(gout).reshape(P, 32)
        sc, isc = _loss_scale(gout)
        h16 = lambda n: torch.empty(P, n, device=dev, dtype=HALF)
        # rgb layer: dYr = scale * gout as fp16 (64 columns, zero padded)
        dYr = h16(64)
        call("moda_split16", ptr(gout), 32, 32, ptr(sc), ptr(dYr), None, 64, 64, P, stream())
        iW, ib = 2 * D + 6, 2 * D + 7
        _wgrad(dYr, 64, dfe, WhP, P, g[iW], 0, oc, Wh, isc, dbias=g[ib])
        d_dfe = h16(WhP)
        _tcl(dYr, 64, None, 0, pk.WrT, P, WhP, mask=dfe, y16=d_dfe)
        # dir layer
        _wgrad(d_dfe, WhP, fin, W, P, g[2 * D + 2], 0, Wh, W, isc, dbias=g[2 * D + 3])
        d_fin = h16(W)
        _tcl(d_dfe, WhP, None, 0, pk.WdT, P, W, y16=d_fin)
        # final layer (no activation); the discarded sigma head gets no gradient (nerf.py:178)
        _wgrad(d_fin, W, H[D - 1], W, P, g[2 * D], 0, W, W, isc, dbias=g[2 * D + 1])
        bufs = [h16(W), d_fin]
        _tcl(d_fin, W, None, 0, pk.WfT, P, W, mask=H[D - 1], y16=bufs[0])
        cur = 0
        d_pe = h16(64)
        pe_written = False
        for i in range(D - 1, -1, -1):
            dY = bufs[cur]
            if i == 0:
                _wgrad(dY, W, A0, 64, P, g[0], 0, W, 63, isc, dbias=g[1])
                _tcl(dY, W, None, 0, pk.T_pe1, P, 64, y16=d_pe, acc16=1 if pe_written else 0)
            else:
                if i == skip:
                    _wgrad(dY, W, A0, 64, P, g[2 * i], 0, W, 63, isc)
                    _wgrad(dY, W, H[i - 1], W, P, g[2 * i], 63, W, W, isc, dbias=g[2 * i + 1])
                    _tcl(dY, W, None, 0, pk.T_pes, P, 64, y16=d_pe)
                    pe_written = True
                else:
                    _wgrad(dY, W, H[i - 1], W, P, g[2 * i], 0, W, W, isc, dbias=g[2 * i + 1])
                nxt = 1 - cur
                _tcl(dY, W, None, 0, pk.T[i], P, W, mask=H[i - 1], y16=bufs[nxt])
                cur = nxt
        gxyz = torch.empty(P, 3, device=dev, dtype=torch.float32)
        wa, _ = _win_array(win)
        call("moda_pe16_bwd", ptr(xyz), ptr(d_pe), None, 64, ptr(gxyz), P, len(win), wa, ptr(isc), 0, stream())
        ctx.act = None
        gret[2 * D + 4] = gret[2 * D + 5] = None
        return (gxyz.reshape(xyz_shape), None, None, None, None, None) + tuple(gret)
