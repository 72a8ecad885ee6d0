"""Tensor-core execution of the 8x256 canonical MLP (``nerf_coarse``, nnutils/nerf.py:147-198) on the
virtual input [PE(xyz) | dir_embedded | env_code] that ``inference`` assembles (nnutils/rendering.py:156-163,
nnutils/geom_utils.py:19-57).

fp16 operands / fp32 accumulation on tcgen05 (csrc/tc_gemm.cu); activations are kept in fp16 between layers;
the per-ray-constant dir/env columns of the direction layer are hoisted into a per-ray bias; the backward
gradient chain runs in fp16 under a dynamic power-of-two loss scale (csrc/tc_support.cu: loss_scale) that is
divided out again in every fp32 result.  Precision of this mode is discussed in DESIGN.md ("precision").
"""
import ctypes

import torch

from ._lib import call, ptr, stream, f32
from .ops import _win_array

HALF = torch.float16


def supported(model, n_dir_cols):
    """The fast path covers the reference's nerf_coarse shape (moda.py:271-273): 8x256, skip at layer 5."""
    return (model.D == 8 and model.W == 256 and list(model.skips) == [4] and model.in_channels_xyz == 63
            and model.in_channels_dir == n_dir_cols and model.out_channels == 3 and not model.raw_feat)


def _tcl(A1, K1, A2, K2, B, M, N, bias=None, rowbias=None, rep=1, relu=0, mask=None, rv=None, cv=None, rscale=None,
         y16=None, acc16=0, y32=None, oscale=None):
    call("moda_tc_linear", ptr(A1), A1.stride(0), K1, ptr(A2) if A2 is not None else None,
         A2.stride(0) if A2 is not None else 0, K2, ptr(B), B.stride(0), M, N, ptr(bias), ptr(rowbias), rep, relu,
         ptr(mask), mask.stride(0) if mask is not None else 0, ptr(rv), ptr(cv), ptr(rscale),
         ptr(y16), y16.stride(0) if y16 is not None else 0, acc16, ptr(y32), y32.stride(0) if y32 is not None else 0,
         ptr(oscale), stream())


def _wgrad(dY, N, X, K, M, dW, col0, k_valid, oscale, dbias=None):
    """dW[:, col0:col0+k_valid] += dY^T X; with ``dbias`` also the bias gradient (column sums of dY)."""
    call("moda_tc_wgrad", ptr(dY), dY.stride(0), N, ptr(X), X.stride(0), K, M, ptr(dW) + 4 * col0, dW.stride(0),
         N, k_valid, ptr(oscale), ptr(dbias), stream())


def _pack(src, cols, col0, out, out_col0, out_rows, width, transpose):
    call("moda_pack16", ptr(src), src.stride(0), src.shape[0], cols, col0, ptr(out) + 2 * out_col0, None, None,
         out.stride(0), out_rows, width, int(transpose), stream())


class PackedWeights:
    """fp16 copies of the trunk weights in the layouts the tensor-core kernels read (K-major, 64-col chunks)."""

    def __init__(self, params, need_backward):
        dev = params[0].device
        h = lambda r, c: torch.empty(r, c, device=dev, dtype=HALF)
        W = [params[2 * i] for i in range(8)]
        Wf, Wd = params[16], params[18]
        self.fwd = []
        for i in range(8):
            if i == 0:
                t = h(256, 64)
                _pack(W[0], 63, 0, t, 0, 256, 64, False)
            elif i == 4:
                t = h(256, 320)
                _pack(W[4], 63, 0, t, 0, 256, 64, False)
                _pack(W[4], 256, 63, t, 64, 256, 256, False)
            else:
                t = h(256, 256)
                _pack(W[i], 256, 0, t, 0, 256, 256, False)
            self.fwd.append(t)
        self.Wf = h(256, 256)
        _pack(Wf, 256, 0, self.Wf, 0, 256, 256, False)
        self.Wd = h(128, 256)
        _pack(Wd, 256, 0, self.Wd, 0, 128, 256, False)
        if need_backward:
            # transposed copies: B[N = input channel, K = output channel] for the data-gradient GEMMs
            self.T = [None] * 8
            for i in range(1, 8):
                t = h(256, 256)
                _pack(W[i], 256, 63 if i == 4 else 0, t, 0, 256, 256, True)
                self.T[i] = t
            self.T_pe1 = h(64, 256)
            _pack(W[0], 63, 0, self.T_pe1, 0, 64, 256, True)
            self.T_pe5 = h(64, 256)
            _pack(W[4], 63, 0, self.T_pe5, 0, 64, 256, True)
            self.WfT = h(256, 256)
            _pack(Wf, 256, 0, self.WfT, 0, 256, 256, True)
            self.WdT = h(256, 128)
            _pack(Wd, 256, 0, self.WdT, 0, 256, 128, True)


class TrunkTcFn(torch.autograd.Function):
    """apply(xyz (P,3), dir_embedded (R,cd), env_code (R,ce) | None, S, win, *params) -> raw (P,4) [rgb | sigma]."""

    @staticmethod
    def forward(ctx, xyz, dir_emb, env, S, win, *params):
        xyz_shape = xyz.shape
        xyz = f32(xyz).reshape(-1, 3)
        P = xyz.shape[0]
        dev = xyz.device
        params = [f32(p) for p in params]
        need_bw = any(ctx.needs_input_grad)
        pk = PackedWeights(params, need_bw)
        b = [params[2 * i + 1] for i in range(8)]
        Wf, bf, Wd, bd, Ws, bs, Wr, br = params[16:24]
        code = f32(dir_emb) if env is None else torch.cat([f32(dir_emb), f32(env)], -1)
        R, cc = code.shape
        assert R * S == P and Wd.shape[1] == 256 + cc
        h16 = lambda n: torch.empty(P, n, device=dev, dtype=HALF)
        wa, nw = _win_array(win)
        F = len(win)
        A0 = h16(64)
        call("moda_pe16_fwd", ptr(xyz), ptr(A0), None, 64, P, F, wa, stream())
        # per-ray bias of the direction layer: Wd[:, 256:] [dir | env] + bd   (tiny fp32 GEMM, M = rays)
        rb = torch.empty(R, 128, device=dev, dtype=torch.float32)
        one = lambda v: (ctypes.c_int * 1)(v)
        call("moda_linear_fwd", R, 128, 1, (ctypes.c_void_p * 1)(ptr(code)), one(cc), one(cc), one(0), one(1), None, 0,
             ptr(Wd) + 4 * 256, Wd.shape[1], ptr(bd), 0, ptr(rb), 128, stream())
        H = []
        for i in range(8):
            y = h16(256)
            if i == 0:
                _tcl(A0, 64, None, 0, pk.fwd[0], P, 256, bias=b[0], relu=1, y16=y)
            elif i == 4:
                _tcl(A0, 64, H[3], 256, pk.fwd[4], P, 256, bias=b[4], relu=1, y16=y)
            else:
                _tcl(H[i - 1], 256, None, 0, pk.fwd[i], P, 256, bias=b[i], relu=1, y16=y)
            H.append(y)
        fin = h16(256)
        _tcl(H[7], 256, None, 0, pk.Wf, P, 256, bias=bf, y16=fin)
        dfe = h16(128)
        _tcl(fin, 256, None, 0, pk.Wd, P, 128, rowbias=rb, rep=S, relu=1, y16=dfe)
        raw = torch.empty(P, 4, device=dev, dtype=torch.float32)
        call("moda_head_fwd", ptr(H[7]), ptr(dfe), ptr(Ws), ptr(bs), ptr(Wr), ptr(br), ptr(raw), P, stream())
        if need_bw:
            ctx.save_for_backward(xyz, code, raw, *params)
            ctx.act = (A0, H, fin, dfe, pk)
            ctx.meta = (S, win, dir_emb.shape[-1], env is not None, xyz_shape)
        return raw

    @staticmethod
    def backward(ctx, graw):
        xyz, code, raw = ctx.saved_tensors[:3]
        params = list(ctx.saved_tensors[3:])
        A0, H, fin, dfe, pk = ctx.act
        S, win, cd, has_env, xyz_shape = ctx.meta
        P, dev = xyz.shape[0], xyz.device
        R, cc = code.shape
        b = [params[2 * i + 1] for i in range(8)]
        Wf, bf, Wd, bd, Ws, bs, Wr, br = params[16:24]
        g = [torch.zeros_like(p) for p in params]
        graw = f32(graw).reshape(P, 4)
        h16 = lambda n: torch.empty(P, n, device=dev, dtype=HALF)
        # dynamic power-of-two loss scale for the fp16 gradient chain
        scale2 = torch.empty(2, device=dev, dtype=torch.float32)
        work = torch.empty(1, device=dev, dtype=torch.int32)
        call("moda_loss_scale", ptr(graw), P * 4, 1024.0, work.data_ptr(), ptr(scale2), stream())
        sc, isc = scale2[0:1], scale2[1:2]
        # heads
        d_dfe = h16(128)
        gsig = torch.empty(P, device=dev, dtype=torch.float32)
        call("moda_head_bwd", ptr(H[7]), ptr(dfe), ptr(raw), ptr(graw), ptr(Wr), ptr(sc), ptr(d_dfe), ptr(gsig),
             ptr(g[22]), ptr(g[23]), ptr(g[20]), ptr(g[21]), P, stream())
        # direction layer: hoisted per-ray part (fp32, M = rays) ...
        grb = torch.empty(R, 128, device=dev, dtype=torch.float32)
        call("moda_segsum16", ptr(d_dfe), 128, ptr(grb), R, S, 128, ptr(isc), 0, stream())
        gcode = torch.empty(R, cc, device=dev, dtype=torch.float32)
        call("moda_linear_dgrad", R, 128, cc, ptr(grb), 128, ptr(Wd), Wd.shape[1], 256, None, 0, 0, ptr(gcode), cc,
             stream())
        one = lambda v: (ctypes.c_int * 1)(v)
        call("moda_linear_wgrad", R, 128, 1, (ctypes.c_void_p * 1)(ptr(code)), one(cc), one(cc), one(0), one(1), None, 0,
             ptr(grb), 128, ptr(g[18]), Wd.shape[1], 256, ptr(g[19]), stream())
        # ... and the per-sample part on tensor cores
        _wgrad(d_dfe, 128, fin, 256, P, g[18], 0, 256, isc)
        d_fin = h16(256)
        _tcl(d_dfe, 128, None, 0, pk.WdT, P, 256, y16=d_fin)
        # final layer (no activation) + sigma head's rank-1 data gradient, masked by relu(H8)
        _wgrad(d_fin, 256, H[7], 256, P, g[16], 0, 256, isc, dbias=g[17])
        bufs = [h16(256), d_fin]  # d_fin's storage is recycled once consumed
        dY = bufs[0]
        _tcl(d_fin, 256, None, 0, pk.WfT, P, 256, mask=H[7], rv=gsig, cv=Ws.reshape(-1), rscale=sc, y16=dY)
        cur = 0
        d_pe = h16(64)
        for i in range(7, -1, -1):
            dY = bufs[cur]
            if i == 0:
                _wgrad(dY, 256, A0, 64, P, g[0], 0, 63, isc, dbias=g[1])
                _tcl(dY, 256, None, 0, pk.T_pe1, P, 64, y16=d_pe, acc16=1)
            else:
                if i == 4:
                    _wgrad(dY, 256, A0, 64, P, g[8], 0, 63, isc)
                    _wgrad(dY, 256, H[3], 256, P, g[8], 63, 256, isc, dbias=g[9])
                    _tcl(dY, 256, None, 0, pk.T_pe5, P, 64, y16=d_pe)
                else:
                    _wgrad(dY, 256, H[i - 1], 256, P, g[2 * i], 0, 256, isc, dbias=g[2 * i + 1])
                nxt = 1 - cur
                _tcl(dY, 256, None, 0, pk.T[i], P, 256, mask=H[i - 1], y16=bufs[nxt])
                cur = nxt
        gxyz = torch.empty(P, 3, device=dev, dtype=torch.float32)
        wa, _ = _win_array(win)
        call("moda_pe16_bwd", ptr(xyz), ptr(d_pe), None, 64, ptr(gxyz), P, len(win), wa, ptr(isc), 0, stream())
        ctx.act = None
        gdir = gcode[:, :cd].contiguous()
        genv = gcode[:, cd:].contiguous() if has_env else None
        return (gxyz.reshape(xyz_shape), gdir, genv, None, None) + tuple(g)
