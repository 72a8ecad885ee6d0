"""Tensor-core execution of ``nerf_skin`` (5x64 MLP, nnutils/moda.py:325-329; NeRF.forward nerf.py:147-198 with
raw_feat=True, in_channels_dir=0) on the virtual input [PE(xyz) | pose code] of ``mlp_skinning``
(nnutils/geom_utils.py:219-229).

The delta skinning logits feed a softmax whose Gaussian logits are O(100) (geom_utils.py:265-266), so this MLP
needs fp32-class accuracy.  Every operand is therefore an fp16 (hi, lo) pair with
value = hi + lo and each layer evaluates  hi Whi^T + lo Whi^T + hi Wlo^T  as ONE tcgen05 GEMM over the
concatenated K dimension (csrc/tc_gemm.cu: moda_tc_linear_split / moda_tc_wgrad_split): ~22 mantissa bits.
The per-ray-constant pose code (128 of the 191 input columns of layers 1 and 5) is hoisted into a per-ray bias.
All activations are 64 columns wide (dir layer 32 and output 25 are zero-padded).
"""
import ctypes

import torch

from ._lib import call, ptr, stream, f32
from .ops import _win_array

HALF = torch.float16
WD = 64  # padded width of every activation


def supported(model, n_code):
    return (model.D == 5 and model.W == 64 and list(model.skips) == [4] and model.in_channels_dir == 0
            and model.in_channels_xyz == 63 + n_code and model.raw_feat and model.out_channels <= 32
            and model.dir_encoding[0].weight.shape[0] == 32)


def _pk3(src, cols, col0, B3, K, sub, width, out_rows, transpose):
    """Writes one sub-block of a [hi | hi | lo] operand: hi at column `sub`, its copy at K+sub, lo at 2K+sub."""
    base = ptr(B3)
    call("moda_pack16", ptr(src), src.stride(0), src.shape[0], cols, col0, base + 2 * sub, base + 2 * (K + sub),
         base + 2 * (2 * K + sub), B3.stride(0), out_rows, width, int(transpose), stream())


def _sl(A1, A2, B3, M, N, bias=None, rowbias=None, rep=1, relu=0, mask=None, out=None, y32=None, oscale=None):
    a2h, a2l = (A2 if A2 is not None else (None, None))
    yh, yl = (out if out is not None else (None, None))
    call("moda_tc_linear_split", ptr(A1[0]), ptr(A1[1]), A1[0].stride(0), WD, ptr(a2h), ptr(a2l),
         a2h.stride(0) if a2h is not None else 0, WD if a2h is not None else 0, ptr(B3), B3.stride(0), M, N,
         ptr(bias), ptr(rowbias), rep, relu, ptr(mask), mask.stride(0) if mask is not None else 0, ptr(yh), ptr(yl),
         yh.stride(0) if yh is not None else 0, 0, ptr(y32), y32.stride(0) if y32 is not None else 0, ptr(oscale),
         stream())


def _wg(dY, X, M, dW, col0, n_valid, k_valid, oscale):
    call("moda_tc_wgrad_split", ptr(dY[0]), ptr(dY[1]), dY[0].stride(0), WD, ptr(X[0]), ptr(X[1]), X[0].stride(0), WD, M,
         ptr(dW) + 4 * col0, dW.stride(0), n_valid, k_valid, ptr(oscale), stream())


class SkinPacked:
    def __init__(self, params, nc, need_backward):
        dev = params[0].device
        z = lambda r, c: torch.zeros(r, c, device=dev, dtype=HALF)
        W = [params[2 * i] for i in range(5)]
        Wf, Wd, Wr = params[10], params[12], params[16]
        oc = Wr.shape[0]
        self.L = []
        for i in range(5):
            if i == 0:
                t = z(64, 192)
                _pk3(W[0], 63, 0, t, 64, 0, 64, 64, False)
            elif i == 4:
                t = z(64, 384)
                _pk3(W[4], 63, 0, t, 128, 0, 64, 64, False)
                _pk3(W[4], 64, 63 + nc, t, 128, 64, 64, 64, False)
            else:
                t = z(64, 192)
                _pk3(W[i], 64, 0, t, 64, 0, 64, 64, False)
            self.L.append(t)
        self.Wf = z(64, 192)
        _pk3(Wf, 64, 0, self.Wf, 64, 0, 64, 64, False)
        self.Wd = z(32, 192)
        _pk3(Wd, 64, 0, self.Wd, 64, 0, 64, 32, False)
        self.Wr = z(32, 192)                       # K = 32 dfe channels padded to 64, N = oc padded to 32
        _pk3(Wr, 32, 0, self.Wr, 64, 0, 64, 32, False)
        if need_backward:
            self.T = [None] * 5                    # B[N = in channel, K = out channel]
            for i in range(1, 5):
                t = z(64, 192)
                _pk3(W[i], 64, (63 + nc) if i == 4 else 0, t, 64, 0, 64, 64, True)
                self.T[i] = t
            self.T_pe = z(64, 384)                 # [dY5 | dY1] -> dPE
            _pk3(W[4], 63, 0, self.T_pe, 128, 0, 64, 64, True)
            _pk3(W[0], 63, 0, self.T_pe, 128, 64, 64, 64, True)
            self.WfT = z(64, 192)
            _pk3(Wf, 64, 0, self.WfT, 64, 0, 64, 64, True)
            self.WdT = z(64, 192)                  # N = 64 fin channels, K = 32 dfe channels (padded)
            _pk3(Wd, 64, 0, self.WdT, 64, 0, 32, 64, True)
            self.WrT = z(32, 192)                  # N = 32 dfe channels, K = oc logits (padded)
            _pk3(Wr, 32, 0, self.WrT, 64, 0, oc, 32, True)


class SkinMlpTcFn(torch.autograd.Function):
    """apply(pts (..,3), code (Rc,nc) with Rc in {rays, 1}, S, win, *params) -> (P, 32) fp32 delta logits,
    columns >= out_channels are zero (a row pitch the skinning kernels accept directly)."""

    @staticmethod
    def forward(ctx, pts, code, S, win, *params):
        pshape = pts.shape
        pts = f32(pts).reshape(-1, 3)
        P, dev = pts.shape[0], pts.device
        params = [f32(p) for p in params]
        code = f32(code).reshape(-1, code.shape[-1])
        Rc, nc = code.shape
        rep = S if Rc * S == P else P
        assert Rc * rep == P, "pose code rows do not match the points"
        need_bw = any(ctx.needs_input_grad)
        pk = SkinPacked(params, nc, need_bw)
        b = [params[2 * i + 1] for i in range(5)]
        W = [params[2 * i] for i in range(5)]
        bf, bd, br = params[11], params[13], params[17]
        oc = br.shape[0]
        pair = lambda zero=False: tuple((torch.zeros if zero else torch.empty)(P, WD, device=dev, dtype=HALF) for _ in range(2))
        wa, _ = _win_array(win)
        A0 = pair()
        call("moda_pe16_fwd", ptr(pts), ptr(A0[0]), ptr(A0[1]), WD, P, len(win), wa, stream())
        one = lambda v: (ctypes.c_int * 1)(v)

        def code_bias(Wl, bl):
            rb = torch.empty(Rc, 64, device=dev, dtype=torch.float32)
            call("moda_linear_fwd", Rc, 64, 1, (ctypes.c_void_p * 1)(ptr(code)), one(nc), one(nc), one(0), one(1), None, 0,
                 ptr(Wl) + 4 * 63, Wl.shape[1], ptr(bl), 0, ptr(rb), 64, stream())
            return rb

        rb1, rb5 = code_bias(W[0], b[0]), code_bias(W[4], b[4])
        H = []
        for i in range(5):
            y = pair()
            if i == 0:
                _sl(A0, None, pk.L[0], P, 64, rowbias=rb1, rep=rep, relu=1, out=y)
            elif i == 4:
                _sl(A0, H[3], pk.L[4], P, 64, rowbias=rb5, rep=rep, relu=1, out=y)
            else:
                _sl(H[i - 1], None, pk.L[i], P, 64, bias=b[i], relu=1, out=y)
            H.append(y)
        fin = pair()
        _sl(H[4], None, pk.Wf, P, 64, bias=bf, out=fin)
        dfe = pair(zero=True)
        _sl(fin, None, pk.Wd, P, 32, bias=bd, relu=1, out=dfe)
        br32 = torch.zeros(32, device=dev, dtype=torch.float32)
        br32[:oc] = br
        out = torch.empty(P, 32, device=dev, dtype=torch.float32)
        _sl(dfe, None, pk.Wr, P, 32, bias=br32, y32=out)
        if need_bw:
            ctx.save_for_backward(pts, code, *params)
            ctx.act = (A0, H, fin, dfe, pk)
            ctx.meta = (S, win, rep, pshape, oc)
        return out

    @staticmethod
    def backward(ctx, gout):
        """Plain fp16 gradient chain (hi halves only) under the dynamic loss scale.  The forward needs fp32-class
        logits (they sit next to O(100) Gaussian logits inside a softmax); its adjoint only has to deliver
        gradients to ~1e-3 relative, which fp16 operands with fp32 accumulation do ("precision")."""
        pts, code = ctx.saved_tensors[:2]
        params = list(ctx.saved_tensors[2:])
        A0, H, fin, dfe, pk = ctx.act
        S, win, rep, pshape, oc = ctx.meta
        P, dev = pts.shape[0], pts.device
        Rc, nc = code.shape
        W = [params[2 * i] for i in range(5)]
        g = [torch.zeros_like(p) for p in params]
        gout = f32(gout).reshape(P, 32)
        h16 = lambda zero=False: (torch.zeros if zero else torch.empty)(P, WD, device=dev, dtype=HALF)
        scale2 = torch.empty(2, device=dev, dtype=torch.float32)
        work = torch.empty(1, device=dev, dtype=torch.int32)
        call("moda_loss_scale", ptr(gout), P * 32, 1024.0, work.data_ptr(), ptr(scale2), stream())
        sc, isc = scale2[0:1], scale2[1:2]
        G = h16()
        call("moda_split16", ptr(gout), 32, 32, ptr(sc), ptr(G), None, WD, WD, P, stream())
        one = lambda v: (ctypes.c_int * 1)(v)
        gcode = torch.zeros_like(code)

        def lin(A1, A2, B3, N, mask=None, out=None):
            """out (P,N) fp16 = [A1 | A2] Bhi^T (masked); Bhi = the leading (hi) block of a [hi | hi | lo] operand"""
            K2 = WD if A2 is not None else 0
            call("moda_tc_linear", ptr(A1), A1.stride(0), WD, ptr(A2), A2.stride(0) if A2 is not None else 0, K2,
                 ptr(B3), B3.stride(0), P, N, None, None, 1, 0, ptr(mask), mask.stride(0) if mask is not None else 0,
                 None, None, None, ptr(out), out.stride(0), 0, None, 0, None, stream())

        def wg(dY, X, dW, col0, n_valid, k_valid, dbias=None):
            call("moda_tc_wgrad", ptr(dY), dY.stride(0), WD, ptr(X), X.stride(0), WD, P, ptr(dW) + 4 * col0,
                 dW.stride(0), n_valid, k_valid, ptr(isc), ptr(dbias), stream())

        def code_part(dY, Wl, gW, gb):
            """hoisted pose-code columns of layers 1 / 5: everything happens at ray (or single-row) level in fp32"""
            rbg = torch.zeros(Rc, 64, device=dev, dtype=torch.float32)
            if Rc == 1:
                call("moda_colsum16", ptr(dY), WD, ptr(rbg), P, WD, ptr(isc), stream())
            else:
                call("moda_segsum16", ptr(dY), WD, ptr(rbg), Rc, rep, 64, ptr(isc), 0, stream())
            call("moda_linear_dgrad", Rc, 64, nc, ptr(rbg), 64, ptr(Wl), Wl.shape[1], 63, None, 0, 1, ptr(gcode), nc,
                 stream())
            call("moda_linear_wgrad", Rc, 64, 1, (ctypes.c_void_p * 1)(ptr(code)), one(nc), one(nc), one(0), one(1), None,
                 0, ptr(rbg), 64, ptr(gW), Wl.shape[1], 63, ptr(gb), stream())

        # output layer (oc logits <- 32 dfe channels)
        wg(G, dfe[0], g[16], 0, oc, 32, dbias=g[17])
        d_dfe = h16(zero=True)
        lin(G, None, pk.WrT, 32, mask=dfe[0], out=d_dfe)
        # dir layer (32 <- 64)
        wg(d_dfe, fin[0], g[12], 0, 32, 64, dbias=g[13])
        d_fin = h16()
        lin(d_dfe, None, pk.WdT, 64, out=d_fin)
        # final layer (64 <- 64, no activation)
        wg(d_fin, H[4][0], g[10], 0, 64, 64, dbias=g[11])
        dY5 = h16()
        lin(d_fin, None, pk.WfT, 64, mask=H[4][0], out=dY5)
        # layer 5: [PE | code | h4]
        wg(dY5, A0[0], g[8], 0, 64, 63)
        wg(dY5, H[3][0], g[8], 63 + nc, 64, 64)
        code_part(dY5, W[4], g[8], g[9])
        cur, free = d_fin, h16()
        lin(dY5, None, pk.T[4], 64, mask=H[3][0], out=cur)
        for i in (3, 2, 1):
            wg(cur, H[i - 1][0], g[2 * i], 0, 64, 64, dbias=g[2 * i + 1])
            lin(cur, None, pk.T[i], 64, mask=H[i - 1][0], out=free)
            cur, free = free, cur
        dY1 = cur
        wg(dY1, A0[0], g[0], 0, 64, 63)
        code_part(dY1, W[0], g[0], g[1])
        d_pe = free
        lin(dY5, dY1, pk.T_pe, 64, out=d_pe)
        gpts = torch.empty(P, 3, device=dev, dtype=torch.float32)
        wa, _ = _win_array(win)
        call("moda_pe16_bwd", ptr(pts), ptr(d_pe), None, WD, ptr(gpts), P, len(win), wa, ptr(isc), 0, stream())
        ctx.act = None
        # the sigma head of nerf_skin is computed and discarded in the reference (nerf.py:178): no gradient
        g[14] = g[15] = None
        return (gpts.reshape(pshape), gcode, None, None) + tuple(g)
