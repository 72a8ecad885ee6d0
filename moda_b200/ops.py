"""torch.autograd.Function wrappers around the C ABI (one per CUDA op, forward + hand-written backward).

These are the only places that touch device pointers.  Each Function mirrors a reference function (cited
in its docstring); the public, reference-named API lives in nerf.py / dual_quat.py / geom_utils.py /
rendering.py and is built from these.
"""
import ctypes

import functools
import torch

from . import _lib
from ._lib import call, ptr, stream, f32

SEG_DENSE, SEG_BCAST, SEG_PE = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 2
DQ_QCONJ, DQ_CCONJ, DQ_NORMALIZE, DQ_INVERSE, Q_NORMALIZE = 0, 1, 2, 3, 4


def _win_array(win):
    if win is None or len(win) == 0:
        return None, 0
    arr = (ctypes.c_float * len(win))(*[float(w) for w in win])
    return arr, len(win)


def pe_window(n_freqs, alpha):
    """nerf.py:63-66, evaluated in fp32 exactly like the reference does (torch ops on a CPU tensor).  A pure function of two
    numbers asked for several times per step: memoised."""
    if n_freqs <= 0:
        return []
    return list(_pe_window(int(n_freqs), float(n_freqs if alpha is None else alpha)))


@functools.lru_cache(maxsize=64)
def _pe_window(n_freqs, alpha):
    import math
    a = torch.as_tensor(alpha, dtype=torch.float32).cpu()
    w = a - torch.arange(n_freqs, dtype=torch.float32)
    w = torch.clamp(w, 0.0, 1.0)
    w = 0.5 * (1 + torch.cos(math.pi * w + math.pi))
    return tuple(float(x) for x in w)


class _P:
    """(pointer, leading dimension) pair with element-offset arithmetic."""

    def __init__(self, t, off=0, ld=None):
        self.t = t
        self.addr = ptr(t) + 4 * off
        self.ld = ld if ld is not None else (t.shape[-1] if t.dim() > 1 else 1)


class Seg:
    """One column segment of the virtual A operand of a linear layer (see csrc/gemm.cu)."""

    def __init__(self, kind, tensor, width, aux=1, col=0, ld=None):
        self.kind, self.tensor, self.width, self.aux, self.col = kind, tensor, width, aux, col
        self.ld = ld if ld is not None else tensor.shape[-1]


def _seg_arrays(segs):
    n = len(segs)
    ptrs = (ctypes.c_void_p * n)(*[ptr(s.tensor) + 4 * s.col for s in segs])
    lds = (ctypes.c_int * n)(*[s.ld for s in segs])
    wid = (ctypes.c_int * n)(*[s.width for s in segs])
    typ = (ctypes.c_int * n)(*[s.kind for s in segs])
    aux = (ctypes.c_int * n)(*[s.aux for s in segs])
    return n, ptrs, lds, wid, typ, aux


def linear_fwd(M, N, segs, win, W, bias, act, Y):
    """Y (_P) = act(A W^T + b); A assembled from ``segs``."""
    n, ptrs, lds, wid, typ, aux = _seg_arrays(segs)
    wa, nw = _win_array(win)
    call("moda_linear_fwd", M, N, n, ptrs, lds, wid, typ, aux, wa, nw, ptr(W), W.shape[1], ptr(bias), act,
         Y.addr, Y.ld, stream())


def linear_dgrad(M, N, K, dY, W, k0, mask, accumulate, dA):
    """dA (_P, M x K) (=|+=) dY (_P, M x N) @ W[:, k0:k0+K], then relu-masked by ``mask`` (_P) if given."""
    call("moda_linear_dgrad", M, N, K, dY.addr, dY.ld, ptr(W), W.shape[1], k0,
         mask.addr if mask is not None else None, mask.ld if mask is not None else 0, int(accumulate),
         dA.addr, dA.ld, stream())


def linear_wgrad(M, N, segs, win, dY, dW, dbias, k0=0):
    n, ptrs, lds, wid, typ, aux = _seg_arrays(segs)
    wa, nw = _win_array(win)
    call("moda_linear_wgrad", M, N, n, ptrs, lds, wid, typ, aux, wa, nw, dY.addr, dY.ld, ptr(dW), dW.shape[1], k0,
         ptr(dbias), stream())


# --------------------------------------------------------------------------------------------------
class EmbedFn(torch.autograd.Function):
    """Embedding.forward (nnutils/nerf.py:35-75)."""

    @staticmethod
    def forward(ctx, x, n_freqs, win):
        shape = x.shape
        C = shape[-1]
        x2 = f32(x).reshape(-1, C)
        M = x2.shape[0]
        out = torch.empty(M, C * (1 + 2 * n_freqs), device=x.device, dtype=torch.float32)
        wa, _ = _win_array(win)
        call("moda_embed_fwd", ptr(x2), C, ptr(out), out.shape[1], M, C, n_freqs, wa, stream())
        ctx.save_for_backward(x2)
        ctx.meta = (n_freqs, win, shape)
        return out.reshape(shape[:-1] + (out.shape[1],))

    @staticmethod
    def backward(ctx, g):
        (x2,) = ctx.saved_tensors
        n_freqs, win, shape = ctx.meta
        C = shape[-1]
        g2 = f32(g).reshape(x2.shape[0], -1)
        gx = torch.empty_like(x2)
        wa, _ = _win_array(win)
        call("moda_embed_bwd", ptr(x2), C, ptr(g2), g2.shape[1], ptr(gx), C, x2.shape[0], C, n_freqs, wa, 0,
             stream())
        return gx.reshape(shape), None, None


# --------------------------------------------------------------------------------------------------
class MlpSpec:
    """Architecture of one ``NeRF`` (nnutils/nerf.py:84-136) plus how its input is assembled.

    ``xyz_segs`` / ``dir_segs`` are lists of (kind, tensor_index, width, aux, col) describing the columns of
    the reference's ``input_xyz`` / ``input_dir`` in terms of the tensors handed to ``MlpFn.apply``.
    """

    def __init__(self, D, W, out_channels, skips, raw_feat, sigma_only, xyz_segs, dir_segs, win, n_inputs,
                 dense_cols=None):
        self.D, self.W, self.out_channels, self.skips = D, W, out_channels, tuple(skips)
        self.raw_feat, self.sigma_only = raw_feat, sigma_only
        self.xyz_segs, self.dir_segs, self.win, self.n_inputs = xyz_segs, dir_segs, win, n_inputs
        self.dense_cols = dense_cols or {}


def _mk_segs(desc, inputs):
    segs = []
    for (kind, idx, width, aux, col) in desc:
        t = inputs[idx]
        segs.append(Seg(kind, t, width, aux, col, t.shape[-1]))
    return segs


class MlpFn(torch.autograd.Function):
    """NeRF.forward (nnutils/nerf.py:147-198) with evaluate_mlp's input assembly (geom_utils.py:19-57).

    apply(spec, M, *inputs, *params); params = [W1,b1,...,WD,bD, Wf,bf, Wd,bd, Ws,bs, Wr,br].
    Returns (M,4) [rgb|sigma], or (M,1) sigma when spec.sigma_only, or (M,out) raw features.
    """

    @staticmethod
    def forward(ctx, spec, M, *tensors):
        inputs = [f32(t) for t in tensors[:spec.n_inputs]]
        params = [f32(t) for t in tensors[spec.n_inputs:]]
        D, W = spec.D, spec.W
        dev = params[0].device
        xyz = _mk_segs(spec.xyz_segs, inputs)
        dirs = _mk_segs(spec.dir_segs, inputs)
        acts = []
        h = None
        for i in range(D):
            if i == 0:
                segs = xyz
            elif i in spec.skips:
                segs = xyz + [Seg(SEG_DENSE, h, W)]
            else:
                segs = [Seg(SEG_DENSE, h, W)]
            out = torch.empty(M, W, device=dev, dtype=torch.float32)
            linear_fwd(M, W, segs, spec.win, params[2 * i], params[2 * i + 1], ACT_RELU, _P(out))
            acts.append(out)
            h = out
        Wf, bf, Wd, bd, Ws, bs, Wr, br = params[2 * D:2 * D + 8]
        fin = dfe = None
        if spec.sigma_only:
            res = torch.empty(M, 1, device=dev, dtype=torch.float32)
            linear_fwd(M, 1, [Seg(SEG_DENSE, h, W)], None, Ws, bs, ACT_NONE, _P(res))
        else:
            fin = torch.empty(M, W, device=dev, dtype=torch.float32)
            linear_fwd(M, W, [Seg(SEG_DENSE, h, W)], None, Wf, bf, ACT_NONE, _P(fin))
            Wh = Wd.shape[0]
            dfe = torch.empty(M, Wh, device=dev, dtype=torch.float32)
            linear_fwd(M, Wh, [Seg(SEG_DENSE, fin, W)] + dirs, spec.win, Wd, bd, ACT_RELU, _P(dfe))
            oc = spec.out_channels
            if spec.raw_feat:
                res = torch.empty(M, oc, device=dev, dtype=torch.float32)
                linear_fwd(M, oc, [Seg(SEG_DENSE, dfe, Wh)], None, Wr, br, ACT_NONE, _P(res))
            else:
                res = torch.empty(M, oc + 1, device=dev, dtype=torch.float32)
                linear_fwd(M, oc, [Seg(SEG_DENSE, dfe, Wh)], None, Wr, br, ACT_SIGMOID, _P(res, 0, oc + 1))
                linear_fwd(M, 1, [Seg(SEG_DENSE, h, W)], None, Ws, bs, ACT_NONE, _P(res, oc, oc + 1))
        ctx.spec, ctx.M = spec, M
        ctx.n_in = len(inputs)
        ctx.save_for_backward(*(inputs + params + acts + [t for t in (fin, dfe, res) if t is not None]))
        ctx.has_tail = fin is not None
        return res

    @staticmethod
    def backward(ctx, g):
        spec, M = ctx.spec, ctx.M
        D, W = spec.D, spec.W
        saved = list(ctx.saved_tensors)
        inputs = saved[:ctx.n_in]
        params = saved[ctx.n_in:ctx.n_in + 2 * D + 8]
        rest = saved[ctx.n_in + 2 * D + 8:]
        acts = rest[:D]
        if ctx.has_tail:
            fin, dfe, res = rest[D:D + 3]
        else:
            res = rest[D]
        dev = g.device
        g = f32(g)
        xyz = _mk_segs(spec.xyz_segs, inputs)
        dirs = _mk_segs(spec.dir_segs, inputs)
        Wf, bf, Wd, bd, Ws, bs, Wr, br = params[2 * D:2 * D + 8]
        gparams = [torch.zeros_like(p) for p in params]
        need_in = [ctx.needs_input_grad[2 + i] for i in range(ctx.n_in)]
        gin = [None] * ctx.n_in
        pe_tmp = {}

        def seg_dgrad(seg_desc, seg, dY, N, Wt, k0):
            """gradient of one input segment of a layer whose (masked) output gradient is dY (M x N)."""
            kind, idx, width, aux, col = seg_desc
            if not need_in[idx]:
                return
            if kind == SEG_DENSE:
                if gin[idx] is None:
                    gin[idx] = torch.zeros_like(inputs[idx])
                linear_dgrad(M, N, width, dY, Wt, k0, None, True, _P(gin[idx], col, inputs[idx].shape[-1]))
            elif kind == SEG_BCAST:
                R = inputs[idx].shape[0]
                red = torch.empty(R, N, device=dev, dtype=torch.float32)
                call("moda_segsum", dY.addr, dY.ld, ptr(red), R, aux, N, stream())
                if gin[idx] is None:
                    gin[idx] = torch.zeros_like(inputs[idx])
                linear_dgrad(R, N, width, _P(red), Wt, k0, None, True, _P(gin[idx], col, inputs[idx].shape[-1]))
            else:  # SEG_PE: gradient of the encoding, folded back onto the points
                if idx not in pe_tmp:
                    pe_tmp[idx] = torch.empty(M, width, device=dev, dtype=torch.float32)
                tmp = pe_tmp[idx]
                linear_dgrad(M, N, width, dY, Wt, k0, None, False, _P(tmp))
                first = gin[idx] is None
                if first:
                    gin[idx] = torch.empty_like(inputs[idx])
                C = aux
                F = (width // C - 1) // 2
                wa, _ = _win_array(spec.win)
                call("moda_embed_bwd", ptr(inputs[idx]), inputs[idx].shape[-1], ptr(tmp), width, ptr(gin[idx]),
                     inputs[idx].shape[-1], M, C, F, wa, 0 if first else 1, stream())

        bufs = [torch.empty(M, W, device=dev, dtype=torch.float32) for _ in range(2)]
        hD = acts[D - 1]
        gh = _P(bufs[0])  # gradient w.r.t. the pre-activation of layer D (after masking)
        if spec.sigma_only:
            gsig = _P(g, 0, 1)
            linear_wgrad(M, 1, [Seg(SEG_DENSE, hD, W)], None, gsig, gparams[2 * D + 4], gparams[2 * D + 5])
            linear_dgrad(M, 1, W, gsig, Ws, 0, _P(hD), False, gh)
        else:
            oc = spec.out_channels
            Wh = Wd.shape[0]
            if spec.raw_feat:
                grgb = _P(g, 0, oc)
            else:
                t = torch.empty(M, oc, device=dev, dtype=torch.float32)
                call("moda_act_bwd", ACT_SIGMOID, ptr(res), oc + 1, ptr(g), oc + 1, ptr(t), oc, M, oc, stream())
                grgb = _P(t)
            # rgb layer
            linear_wgrad(M, oc, [Seg(SEG_DENSE, dfe, Wh)], None, grgb, gparams[2 * D + 6], gparams[2 * D + 7])
            gdfe = torch.empty(M, Wh, device=dev, dtype=torch.float32)
            linear_dgrad(M, oc, Wh, grgb, Wr, 0, _P(dfe), False, _P(gdfe))
            # dir layer
            linear_wgrad(M, Wh, [Seg(SEG_DENSE, fin, W)] + dirs, spec.win, _P(gdfe), gparams[2 * D + 2],
                         gparams[2 * D + 3])
            gfin = _P(bufs[1])
            linear_dgrad(M, Wh, W, _P(gdfe), Wd, 0, None, False, gfin)
            k0 = W
            for desc, seg in zip(spec.dir_segs, dirs):
                seg_dgrad(desc, seg, _P(gdfe), Wh, Wd, k0)
                k0 += seg.width
            # final layer (no activation)
            linear_wgrad(M, W, [Seg(SEG_DENSE, hD, W)], None, gfin, gparams[2 * D], gparams[2 * D + 1])
            if spec.raw_feat:
                # nerf_skin: the sigma head is computed and discarded in the reference (nerf.py:178)
                linear_dgrad(M, W, W, gfin, Wf, 0, _P(hD), False, gh)
            else:
                linear_dgrad(M, W, W, gfin, Wf, 0, None, False, gh)
                gsig = _P(g, oc, oc + 1)
                linear_wgrad(M, 1, [Seg(SEG_DENSE, hD, W)], None, gsig, gparams[2 * D + 4], gparams[2 * D + 5])
                linear_dgrad(M, 1, W, gsig, Ws, 0, _P(hD), True, gh)
        cur = 0  # bufs[cur] holds dY of layer i
        for i in range(D - 1, -1, -1):
            dY = _P(bufs[cur])
            if i == 0:
                segs = xyz
            elif i in spec.skips:
                segs = xyz + [Seg(SEG_DENSE, acts[i - 1], W)]
            else:
                segs = [Seg(SEG_DENSE, acts[i - 1], W)]
            linear_wgrad(M, W, segs, spec.win, dY, gparams[2 * i], gparams[2 * i + 1])
            if i == 0 or i in spec.skips:
                k0 = 0
                for desc, seg in zip(spec.xyz_segs, xyz):
                    seg_dgrad(desc, seg, dY, W, params[2 * i], k0)
                    k0 += seg.width
            if i > 0:
                k0 = sum(s.width for s in xyz) if i in spec.skips else 0
                nxt = 1 - cur
                linear_dgrad(M, W, W, dY, params[2 * i], k0, _P(acts[i - 1]), False, _P(bufs[nxt]))
                cur = nxt
        if spec.raw_feat:  # the unused sigma head gets no gradient in the reference either (nerf.py:178)
            gparams[2 * D + 4] = gparams[2 * D + 5] = None
        return (None, None) + tuple(gin) + tuple(gparams)


# --------------------------------------------------------------------------------------------------
class BoneTransformFn(torch.autograd.Function):
    """bone_transform, neudbs branch (nnutils/geom_utils.py:59-111)."""

    @staticmethod
    def forward(ctx, bones, rts):
        B = bones.shape[-2]
        b2 = f32(bones).reshape(-1, B, 10)
        r2 = f32(rts).reshape(-1, B, 8)
        R = r2.shape[0]
        per_ray = int(b2.shape[0] != 1)
        if per_ray and b2.shape[0] != R:
            raise RuntimeError("bone_transform: %d bone sets for %d transforms" % (b2.shape[0], R))
        out = torch.empty(R, B, 10, device=r2.device, dtype=torch.float32)
        call("moda_bone_transform_fwd", ptr(b2), ptr(r2), ptr(out), R, B, per_ray, stream())
        ctx.save_for_backward(b2, r2)
        ctx.meta = (bones.shape, rts.shape, per_ray)
        return out

    @staticmethod
    def backward(ctx, g):
        b2, r2 = ctx.saved_tensors
        bshape, rshape, per_ray = ctx.meta
        R, B = r2.shape[0], r2.shape[1]
        gb = torch.zeros_like(b2)
        gr = torch.empty_like(r2)
        call("moda_bone_transform_bwd", ptr(b2), ptr(r2), ptr(f32(g)), ptr(gb), ptr(gr), R, B, per_ray, stream())
        return gb.reshape(bshape), gr.reshape(rshape)


GCOPIES = 64  # replicated bone-gradient accumulators of the skin-warp adjoint


class SkinWarpFn(torch.autograd.Function):
    """Gaussian skinning weights and/or dual-quaternion blend skinning of (R,S,3) points.

    One kernel pair covers skinning() (geom_utils.py:237-302), dqs_blend_skinning() (:457-517) and the fused
    backward / forward warps of neu_dbs() (:372-456).  mode flags: deform (bones moved by rts first), invert
    (blend with dq_inverse(rts)); outputs chosen by want_y / want_skin.
    """

    @staticmethod
    def forward(ctx, pts, bones, rts, skin_aux, dskin, skin_in, deform, invert, want_y, want_skin):
        pts = f32(pts)
        R, S, _ = pts.shape
        B = bones.shape[-2]
        bones = f32(bones)
        per_ray = 0
        if bones.dim() == 3:
            if bones.shape[0] == R:
                per_ray = 1
            elif bones.shape[0] != 1:
                raise RuntimeError("skinning: %d bone sets for %d rays" % (bones.shape[0], R))
        rts_ = f32(rts).reshape(R, B, 8) if rts is not None else None
        aux = f32(skin_aux) if skin_aux is not None else torch.zeros(2, device=pts.device)
        dsk = f32(dskin) if dskin is not None else None
        ldd = 0
        if dsk is not None:
            # delta logits may arrive zero-padded to a wider row pitch (the tensor-core nerf_skin path writes 32
            # columns for 25 bones): the kernels take the pitch, nothing is compacted
            if dsk.shape[-1] < B:
                raise RuntimeError("dskin has %d channels for %d bones" % (dsk.shape[-1], B))
            if dsk.shape[-1] > B:
                ldd = dsk.shape[-1]
        sin = f32(skin_in) if skin_in is not None else None
        y = torch.empty_like(pts) if want_y else None
        skin = torch.empty(R, S, B, device=pts.device, dtype=torch.float32) if want_skin else None
        call("moda_skin_warp_fwd", ptr(pts), ptr(bones), ptr(rts_), ptr(aux), ptr(dsk), ptr(sin), ptr(y),
             ptr(skin), R, S, B, ldd, per_ray, int(deform), int(invert), stream())
        ctx.save_for_backward(pts, bones, rts_, aux, dsk, sin)
        ctx.meta = (R, S, B, per_ray, int(deform), int(invert), rts.shape if rts is not None else None,
                    skin_aux is not None, ldd)
        return y, skin

    @staticmethod
    def backward(ctx, gy, gskin):
        pts, bones, rts_, aux, dsk, sin = ctx.saved_tensors
        R, S, B, per_ray, deform, invert, rshape, has_aux, ldd = ctx.meta
        gy = f32(gy) if gy is not None else None
        gskin = f32(gskin) if gskin is not None else None
        gpts = torch.empty_like(pts)
        gdsk = torch.empty_like(dsk) if dsk is not None else None  # the kernel zeroes the pad columns of pitched rows
        gsin = torch.empty_like(sin) if sin is not None else None
        grts = torch.zeros_like(rts_) if rts_ is not None else None
        # bones shared by every ray: the rays' contributions go to GCOPIES replicated accumulators that are summed
        # afterwards (same-address L2 atomics from 8192 CTAs serialise: see include/moda_b200.h)
        copies = 1 if per_ray else max(1, min(R, GCOPIES))
        gbones = torch.zeros_like(bones) if per_ray else torch.zeros((copies,) + tuple(bones.shape), device=pts.device,
                                                                      dtype=torch.float32)
        gaux = torch.zeros(copies, 2, device=pts.device, dtype=torch.float32)
        call("moda_skin_warp_bwd", ptr(pts), ptr(bones), ptr(rts_), ptr(aux), ptr(dsk), ptr(sin), ptr(gy),
             ptr(gskin), ptr(gpts), ptr(gdsk), ptr(gsin), ptr(grts), ptr(gbones), ptr(gaux), R, S, B, ldd, per_ray,
             deform, invert, copies, stream())
        if not per_ray:
            gbones = gbones.sum(0)
        gaux = gaux.sum(0)
        return (gpts, gbones, grts.reshape(rshape) if grts is not None else None, gaux if has_aux else None,
                gdsk, gsin, None, None, None, None)


class DqUnaryFn(torch.autograd.Function):
    """dq_normalize / dq_inverse / conjugates / q_normalize (nnutils/dual_quat.py)."""

    @staticmethod
    def forward(ctx, x, op):
        x2 = f32(x)
        width = 4 if op == Q_NORMALIZE else 8
        if x2.shape[-1] != width:
            raise AssertionError("expected last dimension %d" % width)
        out = torch.empty_like(x2)
        call("moda_dq_unary_fwd", op, ptr(x2), ptr(out), x2.numel() // width, stream())
        ctx.save_for_backward(x2)
        ctx.op = op
        return out

    @staticmethod
    def backward(ctx, g):
        (x2,) = ctx.saved_tensors
        width = 4 if ctx.op == Q_NORMALIZE else 8
        gin = torch.empty_like(x2)
        call("moda_dq_unary_bwd", ctx.op, ptr(x2), ptr(f32(g)), ptr(gin), x2.numel() // width, stream())
        return gin, None


class DqMulFn(torch.autograd.Function):
    """q_mul / dq_mul (nnutils/dual_quat.py:14-49)."""

    @staticmethod
    def forward(ctx, a, b, width):
        a2, b2 = f32(a), f32(b)
        if a2.shape != b2.shape or a2.shape[-1] != width:
            raise AssertionError("q_mul/dq_mul expect two equally-sized [*, %d] tensors" % width)
        out = torch.empty_like(a2)
        call("moda_dq_mul_fwd", ptr(a2), ptr(b2), ptr(out), a2.numel() // width, width, stream())
        ctx.save_for_backward(a2, b2)
        ctx.width = width
        return out

    @staticmethod
    def backward(ctx, g):
        a2, b2 = ctx.saved_tensors
        ga, gb = torch.empty_like(a2), torch.empty_like(b2)
        call("moda_dq_mul_bwd", ptr(a2), ptr(b2), ptr(f32(g)), ptr(ga), ptr(gb), a2.numel() // ctx.width, ctx.width,
             stream())
        return ga, gb, None


class SampleRaysFn(torch.autograd.Function):
    """Depth sampling and point generation (nnutils/rendering.py:64-89).  Returns z (R,S), xyz (R,S,3),
    dn (R,3) = rays_d/|rays_d|.  near/far gradients are not propagated (they are in no optimizer group,
    nnutils/train_utils.py:177-222)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, near, far, jitter, perturb, use_disp, S):
        o, d = f32(rays_o), f32(rays_d)
        R = d.shape[0]
        nr, fr = f32(near).reshape(R), f32(far).reshape(R)
        jit = f32(jitter) if jitter is not None else None
        z = torch.empty(R, S, device=d.device, dtype=torch.float32)
        xyz = torch.empty(R, S, 3, device=d.device, dtype=torch.float32)
        dn = torch.empty(R, 3, device=d.device, dtype=torch.float32)
        call("moda_sample_rays_fwd", ptr(o), ptr(d), ptr(nr), ptr(fr), ptr(jit), float(perturb), int(use_disp),
             ptr(z), ptr(xyz), ptr(dn), R, S, stream())
        ctx.save_for_backward(d, z)
        ctx.mark_non_differentiable(z)
        return z, xyz, dn

    @staticmethod
    def backward(ctx, gz, gxyz, gdn):
        d, z = ctx.saved_tensors
        R, S = z.shape
        go, gd = torch.empty_like(d), torch.empty_like(d)
        call("moda_sample_rays_bwd", ptr(d), ptr(z), ptr(f32(gxyz)) if gxyz is not None else None,
             ptr(f32(gdn)) if gdn is not None else None, None, ptr(go), ptr(gd), R, S, stream())
        return go, gd, None, None, None, None, None, None


class PointsFromDepthsFn(torch.autograd.Function):
    """xyz = o + d * z for given (detached) depths (nnutils/rendering.py:112-113)."""

    @staticmethod
    def forward(ctx, rays_o, rays_d, z):
        o, d, z = f32(rays_o), f32(rays_d), f32(z)
        R, S = z.shape
        xyz = torch.empty(R, S, 3, device=d.device, dtype=torch.float32)
        call("moda_points_from_depths", ptr(o), ptr(d), ptr(z), ptr(xyz), R, S, stream())
        ctx.save_for_backward(d, z)
        return xyz

    @staticmethod
    def backward(ctx, gxyz):
        d, z = ctx.saved_tensors
        R, S = z.shape
        go, gd = torch.empty_like(d), torch.empty_like(d)
        call("moda_sample_rays_bwd", ptr(d), ptr(z), ptr(f32(gxyz)), None, None, ptr(go), ptr(gd), R, S, stream())
        return go, gd, None


class CompositeFn(torch.autograd.Function):
    """Density -> alpha -> transmittance -> rgb / depth / silhouette sums (nnutils/rendering.py:183-235) and
    the cycle term sum_s |xa - xb| * w.detach() (:341, :473).

    raw: (R,S,4) [rgb|sigma].  Returns rgb (R,3), depth (R), sil (R), weights (R,S), visibility (R,S), cyc (R).
    """

    @staticmethod
    def forward(ctx, raw, z, rays_d, beta, noise, mask, xa, xb):
        raw, z, d, beta = f32(raw), f32(z), f32(rays_d), f32(beta)
        R, S = z.shape
        C = raw.shape[-1]
        dev = z.device
        noise = f32(noise) if noise is not None else None
        mask = mask.to(torch.uint8).contiguous() if mask is not None else None
        xa = f32(xa) if xa is not None else None
        xb = f32(xb) if xb is not None else None
        o_rgb = torch.empty(R, 3, device=dev)
        o_dep, o_sil = torch.empty(R, device=dev), torch.empty(R, device=dev)
        o_w, o_vis = torch.empty(R, S, device=dev), torch.empty(R, S, device=dev)
        o_cyc = torch.empty(R, device=dev) if xa is not None else None
        call("moda_composite_fwd", ptr(raw), C, ptr(raw) + 4 * (C - 1), C, ptr(z), ptr(d), ptr(beta), ptr(noise),
             ptr(mask), ptr(xa), ptr(xb), ptr(o_rgb), ptr(o_dep), ptr(o_sil), ptr(o_w), ptr(o_vis), ptr(o_cyc), R, S,
             stream())
        ctx.save_for_backward(raw, z, d, beta, noise, mask, xa, xb, o_vis)
        ctx.mark_non_differentiable(o_vis)
        if o_cyc is None:
            o_cyc = torch.zeros(R, device=dev)
        return o_rgb, o_dep, o_sil, o_w, o_vis, o_cyc

    @staticmethod
    def backward(ctx, g_rgb, g_dep, g_sil, g_w, g_vis, g_cyc):
        raw, z, d, beta, noise, mask, xa, xb, vis = ctx.saved_tensors
        R, S = z.shape
        C = raw.shape[-1]
        dev = z.device
        c = lambda t: f32(t) if t is not None else None
        g_raw = torch.empty_like(raw)
        g_beta = torch.zeros(1, device=dev)
        g_nd = torch.empty(R, device=dev)
        g_xa = torch.empty_like(xa) if xa is not None else None
        g_xb = torch.empty_like(xb) if xb is not None else None
        call("moda_composite_bwd", ptr(raw), C, ptr(raw) + 4 * (C - 1), C, ptr(z), ptr(d), ptr(beta), ptr(noise),
             ptr(mask), ptr(xa), ptr(xb), ptr(vis), ptr(c(g_rgb)), ptr(c(g_dep)), ptr(c(g_sil)),
             ptr(c(g_cyc)) if xa is not None else None, ptr(c(g_w)), ptr(g_raw), C, ptr(g_raw) + 4 * (C - 1), C,
             ptr(g_beta), ptr(g_nd), ptr(g_xa), ptr(g_xb), R, S, stream())
        # d|d|/dd: handled here with the same sample_rays_bwd kernel (only its |d| branch is active)
        gd = torch.empty_like(d)
        call("moda_sample_rays_bwd", ptr(d), ptr(z), None, None, ptr(g_nd), None, ptr(gd), R, S, stream())
        return g_raw, None, gd, g_beta.reshape(beta.shape), None, None, g_xa, g_xb


class WeightedSumFn(torch.autograd.Function):
    """out (R,C) = sum_s w (R,S) * v (R,S,C): feature compositing (nnutils/rendering.py:233)."""

    @staticmethod
    def forward(ctx, w, v):
        w, v = f32(w), f32(v)
        R, S = w.shape
        C = v.shape[-1]
        out = torch.empty(R, C, device=w.device, dtype=torch.float32)
        call("moda_wsum_fwd", ptr(w), ptr(v), ptr(out), R, S, C, stream())
        ctx.save_for_backward(w, v)
        return out

    @staticmethod
    def backward(ctx, g):
        w, v = ctx.saved_tensors
        R, S = w.shape
        C = v.shape[-1]
        gw = torch.empty_like(w) if ctx.needs_input_grad[0] else None
        gv = torch.empty_like(v) if ctx.needs_input_grad[1] else None
        call("moda_wsum_bwd", ptr(w), ptr(v), ptr(f32(g)), ptr(gw), ptr(gv), R, S, C, stream())
        return gw, gv


class SamplePdfFn(torch.autograd.Function):
    """sample_pdf (nnutils/rendering.py:582-623); with ``z_vals`` also the sorted union of rendering.py:103-110.
    Outputs are detached in the reference (:105-106), so there is no backward."""

    @staticmethod
    def forward(ctx, bins, weights, n_importance, det, eps, u, z_vals):
        w = f32(weights)
        R = w.shape[0]
        uu = f32(u) if u is not None else None
        if z_vals is not None:
            z = f32(z_vals)
            S = z.shape[1]
            out = torch.empty(R, S + n_importance, device=w.device, dtype=torch.float32)
            call("moda_sample_pdf", ptr(z), None, ptr(w), ptr(uu), ptr(out), R, S, S - 2, n_importance, int(det),
                 float(eps), 1, stream())
        else:
            b = f32(bins)
            n = w.shape[1]
            out = torch.empty(R, n_importance, device=w.device, dtype=torch.float32)
            call("moda_sample_pdf", None, ptr(b), ptr(w), ptr(uu), ptr(out), R, 0, n, n_importance, int(det),
                 float(eps), 0, stream())
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, g):
        return None, None, None, None, None, None, None
