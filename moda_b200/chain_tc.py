"""Fused-chain execution of ``nerf_coarse`` and ``nerf_skin`` (csrc/chain.cu): one persistent tcgen05 kernel per
pass keeps the activations of all layers on chip; only what the other pass needs goes to HBM (fp16 activations
for the weight gradients, ReLU sign bits for the adjoint, fp16 pre-activation gradients for the weight gradients).

Same mathematics, same precision policy and the same autograd interface as the layer-by-layer modules
``trunk_tc`` / ``skin_tc`` (which stay as the reference implementation the chain kernels are tested against):
  nerf_coarse  nnutils/nerf.py:147-198 on [PE(xyz) | dir | env]    fp16 operands, fp32 accumulation
  nerf_skin    same class, 5x64, raw_feat (moda.py:325-329)        forward in split (hi, lo) fp16 precision,
                                                                   adjoint on plain fp16 operands
"""
import ctypes

import torch

from . import config
from ._lib import call, ptr, stream, f32
from .ops import _win_array

HALF = torch.float16
TILE = 128


class _Packer:
    """Collects the fp32 -> fp16 block copies that build one packed-weight matrix and runs them as ONE launch
    (moda_pack16_multi): block = src[:, col0:col0+cols] -> out[:out_rows, out_col0:out_col0+width], zero padded,
    optionally transposed, optionally with the low half of the split-precision pair at column lo_col0."""

    def __init__(self, out):
        self.out, self.jobs = out, []

    def add(self, src, cols, col0, out_col0, out_rows, width, transpose, lo_col0=None):
        self.jobs.append((src, cols, col0, out_col0, out_rows, width, transpose, lo_col0))

    def run(self):
        n, base = len(self.jobs), ptr(self.out)
        P, I = ctypes.c_void_p * n, ctypes.c_int * n
        j = self.jobs
        call("moda_pack16_multi", n, P(*[ptr(x[0]) for x in j]), I(*[x[0].stride(0) for x in j]),
             I(*[x[0].shape[0] for x in j]), I(*[x[1] for x in j]), I(*[x[2] for x in j]),
             P(*[base + 2 * x[3] for x in j]), P(*[(base + 2 * x[7]) if x[7] is not None else None for x in j]),
             self.out.stride(0), I(*[x[4] for x in j]), I(*[x[5] for x in j]), I(*[int(x[6]) for x in j]), stream())
        return self.out


def _wgrad(dY, N, X, K, M, dW, col0, n_valid, k_valid, oscale, dbias=None):
    call("moda_tc_wgrad", ptr(dY), dY.stride(0), N, ptr(X), X.stride(0), K, M, ptr(dW) + 4 * col0, dW.stride(0),
         n_valid, k_valid, ptr(oscale), ptr(dbias), stream())


def _wgrad_multi(jobs, N, K, M, oscale):
    """jobs: list of (dY, X, dW, col0, n_valid, k_valid, dbias | None) sharing the shape (N, K): ONE launch
    (moda_tc_wgrad_multi) whose grid is partitioned among them."""
    n = len(jobs)
    P_, I_ = ctypes.c_void_p * n, ctypes.c_int * n
    call("moda_tc_wgrad_multi", n, P_(*[ptr(j[0]) for j in jobs]), I_(*[j[0].stride(0) for j in jobs]),
         P_(*[ptr(j[1]) for j in jobs]), I_(*[j[1].stride(0) for j in jobs]),
         P_(*[ptr(j[2]) + 4 * j[3] for j in jobs]), I_(*[j[2].stride(0) for j in jobs]), I_(*[j[4] for j in jobs]),
         I_(*[j[5] for j in jobs]), P_(*[ptr(j[6]) if j[6] is not None else None for j in jobs]), N, K, M, ptr(oscale), stream())


def _al16(t):
    """The chain kernels read head vectors with 128-bit loads: a parameter that is a view at an odd offset of some
    larger buffer is copied (a few hundred floats) instead of being refused."""
    return t if t.data_ptr() % 16 == 0 else t.clone()


def _one(v):
    return (ctypes.c_int * 1)(v)


def _small_linear(code, Wl, col0, bias, n):
    """(R, n) fp32 = code Wl[:, col0:col0+nc]^T + bias: the per-ray-constant input columns as a per-ray bias."""
    R, nc = code.shape
    rb = torch.empty(R, n, device=code.device, dtype=torch.float32)
    call("moda_linear_fwd", R, n, 1, (ctypes.c_void_p * 1)(ptr(code)), _one(nc), _one(nc), _one(0), _one(1), None, 0,
         ptr(Wl) + 4 * col0, Wl.shape[1], ptr(bias), 0, ptr(rb), n, stream())
    return rb


_SIDE = {}


class _Alternate:
    """Issues independent kernels alternately on the current stream and on a per-device side stream, so that the tail
    of one persistent weight-gradient kernel (CTAs draining their atomics) overlaps the ramp-up of the next instead
    of serialising behind it.  Everything issued before `with _Alternate(dev) as alt:` is visible to both streams,
    and the current stream waits for the side stream on exit (so buffers may be released by the caller afterwards).
    MODA_B200_SIDE_STREAM=0 keeps everything on the current stream."""

    def __init__(self, dev):
        self.dev, self.i = dev, 0
        self.on = config.side_stream

    def __enter__(self):
        if self.on:
            self.main = torch.cuda.current_stream(self.dev)
            key = (self.dev.index, self.main.cuda_stream)
            if key not in _SIDE:
                _SIDE[key] = torch.cuda.Stream(device=self.dev)
            self.side = _SIDE[key]
            ev = torch.cuda.Event()
            ev.record(self.main)
            self.side.wait_event(ev)
        return self

    def run(self, fn, *a, **k):
        self.i += 1
        if self.on and (self.i & 1) == 0:
            with torch.cuda.stream(self.side):
                fn(*a, **k)
        else:
            fn(*a, **k)

    def __exit__(self, *exc):
        if self.on:
            ev = torch.cuda.Event()
            ev.record(self.side)
            self.main.wait_event(ev)
        return False


class _SideQueue:
    """Weight-gradient launches of a backward pass whose results nothing else in that pass reads (they add into the
    flat gradient buffer of ``parallel.FlatParams``): issued on two side streams and joined with the main stream only
    at the END of the backward pass (autograd engine callback), not at the end of the Function that issued them.
    The HBM-bound weight-gradient kernels then share the GPU with whatever the rest of the pass runs next
    (skin_warp_bwd is FP32-issue bound, the chain kernels latency bound) instead of serialising in front of it.
    The operands are kept alive here until the join, so the caching allocator cannot hand their memory to a later
    main-stream allocation while a side stream still reads it.  MODA_B200_DEFER_WGRAD=0: join inside the Function."""

    def __init__(self, dev, main):
        self.dev, self.main = dev, main
        self.streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
        self.i, self.keep = 0, []

    def fork(self, *keep):
        ev = torch.cuda.Event()
        ev.record(self.main)
        for s in self.streams:
            s.wait_event(ev)
        self.keep.extend(keep)
        # one join per Function and pass (idempotent): runs after the last node of this backward pass was issued
        torch.autograd.Variable._execution_engine.queue_callback(self.join)

    def run(self, fn, *a, **k):
        with torch.cuda.stream(self.streams[self.i & 1]):
            fn(*a, **k)
        self.i += 1

    def join(self):
        cur = torch.cuda.current_stream(self.dev)
        for s in self.streams:
            ev = torch.cuda.Event()
            ev.record(s)
            self.main.wait_event(ev)
            if cur.cuda_stream != self.main.cuda_stream:
                cur.wait_event(ev)
        self.keep.clear()


_QUEUES = {}


def _side_queue(dev, gret):
    """The deferred-join queue of the current stream, or None when the weight gradients must be complete when the
    Function returns (some gradient is handed back to autograd as a tensor, or the switch is off)."""
    if not (config.side_stream and config.defer_wgrad) or any(r is not None for r in gret):
        return None
    main = torch.cuda.current_stream(dev)
    key = (dev.index, main.cuda_stream)
    if key not in _QUEUES:
        _QUEUES[key] = _SideQueue(dev, main)
    return _QUEUES[key]


def _grad_targets(refs, params):
    """One accumulation buffer per parameter plus what backward() hands to autograd for it.  The weight-gradient
    kernels all accumulate (atomic +=).  A parameter re-homed by ``parallel.FlatParams`` carries its gradient as a
    view of the flat gradient buffer ("gradients are written into a flat fp32 buffer whose slices
    are exposed as .grad views"): the kernels add straight into that view and autograd receives None, which saves a
    zero fill and an AccumulateGrad add per parameter and step.  Any other parameter gets a fresh zero tensor that is
    returned to autograd as usual."""
    tg, ret = [], []
    for r, q in zip(refs, params):
        gr = r.grad if getattr(r, "_moda_grad_inplace", False) else None
        if (gr is not None and gr.dtype == torch.float32 and gr.is_contiguous() and gr.device == q.device
                and gr.shape == q.shape):
            tg.append(gr)
            ret.append(None)
        else:
            z = torch.zeros_like(q)
            tg.append(z)
            ret.append(z)
    return tg, ret


def _loss_scale(g):
    scale2 = torch.empty(2, device=g.device, dtype=torch.float32)
    work = torch.empty(1, device=g.device, dtype=torch.int32)
    call("moda_loss_scale", ptr(g), g.numel(), 1024.0, work.data_ptr(), ptr(scale2), stream())
    return scale2[0:1], scale2[1:2]


# ------------------------------------------------------------------------------- final layer folded into the dir layer
def _fold(Wf, bf, Wd, bd, W):
    """xyz_encoding_final (no activation) followed by dir_encoding on [final | rest] (nerf.py:182-190) is one linear map
    of the last hidden activation: W' = Wd[:, :W] Wf and b' = bd + Wd[:, :W] bf.  O(weights) work per call, one launch
    (moda_fold_final, csrc/gemm.cu)."""
    n, dev = Wd.shape[0], Wd.device
    Wp = torch.empty(n, W, device=dev, dtype=torch.float32)
    bp = torch.empty(n, device=dev, dtype=torch.float32)
    call("moda_fold_final", ptr(Wd), Wd.stride(0), ptr(Wf), Wf.stride(0), ptr(bf), ptr(bd), n, W, ptr(Wp), ptr(bp), stream())
    return Wp, bp


def _unfold_grads(gWp, dbp, Wf, bf, Wd, W, g_Wf, g_bf, g_Wd):
    """Maps the gradient of the folded layer back to the parameters it was made of: with W' = Wd1 Wf and b' = bd + Wd1 bf,
    dWf += Wd1^T dW', dbf += Wd1^T db', dWd1 += dW' Wf^T + db' bf^T (dbd = db' is accumulated by the caller).  One launch."""
    call("moda_unfold_final", ptr(gWp), ptr(dbp), ptr(Wd), Wd.stride(0), ptr(Wf), Wf.stride(0), ptr(bf), Wd.shape[0], W,
         ptr(g_Wf), g_Wf.stride(0), ptr(g_bf), ptr(g_Wd), g_Wd.stride(0), stream())


# ------------------------------------------------------------------------------------------------ nerf_coarse
def pack_trunk_fwd(params, fold=None):
    """(256, 38*64) fp16, chunk order documented at moda_chain_trunk_fwd (csrc/chain.cu)."""
    W = [params[2 * i] for i in range(8)]
    Wf, Wd = params[16], params[18]
    nchunk = 34 if fold is not None else 38   # fold: W' (128, 256) replaces [Wfinal, Wdir[:, :256]]
    out = torch.zeros(256, nchunk * 64, device=W[0].device, dtype=HALF)
    pk = _Packer(out)
    _pack = lambda w, cols, col0, _o, out_col0, out_rows, width, tr: pk.add(w, cols, col0, out_col0, out_rows, width, tr)
    col = 0
    for i in range(8):
        if i == 0:
            _pack(W[0], 63, 0, out, col, 256, 64, False); col += 64
        elif i == 4:
            _pack(W[4], 63, 0, out, col, 256, 64, False); col += 64
            _pack(W[4], 256, 63, out, col, 256, 256, False); col += 256
        else:
            _pack(W[i], 256, 0, out, col, 256, 256, False); col += 256
    if fold is not None:
        _pack(fold, 256, 0, out, col, 128, 256, False); col += 256
    else:
        _pack(Wf, 256, 0, out, col, 256, 256, False); col += 256
        _pack(Wd, 256, 0, out, col, 128, 256, False); col += 256
    assert col == nchunk * 64
    return pk.run()


def pack_trunk_bwd(params, fold=None):
    """(256, 42*64) fp16 transposed weights, chunk order documented at moda_chain_trunk_bwd (fold: (256, 38*64), W'^T in
    place of [Wdir^T, Wfinal^T])."""
    W = [params[2 * i] for i in range(8)]
    Wf, Wd = params[16], params[18]
    nchunk = 38 if fold is not None else 42
    out = torch.zeros(256, nchunk * 64, device=W[0].device, dtype=HALF)
    pk = _Packer(out)
    _pack = lambda w, cols, col0, _o, out_col0, out_rows, width, tr: pk.add(w, cols, col0, out_col0, out_rows, width, tr)
    col = 0
    if fold is not None:
        _pack(fold, 256, 0, out, col, 256, 128, True); col += 128
    else:
        _pack(Wd, 256, 0, out, col, 256, 128, True); col += 128
        _pack(Wf, 256, 0, out, col, 256, 256, True); col += 256
    for i in (7, 6, 5):
        _pack(W[i], 256, 0, out, col, 256, 256, True); col += 256
    _pack(W[4], 63, 0, out, col, 64, 256, True); col += 256
    _pack(W[4], 256, 63, out, col, 256, 256, True); col += 256
    for i in (3, 2, 1):
        _pack(W[i], 256, 0, out, col, 256, 256, True); col += 256
    _pack(W[0], 63, 0, out, col, 64, 256, True); col += 256
    assert col == nchunk * 64
    return pk.run()


def trunk_sigma(xyz, win, params):
    """Density only, no gradient (the grid query of mesh extraction: nnutils/train_utils.py:1377-1404 ->
    nerf.py:176-180 with sigma_only=True): xyz (P,3) -> sigma (P,1) through ``moda_chain_trunk_sigma``."""
    xyz = f32(xyz).reshape(-1, 3)
    P, dev = xyz.shape[0], xyz.device
    params = [f32(p.detach()) for p in params]
    Ws, bs = params[20], params[21]
    wa, _ = _win_array(win)
    wpack = pack_trunk_fwd(params)
    biases = (ctypes.c_void_p * 8)(*[ptr(params[2 * i + 1]) for i in range(8)])
    sigma = torch.empty(P, 1, device=dev, dtype=torch.float32)
    call("moda_chain_trunk_sigma", ptr(xyz), P, len(win), wa, ptr(wpack), biases, ptr(_al16(Ws)), ptr(bs), ptr(sigma),
         config.chain_mode(), stream())
    return sigma


class TrunkChainFn(torch.autograd.Function):
    """apply(xyz (P,3), dir_embedded (R,cd), env_code (R,ce) | None, S, win, save, *params) -> raw (P,4) [rgb | sigma].
    ``save``: the caller's ``torch.is_grad_enabled()`` (always False inside Function.forward, and needs_input_grad
    only reflects requires_grad): without it nothing is kept for a backward pass and the kernel stores nothing."""

    @staticmethod
    def forward(ctx, xyz, dir_emb, env, S, win, save, *params):
        xyz_shape = xyz.shape
        xyz = f32(xyz).reshape(-1, 3)
        P, dev = xyz.shape[0], xyz.device
        ctx.param_refs = params
        params = [f32(p) for p in params]
        need_bw = bool(save) and any(ctx.needs_input_grad)
        Wf, bf, Wd, bd, Ws, bs, Wr, br = params[16:24]
        code = f32(dir_emb) if env is None else torch.cat([f32(dir_emb), f32(env)], -1)
        R, cc = code.shape
        assert R * S == P and Wd.shape[1] == 256 + cc
        T = ((P + TILE - 1) // TILE + 1) & ~1   # even: the CTA-pair kernels run tiles two at a time
        wa, _ = _win_array(win)
        fold = config.fold_final
        mode = config.chain_mode()
        if fold:
            Wp, bp = _fold(Wf, bf, Wd, bd, 256)
            rb = _small_linear(code, Wd, 256, bp, 128)
            wpack = pack_trunk_fwd(params, Wp)
        else:
            rb = _small_linear(code, Wd, 256, bd, 128)
            wpack = pack_trunk_fwd(params)
        biases = (ctypes.c_void_p * 9)(*([ptr(params[2 * i + 1]) for i in range(8)] + [ptr(bf)]))
        raw = torch.empty(P, 4, device=dev, dtype=torch.float32)
        if need_bw:
            A0 = torch.empty(P, 64, device=dev, dtype=HALF)
            H = torch.empty(8, P, 256, device=dev, dtype=HALF)
            fin = None if fold else torch.empty(P, 256, device=dev, dtype=HALF)
            dfe = torch.empty(P, 128, device=dev, dtype=HALF)
            bits = torch.empty(8, T, 4, TILE, device=dev, dtype=torch.int64)
        else:
            A0 = H = fin = dfe = bits = None
        call("moda_chain_trunk_fwd", ptr(xyz), P, S, len(win), wa, ptr(wpack), biases, ptr(rb), ptr(_al16(Ws)), ptr(bs),
             ptr(_al16(Wr)),
             ptr(br), ptr(A0), ptr(H), ptr(fin), ptr(dfe), bits.data_ptr() if bits is not None else None, ptr(raw),
             mode, stream())
        if need_bw:
            ctx.save_for_backward(xyz, code, raw, *params)
            ctx.act = (A0, H, fin, dfe, bits)
            ctx.meta = (S, win, dir_emb.shape[-1], env is not None, xyz_shape, mode)
            ctx.Wp = Wp if fold else None   # the folded weights serve the adjoint too
        return raw

    @staticmethod
    def backward(ctx, graw):
        xyz, code, raw = ctx.saved_tensors[:3]
        params = list(ctx.saved_tensors[3:])
        A0, H, fin, dfe, bits = ctx.act
        S, win, cd, has_env, xyz_shape, mode = ctx.meta
        fold = bool(mode & 4)
        P, dev = xyz.shape[0], xyz.device
        R, cc = code.shape
        Wf, bf, Wd, bd, Ws, bs, Wr, br = params[16:24]
        g, gret = _grad_targets(ctx.param_refs, params)
        graw = f32(graw).reshape(P, 4)
        sc, isc = _loss_scale(graw)
        # heads: d_dfe (scaled, masked by dfe > 0), gsig, and the head parameter gradients
        d_dfe = torch.empty(P, 128, device=dev, dtype=HALF)
        gsig = torch.empty(P, device=dev, dtype=torch.float32)
        call("moda_head_bwd", ptr(H[7]), ptr(dfe), ptr(raw), ptr(graw), ptr(Wr), ptr(sc), ptr(d_dfe), ptr(gsig),
             ptr(g[22]), ptr(g[23]), ptr(g[20]), ptr(g[21]), P, stream())
        # direction layer, hoisted per-ray part (fp32, M = rays)
        grb = torch.empty(R, 128, device=dev, dtype=torch.float32)
        call("moda_segsum16", ptr(d_dfe), 128, ptr(grb), R, S, 128, ptr(isc), 0, stream())
        gcode = torch.empty(R, cc, device=dev, dtype=torch.float32)
        call("moda_linear_dgrad", R, 128, cc, ptr(grb), 128, ptr(Wd), Wd.shape[1], 256, None, 0, 0, ptr(gcode), cc,
             stream())
        call("moda_linear_wgrad", R, 128, 1, (ctypes.c_void_p * 1)(ptr(code)), _one(cc), _one(cc), _one(0), _one(1), None,
             0, ptr(grb), 128, ptr(g[18]), Wd.shape[1], 256, ptr(g[19]), stream())
        # the whole data-gradient chain in one kernel
        wpackT = pack_trunk_bwd(params, ctx.Wp if fold else None)
        d_fin = None if fold else torch.empty(P, 256, device=dev, dtype=HALF)
        dY = torch.empty(8, P, 256, device=dev, dtype=HALF)
        d_pe = torch.empty(P, 64, device=dev, dtype=HALF)
        call("moda_chain_trunk_bwd", ptr(d_dfe), ptr(gsig), ptr(_al16(Ws.reshape(-1))), ptr(sc), ptr(wpackT), bits.data_ptr(), P,
             ptr(d_fin), ptr(dY), ptr(d_pe), mode, stream())
        # weight gradients (bias gradients ride along as column sums of the dY operand)
        gxyz = torch.empty(P, 3, device=dev, dtype=torch.float32)
        wa, _ = _win_array(win)
        # the eight 256 x 256 weight gradients (final layer, layers 8..6, the hidden part of layer 5, layers 4..2) in
        # ONE launch, the two PE-input ones (layers 5 and 1) in another
        big = [] if fold else [(d_fin, H[7], g[16], 0, 256, 256, g[17])]
        for i in range(7, 0, -1):
            big.append((dY[i], H[i - 1], g[2 * i], 63 if i == 4 else 0, 256, 256, g[2 * i + 1]))
        pe_jobs = [(dY[4], A0, g[8], 0, 256, 63, None), (dY[0], A0, g[0], 0, 256, 63, g[1])]
        pe_bwd = lambda: call("moda_pe16_bwd", ptr(xyz), ptr(d_pe), None, 64, ptr(gxyz), P, len(win), wa, ptr(isc), 0, stream())
        if fold:
            # folded layer: dW' = d_dfe^T H8 and db' into scratch, then mapped back to dWdir / dWfinal / dbfinal
            gWp = torch.zeros(128, 256, device=dev, dtype=torch.float32)
            dbp = grb.sum(0)   # db' = column sums of d_dfe = sum of the per-ray sums

            def dir_wgrad():
                _wgrad(d_dfe, 128, H[7], 256, P, gWp, 0, 128, 256, isc)
                _unfold_grads(gWp, dbp, Wf, bf, Wd, 256, g[16], g[17], g[18])
        else:
            gWp = dbp = None
            dir_wgrad = lambda: _wgrad(d_dfe, 128, fin, 256, P, g[18], 0, 128, 256, isc)
        q = _side_queue(dev, gret)
        if q is not None:   # weight gradients off the critical path: joined at the end of the backward pass
            q.fork(d_dfe, fin, d_fin, H, dY, A0, isc, g, gWp, dbp, params)
            q.run(_wgrad_multi, big, 256, 256, P, isc)
            q.run(dir_wgrad)
            q.run(_wgrad_multi, pe_jobs, 256, 64, P, isc)
            pe_bwd()
        else:
            with _Alternate(dev) as alt:   # the launches below are independent of each other
                alt.run(dir_wgrad)
                alt.run(_wgrad_multi, big, 256, 256, P, isc)
                alt.run(_wgrad_multi, pe_jobs, 256, 64, P, isc)
                alt.run(pe_bwd)
        ctx.act = None
        gdir = gcode[:, :cd].contiguous()
        genv = gcode[:, cd:].contiguous() if has_env else None
        return (gxyz.reshape(xyz_shape), gdir, genv, None, None, None) + tuple(gret)


# -------------------------------------------------------------------------------------------------- nerf_skin
WD = 64


def pack_skin_fwd(params, nc, fold=None):
    """(64, 18*64) fp16: per layer [Whi | Wlo], order documented at moda_chain_skin_fwd (fold: (64, 16*64), W' (32, 64) in
    place of [Wfinal, Wdir])."""
    W = [params[2 * i] for i in range(5)]
    Wf, Wd, Wr = params[10], params[12], params[16]
    tail = [(fold, 64, 0)] if fold is not None else [(Wf, 64, 0), (Wd, 64, 0)]
    blocks = [(W[0], 63, 0), (W[1], 64, 0), (W[2], 64, 0), (W[3], 64, 0), (W[4], 63, 0), (W[4], 64, 63 + nc)] + tail + [(Wr, 32, 0)]
    out = torch.zeros(64, 2 * len(blocks) * 64, device=W[0].device, dtype=HALF)
    pk = _Packer(out)
    for i, (w, cols, col0) in enumerate(blocks):
        pk.add(w, cols, col0, 128 * i, 64, 64, False, lo_col0=128 * i + 64)
    return pk.run()


def pack_skin_bwd(params, nc, fold=None):
    """(64, 9*64) fp16 transposed (hi) weights, order documented at moda_chain_skin_bwd (fold: (64, 8*64), W'^T in place
    of [Wdir^T, Wfinal^T])."""
    W = [params[2 * i] for i in range(5)]
    Wf, Wd, Wr = params[10], params[12], params[16]
    head = [(fold, 64, 0)] if fold is not None else [(Wd, 64, 0), (Wf, 64, 0)]
    blocks = [(Wr, 32, 0)] + head + [(W[4], 63, 0), (W[4], 64, 63 + nc), (W[3], 64, 0), (W[2], 64, 0), (W[1], 64, 0),
                                     (W[0], 63, 0)]
    out = torch.zeros(64, len(blocks) * 64, device=W[0].device, dtype=HALF)
    pk = _Packer(out)
    for i, (w, cols, col0) in enumerate(blocks):
        pk.add(w, cols, col0, 64 * i, 64, 64, True)
    return pk.run()


class SkinChainFn(torch.autograd.Function):
    """apply(pts (..,3), code (Rc,nc) with Rc in {rays, 1}, S, win, save, *params) -> (P, 32) fp32 delta logits,
    columns >= out_channels are zero (a row pitch the skinning kernels accept directly).  ``save`` as in TrunkChainFn."""

    @staticmethod
    def forward(ctx, pts, code, S, win, save, *params):
        pshape = pts.shape
        pts = f32(pts).reshape(-1, 3)
        P, dev = pts.shape[0], pts.device
        ctx.param_refs = params
        params = [f32(p) for p in params]
        # code None: a net without code columns (nerf_vis, moda.py:344-348) = one shared, empty code row
        code = torch.zeros(1, 0, device=dev, dtype=torch.float32) if code is None else f32(code).reshape(-1, code.shape[-1])
        Rc, nc = code.shape
        rep = S if Rc * S == P else P
        assert Rc * rep == P, "pose code rows do not match the points"
        need_bw = bool(save) and any(ctx.needs_input_grad)
        W = [params[2 * i] for i in range(5)]
        b = [params[2 * i + 1] for i in range(5)]
        bf, bd, br = params[11], params[13], params[17]
        oc = br.shape[0]
        T = ((P + TILE - 1) // TILE + 1) & ~1   # even: the CTA-pair kernels run tiles two at a time
        wa, _ = _win_array(win)
        if nc:
            rb1 = _small_linear(code, W[0], 63, b[0], 64)
            rb5 = _small_linear(code, W[4], 63, b[4], 64)
        else:
            rb1, rb5 = _al16(b[0].reshape(1, 64)), _al16(b[4].reshape(1, 64))
        fold = config.fold_final
        Wp, bp = _fold(params[10], bf, params[12], bd, 64) if fold else (None, bd)
        pad = torch.zeros(2, 64, device=dev, dtype=torch.float32)
        pad[0, :bd.shape[0]] = bp
        pad[1, :oc] = br
        wpack = pack_skin_fwd(params, nc, Wp)
        biases = (ctypes.c_void_p * 8)(ptr(rb1), ptr(b[1]), ptr(b[2]), ptr(b[3]), ptr(rb5), ptr(bf), ptr(pad[0]), ptr(pad[1]))
        out = torch.empty(P, 32, device=dev, dtype=torch.float32)
        if need_bw:
            A0 = torch.empty(P, WD, device=dev, dtype=HALF)
            H = torch.empty(5, P, WD, device=dev, dtype=HALF)
            fin = None if fold else torch.empty(P, WD, device=dev, dtype=HALF)
            dfe = torch.empty(P, WD, device=dev, dtype=HALF)
            bits = torch.empty(6, T, 2, TILE, device=dev, dtype=torch.int64)
        else:
            A0 = H = fin = dfe = bits = None
        call("moda_chain_skin_fwd", ptr(pts), P, rep, len(win), wa, ptr(wpack), biases, ptr(A0), ptr(H), ptr(fin),
             ptr(dfe), bits.data_ptr() if bits is not None else None, ptr(out), int(fold), stream())
        if need_bw:
            ctx.save_for_backward(pts, code, *params)
            ctx.act = (A0, H, fin, dfe, bits)
            ctx.meta = (S, win, rep, pshape, oc, fold)
            ctx.Wp = Wp
        return out

    @staticmethod
    def backward(ctx, gout):
        pts, code = ctx.saved_tensors[:2]
        params = list(ctx.saved_tensors[2:])
        A0, H, fin, dfe, bits = ctx.act
        S, win, rep, pshape, oc, fold = ctx.meta
        P, dev = pts.shape[0], pts.device
        Rc, nc = code.shape
        W = [params[2 * i] for i in range(5)]
        g, gret = _grad_targets(ctx.param_refs, params)
        gout = f32(gout).reshape(P, 32)
        sc, isc = _loss_scale(gout)
        Wf, bf, Wd, bd = params[10], params[11], params[12], params[13]
        wpackT = pack_skin_bwd(params, nc, ctx.Wp if fold else None)
        h16 = lambda: torch.empty(P, WD, device=dev, dtype=HALF)
        G, d_dfe, d_pe = h16(), h16(), h16()
        d_fin = None if fold else h16()
        dY = torch.empty(5, P, WD, device=dev, dtype=HALF)
        call("moda_chain_skin_bwd", ptr(gout), ptr(sc), ptr(wpackT), bits.data_ptr(), P, ptr(G), ptr(d_dfe), ptr(d_fin),
             ptr(dY), ptr(d_pe), int(fold), stream())
        gcode = torch.zeros_like(code)

        def code_part(dYl, Wl, gW, gb, rbg=None):
            """hoisted pose-code columns of layers 1 / 5: everything happens at ray (or single-row) level in fp32.
            rbg: per-ray (or, for a single shared code row, whole-batch) sums of dYl; for the shared row they come
            for free as the bias-gradient output of the layer's weight-gradient kernel."""
            if nc == 0:   # no code columns: only the bias gradient (the column sums) is left
                gb.add_(rbg[0])
                return
            if rbg is None:
                rbg = torch.empty(Rc, 64, device=dev, dtype=torch.float32)
                call("moda_segsum16", ptr(dYl), WD, ptr(rbg), Rc, rep, 64, ptr(isc), 0, stream())
            call("moda_linear_dgrad", Rc, 64, nc, ptr(rbg), 64, ptr(Wl), Wl.shape[1], 63, None, 0, 1, ptr(gcode), nc,
                 stream())
            call("moda_linear_wgrad", Rc, 64, 1, (ctypes.c_void_p * 1)(ptr(code)), _one(nc), _one(nc), _one(0), _one(1),
                 None, 0, ptr(rbg), 64, ptr(gW), Wl.shape[1], 63, ptr(gb), stream())

        shared_row = Rc == 1
        rb4 = torch.zeros(1, 64, device=dev, dtype=torch.float32) if shared_row else None
        rb0 = torch.zeros(1, 64, device=dev, dtype=torch.float32) if shared_row else None
        gpts = torch.empty(P, 3, device=dev, dtype=torch.float32)
        wa, _ = _win_array(win)
        # all nine 64 x 64 weight gradients of the evaluation in ONE launch (bias gradients ride along; for a single
        # shared pose-code row the column sums of dY[4] / dY[0] double as the code-part reductions)
        if fold:
            # folded layer: dW' = d_dfe^T H5 and db' into scratch, mapped back to dWdir / dWfinal / dbfinal after the launch
            gWp = torch.zeros(32, 64, device=dev, dtype=torch.float32)
            dbp = torch.zeros(32, device=dev, dtype=torch.float32)
            mid = [(d_dfe, H[4], gWp, 0, 32, 64, dbp)]
        else:
            gWp = dbp = None
            mid = [(d_dfe, fin, g[12], 0, 32, 64, g[13]), (d_fin, H[4], g[10], 0, 64, 64, g[11])]
        jobs = [(dY[4], A0, g[8], 0, 64, 63, rb4), (dY[0], A0, g[0], 0, 64, 63, rb0),
                (G, dfe, g[16], 0, oc, 32, g[17])] + mid + [(dY[4], H[3], g[8], 63 + nc, 64, 64, None)]
        jobs += [(dY[i], H[i - 1], g[2 * i], 0, 64, 64, g[2 * i + 1]) for i in (3, 2, 1)]

        def rest_jobs(js):
            _wgrad_multi(js, WD, WD, P, isc)
            if fold:
                _unfold_grads(gWp, dbp, Wf, bf, Wd, 64, g[10], g[11], g[12])
                g[13].add_(dbp)
        want_pts = ctx.needs_input_grad[0]   # e.g. nerf_vis is evaluated on detached points (loss_utils.py:125-149)
        pe_bwd = lambda: want_pts and call("moda_pe16_bwd", ptr(pts), ptr(d_pe), None, WD, ptr(gpts), P, len(win), wa,
                                           ptr(isc), 0, stream())
        q = _side_queue(dev, gret)
        if q is not None:
            # the two jobs whose column sums feed code_part stay on the current stream; the other seven are joined at
            # the end of the backward pass
            if shared_row:
                _wgrad_multi(jobs[:2], WD, WD, P, isc)
            q.fork(G, dfe, d_dfe, fin, d_fin, H, dY, A0, isc, g, gWp, dbp, params)
            q.run(rest_jobs, jobs[2:] if shared_row else jobs)
            pe_bwd()
            code_part(dY[4], W[4], g[8], g[9], rb4)
            code_part(dY[0], W[0], g[0], g[1], rb0)
        else:
            rest_jobs(jobs)
            with _Alternate(dev) as alt:
                alt.run(pe_bwd)
                code_part(dY[4], W[4], g[8], g[9], rb4)
                code_part(dY[0], W[0], g[0], g[1], rb0)
        ctx.act = None
        gret[14] = gret[15] = None   # nerf_skin's sigma head is computed and discarded in the reference (nerf.py:178)
        return (gpts.reshape(pshape) if want_pts else None, gcode if nc else None, None, None, None) + tuple(gret)


# ------------------------------------------------------------------------------------------ 5 x 128 raw-feature MLP
FW = 128


def feat_supported(model, embed_xyz, k):
    """nerf_feat's architecture (nnutils/moda.py:447-449) on a plain PE(xyz) input."""
    return (embed_xyz is not None and embed_xyz.N_freqs == 10 and k == 3 and model.raw_feat and model.in_channels_dir == 0
            and model.in_channels_xyz == 63 and model.D == 5 and model.W == FW and list(model.skips) == [4]
            and model.out_channels <= 32 and model.dir_encoding[0].weight.shape[0] == 64)


def _pack_feat_fwd(params, Wp):
    W = [params[2 * i] for i in range(5)]
    Wr = params[16]
    out = torch.zeros(FW, 13 * 64, device=W[0].device, dtype=HALF)
    pk = _Packer(out)
    col = 0
    for w, cols, col0, rows, width in ((W[0], 63, 0, FW, 64), (W[1], FW, 0, FW, FW), (W[2], FW, 0, FW, FW), (W[3], FW, 0, FW, FW),
                                       (W[4], 63, 0, FW, 64), (W[4], FW, 63, FW, FW), (Wp, FW, 0, 64, FW), (Wr, 64, 0, 64, 64)):
        pk.add(w, cols, col0, col, rows, width, False)
        col += width
    assert col == 13 * 64
    return pk.run()


def _pack_feat_bwd(params, Wp):
    W = [params[2 * i] for i in range(5)]
    Wr = params[16]
    out = torch.zeros(FW, 14 * 64, device=W[0].device, dtype=HALF)
    pk = _Packer(out)
    col = 0
    # transposed blocks: rows = input channel of the layer, `width` columns = its output channels (zero padded)
    for w, cols, col0, rows, width in ((Wr, 64, 0, 64, 64), (Wp, FW, 0, FW, 64), (W[4], 63, 0, 64, FW), (W[4], FW, 63, FW, FW),
                                       (W[3], FW, 0, FW, FW), (W[2], FW, 0, FW, FW), (W[1], FW, 0, FW, FW), (W[0], 63, 0, 64, FW)):
        pk.add(w, cols, col0, col, rows, width, True)
        col += width
    assert col == 14 * 64
    return pk.run()


class FeatChainFn(torch.autograd.Function):
    """apply(pts (..,3), win, save, *params) -> (P, 32) fp32 raw features (columns >= out_channels zero): nerf_feat as ONE
    chain kernel per pass (csrc/chain.cu: moda_chain_feat_fwd / _bwd) instead of one tensor-core kernel per layer
    (generic_tc, which stays as the reference implementation this is tested against).  Same conventions as the other
    chain Functions; xyz_encoding_final is always folded into dir_encoding."""

    @staticmethod
    def forward(ctx, pts, win, save, *params):
        pshape = pts.shape
        pts = f32(pts).reshape(-1, 3)
        P, dev = pts.shape[0], pts.device
        ctx.param_refs = params
        params = [f32(p) for p in params]
        need_bw = bool(save) and any(ctx.needs_input_grad)
        b = [params[2 * i + 1] for i in range(5)]
        Wf, bf, Wd, bd, Wr, br = params[10], params[11], params[12], params[13], params[16], params[17]
        oc = br.shape[0]
        Wp, bp = _fold(Wf, bf, Wd, bd, FW)
        pad = torch.zeros(2, 64, device=dev, dtype=torch.float32)
        pad[0] = bp
        pad[1, :oc] = br
        T = ((P + TILE - 1) // TILE + 1) & ~1
        wa, _ = _win_array(win)
        mode = config.chain_mode() & 3
        wpack = _pack_feat_fwd(params, Wp)
        biases = (ctypes.c_void_p * 7)(*([ptr(_al16(x)) for x in b] + [ptr(pad[0]), ptr(pad[1])]))
        out = torch.empty(P, 32, device=dev, dtype=torch.float32)
        if need_bw:
            A0 = torch.empty(P, 64, device=dev, dtype=HALF)
            H = torch.empty(5, P, FW, device=dev, dtype=HALF)
            dfe = torch.empty(P, 64, device=dev, dtype=HALF)
            bits = torch.empty(6, T, 4, TILE, device=dev, dtype=torch.int64)
        else:
            A0 = H = dfe = bits = None
        call("moda_chain_feat_fwd", ptr(pts), P, len(win), wa, ptr(wpack), biases, ptr(A0), ptr(H), ptr(dfe),
             bits.data_ptr() if bits is not None else None, ptr(out), mode, stream())
        if need_bw:
            ctx.save_for_backward(pts, *params)
            ctx.act = (A0, H, dfe, bits)
            ctx.meta = (win, pshape, oc, mode)
            ctx.Wp = Wp
        return out

    @staticmethod
    def backward(ctx, gout):
        pts = ctx.saved_tensors[0]
        params = list(ctx.saved_tensors[1:])
        A0, H, dfe, bits = ctx.act
        win, pshape, oc, mode = ctx.meta
        P, dev = pts.shape[0], pts.device
        Wf, bf, Wd, bd = params[10], params[11], params[12], params[13]
        g, gret = _grad_targets(ctx.param_refs, params)
        gout = f32(gout).reshape(P, 32)
        sc, isc = _loss_scale(gout)
        wpackT = _pack_feat_bwd(params, ctx.Wp)
        h16 = lambda n: torch.empty(P, n, device=dev, dtype=HALF)
        G, d_dfe, d_pe = h16(64), h16(64), h16(64)
        dY = torch.empty(5, P, FW, device=dev, dtype=HALF)
        call("moda_chain_feat_bwd", ptr(gout), ptr(sc), ptr(wpackT), bits.data_ptr(), P, ptr(G), ptr(d_dfe), ptr(dY),
             ptr(d_pe), mode, stream())
        gWp = torch.zeros(64, FW, device=dev, dtype=torch.float32)
        dbp = torch.zeros(64, device=dev, dtype=torch.float32)
        hidden = [(dY[4], H[3], g[8], 63, FW, FW, g[9])] + [(dY[i], H[i - 1], g[2 * i], 0, FW, FW, g[2 * i + 1]) for i in (3, 2, 1)]
        pe_jobs = [(dY[4], A0, g[8], 0, FW, 63, None), (dY[0], A0, g[0], 0, FW, 63, g[1])]
        want_pts = ctx.needs_input_grad[0]
        gpts = torch.empty(P, 3, device=dev, dtype=torch.float32) if want_pts else None
        wa, _ = _win_array(win)

        def wgrads():
            _wgrad_multi(hidden, FW, FW, P, isc)
            _wgrad_multi(pe_jobs, FW, 64, P, isc)
            _wgrad_multi([(d_dfe, H[4], gWp, 0, 64, FW, dbp)], 64, FW, P, isc)
            _wgrad_multi([(G, dfe, g[16], 0, oc, 64, g[17])], 64, 64, P, isc)
            _unfold_grads(gWp, dbp, Wf, bf, Wd, FW, g[10], g[11], g[12])
            g[13].add_(dbp)

        q = _side_queue(dev, gret)
        if q is not None:
            q.fork(G, dfe, d_dfe, H, dY, A0, isc, g, gWp, dbp, params)
            q.run(wgrads)
        else:
            wgrads()
        if want_pts:
            call("moda_pe16_bwd", ptr(pts), ptr(d_pe), None, 64, ptr(gpts), P, len(win), wa, ptr(isc), 0, stream())
        ctx.act = None
        gret[14] = gret[15] = None   # the sigma head is computed and discarded by the reference (nerf.py:178)
        return (gpts.reshape(pshape) if want_pts else None, None, None) + tuple(gret)
