"""ctypes binding of libmoda_b200.so (the C ABI declared in include/moda_b200.h).

There is no CPU fallback: if the shared library is missing, or a tensor is not a contiguous fp32 CUDA
tensor, the call raises.  Pointers are borrowed for the duration of the (stream-ordered) call; every
buffer is a torch-allocated CUDA tensor (ownership).
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MODA_B200_LIB") or os.path.join(_HERE, "libmoda_b200.so")  # override: kernel experiments

_lib = None

c_p = ctypes.c_void_p
c_i = ctypes.c_int
c_ll = ctypes.c_longlong
c_f = ctypes.c_float
c_pp = ctypes.POINTER(ctypes.c_void_p)
c_ip = ctypes.POINTER(ctypes.c_int)
c_fp = ctypes.POINTER(ctypes.c_float)

# name -> argtypes (every entry point returns int; 0 = success).  Keep in sync with include/moda_b200.h;
# tests/test_abi.py checks that each of these symbols is exported and declared in the header.
SIGNATURES = {
    "moda_device_check": [],
    "moda_embed_fwd": [c_p, c_i, c_p, c_i, c_ll, c_i, c_i, c_fp, c_p],
    "moda_embed_bwd": [c_p, c_i, c_p, c_i, c_p, c_i, c_ll, c_i, c_i, c_fp, c_i, c_p],
    "moda_sample_rays_fwd": [c_p, c_p, c_p, c_p, c_p, c_f, c_i, c_p, c_p, c_p, c_i, c_i, c_p],
    "moda_points_from_depths": [c_p, c_p, c_p, c_p, c_i, c_i, c_p],
    "moda_fold_final": [c_p, c_i, c_p, c_i, c_p, c_p, c_i, c_i, c_p, c_p, c_p],
    "moda_unfold_final": [c_p, c_p, c_p, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_i, c_p, c_p, c_i, c_p],
    "moda_sinkhorn_pass": [c_p, c_i, c_i, c_p, c_p, c_p, c_p, c_i, c_f, c_f, c_p, c_p, c_i, c_p, c_p, c_p, c_f, c_p, c_p],
    "moda_sinkhorn_matrix": [c_p, c_p, c_i, c_i, c_i, c_f, c_f, c_p, c_p, c_p],
    "moda_sinkhorn_rows4": [c_p, c_i, c_i, c_p, c_p, c_p],
    "moda_sinkhorn_cols4": [c_p, c_i, c_i, c_p, c_p, c_p],
    "moda_sinkhorn_gcost": [c_p, c_i, c_i, c_p, c_p, c_i, c_p, c_p, c_i, c_f, c_p, c_p, c_p],
    "moda_flow_render_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_f, c_p, c_p, c_p],
    "moda_flow_render_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_f, c_p, c_p, c_p, c_p, c_p, c_p],
    "moda_adamw_flat": [c_p, c_p, c_p, c_p, c_ll, c_p, c_f, c_f, c_f, c_f, c_f, c_p],
    "moda_sample_rays_bwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_p],
    "moda_dq_unary_fwd": [c_i, c_p, c_p, c_ll, c_p],
    "moda_dq_unary_bwd": [c_i, c_p, c_p, c_p, c_ll, c_p],
    "moda_dq_mul_fwd": [c_p, c_p, c_p, c_ll, c_i, c_p],
    "moda_dq_mul_bwd": [c_p, c_p, c_p, c_p, c_p, c_ll, c_i, c_p],
    "moda_bone_transform_fwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "moda_bone_transform_bwd": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "moda_skin_warp_fwd": [c_p] * 8 + [c_i] * 7 + [c_p],
    "moda_skin_warp_bwd": [c_p] * 14 + [c_i] * 8 + [c_p],
    "moda_composite_fwd": [c_p, c_i, c_p, c_i] + [c_p] * 13 + [c_i, c_i, c_p],
    "moda_composite_bwd": [c_p, c_i, c_p, c_i] + [c_p] * 13 + [c_p, c_i, c_p, c_i] + [c_p] * 4 + [c_i, c_i, c_p],
    "moda_linear_fwd": [c_i, c_i, c_i, c_pp, c_ip, c_ip, c_ip, c_ip, c_fp, c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_p],
    "moda_linear_dgrad": [c_i, c_i, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_i, c_i, c_p, c_i, c_p],
    "moda_linear_wgrad": [c_i, c_i, c_i, c_pp, c_ip, c_ip, c_ip, c_ip, c_fp, c_i, c_p, c_i, c_p, c_i, c_i, c_p, c_p],
    "moda_segsum": [c_p, c_i, c_p, c_i, c_i, c_i, c_p],
    "moda_sample_pdf": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_f, c_i, c_p],
    "moda_tc_linear": [c_p, c_i, c_i, c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_p, c_p, c_i, c_i, c_p, c_i, c_p, c_p, c_p,
                       c_p, c_i, c_i, c_p, c_i, c_p, c_p],
    "moda_tc_wgrad": [c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_p, c_p, c_p],
    "moda_tc_wgrad_multi": [c_i, c_pp, c_ip, c_pp, c_ip, c_pp, c_ip, c_ip, c_ip, c_pp, c_i, c_i, c_i, c_p, c_p],
    "moda_tc_linear_split": [c_p, c_p, c_i, c_i, c_p, c_p, c_i, c_i, c_p, c_i, c_i, c_i, c_p, c_p, c_i, c_i, c_p, c_i,
                             c_p, c_p, c_i, c_i, c_p, c_i, c_p, c_p],
    "moda_tc_wgrad_split": [c_p, c_p, c_i, c_i, c_p, c_p, c_i, c_i, c_i, c_p, c_i, c_i, c_i, c_p, c_p, c_p],
    "moda_pe16_fwd": [c_p, c_p, c_p, c_i, c_ll, c_i, c_fp, c_p],
    "moda_pe16_bwd": [c_p, c_p, c_p, c_i, c_p, c_ll, c_i, c_fp, c_p, c_i, c_p],
    "moda_pack16": [c_p, c_i, c_i, c_i, c_i, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_p],
    "moda_pack16_multi": [c_i, c_pp, c_ip, c_ip, c_ip, c_ip, c_pp, c_pp, c_i, c_ip, c_ip, c_ip, c_p],
    "moda_split16": [c_p, c_i, c_i, c_p, c_p, c_p, c_i, c_i, c_ll, c_p],
    "moda_head_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_ll, c_p],
    "moda_head_bwd": [c_p] * 12 + [c_ll, c_p],
    "moda_colsum16": [c_p, c_i, c_p, c_ll, c_i, c_p, c_p],
    "moda_segsum16": [c_p, c_i, c_p, c_i, c_i, c_i, c_p, c_i, c_p],
    "moda_loss_scale": [c_p, c_ll, c_f, c_p, c_p, c_p],
    "moda_chain_trunk_fwd": [c_p, c_ll, c_i, c_i, c_fp, c_p, c_pp] + [c_p] * 11 + [c_i, c_p],
    "moda_chain_trunk_sigma": [c_p, c_ll, c_i, c_fp, c_p, c_pp, c_p, c_p, c_p, c_i, c_p],
    "moda_chain_trunk_bwd": [c_p] * 6 + [c_ll] + [c_p] * 3 + [c_i, c_p],
    "moda_chain_skin_fwd": [c_p, c_ll, c_i, c_i, c_fp, c_p, c_pp] + [c_p] * 6 + [c_i, c_p],
    "moda_chain_feat_fwd": [c_p, c_ll, c_i, c_fp, c_p, c_pp] + [c_p] * 5 + [c_i, c_p],
    "moda_chain_feat_bwd": [c_p] * 4 + [c_ll] + [c_p] * 4 + [c_i, c_p],
    "moda_chain_skin_bwd": [c_p] * 4 + [c_ll] + [c_p] * 5 + [c_i, c_p],
    "moda_chain_set_trace": [c_p],
    "moda_chain_pair_available": [],
    "moda_wsum_fwd": [c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "moda_wsum_bwd": [c_p, c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p],
    "moda_act_bwd": [c_i, c_p, c_i, c_p, c_i, c_p, c_i, c_ll, c_i, c_p],
}


def lib():
    """Loads the shared library once; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libmoda_b200.so is missing at %s: run `python -m moda_b200.build` "
                           "(there is no CPU or PyTorch fallback for this path)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.moda_last_error.restype = ctypes.c_char_p
    L.moda_version.restype = ctypes.c_char_p
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = ctypes.c_int
    L.moda_chain_pair_available.restype = ctypes.c_int   # a query, not a status code
    _lib = L
    return L


LAUNCHES = 0  # number of C-ABI calls that launched kernels (reported by bench.py as gpu_launches)
PROFILE = None  # when a dict: name -> list of (start_event, end_event) recorded on the current stream


def call(name, *args):
    global LAUNCHES
    L = lib()
    if PROFILE is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(L, name)(*args)
        e1.record()
        PROFILE.setdefault(name, []).append((e0, e1))
    else:
        rc = getattr(L, name)(*args)
    if rc != 0:
        raise RuntimeError("%s failed (%d): %s" % (name, rc, L.moda_last_error().decode()))
    LAUNCHES += 1


def profile_summary():
    """name -> (calls, total ms) from the events gathered while PROFILE was a dict (synchronises)."""
    torch.cuda.synchronize()
    out = {}
    for name, evs in (PROFILE or {}).items():
        out[name] = (len(evs), sum(a.elapsed_time(b) for a, b in evs))
    return out


def ptr(t):
    """Device pointer of a contiguous fp32 CUDA tensor (None passes through as NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("moda_b200 ops need CUDA tensors (there is no CPU fallback); got a %s tensor" % t.device)
    if t.dtype not in (torch.float32, torch.float16, torch.uint8, torch.bool):
        raise RuntimeError("moda_b200 ops need float32 tensors; got %s" % t.dtype)
    if not t.is_contiguous():
        raise RuntimeError("moda_b200 ops need contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def f32(t):
    """Contiguous fp32 view/copy (no-op for conforming tensors)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
