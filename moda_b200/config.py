"""Run-time switches of the CUDA path.

precision:
  "fp16"  (default) the 8x256 trunk runs on tcgen05 tensor cores with fp16 operands and fp32 accumulation
          (csrc/tc_gemm.cu); nerf_skin and all skinning / compositing math stay fp32.
  "fp32"  every linear layer runs on the fp32 SIMT kernels (csrc/gemm.cu): the exact mode used to pin parity.
"""
import os

precision = os.environ.get("MODA_B200_PRECISION", "fp16")
# fused: in "fp16" precision, run each MLP pass as ONE chain kernel that keeps activations on chip
# (csrc/chain.cu); False selects the layer-by-layer tensor-core kernels (csrc/tc_gemm.cu), kept as the reference
# implementation the chain kernels are tested against.
fused = os.environ.get("MODA_B200_FUSED", "1") != "0"


class exact:
    """``with config.exact():`` runs the enclosed MLP evaluations on the fp32 SIMT kernels whatever ``precision`` says
    (for small, precision-critical evaluations such as the 8000-point feature lattice of feat_match, whose features go
    through exp((f.v - 1) / 0.03))."""

    def __enter__(self):
        global precision
        self.old = precision
        precision = "fp32"
        return self

    def __exit__(self, *exc):
        global precision
        precision = self.old
        return False


def set_precision(p):
    global precision
    if p not in ("fp16", "fp32"):
        raise ValueError("precision must be 'fp16' or 'fp32'")
    precision = p


# side_stream: the independent weight-gradient launches of a backward pass alternate between the current stream and a
# side stream (tail/ramp-up overlap of the persistent kernels); "0" keeps everything on the current stream.
side_stream = os.environ.get("MODA_B200_SIDE_STREAM", "1") != "0"
# defer_wgrad: weight gradients that accumulate in place into the flat gradient buffer (parallel.FlatParams) are issued
# on side streams and joined at the END of the backward pass (chain_tc._SideQueue) instead of inside their Function:
# the HBM-bound kernels overlap the rest of the pass.  Needs side_stream; "0" joins inside the Function.
defer_wgrad = os.environ.get("MODA_B200_DEFER_WGRAD", "1") != "0"

# Launch mode of the 256-wide chains (csrc/chain.cu), passed to the library with every call (no library-side state):
# trunk_pair:  CTA pairs (tcgen05 cta_group::2, each CTA stages half of every weight chunk).  MODA_B200_TRUNK_PAIR=0: off.
# trunk_slots: two tiles in flight per CTA (needs trunk_pair): the tensor core works on one tile's layer while the other
#              tile's epilogue drains its accumulator.  MODA_B200_TRUNK_SLOTS=1 selects one tile per CTA.
# Results are bit-identical in every mode; the library falls back to single-CTA kernels if a cluster cannot be placed.
trunk_pair = os.environ.get("MODA_B200_TRUNK_PAIR", "1") != "0"
trunk_slots = 1 if os.environ.get("MODA_B200_TRUNK_SLOTS", "2") == "1" else 2


def set_trunk_pair(on):
    """Switches the CTA-pair launch mode of the nerf_coarse chains (takes effect at the next call)."""
    global trunk_pair
    trunk_pair = bool(on)


def set_trunk_slots(n):
    """1 or 2 tiles in flight per CTA for the nerf_coarse chains (2 needs the CTA-pair mode)."""
    global trunk_slots
    if n not in (1, 2):
        raise ValueError("trunk_slots must be 1 or 2")
    trunk_slots = n


# fold_final: the chains run xyz_encoding_final and dir_encoding (two linear maps in a row, nerf.py:182-190) as ONE layer
# with the product weights W' = Wdir[:, :W] Wfinal: one step less per pass, one saved activation and one gradient less,
# one weight-gradient job less; dW' is mapped back to dWdir / dWfinal with two small products.  MODA_B200_FOLD_FINAL=0: off.
fold_final = os.environ.get("MODA_B200_FOLD_FINAL", "1") != "0"


# feat_chain: nerf_feat (5 x 128 raw-feature MLP) as one chain kernel per pass (csrc/chain.cu: moda_chain_feat_*);
# MODA_B200_FEAT_CHAIN=0 keeps the layer-by-layer tensor-core kernels (generic_tc).
feat_chain = os.environ.get("MODA_B200_FEAT_CHAIN", "1") != "0"


def chain_mode():
    """The `mode` argument of moda_chain_trunk_*: bit 0 = CTA pairs, bit 1 = two tile slots, bit 2 = final layer folded."""
    return (1 if trunk_pair else 0) | (2 if (trunk_pair and trunk_slots == 2) else 0) | (4 if fold_final else 0)
