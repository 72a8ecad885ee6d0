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


def set_precision(p):
    global precision
    if p not in ("fp16", "fp32"):
        raise ValueError("precision must be 'fp16' or 'fp32'")
    precision = p


# side_stream: the independent weight-gradient launches of a backward pass alternate between the current stream and a
# side stream (tail/ramp-up overlap of the persistent kernels); "0" keeps everything on the current stream.
side_stream = os.environ.get("MODA_B200_SIDE_STREAM", "1") != "0"

# trunk_pair: run the 256-wide chain kernels as CTA pairs (tcgen05 cta_group::2; csrc/chain.cu, PAIR = 1).  Results
# are bit-identical; +1.7 % on the training step and +5.5 % on the density grid, so the default is on (the library
# falls back to the single-CTA kernels by itself if a cluster launch fails).  MODA_B200_TRUNK_PAIR=0 switches it off.
trunk_pair = os.environ.get("MODA_B200_TRUNK_PAIR", "1") != "0"


def set_trunk_pair(on):
    """Switches the CTA-pair launch mode of the nerf_coarse chains (takes effect at the next call)."""
    global trunk_pair
    from . import _lib
    trunk_pair = bool(on)
    _lib.call("moda_chain_set_pair", int(trunk_pair))
