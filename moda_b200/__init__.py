"""moda_b200 -- B200-native articulated volume renderer behind MoDA's Python API.

Public surface (names and signatures of the reference):
    rendering.render_rays, rendering.inference, rendering.inference_deform, rendering.sample_pdf
    nerf.NeRF, nerf.Embedding
    geom_utils.bone_transform / skinning / gauss_mlp_skinning / mlp_skinning / dqs_blend_skinning / neu_dbs /
               evaluate_mlp / vec_to_sim3
    dual_quat.q_mul / dq_mul / dq_normalize / dq_inverse / dq_quaternion_conjugate / dq_combined_conjugate
Everything executes in hand-written sm_100a CUDA (libmoda_b200.so, C ABI in include/moda_b200.h); there is
no CPU, Triton or PyTorch-eager fallback.
"""
__version__ = "0.1.0"

_LAZY = ("rendering", "nerf", "geom_utils", "dual_quat", "ops", "synth", "models", "parallel", "extract")


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return importlib.import_module("." + name, __name__)
    if name == "render_rays":
        from .rendering import render_rays
        return render_rays
    raise AttributeError(name)
