"""TEST INFRASTRUCTURE ONLY -- loader for the *real* MoDA reference (read-only at /root/reference).

Only usable in the build container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to generate tests/golden/*.npz and by tests that pin oracle/restated.py
against the reference itself.  Nothing in moda_b200/ may import this module.

Recipe follows SURVEY.md section 8(c): the reference is pure Python/PyTorch, its hot path
(nnutils/rendering.py, nnutils/nerf.py, nnutils/geom_utils.py, nnutils/dual_quat.py) imports with
a handful of stub modules for packages that are only used off the hot path.
"""
import os
import sys
import types

REF = os.environ.get("MODA_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "nnutils"))


_loaded = None


def load():
    """Returns a namespace with the reference's hot-path symbols (imports happen once)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    saved_path = list(sys.path)
    for p in (REF + "/third_party/pytorch3d", REF + "/third_party", REF + "/nnutils", REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    # geom_utils.py:7, nerf.py:10, loss_utils.py:4, ext_utils/flowlib.py:24-26 import these at
    # module top but never touch them on the rendering path.
    for m in ("trimesh", "png", "matplotlib", "matplotlib.pyplot", "matplotlib.colors",
              "matplotlib.cm"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = types.ModuleType(m)
    cwd = os.getcwd()
    try:
        from nnutils import rendering, nerf, geom_utils  # noqa
        import dual_quat
    finally:
        os.chdir(cwd)
        # the reference tree carries generic top-level names (tests/, utils/): do not leave it on the path
        sys.path[:] = saved_path
    ns = types.SimpleNamespace(rendering=rendering, nerf=nerf, geom_utils=geom_utils,
                               dual_quat=dual_quat, render_rays=rendering.render_rays,
                               NeRF=nerf.NeRF, Embedding=nerf.Embedding)
    _loaded = ns
    return ns
