"""Import the unmodified reference (read-only, /root/reference) for tests and golden-vector generation.

The reference's sampler module imports a few third-party packages that are absent here and that the sampling
path never calls (plotting, FID, pairwise): they are stubbed at import time.  Nothing in the
product package or in the GPU-side tests depends on this file; it is skipped when the reference is absent.
"""
import importlib
import itertools
import os
import sys
import types

REF = os.environ.get("NLC_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "src"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


def load():
    """Returns a namespace with the reference modules the hot path needs."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    try:
        import distutils.util  # noqa: F401  (must precede the more_itertools stub: setuptools imports it)
    except Exception:
        pass
    try:
        import more_itertools  # noqa: F401
    except Exception:
        _stub("more_itertools", pairwise=itertools.pairwise)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            mp = _stub("matplotlib")
            mp.pyplot = _stub("matplotlib.pyplot")
            mp.animation = _stub("matplotlib.animation")
    try:
        import pytorch_fid  # noqa: F401
    except Exception:
        pf = _stub("pytorch_fid")
        pf.fid_score = _stub("pytorch_fid.fid_score", calculate_fid_given_paths=None, compute_statistics_of_path=None,
                             calculate_frechet_distance=None)
        pf.inception = _stub("pytorch_fid.inception", InceptionV3=None)
    ns = types.SimpleNamespace()
    for name in ("src.schedulers", "src.unet_ddim", "src.utils", "src.experiments", "src.constraint_functions",
                 "functions.svd_operators"):
        try:
            setattr(ns, name.split(".")[-1], importlib.import_module(name))
        except Exception as e:  # keep going: callers check for what they need
            setattr(ns, name.split(".")[-1], None)
            setattr(ns, name.split(".")[-1] + "_error", e)
    return ns
