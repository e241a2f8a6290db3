"""hinge_b200 — B200-native (sm_100a) implementation of the HINGE hot path.

The product is the C-ABI library `_build/libhinge_b200.so` (CUDA kernels + host
front-end, see include/hinge_b200.h).  This package is the thin Python mirror
of that ABI — it computes nothing itself and has no Python or CPU fallback:

    Context            the hg_ctx_* / hg_set_* / hg_filter* / hg_maximal / hg_layout calls
    main_filter(argv)  what `Reads_filter`, `get_maximal_reads`, `hinging` do,
    main_maximal(argv) same flags and files as the reference executables
    main_layout(argv)

`python -m hinge_b200.build` compiles the library (nvcc, sm_100a).  Until it
exists every attribute of this package except `build` raises ImportError.
"""
import os as _os

LIB_PATH = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "_build", "libhinge_b200.so")

if _os.path.exists(LIB_PATH):
    from ._lib import lib, HingeError  # noqa: F401
    from .api import Context, FilterParams, LayoutParams, main_filter, main_maximal, main_layout  # noqa: F401
else:
    def __getattr__(name):
        if name in ("build", "__path__", "__file__", "__spec__", "__loader__"):
            raise AttributeError(name)
        raise ImportError(
            "hinge_b200: %s is missing — build it with `python -m hinge_b200.build` (nvcc, sm_100a). "
            "There is no Python/CPU fallback for the CUDA path." % LIB_PATH)
