"""hinge_b200 — B200-native (sm_100a) implementation of the HINGE hot path.

The product is the C-ABI library `_build/libhinge_b200.so` (CUDA kernels + host
front-end, see include/hinge_b200.h).  This package is the thin Python mirror
of that ABI: it loads the library (and fails loudly when it is missing — there
is no Python or CPU fallback) and exposes

    Context            the hg_ctx_* / hg_set_* / hg_filter* / hg_maximal / hg_layout calls
    main_filter(argv)  what `Reads_filter`, `get_maximal_reads`, `hinging` do,
    main_maximal(argv) same flags and files as the reference executables
    main_layout(argv)
"""
from ._lib import LIB_PATH, lib, HingeError  # noqa: F401
from .api import Context, FilterParams, LayoutParams, main_filter, main_maximal, main_layout  # noqa: F401
