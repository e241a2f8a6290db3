"""warp_sort_exact (hg_order.h), the warp-parallel restatement of libstdc++'s std::sort used by the
hinge call's order-exact path, against the real std::sort: element for element on tie-heavy arrays,
sorted / reversed inputs and median-of-3 killer sequences (heap-sort fallback)."""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _killer(n):
    k = n // 2
    v = np.zeros(n, np.int32)
    for i in range(k):
        v[i] = i + 1 if i % 2 == 0 else k + i + (0 if k % 2 == 0 else 1)
        v[k + i] = 2 * (i + 1)
    if n % 2:
        v[n - 1] = n
    return v


def test_warp_sort_matches_std_sort(built):
    from hinge_b200 import Context
    from hinge_b200._lib import lib

    lib.hg_debug_warp_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    lib.hg_debug_std_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    rng = np.random.default_rng(7)
    arrays = []
    for n in [1, 2, 3, 15, 16, 17, 18, 31, 32, 33, 47, 64, 65, 100, 127, 200, 257, 316, 500, 1000, 2048, 5000]:
        for distinct in [1, 2, 3, 5, 17, 100, 1000, 1 << 30]:
            for rep in range(4):
                k = rng.integers(0, distinct, n).astype(np.int32)
                if rep == 2:
                    k.sort()
                if rep == 3:
                    k = np.sort(k)[::-1].copy()
                arrays.append(k)
    for n in [64, 100, 1000, 4096, 30000]:
        arrays += [_killer(n), -_killer(n), _killer(n) // 3]
    off = np.zeros(len(arrays) + 1, np.int32)
    np.cumsum([len(a) for a in arrays], out=off[1:])
    data = np.zeros((off[-1], 2), np.int32)
    data[:, 0] = np.concatenate(arrays)
    for w in range(len(arrays)):
        data[off[w]:off[w + 1], 1] = np.arange(off[w + 1] - off[w])
    ctx = Context(0)
    for cta, desc in ((0, 0), (0, 1), (1, 0), (1, 1)):  # one warp per array / one CTA per array
        if cta:
            os.environ["HINGE_B200_DEBUG_SORT_CTA"] = "1"
        else:
            os.environ.pop("HINGE_B200_DEBUG_SORT_CTA", None)
        want, got = data.copy(), data.copy()
        assert lib.hg_debug_std_sort(want.ctypes.data, off.ctypes.data, len(arrays), desc) == 0
        assert lib.hg_debug_warp_sort(ctx._h, got.ctypes.data, off.ctypes.data, len(arrays), desc) == 0
        bad = np.nonzero((want != got).any(axis=1))[0]
        if len(bad):
            w = int(np.searchsorted(off, bad[0], side="right") - 1)
            raise AssertionError("array %d (n=%d, descending=%d, cta=%d) differs from std::sort at element %d"
                                 % (w, off[w + 1] - off[w], desc, cta, bad[0] - off[w]))
    os.environ.pop("HINGE_B200_DEBUG_SORT_CTA", None)
    ctx.close()
