"""Host-side batch plan of the flat filter kernels (hg_filter_flat.cu::flat_plan): the batches
tile the planned read range, respect the histogram and read-count limits, and every read's
profile sits where the kernels will look for it.  CPU only (the library loads without a GPU)."""
import ctypes as C

import numpy as np
import pytest

K_BINS, K_READS = 4096, 512  # kFlatBins, kFlatMaxReads (hinge_b200/csrc/hg_filter.h)


def _plan(rlen, lo, hi, cut_off):
    from hinge_b200._lib import lib

    lib.hg_debug_flat_plan.restype = C.c_int
    lib.hg_debug_flat_plan.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_int32, C.c_void_p]
    rlen = np.ascontiguousarray(rlen, np.int32)
    cap = len(rlen) + 2
    batch = np.zeros((cap, 2), np.int32)
    rbase = np.zeros(len(rlen), np.int32)
    nb = lib.hg_debug_flat_plan(rlen.ctypes.data, len(rlen), lo, hi, cut_off, batch.ctypes.data, cap,
                                rbase.ctypes.data)
    assert nb >= 0
    return batch[:nb + 1], rbase


def _bins(rlen, cut_off):
    return (rlen + abs(cut_off)) // 40 + 3


@pytest.mark.parametrize("case", ["pacbio", "tiny", "mixed_with_giants", "one_read", "range"])
def test_plan_invariants(built, case):
    rng = np.random.default_rng(5)
    cut_off, lo = 300, 0
    if case == "pacbio":
        rlen = np.maximum(1000, rng.normal(3500, 1500, 20000)).astype(np.int32)
    elif case == "tiny":
        rlen, cut_off = rng.integers(40, 200, 5000).astype(np.int32), 0
    elif case == "mixed_with_giants":
        rlen = np.maximum(500, rng.normal(24000, 8000, 3000)).astype(np.int32)
        rlen[[0, 17, 1500, 2999]] = [400000, 170000, 163000, 900000]  # around / beyond 4096 bins
    elif case == "one_read":
        rlen = np.array([12345], np.int32)
    else:
        rlen = np.maximum(1000, rng.normal(3500, 1500, 5000)).astype(np.int32)
        lo = 1234
    hi = len(rlen) - (77 if case == "range" else 0)
    batch, rbase = _plan(rlen, lo, hi, cut_off)
    nbz = _bins(rlen.astype(np.int64), cut_off)
    first = batch[:, 0]
    assert first[0] == lo and first[-1] == hi and np.all(np.diff(first) > 0), "batches tile [lo, hi) in order"
    assert np.all(rbase[:lo] == -1) and np.all(rbase[hi:] == -1), "reads outside the range are not planned"
    for b in range(len(batch) - 1):
        f0, f1, used = int(first[b]), int(first[b + 1]), int(batch[b, 1])
        assert f1 - f0 <= K_READS
        pos = 0
        for r in range(f0, f1):
            if nbz[r] > K_BINS:
                assert rbase[r] == -1, "reads longer than a batch go the generic way"
                continue
            assert rbase[r] == pos, "profiles are laid end to end"
            pos += int(nbz[r])
        assert pos == used and used <= K_BINS
        # greedy: the next read did not fit (or the read-count limit closed the batch)
        if f1 < hi and nbz[f1] <= K_BINS:
            assert used + nbz[f1] > K_BINS or f1 - f0 == K_READS
