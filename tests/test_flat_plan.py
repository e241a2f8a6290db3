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
        if any(nbz[r] > K_BINS for r in range(f0, f1)):
            assert f1 - f0 == 1 and used == 0, "a read that fits no batch is a batch of its own"
        # greedy: the next read did not fit (or the read-count limit closed the batch)
        elif f1 < hi and nbz[f1] <= K_BINS:
            assert used + nbz[f1] > K_BINS or f1 - f0 == K_READS


K_SLAB = 4096


def _plan2(rlen, read_off, lo, hi, cut_off):
    from hinge_b200._lib import lib
    lib.hg_debug_flat_plan2.restype = C.c_int
    lib.hg_debug_flat_plan2.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    rlen = np.ascontiguousarray(rlen, np.int32)
    read_off = np.ascontiguousarray(read_off, np.int64)
    n = len(rlen)
    cap = n + 2
    batch = np.zeros((cap, 2), np.int32)
    rbase, rbatch, nrec_b = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(cap, np.int32)
    nb = lib.hg_debug_flat_plan2(rlen.ctypes.data, read_off.ctypes.data, n, lo, hi, cut_off, batch.ctypes.data, cap,
                                 rbase.ctypes.data, rbatch.ctypes.data, nrec_b.ctypes.data)
    assert nb >= 0
    return batch[:nb + 1], rbase, rbatch, nrec_b[:nb]


@pytest.mark.parametrize("case", ["pacbio", "deep", "empty_reads"])
def test_plan_bounded_by_records(built, case):
    """With the CSR the batches are also bounded by kSlabRecords records (what the TMA-staged form of the
    profile kernel holds in shared memory); deeper pile-ups are batches of their own on the generic path."""
    rng = np.random.default_rng(11)
    n, cut_off = 6000, 300
    rlen = np.maximum(1000, rng.normal(3500, 1500, n)).astype(np.int32)
    if case == "pacbio":
        nrec = rng.poisson(90, n)
    elif case == "deep":
        nrec = rng.poisson(400, n)
        nrec[[5, 700, 701, 5999]] = [5000, 4097, 4096, 20000]
    else:
        nrec = rng.poisson(60, n) * (rng.random(n) < 0.5)
    read_off = np.concatenate([[0], np.cumsum(nrec)]).astype(np.int64)
    lo, hi = 10, n - 5
    batch, rbase, rbatch, nrec_b = _plan2(rlen, read_off, lo, hi, cut_off)
    nbz = _bins(rlen.astype(np.int64), cut_off)
    first = batch[:, 0]
    assert first[0] == lo and first[-1] == hi and np.all(np.diff(first) > 0)
    assert np.all(rbatch[:lo] == -1) and np.all(rbatch[hi:] == -1)
    for b in range(len(batch) - 1):
        f0, f1, used = int(first[b]), int(first[b + 1]), int(batch[b, 1])
        assert np.all(rbatch[f0:f1] == b)
        fit = (nbz[f0:f1] <= K_BINS) & (nrec[f0:f1] <= K_SLAB)
        if not fit.all():
            assert f1 - f0 == 1 and used == 0 and rbase[f0] == -1 and nrec_b[b] == 0
            continue
        assert used == nbz[f0:f1].sum() <= K_BINS and nrec[f0:f1].sum() <= K_SLAB and f1 - f0 <= K_READS
        assert nrec_b[b] == nrec[f0:f1].sum(), "the batch descriptor carries its record count"
        if f1 < hi and nbz[f1] <= K_BINS and nrec[f1] <= K_SLAB:
            assert (used + nbz[f1] > K_BINS or nrec[f0:f1].sum() + nrec[f1] > K_SLAB or f1 - f0 == K_READS)
