"""Array-level parity of hg_filter (C ABI, through the Python mirror) with the
oracle on in-memory synthetic batches, including tie-heavy data that drives the
hinge call through its order-exact sort path, sharded runs, and edge cases."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def _run_both(built, synth_kw, options=()):
    import hgsynth
    import oraclelib
    from hinge_b200 import api

    s = hgsynth.Synth(**synth_kw)
    s.generate(want_trace=False, threads=4)
    cols = {k: v.copy() for k, v in s.cols().items()}
    rlen, qv_off, qv = s.rlen.copy(), s.qv_off.copy(), s.qv.copy()
    orc = oraclelib.Oracle(rlen, qv_off, qv, 100, cols)
    want = orc.filter()
    orc.close()
    ctx = api.Context(0)
    for opt, val in options:
        ctx.set_option(opt, val)
    ctx.set_reads(rlen, qv_off, qv, 100)
    ctx.set_overlaps(len(cols["aread"]), cols)
    summ = ctx.filter(api.FilterParams())
    got = ctx.filter_fetch(int(summ.n_annotations))
    ctx.close()
    s.close()
    return got, want, summ


def _assert_equal(got, want, summ):
    assert (summ.r_begin, summ.r_end, summ.cov_est, summ.min_cov) == tuple(want["summary"])
    for k in ("mask", "cmask", "flags", "anno_off", "anno_pos", "anno_type", "hinge_keep"):
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)


def test_filter_arrays_match_oracle(built):
    got, want, summ = _run_both(built, dict(genome_len=600000, coverage=40.0, seed=77, n_families=4))
    _assert_equal(got, want, summ)
    assert want["hinge_keep"].sum() > 0, "fixture should call hinges"


def test_both_forms_of_the_profile_kernel_match_oracle(built):
    """HG_OPT_PROFILE_KERNEL = 1 forces the four-event 40-bp form of K1 (the one that also serves
    cut-offs off the 20-bp grid), 3 the TMA-staged persistent form; the default picks the 20-bp
    start / end histogram here."""
    from hinge_b200 import api

    kw = dict(genome_len=500000, coverage=45.0, seed=78, n_families=4)
    for form in (0, 1, 3):
        got, want, summ = _run_both(built, kw, options=[(api.HG_OPT_PROFILE_KERNEL, form)])
        _assert_equal(got, want, summ)


def test_tma_staged_profile_kernel_on_long_reads(built):
    """HG_OPT_PROFILE_KERNEL = 3: the persistent form of K1 that stages the abpos / aepos columns and the
    per-read tables of a batch in shared memory with TMA bulk copies (batches bounded by record volume,
    deeper pile-ups on the generic path), on the long-read shape with repeat-induced deep pile-ups."""
    from hinge_b200 import api

    got, want, summ = _run_both(built, dict(genome_len=3000000, coverage=40.0, read_mean=24000, read_sd=8000,
                                           read_min=2000, seed=99, frag_prob=1.2, n_families=12),
                                options=[(api.HG_OPT_PROFILE_KERNEL, 3)])
    _assert_equal(got, want, summ)


def test_enqueued_runs_match_oracle(built):
    """hg_filter_enqueue three times back to back (no wait in between), one hg_filter_finish: the results
    are those of one hg_filter."""
    import hgsynth
    import oraclelib
    from hinge_b200 import api

    s = hgsynth.Synth(genome_len=500000, coverage=45.0, seed=78, n_families=4)
    s.generate(want_trace=False, threads=4)
    cols = {k: v.copy() for k, v in s.cols().items()}
    rlen, qv_off, qv = s.rlen.copy(), s.qv_off.copy(), s.qv.copy()
    orc = oraclelib.Oracle(rlen, qv_off, qv, 100, cols)
    want = orc.filter()
    orc.close()
    ctx = api.Context(0)
    ctx.set_reads(rlen, qv_off, qv, 100)
    ctx.set_overlaps(len(cols["aread"]), cols)
    params = api.FilterParams()
    ctx.filter(params)  # sizes the annotation pool
    for _ in range(3):
        ctx.filter_enqueue(params)
    rc, summ = ctx.filter_finish()
    assert rc == 0
    got = ctx.filter_fetch(int(summ.n_annotations))
    ctx.close()
    s.close()
    _assert_equal(got, want, summ)


def test_c5_shaped_long_reads_with_fragmented_alignments(built):
    """BASELINE configs[4] shape at 1/40 of its size: reads N(24000, 8000) (~600 coverage bins and ~200
    records each, ~6 reads per batch of the flat kernels), most pairs reported as two or three local
    alignments."""
    got, want, summ = _run_both(built, dict(genome_len=7500000, coverage=40.0, read_mean=24000, read_sd=8000,
                                           read_min=2000, seed=4321, frag_prob=1.2, n_families=20))
    _assert_equal(got, want, summ)
    assert len(want["anno_pos"]) > 100 and want["hinge_keep"].sum() > 0


def test_annotation_pool_overflow_is_retried(built):
    """A pool of 16 entries for a batch with hundreds of annotations: hg_filter grows it and reruns the
    stage (HG_RETRY_POOL inside), the results are the oracle's."""
    from hinge_b200 import api

    got, want, summ = _run_both(built, dict(genome_len=600000, coverage=40.0, seed=77, n_families=4),
                                options=[(api.HG_OPT_ANNO_POOL, 16)])
    assert len(want["anno_pos"]) > 64
    _assert_equal(got, want, summ)


def test_tie_heavy_data_uses_order_exact_path(built):
    # no end jitter: every repeat-induced alignment starts exactly on the repeat
    # boundary, and reads longer than the repeats see both boundaries, so end lists
    # are full of equal positions with different overhang classes
    got, want, summ = _run_both(built, dict(genome_len=500000, coverage=60.0, seed=5, n_families=8, jitter=0,
                                           read_mean=8000, read_sd=2500, read_min=2000,
                                           rep_min_len=1500, rep_max_len=3500, rep_min_copies=3,
                                           rep_max_copies=6))
    _assert_equal(got, want, summ)
    assert summ.n_exact_order > 0, "expected position ties that need the order-exact sort"


def test_long_reads_take_the_big_histogram_path(built):
    # a few reads far longer than the 99.9th percentile exceed the shared-memory histogram
    got, want, summ = _run_both(built, dict(genome_len=400000, coverage=30.0, seed=9, read_mean=4000,
                                           read_sd=6000, read_min=1000, read_max=120000, n_families=2))
    _assert_equal(got, want, summ)


def test_records_out_of_order_are_rejected(built):
    from hinge_b200 import api, HingeError

    rlen = np.array([5000, 6000, 7000], np.int32)
    cols = {"aread": np.array([1, 0], np.int32), "bread": np.array([0, 1], np.int32),
            "abpos": np.array([0, 0], np.int32), "aepos": np.array([2000, 2000], np.int32),
            "bbpos": np.array([0, 0], np.int32), "bepos": np.array([2000, 2000], np.int32),
            "flags": np.array([0, 0], np.int32)}
    ctx = api.Context(0)
    ctx.set_reads(rlen)
    with pytest.raises(HingeError):
        ctx.set_overlaps(2, cols)
    ctx.close()
