"""Edge cases on hand-made inputs, product CLI vs oracle CLI, all output files:
read pairs with more than 16 overlaps (std::sort stops being stable there), 16-bit
traces (tspace > 125), reads without any overlap at both ends of the id range and
in the middle, self-overlaps with the telomere switches on, no QV track, and INI
variants that switch the masks."""
import os

import numpy as np
import pytest

import handmade as hm
import hingetest as ht

pytestmark = pytest.mark.gpu
ALL = ht.FILTER_OUT + ht.MAXIMAL_OUT + ht.LAYOUT_OUT


def _run_all(work, root, ini=ht.INI):
    for stage in ("filter", "maximal", "layout"):
        ht.run_stage("oracle", stage, work, root, "ora", ini=ini)
        ht.run_stage("product", stage, work, root, "gpu", ini=ini)
    ht.assert_same_files(work, "gpu", "ora", ALL)


def _tiling(rng, n_read, rlen, genome, cov_step):
    """Reads tiled along a line with dense genuine overlaps (both directions)."""
    starts = np.sort(rng.integers(0, genome - max(rlen), n_read))
    recs = []
    for i in range(n_read):
        for j in range(i + 1, n_read):
            s, e = max(starts[i], starts[j]), min(starts[i] + rlen[i], starts[j] + rlen[j])
            if e - s >= 1000:
                recs += hm.both_directions(i, j, int(s - starts[i]), int(e - starts[i]), int(s - starts[j]),
                                           int(e - starts[j]), 0, rlen)
    return recs


TMA_CASES = {"test_reads_without_overlaps_and_self_overlaps", "test_very_deep_pileups_take_the_fallbacks",
             "test_long_read_and_overlaps_inside_one_bin", "test_many_tiny_reads_per_batch",
             "test_reads_just_under_and_over_the_batch_capacity",
             "test_batch_with_more_records_than_the_prefix_counts_hold"}


@pytest.fixture(params=["default", "tma"], autouse=True)
def profile_kernel_form(request, monkeypatch):
    """The edge cases that bear on the batch plan run twice: with the default forms of the coverage-profile
    kernel and with the TMA-staged persistent one (HINGE_B200_PROFILE_KERNEL=3, read by the executables)."""
    if request.param == "tma":
        if request.node.originalname not in TMA_CASES:
            pytest.skip("default form only")
        monkeypatch.setenv("HINGE_B200_PROFILE_KERNEL", "3")
    else:
        monkeypatch.delenv("HINGE_B200_PROFILE_KERNEL", raising=False)


def test_many_overlaps_per_pair_and_ties(built, tmp_path):
    rng = np.random.default_rng(1)
    n = 60
    rlen = [int(x) for x in rng.integers(6000, 12000, n)]
    recs = _tiling(rng, n, rlen, 60000, 0)
    # 40 extra overlaps for a few pairs, many with identical lengths (tandem-repeat like)
    for (a, b) in [(3, 7), (10, 12), (20, 21)]:
        for k in range(40):
            ab = 100 * (k % 20) + 50
            ln = 1500 + 100 * (k % 3)
            recs += hm.both_directions(a, b, ab, ab + ln, 200 + 37 * k % 900, 200 + 37 * k % 900 + ln, 0, rlen)
    hm.write_fixture(str(tmp_path), "H", rlen, recs, tspace=100, qv="good")
    _run_all(str(tmp_path), "H")


def test_sixteen_bit_traces(built, tmp_path):
    rng = np.random.default_rng(2)
    n = 50
    rlen = [int(x) for x in rng.integers(7000, 15000, n)]
    recs = _tiling(rng, n, rlen, 70000, 0)
    for (a, b) in [(5, 30), (6, 31), (7, 32)]:  # a few complemented ones
        recs += hm.both_directions(a, b, 500, 3500, 1000, 4000, 1, rlen)
    hm.write_fixture(str(tmp_path), "W", rlen, recs, tspace=200, qv="good")
    _run_all(str(tmp_path), "W")


def test_reads_without_overlaps_and_self_overlaps(built, tmp_path):
    rng = np.random.default_rng(3)
    n = 70
    rlen = [int(x) for x in rng.integers(11000, 16000, n)]
    core = list(range(5, 60))  # reads 0-4 and 60-69 have no record at all, neither has 33
    core.remove(33)
    sub = [rlen[i] for i in core]
    recs = []
    for (i, j, ab, ae, bb, be, c) in _tiling(rng, len(core), sub, 90000, 0):
        recs.append((core[i], core[j], ab, ae, bb, be, c))
    for r in (8, 9, 40):  # heavy self-overlaps: > 4.5x the read length in total (filter.cpp:552-561)
        for k in range(30):
            recs.append((r, r, 100 + 10 * k, 100 + 10 * k + 9000, 300 + 5 * k, 300 + 5 * k + 9000, 0))
    hm.write_fixture(str(tmp_path), "E", rlen, recs, tspace=100, qv=None)
    ini = os.path.join(str(tmp_path), "telomere.ini")
    with open(ini, "w") as f:
        f.write(open(ht.INI).read() + "del_telomere = 1\ndel_telomeres = 1\nnum_events_telomere = 0\n")
    _run_all(str(tmp_path), "E", ini=ini)
    assert os.path.getsize(os.path.join(str(tmp_path), "ora.self.flag")) > 0, "fixture should flag self-overlapping reads"


VARIANTS = [
    {"use_qv": "false"},
    {"coverage": "false"},
    {"min_cov": "40", "ec": "90"},
    {"theta": "100", "theta2": "50", "aln_threshold": "2500", "length_threshold": "4000",
     "hinge_tolerance": "250", "matching_hinge_slack": "400", "min_connected_component_size": "2"},
    {"cut_off": "0"},
    {"cut_off": "-1", "min_cov": "2"},
    {"cut_off": "130"},
    {"ec": "-30", "min_cov": "-5"},  # MIN_COV < 0: covered runs are not closed by read ends any more
    {"no_hinge_region": "900", "hinge_min_support": "4", "hinge_unbridged": "3", "hinge_min_pileup": "4",
     "hinge_tolerance_length": "150", "repeat_annotation_gap_threshold": "800",
     "min_repeat_annotation_threshold": "6", "max_repeat_annotation_threshold": "9", "use_two_matches": "0"},
]


def write_ini(path, overrides):
    filt = {"length_threshold": "1000;", "aln_threshold": "1000;", "min_cov": "5;", "cut_off": "300;",
            "theta": "300;", "use_qv": "true;"}
    layout = {"hinge_slack": "1000", "min_connected_component_size": "8"}
    layout_keys = {"hinge_slack", "hinge_tolerance", "matching_hinge_slack", "min_connected_component_size",
                   "use_two_matches", "kill_hinge_overlap", "kill_hinge_internal", "del_telomere", "del_telomeres"}
    for k, v in overrides.items():
        (layout if k in layout_keys else filt)[k] = v
    with open(path, "w") as f:
        f.write("[filter]\n" + "".join("%s = %s\n" % kv for kv in filt.items()))
        f.write("\n[layout]\n" + "".join("%s = %s\n" % kv for kv in layout.items()))


@pytest.mark.parametrize("overrides", VARIANTS)
def test_ini_variants(built, tmp_path, overrides):
    root, _ = ht.materialize("synth_small", str(tmp_path))
    ini = os.path.join(str(tmp_path), "variant.ini")
    write_ini(ini, overrides)
    _run_all(str(tmp_path), root, ini=ini)


def _run_filter(work, root, ini=ht.INI):
    ht.run_stage("oracle", "filter", work, root, "ora", ini=ini)
    ht.run_stage("product", "filter", work, root, "gpu", ini=ini)
    ht.assert_same_files(work, "gpu", "ora", ht.FILTER_OUT)


def test_very_deep_pileups_take_the_fallbacks(built, tmp_path):
    """Pile-ups deeper than the 16-bit histogram halves can count (> 32000 records: per-read
    fallback kernels of both phases) and a median coverage >= 4095 (radix select)."""
    rlen = [6000, 6100, 6200, 6300]
    recs = []
    for a in range(4):
        for b in range(4):
            if a == b:
                continue
            for k in range(11000):
                ab, ae = (k * 7) % 200, rlen[a] - (k * 11) % 300
                bb, be = (k * 5) % 150, rlen[b] - (k * 13) % 250
                recs.append((a, b, ab, ae, bb, be, k & 1))
    hm.write_fixture(str(tmp_path), "D", rlen, recs, tspace=100, qv="good")
    _run_filter(str(tmp_path), "D")
    cov = open(os.path.join(str(tmp_path), "ora.coverage.txt")).readline().split()
    assert max(int(x.split(",")[1]) for x in cov[2:]) > 32000


def test_long_read_and_overlaps_inside_one_bin(built, tmp_path):
    """A read with more coverage bins than a batch holds (per-read fallback by length) and records
    that start and end inside one 40-bp bin, including one that alone sets a profile's length."""
    rng = np.random.default_rng(5)
    n = 40
    rlen = [int(x) for x in rng.integers(7000, 12000, n)] + [200000]
    recs = _tiling(rng, n, rlen[:n], 60000, 0)
    long_id = n
    for i in range(0, n, 2):  # the long read overlaps every other read somewhere along its length
        ln = min(rlen[i], 6000)
        at = 4000 * i + 123
        recs += hm.both_directions(long_id, i, at, at + ln, 0, ln, i % 4 == 0, rlen)
    for r in (3, 4, 5):  # 17-bp records inside one bin; for read 5 beyond every other record's end
        pos = 40 * 100 + 3 if r != 5 else 40 * (rlen[5] // 40 - 1) + 2
        recs.append((r, 20, pos, pos + 17, 1000, 1017, 0))
    hm.write_fixture(str(tmp_path), "L", rlen, recs, tspace=100, qv="good")
    _run_filter(str(tmp_path), "L")


def test_many_tiny_reads_per_batch(built, tmp_path):
    """Hundreds of very short reads: batches of the flat kernels close on the read-count limit,
    not on the bin limit; no read reaches the 5000 bp the coverage estimate wants."""
    rng = np.random.default_rng(8)
    n = 1500
    rlen = [int(x) for x in rng.integers(60, 140, n)]
    recs = []
    for i in range(n - 1):
        for d in (1, 2, 3):
            j = i + d
            if j < n:
                ln = min(rlen[i], rlen[j]) - int(rng.integers(5, 20))
                recs += hm.both_directions(i, j, rlen[i] - ln, rlen[i], 0, ln, 0, rlen)
    hm.write_fixture(str(tmp_path), "T", rlen, recs, tspace=100, qv=None)
    ini = os.path.join(str(tmp_path), "tiny.ini")
    write_ini(ini, {"cut_off": "0", "length_threshold": "50", "aln_threshold": "30", "theta": "10"})
    _run_filter(str(tmp_path), "T", ini=ini)


def test_reads_just_under_and_over_the_batch_capacity(built, tmp_path):
    """(rlen + cut_off) / 40 + 3 coverage bins per read: 163459 bp is the longest read that still fits
    the 4096-bin batch of the flat kernels (alone in its batch), 163460 bp the shortest that does not."""
    rng = np.random.default_rng(11)
    n = 30
    rlen = [int(x) for x in rng.integers(8000, 14000, n)] + [163459, 163460, 163419, 163420]
    recs = _tiling(rng, n, rlen[:n], 70000, 0)
    for big in range(n, n + 4):
        for i in range(n):
            ln = min(rlen[i], 7000)
            at = int(rng.integers(0, rlen[big] - ln))
            recs += hm.both_directions(big, i, at, at + ln, 0, ln, i % 3 == 0, rlen)
        # records that reach the very last base: the last event bins of both profiles
        recs += hm.both_directions(big, 0, rlen[big] - 5000, rlen[big], 100, 5100, 0, rlen)
    hm.write_fixture(str(tmp_path), "B", rlen, recs, tspace=100, qv="good")
    _run_filter(str(tmp_path), "B")


def test_batch_with_more_records_than_the_prefix_counts_hold(built, tmp_path):
    """Five reads of 15000 records each: every pile-up fits the packed 16-bit counters, the batch
    (75000 records) does not fit the 16-bit prefix counts of K1's second form -> both phases leave the
    whole batch to the per-read fallbacks."""
    rlen = [6000, 6100, 6200, 6300, 6400, 9000]
    recs = []
    for a in range(5):
        for k in range(15000):
            b = (a + 1 + k % 4) % 5
            if b == a:
                b = 5
            ab, ae = (k * 7) % 900, rlen[a] - (k * 11) % 900
            bb, be = (k * 5) % 150, rlen[b] - (k * 13) % 250
            recs.append((a, b, ab, ae, bb, be, k & 1))
    recs += hm.both_directions(5, 0, 100, 5000, 200, 5100, 0, rlen)
    hm.write_fixture(str(tmp_path), "O", rlen, recs, tspace=100, qv="good")
    _run_filter(str(tmp_path), "O")
