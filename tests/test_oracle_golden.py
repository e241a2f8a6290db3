"""The oracle is only worth something if it IS the reference: its output files
must equal, byte for byte, what the unmodified reference binaries wrote for the
committed fixtures (tests/golden, made by make_golden.py), and — where the
reference binaries are available (oracle/_ref) — what they write for fresh
synthetic inputs."""
import os

import pytest

import hingetest as ht

ALL = ht.FILTER_OUT + ht.MAXIMAL_OUT + ht.LAYOUT_OUT


@pytest.mark.parametrize("name", ht.FIXTURES)
def test_oracle_reproduces_golden(built, tmp_path, name):
    root, _ = ht.materialize(name, str(tmp_path))
    for stage in ("filter", "maximal", "layout"):
        ht.run_stage("oracle", stage, str(tmp_path), root, "ora")
    ht.assert_matches_golden(name, str(tmp_path), "ora", ALL)


@pytest.mark.skipif(not ht.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("args", [
    ["--genome", 500000, "--cov", 35, "--seed", 101, "--families", 4],
    ["--genome", 400000, "--cov", 50, "--seed", 202, "--jitter", 0, "--read-mean", 7000, "--read-sd", 2500,
     "--read-min", 2000, "--rep-min", 1500, "--rep-max", 3500, "--copies-min", 3, "--copies-max", 5,
     "--families", 6],
])
def test_oracle_equals_reference_on_fresh_synthetic(built, tmp_path, args):
    ht.synth(str(tmp_path), args, "S")
    for stage in ("filter", "maximal", "layout"):
        ht.run_stage("reference", stage, str(tmp_path), "S", "S", out="S")
    for stage in ("filter", "maximal", "layout"):
        ht.run_stage("oracle", stage, str(tmp_path), "S", "ora")
    ht.assert_same_files(str(tmp_path), "ora", "S", ALL)


@pytest.mark.skipif(not ht.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_equals_reference_on_edge_cases(built, tmp_path):
    """The hand-made edge-case fixtures and INI variants of tests/test_gpu_edge_cases.py, oracle vs the
    unmodified reference binaries (this is what pins the oracle on those inputs)."""
    import pathlib

    import test_gpu_edge_cases as ec

    def oracle_vs_reference(work, root, ini=ht.INI):
        if not os.path.exists(os.path.join(work, "." + root + ".bps")):  # the reference loads the bases
            import json

            meta = json.load(open(os.path.join(ht.GOLDEN, "synth_small", "fixture.json")))
            ht.synth(work, meta["synth_args"], root)
        for stage in ("filter", "maximal", "layout"):
            ht.run_stage("oracle", stage, work, root, "ora", ini=ini)
            ht.run_stage("reference", stage, work, root, "ref", out="ref", ini=ini)
        ht.assert_same_files(work, "ora", "ref", ALL)

    saved = ec._run_all
    ec._run_all = oracle_vs_reference
    try:
        for k, fn in enumerate((ec.test_many_overlaps_per_pair_and_ties, ec.test_sixteen_bit_traces,
                                ec.test_reads_without_overlaps_and_self_overlaps)):
            d = tmp_path / ("case%d" % k)
            d.mkdir()
            fn(True, pathlib.Path(d))
        for k, ov in enumerate(ec.VARIANTS):
            d = tmp_path / ("ini%d" % k)
            d.mkdir()
            ec.test_ini_variants(True, pathlib.Path(d), ov)
    finally:
        ec._run_all = saved


@pytest.mark.skipif(not ht.have_reference(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("name,nparts", [("synth_small", 3), ("synth_long", 2)])
def test_oracle_equals_reference_with_mlas(built, tmp_path, name, nparts):
    """--mlas: the .las split at A-read boundaries (LAsplit).  The reference loops over the parts and
    carries MIN_COV and the masks from part to part, closes .repeat.txt after part 0 and drops the last
    read of every part from .hinges.txt (filter.cpp:474-1109), so its results differ from a single-file
    run; the oracle restates exactly that (FilterCarry) and must write the same bytes."""
    import json

    work = str(tmp_path)
    meta = json.load(open(os.path.join(ht.GOLDEN, name, "fixture.json")))
    root = meta["root"]
    ht.synth(work, meta["synth_args"], root)  # with the bases: the reference loads them
    ht.split_las(os.path.join(work, root + ".las"), os.path.join(work, "P"), nparts)
    for stage in ("filter", "maximal", "layout"):
        ht.run_stage_mlas("reference", stage, work, root, "P", "ref", out="ref")
        ht.run_stage_mlas("oracle", stage, work, root, "P", "ora")
    ht.assert_same_files(work, "ora", "ref", ALL)
    # and the multi-part run really is a different computation
    ht.run_stage("oracle", "filter", work, root, "one")
    assert ht.sha256(os.path.join(work, "one.hinges.txt")) != ht.sha256(os.path.join(work, "ora.hinges.txt"))
