"""filter -> maximal -> layout through the product's `hinge` front-end (CUDA via
the C ABI): every output file must equal the reference's golden bytes, and the
oracle's on fresh synthetic inputs (including each stage fed with the ORACLE's
inter-stage files, so a stage cannot hide behind an upstream mismatch)."""
import os
import shutil

import pytest

import hingetest as ht

pytestmark = pytest.mark.gpu
ALL = ht.FILTER_OUT + ht.MAXIMAL_OUT + ht.LAYOUT_OUT


@pytest.mark.parametrize("name", ht.FIXTURES)
def test_pipeline_matches_golden(built, tmp_path, name):
    root, _ = ht.materialize(name, str(tmp_path))
    for stage in ("filter", "maximal", "layout"):
        ht.run_stage("product", stage, str(tmp_path), root, "gpu")
    ht.assert_matches_golden(name, str(tmp_path), "gpu", ALL)


@pytest.mark.parametrize("args", [
    ["--genome", 1000000, "--cov", 30, "--seed", 7],
    ["--genome", 500000, "--cov", 50, "--seed", 202, "--jitter", 0, "--read-mean", 7000, "--read-sd", 2500,
     "--read-min", 2000, "--rep-min", 1500, "--rep-max", 3500, "--copies-min", 3, "--copies-max", 5,
     "--families", 6],
    ["--genome", 300000, "--cov", 80, "--seed", 31, "--read-mean", 12000, "--read-sd", 5000, "--read-min", 3000,
     "--families", 3, "--rep-min", 8000, "--rep-max", 20000],
])
def test_stages_match_oracle_on_fresh_synthetic(built, tmp_path, args):
    work = str(tmp_path)
    ht.synth(work, args + ["--bps", "0"], "S")
    for stage in ("filter", "maximal", "layout"):
        ht.run_stage("oracle", stage, work, "S", "ora")
    # each product stage on the oracle's inter-stage files
    for ext in ("mas", "repeat.txt", "hinges.txt", "max"):
        shutil.copy(os.path.join(work, "ora." + ext), os.path.join(work, "mix." + ext))
    ht.run_stage("product", "maximal", work, "S", "mix")
    ht.assert_same_files(work, "mix", "ora", ht.MAXIMAL_OUT)
    ht.run_stage("product", "layout", work, "S", "mix")
    ht.assert_same_files(work, "mix", "ora", ht.LAYOUT_OUT)
    # and the whole product pipeline end to end
    for stage in ("filter", "maximal", "layout"):
        ht.run_stage("product", stage, work, "S", "gpu")
    ht.assert_same_files(work, "gpu", "ora", ALL)


@pytest.mark.parametrize("name,nparts", [("synth_small", 3), ("synth_long", 4), ("dal_small", 2)])
def test_mlas_pipeline_matches_oracle(built, tmp_path, name, nparts):
    """`--mlas` (what every demo script of the reference passes, demo/ecoli_demo/run.sh:21-25): the .las
    split at A-read boundaries; all three stages, all output files, against the oracle's restatement of
    the reference's part loop (pinned to the reference binaries in tests/test_oracle_golden.py)."""
    work = str(tmp_path)
    root, _ = ht.materialize(name, work)
    ht.split_las(os.path.join(work, root + ".las"), os.path.join(work, "P"), nparts)
    for stage in ("filter", "maximal", "layout"):
        ht.run_stage_mlas("oracle", stage, work, root, "P", "ora")
        ht.run_stage_mlas("product", stage, work, root, "P", "gpu")
    ht.assert_same_files(work, "gpu", "ora", ALL)
