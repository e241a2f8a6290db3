"""Parity of the product's `hinge filter` (CUDA, through the C ABI) with the
reference's golden outputs and with the oracle on fresh synthetic inputs."""
import pytest

import hingetest as ht

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ht.FIXTURES)
def test_filter_matches_golden(built, tmp_path, name):
    root, _ = ht.materialize(name, str(tmp_path))
    ht.run_stage("product", "filter", str(tmp_path), root, "gpu")
    ht.assert_matches_golden(name, str(tmp_path), "gpu", ht.FILTER_OUT)


@pytest.mark.parametrize("args", [
    ["--genome", 1000000, "--cov", 30, "--seed", 7],
    ["--genome", 800000, "--cov", 60, "--seed", 99, "--read-mean", 6000, "--read-sd", 3000, "--families", 6],
    ["--genome", 500000, "--cov", 20, "--seed", 3, "--qv-bad", 0.02, "--families", 4, "--copies-max", 6],
])
def test_filter_matches_oracle_on_fresh_synthetic(built, tmp_path, args):
    ht.synth(str(tmp_path), args + ["--bps", "0"], "S")
    ht.run_stage("oracle", "filter", str(tmp_path), "S", "ora")
    ht.run_stage("product", "filter", str(tmp_path), "S", "gpu")
    ht.assert_same_files(str(tmp_path), "gpu", "ora", ht.FILTER_OUT)
