"""The C-ABI library loads, exports every symbol include/hinge_b200.h declares,
and refuses to compute without a CUDA device (no CPU fallback)."""
import os
import re
import subprocess

import pytest

import hingetest as ht

ROOT = ht.ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "hinge_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(built):
    import ctypes

    lib = ctypes.CDLL(ht.LIB)
    names = _declared_symbols()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_python_mirror_covers_the_header(built):
    from hinge_b200 import _lib

    assert not _lib.MISSING, _lib.MISSING
    assert sorted(_lib.PROTOTYPES) == _declared_symbols()


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_gpu_means_loud_failure_not_fallback(built, tmp_path):
    from hinge_b200 import Context, HingeError

    with pytest.raises(HingeError):
        Context(0)
    # the CLI reads its inputs (same counts as the reference logs), then stops with exit code 1
    root, meta = ht.materialize("dal_small", str(tmp_path))
    r = ht.run_stage("product", "filter", str(tmp_path), root, "gpu", check=False)
    assert r.returncode == 1
    assert "# Reads: 278" in r.stdout and "# Alignments: 11978" in r.stdout
    assert "no CPU fallback" in r.stdout
    assert not os.path.exists(os.path.join(str(tmp_path), "gpu.mas"))


def test_cli_error_conventions(built, tmp_path):
    # filter.cpp:218-226: flag-combination errors exit with status 1
    r = subprocess.run([ht.HINGE, "filter", "--db", "x", "--las", "y", "--paf", "z", "--config", ht.INI],
                       cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "Pass in either a db and a las or a fasta and a paf" in r.stdout
    r = subprocess.run([ht.HINGE, "filter", "--db", "x", "--config", ht.INI], cwd=str(tmp_path),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "Pass in at least one of the following two combinations" in r.stdout
    r = subprocess.run([ht.HINGE, "layout", "--db", "x", "--las", "y", "--config", ht.INI], cwd=str(tmp_path),
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1  # --prefix and --out are required for layout (hinging.cpp:627-628)
    r = subprocess.run([ht.HINGE, "filter", "--db", "missing", "--las", "missing", "--config", ht.INI],
                       cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 1 and "Could not open database" in r.stdout


def test_missing_mask_file_exits_with_one(built, tmp_path):
    # maximal / layout before filter: exit code 1 like every other error path; the CUDA context that is
    # being created beside the input reading must be joined, not left to std::terminate (SIGABRT)
    root, meta = ht.materialize("dal_small", str(tmp_path))
    for stage in ("maximal", "layout"):
        r = ht.run_stage("product", stage, str(tmp_path), root, "nofilter", check=False)
        assert r.returncode == 1, (stage, r.returncode, r.stdout[-500:])
        assert "cannot read nofilter.mas" in r.stdout
