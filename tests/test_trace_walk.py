"""The trace walk of maximal / layout (classify_record, hinge_b200/csrc/hg_layout.cu) against a plain
restatement of the reference's ProcessAlignment on random matches, masks and traces.  The function's SOURCE
TEXT is cut out of the .cu file and compiled for the host (tests/native/test_walk.cpp): no GPU needed, and it
is the product's code that runs, not a copy."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_counting_walk_matches_the_reference_walk(tmp_path):
    src = open(os.path.join(ROOT, "hinge_b200", "csrc", "hg_layout.cu")).read()
    a = src.index("struct Match {")
    b = src.index("__device__ __forceinline__ int raw_length")
    body = src[a:b]
    assert "classify_record(" in body and "trace_value" in body
    (tmp_path / "classify_record.inc").write_text(body)
    exe = str(tmp_path / "test_walk")
    subprocess.run(["g++", "-O2", "-std=gnu++17", "-I", str(tmp_path), "-I", os.path.join(ROOT, "hinge_b200", "csrc"),
                    os.path.join(ROOT, "tests", "native", "test_walk.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "400000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"active (\d+) mismatches (\d+)", r.stdout)
    assert m and int(m.group(2)) == 0 and int(m.group(1)) > 10000, r.stdout
