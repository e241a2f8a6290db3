"""hg_order.h restates libstdc++'s std::sort; the native test compares it with
the real thing element for element (ties, killer sequences, both directions)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_std_sort_exact_matches_libstdcxx(tmp_path):
    exe = str(tmp_path / "test_order")
    subprocess.run(["g++", "-O2", "-std=gnu++17", os.path.join(ROOT, "tests", "native", "test_order.cpp"), "-o", exe],
                   check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stdout
    assert r.stdout.startswith("OK"), r.stdout
