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


def test_hash_iteration_order_matches_libstdcxx(tmp_path):
    """The iteration order of std::unordered_map<int, T> (candidate pre-sort order of the layout stage,
    hinging.cpp:532) restated in hg_order.h vs the real container."""
    exe = str(tmp_path / "test_hash_order")
    subprocess.run(["g++", "-O2", "-std=gnu++17", os.path.join(ROOT, "tests", "native", "test_hash_order.cpp"),
                    "-o", exe], check=True)
    r = subprocess.run([exe], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and "identical to std::unordered_map" in r.stdout, r.stdout
