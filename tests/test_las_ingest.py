"""The chunked, verified .las ingest returns what one sequential walk returns (native test)."""
import lzma
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chunked_ingest_matches_sequential_walk(tmp_path):
    import sys

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import handmade as hm
    import numpy as np

    exe = str(tmp_path / "test_las")
    subprocess.run(["g++", "-O2", "-std=gnu++17", "-pthread", os.path.join(ROOT, "tests", "native", "test_las.cpp"),
                    os.path.join(ROOT, "hinge_b200", "csrc", "hg_io.cpp"), "-o", exe], check=True)
    files = []
    # a real daligner file
    real = str(tmp_path / "D.las")
    with lzma.open(os.path.join(ROOT, "tests", "golden", "dal_small", "D.las.xz")) as f, open(real, "wb") as g:
        g.write(f.read())
    files.append(real)
    # hand-made: 8-bit and 16-bit traces, many short records
    rng = np.random.default_rng(11)
    for tspace, root in ((100, "A"), (200, "B")):
        rlen = [int(x) for x in rng.integers(5000, 9000, 30)]
        recs = []
        for a in range(30):
            for b in range(30):
                if a != b and (a * 7 + b) % 3 == 0:
                    ab = int(rng.integers(0, 2000))
                    ln = int(rng.integers(1000, 2900))
                    bb = int(rng.integers(0, 2000))
                    recs.append((a, b, ab, ab + ln, bb, bb + ln - int(rng.integers(0, 40)), int(rng.integers(0, 2))))
        hm.write_fixture(str(tmp_path), root, rlen, recs, tspace=tspace, qv=None)
        files.append(str(tmp_path / (root + ".las")))
    r = subprocess.run([exe] + files, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout
    print(r.stdout)
