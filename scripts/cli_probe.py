"""Times the `hinge filter` executable on a synthetic sample (phase breakdown on stderr)."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import hgsynth
mb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
work = tempfile.mkdtemp(prefix="cli_probe_")
s = hgsynth.Synth(genome_len=int(mb * 1e6), coverage=50.0, read_mean=3500, read_sd=1500, read_min=1000, seed=1234)
n = s.generate(want_trace=True, threads=os.cpu_count() or 8)
s.write_db(work, "S", with_bps=True, with_qv=True); s.write_las(os.path.join(work, "S.las")); s.close()
print("overlaps", n, "las MB", os.path.getsize(os.path.join(work, "S.las")) / 1e6)
exe = os.path.join(ROOT, "hinge_b200", "_build", "hinge")
ini = os.path.join(ROOT, "tests", "golden", "nominal.ini")
for env_extra in ({}, {}, {"HINGE_B200_IO_THREADS": "1"}):
    env = dict(os.environ, HINGE_B200_TIMING="1", **env_extra)
    t0 = time.perf_counter()
    r = subprocess.run([exe, "filter", "--db", "S", "--las", "S.las", "-x", "gpu", "--config", ini], cwd=work,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env)
    print("---- wall %.3f s %s" % (time.perf_counter() - t0, env_extra))
    print(r.stderr)
