"""Times the `hinge filter` executable on synthetic samples (phase breakdown on stderr): how the time to
create the CUDA context depends on the size of the input, on the host threads of the ingest running
beside it, and on an nvidia-smi query loop in the background (bench.py's clock sampler).
    python scripts/cli_probe.py [genome Mb ...]"""
import os, subprocess, sys, tempfile, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import hgsynth
exe = os.path.join(ROOT, "hinge_b200", "_build", "hinge")
ini = os.path.join(ROOT, "tests", "golden", "nominal.ini")


def run(work, env_extra, smi=False):
    env = dict(os.environ, HINGE_B200_TIMING="1", **env_extra)
    loop = None
    if smi:
        loop = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader", "-lms", "100"],
                                stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        time.sleep(0.3)
    t0 = time.perf_counter()
    r = subprocess.run([exe, "filter", "--db", "S", "--las", "S.las", "-x", "gpu", "--config", ini], cwd=work,
                       stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env)
    wall = time.perf_counter() - t0
    if loop:
        loop.terminate()
    keep = [ln.split("]")[1].strip() for ln in r.stderr.splitlines() if "timing]" in ln and
            any(k in ln for k in ("context", "read db", "hg_set_overlaps", "hg_filter", "fetch", "write output"))]
    print("wall %.3f s  %-32s smi=%d | %s" % (wall, env_extra, smi, " | ".join(" ".join(k.split()) for k in keep)), flush=True)


for mb in [float(x) for x in sys.argv[1:]] or [4.0, 64.0]:
    work = tempfile.mkdtemp(prefix="cli_probe_")
    s = hgsynth.Synth(genome_len=int(mb * 1e6), coverage=40.0, read_mean=24000, read_sd=8000, read_min=2000, seed=4321,
                      frag_prob=1.2)
    n = s.generate(want_trace=True, threads=os.cpu_count() or 8)
    s.write_db(work, "S", with_bps=True, with_qv=True); s.write_las(os.path.join(work, "S.las")); s.close()
    print("==== %g Mb: overlaps %d, las %.0f MB" % (mb, n, os.path.getsize(os.path.join(work, "S.las")) / 1e6), flush=True)
    run(work, {}, False)   # pages everything in
    for smi in (False, True):
        for extra in ({}, {"HINGE_B200_IO_THREADS": "2"}):
            run(work, extra, smi)
    shutil.rmtree(work, ignore_errors=True)
