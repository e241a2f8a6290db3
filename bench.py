#!/usr/bin/env python3
"""bench.py — overlap-records/s through the HINGE hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU stage)

A *step* is one pass of the `hinge filter` stage (coverage profiles + estimate, masks,
repeat annotation, hinge calls) over one batch of synthetic overlap records:

  value     overlaps/s with the struct-of-arrays already resident in HBM,
            timed with CUDA events on the context's stream, max over ranks
  e2e       the same stage through the C ABI with HOST buffers: pinned-host ->
            device copy of the records, the kernels, and the device -> host
            read of the results, all inside the timed region
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event time,
            against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the unmodified reference `Reads_filter` (oracle/_ref) timed on
            a bounded sample of the same workload on this box's host cores
  cli_filter    (informational) the product's own `hinge filter` executable on
            that same sample, file to file, with its phase breakdown

Workload = BASELINE.json configs[2]: synthetic 50 Mb genome, 50x, reads
N(3500,1500) >= 1000 bp, ~52 M overlaps per GPU (weak scaling: the genome grows
with the number of GPUs; reads shard by A-read id; a 16 KB coverage histogram is
all-reduced and the masks, 4 B per read, are all-gathered over NCCL between the
phases of the stage).
"""
import argparse
import ctypes
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

METRIC = "overlap-records/sec through filter+hinge"
UNIT = "overlaps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--genome-mb", type=float, default=50.0, help="genome size per GPU (Mb)")
    ap.add_argument("--cov", type=float, default=50.0)
    ap.add_argument("--read-mean", type=int, default=3500)
    ap.add_argument("--read-sd", type=int, default=1500)
    ap.add_argument("--sample-mb", type=float, default=4.0, help="genome size of the CPU-baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--spread", type=int, default=None, help="HG_OPT_SCATTER_SPREAD (tuning aid)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-downstream", action="store_true", help="skip the maximal/layout timing")
    return ap.parse_args()


def workload_name(args, world):
    return ("synthetic %g Mb genome x %d GPU(s), %gx, reads N(%d,%d)>=1000, planted repeats; step = hinge "
            "filter (coverage estimate + masks + repeat annotation + hinge calls)"
            % (args.genome_mb, world, args.cov, args.read_mean, args.read_sd))


def synth_kwargs(args, genome_mb):
    return dict(genome_len=int(genome_mb * 1e6), coverage=args.cov, read_mean=args.read_mean,
                read_sd=args.read_sd, read_min=1000, seed=args.seed)


def host_threads(world=1):
    return max(1, (os.cpu_count() or 8) // max(1, world))


# ----------------------------------------------------------------------------- clocks


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm


def time_reference_filter(args, steps, warmup, with_cli=False):
    """Times the reference's own `Reads_filter` (1 thread: the reference has no parallel region) on a bounded
    sample of the workload.  Falls back to the oracle port when oracle/_ref is absent."""
    import hgsynth

    ref = os.path.join(ROOT, "oracle", "_ref", "bin", "Reads_filter")
    kind = "reference"
    if not os.path.exists(ref):
        ref = os.path.join(ROOT, "oracle", "_build", "hinge_oracle")
        kind = "port"
    ini = os.path.join(ROOT, "tests", "golden", "nominal.ini")
    work = tempfile.mkdtemp(prefix="hinge_bench_ref_")
    try:
        s = hgsynth.Synth(**synth_kwargs(args, args.sample_mb))
        novl = s.generate(want_trace=True, threads=host_threads())
        s.write_db(work, "S", with_bps=True, with_qv=True)
        s.write_las(os.path.join(work, "S.las"))
        n_read = s.n_read
        s.close()
        cmd = [ref] + (["filter"] if kind == "port" else []) + ["--db", "S", "--las", "S.las", "-x", "ref",
                                                                 "--config", ini]
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            subprocess.run(cmd, cwd=work, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        cli = None
        if with_cli:
            # the same files through the product's own `hinge filter` (process start, CUDA context, .las
            # parse, H2D, kernels, D2H, all output files): the drop-in, file-to-file comparison
            exe = os.path.join(ROOT, "hinge_b200", "_build", "hinge")
            mine = [exe, "filter", "--db", "S", "--las", "S.las", "-x", "gpu", "--config", ini]
            env = dict(os.environ, HINGE_B200_TIMING="1")
            best, phases = None, ""
            for _ in range(2):
                t0 = time.perf_counter()
                r = subprocess.run(mine, cwd=work, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                                   text=True, env=env)
                dt = time.perf_counter() - t0
                if best is None or dt < best:
                    best, phases = dt, r.stderr
            same = all(open(os.path.join(work, "gpu." + e), "rb").read() == open(os.path.join(work, "ref." + e), "rb").read()
                       for e in ("mas", "cmas", "repeat.txt", "hinges.txt", "coverage.txt"))
            cli = {"seconds": best, "overlaps_per_s": novl / best, "reference_seconds": sum(times) / len(times),
                   "outputs_identical_to_reference": same,
                   "phases_ms": {ln.split("]")[1].rsplit(None, 2)[0].strip(): float(ln.split()[-2])
                                 for ln in phases.splitlines() if "timing]" in ln}}
    finally:
        shutil.rmtree(work, ignore_errors=True)
    sec = sum(times) / len(times)
    return {
        "value": novl / sec, "unit": UNIT, "cores": 1, "kind": kind, "seconds_per_pass": sec,
        "sample": "%s on a %g Mb / %gx sample of the workload (%d reads, %d overlaps, .las on page cache -> "
                  "output files), single thread (the reference has no parallel region), %d host cores available"
                  % (os.path.basename(ref), args.sample_mb, args.cov, n_read, novl, os.cpu_count() or 0),
        "cli": cli,
    }, novl


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    t0 = time.perf_counter()
    base, novl = time_reference_filter(args, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * base["seconds_per_pass"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus), "sample": base["sample"]},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- downstream stages


def time_downstream(args, syn, ctx, api, filt, np, torch):
    """hg_maximal and hg_layout on the same batch, fed with the filter's results (host buffers, traces
    included): wall time of the C-ABI call and device time of its kernels.  Not part of `value`."""
    t0 = time.perf_counter()
    novl = syn.generate(want_trace=True, threads=host_threads())
    cols = syn.cols()
    trace_off, trace = syn.trace()
    gen_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    ctx.set_overlaps(novl, cols, trace_off=trace_off, trace=trace, tbytes=1, where=api.HG_MEM_HOST)
    torch.cuda.synchronize()
    load_s = time.perf_counter() - t0
    lp = api.LayoutParams()
    mask = filt["mask"]
    out = {"overlaps": novl, "trace_bytes": int(trace_off[-1]), "generate_s": round(gen_s, 2), "h2d_s": round(load_s, 3)}
    os.environ.setdefault("HINGE_B200_SKIP_CONTAINED_TXT", "1")
    for rep in range(2):
        t0 = time.perf_counter()
        maximal, _, ms_dev = ctx.maximal(lp, mask)
        out["maximal"] = {"wall_ms": 1e3 * (time.perf_counter() - t0), "device_ms": ms_dev,
                          "maximal_reads": int(maximal.sum())}
    n = len(mask)
    off = filt["anno_off"]
    keep = filt["hinge_keep"].astype(bool)
    per_read = np.repeat(np.arange(n), np.diff(off))
    hin_off = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(per_read[keep], minlength=n), out=hin_off[1:])
    rep_csr = (off, filt["anno_pos"], filt["anno_type"])
    hin_csr = (hin_off, filt["anno_pos"][keep], filt["anno_type"][keep])
    for rep in range(2):
        t0 = time.perf_counter()
        edges, ms_dev = ctx.layout(lp, mask, maximal, rep_csr, hin_csr)
        out["layout"] = {"wall_ms": 1e3 * (time.perf_counter() - t0), "device_ms": ms_dev, "edges": len(edges)}
    return out


# ----------------------------------------------------------------------------- our arm


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    # the contract is ONE JSON line on stdout: libraries that chat there (NCCL prints its version at
    # the first communicator) are sent to stderr until the line is ready
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist

    import hgsynth
    import hinge_b200 as hb
    from hinge_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product path has no CPU fallback")
    # CPU baseline + the file-to-file run of the product's own executable on the same sample: first,
    # while this process holds no CUDA context yet (a second context on a busy GPU takes seconds to
    # create and would be charged to the executable)
    base = None
    if world == 1 and not args.no_cpu_baseline:
        base, _ = time_reference_filter(args, 1, 0, with_cli=True)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # ---- synthetic batch: every rank derives the same global read table, then generates only the
    # records of its own A-read range
    t_gen = time.perf_counter()
    syn = hgsynth.Synth(**synth_kwargs(args, args.genome_mb * world))
    n_read = syn.n_read
    chunk = (n_read + world - 1) // world
    a_lo, a_hi = rank * chunk, min(n_read, (rank + 1) * chunk)
    novl = syn.generate(a_lo, a_hi, want_trace=False, threads=host_threads(world))
    cols_np = syn.cols()
    names = ["aread", "bread", "abpos", "aepos", "bbpos", "bepos", "flags"]
    t_gen = time.perf_counter() - t_gen

    stream = torch.cuda.current_stream()
    ctx = api.Context(local, stream.cuda_stream)
    ctx.set_option(api.HG_OPT_PROFILE, 1)
    if args.spread is not None:
        ctx.set_option(api.HG_OPT_SCATTER_SPREAD, args.spread)
    ctx.set_reads(syn.rlen, syn.qv_off, syn.qv, 100)
    params = api.FilterParams()

    # per-read arrays that cross shards live in torch tensors so NCCL can all-gather them in place
    from hinge_b200.sharding import ShardedArrays, run_filter_sharded

    arrays = ShardedArrays(n_read, rank, world, dev)
    assert (arrays.lo, arrays.hi) == (a_lo, a_hi)
    arrays.bind(ctx)

    def run_stage():
        """The sharded form of hg_filter: phase1 | all-gather means | phase2 | all-gather masks | phase3."""
        rc, s = run_filter_sharded(ctx, params, arrays)
        if rc == api.HG_RETRY_POOL:
            raise RuntimeError("annotation pool overflow")
        return s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- arm 1: records resident in HBM
    cols_dev = {k: torch.from_numpy(cols_np[k]).to(dev) for k in names}
    ctx.set_overlaps(novl, cols_dev, where=api.HG_MEM_DEVICE, a_lo=a_lo, a_hi=a_hi)
    # the timed region lasts milliseconds: the sampler starts ahead of the warm-up so that nvidia-smi is
    # already reporting while the same kernels run
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        time.sleep(0.3)
    for _ in range(max(3, args.warmup)):
        summary = run_stage()
    barrier()
    launches0 = api.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ktimes = {}
    ev0.record(stream)
    for _ in range(args.steps):
        summary = run_stage()
        for k, v in ctx.filter_kernel_times().items():
            ktimes.setdefault(k, []).append(v)
    ev1.record(stream)
    barrier()
    launches = api.launch_count() - launches0
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    clocks = sampler.stop() if sampler else None
    total_ovl = sum_over_ranks(float(novl))
    ms_step = ms_total / args.steps
    value = total_ovl / (ms_step * 1e-3)

    # ---- arm 2: end to end through the C ABI with host buffers
    pinned = {k: torch.from_numpy(cols_np[k]).pin_memory() for k in names}
    h2d = sum(t.numel() * t.element_size() for t in pinned.values())
    e2e_times, d2h = [], 0
    for i in range(1 + args.e2e_steps):
        barrier()
        t0 = time.perf_counter()
        ctx.set_overlaps(novl, pinned, where=api.HG_MEM_HOST, a_lo=a_lo, a_hi=a_hi)
        s = run_stage()
        res = ctx.filter_fetch(int(s.n_annotations))
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        d2h = sum(v.nbytes for v in res.values())
        if i > 0:
            e2e_times.append(dt)
    e2e_s = sum(e2e_times) / len(e2e_times)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        kavg = {k: sum(v) / len(v) for k, v in ktimes.items()}
        # coverage bins of the owned reads (cut_off 300, filter.cpp:386: 40-bp bins)
        summary_bins = int(((syn.rlen[a_lo:a_hi].astype(np.int64) + params.cut_off) // 40 + 3).sum())
        owned = a_hi - a_lo
        # algorithmic bytes of one launch.  K1 (profile build) reads aread / abpos /
        # aepos (12 B per record; bread only in batches with self-overlaps), ~30 B per read of offsets,
        # lengths and plan, and writes the scanned profiles (4 B per coverage bin) and 9 B per read;
        # K2 (mask + annotation) reads the profiles back plus ~60 B per read of inputs and results.
        bins = float(summary_bins)
        kbytes = {"profile": 12.0 * novl + 4.0 * bins + 39.0 * owned, "mask_anno": 4.0 * bins + 60.0 * owned}
        dom = max(kbytes, key=lambda k: kavg[k])
        achieved = kbytes[dom] / (kavg[dom] * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args, world), "n_reads": n_read, "overlaps": int(total_ovl),
                       "overlaps_per_gpu": novl, "l2": "inputs (%.0f MB of records per GPU) exceed the 126 MB L2"
                       % (28.0 * novl / 1e6), "parallelism": "reads sharded by A-read id x%d" % world,
                       "cov_est": int(summary.cov_est), "annotations_rank0": int(summary.n_annotations),
                       "generate_s": round(t_gen, 2)},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": kbytes[dom], "kernel_ms": kavg[dom]},
            "kernel_ms": kavg,
            "filter_scan": {"bytes_per_overlap": 32, "gbs": 32.0 * novl / (ms_step * 1e-3) / 1e9,
                            "frac_of_peak": 32.0 * novl / (ms_step * 1e-3) / 1e9 / peak},
            "e2e": {"value": total_ovl / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_s},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if base is not None:
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
            # informational: the same sample, file to file, through the product's `hinge filter` executable
            line["cli_filter"] = base["cli"]
    if world == 1 and not args.no_downstream:
        # informational: the two stages downstream of the filter on the same batch (they need the trace)
        line["downstream_stages"] = time_downstream(args, syn, ctx, api, res, np, torch)
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
