#!/usr/bin/env python3
"""bench.py — overlap-records/s through the HINGE hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (our arm; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's CPU stage)

A *step* is one pass of the `hinge filter` stage (coverage profiles + estimate, masks,
repeat annotation, hinge calls) over one batch of synthetic overlap records.

Workloads (BASELINE.json configs):
  c5 (default, the line's `value`)  configs[4]: 300 Mb genome, 40x, reads N(24000,8000) >= 2000,
      long overlaps reported as several local alignments -> ~100 M overlap records.  The set is
      FIXED: N GPUs shard it by A-read id, balanced on record volume ("scaling": "strong").
  c3 (reported under `c3`)          configs[2]: 50 Mb genome per GPU, 50x, reads N(3500,1500)
      >= 1000, ~62 M records per GPU; the genome grows with N (weak scaling).

Per workload:
  value     overlaps/s with the struct-of-arrays already resident in HBM, timed with CUDA events
            on the context's stream, max over ranks
  parity    outside the timed region: the results of the timed configuration are compared with
            the CPU oracle run on the same batch (oracle/, test infrastructure) and, for N > 1,
            rank 0 also reruns the whole set on one context and compares the gathered shards
  roofline  the dominant kernel's algorithmic bytes / its CUDA-event time against the measured
            HBM copy bandwidth (MEASURED_PEAKS.json)
  e2e_arrays  the stage through the array-level C ABI with pinned HOST buffers (H2D + kernels + D2H)

Once per line (N = 1):
  e2e       file to file: the product's `hinge filter` executable (.db/.las on the page cache ->
            all output files) on a bounded sample of the c5 workload -- the SAME files the
            reference arm / cpu_baseline run `Reads_filter` on, outputs compared byte for byte
  cpu_baseline  the unmodified reference `Reads_filter` (oracle/_ref) on that sample
"""
import argparse
import json
import math
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "overlap-records/sec through filter+hinge"
UNIT = "overlaps/s"
NAMES = ["aread", "bread", "abpos", "aepos", "bbpos", "bepos", "flags"]

PRESETS = {
    # BASELINE.json configs[4] / SURVEY.md section 8(d) "S-100M"
    "c5": dict(genome_mb=300.0, cov=40.0, read_mean=24000, read_sd=8000, read_min=2000, frag=1.2, seed=4321,
               scaling="strong", sample_mb=64.0,
               name="synthetic 300 Mb genome, 40x, reads N(24000,8000)>=2000, ~100 M overlaps (BASELINE configs[4])"),
    # BASELINE.json configs[2] / "S-50M"
    "c3": dict(genome_mb=50.0, cov=50.0, read_mean=3500, read_sd=1500, read_min=1000, frag=0.0, seed=1234,
               scaling="weak", sample_mb=8.0,
               name="synthetic 50 Mb genome per GPU, 50x, reads N(3500,1500)>=1000, ~62 M overlaps per GPU "
                    "(BASELINE configs[2])"),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c5", choices=sorted(PRESETS), help="workload of the line's `value`")
    ap.add_argument("--also", default=None, help="second workload reported under its own key ('' = none); "
                    "default: the other preset")
    ap.add_argument("--genome-mb", type=float, default=None, help="override the preset's genome size (Mb)")
    ap.add_argument("--cov", type=float, default=None)
    ap.add_argument("--read-mean", type=int, default=None)
    ap.add_argument("--read-sd", type=int, default=None)
    ap.add_argument("--frag", type=float, default=None)
    ap.add_argument("--sample-mb", type=float, default=None, help="genome size of the file-to-file / CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--sync-steps", action="store_true", help="one blocking hg_filter per timed step instead of "
                    "hg_filter_enqueue back to back + one hg_filter_finish")
    ap.add_argument("--spread", type=int, default=None, help="HG_OPT_SCATTER_SPREAD (tuning aid)")
    ap.add_argument("--profile-kernel", type=int, default=None, help="HG_OPT_PROFILE_KERNEL (tuning aid)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: phase exchange through NVLink peer memory inside the kernels, or NCCL calls")
    ap.add_argument("--no-verify", action="store_true", help="skip the parity legs")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip cpu_baseline and the file-to-file e2e")
    ap.add_argument("--no-downstream", action="store_true", help="skip the maximal/layout timing")
    return ap.parse_args()


def preset(args, name):
    p = dict(PRESETS[name])
    for k in ("genome_mb", "cov", "read_mean", "read_sd", "frag", "seed", "sample_mb"):
        v = getattr(args, k, None)
        if v is not None and name == args.config:
            p[k] = v
    return p


def synth_kwargs(p, genome_mb):
    return dict(genome_len=int(genome_mb * 1e6), coverage=p["cov"], read_mean=p["read_mean"],
                read_sd=p["read_sd"], read_min=p["read_min"], seed=p["seed"], frag_prob=p["frag"])


def host_threads(world=1):
    return max(1, (os.cpu_count() or 8) // max(1, world))


def bind_to_gpu_numa(torch, index):
    """Ranks of a multi-GPU run stay on the cores of their GPU's NUMA node, so that the pinned host buffers
    of the end-to-end leg (first touch) and the copies out of them are local to the GPU's PCIe root."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError):
        return None


# ----------------------------------------------------------------------------- clocks


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval_ms=20):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", str(interval_ms)],
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


# ----------------------------------------------------------------------------- file-to-file legs


def parse_phases(stderr_text):
    out = {}
    for ln in stderr_text.splitlines():
        if "timing]" not in ln:
            continue
        try:
            out[ln.split("]")[1].rsplit(None, 2)[0].strip()] = float(ln.split()[-2])
        except (IndexError, ValueError):
            pass
    return out


def file_to_file(p, steps, warmup, with_cli):
    """The reference's own `Reads_filter` (1 thread: it has no parallel region) and, optionally, the
    product's `hinge filter` executable on the same files: a bounded sample of the workload written
    as a DAZZ_DB + .las.  Falls back to the oracle port when oracle/_ref is absent."""
    import hgsynth

    ref = os.path.join(ROOT, "oracle", "_ref", "bin", "Reads_filter")
    kind = "reference"
    if not os.path.exists(ref):
        ref = os.path.join(ROOT, "oracle", "_build", "hinge_oracle")
        kind = "port"
    ini = os.path.join(ROOT, "tests", "golden", "nominal.ini")
    work = tempfile.mkdtemp(prefix="hinge_bench_f2f_")
    try:
        s = hgsynth.Synth(**synth_kwargs(p, p["sample_mb"]))
        novl = s.generate(want_trace=True, threads=host_threads())
        s.write_db(work, "S", with_bps=True, with_qv=True)
        s.write_las(os.path.join(work, "S.las"))
        n_read = s.n_read
        s.close()
        las_bytes = os.path.getsize(os.path.join(work, "S.las"))
        cli = None
        if with_cli:
            # first, while nothing else holds the GPU: process start, CUDA context, .las parse, H2D,
            # kernels, D2H, all output files
            exe = os.path.join(ROOT, "hinge_b200", "_build", "hinge")
            mine = [exe, "filter", "--db", "S", "--las", "S.las", "-x", "gpu", "--config", ini]
            env = dict(os.environ, HINGE_B200_TIMING="1")
            runs = []
            # clocks are sampled over this timed region too; the running nvidia-smi also keeps the driver
            # attached to the GPU between the runs (on a box without persistence mode every process start
            # would otherwise pay the driver's own cold start, ~0.7 s, which is not the executable's doing)
            # (sampled every 250 ms only: NVML queries at the 20 ms rate of the kernel-timed region contend with
            # the creation of the executable's CUDA context and were seen to double it, scripts/cli_probe.py)
            sampler = ClockSampler(0, interval_ms=250)
            time.sleep(0.3)
            for _ in range(3):
                t0 = time.perf_counter()
                r = subprocess.run(mine, cwd=work, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                                   text=True, env=env)
                runs.append((time.perf_counter() - t0, r.stderr))
            runs = runs[1:]  # the first run pages the executable and the CUDA libraries in
            best, phases = min(runs, key=lambda x: x[0])
            cli_clocks = sampler.stop()
            cli = {"seconds": best, "clocks": cli_clocks, "seconds_all": [round(x[0], 3) for x in runs], "overlaps_per_s": novl / best,
                   "phases_ms": parse_phases(phases)}
        cmd = [ref] + (["filter"] if kind == "port" else []) + ["--db", "S", "--las", "S.las", "-x", "ref",
                                                                 "--config", ini]
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            subprocess.run(cmd, cwd=work, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        if cli:
            exts = ("mas", "cmas", "repeat.txt", "hinges.txt", "coverage.txt")
            cli["outputs_identical_to_reference"] = all(
                open(os.path.join(work, "gpu." + e), "rb").read() == open(os.path.join(work, "ref." + e), "rb").read()
                for e in exts)
            cli["files_compared"] = list(exts)
            cli["reference_seconds"] = sum(times) / len(times)
    finally:
        shutil.rmtree(work, ignore_errors=True)
    sec = sum(times) / len(times)
    sample = ("%s on a %g Mb sample of the workload (%d reads, %d overlaps, %.2f GB .las with traces on the page "
              "cache -> all output files), single thread (the reference has no parallel region), %d host cores "
              "available" % (os.path.basename(ref), p["sample_mb"], n_read, novl, las_bytes / 1e9, os.cpu_count() or 0))
    return {"value": novl / sec, "unit": UNIT, "cores": 1, "kind": kind, "seconds_per_pass": sec, "sample": sample,
            "sample_overlaps": novl, "las_bytes": las_bytes, "cli": cli}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    p = preset(args, args.config)
    steps, warmup = max(1, min(args.steps, 2)), min(args.warmup, 1)
    t0 = time.perf_counter()
    base = file_to_file(p, steps, warmup, with_cli=False)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * base["seconds_per_pass"],
        "higher_is_better": True, "scaling": p["scaling"], "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": p["name"], "sample": base["sample"]},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- downstream stages


def time_downstream(syn, ctx, api, filt, np, torch):
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


FIELDS = ("mask", "cmask", "flags", "anno_off", "anno_pos", "anno_type", "hinge_keep")


def compare_results(np, got, want, lo=0, hi=None):
    """Element-for-element comparison of two filter results on the reads [lo, hi); returns the names
    of the arrays that differ.  `got` may hold only the annotations of its own reads."""
    n = len(want["anno_off"]) - 1
    hi = n if hi is None else hi
    return compare_slice(np, slice_result(np, got, lo, hi), want)


def slice_result(np, res, lo, hi):
    """The part of a filter result that belongs to the reads [lo, hi) (what a rank sends for checking)."""
    a0, a1 = int(res["anno_off"][lo]), int(res["anno_off"][hi])
    return {"lo": lo, "hi": hi, "mask": res["mask"][lo:hi].copy(), "cmask": res["cmask"][lo:hi].copy(),
            "flags": res["flags"][lo:hi].copy(), "anno_cnt": np.diff(res["anno_off"][lo:hi + 1]),
            "anno_pos": res["anno_pos"][a0:a1].copy(), "anno_type": res["anno_type"][a0:a1].copy(),
            "hinge_keep": res["hinge_keep"][a0:a1].copy()}


def compare_slice(np, part, want):
    lo, hi = part["lo"], part["hi"]
    bad = [k for k in ("mask", "cmask", "flags") if not np.array_equal(part[k], want[k][lo:hi])]
    if not np.array_equal(part["anno_cnt"], np.diff(want["anno_off"][lo:hi + 1])):
        bad.append("anno_off")
    else:
        w0, w1 = want["anno_off"][lo], want["anno_off"][hi]
        bad += [k for k in ("anno_pos", "anno_type", "hinge_keep") if not np.array_equal(part[k], want[k][w0:w1])]
    return bad


def digest(np, a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


class Bench:
    def __init__(self, args):
        import numpy as np
        import torch
        import torch.distributed as dist

        import hgsynth
        from hinge_b200 import api, sharding

        self.np, self.torch, self.dist, self.hgsynth, self.api, self.sharding = np, torch, dist, hgsynth, api, sharding
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream()
        self.numa = bind_to_gpu_numa(torch, self.local) if self.world > 1 else None
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            self.peak, self.peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            self.peak, self.peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    # ---- collectives on scalars
    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def _reduce(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.MAX)

    def sum_over_ranks(self, x):
        return self._reduce(x, self.dist.ReduceOp.SUM)

    # ---- one workload
    def run(self, pname, with_clocks, downstream):
        np, torch, api, args = self.np, self.torch, self.api, self.args
        p = preset(args, pname)
        world, rank = self.world, self.rank
        strong = p["scaling"] == "strong"
        t_gen = time.perf_counter()
        syn = self.hgsynth.Synth(**synth_kwargs(p, p["genome_mb"] * (1 if strong else world)))
        n_read = syn.n_read
        rlen = syn.rlen
        # shards: contiguous A-read ranges of equal record VOLUME (pile-up depth goes with read length,
        # so the cut is on cumulative read length; a .las front-end cuts on its record offsets instead)
        arrays = self.sharding.ShardedArrays(n_read, rank, world, self.dev, weights=rlen,
                                             equal_slices=(args.exchange == "nccl"))
        a_lo, a_hi = arrays.lo, arrays.hi
        novl = syn.generate(a_lo, a_hi, want_trace=False, threads=host_threads(world))
        cols_np = syn.cols()  # views into the generator's buffers: valid until its next generate()
        pile = np.bincount(cols_np["aread"] - a_lo, minlength=a_hi - a_lo)[:a_hi - a_lo]  # records per owned read
        t_gen = time.perf_counter() - t_gen

        ctx = api.Context(self.local, self.stream.cuda_stream)
        ctx.set_option(api.HG_OPT_PROFILE, 1)
        if args.spread is not None:
            ctx.set_option(api.HG_OPT_SCATTER_SPREAD, args.spread)
        if args.profile_kernel is not None:
            ctx.set_option(api.HG_OPT_PROFILE_KERNEL, args.profile_kernel)
        ctx.set_reads(rlen, syn.qv_off, syn.qv, 100)
        params = api.FilterParams()
        arrays.bind(ctx, exchange=args.exchange)

        def run_stage():
            rc, s = self.sharding.run_filter_sharded(ctx, params, arrays)
            if rc != 0:
                raise RuntimeError("hg_filter: status %d" % rc)
            return s

        # ---- arm 1: records resident in HBM
        cols_dev = {k: torch.from_numpy(cols_np[k]).to(self.dev) for k in NAMES}
        ctx.set_overlaps(novl, cols_dev, where=api.HG_MEM_DEVICE, a_lo=a_lo, a_hi=a_hi)
        arrays.set_global_range(ctx, int(cols_np["aread"][0]), int(cols_np["aread"][-1]))
        # the timed region lasts milliseconds: the sampler starts ahead of the warm-up so that nvidia-smi
        # is already reporting while the same kernels run
        sampler = ClockSampler(self.local) if (rank == 0 and with_clocks) else None
        if sampler:
            time.sleep(0.3)
        warm = max(3, args.warmup)
        for _ in range(warm):
            summary = run_stage()
        self.barrier()
        launches0 = api.launch_count()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ktimes = {}
        # The timed steps are enqueued back to back (hg_filter_enqueue: the launches of one stage) and waited
        # for once (hg_filter_finish): device time of K stages without K host round trips; with the NCCL
        # exchange the host takes part in every step, so there each step is a full hg_filter.
        queued = (world == 1 or arrays.exchange == "peer") and not args.sync_steps
        ev0.record(self.stream)
        if queued:
            for _ in range(args.steps):
                ctx.filter_enqueue(params)
            ev1.record(self.stream)
            rc, summary = ctx.filter_finish()
            if rc != 0:
                raise RuntimeError("hg_filter_finish: status %d" % rc)
        else:
            for _ in range(args.steps):
                summary = run_stage()
            ev1.record(self.stream)
        self.barrier()
        launches = api.launch_count() - launches0
        ms_total = self.max_over_ranks(ev0.elapsed_time(ev1))
        # per-kernel device times (CUDA events between the launches): a few more steps, one at a time
        for _ in range(min(args.steps, 5)):
            summary = run_stage()
            for k, v in ctx.filter_kernel_times().items():
                ktimes.setdefault(k, []).append(v)
        clocks = sampler.stop() if sampler else None
        total_ovl = self.sum_over_ranks(float(novl))
        ms_step = ms_total / args.steps
        result = ctx.filter_fetch(int(summary.n_annotations))  # of the timed configuration: what parity checks

        # ---- arm 2: through the array-level C ABI with pinned host buffers
        pinned = {k: torch.from_numpy(cols_np[k]).pin_memory() for k in NAMES}
        h2d = sum(t.numel() * t.element_size() for t in pinned.values())
        e2e_times, d2h = [], 0
        for i in range(1 + args.e2e_steps):
            self.barrier()
            t0 = time.perf_counter()
            ctx.set_overlaps(novl, pinned, where=api.HG_MEM_HOST, a_lo=a_lo, a_hi=a_hi)
            s = run_stage()
            res = ctx.filter_fetch(int(s.n_annotations))
            torch.cuda.synchronize()
            dt = self.max_over_ranks(time.perf_counter() - t0)
            d2h = sum(v.nbytes for v in res.values())
            if i > 0:
                e2e_times.append(dt)
        e2e_s = sum(e2e_times) / len(e2e_times)
        del pinned

        # ---- parity (outside every timed region)
        parity = None
        if not args.no_verify:
            parity = self.verify(pname, p, syn, arrays, summary, result, cols_np, novl)

        out = None
        if rank == 0:
            kavg = {k: sum(v) / len(v) for k, v in ktimes.items()}
            owned = a_hi - a_lo
            # algorithmic bytes of one launch.  K1 (profile build) reads aread / abpos /
            # aepos (12 B per record; bread only in batches with self-overlaps), ~30 B per read of offsets,
            # lengths and plan, and writes the packed profiles (4 B per 40-bp coverage bin, cut_off 300:
            # (rlen + 300) / 40 + 3 bins per read) and 9 B per read; K2 (mask + annotation) reads the
            # profiles back plus ~60 B per read of inputs and results.
            bins = float(((rlen[a_lo:a_hi].astype(np.int64) + params.cut_off) // 40 + 3).sum())
            # K4 (hinge calls): per annotation of an annotated read one pass over a 4-byte position column
            # of its pile-up (the selection's first step), 48 B of work item per read; the second step
            # (24 B + two gathers per record near the annotation) touches a few per cent of that
            n_anno = np.diff(result["anno_off"][a_lo:a_hi + 1])
            k4 = float((n_anno * pile).sum()) * 4.0 + 48.0 * float((n_anno > 0).sum())
            # the TMA-staged form of K1 (--profile-kernel 3) reads abpos / aepos only: 8 B per record; K2 also
            # writes and reads back two bits per bin (its two kernels' bit maps)
            rec_b = 8.0 if args.profile_kernel == 3 else 12.0
            kbytes = {"profile": rec_b * novl + 4.0 * bins + 39.0 * owned, "mask_anno": 4.5 * bins + 60.0 * owned,
                      "hinge_call": k4}
            dom = max(kbytes, key=lambda k: kavg[k])
            achieved = kbytes[dom] / (kavg[dom] * 1e-3) / 1e9
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tpath):
                traffic = json.load(open(tpath)).get(pname, {}).get(dom)
            out = {
                "value": total_ovl / (ms_step * 1e-3), "ms_per_step": ms_step, "scaling": p["scaling"],
                "config": {"workload": p["name"] + "; step = hinge filter (coverage profiles + estimate, masks, "
                           "repeat annotation, hinge calls)", "n_reads": n_read, "overlaps": int(total_ovl),
                           "overlaps_rank0": novl, "reads_rank0": owned,
                           "l2": "inputs (%.0f MB of records on rank 0) exceed the 126 MB L2" % (28.0 * novl / 1e6),
                           "parallelism": "reads sharded by A-read id x%d, balanced on record volume; exchange: %s"
                           % (world, "none" if world == 1 else arrays.exchange),
                           "steps": "enqueued back to back (hg_filter_enqueue), one hg_filter_finish" if queued
                           else "one blocking hg_filter per step",
                           "cov_est": int(summary.cov_est), "annotations_rank0": int(summary.n_annotations),
                           "hinges_rank0": int(result["hinge_keep"].sum()),
                           "exact_order_annotations_rank0": int(summary.n_exact_order),
                           "generate_s": round(t_gen, 2)},
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": self.peak, "unit": "GB/s",
                             "frac": achieved / self.peak, "traffic": traffic, "peak_source": self.peak_src,
                             "algorithmic_bytes_per_launch": kbytes[dom], "kernel_ms": kavg[dom],
                             "all_kernels": {k: {"ms": kavg[k], "algorithmic_bytes": kbytes[k],
                                                 "frac": kbytes[k] / (kavg[k] * 1e-3) / 1e9 / self.peak}
                                             for k in kbytes}},
                "kernel_ms": kavg,
                "filter_scan": {"bytes_per_overlap": 32, "gbs": 32.0 * novl / (ms_step * 1e-3) / 1e9,
                                "frac_of_peak": 32.0 * novl / (ms_step * 1e-3) / 1e9 / self.peak},
                "e2e_arrays": {"value": total_ovl / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                               "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * e2e_s, "numa_node_rank0": self.numa},
                "parity": parity, "gpu_launches": int(launches), "clocks": clocks,
            }
            if downstream and world == 1:
                out["downstream_stages"] = time_downstream(syn, ctx, api, result, np, torch)
        ctx.close()
        syn.close()
        del cols_dev
        torch.cuda.empty_cache()
        return out

    def verify(self, pname, p, syn, arrays, summary, result, cols_np, novl):
        """Parity of the timed configuration's results.  N = 1: against the CPU oracle on the same batch.
        N > 1: rank 0 reruns the whole set on ONE context and compares every rank's shard with it (and,
        for the fixed-size workload, that single-context result with the oracle)."""
        np, torch, api, dist = self.np, self.torch, self.api, self.dist
        world, rank = self.world, self.rank
        t0 = time.perf_counter()
        info = {"checked": True, "identical": None, "against": [], "differing": []}
        mine = slice_result(np, result, arrays.lo, arrays.hi)
        mine.update(cov_est=int(summary.cov_est), min_cov=int(summary.min_cov))
        if world > 1:
            # the shards' masks were exchanged during the run: every rank holds all of them (digest travels)
            mine["mask_all_sha256"] = digest(np, arrays.gathered_mask(syn.n_read))
            parts = [None] * world if rank == 0 else None
            dist.gather_object(mine, parts, dst=0)
        else:
            parts = [mine]
        if rank == 0:
            full = None
            if world > 1:
                novl_all = syn.generate(0, syn.n_read, want_trace=False, threads=host_threads())
                cols_all = syn.cols()
                ref = api.Context(self.local, self.stream.cuda_stream)
                ref.set_reads(syn.rlen, syn.qv_off, syn.qv, 100)
                ref.set_overlaps(novl_all, cols_all)
                s1 = ref.filter(api.FilterParams())
                full = ref.filter_fetch(int(s1.n_annotations))
                full["summary"] = np.array([s1.r_begin, s1.r_end, s1.cov_est, s1.min_cov])
                ref.close()
                want_digest = digest(np, full["mask"])
                for r, part in enumerate(parts):
                    bad = compare_slice(np, part, full)
                    if (part["cov_est"], part["min_cov"]) != (int(s1.cov_est), int(s1.min_cov)):
                        bad.append("cov_est/min_cov")
                    if part["mask_all_sha256"] != want_digest:
                        bad.append("exchanged masks")
                    info["differing"] += ["rank%d:%s" % (r, b) for b in bad]
                info["against"].append("single-context GPU run of the whole set (%d overlaps) vs the %d shards"
                                       % (novl_all, world))
            else:
                novl_all, cols_all = novl, cols_np
                full = dict(result)
                full["summary"] = np.array([summary.r_begin, summary.r_end, summary.cov_est, summary.min_cov])
            if world == 1 or p["scaling"] == "strong":
                import oraclelib

                orc = oraclelib.Oracle(syn.rlen, syn.qv_off, syn.qv, 100, cols_all, threads=host_threads())
                want = orc.filter()
                orc.close()
                bad = compare_results(np, full, want)
                if tuple(int(x) for x in full["summary"][2:4]) != tuple(int(x) for x in want["summary"][2:4]):
                    bad.append("cov_est/min_cov")
                info["differing"] += ["oracle:%s" % b for b in bad]
                info["against"].append("CPU oracle (oracle/, %d host threads) on the whole batch (%d overlaps, "
                                       "%d annotations, %d hinges)" % (host_threads(), novl_all, len(want["anno_pos"]),
                                                                       int(want["hinge_keep"].sum())))
            info["identical"] = len(info["differing"]) == 0
            info["seconds"] = round(time.perf_counter() - t0, 1)
        self.barrier()
        return info


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    # the contract is ONE JSON line on stdout: libraries that chat there (NCCL prints its version at
    # the first communicator) are sent to stderr until the line is ready
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    # file-to-file legs first, while this process holds no CUDA context yet (a second context on a busy
    # GPU takes seconds to create and would be charged to the executable)
    f2f = None
    if world == 1 and not args.no_cpu_baseline:
        f2f = file_to_file(preset(args, args.config), 1, 0, with_cli=True)

    b = Bench(args)
    main_res = b.run(args.config, with_clocks=True, downstream=not args.no_downstream)
    also = args.also if args.also is not None else ("c3" if args.config == "c5" else "c5")
    also_res = b.run(also, with_clocks=False, downstream=False) if also else None

    if rank == 0:
        line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
                "scaling": main_res["scaling"], "vs_baseline": None, "dtype": "int32", "data": "synthetic"}
        for k in ("config", "roofline", "kernel_ms", "filter_scan", "parity", "e2e_arrays", "gpu_launches", "clocks",
                  "downstream_stages"):
            if k in main_res:
                line[k] = main_res[k]
        if also_res:
            also_res.pop("clocks", None)
            line[also] = also_res
            line["gpu_launches"] += also_res["gpu_launches"]
        if f2f is not None:
            cli = f2f["cli"]
            line["cpu_baseline"] = {k: f2f[k] for k in ("value", "unit", "cores", "kind", "sample")}
            # the headline end-to-end number: the drop-in executable, file to file, on the very files the
            # reference ran on; h2d / d2h are the bytes the executable moves (7 int32 columns per record in,
            # masks, flags, annotations and the coverage profiles out)
            line["e2e"] = {"value": cli["overlaps_per_s"], "unit": UNIT,
                           "h2d_bytes_per_step": int(28 * f2f["sample_overlaps"]),
                           "d2h_bytes_per_step": int(cli["phases_ms"].get("d2h bytes", 0)),
                           "ms_per_step": 1e3 * cli["seconds"], "path": "hinge filter executable, .db/.las on the "
                           "page cache -> all output files (process start and CUDA context creation included)",
                           "same_files_as_cpu_baseline": True, "overlaps": f2f["sample_overlaps"],
                           "outputs_identical_to_reference": cli["outputs_identical_to_reference"],
                           "phases_ms": cli["phases_ms"], "reference_seconds": cli["reference_seconds"],
                           "seconds_all": cli["seconds_all"], "clocks": cli["clocks"]}
        else:
            line["e2e"] = dict(main_res["e2e_arrays"], path="array-level C ABI, pinned host buffers (file-to-file "
                               "leg skipped: N > 1 or --no-cpu-baseline)")
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    os.dup2(2, 1)
    if world > 1:
        b.dist.destroy_process_group()


if __name__ == "__main__":
    main()
