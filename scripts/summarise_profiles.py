#!/usr/bin/env python3
"""Turns files of `scripts/gpu_session.sh` sessions (gpurun_out/) into the tracked summaries under
profiles/: bench lines, the per-kernel share of a step from ncu launch lists, the key
`ncu --set full` metrics of each captured kernel and profiles/traffic.json (DRAM bytes per launch and
workload, read by bench.py for `roofline.traffic`).

    python scripts/summarise_profiles.py r02 gpurun_out/r02x_k_profile_flat2_c3.ncu-rep gpurun_out/r02x_launches.csv ...
"""
import csv
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__inst_executed.sum", "smsp__thread_inst_executed.sum",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "launch__waves_per_multiprocessor",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_op_shared_st.sum", "smsp__inst_executed_op_global_ld.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
    "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warp_latency_issue_stalled_wait.ratio",
    "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
    "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
    "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
    "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
    "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
    "smsp__average_warp_latency_issue_stalled_membar.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
]

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def ncu_raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return []
    head, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        out.append({k: (v, u) for k, u, v in zip(head, units, row)})
    return out


def short_name(full):
    m = re.search(r"(k_[A-Za-z0-9_]+)", full)
    return m.group(1) if m else full[:40]


def launch_table(path, title):
    """Per-kernel device time of ONE steady-state filter step (from one K1 launch to the next) out of an
    `ncu --metrics gpu__time_duration.sum` launch list."""
    rows = list(csv.reader(open(path)))
    hdr = next((r for r in rows if r and r[0] == "ID"), None)
    if hdr is None:
        return None
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    body = [r for r in rows if len(r) > iv and r[0].isdigit()]
    names = [short_name(r[ik]) for r in body]
    ns = [float(r[iv].replace(",", "")) for r in body]
    starts = [i for i, n in enumerate(names) if n.startswith("k_profile_flat")]
    lines = ["# %s" % title, "",
             "`ncu --metrics gpu__time_duration.sum --clock-control none` serialises launches and runs them with a",
             "cold cache, so only the SHARES are comparable with the CUDA-event times in the bench line (the size",
             "tiers of the exact-order hinge kernels, for one, overlap on forked streams in a real run).", ""]
    if len(starts) >= 3:
        a, b = starts[1], starts[2]
        agg = {}
        for n, t in zip(names[a:b], ns[a:b]):
            agg.setdefault(n, [0, 0.0])
            agg[n][0] += 1
            agg[n][1] += t
        tot = sum(v[1] for v in agg.values())
        lines += ["| kernel | launches | time (us) | share |", "|---|---|---|---|"]
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            lines.append("| %s | %d | %.1f | %.1f %% |" % (n, c, t / 1e3, 100 * t / tot))
        lines.append("| **total** | %d | %.1f | |" % (b - a, tot / 1e3))
    lines += ["", "all %d launches of the capture: the .csv next to this file" % len(body)]
    return "\n".join(lines) + "\n"


# kernel -> the bench line's kernel group (traffic.json holds DRAM bytes per launch and workload; a group
# of several kernels gets the sum of its captured members)
ALIAS = {"k_mask_bits_flat": "mask_anno", "k_mask_walk": "mask_anno", "k_profile_flat": "profile",
         "k_profile_flat2": "profile", "k_hinge_call": "hinge_call"}


def main():
    """summarise_profiles.py <name under profiles/> <file> ...
    files: gpurun_out/*.ncu-rep (a `_c3` / `_c5` suffix names the workload), *_launches*.csv, anything else
    is copied as is."""
    dst = sys.argv[1]
    os.makedirs(PROF, exist_ok=True)
    traffic_path = os.path.join(PROF, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    traffic = {k: v for k, v in traffic.items() if isinstance(v, dict)}  # round-1 layout: flat keys
    for f in sys.argv[2:]:
        base = os.path.basename(f)
        tail = base.split("_", 1)[1] if "_" in base else base
        if base.endswith(".ncu-rep"):
            cfg = "c5" if "_c5" in base else ("c3" if "_c3" in base else "")
            for d in ncu_raw(f):
                kname = short_name(d["Kernel Name"][0])
                lines = ["# ncu --set full: %s (%s%s)" % (d["Kernel Name"][0], dst, ", workload " + cfg if cfg else ""), ""]
                for k in KEYS:
                    if k in d and d[k][0] != "":
                        lines.append("%-80s %s %s" % (k, d[k][0], d[k][1]))
                name = "%s_%s%s_ncu.txt" % (dst, kname, "_" + cfg if cfg else "")
                open(os.path.join(PROF, name), "w").write("\n".join(lines) + "\n")
                try:
                    rd = float(d["dram__bytes_read.sum"][0].replace(",", "")) * SCALE[d["dram__bytes_read.sum"][1]]
                    wr = float(d["dram__bytes_write.sum"][0].replace(",", "")) * SCALE[d["dram__bytes_write.sum"][1]]
                    if kname in ALIAS and cfg:
                        t = traffic.setdefault(cfg, {})
                        parts = t.setdefault(ALIAS[kname] + "_kernels", {})
                        if not isinstance(parts, dict):
                            parts = t[ALIAS[kname] + "_kernels"] = {}
                        parts[kname] = rd + wr
                        t[ALIAS[kname]] = sum(parts.values())
                        t[ALIAS[kname] + "_source"] = "%s_*_%s_ncu.txt" % (dst, cfg)
                except (KeyError, ValueError):
                    pass
        elif "launches" in base and base.endswith(".csv"):
            shutil.copy(f, os.path.join(PROF, "%s_%s" % (dst, tail)))
            md = launch_table(f, "ncu launch list (%s, %s): per-kernel device time of one steady-state filter step" % (dst, tail))
            if md:
                open(os.path.join(PROF, "%s_%s" % (dst, tail.replace(".csv", ".md"))), "w").write(md)
        else:
            shutil.copy(f, os.path.join(PROF, "%s_%s" % (dst, tail)))
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    print("profiles/ updated:", dst)


if __name__ == "__main__":
    main()
