#!/usr/bin/env python3
"""Turns one `scripts/profile_gpu.sh <tag>` session (gpurun_out/<tag>_*) into the tracked
summaries under profiles/: bench lines, the per-kernel share of a step from the ncu launch
list, the key `ncu --set full` metrics of each captured kernel and profiles/traffic.json
(DRAM bytes per launch, read by bench.py for `roofline.traffic`).

    python scripts/summarise_profiles.py r01a [r01]      (source tag, name under profiles/)
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


def main():
    src = sys.argv[1]
    dst = sys.argv[2] if len(sys.argv) > 2 else src
    os.makedirs(PROF, exist_ok=True)
    for suffix in ("bench.json", "bench_reference.json", "pytest_gpu.log", "smoke.log", "gpu.txt", "launches.csv"):
        p = os.path.join(OUT, "%s_%s" % (src, suffix))
        if os.path.exists(p):
            shutil.copy(p, os.path.join(PROF, "%s_%s" % (dst, suffix)))

    # ---- launch list -> per-kernel share of the steady-state step
    lp = os.path.join(OUT, "%s_launches.csv" % src)
    if os.path.exists(lp):
        rows = [r for r in csv.reader(open(lp)) if len(r) > 14 and r[0].isdigit()]
        names = [short_name(r[4]) if "hg::" in r[4] else "torch:" + r[4][:48] for r in rows]
        ns = [float(r[14].replace(",", "")) for r in rows]
        # one step = from a k_profile_flat launch to the next; the second one is a warm, device-resident step
        # (later ones belong to the end-to-end arm, which re-ingests: k_csr_validate, k_max_pileup)
        starts = [i for i, n in enumerate(names) if n == "k_profile_flat"]
        lines = ["# ncu launch list (%s): per-kernel device time of ONE steady-state filter step" % dst, "",
                 "`ncu --metrics gpu__time_duration.sum --clock-control none` serialises launches and runs them with a",
                 "cold cache, so only the SHARES are comparable with the CUDA-event times in the bench line.", ""]
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
        lines += ["", "all %d launches seen: see %s_launches.csv" % (len(rows), dst)]
        open(os.path.join(PROF, "%s_launches.md" % dst), "w").write("\n".join(lines) + "\n")

    # ---- ncu --set full captures
    traffic_path = os.path.join(PROF, "traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    alias = {"k_mask_anno_flat": "mask_anno", "k_profile_flat": "profile", "k_hinge_call": "hinge_call",
             "k_classify_pairs": "classify_pairs"}
    for f in sorted(os.listdir(OUT)):
        if not (f.startswith(src + "_k_") and f.endswith(".ncu-rep")):
            continue
        for d in ncu_raw(os.path.join(OUT, f)):
            kname = short_name(d["Kernel Name"][0])
            lines = ["# ncu --set full: %s (%s)" % (d["Kernel Name"][0], dst), ""]
            for k in KEYS:
                if k in d and d[k][0] != "":
                    lines.append("%-80s %s %s" % (k, d[k][0], d[k][1]))
            open(os.path.join(PROF, "%s_%s_ncu.txt" % (dst, kname)), "w").write("\n".join(lines) + "\n")
            try:
                rd = float(d["dram__bytes_read.sum"][0].replace(",", "")) * SCALE[d["dram__bytes_read.sum"][1]]
                wr = float(d["dram__bytes_write.sum"][0].replace(",", "")) * SCALE[d["dram__bytes_write.sum"][1]]
                if kname in alias:
                    traffic[alias[kname]] = rd + wr
                    traffic[alias[kname] + "_source"] = "%s_%s_ncu.txt" % (dst, kname)
            except (KeyError, ValueError):
                pass
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    print("profiles/ updated from", src)


if __name__ == "__main__":
    main()
