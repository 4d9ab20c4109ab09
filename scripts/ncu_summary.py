#!/usr/bin/env python
"""Summarise .ncu-rep captures into profiles/<round>/: a markdown table of the headline counters
per kernel and ncu_traffic.json (DRAM bytes per problem per launch of the PDIPM iteration kernel,
consumed by bench.py's roofline.traffic).

    python scripts/ncu_summary.py r01 gpurun_out/prof_iter.ncu-rep [gpurun_out/prof_mpc.ncu-rep ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor pipe % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier (warps/issue)"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    rnd, reps = sys.argv[1], sys.argv[2:]
    outdir = os.path.join(ROOT, "profiles", rnd)
    os.makedirs(outdir, exist_ok=True)
    md = [f"# ncu summaries ({rnd})", "",
          "`ncu --set full --clock-control none --import-source on`, one launch per kernel; numbers under the profiler "
          "are diagnostic only (bench values come from `bench.py` without ncu).", ""]
    traffic = {}
    for path in reps:
        hdr, units, rows = raw(path)
        for r in rows:
            name = r[hdr.index("Kernel Name")]
            md += [f"## `{name}`  ({os.path.basename(path)})", "", "| counter | value | unit |", "|---|---|---|"]
            vals = {}
            for key, label in KEYS:
                if key in hdr:
                    i = hdr.index(key)
                    vals[key] = (r[i], units[i])
                    md.append(f"| {label} (`{key}`) | {r[i]} | {units[i]} |")
            md.append("")
            if "k_fast_iter" in name or "k_pdipm_iter" in name:
                big = "k_pdipm_iter" in name
                def to_bytes(k):
                    v, u = vals[k]
                    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                    return float(v) * mult
                grid = float(vals["launch__grid_size"][0])
                rec = {"kernel": name, "problems_in_capture": grid,
                           "dram_bytes_per_problem_per_launch": (to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")) / grid,
                           "source": os.path.basename(path)}
                if big:
                    json.dump(rec, open(os.path.join(outdir, "ncu_traffic_large_qp.json"), "w"), indent=1)
                else:
                    traffic = rec
    open(os.path.join(outdir, f"ncu_summary_{rnd}.md"), "w").write("\n".join(md) + "\n")
    if traffic:
        json.dump(traffic, open(os.path.join(outdir, "ncu_traffic.json"), "w"), indent=1)
    print("wrote", outdir, traffic)


if __name__ == "__main__":
    main()
