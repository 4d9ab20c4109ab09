#!/usr/bin/env python
"""Round-2 ncu summary: profiles/r02/ncu_summary_r02.md (headline counters of every captured launch, incl. the FP64 pipe and
the DMMA sub-pipe of the tensor pipe), profiles/r02/ncu_traffic.json (DRAM bytes per problem of the first k_wres_chunk launch
and of the whole solve; consumed by bench.py's roofline.traffic).

    python scripts/ncu_summary2.py gpurun_out/r2fin/prof_wres.ncu-rep [gpurun_out/r2fin/prof_rex.ncu-rep]
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs/thread"), ("launch__occupancy_limit_registers", "occupancy limit (regs), CTAs/SM"),
        ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe (DFMA + DMMA) cycles active %"),
        ("sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe, DMMA sub-pipe cycles active %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe cycles active %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_sector_hit_rate.pct", "L2 hit rate %"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (warps/issue)"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait"),
        ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle")]
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw(path):
    """a .ncu-rep report, or the `--page raw --csv` dump of one (the GPU box only returns the dumps: 64 MiB cap)"""
    if path.endswith(".csv"):
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    outdir = os.path.join(ROOT, "profiles", "r02")
    os.makedirs(outdir, exist_ok=True)
    md = ["# ncu summaries (r02)", "",
          "`ncu --set full --clock-control none --import-source on` (scripts/gpu/r2_prof.sh).  QP: the five launches of the second "
          "forward+backward call at the bench batch (`scripts/run_qp_once.py 32768 2`: pre-factorisation, chunk 1 = initial point + "
          "iterations 0-9, chunk 2 = iterations 10-19, the repair launch, backward).  MPC: the rex-quadrotor AL solve "
          "(`scripts/run_rex_once.py`, B = 1024, T = 40).  Numbers under the profiler are diagnostic only (bench values come from "
          "`bench.py` without ncu).", ""]
    per, tot = [], 0.0
    for path in sys.argv[1:]:
        hdr, units, rows = raw(path)
        for r in rows:
            name = r[hdr.index("Kernel Name")]
            md += [f"## `{name}`  ({os.path.basename(path)})", "", "| counter | value | unit |", "|---|---|---|"]
            for k, l in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    md.append(f"| {l} (`{k}`) | {r[i]} | {units[i]} |")
            md.append("")
            if "k_wres" in name:
                i, j = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
                b = float(r[i]) * MULT[units[i]] + float(r[j]) * MULT[units[j]]
                grid = float(r[hdr.index("launch__grid_size")])
                per.append((name, b / grid))
                tot += b / grid
    if per:
        md += ["## DRAM traffic per QP solve", "", "| launch | DRAM bytes per problem |", "|---|---|"]
        md += [f"| `{n[:60]}` | {b:,.0f} |" for n, b in per]
        md += [f"| **whole solve ({len(per)} launches)** | **{tot:,.0f}** = {tot / 68160:.2f} x the 68,160 algorithmic bytes (r01: ~1.5 MB = 22 x) |", ""]
        chunk = [b for n, b in per if "k_wres_chunk" in n]
        json.dump({"kernel": "k_wres_chunk<8,30,60,1> (first launch of a call: initial point + 10 iterations)", "problems_in_capture": 32768,
                   "dram_bytes_per_problem_per_launch": chunk[0] if chunk else None, "dram_bytes_per_solve_all_launches": tot,
                   "source": " ".join(os.path.basename(p) for p in sys.argv[1:]) + " (scripts/gpu/r2_prof.sh, scripts/ncu_summary2.py)"},
                  open(os.path.join(outdir, "ncu_traffic.json"), "w"), indent=1)
    open(os.path.join(outdir, "ncu_summary_r02.md"), "w").write("\n".join(md) + "\n")
    print("wrote", outdir, per, tot)


if __name__ == "__main__":
    main()
