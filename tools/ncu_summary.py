#!/usr/bin/env python3
"""Summarise an ncu report (`ncu --set full`) as JSON: one entry per profiled launch with duration, instructions, DRAM bytes,
occupancy limits, hit rates, pipe utilisation and the stall shares of the warp-state samples.

usage: tools/ncu_summary.py <report.ncu-rep> [note] > profiles/<name>.json
"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = {"_note": sys.argv[2] if len(sys.argv) > 2 else "", "launches": []}
    for r in rows[2:]:
        e = {"kernel": r[idx["Kernel Name"]]}
        for k in KEEP:
            if k in idx:
                e[k] = {"value": r[idx[k]], "unit": units[idx[k]]}
        st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[i] or 0) for h, i in idx.items()
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
        tot = sum(st.values()) or 1.0
        e["stall_share_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(st.items(), key=lambda kv: -kv[1]) if v / tot >= 0.02}
        out["launches"].append(e)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
