#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small JSON: one record per profiled launch
with the metrics the roofline discussion in DESIGN.md uses.  Usage:
    python tools/ncu_summary.py gpurun_out/prof_full.ncu-rep profiles/r01_ncu_full_summary.json
"""
import csv
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum",
    "sm__cycles_elapsed.max",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed.sum",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_xu.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__cluster_size",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_membar_per_warp_active.pct",
    "smsp__warp_issue_stalled_sleeping_per_warp_active.pct",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        rec = {"kernel": r[hdr.index("Kernel Name")].split("(")[0], "id": r[hdr.index("ID")]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                v = r[i].replace(",", "")
                try:
                    rec[k] = float(v)
                except ValueError:
                    rec[k] = v
                rec.setdefault("_units", {})[k] = units[i]
        recs.append(rec)
    json.dump({"source": rep, "command": "ncu --set full --clock-control none --import-source on ... "
               "python bench.py --steps 1 --warmup 1 --no-cpu-baseline", "launches": recs},
              open(out, "w"), indent=1)
    for rec in recs:
        print(rec["kernel"], rec.get("gpu__time_duration.sum"), rec["_units"].get("gpu__time_duration.sum"),
              "tensor%", rec.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
              "dramR", rec.get("dram__bytes_read.sum"), rec["_units"].get("dram__bytes_read.sum"),
              "dramW", rec.get("dram__bytes_write.sum"), "L2hit", rec.get("lts__t_sector_hit_rate.pct"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
