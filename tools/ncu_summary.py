#!/usr/bin/env python
"""Text summary of an .ncu-rep (the format of profiles/*_ncu_summary.txt): per captured launch the duration, DRAM bytes,
occupancy / issue figures and the warp-stall ratios above 0.4, plus DRAM traffic against the algorithmic bytes.

    python tools/ncu_summary.py prof.ncu-rep <algorithmic bytes per item> <items per launch> ["title line"]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, per_item, items = sys.argv[1], float(sys.argv[2]), float(sys.argv[3])
    if len(sys.argv) > 4:
        print(sys.argv[4])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        print(d["Kernel Name"])
        for k in KEYS:
            if k in d:
                print(f"  {k:<75s} {d[k]:>16s} {unit[k]}")
        stalls = []
        for k, v in d.items():
            if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
                try:
                    stalls.append((float(v), k))
                except ValueError:
                    pass
        for f, k in sorted(stalls, reverse=True):
            if f >= 0.4:
                print(f"  {f:8.3f} {k}")
        traffic = sum(float(d[k]) * SCALE[unit[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        alg = per_item * items
        print(f"  dram traffic per launch: {traffic / 1e9:.4f} GB; algorithmic {per_item:.0f} B x {items:.0f} items = "
              f"{alg / 1e9:.4f} GB; ratio {traffic / alg:.3f}")


if __name__ == "__main__":
    main()
