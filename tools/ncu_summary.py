#!/usr/bin/env python
"""Summarise an ncu report (`ncu -i X.ncu-rep --page raw --csv`) into a small tracked text table:
per captured launch the duration, DRAM bytes, DRAM / SM throughput, registers, occupancy, instruction
count, and the top stall reasons."""
import csv
import subprocess
import sys

KEYS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("launch__registers_per_thread", "regs"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
        ("smsp__inst_executed.sum", "warp_inst"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
STALLS = ["barrier", "long_scoreboard", "short_scoreboard", "wait", "math_pipe_throttle", "not_selected", "mio_throttle",
          "lg_throttle", "branch_resolving", "dispatch_stall", "membar", "sleeping", "no_instruction", "imc_miss", "drain"]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print("kernel:", d[idx["Kernel Name"]][:110])
        for key, name in KEYS:
            if key in idx:
                print(f"  {name:10s} {d[idx[key]]:>18s} {units[idx[key]]}")
        st = []
        for s in STALLS:
            k = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
            if k in idx:
                try:
                    st.append((float(d[idx[k]]), s))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  stalls/issue: " + ", ".join(f"{n}={v:.2f}" for v, n in st[:6]))
        print()


if __name__ == "__main__":
    main(sys.argv[1])
