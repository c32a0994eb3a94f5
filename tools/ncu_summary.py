#!/usr/bin/env python
"""Write a text summary of an .ncu-rep (key raw metrics + stall breakdown) for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt"""
import csv
import io
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_stalls  # noqa: E402

WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_elapsed.max", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on : {os.path.basename(rep)}")
    for n, r in enumerate(rows[2:]):
        print(f"\n## launch {n}: {r[hdr.index('Kernel Name')][:150]}")
        rd = wr = None
        for w in WANT:
            if w in hdr:
                v, u = r[hdr.index(w)], units[hdr.index(w)]
                print(f"  {w:70s} {v} {u}")
                if w == "dram__bytes_read.sum":
                    rd = (float(v), u)
                if w == "dram__bytes_write.sum":
                    wr = (float(v), u)
        if rd and wr:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
            tot = rd[0] * scale.get(rd[1], 1) + wr[0] * scale.get(wr[1], 1)
            print(f"  {'traffic = dram read + write per launch':70s} {tot:.6g} byte")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    tmp = "/tmp/_ncu_src.csv"
    open(tmp, "w").write(src)
    print("\n## warp-stall sampling, launch 0")
    ncu_stalls.main(tmp, 0)


if __name__ == "__main__":
    main(sys.argv[1])
