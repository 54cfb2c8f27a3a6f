#!/usr/bin/env python
"""Per-launch table from an `ncu --set full` report: duration, DRAM bytes read / written, DRAM GB/s,
SM throughput %, registers.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_ncu_summary.txt
(`ncu -i <rep> --page raw --csv --print-units base` is the only thing it runs.)"""
import csv
import io
import subprocess
import sys


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    col = lambda r, name: r[ix[name]]
    print("%-64s %14s %4s %9s %10s %10s %8s %6s" % ("kernel", "grid", "regs", "us", "dram_rd_MB", "dram_wr_MB", "DRAM GB/s", "sm%"))
    for r in rows[2:]:
        name = col(r, "Kernel Name").split("(")[0][:64]
        us = float(col(r, "gpu__time_duration.sum")) / 1e3
        rd, wr = float(col(r, "dram__bytes_read.sum")), float(col(r, "dram__bytes_write.sum"))
        print("%-64s %14s %4s %9.1f %10.1f %10.1f %8.0f %6.1f" % (
            name, col(r, "Grid Size").replace(" ", ""), col(r, "launch__registers_per_thread"), us, rd / 1e6, wr / 1e6,
            (rd + wr) / us / 1e3, float(col(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"))))


if __name__ == "__main__":
    main(sys.argv[1])
