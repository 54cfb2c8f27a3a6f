#!/usr/bin/env python
"""Per-kernel bound analysis from one `ncu --set full` capture: duration, DRAM / L2 / L1 throughput (% of peak), L2 hit rate,
tensor-pipe and issue-slot activity, resident warps, and the warp-stall reasons with the largest share -- one row per
distinct kernel (first launch), longest first.  usage: ncu_top_kernels.py capture.ncu-rep [min_us]"""
import csv
import subprocess
import sys


def main(rep, min_us=20.0):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]

    def num(r, k):
        try:
            return float(r[idx[k]].replace(",", ""))
        except (KeyError, ValueError):
            return float("nan")
    seen, out = set(), []
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if name in seen:
            continue
        seen.add(name)
        us = num(r, "gpu__time_duration.sum")
        if us < min_us:
            continue
        top = sorted(((num(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall), reverse=True)[:3]
        out.append((us, name, r, top))
    out.sort(reverse=True)
    print("| kernel | us | DRAM % | L2 % | L2 hit % | L1 % | tensor % | issue % | warps % | top stalls (warps per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for us, name, r, top in out:
        short = name.replace("void ", "").replace("umma::", "")
        short = short[:short.index("(")] if "(" in short else short
        print("| `%s` | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %.0f | %s |" % (
            short[:64], us, num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            num(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"), num(r, "lts__t_sector_hit_rate.pct"),
            num(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
            num(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            ", ".join("%s %.1f" % (n, v) for v, n in top)))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 20.0)
