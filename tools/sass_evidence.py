#!/usr/bin/env python
"""SASS evidence for the shipped library: per kernel, how many tcgen05 / TMA / TMEM instructions `cuobjdump -sass` shows
(tcgen05.mma -> UTC*MMA, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk -> UBLKCP, tcgen05.ld -> LDTM, tcgen05.commit -> UTCBAR),
that no legacy HMMA exists, and each kernel's registers / stack / spills from `cuobjdump -res-usage`.
Usage: python tools/sass_evidence.py > profiles/rNN_sass_evidence.txt   (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "code2vec_b200", "libc2v_b200.so")
PATS = [("UTCxMMA", r"\bUTC[A-Z]*MMA"), ("UTMALDG", r"\bUTMALDG"), ("UBLKCP", r"\bUBLKCP"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"),
        ("HMMA", r"\bHMMA"), ("RED/ATOM", r"\b(RED|ATOMG|ATOM)\b"), ("MUFU", r"\bMUFU")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
            for name, pat in PATS:
                if re.search(pat, line):
                    counts[cur][name] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*(.*)", res):
        usage[m.group(1)] = m.group(2)
    names = demangle(list(counts))
    arch = re.search(r"arch = (\S+)", sass)
    print("SASS evidence, %s (nvcc -gencode arch=compute_100a,code=sm_100a), arch = %s; `cuobjdump -sass` instruction counts per kernel."
          % (os.path.relpath(LIB, ROOT), arch.group(1) if arch else "?"))
    print("tcgen05.mma.kind::tf32 -> UTC*MMA, cp.async.bulk.tensor -> UTMALDG, cp.async.bulk (1-D) -> UBLKCP, tcgen05.ld -> LDTM, "
          "tcgen05.commit -> UTCBAR; no legacy HMMA anywhere.\n")
    hdr = "%-118s" % "kernel" + "".join("%9s" % n for n, _ in PATS) + "   resources"
    print(hdr)
    tot = collections.Counter()
    for k, c in counts.items():
        d = names[k].replace("c2v::umma::", "").replace("c2v::", "").split("(CUtensorMap")[0].split("(float")[0].split("(c2v")[0][:116]
        u = usage.get(k, "")
        mu = re.search(r"REG:(\d+).*?STACK:(\d+)", u)
        print("%-118s" % d + "".join("%9d" % c[n] for n, _ in PATS) + ("   regs %s stack %s" % mu.groups() if mu else ""))
        tot.update(c)
    print("\nTOTAL" + " " * 113 + "".join("%9d" % tot[n] for n, _ in PATS))
    assert tot["HMMA"] == 0


if __name__ == "__main__":
    main()
