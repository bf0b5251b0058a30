#!/usr/bin/env python3
"""Per-opcode and per-instruction hot spots of an .ncu-rep (source page, SASS view).
    python tools/ncu_sass_hot.py file.ncu-rep [top_n]"""
import collections
import csv
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r][0]
    h = rows[hi]
    ix = {k: h.index(k) for k in ("Source", "# Samples", "Instructions Executed", "stall_long_sb", "stall_short_sb",
                                   "stall_barrier", "stall_wait", "stall_math", "stall_mio", "stall_tex")}
    ops = collections.defaultdict(lambda: [0, 0])
    ins = []
    tot_i = tot_s = 0
    for r in rows[hi + 1:]:
        if len(r) <= ix["Instructions Executed"]:
            continue
        try:
            n = int(r[ix["Instructions Executed"]] or 0)
            s = int(r[ix["# Samples"]] or 0)
        except ValueError:
            continue
        src = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]].strip())
        op = src.split()[0].split(".")[0] if src else "?"
        ops[op][0] += n
        ops[op][1] += s
        tot_i += n
        tot_s += s
        ins.append((s, n, r[ix["Source"]].strip()[:90], {k: r[ix[k]] for k in ("stall_long_sb", "stall_short_sb", "stall_barrier", "stall_wait", "stall_math", "stall_mio", "stall_tex")}))
    print("total warp-instructions %d, samples %d" % (tot_i, tot_s))
    print("%-12s %12s %6s %10s %6s" % ("opcode", "executed", "%", "samples", "%"))
    for op, (n, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-12s %12d %5.1f%% %10d %5.1f%%" % (op, n, 100.0 * n / max(tot_i, 1), s, 100.0 * s / max(tot_s, 1)))
    print("\nhottest instructions by stall samples:")
    for s, n, src, st in sorted(ins, key=lambda t: -t[0])[:top]:
        big = {k[6:]: v for k, v in st.items() if v not in ("", "0")}
        print("%7d %9d  %-90s %s" % (s, n, src, big))


if __name__ == "__main__":
    main()
