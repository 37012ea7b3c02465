#!/usr/bin/env python
"""Top SASS instructions by stall samples of an .ncu-rep source page (run here, no GPU): python tools/ncu_src.py rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]
ix = {n: i for i, n in enumerate(h)}
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
body = [r for r in rows[hi + 1:] if len(r) == len(h)]
tot = sum(int(r[ix["# Samples"]]) for r in body)
print("total samples", tot, "instructions", len(body))
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:top]
for i in sorted(order):
    r = body[i]
    why = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print("%5d %5.1f%% exec %9s  %-70s %s" % (i, 100.0 * int(r[ix["# Samples"]]) / max(tot, 1), r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70],
                                        " ".join("%s:%d" % (n, c) for c, n in why if c)))
