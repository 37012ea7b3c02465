#!/usr/bin/env python
"""Markdown summary of an ncu launch list (+ optional --set full captures), run here without a GPU.

    python tools/ncu_summary.py --launches gpurun_out/X_launches.csv --title "..." [--plain-ms 3.1] \
        [--rep name=gpurun_out/a.ncu-rep ...] > profiles/rNN_ncu_summary.md
"""
import argparse
import collections
import csv
import subprocess

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    ix = {n: i for i, n in enumerate(h)}
    out = []
    for r in rows[hi + 1:]:
        if len(r) < len(h) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        u = r[ix["Metric Unit"]]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        out.append((r[ix["Kernel Name"]], r[ix["Grid Size"]], r[ix["Block Size"]], v))
    return out


def short(name, n=72):
    name = name.replace("void ", "").replace("egr::", "").replace("unnamed>::", "").replace("<unnamed>::", "")
    return name[:n]


def rep_metrics(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    res = []
    for v in rows[2:]:
        d = {"Kernel Name": v[h.index("Kernel Name")]}
        for k in KEYS:
            if k in h:
                d[k] = (v[h.index(k)], u[h.index(k)])
        res.append(d)
    return res


def stall_summary(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    try:
        hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    except StopIteration:
        return None
    h = rows[hi]
    ix = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    body = [r for r in rows[hi + 1:] if len(r) == len(h) and r[0] != "Address"]      # several kernels: repeated headers
    tot = sum(int(r[ix["# Samples"]]) for r in body) or 1
    agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
    return ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:5] if v)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--launches", required=True)
    ap.add_argument("--title", default="ncu summary")
    ap.add_argument("--plain-ms", type=float, default=None)
    ap.add_argument("--note", default="")
    ap.add_argument("--rep", action="append", default=[])
    ap.add_argument("--traffic-json", default=None,
                    help="also write {stage: dram bytes per launch} of the --rep captures (name = bench.py stage name) to this file; "
                         "bench.py reads profiles/ncu_traffic.json for roofline.traffic")
    a = ap.parse_args()
    if a.traffic_json:
        import json
        import os
        stages = {}
        for spec in a.rep:
            name, path = spec.split("=", 1)
            ms = rep_metrics(path)
            if not ms:
                continue
            d = ms[0]

            def num(k):
                v, u = d.get(k, ("0", ""))
                x = float(v.replace(",", ""))
                return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
            dur, du = d.get("gpu__time_duration.sum", ("0", "us"))
            dur = float(dur.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(du, 1.0)
            stages[name] = {"kernel": short(d["Kernel Name"], 120), "dram_read_bytes": int(num("dram__bytes_read.sum")),
                            "dram_write_bytes": int(num("dram__bytes_write.sum")), "duration_us_under_ncu": dur,
                            "file": os.path.basename(path)}
        json.dump({"source": a.title, "stages": stages}, open(a.traffic_json, "w"), indent=1)
    L = launches(a.launches)
    tot = sum(x[3] for x in L)
    print("# %s\n" % a.title)
    print("Command (after the same command exited 0 without ncu): `ncu --profile-from-start off --metrics gpu__time_duration.sum "
          "--clock-control none --csv python tools/profile_step.py`")
    if a.plain_ms:
        print("Plain run of the same command: step %.3f ms (CUDA events). ncu per-launch times are cold-cache and serialised "
              "(no PDL overlap, no stream overlap): compare SHARES." % a.plain_ms)
    if a.note:
        print("\n" + a.note)
    print("\n## Launch list of one step (%d launches, sum %.1f us)\n" % (len(L), tot))
    print("| # | kernel | grid | block | us | share |\n|---|---|---|---|---|---|")
    for i, (n, g, b, v) in enumerate(L):
        print("| %d | `%s` | %s | %s | %.1f | %.1f%% |" % (i, short(n), g, b, v, 100 * v / tot))
    agg = collections.OrderedDict()
    for n, g, b, v in L:
        k = short(n.split("(")[0], 90)
        c = agg.setdefault(k, [0, 0.0])
        c[0] += 1
        c[1] += v
    print("\n## By kernel\n\n| kernel | launches | us | share |\n|---|---|---|---|")
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, c, v, 100 * v / tot))
    for spec in a.rep:
        name, path = spec.split("=", 1)
        print("\n## `--set full` capture: %s\n\nFile: `%s` (scratch); `ncu --set full --clock-control none --import-source on`\n" % (name, path))
        for d in rep_metrics(path):
            print("Kernel: `%s`\n\n| metric | value |\n|---|---|" % short(d["Kernel Name"], 120))
            for k in KEYS:
                if k in d:
                    print("| %s | %s %s |" % (k, d[k][0], d[k][1]))
        st = stall_summary(path)
        if st:
            print("\nWarp-stall sampling (all samples): %s" % st)


if __name__ == "__main__":
    main()
