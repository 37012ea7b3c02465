#!/usr/bin/env python
"""SASS evidence for the tensor-core / TMA path, generated here without a GPU:

    python tools/sass_excerpt.py > profiles/sass_tc_kernels.txt

Per kernel of egorear_b200/libegorear_b200.so: counts of the Blackwell mnemonics that prove tcgen05 / TMEM / TMA
(B200_PROFILING.md: UTCHMMA = tcgen05.mma kind::f16, UTCQMMA / UTC*MMA other kinds, LDTM = tcgen05.ld, UTMALDG / UTMASTG =
cp.async.bulk.tensor load / store, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops), then the first lines of each kind
inside the main GEMM kernel as a literal excerpt.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "egorear_b200", "libegorear_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "SYNCS", "IDP.4A", "HFMA2", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    rows, excerpts = [], {}
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        try:
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        except FileNotFoundError:
            dem = name
        cnt = collections.Counter()
        for line in f.splitlines():
            m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]+)", line)
            if not m:
                continue
            op = m.group(1)
            for p in PAT:
                if op == p or op.startswith(p + "."):
                    cnt[p] += 1
                    if "gemm_tc_kernel<(int)128, __nv_bfloat16, __nv_bfloat16>" in dem:
                        excerpts.setdefault(p, [])
                        if len(excerpts[p]) < 3:
                            excerpts[p].append(line.strip())
        if any(cnt[p] for p in PAT[:11]):
            rows.append((dem, cnt))
    print("SASS of egorear_b200/libegorear_b200.so (cuobjdump -sass, sm_100a), kernels that use tcgen05 / TMEM / TMA\n")
    print("%-96s %s" % ("kernel", " ".join("%8s" % p for p in PAT)))
    tot = collections.Counter()
    for dem, cnt in sorted(rows):
        short = dem.replace("egr::", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
        short = re.sub(r">\(.*$", ">", short)
        print("%-96s %s" % (short[:96], " ".join("%8d" % cnt[p] for p in PAT)))
        tot.update(cnt)
    print("%-96s %s" % ("TOTAL (%d kernels)" % len(rows), " ".join("%8d" % tot[p] for p in PAT)))
    print("\nexcerpt: gemm_tc_kernel<128, bf16, bf16>")
    for p, lines in excerpts.items():
        for l in lines:
            print("  " + l)
    return 0


if __name__ == "__main__":
    sys.exit(main())
