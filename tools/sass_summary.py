#!/usr/bin/env python
"""Static evidence for the built library (no GPU needed): per kernel, the SASS mnemonic counts that show which
hardware paths the code uses (DMMA = FP64 tensor pipe, UBLKCP = bulk TMA copy, LDGSTS = async global->shared,
SYNCS = mbarrier, MUFU.EX2 = SFU exponential, which the FP64 paths must NOT depend on), plus registers, shared
memory and spill bytes from the cubin's resource usage.

    python tools/sass_summary.py > profiles/r1_sass_summary.txt
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gpx_b200", "libgpxb200.so")
CUOBJDUMP = os.environ.get("CUOBJDUMP", "/usr/local/cuda/bin/cuobjdump")
CUFILT = os.environ.get("CUFILT", "/usr/local/cuda/bin/cu++filt")
WATCH = ("DMMA", "DFMA", "DADD", "DMUL", "UBLKCP", "LDGSTS", "SYNCS", "LDG", "STG", "LDS", "STS", "SHFL", "BAR",
         "MUFU", "RED", "ATOM", "STL", "LDL")


def demangle(names):
    r = subprocess.run([CUFILT] + list(names), capture_output=True, text=True)
    out = r.stdout.strip().splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def short(name, width=88):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*\)$", "", name)  # drop the argument list, keep template arguments
    return name if len(name) <= width else name[: width - 1] + "~"


def main():
    sass = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True).stdout
    counts, total = collections.defaultdict(collections.Counter), collections.Counter()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(?:\.[A-Za-z0-9_.]+)?\s", ln)
        if m and cur:
            op = m.group(1)
            total[cur] += 1
            if op in WATCH:
                counts[cur][op] += 1
    res = subprocess.run([CUOBJDUMP, "-res-usage", LIB], capture_output=True, text=True).stdout
    usage, fn = {}, None
    for ln in res.splitlines():
        m = re.match(r"\s*Function (\S+):", ln)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in ln:
            d = dict(re.findall(r"(REG|SHARED|STACK|LOCAL):(\d+)", ln))
            usage[fn] = d
            fn = None
    names = sorted(total)
    dm = demangle(names)
    cols = ("DMMA", "DFMA", "UBLKCP", "LDGSTS", "SYNCS", "MUFU", "STL", "LDL")
    print("# SASS summary of gpx_b200/libgpxb200.so (sm_100a), from `cuobjdump -sass` / `-res-usage`")
    print("# DMMA = FP64 tensor-pipe MMA (mma.sync.m8n8k4.f64); UBLKCP = bulk TMA copy (cp.async.bulk);")
    print("# LDGSTS = cp.async; SYNCS = mbarrier ops; STL/LDL = local-memory (spill or indexed array) accesses;")
    print("# STACK = bytes of per-thread stack (0 = nothing spilled).")
    hdr = f"{'kernel':<90}{'instr':>7}" + "".join(f"{c:>8}" for c in cols) + f"{'REG':>5}{'SMEM':>8}{'STACK':>7}"
    print(hdr)
    for n in sorted(names, key=lambda k: -total[k]):
        u = usage.get(n, {})
        print(f"{short(dm[n]):<90}{total[n]:>7}" + "".join(f"{counts[n][c]:>8}" for c in cols)
              + f"{u.get('REG', '?'):>5}{u.get('SHARED', '?'):>8}{u.get('STACK', '?'):>7}")
    agg = collections.Counter()
    for n in names:
        agg.update(counts[n])
    print("\n# totals over %d kernels: " % len(names) + ", ".join(f"{k}={agg[k]}" for k in WATCH if agg[k]))
    spilled = [short(dm[n]) for n in names if int(usage.get(n, {}).get("STACK", 0) or 0) > 0]
    print("# kernels with a non-zero stack frame: %d" % len(spilled))
    for s in spilled:
        print("#   " + s)
    return 0


if __name__ == "__main__":
    sys.exit(main())
