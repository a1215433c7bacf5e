"""Per-kernel shares of ONE LML+grad step from an ncu launch list (gpu__time_duration.sum) of bench.py.

    python tools/launch_shares.py profiles/r2_bench_launches_N32768.csv.gz > profiles/r2_bench_launch_shares_N32768.txt

A step is delimited by the launches of sk::mean_kernel (the first kernel of every gpxb_gpr_lml_grad call); the LAST
complete step of the run is taken.  Times under ncu are cold-cache and serialised: compare shares, not absolutes."""
import collections
import csv
import gzip
import re
import sys


def main():
    path = sys.argv[1]
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    hdr = None
    seq = []
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            u = d["Metric Unit"]
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else (v * 1e6 if u == "s" else v))
            name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("gpxb::<", "").replace("void ", "")
            seq.append((name, v))
    starts = [i for i, (n, _) in enumerate(seq) if "mean_kernel" in n]
    if len(starts) < 2:
        raise SystemExit("fewer than two steps in the launch list")
    a, b = starts[-2], starts[-1]
    step = seq[a:b]
    tot = sum(v for _, v in step)
    acc = collections.defaultdict(lambda: [0, 0.0])
    for n, v in step:
        acc[n][0] += 1
        acc[n][1] += v
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none launch list: {path}")
    print(f"# one LML+grad step at N=32768 (launches {a}..{b - 1} of {len(seq)}); times are cold-cache and serialised: compare SHARES")
    print(f"# total {tot / 1000:.1f} ms over {len(step)} launches")
    print(f"{'kernel':70s} {'n':>6s} {'total_us':>12s} {'share':>7s}")
    for n, (c, v) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:70]:70s} {c:6d} {v:12.1f} {100 * v / tot:6.2f}%")


if __name__ == "__main__":
    main()
