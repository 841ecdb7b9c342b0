"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, time, share of the run."""
import collections
import csv
import re
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = row["Metric Unit"]
        t *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(unit, 1e-6)
        m = re.search(r"gemm_f64_dmma<(\d+), (\d+), (\d+), (\d+), (\d+), (\d+), \(bool\)(\d), \(bool\)(\d), (\d+)>", name)
        if m:
            short = "gemm_f64_dmma<BM%s,BN%s,BK%s,akmajor%s,bkmajor%s,vec%s>" % (m.group(1), m.group(2), m.group(3), m.group(7), m.group(8), m.group(9))
        else:
            m3 = re.search(r"gemm_f64_tma<\(bool\)(\d), \(bool\)(\d)>", name)
            if m3:
                short = "gemm_f64_tma<akmajor%s,bkmajor%s>" % (m3.group(1), m3.group(2))
                agg[short][0] += 1
                agg[short][1] += t
                continue
            m2 = re.search(r"(k_\w+|permute_\w+|splitk_reduce|scale_c|fused_\w+|triples_\w+)", name)
            short = m2.group(1) if m2 else name[:70]
        agg[short][0] += 1
        agg[short][1] += t
    tot = sum(v[1] for v in agg.values())
    lines_out = ["# %s: %d launches, %.2f ms summed device time (cold-cache, serialised under ncu: compare SHARES)" % (
        path, sum(v[0] for v in agg.values()), tot), "%-62s %7s %12s %7s" % ("kernel", "count", "ms", "share")]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines_out.append("%-62s %7d %12.3f %6.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    txt = "\n".join(lines_out) + "\n"
    if out:
        open(out, "w").write(txt)
    print(txt)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
