"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and share."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ik, ig, iv, iu = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(list)
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[iu], 1e-3)
        name = r[ik]
        for cut in ("(", ):
            if name.startswith("void "):
                name = name[5:]
        name = name.split("(")[0][:70]
        agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    n = sum(len(v) for v in agg.values())
    idrk = sum(sum(v) for k, v in agg.items() if "idrk::" in k)
    print("launches %d, summed kernel time %.1f us (idrk kernels %.1f%%, torch glue %.1f%%)" % (n, tot, 100 * idrk / tot, 100 - 100 * idrk / tot))
    print("%6s %10s %8s %6s  %s" % ("count", "total_us", "avg_us", "share", "kernel"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%6d %10.1f %8.2f %5.1f%%  %s" % (len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot, k))


if __name__ == "__main__":
    main(sys.argv[1])
