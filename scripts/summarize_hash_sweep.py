"""Table of a scripts/hash_microbench.py JSON-lines sweep (profiles/r02_hash_encode_sweep.txt)."""
import json
import sys


def main(path):
    rows = []
    for line in open(path):
        try:
            rows.append(json.loads(line))
        except ValueError:
            pass
    print("# python scripts/hash_microbench.py --n 1048576 16777216 67108864 268435456 --log2T 14 16 18 19 20 22 24   (1 x B200, L=16, F=2, C=16,")
    print("# median of 10 after 3 warm-ups, 256 MB L2 flush between runs; 268435456 points run as 4 launches over 2^26-point chunks of ONE 3 GB input)")
    print("# frac = algorithmic bytes/point (fwd 1304 | 408, table-grad bwd 2188 | 396, bwd+dx 3364 | 548 for trilinear | reference) / time / measured HBM peak %.1f GB/s" % rows[0]["hbm_peak_gbs_all_gpus"])
    print("# plain = points walked as given (uniform random); sort = the library's Morton radix sort of the same batch (sort_ms); +sort = Z-order walk")
    print("# through the permutation WITH the sort inside the time; presorted = batch already in Z-order; pair = forward + table-grad backward (one sort)")
    print("%-9s %5s %10s | %-21s %-21s %-21s | %7s | %-21s %-21s | %-21s %-21s | %-15s" % (
        "mode", "log2T", "n", "fwd plain Gpts/s frac", "bwd plain", "bwd+dx plain", "sort ms", "fwd +sort", "bwd +sort", "fwd presorted",
        "bwd presorted", "pair +sort/plain"))

    def cell(r, key, ms=None):
        if key + "_mpts" not in r:
            return "%-21s" % "-"
        return "%7.2f  %5.2f        " % (r[key + "_mpts"] / 1e3, r[key + "_frac"])
    for r in rows:
        extra = "%5.2f / %5.2f" % (r["pair_one_sort_frac"], r["pair_plain_frac"]) if "pair_one_sort_frac" in r else "-"
        print("%-9s %5d %10d | %s %s %s | %7s | %s %s | %s %s | %-15s" % (
            r["mode"], r["log2T"], r["n"], cell(r, "fwd"), cell(r, "bwd"), cell(r, "bwd_with_dx"),
            ("%.2f" % r["sort_ms"]) if "sort_ms" in r else "-", cell(r, "fwd_sort_included"), cell(r, "bwd_sort_included"),
            cell(r, "fwd_presorted"), cell(r, "bwd_presorted"), extra))


if __name__ == "__main__":
    main(sys.argv[1])
