"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: agg_launches.py file.csv [top]"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("p5::", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg[n][0] += 1
        agg[n][1] += v
    tot = sum(v[1] for v in agg.values())
    print("total_us %.1f launches %d" % (tot, sum(v[0] for v in agg.values())))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print("%10.1f us %5.1f%% %6d x %8.2f us  %s" % (t, 100 * t / tot, c, t / c, n[:80]))


if __name__ == "__main__":
    main()
