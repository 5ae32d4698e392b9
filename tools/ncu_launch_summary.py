"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.defaultdict(list)
    for row in r:
        try:
            v = float(row[vi].replace(',', ''))
        except ValueError:
            continue
        v = v / 1000 if row[ui] == 'ns' else (v * 1000 if row[ui] == 'ms' else v)
        agg[re.sub(r'\(.*', '', row[ki])[:70]].append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':70s} {'n':>5s} {'mean_us':>9s} {'total_ms':>9s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:70s} {len(v):5d} {sum(v)/len(v):9.1f} {sum(v)/1000:9.2f} {sum(v)/tot*100:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
