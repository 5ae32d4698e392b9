"""Top warp-stall sampling locations from `ncu -i rep --page source --csv --kernel-name ...` output."""
import csv
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main(path, n=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot = sum(num(r[idx['# Samples']]) for r in data)
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {h: sum(num(r[idx[h]]) for r in data) for h in stalls}
    print("total samples", tot)
    print("stall mix:", {k: f"{v/tot*100:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for r in sorted(data, key=lambda r: -num(r[idx['# Samples']]))[:n]:
        s = num(r[idx['# Samples']])
        st = sorted([(num(r[idx[h]]), h) for h in stalls], reverse=True)[:2]
        print(f"{s:7.0f} {s/tot*100:5.1f}%  {r[idx['Source']][:100]:100s} {[(int(a), b) for a, b in st]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
