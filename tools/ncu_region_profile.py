"""Per-code-region breakdown of an `ncu --page source --csv` export: executed warp instructions and stall samples
between marker instructions (regions are given as hex offsets from the kernel's first address)."""
import csv
import sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main(path, bounds):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = []
    for r in rows[2:]:            # first kernel of the report only
        if r and r[0] == 'Kernel Name':
            break
        data.append(r)
    base = int(data[0][idx["Address"]], 16)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    bounds = sorted(bounds) + [1 << 40]
    regs = [dict(lo=bounds[i], hi=bounds[i + 1], inst=0.0, samp=0.0, st={h: 0.0 for h in stalls}) for i in range(len(bounds) - 1)]
    for r in data:
        off = int(r[idx["Address"]], 16) - base
        for g in regs:
            if g["lo"] <= off < g["hi"]:
                g["inst"] += num(r[idx["Instructions Executed"]])
                g["samp"] += num(r[idx["# Samples"]])
                for h in stalls:
                    g["st"][h] += num(r[idx[h]])
    ti = sum(g["inst"] for g in regs) or 1
    ts = sum(g["samp"] for g in regs) or 1
    for g in regs:
        top = sorted(g["st"].items(), key=lambda kv: -kv[1])[:4]
        print(f"[{g['lo']:#7x},{min(g['hi'], 0xfffff):#7x})  inst {g['inst']:10.0f} ({g['inst']/ti*100:5.1f}%)  samples {g['samp']:8.0f} ({g['samp']/ts*100:5.1f}%)  "
              + ", ".join(f"{k[6:]} {v/max(g['samp'],1)*100:.0f}%" for k, v in top))


if __name__ == "__main__":
    main(sys.argv[1], [int(x, 16) for x in sys.argv[2:]])
