"""Text summary of an ncu --set full report: key raw metrics per kernel + top stall locations (needs ncu on PATH)."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("=" * 100)
        print(r[idx["Kernel Name"]][:120])
        for k in KEYS:
            if k in idx:
                print(f"  {k:75s} {r[idx[k]]:>18s} {units[idx[k]]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for si, st in enumerate(starts[:1]):
        en = starts[si + 1] if si + 1 < len(starts) else len(rows)
        hdr = rows[st + 1]
        idx = {h: i for i, h in enumerate(hdr)}
        data = rows[st + 2:en]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return 0.0
        tot = sum(num(r[idx["# Samples"]]) for r in data) or 1.0
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = sorted(((sum(num(r[idx[h]]) for r in data), h) for h in stalls), reverse=True)[:8]
        print("-" * 100)
        print("warp-stall sampling mix:", ", ".join(f"{h[6:]} {v / tot * 100:.1f}%" for v, h in agg))
        print("top sampled instructions:")
        for r in sorted(data, key=lambda r: -num(r[idx["# Samples"]]))[:18]:
            s = num(r[idx["# Samples"]])
            top = max(stalls, key=lambda h: num(r[idx[h]]))
            print(f"  {s / tot * 100:5.1f}%  {r[idx['Source']][:84]:84s} {top[6:]}")


if __name__ == "__main__":
    main(sys.argv[1])
