#!/usr/bin/env python
"""Markdown summary of an `ncu -i X.ncu-rep --page raw --csv` dump: one row per captured launch with the metrics the
roofline discussion uses (B200_PROFILING.md).   python tools/ncu_summary.py raw.csv "title" > profiles/ncu_rNN_x.md"""
import csv
import sys

COLS = [("gpu__time_duration.sum", "time"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX %"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"), ("dram__bytes_read.sum", "DRAM read"),
        ("dram__bytes_write.sum", "DRAM write"), ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM bytes"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
        ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)")]


def main(path, title):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(k, n) for k, n in COLS if k in idx]
    print(f"# {title}\n")
    print(f"`ncu --set full --clock-control none` ({path}); per launch (cold-cache, serialised).\n")
    print("| kernel | grid | " + " | ".join(n for _, n in cols) + " |")
    print("|---|---|" + "---:|" * len(cols))
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        cells = []
        for k, _ in cols:
            v, u = r[idx[k]], units[idx[k]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            cells.append(f"{v} {u}".strip() if u not in ("%", "") else v)
        print(f"| `{name}` | {r[idx['Grid Size']]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
