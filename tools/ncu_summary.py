#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read with `ncu -i X.ncu-rep --page raw --csv`) as a markdown table.
usage: ncu_summary.py raw.csv [label ...]   labels name the captured launches in order (two captures per label are averaged)"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "MB read"), ("dram__bytes_write.sum", "MB written"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe % (elapsed)"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/smem %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid")]


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if len(r) < 10:
            continue
        d = {"kernel": r[idx["Kernel Name"]].split("(")[0].replace("void ", "")}
        for k, _ in KEYS:
            if k in idx:
                v = float(r[idx[k]].replace(",", ""))
                u = units[idx[k]]
                if u == "byte":
                    v /= 1e6
                elif u == "Kbyte":
                    v /= 1e3
                elif u == "Gbyte":
                    v *= 1e3
                elif u == "ns":
                    v /= 1e3
                elif u == "ms":
                    v *= 1e3
                d[k] = v
        out.append(d)
    return out


def main():
    rows = load(sys.argv[1])
    labels = sys.argv[2:]
    per = max(1, len(rows) // len(labels)) if labels else 1
    print("| launch | kernel | " + " | ".join(n for _, n in KEYS) + " |")
    print("|---|---|" + "---|" * len(KEYS))
    for i in range(0, len(rows), per):
        grp = rows[i:i + per]
        lab = labels[i // per] if labels and i // per < len(labels) else str(i)
        vals = []
        for k, _ in KEYS:
            xs = [g[k] for g in grp if k in g]
            vals.append(f"{sum(xs) / len(xs):.1f}" if xs else "-")
        print(f"| {lab} | {grp[0]['kernel'][:48]} | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    main()
