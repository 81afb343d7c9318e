#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (and optionally by grid)."""
import collections
import csv
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    out = []
    for x in csv.DictReader(lines):
        v = float(x["Metric Value"].replace(",", ""))
        u = x["Metric Unit"]
        ns = v * 1e3 if u.startswith("us") else (v * 1e6 if u.startswith("ms") else v)
        out.append((x["Kernel Name"].split("(")[0].replace("void ", ""), x["Grid Size"], ns))
    return out


def main():
    rows = load(sys.argv[1])
    steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    agg = collections.defaultdict(lambda: [0, 0.0])
    for k, _, ns in rows:
        agg[k][0] += 1
        agg[k][1] += ns
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot / 1e6:.2f} ms over {len(rows)} launches ({tot / 1e6 / steps:.2f} ms/step for {steps:g} steps)")
    print("| kernel | launches/step | ms/step | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
        print(f"| {k[:80]} | {v[0] / steps:.1f} | {v[1] / 1e6 / steps:.3f} | {100 * v[1] / tot:.1f}% |")
    if len(sys.argv) > 3:
        pat = sys.argv[3]
        print(f"\nlargest launches matching '{pat}':")
        sel = sorted([r for r in rows if pat in r[0]], key=lambda r: -r[2])[:40]
        for k, g, ns in sel:
            print(f"  {ns / 1e3:9.1f} us grid {g}")


if __name__ == "__main__":
    main()
