"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel launches, total/avg us, share."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"^void |gg::", "", name)
        val = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
        rows.append((name, us))
    agg = OrderedDict()
    for n, us in rows:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {t:.1f} | {t / c:.1f} | {t / tot:.3f} |")
    print(f"| all | {len(rows)} | {tot:.1f} | | |")


if __name__ == "__main__":
    main(sys.argv[1])
