"""Aggregates an `ncu --page source --csv` dump: stall mix and the hottest SASS rows of the first kernel in the file.
usage: ncu -i rep --page source --csv -k regex:<kernel> > src.csv ; python tools/ncu_source_summary.py src.csv [top]"""
import csv
import sys
from collections import defaultdict


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    out = []
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            h = rows[i + 1]
            ci = {c: k for k, c in enumerate(h)}
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                body.append(rows[j])
                j += 1
            out.append((name, ci, body))
            i = j
        else:
            i += 1
    for name, ci, body in out[:1]:
        print(name, len(body), "SASS rows")
        tot_i = sum(float(r[ci["Instructions Executed"]] or 0) for r in body if len(r) > ci["Instructions Executed"])
        tot_s = sum(float(r[ci["# Samples"]] or 0) for r in body if len(r) > ci["# Samples"])
        print("instructions executed (one launch)", tot_i, "samples", tot_s)
        stall_cols = [c for c in ci if c.startswith("stall_") and "Not Issued" not in c]
        st = defaultdict(float)
        for r in body:
            for c in stall_cols:
                try:
                    st[c] += float(r[ci[c]] or 0)
                except (ValueError, IndexError):
                    pass
        tot = sum(st.values()) or 1
        print("stalls:", ", ".join(f"{k[6:]} {v / tot:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
        body_s = sorted([r for r in body if len(r) > ci["# Samples"]], key=lambda r: -float(r[ci["# Samples"]] or 0))
        for r in body_s[:top]:
            print(f'{r[ci["Address"]][-5:]:>6} smp {float(r[ci["# Samples"]] or 0):7.0f} inst {float(r[ci["Instructions Executed"]] or 0):10.0f} thr {r[ci["Avg. Threads Executed"]][:5]:>5}  {r[ci["Source"]][:90]}')


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
