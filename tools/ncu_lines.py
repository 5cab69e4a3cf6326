"""Per CUDA source line: instructions executed and stall samples (ncu --page source --csv --print-source sass,cuda).
usage: ncu -i rep --page source --csv --print-source sass,cuda -k regex:<kernel> > x.csv; python tools/ncu_lines.py x.csv [top]"""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    hdr = None
    cur_file = ""
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = {c: k for k, c in enumerate(r)}
            # the header has "Source" twice (CUDA text, SASS text): take the first
            hdr["Source"] = r.index("Source")
            continue
        if hdr and r[0].isdigit():
            try:
                lines.append((cur_file, int(r[0]), r[hdr["Source"]].strip()[:100], float(r[hdr["Instructions Executed"]] or 0), float(r[hdr["# Samples"]] or 0)))
            except (ValueError, IndexError):
                pass
    ti = sum(l[3] for l in lines) or 1
    ts = sum(l[4] for l in lines) or 1
    print("total instr %.0f samples %.0f" % (ti, ts))
    for f, ln, src, ins, smp in sorted(lines, key=lambda l: -l[3])[:top]:
        print(f"{f}:{ln:5d} instr {ins / ti:6.3f} smp {smp / ts:6.3f}  {src}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
