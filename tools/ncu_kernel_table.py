"""Per-kernel table of an `ncu --set full` report (raw page): time, DRAM bytes, utilisation; optionally writes the per-scan
DRAM traffic as JSON (profiles/<tag>_traffic.json) for bench.py's roofline.traffic.
usage: python tools/ncu_kernel_table.py report.ncu-rep scans_per_launch [traffic.json]"""
import csv
import json
import re
import subprocess
import sys

ALIAS = {"k_spiral_skew": "k_spiral", "k_detect_tma": "k_detect", "k_detect_ldg": "k_detect"}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]


def main(rep, scans, out_json=None):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, units = rows[0], rows[1]
    ci = {c: k for k, c in enumerate(h)}
    table = {}
    for r in rows[2:]:
        name = re.sub(r"^void |gg::|\(.*|<.*", "", r[ci["Kernel Name"]])
        grid = float(r[ci["launch__grid_size"]])
        d = {w: float(r[ci[w]].replace(",", "")) for w in WANT if w in ci and r[ci[w]] not in ("", "n/a")}
        d["_units"] = {w: units[ci[w]] for w in WANT if w in ci}
        if name not in table or grid > table[name]["launch__grid_size"]:
            table[name] = d                     # keep the batched launch (largest grid) of every kernel
    print("| kernel | grid x block | time us | DRAM read MB | DRAM write MB | DRAM % | SM % | issue % | warps % | regs | L1 hit % | L2 hit % | warp instr M | thr/instr |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    traffic = {}
    for name, d in sorted(table.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        u = d["_units"]
        tu = u["gpu__time_duration.sum"]
        t_us = d["gpu__time_duration.sum"] * {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6}.get(tu, 1.0)
        scale = {"Mbyte": 1.0, "Gbyte": 1000.0, "Kbyte": 0.001, "byte": 1e-6}
        rd = d["dram__bytes_read.sum"] * scale.get(u["dram__bytes_read.sum"], 1.0)
        wr = d["dram__bytes_write.sum"] * scale.get(u["dram__bytes_write.sum"], 1.0)
        print(f"| {name} | {int(d['launch__grid_size'])} x {int(d['launch__block_size'])} | {t_us:.0f} | {rd:.0f} | {wr:.0f} | "
              f"{d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 0):.1f} | {d.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', 0):.1f} | "
              f"{d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.1f} | {d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):.1f} | "
              f"{int(d.get('launch__registers_per_thread', 0))} | {d.get('l1tex__t_sector_hit_rate.pct', 0):.0f} | {d.get('lts__t_sector_hit_rate.pct', 0):.0f} | "
              f"{d.get('smsp__inst_executed.sum', 0) / 1e6:.0f} | {d.get('smsp__thread_inst_executed_per_inst_executed.ratio', 0):.1f} |")
        traffic[ALIAS.get(name, name)] = (rd + wr) * 1e6 / scans
    if out_json:
        json.dump({"source": rep, "scans_per_launch": scans, "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch / scans per launch",
                   "bytes_per_scan": traffic}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else None)
