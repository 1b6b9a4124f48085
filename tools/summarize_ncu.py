#!/usr/bin/env python
"""`ncu -i X.ncu-rep --page raw --csv` -> a markdown table of the metrics the roofline rests on.

    python tools/summarize_ncu.py gpurun_out/prof_raw.csv > profiles/<name>.md"""
import csv
import json
import sys

M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
     ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
     ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefront %"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
     ("smsp__warps_active.avg.per_cycle_active", "warps/SMSP"),
     ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    print("| kernel | " + " | ".join(n for _, n in M) + " |")
    print("|---|" + "---:|" * len(M))
    traffic = {}
    for r in data:
        name = r[col["Kernel Name"]]
        short = name.split("(")[0].replace("void ", "").replace("unflow::", "")
        cells = []
        for key, _ in M:
            if key not in col:
                cells.append("-")
                continue
            v, u = r[col[key]], units[col[key]]
            try:
                f = float(v.replace(",", ""))
                cells.append(("%.1f %s" % (f, u)) if u not in ("", "%") else ("%.1f" % f if "." in v else v))
            except ValueError:
                cells.append(v)
        print("| `%s` | " % short[:46] + " | ".join(cells) + " |")
        try:
            rd = float(r[col["dram__bytes_read.sum"]]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]]) * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[units[col["dram__bytes_write.sum"]]]
            traffic.setdefault(short, []).append(int(rd + wr))
        except Exception:
            pass
    if len(sys.argv) > 2:
        json.dump(traffic, open(sys.argv[2], "w"), indent=1)


if __name__ == "__main__":
    main()
