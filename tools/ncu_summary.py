"""One line per profiled launch from an `ncu -i X.ncu-rep --page raw --csv` dump: duration, DRAM
bytes, registers, SM / tensor-pipe / L2 utilisation, achieved warps.
  python tools/ncu_summary.py raw.csv > profiles/rNN_..._summary.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("time_us", "gpu__time_duration.sum"), ("dram_rd_MB", "dram__bytes_read.sum"),
        ("dram_wr_MB", "dram__bytes_write.sum"), ("regs", "launch__registers_per_thread"),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("tensor_pipe_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active")]


def scale(v, u, name):
    try:
        x = float(v)
    except ValueError:
        return v
    if name.endswith("_MB"):
        x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
    if name == "time_us":
        x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    return f"{x:.2f}"


print("kernel | grid | block | " + " | ".join(n for n, _ in want))
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]][:70]
    cells = [name, r[col["Grid Size"]], r[col["Block Size"]]]
    for n, m in want:
        cells.append(scale(r[col[m]], units[col[m]], n) if m in col else "-")
    print(" | ".join(cells))
