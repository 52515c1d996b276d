"""Print a compact per-kernel table from `ncu -i rep --page raw --csv` output."""
import csv
import sys

COLS = [
    ("gpu__time_duration.sum", "dur_us", 1e-3),
    ("dram__bytes_read.sum", "rd_MB", 1e-6),
    ("dram__bytes_write.sum", "wr_MB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "hmma%", 1),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tens%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
    ("launch__registers_per_thread", "regs", 1),
    ("launch__grid_size", "grid", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf", 1),
    ("lts__t_sector_hit_rate.pct", "l2hit%", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
]
rd = csv.reader(open(sys.argv[1]))
hdr = next(rd)
units = next(rd)
idx = {h: i for i, h in enumerate(hdr)}
pat = sys.argv[2] if len(sys.argv) > 2 else ""
if pat == "--list":
    for h in hdr:
        if len(sys.argv) < 4 or sys.argv[3] in h:
            print(h)
    sys.exit(0)
avail = [(m, n, s) for m, n, s in COLS if m in idx]
print("%-4s %-52s " % ("id", "kernel") + " ".join("%9s" % n for _, n, _ in avail))
for row in rd:
    name = row[idx["Kernel Name"]]
    if pat and pat not in name:
        continue
    vals = []
    for m, n, s in avail:
        try:
            v = float(row[idx[m]].replace(",", ""))
            u = units[idx[m]]
            if n == "dur_us":
                v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
            elif n.endswith("_MB"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            vals.append("%9.1f" % v)
        except (ValueError, KeyError):
            vals.append("%9s" % "-")
    print("%-4s %-52s " % (row[idx["ID"]], name[:52]) + " ".join(vals))
