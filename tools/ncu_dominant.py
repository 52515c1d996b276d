"""profiles/r02_ncu_dominant.json from an `ncu --set full` capture: DRAM bytes per launch of the dominant kernel
(the FFN lin1 + GELU GEMM, gemm_kernel<256, 1, false, true>) — what bench.py reports as roofline.traffic.

    ncu -i gpurun_out/<rep>.ncu-rep --page raw --csv > /tmp/raw.csv ; python tools/ncu_dominant.py /tmp/raw.csv <rep name>
"""
import csv
import json
import os
import sys

rd = csv.reader(open(sys.argv[1]))
hdr = next(rd)
units = next(rd)
idx = {h: i for i, h in enumerate(hdr)}


def val(row, key):
    v = float(row[idx[key]].replace(",", ""))
    u = units[idx[key]]
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)


rows = [r for r in rd if "gemm_kernel<256, 1," in r[idx["Kernel Name"]] or "gemm_kernel<256, 1, 0, 1>" in r[idx["Kernel Name"]]]
assert rows, "no GELU GEMM launch in the capture"
out = {"kernel": "gemm_kernel<256, M3P_EPI_GELU, bf16 out, cta_group::2> (FFN lin1 14592 x 3072 x 768)", "launches": len(rows),
       "dram_bytes_read": sum(val(r, "dram__bytes_read.sum") for r in rows) / len(rows),
       "dram_bytes_write": sum(val(r, "dram__bytes_write.sum") for r in rows) / len(rows),
       "duration_us_under_ncu": sum(val(r, "gpu__time_duration.sum") for r in rows) / len(rows),
       "tensor_pipe_pct": sum(float(r[idx["sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active"]])
                              for r in rows) / len(rows) if "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active" in idx else None,
       "source": "ncu --set full --clock-control none, %s" % (sys.argv[2] if len(sys.argv) > 2 else "capture")}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_ncu_dominant.json")
json.dump(out, open(p, "w"), indent=1)
print(json.dumps(out))
