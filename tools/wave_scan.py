"""How does the N = 768 GEMM time grow with the number of tile waves?  units = 3 * m_tiles on 74 CTA pairs: the slope is
the per-wave cost, the intercept the fixed launch / ramp / drain cost.  CUDA-event medians, L2 flushed.
    python tools/wave_scan.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from m3p_b200 import ops, lib as L

ops.use_current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(name, m, n, k, epi, f32=False, drop_p=0.0, ln=False, reps=15):
    a = torch.randn(m, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(n, device="cuda")
    odt = torch.float32 if f32 else torch.bfloat16
    out = torch.empty(m, n, device="cuda", dtype=odt)
    kw = {}
    if f32:
        kw["out_f32"] = True
    if epi == L.M3P_EPI_GELU:
        kw["out2"] = torch.empty_like(out)
    if epi in (L.M3P_EPI_DROP_RES, L.M3P_EPI_DGELU):
        kw["aux"] = torch.randn(m, n, device="cuda").to(odt)
    if epi == L.M3P_EPI_DROP_RES:
        kw["drop_p"], kw["seed"] = drop_p, 5
    if ln:
        kw["aux_ln"] = (torch.randn(m, device="cuda"), torch.rand(m, device="cuda") + 0.5, torch.randn(n, device="cuda"),
                        torch.randn(n, device="cuda"), None, 0)
    ts = []
    for it in range(reps + 3):
        flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        ops.linear(a, w, bias, out, epi=epi, **kw)
        t1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(t0.elapsed_time(t1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for name, n, k, epi, kw in (
        ("LINEAR bf16      n=768 k=768 ", 768, 768, L.M3P_EPI_LINEAR, {}),
        ("DROP_RES f32+LN  n=768 k=768 ", 768, 768, L.M3P_EPI_DROP_RES, dict(f32=True, drop_p=0.1, ln=True)),
        ("DROP_RES f32+LN  n=768 k=3072", 768, 3072, L.M3P_EPI_DROP_RES, dict(f32=True, drop_p=0.1, ln=True)),
        ("LINEAR bf16      n=2304 k=768", 2304, 768, L.M3P_EPI_LINEAR, {}),
        ("GELU             n=3072 k=768", 3072, 768, L.M3P_EPI_GELU, {}),
        ("DGELU            n=3072 k=768", 3072, 768, L.M3P_EPI_DGELU, {})):
    ntile = (n + 255) // 256
    line = []
    for waves in (1, 2, 3, 4, 6):
        mt = (74 * waves) // ntile  # m tiles so that units <= 74 * waves (just under `waves` full waves)
        m = mt * 256
        line.append("%dw(%3d units) %6.1f" % (waves, mt * ntile, run(name, m, n, k, epi, **kw)))
    print(name, " | ".join(line), flush=True)
