"""The step's main-stream GEMMs at M = 14592 alone (CUDA events, L2 flushed): compare builds / env switches
(M3P_GEMM_TAIL_BALANCE=0|1).   python tools/bench_step_gemms.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from m3p_b200 import ops, lib as L

M = int(sys.argv[1]) if len(sys.argv) > 1 else 14592
ops.use_current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(name, n, k, epi, f32=False, drop_p=0.0, ln=False, b_mn=False, reps=15):
    a = torch.randn(M, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    if b_mn:
        w = w.t().contiguous()
    bias = torch.randn(n, device="cuda")
    odt = torch.float32 if f32 else torch.bfloat16
    out = torch.empty(M, n, device="cuda", dtype=odt)
    kw = dict(b_mn=b_mn)
    if f32:
        kw["out_f32"] = True
    if epi == L.M3P_EPI_GELU:
        kw["out2"] = torch.empty_like(out)
    if epi in (L.M3P_EPI_DROP_RES, L.M3P_EPI_DGELU):
        kw["aux"] = torch.randn(M, n, device="cuda").to(odt)
    if epi == L.M3P_EPI_DROP_RES:
        kw["drop_p"], kw["seed"] = drop_p, 5
    if ln:
        kw["aux_ln"] = (torch.randn(M, device="cuda"), torch.rand(M, device="cuda") + 0.5, torch.randn(n, device="cuda"),
                        torch.randn(n, device="cuda"), None, 0)
    ts = []
    for it in range(reps + 3):
        flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        ops.gemm(a, w, M, n, k, out, bias=bias if epi != L.M3P_EPI_DGELU else None, epi=epi, **kw)
        t1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(t0.elapsed_time(t1) * 1e3)
    ts.sort()
    print("%-34s n=%4d k=%4d  median %6.1f us  min %6.1f" % (name, n, k, ts[len(ts) // 2], ts[0]), flush=True)


print("M3P_GEMM_TAIL_BALANCE =", os.environ.get("M3P_GEMM_TAIL_BALANCE", "1 (default)"))
run("qkv           LINEAR", 2304, 768, L.M3P_EPI_LINEAR)
run("out_lin       DROP_RES f32 + LN", 768, 768, L.M3P_EPI_DROP_RES, f32=True, drop_p=0.1, ln=True)
run("lin1          GELU", 3072, 768, L.M3P_EPI_GELU)
run("lin2          DROP_RES f32 + LN", 768, 3072, L.M3P_EPI_DROP_RES, f32=True, drop_p=0.1, ln=True)
run("lin2 dgrad    DGELU (MN-major W)", 3072, 768, L.M3P_EPI_DGELU, b_mn=True)
run("lin1 dgrad    DROP_RES f32", 768, 3072, L.M3P_EPI_DROP_RES, f32=True, b_mn=True)
run("out_lin dgrad LINEAR", 768, 768, L.M3P_EPI_LINEAR, b_mn=True)
run("qkv dgrad     DROP_RES f32", 768, 2304, L.M3P_EPI_DROP_RES, f32=True, b_mn=True)
