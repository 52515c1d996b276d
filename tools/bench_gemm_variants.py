"""Micro-benchmark of the N = 768 residual GEMM epilogues (CUDA events, 20 launches each, L2 flushed or warm)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from m3p_b200 import ops, lib as L

M = 14592
ops.use_current_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run(name, n, k, epi, f32=False, drop_p=0.0, cold=True, aux_small=False, reps=20):
    a = torch.randn(M, k, device="cuda").bfloat16()
    w = (torch.randn(n, k, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(n, device="cuda")
    odt = torch.float32 if f32 else torch.bfloat16
    out = torch.empty(M, n, device="cuda", dtype=odt)
    kw = {}
    if f32:
        kw["out_f32"] = True
    if epi == L.M3P_EPI_DROP_RES:
        kw["aux"] = torch.randn(M, n, device="cuda").to(odt)
        kw["drop_p"], kw["seed"] = drop_p, 5
    ts = []
    for it in range(reps + 3):
        if cold:
            flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        ops.linear(a, w, bias, out, epi=epi, **kw)
        t1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(t0.elapsed_time(t1) * 1e3)
    ts.sort()
    print("%-44s n=%4d k=%4d %s  median %6.1f us  min %6.1f" % (name, n, k, "cold" if cold else "warm", ts[len(ts) // 2], ts[0]), flush=True)


for cold in (True, False):
    run("LINEAR bf16", 768, 768, L.M3P_EPI_LINEAR, cold=cold)
    run("DROP_RES bf16 p=0", 768, 768, L.M3P_EPI_DROP_RES, cold=cold)
    run("DROP_RES bf16 p=0.1", 768, 768, L.M3P_EPI_DROP_RES, drop_p=0.1, cold=cold)
    run("DROP_RES f32 p=0", 768, 768, L.M3P_EPI_DROP_RES, f32=True, cold=cold)
    run("DROP_RES f32 p=0.1", 768, 768, L.M3P_EPI_DROP_RES, f32=True, drop_p=0.1, cold=cold)
    run("lin2 DROP_RES f32 p=0.1", 768, 3072, L.M3P_EPI_DROP_RES, f32=True, drop_p=0.1, cold=cold)
    run("lin2 LINEAR bf16", 768, 3072, L.M3P_EPI_LINEAR, cold=cold)
    run("qkv-dgrad-like DROP_RES f32 p=0", 768, 2304, L.M3P_EPI_DROP_RES, f32=True, cold=cold)
    run("lin1 GELU", 3072, 768, L.M3P_EPI_GELU, cold=cold) if False else None
