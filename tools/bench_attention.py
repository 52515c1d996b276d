"""Attention kernels alone at the bench shape (B=64, S=228, H=12, p=0.1): CUDA-event medians, and a checksum of the
backward's output so two builds / kernel variants (M3P_ATTN_BWD=oneshot|persist) can be compared bit for bit.
    python tools/bench_attention.py [B] [S]"""
import os, sys, hashlib
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from m3p_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 228
H, d = 12, 768
M = B * S
ops.use_current_stream()
g = torch.Generator(device="cuda").manual_seed(1)
qkv = (torch.randn(M, 3 * d, device="cuda", generator=g) * 0.8).bfloat16()
dctx = (torch.randn(M, d, device="cuda", generator=g) * 0.1).bfloat16()
seqlen = torch.full((B,), S, dtype=torch.int32, device="cuda")
seqlen[1::3] = max(1, S - 37)
seqlen[2::5] = max(1, S // 2 + 3)
ctx = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B * H * S, device="cuda", dtype=torch.float32)
dqkv = torch.empty(M, 3 * d, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
scale = 0.125


def timeit(fn, reps=20, cold=True):
    ts = []
    for it in range(reps + 3):
        if cold:
            flush.zero_()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        fn()
        t1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(t0.elapsed_time(t1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for p_att in (0.1, 0.0):
    f = lambda: ops.attention_fwd(qkv, seqlen, B, S, H, scale, p_att, 77, ctx, lse)
    b = lambda: ops.attention_bwd(qkv, seqlen, B, S, H, scale, p_att, 77, ctx, lse, dctx, dqkv)
    f()
    dqkv.zero_()
    b()
    torch.cuda.synchronize()
    sig = hashlib.sha256(dqkv.view(torch.int16).cpu().numpy().tobytes()).hexdigest()[:16]
    sigf = hashlib.sha256(ctx.view(torch.int16).cpu().numpy().tobytes()).hexdigest()[:16]
    for cold in (True, False):
        mf, nf = timeit(f, cold=cold)
        mb, nb = timeit(b, cold=cold)
        print("B=%d S=%d p=%.1f %s  fwd median %6.1f us (min %6.1f)  bwd median %6.1f us (min %6.1f)  ctx %s dqkv %s  finite %s" % (
            B, S, p_att, "cold" if cold else "warm", mf, nf, mb, nb, sigf, sig, bool(torch.isfinite(dqkv.float()).all())), flush=True)
