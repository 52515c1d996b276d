"""GPU bring-up of the tcgen05 GEMM through the C ABI (run on the B200 box).

    python tools/bringup_gemm.py <case>       # one case in this process
    python tools/bringup_gemm.py all          # every case, each in its own subprocess + timeout

Each case prints one JSON line: shape, operand majors, max/rel error against torch fp32 matmul on
the same bf16 inputs, and (for timing cases) achieved TFLOP/s with CUDA events.
"""
import ctypes
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def _run_gemm(lib, L, torch, a, b, m, n, k, a_mn, b_mn, epi=0, out_f32=False, accumulate=False, split_k=1,
              bias=None, aux=None, alpha=1.0, drop_p=0.0, seed=0, out=None, dbg=None):
    dev = a.device
    if out is None:
        out = torch.empty(m, n, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
        if accumulate:
            out.zero_()
    out2 = torch.empty(m, n, device=dev, dtype=torch.bfloat16) if epi == L.M3P_EPI_GELU else None
    g = L.GemmArgs()
    g.a, g.b = a.data_ptr(), b.data_ptr()
    g.m, g.n, g.k = m, n, k
    g.lda, g.ldb = a.stride(0), b.stride(0)
    g.a_mn_major, g.b_mn_major = int(a_mn), int(b_mn)
    g.epilogue = epi
    g.out_f32, g.accumulate, g.split_k = int(out_f32), int(accumulate), split_k
    g.alpha = alpha
    g.bias = bias.data_ptr() if bias is not None else None
    g.out, g.ldo = out.data_ptr(), out.stride(0)
    g.out2, g.ldo2 = (out2.data_ptr(), out2.stride(0)) if out2 is not None else (None, 0)
    g.aux, g.ldaux = (aux.data_ptr(), aux.stride(0)) if aux is not None else (None, 0)
    g.drop_p, g.seed = drop_p, seed
    stream = torch.cuda.current_stream().cuda_stream
    if dbg is None:
        L.check(lib.m3p_gemm_bf16(ctypes.byref(g), ctypes.c_void_p(stream)), "m3p_gemm_bf16")
    else:
        L.check(lib.m3p_gemm_bf16_debug(ctypes.byref(g), *dbg, ctypes.c_void_p(stream)), "m3p_gemm_bf16_debug")
    return out, out2


def _operands(torch, m, n, k, a_mn, b_mn, seed=0):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    A = (torch.randn(m, k, device="cuda", generator=gen) * 0.5).to(torch.bfloat16)   # logical [m,k]
    B = (torch.randn(n, k, device="cuda", generator=gen) * 0.5).to(torch.bfloat16)   # logical [n,k]
    a_st = A.t().contiguous() if a_mn else A
    b_st = B.t().contiguous() if b_mn else B
    ref = A.float() @ B.float().t()
    return A, B, a_st, b_st, ref


def _err(torch, got, ref):
    d = (got.float() - ref).abs()
    return float(d.max()), float(d.norm() / (ref.norm() + 1e-30))


def case_layout(name, m, n, k, a_mn, b_mn, sweep=False):
    import torch
    from m3p_b200 import lib as L
    lib = L.load()
    L.check(lib.m3p_device_check(), "device_check")
    A, B, a_st, b_st, ref = _operands(torch, m, n, k, a_mn, b_mn)
    out, _ = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, a_mn, b_mn)
    torch.cuda.synchronize()
    mx, rel = _err(torch, out, ref)
    res = {"case": name, "m": m, "n": n, "k": k, "a_mn": a_mn, "b_mn": b_mn, "max_err": mx, "rel_err": rel,
           "ok": rel < 1e-2}
    print(json.dumps(res), flush=True)
    if not res["ok"] and sweep:
        # descriptor hypothesis sweep for MN-major operands: (lbo, sbo, kstep)
        cands = [(8192, 1024, 2048), (1024, 8192, 2048), (8192, 1024, 32), (1024, 8192, 32),
                 (128, 1024, 2048), (1024, 128, 2048), (8192, 128, 2048), (16, 1024, 2048),
                 (8192, 1024, 256), (1024, 8192, 256), (2048, 1024, 2048), (1024, 2048, 2048)]
        for ca in (cands if a_mn else [(-1, -1, -1)]):
            for cb in (cands if b_mn else [(-1, -1, -1)]):
                o, _ = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, a_mn, b_mn, dbg=(*ca, *cb))
                torch.cuda.synchronize()
                mx, rel = _err(torch, o, ref)
                print(json.dumps({"case": name + "/sweep", "a": ca, "b": cb, "rel_err": rel, "ok": rel < 1e-2}),
                      flush=True)
    return res["ok"]


def case_epilogues():
    import torch
    from m3p_b200 import lib as L
    lib = L.load()
    m, n, k = 300, 392, 200   # ragged in every dimension (n % 8 == 0 for vector stores)
    A, B, a_st, b_st, ref = _operands(torch, m, n, k, False, False, seed=1)
    bias = torch.randn(n, device="cuda")
    aux = torch.randn(m, n, device="cuda").to(torch.bfloat16)
    ok_all = True

    def report(name, got, want, tol=1e-2):
        nonlocal ok_all
        mx, rel = _err(torch, got, want)
        ok = rel < tol
        ok_all &= ok
        print(json.dumps({"case": "epi/" + name, "max_err": mx, "rel_err": rel, "ok": ok}), flush=True)

    o, _ = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, 0, 0, bias=bias, alpha=0.5)
    report("linear_bias_alpha", o, 0.5 * ref + bias)
    o, _ = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, 0, 0, out_f32=True, bias=bias)
    report("linear_f32", o, ref + bias, 1e-5)
    acc = torch.ones(m, n, device="cuda")
    o, _ = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, 0, 0, out_f32=True, accumulate=True, split_k=3, out=acc)
    report("splitk3_accumulate", o, ref + 1.0, 1e-5)
    o, o2 = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, 0, 0, epi=L.M3P_EPI_GELU, bias=bias)
    u = (ref + bias)
    report("gelu_grad_stash", o, 0.5 * (1 + torch.erf(u / 2 ** 0.5)) + u * torch.exp(-0.5 * u * u) / (2 * 3.141592653589793) ** 0.5)
    report("gelu_act", o2, 0.5 * u * (1 + torch.erf(u / 2 ** 0.5)))
    o, _ = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, 0, 0, epi=L.M3P_EPI_DROP_RES, bias=bias, aux=aux)
    report("res_nodrop", o, ref + bias + aux.float())
    o, _ = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, 0, 0, epi=L.M3P_EPI_DROP_RES, bias=bias, aux=aux,
                     drop_p=0.1, seed=1234)
    kept = ((o.float() - aux.float()).abs() > 0).float().mean().item()
    v = (ref + bias) / 0.9
    d = o.float() - aux.float()
    frac_ok = float((((d - v).abs() < 0.02 * v.abs() + 0.05) | (d.abs() < 1e-6)).float().mean())
    print(json.dumps({"case": "epi/dropout", "keep_frac": kept, "frac_consistent": frac_ok,
                      "ok": abs(kept - 0.9) < 0.01 and frac_ok > 0.999}), flush=True)
    o, _ = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, 0, 0, epi=L.M3P_EPI_DGELU, aux=aux)
    report("dgelu_mul", o, ref * aux.float())
    o, _ = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, 0, 0, epi=L.M3P_EPI_TANH, bias=bias, alpha=0.1)
    report("tanh", o, torch.tanh(0.1 * ref + bias))
    t = torch.tanh(aux.float()).to(torch.bfloat16)
    o, _ = _run_gemm(lib, L, torch, a_st, b_st, m, n, k, 0, 0, epi=L.M3P_EPI_DTANH, aux=t)
    report("dtanh", o, ref * (1 - t.float() ** 2))
    # odd n (scalar store path), tiny k
    m2, n2, k2 = 70, 250, 40
    A, B, a_st, b_st, ref = _operands(torch, m2, n2, k2, False, False, seed=2)
    outp = torch.zeros(m2, 256, device="cuda", dtype=torch.bfloat16)
    o, _ = _run_gemm(lib, L, torch, a_st, b_st, m2, n2, k2, 0, 0, out=outp)
    report("ragged_n250", outp[:, :n2], ref)
    torch.cuda.synchronize()
    return ok_all


def case_timing():
    import torch
    from m3p_b200 import lib as L
    lib = L.load()
    shapes = [("qkv", 14592, 2304, 768, 0, 0, 1, False), ("out", 14592, 768, 768, 0, 0, 1, False),
              ("ffn1", 14592, 3072, 768, 0, 0, 1, False), ("ffn2", 14592, 768, 3072, 0, 0, 1, False),
              ("dgrad_ffn2", 14592, 3072, 768, 0, 1, 1, False), ("dgrad_ffn1", 14592, 768, 3072, 0, 1, 1, False),
              ("wgrad_ffn1", 3072, 768, 14592, 1, 1, 4, True), ("wgrad_ffn2", 768, 3072, 14592, 1, 1, 4, True),
              ("wgrad_qkv", 2304, 768, 14592, 1, 1, 5, True), ("wgrad_out", 768, 768, 14592, 1, 1, 8, True),
              ("img", 6400, 768, 2048, 0, 0, 1, False)]
    for name, m, n, k, a_mn, b_mn, split, f32 in shapes:
        A, B, a_st, b_st, ref = _operands(torch, m, n, k, a_mn, b_mn, seed=3)
        out = torch.zeros(m, n, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
        for _ in range(3):
            _run_gemm(lib, L, torch, a_st, b_st, m, n, k, a_mn, b_mn, out_f32=f32, accumulate=f32, split_k=split,
                      out=out)
        torch.cuda.synchronize()
        if f32:
            out.zero_()
            _run_gemm(lib, L, torch, a_st, b_st, m, n, k, a_mn, b_mn, out_f32=f32, accumulate=f32, split_k=split,
                      out=out)
        mx, rel = _err(torch, out, ref)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            _run_gemm(lib, L, torch, a_st, b_st, m, n, k, a_mn, b_mn, out_f32=f32, accumulate=f32, split_k=split,
                      out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        # cuBLAS on the same logical problem, for context
        Af, Bf = A, B
        for _ in range(3):
            torch.matmul(Af, Bf.t())
        e0.record()
        for _ in range(iters):
            torch.matmul(Af, Bf.t())
        e1.record()
        torch.cuda.synchronize()
        ms_cublas = e0.elapsed_time(e1) / iters
        print(json.dumps({"case": "time/" + name, "m": m, "n": n, "k": k, "a_mn": a_mn, "b_mn": b_mn,
                          "split_k": split, "rel_err": rel, "ms": ms, "tflops": 2.0 * m * n * k / ms / 1e9,
                          "cublas_ms": ms_cublas, "cublas_tflops": 2.0 * m * n * k / ms_cublas / 1e9}), flush=True)
    return True


CASES = {
    "tn_small": lambda: case_layout("tn_small", 128, 128, 64, 0, 0),
    "tn_k256": lambda: case_layout("tn_k256", 128, 256, 256, 0, 0),
    "tn_multi": lambda: case_layout("tn_multi", 1000, 776, 520, 0, 0),
    "nn_small": lambda: case_layout("nn_small", 128, 128, 64, 0, 1, sweep=True),
    "nn_multi": lambda: case_layout("nn_multi", 1000, 776, 520, 0, 1),
    "tt_small": lambda: case_layout("tt_small", 128, 128, 64, 1, 0, sweep=True),
    "nt_small": lambda: case_layout("nt_small", 128, 256, 128, 1, 1),
    "nt_multi": lambda: case_layout("nt_multi", 776, 1000, 520, 1, 1),
    "epilogues": case_epilogues,
    "timing": case_timing,
}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        rc = 0
        for name in CASES:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=300)
                if r.returncode != 0:
                    print(json.dumps({"case": name, "ok": False, "returncode": r.returncode}), flush=True)
                    rc = 1
            except subprocess.TimeoutExpired:
                print(json.dumps({"case": name, "ok": False, "timeout": True}), flush=True)
                rc = 1
        sys.exit(rc)
    ok = CASES[which]()
    sys.exit(0 if ok else 1)
