# Phase trace of the GEMM kernel (producer / MMA / epilogue SM-clock stamps per tile) on the step's shapes.
#   gpurun -- 'bash tools/trace_gemm.sh > gpurun_out/gemm_trace.txt 2>&1'
set -e
cd ${GRAFT_REPO_ROOT:-.}
cp m3p_b200/libm3p_sm100.so /tmp/lib_prod.so
M3P_NVCC_EXTRA=-DM3P_GEMM_TRACE python -m m3p_b200.build --force > /dev/null
python - "$@" <<'PY'
import torch
from m3p_b200 import ops, lib as L
M = 14592
def case(name, n, k, epi, f32=False, **kw):
    a = torch.randn(M, k, device='cuda').bfloat16()
    w = torch.randn(n, k, device='cuda').bfloat16() * 0.05
    bias = torch.randn(n, device='cuda')
    out = torch.empty(M, n, device='cuda', dtype=torch.float32 if f32 else torch.bfloat16)
    extra = {}
    if f32:
        extra['out_f32'] = True
    if epi == L.M3P_EPI_GELU:
        extra['out2'] = torch.empty_like(out)
    if epi in (L.M3P_EPI_DROP_RES, L.M3P_EPI_DGELU):
        extra['aux'] = torch.randn(M, n, device='cuda').to(out.dtype)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for it in range(2):
        flush.zero_()
        torch.cuda.synchronize()
        if it == 1:
            print('==== %s  N=%d K=%d' % (name, n, k), flush=True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        ops.linear(a, w, bias, out, epi=epi, **extra, **kw)
        t1.record()
        torch.cuda.synchronize()
        if it == 1:
            print('---- %.1f us' % (t0.elapsed_time(t1) * 1e3), flush=True)
ops.use_current_stream()
import sys
which = sys.argv[1:] or ['all']
if 'f32' in which or 'all' in which:
    case('out_lin LINEAR', 768, 768, L.M3P_EPI_LINEAR)
    case('out_lin DROP_RES bf16', 768, 768, L.M3P_EPI_DROP_RES, drop_p=0.1, seed=5)
    case('out_lin DROP_RES f32 p=0.1', 768, 768, L.M3P_EPI_DROP_RES, f32=True, drop_p=0.1, seed=5)
    case('out_lin DROP_RES f32 p=0', 768, 768, L.M3P_EPI_DROP_RES, f32=True, drop_p=0.0, seed=5)
    case('lin2 DROP_RES f32', 768, 3072, L.M3P_EPI_DROP_RES, f32=True, drop_p=0.1, seed=5)
if 'all' in which:
    case('lin1 GELU', 3072, 768, L.M3P_EPI_GELU)
    case('qkv LINEAR', 2304, 768, L.M3P_EPI_LINEAR)
PY
cp /tmp/lib_prod.so m3p_b200/libm3p_sm100.so
