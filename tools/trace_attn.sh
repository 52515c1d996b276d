set -e
cd $GRAFT_REPO_ROOT
cp m3p_b200/libm3p_sm100.so /tmp/lib_prod.so
M3P_NVCC_EXTRA=-DM3P_ATTN_TRACE python -m m3p_b200.build --force > /dev/null
python - <<'PY'
import torch
from m3p_b200 import ops, lib as L
print('occupancy fwd', L.load().m3p_debug_attn_occupancy(0), 'bwd', L.load().m3p_debug_attn_occupancy(1))
B,S,H=64,228,12; d=H*64
qkv=(torch.randn(B*S,3*d,device='cuda')*0.7).bfloat16()
seqlen=torch.full((B,),S,device='cuda',dtype=torch.int32)
ctx=torch.zeros(B*S,d,device='cuda',dtype=torch.bfloat16); lse=torch.zeros(B*H*S,device='cuda')
dctx=torch.randn(B*S,d,device='cuda').bfloat16(); dqkv=torch.zeros_like(qkv)
for it in range(2):
    ops.attention_fwd(qkv,seqlen,B,S,H,0.125,0.1,7,ctx,lse)
    torch.cuda.synchronize()
    ops.attention_bwd(qkv,seqlen,B,S,H,0.125,0.1,7,ctx,lse,dctx,dqkv)
    torch.cuda.synchronize()
PY
cp /tmp/lib_prod.so m3p_b200/libm3p_sm100.so
