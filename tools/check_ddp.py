"""Run under torchrun with N >= 2 ranks on one box: the data-parallel gradients (each rank steps on its
shard, GradReducer averages over NCCL) must equal the single-process gradients of the same model on
the concatenated batch (ITM groups never straddle a shard boundary)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from m3p_b200.ddp import GradReducer, init_distributed  # noqa: E402
from m3p_b200.train_step import prepare_batch, pretrain_step, synthetic_batch  # noqa: E402
from m3p_b200.transformer import TransformerModel  # noqa: E402

rank, local, world = init_distributed()
cfg = dict(bench.CFG, n_layers=2, n_words=3000, dropout=0.0)
torch.manual_seed(0)
model = TransformerModel(bench.namespace(cfg), is_encoder=True, with_output=True, is_crossModal=True).cuda().train()
bf16 = "--bf16" in sys.argv
reducer = GradReducer(model, reduce_dtype=torch.bfloat16 if bf16 else None)
B = 8
full = synthetic_batch(B * world, cfg["T"], cfg["R"], cfg["n_words"], sample_n=4, seed=77, ragged=True, device="cuda")


def shard(lo, hi):
    out = {}
    for k, v in full.items():
        if k in ("x", "x_img", "image_loc", "x_labels"):
            out[k] = v[:, lo:hi].contiguous()
        elif k == "pos_labels":
            out[k] = v[lo // 4:hi // 4]
        elif k in ("lengths", "lengths_img", "obj_labels", "ori_feats"):
            out[k] = v[lo:hi]
    return prepare_batch(out)  # y_text / pred_mask_text / mrfr_weight of THIS shard (local means)


mine = shard(rank * B, (rank + 1) * B)
for heads in (("rel",), ("mlm", "mrm", "mrfr", "rel")):
    model._grad_ready_hook = reducer._segment_ready
    model._defer_token_grads = True
    model.zero_grad()
    total, _ = pretrain_step(model, mine, 4, heads=heads)
    total.backward()
    reducer.finish()
    torch.cuda.synchronize()
    dp_flat, dp_emb = model._flat_grad.clone(), model._emb_grad.clone()
    # reference: every shard on this rank, no communication, mean of the per-shard gradients
    model._grad_ready_hook = None
    model._defer_token_grads = False
    acc_flat, acc_emb = torch.zeros_like(dp_flat), torch.zeros_like(dp_emb)
    for r in range(world):
        model.zero_grad()
        t, _ = pretrain_step(model, shard(r * B, (r + 1) * B), 4, heads=heads)
        t.backward()
        acc_flat += model._flat_grad
        acc_emb += model._emb_grad
    acc_flat /= world
    acc_emb /= world
    e1 = float((dp_flat - acc_flat).norm() / acc_flat.norm())
    e2 = float((dp_emb - acc_emb).norm() / acc_emb.norm())
    res = torch.tensor([e1, e2], device="cuda")
    dist.all_reduce(res, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"case": "ddp_vs_single", "heads": list(heads), "world": world, "flat_rel_err": float(res[0]),
                          "emb_rel_err": float(res[1]), "reduce": "bf16" if bf16 else "fp32",
                          "ok": float(res.max()) < (4e-3 if bf16 else 1e-5)}), flush=True)
dist.barrier()
dist.destroy_process_group()
