"""One M3P-base training step (64 pairs, fwd+bwd) bracketed by cudaProfilerStart/Stop, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--heads multitask] [--batch 64]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402

import bench  # noqa: E402
from m3p_b200.train_step import pretrain_step, synthetic_batch  # noqa: E402
from m3p_b200.transformer import TransformerModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--heads", default="itm")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--layers", type=int, default=0, help="override the layer count (ncu --set full captures)")
a = ap.parse_args()
cfg = dict(bench.CFG)
if a.layers:
    cfg["n_layers"] = a.layers
torch.manual_seed(0)
model = TransformerModel(bench.namespace(cfg), is_encoder=True, with_output=True, is_crossModal=True).cuda().train()
batch = synthetic_batch(a.batch, cfg["T"], cfg["R"], cfg["n_words"], sample_n=cfg["sample_n"], seed=1234, device="cuda")


def step():
    model.zero_grad()
    total, _ = pretrain_step(model, batch, cfg["sample_n"], bench.HEADS[a.heads])
    total.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.steps):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
