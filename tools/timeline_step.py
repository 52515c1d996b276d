"""In-situ kernel timeline of the training step (warm caches, real stream concurrency) from CUPTI through
torch.profiler — the complement of the ncu launch list, whose per-launch times are cold-cache and serialised.

    python tools/timeline_step.py [--heads multitask] [--steps 3] [--graph] > gpurun_out/timeline.txt
Prints: wall span per step, per-kernel totals (us per step), and the ordered list of one forward / backward layer.
"""
import argparse
import os
import sys
from collections import OrderedDict, defaultdict

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from m3p_b200.train_step import GraphedStep, pretrain_step, synthetic_batch  # noqa: E402
from m3p_b200.transformer import TransformerModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--heads", default="itm")
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--list", action="store_true", help="print every kernel of the last step in start order")
a = ap.parse_args()
cfg = dict(bench.CFG)
torch.manual_seed(0)
model = TransformerModel(bench.namespace(cfg), is_encoder=True, with_output=True, is_crossModal=True).cuda().train()
batch = synthetic_batch(a.batch, cfg["T"], cfg["R"], cfg["n_words"], sample_n=cfg["sample_n"], seed=1234, device="cuda")
heads = bench.HEADS[a.heads]
graphed = GraphedStep(model, batch, cfg["sample_n"], heads, warmup=3) if a.graph else None


def step():
    if graphed is not None:
        graphed.step()
        return
    model.zero_grad()
    total, _ = pretrain_step(model, batch, cfg["sample_n"], heads)
    total.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.device_time_total > 0
       and "Memcpy" not in e.name and "Memset" not in e.name]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
print("# %d kernels in %d steps; wall span %.3f ms per step; summed kernel time %.3f ms per step" % (
    len(evs), a.steps, (t1 - t0) / 1e3 / a.steps, sum(e.device_time_total for e in evs) / 1e3 / a.steps))
agg = defaultdict(lambda: [0, 0.0])
for e in evs:
    n = e.name.split("(")[0].replace("void m3p::", "").replace("m3p::", "")
    agg[n][0] += 1
    agg[n][1] += e.device_time_total
print("%-72s %8s %10s %8s" % ("kernel", "n/step", "us/step", "us/call"))
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print("%-72s %8.1f %10.1f %8.1f" % (n[:72], c / a.steps, t / a.steps, t / c))
if a.list:
    per = len(evs) // a.steps
    base = evs[-per].time_range.start
    for e in evs[-per:]:
        print("%9.1f %8.1f  %s" % (e.time_range.start - base, e.device_time_total,
                                   e.name.split("(")[0].replace("void m3p::", "").replace("m3p::", "")[:80]))
