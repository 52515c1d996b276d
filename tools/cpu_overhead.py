"""How long does the HOST take to enqueue one step (no sync), against the GPU time of the step?"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from m3p_b200.train_step import pretrain_step, synthetic_batch
from m3p_b200.transformer import TransformerModel
cfg = bench.CFG
torch.manual_seed(0)
model = TransformerModel(bench.namespace(cfg), is_encoder=True, with_output=True, is_crossModal=True).cuda().train()
batch = synthetic_batch(64, cfg["T"], cfg["R"], cfg["n_words"], sample_n=4, seed=1234, device="cuda")
def step():
    model.zero_grad()
    total, _ = pretrain_step(model, batch, 4, ("rel",))
    total.backward()
for _ in range(5): step()
torch.cuda.synchronize()
N = 20
t0 = time.perf_counter()
for _ in range(N): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host enqueue %.2f ms/step, wall %.2f ms/step" % ((t1 - t0) / N * 1e3, (t2 - t0) / N * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
