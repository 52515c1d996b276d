"""One small training step (all heads, dropout, side stream) for compute-sanitizer:
    compute-sanitizer --tool initcheck|memcheck|racecheck python tools/sanitize_step.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
from test_gpu_parity import _ns, _model
import m3p_b200.transformer as T
from m3p_b200.train_step import pretrain_step, synthetic_batch

ns = _ns(768, int(os.environ.get("SAN_LAYERS", "2")), 12, 3000, dropout=0.1)
b = synthetic_batch(8, 24, 10, ns.n_words, sample_n=4, seed=5, ragged=True, n_mask_text=3, n_mask_img=2, device="cuda")
model = _model(T, ns)
for _ in range(int(os.environ.get("SAN_STEPS", "2"))):
    model.zero_grad()
    total, _ = pretrain_step(model, b, 4)
    total.backward()
torch.cuda.synchronize()
print("loss", float(total.detach()), "grad norm", float(model._flat_grad.norm()))
