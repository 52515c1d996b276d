"""Run-to-run gradient noise floor of the training step: which heads / stream modes make two identical steps differ
(VERDICT r1 weak #1).  Prints one line per configuration."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_gpu_parity import _ns, _model, _rel
import m3p_b200.transformer as T
from m3p_b200.train_step import pretrain_step, synthetic_batch

ns = _ns(768, 3, 12, 3000, dropout=0.1)
b = synthetic_batch(8, 24, 10, ns.n_words, sample_n=4, seed=5, ragged=True, n_mask_text=3, n_mask_img=2, device="cuda")
for heads in (("rel",), ("mlm",), ("mrm",), ("mrfr",), ("mlm", "mrm", "mrfr", "rel")):
    runs = []
    for overlap in (False, False, True, True):
        model = _model(T, ns)
        model.overlap_grads = overlap
        for _ in range(2):
            model.zero_grad()
            total, _ = pretrain_step(model, b, 4, heads=heads)
            total.backward()
        torch.cuda.synchronize()
        runs.append((float(total.detach()), model._flat_grad.clone(), model._emb_grad.clone()))
    print("heads=%-28s inline-vs-inline flat %.2e emb %.2e | side-vs-inline flat %.2e emb %.2e | side-vs-side flat %.2e"
          % (",".join(heads), _rel(runs[1][1], runs[0][1]), _rel(runs[1][2], runs[0][2]), _rel(runs[2][1], runs[0][1]),
             _rel(runs[2][2], runs[0][2]), _rel(runs[3][1], runs[2][1])), flush=True)
