"""GPU bring-up of the non-GEMM kernels and of the whole model through the C ABI (B200 box).

    python tools/bringup_model.py all          # every case in its own subprocess (a trap in one
    python tools/bringup_model.py <case>       # kernel cannot take the others down)

One JSON line per check.  The parity tests proper live in tests/ (-m gpu); this tool exists to get
per-kernel verdicts and timings out of a single gpurun call.
"""
import json
import math
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)


def _rel(got, ref):
    return float((got.float() - ref.float()).norm() / (ref.float().norm() + 1e-30))


def _out(**kw):
    print(json.dumps(kw), flush=True)


def _attn_ref(torch, qkv, seqlen, B, S, H, scale):
    q, k, v = qkv.view(B, S, 3, H, 64).permute(2, 0, 3, 1, 4)  # (B,H,S,64)
    sc = torch.matmul(q, k.transpose(2, 3)) * scale
    key = torch.arange(S, device=qkv.device)[None, :] < seqlen[:, None]
    sc = sc.masked_fill(~key[:, None, None, :], float("-inf"))
    w = torch.softmax(sc, dim=-1)
    return torch.matmul(w, v).transpose(1, 2).reshape(B * S, H * 64)


def case_attention(B, S, H, ragged, time_it=False):
    import torch
    from m3p_b200 import ops
    torch.manual_seed(0)
    d = H * 64
    qkv = (torch.randn(B * S, 3 * d, device="cuda") * 0.7).to(torch.bfloat16)
    if ragged:
        seqlen = torch.randint(max(1, S // 3), S + 1, (B,), device="cuda", dtype=torch.int32)
        seqlen[0] = S
    else:
        seqlen = torch.full((B,), S, device="cuda", dtype=torch.int32)
    scale = 1.0 / 8.0
    ctx = torch.zeros(B * S, d, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B * H * S, device="cuda")
    ops.attention_fwd(qkv, seqlen, B, S, H, scale, 0.0, 0, ctx, lse)
    torch.cuda.synchronize()
    q32 = qkv.float().requires_grad_(True)
    ref = _attn_ref(torch, q32, seqlen.long(), B, S, H, scale)
    dctx = (torch.randn(B * S, d, device="cuda") * 0.5).to(torch.bfloat16)
    # padded query rows carry no upstream gradient on the real path
    valid = (torch.arange(S, device="cuda")[None, :] < seqlen[:, None]).reshape(B * S, 1)
    dctx = dctx * valid
    ref.backward(dctx.float())
    e_fwd = _rel(ctx * valid, ref.detach() * valid)
    dqkv = torch.zeros(B * S, 3 * d, device="cuda", dtype=torch.bfloat16)
    ops.attention_bwd(qkv, seqlen, B, S, H, scale, 0.0, 0, ctx, lse, dctx, dqkv)
    torch.cuda.synchronize()
    g = q32.grad
    e_dq, e_dk, e_dv = _rel(dqkv[:, :d], g[:, :d]), _rel(dqkv[:, d:2 * d], g[:, d:2 * d]), _rel(dqkv[:, 2 * d:], g[:, 2 * d:])
    res = dict(case="attention", B=B, S=S, H=H, ragged=ragged, fwd=e_fwd, dq=e_dq, dk=e_dk, dv=e_dv,
               ok=max(e_fwd, e_dq, e_dk, e_dv) < 1.5e-2)
    if time_it:
        for name, fn in (("fwd_ms", lambda: ops.attention_fwd(qkv, seqlen, B, S, H, scale, 0.1, 7, ctx, lse)),
                         ("bwd_ms", lambda: ops.attention_bwd(qkv, seqlen, B, S, H, scale, 0.1, 7, ctx, lse, dctx, dqkv))):
            for _ in range(3):
                fn()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                fn()
            t1.record()
            torch.cuda.synchronize()
            res[name] = t0.elapsed_time(t1) / 20
    _out(**res)


def case_attention_dropout():
    import torch
    from m3p_b200 import ops
    B, S, H = 3, 228, 2
    d = H * 64
    torch.manual_seed(1)
    qkv = (torch.randn(B * S, 3 * d, device="cuda") * 0.5).to(torch.bfloat16)
    # v = identity-like probe: with v columns = one-hot of (key % 64) we can read the kept mass
    seqlen = torch.full((B,), S, device="cuda", dtype=torch.int32)
    ctx0, ctx1, ctx2 = (torch.zeros(B * S, d, device="cuda", dtype=torch.bfloat16) for _ in range(3))
    lse = torch.zeros(B * H * S, device="cuda")
    ops.attention_fwd(qkv, seqlen, B, S, H, 0.125, 0.0, 0, ctx0, lse)
    ops.attention_fwd(qkv, seqlen, B, S, H, 0.125, 0.1, 123, ctx1, lse)
    ops.attention_fwd(qkv, seqlen, B, S, H, 0.125, 0.1, 123, ctx2, lse)
    torch.cuda.synchronize()
    same = bool((ctx1 == ctx2).all())
    differs = _rel(ctx1, ctx0)
    # finite-difference check of the backward under dropout: d/dv is linear -> compare with autograd on
    # the implied weights is not available; check instead <dctx, J dv> symmetry on a random direction
    dctx = (torch.randn(B * S, d, device="cuda") * 0.5).to(torch.bfloat16)
    dqkv = torch.zeros(B * S, 3 * d, device="cuda", dtype=torch.bfloat16)
    ops.attention_bwd(qkv, seqlen, B, S, H, 0.125, 0.1, 123, ctx1, lse, dctx, dqkv)
    eps_dir = torch.zeros_like(qkv, dtype=torch.float32)
    eps_dir[:, 2 * d:] = torch.randn(B * S, d, device="cuda") * 0.5          # perturb V only (linear)
    qkv2 = (qkv.float() + eps_dir).to(torch.bfloat16)
    real_dir = qkv2.float() - qkv.float()
    ctx3 = torch.zeros_like(ctx1)
    ops.attention_fwd(qkv2, seqlen, B, S, H, 0.125, 0.1, 123, ctx3, lse)
    torch.cuda.synchronize()
    lhs = float(((ctx3.float() - ctx1.float()) * dctx.float()).sum())
    rhs = float((dqkv.float() * real_dir).sum())
    _out(case="attention_dropout", deterministic=same, rel_change=differs, lhs=lhs, rhs=rhs,
         ok=same and 0.05 < differs < 1.0 and abs(lhs - rhs) < 0.03 * max(abs(lhs), abs(rhs), 1.0))


def case_layernorm():
    import torch
    import torch.nn.functional as F
    from m3p_b200 import ops
    torch.manual_seed(0)
    B, S, d = 5, 37, 768
    rows = B * S
    x = torch.randn(rows, d, device="cuda").to(torch.bfloat16)
    gam, bet = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    seqlen = torch.randint(5, S + 1, (B,), device="cuda", dtype=torch.int32)
    y = torch.empty_like(x)
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    ops.layernorm_fwd(x, gam, bet, y, mean, rstd, 1e-12, seqlen=seqlen, S=S)
    mask = (torch.arange(S, device="cuda")[None, :] < seqlen[:, None]).reshape(rows, 1).float()
    x32 = x.float().requires_grad_(True)
    g32, b32 = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    ref = F.layer_norm(x32, (d,), g32, b32, 1e-12) * mask
    dy = torch.randn(rows, d, device="cuda").to(torch.bfloat16)
    ref.backward(dy.float())
    dx = torch.empty_like(x)
    dgam, dbet, dbias = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    ops.layernorm_bwd(dy, x, mean, rstd, gam, dx, seqlen=seqlen, S=S, dgamma=dgam, dbeta=dbet, dbias=dbias)
    torch.cuda.synchronize()
    errs = dict(fwd=_rel(y, ref.detach()), dx=_rel(dx, x32.grad), dgamma=_rel(dgam, g32.grad), dbeta=_rel(dbet, b32.grad),
                dbias=_rel(dbias, x32.grad.sum(0)))
    pad_zero = float((y.float() * (1 - mask)).abs().max())
    _out(case="layernorm", pad_max=pad_zero, ok=max(errs.values()) < 1e-2 and pad_zero == 0.0, **errs)


def case_cross_entropy():
    import torch
    import torch.nn.functional as F
    from m3p_b200 import ops
    torch.manual_seed(0)
    res = {}
    ok = True
    for n, V, ign in ((64, 1600, -1), (33, 1002, -100)):
        ld = (V + 7) // 8 * 8
        logits = torch.zeros(n, ld, device="cuda", dtype=torch.bfloat16)
        logits[:, :V] = (torch.randn(n, V, device="cuda") * 3).to(torch.bfloat16)
        y = torch.randint(0, V, (n,), device="cuda")
        if ign == -1:
            y[::3] = -1
        l32 = logits[:, :V].float().requires_grad_(True)
        ref = F.cross_entropy(l32, y, ignore_index=ign)
        (ref * 0.7).backward()
        loss, lse, inv = torch.zeros((), device="cuda"), torch.zeros(n, device="cuda"), torch.zeros((), device="cuda")
        ops.cross_entropy_fwd(logits, y, V, ign, loss, lse, inv)
        gs = torch.full((), 0.7, device="cuda")
        dl = torch.empty_like(logits)
        ops.cross_entropy_bwd(logits, y, V, ign, lse, inv, gs, dl)
        torch.cuda.synchronize()
        e1, e2 = abs(float(loss) - float(ref)) / abs(float(ref)), _rel(dl[:, :V], l32.grad)
        res["loss_%d" % V], res["dlogits_%d" % V] = e1, e2
        ok = ok and e1 < 1e-4 and e2 < 1e-2
    _out(case="cross_entropy", ok=ok, **res)


def _namespace(d, L_, H, V, n_langs=1, dropout=0.0, refine_layers=1):
    import argparse
    langs = ["en", "fr", "de", "zh"][:n_langs]
    return argparse.Namespace(
        n_langs=n_langs, n_words=V, eos_index=2, pad_index=1, id2lang={i: l for i, l in enumerate(langs)},
        lang2id={l: i for i, l in enumerate(langs)}, emb_dim=d, n_heads=H, n_layers=L_, n_dec_layers=L_,
        dropout=dropout, attention_dropout=dropout, sinusoidal_embeddings=False, refine_layers=refine_layers,
        attention_setting="v1", use_externel_att=False, gelu_activation=True, share_inout_emb=True, asm=False)


def _step(torch, F, model, batch, sample_n, heads):
    """pretrain_under_step loss assembly (xtrainer.py:2285-2375) over the model API."""
    R = batch["x_img"].shape[0]
    enc = model("jointfwd", x=batch["x"], lengths=batch["lengths"], x_img=batch["x_img"],
                lengths_img=batch["lengths_img"], causal=False, langs=None, image_loc=batch["image_loc"],
                refine_image=False)
    text_out, img_out = enc[R:], enc[:R].transpose(0, 1)
    losses = {}
    total = 0.0
    if "mlm" in heads:
        pm = batch["x_labels"] != -1
        y = batch["x_labels"][batch["x_labels"] > 0]
        _, losses["mlm"] = model("predict", tensor=text_out, pred_mask=pm, y=y, get_scores=False)
        total = total + losses["mlm"]
    if "mrm" in heads:
        _, losses["mrm"] = model("predict", tensor=img_out, pred_mask=None, y=batch["obj_labels"].view(-1),
                                 get_scores=False, is_obj=True)
        total = total + losses["mrm"]
    if "mrfr" in heads:
        reg = model("predict", tensor=img_out, is_mrfr=True)
        sel = batch["obj_labels"].reshape(-1) != -1
        losses["mrfr"] = F.mse_loss(reg.reshape(-1, 2048)[sel].float(), batch["ori_feats"].reshape(-1, 2048)[sel])
        total = total + losses["mrfr"]
    if "rel" in heads:
        sc = model("predict", tensor=enc.transpose(0, 1), is_relation=True)
        ce = F.cross_entropy(sc.view(-1, sample_n), batch["pos_labels"])
        bce = F.binary_cross_entropy_with_logits(sc.view(-1), F.one_hot(batch["pos_labels"], sample_n).float().view(-1))
        losses["rel"] = ce + bce
        total = total + losses["rel"]
    return enc, losses, total


def case_golden(name):
    import torch
    import torch.nn.functional as F
    from m3p_b200.transformer import TransformerModel
    g = torch.load(os.path.join(ROOT, "tests", "golden", name), weights_only=False)
    cfg = g["config"]
    model = TransformerModel(_namespace(cfg["emb_dim"], cfg["n_layers"], cfg["n_heads"], cfg["n_words"], cfg["n_langs"]),
                             is_encoder=True, with_output=True, is_crossModal=True)
    model.load_state_dict(g["state_dict"], strict=False)
    model.cuda().train()
    batch = {k: v.cuda() for k, v in g["batch"].items()}
    batch["x_img"].requires_grad_(True)
    enc, losses, total = _step(torch, F, model, batch, cfg["sample_n"], ("mlm", "mrm", "mrfr", "rel"))
    total.backward()
    torch.cuda.synchronize()
    ref = g["joint"]
    res = dict(case="golden/" + name, enc=_rel(enc.detach().cpu(), ref["enc"]))
    for k in ("mlm", "mrm", "mrfr", "rel"):
        res["loss_" + k] = abs(float(losses[k].detach()) - ref["losses"][k]) / abs(ref["losses"][k])
    res["grad_x_img"] = _rel(batch["x_img"].grad.cpu(), ref["grad_x_img"])
    worst, worst_name = 0.0, ""
    named = dict(model.named_parameters(remove_duplicate=False))
    for k, gr in ref["grads"].items():
        if k == "pred_layer.proj.weight":
            continue
        got = named[k].grad
        if got is None:
            worst, worst_name = 1e9, k + " (missing)"
            break
        if gr.norm() < 1e-7:
            continue
        e = _rel(got.cpu(), gr)
        if e > worst:
            worst, worst_name = e, k
    res["worst_grad"], res["worst_grad_name"] = worst, worst_name
    S = enc.shape[0]
    mask = torch.arange(S)[:, None] < (g["batch"]["lengths"] + g["batch"]["lengths_img"])[None, :]
    res["pad_max"] = float(enc.detach().cpu().float()[~mask].abs().max()) if (~mask).any() else 0.0
    res["ok"] = res["enc"] < 2e-2 and worst < 5e-2 and res["pad_max"] == 0.0 and \
        max(res["loss_" + k] for k in ("mlm", "mrm", "mrfr", "rel")) < 2e-2
    _out(**res)


def case_step(B, L_=12, d=768, H=12, V=250002, heads=("rel",), dropout=0.1, steps=10):
    import torch
    import torch.nn.functional as F
    from m3p_b200.transformer import TransformerModel
    from oracle import m3p_oracle as O
    torch.manual_seed(0)
    model = TransformerModel(_namespace(d, L_, H, V, 1, dropout), is_encoder=True, with_output=True, is_crossModal=True)
    model.cuda().train()
    batch = {k: v.cuda() for k, v in O.synthetic_batch(B, 128, 100, V, sample_n=4, seed=1234).items()}
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3 + steps):
        if it == 3:
            torch.cuda.synchronize()
            t0.record()
        model.zero_grad()
        enc, losses, total = _step(torch, F, model, batch, 4, heads)
        total.backward()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    gn = float(model._flat_grad.norm())
    _out(case="step", B=B, L=L_, d=d, heads=list(heads), dropout=dropout, ms_per_step=ms, pairs_per_s=B / ms * 1e3,
         loss=float(total.detach()), grad_norm=gn, finite=math.isfinite(gn) and math.isfinite(float(total.detach())),
         mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30, ok=math.isfinite(gn))


CASES = {
    "attn_small": lambda: case_attention(2, 20, 2, False),
    "attn_128": lambda: case_attention(2, 128, 2, True),
    "attn_228": lambda: case_attention(3, 228, 2, True),
    "attn_256": lambda: case_attention(2, 256, 1, True),
    "attn_time": lambda: case_attention(64, 228, 12, False, time_it=True),
    "attn_dropout": case_attention_dropout,
    "layernorm": case_layernorm,
    "cross_entropy": case_cross_entropy,
    "golden_tiny": lambda: case_golden("c1_tiny.pt"),
    "golden_ragged": lambda: case_golden("c1_ragged_langs.pt"),
    "step_small": lambda: case_step(8, L_=2, V=5000, heads=("mlm", "mrm", "mrfr", "rel"), steps=3),
    "step_base": lambda: case_step(64),
    "step_base_multitask": lambda: case_step(64, heads=("mlm", "mrm", "mrfr", "rel")),
}

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        rc = 0
        for name in CASES:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=300)
                if r.returncode != 0:
                    _out(case=name, ok=False, error="exit code %d" % r.returncode)
                    rc = 1
            except subprocess.TimeoutExpired:
                _out(case=name, ok=False, error="timeout")
                rc = 1
        sys.exit(rc)
    CASES[which]()
