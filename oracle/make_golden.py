"""Generate golden vectors from the UNMODIFIED reference (run in the build container only).

    python oracle/make_golden.py            # writes tests/golden/*.pt and state_dict_keys.json

Imports /root/reference/M3P/src/model/transformer.py (never copied), builds TransformerModel exactly
as model/__init__.py:93 does, runs jointfwd / fwd / crossfwd / predict and the pretrain_under_step loss
assembly (restated from xtrainer.py:2285-2375, because xtrainer itself needs apex) on seeded
synthetic inputs with dropout = 0, and stores inputs, parameters, outputs, losses and gradients.
The fixtures travel to the GPU box; /root/reference does not.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("M3P_REF", "/root/reference/M3P")

from oracle import m3p_oracle as O  # noqa: E402


def ref_params(emb_dim, n_layers, n_heads, n_words, n_langs=1, dropout=0.0):
    langs = ["en", "fr", "de", "zh"][:n_langs]
    return argparse.Namespace(
        n_langs=n_langs, n_words=n_words, eos_index=2, pad_index=1,
        id2lang={i: l for i, l in enumerate(langs)}, lang2id={l: i for i, l in enumerate(langs)},
        emb_dim=emb_dim, n_heads=n_heads, n_layers=n_layers, n_dec_layers=n_layers,
        dropout=dropout, attention_dropout=dropout, sinusoidal_embeddings=False, refine_layers=1,
        attention_setting="v1", use_externel_att=False, gelu_activation=True, share_inout_emb=True, asm=False)


def build_reference(p, seed=0):
    sys.path.insert(0, REF)
    from src.model.transformer import TransformerModel
    torch.manual_seed(seed)
    m = TransformerModel(p, is_encoder=True, with_output=True, is_crossModal=True)
    # LayerNorm affine / biases away from their trivial init so the fixtures exercise them
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for n, t in m.named_parameters():
            if "LayerNorm" in n or "layer_norm" in n:
                t.add_(0.1 * torch.randn(t.shape, generator=g))
    m.train()
    return m


# the new cases keep a representative subset of the gradients (fixtures stay small): the embedding tables, the
# LayerNorms, one matrix and one bias of the first / last layer, the heads they exercise
KEEP_GRADS = ("embeddings.weight", "position_embeddings.weight", "cross_lang_embeddings.weight", "layer_norm_emb.weight",
              "layer_norm_emb.bias", "attentions.0.q_lin.weight", "attentions.0.k_lin.bias", "attentions.1.out_lin.weight",
              "layer_norm1.0.weight", "ffns.1.lin2.bias", "ffns.0.lin1.bias", "layer_norm2.1.weight", "layer_norm2.1.bias",
              "pred_layer.proj.bias", "pooled_layer.dense.weight", "pooled_layer.dense.bias", "seq_relationship.weight",
              "seq_relationship.bias", "pooled_layer2.dense.weight", "pooled_layer2.dense.bias", "seq_relationship2.weight",
              "seq_relationship2.bias", "image_embeddings.image_embeddings.bias", "image_embeddings.LayerNorm.weight",
              "image_embeddings.image_location_embeddings.weight")


def kept_grads(m):
    """(subset of gradients, names of every parameter that received a non-zero gradient)."""
    live = sorted(n for n, p_ in m.named_parameters() if p_.grad is not None and float(p_.grad.abs().max()) > 0)
    return {n: dict(m.named_parameters())[n].grad.detach().clone() for n in live if n in KEEP_GRADS}, live


def import_xtrainer():
    """The reference trainer module with `apex` stubbed (it is imported at module level, xtrainer.py:24, and absent
    here) and Tensor.cuda() made the identity: lets the generator call the trainer's own pure-torch helpers on CPU."""
    import types
    if "apex" not in sys.modules:
        apex = types.ModuleType("apex")
        apex.parallel = types.ModuleType("apex.parallel")
        apex.parallel.DistributedDataParallel = type("DistributedDataParallel", (), {})
        sys.modules["apex"], sys.modules["apex.parallel"] = apex, apex.parallel
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import src.xtrainer as X
    torch.Tensor.cuda = lambda self, *a, **k: self
    return X


def run_case(name, emb_dim, n_layers, n_heads, n_words, B, T, R, n_langs, ragged, seed):
    p = ref_params(emb_dim, n_layers, n_heads, n_words, n_langs)
    m = build_reference(p, seed)
    sample_n = 2 if B % 4 else 4
    batch = O.synthetic_batch(B, T, R, n_words, sample_n=sample_n, seed=seed + 100, ragged=ragged,
                              n_mask_text=3, n_mask_img=2)
    out = {"config": dict(emb_dim=emb_dim, n_layers=n_layers, n_heads=n_heads, n_words=n_words, n_langs=n_langs,
                          B=B, T=T, R=R, sample_n=sample_n), "batch": batch}
    x_img = batch["x_img"].clone().requires_grad_(True)

    # ---- jointfwd + the four heads, loss assembly as xtrainer.py:2285-2375 ----
    enc = m("jointfwd", x=batch["x"], lengths=batch["lengths"], x_img=x_img, lengths_img=batch["lengths_img"],
            causal=False, langs=None, image_loc=batch["image_loc"], refine_image=False)
    text_out = enc[R:]
    img_out = enc[:R].transpose(0, 1)
    y_text, pm_text = O.get_mask_(batch["x_labels"])
    mlm_scores, mlm_loss = m("predict", tensor=text_out, pred_mask=pm_text, y=y_text, get_scores=False)
    obj_scores, mrm_loss = m("predict", tensor=img_out, pred_mask=None, y=batch["obj_labels"].view(-1),
                             get_scores=False, is_obj=True)
    reg = m("predict", tensor=img_out, is_mrfr=True)
    sel = batch["obj_labels"].reshape(-1) != -1
    mrfr_loss = F.mse_loss(reg.reshape(-1, 2048)[sel], batch["ori_feats"].reshape(-1, 2048)[sel])
    rel_scores = m("predict", tensor=enc.transpose(0, 1), is_relation=True)
    ce = F.cross_entropy(rel_scores.view(-1, sample_n), batch["pos_labels"])
    bce = F.binary_cross_entropy_with_logits(rel_scores.view(-1),
                                             F.one_hot(batch["pos_labels"], sample_n).float().view(-1))
    rel_loss = ce + bce
    total = mlm_loss + mrm_loss + mrfr_loss + rel_loss
    total.backward()
    out["joint"] = dict(enc=enc.detach(), mlm_scores=mlm_scores.detach(), obj_scores=obj_scores.detach(),
                        mrfr=reg.detach(), rel_scores=rel_scores.detach(),
                        losses=dict(mlm=mlm_loss.item(), mrm=mrm_loss.item(), mrfr=mrfr_loss.item(),
                                    rel=rel_loss.item(), total=total.item()),
                        grad_x_img=x_img.grad.detach().clone())
    grads = {n: t.grad.detach().clone() for n, t in m.named_parameters() if t.grad is not None}
    out["joint"]["grads"] = grads
    out["no_grad_params"] = sorted(n for n, t in m.named_parameters() if t.grad is None)
    m.zero_grad()

    # ---- text streams: fwd, crossfwd (with and without langs) ----
    with torch.no_grad():
        out["fwd_text"] = m("fwd", x=batch["x"], lengths=batch["lengths"], causal=False)
        out["crossfwd_text"] = m("crossfwd", x=batch["x"], lengths=batch["lengths"], causal=False, stream_="text")
        if n_langs > 1:
            langs = torch.randint(0, n_langs, batch["x"].shape, generator=torch.Generator().manual_seed(seed))
            out["langs"] = langs
            out["crossfwd_text_langs"] = m("crossfwd", x=batch["x"], lengths=batch["lengths"], causal=False,
                                           stream_="text", langs=langs)
        out["fwd_image"] = m("fwd", x=batch["x_img"], lengths=batch["lengths_img"], causal=False, cross_modal=True,
                             image_loc=batch["image_loc"])
        out["crossfwd_img"] = m("crossfwd", x=batch["x_img"], lengths=batch["lengths_img"], causal=False, stream_="img",
                                langs=None, cross_modal=True, image_loc=batch["image_loc"])

    # ---- text-stream BACKWARD (mlm_step, xtrainer.py:734-770): fwd / crossfwd (+ positions, + langs) -> MLM head ->
    #      loss.backward(); every gradient incl. cross_lang_embeddings (transformer.py:1056-1057) ----
    gpos = torch.Generator().manual_seed(seed + 5)
    positions = torch.stack([torch.randperm(T, generator=gpos) for _ in range(B)], dim=1)   # (T, B), reset positions
    out["positions"] = positions
    text_cases = {"fwd": dict(mode="fwd"), "crossfwd": dict(mode="crossfwd", stream_="text"),
                  "crossfwd_positions": dict(mode="crossfwd", stream_="text", positions=positions)}
    if n_langs > 1:
        text_cases["crossfwd_langs"] = dict(mode="crossfwd", stream_="text", langs=out["langs"])
        text_cases["crossfwd_langs_positions"] = dict(mode="crossfwd", stream_="text", langs=out["langs"],
                                                      positions=positions)
    out["text_bwd"] = {}
    wsum = torch.randn(T, B, emb_dim, generator=torch.Generator().manual_seed(seed + 6))
    out["text_bwd_weight"] = wsum
    for cname, kw in text_cases.items():
        kw = dict(kw)
        mode = kw.pop("mode")
        m.zero_grad()
        t = m(mode, x=batch["x"], lengths=batch["lengths"], causal=False, **kw)
        _, loss = m("predict", tensor=t, pred_mask=pm_text, y=y_text, get_scores=False)
        tot = loss + 0.01 * (t * wsum).sum()           # the weighted sum reaches every row, not only the masked ones
        tot.backward()
        gr, live = kept_grads(m)
        out["text_bwd"][cname] = dict(out=t.detach().clone(), loss=loss.item(), total=tot.item(), grads=gr, live=live)
    m.zero_grad()

    # ---- CLCM second pass (pretrain_under_step, i2t branch, xtrainer.py:2379-2393): jointfwd on the code-switched
    #      caption x2 with the SAME regions -> predict(is_clcm=True) -> BCE against clcm_labels ----
    b2 = O.synthetic_batch(B, T, R, n_words, sample_n=sample_n, seed=seed + 200, ragged=ragged, n_mask_text=3, n_mask_img=2)
    clcm_labels = torch.randint(0, 2, (B,), generator=torch.Generator().manual_seed(seed + 7))
    enc2 = m("jointfwd", x=b2["x"], lengths=b2["lengths"], x_img=batch["x_img"], lengths_img=batch["lengths_img"],
             causal=False, langs=None, image_loc=batch["image_loc"], refine_image=False)
    scores2 = m("predict", tensor=enc2.transpose(0, 1), is_clcm=True)
    bce2 = F.binary_cross_entropy_with_logits(scores2.view(-1), clcm_labels.view(-1).float())
    bce2.backward()
    gr, live = kept_grads(m)
    out["clcm"] = dict(x2=b2["x"], lengths2=b2["lengths"], clcm_labels=clcm_labels, scores=scores2.detach().clone(),
                       loss=bce2.item(), grads=gr, live=live)
    m.zero_grad()

    # ---- FreeLB (freelb_t2i_step / freelb_i2t_step, xtrainer.py:2021-2223) with the reference's own delta helpers
    #      (deal_freelb_delta / update_freelb_delta / deal_image_freelb_delta / update_image_freelb_delta,
    #      :2700-2851) called through a stub `self`; no optimizer step between the ascent steps (the AMP
    #      accumulate branch of free_optimize, :2778-2791), so gradients of the three steps accumulate ----
    X = import_xtrainer()
    stub = object.__new__(X.XTrainer)
    torch.manual_seed(seed + 8)
    x1 = batch["x"]
    embeds_init, delta = X.XTrainer.deal_freelb_delta(stub, m, x1.transpose(0, 1), batch["lengths"])
    image_delta = X.XTrainer.deal_image_freelb_delta(stub, batch["x_img"])
    fl = dict(delta0=delta.clone(), image_delta0=image_delta.clone(), steps=[])
    adv_steps = 3
    for astep in range(adv_steps):
        delta.requires_grad_()
        text_imb = delta + embeds_init
        image_delta.requires_grad_()
        img_imb = batch["x_img"] + image_delta
        enc_f = m("jointfwd", x=x1, lengths=batch["lengths"], x_img=img_imb, lengths_img=batch["lengths_img"], causal=False,
                  langs=None, image_loc=batch["image_loc"], refine_image=False, text_embed=text_imb)
        rs = m("predict", tensor=enc_f.transpose(0, 1), is_relation=True)
        ce_f = F.cross_entropy(rs.view(-1, sample_n), batch["pos_labels"])
        bce_f = F.binary_cross_entropy_with_logits(rs.view(-1), F.one_hot(batch["pos_labels"], sample_n).float().view(-1))
        loss_f = (ce_f + bce_f) / (1.0 * adv_steps)
        loss_f.backward()
        rec = dict(loss=loss_f.item(), delta_grad=delta.grad.detach().clone(), image_delta_grad=image_delta.grad.detach().clone())
        if astep < adv_steps - 1:
            embeds_init, delta = X.XTrainer.update_freelb_delta(stub, m, delta, embeds_init, x1.transpose(0, 1))
            image_delta = X.XTrainer.update_image_freelb_delta(stub, batch["x_img"], image_delta)
            rec.update(delta_next=delta.clone(), image_delta_next=image_delta.clone())
        fl["steps"].append(rec)
    fl["grads"], fl["live"] = kept_grads(m)
    out["freelb"] = fl
    m.zero_grad()
    sd = m.state_dict()
    out["state_dict"] = {k: v.clone() for k, v in sd.items()
                         if not k.startswith(("refine_embeddings", "cross_alignment", "encoder_attn", "layer_norm15",
                                              "latent_transforms", "original_transforms",
                                              "image_embeddings.image_distbution_embeddings",
                                              "pred_layer.proj.weight"))}  # tied to embeddings.weight
    keys = {k: list(v.shape) for k, v in sd.items()}
    return out, keys


def run_optimizer_case(seed=3, n=1003, steps=6, max_norm=5.0):
    """Trainer.optimize's clip + step (xtrainer.py:222-228) with the reference's own AdamInverseSqrtWithWarmup
    (optim.py:89-139) on two parameter tensors and seeded gradients; `get_optimizer` itself cannot run on
    Python >= 3.11 (inspect.getargspec, optim.py:264), so the class is constructed directly."""
    import warnings
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from src.optim import AdamInverseSqrtWithWarmup
    g = torch.Generator().manual_seed(seed)
    params = [torch.nn.Parameter(torch.randn(n, generator=g)), torch.nn.Parameter(torch.randn(17, 5, generator=g))]
    kw = dict(lr=1e-2, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.01, warmup_updates=3, warmup_init_lr=1e-4)
    opt = AdamInverseSqrtWithWarmup(params, **kw)
    out = {"kw": kw, "max_norm": max_norm, "p0": [p.detach().clone() for p in params], "grads": [], "params": [],
           "lrs": [], "norms": []}
    for s in range(steps):
        scale = 40.0 if s % 2 == 0 else 0.05  # alternate clipped / unclipped steps
        grads = [torch.randn(p.shape, generator=g) * scale for p in params]
        for p, gr in zip(params, grads):
            p.grad = gr.clone()
        norm = torch.nn.utils.clip_grad_norm_(params, max_norm)
        out["lrs"].append(opt.param_groups[0]["lr"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            opt.step()
        out["grads"].append(grads)
        out["norms"].append(float(norm))
        out["params"].append([p.detach().clone() for p in params])
    # learning-rate schedules of the reference classes (pure Python: get_lr_for_step), optim.py:129-133, 184-201
    from src.optim import AdamCosineWithWarmup
    dummy = [torch.nn.Parameter(torch.zeros(1))]
    inv = AdamInverseSqrtWithWarmup(dummy, lr=1e-4, warmup_updates=4000, warmup_init_lr=1e-7)
    cos1 = AdamCosineWithWarmup(dummy, lr=1e-4, warmup_updates=100, warmup_init_lr=1e-7, min_lr=1e-9, init_period=500,
                                period_mult=1, lr_shrink=0.75)
    cos2 = AdamCosineWithWarmup(dummy, lr=1e-4, warmup_updates=100, warmup_init_lr=1e-7, min_lr=1e-9, init_period=300,
                                period_mult=2, lr_shrink=0.5)
    steps_ = [0, 1, 50, 99, 100, 101, 399, 400, 401, 777, 1000, 3999, 4000, 4001, 10000, 123456]
    out["schedules"] = {"steps": steps_, "inverse_sqrt": [inv.get_lr_for_step(n) for n in steps_],
                        "cosine_mult1": [cos1.get_lr_for_step(n) for n in steps_],
                        "cosine_mult2": [cos2.get_lr_for_step(n) for n in steps_]}
    return out


if __name__ == "__main__":
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    torch.save(run_optimizer_case(), os.path.join(gold, "adam_inverse_sqrt.pt"))
    print("adam_inverse_sqrt.pt", os.path.getsize(os.path.join(gold, "adam_inverse_sqrt.pt")))
    if "--optimizer-only" in sys.argv:
        sys.exit(0)
    # C1 of BASELINE.json: 2 layers / 128 hidden, 16 text + 4 region tokens, batch 2
    c1, keys = run_case("c1", 128, 2, 2, 1000, B=2, T=16, R=4, n_langs=1, ragged=False, seed=0)
    torch.save(c1, os.path.join(gold, "c1_tiny.pt"))
    # ragged lengths + several languages (cross_lang_embeddings) + odd sizes
    c1r, keys_r = run_case("c1_ragged", 128, 2, 2, 600, B=4, T=12, R=5, n_langs=3, ragged=True, seed=7)
    torch.save(c1r, os.path.join(gold, "c1_ragged_langs.pt"))
    with open(os.path.join(gold, "state_dict_keys.json"), "w") as f:
        json.dump({"c1_tiny": keys, "c1_ragged_langs": keys_r}, f, indent=0, sort_keys=True)
    for n in ("c1_tiny.pt", "c1_ragged_langs.pt", "state_dict_keys.json"):
        print(n, os.path.getsize(os.path.join(gold, n)))
