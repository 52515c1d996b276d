"""Pin the CPU oracle (oracle/m3p_oracle.py) against outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by oracle/make_golden.py, which imports
/root/reference/M3P/src/model/transformer.py.  Everything here runs on CPU in fp32.
"""
import os

import pytest
import torch

from oracle import m3p_oracle as O

CASES = ["c1_tiny.pt", "c1_ragged_langs.pt"]


def _load(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name), weights_only=False)
    sd = dict(g["state_dict"])
    sd["pred_layer.proj.weight"] = sd["embeddings.weight"]  # tied (transformer.py:728-729)
    return g, sd


def _close(a, b, tol=2e-5):
    denom = b.norm().item() + 1e-12
    return (a - b).norm().item() / denom < tol


@pytest.mark.parametrize("name", CASES)
def test_jointfwd_heads_and_grads_match_reference(golden_dir, name):
    g, sd = _load(golden_dir, name)
    cfg, batch = g["config"], g["batch"]
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "pred_layer.proj.weight"}
    leaf["pred_layer.proj.weight"] = leaf["embeddings.weight"]
    b = dict(batch)
    b["x_img"] = batch["x_img"].clone().requires_grad_(True)
    enc, losses, total = O.pretrain_step_losses(leaf, cfg["n_layers"], cfg["n_heads"], b, cfg["sample_n"])
    total.backward()
    ref = g["joint"]
    assert _close(enc.detach(), ref["enc"])
    for k in ("mlm", "mrm", "mrfr", "rel"):
        assert abs(losses[k].item() - ref["losses"][k]) < 1e-5 * max(1.0, abs(ref["losses"][k])), k
    assert abs(total.item() - ref["losses"]["total"]) < 1e-5 * abs(ref["losses"]["total"])
    assert _close(b["x_img"].grad, ref["grad_x_img"], 1e-4)
    # padded rows of the encoder output are exactly zero (SURVEY appendix A invariants)
    S = enc.shape[0]
    mask = torch.arange(S)[:, None] < (batch["lengths"] + batch["lengths_img"])[None, :]
    assert enc.detach()[~mask].abs().max().item() == 0.0 if (~mask).any() else True
    n_checked = 0
    for k, gref in ref["grads"].items():
        if k == "pred_layer.proj.weight" or k not in leaf:
            continue
        got = leaf[k].grad
        assert got is not None, k
        assert _close(got, gref, 1e-4), k
        n_checked += 1
    assert n_checked > 40
    # parameters the reference leaves without gradient on this path stay untouched here too
    for k in g["no_grad_params"]:
        if k in leaf:
            assert leaf[k].grad is None, k


@pytest.mark.parametrize("name", CASES)
def test_head_scores_match_reference(golden_dir, name):
    g, sd = _load(golden_dir, name)
    cfg, batch = g["config"], g["batch"]
    R = cfg["R"]
    enc = g["joint"]["enc"]
    y_text, pm = O.get_mask_(batch["x_labels"])
    scores, _ = O.predict_mlm(sd, enc[R:], pm, y_text)
    assert _close(scores, g["joint"]["mlm_scores"])
    oscores, _ = O.predict_obj(sd, enc[:R].transpose(0, 1), batch["obj_labels"].reshape(-1))
    assert _close(oscores, g["joint"]["obj_scores"])
    assert _close(O.predict_mrfr(sd, enc[:R].transpose(0, 1)), g["joint"]["mrfr"])
    assert _close(O.predict_relation(sd, enc.transpose(0, 1)), g["joint"]["rel_scores"])


@pytest.mark.parametrize("name", CASES)
def test_text_and_image_streams_match_reference(golden_dir, name):
    g, sd = _load(golden_dir, name)
    cfg, batch = g["config"], g["batch"]
    L, H = cfg["n_layers"], cfg["n_heads"]
    with torch.no_grad():
        assert _close(O.fwd_text(sd, L, H, batch["x"], batch["lengths"]), g["fwd_text"])
        assert _close(O.crossfwd_text(sd, L, H, batch["x"], batch["lengths"]), g["crossfwd_text"])
        if "crossfwd_text_langs" in g:
            got = O.crossfwd_text(sd, L, H, batch["x"], batch["lengths"], langs=g["langs"])
            assert _close(got, g["crossfwd_text_langs"])
            assert not _close(got, g["crossfwd_text"], 1e-3)  # langs matter for crossfwd, not for fwd
        assert _close(O.fwd_image(sd, L, H, batch["x_img"], batch["lengths_img"], batch["image_loc"]), g["fwd_image"])


def test_get_masks_matches_reference_semantics():
    lengths = torch.tensor([3, 0, 5])
    mask, attn = O.get_masks(5, lengths)
    assert mask.tolist() == [[True] * 3 + [False] * 2, [False] * 5, [True] * 5]
    assert attn is mask


def test_optimizer_oracle_matches_reference_adam(golden_dir):
    """clip_grad_norm_ + AdamInverseSqrtWithWarmup (xtrainer.py:222-228, optim.py:45-139): the oracle's
    restatement reproduces the reference's parameters after every one of 6 steps (clipped and unclipped)."""
    g = torch.load(os.path.join(golden_dir, "adam_inverse_sqrt.pt"), weights_only=False)
    kw = g["kw"]
    params = [p.clone() for p in g["p0"]]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    lr = kw["warmup_init_lr"]
    for s, (grads, want, want_lr, want_norm) in enumerate(zip(g["grads"], g["params"], g["lrs"], g["norms"])):
        assert abs(lr - want_lr) <= 1e-12 + 1e-9 * abs(want_lr)
        coef, total = O.clip_coef(grads, g["max_norm"])
        assert abs(float(total) - want_norm) < 1e-4 * want_norm
        for p, gr, mi, vi, w in zip(params, grads, m, v, want):
            O.adam_step(p, gr * coef, mi, vi, s + 1, lr, kw["betas"], kw["eps"], kw["weight_decay"])
            assert _close(p, w, 1e-6)
        lr = O.lr_inverse_sqrt(s + 1, kw["lr"], kw["warmup_updates"], kw["warmup_init_lr"])


@pytest.mark.parametrize("name", CASES)
def test_rounding_matched_mode_stays_close_to_the_reference(golden_dir, name):
    """oracle.rounding_matched() (bf16 rounding at the kernels' rounding points, straight-through gradients) is the
    same algorithm: on the reference's own fixtures it stays within bf16 noise of the fp32 reference outputs and
    gradients, and leaves padded rows exactly zero."""
    g, sd = _load(golden_dir, name)
    cfg, batch = g["config"], g["batch"]
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "pred_layer.proj.weight"}
    leaf["pred_layer.proj.weight"] = leaf["embeddings.weight"]
    with O.rounding_matched():
        enc, losses, total = O.pretrain_step_losses(leaf, cfg["n_layers"], cfg["n_heads"], batch, cfg["sample_n"])
    total.backward()
    ref = g["joint"]
    assert _close(enc.detach(), ref["enc"], 2e-2)
    for k in ("mlm", "mrm", "mrfr", "rel"):
        assert abs(float(losses[k]) - float(ref["losses"][k])) < 2e-2 * abs(float(ref["losses"][k])), k
    assert _close(leaf["ffns.0.lin1.weight"].grad, ref["grads"]["ffns.0.lin1.weight"], 5e-2)
    S = enc.shape[0]
    lengths = batch["lengths"] + batch["lengths_img"]
    pad = torch.arange(S)[:, None] >= lengths[None, :]
    assert float(enc.detach()[pad].abs().max()) == 0.0 if bool(pad.any()) else True


# ---------------------------------------------------------------------------------------------------
# text-stream backward, CLCM second pass, FreeLB: fixtures generated from the reference's own code
# ---------------------------------------------------------------------------------------------------
def _leaf(sd):
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k != "pred_layer.proj.weight"}
    leaf["pred_layer.proj.weight"] = leaf["embeddings.weight"]
    return leaf


@pytest.mark.parametrize("name", CASES)
def test_text_stream_backward_matches_reference(golden_dir, name):
    """mlm_step's path (xtrainer.py:734-770): fwd / crossfwd (+ reset positions, + langs) -> MLM head -> backward;
    outputs, losses and the gradients of the embedding tables (cross_lang_embeddings included), layer_norm_emb and
    the layers against the reference's."""
    g, sd = _load(golden_dir, name)
    cfg, batch = g["config"], g["batch"]
    L, H = cfg["n_layers"], cfg["n_heads"]
    y_text, pm = O.get_mask_(batch["x_labels"])
    for cname, ref in g["text_bwd"].items():
        leaf = _leaf(sd)
        kw = dict(positions=g["positions"] if "positions" in cname else None)
        if cname == "fwd":
            t = O.fwd_text(leaf, L, H, batch["x"], batch["lengths"])
        else:
            t = O.crossfwd_text(leaf, L, H, batch["x"], batch["lengths"], langs=g["langs"] if "langs" in cname else None, **kw)
        _, loss = O.predict_mlm(leaf, t, pm, y_text)
        tot = loss + 0.01 * (t * g["text_bwd_weight"]).sum()
        tot.backward()
        assert _close(t.detach(), ref["out"]), cname
        assert abs(loss.item() - ref["loss"]) < 1e-5 * abs(ref["loss"]), cname
        live = sorted(k for k, v in leaf.items() if v.grad is not None and float(v.grad.abs().max()) > 0
                      and k != "pred_layer.proj.weight")
        assert live == [k for k in ref["live"] if k != "pred_layer.proj.weight"], cname
        for k, gr in ref["grads"].items():
            assert _close(leaf[k].grad, gr, 1e-4), (cname, k)
        assert ("cross_lang_embeddings.weight" in ref["grads"]) == ("langs" in cname)


@pytest.mark.parametrize("name", CASES)
def test_clcm_second_pass_matches_reference(golden_dir, name):
    g, sd = _load(golden_dir, name)
    cfg, batch, c = g["config"], g["batch"], g["clcm"]
    leaf = _leaf(sd)
    enc2 = O.jointfwd(leaf, cfg["n_layers"], cfg["n_heads"], c["x2"], c["lengths2"], batch["x_img"], batch["lengths_img"],
                      batch["image_loc"])
    scores = O.predict_relation(leaf, enc2.transpose(0, 1), clcm=True)
    loss = O.clcm_loss(scores, c["clcm_labels"])
    loss.backward()
    assert _close(scores.detach(), c["scores"]) and abs(loss.item() - c["loss"]) < 1e-5 * abs(c["loss"])
    for k, gr in c["grads"].items():
        assert _close(leaf[k].grad, gr, 1e-4), k
    assert leaf["pooled_layer.dense.weight"].grad is None  # the ITM pooler is not on this pass


@pytest.mark.parametrize("name", CASES)
def test_freelb_helpers_and_trajectory_match_reference(golden_dir, name):
    """The FreeLB perturbation helpers (train_step.init_adv_delta / ascend_adv_delta: host-side tensor prep, run here
    on CPU) reproduce the reference's deal_* / update_* functions bit for bit given the same RNG state, and the
    oracle reproduces the three ascent steps (losses, perturbation gradients, accumulated parameter gradients —
    the token table included, which is trained through embeds_init)."""
    from m3p_b200.train_step import ascend_adv_delta, init_adv_delta
    g, sd = _load(golden_dir, name)
    cfg, batch, fl = g["config"], g["batch"], g["freelb"]
    seed = {"c1_tiny.pt": 0, "c1_ragged_langs.pt": 7}[name]
    B, T, R, d = cfg["B"], cfg["T"], cfg["R"], cfg["emb_dim"]
    torch.manual_seed(seed + 8)
    d0 = init_adv_delta(torch.zeros(B, T, d), batch["lengths"] * d)
    i0 = init_adv_delta(batch["x_img"], torch.full((R,), 2048.0))
    assert torch.equal(d0, fl["delta0"]) and torch.equal(i0, fl["image_delta0"])
    leaf = _leaf(sd)
    delta_t, delta_i = d0, i0
    for s, rec in enumerate(fl["steps"]):
        delta_t = delta_t.detach().requires_grad_(True)
        delta_i = delta_i.detach().requires_grad_(True)
        emb = torch.nn.functional.embedding(batch["x"].transpose(0, 1), leaf["embeddings.weight"], padding_idx=1)
        enc = O.jointfwd(leaf, cfg["n_layers"], cfg["n_heads"], batch["x"], batch["lengths"], batch["x_img"] + delta_i,
                         batch["lengths_img"], batch["image_loc"], text_embed=emb + delta_t)
        loss = O.relation_loss(O.predict_relation(leaf, enc.transpose(0, 1)), batch["pos_labels"], cfg["sample_n"]) / 3.0
        loss.backward()
        assert abs(loss.item() - rec["loss"]) < 1e-5 * abs(rec["loss"]), s
        assert _close(delta_t.grad, rec["delta_grad"], 1e-4) and _close(delta_i.grad, rec["image_delta_grad"], 1e-4)
        if "delta_next" in rec:
            assert torch.equal(ascend_adv_delta(delta_t, rec["delta_grad"]), rec["delta_next"])
            assert torch.equal(ascend_adv_delta(delta_i, rec["image_delta_grad"]), rec["image_delta_next"])
            delta_t, delta_i = rec["delta_next"], rec["image_delta_next"]
    for k, gr in fl["grads"].items():
        assert _close(leaf[k].grad, gr, 1e-4), k
    assert "embeddings.weight" in fl["grads"]
