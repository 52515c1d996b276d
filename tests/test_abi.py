"""The C-ABI library builds, loads without a GPU, and exports exactly what include/m3p_b200.h declares
(no compute calls here: those need a B200 and live in the -m gpu tests)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "m3p_b200.h")


def _declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"M3P_API\s+[\w\s\*]+?\b(m3p_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from m3p_b200 import build, lib as L
    build.build()
    return L.load()


def test_header_declares_the_path():
    syms = _declared_symbols()
    for must in ("m3p_gemm_bf16", "m3p_attention_fwd", "m3p_attention_bwd", "m3p_layernorm_fwd", "m3p_layernorm_bwd",
                 "m3p_embed_fwd", "m3p_embed_bwd_route", "m3p_cross_entropy_fwd", "m3p_cross_entropy_bwd"):
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    for s in _declared_symbols():
        assert hasattr(lib, s), "libm3p_sm100.so does not export %s" % s


def test_ctypes_table_covers_the_header(lib):
    from m3p_b200 import lib as L
    declared = set(_declared_symbols()) - {"m3p_version", "m3p_last_error"}
    assert declared == set(L.PROTOTYPES), declared ^ set(L.PROTOTYPES)


def test_version_and_error_string(lib):
    assert lib.m3p_version() >= 100
    assert isinstance(lib.m3p_last_error(), bytes)


def test_struct_sizes_match_the_header(lib):
    """ctypes mirrors of the argument structs have the C layout (checked against a gcc-compiled probe)."""
    import subprocess
    import tempfile
    from m3p_b200 import lib as L
    probe = r'''
#include <stdio.h>
#include "m3p_b200.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu %zu\n", sizeof(m3p_gemm_args), sizeof(m3p_attn_args), sizeof(m3p_ln_bwd_args),
         sizeof(m3p_embed_args), sizeof(m3p_embed_bwd_args), sizeof(m3p_adam_args), sizeof(m3p_ln_fwd_args));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "p.c")
        open(c, "w").write(probe)
        exe = os.path.join(td, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    got = [ctypes.sizeof(t) for t in (L.GemmArgs, L.AttnArgs, L.LnBwdArgs, L.EmbedArgs, L.EmbedBwdArgs, L.AdamArgs, L.LnFwdArgs)]
    assert got == sizes


def test_no_gpu_means_loud_failure(lib):
    """Without a B200 the product path raises instead of computing something else."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from m3p_b200 import lib as L, ops
    with pytest.raises(L.M3PError):
        ops.device_check()


def test_train_x_parser_mirrors_reference_flags():
    """m3p_b200.train_x accepts the reference's flag names (train_x.py:29-391) for the subset the path serves."""
    from m3p_b200 import train_x
    p = train_x.get_parser().parse_args(
        ["--emb_dim", "1024", "--n_layers", "24", "--n_heads", "16", "--gelu_activation", "true", "--batch_size", "24",
         "--sample_n", "4", "--bptt", "128", "--max_region_num", "100", "--amp", "1", "--fp16", "true",
         "--accumulate_gradients", "4", "--optimizer", "adam_inverse_sqrt,beta1=0.9,beta2=0.98,lr=0.00005",
         "--clip_grad_norm", "5", "--cross_rel_steps", "coco-img", "--lambda_rel", "1"])
    ns = train_x.model_namespace(p)
    assert (ns.emb_dim, ns.n_layers, ns.n_heads, ns.n_words, ns.pad_index, ns.eos_index) == (1024, 24, 16, 250002, 1, 2)
    assert p.accumulate_gradients == 4 and p.clip_grad_norm == 5


# ---------------------------------------------------------------------------------------------------
# host logic that needs no GPU: optimizer DSL and schedules, batch masking, retrieval bookkeeping
# ---------------------------------------------------------------------------------------------------
def test_optimizer_dsl_and_schedules_match_the_reference(golden_dir):
    """get_optimizer's string DSL (optim.py:211-270) and the LR schedules of AdamInverseSqrtWithWarmup /
    AdamCosineWithWarmup (:129-133, :184-201) against values produced by the reference classes."""
    import torch
    from m3p_b200 import optim
    g = torch.load(os.path.join(golden_dir, "adam_inverse_sqrt.pt"), weights_only=False)["schedules"]
    p = [torch.nn.Parameter(torch.zeros(4))]
    o = optim.get_optimizer(p, "adam_inverse_sqrt,beta1=0.9,beta2=0.98,lr=0.0001", clip_grad_norm=5.0)
    assert isinstance(o, optim.AdamInverseSqrtWithWarmup) and o.param_groups[0]["betas"] == (0.9, 0.98)
    assert o.clip_grad_norm == 5.0 and o.param_groups[0]["lr"] == 1e-7 and o.param_groups[0]["num_updates"] == 0
    for n, want in zip(g["steps"], g["inverse_sqrt"]):
        assert abs(o.get_lr_for_step(n) - want) <= 1e-12 * max(1.0, abs(want)) + 1e-18
    c1 = optim.get_optimizer(p, "adam_cosine,lr=0.0001,warmup_updates=100,min_lr=0.000000001,init_period=500,lr_shrink=0.75")
    c2 = optim.AdamCosineWithWarmup(p, lr=1e-4, warmup_updates=100, warmup_init_lr=1e-7, min_lr=1e-9, init_period=300,
                                    period_mult=2, lr_shrink=0.5)
    for n, w1, w2 in zip(g["steps"], g["cosine_mult1"], g["cosine_mult2"]):
        assert abs(c1.get_lr_for_step(n) - w1) <= 1e-9 * abs(w1) + 1e-18
        assert abs(c2.get_lr_for_step(n) - w2) <= 1e-9 * abs(w2) + 1e-18
    a = optim.get_optimizer(p, "adam,lr=0.001")
    assert type(a) is optim.Adam and a.param_groups[0]["lr"] == 0.001
    with pytest.raises(NotImplementedError):
        optim.get_optimizer(p, "sgd,lr=0.1")          # torch.optim pass-throughs are outside the B200 path
    with pytest.raises(Exception):
        optim.get_optimizer(p, "adam,momentum=0.9")   # unexpected parameter, as the reference rejects it
    with pytest.raises(Exception):
        optim.get_optimizer(p, "nadam")


def test_mask_out_follows_the_reference_rules():
    """Trainer.mask_out (xtrainer.py:385-434): never position 0 or padding, count rounded down to a multiple of 8,
    targets are the original tokens, replacements are <mask> / same / random in roughly 80/10/10."""
    import torch
    from m3p_b200.train_step import mask_out, synthetic_batch
    b = synthetic_batch(32, 64, 2, 1000, sample_n=4, seed=1, ragged=True)
    g = torch.Generator().manual_seed(0)
    x, y, pm = mask_out(b["x"], b["lengths"], 1000, word_pred=0.15, generator=g)
    n = int(pm.sum())
    assert n > 0 and n % 8 == 0 and not bool(pm[0].any()) and not bool((b["x"][pm] == 1).any())
    assert torch.equal(y, b["x"][pm]) and torch.equal(x[~pm], b["x"][~pm])
    frac_mask = float((x[pm] == 999).float().mean())
    frac_same = float((x[pm] == y).float().mean())
    assert 0.65 < frac_mask < 0.92 and 0.03 < frac_same < 0.25


def test_recall_at_k_counts_like_the_reference():
    import torch
    from m3p_b200.evaluate import recall_at_k
    sc = torch.tensor([[0.9, 0.1, 0.2, 0.0], [0.8, 0.7, 0.1, 0.6]])
    lab = torch.tensor([[1, 0, 0, 0], [0, 0, 0, 1]])
    i2t, t2i = recall_at_k(sc, lab, ks=(1, 2, 3))
    assert i2t == {1: 0.5, 2: 0.5, 3: 1.0} and t2i == {1: 0.5, 2: 0.5, 3: 0.5}


def test_state_dict_names_and_shapes_equal_the_reference(golden_dir):
    """Boundary (b): a reference checkpoint loads unchanged — TransformerModel.state_dict() has exactly the keys
    and shapes the reference module has (tests/golden/state_dict_keys.json, written by oracle/make_golden.py from
    the reference class), for both fixture configurations."""
    import argparse
    import json
    from m3p_b200.transformer import TransformerModel
    want = json.load(open(os.path.join(golden_dir, "state_dict_keys.json")))
    for case, (d, n_words, n_langs) in (("c1_tiny", (128, 1000, 1)), ("c1_ragged_langs", (128, 600, 3))):
        langs = ["en", "fr", "de", "zh"][:n_langs]
        ns = argparse.Namespace(
            n_langs=n_langs, n_words=n_words, eos_index=2, pad_index=1, id2lang=dict(enumerate(langs)),
            lang2id={l: i for i, l in enumerate(langs)}, emb_dim=d, n_heads=2, n_layers=2, n_dec_layers=2, dropout=0.0,
            attention_dropout=0.0, sinusoidal_embeddings=False, refine_layers=1, attention_setting="v1",
            use_externel_att=False, gelu_activation=True, share_inout_emb=True, asm=False)
        m = TransformerModel(ns, is_encoder=True, with_output=True, is_crossModal=True)
        got = {k: list(v.shape) for k, v in m.state_dict().items()}
        assert got == want[case], sorted(set(got) ^ set(want[case]))


def test_get_masks_on_cpu():
    """E0 on CPU: the product's get_masks equals the reference's non-causal masks (transformer.py:59-78)."""
    import torch
    from m3p_b200.transformer import get_masks
    mask, attn = get_masks(5, torch.tensor([3, 0, 5]), False)
    assert mask.tolist() == [[True] * 3 + [False] * 2, [False] * 5, [True] * 5] and attn is mask
    with pytest.raises(AssertionError):
        get_masks(4, torch.tensor([5]), False)
