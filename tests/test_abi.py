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
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(m3p_gemm_args), sizeof(m3p_attn_args), sizeof(m3p_ln_bwd_args),
         sizeof(m3p_embed_args), sizeof(m3p_embed_bwd_args), sizeof(m3p_adam_args));
  return 0;
}'''
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "p.c")
        open(c, "w").write(probe)
        exe = os.path.join(td, "p")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    got = [ctypes.sizeof(t) for t in (L.GemmArgs, L.AttnArgs, L.LnBwdArgs, L.EmbedArgs, L.EmbedBwdArgs, L.AdamArgs)]
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
