"""Statistical check (CPU, numpy) of the counter-based dropout generators the kernels use — restated from
m3p_b200/csrc/common.cuh (`drop_hash`: one 32-bit hash per element pair, 16-bit threshold per element) and
m3p_b200/csrc/attention.cu (`attn_drop_next`: one hash per 32-key chunk seeding an LCG advanced once per key pair).
Dropout cannot be bit-compared with PyTorch's Philox stream, so what is pinned here is what a
dropout mask must satisfy: the keep rate, independence between neighbours / chunks, and the drops-per-chunk
distribution of Binomial(32, p)."""
from math import comb

import numpy as np

M = np.uint64(0xFFFFFFFF)


def drop_hash(idx, lo, hi):
    x = ((idx ^ lo) * np.uint64(0x9E3779B1) + hi) & M
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x85EBCA6B)) & M
    x ^= x >> np.uint64(13)
    return x


def _corr(a, b):
    return float((a * b).mean() / np.sqrt((a * a).mean() * (b * b).mean()))


def _check(keep, p):
    n = keep.size
    sigma = np.sqrt(p * (1 - p) / n)
    assert abs(keep.mean() - (1 - p)) < 4 * sigma
    assert np.abs(keep.mean(0) - (1 - p)).max() < 5 * np.sqrt(p * (1 - p) / keep.shape[0])
    k = keep.astype(np.float64) - keep.mean()
    bound = 5.0 / np.sqrt(n)
    assert abs(_corr(k[:, 0::2], k[:, 1::2])) < bound      # the two halves of one 32-bit draw
    assert abs(_corr(k[:, :-2], k[:, 2:])) < bound         # neighbouring draws
    assert abs(_corr(k[:-1], k[1:])) < bound               # same position, neighbouring chunks / rows
    drops = (~keep).sum(1)
    width = keep.shape[1]
    for i in range(8):
        want = comb(width, i) * p ** i * (1 - p) ** (width - i)
        got = float((drops == i).mean())
        assert abs(got - want) < 5 * np.sqrt(want * (1 - want) / keep.shape[0]) + 1e-4, (i, got, want)


def test_elementwise_generator():
    """drop_hash as the GEMM epilogue / LayerNorm / embedding kernels use it: pair index = element index >> 1."""
    p, n, width = 0.1, 100_000, 32
    thr = np.uint64(int(p * 65536 + 0.5))
    lo, hi = np.uint64(0x5DEECE66), np.uint64(0x0000000D)
    pair = np.arange(n * width // 2, dtype=np.uint64)
    h = drop_hash(pair, lo, hi)
    keep = np.empty((n * width,), dtype=bool)
    keep[0::2] = (h & np.uint64(0xFFFF)) >= thr
    keep[1::2] = (h >> np.uint64(16)) >= thr
    _check(keep.reshape(n, width), float(thr) / 65536)


def test_attention_chunk_generator():
    """attn_drop_next: chunk seed = drop_hash(pair index of the chunk's first key), then 16 LCG steps."""
    p, n = 0.1, 100_000
    thr = np.uint64(int(p * 65536 + 0.5))
    lo, hi = np.uint64(0x12345678), np.uint64(0x9ABCDEF0)
    x = drop_hash(np.arange(n, dtype=np.uint64) * np.uint64(16), lo, hi)   # consecutive 32-key chunks
    keep = np.zeros((n, 32), dtype=bool)
    for j in range(16):
        h = x ^ (x >> np.uint64(16))
        keep[:, 2 * j] = (h & np.uint64(0xFFFF)) >= thr
        keep[:, 2 * j + 1] = (h >> np.uint64(16)) >= thr
        x = (x * np.uint64(0x2C9277B5) + np.uint64(0xAC564B05)) & M
    _check(keep, float(thr) / 65536)
