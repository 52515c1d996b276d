"""World-size-2 test of the data-parallel gradient exchange on CPU (gloo): the hook-driven, per-segment
all-reduce of m3p_b200/ddp.py averages every slice exactly once, with and without overlap."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeModel:
    """Just the attributes GradReducer touches (the real model needs a B200)."""

    def __init__(self, rank):
        self._flat_grad = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        self._emb_grad = torch.full((50, 8), float(rank + 1))
        self._proj_grad = self._emb_grad
        self._grad_ready_hook = None
        self._segments = {"embed": (0, 100), "layer0": (100, 400), "layer1": (400, 700), "heads": (700, 1000)}


def _worker(rank, world, port, overlap, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from m3p_b200.ddp import GradReducer
    m = _FakeModel(rank)
    red = GradReducer(m, overlap=overlap)
    assert m._grad_ready_hook is not None
    # the order the backward announces them: heads, last layer ... first layer; "embed" is left to finish()
    for name in ("heads", "layer1", "layer0"):
        m._grad_ready_hook(name, *m._segments[name])
    m._grad_ready_hook("layer0", *m._segments["layer0"])  # announcing twice must not reduce twice
    red.finish()
    want = torch.arange(1000, dtype=torch.float32) * (sum(range(1, world + 1)) / world)
    ok = torch.allclose(m._flat_grad, want) and torch.allclose(m._emb_grad, torch.full((50, 8), (world + 1) / 2))
    # a second step reuses the reducer
    m._flat_grad.fill_(float(rank))
    red.finish()
    ok = ok and torch.allclose(m._flat_grad, torch.full((1000,), (world - 1) / 2))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _worker_sparse(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from m3p_b200.ddp import GradReducer
    m = _FakeModel(rank)
    V, d = 50, 8
    g = torch.Generator().manual_seed(100 + rank)
    ids = torch.randint(0, V, (12,), generator=g)          # duplicates on purpose
    rows = torch.randn(12, d, generator=g)
    m._emb_grad = torch.zeros(V, d).index_add_(0, ids, rows)
    m._proj_grad = m._emb_grad
    m._emb_touched, m._emb_dense_dirty = [ids], False
    dense = [torch.zeros(V, d) for _ in range(world)]
    dist.all_gather(dense, m._emb_grad.clone())
    want = sum(dense) / world
    red = GradReducer(m)
    red.finish()
    ok = torch.allclose(m._emb_grad, want, atol=1e-6) and not m._emb_dense_dirty
    # the ids every rank touched are remembered, so the next zero_grad can clear exactly those rows
    cleared = m._emb_grad.clone().index_fill_(0, m._emb_touched[0], 0.0)
    ok = ok and float(cleared.abs().max()) == 0.0
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _worker_deferred(rank, world, port, q):
    """Tied MLM head: the dense head contribution is all-reduced when the heads finish, the embedding gather's
    per-position rows are exchanged at the end; the sum must equal the dense average of (head + gather) grads."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from m3p_b200.ddp import GradReducer
    m = _FakeModel(rank)
    V, d, T, B = 50, 8, 5, 3
    m.pad_index = 1
    g = torch.Generator().manual_seed(200 + rank)
    head = torch.randn(V, d, generator=g)                   # dense dE from the MLM head
    x = torch.randint(0, V, (T, B), generator=g)
    x[-1, 0] = 1                                            # a padding token: must receive nothing
    g_pos = torch.randn(B, T, d, generator=g)               # per-position gradient of the gather, batch-major
    full = head.clone()
    keep = (x.t().reshape(-1) != 1).float().unsqueeze(1)
    full.index_add_(0, x.t().reshape(-1), g_pos.reshape(-1, d) * keep)
    dense = [torch.zeros(V, d) for _ in range(world)]
    dist.all_gather(dense, full)
    want = sum(dense) / world
    m._emb_grad = head.clone()
    m._proj_grad = m._emb_grad
    m._emb_dense_dirty = True
    red = GradReducer(m)
    assert m._defer_token_grads
    m._grad_ready_hook("heads", *m._segments["heads"])      # dense part goes out here
    m._deferred_token_grads = [(x, g_pos)]                  # what _encode_backward leaves behind
    red.finish()
    ok = torch.allclose(m._emb_grad, want, atol=1e-5) and m._emb_dense_dirty and not m._deferred_token_grads
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_deferred_token_rows_after_early_dense_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_deferred, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_row_sparse_embedding_exchange_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sparse, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def test_overlapped_segment_allreduce_world2():
    _run(True)


def test_end_of_backward_allreduce_world2():
    _run(False)


def test_single_process_is_a_noop():
    from m3p_b200.ddp import GradReducer
    m = _FakeModel(0)
    before = m._flat_grad.clone()
    red = GradReducer(m)
    assert m._grad_ready_hook is None
    red.finish()
    assert torch.equal(m._flat_grad, before)


def _worker_accumulate(rank, world, port, q):
    """Gradient accumulation (`--accumulate_gradients 2`, FreeLB's ascent steps): micro-step 1 runs under
    reducer.accumulate() — its hooks send nothing — micro-step 2 adds local gradients to the same buffers and
    announces the slices; every slice must end up averaged exactly once, with both micro-steps inside."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from m3p_b200.ddp import GradReducer
    m = _FakeModel(rank)
    m._emb_touched, m._emb_dense_dirty, m._deferred_token_grads = None, True, []
    red = GradReducer(m)
    step1 = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    step2 = torch.ones(1000) * (10.0 * (rank + 1))
    m._flat_grad = step1.clone()
    with red.accumulate():
        assert not m._defer_token_grads
        for name in ("heads", "layer1", "layer0"):
            m._grad_ready_hook(name, *m._segments[name])
    ok = torch.equal(m._flat_grad, step1) and not red._works and not red._done and m._defer_token_grads
    m._flat_grad += step2                                    # the second backward accumulates on top
    for name in ("heads", "layer1", "layer0"):
        m._grad_ready_hook(name, *m._segments[name])
    red.finish()
    mean = sum(range(1, world + 1)) / world
    want = torch.arange(1000, dtype=torch.float32) * mean + 10.0 * mean
    ok = ok and torch.allclose(m._flat_grad, want)
    # entering accumulate() after something was already sent is a usage error, not silent corruption
    m._grad_ready_hook("heads", *m._segments["heads"])
    try:
        with red.accumulate():
            pass
        ok = False
    except RuntimeError:
        pass
    red.finish()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gradient_accumulation_sends_nothing_until_the_last_backward_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_accumulate, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
