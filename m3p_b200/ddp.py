"""Data-parallel gradient exchange for the M3P encoder path: one process per GPU, NCCL over NVLink /
NVSwitch through torch.distributed (the reference: apex.parallel.DistributedDataParallel with
delay_allreduce=True, M3P/src/xtrainer.py:77-83 — one flat sum-allreduce at the END of backward,
divided by world size, nothing overlapped).

Here the gradients already live in one flat fp32 buffer in reverse-usable order, so the exchange is
a handful of large in-place NCCL all-reduces (op = AVG, Apex's sum / world) on contiguous slices,
issued from inside the backward as soon as the last kernel writing a slice has been enqueued:
NCCL's stream waits for the compute stream at that point and then runs concurrently with the rest
of the backward.  Only the token-embedding matrix (68 % of the bytes) is reduced after the backward,
because its scatter-add is the very last kernel.  No forward collectives; loss terms stay local means
(DP average of local means == the reference's semantics).
"""
import torch
import torch.distributed as dist


class GradReducer:
    """reduce_dtype=torch.bfloat16: every dense slice is cast to a persistent bf16 staging buffer, all-reduced (AVG)
    and cast back into the fp32 gradient buffer, all on a dedicated communication stream — half the NVLink bytes and
    half the time NCCL's CTAs compete with the persistent GEMMs of the backward (Apex DDP, the reference's
    wrapper, all-reduces the half-precision AMP gradients the same way); the row-sparse embedding exchange sends
    bf16 rows.  None keeps the exchange in fp32 (bit-comparable with a single-process run, tools/check_ddp.py)."""

    def __init__(self, model, group=None, overlap=True, reduce_dtype=None):
        self.model = model
        self.group = group
        self.overlap = overlap
        self.reduce_dtype = reduce_dtype
        self._stage16 = {}   # (data_ptr, numel) of an fp32 slice -> persistent bf16 staging buffer
        self._comm = None    # communication stream of the bf16 path
        self._comm_pending = False
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._works = []
        self._done = set()
        self._post = []
        self._dense_sent = False
        self._gather = {}  # persistent (all_ids, all_rows) buffers of the row-sparse exchange, keyed by size
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"  # gloo has no AVG: SUM then scale
        model._grad_ready_hook = self._segment_ready if self.world > 1 else None
        model._defer_token_grads = self.world > 1 and overlap
        self._dense_sent = False
        self._accumulating = False

    def accumulate(self):
        """`with reducer.accumulate():` around every backward of a step EXCEPT the last one (gradient accumulation,
        `--accumulate_gradients`, xtrainer.py:229-243; the three ascent steps of FreeLB): nothing is sent from inside
        those backwards — a slice reduced during micro-step 1 would be averaged once and then receive un-averaged
        local gradients from micro-step 2 while possibly still in flight on NCCL's stream.  The last backward (outside
        the context) announces the slices as usual; they then hold the accumulated sum.  (Apex DDP with
        delay_allreduce does the same: one all-reduce when the last backward ends.)"""
        return _NoSync(self)

    def _allreduce_bf16(self, buf):
        from . import ops
        n = buf.numel()
        key = (buf.data_ptr(), n)
        if key not in self._stage16:
            self._stage16[key] = torch.empty(n, dtype=torch.bfloat16, device=buf.device)
        tmp = self._stage16[key]
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=buf.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self._comm.wait_event(ev)                      # the slice is final on the compute stream
        with torch.cuda.stream(self._comm), ops.on_stream(self._comm):
            flat = buf.reshape(-1)
            ops.cast_f32_bf16(flat, tmp, n, 1.0 if self._avg else 1.0 / self.world)
            dist.all_reduce(tmp, op=dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM, group=self.group)
            ops.cast_bf16_f32(tmp, flat, n)
        self._comm_pending = True

    def _join_comm(self):
        if self._comm_pending:
            ev = torch.cuda.Event()
            ev.record(self._comm)
            torch.cuda.current_stream().wait_event(ev)
            self._comm_pending = False

    def _allreduce(self, buf):
        if self.reduce_dtype == torch.bfloat16 and buf.is_cuda and buf.dtype == torch.float32 and buf.numel() % 8 == 0 \
                and buf.is_contiguous():
            return self._allreduce_bf16(buf)
        if self._avg:
            self._works.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:
            self._works.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._post.append(buf)

    # segments are (lo, hi) element ranges of model._flat_grad
    def _segment_ready(self, name, lo, hi):
        if self.world == 1 or self._accumulating or (lo, hi) in self._done:
            return
        self._done.add((lo, hi))
        buf = self.model._flat_grad[lo:hi]
        if self.overlap:
            self._allreduce(buf)
            m = self.model
            if name == "heads" and getattr(m, "_emb_dense_dirty", False) and getattr(m, "_defer_token_grads", False) \
                    and m._proj_grad is not None and not self._dense_sent:
                # the MLM head has just added its dense V x d contribution (68 % of all gradient bytes): send it now,
                # underneath the whole encoder backward; the embedding gather's rows follow in finish()
                self._allreduce(m._proj_grad)
                self._dense_sent = True

    def _exchange_embedding_rows(self, m, touched):
        """Row-sparse average of the token-embedding gradient.  Without the MLM head only the rows of the
        batch's tokens are non-zero (<= T*B of V = 250 002 rows), so instead of all-reducing the dense
        V x d matrix (68 % of all gradient bytes) every rank all-gathers (ids, rows):
            rows_r = G_r[ids_r] / (multiplicity of the id in ids_r * world)
            G     <- 0 on ids_r ;  G[ids_all] += rows_all        (duplicates re-sum to the full row)
        which equals sum_r G_r / world, the dense result."""
        g = m._emb_grad
        ids = torch.cat([t.reshape(-1) for t in touched])
        cnt = torch.zeros(g.shape[0], dtype=torch.float32, device=g.device)
        cnt.index_add_(0, ids, torch.ones(ids.numel(), dtype=torch.float32, device=g.device))
        rows = g.index_select(0, ids) / (cnt.index_select(0, ids) * self.world).unsqueeze(1)
        # persistent receive buffers: the NEXT step's zero_grad clears the rows listed in all_ids, so under CUDA-graph
        # replay (and across eager steps) that tensor must stay where it is
        rdt = self.reduce_dtype if (self.reduce_dtype is not None and g.is_cuda) else g.dtype
        key = (self.world * ids.numel(), g.shape[1], rdt, ids.device)
        if key not in self._gather:
            self._gather[key] = (torch.empty(key[0], dtype=ids.dtype, device=ids.device),
                                 torch.empty(key[0], key[1], dtype=rdt, device=g.device))
        all_ids, all_rows = self._gather[key]
        if all_rows.dtype != rows.dtype:
            rows = rows.to(all_rows.dtype)
        dist.all_gather_into_tensor(all_ids, ids, group=self.group)
        dist.all_gather_into_tensor(all_rows, rows, group=self.group)
        g.index_fill_(0, ids, 0.0)
        self._scatter_add(m, all_rows, all_ids, g)
        m._emb_touched = [all_ids]  # what the next zero_grad has to clear

    @staticmethod
    def _scatter_add(m, rows, ids, g):
        """g[ids[i]] += rows[i]: m3p_scatter_add_rows_f32 on the GPU (one launch, float4 reductions), index_add_ on
        the CPU test double."""
        if g.is_cuda:
            from . import ops
            ops.use_current_stream()
            ops.scatter_add_rows_f32(rows, ids, -1, g, ids.numel(), g.shape[1])
        else:
            g.index_add_(0, ids, rows)

    def _exchange_deferred_rows(self, m, deferred):
        """Tied MLM head under data parallelism: the dense part of d E is already in flight; the embedding
        gather's contribution was kept per token position (B, T, d).  All ranks all-gather (ids, rows / world) and
        add every rank's rows into the (by then averaged) dense buffer: sum_r (dense_r + gather_r) / world."""
        g = m._emb_grad
        ids = torch.cat([x.t().reshape(-1) for x, _ in deferred])                  # batch-major, like the rows
        rows = torch.cat([gp.reshape(-1, g.shape[1]) for _, gp in deferred])
        rows = rows * ((ids != m.pad_index).to(rows.dtype) / self.world).unsqueeze(1)  # padding_idx gets no gradient
        rdt = self.reduce_dtype if (self.reduce_dtype is not None and g.is_cuda) else g.dtype
        key = ("deferred", self.world * ids.numel(), g.shape[1], rdt, ids.device)
        if key not in self._gather:
            self._gather[key] = (torch.empty(key[1], dtype=ids.dtype, device=ids.device),
                                 torch.empty(key[1], key[2], dtype=rdt, device=g.device))
        all_ids, all_rows = self._gather[key]
        if all_rows.dtype != rows.dtype:
            rows = rows.to(all_rows.dtype)
        dist.all_gather_into_tensor(all_ids, ids, group=self.group)
        dist.all_gather_into_tensor(all_rows, rows, group=self.group)
        for w in self._works:
            w.wait()  # the dense all-reduce (and everything else in flight) has landed on the current stream
        self._join_comm()
        for buf in self._post:
            buf.div_(self.world)
        self._works, self._post = [], []
        self._scatter_add(m, all_rows, all_ids, g)

    def finish(self):
        """Call after backward(): reduces whatever has not been sent yet and joins the NCCL stream."""
        m = self.model
        if self.world > 1:
            if m._flat_grad is not None:
                # anything not announced by a hook (or everything, when overlap is off)
                covered = sorted(self._done)
                pos = 0
                rest = []
                for lo, hi in covered:
                    if lo > pos:
                        rest.append((pos, lo))
                    pos = max(pos, hi)
                if pos < m._flat_grad.numel():
                    rest.append((pos, m._flat_grad.numel()))
                if not self.overlap:
                    rest = [(0, m._flat_grad.numel())]  # nothing was sent from the hooks
                for lo, hi in rest:
                    self._allreduce(m._flat_grad[lo:hi])
                touched = getattr(m, "_emb_touched", None)
                deferred, m._deferred_token_grads = getattr(m, "_deferred_token_grads", []), []
                if self._dense_sent:
                    if m._proj_grad is not m._emb_grad:     # untied: the projection went early, the table as usual
                        if touched and not m._emb_dense_dirty:
                            self._exchange_embedding_rows(m, touched)
                        else:
                            self._allreduce(m._emb_grad)
                    elif deferred:
                        self._exchange_deferred_rows(m, deferred)
                    m._emb_dense_dirty = True
                else:
                    if touched and not getattr(m, "_emb_dense_dirty", True):
                        self._exchange_embedding_rows(m, touched)
                    else:
                        self._allreduce(m._emb_grad)
                        m._emb_dense_dirty = True  # rows touched by OTHER ranks are now non-zero here too
                    if m._proj_grad is not None and m._proj_grad is not m._emb_grad:
                        self._allreduce(m._proj_grad)
            for w in self._works:
                w.wait()  # current stream waits for NCCL's
            self._join_comm()
            for buf in self._post:
                buf.div_(self.world)
        self._works = []
        self._done = set()
        self._post = []
        self._dense_sent = False


class _NoSync:
    def __init__(self, reducer):
        self.r = reducer

    def __enter__(self):
        r = self.r
        if r._works or r._done:
            raise RuntimeError("GradReducer.accumulate(): a previous backward already sent gradients; wrap every backward "
                               "but the last one, and call finish() after the last")
        self.prev = (r._accumulating, r.model._defer_token_grads)
        r._accumulating, r.model._defer_token_grads = True, False

    def __exit__(self, *exc):
        self.r._accumulating, self.r.model._defer_token_grads = self.prev


def init_distributed():
    """torchrun-style env:// initialisation (reference: M3P/src/slurm.py:46-170, SLURM handling dropped)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, init_method="env://", world_size=world, rank=rank, **kw)
    return rank, local, world
