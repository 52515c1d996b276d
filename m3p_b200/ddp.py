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
    def __init__(self, model, group=None, overlap=True):
        self.model = model
        self.group = group
        self.overlap = overlap
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._works = []
        self._done = set()
        self._post = []
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"  # gloo has no AVG: SUM then scale
        model._grad_ready_hook = self._segment_ready if self.world > 1 else None

    def _allreduce(self, buf):
        if self._avg:
            self._works.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:
            self._works.append(dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self._post.append(buf)

    # segments are (lo, hi) element ranges of model._flat_grad
    def _segment_ready(self, name, lo, hi):
        if self.world == 1 or (lo, hi) in self._done:
            return
        self._done.add((lo, hi))
        buf = self.model._flat_grad[lo:hi]
        if self.overlap:
            self._allreduce(buf)

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
                self._allreduce(m._emb_grad)
                if m._proj_grad is not None and m._proj_grad is not m._emb_grad:
                    self._allreduce(m._proj_grad)
            for w in self._works:
                w.wait()  # current stream waits for NCCL's
            for buf in self._post:
                buf.div_(self.world)
        self._works = []
        self._done = set()
        self._post = []


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
