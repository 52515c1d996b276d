"""The trainer-side step that drives the hot path, restated from the reference trainer (which cannot
be imported: it needs apex): `XTrainer.pretrain_under_step` (M3P/src/xtrainer.py:2234-2402),
`t2i_step`/`i2t_step` (:1888-2018) and `get_mask_` (:2226-2232), with the host syncs removed
(`.item()`, `.cpu().numpy()`, CPU-side ITM loss at :2367-2370 all stay on the device).

Batches are dicts shaped like `retrieval_pretrain_collate` output (xtrainer.py:960-1045):
  x (T,B) int64, lengths (B,), x_img (R,B,2048) fp32, lengths_img (B,), image_loc (R,B,5) fp32,
  x_labels (T,B) int64 (-1 = not masked), obj_labels (B,R) int64 (-1 = not masked),
  ori_feats (B,R,2048) fp32, pos_labels (B/sample_n,) int64.
"""
import os

import torch
import torch.nn.functional as F


def get_mask_(labels):
    """xtrainer.py:2226-2232 — host-side batch preparation (boolean indexing syncs; do it once per batch)."""
    return labels[labels > 0], labels != -1


def prepare_batch(batch):
    """Adds the derived tensors the step needs (`y_text`, `pred_mask_text`, `mrfr_weight`) so the step
    itself never synchronises with the host."""
    b = dict(batch)
    if "x_labels" in b and "y_text" not in b:
        b["y_text"], b["pred_mask_text"] = get_mask_(b["x_labels"])
    if "obj_labels" in b and "mrfr_weight" not in b:
        sel = (b["obj_labels"].reshape(-1) != -1).to(torch.float32)
        b["mrfr_weight"] = sel / (sel.sum().clamp_min(1.0) * 2048.0)
    return b


class _MaskedMse(torch.autograd.Function):
    """F.mse_loss(pred[sel], target[sel]) of the MRFR objective (xtrainer.py:2333-2348) as two kernels
    (m3p_masked_mse_fwd / _bwd): pred (n, d) bf16, target (n, d) fp32, weight (n,) = sel / (n_sel * d)."""

    @staticmethod
    def forward(ctx, pred, target, weight):
        from . import ops
        ops.use_current_stream()
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        ops.masked_mse_fwd(pred, target, weight, loss)
        ctx.save_for_backward(pred, target, weight)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        from . import ops
        ops.use_current_stream()
        pred, target, weight = ctx.saved_tensors
        dpred = torch.empty_like(pred)
        ops.masked_mse_bwd(pred, target, weight, dloss.to(torch.float32).contiguous(), dpred)
        return dpred, None, None


def masked_mse(pred, target, weight):
    """sum_rows weight[row] * sum_f (pred - target)^2 on the device; CPU tensors (host-side tests of the step logic)
    take the same expression in torch."""
    if not pred.is_cuda:
        diff = pred.float() - target
        return (diff * diff * weight[:, None]).sum()
    pred = pred if pred.dtype == torch.bfloat16 else pred.to(torch.bfloat16)
    return _MaskedMse.apply(pred.contiguous(), target.contiguous(), weight.contiguous())


class _RelationLoss(torch.autograd.Function):
    """The ITM loss and its gradient in one kernel (m3p_relation_loss)."""

    @staticmethod
    def forward(ctx, scores, pos_labels, sample_n, w_multi, w_bin):
        from . import ops
        ops.use_current_stream()
        s = scores.reshape(-1).to(torch.float32).contiguous()
        loss = torch.empty((), dtype=torch.float32, device=s.device)
        ds = torch.empty_like(s)
        ops.relation_loss(s, pos_labels.contiguous(), sample_n, w_multi, w_bin, loss, ds)
        ctx.save_for_backward(ds)
        ctx.shape = scores.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (ds,) = ctx.saved_tensors
        return (ds * dloss).view(ctx.shape), None, None, None, None


def relation_loss(scores, pos_labels, sample_n, w_multi=1.0, w_bin=1.0):
    """xtrainer.py:2359-2372 / 1917-1942: CE over groups of sample_n + BCE against the one-hot positive (one kernel
    on the device; the same expression in torch for CPU tensors)."""
    if scores.is_cuda:
        return _RelationLoss.apply(scores, pos_labels, sample_n, float(w_multi), float(w_bin))
    ce = F.cross_entropy(scores.view(-1, sample_n), pos_labels)
    onehot = F.one_hot(pos_labels, sample_n).to(scores.dtype)
    bce = F.binary_cross_entropy_with_logits(scores.view(-1), onehot.view(-1))
    return w_multi * ce + w_bin * bce


def pretrain_step(model, batch, sample_n=4, heads=("mlm", "mrm", "mrfr", "rel"), lambdas=None):
    """jointfwd + the selected heads + loss assembly (xtrainer.py:2281-2375).  `heads=("rel",)` is the
    fine-tune ITM step (t2i_step / i2t_step, :1911-1942).  Returns (total_loss, dict of losses)."""
    lam = dict(mlm=1.0, mrm=1.0, mrfr=1.0, rel=1.0, clcm=1.0)
    if lambdas:
        lam.update(lambdas)
    R = batch["x_img"].shape[0]
    enc = model("jointfwd", x=batch["x"], lengths=batch["lengths"], x_img=batch["x_img"],
                lengths_img=batch["lengths_img"], causal=False, langs=None, image_loc=batch["image_loc"],
                refine_image=False)
    text_out = enc[R:]                     # :2287-2289
    img_out = enc[:R].transpose(0, 1)
    losses = {}
    total = None

    def add(name, value):
        nonlocal total
        losses[name] = value
        total = lam[name] * value if total is None else total + lam[name] * value

    if "mlm" in heads:
        _, l = model("predict", tensor=text_out, pred_mask=batch["pred_mask_text"], y=batch["y_text"], get_scores=False)
        add("mlm", l)
    if "mrm" in heads:
        _, l = model("predict", tensor=img_out, pred_mask=None, y=batch["obj_labels"].reshape(-1), get_scores=False,
                     is_obj=True)
        add("mrm", l)
    if "mrfr" in heads:
        reg = model("predict", tensor=img_out, is_mrfr=True)
        # == F.mse_loss(reg[sel], target[sel]) (:2334-2348) without the boolean gather
        add("mrfr", masked_mse(reg.reshape(-1, 2048), batch["ori_feats"].reshape(-1, 2048), batch["mrfr_weight"]))
    if "rel" in heads:
        scores = model("predict", tensor=enc.transpose(0, 1), is_relation=True)
        add("rel", relation_loss(scores, batch["pos_labels"], sample_n))
    if "clcm" in heads:
        # i2t branch (:2379-2393): a second jointfwd over the code-switched caption x2 with the SAME regions, the
        # second pooler + classifier (is_clcm), BCE against clcm_labels; added to the total with weight 1
        enc2 = model("jointfwd", x=batch["x2"], lengths=batch["lengths2"], x_img=batch["x_img"],
                     lengths_img=batch["lengths_img"], causal=False, langs=None, image_loc=batch["image_loc"],
                     refine_image=False)
        scores2 = model("predict", tensor=enc2.transpose(0, 1), is_clcm=True)
        add("clcm", F.binary_cross_entropy_with_logits(scores2.view(-1), batch["clcm_labels"].view(-1).to(scores2.dtype)))
    return total, losses


def mask_out(x, lengths, n_words, pad_index=1, mask_index=None, word_pred=0.15, pred_probs=(0.8, 0.1, 0.1),
             round_to=8, generator=None):
    """Trainer.mask_out (xtrainer.py:385-434, the sample_alpha == 0 branch of the published recipes): choose
    word_pred of the positions (never padding, never position 0), round the count down to a multiple of 8 as the
    reference does for fp16 (:409-415), and replace each chosen token by <mask> / itself / a random token with
    probabilities pred_probs.  Host-side batch preparation like the reference's (numpy there, torch here);
    returns (x_masked, y, pred_mask) with y = the original tokens in (slen, bs) row-major order."""
    slen, bs = x.shape
    mask_index = n_words - 1 if mask_index is None else mask_index  # tokenization.py:79-81: <mask> is the last id
    pred_mask = torch.rand(slen, bs, generator=generator) <= word_pred
    pred_mask &= x.cpu() != pad_index
    pred_mask[0] = False
    flat = pred_mask.view(-1)
    n1 = int(flat.sum())
    if round_to > 1:
        n2 = max(n1 % round_to, round_to * (n1 // round_to))
        if n2 != n1:
            flat[torch.nonzero(flat).view(-1)[:n1 - n2]] = False
    if int(flat.sum()) == 0:
        pred_mask[0, 0] = True
    pred_mask = pred_mask.to(x.device)
    x_real = x[pred_mask]
    x_rand = torch.randint(0, n_words, x_real.shape, generator=generator).to(x.device)
    probs = torch.multinomial(torch.tensor(pred_probs), len(x_real), replacement=True, generator=generator).to(x.device)
    x_new = torch.where(probs == 0, torch.full_like(x_real, mask_index), torch.where(probs == 1, x_real, x_rand))
    x = x.masked_scatter(pred_mask, x_new)
    assert 0 <= int(x.min()) <= int(x.max()) < n_words
    return x, x_real, pred_mask


def mlm_step(model, x, lengths, pred_mask, y, positions=None, langs=None, lambda_coeff=1.0):
    """Trainer.mlm_step (xtrainer.py:734-770), the xMLM / TLM objective: text stream through `crossfwd`
    (+ language embeddings when `langs` is given), MLM head on the masked positions.  Inputs are what
    `mask_out` returns, already on the device.  Returns lambda_coeff * loss."""
    tensor = model("crossfwd", stream_="text", x=x, lengths=lengths, positions=positions, langs=langs, causal=False)
    _, loss = model("predict", tensor=tensor, pred_mask=pred_mask, y=y, get_scores=False)
    return lambda_coeff * loss


def init_adv_delta(like, row_dims, adv_init_mag=1e-4, norm_type="l2", generator=None):
    """Initial FreeLB perturbation (deal_freelb_delta / deal_image_freelb_delta, xtrainer.py:2700-2736): uniform in
    [-1, 1] scaled per slice of dim 0 by adv_init_mag / sqrt(row_dims) ("l2"; row_dims = lengths * dim for the token
    embeddings (B, T, d), the feature width for the (R, B, 2048) region features — the reference scales those per
    region index, and so does this), or uniform in [-adv_init_mag, adv_init_mag] ("linf"); zeros when the magnitude
    is 0."""
    if adv_init_mag <= 0:
        return torch.zeros_like(like)
    if norm_type == "linf":
        return torch.zeros_like(like).uniform_(-adv_init_mag, adv_init_mag, generator=generator)
    if norm_type != "l2":
        raise NotImplementedError("Norm type {} not specified.".format(norm_type))
    noise = torch.zeros_like(like).uniform_(-1, 1, generator=generator)
    mag = adv_init_mag / torch.sqrt(row_dims.to(torch.float32))
    return (noise * mag.to(like.device).view(-1, *([1] * (like.dim() - 1)))).detach()


def ascend_adv_delta(delta, delta_grad, adv_lr=1e-3, adv_max_norm=1e-2, norm_type="l2"):
    """One ascent step on a FreeLB perturbation (update_freelb_delta / update_image_freelb_delta,
    xtrainer.py:2793-2851): normalise the gradient per slice of dim 0, step by adv_lr, project back onto the
    adv_max_norm ball (l2) or box (linf)."""
    n0 = delta.size(0)
    bshape = (n0,) + (1,) * (delta.dim() - 1)
    g = delta_grad.detach()
    if norm_type == "l2":
        gn = g.reshape(n0, -1).norm(dim=1).clamp(min=1e-8).view(bshape)
        delta = (delta.detach() + adv_lr * g / gn).detach()
        if adv_max_norm > 0:
            dn = delta.reshape(n0, -1).float().norm(p=2, dim=1)
            over = (dn > adv_max_norm).to(delta.dtype)
            delta = (delta * (adv_max_norm / dn * over + (1 - over)).view(bshape)).detach()
    elif norm_type == "linf":
        gn = g.reshape(n0, -1).norm(dim=1, p=float("inf")).clamp(min=1e-8).view(bshape)
        delta = (delta.detach() + adv_lr * g / gn).detach()
        if adv_max_norm > 0:
            delta = delta.clamp(-adv_max_norm, adv_max_norm).detach()
    else:
        raise NotImplementedError("Norm type {} not specified.".format(norm_type))
    return delta


def freelb_relation_step(model, batch, sample_n=4, adv_steps=3, adv_init_mag=1e-4, adv_lr=1e-3, adv_max_norm=1e-2,
                         norm_type="l2", optimizer=None, reducer=None, init=None, trace=None):
    """FreeLB variant of the ITM fine-tune step (freelb_t2i_step / freelb_i2t_step, xtrainer.py:2021-2223,
    `--is_freelb`): `adv_steps` ascent steps on perturbations of the token embeddings
    (`jointfwd(text_embed=embeds_init + delta)`, transformer.py:910-913) and of the region features.  `embeds_init =
    model.embeddings(ids)` stays attached to the graph, so the table is trained through it, and is looked up again
    after every ascent step (:2820-2823).  Each step's loss is divided by adv_steps and back-propagated:
      * optimizer given  — the reference's `free_optimize` without AMP (:2766-2776): zero_grad, backward, clip +
        optimizer step at EVERY ascent step;
      * optimizer = None — its AMP branch inside an accumulation window (:2789-2791): the parameter gradients of
        the ascent steps accumulate in the flat buffer and the caller steps afterwards.
    `reducer` (ddp.GradReducer): accumulated backwards run under `reducer.accumulate()`; `finish()` follows every
    backward that is followed by an optimizer step, or the last one.  `init` = (delta_text, delta_img) overrides
    the random initialisation (tests); `trace`, if a list, receives (loss, delta_text, delta_img) per step.
    Returns the summed loss (a device scalar; the reference logs the same sum)."""
    x = batch["x"]
    ids = x.transpose(0, 1)
    embeds_init = model.embeddings(ids)                                               # (B, T, d) fp32, :2700-2705
    if init is not None:
        delta_t, delta_i = init
    else:
        delta_t = init_adv_delta(embeds_init, batch["lengths"] * embeds_init.size(-1), adv_init_mag, norm_type)
        rdims = torch.full((batch["x_img"].size(0),), float(batch["x_img"].size(-1)))
        delta_i = init_adv_delta(batch["x_img"], rdims, adv_init_mag, norm_type)
    total = 0.0
    for astep in range(adv_steps):
        last = astep == adv_steps - 1
        delta_t = delta_t.detach().requires_grad_(True)
        delta_i = delta_i.detach().requires_grad_(True)
        enc = model("jointfwd", x=x, lengths=batch["lengths"], x_img=batch["x_img"] + delta_i,
                    lengths_img=batch["lengths_img"], causal=False, image_loc=batch["image_loc"],
                    text_embed=delta_t + embeds_init)
        scores = model("predict", tensor=enc.transpose(0, 1), is_relation=True)
        loss = relation_loss(scores, batch["pos_labels"], sample_n) / (1.0 * adv_steps)
        if optimizer is not None:
            model.zero_grad()
        if reducer is not None and optimizer is None and not last:
            with reducer.accumulate():
                loss.backward()
        else:
            loss.backward()
            if reducer is not None:
                reducer.finish()
        if optimizer is not None:
            optimizer.step()
        total = total + loss.detach()
        if trace is not None:
            trace.append((loss.detach(), delta_t.detach(), delta_i.detach(), delta_t.grad, delta_i.grad))
        if last:
            break
        delta_t = ascend_adv_delta(delta_t, delta_t.grad, adv_lr, adv_max_norm, norm_type)
        delta_i = ascend_adv_delta(delta_i, delta_i.grad, adv_lr, adv_max_norm, norm_type)
        embeds_init = model.embeddings(ids)                                           # :2820-2823
    return total


def synthetic_batch(B, T, R, n_words, sample_n=4, seed=1234, ragged=False, n_mask_text=16, n_mask_img=16,
                    device="cpu", feat_dim=2048):
    """Seeded synthetic batch in the layout and value distributions of the reference pipeline
    (SURVEY.md §8d): unit-L2 region features (dataset_pretrain.py:287,379), normalised 5-d boxes
    (:298-300), <s>=0 / </s>=2 / <pad>=1 framing (xtrainer.py:829-880), masked regions zeroed (:258-292)."""
    g = torch.Generator().manual_seed(seed)
    lengths = torch.full((B,), T, dtype=torch.long)
    if ragged:
        lengths = torch.randint(max(4, T // 4), T + 1, (B,), generator=g)
        lengths[0] = T
    x = torch.randint(4, n_words - 2, (T, B), generator=g)
    x[0] = 0
    ar = torch.arange(T)[:, None]
    x = torch.where(ar == (lengths - 1)[None, :], torch.full_like(x, 2), x)
    x = torch.where(ar >= lengths[None, :], torch.full_like(x, 1), x)
    x_img = F.normalize(torch.randn(R, B, feat_dim, generator=g), dim=-1)
    loc = torch.rand(R, B, 5, generator=g)
    image_loc = loc / loc.norm(dim=-1, keepdim=True)
    lengths_img = torch.full((B,), R, dtype=torch.long)
    ori_feats = x_img.transpose(0, 1).clone()
    x_labels = torch.full((T, B), -1, dtype=torch.long)
    obj_labels = torch.full((B, R), -1, dtype=torch.long)
    for b in range(B):
        n_valid = int(lengths[b]) - 2
        k = min(n_mask_text, max(n_valid, 0))
        if k > 0:
            pos = torch.randperm(n_valid, generator=g)[:k] + 1
            x_labels[pos, b] = torch.randint(4, n_words - 2, (k,), generator=g)
        posi = torch.randperm(R, generator=g)[:min(n_mask_img, R)]
        obj_labels[b, posi] = torch.randint(1, 1600, (len(posi),), generator=g)
        x_img[posi, b] = 0.0
    assert B % sample_n == 0
    pos_labels = torch.randint(0, sample_n, (B // sample_n,), generator=g)
    batch = dict(x=x, lengths=lengths, x_img=x_img, lengths_img=lengths_img, image_loc=image_loc, x_labels=x_labels,
                 obj_labels=obj_labels, ori_feats=ori_feats, pos_labels=pos_labels)
    batch = prepare_batch(batch)
    if device != "cpu":
        batch = {k: v.to(device) for k, v in batch.items()}
    return batch


class GraphedStep:
    """One training step (zero_grad + forward + loss + backward) captured ONCE into a CUDA graph and replayed:
    ~320 kernel launches collapse into a single graph launch, removing the host from the step.  Inputs are
    static device tensors (`self.batch`); `step(new_batch)` copies a new batch into them first (H2D from pinned
    host memory is asynchronous).  Dropout stays fresh across replays because the seeds live in a device word
    bumped inside the graph (TransformerModel._next_seed)."""

    def __init__(self, model, batch, sample_n=4, heads=("rel",), lambdas=None, warmup=3, after_backward=None,
                 capture_error_mode="global"):
        self.model, self.batch = model, {k: v.clone() for k, v in batch.items()}
        self.sample_n, self.heads, self.lambdas, self.after_backward = sample_n, heads, lambdas, after_backward
        model.invalidate_operands()  # the captured step always re-casts the operands and clears the gradients:
        model._grads_clean = False   # what the graph contains must not depend on what an optimizer did before
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up off the default stream: allocations, kernel attributes, scratch
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        from . import ops
        n0 = ops.LAUNCHES
        # after_backward may enqueue NCCL collectives (ddp.GradReducer.finish): they are captured like any other
        # kernel; pass capture_error_mode="thread_local" then, so NCCL's watchdog thread may keep polling events
        # The activation-gradient chain is captured on a HIGH-priority stream and the parameter-gradient work of
        # `_SideQueue` on a default-priority one: when both have thread blocks pending, an SM that frees up goes to the
        # chain, and the weight-gradient GEMMs fill what the chain leaves idle (kernel nodes keep their stream's priority).
        hi = torch.cuda.Stream(priority=-1) if os.environ.get("M3P_CHAIN_PRIORITY", "1") != "0" else None
        with torch.cuda.graph(self.graph, stream=hi, capture_error_mode=capture_error_mode):
            self.loss = self._eager()
        self.launches_per_step = ops.LAUNCHES - n0

    def _eager(self):
        # the gradient clear (two memsets + a row clear, ~75 us of pure HBM traffic) runs on the model's side stream
        # underneath the tensor-bound forward; the backward waits for it
        cur = torch.cuda.current_stream()
        side = self.model._side_stream_for(self.model._flat.device)
        fork, cleared = torch.cuda.Event(), torch.cuda.Event()
        fork.record(cur)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            self.model.zero_grad()
            cleared.record(side)
        total, _ = pretrain_step(self.model, self.batch, self.sample_n, self.heads, self.lambdas)
        cur.wait_event(cleared)
        total.backward()
        self._snapshot_touched()
        if self.after_backward is not None:
            self.after_backward()
        return total.detach()

    def _snapshot_touched(self):
        """The next zero_grad clears the token-embedding rows listed in model._emb_touched.  Those lists alias the
        graph's static input tensors, which `step(new_batch)` overwrites before the next replay — the replayed clear
        would then miss the rows the previous batch touched.  Copy the ids into persistent buffers INSIDE the step
        (captured with it), so every replay clears exactly what its predecessor wrote."""
        m = self.model
        touched = m._emb_touched
        if not touched:
            return
        if not hasattr(self, "_touched_bufs"):
            self._touched_bufs = [torch.empty_like(t) for t in touched]
        if len(self._touched_bufs) != len(touched) or any(b.shape != t.shape for b, t in zip(self._touched_bufs, touched)):
            return  # different step shape (not reachable through this class): keep the aliasing lists
        for b, t in zip(self._touched_bufs, touched):
            if b.data_ptr() != t.data_ptr():
                b.copy_(t)
        m._emb_touched = list(self._touched_bufs)

    def release(self):
        """Drop the captured graph (needed before torch.distributed.destroy_process_group() when collectives
        were captured: NCCL requires graphs that hold its kernels to be destroyed before the communicator)."""
        self.graph.reset()
        self.graph = None

    def step(self, new_batch=None):
        if new_batch is not None:
            for k, v in new_batch.items():
                if k in self.batch:
                    self.batch[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.loss

    # -- input pipelining: the host -> device copy of batch i+1 runs on a copy stream underneath step i ----------
    def prefetch(self, host_batch):
        """Start the asynchronous H2D copy of the NEXT batch (pinned host tensors) into a staging set on a copy
        stream; `step_prefetched()` then moves it into the graph's static inputs with device-to-device copies
        (a few tens of microseconds) instead of waiting for PCIe in front of the replay."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._staging = {k: torch.empty_like(v) for k, v in self.batch.items()}
            self._staged_ready = torch.cuda.Event()
            self._staging_free = torch.cuda.Event()
            self._staging_free.record(torch.cuda.current_stream())
            self._staged_keys = ()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._staging_free)  # the previous staged batch has been consumed
            keys = []
            for k, v in host_batch.items():
                if k in self._staging:
                    self._staging[k].copy_(v, non_blocking=True)
                    keys.append(k)
            self._staged_ready.record(self._copy_stream)
        self._staged_keys = tuple(keys)

    def step_prefetched(self, next_host_batch=None):
        """Run one step on the batch staged by the last `prefetch()`, and start prefetching `next_host_batch`."""
        if not hasattr(self, "_staged_ready"):
            raise RuntimeError("step_prefetched() needs a batch staged by prefetch() first")
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged_ready)
        for k in self._staged_keys:
            self.batch[k].copy_(self._staging[k], non_blocking=True)
        self._staging_free.record(cur)
        if next_host_batch is not None:
            self.prefetch(next_host_batch)
        self.graph.replay()
        return self.loss
