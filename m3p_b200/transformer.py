"""B200-native stand-in for the reference's `src.model.transformer.TransformerModel` on the
encoder training path (reference: M3P/src/model/transformer.py).

Same constructor (`TransformerModel(params, is_encoder, with_output, is_crossModal)`, built as in
model/__init__.py:93), same `forward(mode, **kwargs)` dispatcher (:731-751) for the modes
`jointfwd` (:878-968), `fwd` (:753-876), `crossfwd` (:970-1114) and `predict` (:1183-1214), same
module-level `get_masks` (:59-78), and the same `state_dict` names and shapes, so a reference
checkpoint loads unchanged and `XTrainer`'s calls (`model('jointfwd', ...)`, `model('predict', ...)`,
`loss.backward()`) keep working.  Everything between those calls runs in hand-written sm_100a
kernels reached through the C ABI (include/m3p_b200.h): there is no PyTorch implementation of the
math in this file and no fallback — sub-modes outside the encoder training path (causal masks, KV
cache, decoder cross-attention, the AoA refiner) raise NotImplementedError.

Memory layout (HBM):
  * hot parameters live in ONE flat fp32 buffer (`_flat`), every nn.Parameter is a view into it, with
    a bf16 mirror (`_flat16`) that feeds the tensor cores and a flat fp32 gradient buffer
    (`_flat_grad`, `param.grad` are views) that kernels accumulate into and that data-parallel
    training all-reduces in a few large NCCL buckets (m3p_b200/ddp.py).  q/k/v weights (and biases)
    are adjacent, so the packed [3d, d] QKV operand is a view, not a copy.
  * activations are bf16, batch-major rows (row = b*S + s); the (slen, bs, dim) tensors of the
    reference API are transposed views of them.
"""
import math
import os

import torch
import torch.nn as nn

from . import lib as L
from . import ops

N_MAX_POSITIONS = 514  # transformer.py:16
LN_EPS = 1e-12         # transformer.py:244,659,694,709
FEAT_DIM = 2048        # region feature width (transformer.py:239)
LOC_DIM = 5
N_OBJ = 1600           # transformer.py:570

_BF16 = torch.bfloat16
_F32 = torch.float32


def get_masks(slen, lengths, causal, k=None):
    """Hidden-state mask (and the identical attention mask) — transformer.py:59-78, non-causal."""
    if causal:
        raise NotImplementedError("causal masks are outside the B200 encoder training path")
    if not lengths.is_cuda:  # the reference's check (:63); on the device it would cost a host sync per call
        assert lengths.max().item() <= slen
    alen = torch.arange(slen, dtype=torch.long, device=lengths.device)
    mask = alen < lengths[:, None]
    return mask, mask


class _Node(nn.Module):
    """Name-space holder so parameters register under the reference's dotted names."""


class _EmbeddingNode(_Node):
    """`model.embeddings` — callable like the reference's nn.Embedding(n_words, dim, padding_idx)
    (transformer.py:21-26, 658): `model.embeddings(input_ids)` is what FreeLB builds `embeds_init` from
    (xtrainer.py:2700-2705, 2820-2823), attached to the graph so the table is trained through it."""

    def forward(self, input_ids):
        return _EmbedLookupFn.apply(self.weight, input_ids, self.__dict__["_owner"]())


def _align(n, a=64):
    return (n + a - 1) // a * a


def _linear_init_(w, b, fan_in):
    # torch.nn.Linear default (transformer.py:29-34 keeps it): U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    bound = 1.0 / math.sqrt(fan_in)
    w.uniform_(-bound, bound)
    if b is not None:
        b.uniform_(-bound, bound)


class TransformerModel(nn.Module):
    ATTRIBUTES = ['encoder', 'with_output', 'eos_index', 'pad_index', 'n_langs', 'n_words', 'dim', 'n_layers',
                  'n_heads', 'hidden_dim', 'dropout', 'attention_dropout', 'asm', 'asm_cutoffs', 'asm_div_value']

    def __init__(self, params, is_encoder, with_output, is_crossModal=False):
        super().__init__()
        if not is_encoder:
            raise NotImplementedError("the B200 path implements the encoder (is_encoder=True) only")
        if not is_crossModal:
            # the reference constructor itself only works with is_crossModal=True (transformer.py:673-698)
            raise NotImplementedError("construct with is_crossModal=True as model/__init__.py:93 does")
        if getattr(params, "asm", False):
            raise NotImplementedError("adaptive softmax (asm=True) is outside the B200 path")
        if getattr(params, "sinusoidal_embeddings", False):
            raise NotImplementedError("sinusoidal position embeddings are outside the B200 path")
        if not getattr(params, "gelu_activation", True):
            raise NotImplementedError("only the erf-GELU FFN (gelu_activation=True) is implemented")
        self.is_encoder, self.is_decoder = True, False
        self.with_output = with_output
        self.is_crossModal = is_crossModal
        self.n_langs = params.n_langs
        self.n_words = params.n_words
        self.eos_index = params.eos_index
        self.pad_index = params.pad_index
        self.id2lang = params.id2lang
        self.lang2id = params.lang2id
        self.english_only = not (self.n_langs > 1)
        assert len(self.id2lang) == len(self.lang2id) == self.n_langs
        self.dim = params.emb_dim
        self.hidden_dim = self.dim * 4
        self.n_heads = params.n_heads
        self.n_layers = params.n_layers
        self.dropout = params.dropout
        self.attention_dropout = params.attention_dropout
        self.share_inout_emb = bool(getattr(params, "share_inout_emb", True))
        self.refine_layers = int(getattr(params, "refine_layers", 6))
        assert self.dim % self.n_heads == 0, 'transformer dim must be a multiple of n_heads'
        if self.dim // self.n_heads != 64:
            raise NotImplementedError("the tcgen05 attention kernel needs head dim 64 (M3P-base 768/12, large 1024/16); "
                                      "got %d" % (self.dim // self.n_heads))
        self._seed_base = 0x5DEECE66D
        self._step = 0
        self._seed_words = None
        self._grad_ready_hook = None  # set by ddp.GradReducer: (name, lo, hi) slice of _flat_grad is final
        self._device_checked = False
        self._emb_touched = None      # token-id tensors scattered into _emb_grad since it was last cleared
        self._emb_dense_dirty = True  # True once a dense update (tied MLM head) or foreign writer touched it
        self._defer_token_grads = False      # set by ddp.GradReducer (world > 1): see _encode_backward
        self._deferred_token_grads = []      # [(token ids (T,B), per-position gradient (B,T,d) fp32)]
        self._bwd_trace = None  # a list here makes _encode_backward record its intermediates (stage-parity tests)
        self.overlap_grads = os.environ.get("M3P_SIDE_STREAM", "1") != "0"
        self._side_streams = {}
        # set by m3p_b200.optim after a fused step (which also writes the bf16 operand copies and zeroes the
        # gradients): lets refresh_operands() / zero_grad() skip work that has already been done
        self._operands_valid = False
        self._emb16_valid = False
        self._emb16_event = None  # set by _start_operand_refresh when the projection matrix is being cast on the side stream
        self._grads_clean = False
        self._build_parameters()
        from .optim import tag_parameters
        tag_parameters(self)

    # ------------------------------------------------------------------------------------------
    # parameters
    # ------------------------------------------------------------------------------------------
    def _hot_spec(self):
        d, hd, V = self.dim, self.hidden_dim, self.n_words
        spec = [("position_embeddings.weight", (N_MAX_POSITIONS, d), "emb")]
        if self.n_langs > 1:
            spec.append(("cross_lang_embeddings.weight", (self.n_langs, d), "emb"))
        spec += [
            ("layer_norm_emb.weight", (d,), "one"), ("layer_norm_emb.bias", (d,), "zero"),
            ("image_embeddings.image_embeddings.weight", (d, FEAT_DIM), "lin"),
            ("image_embeddings.image_embeddings.bias", (d,), "linb:%d" % FEAT_DIM),
            ("image_embeddings.image_location_embeddings.weight", (d, LOC_DIM), "lin"),
            ("image_embeddings.image_location_embeddings.bias", (d,), "linb:%d" % LOC_DIM),
            ("image_embeddings.LayerNorm.weight", (d,), "one"), ("image_embeddings.LayerNorm.bias", (d,), "zero"),
        ]
        for i in range(self.n_layers):
            a = "attentions.%d." % i
            spec += [(a + "q_lin.weight", (d, d), "lin"), (a + "k_lin.weight", (d, d), "lin"),
                     (a + "v_lin.weight", (d, d), "lin"),
                     (a + "q_lin.bias", (d,), "linb:%d" % d), (a + "k_lin.bias", (d,), "linb:%d" % d),
                     (a + "v_lin.bias", (d,), "linb:%d" % d),
                     (a + "out_lin.weight", (d, d), "lin"), (a + "out_lin.bias", (d,), "linb:%d" % d),
                     ("layer_norm1.%d.weight" % i, (d,), "one"), ("layer_norm1.%d.bias" % i, (d,), "zero"),
                     ("ffns.%d.lin1.weight" % i, (hd, d), "lin"), ("ffns.%d.lin1.bias" % i, (hd,), "linb:%d" % d),
                     ("ffns.%d.lin2.weight" % i, (d, hd), "lin"), ("ffns.%d.lin2.bias" % i, (d,), "linb:%d" % hd),
                     ("layer_norm2.%d.weight" % i, (d,), "one"), ("layer_norm2.%d.bias" % i, (d,), "zero")]
        for pl, sr in (("pooled_layer", "seq_relationship"), ("pooled_layer2", "seq_relationship2")):
            spec += [(pl + ".dense.weight", (d, d), "lin"), (pl + ".dense.bias", (d,), "linb:%d" % d),
                     (sr + ".weight", (1, d), "lin"), (sr + ".bias", (1,), "linb:%d" % d)]
        spec += [("mrfr_dense.weight", (FEAT_DIM, d), "lin"), ("mrfr_dense.bias", (FEAT_DIM,), "linb:%d" % d),
                 ("transformer_obj.dense.weight", (d, d), "lin"), ("transformer_obj.dense.bias", (d,), "linb:%d" % d),
                 ("transformer_obj.LayerNorm.weight", (d,), "one"), ("transformer_obj.LayerNorm.bias", (d,), "zero")]
        if self.with_output:
            spec += [("pred_obj_layer.proj.weight", (N_OBJ, d), "lin"),
                     ("pred_obj_layer.proj.bias", (N_OBJ,), "linb:%d" % d),
                     ("pred_layer.proj.bias", (V,), "linb:%d" % d)]
        return spec

    def _cold_spec(self):
        """Parameters of the reference that never receive a gradient on this path:
        kept (same names / shapes / init family) so checkpoints round-trip; no kernel reads them."""
        d, hd = self.dim, self.hidden_dim
        spec = [("image_embeddings.image_distbution_embeddings", d, N_OBJ)]
        spec += [("cross_alignment.align_output", d, d), ("cross_alignment.att_weight_c", 1, d),
                 ("cross_alignment.att_weight_cq", 1, d), ("cross_alignment.att_weight_q", 1, d),
                 ("cross_alignment.layer_norm", d, None)]
        for i in range(2):
            spec += [("latent_transforms.%d.out_dense" % i, d, 2 * d), ("latent_transforms.%d.x_to_logvar" % i, d, d),
                     ("latent_transforms.%d.x_to_mu" % i, d, d),
                     ("original_transforms.%d.LayerNorm" % i, d, None), ("original_transforms.%d.dense" % i, d, d),
                     ("original_transforms.%d.dense_mu" % i, d, d)]
        for i in range(self.refine_layers):
            r = "refine_embeddings.layers.%d." % i
            spec += [(r + "feed_forward.lin1", hd, d), (r + "feed_forward.lin2", d, hd),
                     (r + "self_attn.aoa_layer.0", 2 * d, 2 * d)]
            spec += [(r + "self_attn.linears.%d" % j, d, d) for j in range(3)]
            spec += [(r + "sublayer.%d.norm" % j, d, None) for j in range(2)]
        spec.append(("refine_embeddings.norm", d, None))
        for i in range(self.n_layers):
            spec.append(("layer_norm15.%d" % i, d, None))
            spec += [("encoder_attn.%d.%s_lin" % (i, n), d, d) for n in ("q", "k", "v", "out")]
        return spec

    def _register(self, name, param):
        parts = name.split(".")
        node = self
        for p in parts[:-1]:
            if not hasattr(node, p):
                if node is self and p == "embeddings":
                    import weakref
                    emb_node = _EmbeddingNode()
                    emb_node.__dict__["_owner"] = weakref.ref(self)
                    node.add_module(p, emb_node)
                else:
                    node.add_module(p, _Node())
            node = getattr(node, p)
        node.register_parameter(parts[-1], param)

    def _build_parameters(self):
        spec = self._hot_spec()
        self._hot_names = [n for n, _, _ in spec]
        self._hot_shapes = {n: s for n, s, _ in spec}
        off = 0
        self._hot_off = {}
        for n, s, _ in spec:
            self._hot_off[n] = off
            off += _align(int(math.prod(s)))
        self._flat_numel = off
        # contiguous slices of the flat buffers, in the order the backward finishes them
        first_layer = self._hot_off["attentions.0.q_lin.weight"] if self.n_layers > 0 else self._hot_off["pooled_layer.dense.weight"]
        heads_lo = self._hot_off["pooled_layer.dense.weight"]
        self._segments = {"embed": (0, first_layer), "heads": (heads_lo, off)}
        for i in range(self.n_layers):
            hi = self._hot_off["attentions.%d.q_lin.weight" % (i + 1)] if i + 1 < self.n_layers else heads_lo
            self._segments["layer%d" % i] = (self._hot_off["attentions.%d.q_lin.weight" % i], hi)
        flat = torch.zeros(off, dtype=_F32)
        with torch.no_grad():
            for n, s, kind in spec:
                v = flat[self._hot_off[n]:self._hot_off[n] + math.prod(s)].view(s)
                if kind == "one":
                    v.fill_(1.0)
                elif kind == "emb":
                    v.normal_(0.0, self.dim ** -0.5)
                elif kind == "lin":
                    _linear_init_(v, None, s[-1])
                elif kind.startswith("linb:"):
                    b = 1.0 / math.sqrt(int(kind[5:]))
                    v.uniform_(-b, b)
            emb = torch.empty(self.n_words, self.dim, dtype=_F32).normal_(0.0, self.dim ** -0.5)
            emb[self.pad_index].zero_()  # transformer.py:21-26
        self._flat = flat
        self._flat16 = None
        self._flat_grad = None
        self._emb16 = None
        self._emb_grad = None
        self._proj_grad = None
        self._hot = {}
        for n, s, _ in spec:
            p = nn.Parameter(flat[self._hot_off[n]:self._hot_off[n] + math.prod(s)].view(s))
            self._hot[n] = p
            self._register(n, p)
        self.__dict__["_emb"] = nn.Parameter(emb)  # plain attribute: registered once, under its reference name
        self._register("embeddings.weight", self._emb)
        if self.with_output:
            if self.share_inout_emb:
                self._register("pred_layer.proj.weight", self._emb)  # tied, transformer.py:728-729
                self.__dict__["_proj"] = self._emb
            else:
                w = torch.empty(self.n_words, self.dim, dtype=_F32)
                _linear_init_(w, None, self.dim)
                self.__dict__["_proj"] = nn.Parameter(w)
                self._register("pred_layer.proj.weight", self._proj)
        else:
            self.__dict__["_proj"] = None
        for name, out_f, in_f in self._cold_spec():
            if in_f is None:  # LayerNorm
                self._register(name + ".weight", nn.Parameter(torch.ones(out_f)))
                self._register(name + ".bias", nn.Parameter(torch.zeros(out_f)))
            else:
                w, b = torch.empty(out_f, in_f), torch.empty(out_f)
                with torch.no_grad():
                    _linear_init_(w, b, in_f)
                self._register(name + ".weight", nn.Parameter(w))
                self._register(name + ".bias", nn.Parameter(b))

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self._reflatten()
        return self

    def _reflatten(self):
        """Re-pack the hot parameters into one flat buffer after .cuda() / .to() moved them."""
        dev = self._emb.device
        for n in self._hot_names:
            if self._hot[n].dtype != _F32:
                raise NotImplementedError("hot parameters stay fp32 masters (bf16 operand copies are internal)")
        flat = torch.zeros(self._flat_numel, dtype=_F32, device=dev)
        with torch.no_grad():
            for n in self._hot_names:
                p = self._hot[n]
                view = flat[self._hot_off[n]:self._hot_off[n] + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                p.grad = None
        self._flat = flat
        self._flat16 = None
        self._flat_grad = None
        self._emb16 = None
        self._emb_grad = None
        self._proj_grad = None
        self._emb_touched, self._emb_dense_dirty = None, True
        self._emb16_event = None
        self.invalidate_operands()
        self._grads_clean = False

    def invalidate_operands(self):
        """Call after editing parameters by hand: the next forward re-casts the bf16 operand copies."""
        self._operands_valid = False
        self._emb16_valid = False

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_operands()
        return out

    # -- views ------------------------------------------------------------------------------------
    def _w32(self, name):
        return self._hot[name].data

    def _view(self, buf, name, numel=None, shape=None):
        o = self._hot_off[name]
        s = self._hot_shapes[name] if shape is None else shape
        n = math.prod(s) if numel is None else numel
        return buf[o:o + n].view(s)

    def _w16(self, name, shape=None):
        return self._view(self._flat16, name, shape=shape)

    def _g(self, name, shape=None):
        return self._view(self._flat_grad, name, shape=shape)

    def hot_parameters(self):
        """(name, parameter) of every parameter this path trains, flat-buffer order, embeddings last."""
        out = [(n, self._hot[n]) for n in self._hot_names]
        out.append(("embeddings.weight", self._emb))
        if self.with_output and not self.share_inout_emb:
            out.append(("pred_layer.proj.weight", self._proj))
        return out

    def refresh_operands(self, embeddings=False):
        """bf16 tensor-core copies of the fp32 masters (one cast kernel over the flat buffer)."""
        ops.use_current_stream()
        if self._flat16 is None or not self._operands_valid:
            if self._flat16 is None:
                self._flat16 = torch.empty(self._flat_numel, dtype=_BF16, device=self._flat.device)
            ops.cast_f32_bf16(self._flat, self._flat16, self._flat_numel)
        if embeddings:
            ev, self._emb16_event = self._emb16_event, None
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)  # cast by _start_operand_refresh on the side stream
            elif self._emb16 is None or not self._emb16_valid:
                if self._emb16 is None:
                    self._emb16 = torch.empty(self._proj.shape, dtype=_BF16, device=self._flat.device)
                ops.cast_f32_bf16(self._proj.data, self._emb16, self._proj.numel())

    def _start_operand_refresh(self):
        """The encoder forward's version of refresh_operands(): the cast of everything the first layer does not need
        (layers 1.., the heads and — if the MLM head ran in the previous step — the V x d projection matrix, 0.55 + 1.15
        GB of pure HBM traffic) goes to the side stream underneath the embedding stage and layer 0; `_encode` waits
        for it before layer 1, the MLM head before its GEMM.  Returns the event to wait for (or None)."""
        ops.use_current_stream()
        dev = self._flat.device
        self._emb16_event = None  # a previous forward's cast finished long ago (its event was joined in _encode)
        need_flat = self._flat16 is None or not self._operands_valid
        need_emb = self.with_output and self._emb16 is not None and not self._emb16_valid
        if not (need_flat or need_emb):
            return None
        if self._flat16 is None:
            self._flat16 = torch.empty(self._flat_numel, dtype=_BF16, device=dev)
        split = self._segments["layer1"][0] if self.n_layers > 1 else self._flat_numel
        if not self.overlap_grads:
            split = self._flat_numel
        if need_flat:
            ops.cast_f32_bf16(self._flat, self._flat16, split)
        ev_rest = None
        if (need_flat and split < self._flat_numel) or need_emb:
            cur, side = torch.cuda.current_stream(), self._side_stream_for(dev)
            fork = torch.cuda.Event()
            fork.record(cur)
            side.wait_event(fork)
            with torch.cuda.stream(side), ops.on_stream(side):
                if need_flat and split < self._flat_numel:
                    ops.cast_f32_bf16(self._flat[split:], self._flat16[split:], self._flat_numel - split)
                    ev_rest = torch.cuda.Event()
                    ev_rest.record(side)
                if need_emb:
                    ops.cast_f32_bf16(self._proj.data, self._emb16, self._proj.numel())
                    self._emb16_event = torch.cuda.Event()
                    self._emb16_event.record(side)
        return ev_rest

    def attach_grads(self, zero=False):
        """Make every hot `param.grad` a view of the flat gradient buffer (kernels accumulate there)."""
        fresh = self._flat_grad is None
        if fresh:
            self._flat_grad = torch.zeros(self._flat_numel, dtype=_F32, device=self._flat.device)
            self._emb_grad = torch.zeros_like(self._emb.data)
            self._emb_touched, self._emb_dense_dirty = [], False
            self._proj_grad = self._emb_grad if self._proj is self._emb else (
                torch.zeros_like(self._proj.data) if self.with_output else None)
        detached = fresh
        for n in self._hot_names:
            p = self._hot[n]
            if p.grad is None:
                detached = True
                p.grad = self._view(self._flat_grad, n)
        if self._emb.grad is None:
            detached = True
            self._emb.grad = self._emb_grad
        if self.with_output and self._proj is not self._emb and self._proj.grad is None:
            detached = True
            self._proj.grad = self._proj_grad
        # zero_grad(set_to_none=True) dropped the views: the buffers still hold the last step's sums
        if ((detached and not fresh) or zero) and not self._grads_clean:
            self._flat_grad.zero_()
            self._zero_emb_grad()
            if self._proj_grad is not None and self._proj_grad is not self._emb_grad:
                self._proj_grad.zero_()
        self._grads_clean = bool(zero or fresh)  # False once a backward is about to accumulate

    def _zero_emb_grad(self):
        """The token-embedding gradient is 68 % of all gradient bytes (V x d fp32) but, without the MLM head,
        only the rows of the tokens in the batch were touched: clear just those instead of a 768 MB memset."""
        touched, self._emb_touched = self._emb_touched, []
        if self._emb_dense_dirty or touched is None:
            self._emb_grad.zero_()
        else:
            for ids in touched:
                self._emb_grad.index_fill_(0, ids.reshape(-1), 0.0)
        self._emb_dense_dirty = False

    def zero_grad(self, set_to_none=False):
        """Zero the flat gradient buffers in two memsets and keep `param.grad` attached."""
        if self._flat_grad is None or set_to_none:
            return super().zero_grad(set_to_none=True)
        self.attach_grads(zero=True)

    # ------------------------------------------------------------------------------------------
    # dispatcher (transformer.py:731-751)
    # ------------------------------------------------------------------------------------------
    def forward(self, mode, **kwargs):
        if mode == 'fwd':
            return self.fwd(**kwargs)
        elif mode == 'crossfwd':
            return self.crossfwd(**kwargs)
        elif mode == 'jointfwd':
            return self.jointfwd(**kwargs)
        elif mode == 'predict':
            return self.predict(**kwargs)
        elif mode in ('ImageEmbed', 'GAN', 'transform'):
            raise NotImplementedError("mode %r is outside the B200 encoder training path (SURVEY.md §2.1 #2)" % mode)
        else:
            raise Exception("Unknown mode: %s" % mode)

    def _drop(self):
        return (self.dropout if self.training else 0.0), (self.attention_dropout if self.training else 0.0)

    _SEED_SLOTS = 8
    _SEED_INC = 0x1E3779B97F4A7C15  # odd: the per-call device word walks through all 2^64 values

    def _next_seed(self):
        """Dropout seeding that survives CUDA-graph replay: the host-side seed is a constant, the per-call
        variation lives in a device word that this call bumps (one tiny kernel, captured with the graph)
        and registers with the library (m3p_set_seed_mix); kernels XOR it in at start.  A ring of words
        lets several forwards be in flight before their backwards (e.g. the CLCM second pass) — the
        backward re-registers the word its forward used.  Returns (host seed, device word tensor)."""
        if self._seed_words is None or self._seed_words.device != self._flat.device:
            # distinct start values: slot j after k bumps holds j * C + k * INC, so no two calls ever share a word
            self._seed_words = torch.arange(self._SEED_SLOTS, dtype=torch.int64, device=self._flat.device) * 0x2545F4914F6CDD1D
        slot = self._step % self._SEED_SLOTS
        self._step += 1
        word = self._seed_words[slot:slot + 1]
        if self.training:
            word.add_(self._SEED_INC)
        L.load().m3p_set_seed_mix(word.data_ptr())
        return self._seed_base, word

    def _side_stream_for(self, dev):
        key = (dev.type, dev.index)
        if key not in self._side_streams:
            self._side_streams[key] = torch.cuda.Stream(device=dev)
        return self._side_streams[key]

    def _require_cuda(self, t):
        if not t.is_cuda:
            raise L.M3PError("m3p_b200 runs on a B200 only: got a %s tensor (there is no CPU fallback; "
                             "move the model and inputs to cuda)" % t.device)

    def jointfwd(self, x, lengths, x_img, lengths_img, causal=False, positions=None, langs=None, image_loc=None,
                 refine_image=False, is_latent=False, text_embed=None, image_prep=None):
        """Image regions ++ text tokens through the encoder — transformer.py:878-968.
        x (T,B) int64, x_img (R,B,2048), image_loc (R,B,5) -> (R+T, B, dim).  `langs`/`positions` are
        accepted and ignored exactly as the reference does (:932-938).
        `image_prep` (extension): dict(normalize=True, zero_mask=(B,R) bool, ori_out=(B,R,2048) fp32 or
        None) — x_img then holds RAW region features and the masking / L2-normalisation the reference's dataset does
        on the host (dataset_pretrain.py:258-292,379) runs inside the cast kernel."""
        if causal or refine_image or is_latent:
            raise NotImplementedError("jointfwd: causal / refine_image / is_latent are outside the B200 path")
        slen, bs = x.size()
        assert lengths.size(0) == bs
        assert image_loc is not None
        self._require_cuda(x)
        spec = dict(kind="joint", B=bs, T=slen, R=x_img.size(0), x=x, lengths=lengths + lengths_img, x_img=x_img,
                    image_loc=image_loc, positions=None, langs=None, image_prep=image_prep,
                    flags=L.M3P_EMB_POS | L.M3P_EMB_MASK_PRE | L.M3P_EMB_LN | L.M3P_EMB_DROP2)
        if image_prep is not None and x_img.requires_grad:
            raise NotImplementedError("image_prep with d x_img (FreeLB) is not supported: perturb the prepared features")
        return _EncoderFn.run(self, spec, x_img, text_embed)

    def fwd(self, x, lengths, causal, src_enc=None, src_len=None, positions=None, langs=None, cache=None,
            enc_mask=None, cross_modal=False, image_loc=None, refine_image=False, refine_encoder=False,
            image_fusion=False, image_enc=None, image_mask=None, image_dist=None):
        """transformer.py:753-876: text (LN_emb(tok+pos) -> dropout -> mask; `langs` unused, :829) or,
        with cross_modal=True, image regions (image_embeddings -> mask)."""
        if causal or src_enc is not None or cache is not None or refine_image or refine_encoder or image_fusion \
                or image_dist is not None:
            raise NotImplementedError("fwd: causal / decoder / cache / refine / fusion are outside the B200 path")
        self._require_cuda(x)
        bs = x.size(1)
        assert lengths.size(0) == bs
        if cross_modal:
            spec = dict(kind="image", B=bs, T=0, R=x.size(0), x=None, lengths=lengths, x_img=x, image_loc=image_loc,
                        positions=None, langs=None, flags=L.M3P_EMB_MASK_POST)
            return _EncoderFn.run(self, spec, x, None)
        slen = x.size(0)
        if positions is not None:
            assert positions.size() == (slen, bs)
        spec = dict(kind="text", B=bs, T=slen, R=0, x=x, lengths=lengths, x_img=None, image_loc=None,
                    positions=positions, langs=None,
                    flags=L.M3P_EMB_POS | L.M3P_EMB_LN | L.M3P_EMB_DROP2 | L.M3P_EMB_MASK_POST)
        return _EncoderFn.run(self, spec, None, None)

    def crossfwd(self, x, lengths, causal, stream_='text', src_enc=None, src_len=None, positions=None, langs=None,
                 cache=None, enc_mask=None, image_loc=None, refine_image=False, refine_encoder=False,
                 image_fusion=False, image_enc=None, image_mask=None, cross_modal=True, image_dist=None,
                 is_latent=False, text_embed=None):
        """transformer.py:970-1114, text stream (+ cross_lang_embeddings when `langs` is given, :1056-1057)."""
        assert stream_ in ['img', 'text']
        if causal or src_enc is not None or cache is not None or refine_image or refine_encoder or image_fusion \
                or image_dist is not None or is_latent:
            raise NotImplementedError("crossfwd: causal / decoder / cache / refine / fusion are outside the B200 path")
        self._require_cuda(x)
        if stream_ == 'img':
            # image stream (:1044-1049): BertImageEmbeddings -> dropout -> mask, no layer_norm_emb
            if langs is not None:
                raise NotImplementedError("crossfwd(stream_='img') with langs: the reference runs it with langs=None "
                                          "('currently we set langs=None', transformer.py:1046); not built")
            assert image_loc is not None and lengths.size(0) == x.size(1)
            spec = dict(kind="image", B=x.size(1), T=0, R=x.size(0), x=None, lengths=lengths, x_img=x, image_loc=image_loc,
                        positions=None, langs=None, flags=L.M3P_EMB_DROP2 | L.M3P_EMB_MASK_POST)
            return _EncoderFn.run(self, spec, x, None)
        slen, bs = x.size()
        assert lengths.size(0) == bs
        if positions is not None:
            assert positions.size() == (slen, bs)
        if langs is not None:
            assert langs.size() == (slen, bs)
            if self.n_langs <= 1:
                raise AttributeError("cross_lang_embeddings does not exist when n_langs == 1 (transformer.py:656)")
        spec = dict(kind="text", B=bs, T=slen, R=0, x=x, lengths=lengths, x_img=None, image_loc=None,
                    positions=positions, langs=langs,
                    flags=L.M3P_EMB_POS | L.M3P_EMB_LN | L.M3P_EMB_DROP2 | L.M3P_EMB_MASK_POST)
        return _EncoderFn.run(self, spec, None, text_embed)

    def predict(self, tensor, pred_mask=None, y=None, get_scores=None, is_obj=False, is_relation=False, is_mrfr=False,
                is_clcm=False):
        """Heads — transformer.py:1183-1214."""
        self._require_cuda(tensor)
        if is_relation:
            return _RelationFn.apply(tensor, self, "pooled_layer", "seq_relationship", self._grad_token())
        if is_clcm:
            return _RelationFn.apply(tensor, self, "pooled_layer2", "seq_relationship2", self._grad_token())
        if is_mrfr:
            return _MrfrFn.apply(tensor, self, self._grad_token())
        if is_obj:
            return _ObjFn.apply(tensor, y, self, bool(get_scores), self._grad_token())
        assert pred_mask is not None and y is not None
        return _MlmFn.apply(tensor, pred_mask, y, self, bool(get_scores), self._grad_token())

    def _grad_token(self):
        """A 0-d tensor that requires grad iff the parameters do: lets the head Functions run their
        backward (which writes parameter gradients straight into the flat buffer) even when the
        encoder output was detached."""
        if torch.is_grad_enabled() and self._emb.requires_grad:
            return self._flat.new_zeros((), requires_grad=True)
        return self._flat.new_zeros(())

    # ------------------------------------------------------------------------------------------
    # encoder engine
    # ------------------------------------------------------------------------------------------
    def _layer_views(self, i):
        d, hd = self.dim, self.hidden_dim
        a = "attentions.%d." % i
        w16, g = self._w16, self._g
        return dict(
            wqkv=w16(a + "q_lin.weight", (3 * d, d)), bqkv=self._view(self._flat, a + "q_lin.bias", shape=(3 * d,)),
            wo=w16(a + "out_lin.weight"), bo=self._w32(a + "out_lin.bias"),
            g1=self._w32("layer_norm1.%d.weight" % i), b1=self._w32("layer_norm1.%d.bias" % i),
            w1=w16("ffns.%d.lin1.weight" % i), bb1=self._w32("ffns.%d.lin1.bias" % i),
            w2=w16("ffns.%d.lin2.weight" % i), bb2=self._w32("ffns.%d.lin2.bias" % i),
            g2=self._w32("layer_norm2.%d.weight" % i), b2=self._w32("layer_norm2.%d.bias" % i),
            name=a, idx=i)

    def _layer_grads(self, i):
        d = self.dim
        a = "attentions.%d." % i
        g = self._g
        return dict(
            wqkv=g(a + "q_lin.weight", (3 * d, d)), bqkv=g(a + "q_lin.bias", (3 * d,)),
            wo=g(a + "out_lin.weight"), bo=g(a + "out_lin.bias"),
            g1=g("layer_norm1.%d.weight" % i), b1=g("layer_norm1.%d.bias" % i),
            w1=g("ffns.%d.lin1.weight" % i), bb1=g("ffns.%d.lin1.bias" % i),
            w2=g("ffns.%d.lin2.weight" % i), bb2=g("ffns.%d.lin2.bias" % i),
            g2=g("layer_norm2.%d.weight" % i), b2=g("layer_norm2.%d.bias" % i))

    def _encode(self, spec, x_img, text_embed, need_grad):
        """Forward of the embedding stage + layer loop.  Returns (h [B*S, d] bf16, stash)."""
        ops.use_current_stream()
        if not self._device_checked:
            ops.device_check()
            self._device_checked = True
        dev = self._flat.device
        B, T, R, d, H = spec["B"], spec["T"], spec["R"], self.dim, self.n_heads
        S = R + T
        M = B * S
        if S > 256:
            raise NotImplementedError("sequence length %d > 256 is not supported by the fused attention kernel" % S)
        p_drop, p_att = self._drop()
        seed, seed_word = self._next_seed()
        ev_operands = self._start_operand_refresh()
        e = lambda *s, dt=_BF16: torch.empty(*s, dtype=dt, device=dev)
        seqlen = spec["lengths"].to(device=dev, dtype=torch.int32).contiguous()
        st = dict(spec=spec, seqlen=seqlen, seed=seed, seed_word=seed_word, p_drop=p_drop, p_att=p_att, B=B, T=T, R=R,
                  S=S, M=M, need_grad=need_grad, layers=[])

        # ---- embedding stage (transformer.py:897-943 / 820-831 / 1044-1062) ----
        a = L.EmbedArgs()
        a.B, a.R, a.T, a.d = B, R, T, d
        a.flags, a.eps, a.drop_p = spec["flags"], LN_EPS, p_drop
        a.seed_img, a.seed_emb = seed ^ 0x1111, seed ^ 0x2222
        keep = []
        if R > 0:
            xi = x_img.detach()
            if xi.dtype != _F32 or not xi.is_contiguous():
                xi = xi.to(_F32).contiguous()
            ximg16 = e(B * R, FEAT_DIM)
            prep = spec.get("image_prep")
            if prep is None:
                ops.permute_cast(xi, ximg16, R, B, FEAT_DIM)  # (R,B,F) fp32 -> (B,R,F) bf16
            else:
                # raw region features: zero the masked regions, L2-normalise, cast + permute in one pass (8f3)
                zm = prep.get("zero_mask")
                zm = None if zm is None else zm.to(device=dev, dtype=torch.uint8).contiguous()
                ori = prep.get("ori_out")
                ops.region_prep(xi, zm, prep.get("normalize", True), ximg16, ori, R, B, FEAT_DIM)
            e_img = e(B * R, d, dt=_F32)
            ops.linear(ximg16, self._w16("image_embeddings.image_embeddings.weight"),
                       self._w32("image_embeddings.image_embeddings.bias"), e_img, out_f32=True)
            loc = spec["image_loc"].detach().to(_F32).contiguous()
            img_mean, img_rstd = e(B * R, dt=_F32), e(B * R, dt=_F32)
            a.e_img, a.image_loc = e_img.data_ptr(), loc.data_ptr()
            a.w_loc = self._w32("image_embeddings.image_location_embeddings.weight").data_ptr()
            a.b_loc = self._w32("image_embeddings.image_location_embeddings.bias").data_ptr()
            a.ln_img_g = self._w32("image_embeddings.LayerNorm.weight").data_ptr()
            a.ln_img_b = self._w32("image_embeddings.LayerNorm.bias").data_ptr()
            a.img_mean, a.img_rstd = img_mean.data_ptr(), img_rstd.data_ptr()
            st.update(ximg16=ximg16, e_img=e_img, loc=loc, img_mean=img_mean, img_rstd=img_rstd)
        if T > 0:
            xt = spec["x"].contiguous()
            a.x = xt.data_ptr()
            a.tok_emb = self._emb.data.data_ptr()
            keep.append(xt)
            st["x"] = xt
            if text_embed is not None:
                te = text_embed.detach().to(_F32).contiguous()
                assert te.shape == (B, T, d)
                a.text_embed = te.data_ptr()
                keep.append(te)
            if spec["positions"] is not None:
                pos = spec["positions"].contiguous()
                a.positions = pos.data_ptr()
                st["positions"] = pos
            if spec["langs"] is not None:
                lg = spec["langs"].contiguous()
                a.langs = lg.data_ptr()
                a.lang_emb = self._w32("cross_lang_embeddings.weight").data_ptr()
                st["langs"] = lg
        a.pos_emb = self._w32("position_embeddings.weight").data_ptr()
        a.seqlen = seqlen.data_ptr()
        a.ln_emb_g = self._w32("layer_norm_emb.weight").data_ptr()
        a.ln_emb_b = self._w32("layer_norm_emb.bias").data_ptr()
        # residual stream: the pre-LayerNorm sums x1 / x2 are fp32; a LayerNorm kernel only writes the bf16 copy the
        # tensor cores read (h), and the residual add of the NEXT linear recomputes the fp32 LayerNorm output from the
        # stored sum, mean, rstd, gamma, beta in its GEMM epilogue (res_ln).  Only the embedding output exists as an
        # fp32 tensor (h32).  Nothing on the residual path is ever rounded to bf16 (the reference adds in fp32,
        # transformer.py:951-952,956).
        h, h32 = e(M, d), e(M, d, dt=_F32)
        if spec["flags"] & L.M3P_EMB_LN:
            y_pre, emb_mean, emb_rstd = e(M, d, dt=_F32), e(M, dt=_F32), e(M, dt=_F32)
            a.y_pre, a.emb_mean, a.emb_rstd = y_pre.data_ptr(), emb_mean.data_ptr(), emb_rstd.data_ptr()
            st.update(y_pre=y_pre, emb_mean=emb_mean, emb_rstd=emb_rstd)
        a.h0, a.h0_f32 = h.data_ptr(), h32.data_ptr()
        ops.embed_fwd(a)
        del keep

        # ---- layer loop (transformer.py:947-958) ----
        scale = 1.0 / math.sqrt(d // H)
        res, res_ln = h32, None  # what the next residual add reads: an fp32 tensor, or (pre-LN sum, LayerNorm recipe)
        recompute = os.environ.get("M3P_RES_LN", "1") != "0"  # 0: LayerNorm kernels also write an fp32 copy (A/B runs)
        for i in range(self.n_layers):
            if i == 1 and ev_operands is not None:
                torch.cuda.current_stream().wait_event(ev_operands)  # bf16 copies of layers 1.. and the heads are ready
                ev_operands = None
            w = self._layer_views(i)
            s1, s2, sa = seed ^ (0x100 * (i + 1) + 1), seed ^ (0x100 * (i + 1) + 2), seed ^ (0x100 * (i + 1) + 3)
            qkv = e(M, 3 * d)
            ops.linear(h, w["wqkv"], w["bqkv"], qkv)
            ctx, lse = e(M, d), e(B * H * S, dt=_F32)
            ops.attention_fwd(qkv, seqlen, B, S, H, scale, p_att, sa, ctx, lse)
            x1 = e(M, d, dt=_F32)
            ops.linear(ctx, w["wo"], w["bo"], x1, epi=L.M3P_EPI_DROP_RES, aux=res, aux_ln=res_ln, drop_p=p_drop, seed=s1,
                       out_f32=True)
            h1, mean1, rstd1 = e(M, d), e(M, dt=_F32), e(M, dt=_F32)
            h1_32 = None if recompute else e(M, d, dt=_F32)
            ops.layernorm_fwd(x1, w["g1"], w["b1"], h1, mean1, rstd1, LN_EPS, y32=h1_32)
            gp, g = e(M, 4 * d), e(M, 4 * d)  # gelu'(u) (stash for the backward) and gelu(u)
            ops.linear(h1, w["w1"], w["bb1"], gp, epi=L.M3P_EPI_GELU, out2=g)
            x2 = e(M, d, dt=_F32)
            if recompute:
                ops.linear(g, w["w2"], w["bb2"], x2, epi=L.M3P_EPI_DROP_RES, aux=x1,
                           aux_ln=(mean1, rstd1, w["g1"], w["b1"], None, 0), drop_p=p_drop, seed=s2, out_f32=True)
            else:
                ops.linear(g, w["w2"], w["bb2"], x2, epi=L.M3P_EPI_DROP_RES, aux=h1_32, drop_p=p_drop, seed=s2, out_f32=True)
            hn, mean2, rstd2 = e(M, d), e(M, dt=_F32), e(M, dt=_F32)
            hn32 = None if (recompute or i + 1 == self.n_layers) else e(M, d, dt=_F32)
            ops.layernorm_fwd(x2, w["g2"], w["b2"], hn, mean2, rstd2, LN_EPS, seqlen=seqlen, S=S, y32=hn32)
            res, res_ln = (x2, (mean2, rstd2, w["g2"], w["b2"], seqlen, S)) if recompute else (hn32, None)
            if need_grad:
                st["layers"].append(dict(h=h, qkv=qkv, ctx=ctx, lse=lse, x1=x1, h1=h1, mean1=mean1, rstd1=rstd1, gp=gp,
                                         g=g, x2=x2, mean2=mean2, rstd2=rstd2, s1=s1, s2=s2, sa=sa))
            h = hn
        if ev_operands is not None:
            torch.cuda.current_stream().wait_event(ev_operands)
        if self._emb16_event is not None:  # join the side stream here (free: the cast ended layers ago) so that a step
            torch.cuda.current_stream().wait_event(self._emb16_event)  # without the MLM head leaves nothing unjoined
        return h, st

    def _encode_backward(self, st, dh, want_dximg, want_dtext):
        """Backward of `_encode`: dh [B*S, d] bf16 -> parameter gradients (accumulated into the flat
        buffer) and, on request, d x_img (R,B,2048) / d text_embed (B,T,d) for FreeLB.  The residual-gradient chain
        (dh -> dx2 -> dh1 -> dx1 -> dh of the layer below) is fp32 end to end, like autograd's in the reference; the
        branch gradients that feed tensor-core GEMMs (dx*d, du, dctx, dqkv) are bf16 operand copies."""
        ops.use_current_stream()
        L.load().m3p_set_seed_mix(st["seed_word"].data_ptr())  # the masks this forward drew
        self.attach_grads()
        dev = self._flat.device
        B, T, R, S, M, d, H = st["B"], st["T"], st["R"], st["S"], st["M"], self.dim, self.n_heads
        p_drop, p_att, seqlen = st["p_drop"], st["p_att"], st["seqlen"]
        e = lambda *s, dt=_BF16: torch.empty(*s, dtype=dt, device=dev)
        scale = 1.0 / math.sqrt(d // H)
        hook = self._grad_ready_hook
        if hook is not None:
            hook("heads", *self._segments["heads"])  # head backward ran before the encoder's
        # The activation-gradient chain (LN rows -> dgrad -> dgrad -> LN rows -> dgrad -> attention -> dgrad) runs on
        # the current stream; everything that only produces PARAMETER gradients (wgrad GEMMs, LN column pass, bias
        # column sums) goes to a side stream, so HBM-bound reductions run underneath tensor-bound GEMMs and the
        # persistent GEMMs of one stream fill the partially-empty last wave of the other's.
        sq = _SideQueue(self._side_stream_for(dev) if self.overlap_grads else None, hook)
        for i in reversed(range(self.n_layers)):
            w, gr, s = self._layer_views(i), self._layer_grads(i), st["layers"][i]
            # layer_norm2 (+ row mask) and the FFN dropout           (:956-958, :226)
            dx2, dx2d = e(M, d, dt=_F32), e(M, d)
            # col_scratch: the row pass also accumulates the column sums into per-CTA partials (one read of dh / x2);
            # the "cols" call on the side stream only adds the partials into the parameter gradients
            scr2, scr1 = e(512 * 3 * d, dt=_F32), e(512 * 3 * d, dt=_F32)
            ln2 = dict(seqlen=seqlen, S=S, dx_drop=dx2d, dx_drop_p=p_drop, dx_seed=s["s2"], dgamma=gr["g2"],
                       dbeta=gr["b2"], dbias=gr["bb2"], col_scratch=scr2)
            ops.layernorm_bwd(dh, s["x2"], s["mean2"], s["rstd2"], w["g2"], dx2, phase="rows", **ln2)
            dx2d_ = dx2d
            sq.fork()
            sq.run(lambda: ops.layernorm_bwd(dh, s["x2"], s["mean2"], s["rstd2"], w["g2"], dx2, phase="cols", **ln2))
            sq.run(lambda: ops.wgrad(dx2d_, s["g"], gr["w2"]))                  # lin2 wgrad            (:225)
            # lin2 dgrad fused with gelu'(u)                                   (:224-225)
            du = e(M, 4 * d)
            ops.dgrad(dx2d_, w["w2"], du, epi=L.M3P_EPI_DGELU, aux=s["gp"], colsum=gr["bb1"])  # + lin1 bias grad
            sq.fork()
            sq.run(lambda: ops.wgrad(du, s["h1"], gr["w1"]))                    # lin1 wgrad            (:223)
            # lin1 dgrad + residual branch                                     (:223, :956)
            dh1 = e(M, d, dt=_F32)
            ops.dgrad(du, w["w1"], dh1, epi=L.M3P_EPI_DROP_RES, aux=dx2, out_f32=True)
            # layer_norm1 and the attention-output dropout              (:951-953)
            dx1, dx1d = e(M, d, dt=_F32), e(M, d)
            ln1 = dict(dx_drop=dx1d, dx_drop_p=p_drop, dx_seed=s["s1"], dgamma=gr["g1"], dbeta=gr["b1"], dbias=gr["bo"],
                       col_scratch=scr1)
            ops.layernorm_bwd(dh1, s["x1"], s["mean1"], s["rstd1"], w["g1"], dx1, phase="rows", **ln1)
            dx1d_ = dx1d
            sq.fork()
            sq.run(lambda: ops.layernorm_bwd(dh1, s["x1"], s["mean1"], s["rstd1"], w["g1"], dx1, phase="cols", **ln1))
            sq.run(lambda: ops.wgrad(dx1d_, s["ctx"], gr["wo"]))
            dctx = e(M, d)
            ops.dgrad(dx1d_, w["wo"], dctx)
            # attention core                                            (:197-205)
            dqkv = e(M, 3 * d)
            ops.attention_bwd(s["qkv"], seqlen, B, S, H, scale, p_att, s["sa"], s["ctx"], s["lse"], dctx, dqkv)
            sq.fork()
            sq.run(lambda: ops.colsum(dqkv, gr["bqkv"]))
            sq.run(lambda: ops.wgrad(dqkv, s["h"], gr["wqkv"]))                 # q/k/v projections     (:178-181)
            dhp = e(M, d, dt=_F32)
            ops.dgrad(dqkv, w["wqkv"], dhp, epi=L.M3P_EPI_DROP_RES, aux=dx1, out_f32=True)
            # the side stream may still be reading these: keep them alive until the join one layer later
            sq.close("layer%d" % i, self._segments["layer%d" % i], (dh, s, dx2, dx2d, du, dh1, dx1, dx1d, dqkv, scr1, scr2))
            if self._bwd_trace is not None:  # tests only: every intermediate of this layer's backward, for stage parity
                self._bwd_trace.append(dict(layer=i, stash=s, dh=dh, dx2=dx2, dx2d=dx2d, du=du, dh1=dh1, dx1=dx1,
                                            dx1d=dx1d, dctx=dctx, dqkv=dqkv, dhp=dhp))
            dh = dhp
            st["layers"][i] = None
            del s, du, dx2, dx2d, dh1, dx1, dx1d, dqkv, dctx
        sq.join_all()
        # ---- embedding stage ----
        spec = st["spec"]
        flags = spec["flags"]
        d_ximg = d_text = None
        if flags & L.M3P_EMB_LN:
            dy_pre = e(M, d, dt=_F32)
            ops.layernorm_bwd(dh, st["y_pre"], st["emb_mean"], st["emb_rstd"], self._w32("layer_norm_emb.weight"), dy_pre,
                              seqlen=seqlen, S=S, dy_drop_p=p_drop if (flags & L.M3P_EMB_DROP2) else 0.0,
                              dy_seed=st["seed"] ^ 0x2222, dgamma=self._g("layer_norm_emb.weight"),
                              dbeta=self._g("layer_norm_emb.bias"), col_scratch=e(512 * 3 * d, dt=_F32))
        else:
            dy_pre = dh.to(_F32)  # fwd(cross_modal=True): image rows only, mask already applied upstream
        b = L.EmbedBwdArgs()
        b.B, b.R, b.T, b.d, b.flags = B, R, T, d, flags
        b.dy_pre, b.seqlen = dy_pre.data_ptr(), seqlen.data_ptr()
        b.pad_index = self.pad_index
        b.drop_p, b.seed_emb = p_drop, st["seed"] ^ 0x2222
        if T > 0:
            b.x = st["x"].data_ptr()
            if want_dtext:
                d_text = e(B, T, d, dt=_F32)
                b.d_text_embed = d_text.data_ptr()
            elif self._defer_token_grads and self._emb_dense_dirty:
                # data-parallel step with the tied MLM head: the dense head contribution to d E is already being
                # all-reduced (ddp.GradReducer sent it when the heads finished), so the gather's contribution must not
                # be scattered into that buffer now — keep it per position and let the reducer exchange it as rows
                # (position / language gradients are routed to their own tables independently of d_text_embed, so
                # reset `positions` — the xMLM / TLM step — take this path too)
                g_pos = e(B, T, d, dt=_F32)
                b.d_text_embed = g_pos.data_ptr()
                self._deferred_token_grads.append((st["x"], g_pos))
            else:
                b.d_tok_emb = self._emb_grad.data_ptr()
                if self._emb_touched is not None:
                    self._emb_touched.append(st["x"])
            if "positions" in st:
                b.positions = st["positions"].data_ptr()
            if "langs" in st:
                b.langs = st["langs"].data_ptr()
                b.d_lang_emb = self._g("cross_lang_embeddings.weight").data_ptr()
        if flags & L.M3P_EMB_POS:
            b.d_pos_emb = self._g("position_embeddings.weight").data_ptr()
        if R > 0:
            dy_img = e(B * R, d, dt=_F32)
            b.dy_img = dy_img.data_ptr()
        ops.embed_bwd_route(b)
        if R > 0:
            de = e(B * R, d)
            ops.layernorm_bwd(dy_img, st["e_img"], st["img_mean"], st["img_rstd"],
                              self._w32("image_embeddings.LayerNorm.weight"), de, dy_drop_p=p_drop,
                              dy_seed=st["seed"] ^ 0x1111, dgamma=self._g("image_embeddings.LayerNorm.weight"),
                              dbeta=self._g("image_embeddings.LayerNorm.bias"),
                              dbias=self._g("image_embeddings.image_embeddings.bias"), col_scratch=e(512 * 3 * d, dt=_F32))
            # the three parameter gradients that hang off `de` are independent: the location-embedding ones (two
            # small HBM-bound kernels) run on the side stream underneath the projection's weight-gradient GEMM
            sq.fork()
            sq.run(lambda: ops.colsum(de, self._g("image_embeddings.image_location_embeddings.bias")))
            sq.run(lambda: ops.loc_wgrad(de, st["loc"], self._g("image_embeddings.image_location_embeddings.weight"), B, R, d))
            ops.wgrad(de, st["ximg16"], self._g("image_embeddings.image_embeddings.weight"))
            if want_dximg:
                dxi = e(B * R, FEAT_DIM)
                ops.dgrad(de, self._w16("image_embeddings.image_embeddings.weight"), dxi)
                d_ximg = dxi.view(B, R, FEAT_DIM).transpose(0, 1).to(_F32)
            sq.close("embed_img", None, (de,))
            sq.join_all()
        if hook is not None:
            hook("embed", *self._segments["embed"])
        return d_ximg, d_text


class _SideQueue:
    """Fork/join bookkeeping for the side stream of `_encode_backward` (works eagerly and under CUDA-graph
    capture, where the events become graph edges).  `fork()` makes the side stream wait for what the main
    stream has enqueued so far; `run(fn)` enqueues fn's kernels on the side stream; `close()` marks a group
    (one layer) and defers its join by one group so the main chain never waits for the side stream's most
    recent work; `join_all()` makes the main stream wait for everything.  With stream=None everything runs
    inline on the main stream."""

    def __init__(self, stream, hook):
        self.side, self.hook = stream, hook
        self.main = torch.cuda.current_stream() if stream is not None else None
        self.pending = []

    def fork(self):
        if self.side is not None:
            ev = torch.cuda.Event()
            ev.record(self.main)
            self.side.wait_event(ev)

    def run(self, fn):
        if self.side is None:
            fn()
        else:
            with ops.on_stream(self.side):
                fn()

    def _join(self, upto):
        while len(self.pending) > upto:
            ev, name, seg, keep = self.pending.pop(0)
            if ev is not None:
                self.main.wait_event(ev)
            del keep
            if self.hook is not None and seg is not None:  # seg None: a group with no gradient segment of its own
                self.hook(name, *seg)

    def close(self, name, seg, keep):
        ev = None
        if self.side is not None:
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.pending.append((ev, name, seg, keep if self.side is not None else None))
        self._join(1 if self.side is not None else 0)

    def join_all(self):
        self._join(0)


class _EncoderFn(torch.autograd.Function):
    """jointfwd / fwd / crossfwd as one autograd node: forward and backward are kernel sequences."""

    @staticmethod
    def run(model, spec, x_img, text_embed):
        need = torch.is_grad_enabled() and model._emb.requires_grad
        token = model._flat.new_zeros((), requires_grad=need)
        xi = x_img if x_img is not None else model._flat.new_zeros(())
        te = text_embed if text_embed is not None else model._flat.new_zeros(())
        return _EncoderFn.apply(token, xi, te, model, spec, x_img is not None, text_embed is not None)

    @staticmethod
    def forward(ctx, token, x_img, text_embed, model, spec, has_img, has_text_embed):
        need = ctx.needs_input_grad[0] or (has_img and ctx.needs_input_grad[1]) or \
            (has_text_embed and ctx.needs_input_grad[2])
        h, st = model._encode(spec, x_img if has_img else None, text_embed if has_text_embed else None, need)
        ctx.model, ctx.st = model, st
        ctx.has_img, ctx.has_text_embed = has_img, has_text_embed
        return h.view(spec["B"], st["S"], model.dim).transpose(0, 1)  # (slen, bs, dim), :964

    @staticmethod
    def backward(ctx, dout):
        model, st = ctx.model, ctx.st
        dh = dout.transpose(0, 1).to(_BF16).contiguous().view(st["M"], model.dim)
        want_img = ctx.has_img and ctx.needs_input_grad[1]
        want_txt = ctx.has_text_embed and ctx.needs_input_grad[2]
        d_ximg, d_text = model._encode_backward(st, dh, want_img, want_txt)
        ctx.st = None
        return None, d_ximg, d_text, None, None, None, None


class _EmbedLookupFn(torch.autograd.Function):
    """Token-embedding lookup on its own, fp32 in / fp32 out; backward scatter-adds into the flat embedding gradient
    (padding_idx rows get none, as nn.Embedding's)."""

    @staticmethod
    def forward(ctx, weight, ids, model):
        model._require_cuda(ids)
        ops.use_current_stream()
        idx = ids.contiguous()
        out = torch.empty(*idx.shape, weight.shape[1], dtype=_F32, device=weight.device)
        ops.gather_rows_f32(weight.data, idx.view(-1), out, idx.numel(), weight.shape[1])
        ctx.model, ctx.idx = model, idx
        return out

    @staticmethod
    def backward(ctx, g):
        model, idx = ctx.model, ctx.idx
        ops.use_current_stream()
        model.attach_grads()
        gc = g.to(_F32).contiguous()
        ops.scatter_add_rows_f32(gc, idx.view(-1), model.pad_index, model._emb_grad, idx.numel(), gc.shape[-1])
        if model._emb_touched is not None:
            model._emb_touched.append(idx)
        return None, None, None


def _rows_of(tensor):
    """(n_outer, n_inner, d) view -> (base tensor, n_inner, stride_outer, stride_inner) for the row gather."""
    assert tensor.dim() == 3 and tensor.stride(2) == 1 and tensor.dtype == _BF16
    return tensor, tensor.size(1), tensor.stride(0), tensor.stride(1)


class _RelationFn(torch.autograd.Function):
    """BertPooler + seq_relationship (ITM / CLCM head) — transformer.py:546-558, 713-716, 1195-1201."""

    @staticmethod
    def forward(ctx, tensor, model, pl, sr, token):
        ops.use_current_stream()
        dev = tensor.device
        t = tensor if tensor.dtype == _BF16 else tensor.to(_BF16)
        B, S, d = t.shape
        idx = torch.arange(B, device=dev, dtype=torch.int64)
        first = torch.empty(B, d, dtype=_BF16, device=dev)
        ops.gather_rows(t, idx, 1, t.stride(0), 0, first, B, d)  # hidden_states[:, 0]
        pooled = torch.empty(B, d, dtype=_BF16, device=dev)
        ops.linear(first, model._w16(pl + ".dense.weight"), model._w32(pl + ".dense.bias"), pooled, epi=L.M3P_EPI_TANH)
        scores = torch.empty(B, dtype=_F32, device=dev)
        ops.rowdot_fwd(pooled, model._w32(sr + ".weight"), model._w32(sr + ".bias"), scores)
        ctx.model, ctx.names = model, (pl, sr)
        ctx.shape = (B, S, d)
        ctx.save_for_backward(first, pooled, idx)
        return scores.view(B, 1)

    @staticmethod
    def backward(ctx, dscores):
        ops.use_current_stream()
        model = ctx.model
        pl, sr = ctx.names
        first, pooled, idx = ctx.saved_tensors
        B, S, d = ctx.shape
        dev = first.device
        model.attach_grads()
        ds = dscores.reshape(B).to(_F32).contiguous()
        dpre = torch.empty(B, d, dtype=_BF16, device=dev)
        ops.rowdot_bwd(ds, pooled, model._w32(sr + ".weight"), dpre, model._g(sr + ".weight"), model._g(sr + ".bias"),
                       tanh_grad=True)
        dfirst = torch.empty(B, d, dtype=_BF16, device=dev)
        ops.dgrad(dpre, model._w16(pl + ".dense.weight"), dfirst)
        ops.wgrad(dpre, first, model._g(pl + ".dense.weight"))
        ops.colsum(dpre, model._g(pl + ".dense.bias"))
        dt = None
        if ctx.needs_input_grad[0]:
            dt = torch.zeros(B, S, d, dtype=_BF16, device=dev)
            ops.scatter_rows(dfirst, idx, 1, S * d, 0, dt, B, d)
        return dt, None, None, None, None


class _MrfrFn(torch.autograd.Function):
    """mrfr_dense — transformer.py:718,1202-1204: (B,R,d) -> (B,R,2048)."""

    @staticmethod
    def forward(ctx, tensor, model, token):
        ops.use_current_stream()
        dev = tensor.device
        t = tensor if tensor.dtype == _BF16 else tensor.to(_BF16)
        B, R, d = t.shape
        idx = torch.arange(B * R, device=dev, dtype=torch.int64)
        rows = torch.empty(B * R, d, dtype=_BF16, device=dev)
        ops.gather_rows(t, idx, R, t.stride(0), t.stride(1), rows, B * R, d)
        out = torch.empty(B * R, FEAT_DIM, dtype=_BF16, device=dev)
        ops.linear(rows, model._w16("mrfr_dense.weight"), model._w32("mrfr_dense.bias"), out)
        ctx.model, ctx.shape = model, (B, R, d)
        ctx.save_for_backward(rows)
        return out.view(B, R, FEAT_DIM)

    @staticmethod
    def backward(ctx, dout):
        ops.use_current_stream()
        model = ctx.model
        (rows,) = ctx.saved_tensors
        B, R, d = ctx.shape
        model.attach_grads()
        do = dout.reshape(B * R, FEAT_DIM).to(_BF16).contiguous()
        ops.wgrad(do, rows, model._g("mrfr_dense.weight"))
        ops.colsum(do, model._g("mrfr_dense.bias"))
        dt = None
        if ctx.needs_input_grad[0]:
            dt = torch.empty(B * R, d, dtype=_BF16, device=rows.device)
            ops.dgrad(do, model._w16("mrfr_dense.weight"), dt)
            dt = dt.view(B, R, d)
        return dt, None, None


class _ObjFn(torch.autograd.Function):
    """BertPredictionHeadTransform + ObjPredLayer (MRM head) — transformer.py:595-606, 576-584."""

    @staticmethod
    def forward(ctx, tensor, y, model, get_scores, token):
        ops.use_current_stream()
        dev = tensor.device
        t = tensor if tensor.dtype == _BF16 else tensor.to(_BF16)
        B, R, d = t.shape
        n = B * R
        idx = torch.arange(n, device=dev, dtype=torch.int64)
        rows = torch.empty(n, d, dtype=_BF16, device=dev)
        ops.gather_rows(t, idx, R, t.stride(0), t.stride(1), rows, n, d)
        gp, g = torch.empty(n, d, dtype=_BF16, device=dev), torch.empty(n, d, dtype=_BF16, device=dev)
        ops.linear(rows, model._w16("transformer_obj.dense.weight"), model._w32("transformer_obj.dense.bias"), gp,
                   epi=L.M3P_EPI_GELU, out2=g)
        tn, mean, rstd = torch.empty(n, d, dtype=_BF16, device=dev), torch.empty(n, dtype=_F32, device=dev), \
            torch.empty(n, dtype=_F32, device=dev)
        ops.layernorm_fwd(g, model._w32("transformer_obj.LayerNorm.weight"), model._w32("transformer_obj.LayerNorm.bias"),
                          tn, mean, rstd, LN_EPS)
        logits = torch.empty(n, N_OBJ, dtype=_BF16, device=dev)
        ops.linear(tn, model._w16("pred_obj_layer.proj.weight"), model._w32("pred_obj_layer.proj.bias"), logits)
        yv = y.reshape(-1).contiguous()
        loss, lse, inv = torch.empty((), dtype=_F32, device=dev), torch.empty(n, dtype=_F32, device=dev), \
            torch.empty((), dtype=_F32, device=dev)
        ops.cross_entropy_fwd(logits, yv, N_OBJ, -1, loss, lse, inv)
        ctx.model, ctx.shape = model, (B, R, d)
        ctx.save_for_backward(rows, gp, g, tn, mean, rstd, logits, yv, lse, inv)
        scores = logits.float() if get_scores else logits
        ctx.mark_non_differentiable(scores)
        return scores, loss

    @staticmethod
    def backward(ctx, _dscores, dloss):
        ops.use_current_stream()
        model = ctx.model
        rows, gp, g, tn, mean, rstd, logits, yv, lse, inv = ctx.saved_tensors
        B, R, d = ctx.shape
        n, dev = B * R, rows.device
        model.attach_grads()
        gs = dloss.to(_F32).contiguous()
        dlog = torch.empty_like(logits)
        ops.cross_entropy_bwd(logits, yv, N_OBJ, -1, lse, inv, gs, dlog)
        ops.wgrad(dlog, tn, model._g("pred_obj_layer.proj.weight"))
        ops.colsum(dlog, model._g("pred_obj_layer.proj.bias"))
        dtn = torch.empty(n, d, dtype=_BF16, device=dev)
        ops.dgrad(dlog, model._w16("pred_obj_layer.proj.weight"), dtn)
        dg = torch.empty(n, d, dtype=_BF16, device=dev)
        ops.layernorm_bwd(dtn, g, mean, rstd, model._w32("transformer_obj.LayerNorm.weight"), dg,
                          dgamma=model._g("transformer_obj.LayerNorm.weight"),
                          dbeta=model._g("transformer_obj.LayerNorm.bias"))
        du = torch.empty(n, d, dtype=_BF16, device=dev)
        ops.gelu_bwd(dg, gp, du)
        ops.wgrad(du, rows, model._g("transformer_obj.dense.weight"))
        ops.colsum(du, model._g("transformer_obj.dense.bias"))
        dt = None
        if ctx.needs_input_grad[0]:
            dt = torch.empty(n, d, dtype=_BF16, device=dev)
            ops.dgrad(du, model._w16("transformer_obj.dense.weight"), dt)
            dt = dt.view(B, R, d)
        return dt, None, None, None, None


class _MlmFn(torch.autograd.Function):
    """PredLayer on the rows selected by pred_mask (MLM head) — transformer.py:104-117, 1206-1212."""

    @staticmethod
    def forward(ctx, tensor, pred_mask, y, model, get_scores, token):
        ops.use_current_stream()
        dev = tensor.device
        t = tensor if tensor.dtype == _BF16 else tensor.to(_BF16)
        slen, bs, d = t.shape
        V = model.n_words
        n = int(y.numel())  # == pred_mask.sum() by the reference's contract (:1189-1190); no device sync
        # seq-major order of tensor[pred_mask] (:1206)
        idx = torch.nonzero_static(pred_mask.reshape(-1), size=n).reshape(-1)
        rows = torch.empty(n, d, dtype=_BF16, device=dev)
        ops.gather_rows(t, idx, bs, t.stride(0), t.stride(1), rows, n, d)
        model.refresh_operands(embeddings=True)
        ldv = (V + 7) // 8 * 8
        logits = torch.empty(n, ldv, dtype=_BF16, device=dev)
        ops.gemm(rows, model._emb16, n, V, d, logits, bias=model._w32("pred_layer.proj.bias"))
        yv = y.contiguous()
        loss, lse, inv = torch.empty((), dtype=_F32, device=dev), torch.empty(n, dtype=_F32, device=dev), \
            torch.empty((), dtype=_F32, device=dev)
        ops.cross_entropy_fwd(logits, yv, V, -100, loss, lse, inv)
        ctx.model, ctx.shape, ctx.inplace = model, (slen, bs, d, n), not get_scores
        ctx.save_for_backward(rows, logits, yv, lse, inv, idx)
        scores = logits[:, :V].float() if get_scores else logits[:, :V]  # bf16 view is only valid until backward
        ctx.mark_non_differentiable(scores)
        return scores, loss

    @staticmethod
    def backward(ctx, _dscores, dloss):
        ops.use_current_stream()
        model = ctx.model
        rows, logits, yv, lse, inv, idx = ctx.saved_tensors
        slen, bs, d, n = ctx.shape
        V, dev = model.n_words, rows.device
        model.attach_grads()
        gs = dloss.to(_F32).contiguous()
        dlog = logits if ctx.inplace else torch.empty_like(logits)
        ops.cross_entropy_bwd(logits, yv, V, -100, lse, inv, gs, dlog)  # in place: logits -> dlogits
        # dE[V][d] += dlogits^T rows  (tied with the input embedding gradient, transformer.py:728-729)
        if model._proj_grad is model._emb_grad:
            model._emb_dense_dirty = True
        ops.gemm(dlog, rows, V, d, n, model._proj_grad, a_mn=True, b_mn=True, out_f32=True, accumulate=True,
                 split_k=1, ldo=d)
        ops.colsum(dlog, model._g("pred_layer.proj.bias"), rows=n, n=V)  # ragged V: whole vectors inside the pitch
        dt = None
        if ctx.needs_input_grad[0]:
            # d rows = dlogits E: a [n x d] output (12 tiles of 256 x 256 for n = 1024) with K = V = 250 002 — without a
            # K split 12 clusters would each run ~3 900 k-blocks while 62 idle.  Every K split stores its fp32 partial
            # in its own slab and one kernel adds the slabs in index order and rounds once: no atomics, so the bf16
            # gradient entering the encoder backward cannot flip with the arrival order of the partial sums.
            tiles = ((n + 255) // 256) * ((d + 255) // 256)
            kblocks = (V + 63) // 64
            split = max(1, min(32, 148 // max(tiles, 1), kblocks))
            split = -(-kblocks // -(-kblocks // split))  # the library drops empty splits: ceil(kb / ceil(kb / split))
            slabs = torch.empty(split, n, d, dtype=_F32, device=dev)
            ops.gemm(dlog, model._emb16, n, d, V, slabs, b_mn=True, out_f32=True, split_k=split, ldo=d,
                     split_stride=n * d if split > 1 else 0)
            drows = torch.empty(n, d, dtype=_BF16, device=dev)
            ops.sum_slabs_bf16(slabs, split, n * d, drows, n * d)
            dt = torch.zeros(slen, bs, d, dtype=_BF16, device=dev)
            ops.scatter_rows(drows, idx, bs, bs * d, d, dt, n, d)
        return dt, None, None, None, None, None
