"""CPU oracle for the M3P encoder training path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain fp32 PyTorch restatement (functional, over a state_dict) of the reference algorithm in
/root/reference/M3P/src/model/transformer.py and of the loss assembly in
/root/reference/M3P/src/xtrainer.py.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; the product path (m3p_b200/) never
does, and fails loudly when its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors, so this restatement is
pinned against outputs of the reference itself, generated in the build container by
oracle/make_golden.py (which imports the reference from /root/reference) and committed under
tests/golden/.  tests/test_oracle_golden.py checks every fixture.

Every function cites the reference lines it follows.  Dropout is identity here (the parity bar is
stated at dropout = attention_dropout = 0; dropout itself is validated separately).
"""
import math

import torch
import torch.nn.functional as F

LN_EPS = 1e-12  # transformer.py:244,659,694,709


def gelu(x):
    """erf-form GELU — transformer.py:48-56."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def get_masks(slen, lengths, causal=False):
    """transformer.py:59-78 (non-causal branch: attn_mask is mask)."""
    assert not causal, "causal masks are outside the encoder training path"
    alen = torch.arange(slen, dtype=torch.long, device=lengths.device)
    mask = alen < lengths[:, None]
    return mask, mask


# ---- rounding-matched mode ---------------------------------------------------------------------------
# `with rounding_matched():` makes the SAME restatement round to bf16 at exactly the points where the B200 kernels
# do (tensor-core operands = weights and the bf16 operand copies of activations; fp32 accumulation, softmax,
# LayerNorm statistics, biases, the whole residual stream and losses stay fp32), with a straight-through gradient.  What remains between this
# mode and the kernels is accumulation order (and a rare 1-ulp flip of a bf16 rounding), which is what the
# north_star's 1e-3 bound can meaningfully be stated against; the default mode is the reference's fp32 math.
_ROUND = False


class _RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def _r(x):
    return _RoundSTE.apply(x) if _ROUND else x


class rounding_matched:
    def __enter__(self):
        global _ROUND
        self.prev, _ROUND = _ROUND, True

    def __exit__(self, *exc):
        global _ROUND
        _ROUND = self.prev


def _linear(sd, prefix, x):
    """nn.Linear; in rounding-matched mode the two GEMM operands are bf16, the bias and the result fp32."""
    return F.linear(_r(x), _r(sd[prefix + ".weight"]), sd[prefix + ".bias"])


def _layer_norm(sd, prefix, x):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], LN_EPS)


def image_embeddings(sd, x_img, image_loc):
    """BertImageEmbeddings.forward — transformer.py:247-269 (input_dist is None on this path)."""
    e = _linear(sd, "image_embeddings.image_embeddings", x_img) + \
        F.linear(image_loc, sd["image_embeddings.image_location_embeddings.weight"],
                 sd["image_embeddings.image_location_embeddings.bias"])  # K = 5: plain fp32 FMAs in the kernel
    return _layer_norm(sd, "image_embeddings.LayerNorm", e)


def attention(sd, i, n_heads, h, mask):
    """MultiHeadAttention.forward, self-attention branch — transformer.py:149-210."""
    bs, qlen, dim = h.shape
    dh = dim // n_heads
    p = "attentions.%d." % i

    def shape(x):
        return x.view(bs, -1, n_heads, dh).transpose(1, 2)

    q = shape(_r(_linear(sd, p + "q_lin", h)))              # the packed QKV projection is stored in bf16
    k = shape(_r(_linear(sd, p + "k_lin", h)))
    v = shape(_r(_linear(sd, p + "v_lin", h)))
    q = q / math.sqrt(dh)                                   # :197 (after the bias; 1/8 is exact in bf16)
    scores = torch.matmul(q, k.transpose(2, 3))             # :198
    scores = scores.masked_fill((mask == 0).view(bs, 1, 1, qlen), -float("inf"))  # :199-200
    if _ROUND:
        # kernel: p = exp(s - max) enters the P V product as bf16; the row sum is taken over the fp32 values
        e = torch.exp(scores - scores.max(dim=-1, keepdim=True).values)
        ctx = torch.matmul(_r(e), v) / e.sum(dim=-1, keepdim=True)
        ctx = _r(ctx).transpose(1, 2).contiguous().view(bs, qlen, dim)
    else:
        weights = F.softmax(scores.float(), dim=-1).type_as(scores)                   # :202
        ctx = torch.matmul(weights, v).transpose(1, 2).contiguous().view(bs, qlen, dim)  # :204-205
    return _linear(sd, p + "out_lin", ctx)                  # :208


def ffn(sd, i, h):
    """TransformerFFN.forward — transformer.py:222-227."""
    p = "ffns.%d." % i
    return _linear(sd, p + "lin2", _r(gelu(_linear(sd, p + "lin1", h))))


def layers(sd, n_layers, n_heads, h, mask):
    """The layer loop shared by fwd / jointfwd / crossfwd — transformer.py:947-958, 842-864."""
    m = mask.unsqueeze(-1).to(h.dtype)
    for i in range(n_layers):
        # kernels: the residual stream (pre-LayerNorm sums, and the copy of every LayerNorm output that the next
        # residual add reads) is fp32; only the tensor-core operand copy of a LayerNorm output is bf16, and that
        # rounding happens inside _linear
        h = _layer_norm(sd, "layer_norm1.%d" % i, h + attention(sd, i, n_heads, h, mask))
        h = _layer_norm(sd, "layer_norm2.%d" % i, h + ffn(sd, i, h)) * m
    return _r(h)  # the encoder output handed to the heads is the bf16 operand copy


def embed_joint(sd, x, lengths, x_img, lengths_img, image_loc, text_embed=None):
    """Embedding stage of jointfwd — transformer.py:897-943.  Returns (h (bs, R+T, d) fp32, mask)."""
    slen, bs = x.shape
    img = image_embeddings(sd, x_img.transpose(0, 1), image_loc.transpose(0, 1))      # :897-901
    txt = text_embed if text_embed is not None else F.embedding(x.transpose(0, 1), sd["embeddings.weight"])
    c_slen = img.shape[1] + slen
    mask, _ = get_masks(c_slen, lengths_img + lengths)                                  # :916-919
    h = torch.cat([img, txt], dim=1)                                                    # :929
    h = h + sd["position_embeddings.weight"][:c_slen].unsqueeze(0)                      # :932-936
    h = h * mask.unsqueeze(-1).to(h.dtype)                                              # :940
    return _layer_norm(sd, "layer_norm_emb", h), mask                                   # :942


def jointfwd(sd, n_layers, n_heads, x, lengths, x_img, lengths_img, image_loc, text_embed=None):
    """TransformerModel.jointfwd — transformer.py:878-968.

    x (T,B) int64; x_img (R,B,2048); image_loc (R,B,5); returns (R+T, B, d).  `langs` is accepted and
    ignored by the reference (:937-938).  Order: mask -> LN_emb (:940-942).
    """
    h, mask = embed_joint(sd, x, lengths, x_img, lengths_img, image_loc, text_embed)
    h = layers(sd, n_layers, n_heads, h, mask)
    return h.transpose(0, 1)


def crossfwd_text(sd, n_layers, n_heads, x, lengths, positions=None, langs=None, text_embed=None):
    """TransformerModel.crossfwd(stream_='text') — transformer.py:970-1114; `fwd` text path
    (:753-876) is the same computation with langs ignored (:829)."""
    slen, bs = x.shape
    mask, _ = get_masks(slen, lengths)
    h = text_embed if text_embed is not None else F.embedding(x.transpose(0, 1), sd["embeddings.weight"])
    if positions is None:
        h = h + sd["position_embeddings.weight"][:slen].unsqueeze(0)
    else:
        h = h + F.embedding(positions.transpose(0, 1), sd["position_embeddings.weight"])
    if langs is not None:
        h = h + F.embedding(langs.transpose(0, 1), sd["cross_lang_embeddings.weight"])   # :1056-1057
    h = _layer_norm(sd, "layer_norm_emb", h)                                            # :1058
    h = h * mask.unsqueeze(-1).to(h.dtype)                                              # :1062
    h = layers(sd, n_layers, n_heads, h, mask)
    return h.transpose(0, 1)


def fwd_text(sd, n_layers, n_heads, x, lengths, positions=None):
    """TransformerModel.fwd text path — transformer.py:824-831 (no language embedding)."""
    return crossfwd_text(sd, n_layers, n_heads, x, lengths, positions=positions, langs=None)


def fwd_image(sd, n_layers, n_heads, x_img, lengths, image_loc):
    """TransformerModel.fwd(cross_modal=True) — transformer.py:822,831: image embeddings * mask,
    no layer_norm_emb, no position embedding."""
    R, bs = x_img.shape[0], x_img.shape[1]
    mask, _ = get_masks(R, lengths)
    h = image_embeddings(sd, x_img.transpose(0, 1), image_loc.transpose(0, 1))
    h = h * mask.unsqueeze(-1).to(h.dtype)
    h = layers(sd, n_layers, n_heads, h, mask)
    return h.transpose(0, 1)


# ---- heads: TransformerModel.predict — transformer.py:1183-1214 ------------------------------------

def predict_mlm(sd, tensor, pred_mask, y):
    """default branch -> PredLayer.forward (:104-117); tensor (S',B,d), pred_mask (S',B) bool."""
    dim = tensor.shape[-1]
    rows = tensor[pred_mask.unsqueeze(-1).expand_as(tensor)].view(-1, dim)
    scores = _r(F.linear(_r(rows), _r(sd["pred_layer.proj.weight"]), sd["pred_layer.proj.bias"]))  # bf16 logits
    return scores, F.cross_entropy(scores, y, reduction="mean")


def predict_obj(sd, tensor, y):
    """is_obj: BertPredictionHeadTransform (:602-606) -> ObjPredLayer (:576-584); tensor (B,R,d)."""
    t = _r(_layer_norm(sd, "transformer_obj.LayerNorm", _r(gelu(_linear(sd, "transformer_obj.dense", tensor)))))
    scores = _r(_linear(sd, "pred_obj_layer.proj", t)).view(-1, sd["pred_obj_layer.proj.weight"].shape[0])
    return scores, F.cross_entropy(scores, y, reduction="mean", ignore_index=-1)


def predict_mrfr(sd, tensor):
    """is_mrfr (:1202-1204)."""
    return _r(_linear(sd, "mrfr_dense", tensor))


def predict_relation(sd, tensor, clcm=False):
    """is_relation / is_clcm: BertPooler (:552-558) on position 0 of the batch-first tensor, then
    seq_relationship (:713,1195-1201)."""
    pl, sr = ("pooled_layer2", "seq_relationship2") if clcm else ("pooled_layer", "seq_relationship")
    pooled = _r(torch.tanh(_linear(sd, pl + ".dense", tensor[:, 0])))
    return F.linear(pooled, sd[sr + ".weight"], sd[sr + ".bias"])  # 1-row projection: fp32 weights in the kernel


# ---- loss assembly: XTrainer.pretrain_under_step — xtrainer.py:2285-2375 ---------------------------

def get_mask_(labels):
    """xtrainer.py:2226-2232: pred_mask = labels != -1 ; y = labels[labels > 0]."""
    return labels[labels > 0], labels != -1


def relation_loss(scores, pos_labels, sample_n, w_multi=1.0, w_bin=1.0):
    """xtrainer.py:2359-2372: CE over groups of sample_n + BCE against the one-hot positive."""
    ce = F.cross_entropy(scores.view(-1, sample_n), pos_labels)
    onehot = F.one_hot(pos_labels, sample_n).to(scores.dtype)
    bce = F.binary_cross_entropy_with_logits(scores.view(-1), onehot.view(-1))
    return w_multi * ce + w_bin * bce


def clcm_loss(scores, clcm_labels):
    """xtrainer.py:2389-2391: BCE of the second pass's is_clcm scores against clcm_labels."""
    return F.binary_cross_entropy_with_logits(scores.view(-1), clcm_labels.view(-1).float())


def pretrain_step_losses(sd, n_layers, n_heads, batch, sample_n, heads=("mlm", "mrm", "mrfr", "rel")):
    """One multitask step (xtrainer.py:2281-2375) on a batch dict with keys
    x, lengths, x_img, lengths_img, image_loc, x_labels (T,B), obj_labels (B,R), ori_feats (B,R,2048),
    pos_labels (B/sample_n,).  Returns (encoder_out, dict of losses, total)."""
    R = batch["x_img"].shape[0]
    out = jointfwd(sd, n_layers, n_heads, batch["x"], batch["lengths"], batch["x_img"], batch["lengths_img"],
                   batch["image_loc"])
    text_out = out[R:]
    img_out = out[:R].transpose(0, 1)
    losses = {}
    total = 0.0
    if "mlm" in heads:
        y_text, pm_text = get_mask_(batch["x_labels"])
        if pm_text.sum() > 0:
            _, losses["mlm"] = predict_mlm(sd, text_out, pm_text, y_text)
            total = total + losses["mlm"]
    obj_labels = batch["obj_labels"]
    if "mrm" in heads and (obj_labels != -1).sum() > 0:
        _, losses["mrm"] = predict_obj(sd, img_out, obj_labels.reshape(-1))
        total = total + losses["mrm"]
    if "mrfr" in heads and (obj_labels != -1).sum() > 0:
        reg = predict_mrfr(sd, img_out).reshape(-1, 2048)
        sel = obj_labels.reshape(-1) != -1
        losses["mrfr"] = F.mse_loss(reg[sel], batch["ori_feats"].reshape(-1, 2048)[sel])
        total = total + losses["mrfr"]
    if "rel" in heads:
        scores = predict_relation(sd, out.transpose(0, 1))
        losses["rel"] = relation_loss(scores, batch["pos_labels"], sample_n)
        total = total + losses["rel"]
    return out, losses, total


# ---- synthetic batches (imitates retrieval_pretrain_collate xtrainer.py:960-1045) ----

def clip_coef(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ as the trainer calls it (xtrainer.py:222-225): one global L2 norm over all
    gradients, coefficient max_norm / (norm + 1e-6) clamped to 1.  Returns (coef, total_norm)."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return coef, total


def adam_step(p, g, exp_avg, exp_avg_sq, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
    """optim.py:45-86, one parameter tensor, in place; `step` is the 1-based update count (:68)."""
    b1, b2 = betas
    exp_avg.mul_(b1).add_(g, alpha=1 - b1)                       # :72
    exp_avg_sq.mul_(b2).addcmul_(g, g, value=1 - b2)             # :73
    denom = exp_avg_sq.sqrt().add_(eps)                          # :74
    step_size = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)  # :76-78
    if weight_decay != 0:
        p.add_(p, alpha=-weight_decay * lr)                      # :80-81
    p.addcdiv_(exp_avg, denom, value=-step_size)                 # :83
    return p


def lr_inverse_sqrt(num_updates, lr, warmup_updates=4000, warmup_init_lr=1e-7, exp_factor=0.5):
    """AdamInverseSqrtWithWarmup.get_lr_for_step — optim.py:129-133."""
    if num_updates < warmup_updates:
        return warmup_init_lr + num_updates * (lr - warmup_init_lr) / warmup_updates
    return lr * warmup_updates ** exp_factor * (num_updates ** -exp_factor)


def synthetic_batch(B, T, R, n_words, sample_n=4, seed=1234, ragged=False, n_mask_text=None, n_mask_img=None,
                    feat_dim=2048):
    g = torch.Generator().manual_seed(seed)
    lengths = torch.full((B,), T, dtype=torch.long)
    if ragged:
        lengths = torch.randint(max(4, T // 4), T + 1, (B,), generator=g)
        lengths[0] = T
    x = torch.randint(4, n_words - 2, (T, B), generator=g)
    x[0] = 0                                                  # <s>   (xtrainer.py:829-880)
    for b in range(B):
        x[lengths[b] - 1, b] = 2                              # </s>
        x[lengths[b]:, b] = 1                                 # <pad>
    x_img = F.normalize(torch.randn(R, B, feat_dim, generator=g), dim=-1)     # dataset_pretrain.py:287,379
    loc = torch.rand(R, B, 5, generator=g)
    image_loc = loc / loc.norm(dim=-1, keepdim=True)                          # :298-300
    lengths_img = torch.full((B,), R, dtype=torch.long)
    n_mt = n_mask_text if n_mask_text is not None else max(1, (T * 15 // 100) // 8 * 8 or 2)
    n_mi = n_mask_img if n_mask_img is not None else max(1, R * 16 // 100)
    x_labels = torch.full((T, B), -1, dtype=torch.long)
    obj_labels = torch.full((B, R), -1, dtype=torch.long)
    ori_feats = x_img.transpose(0, 1).clone()
    for b in range(B):
        n_valid = int(lengths[b]) - 2
        k = min(n_mt, max(n_valid, 0))
        if k > 0:
            pos = torch.randperm(n_valid, generator=g)[:k] + 1
            x_labels[pos, b] = torch.randint(4, n_words - 2, (k,), generator=g)
        posi = torch.randperm(R, generator=g)[:n_mi]
        obj_labels[b, posi] = torch.randint(1, 1600, (len(posi),), generator=g)
        x_img[posi, b] = 0.0                                  # masked regions are zeroed (:258-292)
    assert B % sample_n == 0
    pos_labels = torch.randint(0, sample_n, (B // sample_n,), generator=g)
    return dict(x=x, lengths=lengths, x_img=x_img, lengths_img=lengths_img, image_loc=image_loc,
                x_labels=x_labels, obj_labels=obj_labels, ori_feats=ori_feats, pos_labels=pos_labels)
