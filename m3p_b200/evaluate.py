"""Forward-only image-text retrieval evaluation on the B200 path (reference:
`XEvaluator.evaluate_image_retrieval`, M3P/src/evaluation/xevaluator.py:1528-1657).

The reference scores EVERY (image, caption) pair with the full joint encoder + ITM head — `jointfwd` on the image
repeated against a split of the captions, `predict(is_relation=True)` on position 0 — in `eval()` mode under
`torch.no_grad()`, then counts recall@1/5/10 in both directions on the CPU.  Same calls here (so the same kernels
as training, without stashes), with the score matrix and the recall counts kept on the device; the pair count is
quadratic, which makes this the longest phase of the fine-tune recipe.
"""
import torch


@torch.no_grad()
def matching_scores(model, x_img, image_loc, captions, cap_lengths, pairs_per_call=256):
    """scores[i, c] = ITM logit of image i with caption c.

    x_img (R, n_img, 2048) fp32, image_loc (R, n_img, 5), captions (T, n_cap) int64, cap_lengths (n_cap,).
    Pairs are scored `pairs_per_call` at a time: one image against a block of captions per call, like the
    reference's `img_input.repeat(1, split_len, 1, 1)` (:1564-1565)."""
    was_training = model.training
    model.eval()
    R, n_img = x_img.shape[0], x_img.shape[1]
    n_cap = captions.shape[1]
    dev = x_img.device
    scores = torch.empty(n_img, n_cap, dtype=torch.float32, device=dev)
    img_len_full = torch.full((pairs_per_call,), R, dtype=torch.long, device=dev)
    for i in range(n_img):
        xi, li = x_img[:, i:i + 1], image_loc[:, i:i + 1]
        for c0 in range(0, n_cap, pairs_per_call):
            c1 = min(c0 + pairs_per_call, n_cap)
            n = c1 - c0
            enc = model("jointfwd", x=captions[:, c0:c1], lengths=cap_lengths[c0:c1], x_img=xi.expand(R, n, -1).contiguous(),
                        lengths_img=img_len_full[:n], causal=False, langs=None, image_loc=li.expand(R, n, -1).contiguous(),
                        refine_image=False)                                               # :1590-1594
            s = model("predict", tensor=enc.transpose(0, 1), is_relation=True)               # :1598
            scores[i, c0:c1] = s.view(-1).float()
    if was_training:
        model.train()
    return scores


def recall_at_k(scores, labels, ks=(1, 5, 10)):
    """Recall@k as the reference counts it (:1622-1654): `labels[i, c] == 1` marks caption c as belonging to
    image i.  Returns ({k: i2t recall}, {k: t2i recall}): image -> sentence = the best-ranked caption of an image
    that is correct, per image; sentence -> image = per caption."""
    kmax = max(ks)

    def direction(sc, lab):
        n = sc.shape[0]
        _, pred = sc.topk(min(kmax, sc.shape[1]), dim=-1)
        hit = lab.gather(1, pred) == 1                                   # (n, kmax): is the j-th ranked item correct
        first = torch.where(hit.any(dim=1), hit.float().argmax(dim=1), torch.full((n,), 10 ** 6, device=sc.device))
        return {k: float((first < k).float().sum()) / n for k in ks}

    return direction(scores, labels), direction(scores.t(), labels.t())


def evaluate_image_retrieval(model, x_img, image_loc, captions, cap_lengths, seq_per_img=5, pairs_per_call=256):
    """xevaluator.py:1528-1657 on in-memory test data: captions [seq_per_img * i, seq_per_img * (i + 1)) belong to
    image i.  Returns (t2i_r1, t2i_r5, t2i_r10, i2t_r1, i2t_r5, i2t_r10) like the reference."""
    n_img, n_cap = x_img.shape[1], captions.shape[1]
    assert n_cap == n_img * seq_per_img
    scores = matching_scores(model, x_img, image_loc, captions, cap_lengths, pairs_per_call)
    labels = (torch.arange(n_cap, device=scores.device)[None, :] // seq_per_img ==
              torch.arange(n_img, device=scores.device)[:, None]).long()
    i2t, t2i = recall_at_k(scores, labels)
    return t2i[1], t2i[5], t2i[10], i2t[1], i2t[5], i2t[10]
