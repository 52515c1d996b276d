"""`train_x.py`-compatible entry for the B200 encoder training path (reference: M3P/train_x.py — parser :29-391,
`main` :394-508; trainer loop `Trainer.optimize` / `iter` / `print_stats`, M3P/src/xtrainer.py:205-289).

    python -m m3p_b200.train_x --emb_dim 768 --n_layers 12 --n_heads 12 --dropout 0.1 --attention_dropout 0.1 \
        --batch_size 16 --sample_n 4 --bptt 128 --max_region_num 100 \
        --optimizer adam_inverse_sqrt,beta1=0.9,beta2=0.98,lr=0.0001 --clip_grad_norm 5 \
        --cross_rel_steps coco-img --cross_mlm_steps coco-img --cross_mrm_steps coco-img --cross_mrfr_steps coco-img \
        --epoch_size 2000 --max_epoch 1
    torchrun --nproc-per-node 8 -m m3p_b200.train_x ...          # data parallel, NCCL

Same flag names and meanings as the reference for the subset this path serves (model shape, dropout, batch
geometry, optimizer DSL, gradient clipping / accumulation, loss weights, which pre-training heads are active,
checkpoint save / reload); flags of the reference's other subsystems (datasets, tokenizer, evaluation, decoding,
SLURM) are accepted and ignored only when they cannot change the computation, otherwise rejected.  The
reference's data pipeline needs h5py / lmdb and private datasets, so batches are the
seeded synthetic ones of `train_step.synthetic_batch` (layout of `retrieval_pretrain_collate`,
xtrainer.py:960-1045); `effective batch = batch_size * sample_n` pairs per GPU as in the reference.
Everything on the device runs through libm3p_sm100.so; there is no CPU fallback.
"""
import argparse
import json
import os
import time

import torch


def bool_flag(s):
    """utils.py:39-48."""
    if s.lower() in ("off", "false", "0"):
        return False
    if s.lower() in ("on", "true", "1"):
        return True
    raise argparse.ArgumentTypeError("Invalid value for a boolean flag!")


def get_parser():
    p = argparse.ArgumentParser(description="M3P encoder pre-training / ITM fine-tuning on B200 (synthetic data)")
    # experiment (train_x.py:36-46)
    p.add_argument("--dump_path", type=str, default="./dumped/")
    p.add_argument("--exp_name", type=str, default="")
    p.add_argument("--exp_id", type=str, default="")
    p.add_argument("--save_periodic", type=int, default=0)
    # model (train_x.py:57-86)
    p.add_argument("--emb_dim", type=int, default=768)
    p.add_argument("--n_layers", type=int, default=12)
    p.add_argument("--n_heads", type=int, default=12)
    p.add_argument("--dropout", type=float, default=0.1)
    p.add_argument("--attention_dropout", type=float, default=0.1)
    p.add_argument("--gelu_activation", type=bool_flag, default=True)
    p.add_argument("--share_inout_emb", type=bool_flag, default=True)
    p.add_argument("--sinusoidal_embeddings", type=bool_flag, default=False)
    p.add_argument("--refine_layers", type=int, default=6)
    p.add_argument("--refine_image", type=bool_flag, default=False)
    p.add_argument("--is_cross_modal", type=bool_flag, default=True)
    p.add_argument("--max_vocab", type=int, default=-1)
    p.add_argument("--n_words", type=int, default=250002, help="XLM-R vocabulary + <mask> (tokenization.py:79-81)")
    p.add_argument("--n_langs", type=int, default=1)
    # batch geometry (train_x.py:131-160)
    p.add_argument("--batch_size", type=int, default=16)
    p.add_argument("--sample_n", type=int, default=4)
    p.add_argument("--bptt", type=int, default=128)
    p.add_argument("--max_region_num", type=int, default=100)
    # optimisation (train_x.py:164-176, 371-377)
    p.add_argument("--optimizer", type=str, default="adam_inverse_sqrt,beta1=0.9,beta2=0.98,lr=0.0001")
    p.add_argument("--clip_grad_norm", type=float, default=5)
    p.add_argument("--accumulate_gradients", type=int, default=1)
    p.add_argument("--amp", type=int, default=-1, help="ignored: the path is bf16 tensor cores + fp32 masters (Apex removed)")
    p.add_argument("--fp16", type=bool_flag, default=False, help="ignored, see --amp")
    p.add_argument("--epoch_size", type=int, default=1000, help="pairs per epoch (per process)")
    p.add_argument("--max_epoch", type=int, default=1)
    # loss weights (train_x.py:186-212)
    for k in ("mlm", "mrm", "mrfr", "rel"):
        p.add_argument("--lambda_" + k, type=str, default="1")
    p.add_argument("--bin_cls_loss_weight", type=float, default=1)
    p.add_argument("--multi_cls_loss_weight", type=float, default=1)
    # which heads are active: the reference's step lists (non-empty = on), train_x.py:232-246
    p.add_argument("--cross_rel_steps", type=str, default="coco-img")
    p.add_argument("--cross_mlm_steps", type=str, default="")
    p.add_argument("--cross_mrm_steps", type=str, default="")
    p.add_argument("--cross_mrfr_steps", type=str, default="")
    # reload / checkpoints (train_x.py:250-254)
    p.add_argument("--reload_model", type=str, default="")
    p.add_argument("--save_every_epoch", type=bool_flag, default=False)
    # distributed (train_x.py:277-280): torchrun sets the env
    p.add_argument("--local_rank", type=int, default=-1)
    p.add_argument("--master_port", type=int, default=-1)
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--cuda_graph", type=bool_flag, default=True, help="capture forward+backward(+all-reduce) once and replay")
    return p


def model_namespace(params):
    """The attributes TransformerModel reads (transformer.py:627-645), as check_data_params / check_model_params
    would have set them (loader.py:147-153, model/__init__.py:19-65)."""
    langs = ["l%d" % i for i in range(params.n_langs)]
    return argparse.Namespace(
        n_langs=params.n_langs, n_words=params.n_words, eos_index=2, pad_index=1, id2lang=dict(enumerate(langs)),
        lang2id={l: i for i, l in enumerate(langs)}, emb_dim=params.emb_dim, n_heads=params.n_heads, n_layers=params.n_layers,
        n_dec_layers=params.n_layers, dropout=params.dropout, attention_dropout=params.attention_dropout,
        sinusoidal_embeddings=params.sinusoidal_embeddings, refine_layers=params.refine_layers, attention_setting="v1",
        use_externel_att=False, gelu_activation=params.gelu_activation, share_inout_emb=params.share_inout_emb, asm=False)


def main(params):
    from . import ops
    from .ddp import GradReducer, init_distributed
    from .optim import get_optimizer
    from .train_step import GraphedStep, pretrain_step, synthetic_batch
    from .transformer import TransformerModel

    if not torch.cuda.is_available():
        raise SystemExit("m3p_b200.train_x needs a B200 (no CPU fallback)")
    if params.refine_image or not params.is_cross_modal:
        raise NotImplementedError("refine_image / is_cross_modal=False are outside the B200 path")
    rank, local, world = init_distributed()
    dev = torch.device("cuda", local)
    ops.device_check()
    torch.manual_seed(params.seed)
    model = TransformerModel(model_namespace(params), is_encoder=True, with_output=True, is_crossModal=True)
    if params.reload_model:  # model/__init__.py:96-105: strip DDP's "module." prefix, tolerate missing keys
        sd = torch.load(params.reload_model, map_location="cpu", weights_only=False)
        sd = sd.get("model", sd)
        sd = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in sd.items()}
        missing = model.load_state_dict(sd, strict=False)
        if rank == 0:
            print("reloaded %s (%d missing keys)" % (params.reload_model, len(missing.missing_keys)), flush=True)
    model = model.to(dev).train()
    reducer = GradReducer(model)
    optimizer = get_optimizer([p for p in model.parameters() if p.requires_grad], params.optimizer,
                              clip_grad_norm=params.clip_grad_norm)
    heads = tuple(h for h, on in (("mlm", params.cross_mlm_steps), ("mrm", params.cross_mrm_steps),
                                  ("mrfr", params.cross_mrfr_steps), ("rel", params.cross_rel_steps)) if on)
    assert heads, "no head is active: set at least one of --cross_rel_steps / --cross_mlm_steps / ..."
    lambdas = {k: float(getattr(params, "lambda_" + k)) for k in ("mlm", "mrm", "mrfr", "rel")}
    B = params.batch_size * params.sample_n  # retrieval_pretrain_collate flattens (bs, sample_n) (xtrainer.py:1031-1045)
    n_data = 8  # distinct synthetic batches cycled through pinned host memory
    host = [synthetic_batch(B, params.bptt, params.max_region_num, params.n_words, sample_n=params.sample_n,
                            seed=1234 + 97 * rank + i) for i in range(n_data)]
    host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    first = {k: v.to(dev) for k, v in host[0].items()}
    acc = max(1, params.accumulate_gradients)

    graphed = None
    if params.cuda_graph and acc == 1:
        graphed = GraphedStep(model, first, params.sample_n, heads, lambdas, warmup=2,
                              after_backward=reducer.finish if world > 1 else None,
                              capture_error_mode="thread_local" if world > 1 else "global")

    os.makedirs(params.dump_path, exist_ok=True)
    n_iter, n_pairs, t_last, pairs_last = 0, 0, time.time(), 0
    for epoch in range(params.max_epoch):
        seen = 0
        while seen < params.epoch_size:
            batch = host[n_iter % n_data]
            if graphed is not None:
                # pipelined feed: this batch was staged during the previous step; stage the next one now
                if n_iter == 0:
                    graphed.prefetch(batch)
                loss = graphed.step_prefetched(host[(n_iter + 1) % n_data])
            else:
                if n_iter % acc == 0:
                    model.zero_grad()
                dbatch = {k: v.to(dev, non_blocking=True) for k, v in batch.items()}
                loss, _ = pretrain_step(model, dbatch, params.sample_n, heads, lambdas)
                if (n_iter + 1) % acc == 0:
                    (loss / acc).backward()     # the last micro-step announces the (accumulated) slices to NCCL
                    reducer.finish()
                else:
                    with reducer.accumulate():  # earlier micro-steps must not send anything (ddp.GradReducer.accumulate)
                        (loss / acc).backward()
                loss = loss.detach()
            if (n_iter + 1) % acc == 0:
                optimizer.step()  # clip + Adam + bf16 operand refresh + gradient clear (xtrainer.py:222-228)
            n_iter += 1
            seen += B
            n_pairs += B
            if n_iter % 20 == 0 and rank == 0:  # xtrainer.py:254-289 (every 5 iterations there)
                torch.cuda.synchronize()
                now = time.time()
                print("%7i - %8.2f pairs/s/GPU - loss %.4f - lr %.4e - grad norm %.3f" % (
                    n_iter, (n_pairs - pairs_last) / (now - t_last), float(loss), optimizer.param_groups[0]["lr"],
                    float(optimizer.last_grad_norm.sqrt()) if optimizer.last_grad_norm is not None else float("nan")),
                    flush=True)
                t_last, pairs_last = now, n_pairs
        if rank == 0:
            print("__log__:%s" % json.dumps({"epoch": epoch, "n_iter": n_iter, "loss": float(loss)}), flush=True)
            if params.save_every_epoch or epoch == params.max_epoch - 1:
                # xtrainer.py:511-529: weights + params only
                torch.save({"model": model.state_dict(), "params": vars(params)},
                           os.path.join(params.dump_path, "checkpoint-%d.pth" % epoch))
    if world > 1:
        torch.cuda.synchronize()
        if graphed is not None:
            graphed.release()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main(get_parser().parse_args())
