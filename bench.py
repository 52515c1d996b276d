"""Headline benchmark: image-text pairs/s, forward + backward, M3P-base (12L/768H/12 heads, 100 regions +
128 tokens = 228-token joint sequence, bf16 tensor-core operands), 64 pairs per GPU (BASELINE.json
configs[1]; data-parallel weak scaling for --gpus > 1, configs[2]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--heads itm|multitask] [--impl reference]

One step = jointfwd + prediction head(s) + loss + backward through the reference-facing API
(`model('jointfwd', ...)`, `model('predict', ...)`, `loss.backward()`), gradients zeroed and the bf16
operand copies refreshed inside the timed region, plus the NCCL gradient all-reduce when N > 1.  The
optimizer update is not part of the metric.  Prints ONE JSON line on rank 0.

`--impl reference` times the reference algorithm's CPU implementation (the oracle port of
transformer.py / xtrainer.py under oracle/, fp32 PyTorch on all host cores) on a bounded sample of
the same workload; it is the one place besides tests/ and smoke() that executes oracle/.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GF_PER_PAIR = {"itm": 122.886, "multitask": 143.356}  # BASELINE.md §3 (matmul FLOPs, fwd+bwd = 3x fwd)
GF_PER_PAIR_LARGE = {"itm": 429.714, "multitask": 457.167}  # SURVEY.md §8d, M3P-large 24L/1024H (BASELINE configs[4])
HEADS = {"itm": ("rel",), "multitask": ("mlm", "mrm", "mrfr", "rel")}
CFG = dict(emb_dim=768, n_layers=12, n_heads=12, n_words=250002, T=128, R=100, sample_n=4, dropout=0.1)
CPU_SAMPLE_PAIRS = 4


def namespace(cfg, dropout=None):
    return argparse.Namespace(
        n_langs=1, n_words=cfg["n_words"], eos_index=2, pad_index=1, id2lang={0: "en"}, lang2id={"en": 0},
        emb_dim=cfg["emb_dim"], n_heads=cfg["n_heads"], n_layers=cfg["n_layers"], n_dec_layers=cfg["n_layers"],
        dropout=cfg["dropout"] if dropout is None else dropout,
        attention_dropout=cfg["dropout"] if dropout is None else dropout, sinusoidal_embeddings=False,
        refine_layers=1, attention_setting="v1", use_externel_att=False, gelu_activation=True, share_inout_emb=True,
        asm=False)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(sustained=d.get("bf16_tflops_sustained", 1400.0), burst=d.get("bf16_tflops", 1590.0),
                    hbm=d.get("hbm_gbs", 6650.0), source="measured (MEASURED_PEAKS.json)")
    return dict(sustained=1400.0, burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ts, r in self.rows:
            if self.t0 is not None and not (self.t0 <= ts <= (self.t1 or ts)):
                continue  # only samples taken DURING the timed region
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


REF_STAGE = os.path.join(ROOT, "baseline", "_ref", "M3P")  # staged by __graft_entry__.build() (git-ignored)


def reference_available():
    return os.path.exists(os.path.join(REF_STAGE, "src", "model", "transformer.py"))


def _reference_model(device):
    """The UNMODIFIED reference class (microsoft/M3P src/model/transformer.py, staged under baseline/_ref/ by
    build()), constructed as model/__init__.py:93 does, M3P-base shape, fp32."""
    import torch
    if REF_STAGE not in sys.path:
        sys.path.insert(0, REF_STAGE)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src.model.transformer import TransformerModel as RefModel
    torch.manual_seed(0)
    return RefModel(namespace(CFG, 0.0), is_encoder=True, with_output=True, is_crossModal=True).to(device).train()


def _reference_step(model, batch, heads):
    """pretrain_under_step's forward + loss assembly (xtrainer.py:2281-2375) on the reference class + backward."""
    import torch.nn.functional as F
    R = batch["x_img"].shape[0]
    enc = model("jointfwd", x=batch["x"], lengths=batch["lengths"], x_img=batch["x_img"], lengths_img=batch["lengths_img"],
                causal=False, langs=None, image_loc=batch["image_loc"], refine_image=False)
    total = 0.0
    if "mlm" in heads:
        total = total + model("predict", tensor=enc[R:], pred_mask=batch["pred_mask_text"], y=batch["y_text"], get_scores=False)[1]
    if "mrm" in heads:
        total = total + model("predict", tensor=enc[:R].transpose(0, 1), pred_mask=None, y=batch["obj_labels"].view(-1),
                              get_scores=False, is_obj=True)[1]
    if "mrfr" in heads:
        reg = model("predict", tensor=enc[:R].transpose(0, 1), is_mrfr=True).reshape(-1, 2048)
        sel = batch["obj_labels"].reshape(-1) != -1
        total = total + F.mse_loss(reg[sel], batch["ori_feats"].reshape(-1, 2048)[sel])
    if "rel" in heads:
        sc = model("predict", tensor=enc.transpose(0, 1), is_relation=True)
        n = CFG["sample_n"]
        total = total + F.cross_entropy(sc.view(-1, n), batch["pos_labels"]) + F.binary_cross_entropy_with_logits(
            sc.view(-1), F.one_hot(batch["pos_labels"], n).float().view(-1))
    model.zero_grad(set_to_none=True)
    total.backward()
    return total


def cpu_reference_pairs_per_s(heads, steps, warmup, threads=None):
    """The reference's own CPU path on the box's host cores: the unmodified reference class when it has been staged
    (kind "reference"), else the oracle port of the same algorithm (kind "port").  Full M3P-base weights, fp32, a
    B = CPU_SAMPLE_PAIRS sample of the batch.  Returns (pairs/s, threads, s/step, kind)."""
    import torch
    from m3p_b200.train_step import synthetic_batch
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    B = CPU_SAMPLE_PAIRS
    if reference_available():
        model = _reference_model("cpu")
        batch = synthetic_batch(B, CFG["T"], CFG["R"], CFG["n_words"], sample_n=CFG["sample_n"], seed=1234)
        times = []
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            _reference_step(model, batch, heads)
            if it >= warmup:
                times.append(time.perf_counter() - t0)
        return B * len(times) / sum(times), threads, sum(times) / len(times), "reference"
    return _cpu_port_pairs_per_s(heads, steps, warmup, threads) + ("port",)


def reference_on_b200(heads, B=16, steps=3):
    """Context line: the same unmodified reference class in eager PyTorch ON THE B200 (fp32, and under
    torch.autocast(bfloat16)) — what a user of the reference gets on this GPU without this library."""
    import torch
    from m3p_b200.train_step import synthetic_batch
    if not reference_available():
        return None
    out = {"pairs_per_step": B, "kind": "unmodified reference class, eager PyTorch, dropout 0"}
    model = _reference_model("cuda")
    batch = synthetic_batch(B, CFG["T"], CFG["R"], CFG["n_words"], sample_n=CFG["sample_n"], seed=1234, device="cuda")
    for name, ac in (("fp32", False), ("autocast_bf16", True)):
        ts = []
        for it in range(steps + 2):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=ac):
                _reference_step(model, batch, heads)
            t1.record()
            torch.cuda.synchronize()
            if it >= 2:
                ts.append(t0.elapsed_time(t1))
        out[name + "_pairs_per_s"] = B * len(ts) / (sum(ts) * 1e-3)
    del model
    torch.cuda.empty_cache()
    return out


def _cpu_port_pairs_per_s(heads, steps, warmup, threads):
    """Fallback when the reference has not been staged: oracle/m3p_oracle.py (fp32 PyTorch restatement of
    transformer.py + xtrainer.py loss assembly)."""
    import torch
    from m3p_b200.transformer import TransformerModel
    from m3p_b200.train_step import synthetic_batch
    from oracle import m3p_oracle as O
    torch.manual_seed(0)
    model = TransformerModel(namespace(CFG, 0.0), is_encoder=True, with_output=True, is_crossModal=True)
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()
            if k.split(".")[0] in ("embeddings", "position_embeddings", "layer_norm_emb", "image_embeddings", "attentions",
                                   "layer_norm1", "ffns", "layer_norm2", "pooled_layer", "seq_relationship", "mrfr_dense",
                                   "transformer_obj", "pred_obj_layer") or k == "pred_layer.proj.bias"}
    leaf["pred_layer.proj.weight"] = leaf["embeddings.weight"]
    del model, sd
    B = CPU_SAMPLE_PAIRS
    batch = synthetic_batch(B, CFG["T"], CFG["R"], CFG["n_words"], sample_n=CFG["sample_n"], seed=1234)
    times = []
    for it in range(warmup + steps):
        for v in leaf.values():
            v.grad = None
        t0 = time.perf_counter()
        _, _, total = O.pretrain_step_losses(leaf, CFG["n_layers"], CFG["n_heads"], batch, CFG["sample_n"], heads=heads)
        total.backward()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return B * len(times) / sum(times), threads, sum(times) / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    heads = HEADS[args.heads]
    v, threads, spt, kind = cpu_reference_pairs_per_s(heads, args.steps, args.warmup)
    sample = "%d pairs/step x %d steps, M3P-base fp32, %s on %d host threads" % (
        CPU_SAMPLE_PAIRS, args.steps, "unmodified reference class (baseline/_ref)" if kind == "reference" else "oracle port", threads)
    print(json.dumps({
        "impl": "reference", "metric": "image-text pairs/sec fwd+bwd, M3P-base 228-tok seq", "value": v, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": spt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.heads), "sample": sample},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def workload_name(heads):
    h = "ITM head (t2i/i2t fine-tune step)" if heads == "itm" else "xMLM-style MLM + MRM + MRFR + ITM heads (multitask step)"
    return "M3P-%s %dL/%dH/%dh jointfwd fwd+bwd, 100 regions + 128 tokens, 64 pairs/GPU, " % (
        "base" if CFG["emb_dim"] == 768 else "large", CFG["n_layers"], CFG["emb_dim"], CFG["n_heads"]) + h


def time_dominant_kernel(torch, ops, L):
    """The FFN GEMM (M=14592, N=3072, K=768; 2/3 of the linear FLOPs are this shape or its transposes)
    timed alone with CUDA events: the `roofline.dominant_kernel` entry."""
    ops.use_current_stream()  # the step ran on the graph's capture stream: latch torch's current stream again
    m, n, k = 14592, 3072, 768
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    w = torch.randn(n, k, device="cuda").to(torch.bfloat16)
    bias = torch.randn(n, device="cuda")
    out, out2 = torch.empty(m, n, device="cuda", dtype=torch.bfloat16), torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    times = []
    for it in range(13):
        flush.zero_()  # 256 MB > the 126 MB L2
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        ops.linear(a, w, bias, out, epi=L.M3P_EPI_GELU, out2=out2)
        t1.record()
        torch.cuda.synchronize()
        if it >= 3:
            times.append(t0.elapsed_time(t1))
    ms = statistics.median(times)
    return 2.0 * m * n * k / (ms * 1e-3) / 1e12, ms


def dominant_kernel_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/r02_ncu_dominant.json, written by tools/ncu_dominant.py from the .ncu-rep);
    None when no capture has been committed."""
    p = os.path.join(ROOT, "profiles", "r02_ncu_dominant.json")
    if not os.path.exists(p):
        return None, None
    d = json.load(open(p))
    return d.get("dram_bytes_read", 0) + d.get("dram_bytes_write", 0), d


def time_optimizer(torch, model):
    """the fused clip + Adam step over the model's flat buffers, timed alone (CUDA events) against the
    HBM roofline: sumsq reads 4 B/param; the Adam pass reads p, g, m, v and writes p, m, v, g (zeroed) in fp32 plus
    the bf16 operand copy = 34 B/param."""
    from m3p_b200 import optim
    opt = optim.get_optimizer([p for p in model.parameters() if p.requires_grad], "adam,lr=0.00001", clip_grad_norm=5.0)
    n_params = model._flat_numel + model._emb.numel()
    ts = []
    for it in range(7):
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        opt.step()
        t1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(t0.elapsed_time(t1))
    ms = statistics.median(ts)
    nbytes = n_params * 38.0
    pk = peaks()
    return {"ms": ms, "params": n_params, "bytes_per_param": 38, "achieved_gbs": nbytes / (ms * 1e-3) / 1e9,
            "peak_gbs": pk["hbm"], "frac": nbytes / (ms * 1e-3) / 1e9 / pk["hbm"],
            "note": "clip_grad_norm (m3p_sumsq_f32, 4 B/param) + m3p_adam_step (34 B/param) on the flat fp32 buffers, "
                    "timed alone; not part of the fwd+bwd metric"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--heads", default="itm", choices=["itm", "multitask"])
    ap.add_argument("--batch", type=int, default=64, help="pairs per GPU")
    ap.add_argument("--model", default="base", choices=["base", "large"],
                    help="base = the headline M3P-base workload; large = BASELINE configs[4] (24L/1024H/16 heads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.model == "large":
        CFG.update(emb_dim=1024, n_layers=24, n_heads=16)
        GF_PER_PAIR.update(GF_PER_PAIR_LARGE)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its banner there)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from m3p_b200 import lib as L, ops
    from m3p_b200.ddp import GradReducer, init_distributed
    from m3p_b200.train_step import GraphedStep, pretrain_step, synthetic_batch
    from m3p_b200.transformer import TransformerModel

    rank, local, world = init_distributed()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the M3P hot path has no CPU fallback (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    ops.device_check()
    heads = HEADS[args.heads]
    B = args.batch
    torch.manual_seed(0)
    model = TransformerModel(namespace(CFG), is_encoder=True, with_output=True, is_crossModal=True).cuda().train()
    reduce_dtype = None if os.environ.get("M3P_DDP_FP32", "0") == "1" else torch.bfloat16
    reducer = GradReducer(model, overlap=os.environ.get("M3P_DDP_OVERLAP", "1") != "0", reduce_dtype=reduce_dtype)
    host = synthetic_batch(B, CFG["T"], CFG["R"], CFG["n_words"], sample_n=CFG["sample_n"], seed=1234 + rank)
    host = {k: v.pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    step_keys = ["x", "lengths", "x_img", "lengths_img", "image_loc", "pos_labels"]
    if args.heads == "multitask":
        step_keys += ["pred_mask_text", "y_text", "obj_labels", "ori_feats", "mrfr_weight"]
    h2d_bytes = sum(host[k].numel() * host[k].element_size() for k in step_keys)

    use_graph = not args.no_graph
    graphed = None
    if use_graph:  # for N > 1 the NCCL gradient exchange is captured into the same graph
        graphed = GraphedStep(model, resident, CFG["sample_n"], heads, warmup=3,
                              after_backward=reducer.finish if world > 1 else None,
                              capture_error_mode="thread_local" if world > 1 else "global")

    def step(batch):
        if graphed is not None:  # `batch` is None (replay on the static inputs) or a dict of new inputs to copy in
            return graphed.step(None if batch is resident else batch)
        model.zero_grad()
        total, _ = pretrain_step(model, batch, CFG["sample_n"], heads)
        total.backward()
        reducer.finish()
        return total

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.LAUNCHES
        t0.record()
        for _ in range(n):
            fn()
        t1.record()
        barrier()
        ms = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, ops.LAUNCHES - l0

    # ---- kernel-resident arm: inputs already in HBM ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # nvidia-smi takes a moment to start: launch it before the warm-up
    for _ in range(args.warmup):
        step(resident)
    sampler.mark_begin()
    ms, launches = timed(lambda: step(resident), args.steps)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end arm: pinned host inputs -> H2D -> step -> loss D2H, every step ----
    def e2e_step():
        if graphed is not None:
            b = {k: host[k] for k in step_keys}  # pinned host tensors, copied into the graph's static inputs
        else:
            b = dict(resident)
            for k in step_keys:
                b[k] = host[k].to(dev, non_blocking=True)
        return float(step(b).detach())  # device -> host read of the loss

    if graphed is not None:
        # pipelined input feed: the H2D copy of batch i+1 (pinned host -> staging, copy stream) runs underneath step i;
        # every step still copies its own inputs from the host and reads its loss back inside the timed region
        hb = {k: host[k] for k in step_keys}
        graphed.prefetch(hb)

        # The loss of every step is copied to pinned host memory inside the timed region (4 B D2H per step, enqueued
        # right behind the replay); the CPU reads the value one step late, so it never sits in a synchronisation while
        # the GPU waits for the next launch.
        loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        loss_done = [torch.cuda.Event(), torch.cuda.Event()]
        e2e_state = {"i": 0, "last": float("nan")}

        def e2e_step():  # noqa: F811
            i = e2e_state["i"]
            loss = graphed.step_prefetched(hb)
            loss_host[i & 1:(i & 1) + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_done[i & 1].record()
            if i > 0:
                loss_done[(i - 1) & 1].synchronize()
                e2e_state["last"] = float(loss_host[(i - 1) & 1])
            e2e_state["i"] = i + 1
            return e2e_state["last"]

    for _ in range(3):
        e2e_step()
    ms_e2e, _ = timed(e2e_step, args.steps)

    pairs = B * world * args.steps
    value = pairs / (ms * 1e-3)
    e2e = pairs / (ms_e2e * 1e-3)
    pk = peaks()
    gf = GF_PER_PAIR[args.heads]
    achieved = value * gf / 1e3 / world  # TFLOP/s per GPU
    out = None
    if rank == 0:
        dom_tf, dom_ms = time_dominant_kernel(torch, ops, L)
        dom_traffic, dom_src = dominant_kernel_traffic()
        out = {
            "metric": "image-text pairs/sec fwd+bwd, M3P-%s 228-tok seq" % args.model, "value": value, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args.heads), "global_batch": B * world, "seq_len": CFG["T"] + CFG["R"],
                       "parallelism": "dp%d" % world, "dropout": CFG["dropout"], "vocab": CFG["n_words"],
                       "grad_allreduce": None if world == 1 else ("bf16 slices on a comm stream inside backward + bf16 embedding rows"
                                                                    if reduce_dtype is not None else "fp32"),
                       "launch": "CUDA graph replay (one capture of zero_grad+fwd+loss+bwd%s)" % (" + NCCL gradient exchange" if world > 1 else "") if graphed is not None
                       else "per-kernel launches from Python",
                       "l2": "per-step working set (0.18 GB bf16 weights + >4 GB activations) >> 126 MB L2; no flush needed",
                       "gflop_per_pair": gf},
            "e2e": {"value": e2e, "unit": "pairs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "feed": "pinned host -> staging on a copy stream (overlaps the previous step) -> D2D into the graph's "
                            "inputs -> replay -> loss D2H into pinned memory (4 B, read by the CPU one step late)" if graphed is not None
                            else "H2D on the compute stream -> step -> loss.item()"},
            "gpu_launches": (graphed.launches_per_step * args.steps) if graphed is not None else launches,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["sustained"], "traffic": dom_traffic,
                         "scope": "whole fwd+bwd step per GPU: pairs/s x %.3f GFLOP/pair (BASELINE.md §3) over the "
                                  "sustained cuBLAS bf16 peak, %s" % (gf, pk["source"]),
                         "dominant_kernel": {"name": "gemm_kernel<256, GELU> FFN lin1 14592x3072x768", "ms": dom_ms,
                                             "achieved": dom_tf, "peak": pk["burst"], "frac": dom_tf / pk["burst"],
                                             "note": "timed alone, L2 flushed, vs burst peak",
                                             "traffic": dom_traffic,
                                             "algorithmic_bytes": 14592 * 768 * 2 + 3072 * 768 * 2 + 2 * 14592 * 3072 * 2,
                                             "traffic_note": "dram__bytes_read + dram__bytes_write of this kernel per "
                                             "launch from the committed ncu --set full capture (%s); `roofline.traffic` "
                                             "repeats it (the step itself is tensor-bound: bytes are not its roofline)"
                                             % (dom_src.get("source") if dom_src else "no capture committed")}},
        }
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            out["optimizer"] = time_optimizer(torch, model)
            out["reference_on_b200"] = reference_on_b200(heads)
            del model, graphed
            torch.cuda.empty_cache()
            v, threads, _, kind = cpu_reference_pairs_per_s(heads, 3, 1)
            out["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": threads, "kind": kind,
                                   "sample": "%d pairs/step x 3 steps after 1 warm-up, same M3P-base weights shape and batch "
                                             "generator, fp32, %s" % (CPU_SAMPLE_PAIRS, "unmodified reference class "
                                             "(baseline/_ref/M3P)" if kind == "reference" else "oracle port (oracle/m3p_oracle.py)")}
        else:
            out["cpu_baseline"] = None
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        if graphed is not None:  # NCCL: graphs holding captured collectives must go before the communicator does
            graphed.release()
        del graphed
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
