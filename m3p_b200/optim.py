"""Optimizers of the M3P trainer on the B200 path (reference: M3P/src/optim.py — `Adam` :16-86,
`AdamInverseSqrtWithWarmup` :89-139, `AdamCosineWithWarmup` :142-208, `get_optimizer` :211-270 — and the
gradient clipping of `Trainer.optimize`, M3P/src/xtrainer.py:222-225,234-237).

Same class names, constructor arguments, `param_groups` keys (`lr`, `num_updates`) and update rule as the
reference, but the step is not a Python loop over ~400 tensors x ~10 elementwise kernels: the trained
parameters of `m3p_b200.transformer.TransformerModel` live in flat fp32 buffers, so one step is

    m3p_sumsq_f32  x (number of gradient buffers)      global gradient norm, stays on the device
    m3p_adam_step  x (number of gradient buffers)      clip + Adam + bf16 operand copy + gradient zeroing

(include/m3p_b200.h).  Parameters that do not belong to a flat buffer (the reference's never-trained modules)
are skipped exactly as the reference skips `p.grad is None`; if one of them does have a gradient it goes through
the same kernel on its own.  There is no PyTorch fallback.

Gradient clipping: the reference clips in the trainer, between backward and step.  Here the optimizer does it
(`opt.clip_grad_norm = 5.0`, or `get_optimizer(..., clip_grad_norm=5.0)`), because the norm never has to leave
the device that way; `torch.nn.utils.clip_grad_norm_` on the parameter list still works (the gradients are views
of the flat buffer) but costs ~400 small kernels.
"""
import math
import re
import weakref

import torch

from . import ops

_F32 = torch.float32


def _flat_buffers(params):
    """Partition `params` into (models, loose): the TransformerModels that own them (their flat buffers are
    stepped as a whole) and the parameters that belong to no flat buffer."""
    models, loose, seen = [], [], set()
    for p in params:
        owner = getattr(p, "_m3p_owner", None)
        m = owner() if owner is not None else None
        if m is None:
            loose.append(p)
        elif id(m) not in seen:
            seen.add(id(m))
            models.append(m)
    return models, loose


class Adam(torch.optim.Optimizer):
    """optim.py:16-86 — Adam without amsgrad; `weight_decay` decays the weights directly (:80-81)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, clip_grad_norm=0.0,
                 zero_grad_in_step=True, refresh_operands_in_step=True):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter at index 0: {}".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter at index 1: {}".format(betas[1]))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.clip_grad_norm = float(clip_grad_norm)
        self.zero_grad_in_step = zero_grad_in_step
        self.refresh_operands_in_step = refresh_operands_in_step
        self.n_steps = 0       # the reference keeps state['step'] per parameter; they all advance together
        self._moments = {}     # data_ptr of a parameter buffer -> (exp_avg, exp_avg_sq)
        self._sumsq = None
        self.last_grad_norm = None  # device scalar: sqrt of it is the pre-clip global gradient norm

    # -- buffers ------------------------------------------------------------------------------------
    def _state_for(self, buf):
        key = buf.data_ptr()
        if key not in self._moments:
            self._moments[key] = (torch.zeros_like(buf), torch.zeros_like(buf))
        return self._moments[key]

    def _work_list(self):
        """[(param buffer, grad buffer, bf16 copy or None, model or None)] for this step."""
        work = []
        for group in self.param_groups:
            models, loose = _flat_buffers(group["params"])
            for m in models:
                if m._flat_grad is None:
                    continue  # no backward yet
                if not m._flat.is_cuda:
                    raise RuntimeError("m3p_b200.optim runs on a B200 only (no CPU fallback): move the model to cuda")
                p16 = None
                if self.refresh_operands_in_step:
                    if m._flat16 is None:
                        m._flat16 = torch.empty(m._flat_numel, dtype=torch.bfloat16, device=m._flat.device)
                    p16 = m._flat16
                work.append((group, m._flat, m._flat_grad, p16, m))
                # the MLM head's bf16 copy of the (tied) projection matrix, once the head has been used
                e16 = m._emb16 if (self.refresh_operands_in_step and m._emb16 is not None) else None
                tied = m._proj is m._emb
                work.append((group, m._emb.data, m._emb_grad, e16 if tied else None, m))
                if m._proj is not None and not tied:
                    work.append((group, m._proj.data, m._proj_grad, e16, m))
            for p in loose:
                if p.grad is None:
                    continue  # optim.py:55-56
                if not (p.is_cuda and p.dtype == _F32 and p.is_contiguous() and p.grad.is_contiguous()):
                    raise RuntimeError("m3p_b200.optim: loose parameters must be contiguous fp32 CUDA tensors")
                work.append((group, p.data, p.grad.data, None, None))
        return work

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        work = self._work_list()
        if not work:
            return loss
        ops.use_current_stream()
        self.n_steps += 1
        sumsq = None
        if self.clip_grad_norm > 0:
            dev = work[0][1].device
            if self._sumsq is None or self._sumsq.device != dev:
                self._sumsq = torch.zeros(1, dtype=_F32, device=dev)
            self._sumsq.zero_()
            for _, _, g, _, _ in work:
                ops.sumsq(g, self._sumsq)
            sumsq = self.last_grad_norm = self._sumsq
        for group, p, g, p16, m in work:
            exp_avg, exp_avg_sq = self._state_for(p)
            b1, b2 = group["betas"]
            ops.adam_step(p, g, exp_avg, exp_avg_sq, self.n_steps, group["lr"], b1, b2, group["eps"],
                          weight_decay=group["weight_decay"], p16=p16, grad_sumsq=sumsq,
                          max_grad_norm=self.clip_grad_norm, zero_grad=self.zero_grad_in_step)
            if m is not None:
                if p is m._flat:
                    m._operands_valid = p16 is not None   # _flat16 mirrors the updated masters (or is stale)
                elif p.data_ptr() == m._proj.data_ptr() if m._proj is not None else False:
                    m._emb16_valid = p16 is not None
                if self.zero_grad_in_step:
                    m._grads_clean = True          # zero_grad() has nothing left to do until the next backward
                    m._emb_touched, m._emb_dense_dirty = [], False
        return loss

    # -- checkpoint plumbing (xtrainer.py:531-560 stores optimizer.state_dict(); reload keeps only
    #    num_updates / lr, :583-599) ------------------------------------------------------------------
    def state_dict(self):
        groups = [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]
        return {"param_groups": groups, "n_steps": self.n_steps,
                "moments": [(m.clone(), v.clone()) for m, v in self._moments.values()]}

    def load_state_dict(self, sd, restore_moments=False):
        """The reference's checkpoint reload keeps only `num_updates` and `lr` of each param group and starts Adam's
        moments AND its bias-correction step from zero (xtrainer.py:579-590, "Not reloading checkpoint optimizer"):
        that is the default here — moments dropped, n_steps = 0, so the bias correction matches the zeroed moments.
        restore_moments=True resumes exactly instead (moments and step count of this optimizer's own state_dict)."""
        for g, saved in zip(self.param_groups, sd["param_groups"]):
            g.update({k: v for k, v in saved.items() if k != "params"})
        if restore_moments and sd.get("moments") is not None:
            work = self._work_list()
            if len(work) != len(sd["moments"]):
                raise ValueError("optimizer state has %d moment buffers, this optimizer %d (run one backward first so "
                                 "the flat gradient buffers exist)" % (len(sd["moments"]), len(work)))
            for (_, p, _, _, _), (m, v) in zip(work, sd["moments"]):
                self._moments[p.data_ptr()] = (m.to(p.device).clone(), v.to(p.device).clone())
            self.n_steps = int(sd.get("n_steps", 0))
        else:
            self._moments = {}
            self.n_steps = 0


# ---- learning-rate schedules: pure functions of the update count (the classes below only hold their settings) ----

def warmup_lr(n, init_lr, peak_lr, warmup):
    """Linear ramp init_lr -> peak_lr over `warmup` updates (both reference schedules start with it)."""
    return init_lr + n * (peak_lr - init_lr) / warmup


def inverse_sqrt_lr(n, peak_lr, warmup, init_lr=1e-7, exponent=0.5):
    """optim.py:129-133: after the ramp, peak_lr * (warmup / n) ** exponent."""
    if n < warmup:
        return warmup_lr(n, init_lr, peak_lr, warmup)
    return peak_lr * warmup ** exponent * n ** (-exponent)


def cosine_lr(n, peak_lr, warmup, init_lr=1e-7, floor_lr=1e-9, period=1000000, growth=1, shrink=0.75):
    """optim.py:184-201: after the ramp, cosine cycles; cycle c lasts period * growth**c updates and runs between
    floor_lr * shrink**c and peak_lr * shrink**c."""
    if n < warmup:
        return warmup_lr(n, init_lr, peak_lr, warmup)
    t = n - warmup
    if growth == 1:
        cycle = math.floor(t / period)
        length, into = period, t - period * cycle
    else:
        cycle = math.floor(math.log(1 - t / period * (1 - growth), growth))
        length = period * growth ** cycle
        into = t - (1 - growth ** cycle) / (1 - growth) * period
    lo, hi = floor_lr * shrink ** cycle, peak_lr * shrink ** cycle
    return lo + 0.5 * (hi - lo) * (1 + math.cos(math.pi * into / length))


class _ScheduledAdam(Adam):
    """Adam whose param_groups carry `num_updates` and get a new `lr` after every step (optim.py:135-139, 203-208)."""

    def __init__(self, params, start_lr, **kw):
        super().__init__(params, lr=start_lr, **kw)
        for group in self.param_groups:
            group["num_updates"] = 0

    def get_lr_for_step(self, num_updates):
        raise NotImplementedError

    def step(self, closure=None):
        loss = super().step(closure)
        for group in self.param_groups:
            group["num_updates"] += 1
            group["lr"] = self.get_lr_for_step(group["num_updates"])
        return loss


class AdamInverseSqrtWithWarmup(_ScheduledAdam):
    """optim.py:89-139 (same constructor arguments)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, warmup_updates=4000,
                 warmup_init_lr=1e-7, exp_factor=0.5, **kw):
        super().__init__(params, warmup_init_lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)
        self.peak_lr, self.warmup_updates, self.warmup_init_lr, self.exp_factor = lr, warmup_updates, warmup_init_lr, exp_factor

    def get_lr_for_step(self, num_updates):
        return inverse_sqrt_lr(num_updates, self.peak_lr, self.warmup_updates, self.warmup_init_lr, self.exp_factor)


class AdamCosineWithWarmup(_ScheduledAdam):
    """optim.py:142-208 (same constructor arguments)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, warmup_updates=4000,
                 warmup_init_lr=1e-7, min_lr=1e-9, init_period=1000000, period_mult=1, lr_shrink=0.75, **kw):
        super().__init__(params, warmup_init_lr, betas=betas, eps=eps, weight_decay=weight_decay, **kw)
        self.cfg = dict(peak_lr=lr, warmup=warmup_updates, init_lr=warmup_init_lr, floor_lr=min_lr, period=init_period,
                        growth=period_mult, shrink=lr_shrink)

    def get_lr_for_step(self, num_updates):
        return cosine_lr(num_updates, **self.cfg)


# ---- the optimizer string DSL of train_x.py's --optimizer flag (optim.py:211-270) -----------------------------
_NUMBER = re.compile(r"^[+-]?(\d+(\.\d*)?|\.\d+)$")
_FAMILY = {
    "adam": (Adam, {"lr", "eps", "weight_decay"}),
    "adam_inverse_sqrt": (AdamInverseSqrtWithWarmup, {"lr", "eps", "weight_decay", "warmup_updates", "warmup_init_lr",
                                                      "exp_factor"}),
    "adam_cosine": (AdamCosineWithWarmup, {"lr", "eps", "weight_decay", "warmup_updates", "warmup_init_lr", "min_lr",
                                           "init_period", "period_mult", "lr_shrink"}),
}
_NOT_BUILT = ("adadelta", "adagrad", "adamax", "asgd", "rmsprop", "rprop", "sgd")  # torch.optim pass-throughs there
_INTEGER_ARGS = ("warmup_updates", "init_period")


def get_optimizer(parameters, s, **extra):
    """"adam_inverse_sqrt,beta1=0.9,beta2=0.98,lr=0.0001" -> optimizer, like the reference's parser: a method name
    followed by comma-separated name=number pairs; beta1 / beta2 fold into `betas`; an argument the class does not
    take is an error.  Only the Adam family of the published recipes is built for the B200 path — the reference's
    torch.optim pass-throughs raise NotImplementedError instead of silently running a different code path.
    `extra` forwards keyword arguments such as clip_grad_norm=5.0."""
    method, _, tail = s.partition(",")
    given = {}
    for item in filter(None, tail.split(",")):
        name, eq, value = item.partition("=")
        if not eq or "=" in value or _NUMBER.match(value) is None:
            raise AssertionError("optimizer argument %r is not of the form name=number" % item)
        given[name] = float(value)
    if method in _NOT_BUILT:
        raise NotImplementedError("optimizer %r is outside the B200 path (only the Adam family is fused)" % method)
    if method not in _FAMILY:
        raise Exception('Unknown optimization method: "%s"' % method)
    cls, accepted = _FAMILY[method]
    kwargs = {"betas": (given.pop("beta1", 0.9), given.pop("beta2", 0.999))}
    unknown = sorted(set(given) - accepted)
    if unknown:
        raise Exception('Unexpected parameters: expected "%s", got "%s"' % (sorted(accepted), unknown))
    kwargs.update({k: (int(v) if k in _INTEGER_ARGS else v) for k, v in given.items()})
    kwargs.update(extra)
    return cls(parameters, **kwargs)


def tag_parameters(model):
    """Called by TransformerModel: lets an optimizer built from a bare parameter list find the flat buffers."""
    ref = weakref.ref(model)
    for _, p in model.hot_parameters():
        p._m3p_owner = ref
