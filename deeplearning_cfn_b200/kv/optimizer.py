"""Optimizers and the per-key `Updater` used by KVStore servers and local updates.

Reference: `MX/python/mxnet/optimizer/optimizer.py` - the registry + 17 optimizers (:450-1502), `Updater`
(lazy per-key state, `update_multi_precision`, (de)serialisation :1504-1556), the `Test` optimizer
`w += rescale_grad * g` that the distributed tests rely on (:1486-1499), and `lr_scheduler.py`.
Math follows the reference kernels (optimizer_op-inl.h); tensors are torch tensors on any device.  The
flat-buffer fused CUDA versions of SGD / momentum / Adam live in csrc/cuda/elementwise.cu and are used
by the bucketed trainer; these per-key versions serve the KVStore API.
"""
from __future__ import annotations

import math
import pickle

import torch

_REGISTRY: dict = {}


def register(cls):
    _REGISTRY[cls.__name__.lower()] = cls
    return cls


def create(name, **kwargs):
    if isinstance(name, Optimizer):
        return name
    try:
        return _REGISTRY[name.lower()](**kwargs)
    except KeyError:
        raise ValueError(f"unknown optimizer {name!r}; known: {sorted(_REGISTRY)}")


# --------------------------------------------------------------------------- lr schedulers
class LRScheduler:
    def __init__(self, base_lr: float = 0.01):
        self.base_lr = base_lr

    def __call__(self, num_update: int) -> float:
        return self.base_lr


class FactorScheduler(LRScheduler):
    def __init__(self, step: int, factor: float = 1.0, stop_factor_lr: float = 1e-8, base_lr: float = 0.01):
        super().__init__(base_lr)
        if step < 1:
            raise ValueError("Schedule step must be greater or equal than 1 round")
        self.step, self.factor, self.stop = step, factor, stop_factor_lr

    def __call__(self, num_update):
        return max(self.base_lr * self.factor ** (num_update // self.step), self.stop)


class MultiFactorScheduler(LRScheduler):
    def __init__(self, step: list, factor: float = 1.0, base_lr: float = 0.01):
        super().__init__(base_lr)
        if any(b <= a for a, b in zip(step, step[1:])) or (step and step[0] < 1):
            raise ValueError("Schedule step must be an increasing integer list")
        self.step, self.factor = list(step), factor

    def __call__(self, num_update):
        return self.base_lr * self.factor ** sum(1 for s in self.step if num_update > s)


class PolyScheduler(LRScheduler):
    def __init__(self, max_update: int, base_lr: float = 0.01, pwr: float = 2.0):
        super().__init__(base_lr)
        self.max_update, self.pwr = max_update, pwr

    def __call__(self, num_update):
        if num_update >= self.max_update:
            return 0.0
        return self.base_lr * (1.0 - num_update / self.max_update) ** self.pwr


# --------------------------------------------------------------------------- base
class Optimizer:
    def __init__(self, rescale_grad: float = 1.0, learning_rate: float = 0.01, wd: float = 0.0,
                 clip_gradient: float | None = None, lr_scheduler: LRScheduler | None = None,
                 multi_precision: bool = False, begin_num_update: int = 0, param_idx2name: dict | None = None, **_):
        self.rescale_grad = rescale_grad
        self.lr = learning_rate
        self.wd = wd
        self.clip_gradient = clip_gradient
        self.lr_scheduler = lr_scheduler
        if lr_scheduler is not None:
            lr_scheduler.base_lr = learning_rate
        self.multi_precision = multi_precision
        self.begin_num_update = self.num_update = begin_num_update
        self._index_update_count: dict = {}
        self.lr_mult: dict = {}
        self.wd_mult: dict = {}
        self.idx2name = dict(param_idx2name or {})

    # -- bookkeeping identical in spirit to optimizer.py:_update_count/_get_lr/_get_wd
    def _update_count(self, index):
        c = self._index_update_count.get(index, self.begin_num_update) + 1
        self._index_update_count[index] = c
        self.num_update = max(c, self.num_update)
        return c

    def _get_lr(self, index):
        lr = self.lr_scheduler(self.num_update) if self.lr_scheduler is not None else self.lr
        name = self.idx2name.get(index, index)
        return lr * self.lr_mult.get(name, self.lr_mult.get(index, 1.0))

    def _get_wd(self, index):
        name = self.idx2name.get(index, index)
        return self.wd * self.wd_mult.get(name, self.wd_mult.get(index, 1.0))

    def set_learning_rate(self, lr):
        if self.lr_scheduler is not None:
            raise UserWarning("LRScheduler of the optimizer has already been defined")
        self.lr = lr

    @property
    def learning_rate(self):
        return self.lr_scheduler(self.num_update) if self.lr_scheduler is not None else self.lr

    def set_lr_mult(self, d):
        self.lr_mult = dict(d)

    def set_wd_mult(self, d):
        self.wd_mult = dict(d)

    def _grad(self, grad, weight=None, wd=0.0):
        g = grad.to(torch.float32) * self.rescale_grad
        if self.clip_gradient is not None:
            g = g.clamp(-self.clip_gradient, self.clip_gradient)
        if weight is not None and wd:
            g = g + wd * weight
        return g

    def create_state(self, index, weight):
        return None

    def create_state_multi_precision(self, index, weight):
        if self.multi_precision and weight.dtype in (torch.float16, torch.bfloat16):
            master = weight.detach().to(torch.float32).clone()
            return (master, self.create_state(index, master))
        return self.create_state(index, weight)

    def update(self, index, weight, grad, state):
        raise NotImplementedError

    def update_multi_precision(self, index, weight, grad, state):
        if self.multi_precision and weight.dtype in (torch.float16, torch.bfloat16):
            master, inner = state
            self.update(index, master, grad, inner)
            weight.copy_(master.to(weight.dtype))
        else:
            self.update(index, weight, grad, state)

    def update_row_sparse(self, index, weight, row_ids, grad_rows, state):
        """Lazy update with a row-sparse gradient: only the rows in `row_ids` (duplicates are summed) and their state
        rows change (reference SGDMomStdDnsRspDns / AdamStdDnsRspDns / _sparse_adagrad_update, optimizer_op.cu:33-262).
        CUDA fp32 tables with a natively known optimizer run on one dlcfn kernel; everything else gathers the rows,
        runs the dense rule on them and scatters them back."""
        from ..ops import sparse as sparse_ops
        rows, g = sparse_ops.coalesce(row_ids.to(weight.device), grad_rows.to(weight.device))
        cfg = self.native_config()
        if weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous() and cfg is not None:
            d = dict(kv.split("=") for kv in cfg.split(";"))
            kind = d.pop("kind")
            hp = {k: float(v) for k, v in d.items()}
            t = self._update_count(index)
            if kind == "adam":
                hp["lr"] *= (1.0 - hp["beta2"] ** t) ** 0.5 / (1.0 - hp["beta1"] ** t)
            elif kind == "adamax":
                hp["lr"] /= (1.0 - hp["beta1"] ** t)
            st = state if isinstance(state, (tuple, list)) else (state,)
            s1 = st[0] if len(st) > 0 else None
            s2 = st[1] if len(st) > 1 else None
            sparse_ops.sparse_row_update(kind, rows, g, weight, s1, s2, None, **hp)
            return
        gather = lambda s: None if s is None else s[rows].clone()        # noqa: E731
        w_sub = weight[rows].clone()
        if isinstance(state, (tuple, list)):
            st_sub = type(state)(gather(s) for s in state)
        else:
            st_sub = gather(state)
        self.update(index, w_sub, g.view(w_sub.shape), st_sub)
        weight[rows] = w_sub
        if isinstance(state, (tuple, list)):
            for s, ss in zip(state, st_sub):
                if s is not None:
                    s[rows] = ss
        elif state is not None:
            state[rows] = st_sub

    def native_config(self) -> str | None:
        """`kind=...;lr=...` string understood by the C++ server / CUDA kernels, or None if this optimizer
        has no native implementation (then it is pickled to the servers, kvstore.py:480-497)."""
        return None

    def _cfg(self, kind, **extra):
        if self.lr_scheduler is not None or self.lr_mult or self.wd_mult:
            return None
        clip = self.clip_gradient if self.clip_gradient is not None else -1.0
        items = dict(kind=kind, lr=self.lr, wd=self.wd, rescale=self.rescale_grad, clip=clip, **extra)
        return ";".join(f"{k}={v}" for k, v in items.items())


@register
class SGD(Optimizer):
    """w = w - lr*(g + wd*w) or with momentum m = mom*m - lr*(g + wd*w); w += m  (optimizer_op-inl.h:86-323)."""

    def __init__(self, momentum: float = 0.0, lazy_update: bool = True, **kw):
        super().__init__(**kw)
        self.momentum = momentum
        self.lazy_update = lazy_update

    def create_state(self, index, weight):
        return torch.zeros_like(weight, dtype=torch.float32) if self.momentum != 0.0 else None

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        g = self._grad(grad)
        w = weight.to(torch.float32) if weight.dtype != torch.float32 else weight
        if state is not None:
            state.mul_(self.momentum).add_(w, alpha=-lr * wd).add_(g, alpha=-lr)
            w = w + state
        else:
            w = (1.0 - lr * wd) * w - lr * g
        weight.copy_(w.to(weight.dtype))

    def native_config(self):
        kind = "sgd_mom" if self.momentum != 0.0 else "sgd"
        clip = self.clip_gradient if self.clip_gradient is not None else -1.0
        return (f"kind={kind};lr={self.lr};wd={self.wd};momentum={self.momentum};rescale={self.rescale_grad};"
                f"clip={clip}") if self.lr_scheduler is None and not self.lr_mult and not self.wd_mult else None


@register
class ccSGD(SGD):
    """Deprecated alias kept by the reference (optimizer.py: ccSGD)."""


@register
class NAG(SGD):
    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        g = self._grad(grad, weight.float(), wd)
        if state is not None:
            state.mul_(self.momentum).add_(g)
            weight.add_((-lr * (g + self.momentum * state)).to(weight.dtype))
        else:
            weight.add_((-lr * g).to(weight.dtype))

    def native_config(self):
        return self._cfg("nag", momentum=self.momentum) if self.momentum != 0.0 else SGD.native_config(self)


@register
class Signum(Optimizer):
    def __init__(self, learning_rate=0.01, momentum=0.9, wd_lh=0.0, **kw):
        super().__init__(learning_rate=learning_rate, **kw)
        self.momentum, self.wd_lh = momentum, wd_lh

    def create_state(self, index, weight):
        return torch.zeros_like(weight, dtype=torch.float32) if self.momentum != 0.0 else None

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        g = self._grad(grad, weight.float(), wd)
        if state is not None:
            state.mul_(self.momentum).add_(g, alpha=-(1 - self.momentum))
            step = torch.sign(state)
        else:
            step = -torch.sign(g)
        weight.copy_(((1 - lr * self.wd_lh) * weight.float() + lr * step).to(weight.dtype))

    def native_config(self):
        if self.momentum != 0.0:
            return self._cfg("signum", momentum=self.momentum, beta1=self.wd_lh)
        return self._cfg("signsgd", beta1=self.wd_lh)


@register
class SGLD(Optimizer):
    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        g = self._grad(grad, weight.float(), wd)
        noise = torch.randn_like(g) * math.sqrt(lr)
        weight.add_((-lr / 2 * g + noise).to(weight.dtype))


@register
class DCASGD(Optimizer):
    def __init__(self, momentum=0.0, lamda=0.04, **kw):
        super().__init__(**kw)
        self.momentum, self.lamda = momentum, lamda

    def create_state(self, index, weight):
        mom = torch.zeros_like(weight, dtype=torch.float32) if self.momentum != 0.0 else None
        return (mom, weight.detach().float().clone())

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        g = self._grad(grad)
        mom, prev = state
        w = weight.float()
        d = -lr * (g + wd * w + self.lamda * g * g * (w - prev))
        if mom is not None:
            mom.mul_(self.momentum).add_(d)
            d = mom
        prev.copy_(w)
        weight.copy_((w + d).to(weight.dtype))


@register
class Adam(Optimizer):
    """m,v moments; bias correction folded into lr (optimizer.py:1080-1082, optimizer_op-inl.h:840-875)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, lazy_update=True, **kw):
        super().__init__(learning_rate=learning_rate, **kw)
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon

    def create_state(self, index, weight):
        return (torch.zeros_like(weight, dtype=torch.float32), torch.zeros_like(weight, dtype=torch.float32))

    def update(self, index, weight, grad, state):
        t = self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        lr *= math.sqrt(1.0 - self.beta2 ** t) / (1.0 - self.beta1 ** t)
        g = self._grad(grad, weight.float(), wd)
        m, v = state
        m.mul_(self.beta1).add_(g, alpha=1 - self.beta1)
        v.mul_(self.beta2).addcmul_(g, g, value=1 - self.beta2)
        weight.copy_((weight.float() - lr * m / (v.sqrt() + self.epsilon)).to(weight.dtype))

    def native_config(self):
        clip = self.clip_gradient if self.clip_gradient is not None else -1.0
        return (f"kind=adam;lr={self.lr};wd={self.wd};beta1={self.beta1};beta2={self.beta2};eps={self.epsilon};"
                f"rescale={self.rescale_grad};clip={clip}") if self.lr_scheduler is None else None


@register
class AdaGrad(Optimizer):
    def __init__(self, eps=1e-7, **kw):
        super().__init__(**kw)
        self.float_stable_eps = eps

    def create_state(self, index, weight):
        return torch.zeros_like(weight, dtype=torch.float32)

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        g = self._grad(grad)
        state.addcmul_(g, g)
        w = weight.float()
        weight.copy_((w - lr * (g / (state + self.float_stable_eps).sqrt() + wd * w)).to(weight.dtype))

    def native_config(self):
        return self._cfg("adagrad", eps=self.float_stable_eps)


@register
class GroupAdaGrad(Optimizer):
    """contrib.optimizer.GroupAdaGrad: one accumulator per row (embedding-style tables),
    hist += mean(g^2, axis=1); w -= lr * g / sqrt(hist + eps)  (contrib/optimizer_op.cu group_adagrad_update)."""

    def __init__(self, eps=1e-5, **kw):
        super().__init__(**kw)
        self.float_stable_eps = eps

    def create_state(self, index, weight):
        assert weight.dim() == 2, "GroupAdaGrad needs 2-D weights"
        return torch.zeros((weight.shape[0], 1), dtype=torch.float32, device=weight.device)

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr = self._get_lr(index)
        assert self._get_wd(index) == 0, "weight decay is not supported by GroupAdaGrad"
        if weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous():
            from ..ops import sparse as sparse_ops
            clip = self.clip_gradient if self.clip_gradient is not None else -1.0
            sparse_ops.group_adagrad_update(None, grad, weight, state.view(-1), lr=lr, eps=self.float_stable_eps,
                                            rescale=self.rescale_grad, clip=clip)
            return
        g = self._grad(grad)
        state.add_((g * g).mean(dim=1, keepdim=True))
        weight.copy_((weight.float() - lr * g / (state + self.float_stable_eps).sqrt()).to(weight.dtype))

    def update_row_sparse(self, index, weight, row_ids, grad_rows, state):
        from ..ops import sparse as sparse_ops
        rows, g = sparse_ops.coalesce(row_ids.to(weight.device), grad_rows.to(weight.device))
        self._update_count(index)
        lr = self._get_lr(index)
        if weight.is_cuda and weight.dtype == torch.float32 and weight.is_contiguous():
            clip = self.clip_gradient if self.clip_gradient is not None else -1.0
            sparse_ops.group_adagrad_update(rows, g, weight, state.view(-1), lr=lr, eps=self.float_stable_eps,
                                            rescale=self.rescale_grad, clip=clip)
            return
        g = self._grad(g)
        h = state[rows] + (g * g).mean(dim=1, keepdim=True)
        state[rows] = h
        weight[rows] = (weight[rows].float() - lr * g / (h + self.float_stable_eps).sqrt()).to(weight.dtype)


@register
class RMSProp(Optimizer):
    def __init__(self, learning_rate=0.001, gamma1=0.9, gamma2=0.9, epsilon=1e-8, centered=False,
                 clip_weights=None, **kw):
        super().__init__(learning_rate=learning_rate, **kw)
        self.gamma1, self.gamma2, self.epsilon = gamma1, gamma2, epsilon
        self.centered, self.clip_weights = centered, clip_weights

    def create_state(self, index, weight):
        z = lambda: torch.zeros_like(weight, dtype=torch.float32)  # noqa: E731
        return (z(), z(), z()) if self.centered else (z(),)

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        g = self._grad(grad, weight.float(), wd)
        w = weight.float()
        if not self.centered:
            (n,) = state
            n.mul_(self.gamma1).addcmul_(g, g, value=1 - self.gamma1)
            w = w - lr * g / (n + self.epsilon).sqrt()
        else:
            n, gavg, delta = state
            n.mul_(self.gamma1).addcmul_(g, g, value=1 - self.gamma1)
            gavg.mul_(self.gamma1).add_(g, alpha=1 - self.gamma1)
            delta.mul_(self.gamma2).add_(-lr * g / (n - gavg * gavg + self.epsilon).sqrt())
            w = w + delta
        if self.clip_weights:
            w = w.clamp(-self.clip_weights, self.clip_weights)
        weight.copy_(w.to(weight.dtype))

    def native_config(self):
        if self.centered:
            return None          # three states: python updater
        return self._cfg("rmsprop", beta1=self.gamma1, beta2=self.clip_weights or -1.0, eps=self.epsilon)


@register
class AdaDelta(Optimizer):
    def __init__(self, rho=0.90, epsilon=1e-5, **kw):
        super().__init__(**kw)
        self.rho, self.epsilon = rho, epsilon

    def create_state(self, index, weight):
        return (torch.zeros_like(weight, dtype=torch.float32), torch.zeros_like(weight, dtype=torch.float32))

    def update(self, index, weight, grad, state):
        self._update_count(index)
        wd = self._get_wd(index)
        g = self._grad(grad)
        acc_g, acc_d = state
        acc_g.mul_(self.rho).addcmul_(g, g, value=1 - self.rho)
        d = (acc_d + self.epsilon).sqrt() / (acc_g + self.epsilon).sqrt() * g
        acc_d.mul_(self.rho).addcmul_(d, d, value=1 - self.rho)
        w = weight.float()
        weight.copy_((w - d - wd * w).to(weight.dtype))

    def native_config(self):
        return self._cfg("adadelta", beta1=self.rho, eps=self.epsilon)


@register
class Ftrl(Optimizer):
    def __init__(self, lamda1=0.01, learning_rate=0.1, beta=1, **kw):
        super().__init__(learning_rate=learning_rate, **kw)
        self.lamda1, self.beta = lamda1, beta

    def create_state(self, index, weight):
        return (torch.zeros_like(weight, dtype=torch.float32), torch.zeros_like(weight, dtype=torch.float32))

    def update(self, index, weight, grad, state):
        self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        g = self._grad(grad)
        z, n = state
        w = weight.float()
        z.add_(g - ((n + g * g).sqrt() - n.sqrt()) * w / lr)
        n.addcmul_(g, g)
        neww = (torch.sign(z) * self.lamda1 - z) / ((self.beta + n.sqrt()) / lr + wd) * (z.abs() > self.lamda1)
        weight.copy_(neww.to(weight.dtype))

    def native_config(self):
        return self._cfg("ftrl", beta1=self.lamda1, beta2=self.beta)


@register
class Adamax(Optimizer):
    def __init__(self, learning_rate=0.002, beta1=0.9, beta2=0.999, **kw):
        super().__init__(learning_rate=learning_rate, **kw)
        self.beta1, self.beta2 = beta1, beta2

    def create_state(self, index, weight):
        return (torch.zeros_like(weight, dtype=torch.float32), torch.zeros_like(weight, dtype=torch.float32))

    def update(self, index, weight, grad, state):
        t = self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        lr /= (1.0 - self.beta1 ** t)
        g = self._grad(grad, weight.float(), wd)
        m, u = state
        m.mul_(self.beta1).add_(g, alpha=1 - self.beta1)
        torch.maximum(u * self.beta2, g.abs(), out=u)
        weight.copy_((weight.float() - lr * m / (u + 1e-38)).to(weight.dtype))

    def native_config(self):
        return self._cfg("adamax", beta1=self.beta1, beta2=self.beta2)


@register
class Nadam(Optimizer):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, schedule_decay=0.004, **kw):
        super().__init__(learning_rate=learning_rate, **kw)
        self.beta1, self.beta2, self.epsilon, self.schedule_decay = beta1, beta2, epsilon, schedule_decay
        self.m_schedule = 1.0

    def create_state(self, index, weight):
        return (torch.zeros_like(weight, dtype=torch.float32), torch.zeros_like(weight, dtype=torch.float32))

    def update(self, index, weight, grad, state):
        t = self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        g = self._grad(grad, weight.float(), wd)
        mom_t = self.beta1 * (1.0 - 0.5 * 0.96 ** (t * self.schedule_decay))
        mom_t1 = self.beta1 * (1.0 - 0.5 * 0.96 ** ((t + 1) * self.schedule_decay))
        self.m_schedule *= mom_t
        m_schedule_next = self.m_schedule * mom_t1
        m, v = state
        m.mul_(self.beta1).add_(g, alpha=1 - self.beta1)
        v.mul_(self.beta2).addcmul_(g, g, value=1 - self.beta2)
        g_prime = g / (1.0 - self.m_schedule)
        m_prime = m / (1.0 - m_schedule_next)
        v_prime = v / (1.0 - self.beta2 ** t)
        m_bar = (1.0 - mom_t) * g_prime + mom_t1 * m_prime
        weight.copy_((weight.float() - lr * m_bar / (v_prime.sqrt() + self.epsilon)).to(weight.dtype))


@register
class FTML(Optimizer):
    def __init__(self, beta1=0.6, beta2=0.999, epsilon=1e-8, **kw):
        super().__init__(**kw)
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon

    def create_state(self, index, weight):
        z = lambda: torch.zeros_like(weight, dtype=torch.float32)  # noqa: E731
        return (z(), z(), z())

    def update(self, index, weight, grad, state):
        t = self._update_count(index)
        lr, wd = self._get_lr(index), self._get_wd(index)
        g = self._grad(grad, weight.float(), wd)
        d_prev, v, z = state
        v.mul_(self.beta2).addcmul_(g, g, value=1 - self.beta2)
        d_t = (1 - self.beta1 ** t) / lr * ((v / (1 - self.beta2 ** t)).sqrt() + self.epsilon)
        sigma = d_t - self.beta1 * d_prev
        z.mul_(self.beta1).add_((1 - self.beta1) * g - sigma * weight.float())
        d_prev.copy_(d_t)
        weight.copy_((-z / d_t).to(weight.dtype))


@register
class LBSGD(SGD):
    """Large-batch SGD with LARS-style layer-wise adaptive rate (reference optimizer.py: LBSGD, simplified to
    the `lars` strategy; warm-up handled by the lr scheduler)."""

    def __init__(self, eta: float = 0.001, **kw):
        kw.pop("warmup_strategy", None)
        kw.pop("warmup_epochs", None)
        kw.pop("batch_scale", None)
        kw.pop("updates_per_epoch", None)
        kw.pop("begin_epoch", None)
        kw.pop("num_epochs", None)
        super().__init__(**kw)
        self.eta = eta

    def update(self, index, weight, grad, state):
        w2 = float(weight.float().norm())
        g2 = float(grad.float().norm()) * self.rescale_grad
        ratio = 1.0
        if w2 > 0 and g2 > 0:
            ratio = self.eta * w2 / (g2 + self._get_wd(index) * w2 + 1e-9)
        old = self.lr
        self.lr = old * min(ratio, 1.0) if ratio < 1.0 else old
        try:
            super().update(index, weight, grad, state)
        finally:
            self.lr = old


@register
class Test(Optimizer):
    """w += rescale_grad * g - the closed-form optimizer of the distributed tests (optimizer.py:1486-1499)."""

    def create_state(self, index, weight):
        return torch.zeros_like(weight)

    def update(self, index, weight, grad, state):
        weight.add_((grad.to(torch.float32) * self.rescale_grad).to(weight.dtype))

    def native_config(self):
        return f"kind=test;rescale={self.rescale_grad}"


# --------------------------------------------------------------------------- updater
class Updater:
    """Callable (key, grad, weight) that owns the per-key optimizer state (optimizer.py:1504-1556)."""

    def __init__(self, optimizer: Optimizer):
        self.optimizer = optimizer
        self.states: dict = {}
        self.states_synced: dict = {}

    def __call__(self, index, grad, weight):
        if index not in self.states:
            self.states[index] = self.optimizer.create_state_multi_precision(index, weight)
        self.optimizer.update_multi_precision(index, weight, grad, self.states[index])

    def update_row_sparse(self, index, row_ids, grad_rows, weight):
        """Row-sparse flavour of __call__ (the reference passes a RowSparseNDArray gradient to the same updater)."""
        if index not in self.states:
            self.states[index] = self.optimizer.create_state(index, weight)
        self.optimizer.update_row_sparse(index, weight, row_ids, grad_rows, self.states[index])

    def get_states(self, dump_optimizer: bool = False) -> bytes:
        cpu = {k: _to_cpu(v) for k, v in self.states.items()}
        return pickle.dumps((cpu, self.optimizer) if dump_optimizer else cpu)

    def set_states(self, blob: bytes) -> None:
        obj = pickle.loads(blob)
        if isinstance(obj, tuple) and len(obj) == 2 and isinstance(obj[1], Optimizer):
            self.states, self.optimizer = obj
        else:
            self.states = obj


def _to_cpu(s):
    if isinstance(s, torch.Tensor):
        return s.detach().cpu()
    if isinstance(s, (tuple, list)):
        return type(s)(_to_cpu(x) for x in s)
    return s


def get_updater(optimizer: Optimizer) -> Updater:
    return Updater(optimizer)
