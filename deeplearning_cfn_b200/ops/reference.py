"""Plain-PyTorch fp32 reference implementations of every native op.

Two uses: (1) the numerics tests compare each CUDA kernel against these, (2) the CPU (gloo) control
path of the KVStore / trainer uses them so the whole stack is testable without a GPU.  They are never
selected on a CUDA tensor.
"""
from __future__ import annotations

import torch


def elementwise_sum(inputs, out, scale: float = 1.0):
    acc = torch.zeros_like(inputs[0], dtype=torch.float32)
    for t in inputs:
        acc += t.float()
    out.copy_((acc * scale).to(out.dtype))
    return out


def fused_optimizer(kind, grad, master, s1, s2, wout, *, lr, wd=0.0, momentum=0.0, beta1=0.9, beta2=0.999,
                    eps=1e-8, rescale=1.0, clip=-1.0):
    """Same math as csrc/cuda/elementwise.cu:opt_apply (reference optimizer_op-inl.h:86-427, 840-875)."""
    g = grad.float().reshape(-1) * rescale
    if clip > 0:
        g = g.clamp(-clip, clip)
    w = master.reshape(-1)
    if kind == "sgd":
        w.mul_(1.0 - lr * wd).add_(g, alpha=-lr)
    elif kind == "sgd_mom":
        s1.mul_(momentum).add_(w, alpha=-lr * wd).add_(g, alpha=-lr)
        w.add_(s1)
    elif kind == "adam":
        g = g + wd * w
        s1.mul_(beta1).add_(g, alpha=1.0 - beta1)
        s2.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
        w.addcdiv_(s1, s2.sqrt() + eps, value=-lr)
    elif kind == "test":
        w.add_(g)
    elif kind == "none":
        w.copy_(g)
    elif kind == "nag":
        g = g + wd * w
        s1.mul_(momentum).add_(g)
        w.add_(g + momentum * s1, alpha=-lr)
    elif kind == "signum":      # beta1 carries wd_lh
        g = g + wd * w
        s1.mul_(momentum).add_(g, alpha=-(1.0 - momentum))
        w.mul_(1.0 - lr * beta1).add_(torch.sign(s1), alpha=lr)
    elif kind == "signsgd":
        g = g + wd * w
        w.mul_(1.0 - lr * beta1).add_(torch.sign(g), alpha=-lr)
    elif kind == "adagrad":
        s1.addcmul_(g, g)
        w.sub_(lr * (g / (s1 + eps).sqrt() + wd * w))
    elif kind == "rmsprop":     # beta1 = gamma1, beta2 = clip_weights (<= 0: off)
        g = g + wd * w
        s1.mul_(beta1).addcmul_(g, g, value=1.0 - beta1)
        w.sub_(lr * g / (s1 + eps).sqrt())
        if beta2 > 0:
            w.clamp_(-beta2, beta2)
    elif kind == "adadelta":    # beta1 = rho
        s1.mul_(beta1).addcmul_(g, g, value=1.0 - beta1)
        d = (s2 + eps).sqrt() / (s1 + eps).sqrt() * g
        s2.mul_(beta1).addcmul_(d, d, value=1.0 - beta1)
        w.sub_(d + wd * w)
    elif kind == "ftrl":        # s1 = z, s2 = n, beta1 = lamda1, beta2 = beta
        s1.add_(g - ((s2 + g * g).sqrt() - s2.sqrt()) * w / lr)
        s2.addcmul_(g, g)
        w.copy_((torch.sign(s1) * beta1 - s1) / ((beta2 + s2.sqrt()) / lr + wd) * (s1.abs() > beta1))
    elif kind == "adamax":      # lr pre-divided by (1 - beta1^t)
        g = g + wd * w
        s1.mul_(beta1).add_(g, alpha=1.0 - beta1)
        torch.maximum(s2 * beta2, g.abs(), out=s2)
        w.sub_(lr * s1 / (s2 + 1e-38))
    else:
        raise ValueError(kind)
    if wout is not None:
        wout.reshape(-1).copy_(w.to(wout.dtype))


def _q2_shift(j: torch.Tensor) -> torch.Tensor:
    return 8 * (j // 4) + 6 - 2 * (j % 4)


def quantize_2bit(grad, residual, threshold: float):
    """Returns int32 words; updates residual in place (reference gradient_compression-inl.h:39-93)."""
    n = grad.numel()
    r = residual.reshape(-1)
    r += grad.float().reshape(-1)
    code = torch.zeros(n, dtype=torch.int64, device=grad.device)
    pos = r >= threshold
    neg = r <= -threshold
    code[pos] = 3
    code[neg & ~pos] = 2
    r[pos] -= threshold
    r[neg & ~pos] += threshold
    nwords = (n + 15) // 16
    pad = torch.zeros(nwords * 16, dtype=torch.int64, device=grad.device)
    pad[:n] = code
    j = torch.arange(16, device=grad.device)
    words = (pad.view(nwords, 16) << _q2_shift(j)).sum(1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)
    return words.to(torch.int32)


def dequantize_2bit(words, n: int, threshold: float):
    w = words.to(torch.int64) & 0xFFFFFFFF
    j = torch.arange(16, device=words.device)
    codes = (w.view(-1, 1) >> _q2_shift(j)) & 3
    vals = torch.zeros(codes.shape, dtype=torch.float32, device=words.device)
    vals[codes == 3] = threshold
    vals[codes == 2] = -threshold
    return vals.reshape(-1)[:n]


def normalize_u8_nhwc(src, mean, std):
    m = torch.tensor(mean, dtype=torch.float32, device=src.device)
    s = torch.tensor(std, dtype=torch.float32, device=src.device)
    return ((src.float() / 255.0 - m) / s).to(torch.bfloat16)
