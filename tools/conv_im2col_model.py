"""Host model of the implicit-GEMM convolution producer (csrc/cuda/gemm_tcgen05.cu, kgemm<..., CONV>).

`im2col_load` models one `cp.async.bulk.tensor.4d...im2col` as this code base understands it (cuTensorMapEncodeIm2col:
base positions inside the box [lower, extent-1+upper], traversed w -> h -> n by the traversal stride, each displaced by the
tap offset, out-of-tensor elements zero).  `fprop / dgrad / wgrad` then replay exactly the index arithmetic of the kernel's
producer warp (k-block -> tap / channel slice, tile -> first pixel, reversed taps for dgrad, column panels for wgrad) with
numpy tiles and are compared with torch's convolutions - so the host-visible part of the design is checked without a GPU;
what remains for the GPU run is whether the hardware matches `im2col_load`.

    python tools/conv_im2col_model.py
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def im2col_load(t, lower, upper, stride, c, w, h, n, off_w, off_h, pixels):
    """t: [N,H,W,C]; returns [pixels, 64]."""
    N, H, W, C = t.shape
    out = np.zeros((pixels, 64), dtype=np.float64)
    for i in range(pixels):
        ww, hh = w + off_w, h + off_h
        if 0 <= n < N and 0 <= hh < H and 0 <= ww < W:
            cc = min(64, max(0, C - c))
            out[i, :cc] = t[n, hh, ww, c:c + cc]
        w += stride
        if w > W - 1 + upper[0]:
            w = lower[0]
            h += stride
            if h > H - 1 + upper[1]:
                h = lower[1]
                n += 1
    return out


def _tile2d(mat, row0, rows, col0, cols):
    out = np.zeros((rows, cols))
    r1, c1 = min(mat.shape[0], row0 + rows), min(mat.shape[1], col0 + cols)
    if r1 > row0 and c1 > col0 and row0 >= 0 and col0 >= 0:
        out[:r1 - row0, :c1 - col0] = mat[row0:r1, col0:c1]
    return out


def fprop(x, wt, stride, pad, BN=128):
    N, H, W, C = x.shape
    K, R, S, _ = wt.shape
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    M, cb = N * P * Q, C // 64
    wm = wt.reshape(K, R * S * C)
    y = np.zeros((M, K))
    lower, upper = (-pad, -pad), (pad - (S - 1), pad - (R - 1))
    for m_idx in range(0, M, 128):
        q0, t0 = m_idx % Q, m_idx // Q
        p0, n0 = t0 % P, t0 // P
        for n_idx in range(0, K, BN):
            acc = np.zeros((128, BN))
            for kb in range(R * S * cb):
                rs, c0 = kb // cb, (kb % cb) * 64
                r, s = rs // S, rs % S
                a = im2col_load(x, lower, upper, stride, c0, q0 * stride + lower[0], p0 * stride + lower[1], n0, s, r, 128)
                b = _tile2d(wm, n_idx, BN, kb * 64, 64)
                acc += a @ b.T
            rows, cols = min(128, M - m_idx), min(BN, K - n_idx)
            y[m_idx:m_idx + rows, n_idx:n_idx + cols] = acc[:rows, :cols]
    return y.reshape(N, P, Q, K)


def dgrad(dy, wt, pad, H, W, BN=128):
    N, P, Q, K = dy.shape
    _, R, S, C = wt.shape
    M, cb = N * H * W, K // 64
    wm = wt.reshape(K, R * S * C)
    lower = (pad - (S - 1), pad - (R - 1))
    upper = (lower[0] + W - Q, lower[1] + H - P)
    dx = np.zeros((M, C))
    for m_idx in range(0, M, 128):
        q0, t0 = m_idx % W, m_idx // W
        p0, n0 = t0 % H, t0 // H
        for n_idx in range(0, C, BN):
            acc = np.zeros((128, BN))
            for kb in range(R * S * cb):
                rs, c0 = kb // cb, (kb % cb) * 64
                r, s = rs // S, rs % S
                a = im2col_load(dy, lower, upper, 1, c0, q0 + lower[0], p0 + lower[1], n0, s, r, 128)
                # MN-major B: [64 k rows = cout c0.., BN n cols] from W viewed [K][RS*C] at the reversed tap
                b = _tile2d(wm, c0, 64, (R * S - 1 - rs) * C + n_idx, BN)
                acc += a @ b
            rows, cols = min(128, M - m_idx), min(BN, C - n_idx)
            dx[m_idx:m_idx + rows, n_idx:n_idx + cols] = acc[:rows, :cols]
    return dx.reshape(N, H, W, C)


def wgrad(x, dy, R, S, stride, pad, BN=128):
    N, H, W, C = x.shape
    _, P, Q, K = dy.shape
    cb = C // 64
    Kdim, Ncols = N * P * Q, R * S * C
    dym = dy.reshape(Kdim, K)
    lower, upper = (-pad, -pad), (pad - (S - 1), pad - (R - 1))
    dw = np.zeros((K, Ncols))
    last_panel = R * S * cb - 1
    for m_idx in range(0, K, 128):
        for n_idx in range(0, Ncols, BN):
            acc = np.zeros((128, BN))
            for k_idx in range(0, Kdim, 64):
                q0, t0 = k_idx % Q, k_idx // Q
                p0, n0 = t0 % P, t0 // P
                a = _tile2d(dym, k_idx, 64, m_idx, 128)                 # MN-major A: [64 k, 128 m]
                b = np.zeros((64, BN))
                for j in range(BN // 64):
                    panel = min(n_idx // 64 + j, last_panel)
                    rs, c0 = panel // cb, (panel % cb) * 64
                    r, s = rs // S, rs % S
                    b[:, 64 * j:64 * j + 64] = im2col_load(x, lower, upper, stride, c0, q0 * stride + lower[0],
                                                           p0 * stride + lower[1], n0, s, r, 64)
                acc += a.T @ b
            rows, cols = min(128, K - m_idx), min(BN, Ncols - n_idx)
            dw[m_idx:m_idx + rows, n_idx:n_idx + cols] = acc[:rows, :cols]
    return dw.reshape(K, R, S, C)


def check(N, C, H, W, K, R, stride, pad, BN=128):
    g = torch.Generator().manual_seed(N * 1000 + H)
    x = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(K, C, R, R, generator=g, dtype=torch.float64)
    y = F.conv2d(x, w, None, stride, pad)
    dy = torch.randn(y.shape, generator=g, dtype=torch.float64)
    xn, wn = x.permute(0, 2, 3, 1).numpy(), w.permute(0, 2, 3, 1).numpy()          # NHWC, KRSC
    dyn = dy.permute(0, 2, 3, 1).numpy()
    errs = {"fprop": np.abs(fprop(xn, wn, stride, pad, BN) - y.permute(0, 2, 3, 1).numpy()).max()}
    dwr = torch.nn.grad.conv2d_weight(x, w.shape, dy, stride, pad).permute(0, 2, 3, 1).numpy()
    errs["wgrad"] = np.abs(wgrad(xn, dyn, R, R, stride, pad, BN) - dwr).max()
    if stride == 1 and K % 64 == 0:      # dgrad contracts over (tap, 64 output channels): Cout % 64 == 0
        dxr = torch.nn.grad.conv2d_input(x.shape, w, dy, 1, pad).permute(0, 2, 3, 1).numpy()
        errs["dgrad"] = np.abs(dgrad(dyn, wn, pad, H, W, BN) - dxr).max()
    return errs


CASES = [(1, 64, 7, 9, 128, 3, 1, 1), (1, 192, 5, 5, 64, 5, 1, 2), (2, 64, 9, 13, 72, 3, 1, 1), (1, 128, 6, 6, 64, 3, 1, 1), (2, 64, 15, 15, 64, 3, 2, 1), (1, 64, 12, 12, 96, 5, 1, 2),
         (2, 64, 10, 10, 64, 1, 1, 0), (2, 64, 16, 16, 128, 1, 2, 0), (2, 64, 8, 8, 64, 3, 1, 0), (1, 64, 14, 14, 200, 3, 2, 1)]

if __name__ == "__main__":
    for c in CASES:
        print(c, {k: f"{v:.2e}" for k, v in check(*c).items()})
