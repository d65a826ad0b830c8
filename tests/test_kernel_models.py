"""Host transliterations of the index arithmetic of small kernels that were written after the round's GPU budget was spent
(csrc/cuda/elementwise.cu: kcolsum_partial / kcolsum_final, kbatch_sum_bf16): the same loops over (blockIdx, threadIdx) with
numpy in place of the loads, checked against the plain reduction.  An off-by-one in a stride, a slice that is visited twice or
never, or a partial-buffer index that collides would show up here; the numerics on the device are in tests/test_gpu_kernels.py."""
import numpy as np
import pytest


def _colsum_model(x: np.ndarray, ld: int) -> np.ndarray:
    """x: flat float array holding an [M, N] matrix with row stride ld (elements); returns out[N]."""
    M, N = x.shape
    flat = np.zeros(M * ld, dtype=np.float64)
    for r in range(M):
        flat[r * ld: r * ld + N] = x[r]
    N8, ld8 = N // 8, ld // 8
    s = (M + 63) // 64
    slices = 1 if s < 1 else (64 if s > 64 else s)
    grid_x, grid_y = (N8 + 31) // 32, slices
    part = np.full(slices * N, np.nan)                     # NaN = "never written": the final sum must not read such a slot
    visited = np.zeros((M, N8), dtype=np.int32)
    for by in range(grid_y):
        for bx in range(grid_x):
            sh = np.zeros((8, 32, 8))
            for tid in range(256):
                cx, ty = tid & 31, tid >> 5
                col8 = bx * 32 + cx
                a = np.zeros(8)
                if col8 < N8:
                    r = by * 8 + ty
                    step = grid_y * 8
                    while r + 3 * step < M:                # the 4-way unrolled loop
                        for j in range(4):
                            rr = r + j * step
                            a += flat[rr * ld8 * 8 + col8 * 8: rr * ld8 * 8 + col8 * 8 + 8]
                            visited[rr, col8] += 1
                        r += 4 * step
                    while r < M:
                        a += flat[r * ld8 * 8 + col8 * 8: r * ld8 * 8 + col8 * 8 + 8]
                        visited[r, col8] += 1
                        r += step
                sh[ty, cx] = a
            for tid in range(256):
                cx, ty = tid & 31, tid >> 5
                col8 = bx * 32 + cx
                if ty == 0 and col8 < N8:
                    for k in range(8):
                        part[by * N8 * 8 + col8 * 8 + k] = sh[:, cx, k].sum()
    assert (visited == 1).all(), "every (row, column vector) must be read exactly once"
    out = np.zeros(N)
    for c in range(N):                                     # kcolsum_final
        out[c] = sum(part[s_ * N + c] for s_ in range(slices))
    return out


@pytest.mark.parametrize("M,N,ld", [(1, 8, 8), (77, 1000, 1000), (256, 1000, 1000), (130, 72, 80), (4100, 264, 264), (64, 8, 16)])
def test_colsum_kernel_index_model(M, N, ld):
    rng = np.random.default_rng(M * 1000 + N)
    x = rng.standard_normal((M, N))
    got = _colsum_model(x, ld)
    assert not np.isnan(got).any()
    np.testing.assert_allclose(got, x.sum(0), rtol=1e-9, atol=1e-9)


def _batch_sum_model(x: np.ndarray, blocks_cap: int = 148 * 16) -> np.ndarray:
    """x [N, inner] -> out[inner]: 8 elements per thread, grid-stride over the inner vectors, 8 rows of loads in flight."""
    N, inner = x.shape
    inner8 = inner // 8
    blocks = min((inner8 + 255) // 256, blocks_cap)
    out = np.full(inner, np.nan)
    for b in range(blocks):
        for t in range(256):
            i = b * 256 + t
            while i < inner8:
                a = np.zeros(8)
                n = 0
                while n + 8 <= N:
                    for j in range(8):
                        a += x[n + j, i * 8: i * 8 + 8]
                    n += 8
                while n < N:
                    a += x[n, i * 8: i * 8 + 8]
                    n += 1
                assert np.isnan(out[i * 8]), "an output vector is written once"
                out[i * 8: i * 8 + 8] = a
                i += blocks * 256
    return out


@pytest.mark.parametrize("N,inner,cap", [(1, 8, 2368), (7, 192, 2368), (16, 4096, 2368), (19, 2056 * 8, 2), (256, 64, 2368)])
def test_batch_sum_kernel_index_model(N, inner, cap):
    rng = np.random.default_rng(N + inner)
    x = rng.standard_normal((N, inner))
    got = _batch_sum_model(x, cap)
    assert not np.isnan(got).any()
    np.testing.assert_allclose(got, x.sum(0), rtol=1e-9, atol=1e-9)
