"""A stand-in GPU for the host test tier.

Every Python wrapper around a CUDA kernel (`ops/*`, the trainer's fused optimizer ...) takes its CUDA branch only when its
tensors report `is_cuda`, which they never do on the CPU-only test box - so a wrong argument count, a tensor constructor
that does not exist, or a shape bug in such a branch would first show up on the GPU box.  `mock_gpu()` makes the host tier
walk those branches:

  * `torch.Tensor.is_cuda` reports True,
  * `native.lib()` returns a proxy whose `dlcfn_*` entries do NOT launch anything: they push every argument through the
    declared ctypes `argtypes` (count and convertibility, exactly what ctypes would check on the real call), log the call and
    return 0 (success) - except the GEMM / convolution / column-sum entry points, which first apply the argument checks their C
    counterparts make before touching the device (alignment, leading dimensions, channel multiples, geometry) and return the
    same error code, so that a shape a GPU test or a model would be refused with on the box is refused here too,
  * `native.stream_ptr()` returns the null stream.

Outputs of the "kernels" are whatever `torch.empty` left in the buffers; the tests built on this check control flow, shapes,
dtypes, strides and which kernels a code path launches - numerics stay with the real GPU tests.
"""
from __future__ import annotations

import contextlib
from unittest import mock

import torch


def _ptr_val(p) -> int:
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    import ctypes
    return ctypes.cast(p, ctypes.c_void_p).value or 0


def _check_gemm(A, lda, a_mn, b_shards, n_shards, rows_per_shard, ldb, b_mn, D, ldd, d_dtype, bias, M, N, K, accumulate, relu,
                block_n, num_sms, stats, stream) -> int:
    """The argument checks of dlcfn_gemm_bf16_ex (csrc/cuda/gemm_tcgen05.cu) up to its first CUDA call, plus the alignment
    rules cuTensorMapEncodeTiled enforces on the operands: 1 = cudaErrorInvalidValue, what the real entry point would return."""
    import ctypes
    if M <= 0 or N <= 0 or K <= 0:
        return 0
    if n_shards < 1 or n_shards > 8:
        return 1
    if (lda & 7) or (ldb & 7) or (_ptr_val(A) & 15):
        return 1
    arr = ctypes.cast(b_shards, ctypes.POINTER(ctypes.c_void_p))
    if any((arr[i] or 0) & 15 for i in range(n_shards)):
        return 1
    if (a_mn and lda < M) or (not a_mn and lda < K):
        return 1
    if n_shards == 1 and ((b_mn and ldb < N) or (not b_mn and ldb < K)):
        return 1
    if ldd < N:
        return 1
    tma_store = d_dtype == 1 and not accumulate and N >= 64 and (ldd & 7) == 0 and (_ptr_val(D) & 15) == 0
    if _ptr_val(stats) and not tma_store:
        return 1
    if not tma_store and d_dtype == 1 and ((ldd & 7) or (_ptr_val(D) & 15)) and N >= 8:
        return 1                        # the direct bf16 epilogue stores 8 columns (16 bytes) at a time
    return 0


def _check_conv(op, x, other, out, out_dtype, n, h, w, c, k, cin, r, s, pad, stride, oh, ow, accumulate, block_n, num_sms,
                xs_w, xs_h, xs_n, stats, stream) -> int:
    """The argument checks of dlcfn_conv_bf16_ex before its first CUDA call."""
    if op < 0 or op > 2 or min(n, h, w, r, s, stride) <= 0 or pad < 0:
        return 1
    if (c % 64) or (k % 8) or (cin % 8) or r * s > 49:
        return 1
    if (_ptr_val(x) & 15) or (_ptr_val(other) & 15) or (_ptr_val(out) & 15):
        return 1
    if _ptr_val(stats) and op != 0:
        return 1
    if op in (0, 2):
        P, Q = (h + 2 * pad - r) // stride + 1, (w + 2 * pad - s) // stride + 1
        if P != oh or Q != ow or c != cin:
            return 1
        if op == 0 and _ptr_val(stats) and not (out_dtype == 1 and not accumulate and k >= 64):
            return 1
        return 0
    if c != k:                          # dgrad: x = dy [n, h, w, k], out = dx [n, oh, ow, cin]
        return 1
    if h != (oh + 2 * pad - r) // stride + 1 or w != (ow + 2 * pad - s) // stride + 1 or r > 8 or s > 8:
        return 1
    return 0


def _check_colsum(x, out, M, N, ld, workspace, stream) -> int:
    return 1 if (M <= 0 or N <= 0 or (N & 7) or (ld & 7) or (_ptr_val(x) & 15)) else 0


def _check_batch_sum(x, out, N, inner, stream) -> int:
    return 1 if (N <= 0 or inner <= 0 or (inner & 7) or (_ptr_val(x) & 15) or (_ptr_val(out) & 15)) else 0


def _check_conv_plain(op, x, other, out, out_dtype, n, h, w, c, k, cin, r, s, pad, stride, oh, ow, accumulate, block_n, num_sms,
                      stream) -> int:
    """dlcfn_conv_bf16 = dlcfn_conv_bf16_ex with a dense NHWC view and no statistics."""
    return _check_conv(op, x, other, out, out_dtype, n, h, w, c, k, cin, r, s, pad, stride, oh, ow, accumulate, block_n, num_sms,
                       0, 0, 0, None, stream)


_VALIDATORS = {"dlcfn_gemm_bf16_ex": _check_gemm, "dlcfn_conv_bf16_ex": _check_conv, "dlcfn_conv_bf16": _check_conv_plain,
               "dlcfn_colsum_bf16": _check_colsum, "dlcfn_batch_sum_bf16": _check_batch_sum}


class _Checker:
    def __init__(self, real, name, log):
        object.__setattr__(self, "real", real)
        object.__setattr__(self, "name", name)
        object.__setattr__(self, "log", log)

    def __setattr__(self, k, v):            # argtypes / restype declarations land on the real function object
        setattr(self.real, k, v)

    def __getattr__(self, k):
        return getattr(self.real, k)

    def __call__(self, *args):
        at = self.real.argtypes
        assert at is not None, f"{self.name}: called without an argtypes declaration"
        assert len(args) == len(at), f"{self.name}: {len(args)} arguments passed, {len(at)} declared"
        for i, (a, t) in enumerate(zip(args, at)):
            try:
                t.from_param(a)
            except Exception as e:        # noqa: BLE001
                raise AssertionError(f"{self.name}: argument {i} = {a!r} is not convertible to {t.__name__}: {e}") from e
        self.log.append(self.name)
        v = _VALIDATORS.get(self.name)
        return int(v(*args)) if v is not None else 0


class _FakeLib:
    def __init__(self, real, log):
        object.__setattr__(self, "_real", real)
        object.__setattr__(self, "_log", log)
        object.__setattr__(self, "_cache", {})

    def __getattr__(self, name):
        real = getattr(self._real, name)
        if not name.startswith("dlcfn_"):
            return real
        if name not in self._cache:
            self._cache[name] = _Checker(real, name, self._log)
        return self._cache[name]

    def __setattr__(self, name, v):
        setattr(self._real, name, v)


@contextlib.contextmanager
def mock_gpu():
    """Yields the list of `dlcfn_*` functions "launched" inside the block, in order."""
    from deeplearning_cfn_b200.ops import native
    log: list[str] = []
    fake = _FakeLib(native.lib(), log)
    with mock.patch.object(native, "lib", lambda: fake), \
            mock.patch.object(native, "stream_ptr", lambda stream=None: 0), \
            mock.patch.object(torch.Tensor, "is_cuda", property(lambda self: True)):
        yield log
