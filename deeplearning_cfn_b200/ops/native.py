"""ctypes bindings for the hand-written sm_100a kernel library (csrc/cuda -> _lib/libdlcfn_cuda.so).

There is exactly one code path on a GPU: these kernels.  If the shared object is missing on a
machine with a GPU the import fails loudly (no silent PyTorch fallback); on a CPU-only machine the
ops raise `NativeUnavailable` when called, and pure-PyTorch *reference* implementations used by the
numerics tests live in `ops/reference.py`.
"""
from __future__ import annotations

import ctypes
import threading
from ctypes import c_float, c_int, c_longlong, c_size_t, c_uint32, c_void_p
from pathlib import Path

import torch

_LIB_PATH = Path(__file__).resolve().parent.parent / "_lib" / "libdlcfn_cuda.so"
_lib = None
_lock = threading.Lock()

# number of dlcfn kernels launched by this process (bench.py reports it as gpu_launches)
launch_count = 0


class NativeUnavailable(RuntimeError):
    pass


DT_F32, DT_BF16 = 0, 1
OPT_KINDS = {"sgd": 0, "sgd_mom": 1, "adam": 2, "test": 3, "none": 4, "nag": 5, "signum": 6, "signsgd": 7,
             "adagrad": 8, "rmsprop": 9, "adadelta": 10, "ftrl": 11, "adamax": 12}
OPT_USES_S1 = {"sgd_mom", "adam", "nag", "signum", "adagrad", "rmsprop", "adadelta", "ftrl", "adamax"}
OPT_USES_S2 = {"adam", "adadelta", "ftrl", "adamax"}


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return DT_F32
    if dt == torch.bfloat16:
        return DT_BF16
    raise TypeError(f"dlcfn kernels support float32 and bfloat16, got {dt}")


def lib_path() -> Path:
    return _LIB_PATH


def available() -> bool:
    return _LIB_PATH.exists() and torch.cuda.is_available()


def lib():
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not _LIB_PATH.exists():
            raise NativeUnavailable(
                f"{_LIB_PATH} is missing: run `python -m deeplearning_cfn_b200._build` (needs nvcc); "
                "the CUDA path has no PyTorch fallback")
        L = ctypes.CDLL(str(_LIB_PATH))
        _declare(L)
        _lib = L
    return _lib


def _declare(L) -> None:
    vp, sz, i, f, u32, ll = c_void_p, c_size_t, c_int, c_float, c_uint32, c_longlong
    L.dlcfn_elementwise_sum.argtypes = [vp, i, vp, sz, i, f, vp]
    L.dlcfn_cast_scale.argtypes = [vp, vp, sz, i, i, f, vp]
    L.dlcfn_fused_optimizer.argtypes = [i, vp, i, vp, vp, vp, vp, i, sz, f, f, f, f, f, f, f, f, vp]
    L.dlcfn_normalize_u8_nhwc.argtypes = [vp, vp, sz, vp, vp, vp]
    L.dlcfn_quantize_2bit.argtypes = [vp, i, vp, vp, sz, f, vp]
    L.dlcfn_dequantize_2bit.argtypes = [vp, vp, sz, f, i, vp]
    L.dlcfn_unique_rows_temp_bytes.argtypes = [ll]
    L.dlcfn_unique_rows_temp_bytes.restype = sz
    L.dlcfn_unique_rows.argtypes = [vp, sz, ll, vp, vp, sz, vp, vp, vp]
    L.dlcfn_symm_create.argtypes = [sz, i, i, ctypes.POINTER(vp)]
    L.dlcfn_symm_get_handle.argtypes = [vp, vp]
    L.dlcfn_symm_create_external.argtypes = [vp, sz, i, i, ctypes.POINTER(vp)]
    L.dlcfn_symm_set_multicast.argtypes = [vp, vp]
    L.dlcfn_symm_has_multicast.argtypes = [vp]
    L.dlcfn_allreduce_ex.argtypes = [vp, sz, sz, sz, i, f, i, i, i, vp]
    L.dlcfn_symm_open_peer.argtypes = [vp, i, vp]
    L.dlcfn_symm_set_peer_ptr.argtypes = [vp, i, vp]
    L.dlcfn_symm_allocation.argtypes = [vp]
    L.dlcfn_symm_allocation.restype = vp
    L.dlcfn_symm_finalize.argtypes = [vp]
    L.dlcfn_symm_heap_ptr.argtypes = [vp, i]
    L.dlcfn_symm_heap_ptr.restype = vp
    L.dlcfn_symm_check_error.argtypes = [vp, vp, ctypes.POINTER(i)]
    ull = ctypes.c_ulonglong
    P = ctypes.POINTER
    L.dlcfn_vmm_supported.argtypes = [i]
    L.dlcfn_vmm_supported.restype = c_int
    L.dlcfn_vmm_granularity.argtypes = [i, i, i]
    L.dlcfn_vmm_granularity.restype = sz
    L.dlcfn_vmm_alloc.argtypes = [sz, i, sz, P(vp), P(ull), P(c_int)]
    L.dlcfn_vmm_import.argtypes = [i, sz, i, sz, P(vp), P(ull)]
    L.dlcfn_vmm_mc_create.argtypes = [sz, i, P(ull), P(c_int)]
    L.dlcfn_vmm_mc_import.argtypes = [i, P(ull)]
    L.dlcfn_vmm_mc_add_device.argtypes = [ull, i]
    L.dlcfn_vmm_mc_bind_map.argtypes = [ull, ull, sz, i, sz, P(vp)]
    L.dlcfn_vmm_unmap_release.argtypes = [vp, ull, sz]
    L.dlcfn_probe_fabric.argtypes = [i, P(c_int), P(c_int), P(c_int), P(c_int)]
    for fn in ("dlcfn_vmm_alloc", "dlcfn_vmm_import", "dlcfn_vmm_mc_create", "dlcfn_vmm_mc_import", "dlcfn_vmm_mc_add_device",
               "dlcfn_vmm_mc_bind_map", "dlcfn_vmm_unmap_release", "dlcfn_probe_fabric"):
        getattr(L, fn).restype = c_int
    L.dlcfn_symm_set_debug.argtypes = [vp, i, u32, i]
    L.dlcfn_symm_set_debug.restype = c_int
    L.dlcfn_symm_poll_error_async.argtypes = [vp, vp, vp]
    L.dlcfn_symm_poll_error_async.restype = c_int
    L.dlcfn_symm_destroy.argtypes = [vp]
    L.dlcfn_barrier.argtypes = [vp, i, vp]
    L.dlcfn_allreduce.argtypes = [vp, sz, sz, i, f, i, i, vp]
    L.dlcfn_broadcast.argtypes = [vp, sz, sz, i, i, i, vp]
    L.dlcfn_ps_sync_step.argtypes = [vp, sz, sz, sz, sz, sz, u32, vp, vp, vp, i, i, f, f, f, f, f, f, f, f, i, i, vp]
    L.dlcfn_ps_async_worker.argtypes = [vp, u32, i, i, vp]
    L.dlcfn_ps_sync_step_ex.argtypes = [vp, sz, sz, sz, sz, sz, u32, u32, vp, vp, vp, i, i, f, f, f, f, f, f, f, f, i, i, i, vp]
    L.dlcfn_ps_sync_step_ex.restype = c_int
    L.dlcfn_ps_async_server.argtypes = [vp, sz, sz, sz, sz, u32, vp, vp, vp, i, i, f, f, f, f, f, f, f, f, i, i, vp]
    L.dlcfn_gemm_bf16.argtypes = [vp, ll, i, vp, i, i, ll, i, vp, ll, i, vp, i, i, i, i, i, i, i, vp]
    L.dlcfn_conv_bf16.argtypes = [i, vp, vp, vp, i] + [i] * 15 + [vp]
    L.dlcfn_conv_bf16.restype = c_int
    L.dlcfn_conv_bf16_ex.argtypes = [i, vp, vp, vp, i] + [i] * 15 + [ll, ll, ll, vp, vp]
    L.dlcfn_conv_bf16_ex.restype = c_int
    L.dlcfn_gemm_bf16_ex.argtypes = [vp, ll, i, vp, i, i, ll, i, vp, ll, i, vp, i, i, i, i, i, i, i, vp, vp]
    L.dlcfn_gemm_bf16_ex.restype = c_int
    L.dlcfn_colsum_slices.argtypes = [ll]
    L.dlcfn_colsum_slices.restype = c_int
    L.dlcfn_colsum_bf16.argtypes = [vp, vp, ll, i, ll, vp, vp]
    L.dlcfn_colsum_bf16.restype = c_int
    L.dlcfn_batch_sum_bf16.argtypes = [vp, vp, i, ll, vp]
    L.dlcfn_batch_sum_bf16.restype = c_int
    L.dlcfn_gemm_bf16_staged.argtypes = [vp, vp, ll, i, sz, sz, sz, u32, i, i, i, vp, ll, i, vp, i, i, i, i, i, i, i, vp]
    L.dlcfn_gemm_bf16_staged.restype = c_int
    L.dlcfn_fastdiv_host.argtypes = [i, i]
    L.dlcfn_fastdiv_host.restype = c_int
    L.dlcfn_gemm_stats_rows.argtypes = []
    L.dlcfn_gemm_stats_rows.restype = c_int
    for fn in ("dlcfn_elementwise_sum", "dlcfn_cast_scale", "dlcfn_fused_optimizer", "dlcfn_normalize_u8_nhwc",
               "dlcfn_quantize_2bit", "dlcfn_dequantize_2bit", "dlcfn_unique_rows", "dlcfn_symm_create",
               "dlcfn_symm_get_handle", "dlcfn_symm_open_peer", "dlcfn_symm_set_peer_ptr", "dlcfn_symm_finalize",
               "dlcfn_symm_check_error", "dlcfn_symm_destroy", "dlcfn_barrier", "dlcfn_allreduce",
               "dlcfn_broadcast", "dlcfn_ps_sync_step", "dlcfn_ps_async_worker", "dlcfn_ps_async_server",
               "dlcfn_gemm_bf16", "dlcfn_symm_create_external", "dlcfn_symm_set_multicast", "dlcfn_symm_has_multicast",
               "dlcfn_allreduce_ex"):
        getattr(L, fn).restype = c_int


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"dlcfn native call {what} failed with CUDA error {rc}")


def stream_ptr(stream: torch.cuda.Stream | None = None) -> int:
    s = stream if stream is not None else torch.cuda.current_stream()
    return s.cuda_stream


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------- elementwise
def elementwise_sum(inputs: list[torch.Tensor], out: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """out = scale * sum_i inputs[i]  (reference C5 ElementwiseSum, ndarray_function-inl.h:388-433)."""
    n = out.numel()
    for t in inputs:
        assert t.is_cuda and t.is_contiguous() and t.numel() == n and t.dtype == out.dtype
    arr = (c_void_p * len(inputs))(*[t.data_ptr() for t in inputs])
    check(lib().dlcfn_elementwise_sum(ctypes.cast(arr, c_void_p), len(inputs), out.data_ptr(), n,
                                      dtype_code(out.dtype), float(scale), stream_ptr()), "elementwise_sum")
    _count()
    return out


def cast_scale(src: torch.Tensor, dst: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    assert src.is_contiguous() and dst.is_contiguous() and src.numel() == dst.numel()
    check(lib().dlcfn_cast_scale(src.data_ptr(), dst.data_ptr(), src.numel(), dtype_code(src.dtype),
                                 dtype_code(dst.dtype), float(scale), stream_ptr()), "cast_scale")
    _count()
    return dst


def fused_optimizer(kind: str, grad: torch.Tensor, master: torch.Tensor, s1: torch.Tensor | None,
                    s2: torch.Tensor | None, wout: torch.Tensor | None, *, lr: float, wd: float = 0.0,
                    momentum: float = 0.0, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                    rescale: float = 1.0, clip: float = -1.0) -> None:
    """One launch over a flat parameter buffer (reference O1/O2, optimizer_op-inl.h:86-427, 840-875)."""
    n = master.numel()
    assert master.dtype == torch.float32 and grad.numel() == n
    check(lib().dlcfn_fused_optimizer(OPT_KINDS[kind], grad.data_ptr(), dtype_code(grad.dtype), master.data_ptr(),
                                      _ptr(s1), _ptr(s2), _ptr(wout),
                                      dtype_code(wout.dtype) if wout is not None else 0, n, lr, wd, momentum,
                                      beta1, beta2, eps, rescale, clip, stream_ptr()), "fused_optimizer")
    _count()


def colsum(x: torch.Tensor) -> torch.Tensor:
    """fp32 column sums of a bf16 [M, N] matrix (bias gradients), deterministic two-stage reduction at streaming speed."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1
    M, N = x.shape
    L = lib()
    ws = torch.empty(L.dlcfn_colsum_slices(M) * N, dtype=torch.float32, device=x.device)
    out = torch.empty(N, dtype=torch.float32, device=x.device)
    check(L.dlcfn_colsum_bf16(x.data_ptr(), out.data_ptr(), M, N, x.stride(0), ws.data_ptr(), stream_ptr()), "colsum_bf16")
    _count(2)
    return out


def batch_sum(x: torch.Tensor) -> torch.Tensor:
    """fp32 sum over dim 0 of a bf16 tensor whose dims 1.. are dense in memory (any memory format): one streaming pass."""
    assert x.is_cuda and x.dtype == torch.bfloat16
    N = x.shape[0]
    inner = x[0].numel()
    assert x.stride(0) == inner and inner % 8 == 0
    out = torch.empty_strided(x.shape[1:], x.stride()[1:], dtype=torch.float32, device=x.device)
    check(lib().dlcfn_batch_sum_bf16(x.data_ptr(), out.data_ptr(), N, inner, stream_ptr()), "batch_sum_bf16")
    _count()
    return out


def normalize_u8_nhwc(src: torch.Tensor, dst: torch.Tensor, mean, std) -> torch.Tensor:
    assert src.dtype == torch.uint8 and dst.dtype == torch.bfloat16 and src.numel() == dst.numel()
    m = (c_float * 3)(*mean)
    s = (c_float * 3)(*std)
    check(lib().dlcfn_normalize_u8_nhwc(src.data_ptr(), dst.data_ptr(), src.numel(), ctypes.cast(m, c_void_p),
                                        ctypes.cast(s, c_void_p), stream_ptr()), "normalize_u8_nhwc")
    _count()
    return dst


def quantize_2bit(grad: torch.Tensor, residual: torch.Tensor, out_words: torch.Tensor, threshold: float) -> None:
    n = grad.numel()
    assert residual.dtype == torch.float32 and residual.numel() == n
    assert out_words.dtype == torch.int32 and out_words.numel() == (n + 15) // 16
    check(lib().dlcfn_quantize_2bit(grad.data_ptr(), dtype_code(grad.dtype), residual.data_ptr(),
                                    out_words.data_ptr(), n, threshold, stream_ptr()), "quantize_2bit")
    _count()


def dequantize_2bit(words: torch.Tensor, out: torch.Tensor, threshold: float, accumulate: bool = False) -> None:
    assert out.dtype == torch.float32
    check(lib().dlcfn_dequantize_2bit(words.data_ptr(), out.data_ptr(), out.numel(), threshold, int(accumulate),
                                      stream_ptr()), "dequantize_2bit")
    _count()


def unique_rows(ids: torch.Tensor, num_rows: int) -> torch.Tensor:
    """Sorted unique row ids (reference Q3, kvstore_utils.cu:41-96)."""
    assert ids.dtype == torch.int64 and ids.is_cuda
    L = lib()
    tb = L.dlcfn_unique_rows_temp_bytes(num_rows)
    present = torch.empty(num_rows, dtype=torch.uint8, device=ids.device)
    temp = torch.empty(max(tb, 1), dtype=torch.uint8, device=ids.device)
    out = torch.empty(num_rows, dtype=torch.int64, device=ids.device)
    n_out = torch.zeros(1, dtype=torch.int64, device=ids.device)
    check(L.dlcfn_unique_rows(ids.data_ptr(), ids.numel(), num_rows, present.data_ptr(), temp.data_ptr(), tb,
                              out.data_ptr(), n_out.data_ptr(), stream_ptr()), "unique_rows")
    _count(2)
    return out[: int(n_out.item())]


# --------------------------------------------------------------------------- GEMM
def gemm_bf16(a: torch.Tensor, b, *, a_mn: bool = False, b_mn: bool = False, out: torch.Tensor | None = None,
              out_dtype: torch.dtype = torch.bfloat16, bias: torch.Tensor | None = None, accumulate: bool = False,
              relu: bool = False, m: int | None = None, n: int | None = None, k: int | None = None,
              shard_ptrs: list[int] | None = None, rows_per_shard: int = 0, ldb: int | None = None,
              block_n: int = 0, stats: torch.Tensor | None = None) -> torch.Tensor:
    """D[M,N] = A[M,K] @ B[N,K]^T (+bias) on the tcgen05 tensor cores.

    a_mn / b_mn select the MN-major storage ([K, rows]) of an operand.  With `shard_ptrs`, B lives
    row-sharded in peer HBM and is pulled through TMA inside the kernel (the fused pull + GEMM path).
    """
    assert a.dtype == torch.bfloat16 and a.is_cuda and a.dim() == 2 and a.stride(1) == 1
    if not a_mn:
        M, K = a.shape
    else:
        K, M = a.shape
    if shard_ptrs is None:
        assert b.dtype == torch.bfloat16 and b.dim() == 2 and b.stride(1) == 1
        if not b_mn:
            N, Kb = b.shape
        else:
            Kb, N = b.shape
        assert Kb == K, (a.shape, b.shape, a_mn, b_mn)
        ptrs = [b.data_ptr()]
        ldb_ = b.stride(0)
    else:
        assert n is not None and ldb is not None
        N = n
        ptrs = list(shard_ptrs)
        ldb_ = ldb
    M = m if m is not None else M
    K = k if k is not None else K
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=a.device)
    assert out.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() >= N
    arr = (c_void_p * len(ptrs))(*ptrs)
    rc = lib().dlcfn_gemm_bf16_ex(a.data_ptr(), a.stride(0), int(a_mn), ctypes.cast(arr, c_void_p), len(ptrs),
                                  rows_per_shard, ldb_, int(b_mn), out.data_ptr(), out.stride(0),
                                  dtype_code(out.dtype), _ptr(bias), M, N, K, int(accumulate), int(relu), block_n, 0,
                                  _ptr(stats), stream_ptr())
    if rc != 0:
        raise RuntimeError(f"dlcfn gemm_bf16 failed with CUDA error {rc}: M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn} "
                           f"lda={a.stride(0)} ldb={ldb_} ldd={out.stride(0)} a%16={a.data_ptr() % 16} "
                           f"b%16={ptrs[0] % 16} d%16={out.data_ptr() % 16} shards={len(ptrs)} block_n={block_n}")
    _count()
    return out


def gemm_bf16_staged(heap_ctx, a: torch.Tensor, *, shard_off: int, stage_off: int, flags_off: int, seq: int, rows_per_shard: int,
                     cols: int, b_mn: bool, n: int, k: int, out: torch.Tensor | None = None, bias: torch.Tensor | None = None,
                     a_mn: bool = False, block_n: int = 0, channel: int = -1) -> torch.Tensor:
    """D = A * B^T with B row-sharded over the ranks of a symmetric heap: the all-gather happens INSIDE the GEMM (every rank
    multicasts its shard into all stage buffers through the NVSwitch, the TMA producer consumes chunks as their arrival counters
    tick - csrc/cuda/gemm_tcgen05.cu, GemmParams::sg).  Collective: every rank launches the same call with the same `seq`."""
    assert a.dtype == torch.bfloat16 and a.is_cuda and a.dim() == 2 and a.stride(1) == 1
    M = a.shape[1] if a_mn else a.shape[0]
    if out is None:
        out = torch.empty(M, n, dtype=torch.bfloat16, device=a.device)
    rc = lib().dlcfn_gemm_bf16_staged(heap_ctx, a.data_ptr(), a.stride(0), int(a_mn), shard_off, stage_off, flags_off, int(seq),
                                      rows_per_shard, cols, int(b_mn), out.data_ptr(), out.stride(0), dtype_code(out.dtype),
                                      _ptr(bias), M, n, k, 0, 0, block_n, channel, stream_ptr())
    if rc != 0:
        raise RuntimeError(f"dlcfn gemm_bf16_staged failed with CUDA error {rc}: M={M} N={n} K={k} b_mn={b_mn} "
                           f"rows_per_shard={rows_per_shard} cols={cols} seq={seq}")
    _count()
    return out


# ------------------------------------------------------------------ implicit-GEMM convolution (EXPERIMENTAL)
def _nhwc_ok(t: torch.Tensor) -> bool:
    return t.dim() == 4 and t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous(memory_format=torch.channels_last)


_stats_rows = 0


def stats_rows() -> int:
    """Rows of the partial-statistics buffer the GEMM / convolution epilogues write (4 per SM)."""
    global _stats_rows
    if not _stats_rows:
        _stats_rows = int(lib().dlcfn_gemm_stats_rows())
    return _stats_rows


def new_stats(channels: int, device) -> torch.Tensor:
    """Uninitialised [rows, 2, C] fp32 buffer for the fused BatchNorm statistics (the launch zero-fills it)."""
    return torch.empty((stats_rows(), 2, channels), dtype=torch.float32, device=device)


def conv_fprop(x: torch.Tensor, weight: torch.Tensor, stride: int, pad: int, block_n: int = 0,
               stats: torch.Tensor | None = None) -> torch.Tensor:
    """x [N,C,H,W] channels_last, weight [K,C,R,S] channels_last (memory [K][R][S][C]) -> y [N,K,P,Q] channels_last."""
    assert _nhwc_ok(x) and _nhwc_ok(weight)
    N, C, H, W = x.shape
    K, _, R, S = weight.shape
    P, Q = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    y = torch.empty((N, K, P, Q), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    rc = lib().dlcfn_conv_bf16_ex(0, x.data_ptr(), weight.data_ptr(), y.data_ptr(), 1, N, H, W, C, K, C, R, S, pad, stride, P, Q,
                                  0, block_n, 0, 0, 0, 0, _ptr(stats), stream_ptr())
    if rc != 0:
        raise RuntimeError(f"dlcfn conv fprop failed with CUDA error {rc}: x={tuple(x.shape)} w={tuple(weight.shape)} "
                           f"stride={stride} pad={pad}")
    _count()
    return y


def conv_dgrad(dy: torch.Tensor, weight: torch.Tensor, pad: int, in_hw: tuple[int, int], block_n: int = 0,
               stride: int = 1) -> torch.Tensor:
    """Input gradient: dy [N,K,P,Q] channels_last, weight [K,C,R,S] channels_last -> dx [N,C,H,W] channels_last.  A strided
    convolution's gradient runs as one implicit GEMM per output-parity class (csrc/cuda/gemm_tcgen05.cu:dlcfn_conv_bf16)."""
    assert _nhwc_ok(dy) and _nhwc_ok(weight)
    N, K, P, Q = dy.shape
    _, C, R, S = weight.shape
    H, W = in_hw
    dx = torch.empty((N, C, H, W), dtype=torch.bfloat16, device=dy.device, memory_format=torch.channels_last)
    rc = lib().dlcfn_conv_bf16(1, dy.data_ptr(), weight.data_ptr(), dx.data_ptr(), 1, N, P, Q, K, K, C, R, S, pad, stride, H, W,
                               0, block_n, 0, stream_ptr())
    if rc != 0:
        raise RuntimeError(f"dlcfn conv dgrad failed with CUDA error {rc}: dy={tuple(dy.shape)} w={tuple(weight.shape)} pad={pad} "
                           f"stride={stride} in_hw={in_hw}")
    _count(stride * stride)
    return dx


def conv_wgrad(x: torch.Tensor, dy: torch.Tensor, kernel: tuple[int, int], stride: int, pad: int,
               block_n: int = 0) -> torch.Tensor:
    """dw [K,C,R,S] channels_last = dy^T * im2col(x); split-K over the N*P*Q pixels, deterministic reduction."""
    assert _nhwc_ok(x) and _nhwc_ok(dy)
    N, C, H, W = x.shape
    _, K, P, Q = dy.shape
    R, S = kernel
    dw = torch.empty((K, C, R, S), dtype=torch.bfloat16, device=x.device, memory_format=torch.channels_last)
    rc = lib().dlcfn_conv_bf16(2, x.data_ptr(), dy.data_ptr(), dw.data_ptr(), 1, N, H, W, C, K, C, R, S, pad, stride, P, Q,
                               0, block_n, 0, stream_ptr())
    if rc != 0:
        raise RuntimeError(f"dlcfn conv wgrad failed with CUDA error {rc}: x={tuple(x.shape)} dy={tuple(dy.shape)} "
                           f"kernel={kernel} stride={stride} pad={pad}")
    _count(2)
    return dw


# ------------------------------------------------------------------ ResNet stem: 4x4 / stride 1 over the 16-channel s2d image
def _stem_view(s2d: torch.Tensor):
    """The space-to-depth image [N,16,Hs,Ws] (channels_last) seen as [N, Hs, Ws-3, 64] with a W stride of 16 elements:
    one 64-channel "pixel" = 4 adjacent columns x 16 channels = one 128-byte TMA row, so the 4x4 filter becomes 4 (rows) x 1
    taps of the implicit GEMM (dlcfn_conv_bf16_ex, overlapping view)."""
    assert _nhwc_ok(s2d) and s2d.shape[1] == 16
    N, _, Hs, Ws = s2d.shape
    return N, Hs, Ws - 3, 16, Ws * 16, Hs * Ws * 16


def stem_conv_fprop(s2d: torch.Tensor, w4: torch.Tensor, stats: torch.Tensor | None = None) -> torch.Tensor:
    """s2d [N,16,Hs,Ws], w4 [K,16,4,4] (both channels_last; w4 memory [K][r][s][c] = [K][4][64]) -> y [N,K,Hs-3,Ws-3]."""
    assert _nhwc_ok(w4) and tuple(w4.shape[1:]) == (16, 4, 4)
    N, H, Wv, xs_w, xs_h, xs_n = _stem_view(s2d)
    K = w4.shape[0]
    P, Q = H - 3, Wv
    y = torch.empty((N, K, P, Q), dtype=torch.bfloat16, device=s2d.device, memory_format=torch.channels_last)
    rc = lib().dlcfn_conv_bf16_ex(0, s2d.data_ptr(), w4.data_ptr(), y.data_ptr(), 1, N, H, Wv, 64, K, 64, 4, 1, 0, 1, P, Q,
                                  0, 0, 0, xs_w, xs_h, xs_n, _ptr(stats), stream_ptr())
    if rc != 0:
        raise RuntimeError(f"dlcfn stem conv fprop failed with CUDA error {rc}: s2d={tuple(s2d.shape)} w4={tuple(w4.shape)}")
    _count()
    return y


def stem_conv_wgrad(s2d: torch.Tensor, dy: torch.Tensor) -> torch.Tensor:
    """dw4 [K,16,4,4] channels_last = dy^T * im2col(s2d) through the same overlapping view (split-K, deterministic)."""
    assert _nhwc_ok(dy)
    N, H, Wv, xs_w, xs_h, xs_n = _stem_view(s2d)
    K = dy.shape[1]
    P, Q = H - 3, Wv
    assert tuple(dy.shape) == (N, K, P, Q)
    dw = torch.empty((K, 16, 4, 4), dtype=torch.bfloat16, device=s2d.device, memory_format=torch.channels_last)
    rc = lib().dlcfn_conv_bf16_ex(2, s2d.data_ptr(), dy.data_ptr(), dw.data_ptr(), 1, N, H, Wv, 64, K, 64, 4, 1, 0, 1, P, Q,
                                  0, 0, 0, xs_w, xs_h, xs_n, None, stream_ptr())
    if rc != 0:
        raise RuntimeError(f"dlcfn stem conv wgrad failed with CUDA error {rc}: s2d={tuple(s2d.shape)} dy={tuple(dy.shape)}")
    _count(2)
    return dw
