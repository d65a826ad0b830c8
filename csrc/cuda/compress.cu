// Gradient compression + row-id utilities (sm_100a).
//
//  * 2-bit quantisation with error feedback  <- reference Q1/Q2
//      (gradient_compression-inl.h:39-140: 16 values -> one 32-bit word, codes 11=+thr 10=-thr 00=0,
//       value j of a block lives in byte j>>2 under mask 0xC0>>(2*(j&3)); residual carries the error)
//  * unique row ids for row-sparse pull       <- reference Q3 (kvstore_utils.cu:41-96, CUB sort+unique)
//
// B200 shape: one thread owns one output word = 16 consecutive inputs, read as 4x16B (fp32) or
// 2x16B (bf16) vector loads, so a warp streams 2 KiB contiguous per iteration; the residual is
// updated in the same pass.  Unique is done without a sort: row ids are bounded by the table's
// row count, so a presence bitmap + a flagged select (CUB) gives the sorted unique set in O(rows).
#include "common.cuh"
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

namespace dlcfn {

__device__ __forceinline__ uint32_t q2_code(float& r, float g, float thr) {
  r += g;
  if (r >= thr) { r -= thr; return 3u; }
  if (r <= -thr) { r += thr; return 2u; }
  return 0u;
}
__device__ __forceinline__ int q2_shift(int j) { return 8 * (j >> 2) + 6 - 2 * (j & 3); }

template <typename G>
__global__ void __launch_bounds__(256) kquant2(const G* __restrict__ grad, float* __restrict__ residual,
                                               uint32_t* __restrict__ out, size_t n, float thr) {
  size_t nwords = (n + 15) / 16;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < nwords; w += stride) {
    size_t base = w * 16;
    uint32_t word = 0;
    if (base + 16 <= n) {
      float g[16];
      if constexpr (sizeof(G) == 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float4 x = reinterpret_cast<const float4*>(grad + base)[k];
          g[4 * k] = x.x; g[4 * k + 1] = x.y; g[4 * k + 2] = x.z; g[4 * k + 3] = x.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          uint4 x = reinterpret_cast<const uint4*>(grad + base)[k];
          g[8 * k] = bf16_lo(x.x); g[8 * k + 1] = bf16_hi(x.x);
          g[8 * k + 2] = bf16_lo(x.y); g[8 * k + 3] = bf16_hi(x.y);
          g[8 * k + 4] = bf16_lo(x.z); g[8 * k + 5] = bf16_hi(x.z);
          g[8 * k + 6] = bf16_lo(x.w); g[8 * k + 7] = bf16_hi(x.w);
        }
      }
      float r[16];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 x = reinterpret_cast<const float4*>(residual + base)[k];
        r[4 * k] = x.x; r[4 * k + 1] = x.y; r[4 * k + 2] = x.z; r[4 * k + 3] = x.w;
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) word |= q2_code(r[j], g[j], thr) << q2_shift(j);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        reinterpret_cast<float4*>(residual + base)[k] =
            make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
    } else {
      for (int j = 0; base + j < n; ++j) {
        float g;
        if constexpr (sizeof(G) == 4) g = grad[base + j];
        else g = __bfloat162float(grad[base + j]);
        float r = residual[base + j];
        word |= q2_code(r, g, thr) << q2_shift(j);
        residual[base + j] = r;
      }
    }
    out[w] = word;
  }
}

// accumulate != 0: out += dequant(word)  (server-side merge of compressed pushes)
__global__ void __launch_bounds__(256) kdequant2(const uint32_t* __restrict__ in, float* __restrict__ out,
                                                 size_t n, float thr, int accumulate) {
  size_t nwords = (n + 15) / 16;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t w = blockIdx.x * (size_t)blockDim.x + threadIdx.x; w < nwords; w += stride) {
    uint32_t word = in[w];
    size_t base = w * 16;
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t c = (word >> q2_shift(j)) & 3u;
      v[j] = c == 3u ? thr : (c == 2u ? -thr : 0.f);
    }
    if (base + 16 <= n) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 o = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        if (accumulate) {
          float4 p = reinterpret_cast<float4*>(out + base)[k];
          o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
        }
        reinterpret_cast<float4*>(out + base)[k] = o;
      }
    } else {
      for (int j = 0; base + j < n; ++j) out[base + j] = accumulate ? out[base + j] + v[j] : v[j];
    }
  }
}

__global__ void __launch_bounds__(256) kmark_rows(const int64_t* __restrict__ ids, size_t n, int64_t num_rows,
                                                  uint8_t* __restrict__ present) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t r = ids[i];
    if (r >= 0 && r < num_rows) present[r] = 1;
  }
}

}  // namespace dlcfn

using namespace dlcfn;

static inline int grid1d(size_t n, int threads) {
  size_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

DLCFN_API int dlcfn_quantize_2bit(const void* grad, int grad_dtype, float* residual, uint32_t* out, size_t n,
                                  float threshold, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (n == 0) return 0;
  size_t nwords = (n + 15) / 16;
  if (grad_dtype == 0)
    kquant2<float><<<grid1d(nwords, 256), 256, 0, stream>>>((const float*)grad, residual, out, n, threshold);
  else if (grad_dtype == 1)
    kquant2<__nv_bfloat16><<<grid1d(nwords, 256), 256, 0, stream>>>((const __nv_bfloat16*)grad, residual, out,
                                                                   n, threshold);
  else
    return (int)cudaErrorInvalidValue;
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_dequantize_2bit(const uint32_t* in, float* out, size_t n, float threshold, int accumulate,
                                    cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (n == 0) return 0;
  kdequant2<<<grid1d((n + 15) / 16, 256), 256, 0, stream>>>(in, out, n, threshold, accumulate);
  return (int)cudaGetLastError();
}

// Sorted unique row ids.  `present` is a caller-provided num_rows-byte scratch, `temp` a scratch of
// at least dlcfn_unique_rows_temp_bytes(num_rows) bytes, `n_out` a device int64.
DLCFN_API size_t dlcfn_unique_rows_temp_bytes(int64_t num_rows) {
  size_t bytes = 0;
  cub::CountingInputIterator<int64_t> it(0);
  cub::DeviceSelect::Flagged((void*)nullptr, bytes, it, (const uint8_t*)nullptr, (int64_t*)nullptr,
                             (int64_t*)nullptr, num_rows);
  return bytes;
}

DLCFN_API int dlcfn_unique_rows(const int64_t* ids, size_t n, int64_t num_rows, uint8_t* present, void* temp,
                                size_t temp_bytes, int64_t* out, int64_t* n_out, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  DLCFN_CHECK(cudaMemsetAsync(present, 0, (size_t)num_rows, stream));
  if (n) kmark_rows<<<grid1d(n, 256), 256, 0, stream>>>(ids, n, num_rows, present);
  cub::CountingInputIterator<int64_t> it(0);
  DLCFN_CHECK(cub::DeviceSelect::Flagged(temp, temp_bytes, it, present, out, n_out, num_rows, stream));
  return (int)cudaGetLastError();
}
