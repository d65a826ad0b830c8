// Memory-bound elementwise kernels of the dlcfn runtime (sm_100a).
//
//  * N-way elementwise sum           <- reference C5  (ndarray_function-inl.h:388-433)
//  * cast/scale of gradient buckets  <- Horovod fusion-buffer pack step (run.sh:70-81)
//  * fused multi-precision SGD / SGD-momentum / Adam over flat buffers
//                                    <- reference O1/O2 (optimizer_op-inl.h:86-427, 840-875)
//  * uint8 NHWC image -> normalised bf16 NHWC (input pipeline tail, fused cast+normalise)
//
// Design: every kernel is a 16-byte-vectorised grid-stride loop sized to
// 148 SMs x resident CTAs; fp32 math, bf16 (or fp32) storage.  HBM3e is the
// bound, so each kernel touches every byte exactly once.
#include "common.cuh"
#include "optimizer.cuh"

namespace dlcfn {

static inline int grid_for(size_t nvec, int threads, int max_blocks = 148 * 8) {
  size_t b = (nvec + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > (size_t)max_blocks) b = max_blocks;
  return (int)b;
}

// ------------------------------------------------------------------ sum -----
struct SumArgs {
  const void* in[DLCFN_MAX_RANKS * 2];
  int n_in;
};

template <typename T>
__global__ void __launch_bounds__(256) ksum(SumArgs a, T* __restrict__ out, size_t n, float scale);

template <>
__global__ void __launch_bounds__(256) ksum<float>(SumArgs a, float* __restrict__ out, size_t n,
                                                    float scale) {
  size_t nvec = n / 4;
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = tid; v < nvec; v += stride) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int k = 0; k < a.n_in; ++k) {
      float4 x = reinterpret_cast<const float4*>(a.in[k])[v];
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale;
    reinterpret_cast<float4*>(out)[v] = acc;
  }
  for (size_t i = nvec * 4 + tid; i < n; i += stride) {
    float acc = 0.f;
    for (int k = 0; k < a.n_in; ++k) acc += reinterpret_cast<const float*>(a.in[k])[i];
    out[i] = acc * scale;
  }
}

template <>
__global__ void __launch_bounds__(256) ksum<__nv_bfloat16>(SumArgs a, __nv_bfloat16* __restrict__ out,
                                                            size_t n, float scale) {
  size_t nvec = n / 8;
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = tid; v < nvec; v += stride) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (int k = 0; k < a.n_in; ++k) {
      uint4 x = reinterpret_cast<const uint4*>(a.in[k])[v];
      acc[0] += bf16_lo(x.x); acc[1] += bf16_hi(x.x);
      acc[2] += bf16_lo(x.y); acc[3] += bf16_hi(x.y);
      acc[4] += bf16_lo(x.z); acc[5] += bf16_hi(x.z);
      acc[6] += bf16_lo(x.w); acc[7] += bf16_hi(x.w);
    }
    uint4 o;
    o.x = pack_bf16(acc[0] * scale, acc[1] * scale);
    o.y = pack_bf16(acc[2] * scale, acc[3] * scale);
    o.z = pack_bf16(acc[4] * scale, acc[5] * scale);
    o.w = pack_bf16(acc[6] * scale, acc[7] * scale);
    reinterpret_cast<uint4*>(out)[v] = o;
  }
  for (size_t i = nvec * 8 + tid; i < n; i += stride) {
    float acc = 0.f;
    for (int k = 0; k < a.n_in; ++k)
      acc += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(a.in[k])[i]);
    out[i] = __float2bfloat16(acc * scale);
  }
}

// ----------------------------------------------------------- cast/scale -----
// out(bf16) = in(fp32) * scale  — the "bucket cast/scale" half of the Horovod path when the
// model keeps fp32 grads; bf16 models skip it (grads are produced in the wire dtype).
__global__ void __launch_bounds__(256) kcast_scale_f32_bf16(const float* __restrict__ in,
                                                             __nv_bfloat16* __restrict__ out, size_t n,
                                                             float scale) {
  size_t nvec = n / 8;
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = tid; v < nvec; v += stride) {
    float4 a = reinterpret_cast<const float4*>(in)[2 * v];
    float4 b = reinterpret_cast<const float4*>(in)[2 * v + 1];
    uint4 o;
    o.x = pack_bf16(a.x * scale, a.y * scale);
    o.y = pack_bf16(a.z * scale, a.w * scale);
    o.z = pack_bf16(b.x * scale, b.y * scale);
    o.w = pack_bf16(b.z * scale, b.w * scale);
    reinterpret_cast<uint4*>(out)[v] = o;
  }
  for (size_t i = nvec * 8 + tid; i < n; i += stride) out[i] = __float2bfloat16(in[i] * scale);
}

__global__ void __launch_bounds__(256) kcast_scale_bf16_f32(const __nv_bfloat16* __restrict__ in,
                                                             float* __restrict__ out, size_t n,
                                                             float scale) {
  size_t nvec = n / 8;
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = tid; v < nvec; v += stride) {
    uint4 x = reinterpret_cast<const uint4*>(in)[v];
    float4 a = make_float4(bf16_lo(x.x) * scale, bf16_hi(x.x) * scale, bf16_lo(x.y) * scale,
                           bf16_hi(x.y) * scale);
    float4 b = make_float4(bf16_lo(x.z) * scale, bf16_hi(x.z) * scale, bf16_lo(x.w) * scale,
                           bf16_hi(x.w) * scale);
    reinterpret_cast<float4*>(out)[2 * v] = a;
    reinterpret_cast<float4*>(out)[2 * v + 1] = b;
  }
  for (size_t i = nvec * 8 + tid; i < n; i += stride) out[i] = __bfloat162float(in[i]) * scale;
}

// ------------------------------------------------------------ optimizers ----
// One launch updates a whole flat parameter buffer.  Multi-precision layout
// (reference MP_SGDMomKernel, optimizer_op-inl.h:403-427): fp32 master weight +
// fp32 state in HBM, bf16 gradient in, bf16 model weight out.
// The per-element math for all 13 optimizer kinds is in optimizer.cuh (shared with the fused PS kernels).
// G = gradient storage type, W = model-weight storage type (master is always fp32)
template <typename G, typename W>
__global__ void __launch_bounds__(256) kopt(OptArgs o, const G* __restrict__ grad,
                                            float* __restrict__ master, float* __restrict__ s1,
                                            float* __restrict__ s2, W* __restrict__ wout, size_t n) {
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t nvec = n / 4;
  for (size_t v = tid; v < nvec; v += stride) {
    float g[4];
    if constexpr (sizeof(G) == 2) {
      uint2 x = reinterpret_cast<const uint2*>(grad)[v];
      g[0] = bf16_lo(x.x); g[1] = bf16_hi(x.x); g[2] = bf16_lo(x.y); g[3] = bf16_hi(x.y);
    } else {
      float4 x = reinterpret_cast<const float4*>(grad)[v];
      g[0] = x.x; g[1] = x.y; g[2] = x.z; g[3] = x.w;
    }
    float4 w4 = reinterpret_cast<float4*>(master)[v];
    float w[4] = {w4.x, w4.y, w4.z, w4.w};
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    if (opt_uses_s1(o.kind)) {
      float4 t = reinterpret_cast<float4*>(s1)[v];
      a[0] = t.x; a[1] = t.y; a[2] = t.z; a[3] = t.w;
    }
    if (opt_uses_s2(o.kind)) {
      float4 t = reinterpret_cast<float4*>(s2)[v];
      b[0] = t.x; b[1] = t.y; b[2] = t.z; b[3] = t.w;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) opt_apply(o, g[k], w[k], a[k], b[k]);
    reinterpret_cast<float4*>(master)[v] = make_float4(w[0], w[1], w[2], w[3]);
    if (opt_uses_s1(o.kind))
      reinterpret_cast<float4*>(s1)[v] = make_float4(a[0], a[1], a[2], a[3]);
    if (opt_uses_s2(o.kind)) reinterpret_cast<float4*>(s2)[v] = make_float4(b[0], b[1], b[2], b[3]);
    if (wout != nullptr) {
      if constexpr (sizeof(W) == 2) {
        uint2 ov;
        ov.x = pack_bf16(w[0], w[1]);
        ov.y = pack_bf16(w[2], w[3]);
        reinterpret_cast<uint2*>(wout)[v] = ov;
      } else {
        reinterpret_cast<float4*>(wout)[v] = make_float4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  for (size_t i = nvec * 4 + tid; i < n; i += stride) {
    float g;
    if constexpr (sizeof(G) == 2) g = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(grad)[i]);
    else g = grad[i];
    float w = master[i];
    float a = opt_uses_s1(o.kind) ? s1[i] : 0.f;
    float b = opt_uses_s2(o.kind) ? s2[i] : 0.f;
    opt_apply(o, g, w, a, b);
    master[i] = w;
    if (opt_uses_s1(o.kind)) s1[i] = a;
    if (opt_uses_s2(o.kind)) s2[i] = b;
    if (wout != nullptr) {
      if constexpr (sizeof(W) == 2) reinterpret_cast<__nv_bfloat16*>(wout)[i] = __float2bfloat16(w);
      else wout[i] = w;
    }
  }
}

// ------------------------------------------------------- input pipeline -----
// uint8 NHWC -> bf16 NHWC with per-channel (x/255 - mean)/std, C == 3.
__global__ void __launch_bounds__(256) knorm_u8_bf16(const uint8_t* __restrict__ in,
                                                      __nv_bfloat16* __restrict__ out, size_t n,
                                                      float m0, float m1, float m2, float s0, float s1,
                                                      float s2) {
  // 48 bytes = 16 pixels per thread-iteration keeps channel phase fixed (48 % 3 == 0)
  size_t nchunk = n / 48;
  size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  const float mean[3] = {m0, m1, m2};
  const float inv[3] = {1.f / (255.f * s0), 1.f / (255.f * s1), 1.f / (255.f * s2)};
  const float off[3] = {m0 / s0, m1 / s1, m2 / s2};
  (void)mean;
  for (size_t c = tid; c < nchunk; c += stride) {
    const uint4* src = reinterpret_cast<const uint4*>(in + c * 48);
    uint4 q[3] = {src[0], src[1], src[2]};
    const uint8_t* b = reinterpret_cast<const uint8_t*>(q);
    uint32_t o[24];
#pragma unroll
    for (int k = 0; k < 24; ++k) {
      int c0 = (2 * k) % 3, c1 = (2 * k + 1) % 3;
      o[k] = pack_bf16(b[2 * k] * inv[c0] - off[c0], b[2 * k + 1] * inv[c1] - off[c1]);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + c * 48);
#pragma unroll
    for (int k = 0; k < 6; ++k) dst[k] = make_uint4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
  }
  for (size_t i = nchunk * 48 + tid; i < n; i += stride) {
    int ch = i % 3;
    out[i] = __float2bfloat16(in[i] * inv[ch] - off[ch]);
  }
}

}  // namespace dlcfn

namespace dlcfn {
// out[i] (fp32) = sum_n x[n * inner + i] for a bf16 [N, inner] tensor, 8 elements per thread, 8 rows of loads in flight.
// The fused ResNet stem's backward sums the output gradient over the batch with it (models/resnet.py:_FusedStemFn); torch's
// strided reduce_kernel took 3.5 ms for the same 411 MB read (profiles/r2).
__global__ void __launch_bounds__(256) kbatch_sum_bf16(const uint4* __restrict__ x, float* __restrict__ out, int N, long long inner8) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < inner8; i += (long long)gridDim.x * blockDim.x) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = 0.f;
    int n = 0;
    for (; n + 8 <= N; n += 8) {
      uint4 v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = __ldg(x + (long long)(n + j) * inner8 + i);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[0] += bf16_lo(v[j].x); a[1] += bf16_hi(v[j].x); a[2] += bf16_lo(v[j].y); a[3] += bf16_hi(v[j].y);
        a[4] += bf16_lo(v[j].z); a[5] += bf16_hi(v[j].z); a[6] += bf16_lo(v[j].w); a[7] += bf16_hi(v[j].w);
      }
    }
    for (; n < N; ++n) {
      const uint4 v = __ldg(x + (long long)n * inner8 + i);
      a[0] += bf16_lo(v.x); a[1] += bf16_hi(v.x); a[2] += bf16_lo(v.y); a[3] += bf16_hi(v.y);
      a[4] += bf16_lo(v.z); a[5] += bf16_hi(v.z); a[6] += bf16_lo(v.w); a[7] += bf16_hi(v.w);
    }
    float4* o = reinterpret_cast<float4*>(out + i * 8);
    o[0] = make_float4(a[0], a[1], a[2], a[3]);
    o[1] = make_float4(a[4], a[5], a[6], a[7]);
  }
}

// Column sums of a bf16 [M, N] matrix (bias gradient db = sum_rows dY of the tcgen05 Linear): block = 32 column vectors
// (256 columns) x 8 row lanes, grid = (column strips, row slices); per-slice partials [slices][N] then one fixed-order sum -
// deterministic.  torch's reduce_kernel took 2.07 ms per BERT-large step for the same 3.6 GB (profiles/r2/profile_bert.txt).
__global__ void __launch_bounds__(256) kcolsum_partial(const uint4* __restrict__ x, float* __restrict__ part, long long M, int N8,
                                                        long long ld8) {
  __shared__ float sh[8][32][9];
  const int cx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col8 = blockIdx.x * 32 + cx;
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.f;
  if (col8 < N8) {
    long long r = (long long)blockIdx.y * 8 + ty;
    const long long step = (long long)gridDim.y * 8;
    for (; r + 3 * step < M; r += 4 * step) {
      const uint4 v0 = __ldg(x + r * ld8 + col8), v1 = __ldg(x + (r + step) * ld8 + col8);
      const uint4 v2 = __ldg(x + (r + 2 * step) * ld8 + col8), v3 = __ldg(x + (r + 3 * step) * ld8 + col8);
      const uint4 vs[4] = {v0, v1, v2, v3};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a[0] += bf16_lo(vs[j].x); a[1] += bf16_hi(vs[j].x); a[2] += bf16_lo(vs[j].y); a[3] += bf16_hi(vs[j].y);
        a[4] += bf16_lo(vs[j].z); a[5] += bf16_hi(vs[j].z); a[6] += bf16_lo(vs[j].w); a[7] += bf16_hi(vs[j].w);
      }
    }
    for (; r < M; r += step) {
      const uint4 v = __ldg(x + r * ld8 + col8);
      a[0] += bf16_lo(v.x); a[1] += bf16_hi(v.x); a[2] += bf16_lo(v.y); a[3] += bf16_hi(v.y);
      a[4] += bf16_lo(v.z); a[5] += bf16_hi(v.z); a[6] += bf16_lo(v.w); a[7] += bf16_hi(v.w);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) sh[ty][cx][k] = a[k];
  __syncthreads();
  if (ty == 0 && col8 < N8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = 0.f;
#pragma unroll
      for (int y = 0; y < 8; ++y) t += sh[y][cx][k];
      part[(size_t)blockIdx.y * N8 * 8 + (size_t)col8 * 8 + k] = t;
    }
  }
}
__global__ void __launch_bounds__(256) kcolsum_final(const float* __restrict__ part, float* __restrict__ out, int N, int slices) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float t = 0.f;
  for (int s = 0; s < slices; ++s) t += part[(size_t)s * N + c];
  out[c] = t;
}
}  // namespace dlcfn

using namespace dlcfn;

// out[N] (fp32) = column sums of bf16 x[M, N] (row stride ld elements; N, ld multiples of 8).  workspace: slices * N floats
// with slices = dlcfn_colsum_slices(M).
DLCFN_API int dlcfn_colsum_slices(long long M) {
  long long s = (M + 63) / 64;
  return (int)(s < 1 ? 1 : (s > 64 ? 64 : s));
}
DLCFN_API int dlcfn_colsum_bf16(const void* x, float* out, long long M, int N, long long ld, float* workspace, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (M <= 0 || N <= 0 || (N & 7) || (ld & 7) || (((uintptr_t)x) & 15)) return (int)cudaErrorInvalidValue;
  const int slices = dlcfn_colsum_slices(M);
  dim3 grid((N / 8 + 31) / 32, slices);
  kcolsum_partial<<<grid, 256, 0, stream>>>((const uint4*)x, workspace, M, N / 8, ld / 8);
  kcolsum_final<<<(N + 255) / 256, 256, 0, stream>>>(workspace, out, N, slices);
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_batch_sum_bf16(const void* x, float* out, int N, long long inner, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (N <= 0 || inner <= 0 || (inner & 7) || (((uintptr_t)x) & 15) || (((uintptr_t)out) & 15)) return (int)cudaErrorInvalidValue;
  const long long inner8 = inner / 8;
  long long blocks = (inner8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  kbatch_sum_bf16<<<(int)blocks, 256, 0, stream>>>((const uint4*)x, out, N, inner8);
  return (int)cudaGetLastError();
}

// dtype codes shared with python: 0 = fp32, 1 = bf16
DLCFN_API int dlcfn_elementwise_sum(const void* const* inputs, int n_in, void* out, size_t n, int dtype,
                                    float scale, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (n_in < 1 || n_in > DLCFN_MAX_RANKS * 2) return (int)cudaErrorInvalidValue;
  if (n == 0) return 0;
  SumArgs a;
  a.n_in = n_in;
  for (int i = 0; i < n_in; ++i) a.in[i] = inputs[i];
  if (dtype == 0) {
    ksum<float><<<grid_for(n / 4 + 1, 256), 256, 0, stream>>>(a, (float*)out, n, scale);
  } else if (dtype == 1) {
    ksum<__nv_bfloat16><<<grid_for(n / 8 + 1, 256), 256, 0, stream>>>(a, (__nv_bfloat16*)out, n, scale);
  } else {
    return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_cast_scale(const void* in, void* out, size_t n, int in_dtype, int out_dtype, float scale,
                               cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (n == 0) return 0;
  int g = grid_for(n / 8 + 1, 256);
  if (in_dtype == 0 && out_dtype == 1)
    kcast_scale_f32_bf16<<<g, 256, 0, stream>>>((const float*)in, (__nv_bfloat16*)out, n, scale);
  else if (in_dtype == 1 && out_dtype == 0)
    kcast_scale_bf16_f32<<<g, 256, 0, stream>>>((const __nv_bfloat16*)in, (float*)out, n, scale);
  else
    return (int)cudaErrorInvalidValue;
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_fused_optimizer(int kind, const void* grad, int grad_dtype, float* master, float* s1,
                                    float* s2, void* wout, int w_dtype, size_t n, float lr, float wd,
                                    float momentum, float beta1, float beta2, float eps, float rescale,
                                    float clip, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (n == 0) return 0;
  OptArgs o{kind, lr, wd, momentum, beta1, beta2, eps, rescale, clip};
  int g = grid_for(n / 4 + 1, 256);
  if (grad_dtype == 1 && w_dtype == 1)
    kopt<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, stream>>>(o, (const __nv_bfloat16*)grad, master, s1, s2,
                                                               (__nv_bfloat16*)wout, n);
  else if (grad_dtype == 0 && w_dtype == 0)
    kopt<float, float><<<g, 256, 0, stream>>>(o, (const float*)grad, master, s1, s2, (float*)wout, n);
  else if (grad_dtype == 0 && w_dtype == 1)
    kopt<float, __nv_bfloat16><<<g, 256, 0, stream>>>(o, (const float*)grad, master, s1, s2,
                                                       (__nv_bfloat16*)wout, n);
  else if (grad_dtype == 1 && w_dtype == 0)
    kopt<__nv_bfloat16, float><<<g, 256, 0, stream>>>(o, (const __nv_bfloat16*)grad, master, s1, s2,
                                                       (float*)wout, n);
  else
    return (int)cudaErrorInvalidValue;
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_normalize_u8_nhwc(const uint8_t* in, void* out_bf16, size_t n, const float* mean3,
                                      const float* std3, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (n == 0) return 0;
  knorm_u8_bf16<<<grid_for(n / 48 + 1, 256), 256, 0, stream>>>(in, (__nv_bfloat16*)out_bf16, n, mean3[0],
                                                               mean3[1], mean3[2], std3[0], std3[1],
                                                               std3[2]);
  return (int)cudaGetLastError();
}
