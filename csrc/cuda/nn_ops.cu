// Model-compute kernels off the GEMM path (sm_100a): softmax cross-entropy, local response normalisation,
// depthwise convolution, strided NHWC sub-sampling for 1x1/stride-s convolutions, row-sparse optimizer updates.
//
// Reference inventory (SURVEY 2.9 d/c):
//   M7  SoftmaxOutput: mshadow Softmax kernels, one block per row (tensor_gpu-inl.cuh:558-571, softmax-inl.h:166-266)
//   M9  LRN: cuDNN LRN (nn/lrn.cu)
//   M10 depthwise convolution: TF-derived hand kernels (nn/depthwise_convolution_tf.cuh:76,197,329)
//   O3  row-sparse optimizer updates SGDMomStdDnsRspDnsKernel / AdamStdDnsRspDnsKernel / _sparse_adagrad_update
//       (optimizer_op.cu:33,128,262) and contrib group_adagrad (contrib/optimizer_op.cu)
//
// B200 shape: all of these are bandwidth-bound, so every kernel streams NHWC / row-major data with 16-byte
// accesses where the layout allows and touches each element the minimum number of times:
//   * softmax-CE forward = one pass that keeps an online (max, sum-exp) pair per row and emits loss + logsumexp;
//     backward re-reads the logits once and writes (softmax - onehot) * row_scale: 2 reads + 1 write in total;
//   * LRN stages a [pixels, C] tile in shared memory once and does the window sums there (fwd: 1 read 1 write);
//   * depthwise conv owns 8 channels per thread, weights kept as [kh, kw, C] fp32;
//   * sparse updates apply optimizer.cuh:opt_apply on the touched rows only (the reference's lazy update).
#include "common.cuh"
#include "optimizer.cuh"

namespace dlcfn {

__device__ __forceinline__ void unpack8n(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8n(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// ----------------------------------------------------------------------------------------------- softmax-CE
// G threads cooperate on one row (G = 32: one warp per row, 8 rows per CTA; G = 256: one CTA per row).
template <int G>
__device__ __forceinline__ void group_reduce_max_sum(float& m, float& s, float* smem) {
  // combine online-softmax pairs: (m, s) (+) (m', s') = (max, s*exp(m-max) + s'*exp(m'-max))
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o);
    const float s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mm = fmaxf(m, m2);
    s = (mm == -INFINITY) ? 0.f : s * __expf(m - mm) + s2 * __expf(m2 - mm);
    m = mm;
  }
  if constexpr (G > 32) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { smem[2 * warp] = m; smem[2 * warp + 1] = s; }
    __syncthreads();
    float M = -INFINITY, S = 0.f;
#pragma unroll
    for (int w = 0; w < G / 32; ++w) {
      const float m2 = smem[2 * w], s2 = smem[2 * w + 1];
      const float mm = fmaxf(M, m2);
      S = (mm == -INFINITY) ? 0.f : S * __expf(M - mm) + s2 * __expf(m2 - mm);
      M = mm;
    }
    m = M; s = S;
    __syncthreads();
  }
}

template <typename T, int G>
__global__ void __launch_bounds__(256) ksoftmax_ce_fwd(const T* __restrict__ x, const int64_t* __restrict__ label,
                                                        float* __restrict__ loss, float* __restrict__ lse,
                                                        long long M, int C, long long ldx, long long ignore_index) {
  __shared__ float smem[16];
  const int rows_per_cta = 256 / G;
  const int sub = threadIdx.x / G, lane = threadIdx.x % G;
  for (long long row0 = (long long)blockIdx.x * rows_per_cta; row0 < M; row0 += (long long)gridDim.x * rows_per_cta) {
    const long long row = row0 + sub;
    float m = -INFINITY, s = 0.f;
    if (row < M) {
      const T* xr = x + row * ldx;
      for (int c = lane; c < C; c += G) {
        const float v = to_f<T>(xr[c]);
        if (v > m) { s = s * __expf(m - v) + 1.f; m = v; }
        else s += __expf(v - m);
      }
    }
    group_reduce_max_sum<G>(m, s, smem);
    if (row < M && lane == 0) {
      const float l = m + __logf(s);
      lse[row] = l;
      const long long y = label[row];
      loss[row] = (y == ignore_index || y < 0 || y >= C) ? 0.f : l - to_f<T>(x[row * ldx + y]);
    }
  }
}

// dx[row, c] = (exp(x - lse) - [c == label]) * row_scale[row]   (row_scale = grad_out / #valid rows, 0 if ignored)
// dx rows are `wdx` wide (>= C): columns C..wdx-1 are padding of the producing GEMM and get exact zeros.
template <typename T>
__global__ void __launch_bounds__(256) ksoftmax_ce_bwd(const T* __restrict__ x, const int64_t* __restrict__ label,
                                                        const float* __restrict__ lse,
                                                        const float* __restrict__ row_scale, T* __restrict__ dx,
                                                        long long M, int C, int wdx, long long ldx, long long lddx) {
  const long long total = M * (long long)wdx;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / wdx;
    const int c = (int)(t - row * wdx);
    const float sc = row_scale[row];
    float g = 0.f;
    if (sc != 0.f && c < C) {
      g = __expf(to_f<T>(x[row * ldx + c]) - lse[row]);
      if ((long long)c == label[row]) g -= 1.f;
      g *= sc;
    }
    dx[row * lddx + c] = from_f<T>(g);
  }
}

// ------------------------------------------------------------------------------------------------------ LRN
// y_c = x_c * s_c^-beta, s_c = k + alpha * sum_{j=c-lo}^{c+hi} x_j^2   (alpha already divided by n when the
// caller wants torch / cuDNN semantics; TF's depth_radius form passes alpha as is).
struct LRNArgs {
  long long npix;
  int C, lo, hi, pix_per_cta;
  float alpha, beta, k;
};

__global__ void __launch_bounds__(256) klrn_fwd(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                 float* __restrict__ scale_out, LRNArgs a) {
  extern __shared__ float sm[];  // x^2 tile [pix_per_cta, C]
  const int tile = a.pix_per_cta * a.C;
  for (long long p0 = (long long)blockIdx.x * a.pix_per_cta; p0 < a.npix; p0 += (long long)gridDim.x * a.pix_per_cta) {
    const long long base = p0 * a.C;
    const long long lim = min((long long)tile, (a.npix - p0) * a.C);
    for (int e = threadIdx.x; e < lim; e += blockDim.x) {
      const float v = __bfloat162float(x[base + e]);
      sm[e] = v * v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < lim; e += blockDim.x) {
      const int c = e % a.C, row = e - c;
      float acc = 0.f;
      const int j0 = max(0, c - a.lo), j1 = min(a.C - 1, c + a.hi);
      for (int j = j0; j <= j1; ++j) acc += sm[row + j];
      const float s = a.k + a.alpha * acc;
      if (scale_out) scale_out[base + e] = s;
      y[base + e] = __float2bfloat16(__bfloat162float(x[base + e]) * __powf(s, -a.beta));
    }
    __syncthreads();
  }
}

// dx_c = dy_c s_c^-beta - 2 alpha beta x_c * sum_{j : c in window(j)} dy_j x_j s_j^(-beta-1)
__global__ void __launch_bounds__(256) klrn_bwd(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                 const float* __restrict__ scale, __nv_bfloat16* __restrict__ dx,
                                                 LRNArgs a) {
  extern __shared__ float sm[];  // t_j = dy_j x_j s_j^(-beta-1)
  const int tile = a.pix_per_cta * a.C;
  for (long long p0 = (long long)blockIdx.x * a.pix_per_cta; p0 < a.npix; p0 += (long long)gridDim.x * a.pix_per_cta) {
    const long long base = p0 * a.C;
    const long long lim = min((long long)tile, (a.npix - p0) * a.C);
    for (int e = threadIdx.x; e < lim; e += blockDim.x) {
      const float s = scale[base + e];
      sm[e] = __bfloat162float(dy[base + e]) * __bfloat162float(x[base + e]) * __powf(s, -a.beta - 1.f);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < lim; e += blockDim.x) {
      const int c = e % a.C, row = e - c;
      // channels j whose window [j-lo, j+hi] contains c: j in [c-hi, c+lo]
      float acc = 0.f;
      const int j0 = max(0, c - a.hi), j1 = min(a.C - 1, c + a.lo);
      for (int j = j0; j <= j1; ++j) acc += sm[row + j];
      const float s = scale[base + e];
      const float g = __bfloat162float(dy[base + e]) * __powf(s, -a.beta) -
                      2.f * a.alpha * a.beta * __bfloat162float(x[base + e]) * acc;
      dx[base + e] = __float2bfloat16(g);
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------- depthwise convolution
struct DWGeom {
  int N, H, W, CV, OH, OW, K, S, P;
};

// x [N,H,W,C] bf16, w [K,K,C] fp32, y [N,OH,OW,C] bf16; one thread = 8 channels of one output pixel
__global__ void __launch_bounds__(256) kdw_fwd(const uint4* __restrict__ x, const float* __restrict__ w,
                                                uint4* __restrict__ y, DWGeom g) {
  const long long total = (long long)g.N * g.OH * g.OW * g.CV;
  const int C = g.CV * 8;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % g.CV);
    long long r = t / g.CV;
    const int ow = (int)(r % g.OW); r /= g.OW;
    const int oh = (int)(r % g.OH);
    const int n = (int)(r / g.OH);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int kh = 0; kh < g.K; ++kh) {
      const int ih = oh * g.S - g.P + kh;
      if (ih < 0 || ih >= g.H) continue;
      for (int kw = 0; kw < g.K; ++kw) {
        const int iw = ow * g.S - g.P + kw;
        if (iw < 0 || iw >= g.W) continue;
        float xv[8];
        unpack8n(x[(((long long)n * g.H + ih) * g.W + iw) * g.CV + cv], xv);
        const float4* wp = reinterpret_cast<const float4*>(w + ((long long)kh * g.K + kw) * C + cv * 8);
        const float4 w0 = wp[0], w1 = wp[1];
        acc[0] += xv[0] * w0.x; acc[1] += xv[1] * w0.y; acc[2] += xv[2] * w0.z; acc[3] += xv[3] * w0.w;
        acc[4] += xv[4] * w1.x; acc[5] += xv[5] * w1.y; acc[6] += xv[6] * w1.z; acc[7] += xv[7] * w1.w;
      }
    }
    y[t] = pack8n(acc);
  }
}

// dx[n,h,w,c] = sum_{kh,kw} dy[n,(h+P-kh)/S,(w+P-kw)/S,c] * w[kh,kw,c]  (where the division is exact)
__global__ void __launch_bounds__(256) kdw_bwd_data(const uint4* __restrict__ dy, const float* __restrict__ w,
                                                     uint4* __restrict__ dx, DWGeom g) {
  const long long total = (long long)g.N * g.H * g.W * g.CV;
  const int C = g.CV * 8;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % g.CV);
    long long r = t / g.CV;
    const int iw = (int)(r % g.W); r /= g.W;
    const int ih = (int)(r % g.H);
    const int n = (int)(r / g.H);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int kh = 0; kh < g.K; ++kh) {
      const int a = ih + g.P - kh;
      if (a < 0 || a % g.S) continue;
      const int oh = a / g.S;
      if (oh >= g.OH) continue;
      for (int kw = 0; kw < g.K; ++kw) {
        const int b = iw + g.P - kw;
        if (b < 0 || b % g.S) continue;
        const int ow = b / g.S;
        if (ow >= g.OW) continue;
        float dv[8];
        unpack8n(dy[(((long long)n * g.OH + oh) * g.OW + ow) * g.CV + cv], dv);
        const float4* wp = reinterpret_cast<const float4*>(w + ((long long)kh * g.K + kw) * C + cv * 8);
        const float4 w0 = wp[0], w1 = wp[1];
        acc[0] += dv[0] * w0.x; acc[1] += dv[1] * w0.y; acc[2] += dv[2] * w0.z; acc[3] += dv[3] * w0.w;
        acc[4] += dv[4] * w1.x; acc[5] += dv[5] * w1.y; acc[6] += dv[6] * w1.z; acc[7] += dv[7] * w1.w;
      }
    }
    dx[t] = pack8n(acc);
  }
}

// partial[b, tap, c] = sum over the pixels of slice b of x[n, oh*S-P+kh, ow*S-P+kw, c] * dy[n,oh,ow,c].
// grid (B, K*K); block 256 threads = (256/CVB pixel lanes) x CVB channel vectors, looped over CV.
__global__ void __launch_bounds__(256) kdw_bwd_filter(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                       float* __restrict__ partial, DWGeom g) {
  __shared__ float red[256 * 8];
  const int tap = blockIdx.y, kh = tap / g.K, kw = tap % g.K;
  const int C = g.CV * 8;
  const long long npix = (long long)g.N * g.OH * g.OW;
  const long long per = (npix + gridDim.x - 1) / gridDim.x;
  const long long p_lo = blockIdx.x * per, p_hi = min(npix, p_lo + per);
  const int cvb = min(g.CV, 32);            // channel vectors handled concurrently
  const int lanes = 256 / cvb;              // pixel lanes
  const int my_cv = threadIdx.x % cvb, my_lane = threadIdx.x / cvb;
  for (int cv0 = 0; cv0 < g.CV; cv0 += cvb) {
    const int cv = cv0 + my_cv;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (cv < g.CV && my_lane < lanes) {
      for (long long p = p_lo + my_lane; p < p_hi; p += lanes) {
        const int ow = (int)(p % g.OW);
        const long long q = p / g.OW;
        const int oh = (int)(q % g.OH);
        const int n = (int)(q / g.OH);
        const int ih = oh * g.S - g.P + kh, iw = ow * g.S - g.P + kw;
        if (ih < 0 || ih >= g.H || iw < 0 || iw >= g.W) continue;
        float xv[8], dv[8];
        unpack8n(x[(((long long)n * g.H + ih) * g.W + iw) * g.CV + cv], xv);
        unpack8n(dy[p * g.CV + cv], dv);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += xv[k] * dv[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) red[threadIdx.x * 8 + k] = acc[k];
    __syncthreads();
    // thread (lane 0, cv) sums the pixel lanes in a fixed order: deterministic
    if (my_lane == 0 && cv < g.CV) {
      float out[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int l = 0; l < lanes; ++l)
#pragma unroll
        for (int k = 0; k < 8; ++k) out[k] += red[(l * cvb + my_cv) * 8 + k];
      float* dst = partial + ((long long)blockIdx.x * g.K * g.K + tap) * C + cv * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) dst[k] = out[k];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) kdw_filter_finalize(const float* __restrict__ partial, float* __restrict__ dw,
                                                            int B, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += partial[(long long)b * n + i];
    dw[i] = s;
  }
}

// ------------------------------------------------------------------------------- strided sub-sampling (NHWC)
// y[n,oh,ow,:] = x[n,oh*S,ow*S,:] - the gather in front of a 1x1 / stride-S convolution evaluated as a GEMM
__global__ void __launch_bounds__(256) ksubsample_fwd(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H,
                                                       int W, int CV, int OH, int OW, int S) {
  const long long total = (long long)N * OH * OW * CV;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    long long r = t / CV;
    const int ow = (int)(r % OW); r /= OW;
    const int oh = (int)(r % OH);
    const int n = (int)(r / OH);
    y[t] = x[(((long long)n * H + oh * S) * W + ow * S) * CV + cv];
  }
}

// dx[n,h,w,:] = dy[n,h/S,w/S,:] if h % S == 0 and w % S == 0 (and inside OH x OW) else 0
__global__ void __launch_bounds__(256) ksubsample_bwd(const uint4* __restrict__ dy, uint4* __restrict__ dx, int N, int H,
                                                       int W, int CV, int OH, int OW, int S) {
  const long long total = (long long)N * H * W * CV;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % CV);
    long long r = t / CV;
    const int w = (int)(r % W); r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (h % S == 0 && w % S == 0 && h / S < OH && w / S < OW)
      v = dy[(((long long)n * OH + h / S) * OW + w / S) * CV + cv];
    dx[t] = v;
  }
}

// ----------------------------------------------------------------------------------- row-sparse optimizer
// Applies optimizer.cuh:opt_apply to the rows listed in `rows` only (unique ids; the reference's lazy update).
template <typename G, typename W>
__global__ void __launch_bounds__(256) ksparse_opt(OptArgs o, const int64_t* __restrict__ rows, long long nrows_touched,
                                                    long long num_rows, int row_len, const G* __restrict__ grad,
                                                    float* __restrict__ master, float* __restrict__ s1,
                                                    float* __restrict__ s2, W* __restrict__ wout) {
  const long long total = nrows_touched * row_len;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long r = t / row_len;
    const int c = (int)(t - r * row_len);
    const long long row = rows[r];
    if (row < 0 || row >= num_rows) continue;
    const long long i = row * row_len + c;
    float w = master[i];
    float a = opt_uses_s1(o.kind) ? s1[i] : 0.f;
    float b = opt_uses_s2(o.kind) ? s2[i] : 0.f;
    opt_apply(o, to_f<G>(grad[t]), w, a, b);
    master[i] = w;
    if (opt_uses_s1(o.kind)) s1[i] = a;
    if (opt_uses_s2(o.kind)) s2[i] = b;
    if (wout != nullptr) wout[i] = from_f<W>(w);
  }
}

// contrib group_adagrad: one history value per row, hist += mean(g_row^2); w_row -= lr * g_row / sqrt(hist + eps).
// One warp per touched row.
template <typename G>
__global__ void __launch_bounds__(256) kgroup_adagrad(const int64_t* __restrict__ rows, long long nrows_touched,
                                                       long long num_rows, int row_len, const G* __restrict__ grad,
                                                       float* __restrict__ weight, float* __restrict__ hist, float lr,
                                                       float eps, float rescale, float clip) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < nrows_touched; r += nwarps) {
    const long long row = rows ? rows[r] : r;
    if (row < 0 || row >= num_rows) continue;
    float ss = 0.f;
    for (int c = lane; c < row_len; c += 32) {
      float g = to_f<G>(grad[r * row_len + c]) * rescale;
      if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
      ss += g * g;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float h = hist[row] + ss / (float)row_len;
    const float inv = lr / sqrtf(h + eps);
    for (int c = lane; c < row_len; c += 32) {
      float g = to_f<G>(grad[r * row_len + c]) * rescale;
      if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
      weight[row * row_len + c] -= inv * g;
    }
    __syncwarp();
    if (lane == 0) hist[row] = h;
  }
}

}  // namespace dlcfn

using namespace dlcfn;

static inline int nn_grid(long long total, int per_block = 256) {
  long long b = (total + per_block - 1) / per_block;
  if (b > 148LL * 16) b = 148LL * 16;
  return b < 1 ? 1 : (int)b;
}

// dtype: 0 fp32, 1 bf16.  loss / lse: fp32 [M].
DLCFN_API int dlcfn_softmax_ce_fwd(const void* x, int dtype, const int64_t* label, float* loss, float* lse, long long M,
                                   int C, long long ldx, long long ignore_index, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (M == 0) return 0;
  if (C < 1) return (int)cudaErrorInvalidValue;
  if (C <= 2048) {
    const int grid = nn_grid(M, 8);
    if (dtype == 0) ksoftmax_ce_fwd<float, 32><<<grid, 256, 0, stream>>>((const float*)x, label, loss, lse, M, C, ldx, ignore_index);
    else if (dtype == 1)
      ksoftmax_ce_fwd<__nv_bfloat16, 32><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, label, loss, lse, M, C, ldx, ignore_index);
    else return (int)cudaErrorInvalidValue;
  } else {
    const int grid = nn_grid(M, 1);
    if (dtype == 0) ksoftmax_ce_fwd<float, 256><<<grid, 256, 0, stream>>>((const float*)x, label, loss, lse, M, C, ldx, ignore_index);
    else if (dtype == 1)
      ksoftmax_ce_fwd<__nv_bfloat16, 256><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, label, loss, lse, M, C, ldx, ignore_index);
    else return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_softmax_ce_bwd(const void* x, int dtype, const int64_t* label, const float* lse,
                                   const float* row_scale, void* dx, long long M, int C, int wdx, long long ldx,
                                   long long lddx, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (M == 0) return 0;
  if (wdx < C || lddx < wdx) return (int)cudaErrorInvalidValue;
  const int grid = nn_grid(M * (long long)wdx);
  if (dtype == 0)
    ksoftmax_ce_bwd<float><<<grid, 256, 0, stream>>>((const float*)x, label, lse, row_scale, (float*)dx, M, C, wdx, ldx, lddx);
  else if (dtype == 1)
    ksoftmax_ce_bwd<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, label, lse, row_scale,
                                                              (__nv_bfloat16*)dx, M, C, wdx, ldx, lddx);
  else return (int)cudaErrorInvalidValue;
  return (int)cudaGetLastError();
}

static inline int lrn_pix_per_cta(int C) {
  int p = 8192 / C;  // 32 KiB of fp32 tile
  if (p > 64) p = 64;
  return p < 1 ? 1 : p;
}

// x, y bf16 [npix, C]; scale fp32 [npix, C] saved for backward (may be null for inference)
DLCFN_API int dlcfn_lrn_fwd(const void* x, void* y, float* scale, long long npix, int C, int size, float alpha,
                            float beta, float k, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (npix == 0) return 0;
  if (C < 1 || C > 8192 || size < 1) return (int)cudaErrorInvalidValue;
  LRNArgs a{npix, C, size / 2, (size - 1) / 2, lrn_pix_per_cta(C), alpha, beta, k};
  const int grid = nn_grid(npix, a.pix_per_cta);
  klrn_fwd<<<grid, 256, (size_t)a.pix_per_cta * C * 4, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, scale, a);
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_lrn_bwd(const void* x, const void* dy, const float* scale, void* dx, long long npix, int C, int size,
                            float alpha, float beta, float k, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (npix == 0) return 0;
  if (C < 1 || C > 8192 || size < 1) return (int)cudaErrorInvalidValue;
  LRNArgs a{npix, C, size / 2, (size - 1) / 2, lrn_pix_per_cta(C), alpha, beta, k};
  const int grid = nn_grid(npix, a.pix_per_cta);
  klrn_bwd<<<grid, 256, (size_t)a.pix_per_cta * C * 4, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, scale,
                                                                  (__nv_bfloat16*)dx, a);
  return (int)cudaGetLastError();
}

static inline bool dw_geom(DWGeom* g, int N, int H, int W, int C, int K, int S, int P) {
  if (C % 8 || K < 1 || K > 11 || S < 1 || P < 0) return false;
  *g = DWGeom{N, H, W, C / 8, (H + 2 * P - K) / S + 1, (W + 2 * P - K) / S + 1, K, S, P};
  return g->OH > 0 && g->OW > 0;
}

DLCFN_API int dlcfn_depthwise_fwd(const void* x, const float* w_khwc, void* y, int N, int H, int W, int C, int K, int S,
                                  int P, cudaStream_t stream) {
  (void)cudaGetLastError();
  DWGeom g;
  if (!dw_geom(&g, N, H, W, C, K, S, P)) return (int)cudaErrorInvalidValue;
  const long long total = (long long)N * g.OH * g.OW * g.CV;
  if (total == 0) return 0;
  kdw_fwd<<<nn_grid(total), 256, 0, stream>>>((const uint4*)x, w_khwc, (uint4*)y, g);
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_depthwise_bwd_data(const void* dy, const float* w_khwc, void* dx, int N, int H, int W, int C, int K,
                                       int S, int P, cudaStream_t stream) {
  (void)cudaGetLastError();
  DWGeom g;
  if (!dw_geom(&g, N, H, W, C, K, S, P)) return (int)cudaErrorInvalidValue;
  const long long total = (long long)N * H * W * g.CV;
  if (total == 0) return 0;
  kdw_bwd_data<<<nn_grid(total), 256, 0, stream>>>((const uint4*)dy, w_khwc, (uint4*)dx, g);
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_depthwise_filter_blocks(int N, int OH, int OW) {
  long long npix = (long long)N * OH * OW;
  long long b = (npix + 2047) / 2048;
  if (b > 148) b = 148;
  return b < 1 ? 1 : (int)b;
}

// partial: fp32 scratch of dlcfn_depthwise_filter_blocks(...) * K*K*C floats; dw: fp32 [K,K,C]
DLCFN_API int dlcfn_depthwise_bwd_filter(const void* x, const void* dy, float* partial, float* dw, int N, int H, int W,
                                         int C, int K, int S, int P, cudaStream_t stream) {
  (void)cudaGetLastError();
  DWGeom g;
  if (!dw_geom(&g, N, H, W, C, K, S, P)) return (int)cudaErrorInvalidValue;
  const int B = dlcfn_depthwise_filter_blocks(N, g.OH, g.OW);
  kdw_bwd_filter<<<dim3(B, K * K), 256, 0, stream>>>((const uint4*)x, (const uint4*)dy, partial, g);
  DLCFN_CHECK(cudaGetLastError());
  const int n = K * K * C;
  kdw_filter_finalize<<<nn_grid(n), 256, 0, stream>>>(partial, dw, B, n);
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_subsample_nhwc_fwd(const void* x, void* y, int N, int H, int W, int C, int S, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (C % 8 || S < 1) return (int)cudaErrorInvalidValue;
  const int OH = (H - 1) / S + 1, OW = (W - 1) / S + 1;
  const long long total = (long long)N * OH * OW * (C / 8);
  if (total == 0) return 0;
  ksubsample_fwd<<<nn_grid(total), 256, 0, stream>>>((const uint4*)x, (uint4*)y, N, H, W, C / 8, OH, OW, S);
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_subsample_nhwc_bwd(const void* dy, void* dx, int N, int H, int W, int C, int S, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (C % 8 || S < 1) return (int)cudaErrorInvalidValue;
  const int OH = (H - 1) / S + 1, OW = (W - 1) / S + 1;
  const long long total = (long long)N * H * W * (C / 8);
  if (total == 0) return 0;
  ksubsample_bwd<<<nn_grid(total), 256, 0, stream>>>((const uint4*)dy, (uint4*)dx, N, H, W, C / 8, OH, OW, S);
  return (int)cudaGetLastError();
}

// rows: int64 [nrows_touched] unique row ids; grad [nrows_touched, row_len]; master/s1/s2 fp32 [num_rows, row_len];
// wout (optional) model-dtype copy of the weight, same shape.
DLCFN_API int dlcfn_sparse_row_update(int kind, const int64_t* rows, long long nrows_touched, long long num_rows,
                                      int row_len, const void* grad, int grad_dtype, float* master, float* s1, float* s2,
                                      void* wout, int w_dtype, float lr, float wd, float momentum, float beta1,
                                      float beta2, float eps, float rescale, float clip, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (nrows_touched == 0 || row_len == 0) return 0;
  OptArgs o{kind, lr, wd, momentum, beta1, beta2, eps, rescale, clip};
  const int grid = nn_grid(nrows_touched * row_len);
  if (grad_dtype == 0 && (w_dtype == 0 || wout == nullptr))
    ksparse_opt<float, float><<<grid, 256, 0, stream>>>(o, rows, nrows_touched, num_rows, row_len, (const float*)grad,
                                                        master, s1, s2, (float*)wout);
  else if (grad_dtype == 0 && w_dtype == 1)
    ksparse_opt<float, __nv_bfloat16><<<grid, 256, 0, stream>>>(o, rows, nrows_touched, num_rows, row_len,
                                                                (const float*)grad, master, s1, s2, (__nv_bfloat16*)wout);
  else if (grad_dtype == 1 && w_dtype == 1)
    ksparse_opt<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, stream>>>(o, rows, nrows_touched, num_rows, row_len,
                                                                        (const __nv_bfloat16*)grad, master, s1, s2,
                                                                        (__nv_bfloat16*)wout);
  else if (grad_dtype == 1 && (w_dtype == 0 || wout == nullptr))
    ksparse_opt<__nv_bfloat16, float><<<grid, 256, 0, stream>>>(o, rows, nrows_touched, num_rows, row_len,
                                                                (const __nv_bfloat16*)grad, master, s1, s2, (float*)wout);
  else return (int)cudaErrorInvalidValue;
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_group_adagrad_update(const int64_t* rows, long long nrows_touched, long long num_rows, int row_len,
                                         const void* grad, int grad_dtype, float* weight, float* hist, float lr,
                                         float eps, float rescale, float clip, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (nrows_touched == 0 || row_len == 0) return 0;
  const int grid = nn_grid(nrows_touched * 32);
  if (grad_dtype == 0)
    kgroup_adagrad<float><<<grid, 256, 0, stream>>>(rows, nrows_touched, num_rows, row_len, (const float*)grad, weight,
                                                    hist, lr, eps, rescale, clip);
  else if (grad_dtype == 1)
    kgroup_adagrad<__nv_bfloat16><<<grid, 256, 0, stream>>>(rows, nrows_touched, num_rows, row_len,
                                                            (const __nv_bfloat16*)grad, weight, hist, lr, eps, rescale, clip);
  else return (int)cudaErrorInvalidValue;
  return (int)cudaGetLastError();
}
