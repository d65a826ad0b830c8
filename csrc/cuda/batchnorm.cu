// Fused BatchNorm(+ReLU) for NHWC bf16 activations, training mode, forward and backward (sm_100a).
//
// Reference M3/M4: BatchNorm is cuDNN or hand kernels "one block per channel"
// (MX/src/operator/nn/batch_norm.cu:208-572) followed by a separate Activation kernel
// (nn/activation.cu).  On B200 the op is purely HBM-bound, so the design goal is the minimum number of
// passes over the activation tensor, each at streaming bandwidth:
//     forward   stats (1 read)                    + apply/ReLU (1 read, 1 write)      = 3 passes
//     backward  reduce (read x, read dy)          + apply (read x, read dy, write dx) = 5 passes
// The ReLU mask is never stored: backward recomputes  y = gamma*xhat + beta > 0  from x.
//
// Layout: x is [M, C] row-major (M = N*H*W), C % 8 == 0, viewed as 16-byte vectors [M, CV], CV = C/8.
// A block is CV x RL threads (RL row lanes); every thread owns ONE channel vector for its whole life,
// so the per-channel constants stay in registers, and walks rows with a stride of grid*RL; a block reads
// RL*CV*16 contiguous bytes per step (fully coalesced).  Per-channel partial sums go block -> smem ->
// [grid, C] partials in global memory (deterministic, no atomics); a finalize kernel reduces them and
// produces mean / invstd / running stats (forward) or dgamma / dbeta / the dx coefficients (backward).
#include "common.cuh"

namespace dlcfn {

constexpr int kBnThreads = 256;
constexpr int kBnMaxGrid = 148 * 4;

struct BnGeom {
  int CV, RL, threads, grid;
};

static inline BnGeom bn_geom(long long M, int C) {
  BnGeom g;
  g.CV = C / 8;
  g.RL = kBnThreads / g.CV;
  if (g.RL < 1) g.RL = 1;
  g.threads = g.CV * g.RL;  // <= 256 (C <= 2048)
  long long row_groups = (M + (long long)g.RL * 4 - 1) / ((long long)g.RL * 4);
  long long grid = row_groups < kBnMaxGrid ? row_groups : kBnMaxGrid;
  if (C >= 1024 && grid > 148 * 2) grid = 148 * 2;  // keep the partial buffers small for wide layers
  g.grid = grid < 1 ? 1 : (int)grid;
  return g;
}

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// Block-level reduction over the RL row lanes of 2x8 per-thread accumulators, then one partial row per
// block: part[(blockIdx * 2 + k) * C + c].
__device__ __forceinline__ void block_reduce_store(float* a, float* b, int CV, int RL, int C,
                                                   float* __restrict__ part) {
  extern __shared__ float sh[];  // [RL][CV][16]
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  float* mine = sh + ((size_t)ty * CV + tx) * 16;
#pragma unroll
  for (int k = 0; k < 8; ++k) { mine[k] = a[k]; mine[8 + k] = b[k]; }
  __syncthreads();
  if (ty == 0) {
    float ra[8], rb[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { ra[k] = 0.f; rb[k] = 0.f; }
    for (int r = 0; r < RL; ++r) {
      const float* o = sh + ((size_t)r * CV + tx) * 16;
#pragma unroll
      for (int k = 0; k < 8; ++k) { ra[k] += o[k]; rb[k] += o[8 + k]; }
    }
    float* pa = part + ((size_t)blockIdx.x * 2) * C + tx * 8;
    float* pb = pa + C;
    reinterpret_cast<float4*>(pa)[0] = make_float4(ra[0], ra[1], ra[2], ra[3]);
    reinterpret_cast<float4*>(pa)[1] = make_float4(ra[4], ra[5], ra[6], ra[7]);
    reinterpret_cast<float4*>(pb)[0] = make_float4(rb[0], rb[1], rb[2], rb[3]);
    reinterpret_cast<float4*>(pb)[1] = make_float4(rb[4], rb[5], rb[6], rb[7]);
  }
}

// ------------------------------------------------------------ forward -------
// x2 != nullptr: fused residual add - the tensor that gets normalised is z = x + x2, it is written to `z` (bf16) and
// the statistics are taken over the ROUNDED values so that the apply pass, which re-reads z, is consistent.
__global__ void __launch_bounds__(256) kbn_stats(const uint4* __restrict__ x, const uint4* __restrict__ x2,
                                                   uint4* __restrict__ z, long long M, int CV, int RL,
                                                   float* __restrict__ part) {
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  float s[8], q[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s[k] = 0.f; q[k] = 0.f; }
  const long long step = (long long)gridDim.x * RL;
  long long r = (long long)blockIdx.x * RL + ty;
  if (x2 == nullptr) {
    // 4 independent 16-byte loads in flight per thread
    for (; r + 3 * step < M; r += 4 * step) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = ld_stream_v4(x + (r + u * step) * CV + tx);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[8];
        unpack8(v[u], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] = fmaf(f[k], f[k], q[k]); }
      }
    }
    for (; r < M; r += step) {
      float f[8];
      unpack8(ld_stream_v4(x + r * CV + tx), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += f[k]; q[k] = fmaf(f[k], f[k], q[k]); }
    }
  } else {
    for (; r + 3 * step < M; r += 4 * step) {
      uint4 va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        va[u] = ld_stream_v4(x + (r + u * step) * CV + tx);
        vb[u] = ld_stream_v4(x2 + (r + u * step) * CV + tx);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float fa[8], fb[8];
        unpack8(va[u], fa);
        unpack8(vb[u], fb);
#pragma unroll
        for (int k = 0; k < 8; ++k) fa[k] += fb[k];
        const uint4 o = pack8(fa);
        z[(r + u * step) * CV + tx] = o;
        unpack8(o, fa);
#pragma unroll
        for (int k = 0; k < 8; ++k) { s[k] += fa[k]; q[k] = fmaf(fa[k], fa[k], q[k]); }
      }
    }
    for (; r < M; r += step) {
      float fa[8], fb[8];
      unpack8(ld_stream_v4(x + r * CV + tx), fa);
      unpack8(ld_stream_v4(x2 + r * CV + tx), fb);
#pragma unroll
      for (int k = 0; k < 8; ++k) fa[k] += fb[k];
      const uint4 o = pack8(fa);
      z[r * CV + tx] = o;
      unpack8(o, fa);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s[k] += fa[k]; q[k] = fmaf(fa[k], fa[k], q[k]); }
    }
  }
  block_reduce_store(s, q, CV, RL, CV * 8, part);
}

// mean/invstd, running-stat update, and the affine (scale, shift) used by the apply kernel
// Sum the per-block partials of one channel: blockDim = (32 channels, kFinY slices of the block range).  With up to
// 592 partial rows the walk is a chain of dependent L2 round trips, so it is spread over 32 row lanes (19 rows each,
// 4 loads in flight) - the 8-lane version cost ~17 us per launch, 1.7 ms per ResNet-50 step over 102 launches.
constexpr int kFinY = 32;
__device__ __forceinline__ void reduce_partials(const float* __restrict__ part, int nblocks, int C, int c,
                                                float& s, float& q) {
  __shared__ float sh_s[kFinY][33], sh_q[kFinY][33];
  float a = 0.f, b = 0.f;
  if (c < C) {
    int blk = threadIdx.y;
    for (; blk + 3 * kFinY < nblocks; blk += 4 * kFinY) {  // 4 independent row loads in flight
      float a0 = part[((size_t)blk * 2) * C + c], b0 = part[((size_t)blk * 2 + 1) * C + c];
      float a1 = part[((size_t)(blk + kFinY) * 2) * C + c], b1 = part[((size_t)(blk + kFinY) * 2 + 1) * C + c];
      float a2 = part[((size_t)(blk + 2 * kFinY) * 2) * C + c], b2 = part[((size_t)(blk + 2 * kFinY) * 2 + 1) * C + c];
      float a3 = part[((size_t)(blk + 3 * kFinY) * 2) * C + c], b3 = part[((size_t)(blk + 3 * kFinY) * 2 + 1) * C + c];
      a += (a0 + a1) + (a2 + a3);
      b += (b0 + b1) + (b2 + b3);
    }
    for (; blk < nblocks; blk += kFinY) {
      a += part[((size_t)blk * 2) * C + c];
      b += part[((size_t)blk * 2 + 1) * C + c];
    }
  }
  sh_s[threadIdx.y][threadIdx.x] = a;
  sh_q[threadIdx.y][threadIdx.x] = b;
  __syncthreads();
  s = 0.f; q = 0.f;
  if (threadIdx.y == 0) {
#pragma unroll
    for (int y = 0; y < kFinY; ++y) { s += sh_s[y][threadIdx.x]; q += sh_q[y][threadIdx.x]; }
  }
}

__global__ void __launch_bounds__(32 * kFinY) kbn_fwd_finalize(const float* __restrict__ part, int nblocks, int C,
                                 long long M, float eps,
                                 float momentum, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ mean, float* __restrict__ invstd,
                                 float* __restrict__ running_mean, float* __restrict__ running_var,
                                 float* __restrict__ scale_shift) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s, q;
  reduce_partials(part, nblocks, C, c, s, q);
  if (c >= C || threadIdx.y != 0) return;
  const float inv_m = 1.f / (float)M;
  const float mu = s * inv_m;
  float var = fmaf(-mu, mu, q * inv_m);
  var = fmaxf(var, 0.f);
  const float is = rsqrtf(var + eps);
  mean[c] = mu;
  invstd[c] = is;
  if (running_mean != nullptr) {
    const float unbiased = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
  const float g = gamma ? gamma[c] : 1.f, bta = beta ? beta[c] : 0.f;
  scale_shift[c] = g * is;
  scale_shift[C + c] = bta - mu * g * is;
}

__global__ void __launch_bounds__(256) kbn_apply(const uint4* __restrict__ x, uint4* __restrict__ y, long long M,
                                                   int CV, int RL, const float* __restrict__ scale_shift,
                                                   int relu) {
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  const int C = CV * 8;
  float sc[8], sf[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = scale_shift[tx * 8 + k]; sf[k] = scale_shift[C + tx * 8 + k]; }
  const long long step = (long long)gridDim.x * RL;
  long long r = (long long)blockIdx.x * RL + ty;
  for (; r + 3 * step < M; r += 4 * step) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = ld_stream_v4(x + (r + u * step) * CV + tx);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float f[8];
      unpack8(v[u], f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        f[k] = fmaf(f[k], sc[k], sf[k]);
        if (relu) f[k] = fmaxf(f[k], 0.f);
      }
      y[(r + u * step) * CV + tx] = pack8(f);
    }
  }
  for (; r < M; r += step) {
    float f[8];
    unpack8(ld_stream_v4(x + r * CV + tx), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      f[k] = fmaf(f[k], sc[k], sf[k]);
      if (relu) f[k] = fmaxf(f[k], 0.f);
    }
    y[r * CV + tx] = pack8(f);
  }
}

// ------------------------------------------------------------ backward ------
// sums of dy' and dy'*xhat where dy' = dy * [gamma*xhat+beta > 0] (ReLU mask recomputed)
__global__ void __launch_bounds__(256) kbn_bwd_reduce(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                        long long M, int CV, int RL,
                                                        const float* __restrict__ mean,
                                                        const float* __restrict__ invstd,
                                                        const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int relu,
                                                        float* __restrict__ part) {
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  float mu[8], is[8], g[8], b[8], s1[8], s2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = tx * 8 + k;
    mu[k] = mean[c]; is[k] = invstd[c];
    g[k] = gamma ? gamma[c] : 1.f; b[k] = beta ? beta[c] : 0.f;
    s1[k] = 0.f; s2[k] = 0.f;
  }
  const long long step = (long long)gridDim.x * RL;
  long long r = (long long)blockIdx.x * RL + ty;
  // 4 rows x 2 tensors = 8 independent 16-byte loads in flight per thread (ncu: 2 rows left DRAM at 59 %)
  for (; r + 3 * step < M; r += 4 * step) {
    uint4 vx[4], vd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vx[u] = ld_stream_v4(x + (r + u * step) * CV + tx);
      vd[u] = ld_stream_v4(dy + (r + u * step) * CV + tx);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float fx[8], fd[8];
      unpack8(vx[u], fx);
      unpack8(vd[u], fd);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (fx[k] - mu[k]) * is[k];
        const float d = (relu && fmaf(xh, g[k], b[k]) <= 0.f) ? 0.f : fd[k];
        s1[k] += d;
        s2[k] = fmaf(d, xh, s2[k]);
      }
    }
  }
  for (; r < M; r += step) {
    float fx[8], fd[8];
    unpack8(ld_stream_v4(x + r * CV + tx), fx);
    unpack8(ld_stream_v4(dy + r * CV + tx), fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (fx[k] - mu[k]) * is[k];
      const float d = (relu && fmaf(xh, g[k], b[k]) <= 0.f) ? 0.f : fd[k];
      s1[k] += d;
      s2[k] = fmaf(d, xh, s2[k]);
    }
  }
  block_reduce_store(s1, s2, CV, RL, CV * 8, part);
}

// dgamma = sum(dy'*xhat), dbeta = sum(dy'); coef = {gamma*invstd, sum(dy')/M, sum(dy'*xhat)/M}
__global__ void __launch_bounds__(32 * kFinY) kbn_bwd_finalize(const float* __restrict__ part, int nblocks, int C,
                                 long long M,
                                 const float* __restrict__ gamma, const float* __restrict__ invstd,
                                 float* __restrict__ dgamma, float* __restrict__ dbeta, int accumulate_grads,
                                 float* __restrict__ coef) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s1, s2;
  reduce_partials(part, nblocks, C, c, s1, s2);
  if (c >= C || threadIdx.y != 0) return;
  if (dgamma) dgamma[c] = accumulate_grads ? dgamma[c] + s2 : s2;
  if (dbeta) dbeta[c] = accumulate_grads ? dbeta[c] + s1 : s1;
  const float inv_m = 1.f / (float)M;
  coef[c] = (gamma ? gamma[c] : 1.f) * invstd[c];
  coef[C + c] = s1 * inv_m;
  coef[2 * C + c] = s2 * inv_m;
}

__global__ void __launch_bounds__(256) kbn_bwd_apply(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                       uint4* __restrict__ dx, long long M, int CV, int RL,
                                                       const float* __restrict__ mean,
                                                       const float* __restrict__ invstd,
                                                       const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int relu,
                                                       const float* __restrict__ coef,
                                                       const uint4* __restrict__ extra) {
  // extra != nullptr: dx += extra (the gradient arriving at the same tensor through the residual connection)
  const int tx = threadIdx.x % CV, ty = threadIdx.x / CV;
  const int C = CV * 8;
  float mu[8], is[8], g[8], b[8], c0[8], c1[8], c2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int c = tx * 8 + k;
    mu[k] = mean[c]; is[k] = invstd[c];
    g[k] = gamma ? gamma[c] : 1.f; b[k] = beta ? beta[c] : 0.f;
    c0[k] = coef[c]; c1[k] = coef[C + c]; c2[k] = coef[2 * C + c];
  }
  const long long step = (long long)gridDim.x * RL;
  long long r = (long long)blockIdx.x * RL + ty;
  for (; r + 3 * step < M; r += 4 * step) {
    uint4 vx[4], vd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      vx[u] = ld_stream_v4(x + (r + u * step) * CV + tx);
      vd[u] = ld_stream_v4(dy + (r + u * step) * CV + tx);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float fx[8], fd[8], o[8];
      unpack8(vx[u], fx);
      unpack8(vd[u], fd);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float xh = (fx[k] - mu[k]) * is[k];
        const float d = (relu && fmaf(xh, g[k], b[k]) <= 0.f) ? 0.f : fd[k];
        o[k] = c0[k] * (d - c1[k] - xh * c2[k]);
      }
      if (extra != nullptr) {
        float ex[8];
        unpack8(ld_stream_v4(extra + (r + u * step) * CV + tx), ex);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] += ex[k];
      }
      dx[(r + u * step) * CV + tx] = pack8(o);
    }
  }
  for (; r < M; r += step) {
    float fx[8], fd[8], o[8];
    unpack8(ld_stream_v4(x + r * CV + tx), fx);
    unpack8(ld_stream_v4(dy + r * CV + tx), fd);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float xh = (fx[k] - mu[k]) * is[k];
      const float d = (relu && fmaf(xh, g[k], b[k]) <= 0.f) ? 0.f : fd[k];
      o[k] = c0[k] * (d - c1[k] - xh * c2[k]);
    }
    if (extra != nullptr) {
      float ex[8];
      unpack8(ld_stream_v4(extra + r * CV + tx), ex);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += ex[k];
    }
    dx[r * CV + tx] = pack8(o);
  }
}

}  // namespace dlcfn

using namespace dlcfn;

// workspace (floats) needed by either direction: partials [grid][2][C] + 3*C coefficients
DLCFN_API size_t dlcfn_bn_workspace_floats(long long M, int C) {
  BnGeom g = bn_geom(M, C);
  return (size_t)g.grid * 2 * C + 3 * (size_t)C;
}

static inline bool bn_supported(long long M, int C) { return C % 8 == 0 && C >= 8 && C <= 2048 && M > 0; }

// y = relu?(gamma * (x - mean) * invstd + beta); writes mean/invstd (saved for backward) and updates the
// running statistics.  x, y: bf16 [M, C].  gamma/beta/mean/invstd/running_*: fp32 [C].  3 launches.
DLCFN_API int dlcfn_bn_forward_ex(const void* x, const void* x2, void* z, void* y, long long M, int C,
                                  const float* gamma, const float* beta, float* mean, float* invstd,
                                  float* running_mean, float* running_var, float momentum, float eps, int relu,
                                  float* workspace, cudaStream_t stream);

DLCFN_API int dlcfn_bn_forward(const void* x, void* y, long long M, int C, const float* gamma, const float* beta,
                               float* mean, float* invstd, float* running_mean, float* running_var,
                               float momentum, float eps, int relu, float* workspace, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  return dlcfn_bn_forward_ex(x, nullptr, nullptr, y, M, C, gamma, beta, mean, invstd, running_mean, running_var, momentum,
                             eps, relu, workspace, stream);
}

// x2 != nullptr: z = x + x2 is materialised (bf16) and normalised: y = relu?(bn(z)).  4 launches worth of work in 3.
DLCFN_API int dlcfn_bn_forward_ex(const void* x, const void* x2, void* z, void* y, long long M, int C,
                                  const float* gamma, const float* beta, float* mean, float* invstd,
                                  float* running_mean, float* running_var, float momentum, float eps, int relu,
                                  float* workspace, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (!bn_supported(M, C)) return (int)cudaErrorInvalidValue;
  BnGeom g = bn_geom(M, C);
  float* part = workspace;
  float* scale_shift = workspace + (size_t)g.grid * 2 * C;
  const size_t sh = (size_t)g.threads * 16 * sizeof(float);
  kbn_stats<<<g.grid, g.threads, sh, stream>>>((const uint4*)x, (const uint4*)x2, (uint4*)z, M, g.CV, g.RL, part);
  const void* xin = x2 ? z : x;
  kbn_fwd_finalize<<<(C + 31) / 32, dim3(32, kFinY), 0, stream>>>(part, g.grid, C, M, eps, momentum, gamma, beta, mean,
                                                        invstd, running_mean, running_var, scale_shift);
  kbn_apply<<<g.grid, g.threads, 0, stream>>>((const uint4*)xin, (uint4*)y, M, g.CV, g.RL, scale_shift, relu);
  return (int)cudaGetLastError();
}

// Training forward whose statistics pass already happened in the epilogue of the producing convolution
// (gemm_tcgen05.cu, GemmParams::stats): `part` = [part_rows][2][C] partial sums / sums of squares of x.  2 launches.
DLCFN_API int dlcfn_bn_forward_pre(const void* x, void* y, long long M, int C, const float* gamma, const float* beta,
                                   float* mean, float* invstd, float* running_mean, float* running_var, float momentum,
                                   float eps, int relu, const float* part, int part_rows, float* workspace,
                                   cudaStream_t stream) {
  (void)cudaGetLastError();
  if (!bn_supported(M, C) || part == nullptr || part_rows <= 0) return (int)cudaErrorInvalidValue;
  BnGeom g = bn_geom(M, C);
  float* scale_shift = workspace + (size_t)g.grid * 2 * C;
  kbn_fwd_finalize<<<(C + 31) / 32, dim3(32, kFinY), 0, stream>>>(part, part_rows, C, M, eps, momentum, gamma, beta, mean,
                                                        invstd, running_mean, running_var, scale_shift);
  kbn_apply<<<g.grid, g.threads, 0, stream>>>((const uint4*)x, (uint4*)y, M, g.CV, g.RL, scale_shift, relu);
  return (int)cudaGetLastError();
}

// inference / frozen-statistics forward: one pass
DLCFN_API int dlcfn_bn_inference(const void* x, void* y, long long M, int C, const float* gamma,
                                 const float* beta, const float* running_mean, const float* running_var,
                                 float eps, int relu, float* workspace, cudaStream_t stream);

DLCFN_API int dlcfn_bn_backward_ex(const void* x, const void* dy, const void* extra, void* dx, long long M, int C,
                                   const float* gamma, const float* beta, const float* mean, const float* invstd,
                                   float* dgamma, float* dbeta, int accumulate_grads, int relu, float* workspace,
                                   cudaStream_t stream);

DLCFN_API int dlcfn_bn_backward(const void* x, const void* dy, void* dx, long long M, int C, const float* gamma,
                                const float* beta, const float* mean, const float* invstd, float* dgamma,
                                float* dbeta, int accumulate_grads, int relu, float* workspace,
                                cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  return dlcfn_bn_backward_ex(x, dy, nullptr, dx, M, C, gamma, beta, mean, invstd, dgamma, dbeta, accumulate_grads, relu,
                              workspace, stream);
}

// extra != nullptr: dx = bn_backward(dy) + extra  (gradient of the fused residual add)
DLCFN_API int dlcfn_bn_backward_ex(const void* x, const void* dy, const void* extra, void* dx, long long M, int C,
                                   const float* gamma, const float* beta, const float* mean, const float* invstd,
                                   float* dgamma, float* dbeta, int accumulate_grads, int relu, float* workspace,
                                   cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (!bn_supported(M, C)) return (int)cudaErrorInvalidValue;
  BnGeom g = bn_geom(M, C);
  float* part = workspace;
  float* coef = workspace + (size_t)g.grid * 2 * C;
  const size_t sh = (size_t)g.threads * 16 * sizeof(float);
  kbn_bwd_reduce<<<g.grid, g.threads, sh, stream>>>((const uint4*)x, (const uint4*)dy, M, g.CV, g.RL, mean,
                                                    invstd, gamma, beta, relu, part);
  kbn_bwd_finalize<<<(C + 31) / 32, dim3(32, kFinY), 0, stream>>>(part, g.grid, C, M, gamma, invstd, dgamma, dbeta,
                                                        accumulate_grads, coef);
  kbn_bwd_apply<<<g.grid, g.threads, 0, stream>>>((const uint4*)x, (const uint4*)dy, (uint4*)dx, M, g.CV, g.RL,
                                                  mean, invstd, gamma, beta, relu, coef, (const uint4*)extra);
  return (int)cudaGetLastError();
}

namespace dlcfn {
__global__ void kbn_inf_coef(int C, const float* __restrict__ gamma, const float* __restrict__ beta,
                             const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                             float* __restrict__ scale_shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float is = rsqrtf(rv[c] + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale_shift[c] = g * is;
  scale_shift[C + c] = b - rm[c] * g * is;
}
}  // namespace dlcfn

DLCFN_API int dlcfn_bn_inference(const void* x, void* y, long long M, int C, const float* gamma,
                                 const float* beta, const float* running_mean, const float* running_var,
                                 float eps, int relu, float* workspace, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (!bn_supported(M, C)) return (int)cudaErrorInvalidValue;
  BnGeom g = bn_geom(M, C);
  float* scale_shift = workspace + (size_t)g.grid * 2 * C;
  kbn_inf_coef<<<(C + 127) / 128, 128, 0, stream>>>(C, gamma, beta, running_mean, running_var, eps, scale_shift);
  kbn_apply<<<g.grid, g.threads, 0, stream>>>((const uint4*)x, (uint4*)y, M, g.CV, g.RL, scale_shift, relu);
  return (int)cudaGetLastError();
}
