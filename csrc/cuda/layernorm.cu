// Fused dropout + residual add + LayerNorm for transformer blocks (sm_100a), forward and backward.
//
//   z = x + dropout(h) ;  y = LayerNorm(z) * gamma + beta            (post-LN BERT: both sub-layers end like this)
//
// The PyTorch-eager form is four kernels forward (dropout: read h, write h' + mask; add: read h', x, write z; LN: read
// z, write y) and about eight backward.  Here one kernel each way touches every tensor once:
//   forward   reads h, x             writes z, y, 1 mask BIT per element, mean/rstd per row
//   backward  reads dy, z, mask      writes dz (gradient of the residual branch), dh = dz * mask / keep, and per-CTA
//             partial sums of dgamma / dbeta that a small finalize kernel reduces in a fixed order (deterministic)
// One warp owns one row: the row lives in registers (H / 32 values per lane, H <= 8192, H % 256 == 0), statistics are
// two warp-shuffle reductions, all global traffic is 16-byte vectors.  The dropout mask comes from a counter-based
// generator (Philox-4x32-10 keyed by (seed, element index)), so the same (seed, offset) reproduces the same mask
// regardless of the launch geometry; p = 0 skips the generator and the mask entirely.
//
// Reference: the tree has no transformer; the BERT-large configuration comes from BASELINE.json and the
// ops it replaces are LayerNorm (src/operator/nn/layer_norm.cu), Dropout (nn/dropout.cu) and elemwise add.
#include "common.cuh"

namespace dlcfn {

__device__ __forceinline__ void unpack8l(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8l(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

// ---- Philox-4x32-10 ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint2 key, uint4 ctr) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

// keep-mask of 8 consecutive elements starting at element index `e` (e % 8 == 0): one Philox call gives 4 words, two
// 16-bit lanes of each decide one element
__device__ __forceinline__ uint32_t keep_mask8(unsigned long long seed, unsigned long long offset, unsigned long long e,
                                               uint32_t thresh16) {
  const unsigned long long c = offset + (e >> 3);
  const uint4 r = philox4x32(make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)),
                             make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x6c6e6466u, 0u));
  uint32_t m = 0;
  m |= ((r.x & 0xffffu) >= thresh16) << 0; m |= ((r.x >> 16) >= thresh16) << 1;
  m |= ((r.y & 0xffffu) >= thresh16) << 2; m |= ((r.y >> 16) >= thresh16) << 3;
  m |= ((r.z & 0xffffu) >= thresh16) << 4; m |= ((r.z >> 16) >= thresh16) << 5;
  m |= ((r.w & 0xffffu) >= thresh16) << 6; m |= ((r.w >> 16) >= thresh16) << 7;
  return m;   // bit k set = element e + k is kept
}

struct LNArgs {
  long long rows;
  int H;                      // row length, H % 256 == 0, H <= 8192
  float eps, p, scale;        // dropout probability and 1 / (1 - p)
  unsigned long long seed, offset;
};

// VPL = 16-byte vectors per lane = H / 256
template <int VPL>
__global__ void __launch_bounds__(256) kdal_fwd(const uint4* __restrict__ h, const uint4* __restrict__ x,
                                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                                 uint4* __restrict__ z, uint4* __restrict__ y, uint8_t* __restrict__ mask,
                                                 float* __restrict__ mean, float* __restrict__ rstd, LNArgs a) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int HV = a.H / 8;
  const uint32_t thresh16 = (uint32_t)(a.p * 65536.f);
  for (long long row = warp; row < a.rows; row += nwarps) {
    float v[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int cv = i * 32 + lane;                       // vector index inside the row: coalesced across the warp
      float fh[8], fx[8];
      unpack8l(ld_stream_v4(h + row * HV + cv), fh);
      if (x != nullptr) unpack8l(ld_stream_v4(x + row * HV + cv), fx);
      uint32_t m = 0xffu;
      if (a.p > 0.f) {
        m = keep_mask8(a.seed, a.offset, (unsigned long long)(row * a.H + (long long)cv * 8), thresh16);
        mask[row * HV + cv] = (uint8_t)m;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float t = ((m >> k) & 1u) ? fh[k] * a.scale : 0.f;
        if (x != nullptr) t += fx[k];
        v[i][k] = t;
      }
      const uint4 zz = pack8l(v[i]);
      if (z != nullptr) z[row * HV + cv] = zz;
      unpack8l(zz, v[i]);                                  // statistics over the ROUNDED sum: what backward re-reads
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[i][k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mu = s / (float)a.H;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int k = 0; k < 8; ++k) { const float d = v[i][k] - mu; q = fmaf(d, d, q); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rs = rsqrtf(q / (float)a.H + a.eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int cv = i * 32 + lane;
      const float4 g0 = reinterpret_cast<const float4*>(gamma)[cv * 2], g1 = reinterpret_cast<const float4*>(gamma)[cv * 2 + 1];
      const float4 b0 = reinterpret_cast<const float4*>(beta)[cv * 2], b1 = reinterpret_cast<const float4*>(beta)[cv * 2 + 1];
      const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf((v[i][k] - mu) * rs, g[k], b[k]);
      y[row * HV + cv] = pack8l(o);
    }
  }
}

// dz = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma ;  dh = dz * mask * scale
// partial[(blockIdx * 2 + 0) * H + c] = sum over the CTA's rows of dy * xhat (dgamma), + 1: sum of dy (dbeta)
template <int VPL>
__global__ void __launch_bounds__(256) kdal_bwd(const uint4* __restrict__ dy, const uint4* __restrict__ z,
                                                 const uint8_t* __restrict__ mask, const float* __restrict__ gamma,
                                                 const float* __restrict__ mean, const float* __restrict__ rstd,
                                                 uint4* __restrict__ dz, uint4* __restrict__ dh,
                                                 float* __restrict__ partial, LNArgs a) {
  extern __shared__ float sm[];            // [8 warps][2][H]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int HV = a.H / 8;
  float gsum[VPL][8], bsum[VPL][8];        // this warp's dgamma / dbeta contributions
  float gam[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int cv = i * 32 + lane;
    const float4 g0 = reinterpret_cast<const float4*>(gamma)[cv * 2], g1 = reinterpret_cast<const float4*>(gamma)[cv * 2 + 1];
    gam[i][0] = g0.x; gam[i][1] = g0.y; gam[i][2] = g0.z; gam[i][3] = g0.w;
    gam[i][4] = g1.x; gam[i][5] = g1.y; gam[i][6] = g1.z; gam[i][7] = g1.w;
#pragma unroll
    for (int k = 0; k < 8; ++k) { gsum[i][k] = 0.f; bsum[i][k] = 0.f; }
  }
  for (long long row = warp; row < a.rows; row += nwarps) {
    const float mu = mean[row], rs = rstd[row];
    float g[VPL][8], xh[VPL][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int cv = i * 32 + lane;
      float fd[8], fz[8];
      unpack8l(ld_stream_v4(dy + row * HV + cv), fd);
      unpack8l(ld_stream_v4(z + row * HV + cv), fz);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xh[i][k] = (fz[k] - mu) * rs;
        g[i][k] = fd[k] * gam[i][k];
        s1 += g[i][k];
        s2 = fmaf(g[i][k], xh[i][k], s2);
        gsum[i][k] = fmaf(fd[k], xh[i][k], gsum[i][k]);
        bsum[i][k] += fd[k];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 / (float)a.H, m2 = s2 / (float)a.H;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int cv = i * 32 + lane;
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = rs * (g[i][k] - m1 - xh[i][k] * m2);
      const uint4 packed = pack8l(o);
      dz[row * HV + cv] = packed;
      if (dh != nullptr) {
        if (a.p > 0.f) {
          unpack8l(packed, o);
          const uint32_t m = mask[row * HV + cv];
#pragma unroll
          for (int k = 0; k < 8; ++k) o[k] = ((m >> k) & 1u) ? o[k] * a.scale : 0.f;
          dh[row * HV + cv] = pack8l(o);
        } else {
          dh[row * HV + cv] = packed;
        }
      }
    }
  }
  // CTA-level, fixed-order reduction of the 8 warps' column sums
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = (i * 32 + lane) * 8 + k;
      sm[(w * 2 + 0) * a.H + c] = gsum[i][k];
      sm[(w * 2 + 1) * a.H + c] = bsum[i][k];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < a.H; c += blockDim.x) {
    float sg = 0.f, sb = 0.f;
#pragma unroll
    for (int ww = 0; ww < 8; ++ww) { sg += sm[(ww * 2 + 0) * a.H + c]; sb += sm[(ww * 2 + 1) * a.H + c]; }
    partial[((size_t)blockIdx.x * 2 + 0) * a.H + c] = sg;
    partial[((size_t)blockIdx.x * 2 + 1) * a.H + c] = sb;
  }
}

__global__ void __launch_bounds__(256) kdal_finalize(const float* __restrict__ partial, int nblocks, int H,
                                                      float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float sg = 0.f, sb = 0.f;
  for (int b = 0; b < nblocks; ++b) {
    sg += partial[((size_t)b * 2 + 0) * H + c];
    sb += partial[((size_t)b * 2 + 1) * H + c];
  }
  dgamma[c] = sg;
  dbeta[c] = sb;
}

}  // namespace dlcfn

using namespace dlcfn;

static inline int dal_grid(long long rows) {
  long long b = (rows + 7) / 8;            // 8 warps (rows) per CTA per pass
  if (b > 148 * 2) b = 148 * 2;
  return b < 1 ? 1 : (int)b;
}

DLCFN_API int dlcfn_dal_bwd_blocks(long long rows) { return dal_grid(rows); }

#define DAL_DISPATCH(KERN, ...)                                                          \
  switch (H / 256) {                                                                     \
    case 1: KERN<1> __VA_ARGS__; break;                                                  \
    case 2: KERN<2> __VA_ARGS__; break;                                                  \
    case 3: KERN<3> __VA_ARGS__; break;                                                  \
    case 4: KERN<4> __VA_ARGS__; break;                                                  \
    case 6: KERN<6> __VA_ARGS__; break;                                                  \
    case 8: KERN<8> __VA_ARGS__; break;                                                  \
    default: return (int)cudaErrorInvalidValue;                                          \
  }

// h, x (nullable residual), z (nullable), y: bf16 [rows, H]; gamma/beta fp32 [H]; mask: rows*H/8 bytes (p > 0);
// mean/rstd fp32 [rows].  H in {256, 512, 768, 1024, 1536, 2048}.
DLCFN_API int dlcfn_dropout_add_ln_fwd(const void* h, const void* x, const float* gamma, const float* beta, void* z, void* y,
                                       void* mask, float* mean, float* rstd, long long rows, int H, float eps, float p,
                                       unsigned long long seed, unsigned long long offset, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (rows == 0) return 0;
  if (H % 256 || H > 2048 || p < 0.f || p >= 1.f) return (int)cudaErrorInvalidValue;
  LNArgs a{rows, H, eps, p, p > 0.f ? 1.f / (1.f - p) : 1.f, seed, offset};
  const int grid = dal_grid(rows);
  DAL_DISPATCH(kdal_fwd, <<<grid, 256, 0, stream>>>((const uint4*)h, (const uint4*)x, gamma, beta, (uint4*)z, (uint4*)y,
                                                     (uint8_t*)mask, mean, rstd, a))
  return (int)cudaGetLastError();
}

// partial: fp32 scratch of dlcfn_dal_bwd_blocks(rows) * 2 * H floats; dh nullable (no dropout branch gradient wanted)
DLCFN_API int dlcfn_dropout_add_ln_bwd(const void* dy, const void* z, const void* mask, const float* gamma,
                                       const float* mean, const float* rstd, void* dz, void* dh, float* partial,
                                       float* dgamma, float* dbeta, long long rows, int H, float p, cudaStream_t stream) {
  (void)cudaGetLastError();
  if (rows == 0) return 0;
  if (H % 256 || H > 2048 || p < 0.f || p >= 1.f) return (int)cudaErrorInvalidValue;
  LNArgs a{rows, H, 0.f, p, p > 0.f ? 1.f / (1.f - p) : 1.f, 0ull, 0ull};
  const int grid = dal_grid(rows);
  const size_t smem = (size_t)8 * 2 * H * sizeof(float);
  if (smem > 48 * 1024) {
    static bool configured[9] = {false};
    const int idx = H / 256;
    if (!configured[idx]) {
      cudaError_t e = cudaSuccess;
      switch (idx) {
        case 4: e = cudaFuncSetAttribute(kdal_bwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
        case 6: e = cudaFuncSetAttribute(kdal_bwd<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
        case 8: e = cudaFuncSetAttribute(kdal_bwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); break;
        default: break;
      }
      if (e != cudaSuccess) return (int)e;
      configured[idx] = true;
    }
  }
  DAL_DISPATCH(kdal_bwd, <<<grid, 256, smem, stream>>>((const uint4*)dy, (const uint4*)z, (const uint8_t*)mask, gamma, mean,
                                                        rstd, (uint4*)dz, (uint4*)dh, partial, a))
  DLCFN_CHECK(cudaGetLastError());
  kdal_finalize<<<(H + 255) / 256, 256, 0, stream>>>(partial, grid, H, dgamma, dbeta);
  return (int)cudaGetLastError();
}
