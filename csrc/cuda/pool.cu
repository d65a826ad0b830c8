// NHWC bf16 max-pooling, forward and backward (sm_100a).
//
// Reference M5: cuDNN pooling or hand kernels `pool_max_2d_gpu_kernel` / `unpool_max_2d_gpu_kernel`, one thread
// per output element (MX/src/operator/nn/pool.cuh:96-320, 689-936).  Profiling the B200 ResNet-50 step showed the
// framework's channels-last max-pool taking 0.75 ms forward + 1.73 ms backward for the 112x112x64 stem tensor -
// 10-17x the time its bytes need.  These kernels stream it once: a thread owns 8 channels (one 16-byte vector)
// of one pixel; forward also records the arg-max tap (uint8 per element) so backward is a pure gather with no
// atomics: every INPUT pixel looks at the (at most ceil(k/s)^2) windows covering it and takes dy where it was
// the arg-max.
#include "common.cuh"

namespace dlcfn {

struct PoolGeom {
  int N, H, W, CV, OH, OW, K, S, P;
};

__device__ __forceinline__ void unpack8p(const uint4& v, float* f) {
  f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
  f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}

__global__ void __launch_bounds__(256) kmaxpool_fwd(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                     uint2* __restrict__ idx, PoolGeom g) {
  const long long total = (long long)g.N * g.OH * g.OW * g.CV;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % g.CV);
    long long r = t / g.CV;
    const int ow = (int)(r % g.OW); r /= g.OW;
    const int oh = (int)(r % g.OH);
    const int n = (int)(r / g.OH);
    float best[8];
    uint32_t arg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; arg[k] = 0; }
    for (int kh = 0; kh < g.K; ++kh) {
      const int ih = oh * g.S - g.P + kh;
      if (ih < 0 || ih >= g.H) continue;
      for (int kw = 0; kw < g.K; ++kw) {
        const int iw = ow * g.S - g.P + kw;
        if (iw < 0 || iw >= g.W) continue;
        float f[8];
        unpack8p(x[(((long long)n * g.H + ih) * g.W + iw) * g.CV + cv], f);
        const uint32_t tap = (uint32_t)(kh * g.K + kw);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (f[k] > best[k] || (f[k] != f[k])) { best[k] = f[k]; arg[k] = tap; }
      }
    }
    y[t] = make_uint4(pack_bf16(best[0], best[1]), pack_bf16(best[2], best[3]), pack_bf16(best[4], best[5]),
                      pack_bf16(best[6], best[7]));
    idx[t] = make_uint2(arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24),
                        arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24));
  }
}

__global__ void __launch_bounds__(256) kmaxpool_bwd(const uint4* __restrict__ dy, const uint2* __restrict__ idx,
                                                     uint4* __restrict__ dx, PoolGeom g) {
  const long long total = (long long)g.N * g.H * g.W * g.CV;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(t % g.CV);
    long long r = t / g.CV;
    const int iw = (int)(r % g.W); r /= g.W;
    const int ih = (int)(r % g.H);
    const int n = (int)(r / g.H);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    // windows (oh, ow) with oh*S - P <= ih <= oh*S - P + K - 1
    const int oh_lo = max(0, (ih + g.P - g.K + g.S) / g.S), oh_hi = min(g.OH - 1, (ih + g.P) / g.S);
    const int ow_lo = max(0, (iw + g.P - g.K + g.S) / g.S), ow_hi = min(g.OW - 1, (iw + g.P) / g.S);
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
      const int kh = ih - (oh * g.S - g.P);
      for (int ow = ow_lo; ow <= ow_hi; ++ow) {
        const int kw = iw - (ow * g.S - g.P);
        const uint32_t tap = (uint32_t)(kh * g.K + kw);
        const long long o = (((long long)n * g.OH + oh) * g.OW + ow) * g.CV + cv;
        const uint2 a = idx[o];
        float d[8];
        unpack8p(dy[o], d);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t ak = ((k < 4 ? a.x : a.y) >> (8 * (k & 3))) & 0xffu;
          if (ak == tap) acc[k] += d[k];
        }
      }
    }
    dx[t] = make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]),
                       pack_bf16(acc[6], acc[7]));
  }
}

}  // namespace dlcfn

using namespace dlcfn;

static inline int pool_grid(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148LL * 16) b = 148LL * 16;
  return b < 1 ? 1 : (int)b;
}

// x: bf16 [N,H,W,C] (C % 8 == 0) -> y [N,OH,OW,C], idx: one byte per output element (arg-max tap)
DLCFN_API int dlcfn_maxpool_nhwc_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, int K, int S,
                                     int P, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (C % 8 || K < 1 || K > 15 || S < 1) return (int)cudaErrorInvalidValue;
  PoolGeom g{N, H, W, C / 8, (H + 2 * P - K) / S + 1, (W + 2 * P - K) / S + 1, K, S, P};
  const long long total = (long long)N * g.OH * g.OW * g.CV;
  if (total == 0) return 0;
  kmaxpool_fwd<<<pool_grid(total), 256, 0, stream>>>((const uint4*)x, (uint4*)y, (uint2*)idx, g);
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_maxpool_nhwc_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, int K,
                                     int S, int P, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (C % 8 || K < 1 || K > 15 || S < 1) return (int)cudaErrorInvalidValue;
  PoolGeom g{N, H, W, C / 8, (H + 2 * P - K) / S + 1, (W + 2 * P - K) / S + 1, K, S, P};
  const long long total = (long long)N * H * W * g.CV;
  if (total == 0) return 0;
  kmaxpool_bwd<<<pool_grid(total), 256, 0, stream>>>((const uint4*)dy, (const uint2*)idx, (uint4*)dx, g);
  return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------
// Stem input transform: NHWC 8-channel image (3 real channels, output of the padded bn_data) -> space-to-depth-by-2
// NHWC 16-channel image with the 7x7/pad-3 convolution's zero border folded in:
//     s2d[n, i, j, c*4 + bh*2 + bw] = x[n, 2i+bh-3, 2j+bw-3, c]   (0 outside the image),  c < 3, channels 12..15 = 0
// so that the stem convolution becomes a 4x4 stride-1 convolution with K = 256 (models/resnet.py:StemConv).  One
// pass each way; the framework's pad/reshape/permute/contiguous chain cost 2.7 ms per step on B200.
namespace dlcfn {

__global__ void __launch_bounds__(256) kstem_s2d_fwd(const uint4* __restrict__ x, uint4* __restrict__ y, int N, int H,
                                                      int W, int Hs, int Ws) {
  const long long total = (long long)N * Hs * Ws;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(t % Ws);
    const int i = (int)((t / Ws) % Hs);
    const int n = (int)(t / ((long long)Ws * Hs));
    uint16_t v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = 0;
#pragma unroll
    for (int bh = 0; bh < 2; ++bh)
#pragma unroll
      for (int bw = 0; bw < 2; ++bw) {
        const int r = 2 * i + bh - 3, c = 2 * j + bw - 3;
        if (r >= 0 && r < H && c >= 0 && c < W) {
          const uint4 p = x[((long long)n * H + r) * W + c];
          v[0 * 4 + bh * 2 + bw] = (uint16_t)(p.x & 0xffff);
          v[1 * 4 + bh * 2 + bw] = (uint16_t)(p.x >> 16);
          v[2 * 4 + bh * 2 + bw] = (uint16_t)(p.y & 0xffff);
        }
      }
    uint4 o0, o1;
    o0.x = v[0] | ((uint32_t)v[1] << 16); o0.y = v[2] | ((uint32_t)v[3] << 16);
    o0.z = v[4] | ((uint32_t)v[5] << 16); o0.w = v[6] | ((uint32_t)v[7] << 16);
    o1.x = v[8] | ((uint32_t)v[9] << 16); o1.y = v[10] | ((uint32_t)v[11] << 16);
    o1.z = 0; o1.w = 0;
    y[2 * t] = o0;
    y[2 * t + 1] = o1;
  }
}

__global__ void __launch_bounds__(256) kstem_s2d_bwd(const uint16_t* __restrict__ dy, uint4* __restrict__ dx, int N, int H,
                                                      int W, int Hs, int Ws) {
  const long long total = (long long)N * H * W;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % W);
    const int r = (int)((t / W) % H);
    const int n = (int)(t / ((long long)W * H));
    const int i = (r + 3) >> 1, bh = (r + 3) & 1, j = (c + 3) >> 1, bw = (c + 3) & 1;
    const uint16_t* src = dy + (((long long)n * Hs + i) * Ws + j) * 16 + bh * 2 + bw;
    const uint32_t a = src[0], b = src[4], d = src[8];
    dx[t] = make_uint4(a | (b << 16), d, 0u, 0u);
  }
}

}  // namespace dlcfn

DLCFN_API int dlcfn_stem_s2d_fwd(const void* x8, void* y16, int N, int H, int W, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if ((H & 1) || (W & 1)) return (int)cudaErrorInvalidValue;
  const int Hs = (H + 6) / 2, Ws = (W + 6) / 2;
  const long long total = (long long)N * Hs * Ws;
  if (total == 0) return 0;
  dlcfn::kstem_s2d_fwd<<<pool_grid(total), 256, 0, stream>>>((const uint4*)x8, (uint4*)y16, N, H, W, Hs, Ws);
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_stem_s2d_bwd(const void* dy16, void* dx8, int N, int H, int W, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if ((H & 1) || (W & 1)) return (int)cudaErrorInvalidValue;
  const int Hs = (H + 6) / 2, Ws = (W + 6) / 2;
  const long long total = (long long)N * H * W;
  if (total == 0) return 0;
  dlcfn::kstem_s2d_bwd<<<pool_grid(total), 256, 0, stream>>>((const uint16_t*)dy16, (uint4*)dx8, N, H, W, Hs, Ws);
  return (int)cudaGetLastError();
}

// 3-channel NHWC bf16 image -> 8-channel NHWC (channels 3..7 zero): 6 bytes in, one 16-byte vector out per pixel.
namespace dlcfn {
__global__ void __launch_bounds__(256) kpad3to8(const uint16_t* __restrict__ x, uint4* __restrict__ y, long long npix) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < npix; t += (long long)gridDim.x * blockDim.x) {
    const uint16_t* p = x + 3 * t;
    const uint32_t a = p[0], b = p[1], c = p[2];
    y[t] = make_uint4(a | (b << 16), c, 0u, 0u);
  }
}
}  // namespace dlcfn

DLCFN_API int dlcfn_pad3to8_nhwc(const void* x3, void* y8, long long npix, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (npix == 0) return 0;
  dlcfn::kpad3to8<<<pool_grid(npix), 256, 0, stream>>>((const uint16_t*)x3, (uint4*)y8, npix);
  return (int)cudaGetLastError();
}
