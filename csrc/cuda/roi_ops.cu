// Detection-head kernels for the Mask R-CNN workload (W8: examples/maskrcnn_train.py), NHWC, sm_100a.
// EXPERIMENTAL: written at the end of round 1 without GPU access; gated tests in tests/test_gpu_experimental.py.
//
//   ROIAlign forward / backward   reference: MX/src/operator/contrib/roi_align.cu:40-467 (bilinear sampling grid per bin,
//                                 average of sampling_ratio^2 points; backward scatters with atomics)
//   NMS                           reference: the overlap-bitmask scheme of MX/src/operator/contrib/multi_proposal.cu:275-400
//                                 (64 x 64 box tiles -> one 64-bit suppression word per (box, tile)), followed here by an
//                                 on-device greedy sweep instead of a copy to the host
//
// Layout: features are NHWC ([N,H,W,C], channels_last), so the C threads of a bin read consecutive addresses for each of
// the four bilinear taps - one 128-byte line per warp and tap - and the pooled output [R,PH,PW,C] is written the same way.
#include "common.cuh"

namespace dlcfn {

struct RoiGeom {
  int N, H, W, C, R, PH, PW;
  float scale;
  int sampling_ratio;
  int aligned;
};

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

// per-bin sampling set-up shared by forward and backward
struct BinSetup {
  int batch;
  float start_w, start_h, bin_w, bin_h;
  int grid_w, grid_h;
  float inv_count;
  bool valid;
};

__device__ __forceinline__ BinSetup roi_setup(const float* __restrict__ rois, int r, const RoiGeom& g) {
  BinSetup s;
  const float* roi = rois + (size_t)r * 5;
  s.batch = (int)roi[0];
  const float off = g.aligned ? 0.5f : 0.f;
  s.start_w = roi[1] * g.scale - off;
  s.start_h = roi[2] * g.scale - off;
  float rw = roi[3] * g.scale - off - s.start_w;
  float rh = roi[4] * g.scale - off - s.start_h;
  if (!g.aligned) {   // legacy behaviour: force malformed boxes to 1 x 1
    rw = fmaxf(rw, 1.f);
    rh = fmaxf(rh, 1.f);
  }
  s.bin_w = rw / (float)g.PW;
  s.bin_h = rh / (float)g.PH;
  s.grid_w = g.sampling_ratio > 0 ? g.sampling_ratio : (int)ceilf(rw / (float)g.PW);
  s.grid_h = g.sampling_ratio > 0 ? g.sampling_ratio : (int)ceilf(rh / (float)g.PH);
  const int cnt = max(s.grid_w * s.grid_h, 1);
  s.inv_count = 1.f / (float)cnt;
  s.valid = s.batch >= 0 && s.batch < g.N;
  return s;
}

// bilinear taps of one sample point: four (offset, weight) pairs; returns false when the point lies outside the map
__device__ __forceinline__ bool bilinear_taps(float y, float x, int H, int W, int& y_lo, int& x_lo, int& y_hi, int& x_hi,
                                              float& w1, float& w2, float& w3, float& w4) {
  if (y < -1.f || y > (float)H || x < -1.f || x > (float)W) return false;
  y = fmaxf(y, 0.f);
  x = fmaxf(x, 0.f);
  y_lo = (int)y;
  x_lo = (int)x;
  if (y_lo >= H - 1) { y_hi = y_lo = H - 1; y = (float)y_lo; } else { y_hi = y_lo + 1; }
  if (x_lo >= W - 1) { x_hi = x_lo = W - 1; x = (float)x_lo; } else { x_hi = x_lo + 1; }
  const float ly = y - (float)y_lo, lx = x - (float)x_lo, hy = 1.f - ly, hx = 1.f - lx;
  w1 = hy * hx; w2 = hy * lx; w3 = ly * hx; w4 = ly * lx;
  return true;
}

template <typename T>
__global__ void __launch_bounds__(256) kroi_align_fwd(const T* __restrict__ x, const float* __restrict__ rois,
                                                      T* __restrict__ y, RoiGeom g) {
  const long long total = (long long)g.R * g.PH * g.PW * g.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % g.C);
    long long t = i / g.C;
    const int pw = (int)(t % g.PW);
    t /= g.PW;
    const int ph = (int)(t % g.PH);
    const int r = (int)(t / g.PH);
    const BinSetup s = roi_setup(rois, r, g);
    float acc = 0.f;
    if (s.valid) {
      const T* base = x + (size_t)s.batch * g.H * g.W * g.C + c;
      for (int iy = 0; iy < s.grid_h; ++iy) {
        const float yy = s.start_h + ph * s.bin_h + ((float)iy + 0.5f) * s.bin_h / (float)s.grid_h;
        for (int ix = 0; ix < s.grid_w; ++ix) {
          const float xx = s.start_w + pw * s.bin_w + ((float)ix + 0.5f) * s.bin_w / (float)s.grid_w;
          int yl, xl, yh, xh;
          float w1, w2, w3, w4;
          if (!bilinear_taps(yy, xx, g.H, g.W, yl, xl, yh, xh, w1, w2, w3, w4)) continue;
          acc += w1 * to_f<T>(base[((size_t)yl * g.W + xl) * g.C]) + w2 * to_f<T>(base[((size_t)yl * g.W + xh) * g.C]) +
                 w3 * to_f<T>(base[((size_t)yh * g.W + xl) * g.C]) + w4 * to_f<T>(base[((size_t)yh * g.W + xh) * g.C]);
        }
      }
      acc *= s.inv_count;
    }
    y[i] = from_f<T>(acc);
  }
}

// dx is an fp32 accumulation buffer [N,H,W,C] (zeroed by the caller); red.global.add.f32 per tap, coalesced over channels
template <typename T>
__global__ void __launch_bounds__(256) kroi_align_bwd(const T* __restrict__ dy, const float* __restrict__ rois,
                                                      float* __restrict__ dx, RoiGeom g) {
  const long long total = (long long)g.R * g.PH * g.PW * g.C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % g.C);
    long long t = i / g.C;
    const int pw = (int)(t % g.PW);
    t /= g.PW;
    const int ph = (int)(t % g.PH);
    const int r = (int)(t / g.PH);
    const BinSetup s = roi_setup(rois, r, g);
    if (!s.valid) continue;
    const float gsum = to_f<T>(dy[i]) * s.inv_count;
    float* base = dx + (size_t)s.batch * g.H * g.W * g.C + c;
    for (int iy = 0; iy < s.grid_h; ++iy) {
      const float yy = s.start_h + ph * s.bin_h + ((float)iy + 0.5f) * s.bin_h / (float)s.grid_h;
      for (int ix = 0; ix < s.grid_w; ++ix) {
        const float xx = s.start_w + pw * s.bin_w + ((float)ix + 0.5f) * s.bin_w / (float)s.grid_w;
        int yl, xl, yh, xh;
        float w1, w2, w3, w4;
        if (!bilinear_taps(yy, xx, g.H, g.W, yl, xl, yh, xh, w1, w2, w3, w4)) continue;
        atomicAdd(base + ((size_t)yl * g.W + xl) * g.C, gsum * w1);
        atomicAdd(base + ((size_t)yl * g.W + xh) * g.C, gsum * w2);
        atomicAdd(base + ((size_t)yh * g.W + xl) * g.C, gsum * w3);
        atomicAdd(base + ((size_t)yh * g.W + xh) * g.C, gsum * w4);
      }
    }
  }
}

// ----------------------------------------------------------------------------------------------------------- NMS
// boxes [n,4] (x1,y1,x2,y2), sorted by descending score.  mask[i][jb] bit t set <=> box jb*64+t (> i) overlaps box i by
// more than `thr` (IoU).  One block per (row tile, column tile), 64 threads; column boxes staged in shared memory.
__global__ void __launch_bounds__(64) knms_mask(const float* __restrict__ boxes, int n, float thr,
                                                unsigned long long* __restrict__ mask, int col_blocks) {
  const int row0 = blockIdx.y * 64, col0 = blockIdx.x * 64;
  if (col0 + 64 <= row0) {       // every column box precedes every row box: nothing to record (words stay zero)
    return;
  }
  __shared__ float cb[64 * 4];
  const int ncol = min(64, n - col0), nrow = min(64, n - row0);
  if ((int)threadIdx.x < ncol) {
    const float4 b = reinterpret_cast<const float4*>(boxes)[col0 + threadIdx.x];
    cb[threadIdx.x * 4 + 0] = b.x; cb[threadIdx.x * 4 + 1] = b.y; cb[threadIdx.x * 4 + 2] = b.z; cb[threadIdx.x * 4 + 3] = b.w;
  }
  __syncthreads();
  if ((int)threadIdx.x < nrow) {
    const int i = row0 + threadIdx.x;
    const float4 a = reinterpret_cast<const float4*>(boxes)[i];
    const float area_a = (a.z - a.x) * (a.w - a.y);
    unsigned long long bits = 0;
    const int start = (row0 == col0) ? (int)threadIdx.x + 1 : 0;
    for (int t = start; t < ncol; ++t) {
      const float x1 = fmaxf(a.x, cb[t * 4]), y1 = fmaxf(a.y, cb[t * 4 + 1]);
      const float x2 = fminf(a.z, cb[t * 4 + 2]), y2 = fminf(a.w, cb[t * 4 + 3]);
      const float iw = fmaxf(x2 - x1, 0.f), ih = fmaxf(y2 - y1, 0.f);
      const float inter = iw * ih;
      const float area_b = (cb[t * 4 + 2] - cb[t * 4]) * (cb[t * 4 + 3] - cb[t * 4 + 1]);
      if (inter > thr * (area_a + area_b - inter)) bits |= 1ull << t;
    }
    mask[(size_t)i * col_blocks + blockIdx.x] = bits;
  }
}

// Greedy sweep on the device: one warp walks the boxes in score order; the suppression bitmap lives in shared memory.
__global__ void __launch_bounds__(32) knms_sweep(const unsigned long long* __restrict__ mask, int n, int col_blocks,
                                                 unsigned char* __restrict__ keep, int* __restrict__ num_keep) {
  extern __shared__ unsigned long long removed[];
  for (int w = threadIdx.x; w < col_blocks; w += 32) removed[w] = 0;
  __syncwarp();
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    const bool dead = (removed[i >> 6] >> (i & 63)) & 1ull;     // same word for every lane: broadcast read
    if (threadIdx.x == 0) keep[i] = dead ? 0 : 1;
    if (!dead) {
      ++kept;
      const unsigned long long* row = mask + (size_t)i * col_blocks;
      for (int w = (i >> 6) + threadIdx.x; w < col_blocks; w += 32) removed[w] |= row[w];
    }
    __syncwarp();
  }
  if (threadIdx.x == 0) *num_keep = kept;
}

}  // namespace dlcfn

using namespace dlcfn;

static int roi_grid(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148LL * 16) b = 148LL * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// dtype: 0 fp32, 1 bf16.  x [N,H,W,C], rois [R,5] = (batch index, x1, y1, x2, y2) fp32, y [R,PH,PW,C]
DLCFN_API int dlcfn_roi_align_fwd(const void* x, const float* rois, void* y, int dtype, int N, int H, int W, int C, int R,
                                  int PH, int PW, float scale, int sampling_ratio, int aligned, cudaStream_t stream) {
  if (R <= 0 || C <= 0) return 0;
  RoiGeom g{N, H, W, C, R, PH, PW, scale, sampling_ratio, aligned};
  const long long total = (long long)R * PH * PW * C;
  if (dtype == 0)
    kroi_align_fwd<float><<<roi_grid(total), 256, 0, stream>>>((const float*)x, rois, (float*)y, g);
  else
    kroi_align_fwd<__nv_bfloat16><<<roi_grid(total), 256, 0, stream>>>((const __nv_bfloat16*)x, rois, (__nv_bfloat16*)y, g);
  return (int)cudaGetLastError();
}

// dy [R,PH,PW,C] (dtype), dx_f32 [N,H,W,C] fp32, must be zero on entry
DLCFN_API int dlcfn_roi_align_bwd(const void* dy, const float* rois, float* dx_f32, int dtype, int N, int H, int W, int C,
                                  int R, int PH, int PW, float scale, int sampling_ratio, int aligned, cudaStream_t stream) {
  if (R <= 0 || C <= 0) return 0;
  RoiGeom g{N, H, W, C, R, PH, PW, scale, sampling_ratio, aligned};
  const long long total = (long long)R * PH * PW * C;
  if (dtype == 0)
    kroi_align_bwd<float><<<roi_grid(total), 256, 0, stream>>>((const float*)dy, rois, dx_f32, g);
  else
    kroi_align_bwd<__nv_bfloat16><<<roi_grid(total), 256, 0, stream>>>((const __nv_bfloat16*)dy, rois, dx_f32, g);
  return (int)cudaGetLastError();
}

// boxes [n,4] fp32 sorted by descending score (16-byte aligned); mask: n * ceil(n/64) 64-bit words of scratch (zeroed here);
// keep [n] bytes, num_keep [1] int - both on the device, nothing is copied to the host.
DLCFN_API int dlcfn_nms(const float* boxes, int n, float iou_threshold, unsigned long long* mask, unsigned char* keep,
                        int* num_keep, cudaStream_t stream) {
  if (n <= 0) return (int)cudaMemsetAsync(num_keep, 0, sizeof(int), stream);
  const int col_blocks = (n + 63) / 64;
  if ((size_t)col_blocks * 8 > 200 * 1024) return (int)cudaErrorInvalidValue;      // > 1.6 M boxes: split the call
  DLCFN_CHECK(cudaMemsetAsync(mask, 0, (size_t)n * col_blocks * sizeof(unsigned long long), stream));
  knms_mask<<<dim3(col_blocks, col_blocks), 64, 0, stream>>>(boxes, n, iou_threshold, mask, col_blocks);
  DLCFN_CHECK(cudaGetLastError());
  static bool configured = false;
  if (!configured) {
    DLCFN_CHECK(cudaFuncSetAttribute(knms_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  knms_sweep<<<1, 32, (size_t)col_blocks * 8, stream>>>(mask, n, col_blocks, keep, num_keep);
  return (int)cudaGetLastError();
}
