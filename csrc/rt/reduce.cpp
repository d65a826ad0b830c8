// Host-side gradient reduction (reference K3 CommCPU::ReduceSumCPU / ReduceSumCPUImpl, MX/src/kvstore/comm.h:357-411):
// out = sum_i src_i, taking 4 inputs per pass; arrays of at least `bigarray_bound` elements are cut into 4096-element
// chunks that a small thread team walks with a static stride (the reference's `#pragma omp parallel for` over
// MXNET_KVSTORE_REDUCTION_NTHREADS threads).  `dst` may alias `srcs[0]`.
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#define DLCFN_RT_API extern "C" __attribute__((visibility("default")))

namespace {

inline float bf16_to_f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline uint16_t f_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // quiet NaN
  u += 0x7fffu + ((u >> 16) & 1u);                                            // round to nearest even
  return (uint16_t)(u >> 16);
}

template <typename T>
struct Acc;
template <>
struct Acc<float> {
  static inline float load(const float* p, size_t i) { return p[i]; }
  static inline void store(float* p, size_t i, float v) { p[i] = v; }
};
template <>
struct Acc<uint16_t> {
  static inline float load(const uint16_t* p, size_t i) { return bf16_to_f(p[i]); }
  static inline void store(uint16_t* p, size_t i, float v) { p[i] = f_to_bf16(v); }
};

template <typename T>
void reduce_range(const T* const* srcs, int nsrc, T* dst, size_t lo, size_t hi) {
  // first pass initialises from up to 4 inputs, later passes add 4 more each (comm.h:361-392)
  int i = 0;
  {
    const int m = std::min(4, nsrc);
    for (size_t e = lo; e < hi; ++e) {
      float a = Acc<T>::load(srcs[0], e);
      for (int k = 1; k < m; ++k) a += Acc<T>::load(srcs[k], e);
      Acc<T>::store(dst, e, a);
    }
    i = m;
  }
  while (i < nsrc) {
    const int m = std::min(4, nsrc - i);
    for (size_t e = lo; e < hi; ++e) {
      float a = Acc<T>::load(dst, e);
      for (int k = 0; k < m; ++k) a += Acc<T>::load(srcs[i + k], e);
      Acc<T>::store(dst, e, a);
    }
    i += m;
  }
}

template <typename T>
int reduce_sum(const T* const* srcs, int nsrc, T* dst, size_t n, int nthreads, size_t bigarray_bound) {
  if (nsrc < 1) return -1;
  if (n == 0) return 0;
  constexpr size_t kStep = 4096;
  if (n < bigarray_bound || nthreads <= 1) {
    reduce_range(srcs, nsrc, dst, 0, n);
    return 0;
  }
  const size_t nchunks = (n + kStep - 1) / kStep;
  const int team = (int)std::min<size_t>((size_t)nthreads, nchunks);
  std::vector<std::thread> th;
  th.reserve(team);
  for (int t = 0; t < team; ++t)
    th.emplace_back([=] {
      for (size_t c = (size_t)t; c < nchunks; c += (size_t)team)
        reduce_range(srcs, nsrc, dst, c * kStep, std::min(n, (c + 1) * kStep));
    });
  for (auto& x : th) x.join();
  return 0;
}

}  // namespace

// dtype: 0 = fp32, 1 = bf16 (accumulated in fp32 per pass).  Returns 0, or -1 on a bad argument.
DLCFN_RT_API int dlcfn_reduce_sum_cpu(const void* const* srcs, int nsrc, void* dst, size_t n, int dtype, int nthreads,
                                      size_t bigarray_bound) {
  if (dtype == 0) return reduce_sum<float>((const float* const*)srcs, nsrc, (float*)dst, n, nthreads, bigarray_bound);
  if (dtype == 1)
    return reduce_sum<uint16_t>((const uint16_t* const*)srcs, nsrc, (uint16_t*)dst, n, nthreads, bigarray_bound);
  return -1;
}
