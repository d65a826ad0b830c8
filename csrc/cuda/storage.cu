// Pooled device-memory manager (reference E3: GPUPooledStorageManager / GPUPooledRoundedStorageManager,
// MX/src/storage/pooled_storage_manager.h:58-66,183-206).
//
// The framework's activations live in PyTorch's caching allocator; this pool serves the runtime's own long-lived or
// recurring device buffers (kvstore staging, row-id scratch, compression words) whose sizes repeat every step.
// Same two strategies as the host pool: "Naive" = whole 4 KiB pages, exact-size reuse; "Round" = next power of two
// above a cut-off.  MXNET_GPU_MEM_POOL_RESERVE (percent of total memory that must stay free) triggers a ReleaseAll
// before a fresh cudaMalloc, exactly the reference's back-pressure rule.
#include "common.cuh"
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace {
struct DevicePool {
  std::mutex mu;
  bool round_pow2 = false;
  size_t page = 4096, cutoff = 1 << 24;
  int reserve_pct = 5;
  std::map<size_t, std::vector<void*>> free_;
  std::unordered_map<void*, size_t> live_;
  size_t bytes_live = 0, bytes_pooled = 0, n_alloc = 0, n_reuse = 0;

  size_t SizeClass(size_t n) const {
    if (n == 0) n = 1;
    if (round_pow2 && n > cutoff) {
      size_t p = cutoff;
      while (p < n) p <<= 1;
      return p;
    }
    return (n + page - 1) / page * page;
  }
  void ReleaseAllLocked() {
    for (auto& kv : free_) {
      for (void* p : kv.second) cudaFree(p);
      kv.second.clear();
    }
    bytes_pooled = 0;
  }
};
DevicePool* pool() {
  static DevicePool* p = [] {
    auto* q = new DevicePool;
    const char* t = getenv("MXNET_GPU_MEM_POOL_TYPE");
    q->round_pow2 = t && (strcmp(t, "Round") == 0 || strcmp(t, "round") == 0);
    const char* r = getenv("MXNET_GPU_MEM_POOL_RESERVE");
    if (r) q->reserve_pct = atoi(r);
    const char* c = getenv("MXNET_GPU_MEM_POOL_ROUND_LINEAR_CUTOFF");
    if (c) q->cutoff = (size_t)1 << atoi(c);
    return q;
  }();
  return p;
}
}  // namespace

DLCFN_API int dlcfn_pool_alloc(size_t nbytes, void** out) {
  (void)cudaGetLastError();
  DevicePool* P = pool();
  const size_t sc = P->SizeClass(nbytes);
  std::lock_guard<std::mutex> lk(P->mu);
  auto it = P->free_.find(sc);
  void* p = nullptr;
  if (it != P->free_.end() && !it->second.empty()) {
    p = it->second.back();
    it->second.pop_back();
    P->bytes_pooled -= sc;
    ++P->n_reuse;
  } else {
    size_t free_b = 0, total_b = 0;
    DLCFN_CHECK(cudaMemGetInfo(&free_b, &total_b));
    if (free_b < sc + total_b / 100 * (size_t)P->reserve_pct) P->ReleaseAllLocked();   // back-pressure
    cudaError_t e = cudaMalloc(&p, sc);
    if (e != cudaSuccess) {
      P->ReleaseAllLocked();
      (void)cudaGetLastError();
      e = cudaMalloc(&p, sc);
      if (e != cudaSuccess) return (int)e;
    }
    ++P->n_alloc;
  }
  P->live_[p] = sc;
  P->bytes_live += sc;
  *out = p;
  return 0;
}

DLCFN_API int dlcfn_pool_free(void* p) {
  DevicePool* P = pool();
  std::lock_guard<std::mutex> lk(P->mu);
  auto it = P->live_.find(p);
  if (it == P->live_.end()) return (int)cudaErrorInvalidValue;
  const size_t sc = it->second;
  P->live_.erase(it);
  P->bytes_live -= sc;
  P->free_[sc].push_back(p);
  P->bytes_pooled += sc;
  return 0;
}

DLCFN_API void dlcfn_pool_release_all() {
  DevicePool* P = pool();
  std::lock_guard<std::mutex> lk(P->mu);
  P->ReleaseAllLocked();
}

DLCFN_API void dlcfn_pool_set_strategy(int round_pow2, int cutoff_log2, int reserve_pct) {
  DevicePool* P = pool();
  std::lock_guard<std::mutex> lk(P->mu);
  P->round_pow2 = round_pow2 != 0;
  if (cutoff_log2 > 0) P->cutoff = (size_t)1 << cutoff_log2;
  if (reserve_pct >= 0) P->reserve_pct = reserve_pct;
}

// stats[0..4] = bytes live, bytes pooled, fresh cudaMallocs, reuses, live blocks
DLCFN_API void dlcfn_pool_stats(unsigned long long* stats) {
  DevicePool* P = pool();
  std::lock_guard<std::mutex> lk(P->mu);
  stats[0] = P->bytes_live; stats[1] = P->bytes_pooled; stats[2] = P->n_alloc; stats[3] = P->n_reuse;
  stats[4] = P->live_.size();
}
