// Host storage managers of the native runtime (reference E3: MX/src/storage/ - pooled_storage_manager.h:58-206,
// cpu_shared_storage_manager.h, pinned_memory_storage.h).
//
//  * PooledHost: page-aligned host blocks recycled by size class.  Two strategies as in the reference
//    (MXNET_GPU_MEM_POOL_TYPE): "Naive" rounds a request up to whole pages and reuses exact sizes, "Round" rounds up
//    to the next power of two above a cut-off so that slightly different sizes share blocks.  The python side pins
//    the blocks with cudaHostRegister once (pinning is the expensive part, which is why staging buffers are pooled).
//  * Shm: named POSIX shared-memory segments with a creator/attacher protocol - the data-loader workers hand decoded
//    batches to the trainer process without a copy through a pipe (CPUSharedStorageManager).
#include <cerrno>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <map>
#include <mutex>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <unordered_map>
#include <vector>

#define DLCFN_RT_API extern "C" __attribute__((visibility("default")))

namespace {

struct PooledHost {
  std::mutex mu;
  bool round_pow2 = false;
  size_t page = 4096, cutoff = 1 << 12;
  std::map<size_t, std::vector<void*>> free_;          // size class -> blocks
  std::unordered_map<void*, size_t> live_;             // block -> size class
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
  void* Alloc(size_t n) {
    const size_t sc = SizeClass(n);
    std::lock_guard<std::mutex> lk(mu);
    auto it = free_.find(sc);
    void* p = nullptr;
    if (it != free_.end() && !it->second.empty()) {
      p = it->second.back();
      it->second.pop_back();
      bytes_pooled -= sc;
      ++n_reuse;
    } else {
      if (posix_memalign(&p, page, sc) != 0) return nullptr;
      ++n_alloc;
    }
    live_[p] = sc;
    bytes_live += sc;
    return p;
  }
  int Free(void* p) {
    std::lock_guard<std::mutex> lk(mu);
    auto it = live_.find(p);
    if (it == live_.end()) return -1;
    const size_t sc = it->second;
    live_.erase(it);
    bytes_live -= sc;
    free_[sc].push_back(p);
    bytes_pooled += sc;
    return 0;
  }
  // returns the number of blocks handed back; `out` (may be null) receives up to cap pointers so that the caller
  // can unpin them before they are released
  size_t Drain(void** out, size_t cap, bool release) {
    std::lock_guard<std::mutex> lk(mu);
    size_t k = 0;
    for (auto& kv : free_) {
      for (void* p : kv.second) {
        if (out && k < cap) out[k] = p;
        ++k;
        if (release) free(p);
      }
      if (release) kv.second.clear();
    }
    if (release) bytes_pooled = 0;
    return k;
  }
};

PooledHost* g_pool() {
  static PooledHost* p = [] {
    auto* q = new PooledHost;
    const char* t = getenv("MXNET_GPU_MEM_POOL_TYPE");
    q->round_pow2 = t && (strcmp(t, "Round") == 0 || strcmp(t, "round") == 0);
    const char* c = getenv("MXNET_GPU_MEM_POOL_ROUND_LINEAR_CUTOFF");
    if (c) q->cutoff = (size_t)1 << atoi(c);
    return q;
  }();
  return p;
}

struct ShmSeg {
  void* ptr;
  size_t size;
  std::string name;
  bool owner;
};
std::mutex g_shm_mu;
std::unordered_map<void*, ShmSeg> g_shm;

}  // namespace

DLCFN_RT_API void* dlcfn_storage_alloc(size_t nbytes) { return g_pool()->Alloc(nbytes); }
DLCFN_RT_API int dlcfn_storage_free(void* p) { return g_pool()->Free(p); }
DLCFN_RT_API size_t dlcfn_storage_size_class(size_t nbytes) { return g_pool()->SizeClass(nbytes); }
DLCFN_RT_API void dlcfn_storage_set_strategy(int round_pow2, int cutoff_log2) {
  PooledHost* p = g_pool();
  std::lock_guard<std::mutex> lk(p->mu);
  p->round_pow2 = round_pow2 != 0;
  if (cutoff_log2 > 0) p->cutoff = (size_t)1 << cutoff_log2;
}
// stats[0..4] = bytes live, bytes pooled, fresh allocations, reuses, live blocks
DLCFN_RT_API void dlcfn_storage_stats(uint64_t* stats) {
  PooledHost* p = g_pool();
  std::lock_guard<std::mutex> lk(p->mu);
  stats[0] = p->bytes_live; stats[1] = p->bytes_pooled; stats[2] = p->n_alloc; stats[3] = p->n_reuse;
  stats[4] = p->live_.size();
}
DLCFN_RT_API size_t dlcfn_storage_list_pooled(void** out, size_t cap) { return g_pool()->Drain(out, cap, false); }
DLCFN_RT_API size_t dlcfn_storage_release_all() { return g_pool()->Drain(nullptr, 0, true); }

// create != 0: make a fresh segment of `size` bytes (fails if it exists); else attach to an existing one
// (size may be 0 = take the segment's size).  Returns the mapping or null (errno set).
DLCFN_RT_API void* dlcfn_shm_open(const char* name, size_t size, int create, size_t* out_size) {
  int fd = create ? shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600) : shm_open(name, O_RDWR, 0600);
  if (fd < 0) return nullptr;
  if (create) {
    if (ftruncate(fd, (off_t)size) != 0) { close(fd); shm_unlink(name); return nullptr; }
  } else {
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return nullptr; }
    if (size == 0) size = (size_t)st.st_size;
    if ((size_t)st.st_size < size) { close(fd); errno = EINVAL; return nullptr; }
  }
  void* p = mmap(nullptr, size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { if (create) shm_unlink(name); return nullptr; }
  if (out_size) *out_size = size;
  std::lock_guard<std::mutex> lk(g_shm_mu);
  g_shm[p] = ShmSeg{p, size, name, create != 0};
  return p;
}
// unmaps; the creator also unlinks the name (attachers keep their mappings alive until they close)
DLCFN_RT_API int dlcfn_shm_close(void* p) {
  ShmSeg seg;
  {
    std::lock_guard<std::mutex> lk(g_shm_mu);
    auto it = g_shm.find(p);
    if (it == g_shm.end()) return -1;
    seg = it->second;
    g_shm.erase(it);
  }
  munmap(seg.ptr, seg.size);
  if (seg.owner) shm_unlink(seg.name.c_str());
  return 0;
}
