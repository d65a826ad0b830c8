// Native RecordIO + image-record batch pipeline (reference E7: MX/src/io/ - iter_image_recordio_2.cc, image_aug_default.cc,
// iter_prefetcher.h, iter_batchloader.h; dmlc-core recordio.h / threadediter.h).
//
//   RecReader       read-only mmap of a .rec file + its .idx (key \t offset); random access without locks, records
//                   that the dmlc writer split at embedded magic words (cflag 1/2/3) are re-joined on read
//   RecWriter       append records, write the index on close
//   ImageBatchIter  a team of decoder threads turns records (IRHeader + uint32 h,w,c + raw uint8 HWC pixels) into
//                   fixed-shape uint8 NHWC batches + float labels: crop (center | random), zero pad, random mirror,
//                   deterministic per-sample randomness (epoch, index) so results do not depend on thread timing;
//                   batches are delivered IN ORDER through a ring of slots that live in the pooled host allocator
//                   (storage.cpp) - the python side pins each slot once and issues asynchronous H2D copies from it.
//                   Sharding by (part_index, num_parts) and per-epoch shuffling as in the reference iterator.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <mutex>
#include <random>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <thread>
#include <unistd.h>
#include <vector>

#define DLCFN_RT_API extern "C" __attribute__((visibility("default")))

extern "C" void* dlcfn_storage_alloc(size_t nbytes);
extern "C" int dlcfn_storage_free(void* p);

namespace {

constexpr uint32_t kMagic = 0xCED7230Au;

struct RecReader {
  int fd = -1;
  const uint8_t* base = nullptr;
  size_t size = 0;
  std::vector<int64_t> keys;
  std::vector<uint64_t> offs;

  ~RecReader() {
    if (base) munmap((void*)base, size);
    if (fd >= 0) close(fd);
  }
  // payload of the record starting at `off`; multi-part records are joined into `scratch`
  bool Get(uint64_t off, const uint8_t** p, size_t* n, std::string* scratch) const {
    if (off + 8 > size) return false;
    uint32_t magic, lrec;
    memcpy(&magic, base + off, 4);
    memcpy(&lrec, base + off + 4, 4);
    if (magic != kMagic) return false;
    const uint32_t cflag = lrec >> 29, len = lrec & ((1u << 29) - 1);
    if (off + 8 + len > size) return false;
    if (cflag == 0) {
      *p = base + off + 8;
      *n = len;
      return true;
    }
    // cflag 1 = first part, 2 = middle, 3 = last; the writer dropped one magic word between parts
    scratch->clear();
    uint64_t cur = off;
    for (;;) {
      memcpy(&magic, base + cur, 4);
      memcpy(&lrec, base + cur + 4, 4);
      if (magic != kMagic) return false;
      const uint32_t cf = lrec >> 29, l = lrec & ((1u << 29) - 1);
      if (cur + 8 + l > size) return false;
      scratch->append((const char*)base + cur + 8, l);
      cur += 8 + ((l + 3u) & ~3u);
      if (cf == 3) break;
      const uint32_t m = kMagic;
      scratch->append((const char*)&m, 4);
      if (cur + 8 > size) return false;
    }
    *p = (const uint8_t*)scratch->data();
    *n = scratch->size();
    return true;
  }
};

struct RecWriter {
  FILE* f = nullptr;
  std::string idx_path;
  std::vector<std::pair<int64_t, uint64_t>> index;
  uint64_t pos = 0;
};

inline uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct ImageBatchIter {
  RecReader* rec = nullptr;
  int batch = 0, H = 0, W = 0, C = 0, label_width = 1;
  bool shuffle = false, rand_crop = false, rand_mirror = false, round_batch = true;
  uint64_t seed = 0;
  int nthreads = 1, nslots = 4, keep = 2;
  std::vector<size_t> shard;   // indices into rec->offs owned by this part
  std::vector<size_t> order;   // this epoch's order
  long epoch = 0;
  long nbatches = 0;

  struct Slot {
    uint8_t* data = nullptr;
    float* label = nullptr;
    long batch_id = -1;   // which batch is stored (valid when full)
    bool full = false;
    int nvalid = 0;
  };
  std::vector<Slot> slots;
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<long> next_batch{0};
  long consumed = 0;      // batches handed to the consumer
  long released = 0;      // batches whose slots may be overwritten: [0, released)
  bool stop = false;
  long errors = 0;
  std::vector<std::thread> team;

  size_t DataBytes() const { return (size_t)batch * H * W * C; }

  void BuildOrder() {
    order = shard;
    if (shuffle) {
      std::mt19937_64 g(seed + 0x51ED2701ull * (uint64_t)(epoch + 1));
      std::shuffle(order.begin(), order.end(), g);
    }
    const long n = (long)order.size();
    nbatches = round_batch ? (n + batch - 1) / batch : n / batch;
  }

  void FillSample(size_t rec_index, long sample_global, uint8_t* dst, float* lab, std::string* scratch) {
    const uint8_t* p;
    size_t n;
    memset(dst, 0, (size_t)H * W * C);
    for (int k = 0; k < label_width; ++k) lab[k] = 0.f;
    if (!rec->Get(rec->offs[rec_index], &p, &n, scratch) || n < 24 + 12) {
      std::lock_guard<std::mutex> lk(mu);
      ++errors;
      return;
    }
    uint32_t flag;
    float label;
    memcpy(&flag, p, 4);
    memcpy(&label, p + 4, 4);
    size_t off = 24;
    if (flag == 0) {
      lab[0] = label;
    } else {
      for (uint32_t k = 0; k < flag && off + 4 <= n; ++k, off += 4)
        if ((int)k < label_width) memcpy(&lab[k], p + off, 4);
    }
    if (off + 12 > n) return;
    uint32_t sh, sw, sc;
    memcpy(&sh, p + off, 4);
    memcpy(&sw, p + off + 4, 4);
    memcpy(&sc, p + off + 8, 4);
    off += 12;
    if ((size_t)sh * sw * sc > n - off) {
      std::lock_guard<std::mutex> lk(mu);
      ++errors;
      return;
    }
    const uint8_t* img = p + off;
    const uint64_t r = splitmix(seed ^ splitmix((uint64_t)epoch * 0x100000001B3ull + (uint64_t)sample_global));
    const int ch = std::min<int>(H, (int)sh), cw = std::min<int>(W, (int)sw), cc = std::min<int>(C, (int)sc);
    int y0 = ((int)sh - ch) / 2, x0 = ((int)sw - cw) / 2;      // source offset: center crop
    if (rand_crop) {
      y0 = (int)sh > ch ? (int)((r >> 8) % (uint64_t)((int)sh - ch + 1)) : 0;
      x0 = (int)sw > cw ? (int)((r >> 24) % (uint64_t)((int)sw - cw + 1)) : 0;
    }
    const int dy0 = (H - ch) / 2, dx0 = (W - cw) / 2;          // destination offset when the source is smaller (zero pad)
    const bool mirror = rand_mirror && ((r >> 4) & 1u);
    for (int y = 0; y < ch; ++y) {
      const uint8_t* srow = img + ((size_t)(y0 + y) * sw + x0) * sc;
      uint8_t* drow = dst + ((size_t)(dy0 + y) * W + dx0) * C;
      if (!mirror && cc == (int)sc && cc == C) {
        memcpy(drow, srow, (size_t)cw * C);
      } else {
        for (int x = 0; x < cw; ++x) {
          const uint8_t* s = srow + (size_t)(mirror ? cw - 1 - x : x) * sc;
          uint8_t* d = drow + (size_t)x * C;
          for (int k = 0; k < cc; ++k) d[k] = s[k];
        }
      }
    }
  }

  void Worker() {
    std::string scratch;
    for (;;) {
      const long b = next_batch.fetch_add(1);
      if (b >= nbatches) return;
      Slot& s = slots[(size_t)(b % nslots)];
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || b < released + nslots; });   // slot b % nslots is free again
        if (stop) return;
      }
      const long n = (long)order.size();
      int nvalid = 0;
      for (int j = 0; j < batch; ++j) {
        long idx = b * batch + j;
        if (idx < n) ++nvalid;
        else idx %= n;                                    // round_batch: wrap to the start of the epoch
        FillSample(order[(size_t)idx], idx, s.data + (size_t)j * H * W * C, s.label + (size_t)j * label_width, &scratch);
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        s.batch_id = b;
        s.nvalid = nvalid;
        s.full = true;
      }
      cv.notify_all();
    }
  }

  void Start() {
    next_batch = 0;
    consumed = released = 0;
    stop = false;
    for (auto& s : slots) { s.full = false; s.batch_id = -1; }
    BuildOrder();
    for (int t = 0; t < nthreads; ++t) team.emplace_back([this] { Worker(); });
  }

  void Stop() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    for (auto& t : team) t.join();
    team.clear();
  }

  // 1 = batch delivered, 0 = end of epoch
  int Next(uint8_t** data, float** label, int* nvalid) {
    std::unique_lock<std::mutex> lk(mu);
    // buffers handed out more than `keep` calls ago go back to the decoders
    if (consumed - keep > released) {
      for (long b = released; b < consumed - keep; ++b) slots[(size_t)(b % nslots)].full = false;
      released = consumed - keep;
      cv.notify_all();
    }
    if (consumed >= nbatches) return 0;
    const long b = consumed;
    Slot& s = slots[(size_t)(b % nslots)];
    cv.wait(lk, [&] { return s.full && s.batch_id == b; });
    *data = s.data;
    *label = s.label;
    *nvalid = s.nvalid;
    ++consumed;
    return 1;
  }
};

}  // namespace

// ------------------------------------------------------------------------------------------------ reader / writer
DLCFN_RT_API void* dlcfn_rec_open(const char* rec_path, const char* idx_path) {
  auto* r = new RecReader;
  r->fd = open(rec_path, O_RDONLY);
  struct stat st;
  if (r->fd < 0 || fstat(r->fd, &st) != 0) { delete r; return nullptr; }
  r->size = (size_t)st.st_size;
  if (r->size) {
    void* m = mmap(nullptr, r->size, PROT_READ, MAP_SHARED, r->fd, 0);
    if (m == MAP_FAILED) { delete r; return nullptr; }
    r->base = (const uint8_t*)m;
  }
  FILE* f = idx_path ? fopen(idx_path, "r") : nullptr;
  if (f) {
    long long k, pos;
    while (fscanf(f, "%lld\t%lld", &k, &pos) == 2) { r->keys.push_back(k); r->offs.push_back((uint64_t)pos); }
    fclose(f);
  } else {   // no index: scan the file once
    uint64_t off = 0;
    int64_t k = 0;
    while (off + 8 <= r->size) {
      uint32_t magic, lrec;
      memcpy(&magic, r->base + off, 4);
      memcpy(&lrec, r->base + off + 4, 4);
      if (magic != kMagic) break;
      const uint32_t cf = lrec >> 29, l = lrec & ((1u << 29) - 1);
      if (cf == 0 || cf == 1) { r->keys.push_back(k++); r->offs.push_back(off); }
      off += 8 + ((l + 3u) & ~3u);
    }
  }
  return r;
}
DLCFN_RT_API long dlcfn_rec_num(void* h) { return (long)((RecReader*)h)->offs.size(); }
DLCFN_RT_API long long dlcfn_rec_key(void* h, long i) { return ((RecReader*)h)->keys[(size_t)i]; }
// copies record i into out (cap bytes); returns its length, or -1 on a framing error / too small a buffer
DLCFN_RT_API long dlcfn_rec_read(void* h, long i, void* out, size_t cap) {
  auto* r = (RecReader*)h;
  if (i < 0 || (size_t)i >= r->offs.size()) return -1;
  const uint8_t* p;
  size_t n;
  std::string scratch;
  if (!r->Get(r->offs[(size_t)i], &p, &n, &scratch)) return -1;
  if (out == nullptr) return (long)n;
  if (n > cap) return -1;
  memcpy(out, p, n);
  return (long)n;
}
DLCFN_RT_API void dlcfn_rec_close(void* h) { delete (RecReader*)h; }

DLCFN_RT_API void* dlcfn_rec_writer_open(const char* rec_path, const char* idx_path) {
  auto* w = new RecWriter;
  w->f = fopen(rec_path, "wb");
  if (!w->f) { delete w; return nullptr; }
  w->idx_path = idx_path ? idx_path : "";
  return w;
}
// dmlc framing incl. the split at embedded magic words: a payload containing the (4-byte aligned) magic is written as
// parts with cflag 1 / 2 / 3 and the magic word itself dropped (recordio.cc:WriteRecord)
DLCFN_RT_API int dlcfn_rec_write(void* h, long long key, const void* buf, size_t len) {
  auto* w = (RecWriter*)h;
  if (len >= (1u << 29)) return -1;
  w->index.emplace_back((int64_t)key, w->pos);
  const uint8_t* p = (const uint8_t*)buf;
  const uint32_t magic = kMagic;
  size_t dptr = 0;
  const size_t lower = len & ~(size_t)3;
  auto emit = [&](uint32_t cflag, const uint8_t* data, uint32_t l) {
    const uint32_t lrec = (cflag << 29) | l;
    fwrite(&magic, 4, 1, w->f);
    fwrite(&lrec, 4, 1, w->f);
    if (l) fwrite(data, 1, l, w->f);
    const uint32_t pad = (4 - (l & 3u)) & 3u, zero = 0;
    if (pad) fwrite(&zero, 1, pad, w->f);
    w->pos += 8 + l + pad;
  };
  for (size_t i = 0; i < lower; i += 4) {
    uint32_t v;
    memcpy(&v, p + i, 4);
    if (v == kMagic) {
      emit(dptr == 0 ? 1u : 2u, p + dptr, (uint32_t)(i - dptr));
      dptr = i + 4;
    }
  }
  emit(dptr != 0 ? 3u : 0u, p + dptr, (uint32_t)(len - dptr));
  return 0;
}
DLCFN_RT_API int dlcfn_rec_writer_close(void* h) {
  auto* w = (RecWriter*)h;
  int rc = 0;
  if (w->f) rc = fclose(w->f);
  if (!w->idx_path.empty()) {
    FILE* f = fopen(w->idx_path.c_str(), "w");
    if (!f) rc = -1;
    else {
      for (auto& kv : w->index) fprintf(f, "%lld\t%llu\n", (long long)kv.first, (unsigned long long)kv.second);
      fclose(f);
    }
  }
  delete w;
  return rc;
}

// --------------------------------------------------------------------------------------------- image batch iterator
DLCFN_RT_API void* dlcfn_imgiter_create(void* rec, int batch, int H, int W, int C, int label_width, int shuffle,
                                        unsigned long long seed, int part_index, int num_parts, int nthreads, int prefetch,
                                        int rand_crop, int rand_mirror, int round_batch, int keep) {
  auto* r = (RecReader*)rec;
  if (!r || batch < 1 || H < 1 || W < 1 || C < 1 || num_parts < 1 || part_index < 0 || part_index >= num_parts) return nullptr;
  auto* it = new ImageBatchIter;
  it->rec = r;
  it->batch = batch; it->H = H; it->W = W; it->C = C;
  it->label_width = std::max(1, label_width);
  it->shuffle = shuffle != 0; it->rand_crop = rand_crop != 0; it->rand_mirror = rand_mirror != 0;
  it->round_batch = round_batch != 0;
  it->seed = seed;
  it->nthreads = std::max(1, nthreads);
  it->keep = std::max(1, keep);
  it->nslots = std::max(1, prefetch) + it->keep + 1;
  for (size_t i = (size_t)part_index; i < r->offs.size(); i += (size_t)num_parts) it->shard.push_back(i);
  if (it->shard.empty()) { delete it; return nullptr; }
  it->slots.resize((size_t)it->nslots);
  for (auto& s : it->slots) {
    s.data = (uint8_t*)dlcfn_storage_alloc(it->DataBytes());
    s.label = (float*)dlcfn_storage_alloc((size_t)batch * it->label_width * sizeof(float));
    if (!s.data || !s.label) { delete it; return nullptr; }
  }
  it->epoch = 0;
  it->Start();
  return it;
}
DLCFN_RT_API long dlcfn_imgiter_num_batches(void* h) { return ((ImageBatchIter*)h)->nbatches; }
DLCFN_RT_API long dlcfn_imgiter_num_samples(void* h) { return (long)((ImageBatchIter*)h)->shard.size(); }
DLCFN_RT_API long dlcfn_imgiter_errors(void* h) { return ((ImageBatchIter*)h)->errors; }
DLCFN_RT_API int dlcfn_imgiter_num_slots(void* h) { return ((ImageBatchIter*)h)->nslots; }
// the ring's buffers (stable for the iterator's lifetime): lets the caller page-lock them once
DLCFN_RT_API void* dlcfn_imgiter_slot_data(void* h, int i) { return ((ImageBatchIter*)h)->slots[(size_t)i].data; }
DLCFN_RT_API int dlcfn_imgiter_next(void* h, void** data, void** label, int* nvalid) {
  uint8_t* d = nullptr;
  float* l = nullptr;
  const int rc = ((ImageBatchIter*)h)->Next(&d, &l, nvalid);
  *data = d;
  *label = l;
  return rc;
}
// start the next epoch (new shuffle); outstanding buffers become invalid
DLCFN_RT_API void dlcfn_imgiter_reset(void* h) {
  auto* it = (ImageBatchIter*)h;
  it->Stop();
  ++it->epoch;
  it->Start();
}
DLCFN_RT_API void dlcfn_imgiter_destroy(void* h) {
  auto* it = (ImageBatchIter*)h;
  it->Stop();
  for (auto& s : it->slots) {
    if (s.data) dlcfn_storage_free(s.data);
    if (s.label) dlcfn_storage_free(s.label);
  }
  delete it;
}
