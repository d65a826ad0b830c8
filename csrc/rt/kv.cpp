// KVWorker / KVServer: the KVStoreDist (worker) and KVStoreDistServer (server) logic over Node.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>

#include "ps.h"
#include <memory>

namespace dlcfn {

void UnpairCmd(int z, int* a, int* b) {
  int w = (int)std::floor((std::sqrt(8.0 * z + 1) - 1) / 2);
  int t = (w * w + w) / 2;
  *b = z - t;
  *a = w - *b;
}

size_t DTypeSize(int dtype) { return dtype == kF32 ? 4 : 2; }

static inline float Bf16ToF(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
static inline uint16_t FToBf16(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);  // NaN
  u += 0x7fffu + ((u >> 16) & 1u);                                          // round to nearest even
  return (uint16_t)(u >> 16);
}
static inline float F16ToF(uint16_t h) {
  uint32_t s = (h & 0x8000u) << 16, e = (h >> 10) & 0x1f, m = h & 0x3ff, u;
  if (e == 0) {
    if (m == 0) u = s;
    else { int k = 0; while (!(m & 0x400)) { m <<= 1; ++k; } m &= 0x3ff; u = s | ((uint32_t)(127 - 15 - k + 1) << 23) | (m << 13); }
  } else if (e == 31) u = s | 0x7f800000u | (m << 13);
  else u = s | ((e + 112) << 23) | (m << 13);
  float f; memcpy(&f, &u, 4); return f;
}
static inline uint16_t FToF16(float f) {
  uint32_t u; memcpy(&u, &f, 4);
  uint32_t s = (u >> 16) & 0x8000u; int e = (int)((u >> 23) & 0xff) - 127 + 15; uint32_t m = u & 0x7fffffu;
  if (e >= 31) return (uint16_t)(s | 0x7c00u | (((u & 0x7fffffffu) > 0x7f800000u) ? 0x200 : 0));
  if (e <= 0) {
    if (e < -10) return (uint16_t)s;
    m |= 0x800000u; int shift = 14 - e;
    uint32_t r = m >> shift; if ((m >> (shift - 1)) & 1) ++r; return (uint16_t)(s | r);
  }
  uint32_t r = (uint32_t)(e << 10) | (m >> 13);
  if (m & 0x1000u) ++r;
  return (uint16_t)(s | r);
}

void ToFloat(const void* src, int dtype, size_t n, float* dst) {
  if (dtype == kF32) { memcpy(dst, src, n * 4); return; }
  const uint16_t* p = (const uint16_t*)src;
  if (dtype == kBF16) for (size_t i = 0; i < n; ++i) dst[i] = Bf16ToF(p[i]);
  else for (size_t i = 0; i < n; ++i) dst[i] = F16ToF(p[i]);
}
// dst[i] += src[i] with the conversion fused in (no temporary)
void AccumulateFloat(const void* src, int dtype, size_t n, float* dst) {
  if (dtype == kF32) {
    const char* p = (const char*)src;
    for (size_t i = 0; i < n; ++i) { float v; memcpy(&v, p + 4 * i, 4); dst[i] += v; }
    return;
  }
  const uint16_t* p = (const uint16_t*)src;
  if (dtype == kBF16) for (size_t i = 0; i < n; ++i) dst[i] += Bf16ToF(p[i]);
  else for (size_t i = 0; i < n; ++i) dst[i] += F16ToF(p[i]);
}
void FromFloat(const float* src, int dtype, size_t n, void* dst) {
  if (dtype == kF32) { memcpy(dst, src, n * 4); return; }
  uint16_t* p = (uint16_t*)dst;
  if (dtype == kBF16) for (size_t i = 0; i < n; ++i) p[i] = FToBf16(src[i]);
  else for (size_t i = 0; i < n; ++i) p[i] = FToF16(src[i]);
}

// ======================================================================== worker
KVWorker::KVWorker(Node* node) : node_(node) {
  const char* b = getenv("MXNET_KVSTORE_BIGARRAY_BOUND");
  if (b) bigarray_bound = (size_t)atoll(b);
}

void KVWorker::SetKeyUnit(int key, size_t unit) {
  std::lock_guard<std::mutex> lk(unit_mu_);
  key_unit_[key] = unit;
}

std::vector<KVWorker::Part> KVWorker::Partition(int key, size_t n) const {
  const int S = node_->num_servers();
  std::vector<Part> parts;
  if (n < bigarray_bound || S == 1) {
    // small array: whole key on one server, spread by a multiplicative hash (kvstore_dist.h:547-556)
    int server = (int)(((long long)key * 9973) % S);
    auto range = node_->ServerKeyRange(server);
    parts.push_back(Part{server, range.first + (uint64_t)key, 0, n});
  } else {
    // big array: contiguous even split over all servers (kvstore_dist.h:557-570), boundaries rounded to the key's unit
    size_t unit = 16;
    {
      std::lock_guard<std::mutex> lk(unit_mu_);
      auto it = key_unit_.find(key);
      if (it != key_unit_.end() && it->second > 0) unit = it->second;
    }
    size_t off = 0;
    for (int s = 0; s < S; ++s) {
      size_t end = (size_t)std::llround((double)n / S * (s + 1));
      end = (end + unit / 2) / unit * unit;
      if (s == S - 1 || end > n) end = n;
      if (end < off) end = off;
      const size_t len = end - off;
      auto range = node_->ServerKeyRange(s);
      if (len > 0) parts.push_back(Part{s, range.first + (uint64_t)key, off, len});
      off = end;
    }
  }
  return parts;
}

int KVWorker::Push(int key, const void* data, size_t n, int dtype, int req_type, const int64_t* row_ids,
                   size_t num_rows, size_t row_len) {
  const int ts = node_->NextTimestamp();
  const size_t es = DTypeSize(dtype);
  std::vector<Part> parts;
  if (req_type == kRowSparsePushPull) {
    // rows are routed to the server that owns them under the dense split of the full tensor
    parts = Partition(key, n);  // n = total elements of the dense tensor
  } else if (req_type == kCompressedPushPull) {
    parts = Partition(key, n);  // n = ORIGINAL element count; data = packed 2-bit words (16 values / word)
  } else {
    parts = Partition(key, n);
  }
  {
    std::lock_guard<std::mutex> lk(mu_);
    Track& t = track_[ts];
    t.expected = (int)parts.size();
  }
  for (const Part& p : parts) {
    Message m;
    m.recver = ServerRankToId(p.server);
    m.request = true; m.push = true; m.timestamp = ts;
    m.head = PairCmd(req_type, dtype);
    m.data.emplace_back(reinterpret_cast<const char*>(&p.ps_key), 8);
    if (req_type == kRowSparsePushPull) {
      // select the rows that fall inside [offset, offset+len) (offsets are multiples of row_len for 2-D splits)
      std::string ids, vals;
      for (size_t r = 0; r < num_rows; ++r) {
        size_t begin = (size_t)row_ids[r] * row_len;
        if (begin >= p.offset && begin < p.offset + p.len) {
          int64_t local = (int64_t)((begin - p.offset) / row_len);
          ids.append(reinterpret_cast<const char*>(&local), 8);
          vals.append(reinterpret_cast<const char*>(data) + r * row_len * es, row_len * es);
        }
      }
      uint64_t meta[2] = {p.len, row_len};
      m.data.push_back(std::move(vals));
      m.data.push_back(std::move(ids));
      m.data.emplace_back(reinterpret_cast<const char*>(meta), 16);
    } else if (req_type == kCompressedPushPull) {
      // word-aligned slice of the packed buffer: Partition() keeps every boundary on a multiple of 16 elements
      size_t w0 = p.offset / 16, w1 = (p.offset + p.len + 15) / 16;
      m.data.emplace_back(reinterpret_cast<const char*>(data) + w0 * 4, (w1 - w0) * 4);
      uint64_t meta[1] = {p.len};
      m.data.emplace_back(reinterpret_cast<const char*>(meta), 8);
    } else {
      // dense values leave straight from the caller's buffer (borrowed blob): no copy into the message, none into the frame
      m.blob = reinterpret_cast<const char*>(data) + p.offset * es;
      m.blob_len = p.len * es;
    }
    node_->Send(m);
  }
  return ts;
}

int KVWorker::Pull(int key, void* out, size_t n, int dtype, int req_type, const int64_t* row_ids, size_t num_rows,
                   size_t row_len) {
  const int ts = node_->NextTimestamp();
  std::vector<Part> parts = Partition(key, n);
  {
    std::lock_guard<std::mutex> lk(mu_);
    Track& t = track_[ts];
    t.expected = (int)parts.size();
    t.out = out; t.dtype = dtype; t.parts = parts;
    t.row_sparse = req_type == kRowSparsePushPull; t.row_len = row_len;
  }
  for (const Part& p : parts) {
    Message m;
    m.recver = ServerRankToId(p.server);
    m.request = true; m.push = false; m.timestamp = ts;
    m.head = PairCmd(req_type, dtype);
    m.data.emplace_back(reinterpret_cast<const char*>(&p.ps_key), 8);
    if (req_type == kRowSparsePushPull) {
      std::string ids, pos;
      for (size_t r = 0; r < num_rows; ++r) {
        size_t begin = (size_t)row_ids[r] * row_len;
        if (begin >= p.offset && begin < p.offset + p.len) {
          int64_t local = (int64_t)((begin - p.offset) / row_len), where = (int64_t)r;
          ids.append(reinterpret_cast<const char*>(&local), 8);
          pos.append(reinterpret_cast<const char*>(&where), 8);
        }
      }
      uint64_t meta[2] = {p.len, row_len};
      m.data.push_back(std::move(ids));
      m.data.push_back(std::move(pos));  // echoed back so the reply can be scattered
      m.data.emplace_back(reinterpret_cast<const char*>(meta), 16);
    }
    node_->Send(m);
  }
  return ts;
}

void KVWorker::OnMessage(Message&& m) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = track_.find(m.timestamp);
  if (it == track_.end()) return;
  Track& t = it->second;
  if (!m.push && t.out != nullptr && !m.data.empty()) {
    const int server = IdToRank(m.sender);
    const size_t es = DTypeSize(t.dtype);
    for (const Part& p : t.parts) {
      if (p.server != server) continue;
      if (t.row_sparse) {
        // data[0] = rows, data[1] = destination positions
        const std::string& rows = m.data[0];
        const std::string& pos = m.data.size() > 1 ? m.data[1] : rows;
        size_t nr = pos.size() / 8;
        for (size_t r = 0; r < nr; ++r) {
          int64_t where; memcpy(&where, pos.data() + r * 8, 8);
          memcpy(reinterpret_cast<char*>(t.out) + (size_t)where * t.row_len * es, rows.data() + r * t.row_len * es,
                 t.row_len * es);
        }
      } else {
        memcpy(reinterpret_cast<char*>(t.out) + p.offset * es, m.data[0].data(),
               std::min(m.data[0].size(), p.len * es));
      }
    }
  }
  ++t.received;
  if (t.received >= t.expected) cv_.notify_all();
}

void KVWorker::Wait(int ts) {
  std::unique_lock<std::mutex> lk(mu_);
  cv_.wait(lk, [&] {
    auto it = track_.find(ts);
    return it == track_.end() || it->second.received >= it->second.expected;
  });
  track_.erase(ts);
}

void KVWorker::SendCommand(int head, const std::string& body) {
  const int ts = node_->NextTimestamp();
  {
    std::lock_guard<std::mutex> lk(mu_);
    track_[ts].expected = node_->num_servers();
  }
  for (int s = 0; s < node_->num_servers(); ++s) {
    Message m;
    m.recver = ServerRankToId(s);
    m.request = true; m.simple_app = true; m.timestamp = ts; m.head = head; m.body = body;
    node_->Send(m);
  }
  Wait(ts);
}

// ======================================================================== server
KVServer::KVServer(Node* node) : node_(node) {}

void KVServer::Run() {
  main_thread_ = std::this_thread::get_id();
  running_ = true;
  std::unique_lock<std::mutex> lk(exec_mu_);
  for (;;) {
    exec_cv_.wait(lk, [&] { return !exec_q_.empty() || exec_stop_.load(); });   // never takes mu_ (lock order)
    while (!exec_q_.empty()) {
      ExecTask* t = exec_q_.front();
      exec_q_.pop_front();
      lk.unlock();
      t->fn();
      lk.lock();
      if (t->detached) delete t;
      else { t->done = true; exec_done_cv_.notify_all(); }
    }
    if (exec_stop_.load() && exec_q_.empty()) break;
  }
  running_ = false;
}

void KVServer::ExecOnMain(const std::function<void()>& fn) {
  if (!running_ || std::this_thread::get_id() == main_thread_) { fn(); return; }
  ExecTask t;
  t.fn = fn;
  std::unique_lock<std::mutex> lk(exec_mu_);
  exec_q_.push_back(&t);
  exec_cv_.notify_all();
  exec_done_cv_.wait(lk, [&] { return t.done; });
}

void KVServer::PostToMain(std::function<void()> fn) {
  ExecTask* t = new ExecTask;
  t->fn = std::move(fn);
  t->detached = true;
  std::lock_guard<std::mutex> lk(exec_mu_);
  exec_q_.push_back(t);
  exec_cv_.notify_all();
}

void KVServer::Respond(const Message& req, const std::string* payload) {
  Message r;
  r.recver = req.sender;
  r.request = false; r.push = req.push; r.simple_app = req.simple_app;
  r.timestamp = req.timestamp; r.head = req.head;
  if (payload) { r.blob = payload->data(); r.blob_len = payload->size(); }   // borrowed: Send is synchronous
  node_->Send(r);
}

void KVServer::OnMessage(Message&& m) {
  if (m.simple_app) HandleCommand(std::move(m));
  else HandleData(std::move(m));
}

static bool ParseOpt(const std::string& body, OptimizerConfig* o) {
  // "kind=sgd_mom;lr=0.1;wd=1e-4;momentum=0.9;beta1=..;beta2=..;eps=..;rescale=..;clip=.."
  size_t pos = 0;
  while (pos < body.size()) {
    size_t semi = body.find(';', pos);
    if (semi == std::string::npos) semi = body.size();
    std::string kv = body.substr(pos, semi - pos);
    pos = semi + 1;
    size_t eq = kv.find('=');
    if (eq == std::string::npos) continue;
    std::string k = kv.substr(0, eq), v = kv.substr(eq + 1);
    if (k == "kind") {
      // numbering shared with csrc/cuda/optimizer.cuh
      static const char* const names[] = {"sgd", "sgd_mom", "adam", "test", "none", "nag", "signum", "signsgd",
                                          "adagrad", "rmsprop", "adadelta", "ftrl", "adamax"};
      o->kind = -1;
      for (int i = 0; i < 13; ++i) if (v == names[i]) o->kind = i;
      if (o->kind == 4) o->kind = -1;
    }
    else if (k == "lr") o->lr = (float)atof(v.c_str());
    else if (k == "wd") o->wd = (float)atof(v.c_str());
    else if (k == "momentum") o->momentum = (float)atof(v.c_str());
    else if (k == "beta1") o->beta1 = (float)atof(v.c_str());
    else if (k == "beta2") o->beta2 = (float)atof(v.c_str());
    else if (k == "eps") o->eps = (float)atof(v.c_str());
    else if (k == "rescale") o->rescale = (float)atof(v.c_str());
    else if (k == "clip") o->clip = (float)atof(v.c_str());
  }
  return true;
}

void KVServer::HandleCommand(Message&& m) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    switch (m.head) {
      case kStopServer: stop_ = true; break;
      case kSyncMode: sync_mode_ = true; break;
      case kSetMultiPrecision: multi_precision_ = true; break;
      case kSetGradientCompression: {
        // body "2bit:<threshold>"
        gc_type_ = m.body.rfind("2bit", 0) == 0 ? 2 : 0;
        size_t c = m.body.find(':');
        gc_threshold_ = c == std::string::npos ? 0.5f : (float)atof(m.body.c_str() + c + 1);
        break;
      }
      case kSetProfilerParams: HandleProfilerCommand(m.body); break;
      case kSetOptimizer: ParseOpt(m.body, &opt_); break;
      case kController:
      default: break;
    }
  }
  const bool is_stop = m.head == kStopServer;
  if (py_controller_) {
    // python part + the response that releases the worker run on the Run() thread, in arrival order.  The stop flag
    // is raised by that same task AFTER the response has been sent: raising it here would let Run() return between the
    // flag and the enqueue and drop the response (the worker that sent kStopServer would wait forever).
    auto msg = std::make_shared<Message>(std::move(m));
    PostToMain([this, msg, is_stop] {
      py_controller_(msg->head, msg->body.data(), msg->body.size(), py_ctl_user_);
      Respond(*msg);
      if (is_stop) exec_stop_ = true;
    });
    return;
  }
  Respond(m);
  if (is_stop) {
    exec_stop_ = true;
    std::lock_guard<std::mutex> lk(exec_mu_);
    exec_cv_.notify_all();
  }
}

void KVServer::ApplyUpdate(uint64_t key, Entry& e, size_t begin, size_t n) {
  float* w = e.store.data() + begin;
  const float* g = e.merged.data() + begin;
  if (py_updater_) {
    const int rel = (int)(key - node_->ServerKeyRange(node_->my_rank()).first);
    ExecOnMain([&] { py_updater_(rel, g, w, n, py_user_); });
    return;
  }
  OptimizerConfig& o = opt_;
  if (o.kind < 0) { memcpy(w, g, n * 4); return; }   // no updater: store the aggregate (kvstore_dist_server.h:358-361)
  const bool use1 = o.kind == 1 || o.kind == 2 || o.kind == 5 || o.kind == 6 || (o.kind >= 8 && o.kind <= 12);
  const bool use2 = o.kind == 2 || (o.kind >= 10 && o.kind <= 12);
  if (use1 && e.s1.size() != e.store.size()) e.s1.assign(e.store.size(), 0.f);
  if (use2 && e.s2.size() != e.store.size()) e.s2.assign(e.store.size(), 0.f);
  float lr = o.lr;
  const double t = (double)(e.t + 1);
  if (o.kind == 2)
    lr = (float)(o.lr * std::sqrt(1.0 - std::pow((double)o.beta2, t)) / (1.0 - std::pow((double)o.beta1, t)));
  else if (o.kind == 12)
    lr = (float)(o.lr / (1.0 - std::pow((double)o.beta1, t)));
  float* s1 = e.s1.empty() ? nullptr : e.s1.data() + begin;
  float* s2 = e.s2.empty() ? nullptr : e.s2.data() + begin;
  auto sgn = [](float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); };
  // same per-element math as csrc/cuda/optimizer.cuh:opt_apply
  for (size_t i = 0; i < n; ++i) {
    float gi = g[i] * o.rescale;
    if (o.clip > 0.f) gi = std::min(std::max(gi, -o.clip), o.clip);
    float& wi = w[i];
    switch (o.kind) {
      case 0: wi = (1.f - lr * o.wd) * wi - lr * gi; break;
      case 1: s1[i] = o.momentum * s1[i] - lr * o.wd * wi - lr * gi; wi += s1[i]; break;
      case 2:
        gi += o.wd * wi;
        s1[i] = o.beta1 * s1[i] + (1.f - o.beta1) * gi;
        s2[i] = o.beta2 * s2[i] + (1.f - o.beta2) * gi * gi;
        wi -= lr * s1[i] / (std::sqrt(s2[i]) + o.eps);
        break;
      case 5:
        gi += o.wd * wi;
        s1[i] = o.momentum * s1[i] + gi;
        wi -= lr * (gi + o.momentum * s1[i]);
        break;
      case 6:
        gi += o.wd * wi;
        s1[i] = o.momentum * s1[i] - (1.f - o.momentum) * gi;
        wi = (1.f - lr * o.beta1) * wi + lr * sgn(s1[i]);
        break;
      case 7:
        gi += o.wd * wi;
        wi = (1.f - lr * o.beta1) * wi - lr * sgn(gi);
        break;
      case 8:
        s1[i] += gi * gi;
        wi -= lr * (gi / std::sqrt(s1[i] + o.eps) + o.wd * wi);
        break;
      case 9:
        gi += o.wd * wi;
        s1[i] = o.beta1 * s1[i] + (1.f - o.beta1) * gi * gi;
        wi -= lr * gi / std::sqrt(s1[i] + o.eps);
        if (o.beta2 > 0.f) wi = std::min(std::max(wi, -o.beta2), o.beta2);
        break;
      case 10: {
        s1[i] = o.beta1 * s1[i] + (1.f - o.beta1) * gi * gi;
        const float d = std::sqrt(s2[i] + o.eps) / std::sqrt(s1[i] + o.eps) * gi;
        s2[i] = o.beta1 * s2[i] + (1.f - o.beta1) * d * d;
        wi = wi - d - o.wd * wi;
        break;
      }
      case 11:
        s1[i] += gi - (std::sqrt(s2[i] + gi * gi) - std::sqrt(s2[i])) * wi / lr;
        s2[i] += gi * gi;
        wi = std::fabs(s1[i]) > o.beta1 ? (sgn(s1[i]) * o.beta1 - s1[i]) / ((o.beta2 + std::sqrt(s2[i])) / lr + o.wd) : 0.f;
        break;
      case 12:
        gi += o.wd * wi;
        s1[i] = o.beta1 * s1[i] + (1.f - o.beta1) * gi;
        s2[i] = std::max(o.beta2 * s2[i], std::fabs(gi));
        wi -= lr * s1[i] / (s2[i] + 1e-38f);
        break;
      default: wi += gi; break;  // Test optimizer: w += rescale * g
    }
  }
}

static long long ProfNowUs() {
  using namespace std::chrono;
  return duration_cast<microseconds>(steady_clock::now().time_since_epoch()).count();
}

// body: "config:<file>" | "state:run" | "state:stop" | "dump" (a bare file name is accepted as config)
void KVServer::HandleProfilerCommand(const std::string& body) {
  if (body.rfind("state:", 0) == 0) {
    prof_running_ = body.substr(6) == "run";
  } else if (body == "dump") {
    DumpProfile();
  } else {
    std::lock_guard<std::mutex> lk(prof_mu_);
    prof_file_ = body.rfind("config:", 0) == 0 ? body.substr(7) : body;
    profiler_prefix_ = "rank" + std::to_string(node_->my_rank()) + "_";
  }
}

void KVServer::ProfRecord(const char* name, long long t0_us, uint64_t key, size_t bytes) {
  if (!prof_running_) return;
  const long long now = ProfNowUs();
  std::lock_guard<std::mutex> lk(prof_mu_);
  prof_events_.push_back(ProfEvent{name, t0_us, now - t0_us, key, bytes});
}

int KVServer::DumpProfile() {
  std::lock_guard<std::mutex> lk(prof_mu_);
  std::string dir, base = prof_file_;
  const size_t slash = base.rfind('/');
  if (slash != std::string::npos) { dir = base.substr(0, slash + 1); base = base.substr(slash + 1); }
  const std::string path = dir + "rank" + std::to_string(node_->my_rank()) + "_" + base;
  FILE* f = fopen(path.c_str(), "w");
  if (!f) return -1;
  fprintf(f, "{\"traceEvents\": [");
  for (size_t i = 0; i < prof_events_.size(); ++i) {
    const ProfEvent& e = prof_events_[i];
    fprintf(f, "%s{\"name\": \"%s\", \"cat\": \"kvstore_server\", \"ph\": \"X\", \"pid\": %d, \"tid\": 0, \"ts\": %lld, "
               "\"dur\": %lld, \"args\": {\"key\": %llu, \"bytes\": %zu}}",
            i ? ", " : "", e.name.c_str(), node_->my_rank(), e.ts_us, e.dur_us, (unsigned long long)e.key, e.bytes);
  }
  fprintf(f, "], \"displayTimeUnit\": \"ms\"}\n");
  fclose(f);
  return 0;
}

void KVServer::HandleData(Message&& m) {
  int req_type, dtype;
  UnpairCmd(m.head, &req_type, &dtype);
  uint64_t key = 0;
  if (m.data.empty() || m.data[0].size() < 8) return;
  memcpy(&key, m.data[0].data(), 8);
  const long long prof_t0 = prof_running_ ? ProfNowUs() : 0;
  const size_t prof_bytes = m.data.size() > 1 ? m.data[1].size() : 0;
  const bool prof_push = m.push;
  struct ProfGuard {
    KVServer* s; const char* name; long long t0; uint64_t key; size_t bytes;
    ~ProfGuard() { if (t0) s->ProfRecord(name, t0, key, bytes); }
  } prof_guard{this, prof_push ? "KVStoreDistServerPush" : "KVStoreDistServerPull", prof_t0, key, prof_bytes};
  std::unique_lock<std::mutex> lk(mu_);
  Entry& e = store_[key];

  if (!m.push) {  // ------------------------------------------------ pull
    if (req_type == kRowSparsePushPull) {
      uint64_t meta[2]; memcpy(meta, m.data[3].data(), 16);
      const size_t row_len = meta[1], nr = m.data[1].size() / 8, es = DTypeSize(dtype);
      std::string rows(nr * row_len * es, '\0');
      for (size_t r = 0; r < nr; ++r) {
        int64_t id; memcpy(&id, m.data[1].data() + r * 8, 8);
        if ((size_t)(id + 1) * row_len <= e.store.size())
          FromFloat(e.store.data() + (size_t)id * row_len, dtype, row_len, &rows[r * row_len * es]);
      }
      Message r;
      r.recver = m.sender; r.request = false; r.push = false; r.timestamp = m.timestamp; r.head = m.head;
      r.data.push_back(std::move(rows));
      r.data.push_back(m.data[2]);
      lk.unlock();
      node_->Send(r);
      return;
    }
    std::string out(e.store.size() * DTypeSize(dtype), '\0');
    FromFloat(e.store.data(), dtype, e.store.size(), &out[0]);
    lk.unlock();
    Respond(m, &out);
    return;
  }

  // ---------------------------------------------------------------- push
  size_t n = 0;
  if (req_type == kRowSparsePushPull) { uint64_t meta[2]; memcpy(meta, m.data[3].data(), 16); n = meta[0]; }
  else if (req_type == kCompressedPushPull) { uint64_t meta; memcpy(&meta, m.data[2].data(), 8); n = meta; }
  else n = m.data[1].size() / DTypeSize(dtype);

  if (!e.initialized) {
    // first push initialises the stored value (kvstore_dist_server.h:707-714)
    e.store.assign(n, 0.f);
    if (req_type == kDefaultPushPull) ToFloat(m.data[1].data(), dtype, n, e.store.data());
    e.initialized = true;
    lk.unlock();
    Respond(m);
    return;
  }

  const bool first = e.waiting.empty() || !sync_mode_;
  if (e.merged.size() != e.store.size()) e.merged.assign(e.store.size(), 0.f);
  if (first && req_type != kDefaultPushPull) std::fill(e.merged.begin(), e.merged.end(), 0.f);
  if (req_type == kDefaultPushPull) {
    // the first pusher's values are converted straight into the merge buffer, later ones are added with the conversion
    // fused in: one pass over the data instead of fill + temporary + convert + add
    const size_t cnt = std::min(n, e.merged.size());
    if (first) {
      ToFloat(m.data[1].data(), dtype, cnt, e.merged.data());
      if (cnt < e.merged.size()) std::fill(e.merged.begin() + cnt, e.merged.end(), 0.f);
    } else {
      AccumulateFloat(m.data[1].data(), dtype, cnt, e.merged.data());
    }
  } else if (req_type == kCompressedPushPull) {
    // 2-bit codes, 16 per 32-bit word: 11 -> +thr, 10 -> -thr (gradient_compression-inl.h:96-140)
    const uint32_t* words = reinterpret_cast<const uint32_t*>(m.data[1].data());
    for (size_t i = 0; i < n && i < e.merged.size(); ++i) {
      uint32_t wd = words[i / 16];
      int j = (int)(i % 16);
      uint32_t c = (wd >> (8 * (j >> 2) + 6 - 2 * (j & 3))) & 3u;
      e.merged[i] += c == 3u ? gc_threshold_ : (c == 2u ? -gc_threshold_ : 0.f);
    }
  } else {  // row sparse
    uint64_t meta[2]; memcpy(meta, m.data[3].data(), 16);
    const size_t row_len = meta[1], nr = m.data[2].size() / 8;
    if (e.touched.size() * row_len != e.store.size()) e.touched.assign(e.store.size() / std::max<size_t>(1, row_len), 0);
    if (first) std::fill(e.touched.begin(), e.touched.end(), 0);
    std::vector<float> tmp(row_len);
    for (size_t r = 0; r < nr; ++r) {
      int64_t id; memcpy(&id, m.data[2].data() + r * 8, 8);
      if ((size_t)(id + 1) * row_len > e.merged.size()) continue;
      ToFloat(m.data[1].data() + r * row_len * DTypeSize(dtype), dtype, row_len, tmp.data());
      for (size_t i = 0; i < row_len; ++i) e.merged[(size_t)id * row_len + i] += tmp[i];
      e.touched[(size_t)id] = 1;
    }
  }

  auto apply = [&]() {
    if (req_type == kRowSparsePushPull) {
      const size_t row_len = e.store.size() / std::max<size_t>(1, e.touched.size());
      for (size_t r = 0; r < e.touched.size(); ++r)
        if (e.touched[r]) ApplyUpdate(key, e, r * row_len, row_len);
    } else {
      ApplyUpdate(key, e, 0, e.store.size());
    }
    ++e.t;
  };

  if (!sync_mode_) {
    // async: every push is applied at once, an updater is mandatory (kvstore_dist_server.h:348,715-723)
    apply();
    lk.unlock();
    Respond(m);
    return;
  }
  e.waiting.push_back(std::move(m));
  if ((int)e.waiting.size() == node_->num_workers()) {
    apply();
    std::vector<Message> done;
    done.swap(e.waiting);
    lk.unlock();
    for (const auto& req : done) Respond(req);   // the per-key sync barrier
  }
}

}  // namespace dlcfn
