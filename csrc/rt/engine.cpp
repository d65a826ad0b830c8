// Dependency engine: versioned-variable read/write scheduling on a priority thread pool.
//
// Reference E1 (MX/src/engine/threaded_engine.{h,cc}, threaded_engine_perdevice.cc): every op declares the
// variables it reads and mutates; an op runs when all earlier writers of its read-vars and all earlier
// readers+writers of its write-vars have finished; a priority CPU pool serves `kCPUPrioritized` ops, which is
// how `priority=-index` makes front-layer push/pull jump the queue (model.py:152-155); exceptions propagate
// through variables (threaded_engine.h:179-180); `NaiveEngine` runs everything synchronously for debugging
// (engine.cc:38-44, MXNET_ENGINE_TYPE).
//
// In the B200 design GPU work is ordered by CUDA streams/events, so this engine only schedules HOST-side
// operations: the CPU KVStore's asynchronous push/pull (ordering: pull-after-push on the same key) and
// checkpoint / staging IO.  Same contract, small implementation.
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#define RT_API extern "C" __attribute__((visibility("default")))

namespace dlcfn {

struct Opr;

struct Var {
  // pending queue of (op, is_write) in program order
  std::deque<std::pair<Opr*, bool>> queue;
  int running_reads = 0;
  bool running_write = false;
  int error = 0;  // sticky error code of the last failed writer
};

struct Opr {
  std::function<int()> fn;
  std::vector<int> reads, writes;
  int priority = 0;
  std::atomic<int> wait{0};
  uint64_t seq = 0;
};

class Engine {
 public:
  explicit Engine(int nthreads, bool naive) : naive_(naive) {
    if (!naive_)
      for (int i = 0; i < nthreads; ++i) workers_.emplace_back([this] { Loop(); });
  }
  ~Engine() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }

  int NewVar() {
    std::lock_guard<std::mutex> lk(mu_);
    int id = next_var_++;
    vars_[id];
    return id;
  }

  void Push(std::function<int()> fn, const std::vector<int>& reads, const std::vector<int>& writes, int priority) {
    if (naive_) {
      int rc = fn();
      if (rc) {
        std::lock_guard<std::mutex> lk(mu_);
        for (int v : writes) vars_[v].error = rc;
      }
      return;
    }
    Opr* op = new Opr();
    op->fn = std::move(fn);
    op->reads = reads;
    op->writes = writes;
    op->priority = priority;
    std::lock_guard<std::mutex> lk(mu_);
    op->seq = seq_++;
    ++pending_;
    int blocked = 0;
    for (int v : reads) {
      Var& var = vars_[v];
      if (var.running_write || !var.queue.empty()) { var.queue.emplace_back(op, false); ++blocked; }
      else ++var.running_reads;
    }
    for (int v : writes) {
      Var& var = vars_[v];
      if (var.running_write || var.running_reads > 0 || !var.queue.empty()) { var.queue.emplace_back(op, true); ++blocked; }
      else var.running_write = true;
    }
    op->wait = blocked;
    if (blocked == 0) Enqueue(op);
  }

  int WaitForVar(int v) {
    // a no-op reader: returns when all previously pushed writers of v are done
    std::mutex m;
    std::condition_variable c;
    bool done = false;
    Push([&] { std::lock_guard<std::mutex> lk(m); done = true; c.notify_all(); return 0; }, {v}, {}, 1 << 30);
    std::unique_lock<std::mutex> lk(m);
    c.wait(lk, [&] { return done; });
    std::lock_guard<std::mutex> lk2(mu_);
    int e = vars_[v].error;
    vars_[v].error = 0;
    return e;
  }

  void WaitForAll() {
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
  }

 private:
  struct Cmp {
    bool operator()(Opr* a, Opr* b) const {
      return a->priority != b->priority ? a->priority < b->priority : a->seq > b->seq;
    }
  };
  void Enqueue(Opr* op) {  // mu_ held
    ready_.push(op);
    cv_.notify_one();
  }
  void Loop() {
    for (;;) {
      Opr* op = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || !ready_.empty(); });
        if (stop_ && ready_.empty()) return;
        op = ready_.top();
        ready_.pop();
      }
      int rc = 0;
      try {
        rc = op->fn();
      } catch (...) {
        rc = -1;
      }
      Complete(op, rc);
    }
  }
  void Complete(Opr* op, int rc) {
    std::lock_guard<std::mutex> lk(mu_);
    for (int v : op->reads) {
      Var& var = vars_[v];
      --var.running_reads;
      Dispatch(var);
    }
    for (int v : op->writes) {
      Var& var = vars_[v];
      var.running_write = false;
      if (rc) var.error = rc;
      Dispatch(var);
    }
    delete op;
    if (--pending_ == 0) done_cv_.notify_all();
  }
  void Dispatch(Var& var) {  // mu_ held: admit the head of the queue while compatible
    while (!var.queue.empty()) {
      auto [op, is_write] = var.queue.front();
      if (is_write) {
        if (var.running_reads > 0 || var.running_write) return;
        var.running_write = true;
      } else {
        if (var.running_write) return;
        ++var.running_reads;
      }
      var.queue.pop_front();
      if (--op->wait == 0) Enqueue(op);
      if (is_write) return;
    }
  }

  bool naive_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  bool stop_ = false;
  std::vector<std::thread> workers_;
  std::priority_queue<Opr*, std::vector<Opr*>, Cmp> ready_;
  std::unordered_map<int, Var> vars_;
  int next_var_ = 1;
  uint64_t seq_ = 0;
  int pending_ = 0;
};

}  // namespace dlcfn

using dlcfn::Engine;
typedef int (*dlcfn_engine_fn)(void* user);

RT_API void* dlcfn_engine_create(int nthreads, int naive) {
  const char* t = getenv("MXNET_ENGINE_TYPE");
  if (t && !strcmp(t, "NaiveEngine")) naive = 1;
  if (nthreads <= 0) {
    const char* n = getenv("MXNET_CPU_WORKER_NTHREADS");
    nthreads = n ? atoi(n) : 4;
  }
  return new Engine(nthreads, naive != 0);
}
RT_API void dlcfn_engine_destroy(void* e) { delete (Engine*)e; }
RT_API int dlcfn_engine_new_var(void* e) { return ((Engine*)e)->NewVar(); }
RT_API void dlcfn_engine_push(void* e, dlcfn_engine_fn fn, void* user, const int* reads, int nreads, const int* writes,
                              int nwrites, int priority) {
  std::vector<int> r(reads, reads + nreads), w(writes, writes + nwrites);
  ((Engine*)e)->Push([fn, user] { return fn(user); }, r, w, priority);
}
RT_API int dlcfn_engine_wait_var(void* e, int var) { return ((Engine*)e)->WaitForVar(var); }
RT_API void dlcfn_engine_wait_all(void* e) { ((Engine*)e)->WaitForAll(); }
