// Handle-based C-level KVStore API of libdlcfn_rt.so - the counterpart of the MXKVStore* entry points every MXNet language
// binding sits on (MX/src/c_api/c_api.cc:815-1137: MXKVStoreCreate / Init / Push / Pull / SetUpdater / Barrier / GetRank /
// GetGroupSize / GetType / IsWorkerNode / RunServer / SendCommmandToServers / GetNumDeadNode / Free).  A C or C++ program can
// create and drive a store without Python:
//
//   "local" / "device"      in-process store over host buffers: Init copies the value, Push takes `num` values per key, sums
//                           them with the threaded reducer (csrc/rt/reduce.cpp, the CommCPU path) and either applies the
//                           registered updater or stores the sum; Pull copies the stored value out (kvstore_local.h)
//   "dist_sync" / "dist_async" / "dist_device_sync"
//                           worker side of the parameter-server transport (csrc/rt/kv.cpp KVWorker: key sharding, big-array
//                           split, sync / async servers); needs the DMLC_* environment of a launcher
//
// Values are dense fp32 / bf16 arrays addressed by integer keys (dtype codes of ps.h: kF32 = 0, kBF16 = 1).
// All functions return 0 on success, -1 on error (DLCFNKVGetLastError()).
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "ps.h"

#define RT_API extern "C" __attribute__((visibility("default")))

extern "C" int dlcfn_reduce_sum_cpu(const void* const* srcs, int nsrc, void* dst, size_t n, int dtype, int nthreads, size_t bigarray);
extern "C" int dlcfn_ps_start();
extern "C" void dlcfn_ps_finalize(int do_barrier);
extern "C" int dlcfn_ps_rank();
extern "C" int dlcfn_ps_num_workers();
extern "C" void dlcfn_ps_barrier(int group);
extern "C" int dlcfn_ps_num_dead(int group, int timeout_sec);
extern "C" int dlcfn_ps_is_recovery();
extern "C" int dlcfn_kv_push(int key, const void* data, size_t num_elems, int dtype);
extern "C" int dlcfn_kv_pull(int key, void* out, size_t num_elems, int dtype);
extern "C" void dlcfn_kv_wait(int ts);
extern "C" void dlcfn_kv_send_command(int head, const char* body, size_t len);
extern "C" void dlcfn_kv_set_key_unit(int key, size_t unit);
extern "C" void dlcfn_kv_server_run();

using namespace dlcfn;

namespace {

thread_local std::string g_err;
int fail(const std::string& m) {
  g_err = m;
  return -1;
}

// updater(key, recv, local, num_elems, dtype, user): merge `recv` into `local` in place (MXKVStoreUpdater)
typedef void (*KVUpdater)(int key, const void* recv, void* local, size_t num_elems, int dtype, void* user);

struct Entry {
  std::vector<char> value;
  size_t n = 0;
  int dtype = 0;
};

struct Store {
  std::string type;
  bool dist = false;
  std::map<int, Entry> local;        // local store: the values; dist store: shapes only (value empty)
  KVUpdater updater = nullptr;
  void* updater_user = nullptr;
  std::mutex mu;
};

size_t esize(int dtype) { return dtype == kF32 ? 4 : 2; }

}  // namespace

RT_API const char* DLCFNKVGetLastError() { return g_err.c_str(); }

RT_API int DLCFNKVStoreCreate(const char* type, void** out) {
  if (!type || !out) return fail("null argument");
  std::string t(type);
  std::unique_ptr<Store> s(new Store());
  s->type = t;
  if (t == "local" || t == "device" || t == "local_allreduce_cpu" || t == "local_allreduce_device") {
    s->dist = false;
  } else if (t.rfind("dist", 0) == 0) {
    s->dist = true;
    if (!getenv("DMLC_PS_ROOT_URI") || !getenv("DMLC_PS_ROOT_PORT") || !getenv("DMLC_NUM_WORKER") || !getenv("DMLC_NUM_SERVER"))
      return fail("a dist store is a node of a parameter-server job: DMLC_PS_ROOT_URI / DMLC_PS_ROOT_PORT / DMLC_NUM_WORKER / "
                  "DMLC_NUM_SERVER must be set (start the program through the launcher)");
    if (dlcfn_ps_start() != 0) return fail("cannot join the parameter-server job");
    if (!dlcfn_ps_is_recovery()) dlcfn_ps_barrier(7);
    if (dlcfn_ps_rank() == 0 && t.find("_async") == std::string::npos) dlcfn_kv_send_command(kSyncMode, "", 0);
  } else {
    return fail("unknown kvstore type " + t);
  }
  *out = s.release();
  return 0;
}

RT_API int DLCFNKVStoreFree(void* h) {
  Store* s = (Store*)h;
  if (!s) return 0;
  if (s->dist) {      // destructor protocol of KVStoreDist (kvstore_dist.h:62-76): barrier, rank 0 stops the servers
    dlcfn_ps_barrier(4);
    if (dlcfn_ps_rank() == 0) dlcfn_kv_send_command(kStopServer, "", 0);
    dlcfn_ps_finalize(1);
  }
  delete s;
  return 0;
}

RT_API int DLCFNKVStoreGetType(void* h, const char** out) {
  *out = ((Store*)h)->type.c_str();
  return 0;
}
RT_API int DLCFNKVStoreGetRank(void* h, int* out) {
  *out = ((Store*)h)->dist ? dlcfn_ps_rank() : 0;
  return 0;
}
RT_API int DLCFNKVStoreGetGroupSize(void* h, int* out) {
  *out = ((Store*)h)->dist ? dlcfn_ps_num_workers() : 1;
  return 0;
}
RT_API int DLCFNKVStoreIsWorkerNode(int* out) {
  const char* r = getenv("DMLC_ROLE");
  *out = (!r || !strcmp(r, "worker")) ? 1 : 0;
  return 0;
}
RT_API int DLCFNKVStoreIsServerNode(int* out) {
  const char* r = getenv("DMLC_ROLE");
  *out = (r && !strcmp(r, "server")) ? 1 : 0;
  return 0;
}
RT_API int DLCFNKVStoreIsSchedulerNode(int* out) {
  const char* r = getenv("DMLC_ROLE");
  *out = (r && !strcmp(r, "scheduler")) ? 1 : 0;
  return 0;
}
// Server / scheduler processes of a job call this instead of creating a store: join, serve until the workers' rank 0 sends
// kStopServer (servers) or until every node has left (scheduler), leave.  (MXKVStoreRunServer, c_api.cc:1100-1137; the native
// server applies the optimizers of csrc/cuda/optimizer.cuh itself, so no controller callback is needed from C.)
RT_API int DLCFNKVStoreRunServer() {
  if (dlcfn_ps_start() != 0) return fail("parameter-server node could not start");
  if (!dlcfn_ps_is_recovery()) dlcfn_ps_barrier(7);
  int server = 0;
  DLCFNKVStoreIsServerNode(&server);
  if (server) dlcfn_kv_server_run();
  dlcfn_ps_finalize(1);
  return 0;
}
RT_API int DLCFNKVStoreBarrier(void* h) {
  if (((Store*)h)->dist) dlcfn_ps_barrier(4);
  return 0;
}
RT_API int DLCFNKVStoreGetNumDeadNode(void* h, int node_id, int timeout_sec, int* out) {
  *out = ((Store*)h)->dist ? dlcfn_ps_num_dead(node_id, timeout_sec) : 0;
  return 0;
}
RT_API int DLCFNKVStoreSendCommandToServers(void* h, int head, const char* body) {
  if (((Store*)h)->dist) dlcfn_kv_send_command(head, body ? body : "", body ? strlen(body) : 0);
  return 0;
}
RT_API int DLCFNKVStoreSetUpdater(void* h, KVUpdater f, void* user) {
  Store* s = (Store*)h;
  std::lock_guard<std::mutex> lk(s->mu);
  s->updater = f;
  s->updater_user = user;
  return 0;
}

RT_API int DLCFNKVStoreInit(void* h, int num, const int* keys, const void* const* vals, const size_t* lens, const int* dtypes) {
  Store* s = (Store*)h;
  std::lock_guard<std::mutex> lk(s->mu);
  for (int i = 0; i < num; ++i) {
    if (s->local.count(keys[i])) return fail("duplicate init of key " + std::to_string(keys[i]));
    if (dtypes[i] < 0 || dtypes[i] > 1) return fail("unsupported dtype (0 = fp32, 1 = bf16)");
    Entry& e = s->local[keys[i]];
    e.n = lens[i];
    e.dtype = dtypes[i];
    if (!s->dist) {
      e.value.assign((const char*)vals[i], (const char*)vals[i] + lens[i] * esize(dtypes[i]));
    } else {
      dlcfn_kv_set_key_unit(keys[i], 16);
      if (dlcfn_ps_rank() == 0) dlcfn_kv_wait(dlcfn_kv_push(keys[i], vals[i], lens[i], dtypes[i]));   // rank 0's value initialises the servers
    }
  }
  if (s->dist && !dlcfn_ps_is_recovery()) dlcfn_ps_barrier(4);
  return 0;
}

// `num` (key, value) pairs; several values with the same key are summed first (the per-device gradients of one parameter)
RT_API int DLCFNKVStorePush(void* h, int num, const int* keys, const void* const* vals, int priority) {
  (void)priority;
  Store* s = (Store*)h;
  std::lock_guard<std::mutex> lk(s->mu);
  std::map<int, std::vector<const void*>> grouped;
  for (int i = 0; i < num; ++i) grouped[keys[i]].push_back(vals[i]);
  for (auto& kv : grouped) {
    auto it = s->local.find(kv.first);
    if (it == s->local.end()) return fail("key " + std::to_string(kv.first) + " has not been initialized");
    Entry& e = it->second;
    std::vector<char> merged(e.n * esize(e.dtype));
    if (kv.second.size() == 1) memcpy(merged.data(), kv.second[0], merged.size());
    else if (dlcfn_reduce_sum_cpu(kv.second.data(), (int)kv.second.size(), merged.data(), e.n, e.dtype, 4, 1000000) != 0)
      return fail("reduce failed");
    if (s->dist) {
      dlcfn_kv_wait(dlcfn_kv_push(kv.first, merged.data(), e.n, e.dtype));
    } else if (s->updater) {
      s->updater(kv.first, merged.data(), e.value.data(), e.n, e.dtype, s->updater_user);
    } else {
      e.value.swap(merged);           // no updater: the store keeps the merged value (kvstore_local.h:179-186)
    }
  }
  return 0;
}

RT_API int DLCFNKVStorePull(void* h, int num, const int* keys, void* const* outs, int priority) {
  (void)priority;
  Store* s = (Store*)h;
  std::lock_guard<std::mutex> lk(s->mu);
  for (int i = 0; i < num; ++i) {
    auto it = s->local.find(keys[i]);
    if (it == s->local.end()) return fail("key " + std::to_string(keys[i]) + " has not been initialized");
    Entry& e = it->second;
    if (s->dist) dlcfn_kv_wait(dlcfn_kv_pull(keys[i], outs[i], e.n, e.dtype));
    else memcpy(outs[i], e.value.data(), e.value.size());
  }
  return 0;
}
