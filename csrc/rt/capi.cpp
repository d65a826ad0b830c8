// C API of libdlcfn_rt.so (loaded through ctypes by deeplearning_cfn_b200/utils/rt.py).
// Mirrors the MXKVStore* C entry points the reference exposes to its Python front end
// (MX/src/c_api/c_api.cc:815-1137): create/role queries, push, pull, barrier, run-server, commands,
// dead-node count.
#include <cstring>
#include <memory>

#include "ps.h"

#define RT_API extern "C" __attribute__((visibility("default")))

using namespace dlcfn;

static std::unique_ptr<KVWorker> g_worker;
static std::unique_ptr<KVServer> g_server;

RT_API int dlcfn_ps_start() {
  Node* n = Node::Get();
  if (n->started()) return 0;
  // handlers must exist before the first data message can arrive
  const char* role = getenv("DMLC_ROLE");
  bool is_server = role && !strcmp(role, "server");
  bool is_worker = !role || !strcmp(role, "worker");
  if (is_worker) {
    g_worker.reset(new KVWorker(n));
    n->SetDataHandler([](Message&& m) { g_worker->OnMessage(std::move(m)); });
  } else if (is_server) {
    g_server.reset(new KVServer(n));
    n->SetDataHandler([](Message&& m) { g_server->OnMessage(std::move(m)); });
  }
  return n->Start();
}

RT_API void dlcfn_ps_finalize(int do_barrier) {
  Node::Get()->Finalize(do_barrier != 0);
  g_worker.reset();
  g_server.reset();
}

RT_API int dlcfn_ps_role() { return Node::Get()->role(); }
RT_API int dlcfn_ps_rank() { return Node::Get()->my_rank(); }
RT_API int dlcfn_ps_node_id() { return Node::Get()->my_id(); }
RT_API int dlcfn_ps_num_workers() { return Node::Get()->num_workers(); }
RT_API int dlcfn_ps_num_servers() { return Node::Get()->num_servers(); }
RT_API int dlcfn_ps_is_recovery() { return Node::Get()->is_recovery() ? 1 : 0; }
RT_API void dlcfn_ps_barrier(int group) { Node::Get()->Barrier(group); }
RT_API int dlcfn_ps_num_dead(int group, int timeout_sec) { return Node::Get()->NumDeadNodes(group, timeout_sec); }
RT_API long long dlcfn_ps_bytes_sent() { return Node::Get()->bytes_sent(); }
RT_API long long dlcfn_ps_bytes_recv() { return Node::Get()->bytes_recv(); }

// ---- worker
RT_API int dlcfn_kv_push(int key, const void* data, size_t num_elems, int dtype) {
  return g_worker ? g_worker->Push(key, data, num_elems, dtype) : -1;
}
RT_API int dlcfn_kv_pull(int key, void* out, size_t num_elems, int dtype) {
  return g_worker ? g_worker->Pull(key, out, num_elems, dtype) : -1;
}
RT_API int dlcfn_kv_push_rsp(int key, const void* rows, const int64_t* row_ids, size_t num_rows, size_t row_len,
                             size_t total_elems, int dtype) {
  return g_worker ? g_worker->Push(key, rows, total_elems, dtype, kRowSparsePushPull, row_ids, num_rows, row_len) : -1;
}
RT_API int dlcfn_kv_pull_rsp(int key, void* out_rows, const int64_t* row_ids, size_t num_rows, size_t row_len,
                             size_t total_elems, int dtype) {
  return g_worker ? g_worker->Pull(key, out_rows, total_elems, dtype, kRowSparsePushPull, row_ids, num_rows, row_len) : -1;
}
RT_API int dlcfn_kv_push_2bit(int key, const void* words, size_t original_elems) {
  return g_worker ? g_worker->Push(key, words, original_elems, kF32, kCompressedPushPull) : -1;
}
RT_API void dlcfn_kv_wait(int ts) { if (g_worker) g_worker->Wait(ts); }
RT_API void dlcfn_kv_send_command(int head, const char* body, size_t len) {
  if (g_worker) g_worker->SendCommand(head, std::string(body ? body : "", len));
}
RT_API void dlcfn_kv_set_key_unit(int key, size_t unit) { if (g_worker) g_worker->SetKeyUnit(key, unit); }
RT_API void dlcfn_kv_set_bigarray_bound(size_t n) { if (g_worker) g_worker->bigarray_bound = n; }
// partition introspection: fills up to `cap` entries of (server, offset, len); returns the count
RT_API int dlcfn_kv_partition(int key, size_t num_elems, int* servers, size_t* offsets, size_t* lens, int cap) {
  if (!g_worker) return -1;
  auto parts = g_worker->Partition(key, num_elems);
  int n = 0;
  for (auto& p : parts) {
    if (n >= cap) break;
    servers[n] = p.server; offsets[n] = p.offset; lens[n] = p.len; ++n;
  }
  return (int)parts.size();
}

// ---- server
RT_API void dlcfn_kv_server_set_updater(KVServer::PyUpdater f, void* user) { if (g_server) g_server->SetPyUpdater(f, user); }
RT_API void dlcfn_kv_server_set_controller(KVServer::PyController f, void* user) { if (g_server) g_server->SetPyController(f, user); }
RT_API void dlcfn_kv_server_run() { if (g_server) g_server->Run(); }
