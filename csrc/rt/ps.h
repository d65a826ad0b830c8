// dlcfn parameter-server transport (CPU control/data plane) - the native replacement for ps-lite.
//
// Reference components and their counterparts here:
//   Postoffice (PS/src/postoffice.cc)      -> Node: role/env parsing, id groups, Barrier, key ranges, dead nodes
//   Van / ZMQVan (PS/src/van.cc, zmq_van.h)-> Node: TCP framing, scheduler rendezvous (ADD_NODE -> sorted rank
//                                             assignment -> node table broadcast), receiver threads, heartbeats,
//                                             PS_DROP_MSG fault injection, dead-node replacement (recovery)
//   Resender (PS/src/resender.h)           -> Node: ACK + signature de-dup + timed resend (PS_RESEND)
//   Customer / SimpleApp / KVWorker / KVServer (PS/include/ps/*.h)
//                                          -> KVWorker / KVServer below: request tracking, key-range slicing,
//                                             command channel
//   KVStoreDist + KVStoreDistServer (MX/src/kvstore/kvstore_dist{,_server}.h)
//                                          -> key sharding (EncodeDefaultKey) in KVWorker, sync/async aggregation +
//                                             multi-precision + built-in optimizers in KVServer
// The wire format is a fixed little-endian header + body + blobs (no protobuf / ZeroMQ dependency).
// On a B200 box this transport carries control traffic and the CPU (gloo-free) data path used by the
// tests and by CPU-only nodes; GPU tensors travel through the peer-HBM kernels instead.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <map>
#include <set>
#include <memory>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace dlcfn {

enum Role : int { kScheduler = 0, kServer = 1, kWorker = 2 };
enum Group : int { kSchedulerGroup = 1, kServerGroup = 2, kWorkerGroup = 4 };
enum Control : int { kNone = 0, kAddNode = 1, kBarrier = 2, kTerminate = 3, kHeartbeat = 4, kAck = 5, kNodeTable = 6 };

constexpr int kSchedulerId = 1;
inline int ServerRankToId(int r) { return 8 + 2 * r; }
inline int WorkerRankToId(int r) { return 9 + 2 * r; }
inline int IdToRank(int id) { return (id - 8) / 2; }

struct NodeInfo {
  int role = kWorker;
  int id = -1;
  std::string host;
  int port = 0;
  bool is_recovery = false;
};

struct Message {
  // meta
  int sender = -1, recver = -1;
  int control = kNone;
  int barrier_group = 0;
  int app_id = 0, customer_id = 0;
  int timestamp = -1;
  bool request = false, push = false, simple_app = false;
  int head = 0;          // command / data-type multiplex
  uint64_t sig = 0;      // resender signature (0 = untracked)
  std::string body;
  std::vector<NodeInfo> nodes;            // control payload
  std::vector<std::string> data;          // blobs: keys / vals / lens
  // Optional LAST data element that is only borrowed: sent from the caller's memory after `data` (the receiver sees one
  // more element).  Valid until Send() returns; Send() copies it into `data` first when the message must be kept for
  // retransmission.
  const char* blob = nullptr;
  size_t blob_len = 0;
};

class Node {
 public:
  static Node* Get();
  // reads DMLC_ROLE, DMLC_PS_ROOT_URI/PORT, DMLC_NUM_WORKER/SERVER, DMLC_NODE_HOST, PS_* knobs
  int Start(int customer_id_unused = 0);
  void Finalize(bool do_barrier = true);
  bool started() const { return started_; }

  int role() const { return role_; }
  int my_id() const { return my_.id; }
  int my_rank() const { return role_ == kScheduler ? 0 : IdToRank(my_.id); }
  int num_workers() const { return num_workers_; }
  int num_servers() const { return num_servers_; }
  bool is_recovery() const { return my_.is_recovery; }

  void Barrier(int group);
  int NumDeadNodes(int group, int timeout_sec);
  // uniform split of the 64-bit key space across servers (postoffice.cc:173-184)
  std::pair<uint64_t, uint64_t> ServerKeyRange(int server_rank) const;

  int Send(Message& msg);  // returns bytes sent or -1
  using Handler = std::function<void(Message&&)>;
  void SetDataHandler(Handler h) { data_handler_ = std::move(h); }
  int NextTimestamp() { return timestamp_++; }
  std::vector<int> IdsOfGroup(int group) const;
  long long bytes_sent() const { return bytes_sent_; }
  long long bytes_recv() const { return bytes_recv_; }

 private:
  Node() = default;
  void AcceptLoop();
  void ReadLoop(int fd);
  void HandleControl(Message&& m, int fd);
  void SchedulerAddNode(Message&& m);
  int ConnectTo(const NodeInfo& n);
  int FdFor(int id);
  void HeartbeatLoop();
  void ResendLoop();
  bool SendRaw(int fd, const Message& m);

  bool started_ = false;
  int role_ = kWorker, num_workers_ = 0, num_servers_ = 0, verbose_ = 0;
  NodeInfo my_, scheduler_;
  int listen_fd_ = -1;
  std::atomic<bool> stop_{false};
  std::atomic<bool> ready_{false};
  std::thread accept_thread_, heartbeat_thread_, resend_thread_;
  std::vector<std::thread> readers_;
  std::mutex mu_;                        // node table, fds
  std::condition_variable cv_;
  std::map<int, NodeInfo> nodes_;        // id -> info
  std::map<int, int> send_fd_;           // id -> fd
  std::map<int, std::unique_ptr<std::mutex>> send_mu_;
  std::set<int> stale_fd_;                // peers whose connection must be replaced by the next Send
  std::vector<NodeInfo> pending_;        // scheduler: registrations before assignment
  std::vector<int> pending_fd_;
  std::map<int, int> registered_fd_;     // scheduler: id -> fd of the registration connection
  std::map<int, int> barrier_count_;     // scheduler: group -> arrivals
  bool barrier_done_ = false;
  std::map<int, long long> last_seen_;   // scheduler: id -> epoch seconds
  Handler data_handler_;
  std::atomic<int> timestamp_{0};
  std::atomic<long long> bytes_sent_{0}, bytes_recv_{0};
  // resender
  bool resend_ = false;
  int resend_timeout_ms_ = 1000;
  int drop_pct_ = 0;
  int heartbeat_interval_ = 0;
  std::mutex rs_mu_;
  struct Pending { Message msg; long long sent_ms; int tries; };
  std::unordered_map<uint64_t, Pending> unacked_;
  std::unordered_set<uint64_t> seen_sig_;
  std::atomic<uint64_t> sig_counter_{1};
};

// ---------------------------------------------------------------------------------------------------------
// KV layer
enum RequestType : int { kDefaultPushPull = 0, kRowSparsePushPull = 1, kCompressedPushPull = 2 };
enum CommandType : int { kController = 0, kSetMultiPrecision = 1, kStopServer = 2, kSyncMode = 3,
                         kSetGradientCompression = 4, kSetProfilerParams = 5, kSetOptimizer = 6 };
enum DType : int { kF32 = 0, kBF16 = 1, kF16 = 2 };
// data commands multiplex (request type, dtype) with the Cantor pairing the reference uses
// (kvstore_dist_server.h:68-90)
inline int PairCmd(int a, int b) { return (a + b) * (a + b + 1) / 2 + b; }
void UnpairCmd(int z, int* a, int* b);

struct OptimizerConfig {
  int kind = -1;  // -1 none, else the numbering of csrc/cuda/optimizer.cuh (0 sgd ... 12 adamax)
  float lr = 0.01f, wd = 0.f, momentum = 0.f, beta1 = 0.9f, beta2 = 0.999f, eps = 1e-8f, rescale = 1.f, clip = -1.f;
};

class KVWorker {
 public:
  explicit KVWorker(Node* node);
  // sharding rule of KVStoreDist::EncodeDefaultKey (kvstore_dist.h:532-574)
  struct Part { int server; uint64_t ps_key; size_t offset, len; };
  std::vector<Part> Partition(int key, size_t num_elems) const;
  // Split boundaries of a big array are multiples of the key's unit: 16 elements (one packed word of the 2-bit wire
  // format) and, for keys with more than one dimension, whole rows (row-sparse push / pull address rows by local id) -
  // every request type of a key therefore sees the same shards (upstream splits row-sparse keys on shape[0],
  // kvstore_dist.h:571-600).  Set once per key before its first push.
  void SetKeyUnit(int key, size_t unit);
  int Push(int key, const void* data, size_t num_elems, int dtype, int req_type = kDefaultPushPull,
           const int64_t* row_ids = nullptr, size_t num_rows = 0, size_t row_len = 0);
  int Pull(int key, void* out, size_t num_elems, int dtype, int req_type = kDefaultPushPull,
           const int64_t* row_ids = nullptr, size_t num_rows = 0, size_t row_len = 0);
  void Wait(int ts);
  void SendCommand(int head, const std::string& body);  // to every server, waits for the replies
  void OnMessage(Message&& m);
  size_t bigarray_bound = 1000000;
  std::unordered_map<int, size_t> key_unit_;
  mutable std::mutex unit_mu_;

 private:
  struct Track { int expected = 0, received = 0; void* out = nullptr; int dtype = 0;
                 std::vector<Part> parts; bool row_sparse = false; size_t row_len = 0; };
  Node* node_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::unordered_map<int, Track> track_;
};

class KVServer {
 public:
  explicit KVServer(Node* node);
  void Run();  // blocks until kStopServer from every... (rank-0 worker sends it once)
  void OnMessage(Message&& m);
  using PyUpdater = void (*)(int key, const float* grad, float* weight, size_t n, void* user);
  using PyController = void (*)(int head, const char* body, size_t len, void* user);
  void SetPyUpdater(PyUpdater f, void* user) { py_updater_ = f; py_user_ = user; }
  void SetPyController(PyController f, void* user) { py_controller_ = f; py_ctl_user_ = user; }

 private:
  struct Entry {
    std::vector<float> store;      // fp32 master (multi-precision "store_realt_")
    std::vector<float> merged;
    std::vector<float> s1, s2;     // optimizer state
    std::vector<uint8_t> touched;  // row-sparse: rows seen this round
    std::vector<Message> waiting;  // queued push requests of this round
    bool initialized = false;
    long long t = 0;               // updates applied to this key (Adam / Adamax bias correction)
  };
  void HandleData(Message&& m);
  void HandleCommand(Message&& m);
  void ApplyUpdate(uint64_t key, Entry& e, size_t begin, size_t n);
  // Executor: python callbacks (pickled-optimizer controller, python updater) run on the thread that called
  // Run() - the interpreter's main thread - never on the transport's receiver thread (the reference's
  // single-threaded Executor, kvstore_dist_server.h:95-153).  Besides keeping python single-threaded this is what
  // lets a callback import modules while the server loop itself sits inside `import <package>.kv`.
  struct ExecTask { std::function<void()> fn; bool done = false; bool detached = false; };
  void ExecOnMain(const std::function<void()>& fn);      // blocks until run (inline when no loop is active)
  void PostToMain(std::function<void()> fn);             // fire-and-forget, order preserved
  std::mutex exec_mu_;
  std::condition_variable exec_cv_, exec_done_cv_;
  std::deque<ExecTask*> exec_q_;
  std::atomic<bool> running_{false}, exec_stop_{false};
  std::thread::id main_thread_;
  void Respond(const Message& req, const std::string* payload = nullptr);
  Node* node_;
  std::mutex mu_;
  std::condition_variable cv_;
  bool stop_ = false, sync_mode_ = false, multi_precision_ = false;
  float gc_threshold_ = 0.f;
  int gc_type_ = 0;
  OptimizerConfig opt_;
  std::unordered_map<uint64_t, Entry> store_;
  PyUpdater py_updater_ = nullptr; void* py_user_ = nullptr;
  PyController py_controller_ = nullptr; void* py_ctl_user_ = nullptr;
  std::string profiler_prefix_;
  // server-side profiler (reference kvstore_dist_server.h:276-323: kSetProfilerParams carries config / state / dump
  // requests from worker rank 0; the file name gets a rank<r>_ prefix).  Events = one Chrome-trace 'X' record per request.
  struct ProfEvent { std::string name; long long ts_us, dur_us; uint64_t key; size_t bytes; };
  void HandleProfilerCommand(const std::string& body);
  void ProfRecord(const char* name, long long t0_us, uint64_t key, size_t bytes);
  int DumpProfile();
  std::mutex prof_mu_;
  std::vector<ProfEvent> prof_events_;
  std::string prof_file_ = "profile.json";
  std::atomic<bool> prof_running_{false};
};

// conversions
void ToFloat(const void* src, int dtype, size_t n, float* dst);
void FromFloat(const float* src, int dtype, size_t n, void* dst);
size_t DTypeSize(int dtype);

}  // namespace dlcfn
