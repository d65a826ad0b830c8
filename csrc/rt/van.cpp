// Node = Postoffice + Van + Resender of the dlcfn parameter-server transport (see ps.h).
#include "ps.h"
#include <new>

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

namespace dlcfn {

static const uint32_t kFrameMagic = 0xD1CF2026u;

static long long NowMs() {
  return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now().time_since_epoch())
      .count();
}
static const char* Env(const char* k, const char* dflt = nullptr) {
  const char* v = getenv(k);
  return v ? v : dflt;
}
static int EnvInt(const char* k, int dflt) {
  const char* v = getenv(k);
  return v ? atoi(v) : dflt;
}

// ------------------------------------------------------------------ serialization
namespace {
struct Writer {
  std::string buf;
  void i32(int32_t v) { buf.append(reinterpret_cast<const char*>(&v), 4); }
  void u64(uint64_t v) { buf.append(reinterpret_cast<const char*>(&v), 8); }
  void str(const std::string& s) { u64(s.size()); buf.append(s); }
};
struct Reader {
  const char* p; const char* end; bool ok = true;
  // lengths are compared with the bytes that are LEFT: a length field of a garbage frame may be any 64-bit value, and
  // `p + n` with such an n is pointer overflow (found by UBSAN under tests/test_kvstore.py: garbage connections)
  size_t left() const { return (size_t)(end - p); }
  int32_t i32() { int32_t v = 0; if (left() < 4) { ok = false; return 0; } memcpy(&v, p, 4); p += 4; return v; }
  uint64_t u64() { uint64_t v = 0; if (left() < 8) { ok = false; return 0; } memcpy(&v, p, 8); p += 8; return v; }
  std::string str() {
    uint64_t n = u64();
    if (!ok || n > (uint64_t)left()) { ok = false; return {}; }
    std::string s(p, (size_t)n); p += n; return s;
  }
};

std::string Pack(const Message& m) {
  Writer w;
  size_t total = 128 + m.body.size();                 // one allocation: a gradient blob must not be re-copied as the buffer grows
  for (const auto& n : m.nodes) total += 32 + n.host.size();
  for (const auto& d : m.data) total += 8 + d.size();
  w.buf.reserve(total);      // (a borrowed blob is not copied in: only its length is written, SendRaw sends the bytes)
  w.i32(m.sender); w.i32(m.recver); w.i32(m.control); w.i32(m.barrier_group); w.i32(m.app_id);
  w.i32(m.customer_id); w.i32(m.timestamp);
  w.i32((m.request ? 1 : 0) | (m.push ? 2 : 0) | (m.simple_app ? 4 : 0));
  w.i32(m.head); w.u64(m.sig); w.str(m.body);
  w.i32((int)m.nodes.size());
  for (const auto& n : m.nodes) { w.i32(n.role); w.i32(n.id); w.i32(n.port); w.i32(n.is_recovery ? 1 : 0); w.str(n.host); }
  w.i32((int)m.data.size() + (m.blob ? 1 : 0));
  for (const auto& d : m.data) w.str(d);
  if (m.blob) w.u64(m.blob_len);
  return std::move(w.buf);
}

bool Unpack(const char* begin, size_t size, Message* m) {
  Reader r{begin, begin + size};
  m->sender = r.i32(); m->recver = r.i32(); m->control = r.i32(); m->barrier_group = r.i32(); m->app_id = r.i32();
  m->customer_id = r.i32(); m->timestamp = r.i32();
  int flags = r.i32();
  m->request = flags & 1; m->push = flags & 2; m->simple_app = flags & 4;
  m->head = r.i32(); m->sig = r.u64(); m->body = r.str();
  int nn = r.i32();
  for (int i = 0; i < nn && r.ok; ++i) {
    NodeInfo n; n.role = r.i32(); n.id = r.i32(); n.port = r.i32(); n.is_recovery = r.i32() != 0; n.host = r.str();
    m->nodes.push_back(n);
  }
  int nd = r.i32();
  for (int i = 0; i < nd && r.ok; ++i) m->data.push_back(r.str());
  return r.ok;
}

bool SendAll(int fd, const char* p, size_t n) {
  while (n) {
    ssize_t k = ::send(fd, p, n, MSG_NOSIGNAL);
    if (k <= 0) return false;
    p += k; n -= (size_t)k;
  }
  return true;
}
bool RecvAll(int fd, char* p, size_t n) {
  while (n) {
    ssize_t k = ::recv(fd, p, n, 0);
    if (k <= 0) return false;
    p += k; n -= (size_t)k;
  }
  return true;
}
}  // namespace

// ------------------------------------------------------------------ Node
Node* Node::Get() {
  static Node* n = new Node();
  return n;
}

std::vector<int> Node::IdsOfGroup(int group) const {
  std::vector<int> ids;
  if (group & kSchedulerGroup) ids.push_back(kSchedulerId);
  if (group & kServerGroup) for (int r = 0; r < num_servers_; ++r) ids.push_back(ServerRankToId(r));
  if (group & kWorkerGroup) for (int r = 0; r < num_workers_; ++r) ids.push_back(WorkerRankToId(r));
  return ids;
}

std::pair<uint64_t, uint64_t> Node::ServerKeyRange(int r) const {
  const uint64_t kMax = UINT64_MAX;
  const uint64_t per = kMax / (uint64_t)std::max(1, num_servers_);
  return {per * (uint64_t)r, r == num_servers_ - 1 ? kMax : per * (uint64_t)(r + 1)};
}

int Node::Start(int) {
  if (started_) return 0;
  const char* role = Env("DMLC_ROLE", "worker");
  role_ = !strcmp(role, "scheduler") ? kScheduler : (!strcmp(role, "server") ? kServer : kWorker);
  num_workers_ = EnvInt("DMLC_NUM_WORKER", 1);
  num_servers_ = EnvInt("DMLC_NUM_SERVER", 1);
  verbose_ = EnvInt("PS_VERBOSE", 0);
  resend_ = EnvInt("PS_RESEND", 0) != 0;
  resend_timeout_ms_ = EnvInt("PS_RESEND_TIMEOUT", 1000);
  drop_pct_ = EnvInt("PS_DROP_MSG", 0);
  heartbeat_interval_ = EnvInt("PS_HEARTBEAT_INTERVAL", 0);
  scheduler_.role = kScheduler; scheduler_.id = kSchedulerId;
  scheduler_.host = Env("DMLC_PS_ROOT_URI", "127.0.0.1");
  scheduler_.port = EnvInt("DMLC_PS_ROOT_PORT", 9091);
  my_.role = role_;
  my_.host = role_ == kScheduler ? scheduler_.host : Env("DMLC_NODE_HOST", "127.0.0.1");

  listen_fd_ = ::socket(AF_INET, SOCK_STREAM, 0);
  int one = 1;
  setsockopt(listen_fd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
  sockaddr_in addr{};
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_ANY);
  addr.sin_port = htons(role_ == kScheduler ? scheduler_.port : 0);
  if (::bind(listen_fd_, (sockaddr*)&addr, sizeof(addr)) != 0) { perror("dlcfn ps bind"); return -1; }
  socklen_t len = sizeof(addr);
  getsockname(listen_fd_, (sockaddr*)&addr, &len);
  my_.port = ntohs(addr.sin_port);
  ::listen(listen_fd_, 256);
  stop_ = false;
  accept_thread_ = std::thread(&Node::AcceptLoop, this);

  {
    std::lock_guard<std::mutex> lk(mu_);
    nodes_[kSchedulerId] = scheduler_;
  }
  if (role_ == kScheduler) {
    my_.id = kSchedulerId;
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return ready_.load(); });
  } else {
    Message m;
    m.control = kAddNode; m.request = true; m.recver = kSchedulerId;
    m.nodes.push_back(my_);
    long long t0 = NowMs();
    while (Send(m) < 0) {   // the scheduler may not be listening yet
      if (NowMs() - t0 > 60000) { fprintf(stderr, "dlcfn ps: cannot reach scheduler %s:%d\n", scheduler_.host.c_str(), scheduler_.port); return -1; }
      usleep(50000);
    }
    // A node that (re)joins a running job is only admitted once the scheduler has declared a node of its role dead
    // (heartbeat timeout), so the registration is repeated until the node table arrives; the scheduler ignores
    // duplicates of a registration it already holds.
    std::unique_lock<std::mutex> lk(mu_);
    while (!cv_.wait_for(lk, std::chrono::seconds(2), [&] { return ready_.load(); })) {
      lk.unlock();
      Message again;
      again.control = kAddNode; again.request = true; again.recver = kSchedulerId;
      again.nodes.push_back(my_);
      Send(again);
      lk.lock();
    }
  }
  if (heartbeat_interval_ > 0 && role_ != kScheduler) heartbeat_thread_ = std::thread(&Node::HeartbeatLoop, this);
  if (resend_) resend_thread_ = std::thread(&Node::ResendLoop, this);
  started_ = true;
  if (verbose_) fprintf(stderr, "[dlcfn ps] node %d (role %d) ready at %s:%d\n", my_.id, role_, my_.host.c_str(), my_.port);
  return 0;
}

void Node::Finalize(bool do_barrier) {
  if (!started_) return;
  if (do_barrier) Barrier(kSchedulerGroup | kServerGroup | kWorkerGroup);
  stop_ = true;
  ::shutdown(listen_fd_, SHUT_RDWR);
  ::close(listen_fd_);
  // Outgoing connections: a reader thread may be inside send() on one of them right now (ACKs, barrier replies).
  // Pass 1 only SHUTS the sockets down - safe against a concurrent send(), and it fails a sender that is blocked on a
  // peer that stopped reading.  Pass 2 closes each descriptor while holding that peer's send mutex, so the number cannot
  // be recycled under a sender (found by ThreadSanitizer: close() here raced with send() in ReadLoop -> Send).
  std::vector<std::pair<int, std::mutex*>> peers;
  {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& kv : send_fd_) ::shutdown(kv.second, SHUT_RDWR);
    for (auto& kv : send_mu_) peers.emplace_back(kv.first, kv.second.get());
  }
  for (auto& pm : peers) {
    std::lock_guard<std::mutex> ls(*pm.second);
    std::lock_guard<std::mutex> lk(mu_);
    auto it = send_fd_.find(pm.first);
    if (it != send_fd_.end()) { ::close(it->second); send_fd_.erase(it); }
  }
  if (accept_thread_.joinable()) accept_thread_.join();
  if (heartbeat_thread_.joinable()) heartbeat_thread_.join();
  if (resend_thread_.joinable()) resend_thread_.join();
  for (auto& t : readers_) if (t.joinable()) t.detach();  // blocked in recv on peers that close later
  readers_.clear();
  started_ = false;
  ready_ = false;
}

int Node::ConnectTo(const NodeInfo& n) {
  addrinfo hints{}, *res = nullptr;
  hints.ai_family = AF_INET; hints.ai_socktype = SOCK_STREAM;
  if (getaddrinfo(n.host.c_str(), std::to_string(n.port).c_str(), &hints, &res) != 0 || !res) return -1;
  int fd = ::socket(res->ai_family, res->ai_socktype, res->ai_protocol);
  int rc = ::connect(fd, res->ai_addr, res->ai_addrlen);
  freeaddrinfo(res);
  if (rc != 0) { ::close(fd); return -1; }
  int one = 1;
  setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
  return fd;
}

// Called with the peer's send mutex held (Send): the only place that opens, replaces or closes a peer's descriptor.
int Node::FdFor(int id) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = send_fd_.find(id);
  if (it != send_fd_.end() && stale_fd_.erase(id)) {     // the id was handed to a restarted node: drop the old connection
    ::close(it->second);
    send_fd_.erase(it);
    it = send_fd_.end();
  }
  if (it != send_fd_.end()) return it->second;
  auto n = nodes_.find(id);
  if (n == nodes_.end()) return -1;
  int fd = ConnectTo(n->second);
  if (fd < 0) return -1;
  send_fd_[id] = fd;
  return fd;
}

bool Node::SendRaw(int fd, const Message& m) {
  const std::string payload = Pack(m);
  char hdr[12];
  const uint32_t magic = kFrameMagic;
  const uint64_t n = payload.size() + (m.blob ? m.blob_len : 0);
  memcpy(hdr, &magic, 4);
  memcpy(hdr + 4, &n, 8);
  // header and payload go out as two writes under the peer's send mutex (MSG_MORE keeps them in one segment train)
  // instead of being copied into a third buffer first
  ssize_t k = ::send(fd, hdr, 12, MSG_NOSIGNAL | MSG_MORE);
  if (k != 12) return false;
  if (!SendAll(fd, payload.data(), payload.size())) return false;
  if (m.blob && m.blob_len && !SendAll(fd, m.blob, m.blob_len)) return false;
  bytes_sent_ += (long long)n + 12;
  return true;
}

int Node::Send(Message& msg) {
  msg.sender = my_.id;
  if (resend_ && msg.blob) {            // the resender keeps messages for retransmission: own the bytes
    msg.data.emplace_back(msg.blob, msg.blob_len);
    msg.blob = nullptr;
    msg.blob_len = 0;
  }
  if (resend_ && msg.control == kNone && msg.sig == 0) {
    msg.sig = ((uint64_t)(uint32_t)my_.id << 40) | sig_counter_++;
    std::lock_guard<std::mutex> lk(rs_mu_);
    unacked_[msg.sig] = Pending{msg, NowMs(), 0};
  }
  std::mutex* m = nullptr;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto& slot = send_mu_[msg.recver];
    if (!slot) slot.reset(new std::mutex());
    m = slot.get();
  }
  // one sender per peer at a time, and the descriptor is resolved INSIDE that critical section: nobody can close or
  // replace it between the lookup and the send (lock order: peer send mutex -> mu_)
  std::lock_guard<std::mutex> lk(*m);
  if (stop_) return -1;                              // late ACK / reply from a detached reader after Finalize
  int fd = FdFor(msg.recver);
  if (fd < 0) return -1;
  if (!SendRaw(fd, msg)) {
    std::lock_guard<std::mutex> lk2(mu_);
    ::close(fd);
    send_fd_.erase(msg.recver);
    return -1;
  }
  return 1;
}

void Node::AcceptLoop() {
  while (!stop_) {
    sockaddr_in peer{};
    socklen_t len = sizeof(peer);
    int fd = ::accept(listen_fd_, (sockaddr*)&peer, &len);
    if (fd < 0) {
      if (stop_) return;
      continue;
    }
    int one = 1;
    setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
    std::lock_guard<std::mutex> lk(mu_);
    readers_.emplace_back(&Node::ReadLoop, this, fd);
  }
}

void Node::ReadLoop(int fd) {
  std::mt19937 rng((unsigned)NowMs() ^ (unsigned)fd);
  while (!stop_) {
    char hdr[12];
    if (!RecvAll(fd, hdr, 12)) break;
    uint32_t magic; uint64_t n;
    memcpy(&magic, hdr, 4); memcpy(&n, hdr + 4, 8);
    if (magic != kFrameMagic || n > (1ull << 34)) break;
    std::unique_ptr<char[]> payload(new (std::nothrow) char[n ? n : 1]);     // not zero-filled: recv overwrites all of it
    if (!payload) break;
    if (!RecvAll(fd, payload.get(), n)) break;
    bytes_recv_ += (long long)n + 12;
    Message m;
    if (!Unpack(payload.get(), n, &m)) continue;
    if (role_ == kScheduler && m.sender > 0) {
      std::lock_guard<std::mutex> lk(mu_);
      last_seen_[m.sender] = NowMs() / 1000;
    }
    if (m.control != kNone) {
      HandleControl(std::move(m), fd);
      continue;
    }
    // fault injection: drop received data messages (reference PS_DROP_MSG, van.cc:297-299)
    if (drop_pct_ > 0 && ready_ && (int)(rng() % 100) < drop_pct_) {
      if (verbose_ > 1) fprintf(stderr, "[dlcfn ps] drop message from %d ts %d\n", m.sender, m.timestamp);
      continue;
    }
    if (m.sig != 0) {
      Message ack;
      ack.control = kAck; ack.recver = m.sender; ack.sig = m.sig;
      Send(ack);
      std::lock_guard<std::mutex> lk(rs_mu_);
      if (!seen_sig_.insert(m.sig).second) continue;  // duplicate
    }
    if (data_handler_) data_handler_(std::move(m));
  }
  ::close(fd);
}

void Node::SchedulerAddNode(Message&& m) {
  // called with mu_ held
  NodeInfo n = m.nodes[0];
  if (!ready_) {
    for (const auto& p : pending_)
      if (p.host == n.host && p.port == n.port && p.role == n.role) return;   // repeated registration
    pending_.push_back(n);
    if ((int)pending_.size() < num_workers_ + num_servers_) return;
    // all registered: deterministic rank assignment in host:port order (van.cc:37-85)
    std::sort(pending_.begin(), pending_.end(), [](const NodeInfo& a, const NodeInfo& b) {
      return a.host != b.host ? a.host < b.host : a.port < b.port;
    });
    int ws = 0, ss = 0;
    for (auto& p : pending_) {
      p.id = p.role == kServer ? ServerRankToId(ss++) : WorkerRankToId(ws++);
      nodes_[p.id] = p;
      last_seen_[p.id] = NowMs() / 1000;
    }
    Message table;
    table.control = kNodeTable;
    for (auto& kv : nodes_) table.nodes.push_back(kv.second);
    ready_ = true;
    std::vector<NodeInfo> targets = pending_;
    mu_.unlock();
    for (auto& p : targets) {
      Message t = table;
      t.recver = p.id;
      Send(t);
    }
    mu_.lock();
    cv_.notify_all();
    return;
  }
  // late registration: a restarted node takes over the id of a dead node of the same role (van.cc:86-156)
  for (const auto& kv : nodes_)
    if (kv.second.host == n.host && kv.second.port == n.port && kv.second.role == n.role) {
      // already admitted (its table may have been lost / it re-sent before the table arrived): send the table again
      Message table;
      table.control = kNodeTable;
      for (auto& e : nodes_) table.nodes.push_back(e.second);
      table.recver = kv.first;
      mu_.unlock();
      Send(table);
      mu_.lock();
      return;
    }
  const long long now = NowMs() / 1000;
  const int timeout = EnvInt("PS_HEARTBEAT_TIMEOUT", 30);
  int take = -1;
  for (auto& kv : nodes_) {
    if (kv.first == kSchedulerId || kv.second.role != n.role) continue;
    if (now - last_seen_[kv.first] > timeout) { take = kv.first; break; }
  }
  if (take < 0) return;
  n.id = take; n.is_recovery = true;
  nodes_[take] = n;
  last_seen_[take] = now;
  auto it = send_fd_.find(take);
  if (it != send_fd_.end()) { ::shutdown(it->second, SHUT_RDWR); stale_fd_.insert(take); }   // closed by the next Send (FdFor)
  Message table;
  table.control = kNodeTable;
  for (auto& kv : nodes_) table.nodes.push_back(kv.second);
  std::vector<int> ids;
  for (auto& kv : nodes_) if (kv.first != kSchedulerId) ids.push_back(kv.first);
  mu_.unlock();
  for (int id : ids) {
    Message t = table;
    t.recver = id;
    Send(t);
  }
  mu_.lock();
}

void Node::HandleControl(Message&& m, int) {
  switch (m.control) {
    case kAddNode: {
      if (role_ != kScheduler || m.nodes.empty()) return;
      std::lock_guard<std::mutex> lk(mu_);
      SchedulerAddNode(std::move(m));
      return;
    }
    case kNodeTable: {
      std::lock_guard<std::mutex> lk(mu_);
      for (auto& n : m.nodes) {
        auto old = nodes_.find(n.id);
        if (old != nodes_.end() && (old->second.host != n.host || old->second.port != n.port)) {
          auto it = send_fd_.find(n.id);   // peer moved (recovery): the next Send replaces the connection (FdFor)
          if (it != send_fd_.end()) { ::shutdown(it->second, SHUT_RDWR); stale_fd_.insert(n.id); }
        }
        nodes_[n.id] = n;
        if (my_.id < 0 && n.host == my_.host && n.port == my_.port && n.role == my_.role) my_ = n;
      }
      ready_ = true;
      cv_.notify_all();
      return;
    }
    case kBarrier: {
      if (m.request) {
        if (role_ != kScheduler) return;
        std::vector<int> ids;
        {
          std::lock_guard<std::mutex> lk(mu_);
          int& c = barrier_count_[m.barrier_group];
          ++c;
          if (c < (int)IdsOfGroup(m.barrier_group).size()) return;
          c = 0;
          ids = IdsOfGroup(m.barrier_group);
        }
        // Everybody else is released BEFORE the scheduler's own main thread: at the closing barrier that thread goes on
        // to stop the node and close its connections, and a release still waiting to be sent from this reader thread
        // would be lost - its receiver would sit in the barrier forever (the close() / send() race ThreadSanitizer reported).
        bool self = false;
        for (int id : ids) {
          if (id == kSchedulerId) { self = true; continue; }
          Message r;
          r.control = kBarrier; r.request = false; r.recver = id; r.barrier_group = m.barrier_group;
          Send(r);
        }
        if (self) {
          std::lock_guard<std::mutex> lk(mu_);
          barrier_done_ = true;
          cv_.notify_all();
        }
      } else {
        std::lock_guard<std::mutex> lk(mu_);
        barrier_done_ = true;
        cv_.notify_all();
      }
      return;
    }
    case kHeartbeat: {
      if (role_ == kScheduler && m.request && !m.body.empty()) {
        // dead-node query "group:timeout"
        int group = 0, timeout = 0;
        sscanf(m.body.c_str(), "%d:%d", &group, &timeout);
        int dead = 0;
        if (heartbeat_interval_ > 0) {   // without heartbeats nobody can be declared dead (ps-lite behaviour)
          std::lock_guard<std::mutex> lk(mu_);
          const long long now = NowMs() / 1000;
          for (int id : IdsOfGroup(group))
            if (id != kSchedulerId && now - last_seen_[id] > timeout) ++dead;
        }
        Message r;
        r.control = kHeartbeat; r.request = false; r.recver = m.sender; r.body = std::to_string(dead);
        r.timestamp = m.timestamp;
        Send(r);
      } else if (!m.request) {
        std::lock_guard<std::mutex> lk(mu_);
        last_seen_[-m.timestamp - 1] = atoi(m.body.c_str());  // reply slot keyed by timestamp
        cv_.notify_all();
      }
      return;
    }
    case kAck: {
      std::lock_guard<std::mutex> lk(rs_mu_);
      unacked_.erase(m.sig);
      return;
    }
    default:
      return;
  }
}

void Node::Barrier(int group) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    barrier_done_ = false;
  }
  if (role_ == kScheduler) {
    Message m;
    m.control = kBarrier; m.request = true; m.barrier_group = group; m.sender = kSchedulerId;
    HandleControl(std::move(m), -1);
  } else {
    Message m;
    m.control = kBarrier; m.request = true; m.barrier_group = group; m.recver = kSchedulerId;
    Send(m);
  }
  std::unique_lock<std::mutex> lk(mu_);
  cv_.wait(lk, [&] { return barrier_done_; });
  barrier_done_ = false;
}

int Node::NumDeadNodes(int group, int timeout_sec) {
  if (role_ == kScheduler) {
    std::lock_guard<std::mutex> lk(mu_);
    int dead = 0;
    if (heartbeat_interval_ <= 0) return 0;
    const long long now = NowMs() / 1000;
    for (int id : IdsOfGroup(group))
      if (id != kSchedulerId && now - last_seen_[id] > timeout_sec) ++dead;
    return dead;
  }
  Message m;
  m.control = kHeartbeat; m.request = true; m.recver = kSchedulerId;
  m.timestamp = NextTimestamp();
  m.body = std::to_string(group) + ":" + std::to_string(timeout_sec);
  const int slot = -m.timestamp - 1;
  {
    std::lock_guard<std::mutex> lk(mu_);
    last_seen_.erase(slot);
  }
  if (Send(m) < 0) return -1;
  std::unique_lock<std::mutex> lk(mu_);
  if (!cv_.wait_for(lk, std::chrono::seconds(10), [&] { return last_seen_.count(slot) > 0; })) return -1;
  int dead = (int)last_seen_[slot];
  last_seen_.erase(slot);
  return dead;
}

void Node::HeartbeatLoop() {
  while (!stop_) {
    for (int i = 0; i < heartbeat_interval_ * 10 && !stop_; ++i) usleep(100000);
    if (stop_) break;
    Message m;
    m.control = kHeartbeat; m.request = true; m.recver = kSchedulerId;
    Send(m);
  }
}

void Node::ResendLoop() {
  while (!stop_) {
    usleep((useconds_t)std::max(1, resend_timeout_ms_ / 2) * 1000);
    std::vector<Message> again;
    {
      std::lock_guard<std::mutex> lk(rs_mu_);
      const long long now = NowMs();
      for (auto it = unacked_.begin(); it != unacked_.end();) {
        if (now - it->second.sent_ms > resend_timeout_ms_) {
          if (++it->second.tries > 10) {
            fprintf(stderr, "[dlcfn ps] giving up on message sig %llu after 10 resends\n", (unsigned long long)it->first);
            it = unacked_.erase(it);
            continue;
          }
          it->second.sent_ms = now;
          again.push_back(it->second.msg);
        }
        ++it;
      }
    }
    for (auto& m : again) {
      if (verbose_ > 1) fprintf(stderr, "[dlcfn ps] resend sig %llu to %d\n", (unsigned long long)m.sig, m.recver);
      Send(m);
    }
  }
}

// Test hook (tests/test_kvstore.py: message fuzzing under the sanitizers): decode one frame payload exactly as ReadLoop does and
// re-encode what was understood.  Returns 0 for an undecodable payload, otherwise 1 + the number of node / data entries.
int DebugRoundTrip(const char* payload, size_t n) {
  Message m;
  if (!Unpack(payload, n, &m)) return 0;
  const std::string again = Pack(m);
  return 1 + (int)m.nodes.size() + (int)m.data.size() + (again.empty() ? 1 : 0);
}

}  // namespace dlcfn

extern "C" __attribute__((visibility("default"))) int dlcfn_rt_debug_unpack(const char* payload, size_t n) {
  return dlcfn::DebugRoundTrip(payload, n);
}
