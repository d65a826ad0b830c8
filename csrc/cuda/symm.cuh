// Symmetric peer-HBM heap: the B200 replacement for the reference's EFS shared mount, the pinned
// host comm_buf_ of KVStoreDist (kvstore_dist.h:335-342) and ps-lite's ZeroMQ data plane (C6).
//
// Every rank owns one device allocation laid out as
//     [ signal pad 64 KiB | epoch counters 8 KiB | error word | 128 KiB: one-shot staging 512 KiB ... | 1 MiB: data heap ... ]
// and maps every peer's allocation (CUDA IPC across processes, or direct peer access inside one
// process).  Kernels receive the whole table by value (SymmDev) and address any peer's heap or
// signal pad with plain ld/st over NVLink 5 / NVSwitch.
#pragma once
#include "common.cuh"

namespace dlcfn {

constexpr int kSigChannels = 8;
constexpr int kSigMaxBlocks = 256;
constexpr size_t kSigBytes = (size_t)kSigChannels * kSigMaxBlocks * DLCFN_MAX_RANKS * sizeof(uint32_t);  // 64 KiB
constexpr size_t kEpochOff = kSigBytes;
constexpr size_t kEpochBytes = (size_t)kSigChannels * kSigMaxBlocks * sizeof(uint32_t);                  // 8 KiB
constexpr size_t kErrOff = kEpochOff + kEpochBytes;
// one-shot all-reduce staging: [2 parities][DLCFN_MAX_RANKS sources][kOneShotMax bytes] inside the header, below the heap
constexpr size_t kStageOff = 128u << 10;
constexpr size_t kOneShotMax = 16u << 10;
// flag-in-data ("LL") all-reduce slots: [2 parities][DLCFN_MAX_RANKS sources][2 * kLLMax bytes] (4 B payload + 4 B flag per word)
constexpr size_t kLLOff = kStageOff + 2 * DLCFN_MAX_RANKS * kOneShotMax;      // 384 KiB
constexpr size_t kLLMax = 16u << 10;
constexpr size_t kHeapOff = 1u << 20;
static_assert(kStageOff >= kErrOff + 64 && kLLOff + 2 * DLCFN_MAX_RANKS * 2 * kLLMax <= kHeapOff, "header layout");
constexpr long long kSpinTimeoutNs = 4000000000ll;  // 4 s: a dead peer becomes an error, not a hung GPU

struct SymmDev {
  char* base[DLCFN_MAX_RANKS];     // data heaps
  uint32_t* sig[DLCFN_MAX_RANKS];  // signal pads
  uint32_t* epoch;                 // local, per (channel, block)
  int* err;                        // local error word
  char* mc;                        // NVLS multicast mapping of the data heaps (nullptr when unavailable)
  int rank, world;
  // protocol test hooks (dlcfn_symm_set_debug; all zero / -1 in production):
  long long timeout_ns;            // spin timeout of the flag waits (0 = kSpinTimeoutNs)
  uint32_t stress_ns;              // > 0: every CTA sleeps a pseudo-random 0..stress_ns before it signals (schedule fuzzing)
  int drop_peer;                   // >= 0: this rank never signals that peer (fault injection: the peer must time out)
};

__device__ __forceinline__ uint32_t* sig_slot(const SymmDev& s, int owner, int channel, int block, int src) {
  return s.sig[owner] + ((size_t)(channel * kSigMaxBlocks + block) * DLCFN_MAX_RANKS + src);
}

// All ranks' block `blockIdx.x` rendezvous.  The release store is cumulative over the CTA's earlier
// writes (ordered by the bar.sync), the acquire load makes the peers' earlier writes visible.
__device__ __forceinline__ void block_barrier(const SymmDev& s, int channel, uint32_t val) {
  __syncthreads();
  if ((int)threadIdx.x < s.world) {
    if (s.stress_ns) {             // schedule fuzzing: a different delay per (rank, CTA, peer, round)
      uint32_t h = (uint32_t)s.rank * 0x9E3779B9u ^ (blockIdx.x * 0x85EBCA6Bu) ^ (threadIdx.x * 0xC2B2AE35u) ^ (val * 0x27D4EB2Fu);
      h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12;
      __nanosleep(h % s.stress_ns);
    }
    if ((int)threadIdx.x != s.drop_peer)
      st_release_sys(sig_slot(s, threadIdx.x, channel, blockIdx.x, s.rank), val);
    if (!spin_wait_ge(sig_slot(s, s.rank, channel, blockIdx.x, threadIdx.x), val, s.timeout_ns ? s.timeout_ns : kSpinTimeoutNs))
      atomicExch(s.err, 1);
  }
  __syncthreads();
}

__device__ __forceinline__ uint32_t epoch_load(const SymmDev& s, int channel) {
  return s.epoch[channel * kSigMaxBlocks + blockIdx.x];
}
__device__ __forceinline__ void epoch_store(const SymmDev& s, int channel, uint32_t v) {
  if (threadIdx.x == 0) s.epoch[channel * kSigMaxBlocks + blockIdx.x] = v;
}

// Host-side context (opaque to python; see symm.cu)
struct SymmCtx {
  int rank = 0, world = 1, device = 0;
  size_t heap_bytes = 0;
  char* local = nullptr;                  // whole allocation
  char* mc_alloc = nullptr;               // multicast mapping of the whole allocation (external backend only)
  bool external = false;                  // memory owned by the caller (torch symmetric memory)
  char* peer[DLCFN_MAX_RANKS] = {nullptr};  // whole allocations as mapped here
  bool opened_ipc[DLCFN_MAX_RANKS] = {false};
  SymmDev dev{};
  bool finalized = false;
};

}  // namespace dlcfn
