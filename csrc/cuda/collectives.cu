// Fused compute+collective kernels over the symmetric peer-HBM heap (sm_100a, NVLink 5 / NVSwitch).
//
//  kallreduce      bucket scale (1/world, loss-scale, ...) (+) two-shot all-reduce in ONE kernel:
//                  every rank reduces its 1/W slice by loading the slice from all peers, scales it in
//                  fp32 and stores the result straight into every peer's bucket.
//                  <- replaces Horovod's pack -> ncclAllReduce -> unpack (reference C11, run.sh:70-81)
//  kps_sync        parameter-server step in ONE kernel: push (server loads the gradient shard from
//                  every worker's HBM) (+) sum over workers (+) rescale/clip (+) SGD/momentum/Adam on the
//                  fp32 master copy (+) pull (server stores the new bf16 weights into every worker's HBM).
//                  <- replaces ZPush -> server merged+=recved -> updater -> ZPull -> Broadcast
//                     (reference C6/C7/C2: kvstore_dist.h:311-434, kvstore_dist_server.h:346-375,707-735)
//                  The per-key arrival counter `request.size()==NumWorkers` (kvstore_dist_server.h:348)
//                  becomes the device-side flag barrier at kernel entry.
//  kps_async_*     dist_async: the server applies each worker's push as it arrives, no cross-worker wait
//                  (kvstore_dist_server.h:348,715-723).
//  kbroadcast      root's bucket -> every peer (initial parameter broadcast; reference C2 ncclBcast).
//  kbarrier        device-side rendezvous (reference C8: Postoffice::Barrier through the scheduler).
//
// All of them are latency/NVLink-bound streaming kernels: 16-byte accesses, >= W independent loads in
// flight per thread, a small grid (default 32 CTAs x 512 threads ~ 4 MB in flight) so that the rest of
// the 148 SMs keep running backward while a bucket is on the wire.
#include <cstdlib>
#include "symm.cuh"
#include "optimizer.cuh"

namespace dlcfn {

template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x); f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z); f[3] = __uint_as_float(v.w);
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
  }
  static __device__ __forceinline__ uint4 pack(const float* f) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]),
                      pack_bf16(f[6], f[7]));
  }
};

// ------------------------------------------------------------- barrier ------
__global__ void kbarrier(SymmDev s, int channel) {
  uint32_t e = epoch_load(s, channel);
  block_barrier(s, channel, e + 1);
  epoch_store(s, channel, e + 1);
}

// ----------------------------------------------------------- all-reduce -----
// In place on heap[off, off + 16*nvec).  Slice r = vectors [r*per, (r+1)*per) is reduced by rank r.
template <typename T, int W>
__global__ void __launch_bounds__(512) kallreduce(SymmDev s, size_t off, size_t dst_off, size_t nvec, float scale,
                                                  int channel) {
  constexpr int N = Vec16<T>::N;
  const uint32_t e = epoch_load(s, channel);
  block_barrier(s, channel, e + 1);

  const size_t per = (nvec + W - 1) / W;
  const size_t lo = per * s.rank;
  const size_t hi = (lo + per < nvec) ? lo + per : nvec;
  uint4* buf[W];
  uint4* dst[W];
#pragma unroll
  for (int p = 0; p < W; ++p) {
    buf[p] = reinterpret_cast<uint4*>(s.base[p] + off);
    dst[p] = reinterpret_cast<uint4*>(s.base[p] + dst_off);
  }

  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < hi; v += stride) {
    uint4 x[W];
#pragma unroll
    for (int p = 0; p < W; ++p) x[p] = ld_relaxed_sys_v4(buf[p] + v);
    // fixed summation order (rank 0, 1, ..., W-1): run-to-run deterministic, and since exactly one
    // rank reduces each element every rank ends up with bit-identical results
    float acc[N], t[N];
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = 0.f;
#pragma unroll
    for (int p = 0; p < W; ++p) {
      Vec16<T>::unpack(x[p], t);
#pragma unroll
      for (int k = 0; k < N; ++k) acc[k] += t[k];
    }
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] *= scale;
    const uint4 o = Vec16<T>::pack(acc);
#pragma unroll
    for (int p = 0; p < W; ++p) st_stream_v4(dst[p] + v, o);
  }

  block_barrier(s, channel, e + 2);
  epoch_store(s, channel, e + 2);
}

// NVLS variant: the reduction happens INSIDE the NVSwitch.  One multimem.ld_reduce returns the sum over all
// ranks of a 16-byte vector (fp32 accumulation), one multimem.st writes the scaled result to every rank.
// Per GPU the wire carries 1/W of the bucket in each direction instead of (W-1)/W.
template <typename T>
__device__ __forceinline__ uint4 mc_ld_reduce(const void* p);
template <>
__device__ __forceinline__ uint4 mc_ld_reduce<__nv_bfloat16>(const void* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
template <>
__device__ __forceinline__ uint4 mc_ld_reduce<float>(const void* p) {
  uint4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(*reinterpret_cast<float*>(&r.x)), "=f"(*reinterpret_cast<float*>(&r.y)),
                 "=f"(*reinterpret_cast<float*>(&r.z)), "=f"(*reinterpret_cast<float*>(&r.w))
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void mc_st(void* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p),
               "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
               "f"(__uint_as_float(v.w))
               : "memory");
}

template <typename T>
__global__ void __launch_bounds__(512) kallreduce_mc(SymmDev s, size_t off, size_t dst_off, size_t nvec, float scale,
                                                     int channel) {
  constexpr int N = Vec16<T>::N;
  const uint32_t e = epoch_load(s, channel);
  block_barrier(s, channel, e + 1);
  const int W = s.world;
  const size_t per = (nvec + W - 1) / W;
  const size_t lo = per * s.rank;
  const size_t hi = (lo + per < nvec) ? lo + per : nvec;
  const uint4* src = reinterpret_cast<const uint4*>(s.mc + off);
  uint4* dst = reinterpret_cast<uint4*>(s.mc + dst_off);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t v = lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; v + 3 * stride < hi; v += 4 * stride) {   // four switch reductions in flight per thread
    uint4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = mc_ld_reduce<T>(src + v + u * stride);
    if (scale != 1.f) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float f[N];
        Vec16<T>::unpack(x[u], f);
#pragma unroll
        for (int k = 0; k < N; ++k) f[k] *= scale;
        x[u] = Vec16<T>::pack(f);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) mc_st(dst + v + u * stride, x[u]);
  }
  for (; v < hi; v += stride) {
    uint4 x0 = mc_ld_reduce<T>(src + v);
    if (scale != 1.f) {
      float f[N];
      Vec16<T>::unpack(x0, f);
#pragma unroll
      for (int k = 0; k < N; ++k) f[k] *= scale;
      x0 = Vec16<T>::pack(f);
    }
    mc_st(dst + v, x0);
  }
  block_barrier(s, channel, e + 2);
  epoch_store(s, channel, e + 2);
}

// ONE-SHOT variant for latency-bound sizes (<= kOneShotMax bytes): ONE flag barrier instead of two.
//   1. every rank PUSHES its input into the staging slot [parity][rank] of every peer's header (W stores per vector)
//   2. flag barrier (release after the pushes, acquire before the reads)
//   3. every rank sums the W staged copies out of its OWN memory in rank order (bit-identical on all ranks), scales, stores.
// No exit barrier: the staging slots alternate with the epoch's parity, and a rank can only reach the pushes of call k+2 after
// passing the barrier of call k+1, which every peer enters only after finishing its reads of call k (program order + release).
// In place is safe - the input is read by its owner only.  Reference analogue: the small-key path of CommDevice::Reduce
// (comm.h:598-616) where every key is one copy + one ElementwiseSum.
template <typename T, int W>
__global__ void __launch_bounds__(512) kallreduce_oneshot(SymmDev s, size_t off, size_t dst_off, size_t nvec, float scale,
                                                          int channel) {
  constexpr int N = Vec16<T>::N;
  const uint32_t e = epoch_load(s, channel);
  const size_t parity = e & 1u;
  const uint4* in = reinterpret_cast<const uint4*>(s.base[s.rank] + off);
  const size_t slot = (parity * DLCFN_MAX_RANKS + (size_t)s.rank) * kOneShotMax;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += stride) {
    const uint4 x = in[v];
#pragma unroll
    for (int p = 0; p < W; ++p)      // base[p] - kHeapOff = start of peer p's allocation (header)
      st_stream_v4(reinterpret_cast<uint4*>(s.base[p] - kHeapOff + kStageOff + slot) + v, x);
  }
  block_barrier(s, channel, e + 1);
  uint4* out = reinterpret_cast<uint4*>(s.base[s.rank] + dst_off);
  const char* mine = s.base[s.rank] - kHeapOff + kStageOff + parity * DLCFN_MAX_RANKS * kOneShotMax;
  for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += stride) {
    uint4 x[W];
#pragma unroll
    for (int p = 0; p < W; ++p) x[p] = ld_relaxed_sys_v4(reinterpret_cast<const uint4*>(mine + (size_t)p * kOneShotMax) + v);
    float acc[N], t[N];
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = 0.f;
#pragma unroll
    for (int p = 0; p < W; ++p) {
      Vec16<T>::unpack(x[p], t);
#pragma unroll
      for (int k = 0; k < N; ++k) acc[k] += t[k];
    }
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] *= scale;
    out[v] = Vec16<T>::pack(acc);
  }
  epoch_store(s, channel, e + 1);
}

// FLAG-IN-DATA variant ("LL", after NCCL's low-latency protocol) for the smallest messages (<= kLLMax bytes): NO barrier and NO
// memory fence at all.  Every 4-byte payload word travels in one 8-byte store together with the call's epoch; the receiver
// polls the 8-byte slot until the epoch matches - an 8-byte store is never torn, so a matching flag proves the payload.
//   1. every rank pushes (word, epoch) pairs into slot [parity][rank] of every peer (W 8-byte stores per word)
//   2. every rank polls its own W slots word by word, sums in rank order (bit-identical everywhere), scales, stores.
// Slots alternate with the epoch parity; the reuse argument is the one of kallreduce_oneshot (a rank reaches call k+2 only
// after it has received every peer's data of call k+1, which each peer sends after finishing call k), and a stale slot can
// never be mistaken for fresh data because its flag is an older epoch.  Epochs advance by 1 per call; 0 is never used.
template <typename T, int W>
__global__ void __launch_bounds__(512) kallreduce_ll(SymmDev s, size_t off, size_t dst_off, size_t nwords, float scale, int channel) {
  const uint32_t e = epoch_load(s, channel) + 1;
  const size_t parity = e & 1u;
  const uint32_t* in = reinterpret_cast<const uint32_t*>(s.base[s.rank] + off);
  const size_t slot = (parity * DLCFN_MAX_RANKS + (size_t)s.rank) * 2 * kLLMax;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += stride) {
    const unsigned long long pkt = ((unsigned long long)e << 32) | in[i];
#pragma unroll
    for (int p = 0; p < W; ++p)
      asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long*>(s.base[p] - kHeapOff + kLLOff + slot) + i), "l"(pkt) : "memory");
  }
  uint32_t* out = reinterpret_cast<uint32_t*>(s.base[s.rank] + dst_off);
  const char* mine = s.base[s.rank] - kHeapOff + kLLOff + parity * DLCFN_MAX_RANKS * 2 * kLLMax;
  const long long tmo = s.timeout_ns ? s.timeout_ns : kSpinTimeoutNs;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += stride) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int p = 0; p < W; ++p) {
      const unsigned long long* src = reinterpret_cast<const unsigned long long*>(mine + (size_t)p * 2 * kLLMax) + i;
      unsigned long long pkt;
      long long t0 = 0;
      int it = 0;
      for (;;) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(pkt) : "l"(src) : "memory");
        if ((uint32_t)(pkt >> 32) == e) break;
        if ((++it & 1023) == 0) {
          const long long now = clock_now();
          if (t0 == 0) t0 = now;
          else if (now - t0 > tmo) { atomicExch(s.err, 1); break; }
        }
      }
      const uint32_t w = (uint32_t)pkt;
      if (sizeof(T) == 4) a0 += __uint_as_float(w);
      else { a0 += bf16_lo(w); a1 += bf16_hi(w); }
    }
    out[i] = sizeof(T) == 4 ? __float_as_uint(a0 * scale) : pack_bf16(a0 * scale, a1 * scale);
  }
  __syncthreads();
  epoch_store(s, channel, e);
}

// ------------------------------------------------------------ broadcast -----
__global__ void __launch_bounds__(512) kbroadcast(SymmDev s, size_t off, size_t nvec, int root, int channel) {
  const uint32_t e = epoch_load(s, channel);
  block_barrier(s, channel, e + 1);
  if (s.rank == root) {
    const uint4* src = reinterpret_cast<const uint4*>(s.base[root] + off);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < nvec; v += stride) {
      const uint4 x = src[v];
      for (int p = 1; p < s.world; ++p)
        st_stream_v4(reinterpret_cast<uint4*>(s.base[(root + p) % s.world] + off) + v, x);
    }
  }
  block_barrier(s, channel, e + 2);
  epoch_store(s, channel, e + 2);
}

// ----------------------------------------------------- parameter server -----
struct PSArgs {
  size_t grad_off;       // heap byte offset of the flat gradient buffer (every worker, same offset)
  size_t weight_off;     // heap byte offset of the flat model-weight buffer
  size_t lo, hi;         // this rank's server shard in 16-byte vectors of T; lo == hi: not a server
  uint32_t worker_mask;  // ranks that push gradients and receive weights
  uint32_t pull_mask;    // ranks that receive the new weights (0: same as worker_mask)
  float* master;         // fp32 master weights of the shard (index 0 == element lo*N)
  float* s1;             // optimizer state 1 (momentum / Adam m) or null
  float* s2;             // optimizer state 2 (Adam v) or null
  OptArgs opt;
  int channel;
  int use_mc;            // 1: push = multimem.ld_reduce, pull = multimem.st (ranks that are not workers must hold
                         //    zero gradients: the switch sums over every rank)
};

// Peer pointer tables are indexed by rank with compile-time indices only, so they live in registers.
struct PSPtrs {
  const uint4* g[DLCFN_MAX_RANKS];  // gradient buffer of rank r (peer HBM)
  uint4* w[DLCFN_MAX_RANKS];        // weight buffer of rank r
  uint32_t load_mask, store_mask;
};

__device__ __forceinline__ void ps_make_ptrs(const PSArgs& a, const SymmDev& s, uint32_t worker_mask,
                                             PSPtrs& P) {
#pragma unroll
  for (int r = 0; r < DLCFN_MAX_RANKS; ++r) {
    P.g[r] = reinterpret_cast<const uint4*>(s.base[r] + a.grad_off);
    P.w[r] = reinterpret_cast<uint4*>(s.base[r] + a.weight_off);
  }
  const uint32_t live = (s.world >= 32) ? 0xffffffffu : ((1u << s.world) - 1u);
  P.load_mask = worker_mask & live;
  P.store_mask = ((a.pull_mask ? a.pull_mask : worker_mask) | (1u << s.rank)) & live;  // + the server's own copy
}

template <typename T>
__device__ __forceinline__ uint4 mc_ld_reduce(const void* p);
__device__ __forceinline__ void mc_st(void* p, const uint4& v);

template <typename T>
__device__ __forceinline__ void ps_update_vec(const PSArgs& a, const PSPtrs& P, size_t v, const char* mc = nullptr) {
  constexpr int N = Vec16<T>::N;
  float g[N], t[N];
  if (mc != nullptr) {
    // push through the switch: one load returns the sum over all ranks
    const uint4 xs = mc_ld_reduce<T>(mc + a.grad_off + v * 16);
    Vec16<T>::unpack(xs, g);
  } else {
    uint4 x[DLCFN_MAX_RANKS];
#pragma unroll
    for (int r = 0; r < DLCFN_MAX_RANKS; ++r)
      if ((P.load_mask >> r) & 1u) x[r] = ld_relaxed_sys_v4(P.g[r] + v);
#pragma unroll
    for (int k = 0; k < N; ++k) g[k] = 0.f;
    // fixed rank order: deterministic, independent of which server owns the shard
#pragma unroll
    for (int r = 0; r < DLCFN_MAX_RANKS; ++r)
      if ((P.load_mask >> r) & 1u) {
        Vec16<T>::unpack(x[r], t);
#pragma unroll
        for (int k = 0; k < N; ++k) g[k] += t[k];
      }
  }
  const size_t i0 = (v - a.lo) * N;
  float w[N], m1[N], m2[N];
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    float4 f = reinterpret_cast<const float4*>(a.master + i0)[q];
    w[4 * q] = f.x; w[4 * q + 1] = f.y; w[4 * q + 2] = f.z; w[4 * q + 3] = f.w;
    if (opt_uses_s1(a.opt.kind)) {
      float4 h = reinterpret_cast<const float4*>(a.s1 + i0)[q];
      m1[4 * q] = h.x; m1[4 * q + 1] = h.y; m1[4 * q + 2] = h.z; m1[4 * q + 3] = h.w;
    } else {
      m1[4 * q] = m1[4 * q + 1] = m1[4 * q + 2] = m1[4 * q + 3] = 0.f;
    }
    if (opt_uses_s2(a.opt.kind)) {
      float4 h = reinterpret_cast<const float4*>(a.s2 + i0)[q];
      m2[4 * q] = h.x; m2[4 * q + 1] = h.y; m2[4 * q + 2] = h.z; m2[4 * q + 3] = h.w;
    } else {
      m2[4 * q] = m2[4 * q + 1] = m2[4 * q + 2] = m2[4 * q + 3] = 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < N; ++k) opt_apply(a.opt, g[k], w[k], m1[k], m2[k]);
#pragma unroll
  for (int q = 0; q < N / 4; ++q) {
    reinterpret_cast<float4*>(a.master + i0)[q] = make_float4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    if (opt_uses_s1(a.opt.kind))
      reinterpret_cast<float4*>(a.s1 + i0)[q] = make_float4(m1[4 * q], m1[4 * q + 1], m1[4 * q + 2], m1[4 * q + 3]);
    if (opt_uses_s2(a.opt.kind))
      reinterpret_cast<float4*>(a.s2 + i0)[q] = make_float4(m2[4 * q], m2[4 * q + 1], m2[4 * q + 2], m2[4 * q + 3]);
  }
  const uint4 o = Vec16<T>::pack(w);
  // pull: the server writes the fresh weights into every worker's model buffer (and its own copy)
  if (mc != nullptr) {
    mc_st(const_cast<char*>(mc) + a.weight_off + v * 16, o);   // one store, the switch replicates it
    return;
  }
#pragma unroll
  for (int r = 0; r < DLCFN_MAX_RANKS; ++r)
    if ((P.store_mask >> r) & 1u) st_stream_v4(P.w[r] + v, o);
}

template <typename T>
__global__ void __launch_bounds__(512) kps_sync(SymmDev s, PSArgs a) {
  const uint32_t e = epoch_load(s, a.channel);
  block_barrier(s, a.channel, e + 1);  // every worker's gradients are complete ("all W pushes arrived")

  PSPtrs P;
  ps_make_ptrs(a, s, a.worker_mask, P);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const char* mc = a.use_mc ? s.mc : nullptr;
  for (size_t v = a.lo + blockIdx.x * (size_t)blockDim.x + threadIdx.x; v < a.hi; v += stride)
    ps_update_vec<T>(a, P, v, mc);

  block_barrier(s, a.channel, e + 2);  // every server finished writing: the pull is complete
  epoch_store(s, a.channel, e + 2);
}

// dist_async.  Roles are disjoint (a rank is a server or a worker).  Worker side: announce push #k to
// every server, then wait until every server has answered #k (its weights for this worker are written).
__global__ void kps_async_worker(SymmDev s, uint32_t server_mask, int channel) {
  const uint32_t k = epoch_load(s, channel) + 1;
  __syncthreads();
  if ((int)threadIdx.x < s.world && ((server_mask >> threadIdx.x) & 1u)) {
    st_release_sys(sig_slot(s, threadIdx.x, channel, blockIdx.x, s.rank), k);
    if (!spin_wait_ge(sig_slot(s, s.rank, channel + 1, blockIdx.x, threadIdx.x), k, s.timeout_ns ? s.timeout_ns : kSpinTimeoutNs))
      atomicExch(s.err, 1);
  }
  __syncthreads();
  epoch_store(s, channel, k);
}

// Server side: one launch serves exactly one push from each worker, in arrival order.  Each CTA owns a
// fixed slice of the shard, polls the workers' flags for that slice and applies w -= f(g_w) immediately.
template <typename T>
__global__ void __launch_bounds__(512) kps_async_server(SymmDev s, PSArgs a) {
  __shared__ int sh_next;
  __shared__ uint32_t sh_done;  // bitmask of workers served in this launch
  const uint32_t k = epoch_load(s, a.channel) + 1;
  if (threadIdx.x == 0) sh_done = 0;
  __syncthreads();
  int nworkers = __popc(a.worker_mask);
  const size_t span = a.hi - a.lo;
  const size_t per = (span + gridDim.x - 1) / gridDim.x;
  const size_t blo = a.lo + per * blockIdx.x;
  const size_t bhi = (blo + per < a.hi) ? blo + per : a.hi;
  for (int served = 0; served < nworkers; ++served) {
    if (threadIdx.x == 0) {
      int found = -1;
      long long t0 = clock_now();
      while (found < 0) {
        for (int r = 0; r < s.world && found < 0; ++r) {
          if (!((a.worker_mask >> r) & 1u) || ((sh_done >> r) & 1u)) continue;
          if ((int32_t)(ld_acquire_sys(sig_slot(s, s.rank, a.channel, blockIdx.x, r)) - k) >= 0) found = r;
        }
        if (found < 0 && clock_now() - t0 > (s.timeout_ns ? s.timeout_ns : kSpinTimeoutNs)) {
          atomicExch(s.err, 1);
          found = -2;
        }
      }
      sh_next = found;
      if (found >= 0) sh_done |= 1u << found;
    }
    __syncthreads();
    const int w = sh_next;
    if (w < 0) break;
    PSPtrs P;
    ps_make_ptrs(a, s, 1u << w, P);
    for (size_t v = blo + threadIdx.x; v < bhi; v += blockDim.x) ps_update_vec<T>(a, P, v);
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(sig_slot(s, w, a.channel + 1, blockIdx.x, s.rank), k);
  }
  __syncthreads();
  epoch_store(s, a.channel, k);
}

}  // namespace dlcfn

using namespace dlcfn;

// ============================================================ host API ======
DLCFN_API int dlcfn_symm_create(size_t heap_bytes, int rank, int world, void** out_ctx) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (world < 1 || world > DLCFN_MAX_RANKS || rank < 0 || rank >= world) return (int)cudaErrorInvalidValue;
  SymmCtx* c = new SymmCtx();
  c->rank = rank;
  c->world = world;
  c->heap_bytes = heap_bytes;
  DLCFN_CHECK(cudaGetDevice(&c->device));
  const size_t total = kHeapOff + ((heap_bytes + 255) & ~(size_t)255);
  void* p = nullptr;
  DLCFN_CHECK(cudaMalloc(&p, total));
  DLCFN_CHECK(cudaMemset(p, 0, kHeapOff));
  c->local = (char*)p;
  c->peer[rank] = c->local;
  *out_ctx = c;
  return 0;
}

// Adopt memory that somebody else allocated and mapped on every rank (torch symmetric memory): `local_alloc` is
// this rank's whole allocation (kHeapOff header + heap), peers are registered with dlcfn_symm_set_peer_ptr.
DLCFN_API int dlcfn_symm_create_external(void* local_alloc, size_t total_bytes, int rank, int world, void** out_ctx) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (world < 1 || world > DLCFN_MAX_RANKS || rank < 0 || rank >= world || total_bytes <= kHeapOff)
    return (int)cudaErrorInvalidValue;
  SymmCtx* c = new SymmCtx();
  c->rank = rank;
  c->world = world;
  c->heap_bytes = total_bytes - kHeapOff;
  c->external = true;
  DLCFN_CHECK(cudaGetDevice(&c->device));
  DLCFN_CHECK(cudaMemset(local_alloc, 0, kHeapOff));
  c->local = (char*)local_alloc;
  c->peer[rank] = c->local;
  *out_ctx = c;
  return 0;
}

DLCFN_API int dlcfn_symm_set_multicast(void* ctx, void* mc_alloc) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  ((SymmCtx*)ctx)->mc_alloc = (char*)mc_alloc;
  return 0;
}

// Protocol test hooks (tests/mp/protocol_worker.py): spin timeout in ms (0 = default 4 s), schedule-fuzzing delay bound in
// ns (0 = off), peer that this rank stops signalling (-1 = none; reference fault injection: PS_DROP_MSG, PS/src/van.cc:297-299).
DLCFN_API int dlcfn_symm_set_debug(void* ctx, int timeout_ms, unsigned stress_ns, int drop_peer) {
  SymmCtx* c = (SymmCtx*)ctx;
  c->dev.timeout_ns = (long long)timeout_ms * 1000000ll;
  c->dev.stress_ns = stress_ns;
  c->dev.drop_peer = drop_peer;
  return 0;
}

DLCFN_API int dlcfn_symm_has_multicast(void* ctx) { return ((SymmCtx*)ctx)->dev.mc != nullptr ? 1 : 0; }

DLCFN_API int dlcfn_symm_get_handle(void* ctx, void* out64) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  cudaIpcMemHandle_t h;
  DLCFN_CHECK(cudaIpcGetMemHandle(&h, c->local));
  static_assert(sizeof(h) == 64, "ipc handle size");
  memcpy(out64, &h, 64);
  return 0;
}

DLCFN_API int dlcfn_symm_open_peer(void* ctx, int peer, const void* handle64) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  if (peer < 0 || peer >= c->world || peer == c->rank) return (int)cudaErrorInvalidValue;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  DLCFN_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  c->peer[peer] = (char*)p;
  c->opened_ipc[peer] = true;
  return 0;
}

// single-process multi-GPU: the peer allocation is already addressable (peer access enabled by caller)
DLCFN_API int dlcfn_symm_set_peer_ptr(void* ctx, int peer, void* whole_allocation) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  if (peer < 0 || peer >= c->world) return (int)cudaErrorInvalidValue;
  c->peer[peer] = (char*)whole_allocation;
  return 0;
}

DLCFN_API void* dlcfn_symm_allocation(void* ctx) { return ((SymmCtx*)ctx)->local; }

DLCFN_API int dlcfn_symm_finalize(void* ctx) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  for (int p = 0; p < c->world; ++p)
    if (!c->peer[p]) return (int)cudaErrorInvalidValue;
  for (int p = 0; p < DLCFN_MAX_RANKS; ++p) {
    char* a = c->peer[p < c->world ? p : c->rank];
    c->dev.base[p] = a + kHeapOff;
    c->dev.sig[p] = reinterpret_cast<uint32_t*>(a);
  }
  c->dev.epoch = reinterpret_cast<uint32_t*>(c->local + kEpochOff);
  c->dev.err = reinterpret_cast<int*>(c->local + kErrOff);
  c->dev.mc = c->mc_alloc ? c->mc_alloc + kHeapOff : nullptr;
  c->dev.rank = c->rank;
  c->dev.world = c->world;
  c->dev.timeout_ns = 0;
  c->dev.stress_ns = 0;
  c->dev.drop_peer = -1;
  {   // environment form of the hooks, for whole-job stress runs: DLCFN_SYMM_STRESS_NS, DLCFN_SPIN_TIMEOUT_MS
    const char* e = getenv("DLCFN_SYMM_STRESS_NS");
    if (e) c->dev.stress_ns = (uint32_t)atoi(e);
    e = getenv("DLCFN_SPIN_TIMEOUT_MS");
    if (e && atoi(e) > 0) c->dev.timeout_ns = (long long)atoi(e) * 1000000ll;
  }
  c->finalized = true;
  return 0;
}

DLCFN_API void* dlcfn_symm_heap_ptr(void* ctx, int peer) {
  SymmCtx* c = (SymmCtx*)ctx;
  if (peer < 0 || peer >= c->world || !c->peer[peer]) return nullptr;
  return c->peer[peer] + kHeapOff;
}

// Reads and clears the device error word (set when a flag wait timed out).  Synchronises the stream.
DLCFN_API int dlcfn_symm_check_error(void* ctx, cudaStream_t stream, int* out_err) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  int e = 0;
  DLCFN_CHECK(cudaMemcpyAsync(&e, c->dev.err, sizeof(int), cudaMemcpyDeviceToHost, stream));
  DLCFN_CHECK(cudaStreamSynchronize(stream));
  if (e) DLCFN_CHECK(cudaMemsetAsync(c->dev.err, 0, sizeof(int), stream));
  *out_err = e;
  return 0;
}

// Non-blocking form: copies the error word into PINNED host memory on `stream` and returns; the caller reads it after the
// stream has passed this point (DataParallelTrainer polls it every few steps without draining the GPU queue).
DLCFN_API int dlcfn_symm_poll_error_async(void* ctx, int* host_pinned, cudaStream_t stream) {
  (void)cudaGetLastError();
  SymmCtx* c = (SymmCtx*)ctx;
  if (host_pinned == nullptr) return (int)cudaErrorInvalidValue;
  DLCFN_CHECK(cudaMemcpyAsync(host_pinned, c->dev.err, sizeof(int), cudaMemcpyDeviceToHost, stream));
  return 0;
}

DLCFN_API int dlcfn_symm_destroy(void* ctx) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  for (int p = 0; p < c->world; ++p)
    if (c->opened_ipc[p] && c->peer[p]) cudaIpcCloseMemHandle(c->peer[p]);
  if (c->local && !c->external) cudaFree(c->local);
  delete c;
  return 0;
}

// nblocks <= 0: pick the grid from the message size - the peer ld/st kernels need ~128 CTAs x 512 threads of
// 16-byte requests in flight to fill an NVLink-5 port (measured: 32 CTAs 246 GB/s, 128 CTAs 656 GB/s at 256 MiB,
// N = 2), small messages only pay for the barrier, so they get few CTAs.  `total_vecs` must be rank-independent.
static inline int auto_blocks(size_t total_vecs) {
  size_t b = total_vecs * 16 / (128 * 1024);
  if (b < 8) b = 8;
  if (b > 128) b = 128;
  return (int)b;
}

static inline int clamp_blocks(int nblocks, size_t work_vecs, int threads) {
  if (nblocks <= 0) nblocks = 32;
  if (nblocks > kSigMaxBlocks) nblocks = kSigMaxBlocks;
  size_t need = (work_vecs + threads - 1) / threads;
  if (need < 1) need = 1;
  if ((size_t)nblocks > need) nblocks = (int)need;
  return nblocks;
}

DLCFN_API int dlcfn_barrier(void* ctx, int channel, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  if (!c->finalized || channel < 0 || channel >= kSigChannels) return (int)cudaErrorInvalidValue;
  kbarrier<<<1, 32, 0, stream>>>(c->dev, channel);
  return (int)cudaGetLastError();
}

template <typename T>
static int launch_allreduce(SymmCtx* c, size_t off, size_t dst_off, size_t nvec, float scale, int channel, int nblocks,
                            int algo, cudaStream_t st) {
  // NOTE: every rank must compute the same grid, so the clamp only depends on (nvec, world)
  const size_t per = (nvec + c->world - 1) / c->world;
  if (nblocks <= 0) nblocks = auto_blocks(nvec);
  const int g = clamp_blocks(nblocks, per, 512);
  if (algo == 1) {
    if (c->dev.mc == nullptr) return (int)cudaErrorNotSupported;
    kallreduce_mc<T><<<g, 512, 0, st>>>(c->dev, off, dst_off, nvec, scale, channel);
    return (int)cudaGetLastError();
  }
  if (algo == 3) {
    if (nvec * 16 > kLLMax) return (int)cudaErrorInvalidValue;
    const size_t nwords = nvec * 4;
    // FIXED grid and a channel of its own: the slot parity comes from per-CTA epochs, which must advance in lockstep over all
    // CTAs and calls (a size-dependent grid would let a later call's CTA 0 reuse the parity of an unread slot of CTA 3)
    constexpr int g3 = (int)(kLLMax / 4 / 512);
    channel = kSigChannels - 1;
    switch (c->world) {
#define DLCFN_LL(Wn) case Wn: kallreduce_ll<T, Wn><<<g3, 512, 0, st>>>(c->dev, off, dst_off, nwords, scale, channel); break;
      DLCFN_LL(1) DLCFN_LL(2) DLCFN_LL(3) DLCFN_LL(4) DLCFN_LL(5) DLCFN_LL(6) DLCFN_LL(7) DLCFN_LL(8)
#undef DLCFN_LL
      default: return (int)cudaErrorInvalidValue;
    }
    return (int)cudaGetLastError();
  }
  if (algo == 2) {
    if (nvec * 16 > kOneShotMax) return (int)cudaErrorInvalidValue;
    constexpr int g1 = (int)(kOneShotMax / 16 / 512);  // fixed grid, own channel (see algo 3): per-CTA epochs stay in lockstep
    channel = kSigChannels - 2;
    switch (c->world) {
#define DLCFN_ONESHOT(Wn) case Wn: kallreduce_oneshot<T, Wn><<<g1, 512, 0, st>>>(c->dev, off, dst_off, nvec, scale, channel); break;
      DLCFN_ONESHOT(1) DLCFN_ONESHOT(2) DLCFN_ONESHOT(3) DLCFN_ONESHOT(4) DLCFN_ONESHOT(5) DLCFN_ONESHOT(6) DLCFN_ONESHOT(7) DLCFN_ONESHOT(8)
#undef DLCFN_ONESHOT
      default: return (int)cudaErrorInvalidValue;
    }
    return (int)cudaGetLastError();
  }
  switch (c->world) {
    case 1: kallreduce<T, 1><<<g, 512, 0, st>>>(c->dev, off, dst_off, nvec, scale, channel); break;
    case 2: kallreduce<T, 2><<<g, 512, 0, st>>>(c->dev, off, dst_off, nvec, scale, channel); break;
    case 3: kallreduce<T, 3><<<g, 512, 0, st>>>(c->dev, off, dst_off, nvec, scale, channel); break;
    case 4: kallreduce<T, 4><<<g, 512, 0, st>>>(c->dev, off, dst_off, nvec, scale, channel); break;
    case 5: kallreduce<T, 5><<<g, 512, 0, st>>>(c->dev, off, dst_off, nvec, scale, channel); break;
    case 6: kallreduce<T, 6><<<g, 512, 0, st>>>(c->dev, off, dst_off, nvec, scale, channel); break;
    case 7: kallreduce<T, 7><<<g, 512, 0, st>>>(c->dev, off, dst_off, nvec, scale, channel); break;
    case 8: kallreduce<T, 8><<<g, 512, 0, st>>>(c->dev, off, dst_off, nvec, scale, channel); break;
    default: return (int)cudaErrorInvalidValue;
  }
  return (int)cudaGetLastError();
}

// Fused scale + all-reduce of heap[src_off, +count*sizeof(T)) into heap[dst_off, ...) on every rank
// (dst_off == src_off: in place).  algo 0: peer ld/st two-shot;  algo 1: NVLS multimem (needs a multicast mapping);
// algo 2: one-shot through the header staging slots, <= 16 KiB (one flag barrier); algo 3: flag-in-data ("LL"), <= 16 KiB, no barrier.
DLCFN_API int dlcfn_allreduce_ex(void* ctx, size_t src_off, size_t dst_off, size_t count, int dtype, float scale,
                                 int channel, int nblocks, int algo, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  if (!c->finalized || (src_off & 15) || (dst_off & 15) || channel < 0 || channel >= kSigChannels)
    return (int)cudaErrorInvalidValue;
  if (count == 0) return 0;
  if (dtype == 0)
    return launch_allreduce<float>(c, src_off, dst_off, (count + 3) / 4, scale, channel, nblocks, algo, stream);
  if (dtype == 1)
    return launch_allreduce<__nv_bfloat16>(c, src_off, dst_off, (count + 7) / 8, scale, channel, nblocks, algo, stream);
  return (int)cudaErrorInvalidValue;
}

DLCFN_API int dlcfn_allreduce(void* ctx, size_t off_bytes, size_t count, int dtype, float scale, int channel,
                              int nblocks, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  return dlcfn_allreduce_ex(ctx, off_bytes, off_bytes, count, dtype, scale, channel, nblocks,
                            c->dev.mc != nullptr && c->world > 1 ? 1 : 0, stream);
}

DLCFN_API int dlcfn_broadcast(void* ctx, size_t off_bytes, size_t bytes, int root, int channel, int nblocks,
                              cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  if (!c->finalized || (off_bytes & 15) || root < 0 || root >= c->world) return (int)cudaErrorInvalidValue;
  if (bytes == 0) return 0;
  const size_t nvec = (bytes + 15) / 16;
  const int g = clamp_blocks(nblocks, nvec, 512);
  kbroadcast<<<g, 512, 0, stream>>>(c->dev, off_bytes, nvec, root, channel);
  return (int)cudaGetLastError();
}

// One synchronous parameter-server step, launched by EVERY rank (servers and workers alike).
// shard_lo/shard_hi are element indices (multiples of 16B/sizeof(T)) into the flat buffers.
// algo 1 routes the push through multimem.ld_reduce and the pull through multimem.st (needs a multicast mapping);
// pull_mask != 0 restricts the ranks that receive the updated weights (dedicated servers keep no model replica).
DLCFN_API int dlcfn_ps_sync_step_ex(void* ctx, size_t grad_off, size_t weight_off, size_t total_count,
                                    size_t shard_lo, size_t shard_hi, uint32_t worker_mask, uint32_t pull_mask,
                                    float* master, float* s1, float* s2, int dtype, int opt_kind, float lr, float wd,
                                    float momentum, float beta1, float beta2, float eps, float rescale,
                                    float clip, int channel, int nblocks, int algo, cudaStream_t stream);

DLCFN_API int dlcfn_ps_sync_step(void* ctx, size_t grad_off, size_t weight_off, size_t total_count,
                                 size_t shard_lo, size_t shard_hi, uint32_t worker_mask, float* master,
                                 float* s1, float* s2, int dtype, int opt_kind, float lr, float wd,
                                 float momentum, float beta1, float beta2, float eps, float rescale,
                                 float clip, int channel, int nblocks, cudaStream_t stream) {
  return dlcfn_ps_sync_step_ex(ctx, grad_off, weight_off, total_count, shard_lo, shard_hi, worker_mask, 0u, master, s1, s2,
                               dtype, opt_kind, lr, wd, momentum, beta1, beta2, eps, rescale, clip, channel, nblocks, 0, stream);
}

DLCFN_API int dlcfn_ps_sync_step_ex(void* ctx, size_t grad_off, size_t weight_off, size_t total_count,
                                    size_t shard_lo, size_t shard_hi, uint32_t worker_mask, uint32_t pull_mask,
                                    float* master, float* s1, float* s2, int dtype, int opt_kind, float lr, float wd,
                                    float momentum, float beta1, float beta2, float eps, float rescale,
                                    float clip, int channel, int nblocks, int algo, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  const int N = dtype == 1 ? 8 : 4;
  if (!c->finalized || (grad_off & 15) || (weight_off & 15) || (shard_lo % N) || dtype < 0 || dtype > 1)
    return (int)cudaErrorInvalidValue;
  if (worker_mask == 0 || channel < 0 || channel >= kSigChannels) return (int)cudaErrorInvalidValue;
  if (algo == 1 && c->dev.mc == nullptr) return (int)cudaErrorNotSupported;
  const int use_mc = algo == 1 ? 1 : 0;
  PSArgs a;
  a.grad_off = grad_off;
  a.weight_off = weight_off;
  a.lo = shard_lo / N;
  a.hi = (shard_hi + N - 1) / N;
  a.worker_mask = worker_mask;
  a.pull_mask = pull_mask;
  a.master = master;
  a.s1 = s1;
  a.s2 = s2;
  a.opt = OptArgs{opt_kind, lr, wd, momentum, beta1, beta2, eps, rescale, clip};
  a.channel = channel;
  a.use_mc = use_mc;
  // same grid on every rank: derived from the total size and world only
  const size_t per = ((total_count + N - 1) / N + c->world - 1) / c->world;
  if (nblocks <= 0) nblocks = auto_blocks((total_count + N - 1) / N);
  const int g = clamp_blocks(nblocks, per, 512);
  if (dtype == 1) kps_sync<__nv_bfloat16><<<g, 512, 0, stream>>>(c->dev, a);
  else kps_sync<float><<<g, 512, 0, stream>>>(c->dev, a);
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_ps_async_worker(void* ctx, uint32_t server_mask, int channel, int nblocks,
                                    cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  if (!c->finalized || channel < 0 || channel + 1 >= kSigChannels || nblocks < 1 || nblocks > kSigMaxBlocks)
    return (int)cudaErrorInvalidValue;
  kps_async_worker<<<nblocks, 32, 0, stream>>>(c->dev, server_mask, channel);
  return (int)cudaGetLastError();
}

DLCFN_API int dlcfn_ps_async_server(void* ctx, size_t grad_off, size_t weight_off, size_t shard_lo,
                                    size_t shard_hi, uint32_t worker_mask, float* master, float* s1,
                                    float* s2, int dtype, int opt_kind, float lr, float wd, float momentum,
                                    float beta1, float beta2, float eps, float rescale, float clip,
                                    int channel, int nblocks, cudaStream_t stream) {
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  SymmCtx* c = (SymmCtx*)ctx;
  const int N = dtype == 1 ? 8 : 4;
  if (!c->finalized || (grad_off & 15) || (weight_off & 15) || (shard_lo % N) || dtype < 0 || dtype > 1)
    return (int)cudaErrorInvalidValue;
  if (channel < 0 || channel + 1 >= kSigChannels || nblocks < 1 || nblocks > kSigMaxBlocks)
    return (int)cudaErrorInvalidValue;
  PSArgs a;
  a.grad_off = grad_off;
  a.weight_off = weight_off;
  a.lo = shard_lo / N;
  a.hi = (shard_hi + N - 1) / N;
  a.worker_mask = worker_mask;
  a.master = master;
  a.s1 = s1;
  a.s2 = s2;
  a.opt = OptArgs{opt_kind, lr, wd, momentum, beta1, beta2, eps, rescale, clip};
  a.channel = channel;
  a.use_mc = 0;
  a.pull_mask = 0;
  if (dtype == 1) kps_async_server<__nv_bfloat16><<<nblocks, 512, 0, stream>>>(c->dev, a);
  else kps_async_server<float><<<nblocks, 512, 0, stream>>>(c->dev, a);
  return (int)cudaGetLastError();
}
