// tcgen05 / TMEM / TMA GEMM for sm_100a, with the parameter "pull" fused into the operand loads.
//
//   D[M,N] (bf16|fp32) = A[M,K] * B[N,K]^T (+ bias[N]) (+ D)        fp32 accumulation in TMEM
//
//  * A and B are bf16.  Each may be K-major (row-major [rows,K]) or MN-major (stored [K,rows]); the
//    MN-major forms give dgrad (dX = dY * W) and wgrad (dW = dY^T * X) without materialising transposes.
//  * B may be ROW-SHARDED ACROSS RANKS in the symmetric peer-HBM heap: the producer warp then issues
//    its TMA loads against the owning peer's tensor map, i.e. the weight "pull"/all-gather of the
//    parameter-server path (reference C2: ncclBcast root->all, kvstore_nccl.h:380-396; KVStoreDist
//    PullImpl + Broadcast, kvstore_dist.h:217-269) happens inside the first GEMM that consumes the
//    parameters - tile by tile over NVLink, overlapped with the tcgen05 math - and the gathered weight
//    is never materialised in local HBM.
//
// Kernel shape (one CTA per SM, persistent over output tiles, 192 threads):
//   warp 0      TMA producer   cp.async.bulk.tensor -> 128B-swizzled smem ring (mbarrier full/empty)
//   warp 1      MMA issuer     one elected lane issues tcgen05.mma (UMMA 128 x BLOCK_N x 16), commits
//                              to the smem-empty and tmem-full mbarriers; owns TMEM alloc/dealloc
//   warps 2..9  epilogue       tcgen05.ld 32x32b -> registers -> bias/relu/convert -> 128B-swizzled smem
//                              staging tile -> TMA store (cp.async.bulk.tensor, full 128-byte lines, M/N tails
//                              clipped by the tensor map); fp32 / accumulate / split-K outputs use direct 16B stores;
//                              two warps per TMEM lane quarter (each takes half of the tile's columns and
//                              issues all of its tcgen05.ld before one wait), two TMEM accumulator stages so
//                              the epilogue of tile i overlaps the mainloop of tile i+1
//
// Split-K: when the output has fewer tiles than SMs but a long reduction (wgrad of a convolution:
// dW[Cout,Cin] = dY[M,Cout]^T X[M,Cin] with M ~ 10^5..10^6) the K loop is cut into `splits` slices, every
// (tile, slice) pair is one work item of the persistent grid and writes an fp32 partial tile to a
// stream-ordered workspace; ksplitk_reduce sums the slices in a fixed order (deterministic) and applies
// bias / accumulate / relu / the output conversion.
#include "symm.cuh"
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace dlcfn {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;  // 64 bf16 = 128 B = one swizzle-128B atom row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 320;   // TMA warp + MMA warp + 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kMaxShards = DLCFN_MAX_RANKS;

// Division by a launch-time constant as multiply-high + shift (0 <= n < 2^31): the single-thread TMA producer and MMA issuer
// sit on the critical path of every k-block, and a hardware-less 32-bit division costs them ~150 dependent cycles each
// (six of them per k-block held the im2col kernels at ~650 cycles per k-block, profiles/r2/conv_ncu_before.md).
struct FastDiv {
  uint32_t mul, shr, d;
};
static inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d < 1 ? 1u : (uint32_t)d;
  uint32_t l = 0;
  while ((1ull << l) < f.d) ++l;
  f.shr = 31 + l;
  f.mul = (uint32_t)(((1ull << f.shr) / f.d) + 1ull);       // <= 2^32 - ... fits: d >= 2^(l-1) + 1 for l >= 1, and 2^31 + 1 for d = 1
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv& f) {
  return (int)(((unsigned long long)(uint32_t)n * f.mul) >> f.shr);
}
__device__ __forceinline__ void fdivmod(int n, const FastDiv& f, int& q, int& r) {
  q = fdiv(n, f);
  r = n - q * (int)f.d;
}

struct __align__(64) GemmParams {
  CUtensorMap tmA;
  CUtensorMap tmB[kMaxShards];
  CUtensorMap tmD;      // bf16 output, box {64 cols, 32 rows}, 128B swizzle (valid when tma_store != 0)
  int tma_store;
  int M, N, K;
  int n_shards;         // 1: B is local.  >1: B rows sharded over n_shards peers
  int rows_per_shard;   // B rows (or, for MN-major B: contraction rows... see host) per shard
  void* D;
  long long ldd;        // row stride of D in elements
  const float* bias;    // nullable, length N
  int d_is_fp32;
  int accumulate;       // D += result
  int relu;
  int splits;           // >1: split-K, partial tiles go to ws[split][M][N] (fp32)
  int kb_per_split;     // k-blocks (of kBlockK) per split
  float* ws;
  int split_k_ok;       // host-side switch (DLCFN_GEMM_SPLITK=0 disables)
  // implicit-GEMM convolution (kgemm<..., CONV = true>): one operand is an NHWC activation tensor read through an
  // im2col-mode tensor map (tmA for fprop / dgrad, tmB[0] for wgrad)
  int cv_Q, cv_P;       // base positions per row / rows per image of the im2col traversal (= the output width / height)
  int cv_cblocks;       // channels of the im2col'd tensor / 64
  int cv_S, cv_RS;      // filter width, taps
  int cv_lower_w, cv_lower_h;   // bounding-box lower corner: -pad (fprop, wgrad), pad - (S-1) (dgrad)
  int cv_stride;        // traversal stride (convolution stride; 1 for dgrad)
  int cv_cin;           // dgrad: columns per tap of the weight matrix [Cout][R*S*Cin]
  // filter-tap table: tap t of the reduction reads the im2col'd tensor displaced by (cv_tap_ow[t], cv_tap_oh[t]) from the
  // base position; dgrad multiplies it with weight tap cv_tap_wi[t] (fprop / wgrad walk the taps in storage order).  A
  // strided dgrad is decomposed by output parity: every class has its own (shorter) tap list.
  int cv_ntaps;
  unsigned char cv_tap_ow[52], cv_tap_oh[52], cv_tap_wi[52];
  // output-row remap (strided dgrad): GEMM row m = (n, i, j) of the [cv_P x cv_Q] class grid is pixel
  // (i * out_step + out_h0, j * out_step + out_w0) of the [out_H x out_W] gradient image
  int out_remap, out_H, out_W, out_step, out_h0, out_w0;
  // pipeline shape chosen by the host: ring depth, weights-stationary mode (B panel of bres_bytes resident in shared memory)
  int stages, b_resident, bres_bytes;
  int raster_g;         // pair kernel: M-pair rows per raster group (tiles walk N inside a group of raster_g rows: the group's
                        // A panel stays L2-resident while B streams once per group instead of once per wave)
  int stg_bufs;         // staging tiles per epilogue warp: 2 = a warp fills one tile while TMA still reads the other
  // fused BatchNorm statistics of the stored bf16 outputs: partial rows [4 * grid][2][N] (sum, sum of squares), or null
  float* stats;
  // STAGED PULL (all-gather fused into the GEMM over NVLS, sg_on != 0): B is row-sharded over the ranks of the symmetric heap;
  // at kernel entry every rank PUSHES its shard into the stage buffer of ALL ranks with multimem.st (the switch replicates),
  // sub-chunk by sub-chunk, and bumps that chunk's arrival counter on all ranks with one multimem.red.release; the TMA
  // producer loads B from the LOCAL stage buffer as soon as the 64-row chunks under its tile have arrived - tiles of early
  // chunks start while later chunks are still on the wire.  (Reference C2: ncclBcast / PullImpl + Broadcast before the first
  // consumer, kvstore_nccl.h:380-396, kvstore_dist.h:217-269.)
  SymmDev sg;               // peer tables, multicast base, flag pads (by value, like the collectives)
  int sg_on;
  int sg_channel;           // flag channel of the entry barrier (all ranks have left the previous user of the stage buffer)
  const void* sg_shard;     // this rank's shard [sg_rows_per_shard][sg_cols] bf16, row stride sg_ld
  size_t sg_stage_off;      // stage buffer (full stored matrix) and arrival counters, byte offsets in the heap
  size_t sg_flags_off;
  int sg_rows_per_shard, sg_cols;
  long long sg_ld;
  uint32_t sg_want;         // arrival count that marks a chunk complete for THIS launch (launch sequence x sub-chunks)
  // launch-time divisors (filled by fill_divisors): m tiles, tiles, im2col row length / rows per image / channel blocks
  FastDiv fd_num_m, fd_tiles, fd_Q, fd_P, fd_cblocks;
};

// ------------------------------------------------------------------ PTX -----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// im2col-mode load (NHWC activations, tensor map from cuTensorMapEncodeIm2col): `pixels` consecutive base positions
// starting at (w, h, n) - wrapping row by row, image by image inside the bounding box - each displaced by the filter-tap
// offset (off_w, off_h), `channels` channels from c: lands as [pixels][channels] rows, out-of-image elements zero-filled.
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h,
                                                   int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
// multicast load: the box lands at the same shared-memory offset of every CTA in `cta_mask`, and every destination
// CTA's mbarrier (same offset) receives the complete_tx for its copy
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
// MMA-completion arrive on the mbarrier at the same offset of every CTA in `cta_mask`
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all but the most recent bulk store have finished reading shared memory (double-buffered staging tiles)
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, 128B swizzle (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4  [16,30) LBO>>4  [32,46) SBO>>4  [46,48) version=1  [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

// ---------------------------------------------------------------- staged pull (all-gather fused into the GEMM) ----
// Phase 0, run by the epilogue warps (idle until the first accumulator is ready): push this rank's shard of B into the stage
// buffer of EVERY rank through the NVLS multicast mapping, (64-row chunk) x (256-column slice) work items spread over the
// CTAs, and bump the chunk's arrival counter on all ranks.  An optional entry barrier (sg_channel >= 0) orders the launch
// after every rank's previous user of the stage buffer; the default host path alternates two stage buffers instead: a rank
// can only reach launch i+2 after receiving every rank's chunks of launch i+1, which they push after finishing launch i.
template <int kPushThreads>
__device__ __forceinline__ void staged_push(const GemmParams& p, int pt) {
  const SymmDev& sg = p.sg;
  if (p.sg_channel >= 0) {
    asm volatile("bar.sync 1, %0;" ::"n"(kPushThreads) : "memory");
    if (pt < sg.world) {
      const uint32_t e = sg.epoch[p.sg_channel * kSigMaxBlocks + blockIdx.x];
      st_release_sys(sig_slot(sg, pt, p.sg_channel, blockIdx.x, sg.rank), e + 1);
      if (!spin_wait_ge(sig_slot(sg, sg.rank, p.sg_channel, blockIdx.x, pt), e + 1, sg.timeout_ns ? sg.timeout_ns : kSpinTimeoutNs))
        atomicExch(sg.err, 1);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kPushThreads) : "memory");
    if (pt == 0) sg.epoch[p.sg_channel * kSigMaxBlocks + blockIdx.x] += 1;
  }
  const int chunks = p.sg_rows_per_shard / 64;
  const int subs = (p.sg_cols + 255) / 256;
  const int cvec = p.sg_cols / 8;                 // 16-byte vectors per row
  char* mc_stage = sg.mc + p.sg_stage_off;
  uint32_t* mc_flags = reinterpret_cast<uint32_t*>(sg.mc + p.sg_flags_off);
  for (int w = blockIdx.x; w < chunks * subs; w += gridDim.x) {
    const int chunk = w / subs, sub = w - chunk * subs;
    const int v0 = sub * 32, v1 = min(v0 + 32, cvec);     // vectors [v0, v1) of each row
    const int nv = v1 - v0;
    uint4 val[8];
    int cnt = 0;
    for (int i = pt; i < 64 * nv; i += kPushThreads) {      // <= 8 (4-warp epilogue: 16) vectors per thread, loads first
      const int r = i / nv, cv = v0 + (i - r * nv);
      if (cnt < 8)
        val[cnt] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.sg_shard) +
                                                       (size_t)(chunk * 64 + r) * p.sg_ld) + cv);
      ++cnt;
    }
    cnt = 0;
    for (int i = pt; i < 64 * nv; i += kPushThreads) {
      const int r = i / nv, cv = v0 + (i - r * nv);
      const size_t grow = (size_t)sg.rank * p.sg_rows_per_shard + chunk * 64 + r;   // row of the full stored matrix
      const uint4 x = cnt < 8 ? val[cnt] : __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.sg_shard) +
                                                                               (size_t)(chunk * 64 + r) * p.sg_ld) + cv);
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_stage + (grow * p.sg_cols + (size_t)cv * 8) * 2),
                   "f"(__uint_as_float(x.x)), "f"(__uint_as_float(x.y)), "f"(__uint_as_float(x.z)), "f"(__uint_as_float(x.w))
                   : "memory");
      ++cnt;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kPushThreads) : "memory");     // every thread's stores precede the arrival
    if (pt == 0) {
      const size_t gchunk = (size_t)sg.rank * chunks + chunk;
      asm volatile("multimem.red.release.sys.global.add.u32 [%0], %1;" ::"l"(mc_flags + gchunk), "r"(1u) : "memory");
    }
  }
}

// Producer side: wait until the 64-row chunks covering stored rows [row0, row0 + nrows) of B have arrived in the local stage
// buffer; confirmed chunks are remembered in a caller-owned bitmap (512 chunks = 32768 stored rows) so that every chunk is
// polled once per CTA, and the proxy fence is only paid when something new was confirmed.
__device__ __forceinline__ void staged_wait_rows(const GemmParams& p, unsigned long long* seen, int row0, int nrows) {
  const uint32_t* flags = reinterpret_cast<const uint32_t*>(p.sg.base[p.sg.rank] + p.sg_flags_off);
  const int total = p.sg_rows_per_shard / 64 * p.sg.world;
  int c1 = (row0 + nrows + 63) / 64;
  if (c1 > total) c1 = total;
  bool fresh = false;
  for (int c = row0 / 64; c < c1; ++c) {
    const bool tracked = c < 512;
    if (tracked && ((seen[c >> 6] >> (c & 63)) & 1ull)) continue;
    if (!spin_wait_ge(flags + c, p.sg_want, p.sg.timeout_ns ? p.sg.timeout_ns : kSpinTimeoutNs)) atomicExch(p.sg.err, 1);
    if (tracked) seen[c >> 6] |= 1ull << (c & 63);
    fresh = true;
  }
  if (fresh) asm volatile("fence.proxy.async;" ::: "memory");      // the generic-proxy acquire orders the TMA (async proxy) reads
}

// ---------------------------------------------------------------- kernel ----
constexpr int kMaxStages = 12;
constexpr int kStagingBytes = 8 * 4096;             // one 32-row x 128-byte store tile per epilogue warp
constexpr int kBarrierBytes = 256;                  // full[12] empty[12] tfull[2] tempty[2] bfull[1] + tmem slot
constexpr int kSmemLimit = 227 * 1024;

template <int BLOCK_N, bool A_MN, bool B_MN>
struct GemmCfg {
  static constexpr int kABytes = kBlockM * kBlockK * 2;   // 16 KiB
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;   // 8 / 16 / 32 KiB
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);   // streaming-B default
  static constexpr int kTmemCols = (2 * BLOCK_N < 32) ? 32 : 2 * BLOCK_N;             // two accumulator stages
  static constexpr int kEpiWarps = (BLOCK_N == 64) ? 4 : 8;                           // 64 columns per warp at least
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + kBarrierBytes;
};

// MC = true (DLCFN_GEMM_MULTICAST=1, validated on B200 in round 2, not faster than the pair kernel): the grid is launched as
// clusters of two CTAs that work on M-adjacent tiles of the same N block.  They share the B tile: each CTA fetches HALF of it
// and multicasts that half into both CTAs' shared memory.  A stage of either CTA may only be refilled when BOTH CTAs' MMAs
// have consumed it (the peer's multicast writes into it), so the smem-empty barriers take two arrivals: each CTA's
// tcgen05.commit is itself multicast to both.  Everything else (TMEM, epilogue, TMA store) is per CTA and unchanged.
//
// CONV = true: implicit-GEMM convolution on NHWC tensors.  Only the producer changes - the activation operand is fetched by
// im2col-mode TMA, which performs the (pixel, tap) -> input-pixel arithmetic, the halo zero-fill and the row / image wrap in
// hardware:
//   fprop <A_MN=0,B_MN=0>  Y[NPQ, K]     = im2col(X)[NPQ, RS*C] * W[K, RS*C]^T      k-block -> (tap, 64 channels)
//   dgrad <A_MN=0,B_MN=1>  dX[NHW, C]    = im2col(dY)[NHW, RS*K] * W'[RS*K, C]      W' = taps reversed, read MN-major
//   wgrad <A_MN=1,B_MN=1>  dW[K, RS*C]   = dY[NPQ, K]^T * im2col(X)[NPQ, RS*C]      k-block = 64 pixels, split-K
// MMA issue, TMEM double buffering, epilogue and TMA store are those of the plain GEMM.
//
// p.b_resident (weights-stationary, chosen by the host when the whole [BLOCK_N x K] weight panel fits beside the A ring):
// the B operand of an N tile is loaded ONCE per CTA into its own shared-memory region and every M tile of that panel reuses
// it; the ring then carries A only (more stages).  The convolutions of a CNN are tall-skinny (M = N*P*Q pixels, K = R*S*C
// <= a few thousand): without this every 128-pixel tile re-fetches the same weights through the L2->SM crossbar, which is
// the bound of the small-channel layers (profiles/r2/conv_ncu_*.md).
//
// p.stats (fused BatchNorm statistics): the epilogue also accumulates, per output column (= channel), the sum and the sum
// of squares of the bf16-rounded outputs it stores, and writes them as partial rows [4 * grid][2][N] in the format of
// batchnorm.cu:kbn_fwd_finalize - the BatchNorm that follows the convolution skips its statistics pass over the tensor.
template <int BLOCK_N, bool A_MN, bool B_MN, bool MC = false, bool CONV = false>
__global__ void __launch_bounds__(kGemmThreads, 1) kgemm(const __grid_constant__ GemmParams p) {
  static_assert(!(MC && CONV), "the multicast-pair variant has no convolution producer");
  static_assert(!CONV || B_MN || !A_MN, "convolution modes: fprop <0,0>, dgrad <0,1>, wgrad <1,1>");
  using Cfg = GemmCfg<BLOCK_N, A_MN, B_MN>;
  constexpr int kEpiWarps = Cfg::kEpiWarps;
  const int stages = p.stages;
  const bool bres = !MC && p.b_resident != 0;
  const uint32_t stage_bytes = Cfg::kABytes + (bres ? 0 : Cfg::kBBytes);
  extern __shared__ uint8_t smem_raw[];
  // swizzle-128B atoms need 1024 B alignment
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bres_base = smem_base + stages * stage_bytes;        // 1024-aligned (all sizes are multiples of 8 KiB)
  const uint32_t stg_base = bres_base + (uint32_t)p.bres_bytes;
  const uint32_t bar_base = stg_base + (uint32_t)p.stg_bufs * kStagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + 2 + s); };
  const uint32_t bfull_bar = bar_base + 8u * (2 * kMaxStages + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + (bar_base - smem_base) + 8 * (2 * kMaxStages + 5));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + kBlockM - 1) / kBlockM;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  // MC: work items are PAIRS of M-adjacent tiles, one per cluster; a phantom second tile (odd num_m) loads zeros (TMA
  // out-of-bounds fill) and stores nothing (clipped), so both CTAs always run the same barrier sequence
  const int cta_rank = MC ? (int)cluster_ctarank() : 0;
  const int num_pm = MC ? (num_m + 1) / 2 : num_m;
  const int num_items = MC ? num_pm * num_n : num_tiles * p.splits;
  const int item0 = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int item_step = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    for (int s = 0; s < p.n_shards; ++s) tma_prefetch_desc(&p.tmB[s]);
    if (p.tma_store) tma_prefetch_desc(&p.tmD);
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), MC ? 2 : 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEpiWarps * 32);
    }
    mbar_init(bfull_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(Cfg::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if constexpr (MC) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast into them
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      unsigned long long sg_seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      auto sg_wait_rows = [&](int row0, int nrows) { staged_wait_rows(p, sg_seen, row0, nrows); };
      // one k-block of the B operand for output columns n_idx.. -> shared memory at sb (not used by the MC variant)
      auto load_b = [&](uint32_t sb, int kb, int n_idx, uint32_t bar) {
        const int k_idx = kb * kBlockK;
        if constexpr (CONV && B_MN && !A_MN) {
          // dgrad: the 64 contraction rows are output channels c0.. of list entry rs; the matching weights are the columns
          // of filter tap cv_tap_wi[rs] in W viewed as [Cout][R*S*Cin] (MN-major B: box {64 cin, 64 cout})
          int rs, cb;
          fdivmod(kb, p.fd_cblocks, rs, cb);
          const int c0 = cb * 64;
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_2d(sb + j * 8192, &p.tmB[0], bar, (int)p.cv_tap_wi[rs] * p.cv_cin + n_idx + 64 * j, c0);
        } else if constexpr (CONV && B_MN && A_MN) {
          // wgrad: the 64 contraction rows are output pixels k_idx..; output column panel P = (tap, 64-channel slice)
          int q0, t0, p0, n0;
          fdivmod(k_idx, p.fd_Q, t0, q0);
          fdivmod(t0, p.fd_P, n0, p0);
          const int last_panel = p.cv_ntaps * p.cv_cblocks - 1;
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j) {
            int panel = n_idx / 64 + j;
            if (panel > last_panel) panel = last_panel;    // N tail: columns >= N are never stored
            int rs, cb;
            fdivmod(panel, p.fd_cblocks, rs, cb);
            const int c0 = cb * 64;
            tma_load_im2col_4d(sb + j * 8192, &p.tmB[0], bar, c0, q0 * p.cv_stride + p.cv_lower_w,
                               p0 * p.cv_stride + p.cv_lower_h, n0, (uint16_t)p.cv_tap_ow[rs],
                               (uint16_t)p.cv_tap_oh[rs]);   // box {64 c, 64 pixels}
          }
        } else if constexpr (!B_MN) {
          // rows of B (= output columns n) may live on a peer: pick the owner's tensor map
          if (p.sg_on) sg_wait_rows(n_idx, BLOCK_N);           // staged pull: the chunks under this tile have arrived
          const int shard = (p.n_shards > 1) ? n_idx / p.rows_per_shard : 0;
          const int n_loc = n_idx - shard * p.rows_per_shard;
          tma_load_2d(sb, &p.tmB[shard], bar, k_idx, n_loc);  // box {64 k, BLOCK_N n}
        } else {
          // MN-major B is stored [K, N]: the CONTRACTION rows may live on a peer
          if (p.sg_on) sg_wait_rows(k_idx, kBlockK);
          const int shard = (p.n_shards > 1) ? k_idx / p.rows_per_shard : 0;
          const int k_loc = k_idx - shard * p.rows_per_shard;
#pragma unroll
          for (int j = 0; j < BLOCK_N / 64; ++j)
            tma_load_2d(sb + j * 8192, &p.tmB[shard], bar, n_idx + 64 * j, k_loc);
        }
      };
      int stage = 0;
      uint32_t phase = 0;
      int cur_n = -1;                       // b_resident: the N tile whose weights are in shared memory
      int acc = 0, prev_acc = 0;            // mirror of the MMA issuer's accumulator-stage sequence
      uint32_t acc_phase = 0, prev_acc_phase = 0;
      for (int item = item0; item < num_items; item += item_step) {
        int tile = 0, split = 0, mt, nt;
        if constexpr (MC) { nt = item / num_pm; mt = (item - nt * num_pm) * 2 + cta_rank; }
        else { fdivmod(item, p.fd_tiles, split, tile); fdivmod(tile, p.fd_num_m, nt, mt); }
        const int m_idx = mt * kBlockM, n_idx = nt * BLOCK_N;
        const int kb0 = split * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        // im2col position of the tile's first pixel (fprop / dgrad): once per tile, not per k-block
        int cv_q0 = 0, cv_p0 = 0, cv_n0 = 0, cv_rs = 0, cv_cb = 0;
        if constexpr (CONV && !A_MN) {
          int t0;
          fdivmod(m_idx, p.fd_Q, t0, cv_q0);
          fdivmod(t0, p.fd_P, cv_n0, cv_p0);
          fdivmod(kb0, p.fd_cblocks, cv_rs, cv_cb);
        }
        if (bres) {
          if (n_idx != cur_n) {
            // every MMA that reads the old panel has completed once the previous item's accumulator is committed
            if (cur_n >= 0) mbar_wait(tfull_bar(prev_acc), prev_acc_phase);
            mbar_expect_tx(bfull_bar, (uint32_t)num_kb * Cfg::kBBytes);
            for (int kb = 0; kb < num_kb; ++kb) load_b(bres_base + (uint32_t)kb * Cfg::kBBytes, kb, n_idx, bfull_bar);
            cur_n = n_idx;
          }
          prev_acc = acc; prev_acc_phase = acc_phase;
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          const int k_idx = kb * kBlockK;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = sa + Cfg::kABytes;
          mbar_expect_tx(full_bar(stage), stage_bytes);
          if constexpr (CONV && !A_MN) {
            // fprop / dgrad: k-block = (filter tap rs, 64-channel slice); 128 consecutive output pixels from m_idx
            tma_load_im2col_4d(sa, &p.tmA, full_bar(stage), cv_cb * 64, cv_q0 * p.cv_stride + p.cv_lower_w,
                               cv_p0 * p.cv_stride + p.cv_lower_h, cv_n0, (uint16_t)p.cv_tap_ow[cv_rs],
                               (uint16_t)p.cv_tap_oh[cv_rs]);   // box {64 c, 128 pixels}
            if (++cv_cb == p.cv_cblocks) { cv_cb = 0; ++cv_rs; }
          } else if constexpr (!A_MN) {
            tma_load_2d(sa, &p.tmA, full_bar(stage), k_idx, m_idx);  // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j)  // box {64 m, 64 k} per 128B-wide MN atom
              tma_load_2d(sa + j * 8192, &p.tmA, full_bar(stage), m_idx + 64 * j, k_idx);
          }
          if constexpr (MC) {
            // this CTA's half of the shared B tile, multicast to both CTAs of the pair
            if constexpr (!B_MN) {       // K-major: box {64 k, BLOCK_N/2 n}; the halves are 16 KiB apart (swizzle-atom aligned)
              tma_load_2d_mc(sb + cta_rank * (Cfg::kBBytes / 2), &p.tmB[0], full_bar(stage), k_idx,
                             n_idx + cta_rank * (BLOCK_N / 2), (uint16_t)3);
            } else {                     // MN-major: BLOCK_N/64 boxes {64 n, 64 k}; this CTA issues its half of them
#pragma unroll
              for (int j = 0; j < BLOCK_N / 128; ++j) {
                const int jj = cta_rank * (BLOCK_N / 128) + j;
                tma_load_2d_mc(sb + jj * 8192, &p.tmB[0], full_bar(stage), n_idx + 64 * jj, k_idx, (uint16_t)3);
              }
            }
          } else if (!bres) {
            load_b(sb, kb, n_idx, full_bar(stage));
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
      if constexpr (MC) {
        // producer tail: every slot's last use has been released by BOTH CTAs, i.e. the peer's multicast commits that
        // target this CTA's barriers have landed - only then may this CTA run towards its exit
        for (int s = 0; s < stages; ++s) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32, a=b=bf16, majors, N>>3, M>>4
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                             ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BLOCK_N >> 3) << 17) |
                             ((uint32_t)(kBlockM >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int cur_n = -1;
      uint32_t bphase = 0;
      for (int item = item0; item < num_items; item += item_step) {
        int tile = 0, split = 0;
        if constexpr (!MC) fdivmod(item, p.fd_tiles, split, tile);
        const int kb0 = split * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        if (bres) {
          const int n_idx = fdiv(tile, p.fd_num_m) * BLOCK_N;
          if (n_idx != cur_n) {            // the producer has (re)loaded the weight panel
            mbar_wait(bfull_bar, bphase);
            bphase ^= 1u;
            cur_n = n_idx;
          }
        }
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * stage_bytes;
          const uint32_t sb = bres ? bres_base + (uint32_t)kb * Cfg::kBBytes : sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            // K-major: advance 16 elements = 32 B inside the swizzle atom; SBO = 8 rows * 128 B
            // MN-major: advance 16 k-rows = 2048 B; LBO = stride between 64-wide MN atoms, SBO = 1024
            const uint64_t da = A_MN ? umma_desc(sa + k * 2048, 8192, 1024) : umma_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc(sb + k * 2048, 8192, 1024) : umma_desc(sb + k * 32, 16, 1024);
            tc_mma_bf16(tmem_d, da, db, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          // smem slot reusable once these MMAs have read it (MC: tell the peer too, its multicast writes into this slot)
          if constexpr (MC) tc_commit_mc(empty_bar(stage), (uint16_t)3);
          else tc_commit(empty_bar(stage));
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));  // accumulator complete -> epilogue (and, b_resident, -> the producer's panel swap)
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp < 2 + kEpiWarps) {
    // ===================== epilogue (warps 2..2+kEpiWarps) =====================
    if (!MC && !CONV && p.sg_on) staged_push<kEpiWarps * 32>(p, (int)threadIdx.x - 64);
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;   // which part of the tile's columns this warp owns (8 warps: halves; 4 warps: all)
    constexpr int kChunks = BLOCK_N / (32 * (kEpiWarps / 4));   // 32-column chunks per warp
    constexpr int kBoxes = kChunks / 2;                         // 64-column TMA store boxes per warp
    const bool splitk = p.splits > 1;
    const bool do_stats = p.stats != nullptr && p.tma_store && !splitk;
    float st_s[kBoxes][2], st_q[kBoxes][2];   // fused BN statistics: this lane's two columns of every box
#pragma unroll
    for (int b = 0; b < kBoxes; ++b) { st_s[b][0] = st_s[b][1] = 0.f; st_q[b][0] = st_q[b][1] = 0.f; }
    int st_n = -1;
    auto flush_stats = [&]() {
      if (st_n < 0) return;
      float* row = p.stats + (size_t)(blockIdx.x * 4 + q) * 2 * p.N;
#pragma unroll
      for (int b = 0; b < kBoxes; ++b) {
        const int col = st_n + (half * kChunks + 2 * b) * 32 + 2 * lane;
        if (col + 1 < p.N) {
          *reinterpret_cast<float2*>(row + col) = make_float2(st_s[b][0], st_s[b][1]);
          *reinterpret_cast<float2*>(row + p.N + col) = make_float2(st_q[b][0], st_q[b][1]);
        } else if (col < p.N) {
          row[col] = st_s[b][0];
          row[p.N + col] = st_q[b][0];
        }
        st_s[b][0] = st_s[b][1] = 0.f; st_q[b][0] = st_q[b][1] = 0.f;
      }
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool stg2 = p.stg_bufs == 2;
    uint32_t stg_sel = 0;               // which of this warp's staging tiles the next box goes through
    for (int item = item0; item < num_items; item += item_step) {
      int tile = 0, split = 0, mt, nt;
      if constexpr (MC) { nt = item / num_pm; mt = (item - nt * num_pm) * 2 + cta_rank; }
      else { fdivmod(item, p.fd_tiles, split, tile); fdivmod(tile, p.fd_num_m, nt, mt); }
      const int m_idx = mt * kBlockM, n_idx = nt * BLOCK_N;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m_idx + q * 32 + lane;
      const bool row_ok = row < p.M;
      long long row_pix = row;         // output pixel (row of the [pixels, N] output matrix) this GEMM row is stored to
      if (CONV && p.out_remap) {
        int jj, tt, ii, nn;
        fdivmod(row, p.fd_Q, tt, jj);
        fdivmod(tt, p.fd_P, nn, ii);
        row_pix = ((long long)nn * p.out_H + (long long)ii * p.out_step + p.out_h0) * p.out_W + (long long)jj * p.out_step + p.out_w0;
      }
      // all of this warp's accumulator columns in flight at once, one wait
      uint32_t r[kChunks][32];
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BLOCK_N + (half * kChunks + c) * 32, r[c]);
      tc_wait_ld();
      // the accumulator stage is free as soon as it sits in registers: let the next tile's MMAs start
      tc_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      const bool d_fp32 = splitk || p.d_is_fp32;
      const bool accum = !splitk && p.accumulate;
      const bool relu = !splitk && p.relu;
      const float* bias = splitk ? nullptr : p.bias;
      const long long ldd = splitk ? (long long)p.N : p.ldd;
      void* dbase = splitk ? (void*)(p.ws + (size_t)split * p.M * p.N) : p.D;
      if (p.tma_store && !splitk) {
        if (do_stats && n_idx != st_n) {
          flush_stats();
          st_n = n_idx;
        }
        // bf16 output through shared memory: lane = row writes its 64 columns (128 B) as 8 16-byte chunks at the
        // 128B-swizzled position (chunk ^ (row & 7)) - conflict-free - and one lane hands the tile to TMA.
#pragma unroll
        for (int bx = 0; bx < kBoxes; ++bx) {
          const int colb = n_idx + (half * kChunks + 2 * bx) * 32;   // first column of this 64-wide box
          if (colb < p.N) {
            const uint32_t stg = stg_base + ((uint32_t)(warp - 2) * (stg2 ? 2u : 1u) + (stg2 ? stg_sel : 0u)) * 4096u;
            uint8_t* stg_gen = smem_gen + (stg - smem_base);
            stg_sel ^= 1u;
            if (lane == 0) {                                         // the previous store out of THIS buffer has been read
              if (stg2) tma_store_wait_read1();
              else tma_store_wait_read();
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float o[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) o[t] = __uint_as_float(r[2 * bx + (j >> 2)][(j & 3) * 8 + t]);
              if (bias != nullptr) {
#pragma unroll
                for (int t = 0; t < 8; ++t)
                  if (colb + j * 8 + t < p.N) o[t] += __ldg(bias + colb + j * 8 + t);
              }
              if (relu) {
#pragma unroll
                for (int t = 0; t < 8; ++t) o[t] = fmaxf(o[t], 0.f);
              }
              *reinterpret_cast<uint4*>(stg_gen + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&p.tmD, stg, colb, m_idx + q * 32);       // rows >= M / cols >= N are clipped by the map
              tma_store_commit();
            }
            if (do_stats) {
              // column sums over this warp's 32 rows, read back from the staging tile (the values BatchNorm will see):
              // lane l owns columns 2l, 2l+1 = one 4-byte word per row, a conflict-free 128-byte row read per step.
              // Rows >= M hold exact zeros (their A rows are TMA zero-fill) and add nothing.
              const int chunk = lane >> 2, within = (lane & 3) * 4;
#pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(stg_gen + rr * 128 + ((chunk ^ (rr & 7)) << 4) + within);
                const float a = bf16_lo(w), b = bf16_hi(w);
                st_s[bx][0] += a; st_q[bx][0] = fmaf(a, a, st_q[bx][0]);
                st_s[bx][1] += b; st_q[bx][1] = fmaf(b, b, st_q[bx][1]);
              }
            }
          }
        }
        continue;
      }
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int col0 = n_idx + (half * kChunks + c) * 32;
        if (row_ok && col0 < p.N) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[c][j]);
          if (bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) f[j] += __ldg(bias + col0 + j);
          }
          if (d_fp32) {
            float* d = reinterpret_cast<float*>(dbase) + (size_t)(splitk ? (long long)row : row_pix) * ldd + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j + 3 < p.N) {
                float4 o = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                if (accum) {
                  float4 old = *reinterpret_cast<const float4*>(d + j);
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                *reinterpret_cast<float4*>(d + j) = o;
              } else {
                for (int t = 0; t < 4; ++t)
                  if (col0 + j + t < p.N) {
                    float o = f[j + t] + (accum ? d[j + t] : 0.f);
                    d[j + t] = relu ? fmaxf(o, 0.f) : o;
                  }
              }
            }
          } else {
            __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(dbase) + (size_t)row_pix * ldd + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j + 7 < p.N) {
                float o[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) o[t] = f[j + t];
                if (accum) {
                  uint4 old = *reinterpret_cast<const uint4*>(d + j);
                  o[0] += bf16_lo(old.x); o[1] += bf16_hi(old.x); o[2] += bf16_lo(old.y); o[3] += bf16_hi(old.y);
                  o[4] += bf16_lo(old.z); o[5] += bf16_hi(old.z); o[6] += bf16_lo(old.w); o[7] += bf16_hi(old.w);
                }
                if (relu) {
#pragma unroll
                  for (int t = 0; t < 8; ++t) o[t] = fmaxf(o[t], 0.f);
                }
                *reinterpret_cast<uint4*>(d + j) = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                                                              pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
              } else {
                for (int t = 0; t < 8; ++t)
                  if (col0 + j + t < p.N) {
                    float o = f[j + t] + (accum ? __bfloat162float(d[j + t]) : 0.f);
                    d[j + t] = __float2bfloat16(relu ? fmaxf(o, 0.f) : o);
                  }
              }
            }
          }
        }
      }
    }
    if (do_stats) flush_stats();
    if (p.tma_store && lane == 0) tma_store_wait_read();   // staging smem must outlive the last bulk store's reads
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC) cluster_sync_all();   // no CTA leaves while its peer may still multicast into it or signal its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols));
  }
}


// ====================================================================================================================
// CTA-PAIR kernel (EXPERIMENTAL, written at the end of round 1 without GPU access; DLCFN_GEMM_2CTA=1).
//
// tcgen05.mma.cta_group::2: the two CTAs of a cluster (the two SMs of a TPC) execute ONE 256 x 256 x 16 MMA per
// instruction.  CTA r holds rows [128r, 128r+128) of the A tile and columns [128r, 128r+128) of the B tile in ITS shared
// memory and rows [128r, 128r+128) of the accumulator in ITS tensor memory; the tensor cores of each SM read the B half
// they do not own from the partner's shared memory.  Per k-block a CTA therefore pulls 16 KiB (A) + 16 KiB (B half) from
// L2 instead of 16 + 32 KiB, and the shared-memory reads per MMA halve as well - the single-CTA kernel is limited by
// exactly those two (26 TB/s of L2 reads at full rate, 81 % tensor-pipe utilisation in profiles/gemm_tcgen05_8192_ncu.txt).
//
// Protocol (after CUTLASS's sm100 2-SM pipeline, re-derived here):
//   * both CTAs run a TMA producer; their loads use the .cta_group::2 form whose completion bytes are credited to the
//     LEADER's (rank 0) full barrier (barrier address with the CTA bit cleared), which expects the bytes of both CTAs
//   * only the leader runs the MMA issuer; its tcgen05.commit is multicast to both CTAs: to the smem-empty barriers (each
//     producer refills its own ring) and to the tmem-full barriers (each CTA's epilogue drains its own 128 rows)
//   * epilogue warps of both CTAs arrive on the LEADER's tmem-empty barrier (mapa + remote arrive), 16 arrivals per stage
//   * TMEM is allocated / freed with cta_group::2 by warp 1 of both CTAs; cluster barriers bracket the kernel
constexpr int k2BlockN = 256;
constexpr int k2HalfN = 128;
constexpr int k2Stages = 5;            // 5 x 32 KiB ring + two staging tiles per epilogue warp (64 KiB)
constexpr int k2ABytes = kBlockM * kBlockK * 2;     // 16 KiB
constexpr int k2BBytes = k2HalfN * kBlockK * 2;     // 16 KiB (this CTA's half of the B tile)
constexpr int k2StageBytes = k2ABytes + k2BBytes;
constexpr int k2StagingBytes = 2 * 8 * 4096;
constexpr int k2SmemBytes = k2Stages * k2StageBytes + k2StagingBytes + 1024 + 256;
constexpr int k2TmemCols = 2 * k2BlockN;            // two accumulator stages = all 512 columns

__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}
// arrive on the barrier at the same shared-memory offset of cluster CTA `rank`.  RELAXED: the arrival only says "this warp's
// tcgen05.ld of the accumulator stage has completed" (tcgen05.wait::ld + fence::before_thread_sync precede it) - no memory
// is published.  The default .release.cluster form compiles to MEMBAR.ALL.CTA + ERRBAR in front of the arrive, which held 37 %
// of the pair kernel's stall samples on short-K layers (profiles/r2/pair_ncu_1x1_64_256.md).
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(bar), "r"(rank)
      : "memory");
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kGemmThreads, 1) kgemm_pair(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stg_base = smem_base + k2Stages * k2StageBytes;
  const uint32_t bar_base = stg_base + k2StagingBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (k2Stages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * k2Stages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * k2Stages + 2 + s); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + k2Stages * k2StageBytes + k2StagingBytes +
                                                    8 * (2 * k2Stages + 4));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = (int)cluster_ctarank();
  const bool leader = cta_rank == 0;

  const int num_m = (p.M + kBlockM - 1) / kBlockM;
  const int num_pm = (num_m + 1) / 2;                 // pairs of M tiles; a phantom second tile loads zeros, stores nothing
  const int num_n = (p.N + k2BlockN - 1) / k2BlockN;
  const int num_items = num_pm * num_n;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int item0 = (int)(blockIdx.x >> 1), item_step = (int)(gridDim.x >> 1);
  // grouped raster: items walk the N tiles of a group of raster_g M-pair rows before moving to the next group
  const int rg = p.raster_g > 0 ? p.raster_g : num_pm;
  auto pair_raster = [&](int item, int& pm, int& nt) {
    const int per_group = rg * num_n;
    const int g = item / per_group, within = item - g * per_group;
    const int rows = min(rg, num_pm - g * rg);
    nt = within / rows;
    pm = g * rg + (within - nt * rows);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB[0]);
    tma_prefetch_desc(&p.tmD);
    for (int s = 0; s < k2Stages; ++s) {
      mbar_init(full_bar(s), 1);      // used on the leader only: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(empty_bar(s), 1);     // one multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);     // one multicast tcgen05.commit
      mbar_init(tempty_bar(s), 16);   // used on the leader only: 8 epilogue warps x 2 CTAs
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();                 // both CTAs' barriers exist before any remote arrive / multicast commit / pair load
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(k2TmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      unsigned long long sg_seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int item = item0; item < num_items; item += item_step) {
        int pm, nt;
        pair_raster(item, pm, nt);
        const int m_idx = (pm * 2 + cta_rank) * kBlockM;
        const int n_idx = nt * k2BlockN + cta_rank * k2HalfN;     // this CTA's half of the B tile
        for (int kb = 0; kb < num_kb; ++kb) {
          const int k_idx = kb * kBlockK;
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa = smem_base + stage * k2StageBytes;
          const uint32_t sb = sa + k2ABytes;
          const uint32_t lbar = full_bar(stage) & 0xFEFFFFFFu;                 // the leader CTA's barrier
          if (leader) mbar_expect_tx(full_bar(stage), 2 * k2StageBytes);
          if constexpr (!A_MN) {
            tma_load_2d_pair(sa, &p.tmA, lbar, k_idx, m_idx);                  // box {64 k, 128 m}
          } else {
#pragma unroll
            for (int j = 0; j < kBlockM / 64; ++j) tma_load_2d_pair(sa + j * 8192, &p.tmA, lbar, m_idx + 64 * j, k_idx);
          }
          if constexpr (!B_MN) {
            if (p.sg_on) staged_wait_rows(p, sg_seen, n_idx, k2HalfN);         // staged pull: this half tile's rows have arrived
            tma_load_2d_pair(sb, &p.tmB[0], lbar, k_idx, n_idx);               // box {64 k, 128 n}
          } else {
            if (p.sg_on) staged_wait_rows(p, sg_seen, k_idx, kBlockK);
#pragma unroll
            for (int j = 0; j < k2HalfN / 64; ++j) tma_load_2d_pair(sb + j * 8192, &p.tmB[0], lbar, n_idx + 64 * j, k_idx);
          }
          if (++stage == k2Stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(k2BlockN >> 3) << 17) | ((uint32_t)((2 * kBlockM) >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = item0; item < num_items; item += item_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * k2BlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_base + stage * k2StageBytes;
          const uint32_t sb = sa + k2ABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t da = A_MN ? umma_desc(sa + k * 2048, 8192, 1024) : umma_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? umma_desc(sb + k * 2048, 8192, 1024) : umma_desc(sb + k * 32, 16, 1024);
            tc_mma_bf16_pair(tmem_d, da, db, idesc, (kb != 0 || k != 0) ? 1u : 0u);
          }
          tc_commit_pair(empty_bar(stage));      // both producers may refill this slot
          if (++stage == k2Stages) { stage = 0; phase ^= 1u; }
        }
        tc_commit_pair(tfull_bar(acc));          // both epilogues may drain this accumulator stage
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9 of both CTAs): bf16 through swizzled staging + TMA store ==========
    if (p.sg_on) staged_push<256>(p, (int)threadIdx.x - 64);     // staged pull, phase 0: push this rank's shard (see staged_push)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int kChunks = k2BlockN / 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t stg_sel = 0;           // two staging tiles per warp: fill one while TMA still reads the other
    for (int item = item0; item < num_items; item += item_step) {
      int pm, nt;
      pair_raster(item, pm, nt);
      const int m_idx = (pm * 2 + cta_rank) * kBlockM;
      const int n_idx = nt * k2BlockN;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      uint32_t r[kChunks][32];
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        tc_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * k2BlockN + (half * kChunks + c) * 32, r[c]);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(tempty_bar(acc), 0);     // this warp's quarter of the stage is in registers
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
#pragma unroll
      for (int bx = 0; bx < kChunks / 2; ++bx) {
        const int colb = n_idx + (half * kChunks + 2 * bx) * 32;
        if (colb < p.N) {
          const uint32_t stg = stg_base + ((uint32_t)(warp - 2) * 2u + stg_sel) * 4096u;
          uint8_t* stg_gen = smem_gen + (stg - smem_base);
          stg_sel ^= 1u;
          if (lane == 0) tma_store_wait_read1();
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float o[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) o[t] = __uint_as_float(r[2 * bx + (j >> 2)][(j & 3) * 8 + t]);
            if (p.bias != nullptr) {
#pragma unroll
              for (int t = 0; t < 8; ++t)
                if (colb + j * 8 + t < p.N) o[t] += __ldg(p.bias + colb + j * 8 + t);
            }
            if (p.relu) {
#pragma unroll
              for (int t = 0; t < 8; ++t) o[t] = fmaxf(o[t], 0.f);
            }
            *reinterpret_cast<uint4*>(stg_gen + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&p.tmD, stg, colb, m_idx + q * 32);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_read();
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();     // the partner has consumed everything that lives in this CTA's shared / tensor memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(k2TmemCols));
  }
}

// Sum the split-K partials, then bias / accumulate / relu / convert.  A CTA is 32 float4 columns x 8 slice lanes:
// lane group g sums slices g, g+8, g+16, ... (4 loads in flight), the 8 partial sums meet in shared memory and are
// added in a fixed order, so the result does not depend on scheduling.  (The first version walked all slices
// serially per thread: 23 us for a 64x256 output with 148 slices - latency-bound, see profiles/r1_ncu.)
__global__ void __launch_bounds__(256) ksplitk_reduce(const float* __restrict__ ws, int splits, int M, int N, void* D,
                                                       long long ldd, int d_is_fp32, const float* __restrict__ bias,
                                                       int accumulate, int relu) {
  __shared__ float4 part[8][32];
  const long long MN = (long long)M * N;
  const long long nvec = MN / 4;   // N % 4 == 0 is guaranteed by the host
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  for (long long v0 = (long long)blockIdx.x * 32; v0 < nvec; v0 += (long long)gridDim.x * 32) {
    const long long v = v0 + lane;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < nvec) {
      int sidx = grp;
      for (; sidx + 24 < splits; sidx += 32) {
        const float4 b0 = reinterpret_cast<const float4*>(ws + (size_t)sidx * MN)[v];
        const float4 b1 = reinterpret_cast<const float4*>(ws + (size_t)(sidx + 8) * MN)[v];
        const float4 b2 = reinterpret_cast<const float4*>(ws + (size_t)(sidx + 16) * MN)[v];
        const float4 b3 = reinterpret_cast<const float4*>(ws + (size_t)(sidx + 24) * MN)[v];
        a.x += b0.x; a.y += b0.y; a.z += b0.z; a.w += b0.w;
        a.x += b1.x; a.y += b1.y; a.z += b1.z; a.w += b1.w;
        a.x += b2.x; a.y += b2.y; a.z += b2.z; a.w += b2.w;
        a.x += b3.x; a.y += b3.y; a.z += b3.z; a.w += b3.w;
      }
      for (; sidx < splits; sidx += 8) {
        const float4 b0 = reinterpret_cast<const float4*>(ws + (size_t)sidx * MN)[v];
        a.x += b0.x; a.y += b0.y; a.z += b0.z; a.w += b0.w;
      }
    }
    part[grp][lane] = a;
    __syncthreads();
    if (grp == 0 && v < nvec) {
#pragma unroll
      for (int g = 1; g < 8; ++g) {
        const float4 b0 = part[g][lane];
        a.x += b0.x; a.y += b0.y; a.z += b0.z; a.w += b0.w;
      }
      const long long e = v * 4;
      const int row = (int)(e / N), col = (int)(e - (long long)row * N);
      if (bias != nullptr) { a.x += bias[col]; a.y += bias[col + 1]; a.z += bias[col + 2]; a.w += bias[col + 3]; }
      if (d_is_fp32) {
        float* d = reinterpret_cast<float*>(D) + (size_t)row * ldd + col;
        if (accumulate) { a.x += d[0]; a.y += d[1]; a.z += d[2]; a.w += d[3]; }
        if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
        *reinterpret_cast<float4*>(d) = a;
      } else {
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(D) + (size_t)row * ldd + col;
        if (accumulate) {
          const uint2 old = *reinterpret_cast<const uint2*>(d);
          a.x += bf16_lo(old.x); a.y += bf16_hi(old.x); a.z += bf16_lo(old.y); a.w += bf16_hi(old.y);
        }
        if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
        *reinterpret_cast<uint2*>(d) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ host ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)f;
  }
  return fn;
}

// 2-D bf16 tensor [rows, cols] (cols contiguous, row stride ld elements), box {box_cols, box_rows}.
// A descriptor depends only on (pointer, shape, stride, box), and a training step presents the same few hundred
// combinations every iteration (the caching allocator hands back the same blocks), so encoded maps are kept in a
// small per-thread direct-mapped cache instead of calling into the driver three times per GEMM.
struct MapKey {
  const void* ptr;
  uint64_t rows, cols, ld;
  uint32_t box_cols, box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_cols == o.box_cols && box_rows == o.box_rows;
  }
};
struct MapSlot {
  MapKey key;
  CUtensorMap map;
  bool valid;
};
constexpr int kMapCacheSlots = 1024;

static int make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_cols,
                    uint32_t box_rows) {
  static thread_local MapSlot* cache = nullptr;
  if (!cache) {
    cache = static_cast<MapSlot*>(aligned_alloc(alignof(MapSlot) < 64 ? 64 : alignof(MapSlot), kMapCacheSlots * sizeof(MapSlot)));
    if (cache) memset(cache, 0, kMapCacheSlots * sizeof(MapSlot));
  }
  const MapKey key{ptr, rows, cols, ld, box_cols, box_rows};
  uint64_t h = (uint64_t)(uintptr_t)ptr * 0x9E3779B97F4A7C15ull;
  h ^= (rows * 0xC2B2AE3D27D4EB4Full) ^ (cols * 0x165667B19E3779F9ull) ^ (ld << 17) ^ ((uint64_t)box_cols << 40) ^ box_rows;
  MapSlot* slot = cache ? &cache[(h >> 20) % kMapCacheSlots] : nullptr;
  if (slot && slot->valid && slot->key == key) {
    memcpy(m, &slot->map, sizeof(CUtensorMap));
    return 0;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return (int)cudaErrorNotSupported;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[dlcfn] cuTensorMapEncodeTiled failed (%d): ptr=%p rows=%llu cols=%llu ld=%llu box=%ux%u\n", (int)r, ptr,
            (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
    return 10000 + (int)r;
  }
  if (slot) {
    slot->key = key;
    memcpy(&slot->map, m, sizeof(CUtensorMap));
    slot->valid = true;
  }
  return 0;
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool CONV = false>
static int launch(GemmParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N, A_MN, B_MN>;
  auto kern = kgemm<BLOCK_N, A_MN, B_MN, false, CONV>;
  static bool configured = false;
  if (!configured) {
    cudaError_t ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit);
    if (ae != cudaSuccess) {
      fprintf(stderr, "[dlcfn] cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(ae));
      return (int)ae;
    }
    configured = true;
  }
  const int num_m = (p.M + kBlockM - 1) / kBlockM;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles = num_m * num_n;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  p.fd_num_m = make_fastdiv(num_m);
  p.fd_tiles = make_fastdiv(tiles);
  p.fd_Q = make_fastdiv(p.cv_Q);
  p.fd_P = make_fastdiv(p.cv_P);
  p.fd_cblocks = make_fastdiv(p.cv_cblocks);
  // split-K when the tiles alone cannot fill the machine and the reduction is long enough to cut
  p.splits = 1;
  p.kb_per_split = num_kb;
  p.ws = nullptr;
  if (p.split_k_ok && tiles * 2 <= num_sms && num_kb >= 8 && (p.N % 4) == 0) {
    int splits = num_sms / tiles;
    if (splits > num_kb / 4) splits = num_kb / 4;
    if (splits >= 2) {
      p.kb_per_split = (num_kb + splits - 1) / splits;
      p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;   // every slice owns >= 1 k-block
    }
  }
  // pipeline shape: weights-stationary when B is a weight panel (A is the activation operand: fprop / dgrad forms), the whole
  // [BLOCK_N x K] panel fits beside at least 4 A stages, and every CTA works on more than one M tile of it
  static const bool bres_env = [] { const char* e = getenv("DLCFN_GEMM_BRES"); return !(e && e[0] == '0'); }();
  // short reductions are epilogue-paced: two staging tiles per warp let a warp convert the next box while TMA drains the last
  p.stg_bufs = (p.tma_store && num_kb <= 12) ? 2 : 1;
  const int fixed = p.stg_bufs * kStagingBytes + 1024 + kBarrierBytes;
  p.b_resident = 0;
  p.bres_bytes = 0;
  const long long panel = (long long)num_kb * Cfg::kBBytes;
  if (bres_env && !A_MN && p.n_shards == 1 && p.splits == 1 && tiles >= 2 * num_sms &&
      panel + 4 * Cfg::kABytes + fixed <= kSmemLimit) {
    p.b_resident = 1;
    p.bres_bytes = (int)panel;
  }
  const int stage_bytes = Cfg::kABytes + (p.b_resident ? 0 : Cfg::kBBytes);
  int stages = (kSmemLimit - fixed - p.bres_bytes) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (!p.b_resident && stages > Cfg::kStages + 2) stages = Cfg::kStages + 2;
  p.stages = stages;
  const int smem_bytes = stages * stage_bytes + p.bres_bytes + fixed;
  if (p.splits > 1) {
    // stream-ordered workspace from the device's default pool; the pool keeps its memory across synchronisation points
    // (the default release threshold 0 hands everything back to the driver at every sync: 1+ ms per re-allocation)
    static thread_local int pool_dev = -1;
    int dev = 0;
    DLCFN_CHECK(cudaGetDevice(&dev));
    if (pool_dev != dev) {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t keep = ~0ull;
        (void)cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      }
      pool_dev = dev;
    }
    void* ws = nullptr;
    DLCFN_CHECK(cudaMallocAsync(&ws, (size_t)p.splits * p.M * p.N * sizeof(float), stream));
    p.ws = (float*)ws;
  }
  int grid = tiles * p.splits;
  if (grid > num_sms) grid = num_sms;
  if (p.stats != nullptr) {
    // rows of CTAs / N tiles that never meet stay zero
    if (p.splits > 1 || !p.tma_store) return (int)cudaErrorInvalidValue;
    DLCFN_CHECK(cudaMemsetAsync(p.stats, 0, (size_t)4 * num_sms * 2 * p.N * sizeof(float), stream));
  }
  kern<<<grid, kGemmThreads, smem_bytes, stream>>>(p);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess)
    fprintf(stderr, "[dlcfn] kgemm launch failed: %s (grid %d, smem %d)\n", cudaGetErrorString(le), grid, smem_bytes);
  if (p.splits > 1) {
    if (le == cudaSuccess) {
      const long long nvec = (long long)p.M * p.N / 4;
      long long rb = (nvec + 31) / 32;
      if (rb > 148 * 8) rb = 148 * 8;
      ksplitk_reduce<<<(int)rb, 256, 0, stream>>>(p.ws, p.splits, p.M, p.N, p.D, p.ldd, p.d_is_fp32, p.bias, p.accumulate,
                                                  p.relu);
      le = cudaGetLastError();
    }
    cudaFreeAsync(p.ws, stream);
  }
  return (int)le;
}


// ------------------------------------------------------------------------------------------ implicit-GEMM convolution
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeIm2colFn get_encode_im2col() {
  static EncodeIm2colFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeIm2colFn)f;
  }
  return fn;
}

struct Im2colKey {
  const void* ptr;
  int n, h, w, c, lower_w, lower_h, upper_w, upper_h, stride, pixels;
  long long xs_w, xs_h, xs_n;
  bool operator==(const Im2colKey& o) const {
    return ptr == o.ptr && n == o.n && h == o.h && w == o.w && c == o.c && lower_w == o.lower_w && lower_h == o.lower_h &&
           upper_w == o.upper_w && upper_h == o.upper_h && stride == o.stride && pixels == o.pixels && xs_w == o.xs_w &&
           xs_h == o.xs_h && xs_n == o.xs_n;
  }
};
struct Im2colSlot {
  Im2colKey key;
  CUtensorMap map;
  bool valid;
};
constexpr int kIm2colCacheSlots = 256;

// NHWC bf16 tensor [n, h, w, c] as an im2col-mode map: 64 channels x `pixels` base positions per load, 128B swizzle.
// Base positions live in the box [lower, extent - 1 + upper] per spatial dimension and advance by `stride`.
static int make_map_im2col(CUtensorMap* m, const void* ptr, int n, int h, int w, int c, int lower_w, int lower_h, int upper_w,
                           int upper_h, int stride, int pixels, long long xs_w = 0, long long xs_h = 0, long long xs_n = 0) {
  if (xs_w <= 0) xs_w = c;
  if (xs_h <= 0) xs_h = (long long)w * xs_w;
  if (xs_n <= 0) xs_n = (long long)h * xs_h;
  static thread_local Im2colSlot* cache = nullptr;
  if (!cache) {
    cache = static_cast<Im2colSlot*>(aligned_alloc(64, kIm2colCacheSlots * sizeof(Im2colSlot)));
    if (cache) memset(cache, 0, kIm2colCacheSlots * sizeof(Im2colSlot));
  }
  Im2colKey key;
  memset(&key, 0, sizeof(key));     // padding bytes take part in nothing, but keep the slot image deterministic
  key.ptr = ptr; key.n = n; key.h = h; key.w = w; key.c = c; key.lower_w = lower_w; key.lower_h = lower_h;
  key.upper_w = upper_w; key.upper_h = upper_h; key.stride = stride; key.pixels = pixels;
  key.xs_w = xs_w; key.xs_h = xs_h; key.xs_n = xs_n;
  uint64_t hsh = (uint64_t)(uintptr_t)ptr * 0x9E3779B97F4A7C15ull;
  hsh ^= ((uint64_t)n << 48) ^ ((uint64_t)h << 36) ^ ((uint64_t)w << 24) ^ ((uint64_t)c << 8) ^ ((uint64_t)pixels << 2) ^
         (uint64_t)(lower_w + 8) ^ ((uint64_t)stride << 5);
  Im2colSlot* slot = cache ? &cache[(hsh >> 22) % kIm2colCacheSlots] : nullptr;
  if (slot && slot->valid && slot->key == key) {
    memcpy(m, &slot->map, sizeof(CUtensorMap));
    return 0;
  }
  EncodeIm2colFn enc = get_encode_im2col();
  if (!enc) return (int)cudaErrorNotSupported;
  cuuint64_t gdim[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t gstr[3] = {(cuuint64_t)xs_w * 2, (cuuint64_t)xs_h * 2, (cuuint64_t)xs_n * 2};
  int lower[2] = {lower_w, lower_h};
  int upper[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, lower, upper, 64,
                   (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[dlcfn] cuTensorMapEncodeIm2col failed (%d): ptr=%p nhwc=%dx%dx%dx%d lower=(%d,%d) upper=(%d,%d) stride=%d "
            "pixels=%d\n", (int)r, ptr, n, h, w, c, lower_w, lower_h, upper_w, upper_h, stride, pixels);
    return 10000 + (int)r;
  }
  // Driver issue with im2col descriptors of tensors smaller than 128 KiB (drivers up to CUDA 13.1; CUTLASS carries the same
  // fix-up in its im2col descriptor builder): bit 21 of the descriptor's second 64-bit word must be cleared.
  {
    static int driver_version = -1;
    if (driver_version < 0 && cudaDriverGetVersion(&driver_version) != cudaSuccess) driver_version = 0;
    const unsigned long long bytes = (unsigned long long)n * h * w * c * 2ull;
    if (driver_version <= 13010 && bytes < 131072ull) reinterpret_cast<uint64_t*>(m)[1] &= ~(1ull << 21);
  }
  if (slot) {
    slot->key = key;
    memcpy(&slot->map, m, sizeof(CUtensorMap));
    slot->valid = true;
  }
  return 0;
}

// EXPERIMENTAL multicast-pair launch (BLOCK_N = 256, local B, no split-K): clusters of 2 CTAs, grid = 2 x pairs
template <bool A_MN, bool B_MN>
static int launch_mc(GemmParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = GemmCfg<256, A_MN, B_MN>;
  auto kern = kgemm<256, A_MN, B_MN, true>;
  static bool configured = false;
  if (!configured) {
    cudaError_t ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (ae != cudaSuccess) return (int)ae;
    configured = true;
  }
  p.splits = 1;
  p.kb_per_split = (p.K + kBlockK - 1) / kBlockK;
  p.ws = nullptr;
  p.stages = Cfg::kStages;
  p.b_resident = 0;
  p.bres_bytes = 0;
  p.stg_bufs = 1;
  const int num_m = (p.M + kBlockM - 1) / kBlockM, num_n = (p.N + 255) / 256;
  const int pair_tiles = ((num_m + 1) / 2) * num_n;
  int pairs = num_sms / 2;
  if (pairs > pair_tiles) pairs = pair_tiles;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
  cfg.blockDim = dim3(kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
  if (le != cudaSuccess)
    fprintf(stderr, "[dlcfn] kgemm (multicast pair) launch failed: %s (grid %d)\n", cudaGetErrorString(le), 2 * pairs);
  return (int)le;
}


// EXPERIMENTAL CTA-pair launch (256 x 256 pair tiles, local B, bf16 TMA-store output, no split-K)
template <bool A_MN, bool B_MN>
static int launch_pair(GemmParams& p, int num_sms, cudaStream_t stream) {
  auto kern = kgemm_pair<A_MN, B_MN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t ae = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, k2SmemBytes);
    if (ae != cudaSuccess) return (int)ae;
    configured = true;
  }
  p.splits = 1;
  p.kb_per_split = (p.K + kBlockK - 1) / kBlockK;
  p.ws = nullptr;
  const int num_m = (p.M + kBlockM - 1) / kBlockM, num_n = (p.N + k2BlockN - 1) / k2BlockN;
  const int pair_tiles = ((num_m + 1) / 2) * num_n;
  // raster group: as many M-pair rows as keep the group's A panel (rows x K bf16) within ~1/4 of the 126 MB L2, at least 2
  // rows, and only when there are several N tiles to walk (otherwise M-fastest order is already optimal)
  {
    static const int rg_env = [] { const char* e = getenv("DLCFN_GEMM_RASTER"); return e ? atoi(e) : -1; }();
    const long long row_bytes = 256ll * p.K * 2;
    int rg = (int)((32ll << 20) / (row_bytes > 0 ? row_bytes : 1));
    if (rg < 2) rg = 2;
    if (rg > 16) rg = 16;
    p.raster_g = (num_n >= 2) ? rg : 0;
    if (rg_env >= 0) p.raster_g = rg_env;
  }
  int pairs = num_sms / 2;
  if (pairs > pair_tiles) pairs = pair_tiles;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
  cfg.blockDim = dim3(kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = k2SmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, p);
  if (le != cudaSuccess)
    fprintf(stderr, "[dlcfn] kgemm_pair launch failed: %s (grid %d, smem %d)\n", cudaGetErrorString(le), 2 * pairs, k2SmemBytes);
  return (int)le;
}

}  // namespace dlcfn

using namespace dlcfn;

// Host evaluation of the launch-time divisor arithmetic the producers use on the device (make_fastdiv + the expression of
// fdiv), exported for the CPU test tier: quotient of n / d for 0 <= n < 2^31, d >= 1.
DLCFN_API int dlcfn_fastdiv_host(int n, int d) {
  const FastDiv f = make_fastdiv(d);
  return (int)(((unsigned long long)(uint32_t)n * f.mul) >> f.shr);
}

// a_mn / b_mn: 0 = K-major ([rows, K] row-major), 1 = MN-major ([K, rows] row-major).
// b_shards: array of n_shards device pointers (peer-mapped); shard s holds rows
//   [s*rows_per_shard, (s+1)*rows_per_shard) of B's ROW dimension as stored
//   (K-major: output columns n; MN-major: contraction index k).  n_shards == 1: plain local GEMM.
// lda / ldb: row strides in elements of the stored matrices.  d_dtype: 0 fp32, 1 bf16.
// stats (nullable): fp32 [dlcfn_gemm_stats_rows()][2][N] partial column sums / sums of squares of the stored bf16 output
// (fused BatchNorm statistics, see kgemm); needs the bf16 TMA-store output form.
DLCFN_API int dlcfn_gemm_stats_rows() {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  return 4 * sms;
}

DLCFN_API int dlcfn_gemm_bf16_ex(const void* A, long long lda, int a_mn, const void* const* b_shards, int n_shards,
                                 int rows_per_shard, long long ldb, int b_mn, void* D, long long ldd, int d_dtype,
                                 const float* bias, int M, int N, int K, int accumulate, int relu, int block_n,
                                 int num_sms, float* stats, cudaStream_t stream);

DLCFN_API int dlcfn_gemm_bf16(const void* A, long long lda, int a_mn, const void* const* b_shards, int n_shards,
                              int rows_per_shard, long long ldb, int b_mn, void* D, long long ldd, int d_dtype,
                              const float* bias, int M, int N, int K, int accumulate, int relu, int block_n,
                              int num_sms, cudaStream_t stream) {
  return dlcfn_gemm_bf16_ex(A, lda, a_mn, b_shards, n_shards, rows_per_shard, ldb, b_mn, D, ldd, d_dtype, bias, M, N, K,
                            accumulate, relu, block_n, num_sms, nullptr, stream);
}

struct StagedPull {
  SymmCtx* ctx;
  size_t shard_off, stage_off, flags_off;
  int rows_per_shard, cols, channel;
  uint32_t seq;
};
static int gemm_impl(const void* A, long long lda, int a_mn, const void* const* b_shards, int n_shards, int rows_per_shard,
                     long long ldb, int b_mn, void* D, long long ldd, int d_dtype, const float* bias, int M, int N, int K,
                     int accumulate, int relu, int block_n, int num_sms, float* stats, const StagedPull* sp, cudaStream_t stream);

DLCFN_API int dlcfn_gemm_bf16_ex(const void* A, long long lda, int a_mn, const void* const* b_shards, int n_shards,
                                 int rows_per_shard, long long ldb, int b_mn, void* D, long long ldd, int d_dtype,
                                 const float* bias, int M, int N, int K, int accumulate, int relu, int block_n,
                                 int num_sms, float* stats, cudaStream_t stream) {
  return gemm_impl(A, lda, a_mn, b_shards, n_shards, rows_per_shard, ldb, b_mn, D, ldd, d_dtype, bias, M, N, K, accumulate, relu,
                   block_n, num_sms, stats, nullptr, stream);
}

// All-gather fused into the GEMM (see GemmParams::sg): B's stored matrix [world * rows_per_shard][cols] is row-sharded over
// the ranks of the symmetric heap `ctx` (this rank's shard at heap offset shard_off); stage_off / flags_off are heap offsets
// of a [world * rows_per_shard][cols] bf16 stage buffer and of world * rows_per_shard / 64 arrival counters (zero at
// allocation, never reset); `seq` = 1, 2, 3, ... counts the launches that used THESE counters (same on every rank).  Every
// rank of the heap must launch the same call.  b_mn = 0: B = stored matrix is [N][K] (forward); b_mn = 1: [K][N] (dgrad).
DLCFN_API int dlcfn_gemm_bf16_staged(void* ctx, const void* A, long long lda, int a_mn, size_t shard_off, size_t stage_off,
                                     size_t flags_off, unsigned seq, int rows_per_shard, int cols, int b_mn, void* D,
                                     long long ldd, int d_dtype, const float* bias, int M, int N, int K, int accumulate,
                                     int relu, int block_n, int channel, cudaStream_t stream) {
  SymmCtx* c = (SymmCtx*)ctx;
  if (c == nullptr || !c->finalized || c->dev.mc == nullptr) return (int)cudaErrorNotSupported;
  if (rows_per_shard <= 0 || (rows_per_shard % 64) || cols <= 0 || (cols % 8) || (shard_off & 15) || (stage_off & 15) ||
      (flags_off & 3) || channel >= kSigChannels)       // channel < 0: no entry barrier (the caller alternates two stage buffers)
    return (int)cudaErrorInvalidValue;
  const int rows_total = rows_per_shard * c->world;
  if (!b_mn && (N > rows_total || K != cols)) return (int)cudaErrorInvalidValue;
  if (b_mn && (K > rows_total || N > cols)) return (int)cudaErrorInvalidValue;
  StagedPull sp{c, shard_off, stage_off, flags_off, rows_per_shard, cols, channel, seq};
  const void* stage_local = c->dev.base[c->rank] + stage_off;
  const void* shards[1] = {stage_local};
  return gemm_impl(A, lda, a_mn, shards, 1, 0, cols, b_mn, D, ldd, d_dtype, bias, M, N, K, accumulate, relu, block_n, 0, nullptr,
                   &sp, stream);
}

static int gemm_impl(const void* A, long long lda, int a_mn, const void* const* b_shards, int n_shards, int rows_per_shard,
                     long long ldb, int b_mn, void* D, long long ldd, int d_dtype, const float* bias, int M, int N, int K,
                     int accumulate, int relu, int block_n, int num_sms, float* stats, const StagedPull* sp, cudaStream_t stream) {
  // cuTensorMapEncodeTiled is a driver call and needs a context current on THIS thread.  Autograd
  // worker threads for device 0 have none until the first runtime call (torch skips cudaSetDevice
  // when the ordinal already matches), so bind the primary context once per thread.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    DLCFN_CHECK(cudaFree(nullptr));
    ctx_bound = true;
  }
  (void)cudaGetLastError();  // drop stale non-sticky errors left by other libraries
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (n_shards < 1 || n_shards > kMaxShards) return (int)cudaErrorInvalidValue;
  if ((lda & 7) || (ldb & 7) || (((uintptr_t)A) & 15)) return (int)cudaErrorInvalidValue;
  if (num_sms <= 0) {
    int dev = 0;
    DLCFN_CHECK(cudaGetDevice(&dev));
    DLCFN_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (block_n == 64 && a_mn) block_n = 128;      // the 64-wide tile exists for the activation-operand forms only
  if (block_n != 64 && block_n != 128 && block_n != 256 && N <= 64 && !a_mn && n_shards == 1) block_n = 64;
  if (block_n != 64 && block_n != 128 && block_n != 256) {
    // 128x256 tiles feed the tensor pipe better (measured 1.35 vs 1.0 PFLOP/s at 8192^3) but halve the tile count.
    // Score each width by  pipe rate x useful fraction of the padded tile grid x occupancy of the last wave  and take
    // the better one; e.g. M=8192, N=1024: 256 wide tiles = 1.73 waves (86 %) x 1.35 beats 512 narrow tiles (86 %) x 1.0.
    auto score = [&](int bn, double rate) {
      const long long tm = (M + kBlockM - 1) / kBlockM, tn = (N + bn - 1) / bn;
      const long long tiles = tm * tn;
      const double useful = ((double)M * (double)N) / ((double)(tm * kBlockM) * (double)(tn * bn));
      const long long waves = (tiles + num_sms - 1) / num_sms;
      return rate * useful * ((double)tiles / (double)(waves * num_sms));
    };
    block_n = (N >= 256 && score(256, 1.35) > score(128, 1.0)) ? 256 : 128;
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.n_shards = n_shards;
  p.rows_per_shard = rows_per_shard;
  p.D = D; p.ldd = ldd; p.bias = bias;
  p.d_is_fp32 = d_dtype == 0;
  p.accumulate = accumulate;
  p.relu = relu;
  p.stats = stats;
  static const bool splitk_env = [] { const char* e = getenv("DLCFN_GEMM_SPLITK"); return !(e && e[0] == '0'); }();
  p.split_k_ok = (splitk_env && stats == nullptr) ? 1 : 0;
  if (sp != nullptr) {
    p.sg = sp->ctx->dev;
    p.sg_on = 1;
    p.sg_channel = sp->channel;
    p.sg_shard = sp->ctx->dev.base[sp->ctx->rank] + sp->shard_off;
    p.sg_stage_off = sp->stage_off;
    p.sg_flags_off = sp->flags_off;
    p.sg_rows_per_shard = sp->rows_per_shard;
    p.sg_cols = sp->cols;
    p.sg_ld = sp->cols;
    p.sg_want = sp->seq * (uint32_t)((sp->cols + 255) / 256);
  }
  // EXPERIMENTAL multicast-pair kernel: opt-in, wide tiles, local B, enough tile pairs to fill the machine, and a
  // reduction too short for split-K to matter (it is not combined with split-K)
  static const bool mc_env = [] { const char* e = getenv("DLCFN_GEMM_MULTICAST"); return e && e[0] == '1'; }();
  const bool use_mc = mc_env && sp == nullptr && stats == nullptr && block_n == 256 && n_shards == 1 &&
                      (long long)(((M + kBlockM - 1) / kBlockM + 1) / 2) * ((N + 255) / 256) >= num_sms / 2;
  // CTA-pair kernel (cta_group::2, validated on B200 in round 2: 983 vs 629 TFLOP/s single-CTA at 16384x1024x1024): default
  // for bf16 TMA-store outputs of GEMMs with enough 256 x 256 pair tiles; DLCFN_GEMM_2CTA=0 disables
  static const bool pair_env = [] { const char* e = getenv("DLCFN_GEMM_2CTA"); return !(e && e[0] == '0'); }();
  static const bool tma_store_env2 = [] { const char* e = getenv("DLCFN_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
  const bool use_pair = pair_env && stats == nullptr && !use_mc && n_shards == 1 && N >= 256 && tma_store_env2 && d_dtype == 1 && !accumulate &&
                        (ldd & 7) == 0 && (((uintptr_t)D) & 15) == 0 &&
                        (long long)(((M + kBlockM - 1) / kBlockM + 1) / 2) * ((N + 255) / 256) >= num_sms / 4;
  if (use_pair) block_n = 256;
  int rc;
  if (!a_mn) rc = make_map(&p.tmA, A, M, K, lda, kBlockK, kBlockM);
  else rc = make_map(&p.tmA, A, K, M, lda, 64, kBlockK);
  if (rc) return rc;
  for (int s = 0; s < n_shards; ++s) {
    if (!b_mn) {
      const uint64_t rows = n_shards > 1 ? (uint64_t)rows_per_shard : (uint64_t)N;
      rc = make_map(&p.tmB[s], b_shards[s], rows, K, ldb, kBlockK, (use_mc || use_pair) ? block_n / 2 : block_n);   // MC / pair: half tile per CTA
    } else {
      const uint64_t rows = n_shards > 1 ? (uint64_t)rows_per_shard : (uint64_t)K;
      rc = make_map(&p.tmB[s], b_shards[s], rows, N, ldb, 64, kBlockK);
    }
    if (rc) return rc;
  }
  // bf16, non-accumulating output with a 16-byte aligned base / row stride goes out through TMA
  static const bool tma_store_env = [] { const char* e = getenv("DLCFN_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
  if (tma_store_env && d_dtype == 1 && !accumulate && N >= 64 && (ldd & 7) == 0 && (((uintptr_t)D) & 15) == 0) {
    rc = make_map(&p.tmD, D, M, N, ldd, 64, 32);
    if (rc) return rc;
    p.tma_store = 1;
  }
  if (stats != nullptr && !p.tma_store) return (int)cudaErrorInvalidValue;
  if (n_shards > 1) {
    // a tile must not straddle two shards
    if (!b_mn && (rows_per_shard % block_n)) return (int)cudaErrorInvalidValue;
    if (b_mn && (rows_per_shard % kBlockK)) return (int)cudaErrorInvalidValue;
  }
  if (use_pair) {
    if (!a_mn && !b_mn) return launch_pair<false, false>(p, num_sms, stream);
    if (!a_mn && b_mn) return launch_pair<false, true>(p, num_sms, stream);
    if (a_mn && !b_mn) return launch_pair<true, false>(p, num_sms, stream);
    return launch_pair<true, true>(p, num_sms, stream);
  }
  if (use_mc) {
    if (!a_mn && !b_mn) return launch_mc<false, false>(p, num_sms, stream);
    if (!a_mn && b_mn) return launch_mc<false, true>(p, num_sms, stream);
    if (a_mn && !b_mn) return launch_mc<true, false>(p, num_sms, stream);
    return launch_mc<true, true>(p, num_sms, stream);
  }
#define DLCFN_GEMM_CASE(BN, AM, BM) \
  if (block_n == BN && (bool)a_mn == AM && (bool)b_mn == BM) return launch<BN, AM, BM>(p, num_sms, stream);
  DLCFN_GEMM_CASE(64, false, false)
  DLCFN_GEMM_CASE(64, false, true)
  DLCFN_GEMM_CASE(128, false, false)
  DLCFN_GEMM_CASE(128, false, true)
  DLCFN_GEMM_CASE(128, true, false)
  DLCFN_GEMM_CASE(128, true, true)
  DLCFN_GEMM_CASE(256, false, false)
  DLCFN_GEMM_CASE(256, false, true)
  DLCFN_GEMM_CASE(256, true, false)
  DLCFN_GEMM_CASE(256, true, true)
#undef DLCFN_GEMM_CASE
  return (int)cudaErrorInvalidValue;
}

// Implicit-GEMM convolution, NHWC bf16 (kgemm<..., CONV>).  Square stride, symmetric padding, no dilation.
//   op 0 fprop:  x [n,h,w,c]   * wt [k,r,s,c] -> out = y [n,p,q,k]            (any stride)
//   op 1 dgrad:  x = dy [n,p,q,k] (pass its dims as n,h,w,c = n,p,q,k), wt [k,r,s,cin] -> out = dx [n,oh,ow,cin]; any stride:
//                a strided dgrad runs one implicit GEMM per output-parity class (the taps that reach pixels of that
//                parity), each storing to its strided positions of dx; classes no tap reaches are zero-filled
//   op 2 wgrad:  x [n,h,w,c], other = dy [n,p,q,k] -> out = dw [k,r,s,c]
// `k` is always the filter count (Cout), `cin` the filter depth; channel counts of the im2col'd tensor must be multiples
// of 64 (one 128-byte swizzle row per k-block), all others multiples of 8.
static int conv_launch(GemmParams& p, int block_n, bool a_mn, bool b_mn, int num_sms, cudaStream_t stream) {
#define DLCFN_CONV_CASE(BN, AM, BM) \
  if (block_n == BN && a_mn == AM && b_mn == BM) return launch<BN, AM, BM, true>(p, num_sms, stream);
  DLCFN_CONV_CASE(64, false, false)
  DLCFN_CONV_CASE(128, false, false)
  DLCFN_CONV_CASE(256, false, false)
  DLCFN_CONV_CASE(64, false, true)
  DLCFN_CONV_CASE(128, false, true)
  DLCFN_CONV_CASE(256, false, true)
  DLCFN_CONV_CASE(128, true, true)
  DLCFN_CONV_CASE(256, true, true)
#undef DLCFN_CONV_CASE
  return (int)cudaErrorInvalidValue;
}

// _ex: `xs_w / xs_h / xs_n` (elements; 0 = dense NHWC) are the strides of the im2col'd tensor's w / h / n dimensions - an
// OVERLAPPING view is allowed: ResNet's stem reads its 16-channel space-to-depth image as [n, 115, 112, 64] with xs_w = 16,
// i.e. a "pixel" of 64 channels = 4 adjacent columns x 16 channels, which turns the 4x4 filter into 4 (r) x 1 taps of whole
// 128-byte rows.  `stats`: fused BatchNorm statistics of the fprop output (see dlcfn_gemm_bf16_ex).
DLCFN_API int dlcfn_conv_bf16_ex(int op, const void* x, const void* other, void* out, int out_dtype, int n, int h, int w, int c,
                                 int k, int cin, int r, int s, int pad, int stride, int oh, int ow, int accumulate,
                                 int block_n, int num_sms, long long xs_w, long long xs_h, long long xs_n, float* stats,
                                 cudaStream_t stream);

DLCFN_API int dlcfn_conv_bf16(int op, const void* x, const void* other, void* out, int out_dtype, int n, int h, int w, int c,
                              int k, int cin, int r, int s, int pad, int stride, int oh, int ow, int accumulate,
                              int block_n, int num_sms, cudaStream_t stream) {
  return dlcfn_conv_bf16_ex(op, x, other, out, out_dtype, n, h, w, c, k, cin, r, s, pad, stride, oh, ow, accumulate, block_n,
                            num_sms, 0, 0, 0, nullptr, stream);
}

DLCFN_API int dlcfn_conv_bf16_ex(int op, const void* x, const void* other, void* out, int out_dtype, int n, int h, int w, int c,
                                 int k, int cin, int r, int s, int pad, int stride, int oh, int ow, int accumulate,
                                 int block_n, int num_sms, long long xs_w, long long xs_h, long long xs_n, float* stats,
                                 cudaStream_t stream) {
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    DLCFN_CHECK(cudaFree(nullptr));
    ctx_bound = true;
  }
  (void)cudaGetLastError();
  if (op < 0 || op > 2 || n <= 0 || h <= 0 || w <= 0 || r <= 0 || s <= 0 || stride <= 0 || pad < 0) return (int)cudaErrorInvalidValue;
  if ((c % 64) || (k % 8) || (cin % 8) || r * s > 49) return (int)cudaErrorInvalidValue;
  if ((((uintptr_t)x) & 15) || (((uintptr_t)other) & 15) || (((uintptr_t)out) & 15)) return (int)cudaErrorInvalidValue;
  if (num_sms <= 0) {
    int dev = 0;
    DLCFN_CHECK(cudaGetDevice(&dev));
    DLCFN_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.n_shards = 1;
  p.D = out;
  p.d_is_fp32 = out_dtype == 0;
  p.accumulate = accumulate;
  static const bool splitk_env = [] { const char* e = getenv("DLCFN_GEMM_SPLITK"); return !(e && e[0] == '0'); }();
  p.split_k_ok = splitk_env ? 1 : 0;
  if (stats != nullptr && op != 0) return (int)cudaErrorInvalidValue;
  p.stats = stats;
  if (stats != nullptr) p.split_k_ok = 0;
  p.cv_S = s;
  p.cv_RS = r * s;
  p.cv_cblocks = c / 64;
  p.cv_stride = stride;
  p.cv_cin = cin;
  static const bool tma_store_env = [] { const char* e = getenv("DLCFN_GEMM_TMA_STORE"); return !(e && e[0] == '0'); }();
  int rc;
  if (op == 0 || op == 2) {
    // im2col of the forward input: base positions = output pixels, taps in storage order
    const int P = (h + 2 * pad - r) / stride + 1, Q = (w + 2 * pad - s) / stride + 1;
    if (P != oh || Q != ow || c != cin) return (int)cudaErrorInvalidValue;
    p.cv_P = P; p.cv_Q = Q;
    p.cv_lower_w = -pad; p.cv_lower_h = -pad;
    p.cv_ntaps = r * s;
    for (int t = 0; t < r * s; ++t) {
      p.cv_tap_ow[t] = (unsigned char)(t % s);
      p.cv_tap_oh[t] = (unsigned char)(t / s);
      p.cv_tap_wi[t] = (unsigned char)t;
    }
    // the last base position of a row is (Q-1)*stride - pad, which must be the box's upper end w - 1 + upper_w or just
    // inside it: upper = pad - (s - 1) is the exact end for stride 1 and at most stride - 1 positions further otherwise
    const int upper_w = pad - (s - 1), upper_h = pad - (r - 1);
    bool a_mn, b_mn;
    if (op == 0) {
      p.M = n * P * Q; p.N = k; p.K = r * s * c;
      p.ldd = k;
      rc = make_map_im2col(&p.tmA, x, n, h, w, c, -pad, -pad, upper_w, upper_h, stride, kBlockM, xs_w, xs_h, xs_n);
      if (rc) return rc;
      if (block_n != 64 && block_n != 128 && block_n != 256) block_n = (k >= 256) ? 256 : (k > 64 ? 128 : 64);
      rc = make_map(&p.tmB[0], other, k, (uint64_t)r * s * c, (uint64_t)r * s * c, kBlockK, block_n);   // W as [k][rs*c], K-major
      if (rc) return rc;
      a_mn = false; b_mn = false;
    } else {
      p.M = k; p.N = r * s * c; p.K = n * P * Q;
      p.ldd = (long long)r * s * c;
      rc = make_map(&p.tmA, other, (uint64_t)n * P * Q, k, k, 64, kBlockK);                       // dy as [npq][k], MN-major A
      if (rc) return rc;
      rc = make_map_im2col(&p.tmB[0], x, n, h, w, c, -pad, -pad, upper_w, upper_h, stride, kBlockK, xs_w, xs_h, xs_n);
      if (rc) return rc;
      // every (tile, k-slice) work item streams its whole slice of BOTH operands through the L2 -> SM crossbar, which is the
      // bound of these kernels: wide tiles halve the re-reads of dy (profiles/r2/conv_ncu_before.md)
      if (block_n != 128 && block_n != 256) block_n = (p.N >= 512) ? 256 : 128;
      a_mn = true; b_mn = true;
    }
    if (tma_store_env && out_dtype == 1 && !accumulate && p.N >= 64 && (p.ldd & 7) == 0) {
      rc = make_map(&p.tmD, out, p.M, p.N, p.ldd, 64, 32);
      if (rc) return rc;
      p.tma_store = 1;
    }
    if (stats != nullptr && !p.tma_store) return (int)cudaErrorInvalidValue;
    return conv_launch(p, block_n, a_mn, b_mn, num_sms, stream);
  }
  // ---------------------------------------------------------------- dgrad: x is dy [n, h, w, c = k]; out = dx [n, oh, ow, cin]
  if (c != k) return (int)cudaErrorInvalidValue;
  if ((h + 0) != (oh + 2 * pad - r) / stride + 1 || w != (ow + 2 * pad - s) / stride + 1) return (int)cudaErrorInvalidValue;
  rc = make_map(&p.tmB[0], other, k, (uint64_t)r * s * cin, (uint64_t)r * s * cin, 64, kBlockK);  // W as [k][rs*cin], MN-major B
  if (rc) return rc;
  if (block_n != 64 && block_n != 128 && block_n != 256) block_n = (cin >= 256) ? 256 : (cin > 64 ? 128 : 64);
  p.ldd = cin;
  p.N = cin;
  p.cv_stride = 1;                 // base positions walk dy's grid with unit steps in every class
  // per output-parity class (ph, pw): taps with (ph + pad - r') % stride == 0 reach it, displaced by (ph + pad - r') / stride
  auto taps_1d = [&](int phase, int taps, int* idx, int* disp) {
    int cnt = 0;
    for (int t = 0; t < taps; ++t) {
      const int v = phase + pad - t;
      if (((v % stride) + stride) % stride != 0) continue;
      idx[cnt] = t;
      disp[cnt] = (v >= 0) ? v / stride : -((-v) / stride);     // exact division
      ++cnt;
    }
    return cnt;
  };
  bool any_empty = false;
  for (int ph = 0; ph < stride && ph < oh; ++ph)
    for (int pw = 0; pw < stride && pw < ow; ++pw) {
      int ti[8], td[8];
      if (r > 8 || s > 8) return (int)cudaErrorInvalidValue;
      if (taps_1d(ph, r, ti, td) == 0 || taps_1d(pw, s, ti, td) == 0) any_empty = true;
    }
  if (any_empty && !accumulate)
    DLCFN_CHECK(cudaMemsetAsync(out, 0, (size_t)n * oh * ow * cin * (out_dtype == 0 ? 4 : 2), stream));
  for (int ph = 0; ph < stride && ph < oh; ++ph) {
    for (int pw = 0; pw < stride && pw < ow; ++pw) {
      int ri[8], rd[8], si[8], sd[8];
      const int nr = taps_1d(ph, r, ri, rd), ns = taps_1d(pw, s, si, sd);
      if (nr == 0 || ns == 0) continue;
      const int Hp = (oh - ph + stride - 1) / stride, Wp = (ow - pw + stride - 1) / stride;
      int dmin_h = rd[0], dmin_w = sd[0];
      for (int t = 1; t < nr; ++t) dmin_h = rd[t] < dmin_h ? rd[t] : dmin_h;
      for (int t = 1; t < ns; ++t) dmin_w = sd[t] < dmin_w ? sd[t] : dmin_w;
      GemmParams q = p;
      q.cv_ntaps = nr * ns;
      for (int a = 0; a < nr; ++a)
        for (int b = 0; b < ns; ++b) {
          const int t = a * ns + b;
          q.cv_tap_oh[t] = (unsigned char)(rd[a] - dmin_h);
          q.cv_tap_ow[t] = (unsigned char)(sd[b] - dmin_w);
          q.cv_tap_wi[t] = (unsigned char)(ri[a] * s + si[b]);
        }
      q.cv_P = Hp; q.cv_Q = Wp;
      q.cv_lower_w = dmin_w; q.cv_lower_h = dmin_h;
      q.M = n * Hp * Wp;
      q.K = nr * ns * k;
      // base positions per row = w + upper - lower = Wp
      rc = make_map_im2col(&q.tmA, x, n, h, w, c, dmin_w, dmin_h, dmin_w + Wp - w, dmin_h + Hp - h, 1, kBlockM);
      if (rc) return rc;
      if (stride == 1) {
        if (tma_store_env && out_dtype == 1 && !accumulate && q.N >= 64 && (q.ldd & 7) == 0) {
          rc = make_map(&q.tmD, out, q.M, q.N, q.ldd, 64, 32);
          if (rc) return rc;
          q.tma_store = 1;
        }
      } else {
        q.out_remap = 1;
        q.split_k_ok = 0;          // the split-K reduce kernel stores plain rows
        q.out_H = oh; q.out_W = ow; q.out_step = stride; q.out_h0 = ph; q.out_w0 = pw;
      }
      rc = conv_launch(q, block_n, false, true, num_sms, stream);
      if (rc) return rc;
    }
  }
  return 0;
}
