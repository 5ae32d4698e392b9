// Warp-specialised, mbarrier-pipelined tcgen05 kernels of the F-FNO layer (width 64 / hidden 256), sm_100a.
//
// One CTA per SM, 9 warps:
//   warps 0-3  epilogue   (TMEM lanes 0..127 -> registers -> bias/activation/split or global stores)
//   warp  4    MMA issuer (lane 0 issues tcgen05.mma + tcgen05.commit)
//   warps 5-8  loaders    (coalesced FP32 global loads -> BF16 hi/lo split -> swizzled operand tiles in smem)
// Operand tiles in shared memory and accumulators in tensor memory are double-buffered and handed over with
// full/empty mbarriers, so the global loads of tile t+1, the MMAs of tile t and the epilogue of tile t-1 overlap.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "umma.cuh"
#include "umma_kernels.cuh"

namespace ffno {

extern thread_local long long g_launch_counter;
using namespace umma;

// Optional in-kernel timeline (diagnostics): block 0 records clock64() at pipeline events into g_timeline
// [role 8][tile 16][event 8]; enabled by ffno_debug_timeline(1).  Costs one predicated store per event.
__device__ long long g_timeline[8 * 16 * 8];
__device__ int g_timeline_on = 0;
#ifdef FFNO_TIMELINE
// g_timeline_on selects the kernel that records: 1 FF, 2 forward transforms, 3 inverse transforms, 4 mode mix
// (tl_on is read from g_timeline_on ONCE at kernel start: a flag load per stamp would put an L2 round trip in front of
// every clock read and distort the very timeline it records)
#define TL(role, tile_n, ev)                                                                      \
  do {                                                                                            \
    if (tl_on && blockIdx.x == 0 && (tile_n) < 16 && (threadIdx.x & 31) == 0)                     \
      g_timeline[((role) * 16 + (tile_n)) * 8 + (ev)] = clock64();                                \
  } while (0)
#else
#define TL(role, tile_n, ev) do {} while (0)
#endif

namespace {

// Programmatic dependent launch: the kernel may begin (barrier init, TMEM allocation, constant-operand loads) while
// its predecessor on the stream drains; griddepcontrol.wait then blocks until the predecessor has fully completed.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// axis / mix kernels: 4 epilogue warps, 1 MMA warp, 16 loader-converter warps.  The FP32 -> BF16 hi/lo conversion is
// ~25 instructions per float4 and is what bounds these kernels, so it is spread over as many warps as fit.
constexpr int kLoaders = 512;
constexpr int kEpiWarps = 8;                   // two teams of 4 (one per TMEM lane quadrant), splitting the columns
// Warp order: epilogue warps first, then the MMA issuer, then the converter / loader warps.  (Tried: converters
// first so that the short, latency-critical roles get the higher warp ids, which the warp schedulers of this
// architecture family are said to favour — 1.90 -> 2.01 ms per forward, rejected.)
constexpr int kEpiWarp0 = 0;                              // warps 0..7 (warp & 3 = TMEM lane quadrant)
constexpr int kMmaWarp = kEpiWarp0 + kEpiWarps;           // warp 8
constexpr int kLoaderThread0 = (kMmaWarp + 1) * 32;       // warps 9..24
constexpr int kThreads = kLoaderThread0 + kLoaders;       // 800
constexpr int kPollWarp = kThreads / 32, kPubWarp = kPollWarp + 1;   // pipelined kernels: +2 synchronisation warps
constexpr int kPipeThreads = kThreads + 64;
// axis transform: the activation tiles arrive by TMA, issued by one producer warp (which in the pipelined kernel also
// does the poller's job); the pipelined kernel adds the publisher warp
constexpr int kAxProdWarp = kThreads / 32, kAxPubWarp = kAxProdWarp + 1;
constexpr int kAxThreads = kThreads + 32, kAxPipeThreads = kThreads + 64;
constexpr int kLdPerThread = 2048 / kLoaders;  // float4 per thread per 32 KB work item

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// TMA tile load (SASS: UTMALDG): one 4-D box of the tensor described by `tmap` -> shared memory, completion (bytes) on
// an mbarrier.  Out-of-range coordinates read as zeros, which is how tails and dead groups are padded.
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ float4 ldg_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
// Bulk L2 prefetch of a contiguous, 16-byte aligned range (no registers, no completion tracking).
__device__ __forceinline__ void prefetch_l2_bulk(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}
// Ampere-style async copy (SASS: LDGSTS): 16 bytes global -> shared, L2-only; src_bytes = 0 zero-fills.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 3-pass BF16 product over KSTEPS K steps of 16 (K-major SW128 operands within one 64-wide K block): descriptors are
// formed by adding the K offset (32 B >> 4 = 2 per step) to precomputed bases, fully unrolled — the single issuing
// thread must spend a handful of instructions per tcgen05.mma, not a descriptor rebuild, or it becomes the bottleneck.
template <int KSTEPS>
__device__ __forceinline__ void issue3_kmajor(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo,
                                              uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const uint64_t a = (pass == 2) ? a_lo : a_hi;
    const uint64_t b = (pass == 1) ? b_lo : b_hi;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
      umma_bf16_ss(tmem_d, a + (uint64_t)(ks * 2), b + (uint64_t)(ks * 2), idesc, (pass == 0 && ks == 0) ? acc_first : 1u);
  }
}

#ifndef FFNO_KO
#define FFNO_KO 0
#endif
#define KO_TILES(x) ((FFNO_KO & 1) ? 0 : (x))
#define KO_PASSES(bit) ((FFNO_KO & (bit)) ? 1 : 3)
template <int KSTEPS>
__device__ __forceinline__ void issue3_kmajor_elect(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi,
                                                    uint64_t b_lo, uint32_t idesc, uint32_t acc_first, int npass = 3) {
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    if (pass >= npass) break;
    const uint64_t a = (pass == 2) ? a_lo : a_hi;
    const uint64_t b = (pass == 1) ? b_lo : b_hi;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
      umma_bf16_ss_elect(tmem_d, a + (uint64_t)(ks * 2), b + (uint64_t)(ks * 2), idesc,
                         (pass == 0 && ks == 0) ? acc_first : 1u);
  }
}

__device__ __forceinline__ void store_split4_at(uint8_t* tile_hi, uint8_t* tile_lo, uint32_t off, float4 v) {
  uint32_t h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  *reinterpret_cast<uint2*>(tile_hi + off) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(tile_lo + off) = make_uint2(l0, l1);
}


// =======================================================================================================
// Stage-pipelined execution (launch_stack_pipe): the transform / mix / inverse / FF kernels are launched ONCE per
// forward, side by side on disjoint SMs, and each walks the whole (layer, tile) sequence of its stage.  A "unit" is the
// smallest group of samples whose tiles align in every stage (two samples on the 64 x 64 grid); a stage may start a
// unit of layer l when its producer's arrival counter of that unit has reached the layer's target:
//     forward transforms <- FF of layer l-1      mode mix <- forward transforms      inverse <- mode mix      FF <- inverse
// Counters are cumulative over the layers of one forward (zeroed by a memset before the kernels), bumped once per
// epilogue warp and finished tile with a gpu-scope release, read with a gpu-scope acquire.  Every dependency points to
// an earlier element of the (layer, unit, stage) order, every CTA walks its tiles in that order and all CTAs of the
// four kernels are resident at once (grid sizes add up to at most the SM count, one CTA per SM), so waits always end;
// the chain FF(l-1,u) -> fwd(l,u) -> mix -> inv -> FF(l,u) also orders every reuse of the F / R / s / x buffers.
// =======================================================================================================
struct PipeDesc {
  int n_layers;                  // 0 = plain single-launch behaviour, everything below ignored
  int tiles_per_unit[3];         // per axis: tiles of this kernel (per mode for the mix) that make up one unit
  int wait_lag;                  // the loader of (layer l, unit u) waits for wait_ctr[u] >= (l + wait_lag) * wait_count
  unsigned wait_count[3];        // per axis: arrivals the producer posts per layer and unit
  const unsigned* wait_ctr[3];   // per axis: the producer's counters [unit]; NULL = nothing to wait for
  unsigned* done_ctr[3];         // per axis: this stage's counters [unit]
  unsigned long long* dbg_ts;    // diagnostics: [64] globaltimer at which CTA 0's loaders found (layer, unit) ready
  unsigned long long* dbg;       // diagnostics (FFNO_B200_PIPE_DEBUG=1): per CTA {cycles blocked on the producer, cycles
                                 // in the kernel, start, end (globaltimer ns)} of the polling loader warp; NULL = off
};
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
constexpr int kCtrStride = 32;     // unsigned per counter: every counter has its own 128-byte line (no false sharing
                                   // between the line a producer bumps and the lines other consumers poll)
// One thread: block until *ctr >= target (bounded, like the mbarrier waits: a protocol bug must surface as an error).
// Relaxed polls + one acquire fence at the end: a gpu-scope acquire invalidates the SM's L1, once per wait is enough.
__device__ __forceinline__ void pipe_wait(const unsigned* ctr, unsigned target) {
  if (target == 0u) return;
  if (ld_relaxed_gpu(ctr) < target) {
    const long long t0 = clock64();
    while (ld_relaxed_gpu(ctr) < target) {
      __nanosleep(100);
      if (clock64() - t0 > (1ll << 33)) {
        atomicExch(&g_umma_timeout_flag, 2u);
        __trap();
      }
    }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
// Synchronisation is kept off the working warps: in the pipelined kernels two extra warps do it.
//  * The POLLER warp walks the CTA's tile sequence ahead of the loaders, waits (gpu-scope acquire) for each new unit's
//    producer counter and publishes the unit's sequence number in a shared-memory word; loaders / store warps only
//    watch that word (cta-scope acquire), so no gpu-scope fence ever stalls a warp with copies or stores in flight.
//  * The PUBLISHER warp waits until the epilogue warps have issued a tile's global stores (a shared-memory counter
//    they bump with cta-scope release after their last store), then posts ONE arrival for the tile with a gpu-scope
//    release.  The release is cumulative over the epilogue warps' stores (the same pattern as a grid barrier:
//    bar / cta-scope synchronisation first, one thread fences and signals), and only the publisher waits for them to
//    drain.
__device__ __forceinline__ int ld_acquire_cta_smem(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta_smem(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// one thread: block until the shared-memory word reaches `want` (bounded)
__device__ __forceinline__ void smem_wait_ge(const int* p, int want) {
  if (ld_acquire_cta_smem(p) >= want) return;
  const long long tb = clock64();
  while (ld_acquire_cta_smem(p) < want) {
    __nanosleep(40);
    if (clock64() - tb > (1ll << 33)) {
      atomicExch(&g_umma_timeout_flag, 3u);
      __trap();
    }
  }
}
// Whole warp (warp-uniform arguments): wait until the poller has seen unit `seq` ready.  Returns the cycles lane 0 waited.
__device__ __forceinline__ long long wait_unit_ready(const int* s_seq, int seq) {
  long long blocked = 0;
  if ((threadIdx.x & 31) == 0) {
    const long long ta = clock64();
    smem_wait_ge(s_seq, seq);
    blocked = clock64() - ta;
  }
  __syncwarp();
  return blocked;
}
// Whole epilogue warp, after its last global store of a tile: tell the publisher (cta-scope release).
__device__ __forceinline__ void stores_issued(int* s_pub) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) asm volatile("red.release.cta.shared::cta.add.s32 [%0], %1;" ::"r"(smem_u32(s_pub)), "r"(1) : "memory");
}
// Publisher thread: one arrival on the stage counter (gpu-scope release, cumulative over what it has observed).
__device__ __forceinline__ void pipe_publish(unsigned* ctr) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
}
// Coherent streaming load (L2 only): data written by another kernel DURING this kernel's lifetime must not come from
// the non-coherent path or from a stale L1 line.
__device__ __forceinline__ float4 ldg_cg(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

}  // namespace

// =======================================================================================================
// Axis transform (forward / inverse truncated DFT), up to 3 axes per launch (blockIdx.y = axis)
// =======================================================================================================
constexpr int AXP_A_STAGE = 32768;                 // operand stage: hi 16 KB | lo 16 KB
constexpr int AXP_BAR = 2 * AXP_A_STAGE;           // 65536
constexpr int AXP_STAGING = AXP_BAR + 1024;        // ring x 32 KB of raw FP32 (cp.async landing zone)
constexpr int kAxStages = 4;                       // deepest staging ring (power of two); shallower when a large table
                                                   // (C4: 128 KB) leaves no room.  The ring depth x 32 KB is what a CTA
                                                   // keeps in flight: the loaders are bound by latency / depth
constexpr int axp_table_offset(int ring) { return AXP_STAGING + ring * 32768; }   // the DFT table image follows the ring

struct AxisSet {
  AxisXform ax[3];
  int n_tiles[3];
  int tmem_cols[3];
  int ring[3];      // cp.async staging ring depth of the axis (1 or 2)
  int reverse;      // walk the tiles from the last to the first (see launch_axis_pipe)
  const float* X_odd[3];   // kPipe: input of the odd layers (the residual stream ping-pongs between two buffers)
  PipeDesc pipe;
  // TMA descriptors of the inputs, viewed as [outer][n_in][inner / 64][64] floats with a box of one 64 x 64 group
  alignas(64) CUtensorMap tm[3];
  alignas(64) CUtensorMap tm_odd[3];
};

// kPipe: the CTA walks tile, tile + grid, ... of every layer back to back (grid <= n_tiles, n_groups even), its loaders
// wait for the unit's producer and its epilogue warps post the unit's completion (see PipeDesc).
// kRmw: instantiation for launches with an accumulating axis (summed spectral output of the taps / standalone / backward
// paths): its epilogue batches the read-modify-write; kept out of the plain instantiation, whose register budget it breaks.
template <bool kPipe, bool kRmw = false>
__global__ void __launch_bounds__(kPipe ? kAxPipeThreads : kAxThreads, 1) axis_pipe_kernel(const __grid_constant__ AxisSet set) {
  const AxisXform& p = set.ax[blockIdx.y];
  const int n_tiles = set.n_tiles[blockIdx.y];
  if ((int)blockIdx.x >= n_tiles) return;
  const int n_layers = kPipe ? set.pipe.n_layers : 1;
  const int my_tiles = KO_TILES((int)(((long long)n_layers * n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x));
  const int tmem_cols = set.tmem_cols[blockIdx.y];
  const int stage_cols = tmem_cols >> 1;
  [[maybe_unused]] const bool tl_on = g_timeline_on == (p.n_in >= p.n_out ? 2 : 3);      // timeline build: forward / inverse

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AXP_BAR);
  uint64_t* a_full = bars;          // [2]
  uint64_t* a_empty = bars + 2;     // [2]
  uint64_t* d_full = bars + 4;      // [2]
  uint64_t* d_empty = bars + 6;     // [2]
  uint64_t* bar_w = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  int* s_pub = reinterpret_cast<int*>(bars + 10);      // [2] kPipe: warps of epilogue team t that have issued their stores
  uint64_t* stg_full = bars + 12;   // [4] staging slot landed (TMA transaction bytes)
  uint64_t* stg_empty = bars + 16;  // [4] every converter thread has taken its part of the slot
  const int ring = set.ring[blockIdx.y];
  const int ring_log2 = ring == 4 ? 2 : (ring == 2 ? 1 : 0);
  uint8_t* sB = smem + axp_table_offset(ring);
  const uint32_t b_half = (uint32_t)p.kchunks * p.npad * 128u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    s_pub[0] = s_pub[1] = 0;
    // the (constant) table image first: its copy runs under TMEM allocation, barrier set-up and the wait for the
    // predecessor grid instead of after them
    mbar_init(bar_w, 1);
    fence_barrier_init();
    {
      const uint32_t total = 2u * b_half;
      mbar_expect_tx(bar_w, total);
      for (uint32_t off = 0; off < total; off += 32768u) {
        const uint32_t nb = (total - off) < 32768u ? (total - off) : 32768u;
        bulk_g2s(sB + off, p.table + off, nb, bar_w);
      }
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], kLoaders);
      mbar_init(&a_empty[i], 1);
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], p.npad > 32 ? 256 : 128);   // see the epilogue: who drains a stage
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&stg_full[i], 1);
      mbar_init(&stg_empty[i], kLoaders);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long gpi = p.inner >> 6;
  const long long n_groups = p.outer * gpi;
  pdl_launch_dependents();
  if (warp != kMmaWarp) pdl_wait();     // the MMA warp first starts the (constant) table copy, then waits too

  if (warp >= kEpiWarp0 && warp < kMmaWarp) {
    // ---------------------------------------------------------------- epilogue: 8 independent warps
    // TMEM lane = inner element, column = output index, and inner is the contiguous dimension of Y: a warp's 32 lanes
    // of one column are 128 contiguous bytes in global memory, so every warp drains its own lane quadrant straight
    // from registers with one STG.32 per column (one full line per instruction, no staging, no block barrier).
    //   npad <= 32 (one chunk per tile): warps 0-3 / 4-7 take alternate tiles, each drains a stage alone;
    //   npad  > 32: both sets work on every tile, set t takes chunks t, t+2, ... and both release the stage.
    const int team = (warp - kEpiWarp0) >> 2, quad = warp & 3;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int total_chunks = (p.npad + 31) >> 5;
    const bool split_tiles = total_chunks == 1;
    const int gq = quad >> 1;                                  // which of the tile's two 64-element groups
    const int in_group = (quad & 1) * 32 + lane;
    const unsigned stride = (unsigned)p.inner;
    int tile = (int)blockIdx.x - (int)gridDim.x;
    for (int n = 0; n < my_tiles; ++n) {
      tile += gridDim.x;
      if (kPipe && tile >= n_tiles) tile -= n_tiles;       // next layer
      const int ds = n & 1;
      if (split_tiles && ds != team) continue;
      // Where the tile goes, BEFORE waiting for its accumulator: the division and the 64-bit multiplies are a
      // dependent chain of several hundred cycles (in-kernel timeline: 750-950 between the accumulator becoming ready
      // and the tcgen05.ld when computed after the wait) that belongs under the wait, not behind it.
      const int ptile = (!kPipe && set.reverse) ? n_tiles - 1 - tile : tile;
      const long long G = (long long)ptile * 2 + gq;             // warp-uniform
      const bool live = (FFNO_KO & 2) ? false : G < n_groups;
      const unsigned uo = live ? (unsigned)G / (unsigned)gpi : 0u;
      const unsigned ug = live ? (unsigned)G - uo * (unsigned)gpi : 0u;
      float* ybase = p.Y + ((long long)uo * p.n_out) * p.inner + (long long)ug * 64 + in_group;
      const uint32_t taddr = tmem + lane_base + (uint32_t)(ds * stage_cols);
      // Pin both addresses in registers HERE: without the empty asm ptxas sinks the whole chain (and the constant-bank
      // reloads of the geometry) behind the wait, right in front of the tcgen05.ld (SASS: ~55 instructions and an LDCU
      // between the barrier test and LDTM; in-kernel timeline: 700-950 cycles with the accumulator stage held, 210 now).
      asm volatile("" ::"l"(ybase), "r"(taddr), "r"(stride) : "memory");
      mbar_wait(&d_full[ds], (uint32_t)(n >> 1) & 1u);
      tc_fence_after();
      if (warp == kEpiWarp0 && blockIdx.y == 0) TL(5, n, 0);
      const int ch_begin = split_tiles ? 0 : team, ch_step = split_tiles ? 1 : 2;
      bool released = false;
#pragma unroll 1
      for (int ch = ch_begin; ch < total_chunks; ch += ch_step) {
        const int c0 = ch * 32;
        uint32_t v[32];
        if (warp == kEpiWarp0 && blockIdx.y == 0) TL(5, n, 3);
        tmem_ld32(taddr + (uint32_t)c0, v);
        if (warp == kEpiWarp0 && blockIdx.y == 0) TL(5, n, 4);
        tmem_ld_wait();
        if (warp == kEpiWarp0 && blockIdx.y == 0) TL(5, n, 5);
        if (ch + ch_step >= total_chunks) {   // this warp's last read of the stage
          tc_fence_before();
          mbar_arrive(&d_empty[ds]);
          released = true;
          if (warp == kEpiWarp0 && blockIdx.y == 0) TL(5, n, 1);
        }
        if (live) {
          float* yp = ybase + (size_t)c0 * stride;
          const int ncols = p.n_out - c0;                        // >= 32: full chunk
          if (!p.accumulate && ncols >= 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) yp[(size_t)j * stride] = __uint_as_float(v[j]);
          } else if (kRmw && p.accumulate && ncols >= 32) {
            // read-modify-write (summed spectral output of the taps / standalone / backward paths): several loads in
            // flight before the first add — a load -> add -> store chain per column costs one memory latency each
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += 8) {      // 8 at a time: the register budget of the kernel (72) allows no more
              float old[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) old[j] = yp[(size_t)(j0 + j) * stride];
#pragma unroll
              for (int j = 0; j < 8; ++j) yp[(size_t)(j0 + j) * stride] = old[j] + __uint_as_float(v[j0 + j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncols) {
                float* dst = yp + (size_t)j * stride;
                *dst = p.accumulate ? *dst + __uint_as_float(v[j]) : __uint_as_float(v[j]);
              }
          }
        }
      }
      if (!split_tiles && !released) {        // a warp with no chunk in this tile still owes its arrivals
        tc_fence_before();
        mbar_arrive(&d_empty[ds]);
      }
      if (kPipe) stores_issued(&s_pub[team]);
      if (warp == kEpiWarp0 && blockIdx.y == 0) TL(5, n, 2);
    }
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- MMA issuer
    {   // the whole warp walks the schedule (warp-uniform); elect.sync picks the issuing lane per instruction
      pdl_wait();
      mbar_wait(bar_w, 0);
      const uint32_t idesc = make_idesc_bf16(128, p.npad, 1, 0);
      int item = 0;
      if (p.kchunks == 1 && !(FFNO_KO & 8)) {
        // One operand stage per tile (axis lengths <= 64: C2, C5, every inverse with <= 32 modes): operand stage and
        // accumulator stage are both n & 1, so with the tile loop unrolled by two every descriptor is a loop invariant.
        // The generic loop below rebuilds them per tile — ~85 SASS instructions (shifts, masks, R2UR, two constant-bank
        // reloads) between the barrier test and the first tcgen05.mma; in-kernel timeline: 1000-1100 -> 650-800 cycles of
        // issue per tile, forward-transform launch 17.7 -> 17.2 us.
        const int ksteps = (p.n_in + 15) >> 4;
        const uint64_t dA0h = make_smem_desc_sw128(smem_u32(smem), 8192u, 1024u);
        const uint64_t dBh = desc_kmajor(smem_u32(sB), 0);
        const uint64_t dBl = dBh + (uint64_t)(b_half >> 4);
        auto issue_tile = [&](auto STAGE, int n) {
          constexpr int st = decltype(STAGE)::value;
          const uint32_t ph = (uint32_t)(n >> 1) & 1u;
          mbar_wait(&d_empty[st], ph ^ 1u);
          mbar_wait(&a_full[st], ph);
          tc_fence_after();
          if (lane == 0 && blockIdx.y == 0) TL(6, n, 0);
          const uint64_t dAh = dA0h + (uint64_t)(st * (AXP_A_STAGE >> 4)), dAl = dAh + (uint64_t)(16384 >> 4);
          const uint32_t d_addr = tmem + (uint32_t)(st * stage_cols);
          auto issue = [&](auto KS) {
            constexpr int kSteps = decltype(KS)::value;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
              const uint64_t a = (pass == 2) ? dAl : dAh, b = (pass == 1) ? dBl : dBh;
#pragma unroll
              for (int ks = 0; ks < kSteps; ++ks)
                umma_bf16_ss_elect(d_addr, a + (uint64_t)(ks * (2048 >> 4)), b + (uint64_t)(ks * 2), idesc,
                                   (pass == 0 && ks == 0) ? 0u : 1u);
            }
          };
          switch (ksteps) {       // warp-uniform
            case 4: issue(std::integral_constant<int, 4>{}); break;
            case 3: issue(std::integral_constant<int, 3>{}); break;
            case 2: issue(std::integral_constant<int, 2>{}); break;
            default: issue(std::integral_constant<int, 1>{}); break;
          }
          umma_commit_elect(&a_empty[st]);
          umma_commit_elect(&d_full[st]);
          if (lane == 0 && blockIdx.y == 0) TL(6, n, 1);
        };
        for (int n = 0; n < my_tiles; n += 2) {
          issue_tile(std::integral_constant<int, 0>{}, n);
          if (n + 1 < my_tiles) issue_tile(std::integral_constant<int, 1>{}, n + 1);
        }
      } else
      for (int n = 0; n < my_tiles; ++n) {
        const int ds = n & 1;
        mbar_wait(&d_empty[ds], ((uint32_t)(n >> 1) & 1u) ^ 1u);
        const uint32_t d_addr = tmem + (uint32_t)(ds * stage_cols);
        for (int kc = 0; kc < p.kchunks; ++kc, ++item) {
          const int as = item & 1;
          mbar_wait(&a_full[as], (uint32_t)(item >> 1) & 1u);
          tc_fence_after();
          if (lane == 0 && blockIdx.y == 0) TL(6, n, 0);
          const int rem = p.n_in - kc * 64;
          const int ksteps = rem >= 64 ? 4 : (rem + 15) / 16;
          // MN-major A: 64-wide mn blocks 8 KB apart (LBO), 8-row k groups 1 KB apart (SBO); one K step = 2 groups
          const uint64_t dAh = make_smem_desc_sw128(smem_u32(smem + as * AXP_A_STAGE), 8192u, 1024u);
          const uint64_t dAl = dAh + (uint64_t)(16384 >> 4);
          const uint64_t dBh = desc_kmajor(smem_u32(sB) + (uint32_t)kc * ((uint32_t)p.npad * 128u), 0);
          const uint64_t dBl = dBh + (uint64_t)(b_half >> 4);
          const uint32_t acc0 = kc > 0 ? 1u : 0u;
          auto issue = [&](auto KS) {
            constexpr int kSteps = decltype(KS)::value;
#pragma unroll
            for (int pass = 0; pass < KO_PASSES(8); ++pass) {
              const uint64_t a = (pass == 2) ? dAl : dAh, b = (pass == 1) ? dBl : dBh;
#pragma unroll
              for (int ks = 0; ks < kSteps; ++ks)
                umma_bf16_ss_elect(d_addr, a + (uint64_t)(ks * (2048 >> 4)), b + (uint64_t)(ks * 2), idesc,
                                   (pass == 0 && ks == 0) ? acc0 : 1u);
            }
          };
          switch (ksteps) {       // warp-uniform
            case 4: issue(std::integral_constant<int, 4>{}); break;
            case 3: issue(std::integral_constant<int, 3>{}); break;
            case 2: issue(std::integral_constant<int, 2>{}); break;
            default: issue(std::integral_constant<int, 1>{}); break;
          }
          umma_commit_elect(&a_empty[as]);
        }
        umma_commit_elect(&d_full[ds]);
        if (lane == 0 && blockIdx.y == 0) TL(6, n, 1);
      }
    }
    __syncwarp();
  } else if (warp == kAxProdWarp) {
    // ---------------------------------------------------------------- TMA producer (+ kPipe: dependency waits)
    // One thread walks the CTA's tile sequence `ring` items ahead of the converters: it waits until every converter
    // has taken its part of a staging slot, (kPipe) until the tile's unit is complete in the producer stage, then
    // fetches the item's two 64 x 64 groups with one tensor-map load each.  Rows beyond n_in and groups beyond the
    // tensor read as zeros.
    if (lane == 0) {
      const CUtensorMap* tm = &set.tm[blockIdx.y];
      tma_prefetch_desc(tm);
      if (kPipe) tma_prefetch_desc(&set.tm_odd[blockIdx.y]);
      pdl_wait();
      const unsigned* ctr = set.pipe.wait_ctr[blockIdx.y];
      const unsigned cnt = set.pipe.wait_count[blockIdx.y];
      const int tpu = kPipe ? set.pipe.tiles_per_unit[blockIdx.y] : 1;
      long long dbg_blocked = 0;
      const long long dbg_c0 = kPipe ? clock64() : 0;
      const unsigned long long dbg_t0 = kPipe ? globaltimer_ns() : 0ull;
      int tile = (int)blockIdx.x - (int)gridDim.x, layer = 0, last = -1, j = 0;
      for (int n = 0; n < my_tiles; ++n) {
        tile += gridDim.x;
        if (kPipe && tile >= n_tiles) {
          tile -= n_tiles;
          ++layer;
          tm = (layer & 1) ? &set.tm_odd[blockIdx.y] : &set.tm[blockIdx.y];
        }
        if (kPipe) {
          const int unit = tile / tpu, seq = layer * (n_tiles + 1) + unit;
          if (seq != last) {
            const long long ta = clock64();
            pipe_wait(ctr + unit * kCtrStride, (unsigned)(layer + set.pipe.wait_lag) * cnt);
            asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy writes of the producer stage -> TMA reads
            dbg_blocked += clock64() - ta;
            last = seq;
            if (set.pipe.dbg_ts && blockIdx.x == 0 && blockIdx.y == 0) {
              const int q = layer * (n_tiles / tpu) + unit;
              if (q < 64) set.pipe.dbg_ts[q] = globaltimer_ns();
            }
          }
        }
        const unsigned G0 = 2u * (unsigned)((!kPipe && set.reverse) ? n_tiles - 1 - tile : tile), G1 = G0 + 1u;
        const unsigned o0 = G0 / (unsigned)gpi, u0 = G0 - o0 * (unsigned)gpi;
        const unsigned o1 = G1 / (unsigned)gpi, u1 = G1 - o1 * (unsigned)gpi;     // o1 == outer for a dead group: zeros
        for (int kc = 0; kc < p.kchunks; ++kc, ++j) {
          const int slot = j & (ring - 1), use = j >> ring_log2;
          if (use > 0) mbar_wait(&stg_empty[slot], (uint32_t)(use - 1) & 1u);
          uint8_t* dst = smem + AXP_STAGING + slot * 32768;
          mbar_expect_tx(&stg_full[slot], 32768u);
          tma_load_4d(dst, tm, 0, (int)u0, kc * 64, (int)o0, &stg_full[slot]);
          tma_load_4d(dst + 16384, tm, 0, (int)u1, kc * 64, (int)o1, &stg_full[slot]);
        }
      }
      if (kPipe && set.pipe.dbg) {
        unsigned long long* d = set.pipe.dbg + 4 * (blockIdx.y * gridDim.x + blockIdx.x);
        d[0] = (unsigned long long)dbg_blocked;
        d[1] = (unsigned long long)(clock64() - dbg_c0);
        d[2] = dbg_t0;
        d[3] = globaltimer_ns();
      }
    }
  } else if (kPipe && warp == kAxPubWarp) {
    // ---------------------------------------------------------------- publisher (see pipe_publish)
    if (lane == 0) {
      const bool split_tiles = p.npad <= 32;       // one team stores a tile (teams alternate), else both do
      unsigned* ctr = set.pipe.done_ctr[blockIdx.y];
      const int tpu = set.pipe.tiles_per_unit[blockIdx.y];
      int tile = (int)blockIdx.x - (int)gridDim.x, need0 = 0, need1 = 0;
      for (int n = 0; n < my_tiles; ++n) {
        tile += gridDim.x;
        if (tile >= n_tiles) tile -= n_tiles;
        if (!split_tiles || (n & 1) == 0) { need0 += 4; smem_wait_ge(&s_pub[0], need0); }
        if (!split_tiles || (n & 1) == 1) { need1 += 4; smem_wait_ge(&s_pub[1], need1); }
        pipe_publish(ctr + (tile / tpu) * kCtrStride);
      }
    }
  } else if (warp > kMmaWarp && warp < kAxProdWarp) {
    // ---------------------------------------------------------------- converters
    // A staging slot holds one item as the TMA delivers it: [group 2][axis index 64][64 floats].  Every thread takes
    // four float4 of it into registers, releases the slot, and — once the MMA warp has retired the operand stage —
    // writes them as BF16 hi / lo into the MN-major SWIZZLE_128B operand tiles.
    const int lt = tid - kLoaderThread0;
    const int n_items = my_tiles * p.kchunks;
    const int gsel = (lt >> 4) & 1, c4 = lt & 15, rsub = lt >> 5;            // rsub: 0..15
    const uint8_t* stg_thr = smem + AXP_STAGING + gsel * 16384 + rsub * 256 + c4 * 16;     // + it * 4096: rows rsub + 16 it
    for (int item = 0; item < n_items; ++item) {
      const int slot = item & (ring - 1);
      mbar_wait(&stg_full[slot], (uint32_t)(item >> ring_log2) & 1u);
      if (lt < 32 && blockIdx.y == 0) TL(7, item, 0);
      const uint8_t* src = stg_thr + slot * 32768;
      float4 v[kLdPerThread];
#pragma unroll
      for (int it = 0; it < kLdPerThread; ++it) v[it] = (FFNO_KO & 4) ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(src + it * 4096);
      const int as = item & 1;
      mbar_wait(&a_empty[as], ((uint32_t)(item >> 1) & 1u) ^ 1u);
      if (lt < 32 && blockIdx.y == 0) TL(7, item, 1);
      uint8_t* sAh = smem + as * AXP_A_STAGE;
      uint8_t* sAl = sAh + 16384;
#pragma unroll
      for (int it = 0; it < kLdPerThread; ++it) {
        const int il = it * 16 + rsub;
        const uint32_t off = (uint32_t)gsel * 8192u + (uint32_t)(il >> 3) * 1024u + (uint32_t)(il & 7) * 128u +
                             (uint32_t)(((c4 >> 1) ^ (il & 7)) << 4) + (uint32_t)(c4 & 1) * 8u;
        if (!(FFNO_KO & 4)) store_split4_at(sAh, sAl, off, v[it]);
      }
      // The slot is released only now: mbarrier operations are not ordered behind shared-memory loads still in flight,
      // so the arrival must come after instructions that consumed the loaded registers (measured: releasing right after
      // the LDS lets the next TMA overwrite data that has not been read yet).
      mbar_arrive(&stg_empty[slot]);
      fence_proxy_async_smem();
      mbar_arrive(&a_full[as]);
      if (lt < 32 && blockIdx.y == 0) TL(7, item, 2);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

// ---- TMA descriptors (host) ------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}
// Input of an axis transform, X[outer][n_in][inner] FP32, as the 4-D tensor [outer][n_in][inner / 64][64] with a box
// of one group: 64 axis indices x 64 contiguous floats (16 KB).  Coordinates past n_in / outer read as zeros.
static int encode_axis_tmap(CUtensorMap* tm, const float* X, long long outer, int n_in, long long inner) {
  EncodeTiledFn fn = encode_tiled_fn();
  FFNO_REQUIRE(fn != nullptr, FFNO_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  FFNO_REQUIRE(((uintptr_t)X & 15) == 0, FFNO_ERR_BAD_ARG, "axis transform input must be 16-byte aligned");
  const cuuint64_t dims[4] = {64, (cuuint64_t)(inner / 64), (cuuint64_t)n_in, (cuuint64_t)outer};
  const cuuint64_t strides[3] = {256, (cuuint64_t)inner * 4, (cuuint64_t)n_in * (cuuint64_t)inner * 4};
  const cuuint32_t box[4] = {64, 1, 64, 1}, estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(X), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  FFNO_REQUIRE(r == CUDA_SUCCESS, FFNO_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%lld][%d][%lld]", (int)r, outer,
               n_in, inner);
  return FFNO_OK;
}

static int axis_ring_depth(int n_in, int n_out) {      // deepest staging ring that leaves room for the table, 0 = none
  static const int max_ring = [] {
    const char* e = getenv("FFNO_B200_AXIS_RING");       // A/B switch: 1, 2 or 4
    const int v = e ? atoi(e) : kAxStages;
    return (v == 1 || v == 2 || v == 4) ? v : kAxStages;
  }();
  for (int ring = max_ring; ring >= 1; ring >>= 1)
    if (axp_table_offset(ring) + table_image_bytes(n_in, n_out) <= (size_t)227 * 1024) return ring;
  return 0;
}
bool axis_pipe_fits(int n_in, int n_out) { return n_out <= 256 && axis_ring_depth(n_in, n_out) > 0; }

int launch_axis_pipe(const AxisXform* axes, int n_axes, int sm_count, cudaStream_t st, bool reverse) {
  FFNO_REQUIRE(n_axes >= 1 && n_axes <= 3, FFNO_ERR_BAD_ARG, "axis_pipe: n_axes=%d", n_axes);
  AxisSet set{};
  set.reverse = reverse ? 1 : 0;
  size_t smem = 0;
  int max_tiles = 0;
  for (int a = 0; a < n_axes; ++a) {
    const AxisXform& p = axes[a];
    FFNO_REQUIRE(p.inner % 64 == 0 && p.inner < (1ll << 23), FFNO_ERR_UNSUPPORTED,
                 "axis_pipe: inner=%lld must be a multiple of 64 below 2^23", p.inner);
    FFNO_REQUIRE(p.npad >= 16 && p.npad <= 256 && p.npad % 16 == 0, FFNO_ERR_UNSUPPORTED, "axis_pipe: npad=%d", p.npad);
    const int ring = axis_ring_depth(p.n_in, p.n_out);
    FFNO_REQUIRE(ring > 0, FFNO_ERR_UNSUPPORTED, "axis_pipe: table does not fit in shared memory");
    set.ring[a] = ring;
    const size_t need = axp_table_offset(ring) + table_image_bytes(p.n_in, p.n_out);
    FFNO_REQUIRE(p.outer * (p.inner / 64) < (1ll << 31), FFNO_ERR_UNSUPPORTED, "axis_pipe: too many groups");
    smem = need > smem ? need : smem;
    set.ax[a] = p;
    const long long n_groups = p.outer * (p.inner / 64);
    set.n_tiles[a] = (int)((n_groups + 1) / 2);
    int cols = 32;
    while (cols < 2 * p.npad) cols *= 2;
    set.tmem_cols[a] = cols;
    max_tiles = set.n_tiles[a] > max_tiles ? set.n_tiles[a] : max_tiles;
    if (n_groups > 0) FFNO_TRY(encode_axis_tmap(&set.tm[a], p.X, p.outer, p.n_in, p.inner));
  }
  if (max_tiles == 0) return FFNO_OK;
  bool rmw = false;
  for (int a = 0; a < n_axes; ++a) rmw |= axes[a].accumulate != 0;
  if (rmw) {
    FFNO_TRY(ensure_dynamic_smem(axis_pipe_kernel<false, true>, smem));
  } else {
    FFNO_TRY(ensure_dynamic_smem(axis_pipe_kernel<false>, smem));
  }
  int per_axis = sm_count / n_axes;
  if (per_axis < 1) per_axis = 1;
  const int gx = max_tiles < per_axis ? max_tiles : per_axis;
  set.pipe = PipeDesc{};
  if (rmw) FFNO_CUDA_CHECK(launch_pdl(axis_pipe_kernel<false, true>, dim3(gx, n_axes), dim3(kAxThreads), smem, st, set));
  else FFNO_CUDA_CHECK(launch_pdl(axis_pipe_kernel<false>, dim3(gx, n_axes), dim3(kAxThreads), smem, st, set));
  ++g_launch_counter;
  return FFNO_OK;
}

// =======================================================================================================
// Per-mode complex channel mix, all axes in one launch (blockIdx.z = axis, blockIdx.y = mode)
//   work item = (row tile, K block): K block 0 = real parts, 1 = imaginary parts of the 128-wide mix row
// =======================================================================================================
constexpr int MXP_A_STAGE = 32768;
constexpr int MXP_B = 2 * MXP_A_STAGE;             // 65536: B image 64 KB
constexpr int MXP_BAR = MXP_B + 65536;             // 131072
constexpr int MXP_STAGING = MXP_BAR + 1024;        // kMxStages x 32 KB raw FP32 (cp.async landing zone)
constexpr int kMxStages = 2;
constexpr int MXP_OUT = MXP_STAGING + kMxStages * 32768;       // 32 KB FP32 output staging (one re / im segment)
constexpr int MXP_TOTAL = MXP_OUT + 32768;                     // 230400 <= 232448

struct MixSet {
  MixAxis ax[3];
  int tiles_per_cta[3];
  int reverse;
  PipeDesc pipe;
};

// kPipe: the CTA of (axis, mode) walks its row tiles of every layer back to back with its weight image stationary
// (the mode weights must be shared by the layers), waits for the unit's forward transforms and posts completion.
template <bool kPipe>
__global__ void __launch_bounds__(kPipe ? kPipeThreads : kThreads, 1) mix_pipe_kernel(MixSet set) {
  const MixAxis& ax = set.ax[blockIdx.z];
  const int k = blockIdx.y;
  if (k >= ax.K) return;
  const long long M = ax.outer * ax.p_inner;
  const int n_tiles = (int)((M + 127) / 128);
  const int tpc = set.tiles_per_cta[blockIdx.z];
  const int tile_begin = blockIdx.x * tpc;
  if (tile_begin >= n_tiles) return;
  const int tile_end = min(n_tiles, tile_begin + tpc);
  const int n_layers = kPipe ? set.pipe.n_layers : 1;
  const int my_tiles = KO_TILES((tile_end - tile_begin) * n_layers);
  [[maybe_unused]] const bool tl_on = g_timeline_on == 4;
  [[maybe_unused]] const bool tl_cta = blockIdx.y == 0 && blockIdx.z == 0;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem + MXP_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MXP_BAR);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + 2;
  uint64_t* d_full = bars + 4;
  uint64_t* d_empty = bars + 6;
  uint64_t* bar_w = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  int* s_seq = reinterpret_cast<int*>(bars + 10);
  int* s_pub = s_seq + 1;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    *s_seq = -1;
    s_pub[0] = s_pub[1] = 0;
    mbar_init(bar_w, 1);                       // weight image first: see axis_pipe_kernel
    fence_barrier_init();
    {
      mbar_expect_tx(bar_w, kMixImageBytes);
      const uint8_t* img = ax.image + (long long)k * kMixImageBytes;
      bulk_g2s(sB, img, 32768, bar_w);
      bulk_g2s(sB + 32768, img + 32768, 32768, bar_w);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], kLoaders);
      mbar_init(&a_empty[i], 1);
      mbar_init(&d_full[i], 1);
      mbar_init(&d_empty[i], kEpiWarps * 32);
    }
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const long long inner = ax.p_inner * 64;
  pdl_launch_dependents();
  if (warp != kMmaWarp) pdl_wait();

  if (warp >= kEpiWarp0 && warp < kMmaWarp) {
    // two epilogue teams: team 0 drains the real segment (columns 0..63), team 1 the imaginary one, each 32 columns
    // at a time through its own swizzled 16 KB staging tile so the global stores are full 128-byte lines
    const int team = (warp - kEpiWarp0) >> 2, rt = tid & 127, seg = team;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint8_t* sOut = smem + MXP_OUT + team * 16384;
    const int rq = rt >> 3, cq = rt & 7;                    // coalesced phase: rows rq + 16 it, float4 column cq
    const uint8_t* so_rd = sOut + rq * 128 + ((cq ^ (rq & 7)) << 4);        // (rq + 16 it) & 7 == rq & 7
    const unsigned p_in = (unsigned)ax.p_inner;
    const unsigned q16 = 16u / p_in, r16 = 16u % p_in;
    const long long o_stride = (long long)ax.K * 2 * inner;
    float* out_seg = ax.R + ((long long)k * 2 + seg) * inner + cq * 4;
    int tile = tile_begin - 1;
    for (int n = 0; n < my_tiles; ++n) {
      if (++tile == tile_end) tile = tile_begin;         // kPipe: next layer
      const int ds = n & 1;
      if (tl_cta && (warp & 3) == 0) TL(team ? 6 : 5, n, 0);
      mbar_wait(&d_full[ds], (uint32_t)(n >> 1) & 1u);
      tc_fence_after();
      if (tl_cta && (warp & 3) == 0) TL(team ? 6 : 5, n, 1);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + (uint32_t)(ds * 128 + seg * 64 + half * 32), v);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(&d_empty[ds]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
          *reinterpret_cast<float4*>(sOut + rt * 128 + ((e ^ (rt & 7)) << 4)) =
              make_float4(__uint_as_float(v[e * 4]), __uint_as_float(v[e * 4 + 1]), __uint_as_float(v[e * 4 + 2]),
                          __uint_as_float(v[e * 4 + 3]));
        if (team == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        {
          // rows rq + 16 it: (outer, position) of the first by one division, the rest by a constant step
          // (kept inside the drain: hoisting the eight row pointers above the accumulator wait costs 16 registers the
          // kernel does not have — it spilled and the forward got slower, 1.90 -> 1.97 ms)
          const long long row_first = (long long)((!kPipe && set.reverse) ? n_tiles - 1 - tile : tile) * 128 + rq;
          const unsigned rf = row_first < M ? (unsigned)row_first : 0u;
          unsigned uo = rf / p_in, pp = rf - uo * p_in;
          const int rows_left = (FFNO_KO & 128) ? 0 : ((M - row_first) > 0 ? (int)((M - row_first) < 128 ? (M - row_first) : 128) : 0);   // rows rq .. M-1
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            if (it * 16 < rows_left) {
              float* dst = out_seg + (long long)uo * o_stride + (long long)pp * 64 + half * 32;
              *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(so_rd + it * 2048);
            }
            uo += q16;
            pp += r16;
            if (pp >= p_in) { pp -= p_in; ++uo; }
          }
        }
        if (team == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
        else asm volatile("bar.sync 2, 128;" ::: "memory");
      }
      if (tl_cta && (warp & 3) == 0) TL(team ? 6 : 5, n, 2);
      if (kPipe) stores_issued(&s_pub[team]);
    }
  } else if (warp == kMmaWarp) {
    {
      pdl_wait();
      mbar_wait(bar_w, 0);
      constexpr uint32_t IDESC = make_idesc_bf16(128, 128, 0, 0);
      const uint64_t dBh = desc_kmajor(smem_u32(sB), 0), dBl = desc_kmajor(smem_u32(sB) + 32768u, 0);
      const uint64_t dAh = desc_kmajor(smem_u32(smem), 0), dAl = desc_kmajor(smem_u32(smem) + 16384u, 0);
      int item = 0;
      for (int n = 0; n < my_tiles; ++n) {
        const int ds = n & 1;
        mbar_wait(&d_empty[ds], ((uint32_t)(n >> 1) & 1u) ^ 1u);
        const uint32_t d_addr = tmem + (uint32_t)(ds * 128);
        for (int kb = 0; kb < 2; ++kb, ++item) {
          const int as = item & 1;
          mbar_wait(&a_full[as], (uint32_t)(item >> 1) & 1u);
          tc_fence_after();
          if (tl_cta) TL(4, n, kb);
          const uint64_t a_off = (uint64_t)(as * (MXP_A_STAGE >> 4)), b_off = (uint64_t)(kb * (16384 >> 4));
          issue3_kmajor_elect<4>(d_addr, dAh + a_off, dAl + a_off, dBh + b_off, dBl + b_off, IDESC, kb > 0 ? 1u : 0u, KO_PASSES(256));
          umma_commit_elect(&a_empty[as]);
        }
        umma_commit_elect(&d_full[ds]);
      }
    }
    __syncwarp();
  } else if (kPipe && warp == kPollWarp) {
    if (lane == 0) {      // poller
      const unsigned* ctr = set.pipe.wait_ctr[blockIdx.z];
      const unsigned cnt = set.pipe.wait_count[blockIdx.z];
      const int tpu = set.pipe.tiles_per_unit[blockIdx.z];
      int tile = tile_begin - 1, layer = 0, last = -1;
      for (int n = 0; n < my_tiles; ++n) {
        if (++tile == tile_end) { tile = tile_begin; ++layer; }
        const int unit = tile / tpu, seq = layer * (n_tiles + 1) + unit;
        if (seq != last) {
          pipe_wait(ctr + unit * kCtrStride, (unsigned)(layer + set.pipe.wait_lag) * cnt);
          st_release_cta_smem(s_seq, seq);
          last = seq;
        }
      }
    }
  } else if (kPipe && warp == kPubWarp) {
    if (lane == 0) {      // publisher: both teams store a part of every tile
      unsigned* ctr = set.pipe.done_ctr[blockIdx.z];
      const int tpu = set.pipe.tiles_per_unit[blockIdx.z];
      int tile = tile_begin - 1, need = 0;
      for (int n = 0; n < my_tiles; ++n) {
        if (++tile == tile_end) tile = tile_begin;
        need += 4;
        smem_wait_ge(&s_pub[0], need);
        smem_wait_ge(&s_pub[1], need);
        pipe_publish(ctr + (tile / tpu) * kCtrStride);
      }
    }
  } else {
    // cp.async staging ring, one private slice per thread (see axis_pipe_kernel)
    const int lt = tid - kLoaderThread0;
    const int n_items = my_tiles * 2;
    const int c4 = lt & 15, rsub = lt >> 4;                                   // rsub: 0..31
    int iss_tile = tile_begin, iss_layer = 0, ready_seq = -1;                 // tile / layer of the next issue
    long long dbg_blocked = 0;
    const long long dbg_c0 = kPipe ? clock64() : 0;
    const unsigned long long dbg_t0 = kPipe ? globaltimer_ns() : 0ull;
    const unsigned p_in = (unsigned)ax.p_inner;
    uint8_t* stg_base = smem + MXP_STAGING + lt * 16;
    const unsigned q32 = 32u / p_in, r32 = 32u % p_in;
    const long long o_stride = (long long)ax.K * 2 * inner;
    const float* f_mode = ax.F + (long long)k * 2 * inner + c4 * 4;
    auto issue = [&](int item) {
      if (item < n_items) {
        const int ltile = iss_tile, half = item & 1;
        if (kPipe && half == 0) {         // first half of a tile: the unit's forward transforms must be complete
          const int unit = ltile / set.pipe.tiles_per_unit[blockIdx.z];
          const int seq = iss_layer * (n_tiles + 1) + unit;
          if (seq != ready_seq) {
            dbg_blocked += wait_unit_ready(s_seq, seq);
            ready_seq = seq;
            if (set.pipe.dbg_ts && lt == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
              const int q = iss_layer * (n_tiles / set.pipe.tiles_per_unit[0]) + unit;
              if (q < 64) set.pipe.dbg_ts[q] = globaltimer_ns();
            }
          }
        }
        if (half == 1 && ++iss_tile == tile_end) { iss_tile = tile_begin; ++iss_layer; }
        const int tile = (!kPipe && set.reverse) ? n_tiles - 1 - ltile : ltile;
        uint8_t* dst = stg_base + (item % kMxStages) * 32768;
        const long long row_first = (long long)tile * 128 + rsub;          // rows rsub + 32 it
        const unsigned rf = row_first < M ? (unsigned)row_first : 0u;
        unsigned o = rf / p_in, pp = rf - o * p_in;
        const long long left = M - row_first;
        const float* f_half = f_mode + (long long)half * inner;
#pragma unroll
        for (int it = 0; it < kLdPerThread; ++it) {
          const bool ok = (FFNO_KO & 512) ? false : it * 32 < left;
          const float* src = f_half + (long long)o * o_stride + (long long)pp * 64;
          cp_async16(dst + it * (kLoaders * 16), ok ? (const void*)src : (const void*)ax.F, ok ? 16u : 0u);
          o += q32;
          pp += r32;
          if (pp >= p_in) { pp -= p_in; ++o; }
        }
      }
      cp_async_commit();
    };
    for (int q = 0; q < kMxStages; ++q) issue(q);
    for (int item = 0; item < n_items; ++item) {
      cp_async_wait<kMxStages - 1>();
      if (tl_cta && lt < 32) TL(7, item, 0);
      const uint8_t* src = stg_base + (item % kMxStages) * 32768;
      float4 v[kLdPerThread];
#pragma unroll
      for (int it = 0; it < kLdPerThread; ++it) v[it] = *reinterpret_cast<const float4*>(src + it * (kLoaders * 16));
      const int as = item & 1;
      mbar_wait(&a_empty[as], ((uint32_t)(item >> 1) & 1u) ^ 1u);
      uint8_t* sAh = smem + as * MXP_A_STAGE;
      uint8_t* sAl = sAh + 16384;
#pragma unroll
      for (int it = 0; it < kLdPerThread; ++it)
        store_split4_at(sAh, sAl, kmajor_sw128_offset(it * 32 + rsub, c4 * 4), v[it]);
      fence_proxy_async_smem();
      mbar_arrive(&a_full[as]);
      if (tl_cta && lt < 32) TL(7, item, 1);
      issue(item + kMxStages);
      if (tl_cta && lt < 32) TL(7, item, 2);
    }
    cp_async_wait<0>();
    if (kPipe && lt == 0 && set.pipe.dbg) {
      unsigned long long* d = set.pipe.dbg + 4 * (blockIdx.z * gridDim.y + blockIdx.y);
      d[0] = (unsigned long long)dbg_blocked;
      d[1] = (unsigned long long)(clock64() - dbg_c0);
      d[2] = dbg_t0;
      d[3] = globaltimer_ns();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int launch_mix_pipe(const MixAxis* axes, int n_axes, int sm_count, cudaStream_t st, bool reverse) {
  FFNO_REQUIRE(n_axes >= 1 && n_axes <= 3, FFNO_ERR_BAD_ARG, "mix_pipe: n_axes=%d", n_axes);
  FFNO_TRY(ensure_dynamic_smem(mix_pipe_kernel<false>, MXP_TOTAL));
  MixSet set;
  set.pipe = PipeDesc{};
  set.reverse = reverse ? 1 : 0;
  int maxK = 0, total_modes = 0;
  long long total_tiles = 0;
  for (int a = 0; a < n_axes; ++a) {
    set.ax[a] = axes[a];
    maxK = axes[a].K > maxK ? axes[a].K : maxK;
    total_modes += axes[a].K;
    total_tiles += (long long)axes[a].K * ((axes[a].outer * axes[a].p_inner + 127) / 128);
  }
  if (total_tiles == 0) return FFNO_OK;
  int grid_x = 1;
  for (int a = 0; a < n_axes; ++a) {
    const long long tiles = (axes[a].outer * axes[a].p_inner + 127) / 128;
    long long ctas_per_mode = (long long)sm_count / (total_modes > 0 ? total_modes : 1);
    if (ctas_per_mode < 1) ctas_per_mode = 1;
    long long tpc = (tiles + ctas_per_mode - 1) / ctas_per_mode;
    if (tpc < 1) tpc = 1;
    set.tiles_per_cta[a] = (int)tpc;
    const int gx = (int)((tiles + tpc - 1) / tpc);
    grid_x = gx > grid_x ? gx : grid_x;
  }
  FFNO_CUDA_CHECK(launch_pdl(mix_pipe_kernel<false>, dim3(grid_x, maxK, n_axes), dim3(kThreads), (size_t)MXP_TOTAL, st, set));
  ++g_launch_counter;
  return FFNO_OK;
}

// =======================================================================================================
// FeedForward + residual: the hidden activations never touch shared memory.
//   G1   : D1[128 x 256] = A1 (smem, 2 stages) x W1 (smem)                       (SS, N = 128 halves)
//   epi  : D1 chunk -> +b1, ReLU, BF16 hi/lo -> tcgen05.st into TMEM operand stage A2[team]
//   G2   : D2[128 x 64] += A2 (TMEM) x W2 (smem)                                  (TS: A operand from tensor memory)
//   store: D2 -> +b2 -> swizzled FP32 staging tile in smem -> coalesced (+ residual) global stores
// The loader sums up to three per-axis spectral outputs (s = s_x + s_y (+ s_z), grid_2d.py:94) while it splits, so
// the inverse transforms never read-modify-write a shared buffer.
// TMEM columns: D1 0..255 | D2 stage 256 + 64 t | A2 stage 384 + 64 t (hi 0..31, lo 32..63).
// =======================================================================================================
constexpr int FF3_W = 0;
constexpr int FF3_A1 = 131072;                     // 2 stages x (hi 16 KB | lo 16 KB)
constexpr int FF3_OUT = FF3_A1 + 65536;            // 196608: 128 rows x 256 B staging
constexpr int FF3_BIAS = FF3_OUT + 32768;          // 229376
constexpr int FF3_BAR = FF3_BIAS + 2 * 320 * 4;    // 231936: two bias sets (kPipe: the layer's parity selects the live one)
constexpr int FF3_HEAD = FF3_BAR + 208;            // 64 floats: folded head weights of a 1-output head
constexpr int FF3_TOTAL = FF3_HEAD + 256;          // 232400 <= 232448
constexpr int kFF3Threads = 576;                   // 8 chunk-epilogue + 4 store + G1 issuer + 4 loader + G2 issuer warps
constexpr int kFF3G1Warp = 12, kFF3G2Warp = 17;
constexpr int kFF3PollWarp = 18, kFF3PubWarp = 19, kFF3PipeThreads = kFF3Threads + 64;   // pipelined kernel only

struct FFArgs {
  const float *s0, *s1, *s2, *residual;
  float *x_out, *b_out;
  const uint8_t* image;
  const float *b1, *b2, *head_w, *head_b;
  float* forecast;
  long long P;
  int n_tiles, reverse;
  // kPipe: every layer in one launch.  Layer l reads the residual stream from xbuf[l & 1] and writes xbuf[(l + 1) & 1]
  // (nothing on the last layer, whose store warps apply the fused head instead); weights / biases come from layers[l].
  const FFLayerArgs* layers;
  float* xbuf[2];
  PipeDesc pipe;
};

template <bool kPipe>
__global__ void __launch_bounds__(kPipe ? kFF3PipeThreads : kFF3Threads, 1)  // 18 warps are allocated as 20: 96 registers/thread
ff_ts_kernel(const FFArgs a) {
  const float* __restrict__ s0 = a.s0;
  const float* __restrict__ s1 = a.s1;
  const float* __restrict__ s2 = a.s2;
  const float* __restrict__ residual = a.residual;
  float* __restrict__ x_out = a.x_out;
  float* __restrict__ b_out = a.b_out;
  const uint8_t* __restrict__ image = kPipe ? a.layers[0].image : a.image;
  const float* __restrict__ b1 = kPipe ? a.layers[0].b1 : a.b1;
  const float* __restrict__ b2 = kPipe ? a.layers[0].b2 : a.b2;
  const float* __restrict__ head_w = a.head_w;
  const float* __restrict__ head_b = a.head_b;
  float* __restrict__ forecast = a.forecast;
  const long long P = a.P;
  const int n_tiles = a.n_tiles, reverse = kPipe ? 0 : a.reverse;
  const int n_layers = kPipe ? a.pipe.n_layers : 1;
  const int my_tiles = KO_TILES((int)(((long long)n_layers * n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x));
  [[maybe_unused]] const bool tl_on = g_timeline_on == 1;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sb1 = reinterpret_cast<float*>(smem + FF3_BIAS);
  float* sb2 = sb1 + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FF3_BAR);
  uint64_t* a1_full = bars;         // [2] 128 (loaders)
  uint64_t* a1_empty = bars + 2;    // [2] commit
  uint64_t* d1_full = bars + 4;     // [2] commit
  uint64_t* d1_empty = bars + 6;    // [2] 256 (both epilogue teams)
  uint64_t* a2_full = bars + 8;     // [4] 128 (team t, 32-column part p: stage 2 t + p)
  uint64_t* a2_empty = bars + 12;   // [4] commit
  uint64_t* d2_full = bars + 16;    // [2] commit
  uint64_t* d2_empty = bars + 18;   // [2] 128 (store warps)
  uint64_t* bar_w = bars + 20;      // W1 image (first 64 KB) landed
  uint64_t* bar_w2 = bars + 21;     // W2 image (second 64 KB) landed: only GEMM2 needs it
  uint64_t* w_free = bars + 22;     // [2] kPipe: every GEMM1 / GEMM2 MMA issued so far has retired (W1 / W2 may be replaced)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  int* s_seq = reinterpret_cast<int*>(tmem_slot + 1);      // kPipe: last unit the poller has seen ready
  int* s_pub = s_seq + 1;                                  // kPipe: store warps that have issued their tile's stores

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(bar_w, 1);                       // weight images first: see axis_pipe_kernel
    mbar_init(bar_w2, 1);
    fence_barrier_init();
    mbar_expect_tx(bar_w, 65536);
    for (int i = 0; i < 2; ++i) bulk_g2s(smem + FF3_W + i * 32768, image + i * 32768, 32768, bar_w);
    mbar_expect_tx(bar_w2, 65536);
    for (int i = 2; i < 4; ++i) bulk_g2s(smem + FF3_W + i * 32768, image + i * 32768, 32768, bar_w2);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a1_full[i], 128);
      mbar_init(&a1_empty[i], 1);
      mbar_init(&d1_full[i], 1);
      mbar_init(&d1_empty[i], 256);
      mbar_init(&d2_full[i], 1);
      mbar_init(&d2_empty[i], 128);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&a2_full[i], 128);
      mbar_init(&a2_empty[i], 1);
    }
    mbar_init(&w_free[0], 1);
    mbar_init(&w_free[1], 1);
    *s_seq = -1;
    *s_pub = 0;
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = tid; i < 256; i += kFF3Threads) sb1[i] = b1 ? b1[i] : 0.f;
  if (tid < 64) sb2[tid] = b2 ? b2[tid] : 0.f;
  float* sHead = reinterpret_cast<float*>(smem + FF3_HEAD);
  if (tid < 64) sHead[tid] = forecast ? head_w[tid] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_launch_dependents();
  if (warp != kFF3G1Warp && warp != kFF3G2Warp) pdl_wait();

  if (warp < 8) {
    // ---------------------------------------------------------------- chunk epilogue teams (thread = row)
    const int team = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t a2_addr = tmem + lane_base + (uint32_t)(384 + team * 64);
    int tile = (int)blockIdx.x - (int)gridDim.x, layer = 0;
    for (int n = 0; n < my_tiles; ++n) {
      tile += gridDim.x;
      if (kPipe && tile >= n_tiles) { tile -= n_tiles; ++layer; }
      if (kPipe && tile < (int)gridDim.x && layer > 0) {
        // first tile of a new layer: its b1 goes into the bias set of the layer's parity (the other set may still be
        // read by a slower warp of the role; the gpu-scope acquires of the loaders keep invalidating L1, so the hot
        // loop must not read biases from global memory)
        sb1[(layer & 1) * 320 + tid] = __ldg(a.layers[layer].b1 + tid);
        asm volatile("bar.sync 3, 256;" ::: "memory");
      }
      const float* bl = sb1 + (kPipe ? (layer & 1) * 320 : 0);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int j = 2 * h + team;              // chunk of hidden units j*64 .. j*64+63
        const int q = 2 * n + h;                 // running chunk index of this team (its A2 stage is `team`)
        mbar_wait(&d1_full[h], (uint32_t)n & 1u);
        tc_fence_after();
        if ((warp & 3) == 0) TL(team ? 5 : 0, n, h * 4 + 0);
        const float* bj = bl + j * 64;
#pragma unroll
        for (int part = 0; part < 2; ++part) {      // 32 columns = one A2 stage at a time (~64 live registers)
          uint32_t v[32];
          tmem_ld32(tmem + lane_base + (uint32_t)(h * 128 + team * 64 + part * 32), v);
          tmem_ld_wait();
          if (part == 1) {                          // D1 chunk fully read: release this half to the MMA warp
            tc_fence_before();
            mbar_arrive(&d1_empty[h]);
            if ((warp & 3) == 0) TL(team ? 5 : 0, n, h * 4 + 1);
          }
          uint32_t hi[16], lo[16];
#pragma unroll
          for (int e = 0; e < 8; ++e) {             // +b1, ReLU, BF16 hi/lo: 6 instructions per element pair
            const float4 bb = *reinterpret_cast<const float4*>(bj + part * 32 + 4 * e);
            split2_relu(fadd2(make_float2(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1])), make_float2(bb.x, bb.y)),
                        hi[2 * e], lo[2 * e]);
            split2_relu(fadd2(make_float2(__uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3])), make_float2(bb.z, bb.w)),
                        hi[2 * e + 1], lo[2 * e + 1]);
          }
          // each 32-column part is its own operand stage (K = 32 of GEMM2): the team only waits for the MMAs that
          // read the same part of its PREVIOUS chunk, issued a whole part earlier than it needs the slot back
          const int stage = team * 2 + part;
          mbar_wait(&a2_empty[stage], ((uint32_t)q & 1u) ^ 1u);
          tc_fence_after();
          if ((warp & 3) == 0 && part == 0) TL(team ? 5 : 0, n, h * 4 + 2);
          tmem_st16(a2_addr + (uint32_t)(part * 16), hi);
          tmem_st16(a2_addr + 32u + (uint32_t)(part * 16), lo);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&a2_full[stage]);
        }
        if ((warp & 3) == 0) TL(team ? 5 : 0, n, h * 4 + 3);
      }
    }
  } else if (warp < 12) {
    // ---------------------------------------------------------------- store warps
    // Every address below is (tile base) + (per-thread constant) + (compile-time step): the role is one serial
    // instruction stream per warp, so its speed is its instruction count.
    const int rt = tid - 256;
    const int rq = rt >> 3, cq = rt & 7;       // coalesced phase: rows rq + 16 it, float4 column cq (+ 8 for half 1)
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint8_t* sOut = smem + FF3_OUT;
    uint8_t* so_wr = sOut + rt * 256;          // accumulator phase: thread = row rt, chunk c4 at ((c4 ^ (rt & 15)) << 4)
    const uint32_t wsw = (uint32_t)(rt & 15) << 4;
    const uint8_t* so_rd0 = sOut + rq * 256 + ((cq ^ rq) << 4);            // (rq + 16 it) & 15 == rq
    const uint8_t* so_rd1 = sOut + rq * 256 + (((8 + cq) ^ rq) << 4);
    const float hb = forecast ? head_b[0] : 0.f;
    int tile = (int)blockIdx.x - (int)gridDim.x, layer = 0, ready_seq = -1;
    const float* b2l = sb2;
    for (int n = 0; n < my_tiles; ++n) {
      tile += gridDim.x;
      if (kPipe && tile >= n_tiles) { tile -= n_tiles; ++layer; }
      if (kPipe) {
        // this layer's buffers; the fused head replaces the residual-stream store on the last layer
        const bool last = layer == n_layers - 1;
        residual = a.xbuf[layer & 1];
        x_out = last ? nullptr : a.xbuf[(layer + 1) & 1];
        forecast = last ? a.forecast : nullptr;
        if (tile < (int)gridDim.x && layer > 0) {       // first tile of a new layer: b2 into the parity's bias set
          if (rt < 64) sb2[(layer & 1) * 320 + rt] = __ldg(a.layers[layer].b2 + rt);
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
        b2l = sb2 + (layer & 1) * 320;
        // the rows read (residual) and overwritten (x of layer l - 1) below belong to FF(l - 1, unit): it must be complete
        const int unit = tile / a.pipe.tiles_per_unit[0];
        const int seq = layer * (n_tiles + 1) + unit;
        if (seq != ready_seq) {      // (ready inverse transforms of (l, unit) imply a complete FF(l - 1, unit))
          wait_unit_ready(s_seq, seq);
          ready_seq = seq;
        }
      }
      const long long row0 = (long long)(reverse ? n_tiles - 1 - tile : tile) * 128;
      const int rows_left = (FFNO_KO & 16) ? 0 : ((P - row0) < 128 ? (int)(P - row0) : 128);
      const long long gofs = row0 * 64 + rq * 64 + cq * 4;
      if (warp == 8) TL(1, n, 0);
      // The tile is finished in two 32-channel halves so that only 8 float4 of the residual are live while the
      // accumulator is in registers: half 0 is fetched before the wait (its rows were pulled into L2 by the loaders'
      // bulk prefetch one tile earlier), half 1 once the accumulator registers are dead.
      float4 r0[8], r1[8];
      const float* rp = residual ? residual + gofs : s0 + gofs;
#pragma unroll
      for (int it = 0; it < 8; ++it)
        r0[it] = (residual && it * 16 + rq < rows_left) ? (kPipe ? ldg_cg(rp + it * 1024) : ldg_stream(rp + it * 1024))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      const int ds = n & 1;
      mbar_wait(&d2_full[ds], (uint32_t)(n >> 1) & 1u);
      tc_fence_after();
      if (warp == 8) TL(1, n, 1);
      float hacc = 0.f;        // fused 1-output head on the last layer: forecast = <b_row, w_eff> + b_eff
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + (uint32_t)(256 + ds * 64 + half * 32), v);
        tmem_ld_wait();
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(&d2_empty[ds]);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c4 = half * 8 + e;
          const float4 bb = *reinterpret_cast<const float4*>(b2l + c4 * 4);
          float4 b;
          b.x = __uint_as_float(v[e * 4 + 0]) + bb.x;
          b.y = __uint_as_float(v[e * 4 + 1]) + bb.y;
          b.z = __uint_as_float(v[e * 4 + 2]) + bb.z;
          b.w = __uint_as_float(v[e * 4 + 3]) + bb.w;
          *reinterpret_cast<float4*>(so_wr + (((uint32_t)c4 << 4) ^ wsw)) = b;      // XOR swizzle: conflict-free
          if (forecast) {
            const float4 hw = *reinterpret_cast<const float4*>(sHead + c4 * 4);
            hacc = fmaf(b.x, hw.x, fmaf(b.y, hw.y, fmaf(b.z, hw.z, fmaf(b.w, hw.w, hacc))));
          }
        }
      }
      if (forecast && rt < rows_left) forecast[row0 + rt] = hacc + hb;
#pragma unroll
      for (int it = 0; it < 8; ++it)
        r1[it] = (residual && it * 16 + rq < rows_left) ? (kPipe ? ldg_cg(rp + it * 1024 + 32) : ldg_stream(rp + it * 1024 + 32))
                                                       : make_float4(0.f, 0.f, 0.f, 0.f);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (warp == 8) TL(1, n, 2);
      if (b_out) {             // last layer only
        float* bo = b_out + gofs;
#pragma unroll
        for (int it = 0; it < 8; ++it)
          if (it * 16 + rq < rows_left) {
            *reinterpret_cast<float4*>(bo + it * 1024) = *reinterpret_cast<const float4*>(so_rd0 + it * 4096);
            *reinterpret_cast<float4*>(bo + it * 1024 + 32) = *reinterpret_cast<const float4*>(so_rd1 + it * 4096);
          }
      }
      if (x_out) {
        float* xo = x_out + gofs;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const float4 b = *reinterpret_cast<const float4*>(so_rd0 + it * 4096);
          if (it * 16 + rq < rows_left)
            *reinterpret_cast<float4*>(xo + it * 1024) =
                make_float4(b.x + r0[it].x, b.y + r0[it].y, b.z + r0[it].z, b.w + r0[it].w);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const float4 b = *reinterpret_cast<const float4*>(so_rd1 + it * 4096);
          if (it * 16 + rq < rows_left)
            *reinterpret_cast<float4*>(xo + it * 1024 + 32) =
                make_float4(b.x + r1[it].x, b.y + r1[it].y, b.z + r1[it].z, b.w + r1[it].w);
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (kPipe) stores_issued(s_pub);
      if (warp == 8) TL(1, n, 3);
    }
  } else if (warp == kFF3G1Warp) {
    // ---------------------------------------------------------------- GEMM1 issuer
    // GEMM1 and GEMM2 have independent dependency chains (A1/D1 vs A2/D2 barriers), so each gets its own issuing
    // warp: the tensor pipe takes the instructions in arrival order, neither chain waits behind the other's
    // operands, and the ~10 SASS instructions of descriptor set-up per tcgen05.mma are spread over two warps.
    pdl_wait();
    mbar_wait(bar_w, 0);
    constexpr uint32_t IDESC_G1 = make_idesc_bf16(128, 128, 0, 0);
    const uint32_t sW = smem_u32(smem + FF3_W);
    const uint64_t dA1h = desc_kmajor(smem_u32(smem + FF3_A1), 0), dA1l = desc_kmajor(smem_u32(smem + FF3_A1) + 16384u, 0);
    const uint64_t dW1h = desc_kmajor(sW, 0), dW1l = desc_kmajor(sW + 32768u, 0);
    constexpr uint64_t kStage = 32768 >> 4, kHalf = 16384 >> 4;
    int tile = (int)blockIdx.x - (int)gridDim.x, layer = 0;
    uint32_t w_phase = 1u, free_phase = 0u;
    for (int n = 0; n < my_tiles; ++n) {
      tile += gridDim.x;
      if (kPipe && tile >= n_tiles) {
        // next layer: once every GEMM1 issued so far has retired, replace the W1 image (the GEMM2s, chunk epilogues and
        // stores of the previous tile keep running meanwhile) and wait for it to land
        tile -= n_tiles;
        ++layer;
        umma_commit_elect(&w_free[0]);
        mbar_wait(&w_free[0], free_phase);
        free_phase ^= 1u;
        if (lane == 0) {
          mbar_expect_tx(bar_w, 65536);
          const uint8_t* img = a.layers[layer].image;
          for (int i = 0; i < 2; ++i) bulk_g2s(smem + FF3_W + i * 32768, img + i * 32768, 32768, bar_w);
        }
        __syncwarp();
        mbar_wait(bar_w, w_phase);
        w_phase ^= 1u;
      }
      const int stn = n & 1;
      mbar_wait(&a1_full[stn], (uint32_t)(n >> 1) & 1u);
      if (lane == 0) TL(2, n, 0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&d1_empty[h], ((uint32_t)n & 1u) ^ 1u);
        tc_fence_after();
        if (lane == 0) TL(2, n, 1 + 2 * h);
        issue3_kmajor_elect<4>(tmem + (uint32_t)(h * 128), dA1h + stn * kStage, dA1l + stn * kStage, dW1h + h * kHalf,
                               dW1l + h * kHalf, IDESC_G1, 0u, KO_PASSES(32));
        umma_commit_elect(&d1_full[h]);
        if (h == 1) umma_commit_elect(&a1_empty[stn]);
        if (lane == 0) TL(2, n, 2 + 2 * h);
      }
    }
    __syncwarp();
  } else if (warp == kFF3G2Warp) {
    // ---------------------------------------------------------------- GEMM2 issuer (A operand from tensor memory)
    pdl_wait();
    mbar_wait(bar_w2, 0);
    constexpr uint32_t IDESC_G2 = make_idesc_bf16(128, 64, 0, 0);
    const uint32_t sW = smem_u32(smem + FF3_W);
    const uint64_t dW2h = desc_kmajor(sW + 65536u, 0), dW2l = desc_kmajor(sW + 98304u, 0);
    constexpr uint64_t kBlk = 8192 >> 4;
    int tile = (int)blockIdx.x - (int)gridDim.x, layer = 0;
    uint32_t w_phase = 1u, free_phase = 0u;
    for (int n = 0; n < my_tiles; ++n) {
      tile += gridDim.x;
      if (kPipe && tile >= n_tiles) {             // next layer: same hand-over for the W2 image
        tile -= n_tiles;
        ++layer;
        umma_commit_elect(&w_free[1]);
        mbar_wait(&w_free[1], free_phase);
        free_phase ^= 1u;
        if (lane == 0) {
          mbar_expect_tx(bar_w2, 65536);
          const uint8_t* img = a.layers[layer].image;
          for (int i = 2; i < 4; ++i) bulk_g2s(smem + FF3_W + i * 32768, img + i * 32768, 32768, bar_w2);
        }
        __syncwarp();
        mbar_wait(bar_w2, w_phase);
        w_phase ^= 1u;
      }
      const int ds = n & 1;
      const uint32_t d2 = tmem + (uint32_t)(256 + ds * 64);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int team = j & 1, q = 2 * n + (j >> 1);
        const uint32_t a_hi = tmem + (uint32_t)(384 + team * 64), a_lo = a_hi + 32u;
        const uint64_t bh = dW2h + j * kBlk, bl = dW2l + j * kBlk;
#pragma unroll
        for (int part = 0; part < 2; ++part) {      // K = 32 per operand stage: 2 K steps x 3 passes
          const int stage = team * 2 + part;
          if (lane == 0 && j == 0 && part == 0) TL(4, n, 0);
          mbar_wait(&a2_full[stage], (uint32_t)q & 1u);
          if (lane == 0 && j == 0 && part == 0) TL(4, n, 1);
          if (j == 0 && part == 0) mbar_wait(&d2_empty[ds], ((uint32_t)(n >> 1) & 1u) ^ 1u);
          tc_fence_after();
          if (lane == 0 && j == 0 && part == 0) TL(4, n, 2);
          if (lane == 0 && j == 0 && part == 1) TL(4, n, 4);
#pragma unroll
          for (int pass = 0; pass < KO_PASSES(32); ++pass) {
            const uint32_t a = (pass == 2) ? a_lo : a_hi;
            const uint64_t b = (pass == 1) ? bl : bh;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const int kk = part * 2 + ks;
              umma_bf16_ts_elect(d2, a + (uint32_t)(kk * 8), b + (uint64_t)(kk * 2), IDESC_G2,
                                 (pass == 0 && ks == 0 && part == 0) ? (j > 0 ? 1u : 0u) : 1u);
            }
          }
          umma_commit_elect(&a2_empty[stage]);
          if (lane == 0 && j == 0) TL(4, n, part == 0 ? 3 : 5);
        }
        if (j == 3) umma_commit_elect(&d2_full[ds]);
        if (lane == 0 && j == 1) TL(4, n, 6);
        if (lane == 0 && j == 3) TL(4, n, 7);
      }
    }
    __syncwarp();
  } else if (kPipe && warp == kFF3PollWarp) {
    if (lane == 0) {      // poller: the unit's inverse transforms of every axis
      const int tpu = a.pipe.tiles_per_unit[0];
      int tile = (int)blockIdx.x - (int)gridDim.x, layer = 0, last = -1;
      for (int n = 0; n < my_tiles; ++n) {
        tile += gridDim.x;
        if (tile >= n_tiles) { tile -= n_tiles; ++layer; }
        const int unit = tile / tpu, seq = layer * (n_tiles + 1) + unit;
        if (seq != last) {
          pipe_wait(a.pipe.wait_ctr[0] + unit * kCtrStride, (unsigned)(layer + 1) * a.pipe.wait_count[0]);
          pipe_wait(a.pipe.wait_ctr[1] + unit * kCtrStride, (unsigned)(layer + 1) * a.pipe.wait_count[1]);
          st_release_cta_smem(s_seq, seq);
          last = seq;
        }
      }
    }
  } else if (kPipe && warp == kFF3PubWarp) {
    if (lane == 0) {      // publisher
      const int tpu = a.pipe.tiles_per_unit[0];
      int tile = (int)blockIdx.x - (int)gridDim.x, need = 0;
      for (int n = 0; n < my_tiles; ++n) {
        tile += gridDim.x;
        if (tile >= n_tiles) tile -= n_tiles;
        need += 4;
        smem_wait_ge(s_pub, need);
        pipe_publish(a.pipe.done_ctr[0] + (tile / tpu) * kCtrStride);
      }
    }
  } else {
    // ---------------------------------------------------------------- loaders: (s0 + s1 + s2) tile -> A1[stage]
    // Two rounds of 8 float4 per source and thread (64 data registers live at most: more spills under the 96-register
    // cap of a 17-warp CTA); the next tile of every source, and of the residual, is pulled into L2 by one bulk
    // prefetch per buffer so that the rounds see L2 latency, not DRAM latency.
    const int lt = tid - 13 * 32;        // warps 13..16
    auto prefetch_tile = [&](int tile) {
      const long long row0 = (long long)(reverse ? n_tiles - 1 - tile : tile) * 128;
      const long long rows = (P - row0) < 128 ? (P - row0) : 128;
      const uint32_t bytes = (uint32_t)rows * 256u;
      prefetch_l2_bulk(s0 + row0 * 64, bytes);
      if (s1) prefetch_l2_bulk(s1 + row0 * 64, bytes);
      if (s2) prefetch_l2_bulk(s2 + row0 * 64, bytes);
      if (residual) prefetch_l2_bulk(residual + row0 * 64, bytes);
    };
    if (!kPipe && lt == 0 && (int)blockIdx.x < n_tiles) prefetch_tile(blockIdx.x);
    const int lr = lt >> 4, lc = lt & 15;      // rows lr + 8 i (i = 0..15), float4 column lc
    const uint32_t a_off = (uint32_t)lr * 128u + (uint32_t)(((lc >> 1) ^ lr) << 4) + (uint32_t)(lc & 1) * 8u;   // + 1024 i
    int tile = (int)blockIdx.x - (int)gridDim.x, layer = 0, ready_seq = -1;
    long long dbg_blocked = 0;
    const long long dbg_c0 = kPipe ? clock64() : 0;
    const unsigned long long dbg_t0 = kPipe ? globaltimer_ns() : 0ull;
    for (int n = 0; n < my_tiles; ++n) {
      tile += gridDim.x;
      if (kPipe && tile >= n_tiles) { tile -= n_tiles; ++layer; }
      if (kPipe) {      // the unit's inverse transforms of every axis must be complete
        const int unit = tile / a.pipe.tiles_per_unit[0];
        const int seq = layer * (n_tiles + 1) + unit;
        if (seq != ready_seq) {
          dbg_blocked += wait_unit_ready(s_seq, seq);
          ready_seq = seq;
          if (a.pipe.dbg_ts && lt == 0 && blockIdx.x == 0) {
            const int q = layer * (n_tiles / a.pipe.tiles_per_unit[0]) + unit;
            if (q < 64) a.pipe.dbg_ts[q] = globaltimer_ns();
          }
        }
      }
      const long long row0 = (long long)(reverse ? n_tiles - 1 - tile : tile) * 128;
      const int rows_left = (P - row0) < 128 ? (int)(P - row0) : 128;
      const long long gofs = row0 * 64 + lr * 64 + lc * 4;
      const int st = n & 1;
      if (lt < 32) TL(3, n, 0);
      if (!kPipe && lt == 0 && tile + (int)gridDim.x < n_tiles) prefetch_tile(tile + (int)gridDim.x);
      if (kPipe && lt == 0 && n + 1 < my_tiles) {
        // pull the residual rows of this CTA's next tile into L2: they were written a whole layer ago (evicted since),
        // and the store warps read them inside their serial per-tile chain.  (Across a layer boundary the rows may
        // still be in the making: a prefetch of a line that is rewritten later is harmless, L2 is the coherence point.)
        int nt = tile + (int)gridDim.x, nl = layer;
        if (nt >= n_tiles) { nt -= n_tiles; ++nl; }
        prefetch_l2_bulk(a.xbuf[nl & 1] + (long long)nt * 128 * 64, 128u * 256u);
      }
      uint8_t* sA1h = smem + FF3_A1 + st * 32768 + a_off;
      uint8_t* sA1l = sA1h + 16384;
      const float* p0 = s0 + gofs;
      const float* p1 = s1 ? s1 + gofs : p0;
      const float* p2 = s2 ? s2 + gofs : p0;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float4 v[8], t[8];
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int i = half * 8 + it;
#if FFNO_KO & 64
          v[it] = make_float4(0.1f * (float)lr, 0.2f, 0.3f, 0.4f);
#else
          v[it] = (i * 8 + lr < rows_left) ? (kPipe ? ldg_cg(p0 + i * 512) : ldg_stream(p0 + i * 512)) : make_float4(0.f, 0.f, 0.f, 0.f);
#endif
        }
#if FFNO_KO & 64
        if (false) {
#else
        if (s1) {
#endif
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int i = half * 8 + it;
            t[it] = (i * 8 + lr < rows_left) ? (kPipe ? ldg_cg(p1 + i * 512) : ldg_stream(p1 + i * 512)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            v[it].x += t[it].x; v[it].y += t[it].y; v[it].z += t[it].z; v[it].w += t[it].w;
          }
        }
        if (s2) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int i = half * 8 + it;
            t[it] = (i * 8 + lr < rows_left) ? (kPipe ? ldg_cg(p2 + i * 512) : ldg_stream(p2 + i * 512)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            v[it].x += t[it].x; v[it].y += t[it].y; v[it].z += t[it].z; v[it].w += t[it].w;
          }
        }
        if (half == 0) {
          if (lt < 32) TL(3, n, 1);
          mbar_wait(&a1_empty[st], ((uint32_t)(n >> 1) & 1u) ^ 1u);
          if (lt < 32) TL(3, n, 2);
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) store_split4_at(sA1h, sA1l, (uint32_t)(half * 8 + it) * 1024u, v[it]);
      }
      fence_proxy_async_smem();
      mbar_arrive(&a1_full[st]);
      if (lt < 32) TL(3, n, 3);
    }
    if (kPipe && lt == 0 && a.pipe.dbg) {
      unsigned long long* d = a.pipe.dbg + 4 * blockIdx.x;
      d[0] = (unsigned long long)dbg_blocked;
      d[1] = (unsigned long long)(clock64() - dbg_c0);
      d[2] = dbg_t0;
      d[3] = globaltimer_ns();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_ff_ts(const float* s0, const float* s1, const float* s2, const float* residual, float* x_out, float* b_out,
                 const uint8_t* image, const float* b1, const float* b2, long long P, int sm_count, cudaStream_t st,
                 const float* head_w, const float* head_b, float* forecast, bool reverse) {
  if (P == 0) return FFNO_OK;
  FFNO_TRY(ensure_dynamic_smem(ff_ts_kernel<false>, FF3_TOTAL));
  const int n_tiles = ceil_div(P, 128);
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  FFArgs a{};
  a.s0 = s0; a.s1 = s1; a.s2 = s2; a.residual = residual; a.x_out = x_out; a.b_out = b_out;
  a.image = image; a.b1 = b1; a.b2 = b2; a.head_w = head_w; a.head_b = head_b; a.forecast = forecast;
  a.P = P; a.n_tiles = n_tiles; a.reverse = reverse ? 1 : 0;
  FFNO_CUDA_CHECK(launch_pdl(ff_ts_kernel<false>, dim3(grid), dim3(kFF3Threads), (size_t)FF3_TOTAL, st, a));
  ++g_launch_counter;
  return FFNO_OK;
}


// =======================================================================================================
// Stage-pipelined forward of the whole layer stack: four launches per forward (see PipeDesc).
// =======================================================================================================
static int pipe_axis_set(const AxisXform* axes, int n_axes, int n_units, const float* x_odd, AxisSet* set, size_t* smem,
                         int* max_tiles) {
  *smem = 0;
  *max_tiles = 0;
  set->reverse = 0;
  for (int a = 0; a < n_axes; ++a) {
    const AxisXform& p = axes[a];
    FFNO_REQUIRE(p.inner % 64 == 0 && p.inner < (1ll << 23) && p.npad >= 16 && p.npad <= 256 && p.npad % 16 == 0,
                 FFNO_ERR_UNSUPPORTED, "stack_pipe: axis %d geometry", a);
    const int ring = axis_ring_depth(p.n_in, p.n_out);
    FFNO_REQUIRE(ring > 0, FFNO_ERR_UNSUPPORTED, "stack_pipe: table does not fit in shared memory");
    set->ring[a] = ring;
    const size_t need = axp_table_offset(ring) + table_image_bytes(p.n_in, p.n_out);
    *smem = need > *smem ? need : *smem;
    set->ax[a] = p;
    const long long n_groups = p.outer * (p.inner / 64);
    FFNO_REQUIRE(n_groups % (2 * n_units) == 0 && n_groups < (1ll << 31), FFNO_ERR_UNSUPPORTED,
                 "stack_pipe: %lld groups of axis %d do not split into %d units of whole tiles", n_groups, a, n_units);
    set->n_tiles[a] = (int)(n_groups / 2);
    set->pipe.tiles_per_unit[a] = set->n_tiles[a] / n_units;
    FFNO_TRY(encode_axis_tmap(&set->tm[a], p.X, p.outer, p.n_in, p.inner));
    FFNO_TRY(encode_axis_tmap(&set->tm_odd[a], x_odd ? x_odd : p.X, p.outer, p.n_in, p.inner));
    int cols = 32;
    while (cols < 2 * p.npad) cols *= 2;
    set->tmem_cols[a] = cols;
    *max_tiles = set->n_tiles[a] > *max_tiles ? set->n_tiles[a] : *max_tiles;
  }
  return FFNO_OK;
}


size_t stack_pipe_counter_bytes(int n_units) { return (size_t)10 * n_units * kCtrStride * sizeof(unsigned); }

bool stack_pipe_geometry_ok(const AxisXform* fwd, const MixAxis* mix, int n_axes, long long P, int n_units) {
  if (n_axes != 2 || n_units < 4 || P % (128ll * n_units) != 0) return false;
  for (int a = 0; a < n_axes; ++a) {
    const long long n_groups = fwd[a].outer * (fwd[a].inner / 64);
    if (fwd[a].inner % 64 != 0 || n_groups % (2 * n_units) != 0) return false;
    const long long rows = mix[a].outer * mix[a].p_inner;
    if (rows % (128ll * n_units) != 0) return false;
  }
  return true;
}

int launch_stack_pipe(const StackPipeArgs& A) {
  const int n_units = A.n_units, na = A.n_axes, L = A.n_layers;
  FFNO_REQUIRE(na == 2, FFNO_ERR_UNSUPPORTED, "stack_pipe: %d axes", na);
  FFNO_REQUIRE(stack_pipe_geometry_ok(A.fwd, A.mix, na, A.P, n_units), FFNO_ERR_UNSUPPORTED, "stack_pipe: geometry");
  const size_t cs = (size_t)n_units * kCtrStride;          // one 128-byte line per counter
  unsigned* ctr_fwd[3], *ctr_mix[3], *ctr_inv[3], *ctr_ff = A.counters + 9 * cs;
  for (int a = 0; a < 3; ++a) {
    ctr_fwd[a] = A.counters + a * cs;
    ctr_mix[a] = A.counters + (3 + a) * cs;
    ctr_inv[a] = A.counters + (6 + a) * cs;
  }
  const int ff_tiles = (int)(A.P / 128), ff_tpu = ff_tiles / n_units;
  const unsigned ff_arrivals = (unsigned)ff_tpu;              // the publisher warps post one arrival per finished tile

  AxisSet fset{}, iset{};
  size_t fsmem, ismem;
  int fmax, imax;
  FFNO_TRY(pipe_axis_set(A.fwd, na, n_units, A.x_odd, &fset, &fsmem, &fmax));
  FFNO_TRY(pipe_axis_set(A.inv, na, n_units, nullptr, &iset, &ismem, &imax));
  MixSet mset{};
  int maxK = 0;
  unsigned fwd_arr[3] = {0, 0, 0}, mix_arr[3] = {0, 0, 0}, inv_arr[3] = {0, 0, 0};
  for (int a = 0; a < na; ++a) {
    mset.ax[a] = A.mix[a];
    const int tiles = (int)(A.mix[a].outer * A.mix[a].p_inner / 128);
    mset.tiles_per_cta[a] = tiles;
    mset.pipe.tiles_per_unit[a] = tiles / n_units;
    maxK = A.mix[a].K > maxK ? A.mix[a].K : maxK;
    fwd_arr[a] = (unsigned)fset.pipe.tiles_per_unit[a];
    mix_arr[a] = (unsigned)A.mix[a].K * (unsigned)mset.pipe.tiles_per_unit[a];
    inv_arr[a] = (unsigned)iset.pipe.tiles_per_unit[a];
  }
  // forward transforms: wait for the FF of the previous layer (nothing at layer 0), read x of the layer's parity
  fset.pipe.n_layers = L;
  fset.pipe.wait_lag = 0;
  for (int a = 0; a < na; ++a) {
    fset.pipe.wait_ctr[a] = ctr_ff;
    fset.pipe.wait_count[a] = ff_arrivals;
    fset.pipe.done_ctr[a] = ctr_fwd[a];
    fset.X_odd[a] = A.x_odd;
    iset.X_odd[a] = A.inv[a].X;
  }
  mset.reverse = 0;
  mset.pipe.n_layers = L;
  mset.pipe.wait_lag = 1;
  iset.pipe.n_layers = L;
  iset.pipe.wait_lag = 1;
  for (int a = 0; a < na; ++a) {
    mset.pipe.wait_ctr[a] = ctr_fwd[a];
    mset.pipe.wait_count[a] = fwd_arr[a];
    mset.pipe.done_ctr[a] = ctr_mix[a];
    iset.pipe.wait_ctr[a] = ctr_mix[a];
    iset.pipe.wait_count[a] = mix_arr[a];
    iset.pipe.done_ctr[a] = ctr_inv[a];
  }
  FFArgs fa{};
  fa.s0 = A.inv[0].Y;
  fa.s1 = A.inv[1].Y;
  fa.s2 = nullptr;
  fa.head_w = A.head_w;
  fa.head_b = A.head_b;
  fa.forecast = A.forecast;
  fa.P = A.P;
  fa.n_tiles = ff_tiles;
  fa.layers = A.ff_layers;
  fa.xbuf[0] = A.xbuf[0];
  fa.xbuf[1] = A.xbuf[1];
  fa.pipe.n_layers = L;
  fa.pipe.tiles_per_unit[0] = ff_tpu;
  fa.pipe.wait_lag = 1;
  fa.pipe.wait_ctr[0] = ctr_inv[0];
  fa.pipe.wait_ctr[1] = ctr_inv[1];
  fa.pipe.wait_count[0] = inv_arr[0];
  fa.pipe.wait_count[1] = inv_arr[1];
  fa.pipe.done_ctr[0] = ctr_ff;

  // SM budget: one CTA per (axis, mode) for the mix; the rest split between the three streaming stages
  int g_fwd = A.sms[0] / na, g_inv = A.sms[2] / na, g_ff = A.sms[3];
  g_fwd = g_fwd < 1 ? 1 : (g_fwd > fmax ? fmax : g_fwd);
  g_inv = g_inv < 1 ? 1 : (g_inv > imax ? imax : g_inv);
  g_ff = g_ff < 1 ? 1 : (g_ff > ff_tiles ? ff_tiles : g_ff);
  for (int a = 0; a < na; ++a)
    FFNO_REQUIRE(g_fwd <= fset.n_tiles[a] && g_inv <= iset.n_tiles[a], FFNO_ERR_UNSUPPORTED, "stack_pipe: grid > tiles");

  FFNO_TRY(ensure_dynamic_smem(axis_pipe_kernel<true>, fsmem > ismem ? fsmem : ismem));
  FFNO_TRY(ensure_dynamic_smem(mix_pipe_kernel<true>, MXP_TOTAL));
  FFNO_TRY(ensure_dynamic_smem(ff_ts_kernel<true>, FF3_TOTAL));

  if (A.dbg) {
    unsigned long long* d = A.dbg;
    d += 8;                                     // header written by the host wrapper: grid sizes
    fset.pipe.dbg = d;
    d += 4 * na * g_fwd;
    mset.pipe.dbg = d;
    d += 4 * na * maxK;
    iset.pipe.dbg = d;
    d += 4 * na * g_inv;
    fa.pipe.dbg = d;
    unsigned long long* ts = A.dbg + 8 + 4 * 256;
    fset.pipe.dbg_ts = ts;
    mset.pipe.dbg_ts = ts + 64;
    iset.pipe.dbg_ts = ts + 128;
    fa.pipe.dbg_ts = ts + 192;
    const unsigned long long hdr[8] = {(unsigned long long)(na * g_fwd), (unsigned long long)(na * maxK),
                                       (unsigned long long)(na * g_inv), (unsigned long long)g_ff, 0, 0, 0, 0};
    FFNO_CUDA_CHECK(cudaMemcpyAsync(A.dbg, hdr, sizeof(hdr), cudaMemcpyHostToDevice, A.streams[0]));
  }
  cudaStream_t st = A.streams[0];
  // diagnostics (A.only_stage >= 0): one stage alone with every dependency pre-satisfied — its raw throughput
  FFNO_CUDA_CHECK(cudaMemsetAsync(A.counters, A.only_stage >= 0 ? 0x3f : 0, stack_pipe_counter_bytes(n_units), st));
  FFNO_CUDA_CHECK(cudaEventRecord(A.fork, st));
  for (int i = 1; i < 4; ++i) FFNO_CUDA_CHECK(cudaStreamWaitEvent(A.streams[i], A.fork, 0));
  const int only = A.only_stage;
  if (only < 0 || only == 0) axis_pipe_kernel<true><<<dim3(g_fwd, na), kAxPipeThreads, fsmem, A.streams[0]>>>(fset);
  if (only < 0 || only == 1) mix_pipe_kernel<true><<<dim3(1, maxK, na), kPipeThreads, MXP_TOTAL, A.streams[1]>>>(mset);
  if (only < 0 || only == 2) axis_pipe_kernel<true><<<dim3(g_inv, na), kAxPipeThreads, ismem, A.streams[2]>>>(iset);
  if (only < 0 || only == 3) ff_ts_kernel<true><<<dim3(g_ff), kFF3PipeThreads, FF3_TOTAL, A.streams[3]>>>(fa);
  g_launch_counter += 4;
  const cudaError_t e = cudaGetLastError();
  for (int i = 1; i < 4; ++i) {          // always rejoin (a forked stream must rejoin a capture even after an error)
    cudaEventRecord(A.join[i - 1], A.streams[i]);
    cudaStreamWaitEvent(st, A.join[i - 1], 0);
  }
  if (e != cudaSuccess) return set_error(FFNO_ERR_CUDA, "stack_pipe launch failed: %s", cudaGetErrorString(e));
  return FFNO_OK;
}

// Two tiny kernels that can only both succeed when they run at the same time (each raises its flag and waits ~1 ms for
// the other's): tells whether kernels of different streams really execute concurrently in this process — they do not
// under a profiler that serialises launches or with CUDA_LAUNCH_BLOCKING=1, where the stage-pipelined forward must not
// be used (its stages wait for each other).
__global__ void concurrency_probe_kernel(unsigned* flags, int me) {
  atomicExch(&flags[me], 1u);
  __threadfence();
  const long long t0 = clock64();
  unsigned seen = 0;
  while (clock64() - t0 < 2000000ll) {
    seen = ld_acquire_gpu(&flags[1 - me]);
    if (seen) break;
    __nanosleep(200);
  }
  flags[2 + me] = seen;
}
int probe_stream_concurrency(cudaStream_t s0, cudaStream_t s1, unsigned* dev_flags4, bool* concurrent) {
  *concurrent = false;
  FFNO_CUDA_CHECK(cudaMemsetAsync(dev_flags4, 0, 4 * sizeof(unsigned), s0));
  FFNO_CUDA_CHECK(cudaStreamSynchronize(s0));
  concurrency_probe_kernel<<<1, 1, 0, s0>>>(dev_flags4, 0);
  concurrency_probe_kernel<<<1, 1, 0, s1>>>(dev_flags4, 1);
  g_launch_counter += 2;
  FFNO_LAUNCH_CHECK("concurrency_probe_kernel");
  FFNO_CUDA_CHECK(cudaStreamSynchronize(s0));
  FFNO_CUDA_CHECK(cudaStreamSynchronize(s1));
  unsigned h[4] = {0, 0, 0, 0};
  FFNO_CUDA_CHECK(cudaMemcpy(h, dev_flags4, sizeof(h), cudaMemcpyDeviceToHost));
  *concurrent = h[2] != 0 && h[3] != 0;
  return FFNO_OK;
}

int debug_timeline(int enable, long long* host_out /*[1024] or NULL*/) {
  if (host_out) FFNO_CUDA_CHECK(cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(long long) * 1024));
  if (enable >= 0) {
    long long zero[1024] = {0};
    FFNO_CUDA_CHECK(cudaMemcpyToSymbol(g_timeline, zero, sizeof(zero)));
    FFNO_CUDA_CHECK(cudaMemcpyToSymbol(g_timeline_on, &enable, sizeof(int)));
  }
  return FFNO_OK;
}

}  // namespace ffno
