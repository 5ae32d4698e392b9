// Warp-specialised, mbarrier-pipelined tcgen05 kernels of the F-FNO layer (width 64 / hidden 256), sm_100a.
//
// One CTA per SM, 9 warps:
//   warps 0-3  epilogue   (TMEM lanes 0..127 -> registers -> bias/activation/split or global stores)
//   warp  4    MMA issuer (lane 0 issues tcgen05.mma + tcgen05.commit)
//   warps 5-8  loaders    (coalesced FP32 global loads -> BF16 hi/lo split -> swizzled operand tiles in smem)
// Operand tiles in shared memory and accumulators in tensor memory are double-buffered and handed over with
// full/empty mbarriers, so the global loads of tile t+1, the MMAs of tile t and the epilogue of tile t-1 overlap.
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
#define TL(role, tile_n, ev)                                                                      \
  do {                                                                                            \
    if (blockIdx.x == 0 && (tile_n) < 16 && (threadIdx.x & 31) == 0)    /* store only: no flag load */ \
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
constexpr int kMmaWarp = kEpiWarps;
constexpr int kLoaderThread0 = (kEpiWarps + 1) * 32;
constexpr int kThreads = kLoaderThread0 + kLoaders;       // 800
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

template <int KSTEPS>
__device__ __forceinline__ void issue3_kmajor_elect(uint32_t tmem_d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi,
                                                    uint64_t b_lo, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
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

}  // namespace

// =======================================================================================================
// Axis transform (forward / inverse truncated DFT), up to 3 axes per launch (blockIdx.y = axis)
// =======================================================================================================
constexpr int AXP_A_STAGE = 32768;                 // operand stage: hi 16 KB | lo 16 KB
constexpr int AXP_BAR = 2 * AXP_A_STAGE;           // 65536
constexpr int AXP_STAGING = AXP_BAR + 1024;        // ring x 32 KB of raw FP32 (cp.async landing zone)
constexpr int kAxStages = 2;                       // ring depth; 1 when a large table (C4: 128 KB) leaves no room for 2
constexpr int axp_table_offset(int ring) { return AXP_STAGING + ring * 32768; }   // the DFT table image follows the ring

struct AxisSet {
  AxisXform ax[3];
  int n_tiles[3];
  int tmem_cols[3];
  int ring[3];      // cp.async staging ring depth of the axis (1 or 2)
  int reverse;      // walk the tiles from the last to the first (see launch_axis_pipe)
};

__global__ void __launch_bounds__(kThreads, 1) axis_pipe_kernel(AxisSet set) {
  const AxisXform& p = set.ax[blockIdx.y];
  const int n_tiles = set.n_tiles[blockIdx.y];
  if ((int)blockIdx.x >= n_tiles) return;
  const int tmem_cols = set.tmem_cols[blockIdx.y];
  const int stage_cols = tmem_cols >> 1;

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AXP_BAR);
  uint64_t* a_full = bars;          // [2]
  uint64_t* a_empty = bars + 2;     // [2]
  uint64_t* d_full = bars + 4;      // [2]
  uint64_t* d_empty = bars + 6;     // [2]
  uint64_t* bar_w = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  const int ring = set.ring[blockIdx.y];
  uint8_t* sB = smem + axp_table_offset(ring);
  const uint32_t b_half = (uint32_t)p.kchunks * p.npad * 128u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
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

  if (warp < kEpiWarps) {
    // ---------------------------------------------------------------- epilogue: 8 independent warps
    // TMEM lane = inner element, column = output index, and inner is the contiguous dimension of Y: a warp's 32 lanes
    // of one column are 128 contiguous bytes in global memory, so every warp drains its own lane quadrant straight
    // from registers with one STG.32 per column (one full line per instruction, no staging, no block barrier).
    //   npad <= 32 (one chunk per tile): warps 0-3 / 4-7 take alternate tiles, each drains a stage alone;
    //   npad  > 32: both sets work on every tile, set t takes chunks t, t+2, ... and both release the stage.
    const int team = warp >> 2, quad = warp & 3;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int total_chunks = (p.npad + 31) >> 5;
    const bool split_tiles = total_chunks == 1;
    const int gq = quad >> 1;                                  // which of the tile's two 64-element groups
    const int in_group = (quad & 1) * 32 + lane;
    const unsigned stride = (unsigned)p.inner;
    int n = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      const int ds = n & 1;
      if (split_tiles && ds != team) continue;
      mbar_wait(&d_full[ds], (uint32_t)(n >> 1) & 1u);
      tc_fence_after();
      if (warp == 0 && blockIdx.y == 0) TL(5, n, 0);
      const int ptile = set.reverse ? n_tiles - 1 - tile : tile;
      const long long G = (long long)ptile * 2 + gq;             // warp-uniform
      const bool live = G < n_groups;
      const unsigned uo = live ? (unsigned)G / (unsigned)gpi : 0u;
      const unsigned ug = live ? (unsigned)G - uo * (unsigned)gpi : 0u;
      float* ybase = p.Y + ((long long)uo * p.n_out) * p.inner + (long long)ug * 64 + in_group;
      const int ch_begin = split_tiles ? 0 : team, ch_step = split_tiles ? 1 : 2;
      bool released = false;
#pragma unroll 1
      for (int ch = ch_begin; ch < total_chunks; ch += ch_step) {
        const int c0 = ch * 32;
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + (uint32_t)(ds * stage_cols + c0), v);
        tmem_ld_wait();
        if (ch + ch_step >= total_chunks) {   // this warp's last read of the stage
          tc_fence_before();
          mbar_arrive(&d_empty[ds]);
          released = true;
          if (warp == 0 && blockIdx.y == 0) TL(5, n, 1);
        }
        if (live) {
          float* yp = ybase + (size_t)c0 * stride;
          const int ncols = p.n_out - c0;                        // >= 32: full chunk
          if (!p.accumulate && ncols >= 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) yp[(size_t)j * stride] = __uint_as_float(v[j]);
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
      if (warp == 0 && blockIdx.y == 0) TL(5, n, 2);
    }
  } else if (warp == kMmaWarp) {
    // ---------------------------------------------------------------- MMA issuer
    {   // the whole warp walks the schedule (warp-uniform); elect.sync picks the issuing lane per instruction
      pdl_wait();
      mbar_wait(bar_w, 0);
      const uint32_t idesc = make_idesc_bf16(128, p.npad, 1, 0);
      int item = 0, n = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
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
            for (int pass = 0; pass < 3; ++pass) {
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
  } else {
    // ---------------------------------------------------------------- loaders
    // Each thread streams its own 16 x 16 B of every work item through a private slice of the staging ring with
    // cp.async (kAxStages items = 96 KB in flight per CTA, no registers held), then converts the item that has
    // landed: FP32 -> BF16 hi/lo, MN-major SWIZZLE_128B operand stage.  A thread only ever reads what it copied,
    // so the ring needs no cross-thread synchronisation (per-thread cp.async groups).
    const int lt = tid - kLoaderThread0;
    const int my_tiles = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int n_items = my_tiles * p.kchunks;
    const int gsel = (lt >> 4) & 1, c4 = lt & 15, rsub = lt >> 5;            // rsub: 0..15
    uint8_t* stg_base = smem + AXP_STAGING + lt * 16;
    const long long row_step = 16 * p.inner;                 // floats between the rows this thread copies
    // The issue sequence walks this CTA's tiles in order, so the (outer, group) pair of the next tile follows from
    // the previous one by a precomputed step: two divisions per kernel instead of two per work item.
    const int sgn = set.reverse ? -1 : 1;
    const unsigned step_groups = 2u * gridDim.x;
    const int step_o = (int)(step_groups / (unsigned)gpi), step_g = (int)(step_groups % (unsigned)gpi);
    long long curG = (long long)(set.reverse ? n_tiles - 1 - (int)blockIdx.x : (int)blockIdx.x) * 2 + gsel;
    // (a dead group G == n_groups of an odd tail still gets its true (outer, group): it is stepped from, never read)
    int cur_o = (int)((unsigned)curG / (unsigned)gpi);
    int cur_g = (int)((unsigned)curG - (unsigned)cur_o * (unsigned)gpi);
    int iss_kc = 0, iss_left = n_items;
    const long long thr_off = (long long)rsub * p.inner + c4 * 4;
    auto issue = [&](int slot) {
      if (iss_left > 0) {
        --iss_left;
        const bool live = curG < n_groups;
        const int i0 = iss_kc * 64 + rsub;
        const float* src = p.X + ((long long)cur_o * p.n_in + iss_kc * 64) * p.inner + (long long)cur_g * 64 + thr_off;
        uint8_t* dst = stg_base + slot * 32768;
#pragma unroll
        for (int it = 0; it < kLdPerThread; ++it) {
          const bool ok = live && (i0 + it * 16) < p.n_in;
          cp_async16(dst + it * (kLoaders * 16), ok ? (const void*)src : (const void*)p.X, ok ? 16u : 0u);
          src += row_step;
        }
        if (++iss_kc == p.kchunks) {
          iss_kc = 0;
          curG += sgn * (long long)step_groups;
          cur_o += sgn * step_o;
          cur_g += sgn * step_g;
          if (cur_g >= (int)gpi) { cur_g -= (int)gpi; ++cur_o; }
          if (cur_g < 0) { cur_g += (int)gpi; --cur_o; }
        }
      }
      cp_async_commit();
    };
    for (int q = 0; q < ring; ++q) issue(q);
    for (int item = 0; item < n_items; ++item) {
      if (ring == 2) cp_async_wait<1>(); else cp_async_wait<0>();      // warp-uniform
      if (lt < 32 && blockIdx.y == 0) TL(7, item, 0);
      const uint8_t* src = stg_base + (item & (ring - 1)) * 32768;
      float4 v[kLdPerThread];
#pragma unroll
      for (int it = 0; it < kLdPerThread; ++it) v[it] = *reinterpret_cast<const float4*>(src + it * (kLoaders * 16));
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
        store_split4_at(sAh, sAl, off, v[it]);
      }
      fence_proxy_async_smem();
      mbar_arrive(&a_full[as]);
      if (lt < 32 && blockIdx.y == 0) TL(7, item, 2);
      issue(item & (ring - 1));      // refill the slot just drained with item + ring
      if (lt < 32 && blockIdx.y == 0) TL(7, item, 3);
    }
    cp_async_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

static int axis_ring_depth(int n_in, int n_out) {      // deepest staging ring that leaves room for the table, 0 = none
  for (int ring = kAxStages; ring >= 1; --ring)
    if (axp_table_offset(ring) + table_image_bytes(n_in, n_out) <= (size_t)227 * 1024) return ring;
  return 0;
}
bool axis_pipe_fits(int n_in, int n_out) { return n_out <= 256 && axis_ring_depth(n_in, n_out) > 0; }

int launch_axis_pipe(const AxisXform* axes, int n_axes, int sm_count, cudaStream_t st, bool reverse) {
  FFNO_REQUIRE(n_axes >= 1 && n_axes <= 3, FFNO_ERR_BAD_ARG, "axis_pipe: n_axes=%d", n_axes);
  AxisSet set;
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
  }
  if (max_tiles == 0) return FFNO_OK;
  FFNO_TRY(ensure_dynamic_smem(axis_pipe_kernel, smem));
  int per_axis = sm_count / n_axes;
  if (per_axis < 1) per_axis = 1;
  const int gx = max_tiles < per_axis ? max_tiles : per_axis;
  FFNO_CUDA_CHECK(launch_pdl(axis_pipe_kernel, dim3(gx, n_axes), dim3(kThreads), smem, st, set));
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
};

__global__ void __launch_bounds__(kThreads, 1) mix_pipe_kernel(MixSet set) {
  const MixAxis& ax = set.ax[blockIdx.z];
  const int k = blockIdx.y;
  if (k >= ax.K) return;
  const long long M = ax.outer * ax.p_inner;
  const int n_tiles = (int)((M + 127) / 128);
  const int tpc = set.tiles_per_cta[blockIdx.z];
  const int tile_begin = blockIdx.x * tpc;
  if (tile_begin >= n_tiles) return;
  const int tile_end = min(n_tiles, tile_begin + tpc);

  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sB = smem + MXP_B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MXP_BAR);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + 2;
  uint64_t* d_full = bars + 4;
  uint64_t* d_empty = bars + 6;
  uint64_t* bar_w = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
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

  if (warp < kEpiWarps) {
    // two epilogue teams: team 0 drains the real segment (columns 0..63), team 1 the imaginary one, each 32 columns
    // at a time through its own swizzled 16 KB staging tile so the global stores are full 128-byte lines
    const int team = warp >> 2, rt = tid & 127, seg = team;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint8_t* sOut = smem + MXP_OUT + team * 16384;
    const int rq = rt >> 3, cq = rt & 7;                    // coalesced phase: rows rq + 16 it, float4 column cq
    const uint8_t* so_rd = sOut + rq * 128 + ((cq ^ (rq & 7)) << 4);        // (rq + 16 it) & 7 == rq & 7
    const unsigned p_in = (unsigned)ax.p_inner;
    const unsigned q16 = 16u / p_in, r16 = 16u % p_in;
    const long long o_stride = (long long)ax.K * 2 * inner;
    float* out_seg = ax.R + ((long long)k * 2 + seg) * inner + cq * 4;
    int n = 0;
    for (int tile = tile_begin; tile < tile_end; ++tile, ++n) {
      const int ds = n & 1;
      mbar_wait(&d_full[ds], (uint32_t)(n >> 1) & 1u);
      tc_fence_after();
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
          const long long row_first = (long long)(set.reverse ? n_tiles - 1 - tile : tile) * 128 + rq;
          const unsigned rf = row_first < M ? (unsigned)row_first : 0u;
          unsigned uo = rf / p_in, pp = rf - uo * p_in;
          const int rows_left = (M - row_first) > 0 ? (int)((M - row_first) < 128 ? (M - row_first) : 128) : 0;   // rows rq .. M-1
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
    }
  } else if (warp == kMmaWarp) {
    {
      pdl_wait();
      mbar_wait(bar_w, 0);
      constexpr uint32_t IDESC = make_idesc_bf16(128, 128, 0, 0);
      const uint64_t dBh = desc_kmajor(smem_u32(sB), 0), dBl = desc_kmajor(smem_u32(sB) + 32768u, 0);
      const uint64_t dAh = desc_kmajor(smem_u32(smem), 0), dAl = desc_kmajor(smem_u32(smem) + 16384u, 0);
      int item = 0, n = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++n) {
        const int ds = n & 1;
        mbar_wait(&d_empty[ds], ((uint32_t)(n >> 1) & 1u) ^ 1u);
        const uint32_t d_addr = tmem + (uint32_t)(ds * 128);
        for (int kb = 0; kb < 2; ++kb, ++item) {
          const int as = item & 1;
          mbar_wait(&a_full[as], (uint32_t)(item >> 1) & 1u);
          tc_fence_after();
          const uint64_t a_off = (uint64_t)(as * (MXP_A_STAGE >> 4)), b_off = (uint64_t)(kb * (16384 >> 4));
          issue3_kmajor_elect<4>(d_addr, dAh + a_off, dAl + a_off, dBh + b_off, dBl + b_off, IDESC, kb > 0 ? 1u : 0u);
          umma_commit_elect(&a_empty[as]);
        }
        umma_commit_elect(&d_full[ds]);
      }
    }
    __syncwarp();
  } else {
    // cp.async staging ring, one private slice per thread (see axis_pipe_kernel)
    const int lt = tid - kLoaderThread0;
    const int n_items = (tile_end - tile_begin) * 2;
    const int c4 = lt & 15, rsub = lt >> 4;                                   // rsub: 0..31
    const unsigned p_in = (unsigned)ax.p_inner;
    uint8_t* stg_base = smem + MXP_STAGING + lt * 16;
    const unsigned q32 = 32u / p_in, r32 = 32u % p_in;
    const long long o_stride = (long long)ax.K * 2 * inner;
    const float* f_mode = ax.F + (long long)k * 2 * inner + c4 * 4;
    auto issue = [&](int item) {
      if (item < n_items) {
        const int ltile = tile_begin + (item >> 1), half = item & 1;
        const int tile = set.reverse ? n_tiles - 1 - ltile : ltile;
        uint8_t* dst = stg_base + (item % kMxStages) * 32768;
        const long long row_first = (long long)tile * 128 + rsub;          // rows rsub + 32 it
        const unsigned rf = row_first < M ? (unsigned)row_first : 0u;
        unsigned o = rf / p_in, pp = rf - o * p_in;
        const long long left = M - row_first;
        const float* f_half = f_mode + (long long)half * inner;
#pragma unroll
        for (int it = 0; it < kLdPerThread; ++it) {
          const bool ok = it * 32 < left;
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
      issue(item + kMxStages);
    }
    cp_async_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int launch_mix_pipe(const MixAxis* axes, int n_axes, int sm_count, cudaStream_t st, bool reverse) {
  FFNO_REQUIRE(n_axes >= 1 && n_axes <= 3, FFNO_ERR_BAD_ARG, "mix_pipe: n_axes=%d", n_axes);
  FFNO_TRY(ensure_dynamic_smem(mix_pipe_kernel, MXP_TOTAL));
  MixSet set;
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
  FFNO_CUDA_CHECK(launch_pdl(mix_pipe_kernel, dim3(grid_x, maxK, n_axes), dim3(kThreads), (size_t)MXP_TOTAL, st, set));
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
constexpr int FF3_BAR = FF3_BIAS + 320 * 4;        // 230656
constexpr int FF3_HEAD = FF3_BAR + 192;            // 64 floats: folded head weights of a 1-output head
constexpr int FF3_TOTAL = FF3_HEAD + 256;          // 231104 <= 232448
constexpr int kFF3Threads = 576;                   // 8 chunk-epilogue + 4 store + G1 issuer + 4 loader + G2 issuer warps
constexpr int kFF3G1Warp = 12, kFF3G2Warp = 17;

__global__ void __launch_bounds__(kFF3Threads, 1)  // 18 warps are allocated as 20: 96 registers/thread is the cap
ff_ts_kernel(const float* __restrict__ s0, const float* __restrict__ s1, const float* __restrict__ s2,
             const float* __restrict__ residual, float* __restrict__ x_out, float* __restrict__ b_out,
             const uint8_t* __restrict__ image, const float* __restrict__ b1, const float* __restrict__ b2,
             const float* __restrict__ head_w, const float* __restrict__ head_b, float* __restrict__ forecast,
             long long P, int n_tiles, int reverse) {
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

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
    int n = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int j = 2 * h + team;              // chunk of hidden units j*64 .. j*64+63
        const int q = 2 * n + h;                 // running chunk index of this team (its A2 stage is `team`)
        mbar_wait(&d1_full[h], (uint32_t)n & 1u);
        tc_fence_after();
        if ((warp & 3) == 0) TL(team ? 5 : 0, n, h * 4 + 0);
        const float* bj = sb1 + j * 64;
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
    int n = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      const long long row0 = (long long)(reverse ? n_tiles - 1 - tile : tile) * 128;
      const int rows_left = (P - row0) < 128 ? (int)(P - row0) : 128;
      const long long gofs = row0 * 64 + rq * 64 + cq * 4;
      if (warp == 8) TL(1, n, 0);
      // The tile is finished in two 32-channel halves so that only 8 float4 of the residual are live while the
      // accumulator is in registers: half 0 is fetched before the wait (its rows were pulled into L2 by the loaders'
      // bulk prefetch one tile earlier), half 1 once the accumulator registers are dead.
      float4 r0[8], r1[8];
      const float* rp = residual ? residual + gofs : s0 + gofs;
#pragma unroll
      for (int it = 0; it < 8; ++it)
        r0[it] = (residual && it * 16 + rq < rows_left) ? ldg_stream(rp + it * 1024) : make_float4(0.f, 0.f, 0.f, 0.f);
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
          const float4 bb = *reinterpret_cast<const float4*>(sb2 + c4 * 4);
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
        r1[it] = (residual && it * 16 + rq < rows_left) ? ldg_stream(rp + it * 1024 + 32) : make_float4(0.f, 0.f, 0.f, 0.f);
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
    int n = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      const int stn = n & 1;
      mbar_wait(&a1_full[stn], (uint32_t)(n >> 1) & 1u);
      if (lane == 0) TL(2, n, 0);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&d1_empty[h], ((uint32_t)n & 1u) ^ 1u);
        tc_fence_after();
        if (lane == 0) TL(2, n, 1 + 2 * h);
        issue3_kmajor_elect<4>(tmem + (uint32_t)(h * 128), dA1h + stn * kStage, dA1l + stn * kStage, dW1h + h * kHalf,
                               dW1l + h * kHalf, IDESC_G1, 0u);
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
    int n = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
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
          for (int pass = 0; pass < 3; ++pass) {
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
    if (lt == 0 && (int)blockIdx.x < n_tiles) prefetch_tile(blockIdx.x);
    const int lr = lt >> 4, lc = lt & 15;      // rows lr + 8 i (i = 0..15), float4 column lc
    const uint32_t a_off = (uint32_t)lr * 128u + (uint32_t)(((lc >> 1) ^ lr) << 4) + (uint32_t)(lc & 1) * 8u;   // + 1024 i
    int n = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++n) {
      const long long row0 = (long long)(reverse ? n_tiles - 1 - tile : tile) * 128;
      const int rows_left = (P - row0) < 128 ? (int)(P - row0) : 128;
      const long long gofs = row0 * 64 + lr * 64 + lc * 4;
      const int st = n & 1;
      if (lt < 32) TL(3, n, 0);
      if (lt == 0 && tile + (int)gridDim.x < n_tiles) prefetch_tile(tile + (int)gridDim.x);
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
          v[it] = (i * 8 + lr < rows_left) ? ldg_stream(p0 + i * 512) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (s1) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int i = half * 8 + it;
            t[it] = (i * 8 + lr < rows_left) ? ldg_stream(p1 + i * 512) : make_float4(0.f, 0.f, 0.f, 0.f);
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
            t[it] = (i * 8 + lr < rows_left) ? ldg_stream(p2 + i * 512) : make_float4(0.f, 0.f, 0.f, 0.f);
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

int launch_ff_ts(const float* s0, const float* s1, const float* s2, const float* residual, float* x_out, float* b_out,
                 const uint8_t* image, const float* b1, const float* b2, long long P, int sm_count, cudaStream_t st,
                 const float* head_w, const float* head_b, float* forecast, bool reverse) {
  if (P == 0) return FFNO_OK;
  FFNO_TRY(ensure_dynamic_smem(ff_ts_kernel, FF3_TOTAL));
  const int n_tiles = ceil_div(P, 128);
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  FFNO_CUDA_CHECK(launch_pdl(ff_ts_kernel, dim3(grid), dim3(kFF3Threads), (size_t)FF3_TOTAL, st, s0, s1, s2, residual, x_out,
                             b_out, image, b1, b2, head_w, head_b, forecast, P, n_tiles, reverse ? 1 : 0));
  ++g_launch_counter;
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
