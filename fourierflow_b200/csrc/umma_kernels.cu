// tcgen05 (UMMA) kernels of the F-FNO layer, width 64 / hidden 256.  sm_100a only.
//
// Common shape of every kernel here: one CTA per SM (persistent over 128-row tiles), 128 threads = one
// warpgroup.  FP32 activations are loaded coalesced from global memory, split into BF16 hi/lo and stored
// as K-major SWIZZLE_128B operand tiles; constant operands (weights) arrive pre-split as a ready-made
// shared-memory image via cp.async.bulk; one elected thread issues tcgen05.mma (3 BF16 passes, FP32
// accumulate in tensor memory); all four warps read the accumulator back with tcgen05.ld (thread = row).
#include "umma.cuh"
#include "umma_kernels.cuh"

namespace ffno {

extern thread_local long long g_launch_counter;
using namespace umma;

namespace {

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
// 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
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

// Split a float4 (4 consecutive K elements) and store the two 8-byte bf16 groups into K-major SW128 tiles.
__device__ __forceinline__ void store_split4(uint8_t* tile_hi, uint8_t* tile_lo, int row, int k, float4 v) {
  uint32_t h0, l0, h1, l1;
  split2(v.x, v.y, h0, l0);
  split2(v.z, v.w, h1, l1);
  const uint32_t off = kmajor_sw128_offset(row, k);
  *reinterpret_cast<uint2*>(tile_hi + off) = make_uint2(h0, h1);
  *reinterpret_cast<uint2*>(tile_lo + off) = make_uint2(l0, l1);
}

// Issue the 3-pass product D (+)= A*B over `ksteps` K steps of 16 for K-major SW128 operands whose K blocks
// (64 elements wide) are `a_kblock_bytes` / `b_kblock_bytes` apart.
__device__ __forceinline__ void issue_3pass(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                            int ksteps, uint32_t a_kblock_bytes, uint32_t b_kblock_bytes, uint32_t idesc,
                                            bool accumulate_first) {
  uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll 1
  for (int pass = 0; pass < 3; ++pass) {
    const uint32_t a = (pass == 2) ? a_lo : a_hi;
    const uint32_t b = (pass == 1) ? b_lo : b_hi;
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
      const int kb = ks >> 2, kk = (ks & 3) * 16;
      umma_bf16_ss(tmem_d, desc_kmajor(a + kb * a_kblock_bytes, kk), desc_kmajor(b + kb * b_kblock_bytes, kk), idesc, acc);
      acc = 1u;
    }
  }
}

}  // namespace

// =======================================================================================================
// weight images
// =======================================================================================================
__global__ void __launch_bounds__(256)
pack_ff_image_kernel(const float* __restrict__ w1t, const float* __restrict__ w2t, uint8_t* __restrict__ image) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;      // one thread per weight element of W1 then W2
  if (idx >= 2 * 64 * 256) return;
  float v;
  uint32_t off;
  if (idx < 64 * 256) {
    // GEMM1 B operand: rows n = hidden unit, K = input channel; 4 chunks of 64 hidden units (8 KB each)
    int h = idx / 64, c = idx % 64;
    v = w1t[c * 256 + h];
    off = (uint32_t)(h / 64) * 8192u + kmajor_sw128_offset(h % 64, c);
  } else {
    // GEMM2 B operand: rows o = output channel, K = hidden unit in 4 K blocks of 64 (8 KB each)
    int t = idx - 64 * 256;
    int o = t / 256, h = t % 256;
    v = w2t[h * 64 + o];
    off = 65536u + (uint32_t)(h / 64) * 8192u + kmajor_sw128_offset(o, h % 64);
  }
  __nv_bfloat16 hi = __float2bfloat16_rn(v);
  __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  *reinterpret_cast<__nv_bfloat16*>(image + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(image + off + 32768u) = lo;
}

int launch_pack_ff_image(const float* w1t, const float* w2t, uint8_t* image, cudaStream_t st) {
  pack_ff_image_kernel<<<ceil_div(2 * 64 * 256, 256), 256, 0, st>>>(w1t, w2t, image);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("pack_ff_image_kernel");
  return FFNO_OK;
}

__global__ void __launch_bounds__(256)
pack_mix_image_kernel(const float* __restrict__ wblk, uint8_t* __restrict__ image, int K) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)K * 128 * 128) return;
  int n = (int)(idx % 128);                 // output index (re|im, channel)  -> B row
  int kk = (int)((idx / 128) % 128);        // input index                    -> K
  int k = (int)(idx / (128 * 128));
  float v = wblk[idx];                      // Wblk[k][kk][n]
  uint32_t off = (uint32_t)(kk / 64) * 16384u + kmajor_sw128_offset(n, kk % 64);
  uint8_t* img = image + (long long)k * kMixImageBytes;
  __nv_bfloat16 hi = __float2bfloat16_rn(v);
  __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  *reinterpret_cast<__nv_bfloat16*>(img + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(img + off + 32768u) = lo;
}

int launch_pack_mix_image(const float* wblk, uint8_t* image, int K, cudaStream_t st) {
  long long total = (long long)K * 128 * 128;
  pack_mix_image_kernel<<<ceil_div(total, 256), 256, 0, st>>>(wblk, image, K);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("pack_mix_image_kernel");
  return FFNO_OK;
}

// =======================================================================================================
// FeedForward + residual  (modules/feedforward.py:21-24 + factorized_fno/grid_2d.py:169)
//   tile = 128 points; hidden processed in 4 chunks of 64:
//     D1 = s_tile * W1[chunk]^T        (M128 N64 K64, 3 passes)   -> +b1, ReLU, split -> A2
//     D2 += A2 * W2[:, chunk]^T        (M128 N64 K64, 3 passes)
//   x_out = residual + D2 + b2
// =======================================================================================================
constexpr int FF_SMEM_W = 131072;
constexpr int FF_SMEM_A1 = FF_SMEM_W;                 // hi 16 KB | lo 16 KB
constexpr int FF_SMEM_A2 = FF_SMEM_A1 + 32768;        // hi 16 KB | lo 16 KB
constexpr int FF_SMEM_BIAS = FF_SMEM_A2 + 32768;      // b1[256] b2[64]
constexpr int FF_SMEM_BAR = FF_SMEM_BIAS + 320 * 4;
constexpr int FF_SMEM_TOTAL = FF_SMEM_BAR + 64;

__global__ void __launch_bounds__(128, 1)
ff_umma_kernel(const float* __restrict__ s, const float* __restrict__ residual, float* __restrict__ x_out,
               float* __restrict__ b_out, const uint8_t* __restrict__ image, const float* __restrict__ b1,
               const float* __restrict__ b2, long long P, int n_tiles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sW1h = smem;
  uint8_t* sW1l = smem + 32768;
  uint8_t* sW2h = smem + 65536;
  uint8_t* sW2l = smem + 98304;
  uint8_t* sA1h = smem + FF_SMEM_A1;
  uint8_t* sA1l = sA1h + 16384;
  uint8_t* sA2h = smem + FF_SMEM_A2;
  uint8_t* sA2l = sA2h + 16384;
  float* sb1 = reinterpret_cast<float*>(smem + FF_SMEM_BIAS);
  float* sb2 = sb1 + 256;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + FF_SMEM_BAR);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  for (int i = tid; i < 256; i += 128) sb1[i] = b1 ? b1[i] : 0.f;
  if (tid < 64) sb2[tid] = b2 ? b2[tid] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_d1 = tmem, tmem_d2 = tmem + 64;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  if (tid == 0) {
    mbar_expect_tx(bar_w, FF_SMEM_W);
    for (int i = 0; i < 4; ++i) bulk_g2s(smem + i * 32768, image + i * 32768, 32768, bar_w);
  }
  mbar_wait(bar_w, 0);

  constexpr uint32_t IDESC = make_idesc_bf16(128, 64, 0, 0);
  uint32_t phase = 0;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = (long long)tile * 128;
    // ---- s tile -> A1 (bf16 hi/lo, K-major SW128) ---------------------------------------------------
    {
      float4 v[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int idx = it * 128 + tid, r = idx >> 4, c4 = idx & 15;
        v[it] = (row0 + r < P) ? ldg_stream(s + (row0 + r) * 64 + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int idx = it * 128 + tid, r = idx >> 4, c4 = idx & 15;
        store_split4(sA1h, sA1l, r, c4 * 4, v[it]);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();

#pragma unroll 1
    for (int j = 0; j < 4; ++j) {
      if (tid == 0) {
        tc_fence_after();
        issue_3pass(tmem_d1, smem_u32(sA1h), smem_u32(sA1l), smem_u32(sW1h) + j * 8192, smem_u32(sW1l) + j * 8192, 4, 0, 0,
                    IDESC, false);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
      // ---- epilogue 1: D1 -> +b1 -> ReLU -> split -> A2 (row = this thread) ---------------------------
      {
        uint32_t v0[32], v1[32];
        tmem_ld32(tmem_d1 + lane_base, v0);
        tmem_ld32(tmem_d1 + lane_base + 32, v1);
        tmem_ld_wait();
        const float* bj = sb1 + j * 64;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = c * 8 + q * 2;
            float a = __uint_as_float(col < 32 ? v0[col] : v1[col - 32]) + bj[col];
            float b = __uint_as_float(col + 1 < 32 ? v0[col + 1] : v1[col + 1 - 32]) + bj[col + 1];
            split2(fmaxf(a, 0.f), fmaxf(b, 0.f), hi[q], lo[q]);
          }
          const uint32_t off = (uint32_t)tid * 128u + (uint32_t)((c ^ (tid & 7)) << 4);
          *reinterpret_cast<uint4*>(sA2h + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(sA2l + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        issue_3pass(tmem_d2, smem_u32(sA2h), smem_u32(sA2l), smem_u32(sW2h) + j * 8192, smem_u32(sW2l) + j * 8192, 4, 0, 0,
                    IDESC, j > 0);
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
    }
    // ---- final epilogue: x_out = residual + D2 + b2 ----------------------------------------------------
    {
      uint32_t v0[32], v1[32];
      tmem_ld32(tmem_d2 + lane_base, v0);
      tmem_ld32(tmem_d2 + lane_base + 32, v1);
      tmem_ld_wait();
      const long long row = row0 + tid;
      if (row < P) {
#pragma unroll
        for (int c4 = 0; c4 < 16; ++c4) {
          float4 b;
          b.x = __uint_as_float(c4 < 8 ? v0[c4 * 4 + 0] : v1[c4 * 4 - 32 + 0]) + sb2[c4 * 4 + 0];
          b.y = __uint_as_float(c4 < 8 ? v0[c4 * 4 + 1] : v1[c4 * 4 - 32 + 1]) + sb2[c4 * 4 + 1];
          b.z = __uint_as_float(c4 < 8 ? v0[c4 * 4 + 2] : v1[c4 * 4 - 32 + 2]) + sb2[c4 * 4 + 2];
          b.w = __uint_as_float(c4 < 8 ? v0[c4 * 4 + 3] : v1[c4 * 4 - 32 + 3]) + sb2[c4 * 4 + 3];
          if (b_out) *reinterpret_cast<float4*>(b_out + row * 64 + c4 * 4) = b;
          if (x_out) {
            float4 o = b;
            if (residual) {
              float4 r = *reinterpret_cast<const float4*>(residual + row * 64 + c4 * 4);
              o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
            }
            *reinterpret_cast<float4*>(x_out + row * 64 + c4 * 4) = o;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int launch_ff_umma(const float* s, const float* residual, float* x_out, float* b_out, const uint8_t* image,
                   const float* b1, const float* b2, long long P, int sm_count, cudaStream_t st) {
  if (P == 0) return FFNO_OK;
  static bool configured = false;
  if (!configured) {
    FFNO_CUDA_CHECK(cudaFuncSetAttribute(ff_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FF_SMEM_TOTAL));
    configured = true;
  }
  const int n_tiles = ceil_div(P, 128);
  const int grid = n_tiles < sm_count ? n_tiles : sm_count;
  ff_umma_kernel<<<grid, 128, FF_SMEM_TOTAL, st>>>(s, residual, x_out, b_out, image, b1, b2, P, n_tiles);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("ff_umma_kernel");
  return FFNO_OK;
}

// =======================================================================================================
// Per-mode complex channel mix (grid_2d.py:65-68) as a real [pts x 128] * [128 x 128] GEMM per mode.
//   F/R layout per axis: [outer][k][re|im][p_inner][64]
// =======================================================================================================
constexpr int MIX_SMEM_B = 0;                          // hi 32 KB | lo 32 KB  (2 K blocks of 128 rows x 128 B)
constexpr int MIX_SMEM_A = 65536;                      // hi 32 KB | lo 32 KB
constexpr int MIX_SMEM_BAR = 131072;
constexpr int MIX_SMEM_TOTAL = MIX_SMEM_BAR + 64;
constexpr int MIX_MAX_AXES = 3;

struct MixParams {
  MixAxis ax[MIX_MAX_AXES];
  int tiles_per_cta[MIX_MAX_AXES];
};

__global__ void __launch_bounds__(128, 1) mix_umma_kernel(MixParams prm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const MixAxis& ax = prm.ax[blockIdx.z];
  const int k = blockIdx.y;
  if (k >= ax.K) return;
  const long long M = ax.outer * ax.p_inner;
  const int n_tiles = (int)((M + 127) / 128);
  const int tpc = prm.tiles_per_cta[blockIdx.z];
  const int tile_begin = blockIdx.x * tpc;
  if (tile_begin >= n_tiles) return;
  const int tile_end = min(n_tiles, tile_begin + tpc);

  uint8_t* sBh = smem + MIX_SMEM_B;
  uint8_t* sBl = sBh + 32768;
  uint8_t* sAh = smem + MIX_SMEM_A;
  uint8_t* sAl = sAh + 32768;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + MIX_SMEM_BAR);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (tid == 0) {
    mbar_expect_tx(bar_w, kMixImageBytes);
    const uint8_t* img = ax.image + (long long)k * kMixImageBytes;
    bulk_g2s(sBh, img, 32768, bar_w);
    bulk_g2s(sBl, img + 32768, 32768, bar_w);
  }
  constexpr uint32_t IDESC = make_idesc_bf16(128, 128, 0, 0);
  const long long inner = ax.p_inner * 64;
  uint32_t phase = 0;
  bool w_ready = false;

  for (int tile = tile_begin; tile < tile_end; ++tile) {
    const long long row0 = (long long)tile * 128;
    // ---- A tile: 128 points x (re 64 | im 64) -> bf16 hi/lo, 2 K blocks ---------------------------------
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {              // K block = re (0) / im (1) segment
      float4 v[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int idx = it * 128 + tid, r = idx >> 4, c4 = idx & 15;
        const long long row = row0 + r;
        if (row < M) {
          const long long o = row / ax.p_inner, p = row - o * ax.p_inner;
          v[it] = ldg_stream(ax.F + ((o * ax.K + k) * 2 + half) * inner + p * 64 + c4 * 4);
        } else {
          v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int idx = it * 128 + tid, r = idx >> 4, c4 = idx & 15;
        store_split4(sAh + half * 16384, sAl + half * 16384, r, c4 * 4, v[it]);
      }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (!w_ready) {
      mbar_wait(bar_w, 0);
      w_ready = true;
    }
    if (tid == 0) {
      tc_fence_after();
      issue_3pass(tmem, smem_u32(sAh), smem_u32(sAl), smem_u32(sBh), smem_u32(sBl), 8, 16384, 16384, IDESC, false);
      umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue: D[128 x 128] -> R rows (re segment | im segment) ------------------------------------
    {
      const long long row = row0 + tid;
      long long o = 0, p = 0;
      if (row < M) { o = row / ax.p_inner; p = row - o * ax.p_inner; }
#pragma unroll 1
      for (int q = 0; q < 4; ++q) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + q * 32, v);
        tmem_ld_wait();
        if (row < M) {
          float* dst = ax.R + ((o * ax.K + k) * 2 + (q >> 1)) * inner + p * 64 + (q & 1) * 32;
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4)
            *reinterpret_cast<float4*>(dst + c4 * 4) =
                make_float4(__uint_as_float(v[c4 * 4]), __uint_as_float(v[c4 * 4 + 1]), __uint_as_float(v[c4 * 4 + 2]),
                            __uint_as_float(v[c4 * 4 + 3]));
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 128);
}

int launch_mix_umma(const MixAxis* axes, int n_axes, int sm_count, cudaStream_t st) {
  FFNO_REQUIRE(n_axes >= 1 && n_axes <= MIX_MAX_AXES, FFNO_ERR_BAD_ARG, "mix: n_axes=%d", n_axes);
  static bool configured = false;
  if (!configured) {
    FFNO_CUDA_CHECK(cudaFuncSetAttribute(mix_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MIX_SMEM_TOTAL));
    configured = true;
  }
  MixParams prm;
  int maxK = 0, total_modes = 0;
  long long total_tiles = 0;
  for (int a = 0; a < n_axes; ++a) {
    prm.ax[a] = axes[a];
    maxK = axes[a].K > maxK ? axes[a].K : maxK;
    total_modes += axes[a].K;
    total_tiles += (long long)axes[a].K * ((axes[a].outer * axes[a].p_inner + 127) / 128);
  }
  if (total_tiles == 0) return FFNO_OK;
  // aim at ~one CTA per SM: every CTA keeps one mode's weights resident and walks `tpc` row tiles
  int grid_x = 1;
  for (int a = 0; a < n_axes; ++a) {
    const long long tiles = (axes[a].outer * axes[a].p_inner + 127) / 128;
    long long ctas_per_mode = (long long)sm_count / (total_modes > 0 ? total_modes : 1);
    if (ctas_per_mode < 1) ctas_per_mode = 1;
    long long tpc = (tiles + ctas_per_mode - 1) / ctas_per_mode;
    if (tpc < 1) tpc = 1;
    prm.tiles_per_cta[a] = (int)tpc;
    const int gx = (int)((tiles + tpc - 1) / tpc);
    grid_x = gx > grid_x ? gx : grid_x;
  }
  dim3 grid(grid_x, maxK, n_axes);
  mix_umma_kernel<<<grid, 128, MIX_SMEM_TOTAL, st>>>(prm);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("mix_umma_kernel");
  return FFNO_OK;
}

// =======================================================================================================
// Truncated real DFT along a strided axis (forward: torch.fft.rfft(...)[:K] at grid_2d.py:58,76; inverse:
// torch.fft.irfft of the zero-padded bins at :72,:90) as a tcgen05 product against a host-built table.
//   tile = 2 groups of 64 consecutive `inner` elements (M = 128), K = the transformed axis in chunks of 64,
//   N = npad output rows.  The X tile is already "MN-major" in memory (for a fixed axis index the 64 inner
//   elements are contiguous), so it is split to bf16 and stored without any transpose.
// =======================================================================================================
constexpr int AX_SMEM_A = 0;                       // hi 16 KB | lo 16 KB : [mb 2][kg 8][8 x 128 B]
constexpr int AX_SMEM_BAR = 32768;
constexpr int AX_SMEM_B = 32768 + 1024;            // table image (1024-aligned)

size_t table_image_bytes(int n_in, int n_out) {
  const int npad = (n_out + 15) / 16 * 16, kchunks = (n_in + 63) / 64;
  return (size_t)2 * kchunks * npad * 128;
}

__global__ void __launch_bounds__(256)
pack_table_image_kernel(const float* __restrict__ T, int ldt, int n_in, int n_out, int npad, int kchunks,
                        uint8_t* __restrict__ image) {
  const int total = kchunks * 64 * npad;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = idx % npad, i = idx / npad;
  const float v = (i < n_in && j < n_out) ? T[(long long)i * ldt + j] : 0.f;
  const uint32_t half = (uint32_t)kchunks * npad * 128u;
  const uint32_t off = (uint32_t)(i / 64) * ((uint32_t)npad * 128u) + kmajor_sw128_offset(j, i % 64);
  __nv_bfloat16 hi = __float2bfloat16_rn(v);
  __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  *reinterpret_cast<__nv_bfloat16*>(image + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(image + half + off) = lo;
}

int launch_pack_table_image(const float* T, int ldt, int n_in, int n_out, uint8_t* image, cudaStream_t st) {
  const int npad = (n_out + 15) / 16 * 16, kchunks = (n_in + 63) / 64;
  const int total = kchunks * 64 * npad;
  pack_table_image_kernel<<<ceil_div(total, 256), 256, 0, st>>>(T, ldt, n_in, n_out, npad, kchunks, image);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("pack_table_image_kernel");
  return FFNO_OK;
}

__global__ void __launch_bounds__(128, 2) axis_umma_kernel(AxisXform p, int n_tiles, int tmem_cols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sAh = smem + AX_SMEM_A;
  uint8_t* sAl = sAh + 16384;
  uint64_t* bar_w = reinterpret_cast<uint64_t*>(smem + AX_SMEM_BAR);
  uint64_t* bar_mma = bar_w + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 2);
  uint8_t* sB = smem + AX_SMEM_B;
  const uint32_t b_half = (uint32_t)p.kchunks * p.npad * 128u;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_mma, 1);
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
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  if (tid == 0) {
    const uint32_t total = 2u * b_half;
    mbar_expect_tx(bar_w, total);
    for (uint32_t off = 0; off < total; off += 32768u) {
      const uint32_t n = (total - off) < 32768u ? (total - off) : 32768u;
      bulk_g2s(sB + off, p.table + off, n, bar_w);
    }
  }
  const uint32_t idesc = make_idesc_bf16(128, p.npad, /*a_mn=*/1, /*b_mn=*/0);
  const long long gpi = p.inner >> 6;                 // 64-element groups per outer index
  const long long n_groups = p.outer * gpi;
  uint32_t phase = 0;
  bool w_ready = false;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long G0 = (long long)tile * 2;
#pragma unroll 1
    for (int kc = 0; kc < p.kchunks; ++kc) {
      // ---- X chunk (64 axis indices x 128 inner elements) -> bf16 hi/lo, MN-major SW128 ----------------
      {
        float4 v[16];
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int idx = it * 128 + tid, il = idx >> 5, gsel = (idx >> 4) & 1, c4 = idx & 15;
          const long long G = G0 + gsel;
          const int i = kc * 64 + il;
          if (G < n_groups && i < p.n_in) {
            const long long o = G / gpi, g = G - o * gpi;
            v[it] = ldg_stream(p.X + ((o * p.n_in + i) * p.inner + g * 64 + c4 * 4));
          } else {
            v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int it = 0; it < 16; ++it) {
          const int idx = it * 128 + tid, il = idx >> 5, gsel = (idx >> 4) & 1, c4 = idx & 15;
          uint32_t h0, l0, h1, l1;
          split2(v[it].x, v[it].y, h0, l0);
          split2(v[it].z, v[it].w, h1, l1);
          const uint32_t off = (uint32_t)gsel * 8192u + (uint32_t)(il >> 3) * 1024u + (uint32_t)(il & 7) * 128u +
                               (uint32_t)(((c4 >> 1) ^ (il & 7)) << 4) + (uint32_t)(c4 & 1) * 8u;
          *reinterpret_cast<uint2*>(sAh + off) = make_uint2(h0, h1);
          *reinterpret_cast<uint2*>(sAl + off) = make_uint2(l0, l1);
        }
      }
      fence_proxy_async_smem();
      __syncthreads();
      if (!w_ready) {
        mbar_wait(bar_w, 0);
        w_ready = true;
      }
      if (tid == 0) {
        tc_fence_after();
        const int rem = p.n_in - kc * 64;
        const int ksteps = rem >= 64 ? 4 : (rem + 15) / 16;
        const uint32_t b_blk = smem_u32(sB) + (uint32_t)kc * ((uint32_t)p.npad * 128u);
        uint32_t acc = kc > 0 ? 1u : 0u;
#pragma unroll 1
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t a = smem_u32(pass == 2 ? sAl : sAh);
          const uint32_t b = b_blk + (pass == 1 ? b_half : 0u);
#pragma unroll 1
          for (int ks = 0; ks < ksteps; ++ks) {
            // MN-major A: 64-wide mn blocks 8 KB apart (LBO), 8-row k groups 1 KB apart (SBO); one K step = 2 groups
            const uint64_t ad = make_smem_desc_sw128(a + (uint32_t)ks * 2048u, 8192u, 1024u);
            umma_bf16_ss(tmem, ad, desc_kmajor(b, ks * 16), idesc, acc);
            acc = 1u;
          }
        }
        umma_commit(bar_mma);
      }
      mbar_wait(bar_mma, phase);
      phase ^= 1;
      tc_fence_after();
    }
    // ---- epilogue: D[128 inner elements x n_out] -> Y[o][j][inner] (coalesced along inner) ---------------
    {
      const long long G = G0 + (tid >> 6);
      const bool live = G < n_groups;
      long long o = 0, g = 0;
      if (live) { o = G / gpi; g = G - o * gpi; }
      float* ybase = p.Y + (o * p.n_out) * p.inner + g * 64 + (tid & 63);
#pragma unroll 1
      for (int c0 = 0; c0 < p.npad; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + lane_base + c0, v);
        tmem_ld_wait();
        if (live) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (c0 + j < p.n_out) {
              float* dst = ybase + (long long)(c0 + j) * p.inner;
              float r = __uint_as_float(v[j]);
              if (p.accumulate) r += *dst;
              *dst = r;
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, (uint32_t)tmem_cols);
}

int launch_axis_umma(const AxisXform& p, int sm_count, cudaStream_t st) {
  FFNO_REQUIRE(p.inner % 64 == 0, FFNO_ERR_UNSUPPORTED, "axis_umma: inner=%lld not a multiple of 64", p.inner);
  FFNO_REQUIRE(p.npad >= 16 && p.npad <= 256 && p.npad % 16 == 0, FFNO_ERR_UNSUPPORTED, "axis_umma: npad=%d", p.npad);
  const size_t img = table_image_bytes(p.n_in, p.n_out);
  const size_t smem = AX_SMEM_B + img;
  FFNO_REQUIRE(smem <= 227 * 1024, FFNO_ERR_UNSUPPORTED, "axis_umma: table image %zu B does not fit in shared memory", img);
  const long long n_groups = p.outer * (p.inner / 64);
  if (n_groups == 0) return FFNO_OK;
  static size_t configured = 0;
  if (smem > configured) {
    FFNO_CUDA_CHECK(cudaFuncSetAttribute(axis_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  int tmem_cols = 32;
  while (tmem_cols < p.npad) tmem_cols *= 2;
  const int n_tiles = (int)((n_groups + 1) / 2);
  int ctas_per_sm = (int)((227 * 1024) / (smem + 1024));
  if (ctas_per_sm > 2) ctas_per_sm = 2;
  if (ctas_per_sm * tmem_cols > 512) ctas_per_sm = 512 / tmem_cols;
  if (ctas_per_sm < 1) ctas_per_sm = 1;
  const int max_grid = sm_count * ctas_per_sm;
  const int grid = n_tiles < max_grid ? n_tiles : max_grid;
  axis_umma_kernel<<<grid, 128, smem, st>>>(p, n_tiles, tmem_cols);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("axis_umma_kernel");
  return FFNO_OK;
}

// =======================================================================================================
// Self-test of the descriptor / layout conventions (one CTA).
// =======================================================================================================
__device__ __forceinline__ uint32_t mnmajor_sw128_offset(int mn, int k, int k_total) {
  // atom = 8 k rows x 64 mn (128 B per k row); k groups 1024 B apart; 64-wide mn blocks after all k groups
  const int mb = mn >> 6, m = mn & 63, kg = k >> 3, kr = k & 7;
  const int chunk = (m >> 3) ^ kr;
  return (uint32_t)mb * (uint32_t)(k_total / 8) * 1024u + (uint32_t)kg * 1024u + (uint32_t)kr * 128u +
         (uint32_t)chunk * 16u + (uint32_t)(m & 7) * 2u;
}

__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(const uint16_t* __restrict__ A, const uint16_t* __restrict__ B, float* __restrict__ D, int N, int K,
                     int a_mn, int b_mn, int variant) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sA = smem;                    // up to 128 x 256 bf16 = 64 KB
  uint8_t* sB = smem + 65536;            // up to 256 x 256 bf16 = 128 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 196608);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  for (int i = tid; i < 128 * K; i += 128) {
    int r = i / K, k = i % K;
    uint32_t off = a_mn ? mnmajor_sw128_offset(r, k, K) : (uint32_t)(k / 64) * (128u * 128u) + kmajor_sw128_offset(r, k % 64);
    *reinterpret_cast<uint16_t*>(sA + off) = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    uint32_t off = b_mn ? mnmajor_sw128_offset(r, k, K) : (uint32_t)(k / 64) * ((uint32_t)N * 128u) + kmajor_sw128_offset(r, k % 64);
    *reinterpret_cast<uint16_t*>(sB + off) = B[i];
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool a_tmem = (variant & 2) != 0;          // A operand from tensor memory (K = 64): columns [256-N.. hmm
  if (a_tmem) {
    // row = this thread; pack (k, k+1) into one 32-bit column, low half = even k; A lives at columns 256.. of a
    // 512-column allocation is not available here (256 allocated): use columns N..N+31 (N <= 192 in this mode)
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = (uint32_t)A[tid * K + 2 * c] | ((uint32_t)A[tid * K + 2 * c + 1] << 16);
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)N, v);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, a_mn, b_mn);
    const uint32_t mn_lbo = (uint32_t)(K / 8) * 1024u, mn_sbo = 1024u;
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad, bd;
      if (a_mn) ad = (variant & 1) ? make_smem_desc_sw128(smem_u32(sA) + ks * 2048, mn_sbo, mn_lbo)
                                   : make_smem_desc_sw128(smem_u32(sA) + ks * 2048, mn_lbo, mn_sbo);
      else ad = desc_kmajor(smem_u32(sA) + (ks >> 2) * (128 * 128), (ks & 3) * 16);
      if (b_mn) bd = (variant & 1) ? make_smem_desc_sw128(smem_u32(sB) + ks * 2048, mn_sbo, mn_lbo)
                                   : make_smem_desc_sw128(smem_u32(sB) + ks * 2048, mn_lbo, mn_sbo);
      else bd = desc_kmajor(smem_u32(sB) + (ks >> 2) * (N * 128), (ks & 3) * 16);
      if (a_tmem) umma_bf16_ts(tmem, tmem + (uint32_t)N + (uint32_t)ks * 8u, bd, idesc, ks > 0 ? 1u : 0u);
      else umma_bf16_ss(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32 && c0 + j < N; ++j) D[(long long)tid * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int launch_umma_selftest(const uint16_t* A, const uint16_t* B, float* D, int N, int K, int a_mn, int b_mn, int variant,
                         cudaStream_t st) {
  FFNO_REQUIRE(N % 16 == 0 && N >= 16 && N <= 256 && K % 16 == 0 && K >= 16 && K <= 256, FFNO_ERR_BAD_ARG,
               "selftest: N=%d K=%d", N, K);
  FFNO_REQUIRE((!a_mn && !b_mn) || K % 8 == 0, FFNO_ERR_BAD_ARG, "selftest: K");
  FFNO_REQUIRE(!b_mn || N % 64 == 0, FFNO_ERR_BAD_ARG, "selftest: MN-major B needs N %% 64 == 0");
  FFNO_REQUIRE(!(variant & 2) || (K == 64 && N <= 192 && !a_mn), FFNO_ERR_BAD_ARG, "selftest: TMEM-A mode needs K=64, N<=192");
  const int smem = 196608 + 64;
  FFNO_CUDA_CHECK(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 128, smem, st>>>(A, B, D, N, K, a_mn, b_mn, variant);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("umma_selftest_kernel");
  return FFNO_OK;
}

}  // namespace ffno
