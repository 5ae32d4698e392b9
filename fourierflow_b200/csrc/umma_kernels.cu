// Constant-operand images of the tcgen05 (UMMA) kernels of the F-FNO layer (width 64 / hidden 256) and the hardware
// self-test of the descriptor conventions.  sm_100a only.
//
// Weights and DFT tables are pre-split into BF16 hi/lo and stored as ready-made shared-memory images (K-major,
// SWIZZLE_128B) that the pipelined kernels in umma_pipelined.cu fetch verbatim with cp.async.bulk.
#include "umma.cuh"
#include "umma_kernels.cuh"

namespace ffno {

extern thread_local long long g_launch_counter;
using namespace umma;


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
// Truncated real DFT along a strided axis (forward: torch.fft.rfft(...)[:K] at grid_2d.py:58,76; inverse:
// torch.fft.irfft of the zero-padded bins at :72,:90) as a tcgen05 product against a host-built table.
//   tile = 2 groups of 64 consecutive `inner` elements (M = 128), K = the transformed axis in chunks of 64,
//   N = npad output rows.  The X tile is already "MN-major" in memory (for a fixed axis index the 64 inner
//   elements are contiguous), so it is split to bf16 and stored without any transpose.
// =======================================================================================================
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
