// Generic FP32 kernels of the F-FNO forward (see generic_kernels.cuh).  Hand-written CUDA, sm_100a.
#include "generic_kernels.cuh"

namespace ffno {

thread_local long long g_launch_counter = 0;   // bumped by every launch wrapper (see plan.cu)

// ------------------------------------------------------------------------------------------------
// Truncated real DFT along a strided axis, as a skinny matrix product with a host-built table.
// One thread owns one float4 "column" (o, inner4) and all of its n_out outputs, JC at a time.
// Loads of X are coalesced along `inner`; table reads are warp-uniform (broadcast through L1).
// ------------------------------------------------------------------------------------------------
template <int JC>
__global__ void __launch_bounds__(128)
axis_transform_kernel(const float4* __restrict__ X, const float* __restrict__ T, float4* __restrict__ Y,
                      long long ncols, long long inner4, int n_in, int n_out, int ldt, int accumulate) {
  long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  long long o = col / inner4, i4 = col - o * inner4;
  const float4* xp = X + o * (long long)n_in * inner4 + i4;
  float4* yp = Y + o * (long long)n_out * inner4 + i4;
  for (int j0 = 0; j0 < n_out; j0 += JC) {
    float4 acc[JC];
#pragma unroll
    for (int j = 0; j < JC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = 0; i < n_in; ++i) {
      float4 v = xp[(long long)i * inner4];
      const float4* t4 = reinterpret_cast<const float4*>(T + (long long)i * ldt + j0);
#pragma unroll
      for (int j = 0; j < JC; j += 4) {
        float4 c = __ldg(t4 + j / 4);
        acc[j + 0].x = fmaf(c.x, v.x, acc[j + 0].x); acc[j + 0].y = fmaf(c.x, v.y, acc[j + 0].y);
        acc[j + 0].z = fmaf(c.x, v.z, acc[j + 0].z); acc[j + 0].w = fmaf(c.x, v.w, acc[j + 0].w);
        acc[j + 1].x = fmaf(c.y, v.x, acc[j + 1].x); acc[j + 1].y = fmaf(c.y, v.y, acc[j + 1].y);
        acc[j + 1].z = fmaf(c.y, v.z, acc[j + 1].z); acc[j + 1].w = fmaf(c.y, v.w, acc[j + 1].w);
        acc[j + 2].x = fmaf(c.z, v.x, acc[j + 2].x); acc[j + 2].y = fmaf(c.z, v.y, acc[j + 2].y);
        acc[j + 2].z = fmaf(c.z, v.z, acc[j + 2].z); acc[j + 2].w = fmaf(c.z, v.w, acc[j + 2].w);
        acc[j + 3].x = fmaf(c.w, v.x, acc[j + 3].x); acc[j + 3].y = fmaf(c.w, v.y, acc[j + 3].y);
        acc[j + 3].z = fmaf(c.w, v.z, acc[j + 3].z); acc[j + 3].w = fmaf(c.w, v.w, acc[j + 3].w);
      }
    }
#pragma unroll
    for (int j = 0; j < JC; ++j) {
      if (j0 + j < n_out) {
        float4* dst = yp + (long long)(j0 + j) * inner4;
        float4 r = acc[j];
        if (accumulate) {
          float4 p = *dst;
          r.x += p.x; r.y += p.y; r.z += p.z; r.w += p.w;
        }
        *dst = r;
      }
    }
  }
}

int launch_axis_transform(const float* X, const float* T, float* Y, long long outer, int n_in, int n_out,
                          long long inner, bool accumulate, cudaStream_t st) {
  FFNO_REQUIRE(inner % 4 == 0, FFNO_ERR_UNSUPPORTED, "axis transform: inner=%lld not a multiple of 4", inner);
  constexpr int JC = 16;
  int ldt = (n_out + JC - 1) / JC * JC;   // tables are allocated with this padded leading dimension
  long long inner4 = inner / 4, ncols = outer * inner4;
  if (ncols == 0) return FFNO_OK;
  axis_transform_kernel<JC><<<ceil_div(ncols, 128), 128, 0, st>>>(
      reinterpret_cast<const float4*>(X), T, reinterpret_cast<float4*>(Y), ncols, inner4, n_in, n_out, ldt,
      accumulate ? 1 : 0);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("axis_transform_kernel");
  return FFNO_OK;
}

// ------------------------------------------------------------------------------------------------
// Tiled FP32 SGEMM (64x64 block tile, 4x4 per thread, K chunks of 16) with functor A-loader / epilogue.
// ------------------------------------------------------------------------------------------------
template <class ALoad, class Epi>
__global__ void __launch_bounds__(256)
sgemm64_kernel(ALoad a, const float* __restrict__ Bt, long long b_batch_stride, long long M, int N, int K,
               Epi epi) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int z = blockIdx.z;
  const long long m0 = (long long)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const float* B = Bt + (long long)z * b_batch_stride;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += 16) {
    {
      int r = tid / 4, kk = (tid % 4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m0 + r < M && k0 + kk < K) v = a.load4(z, m0 + r, k0 + kk);
      As[kk + 0][r] = v.x; As[kk + 1][r] = v.y; As[kk + 2][r] = v.z; As[kk + 3][r] = v.w;
    }
    {
      int k = tid / 16, nn = (tid % 16) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + k < K && n0 + nn < N)
        v = __ldg(reinterpret_cast<const float4*>(B + (long long)(k0 + k) * N + n0 + nn));
      *reinterpret_cast<float4*>(&Bs[k][nn]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int col = n0 + tx * 4;
  if (col < N) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      long long row = m0 + ty * 4 + i;
      if (row < M) epi.store4(z, row, col, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
  }
}

// The same product with a 128 x TN block tile (TN = 128 or 64) and 8 x (TN / 16) outputs per thread: half the
// shared-memory loads per FMA of the 64 x 64 kernel.  Used when the shape fills the tile (the FF linears of the training
// path and of the generic forward); N % TN == 0 and K % 16 == 0 are required, rows are bounds-checked.
template <int TN, class ALoad, class Epi>
__global__ void __launch_bounds__(256)
sgemm128_kernel(ALoad a, const float* __restrict__ Bt, long long b_batch_stride, long long M, int N, int K, Epi epi) {
  constexpr int CN = TN / 16;            // output columns per thread: 8 or 4
  __shared__ float As[16][128 + 4];
  __shared__ float Bs[16][TN + 4];
  const int z = blockIdx.z;
  const long long m0 = (long long)blockIdx.x * 128;
  const int n0 = blockIdx.y * TN;
  const float* B = Bt + (long long)z * b_batch_stride;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[8][CN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < CN; ++j) acc[i][j] = 0.f;

  // the tiles of step k0 + 16 are fetched into registers while step k0 is multiplied (one memory latency per step
  // would otherwise sit between the two barriers)
  float4 va[2], vb[TN / 64];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {         // A tile: 128 rows x 16 k
      const int r = tid / 4 + 64 * j, kk = (tid % 4) * 4;
      va[j] = (m0 + r < M) ? a.load4(z, m0 + r, k0 + kk) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < TN / 64; ++j) {   // B tile: 16 k x TN columns
      const int idx = tid + 256 * j, k = idx / (TN / 4), nn = (idx % (TN / 4)) * 4;
      vb[j] = __ldg(reinterpret_cast<const float4*>(B + (long long)(k0 + k) * N + n0 + nn));
    }
  };
  // (TN = 128: the 64 accumulators leave no registers for the look-ahead — 153 registers, one block per SM, measured
  // 117 -> 167 us; that tile fetches at the top of its own step instead)
  constexpr bool kAhead = TN == 64;
  if (kAhead) fetch(0);
  for (int k0 = 0; k0 < K; k0 += 16) {
    if (!kAhead) fetch(k0);
#pragma unroll
    for (int j = 0; j < 2; ++j) {         // transposed into As[k][row]
      const int r = tid / 4 + 64 * j, kk = (tid % 4) * 4;
      As[kk + 0][r] = va[j].x; As[kk + 1][r] = va[j].y; As[kk + 2][r] = va[j].z; As[kk + 3][r] = va[j].w;
    }
#pragma unroll
    for (int j = 0; j < TN / 64; ++j) {
      const int idx = tid + 256 * j, k = idx / (TN / 4), nn = (idx % (TN / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[k][nn]) = vb[j];
    }
    __syncthreads();
    if (kAhead && k0 + 16 < K) fetch(k0 + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float br[CN];
      {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        br[0] = b0.x; br[1] = b0.y; br[2] = b0.z; br[3] = b0.w;
        if (CN == 8) {
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][(TN / 2) + tx * 4]);
          br[CN - 4] = b1.x; br[CN - 3] = b1.y; br[CN - 2] = b1.z; br[CN - 1] = b1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < CN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const long long row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row < M) {
      epi.store4(z, row, n0 + tx * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
      if (CN == 8)
        epi.store4(z, row, n0 + (TN / 2) + tx * 4,
                   make_float4(acc[i][CN - 4], acc[i][CN - 3], acc[i][CN - 2], acc[i][CN - 1]));
    }
  }
}

// ---- mode mix ----------------------------------------------------------------------------------
struct MixGeom {
  const float* F;
  float* R;
  int K, C;
  long long p_inner, inner;   // inner = p_inner * C
};
struct MixALoad {
  MixGeom g;
  __device__ float4 load4(int k, long long row, int kk) const {
    long long o = row / g.p_inner, p = row - o * g.p_inner;
    int ri = kk / g.C, ci = kk - ri * g.C;
    return __ldg(reinterpret_cast<const float4*>(g.F + ((o * g.K + k) * 2 + ri) * g.inner + p * g.C + ci));
  }
};
struct MixEpi {
  MixGeom g;
  __device__ void store4(int k, long long row, int col, float4 v) const {
    long long o = row / g.p_inner, p = row - o * g.p_inner;
    int ro = col / g.C, co = col - ro * g.C;
    *reinterpret_cast<float4*>(g.R + ((o * g.K + k) * 2 + ro) * g.inner + p * g.C + co) = v;
  }
};

int launch_mode_mix(const float* F, const float* Wblk, float* R, long long outer, int K, long long p_inner,
                    int C, cudaStream_t st) {
  FFNO_REQUIRE(C % 4 == 0, FFNO_ERR_UNSUPPORTED, "mode mix: width %d not a multiple of 4", C);
  MixGeom g{F, R, K, C, p_inner, p_inner * C};
  long long M = outer * p_inner;
  if (M == 0 || K == 0) return FFNO_OK;
  dim3 grid(ceil_div(M, 64), ceil_div(2 * C, 64), K);
  sgemm64_kernel<<<grid, 256, 0, st>>>(MixALoad{g}, Wblk, (long long)4 * C * C, M, 2 * C, 2 * C, MixEpi{g});
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("sgemm64_kernel<mix>");
  return FFNO_OK;
}

// ---- linear ------------------------------------------------------------------------------------
struct LinALoad {
  const float* x;
  int K;
  __device__ float4 load4(int, long long row, int kk) const {
    return __ldg(reinterpret_cast<const float4*>(x + row * K + kk));
  }
};
struct LinEpi {
  const float* bias;
  const float* residual;
  float* y;
  float* y_pre;
  int N, relu;
  const float* mask;      // backward of a ReLU: zero the outputs whose mask[row][col] <= 0 (NULL: none)
  __device__ void store4(int, long long row, int col, float4 v) const {
    if (mask) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(mask + row * N + col));
      v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f; v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
    }
    if (bias) {
      float4 b = __ldg(reinterpret_cast<const float4*>(bias + col));
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
    if (y_pre) *reinterpret_cast<float4*>(y_pre + row * N + col) = v;
    if (residual) {
      float4 r = __ldg(reinterpret_cast<const float4*>(residual + row * N + col));
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (y) *reinterpret_cast<float4*>(y + row * N + col) = v;
  }
};

int launch_linear(const float* x, const float* Wt, const float* bias, const float* residual, float* y,
                  float* y_pre, long long P, int K, int N, bool relu, cudaStream_t st, const float* relu_mask) {
  FFNO_REQUIRE(K % 4 == 0 && N % 4 == 0, FFNO_ERR_UNSUPPORTED,
               "linear: in=%d / out=%d must be multiples of 4", K, N);
  if (P == 0) return FFNO_OK;
  const LinEpi epi{bias, residual, y, y_pre, N, relu ? 1 : 0, relu_mask};
  if (P >= 4096 && K % 16 == 0 && N % 128 == 0) {
    sgemm128_kernel<128><<<dim3(ceil_div(P, 128), N / 128, 1), 256, 0, st>>>(LinALoad{x, K}, Wt, 0, P, N, K, epi);
  } else if (P >= 4096 && K % 16 == 0 && N % 64 == 0) {
    sgemm128_kernel<64><<<dim3(ceil_div(P, 128), N / 64, 1), 256, 0, st>>>(LinALoad{x, K}, Wt, 0, P, N, K, epi);
  } else {
    dim3 grid(ceil_div(P, 64), ceil_div(N, 64), 1);
    sgemm64_kernel<<<grid, 256, 0, st>>>(LinALoad{x, K}, Wt, 0, P, N, K, epi);
  }
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("sgemm_kernel<linear>");
  return FFNO_OK;
}

// ---- LayerNorm + residual: one warp per point -----------------------------------------------------
__global__ void __launch_bounds__(256)
layernorm_residual_kernel(const float* __restrict__ y, const float* __restrict__ gamma,
                          const float* __restrict__ beta, const float* __restrict__ residual,
                          float* __restrict__ out, float* __restrict__ b_out, long long P, int C) {
  long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 32;
  int lane = threadIdx.x % 32;
  if (p >= P) return;
  const float* row = y + p * C;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += row[c];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  float mean = s / C;
  float v = 0.f;
  for (int c = lane; c < C; c += 32) { float d = row[c] - mean; v += d * d; }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  float rstd = rsqrtf(v / C + 1e-5f);
  for (int c = lane; c < C; c += 32) {
    float b = (row[c] - mean) * rstd * gamma[c] + beta[c];
    if (b_out) b_out[p * C + c] = b;
    if (out) out[p * C + c] = residual ? residual[p * C + c] + b : b;
  }
}

int launch_layernorm_residual(const float* y, const float* gamma, const float* beta, const float* residual,
                              float* out, float* b_out, long long P, int C, cudaStream_t st) {
  if (P == 0) return FFNO_OK;
  layernorm_residual_kernel<<<ceil_div(P * 32, 256), 256, 0, st>>>(y, gamma, beta, residual, out, b_out, P, C);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("layernorm_residual_kernel");
  return FFNO_OK;
}

// ---- lift / head ---------------------------------------------------------------------------------
struct LiftDev {
  int ndim, size[3], pad[3], in_features, append_grid, C;
  const float* grid_vals[3];   // unused here (kept for ABI symmetry)
};

__device__ __forceinline__ float linspace01(int i, int n) {
  // float32(np.linspace(0, 1, n)[i]): numpy computes i * (1/(n-1)) in float64 and pins the last sample.
  if (n <= 1) return 0.f;
  if (i == n - 1) return 1.f;
  return (float)((double)i * (1.0 / (double)(n - 1)));
}

// IDX = unsigned when the flattened index fits 32 bits (every BASELINE shape): the per-thread coordinate split is a
// handful of 32-bit divisions instead of 64-bit ones, which were most of this kernel's time.
template <typename IDX>
__global__ void __launch_bounds__(256)
lift_kernel(const float* __restrict__ x, const float* __restrict__ Wt, const float* __restrict__ bias,
            float4* __restrict__ out, long long total4, LiftGeom g) {
  const long long idx64 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx64 >= total4) return;
  const IDX idx = (IDX)idx64;
  const IDX C4 = (IDX)(g.C / 4);
  const int c4 = (int)(idx % C4);
  IDX rem = idx / C4;
  int coord[3] = {0, 0, 0};
  bool in_pad = false;
  for (int a = g.ndim - 1; a >= 0; --a) {
    const IDX ext = (IDX)(g.size[a] + g.pad[a]);
    const IDX q = rem / ext;
    coord[a] = (int)(rem - q * ext);
    rem = q;
    in_pad |= coord[a] >= g.size[a];
  }
  const long long b = (long long)rem;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!in_pad) {
    long long src = b;
    for (int a = 0; a < g.ndim; ++a) src = src * g.size[a] + coord[a];
    if (bias) acc = __ldg(reinterpret_cast<const float4*>(bias) + c4);
    const float* xr = x + src * g.in_features;
    for (int j = 0; j < g.in_features; ++j) {
      float v = __ldg(xr + j);
      float4 w = __ldg(reinterpret_cast<const float4*>(Wt + (long long)j * g.C) + c4);
      acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y);
      acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
    }
    if (g.append_grid) {
      for (int a = 0; a < g.ndim; ++a) {
        float v = linspace01(coord[a], g.size[a]);
        float4 w = __ldg(reinterpret_cast<const float4*>(Wt + (long long)(g.in_features + a) * g.C) + c4);
        acc.x = fmaf(v, w.x, acc.x); acc.y = fmaf(v, w.y, acc.y);
        acc.z = fmaf(v, w.z, acc.z); acc.w = fmaf(v, w.w, acc.w);
      }
    }
  }
  out[idx64] = acc;
}

// Periodic grids (no padding, no appended coordinates — FNOFactorized2DBlock, grid_2d.py:157): the lift is a plain
// [points x in] x [in x C] product.  One thread keeps its float4 column of the weights and the bias in registers and
// walks kLiftPts points, so the kernel is a streaming write of the activations instead of a division per float4.
constexpr int kLiftPts = 8, kLiftMaxIn = 8;
__global__ void __launch_bounds__(256)
lift_plain_kernel(const float* __restrict__ x, const float* __restrict__ Wt, const float* __restrict__ bias,
                  float4* __restrict__ out, long long pts, int in_features, int C4) {
  const int c4 = threadIdx.x % C4, sub = threadIdx.x / C4, per_block = blockDim.x / C4;
  float4 w[kLiftMaxIn];
#pragma unroll
  for (int j = 0; j < kLiftMaxIn; ++j)
    w[j] = j < in_features ? __ldg(reinterpret_cast<const float4*>(Wt + (long long)j * C4 * 4) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const long long p0 = ((long long)blockIdx.x * per_block + sub) * kLiftPts;
#pragma unroll
  for (int i = 0; i < kLiftPts; ++i) {
    const long long pt = p0 + i;
    if (pt >= pts) return;
    const float* xr = x + pt * in_features;
    float4 acc = b;
#pragma unroll
    for (int j = 0; j < kLiftMaxIn; ++j)
      if (j < in_features) {
        const float v = __ldg(xr + j);
        acc.x = fmaf(v, w[j].x, acc.x); acc.y = fmaf(v, w[j].y, acc.y);
        acc.z = fmaf(v, w[j].z, acc.z); acc.w = fmaf(v, w[j].w, acc.w);
      }
    out[pt * C4 + c4] = acc;
  }
}

int launch_lift(const float* x, const float* Wt, const float* bias, float* out, int batch, const LiftGeom& g,
                cudaStream_t st) {
  FFNO_REQUIRE(g.C % 4 == 0, FFNO_ERR_UNSUPPORTED, "lift: width %d not a multiple of 4", g.C);
  long long pts = batch;
  for (int a = 0; a < g.ndim; ++a) pts *= (g.size[a] + g.pad[a]);
  long long total4 = pts * (g.C / 4);
  if (total4 == 0) return FFNO_OK;
  bool plain = !g.append_grid && g.in_features <= kLiftMaxIn && 256 % (g.C / 4) == 0;
  for (int a = 0; a < g.ndim; ++a) plain &= g.pad[a] == 0;
  if (plain) {
    const int C4 = g.C / 4, per_block = 256 / C4;
    lift_plain_kernel<<<ceil_div(pts, (long long)per_block * kLiftPts), 256, 0, st>>>(x, Wt, bias, reinterpret_cast<float4*>(out),
                                                                                      pts, g.in_features, C4);
    ++g_launch_counter;
    FFNO_LAUNCH_CHECK("lift_plain_kernel");
    return FFNO_OK;
  }
  if (total4 < (1ll << 31))
    lift_kernel<unsigned><<<ceil_div(total4, 256), 256, 0, st>>>(x, Wt, bias, reinterpret_cast<float4*>(out), total4, g);
  else
    lift_kernel<unsigned long long><<<ceil_div(total4, 256), 256, 0, st>>>(x, Wt, bias, reinterpret_cast<float4*>(out),
                                                                          total4, g);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("lift_kernel");
  return FFNO_OK;
}

template <int MAXO>
__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ b, const float* __restrict__ Weff, const float* __restrict__ beff,
            float* __restrict__ y, long long n_pts, LiftGeom g, int out_features, int accumulate) {
  // 16 lanes per (cropped) point
  long long gp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / 16;
  int sub = threadIdx.x % 16;
  bool active = gp < n_pts;
  long long pidx = active ? gp : 0;
  int coord[3] = {0, 0, 0};
  long long rem = pidx;
  for (int a = g.ndim - 1; a >= 0; --a) { coord[a] = (int)(rem % g.size[a]); rem /= g.size[a]; }
  long long src = rem;
  for (int a = 0; a < g.ndim; ++a) src = src * (g.size[a] + g.pad[a]) + coord[a];
  float acc[MAXO];
#pragma unroll
  for (int j = 0; j < MAXO; ++j) acc[j] = 0.f;
  const int C4 = g.C / 4;
  for (int c4 = sub; c4 < C4; c4 += 16) {
    float4 v = __ldg(reinterpret_cast<const float4*>(b + src * g.C) + c4);
#pragma unroll
    for (int j = 0; j < MAXO; ++j) {
      if (j < out_features) {
        float4 w = __ldg(reinterpret_cast<const float4*>(Weff + (long long)j * g.C) + c4);
        acc[j] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[j]))));
      }
    }
  }
#pragma unroll
  for (int j = 0; j < MAXO; ++j)
    for (int o = 8; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o, 16);
  if (active && sub == 0)
    for (int j = 0; j < out_features && j < MAXO; ++j) {
      float v = acc[j] + beff[j];
      if (accumulate) v += y[pidx * out_features + j];
      y[pidx * out_features + j] = v;
    }
}

int launch_head(const float* b, const float* Weff, const float* beff, float* y, int batch, const LiftGeom& g,
                int out_features, bool accumulate, cudaStream_t st) {
  FFNO_REQUIRE(out_features >= 1 && out_features <= 8, FFNO_ERR_UNSUPPORTED,
               "head: out_features=%d not in [1,8]", out_features);
  long long pts = batch;
  for (int a = 0; a < g.ndim; ++a) pts *= g.size[a];
  if (pts == 0) return FFNO_OK;
  head_kernel<8><<<ceil_div(pts * 16, 256), 256, 0, st>>>(b, Weff, beff, y, pts, g, out_features,
                                                          accumulate ? 1 : 0);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("head_kernel");
  return FFNO_OK;
}

// ---- parameter preparation -------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
weight_fold_transpose_kernel(const float* __restrict__ v, const float* __restrict__ g,
                             float* __restrict__ w_t, int out, int in) {
  int o = blockIdx.x;
  __shared__ float red[4];
  float scale = 1.f;
  if (g) {
    float s = 0.f;
    for (int i = threadIdx.x; i < in; i += blockDim.x) { float t = v[(long long)o * in + i]; s = fmaf(t, t, s); }
    for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    if (threadIdx.x % 32 == 0) red[threadIdx.x / 32] = s;
    __syncthreads();
    float tot = red[0] + red[1] + red[2] + red[3];
    scale = g[o] / sqrtf(tot);
  }
  for (int i = threadIdx.x; i < in; i += blockDim.x) w_t[(long long)i * out + o] = v[(long long)o * in + i] * scale;
}

int launch_weight_fold_transpose(const float* v, const float* g, float* w_t, int out, int in, cudaStream_t st) {
  if (out == 0 || in == 0) return FFNO_OK;
  weight_fold_transpose_kernel<<<out, 128, 0, st>>>(v, g, w_t, out, in);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("weight_fold_transpose_kernel");
  return FFNO_OK;
}

__global__ void __launch_bounds__(256)
pack_mix_weights_kernel(const float* __restrict__ w, float* __restrict__ Wblk, int C, int K) {
  // w[i][o][k][2] -> Wblk[k][ri*C+i][ro*C+o] = [[Wr, Wi], [-Wi, Wr]]
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)K * 4 * C * C;
  if (idx >= total) return;
  int n2 = 2 * C;
  int col = (int)(idx % n2);
  long long t = idx / n2;
  int row = (int)(t % n2);
  int k = (int)(t / n2);
  int ri = row / C, i = row % C, ro = col / C, o = col % C;
  float wr = w[(((long long)i * C + o) * K + k) * 2 + 0];
  float wi = w[(((long long)i * C + o) * K + k) * 2 + 1];
  float val = (ri == ro) ? wr : (ri == 0 ? wi : -wi);
  Wblk[idx] = val;
}

__global__ void __launch_bounds__(256)
pack_mix_weights_dct_kernel(const float* __restrict__ w, float* __restrict__ Wblk, int C, int Kpairs, int Kc) {
  // real weights w[i][o][j] (factorized_cno/mesh_3d.py:33) -> Wblk[k][ri*C+i][ro*C+o] = (ri == ro) ? w[i][o][2k+ri] : 0
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)Kpairs * 4 * C * C;
  if (idx >= total) return;
  int n2 = 2 * C;
  int col = (int)(idx % n2);
  long long t = idx / n2;
  int row = (int)(t % n2);
  int k = (int)(t / n2);
  int ri = row / C, i = row % C, ro = col / C, o = col % C;
  const int j = 2 * k + ri;
  Wblk[idx] = (ri == ro && j < Kc) ? w[((long long)i * C + o) * Kc + j] : 0.f;
}

int launch_pack_mix_weights_dct(const float* w, float* Wblk, int C, int Kpairs, int Kc, cudaStream_t st) {
  long long total = (long long)Kpairs * 4 * C * C;
  if (total == 0) return FFNO_OK;
  pack_mix_weights_dct_kernel<<<ceil_div(total, 256), 256, 0, st>>>(w, Wblk, C, Kpairs, Kc);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("pack_mix_weights_dct_kernel");
  return FFNO_OK;
}

__global__ void __launch_bounds__(256)
c2c_combine_kernel(const float4* __restrict__ Y, float4* __restrict__ out, long long n4, long long row4, int R, int c4, float sgn) {
  // n4 = outer * R * q * 2 * c4 float4 outputs; row4 = q * 2 * c4 float4 per (o, r) row
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4) return;
  const long long within = idx % row4, orow = idx / row4;          // orow = o * R + r
  const long long o = orow / R, r = orow % R;
  const int ri = (int)((within / c4) & 1);
  const long long base = (o * 2 * R + r) * row4;
  const float4 c = Y[base + within];
  const float4 sn = Y[base + (long long)R * row4 + (ri ? within - c4 : within + c4)];     // sine sums of the OTHER part
  const float f = ri ? -sgn : sgn;
  out[idx] = make_float4(fmaf(f, sn.x, c.x), fmaf(f, sn.y, c.y), fmaf(f, sn.z, c.z), fmaf(f, sn.w, c.w));
}

int launch_c2c_combine(const float* Y, float* out, long long outer, int R, long long q, int C, float sgn, cudaStream_t st) {
  FFNO_REQUIRE(C % 4 == 0, FFNO_ERR_UNSUPPORTED, "c2c combine: width %d not a multiple of 4", C);
  const int c4 = C / 4;
  const long long row4 = q * 2 * c4, n4 = outer * R * row4;
  if (n4 == 0) return FFNO_OK;
  c2c_combine_kernel<<<ceil_div(n4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(Y), reinterpret_cast<float4*>(out), n4,
                                                        row4, R, c4, sgn);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("c2c_combine_kernel");
  return FFNO_OK;
}

__global__ void __launch_bounds__(256)
c2c_expand_kernel(const float4* __restrict__ d, float4* __restrict__ Y, long long n4, long long row4, int R, int c4, float sgn) {
  // one thread per float4 of d: writes the cosine-row copy and the sine-row entry of the OTHER part
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n4) return;
  const long long within = idx % row4, orow = idx / row4;
  const long long o = orow / R, r = orow % R;
  const int ri = (int)((within / c4) & 1);
  const long long base = (o * 2 * R + r) * row4;
  const float4 v = d[idx];
  Y[base + within] = v;
  const float f = ri ? -sgn : sgn;      // d_re -> Ys_im * sgn, d_im -> Ys_re * (-sgn)
  Y[base + (long long)R * row4 + (ri ? within - c4 : within + c4)] = make_float4(f * v.x, f * v.y, f * v.z, f * v.w);
}

int launch_c2c_expand(const float* d, float* Y, long long outer, int R, long long q, int C, float sgn, cudaStream_t st) {
  FFNO_REQUIRE(C % 4 == 0, FFNO_ERR_UNSUPPORTED, "c2c expand: width %d not a multiple of 4", C);
  const int c4 = C / 4;
  const long long row4 = q * 2 * c4, n4 = outer * R * row4;
  if (n4 == 0) return FFNO_OK;
  c2c_expand_kernel<<<ceil_div(n4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(d), reinterpret_cast<float4*>(Y), n4, row4,
                                                       R, c4, sgn);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("c2c_expand_kernel");
  return FFNO_OK;
}

int launch_pack_mix_weights(const float* w, float* Wblk, int C, int K, cudaStream_t st) {
  long long total = (long long)K * 4 * C * C;
  if (total == 0) return FFNO_OK;
  pack_mix_weights_kernel<<<ceil_div(total, 256), 256, 0, st>>>(w, Wblk, C, K);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("pack_mix_weights_kernel");
  return FFNO_OK;
}

__global__ void fold_head_kernel(const float* __restrict__ W0t, const float* __restrict__ b0,
                                 const float* __restrict__ W1t, const float* __restrict__ b1,
                                 float* __restrict__ Weff, float* __restrict__ beff, int C, int H, int out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < out * C) {
    int j = idx / C, c = idx % C;
    double s = 0.0;
    for (int h = 0; h < H; ++h) s += (double)W1t[(long long)h * out + j] * (double)W0t[(long long)c * H + h];
    Weff[idx] = (float)s;
  }
  if (idx < out) {
    double s = b1 ? (double)b1[idx] : 0.0;
    if (b0)
      for (int h = 0; h < H; ++h) s += (double)W1t[(long long)h * out + idx] * (double)b0[h];
    beff[idx] = (float)s;
  }
}

int launch_fold_head(const float* W0t, const float* b0, const float* W1t, const float* b1, float* Weff,
                     float* beff, int C, int H, int out, cudaStream_t st) {
  fold_head_kernel<<<ceil_div((long long)out * C, 128), 128, 0, st>>>(W0t, b0, W1t, b1, Weff, beff, C, H, out);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("fold_head_kernel");
  return FFNO_OK;
}

// ---- rollout glue --------------------------------------------------------------------------------
// Velocity features of the torus_kochkov rollout (routines/grid_2d_markov.py:82-93, :206-220, :268-285):
//   w_hat = rfftn(w, 'backward');  psi_hat = -w_hat / lap,  lap = (2 pi i)^2 (kx^2 + ky^2), lap[0,0] = 1
//   q = irfftn(2 pi i ky psi_hat) = psi_y,   v = irfftn(-2 pi i kx psi_hat) = -psi_x
// with kx = fftfreq(X, Lx/X), ky = rfftfreq(Y, Ly/Y).  With g = 1 / (2 pi (kx^2 + ky^2)) (0 at DC, where ky = kx = 0
// kill the product anyway) this is q_hat = i ky g w_hat, v_hat = -i kx g w_hat.  One scalar field per sample — a
// 1/64th of one axis of one layer — so the four separable passes are plain FP32 DFT sums over twiddles kept in shared
// memory (exact for any X, Y); the C2R pass drops Im(DC) / Im(Nyquist) like torch's irfftn.
namespace {

constexpr int kVelMaxN = 1024;

__device__ __forceinline__ void fill_twiddles(float* c, float* s, int N) {       // e^{+2 pi i r / N}, r = 0..N-1
  for (int r = threadIdx.x; r < N; r += blockDim.x) sincospif(2.0f * (float)r / (float)N, &s[r], &c[r]);
  __syncthreads();
}

// A[b][x][my] = sum_y w[b][x][y] e^{-2 pi i my y / Y},  my = 0..Y/2
__global__ void __launch_bounds__(256)
vel_fwd_y_kernel(const float* __restrict__ w, long long stride_b, long long stride_xy, float* __restrict__ Are,
                 float* __restrict__ Aim, long long total, int X, int Y, int Yh) {
  __shared__ float tc[kVelMaxN], ts[kVelMaxN];
  fill_twiddles(tc, ts, Y);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int my = (int)(idx % Yh);
  const long long bx = idx / Yh;
  const long long b = bx / X;
  const int x = (int)(bx - b * X);
  const float* row = w + b * stride_b + (long long)x * Y * stride_xy;
  float re = 0.f, im = 0.f;
  int r = 0;
  for (int y = 0; y < Y; ++y) {
    const float v = row[(long long)y * stride_xy];
    re = fmaf(v, tc[r], re);
    im = fmaf(-v, ts[r], im);
    r += my;
    if (r >= Y) r -= Y;
  }
  Are[idx] = re;
  Aim[idx] = im;
}

// W[b][mx][my] = sum_x A[b][x][my] e^{-2 pi i mx x / X};  Q = i ky g W,  V = -i kx g W
__global__ void __launch_bounds__(256)
vel_fwd_x_mul_kernel(const float* __restrict__ Are, const float* __restrict__ Aim, float* __restrict__ Qre,
                     float* __restrict__ Qim, float* __restrict__ Vre, float* __restrict__ Vim, long long total, int X,
                     int Yh, float Lx, float Ly) {
  __shared__ float tc[kVelMaxN], ts[kVelMaxN];
  fill_twiddles(tc, ts, X);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int my = (int)(idx % Yh);
  const long long bm = idx / Yh;
  const long long b = bm / X;
  const int mx = (int)(bm - b * X);
  const float* are = Are + b * X * Yh + my;
  const float* aim = Aim + b * X * Yh + my;
  float re = 0.f, im = 0.f;
  int r = 0;
  for (int x = 0; x < X; ++x) {
    const float ar = are[(long long)x * Yh], ai = aim[(long long)x * Yh];
    const float c = tc[r], s = ts[r];                 // (ar + i ai)(c - i s)
    re = fmaf(ar, c, fmaf(ai, s, re));
    im = fmaf(ai, c, fmaf(-ar, s, im));
    r += mx;
    if (r >= X) r -= X;
  }
  const int mxs = mx < (X + 1) / 2 ? mx : mx - X;     // fftfreq order: 0 .. ceil(X/2)-1, -floor(X/2) .. -1
  const float kx = (float)mxs / Lx, ky = (float)my / Ly;
  const float k2 = kx * kx + ky * ky;
  const float g = k2 > 0.f ? 1.0f / (6.283185307179586f * k2) : 0.f;
  const float gq = ky * g, gv = -kx * g;              // i g (re + i im) = g (-im + i re)
  Qre[idx] = -gq * im;
  Qim[idx] = gq * re;
  Vre[idx] = -gv * im;
  Vim[idx] = gv * re;
}

// B[b][x][my] = sum_mx F[b][mx][my] e^{+2 pi i mx x / X}  for F in {Q, V}
__global__ void __launch_bounds__(256)
vel_inv_x_kernel(const float* __restrict__ Qre, const float* __restrict__ Qim, const float* __restrict__ Vre,
                 const float* __restrict__ Vim, float* __restrict__ BQre, float* __restrict__ BQim,
                 float* __restrict__ BVre, float* __restrict__ BVim, long long total, int X, int Yh) {
  __shared__ float tc[kVelMaxN], ts[kVelMaxN];
  fill_twiddles(tc, ts, X);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int my = (int)(idx % Yh);
  const long long bx = idx / Yh;
  const long long b = bx / X;
  const int x = (int)(bx - b * X);
  const long long base = b * X * Yh + my;
  float qr = 0.f, qi = 0.f, vr = 0.f, vi = 0.f;
  int r = 0;
  for (int mx = 0; mx < X; ++mx) {
    const long long o = base + (long long)mx * Yh;
    const float c = tc[r], s = ts[r];                 // (a + i b)(c + i s)
    const float a0 = Qre[o], b0 = Qim[o], a1 = Vre[o], b1 = Vim[o];
    qr = fmaf(a0, c, fmaf(-b0, s, qr));
    qi = fmaf(b0, c, fmaf(a0, s, qi));
    vr = fmaf(a1, c, fmaf(-b1, s, vr));
    vi = fmaf(b1, c, fmaf(a1, s, vi));
    r += x;
    if (r >= X) r -= X;
  }
  BQre[idx] = qr; BQim[idx] = qi; BVre[idx] = vr; BVim[idx] = vi;
}

// C2R along y: f[b][x][y] = (1 / XY) sum_my c_my Re(B[b][x][my] e^{+2 pi i my y / Y}),  c = 1 (DC, Nyquist) or 2
__global__ void __launch_bounds__(256)
vel_inv_y_kernel(const float* __restrict__ BQre, const float* __restrict__ BQim, const float* __restrict__ BVre,
                 const float* __restrict__ BVim, float* __restrict__ q, float* __restrict__ v, long long total, int X,
                 int Y, int Yh) {
  __shared__ float tc[kVelMaxN], ts[kVelMaxN];
  fill_twiddles(tc, ts, Y);
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int y = (int)(idx % Y);
  const long long bx = idx / Y;
  const long long base = bx * Yh;
  float aq = 0.f, av = 0.f;
  int r = 0;
  for (int my = 0; my < Yh; ++my) {
    const float cm = (my == 0 || (2 * my == Y)) ? 1.f : 2.f;
    const float c = cm * tc[r], s = cm * ts[r];
    aq = fmaf(BQre[base + my], c, fmaf(-BQim[base + my], s, aq));
    av = fmaf(BVre[base + my], c, fmaf(-BVim[base + my], s, av));
    r += y;
    if (r >= Y) r -= Y;
  }
  const float scale = 1.0f / ((float)X * (float)Y);
  q[idx] = aq * scale;
  v[idx] = av * scale;
}

// ---- shared-memory versions (X <= kVelTileX): one block per grid row for the y passes, one block per (sample, 4
// ky columns) for the fused x passes (forward DFT, multiply, inverse DFT: W, Q, V never leave the SM)
constexpr int kVelTileX = 256;
constexpr int kVelTM = 4;

__global__ void __launch_bounds__(128)
vel_rows_fwd_kernel(const float* __restrict__ w, long long stride_b, long long stride_xy, float* __restrict__ Are,
                    float* __restrict__ Aim, int X, int Y, int Yh) {
  __shared__ float tc[kVelMaxN], ts[kVelMaxN], row[kVelMaxN];
  const long long bx = blockIdx.x;
  const long long b = bx / X;
  const int x = (int)(bx - b * X);
  const float* src = w + b * stride_b + (long long)x * Y * stride_xy;
  for (int y = threadIdx.x; y < Y; y += blockDim.x) row[y] = src[(long long)y * stride_xy];
  fill_twiddles(tc, ts, Y);                            // ends with __syncthreads()
  for (int my = threadIdx.x; my < Yh; my += blockDim.x) {
    float re = 0.f, im = 0.f;
    int r = 0;
    for (int y = 0; y < Y; ++y) {
      re = fmaf(row[y], tc[r], re);
      im = fmaf(-row[y], ts[r], im);
      r += my;
      if (r >= Y) r -= Y;
    }
    Are[bx * Yh + my] = re;
    Aim[bx * Yh + my] = im;
  }
}

__global__ void __launch_bounds__(256)
vel_cols_kernel(const float* __restrict__ Are, const float* __restrict__ Aim, float* __restrict__ BQre,
                float* __restrict__ BQim, float* __restrict__ BVre, float* __restrict__ BVim, int X, int Yh, float Lx,
                float Ly) {
  __shared__ float tc[kVelTileX], ts[kVelTileX];
  __shared__ float a_re[kVelTileX][kVelTM], a_im[kVelTileX][kVelTM];
  __shared__ float q_re[kVelTileX][kVelTM], q_im[kVelTileX][kVelTM], v_re[kVelTileX][kVelTM], v_im[kVelTileX][kVelTM];
  const int my0 = blockIdx.x * kVelTM;
  const long long base = (long long)blockIdx.y * X * Yh;
  for (int i = threadIdx.x; i < X * kVelTM; i += blockDim.x) {
    const int x = i / kVelTM, j = i - x * kVelTM;
    const bool ok = my0 + j < Yh;
    a_re[x][j] = ok ? Are[base + (long long)x * Yh + my0 + j] : 0.f;
    a_im[x][j] = ok ? Aim[base + (long long)x * Yh + my0 + j] : 0.f;
  }
  fill_twiddles(tc, ts, X);
  for (int mx = threadIdx.x; mx < X; mx += blockDim.x) {          // W = forward DFT along x, then the multipliers
    float re[kVelTM], im[kVelTM];
#pragma unroll
    for (int j = 0; j < kVelTM; ++j) re[j] = im[j] = 0.f;
    int r = 0;
    for (int x = 0; x < X; ++x) {
      const float c = tc[r], s = ts[r];                             // (ar + i ai)(c - i s)
#pragma unroll
      for (int j = 0; j < kVelTM; ++j) {
        re[j] = fmaf(a_re[x][j], c, fmaf(a_im[x][j], s, re[j]));
        im[j] = fmaf(a_im[x][j], c, fmaf(-a_re[x][j], s, im[j]));
      }
      r += mx;
      if (r >= X) r -= X;
    }
    const int mxs = mx < (X + 1) / 2 ? mx : mx - X;
    const float kx = (float)mxs / Lx;
#pragma unroll
    for (int j = 0; j < kVelTM; ++j) {
      const float ky = (float)(my0 + j) / Ly;
      const float k2 = kx * kx + ky * ky;
      const float g = k2 > 0.f ? 1.0f / (6.283185307179586f * k2) : 0.f;
      const float gq = ky * g, gv = -kx * g;
      q_re[mx][j] = -gq * im[j];
      q_im[mx][j] = gq * re[j];
      v_re[mx][j] = -gv * im[j];
      v_im[mx][j] = gv * re[j];
    }
  }
  __syncthreads();
  for (int x = threadIdx.x; x < X; x += blockDim.x) {             // inverse DFT along x of Q and V
    float qr[kVelTM], qi[kVelTM], vr[kVelTM], vi[kVelTM];
#pragma unroll
    for (int j = 0; j < kVelTM; ++j) qr[j] = qi[j] = vr[j] = vi[j] = 0.f;
    int r = 0;
    for (int mx = 0; mx < X; ++mx) {
      const float c = tc[r], s = ts[r];                             // (a + i b)(c + i s)
#pragma unroll
      for (int j = 0; j < kVelTM; ++j) {
        qr[j] = fmaf(q_re[mx][j], c, fmaf(-q_im[mx][j], s, qr[j]));
        qi[j] = fmaf(q_im[mx][j], c, fmaf(q_re[mx][j], s, qi[j]));
        vr[j] = fmaf(v_re[mx][j], c, fmaf(-v_im[mx][j], s, vr[j]));
        vi[j] = fmaf(v_im[mx][j], c, fmaf(v_re[mx][j], s, vi[j]));
      }
      r += x;
      if (r >= X) r -= X;
    }
#pragma unroll
    for (int j = 0; j < kVelTM; ++j)
      if (my0 + j < Yh) {
        const long long o = base + (long long)x * Yh + my0 + j;
        BQre[o] = qr[j]; BQim[o] = qi[j]; BVre[o] = vr[j]; BVim[o] = vi[j];
      }
  }
}

__global__ void __launch_bounds__(128)
vel_rows_inv_kernel(const float* __restrict__ BQre, const float* __restrict__ BQim, const float* __restrict__ BVre,
                    const float* __restrict__ BVim, float* __restrict__ q, float* __restrict__ v, int X, int Y, int Yh) {
  __shared__ float tc[kVelMaxN], ts[kVelMaxN];
  __shared__ float bq_re[kVelMaxN / 2 + 1], bq_im[kVelMaxN / 2 + 1], bv_re[kVelMaxN / 2 + 1], bv_im[kVelMaxN / 2 + 1];
  const long long bx = blockIdx.x;
  for (int my = threadIdx.x; my < Yh; my += blockDim.x) {
    const float cm = (my == 0 || (2 * my == Y)) ? 1.f : 2.f;       // Hermitian half: bins >= 1 count twice
    bq_re[my] = cm * BQre[bx * Yh + my];
    bq_im[my] = cm * BQim[bx * Yh + my];
    bv_re[my] = cm * BVre[bx * Yh + my];
    bv_im[my] = cm * BVim[bx * Yh + my];
  }
  fill_twiddles(tc, ts, Y);
  const float scale = 1.0f / ((float)X * (float)Y);
  for (int y = threadIdx.x; y < Y; y += blockDim.x) {
    float aq = 0.f, av = 0.f;
    int r = 0;
    for (int my = 0; my < Yh; ++my) {
      aq = fmaf(bq_re[my], tc[r], fmaf(-bq_im[my], ts[r], aq));
      av = fmaf(bv_re[my], tc[r], fmaf(-bv_im[my], ts[r], av));
      r += y;
      if (r >= Y) r -= Y;
    }
    q[bx * Y + y] = aq * scale;
    v[bx * Y + y] = av * scale;
  }
}

}  // namespace

size_t velocity_scratch_floats(int batch, int X, int Y) { return (size_t)10 * batch * X * (Y / 2 + 1); }

int launch_velocity(const float* w, long long stride_b, long long stride_xy, int batch, int X, int Y, float Lx, float Ly,
                    float* q, float* v, float* scratch, cudaStream_t st) {
  FFNO_REQUIRE(X >= 1 && Y >= 1 && X <= kVelMaxN && Y <= kVelMaxN, FFNO_ERR_UNSUPPORTED,
               "velocity features: grid %dx%d exceeds %d per axis", X, Y, kVelMaxN);
  FFNO_REQUIRE(Lx > 0.f && Ly > 0.f, FFNO_ERR_BAD_ARG, "velocity features: domain lengths must be positive");
  if (batch == 0) return FFNO_OK;
  const int Yh = Y / 2 + 1;
  const long long nh = (long long)batch * X * Yh, nr = (long long)batch * X * Y;
  float* a[10];
  for (int i = 0; i < 10; ++i) a[i] = scratch + (size_t)i * nh;
  if (X <= kVelTileX && batch <= 65535) {   // shared-memory passes: 3 launches, the x-direction spectra stay on chip
    const long long rows = (long long)batch * X;
    FFNO_REQUIRE(rows < (1ll << 31), FFNO_ERR_UNSUPPORTED, "velocity features: too many rows");
    vel_rows_fwd_kernel<<<(unsigned)rows, 128, 0, st>>>(w, stride_b, stride_xy, a[0], a[1], X, Y, Yh);
    FFNO_LAUNCH_CHECK("vel_rows_fwd_kernel");
    vel_cols_kernel<<<dim3((unsigned)ceil_div(Yh, kVelTM), (unsigned)batch), 256, 0, st>>>(a[0], a[1], a[6], a[7], a[8],
                                                                                        a[9], X, Yh, Lx, Ly);
    FFNO_LAUNCH_CHECK("vel_cols_kernel");
    vel_rows_inv_kernel<<<(unsigned)rows, 128, 0, st>>>(a[6], a[7], a[8], a[9], q, v, X, Y, Yh);
    FFNO_LAUNCH_CHECK("vel_rows_inv_kernel");
    g_launch_counter += 3;
    return FFNO_OK;
  }
  vel_fwd_y_kernel<<<ceil_div(nh, 256), 256, 0, st>>>(w, stride_b, stride_xy, a[0], a[1], nh, X, Y, Yh);
  FFNO_LAUNCH_CHECK("vel_fwd_y_kernel");
  vel_fwd_x_mul_kernel<<<ceil_div(nh, 256), 256, 0, st>>>(a[0], a[1], a[2], a[3], a[4], a[5], nh, X, Yh, Lx, Ly);
  FFNO_LAUNCH_CHECK("vel_fwd_x_mul_kernel");
  vel_inv_x_kernel<<<ceil_div(nh, 256), 256, 0, st>>>(a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9], nh, X, Yh);
  FFNO_LAUNCH_CHECK("vel_inv_x_kernel");
  vel_inv_y_kernel<<<ceil_div(nr, 256), 256, 0, st>>>(a[6], a[7], a[8], a[9], q, v, nr, X, Y, Yh);
  FFNO_LAUNCH_CHECK("vel_inv_y_kernel");
  g_launch_counter += 4;
  return FFNO_OK;
}

__device__ __forceinline__ float torch_linspace(int i, int steps, float low, float high) {
  // torch.linspace float32 kernel: symmetric evaluation around the midpoint.
  if (steps <= 1) return low;
  float step = (high - low) / (float)(steps - 1);
  int half = steps / 2;
  return (i < half) ? (low + step * (float)i) : (high - step * (float)(steps - 1 - i));
}

// features of one rollout step, normalised: [w, (q, v,) gx, gy]
template <int NF>
__global__ void __launch_bounds__(256)
rollout_features_kernel(const float* __restrict__ frame, long long stride_b, int stride_xy,
                        const float* __restrict__ q, const float* __restrict__ v, float* __restrict__ feat,
                        long long total, int X, int Y, float low, float high, MeanStd ms) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int y = (int)(idx % Y);
  long long t = idx / Y;
  int x = (int)(t % X);
  long long b = t / X;
  float w = frame[b * stride_b + ((long long)x * Y + y) * stride_xy];
  float gx = torch_linspace(x, X, low, high), gy = torch_linspace(y, Y, low, high);
  float* f = feat + idx * NF;
  int c = 0;
  f[c] = (w - ms.m[c]) / ms.s[c]; ++c;
  if (NF == 5) {
    f[c] = (q[idx] - ms.m[c]) / ms.s[c]; ++c;
    f[c] = (v[idx] - ms.m[c]) / ms.s[c]; ++c;
  }
  f[c] = (gx - ms.m[c]) / ms.s[c]; ++c;
  f[c] = (gy - ms.m[c]) / ms.s[c];
}

int launch_rollout_features(const float* frame, long long frame_stride_b, int frame_stride_xy, const float* q,
                            const float* v, float* feat, int batch, int X, int Y, float low, float high,
                            const MeanStd& mean_std, cudaStream_t st) {
  long long total = (long long)batch * X * Y;
  if (total == 0) return FFNO_OK;
  if (q)
    rollout_features_kernel<5><<<ceil_div(total, 256), 256, 0, st>>>(frame, frame_stride_b, frame_stride_xy, q, v, feat,
                                                                     total, X, Y, low, high, mean_std);
  else
    rollout_features_kernel<3><<<ceil_div(total, 256), 256, 0, st>>>(frame, frame_stride_b, frame_stride_xy, q, v, feat,
                                                                     total, X, Y, low, high, mean_std);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("rollout_features_kernel");
  return FFNO_OK;
}

// features of one rollout step with the optional force / mu channels: [w, (q, v,) gx, gy, (f,) (mu)]
__global__ void __launch_bounds__(256)
rollout_features_ex_kernel(const float* __restrict__ frame, long long stride_b, int stride_xy,
                           const float* __restrict__ q, const float* __restrict__ v, const float* __restrict__ force,
                           int force_steps, int t, const float* __restrict__ mu, float* __restrict__ feat, int nf,
                           long long total, int X, int Y, float low, float high, MeanStd ms) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int y = (int)(idx % Y);
  long long r = idx / Y;
  int x = (int)(r % X);
  long long b = r / X;
  float* f = feat + idx * nf;
  int c = 0;
  f[c] = (frame[b * stride_b + ((long long)x * Y + y) * stride_xy] - ms.m[c]) / ms.s[c]; ++c;
  if (q) {
    f[c] = (q[idx] - ms.m[c]) / ms.s[c]; ++c;
    f[c] = (v[idx] - ms.m[c]) / ms.s[c]; ++c;
  }
  f[c] = (torch_linspace(x, X, low, high) - ms.m[c]) / ms.s[c]; ++c;
  f[c] = (torch_linspace(y, Y, low, high) - ms.m[c]) / ms.s[c]; ++c;
  if (force) { f[c] = (force[idx * force_steps + t] - ms.m[c]) / ms.s[c]; ++c; }
  if (mu) { f[c] = (mu[b] - ms.m[c]) / ms.s[c]; }
}

int launch_rollout_features_ex(const float* frame, long long frame_stride_b, int frame_stride_xy, const float* q,
                               const float* v, const float* force, int force_steps, int t, const float* mu, float* feat,
                               int batch, int X, int Y, float low, float high, const MeanStd& mean_std, cudaStream_t st) {
  long long total = (long long)batch * X * Y;
  if (total == 0) return FFNO_OK;
  const int nf = 3 + (q ? 2 : 0) + (force ? 1 : 0) + (mu ? 1 : 0);
  rollout_features_ex_kernel<<<ceil_div(total, 256), 256, 0, st>>>(frame, frame_stride_b, frame_stride_xy, q, v, force,
                                                                   force_steps, force_steps > 1 ? t : 0, mu, feat, nf, total,
                                                                   X, Y, low, high, mean_std);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("rollout_features_ex_kernel");
  return FFNO_OK;
}

__global__ void __launch_bounds__(256)
rollout_denorm_kernel(const float* __restrict__ fc, float* __restrict__ preds, long long total, int n_steps,
                      int t, MeanStd ms) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  preds[idx * n_steps + t] = fc[idx] * ms.s[0] + ms.m[0];
}

int launch_rollout_denorm(const float* forecast, float* preds, int batch, int XY, int n_steps, int t,
                          const MeanStd& mean_std, cudaStream_t st) {
  long long total = (long long)batch * XY;
  if (total == 0) return FFNO_OK;
  rollout_denorm_kernel<<<ceil_div(total, 256), 256, 0, st>>>(forecast, preds, total, n_steps, t, mean_std);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("rollout_denorm_kernel");
  return FFNO_OK;
}

}  // namespace ffno

namespace ffno {

__global__ void __launch_bounds__(256)
linear_any_kernel(const float* __restrict__ x, const float* __restrict__ Wt, const float* __restrict__ bias,
                  float* __restrict__ y, long long total, int K, int N, int relu) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  long long row = idx / N;
  int col = (int)(idx - row * N);
  float acc = bias ? bias[col] : 0.f;
  const float* xr = x + row * K;
  for (int k = 0; k < K; ++k) acc = fmaf(__ldg(xr + k), __ldg(Wt + (long long)k * N + col), acc);
  y[idx] = relu ? fmaxf(acc, 0.f) : acc;
}

int launch_linear_any(const float* x, const float* Wt, const float* bias, float* y, long long P, int K, int N,
                      bool relu, cudaStream_t st) {
  if (K % 4 == 0 && N % 4 == 0) return launch_linear(x, Wt, bias, nullptr, y, nullptr, P, K, N, relu, st);
  long long total = P * N;
  if (total == 0) return FFNO_OK;
  linear_any_kernel<<<ceil_div(total, 256), 256, 0, st>>>(x, Wt, bias, y, total, K, N, relu ? 1 : 0);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("linear_any_kernel");
  return FFNO_OK;
}

__global__ void __launch_bounds__(256)
rel_l2_kernel(const float* __restrict__ x, long long xsb, long long xsi, const float* __restrict__ y,
              long long ysb, long long ysi, long long n, float* __restrict__ out) {
  const int b = blockIdx.x;
  double d2 = 0.0, y2 = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    float xv = x[b * xsb + i * xsi], yv = y[b * ysb + i * ysi];
    float d = xv - yv;
    d2 += (double)d * d;
    y2 += (double)yv * yv;
  }
  __shared__ double sd[8], sy[8];
  for (int o = 16; o > 0; o >>= 1) { d2 += __shfl_xor_sync(0xffffffffu, d2, o); y2 += __shfl_xor_sync(0xffffffffu, y2, o); }
  if (threadIdx.x % 32 == 0) { sd[threadIdx.x / 32] = d2; sy[threadIdx.x / 32] = y2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, c = 0;
    for (int w = 0; w < 8; ++w) { a += sd[w]; c += sy[w]; }
    out[b] = (float)(sqrt(a) / sqrt(c));
  }
}

int launch_rel_l2(const float* x, long long xsb, long long xsi, const float* y, long long ysb, long long ysi,
                  int batch, long long n, float* out, cudaStream_t st) {
  if (batch == 0) return FFNO_OK;
  rel_l2_kernel<<<batch, 256, 0, st>>>(x, xsb, xsi, y, ysb, ysi, n, out);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("rel_l2_kernel");
  return FFNO_OK;
}

}  // namespace ffno
