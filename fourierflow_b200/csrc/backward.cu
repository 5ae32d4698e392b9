// Backward pass of the F-FNO stack — generic FP32 (FFMA) kernels.
//
// What the reference gets from torch.autograd for its one-step training loss (routines/grid_2d_markov.py:172-193:
// forecast -> LpLoss) is written out here as explicit adjoints: every piece of the forward is a real-linear map
// (truncated DFT tables, per-mode block matrices, linears) or a ReLU, so the backward is the same kernels run with
// transposed tables / weights (generic_kernels.cu) plus the reductions over points below that produce the parameter
// gradients.  Parameter gradients ACCUMULATE (+=) into the caller's buffers, like `.grad` does.
#include "backward.cuh"
#include "generic_kernels.cuh"

namespace ffno {

extern thread_local long long g_launch_counter;

namespace {

// Rows of a plain row-major matrix [rows][ld], `ncols` valid columns.
struct RowLoad {
  const float* x;
  int ld, ncols;
  __device__ float4 load4(int, long long row, int col) const {
    const float* p = x + row * ld + col;
    if (col + 3 < ncols && (ld & 3) == 0) return __ldg(reinterpret_cast<const float4*>(p));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (col < ncols) v.x = __ldg(p);
    if (col + 1 < ncols) v.y = __ldg(p + 1);
    if (col + 2 < ncols) v.z = __ldg(p + 2);
    if (col + 3 < ncols) v.w = __ldg(p + 3);
    return v;
  }
};
// Rows of the spectra of one mode: F[o][k][re|im][p][c] seen as [outer * p_inner][2C] (generic_kernels.cu: MixALoad)
struct ModeRowLoad {
  const float* F;
  int K, C;
  long long p_inner, inner;
  __device__ float4 load4(int k, long long row, int col) const {
    long long o = row / p_inner, p = row - o * p_inner;
    int ri = col / C, ci = col - ri * C;
    return __ldg(reinterpret_cast<const float4*>(F + ((o * K + k) * 2 + ri) * inner + p * C + ci));
  }
};
struct PlainAdd {      // C[m][n] += v
  float* out;
  int N;
  __device__ void add(int, int m, int n, float v) const { atomicAdd(out + (long long)m * N + n, v); }
};
// D[k][(ri,ci)][(ro,co)] = sum_rows F * dR folded into the complex weight gradient (grid_2d.py:65-68):
//   Wblk = [[Wr, Wi], [-Wi, Wr]]  =>  dWr = D[0][0] + D[1][1],  dWi = D[0][1] - D[1][0];   w[ci][co][k][re|im]
struct MixWeightAdd {
  float* dw;
  int C, K;
  int dct_kc;      // > 0: DCT variant, real weights [C][C][dct_kc]; pair k carries coefficients 2k (Re rows) and 2k + 1 (Im rows)
  float* dw_hi;    // rfft2 variant: modes k >= ksplit belong to a second weight tensor of the same [C][C][ksplit][2] layout
  int ksplit;
  __device__ void add(int k, int m, int n, float v) const {
    int ri = m / C, ci = m - ri * C, ro = n / C, co = n - ro * C;
    if (ksplit > 0) {
      float* dst = (k >= ksplit ? dw_hi : dw) + (((long long)ci * C + co) * ksplit + (k >= ksplit ? k - ksplit : k)) * 2;
      if (ri == ro) atomicAdd(dst, v);
      else atomicAdd(dst + 1, ri == 0 ? v : -v);
      return;
    }
    if (dct_kc > 0) {      // only the diagonal blocks of [[W_2k, 0], [0, W_2k+1]] are parameters (pack_mix_weights_dct_kernel)
      const int j = 2 * k + ri;
      if (ri == ro && j < dct_kc) atomicAdd(dw + ((long long)ci * C + co) * dct_kc + j, v);
      return;
    }
    float* dst = dw + (((long long)ci * C + co) * K + k) * 2;
    if (ri == ro) atomicAdd(dst, v);
    else atomicAdd(dst + 1, ri == 0 ? v : -v);
  }
};

// C[z][M][N] += sum_{r in this block's row slice} A(z, r, m) * B(z, r, n): 64 x 64 tile, 4 x 4 per thread, 16 rows per step.
// blockIdx.z = z * splits + split.
template <class LoadA, class LoadB, class Epi>
__global__ void __launch_bounds__(256)
atb_kernel(LoadA a, LoadB b, long long rows, int M, int N, int splits, Epi epi) {
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int z = blockIdx.z / splits, split = blockIdx.z - z * splits;
  const long long per = (rows + splits - 1) / splits;
  const long long r_begin = (long long)split * per, r_end = r_begin + per < rows ? r_begin + per : rows;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int lr = tid / 16, lc = (tid % 16) * 4;      // loader: row lr of the 16-row step, columns lc..lc+3
  for (long long r0 = r_begin; r0 < r_end; r0 += 16) {
    float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
    if (r0 + lr < r_end) {
      if (m0 + lc < M) va = a.load4(z, r0 + lr, m0 + lc);
      if (n0 + lc < N) vb = b.load4(z, r0 + lr, n0 + lc);
    }
    *reinterpret_cast<float4*>(&As[lr][lc]) = va;
    *reinterpret_cast<float4*>(&Bs[lr][lc]) = vb;
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
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) epi.add(z, m, n, acc[i][j]);
    }
}

// The same reduction with a TM x TN block tile (128 x 64, 64 x 128 or 128 x 128) and (TM / 16) x (TN / 16) outputs per
// thread; M % TM == 0 and N % TN == 0 required.
template <int TM, int TN, class LoadA, class LoadB, class Epi>
__global__ void __launch_bounds__(256)
atb_big_kernel(LoadA a, LoadB b, long long rows, int M, int N, int splits, Epi epi) {
  constexpr int CM = TM / 16, CN = TN / 16;
  __shared__ float As[16][TM + 4];
  __shared__ float Bs[16][TN + 4];
  const int z = blockIdx.z / splits, split = blockIdx.z - z * splits;
  const long long per = (rows + splits - 1) / splits;
  const long long r_begin = (long long)split * per, r_end = r_begin + per < rows ? r_begin + per : rows;
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  float acc[CM][CN];
#pragma unroll
  for (int i = 0; i < CM; ++i)
#pragma unroll
    for (int j = 0; j < CN; ++j) acc[i][j] = 0.f;
  // the rows of step r0 + 16 are fetched into registers while step r0 is multiplied
  float4 va[TM / 64], vb[TN / 64];
  auto fetch = [&](long long r0) {
#pragma unroll
    for (int j = 0; j < TM / 64; ++j) {
      const int idx = tid + 256 * j, lr = idx / (TM / 4), lc = (idx % (TM / 4)) * 4;
      va[j] = (r0 + lr < r_end) ? a.load4(z, r0 + lr, m0 + lc) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int j = 0; j < TN / 64; ++j) {
      const int idx = tid + 256 * j, lr = idx / (TN / 4), lc = (idx % (TN / 4)) * 4;
      vb[j] = (r0 + lr < r_end) ? b.load4(z, r0 + lr, n0 + lc) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  if (r_begin < r_end) fetch(r_begin);
  for (long long r0 = r_begin; r0 < r_end; r0 += 16) {
#pragma unroll
    for (int j = 0; j < TM / 64; ++j) {
      const int idx = tid + 256 * j, lr = idx / (TM / 4), lc = (idx % (TM / 4)) * 4;
      *reinterpret_cast<float4*>(&As[lr][lc]) = va[j];
    }
#pragma unroll
    for (int j = 0; j < TN / 64; ++j) {
      const int idx = tid + 256 * j, lr = idx / (TN / 4), lc = (idx % (TN / 4)) * 4;
      *reinterpret_cast<float4*>(&Bs[lr][lc]) = vb[j];
    }
    __syncthreads();
    if (r0 + 16 < r_end) fetch(r0 + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float ar[CM], br[CN];
#pragma unroll
      for (int h = 0; h < CM / 4; ++h) {
        const float4 v = *reinterpret_cast<const float4*>(&As[kk][h * 64 + ty * 4]);
        ar[4 * h] = v.x; ar[4 * h + 1] = v.y; ar[4 * h + 2] = v.z; ar[4 * h + 3] = v.w;
      }
#pragma unroll
      for (int h = 0; h < CN / 4; ++h) {
        const float4 v = *reinterpret_cast<const float4*>(&Bs[kk][h * 64 + tx * 4]);
        br[4 * h] = v.x; br[4 * h + 1] = v.y; br[4 * h + 2] = v.z; br[4 * h + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < CM; ++i)
#pragma unroll
        for (int j = 0; j < CN; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < CM; ++i)
#pragma unroll
    for (int j = 0; j < CN; ++j) {
      const int m = m0 + (i / 4) * 64 + ty * 4 + (i % 4), n = n0 + (j / 4) * 64 + tx * 4 + (j % 4);
      epi.add(z, m, n, acc[i][j]);
    }
}

int pick_splits(long long rows, int tiles, int sm_count) {
  // enough blocks to fill the GPU twice, at least 256 rows per block
  long long want = (2ll * sm_count + tiles - 1) / tiles;
  long long cap = (rows + 255) / 256;
  long long s = want < cap ? want : cap;
  return (int)(s < 1 ? 1 : s);
}

__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, float* __restrict__ out, long long rows, int N, int rows_per_block) {
  // out[n] += sum_r x[r][n]; thread t owns column t % N (N <= 256) and rows t / N, t / N + 256 / N, ...
  const int per = blockDim.x / N;
  const int n = threadIdx.x % N, sub = threadIdx.x / N;
  if (sub >= per) return;
  const long long r0 = (long long)blockIdx.x * rows_per_block;
  const long long r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float s = 0.f;
  for (long long r = r0 + sub; r < r1; r += per) s += x[r * N + n];
  atomicAdd(out + n, s);
}

__global__ void __launch_bounds__(256)
relu_bwd_kernel(float4* __restrict__ dh, const float4* __restrict__ h, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 g = dh[i];
  const float4 a = h[i];
  g.x = a.x > 0.f ? g.x : 0.f;
  g.y = a.y > 0.f ? g.y : 0.f;
  g.z = a.z > 0.f ? g.z : 0.f;
  g.w = a.w > 0.f ? g.w : 0.f;
  dh[i] = g;
}

__global__ void __launch_bounds__(256)
axpy_kernel(float* __restrict__ y, const float* __restrict__ x, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}

// w = g v / ||v||_row (linear.py:49, torch weight_norm dim 0):  dg = <dw, v> / ||v||,
// dv = (g / ||v||) dw - (g <dw, v> / ||v||^3) v.  One block per output row; += into dg, dv.
__global__ void __launch_bounds__(128)
wnorm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v, const float* __restrict__ g,
                 float* __restrict__ dg, float* __restrict__ dv, int in) {
  const int o = blockIdx.x;
  __shared__ float red[2][4];
  float nn = 0.f, dot = 0.f;
  for (int i = threadIdx.x; i < in; i += blockDim.x) {
    const float t = v[(long long)o * in + i];
    nn = fmaf(t, t, nn);
    dot = fmaf(dw[(long long)o * in + i], t, dot);
  }
  for (int k = 16; k > 0; k >>= 1) {
    nn += __shfl_xor_sync(0xffffffffu, nn, k);
    dot += __shfl_xor_sync(0xffffffffu, dot, k);
  }
  if (threadIdx.x % 32 == 0) { red[0][threadIdx.x / 32] = nn; red[1][threadIdx.x / 32] = dot; }
  __syncthreads();
  nn = red[0][0] + red[0][1] + red[0][2] + red[0][3];
  dot = red[1][0] + red[1][1] + red[1][2] + red[1][3];
  const float norm = sqrtf(nn), gg = g[o];
  if (threadIdx.x == 0) dg[o] += dot / norm;
  const float a = gg / norm, b = gg * dot / (norm * nn);
  for (int i = threadIdx.x; i < in; i += blockDim.x)
    dv[(long long)o * in + i] += a * dw[(long long)o * in + i] - b * v[(long long)o * in + i];
}

__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  // dst[z][c][r] = src[z][r][c]
  __shared__ float tile[32][33];
  const long long zoff = (long long)blockIdx.z * rows * cols;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;
  for (int i = ty; i < 32; i += 8)
    if (r0 + i < rows && c0 + tx < cols) tile[i][tx] = src[zoff + (long long)(r0 + i) * cols + c0 + tx];
  __syncthreads();
  for (int i = ty; i < 32; i += 8)
    if (c0 + i < cols && r0 + tx < rows) dst[zoff + (long long)(c0 + i) * rows + r0 + tx] = tile[tx][i];
}

// LpLoss.rel (modules/loss.py:33-46): out[b] = ||x_b - y_b|| / ||y_b||;  dx[b][i] = gout[b] (x - y) / (||x - y|| ||y||)
__global__ void __launch_bounds__(256)
rel_l2_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ gout,
                  float* __restrict__ dx, long long n) {
  const long long b = blockIdx.x;
  const float* xb = x + b * n;
  const float* yb = y + b * n;
  __shared__ float red[2][8];
  float dd = 0.f, yy = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = xb[i] - yb[i];
    dd = fmaf(d, d, dd);
    yy = fmaf(yb[i], yb[i], yy);
  }
  for (int k = 16; k > 0; k >>= 1) {
    dd += __shfl_xor_sync(0xffffffffu, dd, k);
    yy += __shfl_xor_sync(0xffffffffu, yy, k);
  }
  if (threadIdx.x % 32 == 0) { red[0][threadIdx.x / 32] = dd; red[1][threadIdx.x / 32] = yy; }
  __syncthreads();
  dd = yy = 0.f;
  for (int k = 0; k < 8; ++k) { dd += red[0][k]; yy += red[1][k]; }
  const float scale = gout[b] / (sqrtf(dd) * sqrtf(yy));
  for (long long i = threadIdx.x; i < n; i += blockDim.x) dx[b * n + i] = scale * (xb[i] - yb[i]);
}

// Mesh variants (mesh_3d.py:161-166, :173): the lift zero-pads every spatial axis on the high side, the head reads the
// unpadded region.  dir 0: dense[b][coord][C] = padded[b][coord][C];  dir 1: padded = dense inside, 0 in the padding.
__global__ void __launch_bounds__(256)
crop_pad_kernel(float4* __restrict__ dense, float4* __restrict__ padded, long long total4, LiftGeom g, int dir) {
  // one thread per float4 of the PADDED tensor
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total4) return;
  const int C4 = g.C / 4;
  const int c4 = (int)(idx % C4);
  long long rem = idx / C4;
  int coord[3] = {0, 0, 0};
  bool in_pad = false;
  for (int a = g.ndim - 1; a >= 0; --a) {
    const int ext = g.size[a] + g.pad[a];
    coord[a] = (int)(rem % ext);
    rem /= ext;
    in_pad |= coord[a] >= g.size[a];
  }
  if (in_pad) {
    if (dir == 1) padded[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  long long src = rem;
  for (int a = 0; a < g.ndim; ++a) src = src * g.size[a] + coord[a];
  if (dir == 0) dense[src * C4 + c4] = padded[idx];
  else padded[idx] = dense[src * C4 + c4];
}

__device__ __forceinline__ float linspace01_bwd(int i, int n) {      // float32(np.linspace(0, 1, n)[i]), as lift_kernel does
  if (n <= 1) return 0.f;
  if (i == n - 1) return 1.f;
  return (float)((double)i * (1.0 / (double)(n - 1)));
}

// rows of the lift's input as the linear sees them: [x (in_features) | linspace grid coordinate per axis] (mesh_3d.py:178-189)
__global__ void __launch_bounds__(256)
lift_rows_kernel(const float* __restrict__ x, float* __restrict__ rows, long long n_pts, LiftGeom g) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  const int in_total = g.in_features + (g.append_grid ? g.ndim : 0);
  float* r = rows + p * in_total;
  for (int j = 0; j < g.in_features; ++j) r[j] = x[p * g.in_features + j];
  if (g.append_grid) {
    long long rem = p;
    int coord[3] = {0, 0, 0};
    for (int a = g.ndim - 1; a >= 0; --a) { coord[a] = (int)(rem % g.size[a]); rem /= g.size[a]; }
    for (int a = 0; a < g.ndim; ++a) r[g.in_features + a] = linspace01_bwd(coord[a], g.size[a]);
  }
}

__global__ void __launch_bounds__(256)
take_cols_kernel(const float* __restrict__ src, float* __restrict__ dst, long long rows, int ld, int ncols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ncols) return;
  const long long r = i / ncols;
  dst[i] = src[r * ld + (i - r * ncols)];
}

}  // namespace

int launch_crop_pad(float* dense, float* padded, int batch, const LiftGeom& g, bool to_padded, cudaStream_t st) {
  long long total4 = (long long)batch * (g.C / 4);
  for (int a = 0; a < g.ndim; ++a) total4 *= g.size[a] + g.pad[a];
  if (total4 == 0) return FFNO_OK;
  crop_pad_kernel<<<ceil_div(total4, 256), 256, 0, st>>>(reinterpret_cast<float4*>(dense), reinterpret_cast<float4*>(padded),
                                                         total4, g, to_padded ? 1 : 0);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("crop_pad_kernel");
  return FFNO_OK;
}

int launch_lift_rows(const float* x, float* rows, long long n_pts, const LiftGeom& g, cudaStream_t st) {
  if (n_pts == 0) return FFNO_OK;
  lift_rows_kernel<<<ceil_div(n_pts, 256), 256, 0, st>>>(x, rows, n_pts, g);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("lift_rows_kernel");
  return FFNO_OK;
}

int launch_take_cols(const float* src, float* dst, long long rows, int ld, int ncols, cudaStream_t st) {
  if (rows == 0 || ncols == 0) return FFNO_OK;
  take_cols_kernel<<<ceil_div(rows * ncols, 256), 256, 0, st>>>(src, dst, rows, ld, ncols);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("take_cols_kernel");
  return FFNO_OK;
}

int launch_linear_wgrad(const float* dy, const float* x, float* dw, long long P, int out, int in, int sm_count,
                        cudaStream_t st) {
  if (P == 0 || out == 0 || in == 0) return FFNO_OK;
  const RowLoad la{dy, out, out}, lb{x, in, in};
  const PlainAdd epi{dw, in};
  if (P >= 4096 && out % 128 == 0 && in % 64 == 0) {
    const int splits = pick_splits(P, (out / 128) * (in / 64), sm_count);
    atb_big_kernel<128, 64><<<dim3(out / 128, in / 64, splits), 256, 0, st>>>(la, lb, P, out, in, splits, epi);
  } else if (P >= 4096 && out % 64 == 0 && in % 128 == 0) {
    const int splits = pick_splits(P, (out / 64) * (in / 128), sm_count);
    atb_big_kernel<64, 128><<<dim3(out / 64, in / 128, splits), 256, 0, st>>>(la, lb, P, out, in, splits, epi);
  } else {
    const int tiles = ceil_div(out, 64) * ceil_div(in, 64);
    const int splits = pick_splits(P, tiles, sm_count);
    dim3 grid(ceil_div(out, 64), ceil_div(in, 64), splits);
    atb_kernel<<<grid, 256, 0, st>>>(la, lb, P, out, in, splits, epi);
  }
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("atb_kernel<linear>");
  return FFNO_OK;
}

int launch_mix_wgrad(const float* F, const float* dR, float* dw, long long outer, int K, long long p_inner, int C,
                     int sm_count, cudaStream_t st, int dct_kc, float* dw_hi, int ksplit) {
  const long long rows = outer * p_inner;
  if (rows == 0 || K == 0) return FFNO_OK;
  FFNO_REQUIRE(C % 4 == 0, FFNO_ERR_UNSUPPORTED, "mix wgrad: width %d not a multiple of 4", C);
  const int n2 = 2 * C;
  const long long inner = p_inner * C;
  const ModeRowLoad la{F, K, C, p_inner, inner}, lb{dR, K, C, p_inner, inner};
  const MixWeightAdd epi{dw, C, K, dct_kc, dw_hi, ksplit};
  if (n2 % 128 == 0 && rows >= 1024) {
    const int splits = pick_splits(rows, (n2 / 128) * (n2 / 128) * K, sm_count);
    FFNO_REQUIRE((long long)K * splits < 65536, FFNO_ERR_UNSUPPORTED, "mix wgrad: %d modes x %d splits", K, splits);
    atb_big_kernel<128, 128><<<dim3(n2 / 128, n2 / 128, K * splits), 256, 0, st>>>(la, lb, rows, n2, n2, splits, epi);
  } else {
    const int tiles = ceil_div(n2, 64) * ceil_div(n2, 64) * K;
    const int splits = pick_splits(rows, tiles, sm_count);
    FFNO_REQUIRE((long long)K * splits < 65536, FFNO_ERR_UNSUPPORTED, "mix wgrad: %d modes x %d splits", K, splits);
    dim3 grid(ceil_div(n2, 64), ceil_div(n2, 64), K * splits);
    atb_kernel<<<grid, 256, 0, st>>>(la, lb, rows, n2, n2, splits, epi);
  }
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("atb_kernel<mix>");
  return FFNO_OK;
}

int launch_colsum(const float* x, float* out, long long rows, int N, cudaStream_t st) {
  if (rows == 0 || N == 0) return FFNO_OK;
  FFNO_REQUIRE(N <= 256, FFNO_ERR_UNSUPPORTED, "colsum: %d columns", N);
  const int rpb = rows >= (1 << 16) ? 128 : 1024;      // enough blocks to fill the GPU; one atomic per thread and block
  colsum_kernel<<<ceil_div(rows, rpb), 256, 0, st>>>(x, out, rows, N, rpb);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("colsum_kernel");
  return FFNO_OK;
}

int launch_relu_bwd(float* dh, const float* h, long long n, cudaStream_t st) {
  if (n == 0) return FFNO_OK;
  FFNO_REQUIRE(n % 4 == 0, FFNO_ERR_UNSUPPORTED, "relu_bwd: n %% 4");
  relu_bwd_kernel<<<ceil_div(n / 4, 256), 256, 0, st>>>(reinterpret_cast<float4*>(dh), reinterpret_cast<const float4*>(h),
                                                        n / 4);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("relu_bwd_kernel");
  return FFNO_OK;
}

int launch_axpy(float* y, const float* x, long long n, cudaStream_t st) {
  if (n == 0) return FFNO_OK;
  axpy_kernel<<<ceil_div(n, 256), 256, 0, st>>>(y, x, n);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("axpy_kernel");
  return FFNO_OK;
}

int launch_wnorm_bwd(const float* dw, const float* v, const float* g, float* dg, float* dv, int out, int in,
                     cudaStream_t st) {
  if (out == 0 || in == 0) return FFNO_OK;
  wnorm_bwd_kernel<<<out, 128, 0, st>>>(dw, v, g, dg, dv, in);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("wnorm_bwd_kernel");
  return FFNO_OK;
}

int launch_transpose(const float* src, float* dst, int rows, int cols, int batch, cudaStream_t st) {
  if (rows == 0 || cols == 0 || batch == 0) return FFNO_OK;
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), batch);
  transpose_kernel<<<grid, 256, 0, st>>>(src, dst, rows, cols);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("transpose_kernel");
  return FFNO_OK;
}

int launch_rel_l2_bwd(const float* x, const float* y, const float* gout, float* dx, int batch, long long n,
                      cudaStream_t st) {
  if (batch == 0 || n == 0) return FFNO_OK;
  rel_l2_bwd_kernel<<<batch, 256, 0, st>>>(x, y, gout, dx, n);
  ++g_launch_counter;
  FFNO_LAUNCH_CHECK("rel_l2_bwd_kernel");
  return FFNO_OK;
}

}  // namespace ffno
