// Generic FP32 (FFMA) kernels of the F-FNO forward: any width (multiple of 4), any axis length, any
// mode count, 2-D/3-D, LayerNorm/fork options.  They are the always-available CUDA path and the
// on-device cross-check for the tcgen05 kernels (umma_*.cu), which take over when the shape qualifies.
#pragma once
#include "common.cuh"

namespace ffno {

// Y[o][j][inner] (+)= sum_i T[i][j] * X[o][i][inner]      (truncated real DFT along a strided axis)
//   forward : i = l (axis length L), j = k' (2K rows: Re,Im interleaved), T = D^T
//   inverse : i = k',                j = l,                               T = E^T
// Reference: torch.fft.rfft / irfft calls at modules/factorized_fno/grid_2d.py:58,72,76,90.
int launch_axis_transform(const float* X, const float* T, float* Y, long long outer, int n_in,
                          int n_out, long long inner, bool accumulate, cudaStream_t st);

// Per-mode complex channel mix as a real block GEMM (grid_2d.py:65-68 "bixy,ioy->boxy"):
//   R[o][k][ro][p][co] = sum_{ri,ci} F[o][k][ri][p][ci] * W[k][ri*C+ci][ro*C+co]
// with inner = P_in * C (P_in = points per line after the transformed axis).
int launch_mode_mix(const float* F, const float* Wblk, float* R, long long outer, int K,
                    long long p_inner, int C, cudaStream_t st);

// y[P][N] = act(x[P][K] @ Wt[K][N] + bias) (+ residual[P][N]);  optional second store y2 (pre-residual)
// relu_mask (backward of a ReLU): outputs whose relu_mask[p][n] <= 0 are zeroed
int launch_linear(const float* x, const float* Wt, const float* bias, const float* residual, float* y,
                  float* y_pre, long long P, int K, int N, bool relu, cudaStream_t st, const float* relu_mask = nullptr);

// b = LayerNorm(y) * g + beta over the last dim (C); out = residual + b (residual may be NULL);
// b_out optional.  Reference: modules/feedforward.py:17.
int launch_layernorm_residual(const float* y, const float* gamma, const float* beta,
                              const float* residual, float* out, float* b_out, long long P, int C,
                              cudaStream_t st);

// Lift: optional linspace grid append (mesh_3d.py:178-189), Linear(in->C) (grid_2d.py:157), zero pad on
// the high side of every spatial axis (mesh_3d.py:164-166).  out: [B, size+pad.., C]
struct LiftGeom {
  int ndim;
  int size[3];
  int pad[3];
  int in_features;   // of the given tensor
  int append_grid;
  int C;
};
int launch_lift(const float* x, const float* Wt /*[in_total][C]*/, const float* bias, float* out,
                int batch, const LiftGeom& g, cudaStream_t st);

// Head on the cropped last backcast (grid_2d.py:171-172; mesh_3d.py:173-174) with the two activation-free
// linears pre-multiplied: y[p][j] = sum_c b[p][c] * Weff[j][c] + beff[j].
int launch_head(const float* b, const float* Weff /*[out][C]*/, const float* beff, float* y, int batch,
                const LiftGeom& g, int out_features, bool accumulate, cudaStream_t st);

// w_t[in][out] = v[out][in] * g[out] / ||v[out][:]||   (linear.py:49; g == NULL -> plain transpose)
int launch_weight_fold_transpose(const float* v, const float* g, float* w_t, int out, int in,
                                 cudaStream_t st);

// fourier_weight [Cin][Cout][K][2] -> real block matrices Wblk[k][2C][2C]
int launch_pack_mix_weights(const float* w, float* Wblk, int C, int K, cudaStream_t st);
// DCT variant: real weights [Cin][Cout][Kc] -> block-diagonal pair matrices Wblk[k][2C][2C] = diag(W_2k, W_2k+1)
int launch_pack_mix_weights_dct(const float* w, float* Wblk, int C, int Kpairs, int Kc, cudaStream_t st);

// Complex-to-complex DFT along an axis as TWO real table products + this combine: Y[o][2R][q][2][C] holds, for each of the
// R output rows, the cosine sums (rows 0..R-1) and the sine sums (rows R..2R-1) of the (Re, Im) input pairs;
//   out[o][r][q][0][c] = Yc[..0..] + sgn * Ys[..1..],   out[o][r][q][1][c] = Yc[..1..] - sgn * Ys[..0..]
// sgn = +1: multiply by e^{-i theta} (forward), -1: by e^{+i theta} (inverse).  (torch.fft.rfft2 / irfft2 along dim -2,
// zongyi_fno/grid_plus_2d.py:57,78.)
int launch_c2c_combine(const float* Y, float* out, long long outer, int R, long long q, int C, float sgn, cudaStream_t st);
// Its transpose (backward pass): d[o][R][q][2][C] -> Y[o][2R][q][2][C] with Yc = d, Ys[..0..] = -sgn * d[..1..],
// Ys[..1..] = sgn * d[..0..].
int launch_c2c_expand(const float* d, float* Y, long long outer, int R, long long q, int C, float sgn, cudaStream_t st);

// Weff[j][c] = sum_h W1t[h][j] * W0t[c][h];  beff[j] = sum_h W1t[h][j]*b0[h] + b1[j]   (fp64 accumulate)
int launch_fold_head(const float* W0t /*[C][H]*/, const float* b0, const float* W1t /*[H][out]*/,
                     const float* b1, float* Weff, float* beff, int C, int H, int out, cudaStream_t st);

// Rollout glue (routines/grid_2d_markov.py:286-306, modules/normalizer.py:51,62)
//   feat[b][x][y][0] = (frame - mean0)/std0 ; feat[..][1] = (lin(x) - mean1)/std1 ; [2] likewise
struct MeanStd { float m[8]; float s[8]; };   // passed by value: no device copy, graph-capturable
// q, v: velocity features of the frame (NULL: 3 features [w, gx, gy]; else 5: [w, q, v, gx, gy])
int launch_rollout_features(const float* frame, long long frame_stride_b, int frame_stride_xy, const float* q,
                            const float* v, float* feat, int batch, int X, int Y, float low, float high,
                            const MeanStd& ms, cudaStream_t st);
// The same with the torus_vis extras appended after the position grid (routines/grid_2d_markov.py:246-260, :288-291):
// force [B, X, Y, force_steps] (frame `t`, or the only one of a static forcing; NULL: no channel), mu [B] (NULL: none).
int launch_rollout_features_ex(const float* frame, long long frame_stride_b, int frame_stride_xy, const float* q,
                               const float* v, const float* force, int force_steps, int t, const float* mu, float* feat,
                               int batch, int X, int Y, float low, float high, const MeanStd& ms, cudaStream_t st);
// (q, v) = (psi_y, -psi_x) from the vorticity w[b] at w + b * stride_b + (x * Y + y) * stride_xy on an Lx x Ly periodic
// domain (routines/grid_2d_markov.py:206-220); scratch: velocity_scratch_floats(batch, X, Y) floats
size_t velocity_scratch_floats(int batch, int X, int Y);
int launch_velocity(const float* w, long long stride_b, long long stride_xy, int batch, int X, int Y, float Lx, float Ly,
                    float* q, float* v, float* scratch, cudaStream_t st);
//   preds[b][x][y][t] = fc[b][x][y] * std0 + mean0
int launch_rollout_denorm(const float* forecast, float* preds, int batch, int XY, int n_steps, int t,
                          const MeanStd& ms, cudaStream_t st);

}  // namespace ffno

namespace ffno {
// y[P][N] = act(x[P][K] @ Wt[K][N] + bias) for any K, N (one thread per output; small/odd shapes)
int launch_linear_any(const float* x, const float* Wt, const float* bias, float* y, long long P, int K, int N,
                      bool relu, cudaStream_t st);
// out[b] = ||x_b - y_b|| / ||y_b||   (modules/loss.py:33-46)
int launch_rel_l2(const float* x, long long xsb, long long xsi, const float* y, long long ysb, long long ysi,
                  int batch, long long n, float* out, cudaStream_t st);
}  // namespace ffno
