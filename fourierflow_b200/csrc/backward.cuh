// Backward-pass kernels (generic FP32): reductions over points for the parameter gradients and the small adjoint
// helpers; the data-path adjoints reuse generic_kernels.cu with transposed tables / weights.  See backward.cu.
#pragma once
#include "common.cuh"

namespace ffno {

// dw[out][in] += sum_p dy[p][out] * x[p][in]
int launch_linear_wgrad(const float* dy, const float* x, float* dw, long long P, int out, int in, int sm_count,
                        cudaStream_t st);
// dw[ci][co][k][re|im] += the complex-weight gradient of the per-mode mix (grid_2d.py:65-68) from the forward spectra
// F and the gradient of the mixed spectra dR, both [outer][K][2][p_inner][C]
int launch_mix_wgrad(const float* F, const float* dR, float* dw, long long outer, int K, long long p_inner, int C,
                     int sm_count, cudaStream_t st, int dct_kc = 0, float* dw_hi = nullptr, int ksplit = 0);
// out[n] += sum_r x[r][n]   (N <= 256)
int launch_colsum(const float* x, float* out, long long rows, int N, cudaStream_t st);
// dh[i] = h[i] > 0 ? dh[i] : 0
int launch_relu_bwd(float* dh, const float* h, long long n, cudaStream_t st);
// y += x
int launch_axpy(float* y, const float* x, long long n, cudaStream_t st);
// torch weight_norm (dim 0) backward: dg[out], dv[out][in] += from dw[out][in]
int launch_wnorm_bwd(const float* dw, const float* v, const float* g, float* dg, float* dv, int out, int in,
                     cudaStream_t st);
// dst[z][c][r] = src[z][r][c]
int launch_transpose(const float* src, float* dst, int rows, int cols, int batch, cudaStream_t st);
struct LiftGeom;
// mesh variants: copy between the padded activation layout and the dense (unpadded) one; to_padded zero-fills the padding
int launch_crop_pad(float* dense, float* padded, int batch, const LiftGeom& g, bool to_padded, cudaStream_t st);
// rows[p] = [x[p][0..in) | linspace grid coordinates of point p] — the lift's input as its linear sees it
int launch_lift_rows(const float* x, float* rows, long long n_pts, const LiftGeom& g, cudaStream_t st);
// dst[r][0..ncols) = src[r][0..ncols) of a [rows][ld] matrix
int launch_take_cols(const float* src, float* dst, long long rows, int ld, int ncols, cudaStream_t st);
// LpLoss.rel backward (modules/loss.py:33-46), contiguous x, y [batch][n]
int launch_rel_l2_bwd(const float* x, const float* y, const float* gout, float* dx, int batch, long long n,
                      cudaStream_t st);

}  // namespace ffno
