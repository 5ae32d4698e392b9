// tcgen05 (UMMA) fast path of the F-FNO layer: interface used by plan.cu.  Implemented in umma_path.cu.
#pragma once
#include "common.cuh"

namespace ffno {

struct UmmaState;

// Prepared FP32 parameters of one layer (device pointers owned by the plan) that the UMMA path re-packs
// into bf16 hi/lo operand tiles.
struct UmmaLayerSrc {
  const float* wmix[3];   // per axis [K][2C][2C] real block matrices
  const float* w1t;       // [C][4C]  folded, transposed FF linear 0
  const float* b1;        // [4C]
  const float* w2t;       // [4C][C]
  const float* b2;        // [C]
};

bool umma_supported(const ffno_desc* d, const int ext[3]);
const char* umma_why_not(const ffno_desc* d, const int ext[3]);
int umma_create(UmmaState** out, const ffno_desc* d, const int ext[3]);
void umma_destroy(UmmaState* s);
// The state serves the FeedForward only (plans whose spectral layer runs on the FP32 kernels: FFNO_TRANSFORM_RFFT2):
// no table / mode-weight images are built, only umma_ff_fwd / umma_ff_layer may be called.
void umma_set_ff_only(UmmaState* s);
int umma_ff_layer(UmmaState* s, int layer, const float* s_in, const float* residual, int batch, float* x_next, float* b_out,
                  cudaStream_t st);
// Persistent-kernel grid cap for the following launches (0 = all SMs): lets independent batch chunks on different
// streams share the GPU side by side.
void umma_set_sm_limit(UmmaState* s, int n);
int umma_load_params(UmmaState* s, const UmmaLayerSrc* layers, float* const d_fwd[3], float* const d_inv[3],
                     cudaStream_t st);
size_t umma_workspace_floats(const UmmaState* s, int batch);

// One full layer: x_next = x + FF(spectral(x)); optionally also materialises s (spectral output) and
// b (FF output before the residual; the head input on the last layer).
// Optional fused head (last layer, 1-output head, no padding): the FF writes forecast[p] = <b_p, w> + b directly.
struct UmmaFusedHead {
  const float* w = nullptr;
  const float* b = nullptr;
  float* forecast = nullptr;
};
bool umma_can_fuse_head(const UmmaState* s, bool want_s);
// Stage-pipelined forward of the whole stack (four persistent, flag-synchronised launches per forward instead of four
// per layer).  umma_pipeline_unit: samples per pipeline unit for this batch, 0 when the plan / batch does not qualify
// (3-D, padded or forked stacks, unshared mode weights, fewer than 4 units, kernels of different streams not running
// concurrently in this process, FFNO_B200_PERSIST=0).
int umma_pipeline_unit(const UmmaState* s, int batch);
int umma_pipe_debug(const UmmaState* s, unsigned long long* host_out, int n_words);   // diagnostics, see ffno_debug_pipe_stats
int umma_stack_fwd_pipelined(UmmaState* s, float* xa, float* xb, int batch, float* s_out, float* F, float* R, float* ws,
                             const UmmaFusedHead& head, cudaStream_t st);
int umma_layer_fwd(UmmaState* s, int layer, const float* x, int batch, float* x_next, float* s_out, float* b_out,
                   float* F, float* R, float* ws, bool want_s, bool want_b, cudaStream_t st,
                   const UmmaFusedHead* head = nullptr);
// accumulate: s_out += the operator's output instead of s_out = (the backward adds the adjoint onto the gradient stream)
int umma_spectral_fwd(UmmaState* s, int layer, const float* x, int batch, float* s_out, float* F, float* R,
                      float* ws, cudaStream_t st, bool accumulate = false);
// Per-axis outputs, no accumulation (falls back to the summed path into s_axis[0] when a kernel does not qualify;
// returns the number of buffers written through *n_written).
int umma_spectral_split_fwd(UmmaState* s, int layer, const float* x, int batch, float* const s_axis[3], float* F, float* R,
                            float* ws, int* n_written, cudaStream_t st);
// F_a = Fwd_a x for every axis, axis a at the same offset of F as in the calls above (the backward's weight gradient
// needs the forward spectra again; nothing else of the spectral operator runs).
int umma_forward_spectra(UmmaState* s, const float* x, int batch, float* F, cudaStream_t st);
size_t umma_spec_offset(const UmmaState* s, int batch, int axis);
int umma_ff_fwd(UmmaState* s, int layer, const float* s_in, const float* residual, int batch, float* y, float* ws,
                cudaStream_t st);

}  // namespace ffno
