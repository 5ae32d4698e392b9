// tcgen05 kernels of the F-FNO layer for width C = 64, FF factor 4 (hidden 256): declarations.
#pragma once
#include "common.cuh"

namespace ffno {

constexpr int kUmmaC = 64;
constexpr int kUmmaH = 256;
constexpr int kFFImageBytes = 131072;        // W1 hi|lo (32 KB each) + W2 hi|lo (32 KB each), smem image
constexpr int kMixImageBytes = 65536;        // per (axis, mode): B hi (32 KB) | B lo (32 KB), smem image

// w1t [64][256], w2t [256][64] (folded, transposed FP32) -> bf16 hi/lo operand image (K-major, SWIZZLE_128B)
int launch_pack_ff_image(const float* w1t, const float* w2t, uint8_t* image, cudaStream_t st);
// Wblk [K][128][128] (real block form of the complex mode weights) -> K images of kMixImageBytes
int launch_pack_mix_image(const float* wblk, uint8_t* image, int K, cudaStream_t st);

// Per-mode complex channel mix of one or more axes in one launch.
struct MixAxis {
  const float* F;
  float* R;
  const uint8_t* image;   // K images
  long long outer, p_inner;
  int K;
};

// Truncated real DFT along a strided axis on tcgen05 (forward and inverse are the same product with a
// different table):   Y[o][j][inner] (+)= sum_i T[i][j] * X[o][i][inner],   inner % 64 == 0, n_out <= 256
//   A = X tile (128 inner elements x n_in), stored MN-major straight from its natural layout;
//   B = table image [npad rows j][n_in] K-major (bf16 hi | lo), built once per plan by launch_pack_table_image.
struct AxisXform {
  const float* X;
  float* Y;
  const uint8_t* table;    // image: hi half then lo half, each kchunks * npad * 128 bytes
  long long outer, inner;
  int n_in, n_out, npad, kchunks, accumulate;
};
size_t table_image_bytes(int n_in, int n_out);
int launch_pack_table_image(const float* T /*[n_in][ldt]*/, int ldt, int n_in, int n_out, uint8_t* image,
                            cudaStream_t st);

// Warp-specialised, double-buffered kernels (umma_pipelined.cu): loads / MMAs / epilogues overlap.
bool axis_pipe_fits(int n_in, int n_out);
// `reverse`: walk the tiles last-to-first.  Consecutive launches of the layer loop alternate the direction so that a
// kernel starts with the data its predecessor wrote LAST (still resident in the 126 MB L2) instead of re-streaming it
// in the producer's order, which is the LRU worst case for a ~100 MB working set.
int launch_axis_pipe(const AxisXform* axes, int n_axes, int sm_count, cudaStream_t st, bool reverse = false);
int launch_mix_pipe(const MixAxis* axes, int n_axes, int sm_count, cudaStream_t st, bool reverse = false);

// FF: hidden activations staged in tensor memory (TS-mode GEMM2), coalesced output through a smem staging tile,
// loader sums up to three spectral partial outputs (s1 / s2 may be NULL).
// head_w/head_b/forecast (optional): fused folded 1-output head on the FF output, forecast[p] = <b_p, head_w> + head_b.
int launch_ff_ts(const float* s0, const float* s1, const float* s2, const float* residual, float* x_out, float* b_out,
                 const uint8_t* image, const float* b1, const float* b2, long long P, int sm_count, cudaStream_t st,
                 const float* head_w = nullptr, const float* head_b = nullptr, float* forecast = nullptr,
                 bool reverse = false);

// ---- stage-pipelined forward of the whole layer stack (umma_pipelined.cu: PipeDesc) -----------------------------
// Per-layer FeedForward parameters (device array, one entry per layer).
struct FFLayerArgs {
  const uint8_t* image;
  const float* b1;
  const float* b2;
};
struct StackPipeArgs {
  int n_layers, n_axes, n_units;
  AxisXform fwd[3];            // X = residual-stream buffer of the EVEN layers (xbuf[0]), Y = F of the axis
  const float* x_odd;          // residual-stream buffer of the odd layers (xbuf[1])
  MixAxis mix[3];              // image = the mode weights shared by every layer
  AxisXform inv[3];            // X = R of the axis, Y = the axis' spectral output s_a
  float* xbuf[2];
  const FFLayerArgs* ff_layers;
  const float* head_w;
  const float* head_b;
  float* forecast;
  long long P;                 // points of the whole batch
  unsigned* counters;          // stack_pipe_counter_bytes(n_units) of arrival counters (zeroed by the call)
  unsigned long long* dbg;     // diagnostics buffer or NULL
  int only_stage;              // diagnostics: >= 0 runs that stage alone with its dependencies pre-satisfied; -1 = all
  int sms[4];                  // CTAs of the forward-transform / mix / inverse-transform / FF stage
  cudaStream_t streams[4];     // [0] = the caller's stream; the others are forked from it and joined back
  cudaEvent_t fork, join[3];
};
size_t stack_pipe_counter_bytes(int n_units);
bool stack_pipe_geometry_ok(const AxisXform* fwd, const MixAxis* mix, int n_axes, long long P, int n_units);
int launch_stack_pipe(const StackPipeArgs& a);
// Do kernels of two streams really run at the same time in this process (false under serialising profilers)?
int probe_stream_concurrency(cudaStream_t s0, cudaStream_t s1, unsigned* dev_flags4, bool* concurrent);

// Diagnostics: in-kernel clock64 timeline of block 0 of ff_ts_kernel ([role 8][tile 16][event 8]).
int debug_timeline(int enable, long long* host_out);

// Diagnostics: D[128][N] = A[128][K] * B[N][K]^T with bf16 inputs (raw ushort), one CTA, every layout variant
// the kernels rely on (a_mn / b_mn: operand stored MN-major; variant: LBO/SBO interpretation under test).
int launch_umma_selftest(const uint16_t* A, const uint16_t* B, float* D, int N, int K, int a_mn, int b_mn,
                         int variant, cudaStream_t st);

}  // namespace ffno
