/*
 * ffno_b200.h — C ABI of the B200-native Factorized-FNO forward (libffno_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of alasdairtran/fourierflow: the F-FNO layer-stack
 * forward.  The reference has no FFI (it is eager PyTorch); what binds here is the reference's
 * nn.Module surface, mirrored by fourierflow_b200/modules (same class names, constructor arguments,
 * forward signatures and state_dict keys).  Every entry point below names the reference code it
 * replaces (paths relative to the reference tree, fourierflow/...).
 *
 * Conventions
 *   - All tensors are fp32, contiguous, channels-LAST: activations [B, S1, S2(, S3), C] — the layout
 *     the reference modules take and return (modules/factorized_fno/grid_2d.py:43, :96).
 *   - Pointers are DEVICE pointers unless the function name ends in _host.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises
 *     except the *_host variants (which copy H2D/D2H on that stream and wait for it).
 *   - Return value: 0 on success, negative ffno_status otherwise; ffno_last_error() returns a
 *     thread-local human-readable message.  Nothing throws or aborts across the ABI.
 *   - Ownership: the caller owns every tensor and the workspace.  The library owns only what lives
 *     behind an ffno_plan* (DFT tables, folded/packed weights, TMA descriptors), created/destroyed by
 *     paired calls.  No CPU fallback exists: unsupported shapes return FFNO_ERR_UNSUPPORTED.
 *   - Threading: entry points are re-entrant; one plan must not be used from two threads at once.
 */
#ifndef FFNO_B200_H
#define FFNO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FFNO_API __attribute__((visibility("default")))
#else
#define FFNO_API
#endif

#define FFNO_ABI_VERSION 2
#define FFNO_MAX_DIMS 3
#define FFNO_MAX_FF_LAYERS 4

typedef enum {
  FFNO_OK = 0,
  FFNO_ERR_BAD_ARG = -1,      /* null pointer, inconsistent sizes */
  FFNO_ERR_UNSUPPORTED = -2,  /* shape/option the CUDA path does not implement */
  FFNO_ERR_CUDA = -3,         /* a CUDA runtime/driver call failed (message has the cudaError) */
  FFNO_ERR_WORKSPACE = -4,    /* workspace too small */
  FFNO_ERR_STATE = -5         /* e.g. forward before ffno_plan_load_params */
} ffno_status;

typedef enum {                /* SpectralConv2d `mode`, grid_2d.py:44,64,69 */
  FFNO_MODE_FULL = 0,
  FFNO_MODE_LOW_PASS = 1,
  FFNO_MODE_NO_FOURIER = 2
} ffno_spectral_mode;

typedef enum {                /* per-axis transform of the spectral layer */
  FFNO_TRANSFORM_RFFT = 0,    /* F-FNO: ortho rfft, K complex bins kept, complex channel mix (factorized_fno/*.py) */
  FFNO_TRANSFORM_DCT = 1,     /* factorized CNO sibling: ortho DCT-II, first K coefficients kept, REAL channel mix,
                                 inverse DCT-III (factorized_cno/mesh_3d.py:55-112, modules/dct.py:16-88);
                                 fourier_weight is [C, C, K_a] */
  FFNO_TRANSFORM_RFFT2 = 2    /* un-factorized sibling FNOPlus2DBlock (zongyi_fno/grid_plus_2d.py:52-83): ortho rfft2, the
                                 two K x K corner blocks (rows 0..K-1 and M-K..M-1, columns 0..K-1) mixed per 2-D mode,
                                 irfft2.  2-D only, modes[0] == modes[1] == K, 2K <= size[0]; fourier_weight[0] / [1] are
                                 the [C, C, K, K, 2] weights of the low / high row block; spectral layer on the FP32 kernels,
                                 FeedForward on tcgen05 when width = 64, ff_factor = 4 */
} ffno_transform;

typedef enum {                /* which implementation a plan may use */
  FFNO_PATH_AUTO = 0,         /* tcgen05 kernels when the shape qualifies, else the generic FP32 kernels */
  FFNO_PATH_GENERIC = 1,      /* FP32 FFMA kernels only (any width / length / mode count) */
  FFNO_PATH_UMMA = 2          /* tcgen05 kernels only; plan creation fails if the shape does not qualify */
} ffno_path;

/* Static description of one operator stack (everything except the batch size).
 * Mirrors the constructor arguments of FNOFactorized2DBlock (grid_2d.py:103-107),
 * FNOFactorizedMesh2D (mesh_2d.py:108-109) and FNOFactorizedMesh3D (mesh_3d.py:116-118). */
typedef struct {
  int32_t abi_version;            /* = FFNO_ABI_VERSION */
  int32_t ndim;                   /* spatial axes: 2 or 3 */
  int32_t size[FFNO_MAX_DIMS];    /* spatial extent of the INPUT, tensor-axis order (X, Y, Z) */
  int32_t pad[FFNO_MAX_DIMS];     /* zeros appended on the high side of each axis after the lift
                                     (mesh variants: 8, mesh_3d.py:120,165; periodic grid: 0) */
  int32_t modes[FFNO_MAX_DIMS];   /* retained rfft bins per tensor axis */
  int32_t width;                  /* C */
  int32_t in_features;            /* width of the tensor given to forward (before grid append) */
  int32_t append_grid;            /* 1: concatenate linspace(0,1,size_a) per axis (mesh_3d.py:178-189) */
  int32_t out_features;           /* head output width (1, or output_dim) */
  int32_t head_hidden;            /* 128 (grid_2d.py:150-152) */
  int32_t n_layers;
  int32_t ff_factor;              /* hidden = ff_factor * width (feedforward.py:11-12) */
  int32_t n_ff_layers;            /* feedforward.py:10 (shipped configs: 2) */
  int32_t layer_norm;             /* LayerNorm(width) on the FF output (feedforward.py:17) */
  int32_t use_fork;               /* forecast fork (grid_2d.py:48,164-167) */
  int32_t spectral_mode;          /* ffno_spectral_mode */
  int32_t path;                   /* ffno_path */
  int32_t transform;              /* ffno_transform (ABI 2) */
} ffno_desc;

/* One WNLinear (modules/linear.py:41-51).  Either `weight` (plain nn.Linear) or weight_g+weight_v
 * (torch weight_norm: w = g * v / ||v||_row, recomputed by the reference on every call). */
typedef struct {
  const float* weight;            /* [out, in] or NULL */
  const float* weight_g;          /* [out] (stored [out,1]) or NULL */
  const float* weight_v;          /* [out, in] or NULL */
  const float* bias;              /* [out] or NULL */
  int32_t in_features;
  int32_t out_features;
} ffno_linear_params;

/* FeedForward (modules/feedforward.py:6-24): n_ff_layers linears (+ optional LayerNorm on the last) */
typedef struct {
  ffno_linear_params linear[FFNO_MAX_FF_LAYERS];
  const float* ln_weight;         /* [width] or NULL */
  const float* ln_bias;
} ffno_ff_params;

/* One spectral layer (SpectralConv2d, grid_2d.py:10-49).  fourier_weight is indexed by TENSOR AXIS
 * (0 = X = first spatial axis), each [C, C, K_a, 2] fp32 with (re, im) innermost.  The host mirror
 * resolves the reference's inconsistent ParameterList order (grid_2d: [0]->Y,[1]->X; mesh: [0]->X ...). */
typedef struct {
  const float* fourier_weight[FFNO_MAX_DIMS];
  ffno_ff_params backcast_ff;
  ffno_ff_params forecast_ff;     /* only read when use_fork */
} ffno_layer_params;

typedef struct {
  ffno_linear_params in_proj;     /* lift, grid_2d.py:112,157 */
  ffno_linear_params out0;        /* head, grid_2d.py:150-152: Linear(C,128) -> Linear(128,out), no activation */
  ffno_linear_params out1;
  const ffno_layer_params* layers;
  int32_t n_layers;
} ffno_block_params;

/* Optional per-layer taps for per-layer parity (device pointers, any may be NULL). */
typedef struct {
  float* lift;                    /* [B, S.., C]   x after in_proj (+pad) */
  float* const* x_after;          /* n_layers pointers: x after the residual of layer l */
  float* const* spectral;         /* n_layers pointers: forward_fourier output of layer l */
  float* b_last;                  /* backcast of the last layer (head input) */
  float* const* forecast_list;    /* use_fork: n_layers pointers [B, S.., out_features] */
} ffno_taps;

typedef struct ffno_plan ffno_plan;

FFNO_API const char* ffno_last_error(void);
FFNO_API int ffno_abi_version(void);
/* 1 if a CUDA device of compute capability 10.x is present and usable. */
FFNO_API int ffno_device_ok(void);

/* Plan lifetime.  Builds DFT tables (host double -> device) and reserves folded/packed weights. */
FFNO_API int ffno_plan_create(const ffno_desc* desc, ffno_plan** out_plan);
FFNO_API int ffno_plan_destroy(ffno_plan* plan);
/* 1 if the plan runs tcgen05 (UMMA) kernels (for FFNO_TRANSFORM_RFFT2 plans: the FeedForward only), 0 if it runs the
   generic FP32 kernels throughout. */
FFNO_API int ffno_plan_uses_umma(const ffno_plan* plan);

/* Fold weight-norm (linear.py:49), re-layout spectral weights to per-mode real block matrices and
 * (UMMA path) split to bf16 hi/lo operand tiles.  Call again whenever a parameter changes. */
FFNO_API int ffno_plan_load_params(ffno_plan* plan, const ffno_block_params* params, void* stream);

/* Bytes of scratch the forward needs for `batch` samples. */
FFNO_API size_t ffno_workspace_bytes(const ffno_plan* plan, int32_t batch);

/* Whole-stack forward: replaces FNOFactorized2DBlock.forward (grid_2d.py:154-177) and
 * FNOFactorizedMesh2D/3D.forward (mesh_2d.py:149-165, mesh_3d.py:160-176).
 *   x        [B, size.., in_features]
 *   forecast [B, size.., out_features]
 * taps may be NULL. */
FFNO_API int ffno_block_fwd(ffno_plan* plan, const float* x, int32_t batch, float* forecast,
                   const ffno_taps* taps, void* workspace, size_t workspace_bytes, void* stream);

/* Same through HOST buffers: copies x H2D, runs, copies forecast D2H, waits.  Device scratch is the
 * caller's `workspace` (device), which already contains the two staging areas.
 * Calls without taps replay a CUDA graph of the launch sequence from their second use on (same batch and
 * workspace); set FFNO_B200_GRAPH=0 to force eager launches. */
FFNO_API size_t ffno_workspace_bytes_host(const ffno_plan* plan, int32_t batch);
FFNO_API int ffno_block_fwd_host(ffno_plan* plan, const float* x_host, int32_t batch, float* forecast_host,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Layer-level entry points (module-level drop-ins).  `layer` indexes the loaded parameters.
 *   ffno_spectral_fwd   = SpectralConv2d.forward_fourier (grid_2d.py:51-99; mesh_3d.py:55-112)
 *   ffno_ff_fwd         = FeedForward.forward (feedforward.py:21-24); which: 0 backcast, 1 forecast;
 *                         if residual != NULL, y = residual + FF(s)  (grid_2d.py:169)
 * x, s, y: [B, S.. (padded extent), C]. */
FFNO_API int ffno_spectral_fwd(ffno_plan* plan, int32_t layer, const float* x, int32_t batch, float* s,
                      void* workspace, size_t workspace_bytes, void* stream);
FFNO_API int ffno_ff_fwd(ffno_plan* plan, int32_t layer, int32_t which, const float* s, const float* residual,
                int32_t batch, float* y, void* workspace, size_t workspace_bytes, void* stream);

/* The spectral operator exactly as the layer loop runs it (tcgen05 path: forward transforms of all axes, mode mixes of
 * all axes, inverse transforms of all axes — 3 launches): s_axis[a] receives axis a's contribution; forward_fourier's
 * result is their sum (grid_2d.py:94), which the FeedForward kernel forms while loading.  On plans that run the generic
 * kernels only s_axis[0] is written (already summed) and the other pointers are ignored. */
FFNO_API int ffno_spectral_split_fwd(ffno_plan* plan, int32_t layer, const float* x, int32_t batch, float* const* s_axis,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Stateless WNLinear.forward (modules/linear.py:41-51): y[rows, out] = act(x[rows, in] @ W^T + b).
 * `scratch` (device, >= in*out*4 bytes) receives the folded, transposed weight. */
FFNO_API int ffno_linear_fwd(const ffno_linear_params* lin, const float* x, int64_t rows, float* y, int32_t relu,
                    void* scratch, size_t scratch_bytes, void* stream);
/* nn.LayerNorm(C) over the last dim, eps 1e-5 (modules/feedforward.py:17). */
FFNO_API int ffno_layernorm_fwd(const float* x, const float* weight, const float* bias, int64_t rows, int32_t C,
                       float* y, void* stream);
/* LpLoss.rel per sample (modules/loss.py:33-46): out[b] = ||x_b - y_b||_2 / ||y_b||_2 over n elements;
 * element (b, i) of x lives at x[b*x_stride_b + i*x_stride_i] (same for y).  The mean over the (sharded)
 * batch is the caller's one collective. */
FFNO_API int ffno_rel_l2(const float* x, int64_t x_stride_b, int64_t x_stride_i, const float* y, int64_t y_stride_b,
                int64_t y_stride_i, int32_t batch, int64_t n, float* out, void* stream);

/* 10-step Markov rollout of Grid2DMarkovExperiment._valid_step (routines/grid_2d_markov.py:195-326)
 * for the torus_li/markov configuration (use_position, should_normalize; 2-D periodic grid only) and, for plans
 * with in_features = 5, the torus_kochkov one (use_velocity: features [w, q, v, gx, gy], see ffno_velocity_fwd):
 *   frame0 [B, X, Y]        ground-truth vorticity fed to step 0
 *   mean/std [in_features]  Normalizer statistics (modules/normalizer.py:68-77), host pointers
 *   preds  [B, X, Y, n_steps]
 * Each step: concat(frame, position grid linspace(low, high)) -> normalise -> stack forward ->
 * de-normalise with channel 0.  The loss reduction is left to the caller (needs the targets). */
FFNO_API int ffno_rollout_fwd(ffno_plan* plan, const float* frame0, int32_t batch, int32_t n_steps,
                     const float* mean_host, const float* std_host, float low, float high,
                     float* preds, void* workspace, size_t workspace_bytes, void* stream);
FFNO_API size_t ffno_rollout_workspace_bytes(const ffno_plan* plan, int32_t batch, int32_t n_steps);

/* The same rollout with the optional feature channels of the torus_vis / torus_vis_force configurations
 * (append_force, append_mu: routines/grid_2d_markov.py:146-152 training-time, :246-260 and :288-291 in _valid_step).
 * Features per step: [w, (q, v,) gx, gy, (f,) (mu)]; plan in_features must equal their count. */
typedef struct ffno_rollout_extras {
  int32_t use_velocity;   /* 1: stream-function velocities q, v after w (as ffno_rollout_fwd does for in_features = 5) */
  int32_t force_steps;    /* 0: no forcing channel; 1: static forcing; n_steps: frame t of the forcing at step t */
  const float* force;     /* device [B, X, Y, force_steps] (the reference's batch['f'], 4-D: its last n_steps frames) */
  const float* mu;        /* device [B] viscosity per sample, broadcast over the grid; NULL: no channel */
} ffno_rollout_extras;
FFNO_API int ffno_rollout_fwd_ex(ffno_plan* plan, const float* frame0, int32_t batch, int32_t n_steps,
                        const float* mean_host, const float* std_host, float low, float high,
                        const ffno_rollout_extras* extras, float* preds, void* workspace, size_t workspace_bytes,
                        void* stream);
FFNO_API size_t ffno_rollout_workspace_bytes_ex(const ffno_plan* plan, int32_t batch, int32_t n_steps,
                                       int32_t use_velocity, int32_t force_steps, int32_t has_mu);

/* ---- Backward pass ------------------------------------------------------------------------------
 * What torch.autograd computes for the reference's training step (routines/grid_2d_markov.py:172-193 `_training_step`:
 * forecast -> Normalizer.inverse -> LpLoss; routines/base.py:27-52 applies the optimizer), written out as explicit
 * adjoints.  All three transforms (RFFT2: FP32 kernels; its two
 * fourier_weight gradients are given together or not at all).  Supported stacks: the 2-D grid block and
 * the mesh variants (grid append, zero padding and crop are part of the adjoint), n_ff_layers = 2, no LayerNorm, no fork,
 * mode 'full' (every torus_li / torus_kochkov / torus_vis / plasticity / airfoil F-FNO and F-CNO config).  The forward is
 * recomputed inside the call, so no state is carried between the forward and the backward; which kernels run the
 * recompute and the spectral adjoint is selected by ffno_plan_set_backward_mode.
 *
 * Gradient buffers have the shapes of the parameters they belong to and are ACCUMULATED into (+=), like `.grad`: zero
 * them first; a parameter shared by several layers (share_weight) is given the same buffer in each layer and receives
 * the sum.  NULL pointers skip that gradient. */
typedef struct {
  float* weight;                  /* [out, in]  (plain linear)            */
  float* weight_g;                /* [out]      (weight-normed linear)    */
  float* weight_v;                /* [out, in]                            */
  float* bias;                    /* [out] */
} ffno_linear_grads;
typedef struct {
  float* fourier_weight[FFNO_MAX_DIMS];        /* shaped like the parameter: [C, C, K_a, 2] | DCT [C, C, K_a] | RFFT2 [C, C, K, K, 2] x 2 */
  ffno_linear_grads backcast_ff[FFNO_MAX_FF_LAYERS];
} ffno_layer_grads;
typedef struct {
  ffno_linear_grads in_proj, out0, out1;
  const ffno_layer_grads* layers;
  int32_t n_layers;
} ffno_block_grads;
/*   x           [B, S.., in_features]   the forward's input
 *   d_forecast  [B, S.., out_features]  gradient of the loss w.r.t. the forecast
 *   params      the parameters the plan was loaded with (the weight-norm backward needs the raw g, v)
 *   dx          [B, S.., in_features] or NULL: gradient w.r.t. the input (overwritten) */
FFNO_API size_t ffno_block_bwd_workspace_bytes(const ffno_plan* plan, int32_t batch);
FFNO_API int ffno_block_bwd(ffno_plan* plan, const ffno_block_params* params, const float* x, const float* d_forecast,
                   int32_t batch, const ffno_block_grads* grads, float* dx, void* workspace, size_t workspace_bytes,
                   void* stream);
/* Precision / speed of ffno_block_bwd on plans that use the tcgen05 kernels (initial value from FFNO_B200_BWD):
 *   0 default : forward recompute on the FP32 kernels (ReLU masks equal the reference's to FP32 round-off), spectral
 *               adjoint on the tcgen05 kernels — every gradient tensor within ~1e-5 of its own max
 *   1 fast    : recompute on the tcgen05 kernels too (1.4x faster); hidden units within ~1e-5 of the ReLU kink may
 *               take the other branch: the gradient of the tcgen05-computed forward, not of the reference's
 *   2 fp32    : everything on the FP32 kernels */
/* Backward of the LAYER LOOP alone, for plans loaded without lift / head (the interior of the geo-F-FNO,
 * factorized_fno/point_cloud_2d.py:198-210):  x_{l+1} = x_l + backcast_ff_l(spectral_l(x_l)) + bias,  l = 0 .. n_layers-1,
 * `bias` [points, C] broadcast over the batch (NULL: none).  d_out = dL/dx_{n_layers} [B, points, C]; writes
 * dx = dL/dx_0 (NULL: skipped), accumulates the layer gradients (grads->layers; in_proj / out0 / out1 are ignored) and
 * d_bias += dL/dbias (NULL: skipped).  FP32 kernels; workspace as ffno_block_bwd. */
FFNO_API int ffno_layers_bwd(ffno_plan* plan, const ffno_block_params* params, const float* x0, const float* d_out,
                    const float* bias, int32_t batch, const ffno_block_grads* grads, float* dx, float* d_bias,
                    void* workspace, size_t workspace_bytes, void* stream);
FFNO_API int ffno_plan_set_backward_mode(ffno_plan* plan, int32_t mode);
/* Backward of ffno_rel_l2 for contiguous x, y [batch, n]: dx[b, i] = g_out[b] * d out[b] / d x[b, i]
 * (modules/loss.py:33-46; the mean over the batch is the caller's g_out = 1 / batch). */
FFNO_API int ffno_rel_l2_bwd(const float* x, const float* y, const float* g_out, int32_t batch, int64_t n, float* dx,
                    void* stream);

/* Velocity features of the torus_kochkov rollout (use_velocity=True, routines/grid_2d_markov.py:82-93, :206-220):
 * (q, v) = (psi_y, -psi_x) of the stream function of the vorticity w (w = -laplace psi) on a periodic
 * length_x x length_y domain, i.e. irfftn(+-2 pi i k psi_hat) with the reference's kx/ky/lap buffers.
 *   w element (b, x, y) at w[b*stride_b + (x*Y + y)*stride_xy];  q, v: contiguous [B, X, Y].
 * A plan built with in_features = 5 runs ffno_rollout_fwd with the features [w, q, v, gx, gy] and recomputes q, v from
 * every fed-back forecast; ffno_plan_set_domain sets its domain lengths (default 2 pi x 2 pi, the reference default). */
FFNO_API size_t ffno_velocity_scratch_bytes(int32_t batch, int32_t X, int32_t Y);
FFNO_API int ffno_velocity_fwd(const float* w, int64_t stride_b, int64_t stride_xy, int32_t batch, int32_t X, int32_t Y,
                      float length_x, float length_y, float* q, float* v, void* scratch, size_t scratch_bytes,
                      void* stream);
FFNO_API int ffno_plan_set_domain(ffno_plan* plan, float length_x, float length_y);

/* Diagnostics: one tcgen05 product D[128][N] = A[128][K] * B[N][K]^T (bf16 bit patterns in, FP32 out) through
 * the operand layouts / descriptors the kernels use (a_mn / b_mn: operand stored MN-major; variant selects the
 * LBO/SBO convention under test).  Used by tests/test_gpu_umma.py to pin the layout conventions on hardware. */
FFNO_API int ffno_umma_selftest(const uint16_t* A, const uint16_t* B, float* D, int32_t N, int32_t K, int32_t a_mn,
                       int32_t b_mn, int32_t variant, void* stream);

/* Diagnostics: arm (enable = 1) / disarm (0) / only read (-1) the in-kernel clock64 timeline of block 0 of the
 * pipelined FF kernel; host_out (may be NULL) receives [role 8][tile 16][event 8] cycle stamps. */
FFNO_API int ffno_debug_timeline(int32_t enable, int64_t* host_out);

/* Number of kernel launches the last ffno_block_fwd / ffno_rollout_fwd on this plan enqueued. */
FFNO_API int64_t ffno_plan_last_launch_count(const ffno_plan* plan);

/* 1 once a CUDA graph of the stack forward (ffno_block_fwd without taps, ffno_block_fwd_host) or of the rollout
 * (ffno_rollout_fwd) has been captured and is what the next identical call replays; 0 while calls still launch
 * kernel by kernel (first two calls of a shape, FFNO_B200_GRAPH=0, or capture refused).  Work submitted on the
 * legacy default stream is captured and replayed on a plan-owned stream fenced with events, since that stream
 * cannot be captured. */
FFNO_API int ffno_plan_graph_active(const ffno_plan* plan);

/* Samples per pipeline unit when ffno_block_fwd / ffno_rollout_fwd run `batch` samples through the STAGE-PIPELINED
 * forward — the transform / mix / inverse / FeedForward kernels launched once per forward as persistent,
 * flag-synchronised stages on disjoint SMs (replaces the layer loop of grid_2d.py:160-169 as a whole) — or 0 when this
 * plan and batch use one launch per stage and layer (3-D, padded or forked stacks, unshared mode weights, fewer than
 * four units, FFNO_B200_PERSIST not set to 1 (the path is opt-in), or a process whose kernels do not run concurrently,
 * e.g. under a profiler). */
FFNO_API int ffno_plan_pipeline_unit(const ffno_plan* plan, int32_t batch);

/* Diagnostics of the stage-pipelined forward (plans created with FFNO_B200_PIPE_DEBUG=1, FFNO_B200_GRAPH=0): copies
 * up to n_words 64-bit words to the host — 8 header words {CTAs of the forward-transform, mix, inverse-transform and
 * FF stage, 0...} followed, stage after stage, by 4 words per CTA of the last forward: SM cycles its loaders spent
 * blocked on the producer stage, SM cycles in the kernel, start and end time (globaltimer ns).  Returns the number
 * of words written, or a negative status. */
FFNO_API int ffno_debug_pipe_stats(const ffno_plan* plan, uint64_t* host_out, int32_t n_words);

#ifdef __cplusplus
}
#endif
#endif /* FFNO_B200_H */
