// tcgen05 (UMMA) fast path of the F-FNO layer for width 64 / FF factor 4: state, parameter images, dispatch.
//
// Per layer (fast path, every axis on the pipelined kernels): all forward transforms -> all mode mixes -> all inverse
// transforms (one output per axis) -> FeedForward (+ sum over axes, + residual), 4 launches; or the same stages as
// persistent, flag-synchronised roles launched once per forward (umma_persistent.cu).  An axis whose DFT table does
// not fit the pipelined transform kernel (n_out > 256 or a table image beyond shared memory) runs that transform on
// the FP32 table kernel of generic_kernels.cu; its mix still runs on tcgen05.
#include <cstdlib>
#include <map>
#include <vector>

#include "generic_kernels.cuh"
#include "umma_kernels.cuh"
#include "umma_path.cuh"

namespace ffno {

struct UmmaLayer {
  uint8_t* mix_image[3] = {nullptr, nullptr, nullptr};
  uint8_t* ff_image = nullptr;
  const float* b1 = nullptr;
  const float* b2 = nullptr;
};

struct UmmaState {
  ffno_desc d;
  int ext[3];
  int sm_count = 148;                                    // CTAs a launch may use (<= hw_sm_count)
  int hw_sm_count = 148;
  std::vector<UmmaLayer> layers;
  std::vector<void*> owned;
  std::map<std::pair<const void*, int>, uint8_t*> mix_cache;   // (source, axis) -> image
  std::map<const void*, uint8_t*> ff_cache;
  float* d_fwd[3] = {nullptr, nullptr, nullptr};
  float* d_inv[3] = {nullptr, nullptr, nullptr};
  bool zigzag = true;                                    // FFNO_ZIGZAG=0: every kernel walks its tiles first-to-last
  bool ff_only = false;                                  // only the FeedForward runs here (FFNO_TRANSFORM_RFFT2 plans)
  uint8_t* fwd_image[3] = {nullptr, nullptr, nullptr};   // tcgen05 table images (NULL -> FP32 table kernel)
  uint8_t* inv_image[3] = {nullptr, nullptr, nullptr};
  // stage-pipelined forward (launch_stack_pipe): FFNO_B200_PERSIST=1 turns it on, FFNO_B200_PIPE_SMS=f,m,i,ff sets
  // the CTAs per stage
  bool persist_env = false;
  bool concurrent = false;                               // probe_stream_concurrency at parameter load
  bool mix_shared = false;                               // every layer uses the same mode-weight images
  cudaStream_t pstream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t pfork = nullptr, pjoin[3] = {nullptr, nullptr, nullptr};
  unsigned* counters = nullptr;
  int counter_units = 0;
  FFLayerArgs* ff_layers_dev = nullptr;
  int pipe_sms[4] = {0, 0, 0, 0};
  unsigned long long* pipe_dbg = nullptr;                // FFNO_B200_PIPE_DEBUG=1: per-CTA wait statistics
};

static int pad16i(int n) { return (n + 15) / 16 * 16; }
constexpr int kPipeDebugWords = 8 + 4 * 256 + 4 * 64;
static int alloc_bytes(UmmaState* s, size_t bytes, uint8_t** out);
static bool axis_fits_umma(int n_in, int n_out) { return axis_pipe_fits(n_in, n_out); }

const char* umma_why_not(const ffno_desc* d, const int*) {
  if (d->width != kUmmaC) return "width != 64";
  if (d->ff_factor != 4) return "ff_factor != 4";
  if (d->n_ff_layers != 2) return "n_ff_layers != 2";
  if (d->layer_norm) return "layer_norm";
  if (d->use_fork) return "use_fork";
  if (d->spectral_mode == FFNO_MODE_NO_FOURIER) return "mode no-fourier";
  return nullptr;
}
bool umma_supported(const ffno_desc* d, const int* ext) { return umma_why_not(d, ext) == nullptr; }

int umma_create(UmmaState** out, const ffno_desc* d, const int ext[3]) {
  UmmaState* s = new UmmaState();
  s->d = *d;
  for (int a = 0; a < 3; ++a) s->ext[a] = ext[a];
  int dev = 0;
  FFNO_CUDA_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  FFNO_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    delete s;
    return set_error(FFNO_ERR_UNSUPPORTED, "tcgen05 path needs compute capability 10.x, device is %d.%d", prop.major,
                     prop.minor);
  }
  s->sm_count = s->hw_sm_count = prop.multiProcessorCount;
  const char* zz = getenv("FFNO_ZIGZAG");
  s->zigzag = !(zz && zz[0] == '0');
  const char* pe = getenv("FFNO_B200_PERSIST");
  s->persist_env = pe && pe[0] == '1';       // opt-in: measured slower than one launch per stage and layer
  const char* pd = getenv("FFNO_B200_PIPE_DEBUG");
  if (pd && pd[0] == '1') {
    uint8_t* raw = nullptr;
    if (alloc_bytes(s, kPipeDebugWords * 8, &raw) == FFNO_OK) {
      s->pipe_dbg = reinterpret_cast<unsigned long long*>(raw);
      cudaMemset(raw, 0, kPipeDebugWords * 8);
    }
  }
  const char* ps = getenv("FFNO_B200_PIPE_SMS");
  if (ps) sscanf(ps, "%d,%d,%d,%d", &s->pipe_sms[0], &s->pipe_sms[1], &s->pipe_sms[2], &s->pipe_sms[3]);
  s->layers.resize(d->n_layers);
  *out = s;
  return FFNO_OK;
}

void umma_set_ff_only(UmmaState* s) { s->ff_only = true; }

void umma_set_sm_limit(UmmaState* s, int n) {
  s->sm_count = (n >= 1 && n < s->hw_sm_count) ? n : s->hw_sm_count;
}

void umma_destroy(UmmaState* s) {
  if (!s) return;
  for (int i = 0; i < 3; ++i) {
    if (s->pstream[i]) cudaStreamDestroy(s->pstream[i]);
    if (s->pjoin[i]) cudaEventDestroy(s->pjoin[i]);
  }
  if (s->pfork) cudaEventDestroy(s->pfork);
  for (void* p : s->owned) cudaFree(p);
  delete s;
}

static int alloc_bytes(UmmaState* s, size_t bytes, uint8_t** out) {
  void* p = nullptr;
  FFNO_CUDA_CHECK(cudaMalloc(&p, bytes));
  s->owned.push_back(p);
  *out = static_cast<uint8_t*>(p);
  return FFNO_OK;
}

int umma_load_params(UmmaState* s, const UmmaLayerSrc* layers, float* const d_fwd[3], float* const d_inv[3],
                     cudaStream_t st) {
  for (int a = 0; a < 3; ++a) { s->d_fwd[a] = d_fwd[a]; s->d_inv[a] = d_inv[a]; }
  for (int a = 0; a < s->d.ndim && !s->ff_only; ++a) {
    const int Ln = s->ext[a], K2 = 2 * s->d.modes[a];
    if (!s->fwd_image[a] && axis_fits_umma(Ln, K2)) {
      FFNO_TRY(alloc_bytes(s, table_image_bytes(Ln, K2), &s->fwd_image[a]));
      FFNO_TRY(launch_pack_table_image(d_fwd[a], pad16i(K2), Ln, K2, s->fwd_image[a], st));
    }
    if (!s->inv_image[a] && axis_fits_umma(K2, Ln)) {
      FFNO_TRY(alloc_bytes(s, table_image_bytes(K2, Ln), &s->inv_image[a]));
      FFNO_TRY(launch_pack_table_image(d_inv[a], pad16i(Ln), K2, Ln, s->inv_image[a], st));
    }
  }
  std::map<std::pair<const void*, int>, bool> mix_done;
  std::map<const void*, bool> ff_done;
  for (int l = 0; l < s->d.n_layers; ++l) {
    UmmaLayer& L = s->layers[l];
    const UmmaLayerSrc& src = layers[l];
    if (s->d.spectral_mode == FFNO_MODE_FULL && !s->ff_only) {
      for (int a = 0; a < s->d.ndim; ++a) {
        auto key = std::make_pair((const void*)src.wmix[a], a);
        uint8_t*& img = s->mix_cache[key];
        if (!img) FFNO_TRY(alloc_bytes(s, (size_t)s->d.modes[a] * kMixImageBytes, &img));
        if (!mix_done[key]) {
          FFNO_TRY(launch_pack_mix_image(src.wmix[a], img, s->d.modes[a], st));
          mix_done[key] = true;
        }
        L.mix_image[a] = img;
      }
    }
    uint8_t*& img = s->ff_cache[(const void*)src.w1t];
    if (!img) FFNO_TRY(alloc_bytes(s, kFFImageBytes, &img));
    if (!ff_done[(const void*)src.w1t]) {
      FFNO_TRY(launch_pack_ff_image(src.w1t, src.w2t, img, st));
      ff_done[(const void*)src.w1t] = true;
    }
    L.ff_image = img;
    L.b1 = src.b1;
    L.b2 = src.b2;
  }
  // ---- stage-pipelined forward: per-layer FF table on the device, side streams, concurrency probe ----------------
  s->mix_shared = s->d.spectral_mode == FFNO_MODE_FULL;
  for (int l = 1; l < s->d.n_layers; ++l)
    for (int a = 0; a < s->d.ndim; ++a) s->mix_shared &= s->layers[l].mix_image[a] == s->layers[0].mix_image[a];
  if (s->persist_env && s->d.ndim == 2) {
    if (!s->ff_layers_dev) {
      uint8_t* raw = nullptr;
      FFNO_TRY(alloc_bytes(s, sizeof(FFLayerArgs) * s->d.n_layers, &raw));
      s->ff_layers_dev = reinterpret_cast<FFLayerArgs*>(raw);
    }
    std::vector<FFLayerArgs> host(s->d.n_layers);
    for (int l = 0; l < s->d.n_layers; ++l) host[l] = FFLayerArgs{s->layers[l].ff_image, s->layers[l].b1, s->layers[l].b2};
    FFNO_CUDA_CHECK(cudaStreamSynchronize(st));
    FFNO_CUDA_CHECK(cudaMemcpy(s->ff_layers_dev, host.data(), sizeof(FFLayerArgs) * host.size(), cudaMemcpyHostToDevice));
    if (!s->pfork) {
      FFNO_CUDA_CHECK(cudaEventCreateWithFlags(&s->pfork, cudaEventDisableTiming));
      for (int i = 0; i < 3; ++i) {
        FFNO_CUDA_CHECK(cudaStreamCreateWithFlags(&s->pstream[i], cudaStreamNonBlocking));
        FFNO_CUDA_CHECK(cudaEventCreateWithFlags(&s->pjoin[i], cudaEventDisableTiming));
      }
      uint8_t* flags = nullptr;
      FFNO_TRY(alloc_bytes(s, 4 * sizeof(unsigned), &flags));
      FFNO_TRY(probe_stream_concurrency(s->pstream[0], s->pstream[1], reinterpret_cast<unsigned*>(flags), &s->concurrent));
    }
  }
  return FFNO_OK;
}

size_t umma_workspace_floats(const UmmaState* s, int batch) {
  size_t U = (size_t)batch * kUmmaC;
  for (int a = 0; a < s->d.ndim; ++a) U *= s->ext[a];
  return (size_t)(s->d.ndim - 1) * U;      // per-axis spectral outputs of axes 1.. (axis 0 uses the plan's `s`)
}

// F / R hold every axis back to back: axis a starts at spec_offset(a) floats.
static size_t spec_offset(const UmmaState* s, int batch, int axis) {
  size_t P = (size_t)batch;
  for (int a = 0; a < s->d.ndim; ++a) P *= s->ext[a];
  size_t off = 0;
  for (int a = 0; a < axis; ++a) off += P / s->ext[a] * 2 * s->d.modes[a] * kUmmaC;
  return off;
}

static void axis_geom(const UmmaState* s, int batch, int a, long long* outer, long long* p_inner) {
  *outer = batch;
  *p_inner = 1;
  for (int i = 0; i < a; ++i) *outer *= s->ext[i];
  for (int i = a + 1; i < s->d.ndim; ++i) *p_inner *= s->ext[i];
}

// Summed spectral operator, one axis after the other (reference order: last axis first, grid_2d.py:57,75): used when
// a caller wants the sum materialised (taps, the standalone ffno_spectral_fwd) or an axis does not fit the pipelined
// transform kernel.
int umma_spectral_fwd(UmmaState* s, int layer, const float* x, int batch, float* s_out, float* F, float* R, float*,
                      cudaStream_t st, bool accumulate) {
  const UmmaLayer& L = s->layers[layer];
  const bool full = s->d.spectral_mode == FFNO_MODE_FULL;
  bool first = !accumulate;
  for (int a = s->d.ndim - 1; a >= 0; --a) {
    long long outer, p_inner;
    axis_geom(s, batch, a, &outer, &p_inner);
    const int Ln = s->ext[a], K = s->d.modes[a];
    const long long inner = p_inner * kUmmaC;
    float* Fa = F + spec_offset(s, batch, a);
    float* Ra = R + spec_offset(s, batch, a);
    if (s->fwd_image[a]) {
      AxisXform t{x, Fa, s->fwd_image[a], outer, inner, Ln, 2 * K, pad16i(2 * K), (Ln + 63) / 64, 0};
      FFNO_TRY(launch_axis_pipe(&t, 1, s->sm_count, st));
    } else {
      FFNO_TRY(launch_axis_transform(x, s->d_fwd[a], Fa, outer, Ln, 2 * K, inner, false, st));
    }
    const float* src = Fa;
    if (full) {
      MixAxis ax{Fa, Ra, L.mix_image[a], outer, p_inner, K};
      FFNO_TRY(launch_mix_pipe(&ax, 1, s->sm_count, st));
      src = Ra;
    }
    if (s->inv_image[a]) {
      AxisXform t{src, s_out, s->inv_image[a], outer, inner, 2 * K, Ln, pad16i(Ln), (2 * K + 63) / 64, first ? 0 : 1};
      FFNO_TRY(launch_axis_pipe(&t, 1, s->sm_count, st));
    } else {
      FFNO_TRY(launch_axis_transform(src, s->d_inv[a], s_out, outer, 2 * K, Ln, inner, !first, st));
    }
    first = false;
  }
  return FFNO_OK;
}

// Per-axis spectral outputs (no accumulation): s_axis[a] receives the inverse transform of axis a; the FF loader sums
// them.  3 launches per layer: all forward transforms, all mode mixes, all inverse transforms.
static bool all_axes_pipe(const UmmaState* s) {
  bool ok = true;
  for (int a = 0; a < s->d.ndim; ++a) ok &= (s->fwd_image[a] != nullptr) && (s->inv_image[a] != nullptr);
  return ok;
}

static int spectral_split(UmmaState* s, const UmmaLayer& L, const float* x, int batch, float* const s_axis[3], float* F,
                          float* R, cudaStream_t st) {
  AxisXform fwd[3], inv[3];
  MixAxis mix[3];
  const bool full = s->d.spectral_mode == FFNO_MODE_FULL;
  for (int a = 0; a < s->d.ndim; ++a) {
    long long outer, p_inner;
    axis_geom(s, batch, a, &outer, &p_inner);
    const int Ln = s->ext[a], K = s->d.modes[a];
    float* Fa = F + spec_offset(s, batch, a);
    float* Ra = R + spec_offset(s, batch, a);
    fwd[a] = AxisXform{x, Fa, s->fwd_image[a], outer, p_inner * kUmmaC, Ln, 2 * K, pad16i(2 * K), (Ln + 63) / 64, 0};
    mix[a] = MixAxis{Fa, Ra, L.mix_image[a], outer, p_inner, K};
    inv[a] = AxisXform{full ? Ra : Fa, s_axis[a], s->inv_image[a], outer, p_inner * kUmmaC, 2 * K, Ln, pad16i(Ln),
                       (2 * K + 63) / 64, 0};
  }
  // direction of the tile walk alternates launch to launch (fwd ->, mix <-, inv ->, FF <-): see umma_kernels.cuh
  const bool zz = s->zigzag;
  FFNO_TRY(launch_axis_pipe(fwd, s->d.ndim, s->sm_count, st, false));
  if (full) FFNO_TRY(launch_mix_pipe(mix, s->d.ndim, s->sm_count, st, zz));
  return launch_axis_pipe(inv, s->d.ndim, s->sm_count, st, false);
}

int umma_spectral_split_fwd(UmmaState* s, int layer, const float* x, int batch, float* const s_axis[3], float* F, float* R,
                            float* ws, int* n_written, cudaStream_t st) {
  if (all_axes_pipe(s)) {
    *n_written = s->d.ndim;
    return spectral_split(s, s->layers[layer], x, batch, s_axis, F, R, st);
  }
  *n_written = 1;
  return umma_spectral_fwd(s, layer, x, batch, s_axis[0], F, R, ws, st);
}

size_t umma_spec_offset(const UmmaState* s, int batch, int axis) { return spec_offset(s, batch, axis); }

int umma_forward_spectra(UmmaState* s, const float* x, int batch, float* F, cudaStream_t st) {
  AxisXform fwd[3];
  for (int a = 0; a < s->d.ndim; ++a) {
    long long outer, p_inner;
    axis_geom(s, batch, a, &outer, &p_inner);
    const int Ln = s->ext[a], K = s->d.modes[a];
    float* Fa = F + spec_offset(s, batch, a);
    fwd[a] = AxisXform{x, Fa, s->fwd_image[a], outer, p_inner * kUmmaC, Ln, 2 * K, pad16i(2 * K), (Ln + 63) / 64, 0};
    if (!all_axes_pipe(s)) {
      if (s->fwd_image[a]) FFNO_TRY(launch_axis_pipe(&fwd[a], 1, s->sm_count, st));
      else FFNO_TRY(launch_axis_transform(x, s->d_fwd[a], Fa, outer, Ln, 2 * K, p_inner * kUmmaC, false, st));
    }
  }
  if (all_axes_pipe(s)) return launch_axis_pipe(fwd, s->d.ndim, s->sm_count, st, false);
  return FFNO_OK;
}

int umma_ff_fwd(UmmaState* s, int layer, const float* s_in, const float* residual, int batch, float* y, float*,
                cudaStream_t st) {
  const UmmaLayer& L = s->layers[layer];
  long long P = batch;
  for (int a = 0; a < s->d.ndim; ++a) P *= s->ext[a];
  float* xo = residual ? y : nullptr;
  float* bo = residual ? nullptr : y;
  return launch_ff_ts(s_in, nullptr, nullptr, residual, xo, bo, L.ff_image, L.b1, L.b2, P, s->sm_count, st);
}

// FeedForward of a layer whose spectral output `s_in` was produced elsewhere (the FP32 rfft2 passes of an
// FFNO_TRANSFORM_RFFT2 plan): x_next = residual + FF(s_in) and / or b_out = FF(s_in).
int umma_ff_layer(UmmaState* s, int layer, const float* s_in, const float* residual, int batch, float* x_next, float* b_out,
                  cudaStream_t st) {
  const UmmaLayer& L = s->layers[layer];
  long long P = batch;
  for (int a = 0; a < s->d.ndim; ++a) P *= s->ext[a];
  return launch_ff_ts(s_in, nullptr, nullptr, x_next ? residual : nullptr, x_next, b_out, L.ff_image, L.b1, L.b2, P,
                      s->sm_count, st);
}

// ---- stage-pipelined forward ---------------------------------------------------------------------------------------
// Samples per unit: the smallest group whose tiles align in every stage (0 = this batch cannot be pipelined).
int umma_pipeline_unit(const UmmaState* s, int batch) {
  bool nopad = true;
  for (int a = 0; a < s->d.ndim; ++a) nopad &= s->d.pad[a] == 0;
  if (s->ff_only || !s->persist_env || !s->concurrent || !s->mix_shared || s->d.ndim != 2 || !all_axes_pipe(s) || !nopad ||
      s->d.out_features != 1 || s->d.use_fork || s->ff_layers_dev == nullptr || s->sm_count != s->hw_sm_count)
    return 0;
  int mix_ctas = 0;
  for (int a = 0; a < s->d.ndim; ++a) mix_ctas = s->d.modes[a] > mix_ctas ? s->d.modes[a] : mix_ctas;
  mix_ctas *= s->d.ndim;
  if (mix_ctas + 3 * s->d.ndim > s->hw_sm_count) return 0;
  for (int u = 1; u <= 8; u *= 2) {
    if (batch % u != 0 || batch / u < 4) continue;
    AxisXform fwd[3];
    MixAxis mix[3];
    long long P = batch;
    for (int a = 0; a < s->d.ndim; ++a) {
      long long outer, p_inner;
      axis_geom(s, batch, a, &outer, &p_inner);
      fwd[a] = AxisXform{nullptr, nullptr, nullptr, outer, p_inner * kUmmaC, s->ext[a], 2 * s->d.modes[a], 0, 0, 0};
      mix[a] = MixAxis{nullptr, nullptr, nullptr, outer, p_inner, s->d.modes[a]};
      P *= s->ext[a];
    }
    if (stack_pipe_geometry_ok(fwd, mix, s->d.ndim, P, batch / u)) return u;
  }
  return 0;
}

// The whole layer stack on the lifted activations in xa: four launches (forward transforms, mode mixes, inverse
// transforms, FeedForward + residual + fused head on the last layer), each walking every layer.
int umma_stack_fwd_pipelined(UmmaState* s, float* xa, float* xb, int batch, float* s_out, float* F, float* R, float* ws,
                             const UmmaFusedHead& head, cudaStream_t st) {
  const int u = umma_pipeline_unit(s, batch);
  FFNO_REQUIRE(u > 0, FFNO_ERR_STATE, "stage-pipelined forward requested for a batch that does not qualify");
  const int n_units = batch / u;
  if (s->counter_units < n_units) {
    uint8_t* raw = nullptr;
    FFNO_TRY(alloc_bytes(s, stack_pipe_counter_bytes(n_units), &raw));
    s->counters = reinterpret_cast<unsigned*>(raw);
    s->counter_units = n_units;
  }
  StackPipeArgs A{};
  A.n_layers = s->d.n_layers;
  A.n_axes = s->d.ndim;
  A.n_units = n_units;
  long long P = batch;
  size_t U = (size_t)batch * kUmmaC;
  for (int a = 0; a < s->d.ndim; ++a) { P *= s->ext[a]; U *= s->ext[a]; }
  float* s_axis[3] = {s_out, ws, ws + U};
  int maxK = 0;
  for (int a = 0; a < s->d.ndim; ++a) {
    long long outer, p_inner;
    axis_geom(s, batch, a, &outer, &p_inner);
    const int Ln = s->ext[a], K = s->d.modes[a];
    float* Fa = F + spec_offset(s, batch, a);
    float* Ra = R + spec_offset(s, batch, a);
    A.fwd[a] = AxisXform{xa, Fa, s->fwd_image[a], outer, p_inner * kUmmaC, Ln, 2 * K, pad16i(2 * K), (Ln + 63) / 64, 0};
    A.mix[a] = MixAxis{Fa, Ra, s->layers[0].mix_image[a], outer, p_inner, K};
    A.inv[a] = AxisXform{Ra, s_axis[a], s->inv_image[a], outer, p_inner * kUmmaC, 2 * K, Ln, pad16i(Ln), (2 * K + 63) / 64, 0};
    maxK = K > maxK ? K : maxK;
  }
  A.x_odd = xb;
  A.xbuf[0] = xa;
  A.xbuf[1] = xb;
  A.ff_layers = s->ff_layers_dev;
  A.head_w = head.w;
  A.head_b = head.b;
  A.forecast = head.forecast;
  A.P = P;
  A.counters = s->counters;
  A.dbg = s->pipe_dbg;
  {
    const char* os = getenv("FFNO_B200_PIPE_ONLY");      // diagnostics (tools/pipe_stage_alone.py); results are garbage
    A.only_stage = os ? atoi(os) : -1;
  }
  if (s->pipe_sms[0] > 0 && s->pipe_sms[2] > 0 && s->pipe_sms[3] > 0) {
    for (int i = 0; i < 4; ++i) A.sms[i] = s->pipe_sms[i];
  } else {
    // One CTA per (axis, mode) mixes; the other SMs are shared in proportion to the stages' measured steady-state
    // tile costs inside the pipeline (B200, 64 x 64 grid, tools/pipe_stats.py: 0.95 us per forward-transform tile,
    // 1.4 us per inverse-transform tile, 2.5 us per FF tile — DESIGN.md §5).
    const int mix_ctas = maxK * s->d.ndim, rest = s->hw_sm_count - mix_ctas, nd = s->d.ndim;
    double tiles_ax = 0.0;
    for (int a = 0; a < nd; ++a) tiles_ax += (double)(A.fwd[a].outer * (A.fwd[a].inner / 64) / 2);
    const double w_f = 0.95 * tiles_ax, w_i = 1.4 * tiles_ax, w_ff = 2.5 * (double)(P / 128), tot = w_f + w_i + w_ff;
    int f = (int)(rest * w_f / tot + 0.5) / nd * nd, i = (int)(rest * w_i / tot + 0.5) / nd * nd;
    if (f < nd) f = nd;
    if (i < nd) i = nd;
    A.sms[0] = f;
    A.sms[1] = mix_ctas;
    A.sms[2] = i;
    A.sms[3] = rest - f - i;
  }
  FFNO_REQUIRE(A.sms[0] + maxK * s->d.ndim + A.sms[2] + A.sms[3] <= s->hw_sm_count && A.sms[3] >= 1, FFNO_ERR_BAD_ARG,
               "stage-pipelined forward: %d + %d + %d + %d CTAs exceed the %d SMs (all stages must be resident at once)",
               A.sms[0], maxK * s->d.ndim, A.sms[2], A.sms[3], s->hw_sm_count);
  A.streams[0] = st;
  for (int i = 0; i < 3; ++i) { A.streams[i + 1] = s->pstream[i]; A.join[i] = s->pjoin[i]; }
  A.fork = s->pfork;
  return launch_stack_pipe(A);
}

int umma_pipe_debug(const UmmaState* s, unsigned long long* host_out, int n_words) {
  FFNO_REQUIRE(s->pipe_dbg != nullptr, FFNO_ERR_STATE, "plan was not created with FFNO_B200_PIPE_DEBUG=1");
  const int n = n_words < kPipeDebugWords ? n_words : kPipeDebugWords;
  FFNO_CUDA_CHECK(cudaMemcpy(host_out, s->pipe_dbg, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return n;
}

bool umma_can_fuse_head(const UmmaState* s, bool want_s) {
  bool nopad = true;
  for (int a = 0; a < s->d.ndim; ++a) nopad &= s->d.pad[a] == 0;
  return !want_s && all_axes_pipe(s) && s->d.out_features == 1 && nopad && !s->d.use_fork;
}

int umma_layer_fwd(UmmaState* s, int layer, const float* x, int batch, float* x_next, float* s_out, float* b_out,
                   float* F, float* R, float* ws, bool want_s, bool want_b, cudaStream_t st, const UmmaFusedHead* head) {
  const UmmaLayer& L = s->layers[layer];
  long long P = batch;
  for (int a = 0; a < s->d.ndim; ++a) P *= s->ext[a];
  float* bo = want_b ? b_out : nullptr;
  if (!want_s && all_axes_pipe(s)) {
    // fast path: per-axis inverse outputs s_out (axis 0) and ws + (a-1) * U (axes 1, 2), summed by the FF loader
    const size_t U = (size_t)P * kUmmaC;
    float* s_axis[3] = {s_out, ws, ws + U};
    FFNO_TRY(spectral_split(s, L, x, batch, s_axis, F, R, st));
    return launch_ff_ts(s_axis[0], s->d.ndim > 1 ? s_axis[1] : nullptr, s->d.ndim > 2 ? s_axis[2] : nullptr,
                        x_next ? x : nullptr, x_next, bo,
                        L.ff_image, L.b1, L.b2, P, s->sm_count, st, head ? head->w : nullptr, head ? head->b : nullptr,
                        head ? head->forecast : nullptr, s->zigzag);
  }
  FFNO_TRY(umma_spectral_fwd(s, layer, x, batch, s_out, F, R, ws, st));
  return launch_ff_ts(s_out, nullptr, nullptr, x, x_next, bo, L.ff_image, L.b1, L.b2, P, s->sm_count, st);
}

}  // namespace ffno
