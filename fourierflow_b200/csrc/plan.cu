// libffno_b200: plan management and the extern "C" entry points declared in include/ffno_b200.h.
//
// A plan owns (a) the truncated real-DFT tables of every axis, built on the host in double precision,
// (b) the prepared parameters: weight-norm folded + transposed linears, per-mode real block matrices of
// the spectral weights, the pre-multiplied head, and (UMMA path) their bf16 hi/lo operand tiles.
// The forward enqueues kernels on the caller's stream and never synchronises.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <vector>

#include "backward.cuh"
#include "generic_kernels.cuh"
#include "umma_kernels.cuh"
#include "umma_path.cuh"

namespace ffno {

extern thread_local long long g_launch_counter;

char* error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}
int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

int ensure_dynamic_smem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> granted;
  int dev = 0;
  FFNO_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = granted[std::make_pair(func, dev)];
  if (bytes > have) {
    FFNO_CUDA_CHECK(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    have = bytes;
  }
  return FFNO_OK;
}

struct Lin {
  float* wt = nullptr;     // [in][out] folded, transposed
  float* bias = nullptr;   // [out] (copy) or nullptr
  int in = 0, out = 0;
};
struct FFW {
  Lin lin[FFNO_MAX_FF_LAYERS];
  float* ln_w = nullptr;
  float* ln_b = nullptr;
};
struct LayerW {
  float* wmix[FFNO_MAX_DIMS] = {nullptr, nullptr, nullptr};   // [K][2C][2C]
  FFW back, fork;
};

}  // namespace ffno

using namespace ffno;

struct ffno_plan {
  ffno_desc d;
  int ext[3] = {1, 1, 1};          // size + pad
  long long pts = 0;               // padded points per sample
  long long pts_in = 0;            // input points per sample
  int in_total = 0;                // in_features + appended grid channels
  bool use_umma = false;
  bool ff_umma = false;      // FFNO_TRANSFORM_RFFT2: spectral layer on the FP32 kernels, FeedForward on ff_ts_kernel
  float* d_fwd[3] = {nullptr, nullptr, nullptr};   // [L][ld(2K)]
  float* d_inv[3] = {nullptr, nullptr, nullptr};   // [2K][ld(L)]
  float* d_fwdT[3] = {nullptr, nullptr, nullptr};  // [2K][ld(L)]: transpose of d_fwd — adjoint of the forward transform
  float* d_invT[3] = {nullptr, nullptr, nullptr};  // [L][ld(2K)]: transpose of d_inv — adjoint of the inverse transform
  bool loaded = false;
  bool has_io = false;             // lift + head parameters were given (false for a bare spectral layer)
  std::vector<LayerW> layers;
  Lin lift;
  float* head_w = nullptr;         // [out][C]
  float* head_b = nullptr;         // [out]
  Lin out0, out1;                  // scratch for the head fold
  std::vector<void*> owned;        // every cudaMalloc of this plan
  std::map<const void*, float*> dedup;   // source pointer -> prepared buffer (shared weights)
  int64_t last_launches = 0;
  UmmaState* umma = nullptr;
  // Backward on the tcgen05 kernels: the adjoint of the spectral operator is the SAME three kernels with the transposed
  // tables (d_invT in the forward-transform role, d_fwdT in the inverse role) and the transposed mode blocks — a second
  // kernel state built on first use and refreshed when the parameters change.
  UmmaState* umma_adj = nullptr;
  bool adj_stale = true;
  int dct_modes[3] = {0, 0, 0};      // FFNO_TRANSFORM_DCT: coefficients kept per axis (d.modes then holds PAIRS of them)
  bool bwd_fp32_recompute = false, bwd_fp32_adjoint = false;
  bool bwd_fp32 = false;                     // FFNO_B200_BWD=fp32: backward entirely on the FP32 kernels (see ffno_block_bwd)
  std::map<const float*, float*> wmixT;      // forward block matrices -> their per-mode transposes

  // CUDA-graph replay of the launch sequence (kills ~120 launch gaps per forward).  A slot is keyed by everything
  // baked into the captured kernel arguments; the first call with a new key runs eagerly, the second captures.
  struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    int batch = -1, n_steps = 0, seen = 0, variant = 0;
    void* ws = nullptr;
    MeanStd ms{};
    float low = 0.f, high = 0.f;
    int64_t launches = 0;
    void reset() {
      if (exec) cudaGraphExecDestroy(exec);
      exec = nullptr;
      batch = -1;
      seen = 0;
    }
  };
  GraphSlot g_block, g_rollout;
  bool graphs = true;
  int graph_failures = 0;          // refused captures; graphs are given up after kMaxGraphFailures of them
  // Capture and replay need a capturable stream.  The legacy default stream (handle 0 — what
  // torch.cuda.current_stream() is unless the caller opened a side stream) is not, so work arriving on it runs on this
  // plan-owned non-blocking stream, fenced against the caller's stream with an event on either side.
  cudaStream_t own_stream = nullptr;
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  float domain[2] = {6.283185307179586f, 6.283185307179586f};   // periodic domain lengths of the velocity features

  // Batch chunking of the stack forward (samples are independent): the batch is cut into chunks of
  // `chunk` samples that run the whole layer stack one after the other on the same (small, L2-resident) workspace
  // slot, or — with n_streams > 1 — side by side on forked streams, each kernel limited to 1/n_streams of the SMs so
  // that one chunk's pipeline fill / drain overlaps the other's steady state.  0 = off.
  int chunk = 0;
  int n_streams = 1;
  int sm_count = 148;
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[3] = {nullptr, nullptr, nullptr};
  int chunk_for(int batch) const {
    if (!use_umma) return 0;
    int c = chunk;
    if (c <= 0 && n_streams > 1) c = (batch + n_streams - 1) / n_streams;
    return (c > 0 && c < batch) ? c : 0;
  }

  LiftGeom geom() const {
    LiftGeom g;
    g.ndim = d.ndim;
    for (int a = 0; a < 3; ++a) { g.size[a] = a < d.ndim ? d.size[a] : 1; g.pad[a] = a < d.ndim ? d.pad[a] : 0; }
    g.in_features = d.in_features;
    g.append_grid = d.append_grid;
    g.C = d.width;
    return g;
  }
  int hidden() const { return d.width * d.ff_factor; }
};

namespace {

int dev_alloc(ffno_plan* p, size_t bytes, float** out) {
  void* ptr = nullptr;
  FFNO_CUDA_CHECK(cudaMalloc(&ptr, bytes ? bytes : 4));
  p->owned.push_back(ptr);
  *out = static_cast<float*>(ptr);
  return FFNO_OK;
}

int pad16(int n) { return (n + 15) / 16 * 16; }

// Truncated ortho real DFT tables (double precision on the host).  See oracle/ffno_oracle.py
// dft_forward_matrix / dft_inverse_matrix for the independent statement the CPU tests pin to torch.fft:
//   forward  T[l][2k]   =  cos(2 pi k l / L) / sqrt(L),  T[l][2k+1] = -sin(2 pi k l / L) / sqrt(L)
//   inverse  T[2k][l]   =  c_k cos(..) / sqrt(L),        T[2k+1][l] = -c_k sin(..) / sqrt(L)
// c_0 = 1, c_k = 2 (Hermitian half), c_{L/2} = 1 for even L; sin(0) = sin(pi l) = 0 drops Im(DC), Im(Nyquist)
// exactly as the C2R transform at grid_2d.py:72 does.
// FFNO_TRANSFORM_RFFT2, row axis (complex-to-complex over the 2K retained rows kx = 0..K-1, L-K..L-1): cosine and sine
// tables side by side, combined by launch_c2c_combine.   forward [L][4K]: T[l][j] = cos(th_jl)/sqrt(L), T[l][2K + j] = sin;
// inverse [2K][2L]: T[j][l] = cos(th_jl)/sqrt(L), T[j][L + l] = sin, th_jl = 2 pi kx_j l / L.
int build_c2c_tables(ffno_plan* p) {
  const int L = p->ext[0], K = p->d.modes[0], R = 2 * K;
  const int ldf = pad16(2 * R), ldi = pad16(2 * L);
  std::vector<float> f((size_t)L * ldf, 0.f), inv((size_t)R * ldi, 0.f);
  const int ldft = pad16(L), ldit = pad16(R);                       // their transposes (backward pass)
  std::vector<float> fT((size_t)2 * R * ldft, 0.f), invT((size_t)2 * L * ldit, 0.f);
  const double s = 1.0 / std::sqrt((double)L);
  for (int l = 0; l < L; ++l)
    for (int j = 0; j < R; ++j) {
      const int kx = j < K ? j : L - R + j;
      const double ang = 2.0 * M_PI * (double)(((long long)kx * l) % L) / (double)L;
      const double c = std::cos(ang) * s, sn = std::sin(ang) * s;
      f[(size_t)l * ldf + j] = (float)c;
      f[(size_t)l * ldf + R + j] = (float)sn;
      inv[(size_t)j * ldi + l] = (float)c;
      inv[(size_t)j * ldi + L + l] = (float)sn;
      fT[(size_t)j * ldft + l] = (float)c;
      fT[(size_t)(R + j) * ldft + l] = (float)sn;
      invT[(size_t)l * ldit + j] = (float)c;
      invT[(size_t)(L + l) * ldit + j] = (float)sn;
    }
  FFNO_TRY(dev_alloc(p, f.size() * 4, &p->d_fwd[0]));
  FFNO_TRY(dev_alloc(p, inv.size() * 4, &p->d_inv[0]));
  FFNO_TRY(dev_alloc(p, fT.size() * 4, &p->d_fwdT[0]));
  FFNO_TRY(dev_alloc(p, invT.size() * 4, &p->d_invT[0]));
  FFNO_CUDA_CHECK(cudaMemcpy(p->d_fwd[0], f.data(), f.size() * 4, cudaMemcpyHostToDevice));
  FFNO_CUDA_CHECK(cudaMemcpy(p->d_inv[0], inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
  FFNO_CUDA_CHECK(cudaMemcpy(p->d_fwdT[0], fT.data(), fT.size() * 4, cudaMemcpyHostToDevice));
  FFNO_CUDA_CHECK(cudaMemcpy(p->d_invT[0], invT.data(), invT.size() * 4, cudaMemcpyHostToDevice));
  return FFNO_OK;
}

int build_tables(ffno_plan* p) {
  for (int a = 0; a < p->d.ndim; ++a) {
    if (p->d.transform == FFNO_TRANSFORM_RFFT2 && a == 0) {
      FFNO_TRY(build_c2c_tables(p));
      continue;
    }
    const int L = p->ext[a], K = p->d.modes[a];
    const int ldf = pad16(2 * K), ldi = pad16(L);
    std::vector<float> f((size_t)L * ldf, 0.f), inv((size_t)2 * K * ldi, 0.f);
    std::vector<float> fT((size_t)2 * K * ldi, 0.f), invT((size_t)L * ldf, 0.f);      // backward pass (ffno_block_bwd)
    const double s = 1.0 / std::sqrt((double)L);
    if (p->d.transform == FFNO_TRANSFORM_DCT) {
      // Ortho DCT-II basis (modules/dct.py:16-45 with norm='ortho'): X_j = c_j sum_l x_l cos(pi (2 l + 1) j / (2 L)),
      // c_0 = sqrt(1/L), c_j = sqrt(2/L); the inverse (DCT-III, dct.py:48-88) is the transpose.  Coefficients 2k and
      // 2k + 1 ride in the (Re, Im) rows of "mode" k, so every kernel downstream is the F-FNO one: the mix of pair k is
      // the block-diagonal [[W_2k, 0], [0, W_2k+1]] (pack_mix_weights_dct).  Rows past the kept count stay zero.
      for (int l = 0; l < L; ++l)
        for (int j = 0; j < p->dct_modes[a]; ++j) {
          const long long num = ((long long)(2 * l + 1) * j) % (4ll * L);      // angle = pi * num / (2 L), period 4 L
          const double v = std::cos(M_PI * (double)num / (2.0 * L)) * (j == 0 ? s : s * std::sqrt(2.0));
          f[(size_t)l * ldf + j] = (float)v;
          invT[(size_t)l * ldf + j] = (float)v;
          inv[(size_t)j * ldi + l] = (float)v;
          fT[(size_t)j * ldi + l] = (float)v;
        }
    } else
    for (int l = 0; l < L; ++l)
      for (int k = 0; k < K; ++k) {
        // reduce k*l mod L before the multiply: keeps the angle small and exact for large L
        long long kl = ((long long)k * l) % L;
        double ang = 2.0 * M_PI * (double)kl / (double)L;
        double c = std::cos(ang) * s, sn = std::sin(ang) * s;
        double ck = (k == 0) ? 1.0 : ((L % 2 == 0 && k == L / 2) ? 1.0 : 2.0);
        f[(size_t)l * ldf + 2 * k] = (float)c;
        f[(size_t)l * ldf + 2 * k + 1] = (float)(-sn);
        inv[(size_t)(2 * k) * ldi + l] = (float)(ck * c);
        inv[(size_t)(2 * k + 1) * ldi + l] = (float)(-ck * sn);
        fT[(size_t)(2 * k) * ldi + l] = (float)c;
        fT[(size_t)(2 * k + 1) * ldi + l] = (float)(-sn);
        invT[(size_t)l * ldf + 2 * k] = (float)(ck * c);
        invT[(size_t)l * ldf + 2 * k + 1] = (float)(-ck * sn);
      }
    FFNO_TRY(dev_alloc(p, f.size() * 4, &p->d_fwd[a]));
    FFNO_TRY(dev_alloc(p, inv.size() * 4, &p->d_inv[a]));
    FFNO_CUDA_CHECK(cudaMemcpy(p->d_fwd[a], f.data(), f.size() * 4, cudaMemcpyHostToDevice));
    FFNO_CUDA_CHECK(cudaMemcpy(p->d_inv[a], inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
    FFNO_TRY(dev_alloc(p, fT.size() * 4, &p->d_fwdT[a]));
    FFNO_TRY(dev_alloc(p, invT.size() * 4, &p->d_invT[a]));
    FFNO_CUDA_CHECK(cudaMemcpy(p->d_fwdT[a], fT.data(), fT.size() * 4, cudaMemcpyHostToDevice));
    FFNO_CUDA_CHECK(cudaMemcpy(p->d_invT[a], invT.data(), invT.size() * 4, cudaMemcpyHostToDevice));
  }
  return FFNO_OK;
}

int validate_desc(const ffno_desc* d) {
  FFNO_REQUIRE(d != nullptr, FFNO_ERR_BAD_ARG, "desc is NULL");
  FFNO_REQUIRE(d->abi_version == FFNO_ABI_VERSION, FFNO_ERR_BAD_ARG, "ABI version %d != %d", d->abi_version,
               FFNO_ABI_VERSION);
  FFNO_REQUIRE(d->ndim == 2 || d->ndim == 3, FFNO_ERR_UNSUPPORTED, "ndim=%d (need 2 or 3)", d->ndim);
  FFNO_REQUIRE(d->width > 0 && d->width % 4 == 0, FFNO_ERR_UNSUPPORTED, "width=%d must be a positive multiple of 4",
               d->width);
  FFNO_REQUIRE(d->n_layers >= 1, FFNO_ERR_BAD_ARG, "n_layers=%d", d->n_layers);
  FFNO_REQUIRE(d->n_ff_layers >= 1 && d->n_ff_layers <= FFNO_MAX_FF_LAYERS, FFNO_ERR_UNSUPPORTED,
               "n_ff_layers=%d not in [1,%d]", d->n_ff_layers, FFNO_MAX_FF_LAYERS);
  FFNO_REQUIRE(d->ff_factor >= 1, FFNO_ERR_BAD_ARG, "ff_factor=%d", d->ff_factor);
  FFNO_REQUIRE(d->in_features >= 1 && d->out_features >= 1 && d->out_features <= 8, FFNO_ERR_UNSUPPORTED,
               "in_features=%d out_features=%d", d->in_features, d->out_features);
  FFNO_REQUIRE(d->head_hidden >= 1, FFNO_ERR_BAD_ARG, "head_hidden=%d", d->head_hidden);
  FFNO_REQUIRE(d->spectral_mode >= 0 && d->spectral_mode <= 2, FFNO_ERR_BAD_ARG, "spectral_mode=%d", d->spectral_mode);
  FFNO_REQUIRE(d->transform >= FFNO_TRANSFORM_RFFT && d->transform <= FFNO_TRANSFORM_RFFT2, FFNO_ERR_BAD_ARG,
               "transform=%d", d->transform);
  FFNO_REQUIRE(d->transform != FFNO_TRANSFORM_DCT || d->spectral_mode == FFNO_MODE_FULL, FFNO_ERR_UNSUPPORTED,
               "the DCT variant has no low-pass / no-fourier mode");
  if (d->transform == FFNO_TRANSFORM_RFFT2) {
    // zongyi_fno/grid_plus_2d.py:52-83: one mode count for both axes; 'low-pass' is a bare `raise` there (:74-75)
    FFNO_REQUIRE(d->ndim == 2, FFNO_ERR_UNSUPPORTED, "the rfft2 (FNOPlus2DBlock) variant is 2-D, ndim=%d", d->ndim);
    FFNO_REQUIRE(d->spectral_mode != FFNO_MODE_LOW_PASS, FFNO_ERR_UNSUPPORTED, "the rfft2 variant has no low-pass mode");
    FFNO_REQUIRE(d->modes[0] == d->modes[1], FFNO_ERR_BAD_ARG, "rfft2: modes %d != %d", d->modes[0], d->modes[1]);
    FFNO_REQUIRE(2 * d->modes[0] <= d->size[0] + d->pad[0], FFNO_ERR_BAD_ARG,
                 "rfft2: the two %d-row blocks overlap on a %d-row grid", d->modes[0], d->size[0] + d->pad[0]);
    FFNO_REQUIRE(d->path != FFNO_PATH_UMMA, FFNO_ERR_UNSUPPORTED, "the rfft2 variant runs on the FP32 kernels only");
  }
  for (int a = 0; a < d->ndim; ++a) {
    FFNO_REQUIRE(d->size[a] >= 1 && d->pad[a] >= 0, FFNO_ERR_BAD_ARG, "size[%d]=%d pad=%d", a, d->size[a], d->pad[a]);
    int L = d->size[a] + d->pad[a];
    // the reference's slice-assign `out_ft[..., :modes] = einsum(x_ft[..., :modes], W)` raises when
    // modes > L//2+1 — same contract here.
    if (d->transform == FFNO_TRANSFORM_DCT)
      FFNO_REQUIRE(d->modes[a] >= 1 && d->modes[a] <= L, FFNO_ERR_BAD_ARG,
                   "modes[%d]=%d exceeds the %d DCT coefficients of a length-%d axis", a, d->modes[a], L, L);
    else
    FFNO_REQUIRE(d->modes[a] >= 1 && d->modes[a] <= L / 2 + 1, FFNO_ERR_BAD_ARG,
                 "modes[%d]=%d exceeds the %d rfft bins of a length-%d axis", a, d->modes[a], L / 2 + 1, L);
  }
  return FFNO_OK;
}

// ---- parameter preparation ---------------------------------------------------------------------------
int prep_linear(ffno_plan* p, const ffno_linear_params& src, int in, int out, Lin* dst, cudaStream_t st) {
  FFNO_REQUIRE(src.in_features == in && src.out_features == out, FFNO_ERR_BAD_ARG,
               "linear shape [%d,%d] given, [%d,%d] expected", src.out_features, src.in_features, out, in);
  const float* v = src.weight ? src.weight : src.weight_v;
  const float* g = src.weight ? nullptr : src.weight_g;
  FFNO_REQUIRE(v != nullptr, FFNO_ERR_BAD_ARG, "linear has neither weight nor weight_v");
  FFNO_REQUIRE(src.weight || src.weight_g, FFNO_ERR_BAD_ARG, "weight_v without weight_g");
  dst->in = in;
  dst->out = out;
  if (!dst->wt) FFNO_TRY(dev_alloc(p, (size_t)in * out * 4, &dst->wt));
  FFNO_TRY(launch_weight_fold_transpose(v, g, dst->wt, out, in, st));
  if (src.bias) {
    if (!dst->bias) FFNO_TRY(dev_alloc(p, (size_t)out * 4, &dst->bias));
    FFNO_CUDA_CHECK(cudaMemcpyAsync(dst->bias, src.bias, (size_t)out * 4, cudaMemcpyDeviceToDevice, st));
  } else {
    dst->bias = nullptr;
  }
  return FFNO_OK;
}

int prep_ff(ffno_plan* p, const ffno_ff_params& src, FFW* dst, cudaStream_t st) {
  const int C = p->d.width, H = p->hidden(), n = p->d.n_ff_layers;
  for (int i = 0; i < n; ++i) {
    int in = (i == 0) ? C : H, out = (i == n - 1) ? C : H;     // feedforward.py:11-12
    FFNO_TRY(prep_linear(p, src.linear[i], in, out, &dst->lin[i], st));
  }
  if (p->d.layer_norm) {
    FFNO_REQUIRE(src.ln_weight && src.ln_bias, FFNO_ERR_BAD_ARG, "layer_norm set but LayerNorm params missing");
    if (!dst->ln_w) FFNO_TRY(dev_alloc(p, (size_t)C * 4, &dst->ln_w));
    if (!dst->ln_b) FFNO_TRY(dev_alloc(p, (size_t)C * 4, &dst->ln_b));
    FFNO_CUDA_CHECK(cudaMemcpyAsync(dst->ln_w, src.ln_weight, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
    FFNO_CUDA_CHECK(cudaMemcpyAsync(dst->ln_b, src.ln_bias, (size_t)C * 4, cudaMemcpyDeviceToDevice, st));
  }
  return FFNO_OK;
}

// ---- workspace carving ---------------------------------------------------------------------------------
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  float* take(size_t n_floats) {
    float* r = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += (n_floats * 4 + 255) / 256 * 256;
    return r;
  }
};

struct Workspace {
  float *xa, *xb, *s, *b, *h0, *h1, *F, *R, *tmp, *f, *umma;
  float *io_in, *io_out;      // fixed-address staging of the stack input / forecast (CUDA-graph replay, host entry)
  size_t bytes;
};

Workspace carve(const ffno_plan* p, int batch, void* base) {
  Workspace w{};
  Carver c(base);
  const long long P = (long long)batch * p->pts;
  const size_t U = (size_t)P * p->d.width;
  w.xa = c.take(U);
  w.xb = c.take(U);
  w.s = c.take(U);
  w.b = c.take(U);
  size_t spec = 0;
  for (int a = 0; a < p->d.ndim; ++a) {
    spec += U / p->ext[a] * 2 * p->d.modes[a];     // every axis' spectra live side by side
  }
  if (p->d.transform == FFNO_TRANSFORM_RFFT2)       // largest intermediate: cos | sin sums of the inverse row transform
    spec = (size_t)batch * 2 * p->ext[0] * 2 * p->d.modes[1] * p->d.width;
  w.F = c.take(spec);
  w.R = c.take(spec);
  if (!p->use_umma || p->d.n_ff_layers != 2) {
    w.h0 = c.take(p->d.n_ff_layers > 1 ? (size_t)P * p->hidden() : 0);
    w.h1 = c.take(p->d.n_ff_layers > 2 ? (size_t)P * p->hidden() : 0);
  }
  w.tmp = c.take(p->d.layer_norm ? U : 0);
  w.f = c.take(p->d.use_fork ? U : 0);
  w.umma = c.take(p->use_umma ? umma_workspace_floats(p->umma, batch) : 0);
  w.io_in = c.take((size_t)batch * p->pts_in * p->d.in_features);
  w.io_out = c.take((size_t)batch * p->pts_in * p->d.out_features);
  w.bytes = c.off;
  return w;
}

// ---- generic forward pieces ---------------------------------------------------------------------------
// rfft2 -> two K x K corner blocks -> per-(kx, ky) channel mix -> irfft2 (zongyi_fno/grid_plus_2d.py:52-83) as 1-D passes:
// real DFT along the columns (the F-FNO table), complex DFT along the rows (cos | sin tables + combine), and back.
int spectral_rfft2(ffno_plan* p, const LayerW& lw, const float* x, int batch, float* s, float* F, float* R, cudaStream_t st) {
  const int C = p->d.width, M = p->ext[0], N = p->ext[1], K = p->d.modes[1], Rr = 2 * K;
  const long long row = (long long)2 * K * C;                                        // one (ky, ri, c) row of spectra
  FFNO_TRY(launch_axis_transform(x, p->d_fwd[1], F, (long long)batch * M, N, 2 * K, C, false, st));       // F: [B][M][K][2][C]
  FFNO_TRY(launch_axis_transform(F, p->d_fwd[0], R, batch, M, 2 * Rr, row, false, st));                    // R: [B][4K][K][2][C]
  FFNO_TRY(launch_c2c_combine(R, F, batch, Rr, K, C, 1.f, st));                                            // F: [B][2K][K][2][C]
  FFNO_TRY(launch_mode_mix(F, lw.wmix[0], R, batch, Rr * K, 1, C, st));                                    // R: same layout
  FFNO_TRY(launch_axis_transform(R, p->d_inv[0], F, batch, Rr, 2 * M, row, false, st));                    // F: [B][2M][K][2][C]
  FFNO_TRY(launch_c2c_combine(F, R, batch, M, K, C, -1.f, st));                                            // R: [B][M][K][2][C]
  return launch_axis_transform(R, p->d_inv[1], s, (long long)batch * M, 2 * K, N, C, false, st);
}

int spectral_generic(ffno_plan* p, const LayerW& lw, const float* x, int batch, float* s, float* F, float* R,
                     cudaStream_t st) {
  if (p->d.transform == FFNO_TRANSFORM_RFFT2) return spectral_rfft2(p, lw, x, batch, s, F, R, st);
  const int C = p->d.width;
  bool first = true;
  for (int a = p->d.ndim - 1; a >= 0; --a) {     // reference order: last axis first (grid_2d.py:57,75)
    long long outer = batch, p_inner = 1;
    for (int i = 0; i < a; ++i) outer *= p->ext[i];
    for (int i = a + 1; i < p->d.ndim; ++i) p_inner *= p->ext[i];
    const int L = p->ext[a], K = p->d.modes[a];
    const long long inner = p_inner * C;
    FFNO_TRY(launch_axis_transform(x, p->d_fwd[a], F, outer, L, 2 * K, inner, false, st));
    const float* src = F;
    if (p->d.spectral_mode == FFNO_MODE_FULL) {
      FFNO_TRY(launch_mode_mix(F, lw.wmix[a], R, outer, K, p_inner, C, st));
      src = R;
    }
    FFNO_TRY(launch_axis_transform(src, p->d_inv[a], s, outer, 2 * K, L, inner, !first, st));
    first = false;
  }
  return FFNO_OK;
}

// y = FF(s); if residual: out = residual + y, b_out = y (either may be NULL)
int ff_generic(ffno_plan* p, const FFW& ff, const float* s, const float* residual, long long P, float* out,
               float* b_out, const Workspace& w, cudaStream_t st) {
  const int n = p->d.n_ff_layers, C = p->d.width;
  const float* cur = s;
  float* hbuf[2] = {w.h0, w.h1};
  for (int i = 0; i < n - 1; ++i) {
    float* dst = hbuf[i & 1];
    FFNO_TRY(launch_linear(cur, ff.lin[i].wt, ff.lin[i].bias, nullptr, dst, nullptr, P, ff.lin[i].in,
                           ff.lin[i].out, true, st));
    cur = dst;
  }
  const Lin& last = ff.lin[n - 1];
  if (!p->d.layer_norm) {
    FFNO_TRY(launch_linear(cur, last.wt, last.bias, residual, out, b_out, P, last.in, last.out, false, st));
    if (!residual && !out) return FFNO_OK;
  } else {
    FFNO_TRY(launch_linear(cur, last.wt, last.bias, nullptr, w.tmp, nullptr, P, last.in, last.out, false, st));
    FFNO_TRY(launch_layernorm_residual(w.tmp, ff.ln_w, ff.ln_b, residual, out, b_out, P, C, st));
  }
  return FFNO_OK;
}

int check_ready(const ffno_plan* p, int batch, const void* ws, size_t ws_bytes, size_t need) {
  FFNO_REQUIRE(p != nullptr, FFNO_ERR_BAD_ARG, "plan is NULL");
  FFNO_REQUIRE(p->loaded, FFNO_ERR_STATE, "ffno_plan_load_params has not been called");
  FFNO_REQUIRE(batch >= 0, FFNO_ERR_BAD_ARG, "batch=%d", batch);
  FFNO_REQUIRE(ws != nullptr || need == 0, FFNO_ERR_WORKSPACE, "workspace is NULL");
  FFNO_REQUIRE(ws_bytes >= need, FFNO_ERR_WORKSPACE, "workspace %zu B < required %zu B", ws_bytes, need);
  FFNO_REQUIRE(((uintptr_t)ws & 255) == 0, FFNO_ERR_WORKSPACE, "workspace must be 256-byte aligned");
  return FFNO_OK;
}

int block_fwd_core(ffno_plan* p, const float* x, int batch, float* forecast, const ffno_taps* taps,
                   void* workspace, cudaStream_t st) {
  const Workspace w = carve(p, batch, workspace);
  const LiftGeom g = p->geom();
  const long long P = (long long)batch * p->pts;
  const size_t Ubytes = (size_t)P * p->d.width * 4;
  const int nl = p->d.n_layers;
  if (batch == 0) return FFNO_OK;

  FFNO_TRY(launch_lift(x, p->lift.wt, p->lift.bias, w.xa, batch, g, st));
  if (taps && taps->lift) FFNO_CUDA_CHECK(cudaMemcpyAsync(taps->lift, w.xa, Ubytes, cudaMemcpyDeviceToDevice, st));

  if (p->use_umma && !taps && umma_pipeline_unit(p->umma, batch) > 0) {
    // every layer in four stage-pipelined launches (umma_pipelined.cu: PipeDesc); the last FF applies the head
    UmmaFusedHead fh;
    fh.w = p->head_w;
    fh.b = p->head_b;
    fh.forecast = forecast;
    return umma_stack_fwd_pipelined(p->umma, w.xa, w.xb, batch, w.s, w.F, w.R, w.umma, fh, st);
  }

  float* cur = w.xa;
  float* nxt = w.xb;
  bool head_fused = false;
  for (int l = 0; l < nl; ++l) {
    const LayerW& lw = p->layers[l];
    const bool last = (l == nl - 1);
    if (p->use_umma) {
      const bool want_s = (taps && taps->spectral && taps->spectral[l]) || p->d.use_fork;
      if (last && !taps && umma_can_fuse_head(p->umma, want_s)) {
        // last layer: the FF store warps apply the folded head themselves — no `b` round trip, no head launch
        UmmaFusedHead fh;
        fh.w = p->head_w;
        fh.b = p->head_b;
        fh.forecast = forecast;
        // (the residual stream is not needed after the last layer: no x_next store either)
        FFNO_TRY(umma_layer_fwd(p->umma, l, cur, batch, nullptr, w.s, w.b, w.F, w.R, w.umma, false, false, st, &fh));
        head_fused = true;
      } else {
        // the residual stream is dead after the last layer (the head reads b): only a tap asks for it there
        const bool want_x = !last || (taps && taps->x_after && taps->x_after[l]);
        FFNO_TRY(umma_layer_fwd(p->umma, l, cur, batch, want_x ? nxt : nullptr, w.s, w.b, w.F, w.R, w.umma, want_s,
                                last || p->d.use_fork, st));
      }
    } else {
      const float* s = cur;
      if (p->d.spectral_mode != FFNO_MODE_NO_FOURIER) {
        FFNO_TRY(spectral_generic(p, lw, cur, batch, w.s, w.F, w.R, st));
        s = w.s;
      }
      if (p->ff_umma) FFNO_TRY(umma_ff_layer(p->umma, l, s, cur, batch, nxt, (last || taps) ? w.b : nullptr, st));
      else FFNO_TRY(ff_generic(p, lw.back, s, cur, P, nxt, w.b, w, st));
    }
    if (p->d.use_fork) {
      const float* s = (p->d.spectral_mode != FFNO_MODE_NO_FOURIER) ? w.s : cur;
      FFNO_TRY(ff_generic(p, lw.fork, s, nullptr, P, nullptr, w.f, w, st));
      if (taps && taps->forecast_list && taps->forecast_list[l])
        FFNO_TRY(launch_head(w.f, p->head_w, p->head_b, taps->forecast_list[l], batch, g, p->d.out_features, false, st));
      FFNO_TRY(launch_head(w.f, p->head_w, p->head_b, forecast, batch, g, p->d.out_features, l > 0, st));
    }
    if (taps && taps->spectral && taps->spectral[l] && p->d.spectral_mode != FFNO_MODE_NO_FOURIER)
      FFNO_CUDA_CHECK(cudaMemcpyAsync(taps->spectral[l], w.s, Ubytes, cudaMemcpyDeviceToDevice, st));
    if (taps && taps->x_after && taps->x_after[l])
      FFNO_CUDA_CHECK(cudaMemcpyAsync(taps->x_after[l], nxt, Ubytes, cudaMemcpyDeviceToDevice, st));
    float* t = cur; cur = nxt; nxt = t;
  }
  if (taps && taps->b_last) FFNO_CUDA_CHECK(cudaMemcpyAsync(taps->b_last, w.b, Ubytes, cudaMemcpyDeviceToDevice, st));
  if (!p->d.use_fork && !head_fused)
    FFNO_TRY(launch_head(w.b, p->head_w, p->head_b, forecast, batch, g, p->d.out_features, false, st));
  return FFNO_OK;
}


// Bytes of the caller's workspace the stack forward may touch: the full-batch carve plus one slot per concurrent chunk.
size_t stack_ws_bytes(const ffno_plan* p, int batch) {
  size_t bytes = carve(p, batch, nullptr).bytes;
  const int c = p->chunk_for(batch);
  if (c > 0) bytes += (size_t)(p->n_streams > 1 ? p->n_streams : 1) * carve(p, c, nullptr).bytes;
  return bytes;
}

int block_fwd_impl(ffno_plan* p, const float* x, int batch, float* forecast, const ffno_taps* taps,
                   void* workspace, cudaStream_t st) {
  const int c = taps ? 0 : p->chunk_for(batch);
  if (c == 0) return block_fwd_core(p, x, batch, forecast, taps, workspace, st);
  const int S = p->n_streams > 1 ? p->n_streams : 1;
  char* slots = static_cast<char*>(workspace) + carve(p, batch, nullptr).bytes;
  const size_t slot_bytes = carve(p, c, nullptr).bytes;
  const size_t in_stride = (size_t)p->pts_in * p->d.in_features, out_stride = (size_t)p->pts_in * p->d.out_features;
  if (S > 1) {
    umma_set_sm_limit(p->umma, 0);
    FFNO_CUDA_CHECK(cudaEventRecord(p->ev_fork, st));
    for (int i = 1; i < S; ++i) FFNO_CUDA_CHECK(cudaStreamWaitEvent(p->aux[i - 1], p->ev_fork, 0));
  }
  int status = FFNO_OK, n = 0;
  for (int b0 = 0; b0 < batch && status == FFNO_OK; b0 += c, ++n) {
    const int nb = batch - b0 < c ? batch - b0 : c;
    const int lane = n % S;
    if (S > 1) umma_set_sm_limit(p->umma, p->sm_count / S);
    status = block_fwd_core(p, x + b0 * in_stride, nb, forecast + b0 * out_stride, nullptr, slots + lane * slot_bytes,
                            lane == 0 ? st : p->aux[lane - 1]);
  }
  if (S > 1) {
    umma_set_sm_limit(p->umma, 0);
    for (int i = 1; i < S; ++i) {         // always join, also after an error: the forked streams must rejoin a capture
      FFNO_CUDA_CHECK(cudaEventRecord(p->ev_join[i - 1], p->aux[i - 1]));
      FFNO_CUDA_CHECK(cudaStreamWaitEvent(st, p->ev_join[i - 1], 0));
    }
  }
  return status;
}

constexpr int kMaxGraphFailures = 3;

// Two stage-pipelined forwards must never be in flight at once on one device: each needs ALL of its four kernels
// resident (they wait for each other), and two half-resident sets would wait forever.  Forwards of this process are
// therefore chained through one event per device: a forward starts after the previous one has ended, whatever streams
// the callers use.  (Skipped while the caller's stream is being captured: the events would become part of that graph.)
struct SerialGuard {
  std::mutex mu;
  std::map<int, cudaEvent_t> ev;
  cudaEvent_t get() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    cudaEvent_t& e = ev[dev];
    if (!e && cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) e = nullptr;
    return e;
  }
};
SerialGuard& serial_guard() {
  static SerialGuard g;
  return g;
}
bool stream_is_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs != cudaStreamCaptureStatusNone;
}
int guard_begin(const ffno_plan* p, cudaStream_t st) {
  if (!p->use_umma || stream_is_capturing(st)) return FFNO_OK;
  if (cudaEvent_t e = serial_guard().get()) FFNO_CUDA_CHECK(cudaStreamWaitEvent(st, e, 0));
  return FFNO_OK;
}
int guard_end(const ffno_plan* p, cudaStream_t st) {
  if (!p->use_umma || stream_is_capturing(st)) return FFNO_OK;
  if (cudaEvent_t e = serial_guard().get()) FFNO_CUDA_CHECK(cudaEventRecord(e, st));
  return FFNO_OK;
}

bool is_legacy_stream(cudaStream_t st) { return st == nullptr || st == cudaStreamLegacy; }

// The stream the forward actually runs on (see ffno_plan::own_stream), ordered after everything already enqueued on
// the caller's stream.
int fence_in(ffno_plan* p, cudaStream_t caller, cudaStream_t* run) {
  *run = caller;
  if (!p->graphs || !is_legacy_stream(caller)) return FFNO_OK;
  if (!p->own_stream) {
    FFNO_CUDA_CHECK(cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking));
    FFNO_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
    FFNO_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming));
  }
  FFNO_CUDA_CHECK(cudaEventRecord(p->ev_in, caller));
  FFNO_CUDA_CHECK(cudaStreamWaitEvent(p->own_stream, p->ev_in, 0));
  *run = p->own_stream;
  return FFNO_OK;
}
// ... and the caller's stream ordered after the forward.
int fence_out(ffno_plan* p, cudaStream_t caller, cudaStream_t run) {
  if (run == caller) return FFNO_OK;
  FFNO_CUDA_CHECK(cudaEventRecord(p->ev_out, run));
  FFNO_CUDA_CHECK(cudaStreamWaitEvent(caller, p->ev_out, 0));
  return FFNO_OK;
}

// Capture `body` (which only enqueues kernels on `st`) into an executable graph.  Returns false (and leaves the
// stream usable) if capture is not possible; the caller then launches eagerly.
template <class Body>
bool capture_graph(cudaStream_t st, cudaGraphExec_t* exec, int64_t* launches, Body body) {
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  const long long before = g_launch_counter;
  const int status = body();
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(st, &graph);
  *launches = g_launch_counter - before;
  if (status != FFNO_OK || e != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return false;
  }
  const cudaError_t ei = cudaGraphInstantiate(exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ei != cudaSuccess) {
    cudaGetLastError();
    *exec = nullptr;
    return false;
  }
  return true;
}

// Stack forward on the fixed staging buffers w.io_in -> w.io_out, replayed from a graph when possible.
int block_fwd_staged_on(ffno_plan* p, int batch, void* workspace, cudaStream_t st) {
  const Workspace w = carve(p, batch, workspace);
  ffno_plan::GraphSlot& g = p->g_block;
  if (p->graphs && !stream_is_capturing(st)) {      // (a caller capturing its own graph gets the plain launches)
    if (g.exec && g.batch == batch && g.ws == workspace) {
      FFNO_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
      p->last_launches = g.launches;
      return FFNO_OK;
    }
    if (g.batch == batch && g.ws == workspace && g.seen >= 1 && !g.exec) {
      if (capture_graph(st, &g.exec, &g.launches, [&] { return block_fwd_impl(p, w.io_in, batch, w.io_out, nullptr, workspace, st); })) {
        FFNO_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
        p->last_launches = g.launches;
        return FFNO_OK;
      }
      if (++p->graph_failures >= kMaxGraphFailures) p->graphs = false;   // capture keeps being refused: stay eager
    } else if (g.batch != batch || g.ws != workspace) {
      g.reset();
      g.batch = batch;
      g.ws = workspace;
    }
    g.seen++;
  }
  const long long before = g_launch_counter;
  const int status = block_fwd_impl(p, w.io_in, batch, w.io_out, nullptr, workspace, st);
  p->last_launches = g_launch_counter - before;
  return status;
}

int block_fwd_staged(ffno_plan* p, int batch, void* workspace, cudaStream_t caller) {
  cudaStream_t run;
  FFNO_TRY(fence_in(p, caller, &run));
  FFNO_TRY(guard_begin(p, run));
  const int status = block_fwd_staged_on(p, batch, workspace, run);
  FFNO_TRY(guard_end(p, run));
  FFNO_TRY(fence_out(p, caller, run));
  return status;
}

}  // namespace

// =====================================================================================================
// extern "C"
// =====================================================================================================
extern "C" {

const char* ffno_last_error(void) { return error_buffer(); }
int ffno_abi_version(void) { return FFNO_ABI_VERSION; }

int ffno_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return 0; }
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

static int create_chunk_streams(ffno_plan* p) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess)
    p->sm_count = prop.multiProcessorCount;
  else
    cudaGetLastError();
  if (!p->use_umma || p->n_streams <= 1) { p->n_streams = 1; return FFNO_OK; }
  FFNO_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
  for (int i = 0; i + 1 < p->n_streams; ++i) {
    FFNO_CUDA_CHECK(cudaStreamCreateWithFlags(&p->aux[i], cudaStreamNonBlocking));
    FFNO_CUDA_CHECK(cudaEventCreateWithFlags(&p->ev_join[i], cudaEventDisableTiming));
  }
  return FFNO_OK;
}

int ffno_plan_create(const ffno_desc* desc, ffno_plan** out_plan) {
  FFNO_REQUIRE(out_plan != nullptr, FFNO_ERR_BAD_ARG, "out_plan is NULL");
  *out_plan = nullptr;
  FFNO_TRY(validate_desc(desc));
  ffno_plan* p = new (std::nothrow) ffno_plan();
  FFNO_REQUIRE(p != nullptr, FFNO_ERR_BAD_ARG, "out of host memory");
  p->d = *desc;
  if (desc->transform == FFNO_TRANSFORM_DCT)
    for (int a = 0; a < desc->ndim; ++a) {
      p->dct_modes[a] = desc->modes[a];
      p->d.modes[a] = (desc->modes[a] + 1) / 2;      // coefficient pairs: the "complex modes" of every kernel downstream
    }
  p->pts = 1;
  p->pts_in = 1;
  for (int a = 0; a < desc->ndim; ++a) {
    p->ext[a] = desc->size[a] + desc->pad[a];
    p->pts *= p->ext[a];
    p->pts_in *= desc->size[a];
  }
  p->in_total = desc->in_features + (desc->append_grid ? desc->ndim : 0);
  {
    const char* g = getenv("FFNO_B200_GRAPH");
    p->graphs = !(g && g[0] == '0');
  }
  {
    // Backward modes (ffno_block_bwd).  default: forward recompute on the FP32 kernels (the ReLU masks are the reference's
    // to FP32 round-off), spectral adjoint on the tcgen05 kernels.  "fast": recompute on the tcgen05 kernels too — 1.4x
    // faster, but activations carry the forward's ~1e-5 noise and hidden units that close to the ReLU kink flip their
    // mask, a discrete change of single gradient terms (up to ~1 % of a bias gradient entry on a 8 k-point batch).
    // "fp32": everything on the FP32 kernels.  "fp32-adjoint": diagnostics (tcgen05 recompute, FP32 adjoint).
    const char* b = getenv("FFNO_B200_BWD");
    p->bwd_fp32 = b && strcmp(b, "fp32") == 0;
    p->bwd_fp32_recompute = !(b && (strcmp(b, "fast") == 0 || strcmp(b, "fp32-adjoint") == 0));
    p->bwd_fp32_adjoint = b && strcmp(b, "fp32-adjoint") == 0;
  }
  {
    const char* c = getenv("FFNO_B200_CHUNK");
    const char* ns = getenv("FFNO_B200_STREAMS");
    p->chunk = c ? atoi(c) : 0;
    p->n_streams = ns ? atoi(ns) : 1;
    if (p->n_streams < 1) p->n_streams = 1;
    if (p->n_streams > 4) p->n_streams = 4;
    if (p->chunk < 0) p->chunk = 0;
  }
  p->layers.resize(desc->n_layers);
  int st = build_tables(p);
  if (st == FFNO_OK) {
    const bool ok = umma_supported(&p->d, p->ext) && desc->transform != FFNO_TRANSFORM_RFFT2;
    if (desc->path == FFNO_PATH_UMMA && !ok)
      st = set_error(FFNO_ERR_UNSUPPORTED, "shape does not qualify for the tcgen05 path: %s", umma_why_not(&p->d, p->ext));
    p->use_umma = ok && desc->path != FFNO_PATH_GENERIC;
    if (st == FFNO_OK && p->use_umma) st = umma_create(&p->umma, &p->d, p->ext);
    // rfft2 plans: only the FeedForward (width 64 -> 256 -> 64, the block's other half) qualifies for tcgen05
    p->ff_umma = desc->transform == FFNO_TRANSFORM_RFFT2 && desc->path != FFNO_PATH_GENERIC && umma_supported(&p->d, p->ext) &&
                 ffno_device_ok() == 1;
    if (st == FFNO_OK && p->ff_umma) {
      st = umma_create(&p->umma, &p->d, p->ext);
      if (st == FFNO_OK) umma_set_ff_only(p->umma);
    }
  }
  if (st == FFNO_OK) st = create_chunk_streams(p);
  if (st != FFNO_OK) {
    ffno_plan_destroy(p);
    return st;
  }
  *out_plan = p;
  return FFNO_OK;
}

int ffno_plan_destroy(ffno_plan* plan) {
  if (!plan) return FFNO_OK;
  plan->g_block.reset();
  plan->g_rollout.reset();
  if (plan->own_stream) cudaStreamDestroy(plan->own_stream);
  if (plan->ev_in) cudaEventDestroy(plan->ev_in);
  if (plan->ev_out) cudaEventDestroy(plan->ev_out);
  for (int i = 0; i < 3; ++i) {
    if (plan->aux[i]) cudaStreamDestroy(plan->aux[i]);
    if (plan->ev_join[i]) cudaEventDestroy(plan->ev_join[i]);
  }
  if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
  if (plan->umma) umma_destroy(plan->umma);
  if (plan->umma_adj) umma_destroy(plan->umma_adj);
  for (void* ptr : plan->owned) cudaFree(ptr);
  delete plan;
  return FFNO_OK;
}

int ffno_plan_uses_umma(const ffno_plan* plan) { return plan && (plan->use_umma || plan->ff_umma) ? 1 : 0; }

int ffno_plan_load_params(ffno_plan* p, const ffno_block_params* prm, void* stream) {
  FFNO_REQUIRE(p && prm, FFNO_ERR_BAD_ARG, "plan/params is NULL");
  FFNO_REQUIRE(prm->n_layers == p->d.n_layers && prm->layers, FFNO_ERR_BAD_ARG, "params.n_layers=%d, plan has %d",
               prm->n_layers, p->d.n_layers);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = p->d.width;
  p->g_block.reset();          // captured graphs point at the previous parameter buffers' contents: rebuild
  p->g_rollout.reset();
  p->adj_stale = true;
  p->has_io = prm->in_proj.weight || prm->in_proj.weight_v;
  if (p->has_io) {
    FFNO_TRY(prep_linear(p, prm->in_proj, p->in_total, C, &p->lift, st));
    FFNO_TRY(prep_linear(p, prm->out0, C, p->d.head_hidden, &p->out0, st));
    FFNO_TRY(prep_linear(p, prm->out1, p->d.head_hidden, p->d.out_features, &p->out1, st));
    if (!p->head_w) FFNO_TRY(dev_alloc(p, (size_t)p->d.out_features * C * 4, &p->head_w));
    if (!p->head_b) FFNO_TRY(dev_alloc(p, (size_t)p->d.out_features * 4, &p->head_b));
    FFNO_TRY(launch_fold_head(p->out0.wt, p->out0.bias, p->out1.wt, p->out1.bias, p->head_w, p->head_b, C,
                              p->d.head_hidden, p->d.out_features, st));
  }

  // Spectral weights: prepared once per distinct source pointer (share_weight=True aliases one
  // ParameterList across all layers, grid_2d.py:125-147).  FF weights likewise (share_fork).
  std::map<const void*, float*> seen_mix;
  std::map<const void*, int> seen_ff_back, seen_ff_fork;
  for (int l = 0; l < p->d.n_layers; ++l) {
    const ffno_layer_params& src = prm->layers[l];
    LayerW& dst = p->layers[l];
    if (p->d.spectral_mode == FFNO_MODE_FULL && p->d.transform == FFNO_TRANSFORM_RFFT2) {
      // [C][C][K][K][2] x 2 (low / high row block) -> Wblk[(row j, column ky)][2C][2C], rows j = 0..2K-1
      const int K = p->d.modes[0];
      const float* w0 = src.fourier_weight[0];
      const float* w1 = src.fourier_weight[1];
      FFNO_REQUIRE(w0 && w1, FFNO_ERR_BAD_ARG, "layer %d: fourier_weight[0/1] is NULL", l);
      const size_t half = (size_t)K * K * 4 * C * C;
      const void* key = (const void*)w0;
      auto it = seen_mix.find(key);
      if (it != seen_mix.end()) {
        dst.wmix[0] = it->second;
      } else {
        float*& slot = p->dedup[(const void*)((uintptr_t)(l * 4 + 1))];
        if (!slot) FFNO_TRY(dev_alloc(p, 2 * half * 4, &slot));
        FFNO_TRY(launch_pack_mix_weights(w0, slot, C, K * K, st));
        FFNO_TRY(launch_pack_mix_weights(w1, slot + half, C, K * K, st));
        dst.wmix[0] = slot;
        seen_mix[key] = slot;
      }
    } else if (p->d.spectral_mode == FFNO_MODE_FULL) {
      for (int a = 0; a < p->d.ndim; ++a) {
        const float* wsrc = src.fourier_weight[a];
        FFNO_REQUIRE(wsrc != nullptr, FFNO_ERR_BAD_ARG, "layer %d: fourier_weight[%d] is NULL", l, a);
        const void* key = (const void*)((uintptr_t)wsrc ^ ((uintptr_t)a << 60));
        auto it = seen_mix.find(key);
        if (it != seen_mix.end()) { dst.wmix[a] = it->second; continue; }
        float*& slot = p->dedup[(const void*)((uintptr_t)(l * 4 + a + 1))];   // stable per (layer, axis) slot
        if (!slot) FFNO_TRY(dev_alloc(p, (size_t)p->d.modes[a] * 4 * C * C * 4, &slot));
        if (p->d.transform == FFNO_TRANSFORM_DCT)
          FFNO_TRY(launch_pack_mix_weights_dct(wsrc, slot, C, p->d.modes[a], p->dct_modes[a], st));
        else
          FFNO_TRY(launch_pack_mix_weights(wsrc, slot, C, p->d.modes[a], st));
        dst.wmix[a] = slot;
        seen_mix[key] = slot;
      }
    }
    {
      const void* key = src.backcast_ff.linear[0].weight ? (const void*)src.backcast_ff.linear[0].weight
                                                         : (const void*)src.backcast_ff.linear[0].weight_v;
      auto it = seen_ff_back.find(key);
      if (it != seen_ff_back.end()) dst.back = p->layers[it->second].back;
      else { FFNO_TRY(prep_ff(p, src.backcast_ff, &dst.back, st)); seen_ff_back[key] = l; }
    }
    if (p->d.use_fork) {
      const void* key = src.forecast_ff.linear[0].weight ? (const void*)src.forecast_ff.linear[0].weight
                                                         : (const void*)src.forecast_ff.linear[0].weight_v;
      auto it = seen_ff_fork.find(key);
      if (it != seen_ff_fork.end()) dst.fork = p->layers[it->second].fork;
      else { FFNO_TRY(prep_ff(p, src.forecast_ff, &dst.fork, st)); seen_ff_fork[key] = l; }
    }
  }
  if (p->use_umma || p->ff_umma) {
    std::vector<UmmaLayerSrc> srcs(p->d.n_layers);
    for (int l = 0; l < p->d.n_layers; ++l) {
      for (int a = 0; a < 3; ++a) srcs[l].wmix[a] = p->layers[l].wmix[a];
      srcs[l].w1t = p->layers[l].back.lin[0].wt;
      srcs[l].b1 = p->layers[l].back.lin[0].bias;
      srcs[l].w2t = p->layers[l].back.lin[1].wt;
      srcs[l].b2 = p->layers[l].back.lin[1].bias;
    }
    FFNO_TRY(umma_load_params(p->umma, srcs.data(), p->d_fwd, p->d_inv, st));
  }
  // Forward kernels are launched with programmatic dependent launch and read these prepared parameters in their
  // prologue, possibly before their stream predecessor has finished: make sure nothing of the load is still in flight.
  FFNO_CUDA_CHECK(cudaStreamSynchronize(st));
  p->loaded = true;
  return FFNO_OK;
}

size_t ffno_workspace_bytes(const ffno_plan* plan, int32_t batch) {
  if (!plan || batch < 0) return 0;
  return stack_ws_bytes(plan, batch);
}

int ffno_block_fwd(ffno_plan* p, const float* x, int32_t batch, float* forecast, const ffno_taps* taps,
                   void* workspace, size_t workspace_bytes, void* stream) {
  FFNO_TRY(check_ready(p, batch, workspace, workspace_bytes, ffno_workspace_bytes(p, batch)));
  if (batch == 0) { p->last_launches = 0; return FFNO_OK; }
  FFNO_REQUIRE(x && forecast, FFNO_ERR_BAD_ARG, "x/forecast is NULL");
  FFNO_REQUIRE(p->has_io, FFNO_ERR_STATE, "plan was loaded without lift/head parameters");
  cudaStream_t cst = static_cast<cudaStream_t>(stream);
  if (!taps && p->graphs) {
    const Workspace w = carve(p, batch, workspace);
    const size_t in_b = (size_t)batch * p->pts_in * p->d.in_features * 4;
    const size_t out_b = (size_t)batch * p->pts_in * p->d.out_features * 4;
    FFNO_CUDA_CHECK(cudaMemcpyAsync(w.io_in, x, in_b, cudaMemcpyDeviceToDevice, cst));
    FFNO_TRY(block_fwd_staged(p, batch, workspace, cst));
    FFNO_CUDA_CHECK(cudaMemcpyAsync(forecast, w.io_out, out_b, cudaMemcpyDeviceToDevice, cst));
    return FFNO_OK;
  }
  const long long before = g_launch_counter;
  FFNO_TRY(guard_begin(p, cst));
  int st = block_fwd_impl(p, x, batch, forecast, taps, workspace, cst);
  p->last_launches = g_launch_counter - before;
  FFNO_TRY(guard_end(p, cst));
  return st;
}

size_t ffno_workspace_bytes_host(const ffno_plan* plan, int32_t batch) { return ffno_workspace_bytes(plan, batch); }

int ffno_block_fwd_host(ffno_plan* p, const float* x_host, int32_t batch, float* forecast_host, void* workspace,
                        size_t workspace_bytes, void* stream) {
  FFNO_TRY(check_ready(p, batch, workspace, workspace_bytes, ffno_workspace_bytes(p, batch)));
  if (batch == 0) { p->last_launches = 0; return FFNO_OK; }
  FFNO_REQUIRE(x_host && forecast_host, FFNO_ERR_BAD_ARG, "x_host/forecast_host is NULL");
  FFNO_REQUIRE(p->has_io, FFNO_ERR_STATE, "plan was loaded without lift/head parameters");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const Workspace w = carve(p, batch, workspace);
  const size_t in_b = (size_t)batch * p->pts_in * p->d.in_features * 4;
  const size_t out_b = (size_t)batch * p->pts_in * p->d.out_features * 4;
  FFNO_CUDA_CHECK(cudaMemcpyAsync(w.io_in, x_host, in_b, cudaMemcpyHostToDevice, st));
  FFNO_TRY(block_fwd_staged(p, batch, workspace, st));
  FFNO_CUDA_CHECK(cudaMemcpyAsync(forecast_host, w.io_out, out_b, cudaMemcpyDeviceToHost, st));
  FFNO_CUDA_CHECK(cudaStreamSynchronize(st));
  return FFNO_OK;
}

int ffno_spectral_fwd(ffno_plan* p, int32_t layer, const float* x, int32_t batch, float* s, void* workspace,
                      size_t workspace_bytes, void* stream) {
  FFNO_TRY(check_ready(p, batch, workspace, workspace_bytes, ffno_workspace_bytes(p, batch)));
  FFNO_REQUIRE(layer >= 0 && layer < p->d.n_layers, FFNO_ERR_BAD_ARG, "layer=%d", layer);
  FFNO_REQUIRE(x && s, FFNO_ERR_BAD_ARG, "x/s is NULL");
  FFNO_REQUIRE(p->d.spectral_mode != FFNO_MODE_NO_FOURIER, FFNO_ERR_STATE, "plan was built with mode=no-fourier");
  const Workspace w = carve(p, batch, workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (batch == 0) return FFNO_OK;
  if (p->use_umma) return umma_spectral_fwd(p->umma, layer, x, batch, s, w.F, w.R, w.umma, st);
  return spectral_generic(p, p->layers[layer], x, batch, s, w.F, w.R, st);
}

int ffno_spectral_split_fwd(ffno_plan* p, int32_t layer, const float* x, int32_t batch, float* const* s_axis,
                            void* workspace, size_t workspace_bytes, void* stream) {
  FFNO_TRY(check_ready(p, batch, workspace, workspace_bytes, ffno_workspace_bytes(p, batch)));
  FFNO_REQUIRE(layer >= 0 && layer < p->d.n_layers, FFNO_ERR_BAD_ARG, "layer=%d", layer);
  FFNO_REQUIRE(x && s_axis && s_axis[0], FFNO_ERR_BAD_ARG, "x/s_axis is NULL");
  FFNO_REQUIRE(p->d.spectral_mode != FFNO_MODE_NO_FOURIER, FFNO_ERR_STATE, "plan was built with mode=no-fourier");
  const Workspace w = carve(p, batch, workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (batch == 0) return FFNO_OK;
  const long long before = g_launch_counter;
  int status;
  if (p->use_umma) {
    for (int a = 1; a < p->d.ndim; ++a) FFNO_REQUIRE(s_axis[a], FFNO_ERR_BAD_ARG, "s_axis[%d] is NULL", a);
    float* bufs[3] = {s_axis[0], p->d.ndim > 1 ? s_axis[1] : nullptr, p->d.ndim > 2 ? s_axis[2] : nullptr};
    int n_written = 0;
    status = umma_spectral_split_fwd(p->umma, layer, x, batch, bufs, w.F, w.R, w.umma, &n_written, st);
    for (int a = n_written; status == FFNO_OK && a < p->d.ndim; ++a)       // summed fallback: the others contribute zero
      FFNO_CUDA_CHECK(cudaMemsetAsync(bufs[a], 0, (size_t)batch * p->pts * p->d.width * 4, st));
  } else {
    status = spectral_generic(p, p->layers[layer], x, batch, s_axis[0], w.F, w.R, st);
    for (int a = 1; status == FFNO_OK && a < p->d.ndim; ++a)
      if (s_axis[a]) FFNO_CUDA_CHECK(cudaMemsetAsync(s_axis[a], 0, (size_t)batch * p->pts * p->d.width * 4, st));
  }
  p->last_launches = g_launch_counter - before;
  return status;
}

int ffno_ff_fwd(ffno_plan* p, int32_t layer, int32_t which, const float* s, const float* residual, int32_t batch,
                float* y, void* workspace, size_t workspace_bytes, void* stream) {
  FFNO_TRY(check_ready(p, batch, workspace, workspace_bytes, ffno_workspace_bytes(p, batch)));
  FFNO_REQUIRE(layer >= 0 && layer < p->d.n_layers, FFNO_ERR_BAD_ARG, "layer=%d", layer);
  FFNO_REQUIRE(which == 0 || (which == 1 && p->d.use_fork), FFNO_ERR_BAD_ARG, "which=%d", which);
  FFNO_REQUIRE(s && y, FFNO_ERR_BAD_ARG, "s/y is NULL");
  const Workspace w = carve(p, batch, workspace);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long P = (long long)batch * p->pts;
  if (batch == 0) return FFNO_OK;
  const FFW& ff = which == 0 ? p->layers[layer].back : p->layers[layer].fork;
  if ((p->use_umma || p->ff_umma) && which == 0 && p->d.n_ff_layers == 2 && !p->d.layer_norm)
    return umma_ff_fwd(p->umma, layer, s, residual, batch, y, w.umma, st);
  if (residual) return ff_generic(p, ff, s, residual, P, y, nullptr, w, st);
  return ff_generic(p, ff, s, nullptr, P, nullptr, y, w, st);
}

int ffno_linear_fwd(const ffno_linear_params* lin, const float* x, int64_t rows, float* y, int32_t relu,
                    void* scratch, size_t scratch_bytes, void* stream) {
  FFNO_REQUIRE(lin && x && y, FFNO_ERR_BAD_ARG, "NULL argument");
  const int in = lin->in_features, out = lin->out_features;
  FFNO_REQUIRE(in >= 1 && out >= 1 && rows >= 0, FFNO_ERR_BAD_ARG, "linear [%d,%d] rows=%lld", out, in, (long long)rows);
  FFNO_REQUIRE(scratch && scratch_bytes >= (size_t)in * out * 4, FFNO_ERR_WORKSPACE, "scratch < %zu B", (size_t)in * out * 4);
  const float* v = lin->weight ? lin->weight : lin->weight_v;
  FFNO_REQUIRE(v && (lin->weight || lin->weight_g), FFNO_ERR_BAD_ARG, "linear has no usable weight");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* wt = static_cast<float*>(scratch);
  FFNO_TRY(launch_weight_fold_transpose(v, lin->weight ? nullptr : lin->weight_g, wt, out, in, st));
  return launch_linear_any(x, wt, lin->bias, y, rows, in, out, relu != 0, st);
}

int ffno_layernorm_fwd(const float* x, const float* weight, const float* bias, int64_t rows, int32_t C, float* y,
                       void* stream) {
  FFNO_REQUIRE(x && weight && bias && y && C >= 1, FFNO_ERR_BAD_ARG, "NULL argument");
  return launch_layernorm_residual(x, weight, bias, nullptr, y, nullptr, rows, C, static_cast<cudaStream_t>(stream));
}

int ffno_rel_l2(const float* x, int64_t x_stride_b, int64_t x_stride_i, const float* y, int64_t y_stride_b,
                int64_t y_stride_i, int32_t batch, int64_t n, float* out, void* stream) {
  FFNO_REQUIRE(x && y && out && batch >= 0 && n >= 1, FFNO_ERR_BAD_ARG, "bad argument");
  return launch_rel_l2(x, x_stride_b, x_stride_i, y, y_stride_b, y_stride_i, batch, n, out,
                       static_cast<cudaStream_t>(stream));
}

static bool rollout_layout(const ffno_plan* p, const ffno_rollout_extras* ex, bool* use_velocity, bool* has_force,
                           bool* has_mu) {
  // feature count <-> options: [w, (q, v,) gx, gy, (f,) (mu)]; without extras 5 features mean the velocity set
  *use_velocity = ex ? ex->use_velocity != 0 : p->d.in_features == 5;
  *has_force = ex && ex->force_steps > 0;
  *has_mu = ex && ex->mu != nullptr;
  return p->d.in_features == 3 + (*use_velocity ? 2 : 0) + (*has_force ? 1 : 0) + (*has_mu ? 1 : 0);
}

size_t ffno_rollout_workspace_bytes_ex(const ffno_plan* plan, int32_t batch, int32_t n_steps, int32_t use_velocity,
                                       int32_t force_steps, int32_t has_mu) {
  if (!plan || batch < 0 || n_steps < 0 || force_steps < 0) return 0;
  size_t frame = ((size_t)batch * plan->pts_in * 4 + 255) / 256 * 256;
  size_t preds = ((size_t)batch * plan->pts_in * n_steps * 4 + 255) / 256 * 256;
  size_t vel = 0;
  if (use_velocity && plan->d.ndim == 2)      // q, v and the scratch of the four DFT passes
    vel = 2 * frame + (velocity_scratch_floats(batch, plan->d.size[0], plan->d.size[1]) * 4 + 255) / 256 * 256;
  size_t force = ((size_t)batch * plan->pts_in * force_steps * 4 + 255) / 256 * 256;
  size_t mu = has_mu ? ((size_t)batch * 4 + 255) / 256 * 256 : 0;
  return ffno_workspace_bytes(plan, batch) + frame + preds + vel + force + mu;
}

size_t ffno_rollout_workspace_bytes(const ffno_plan* plan, int32_t batch, int32_t n_steps) {
  return ffno_rollout_workspace_bytes_ex(plan, batch, n_steps, plan && plan->d.in_features == 5, 0, 0);
}

int ffno_rollout_fwd_ex(ffno_plan* p, const float* frame0, int32_t batch, int32_t n_steps, const float* mean_host,
                        const float* std_host, float low, float high, const ffno_rollout_extras* ex, float* preds,
                        void* workspace, size_t workspace_bytes, void* stream) {
  FFNO_REQUIRE(p != nullptr, FFNO_ERR_BAD_ARG, "plan is NULL");
  bool use_velocity, has_force, has_mu;
  const bool layout_ok = rollout_layout(p, ex, &use_velocity, &has_force, &has_mu);
  const int force_steps = has_force ? ex->force_steps : 0;
  FFNO_TRY(check_ready(p, batch, workspace, workspace_bytes,
                       ffno_rollout_workspace_bytes_ex(p, batch, n_steps, use_velocity, force_steps, has_mu)));
  if (batch == 0) { p->last_launches = 0; return FFNO_OK; }
  FFNO_REQUIRE(frame0 && preds && mean_host && std_host, FFNO_ERR_BAD_ARG, "NULL argument");
  FFNO_REQUIRE(p->d.ndim == 2 && p->d.out_features == 1 && !p->d.append_grid && p->d.pad[0] == 0 && p->d.pad[1] == 0 &&
                   !p->d.use_fork,
               FFNO_ERR_UNSUPPORTED, "rollout needs the Grid2DMarkovExperiment layout: 2-D periodic grid, out=1, no fork");
  FFNO_REQUIRE(layout_ok, FFNO_ERR_BAD_ARG,
               "rollout features [w%s, gx, gy%s%s] do not match the plan's in_features=%d", use_velocity ? ", q, v" : "",
               has_force ? ", f" : "", has_mu ? ", mu" : "", p->d.in_features);
  FFNO_REQUIRE(n_steps >= 1, FFNO_ERR_BAD_ARG, "n_steps=%d", n_steps);
  FFNO_REQUIRE(!has_force || (ex->force && (force_steps == 1 || force_steps == n_steps)), FFNO_ERR_BAD_ARG,
               "force_steps=%d must be 1 (static forcing) or n_steps=%d, with a non-NULL force", force_steps, n_steps);
  FFNO_REQUIRE(p->has_io, FFNO_ERR_STATE, "plan was loaded without lift/head parameters");
  cudaStream_t caller = static_cast<cudaStream_t>(stream), st;
  FFNO_TRY(fence_in(p, caller, &st));
  const int X = p->d.size[0], Y = p->d.size[1];
  const Workspace w = carve(p, batch, workspace);
  char* base = static_cast<char*>(workspace);
  const size_t frame_b = (size_t)batch * X * Y * 4;
  const size_t stack_b = stack_ws_bytes(p, batch);
  const size_t frame_al = (frame_b + 255) / 256 * 256;
  const size_t preds_al = (frame_b * n_steps + 255) / 256 * 256;
  char* cur = base + stack_b;
  auto take = [&](size_t bytes) { float* r = reinterpret_cast<float*>(cur); cur += bytes; return r; };
  float* frame_st = take(frame_al);
  float* preds_st = take(preds_al);
  float* vel_q = use_velocity ? take(frame_al) : nullptr;
  float* vel_v = use_velocity ? take(frame_al) : nullptr;
  float* vel_scratch = use_velocity ? take((velocity_scratch_floats(batch, X, Y) * 4 + 255) / 256 * 256) : nullptr;
  float* force_st = has_force ? take((frame_b * force_steps + 255) / 256 * 256) : nullptr;
  float* mu_st = has_mu ? take(((size_t)batch * 4 + 255) / 256 * 256) : nullptr;
  MeanStd ms{};
  for (int i = 0; i < p->d.in_features; ++i) { ms.m[i] = mean_host[i]; ms.s[i] = std_host[i]; }
  const bool extras = has_force || has_mu;

  // the whole step loop: features -> layer stack -> de-normalise, each forecast feeding the next step
  // (routines/grid_2d_markov.py:263-321); every pointer it touches lives in the caller's workspace
  auto body = [&]() -> int {
    for (int t = 0; t < n_steps; ++t) {
      // the frame fed to this step: ground truth at t = 0, then the previous (de-normalised) forecast
      const float* frame = t == 0 ? frame_st : preds_st + (t - 1);
      const long long fs_b = t == 0 ? (long long)X * Y : (long long)X * Y * n_steps;
      const int fs_xy = t == 0 ? 1 : n_steps;
      if (use_velocity)       // recomputed from every fed-back forecast (grid_2d_markov.py:268-285)
        FFNO_TRY(launch_velocity(frame, fs_b, fs_xy, batch, X, Y, p->domain[0], p->domain[1], vel_q, vel_v, vel_scratch, st));
      if (extras)
        FFNO_TRY(launch_rollout_features_ex(frame, fs_b, fs_xy, vel_q, vel_v, force_st, force_steps, t, mu_st, w.io_in, batch,
                                            X, Y, low, high, ms, st));
      else
        FFNO_TRY(launch_rollout_features(frame, fs_b, fs_xy, vel_q, vel_v, w.io_in, batch, X, Y, low, high, ms, st));
      FFNO_TRY(block_fwd_impl(p, w.io_in, batch, w.io_out, nullptr, workspace, st));
      FFNO_TRY(launch_rollout_denorm(w.io_out, preds_st, batch, X * Y, n_steps, t, ms, st));
    }
    return FFNO_OK;
  };

  // inputs into fixed workspace addresses, so that a captured graph stays valid whatever the caller passes next
  FFNO_CUDA_CHECK(cudaMemcpyAsync(frame_st, frame0, frame_b, cudaMemcpyDeviceToDevice, st));
  if (has_force) FFNO_CUDA_CHECK(cudaMemcpyAsync(force_st, ex->force, frame_b * force_steps, cudaMemcpyDeviceToDevice, st));
  if (has_mu) FFNO_CUDA_CHECK(cudaMemcpyAsync(mu_st, ex->mu, (size_t)batch * 4, cudaMemcpyDeviceToDevice, st));
  FFNO_TRY(guard_begin(p, st));
  ffno_plan::GraphSlot& g = p->g_rollout;
  const int variant = (use_velocity ? 1 : 0) | (force_steps << 2) | (has_mu ? 2 : 0);
  bool done = false;
  if (p->graphs && !stream_is_capturing(st)) {
    const bool same = g.batch == batch && g.ws == workspace && g.n_steps == n_steps && g.low == low && g.high == high &&
                      g.variant == variant && memcmp(&g.ms, &ms, sizeof(ms)) == 0;
    if (same && g.exec) {
      FFNO_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
      p->last_launches = g.launches;
      done = true;
    } else if (same && g.seen >= 1) {
      if (capture_graph(st, &g.exec, &g.launches, body)) {
        FFNO_CUDA_CHECK(cudaGraphLaunch(g.exec, st));
        p->last_launches = g.launches;
        done = true;
      } else if (++p->graph_failures >= kMaxGraphFailures) {
        p->graphs = false;
      }
    } else if (!same) {
      g.reset();
      g.batch = batch; g.ws = workspace; g.n_steps = n_steps; g.low = low; g.high = high; g.ms = ms; g.variant = variant;
    }
    if (!done) g.seen++;
  }
  if (!done) {
    const long long before = g_launch_counter;
    FFNO_TRY(body());
    p->last_launches = g_launch_counter - before;
  }
  FFNO_TRY(guard_end(p, st));
  FFNO_CUDA_CHECK(cudaMemcpyAsync(preds, preds_st, frame_b * n_steps, cudaMemcpyDeviceToDevice, st));
  return fence_out(p, caller, st);
}

int ffno_rollout_fwd(ffno_plan* p, const float* frame0, int32_t batch, int32_t n_steps, const float* mean_host,
                     const float* std_host, float low, float high, float* preds, void* workspace,
                     size_t workspace_bytes, void* stream) {
  FFNO_REQUIRE(p != nullptr, FFNO_ERR_BAD_ARG, "plan is NULL");
  FFNO_REQUIRE(p->d.in_features == 3 || p->d.in_features == 5, FFNO_ERR_UNSUPPORTED,
               "rollout needs the torus_li/markov (in=3) or torus_kochkov (in=5, velocity features) layout; "
               "force / mu channels go through ffno_rollout_fwd_ex");
  return ffno_rollout_fwd_ex(p, frame0, batch, n_steps, mean_host, std_host, low, high, nullptr, preds, workspace,
                             workspace_bytes, stream);
}

// ---- backward pass (include/ffno_b200.h: ffno_block_bwd) ---------------------------------------------------------
namespace {

struct BwdWs {
  float *xs, *ss, *b, *gx, *gb, *ds, *h, *dh, *F, *dR, *dF, *R, *h0, *dh0, *wT, *dwf, *fc, *fwd_ws, *b_dense, *g_dense, *rows_in, *d_rows;
  size_t bytes;
};

size_t bwd_scratch_floats(const ffno_plan* p) {
  const size_t C = p->d.width, H = p->hidden(), Hh = p->d.head_hidden;
  size_t m = C * H;
  if (C * Hh > m) m = C * Hh;
  if (Hh * (size_t)p->d.out_features > m) m = Hh * (size_t)p->d.out_features;
  if ((size_t)p->in_total * C > m) m = (size_t)p->in_total * C;
  for (int a = 0; a < p->d.ndim; ++a)
    if ((size_t)p->d.modes[a] * 4 * C * C > m) m = (size_t)p->d.modes[a] * 4 * C * C;
  if (p->d.transform == FFNO_TRANSFORM_RFFT2) {      // 2 K^2 mode blocks
    const size_t n = (size_t)2 * p->d.modes[0] * p->d.modes[1] * 4 * C * C;
    if (n > m) m = n;
  }
  return m;
}

BwdWs carve_bwd(const ffno_plan* p, int batch, void* base) {
  BwdWs w{};
  Carver c(base);
  const long long P = (long long)batch * p->pts;
  const size_t U = (size_t)P * p->d.width, PH = (size_t)P * p->hidden();
  w.xs = c.take(U * p->d.n_layers);      // x_0 .. x_{L-1}: the input of every layer
  w.ss = c.take(U * p->d.n_layers);      // the spectral output of every layer
  w.b = c.take(U);
  w.gx = c.take(U);
  w.gb = c.take(U);
  w.ds = c.take(U);
  w.h = c.take(PH);
  w.dh = c.take(PH);
  size_t spec = 0;
  for (int a = 0; a < p->d.ndim; ++a) spec += U / p->ext[a] * 2 * p->d.modes[a];     // every axis back to back
  if (p->d.transform == FFNO_TRANSFORM_RFFT2) spec = (size_t)batch * 2 * p->ext[0] * 2 * p->d.modes[1] * p->d.width;
  w.F = c.take(spec);
  w.R = c.take(spec);
  w.dR = c.take(spec);
  w.dF = c.take(spec);
  w.h0 = c.take((size_t)P * p->d.head_hidden);
  w.dh0 = c.take((size_t)P * p->d.head_hidden);
  w.wT = c.take(bwd_scratch_floats(p));
  w.dwf = c.take(bwd_scratch_floats(p));
  w.fc = c.take((size_t)batch * p->pts_in * p->d.out_features);
  const bool mesh = p->pts != p->pts_in || p->d.append_grid;       // padded / grid-appended stacks: dense copies
  const size_t Pin = (size_t)batch * p->pts_in;
  w.b_dense = c.take(mesh ? Pin * p->d.width : 0);
  w.g_dense = c.take(mesh ? Pin * p->d.width : 0);
  w.rows_in = c.take(mesh ? Pin * p->in_total : 0);
  w.d_rows = c.take(mesh ? Pin * p->in_total : 0);
  w.fwd_ws = c.take(p->use_umma ? (stack_ws_bytes(p, batch) + 3) / 4 : 0);      // the tcgen05 forward's own workspace
  w.bytes = c.off;
  return w;
}

// gradient of one (weight-normed or plain) linear from the gradient dwf[out][in] of its folded weight
int linear_param_grads(const ffno_linear_params& prm, const ffno_linear_grads& g, const float* dwf, int out, int in,
                       cudaStream_t st) {
  if (prm.weight) {
    if (g.weight) FFNO_TRY(launch_axpy(g.weight, dwf, (long long)out * in, st));
  } else if (g.weight_g && g.weight_v) {
    FFNO_TRY(launch_wnorm_bwd(dwf, prm.weight_v, prm.weight_g, g.weight_g, g.weight_v, out, in, st));
  } else {
    FFNO_REQUIRE(!g.weight_g && !g.weight_v, FFNO_ERR_BAD_ARG, "weight_g and weight_v gradients come together");
  }
  return FFNO_OK;
}

// One linear y[P][out] = x[P][in] W^T + b, W given folded + transposed as wt[in][out]: accumulates the parameter
// gradients and (dx != NULL) writes dx[P][in] = dy W.  `wide` selects the 64-wide GEMM (dims multiples of 4).
int linear_bwd(const ffno_plan* p, const Lin& lin, const ffno_linear_params& prm, const ffno_linear_grads& g,
               const float* x, const float* dy, float* dx, long long P, const BwdWs& w, cudaStream_t st,
               const float* relu_mask = nullptr) {
  const int in = lin.in, out = lin.out;
  const bool want_w = g.weight || g.weight_v;
  if (want_w) {
    FFNO_CUDA_CHECK(cudaMemsetAsync(w.dwf, 0, (size_t)in * out * 4, st));
    FFNO_TRY(launch_linear_wgrad(dy, x, w.dwf, P, out, in, p->sm_count, st));
    FFNO_TRY(linear_param_grads(prm, g, w.dwf, out, in, st));
  }
  if (g.bias && prm.bias) FFNO_TRY(launch_colsum(dy, g.bias, P, out, st));
  if (dx) {
    FFNO_TRY(launch_transpose(lin.wt, w.wT, in, out, 1, st));      // wT[out][in]: the K x N operand of dx = dy W
    if (in % 4 == 0 && out % 4 == 0) {
      FFNO_TRY(launch_linear(dy, w.wT, nullptr, nullptr, dx, nullptr, P, out, in, false, st, relu_mask));
    } else {
      FFNO_REQUIRE(!relu_mask, FFNO_ERR_UNSUPPORTED, "masked linear backward needs dimensions that are multiples of 4");
      FFNO_TRY(launch_linear_any(dy, w.wT, nullptr, dx, P, out, in, false, st));
    }
  }
  return FFNO_OK;
}

}  // namespace

namespace {
// (Re)build the adjoint kernel state: transposed mode blocks + the transposed tables in swapped roles.
int ensure_adjoint(ffno_plan* p, cudaStream_t st) {
  if (!p->use_umma || p->bwd_fp32 || !p->adj_stale) return FFNO_OK;
  const int C = p->d.width;
  if (!p->umma_adj) FFNO_TRY(umma_create(&p->umma_adj, &p->d, p->ext));
  std::map<const float*, bool> done;
  std::vector<UmmaLayerSrc> srcs(p->d.n_layers);
  for (int l = 0; l < p->d.n_layers; ++l) {
    for (int a = 0; a < 3; ++a) srcs[l].wmix[a] = nullptr;
    for (int a = 0; a < p->d.ndim; ++a) {
      const float* w = p->layers[l].wmix[a];
      float*& wt = p->wmixT[w];
      if (!wt) FFNO_TRY(dev_alloc(p, (size_t)p->d.modes[a] * 4 * C * C * 4, &wt));
      if (!done[w]) {
        FFNO_TRY(launch_transpose(w, wt, 2 * C, 2 * C, p->d.modes[a], st));
        done[w] = true;
      }
      srcs[l].wmix[a] = wt;
    }
    // the FF images of the adjoint state are never used (the FF backward needs the ReLU mask): any valid weights do
    srcs[l].w1t = p->layers[l].back.lin[0].wt;
    srcs[l].b1 = p->layers[l].back.lin[0].bias;
    srcs[l].w2t = p->layers[l].back.lin[1].wt;
    srcs[l].b2 = p->layers[l].back.lin[1].bias;
  }
  FFNO_TRY(umma_load_params(p->umma_adj, srcs.data(), p->d_invT, p->d_fwdT, st));
  p->adj_stale = false;
  return FFNO_OK;
}
}  // namespace

size_t ffno_block_bwd_workspace_bytes(const ffno_plan* plan, int32_t batch) {
  if (!plan || batch < 0) return 0;
  return carve_bwd(plan, batch, nullptr).bytes;
}

// Layer loop alone (ffno_layers_bwd): no lift / head; x is x_0, d_forecast is dL/dx_L, every layer adds `bias`.
struct LayersOnly {
  const float* bias;      // [pts, C] added after every layer's residual sum, broadcast over the batch (NULL: none)
  float* d_bias;          // += its gradient (NULL: skipped)
};

static int block_bwd_core(ffno_plan* p, const ffno_block_params* prm, const float* x, const float* d_forecast, int32_t batch,
                          const ffno_block_grads* grads, float* dx, void* workspace, size_t workspace_bytes, void* stream,
                          const LayersOnly* lo) {
  FFNO_TRY(check_ready(p, batch, workspace, workspace_bytes, ffno_block_bwd_workspace_bytes(p, batch)));
  FFNO_REQUIRE(prm && grads && x && d_forecast, FFNO_ERR_BAD_ARG, "NULL argument");
  FFNO_REQUIRE(prm->n_layers == p->d.n_layers && grads->n_layers == p->d.n_layers && prm->layers && grads->layers,
               FFNO_ERR_BAD_ARG, "params / grads must describe the plan's %d layers", p->d.n_layers);
  FFNO_REQUIRE(lo || p->has_io, FFNO_ERR_STATE, "plan was loaded without lift/head parameters");
  FFNO_REQUIRE(!lo || (p->pts == p->pts_in && !p->d.append_grid), FFNO_ERR_UNSUPPORTED, "layer-loop backward: no padding / grid append");
  FFNO_REQUIRE(p->d.n_ff_layers == 2 && !p->d.layer_norm && !p->d.use_fork && p->d.spectral_mode == FFNO_MODE_FULL,
               FFNO_ERR_UNSUPPORTED, "backward is implemented for n_ff_layers = 2, no LayerNorm, no fork, mode 'full'");
  const bool mesh = p->pts != p->pts_in || p->d.append_grid;       // zero-padded / grid-appended (mesh_3d.py:161-166)
  if (batch == 0) return FFNO_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long before = g_launch_counter;
  const BwdWs w = carve_bwd(p, batch, workspace);
  const LiftGeom g = p->geom();
  const int C = p->d.width, H = p->hidden(), Hh = p->d.head_hidden, O = p->d.out_features, nl = p->d.n_layers;
  const long long P = (long long)batch * p->pts;
  const size_t U = (size_t)P * C;
  Workspace wf{};            // ff_generic only needs the hidden buffer
  wf.h0 = w.h;

  // ---- 1. forward again, keeping the input x_l and the spectral output s_l of every layer (and the last backcast)
  // tc: the tcgen05 kernels take part (see the FFNO_B200_BWD modes at plan creation)
  const bool tc = p->use_umma && !p->bwd_fp32 && !lo;      // (the layer-loop form runs on the FP32 kernels)
  const bool tc_fwd = tc && !p->bwd_fp32_recompute, tc_adj = tc && !p->bwd_fp32_adjoint;
  if (tc) FFNO_TRY(ensure_adjoint(p, st));
  if (tc_fwd) {
    // the tcgen05 forward with taps: the same kernels as inference
    FFNO_TRY(ensure_adjoint(p, st));
    std::vector<float*> xa(nl), sa(nl);
    for (int l = 0; l < nl; ++l) {
      xa[l] = l + 1 < nl ? w.xs + (size_t)(l + 1) * U : nullptr;      // x after layer l = input of layer l + 1
      sa[l] = w.ss + (size_t)l * U;
    }
    ffno_taps taps{};
    taps.lift = w.xs;
    taps.x_after = xa.data();
    taps.spectral = sa.data();
    taps.b_last = w.b;
    FFNO_TRY(block_fwd_core(p, x, batch, w.fc, &taps, w.fwd_ws, st));
  } else {
    if (lo) FFNO_CUDA_CHECK(cudaMemcpyAsync(w.xs, x, U * 4, cudaMemcpyDeviceToDevice, st));
    else FFNO_TRY(launch_lift(x, p->lift.wt, p->lift.bias, w.xs, batch, g, st));
    for (int l = 0; l < nl; ++l) {
      const LayerW& lw = p->layers[l];
      float* xl = w.xs + (size_t)l * U;
      float* sl = w.ss + (size_t)l * U;
      FFNO_TRY(spectral_generic(p, lw, xl, batch, sl, w.F, w.R, st));
      // x_{l+1} = x_l + b_l; the last layer's residual sum is dead (the head reads b, grid_2d.py:170-172)
      FFNO_TRY(ff_generic(p, lw.back, sl, xl, P, l + 1 < nl ? xl + U : nullptr, l + 1 < nl ? nullptr : w.b, wf, st));
      if (lo && lo->bias && l + 1 < nl)      // x_{l+1} += bias, every sample (point_cloud_2d.py:205-206)
        for (int bi = 0; bi < batch; ++bi) FFNO_TRY(launch_axpy(xl + U + (size_t)bi * (U / batch), lo->bias, (long long)(U / batch), st));
    }
  }
  // ---- 2. head: forecast = out1(out0(b)) on the unpadded region (mesh_3d.py:173-174)
  const long long Pin = (long long)batch * p->pts_in;
  const float* b_rows = w.b;
  if (lo) {
    // no head: the incoming gradient is dL/dx_L, which reaches the last layer's backcast AND its residual input
    FFNO_CUDA_CHECK(cudaMemcpyAsync(w.gb, d_forecast, U * 4, cudaMemcpyDeviceToDevice, st));
    FFNO_CUDA_CHECK(cudaMemcpyAsync(w.gx, d_forecast, U * 4, cudaMemcpyDeviceToDevice, st));
  } else {
  if (mesh) {
    FFNO_TRY(launch_crop_pad(w.b_dense, w.b, batch, g, false, st));
    b_rows = w.b_dense;
  }
  FFNO_TRY(launch_linear(b_rows, p->out0.wt, p->out0.bias, nullptr, w.h0, nullptr, Pin, C, Hh, false, st));
  FFNO_TRY(linear_bwd(p, p->out1, prm->out1, grads->out1, w.h0, d_forecast, w.dh0, Pin, w, st));
  FFNO_TRY(linear_bwd(p, p->out0, prm->out0, grads->out0, b_rows, w.dh0, mesh ? w.g_dense : w.gb, Pin, w, st));
  if (mesh) FFNO_TRY(launch_crop_pad(w.g_dense, w.gb, batch, g, true, st));      // zero gradient in the padding
  }

  // ---- 3. layers, last to first.  gx = dL/dx_{l+1}; the layer's backcast gradient is gx (x_{l+1} = x_l + b_l), or
  //         the head's for the last layer
  if (!lo) FFNO_CUDA_CHECK(cudaMemsetAsync(w.gx, 0, U * 4, st));
  for (int l = nl - 1; l >= 0; --l) {
    if (lo && lo->d_bias)      // bias enters x_{l+1} of every layer: d_bias += sum over the samples of dL/dx_{l+1}
      for (int bi = 0; bi < batch; ++bi) FFNO_TRY(launch_axpy(lo->d_bias, w.gx + (size_t)bi * (U / batch), (long long)(U / batch), st));
    const LayerW& lw = p->layers[l];
    const ffno_layer_params& lp = prm->layers[l];
    const ffno_layer_grads& lg = grads->layers[l];
    const float* xl = w.xs + (size_t)l * U;
    const float* sl = w.ss + (size_t)l * U;
    const float* gb = l == nl - 1 ? w.gb : w.gx;
    // FeedForward (feedforward.py:6-24): h = relu(W1 s + b1), b = W2 h + b2
    FFNO_TRY(launch_linear(sl, lw.back.lin[0].wt, lw.back.lin[0].bias, nullptr, w.h, nullptr, P, C, H, true, st));
    // (the ReLU's backward is the mask h > 0 applied in the epilogue of the g_h GEMM)
    FFNO_TRY(linear_bwd(p, lw.back.lin[1], lp.backcast_ff.linear[1], lg.backcast_ff[1], w.h, gb, w.dh, P, w, st, w.h));
    FFNO_TRY(linear_bwd(p, lw.back.lin[0], lp.backcast_ff.linear[0], lg.backcast_ff[0], sl, w.dh, w.ds, P, w, st));
    // spectral operator (grid_2d.py:51-99): s = sum_a Inv_a Mix_a Fwd_a x  =>  gx += sum_a Fwd_a^T Mix_a^T Inv_a^T ds
    if (p->d.transform == FFNO_TRANSFORM_RFFT2) {
      // un-factorized variant (zongyi_fno/grid_plus_2d.py:52-83): s = InvN C- InvRow Mix C+ FwdRow FwdN x (spectral_rfft2);
      // the adjoint walks the same seven linear maps backwards with transposed tables, combine^T = expand
      const int M = p->ext[0], N = p->ext[1], K = p->d.modes[1], Rr = 2 * K;
      const long long row = (long long)2 * K * C;
      FFNO_TRY(launch_axis_transform(w.ds, p->d_invT[1], w.dR, (long long)batch * M, N, 2 * K, C, false, st));
      FFNO_TRY(launch_c2c_expand(w.dR, w.dF, batch, M, K, C, -1.f, st));
      FFNO_TRY(launch_axis_transform(w.dF, p->d_invT[0], w.dR, batch, 2 * M, Rr, row, false, st));      // dR: d loss / d R
      if (lg.fourier_weight[0] || lg.fourier_weight[1]) {
        FFNO_REQUIRE(lg.fourier_weight[0] && lg.fourier_weight[1], FFNO_ERR_BAD_ARG,
                     "rfft2: both fourier_weight gradients or none");
        FFNO_TRY(launch_axis_transform(xl, p->d_fwd[1], w.F, (long long)batch * M, N, 2 * K, C, false, st));
        FFNO_TRY(launch_axis_transform(w.F, p->d_fwd[0], w.R, batch, M, 2 * Rr, row, false, st));
        FFNO_TRY(launch_c2c_combine(w.R, w.F, batch, Rr, K, C, 1.f, st));                               // F: the mix input
        FFNO_TRY(launch_mix_wgrad(w.F, w.dR, lg.fourier_weight[0], batch, Rr * K, 1, C, p->sm_count, st, 0,
                                  lg.fourier_weight[1], K * K));
      }
      FFNO_TRY(launch_transpose(lw.wmix[0], w.wT, 2 * C, 2 * C, Rr * K, st));
      FFNO_TRY(launch_mode_mix(w.dR, w.wT, w.dF, batch, Rr * K, 1, C, st));
      FFNO_TRY(launch_c2c_expand(w.dF, w.dR, batch, Rr, K, C, 1.f, st));
      FFNO_TRY(launch_axis_transform(w.dR, p->d_fwdT[0], w.dF, batch, 2 * Rr, M, row, false, st));
      FFNO_TRY(launch_axis_transform(w.dF, p->d_fwdT[1], w.gx, (long long)batch * M, 2 * K, N, C, true, st));
      continue;
    }
    if (tc_adj) {
      // tcgen05: the forward's three kernels on the adjoint state; its "F" buffer ends up holding dR_a = Inv_a^T ds
      // of every axis, which — with the forward spectra F_a of x_l — gives the weight gradients
      FFNO_TRY(umma_spectral_fwd(p->umma_adj, l, w.ds, batch, w.gx, w.dR, w.dF, nullptr, st, true));
      bool want_w = false;
      for (int a = 0; a < p->d.ndim; ++a) want_w |= lg.fourier_weight[a] != nullptr;
      if (want_w) {
        FFNO_TRY(umma_forward_spectra(p->umma, xl, batch, w.F, st));
        for (int a = 0; a < p->d.ndim; ++a) {
          if (!lg.fourier_weight[a]) continue;
          long long outer = batch, p_inner = 1;
          for (int i = 0; i < a; ++i) outer *= p->ext[i];
          for (int i = a + 1; i < p->d.ndim; ++i) p_inner *= p->ext[i];
          const size_t off = umma_spec_offset(p->umma, batch, a);
          FFNO_TRY(launch_mix_wgrad(w.F + off, w.dR + off, lg.fourier_weight[a], outer, p->d.modes[a], p_inner, C,
                                    p->sm_count, st, p->dct_modes[a]));
        }
      }
      continue;
    }
    for (int a = p->d.ndim - 1; a >= 0; --a) {
      long long outer = batch, p_inner = 1;
      for (int i = 0; i < a; ++i) outer *= p->ext[i];
      for (int i = a + 1; i < p->d.ndim; ++i) p_inner *= p->ext[i];
      const int L = p->ext[a], K = p->d.modes[a];
      const long long inner = p_inner * C;
      FFNO_TRY(launch_axis_transform(w.ds, p->d_invT[a], w.dR, outer, L, 2 * K, inner, false, st));
      if (lg.fourier_weight[a]) {
        FFNO_TRY(launch_axis_transform(xl, p->d_fwd[a], w.F, outer, L, 2 * K, inner, false, st));
        FFNO_TRY(launch_mix_wgrad(w.F, w.dR, lg.fourier_weight[a], outer, K, p_inner, C, p->sm_count, st, p->dct_modes[a]));
      }
      FFNO_TRY(launch_transpose(lw.wmix[a], w.wT, 2 * C, 2 * C, K, st));
      FFNO_TRY(launch_mode_mix(w.dR, w.wT, w.dF, outer, K, p_inner, C, st));
      FFNO_TRY(launch_axis_transform(w.dF, p->d_fwdT[a], w.gx, outer, 2 * K, L, inner, true, st));
    }
  }

  // ---- 4. lift (grid_2d.py:157; mesh_3d.py:161-166: grid coordinates appended, linear, zero padding)
  if (lo) {
    if (dx) FFNO_CUDA_CHECK(cudaMemcpyAsync(dx, w.gx, U * 4, cudaMemcpyDeviceToDevice, st));
  } else if (mesh) {
    FFNO_TRY(launch_crop_pad(w.g_dense, w.gx, batch, g, false, st));
    FFNO_TRY(launch_lift_rows(x, w.rows_in, Pin, g, st));
    FFNO_TRY(linear_bwd(p, p->lift, prm->in_proj, grads->in_proj, w.rows_in, w.g_dense, dx ? w.d_rows : nullptr, Pin, w, st));
    if (dx) FFNO_TRY(launch_take_cols(w.d_rows, dx, Pin, p->in_total, p->d.in_features, st));
  } else {
    FFNO_TRY(linear_bwd(p, p->lift, prm->in_proj, grads->in_proj, x, w.gx, dx, P, w, st));
  }
  p->last_launches = g_launch_counter - before;
  (void)O;
  return FFNO_OK;
}

int ffno_block_bwd(ffno_plan* p, const ffno_block_params* prm, const float* x, const float* d_forecast, int32_t batch,
                   const ffno_block_grads* grads, float* dx, void* workspace, size_t workspace_bytes, void* stream) {
  return block_bwd_core(p, prm, x, d_forecast, batch, grads, dx, workspace, workspace_bytes, stream, nullptr);
}

int ffno_layers_bwd(ffno_plan* p, const ffno_block_params* prm, const float* x0, const float* d_out, const float* bias,
                    int32_t batch, const ffno_block_grads* grads, float* dx, float* d_bias, void* workspace,
                    size_t workspace_bytes, void* stream) {
  const LayersOnly lo{bias, d_bias};
  return block_bwd_core(p, prm, x0, d_out, batch, grads, dx, workspace, workspace_bytes, stream, &lo);
}

int ffno_plan_set_backward_mode(ffno_plan* plan, int32_t mode) {
  FFNO_REQUIRE(plan != nullptr, FFNO_ERR_BAD_ARG, "plan is NULL");
  FFNO_REQUIRE(mode >= 0 && mode <= 2, FFNO_ERR_BAD_ARG, "backward mode %d (0 = default, 1 = fast, 2 = fp32)", mode);
  plan->bwd_fp32 = mode == 2;
  plan->bwd_fp32_recompute = mode != 1;
  plan->bwd_fp32_adjoint = false;
  return FFNO_OK;
}

int ffno_rel_l2_bwd(const float* x, const float* y, const float* g_out, int32_t batch, int64_t n, float* dx, void* stream) {
  FFNO_REQUIRE(x && y && g_out && dx && batch >= 0 && n >= 1, FFNO_ERR_BAD_ARG, "bad argument");
  return launch_rel_l2_bwd(x, y, g_out, dx, batch, n, static_cast<cudaStream_t>(stream));
}

int ffno_umma_selftest(const uint16_t* A, const uint16_t* B, float* D, int32_t N, int32_t K, int32_t a_mn,
                       int32_t b_mn, int32_t variant, void* stream) {
  FFNO_REQUIRE(A && B && D, FFNO_ERR_BAD_ARG, "NULL argument");
  return launch_umma_selftest(A, B, D, N, K, a_mn, b_mn, variant, static_cast<cudaStream_t>(stream));
}

int ffno_debug_timeline(int32_t enable, int64_t* host_out) {
  return debug_timeline(enable, reinterpret_cast<long long*>(host_out));
}

int ffno_plan_set_domain(ffno_plan* plan, float length_x, float length_y) {
  FFNO_REQUIRE(plan != nullptr, FFNO_ERR_BAD_ARG, "plan is NULL");
  FFNO_REQUIRE(length_x > 0.f && length_y > 0.f, FFNO_ERR_BAD_ARG, "domain lengths must be positive");
  if (plan->domain[0] != length_x || plan->domain[1] != length_y) plan->g_rollout.reset();   // baked into the graph
  plan->domain[0] = length_x;
  plan->domain[1] = length_y;
  return FFNO_OK;
}

size_t ffno_velocity_scratch_bytes(int32_t batch, int32_t X, int32_t Y) {
  if (batch < 0 || X < 1 || Y < 1) return 0;
  return velocity_scratch_floats(batch, X, Y) * sizeof(float);
}

int ffno_velocity_fwd(const float* w, int64_t stride_b, int64_t stride_xy, int32_t batch, int32_t X, int32_t Y,
                      float length_x, float length_y, float* q, float* v, void* scratch, size_t scratch_bytes,
                      void* stream) {
  FFNO_REQUIRE(batch >= 0, FFNO_ERR_BAD_ARG, "batch=%d", batch);
  if (batch == 0) return FFNO_OK;
  FFNO_REQUIRE(w && q && v && scratch, FFNO_ERR_BAD_ARG, "NULL argument");
  FFNO_REQUIRE(scratch_bytes >= ffno_velocity_scratch_bytes(batch, X, Y), FFNO_ERR_WORKSPACE, "scratch < %zu B",
               ffno_velocity_scratch_bytes(batch, X, Y));
  return launch_velocity(w, stride_b, stride_xy, batch, X, Y, length_x, length_y, q, v, static_cast<float*>(scratch),
                         static_cast<cudaStream_t>(stream));
}

int64_t ffno_plan_last_launch_count(const ffno_plan* plan) { return plan ? plan->last_launches : 0; }

int ffno_debug_pipe_stats(const ffno_plan* plan, uint64_t* host_out, int32_t n_words) {
  FFNO_REQUIRE(plan && plan->use_umma && host_out && n_words > 0, FFNO_ERR_BAD_ARG, "bad argument");
  return umma_pipe_debug(plan->umma, reinterpret_cast<unsigned long long*>(host_out), n_words);
}

int ffno_plan_pipeline_unit(const ffno_plan* plan, int32_t batch) {
  return plan && plan->use_umma && batch > 0 ? umma_pipeline_unit(plan->umma, batch) : 0;
}

int ffno_plan_graph_active(const ffno_plan* plan) {
  return plan && (plan->g_block.exec || plan->g_rollout.exec) ? 1 : 0;
}

}  // extern "C"
