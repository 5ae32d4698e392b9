// tcgen05 (UMMA) fast path — placeholder until the kernels land; every shape reports "unsupported"
// so plans fall back to the generic FP32 CUDA kernels (still CUDA, never CPU).
#include "umma_path.cuh"

namespace ffno {
struct UmmaState {};
bool umma_supported(const ffno_desc*, const int*) { return false; }
const char* umma_why_not(const ffno_desc*, const int*) { return "tcgen05 path not built yet"; }
int umma_create(UmmaState**, const ffno_desc*, const int*) { return set_error(FFNO_ERR_UNSUPPORTED, "no umma"); }
void umma_destroy(UmmaState*) {}
int umma_load_params(UmmaState*, const UmmaLayerSrc*, float* const*, float* const*, cudaStream_t) { return FFNO_OK; }
size_t umma_workspace_floats(const UmmaState*, int) { return 0; }
int umma_layer_fwd(UmmaState*, int, const float*, int, float*, float*, float*, float*, float*, float*, bool, bool,
                   cudaStream_t) { return set_error(FFNO_ERR_UNSUPPORTED, "no umma"); }
int umma_spectral_fwd(UmmaState*, int, const float*, int, float*, float*, float*, float*, cudaStream_t) {
  return set_error(FFNO_ERR_UNSUPPORTED, "no umma"); }
int umma_ff_fwd(UmmaState*, int, const float*, const float*, int, float*, float*, cudaStream_t) {
  return set_error(FFNO_ERR_UNSUPPORTED, "no umma"); }
}  // namespace ffno
