// Shared helpers of libffno_b200: error reporting across the C ABI, launch checks.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/ffno_b200.h"

namespace ffno {

// Thread-local last-error message returned by ffno_last_error().
char* error_buffer();
int set_error(int code, const char* fmt, ...);

#define FFNO_CUDA_CHECK(expr)                                                                  \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::ffno::set_error(FFNO_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                  \
                               cudaGetErrorString(_e), __FILE__, __LINE__);                    \
  } while (0)

#define FFNO_LAUNCH_CHECK(name)                                                                \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess)                                                                     \
      return ::ffno::set_error(FFNO_ERR_CUDA, "launch of %s failed: %s", name,                 \
                               cudaGetErrorString(_e));                                        \
  } while (0)

#define FFNO_REQUIRE(cond, code, ...)                                                          \
  do {                                                                                         \
    if (!(cond)) return ::ffno::set_error(code, __VA_ARGS__);                                  \
  } while (0)

#define FFNO_TRY(expr)                                                                         \
  do {                                                                                         \
    int _s = (expr);                                                                           \
    if (_s != FFNO_OK) return _s;                                                              \
  } while (0)

inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Opt a kernel into `bytes` of dynamic shared memory on the CURRENT device.  The attribute belongs to the
// (function, device) pair, so what has been granted is remembered per pair, under a mutex: plans on several GPUs of
// one process each get their opt-in, and concurrent callers do not race (include/ffno_b200.h: re-entrancy).
int ensure_dynamic_smem(const void* func, size_t bytes);
template <class... A>
int ensure_dynamic_smem(void (*kernel)(A...), size_t bytes) {
  return ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), bytes);
}

}  // namespace ffno
