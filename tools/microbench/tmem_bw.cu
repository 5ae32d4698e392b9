// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM as a function of the number of issuing warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bw tmem_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NCOL>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t* v);
template <>
__device__ __forceinline__ void ld<32>(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      :
      : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
        "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
        "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
        "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// mode 0: ld only; 1: st only
__global__ void k(int iters, int mode, long long* out_cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64 % 512);
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  st32(base, v);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
      ld<32>(base + (uint32_t)((it & 1) * 32), v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc += v[it & 31];
    } else {
      v[0] = acc + it;
      st32(base + (uint32_t)((it & 1) * 32), v);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  long long* d_c;
  uint32_t* d_s;
  cudaMalloc(&d_c, 148 * 8);
  cudaMalloc(&d_s, 148 * 1024 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 4, 8, 16}) {
      k<<<1, warps * 32>>>(iters, mode, d_c, d_s);
      k<<<1, warps * 32>>>(iters, mode, d_c, d_s);
      cudaError_t e = cudaDeviceSynchronize();
      long long c = 0;
      cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)iters * warps * 32 * 32 * 4;
      printf("%s warps=%2d  cycles/iter=%.1f  bytes/cycle/SM=%.1f  (%s)\n", mode ? "tcgen05.st" : "tcgen05.ld", warps,
             (double)c / iters, bytes / c, cudaGetErrorString(e));
    }
  return 0;
}
