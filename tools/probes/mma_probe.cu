// How long does a chain of small tcgen05.mma (M = 128, N = 32..128, K = 16, bf16, operands in shared memory) take when the
// MMAs accumulate into ONE TMEM accumulator vs when consecutive MMAs alternate between 2 / 3 / 4 accumulators?
// (the sparse convolutions issue six such MMAs per 32-channel stage into the same accumulator)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mma_probe tools/probes/mma_probe.cu && /tmp/mma_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe(int n, int n_acc, int iters, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0));
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem), sb = sa + 16384;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = desc(sa, 16, 1024, 2);                 // SWIZZLE_128B A tile, as the convolution uses it
    const uint64_t db = desc(sb, (uint32_t)n * 16, 128, 0);    // no-swizzle weight stage
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 6; ++j) umma(slot + (uint32_t)((j % n_acc) * 128), da + (j & 1) * 2, db, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(b0) : "memory");
    const long long t1 = clock64();
    while (!try_wait(b0, 0)) {}
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 2 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int iters = 2000;
  for (int n : {32, 64, 96, 128, 256})
    for (int n_acc : {1, 2, 3, 4}) {
      if (n_acc * 128 > 512 || (n > 128 && n_acc > 2)) continue;
      probe<<<1, 128, 65536>>>(n, n_acc, iters, out);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
      printf("N = %3d, %d accumulator(s): issue %.1f cycles per MMA, issue + drain %.1f cycles per MMA (ideal %d)\n", n, n_acc,
             (double)out[0] / iters / 6, (double)out[1] / iters / 6, 128 * n / 256);
    }
  return 0;
}
