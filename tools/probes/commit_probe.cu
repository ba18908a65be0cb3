// What does one tcgen05.commit cost?  (a) latency: commit -> mbarrier phase completes (no MMAs pending);
// (b) throughput: back-to-back commits onto 8 barriers, waiting only every 8th; (c) mbarrier.try_wait on a barrier whose
// phase already completed; (d) plain mbarrier.arrive + try_wait round trip inside one thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/commit_probe tools/probes/commit_probe.cu && /tmp/commit_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) { while (!try_wait(bar, parity)) {} }
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__global__ void probe(int iters, long long* out) {
  __shared__ uint64_t bars[8];
  __shared__ uint32_t slot;
  const uint32_t b0 = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0)
    for (int i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b0 + 8 * i));
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { commit(b0); wait(b0, i & 1); }                    // (a)
    long long t1 = clock64();
    for (int i = 0; i < iters; ++i) {                                                    // (b)
      for (int j = 0; j < 8; ++j) commit(b0 + 8 * j);
      wait(b0 + 56, i & 1);
    }
    // barrier 0 has completed iters + iters phases; barriers 1..7 iters phases
    long long t2 = clock64();
    uint32_t acc = 0;
    const uint32_t par = ((iters - 1) & 1);
    for (int i = 0; i < iters; ++i) acc += try_wait(b0 + 8, par);                       // (c) already complete
    long long t3 = clock64();
    for (int i = 0; i < iters; ++i) { arrive(b0 + 16); wait(b0 + 16, (iters + i) & 1); } // (d)
    long long t4 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(slot) : "memory");
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 8 * sizeof(long long));
  const int iters = 2000;
  probe<<<1, 64>>>(iters, out);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  printf("commit -> wait round trip      : %.1f cycles\n", (double)out[0] / iters);
  printf("8 commits + 1 wait             : %.1f cycles per commit\n", (double)out[1] / iters / 8);
  printf("try_wait on a completed phase  : %.1f cycles (hits %lld)\n", (double)out[2] / iters, out[4]);
  printf("arrive -> wait round trip      : %.1f cycles\n", (double)out[3] / iters);
  return 0;
}
