// Throughput probes behind the design of the pair-packed sparse convolution (spconv_pk.cu): how fast can 8 warps of one
// SM (a) read accumulators out of TMEM (tcgen05.ld 32x32b), (b) permute 32-bit values across lanes (shfl.idx), and
// (c) do both in the pattern of the per-group epilogue (ld 16 columns, two indexed shuffles + adds per column).
// One CTA per SM, cycles from clock64(); prints bytes or instructions per cycle per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/epi_probe tools/probes/epi_probe.cu && /tmp/epi_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// mode 0: x16 loads, wait after each;  1: x32 loads, wait after each;  2: two x16 loads in flight per wait
// 3: shfl.idx only (16 independent shuffles per iteration);  4: epilogue pattern: ld16 + wait + 32 shfl + 32 fadd
// 5: pattern 4 with the next ld16 issued before the shuffles of the current one (software pipelined)
__global__ void __launch_bounds__(256) probe(int mode, int iters, int warps, float* out, long long* cyc) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  const int src0 = (lane * 7 + 3) & 31, src1 = (lane * 13 + 5) & 31;
  __syncthreads();
  const long long t0 = clock64();
  if (warp < warps) {
    if (mode == 0) {
      for (int it = 0; it < iters; ++it) {
        uint32_t r[16];
        ld16(base + ((it * 16) & 255) + (warp >> 2) * 256, r);
        ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += __uint_as_float(r[i]);
      }
    } else if (mode == 1) {
      for (int it = 0; it < iters; ++it) {
        uint32_t r[32];
        ld32(base + ((it * 32) & 255) + (warp >> 2) * 256, r);
        ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] += __uint_as_float(r[i]);
      }
    } else if (mode == 2) {
      for (int it = 0; it < iters; it += 2) {
        uint32_t r[16], s[16];
        ld16(base + ((it * 16) & 255) + (warp >> 2) * 256, r);
        ld16(base + ((it * 16 + 16) & 255) + (warp >> 2) * 256, s);
        ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += __uint_as_float(r[i]) + __uint_as_float(s[i]);
      }
    } else if (mode == 3) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (float)(lane + i);
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] += __shfl_sync(0xffffffffu, v[i], (src0 + it) & 31);
      }
    } else if (mode == 4) {
      for (int it = 0; it < iters; ++it) {
        uint32_t r[16];
        ld16(base + ((it * 16) & 255) + (warp >> 2) * 256, r);
        ld_wait();
        const bool p0 = ((lane + it) & 3) != 0, p1 = ((lane + it) & 7) < 3;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = __shfl_sync(0xffffffffu, __uint_as_float(r[i]), src0);
          const float b = __shfl_sync(0xffffffffu, __uint_as_float(r[i]), src1);
          if (p0) acc[i] += a;
          if (p1) acc[16 + i] += b;
        }
      }
    } else {
      uint32_t r[16], s[16];
      ld16(base + (warp >> 2) * 256, r);
      for (int it = 0; it < iters; it += 2) {
        ld_wait();
        ld16(base + (((it + 1) * 16) & 255) + (warp >> 2) * 256, s);
        const bool p0 = ((lane + it) & 3) != 0, p1 = ((lane + it) & 7) < 3;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = __shfl_sync(0xffffffffu, __uint_as_float(r[i]), src0);
          const float b = __shfl_sync(0xffffffffu, __uint_as_float(r[i]), src1);
          if (p0) acc[i] += a;
          if (p1) acc[16 + i] += b;
        }
        ld_wait();
        ld16(base + (((it + 2) * 16) & 255) + (warp >> 2) * 256, r);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = __shfl_sync(0xffffffffu, __uint_as_float(s[i]), src0);
          const float b = __shfl_sync(0xffffffffu, __uint_as_float(s[i]), src1);
          if (p0) acc[i] += a;
          if (p1) acc[16 + i] += b;
        }
      }
      ld_wait();
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sms * 256 * sizeof(float));
  cudaMallocManaged(&cyc, sms * sizeof(long long));
  const int iters = 4096;
  const char* names[] = {"ldtm x16 (64 B/lane)", "ldtm x32 (128 B/lane)", "ldtm 2 x x16 per wait", "shfl.idx x16",
                         "epilogue pattern (ld16 + 32 shfl + 32 fadd)", "epilogue pattern, ld one chunk ahead"};
  for (int mode = 0; mode < 6; ++mode)
    for (int warps = 4; warps <= 8; warps += 4) {
      probe<<<sms, 256>>>(mode, iters, warps, out, cyc);
      probe<<<sms, 256>>>(mode, iters, warps, out, cyc);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(cudaGetLastError())); return 1; }
      double c = 0;
      for (int b = 0; b < sms; ++b) c += (double)cyc[b];
      c /= sms;
      const double per_it = c / iters;
      if (mode <= 2) {
        const double bytes = (mode == 1 ? 32.0 : 16.0) * 4 * 32 * warps;    // per iteration, all warps
        printf("%-48s warps %d: %8.1f cyc/iter/warp-set  -> %6.1f B/cyc/SM\n", names[mode], warps, per_it, bytes / per_it);
      } else if (mode == 3) {
        printf("%-48s warps %d: %8.1f cyc/iter -> %5.2f warp-shfl/cyc/SM\n", names[mode], warps, per_it, 16.0 * warps / per_it);
      } else {
        printf("%-48s warps %d: %8.1f cyc per 16-column chunk per warp-set (%d warps) -> Cout=96 group: %6.0f cyc\n",
               names[mode], warps, per_it, warps, per_it * 3);
      }
    }
  return 0;
}
