// Throughput of tile::gather4 row gathers (the sparse-conv A-operand pattern): persistent CTAs, a ring of 16 KB stages
// (128 rows x 128 B), one producer warp (lane l gathers rows 4l..4l+3 of the stage), one consumer warp that only
// releases stages.  Rows come from a random neighbour table with ~46 % valid entries (-1 = zero fill).
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_gather_bw tools/probes/tma_gather_bw.cu && /tmp/tma_gather_bw
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
  unsigned spins = 0;
  while (!try_wait(bar, parity)) if (++spins > (1u << 24)) __trap();
}

// stages_total stages, stage s of CTA b = global stage b + s * gridDim.x; idx: [stages_total][128] rows
__global__ void __launch_bounds__(288) gather_bw(const __grid_constant__ CUtensorMap tm, const int* __restrict__ idx,
                                                 int stages_total, int S, int slabs, unsigned long long* sink, int W) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * 16384);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  const uint32_t st0 = (uint32_t)__cvta_generic_to_shared(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s));          // full
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * (S + s)));    // empty
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  int n = 0;
  if (warp < W) {
    // producer warp w takes the CTA's stages n with n % W == w
    // neighbour rows are fetched AHEAD stages before they are needed (register ring), as a real producer would
    constexpr int AHEAD = 8;
    int4 ring[AHEAD];
#pragma unroll
    for (int a = 0; a < AHEAD; ++a) {
      const long long g = blockIdx.x + (long long)(a * W + warp) * gridDim.x;
      ring[a] = g < stages_total ? __ldg(reinterpret_cast<const int4*>(idx + (size_t)g * 128) + lane) : make_int4(-1, -1, -1, -1);
    }
    n = warp;
    for (long long g = blockIdx.x + (long long)warp * gridDim.x; g < stages_total;) {
#pragma unroll
      for (int a = 0; a < AHEAD; ++a) {
        if (g < stages_total) {
          const int s = n % S;
          const int4 r = ring[a];
          const long long gn = g + (long long)AHEAD * W * gridDim.x;
          if (gn < stages_total) ring[a] = __ldg(reinterpret_cast<const int4*>(idx + (size_t)gn * 128) + lane);
          wait(bar0 + 8 * (S + s), ((n / S) & 1) ^ 1);
          if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(16384u) : "memory");
          __syncwarp();
          const int col = (g % slabs) * 64;
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
              " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(st0 + s * 16384 + lane * 512),
              "l"(&tm), "r"(bar0 + 8 * s), "r"(col), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w)
              : "memory");
          g += (long long)W * gridDim.x;
          n += W;
        }
      }
    }
  } else if (warp == W) {
    unsigned long long acc = 0;
    for (int g = blockIdx.x; g < stages_total; g += gridDim.x, ++n) {
      const int s = n % S;
      wait(bar0 + 8 * s, (n / S) & 1);
      acc += *reinterpret_cast<const unsigned long long*>(smem + (size_t)s * 16384 + lane * 8);
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8 * (S + s)) : "memory");
    }
    if (acc == 0x1234567ull) *sink = acc;
  }
}

int main() {
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q);
  const int R = 150000, C = 192;                       // 150k rows x 96 channels as bf16 hi/lo pairs (384 B per row)
  const int slabs = 3, K = 27;
  const int tiles = (R + 127) / 128;
  const int stages_total = tiles * K * slabs;
  uint16_t* d;
  cudaMalloc(&d, (size_t)R * C * 2);
  cudaMemset(d, 1, (size_t)R * C * 2);
  std::vector<int> h((size_t)stages_total * 128);
  srand(3);
  for (int t = 0; t < tiles; ++t)
    for (int k = 0; k < K; ++k) {
      std::vector<int> rows(128);
      for (int r = 0; r < 128; ++r) rows[r] = (rand() % 100 < 46) ? (int)(((long long)rand() * 32768 + rand()) % R) : -1;
      for (int c = 0; c < slabs; ++c) {
        const size_t g = ((size_t)t * K + k) * slabs + c;
        for (int r = 0; r < 128; ++r) h[g * 128 + r] = rows[r];
      }
    }
  int* di;
  cudaMalloc(&di, h.size() * 4);
  cudaMemcpy(di, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  unsigned long long* sink;
  cudaMalloc(&sink, 8);
  double valid = 0;
  for (int v : h) valid += v >= 0;
  for (int sw = 1; sw < 2; ++sw) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rc); return 1; }
    for (int occ = 1; occ <= 2; ++occ)
     for (int W = 1; W <= 8; W *= 2)
      for (int S = (occ == 1 ? 8 : 4); S <= (occ == 1 ? 8 : 6); S += 2) {
        const size_t smem = (size_t)S * 16384 + 256;
        cudaFuncSetAttribute(gather_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        const int grid = 148 * occ;
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        gather_bw<<<grid, 32 * (W + 1), smem>>>(tm, di, stages_total, S, slabs, sink, W);
        cudaEventRecord(e0);
        for (int it = 0; it < 5; ++it) gather_bw<<<grid, 32 * (W + 1), smem>>>(tm, di, stages_total, S, slabs, sink, W);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= 5;
        printf("swizzle %d  CTAs/SM %d  producer warps %d  stages %2d: %s  %.4f ms  %.0f stages/ms  valid rows %.1f GB/s  (all rows %.1f GB/s)\n", sw, occ, W, S,
               cudaGetErrorString(e), ms, stages_total / ms, valid * 128 / ms / 1e6, (double)stages_total * 16384 / ms / 1e6);
        if (e != cudaSuccess) return 2;
      }
  }
  return 0;
}
