// Probe of cp.async.bulk.tensor.2d ... tile::gather4 on sm_100a: which tensor-map box works, where the four rows land
// in shared memory under SWIZZLE_NONE / SWIZZLE_128B, and what a negative / out-of-range row index produces.
// nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/tma_gather_probe tools/probes/tma_gather_probe.cu && /tmp/tma_gather_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, int r0, int r1, int r2, int r3, int col, uint16_t* out,
                      uint32_t expect) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4096);
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar);
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) reinterpret_cast<uint16_t*>(smem)[i] = 0xFFFF;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(expect) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(&tm), "r"(bar_a), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
        : "memory");
  }
  // bounded wait
  uint32_t ok = 0;
  for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar_a), "r"(0u)
        : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(smem)[i];
  if (threadIdx.x == 0) out[2048] = (uint16_t)ok;
}

int main() {
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  const int R = 64, C = 128;   // rows x bf16 columns (256 B per row)
  std::vector<uint16_t> h(R * C);
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) h[r * C + c] = (uint16_t)(r * 256 + c);   // value = row<<8 | col
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2);
  cudaMalloc(&o, 4098 * 2);
  cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
  for (int sw = 0; sw < 2; ++sw)
    for (int boxrows = 1; boxrows <= 1; ++boxrows) {   // box {64,4} raises an illegal-instruction error (measured)
      CUtensorMap tm;
      cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)R};
      cuuint64_t strides[1] = {(cuuint64_t)C * 2};
      cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
      cuuint32_t estr[2] = {1, 1};
      CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("== swizzle %s, box {64,%d}: encode rc %d\n", sw ? "128B" : "none", boxrows, (int)rc);
      if (rc != CUDA_SUCCESS) continue;
      for (int variant = 0; variant < 2; ++variant) {
        const int rows[4] = {5, variant ? -1 : 7, 9, variant ? 1000 : 11};
        cudaMemset(o, 0, 4098 * 2);
        probe<<<1, 32, 8192>>>(tm, rows[0], rows[1], rows[2], rows[3], 64, o, 4 * 128);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<uint16_t> r(2049);
        cudaMemcpy(r.data(), o, 2049 * 2, cudaMemcpyDeviceToHost);
        printf("  rows {%d,%d,%d,%d} col 64: %s, barrier %s\n", rows[0], rows[1], rows[2], rows[3], cudaGetErrorString(e),
               r[2048] ? "completed" : "TIMEOUT");
        if (e != cudaSuccess) return 2;
        for (int line = 0; line < 8; ++line) {       // first 8 x 128 B of the destination, one 16-byte chunk = first value
          printf("   +%4d:", line * 128);
          for (int ch = 0; ch < 8; ++ch) {
            const uint16_t v = r[line * 64 + ch * 8];
            if (v == 0xFFFF) printf("  ----");
            else printf(" %2d:%3d", v >> 8, v & 0xFF);
          }
          printf("\n");
        }
      }
    }
  // ---- does an absent row (index -1) touch memory below the tensor base?  Put the tensor at the start of its own
  //      large allocation (the virtual addresses below it are not mapped) and gather {-1, 0, -1, 1}.
  {
    uint16_t* big;
    cudaMalloc(&big, (size_t)64 << 20);
    cudaMemset(big, 0x11, (size_t)64 << 20);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)0x7FFFFFFF};
    cuuint64_t strides[1] = {(cuuint64_t)C * 2};
    cuuint32_t box[2] = {64, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, big, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== tensor at the start of a 64 MB allocation %p, rows {-1,0,-1,1}: encode rc %d\n", (void*)big, (int)rc);
    for (int rep = 0; rep < 3; ++rep) {
      probe<<<1, 32, 8192>>>(tm, -1, 0, -1 - rep * 1000, 1, 0, o, 4 * 128);
      cudaError_t e = cudaDeviceSynchronize();
      printf("   rep %d: %s\n", rep, cudaGetErrorString(e));
      if (e != cudaSuccess) return 3;
    }
  }
  return 0;
}
