// Fourier positional encoding of the full-resolution voxels (models/agile3d.py:141-161,
// models/position_embedding.py:13-41,123-152).  Two kernels: per-scene xyz range, then sin/cos.
// Precise sinf/cosf (no fast-math): the arguments reach tens of radians and the result feeds fp32 parity.
#include <float.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace ag3d {

__device__ __forceinline__ unsigned enc_f(float f) {  // order-preserving float -> uint
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void range_init_kernel(unsigned* enc, int n_scenes) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_scenes * 6) enc[i] = ((i % 6) < 3) ? 0xFFFFFFFFu : 0u;  // min slots start at +max, max slots at -max
}

// grid (blocks_per_scene, n_scenes)
__global__ void range_kernel(const float* __restrict__ xyz, const int* __restrict__ offsets, unsigned* enc) {
  const int b = blockIdx.y;
  const long long lo = offsets[b], hi = offsets[b + 1];
  float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi;
       i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float v = __ldg(xyz + i * 3 + a);
      mn[a] = fminf(mn[a], v);
      mx[a] = fmaxf(mx[a], v);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) {
      mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
      mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      atomicMin(enc + b * 6 + a, enc_f(mn[a]));
      atomicMax(enc + b * 6 + 3 + a, enc_f(mx[a]));
    }
  }
}

__global__ void range_decode_kernel(const unsigned* enc, float* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = dec_f(enc[i]);
}

// one thread per (voxel, frequency j): out[v][j] = sin(t), out[v][half + j] = cos(t)
__global__ void posenc_kernel(const float* __restrict__ xyz, const int* __restrict__ offsets, int n_scenes,
                              const unsigned* __restrict__ enc, const float* __restrict__ gauss_B, int half,
                              float* __restrict__ out, float* __restrict__ out_split, long long n_total) {
  const int b = blockIdx.y;
  const long long lo = offsets[b], hi = offsets[b + 1];
  float mn[3], inv_den[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    mn[a] = dec_f(enc[b * 6 + a]);
    inv_den[a] = dec_f(enc[b * 6 + 3 + a]) - mn[a];
  }
  const long long total = (hi - lo) * half;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long v = lo + t / half;
    const int j = (int)(t % half);
    float arg = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      // ((xyz - min) * 1) / (max - min) + 0, then * 2*pi   (shift_scale_points + "xyz *= 2*np.pi")
      float u = (__ldg(xyz + v * 3 + a) - mn[a]) / inv_den[a];
      u *= 6.283185307179586f;
      arg = fmaf(u, __ldg(gauss_B + a * half + j), arg);
    }
    float s, c;
    sincosf(arg, &s, &c);
    out[v * (2 * half) + j] = s;
    out[v * (2 * half) + half + j] = c;
    if (out_split) {      // the same row as "split" bf16 pairs: 32-channel slab = 32 hi | 32 lo (csrc/spconv_tc.cu)
      __nv_bfloat16* row = reinterpret_cast<__nv_bfloat16*>(out_split + v * (2 * half));
      const int cs = j, cc = half + j;
      const __nv_bfloat16 sh = __float2bfloat16_rn(s), ch = __float2bfloat16_rn(c);
      row[(cs >> 5) * 64 + (cs & 31)] = sh;
      row[(cs >> 5) * 64 + 32 + (cs & 31)] = __float2bfloat16_rn(s - __bfloat162float(sh));
      row[(cc >> 5) * 64 + (cc & 31)] = ch;
      row[(cc >> 5) * 64 + 32 + (cc & 31)] = __float2bfloat16_rn(c - __bfloat162float(ch));
    }
  }
}

}  // namespace ag3d

using namespace ag3d;

extern "C" {

size_t ag3d_posenc_workspace_bytes(int32_t n_scenes) { return (size_t)(n_scenes * 6 + n_scenes + 1 + 8) * 4; }

static int posenc_launch(const float* xyz, const int32_t* scene_offsets_host, int32_t n_scenes, const float* gauss_B,
                         int32_t d_pos, float* out, float* out_split, float* range_out, void* ws, size_t ws_bytes,
                         ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_scenes >= 1 && n_scenes < 65535, "n_scenes out of range");
  AG3D_CHECK_ARG(d_pos > 0 && d_pos % 2 == 0, "d_pos must be even");
  AG3D_CHECK_ARG(xyz && scene_offsets_host && gauss_B && out, "bad pointers");
  if (!ws || ws_bytes < ag3d_posenc_workspace_bytes(n_scenes)) {
    set_error("posenc workspace too small");
    return AG3D_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  unsigned* enc = static_cast<unsigned*>(ws);
  int* offsets = reinterpret_cast<int*>(enc + n_scenes * 6);
  const long long n_total = scene_offsets_host[n_scenes];
  long long max_scene = 0;
  for (int b = 0; b < n_scenes; ++b) {
    long long len = (long long)scene_offsets_host[b + 1] - scene_offsets_host[b];
    AG3D_CHECK_ARG(len > 0, "empty scene");
    if (len > max_scene) max_scene = len;
  }
  // offsets are tiny (n_scenes+1 ints): pageable async copy is staged by the runtime before returning
  AG3D_CUDA(cudaMemcpyAsync(offsets, scene_offsets_host, (size_t)(n_scenes + 1) * 4, cudaMemcpyHostToDevice, st));
  range_init_kernel<<<(n_scenes * 6 + 127) / 128, 128, 0, st>>>(enc, n_scenes);
  AG3D_LAUNCH_CHECK("range_init");
  int bx = (int)((max_scene + 256 * 8 - 1) / (256 * 8));
  if (bx > sm_count() * 2) bx = sm_count() * 2;
  if (bx < 1) bx = 1;
  range_kernel<<<dim3(bx, n_scenes), 256, 0, st>>>(xyz, offsets, enc);
  AG3D_LAUNCH_CHECK("range");
  if (range_out) {
    range_decode_kernel<<<(n_scenes * 6 + 127) / 128, 128, 0, st>>>(enc, range_out, n_scenes * 6);
    AG3D_LAUNCH_CHECK("range_decode");
  }
  const int half = d_pos / 2;
  long long work = max_scene * half;
  int px = (int)((work + 256 * 4 - 1) / (256 * 4));
  if (px > sm_count() * 8) px = sm_count() * 8;
  if (px < 1) px = 1;
  posenc_kernel<<<dim3(px, n_scenes), 256, 0, st>>>(xyz, offsets, n_scenes, enc, gauss_B, half, out, out_split, n_total);
  AG3D_LAUNCH_CHECK("posenc");
  return AG3D_OK;
}

int ag3d_fourier_posenc(const float* xyz, const int32_t* scene_offsets_host, int32_t n_scenes,
                        const float* gauss_B, int32_t d_pos, float* out, float* range_out, void* ws,
                        size_t ws_bytes, ag3d_stream_t stream) {
  return posenc_launch(xyz, scene_offsets_host, n_scenes, gauss_B, d_pos, out, nullptr, range_out, ws, ws_bytes, stream);
}

int ag3d_fourier_posenc_split(const float* xyz, const int32_t* scene_offsets_host, int32_t n_scenes,
                              const float* gauss_B, int32_t d_pos, float* out, float* out_split, float* range_out,
                              void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(out_split && d_pos % 64 == 0, "split rows need d_pos % 64 == 0");
  return posenc_launch(xyz, scene_offsets_host, n_scenes, gauss_B, d_pos, out, out_split, range_out, ws, ws_bytes, stream);
}

}  // extern "C"
