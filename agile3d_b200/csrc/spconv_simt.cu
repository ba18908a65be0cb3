// K4 (exact-fp32 variant): output-stationary gather implicit GEMM on the CUDA cores, plus the 3->32 stem
// evaluated directly against the hash table.  This is the bit-for-bit-fp32 arithmetic path (FFMA, fp32
// accumulate) that the tcgen05 3xTF32 path (spconv_tc.cu) is validated against at full size.
#include <cuda_bf16.h>

#include "common.cuh"

namespace ag3d {

// out tile BM x BN per CTA, 256 threads, thread tile 4 x TN, K-chunk 16 channels.
constexpr int BM = 64;
constexpr int BK = 16;
constexpr int SIMT_THREADS = 256;

template <int BN>
__global__ void __launch_bounds__(SIMT_THREADS)
spconv_simt_kernel(const float* __restrict__ in, int in_ld, int cin, const int* __restrict__ nbr, int K,
                   long long n_out, const float* __restrict__ weight, int cout, const float* __restrict__ scale,
                   const float* __restrict__ shift, const float* __restrict__ residual, int res_ld,
                   float* __restrict__ out, int out_ld, int flags) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN];
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const long long row0 = (long long)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;

  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  // gather role: thread -> (row a_r of the tile, 4-float chunk a_c of the 16-channel slab)
  const int a_r = tid >> 2, a_c = tid & 3;
  const long long my_row = row0 + a_r;
  // weight role: thread -> (kk, 4-float chunk)
  constexpr int B_VEC = BK * BN / 4;  // float4 per slab
  const int b_kk = tid / (BN / 4), b_c = tid % (BN / 4);

  for (int k = 0; k < K; ++k) {
    int src = -1;
    if (my_row < n_out) src = nbr ? __ldg(nbr + (long long)k * n_out + my_row) : (int)my_row;
    if (!__syncthreads_or(src >= 0)) continue;  // no voxel of this tile has a neighbour at offset k
    const float* wk = weight + (long long)k * cin * cout;
    for (int c0 = 0; c0 < cin; c0 += BK) {
      float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src >= 0) av = __ldg(reinterpret_cast<const float4*>(in + (long long)src * in_ld + c0 + a_c * 4));
      As[a_c * 4 + 0][a_r] = av.x;
      As[a_c * 4 + 1][a_r] = av.y;
      As[a_c * 4 + 2][a_r] = av.z;
      As[a_c * 4 + 3][a_r] = av.w;
      if (tid < B_VEC) {
        const float4 bv =
            __ldg(reinterpret_cast<const float4*>(wk + (long long)(c0 + b_kk) * cout + col0 + b_c * 4));
        *reinterpret_cast<float4*>(&Bs[b_kk][b_c * 4]) = bv;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        float b[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
        const float ar[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(ar[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  // fused epilogue: folded BatchNorm / bias, residual, ReLU, write into a channel slice
  const bool relu = flags & AG3D_RELU;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = row0 + ty * 4 + i;
    if (r >= n_out) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int c = col0 + tx * TN + j;
      float v = acc[i][j];
      if (scale) v *= __ldg(scale + c);
      if (shift) v += __ldg(shift + c);
      if (residual) v += __ldg(residual + r * res_ld + c);
      if (relu) v = fmaxf(v, 0.f);
      out[r * out_ld + c] = v;
    }
  }
}

// Stem: one warp per output voxel, lane = output channel.  The 125 offsets are probed 32 at a time.
constexpr int STEM_CIN = 3;
constexpr int STEM_COUT = 32;

template <bool BRICKS>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const int4* __restrict__ coords, const float* __restrict__ feats, long long n,
                 const Slot* __restrict__ table, unsigned long long mask, const int* __restrict__ brick_rows, int ksize,
                 const float* __restrict__ weight, const float* __restrict__ scale,
                 const float* __restrict__ shift, float* __restrict__ out, int out_ld, int flags) {
  extern __shared__ __align__(16) float w_s[];  // [K][3][32] weights, then the warps' hit lists [8][128] float4
  const int K = ksize * ksize * ksize;
  float4* hit_s = reinterpret_cast<float4*>(w_s + ((K * STEM_CIN * STEM_COUT + 3) & ~3));
  for (int i = threadIdx.x; i < K * STEM_CIN * STEM_COUT; i += blockDim.x) w_s[i] = __ldg(weight + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int half = ksize / 2;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long step = (long long)gridDim.x * (blockDim.x >> 5);
  for (; row < n; row += step) {
    const int4 c = __ldg(coords + row);
    float acc = 0.f;
    // All probes of the voxel are issued before any result is used (up to 4 independent hash chains per lane), then
    // every lane fetches the 3 features of its own hits (independent loads), and only then the hits are folded in
    // ascending offset order: no global load sits on the serial accumulation chain.
    int src[4];
    float f0[4], f1[4], f2[4];
    if constexpr (BRICKS) {     // `table` is the tensor-stride-4 table: 8 probes + reads of the bricks' row lists
      brick_window_find(c, lane, ksize, K, table, mask, brick_rows, src);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = j * 32 + lane;
        src[j] = -1;
        if (k < K) {
          int r = k;
          const int jx = r % ksize; r /= ksize;
          const int jy = r % ksize; r /= ksize;
          const int x = c.y + jx - half, y = c.z + jy - half, z = c.w + r - half;
          if (coord_in_range(c.x, x, y, z)) src[j] = table_find(table, mask, pack_key(c.x, x, y, z));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f0[j] = f1[j] = f2[j] = 0.f;
      if (src[j] >= 0) {
        const float* f = feats + (long long)src[j] * STEM_CIN;
        f0[j] = __ldg(f + 0); f1[j] = __ldg(f + 1); f2[j] = __ldg(f + 2);
      }
    }
    // The hits are compacted, in ascending offset order, into the warp's list in shared memory (features + weight row
    // offset); the fold is then one broadcast LDS.128 + three weight LDS + three FMAs per hit.
    {
      float4* list = hit_s + (threadIdx.x >> 5) * 128;
      int base = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const unsigned hits = __ballot_sync(0xffffffffu, src[j] >= 0);
        if (src[j] >= 0)
          list[base + __popc(hits & ((1u << lane) - 1u))] =
              make_float4(f0[j], f1[j], f2[j], __int_as_float((j * 32 + lane) * STEM_CIN * STEM_COUT));
        base += __popc(hits);
      }
      __syncwarp();
      for (int i = 0; i < base; ++i) {
        const float4 e = list[i];
        const float* w = w_s + __float_as_int(e.w) + lane;
        acc = fmaf(e.x, w[0], acc);
        acc = fmaf(e.y, w[STEM_COUT], acc);
        acc = fmaf(e.z, w[2 * STEM_COUT], acc);
      }
      __syncwarp();
    }
    float v = acc;
    if (scale) v *= __ldg(scale + lane);
    if (shift) v += __ldg(shift + lane);
    if (flags & AG3D_RELU) v = fmaxf(v, 0.f);
    if (flags & AG3D_OUT_SPLIT) {      // bf16 hi/lo pair rows: 32-channel slab = 64 B hi | 64 B lo
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(out + row * out_ld) + lane;
      dst[0] = hi;
      dst[32] = lo;
    } else {
      out[row * out_ld + lane] = v;
    }
  }
}

int spconv_tc_launch(const float* in, long long n_in, int in_ld, int cin, const int* nbr, int K, long long n_out,
                     const void* wprep, int cout, const float* scale, const float* shift, const float* residual,
                     int res_ld, float* out, int out_ld, int flags, void* ws, size_t ws_bytes, cudaStream_t st);
bool spconv_tc_supported(int cin, int cout);
int spconv_pk_launch(const float* in, long long n_in, int in_ld, int cin, const int* nbr, int K, long long n_out,
                     const void* wprep, int cout, const float* scale, const float* shift, const float* residual,
                     int res_ld, float* out, int out_ld, int flags, cudaStream_t st);
bool spconv_pk_supported(int cin, int cout, int K);
bool spconv_pk_preferred(long long n_out, int K, int cin, int cout);
size_t spconv_tc_workspace_bytes(long long n_out, int K, int cout);

}  // namespace ag3d

using namespace ag3d;

extern "C" {

int ag3d_spconv_fwd_rows(const float* in, int64_t n_in, int32_t in_ld, int32_t cin, const int32_t* nbr, int32_t K,
                         int64_t n_out, const float* weight, const void* weight_tc, int32_t cout, const float* scale,
                         const float* shift, const float* residual, int32_t res_ld, float* out, int32_t out_ld,
                         int32_t flags, int32_t algo, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_out > 0 && n_out < 2147483647LL, "row count out of range");
  AG3D_CHECK_ARG(n_in >= 0 && n_in < 2147483647LL && (nbr || n_in == 0 || n_in == n_out), "input row count");
  AG3D_CHECK_ARG(cin > 0 && cin % 32 == 0 && cout > 0 && cout % 32 == 0, "cin and cout must be multiples of 32");
  AG3D_CHECK_ARG(K >= 1 && (nbr || K == 1), "K > 1 needs a neighbour table");
  AG3D_CHECK_ARG(in && out && aligned16(in) && aligned16(out), "bad pointers");
  AG3D_CHECK_ARG(weight || weight_tc, "need weight (fp32) or weight_tc (prepared)");
  AG3D_CHECK_ARG(in_ld % 4 == 0 && out_ld % 4 == 0 && in_ld >= cin && out_ld >= cout, "leading dims");
  AG3D_CHECK_ARG(!residual || (res_ld >= cout && aligned16(residual) && res_ld % 4 == 0), "residual leading dim");
  cudaStream_t st = as_stream(stream);
  if (algo == AG3D_ALGO_AUTO) {
    algo = (weight_tc && K <= 32 && spconv_tc_supported(cin, cout)) ? AG3D_ALGO_TC : AG3D_ALGO_SIMT;
    if (algo == AG3D_ALGO_TC && nbr && n_in > 0 && (flags & AG3D_IN_SPLIT) && spconv_pk_supported(cin, cout, K) &&
        spconv_pk_preferred(n_out, K, cin, cout))
      algo = AG3D_ALGO_TC_PACKED;
  }
  if (algo == AG3D_ALGO_TC_PACKED)
    return spconv_pk_launch(in, n_in, in_ld, cin, nbr, K, n_out, weight_tc, cout, scale, shift, residual, res_ld, out,
                            out_ld, flags, st);
  if (algo == AG3D_ALGO_TC) {
    AG3D_CHECK_ARG(spconv_tc_supported(cin, cout), "shape not supported by the tcgen05 path");
    return spconv_tc_launch(in, n_in, in_ld, cin, nbr, K, n_out, weight_tc, cout, scale, shift, residual, res_ld, out,
                            out_ld, flags, ws, ws_bytes, st);
  }
  AG3D_CHECK_ARG(algo == AG3D_ALGO_SIMT, "unknown algo");
  AG3D_CHECK_ARG(weight && aligned16(weight), "the fp32 path needs the fp32 weight");
  AG3D_CHECK_ARG(!(flags & (AG3D_IN_SPLIT | AG3D_OUT_SPLIT | AG3D_RES_SPLIT)), "the fp32 path takes fp32 feature rows only");
  const unsigned gx = (unsigned)((n_out + BM - 1) / BM);
  if (cout % 64 == 0) {
    spconv_simt_kernel<64><<<dim3(gx, cout / 64), SIMT_THREADS, 0, st>>>(
        in, in_ld, cin, nbr, K, n_out, weight, cout, scale, shift, residual, res_ld, out, out_ld, flags);
  } else {
    spconv_simt_kernel<32><<<dim3(gx, cout / 32), SIMT_THREADS, 0, st>>>(
        in, in_ld, cin, nbr, K, n_out, weight, cout, scale, shift, residual, res_ld, out, out_ld, flags);
  }
  AG3D_LAUNCH_CHECK("spconv_simt");
  return AG3D_OK;
}

int ag3d_spconv_fwd(const float* in, int32_t in_ld, int32_t cin, const int32_t* nbr, int32_t K, int64_t n_out,
                    const float* weight, const void* weight_tc, int32_t cout, const float* scale, const float* shift,
                    const float* residual, int32_t res_ld, float* out, int32_t out_ld, int32_t flags,
                    int32_t algo, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  // input row count unknown (0): the tensor-core path then gathers with cp.async instead of the TMA engine
  return ag3d_spconv_fwd_rows(in, nbr ? 0 : n_out, in_ld, cin, nbr, K, n_out, weight, weight_tc, cout, scale, shift,
                              residual, res_ld, out, out_ld, flags, algo, ws, ws_bytes, stream);
}

int ag3d_spconv_bwd_data(const float* dout, int32_t dout_ld, int32_t cout, const int32_t* nbr_t, int32_t K,
                         int64_t n_in, const float* weight_t, const void* weight_t_tc, int32_t cin,
                         const float* residual, int32_t res_ld, float* din, int32_t din_ld, int32_t algo, void* ws,
                         size_t ws_bytes, ag3d_stream_t stream) {
  // the data gradient of a sparse convolution is the sparse convolution over the transposed map with W[k]^T
  return ag3d_spconv_fwd(dout, dout_ld, cout, nbr_t, K, n_in, weight_t, weight_t_tc, cin, nullptr, nullptr, residual,
                         res_ld, din, din_ld, 0, algo, ws, ws_bytes, stream);
}

size_t ag3d_spconv_workspace_bytes(int64_t n_out, int32_t K, int32_t cin, int32_t cout) {
  (void)cin;
  if (n_out <= 0 || K < 1 || K > 32 || cout % 32 != 0 || cout < 32 || cout > 256) return 0;
  return spconv_tc_workspace_bytes(n_out, K, cout);
}

static int stem_conv_launch(const int32_t* coords, const float* feats, int64_t n, const void* table, int64_t cap,
                            const int32_t* brick_rows, int32_t ksize, const float* weight, const float* scale,
                            const float* shift, float* out, int32_t out_ld, int32_t flags, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n > 0 && n < 2147483647LL, "row count out of range");
  AG3D_CHECK_ARG(ksize == 1 || ksize == 3 || ksize == 5, "stem kernel size must be 1, 3 or 5");
  AG3D_CHECK_ARG(coords && aligned16(coords) && feats && weight && out, "bad pointers");
  AG3D_CHECK_ARG(table && aligned16(table) && cap >= 2 && (cap & (cap - 1)) == 0, "bad hash table");
  AG3D_CHECK_ARG(out_ld >= STEM_COUT, "out_ld");
  const int K = ksize * ksize * ksize;
  const size_t smem = (size_t)((K * STEM_CIN * STEM_COUT + 3) & ~3) * sizeof(float) + 8 * 128 * sizeof(float4);
  static bool attr_done = false;
  if (!attr_done) {
    AG3D_CUDA(cudaFuncSetAttribute(stem_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    AG3D_CUDA(cudaFuncSetAttribute(stem_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
    attr_done = true;
  }
  long long blocks = (n + 7) / 8;
  const long long cap_blocks = (long long)sm_count() * 8;
  if (blocks > cap_blocks) blocks = cap_blocks;
  if (brick_rows)
    stem_conv_kernel<true><<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(
        reinterpret_cast<const int4*>(coords), feats, n, static_cast<const Slot*>(table),
        (unsigned long long)(cap - 1), brick_rows, ksize, weight, scale, shift, out, out_ld, flags);
  else
    stem_conv_kernel<false><<<(unsigned)blocks, 256, smem, as_stream(stream)>>>(
        reinterpret_cast<const int4*>(coords), feats, n, static_cast<const Slot*>(table),
        (unsigned long long)(cap - 1), nullptr, ksize, weight, scale, shift, out, out_ld, flags);
  AG3D_LAUNCH_CHECK("stem_conv");
  return AG3D_OK;
}

int ag3d_stem_conv_fwd(const int32_t* coords, const float* feats, int64_t n, const void* table, int64_t cap,
                       int32_t ksize, const float* weight, const float* scale, const float* shift, float* out,
                       int32_t out_ld, int32_t flags, ag3d_stream_t stream) {
  return stem_conv_launch(coords, feats, n, table, cap, nullptr, ksize, weight, scale, shift, out, out_ld, flags, stream);
}

int ag3d_stem_conv_fwd_bricks(const int32_t* coords, const float* feats, int64_t n, const void* table2, int64_t cap2,
                              const int32_t* brick_rows, int32_t ksize, const float* weight, const float* scale,
                              const float* shift, float* out, int32_t out_ld, int32_t flags, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(brick_rows && (ksize == 3 || ksize == 5), "brick stem: brick_rows and kernel size 3 or 5");
  return stem_conv_launch(coords, feats, n, table2, cap2, brick_rows, ksize, weight, scale, shift, out, out_ld, flags,
                          stream);
}

}  // extern "C"
