// Weight gradient of the sparse convolution on the tensor cores:
//
//   dW[k][ci][co] = sum_o  in[nbr[k][o]][ci] * dout[o][co]
//
// Both operands are "split" rows (every 32-channel slab = 64 B bf16 hi | 64 B bf16 lo, x = hi + lo), i.e. plain bf16
// vectors of length 2C, and the contraction runs over the rows, so the kernel is ONE bf16 GEMM per (offset, slab)
// pair with both operands MN-major:
//
//   D[(combo q, hi|lo, ci)][(slab s, hi|lo, co)] = sum_o A[o][...] * B[o][...]         (fp32 accumulators in TMEM)
//   dW = D[hi,hi] + D[hi,lo] + D[lo,hi] + D[lo,lo]                                      (all four products: "bf16x4")
//
// A CTA owns two combos (offset k, 32-channel input slab) = M 128, an output-channel chunk of <= 128 channels
// (N = 2 * chunk <= 256) and a range of rows.  Per WG_ROWS-row stage the TMA engine gathers the two A tiles through the
// neighbour table and the B tiles from consecutive rows (tile::gather4, SWIZZLE_128B: the [rows x 128 B] tiles are
// exactly the MN-major canonical layout, SBO = 1024 between 8-row groups, LBO = one tile between 64-element atoms), one
// elected thread issues WG_ROWS / 16 MMAs (K = 16 rows each), and the epilogue adds the two column halves in registers and writes
// the hi-row / lo-row partial sums, which the deterministic split reduction folds together with the row splits.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace ag3d {

int launch_split_reduce(const float* part, int splits, long long count, long long total, int accumulate, float* dst,
                        cudaStream_t st);

constexpr int WG_PROD_WARPS = 8;
constexpr int WG_THREADS = WG_PROD_WARPS * 32 + 32;     // + MMA warp
constexpr int WG_ROWS = 128;                              // rows (= GEMM K) per stage.  Measured on 150k-row layers: 64-row
                                                          // stages with a deeper ring are 1.4x SLOWER (per-stage signalling)
constexpr uint32_t WG_TILE = WG_ROWS * 128;               // [WG_ROWS x 128 B]
constexpr int WG_GROUPS = WG_ROWS / 4;                    // gather4 instructions per A tile
constexpr int WG_LANES = 2 * WG_GROUPS / WG_PROD_WARPS;   // ... spread over the warps: active lanes per warp

struct WgParams {
  const int* nbr; int K; long long n_out;
  int cin, cout, n_slab_in;
  long long rows_per_split;      // multiple of 128
  int splits;
  int NS;                        // ring stages
  float* partial;                // [K][2 * splits][cin][cout]
};

__device__ __forceinline__ void wg_gather4(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int col, int r0, int r1,
                                           int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_dy, const WgParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  // [0, 256): barriers full[NS] | empty[NS] | acc_full; [256, 260): TMEM base; tiles from the next 1024-byte boundary
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  const uint32_t bar_base = smem_u32(bars);
  auto full = [&](int s) { return bar_base + 8u * s; };
  auto empty = [&](int s) { return bar_base + 8u * (8 + s); };
  const uint32_t acc_full = bar_base + 8u * 16;
  uint32_t tiles0 = smem_u32(smem) + 1024u;
  tiles0 += (1024u - (tiles0 & 1023u)) & 1023u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int combos = p.K * p.n_slab_in;
  const int q0 = 2 * blockIdx.x;                                  // this CTA's combos q0, q0 + 1
  const int c_lo = blockIdx.z * 128;                              // output-channel chunk
  const int cn = min(128, p.cout - c_lo);
  const int nb = cn >> 5;                                         // B tiles per stage
  const int n_tiles = 2 + nb;
  const uint32_t stage_bytes = (uint32_t)n_tiles * WG_TILE;
  const long long r_lo = (long long)blockIdx.y * p.rows_per_split;
  const long long r_hi = min(p.n_out, r_lo + p.rows_per_split);
  const int n_stage = r_hi > r_lo ? (int)((r_hi - r_lo + WG_ROWS - 1) / WG_ROWS) : 0;
  const uint32_t tmem_cols = 2u * cn <= 64u ? 64u : (2u * cn <= 128u ? 128u : 256u);

  if (tid == 0) {
    for (int s = 0; s < p.NS; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == WG_PROD_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < WG_PROD_WARPS) {
    // =========================================================================== producers
    // ptxas serialises a warp's TMA instructions (one per active lane, ~74 cycles each), so the 64 gather4s of the two
    // A tiles are spread over all eight warps (tile = warp & 1, WG_LANES row groups per warp), and every B tile -
    // WG_ROWS CONSECUTIVE rows - is one ordinary 2-D tile load (box 64 x WG_ROWS) issued by the next lane of warp b.
    if (n_stage > 0) {
      const int t = warp & 1;                                     // A tile / combo of this warp
      const int q = q0 + t;
      const bool q_ok = q < combos;
      const int k = q_ok ? q / p.n_slab_in : 0;
      const int col_a = q_ok ? (q % p.n_slab_in) * 64 : 0;
      const bool a_lane = lane < WG_LANES;
      const int g = (warp >> 1) * WG_LANES + lane;                // row group (rows 4g .. 4g+3 of the stage)
      const bool b_lane = lane == WG_LANES && warp < nb;
      const int col_b = ((c_lo >> 5) + warp) * 64;
      const int* nbr_k = p.nbr ? p.nbr + (long long)k * p.n_out : nullptr;
      auto rows_of = [&](int it) {
        const long long o = r_lo + (long long)it * WG_ROWS + 4 * g;
        int v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = -1;
          if (a_lane && q_ok && o + i < r_hi) v[i] = nbr_k ? __ldg(nbr_k + o + i) : (int)(o + i);
        }
        return make_int4(v[0], v[1], v[2], v[3]);
      };
      constexpr int AHEAD = 3;
      int4 ring[AHEAD];
#pragma unroll
      for (int d = 0; d < AHEAD; ++d) ring[d] = d < n_stage ? rows_of(d) : make_int4(-1, -1, -1, -1);
      for (int it = 0; it < n_stage;) {
#pragma unroll
        for (int d = 0; d < AHEAD; ++d) {
          if (it < n_stage) {
            const int4 r = ring[d];
            if (it + AHEAD < n_stage) ring[d] = rows_of(it + AHEAD);
            const int s = it % p.NS;
            mbar_wait(empty(s), (((uint32_t)(it / p.NS)) & 1u) ^ 1u);
            if (warp == 0 && lane == 0) mbar_arrive_expect_tx(full(s), stage_bytes);
            __syncwarp();
            const uint32_t st = tiles0 + (uint32_t)s * stage_bytes;
            if (a_lane) wg_gather4(st + (uint32_t)t * WG_TILE + (uint32_t)g * 512u, &tm_x, full(s), col_a, r.x, r.y, r.z, r.w);
            if (b_lane) {
              const int row = (int)(r_lo + (long long)it * WG_ROWS);
              asm volatile(
                  "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                      st + (uint32_t)(2 + warp) * WG_TILE),
                  "l"(&tm_dy), "r"(full(s)), "r"(col_b), "r"(row)
                  : "memory");
            }
            ++it;
          }
        }
      }
    }
    // =========================================================================== epilogue: warps 0..3 own the 128 rows
    if (warp < 4) {
      mbar_wait(acc_full, 0);
      tc_fence_after();
      const int m = warp * 32 + lane;                             // D row = TMEM lane
      const int q = q0 + (m >> 6), hl = (m >> 5) & 1, ci = m & 31;
      if (q < combos) {
        const int k = q / p.n_slab_in, slab = q % p.n_slab_in;
        float* dst = p.partial + (((size_t)k * (2 * p.splits) + (size_t)(2 * blockIdx.y + hl)) * p.cin + (size_t)(slab * 32 + ci)) * p.cout + c_lo;
        const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int b = 0; b < nb; ++b) {
          float v[64];
          if (n_stage > 0) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) tmem_ld16(t_lane + (uint32_t)(b * 64 + ch * 16), v + ch * 16);
          } else {
#pragma unroll
            for (int e = 0; e < 64; ++e) v[e] = 0.f;
          }
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4)
            *reinterpret_cast<float4*>(dst + b * 32 + e4 * 4) =
                make_float4(v[e4 * 4] + v[32 + e4 * 4], v[e4 * 4 + 1] + v[32 + e4 * 4 + 1], v[e4 * 4 + 2] + v[32 + e4 * 4 + 2],
                            v[e4 * 4 + 3] + v[32 + e4 * 4 + 3]);
        }
      }
      tc_fence_before();
    }
  } else {
    // =========================================================================== MMA issuer
    const uint32_t idesc = umma_idesc_bf16_major(64 * nb, 1, 1);  // M 128, N = 64 nb, A and B MN-major
    // MN-major SWIZZLE_128B: LBO = bytes between 64-element atoms along M/N (one tile), SBO = between 8-row groups
    const uint32_t hi32 = umma_desc_hi32(1024) | (2u << 29);
    for (int it = 0; it < n_stage; ++it) {
      const int s = it % p.NS;
      mbar_wait(full(s), ((uint32_t)(it / p.NS)) & 1u);
      tc_fence_after();
      const uint32_t a0 = tiles0 + (uint32_t)s * stage_bytes, b0 = a0 + 2u * WG_TILE;
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < WG_ROWS / 16; ++j) {                   // 16 rows per step = two 8-row groups = 2048 B
          const uint64_t da = umma_desc_join(hi32, umma_desc_lo32(a0 + j * 2048u, WG_TILE));
          const uint64_t db = umma_desc_join(hi32, umma_desc_lo32(b0 + j * 2048u, WG_TILE));
          umma_bf16(tmem_base, da, db, idesc, (it | j) ? 1u : 0u);
        }
        umma_commit(empty(s));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WG_PROD_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// fp32 rows -> split rows (64 B bf16 hi | 64 B bf16 lo per 32-channel slab); one thread per 8 channels
__global__ void pack_split_kernel(const float* __restrict__ in, int in_ld, int C, long long n, float* __restrict__ out,
                                  int out_ld) {
  const int c8n = C >> 3;
  const long long total = n * c8n;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / c8n;
    const int c0 = (int)(t % c8n) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(in + row * in_ld + c0));
    const float4 b = __ldg(reinterpret_cast<const float4*>(in + row * in_ld + c0 + 4));
    uint32_t h[4], l[4];
    split2(a.x, a.y, h[0], l[0]);
    split2(a.z, a.w, h[1], l[1]);
    split2(b.x, b.y, h[2], l[2]);
    split2(b.z, b.w, h[3], l[3]);
    unsigned char* d = reinterpret_cast<unsigned char*>(out + row * out_ld) + (size_t)(c0 >> 5) * 128 + (size_t)((c0 >> 3) & 3) * 16;
    *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(d + 64) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

typedef CUresult (*WgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WgEncodeFn wg_encode() {
  static WgEncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WgEncodeFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}
static bool wg_row_map(CUtensorMap* tm, const float* base, int ld, int channels, long long rows, int box_rows) {
  WgEncodeFn enc = wg_encode();
  if (!enc || rows <= 0) return false;
  cuuint64_t dims[2] = {(cuuint64_t)channels * 2, (cuuint64_t)rows};       // true extents (see spconv_tc.cu)
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct WgPlan { int groups, chunks, splits; long long rows_per_split; };
static WgPlan wg_plan(long long n_out, int K, int cin, int cout) {
  WgPlan pl;
  const int combos = K * (cin / 32);
  pl.groups = (combos + 1) / 2;
  pl.chunks = (cout + 127) / 128;
  const long long base = (long long)pl.groups * pl.chunks;
  const long long tiles = (n_out + WG_ROWS - 1) / WG_ROWS;
  long long splits = ((long long)sm_count() * 3 + base - 1) / base;        // ~3 waves of CTAs
  splits = std::max(1LL, std::min(splits, (tiles + 3) / 4));               // at least 4 row tiles per CTA
  long long per = (tiles + splits - 1) / splits;
  pl.rows_per_split = per * WG_ROWS;
  pl.splits = (int)((n_out + pl.rows_per_split - 1) / pl.rows_per_split);
  return pl;
}

}  // namespace ag3d

using namespace ag3d;

extern "C" {

int ag3d_pack_split(const float* in, int32_t in_ld, int32_t C, int64_t n, float* out, int32_t out_ld,
                    ag3d_stream_t stream) {
  AG3D_CHECK_ARG(in && out && aligned16(in) && aligned16(out), "pack_split: pointers");
  AG3D_CHECK_ARG(n > 0 && C >= 32 && C % 32 == 0 && in_ld % 4 == 0 && out_ld % 4 == 0 && in_ld >= C && out_ld >= C,
                 "pack_split: shape");
  const long long total = n * (C / 8);
  long long blocks = (total + 255) / 256;
  blocks = std::min<long long>(blocks, (long long)sm_count() * 16);
  pack_split_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(in, in_ld, C, n, out, out_ld);
  AG3D_LAUNCH_CHECK("pack_split");
  return AG3D_OK;
}

int32_t ag3d_spconv_bwd_weight_tc_supported(int32_t K, int32_t cin, int32_t cout) {
  return (K >= 1 && cin >= 32 && cin % 32 == 0 && cout >= 32 && cout % 32 == 0 && wg_encode() != nullptr) ? 1 : 0;
}

size_t ag3d_spconv_bwd_weight_tc_workspace_bytes(int64_t n_out, int32_t K, int32_t cin, int32_t cout) {
  if (n_out <= 0 || K < 1 || cin < 32 || cout < 32) return 0;
  const WgPlan pl = wg_plan(n_out, K, cin, cout);
  return (size_t)2 * pl.splits * K * cin * cout * sizeof(float);
}

int ag3d_spconv_bwd_weight_tc(const float* in_split, int64_t n_in, int32_t in_ld, int32_t cin, const int32_t* nbr,
                              int32_t K, int64_t n_out, const float* dout_split, int32_t dout_ld, int32_t cout,
                              float* dweight, int32_t accumulate, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_out > 0 && n_out < 2147483647LL && n_in > 0 && n_in < 2147483647LL, "bwd_weight_tc: row counts");
  AG3D_CHECK_ARG(K >= 1 && (nbr || (K == 1 && n_in == n_out)), "bwd_weight_tc: K > 1 needs a neighbour table");
  AG3D_CHECK_ARG(cin >= 32 && cin % 32 == 0 && cout >= 32 && cout % 32 == 0, "bwd_weight_tc: channels must be multiples of 32");
  AG3D_CHECK_ARG(in_split && dout_split && dweight && aligned16(in_split) && aligned16(dout_split) && aligned16(dweight),
                 "bwd_weight_tc: pointers");
  AG3D_CHECK_ARG(in_ld % 4 == 0 && dout_ld % 4 == 0 && in_ld >= cin && dout_ld >= cout, "bwd_weight_tc: leading dims");
  const WgPlan pl = wg_plan(n_out, K, cin, cout);
  AG3D_CHECK_ARG(ws && aligned16(ws) && ws_bytes >= (size_t)2 * pl.splits * K * cin * cout * sizeof(float),
                 "bwd_weight_tc: workspace too small (ag3d_spconv_bwd_weight_tc_workspace_bytes)");
  alignas(64) CUtensorMap tm_x, tm_dy;
  AG3D_CHECK_ARG(wg_row_map(&tm_x, in_split, in_ld, cin, n_in, 1) && wg_row_map(&tm_dy, dout_split, dout_ld, cout, n_out, WG_ROWS),
                 "bwd_weight_tc: cuTensorMapEncodeTiled failed");
  WgParams p;
  p.nbr = nbr; p.K = K; p.n_out = n_out; p.cin = cin; p.cout = cout; p.n_slab_in = cin / 32;
  p.rows_per_split = pl.rows_per_split; p.splits = pl.splits;
  p.partial = static_cast<float*>(ws);
  const int nb_max = std::min(cout, 128) / 32;
  const size_t stage = (size_t)(2 + nb_max) * WG_TILE;
  int ns = (int)((size_t)(220 * 1024 - 2048) / stage);
  p.NS = std::max(1, std::min(ns, 8));
  const size_t smem = 2048 + (size_t)p.NS * stage;
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  cudaStream_t st = as_stream(stream);
  wgrad_tc_kernel<<<dim3(pl.groups, pl.splits, pl.chunks), WG_THREADS, smem, st>>>(tm_x, tm_dy, p);
  AG3D_LAUNCH_CHECK("wgrad_tc");
  const long long count = (long long)cin * cout;
  return launch_split_reduce(p.partial, 2 * pl.splits, count, (long long)K * count, accumulate, dweight, st);
}

}  // extern "C"
