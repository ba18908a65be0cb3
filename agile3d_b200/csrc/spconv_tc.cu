// K4 (tensor-core variant): output-stationary sparse convolution as an implicit GEMM on the 5th-gen tensor cores.
//
//   out[o, :] = act( scale * sum_k in[nbr[k][o], :] @ W[k] + shift (+ residual[o, :]) )
//
// Structure (352 threads: 8 producer/epilogue warps, MMA issuer, weight loader, optional second issuer):
//   * the CTA owns T consecutive 128-row output tiles; their fp32 accumulators [128 x Cout] live in TMEM for the whole
//     kernel (two CTAs per SM with T <= 2 in the split-row modes, one CTA with T <= 4 on fp32 rows);
//   * the weight operand is *stationary*: for every (kernel offset k, 32-channel slab c) the pre-split weight
//     slab is brought into shared memory ONCE by the TMA engine (cp.async.bulk) and reused by all T tiles;
//   * the A operand is the 32-channel slab of the 128 neighbour rows of (tile, k), one ring stage per (k, slab, tile):
//       MODE 2 (default for split rows): the TMA engine gathers it - tile::gather4 pulls four rows per instruction
//               into a SWIZZLE_128B tile, absent neighbours (-1) are out-of-range rows = zeros, completion through the
//               stage's mbarrier; the producer warps only issue (slot = warp % NA, 32 / (8/NA) gathers per warp);
//       MODE 1: every thread cp.async-copies its 16-byte pieces of the split rows (row count of the input unknown);
//       MODE 0: fp32 rows through registers, split into bf16 hi/lo there, canonical no-swizzle K-major tiles;
//   * one elected thread issues tcgen05.mma (kind::f16, bf16 inputs, fp32 accumulate): per 16-channel step the three
//     products hi*hi + hi*lo + lo*hi ("bf16x3", relative error <= ~1e-5) accumulate into TMEM;
//   * tcgen05.commit hands shared-memory stages back to the producers and finally signals the epilogue;
//   * the 8 producer warps then become the epilogue: tcgen05.ld the accumulators, apply folded BatchNorm /
//     bias, residual, ReLU and write the channel slice of the output buffer (fp32 or split rows).
// (tile, k) pairs in which no row of the tile has a neighbour are skipped by all roles; levels with few rows split
// the kernel offsets over gridDim.y (deterministic reduction in splitk_reduce_kernel).
// AG3D_TC_* environment switches are measurement aids (tools/tc_probe.sh), not configuration.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace ag3d {

constexpr int TC_BK = 32;                        // input channels per pipeline stage
constexpr int TC_PROD_WARPS = 8;
constexpr int TC_PROD_THREADS = TC_PROD_WARPS * 32;
constexpr int TC_THREADS = TC_PROD_THREADS + 96; // + MMA warp 0 + weight-loader warp + MMA warp 1
constexpr int TC_MAX_T = 4;
constexpr int TC_BAR_BYTES = 512;
constexpr int TC_MAX_STAGES = 32 * 12 * TC_MAX_T;   // K <= 32 offsets, cin <= 384 (12 slabs), T tiles
constexpr int TC_LIST_BYTES = TC_MAX_STAGES * 4;

// ---------------------------------------------------------------------------------------------- weight prep
// W [K][cin][cout] fp32  ->  Wp [K][cin/32][piece 2][kc 4][cout][8] bf16 : the exact shared-memory image of every
// (k, slab) weight stage (UMMA K-major canonical layout with SBO = 128, LBO = cout*16), so one 1-D bulk copy
// loads a stage.
__global__ void weight_prep_kernel(const float* __restrict__ w, int K, int cin, int cout, uint4* __restrict__ wp) {
  const long long total = (long long)K * (cin / TC_BK) * 4 * cout;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(t % cout);
    long long r = t / cout;
    const int kc = (int)(r % 4); r /= 4;
    const int slab = (int)(r % (cin / TC_BK));
    const int k = (int)(r / (cin / TC_BK));
    const float* src = w + ((long long)k * cin + slab * TC_BK + kc * 8) * cout + n;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2(__ldg(src + (2 * e) * (long long)cout), __ldg(src + (2 * e + 1) * (long long)cout), hi[e], lo[e]);
    const long long stage = ((long long)k * (cin / TC_BK) + slab) * (2 * 4 * (long long)cout);
    wp[stage + (0 * 4 + kc) * (long long)cout + n] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    wp[stage + (1 * 4 + kc) * (long long)cout + n] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---------------------------------------------------------------------------------------------- split format
// "split" feature rows: every 32-channel slab (128 B) is stored as 64 B of bf16 hi (32 channels) followed by 64 B of
// bf16 lo (x = hi + lo + O(2^-17 |x|)).  A row of C channels occupies exactly the 4*C bytes of the fp32 row, so
// leading dimensions, 32-channel-aligned slices and concat buffers are unchanged, and a gathered slab is 128
// contiguous bytes whose 16-byte pieces are exactly the UMMA core-matrix rows (4 hi pieces, then 4 lo pieces).
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// byte offset of the hi half of channels c0 .. c0+15 (c0 % 16 == 0) inside a split row; the lo half is 64 B further
__device__ __forceinline__ size_t split_off16(int c0) { return (size_t)(c0 >> 5) * 128 + (size_t)((c0 >> 4) & 1) * 32; }

__device__ __forceinline__ void load_split16(const float* row, int c0, float* v) {
  const uint4* h = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(row) + split_off16(c0));
  const uint4 hh[2] = {__ldg(h), __ldg(h + 1)}, ll[2] = {__ldg(h + 4), __ldg(h + 5)};
  const uint32_t* hp = reinterpret_cast<const uint32_t*>(hh);
  const uint32_t* lp = reinterpret_cast<const uint32_t*>(ll);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    v[2 * e] = __uint_as_float(hp[e] << 16) + __uint_as_float(lp[e] << 16);
    v[2 * e + 1] = __uint_as_float(hp[e] & 0xFFFF0000u) + __uint_as_float(lp[e] & 0xFFFF0000u);
  }
}
__device__ __forceinline__ void store_split16(float* row, int c0, const float* v) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) split2(v[2 * e], v[2 * e + 1], h[e], l[e]);
  uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(row) + split_off16(c0));
  d[0] = make_uint4(h[0], h[1], h[2], h[3]);
  d[1] = make_uint4(h[4], h[5], h[6], h[7]);
  d[4] = make_uint4(l[0], l[1], l[2], l[3]);
  d[5] = make_uint4(l[4], l[5], l[6], l[7]);
}

// ---------------------------------------------------------------------------------------------- main kernel
struct TcParams {
  const float* in; int in_ld; int cin;
  const int* nbr; int K; long long n_out;
  const uint4* wp; int cout;
  const float* scale; const float* shift; const float* residual; int res_ld;
  float* out; int out_ld; int flags;
  int in_split, out_split, res_split;   // feature format: 0 = fp32, 1 = bf16 hi/lo pairs (AG3D_*_SPLIT)
  int T;            // tiles per CTA
  int NA;           // A ring stages
  int NB;           // weight stages
  int na_log2, nb_log2;
  int k_per;        // kernel offsets per CTA row (split-K over gridDim.y); k range = [by*k_per, min(K, (by+1)*k_per))
  float* partial;   // split-K: raw accumulators [gridDim.y][n_out][cout]; NULL = fused epilogue
  int NI;           // MMA-issuing warps (1 or 2): issuer w owns the stages of tiles j with (j & 1) == w
  int debug;        // profiling experiments only (AG3D_TC_DEBUG): 1 = skip MMAs, 2 = skip gather loads (MODE 0/1), 4 = one product
  int tma_out;      // MODE 2, split output: the tile is staged in the (idle) operand ring and stored by the TMA engine
  int cpad;         // TMEM columns per tile (pow2 >= cout)
  int tmem_cols;    // allocation (pow2, 32..512)
};

constexpr uint32_t TMA_STAGE = 16384;   // [128 rows x 128 B] gathered slab, SWIZZLE_128B, 1024-byte aligned

// four rows of a 2-D tensor (row = 128 B here) -> four consecutive 128-byte lines of shared memory; a negative or
// out-of-range row index yields zeros and still counts its bytes on the mbarrier (measured, profiles/r01_e_tma_*).
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int col, int r0, int r1,
                                            int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

// MODE 0: fp32 input rows, gathered through registers and split in the kernel, one CTA per SM
// MODE 1: input rows are bf16 hi/lo pairs ("split"): cp.async gather, few registers, two CTAs per SM
// MODE 2: split rows gathered by the TMA engine (tile::gather4, SWIZZLE_128B operand tiles): one producer WARP per
//         ring slot issues a whole stage with a single instruction (lane l = rows 4l..4l+3)
template <int MODE>
__global__ void __launch_bounds__(TC_THREADS, MODE ? 2 : 1)
spconv_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_out, const TcParams p) {
  constexpr bool SPLIT = MODE != 0;
  constexpr bool TMA = MODE == 2;
  extern __shared__ __align__(1024) unsigned char smem[];
  // barrier block: bars[0..7] a_full, [8..15] a_empty, [16..19] b_full, [20..23] b_empty, [24] acc_full
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  uint32_t* kmask_s = reinterpret_cast<uint32_t*>(smem + 272);   // [TC_MAX_T]
  int* n_stage_s = reinterpret_cast<int*>(smem + 288);
  uint32_t* stage_list = reinterpret_cast<uint32_t*>(smem + TC_BAR_BYTES);   // [TC_MAX_STAGES] (k << 16 | slab << 8 | tile)
  const uint32_t b_stage_bytes = (uint32_t)p.cout * 128u;
  int* nbr_s = reinterpret_cast<int*>(smem + TC_BAR_BYTES + TC_LIST_BYTES);   // [k_per][T*128] this CTA's slice of the map
  const int TR = p.T * TC_BM;
  unsigned char* b_smem = smem + TC_BAR_BYTES + TC_LIST_BYTES + (SPLIT ? (size_t)0 : (size_t)p.k_per * TR * 4);
  unsigned char* a_smem = b_smem + (size_t)p.NB * b_stage_bytes;
  if constexpr (TMA) a_smem += (1024u - (smem_u32(a_smem) & 1023u)) & 1023u;   // swizzle atoms are 1024-byte aligned

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Programmatic dependent launch: let the next kernel of the stream start its prologue as soon as SMs free up
  // (its own griddepcontrol.wait keeps it from touching anything this grid reads or writes).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long tiles_total = (p.n_out + TC_BM - 1) / TC_BM;
  const long long tile0 = (long long)blockIdx.x * p.T;
  const int T_here = (int)min((long long)p.T, tiles_total - tile0);
  const long long row0 = tile0 * TC_BM;
  const int n_slab = p.cin / TC_BK;
  const int k_lo = blockIdx.y * p.k_per;
  const int k_hi = min(p.K, k_lo + p.k_per);
  const int na_mask = p.NA - 1, na_shift = p.na_log2;       // ring sizes are powers of two
  const int nb_mask = p.NB - 1, nb_shift = p.nb_log2;

  const uint32_t bar_base = smem_u32(bars);
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (8 + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (16 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (20 + s); };
  const uint32_t acc_full = bar_base + 8u * 24;

  if (tid == 0) {
    for (int s = 0; s < p.NA; ++s) { mbar_init(a_full(s), TMA ? 1 : (SPLIT ? ((p.debug & 256) ? 4 : 128) : TC_PROD_WARPS / 2)); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < p.NB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), p.NI); }
    mbar_init(acc_full, p.NI);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < TC_MAX_T) kmask_s[tid] = 0;
  if (warp == TC_PROD_WARPS) {   // MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- which kernel offsets does each tile need?  (tile, k) pairs without any neighbour are skipped by all roles.
  // 256 producer threads: thread -> row (tid & 127) of tiles (tid >> 7) and (tid >> 7) + 2; all loads issued first.
  if (tid < TC_PROD_THREADS) {
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
      const int j = (tid >> 7) + 2 * jj;
      const long long row = row0 + (long long)j * TC_BM + (tid & 127);
      uint32_t mine = 0;
      if (j < p.T) {
        int* dst = nbr_s + j * TC_BM + (tid & 127);                 // column of this row in the staged slice
        if (j < T_here && row < p.n_out) {
          if (p.nbr) {
            const int* col = p.nbr + row;
#pragma unroll 9
            for (int k = k_lo; k < k_hi; ++k) {
              const int v = __ldg(col + (long long)k * p.n_out);
              if constexpr (!SPLIT) dst[(k - k_lo) * TR] = v;
              mine |= (v >= 0 ? 1u : 0u) << k;
            }
          } else {
            if constexpr (!SPLIT) dst[0] = (int)row;                // identity map: K == 1, never split
            mine = 1u;
          }
        } else if constexpr (!SPLIT) {
          for (int k = k_lo; k < k_hi; ++k) dst[(k - k_lo) * TR] = -1;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mine |= __shfl_xor_sync(0xffffffffu, mine, o);
      if (lane == 0 && mine && j < TC_MAX_T) atomicOr(&kmask_s[j], mine);
    }
  }
  __syncthreads();
  // ---- stage list in (k, slab, tile) order: lane k of warp 0 emits the stages of offset k
  if (warp == 0) {
    const int k = k_lo + lane;
    uint32_t tiles_k = 0;                                       // bit j: tile j needs offset k
    if (k < k_hi)
      for (int j = 0; j < T_here; ++j) tiles_k |= ((kmask_s[j] >> k) & 1u) << j;
    const int cnt = n_slab * __popc(tiles_k);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    int pos = incl - cnt;
    for (int c = 0; c < n_slab; ++c)
      for (int j = 0; j < T_here; ++j)
        if ((tiles_k >> j) & 1u) stage_list[pos++] = ((uint32_t)k << 16) | ((uint32_t)c << 8) | (uint32_t)j;
    if (lane == 31) *n_stage_s = incl;
  }
  __syncthreads();
  const int n_stage = *n_stage_s;
  const uint32_t tmem_base = *tmem_slot;
  // Everything above read only the neighbour table (built before the previous kernels of the stream were launched),
  // shared and tensor memory.  From here on the grid reads the previous layer's output and writes global memory:
  // wait for the preceding grid to complete and flush (no-op when this grid was launched without the PDL attribute).
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp < TC_PROD_WARPS) {
    // =========================================================================== A producers
    // Two producer groups of 4 warps work on alternating stages (group = warp >> 2 takes stages group, group+2, ...)
    // so that two stages are always being converted/stored concurrently: the per-stage chain (barrier wait -> split
    // -> st.shared -> proxy fence -> arrive) is latency bound when every warp has to touch every stage.
    // thread -> 8 channels (one 16-byte bf16 K-chunk kc) of rows rbase + 32*i, i = 0..3
    constexpr int PG = 2;                                   // producer groups
    constexpr int PR = 4;                                   // rows per thread per stage
    const int grp = warp >> 2;
    const int tg = tid & 127;
    const int kc = tg & 3;
    const int rbase = tg >> 2;                              // 0..31
    uint32_t st_off[PR];
#pragma unroll
    for (int i = 0; i < PR; ++i) {
      const int r = rbase + 32 * i;
      st_off[i] = (uint32_t)(kc * A_LBO + (r >> 3) * 128 + (r & 7) * 16);
    }
    const float* in_kc = p.in + kc * 8;

    if constexpr (TMA) {
      // ---- TMA gather.  The eight producer warps share the NA ring slots: slot = warp % NA, and the 8 / NA warps of
      //      a slot each issue their share of the stage's 32 gather4 instructions (row group g = rows 4g..4g+3).
      //      ptxas serialises a warp's gather4s (one instruction per active lane), so spreading a stage over more
      //      warps raises the issue rate (measured: profiles/r01_e_tma_gather4_bandwidth.txt).  Per stage: the lane's
      //      four neighbour rows (fetched IDX_AHEAD own stages ahead), wait for the slot, one expect_tx, the gathers.
      {
        constexpr int IDX_AHEAD = 4;
        const int slot_id = warp & na_mask, part = warp >> na_shift;
        const int lanes_per = (32 * p.NA) / TC_PROD_WARPS;        // active lanes per warp: 8 (NA = 2), 16, 32
        const bool active = lane < lanes_per;
        const int g = part * lanes_per + lane;                    // row group of this lane
        const uint32_t slot = smem_u32(a_smem) + (uint32_t)slot_id * TMA_STAGE + (uint32_t)g * 512u;
        const uint32_t full = a_full(slot_id), empty = a_empty(slot_id);
        const bool leader = part == 0 && lane == 0;
        const long long rows_here = p.n_out - row0;            // rows of this CTA that exist
        int4 ring[IDX_AHEAD];
        uint32_t cs[IDX_AHEAD];
        auto fetch = [&](int n, int4& r, uint32_t& cslab) {
          const uint32_t e = stage_list[n];
          const int k = (int)(e >> 16), j = (int)(e & 0xFFu);
          cslab = (e >> 8) & 0xFFu;
          const int off = j * TC_BM + 4 * g;
          int v[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            v[i] = -1;
            if (active && off + i < rows_here)
              v[i] = p.nbr ? __ldg(p.nbr + (long long)k * p.n_out + row0 + off + i) : (int)(row0 + off + i);
          }
          r = make_int4(v[0], v[1], v[2], v[3]);
        };
#pragma unroll
        for (int d = 0; d < IDX_AHEAD; ++d) {
          ring[d] = make_int4(-1, -1, -1, -1);
          cs[d] = 0;
          if (slot_id + d * p.NA < n_stage) fetch(slot_id + d * p.NA, ring[d], cs[d]);
        }
        int it = 0;                                            // stages this warp has issued (slot phase)
        for (int n = slot_id; n < n_stage;) {
#pragma unroll
          for (int d = 0; d < IDX_AHEAD; ++d) {
            if (n < n_stage) {
              const int4 r = ring[d];
              const int col = (int)cs[d] * 64;
              if (n + IDX_AHEAD * p.NA < n_stage) fetch(n + IDX_AHEAD * p.NA, ring[d], cs[d]);
              mbar_wait(empty, ((uint32_t)it & 1u) ^ 1u);
              if (leader) mbar_arrive_expect_tx(full, TMA_STAGE);
              __syncwarp();
              if (active) tma_gather4(slot, &tm_in, full, col, r.x, r.y, r.z, r.w);
              n += p.NA;
              ++it;
            }
          }
        }
      }
    } else if constexpr (SPLIT) {
      // ---- input already stored as bf16 hi/lo pairs: the gather is a pure byte copy, done by cp.async straight
      //      into the operand stage (no registers, no ALU); completion is tracked by the stage's mbarrier, so up to
      //      NA stages of gathers are in flight per group.
      const unsigned char* in_b = reinterpret_cast<const unsigned char*>(p.in) + kc * 16;   // hi piece kc; lo at +64
      const size_t row_bytes = (size_t)p.in_ld * 4;
      const long long rows_left = p.n_out - row0 - rbase;   // row (j, i) is real iff j*128 + 32*i < rows_left
      const int* nbr_r = p.nbr ? p.nbr + row0 + rbase : nullptr;
      // The neighbour indices of a stage come from global memory (L2 / HBM latency); they are fetched IDX_AHEAD own
      // stages (= 2 * IDX_AHEAD stages of the CTA) before they are used, in a register ring, so that this latency
      // never sits on the stage loop (with one stage of look-ahead the loop ran at one index round trip per stage).
      constexpr int IDX_AHEAD = 4;
      int idx[IDX_AHEAD][PR];
      uint32_t cs[IDX_AHEAD];
      auto fetch_idx = [&](int n, int (&ix)[PR], uint32_t& cslab) {
        const uint32_t e = stage_list[n];
        const int k = (int)(e >> 16), j = (int)(e & 0xFFu);
        cslab = (e >> 8) & 0xFFu;
#pragma unroll
        for (int i = 0; i < PR; ++i) {
          const int off = j * TC_BM + 32 * i;
          ix[i] = -1;
          if (off < rows_left) ix[i] = nbr_r ? __ldg(nbr_r + (long long)k * p.n_out + off) : (int)(row0 + rbase + off);
        }
      };
#pragma unroll
      for (int d = 0; d < IDX_AHEAD; ++d) {
        cs[d] = 0;
#pragma unroll
        for (int i = 0; i < PR; ++i) idx[d][i] = -1;
        if (grp + d * PG < n_stage) fetch_idx(grp + d * PG, idx[d], cs[d]);
      }
      for (int n = grp; n < n_stage;) {
#pragma unroll
        for (int d = 0; d < IDX_AHEAD; ++d) {
          if (n < n_stage) {
            const int s = n & na_mask;
            mbar_wait(a_empty(s), (((uint32_t)n >> na_shift) & 1u) ^ 1u);
            const uint32_t st = smem_u32(a_smem + (size_t)s * A_STAGE);
            const unsigned char* src = in_b + (size_t)cs[d] * 128;
#pragma unroll
            for (int i = 0; i < PR; ++i) {
              const bool ok = idx[d][i] >= 0 && !(p.debug & 2);
              const unsigned char* g = ok ? src + (size_t)idx[d][i] * row_bytes : src;
              cp_async16_zfill(st + st_off[i], g, ok ? 16u : 0u);
              cp_async16_zfill(st + A_PIECE + st_off[i], g + 64, ok ? 16u : 0u);
            }
            if (!(p.debug & 256) || lane == 0) cp_async_mbar_arrive_noinc(a_full(s));
            if (n + IDX_AHEAD * PG < n_stage) fetch_idx(n + IDX_AHEAD * PG, idx[d], cs[d]);
            n += PG;
          }
        }
      }
    } else {
      int idx_ld[PR];                  // neighbour rows of stage n_issued (prefetched one own-stage ahead)
      uint32_t c_ld = 0;
      int n_issued = grp;
      auto load_idx = [&]() {          // decode stage n_issued and fetch its neighbour rows
        const uint32_t e = stage_list[n_issued];
        const int k = (int)(e >> 16), j = (int)(e & 0xFFu);
        c_ld = (e >> 8) & 0xFFu;
        const int* col = nbr_s + (k - k_lo) * TR + j * TC_BM + rbase;
#pragma unroll
        for (int i = 0; i < PR; ++i) idx_ld[i] = col[32 * i];
      };
      auto issue = [&](float4 (&buf)[2 * PR]) -> bool {
        if (n_issued >= n_stage) return false;
        const float* src = in_kc + c_ld * TC_BK;
  #pragma unroll
        for (int i = 0; i < PR; ++i) {
          buf[2 * i] = buf[2 * i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (idx_ld[i] >= 0 && !(p.debug & 2)) {
            const float4* r = reinterpret_cast<const float4*>(src + (size_t)idx_ld[i] * (size_t)p.in_ld);
            buf[2 * i] = __ldg(r);
            buf[2 * i + 1] = __ldg(r + 1);
          }
        }
        n_issued += PG;
        if (n_issued < n_stage) load_idx();
        return true;
      };
      int n_done = grp;                // next stage this group finishes (ring position)
      auto finish = [&](const float4 (&buf)[2 * PR]) {
        const int s = n_done & na_mask;
        if (!(p.debug & 16)) mbar_wait(a_empty(s), (((uint32_t)n_done >> na_shift) & 1u) ^ 1u);
        unsigned char* st = a_smem + (size_t)s * A_STAGE;
  #pragma unroll
        for (int i = 0; i < PR; ++i) {
          uint32_t h[4], l[4];
          split2(buf[2 * i].x, buf[2 * i].y, h[0], l[0]);
          split2(buf[2 * i].z, buf[2 * i].w, h[1], l[1]);
          split2(buf[2 * i + 1].x, buf[2 * i + 1].y, h[2], l[2]);
          split2(buf[2 * i + 1].z, buf[2 * i + 1].w, h[3], l[3]);
          *reinterpret_cast<uint4*>(st + st_off[i]) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(st + A_PIECE + st_off[i]) = make_uint4(l[0], l[1], l[2], l[3]);
        }
        if (!(p.debug & 8)) fence_proxy_async();   // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full(s));
        n_done += PG;
      };
      if (n_issued < n_stage) load_idx();
      // three own stages of gathers in flight per thread (24 x 16 B): the gather is latency-bound otherwise
      float4 b0[2 * PR], b1[2 * PR], b2[2 * PR];
      bool v0 = issue(b0), v1 = issue(b1), v2 = issue(b2);
      while (v0) {
        finish(b0);
        v0 = issue(b0);
        if (!v1) break;
        finish(b1);
        v1 = issue(b1);
        if (!v2) break;
        finish(b2);
        v2 = issue(b2);
      }
    }

    // =========================================================================== epilogue
    // 8 warps: TMEM lane quarter q = warp & 3, column half = warp >> 2
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int ncol = p.cout >> 1;    // columns per half (multiple of 16)
    const bool relu = p.flags & AG3D_RELU;
    for (int j = 0; j < ((p.debug & 64) ? 0 : T_here); ++j) {
      const long long row = row0 + (long long)j * TC_BM + q * 32 + lane;
      const bool live = kmask_s[j] != 0u;
      for (int c0 = half * ncol; c0 < (half + 1) * ncol; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * p.cpad + c0), v);
        if (row < p.n_out && p.partial) {       // split-K: raw partial sums, reduced + finished by splitk_reduce_kernel
          float* dst = p.partial + ((size_t)blockIdx.y * (size_t)p.n_out + (size_t)row) * p.cout + c0;
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4)
            *reinterpret_cast<float4*>(dst + e4 * 4) =
                live ? make_float4(v[e4 * 4], v[e4 * 4 + 1], v[e4 * 4 + 2], v[e4 * 4 + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
        } else if (row < p.n_out) {
          float4 r4[4];
          if (p.residual) {
            if (p.res_split) {
              load_split16(p.residual + row * p.res_ld, c0, reinterpret_cast<float*>(r4));
            } else {
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4)
                r4[e4] = __ldg(reinterpret_cast<const float4*>(p.residual + row * p.res_ld + c0 + e4 * 4));
            }
          }
          if (!live) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = 0.f;
          }
          if (p.scale) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + e4 * 4));
              v[e4 * 4 + 0] *= s4.x; v[e4 * 4 + 1] *= s4.y; v[e4 * 4 + 2] *= s4.z; v[e4 * 4 + 3] *= s4.w;
            }
          }
          if (p.shift) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 s4 = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + e4 * 4));
              v[e4 * 4 + 0] += s4.x; v[e4 * 4 + 1] += s4.y; v[e4 * 4 + 2] += s4.z; v[e4 * 4 + 3] += s4.w;
            }
          }
          if (p.residual) {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              v[e4 * 4 + 0] += r4[e4].x; v[e4 * 4 + 1] += r4[e4].y; v[e4 * 4 + 2] += r4[e4].z; v[e4 * 4 + 3] += r4[e4].w;
            }
          }
          if (relu) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
          }
          float4* dst = reinterpret_cast<float4*>(p.out + row * p.out_ld + c0);
          if (TMA && p.tma_out) {          // swizzled slab tile in the operand ring: [128 rows x (32 hi | 32 lo)] per slab
            uint32_t h[8], l[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) split2(v[2 * e], v[2 * e + 1], h[e], l[e]);
            const int rr = q * 32 + lane;
            unsigned char* slab = a_smem + (size_t)(c0 >> 5) * TMA_STAGE + rr * 128;
            const int k0 = (c0 >> 3) & 3;                                  // first of the two 8-channel chunks
            *reinterpret_cast<uint4*>(slab + (((k0) ^ (rr & 7)) << 4)) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(slab + (((k0 + 1) ^ (rr & 7)) << 4)) = make_uint4(h[4], h[5], h[6], h[7]);
            *reinterpret_cast<uint4*>(slab + (((4 + k0) ^ (rr & 7)) << 4)) = make_uint4(l[0], l[1], l[2], l[3]);
            *reinterpret_cast<uint4*>(slab + (((5 + k0) ^ (rr & 7)) << 4)) = make_uint4(l[4], l[5], l[6], l[7]);
          } else if (p.out_split) {
            store_split16(p.out + row * p.out_ld, c0, v);
          } else {
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4)
              dst[e4] = make_float4(v[e4 * 4], v[e4 * 4 + 1], v[e4 * 4 + 2], v[e4 * 4 + 3]);
          }
        }
      }
      if (TMA && p.tma_out) {
        // the tile sits in the operand ring in the layout of the output's tensor map: one thread hands it to the TMA engine
        // (full 128-byte row writes, rows past the end clipped) instead of every thread writing 16-byte pieces of its row
        fence_proxy_async();
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid == 0) {
          const int row_t = (int)(row0 + (long long)j * TC_BM);
          for (int sl = 0; sl < (p.cout >> 5); ++sl)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm_out),
                         "r"(smem_u32(a_smem) + (uint32_t)sl * TMA_STAGE), "r"(sl * 64), "r"(row_t)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
    if (TMA && p.tma_out && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else if (warp != TC_PROD_WARPS + 1) {
    // =========================================================================== MMA issuers
    // Issuing is the serial resource of this kernel (one instruction stream per issuer), so two warps share it:
    // issuer w owns the stages of the tiles j with (j & 1) == w (disjoint accumulators), both walk every weight
    // stage and both release it (b_empty counts NI arrivals).
    const int issuer = warp == TC_PROD_WARPS ? 0 : 1;
    if (issuer < p.NI) {
    // The whole warp walks the stage list with warp-uniform state; one elected lane issues the six MMAs of a stage
    // and the commit.  Descriptors of a ring slot are base + slot * stride in the low descriptor word.
    const uint32_t idesc = umma_idesc_bf16(p.cout);
    const uint32_t b_lbo = (uint32_t)p.cout * 16u;
    const uint32_t d_hi32 = umma_desc_hi32(128);
    // A operand: MODE 0/1 canonical no-swizzle tiles (A_STAGE bytes, hi piece | lo piece); MODE 2 one SWIZZLE_128B
    // tile per stage whose 128-byte rows are [hi ch 0-31 | lo ch 0-31]: k-step ks of the hi (lo) product starts
    // 32 ks (64 + 32 ks) bytes into the row, SBO = 1024 (eight rows), layout type 2.
    const uint32_t a_hi32 = TMA ? (umma_desc_hi32(1024) | (2u << 29)) : umma_desc_hi32(128);
    const uint32_t a_lo32 = TMA ? umma_desc_lo32(smem_u32(a_smem), 16) : umma_desc_lo32(smem_u32(a_smem), A_LBO);
    constexpr uint32_t A_SLOT16 = TMA ? (TMA_STAGE >> 4) : (uint32_t)(A_STAGE >> 4);
    constexpr uint32_t A_KS16 = TMA ? 2u : (uint32_t)((2 * A_LBO) >> 4);      // one 16-channel k-step
    constexpr uint32_t A_LO16 = TMA ? 4u : (uint32_t)(A_PIECE >> 4);          // hi piece -> lo piece
    const uint32_t b_lo32 = umma_desc_lo32(smem_u32(b_smem), b_lbo);
    const uint32_t b_lo_off = (4u * b_lbo) >> 4, b_ks_off = (2u * b_lbo) >> 4, b_slot = b_stage_bytes >> 4;
    uint32_t started = 0;            // bit j: accumulator j has been written
    int n_b = -1;                    // weight stage in use
    uint32_t prev_kc = 0xFFFFFFFFu;
    uint32_t b_cur = 0;
    for (int n = 0; n < n_stage; ++n) {
      const uint32_t e = stage_list[n];
      const int j = (int)(e & 0xFFu);
      if ((e >> 8) != prev_kc) {     // first tile of a new (k, slab): release the previous weight stage, take the next
        if (n_b >= 0 && elect_one()) umma_commit(b_empty(n_b & nb_mask));
        ++n_b;
        prev_kc = e >> 8;
        const int sb = n_b & nb_mask;
        mbar_wait(b_full(sb), ((uint32_t)n_b >> nb_shift) & 1u);
        b_cur = b_lo32 + (uint32_t)sb * b_slot;
      }
      if (p.NI == 2 && (j & 1) != issuer) continue;
      const int s = n & na_mask;
      mbar_wait(a_full(s), ((uint32_t)n >> na_shift) & 1u);
      if constexpr (SPLIT && !TMA) { if (!(p.debug & 128)) fence_proxy_async(); }   // cp.async (generic proxy) writes -> tensor core (async proxy) reads
      tc_fence_after();
      const uint32_t a_cur = a_lo32 + (uint32_t)s * A_SLOT16;
      const uint32_t d = tmem_base + (uint32_t)(j * p.cpad);
      const uint32_t acc0 = (started >> j) & 1u;
      started |= 1u << j;
      if (elect_one()) {
        if (!(p.debug & 1)) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {           // two 16-channel MMA steps per 32-channel slab
            const uint64_t da_hi = umma_desc_join(a_hi32, a_cur + ks * A_KS16);
            const uint64_t da_lo = umma_desc_join(a_hi32, a_cur + ks * A_KS16 + A_LO16);
            const uint64_t db_hi = umma_desc_join(d_hi32, b_cur + ks * b_ks_off);
            const uint64_t db_lo = umma_desc_join(d_hi32, b_cur + ks * b_ks_off + b_lo_off);
            umma_bf16(d, da_hi, db_hi, idesc, ks ? 1u : acc0);
            if (!(p.debug & 4)) {
              umma_bf16(d, da_hi, db_lo, idesc, 1u);
              umma_bf16(d, da_lo, db_hi, idesc, 1u);
            }
          }
        }
        umma_commit(a_empty(s));       // stage s may be overwritten once these MMAs have read it
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
    }
  } else {
    // =========================================================================== weight loader (TMA engine)
    int n_b = 0;
    uint32_t prev_kc = 0xFFFFFFFFu;
    for (int n = 0; n < n_stage; ++n) {
      const uint32_t e = stage_list[n];
      if ((e >> 8) == prev_kc) continue;
      prev_kc = e >> 8;
      const int k = (int)(e >> 16), c = (int)((e >> 8) & 0xFFu);
      const int sb = n_b & nb_mask;
      mbar_wait(b_empty(sb), (((uint32_t)n_b >> nb_shift) & 1u) ^ 1u);
      if (elect_one()) {
        if (p.debug & 32) {
          mbar_arrive(b_full(sb));
        } else {
          mbar_arrive_expect_tx(b_full(sb), b_stage_bytes);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wp) +
                                     ((size_t)k * n_slab + c) * (size_t)b_stage_bytes;
          bulk_g2s(smem_u32(b_smem + (size_t)sb * b_stage_bytes), src, b_stage_bytes, b_full(sb));
        }
      }
      __syncwarp();
      ++n_b;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == TC_PROD_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// out = act(scale * sum_y partial[y] + shift (+ residual)), y in fixed order (deterministic); 16 channels per thread
__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int ksplit, long long n_out, int cout,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ residual, int res_ld, float* __restrict__ out,
                                     int out_ld, int flags) {
  const int c16n = cout >> 4;
  const long long total = n_out * c16n;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / c16n;
    const int c = (int)(t % c16n) * 16;
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 0.f;
    for (int y = 0; y < ksplit; ++y) {
      const float4* src = reinterpret_cast<const float4*>(partial + ((size_t)y * n_out + row) * cout + c);
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        const float4 a = __ldg(src + e4);
        v[e4 * 4] += a.x; v[e4 * 4 + 1] += a.y; v[e4 * 4 + 2] += a.z; v[e4 * 4 + 3] += a.w;
      }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      if (scale) v[e] *= __ldg(scale + c + e);
      if (shift) v[e] += __ldg(shift + c + e);
    }
    if (residual) {
      float4 r4[4];
      if (flags & AG3D_RES_SPLIT) {
        load_split16(residual + row * res_ld, c, reinterpret_cast<float*>(r4));
      } else {
#pragma unroll
        for (int e4 = 0; e4 < 4; ++e4) r4[e4] = __ldg(reinterpret_cast<const float4*>(residual + row * res_ld + c) + e4);
      }
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        v[e4 * 4] += r4[e4].x; v[e4 * 4 + 1] += r4[e4].y; v[e4 * 4 + 2] += r4[e4].z; v[e4 * 4 + 3] += r4[e4].w;
      }
    }
    if (flags & AG3D_RELU) {
#pragma unroll
      for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
    }
    float4* dst = reinterpret_cast<float4*>(out + row * out_ld + c);
    if (flags & AG3D_OUT_SPLIT) {
      store_split16(out + row * out_ld, c, v);
    } else {
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) dst[e4] = make_float4(v[e4 * 4], v[e4 * 4 + 1], v[e4 * 4 + 2], v[e4 * 4 + 3]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- host side
static int pow2_at_least(int v, int lo) {
  int r = lo;
  while (r < v) r <<= 1;
  return r;
}

struct TcPlan { int cpad, T, ksplit, k_per; };

// tiles per CTA: minimise waves * (T gathers + one weight stage); weight stage cost relative to a gather = cout/128.
// Small levels (few tiles) are weight-streaming bound on a handful of SMs: split the kernel offsets over gridDim.y.
// CTAs per SM of the split-row variant (AG3D_TC_SPLIT_OCC = 1 | 2, default 2).  Two CTAs hide each other's stage
// round trips; one CTA with four weight stages shared by up to four tiles and an eight-deep gather ring was measured
// slower (0.46 vs 0.30 ms on the 150k-row 96->96 layer).
static int split_occ() {
  static int occ = -1;
  if (occ < 0) { const char* e = getenv("AG3D_TC_SPLIT_OCC"); occ = (e && atoi(e) == 1) ? 1 : 2; }
  return occ;
}

static TcPlan tc_plan(long long n_out, int K, int cout, bool split2 = false) {
  const bool split = split2;       // true: two co-resident CTAs per SM
  TcPlan pl;
  pl.cpad = pow2_at_least(cout, 32);
  // split mode runs two CTAs per SM: each may hold 256 TMEM columns
  const int t_max = std::max(1, std::min(split ? 2 : TC_MAX_T, (split ? 256 : 512) / pl.cpad));
  const long long tiles = (n_out + TC_BM - 1) / TC_BM;
  const int sms = sm_count();
  int best_t = 1;
  double best_cost = 1e30;
  for (int t = 1; t <= t_max; ++t) {
    const long long ctas = (tiles + t - 1) / t;
    const long long waves = (ctas + sms * (split ? 2 : 1) - 1) / (sms * (split ? 2 : 1));
    const double cost = (double)waves * ((double)t + (double)cout / 128.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best_t = t; }
  }
  {
    static int force_t = -1;         // tuning experiments only (AG3D_TC_T)
    if (force_t < 0) { const char* e = getenv("AG3D_TC_T"); force_t = e ? atoi(e) : 0; }
    if (force_t > 0) best_t = std::min(force_t, t_max);
  }
  pl.T = best_t;
  const long long ctas = (tiles + pl.T - 1) / pl.T;
  const long long slots = (long long)sms * (split ? 2 : 1);     // co-resident CTAs
  int ksplit = 1;
  if (K > 1 && ctas * 2 <= slots) ksplit = (int)std::min<long long>(K, slots / ctas);
  pl.k_per = (K + ksplit - 1) / ksplit;
  pl.ksplit = (K + pl.k_per - 1) / pl.k_per;
  return pl;
}

size_t spconv_tc_workspace_bytes(long long n_out, int K, int cout) {
  const TcPlan pl = tc_plan(n_out, K, cout, false);
  const TcPlan ps = tc_plan(n_out, K, cout, split_occ() == 2);
  if (ps.ksplit > pl.ksplit) return (size_t)ps.ksplit * (size_t)n_out * cout * sizeof(float);
  return pl.ksplit > 1 ? (size_t)pl.ksplit * (size_t)n_out * cout * sizeof(float) : 0;
}

bool spconv_tc_supported(int cin, int cout) {
  return cin % TC_BK == 0 && cin >= 32 && cin <= 384 && cout % 32 == 0 && cout >= 32 && cout <= 256;
}

// ---- tensor map of the split input rows for MODE 2: bf16 [rows, 2*cin] with row pitch 4*in_ld bytes, box = one
// 128-byte slab row, SWIZZLE_128B.  The driver entry point is resolved once through the runtime (no libcuda link).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int g_last_tmap_rc = 0;     // CUresult of the last cuTensorMapEncodeTiled (diagnostics)
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {                              // a failed lookup is retried on the next call, never cached
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    (void)cudaGetLastError();
  }
  return fn;
}
bool make_row_map(CUtensorMap* tm, const float* in, int in_ld, int cin, long long n_in) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc) { g_last_tmap_rc = -1; return false; }
  // The map must carry the TRUE row count: with an oversized row extent (the neighbour table never names a row
  // outside the buffer, so 2^31 - 1 looked harmless) the TMA unit raised sporadic illegal-address faults on small
  // levels (measured on B200, tools/tma_model_diag.py); absent neighbours (-1) are out of range and read as zeros.
  if (n_in <= 0) return false;
  cuuint64_t strides[1] = {(cuuint64_t)in_ld * 4};
  cuuint64_t dims[2] = {(cuuint64_t)cin * 2, (cuuint64_t)n_in};
  cuuint32_t box[2] = {64, 1};
  cuuint32_t estr[2] = {1, 1};
  const CUresult rc = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<float*>(in), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  g_last_tmap_rc = (int)rc;
  return rc == CUDA_SUCCESS;
}

static bool make_out_map(CUtensorMap* tm, float* out, int out_ld, int cout, long long n_out) {
  EncodeTiledFn enc = encode_tiled();
  if (!enc || n_out <= 0) return false;
  cuuint64_t strides[1] = {(cuuint64_t)out_ld * 4};
  cuuint64_t dims[2] = {(cuuint64_t)cout * 2, (cuuint64_t)n_out};
  cuuint32_t box[2] = {64, (cuuint32_t)TC_BM};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

int spconv_tc_launch(const float* in, long long n_in, int in_ld, int cin, const int* nbr, int K, long long n_out,
                     const void* wprep, int cout, const float* scale, const float* shift, const float* residual,
                     int res_ld, float* out, int out_ld, int flags, void* ws, size_t ws_bytes, cudaStream_t st) {
  AG3D_CHECK_ARG(K <= 32, "the tensor-core path handles at most 32 kernel offsets");
  AG3D_CHECK_ARG(wprep && aligned16(wprep), "prepared weights missing (ag3d_spconv_tc_prepare_weight)");
  TcParams p;
  p.in = in; p.in_ld = in_ld; p.cin = cin; p.nbr = nbr; p.K = K; p.n_out = n_out;
  p.wp = static_cast<const uint4*>(wprep); p.cout = cout;
  p.scale = scale; p.shift = shift; p.residual = residual; p.res_ld = res_ld;
  p.out = out; p.out_ld = out_ld; p.flags = flags;
  p.in_split = (flags & AG3D_IN_SPLIT) ? 1 : 0;
  p.out_split = (flags & AG3D_OUT_SPLIT) ? 1 : 0;
  p.res_split = (flags & AG3D_RES_SPLIT) ? 1 : 0;
  const bool split = p.in_split != 0;
  const bool two = split && split_occ() == 2;     // two CTAs per SM
  const TcPlan plan = tc_plan(n_out, K, cout, two);
  p.cpad = plan.cpad;
  p.T = plan.T;
  p.k_per = plan.k_per;
  const long long tiles = (n_out + TC_BM - 1) / TC_BM;
  p.partial = nullptr;
  {
    static int dbg = -1, ni = -1;
    if (dbg < 0) { const char* e = getenv("AG3D_TC_DEBUG"); dbg = e ? atoi(e) : 0; }
    if (ni < 0) { const char* e = getenv("AG3D_TC_ISSUERS"); ni = e ? atoi(e) : 1; }
    p.debug = dbg;
    p.NI = (ni == 2 && plan.T > 1) ? 2 : 1;   // measured: a second issuing warp does not pay (tools/tc_probe.sh)
  }
  static int force_na = -1;          // tuning experiments only (AG3D_TC_NA = 2 | 4 | 8)
  if (force_na < 0) { const char* e = getenv("AG3D_TC_NA"); force_na = e ? atoi(e) : 0; }
  if (plan.ksplit > 1) {
    AG3D_CHECK_ARG(ws && aligned16(ws) && ws_bytes >= (size_t)plan.ksplit * (size_t)n_out * cout * sizeof(float),
                   "split-K workspace too small (ag3d_spconv_workspace_bytes)");
    p.partial = static_cast<float*>(ws);
  }
  p.tmem_cols = pow2_at_least(p.T * p.cpad, 32);
  p.NB = (cout <= 128 && !two) ? 4 : 2;
  // gather engine of the split-row variant: TMA tile::gather4 (default) or per-thread cp.async (AG3D_TC_GATHER=cpasync)
  static int want_tma = -1;
  if (want_tma < 0) { const char* e = getenv("AG3D_TC_GATHER"); want_tma = (e && e[0] == 'c') ? 0 : 1; }
  alignas(64) CUtensorMap tm_in;
  memset(&tm_in, 0, sizeof(tm_in));
  const bool tma = split && want_tma && make_row_map(&tm_in, in, in_ld, cin, n_in);
  const size_t a_stage = tma ? (size_t)TMA_STAGE : (size_t)A_STAGE;
  const size_t fixed = TC_BAR_BYTES + TC_LIST_BYTES + (split ? 0 : (size_t)p.k_per * p.T * TC_BM * 4) +
                       (size_t)p.NB * (size_t)cout * 128 + (tma ? 1024 : 0);
  const size_t budget = two ? 110 * 1024 : (split ? 224 * 1024 : 200 * 1024);
  int na = (int)((budget - fixed) / a_stage);
  na = na >= 8 ? 8 : (na >= 4 ? 4 : 2);
  if (force_na == 2 || force_na == 4 || force_na == 8) na = std::min(na, force_na);
  p.NA = na;
  p.na_log2 = na == 8 ? 3 : (na == 4 ? 2 : 1);
  p.nb_log2 = p.NB == 4 ? 2 : 1;
  // split-row output through the TMA engine: the finished tile is staged in the operand ring (cout / 32 slabs of 16 KB)
  static int want_tma_out = -1;
  if (want_tma_out < 0) { const char* e = getenv("AG3D_TC_TMA_OUT"); want_tma_out = (e && e[0] == '0') ? 0 : 1; }
  alignas(64) CUtensorMap tm_out;
  memset(&tm_out, 0, sizeof(tm_out));
  p.tma_out = (tma && want_tma_out && p.out_split && plan.ksplit == 1 && cout / 32 <= na && n_out < 2147483647LL &&
               make_out_map(&tm_out, out, out_ld, cout, n_out)) ? 1 : 0;
  size_t smem = fixed + (size_t)na * a_stage;
  if (!two) smem = std::max(smem, (size_t)116 * 1024);   // one CTA per SM: it may allocate all 512 TMEM columns
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AG3D_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AG3D_CUDA(cudaFuncSetAttribute(spconv_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  const dim3 grid((unsigned)((tiles + p.T - 1) / p.T), (unsigned)plan.ksplit);
  // programmatic dependent launch (opt-in, AG3D_PDL=1): the grid may start while its predecessor drains; it does its
  // prologue (barriers, TMEM allocation, stage list from the neighbour table) and then waits for the predecessor.
  // Measured on the headline step: 22.95 vs 23.01 ms at batch 8, 5.71 vs 5.62 ms at batch 1 (profiles/r02_pdl_ab.txt) -
  // no gain, because the SMs have no room for a second grid's CTAs until the first one's leave, so it stays off.
  static int pdl = -1;
  if (pdl < 0) { const char* e = getenv("AG3D_PDL"); pdl = (e && e[0] == '1') ? 1 : 0; }
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  lattr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cfg.attrs = lattr; cfg.numAttrs = pdl ? 1 : 0;
  if (tma) AG3D_CUDA(cudaLaunchKernelEx(&cfg, spconv_tc_kernel<2>, tm_in, tm_out, p));
  else if (split) AG3D_CUDA(cudaLaunchKernelEx(&cfg, spconv_tc_kernel<1>, tm_in, tm_out, p));
  else AG3D_CUDA(cudaLaunchKernelEx(&cfg, spconv_tc_kernel<0>, tm_in, tm_out, p));
  AG3D_LAUNCH_CHECK("spconv_tc");
  if (plan.ksplit > 1) {
    const long long total = n_out * (cout / 16);
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
    cudaLaunchConfig_t rc{};
    rc.gridDim = dim3((unsigned)blocks); rc.blockDim = dim3(256); rc.dynamicSmemBytes = 0; rc.stream = st;
    rc.attrs = lattr; rc.numAttrs = pdl ? 1 : 0;
    const float* partial_c = p.partial;
    AG3D_CUDA(cudaLaunchKernelEx(&rc, splitk_reduce_kernel, partial_c, plan.ksplit, n_out, cout, scale, shift, residual, res_ld,
                                 out, out_ld, flags));
    AG3D_LAUNCH_CHECK("splitk_reduce");
  }
  return AG3D_OK;
}

}  // namespace ag3d

using namespace ag3d;

extern "C" {

size_t ag3d_spconv_tc_weight_bytes(int32_t K, int32_t cin, int32_t cout) {
  return (size_t)K * (size_t)(cin / TC_BK) * (size_t)cout * 128;
}

int ag3d_spconv_tc_prepare_weight(const float* weight, int32_t K, int32_t cin, int32_t cout, void* wprep,
                                  ag3d_stream_t stream) {
  AG3D_CHECK_ARG(K >= 1 && spconv_tc_supported(cin, cout), "shape not supported by the tensor-core path");
  AG3D_CHECK_ARG(weight && wprep && aligned16(wprep), "bad pointers");
  const long long total = (long long)K * (cin / TC_BK) * 4 * cout;
  long long blocks = (total + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  weight_prep_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(weight, K, cin, cout, static_cast<uint4*>(wprep));
  AG3D_LAUNCH_CHECK("weight_prep");
  return AG3D_OK;
}

}  // extern "C"
