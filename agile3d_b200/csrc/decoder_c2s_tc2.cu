// K9 (split-row variant): click -> scene cross-attention with the voxel tiles fed by the TMA engine.
//
//   ctx[(h,q), :] = sum_v softmax_v( qfold[(h,q)] . (x_v + pos_v)  [label mask] ) * x_v
//
// Same algorithm and partial-result format as decoder_c2s_tc.cu (flash decoding over 64-voxel tiles, folded queries of
// one query group as the stationary M = 128 operand, online softmax with the context accumulator in TMEM, log-sum-exp
// merge by c2s_merge_kernel).  What changed is the operand path: x and pos arrive as "split" rows (every 32-channel
// slab = 64 B bf16 hi | 64 B bf16 lo, the activation format of the tensor-core backbone), so a voxel tile is eight 2-D
// TMA box loads into SWIZZLE_128B shared-memory tiles that the tensor core reads directly:
//   S   = Qf . x^T + Qf . pos^T   [128 x 64]   K-major B tiles, bf16x3 per term; no (x + pos) tile is ever formed
//   ctx += P . [x_hi | x_lo]      [128 x 256]  the same x tile as an MN-major B operand: every 128-byte slab row is 32
//                                              hi and 32 lo channels, so hi and lo products land in separate column
//                                              halves (all four products, "bf16x4") and are folded once at the end
// No thread touches a voxel byte: the 256 loader threads, their global-load latency and the two fp32 -> bf16 hi/lo
// conversions per element of the fp32 variant are gone; separate rings for the x tiles (3 deep: alive until the context GEMM) and the pos tiles (2 deep: dead after the score GEMM) keep two tiles in flight.  The
// probabilities never touch shared memory either (A operand of the context GEMM in tensor memory, see C2_TM_P0).
// Roles: 16 softmax warps (four threads per (head, query) row), 1 MMA-issuer warp, 1 TMA-producer warp.
#include <cuda.h>
#include <float.h>
#include <math.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace ag3d {

constexpr int C2_D = 128;
constexpr int C2_TV = 64;                                // voxels per tile
constexpr int C2_SOFT_WARPS = 16;                        // four threads per (head, query) row, 16 voxels of the tile each
constexpr int C2_THREADS = C2_SOFT_WARPS * 32 + 64;      // + MMA warp + TMA warp
constexpr uint32_t C2_SLAB = C2_TV * 128;                // [64 voxels x 128 B] slab tile (32 channels hi | lo)
constexpr uint32_t C2_TILE = 4 * C2_SLAB;                // x or pos tile: 4 slabs
constexpr int C2_NX = 3;                                 // x ring: a tile lives until its context GEMM has run
constexpr int C2_NP = 2;                                 // pos ring: a tile is dead as soon as its score GEMM has run
constexpr uint32_t C2_OFF_POS = C2_NX * C2_TILE;
constexpr uint32_t C2_QLBO = 128 * 16;                   // Q pieces: no-swizzle K-major, 128 rows x 16 B per 8-channel chunk
constexpr uint32_t C2_QPIECE = 16 * C2_QLBO;             // [128 rows x 128 channels] bf16
constexpr uint32_t C2_OFF_Q = C2_OFF_POS + C2_NP * C2_TILE;
constexpr uint32_t C2_OFF_MISC = C2_OFF_Q + 2 * C2_QPIECE;
constexpr uint32_t C2_MISC = 3072;                       // barriers | TMEM slot | tile-maximum exchange (aliased: partial sums)
constexpr size_t C2_SMEM = C2_OFF_MISC + C2_MISC;        // 232448 B = all 227 KB a CTA may have
// TMEM columns: two score buffers [128 x 64] fp32, the [128 x 256] context accumulator, and two probability buffers -
// P is the A operand of the context GEMM and lives in TENSOR MEMORY ([128 rows x 64 voxels] bf16 = 32 columns hi + 32
// columns lo, two voxels per 32-bit column), written by the softmax threads with tcgen05.st: no shared-memory stores,
// no proxy fence, and with two buffers the softmax of tile i+1 runs under the context GEMM of tile i
constexpr uint32_t C2_TM_S0 = 0, C2_TM_S1 = 64, C2_TM_CTX = 128, C2_TM_P0 = 384, C2_TM_P1 = 448;

struct C2sSplitParams {
  long long nv;
  const float* qfold; int nq; int heads; int nqg;
  const unsigned char* label; const int* q_obj; const int* obj_count;
  float* part_m; float* part_l; float* part_acc;
  int debug;      // measurement aid (AG3D_C2S_DEBUG): 1 no score MMAs, 2 no context MMAs, 4 no tile loads, 8 no softmax math
};

__device__ __forceinline__ void c2_tma_tile(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int col, int row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(col), "r"(row)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]: the A operand (bf16, two K elements per 32-bit column, row = lane) comes from tensor memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ uint32_t c2_row_off(int r) { return (uint32_t)((r >> 3) * 128 + (r & 7) * 16); }

__global__ void __launch_bounds__(C2_THREADS, 1)
c2s_split_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_pos, const C2sSplitParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* Qs = smem + C2_OFF_Q;                 // hi | lo
  unsigned char* misc = smem + C2_OFF_MISC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 128);
  float* tmax_s = reinterpret_cast<float*>(misc + 256);          // [4 parts][128 rows]
  float* lsum_s = tmax_s;                                        // [4 parts][128] (after the last tile)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_base = smem_u32(bars);
  auto full_x = [&](int s) { return bar_base + 8u * s; };
  auto empty_x = [&](int s) { return bar_base + 8u * (3 + s); };
  auto full_p = [&](int s) { return bar_base + 8u * (6 + s); };
  auto empty_p = [&](int s) { return bar_base + 8u * (8 + s); };
  auto s_full = [&](int t) { return bar_base + 8u * (10 + (t & 1)); };
  auto p_ready = [&](int t) { return bar_base + 8u * (12 + (t & 1)); };  // one per probability buffer
  auto g2_done = [&](int t) { return bar_base + 8u * (14 + (t & 1)); };
  const int g = blockIdx.y;
  const int q0 = g * p.nqg;
  const int nq_here = min(p.nqg, p.nq - q0);
  const int HQ = p.heads * nq_here;                    // <= 128

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();              // SWIZZLE_128B tiles need 1024-byte alignment
    for (int s = 0; s < C2_NX; ++s) { mbar_init(full_x(s), 1); mbar_init(empty_x(s), 1); }
    for (int s = 0; s < C2_NP; ++s) { mbar_init(full_p(s), 1); mbar_init(empty_p(s), 1); }
    mbar_init(s_full(0), 1);
    mbar_init(s_full(1), 1);
    mbar_init(p_ready(0), C2_SOFT_WARPS);
    mbar_init(p_ready(1), C2_SOFT_WARPS);
    mbar_init(g2_done(0), 1);
    mbar_init(g2_done(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == C2_SOFT_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // folded queries of this group -> stationary A operand (bf16 hi/lo, K-major, no swizzle): row r = h*nq_here + ql
  for (int idx = tid; idx < 128 * 16; idx += C2_THREADS) {
    const int r = idx & 127, cc = idx >> 7;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (r < HQ) {
      const int h = r / nq_here, ql = r % nq_here;
      const float4* src = reinterpret_cast<const float4*>(p.qfold + ((size_t)h * p.nq + q0 + ql) * C2_D + cc * 8);
      a = __ldg(src);
      b = __ldg(src + 1);
    }
    uint32_t h4[4], l4[4];
    split2(a.x, a.y, h4[0], l4[0]);
    split2(a.z, a.w, h4[1], l4[1]);
    split2(b.x, b.y, h4[2], l4[2]);
    split2(b.z, b.w, h4[3], l4[3]);
    unsigned char* dst = Qs + cc * C2_QLBO + c2_row_off(r);
    *reinterpret_cast<uint4*>(dst) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
    *reinterpret_cast<uint4*>(dst + C2_QPIECE) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long n_tiles = (p.nv + C2_TV - 1) / C2_TV;
  const int n_my = n_tiles > blockIdx.x ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

  if (warp < C2_SOFT_WARPS) {
    // ======================================================================================= softmax warps
    // Four threads per (head, query) row: warp 4*part + q takes voxels 16*part .. 16*part+15 of the tile for the rows of TMEM
    // lane quarter q.  The four exchange their tile maxima through shared memory (named barrier 1 + q, 128 threads).
    const int q = warp & 3, part = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
    int ro = -2;                                                 // object the row is restricted to; -1 none; -2 padding
    if (r < HQ) {
      ro = -1;
      if (p.label) {
        const int o = p.q_obj[q0 + r % nq_here];
        if (p.obj_count[o] > 0) ro = o;                          // all-masked rows are un-masked (agile3d.py:369,375)
      }
    }
    float m_ref = -INFINITY, l_sum = 0.f;
    for (int it = 0; it < n_my; ++it) {
      const long long v0 = ((long long)blockIdx.x + (long long)it * gridDim.x) * C2_TV + part * 16;
      // labels of this thread's 16 voxels (254 = no label mask, 255 = past the end), fetched before the scores are awaited
      uint32_t labw[4];
      if (p.label && v0 + 16 <= p.nv) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.label + v0));
        labw[0] = a.x; labw[1] = a.y; labw[2] = a.z; labw[3] = a.w;
      } else if (!p.label && v0 + 16 <= p.nv) {
#pragma unroll
        for (int w = 0; w < 4; ++w) labw[w] = 0xFEFEFEFEu;
      } else {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          uint32_t acc = 0;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const long long v = v0 + w * 4 + e;
            const uint32_t lab = v < p.nv ? (p.label ? (uint32_t)p.label[v] : 254u) : 255u;
            acc |= lab << (8 * e);
          }
          labw[w] = acc;
        }
      }
      mbar_wait(s_full(it), ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      float s[16];
      tmem_ld16(t_lane + ((it & 1) ? C2_TM_S1 : C2_TM_S0) + (uint32_t)part * 16u, s);
      float tmax = -INFINITY;
#pragma unroll
      for (int v = 0; v < 16; ++v) {
        const int lab = (int)((labw[v >> 2] >> (8 * (v & 3))) & 0xFFu);
        const bool dead = (lab == 255) || (ro == -2) || (ro >= 0 && lab != ro);
        s[v] = dead ? -INFINITY : s[v];
        tmax = fmaxf(tmax, s[v]);
      }
      float* tm = tmax_s + r;
      tm[part * 128] = tmax;
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      tmax = fmaxf(fmaxf(tm[0], tm[128]), fmaxf(tm[256], tm[384]));
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");   // all four have read before the next tile overwrites
      float alpha = 1.f;
      const bool grow = tmax > m_ref;
      if (grow) {
        alpha = (m_ref == -INFINITY) ? 0.f : __expf(m_ref - tmax);
        m_ref = tmax;
        l_sum *= alpha;
      }
      // probabilities of this tile -> P buffer it & 1 (bf16 hi | lo, two voxels per column); the buffer was last read by
      // the context GEMM of tile it - 2
      if (it > 1) {
        mbar_wait(g2_done(it), (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
      }
      float psum = 0.f;
      const uint32_t pcol = ((it & 1) ? C2_TM_P1 : C2_TM_P0) + (uint32_t)part * 8u;
      uint32_t ph_[8], pl_[8];
      if (p.debug & 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) ph_[e] = pl_[e] = 0u;
      } else
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float p0 = (m_ref == -INFINITY) ? 0.f : __expf(s[2 * e] - m_ref);
        const float p1 = (m_ref == -INFINITY) ? 0.f : __expf(s[2 * e + 1] - m_ref);
        psum += p0 + p1;
        split2(p0, p1, ph_[e], pl_[e]);
      }
      tmem_st8(t_lane + pcol, ph_);
      tmem_st8(t_lane + pcol + 32u, pl_);
      l_sum += psum;
      if (it > 0 && __any_sync(0xffffffffu, grow)) {             // rescale this thread's quarter of its context row in TMEM:
        mbar_wait(g2_done(it - 1), ((uint32_t)(it - 1) >> 1) & 1u);   // the context GEMM of the previous tile must be done
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float c[16];
          tmem_ld16(t_lane + C2_TM_CTX + part * 64 + ch * 16, c);
#pragma unroll
          for (int e = 0; e < 16; ++e) c[e] *= alpha;
          tmem_st16(t_lane + C2_TM_CTX + part * 64 + ch * 16, c);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready(it));
    }
    // ---- partial result of this CTA: (m, l, ctx row); the four threads of a row add their partial sums; thread `part`
    //      owns the accumulator columns of 32-channel slab `part`, whose hi / lo column halves are folded here
    if (n_my > 0) {
      mbar_wait(g2_done(n_my - 1), ((uint32_t)(n_my - 1) >> 1) & 1u);
      tc_fence_after();
    }
    lsum_s[part * 128 + r] = l_sum;
    asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
    const size_t prow = ((size_t)g * gridDim.x + blockIdx.x) * 128 + r;
    if (part == 0) {
      p.part_m[prow] = m_ref;
      p.part_l[prow] = (lsum_s[r] + lsum_s[128 + r]) + (lsum_s[256 + r] + lsum_s[384 + r]);
    }
    {
      float hi[32], lo[32];
      if (n_my > 0) {
        tmem_ld16(t_lane + C2_TM_CTX + part * 64, hi);
        tmem_ld16(t_lane + C2_TM_CTX + part * 64 + 16, hi + 16);
        tmem_ld16(t_lane + C2_TM_CTX + part * 64 + 32, lo);
        tmem_ld16(t_lane + C2_TM_CTX + part * 64 + 48, lo + 16);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) hi[e] = lo[e] = 0.f;
      }
      float* dst = p.part_acc + prow * C2_D + part * 32;
#pragma unroll
      for (int e4 = 0; e4 < 8; ++e4)
        *reinterpret_cast<float4*>(dst + e4 * 4) = make_float4(hi[e4 * 4] + lo[e4 * 4], hi[e4 * 4 + 1] + lo[e4 * 4 + 1],
                                                               hi[e4 * 4 + 2] + lo[e4 * 4 + 2], hi[e4 * 4 + 3] + lo[e4 * 4 + 3]);
    }
    tc_fence_before();
  } else if (warp == C2_SOFT_WARPS) {
    // ======================================================================================= MMA issuer
    const uint32_t id1 = umma_idesc_bf16_major(C2_TV, 0, 0);     // S: N = 64, both operands K-major
    const uint32_t id2 = umma_idesc_bf16_major(256, 0, 1);       // ctx: N = 256 (hi | lo per slab), B (= x tile) MN-major
    const uint32_t sw_hi32 = umma_desc_hi32(1024) | (2u << 29);  // SWIZZLE_128B, 1024 B between 8-row groups
    const uint32_t q_hi = smem_u32(Qs), q_lo = q_hi + C2_QPIECE;
    const uint32_t stage0 = smem_u32(smem);
    // The issuer serves whichever is ready: the score GEMM S[t & 1] = Qf . x^T + Qf . pos^T of the next tile (its
    // stage has landed and its score buffer has been consumed) or the context GEMM of the oldest tile whose
    // probabilities are written.  Waiting for them in a fixed order would chain the load latency of tile t + 2 in
    // front of the context GEMM of tile t + 1, i.e. one HBM round trip per tile.
    auto issue_scores = [&](int t) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = tmem_base + ((t & 1) ? C2_TM_S1 : C2_TM_S0);
        const uint32_t tb[2] = {stage0 + (uint32_t)(t % C2_NX) * C2_TILE, stage0 + C2_OFF_POS + (uint32_t)(t % C2_NP) * C2_TILE};
        if (!(p.debug & 1))
#pragma unroll
        for (int j = 0; j < 8; ++j) {                            // K = 128 channels, 16 per step: slab j >> 1, half j & 1
          const uint64_t a_h = umma_desc(q_hi + j * 2 * C2_QLBO, C2_QLBO, 128), a_l = umma_desc(q_lo + j * 2 * C2_QLBO, C2_QLBO, 128);
#pragma unroll
          for (int o = 0; o < 2; ++o) {                          // operand: x tile, pos tile
            const uint32_t b32 = umma_desc_lo32(tb[o] + (uint32_t)(j >> 1) * C2_SLAB, 16) + (uint32_t)(j & 1) * 2u;
            const uint64_t b_h = umma_desc_join(sw_hi32, b32), b_l = umma_desc_join(sw_hi32, b32 + 4u);
            umma_bf16(d, a_h, b_h, id1, (j | o) ? 1u : 0u);
            umma_bf16(d, a_h, b_l, id1, 1u);
            umma_bf16(d, a_l, b_h, id1, 1u);
          }
        }
        umma_commit(s_full(t));
        umma_commit(empty_p(t % C2_NP));                         // the pos tile is dead
      }
      __syncwarp();
    };
    auto issue_context = [&](int it) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t xb = stage0 + (uint32_t)(it % C2_NX) * C2_TILE;
        if (!(p.debug & 2))
#pragma unroll
        for (int j = 0; j < 4; ++j) {                            // K = 64 voxels, 16 per step = 8 TMEM columns of P, 2048 B of x
          const uint32_t pa = tmem_base + ((it & 1) ? C2_TM_P1 : C2_TM_P0) + (uint32_t)j * 8u;
          // MN-major SWIZZLE_128B: LBO = bytes between 64-element atoms along N (one slab tile), SBO = 8-row groups
          const uint64_t b = umma_desc_join(sw_hi32, umma_desc_lo32(xb + (uint32_t)j * 2048u, C2_SLAB));
          umma_bf16_ts(tmem_base + C2_TM_CTX, pa, b, id2, (it | j) ? 1u : 0u);
          umma_bf16_ts(tmem_base + C2_TM_CTX, pa + 32u, b, id2, 1u);
        }
        umma_commit(g2_done(it));
        umma_commit(empty_x(it % C2_NX));
      }
      __syncwarp();
    };
    int ns = 0, nc = 0;                                          // next score / context GEMM to issue
    unsigned spins = 0;
    while (nc < n_my) {
      uint32_t ready = 0;                                        // lane 0 polls, the warp follows (warp-uniform control flow)
      if (lane == 0) {
        if (ns < n_my && ns <= nc + 1 && mbar_try(full_x(ns % C2_NX), (uint32_t)(ns / C2_NX) & 1u) &&
            mbar_try(full_p(ns % C2_NP), (uint32_t)(ns / C2_NP) & 1u)) ready = 1;
        else if (mbar_try(p_ready(nc), ((uint32_t)nc >> 1) & 1u)) ready = 2;
      }
      ready = __shfl_sync(0xffffffffu, ready, 0);
      if (ready == 1) { issue_scores(ns); ++ns; spins = 0; }
      else if (ready == 2) { issue_context(nc); ++nc; spins = 0; }
      else if (++spins > SPIN_LIMIT) __trap();
    }
  } else {
    // ======================================================================================= TMA producer
    const uint32_t stage0 = smem_u32(smem);
    for (int t = 0; t < n_my; ++t) {
      const int row = (int)(((long long)blockIdx.x + (long long)t * gridDim.x) * C2_TV);
      const int sx = t % C2_NX, sp = t % C2_NP;
      mbar_wait(empty_x(sx), ((uint32_t)(t / C2_NX) & 1u) ^ 1u);
      if (elect_one()) {
        const uint32_t dst = stage0 + (uint32_t)sx * C2_TILE;
        if (p.debug & 4) {
          mbar_arrive(full_x(sx));
        } else {
          mbar_arrive_expect_tx(full_x(sx), C2_TILE);
#pragma unroll
          for (int c = 0; c < 4; ++c) c2_tma_tile(dst + (uint32_t)c * C2_SLAB, &tm_x, full_x(sx), c * 64, row);
        }
      }
      __syncwarp();
      mbar_wait(empty_p(sp), ((uint32_t)(t / C2_NP) & 1u) ^ 1u);
      if (elect_one()) {
        const uint32_t dst = stage0 + C2_OFF_POS + (uint32_t)sp * C2_TILE;
        if (p.debug & 4) {
          mbar_arrive(full_p(sp));
        } else {
          mbar_arrive_expect_tx(full_p(sp), C2_TILE);
#pragma unroll
          for (int c = 0; c < 4; ++c) c2_tma_tile(dst + (uint32_t)c * C2_SLAB, &tm_pos, full_p(sp), c * 64, row);
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == C2_SOFT_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side
typedef CUresult (*C2EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static C2EncodeFn c2_encode() {
  static C2EncodeFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<C2EncodeFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}
// split rows [rows, 128 channels] (512 B per row) -> tensor map of bf16 [rows, 256] with a [box_rows x 64] box
bool split_rows_tile_map(CUtensorMap* tm, const float* base, long long rows, int box_rows) {
  C2EncodeFn enc = c2_encode();
  if (!enc || rows <= 0) return false;
  cuuint64_t dims[2] = {256, (cuuint64_t)rows};
  cuuint64_t strides[1] = {512};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<float*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t c2s_tc_workspace_bytes(int nq, int heads);

int c2s_split_launch(const float* x_split, const float* pos_split, long long nv, const float* qfold, int nq, int heads,
                     const unsigned char* label, const int* q_obj, const int* obj_count, void* ws, size_t ws_bytes,
                     cudaStream_t st, float** part_m, float** part_l, float** part_acc, int* n_cta_out, int* nqg_out) {
  AG3D_CHECK_ARG(heads == 8, "tensor-core c2s handles 8 heads");
  AG3D_CHECK_ARG(nv < 2147483647LL, "too many voxels");
  AG3D_CHECK_ARG(ws && aligned16(ws) && ws_bytes >= c2s_tc_workspace_bytes(nq, heads), "c2s workspace too small");
  AG3D_CHECK_ARG(!label || aligned16(label), "label vector must be 16-byte aligned");
  const int groups = (nq + 15) / 16;
  const int nqg = (nq + groups - 1) / groups;
  const long long n_tiles = (nv + C2_TV - 1) / C2_TV;
  int n_cta = std::max(1, sm_count() / groups);
  if (n_cta > n_tiles) n_cta = (int)n_tiles;
  alignas(64) CUtensorMap tm_x, tm_pos;
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_pos, 0, sizeof(tm_pos));
  AG3D_CHECK_ARG(split_rows_tile_map(&tm_x, x_split, nv, C2_TV) && split_rows_tile_map(&tm_pos, pos_split, nv, C2_TV),
                 "cuTensorMapEncodeTiled failed for the voxel rows");
  C2sSplitParams p;
  p.nv = nv; p.qfold = qfold; p.nq = nq; p.heads = heads; p.nqg = nqg;
  p.label = label; p.q_obj = q_obj; p.obj_count = obj_count;
  { const char* e = getenv("AG3D_C2S_DEBUG"); p.debug = e ? atoi(e) : 0; }
  p.part_m = static_cast<float*>(ws);
  p.part_l = p.part_m + (size_t)groups * n_cta * 128;
  p.part_acc = p.part_l + (size_t)groups * n_cta * 128;
  p.part_acc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p.part_acc) + 15) & ~(uintptr_t)15);
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(c2s_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C2_SMEM));
    attr = true;
  }
  c2s_split_kernel<<<dim3(n_cta, groups), C2_THREADS, C2_SMEM, st>>>(tm_x, tm_pos, p);
  AG3D_LAUNCH_CHECK("c2s_split");
  *part_m = p.part_m; *part_l = p.part_l; *part_acc = p.part_acc; *n_cta_out = n_cta; *nqg_out = nqg;
  return AG3D_OK;
}

}  // namespace ag3d
