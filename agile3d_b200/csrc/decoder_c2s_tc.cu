// K9 (tensor-core variant): click -> scene cross-attention as a streaming flash-decoding kernel on tcgen05.
//
//   ctx[(h,q), :] = sum_v softmax_v( qfold[(h,q)] . (x_v + pos_v)  [label mask] ) * x_v
//
// One CTA per SM streams its share of the voxels in tiles of 64; the folded queries of one query group (<= 16
// queries = 128 (head, query) rows) are the stationary M = 128 operand.  Per tile:
//   S   = Qf . (x+pos)^T       [128 x 64]   tcgen05.mma bf16x3, accumulator in TMEM
//   online softmax over the voxels: one thread per (head, query) row reads its 64 scores from TMEM, applies the
//       label mask, keeps a running max / sum; when the max moves the context accumulator row is rescaled in TMEM
//   ctx += P . X               [128 x 128]  tcgen05.mma bf16x3, accumulator resident in TMEM for the whole kernel;
//       X is consumed as an MN-major B operand straight from the row-major voxel tile (no transpose)
// Roles: 8 softmax warps (two threads per (head, query) row, 32 voxels each, maxima exchanged through shared memory),
// 8 loader warps (next tile's x/pos -> bf16 hi/lo operand tiles while the current tile computes), 1 MMA-issuer warp
// that keeps the score GEMM one tile ahead of the softmax (two score buffers in TMEM).  Per-CTA partial
// (max, sum, ctx) results are merged by c2s_merge_kernel (log-sum-exp), as for the SIMT variant.
#include <float.h>
#include <math.h>

#include <algorithm>

#include "tc_common.cuh"

namespace ag3d {

constexpr int CT_D = 128;
constexpr int CT_TV = 64;                          // voxels per tile
constexpr int CT_SOFT_THREADS = 128;
constexpr int CT_LOAD_THREADS = 256;
constexpr int CT_THREADS = CT_SOFT_THREADS + CT_LOAD_THREADS + 32 + 128;   // + MMA warp + 4 softmax helper warps
constexpr int T_LBO = CT_TV * 16 + 16;             // 64-row pieces: bytes between adjacent 8-element chunks (padded)
constexpr int Q_PIECE = 16 * A_LBO;                // [128 rows x 128 ch] piece: 16 channel chunks
constexpr int T_PIECE = 16 * T_LBO;                // [64 voxels x 128 ch] piece
constexpr int P_PIECE = 8 * A_LBO;                 // [128 rows x 64 voxels] piece: 8 voxel chunks
constexpr int CT_MISC = 4096;
constexpr size_t CT_SMEM = CT_MISC + 2 * (size_t)Q_PIECE + 4 * (size_t)T_PIECE + 2 * (size_t)P_PIECE;
constexpr uint32_t CT_TM_S = 0, CT_TM_CTX = 64, CT_TM_S1 = 192;   // two score buffers: MMA1 of tile i+1 runs under softmax(i)

struct C2sParams {
  const float* x; const float* pos; long long nv;
  const float* qfold; int nq; int heads; int nqg;
  const unsigned char* label; const int* q_obj; const int* obj_count;
  float* part_m; float* part_l; float* part_acc;
};

__device__ __forceinline__ uint32_t row_off(int r) { return (uint32_t)((r >> 3) * 128 + (r & 7) * 16); }

__global__ void __launch_bounds__(CT_THREADS, 1) c2s_tc_kernel(const C2sParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 64);
  unsigned char* lab_s = smem + 128;                 // [64] label of each voxel of the tile (254 none, 255 invalid)
  unsigned char* Qs = smem + CT_MISC;                // hi | lo
  unsigned char* XPs = Qs + 2 * Q_PIECE;
  unsigned char* Xs = XPs + 2 * T_PIECE;
  unsigned char* Ps = Xs + 2 * T_PIECE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t xp_ready = bar_base, x_ready = bar_base + 8, p_ready = bar_base + 24, g2_done = bar_base + 32;
  auto s_full = [&](int t) { return bar_base + ((t & 1) ? 40u : 16u); };       // one barrier per score buffer
  float* tmax_s = reinterpret_cast<float*>(smem + 512);          // [2 tiles][2 halves][128 rows] pair exchange
  float* lsum_s = reinterpret_cast<float*>(smem + 512 + 2048);   // [128]
  const int g = blockIdx.y;
  const int q0 = g * p.nqg;
  const int nq_here = min(p.nqg, p.nq - q0);
  const int HQ = p.heads * nq_here;                  // <= 128

  if (tid == 0) {
    mbar_init(xp_ready, CT_LOAD_THREADS / 32);
    mbar_init(x_ready, CT_LOAD_THREADS / 32);
    mbar_init(p_ready, 2 * CT_SOFT_THREADS / 32);
    mbar_init(s_full(0), 1);
    mbar_init(s_full(1), 1);
    mbar_init(g2_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // folded queries of this group -> stationary A operand (bf16 hi/lo, K-major): row r = h*nq_here + ql
  for (int idx = tid; idx < 128 * 16; idx += CT_THREADS) {
    const int r = idx >> 4, cc = idx & 15;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (r < HQ) {
      const int h = r / nq_here, ql = r % nq_here;
      const float4* src = reinterpret_cast<const float4*>(p.qfold + ((size_t)h * p.nq + q0 + ql) * CT_D + cc * 8);
      a = __ldg(src);
      b = __ldg(src + 1);
    }
    uint32_t h4[4], l4[4];
    split2(a.x, a.y, h4[0], l4[0]);
    split2(a.z, a.w, h4[1], l4[1]);
    split2(b.x, b.y, h4[2], l4[2]);
    split2(b.z, b.w, h4[3], l4[3]);
    unsigned char* dst = Qs + cc * A_LBO + row_off(r);
    *reinterpret_cast<uint4*>(dst) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
    *reinterpret_cast<uint4*>(dst + Q_PIECE) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long n_tiles = (p.nv + CT_TV - 1) / CT_TV;

  if (warp < 4 || warp >= 13) {
    // ======================================================================================= softmax warps
    // Two threads per (head, query) row: warp q (q = 0..3) takes voxels 0..31 of the tile, helper warp 13 + ((q + 3) & 3)
    // (the warp with warp % 4 == q: TMEM lane quarter) voxels 32..63.  The pair exchanges its tile maxima through
    // shared memory (named barrier 1 + q), so both follow the same running maximum.
    const int q = warp & 3;
    const int half = warp < 4 ? 0 : 1;
    const int r = q * 32 + lane;                                 // (head, query) row of this thread
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
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      mbar_wait(s_full(it), ((uint32_t)it >> 1) & 1u);
      mbar_wait(x_ready, ph);                                    // labels of this tile are in lab_s
      tc_fence_after();
      float s[32];
      const uint32_t scol = ((it & 1) ? CT_TM_S1 : CT_TM_S) + (uint32_t)half * 32u;
      tmem_ld16(t_lane + scol, s);
      tmem_ld16(t_lane + scol + 16, s + 16);
      float tmax = -INFINITY;
#pragma unroll
      for (int v = 0; v < 32; ++v) {
        const int lab = lab_s[half * 32 + v];
        const bool dead = (lab == 255) || (ro == -2) || (ro >= 0 && lab != ro);
        s[v] = dead ? -INFINITY : s[v];
        tmax = fmaxf(tmax, s[v]);
      }
      tmax_s[((it & 1) * 2 + half) * 128 + r] = tmax;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
      tmax = fmaxf(tmax, tmax_s[((it & 1) * 2 + (half ^ 1)) * 128 + r]);
      if (it > 0) {                                              // GEMM2 of the previous tile is done:
        mbar_wait(g2_done, ph ^ 1u);                             // P buffer reusable, ctx accumulator quiescent
        tc_fence_after();
      }
      float alpha = 1.f;
      const bool grow = tmax > m_ref;
      if (grow) {
        alpha = (m_ref == -INFINITY) ? 0.f : __expf(m_ref - tmax);
        m_ref = tmax;
        l_sum *= alpha;
      }
      if (it > 0 && __any_sync(0xffffffffu, grow)) {             // rescale this thread's half of its context row in TMEM
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float c[16];
          tmem_ld16(t_lane + CT_TM_CTX + half * 64 + ch * 16, c);
#pragma unroll
          for (int e = 0; e < 16; ++e) c[e] *= alpha;
          tmem_st16(t_lane + CT_TM_CTX + half * 64 + ch * 16, c);
        }
        tmem_st_wait();
      }
      float psum = 0.f;
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        float pv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          pv[e] = (m_ref == -INFINITY) ? 0.f : __expf(s[c8 * 8 + e] - m_ref);
          psum += pv[e];
        }
        uint32_t h4[4], l4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2(pv[2 * e], pv[2 * e + 1], h4[e], l4[e]);
        unsigned char* dst = Ps + (half * 4 + c8) * A_LBO + row_off(r);
        *reinterpret_cast<uint4*>(dst) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
        *reinterpret_cast<uint4*>(dst + P_PIECE) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
      }
      l_sum += psum;
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // ---- partial result of this CTA: (m, l, ctx row); the pair adds its two partial sums
    mbar_wait(g2_done, (uint32_t)(it - 1) & 1u);
    tc_fence_after();
    if (half == 1) lsum_s[r] = l_sum;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
    const size_t prow = ((size_t)g * gridDim.x + blockIdx.x) * 128 + r;
    if (half == 0) {
      p.part_m[prow] = m_ref;
      p.part_l[prow] = l_sum + lsum_s[r];
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float c[16];
      tmem_ld16(t_lane + CT_TM_CTX + half * 64 + ch * 16, c);
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4)
        *reinterpret_cast<float4*>(p.part_acc + prow * CT_D + half * 64 + ch * 16 + e4 * 4) =
            make_float4(c[e4 * 4], c[e4 * 4 + 1], c[e4 * 4 + 2], c[e4 * 4 + 3]);
    }
    tc_fence_before();
  } else if (warp < 12) {
    // ======================================================================================= loader warps (4..11)
    const int lt = tid - CT_SOFT_THREADS;
    const int cc = lt & 15, rb = lt >> 4;                        // 8 channels x rows rb + 16 i
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const long long v0 = tile * CT_TV;
      float4 xv[4][2], pv[4][2];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long row = v0 + rb + 16 * i;
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
          xv[i][hlf] = pv[i][hlf] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row < p.nv) {
            xv[i][hlf] = __ldg(reinterpret_cast<const float4*>(p.x + (size_t)row * CT_D + cc * 8 + hlf * 4));
            pv[i][hlf] = __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)row * CT_D + cc * 8 + hlf * 4));
          }
        }
      }
      int lab = 255;
      if (lt < CT_TV && v0 + lt < p.nv) lab = p.label ? (int)p.label[v0 + lt] : 254;
      if (it > 0) mbar_wait(s_full(it - 1), ((uint32_t)(it - 1) >> 1) & 1u);   // score GEMM of the previous tile has read XP
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = xv[i][0], b = xv[i][1], c = pv[i][0], d = pv[i][1];
        uint32_t h4[4], l4[4];
        split2(a.x + c.x, a.y + c.y, h4[0], l4[0]);
        split2(a.z + c.z, a.w + c.w, h4[1], l4[1]);
        split2(b.x + d.x, b.y + d.y, h4[2], l4[2]);
        split2(b.z + d.z, b.w + d.w, h4[3], l4[3]);
        unsigned char* dst = XPs + cc * T_LBO + row_off(rb + 16 * i);
        *reinterpret_cast<uint4*>(dst) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
        *reinterpret_cast<uint4*>(dst + T_PIECE) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(xp_ready);
      if (it > 0) mbar_wait(g2_done, ph ^ 1u);                   // context GEMM of the previous tile has read X
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 a = xv[i][0], b = xv[i][1];
        uint32_t h4[4], l4[4];
        split2(a.x, a.y, h4[0], l4[0]);
        split2(a.z, a.w, h4[1], l4[1]);
        split2(b.x, b.y, h4[2], l4[2]);
        split2(b.z, b.w, h4[3], l4[3]);
        unsigned char* dst = Xs + cc * T_LBO + row_off(rb + 16 * i);
        *reinterpret_cast<uint4*>(dst) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
        *reinterpret_cast<uint4*>(dst + T_PIECE) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
      }
      if (lt < CT_TV) lab_s[lt] = (unsigned char)lab;
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(x_ready);
    }
  } else if (warp == 12) {
    // ======================================================================================= MMA issuer
    // whole warp, warp-uniform state; one elected lane issues (see elect_one in tc_common.cuh)
    {
      const uint32_t id1 = umma_idesc_bf16_major(CT_TV, 0, 0);   // S: N = 64, both operands K-major
      const uint32_t id2 = umma_idesc_bf16_major(CT_D, 0, 1);    // ctx: N = 128, B (= X tile) MN-major
      const uint32_t q_hi = smem_u32(Qs), q_lo = q_hi + Q_PIECE;
      const uint32_t xp_hi = smem_u32(XPs), xp_lo = xp_hi + T_PIECE;
      const uint32_t x_hi = smem_u32(Xs), x_lo = x_hi + T_PIECE;
      const uint32_t p_hi = smem_u32(Ps), p_lo = p_hi + P_PIECE;
      // S[t & 1] = Qf . (x+pos)^T of tile t.  It is issued one tile ahead (under the softmax of tile t - 1), so the
      // tensor pipe has the next scores ready when the softmax warps come back.
      const int n_my = n_tiles > blockIdx.x ? (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
      auto issue_scores = [&](int t) {
        mbar_wait(xp_ready, (uint32_t)t & 1u);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_base + ((t & 1) ? CT_TM_S1 : CT_TM_S);
#pragma unroll
          for (int j = 0; j < 8; ++j) {                            // K = 128 channels, 16 per step
            const uint64_t a_h = umma_desc(q_hi + j * 2 * A_LBO, A_LBO, 128), a_l = umma_desc(q_lo + j * 2 * A_LBO, A_LBO, 128);
            const uint64_t b_h = umma_desc(xp_hi + j * 2 * T_LBO, T_LBO, 128), b_l = umma_desc(xp_lo + j * 2 * T_LBO, T_LBO, 128);
            umma_bf16(d, a_h, b_h, id1, j ? 1u : 0u);
            umma_bf16(d, a_h, b_l, id1, 1u);
            umma_bf16(d, a_l, b_h, id1, 1u);
          }
          umma_commit(s_full(t));
        }
        __syncwarp();
      };
      if (n_my > 0) issue_scores(0);
      for (int it = 0; it < n_my; ++it) {
        const uint32_t ph = (uint32_t)it & 1u;
        if (it + 1 < n_my) issue_scores(it + 1);
        mbar_wait(p_ready, ph);
        mbar_wait(x_ready, ph);
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {                            // K = 64 voxels, 16 per step
            const uint64_t a_h = umma_desc(p_hi + j * 2 * A_LBO, A_LBO, 128), a_l = umma_desc(p_lo + j * 2 * A_LBO, A_LBO, 128);
            // MN-major B: K direction (voxel groups of 8) stride 128 B = "LBO", channel-chunk stride T_LBO = "SBO"
            const uint64_t b_h = umma_desc(x_hi + j * 2 * 128, 128, T_LBO), b_l = umma_desc(x_lo + j * 2 * 128, 128, T_LBO);
            umma_bf16(tmem_base + CT_TM_CTX, a_h, b_h, id2, (it | j) ? 1u : 0u);
            umma_bf16(tmem_base + CT_TM_CTX, a_h, b_l, id2, 1u);
            umma_bf16(tmem_base + CT_TM_CTX, a_l, b_h, id2, 1u);
          }
          umma_commit(g2_done);
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

size_t c2s_tc_workspace_bytes(int nq, int heads) {
  (void)heads;
  const int groups = (nq + 15) / 16;
  return (size_t)groups * (size_t)sm_count() * 128 * (CT_D + 2) * sizeof(float) + 256;
}

// launches the partial kernel; returns the geometry the merge kernel needs
int c2s_tc_launch(const float* x, const float* pos, long long nv, const float* qfold, int nq, int heads,
                  const unsigned char* label, const int* q_obj, const int* obj_count, void* ws, size_t ws_bytes,
                  cudaStream_t st, float** part_m, float** part_l, float** part_acc, int* n_cta_out, int* nqg_out) {
  AG3D_CHECK_ARG(heads == 8, "tensor-core c2s handles 8 heads");
  AG3D_CHECK_ARG(ws && aligned16(ws) && ws_bytes >= c2s_tc_workspace_bytes(nq, heads), "c2s workspace too small");
  const int groups = (nq + 15) / 16;
  const int nqg = (nq + groups - 1) / groups;
  const long long n_tiles = (nv + CT_TV - 1) / CT_TV;
  int n_cta = std::max(1, sm_count() / groups);
  if (n_cta > n_tiles) n_cta = (int)n_tiles;
  C2sParams p;
  p.x = x; p.pos = pos; p.nv = nv; p.qfold = qfold; p.nq = nq; p.heads = heads; p.nqg = nqg;
  p.label = label; p.q_obj = q_obj; p.obj_count = obj_count;
  p.part_m = static_cast<float*>(ws);
  p.part_l = p.part_m + (size_t)groups * n_cta * 128;
  p.part_acc = p.part_l + (size_t)groups * n_cta * 128;
  p.part_acc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(p.part_acc) + 15) & ~(uintptr_t)15);
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(c2s_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CT_SMEM));
    attr = true;
  }
  c2s_tc_kernel<<<dim3(n_cta, groups), CT_THREADS, CT_SMEM, st>>>(p);
  AG3D_LAUNCH_CHECK("c2s_tc");
  *part_m = p.part_m; *part_l = p.part_l; *part_acc = p.part_acc; *n_cta_out = n_cta; *nqg_out = nqg;
  return AG3D_OK;
}

}  // namespace ag3d
