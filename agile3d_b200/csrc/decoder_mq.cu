// K10 for MANY click queries (33 .. 256 per scene): scene -> click cross-attention + residual + LayerNorm + mask head
// when the (head, query) score columns of a voxel tile no longer fit TMEM at once - the tail of the reference's
// iterative-click protocol, where Nq grows to 10 + 20 K (eval_multi_obj.py:116-167, models/agile3d.py:305-321,342-384).
//
// Per 128-voxel tile the queries are walked in GROUPS of 16 (128 score columns = 8 heads x 16):
//   pass 1   S_g = (x+pos) . A_g^T (bf16x3, TMEM, double buffered)  -> per (voxel, head) running max and sum of exp
//   pass 2   S_g again -> P_g = exp(S_g - max) / sum (bf16 hi/lo, smem) -> O += P_g . U_g  (one [128 x 128] accumulator;
//            the heads mix in O, which is why the softmax statistics must be final before the first product: two
//            passes over the scores instead of an online rescale)
//   then     O + bo + x -> LayerNorm -> y (global) and Y (bf16 hi/lo, smem) -> Z = Y . E^T [128 x NQP] -> per-object max
// The queries are taken in object-sorted order (stable; computed on the device by mq_perm_kernel), so the columns of an
// object are one contiguous run of Z and the per-object maximum is a single sweep; objects without a query get -inf,
// as in the reference's torch.max over an empty set would never happen (every object has a click) and as the 32-query
// kernel does.  Operand images (A_g, U_g per group, E) stream through a ring of 16-KB stages by cp.async.bulk in
// exactly the order the MMA issuer consumes them.
#include <float.h>
#include <math.h>

#include <algorithm>
#include <cstdlib>

#include "tc_common.cuh"

namespace ag3d {

constexpr int MQ_D = 128;
constexpr int MQ_QG = 16;                 // queries per group
constexpr int MQ_HQ = 8 * MQ_QG;          // score columns per group
constexpr int MQ_MAXQ = 256;
constexpr int MQ_COMPUTE_THREADS = 256;
constexpr int MQ_THREADS = MQ_COMPUTE_THREADS + 64;
constexpr int MQ_MISC = 8192;
constexpr int MQ_NBR = 5;                 // ring slots
constexpr uint32_t MQ_STAGE = 16384;
constexpr uint32_t MQ_TM_S = 0, MQ_TM_O = 256, MQ_TM_Z = 0;     // S: two buffers of 128 columns; Z reuses them
constexpr size_t MQ_SMEM = MQ_MISC + 8 * (size_t)A_STAGE + (size_t)MQ_NBR * MQ_STAGE;

// ---- operand prep -------------------------------------------------------------------------------------------------
// perm[s] = original index of the s-th query in (object id, original order) order; obj_end[ob] = number of queries with
// object id <= ob.  One block; nq <= 256.
__global__ void mq_perm_kernel(const int* __restrict__ q_obj, int nq, int n_obj, int* __restrict__ perm,
                               int* __restrict__ obj_end) {
  __shared__ int ob_s[MQ_MAXQ];
  const int t = threadIdx.x;
  if (t < nq) ob_s[t] = q_obj[t];
  __syncthreads();
  if (t < nq) {
    const int mine = ob_s[t];
    int pos = 0;
    for (int j = 0; j < nq; ++j) pos += (ob_s[j] < mine || (ob_s[j] == mine && j < t)) ? 1 : 0;
    perm[pos] = t;
  }
  if (t < n_obj) {
    int c = 0;
    for (int j = 0; j < nq; ++j) c += ob_s[j] <= t ? 1 : 0;
    obj_end[t] = c;
  }
}

// images: A [G][4 slabs][2 pieces][4 kc][128 cols][8] | U [G][4 slabs of P columns][2][4][128 ch][8] |
//         E [4 slabs][2 pieces][4 kc][NQP][8]   (uint4 units); cpad [G][128]
__global__ void mq_prep_kernel(const float* __restrict__ A, const float* __restrict__ cvec, const float* __restrict__ U,
                               const float* __restrict__ E, const int* __restrict__ perm, int nq, int G, int NQP,
                               uint4* __restrict__ img, float* __restrict__ cpad) {
  const long long nA = (long long)G * 4 * 4 * MQ_HQ, nU = (long long)G * 4 * 4 * 128, nE = (long long)4 * 4 * NQP;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < (long long)G * MQ_HQ) {
    const int g = (int)(t / MQ_HQ), n = (int)(t % MQ_HQ), h = n / MQ_QG, sq = g * MQ_QG + n % MQ_QG;
    cpad[t] = sq < nq ? cvec[h * nq + perm[sq]] : -INFINITY;
  }
  if (t >= nA + nU + nE) return;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  size_t hi_idx, lo_idx;
  if (t < nA) {                                  // B[n = score column][k = channel]
    const int n = (int)(t % MQ_HQ), kc = (int)((t / MQ_HQ) % 4), s = (int)((t / (MQ_HQ * 4)) % 4), g = (int)(t / (MQ_HQ * 16));
    const int h = n / MQ_QG, sq = g * MQ_QG + n % MQ_QG;
    if (sq < nq) {
      const float* src = A + (size_t)(h * nq + perm[sq]) * MQ_D + s * 32 + kc * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldg(src + e);
    }
    const size_t base = ((size_t)g * 4 + s) * (MQ_STAGE / 16);
    hi_idx = base + (size_t)(0 * 4 + kc) * MQ_HQ + n;
    lo_idx = base + (size_t)(1 * 4 + kc) * MQ_HQ + n;
  } else if (t < nA + nU) {                      // B[n = channel][k = probability column]
    const long long u = t - nA;
    const int n = (int)(u % 128), kc = (int)((u / 128) % 4), s = (int)((u / 512) % 4), g = (int)(u / 2048);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = s * 32 + kc * 8 + e, h = col / MQ_QG, sq = g * MQ_QG + col % MQ_QG;
      if (sq < nq) v[e] = __ldg(U + (size_t)(h * nq + perm[sq]) * MQ_D + n);
    }
    const size_t base = (size_t)G * 4 * (MQ_STAGE / 16) + ((size_t)g * 4 + s) * (MQ_STAGE / 16);
    hi_idx = base + (size_t)(0 * 4 + kc) * 128 + n;
    lo_idx = base + (size_t)(1 * 4 + kc) * 128 + n;
  } else {                                       // B[n = sorted query][k = channel]; hi and lo pieces are separate stages
    const long long u = t - nA - nU;
    const int n = (int)(u % NQP), kc = (int)((u / NQP) % 4), s = (int)(u / (NQP * 4));
    if (n < nq) {
      const float* src = E + (size_t)perm[n] * MQ_D + s * 32 + kc * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldg(src + e);
    }
    const size_t base = (size_t)G * 8 * (MQ_STAGE / 16) + (size_t)s * 2 * (size_t)(NQP * 4);
    hi_idx = base + (size_t)kc * NQP + n;
    lo_idx = base + (size_t)(NQP * 4) + (size_t)kc * NQP + n;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
  img[hi_idx] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  img[lo_idx] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

struct MqParams {
  const float* x; const float* pos; long long nv;
  const uint4* img; const float* cpad; const int* obj_end;
  const float* bo; const float* ln_w; const float* ln_b; float ln_eps;
  int nq, n_obj, G, NQP;
  float* x_out; float* logits; unsigned char* label; int* obj_count;
};

__global__ void __launch_bounds__(MQ_THREADS, 1) s2c_mq_kernel(const MqParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 192);
  int* hist_s = reinterpret_cast<int*>(smem + 256);              // [32]
  int* oend_s = reinterpret_cast<int*>(smem + 384);              // [32]
  float* vec_s = reinterpret_cast<float*>(smem + 512);           // bo | ln_w | ln_b
  float* lnred_s = reinterpret_cast<float*>(smem + 3072);        // [2][128][2]
  unsigned char* XP = smem + MQ_MISC;                            // (x+pos) tile, later the Y tile: 4 slabs
  unsigned char* PT = XP + 4 * (size_t)A_STAGE;                  // probability tile of one group: 4 slabs
  unsigned char* ring = PT + 4 * (size_t)A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t xp_full = bar_base, p_full = bar_base + 8, p_free = bar_base + 16, o_full = bar_base + 24,
                 y_full = bar_base + 32, z_full = bar_base + 40;
  auto s_full = [&](int b) { return bar_base + 8u * (6 + b); };
  auto s_free = [&](int b) { return bar_base + 8u * (8 + b); };
  auto b_full = [&](int s) { return bar_base + 8u * (10 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (16 + s); };

  if (tid == 0) {
    mbar_init(xp_full, 8); mbar_init(p_full, 8); mbar_init(y_full, 8);
    mbar_init(p_free, 1); mbar_init(o_full, 1); mbar_init(z_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(s_full(b), 1); mbar_init(s_free(b), 8); }
    for (int s = 0; s < MQ_NBR; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) { hist_s[tid] = 0; oend_s[tid] = tid < p.n_obj ? p.obj_end[tid] : p.nq; }
  for (int i = tid; i < 128; i += MQ_THREADS) { vec_s[i] = p.bo[i]; vec_s[128 + i] = p.ln_w[i]; vec_s[256 + i] = p.ln_b[i]; }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long n_tiles = (p.nv + TC_BM - 1) / TC_BM;
  const int G = p.G;

  if (warp < 8) {
    // ======================================================================================= compute warps
    const int q4 = warp & 3, g2 = warp >> 2;
    const int r = q4 * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int ld_kc = tid & 3, ld_rb = tid >> 2;
    uint32_t ns = 0, np = 0;
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const long long row0 = tile * TC_BM;
      // ---- P0: x + pos -> bf16 hi/lo slabs.  The previous tile's Z GEMM (which read this region as Y) is complete:
      //      z_full was waited for at the end of the previous iteration.
#pragma unroll
      for (int sp = 0; sp < 2; ++sp) {
        float4 xv[2][2][2], pv[2][2][2];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const long long row = row0 + ld_rb + 64 * i;
            const size_t off = (size_t)row * MQ_D + (sp * 2 + s2) * 32 + ld_kc * 8;
#pragma unroll
            for (int hlf = 0; hlf < 2; ++hlf) {
              xv[s2][i][hlf] = pv[s2][i][hlf] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < p.nv) {
                xv[s2][i][hlf] = *reinterpret_cast<const float4*>(p.x + off + hlf * 4);
                pv[s2][i][hlf] = __ldg(reinterpret_cast<const float4*>(p.pos + off + hlf * 4));
              }
            }
          }
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float4 a = xv[s2][i][0], b = xv[s2][i][1], c = pv[s2][i][0], d = pv[s2][i][1];
            uint32_t h[4], l[4];
            split2(a.x + c.x, a.y + c.y, h[0], l[0]);
            split2(a.z + c.z, a.w + c.w, h[1], l[1]);
            split2(b.x + d.x, b.y + d.y, h[2], l[2]);
            split2(b.z + d.z, b.w + d.w, h[3], l[3]);
            unsigned char* dst = XP + (size_t)(sp * 2 + s2) * A_STAGE + a_piece_off(ld_rb + 64 * i, ld_kc);
            *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(dst + A_PIECE) = make_uint4(l[0], l[1], l[2], l[3]);
          }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(xp_full);

      // ---- pass 1: softmax statistics of this thread's row for heads 4 g2 .. 4 g2 + 3
      float mx[4], sm[4];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) { mx[hh] = -INFINITY; sm[hh] = 0.f; }
      for (int g = 0; g < G; ++g, ++ns) {
        const int b = (int)(ns & 1u);
        mbar_wait(s_full(b), (ns >> 1) & 1u);
        tc_fence_after();
        const float* cp = p.cpad + (size_t)g * MQ_HQ;
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          const int col0 = (4 * g2 + hh) * MQ_QG;
          float sc[MQ_QG];
          tmem_ld16(t_lane + MQ_TM_S + (uint32_t)(b * MQ_HQ + col0), sc);
          float gm = -INFINITY;
#pragma unroll
          for (int i = 0; i < MQ_QG; ++i) { sc[i] += __ldg(cp + col0 + i); gm = fmaxf(gm, sc[i]); }
          const float m_new = fmaxf(mx[hh], gm);          // group 0 always holds a real query: m_new is finite
          float add = 0.f;
#pragma unroll
          for (int i = 0; i < MQ_QG; ++i) add += __expf(sc[i] - m_new);
          sm[hh] = sm[hh] * __expf(mx[hh] - m_new) + add;
          mx[hh] = m_new;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free(b));
      }
      float inv[4];
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) inv[hh] = 1.f / sm[hh];

      // ---- pass 2: probabilities of group g -> smem (A operand of O += P_g . U_g)
      for (int g = 0; g < G; ++g, ++ns, ++np) {
        const int b = (int)(ns & 1u);
        mbar_wait(s_full(b), (ns >> 1) & 1u);
        tc_fence_after();
        const float* cp = p.cpad + (size_t)g * MQ_HQ;
        float pr[4][MQ_QG];
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          const int col0 = (4 * g2 + hh) * MQ_QG;
          tmem_ld16(t_lane + MQ_TM_S + (uint32_t)(b * MQ_HQ + col0), pr[hh]);
#pragma unroll
          for (int i = 0; i < MQ_QG; ++i) pr[hh][i] = __expf(pr[hh][i] + __ldg(cp + col0 + i) - mx[hh]) * inv[hh];
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free(b));
        mbar_wait(p_free, (np & 1u) ^ 1u);                 // the O GEMM of the previous group has read the P tile
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          const int col0 = (4 * g2 + hh) * MQ_QG;
#pragma unroll
          for (int c8 = 0; c8 < MQ_QG / 8; ++c8) {
            uint32_t h[4], l[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) split2(pr[hh][c8 * 8 + 2 * e], pr[hh][c8 * 8 + 2 * e + 1], h[e], l[e]);
            const int col = col0 + c8 * 8;
            unsigned char* dst = PT + (size_t)(col >> 5) * A_STAGE + a_piece_off(r, (col >> 3) & 3);
            *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(dst + A_PIECE) = make_uint4(l[0], l[1], l[2], l[3]);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }

      // ---- o + bo + x -> LayerNorm -> y; this thread: row r, channels 64 g2 .. 64 g2 + 63
      mbar_wait(o_full, ph);
      tc_fence_after();
      const long long row = row0 + r;
      const bool valid = row < p.nv;
      float v[64];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_ld16(t_lane + MQ_TM_O + 64 * g2 + ch * 16, v + ch * 16);
      float sum = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        float4 xr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) xr = *reinterpret_cast<const float4*>(p.x + (size_t)row * MQ_D + 64 * g2 + c4 * 4);
        const float4 b4 = *reinterpret_cast<const float4*>(vec_s + 64 * g2 + c4 * 4);
        v[c4 * 4 + 0] = xr.x + (v[c4 * 4 + 0] + b4.x);
        v[c4 * 4 + 1] = xr.y + (v[c4 * 4 + 1] + b4.y);
        v[c4 * 4 + 2] = xr.z + (v[c4 * 4 + 2] + b4.z);
        v[c4 * 4 + 3] = xr.w + (v[c4 * 4 + 3] + b4.w);
        sum += (v[c4 * 4 + 0] + v[c4 * 4 + 1]) + (v[c4 * 4 + 2] + v[c4 * 4 + 3]);
      }
      lnred_s[(0 * 128 + r) * 2 + g2] = sum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float mean = (lnred_s[(0 * 128 + r) * 2 + 0] + lnred_s[(0 * 128 + r) * 2 + 1]) * (1.f / MQ_D);
      float sq = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) { const float d = v[c] - mean; sq = fmaf(d, d, sq); }
      lnred_s[(1 * 128 + r) * 2 + g2] = sq;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float var = (lnred_s[(1 * 128 + r) * 2 + 0] + lnred_s[(1 * 128 + r) * 2 + 1]) * (1.f / MQ_D);
      const float rstd = 1.f / sqrtf(var + p.ln_eps);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = 64 * g2 + c8 * 8 + e;
          y[e] = (v[c8 * 8 + e] - mean) * rstd * vec_s[128 + c] + vec_s[256 + c];
        }
        if (valid) {
          float* o = p.x_out + (size_t)row * MQ_D + 64 * g2 + c8 * 8;
          *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2(y[2 * e], y[2 * e + 1], h[e], l[e]);
        const int c = 64 * g2 + c8 * 8;
        unsigned char* dst = XP + (size_t)(c >> 5) * A_STAGE + a_piece_off(r, (c >> 3) & 3);
        *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(dst + A_PIECE) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(y_full);

      // ---- mask head: Z columns are the object-sorted queries, object ob owns columns [obj_end[ob-1], obj_end[ob])
      mbar_wait(z_full, ph);
      tc_fence_after();
      if (g2 == 0) {
        float best = -INFINITY, cur = -INFINITY;
        int arg = 0, ob = 0, run_end = oend_s[0];
        for (int c0 = 0; c0 < p.NQP; c0 += 16) {
          float z[16];
          tmem_ld16(t_lane + MQ_TM_Z + (uint32_t)c0, z);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = c0 + i;
            while (ob < p.n_obj && col >= run_end) {         // close the runs that end before this column
              if (valid) p.logits[(size_t)row * p.n_obj + ob] = cur;
              if (cur > best || ob == 0) { best = cur; arg = ob; }
              cur = -INFINITY;
              ++ob;
              run_end = ob < p.n_obj ? oend_s[ob] : 0x7fffffff;
            }
            if (col < p.nq) cur = fmaxf(cur, z[i]);
          }
        }
        while (ob < p.n_obj) {
          if (valid) p.logits[(size_t)row * p.n_obj + ob] = cur;
          if (cur > best || ob == 0) { best = cur; arg = ob; }
          cur = -INFINITY;
          ++ob;
        }
        if (valid) {
          p.label[row] = (unsigned char)arg;
          atomicAdd(&hist_s[arg], 1);
        }
      }
      tc_fence_before();
    }
  } else if (warp == 8) {
    // ======================================================================================= MMA issuer
    const uint32_t id_s = umma_idesc_bf16(MQ_HQ), id_o = umma_idesc_bf16(128), id_z = umma_idesc_bf16(p.NQP);
    const uint32_t d_hi32 = umma_desc_hi32(128);
    const uint32_t xp_lo32 = umma_desc_lo32(smem_u32(XP), A_LBO), pt_lo32 = umma_desc_lo32(smem_u32(PT), A_LBO);
    const uint32_t ring_u32 = smem_u32(ring);
    uint32_t nb = 0, ns = 0, np = 0;
    // one [128 x 32-channel slab] x B stage product: 2 k-steps x 3 products.  b_hi / b_lo: descriptor words of the hi and
    // lo pieces of the B stage
    auto slab = [&](uint32_t a_lo32, int s, uint32_t d, uint32_t idesc, uint32_t b_lbo, uint32_t b_hi, uint32_t b_lo, bool first) {
      const uint32_t ah = a_lo32 + (uint32_t)s * (uint32_t)(A_STAGE >> 4);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const uint64_t da_hi = umma_desc_join(d_hi32, ah + ks * ((2 * A_LBO) >> 4));
        const uint64_t da_lo = umma_desc_join(d_hi32, ah + ks * ((2 * A_LBO) >> 4) + (A_PIECE >> 4));
        const uint64_t db_hi = umma_desc_join(d_hi32, b_hi + ks * ((2u * b_lbo) >> 4));
        const uint64_t db_lo = umma_desc_join(d_hi32, b_lo + ks * ((2u * b_lbo) >> 4));
        umma_bf16(d, da_hi, db_hi, idesc, (first && ks == 0) ? 0u : 1u);
        umma_bf16(d, da_hi, db_lo, idesc, 1u);
        umma_bf16(d, da_lo, db_hi, idesc, 1u);
      }
    };
    // GEMM over 4 slabs whose B stages hold [hi piece | lo piece] (A_g and U_g images: one ring slot per slab)
    auto gemm4 = [&](uint32_t a_lo32, uint32_t d, uint32_t idesc, uint32_t b_lbo, bool accumulate) {
      for (int s = 0; s < 4; ++s, ++nb) {
        const int sb = (int)(nb % MQ_NBR);
        mbar_wait(b_full(sb), (nb / MQ_NBR) & 1u);
        tc_fence_after();
        const uint32_t b_hi = umma_desc_lo32(ring_u32 + (uint32_t)sb * MQ_STAGE, b_lbo);
        if (elect_one()) {
          slab(a_lo32, s, d, idesc, b_lbo, b_hi, b_hi + ((4u * b_lbo) >> 4), !accumulate && s == 0);
          umma_commit(b_empty(sb));
        }
        __syncwarp();
      }
    };
    auto issue_s = [&]() {
      const int b = (int)(ns & 1u);
      mbar_wait(s_free(b), ((ns >> 1) & 1u) ^ 1u);
      tc_fence_after();
      gemm4(xp_lo32, tmem_base + MQ_TM_S + (uint32_t)(b * MQ_HQ), id_s, (uint32_t)MQ_HQ * 16u, false);
      if (elect_one()) umma_commit(s_full(b));
      __syncwarp();
      ++ns;
    };
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      mbar_wait(xp_full, ph);
      tc_fence_after();
      for (int g = 0; g < G; ++g) issue_s();                 // pass 1 (two score buffers: one group ahead of the readers)
      issue_s();                                             // pass 2, group 0
      for (int g = 0; g < G; ++g, ++np) {
        if (g + 1 < G) issue_s();
        mbar_wait(p_full, np & 1u);
        tc_fence_after();
        gemm4(pt_lo32, tmem_base + MQ_TM_O, id_o, 128u * 16u, g > 0);
        if (elect_one()) umma_commit(p_free);
        __syncwarp();
      }
      if (elect_one()) umma_commit(o_full);
      __syncwarp();
      mbar_wait(y_full, ph);
      tc_fence_after();
      const uint32_t e_lbo = (uint32_t)p.NQP * 16u;
      for (int s = 0; s < 4; ++s, nb += 2) {                 // E slabs: hi and lo pieces in consecutive ring slots
        const int sb = (int)(nb % MQ_NBR), sb2 = (int)((nb + 1) % MQ_NBR);
        mbar_wait(b_full(sb), (nb / MQ_NBR) & 1u);
        mbar_wait(b_full(sb2), ((nb + 1) / MQ_NBR) & 1u);
        tc_fence_after();
        if (elect_one()) {
          slab(xp_lo32, s, tmem_base + MQ_TM_Z, id_z, e_lbo, umma_desc_lo32(ring_u32 + (uint32_t)sb * MQ_STAGE, e_lbo),
               umma_desc_lo32(ring_u32 + (uint32_t)sb2 * MQ_STAGE, e_lbo), s == 0);
          umma_commit(b_empty(sb));
          umma_commit(b_empty(sb2));
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(z_full);
      __syncwarp();
    }
  } else {
    // ======================================================================================= operand loader
    const unsigned char* img = reinterpret_cast<const unsigned char*>(p.img);
    const size_t u_base = (size_t)G * 4 * MQ_STAGE, e_base = (size_t)G * 8 * MQ_STAGE;
    const uint32_t e_bytes = (uint32_t)p.NQP * 64u;
    uint32_t nb = 0;
    auto push = [&](const unsigned char* src, uint32_t bytes) {
      const int sb = (int)(nb % MQ_NBR);
      mbar_wait(b_empty(sb), ((nb / MQ_NBR) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(b_full(sb), bytes);
        bulk_g2s(smem_u32(ring + (size_t)sb * MQ_STAGE), src, bytes, b_full(sb));
      }
      __syncwarp();
      ++nb;
    };
    auto push_a = [&](int g) { for (int s = 0; s < 4; ++s) push(img + ((size_t)g * 4 + s) * MQ_STAGE, MQ_STAGE); };
    auto push_u = [&](int g) { for (int s = 0; s < 4; ++s) push(img + u_base + ((size_t)g * 4 + s) * MQ_STAGE, MQ_STAGE); };
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int g = 0; g < G; ++g) push_a(g);
      push_a(0);
      for (int g = 0; g < G; ++g) {
        if (g + 1 < G) push_a(g + 1);
        push_u(g);
      }
      for (int s = 0; s < 8; ++s) push(img + e_base + (size_t)s * e_bytes, e_bytes);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tid < p.n_obj && hist_s[tid]) atomicAdd(p.obj_count + tid, hist_s[tid]);
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

static int mq_groups(int nq) { return (nq + MQ_QG - 1) / MQ_QG; }

size_t s2c_mq_workspace_bytes(int nq) {
  const int G = mq_groups(nq), NQP = G * MQ_QG;
  return (size_t)G * 8 * MQ_STAGE + 8 * (size_t)NQP * 64 + (size_t)G * MQ_HQ * 4 + MQ_MAXQ * 4 + 32 * 4 + 256;
}

int s2c_mq_launch(const float* x, const float* pos, long long nv, const float* A, const float* c, const float* U,
                  const float* bo, const float* ln_w, const float* ln_b, float ln_eps, const float* E,
                  const int* q_obj, int nq, int heads, int n_obj, float* x_out, float* logits, unsigned char* label,
                  int* obj_count, void* ws, size_t ws_bytes, cudaStream_t st) {
  AG3D_CHECK_ARG(heads == 8 && nq >= 1 && nq <= MQ_MAXQ, "the many-query s2c kernel handles 8 heads and at most 256 queries");
  AG3D_CHECK_ARG(ws && aligned16(ws) && ws_bytes >= s2c_mq_workspace_bytes(nq), "s2c workspace too small");
  const int G = mq_groups(nq), NQP = G * MQ_QG;
  unsigned char* w = static_cast<unsigned char*>(ws);
  uint4* img = reinterpret_cast<uint4*>(w);
  size_t off = (size_t)G * 8 * MQ_STAGE + 8 * (size_t)NQP * 64;
  float* cpad = reinterpret_cast<float*>(w + off);
  off += (size_t)G * MQ_HQ * 4;
  int* perm = reinterpret_cast<int*>(w + off);
  off += MQ_MAXQ * 4;
  int* obj_end = reinterpret_cast<int*>(w + off);
  mq_perm_kernel<<<1, MQ_MAXQ, 0, st>>>(q_obj, nq, n_obj, perm, obj_end);
  AG3D_LAUNCH_CHECK("mq_perm");
  const long long total = (long long)G * 16 * MQ_HQ + (long long)G * 16 * 128 + 16LL * NQP;
  mq_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(A, c, U, E, perm, nq, G, NQP, img, cpad);
  AG3D_LAUNCH_CHECK("mq_prep");
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(s2c_mq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MQ_SMEM));
    attr = true;
  }
  MqParams p;
  p.x = x; p.pos = pos; p.nv = nv; p.img = img; p.cpad = cpad; p.obj_end = obj_end; p.bo = bo; p.ln_w = ln_w;
  p.ln_b = ln_b; p.ln_eps = ln_eps; p.nq = nq; p.n_obj = n_obj; p.G = G; p.NQP = NQP; p.x_out = x_out;
  p.logits = logits; p.label = label; p.obj_count = obj_count;
  const long long tiles = (nv + TC_BM - 1) / TC_BM;
  const int grid = (int)std::min<long long>(tiles, sm_count());
  s2c_mq_kernel<<<grid, MQ_THREADS, MQ_SMEM, st>>>(p);
  AG3D_LAUNCH_CHECK("s2c_mq");
  return AG3D_OK;
}

}  // namespace ag3d
