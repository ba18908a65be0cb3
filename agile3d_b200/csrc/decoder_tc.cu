// K10 (tensor-core variant): scene -> click cross-attention + residual + LayerNorm + mask head, one pass over the
// voxels, three chained tcgen05 GEMMs per 128-voxel tile with fp32 accumulators in TMEM:
//   S = (x+pos) . A^T            [128 x HQP]   (bf16x3)      -> per-head softmax in registers -> P (bf16 hi/lo, smem)
//   O = P . U                    [128 x 128]   (bf16x3)      -> + bo + x -> LayerNorm -> y (global) and Y (bf16 hi/lo, smem)
//   Z = Y . E^T                  [128 x 32]    (bf16x3)      -> per-object max -> logits, label, histogram
// The three small right-hand operands (A, U^T, E; <= 176 KB as bf16 hi/lo images, L2 resident) are streamed per
// tile through a ring of shared-memory stages by the TMA engine (cp.async.bulk); the voxel tile itself is read
// from HBM exactly once.  Column layout of S/P: head-major with every head padded to NQ16 (16, 24 or 32) columns so
// that a 16-column TMEM load never straddles heads; padded columns carry a bias of -inf (probability 0).
// Roles: 8 compute warps (tile load + split, softmax, LayerNorm, mask head), 1 MMA-issuer thread, 1 loader thread.
#include <cuda.h>
#include <float.h>
#include <math.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace ag3d {

constexpr int DT_D = 128;
constexpr int DT_COMPUTE_THREADS = 256;
constexpr int DT_THREADS = DT_COMPUTE_THREADS + 64;
constexpr int DT_NQP = 32;               // mask-head columns (queries padded to 32)
constexpr int DT_MISC_BYTES = 8192;      // barriers + small per-CTA arrays
constexpr uint32_t TM_S = 0, TM_O = 256, TM_Z = 384;   // TMEM column offsets

// ---------------------------------------------------------------------------------------------- operand prep
// Builds the bf16 hi/lo shared-memory images of the right-hand operands (consumption order G1 | G2 | G3) and the
// padded score bias.  A, U: [heads*nq, 128] fp32 (row = h*nq + q); E: [nq, 128]; c: [heads*nq].
template <int NQ16>
__global__ void s2c_prep_kernel(const float* __restrict__ A, const float* __restrict__ cvec,
                                const float* __restrict__ U, const float* __restrict__ E, int nq, int heads,
                                uint4* __restrict__ img, float* __restrict__ cpad) {
  constexpr int HQP = 8 * NQ16;
  constexpr int SLABS_P = HQP / 32;
  constexpr int N1 = 4 * 4 * HQP, N2 = SLABS_P * 4 * 128, N3 = 4 * 4 * DT_NQP;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < HQP) {
    const int h = t / NQ16, qq = t % NQ16;
    cpad[t] = (qq < nq && h < heads) ? cvec[h * nq + qq] : -INFINITY;
  }
  if (t >= N1 + N2 + N3) return;
  float v[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = 0.f;
  size_t hi_idx, lo_idx;
  if (t < N1) {                                   // G1: B[n = padded column][k = channel]
    const int n = t % HQP, kc = (t / HQP) % 4, s = t / (HQP * 4);
    const int h = n / NQ16, qq = n % NQ16;
    if (qq < nq) {
      const float* src = A + (size_t)(h * nq + qq) * DT_D + s * 32 + kc * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldg(src + e);
    }
    const size_t base = (size_t)s * (HQP * 8);
    hi_idx = base + (size_t)(0 * 4 + kc) * HQP + n;
    lo_idx = base + (size_t)(1 * 4 + kc) * HQP + n;
  } else if (t < N1 + N2) {                       // G2: B[n = channel][k = padded column]
    const int u = t - N1;
    const int n = u % 128, kc = (u / 128) % 4, s = u / 512;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = s * 32 + kc * 8 + e;
      const int h = col / NQ16, qq = col % NQ16;
      if (qq < nq) v[e] = __ldg(U + (size_t)(h * nq + qq) * DT_D + n);
    }
    const size_t base = (size_t)4 * HQP * 8 + (size_t)s * 1024;
    hi_idx = base + (size_t)(0 * 4 + kc) * 128 + n;
    lo_idx = base + (size_t)(1 * 4 + kc) * 128 + n;
  } else {                                        // G3: B[n = query][k = channel]
    const int u = t - N1 - N2;
    const int n = u % DT_NQP, kc = (u / DT_NQP) % 4, s = u / (DT_NQP * 4);
    if (n < nq) {
      const float* src = E + (size_t)n * DT_D + s * 32 + kc * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = __ldg(src + e);
    }
    const size_t base = (size_t)4 * HQP * 8 + (size_t)SLABS_P * 1024 + (size_t)s * 256;
    hi_idx = base + (size_t)(0 * 4 + kc) * DT_NQP + n;
    lo_idx = base + (size_t)(1 * 4 + kc) * DT_NQP + n;
  }
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split2(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
  img[hi_idx] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  img[lo_idx] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

template <int NQ16>
constexpr size_t s2c_img_bytes() {
  constexpr int HQP = 8 * NQ16;
  return (size_t)(4 * HQP * 8 + (HQP / 32) * 1024 + 4 * 256) * 16;
}

// ---------------------------------------------------------------------------------------------- main kernel
struct S2cParams {
  const float* x; const float* pos; long long nv;
  const uint4* img; const float* cpad;
  const float* bo; const float* ln_w; const float* ln_b; float ln_eps;
  const int* q_obj; int nq; int n_obj;
  float* x_out; float* logits; unsigned char* label; int* obj_count;
};

template <int NQ16>
struct S2cCfg {
  static constexpr int HQP = 8 * NQ16;
  static constexpr int SLABS_P = HQP / 32;
  static constexpr int R_SLABS = SLABS_P > 4 ? SLABS_P : 4;
  // operand ring: 16-KB slots.  A G1 slab image (hi piece | lo piece, HQP * 128 bytes) of the 32-query variant is two
  // slots (hi, lo), so four slots are in flight instead of two 32-KB stages: the ring round trip, not the MMAs, is
  // what a tile waits for (the three GEMMs stream 17 slots per tile)
  static constexpr int B_STAGE = 16384;
  static constexpr int G1_CHUNKS = HQP * 128 > 16384 ? 2 : 1;          // slots per G1 slab
  static constexpr int G1_BYTES = HQP * 128 / G1_CHUNKS;               // bytes per G1 slot
  static constexpr int NBR = (HQP <= 128) ? 6 : 4;
  static constexpr size_t SMEM = DT_MISC_BYTES + (size_t)R_SLABS * A_STAGE + (size_t)NBR * B_STAGE;
};

template <int NQ16>
__global__ void __launch_bounds__(DT_THREADS, 1) s2c_tc_kernel(const S2cParams p) {
  using Cfg = S2cCfg<NQ16>;
  constexpr int HQP = Cfg::HQP, SLABS_P = Cfg::SLABS_P, NBR = Cfg::NBR;
  constexpr uint32_t B_STAGE = Cfg::B_STAGE;
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);            // [0..5] phase barriers, [8..15] ring full, [16..23] ring empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 192);
  int* hist_s = reinterpret_cast<int*>(smem + 256);              // [32]
  int* qobj_s = reinterpret_cast<int*>(smem + 384);              // [32]
  float* vec_s = reinterpret_cast<float*>(smem + 512);           // bo[128] | ln_w[128] | ln_b[128]
  float* cpad_s = reinterpret_cast<float*>(smem + 2048);         // [HQP <= 256]
  float* lnred_s = reinterpret_cast<float*>(smem + 3072);        // [2][128][2]  (sum / sq-dev partials per column half)
  unsigned char* R = smem + DT_MISC_BYTES;                       // XP -> P -> Y operand tiles (bf16 hi/lo slabs)
  unsigned char* ring = R + (size_t)Cfg::R_SLABS * A_STAGE;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t xp_full = bar_base, s_full = bar_base + 8, p_full = bar_base + 16, o_full = bar_base + 24,
                 y_full = bar_base + 32, z_full = bar_base + 40;
  auto b_full = [&](int s) { return bar_base + 8u * (8 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (16 + s); };   // up to 8 ring slots: barriers 8..15 / 16..23

  if (tid == 0) {
    mbar_init(xp_full, 8); mbar_init(p_full, 8); mbar_init(y_full, 8);
    mbar_init(s_full, 1); mbar_init(o_full, 1); mbar_init(z_full, 1);
    for (int s = 0; s < NBR; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) { hist_s[tid] = 0; qobj_s[tid] = (tid < p.nq) ? p.q_obj[tid] : -1; }
  for (int i = tid; i < 128; i += DT_THREADS) {
    vec_s[i] = p.bo[i]; vec_s[128 + i] = p.ln_w[i]; vec_s[256 + i] = p.ln_b[i];
  }
  for (int i = tid; i < HQP; i += DT_THREADS) cpad_s[i] = p.cpad[i];
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long n_tiles = (p.nv + TC_BM - 1) / TC_BM;

  if (warp < 8) {
    // ======================================================================================= compute warps
    const int q4 = warp & 3, g = warp >> 2;
    const int r = q4 * 32 + lane;                       // tile row owned in the TMEM phases
    const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    const int ld_kc = tid & 3, ld_rb = tid >> 2;        // tile-load mapping: 8 channels x rows ld_rb, ld_rb + 64
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const long long row0 = tile * TC_BM;
      // ---- P0: x + pos -> bf16 hi/lo, four 32-channel slabs (UMMA K-major A operand)
#pragma unroll
      for (int sp = 0; sp < 2; ++sp) {
        float4 xv[2][2][2], pv[2][2][2];
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const long long row = row0 + ld_rb + 64 * i;
            const size_t off = (size_t)row * DT_D + (sp * 2 + s2) * 32 + ld_kc * 8;
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
            unsigned char* dst = R + (size_t)(sp * 2 + s2) * A_STAGE + a_piece_off(ld_rb + 64 * i, ld_kc);
            *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(dst + A_PIECE) = make_uint4(l[0], l[1], l[2], l[3]);
          }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(xp_full);

      // ---- P1: per-head softmax over the queries; this thread: row r, heads 4g .. 4g+3
      mbar_wait(s_full, ph);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 4; ++hh) {
        const int col0 = (4 * g + hh) * NQ16;
        float sc[NQ16];
#pragma unroll
        if constexpr (NQ16 % 16 == 0) {
#pragma unroll
          for (int ch = 0; ch < NQ16 / 16; ++ch) tmem_ld16(t_lane + TM_S + col0 + ch * 16, sc + ch * 16);
        } else {
#pragma unroll
          for (int ch = 0; ch < NQ16 / 8; ++ch) tmem_ld8(t_lane + TM_S + col0 + ch * 8, sc + ch * 8);
        }
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < NQ16; ++i) { sc[i] += cpad_s[col0 + i]; mx = fmaxf(mx, sc[i]); }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < NQ16; ++i) { sc[i] = __expf(sc[i] - mx); sum += sc[i]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int c8 = 0; c8 < NQ16 / 8; ++c8) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split2(sc[c8 * 8 + 2 * e] * inv, sc[c8 * 8 + 2 * e + 1] * inv, h[e], l[e]);
          const int col = col0 + c8 * 8;
          unsigned char* dst = R + (size_t)(col >> 5) * A_STAGE + a_piece_off(r, (col >> 3) & 3);
          *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(dst + A_PIECE) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);

      // ---- P2: o + bo + x -> LayerNorm -> y; this thread: row r, channels 64g .. 64g+63
      mbar_wait(o_full, ph);
      tc_fence_after();
      const long long row = row0 + r;
      const bool valid = row < p.nv;
      float v[64];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) tmem_ld16(t_lane + TM_O + 64 * g + ch * 16, v + ch * 16);
      float sum = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        float4 xr = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) xr = *reinterpret_cast<const float4*>(p.x + (size_t)row * DT_D + 64 * g + c4 * 4);
        const float4 b4 = *reinterpret_cast<const float4*>(vec_s + 64 * g + c4 * 4);
        v[c4 * 4 + 0] = xr.x + (v[c4 * 4 + 0] + b4.x);
        v[c4 * 4 + 1] = xr.y + (v[c4 * 4 + 1] + b4.y);
        v[c4 * 4 + 2] = xr.z + (v[c4 * 4 + 2] + b4.z);
        v[c4 * 4 + 3] = xr.w + (v[c4 * 4 + 3] + b4.w);
        sum += (v[c4 * 4 + 0] + v[c4 * 4 + 1]) + (v[c4 * 4 + 2] + v[c4 * 4 + 3]);
      }
      lnred_s[(0 * 128 + r) * 2 + g] = sum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float mean = (lnred_s[(0 * 128 + r) * 2 + 0] + lnred_s[(0 * 128 + r) * 2 + 1]) * (1.f / DT_D);
      float sq = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) { const float d = v[c] - mean; sq = fmaf(d, d, sq); }
      lnred_s[(1 * 128 + r) * 2 + g] = sq;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float var = (lnred_s[(1 * 128 + r) * 2 + 0] + lnred_s[(1 * 128 + r) * 2 + 1]) * (1.f / DT_D);
      const float rstd = 1.f / sqrtf(var + p.ln_eps);
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int c = 64 * g + c8 * 8 + e;
          y[e] = (v[c8 * 8 + e] - mean) * rstd * vec_s[128 + c] + vec_s[256 + c];
        }
        if (valid) {
          float* o = p.x_out + (size_t)row * DT_D + 64 * g + c8 * 8;
          *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
        }
        uint32_t h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split2(y[2 * e], y[2 * e + 1], h[e], l[e]);
        const int c = 64 * g + c8 * 8;
        unsigned char* dst = R + (size_t)(c >> 5) * A_STAGE + a_piece_off(r, (c >> 3) & 3);
        *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(dst + A_PIECE) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(y_full);

      // ---- P3: mask head (column half 0 only: 32 query columns)
      mbar_wait(z_full, ph);
      tc_fence_after();
      if (g == 0) {
        float z[DT_NQP];
        tmem_ld16(t_lane + TM_Z, z);
        tmem_ld16(t_lane + TM_Z + 16, z + 16);
        if (valid) {
          float best = -INFINITY;
          int arg = 0;
          for (int ob = 0; ob < p.n_obj; ++ob) {
            float m = -INFINITY;
#pragma unroll
            for (int q = 0; q < DT_NQP; ++q) m = (qobj_s[q] == ob) ? fmaxf(m, z[q]) : m;
            p.logits[(size_t)row * p.n_obj + ob] = m;
            if (m > best || ob == 0) { best = m; arg = ob; }
          }
          p.label[row] = (unsigned char)arg;
          atomicAdd(&hist_s[arg], 1);
        }
      }
      tc_fence_before();
    }
  } else if (warp == 8) {
    // ======================================================================================= MMA issuer
    // whole warp, warp-uniform state; one elected lane issues (see elect_one in tc_common.cuh)
    {
      const uint32_t id_s = umma_idesc_bf16(HQP), id_o = umma_idesc_bf16(128), id_z = umma_idesc_bf16(DT_NQP);
      const uint32_t d_hi32 = umma_desc_hi32(128);
      const uint32_t r_lo32 = umma_desc_lo32(smem_u32(R), A_LBO);
      const uint32_t ring_u32 = smem_u32(ring);
      int nb = 0, it = 0;
      // one GEMM = `slabs` x (2 k-steps x 3 products).  chunks = ring slots per slab (2: hi piece and lo piece in
      // separate slots, G1 of the 32-query variant; 1: whole slab image in one slot; 0: all slabs in ONE slot, G3)
      auto gemm = [&](int slabs, uint32_t d, uint32_t idesc, uint32_t b_lbo, int chunks, uint32_t b_slab) {
        uint32_t b_hi = 0, b_lo = 0;
        for (int s = 0; s < slabs; ++s) {
          if (chunks || s == 0) {
            const int sb = nb % NBR;
            mbar_wait(b_full(sb), (uint32_t)(nb / NBR) & 1u);
            b_hi = umma_desc_lo32(ring_u32 + (uint32_t)sb * B_STAGE, b_lbo);
            b_lo = b_hi + ((4u * b_lbo) >> 4);
            if (chunks == 2) {
              const int sb2 = (nb + 1) % NBR;
              mbar_wait(b_full(sb2), (uint32_t)((nb + 1) / NBR) & 1u);
              b_lo = umma_desc_lo32(ring_u32 + (uint32_t)sb2 * B_STAGE, b_lbo);
            }
            tc_fence_after();
          }
          const uint32_t off = chunks ? 0u : ((uint32_t)s * b_slab) >> 4;
          const uint32_t ah = r_lo32 + (uint32_t)s * (uint32_t)(A_STAGE >> 4);
          const bool release = chunks || s == slabs - 1;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t da_hi = umma_desc_join(d_hi32, ah + ks * ((2 * A_LBO) >> 4));
              const uint64_t da_lo = umma_desc_join(d_hi32, ah + ks * ((2 * A_LBO) >> 4) + (A_PIECE >> 4));
              const uint64_t db_hi = umma_desc_join(d_hi32, b_hi + off + ks * ((2u * b_lbo) >> 4));
              const uint64_t db_lo = umma_desc_join(d_hi32, b_lo + off + ks * ((2u * b_lbo) >> 4));
              umma_bf16(d, da_hi, db_hi, idesc, (s | ks) ? 1u : 0u);
              umma_bf16(d, da_hi, db_lo, idesc, 1u);
              umma_bf16(d, da_lo, db_hi, idesc, 1u);
            }
            if (release) {
              umma_commit(b_empty(nb % NBR));
              if (chunks == 2) umma_commit(b_empty((nb + 1) % NBR));
            }
          }
          __syncwarp();
          if (release) nb += chunks == 2 ? 2 : 1;
        }
      };
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint32_t ph = (uint32_t)it & 1u;
        mbar_wait(xp_full, ph);
        tc_fence_after();
        gemm(4, tmem_base + TM_S, id_s, (uint32_t)HQP * 16u, Cfg::G1_CHUNKS, 0);
        if (elect_one()) umma_commit(s_full);
        __syncwarp();
        mbar_wait(p_full, ph);
        tc_fence_after();
        gemm(SLABS_P, tmem_base + TM_O, id_o, 128u * 16u, 1, 0);
        if (elect_one()) umma_commit(o_full);
        __syncwarp();
        mbar_wait(y_full, ph);
        tc_fence_after();
        gemm(4, tmem_base + TM_Z, id_z, (uint32_t)DT_NQP * 16u, 0, 4096u);
        if (elect_one()) umma_commit(z_full);
        __syncwarp();
      }
    }
  } else {
    // ======================================================================================= operand loader
    {
      const unsigned char* img = reinterpret_cast<const unsigned char*>(p.img);
      int nb = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        size_t off = 0;
        for (int st = 0; st < 4 * Cfg::G1_CHUNKS + SLABS_P + 1; ++st) {
          const uint32_t bytes = st < 4 * Cfg::G1_CHUNKS ? (uint32_t)Cfg::G1_BYTES : 16384u;
          const int sb = nb % NBR;
          mbar_wait(b_empty(sb), ((uint32_t)(nb / NBR) & 1u) ^ 1u);
          if (elect_one()) {
            mbar_arrive_expect_tx(b_full(sb), bytes);
            bulk_g2s(smem_u32(ring + (size_t)sb * B_STAGE), img + off, bytes, b_full(sb));
          }
          __syncwarp();
          off += bytes;
          ++nb;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (tid < p.n_obj && hist_s[tid]) atomicAdd(p.obj_count + tid, hist_s[tid]);
  if (warp == 8) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int NQ16>
static int s2c_tc_launch_t(const float* x, const float* pos, long long nv, const float* A, const float* c,
                           const float* U, const float* bo, const float* ln_w, const float* ln_b, float ln_eps,
                           const float* E, const int* q_obj, int nq, int heads, int n_obj, float* x_out,
                           float* logits, unsigned char* label, int* obj_count, void* ws, cudaStream_t st) {
  using Cfg = S2cCfg<NQ16>;
  constexpr int HQP = Cfg::HQP;
  uint4* img = static_cast<uint4*>(ws);
  float* cpad = reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + s2c_img_bytes<NQ16>());
  constexpr int total = 4 * 4 * HQP + Cfg::SLABS_P * 4 * 128 + 4 * 4 * DT_NQP;
  s2c_prep_kernel<NQ16><<<(total + 255) / 256, 256, 0, st>>>(A, c, U, E, nq, heads, img, cpad);
  AG3D_LAUNCH_CHECK("s2c_prep");
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(s2c_tc_kernel<NQ16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr = true;
  }
  S2cParams p;
  p.x = x; p.pos = pos; p.nv = nv; p.img = img; p.cpad = cpad; p.bo = bo; p.ln_w = ln_w; p.ln_b = ln_b;
  p.ln_eps = ln_eps; p.q_obj = q_obj; p.nq = nq; p.n_obj = n_obj; p.x_out = x_out; p.logits = logits;
  p.label = label; p.obj_count = obj_count;
  long long tiles = (nv + TC_BM - 1) / TC_BM;
  const int grid = (int)std::min<long long>(tiles, sm_count());
  s2c_tc_kernel<NQ16><<<grid, DT_THREADS, Cfg::SMEM, st>>>(p);
  AG3D_LAUNCH_CHECK("s2c_tc");
  return AG3D_OK;
}

// ============================================================================================== split-row variant
// Same three chained GEMMs, but the voxel tile never passes through registers on its way in or out: x and pos are
// "split" rows (32-channel slabs of 64 B bf16 hi | 64 B bf16 lo) and the TMA engine drops their [128 voxels x 128 B]
// slab tiles into SWIZZLE_128B shared memory, which is directly the K-major A operand of
//   S = pos . A^T + x . A^T      (two passes over the weight slabs: the pos tile of the NEXT voxel tile is fetched as
//                                 soon as the score GEMM of this one has read it, the x tile once the mask head has
//                                 read Y, and the pos pass runs while the x tile is still landing)
// The probabilities P never touch shared memory: they overwrite the scores in TENSOR MEMORY (bf16 hi | lo, two queries
// per 32-bit column) and are the A operand of O = P . U from there (tcgen05.mma with A in TMEM).  The LayerNorm
// output Y is written in the swizzled slab format in place over the x tile (whose values are the residual input), is
// the A operand of the mask-head GEMM, and leaves as split rows through four bulk tensor stores per tile (x_out_split,
// nullable: the last decoder layer's features are never read).
// Roles: 16 compute warps (softmax, LayerNorm, mask head; four threads per voxel row), MMA issuer, operand-ring loader,
// voxel-tile TMA producer.
constexpr int DS_COMPUTE_WARPS = 16;                  // TMEM lane quarter q4 = warp & 3, part g = warp >> 2
constexpr int DS_THREADS = DS_COMPUTE_WARPS * 32 + 96;
constexpr uint32_t DS_SLAB = 16384;                    // [128 voxels x 128 B]
constexpr uint32_t DS_MISC = 8192;

template <int NQ16>
struct S2sCfg {
  static constexpr int HQP = 8 * NQ16;
  static constexpr int SLABS_P = HQP / 32;
  static constexpr int G1_CHUNKS = HQP * 128 > 16384 ? 2 : 1;
  static constexpr int G1_BYTES = HQP * 128 / G1_CHUNKS;
  static constexpr int NBR = 5;
  static constexpr uint32_t OFF_X = 0;
  static constexpr uint32_t OFF_POS = 4 * DS_SLAB;
  static constexpr uint32_t OFF_RING = OFF_POS + 4 * DS_SLAB;
  static constexpr uint32_t OFF_MISC = OFF_RING + NBR * 16384;
  static constexpr size_t SMEM = OFF_MISC + DS_MISC;
};

struct S2sParams {
  long long nv;
  const uint4* img; const float* cpad;
  const float* bo; const float* ln_w; const float* ln_b; float ln_eps;
  const int* q_obj; int nq; int n_obj;
  float* x_out; float* logits; unsigned char* label; int* obj_count;
  int debug;     // measurement aid (AG3D_S2C_DEBUG): 1/2/4 no MMAs in G1/G2/G3, 8 no voxel-tile loads, 16 no operand-ring copies,
                 // 32 no softmax, 64 no LayerNorm phase work
};

__device__ __forceinline__ void ds_tma_tile(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int col, int row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(col), "r"(row)
      : "memory");
}
// byte offset of the 16-byte chunk `ch` (0..3 hi pieces, 4..7 lo pieces of 8 channels) of row r inside a swizzled slab tile
__device__ __forceinline__ uint32_t ds_chunk_off(int r, int ch) { return (uint32_t)(r * 128 + ((ch ^ (r & 7)) << 4)); }
// D[tmem] (+)= A[tmem] . B[smem]: A (bf16, two K elements per 32-bit column, row = lane) read from tensor memory
__device__ __forceinline__ void ds_umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int NQ16>
__global__ void __launch_bounds__(DS_THREADS, 1)
s2c_split_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_pos,
                 const __grid_constant__ CUtensorMap tm_out, const S2sParams p) {
  using Cfg = S2sCfg<NQ16>;
  constexpr int HQP = Cfg::HQP, SLABS_P = Cfg::SLABS_P, NBR = Cfg::NBR;
  constexpr uint32_t B_STAGE = 16384;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* Xs = smem + Cfg::OFF_X;                         // x tile (4 slabs); later Y in place
  unsigned char* Ps = smem + Cfg::OFF_POS;                       // pos tile (4 slabs)
  unsigned char* ring = smem + Cfg::OFF_RING;
  unsigned char* misc = smem + Cfg::OFF_MISC;
  uint64_t* bars = reinterpret_cast<uint64_t*>(misc);            // [0..6] phase barriers, [8..15] ring full, [16..23] ring empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 192);
  int* hist_s = reinterpret_cast<int*>(misc + 256);              // [32]
  int* qobj_s = reinterpret_cast<int*>(misc + 384);              // [32]
  float* vec_s = reinterpret_cast<float*>(misc + 512);           // bo[128] | ln_w[128] | ln_b[128]
  float* cpad_s = reinterpret_cast<float*>(misc + 2048);         // [HQP <= 256]
  float* lnred_s = reinterpret_cast<float*>(misc + 3072);        // [2][128][4]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t x_full = bar_base, s_full = bar_base + 8, p_full = bar_base + 16, o_full = bar_base + 24,
                 y_full = bar_base + 32, z_full = bar_base + 40, pos_full = bar_base + 48;
  auto b_full = [&](int s) { return bar_base + 8u * (8 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (16 + s); };

  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    mbar_init(x_full, 1); mbar_init(pos_full, 1); mbar_init(p_full, DS_COMPUTE_WARPS); mbar_init(y_full, DS_COMPUTE_WARPS);
    mbar_init(s_full, 1); mbar_init(o_full, 1); mbar_init(z_full, 1);
    for (int s = 0; s < NBR; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) { hist_s[tid] = 0; qobj_s[tid] = (tid < p.nq) ? p.q_obj[tid] : -1; }
  for (int i = tid; i < 128; i += DS_THREADS) {
    vec_s[i] = p.bo[i]; vec_s[128 + i] = p.ln_w[i]; vec_s[256 + i] = p.ln_b[i];
  }
  for (int i = tid; i < HQP; i += DS_THREADS) cpad_s[i] = p.cpad[i];
  if (warp == DS_COMPUTE_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long n_tiles = (p.nv + TC_BM - 1) / TC_BM;

  if (warp < DS_COMPUTE_WARPS) {
    // ======================================================================================= compute warps
    const int q4 = warp & 3, g = warp >> 2;             // g = part: heads 2g, 2g+1 of the softmax, 32-channel slab g of the LayerNorm
    const int r = q4 * 32 + lane;                       // tile row owned in all phases
    const uint32_t t_lane = tmem_base + ((uint32_t)(q4 * 32) << 16);
    // ---- P3 (mask head, part 0 only: 32 query columns) of a tile runs AFTER the softmax of the next tile: the mask-head
    //      GEMM and its hand-off then sit under the out-projection GEMM of the next tile instead of on the chain
    auto mask_head = [&](long long row0_prev, uint32_t ph_prev) {
      mbar_wait(z_full, ph_prev);
      tc_fence_after();
      if (g == 0) {
        const long long row = row0_prev + r;
        float z[DT_NQP];
        tmem_ld16(t_lane + TM_Z, z);
        tmem_ld16(t_lane + TM_Z + 16, z + 16);
        if (row < p.nv) {
          float best = -INFINITY;
          int arg = 0;
          for (int ob = 0; ob < p.n_obj; ++ob) {
            float m = -INFINITY;
#pragma unroll
            for (int q = 0; q < DT_NQP; ++q) m = (qobj_s[q] == ob) ? fmaxf(m, z[q]) : m;
            p.logits[(size_t)row * p.n_obj + ob] = m;
            if (m > best || ob == 0) { best = m; arg = ob; }
          }
          p.label[row] = (unsigned char)arg;
          atomicAdd(&hist_s[arg], 1);
        }
      }
      tc_fence_before();
    };
    int it = 0;
    long long row0_prev = -1;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const long long row0 = tile * TC_BM;
      // ---- P1: per-head softmax over the queries; this thread: row r, heads 2g, 2g+1.  The probabilities replace the
      //      scores in tensor memory (hi pairs in columns [0, HQP/2), lo pairs in [HQP/2, HQP)), so the four threads of a
      //      row first read all their scores (named barrier 2 + q4) and only then write.
      mbar_wait(s_full, ph);
      tc_fence_after();
      float sc[2 * NQ16];
      if constexpr (NQ16 % 16 == 0) {
#pragma unroll
        for (int ch = 0; ch < 2 * NQ16 / 16; ++ch) tmem_ld16(t_lane + TM_S + 2 * g * NQ16 + ch * 16, sc + ch * 16);
      } else {
#pragma unroll
        for (int ch = 0; ch < 2 * NQ16 / 8; ++ch) tmem_ld8(t_lane + TM_S + 2 * g * NQ16 + ch * 8, sc + ch * 8);
      }
      asm volatile("bar.sync %0, 128;" ::"r"(2 + q4) : "memory");
      uint32_t ph_[NQ16], pl_[NQ16];                    // 2 * NQ16 probabilities as bf16 pairs
      if (p.debug & 32) {
#pragma unroll
        for (int i = 0; i < NQ16; ++i) ph_[i] = pl_[i] = 0u;
      } else {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          float* s = sc + hh * NQ16;
          const int col0 = (2 * g + hh) * NQ16;
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < NQ16; ++i) { s[i] += cpad_s[col0 + i]; mx = fmaxf(mx, s[i]); }
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < NQ16; ++i) { s[i] = __expf(s[i] - mx); sum += s[i]; }
          const float inv = 1.f / sum;
#pragma unroll
          for (int i = 0; i < NQ16 / 2; ++i) split2(s[2 * i] * inv, s[2 * i + 1] * inv, ph_[hh * (NQ16 / 2) + i], pl_[hh * (NQ16 / 2) + i]);
        }
      }
      {
        const uint32_t c_hi = t_lane + TM_S + (uint32_t)(g * NQ16), c_lo = c_hi + HQP / 2;
#pragma unroll
        for (int ch = 0; ch < NQ16 / 8; ++ch) {
          tmem_st8(c_hi + ch * 8, ph_ + ch * 8);
          tmem_st8(c_lo + ch * 8, pl_ + ch * 8);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
      if (row0_prev >= 0) mask_head(row0_prev, ph ^ 1u);      // mask head of the previous tile (Z is not rewritten before y_full)

      // ---- P2: o + bo + x -> LayerNorm -> y; this thread: row r, channels 32g .. 32g+31 (slab g)
      mbar_wait(o_full, ph);
      mbar_wait(x_full, ph);                            // the x tile was written by the TMA engine: observe its barrier
      tc_fence_after();
      const long long row = row0 + r;
      const bool valid = row < p.nv;
      if (!(p.debug & 64)) {
        float v[32];
        tmem_ld16(t_lane + TM_O + 32 * g, v);
        tmem_ld16(t_lane + TM_O + 32 * g + 16, v + 16);
        unsigned char* slab = Xs + (size_t)g * DS_SLAB;
        float sum = 0.f;
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {                // 8 channels: hi chunk c8, lo chunk 4 + c8
          const uint4 hh = *reinterpret_cast<const uint4*>(slab + ds_chunk_off(r, c8));
          const uint4 ll = *reinterpret_cast<const uint4*>(slab + ds_chunk_off(r, 4 + c8));
          const uint32_t hw[4] = {hh.x, hh.y, hh.z, hh.w}, lw[4] = {ll.x, ll.y, ll.z, ll.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float x0 = __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16);
            const float x1 = __uint_as_float(hw[e] & 0xFFFF0000u) + __uint_as_float(lw[e] & 0xFFFF0000u);
            const int c = c8 * 8 + 2 * e;
            v[c] = x0 + (v[c] + vec_s[32 * g + c]);
            v[c + 1] = x1 + (v[c + 1] + vec_s[32 * g + c + 1]);
            sum += v[c] + v[c + 1];
          }
        }
        lnred_s[(0 * 128 + r) * 4 + g] = sum;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float4 s4 = *reinterpret_cast<const float4*>(lnred_s + (0 * 128 + r) * 4);
        const float mean = ((s4.x + s4.y) + (s4.z + s4.w)) * (1.f / DT_D);
        float sq = 0.f;
#pragma unroll
        for (int c = 0; c < 32; ++c) { const float d = v[c] - mean; sq = fmaf(d, d, sq); }
        lnred_s[(1 * 128 + r) * 4 + g] = sq;
        asm volatile("bar.sync 1, 512;" ::: "memory");
        const float4 q4v = *reinterpret_cast<const float4*>(lnred_s + (1 * 128 + r) * 4);
        const float var = ((q4v.x + q4v.y) + (q4v.z + q4v.w)) * (1.f / DT_D);
        const float rstd = 1.f / sqrtf(var + p.ln_eps);
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
          float y[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int c = 32 * g + c8 * 8 + e;
            y[e] = (v[c8 * 8 + e] - mean) * rstd * vec_s[128 + c] + vec_s[256 + c];
          }
          uint32_t h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) split2(y[2 * e], y[2 * e + 1], h[e], l[e]);
          *reinterpret_cast<uint4*>(slab + ds_chunk_off(r, c8)) = make_uint4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<uint4*>(slab + ds_chunk_off(r, 4 + c8)) = make_uint4(l[0], l[1], l[2], l[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(y_full);

      row0_prev = row0;
    }
    if (row0_prev >= 0) mask_head(row0_prev, (uint32_t)(it - 1) & 1u);
  } else if (warp == DS_COMPUTE_WARPS) {
    // ======================================================================================= MMA issuer
    const uint32_t id_s = umma_idesc_bf16(HQP), id_o = umma_idesc_bf16(128), id_z = umma_idesc_bf16(DT_NQP);
    const uint32_t d_hi32 = umma_desc_hi32(128);                     // ring operands: no swizzle, SBO = 128
    const uint32_t sw_hi32 = umma_desc_hi32(1024) | (2u << 29);      // voxel-side operands: SWIZZLE_128B slab tiles
    const uint32_t x_lo32 = umma_desc_lo32(smem_u32(Xs), 16), pos_lo32 = umma_desc_lo32(smem_u32(Ps), 16);
    const uint32_t ring_u32 = smem_u32(ring);
    int nb = 0, it = 0;
    // one GEMM pass = `slabs` K-slabs x (2 k-steps x 3 products).  chunks: ring slots per slab (2: hi and lo piece in
    // separate slots; 1: whole slab image in one slot; 0: all slabs in ONE slot).  a_smem: descriptor base of the
    // swizzled A slab tiles; 0 = A is P in tensor memory (a_tmem: hi pairs, + HQP/2 columns: lo pairs)
    auto gemm = [&](int slabs, uint32_t d, uint32_t idesc, uint32_t b_lbo, int chunks, uint32_t b_slab, uint32_t a_smem,
                    uint32_t a_tmem, bool first, bool skip) {
      uint32_t b_hi = 0, b_lo = 0;
      for (int s = 0; s < slabs; ++s) {
        if (chunks || s == 0) {
          const int sb = nb % NBR;
          mbar_wait(b_full(sb), (uint32_t)(nb / NBR) & 1u);
          b_hi = umma_desc_lo32(ring_u32 + (uint32_t)sb * B_STAGE, b_lbo);
          b_lo = b_hi + ((4u * b_lbo) >> 4);
          if (chunks == 2) {
            const int sb2 = (nb + 1) % NBR;
            mbar_wait(b_full(sb2), (uint32_t)((nb + 1) / NBR) & 1u);
            b_lo = umma_desc_lo32(ring_u32 + (uint32_t)sb2 * B_STAGE, b_lbo);
          }
          tc_fence_after();
        }
        const uint32_t off = chunks ? 0u : ((uint32_t)s * b_slab) >> 4;
        const bool release = chunks || s == slabs - 1;
        if (elect_one()) {
          if (!skip) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint64_t db_hi = umma_desc_join(d_hi32, b_hi + off + ks * ((2u * b_lbo) >> 4));
              const uint64_t db_lo = umma_desc_join(d_hi32, b_lo + off + ks * ((2u * b_lbo) >> 4));
              const uint32_t acc0 = (!first || s || ks) ? 1u : 0u;
              if (a_smem) {
                const uint32_t ah = a_smem + (uint32_t)s * (DS_SLAB >> 4) + ks * 2u;
                const uint64_t da_hi = umma_desc_join(sw_hi32, ah), da_lo = umma_desc_join(sw_hi32, ah + 4u);
                umma_bf16(d, da_hi, db_hi, idesc, acc0);
                umma_bf16(d, da_hi, db_lo, idesc, 1u);
                umma_bf16(d, da_lo, db_hi, idesc, 1u);
              } else {
                const uint32_t ta = a_tmem + (uint32_t)(s * 16 + ks * 8);     // 16 queries = 8 columns per k-step
                ds_umma_ts(d, ta, db_hi, idesc, acc0);
                ds_umma_ts(d, ta, db_lo, idesc, 1u);
                ds_umma_ts(d, ta + HQP / 2, db_hi, idesc, 1u);
              }
            }
          }
          if (release) {
            umma_commit(b_empty(nb % NBR));
            if (chunks == 2) umma_commit(b_empty((nb + 1) % NBR));
          }
        }
        __syncwarp();
        if (release) nb += chunks == 2 ? 2 : 1;
      }
    };
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      mbar_wait(pos_full, ph);
      tc_fence_after();
      gemm(4, tmem_base + TM_S, id_s, (uint32_t)HQP * 16u, Cfg::G1_CHUNKS, 0, pos_lo32, 0u, true, p.debug & 1);
      mbar_wait(x_full, ph);
      tc_fence_after();
      gemm(4, tmem_base + TM_S, id_s, (uint32_t)HQP * 16u, Cfg::G1_CHUNKS, 0, x_lo32, 0u, false, p.debug & 1);
      if (elect_one()) umma_commit(s_full);
      __syncwarp();
      mbar_wait(p_full, ph);
      tc_fence_after();
      gemm(SLABS_P, tmem_base + TM_O, id_o, 128u * 16u, 1, 0, 0u, tmem_base + TM_S, true, p.debug & 2);
      if (elect_one()) umma_commit(o_full);
      __syncwarp();
      mbar_wait(y_full, ph);
      tc_fence_after();
      gemm(4, tmem_base + TM_Z, id_z, (uint32_t)DT_NQP * 16u, 0, 4096u, x_lo32, 0u, true, p.debug & 4);
      if (elect_one()) umma_commit(z_full);
      __syncwarp();
    }
  } else if (warp == DS_COMPUTE_WARPS + 1) {
    // ======================================================================================= operand loader
    // per tile: the score operand twice (pos pass, x pass), the out-projection slabs, the mask embeddings
    const unsigned char* img = reinterpret_cast<const unsigned char*>(p.img);
    constexpr int N1 = 4 * Cfg::G1_CHUNKS;
    int nb = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int st = 0; st < 2 * N1 + SLABS_P + 1; ++st) {
        const int k = st < N1 ? st : st - N1;                         // position in the image (consumption order G1 | G2 | G3)
        const uint32_t bytes = k < N1 ? (uint32_t)Cfg::G1_BYTES : 16384u;
        const size_t off = k < N1 ? (size_t)k * Cfg::G1_BYTES : (size_t)N1 * Cfg::G1_BYTES + (size_t)(k - N1) * 16384u;
        const int sb = nb % NBR;
        mbar_wait(b_empty(sb), ((uint32_t)(nb / NBR) & 1u) ^ 1u);
        if (elect_one()) {
          if (p.debug & 16) {
            mbar_arrive(b_full(sb));
          } else {
            mbar_arrive_expect_tx(b_full(sb), bytes);
            bulk_g2s(smem_u32(ring + (size_t)sb * B_STAGE), img + off, bytes, b_full(sb));
          }
        }
        __syncwarp();
        ++nb;
      }
    }
  } else {
    // ======================================================================================= voxel-tile producer
    // pos tile of tile i+1: as soon as the score GEMM of tile i is complete (s_full); x tile of tile i+1: once the mask
    // head of tile i has read Y (z_full) and the bulk stores of Y have read it too.  The updated features leave the
    // same way: Y sits in the x region in exactly the layout of the output's tensor map, so one thread stores the tile
    // with four bulk tensor copies (rows past the end are clipped) instead of every thread writing 16-byte pieces.
    auto load_tile = [&](unsigned char* dst, const CUtensorMap* tm, uint32_t bar, long long tile) {
      if (elect_one()) {
        const int row = (int)(tile * TC_BM);
        if (p.debug & 8) {
          mbar_arrive(bar);
        } else {
          mbar_arrive_expect_tx(bar, 4u * DS_SLAB);
#pragma unroll
          for (int c = 0; c < 4; ++c) ds_tma_tile(smem_u32(dst) + (uint32_t)c * DS_SLAB, tm, bar, c * 64, row);
        }
      }
      __syncwarp();
    };
    int it = 0;
    if ((long long)blockIdx.x < n_tiles) {
      load_tile(Ps, &tm_pos, pos_full, blockIdx.x);
      load_tile(Xs, &tm_x, x_full, blockIdx.x);
    }
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const long long next = tile + gridDim.x;
      mbar_wait(s_full, ph);
      if (next < n_tiles) load_tile(Ps, &tm_pos, pos_full, next);
      if (p.x_out) {
        mbar_wait(y_full, ph);                           // Y of this tile is written (and fenced for the async proxy)
        if (elect_one()) {
          const int row = (int)(tile * TC_BM);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&tm_out),
                         "r"(smem_u32(Xs) + (uint32_t)c * DS_SLAB), "r"(c * 64), "r"(row)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        __syncwarp();
      }
      mbar_wait(z_full, ph);
      if (p.x_out && elect_one()) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the store has read Y
      __syncwarp();
      if (next < n_tiles) load_tile(Xs, &tm_x, x_full, next);
    }
    if (p.x_out && elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (tid < p.n_obj && hist_s[tid]) atomicAdd(p.obj_count + tid, hist_s[tid]);
  if (warp == DS_COMPUTE_WARPS) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

bool split_rows_tile_map(CUtensorMap* tm, const float* base, long long rows, int box_rows);   // decoder_c2s_tc2.cu
size_t s2c_tc_workspace_bytes(int nq);

template <int NQ16>
static int s2c_split_launch_t(const float* x, const float* pos, long long nv, const float* A, const float* c,
                              const float* U, const float* bo, const float* ln_w, const float* ln_b, float ln_eps,
                              const float* E, const int* q_obj, int nq, int heads, int n_obj, float* x_out,
                              float* logits, unsigned char* label, int* obj_count, void* ws, cudaStream_t st) {
  using Cfg = S2sCfg<NQ16>;
  constexpr int HQP = Cfg::HQP;
  uint4* img = static_cast<uint4*>(ws);
  float* cpad = reinterpret_cast<float*>(static_cast<unsigned char*>(ws) + s2c_img_bytes<NQ16>());
  constexpr int total = 4 * 4 * HQP + Cfg::SLABS_P * 4 * 128 + 4 * 4 * DT_NQP;
  s2c_prep_kernel<NQ16><<<(total + 255) / 256, 256, 0, st>>>(A, c, U, E, nq, heads, img, cpad);
  AG3D_LAUNCH_CHECK("s2c_prep");
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(s2c_split_kernel<NQ16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    attr = true;
  }
  alignas(64) CUtensorMap tm_x, tm_pos;
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_pos, 0, sizeof(tm_pos));
  AG3D_CHECK_ARG(split_rows_tile_map(&tm_x, x, nv, TC_BM) && split_rows_tile_map(&tm_pos, pos, nv, TC_BM),
                 "cuTensorMapEncodeTiled failed for the voxel rows");
  alignas(64) CUtensorMap tm_out;
  memset(&tm_out, 0, sizeof(tm_out));
  AG3D_CHECK_ARG(split_rows_tile_map(&tm_out, x_out ? x_out : x, nv, TC_BM), "cuTensorMapEncodeTiled failed for the output rows");
  S2sParams p;
  p.nv = nv; p.img = img; p.cpad = cpad; p.bo = bo; p.ln_w = ln_w; p.ln_b = ln_b;
  p.ln_eps = ln_eps; p.q_obj = q_obj; p.nq = nq; p.n_obj = n_obj; p.x_out = x_out; p.logits = logits;
  p.label = label; p.obj_count = obj_count;
  { const char* e = getenv("AG3D_S2C_DEBUG"); p.debug = e ? atoi(e) : 0; }
  long long tiles = (nv + TC_BM - 1) / TC_BM;
  const int grid = (int)std::min<long long>(tiles, sm_count());
  s2c_split_kernel<NQ16><<<grid, DS_THREADS, Cfg::SMEM, st>>>(tm_x, tm_pos, tm_out, p);
  AG3D_LAUNCH_CHECK("s2c_split");
  return AG3D_OK;
}

int s2c_split_launch(const float* x, const float* pos, long long nv, const float* A, const float* c, const float* U,
                     const float* bo, const float* ln_w, const float* ln_b, float ln_eps, const float* E,
                     const int* q_obj, int nq, int heads, int n_obj, float* x_out, float* logits, unsigned char* label,
                     int* obj_count, void* ws, size_t ws_bytes, cudaStream_t st) {
  AG3D_CHECK_ARG(heads == 8 && nq <= 24, "split-row s2c handles 8 heads and at most 24 queries");
  AG3D_CHECK_ARG(nv < 2147483647LL, "too many voxels");
  AG3D_CHECK_ARG(ws && aligned16(ws) && ws_bytes >= s2c_tc_workspace_bytes(nq), "s2c workspace too small");
  if (nq <= 16)
    return s2c_split_launch_t<16>(x, pos, nv, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out, logits,
                                  label, obj_count, ws, st);
  return s2c_split_launch_t<24>(x, pos, nv, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out, logits,
                                label, obj_count, ws, st);
}

// heads are padded to 16, 24 or 32 query columns (AG3D_S2C_PACK=0 disables the 24-column variant: measurement aid)
static int s2c_pad(int nq) {
  static int pack = -1;
  if (pack < 0) { const char* e = getenv("AG3D_S2C_PACK"); pack = (e && e[0] == '0') ? 0 : 1; }
  return nq <= 16 ? 16 : ((nq <= 24 && pack) ? 24 : 32);
}

size_t s2c_tc_workspace_bytes(int nq) {
  (void)nq;
  return s2c_img_bytes<32>() + 256 * 4 + 256;      // the largest variant
}

int s2c_tc_launch(const float* x, const float* pos, long long nv, const float* A, const float* c, const float* U,
                  const float* bo, const float* ln_w, const float* ln_b, float ln_eps, const float* E,
                  const int* q_obj, int nq, int heads, int n_obj, float* x_out, float* logits, unsigned char* label,
                  int* obj_count, void* ws, size_t ws_bytes, cudaStream_t st) {
  AG3D_CHECK_ARG(heads == 8 && nq <= 32, "tensor-core s2c handles 8 heads and at most 32 queries");
  AG3D_CHECK_ARG(ws && aligned16(ws) && ws_bytes >= s2c_tc_workspace_bytes(nq), "s2c workspace too small");
  const int pad = s2c_pad(nq);
  if (pad == 16)
    return s2c_tc_launch_t<16>(x, pos, nv, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out, logits,
                               label, obj_count, ws, st);
  if (pad == 24)
    return s2c_tc_launch_t<24>(x, pos, nv, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out, logits,
                               label, obj_count, ws, st);
  return s2c_tc_launch_t<32>(x, pos, nv, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out, logits,
                             label, obj_count, ws, st);
}

}  // namespace ag3d
