// K9/K10 (exact-fp32 variant): click->scene attention (streaming, split over voxels, LSE merge) and
// scene->click attention fused with residual + LayerNorm + mask head.  FFMA arithmetic, fp32 throughout.
// Both kernels make ONE pass over the voxel features; the key/value (c2s) and query/out (s2c) projections over
// Nv are never materialised (SURVEY.md §7 "algebraic folding").
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace ag3d {

constexpr int D = 128;        // hidden dim (models/agile3d.py: hidden_dim, main.py:44)
constexpr int TV = 64;        // voxels per tile
constexpr int DEC_THREADS = 256;
#define NEG_INF (-INFINITY)

// ================================================================================================ c2s
// grid (n_cta, n_groups); query group g holds nqg = ceil(nq / n_groups) queries -> HQ = heads*nqg <= 32*J rows.
template <int J>
__global__ void __launch_bounds__(DEC_THREADS, 1)
c2s_partial_kernel(const float* __restrict__ x, const float* __restrict__ pos, long long nv,
                   const float* __restrict__ qfold, int nq, int heads, int nqg,
                   const unsigned char* __restrict__ label, const int* __restrict__ q_obj,
                   const int* __restrict__ obj_count, float* __restrict__ part_m, float* __restrict__ part_l,
                   float* __restrict__ part_acc) {
  constexpr int HQP = 32 * J;
  constexpr int QT_LD = HQP + 0;      // Qt[c][hq]
  constexpr int XP_LD = TV + 4;       // XPt[c][v]
  constexpr int P_LD = TV + 1;        // P[hq][v]
  extern __shared__ __align__(16) float smem[];
  float* Qt = smem;                                  // D * QT_LD
  float* Xs = Qt + D * QT_LD;                        // TV * D      (x tile, row-major)
  float* XPt = Xs + TV * D;                          // D * XP_LD   (x+pos, transposed)
  float* Ps = XPt + D * XP_LD;                       // HQP * P_LD  (softmax numerators of the tile)
  float* alpha_s = Ps + HQP * P_LD;                  // HQP
  int* rowobj_s = reinterpret_cast<int*>(alpha_s + HQP);   // HQP: object id a row is restricted to, -1 = unrestricted, -2 = padding
  int* lab_s = rowobj_s + HQP;                       // TV

  const int tid = threadIdx.x;
  const int g = blockIdx.y;
  const int q0 = g * nqg;
  const int nq_here = min(nqg, nq - q0);
  const int HQ = heads * nq_here;

  // ---- stage the folded queries of this group, transposed: Qt[c][r], r = h*nq_here + ql
  for (int idx = tid; idx < HQP * (D / 4); idx += DEC_THREADS) {
    const int r = idx / (D / 4), c4 = idx % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < HQ) {
      const int h = r / nq_here, ql = r % nq_here;
      v = __ldg(reinterpret_cast<const float4*>(qfold + ((long long)h * nq + q0 + ql) * D + c4 * 4));
    }
    Qt[(c4 * 4 + 0) * QT_LD + r] = v.x;
    Qt[(c4 * 4 + 1) * QT_LD + r] = v.y;
    Qt[(c4 * 4 + 2) * QT_LD + r] = v.z;
    Qt[(c4 * 4 + 3) * QT_LD + r] = v.w;
  }
  for (int r = tid; r < HQP; r += DEC_THREADS) {
    int ro = -2;
    if (r < HQ) {
      ro = -1;
      if (label) {
        const int o = q_obj[q0 + r % nq_here];
        if (obj_count[o] > 0) ro = o;   // all-masked rows are un-masked (models/agile3d.py:369,375)
      }
    }
    rowobj_s[r] = ro;
  }

  // phase-1 mapping: 16 row groups x 16 voxel groups
  const int ty = tid >> 4, tx = tid & 15;
  // phase-3 mapping: 32 row groups x 8 channel groups
  const int ty3 = tid >> 3, tx3 = tid & 7;

  float m_run[2 * J], l_run[2 * J];
#pragma unroll
  for (int j = 0; j < 2 * J; ++j) { m_run[j] = NEG_INF; l_run[j] = 0.f; }
  float acc3[J][16];
#pragma unroll
  for (int i = 0; i < J; ++i)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc3[i][c] = 0.f;

  const long long n_tiles = (nv + TV - 1) / TV;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long v0 = tile * TV;
    __syncthreads();   // previous tile's phase 3 is done with Xs / Ps
    // ---- load x and pos tile: Xs row-major, XPt transposed
    for (int idx = tid; idx < TV * (D / 4); idx += DEC_THREADS) {
      const int v = idx / (D / 4), c4 = idx % (D / 4);
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), pv = xv;
      if (v0 + v < nv) {
        xv = __ldg(reinterpret_cast<const float4*>(x + (v0 + v) * D + c4 * 4));
        pv = __ldg(reinterpret_cast<const float4*>(pos + (v0 + v) * D + c4 * 4));
      }
      *reinterpret_cast<float4*>(Xs + v * D + c4 * 4) = xv;
      XPt[(c4 * 4 + 0) * XP_LD + v] = xv.x + pv.x;
      XPt[(c4 * 4 + 1) * XP_LD + v] = xv.y + pv.y;
      XPt[(c4 * 4 + 2) * XP_LD + v] = xv.z + pv.z;
      XPt[(c4 * 4 + 3) * XP_LD + v] = xv.w + pv.w;
    }
    if (tid < TV) lab_s[tid] = (v0 + tid < nv) ? (label ? (int)label[v0 + tid] : 0) : -3;
    __syncthreads();

    // ---- phase 1: S[r][v] = Qt[:, r] . XPt[:, v]
    float s[2 * J][4];
#pragma unroll
    for (int j = 0; j < 2 * J; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
#pragma unroll 4
    for (int c = 0; c < D; ++c) {
      const float4 xp = *reinterpret_cast<const float4*>(XPt + c * XP_LD + tx * 4);
#pragma unroll
      for (int j = 0; j < 2 * J; ++j) {
        const float qv = Qt[c * QT_LD + ty + 16 * j];
        s[j][0] = fmaf(qv, xp.x, s[j][0]);
        s[j][1] = fmaf(qv, xp.y, s[j][1]);
        s[j][2] = fmaf(qv, xp.z, s[j][2]);
        s[j][3] = fmaf(qv, xp.w, s[j][3]);
      }
    }
    // ---- phase 2: mask, online softmax statistics, P tile
    int lab4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) lab4[i] = lab_s[tx * 4 + i];
#pragma unroll
    for (int j = 0; j < 2 * J; ++j) {
      const int r = ty + 16 * j;
      const int ro = rowobj_s[r];
      float tmax = NEG_INF;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const bool dead = (lab4[i] == -3) || (ro == -2) || (ro >= 0 && lab4[i] != ro);
        if (dead) s[j][i] = NEG_INF;
        tmax = fmaxf(tmax, s[j][i]);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
      const float m_new = fmaxf(m_run[j], tmax);
      float a = 1.f, psum = 0.f;
      float p[4] = {0.f, 0.f, 0.f, 0.f};
      if (m_new != NEG_INF) {
        a = __expf(m_run[j] - m_new);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          p[i] = __expf(s[j][i] - m_new);
          psum += p[i];
        }
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
      l_run[j] = l_run[j] * a + psum;
      m_run[j] = m_new;
#pragma unroll
      for (int i = 0; i < 4; ++i) Ps[r * P_LD + tx * 4 + i] = p[i];
      if (tx == 0) alpha_s[r] = a;
    }
    __syncthreads();

    // ---- phase 3: acc[r][:] = acc[r][:] * alpha[r] + sum_v P[r][v] * Xs[v][:]
#pragma unroll
    for (int i = 0; i < J; ++i) {
      const float a = alpha_s[ty3 + 32 * i];
#pragma unroll
      for (int c = 0; c < 16; ++c) acc3[i][c] *= a;
    }
#pragma unroll 2
    for (int v = 0; v < TV; ++v) {
      float xv[16];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 t = *reinterpret_cast<const float4*>(Xs + v * D + tx3 * 16 + c4 * 4);
        xv[c4 * 4 + 0] = t.x; xv[c4 * 4 + 1] = t.y; xv[c4 * 4 + 2] = t.z; xv[c4 * 4 + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < J; ++i) {
        const float p = Ps[(ty3 + 32 * i) * P_LD + v];
#pragma unroll
        for (int c = 0; c < 16; ++c) acc3[i][c] = fmaf(p, xv[c], acc3[i][c]);
      }
    }
  }

  // ---- write this CTA's partial (m, l, acc) for the merge
  const long long pbase = ((long long)g * gridDim.x + blockIdx.x) * HQP;
  if (tx == 0) {
#pragma unroll
    for (int j = 0; j < 2 * J; ++j) {
      part_m[pbase + ty + 16 * j] = m_run[j];
      part_l[pbase + ty + 16 * j] = l_run[j];
    }
  }
#pragma unroll
  for (int i = 0; i < J; ++i) {
    float* dst = part_acc + (pbase + ty3 + 32 * i) * D + tx3 * 16;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4)
      *reinterpret_cast<float4*>(dst + c4 * 4) =
          make_float4(acc3[i][c4 * 4], acc3[i][c4 * 4 + 1], acc3[i][c4 * 4 + 2], acc3[i][c4 * 4 + 3]);
  }
}

// grid heads*nq blocks of 128 threads: log-sum-exp merge of the per-CTA partials
__global__ void __launch_bounds__(D) c2s_merge_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l,
                                                      const float* __restrict__ part_acc, int n_cta, int HQP, int nq, int nqg,
                                                      float* __restrict__ ctx, float* __restrict__ lse_out) {
  // one CTA per (head, query) row, thread = channel.  The partial maxima / sums of the n_cta partial results are folded
  // by the whole CTA first (weights in shared memory), so the channel loop is n_cta independent, unrolled loads.
  __shared__ float w_s[1024];
  __shared__ float red_s[8];
  const int row = blockIdx.x;             // h*nq + q
  const int h = row / nq, q = row % nq;
  const int g = q / nqg, ql = q % nqg;
  const int nq_here = min(nqg, nq - g * nqg);
  const int r = h * nq_here + ql;
  const int c = threadIdx.x, warp = c >> 5, lane = c & 31;
  const long long p0 = (long long)g * n_cta * HQP + r;          // partial i lives at p0 + i * HQP
  float M = NEG_INF;
  for (int i = c; i < n_cta; i += D) M = fmaxf(M, part_m[p0 + (long long)i * HQP]);
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  if (lane == 0) red_s[warp] = M;
  __syncthreads();
  M = fmaxf(fmaxf(red_s[0], red_s[1]), fmaxf(red_s[2], red_s[3]));
  float L = 0.f;
  for (int i = c; i < n_cta; i += D) {
    const float m = part_m[p0 + (long long)i * HQP];
    const float w = (m == NEG_INF || M == NEG_INF) ? 0.f : __expf(m - M);
    w_s[i] = w;
    L += part_l[p0 + (long long)i * HQP] * w;
  }
  for (int o = 16; o > 0; o >>= 1) L += __shfl_xor_sync(0xffffffffu, L, o);
  if (lane == 0) red_s[4 + warp] = L;
  __syncthreads();
  L = (red_s[4] + red_s[5]) + (red_s[6] + red_s[7]);
  float a = 0.f;
  const float* acc = part_acc + p0 * D + c;
#pragma unroll 8
  for (int i = 0; i < n_cta; ++i) a = fmaf(acc[(long long)i * HQP * D], w_s[i], a);
  ctx[(long long)row * D + c] = (L > 0.f) ? a / L : 0.f;
  if (lse_out && c == 0) lse_out[row] = (L > 0.f) ? M + logf(L) : INFINITY;   // +inf: exp(s - lse) = 0 in the backward
}

// ================================================================================================ s2c
// thread (ty, tx): voxels ty*4..+3 ; score columns tx + 16 j (j < J2) ; output channels tx*8..+7
template <int J2>
__global__ void __launch_bounds__(DEC_THREADS)
s2c_mask_kernel(const float* x, const float* __restrict__ pos, long long nv, const float* __restrict__ A,
                const float* __restrict__ cvec, const float* __restrict__ U, const float* __restrict__ bo,
                const float* __restrict__ ln_w, const float* __restrict__ ln_b, float ln_eps,
                const float* __restrict__ E, const int* __restrict__ q_obj, int nq, int heads, int n_obj,
                float* x_out, float* __restrict__ logits, unsigned char* __restrict__ label,
                int* __restrict__ obj_count) {
  constexpr int HQP = 16 * J2;
  constexpr int XP_LD = TV + 4;
  constexpr int SP_LD = (HQP + 1 > D + 1) ? HQP + 1 : D + 1;   // S/P tile, later Y tile
  constexpr int AT_LD = HQP + 1;
  // Bb holds, in turn: a 32-channel slab of A^T [32][AT_LD], a 32-row slab of U [32][D], E^T [D][33]
  extern __shared__ __align__(16) float smem[];
  float* XPt = smem;                       // D * XP_LD ; later Z[TV][33] and per-object maxima
  float* SP = XPt + D * XP_LD;             // TV * SP_LD
  float* Bb = SP + TV * SP_LD;             // max(32*AT_LD, D*33) floats
  __shared__ int hist_s[64];

  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;
  const int HQ = heads * nq;
  const long long v0 = (long long)blockIdx.x * TV;
  if (tid < 64) hist_s[tid] = 0;

  for (int idx = tid; idx < TV * (D / 4); idx += DEC_THREADS) {
    const int v = idx / (D / 4), c4 = idx % (D / 4);
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), pv = xv;
    if (v0 + v < nv) {
      xv = *reinterpret_cast<const float4*>(x + (v0 + v) * D + c4 * 4);
      pv = __ldg(reinterpret_cast<const float4*>(pos + (v0 + v) * D + c4 * 4));
    }
    XPt[(c4 * 4 + 0) * XP_LD + v] = xv.x + pv.x;
    XPt[(c4 * 4 + 1) * XP_LD + v] = xv.y + pv.y;
    XPt[(c4 * 4 + 2) * XP_LD + v] = xv.z + pv.z;
    XPt[(c4 * 4 + 3) * XP_LD + v] = xv.w + pv.w;
  }

  // ---- phase 1: S[v][r] = (x+pos)[v] . A[r] + c[r]
  float s[4][J2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < J2; ++j) s[i][j] = 0.f;
  for (int c0 = 0; c0 < D; c0 += 32) {
    __syncthreads();
    for (int idx = tid; idx < HQP * 8; idx += DEC_THREADS) {      // A[r][c0 .. c0+31] -> At[cc][r]
      const int r = idx >> 3, c4 = idx & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < HQ) v = __ldg(reinterpret_cast<const float4*>(A + (long long)r * D + c0 + c4 * 4));
      Bb[(c4 * 4 + 0) * AT_LD + r] = v.x;
      Bb[(c4 * 4 + 1) * AT_LD + r] = v.y;
      Bb[(c4 * 4 + 2) * AT_LD + r] = v.z;
      Bb[(c4 * 4 + 3) * AT_LD + r] = v.w;
    }
    __syncthreads();
#pragma unroll 4
    for (int cc = 0; cc < 32; ++cc) {
      const float4 xp = *reinterpret_cast<const float4*>(XPt + (c0 + cc) * XP_LD + ty * 4);
#pragma unroll
      for (int j = 0; j < J2; ++j) {
        const float av = Bb[cc * AT_LD + tx + 16 * j];
        s[0][j] = fmaf(xp.x, av, s[0][j]);
        s[1][j] = fmaf(xp.y, av, s[1][j]);
        s[2][j] = fmaf(xp.z, av, s[2][j]);
        s[3][j] = fmaf(xp.w, av, s[3][j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < J2; ++j) {
    const int r = tx + 16 * j;
    const float cb = (r < HQ) ? __ldg(cvec + r) : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) SP[(ty * 4 + i) * SP_LD + r] = s[i][j] + cb;
  }
  __syncthreads();

  // ---- phase 2: softmax over the nq queries of each head, in place
  for (int pidx = tid; pidx < TV * heads; pidx += DEC_THREADS) {
    const int v = pidx / heads, h = pidx % heads;
    float* row = SP + v * SP_LD + h * nq;
    float mx = NEG_INF;
    for (int q = 0; q < nq; ++q) mx = fmaxf(mx, row[q]);
    float sum = 0.f;
    for (int q = 0; q < nq; ++q) {
      const float e = __expf(row[q] - mx);
      row[q] = e;
      sum += e;
    }
    const float inv = 1.f / sum;
    for (int q = 0; q < nq; ++q) row[q] *= inv;
  }

  // ---- phase 3: O[v][:] = P[v][:] @ U
  float o[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) o[i][c] = 0.f;
  for (int r0 = 0; r0 < HQ; r0 += 32) {
    __syncthreads();
    for (int idx = tid; idx < 32 * (D / 4); idx += DEC_THREADS) {
      const int r = idx / (D / 4), c4 = idx % (D / 4);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r0 + r < HQ) v = __ldg(reinterpret_cast<const float4*>(U + (long long)(r0 + r) * D + c4 * 4));
      *reinterpret_cast<float4*>(Bb + r * D + c4 * 4) = v;
    }
    __syncthreads();
    const int rmax = min(32, HQ - r0);
    for (int r = 0; r < rmax; ++r) {
      const float4 u0 = *reinterpret_cast<const float4*>(Bb + r * D + tx * 8);
      const float4 u1 = *reinterpret_cast<const float4*>(Bb + r * D + tx * 8 + 4);
      const float uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = SP[(ty * 4 + i) * SP_LD + r0 + r];
#pragma unroll
        for (int c = 0; c < 8; ++c) o[i][c] = fmaf(p, uu[c], o[i][c]);
      }
    }
  }
  __syncthreads();   // everyone is done reading P and the last U chunk

  // stage E transposed for the mask head: Et[c][q] (stride 33)
  for (int idx = tid; idx < 32 * (D / 4); idx += DEC_THREADS) {
    const int q = idx / (D / 4), c4 = idx % (D / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq) v = __ldg(reinterpret_cast<const float4*>(E + (long long)q * D + c4 * 4));
    Bb[(c4 * 4 + 0) * 33 + q] = v.x;
    Bb[(c4 * 4 + 1) * 33 + q] = v.y;
    Bb[(c4 * 4 + 2) * 33 + q] = v.z;
    Bb[(c4 * 4 + 3) * 33 + q] = v.w;
  }

  // ---- residual + bias + LayerNorm (two-pass), write y to global and to the Y tile (SP region)
  {
    float bo8[8], w8[8], b8[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      bo8[c] = __ldg(bo + tx * 8 + c);
      w8[c] = __ldg(ln_w + tx * 8 + c);
      b8[c] = __ldg(ln_b + tx * 8 + c);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long v = v0 + ty * 4 + i;
      float xr[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (v < nv) {
        const float4 a0 = *reinterpret_cast<const float4*>(x + v * D + tx * 8);
        const float4 a1 = *reinterpret_cast<const float4*>(x + v * D + tx * 8 + 4);
        xr[0] = a0.x; xr[1] = a0.y; xr[2] = a0.z; xr[3] = a0.w;
        xr[4] = a1.x; xr[5] = a1.y; xr[6] = a1.z; xr[7] = a1.w;
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        o[i][c] = xr[c] + (o[i][c] + bo8[c]);
        sum += o[i][c];
      }
#pragma unroll
      for (int sft = 8; sft > 0; sft >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, sft);
      const float mean = sum * (1.f / D);
      float var = 0.f;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float d = o[i][c] - mean;
        var = fmaf(d, d, var);
      }
#pragma unroll
      for (int sft = 8; sft > 0; sft >>= 1) var += __shfl_xor_sync(0xffffffffu, var, sft);
      const float rstd = 1.f / sqrtf(var * (1.f / D) + ln_eps);
      float y[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        y[c] = (o[i][c] - mean) * rstd * w8[c] + b8[c];
        SP[(ty * 4 + i) * SP_LD + tx * 8 + c] = y[c];
      }
      if (v < nv) {
        *reinterpret_cast<float4*>(x_out + v * D + tx * 8) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(x_out + v * D + tx * 8 + 4) = make_float4(y[4], y[5], y[6], y[7]);
      }
    }
  }
  __syncthreads();

  // ---- phase 4: Z[v][q] = Y[v] . E[q]   (q = tx, tx + 16)
  float z[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) z[i][0] = z[i][1] = 0.f;
#pragma unroll 4
  for (int c = 0; c < D; ++c) {
    const float e0 = Bb[c * 33 + tx], e1 = Bb[c * 33 + tx + 16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float yv = SP[(ty * 4 + i) * SP_LD + c];
      z[i][0] = fmaf(yv, e0, z[i][0]);
      z[i][1] = fmaf(yv, e1, z[i][1]);
    }
  }
  float* Zs = XPt;   // [TV][33], XPt is free now
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    Zs[(ty * 4 + i) * 33 + tx] = z[i][0];
    Zs[(ty * 4 + i) * 33 + tx + 16] = z[i][1];
  }
  __syncthreads();

  // ---- phase 5: per-object maxima, logits, label, histogram
  float* Ms = XPt + TV * 33;   // [TV][33]
  if (tid < TV) {
    const int v = tid;
    for (int ob = 0; ob < n_obj; ++ob) Ms[v * 33 + ob] = NEG_INF;
    for (int q = 0; q < nq; ++q) {
      const int ob = __ldg(q_obj + q);
      Ms[v * 33 + ob] = fmaxf(Ms[v * 33 + ob], Zs[v * 33 + q]);
    }
    if (v0 + v < nv) {
      float best = Ms[v * 33];
      int arg = 0;
      for (int ob = 0; ob < n_obj; ++ob) {
        const float val = Ms[v * 33 + ob];
        logits[(v0 + v) * n_obj + ob] = val;
        if (val > best) { best = val; arg = ob; }
      }
      label[v0 + v] = (unsigned char)arg;
      atomicAdd(&hist_s[arg], 1);
    }
  }
  __syncthreads();
  if (tid < n_obj && hist_s[tid]) atomicAdd(obj_count + tid, hist_s[tid]);
}

template <int J2>
constexpr size_t s2c_smem_bytes() {
  constexpr int HQP = 16 * J2;
  constexpr int SP_LD = (HQP + 1 > D + 1) ? HQP + 1 : D + 1;
  constexpr int AT = 32 * (HQP + 1);
  constexpr int BF = (AT > D * 33) ? AT : D * 33;
  return sizeof(float) * (size_t)(D * (TV + 4) + TV * SP_LD + BF);
}

template <int J>
constexpr size_t c2s_smem_bytes() {
  constexpr int HQP = 32 * J;
  return sizeof(float) * (size_t)(D * HQP + TV * D + D * (TV + 4) + HQP * (TV + 1) + HQP) + sizeof(int) * (HQP + TV);
}

static inline int c2s_groups(int nq) { return (nq + 19) / 20; }

size_t c2s_tc_workspace_bytes(int nq, int heads);
int c2s_tc_launch(const float* x, const float* pos, long long nv, const float* qfold, int nq, int heads,
                  const unsigned char* label, const int* q_obj, const int* obj_count, void* ws, size_t ws_bytes,
                  cudaStream_t st, float** part_m, float** part_l, float** part_acc, int* n_cta_out, int* nqg_out);
size_t s2c_tc_workspace_bytes(int nq);
int s2c_tc_launch(const float* x, const float* pos, long long nv, const float* A, const float* c, const float* U,
                  const float* bo, const float* ln_w, const float* ln_b, float ln_eps, const float* E,
                  const int* q_obj, int nq, int heads, int n_obj, float* x_out, float* logits, unsigned char* label,
                  int* obj_count, void* ws, size_t ws_bytes, cudaStream_t st);

int c2s_split_launch(const float* x_split, const float* pos_split, long long nv, const float* qfold, int nq, int heads,
                     const unsigned char* label, const int* q_obj, const int* obj_count, void* ws, size_t ws_bytes,
                     cudaStream_t st, float** part_m, float** part_l, float** part_acc, int* n_cta_out, int* nqg_out);
int s2c_split_launch(const float* x, const float* pos, long long nv, const float* A, const float* c, const float* U,
                     const float* bo, const float* ln_w, const float* ln_b, float ln_eps, const float* E,
                     const int* q_obj, int nq, int heads, int n_obj, float* x_out, float* logits, unsigned char* label,
                     int* obj_count, void* ws, size_t ws_bytes, cudaStream_t st);
size_t s2c_mq_workspace_bytes(int nq);
int s2c_mq_launch(const float* x, const float* pos, long long nv, const float* A, const float* c, const float* U,
                  const float* bo, const float* ln_w, const float* ln_b, float ln_eps, const float* E,
                  const int* q_obj, int nq, int heads, int n_obj, float* x_out, float* logits, unsigned char* label,
                  int* obj_count, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace ag3d

using namespace ag3d;

extern "C" {

size_t ag3d_c2s_workspace_bytes(int32_t nq, int32_t heads) {
  const int groups = c2s_groups(nq);
  const size_t rows = (size_t)groups * (size_t)(sm_count() > 0 ? sm_count() : 148) * 160;
  const size_t simt = rows * (D + 2) * sizeof(float) + 256;
  const size_t tc = c2s_tc_workspace_bytes(nq, heads);
  return simt > tc ? simt : tc;
}

int ag3d_c2s_attn_fwd(const float* x, const float* pos, int64_t nv, const float* qfold, int32_t nq,
                      int32_t heads, const uint8_t* label, const int32_t* q_obj, const int32_t* obj_count,
                      float* ctx, float* lse, int32_t algo, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(nv > 0 && nq > 0, "empty problem");
  AG3D_CHECK_ARG(heads == 8, "heads must be 8 (hidden 128 = 8 x 16)");
  AG3D_CHECK_ARG(x && pos && qfold && ctx && aligned16(x) && aligned16(pos) && aligned16(qfold) && aligned16(ctx),
                 "bad pointers");
  AG3D_CHECK_ARG(!label || (q_obj && obj_count), "a label mask needs q_obj and obj_count");
  if (!ws || ws_bytes < ag3d_c2s_workspace_bytes(nq, heads)) {
    set_error("c2s workspace too small");
    return AG3D_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  if (algo == AG3D_ALGO_AUTO) algo = AG3D_ALGO_TC;
  if (algo == AG3D_ALGO_TC) {
    float *pm, *pl, *pa;
    int n_cta_tc, nqg_tc;
    if (int rc = c2s_tc_launch(x, pos, nv, qfold, nq, heads, label, q_obj, obj_count, ws, ws_bytes, st, &pm, &pl, &pa,
                               &n_cta_tc, &nqg_tc))
      return rc;
    c2s_merge_kernel<<<heads * nq, D, 0, st>>>(pm, pl, pa, n_cta_tc, 128, nq, nqg_tc, ctx, lse);
    AG3D_LAUNCH_CHECK("c2s_merge");
    return AG3D_OK;
  }
  AG3D_CHECK_ARG(algo == AG3D_ALGO_SIMT, "unknown algo");
  const int groups = c2s_groups(nq);
  const int nqg = (nq + groups - 1) / groups;
  int J = (heads * nqg + 31) / 32;                     // template instances: 3, 4, 5
  if (J < 3) J = 3;
  const int HQP = 32 * J;
  long long n_tiles = (nv + TV - 1) / TV;
  int n_cta = sm_count() / groups;
  if (n_cta < 1) n_cta = 1;
  if (n_cta > n_tiles) n_cta = (int)n_tiles;
  float* part_m = static_cast<float*>(ws);
  float* part_l = part_m + (size_t)groups * n_cta * HQP;
  float* part_acc = part_l + (size_t)groups * n_cta * HQP;
  part_acc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(part_acc) + 15) & ~(uintptr_t)15);
  dim3 grid(n_cta, groups);
#define LAUNCH_C2S(JJ)                                                                                       \
  do {                                                                                                       \
    static bool attr = false;                                                                                \
    if (!attr) {                                                                                             \
      AG3D_CUDA(cudaFuncSetAttribute(c2s_partial_kernel<JJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                     (int)c2s_smem_bytes<JJ>()));                                           \
      attr = true;                                                                                           \
    }                                                                                                        \
    c2s_partial_kernel<JJ><<<grid, DEC_THREADS, c2s_smem_bytes<JJ>(), st>>>(                                 \
        x, pos, nv, qfold, nq, heads, nqg, label, q_obj, obj_count, part_m, part_l, part_acc);               \
  } while (0)
  if (J == 3) { LAUNCH_C2S(3); }
  else if (J == 4) { LAUNCH_C2S(4); }
  else { AG3D_CHECK_ARG(J == 5, "internal: query group too large"); LAUNCH_C2S(5); }
#undef LAUNCH_C2S
  AG3D_LAUNCH_CHECK("c2s_partial");
  c2s_merge_kernel<<<heads * nq, D, 0, st>>>(part_m, part_l, part_acc, n_cta, HQP, nq, nqg, ctx, lse);
  AG3D_LAUNCH_CHECK("c2s_merge");
  return AG3D_OK;
}

int ag3d_c2s_attn_fwd_split(const float* x_split, const float* pos_split, int64_t nv, const float* qfold, int32_t nq,
                            int32_t heads, const uint8_t* label, const int32_t* q_obj, const int32_t* obj_count,
                            float* ctx, float* lse, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(nv > 0 && nq > 0, "empty problem");
  AG3D_CHECK_ARG(heads == 8, "heads must be 8 (hidden 128 = 8 x 16)");
  AG3D_CHECK_ARG(x_split && pos_split && qfold && ctx && aligned16(x_split) && aligned16(pos_split) && aligned16(qfold) &&
                     aligned16(ctx), "bad pointers");
  AG3D_CHECK_ARG(!label || (q_obj && obj_count), "a label mask needs q_obj and obj_count");
  if (!ws || ws_bytes < ag3d_c2s_workspace_bytes(nq, heads)) {
    set_error("c2s workspace too small");
    return AG3D_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  float *pm, *pl, *pa;
  int n_cta_tc, nqg_tc;
  if (int rc = c2s_split_launch(x_split, pos_split, nv, qfold, nq, heads, label, q_obj, obj_count, ws, ws_bytes, st, &pm,
                                &pl, &pa, &n_cta_tc, &nqg_tc))
    return rc;
  c2s_merge_kernel<<<heads * nq, D, 0, st>>>(pm, pl, pa, n_cta_tc, 128, nq, nqg_tc, ctx, lse);
  AG3D_LAUNCH_CHECK("c2s_merge");
  return AG3D_OK;
}

size_t ag3d_s2c_workspace_bytes(int32_t nq) {
  if (nq < 1 || nq > 256) return 0;
  return nq <= 32 ? s2c_tc_workspace_bytes(nq) : s2c_mq_workspace_bytes(nq);
}

int ag3d_s2c_mask_fwd(const float* x, const float* pos, int64_t nv, const float* A, const float* c,
                      const float* U, const float* bo, const float* ln_w, const float* ln_b, float ln_eps,
                      const float* E, const int32_t* q_obj, int32_t nq, int32_t heads, int32_t n_obj,
                      float* x_out, float* logits, uint8_t* label, int32_t* obj_count, int32_t algo, void* ws,
                      size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(nv > 0 && nq > 0, "empty problem");
  AG3D_CHECK_ARG(heads == 8, "heads must be 8 (hidden 128 = 8 x 16)");
  AG3D_CHECK_ARG(nq <= 256, "at most 256 click queries per scene");
  AG3D_CHECK_ARG(n_obj >= 1 && n_obj <= 32, "n_obj must be 1..32");
  AG3D_CHECK_ARG(x && pos && A && c && U && bo && ln_w && ln_b && E && q_obj && x_out && logits && label && obj_count,
                 "bad pointers");
  AG3D_CHECK_ARG(aligned16(x) && aligned16(pos) && aligned16(A) && aligned16(U) && aligned16(E) && aligned16(x_out),
                 "pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  if (algo == AG3D_ALGO_AUTO) algo = ws ? AG3D_ALGO_TC : AG3D_ALGO_SIMT;
  if (nq > 32) {      // query groups of 16 with two-pass softmax statistics (decoder_mq.cu); tensor-core path only
    AG3D_CHECK_ARG(algo == AG3D_ALGO_TC, "more than 32 click queries need the tensor-core path (pass the workspace)");
    return s2c_mq_launch(x, pos, nv, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out, logits, label,
                         obj_count, ws, ws_bytes, st);
  }
  if (algo == AG3D_ALGO_TC)
    return s2c_tc_launch(x, pos, nv, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out, logits,
                         label, obj_count, ws, ws_bytes, st);
  AG3D_CHECK_ARG(algo == AG3D_ALGO_SIMT, "unknown algo");
  const unsigned grid = (unsigned)((nv + TV - 1) / TV);
  const int J2 = (heads * nq + 15) / 16;
#define LAUNCH_S2C(JJ)                                                                                      \
  do {                                                                                                      \
    static bool attr = false;                                                                               \
    if (!attr) {                                                                                            \
      AG3D_CUDA(cudaFuncSetAttribute(s2c_mask_kernel<JJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                     (int)s2c_smem_bytes<JJ>()));                                          \
      attr = true;                                                                                          \
    }                                                                                                       \
    s2c_mask_kernel<JJ><<<grid, DEC_THREADS, s2c_smem_bytes<JJ>(), st>>>(                                   \
        x, pos, nv, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out, logits, label,      \
        obj_count);                                                                                         \
  } while (0)
  if (J2 <= 6) { LAUNCH_S2C(6); }
  else if (J2 <= 8) { LAUNCH_S2C(8); }
  else if (J2 <= 10) { LAUNCH_S2C(10); }
  else if (J2 <= 12) { LAUNCH_S2C(12); }
  else { LAUNCH_S2C(16); }
#undef LAUNCH_S2C
  AG3D_LAUNCH_CHECK("s2c_mask");
  return AG3D_OK;
}

int ag3d_s2c_mask_fwd_split(const float* x_split, const float* pos_split, int64_t nv, const float* A, const float* c,
                            const float* U, const float* bo, const float* ln_w, const float* ln_b, float ln_eps,
                            const float* E, const int32_t* q_obj, int32_t nq, int32_t heads, int32_t n_obj,
                            float* x_out_split, float* logits, uint8_t* label, int32_t* obj_count, void* ws,
                            size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(nv > 0 && nq > 0, "empty problem");
  AG3D_CHECK_ARG(heads == 8, "heads must be 8 (hidden 128 = 8 x 16)");
  AG3D_CHECK_ARG(nq <= 24, "the split-row kernel handles at most 24 click queries per scene");
  AG3D_CHECK_ARG(n_obj >= 1 && n_obj <= 32, "n_obj must be 1..32");
  AG3D_CHECK_ARG(x_split && pos_split && A && c && U && bo && ln_w && ln_b && E && q_obj && logits && label && obj_count,
                 "bad pointers");
  AG3D_CHECK_ARG(aligned16(x_split) && aligned16(pos_split) && aligned16(A) && aligned16(U) && aligned16(E) &&
                     (!x_out_split || aligned16(x_out_split)), "pointers must be 16-byte aligned");
  return s2c_split_launch(x_split, pos_split, nv, A, c, U, bo, ln_w, ln_b, ln_eps, E, q_obj, nq, heads, n_obj, x_out_split,
                          logits, label, obj_count, ws, ws_bytes, as_stream(stream));
}

}  // extern "C"
