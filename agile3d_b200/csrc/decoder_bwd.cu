// Backward of the two voxel-streaming decoder kernels (SURVEY.md §8 row a11 for a6/a8/a9):
//   c2s:  ctx[(h,q)] = sum_v softmax_v(qfold[(h,q)] . (x_v + pos_v)) x_v
//   s2c:  x' = LayerNorm(x + softmax_q((x + pos) A^T + c) U + bo),  logits[v,o] = max_{q in o} x'_v . E[q]
// One pass over the voxels per kernel (fp32 FFMA).  Per-voxel gradients (dx) are finished in the kernel; the
// per-voxel factors of the query-side gradients (dS, a, dy, g) are written once and contracted over the voxels by
// ag3d_spconv_bwd_weight's "X^T dY" kernel; column sums (dbo, dln_w, dln_b, dc) are accumulated per CTA.
// Small matrices are row-padded to HQP = 16*J rows ((head, query) pairs) / 32 queries by the caller.
#include <math.h>

#include "common.cuh"

namespace ag3d {

constexpr int BD = 128;         // hidden dim
constexpr int BTV = 64;         // voxels per tile
constexpr int BTHREADS = 256;   // thread (vy, tx): voxels vy*4..+3, columns tx + 16 j
constexpr int NQP = 32;         // padded query count

// acc[i][j] += sum_k As[(vy*4+i)*lda + k] * Bg[k*ldb + tx + 16 j]
template <int J>
__device__ __forceinline__ void tile_mm(const float* As, int lda, const float* __restrict__ Bg, int ldb, int K, int vy,
                                        int tx, float (&acc)[4][J]) {
  const float* a0 = As + (vy * 4) * lda;
#pragma unroll 2
  for (int k = 0; k < K; ++k) {
    const float x0 = a0[k], x1 = a0[lda + k], x2 = a0[2 * lda + k], x3 = a0[3 * lda + k];
    const float* b = Bg + (long long)k * ldb + tx;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const float bv = __ldg(b + 16 * j);
      acc[0][j] = fmaf(x0, bv, acc[0][j]);
      acc[1][j] = fmaf(x1, bv, acc[1][j]);
      acc[2][j] = fmaf(x2, bv, acc[2][j]);
      acc[3][j] = fmaf(x3, bv, acc[3][j]);
    }
  }
}

template <int J>
__device__ __forceinline__ void zero_acc(float (&acc)[4][J]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < J; ++j) acc[i][j] = 0.f;
}

__device__ __forceinline__ float half_warp_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ================================================================================================ c2s backward
template <int J>
__global__ void __launch_bounds__(BTHREADS, 1)
c2s_bwd_kernel(const float* __restrict__ x, const float* __restrict__ pos, long long nv,
               const float* __restrict__ qf, const float* __restrict__ qft, const float* __restrict__ dctx,
               const float* __restrict__ dctxt, const float* __restrict__ lse, const float* __restrict__ dr,
               const int* __restrict__ rowobj, const unsigned char* __restrict__ label, float* __restrict__ dx,
               float* __restrict__ ds_out) {
  constexpr int HQP = 16 * J;
  extern __shared__ __align__(16) float smem[];
  float* Xs = smem;                 // BTV * BD
  float* XPs = Xs + BTV * BD;       // BTV * BD
  float* Ps = XPs + BTV * BD;       // BTV * HQP
  float* DSs = Ps + BTV * HQP;      // BTV * HQP
  const int tid = threadIdx.x, vy = tid >> 4, tx = tid & 15;
  float lse_r[J], dr_r[J];
  int ro_r[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    lse_r[j] = __ldg(lse + tx + 16 * j);
    dr_r[j] = __ldg(dr + tx + 16 * j);
    ro_r[j] = __ldg(rowobj + tx + 16 * j);
  }
  const long long n_tiles = (nv + BTV - 1) / BTV;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long v0 = tile * BTV;
    __syncthreads();
    for (int idx = tid; idx < BTV * (BD / 4); idx += BTHREADS) {
      const int v = idx / (BD / 4), c4 = idx % (BD / 4);
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), pv = xv;
      if (v0 + v < nv) {
        xv = __ldg(reinterpret_cast<const float4*>(x + (v0 + v) * BD) + c4);
        pv = __ldg(reinterpret_cast<const float4*>(pos + (v0 + v) * BD) + c4);
      }
      *reinterpret_cast<float4*>(Xs + v * BD + c4 * 4) = xv;
      *reinterpret_cast<float4*>(XPs + v * BD + c4 * 4) = make_float4(xv.x + pv.x, xv.y + pv.y, xv.z + pv.z, xv.w + pv.w);
    }
    __syncthreads();
    int lab[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long v = v0 + vy * 4 + i;
      lab[i] = (v < nv) ? (label ? (int)label[v] : 0) : -3;
    }
    float p[4][J];
    {
      float s[4][J];
      zero_acc<J>(s);
      tile_mm<J>(XPs, BD, qft, HQP, BD, vy, tx, s);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const bool dead = (lab[i] == -3) || (ro_r[j] == -2) || (ro_r[j] >= 0 && lab[i] != ro_r[j]);
          p[i][j] = dead ? 0.f : __expf(s[i][j] - lse_r[j]);
          Ps[(vy * 4 + i) * HQP + tx + 16 * j] = p[i][j];
        }
    }
    {
      float dp[4][J];
      zero_acc<J>(dp);
      tile_mm<J>(Xs, BD, dctxt, HQP, BD, vy, tx, dp);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long v = v0 + vy * 4 + i;
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const float ds = p[i][j] * (dp[i][j] - dr_r[j]);
          DSs[(vy * 4 + i) * HQP + tx + 16 * j] = ds;
          if (v < nv) ds_out[v * HQP + tx + 16 * j] = ds;
        }
      }
    }
    __syncthreads();
    {
      float o[4][8];
      zero_acc<8>(o);
      tile_mm<8>(Ps, HQP, dctx, BD, HQP, vy, tx, o);
      tile_mm<8>(DSs, HQP, qf, BD, HQP, vy, tx, o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long v = v0 + vy * 4 + i;
        if (v < nv) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dx[v * BD + tx + 16 * j] = o[i][j];
        }
      }
    }
  }
}

// ================================================================================================ s2c backward
// column-sum partial layout per CTA: [dbo 128 | dln_w 128 | dln_b 128 | dc HQP]
template <int J>
__global__ void __launch_bounds__(BTHREADS, 1)
s2c_bwd_kernel(const float* __restrict__ x, const float* __restrict__ pos, long long nv,
               const float* __restrict__ A, const float* __restrict__ At, const float* __restrict__ cvec,
               const float* __restrict__ U, const float* __restrict__ Ut, const float* __restrict__ bo,
               const float* __restrict__ ln_w, const float* __restrict__ ln_b, float ln_eps,
               const float* __restrict__ E, const float* __restrict__ Et, const int* __restrict__ q_obj, int nq,
               int heads, int n_obj, const float* __restrict__ dxo, const float* __restrict__ dlogits,
               float* __restrict__ dx, float* __restrict__ a_out, float* __restrict__ ds_out,
               float* __restrict__ dy_out, float* __restrict__ g_out, float* __restrict__ colpart) {
  constexpr int HQP = 16 * J;
  extern __shared__ __align__(16) float smem[];
  float* B0 = smem;                  // BTV * BD : x+pos, then x', then dy
  float* Ns = B0 + BTV * BD;         // BTV * BD : normalised y
  float* Aa = Ns + BTV * BD;         // BTV * HQP: attention probabilities
  float* DAs = Aa + BTV * HQP;       // BTV * HQP: da, then dS
  float* Gs = DAs + BTV * HQP;       // BTV * NQP: x'.E products
  float* G2s = Gs + BTV * NQP;       // BTV * NQP: routed logit gradients
  __shared__ int qobj_s[NQP];
  const int tid = threadIdx.x, vy = tid >> 4, tx = tid & 15;
  const int HQ = heads * nq;
  if (tid < NQP) qobj_s[tid] = tid < nq ? q_obj[tid] : -1;

  float lw[8], lb[8], bov[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    lw[j] = __ldg(ln_w + tx + 16 * j);
    lb[j] = __ldg(ln_b + tx + 16 * j);
    bov[j] = __ldg(bo + tx + 16 * j);
  }
  float s_bo[8], s_lnw[8], s_lnb[8], s_dc[J];
#pragma unroll
  for (int j = 0; j < 8; ++j) s_bo[j] = s_lnw[j] = s_lnb[j] = 0.f;
#pragma unroll
  for (int j = 0; j < J; ++j) s_dc[j] = 0.f;

  const long long n_tiles = (nv + BTV - 1) / BTV;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long v0 = tile * BTV;
    __syncthreads();
    for (int idx = tid; idx < BTV * (BD / 4); idx += BTHREADS) {
      const int v = idx / (BD / 4), c4 = idx % (BD / 4);
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), pv = xv;
      if (v0 + v < nv) {
        xv = __ldg(reinterpret_cast<const float4*>(x + (v0 + v) * BD) + c4);
        pv = __ldg(reinterpret_cast<const float4*>(pos + (v0 + v) * BD) + c4);
      }
      *reinterpret_cast<float4*>(B0 + v * BD + c4 * 4) = make_float4(xv.x + pv.x, xv.y + pv.y, xv.z + pv.z, xv.w + pv.w);
    }
    __syncthreads();
    // ---- scores + per-head softmax
    {
      float s[4][J];
      zero_acc<J>(s);
      tile_mm<J>(B0, BD, At, HQP, BD, vy, tx, s);
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const float cb = __ldg(cvec + tx + 16 * j);
#pragma unroll
        for (int i = 0; i < 4; ++i) Aa[(vy * 4 + i) * HQP + tx + 16 * j] = s[i][j] + cb;
      }
    }
    __syncthreads();
    for (int pidx = tid; pidx < BTV * heads; pidx += BTHREADS) {
      const int v = pidx / heads, h = pidx % heads;
      float* row = Aa + v * HQP + h * nq;
      float mx = -INFINITY;
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
    if (HQ < HQP)
      for (int idx = tid; idx < BTV * (HQP - HQ); idx += BTHREADS)
        Aa[(idx / (HQP - HQ)) * HQP + HQ + idx % (HQP - HQ)] = 0.f;
    __syncthreads();
    for (int idx = tid; idx < BTV * HQP; idx += BTHREADS) {
      const long long v = v0 + idx / HQP;
      if (v < nv) a_out[v * HQP + idx % HQP] = Aa[idx];
    }
    // ---- y = x + a U + bo, LayerNorm
    float rstd[4];
    {
      float y[4][8];
      zero_acc<8>(y);
      tile_mm<8>(Aa, HQP, U, BD, HQP, vy, tx, y);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long v = v0 + vy * 4 + i;
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xv = (v < nv) ? __ldg(x + v * BD + tx + 16 * j) : 0.f;
          y[i][j] += xv + bov[j];
          sum += y[i][j];
        }
        const float mean = half_warp_sum(sum) * (1.f / BD);
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          y[i][j] -= mean;
          sq = fmaf(y[i][j], y[i][j], sq);
        }
        rstd[i] = rsqrtf(half_warp_sum(sq) * (1.f / BD) + ln_eps);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float n = y[i][j] * rstd[i];
          Ns[(vy * 4 + i) * BD + tx + 16 * j] = n;
          B0[(vy * 4 + i) * BD + tx + 16 * j] = fmaf(n, lw[j], lb[j]);
        }
      }
    }
    __syncthreads();
    // ---- x'.E products, first-maximum routing of the logit gradients
    {
      float pr[4][2];
      zero_acc<2>(pr);
      tile_mm<2>(B0, BD, Et, NQP, BD, vy, tx, pr);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        Gs[(vy * 4 + i) * NQP + tx] = pr[i][0];
        Gs[(vy * 4 + i) * NQP + tx + 16] = pr[i][1];
      }
    }
    __syncthreads();
    for (int idx = tid; idx < BTV * NQP; idx += BTHREADS) {
      const int v = idx / NQP, q = idx % NQP;
      float g = 0.f;
      const long long gv = v0 + v;
      if (q < nq && gv < nv && dlogits) {
        const int o = qobj_s[q];
        const float mine = Gs[v * NQP + q];
        bool first_max = true;
        for (int q2 = 0; q2 < nq; ++q2) {
          if (qobj_s[q2] != o || q2 == q) continue;
          const float other = Gs[v * NQP + q2];
          if (other > mine || (other == mine && q2 < q)) first_max = false;
        }
        if (first_max) g = __ldg(dlogits + gv * n_obj + o);
      }
      G2s[idx] = g;
      if (gv < nv) g_out[gv * NQP + q] = g;
    }
    __syncthreads();
    // ---- d x' (total) -> LayerNorm backward -> dy
    {
      float t[4][8];
      zero_acc<8>(t);
      tile_mm<8>(G2s, NQP, E, BD, NQP, vy, tx, t);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long v = v0 + vy * 4 + i;
        float m1 = 0.f, m2 = 0.f;
        float nrm[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (dxo && v < nv) t[i][j] += __ldg(dxo + v * BD + tx + 16 * j);
          nrm[j] = Ns[(vy * 4 + i) * BD + tx + 16 * j];
          if (v < nv) {
            s_lnb[j] += t[i][j];
            s_lnw[j] = fmaf(t[i][j], nrm[j], s_lnw[j]);
          }
          t[i][j] *= lw[j];
          m1 += t[i][j];
          m2 = fmaf(t[i][j], nrm[j], m2);
        }
        m1 = half_warp_sum(m1) * (1.f / BD);
        m2 = half_warp_sum(m2) * (1.f / BD);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float dyv = rstd[i] * (t[i][j] - m1 - nrm[j] * m2);
          if (v >= nv) dyv = 0.f;
          B0[(vy * 4 + i) * BD + tx + 16 * j] = dyv;
          if (v < nv) {
            dy_out[v * BD + tx + 16 * j] = dyv;
            s_bo[j] += dyv;
          }
        }
      }
    }
    __syncthreads();
    // ---- da = dy U^T ; dS = a (da - sum_q a da) per head
    {
      float da[4][J];
      zero_acc<J>(da);
      tile_mm<J>(B0, BD, Ut, HQP, BD, vy, tx, da);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < J; ++j) DAs[(vy * 4 + i) * HQP + tx + 16 * j] = da[i][j];
    }
    __syncthreads();
    for (int pidx = tid; pidx < BTV * heads; pidx += BTHREADS) {
      const int v = pidx / heads, h = pidx % heads;
      const float* ar = Aa + v * HQP + h * nq;
      float* dr = DAs + v * HQP + h * nq;
      float dot = 0.f;
      for (int q = 0; q < nq; ++q) dot = fmaf(ar[q], dr[q], dot);
      for (int q = 0; q < nq; ++q) dr[q] = ar[q] * (dr[q] - dot);
    }
    if (HQ < HQP)
      for (int idx = tid; idx < BTV * (HQP - HQ); idx += BTHREADS)
        DAs[(idx / (HQP - HQ)) * HQP + HQ + idx % (HQP - HQ)] = 0.f;
    __syncthreads();
    for (int idx = tid; idx < BTV * HQP; idx += BTHREADS) {
      const long long v = v0 + idx / HQP;
      if (v < nv) ds_out[v * HQP + idx % HQP] = DAs[idx];
    }
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) s_dc[j] += DAs[(vy * 4 + i) * HQP + tx + 16 * j];   // rows beyond nv hold zeros
    // ---- dx = dy + dS A
    {
      float o[4][8];
      zero_acc<8>(o);
      tile_mm<8>(DAs, HQP, A, BD, HQP, vy, tx, o);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long v = v0 + vy * 4 + i;
        if (v < nv) {
#pragma unroll
          for (int j = 0; j < 8; ++j) dx[v * BD + tx + 16 * j] = o[i][j] + B0[(vy * 4 + i) * BD + tx + 16 * j];
        }
      }
    }
  }
  // ---- per-CTA column sums: reduce over the 16 voxel groups through shared memory
  __syncthreads();
  constexpr int NCOL = 3 * BD + HQP;
  float* red = smem;   // 16 * NCOL floats <= BTV*BD*2 + ... (16 * (384 + 256) = 10240 floats)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[vy * NCOL + tx + 16 * j] = s_bo[j];
    red[vy * NCOL + BD + tx + 16 * j] = s_lnw[j];
    red[vy * NCOL + 2 * BD + tx + 16 * j] = s_lnb[j];
  }
#pragma unroll
  for (int j = 0; j < J; ++j) red[vy * NCOL + 3 * BD + tx + 16 * j] = s_dc[j];
  __syncthreads();
  for (int c = tid; c < NCOL; c += BTHREADS) {
    float s = 0.f;
    for (int i = 0; i < 16; ++i) s += red[i * NCOL + c];
    colpart[(long long)blockIdx.x * NCOL + c] = s;
  }
}

int launch_split_reduce(const float* part, int splits, long long count, long long total, int accumulate, float* dst,
                        cudaStream_t st);   // train_ops.cu

template <int J>
constexpr size_t c2s_bwd_smem() { return (size_t)(2 * BTV * BD + 2 * BTV * 16 * J) * sizeof(float); }
template <int J>
constexpr size_t s2c_bwd_smem() { return (size_t)(2 * BTV * BD + 2 * BTV * 16 * J + 2 * BTV * NQP) * sizeof(float); }

static int bwd_ctas(long long nv) {
  long long t = (nv + BTV - 1) / BTV;
  if (t > sm_count()) t = sm_count();
  return (int)(t < 1 ? 1 : t);
}

// Point-wise middle of the click -> scene attention backward when its four GEMMs run on the tensor cores
// (ops.c2s_attn_bwd_tc):  P = masked exp(S - lse),  dS = P * (dP - dr), in place (S -> P, dP -> dS).
// Same masking rule as c2s_bwd_kernel: a (row j, voxel v) pair is dead when row j is padding (rowobj -2) or restricted
// to an object other than the voxel's label.
__global__ void __launch_bounds__(256)
c2s_bwd_pointwise_kernel(float* __restrict__ s_p, float* __restrict__ dp_ds, const float* __restrict__ lse,
                         const float* __restrict__ dr, const int* __restrict__ rowobj,
                         const unsigned char* __restrict__ label, long long nv, int hq) {
  const int c4n = hq >> 2;
  const long long total = nv * c4n;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long v = t / c4n;
    const int c = (int)(t % c4n) * 4;
    const int lab = label ? (int)label[v] : 0;
    float4 s4 = *reinterpret_cast<const float4*>(s_p + v * hq + c);
    float4 d4 = *reinterpret_cast<const float4*>(dp_ds + v * hq + c);
    float sv[4] = {s4.x, s4.y, s4.z, s4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int ro = __ldg(rowobj + c + e);
      const bool dead = (ro == -2) || (ro >= 0 && lab != ro);
      const float pv = dead ? 0.f : __expf(sv[e] - __ldg(lse + c + e));
      sv[e] = pv;
      dv[e] = pv * (dv[e] - __ldg(dr + c + e));
    }
    *reinterpret_cast<float4*>(s_p + v * hq + c) = make_float4(sv[0], sv[1], sv[2], sv[3]);
    *reinterpret_cast<float4*>(dp_ds + v * hq + c) = make_float4(dv[0], dv[1], dv[2], dv[3]);
  }
}

// ---- point-wise / row-wise pieces of the scene -> click backward when its GEMMs run as 1x1 tensor-core convolutions
// (ops.s2c_mask_bwd_tc).  Same formulas as s2c_bwd_kernel.

// Per-(voxel, head) row operations on [nv, hqp] matrices, staged through shared memory so that global traffic is
// coalesced (a thread owns one (voxel, head) segment of nq values; rows are padded by one float against bank conflicts):
//   DS = false: softmax over the nq queries, in place on S
//   DS = true : dS = a * (da - sum_q a da), in place on da
// padding columns [heads*nq, hqp) are zeroed.
constexpr int ROWS_TV = 32;
template <bool DS>
__global__ void __launch_bounds__(256)
s2c_rows_kernel(const float* __restrict__ a, float* __restrict__ io, long long nv, int heads, int nq, int hqp) {
  extern __shared__ float sm[];
  const int ld = hqp + 1;
  float* t_io = sm;                       // [ROWS_TV][ld]
  float* t_a = sm + ROWS_TV * ld;         // [ROWS_TV][ld] (DS only)
  const int hq = heads * nq;
  const long long n_tiles = (nv + ROWS_TV - 1) / ROWS_TV;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long v0 = tile * ROWS_TV;
    const int rows = (int)min((long long)ROWS_TV, nv - v0);
    const int n4 = rows * hqp / 4;        // hqp % 4 == 0: the tile is one contiguous, 16-byte aligned span
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = *(reinterpret_cast<const float4*>(io + v0 * hqp) + i);
      const int r = (i * 4) / hqp, c = (i * 4) % hqp;
      float* d = t_io + r * ld + c;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
      if (DS) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(a + v0 * hqp) + i);
        float* e = t_a + r * ld + c;
        e[0] = w.x; e[1] = w.y; e[2] = w.z; e[3] = w.w;
      }
    }
    __syncthreads();
    for (int pidx = threadIdx.x; pidx < rows * heads; pidx += blockDim.x) {
      const int r = pidx / heads, h = pidx % heads;
      float* row = t_io + r * ld + h * nq;
      if (DS) {
        const float* ar = t_a + r * ld + h * nq;
        float dot = 0.f;
        for (int q = 0; q < nq; ++q) dot = fmaf(ar[q], row[q], dot);
        for (int q = 0; q < nq; ++q) row[q] = ar[q] * (row[q] - dot);
      } else {
        float mx = -INFINITY;
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
      if (h == 0)
        for (int c = hq; c < hqp; ++c) t_io[r * ld + c] = 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
      const int r = (i * 4) / hqp, c = (i * 4) % hqp;
      const float* d = t_io + r * ld + c;
      *(reinterpret_cast<float4*>(io + v0 * hqp) + i) = make_float4(d[0], d[1], d[2], d[3]);
    }
    __syncthreads();
  }
}

// LayerNorm statistics of y [nv, 128]: n = (y - mean) * rstd (written), rstd (written); warp per row, 4 channels per lane
__global__ void __launch_bounds__(256)
ln_fwd_stats_kernel(const float* __restrict__ y, long long nv, float eps, float* __restrict__ n_out,
                    float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long step = (long long)gridDim.x * (blockDim.x >> 5);
  for (; row < nv; row += step) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(y + row * BD) + lane);
    float sum = (v.x + v.y) + (v.z + v.w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.f / BD);
    const float4 d = make_float4(v.x - mean, v.y - mean, v.z - mean, v.w - mean);
    float sq = fmaf(d.x, d.x, fmaf(d.y, d.y, fmaf(d.z, d.z, d.w * d.w)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.f / BD) + eps);
    *(reinterpret_cast<float4*>(n_out + row * BD) + lane) = make_float4(d.x * rstd, d.y * rstd, d.z * rstd, d.w * rstd);
    if (lane == 0) rstd_out[row] = rstd;
  }
}

// LayerNorm backward: t = d x' (total), n, rstd -> dy = rstd * (t*lw - mean(t*lw) - n * mean(t*lw*n)); per-CTA column
// partials [dbo = sum dy | dln_w = sum t*n | dln_b = sum t] (lane l owns channels 4l..4l+3 of every row it visits)
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const float* __restrict__ t, const float* __restrict__ n, const float* __restrict__ rstd,
              const float* __restrict__ lw, long long nv, float* __restrict__ dy, float* __restrict__ colpart) {
  __shared__ float red[8][3 * BD];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4 w4 = __ldg(reinterpret_cast<const float4*>(lw) + lane);
  float s_bo[4] = {0.f, 0.f, 0.f, 0.f}, s_w[4] = {0.f, 0.f, 0.f, 0.f}, s_b[4] = {0.f, 0.f, 0.f, 0.f};
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  const long long step = (long long)gridDim.x * (blockDim.x >> 5);
  for (; row < nv; row += step) {
    const float4 tv = __ldg(reinterpret_cast<const float4*>(t + row * BD) + lane);
    const float4 nv4 = __ldg(reinterpret_cast<const float4*>(n + row * BD) + lane);
    const float r = __ldg(rstd + row);
    const float tt[4] = {tv.x, tv.y, tv.z, tv.w}, nn[4] = {nv4.x, nv4.y, nv4.z, nv4.w}, ww[4] = {w4.x, w4.y, w4.z, w4.w};
    float tw[4], m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s_b[e] += tt[e];
      s_w[e] = fmaf(tt[e], nn[e], s_w[e]);
      tw[e] = tt[e] * ww[e];
      m1 += tw[e];
      m2 = fmaf(tw[e], nn[e], m2);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m1 += __shfl_xor_sync(0xffffffffu, m1, o);
      m2 += __shfl_xor_sync(0xffffffffu, m2, o);
    }
    m1 *= (1.f / BD);
    m2 *= (1.f / BD);
    float d[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      d[e] = r * (tw[e] - m1 - nn[e] * m2);
      s_bo[e] += d[e];
    }
    *(reinterpret_cast<float4*>(dy + row * BD) + lane) = make_float4(d[0], d[1], d[2], d[3]);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    red[warp][4 * lane + e] = s_bo[e];
    red[warp][BD + 4 * lane + e] = s_w[e];
    red[warp][2 * BD + 4 * lane + e] = s_b[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 3 * BD; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w][c];
    colpart[(long long)blockIdx.x * (3 * BD) + c] = s;
  }
}

// first-maximum routing of the logit gradients: g[v][q] = dlogits[v][obj(q)] iff q is the first query of its object
// that attains the object's maximum of G[v][.] (models/agile3d.py:348-361 take the max over an object's clicks)
__global__ void __launch_bounds__(256)
s2c_route_kernel(const float* __restrict__ G, const float* __restrict__ dlogits, const int* __restrict__ q_obj, int nq,
                 int n_obj, long long nv, float* __restrict__ g_out) {
  __shared__ int qobj_s[NQP];
  if (threadIdx.x < NQP) qobj_s[threadIdx.x] = threadIdx.x < nq ? q_obj[threadIdx.x] : -1;
  __syncthreads();
  const long long total = nv * NQP;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long v = t / NQP;
    const int q = (int)(t % NQP);
    float g = 0.f;
    if (q < nq && dlogits) {
      const int o = qobj_s[q];
      const float mine = __ldg(G + v * NQP + q);
      bool first_max = true;
      for (int q2 = 0; q2 < nq; ++q2) {
        if (qobj_s[q2] != o || q2 == q) continue;
        const float other = __ldg(G + v * NQP + q2);
        if (other > mine || (other == mine && q2 < q)) first_max = false;
      }
      if (first_max) g = __ldg(dlogits + v * n_obj + o);
    }
    g_out[t] = g;
  }
}

// ---- the same routing for up to 256 queries per scene (training at the tail of the click protocol, engine.py:78-115):
// pass 1: arg[v][o] = first query of object o attaining max_q G[v][q]; pass 2: g[v][q] = dlogits[v][o(q)] iff q == arg[v][o(q)]
__global__ void __launch_bounds__(256)
s2c_route_arg_kernel(const float* __restrict__ G, int ld, const int* __restrict__ q_obj, int nq, int n_obj, long long nv,
                     int* __restrict__ arg) {
  __shared__ int qobj_s[256];
  if (threadIdx.x < 256) qobj_s[threadIdx.x] = threadIdx.x < nq ? q_obj[threadIdx.x] : -1;
  __syncthreads();
  const long long total = nv * n_obj;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long v = t / n_obj;
    const int o = (int)(t % n_obj);
    float best = -INFINITY;
    int a = -1;
    for (int q = 0; q < nq; ++q) {
      if (qobj_s[q] != o) continue;
      const float z = __ldg(G + v * ld + q);
      if (a < 0 || z > best) { best = z; a = q; }
    }
    arg[t] = a;
  }
}
__global__ void __launch_bounds__(256)
s2c_route_scatter_kernel(const int* __restrict__ arg, const float* __restrict__ dlogits, const int* __restrict__ q_obj, int nq,
                         int n_obj, long long nv, int ld, float* __restrict__ g_out) {
  const long long total = nv * ld;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long v = t / ld;
    const int q = (int)(t % ld);
    float g = 0.f;
    if (q < nq && dlogits) {
      const int o = __ldg(q_obj + q);
      if (__ldg(arg + v * n_obj + o) == q) g = __ldg(dlogits + v * n_obj + o);
    }
    g_out[t] = g;
  }
}

}  // namespace ag3d

using namespace ag3d;

extern "C" {

int32_t ag3d_decoder_bwd_rows(int32_t nq, int32_t heads) {
  const int hq = nq * heads;
  for (int J : {6, 8, 10, 12, 14, 16})
    if (hq <= 16 * J) return 16 * J;
  return 0;
}

static unsigned pw_blocks(long long work) {
  long long b = (work + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

int ag3d_s2c_softmax_heads(float* s_a, int64_t nv, int32_t heads, int32_t nq, int32_t hqp, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(s_a && nv > 0 && heads > 0 && nq > 0 && heads * nq <= hqp, "s2c_softmax_heads: arguments");
  AG3D_CHECK_ARG(hqp % 4 == 0 && hqp <= 256 && aligned16(s_a), "s2c_softmax_heads: hqp");
  const size_t smem = (size_t)ROWS_TV * (hqp + 1) * sizeof(float);
  s2c_rows_kernel<false><<<pw_blocks((nv + ROWS_TV - 1) / ROWS_TV * 256), 256, smem, as_stream(stream)>>>(nullptr, s_a, nv, heads,
                                                                                                       nq, hqp);
  AG3D_LAUNCH_CHECK("s2c_softmax_heads");
  return AG3D_OK;
}

int ag3d_s2c_ds(const float* a, float* da_ds, int64_t nv, int32_t heads, int32_t nq, int32_t hqp, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(a && da_ds && nv > 0 && heads > 0 && nq > 0 && heads * nq <= hqp, "s2c_ds: arguments");
  AG3D_CHECK_ARG(hqp % 4 == 0 && hqp <= 256 && aligned16(a) && aligned16(da_ds), "s2c_ds: hqp");
  const size_t smem = (size_t)2 * ROWS_TV * (hqp + 1) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(s2c_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * ROWS_TV * 257 * 4));
    attr = true;
  }
  s2c_rows_kernel<true><<<pw_blocks((nv + ROWS_TV - 1) / ROWS_TV * 256), 256, smem, as_stream(stream)>>>(a, da_ds, nv, heads, nq,
                                                                                                      hqp);
  AG3D_LAUNCH_CHECK("s2c_ds");
  return AG3D_OK;
}

int ag3d_ln_fwd_stats(const float* y, int64_t nv, float eps, float* n_out, float* rstd_out, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(y && n_out && rstd_out && nv > 0 && aligned16(y) && aligned16(n_out), "ln_fwd_stats: arguments");
  ln_fwd_stats_kernel<<<pw_blocks(nv * 32), 256, 0, as_stream(stream)>>>(y, nv, eps, n_out, rstd_out);
  AG3D_LAUNCH_CHECK("ln_fwd_stats");
  return AG3D_OK;
}

size_t ag3d_ln_bwd_workspace_bytes(void) { return (size_t)sm_count() * 4 * 3 * BD * sizeof(float); }

/* colsums: [dbo 128 | dln_w 128 | dln_b 128] */
int ag3d_ln_bwd(const float* t, const float* n, const float* rstd, const float* ln_w, int64_t nv, float* dy,
                float* colsums, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(t && n && rstd && ln_w && dy && colsums && nv > 0, "ln_bwd: arguments");
  AG3D_CHECK_ARG(aligned16(t) && aligned16(n) && aligned16(dy) && aligned16(ln_w), "ln_bwd: alignment");
  long long blocks = (nv + 7) / 8;
  if (blocks > (long long)sm_count() * 4) blocks = (long long)sm_count() * 4;
  AG3D_CHECK_ARG(ws && ws_bytes >= (size_t)blocks * 3 * BD * sizeof(float), "ln_bwd: workspace");
  cudaStream_t st = as_stream(stream);
  ln_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(t, n, rstd, ln_w, nv, dy, static_cast<float*>(ws));
  AG3D_LAUNCH_CHECK("ln_bwd");
  return launch_split_reduce(static_cast<const float*>(ws), (int)blocks, 3 * BD, 3 * BD, 0, colsums, st);
}

int ag3d_s2c_route(const float* G, const float* dlogits, const int32_t* q_obj, int32_t nq, int32_t n_obj, int64_t nv,
                   float* g_out, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(G && q_obj && g_out && nv > 0 && nq > 0 && nq <= NQP, "s2c_route: arguments");
  s2c_route_kernel<<<pw_blocks(nv * NQP), 256, 0, as_stream(stream)>>>(G, dlogits, q_obj, nq, n_obj, nv, g_out);
  AG3D_LAUNCH_CHECK("s2c_route");
  return AG3D_OK;
}

int ag3d_s2c_route_ld(const float* G, int32_t ld, const float* dlogits, const int32_t* q_obj, int32_t nq, int32_t n_obj,
                      int64_t nv, int32_t* arg_ws, float* g_out, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(G && q_obj && g_out && arg_ws && nv > 0 && nq > 0 && nq <= 256 && ld >= nq && n_obj >= 1 && n_obj <= 32,
                 "s2c_route_ld: arguments");
  s2c_route_arg_kernel<<<pw_blocks(nv * n_obj), 256, 0, as_stream(stream)>>>(G, ld, q_obj, nq, n_obj, nv, arg_ws);
  AG3D_LAUNCH_CHECK("s2c_route_arg");
  s2c_route_scatter_kernel<<<pw_blocks(nv * ld), 256, 0, as_stream(stream)>>>(arg_ws, dlogits, q_obj, nq, n_obj, nv, ld, g_out);
  AG3D_LAUNCH_CHECK("s2c_route_scatter");
  return AG3D_OK;
}

int ag3d_c2s_bwd_pointwise(float* s_p, float* dp_ds, const float* lse, const float* dr, const int32_t* rowobj,
                           const uint8_t* label, int64_t nv, int32_t hq, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(nv > 0 && hq >= 4 && hq % 4 == 0, "c2s_bwd_pointwise: shape");
  AG3D_CHECK_ARG(s_p && dp_ds && lse && dr && rowobj && aligned16(s_p) && aligned16(dp_ds), "c2s_bwd_pointwise: pointers");
  const long long total = nv * (hq / 4);
  long long blocks = (total + 255) / 256;
  if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
  c2s_bwd_pointwise_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(s_p, dp_ds, lse, dr, rowobj, label, nv, hq);
  AG3D_LAUNCH_CHECK("c2s_bwd_pointwise");
  return AG3D_OK;
}

int ag3d_c2s_attn_bwd(const float* x, const float* pos, int64_t nv, const float* qf, const float* qft,
                      const float* dctx, const float* dctxt, const float* lse, const float* dr, const int32_t* rowobj,
                      int32_t hqp, const uint8_t* label, float* dx, float* ds_out, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(nv > 0, "c2s_bwd: empty");
  AG3D_CHECK_ARG(x && pos && qf && qft && dctx && dctxt && lse && dr && rowobj && dx && ds_out, "c2s_bwd: pointers");
  AG3D_CHECK_ARG(aligned16(x) && aligned16(pos), "c2s_bwd: alignment");
  cudaStream_t st = as_stream(stream);
  const int grid = bwd_ctas(nv);
#define LAUNCH(JJ)                                                                                          \
  do {                                                                                                      \
    static bool attr = false;                                                                               \
    if (!attr) {                                                                                            \
      AG3D_CUDA(cudaFuncSetAttribute(c2s_bwd_kernel<JJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                     (int)c2s_bwd_smem<JJ>()));                                            \
      attr = true;                                                                                          \
    }                                                                                                       \
    c2s_bwd_kernel<JJ><<<grid, BTHREADS, c2s_bwd_smem<JJ>(), st>>>(x, pos, nv, qf, qft, dctx, dctxt, lse, dr, \
                                                                   rowobj, label, dx, ds_out);              \
  } while (0)
  switch (hqp) {
    case 96: LAUNCH(6); break;
    case 128: LAUNCH(8); break;
    case 160: LAUNCH(10); break;
    case 192: LAUNCH(12); break;
    case 224: LAUNCH(14); break;
    case 256: LAUNCH(16); break;
    default: AG3D_CHECK_ARG(false, "c2s_bwd: hqp must come from ag3d_decoder_bwd_rows");
  }
#undef LAUNCH
  AG3D_LAUNCH_CHECK("c2s_bwd");
  return AG3D_OK;
}

size_t ag3d_s2c_bwd_workspace_bytes(int32_t hqp) { return (size_t)sm_count() * (3 * BD + hqp) * sizeof(float); }

int ag3d_s2c_mask_bwd(const float* x, const float* pos, int64_t nv, const float* A, const float* At, const float* c,
                      const float* U, const float* Ut, const float* bo, const float* ln_w, const float* ln_b,
                      float ln_eps, const float* E, const float* Et, const int32_t* q_obj, int32_t nq, int32_t heads,
                      int32_t n_obj, int32_t hqp, const float* dxo, const float* dlogits, float* dx, float* a_out,
                      float* ds_out, float* dy_out, float* g_out, float* colsums, void* ws, size_t ws_bytes,
                      ag3d_stream_t stream) {
  AG3D_CHECK_ARG(nv > 0 && nq > 0 && nq <= NQP && heads == 8, "s2c_bwd: shape");
  AG3D_CHECK_ARG(x && pos && A && At && c && U && Ut && bo && ln_w && ln_b && E && Et && q_obj, "s2c_bwd: pointers");
  AG3D_CHECK_ARG(dx && a_out && ds_out && dy_out && g_out && colsums, "s2c_bwd: outputs");
  AG3D_CHECK_ARG(aligned16(x) && aligned16(pos), "s2c_bwd: alignment");
  AG3D_CHECK_ARG(ws && ws_bytes >= ag3d_s2c_bwd_workspace_bytes(hqp), "s2c_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int grid = bwd_ctas(nv);
  float* colpart = static_cast<float*>(ws);
#define LAUNCH(JJ)                                                                                             \
  do {                                                                                                         \
    static bool attr = false;                                                                                  \
    if (!attr) {                                                                                               \
      AG3D_CUDA(cudaFuncSetAttribute(s2c_bwd_kernel<JJ>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                     (int)s2c_bwd_smem<JJ>()));                                               \
      attr = true;                                                                                             \
    }                                                                                                          \
    s2c_bwd_kernel<JJ><<<grid, BTHREADS, s2c_bwd_smem<JJ>(), st>>>(x, pos, nv, A, At, c, U, Ut, bo, ln_w, ln_b, \
                                                                   ln_eps, E, Et, q_obj, nq, heads, n_obj, dxo, \
                                                                   dlogits, dx, a_out, ds_out, dy_out, g_out,  \
                                                                   colpart);                                   \
  } while (0)
  switch (hqp) {
    case 96: LAUNCH(6); break;
    case 128: LAUNCH(8); break;
    case 160: LAUNCH(10); break;
    case 192: LAUNCH(12); break;
    case 224: LAUNCH(14); break;
    case 256: LAUNCH(16); break;
    default: AG3D_CHECK_ARG(false, "s2c_bwd: hqp must come from ag3d_decoder_bwd_rows");
  }
#undef LAUNCH
  AG3D_LAUNCH_CHECK("s2c_bwd");
  const long long count = 3 * BD + hqp;
  return launch_split_reduce(colpart, grid, count, count, 0, colsums, st);
}

}  // extern "C"
