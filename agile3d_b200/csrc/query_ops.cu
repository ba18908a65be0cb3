// K11: the click-query side of one decoder layer, fused (models/agile3d.py:273-325 around the two voxel-streaming
// kernels; models/modules/attention_block.py:28-38 (click<->click self-attention), :86-98 (cross-attention wrappers),
// :151-155 (FFN); mask_embed_head / decoder_norm of Agile3d.mask_module, models/agile3d.py:342-347).
//
// The O(Nq) algebra of a layer is ~15 [Nq x 128] x [128 x 128] products, a 128 -> 1024 -> 128 FFN and an Nq x Nq
// attention: launch-latency bound as ~40 library calls, microseconds as three kernels (plus one that assembles the
// queries of a click round).  Every kernel runs one CTA per (QRB = 4 query rows, scene); the operands of a CTA live in
// shared memory, weights are read from a per-layer blob of PRE-TRANSPOSED matrices (ag3d_query_blob_floats, layout
// below) so that the weight reads of a warp are coalesced and L2 resident.  fp32 FFMA throughout: this side carries no
// measurable bytes or flops, only latency.
//   ag3d_query_init       queries / query positions of a click round: clicked voxel features + fourier(click xyz) + time
//                         encoding, learned background queries (agile3d.py:202-264)
//   ag3d_query_fold_c2s   qfold[(h,q),:] = Wk_h^T ((Wq_h (Q+qpos) + bq_h) / sqrt(dh))           -> ag3d_c2s_attn_fwd
//   ag3d_query_update_a   ctx -> per-head value projection, out-proj, residual, LayerNorm (c2s tail); q/k/v of c2c
//   ag3d_query_update_b   c2c attention + out-proj + LN, FFN + LN, the s2c folds (A, c, U) and the mask embeddings E
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace ag3d {

constexpr int QD = 128, QHEADS = 8, QDH = 16, QF = 1024;
constexpr int QRB = 4;                   // query rows per CTA (the kernels are bound by the LDS + FFMA stream of a CTA, so
                                         // few rows per CTA = more SMs at work; every CTA re-reads the L2-resident weights)
constexpr int RPT = QRB / 2;             // rows per thread
constexpr int QT = 256;                  // threads per CTA: thread = (output column o = tid & 127, row half tid >> 7)
constexpr int DD = QD * QD;

// ---- weight blob of one decoder layer (floats).  "T" = stored transposed, [in][out].
constexpr int O_C2S_WQT = 0;
constexpr int O_C2S_BQ = O_C2S_WQT + DD;
constexpr int O_C2S_WK = O_C2S_BQ + QD;          // [(h,d)][c] as stored in in_proj_weight[d:2d]
constexpr int O_C2S_WVT = O_C2S_WK + DD;
constexpr int O_C2S_BV = O_C2S_WVT + DD;
constexpr int O_C2S_WOT = O_C2S_BV + QD;
constexpr int O_C2S_BO = O_C2S_WOT + DD;
constexpr int O_C2S_LNW = O_C2S_BO + QD;
constexpr int O_C2S_LNB = O_C2S_LNW + QD;
constexpr int O_C2C_WQT = O_C2S_LNB + QD;
constexpr int O_C2C_BQ = O_C2C_WQT + DD;
constexpr int O_C2C_WKT = O_C2C_BQ + QD;
constexpr int O_C2C_BK = O_C2C_WKT + DD;
constexpr int O_C2C_WVT = O_C2C_BK + QD;
constexpr int O_C2C_BV = O_C2C_WVT + DD;
constexpr int O_C2C_WOT = O_C2C_BV + QD;
constexpr int O_C2C_BO = O_C2C_WOT + DD;
constexpr int O_C2C_LNW = O_C2C_BO + QD;
constexpr int O_C2C_LNB = O_C2C_LNW + QD;
constexpr int O_FFN_W1T = O_C2C_LNB + QD;        // [128][1024]
constexpr int O_FFN_B1 = O_FFN_W1T + QD * QF;
constexpr int O_FFN_W2T = O_FFN_B1 + QF;         // [1024][128]
constexpr int O_FFN_B2 = O_FFN_W2T + QF * QD;
constexpr int O_FFN_LNW = O_FFN_B2 + QD;
constexpr int O_FFN_LNB = O_FFN_LNW + QD;
constexpr int O_S2C_WKT = O_FFN_LNB + QD;
constexpr int O_S2C_BK = O_S2C_WKT + DD;
constexpr int O_S2C_WVT = O_S2C_BK + QD;
constexpr int O_S2C_BV = O_S2C_WVT + DD;
constexpr int O_S2C_WQ = O_S2C_BV + QD;          // [(h,d)][c] as stored in in_proj_weight[:d]
constexpr int O_S2C_BQ = O_S2C_WQ + DD;
constexpr int O_S2C_WOT = O_S2C_BQ + QD;         // [(h,d)][c] = out_proj.weight^T
constexpr int O_DEC_LNW = O_S2C_WOT + DD;
constexpr int O_DEC_LNB = O_DEC_LNW + QD;
constexpr int O_M1T = O_DEC_LNB + QD;
constexpr int O_M1B = O_M1T + DD;
constexpr int O_M2T = O_M1B + QD;
constexpr int O_M2B = O_M2T + DD;
constexpr int Q_BLOB_FLOATS = O_M2B + QD;

// acc[r] = sum_i xs[(rb + r) * ldx + i] * WT[i * ldw + o]   (r < RPT); xs in shared memory (warp-wide broadcast reads, four
// inputs per LDS.128), WT in global memory (coalesced across o, 16 weight loads in flight per thread)
template <int KDIM>
__device__ __forceinline__ void gemm_rows(const float* xs, int ldx, const float* __restrict__ WT, int ldw, int o, int rb,
                                          float* acc) {
#pragma unroll
  for (int r = 0; r < RPT; ++r) acc[r] = 0.f;
#pragma unroll 1
  for (int i0 = 0; i0 < KDIM; i0 += 16) {
    float w[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) w[u] = __ldg(WT + (size_t)(i0 + u) * ldw + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const float4* xr = reinterpret_cast<const float4*>(xs + (rb + r) * ldx + i0);
#pragma unroll
      for (int u4 = 0; u4 < 4; ++u4) {
        const float4 x4 = xr[u4];
        acc[r] = fmaf(x4.x, w[4 * u4], acc[r]);
        acc[r] = fmaf(x4.y, w[4 * u4 + 1], acc[r]);
        acc[r] = fmaf(x4.z, w[4 * u4 + 2], acc[r]);
        acc[r] = fmaf(x4.w, w[4 * u4 + 3], acc[r]);
      }
    }
  }
}

// LayerNorm of the QRB rows of v [QRB][128] in place (eps 1e-5, biased variance, as nn.LayerNorm); warp w < QRB: row w
__device__ __forceinline__ void layer_norm16(float* v, const float* __restrict__ w, const float* __restrict__ b, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < QRB) {
    float* row = v + warp * QD;
    float x[4], s = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) { x[e] = row[lane + 32 * e]; s += x[e]; }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / QD);
    float sq = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float d = x[e] - mean; sq = fmaf(d, d, sq); }
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = 1.f / sqrtf(sq * (1.f / QD) + eps);
#pragma unroll
    for (int e = 0; e < 4; ++e) row[lane + 32 * e] = (x[e] - mean) * rstd * __ldg(w + lane + 32 * e) + __ldg(b + lane + 32 * e);
  }
}

// ---------------------------------------------------------------------------------------------- query assembly
// row (scene b, query q): src >= 0 -> clicked voxel: fourier(xyz[src]) + time encoding, feature row feats[feat_row]
// (feat_row: the same voxel in the row order of `feats`; NULL = src); src < 0 -> learned background query -(src + 1).   grid = total query rows, 128 threads (thread = channel)
__global__ void query_init_kernel(const float* __restrict__ feats, const float* __restrict__ xyz,
                                  const float* __restrict__ range, const int* __restrict__ src_row,
                                  const int* __restrict__ feat_row, const int* __restrict__ time_idx, const int* __restrict__ scene_of_row,
                                  const float* __restrict__ gauss_B, const float* __restrict__ time_table,
                                  const float* __restrict__ bg_feat, const float* __restrict__ bg_pos,
                                  float* __restrict__ queries, float* __restrict__ qpos, int feats_split) {
  const int row = blockIdx.x, c = threadIdx.x;
  const int src = src_row[row];
  if (src < 0) {
    const int k = -(src + 1);
    queries[(size_t)row * QD + c] = bg_feat[k * QD + c];
    qpos[(size_t)row * QD + c] = bg_pos[k * QD + c];
    return;
  }
  const int b = scene_of_row[row];
  const int j = c & 63;
  float arg = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float mn = range[b * 6 + a], den = range[b * 6 + 3 + a] - mn;
    float u = (xyz[(size_t)src * 3 + a] - mn) / den;
    u *= 6.283185307179586f;
    arg = fmaf(u, __ldg(gauss_B + a * 64 + j), arg);
  }
  float s, co;
  sincosf(arg, &s, &co);
  const float* frow = feats + (size_t)(feat_row ? feat_row[row] : src) * QD;
  float f;
  if (feats_split) {      // "split" rows: 32-channel slab = 32 bf16 hi | 32 bf16 lo
    const unsigned short* hp = reinterpret_cast<const unsigned short*>(frow) + (c >> 5) * 64 + (c & 31);
    f = __uint_as_float((unsigned)hp[0] << 16) + __uint_as_float((unsigned)hp[32] << 16);
  } else {
    f = frow[c];
  }
  queries[(size_t)row * QD + c] = f;
  qpos[(size_t)row * QD + c] = (c < 64 ? s : co) + time_table[(size_t)time_idx[row] * QD + c];
}

// ---------------------------------------------------------------------------------------------- fold for c2s
__global__ void __launch_bounds__(QT) query_fold_c2s_kernel(const float* __restrict__ Q, const float* __restrict__ qpos,
                                                           const float* __restrict__ blob, int nq, float* __restrict__ qfold) {
  __shared__ __align__(16) float xs[QRB * QD], qp[QRB * QD];
  const int b = blockIdx.y, r0 = blockIdx.x * QRB, tid = threadIdx.x, o = tid & 127, rb = (tid >> 7) * RPT;
  const float* Qb = Q + (size_t)b * nq * QD;
  const float* Pb = qpos + (size_t)b * nq * QD;
  for (int i = tid; i < QRB * QD; i += QT) {
    const int q = r0 + i / QD;
    xs[i] = q < nq ? Qb[(size_t)q * QD + (i % QD)] + Pb[(size_t)q * QD + (i % QD)] : 0.f;
  }
  __syncthreads();
  float acc[RPT];
  gemm_rows<QD>(xs, QD, blob + O_C2S_WQT, QD, o, rb, acc);
  const float bq = __ldg(blob + O_C2S_BQ + o);
#pragma unroll
  for (int r = 0; r < RPT; ++r) qp[(rb + r) * QD + o] = (acc[r] + bq) * 0.25f;       // 1 / sqrt(16)
  __syncthreads();
  float* out = qfold + (size_t)b * QHEADS * nq * QD;
  for (int h = 0; h < QHEADS; ++h) {
    gemm_rows<QDH>(qp + h * QDH, QD, blob + O_C2S_WK + (size_t)h * QDH * QD, QD, o, rb, acc);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int q = r0 + rb + r;
      if (q < nq) out[((size_t)h * nq + q) * QD + o] = acc[r];
    }
  }
}

// ---------------------------------------------------------------------------------------------- c2s tail + c2c projections
__global__ void __launch_bounds__(QT) query_update_a_kernel(const float* __restrict__ ctx, const float* __restrict__ Q,
                                                           const float* __restrict__ qpos, const float* __restrict__ blob,
                                                           int nq, float ln_eps, float* __restrict__ q1,
                                                           float* __restrict__ Qh, float* __restrict__ Kh,
                                                           float* __restrict__ Vh) {
  extern __shared__ float sm[];
  float* cs = sm;                         // [8 heads][16 rows][128]
  float* xs = cs + QHEADS * QRB * QD;     // [16][128]
  float* ys = xs + QRB * QD;              // [16][128]
  const int b = blockIdx.y, r0 = blockIdx.x * QRB, tid = threadIdx.x, o = tid & 127, rb = (tid >> 7) * RPT;
  const float* cb = ctx + (size_t)b * QHEADS * nq * QD;
  for (int i = tid; i < QHEADS * QRB * QD; i += QT) {
    const int c = i % QD, r = (i / QD) % QRB, h = i / (QD * QRB), q = r0 + r;
    cs[i] = q < nq ? cb[((size_t)h * nq + q) * QD + c] : 0.f;
  }
  __syncthreads();
  float acc[RPT];
  // per-head value projection: heads[q][(h,d)] = sum_c ctx[(h,q)][c] Wv[(h,d)][c] + bv
  gemm_rows<QD>(cs + (o >> 4) * QRB * QD, QD, blob + O_C2S_WVT, QD, o, rb, acc);
  {
    const float bv = __ldg(blob + O_C2S_BV + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) xs[(rb + r) * QD + o] = acc[r] + bv;
  }
  __syncthreads();
  gemm_rows<QD>(xs, QD, blob + O_C2S_WOT, QD, o, rb, acc);
  {
    const float bo = __ldg(blob + O_C2S_BO + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int q = r0 + rb + r;
      const float res = q < nq ? Q[((size_t)b * nq + q) * QD + o] : 0.f;
      ys[(rb + r) * QD + o] = res + (acc[r] + bo);
    }
  }
  __syncthreads();
  layer_norm16(ys, blob + O_C2S_LNW, blob + O_C2S_LNB, ln_eps);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int q = r0 + rb + r;
    const float v = ys[(rb + r) * QD + o];
    const float pp = q < nq ? qpos[((size_t)b * nq + q) * QD + o] : 0.f;
    if (q < nq) q1[((size_t)b * nq + q) * QD + o] = v;
    xs[(rb + r) * QD + o] = v + pp;
  }
  __syncthreads();
  const float* wts[3] = {blob + O_C2C_WQT, blob + O_C2C_WKT, blob + O_C2C_WVT};
  const float* bss[3] = {blob + O_C2C_BQ, blob + O_C2C_BK, blob + O_C2C_BV};
  float* outs[3] = {Qh, Kh, Vh};
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    gemm_rows<QD>(m == 2 ? ys : xs, QD, wts[m], QD, o, rb, acc);
    const float bias = __ldg(bss[m] + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int q = r0 + rb + r;
      if (q < nq) outs[m][((size_t)b * nq + q) * QD + o] = acc[r] + bias;
    }
  }
}

// ---------------------------------------------------------------------------------------------- c2c attention, FFN, s2c folds, E
__global__ void __launch_bounds__(QT) query_update_b_kernel(const float* __restrict__ q1, const float* __restrict__ Qh,
                                                           const float* __restrict__ Kh, const float* __restrict__ Vh,
                                                           const float* __restrict__ qpos, const float* __restrict__ blob,
                                                           int nq, float ln_eps, float* __restrict__ q3, float* __restrict__ A,
                                                           float* __restrict__ cvec, float* __restrict__ U,
                                                           float* __restrict__ E) {
  extern __shared__ float sm[];
  float* b0 = sm;                          // four [16][128] buffers
  float* b1 = b0 + QRB * QD;
  float* b2 = b1 + QRB * QD;
  float* b3 = b2 + QRB * QD;
  float* hid = b3 + QRB * QD;              // [16][1024]
  const int b = blockIdx.y, r0 = blockIdx.x * QRB, tid = threadIdx.x, o = tid & 127, rb = (tid >> 7) * RPT;
  const int warp = tid >> 5, lane = tid & 31;
  const size_t base = (size_t)b * nq * QD;
  // ---- click <-> click attention: warp = head, rows one after the other, lanes over the keys (nq <= 256: 8 per lane)
  {
    const int h = warp;
    for (int r = 0; r < QRB; ++r) {
      const int q = r0 + r;
      if (q >= nq) {
        if (lane < QDH) b0[r * QD + h * QDH + lane] = 0.f;
        continue;
      }
      float qv[QDH];
      {
        const float4* src = reinterpret_cast<const float4*>(Qh + base + (size_t)q * QD + h * QDH);
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float4 t = __ldg(src + e); qv[4 * e] = t.x; qv[4 * e + 1] = t.y; qv[4 * e + 2] = t.z; qv[4 * e + 3] = t.w; }
      }
      float s[8], mx = -INFINITY;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int j = lane + 32 * m;
        s[m] = -INFINITY;
        if (j < nq) {
          const float4* src = reinterpret_cast<const float4*>(Kh + base + (size_t)j * QD + h * QDH);
          float d = 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 t = __ldg(src + e);
            d = fmaf(qv[4 * e], t.x, d); d = fmaf(qv[4 * e + 1], t.y, d); d = fmaf(qv[4 * e + 2], t.z, d); d = fmaf(qv[4 * e + 3], t.w, d);
          }
          s[m] = d * 0.25f;
        }
        mx = fmaxf(mx, s[m]);
      }
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      float sum = 0.f, ov[QDH];
#pragma unroll
      for (int d = 0; d < QDH; ++d) ov[d] = 0.f;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int j = lane + 32 * m;
        if (j < nq) {
          const float e = expf(s[m] - mx);
          sum += e;
          const float4* src = reinterpret_cast<const float4*>(Vh + base + (size_t)j * QD + h * QDH);
#pragma unroll
          for (int e4 = 0; e4 < 4; ++e4) {
            const float4 t = __ldg(src + e4);
            ov[4 * e4] = fmaf(e, t.x, ov[4 * e4]); ov[4 * e4 + 1] = fmaf(e, t.y, ov[4 * e4 + 1]);
            ov[4 * e4 + 2] = fmaf(e, t.z, ov[4 * e4 + 2]); ov[4 * e4 + 3] = fmaf(e, t.w, ov[4 * e4 + 3]);
          }
        }
      }
      for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
#pragma unroll
        for (int d = 0; d < QDH; ++d) ov[d] += __shfl_xor_sync(0xffffffffu, ov[d], off);
      }
      const float inv = 1.f / sum;
#pragma unroll
      for (int d = 0; d < QDH; ++d)
        if (lane == d) b0[r * QD + h * QDH + d] = ov[d] * inv;
    }
  }
  __syncthreads();
  float acc[RPT];
  // ---- out-proj + residual + LayerNorm -> q2 (b1)
  gemm_rows<QD>(b0, QD, blob + O_C2C_WOT, QD, o, rb, acc);
  {
    const float bo = __ldg(blob + O_C2C_BO + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int q = r0 + rb + r;
      b1[(rb + r) * QD + o] = (q < nq ? q1[base + (size_t)q * QD + o] : 0.f) + (acc[r] + bo);
    }
  }
  __syncthreads();
  layer_norm16(b1, blob + O_C2C_LNW, blob + O_C2C_LNB, ln_eps);
  __syncthreads();
  // ---- FFN: hid = relu(q2 W1^T + b1), q3 = LN(q2 + hid W2^T + b2) (b2 buffer)
  for (int cbk = 0; cbk < QF / QD; ++cbk) {
    gemm_rows<QD>(b1, QD, blob + O_FFN_W1T + cbk * QD, QF, o, rb, acc);
    const float bias = __ldg(blob + O_FFN_B1 + cbk * QD + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) hid[(rb + r) * QF + cbk * QD + o] = fmaxf(acc[r] + bias, 0.f);
  }
  __syncthreads();
  gemm_rows<QF>(hid, QF, blob + O_FFN_W2T, QD, o, rb, acc);
  {
    const float bias = __ldg(blob + O_FFN_B2 + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) b2[(rb + r) * QD + o] = b1[(rb + r) * QD + o] + (acc[r] + bias);
  }
  __syncthreads();
  layer_norm16(b2, blob + O_FFN_LNW, blob + O_FFN_LNB, ln_eps);
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RPT; ++r) {
    const int q = r0 + rb + r;
    const float v = b2[(rb + r) * QD + o];
    if (q < nq) q3[base + (size_t)q * QD + o] = v;
    b0[(rb + r) * QD + o] = v + (q < nq ? qpos[base + (size_t)q * QD + o] : 0.f);      // q3 + qpos
  }
  __syncthreads();
  // ---- s2c folds: kp = (q3+qpos) Wk^T + bk (b1), vp = q3 Wv^T + bv (b3)
  gemm_rows<QD>(b0, QD, blob + O_S2C_WKT, QD, o, rb, acc);
  {
    const float bias = __ldg(blob + O_S2C_BK + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) b1[(rb + r) * QD + o] = acc[r] + bias;
  }
  gemm_rows<QD>(b2, QD, blob + O_S2C_WVT, QD, o, rb, acc);
  {
    const float bias = __ldg(blob + O_S2C_BV + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) b3[(rb + r) * QD + o] = acc[r] + bias;
  }
  __syncthreads();
  {
    float* Ab = A + (size_t)b * QHEADS * nq * QD;
    float* Ub = U + (size_t)b * QHEADS * nq * QD;
    for (int h = 0; h < QHEADS; ++h) {
      gemm_rows<QDH>(b1 + h * QDH, QD, blob + O_S2C_WQ + (size_t)h * QDH * QD, QD, o, rb, acc);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int q = r0 + rb + r;
        if (q < nq) Ab[((size_t)h * nq + q) * QD + o] = acc[r] * 0.25f;
      }
      gemm_rows<QDH>(b3 + h * QDH, QD, blob + O_S2C_WOT + (size_t)h * QDH * QD, QD, o, rb, acc);
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int q = r0 + rb + r;
        if (q < nq) Ub[((size_t)h * nq + q) * QD + o] = acc[r];
      }
    }
    if (tid < QRB * QHEADS) {
      const int r = tid >> 3, h = tid & 7, q = r0 + r;
      if (q < nq) {
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < QDH; ++e) d = fmaf(b1[r * QD + h * QDH + e], __ldg(blob + O_S2C_BQ + h * QDH + e), d);
        cvec[(size_t)b * QHEADS * nq + (size_t)h * nq + q] = d * 0.25f;
      }
    }
  }
  __syncthreads();
  // ---- mask embeddings: E = W2 relu(W1 LN_dec(q3) + b1) + b2   (b2 holds q3)
  layer_norm16(b2, blob + O_DEC_LNW, blob + O_DEC_LNB, ln_eps);
  __syncthreads();
  gemm_rows<QD>(b2, QD, blob + O_M1T, QD, o, rb, acc);
  {
    const float bias = __ldg(blob + O_M1B + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) b0[(rb + r) * QD + o] = fmaxf(acc[r] + bias, 0.f);
  }
  __syncthreads();
  gemm_rows<QD>(b0, QD, blob + O_M2T, QD, o, rb, acc);
  {
    const float bias = __ldg(blob + O_M2B + o);
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const int q = r0 + rb + r;
      if (q < nq) E[base + (size_t)q * QD + o] = acc[r] + bias;
    }
  }
}

}  // namespace ag3d

using namespace ag3d;

extern "C" {

int64_t ag3d_query_blob_floats(void) { return Q_BLOB_FLOATS; }

int ag3d_query_init(const float* feats, const float* xyz, const float* range, const int32_t* src_row,
                    const int32_t* feat_row, const int32_t* time_idx, const int32_t* scene_of_row, int32_t n_rows, const float* gauss_B,
                    const float* time_table, const float* bg_feat, const float* bg_pos, float* queries, float* qpos,
                    int32_t feats_split, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_rows > 0, "no query rows");
  AG3D_CHECK_ARG(feats && xyz && range && src_row && time_idx && scene_of_row && gauss_B && time_table && bg_feat &&
                     bg_pos && queries && qpos, "bad pointers");
  query_init_kernel<<<n_rows, QD, 0, as_stream(stream)>>>(feats, xyz, range, src_row, feat_row, time_idx, scene_of_row, gauss_B,
                                                          time_table, bg_feat, bg_pos, queries, qpos, feats_split);
  AG3D_LAUNCH_CHECK("query_init");
  return AG3D_OK;
}

int ag3d_query_fold_c2s(const float* queries, const float* qpos, const float* blob, int32_t n_scenes, int32_t nq,
                        float* qfold, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_scenes > 0 && nq > 0 && nq <= 256, "1..256 queries per scene");
  AG3D_CHECK_ARG(queries && qpos && blob && qfold, "bad pointers");
  query_fold_c2s_kernel<<<dim3((nq + QRB - 1) / QRB, n_scenes), QT, 0, as_stream(stream)>>>(queries, qpos, blob, nq, qfold);
  AG3D_LAUNCH_CHECK("query_fold_c2s");
  return AG3D_OK;
}

int ag3d_query_update_a(const float* ctx, const float* queries, const float* qpos, const float* blob, int32_t n_scenes,
                        int32_t nq, float ln_eps, float* q1, float* qh, float* kh, float* vh, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_scenes > 0 && nq > 0 && nq <= 256, "1..256 queries per scene");
  AG3D_CHECK_ARG(ctx && queries && qpos && blob && q1 && qh && kh && vh, "bad pointers");
  const size_t smem = (size_t)(QHEADS + 2) * QRB * QD * sizeof(float);
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(query_update_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  query_update_a_kernel<<<dim3((nq + QRB - 1) / QRB, n_scenes), QT, smem, as_stream(stream)>>>(ctx, queries, qpos, blob, nq,
                                                                                             ln_eps, q1, qh, kh, vh);
  AG3D_LAUNCH_CHECK("query_update_a");
  return AG3D_OK;
}

int ag3d_query_update_b(const float* q1, const float* qh, const float* kh, const float* vh, const float* qpos,
                        const float* blob, int32_t n_scenes, int32_t nq, float ln_eps, float* q3, float* A, float* c,
                        float* U, float* E, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_scenes > 0 && nq > 0 && nq <= 256, "1..256 queries per scene");
  AG3D_CHECK_ARG(q1 && qh && kh && vh && qpos && blob && q3 && A && c && U && E, "bad pointers");
  AG3D_CHECK_ARG(aligned16(qh) && aligned16(kh) && aligned16(vh), "projections must be 16-byte aligned");
  const size_t smem = (size_t)(4 * QRB * QD + QRB * QF) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(query_update_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  query_update_b_kernel<<<dim3((nq + QRB - 1) / QRB, n_scenes), QT, smem, as_stream(stream)>>>(q1, qh, kh, vh, qpos, blob, nq,
                                                                                             ln_eps, q3, A, c, U, E);
  AG3D_LAUNCH_CHECK("query_update_b");
  return AG3D_OK;
}

}  // extern "C"
