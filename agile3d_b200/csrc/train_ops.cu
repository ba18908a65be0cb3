// Training-step kernels: batch-statistics BatchNorm forward/backward, the
// weight-gradient contraction of the sparse convolution (also used for every "X^T dY" parameter gradient of the
// decoder), the stem's weight gradient against the hash table, the click-weighted CE + dice loss, the click
// loss-weight map, and the clip + AdamW update over flat parameter buffers.  All fp32, HBM-bound except the
// weight-gradient GEMM (FFMA).  Reductions are two-stage (per-CTA partials, then a deterministic final pass).
#include <math.h>

#include "common.cuh"

namespace ag3d {

// ============================================================================================ column reductions
// MODE 0: (sum z, sum z^2)            -> batch mean / inverse std (+ running-stat update)
// MODE 1: (sum g, sum g*xhat)         -> dbeta, dgamma, with g = dy * (y > 0 if relu), xhat = (z - mean) * invstd
// MODE 2: (sum z, -)                  -> plain column sum (bias gradients)
constexpr int CR_THREADS = 256;

template <int MODE>
__global__ void __launch_bounds__(CR_THREADS)
colreduce_kernel(const float* __restrict__ z, int z_ld, const float* __restrict__ y, int y_ld,
                 const float* __restrict__ dy, int dy_ld, const float* __restrict__ mean,
                 const float* __restrict__ invstd, int C, long long n, int relu, float* __restrict__ part) {
  __shared__ float red[2][CR_THREADS * 4];
  const int cg = C >> 2;               // float4 groups per row
  const int R = CR_THREADS / cg;       // row lanes
  const int tid = threadIdx.x;
  const int rl = tid / cg, c4 = tid % cg;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  if (rl < R) {
    float4 mu = s0, is = s0;
    if (MODE == 1) {
      mu = __ldg(reinterpret_cast<const float4*>(mean) + c4);
      is = __ldg(reinterpret_cast<const float4*>(invstd) + c4);
    }
    // MODE 0 sums z - z[0] (shifted data): E[(z-K)^2] - E[z-K]^2 has no cancellation when K is one of the samples
    if (MODE == 0) mu = __ldg(reinterpret_cast<const float4*>(z) + c4);
    for (long long r = (long long)blockIdx.x * R + rl; r < n; r += (long long)gridDim.x * R) {
      const float4 zv = __ldg(reinterpret_cast<const float4*>(z + r * z_ld) + c4);
      if (MODE == 0) {
        const float dx = zv.x - mu.x, dy_ = zv.y - mu.y, dz_ = zv.z - mu.z, dw = zv.w - mu.w;
        s0.x += dx; s0.y += dy_; s0.z += dz_; s0.w += dw;
        s1.x = fmaf(dx, dx, s1.x); s1.y = fmaf(dy_, dy_, s1.y);
        s1.z = fmaf(dz_, dz_, s1.z); s1.w = fmaf(dw, dw, s1.w);
      } else if (MODE == 2) {
        s0.x += zv.x; s0.y += zv.y; s0.z += zv.z; s0.w += zv.w;
      } else {
        float4 g = __ldg(reinterpret_cast<const float4*>(dy + r * dy_ld) + c4);
        if (relu) {
          const float4 yv = __ldg(reinterpret_cast<const float4*>(y + r * y_ld) + c4);
          g.x = yv.x > 0.f ? g.x : 0.f; g.y = yv.y > 0.f ? g.y : 0.f;
          g.z = yv.z > 0.f ? g.z : 0.f; g.w = yv.w > 0.f ? g.w : 0.f;
        }
        s0.x += g.x; s0.y += g.y; s0.z += g.z; s0.w += g.w;
        s1.x = fmaf(g.x, (zv.x - mu.x) * is.x, s1.x); s1.y = fmaf(g.y, (zv.y - mu.y) * is.y, s1.y);
        s1.z = fmaf(g.z, (zv.z - mu.z) * is.z, s1.z); s1.w = fmaf(g.w, (zv.w - mu.w) * is.w, s1.w);
      }
    }
    float* r0 = &red[0][(rl * cg + c4) * 4];
    float* r1 = &red[1][(rl * cg + c4) * 4];
    r0[0] = s0.x; r0[1] = s0.y; r0[2] = s0.z; r0[3] = s0.w;
    r1[0] = s1.x; r1[1] = s1.y; r1[2] = s1.z; r1[3] = s1.w;
  }
  __syncthreads();
  for (int c = tid; c < C; c += CR_THREADS) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < R; ++i) {
      a += red[0][i * C + c];
      b += red[1][i * C + c];
    }
    part[((long long)blockIdx.x * 2 + 0) * C + c] = a;
    part[((long long)blockIdx.x * 2 + 1) * C + c] = b;
  }
}

// fp64 sum of the per-CTA partials, deterministic: a CTA owns 32 channels, thread (g, c) adds the partials g, g + 16,
// ... of its channel (coalesced over c), and the 16 group sums are folded in a fixed order by the first 32 threads
constexpr int CF_GROUPS = 16;
__global__ void __launch_bounds__(32 * CF_GROUPS)
colreduce_final_kernel(const float* __restrict__ part, int n_cta, int C, long long n, int mode,
                       const float* __restrict__ shift_row, float eps, float momentum,
                       float* __restrict__ running_mean, float* __restrict__ running_var,
                       float* __restrict__ out0, float* __restrict__ out1) {
  __shared__ double sa[CF_GROUPS][32], sb[CF_GROUPS][32];
  const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  double a = 0.0, b = 0.0;
  if (c < C)
    for (int i = grp; i < n_cta; i += CF_GROUPS) {
      a += (double)part[((long long)i * 2 + 0) * C + c];
      b += (double)part[((long long)i * 2 + 1) * C + c];
    }
  sa[grp][cl] = a;
  sb[grp][cl] = b;
  __syncthreads();
  if (grp != 0 || c >= C) return;
  a = 0.0;
  b = 0.0;
  for (int g2 = 0; g2 < CF_GROUPS; ++g2) {
    a += sa[g2][cl];
    b += sb[g2][cl];
  }
  if (mode == 0) {
    const double ms = a / (double)n;                 // mean of the shifted data
    double var = b / (double)n - ms * ms;
    if (var < 0.0) var = 0.0;
    const double m = ms + (double)shift_row[c];
    out0[c] = (float)m;
    out1[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
      const double unb = n > 1 ? var * (double)n / (double)(n - 1) : var;
      running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
      running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unb);
    }
  } else {
    out0[c] = (float)a;
    if (out1) out1[c] = (float)b;
  }
}

// y = act((z - mean) * invstd * gamma + beta (+ residual))
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ z, int z_ld, const float* __restrict__ mean,
                const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                const float* __restrict__ res, int res_ld, int C, long long n, int relu, float* __restrict__ y,
                int y_ld) {
  const int cg = C >> 2;
  const long long total = n * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cg;
    const int c4 = (int)(i % cg);
    const float4 zv = __ldg(reinterpret_cast<const float4*>(z + r * z_ld) + c4);
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c4);
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + c4);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    float4 o;
    o.x = fmaf((zv.x - mu.x) * is.x, ga.x, be.x);
    o.y = fmaf((zv.y - mu.y) * is.y, ga.y, be.y);
    o.z = fmaf((zv.z - mu.z) * is.z, ga.z, be.z);
    o.w = fmaf((zv.w - mu.w) * is.w, ga.w, be.w);
    if (res) {
      const float4 rv = __ldg(reinterpret_cast<const float4*>(res + r * res_ld) + c4);
      o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
    }
    if (relu) {
      o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
    }
    *(reinterpret_cast<float4*>(y + r * y_ld) + c4) = o;
  }
}

// dz = gamma * invstd * (g - sum_g / n - xhat * sum_gx / n);  g (relu-masked dy) optionally written for the
// residual branch.  dz may alias dy.
__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ z, int z_ld, const float* __restrict__ y, int y_ld, const float* dy,
                    int dy_ld, const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ gamma, const float* __restrict__ sum_g,
                    const float* __restrict__ sum_gx, int C, long long n, int relu, float* dz, int dz_ld,
                    float* g_out, int g_ld) {
  const int cg = C >> 2;
  const long long total = n * cg;
  const float inv_n = 1.f / (float)n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cg;
    const int c4 = (int)(i % cg);
    const float4 zv = __ldg(reinterpret_cast<const float4*>(z + r * z_ld) + c4);
    float4 g = *(reinterpret_cast<const float4*>(dy + r * dy_ld) + c4);
    if (relu) {
      const float4 yv = __ldg(reinterpret_cast<const float4*>(y + r * y_ld) + c4);
      g.x = yv.x > 0.f ? g.x : 0.f; g.y = yv.y > 0.f ? g.y : 0.f;
      g.z = yv.z > 0.f ? g.z : 0.f; g.w = yv.w > 0.f ? g.w : 0.f;
    }
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c4);
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + c4);
    const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
    const float4 sg = __ldg(reinterpret_cast<const float4*>(sum_g) + c4);
    const float4 sx = __ldg(reinterpret_cast<const float4*>(sum_gx) + c4);
    float4 o;
    o.x = ga.x * is.x * (g.x - sg.x * inv_n - (zv.x - mu.x) * is.x * sx.x * inv_n);
    o.y = ga.y * is.y * (g.y - sg.y * inv_n - (zv.y - mu.y) * is.y * sx.y * inv_n);
    o.z = ga.z * is.z * (g.z - sg.z * inv_n - (zv.z - mu.z) * is.z * sx.z * inv_n);
    o.w = ga.w * is.w * (g.w - sg.w * inv_n - (zv.w - mu.w) * is.w * sx.w * inv_n);
    if (g_out) *(reinterpret_cast<float4*>(g_out + r * g_ld) + c4) = g;
    *(reinterpret_cast<float4*>(dz + r * dz_ld) + c4) = o;
  }
}

// ============================================================================================ weight gradient
// C_z[m][n] = sum_i A[idx_z(i)][m] * B[i][n]   (idx_z(i) = nbr[z][i], rows with idx < 0 contribute nothing;
// nbr == NULL: identity).  Grid (tiles_m * tiles_n, splits, K); per-split partial sums, reduced deterministically.
constexpr int TG_BM = 64, TG_BN = 64, TG_BK = 16, TG_THREADS = 256;

__global__ void __launch_bounds__(TG_THREADS)
tn_gemm_kernel(const float* __restrict__ A, int lda, int M, const float* __restrict__ B, int ldb, int N,
               const int* __restrict__ idx, long long idx_stride, long long rows, long long rows_per_split,
               int tiles_n, float* __restrict__ part) {
  __shared__ __align__(16) float As[TG_BK][TG_BM];
  __shared__ __align__(16) float Bs[TG_BK][TG_BN];
  __shared__ int s_row[TG_THREADS], s_src[TG_THREADS], s_wcnt[TG_THREADS / 32];
  const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
  const int split = blockIdx.y, z = blockIdx.z;
  const int m0 = tm * TG_BM, n0 = tn * TG_BN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  const int lr = tid >> 4, lc = (tid & 15) * 4;
  const int warp = tid >> 5, lane = tid & 31;
  const int* idz = idx ? idx + (long long)z * idx_stride : nullptr;
  const long long r0 = (long long)split * rows_per_split;
  const long long r1 = min(rows, r0 + rows_per_split);
  const bool a_ok = m0 + lc < M, b_ok = n0 + lc < N;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long r = r0; r < r1; r += TG_THREADS) {
    // ---- compact the (output row, input row) pairs of this 256-row window: offsets without a neighbour cost nothing
    const long long row = r + tid;
    int src = -1;
    if (row < r1) src = idz ? __ldg(idz + row) : (int)row;
    const unsigned bal = __ballot_sync(0xffffffffu, src >= 0);
    if (lane == 0) s_wcnt[warp] = __popc(bal);
    __syncthreads();                       // also: the previous window's last chunk is fully consumed
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < TG_THREADS / 32; ++w) {
      const int c = s_wcnt[w];
      if (w < warp) base += c;
      total += c;
    }
    if (src >= 0) {
      const int p = base + __popc(bal & ((1u << lane) - 1u));
      s_row[p] = (int)(row - r);
      s_src[p] = src;
    }
    __syncthreads();
    if (total == 0) continue;
    // ---- 16 pairs per step, next step's rows prefetched into registers while the current one is multiplied
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = av;
    if (lr < total) {
      if (a_ok) av = __ldg(reinterpret_cast<const float4*>(A + (long long)s_src[lr] * lda + m0 + lc));
      if (b_ok) bv = __ldg(reinterpret_cast<const float4*>(B + (r + s_row[lr]) * ldb + n0 + lc));
    }
    for (int c = 0; c < total; c += TG_BK) {
      *reinterpret_cast<float4*>(&As[lr][lc]) = av;
      *reinterpret_cast<float4*>(&Bs[lr][lc]) = bv;
      __syncthreads();
      av = make_float4(0.f, 0.f, 0.f, 0.f);
      bv = av;
      const int nx = c + TG_BK + lr;
      if (nx < total) {
        if (a_ok) av = __ldg(reinterpret_cast<const float4*>(A + (long long)s_src[nx] * lda + m0 + lc));
        if (b_ok) bv = __ldg(reinterpret_cast<const float4*>(B + (r + s_row[nx]) * ldb + n0 + lc));
      }
#pragma unroll
      for (int kk = 0; kk < TG_BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float ar[4] = {a.x, a.y, a.z, a.w};
        const float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  float* dst = part + ((long long)z * gridDim.y + split) * M * N;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int n = n0 + tx * 4;
    if (n < N) *reinterpret_cast<float4*>(dst + (long long)m * N + n) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  }
}

// dst[z][e] = (accumulate ? dst[z][e] : 0) + sum_s part[z][s][e]
__global__ void split_reduce_kernel(const float* __restrict__ part, int splits, long long count, long long total,
                                    int accumulate, float* __restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long z = i / count, e = i % count;
    const float* p = part + z * splits * count + e;
    float s = accumulate ? dst[i] : 0.f;
    for (int k = 0; k < splits; ++k) s += p[(long long)k * count];
    dst[i] = s;
  }
}

int launch_split_reduce(const float* part, int splits, long long count, long long total, int accumulate, float* dst,
                        cudaStream_t st) {
  long long b = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (b > cap) b = cap;
  split_reduce_kernel<<<(unsigned)b, 256, 0, st>>>(part, splits, count, total, accumulate, dst);
  AG3D_LAUNCH_CHECK("split_reduce");
  return AG3D_OK;
}

static int tn_gemm_splits(long long rows, int M, int N, int nz, int* tiles_m, int* tiles_n, long long* rps) {
  *tiles_m = (M + TG_BM - 1) / TG_BM;
  *tiles_n = (N + TG_BN - 1) / TG_BN;
  const long long base = (long long)(*tiles_m) * (*tiles_n) * nz;
  const long long target = (long long)sm_count() * 8;
  long long splits = (target + base - 1) / base;
  const long long max_splits = (rows + 511) / 512;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long per = (rows + splits - 1) / splits;
  per = (per + TG_THREADS - 1) / TG_THREADS * TG_THREADS;
  splits = (rows + per - 1) / per;
  *rps = per;
  return (int)splits;
}

// ============================================================================================ stem weight gradient
// dW[k][ci][co] += f[src_k(v)][ci] * dz[v][co]: warp per voxel, lane = output channel, probes as in the forward.
template <bool BRICKS>
__global__ void __launch_bounds__(256)
stem_wgrad_kernel(const int4* __restrict__ coords, const float* __restrict__ feats, long long n,
                  const Slot* __restrict__ table, unsigned long long mask, const int* __restrict__ brick_rows, int ksize,
                  const float* __restrict__ dz, int dz_ld, float* __restrict__ part) {
  extern __shared__ float dw_s[];  // [K][3][32]
  const int K = ksize * ksize * ksize;
  for (int i = threadIdx.x; i < K * 96; i += blockDim.x) dw_s[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int half = ksize / 2;
  long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long step = (long long)gridDim.x * (blockDim.x >> 5);
  for (; row < n; row += step) {
    const int4 c = __ldg(coords + row);
    const float g = __ldg(dz + row * dz_ld + lane);
    // probes and neighbour features of all (<= 4 x 32) offsets are fetched before the serial accumulation (as in
    // stem_conv_kernel): no global load on the dependent chain
    int src[4];
    float f0[4], f1[4], f2[4];
    if constexpr (BRICKS) {
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
          const int x = c.y + jx - half, y = c.z + jy - half, zz = c.w + r - half;
          if (coord_in_range(c.x, x, y, zz)) src[j] = table_find(table, mask, pack_key(c.x, x, y, zz));
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      f0[j] = f1[j] = f2[j] = 0.f;
      if (src[j] >= 0) {
        const float* f = feats + (long long)src[j] * 3;
        f0[j] = __ldg(f + 0); f1[j] = __ldg(f + 1); f2[j] = __ldg(f + 2);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      unsigned hits = __ballot_sync(0xffffffffu, src[j] >= 0);
      while (hits) {
        const int b = __ffs(hits) - 1;
        hits &= hits - 1;
        const float a0 = __shfl_sync(0xffffffffu, f0[j], b);
        const float a1 = __shfl_sync(0xffffffffu, f1[j], b);
        const float a2 = __shfl_sync(0xffffffffu, f2[j], b);
        float* w = dw_s + (j * 32 + b) * 96 + lane;
        atomicAdd(w, a0 * g);
        atomicAdd(w + 32, a1 * g);
        atomicAdd(w + 64, a2 * g);
      }
    }
  }
  __syncthreads();
  float* dst = part + (long long)blockIdx.x * K * 96;
  for (int i = threadIdx.x; i < K * 96; i += blockDim.x) dst[i] = dw_s[i];
}

// ============================================================================================ loss
// Per voxel (models/criterion.py:84-111 with multiclass_dice_loss 15-75 taken over the class axis):
//   ce_v   = logsumexp(l_v) - l_v[t_v]
//   dice_v = num > eps ? 1 - (num + eps) / (2/C + eps) : 0,   num = 2 p_v[t_v] / C
//   loss_bce = mean_v(w_v ce_v),  loss_dice = mean_v(w_v dice_v)
// sums[0] += sum_v w_v ce_v, sums[1] += sum_v w_v dice_v  (per-CTA partials, reduced by colreduce_final_kernel).
constexpr int LOSS_MAXC = 32;

__global__ void __launch_bounds__(256)
loss_fwd_kernel(const float* __restrict__ logits, int C, long long n, const int* __restrict__ target,
                const float* __restrict__ w, float eps, float* __restrict__ part) {
  __shared__ float red[2][8];
  float s_ce = 0.f, s_dice = 0.f;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
    const float* l = logits + v * C;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(l + c));
    float sum = 0.f;
    for (int c = 0; c < C; ++c) sum += expf(__ldg(l + c) - mx);
    const int t = __ldg(target + v);
    if ((unsigned)t >= (unsigned)C) {     // the reference asserts 0 <= class id < C (models/criterion.py: multiclass_dice_loss);
      s_ce = s_dice = __int_as_float(0x7fc00000);   // here an invalid label (e.g. the ignore id -1) poisons the loss with NaN
      continue;
    }
    const float lt = __ldg(l + t);
    const float lse = mx + logf(sum);
    const float pt = expf(lt - lse);
    const float wv = __ldg(w + v);
    const float num = 2.f * pt / (float)C;
    const float den = 2.f / (float)C;
    s_ce += wv * (lse - lt);
    s_dice += wv * (num > eps ? 1.f - (num + eps) / (den + eps) : 0.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s_ce += __shfl_xor_sync(0xffffffffu, s_ce, o);
    s_dice += __shfl_xor_sync(0xffffffffu, s_dice, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_ce; red[1][threadIdx.x >> 5] = s_dice; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float a = 0.f;
    for (int i = 0; i < 8; ++i) a += red[threadIdx.x][i];
    part[(long long)blockIdx.x * 4 + threadIdx.x * 2] = a;   // laid out as [cta][2][C=2]: column 0 of each sum
    part[(long long)blockIdx.x * 4 + threadIdx.x * 2 + 1] = 0.f;
  }
}

// dlogits[v][c] (+)= w_v/n * ( g_bce * (p_c - [c == t]) + g_dice * ddice/dl_c ),
// ddice/dl_c = -(2/C)/(2/C + eps) * p_t ([c == t] - p_c) when num > eps.  g = (g_bce, g_dice) on the device.
__global__ void __launch_bounds__(256)
loss_bwd_kernel(const float* __restrict__ logits, int C, long long n, const int* __restrict__ target,
                const float* __restrict__ w, float eps, const float* __restrict__ g, float* __restrict__ dlogits) {
  const float g_bce = __ldg(g), g_dice = __ldg(g + 1);
  const float inv_n = 1.f / (float)n;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
    const float* l = logits + v * C;
    float p[LOSS_MAXC];
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(l + c));
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < LOSS_MAXC; ++c)
      if (c < C) { p[c] = expf(__ldg(l + c) - mx); sum += p[c]; }
    const float inv = 1.f / sum;
    const int t = __ldg(target + v);
    const bool bad = (unsigned)t >= (unsigned)C;       // invalid label: NaN gradients (see loss_fwd_kernel)
    float pt = 0.f;
#pragma unroll
    for (int c = 0; c < LOSS_MAXC; ++c)
      if (c < C) { p[c] *= inv; if (c == t) pt = p[c]; }
    const float num = 2.f * pt / (float)C, den = 2.f / (float)C;
    const float kd = (num > eps) ? -(den / (den + eps)) * pt * g_dice : 0.f;
    const float wv = __ldg(w + v) * inv_n;
#pragma unroll
    for (int c = 0; c < LOSS_MAXC; ++c)
      if (c < C) {
        const float ind = (c == t) ? 1.f : 0.f;
        dlogits[v * C + c] = bad ? __int_as_float(0x7fc00000) : wv * (g_bce * (p[c] - ind) + kd * (ind - p[c]));
      }
  }
}

// utils/seg.py:62-89: w_v = alpha + (beta - alpha) * (1 - min(d_v, tita) / tita), d_v = distance to the nearest click
__global__ void __launch_bounds__(256)
click_weights_kernel(const float* __restrict__ xyz, long long n, const float* __restrict__ clicks, int n_clicks,
                     float alpha, float beta, float tita, float* __restrict__ w) {
  extern __shared__ float ck[];
  for (int i = threadIdx.x; i < n_clicks * 3; i += blockDim.x) ck[i] = __ldg(clicks + i);
  __syncthreads();
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x) {
    const float x = __ldg(xyz + v * 3), y = __ldg(xyz + v * 3 + 1), z = __ldg(xyz + v * 3 + 2);
    float best = INFINITY;
    for (int i = 0; i < n_clicks; ++i) {
      const float dx = x - ck[i * 3], dy = y - ck[i * 3 + 1], dz = z - ck[i * 3 + 2];
      best = fminf(best, dx * dx + dy * dy + dz * dz);
    }
    const float d = sqrtf(best);
    w[v] = alpha + (beta - alpha) * (1.f - fminf(d, tita) / tita);
  }
}

// ============================================================================================ optimizer
__global__ void __launch_bounds__(256)
sqnorm_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ part) {
  __shared__ float red[8];
  float s = 0.f;
  const long long n4 = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (long long i = n4 << 2; i < n; ++i) s = fmaf(g[i], g[i], s);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int i = 0; i < 8; ++i) a += red[i];
    part[blockIdx.x] = a;
  }
}

__global__ void sqnorm_final_kernel(const float* __restrict__ part, int n_cta, float* __restrict__ norm_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0;
    for (int i = 0; i < n_cta; ++i) a += (double)part[i];
    norm_out[0] = (float)sqrt(a);
  }
}

// torch.nn.utils.clip_grad_norm_ (engine.py:148-149) + torch.optim.AdamW (main.py) in one pass over the flat buffers
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             long long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2,
             const float* __restrict__ norm, float max_norm) {
  float coef = 1.f;
  if (norm && max_norm > 0.f) coef = fminf(1.f, max_norm / (__ldg(norm) + 1e-6f));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * mi / denom;
    p[i] = pi;
  }
}

static inline int grid_for(long long work_items, int threads, int per_sm) {
  long long b = (work_items + threads - 1) / threads;
  const long long cap = (long long)sm_count() * per_sm;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

static int colreduce_ctas(long long n, int C) {
  const int R = CR_THREADS / (C / 4);
  long long b = (n + (long long)R * 32 - 1) / ((long long)R * 32);   // >= 32 rows per row lane
  const long long cap = (long long)sm_count() * 4;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace ag3d

using namespace ag3d;

extern "C" {

size_t ag3d_colreduce_workspace_bytes(int32_t C) { return (size_t)sm_count() * 4 * 2 * (size_t)(C > 0 ? C : 1) * sizeof(float); }

static int colreduce_common(int mode, const float* z, int z_ld, const float* y, int y_ld, const float* dy, int dy_ld,
                            const float* mean, const float* invstd, int C, long long n, int relu, float eps,
                            float momentum, float* rm, float* rv, float* out0, float* out1, void* ws, size_t ws_bytes,
                            cudaStream_t st) {
  AG3D_CHECK_ARG(n > 0 && C >= 4 && C % 4 == 0 && C <= 1024, "colreduce: C must be a multiple of 4, <= 1024");
  AG3D_CHECK_ARG(z && aligned16(z) && z_ld % 4 == 0 && z_ld >= C, "colreduce: bad z");
  AG3D_CHECK_ARG(ws && ws_bytes >= ag3d_colreduce_workspace_bytes(C), "colreduce: workspace too small");
  const int n_cta = colreduce_ctas(n, C);
  float* part = static_cast<float*>(ws);
  if (mode == 0)
    colreduce_kernel<0><<<n_cta, CR_THREADS, 0, st>>>(z, z_ld, y, y_ld, dy, dy_ld, mean, invstd, C, n, relu, part);
  else if (mode == 1)
    colreduce_kernel<1><<<n_cta, CR_THREADS, 0, st>>>(z, z_ld, y, y_ld, dy, dy_ld, mean, invstd, C, n, relu, part);
  else
    colreduce_kernel<2><<<n_cta, CR_THREADS, 0, st>>>(z, z_ld, y, y_ld, dy, dy_ld, mean, invstd, C, n, relu, part);
  AG3D_LAUNCH_CHECK("colreduce");
  colreduce_final_kernel<<<(C + 31) / 32, 32 * CF_GROUPS, 0, st>>>(part, n_cta, C, n, mode, z, eps, momentum, rm, rv, out0, out1);
  AG3D_LAUNCH_CHECK("colreduce_final");
  return AG3D_OK;
}

int ag3d_bn_stats(const float* z, int32_t z_ld, int32_t C, int64_t n, float eps, float momentum, float* running_mean,
                  float* running_var, float* mean, float* invstd, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(mean && invstd, "bn_stats: outputs");
  return colreduce_common(0, z, z_ld, nullptr, 0, nullptr, 0, nullptr, nullptr, C, n, 0, eps, momentum, running_mean,
                          running_var, mean, invstd, ws, ws_bytes, as_stream(stream));
}

int ag3d_col_sum(const float* z, int32_t z_ld, int32_t C, int64_t n, float* sum, void* ws, size_t ws_bytes,
                 ag3d_stream_t stream) {
  AG3D_CHECK_ARG(sum, "col_sum: output");
  return colreduce_common(2, z, z_ld, nullptr, 0, nullptr, 0, nullptr, nullptr, C, n, 0, 0.f, 0.f, nullptr, nullptr,
                          sum, nullptr, ws, ws_bytes, as_stream(stream));
}

int ag3d_bn_apply(const float* z, int32_t z_ld, const float* mean, const float* invstd, const float* gamma,
                  const float* beta, const float* residual, int32_t res_ld, int32_t C, int64_t n, int32_t flags,
                  float* y, int32_t y_ld, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n > 0 && C >= 4 && C % 4 == 0, "bn_apply: shape");
  AG3D_CHECK_ARG(z && y && mean && invstd && gamma && beta && aligned16(z) && aligned16(y), "bn_apply: pointers");
  AG3D_CHECK_ARG(z_ld % 4 == 0 && y_ld % 4 == 0 && (!residual || (res_ld % 4 == 0 && aligned16(residual))), "bn_apply: ld");
  bn_apply_kernel<<<grid_for(n * (C / 4), 256, 8), 256, 0, as_stream(stream)>>>(
      z, z_ld, mean, invstd, gamma, beta, residual, res_ld, C, n, flags & AG3D_RELU, y, y_ld);
  AG3D_LAUNCH_CHECK("bn_apply");
  return AG3D_OK;
}

int ag3d_bn_bwd(const float* z, int32_t z_ld, const float* y, int32_t y_ld, const float* dy, int32_t dy_ld,
                const float* mean, const float* invstd, const float* gamma, int32_t C, int64_t n, int32_t flags,
                float* dz, int32_t dz_ld, float* g_out, int32_t g_ld, float* dgamma, float* dbeta, void* ws,
                size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(dy && dz && dgamma && dbeta && mean && invstd && gamma, "bn_bwd: pointers");
  AG3D_CHECK_ARG(aligned16(dy) && aligned16(dz) && dy_ld % 4 == 0 && dz_ld % 4 == 0, "bn_bwd: alignment");
  const int relu = flags & AG3D_RELU;
  AG3D_CHECK_ARG(!relu || (y && aligned16(y) && y_ld % 4 == 0), "bn_bwd: relu needs y");
  AG3D_CHECK_ARG(!g_out || (aligned16(g_out) && g_ld % 4 == 0), "bn_bwd: g_out");
  cudaStream_t st = as_stream(stream);
  if (int rc = colreduce_common(1, z, z_ld, y, y_ld, dy, dy_ld, mean, invstd, C, n, relu, 0.f, 0.f, nullptr, nullptr,
                                dbeta, dgamma, ws, ws_bytes, st))
    return rc;
  bn_bwd_apply_kernel<<<grid_for(n * (C / 4), 256, 8), 256, 0, st>>>(z, z_ld, y, y_ld, dy, dy_ld, mean, invstd, gamma,
                                                                      dbeta, dgamma, C, n, relu, dz, dz_ld, g_out, g_ld);
  AG3D_LAUNCH_CHECK("bn_bwd_apply");
  return AG3D_OK;
}

size_t ag3d_spconv_bwd_weight_workspace_bytes(int64_t n_out, int32_t K, int32_t cin, int32_t cout) {
  if (n_out <= 0 || K < 1 || cin < 4 || cout < 4) return 0;
  int tm, tn;
  long long rps;
  const int splits = tn_gemm_splits(n_out, cin, cout, K, &tm, &tn, &rps);
  return (size_t)splits * K * cin * cout * sizeof(float);
}

int ag3d_spconv_bwd_weight(const float* in, int32_t in_ld, int32_t cin, const int32_t* nbr, int32_t K, int64_t n_out,
                           const float* dout, int32_t dout_ld, int32_t cout, float* dweight, int32_t accumulate,
                           void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_out > 0 && n_out < 2147483647LL, "bwd_weight: row count");
  AG3D_CHECK_ARG(cin >= 4 && cin % 4 == 0 && cout >= 4 && cout % 4 == 0, "bwd_weight: channels must be multiples of 4");
  AG3D_CHECK_ARG(K >= 1 && (nbr || K == 1), "bwd_weight: K > 1 needs a neighbour table");
  AG3D_CHECK_ARG(in && dout && dweight && aligned16(in) && aligned16(dout) && aligned16(dweight), "bwd_weight: pointers");
  AG3D_CHECK_ARG(in_ld % 4 == 0 && dout_ld % 4 == 0 && in_ld >= cin && dout_ld >= cout, "bwd_weight: leading dims");
  int tm, tn;
  long long rps;
  const int splits = tn_gemm_splits(n_out, cin, cout, K, &tm, &tn, &rps);
  AG3D_CHECK_ARG(ws && aligned16(ws) && ws_bytes >= (size_t)splits * K * cin * cout * sizeof(float),
                 "bwd_weight: workspace too small");
  cudaStream_t st = as_stream(stream);
  float* part = static_cast<float*>(ws);
  tn_gemm_kernel<<<dim3(tm * tn, splits, K), TG_THREADS, 0, st>>>(in, in_ld, cin, dout, dout_ld, cout, nbr, n_out, n_out,
                                                                  rps, tn, part);
  AG3D_LAUNCH_CHECK("tn_gemm");
  const long long count = (long long)cin * cout, total = count * K;
  return launch_split_reduce(part, splits, count, total, accumulate, dweight, st);
}

size_t ag3d_stem_bwd_weight_workspace_bytes(int32_t ksize) {
  return (size_t)sm_count() * 2 * (size_t)ksize * ksize * ksize * 96 * sizeof(float);
}

static int stem_wgrad_launch(const int32_t* coords, const float* feats, int64_t n, const void* table, int64_t cap,
                             const int32_t* brick_rows, int32_t ksize, const float* dz, int32_t dz_ld, float* dweight,
                             int32_t accumulate, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n > 0 && n < 2147483647LL, "stem_bwd_weight: row count");
  AG3D_CHECK_ARG(ksize == 1 || ksize == 3 || ksize == 5, "stem kernel size must be 1, 3 or 5");
  AG3D_CHECK_ARG(coords && aligned16(coords) && feats && dz && dweight, "stem_bwd_weight: pointers");
  AG3D_CHECK_ARG(table && aligned16(table) && cap >= 2 && (cap & (cap - 1)) == 0, "bad hash table");
  AG3D_CHECK_ARG(ws && ws_bytes >= ag3d_stem_bwd_weight_workspace_bytes(ksize), "stem_bwd_weight: workspace too small");
  const int K = ksize * ksize * ksize;
  const size_t smem = (size_t)K * 96 * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    AG3D_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    AG3D_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr_done = true;
  }
  cudaStream_t st = as_stream(stream);
  long long blocks = (n + 7) / 8;
  const long long capb = (long long)sm_count() * 2;
  if (blocks > capb) blocks = capb;
  float* part = static_cast<float*>(ws);
  if (brick_rows)
    stem_wgrad_kernel<true><<<(unsigned)blocks, 256, smem, st>>>(reinterpret_cast<const int4*>(coords), feats, n,
                                                                 static_cast<const Slot*>(table),
                                                                 (unsigned long long)(cap - 1), brick_rows, ksize, dz,
                                                                 dz_ld, part);
  else
    stem_wgrad_kernel<false><<<(unsigned)blocks, 256, smem, st>>>(reinterpret_cast<const int4*>(coords), feats, n,
                                                                  static_cast<const Slot*>(table),
                                                                  (unsigned long long)(cap - 1), nullptr, ksize, dz,
                                                                  dz_ld, part);
  AG3D_LAUNCH_CHECK("stem_wgrad");
  const long long count = (long long)K * 96;
  return launch_split_reduce(part, (int)blocks, count, count, accumulate, dweight, st);
}

int ag3d_stem_bwd_weight(const int32_t* coords, const float* feats, int64_t n, const void* table, int64_t cap,
                         int32_t ksize, const float* dz, int32_t dz_ld, float* dweight, int32_t accumulate, void* ws,
                         size_t ws_bytes, ag3d_stream_t stream) {
  return stem_wgrad_launch(coords, feats, n, table, cap, nullptr, ksize, dz, dz_ld, dweight, accumulate, ws, ws_bytes,
                           stream);
}

int ag3d_stem_bwd_weight_bricks(const int32_t* coords, const float* feats, int64_t n, const void* table2, int64_t cap2,
                                const int32_t* brick_rows, int32_t ksize, const float* dz, int32_t dz_ld, float* dweight,
                                int32_t accumulate, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(brick_rows && (ksize == 3 || ksize == 5), "brick stem: brick_rows and kernel size 3 or 5");
  return stem_wgrad_launch(coords, feats, n, table2, cap2, brick_rows, ksize, dz, dz_ld, dweight, accumulate, ws, ws_bytes,
                           stream);
}

size_t ag3d_loss_workspace_bytes(void) { return (size_t)sm_count() * 4 * 4 * sizeof(float); }

int ag3d_loss_fwd(const float* logits, int32_t C, int64_t n, const int32_t* target, const float* w, float eps,
                  float* sums, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n > 0 && C >= 1 && C <= LOSS_MAXC, "loss: 1 <= classes <= 32");
  AG3D_CHECK_ARG(logits && target && w && sums, "loss: pointers");
  AG3D_CHECK_ARG(ws && ws_bytes >= ag3d_loss_workspace_bytes(), "loss: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int n_cta = grid_for(n, 256, 4);
  float* part = static_cast<float*>(ws);
  loss_fwd_kernel<<<n_cta, 256, 0, st>>>(logits, C, n, target, w, eps, part);
  AG3D_LAUNCH_CHECK("loss_fwd");
  // partials are [cta][2][2]; the final kernel sums column c of row 0 / row 1 -> out0[c], out1[c]; only c = 0 is used
  colreduce_final_kernel<<<1, 32 * CF_GROUPS, 0, st>>>(part, n_cta, 2, n, 2, nullptr, 0.f, 0.f, nullptr, nullptr, sums, sums + 2);
  AG3D_LAUNCH_CHECK("loss_final");
  return AG3D_OK;
}

int ag3d_loss_bwd(const float* logits, int32_t C, int64_t n, const int32_t* target, const float* w, float eps,
                  const float* g, float* dlogits, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n > 0 && C >= 1 && C <= LOSS_MAXC, "loss: 1 <= classes <= 32");
  AG3D_CHECK_ARG(logits && target && w && g && dlogits, "loss: pointers");
  loss_bwd_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(logits, C, n, target, w, eps, g, dlogits);
  AG3D_LAUNCH_CHECK("loss_bwd");
  return AG3D_OK;
}

int ag3d_click_loss_weights(const float* xyz, int64_t n, const float* clicks, int32_t n_clicks, float alpha, float beta,
                            float tita, float* w, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n > 0 && n_clicks >= 1 && n_clicks <= 4096, "click_loss_weights: 1..4096 clicks");
  AG3D_CHECK_ARG(xyz && clicks && w, "click_loss_weights: pointers");
  click_weights_kernel<<<grid_for(n, 256, 8), 256, (size_t)n_clicks * 3 * sizeof(float), as_stream(stream)>>>(
      xyz, n, clicks, n_clicks, alpha, beta, tita, w);
  AG3D_LAUNCH_CHECK("click_weights");
  return AG3D_OK;
}

size_t ag3d_grad_norm_workspace_bytes(void) { return (size_t)sm_count() * 8 * sizeof(float); }

int ag3d_grad_norm(const float* g, int64_t n, float* norm_out, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(g && n > 0 && norm_out && aligned16(g), "grad_norm: pointers");
  AG3D_CHECK_ARG(ws && ws_bytes >= ag3d_grad_norm_workspace_bytes(), "grad_norm: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int n_cta = grid_for(n / 4 + 1, 256, 8);
  sqnorm_partial_kernel<<<n_cta, 256, 0, st>>>(g, n, static_cast<float*>(ws));
  AG3D_LAUNCH_CHECK("sqnorm_partial");
  sqnorm_final_kernel<<<1, 32, 0, st>>>(static_cast<const float*>(ws), n_cta, norm_out);
  AG3D_LAUNCH_CHECK("sqnorm_final");
  return AG3D_OK;
}

int ag3d_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int32_t step, const float* grad_norm, float max_norm,
                    ag3d_stream_t stream) {
  AG3D_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "adamw: arguments");
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adamw_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                   bc1, bc2, grad_norm, max_norm);
  AG3D_LAUNCH_CHECK("adamw");
  return AG3D_OK;
}

}  // extern "C"
