// K4, "pair-packed" variant: sparse convolution whose tensor-core tiles hold only (input row, output row) PAIRS.
//
//   out[o, :] = act( scale * sum_k in[nbr[k][o], :] @ W[k] + shift (+ residual[o, :]) )
//
// Why.  The output-stationary kernel (spconv_tc.cu) multiplies, for every kernel offset k, a dense tile of 128
// consecutive output rows although on surface-like voxel clouds only ~37 % of them have a neighbour at k (9.9 of 27
// at 2 cm, measured on the synthetic ScanNet-shape scenes; no row order changes that - tools/order_analysis.py: Morton
// order leaves 26.8 of 27 (tile, k) stages non-empty).  Absent rows cost the same TMA gather slots and the same MMAs
// as present ones.  Here the M = 128 rows of an MMA are the PRESENT pairs of a 256-row "super tile":
//   * a super tile = 4 segments of 64 consecutive output rows; segment q feeds TMEM lane quarter q: for offset k the
//     rows of the segment that have a neighbour are compacted (order preserved) into lanes 32q .. 32q+31 of pass 0 and,
//     if there are more than 32, of pass 1 (mean 24 of 32 lanes filled, i.e. 1.8x fewer MMAs / gathers than dense);
//   * every (k, pass) "group" gets a fresh fp32 accumulator D[128 x Cout] in TMEM (4 in flight); the accumulators of
//     the OUTPUT rows live in registers of the 8 owner warps: warp (q, h) owns the 64 rows of segment q x column half
//     h, lane l = rows 64q+l and 64q+32+l.  When a group completes, the owner warps read their lane quarter with
//     tcgen05.ld and route pair lane -> owner lane with shfl.idx (the rank of a row among the present rows of its
//     segment is a popc of a ballot), adding into their registers.  Reading D per group instead of per tile is the
//     price of packing; it runs beside the MMAs of the next groups (tools/probes/epi_probe.cu: ~840 cycles per Cout = 96
//     group on 8 warps vs 864 cycles of MMA).
// Roles (640 threads, one CTA per SM, persistent over super tiles, setmaxnreg moves registers to the owner warps):
//   warps 0-7   owners / epilogue (184 registers)          warps 8-15  TMA tile::gather4 producers (40 registers)
//   warp 16     tcgen05.mma issuer                         warp 17     weight stages (cp.async.bulk)
//   warps 18-19 planners: compact the neighbour table of the NEXT super tile into shared-memory pair lists,
//               per-row offset masks and per-offset pass counts (double buffered)
// Operand formats are those of spconv_tc.cu MODE 2: split (bf16 hi|lo) input rows gathered into SWIZZLE_128B tiles,
// pre-split weight stage images, bf16x3 products (hi*hi + hi*lo + lo*hi).
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace ag3d {

bool make_row_map(CUtensorMap* tm, const float* in, int in_ld, int cin, long long n_in);   // spconv_tc.cu
extern int g_last_tmap_rc;

constexpr int PK_R = 256;                 // output rows per super tile
constexpr int PK_SEG = 64;                // rows per segment (one TMEM lane quarter)
constexpr int PK_E_WARPS = 8;
constexpr int PK_P_WARPS = 8;
constexpr int PK_WARP_MMA = 16, PK_WARP_W = 17, PK_WARP_PLAN = 18;
constexpr int PK_THREADS = 20 * 32;
constexpr int PK_ND = 4;                  // accumulator buffers (groups in flight between the MMA issuer and the owners)
constexpr int PK_MAX_NA = 8, PK_MAX_NB = 8;
constexpr uint32_t PK_STAGE = 16384;      // [128 pair rows x 128 B] gathered slab, SWIZZLE_128B
constexpr int PK_BAR_BYTES = 512;
// setmaxnreg budgets (owners, producers, issuer group): 256 E + 256 P + 128 M registers-per-thread must equal 61440
template <int COUT> struct PkRegs { static constexpr int E = 160, P = 48, M = 64; };       // 40960 + 12288 + 8192
template <> struct PkRegs<128> { static constexpr int E = 176, P = 40, M = 48; };           // 45056 + 10240 + 6144
constexpr int PK_REGS_unused = 0;    // 256*176 + 256*40 + 128*48 = 61440 = 640 threads x the 96 registers the CTA is launched with (setmaxnreg only redistributes that pool)

struct PkParams {
  const int* nbr; int K; long long n_out;
  const uint4* wp; int cin;
  const float* scale; const float* shift; const float* residual; int res_ld;
  float* out; int out_ld; int flags; int out_split, res_split;
  int NA, NB;
  int n_super;
  long long* prof;  // measurement aid (AG3D_PK_PROF): per-role wait cycles of CTA 0, [role 4][slot 8]
  int debug;        // measurement aid (AG3D_PK_DEBUG): 1 no gathers, 2 no weight copies, 4 no MMAs, 8 no accumulator read-out, 16 no stores, 32 no accumulator hand-shake, 64 no weight ring, 128 synthetic plan (no neighbour-table loads)
  uint32_t plan_bytes;     // one plan buffer: lists [K][4][64] i32 | cnt [K][4] i32 (512 B) | npass [K] i32 (128 B) | mask [256] u32
  uint32_t off_b, off_a;   // shared-memory offsets of the weight ring and (before 1024-byte alignment) the A ring
};

__device__ __forceinline__ void pk_gather4(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int col, int4 r) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(col), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w)
      : "memory");
}
__device__ __forceinline__ void pk_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// mbarrier wait without a function call: ptxas cannot allocate a kernel that changes its register count
// (setmaxnreg) around an ABI call such as the shared mbar_wait_slow()
// try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes (or ~2 us pass) instead of
// spinning.  20 warps share four schedulers here and most of them wait most of the time: hot spin loops took the issue
// slots of the one warp per scheduler that had work (the MMA issuer ran ~3x slower than its instruction stream).
__device__ __forceinline__ uint32_t pk_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(2000u)
      : "memory");
  return ok;
}
__device__ __forceinline__ void pk_wait(uint32_t bar, uint32_t parity) {
  unsigned spins = 0;
  while (!pk_try(bar, parity))
    if (++spins > (1u << 22)) __trap();     // a protocol bug must fail loudly, never hang the GPU (~8 s)
}
// wait that adds the cycles it blocked to acc (profiling builds of the roles pass their counters; cheap otherwise)
__device__ __forceinline__ void pk_wait_t(uint32_t bar, uint32_t parity, long long& acc, bool prof) {
  if (!prof) { pk_wait(bar, parity); return; }
  const long long t0 = clock64();          // try_wait itself may block up to a hardware time limit: time the whole wait
  pk_wait(bar, parity);
  acc += clock64() - t0;
}
__device__ __forceinline__ void pk_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void pk_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ size_t pk_split_off16(int c0) { return (size_t)(c0 >> 5) * 128 + (size_t)((c0 >> 4) & 1) * 32; }

// fused epilogue of one output row x 16 channels: BatchNorm fold / bias, residual, ReLU, store (fp32 or split rows)
__device__ __forceinline__ void pk_store16(const PkParams& p, long long row, int c0, float* v) {
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
    const float* rrow = p.residual + row * p.res_ld;
    if (p.res_split) {
      const uint4* h = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(rrow) + pk_split_off16(c0));
      const uint4 hh[2] = {__ldg(h), __ldg(h + 1)}, ll[2] = {__ldg(h + 4), __ldg(h + 5)};
      const uint32_t* hp = reinterpret_cast<const uint32_t*>(hh);
      const uint32_t* lp = reinterpret_cast<const uint32_t*>(ll);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[2 * e] += __uint_as_float(hp[e] << 16) + __uint_as_float(lp[e] << 16);
        v[2 * e + 1] += __uint_as_float(hp[e] & 0xFFFF0000u) + __uint_as_float(lp[e] & 0xFFFF0000u);
      }
    } else {
#pragma unroll
      for (int e4 = 0; e4 < 4; ++e4) {
        const float4 r4 = __ldg(reinterpret_cast<const float4*>(rrow + c0 + e4 * 4));
        v[e4 * 4 + 0] += r4.x; v[e4 * 4 + 1] += r4.y; v[e4 * 4 + 2] += r4.z; v[e4 * 4 + 3] += r4.w;
      }
    }
  }
  if (p.flags & AG3D_RELU) {
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
  }
  float* orow = p.out + row * p.out_ld;
  if (p.out_split) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split2(v[2 * e], v[2 * e + 1], h[e], l[e]);
    uint4* d = reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(orow) + pk_split_off16(c0));
    d[0] = make_uint4(h[0], h[1], h[2], h[3]);
    d[1] = make_uint4(h[4], h[5], h[6], h[7]);
    d[4] = make_uint4(l[0], l[1], l[2], l[3]);
    d[5] = make_uint4(l[4], l[5], l[6], l[7]);
  } else {
    float4* dst = reinterpret_cast<float4*>(orow + c0);
#pragma unroll
    for (int e4 = 0; e4 < 4; ++e4) dst[e4] = make_float4(v[e4 * 4], v[e4 * 4 + 1], v[e4 * 4 + 2], v[e4 * 4 + 3]);
  }
}

template <int COUT>
__global__ void __launch_bounds__(PK_THREADS, 1) spconv_pk_kernel(const __grid_constant__ CUtensorMap tm_in, const PkParams p) {
  constexpr int HALF = COUT / 2;          // columns per owner warp
  constexpr int NCH = HALF / 16;          // 16-column chunks per owner warp
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 384);
  const uint32_t bar_base = smem_u32(bars);
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (8 + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (16 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (24 + s); };
  auto d_full = [&](int s) { return bar_base + 8u * (32 + s); };
  auto d_empty = [&](int s) { return bar_base + 8u * (36 + s); };
  auto plan_full = [&](int s) { return bar_base + 8u * (40 + s); };
  auto plan_empty = [&](int s) { return bar_base + 8u * (42 + s); };
  unsigned char* plan0 = smem + PK_BAR_BYTES;
  auto plan_list = [&](int buf) { return reinterpret_cast<int*>(plan0 + (size_t)buf * p.plan_bytes); };
  auto plan_cnt = [&](int buf) { return reinterpret_cast<int*>(plan0 + (size_t)buf * p.plan_bytes + (size_t)p.K * 1024); };
  auto plan_npass = [&](int buf) { return reinterpret_cast<int*>(plan0 + (size_t)buf * p.plan_bytes + (size_t)p.K * 1024 + 512); };
  auto plan_mask = [&](int buf) { return reinterpret_cast<uint32_t*>(plan0 + (size_t)buf * p.plan_bytes + (size_t)p.K * 1024 + 640); };
  unsigned char* b_smem = smem + p.off_b;
  unsigned char* a_smem = smem + p.off_a;
  a_smem += (1024u - (smem_u32(a_smem) & 1023u)) & 1023u;     // swizzle atoms are 1024-byte aligned
  const uint32_t b_stage_bytes = (uint32_t)COUT * 128u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool prof_on = p.prof != nullptr;
  const int n_slab = p.cin >> 5;
  const int K = p.K, NA = p.NA, NB = p.NB;

  if (tid == 0) {
    for (int s = 0; s < PK_MAX_NA; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
    for (int s = 0; s < PK_MAX_NB; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
    for (int s = 0; s < PK_ND; ++s) { mbar_init(d_full(s), 1); mbar_init(d_empty(s), PK_E_WARPS); }
    for (int s = 0; s < 2; ++s) { mbar_init(plan_full(s), 1); mbar_init(plan_empty(s), PK_E_WARPS + PK_P_WARPS + 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == PK_WARP_MMA) {
    constexpr uint32_t cols = PK_ND * COUT <= 128 ? 128u : (PK_ND * COUT <= 256 ? 256u : 512u);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < PK_E_WARPS) {
    // =========================================================================== owners / epilogue
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(PkRegs<COUT>::E));
    const int q = warp & 3, h = warp >> 2;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * HALF);
    float acc0[HALF], acc1[HALF];
    int d_slot = 0;
    uint32_t d_phase = 0;
    int it = 0;
    long long w_plan = 0, w_d = 0;
    const long long t_begin = clock64();
    for (int t = blockIdx.x; t < p.n_super; t += gridDim.x, ++it) {
      const int buf = it & 1;
      pk_wait_t(plan_full(buf), (uint32_t)(it >> 1) & 1u, w_plan, prof_on);
      const uint32_t mask0 = plan_mask(buf)[q * PK_SEG + lane], mask1 = plan_mask(buf)[q * PK_SEG + 32 + lane];
      const int* npass = plan_npass(buf);
#pragma unroll
      for (int i = 0; i < HALF; ++i) acc0[i] = acc1[i] = 0.f;
      for (int k = 0; k < K; ++k) {
        const int np = npass[k];
        const uint32_t p0 = (mask0 >> k) & 1u, p1 = (mask1 >> k) & 1u;
        const uint32_t b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
        const int rank0 = __popc(b0 & lt), rank1 = __popc(b0) + __popc(b1 & lt);
        for (int pass = 0; pass < np; ++pass) {
          const int db = d_slot;
          if (p.debug & 32) continue;
          pk_wait_t(d_full(db), d_phase, w_d, prof_on);
          if (++d_slot == PK_ND) { d_slot = 0; d_phase ^= 1u; }
          tc_fence_after();
          const int s0 = rank0 - 32 * pass, s1 = rank1 - 32 * pass;
          const bool ok0 = p0 && (unsigned)s0 < 32u, ok1 = p1 && (unsigned)s1 < 32u;
          if (__ballot_sync(0xffffffffu, ok0 || ok1) && !(p.debug & 8)) {
            const uint32_t taddr = t_lane + (uint32_t)(db * COUT);
            // the next 16-column chunk is in flight while this one is routed (registers permitting: not at Cout = 128)
            constexpr int NBUF = COUT >= 128 ? 1 : 2;
            uint32_t r[NBUF][16];
            pk_ld16(taddr, r[0]);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
              pk_ld_wait();
              if (NBUF == 2 && ch + 1 < NCH) pk_ld16(taddr + (uint32_t)((ch + 1) * 16), r[(ch + 1) % NBUF]);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float va = __shfl_sync(0xffffffffu, __uint_as_float(r[ch % NBUF][i]), s0 & 31);
                const float vb = __shfl_sync(0xffffffffu, __uint_as_float(r[ch % NBUF][i]), s1 & 31);
                if (ok0) acc0[ch * 16 + i] += va;
                if (ok1) acc1[ch * 16 + i] += vb;
              }
              if (NBUF == 1 && ch + 1 < NCH) pk_ld16(taddr + (uint32_t)((ch + 1) * 16), r[0]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty(db));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(plan_empty(buf));
      // ---- fused epilogue straight from the owner registers
      const long long row_a = (long long)t * PK_R + q * PK_SEG + lane, row_b = row_a + 32;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        if (row_a < p.n_out && !(p.debug & 16)) pk_store16(p, row_a, h * HALF + ch * 16, acc0 + ch * 16);
        if (row_b < p.n_out && !(p.debug & 16)) pk_store16(p, row_b, h * HALF + ch * 16, acc1 + ch * 16);
      }
    }
    if (p.prof && blockIdx.x == 0 && tid == 0) { p.prof[0] = clock64() - t_begin; p.prof[1] = w_plan; p.prof[2] = w_d; }
  } else if (warp < PK_E_WARPS + PK_P_WARPS) {
    // =========================================================================== TMA gather producers
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PkRegs<COUT>::P));
    const int pw = warp - PK_E_WARPS, pair = pw >> 1, part = pw & 1;
    const bool active = lane < 16;
    const int g = part * 16 + (lane & 15);            // row group of the stage: pair rows 4g .. 4g+3
    const int q = g >> 3, gi = g & 7;
    const uint32_t a_base = smem_u32(a_smem) + (uint32_t)g * 512u;
    uint32_t n = 0;              // stage counter (low two bits: which producer pair issues it)
    int a_slot = 0;              // ring slot and phase of stage n, kept incrementally (no divisions in the stage loop)
    uint32_t a_phase = 0;
    int it = 0;
    long long w_plan = 0, w_a = 0;
    const long long t_begin = clock64();
    for (int t = blockIdx.x; t < p.n_super; t += gridDim.x, ++it) {
      const int buf = it & 1;
      pk_wait_t(plan_full(buf), (uint32_t)(it >> 1) & 1u, w_plan, prof_on);
      const int* list = plan_list(buf);
      const int* cnt = plan_cnt(buf);
      const int* npass = plan_npass(buf);
      for (int k = 0; k < K; ++k) {
        const int np = npass[k];
        if (np == 0) continue;
        const int4 c4 = *reinterpret_cast<const int4*>(cnt + k * 4);
        const int cq = q == 0 ? c4.x : (q == 1 ? c4.y : (q == 2 ? c4.z : c4.w));
        for (int c = 0; c < n_slab; ++c) {
          for (int pass = 0; pass < np; ++pass) {
            const int slot = a_slot;
            const uint32_t phase = a_phase;
            if (++a_slot == NA) { a_slot = 0; a_phase ^= 1u; }
            if ((int)(n++ & 3u) != pair) continue;
            const int pos = 32 * pass + 4 * gi;
            const bool issue = active && pos < cq;
            int4 rows = make_int4(-1, -1, -1, -1);
            if (issue) rows = *reinterpret_cast<const int4*>(list + (k * 4 + q) * PK_SEG + pos);
            pk_wait_t(a_empty(slot), phase ^ 1u, w_a, prof_on);
            if (p.debug & 1) {
              if (part == 0 && lane == 0) mbar_arrive(a_full(slot));
              continue;
            }
            if (part == 0 && lane == 0) {
              int tot = 0;
              tot += min(8, (max(0, c4.x - 32 * pass) + 3) >> 2);
              tot += min(8, (max(0, c4.y - 32 * pass) + 3) >> 2);
              tot += min(8, (max(0, c4.z - 32 * pass) + 3) >> 2);
              tot += min(8, (max(0, c4.w - 32 * pass) + 3) >> 2);
              mbar_arrive_expect_tx(a_full(slot), (uint32_t)tot * 512u);
            }
            __syncwarp();
            if (issue) pk_gather4(a_base + (uint32_t)slot * PK_STAGE, &tm_in, a_full(slot), c * 64, rows);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(plan_empty(buf));
    }
    if (p.prof && blockIdx.x == 0 && pw == 0 && lane == 0) { p.prof[8] = clock64() - t_begin; p.prof[9] = w_plan; p.prof[10] = w_a; }
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PkRegs<COUT>::M));
    if (warp == PK_WARP_MMA) {
      // =========================================================================== MMA issuer
      const uint32_t idesc = umma_idesc_bf16(COUT);
      const uint32_t b_lbo = (uint32_t)COUT * 16u;
      const uint32_t d_hi32 = umma_desc_hi32(128);
      const uint32_t a_hi32 = umma_desc_hi32(1024) | (2u << 29);          // SWIZZLE_128B, SBO = 1024 (eight rows)
      const uint32_t a_lo32 = umma_desc_lo32(smem_u32(a_smem), 16);
      const uint32_t b_lo32 = umma_desc_lo32(smem_u32(b_smem), b_lbo);
      const uint32_t b_lo_off = (4u * b_lbo) >> 4, b_ks_off = (2u * b_lbo) >> 4, b_stride16 = b_stage_bytes >> 4;
      uint32_t n = 0;
      int a_slot = 0, b_slot = 0, d_slot = 0;                 // ring positions kept incrementally (no divisions per stage)
      uint32_t a_phase = 0, b_phase = 0, d_phase = 0;
      int it = 0;
      long long w_plan = 0, w_a = 0, w_b = 0, w_d = 0;
      const long long t_begin = clock64();
      for (int t = blockIdx.x; t < p.n_super; t += gridDim.x, ++it) {
        const int buf = it & 1;
        pk_wait_t(plan_full(buf), (uint32_t)(it >> 1) & 1u, w_plan, prof_on);
        const int* npass = plan_npass(buf);
        for (int k = 0; k < K; ++k) {
          const int np = npass[k];
          int db[2] = {0, 0};
          for (int c = 0; c < n_slab; ++c) {
            const int sb = b_slot;
            if (!(p.debug & 64)) pk_wait_t(b_full(sb), b_phase, w_b, prof_on);
            if (++b_slot == NB) { b_slot = 0; b_phase ^= 1u; }
            const uint32_t b_cur = b_lo32 + (uint32_t)sb * b_stride16;
            for (int pass = 0; pass < np; ++pass, ++n) {
              if (c == 0) {
                db[pass] = d_slot;
                if (!(p.debug & 32)) pk_wait_t(d_empty(d_slot), d_phase ^ 1u, w_d, prof_on);
                if (++d_slot == PK_ND) { d_slot = 0; d_phase ^= 1u; }
                tc_fence_after();                                   // the owners' tcgen05.ld of this buffer happened-before
              }
              const int slot = a_slot;
              pk_wait_t(a_full(slot), a_phase, w_a, prof_on);     // TMA (async proxy) -> MMA (async proxy): no fence needed
              if (++a_slot == NA) { a_slot = 0; a_phase ^= 1u; }
              const uint32_t a_cur = a_lo32 + (uint32_t)slot * (PK_STAGE >> 4);
              const uint32_t d = tmem_base + (uint32_t)(db[pass] * COUT);
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < ((p.debug & 4) ? 0 : 2); ++ks) {   // two 16-channel steps per 32-channel slab, three products each
                  const uint64_t da_hi = umma_desc_join(a_hi32, a_cur + ks * 2u);
                  const uint64_t da_lo = umma_desc_join(a_hi32, a_cur + ks * 2u + 4u);
                  const uint64_t db_hi = umma_desc_join(d_hi32, b_cur + ks * b_ks_off);
                  const uint64_t db_lo = umma_desc_join(d_hi32, b_cur + ks * b_ks_off + b_lo_off);
                  umma_bf16(d, da_hi, db_hi, idesc, (c > 0 || ks > 0) ? 1u : 0u);
                  umma_bf16(d, da_hi, db_lo, idesc, 1u);
                  umma_bf16(d, da_lo, db_hi, idesc, 1u);
                }
                umma_commit(a_empty(slot));
                if (c == n_slab - 1 && !(p.debug & 32)) umma_commit(d_full(db[pass]));
              }
              __syncwarp();
            }
            if (!(p.debug & 64) && elect_one()) umma_commit(b_empty(sb));   // all passes of this (k, slab) have read the weight stage
            __syncwarp();
          }
        }
        if (lane == 0) mbar_arrive(plan_empty(buf));
        __syncwarp();
      }
      if (p.prof && blockIdx.x == 0 && lane == 0) {
        p.prof[16] = clock64() - t_begin; p.prof[17] = w_plan; p.prof[18] = w_a; p.prof[19] = w_b; p.prof[20] = w_d; p.prof[21] = n;
      }
    } else if (warp == PK_WARP_W) {
      // =========================================================================== weight stages
      int b_slot = 0;
      uint32_t b_phase = 0;
      for (int t = blockIdx.x; t < p.n_super && !(p.debug & 64); t += gridDim.x) {
        for (int k = 0; k < K; ++k) {
          for (int c = 0; c < n_slab; ++c) {
            const int sb = b_slot;
            pk_wait(b_empty(sb), b_phase ^ 1u);
            if (++b_slot == NB) { b_slot = 0; b_phase ^= 1u; }
            if (p.debug & 2) {
              if (elect_one()) mbar_arrive(b_full(sb));
              __syncwarp();
              continue;
            }
            if (elect_one()) {
              mbar_arrive_expect_tx(b_full(sb), b_stage_bytes);
              const unsigned char* src = reinterpret_cast<const unsigned char*>(p.wp) + ((size_t)k * n_slab + c) * (size_t)b_stage_bytes;
              bulk_g2s(smem_u32(b_smem + (size_t)sb * b_stage_bytes), src, b_stage_bytes, b_full(sb));
            }
            __syncwarp();
          }
        }
      }
    } else {
      // =========================================================================== planners (2 warps, 2 segments each)
      const int pw = warp - PK_WARP_PLAN;
      const uint32_t lt = (1u << lane) - 1u;
      int it = 0;
      long long w_plan = 0;
      const long long t_begin = clock64();
      for (int t = blockIdx.x; t < p.n_super; t += gridDim.x, ++it) {
        const int buf = it & 1;
        if (it >= 2) pk_wait_t(plan_empty(buf), (uint32_t)((it >> 1) - 1) & 1u, w_plan, prof_on);
        int* list = plan_list(buf);
        int* cnt = plan_cnt(buf);
#pragma unroll 1
        for (int ss = 0; ss < 2; ++ss) {
          const int s = 2 * pw + ss;
          const long long row_a = (long long)t * PK_R + s * PK_SEG + lane, row_b = row_a + 32;
          const bool in_a = row_a < p.n_out, in_b = row_b < p.n_out;
          uint32_t mask_a = 0, mask_b = 0;
          for (int k0 = 0; k0 < K; k0 += 8) {
            int va[8], vb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              va[j] = vb[j] = -1;
              if (k0 + j < K) {
                const int* col = p.nbr + (long long)(k0 + j) * p.n_out;
                if (p.debug & 128) {
                  va[j] = (lane % 3 == 0 && in_a) ? (int)row_a : -1;
                  vb[j] = (lane % 3 == 1 && in_b) ? (int)row_b : -1;
                } else {
                  if (in_a) va[j] = __ldg(col + row_a);
                  if (in_b) vb[j] = __ldg(col + row_b);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int k = k0 + j;
              if (k < K) {
                const bool pa = va[j] >= 0, pb = vb[j] >= 0;
                const uint32_t ba = __ballot_sync(0xffffffffu, pa), bb = __ballot_sync(0xffffffffu, pb);
                int* dst = list + (k * 4 + s) * PK_SEG;
                if (pa) dst[__popc(ba & lt)] = va[j];
                if (pb) dst[__popc(ba) + __popc(bb & lt)] = vb[j];
                if (lane == 0) cnt[k * 4 + s] = __popc(ba) + __popc(bb);
                mask_a |= (pa ? 1u : 0u) << k;
                mask_b |= (pb ? 1u : 0u) << k;
              }
            }
          }
          plan_mask(buf)[s * PK_SEG + lane] = mask_a;
          plan_mask(buf)[s * PK_SEG + 32 + lane] = mask_b;
        }
        pk_bar_sync(1, 64);                                   // both planner warps have written their segments
        if (pw == 0) {
          if (lane < K) {
            const int4 c4 = *reinterpret_cast<const int4*>(cnt + lane * 4);
            const int m = max(max(c4.x, c4.y), max(c4.z, c4.w));
            plan_npass(buf)[lane] = (m + 31) >> 5;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(plan_full(buf));
        }
        pk_bar_sync(1, 64);                                   // keep the pair in step (cnt is read above)
      }
      if (p.prof && blockIdx.x == 0 && pw == 0 && lane == 0) { p.prof[24] = clock64() - t_begin; p.prof[25] = w_plan; p.prof[26] = it; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == PK_WARP_MMA) {
    constexpr uint32_t cols = PK_ND * COUT <= 128 ? 128u : (PK_ND * COUT <= 256 ? 256u : 512u);
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(cols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------- host side
bool spconv_pk_supported(int cin, int cout, int K) {
  return cin % 32 == 0 && cin >= 32 && cin <= 384 && (cout == 32 || cout == 64 || cout == 96 || cout == 128) && K >= 2 &&
         K <= 32;
}

// AUTO never takes the packed kernel in this version; AG3D_ALGO_TC_PACKED (or AG3D_PK_ALL=1, experiments) selects it.
// Measured on B200 (profiles/r02_pk_vs_dense_150k.txt, 8 x 150k voxels; profiles/r02_pk_role_waits.txt): both kernels are
// bound by shared-memory bandwidth, not by the tensor pipe - a [128 x 32ch] stage moves ~70 KB through shared memory (TMA
// writes 28 KB, the six bf16x3 MMAs read 42 KB; N = 96 MMAs run at 56 cycles = 125 B/cycle of operand reads,
// tools/probes/mma_probe.cu) and the single MMA-issuing warp of the one resident CTA is busy ~930 cycles per stage.
// Packing removes 1.8x of the stages, yet on the 96/128-channel 3x3x3 layers the dense kernel's two co-resident CTAs
// still win (packed 0.7-0.85x); it wins 1.15-1.33x on the stride-2 / transposed maps and the 32-channel layers at
// >= 1 M rows and loses below ~300 k rows (per-CTA prologue, 4-7 super tiles per CTA).
bool spconv_pk_preferred(long long n_out, int K, int cin, int cout) {
  static int all = -1;
  if (all < 0) { const char* a = getenv("AG3D_PK_ALL"); all = a ? atoi(a) : 0; }
  (void)K; (void)cin; (void)cout;
  return all && n_out >= (long long)PK_R * sm_count() * 2;
}

int spconv_pk_launch(const float* in, long long n_in, int in_ld, int cin, const int* nbr, int K, long long n_out,
                     const void* wprep, int cout, const float* scale, const float* shift, const float* residual,
                     int res_ld, float* out, int out_ld, int flags, cudaStream_t st) {
  AG3D_CHECK_ARG(spconv_pk_supported(cin, cout, K) && nbr, "shape not supported by the pair-packed path");
  AG3D_CHECK_ARG((flags & AG3D_IN_SPLIT) && n_in > 0, "the pair-packed path gathers split rows with a known row count");
  AG3D_CHECK_ARG(wprep && aligned16(wprep), "prepared weights missing (ag3d_spconv_tc_prepare_weight)");
  PkParams p;
  p.nbr = nbr; p.K = K; p.n_out = n_out; p.wp = static_cast<const uint4*>(wprep); p.cin = cin;
  p.scale = scale; p.shift = shift; p.residual = residual; p.res_ld = res_ld;
  p.out = out; p.out_ld = out_ld; p.flags = flags;
  p.out_split = (flags & AG3D_OUT_SPLIT) ? 1 : 0;
  p.res_split = (flags & AG3D_RES_SPLIT) ? 1 : 0;
  p.n_super = (int)((n_out + PK_R - 1) / PK_R);
  p.plan_bytes = (uint32_t)K * 1024u + 512u + 128u + 1024u;
  p.off_b = PK_BAR_BYTES + 2 * p.plan_bytes;
  const uint32_t b_stage = (uint32_t)cout * 128u;
  static int force_na = -1, force_nb = -1;      // tuning experiments only
  if (force_na < 0) { const char* e = getenv("AG3D_PK_NA"); force_na = e ? atoi(e) : 0; }
  if (force_nb < 0) { const char* e = getenv("AG3D_PK_NB"); force_nb = e ? atoi(e) : 0; }
  p.NB = force_nb >= 2 && force_nb <= PK_MAX_NB ? force_nb : 3;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("AG3D_PK_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
    static long long* prof = nullptr;
    static int want_prof = -1;
    if (want_prof < 0) { const char* e = getenv("AG3D_PK_PROF"); want_prof = e ? atoi(e) : 0; }
    if (want_prof && !prof) { cudaMallocManaged(&prof, 32 * sizeof(long long)); }
    p.prof = want_prof ? prof : nullptr;
  }
  p.off_a = p.off_b + (uint32_t)p.NB * b_stage;
  const size_t budget = 227 * 1024;
  int na = (int)((budget - p.off_a - 1024) / PK_STAGE);
  na = std::min(na, PK_MAX_NA);
  if (force_na >= 2 && force_na <= PK_MAX_NA) na = std::min(na, force_na);
  AG3D_CHECK_ARG(na >= 2, "shared memory budget");
  p.NA = na;
  const size_t smem = (size_t)p.off_a + 1024 + (size_t)na * PK_STAGE;
  alignas(64) CUtensorMap tm_in;
  memset(&tm_in, 0, sizeof(tm_in));
  if (!make_row_map(&tm_in, in, in_ld, cin, n_in)) {
    set_error("cuTensorMapEncodeTiled failed (CUresult " + std::to_string(g_last_tmap_rc) + ", rows " + std::to_string(n_in) +
              ", ld " + std::to_string(in_ld) + ", cin " + std::to_string(cin) + ")");
    return AG3D_E_INVALID;
  }
  static bool attr = false;
  if (!attr) {
    AG3D_CUDA(cudaFuncSetAttribute(spconv_pk_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AG3D_CUDA(cudaFuncSetAttribute(spconv_pk_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AG3D_CUDA(cudaFuncSetAttribute(spconv_pk_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AG3D_CUDA(cudaFuncSetAttribute(spconv_pk_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  const unsigned grid = (unsigned)std::min(p.n_super, sm_count());
  switch (cout) {
    case 32: spconv_pk_kernel<32><<<grid, PK_THREADS, smem, st>>>(tm_in, p); break;
    case 64: spconv_pk_kernel<64><<<grid, PK_THREADS, smem, st>>>(tm_in, p); break;
    case 96: spconv_pk_kernel<96><<<grid, PK_THREADS, smem, st>>>(tm_in, p); break;
    default: spconv_pk_kernel<128><<<grid, PK_THREADS, smem, st>>>(tm_in, p); break;
  }
  AG3D_LAUNCH_CHECK("spconv_pk");
  if (p.prof) {      // measurement aid only: synchronises
    cudaStreamSynchronize(st);
    static int printed = 0;
    if (printed++ % 13 == 12) {
      const long long* q = p.prof;
      fprintf(stderr, "pk prof (cycles, CTA 0): owners total %lld wait plan %lld d_full %lld | producers total %lld plan %lld a_empty %lld | "
              "mma total %lld plan %lld a_full %lld b_full %lld d_empty %lld stages %lld | planner total %lld wait %lld tiles %lld\n",
              q[0], q[1], q[2], q[8], q[9], q[10], q[16], q[17], q[18], q[19], q[20], q[21], q[24], q[25], q[26]);
    }
  }
  return AG3D_OK;
}

}  // namespace ag3d
