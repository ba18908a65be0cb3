// the parts of the interactive loop that sit around forward_mask in every click round
// (eval_multi_obj.py:118-167) and the voxelisation front end, on the device.
//   ag3d_click_pred        pred = argmax of the mask logits, clicked voxels overwritten with their object id
//                          (eval_multi_obj.py:124-139)
//   ag3d_scene_iou         pred[inverse_map] against the full-resolution labels: per-object intersection / counts
//                          (eval_multi_obj.py:143-148, utils/seg.py:9-17,44-59)
//   ag3d_click_simulate    utils/seg.py:173-226 (get_simulated_clicks + measure_error_size + get_next_click_coo_torch):
//                          error clusters (gt, pred), every error voxel's distance to the nearest voxel outside its
//                          cluster, per cluster the voxel furthest from the border, clusters ranked by that distance.
//                          The reference materialises a torch.cdist matrix per cluster; here one N-body style pass over
//                          shared-memory tiles computes all clusters at once (E x N pair evaluations, no matrix).
//   ag3d_quantize_points   ME.utils.sparse_quantize front end (datasets/InterMultiObj3DSegDataset.py:67-71):
//                          floor(p / q) -> int32 (b, x, y, z); uniqueness / first-occurrence order / inverse map come
//                          from ag3d_downsample with stride 1, unique_map from ag3d_first_rows.
#include <float.h>
#include <limits.h>

#include "common.cuh"

namespace ag3d {

constexpr int CK_MAX_OBJ = 32;                          // object ids 0 .. 31 (labels are uint8 elsewhere, n_obj <= 32)
constexpr int CK_CLUSTERS = CK_MAX_OBJ * CK_MAX_OBJ;    // (gt, pred) pairs
constexpr int CK_TILE = 256;

__global__ void click_pred_kernel(const float* __restrict__ logits, int n_obj, long long nv, int* __restrict__ pred) {
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; v < nv; v += step) {
    int arg = 0;
    if (logits) {
      float best = __ldg(logits + v * n_obj);
      for (int o = 1; o < n_obj; ++o) {
        const float z = __ldg(logits + v * n_obj + o);
        if (z > best) { best = z; arg = o; }              // first maximum, as torch.argmax
      }
    }
    pred[v] = arg;
  }
}
__global__ void click_override_kernel(const int* __restrict__ rows, const int* __restrict__ objs, int n, long long nv,
                                      int* __restrict__ pred) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && rows[i] >= 0 && rows[i] < nv) pred[rows[i]] = objs[i];
}

// counts[o] = (|pred_full == o & labels_full == o|, |pred_full == o|, |labels_full == o|), pred_full = pred[inverse_map]
__global__ void scene_iou_kernel(const int* __restrict__ pred, const long long* __restrict__ inverse_map,
                                 const int* __restrict__ labels_full, long long n_full, int n_obj,
                                 unsigned long long* __restrict__ counts) {
  __shared__ unsigned int c_s[CK_MAX_OBJ * 3];
  for (int i = threadIdx.x; i < CK_MAX_OBJ * 3; i += blockDim.x) c_s[i] = 0;
  __syncthreads();
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n_full; i += step) {
    const int p = pred[inverse_map ? inverse_map[i] : i], l = labels_full[i];
    if (p >= 0 && p < n_obj) atomicAdd(&c_s[p * 3 + 1], 1u);
    if (l >= 0 && l < n_obj) {
      atomicAdd(&c_s[l * 3 + 2], 1u);
      if (p == l) atomicAdd(&c_s[l * 3 + 0], 1u);
    }
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_obj * 3; k += blockDim.x)
    if (c_s[k]) atomicAdd(counts + k, (unsigned long long)c_s[k]);
}

// cluster id of a voxel: gt * 32 + pred where the prediction is wrong, -1 elsewhere; error voxels are appended to a list
__global__ void click_cluster_kernel(const int* __restrict__ pred, const int* __restrict__ gt, long long nv,
                                     int* __restrict__ cid, int* __restrict__ err_rows, int* __restrict__ n_err) {
  long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; v < nv; v += step) {
    const int p = pred[v], g = gt[v];
    const bool err = p != g;
    cid[v] = err ? g * CK_MAX_OBJ + p : -1;
    const unsigned m = __ballot_sync(__activemask(), err);
    if (err) {                                         // one atomic per warp; list order is irrelevant (the result is a max)
      const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(n_err, __popc(m));
      base = __shfl_sync(m, base, leader);
      err_rows[base + __popc(m & ((1u << lane) - 1u))] = (int)v;
    }
  }
}

// One thread per error voxel e: d2[e] = min over all voxels u with cid[u] != cid[e] of |x_e - x_u|^2, voxels staged
// through shared memory; then the per-cluster maximum with the smallest row among equal maxima (torch.where(...)[0][0])
__global__ void __launch_bounds__(CK_TILE)
click_border_distance_kernel(const float* __restrict__ xyz, const int* __restrict__ cid, long long nv,
                             const int* __restrict__ err_rows, const int* __restrict__ n_err_p,
                             unsigned long long* __restrict__ cluster_best) {
  __shared__ float4 tile[CK_TILE];
  const int n_err = *n_err_p;
  for (int e0 = blockIdx.x * CK_TILE; e0 < n_err; e0 += gridDim.x * CK_TILE) {
    const int e = e0 + threadIdx.x;
    const bool live = e < n_err;
    const int row = live ? err_rows[e] : 0;
    const float px = xyz[(size_t)row * 3], py = xyz[(size_t)row * 3 + 1], pz = xyz[(size_t)row * 3 + 2];
    const int mine = live ? cid[row] : -2;
    float best = FLT_MAX;
    for (long long u0 = 0; u0 < nv; u0 += CK_TILE) {
      const long long u = u0 + threadIdx.x;
      __syncthreads();
      if (u < nv) tile[threadIdx.x] = make_float4(xyz[u * 3], xyz[u * 3 + 1], xyz[u * 3 + 2], __int_as_float(cid[u]));
      else tile[threadIdx.x] = make_float4(0.f, 0.f, 0.f, __int_as_float(-3));        // never read: the loop stops at lim
      __syncthreads();
      const int lim = (int)min((long long)CK_TILE, nv - u0);
#pragma unroll 8
      for (int j = 0; j < lim; ++j) {
        const float4 t = tile[j];
        const float dx = px - t.x, dy = py - t.y, dz = pz - t.z;
        const float d2 = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
        if (__float_as_int(t.w) != mine) best = fminf(best, d2);
      }
    }
    if (live && best < FLT_MAX) {
      // maximise (distance, then smaller row): distance bits in the high word, ~row in the low word
      const unsigned long long key = ((unsigned long long)__float_as_uint(sqrtf(best)) << 32) | (unsigned)(~row);
      atomicMax(cluster_best + mine, key);
    }
  }
}

// ranks the non-empty clusters by size (descending; ties in the reference's order of ascending 96 gt + 11 pred) and writes
// the selected ones.  out: [0] = number of new clicks n, then n x (row, object id, cluster id), then the distances as float
// bits; one thread block
__global__ void click_select_kernel(const unsigned long long* __restrict__ cluster_best, int top_n,
                                    const int* __restrict__ perm, int max_new, int* __restrict__ out) {
  __shared__ float size_s[CK_CLUSTERS];
  __shared__ int rank_s[CK_CLUSTERS];
  __shared__ int order_s[CK_CLUSTERS];
  __shared__ int n_s;
  if (threadIdx.x == 0) n_s = 0;
  for (int c = threadIdx.x; c < CK_CLUSTERS; c += blockDim.x) {
    const unsigned long long k = cluster_best[c];
    size_s[c] = k ? __uint_as_float((unsigned)(k >> 32)) : -1.f;          // 0 = the cluster has no voxel
  }
  __syncthreads();
  for (int c = threadIdx.x; c < CK_CLUSTERS; c += blockDim.x) {
    rank_s[c] = -1;
    if (size_s[c] >= 0.f) {
      const int key_c = 96 * (c / CK_MAX_OBJ) + 11 * (c % CK_MAX_OBJ);
      int r = 0;
      for (int o = 0; o < CK_CLUSTERS; ++o) {
        if (o == c || size_s[o] < 0.f) continue;
        const int key_o = 96 * (o / CK_MAX_OBJ) + 11 * (o % CK_MAX_OBJ);
        if (size_s[o] > size_s[c] || (size_s[o] == size_s[c] && key_o < key_c)) ++r;
      }
      rank_s[c] = r;
      order_s[r] = c;
      atomicAdd(&n_s, 1);
    }
  }
  __syncthreads();
  int n = n_s;
  if (top_n >= 0 && n > top_n) n = top_n;
  if (n > max_new) n = max_new;
  if (threadIdx.x == 0) out[0] = n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    int src = perm ? perm[i] : i;                       // the reference shuffles the selected list (random.shuffle)
    if (src < 0 || src >= n) src = i;
    const int c = order_s[src];
    const unsigned long long k = cluster_best[c];
    out[1 + 3 * i] = (int)(~(unsigned)(k & 0xFFFFFFFFull));
    out[2 + 3 * i] = c / CK_MAX_OBJ;                    // the clicked voxel's ground-truth object
    out[3 + 3 * i] = c;
    out[1 + 3 * max_new + i] = (int)(unsigned)(k >> 32);
  }
}

// ---- voxelisation front end
__global__ void quantize_points_kernel(const float* __restrict__ pts, long long n, float q, int batch, int4* __restrict__ coords,
                                       int* __restrict__ status) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  int bad = 0;
  for (; i < n; i += step) {
    // numpy: np.floor(coordinates / quantization_size) in fp32 (IEEE division, no reciprocal)
    const float fx = floorf(__fdiv_rn(pts[i * 3], q)), fy = floorf(__fdiv_rn(pts[i * 3 + 1], q)), fz = floorf(__fdiv_rn(pts[i * 3 + 2], q));
    const bool ok = fabsf(fx) < (float)COORD_LIMIT && fabsf(fy) < (float)COORD_LIMIT && fabsf(fz) < (float)COORD_LIMIT;
    bad += ok ? 0 : 1;
    coords[i] = make_int4(batch, ok ? (int)fx : 0, ok ? (int)fy : 0, ok ? (int)fz : 0);
  }
  if (bad) atomicAdd(status, bad);
}
__global__ void first_rows_kernel(const int* __restrict__ parent, long long n, long long* __restrict__ unique_map) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) atomicMin(reinterpret_cast<unsigned long long*>(unique_map + parent[i]), (unsigned long long)i);
}
__global__ void fill_i64_kernel(long long* p, long long n, long long v) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) p[i] = v;
}

static unsigned ck_blocks(long long work, int threads = 256) {
  long long b = (work + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  return (unsigned)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace ag3d

using namespace ag3d;

extern "C" {

int ag3d_click_pred(const float* logits, int32_t n_obj, int64_t nv, const int32_t* click_rows, const int32_t* click_objs,
                    int32_t n_clicks, int32_t* pred, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(nv > 0 && pred && n_obj >= 1 && n_obj <= CK_MAX_OBJ, "click_pred: arguments");
  AG3D_CHECK_ARG(n_clicks == 0 || (click_rows && click_objs), "click_pred: click arrays");
  cudaStream_t st = as_stream(stream);
  click_pred_kernel<<<ck_blocks(nv), 256, 0, st>>>(logits, n_obj, nv, pred);
  AG3D_LAUNCH_CHECK("click_pred");
  if (n_clicks > 0) {
    click_override_kernel<<<(n_clicks + 127) / 128, 128, 0, st>>>(click_rows, click_objs, n_clicks, nv, pred);
    AG3D_LAUNCH_CHECK("click_override");
  }
  return AG3D_OK;
}

int ag3d_scene_iou(const int32_t* pred, const int64_t* inverse_map, const int32_t* labels_full, int64_t n_full, int32_t n_obj,
                   uint64_t* counts, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(pred && labels_full && counts && n_full > 0 && n_obj >= 1 && n_obj <= CK_MAX_OBJ, "scene_iou: arguments");
  cudaStream_t st = as_stream(stream);
  AG3D_CUDA(cudaMemsetAsync(counts, 0, sizeof(uint64_t) * 3 * n_obj, st));
  scene_iou_kernel<<<ck_blocks(n_full), 256, 0, st>>>(pred, reinterpret_cast<const long long*>(inverse_map), labels_full, n_full,
                                                      n_obj, reinterpret_cast<unsigned long long*>(counts));
  AG3D_LAUNCH_CHECK("scene_iou");
  return AG3D_OK;
}

size_t ag3d_click_simulate_workspace_bytes(int64_t nv) {
  return (size_t)nv * 8 + CK_CLUSTERS * 8 + 256;       // cid[nv] | err_rows[nv] | cluster_best[1024] | n_err
}

int ag3d_click_simulate(const int32_t* pred, const int32_t* gt, const float* xyz, int64_t nv, int32_t top_n,
                        const int32_t* perm, int32_t max_new, int32_t* out, void* ws, size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(pred && gt && xyz && out && nv > 0 && nv < INT_MAX && max_new >= 1 && max_new <= CK_CLUSTERS,
                 "click_simulate: arguments");
  AG3D_CHECK_ARG(ws && aligned16(ws) && ws_bytes >= ag3d_click_simulate_workspace_bytes(nv), "click_simulate: workspace too small");
  cudaStream_t st = as_stream(stream);
  int* cid = static_cast<int*>(ws);
  int* err_rows = cid + nv;
  unsigned long long* best = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(ws) + (((size_t)nv * 8 + 15) & ~(size_t)15));
  int* n_err = reinterpret_cast<int*>(best + CK_CLUSTERS);
  AG3D_CUDA(cudaMemsetAsync(best, 0, CK_CLUSTERS * 8 + 16, st));
  click_cluster_kernel<<<ck_blocks(nv), 256, 0, st>>>(pred, gt, nv, cid, err_rows, n_err);
  AG3D_LAUNCH_CHECK("click_cluster");
  // grid sized for the worst case (every voxel wrong); blocks past n_err exit at once
  long long blocks = (nv + CK_TILE - 1) / CK_TILE;
  if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
  click_border_distance_kernel<<<(unsigned)blocks, CK_TILE, 0, st>>>(xyz, cid, nv, err_rows, n_err, best);
  AG3D_LAUNCH_CHECK("click_border_distance");
  click_select_kernel<<<1, 256, 0, st>>>(best, top_n, perm, max_new, out);
  AG3D_LAUNCH_CHECK("click_select");
  return AG3D_OK;
}

int ag3d_quantize_points(const float* points, int64_t n, float quantization_size, int32_t batch_index, int32_t* coords,
                         int32_t* status, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(points && coords && status && n > 0 && n < INT_MAX && quantization_size > 0.f, "quantize_points: arguments");
  AG3D_CHECK_ARG(aligned16(coords), "quantize_points: coords must be 16-byte aligned");
  quantize_points_kernel<<<ck_blocks(n), 256, 0, as_stream(stream)>>>(points, n, quantization_size, batch_index,
                                                                       reinterpret_cast<int4*>(coords), status);
  AG3D_LAUNCH_CHECK("quantize_points");
  return AG3D_OK;
}

int ag3d_first_rows(const int32_t* parent, int64_t n, int64_t m, int64_t* unique_map, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(parent && unique_map && n > 0 && m > 0, "first_rows: arguments");
  cudaStream_t st = as_stream(stream);
  fill_i64_kernel<<<ck_blocks(m), 256, 0, st>>>(reinterpret_cast<long long*>(unique_map), m, LLONG_MAX);
  AG3D_LAUNCH_CHECK("fill_i64");
  first_rows_kernel<<<ck_blocks(n), 256, 0, st>>>(parent, n, reinterpret_cast<long long*>(unique_map));
  AG3D_LAUNCH_CHECK("first_rows");
  return AG3D_OK;
}

}  // extern "C"
