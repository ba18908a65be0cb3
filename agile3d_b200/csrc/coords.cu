// K1-K3: voxel-coordinate hashing, stride-2 coordinate maps, kernel-map (neighbour table) construction.
// All integer work, HBM/L2-latency bound: one 128-bit coordinate load per row, one 128-bit slot load per probe.
#include <limits.h>

#include <cub/device/device_radix_sort.cuh>
#include <mutex>

#include "common.cuh"

namespace ag3d {

std::atomic<long long> g_kernel_launches{0};
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what) {
  g_last_error = std::string("cuda: ") + what + ": " + cudaGetErrorString(e);
  return AG3D_E_CUDA;
}
int sm_count() {
  static int cached = 0;
  if (!cached) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
    if (cached <= 0) cached = 148;
  }
  return cached;
}

// ------------------------------------------------------------------------------------------------ kernels

__global__ void table_clear_kernel(Slot* table, long long cap) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  int4 e;
  e.x = -1; e.y = -1; e.z = INT_MAX; e.w = -1;
  for (; i < cap; i += step) reinterpret_cast<int4*>(table)[i] = e;
}

// Insert `key`; returns the slot index.  first_row keeps the minimum inserting row (first occurrence).
__device__ __forceinline__ long long table_insert(Slot* table, unsigned long long mask, unsigned long long key,
                                                  int row, bool* was_present) {
  unsigned long long s = mix64(key) & mask;
  while (true) {
    unsigned long long prev = atomicCAS(&table[s].key, EMPTY_KEY, key);
    if (prev == EMPTY_KEY || prev == key) {
      *was_present = (prev == key);
      atomicMin(&table[s].first_row, row);
      return (long long)s;
    }
    s = (s + 1) & mask;
  }
}

__global__ void hash_build_kernel(const int4* __restrict__ coords, long long n, Slot* table,
                                  unsigned long long mask, int* status) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    const int4 c = __ldg(coords + i);
    if (!coord_in_range(c.x, c.y, c.z, c.w)) {
      atomicAdd(status + 1, 1);
      continue;
    }
    bool present;
    long long s = table_insert(table, mask, pack_key(c.x, c.y, c.z, c.w), (int)i, &present);
    if (present) atomicAdd(status + 0, 1);
    else table[s].row = (int)i;   // unique input: row id == input row
  }
}

// --- downsample: pass 1 inserts the coarse key of every fine row, remembering the slot.
// n_dev (nullable): the row count lives on the device (the previous level's count of a chained downsample, not yet
// known to the host); n is then only the upper bound that sized the grid.
__global__ void coarse_insert_kernel(const int4* __restrict__ coords, long long n, const int* __restrict__ n_dev,
                                     int new_stride, Slot* table, unsigned long long mask, int* __restrict__ slot_of_row) {
  if (n_dev) n = *n_dev;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    const int4 c = __ldg(coords + i);
    const int x = floor_div(c.y, new_stride) * new_stride;
    const int y = floor_div(c.z, new_stride) * new_stride;
    const int z = floor_div(c.w, new_stride) * new_stride;
    bool present;
    slot_of_row[i] = (int)table_insert(table, mask, pack_key(c.x, x, y, z), (int)i, &present);
  }
}

// --- exclusive scan over "row i is the first occurrence of its coarse voxel" (3 kernels, 2048 rows per block)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int is_rep(const Slot* table, const int* slot_of_row, long long i) {
  return table[slot_of_row[i]].first_row == (int)i;
}

__global__ void rep_count_kernel(const Slot* __restrict__ table, const int* __restrict__ slot_of_row, long long n,
                                 const int* __restrict__ n_dev, int* __restrict__ block_sums) {
  if (n_dev) n = *n_dev;
  __shared__ int warp_sums[SCAN_THREADS / 32];
  const long long base = (long long)blockIdx.x * SCAN_TILE;
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    long long i = base + j * SCAN_THREADS + threadIdx.x;
    if (i < n) cnt += is_rep(table, slot_of_row, i);
  }
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < SCAN_THREADS / 32; ++w) t += warp_sums[w];
    block_sums[blockIdx.x] = t;
  }
}

__global__ void block_scan_kernel(int* block_sums, int n_blocks, int* out_total) {
  // single block; sequential over chunks of blockDim.x
  __shared__ int buf[1024];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n_blocks; base += blockDim.x) {
    int i = base + threadIdx.x;
    int v = (i < n_blocks) ? block_sums[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < blockDim.x; o <<= 1) {
      int t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    int incl = buf[threadIdx.x];
    if (i < n_blocks) block_sums[i] = carry + incl - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *out_total = carry;
}

__global__ void rep_assign_kernel(const int4* __restrict__ coords, long long n, const int* __restrict__ n_dev,
                                  int new_stride, Slot* table, const int* __restrict__ slot_of_row,
                                  const int* __restrict__ block_offsets, int4* __restrict__ out_coords) {
  if (n_dev) n = *n_dev;
  // rows are assigned to threads in *blocked* order inside the tile so that the scan preserves row order
  __shared__ int warp_sums[SCAN_THREADS / 32];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int flags[SCAN_ITEMS];
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    long long i = base + j;
    flags[j] = (i < n) ? is_rep(table, slot_of_row, i) : 0;
    cnt += flags[j];
  }
  // exclusive scan of cnt across the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = cnt;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  int warp_off = 0;
  for (int w = 0; w < warp; ++w) warp_off += warp_sums[w];
  int pos = block_offsets[blockIdx.x] + warp_off + incl - cnt;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    long long i = base + j;
    if (flags[j]) {
      const int4 c = __ldg(coords + i);
      int4 o;
      o.x = c.x;
      o.y = floor_div(c.y, new_stride) * new_stride;
      o.z = floor_div(c.z, new_stride) * new_stride;
      o.w = floor_div(c.w, new_stride) * new_stride;
      out_coords[pos] = o;
      table[slot_of_row[i]].row = pos;
      ++pos;
    }
  }
}

__global__ void parent_kernel(const Slot* __restrict__ table, const int* __restrict__ slot_of_row, long long n,
                              const int* __restrict__ n_dev, int* __restrict__ parent) {
  if (n_dev) n = *n_dev;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) parent[i] = table[slot_of_row[i]].row;
}

// --- kernel map: one thread per (offset k, output row o), o fastest -> coalesced coordinate loads and table writes
__global__ void kernel_map_kernel(const int4* __restrict__ out_coords, long long n_out, const Slot* __restrict__ table,
                                  unsigned long long mask, int ksize, int step_len, int* __restrict__ nbr,
                                  int* __restrict__ pair_count) {
  const int k = blockIdx.y;
  int r = k;
  const int jx = r % ksize; r /= ksize;
  const int jy = r % ksize; r /= ksize;
  const int jz = r;
  const int half = (ksize & 1) ? ksize / 2 : 0;
  const int dx = (jx - half) * step_len, dy = (jy - half) * step_len, dz = (jz - half) * step_len;
  long long o = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  int hits = 0;
  for (; o < n_out; o += step) {
    const int4 c = __ldg(out_coords + o);
    const int x = c.y + dx, y = c.z + dy, z = c.w + dz;
    int row = -1;
    if (coord_in_range(c.x, x, y, z)) row = table_find(table, mask, pack_key(c.x, x, y, z));
    nbr[(long long)k * n_out + o] = row;
    hits += (row >= 0);
  }
  if (pair_count) {
    for (int s = 16; s > 0; s >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, s);
    if ((threadIdx.x & 31) == 0 && hits) atomicAdd(pair_count + k, hits);
  }
}

__global__ void kernel_map_transposed_kernel(const int4* __restrict__ fine, const int* __restrict__ parent,
                                             long long n, int fine_stride, int* __restrict__ nbr) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  const int cs = 2 * fine_stride;
  for (; i < n; i += step) {
    const int4 c = __ldg(fine + i);
    const int ox = (c.y - floor_div(c.y, cs) * cs) / fine_stride;
    const int oy = (c.z - floor_div(c.z, cs) * cs) / fine_stride;
    const int oz = (c.w - floor_div(c.w, cs) * cs) / fine_stride;
    const int kk = ox + 2 * oy + 4 * oz;
    const int p = parent[i];
#pragma unroll
    for (int k = 0; k < 8; ++k) nbr[(long long)k * n + i] = (k == kk) ? p : -1;
  }
}

// offsets[b] = first row whose scene index is >= b (b = 0 .. max_scenes); offsets[max_scenes + 1] = number of rows whose
// scene index is smaller than their predecessor's (scenes must be contiguous and in batch order)
__global__ void scene_offsets_kernel(const int4* __restrict__ coords, long long n, int max_scenes, int* __restrict__ offsets) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b <= max_scenes) {
    long long lo = 0, hi = n;                       // lower bound of b in the (sorted) scene column
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (__ldg(&coords[mid].x) < b) lo = mid + 1; else hi = mid;
    }
    offsets[b] = (int)lo;
  }
}
__global__ void scene_order_kernel(const int4* __restrict__ coords, long long n, int* __restrict__ bad) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x + 1;
  const long long step = (long long)gridDim.x * blockDim.x;
  int c = 0;
  for (; i < n; i += step) c += __ldg(&coords[i].x) < __ldg(&coords[i - 1].x);
  if (c) atomicAdd(bad, c);
}

// ---- internal row order.  Rows of a level are sorted, scene by scene, by their 3x3x3 neighbour
// pattern (the 27-bit mask of existing offsets): rows of a 128-row tile then agree on which offsets are absent and the
// dense-tile convolution skips those (tile, offset) stages (18.8 instead of 26.9 of 27 at 150k voxels; Morton order: 26.8 -
// tools/order_analysis.py).  The permutation is internal: callers see their own row order (agile3d_b200/model.py).
__global__ void order_key_kernel(const int* __restrict__ nbr, int K, long long n, const int4* __restrict__ coords,
                                 unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    unsigned m = 0;
    for (int k = 0; k < K; ++k) m |= (__ldg(nbr + (long long)k * n + i) >= 0 ? 1u : 0u) << k;
    keys[i] = ((unsigned long long)(unsigned)__ldg(&coords[i].x) << 32) | m;
    vals[i] = (int)i;
  }
}
__global__ void invert_perm_kernel(const int* __restrict__ perm, long long n, int* __restrict__ inv) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) inv[perm[i]] = (int)i;
}
// out[k][i] = remap(nbr[k][perm_out ? perm_out[i] : i]),  remap(v) = v < 0 ? -1 : (inv_in ? inv_in[v] : v)
__global__ void permute_map_kernel(const int* __restrict__ nbr, long long n_out, const int* __restrict__ perm_out,
                                   const int* __restrict__ inv_in, int* __restrict__ out) {
  const int k = blockIdx.y;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n_out; i += step) {
    const long long src = perm_out ? perm_out[i] : i;
    int v = __ldg(nbr + (long long)k * n_out + src);
    if (v >= 0 && inv_in) v = __ldg(inv_in + v);
    out[(long long)k * n_out + i] = v;
  }
}

static inline int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// dst[i] = src[idx[i]] for rows of `row_bytes` bytes (multiple of 4): one thread per 4-byte (or 16-byte) piece, so the
// writes are coalesced and a row's pieces are read by neighbouring threads
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src, int pieces, const int* __restrict__ idx, long long n,
                                   T* __restrict__ dst) {
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x, total = n * pieces;
  for (; t < total; t += step) {
    const long long i = t / pieces;
    const int c = (int)(t - i * pieces);
    dst[t] = __ldg(src + (long long)__ldg(idx + i) * pieces + c);
  }
}

// --- bricks (see brick_window_find in common.cuh): full-resolution row of every cell of every tensor-stride-4 voxel
__global__ void brick_rows_kernel(const int4* __restrict__ coords, const int* __restrict__ parent01,
                                  const int* __restrict__ parent12, long long n, int* __restrict__ brick_rows) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long step = (long long)gridDim.x * blockDim.x;
  for (; i < n; i += step) {
    const int4 c = __ldg(coords + i);
    const int b = __ldg(parent12 + __ldg(parent01 + i));
    brick_rows[(long long)b * 64 + ((c.y & 3) | ((c.z & 3) << 2) | ((c.w & 3) << 4))] = (int)i;
  }
}

}  // namespace ag3d

using namespace ag3d;

extern "C" {

int ag3d_abi_version(void) { return AG3D_ABI_VERSION; }
int64_t ag3d_kernel_launches(void) { return g_kernel_launches.load(std::memory_order_relaxed); }
const char* ag3d_last_error(void) { return g_last_error.c_str(); }

int ag3d_device_info(int32_t* sms, int32_t* major, int32_t* minor) {
  int dev = 0;
  AG3D_CUDA(cudaGetDevice(&dev));
  int a = 0, b = 0, c = 0;
  AG3D_CUDA(cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev));
  AG3D_CUDA(cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev));
  AG3D_CUDA(cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev));
  if (sms) *sms = a;
  if (major) *major = b;
  if (minor) *minor = c;
  return AG3D_OK;
}

int64_t ag3d_hash_capacity(int64_t n) {
  int64_t cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  return cap;
}

static int check_table(const void* table, int64_t cap, int64_t n) {
  AG3D_CHECK_ARG(table && aligned16(table), "hash table must be a 16-byte aligned device pointer");
  AG3D_CHECK_ARG(cap >= 2 * n && cap >= 2 && (cap & (cap - 1)) == 0, "hash capacity must be a power of two >= 2n");
  return AG3D_OK;
}

int ag3d_hash_build(const int32_t* coords, int64_t n, void* table, int64_t cap, int32_t* status,
                    ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n >= 0 && n < INT_MAX, "row count out of range");
  if (int rc = check_table(table, cap, n)) return rc;
  AG3D_CHECK_ARG(status, "status must be a device int32[2]");
  AG3D_CHECK_ARG(n == 0 || (coords && aligned16(coords)), "coords must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  table_clear_kernel<<<grid_for(cap, 256), 256, 0, st>>>(static_cast<Slot*>(table), cap);
  AG3D_LAUNCH_CHECK("table_clear");
  if (n == 0) return AG3D_OK;
  hash_build_kernel<<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const int4*>(coords), n,
                                                      static_cast<Slot*>(table), (unsigned long long)(cap - 1), status);
  AG3D_LAUNCH_CHECK("hash_build");
  return AG3D_OK;
}

size_t ag3d_downsample_workspace_bytes(int64_t n) {
  int64_t blocks = (n + SCAN_TILE - 1) / SCAN_TILE + 1;
  return (size_t)(n + blocks + 8) * sizeof(int32_t);
}

static int downsample_impl(const int32_t* coords, int64_t n, const int32_t* n_dev, int32_t new_stride, void* coarse_table,
                           int64_t cap, int32_t* parent, int32_t* out_coords, int32_t* out_n, void* ws, size_t ws_bytes,
                           ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n > 0 && n < INT_MAX, "row count out of range");
  AG3D_CHECK_ARG(new_stride >= 1, "new_stride must be >= 1 (1 = unique-ify in first-occurrence order)");
  if (int rc = check_table(coarse_table, cap, n)) return rc;
  AG3D_CHECK_ARG(coords && aligned16(coords) && out_coords && aligned16(out_coords), "coords must be 16-byte aligned");
  AG3D_CHECK_ARG(parent && out_n, "parent / out_n missing");
  if (ws_bytes < ag3d_downsample_workspace_bytes(n) || !ws) {
    set_error("downsample workspace too small");
    return AG3D_E_WORKSPACE;
  }
  cudaStream_t st = as_stream(stream);
  Slot* table = static_cast<Slot*>(coarse_table);
  int* slot_of_row = static_cast<int*>(ws);
  int* block_sums = slot_of_row + n;
  const int n_blocks = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  const unsigned long long mask = (unsigned long long)(cap - 1);
  const int4* c4 = reinterpret_cast<const int4*>(coords);
  table_clear_kernel<<<grid_for(cap, 256), 256, 0, st>>>(table, cap);
  AG3D_LAUNCH_CHECK("table_clear");
  coarse_insert_kernel<<<grid_for(n, 256), 256, 0, st>>>(c4, n, n_dev, new_stride, table, mask, slot_of_row);
  AG3D_LAUNCH_CHECK("coarse_insert");
  rep_count_kernel<<<n_blocks, SCAN_THREADS, 0, st>>>(table, slot_of_row, n, n_dev, block_sums);
  AG3D_LAUNCH_CHECK("rep_count");
  block_scan_kernel<<<1, 1024, 0, st>>>(block_sums, n_blocks, out_n);
  AG3D_LAUNCH_CHECK("block_scan");
  rep_assign_kernel<<<n_blocks, SCAN_THREADS, 0, st>>>(c4, n, n_dev, new_stride, table, slot_of_row, block_sums,
                                                       reinterpret_cast<int4*>(out_coords));
  AG3D_LAUNCH_CHECK("rep_assign");
  parent_kernel<<<grid_for(n, 256), 256, 0, st>>>(table, slot_of_row, n, n_dev, parent);
  AG3D_LAUNCH_CHECK("parent");
  return AG3D_OK;
}

int ag3d_downsample(const int32_t* coords, int64_t n, int32_t new_stride, void* coarse_table, int64_t cap,
                    int32_t* parent, int32_t* out_coords, int32_t* out_n, void* ws, size_t ws_bytes,
                    ag3d_stream_t stream) {
  return downsample_impl(coords, n, nullptr, new_stride, coarse_table, cap, parent, out_coords, out_n, ws, ws_bytes, stream);
}

int ag3d_downsample_dev(const int32_t* coords, int64_t n_max, const int32_t* n_dev, int32_t new_stride, void* coarse_table,
                        int64_t cap, int32_t* parent, int32_t* out_coords, int32_t* out_n, void* ws, size_t ws_bytes,
                        ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_dev, "n_dev missing");
  return downsample_impl(coords, n_max, n_dev, new_stride, coarse_table, cap, parent, out_coords, out_n, ws, ws_bytes, stream);
}

int ag3d_scene_offsets(const int32_t* coords, int64_t n, int32_t max_scenes, int32_t* offsets, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n > 0 && n < INT_MAX && max_scenes >= 1 && max_scenes <= 65534, "row / scene count out of range");
  AG3D_CHECK_ARG(coords && aligned16(coords) && offsets, "bad pointers");
  cudaStream_t st = as_stream(stream);
  AG3D_CUDA(cudaMemsetAsync(offsets + max_scenes + 1, 0, sizeof(int32_t), st));
  scene_offsets_kernel<<<(max_scenes + 1 + 127) / 128, 128, 0, st>>>(reinterpret_cast<const int4*>(coords), n, max_scenes, offsets);
  AG3D_LAUNCH_CHECK("scene_offsets");
  scene_order_kernel<<<grid_for(n, 256), 256, 0, st>>>(reinterpret_cast<const int4*>(coords), n, offsets + max_scenes + 1);
  AG3D_LAUNCH_CHECK("scene_order");
  return AG3D_OK;
}

size_t ag3d_row_order_workspace_bytes(int64_t n) {
  size_t temp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, temp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const int*)nullptr, (int*)nullptr, (int)n, 0, 48);
  return (size_t)n * (8 + 8 + 4) + temp + 1024;
}

int ag3d_row_order(const int32_t* nbr, int32_t K, int64_t n, const int32_t* coords, int32_t* perm, int32_t* inv, void* ws,
                   size_t ws_bytes, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n > 0 && n < INT_MAX && K >= 1 && K <= 32, "row_order: shape");
  AG3D_CHECK_ARG(nbr && coords && aligned16(coords) && perm && inv, "row_order: pointers");
  AG3D_CHECK_ARG(ws && aligned16(ws) && ws_bytes >= ag3d_row_order_workspace_bytes(n), "row_order: workspace too small");
  cudaStream_t st = as_stream(stream);
  unsigned char* w = static_cast<unsigned char*>(ws);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(w);
  unsigned long long* keys_out = keys + n;
  int* vals = reinterpret_cast<int*>(keys_out + n);
  void* temp = w + (((size_t)n * 20 + 255) & ~(size_t)255);
  size_t temp_bytes = ws_bytes - (((size_t)n * 20 + 255) & ~(size_t)255);
  order_key_kernel<<<grid_for(n, 256), 256, 0, st>>>(nbr, K, n, reinterpret_cast<const int4*>(coords), keys, vals);
  AG3D_LAUNCH_CHECK("order_key");
  // stable LSD radix sort over (scene, mask): equal patterns keep the caller's order
  AG3D_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_out, vals, perm, (int)n, 0, 48, st));
  g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
  invert_perm_kernel<<<grid_for(n, 256), 256, 0, st>>>(perm, n, inv);
  AG3D_LAUNCH_CHECK("invert_perm");
  return AG3D_OK;
}

int ag3d_permute_map(const int32_t* nbr, int32_t K, int64_t n_out, const int32_t* perm_out, const int32_t* inv_in,
                     int32_t* out, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(nbr && out && nbr != out && n_out > 0 && n_out < INT_MAX && K >= 1, "permute_map: arguments");
  permute_map_kernel<<<dim3(grid_for(n_out, 256), K), 256, 0, as_stream(stream)>>>(nbr, n_out, perm_out, inv_in, out);
  AG3D_LAUNCH_CHECK("permute_map");
  return AG3D_OK;
}

int ag3d_kernel_map(const int32_t* out_coords, int64_t n_out, const void* in_table, int64_t cap, int32_t ksize,
                    int32_t in_tensor_stride, int32_t dilation, int32_t* nbr, int32_t* pair_count,
                    ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_out > 0 && n_out < INT_MAX, "row count out of range");
  AG3D_CHECK_ARG(ksize >= 1 && ksize <= 7, "kernel size must be 1..7");
  AG3D_CHECK_ARG(in_tensor_stride >= 1 && dilation >= 1, "stride/dilation must be >= 1");
  AG3D_CHECK_ARG(in_table && aligned16(in_table) && cap >= 2 && (cap & (cap - 1)) == 0, "bad hash table");
  AG3D_CHECK_ARG(out_coords && aligned16(out_coords) && nbr, "bad pointers");
  const int K = ksize * ksize * ksize;
  dim3 grid(grid_for(n_out, 256), K);
  kernel_map_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const int4*>(out_coords), n_out,
                                                         static_cast<const Slot*>(in_table),
                                                         (unsigned long long)(cap - 1), ksize,
                                                         in_tensor_stride * dilation, nbr, pair_count);
  AG3D_LAUNCH_CHECK("kernel_map");
  return AG3D_OK;
}

int ag3d_kernel_map_transposed(const int32_t* fine_coords, const int32_t* parent, int64_t n_fine,
                               int32_t fine_stride, int32_t* nbr, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n_fine > 0 && n_fine < INT_MAX, "row count out of range");
  AG3D_CHECK_ARG(fine_stride >= 1, "fine_stride must be >= 1");
  AG3D_CHECK_ARG(fine_coords && aligned16(fine_coords) && parent && nbr, "bad pointers");
  kernel_map_transposed_kernel<<<grid_for(n_fine, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const int4*>(fine_coords), parent, n_fine, fine_stride, nbr);
  AG3D_LAUNCH_CHECK("kernel_map_transposed");
  return AG3D_OK;
}

int ag3d_gather_rows(const void* src, int32_t row_bytes, const int32_t* idx, int64_t n, void* dst, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(src && idx && dst && src != dst && n > 0 && n < INT_MAX && row_bytes > 0 && row_bytes % 4 == 0,
                 "gather_rows: arguments");
  if (row_bytes % 16 == 0 && aligned16(src) && aligned16(dst))
    gather_rows_kernel<int4><<<grid_for(n * (row_bytes / 16), 256), 256, 0, as_stream(stream)>>>(
        static_cast<const int4*>(src), row_bytes / 16, idx, n, static_cast<int4*>(dst));
  else
    gather_rows_kernel<int><<<grid_for(n * (row_bytes / 4), 256), 256, 0, as_stream(stream)>>>(
        static_cast<const int*>(src), row_bytes / 4, idx, n, static_cast<int*>(dst));
  AG3D_LAUNCH_CHECK("gather_rows");
  return AG3D_OK;
}

int ag3d_brick_rows(const int32_t* coords, const int32_t* parent01, const int32_t* parent12, int64_t n,
                    int64_t n_bricks, int32_t* brick_rows, ag3d_stream_t stream) {
  AG3D_CHECK_ARG(n > 0 && n < INT_MAX && n_bricks > 0 && n_bricks < INT_MAX / 64, "row count out of range");
  AG3D_CHECK_ARG(coords && aligned16(coords) && parent01 && parent12 && brick_rows, "bad pointers");
  AG3D_CUDA(cudaMemsetAsync(brick_rows, 0xFF, (size_t)n_bricks * 64 * sizeof(int32_t), as_stream(stream)));
  brick_rows_kernel<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(reinterpret_cast<const int4*>(coords), parent01,
                                                                     parent12, n, brick_rows);
  AG3D_LAUNCH_CHECK("brick_rows");
  return AG3D_OK;
}

}  // extern "C"
