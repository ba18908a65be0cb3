// Shared helpers for the agile3d_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/agile3d_b200.h"

namespace ag3d {

extern std::atomic<long long> g_kernel_launches;
void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

#define AG3D_CHECK_ARG(cond, msg)                       \
  do {                                                  \
    if (!(cond)) {                                      \
      ::ag3d::set_error(std::string("invalid: ") + msg); \
      return AG3D_E_INVALID;                            \
    }                                                   \
  } while (0)

#define AG3D_CUDA(call)                                        \
  do {                                                         \
    cudaError_t e__ = (call);                                  \
    if (e__ != cudaSuccess) return ::ag3d::cuda_fail(e__, #call); \
  } while (0)

#define AG3D_LAUNCH_CHECK(name)                                   \
  do {                                                            \
    ::ag3d::g_kernel_launches.fetch_add(1, std::memory_order_relaxed); \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) return ::ag3d::cuda_fail(e__, name);  \
  } while (0)

inline cudaStream_t as_stream(ag3d_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
int sm_count();

// ---------------------------------------------------------------------------------------------- hash table
// 16-byte slot: one 128-bit load per probe.
struct __align__(16) Slot {
  unsigned long long key;  // packed (b,x,y,z); EMPTY_KEY when free
  int first_row;           // smallest source row that produced this key (first occurrence)
  int row;                 // row id of this coordinate in its own level
};
static_assert(sizeof(Slot) == 16, "slot must be 16 bytes");

constexpr unsigned long long EMPTY_KEY = 0xFFFFFFFFFFFFFFFFull;
constexpr int COORD_LIMIT = 32768;

__host__ __device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
  return b >= 0 && b < 65535 && x > -COORD_LIMIT && x < COORD_LIMIT && y > -COORD_LIMIT && y < COORD_LIMIT &&
         z > -COORD_LIMIT && z < COORD_LIMIT;
}

__host__ __device__ __forceinline__ unsigned long long pack_key(int b, int x, int y, int z) {
  return (static_cast<unsigned long long>(static_cast<unsigned>(b) & 0xFFFFu) << 48) |
         (static_cast<unsigned long long>(static_cast<unsigned>(x + COORD_LIMIT) & 0xFFFFu) << 32) |
         (static_cast<unsigned long long>(static_cast<unsigned>(y + COORD_LIMIT) & 0xFFFFu) << 16) |
         static_cast<unsigned long long>(static_cast<unsigned>(z + COORD_LIMIT) & 0xFFFFu);
}

__host__ __device__ __forceinline__ unsigned long long mix64(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

// Read-only probe; returns the slot's row or -1.
__device__ __forceinline__ int table_find(const Slot* __restrict__ table, unsigned long long mask,
                                          unsigned long long key) {
  unsigned long long s = mix64(key) & mask;
  while (true) {
    const int4 raw = __ldg(reinterpret_cast<const int4*>(table + s));
    const unsigned long long k =
        (static_cast<unsigned long long>(static_cast<unsigned>(raw.y)) << 32) | static_cast<unsigned>(raw.x);
    if (k == key) return raw.w;
    if (k == EMPTY_KEY) return -1;
    s = (s + 1) & mask;
  }
}

// ---------------------------------------------------------------------------------------------- bricks
// A "brick" is the 4x4x4 block of full-resolution cells under one tensor-stride-4 voxel (a level-2 row): brick_rows
// [n_level2][64] holds the full-resolution row of every cell (bit = x&3 | (y&3)<<2 | (z&3)<<4), -1 = empty.  A window
// of +-2 cells touches exactly 2x2x2 bricks, so the 125 neighbours of the stem cost 8 probes of the (small, cache
// resident) level-2 table plus 125 four-byte reads out of eight 256-byte lines instead of 125 random 16-byte probes.
// Warp-cooperative: lanes 0..7 probe, every lane then resolves its own offsets k = j*32 + lane (j < 4).
__device__ __forceinline__ void brick_window_find(const int4 c, int lane, int ksize, int K, const Slot* __restrict__ table2,
                                                  unsigned long long mask2, const int* __restrict__ brick_rows,
                                                  int (&src)[4]) {
  const int half = ksize / 2;
  const int bx0 = (c.y - half) >> 2, by0 = (c.z - half) >> 2, bz0 = (c.w - half) >> 2;   // arithmetic shift = floor
  int brick = -1;
  if (lane < 8)
    brick = table_find(table2, mask2, pack_key(c.x, (bx0 + (lane & 1)) * 4, (by0 + ((lane >> 1) & 1)) * 4,
                                               (bz0 + (lane >> 2)) * 4));
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = j * 32 + lane;
    int r = k;
    const int jx = r % ksize; r /= ksize;
    const int jy = r % ksize; r /= ksize;
    const int x = c.y + jx - half, y = c.z + jy - half, z = c.w + r - half;
    const int bi = (((x >> 2) - bx0) & 1) | ((((y >> 2) - by0) & 1) << 1) | ((((z >> 2) - bz0) & 1) << 2);
    const int bid = __shfl_sync(0xffffffffu, brick, bi);
    src[j] = -1;
    if (k < K && bid >= 0 && coord_in_range(c.x, x, y, z))
      src[j] = __ldg(brick_rows + (long long)bid * 64 + ((x & 3) | ((y & 3) << 2) | ((z & 3) << 4)));
  }
}

__device__ __forceinline__ int floor_div(int a, int b) {  // b > 0
  int q = a / b;
  return (a % b != 0 && a < 0) ? q - 1 : q;
}

}  // namespace ag3d
