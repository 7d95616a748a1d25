// Exclusive scans shared by the ingest-side kernels (id_remap.cu, sample.cu): one block for short arrays, three
// passes (tile sums -> scan of the sums -> tile scan + offset) for long ones.
#pragma once
#include "common.cuh"

namespace gr {

// exclusive scan of `count` ints in place by ONE block (~1 us per 4096 elements: short arrays only)
static __global__ void __launch_bounds__(1024) scan1_kernel(int* __restrict__ data, long long count, int* __restrict__ total) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  constexpr int PER = 4;
  for (long long base = 0; base < count; base += 1024 * PER) {
    int v[PER];
    int sum = 0;
    const long long j0 = base + (long long)threadIdx.x * PER;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      v[i] = j0 + i < count ? data[j0 + i] : 0;
      sum += v[i];
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(gr::FULL, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    if (w == 0) {
      int ws = s_warp[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(gr::FULL, ws, o);
        if (lane >= o) ws += t;
      }
      s_warp[lane] = ws;
    }
    __syncthreads();
    int run = s_carry + (w > 0 ? s_warp[w - 1] : 0) + incl - sum;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (j0 + i < count) data[j0 + i] = run;
      run += v[i];
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = run;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

constexpr int SCAN_TILE = 4096;  // elements per block of the multi-block scan (1024 threads x 4)

static __global__ void __launch_bounds__(1024) scan_tile_sums_kernel(const int* __restrict__ data, long long count,
                                                                     int* __restrict__ tile_sums) {
  __shared__ int s_warp[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long j0 = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * 4;
  int sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) sum += j0 + i < count ? data[j0 + i] : 0;
  sum = __reduce_add_sync(FULL, sum);
  if (lane == 0) s_warp[w] = sum;
  __syncthreads();
  if (w == 0) {
    const int t = __reduce_add_sync(FULL, s_warp[lane]);
    if (lane == 0) tile_sums[blockIdx.x] = t;
  }
}

static __global__ void __launch_bounds__(1024) scan_tiles_kernel(int* __restrict__ data, long long count,
                                                                 const int* __restrict__ tile_offsets) {
  __shared__ int s_warp[32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long j0 = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * 4;
  int v[4];
  int sum = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = j0 + i < count ? data[j0 + i] : 0;
    sum += v[i];
  }
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[w] = incl;
  __syncthreads();
  if (w == 0) {
    int ws = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(FULL, ws, o);
      if (lane >= o) ws += t;
    }
    s_warp[lane] = ws;
  }
  __syncthreads();
  int run = tile_offsets[blockIdx.x] + (w > 0 ? s_warp[w - 1] : 0) + incl - sum;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (j0 + i < count) data[j0 + i] = run;
    run += v[i];
  }
}

inline size_t scan_workspace_bytes(long long count) {
  return align_up(sizeof(int) * (size_t)((count + SCAN_TILE - 1) / SCAN_TILE + 1), 256);
}

// exclusive scan of `count` ints in place, *total = their sum; `tiles` = scan_workspace_bytes(count) of scratch
// (may be null for count <= 4 * SCAN_TILE, which one block handles). Integer adds: any order gives the same result.
inline cudaError_t scan_exclusive_i32(int* data, long long count, int* total, int* tiles, cudaStream_t st) {
  if (count <= 4 * SCAN_TILE || tiles == nullptr) {
    scan1_kernel<<<1, 1024, 0, st>>>(data, count, total);
    count_launch();
    return cudaGetLastError();
  }
  const long long n_tiles = (count + SCAN_TILE - 1) / SCAN_TILE;
  scan_tile_sums_kernel<<<(unsigned)n_tiles, 1024, 0, st>>>(data, count, tiles);
  count_launch();
  scan1_kernel<<<1, 1024, 0, st>>>(tiles, n_tiles, total);
  count_launch();
  scan_tiles_kernel<<<(unsigned)n_tiles, 1024, 0, st>>>(data, count, tiles);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gr
