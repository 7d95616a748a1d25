// Single-block exclusive scan shared by the ingest-side kernels (id_remap.cu, sample.cu).
#pragma once
#include "common.cuh"

namespace gr {

// exclusive scan of `count` ints in place by ONE block (ingest-time only; ~40 ms at 500M elements)
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

}  // namespace gr
