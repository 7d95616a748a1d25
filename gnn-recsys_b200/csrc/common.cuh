// Shared device/host helpers for the gnn_recsys_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/gnn_recsys_b200.h"

namespace gr {

void set_error(const std::string& msg);
void count_launch();  // every kernel launch of the library bumps gr_launch_count()

#define GR_REQUIRE(cond, code, msg)                                                   \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      ::gr::set_error(std::string(__func__) + ": " + (msg));                          \
      return (code);                                                                  \
    }                                                                                 \
  } while (0)

#define GR_CUDA(expr)                                                                 \
  do {                                                                                \
    cudaError_t e__ = (expr);                                                         \
    if (e__ != cudaSuccess) {                                                         \
      ::gr::set_error(std::string(__func__) + ": " #expr ": " + cudaGetErrorString(e__)); \
      return GR_E_CUDA;                                                               \
    }                                                                                 \
  } while (0)

#define GR_LAUNCH_CHECK()                                                             \
  do {                                                                                \
    ::gr::count_launch();                                                             \
    cudaError_t e__ = cudaGetLastError();                                             \
    if (e__ != cudaSuccess) {                                                         \
      ::gr::set_error(std::string(__func__) + ": kernel launch: " + cudaGetErrorString(e__)); \
      return GR_E_CUDA;                                                               \
    }                                                                                 \
  } while (0)

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float4 ldg_f4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// streaming (read-once) loads: keep them out of L1
__device__ __forceinline__ int ldg_stream_i32(const int* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace gr
