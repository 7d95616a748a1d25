// Item sweep order for the scoring kernel (stage 1 of get_recs, src/metrics.py:52-77).
//
// The fused shortlist epilogue of gr_score_topk_tc pays for every candidate that beats a user's running threshold. In
// item-id order a user's threshold climbs slowly (S * (1 + ln(I / S)) inserts per sweep); when the items every user is
// likely to score high come FIRST, the thresholds are near their final value after a few tiles and the rest of the
// sweep is the insert-free pipeline. Embedding tables of a recommender share a strong common component (popular items
// score high for everybody), so the order used is: descending cosine to the mean normalised user row `dir`. The order
// only changes WHEN an item is looked at, never the result -- the shortlist keeps real item ids and stage 2 re-scores
// and proves exactly as before.
//
//   gr_score_item_order:  perm[p] = index of the item swept at position p. Counting sort (the LSD radix sort behind
//                         gr_csr_build_i32, stable: ascending index inside a bucket) of 65536 equal-width buckets
//                         between the smallest and the largest cosine.
//   gr_permute_rows:      dst[p] = src[perm[p]] for rows of row_bytes (multiple of 16): the permuted operand table.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "common.cuh"

namespace {

constexpr int ORDER_BUCKETS = 65536;

// monotone float -> int map (and back): integer min / max atomics then order like the floats
__device__ __forceinline__ int ordered_int(float f) {
  const int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

__global__ void order_init_kernel(int* minmax) {
  minmax[0] = 0x7fffffff;          // running min
  minmax[1] = (int)0x80000000u;    // running max
}

// one warp per item row: cos(h_item[i], dir) up to the (positive) norm of dir; 0 for a NaN so that it cannot poison
// the range
__global__ void __launch_bounds__(256) order_cosine_kernel(const float* __restrict__ h_item, long long n, int d,
                                                           const float* __restrict__ dir, float* __restrict__ cosv,
                                                           int* __restrict__ minmax) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  float lo = INFINITY, hi = -INFINITY;
  for (long long i = warp0; i < n; i += n_warps) {
    const float* row = h_item + i * d;
    float dot = 0.f, sq = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float x = gr::ldg_stream_f32(row + c);
      dot = fmaf(x, __ldg(dir + c), dot);
      sq = fmaf(x, x, sq);
    }
    dot = gr::warp_sum(dot);
    sq = gr::warp_sum(sq);
    float v = dot / fmaxf(sqrtf(sq), 1e-12f);
    if (!(v == v)) v = 0.f;
    v = fminf(fmaxf(v, -3.0e38f), 3.0e38f);
    if (lane == 0) cosv[i] = v;
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  if (lane == 0 && lo <= hi) {
    atomicMin(minmax, ordered_int(lo));
    atomicMax(minmax + 1, ordered_int(hi));
  }
}

// bucket 0 = the largest cosine
__global__ void __launch_bounds__(256) order_bucket_kernel(const float* __restrict__ cosv, long long n,
                                                           const int* __restrict__ minmax, int* __restrict__ keys) {
  const float lo = ordered_float(minmax[0]), hi = ordered_float(minmax[1]);
  const float scale = hi > lo ? (float)(ORDER_BUCKETS - 1) / (hi - lo) : 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float b = (hi - cosv[i]) * scale;
    keys[i] = (int)fminf(fmaxf(b, 0.f), (float)(ORDER_BUCKETS - 1));
  }
}

__global__ void __launch_bounds__(256) permute_rows_kernel(const uint4* __restrict__ src, long long n_rows, int vec_per_row,
                                                           const int* __restrict__ perm, uint4* __restrict__ dst) {
  const long long total = n_rows * vec_per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vec_per_row;
    const int c = (int)(i - r * vec_per_row);
    dst[i] = __ldg(src + (long long)__ldg(perm + r) * vec_per_row + c);
  }
}

struct OrderWs { size_t cosv, keys, indptr, dummy, minmax, csr, total; };
OrderWs order_plan(int64_t n) {
  OrderWs w;
  size_t o = 0;
  w.cosv = o; o += gr::align_up((size_t)n * 4, 256);
  w.keys = o; o += gr::align_up((size_t)n * 4, 256);
  w.indptr = o; o += gr::align_up((size_t)(ORDER_BUCKETS + 1) * 4, 256);
  w.dummy = o; o += gr::align_up((size_t)n * 4, 256);
  w.minmax = o; o += 256;
  w.csr = o; o += gr_csr_build_workspace_bytes(n, ORDER_BUCKETS);
  w.total = o;
  return w;
}

}  // namespace

extern "C" size_t gr_score_item_order_workspace_bytes(int64_t n_items) {
  return n_items <= 0 ? 256 : order_plan(n_items).total;
}

extern "C" int gr_score_item_order(const float* h_item, int64_t n_items, int32_t d, const float* dir, int32_t* perm,
                                   void* ws, size_t ws_bytes, gr_stream_t stream) {
  GR_REQUIRE(n_items >= 0 && n_items <= 0x7fffffffLL && d >= 1, GR_E_INVALID, "bad size");
  if (n_items == 0) return GR_OK;
  GR_REQUIRE(h_item && dir && perm, GR_E_INVALID, "null pointer");
  const OrderWs w = order_plan(n_items);
  GR_REQUIRE(ws != nullptr && ws_bytes >= w.total, GR_E_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* base = static_cast<char*>(ws);
  float* cosv = reinterpret_cast<float*>(base + w.cosv);
  int* keys = reinterpret_cast<int*>(base + w.keys);
  int* minmax = reinterpret_cast<int*>(base + w.minmax);
  order_init_kernel<<<1, 1, 0, st>>>(minmax);
  GR_LAUNCH_CHECK();
  const int blocks = (int)std::min<long long>((n_items + 7) / 8, (long long)gr::sm_count() * 8);
  order_cosine_kernel<<<blocks, 256, 0, st>>>(h_item, n_items, d, dir, cosv, minmax);
  GR_LAUNCH_CHECK();
  const int blocks2 = (int)std::min<long long>((n_items + 255) / 256, (long long)gr::sm_count() * 8);
  order_bucket_kernel<<<blocks2, 256, 0, st>>>(cosv, n_items, minmax, keys);
  GR_LAUNCH_CHECK();
  // stable counting sort by bucket: eperm[p] = index of the p-th item in (bucket, index) order
  return gr_csr_build_i32(keys, keys, n_items, ORDER_BUCKETS, reinterpret_cast<int32_t*>(base + w.indptr),
                          reinterpret_cast<int32_t*>(base + w.dummy), perm, nullptr, base + w.csr,
                          ws_bytes - w.csr, stream);
}

extern "C" int gr_permute_rows(const void* src, int64_t n_rows, int64_t row_bytes, const int32_t* perm, void* dst,
                               gr_stream_t stream) {
  GR_REQUIRE(n_rows >= 0 && row_bytes > 0 && row_bytes % 16 == 0 && row_bytes <= (1 << 20), GR_E_INVALID,
             "row_bytes must be a positive multiple of 16");
  if (n_rows == 0) return GR_OK;
  GR_REQUIRE(src && perm && dst && src != dst, GR_E_INVALID, "null pointer or in-place permutation");
  GR_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, GR_E_INVALID,
             "rows must be 16-byte aligned");
  const int vec = (int)(row_bytes / 16);
  const long long total = n_rows * vec;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)gr::sm_count() * 16);
  permute_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), n_rows, vec, perm, static_cast<uint4*>(dst));
  GR_LAUNCH_CHECK();
  return GR_OK;
}
