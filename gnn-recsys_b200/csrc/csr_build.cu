// gr_csr_build_i32: stable COO -> int32 CSR over destination rows on the device.
//
// Replaces what dgl.heterograph + DGL's lazy CSC construction do on the CPU behind update_all
// (reference src/builder.py:377-383, src/model.py:145-147): neighbours of a destination row must stay in edge-id
// order, so the build is a *stable* sort of the edge list by destination id -- a least-significant-digit radix sort
// (8-bit digits, ceil(log2(n_dst) / 8) passes) of (dst, edge id) pairs, then
//   eperm[j]   = edge id held by CSR slot j
//   indices[j] = src[eperm[j]]
//   indptr[v]  = first slot whose destination is >= v
// Bit-exact against oracle.straightline.csr_by_dst / numpy's stable argsort (tests/test_gpu_parity.py).
//
// One pass = three kernels: per-tile digit histograms (bin-major), one exclusive scan over bins x tiles, and a
// scatter in which every warp owns a contiguous slice of its tile and ranks elements with __match_any_sync, so the
// order among equal digits is preserved without atomics.
#include <algorithm>

#include "common.cuh"

namespace {

using gr::FULL;
constexpr int RADIX_BITS = 8;
constexpr int BINS = 1 << RADIX_BITS;
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;
constexpr int ITEMS = 16;                      // elements per thread
constexpr int TILE = THREADS * ITEMS;          // 4096 elements per CTA
constexpr int WARP_SPAN = 32 * ITEMS;          // contiguous elements per warp

__global__ void __launch_bounds__(THREADS) radix_hist_kernel(const int* __restrict__ keys, long long n, int shift,
                                                             int n_tiles, int* __restrict__ hist /*[BINS][n_tiles]*/) {
  __shared__ int s_hist[BINS];
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)tile * TILE;
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const long long j = base + i * THREADS + threadIdx.x;
      if (j < n) atomicAdd(&s_hist[(keys[j] >> shift) & (BINS - 1)], 1);
    }
    __syncthreads();
    hist[(long long)threadIdx.x * n_tiles + tile] = s_hist[threadIdx.x];
    __syncthreads();
  }
}

// exclusive scan of `count` ints in place by ONE block (count = BINS x n_tiles, <= ~32M for 500M edges: ~1 ms)
__global__ void __launch_bounds__(1024) scan_kernel(int* __restrict__ data, long long count) {
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
      s_warp[lane] = ws;  // inclusive over warps
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
}

template <bool FIRST>
__global__ void __launch_bounds__(THREADS) radix_scatter_kernel(const int* __restrict__ keys_in,
                                                                const int* __restrict__ vals_in, long long n, int shift,
                                                                int n_tiles, const int* __restrict__ offsets,
                                                                int* __restrict__ keys_out, int* __restrict__ vals_out) {
  __shared__ int s_cnt[WARPS][BINS];   // per-warp running digit counts, then exclusive bases over warps
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int i = threadIdx.x; i < WARPS * BINS; i += THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const long long base = (long long)tile * TILE + (long long)w * WARP_SPAN;
    int key[ITEMS], val[ITEMS], rank[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {  // group i of this warp = 32 consecutive elements
      const long long j = base + i * 32 + lane;
      const bool ok = j < n;
      key[i] = ok ? keys_in[j] : 0x7fffffff;
      val[i] = ok ? (FIRST ? (int)j : vals_in[j]) : 0;
      const int dgt = ok ? (key[i] >> shift) & (BINS - 1) : BINS;  // BINS = inactive
      const unsigned peers = __match_any_sync(FULL, dgt);
      const int before = __popc(peers & ((1u << lane) - 1u));
      int old = 0;
      if (ok) old = s_cnt[w][dgt];
      __syncwarp();
      if (ok && before == 0) s_cnt[w][dgt] = old + __popc(peers);
      __syncwarp();
      rank[i] = old + before;
    }
    __syncthreads();
    {  // exclusive prefix over warps per digit, plus the global offset of (digit, tile)
      const int dgt = threadIdx.x;
      int run = offsets[(long long)dgt * n_tiles + tile];
#pragma unroll
      for (int ww = 0; ww < WARPS; ++ww) {
        const int c = s_cnt[ww][dgt];
        s_cnt[ww][dgt] = run;
        run += c;
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
      const long long j = base + i * 32 + lane;
      if (j < n) {
        const int pos = s_cnt[w][(key[i] >> shift) & (BINS - 1)] + rank[i];
        keys_out[pos] = key[i];
        vals_out[pos] = val[i];
      }
    }
    __syncthreads();
  }
}

// status[0] |= 1 when some destination id lies outside [0, n_dst): the caller's input is invalid (the host twin
// csr_by_dst_host raises IndexError for it). The build itself stays memory-safe for such input (csr_finish_kernel clamps).
__global__ void csr_validate_kernel(const int* __restrict__ dst, long long nnz, int n_dst, int* __restrict__ status) {
  bool bad = false;
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < nnz; j += (long long)gridDim.x * blockDim.x) {
    const int v = dst[j];
    bad |= v < 0 || v >= n_dst;
  }
  if (__any_sync(FULL, bad) && (threadIdx.x & 31) == 0) atomicOr(status, 1);
}

__global__ void csr_finish_kernel(const int* __restrict__ sorted_dst, const int* __restrict__ eperm,
                                  const int* __restrict__ src, long long nnz, int n_dst, int* __restrict__ indptr,
                                  int* __restrict__ indices) {
  for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < nnz; j += (long long)gridDim.x * blockDim.x) {
    indices[j] = src[eperm[j]];
    // clamped: an out-of-range destination id (flagged by csr_validate_kernel) must never index outside indptr
    const int kcur = min(max(sorted_dst[j], -1), n_dst - 1);
    const int kprev = j > 0 ? min(max(sorted_dst[j - 1], -1), n_dst - 1) : -1;
    for (int v = kprev + 1; v <= kcur; ++v) indptr[v] = (int)j;
    if (j == nnz - 1)
      for (int v = kcur + 1; v <= n_dst; ++v) indptr[v] = (int)nnz;
  }
}

__global__ void fill_kernel(int* p, long long n, int v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

int n_passes(int n_dst) {
  int bits = 1;
  while (bits < 31 && (1LL << bits) < (long long)n_dst) ++bits;
  return (bits + RADIX_BITS - 1) / RADIX_BITS;
}

struct Layout { size_t keys_a, keys_b, vals_b, hist, total; };
Layout layout(int64_t nnz) {
  const int64_t n_tiles = (nnz + TILE - 1) / TILE;
  Layout l;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = gr::align_up(off + b, 256); return o; };
  l.keys_a = take(4 * (size_t)nnz);
  l.keys_b = take(4 * (size_t)nnz);
  l.vals_b = take(4 * (size_t)nnz);
  l.hist = take(4 * (size_t)BINS * (size_t)std::max<int64_t>(n_tiles, 1));
  l.total = std::max<size_t>(off, 256);
  return l;
}

}  // namespace

extern "C" size_t gr_csr_build_workspace_bytes(int64_t nnz, int32_t n_dst) {
  (void)n_dst;
  return layout(nnz < 0 ? 0 : nnz).total;
}

extern "C" int gr_csr_build_i32(const int32_t* src, const int32_t* dst, int64_t nnz, int32_t n_dst, int32_t* indptr,
                                int32_t* indices, int32_t* eperm, int32_t* status_or_null, void* ws, size_t ws_bytes,
                                gr_stream_t stream) {
  GR_REQUIRE(nnz >= 0 && n_dst >= 0, GR_E_INVALID, "negative size");
  GR_REQUIRE(nnz <= 0x7fffffffLL, GR_E_INVALID, "int32 CSR cannot index more than 2^31 - 1 edges");
  GR_REQUIRE(indptr != nullptr, GR_E_INVALID, "null indptr");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (status_or_null != nullptr) GR_CUDA(cudaMemsetAsync(status_or_null, 0, sizeof(int32_t), st));
  if (nnz == 0) {
    fill_kernel<<<std::max(1, std::min((n_dst + 256) / 256, gr::sm_count() * 8)), 256, 0, st>>>(indptr, (long long)n_dst + 1, 0);
    GR_LAUNCH_CHECK();
    return GR_OK;
  }
  GR_REQUIRE(src && dst && indices && eperm, GR_E_INVALID, "null pointer");
  const Layout l = layout(nnz);
  GR_REQUIRE(ws != nullptr && ws_bytes >= l.total, GR_E_WORKSPACE, "workspace too small");
  GR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, GR_E_INVALID, "workspace must be 256-byte aligned");
  char* base = static_cast<char*>(ws);
  int* keys_a = reinterpret_cast<int*>(base + l.keys_a);
  int* keys_b = reinterpret_cast<int*>(base + l.keys_b);
  int* vals_b = reinterpret_cast<int*>(base + l.vals_b);
  int* hist = reinterpret_cast<int*>(base + l.hist);
  const int n_tiles = (int)((nnz + TILE - 1) / TILE);
  const int grid = std::min(n_tiles, gr::sm_count() * 8);
  const int passes = n_passes(n_dst);
  if (status_or_null != nullptr) {
    csr_validate_kernel<<<grid, THREADS, 0, st>>>(dst, nnz, n_dst, status_or_null);
    GR_LAUNCH_CHECK();
  }
  // ping-pong so that the last pass lands in (keys_?, eperm): vals alternate vals_b <-> eperm
  const int* kin = dst;
  const int* vin = nullptr;
  for (int p = 0; p < passes; ++p) {
    const bool last_to_eperm = ((passes - 1 - p) % 2) == 0;
    int* kout = last_to_eperm ? keys_a : keys_b;
    int* vout = last_to_eperm ? eperm : vals_b;
    const int shift = p * RADIX_BITS;
    radix_hist_kernel<<<grid, THREADS, 0, st>>>(kin, nnz, shift, n_tiles, hist);
    GR_LAUNCH_CHECK();
    scan_kernel<<<1, 1024, 0, st>>>(hist, (long long)BINS * n_tiles);
    GR_LAUNCH_CHECK();
    if (p == 0) radix_scatter_kernel<true><<<grid, THREADS, 0, st>>>(kin, vin, nnz, shift, n_tiles, hist, kout, vout);
    else radix_scatter_kernel<false><<<grid, THREADS, 0, st>>>(kin, vin, nnz, shift, n_tiles, hist, kout, vout);
    GR_LAUNCH_CHECK();
    kin = kout;
    vin = vout;
  }
  const int g2 = (int)std::min<int64_t>((nnz + 255) / 256, (int64_t)gr::sm_count() * 16);
  csr_finish_kernel<<<g2, 256, 0, st>>>(kin, eperm, src, nnz, n_dst, indptr, indices);
  GR_LAUNCH_CHECK();
  return GR_OK;
}
