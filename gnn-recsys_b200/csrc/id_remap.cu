// gr_remap_first_appearance_i64: raw ids -> contiguous ids in order of first appearance, on the device.
//
// Replaces create_ids of the reference (src/builder.py:182-227): `df[col].unique()` (pandas keeps first-appearance
// order) + a merge that assigns 0..n_unique-1 in that order. Bit-exact against oracle.straightline.first_appearance_ids:
//   1. open-addressing hash table keyed by the raw id; the value is the SMALLEST position at which the id occurs
//      (atomicCAS claims the slot, atomicMin keeps the first position) -- order-independent, hence deterministic
//   2. flag[i] = 1 iff position i is the first occurrence of raw[i]; exclusive scan of the flags = new id of that id
//   3. new_ids[i] = scan[first position of raw[i]];  uniq_raw[new id] = raw id (the reverse map the reference pickles)
#include <algorithm>

#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr long long EMPTY = (long long)0x8000000000000000ull;  // empty-slot marker; the raw id INT64_MIN itself is kept out
//                                                                of the table and tracked in one dedicated word (`special`)

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {  // splitmix64 finaliser
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}

__global__ void table_init_kernel(long long* keys, int* vals, long long cap, int* special) {
  if (blockIdx.x == 0 && threadIdx.x == 0) special[0] = 0x7fffffff;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < cap; i += (long long)gridDim.x * blockDim.x) {
    keys[i] = EMPTY;
    vals[i] = 0x7fffffff;
  }
}

__global__ void table_insert_kernel(const long long* __restrict__ raw, long long n, long long* keys, int* vals,
                                    unsigned long long mask, int* special) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long k = raw[i];
    if (k == EMPTY) {  // the one raw id that collides with the empty-slot marker
      atomicMin(special, (int)i);
      continue;
    }
    unsigned long long slot = mix64((unsigned long long)k) & mask;
    while (true) {
      const long long prev = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(keys + slot),
                                                  (unsigned long long)EMPTY, (unsigned long long)k);
      if (prev == EMPTY || prev == k) {
        atomicMin(vals + slot, (int)i);
        break;
      }
      slot = (slot + 1) & mask;
    }
  }
}

__device__ __forceinline__ int first_pos(const long long* __restrict__ keys, const int* __restrict__ vals,
                                         unsigned long long mask, long long k, const int* __restrict__ special) {
  if (k == EMPTY) return special[0];
  unsigned long long slot = mix64((unsigned long long)k) & mask;
  while (keys[slot] != k) slot = (slot + 1) & mask;
  return vals[slot];
}

__global__ void flag_first_kernel(const long long* __restrict__ raw, long long n, const long long* __restrict__ keys,
                                  const int* __restrict__ vals, unsigned long long mask, const int* __restrict__ special,
                                  int* __restrict__ first, int* __restrict__ flag) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int f = first_pos(keys, vals, mask, raw[i], special);
    first[i] = f;
    flag[i] = f == (int)i ? 1 : 0;
  }
}

__global__ void assign_kernel(const long long* __restrict__ raw, long long n, const int* __restrict__ first,
                              const int* __restrict__ rank, int* __restrict__ new_ids, long long* __restrict__ uniq_raw) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int f = first[i];
    const int id = rank[f];
    new_ids[i] = id;
    if (f == (int)i && uniq_raw != nullptr) uniq_raw[id] = raw[i];
  }
}

struct Layout { size_t keys, vals, first, rank, tiles, special, total; long long cap; };
Layout layout(int64_t n) {
  Layout l;
  long long cap = 1024;
  while (cap < 2 * n) cap <<= 1;
  l.cap = cap;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = gr::align_up(off + b, 256); return o; };
  l.keys = take(8 * (size_t)cap);
  l.vals = take(4 * (size_t)cap);
  l.first = take(4 * (size_t)std::max<int64_t>(n, 1));
  l.rank = take(4 * (size_t)std::max<int64_t>(n, 1));
  l.tiles = take(gr::scan_workspace_bytes(std::max<int64_t>(n, 1)));
  l.special = take(256);
  l.total = off;
  return l;
}

}  // namespace

extern "C" size_t gr_remap_workspace_bytes(int64_t n) { return layout(n < 0 ? 0 : n).total; }

extern "C" int gr_remap_first_appearance_i64(const int64_t* raw, int64_t n, int32_t* new_ids,
                                             int64_t* uniq_raw_or_null, int32_t* n_unique, void* ws, size_t ws_bytes,
                                             gr_stream_t stream) {
  GR_REQUIRE(n >= 0 && n <= 0x7fffffffLL, GR_E_INVALID, "n must be in [0, 2^31 - 1]");
  GR_REQUIRE(n_unique != nullptr, GR_E_INVALID, "null n_unique");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) {
    GR_CUDA(cudaMemsetAsync(n_unique, 0, sizeof(int32_t), st));
    return GR_OK;
  }
  GR_REQUIRE(raw && new_ids, GR_E_INVALID, "null pointer");
  const Layout l = layout(n);
  GR_REQUIRE(ws != nullptr && ws_bytes >= l.total, GR_E_WORKSPACE, "workspace too small");
  GR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, GR_E_INVALID, "workspace must be 256-byte aligned");
  char* base = static_cast<char*>(ws);
  long long* keys = reinterpret_cast<long long*>(base + l.keys);
  int* vals = reinterpret_cast<int*>(base + l.vals);
  int* first = reinterpret_cast<int*>(base + l.first);
  int* rank = reinterpret_cast<int*>(base + l.rank);
  int* special = reinterpret_cast<int*>(base + l.special);
  const int grid = gr::sm_count() * 8;
  const unsigned long long mask = (unsigned long long)l.cap - 1;
  table_init_kernel<<<grid, 256, 0, st>>>(keys, vals, l.cap, special);
  GR_LAUNCH_CHECK();
  table_insert_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const long long*>(raw), n, keys, vals, mask, special);
  GR_LAUNCH_CHECK();
  flag_first_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const long long*>(raw), n, keys, vals, mask, special, first, rank);
  GR_LAUNCH_CHECK();
  GR_CUDA(gr::scan_exclusive_i32(rank, n, n_unique, reinterpret_cast<int*>(base + l.tiles), st));
  assign_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const long long*>(raw), n, first, rank, new_ids,
                                      reinterpret_cast<long long*>(uniq_raw_or_null));
  GR_LAUNCH_CHECK();
  return GR_OK;
}
