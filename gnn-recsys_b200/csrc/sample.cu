// Sampled message-passing blocks on the device: the step in front of the training-step
// forward (BASELINE config 4). Replaces what dgl.dataloading does on the CPU behind the reference's loaders
// (src/sampling.py:153-207): MultiLayerNeighborSampler(fanouts, replace=False) / MultiLayerFullNeighborSampler
// frontiers (sample_neighbors / in_subgraph, optionally minus the batch's own edges: exclude='reverse_types') and
// negative_sampler.Uniform(k). Block compaction (to_block) is gr_remap_first_appearance_i64 over
// [seeds | frontier sources] -- see gnn-recsys_b200/sampling_device.py.
//
// Randomness is COUNTER BASED so that a frontier does not depend on thread scheduling and can be restated on the CPU
// bit for bit (oracle.straightline.sample_frontier / negative_uniform):
//     hash64(key, ctr) = fin(key + (ctr + 1) * 0x9E3779B97F4A7C15)        (splitmix64 step + finaliser)
//   * fan-out sampling without replacement: every candidate in-edge e of a seed row gets the 32-bit ticket
//     hash64(key, eid(e)) >> 32; the row keeps the `fanout` candidates with the smallest (ticket, CSR slot) -- a uniformly
//     random subset -- and emits them in CSR order (= edge-id order, the order the aggregation kernel sums in)
//   * negatives: dst[e * k + j] = hash64(key, eid_e * k + j) mod n_dst_nodes, src[e * k + j] = src(eid_e)
//     (k consecutive negatives per positive edge: the layout max_margin_loss reshapes by, src/model.py:516)
//
// One warp per seed row. HBM-bound integer work: coalesced 128-byte reads of the row's eperm / indices slice, the
// running `fanout`-best list lives one entry per lane in registers (fanout <= 32), insertions are warp-uniform.
#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr unsigned long long GOLDEN = 0x9E3779B97F4A7C15ull;

__host__ __device__ __forceinline__ unsigned long long fin64(unsigned long long x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  x ^= x >> 31;
  return x;
}
__host__ __device__ __forceinline__ unsigned long long hash64(unsigned long long key, unsigned long long ctr) {
  return fin64(key + (ctr + 1ull) * GOLDEN);
}

// edge id `eid` in the ascending exclusion list?
__device__ __forceinline__ bool excluded(const int* __restrict__ excl, int n_excl, int eid) {
  int lo = 0, hi = n_excl;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int v = __ldg(excl + mid);
    if (v < eid) lo = mid + 1; else hi = mid;
  }
  return lo < n_excl && __ldg(excl + lo) == eid;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
  const unsigned hi = __reduce_max_sync(gr::FULL, (unsigned)(v >> 32));
  const unsigned lo = __reduce_max_sync(gr::FULL, (unsigned)(v >> 32) == hi ? (unsigned)v : 0u);
  return ((unsigned long long)hi << 32) | lo;
}

// counts[i] = number of in-edges the frontier keeps for seed i (written to out_indptr[i]; out_indptr[n_seeds] = 0,
// the exclusive scan that follows turns the array into the block's indptr with the total in the last slot).
__global__ void sample_count_kernel(const int* __restrict__ indptr, const int* __restrict__ eperm,
                                    const long long* __restrict__ seeds, long long n_seeds, int fanout,
                                    const int* __restrict__ excl, int n_excl, int* __restrict__ out_indptr) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  if (warp0 == 0 && lane == 0) out_indptr[n_seeds] = 0;
  if (n_excl == 0) {  // thread per row: min(deg, fanout)
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_seeds; i += (long long)gridDim.x * blockDim.x) {
      const long long row = seeds[i];
      const int deg = indptr[row + 1] - indptr[row];
      out_indptr[i] = fanout > 0 ? min(deg, fanout) : deg;
    }
    return;
  }
  for (long long i = warp0; i < n_seeds; i += n_warps) {
    const long long row = seeds[i];
    const int b = indptr[row], deg = indptr[row + 1] - b;
    int valid = 0;
    for (int base = 0; base < deg; base += 32) {
      const int slot = base + lane;
      bool ok = false;
      if (slot < deg) {
        const int eid = eperm ? gr::ldg_stream_i32(eperm + b + slot) : b + slot;
        ok = !excluded(excl, n_excl, eid);
      }
      valid += __popc(__ballot_sync(gr::FULL, ok));
    }
    if (lane == 0) out_indptr[i] = fanout > 0 ? min(valid, fanout) : valid;
  }
}

__global__ void __launch_bounds__(256) sample_fill_kernel(
    const int* __restrict__ indptr, const int* __restrict__ indices, const int* __restrict__ eperm,
    const long long* __restrict__ seeds, long long n_seeds, int fanout, const int* __restrict__ excl, int n_excl,
    unsigned long long key, const int* __restrict__ out_indptr, long long* __restrict__ out_src,
    int* __restrict__ out_eid) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp0; i < n_seeds; i += n_warps) {
    const long long row = seeds[i];
    const int b = indptr[row], deg = indptr[row + 1] - b;
    const int o = out_indptr[i], cnt = out_indptr[i + 1] - o;
    if (cnt == 0) continue;
    if (fanout <= 0 || cnt < fanout || deg == cnt) {
      // every (non-excluded) in-edge is kept: ordered compaction of the row slice
      int run = 0;
      for (int base = 0; base < deg; base += 32) {
        const int slot = base + lane;
        bool ok = false;
        int eid = 0;
        if (slot < deg) {
          eid = eperm ? gr::ldg_stream_i32(eperm + b + slot) : b + slot;
          ok = n_excl == 0 || !excluded(excl, n_excl, eid);
        }
        const unsigned m = __ballot_sync(gr::FULL, ok);
        if (ok) {
          const int pos = o + run + __popc(m & ((1u << lane) - 1u));
          out_src[pos] = gr::ldg_stream_i32(indices + b + slot);
          out_eid[pos] = eid;
        }
        run += __popc(m);
      }
      continue;
    }
    // keep the `fanout` smallest (ticket, slot) pairs: lane l < fanout holds one list entry
    unsigned long long mine = lane < fanout ? ~0ull : 0ull;
    unsigned long long cur_max = ~0ull;
    for (int base = 0; base < deg; base += 32) {
      const int slot = base + lane;
      unsigned long long c = ~0ull;
      if (slot < deg) {
        const int eid = eperm ? gr::ldg_stream_i32(eperm + b + slot) : b + slot;
        if (n_excl == 0 || !excluded(excl, n_excl, eid))
          c = (hash64(key, (unsigned long long)(unsigned)eid) & 0xffffffff00000000ull) | (unsigned)slot;
      }
      unsigned m = __ballot_sync(gr::FULL, c < cur_max);
      while (m) {
        const int from = __ffs(m) - 1;
        m &= m - 1;
        const unsigned long long cand = __shfl_sync(gr::FULL, c, from);
        if (cand < cur_max) {  // warp-uniform
          const unsigned holders = __ballot_sync(gr::FULL, lane < fanout && mine == cur_max);
          if (lane == __ffs(holders) - 1) mine = cand;
          cur_max = warp_max_u64(mine);
        }
      }
    }
    const bool sel = lane < fanout && mine != ~0ull;
    const unsigned my_slot = (unsigned)mine;
    int rank = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const unsigned sj = __shfl_sync(gr::FULL, my_slot, j);
      const bool selj = __shfl_sync(gr::FULL, (int)sel, j) != 0;
      rank += (selj && sj < my_slot) ? 1 : 0;
    }
    if (sel) {
      out_src[o + rank] = indices[b + my_slot];
      out_eid[o + rank] = eperm ? eperm[b + my_slot] : b + (int)my_slot;
    }
  }
}

__global__ void negative_uniform_kernel(const int* __restrict__ edge_src, const long long* __restrict__ eids,
                                        long long n_pos, int k, unsigned long long n_dst_nodes, unsigned long long key,
                                        long long* __restrict__ out_src, long long* __restrict__ out_dst) {
  const long long total = n_pos * k;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long e = t / k;
    const int j = (int)(t - e * k);
    const unsigned long long eid = (unsigned long long)eids[e];
    out_src[t] = edge_src[eid];
    out_dst[t] = (long long)(hash64(key, eid * (unsigned long long)k + (unsigned long long)j) % n_dst_nodes);
  }
}

int grid_for(long long work_items, int per_block) {
  const long long want = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)gr::sm_count() * 16;
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

}  // namespace

extern "C" uint64_t gr_sample_key(uint64_t seed, uint64_t stream_id) { return hash64(fin64(seed + GOLDEN), stream_id); }

extern "C" int gr_sample_count_i32(const int32_t* indptr, const int32_t* eperm_or_null, const int64_t* seeds,
                                   int64_t n_seeds, int32_t fanout, const int32_t* excl_sorted_or_null, int32_t n_excl,
                                   int32_t* out_indptr, int32_t* total, gr_stream_t stream) {
  GR_REQUIRE(n_seeds >= 0 && n_seeds < 0x7fffffffLL, GR_E_INVALID, "n_seeds must be in [0, 2^31 - 2]");
  GR_REQUIRE(out_indptr && total, GR_E_INVALID, "null output");
  GR_REQUIRE(fanout <= 32, GR_E_UNSUPPORTED, "fanout > 32 (0 or negative = every in-edge)");
  GR_REQUIRE(n_excl >= 0 && (n_excl == 0 || excl_sorted_or_null), GR_E_INVALID, "bad exclusion list");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_seeds > 0) GR_REQUIRE(indptr && seeds, GR_E_INVALID, "null pointer");
  const int grid = n_excl == 0 ? grid_for(n_seeds, 256) : grid_for(n_seeds, 8);
  sample_count_kernel<<<grid, 256, 0, st>>>(indptr, eperm_or_null, reinterpret_cast<const long long*>(seeds), n_seeds,
                                            fanout, excl_sorted_or_null, n_excl, out_indptr);
  GR_LAUNCH_CHECK();
  gr::scan1_kernel<<<1, 1024, 0, st>>>(out_indptr, n_seeds + 1, total);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

extern "C" int gr_sample_fill_i32(const int32_t* indptr, const int32_t* indices, const int32_t* eperm_or_null,
                                  const int64_t* seeds, int64_t n_seeds, int32_t fanout,
                                  const int32_t* excl_sorted_or_null, int32_t n_excl, uint64_t key,
                                  const int32_t* out_indptr, int64_t* out_src, int32_t* out_eid, gr_stream_t stream) {
  GR_REQUIRE(n_seeds >= 0 && n_seeds < 0x7fffffffLL, GR_E_INVALID, "n_seeds must be in [0, 2^31 - 2]");
  GR_REQUIRE(fanout <= 32, GR_E_UNSUPPORTED, "fanout > 32 (0 or negative = every in-edge)");
  GR_REQUIRE(n_excl >= 0 && (n_excl == 0 || excl_sorted_or_null), GR_E_INVALID, "bad exclusion list");
  if (n_seeds == 0) return GR_OK;
  GR_REQUIRE(indptr && indices && seeds && out_indptr && out_src && out_eid, GR_E_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  sample_fill_kernel<<<grid_for(n_seeds, 8), 256, 0, st>>>(indptr, indices, eperm_or_null,
                                                           reinterpret_cast<const long long*>(seeds), n_seeds, fanout,
                                                           excl_sorted_or_null, n_excl, key, out_indptr,
                                                           reinterpret_cast<long long*>(out_src), out_eid);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

extern "C" int gr_negative_uniform_i64(const int32_t* edge_src, const int64_t* eids, int64_t n_pos, int32_t k,
                                       int64_t n_dst_nodes, uint64_t key, int64_t* out_src, int64_t* out_dst,
                                       gr_stream_t stream) {
  GR_REQUIRE(n_pos >= 0 && k >= 0, GR_E_INVALID, "negative size");
  if (n_pos == 0 || k == 0) return GR_OK;
  GR_REQUIRE(n_dst_nodes > 0, GR_E_INVALID, "no destination node to draw from");
  GR_REQUIRE(edge_src && eids && out_src && out_dst, GR_E_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  negative_uniform_kernel<<<grid_for(n_pos * k, 1024), 256, 0, st>>>(
      edge_src, reinterpret_cast<const long long*>(eids), n_pos, k, (unsigned long long)n_dst_nodes, key,
      reinterpret_cast<long long*>(out_src), reinterpret_cast<long long*>(out_dst));
  GR_LAUNCH_CHECK();
  return GR_OK;
}
