// gr_sage_relation_f32 / gr_gather_reduce_f32: ConvLayer.forward for one relation, fused.
//
// Replaces (reference src/model.py:143-162, 226-235 + dgl.nn.HeteroGraphConv): DGL update_all (libdgl SpMM
// copy_u/u_mul_e with mean/max reducer), torch fc_self / fc_neigh sgemm, relu, row L2 norm with where(norm==0,1),
// and the per-destination-type stack+sum/mean/max across relations.
//
// Layout: int32 CSR over destination rows; fp32 row-major feature tables. One warp gathers one destination row:
// 32 lanes x float4 = 512 B per neighbour row (128 columns; two float4 per lane for 256 columns), 8 independent
// 128-bit loads in flight per lane. A CTA owns a tile of R destination rows: gathered means / maxima and the self
// rows are staged in shared memory, then the tile is pushed through z = relu(S.Ws^T + N.Wn^T) with a
// register-blocked FFMA micro-kernel (weights k-major, read through L1), L2-normalised per row and combined into
// `out` (store / add / max) -- h_neigh never makes a round trip through HBM.
//
// Hub rows (> GR_SAGE_LONG_ROW in-edges) are cut into GR_SAGE_CHUNK-edge chunks that are reduced by separate CTAs
// and summed per row in chunk order, so the result is deterministic and no CTA serialises a 10^6-edge row.
//
// Sum order: neighbours are accumulated in CSR (edge-id) order like the sequential CPU loop of the oracle; mean
// divides by the degree (IEEE division) exactly like `sum / clamp(deg, 1)`.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace {

using gr::FULL;
constexpr int LONG_ROW = GR_SAGE_LONG_ROW;
constexpr int CHUNK = GR_SAGE_CHUNK;
constexpr int THREADS = 256;
constexpr int WARPS = THREADS / 32;

struct LongWs {
  int* counters;    // [0] = number of long rows, [1] = number of chunks
  int* long_rows;   // row id of each long row
  int* long_base;   // first chunk slot of each long row
  int* chunk_long;  // long-row index of each chunk
  float* partials;  // [max_chunks][d]
  float* long_agg;  // [max_long][d] reduced neighbour row of each long row
  int max_long, max_chunks;
};

size_t long_ws_layout(int64_t nnz, int d, LongWs* w, char* base) {
  const int64_t max_long = nnz / LONG_ROW + 1;
  const int64_t max_chunks = nnz / CHUNK + max_long + 1;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = gr::align_up(off + bytes, 256);
    return o;
  };
  size_t o_cnt = take(256), o_rows = take(4 * max_long), o_base = take(4 * max_long), o_cl = take(4 * max_chunks);
  size_t o_part = take(sizeof(float) * d * max_chunks), o_agg = take(sizeof(float) * d * max_long);
  if (w) {
    w->counters = reinterpret_cast<int*>(base + o_cnt);
    w->long_rows = reinterpret_cast<int*>(base + o_rows);
    w->long_base = reinterpret_cast<int*>(base + o_base);
    w->chunk_long = reinterpret_cast<int*>(base + o_cl);
    w->partials = reinterpret_cast<float*>(base + o_part);
    w->long_agg = reinterpret_cast<float*>(base + o_agg);
    w->max_long = (int)max_long;
    w->max_chunks = (int)max_chunks;
  }
  return off;
}

// ------------------------------------------------------------------------------------------------ gather
template <bool MAXR>
__device__ __forceinline__ float4 ident4() {
  const float v = MAXR ? -INFINITY : 0.f;
  return make_float4(v, v, v, v);
}

template <bool MAXR>
__device__ __forceinline__ void combine(float4& a, const float4& v) {
  if (MAXR) {
    a.x = fmaxf(a.x, v.x); a.y = fmaxf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w);
  } else {
    a.x = __fadd_rn(a.x, v.x); a.y = __fadd_rn(a.y, v.y); a.z = __fadd_rn(a.z, v.z); a.w = __fadd_rn(a.w, v.w);
  }
}

// Warp-cooperative reduce of neighbour rows for CSR slots [e0, e1) into acc (pre-initialised by the caller).
// Lane l owns columns 4*(l + 32q) .. +3 for q < VN.
template <int VN, bool MAXR>
__device__ __forceinline__ void gather_range(const int* __restrict__ indices, const float* __restrict__ ew,
                                             const float* __restrict__ h, int d, int e0, int e1, int lane,
                                             float4 (&acc)[VN]) {
  constexpr int UNROLL = (VN == 1) ? 8 : 4;
  for (int e = e0; e < e1; e += 32) {
    const int cnt = min(32, e1 - e);
    const int my = lane < cnt ? gr::ldg_stream_i32(indices + e + lane) : 0;
    float myw = 1.f;
    if (ew != nullptr) myw = lane < cnt ? gr::ldg_stream_f32(ew + e + lane) : 0.f;
    for (int j = 0; j < cnt; j += UNROLL) {
      float4 v[UNROLL][VN];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int s = __shfl_sync(FULL, my, (j + u) & 31);
        const float* row = h + (size_t)s * d;
#pragma unroll
        for (int q = 0; q < VN; ++q) {
          const int c = (lane + 32 * q) * 4;
          v[u][q] = (j + u < cnt && c < d) ? gr::ldg_f4(row + c) : ident4<MAXR>();
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (ew != nullptr) {
          const float w = __shfl_sync(FULL, myw, (j + u) & 31);
          if (j + u < cnt) {
#pragma unroll
            for (int q = 0; q < VN; ++q) {
              v[u][q].x = __fmul_rn(v[u][q].x, w); v[u][q].y = __fmul_rn(v[u][q].y, w);
              v[u][q].z = __fmul_rn(v[u][q].z, w); v[u][q].w = __fmul_rn(v[u][q].w, w);
            }
          }
        }
#pragma unroll
        for (int q = 0; q < VN; ++q) combine<MAXR>(acc[q], v[u][q]);
      }
    }
  }
}

template <bool MAXR>
__device__ __forceinline__ float4 finalize4(float4 a, int deg) {
  if (deg == 0) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (!MAXR) {
    const float f = (float)deg;
    a.x = a.x / f; a.y = a.y / f; a.z = a.z / f; a.w = a.w / f;
  }
  return a;
}

// Warp-level lookup of `row` in the (short, unsorted) long-row list.
__device__ __forceinline__ int find_long(const int* __restrict__ long_rows, int n_long, int row, int lane) {
  int found = -1;
  for (int i = lane; i < n_long; i += 32)
    if (long_rows[i] == row) found = i;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) found = max(found, __shfl_xor_sync(FULL, found, o));
  return found;
}

// Reduced neighbour row of `row` -> dst (shared or global, 16-byte aligned rows of d floats).
template <int VN, bool MAXR>
__device__ __forceinline__ void reduce_row_to(const int* __restrict__ indptr, const int* __restrict__ indices,
                                              const float* __restrict__ ew, const float* __restrict__ h, int d,
                                              int row, int lane, const LongWs& lw, float* dst) {
  const int beg = __ldg(indptr + row), end = __ldg(indptr + row + 1);
  const int deg = end - beg;
  float4 acc[VN];
  if (deg > LONG_ROW) {
    const int idx = find_long(lw.long_rows, lw.counters[0], row, lane);
#pragma unroll
    for (int q = 0; q < VN; ++q) {
      const int c = (lane + 32 * q) * 4;
      if (c < d) acc[q] = *reinterpret_cast<const float4*>(lw.long_agg + (size_t)idx * d + c);
    }
  } else {
#pragma unroll
    for (int q = 0; q < VN; ++q) acc[q] = ident4<MAXR>();
    gather_range<VN, MAXR>(indices, ew, h, d, beg, end, lane, acc);
#pragma unroll
    for (int q = 0; q < VN; ++q) acc[q] = finalize4<MAXR>(acc[q], deg);
  }
#pragma unroll
  for (int q = 0; q < VN; ++q) {
    const int c = (lane + 32 * q) * 4;
    if (c < d) *reinterpret_cast<float4*>(dst + c) = acc[q];
  }
}

// ------------------------------------------------------------------------------------------------ hub rows
__global__ void collect_long_rows_kernel(const int* __restrict__ indptr, int64_t row_begin, int64_t row_end,
                                         LongWs lw) {
  for (int64_t r = row_begin + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < row_end;
       r += (int64_t)gridDim.x * blockDim.x) {
    const int deg = indptr[r + 1] - indptr[r];
    if (deg > LONG_ROW) {
      const int nchunk = (deg + CHUNK - 1) / CHUNK;
      const int idx = atomicAdd(lw.counters + 0, 1);
      const int base = atomicAdd(lw.counters + 1, nchunk);
      lw.long_rows[idx] = (int)r;
      lw.long_base[idx] = base;
      for (int c = 0; c < nchunk; ++c) lw.chunk_long[base + c] = idx;
    }
  }
}

// One CTA per chunk: 8 warps x (CHUNK/8) edges each, partial rows combined in warp order.
template <int VN, bool MAXR>
__global__ void __launch_bounds__(THREADS) long_partial_kernel(const int* __restrict__ indptr,
                                                               const int* __restrict__ indices,
                                                               const float* __restrict__ ew,
                                                               const float* __restrict__ h, int d, LongWs lw) {
  extern __shared__ __align__(16) float smem[];  // [WARPS][d]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = lw.counters[1];
  for (int ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int idx = lw.chunk_long[ch];
    const int row = lw.long_rows[idx];
    const int local = ch - lw.long_base[idx];
    const int beg = indptr[row] + local * CHUNK;
    const int end = min(indptr[row + 1], beg + CHUNK);
    constexpr int PER_WARP = CHUNK / WARPS;
    const int e0 = min(end, beg + warp * PER_WARP), e1 = min(end, e0 + PER_WARP);
    float4 acc[VN];
#pragma unroll
    for (int q = 0; q < VN; ++q) acc[q] = ident4<MAXR>();
    gather_range<VN, MAXR>(indices, ew, h, d, e0, e1, lane, acc);
#pragma unroll
    for (int q = 0; q < VN; ++q) {
      const int c = (lane + 32 * q) * 4;
      if (c < d) *reinterpret_cast<float4*>(smem + warp * d + c) = acc[q];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += THREADS) {
      float a = smem[c];
      for (int w = 1; w < WARPS; ++w) a = MAXR ? fmaxf(a, smem[w * d + c]) : __fadd_rn(a, smem[w * d + c]);
      lw.partials[(size_t)ch * d + c] = a;
    }
    __syncthreads();
  }
}

template <bool MAXR>
__global__ void long_reduce_kernel(const int* __restrict__ indptr, int d, LongWs lw) {
  const int n_long = lw.counters[0];
  for (int idx = blockIdx.x; idx < n_long; idx += gridDim.x) {
    const int row = lw.long_rows[idx];
    const int deg = indptr[row + 1] - indptr[row];
    const int nchunk = (deg + CHUNK - 1) / CHUNK;
    const float* p = lw.partials + (size_t)lw.long_base[idx] * d;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      float a = p[c];
      for (int k = 1; k < nchunk; ++k) a = MAXR ? fmaxf(a, p[(size_t)k * d + c]) : __fadd_rn(a, p[(size_t)k * d + c]);
      lw.long_agg[(size_t)idx * d + c] = MAXR ? a : a / (float)deg;
    }
  }
}

// ------------------------------------------------------------------------------------------------ gather only
template <int VN, bool MAXR>
__global__ void __launch_bounds__(THREADS) gather_only_kernel(const int* __restrict__ indptr,
                                                              const int* __restrict__ indices,
                                                              const float* __restrict__ ew,
                                                              const float* __restrict__ h, int d, int64_t row_begin,
                                                              int64_t row_end, LongWs lw, float* __restrict__ agg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t r = row_begin + (int64_t)blockIdx.x * WARPS + warp; r < row_end; r += (int64_t)gridDim.x * WARPS)
    reduce_row_to<VN, MAXR>(indptr, indices, ew, h, d, (int)r, lane, lw, agg + (size_t)r * d);
}

// ------------------------------------------------------------------------------------------------ fused tile
template <int CPL>
__device__ __forceinline__ void load_cols(const float* __restrict__ p, float (&w)[CPL]) {
  if (CPL == 2) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p));
    w[0] = v.x; w[1] = v.y;
  } else {
#pragma unroll
    for (int i = 0; i < CPL / 4; ++i) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
      w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
    }
  }
}

// acc[r][c] += sum_k tile[r][k] * wt[k][col0 + c] for the warp's RPW rows; tile rows have pitch dk (multiple of 4).
template <int RPW, int CPL>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ tile, int dk, const float* __restrict__ wt,
                                          int d_out, int col0, float (&acc)[RPW][CPL]) {
  for (int k = 0; k < dk; k += 4) {
    float4 a[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) a[r] = *reinterpret_cast<const float4*>(tile + r * dk + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      float w[CPL];
      load_cols<CPL>(wt + (size_t)(k + kk) * d_out + col0, w);
#pragma unroll
      for (int r = 0; r < RPW; ++r) {
        const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
#pragma unroll
        for (int c = 0; c < CPL; ++c) acc[r][c] = fmaf(av, w[c], acc[r][c]);
      }
    }
  }
}

struct SageParams {
  const int* indptr; const int* indices; const float* ew;
  const float* h_src; const float* h_dst;
  int64_t row_begin, row_end;
  int dn, ds, dout;
  const float* ws_t; const float* wn_t;
  int l2norm, accumulate;
  float z_scale;
  float* out;
};

// R = rows per tile, CPL = d_out / 32 output columns per lane, VN / VS = float4 per lane of a neighbour / self row.
template <int VN, int VS, int CPL, int R, bool MAXR>
__global__ void __launch_bounds__(THREADS) sage_fused_kernel(SageParams p, LongWs lw) {
  extern __shared__ __align__(16) float smem[];
  float* sN = smem;                // [R][dn]
  float* sS = smem + R * p.dn;     // [R][ds]
  __shared__ int s_next;
  constexpr int RPW = R / WARPS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row0 = p.row_begin + (int64_t)blockIdx.x * R;
  if (threadIdx.x == 0) s_next = 0;
  __syncthreads();

  // phase 1: warps pull rows of the tile dynamically (degree skew), gather -> smem
  while (true) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&s_next, 1);
    r = __shfl_sync(FULL, r, 0);
    if (r >= R) break;
    const int64_t row = row0 + r;
    if (row < p.row_end) {
      reduce_row_to<VN, MAXR>(p.indptr, p.indices, p.ew, p.h_src, p.dn, (int)row, lane, lw, sN + r * p.dn);
#pragma unroll
      for (int q = 0; q < VS; ++q) {
        const int c = (lane + 32 * q) * 4;
        if (c < p.ds) *reinterpret_cast<float4*>(sS + r * p.ds + c) = gr::ldg_f4(p.h_dst + (size_t)row * p.ds + c);
      }
    } else {
      for (int c = lane * 4; c < p.dn; c += 128) *reinterpret_cast<float4*>(sN + r * p.dn + c) = make_float4(0, 0, 0, 0);
      for (int c = lane * 4; c < p.ds; c += 128) *reinterpret_cast<float4*>(sS + r * p.ds + c) = make_float4(0, 0, 0, 0);
    }
  }
  __syncthreads();

  // phase 2: z = relu(S.Ws^T + N.Wn^T); warp w owns rows [w*RPW, +RPW), lane owns columns [lane*CPL, +CPL)
  float acc[RPW][CPL];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int c = 0; c < CPL; ++c) acc[r][c] = 0.f;
  const int col0 = lane * CPL;
  tile_gemm<RPW, CPL>(sS + warp * RPW * p.ds, p.ds, p.ws_t, p.dout, col0, acc);
  tile_gemm<RPW, CPL>(sN + warp * RPW * p.dn, p.dn, p.wn_t, p.dout, col0, acc);

#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int64_t row = row0 + warp * RPW + r;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < CPL; ++c) {
      acc[r][c] = fmaxf(acc[r][c], 0.f);
      ss = fmaf(acc[r][c], acc[r][c], ss);
    }
    if (p.l2norm) {
      ss = gr::warp_sum(ss);
      float nrm = sqrtf(ss);
      if (nrm == 0.f) nrm = 1.f;
#pragma unroll
      for (int c = 0; c < CPL; ++c) acc[r][c] = acc[r][c] / nrm;
    }
    if (row < p.row_end) {
      float* o = p.out + (size_t)row * p.dout + col0;
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        float z = acc[r][c];
        if (p.accumulate == GR_ACC_ADD) z = o[c] + z;
        else if (p.accumulate == GR_ACC_MAX) z = fmaxf(o[c], z);
        acc[r][c] = z * p.z_scale;
      }
      if (CPL == 2) {
        *reinterpret_cast<float2*>(o) = make_float2(acc[r][0], acc[r][1]);
      } else {
#pragma unroll
        for (int i = 0; i < CPL / 4; ++i)
          reinterpret_cast<float4*>(o)[i] = make_float4(acc[r][4 * i], acc[r][4 * i + 1], acc[r][4 * i + 2], acc[r][4 * i + 3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ generic dims
// Any d_neigh / d_self / d_out <= 256 (e.g. raw 2/4-column features when embedding_layer=False): one warp per row,
// scalar column loops, rows staged in shared memory. Correctness path for odd shapes, not a tuned kernel.
template <bool MAXR>
__global__ void __launch_bounds__(128) sage_generic_kernel(SageParams p, LongWs lw) {
  extern __shared__ __align__(16) float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sRow = smem + warp * (p.dn + p.ds);  // [dn] reduced neighbours, then [ds] self
  for (int64_t row = p.row_begin + (int64_t)blockIdx.x * 4 + warp; row < p.row_end; row += (int64_t)gridDim.x * 4) {
    const int beg = p.indptr[row], end = p.indptr[row + 1];
    const int deg = end - beg;
    for (int c0 = 0; c0 < p.dn; c0 += 32) {
      const int c = c0 + lane;
      float a = MAXR ? -INFINITY : 0.f;
      if (c < p.dn) {
        for (int e = beg; e < end; ++e) {
          float v = __ldg(p.h_src + (size_t)p.indices[e] * p.dn + c);
          if (p.ew != nullptr) v = __fmul_rn(v, p.ew[e]);
          a = MAXR ? fmaxf(a, v) : __fadd_rn(a, v);
        }
        if (deg == 0) a = 0.f;
        else if (!MAXR) a = a / (float)deg;
        sRow[c] = a;
      }
    }
    for (int c = lane; c < p.ds; c += 32) sRow[p.dn + c] = __ldg(p.h_dst + (size_t)row * p.ds + c);
    __syncwarp();
    float z[8];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int o = lane + 32 * q;
      z[q] = 0.f;
      if (o < p.dout) {
        float a = 0.f;
        for (int k = 0; k < p.ds; ++k) a = fmaf(sRow[p.dn + k], __ldg(p.ws_t + (size_t)k * p.dout + o), a);
        for (int k = 0; k < p.dn; ++k) a = fmaf(sRow[k], __ldg(p.wn_t + (size_t)k * p.dout + o), a);
        z[q] = fmaxf(a, 0.f);
        ss = fmaf(z[q], z[q], ss);
      }
    }
    float nrm = 1.f;
    if (p.l2norm) {
      nrm = sqrtf(gr::warp_sum(ss));
      if (nrm == 0.f) nrm = 1.f;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int o = lane + 32 * q;
      if (o < p.dout) {
        float v = p.l2norm ? z[q] / nrm : z[q];
        float* dst = p.out + (size_t)row * p.dout + o;
        if (p.accumulate == GR_ACC_ADD) v = *dst + v;
        else if (p.accumulate == GR_ACC_MAX) v = fmaxf(*dst, v);
        *dst = v * p.z_scale;
      }
    }
    __syncwarp();
  }
}

// generic-dims gather only (d not a multiple of 4 or > 256)
template <bool MAXR>
__global__ void gather_generic_kernel(const int* __restrict__ indptr, const int* __restrict__ indices,
                                      const float* __restrict__ ew, const float* __restrict__ h, int d,
                                      int64_t row_begin, int64_t row_end, float* __restrict__ agg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t row = row_begin + (int64_t)blockIdx.x * wpb + warp; row < row_end; row += (int64_t)gridDim.x * wpb) {
    const int beg = indptr[row], end = indptr[row + 1];
    for (int c = lane; c < d; c += 32) {
      float a = MAXR ? -INFINITY : 0.f;
      for (int e = beg; e < end; ++e) {
        float v = __ldg(h + (size_t)indices[e] * d + c);
        if (ew != nullptr) v = __fmul_rn(v, ew[e]);
        a = MAXR ? fmaxf(a, v) : __fadd_rn(a, v);
      }
      if (end == beg) a = 0.f;
      else if (!MAXR) a = a / (float)(end - beg);
      agg[(size_t)row * d + c] = a;
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
template <int VN, bool MAXR>
int launch_long_rows(const int* indptr, const int* indices, const float* ew, const float* h, int d, int64_t row_begin,
                     int64_t row_end, const LongWs& lw, cudaStream_t st) {
  GR_CUDA(cudaMemsetAsync(lw.counters, 0, 256, st));
  const int64_t rows = row_end - row_begin;
  const int g1 = (int)std::min<int64_t>((rows + 255) / 256, (int64_t)gr::sm_count() * 8);
  collect_long_rows_kernel<<<std::max(g1, 1), 256, 0, st>>>(indptr, row_begin, row_end, lw);
  GR_LAUNCH_CHECK();
  const int g2 = std::min(lw.max_chunks, gr::sm_count() * 4);
  long_partial_kernel<VN, MAXR><<<g2, THREADS, sizeof(float) * WARPS * d, st>>>(indptr, indices, ew, h, d, lw);
  GR_LAUNCH_CHECK();
  const int g3 = std::min(lw.max_long, gr::sm_count() * 4);
  long_reduce_kernel<MAXR><<<g3, 128, 0, st>>>(indptr, d, lw);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

template <int VN, int VS, int CPL, int R, bool MAXR>
int launch_fused(const SageParams& p, const LongWs& lw, cudaStream_t st) {
  const size_t smem = sizeof(float) * R * (p.dn + p.ds);
  auto kern = sage_fused_kernel<VN, VS, CPL, R, MAXR>;
  GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t rows = p.row_end - p.row_begin;
  const int64_t tiles = (rows + R - 1) / R;
  kern<<<(unsigned)tiles, THREADS, smem, st>>>(p, lw);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

template <bool MAXR>
int dispatch_fused(const SageParams& p, const LongWs& lw, cudaStream_t st, bool* handled) {
  *handled = true;
  const int vn = (p.dn + 127) / 128, vs = (p.ds + 127) / 128;
#define GR_CASE(VN_, VS_, CPL_, R_) \
  if (vn == VN_ && vs == VS_ && p.dout == 32 * CPL_) return launch_fused<VN_, VS_, CPL_, R_, MAXR>(p, lw, st);
  GR_CASE(1, 1, 4, 64)  // 128 -> 128 (c1, c2, c5)
  GR_CASE(1, 1, 2, 64)  // .. -> 64
  GR_CASE(2, 2, 8, 32)  // 256 -> 256 (c3 hidden)
  GR_CASE(2, 2, 4, 32)  // 256 -> 128 (c3 output)
  GR_CASE(1, 1, 8, 32)  // 128 -> 256
  GR_CASE(2, 2, 2, 32)
#undef GR_CASE
  *handled = false;
  return GR_OK;
}

bool fast_dims(int dn, int ds, int dout) {
  return dn % 4 == 0 && ds % 4 == 0 && dn <= 256 && ds <= 256 && ((dn + 127) / 128 == (ds + 127) / 128) &&
         (dout == 64 || dout == 128 || dout == 256);
}

}  // namespace

extern "C" size_t gr_sage_relation_workspace_bytes(int64_t nnz, int32_t d_neigh) {
  return long_ws_layout(nnz < 0 ? 0 : nnz, d_neigh, nullptr, nullptr);
}

extern "C" int gr_sage_relation_f32(const int32_t* indptr, const int32_t* indices, const float* edge_w_or_null,
                                    int64_t nnz, const float* h_src, const float* h_dst, int64_t row_begin,
                                    int64_t row_end, int32_t d_neigh, int32_t d_self, const float* w_self_t,
                                    const float* w_neigh_t, int32_t d_out, int reducer, int l2norm, int accumulate,
                                    float z_scale, float* out, void* ws, size_t ws_bytes, gr_stream_t stream) {
  GR_REQUIRE(row_begin >= 0 && row_end >= row_begin, GR_E_INVALID, "bad row range");
  GR_REQUIRE(d_neigh > 0 && d_self > 0 && d_out > 0 && d_neigh <= 256 && d_self <= 256 && d_out <= 256, GR_E_INVALID,
             "dimensions must be in [1, 256]");
  GR_REQUIRE(reducer == GR_REDUCE_MEAN || reducer == GR_REDUCE_MAX, GR_E_INVALID, "unknown reducer");
  GR_REQUIRE(accumulate >= GR_ACC_STORE && accumulate <= GR_ACC_MAX, GR_E_INVALID, "unknown accumulate mode");
  if (row_end == row_begin) return GR_OK;
  GR_REQUIRE(indptr && h_src && h_dst && w_self_t && w_neigh_t && out, GR_E_INVALID, "null pointer");
  GR_REQUIRE(nnz == 0 || indices, GR_E_INVALID, "null indices");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SageParams p{indptr, indices, edge_w_or_null, h_src, h_dst, row_begin, row_end, d_neigh, d_self, d_out,
               w_self_t, w_neigh_t, l2norm, accumulate, z_scale, out};
  const bool maxr = reducer == GR_REDUCE_MAX;
  LongWs lw{};
  if (fast_dims(d_neigh, d_self, d_out)) {
    const size_t need = long_ws_layout(nnz, d_neigh, nullptr, nullptr);
    GR_REQUIRE(ws != nullptr && ws_bytes >= need, GR_E_WORKSPACE, "workspace too small");
    GR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, GR_E_INVALID, "workspace must be 256-byte aligned");
    long_ws_layout(nnz, d_neigh, &lw, static_cast<char*>(ws));
    int rc;
    const int vn = (d_neigh + 127) / 128;
    if (maxr) rc = vn == 1 ? launch_long_rows<1, true>(indptr, indices, edge_w_or_null, h_src, d_neigh, row_begin, row_end, lw, st)
                           : launch_long_rows<2, true>(indptr, indices, edge_w_or_null, h_src, d_neigh, row_begin, row_end, lw, st);
    else rc = vn == 1 ? launch_long_rows<1, false>(indptr, indices, edge_w_or_null, h_src, d_neigh, row_begin, row_end, lw, st)
                      : launch_long_rows<2, false>(indptr, indices, edge_w_or_null, h_src, d_neigh, row_begin, row_end, lw, st);
    if (rc != GR_OK) return rc;
    bool handled = false;
    rc = maxr ? dispatch_fused<true>(p, lw, st, &handled) : dispatch_fused<false>(p, lw, st, &handled);
    if (rc != GR_OK) return rc;
    if (handled) return GR_OK;
  }
  // generic shapes: scalar kernel (hub rows are walked by a single warp there; only small / odd-shaped inputs land here)
  const int64_t rows = row_end - row_begin;
  const int grid = (int)std::min<int64_t>((rows + 3) / 4, (int64_t)gr::sm_count() * 16);
  const size_t smem = sizeof(float) * 4 * (d_neigh + d_self);
  if (maxr) sage_generic_kernel<true><<<grid, 128, smem, st>>>(p, lw);
  else sage_generic_kernel<false><<<grid, 128, smem, st>>>(p, lw);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

extern "C" int gr_gather_reduce_f32(const int32_t* indptr, const int32_t* indices, const float* edge_w_or_null,
                                    int64_t nnz, const float* h_src, int64_t row_begin, int64_t row_end, int32_t d,
                                    int reducer, float* agg, void* ws, size_t ws_bytes, gr_stream_t stream) {
  GR_REQUIRE(row_begin >= 0 && row_end >= row_begin && d > 0, GR_E_INVALID, "bad arguments");
  GR_REQUIRE(reducer == GR_REDUCE_MEAN || reducer == GR_REDUCE_MAX, GR_E_INVALID, "unknown reducer");
  if (row_end == row_begin) return GR_OK;
  GR_REQUIRE(indptr && h_src && agg, GR_E_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool maxr = reducer == GR_REDUCE_MAX;
  const int64_t rows = row_end - row_begin;
  if (d % 4 == 0 && d <= 256) {
    const size_t need = long_ws_layout(nnz, d, nullptr, nullptr);
    GR_REQUIRE(ws != nullptr && ws_bytes >= need, GR_E_WORKSPACE, "workspace too small");
    GR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, GR_E_INVALID, "workspace must be 256-byte aligned");
    LongWs lw{};
    long_ws_layout(nnz, d, &lw, static_cast<char*>(ws));
    const int grid = (int)std::min<int64_t>((rows + WARPS - 1) / WARPS, (int64_t)gr::sm_count() * 32);
    int rc;
#define GR_GO(VN_, MAX_)                                                                                             \
  rc = launch_long_rows<VN_, MAX_>(indptr, indices, edge_w_or_null, h_src, d, row_begin, row_end, lw, st);           \
  if (rc != GR_OK) return rc;                                                                                        \
  gather_only_kernel<VN_, MAX_><<<grid, THREADS, 0, st>>>(indptr, indices, edge_w_or_null, h_src, d, row_begin, row_end, lw, agg);
    if (d <= 128) { if (maxr) { GR_GO(1, true) } else { GR_GO(1, false) } }
    else { if (maxr) { GR_GO(2, true) } else { GR_GO(2, false) } }
#undef GR_GO
  } else {
    const int grid = (int)std::min<int64_t>((rows + 3) / 4, (int64_t)gr::sm_count() * 16);
    if (maxr) gather_generic_kernel<true><<<grid, 128, 0, st>>>(indptr, indices, edge_w_or_null, h_src, d, row_begin, row_end, agg);
    else gather_generic_kernel<false><<<grid, 128, 0, st>>>(indptr, indices, edge_w_or_null, h_src, d, row_begin, row_end, agg);
  }
  GR_LAUNCH_CHECK();
  return GR_OK;
}
