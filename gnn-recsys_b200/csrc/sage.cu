// gr_sage_relation_f32 / gr_gather_reduce_f32: ConvLayer.forward for one relation, fused.
//
// Replaces (reference src/model.py:143-162, 226-235 + dgl.nn.HeteroGraphConv): DGL update_all (libdgl SpMM
// copy_u/u_mul_e with mean/max reducer), torch fc_self / fc_neigh sgemm, relu, row L2 norm with where(norm==0,1),
// and the per-destination-type stack+sum/mean/max across relations.
//
// Layout: int32 CSR over destination rows; fp32 row-major feature tables. One warp gathers one destination row:
// 32 lanes x float4 = 512 B per neighbour row (128 columns; two float4 per lane for 256 columns), neighbours taken in
// groups of exactly 8 / 4 / 2 / 1 independent 128-bit loads, summed with packed add.rn.f32x2. A group of four warps owns
// a tile of R destination rows: the reduced neighbour rows and the self rows are staged in shared memory (fp16 hi / lo
// planes, split once by the gathering warp), then the tile is pushed through z = relu(S.Ws^T + N.Wn^T) on the tensor
// cores (3-product hi/lo split mma.sync, fp32 accurate; A fragments by ldmatrix, pre-packed B fragments from L2),
// L2-normalised per row and combined into `out` (store / add / max) -- h_neigh never makes a round trip through HBM.
//
// Hub rows (> GR_SAGE_LONG_ROW in-edges) are cut into GR_SAGE_CHUNK-edge chunks that are reduced by separate CTAs
// and summed per row in chunk order, so the result is deterministic and no CTA serialises a 10^6-edge row.
//
// Sum order: neighbours are accumulated in CSR (edge-id) order like the sequential CPU loop of the oracle; mean
// divides by the degree (IEEE division) exactly like `sum / clamp(deg, 1)`.
#include <cuda_fp16.h>

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

// sm_100 packed fp32 pairs (add.rn.f32x2 / mul.rn.f32x2): two IEEE round-to-nearest results per instruction --
// bit-identical to the scalar forms, half the issue slots of the gather loop.
__device__ __forceinline__ void add2_rn(float& a0, float& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\t"
      "mov.b64 ra, {%0, %1};\n\t"
      "mov.b64 rb, {%2, %3};\n\t"
      "add.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void mul2_rn(float& a0, float& a1, float w) {
  asm("{\n\t.reg .b64 ra, rb;\n\t"
      "mov.b64 ra, {%0, %1};\n\t"
      "mov.b64 rb, {%2, %2};\n\t"
      "mul.rn.f32x2 ra, ra, rb;\n\t"
      "mov.b64 {%0, %1}, ra;\n\t}"
      : "+f"(a0), "+f"(a1) : "f"(w));
}

template <bool MAXR>
__device__ __forceinline__ void combine(float4& a, const float4& v) {
  if (MAXR) {
    a.x = fmaxf(a.x, v.x); a.y = fmaxf(a.y, v.y); a.z = fmaxf(a.z, v.z); a.w = fmaxf(a.w, v.w);
  } else {
    add2_rn(a.x, a.y, v.x, v.y);
    add2_rn(a.z, a.w, v.z, v.w);
  }
}

// Warp-cooperative reduce of neighbour rows for CSR slots [e0, e1) into acc (pre-initialised by the caller).
// Lane l owns columns 4*(l + 32q) .. +3 for q < VN; lanes whose columns lie beyond d read column block 0 instead (no
// predicate in the loop) and their accumulators are meaningless -- callers clear them (clear_tail_lanes).
// Neighbours are consumed in groups of exactly N in {8, 4, 2, 1} (warp-uniform branches, N independent 128-bit loads in
// flight, no padded slots), in CSR order: the sum order is the sequential order of the oracle.
// `hb[q]` = byte address of this lane's column block q in row 0; a neighbour row adds id * row_bytes (one IMAD.WIDE.U32).
template <int N, int VN, bool MAXR>
__device__ __forceinline__ void gather_group(const char* const (&hb)[VN], unsigned row_bytes, int my, float myw,
                                             bool weighted, int j, float4 (&acc)[VN]) {
  float4 v[N][VN];
#pragma unroll
  for (int u = 0; u < N; ++u) {
    const unsigned s = (unsigned)__shfl_sync(FULL, my, j + u);
    const size_t off = (size_t)s * row_bytes;
#pragma unroll
    for (int q = 0; q < VN; ++q) v[u][q] = __ldg(reinterpret_cast<const float4*>(hb[q] + off));
  }
#pragma unroll
  for (int u = 0; u < N; ++u) {
    if (weighted) {
      const float w = __shfl_sync(FULL, myw, j + u);
#pragma unroll
      for (int q = 0; q < VN; ++q) {
        mul2_rn(v[u][q].x, v[u][q].y, w);
        mul2_rn(v[u][q].z, v[u][q].w, w);
      }
    }
#pragma unroll
    for (int q = 0; q < VN; ++q) combine<MAXR>(acc[q], v[u][q]);
  }
}

template <int VN, bool MAXR>
__device__ __forceinline__ void gather_range(const int* __restrict__ indices, const float* __restrict__ ew,
                                             const float* __restrict__ h, int d, int e0, int e1, int lane,
                                             float4 (&acc)[VN]) {
  constexpr int UNROLL = (VN == 1) ? 8 : 4;
  const char* hb[VN];
#pragma unroll
  for (int q = 0; q < VN; ++q) {
    const int c = (lane + 32 * q) * 4;
    hb[q] = reinterpret_cast<const char*>(h + (c < d ? c : 0));
  }
  const unsigned row_bytes = (unsigned)d * 4u;
  const bool weighted = ew != nullptr;
  for (int e = e0; e < e1; e += 32) {
    const int cnt = min(32, e1 - e);
    const int my = lane < cnt ? gr::ldg_stream_i32(indices + e + lane) : 0;
    float myw = 1.f;
    if (weighted) myw = lane < cnt ? gr::ldg_stream_f32(ew + e + lane) : 0.f;
    int j = 0;
    for (; j + UNROLL <= cnt; j += UNROLL) gather_group<UNROLL, VN, MAXR>(hb, row_bytes, my, myw, weighted, j, acc);
    const int rem = cnt - j;
    if (UNROLL == 8 && (rem & 4)) { gather_group<4, VN, MAXR>(hb, row_bytes, my, myw, weighted, j, acc); j += 4; }
    if (rem & 2) { gather_group<2, VN, MAXR>(hb, row_bytes, my, myw, weighted, j, acc); j += 2; }
    if (rem & 1) gather_group<1, VN, MAXR>(hb, row_bytes, my, myw, weighted, j, acc);
  }
}

// accumulators of lanes whose columns lie beyond d (see gather_range) -> 0
template <int VN>
__device__ __forceinline__ void clear_tail_lanes(float4 (&acc)[VN], int d, int lane) {
#pragma unroll
  for (int q = 0; q < VN; ++q)
    if ((lane + 32 * q) * 4 >= d) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// EXACT: IEEE division by the degree (bit-identical to `sum / clamp(deg, 1)`, used by the stand-alone gather);
// otherwise one correctly rounded reciprocal per row and four multiplies (<= 1 ulp apart; the fused kernel's split-MMA
// projection that follows is itself only fp32-accurate, far inside rtol 1e-4).
template <bool MAXR, bool EXACT = true>
__device__ __forceinline__ float4 finalize4(float4 a, int deg) {
  if (deg == 0) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (!MAXR) {
    const float f = (float)deg;
    if (EXACT) {
      a.x = a.x / f; a.y = a.y / f; a.z = a.z / f; a.w = a.w / f;
    } else {
      const float r = __frcp_rn(f);
      a.x *= r; a.y *= r; a.z *= r; a.w *= r;
    }
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

// Reduced neighbour row of `row` in registers: lane l holds columns 4 * (l + 32 q) .. + 3 in acc[q].
template <int VN, bool MAXR>
__device__ __forceinline__ void reduce_row_regs(const int* __restrict__ indptr, const int* __restrict__ indices,
                                                const float* __restrict__ ew, const float* __restrict__ h, int d,
                                                int row, int lane, const LongWs& lw, float4 (&acc)[VN]) {
  const int beg = __ldg(indptr + row), end = __ldg(indptr + row + 1);
  const int deg = end - beg;
  if (deg > LONG_ROW) {
    const int idx = find_long(lw.long_rows, lw.counters[0], row, lane);
#pragma unroll
    for (int q = 0; q < VN; ++q) {
      const int c = (lane + 32 * q) * 4;
      acc[q] = c < d ? *reinterpret_cast<const float4*>(lw.long_agg + (size_t)idx * d + c) : make_float4(0, 0, 0, 0);
    }
  } else {
#pragma unroll
    for (int q = 0; q < VN; ++q) acc[q] = ident4<MAXR>();
    gather_range<VN, MAXR>(indices, ew, h, d, beg, end, lane, acc);
#pragma unroll
    for (int q = 0; q < VN; ++q) acc[q] = finalize4<MAXR, false>(acc[q], deg);
    clear_tail_lanes<VN>(acc, d, lane);
  }
}
__device__ __forceinline__ float absmax4(const float4& v) {
  return fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
}
__device__ __forceinline__ float4 scale4(const float4& v, float s) { return make_float4(v.x * s, v.y * s, v.z * s, v.w * s); }
// power of two that brings m into [1, 2) (1 for m == 0 or non-finite): multiplying by it is exact
__device__ __forceinline__ float pow2_scale(float m) {
  int e = (__float_as_int(m) >> 23) & 0xff;
  if (e == 0 || e == 0xff) return 1.f;
  e = max(min(254 - e, 200), 54);
  return __int_as_float(e << 23);
}
// 1 / s for s = pow2_scale(.) (a normal power of two with exponent field in [54, 200]): exact, no division
__device__ __forceinline__ float pow2_inverse(float s) { return __int_as_float((254 - (__float_as_int(s) >> 23)) << 23); }

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
// Projection epilogue on the tensor cores with fp32 accuracy: 3xTF32. Every fp32 operand is split into
// hi = tf32(x) and lo = tf32(x - hi) (round-to-nearest), and z += lo.hi + hi.lo + hi.hi with fp32 accumulation
// (mma.sync.m16n8k8.tf32): ~3 * 2^-22 relative error, i.e. indistinguishable from an fp32 FFMA GEMM at the
// rtol 1e-4 / atol 1e-5 embedding tolerance, at a fifth of the instructions. The weights arrive pre-split and
// pre-packed in mma B-fragment order (pack_weights_kernel), so a lane fetches both halves of a fragment with ONE
// coalesced 128-bit load; A fragments come from the gathered tile in shared memory (row pitch D + 4: conflict-free).
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = to_tf32(x);
  lo = to_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// packed[(ks * n_tiles + nt) * 32 + lane] = {b0_hi, b1_hi, b0_lo, b1_lo} of the m16n8k8 B fragment of k-step ks
// (8 rows of [W_self^T ; W_neigh^T]) and column tile nt: b0 = W[ks*8 + lane%4][nt*8 + lane/4], b1 = W[.. + 4][..].
__global__ void pack_weights_kernel(const float* __restrict__ ws_t, const float* __restrict__ wn_t, int ds, int dn,
                                    int dout, float4* __restrict__ packed) {
  const int n_tiles = dout / 8, k_steps = (ds + dn) / 8;
  const int total = k_steps * n_tiles * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int lane = i & 31, nt = (i >> 5) % n_tiles, ks = (i >> 5) / n_tiles;
    const int k0 = ks * 8 + (lane & 3), n = nt * 8 + (lane >> 2);
    auto w = [&](int k) { return k < ds ? ws_t[(size_t)k * dout + n] : wn_t[(size_t)(k - ds) * dout + n]; };
    uint32_t h0, l0, h1, l1;
    split_tf32(w(k0), h0, l0);
    split_tf32(w(k0 + 4), h1, l1);
    packed[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
  }
}

// ---- fp16 split variant (default): the same hi/lo idea on mma.sync.m16n8k16.f16 -- half the MMA instructions of the
// tf32 path at twice the rate. fp16 has 11 significant bits like tf32 but only 5 exponent bits, so every tile row is
// scaled by a power of two that brings its largest |value| (self and neighbour part together) into [1, 2), and the
// weights by one global power of two; both are exact, cancel in the L2 normalisation and are divided out otherwise.
// Error: <= 2^-25 absolute per scaled element (lo halves go subnormal), ~0.5 % of the rtol 1e-4 / atol 1e-5 budget,
// independent of the input scale (tests/experiments/exp_split_f16.py).
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split_f16x4(const float4& v, uint2& hi, uint2& lo) {
  split_f16x2(v.x, v.y, hi.x, lo.x);
  split_f16x2(v.z, v.w, hi.y, lo.y);
}
// four 8x8 b16 matrices -> the A fragment of mma.m16n8k16 (lane l supplies the address of row l % 16, k-half l / 16)
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_ptr) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// wscale[0] = power of two bringing max |W| into [1, 2), wscale[1] = its inverse
__global__ void __launch_bounds__(1024) weight_scale_kernel(const float* __restrict__ ws_t, const float* __restrict__ wn_t,
                                                            int n_self, int n_neigh, float* __restrict__ wscale) {
  __shared__ float s_red[32];
  float m = 0.f;
  for (int i = threadIdx.x; i < n_self; i += blockDim.x) m = fmaxf(m, fabsf(ws_t[i]));
  for (int i = threadIdx.x; i < n_neigh; i += blockDim.x) m = fmaxf(m, fabsf(wn_t[i]));
  m = gr::warp_max(m);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = gr::warp_max(threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f);
    if (threadIdx.x == 0) {
      const float sc = pow2_scale(m);
      wscale[0] = sc;
      wscale[1] = 1.f / sc;
    }
  }
}

// packed[(ks * n_tiles + nt) * 32 + lane] = {b0_hi, b1_hi, b0_lo, b1_lo} (each two fp16) of the m16n8k16 B fragment:
// b0 = W[ks*16 + 2*(lane%4) + {0,1}][nt*8 + lane/4], b1 = the same rows + 8; W = wscale * [W_self^T ; W_neigh^T].
__global__ void pack_weights_f16_kernel(const float* __restrict__ ws_t, const float* __restrict__ wn_t, int ds, int dn,
                                        int dout, const float* __restrict__ wscale, uint4* __restrict__ packed) {
  const int n_tiles = dout / 8, k_steps = (ds + dn) / 16;
  const int total = k_steps * n_tiles * 32;
  const float sc = wscale[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int lane = i & 31, nt = (i >> 5) % n_tiles, ks = (i >> 5) / n_tiles;
    const int k0 = ks * 16 + 2 * (lane & 3), n = nt * 8 + (lane >> 2);
    auto w = [&](int k) { return sc * (k < ds ? ws_t[(size_t)k * dout + n] : wn_t[(size_t)(k - ds) * dout + n]); };
    uint4 o;
    split_f16x2(w(k0), w(k0 + 1), o.x, o.z);
    split_f16x2(w(k0 + 8), w(k0 + 9), o.y, o.w);
    packed[i] = o;
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
  const float4* packed;  // pre-split weights in B-fragment order (fused kernel only)
  const float* wscale;   // fp16 variant: {2^p, 2^-p} applied to the weights
  int* tile_counter;     // zero-initialised; groups take tiles from it
};

constexpr int PAD_TF32 = 4, PAD_F16 = 8;  // floats of row padding: conflict-free 32-bit / 64-bit fragment loads

// A CTA is two independent GROUPS of four warps (named barriers); each group walks its own sequence of R-row tiles
// (R = 16 * MT) taken from a global counter: gather phase (LSU / L2 bound) -> projection phase (tensor bound). The six
// groups resident on an SM drift apart, so one group's gather overlaps another group's MMAs -- with one 8-warp tile per
// CTA the three resident CTAs ran in lockstep and the two phases simply added up.
// MT = 16-row m-tiles per warp, NT = 8-column n-tiles per warp (d_out = 32 * NT: the four warps of a group split the
// columns). VN / VS = float4 per lane of a neighbour / self row.
constexpr int GROUP_THREADS = 128;
__device__ __forceinline__ void group_sync(int grp) {
  asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(GROUP_THREADS) : "memory");
}
template <int VN, int VS, int NT, int MT, bool MAXR, bool F16>
__global__ void __launch_bounds__(THREADS, (VN == 1 && NT <= 4) ? 3 : 2) sage_fused_kernel(SageParams p, LongWs lw) {
  constexpr int R = 16 * MT;
  constexpr int PAD = F16 ? PAD_F16 : PAD_TF32;
  extern __shared__ __align__(16) float smem[];
  const int pn = p.dn + PAD, ps = p.ds + PAD;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp >> 2;                     // group of 4 warps
  float* sN = smem + grp * R * (pn + ps);        // tf32 variant: [R][dn + PAD] fp32
  float* sS = sN + R * pn;                       //               [R][ds + PAD] fp32
  // fp16 variant: the same bytes hold FOUR half planes, split once by the gathering warp (the four warps of phase 2
  // would otherwise each redo the split): neighbour hi / lo [R][dn + 8], self hi / lo [R][ds + 8] -- a 16-byte row pad
  // makes both the 8-byte stores of phase 1 and the ldmatrix reads of phase 2 conflict-free
  const int pnh = p.dn + 8, psh = p.ds + 8;
  __half* sNh = reinterpret_cast<__half*>(sN);
  __half* sNl = sNh + R * pnh;
  __half* sSh = sNl + R * pnh;
  __half* sSl = sSh + R * psh;
  __shared__ int s_next_[2], s_tile_[2];
  __shared__ float s_part_[2][4][R];
  __shared__ float s_inv_[2][R];                 // F16: 1 / row scale
  int& s_next = s_next_[grp];
  float (&s_part)[4][R] = s_part_[grp];
  float (&s_inv)[R] = s_inv_[grp];
  const int64_t n_tiles_total = (p.row_end - p.row_begin + R - 1) / R;

 for (;;) {  // persistent: one tile per iteration and group
  if ((threadIdx.x & (GROUP_THREADS - 1)) == 0) {
    s_tile_[grp] = atomicAdd(p.tile_counter, 1);
    s_next = 0;
  }
  group_sync(grp);
  const int64_t tile = s_tile_[grp];
  if (tile >= n_tiles_total) break;
  const int64_t row0 = p.row_begin + tile * R;

  // phase 1: warps pull rows of the tile dynamically (degree skew), gather -> smem
  while (true) {
    int r = 0;
    if (lane == 0) r = atomicAdd(&s_next, 1);
    r = __shfl_sync(FULL, r, 0);
    if (r >= R) break;
    const int64_t row = row0 + r;
    if (row < p.row_end) {
      float4 an[VN], as[VS];
#pragma unroll
      for (int q = 0; q < VS; ++q) {  // in flight while the neighbours are gathered
        const int c = (lane + 32 * q) * 4;
        as[q] = c < p.ds ? gr::ldg_f4(p.h_dst + (size_t)row * p.ds + c) : make_float4(0, 0, 0, 0);
      }
      reduce_row_regs<VN, MAXR>(p.indptr, p.indices, p.ew, p.h_src, p.dn, (int)row, lane, lw, an);
      float sc = 1.f;
      if (F16) {  // one exact power-of-two scale per row, shared by its self and neighbour part
        float m = 0.f;
#pragma unroll
        for (int q = 0; q < VN; ++q) m = fmaxf(m, absmax4(an[q]));
#pragma unroll
        for (int q = 0; q < VS; ++q) m = fmaxf(m, absmax4(as[q]));
        sc = pow2_scale(gr::warp_max(m));
        if (lane == 0) s_inv[r] = pow2_inverse(sc);
      }
#pragma unroll
      for (int q = 0; q < VN; ++q) {
        const int c = (lane + 32 * q) * 4;
        if (c < p.dn) {
          if (F16) {
            uint2 hi, lo;
            split_f16x4(scale4(an[q], sc), hi, lo);
            *reinterpret_cast<uint2*>(sNh + r * pnh + c) = hi;
            *reinterpret_cast<uint2*>(sNl + r * pnh + c) = lo;
          } else {
            *reinterpret_cast<float4*>(sN + r * pn + c) = an[q];
          }
        }
      }
#pragma unroll
      for (int q = 0; q < VS; ++q) {
        const int c = (lane + 32 * q) * 4;
        if (c < p.ds) {
          if (F16) {
            uint2 hi, lo;
            split_f16x4(scale4(as[q], sc), hi, lo);
            *reinterpret_cast<uint2*>(sSh + r * psh + c) = hi;
            *reinterpret_cast<uint2*>(sSl + r * psh + c) = lo;
          } else {
            *reinterpret_cast<float4*>(sS + r * ps + c) = as[q];
          }
        }
      }
    } else if (F16) {  // rows past the end of the shard: zero rows in all four half planes
      if (lane == 0) s_inv[r] = 1.f;
      for (int c = lane * 4; c < p.dn; c += 128) {
        *reinterpret_cast<uint2*>(sNh + r * pnh + c) = make_uint2(0u, 0u);
        *reinterpret_cast<uint2*>(sNl + r * pnh + c) = make_uint2(0u, 0u);
      }
      for (int c = lane * 4; c < p.ds; c += 128) {
        *reinterpret_cast<uint2*>(sSh + r * psh + c) = make_uint2(0u, 0u);
        *reinterpret_cast<uint2*>(sSl + r * psh + c) = make_uint2(0u, 0u);
      }
    } else {
      for (int c = lane * 4; c < p.dn; c += 128) *reinterpret_cast<float4*>(sN + r * pn + c) = make_float4(0, 0, 0, 0);
      for (int c = lane * 4; c < p.ds; c += 128) *reinterpret_cast<float4*>(sS + r * ps + c) = make_float4(0, 0, 0, 0);
    }
  }
  group_sync(grp);

  // phase 2: z = relu([S | N] . [Ws^T ; Wn^T]) on the tensor cores (hi/lo split, 3 products)
  const int cg = warp & 3;                          // column group (d_out / 4 columns); every warp covers all R rows
  const int g = lane >> 2, tig = lane & 3;
  const int n_tiles = p.dout / 8;
  const int ks_self = p.ds / 8, ks_all = (p.ds + p.dn) / 8;
  float acc[MT][NT][4];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
  const float4* wp = p.packed + (size_t)(cg * NT) * 32 + lane;
  float4 bq[NT];
  if (F16) {
    const int ks16_self = p.ds / 16, ks16_all = (p.ds + p.dn) / 16;
    // B fragments double-buffered in two register sets (no copies): while k-step ks runs on one set the loads of
    // k-step ks + 1 (L1 / L2 resident) fill the other
    float4 bq2[NT];
    auto load_b = [&](float4 (&b)[NT], int ks) {
      const float4* src = wp + (size_t)ks * n_tiles * 32;
#pragma unroll
      for (int n = 0; n < NT; ++n) b[n] = __ldg(src + (size_t)n * 32);
    };
    const int lrow = lane & 15, lk = (lane >> 4) * 8;  // this lane's ldmatrix row / k-half
    auto kstep = [&](int ks, const float4 (&b)[NT]) {
      const bool self = ks < ks16_self;
      const int pitch = self ? psh : pnh;
      const int off = lrow * pitch + (self ? ks : ks - ks16_self) * 16 + lk;
      const __half* th = (self ? sSh : sNh) + off;
      const __half* tl = (self ? sSl : sNl) + off;
#pragma unroll
      for (int m = 0; m < MT; ++m) {
        uint32_t ah[4], al[4];
        ldmatrix_x4(ah, th + m * 16 * pitch);
        ldmatrix_x4(al, tl + m * 16 * pitch);
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          const uint32_t bh0 = __float_as_uint(b[n].x), bh1 = __float_as_uint(b[n].y);
          const uint32_t bl0 = __float_as_uint(b[n].z), bl1 = __float_as_uint(b[n].w);
          mma_f16(acc[m][n], al, bh0, bh1);
          mma_f16(acc[m][n], ah, bl0, bl1);
          mma_f16(acc[m][n], ah, bh0, bh1);
        }
      }
    };
    load_b(bq, 0);
    for (int ks = 0; ks < ks16_all; ks += 2) {
      if (ks + 1 < ks16_all) load_b(bq2, ks + 1);
      kstep(ks, bq);
      if (ks + 1 < ks16_all) {
        if (ks + 2 < ks16_all) load_b(bq, ks + 2);
        kstep(ks + 1, bq2);
      }
    }
  } else {
#pragma unroll
  for (int n = 0; n < NT; ++n) bq[n] = __ldg(wp + (size_t)n * 32);
  for (int ks = 0; ks < ks_all; ++ks) {
    float4 bcur[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) bcur[n] = bq[n];
    if (ks + 1 < ks_all) {  // prefetch the next k-step's fragments (L1 / L2 resident)
      const float4* nx = wp + (size_t)(ks + 1) * n_tiles * 32;
#pragma unroll
      for (int n = 0; n < NT; ++n) bq[n] = __ldg(nx + (size_t)n * 32);
    }
    const float* tile = ks < ks_self ? sS : sN;
    const int pitch = ks < ks_self ? ps : pn;
    const int kc = (ks < ks_self ? ks : ks - ks_self) * 8 + tig;
#pragma unroll
    for (int m = 0; m < MT; ++m) {
      const float* a = tile + (m * 16 + g) * pitch + kc;
      uint32_t ah[4], al[4];
      split_tf32(a[0], ah[0], al[0]);
      split_tf32(a[8 * pitch], ah[1], al[1]);
      split_tf32(a[4], ah[2], al[2]);
      split_tf32(a[8 * pitch + 4], ah[3], al[3]);
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        const uint32_t bh0 = __float_as_uint(bcur[n].x), bh1 = __float_as_uint(bcur[n].y);
        const uint32_t bl0 = __float_as_uint(bcur[n].z), bl1 = __float_as_uint(bcur[n].w);
        mma_tf32(acc[m][n], al, bh0, bh1);
        mma_tf32(acc[m][n], ah, bl0, bl1);
        mma_tf32(acc[m][n], ah, bh0, bh1);
      }
    }
  }
  }

  // relu + row sum of squares: a row's columns live in 4 lanes (tig) x 4 warps (cg)
  float ss[MT][2];
#pragma unroll
  for (int m = 0; m < MT; ++m) {
    ss[m][0] = ss[m][1] = 0.f;
#pragma unroll
    for (int n = 0; n < NT; ++n) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[m][n][i] = fmaxf(acc[m][n][i], 0.f);
        ss[m][i >> 1] = fmaf(acc[m][n][i], acc[m][n][i], ss[m][i >> 1]);
      }
    }
  }
  if (p.l2norm) {
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v = ss[m][h];
        v += __shfl_xor_sync(FULL, v, 1);
        v += __shfl_xor_sync(FULL, v, 2);
        if (tig == 0) s_part[cg][m * 16 + h * 8 + g] = v;
      }
    group_sync(grp);
  }
#pragma unroll
  for (int m = 0; m < MT; ++m) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = m * 16 + h * 8 + g;
      const int64_t row = row0 + r;
      float inv_nrm = 1.f;  // z / (|z| or 1 if 0) as z * rcp(|z|): one correctly rounded reciprocal per row
      if (p.l2norm) {
        const float nrm = sqrtf((s_part[0][r] + s_part[1][r]) + (s_part[2][r] + s_part[3][r]));
        inv_nrm = nrm == 0.f ? 1.f : __frcp_rn(nrm);
      }
      float unscale = 1.f;
      if (F16 && !p.l2norm) unscale = s_inv[r] * __ldg(p.wscale + 1);  // exact powers of two
      if (row < p.row_end) {
#pragma unroll
        for (int n = 0; n < NT; ++n) {
          float* o = p.out + (size_t)row * p.dout + cg * (NT * 8) + n * 8 + 2 * tig;
          float z0 = acc[m][n][2 * h], z1 = acc[m][n][2 * h + 1];
          if (p.l2norm) { z0 *= inv_nrm; z1 *= inv_nrm; }
          else if (F16) { z0 *= unscale; z1 *= unscale; }
          if (p.accumulate == GR_ACC_ADD) {
            const float2 prev = *reinterpret_cast<const float2*>(o);
            z0 = prev.x + z0; z1 = prev.y + z1;
          } else if (p.accumulate == GR_ACC_MAX) {
            const float2 prev = *reinterpret_cast<const float2*>(o);
            z0 = fmaxf(prev.x, z0); z1 = fmaxf(prev.y, z1);
          }
          *reinterpret_cast<float2*>(o) = make_float2(z0 * p.z_scale, z1 * p.z_scale);
        }
      }
    }
  }
  group_sync(grp);  // the next tile overwrites this group's shared memory
 }
}

// ------------------------------------------------------------------------------------------------ generic dims
// Any d_neigh / d_self / d_out <= 512 (e.g. raw 2/4-column features when embedding_layer=False, or the reference's
// 192 / 384 / 512-wide presets, main.py:86-87): one warp per row, scalar column loops, rows staged in shared memory.
// Correctness path for shapes the fused tile kernel does not cover, not a tuned kernel.
constexpr int GEN_Q = 16;  // output columns per lane
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
    float z[GEN_Q];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < GEN_Q; ++q) {
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
    for (int q = 0; q < GEN_Q; ++q) {
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

template <int VN, int VS, int NT, int MT, bool MAXR, bool F16>
int launch_fused(const SageParams& p, const LongWs& lw, cudaStream_t st) {
  constexpr int R = 16 * MT;
  const size_t smem = sizeof(float) * 2 * R * (p.dn + p.ds + 2 * (F16 ? PAD_F16 : PAD_TF32));
  auto kern = sage_fused_kernel<VN, VS, NT, MT, MAXR, F16>;
  GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  const int64_t rows = p.row_end - p.row_begin;
  const int64_t tiles = (rows + R - 1) / R;
  const int per_sm = (VN == 1 && NT <= 4) ? 3 : 2;
  const int64_t grid = std::min<int64_t>((tiles + 1) / 2, (int64_t)gr::sm_count() * per_sm);
  kern<<<(unsigned)std::max<int64_t>(grid, 1), THREADS, smem, st>>>(p, lw);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

template <bool MAXR, bool F16>
int dispatch_fused(const SageParams& p, const LongWs& lw, cudaStream_t st, bool* handled) {
  *handled = true;
  const int vn = (p.dn + 127) / 128, vs = (p.ds + 127) / 128;
#define GR_CASE(VN_, VS_, NT_, MT_) \
  if (vn == VN_ && vs == VS_ && p.dout == 32 * NT_) return launch_fused<VN_, VS_, NT_, MT_, MAXR, F16>(p, lw, st);
  GR_CASE(1, 1, 4, 2)  // 128 -> 128 (c1, c2, c5): 64-row tiles
  GR_CASE(1, 1, 2, 2)  // .. -> 64
  GR_CASE(2, 2, 8, 1)  // 256 -> 256 (c3 hidden): 32-row tiles
  GR_CASE(2, 2, 4, 1)  // 256 -> 128 (c3 output)
  GR_CASE(1, 1, 8, 1)  // 128 -> 256
  GR_CASE(2, 2, 2, 1)
#undef GR_CASE
  *handled = false;
  return GR_OK;
}

bool fast_dims(int dn, int ds, int dout) {
  return dn % 8 == 0 && ds % 8 == 0 && dn <= 256 && ds <= 256 && ((dn + 127) / 128 == (ds + 127) / 128) &&
         (dout == 64 || dout == 128 || dout == 256);
}

size_t packed_bytes(int dn, int ds, int dout) { return gr::align_up((size_t)(dn + ds) * dout * 2 * sizeof(float), 256) + 256; }

// [256 bytes: wscale {2^p, 2^-p}][packed B fragments]: what the fused kernel reads instead of the fp32 weights
int pack_weights(const float* w_self_t, const float* w_neigh_t, int d_self, int d_neigh, int d_out, bool f16, char* dst,
                 cudaStream_t st) {
  float* wscale = reinterpret_cast<float*>(dst);
  float4* packed = reinterpret_cast<float4*>(dst + 256);
  if (f16) {
    weight_scale_kernel<<<1, 1024, 0, st>>>(w_self_t, w_neigh_t, d_self * d_out, d_neigh * d_out, wscale);
    GR_LAUNCH_CHECK();
    const int total = (d_neigh + d_self) / 16 * (d_out / 8) * 32;
    pack_weights_f16_kernel<<<(total + 255) / 256, 256, 0, st>>>(w_self_t, w_neigh_t, d_self, d_neigh, d_out, wscale,
                                                                 reinterpret_cast<uint4*>(packed));
    GR_LAUNCH_CHECK();
  } else {
    const int total = (d_neigh + d_self) / 8 * (d_out / 8) * 32;
    pack_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(w_self_t, w_neigh_t, d_self, d_neigh, d_out, packed);
    GR_LAUNCH_CHECK();
  }
  return GR_OK;
}

bool use_f16(int32_t flags, int d_neigh, int d_self) {
  return !(flags & GR_SAGE_FLAG_TF32_EPILOGUE) && d_neigh % 16 == 0 && d_self % 16 == 0;
}

}  // namespace

extern "C" size_t gr_sage_packed_weights_bytes(int32_t d_neigh, int32_t d_self, int32_t d_out) {
  return fast_dims(d_neigh, d_self, d_out) ? packed_bytes(d_neigh, d_self, d_out) : 0;
}

extern "C" int gr_sage_pack_weights(const float* w_self_t, const float* w_neigh_t, int32_t d_neigh, int32_t d_self,
                                    int32_t d_out, int32_t flags, void* packed, gr_stream_t stream) {
  GR_REQUIRE(fast_dims(d_neigh, d_self, d_out), GR_E_INVALID, "these dimensions do not use packed weights");
  GR_REQUIRE(w_self_t && w_neigh_t && packed, GR_E_INVALID, "null pointer");
  GR_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 255) == 0, GR_E_INVALID, "packed buffer must be 256-byte aligned");
  return pack_weights(w_self_t, w_neigh_t, d_self, d_neigh, d_out, use_f16(flags, d_neigh, d_self),
                      static_cast<char*>(packed), static_cast<cudaStream_t>(stream));
}

extern "C" size_t gr_sage_relation_workspace_bytes(int64_t nnz, int32_t d_neigh) {
  return long_ws_layout(nnz < 0 ? 0 : nnz, d_neigh, nullptr, nullptr) + packed_bytes(256, 256, 256);
}

extern "C" int gr_sage_relation_f32(const int32_t* indptr, const int32_t* indices, const float* edge_w_or_null,
                                    int64_t nnz, const float* h_src, const float* h_dst, int64_t row_begin,
                                    int64_t row_end, int32_t d_neigh, int32_t d_self, const float* w_self_t,
                                    const float* w_neigh_t, int32_t d_out, int reducer, int l2norm, int accumulate,
                                    float z_scale, int32_t flags, const void* packed_or_null, float* out, void* ws,
                                    size_t ws_bytes, gr_stream_t stream) {
  GR_REQUIRE(row_begin >= 0 && row_end >= row_begin, GR_E_INVALID, "bad row range");
  GR_REQUIRE(d_neigh > 0 && d_self > 0 && d_out > 0 && d_neigh <= 512 && d_self <= 512 && d_out <= 512, GR_E_INVALID,
             "dimensions must be in [1, 512]");
  GR_REQUIRE(reducer == GR_REDUCE_MEAN || reducer == GR_REDUCE_MAX, GR_E_INVALID, "unknown reducer");
  GR_REQUIRE(accumulate >= GR_ACC_STORE && accumulate <= GR_ACC_MAX, GR_E_INVALID, "unknown accumulate mode");
  if (row_end == row_begin) return GR_OK;
  GR_REQUIRE(indptr && h_src && h_dst && w_self_t && w_neigh_t && out, GR_E_INVALID, "null pointer");
  GR_REQUIRE(nnz == 0 || indices, GR_E_INVALID, "null indices");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SageParams p{indptr, indices, edge_w_or_null, h_src, h_dst, row_begin, row_end, d_neigh, d_self, d_out,
               w_self_t, w_neigh_t, l2norm, accumulate, z_scale, out, nullptr, nullptr, nullptr};
  const bool maxr = reducer == GR_REDUCE_MAX;
  LongWs lw{};
  if (fast_dims(d_neigh, d_self, d_out)) {
    const size_t long_bytes = long_ws_layout(nnz, d_neigh, nullptr, nullptr);
    const size_t need = long_bytes + packed_bytes(d_neigh, d_self, d_out);
    GR_REQUIRE(ws != nullptr && ws_bytes >= need, GR_E_WORKSPACE, "workspace too small");
    GR_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, GR_E_INVALID, "workspace must be 256-byte aligned");
    long_ws_layout(nnz, d_neigh, &lw, static_cast<char*>(ws));
    const bool f16 = use_f16(flags, d_neigh, d_self);
    const char* pk = static_cast<const char*>(packed_or_null);
    if (pk == nullptr) {  // not packed by the caller (gr_sage_pack_weights, once per weight update): pack per call
      char* own = static_cast<char*>(ws) + long_bytes;
      const int rcp = pack_weights(w_self_t, w_neigh_t, d_self, d_neigh, d_out, f16, own, st);
      if (rcp != GR_OK) return rcp;
      pk = own;
    }
    const float* wscale = reinterpret_cast<const float*>(pk);
    const float4* packed = reinterpret_cast<const float4*>(pk + 256);
    p.packed = packed;
    p.wscale = wscale;
    p.tile_counter = lw.counters + 2;  // zeroed by launch_long_rows below (it clears all 256 bytes of counters)
    int rc;
    const int vn = (d_neigh + 127) / 128;
    if (maxr) rc = vn == 1 ? launch_long_rows<1, true>(indptr, indices, edge_w_or_null, h_src, d_neigh, row_begin, row_end, lw, st)
                           : launch_long_rows<2, true>(indptr, indices, edge_w_or_null, h_src, d_neigh, row_begin, row_end, lw, st);
    else rc = vn == 1 ? launch_long_rows<1, false>(indptr, indices, edge_w_or_null, h_src, d_neigh, row_begin, row_end, lw, st)
                      : launch_long_rows<2, false>(indptr, indices, edge_w_or_null, h_src, d_neigh, row_begin, row_end, lw, st);
    if (rc != GR_OK) return rc;
    bool handled = false;
    if (f16) rc = maxr ? dispatch_fused<true, true>(p, lw, st, &handled) : dispatch_fused<false, true>(p, lw, st, &handled);
    else rc = maxr ? dispatch_fused<true, false>(p, lw, st, &handled) : dispatch_fused<false, false>(p, lw, st, &handled);
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
