// relu(fc_preagg(h)) of the mean_nn / pool_nn aggregators on the 5th-generation tensor cores.
//
// Replaces (reference src/model.py:151,158) `F.relu(self.fc_preagg(h_neigh))` -- a bias-free square nn.Linear over EVERY
// source row of a relation, i.e. a dense [N, D] x [D, D] contraction (c3: 22M rows x 256 x 256 per step). The legacy
// path (linear.cu, 3xTF32 mma.sync) ran it at 60 TFLOP/s; here it is a tcgen05 GEMM at fp32 accuracy:
//
//   * fp32 accuracy on 16-bit tensor cores: every operand is split into hi + lo fp16 halves and the product is the
//     3-term sum hi.hi + lo.hi + hi.lo accumulated in fp32 in TMEM (error ~3 * 2^-22 relative). fp16 has only 5 exponent
//     bits, so each input row is first scaled by the power of two that brings its largest |value| into [1, 2) and the
//     weights by one global power of two; both are exact and divided out in the epilogue (split_rows_f16_kernel /
//     split_weights_f16_kernel write the 16-bit operand tables, K-major, [hi D | lo D] per row).
//   * CTA pairs (cta_group::2): M = 256 rows (128 from each CTA) x N = D outputs x K = 16 per MMA. The WEIGHTS are the
//     stationary operand: each CTA keeps its half of the split weight table (D/2 output rows x 2 parts x D: 128 KB at
//     D = 256) in shared memory for the whole kernel, loaded once by TMA; the input rows stream through a TMA ring of
//     [128 rows x 64 K] sub-tiles (hi and lo of one K block per stage). Persistent: a pair walks row tiles
//     pair, pair + n_pairs, ... ; two D-column accumulators double-buffer in TMEM so the epilogue of tile t overlaps the
//     MMAs of tile t + 1.
//   * warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (whole warp walks the loop, one elected
//     lane issues from uniform registers) + TMEM alloc, warps 2-5 = epilogue: one thread per output row, tcgen05.ld 32
//     columns at a time, un-scale, ReLU, 128-bit stores.
// HBM traffic per row: 2 D bytes (16-bit operands) read + 4 D written, plus the split pass (4 D read + 4 D written);
// the MMAs (3 x 2 D^2 flop per row) hide under it.
#include <cuda_fp16.h>

#include <algorithm>

#include "tc5.cuh"

namespace {

using namespace gr::tc5;

constexpr int TILE_M = 128;                    // rows per CTA and tile (TMEM lanes)
constexpr int SUB_A = TILE_M * KBLK * 2;       // [128 rows][64 x fp16] swizzled sub-tile = 16 KB
constexpr int NUM_THREADS = 192;
constexpr int MAX_STAGES = 4;

// power of two that brings m into [1, 2) (1 for m == 0 or non-finite): multiplying by it is exact
__device__ __forceinline__ float pow2_scale(float m) {
  int e = (__float_as_int(m) >> 23) & 0xff;
  if (e == 0 || e == 0xff) return 1.f;
  e = max(min(254 - e, 200), 54);
  return __int_as_float(e << 23);
}
__device__ __forceinline__ float pow2_inverse(float s) { return __int_as_float((254 - (__float_as_int(s) >> 23)) << 23); }

__device__ __forceinline__ void split_store(float v, __half* hi_p, __half* lo_p) {
  const __half h = __float2half_rn(v);
  *hi_p = h;
  *lo_p = __float2half_rn(v - __half2float(h));
}

// one warp per row: xq[r] = [hi(x_r * s_r) | lo(..)] (fp16, 2 * d), inv_scale[r] = 1 / s_r
__global__ void __launch_bounds__(256) split_rows_f16_kernel(const float* __restrict__ x, long long n, int d,
                                                             __half* __restrict__ xq, float* __restrict__ inv_scale) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp; r < n; r += n_warps) {
    float4 v[2];
    float m = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = (lane + 32 * q) * 4;
      v[q] = c < d ? gr::ldg_f4(x + r * d + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      m = fmaxf(m, fmaxf(fmaxf(fabsf(v[q].x), fabsf(v[q].y)), fmaxf(fabsf(v[q].z), fabsf(v[q].w))));
    }
    const float sc = pow2_scale(gr::warp_max(m));
    if (lane == 0) inv_scale[r] = pow2_inverse(sc);
    __half* row = xq + r * (long long)(2 * d);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = (lane + 32 * q) * 4;
      if (c < d) {
        __align__(8) __half hi[4], lo[4];
        split_store(v[q].x * sc, hi + 0, lo + 0);
        split_store(v[q].y * sc, hi + 1, lo + 1);
        split_store(v[q].z * sc, hi + 2, lo + 2);
        split_store(v[q].w * sc, hi + 3, lo + 3);
        *reinterpret_cast<uint2*>(row + c) = *reinterpret_cast<const uint2*>(hi);
        *reinterpret_cast<uint2*>(row + d + c) = *reinterpret_cast<const uint2*>(lo);
      }
    }
  }
}

// wq[o] = [hi(s * W[:, o]) | lo(..)] for the TRANSPOSED weight wt[k][o] (what gr_linear_f32 receives); wscale = {s, 1/s}
__global__ void __launch_bounds__(1024) split_weights_f16_kernel(const float* __restrict__ wt, int d_in, int d_out,
                                                                 __half* __restrict__ wq, float* __restrict__ wscale) {
  __shared__ float s_red[32];
  __shared__ float s_scale;
  float m = 0.f;
  for (int i = threadIdx.x; i < d_in * d_out; i += blockDim.x) m = fmaxf(m, fabsf(wt[i]));
  m = gr::warp_max(m);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = gr::warp_max(threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.f);
    if (threadIdx.x == 0) {
      const float sc = pow2_scale(m);
      s_scale = sc;
      wscale[0] = sc;
      wscale[1] = pow2_inverse(sc);
    }
  }
  __syncthreads();
  const float sc = s_scale;
  for (int i = threadIdx.x; i < d_in * d_out; i += blockDim.x) {
    const int k = i / d_out, o = i % d_out;  // coalesced read of wt, scattered 2-byte writes (64K elements: irrelevant)
    split_store(wt[i] * sc, wq + (size_t)o * 2 * d_in + k, wq + (size_t)o * 2 * d_in + d_in + k);
  }
}

// KB = D / 64 K-blocks; NH = D / 2 weight rows held by each CTA of the pair.
template <int KB>
struct Cfg {
  static constexpr int D = KB * KBLK;
  static constexpr int NH = D / 2;
  static constexpr int SUB_B = NH * KBLK * 2;            // [NH rows][64 x fp16]
  static constexpr int B_BYTES = 2 * KB * SUB_B;         // both parts, all K blocks: resident
  static constexpr int A_STAGE = 2 * SUB_A;              // hi + lo sub-tile of one K block
  static constexpr int STAGES = ((SMEM_LIMIT - 2048 - B_BYTES) / A_STAGE) > MAX_STAGES ? MAX_STAGES
                                                                                        : ((SMEM_LIMIT - 2048 - B_BYTES) / A_STAGE);
  static constexpr int TMEM_COLS = 2 * D;                // two D-column accumulators
  static constexpr size_t SMEM = 1024 + B_BYTES + STAGES * A_STAGE + 512;
};

template <int KB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
linear_tc5_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, long long n_rows,
                  const float* __restrict__ inv_scale, const float* __restrict__ wscale, int relu,
                  float* __restrict__ y) {
  using L = Cfg<KB>;
  constexpr int D = L::D;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = base;
  uint8_t* sA = sB + L::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + L::STAGES * L::A_STAGE);
  uint64_t* full = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty = full + MAX_STAGES;      // [STAGES]  MMA -> TMA
  uint64_t* b_full = empty + MAX_STAGES;    // [1]       weights resident
  uint64_t* t_full = b_full + 1;            // [2]       MMA -> epilogue
  uint64_t* t_empty = t_full + 2;           // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const long long n_tiles = (n_rows + 2 * TILE_M - 1) / (2 * TILE_M);   // a tile = 256 rows = 128 per CTA of the pair
  const long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < L::STAGES; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(b_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 8); }  // 4 epilogue warps x 2 CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(L::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0 && pair < n_tiles) {
      if (leader) mbar_expect_tx(b_full, 2 * L::B_BYTES);
      for (int pa = 0; pa < 2; ++pa)
        for (int kb = 0; kb < KB; ++kb)   // this CTA's NH output rows of the split weight table
          tma_load_2d_pair(sB + (pa * KB + kb) * L::SUB_B, &tm_w, pa * D + kb * KBLK, (int)crank * L::NH, b_full);
      int st = 0;
      uint32_t phase = 0;
      for (long long t = pair; t < n_tiles; t += n_pairs) {
        const int row0 = (int)(t * 2 * TILE_M) + (int)crank * TILE_M;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(empty + st, phase ^ 1u);
          if (leader) mbar_expect_tx(full + st, 2 * L::A_STAGE);   // both CTAs' bytes count on the leader's barrier
          tma_load_2d_pair(sA + st * L::A_STAGE, &tm_x, kb * KBLK, row0, full + st);              // hi
          tma_load_2d_pair(sA + st * L::A_STAGE + SUB_A, &tm_x, D + kb * KBLK, row0, full + st);  // lo
          if (++st == L::STAGES) { st = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (leader CTA): warp-uniform loop, one elected lane issues =================
    if (leader && pair < n_tiles) {
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
      constexpr uint32_t idesc = make_idesc_mn(0u /* fp16 */, 2 * TILE_M, (uint32_t)D);
      mbar_wait(b_full, 0);
      tc_fence_after();
      int st = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (long long t = pair; t < n_tiles; t += n_pairs, ++it) {
        const int buf = (int)(it & 1);
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        const uint32_t d_tmem = tmem_u + (uint32_t)(buf * D);
        mbar_wait(t_empty + buf, aphase ^ 1u);   // drained by the epilogue of the tile two iterations back
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(full + st, phase);
          tc_fence_after();
          const uint64_t a_hi = make_desc_sw128(sA_u + (uint32_t)(st * L::A_STAGE));
          const uint64_t a_lo = make_desc_sw128(sA_u + (uint32_t)(st * L::A_STAGE + SUB_A));
          const uint64_t b_hi = make_desc_sw128(sB_u + (uint32_t)(kb * L::SUB_B));
          const uint64_t b_lo = make_desc_sw128(sB_u + (uint32_t)((KB + kb) * L::SUB_B));
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < KBLK / UMMA_K; ++k) {  // hi.hi + lo.hi + hi.lo
              tc_mma_f16_pair(d_tmem, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
              tc_mma_f16_pair(d_tmem, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, 1u);
              tc_mma_f16_pair(d_tmem, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), idesc, 1u);
            }
            tc_commit_pair(empty + st);
            if (kb == KB - 1) tc_commit_pair(t_full + buf);
          }
          __syncwarp();
          if (++st == L::STAGES) { st = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ================= epilogue: one thread = one output row =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const float inv_w = __ldg(wscale + 1);
    long long it = 0;
    for (long long t = pair; t < n_tiles; t += n_pairs, ++it) {
      const int buf = (int)(it & 1);
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const long long row = t * 2 * TILE_M + (long long)crank * TILE_M + q * 32 + lane;
      const bool live = row < n_rows;
      const float un = live ? __ldg(inv_scale + row) * inv_w : 0.f;
      mbar_wait(t_full + buf, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * D);
      float* out = y + row * D;
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t v[32];
        tc_ld32(taddr + (uint32_t)c0, v);
        tc_ld_wait();
        if (c0 + 32 >= D) {  // last chunk read: hand the accumulator back before the stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(t_empty + buf);
        }
        if (live) {
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 o;
            o.x = __uint_as_float(v[i]) * un; o.y = __uint_as_float(v[i + 1]) * un;
            o.z = __uint_as_float(v[i + 2]) * un; o.w = __uint_as_float(v[i + 3]) * un;
            if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(out + c0 + i) = o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(L::TMEM_COLS));
  }
}

struct Layout { size_t wq, wscale, inv, xq, total; };
Layout layout(int64_t n, int d) {
  Layout l;
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = gr::align_up(off + b, 256); return o; };
  l.wq = take((size_t)d * 2 * d * 2);
  l.wscale = take(256);
  l.inv = take(sizeof(float) * (size_t)std::max<int64_t>(n, 1));
  l.xq = take((size_t)std::max<int64_t>(n, 1) * 2 * d * 2);
  l.total = off;
  return l;
}

template <int KB>
int launch(const float* x, int64_t n, const float* wt, int relu, float* y, char* ws, cudaStream_t st) {
  using L = Cfg<KB>;
  constexpr int D = L::D;
  static_assert(L::STAGES >= 2, "input ring too small");
  const Layout l = layout(n, D);
  __half* wq = reinterpret_cast<__half*>(ws + l.wq);
  float* wscale = reinterpret_cast<float*>(ws + l.wscale);
  float* inv = reinterpret_cast<float*>(ws + l.inv);
  __half* xq = reinterpret_cast<__half*>(ws + l.xq);
  split_weights_f16_kernel<<<1, 1024, 0, st>>>(wt, D, D, wq, wscale);
  GR_LAUNCH_CHECK();
  const int g1 = (int)std::min<int64_t>((n + 7) / 8, (int64_t)gr::sm_count() * 16);
  split_rows_f16_kernel<<<g1, 256, 0, st>>>(x, n, D, xq, inv);
  GR_LAUNCH_CHECK();
  CUtensorMap mx, mw;
  int rc = make_map(&mx, reinterpret_cast<const uint16_t*>(xq), n, 2 * D, GR_ELEM_FP16, TILE_M);
  if (rc != GR_OK) return rc;
  rc = make_map(&mw, reinterpret_cast<const uint16_t*>(wq), D, 2 * D, GR_ELEM_FP16, L::NH);
  if (rc != GR_OK) return rc;
  auto kern = linear_tc5_kernel<KB>;
  GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::SMEM));
  const int64_t tiles = (n + 2 * TILE_M - 1) / (2 * TILE_M);
  const unsigned pairs = (unsigned)std::min<int64_t>(tiles, gr::sm_count() / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = L::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  GR_CUDA(cudaLaunchKernelEx(&cfg, kern, mx, mw, (long long)n, (const float*)inv, (const float*)wscale, relu, y));
  GR_LAUNCH_CHECK();
  return GR_OK;
}

}  // namespace

namespace gr {

// d_in == d_out in {128, 256}, no bias: the shapes of fc_preagg this kernel serves
bool linear_tc5_supported(int d_in, int d_out, bool has_bias) {
  return !has_bias && d_in == d_out && (d_in == 128 || d_in == 256);
}

size_t linear_tc5_workspace_bytes(int64_t n, int d) { return layout(n, d).total; }

int linear_tc5(const float* x, int64_t n, int d, const float* wt, int relu, float* y, void* ws, cudaStream_t st) {
  char* base = static_cast<char*>(ws);
  return d == 128 ? launch<2>(x, n, wt, relu, y, base, st) : launch<4>(x, n, wt, relu, y, base, st);
}

}  // namespace gr
