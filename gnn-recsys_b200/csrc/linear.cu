// gr_linear_f32: y = x . wt (+ bias) (relu), fp32 FFMA.
//   - NodeEmbedding.forward (reference src/model.py:19-24): d_in is 2 / 4, the kernel is a pure output-write stream.
//   - relu(fc_preagg(h)) of mean_nn / pool_nn (reference src/model.py:151,158): square D x D projection of every
//     source row; classic 128x128x8 smem-tiled register-blocked SGEMM (fp32 accuracy is required by the
//     rtol 1e-4 embedding tolerance, so no single-pass TF32/bf16 tensor math here).
#include <algorithm>

#include "common.cuh"

namespace {

// ---- tiny d_in: one thread = one row x 4 output columns; wt and bias are read through L1 ---------------------
__global__ void __launch_bounds__(256) linear_small_kernel(const float* __restrict__ x, int64_t n, int d_in,
                                                           const float* __restrict__ wt,
                                                           const float* __restrict__ bias, int d_out, int relu,
                                                           float* __restrict__ y) {
  const int cols4 = (d_out + 3) >> 2;
  const int64_t total = n * cols4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = (bias != nullptr && c + j < d_out) ? __ldg(bias + c + j) : 0.f;
    for (int k = 0; k < d_in; ++k) {
      const float a = __ldg(x + r * d_in + k);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < d_out) acc[j] = fmaf(a, __ldg(wt + (size_t)k * d_out + c + j), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (relu) acc[j] = fmaxf(acc[j], 0.f);
    if ((d_out & 3) == 0) {
      *reinterpret_cast<float4*>(y + r * d_out + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (c + j < d_out) y[r * d_out + c + j] = acc[j];
    }
  }
}

// ---- tiny d_in, d_out % 4 == 0 and 256 % (d_out / 4) == 0 (the NodeEmbedding shapes: 2 / 4 -> 64 / 128 / 256): a thread
// keeps ONE group of 4 output columns for the whole kernel, so its d_in x 4 weights and 4 biases live in registers; per
// row it reads d_in broadcast inputs, does 4 * d_in FMAs and writes one coalesced 128-bit store. No per-element index
// division, no weight loads in the loop: the kernel is the output-write stream it should be.
template <int DIN>
__global__ void __launch_bounds__(256) linear_small_reg_kernel(const float* __restrict__ x, int64_t n,
                                                               const float* __restrict__ wt,
                                                               const float* __restrict__ bias, int d_out, int relu,
                                                               float* __restrict__ y) {
  const int cols4 = d_out >> 2;
  const int rows_per_block = 256 / cols4;
  const int c = (threadIdx.x % cols4) * 4;
  const int r_in_block = threadIdx.x / cols4;
  float4 w[DIN];
#pragma unroll
  for (int k = 0; k < DIN; ++k) w[k] = __ldg(reinterpret_cast<const float4*>(wt + (size_t)k * d_out + c));
  const float4 b = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t r = (int64_t)blockIdx.x * rows_per_block + r_in_block; r < n; r += (int64_t)gridDim.x * rows_per_block) {
    float a[DIN];
#pragma unroll
    for (int k = 0; k < DIN; ++k) a[k] = __ldg(x + r * DIN + k);
    float4 acc = b;
#pragma unroll
    for (int k = 0; k < DIN; ++k) {  // same order as the generic kernel: bias, then k ascending
      acc.x = fmaf(a[k], w[k].x, acc.x); acc.y = fmaf(a[k], w[k].y, acc.y);
      acc.z = fmaf(a[k], w[k].z, acc.z); acc.w = fmaf(a[k], w[k].w, acc.w);
    }
    if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    *reinterpret_cast<float4*>(y + r * d_out + c) = acc;
  }
}

template <int DIN>
void launch_small_reg(const float* x, int64_t n, const float* wt, const float* bias, int d_out, int relu, float* y,
                      cudaStream_t st) {
  const int rows_per_block = 256 / (d_out / 4);
  const int grid = (int)std::min<int64_t>((n + rows_per_block - 1) / rows_per_block, (int64_t)gr::sm_count() * 16);
  linear_small_reg_kernel<DIN><<<grid, 256, 0, st>>>(x, n, wt, bias, d_out, relu, y);
}

// ---- general: C[M,N] = A[M,K] . B[K,N]; 128x128 tile, BK = 8, 256 threads x (8x8) outputs ---------------------
constexpr int BM = 128, BN = 128, BK = 8;

template <bool VEC>
__global__ void __launch_bounds__(256) linear_tiled_kernel(const float* __restrict__ A, int64_t M, int K,
                                                           const float* __restrict__ B,
                                                           const float* __restrict__ bias, int N, int relu,
                                                           float* __restrict__ C) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int ty = tid >> 4, tx = tid & 15;
  // loader mapping: A: row = tid / 2, k-offset = (tid % 2) * 4;  B: k = tid / 32, col = (tid % 32) * 4
  const int a_row = tid >> 1, a_k = (tid & 1) * 4;
  const int b_k = tid >> 5, b_col = (tid & 31) * 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto load_tiles = [&](int kt, float (&ra)[4], float (&rb)[4]) {
    const int64_t gr_ = m0 + a_row;
    const int gk = kt * BK + a_k;
    if (VEC) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr_ < M && gk < K) v = __ldg(reinterpret_cast<const float4*>(A + gr_ * K + gk));
      ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) ra[j] = (gr_ < M && gk + j < K) ? __ldg(A + gr_ * K + gk + j) : 0.f;
    }
    const int bk = kt * BK + b_k;
    const int bc = n0 + b_col;
    if (VEC) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (bk < K && bc < N) v = __ldg(reinterpret_cast<const float4*>(B + (size_t)bk * N + bc));
      rb[0] = v.x; rb[1] = v.y; rb[2] = v.z; rb[3] = v.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) rb[j] = (bk < K && bc + j < N) ? __ldg(B + (size_t)bk * N + bc + j) : 0.f;
    }
  };
  auto store_tiles = [&](int buf, const float (&ra)[4], const float (&rb)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) As[buf][a_k + j][a_row] = ra[j];
    *reinterpret_cast<float4*>(&Bs[buf][b_k][b_col]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
  };

  const int n_kt = (K + BK - 1) / BK;
  float ra[4], rb[4];
  load_tiles(0, ra, rb);
  store_tiles(0, ra, rb);
  __syncthreads();
  for (int kt = 0; kt < n_kt; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < n_kt) load_tiles(kt + 1, ra, rb);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < n_kt) store_tiles(buf ^ 1, ra, rb);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = m0 + ty * 8 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = n0 + tx * 8 + j;
      if (c >= N) continue;
      float v = acc[i][j] + (bias != nullptr ? __ldg(bias + c) : 0.f);
      if (relu) v = fmaxf(v, 0.f);
      C[r * N + c] = v;
    }
  }
}


// ---- tensor-core path (d_in % 8 == 0, d_out % 32 == 0): 3xTF32 mma.sync, fp32-accurate (see sage.cu) ------------
// CTA = 128 rows x 128 columns, 8 warps as 2 row groups x 4 column groups (64 x 32 per warp); A streams through a
// cp.async double buffer in 32-column chunks (row pitch 36: conflict-free fragment loads); W arrives pre-split in
// B-fragment order (one 128-bit load per fragment and lane).
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
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;  // src-size 0 = zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(bytes) : "memory");
}

__global__ void pack_linear_weights_kernel(const float* __restrict__ wt, int d_in, int d_out, float4* __restrict__ packed) {
  const int n_tiles = d_out / 8, total = d_in / 8 * n_tiles * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int lane = i & 31, nt = (i >> 5) % n_tiles, ks = (i >> 5) / n_tiles;
    const int k0 = ks * 8 + (lane & 3), n = nt * 8 + (lane >> 2);
    uint32_t h0, l0, h1, l1;
    split_tf32(wt[(size_t)k0 * d_out + n], h0, l0);
    split_tf32(wt[(size_t)(k0 + 4) * d_out + n], h1, l1);
    packed[i] = make_float4(__uint_as_float(h0), __uint_as_float(h1), __uint_as_float(l0), __uint_as_float(l1));
  }
}

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_PITCH = TC_BK + 4;

__global__ void __launch_bounds__(256, 2) linear_tc_kernel(const float* __restrict__ A, int64_t M, int K,
                                                           const float4* __restrict__ packed,
                                                           const float* __restrict__ bias, int N, int relu,
                                                           float* __restrict__ C) {
  __shared__ __align__(16) float sA[2][TC_BM][TC_PITCH];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rg = warp & 1, cg = warp >> 1, g = lane >> 2, tig = lane & 3;
  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * TC_BN;
  const int n_tiles = N / 8;
  const int n_chunks = K / TC_BK + (K % TC_BK != 0);
  auto load_chunk = [&](int ch, int buf) {  // 128 rows x 32 floats = 1024 x 16 B, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256, r = idx >> 3, c4 = (idx & 7) * 4;
      const int64_t row = m0 + r;
      const int k = ch * TC_BK + c4;
      const bool ok = row < M && k < K;
      cp_async16(&sA[buf][r][c4], A + (ok ? row * K + k : 0), ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float acc[4][4][4];
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int n = 0; n < 4; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
  const bool col_ok = n0 + cg * 32 < N;
  const float4* wp = packed + (size_t)((n0 + cg * 32) / 8) * 32 + lane;
  load_chunk(0, 0);
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < n_chunks) {
      load_chunk(ch + 1, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const int ks_n = min(TC_BK, K - ch * TC_BK) / 8;
    for (int kk = 0; kk < ks_n; ++kk) {
      const int ks = ch * (TC_BK / 8) + kk;
      float4 b[4];
#pragma unroll
      for (int n = 0; n < 4; ++n) b[n] = col_ok ? __ldg(wp + ((size_t)ks * n_tiles + n) * 32) : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const float* a = &sA[buf][rg * 64 + m * 16 + g][kk * 8 + tig];
        uint32_t ah[4], al[4];
        split_tf32(a[0], ah[0], al[0]);
        split_tf32(a[8 * TC_PITCH], ah[1], al[1]);
        split_tf32(a[4], ah[2], al[2]);
        split_tf32(a[8 * TC_PITCH + 4], ah[3], al[3]);
#pragma unroll
        for (int n = 0; n < 4; ++n) {
          mma_tf32(acc[m][n], al, __float_as_uint(b[n].x), __float_as_uint(b[n].y));
          mma_tf32(acc[m][n], ah, __float_as_uint(b[n].z), __float_as_uint(b[n].w));
          mma_tf32(acc[m][n], ah, __float_as_uint(b[n].x), __float_as_uint(b[n].y));
        }
      }
    }
    __syncthreads();
  }
  if (!col_ok) return;
#pragma unroll
  for (int m = 0; m < 4; ++m)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t row = m0 + rg * 64 + m * 16 + h * 8 + g;
      if (row >= M) continue;
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        const int c = n0 + cg * 32 + n * 8 + 2 * tig;
        float z0 = acc[m][n][2 * h], z1 = acc[m][n][2 * h + 1];
        if (bias != nullptr) { z0 += __ldg(bias + c); z1 += __ldg(bias + c + 1); }
        if (relu) { z0 = fmaxf(z0, 0.f); z1 = fmaxf(z1, 0.f); }
        *reinterpret_cast<float2*>(C + row * N + c) = make_float2(z0, z1);
      }
    }
}

}  // namespace

namespace gr {  // linear_tc5.cu: the tcgen05 path for the square bias-free projections (fc_preagg)
bool linear_tc5_supported(int d_in, int d_out, bool has_bias);
size_t linear_tc5_workspace_bytes(int64_t n, int d);
int linear_tc5(const float* x, int64_t n, int d, const float* wt, int relu, float* y, void* ws, cudaStream_t st);
}  // namespace gr

static size_t legacy_workspace_bytes(int32_t d_in, int32_t d_out) {
  return gr::align_up((size_t)d_in * d_out * 2 * sizeof(float), 256);
}

constexpr int64_t TC5_MIN_ROWS = 256;  // below one row tile the three launches of the tcgen05 path are not worth it

extern "C" size_t gr_linear_workspace_bytes(int64_t n, int32_t d_in, int32_t d_out) {
  if (d_in <= 0 || d_out <= 0) return 256;
  size_t need = legacy_workspace_bytes(d_in, d_out);
  if (gr::linear_tc5_supported(d_in, d_out, false) && n >= TC5_MIN_ROWS)
    need = std::max(need, gr::linear_tc5_workspace_bytes(n, d_in));
  return need;
}

extern "C" int gr_linear_f32(const float* x, int64_t n, int32_t d_in, const float* wt, const float* bias_or_null,
                             int32_t d_out, int relu, int32_t flags, float* y, void* ws, size_t ws_bytes,
                             gr_stream_t stream) {
  GR_REQUIRE(n >= 0 && d_in > 0 && d_out > 0, GR_E_INVALID, "bad shape");
  if (n == 0) return GR_OK;
  GR_REQUIRE(x && wt && y, GR_E_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!(flags & GR_LINEAR_FLAG_LEGACY) && gr::linear_tc5_supported(d_in, d_out, bias_or_null != nullptr) &&
      n >= TC5_MIN_ROWS && ws != nullptr && ws_bytes >= gr::linear_tc5_workspace_bytes(n, d_in) &&
      (reinterpret_cast<uintptr_t>(ws) & 255) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    int major = 0, dev = 0;
    GR_CUDA(cudaGetDevice(&dev));
    GR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major == 10) return gr::linear_tc5(x, n, d_in, wt, relu, y, ws, st);
  }
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 7) == 0);
  const bool reg_ok = d_out % 4 == 0 && d_out <= 1024 && 256 % (d_out / 4) == 0 &&
                      (reinterpret_cast<uintptr_t>(y) & 15) == 0 && (reinterpret_cast<uintptr_t>(wt) & 15) == 0 &&
                      (bias_or_null == nullptr || (reinterpret_cast<uintptr_t>(bias_or_null) & 15) == 0);
  if (d_in <= 8 && reg_ok && (d_in == 2 || d_in == 4 || d_in == 8)) {
    if (d_in == 2) launch_small_reg<2>(x, n, wt, bias_or_null, d_out, relu, y, st);
    else if (d_in == 4) launch_small_reg<4>(x, n, wt, bias_or_null, d_out, relu, y, st);
    else launch_small_reg<8>(x, n, wt, bias_or_null, d_out, relu, y, st);
  } else if (d_in <= 8) {
    const int64_t total = n * ((d_out + 3) / 4);
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)gr::sm_count() * 16);
    linear_small_kernel<<<grid, 256, 0, st>>>(x, n, d_in, wt, bias_or_null, d_out, relu, y);
  } else if (d_in % 8 == 0 && d_out % 32 == 0 && aligned && ws != nullptr &&
             ws_bytes >= legacy_workspace_bytes(d_in, d_out) && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) {
    float4* packed = static_cast<float4*>(ws);
    const int total = d_in / 8 * (d_out / 8) * 32;
    pack_linear_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(wt, d_in, d_out, packed);
    GR_LAUNCH_CHECK();
    dim3 grid((unsigned)((n + TC_BM - 1) / TC_BM), (unsigned)((d_out + TC_BN - 1) / TC_BN));
    linear_tc_kernel<<<grid, 256, 0, st>>>(x, n, d_in, packed, bias_or_null, d_out, relu, y);
  } else {
    dim3 grid((unsigned)((n + BM - 1) / BM), (unsigned)((d_out + BN - 1) / BN));
    const bool vec = (d_in % 4 == 0) && (d_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(wt) & 15) == 0);
    if (vec)
      linear_tiled_kernel<true><<<grid, 256, 0, st>>>(x, n, d_in, wt, bias_or_null, d_out, relu, y);
    else
      linear_tiled_kernel<false><<<grid, 256, 0, st>>>(x, n, d_in, wt, bias_or_null, d_out, relu, y);
  }
  GR_LAUNCH_CHECK();
  return GR_OK;
}
