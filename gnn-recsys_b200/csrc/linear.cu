// gr_linear_f32: y = x . wt (+ bias) (relu), fp32 FFMA.
//   - NodeEmbedding.forward (reference src/model.py:19-24): d_in is 2 / 4, the kernel is a pure output-write stream.
//   - relu(fc_preagg(h)) of mean_nn / pool_nn (reference src/model.py:151,158): square D x D projection of every
//     source row; classic 128x128x8 smem-tiled register-blocked SGEMM (fp32 accuracy is required by the
//     rtol 1e-4 embedding tolerance, so no single-pass TF32/bf16 tensor math here).
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

}  // namespace

extern "C" int gr_linear_f32(const float* x, int64_t n, int32_t d_in, const float* wt, const float* bias_or_null,
                             int32_t d_out, int relu, float* y, gr_stream_t stream) {
  GR_REQUIRE(n >= 0 && d_in > 0 && d_out > 0, GR_E_INVALID, "bad shape");
  if (n == 0) return GR_OK;
  GR_REQUIRE(x && wt && y, GR_E_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (d_in <= 8) {
    const int64_t total = n * ((d_out + 3) / 4);
    const int grid = (int)std::min<int64_t>((total + 255) / 256, (int64_t)gr::sm_count() * 16);
    linear_small_kernel<<<grid, 256, 0, st>>>(x, n, d_in, wt, bias_or_null, d_out, relu, y);
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
