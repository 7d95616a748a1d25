// gr_edge_cosine_f32: CosinePrediction.forward for one etype (reference src/model.py:317-327).
// Replaces two F.normalize passes over *every* node row plus DGL's SDDMM u_dot_v with one fused
// gather + dot + normalise per edge: out[e] = <a, b> / (max(|a|, 1e-12) * max(|b|, 1e-12)).
// Eight lanes per edge when d <= 128 (float4 per lane per 128-byte slice), so a warp scores four edges at a time;
// negatives arrive K-consecutive per positive edge (src/model.py:516), so the source row stays hot in L1.
#include "common.cuh"

namespace {

template <int LANES>  // lanes cooperating on one edge: 8, 16 or 32
__global__ void __launch_bounds__(256) edge_cosine_kernel(const int* __restrict__ u, const int* __restrict__ v,
                                                          int64_t n_edges, const float* __restrict__ hs,
                                                          const float* __restrict__ hd, int d,
                                                          float* __restrict__ out) {
  constexpr int EPW = 32 / LANES;  // edges per warp
  const int lane = threadIdx.x & 31;
  const int sub = lane % LANES, slot = lane / LANES;
  const int64_t warp_global = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t e0 = warp_global * EPW; e0 < n_edges; e0 += n_warps * EPW) {
    const int64_t e = e0 + slot;
    float dot = 0.f, na = 0.f, nb = 0.f;
    if (e < n_edges) {
      const float* a = hs + (size_t)gr::ldg_stream_i32(u + e) * d;
      const float* b = hd + (size_t)gr::ldg_stream_i32(v + e) * d;
      if ((d & 3) == 0) {
        for (int c = sub * 4; c < d; c += LANES * 4) {
          const float4 x = gr::ldg_f4(a + c), y = gr::ldg_f4(b + c);
          dot = fmaf(x.x, y.x, dot); dot = fmaf(x.y, y.y, dot); dot = fmaf(x.z, y.z, dot); dot = fmaf(x.w, y.w, dot);
          na = fmaf(x.x, x.x, na); na = fmaf(x.y, x.y, na); na = fmaf(x.z, x.z, na); na = fmaf(x.w, x.w, na);
          nb = fmaf(y.x, y.x, nb); nb = fmaf(y.y, y.y, nb); nb = fmaf(y.z, y.z, nb); nb = fmaf(y.w, y.w, nb);
        }
      } else {
        for (int c = sub; c < d; c += LANES) {
          const float x = __ldg(a + c), y = __ldg(b + c);
          dot = fmaf(x, y, dot); na = fmaf(x, x, na); nb = fmaf(y, y, nb);
        }
      }
    }
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) {
      dot += __shfl_xor_sync(gr::FULL, dot, o);
      na += __shfl_xor_sync(gr::FULL, na, o);
      nb += __shfl_xor_sync(gr::FULL, nb, o);
    }
    if (e < n_edges && sub == 0) out[e] = dot / (fmaxf(sqrtf(na), 1e-12f) * fmaxf(sqrtf(nb), 1e-12f));
  }
}

}  // namespace

extern "C" int gr_edge_cosine_f32(const int32_t* u, const int32_t* v, int64_t n_edges, const float* h_src,
                                  const float* h_dst, int32_t d, float* out, gr_stream_t stream) {
  GR_REQUIRE(n_edges >= 0 && d > 0, GR_E_INVALID, "bad shape");
  if (n_edges == 0) return GR_OK;
  GR_REQUIRE(u && v && h_src && h_dst && out, GR_E_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int lanes = d <= 128 ? 8 : (d <= 256 ? 16 : 32);
  const int epw = 32 / lanes;
  const int64_t warps = (n_edges + epw - 1) / epw;
  const int grid = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)gr::sm_count() * 16);
  if (lanes == 8) edge_cosine_kernel<8><<<grid, 256, 0, st>>>(u, v, n_edges, h_src, h_dst, d, out);
  else if (lanes == 16) edge_cosine_kernel<16><<<grid, 256, 0, st>>>(u, v, n_edges, h_src, h_dst, d, out);
  else edge_cosine_kernel<32><<<grid, 256, 0, st>>>(u, v, n_edges, h_src, h_dst, d, out);
  GR_LAUNCH_CHECK();
  return GR_OK;
}
