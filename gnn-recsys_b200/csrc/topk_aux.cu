// The CUDA-core stages around the tcgen05 scoring kernel of get_recs (reference src/metrics.py:52-77):
//   gr_colmean_normalized_f32  mean of the L2-normalised rows (the item "centre")
//   gr_score_prep              normalise (- centre), pad, round to bf16 / fp16, optional hi/lo split
//   gr_rescore_topk_f32        exact fp32 cosine of the shortlist, final top-k, soundness proof per user
//   gr_score_topk_exact_f32    exact fp32 scoring of all items for listed users (fallback + brute-force checker)
//   gr_topk_merge              row-wise merge of partial (score, id) lists
// Cosine follows nn.CosineSimilarity(dim=1, eps) as the reference calls it (src/metrics.py:58-59):
//   x.y / sqrt(max(|x|^2 |y|^2, eps^2)).
#include <cuda_fp16.h>

#include <algorithm>

#include "common.cuh"

namespace {

using gr::FULL;

// ------------------------------------------------------------------------------------------------ colmean
constexpr int CM_MAXQ = 8;  // columns per lane: d <= 256

__global__ void __launch_bounds__(256) colmean_partial_kernel(const float* __restrict__ x, long long n, int d,
                                                              float* __restrict__ partials) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  float acc[CM_MAXQ];
#pragma unroll
  for (int q = 0; q < CM_MAXQ; ++q) acc[q] = 0.f;
  for (long long r = warp; r < n; r += n_warps) {
    float v[CM_MAXQ];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < CM_MAXQ; ++q) {
      const int c = lane + 32 * q;
      v[q] = c < d ? __ldg(x + r * d + c) : 0.f;
      ss = fmaf(v[q], v[q], ss);
    }
    const float inv = 1.f / fmaxf(sqrtf(gr::warp_sum(ss)), 1e-12f);
#pragma unroll
    for (int q = 0; q < CM_MAXQ; ++q) acc[q] = fmaf(v[q], inv, acc[q]);
  }
#pragma unroll
  for (int q = 0; q < CM_MAXQ; ++q) {
    const int c = lane + 32 * q;
    if (c < d) partials[warp * d + c] = acc[q];
  }
}

// one block = 32 columns x 32 warps: warp w sums partial rows w, w + 32, ... (coalesced 128-byte reads, 4 independent
// loads in flight), then the 32 per-warp sums are added in warp order -- a fixed order, so the centre is deterministic
__global__ void __launch_bounds__(1024) colmean_final_kernel(const float* __restrict__ partials, long long n_partials,
                                                             long long n, int d, float* __restrict__ center) {
  __shared__ float s_part[32][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (c < d) {
    long long p = w;
    for (; p + 96 < n_partials; p += 128) {
      a0 += partials[p * d + c];
      a1 += partials[(p + 32) * d + c];
      a2 += partials[(p + 64) * d + c];
      a3 += partials[(p + 96) * d + c];
    }
    for (; p < n_partials; p += 32) a0 += partials[p * d + c];
  }
  s_part[w][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (w == 0 && c < d) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) a += s_part[i][lane];
    center[c] = n > 0 ? a / (float)n : 0.f;
  }
}

int colmean_grid() { return gr::sm_count() * 4; }

// ------------------------------------------------------------------------------------------------ prep
__device__ __forceinline__ uint16_t to16(float v, int elem_type) {
  if (elem_type == GR_ELEM_FP16) return __half_as_ushort(__float2half_rn(v));
  return __bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float from16(uint16_t b, int elem_type) {
  if (elem_type == GR_ELEM_FP16) return __half2float(__ushort_as_half(b));
  return __bfloat162float(__ushort_as_bfloat16(b));
}

// One row -> its 16-bit parts. `w` = the normalised (centred) value of this lane's column q (0 beyond d). Returns this
// lane's squared residuals: r1 += (w - hi)^2, r2 += (w - hi - lo)^2 (== r1 for parts == 1). Shared by the prep kernel
// and the re-score kernel, so that the user residuals of the proof are those of the rows the GEMM actually read.
__device__ __forceinline__ void quantise(float w, int parts, int elem_type, uint16_t& hi, uint16_t& lo, float& r1,
                                         float& r2) {
  hi = to16(w, elem_type);
  const float e1 = w - from16(hi, elem_type);  // exact in fp32
  r1 = fmaf(e1, e1, r1);
  lo = 0;
  float e2 = e1;
  if (parts == 2) {
    lo = to16(e1, elem_type);
    e2 = e1 - from16(lo, elem_type);
  }
  r2 = fmaf(e2, e2, r2);
}

constexpr int PREP_MAXQ = 8;  // columns per lane: d_pad <= 256

__global__ void __launch_bounds__(256) score_prep_kernel(const float* __restrict__ x, long long n, int d,
                                                         const float* __restrict__ center, int d_pad, int parts,
                                                         int elem_type, uint16_t* __restrict__ out,
                                                         float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int nq = d_pad / 32;  // 2, 4, 6 or 8
  float max_norm = 0.f, min_norm = INFINITY, max_r1 = 0.f, max_r2 = 0.f;
  for (long long r = warp; r < n; r += n_warps) {
    float v[PREP_MAXQ];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < PREP_MAXQ; ++q) {
      const int c = lane + 32 * q;
      v[q] = (q < nq && c < d) ? __ldg(x + r * d + c) : 0.f;
      ss = fmaf(v[q], v[q], ss);
    }
    const float nrm = sqrtf(gr::warp_sum(ss));
    if (nrm > 0.f) min_norm = fminf(min_norm, nrm);
    const float inv = 1.f / fmaxf(nrm, 1e-12f);
    float cs = 0.f, r1 = 0.f, r2 = 0.f;
    uint16_t* o = out + r * (long long)(parts * d_pad);
#pragma unroll
    for (int q = 0; q < PREP_MAXQ; ++q) {
      const int c = lane + 32 * q;
      if (q < nq) {
        float w = v[q] * inv;
        if (center != nullptr && c < d) w -= __ldg(center + c);
        if (c >= d) w = 0.f;
        cs = fmaf(w, w, cs);
        uint16_t hi, lo;
        quantise(w, parts, elem_type, hi, lo, r1, r2);
        o[c] = hi;
        if (parts == 2) o[d_pad + c] = lo;
      }
    }
    max_norm = fmaxf(max_norm, sqrtf(gr::warp_sum(cs)));
    max_r1 = fmaxf(max_r1, sqrtf(gr::warp_sum(r1)));
    max_r2 = fmaxf(max_r2, sqrtf(gr::warp_sum(r2)));
  }
  if (stats != nullptr && lane == 0) {
    atomicMax(reinterpret_cast<int*>(stats), __float_as_int(max_norm));        // non-negative floats order as ints
    if (min_norm < INFINITY) atomicMin(reinterpret_cast<int*>(stats + 1), __float_as_int(min_norm));
    atomicMax(reinterpret_cast<int*>(stats + 2), __float_as_int(max_r2));
    atomicMax(reinterpret_cast<int*>(stats + 3), __float_as_int(max_r1));
  }
}

// ------------------------------------------------------------------------------------------------ exact cosine
__device__ __forceinline__ void dot_norm(const float* __restrict__ a, const float* __restrict__ b, int d, float& dot,
                                         float& nb) {
  dot = 0.f; nb = 0.f;
  if ((d & 3) == 0) {
    for (int c = 0; c < d; c += 4) {
      const float4 x = *reinterpret_cast<const float4*>(a + c);
      const float4 y = gr::ldg_f4(b + c);
      dot = fmaf(x.x, y.x, dot); dot = fmaf(x.y, y.y, dot); dot = fmaf(x.z, y.z, dot); dot = fmaf(x.w, y.w, dot);
      nb = fmaf(y.x, y.x, nb); nb = fmaf(y.y, y.y, nb); nb = fmaf(y.z, y.z, nb); nb = fmaf(y.w, y.w, nb);
    }
  } else {
    for (int c = 0; c < d; ++c) {
      const float x = a[c], y = __ldg(b + c);
      dot = fmaf(x, y, dot); nb = fmaf(y, y, nb);
    }
  }
}
__device__ __forceinline__ float cosine(float dot, float na, float nb, float eps) {
  return dot / sqrtf(fmaxf(na * nb, eps * eps));
}

// better(a, b): a precedes b in the output order (score descending, id ascending)
__device__ __forceinline__ bool better(float sa, int ia, float sb, int ib) { return sa > sb || (sa == sb && ia < ib); }

// ------------------------------------------------------------------------------------------------ rescore
constexpr int RS_WARPS = 8;

struct RescoreErr {
  int elem_type, parts_users, parts_items;
  float acc_err;
  const float* band;  // device scalar or null
};

// |approximate - exact| of one (user, item) score: user residuals ru (final) / ru1 (first level), item table
// statistics Y = max |row|, R / R1 = max final / first-level residual (include/gnn_recsys_b200.h, stage 2)
__device__ __forceinline__ float score_err(float ru, float ru1, float Y, float R, float R1, int elem_type,
                                           int parts_users, int parts_items, float acc_err) {
  const float u_round = elem_type == GR_ELEM_FP16 ? 0x1p-11f : 0x1p-8f;
  float err = ru * Y + (1.f + ru) * R + acc_err * (1.f + ru) * (Y + R);
  if (parts_users == 2 && parts_items == 2) err += ru1 * R1 * (1.f + u_round) * (1.f + u_round);
  return err;
}

// band = 2 x the largest err_u over the user table (its measured residuals, user stats [2] / [3])
__global__ void score_band_kernel(const float* __restrict__ item_stats, const float* __restrict__ user_stats,
                                  RescoreErr em, float* __restrict__ band) {
  const float ru = user_stats[2] * 1.001f, ru1 = user_stats[3] * 1.001f;
  band[0] = 2.f * score_err(ru, ru1, item_stats[0], item_stats[2], item_stats[3], em.elem_type, em.parts_users,
                            em.parts_items, em.acc_err) * 1.001f;
}

__global__ void __launch_bounds__(RS_WARPS * 32) rescore_kernel(
    const float* __restrict__ hu, const float* __restrict__ hi, long long item_id_base, int d,
    const float* __restrict__ center, const float* __restrict__ sl_score, const int* __restrict__ sl_id, int S,
    long long n_users, const float* __restrict__ stats, RescoreErr em, float tie_tol, int k, float eps,
    const int* __restrict__ user_map, int* __restrict__ out_ids, float* __restrict__ out_scores,
    int* __restrict__ overflow_users, int* __restrict__ n_overflow) {
  extern __shared__ __align__(16) float s_user[];  // [RS_WARPS][d]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* xu = s_user + w * d;
  const float Y = stats[0], min_item = stats[1], R = stats[2], R1 = stats[3];
  const float band = em.band != nullptr ? __ldg(em.band) : INFINITY;
  for (long long u = (long long)blockIdx.x * RS_WARPS + w; u < n_users; u += (long long)gridDim.x * RS_WARPS) {
    float pn = 0.f, pc = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float x = __ldg(hu + u * d + c);
      xu[c] = x;
      pn = fmaf(x, x, pn);
      if (center != nullptr) pc = fmaf(x, __ldg(center + c), pc);
    }
    const float nu = gr::warp_sum(pn);
    const float inv_nrm = 1.f / fmaxf(sqrtf(nu), 1e-12f);  // the same operations as score_prep_kernel (d <= 128 there)
    const float xc = gr::warp_sum(pc) * inv_nrm;
    // rounding residuals of this user's operand row, recomputed exactly as stage 0 rounds it
    float r1 = 0.f, r2 = 0.f;
    for (int c = lane; c < d; c += 32) {
      uint16_t qh, ql;
      quantise(xu[c] * inv_nrm, em.parts_users, em.elem_type, qh, ql, r1, r2);
    }
    const float ru1 = sqrtf(gr::warp_sum(r1)) * 1.001f, ru = sqrtf(gr::warp_sum(r2)) * 1.001f;
    const float err = score_err(ru, ru1, Y, R, R1, em.elem_type, em.parts_users, em.parts_items, em.acc_err);
    __syncwarp();
    int id = -1;
    float approx = -INFINITY, e = -INFINITY;
    if (lane < S) {
      id = sl_id[u * S + lane];
      approx = sl_score[u * S + lane];
    }
    const bool valid = id >= 0;
    const int n_valid = __popc(__ballot_sync(FULL, valid));
    // a candidate whose approximate score lies more than 2 err below the k-th best approximate score cannot reach the
    // exact top-k (k candidates score at least tau_k - err exactly): it is neither gathered nor ranked
    const float tau_k = n_valid >= k ? __shfl_sync(FULL, approx, k - 1) : -INFINITY;
    const bool cand = valid && approx >= tau_k - 2.f * err;
    if (!cand) id = -1;
    if ((d & 31) == 0) {
      // 8 lanes per candidate, 4 candidates per pass: every load instruction covers 4 x 128 contiguous bytes
      // (the one-lane-per-candidate walk touched 32 different rows per instruction, half a sector each)
      const int grp = lane >> 3, sub = lane & 7;
      const unsigned cmask = __ballot_sync(FULL, cand);
      for (int pass = 0; pass * 4 < S; ++pass) {
        if (((cmask >> (pass * 4)) & 0xfu) == 0u) continue;  // warp-uniform: none of these four is a candidate
        const int cid = __shfl_sync(FULL, id, min(pass * 4 + grp, 31));
        float dot = 0.f, ni = 0.f;
        if (cid >= 0 && pass * 4 + grp < S) {
          const float* row = hi + (cid - item_id_base) * (long long)d;
          for (int c = sub * 4; c < d; c += 32) {
            const float4 x = *reinterpret_cast<const float4*>(xu + c);
            const float4 y = gr::ldg_f4(row + c);
            dot = fmaf(x.x, y.x, dot); dot = fmaf(x.y, y.y, dot); dot = fmaf(x.z, y.z, dot); dot = fmaf(x.w, y.w, dot);
            ni = fmaf(y.x, y.x, ni); ni = fmaf(y.y, y.y, ni); ni = fmaf(y.z, y.z, ni); ni = fmaf(y.w, y.w, ni);
          }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
          dot += __shfl_xor_sync(FULL, dot, o);
          ni += __shfl_xor_sync(FULL, ni, o);
        }
        const float eg = cosine(dot, nu, ni, eps);
        const float mine = __shfl_sync(FULL, eg, (lane & 3) * 8);  // candidate `lane` was handled by group lane % 4
        if ((lane >> 2) == pass && id >= 0) e = mine;
      }
    } else if (id >= 0) {
      float dot, ni;
      dot_norm(xu, hi + (id - item_id_base) * (long long)d, d, dot, ni);
      e = cosine(dot, nu, ni, eps);
    }
    int rank = 0;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
      const float ei = __shfl_sync(FULL, e, i);
      const int ii = __shfl_sync(FULL, id, i);
      if (ii >= 0 && i != lane && better(ei, ii, e, id)) ++rank;
    }
    const int n_cand = __popc(__ballot_sync(FULL, cand));
    if (cand && rank < k) {
      out_ids[u * k + rank] = id;
      out_scores[u * k + rank] = e;
    }
    if (lane >= n_cand && lane < k) {
      out_ids[u * k + lane] = -1;
      out_scores[u * k + lane] = -INFINITY;
    }
    // soundness: every item stage 1 dropped has an approximate score <= dropped
    float dropped = -INFINITY;
    if (n_valid == S) dropped = __shfl_sync(FULL, approx, S - 1);
    if (n_valid >= k) dropped = fmaxf(dropped, tau_k - band);  // band == +inf: -inf, the rule was off
    if (dropped > -INFINITY) {
      const unsigned kth_mask = __ballot_sync(FULL, cand && rank == k - 1);
      float kth = -INFINITY;
      if (kth_mask) kth = __shfl_sync(FULL, e, __ffs(kth_mask) - 1);
      const float bound = dropped + err + xc;
      const bool clamp_binds = nu > 0.f && nu * min_item * min_item < eps * eps;
      if (lane == 0 && (S < k || kth < bound - tie_tol || clamp_binds)) {
        const int slot = atomicAdd(n_overflow, 1);
        overflow_users[slot] = user_map != nullptr ? user_map[u] : (int)u;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ exact top-k
constexpr int EX_THREADS = 256;
constexpr int EX_KMAX = 64;  // entries per pass (per-thread sorted lists of that length live in shared memory)

// One pass writes entries [k_off, k_off + k_pass) of every listed user's k_total-long row: the k_pass best items that
// come strictly AFTER entry k_off - 1 in the output order (score desc, id asc). k <= EX_KMAX is a single pass; larger k
// (the reference's get_recs takes any k) costs one sweep over the items per EX_KMAX entries.
__global__ void __launch_bounds__(EX_THREADS) exact_topk_kernel(
    const float* __restrict__ hu, const int* __restrict__ user_list, const int* __restrict__ n_list, long long n_users,
    const float* __restrict__ hi, long long n_items, long long item_id_base, int d,
    const long long* __restrict__ bought_indptr, const int* __restrict__ bought_ids, int k_total, int k_off, int k,
    float eps, const float* __restrict__ popularity, float weight_popularity, int* __restrict__ out_ids,
    float* __restrict__ out_scores) {
  extern __shared__ __align__(16) float smem[];
  float* xu = smem;                                            // [d_al]
  const int d_al = (d + 3) & ~3;
  float* ls = xu + d_al;                                       // [k][EX_THREADS]
  int* li = reinterpret_cast<int*>(ls + k * EX_THREADS);       // [k][EX_THREADS]
  __shared__ float red_s[EX_THREADS / 32];
  __shared__ int red_i[EX_THREADS / 32];
  __shared__ int red_t[EX_THREADS / 32];
  __shared__ float s_nu;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const long long count = n_list != nullptr ? (long long)*n_list : n_users;
  for (long long idx = blockIdx.x; idx < count; idx += gridDim.x) {
    const long long u = user_list != nullptr ? (long long)user_list[idx] : idx;
    __syncthreads();
    // resume point of a later pass: everything up to (after_s, after_i) is already in the row
    float after_s = INFINITY;
    int after_i = -1;
    if (k_off > 0) {
      after_s = out_scores[u * k_total + k_off - 1];
      after_i = out_ids[u * k_total + k_off - 1];
      if (after_i < 0) {  // the previous pass ran out of items
        for (int r = tid; r < k; r += EX_THREADS) {
          out_ids[u * k_total + k_off + r] = -1;
          out_scores[u * k_total + k_off + r] = -INFINITY;
        }
        continue;
      }
    }
    float pn = 0.f;
    for (int c = tid; c < d; c += EX_THREADS) {
      const float x = hu[u * d + c];
      xu[c] = x;
      pn = fmaf(x, x, pn);
    }
    pn = gr::warp_sum(pn);
    if (lane == 0) red_s[w] = pn;
    for (int s = 0; s < k; ++s) { ls[s * EX_THREADS + tid] = -INFINITY; li[s * EX_THREADS + tid] = -1; }
    __syncthreads();
    if (tid == 0) {
      float a = 0.f;
      for (int i = 0; i < EX_THREADS / 32; ++i) a += red_s[i];
      s_nu = a;
    }
    __syncthreads();
    const float nu = s_nu;
    long long b0 = 0, b1 = 0;
    if (bought_indptr != nullptr) { b0 = bought_indptr[u]; b1 = bought_indptr[u + 1]; }
    // popularity re-rank (src/metrics.py:69-72): rating = softmax_i(cos) + w * popularity_i. Pass 1: the softmax
    // denominator Z = sum_i exp(cos_i - 1) (cos <= 1, so the shift plays the role of the reference's max)
    float inv_z = 0.f;
    if (popularity != nullptr) {
      float z = 0.f;
      for (long long i = tid; i < n_items; i += EX_THREADS) {
        float dot, ni;
        dot_norm(xu, hi + i * d, d, dot, ni);
        z += expf(cosine(dot, nu, ni, eps) - 1.f);
      }
      z = gr::warp_sum(z);
      __syncthreads();
      if (lane == 0) red_s[w] = z;
      __syncthreads();
      float a = 0.f;
      for (int i = 0; i < EX_THREADS / 32; ++i) a += red_s[i];
      inv_z = 1.f / a;
      __syncthreads();
    }
    float tau = -INFINITY;
    for (long long i = tid; i < n_items; i += EX_THREADS) {
      float dot, ni;
      dot_norm(xu, hi + i * d, d, dot, ni);
      float s = cosine(dot, nu, ni, eps);
      if (popularity != nullptr) s = fmaf(weight_popularity, __ldg(popularity + i), expf(s - 1.f) * inv_z);
      if (s > tau) {
        const int gid = (int)(item_id_base + i);
        if (k_off > 0 && !better(after_s, after_i, s, gid)) continue;  // already emitted by an earlier pass
        long long lo = b0, hi_ = b1;
        while (lo < hi_) {
          const long long mid = (lo + hi_) >> 1;
          if (bought_ids[mid] < gid) lo = mid + 1; else hi_ = mid;
        }
        if (lo < b1 && bought_ids[lo] == gid) continue;
        int j = k - 1;
        while (j > 0 && ls[(j - 1) * EX_THREADS + tid] < s) {
          ls[j * EX_THREADS + tid] = ls[(j - 1) * EX_THREADS + tid];
          li[j * EX_THREADS + tid] = li[(j - 1) * EX_THREADS + tid];
          --j;
        }
        ls[j * EX_THREADS + tid] = s;
        li[j * EX_THREADS + tid] = gid;
        tau = ls[(k - 1) * EX_THREADS + tid];
      }
    }
    // merge the 256 sorted per-thread lists: k rounds of block-wide argmax over the list heads
    int head = 0;
    for (int r = 0; r < k; ++r) {
      float s = head < k ? ls[head * EX_THREADS + tid] : -INFINITY;
      int id = head < k ? li[head * EX_THREADS + tid] : -1;
      if (id < 0) { s = -INFINITY; id = 0x7fffffff; }
      int who = tid;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float s2 = __shfl_xor_sync(FULL, s, o);
        const int i2 = __shfl_xor_sync(FULL, id, o);
        const int w2 = __shfl_xor_sync(FULL, who, o);
        if (better(s2, i2, s, id)) { s = s2; id = i2; who = w2; }
      }
      if (lane == 0) { red_s[w] = s; red_i[w] = id; red_t[w] = who; }
      __syncthreads();
      float bs = red_s[0];
      int bi = red_i[0], bt = red_t[0];
      for (int i = 1; i < EX_THREADS / 32; ++i)
        if (better(red_s[i], red_i[i], bs, bi)) { bs = red_s[i]; bi = red_i[i]; bt = red_t[i]; }
      const bool none = bi == 0x7fffffff;
      if (tid == 0) {
        out_ids[u * k_total + k_off + r] = none ? -1 : bi;
        out_scores[u * k_total + k_off + r] = none ? -INFINITY : bs;
      }
      if (!none && tid == bt) ++head;
      __syncthreads();
    }
  }
}

// ------------------------------------------------------------------------------------------------ merge
__global__ void __launch_bounds__(256) topk_merge_kernel(const float* __restrict__ scores, const int* __restrict__ ids,
                                                         int parts, long long n_users, int k_in, int k_out,
                                                         float* __restrict__ out_scores, int* __restrict__ out_ids) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long u = warp; u < n_users; u += n_warps) {
    const float* ps = scores + ((long long)lane * n_users + u) * k_in;  // lane p walks list p
    const int* pi = ids + ((long long)lane * n_users + u) * k_in;
    int head = 0;
    float hs = -INFINITY;
    int hid = 0x7fffffff;
    auto load_head = [&]() {
      hs = -INFINITY; hid = 0x7fffffff;
      if (lane < parts && head < k_in) {
        const int id = pi[head];
        if (id >= 0) { hid = id; hs = ps[head]; }
      }
    };
    load_head();
    for (int r = 0; r < k_out; ++r) {
      float s = hs;
      int id = hid, who = lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float s2 = __shfl_xor_sync(FULL, s, o);
        const int i2 = __shfl_xor_sync(FULL, id, o);
        const int w2 = __shfl_xor_sync(FULL, who, o);
        if (better(s2, i2, s, id) || (s2 == s && i2 == id && w2 < who)) { s = s2; id = i2; who = w2; }
      }
      const bool none = id == 0x7fffffff;
      if (lane == 0) {
        out_ids[u * k_out + r] = none ? -1 : id;
        out_scores[u * k_out + r] = none ? -INFINITY : s;
      }
      if (!none && lane == who) { ++head; load_head(); }
    }
  }
}

// ------------------------------------------------------------------------------------------------ metrics@k
// recs_to_metrics (src/metrics.py:81-107) as counters: [0] recommended ids, [1] of those in the user's ground truth,
// [2] ground-truth entries (duplicates kept), [3] of those recommended, [4] distinct recommended items.
__global__ void __launch_bounds__(256) metrics_kernel(const int* __restrict__ recs, long long n_users, int k,
                                                      const long long* __restrict__ t_indptr,
                                                      const int* __restrict__ t_ids, unsigned int* __restrict__ bitmap,
                                                      unsigned long long* __restrict__ counters) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (long long u = warp; u < n_users; u += n_warps) {
    const long long b0 = t_indptr[u], b1 = t_indptr[u + 1];
    for (int r0 = 0; r0 < k; r0 += 32) {  // precision side: 32 recommendations per pass
      const int id = r0 + lane < k ? recs[u * k + r0 + lane] : -1;
      if (id >= 0) {
        ++c0;
        long long lo = b0, hi = b1;
        while (lo < hi) {
          const long long mid = (lo + hi) >> 1;
          if (t_ids[mid] < id) lo = mid + 1; else hi = mid;
        }
        if (lo < b1 && t_ids[lo] == id) ++c1;
        atomicOr(bitmap + (id >> 5), 1u << (id & 31));
      }
    }
    for (long long j = b0 + lane; j < ((b1 - b0 + 31) / 32) * 32 + b0; j += 32) {  // recall side
      const int t = j < b1 ? t_ids[j] : -2;
      bool hit = false;
      for (int r0 = 0; r0 < k; r0 += 32) {
        const int id = r0 + lane < k ? recs[u * k + r0 + lane] : -1;
        const int n = min(32, k - r0);
        for (int i = 0; i < n; ++i) hit |= (__shfl_sync(FULL, id, i) == t);
      }
      if (j < b1) { ++c2; if (hit) ++c3; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    c0 += __shfl_xor_sync(FULL, c0, o); c1 += __shfl_xor_sync(FULL, c1, o);
    c2 += __shfl_xor_sync(FULL, c2, o); c3 += __shfl_xor_sync(FULL, c3, o);
  }
  if (lane == 0) {
    atomicAdd(counters + 0, c0); atomicAdd(counters + 1, c1); atomicAdd(counters + 2, c2); atomicAdd(counters + 3, c3);
  }
}

__global__ void popcount_kernel(const unsigned int* __restrict__ bitmap, long long n_words,
                                unsigned long long* __restrict__ counters) {
  unsigned long long c = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_words; i += (long long)gridDim.x * blockDim.x)
    c += __popc(bitmap[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(counters + 4, c);
}

}  // namespace

extern "C" size_t gr_metrics_workspace_bytes(int64_t n_items) {
  return gr::align_up((size_t)((n_items + 31) / 32) * 4 + 4, 256);
}

extern "C" int gr_metrics_at_k(const int32_t* recs, int64_t n_users, int32_t k, const int64_t* truth_indptr,
                               const int32_t* truth_ids, int64_t n_items, uint64_t* counters5, void* ws,
                               size_t ws_bytes, gr_stream_t stream) {
  GR_REQUIRE(n_users >= 0 && k >= 1 && n_items >= 0, GR_E_INVALID, "bad shape");
  GR_REQUIRE(counters5 != nullptr, GR_E_INVALID, "null counters");
  GR_REQUIRE(ws != nullptr && ws_bytes >= gr_metrics_workspace_bytes(n_items), GR_E_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n_words = (n_items + 31) / 32;
  GR_CUDA(cudaMemsetAsync(counters5, 0, 5 * sizeof(uint64_t), st));
  GR_CUDA(cudaMemsetAsync(ws, 0, (size_t)n_words * 4 + 4, st));
  if (n_users == 0) return GR_OK;
  GR_REQUIRE(recs && truth_indptr && truth_ids, GR_E_INVALID, "null pointer");
  const int grid = (int)std::min<int64_t>((n_users + 7) / 8, (int64_t)gr::sm_count() * 16);
  metrics_kernel<<<grid, 256, 0, st>>>(recs, n_users, k, reinterpret_cast<const long long*>(truth_indptr), truth_ids,
                                       static_cast<unsigned int*>(ws), reinterpret_cast<unsigned long long*>(counters5));
  GR_LAUNCH_CHECK();
  popcount_kernel<<<std::max(1, (int)std::min<long long>((n_words + 255) / 256, gr::sm_count() * 8)), 256, 0, st>>>(
      static_cast<unsigned int*>(ws), n_words, reinterpret_cast<unsigned long long*>(counters5));
  GR_LAUNCH_CHECK();
  return GR_OK;
}

extern "C" size_t gr_colmean_workspace_bytes(int64_t n, int32_t d) {
  (void)n;
  return gr::align_up((size_t)colmean_grid() * 8 * (size_t)std::max(d, 1) * sizeof(float), 256);
}

extern "C" int gr_colmean_normalized_f32(const float* x, int64_t n, int32_t d, float* center, void* ws,
                                         size_t ws_bytes, gr_stream_t stream) {
  GR_REQUIRE(n >= 0 && d > 0 && d <= 32 * CM_MAXQ, GR_E_INVALID, "d must be in [1, 256]");
  GR_REQUIRE(center != nullptr && (n == 0 || x != nullptr), GR_E_INVALID, "null pointer");
  GR_REQUIRE(ws != nullptr && ws_bytes >= gr_colmean_workspace_bytes(n, d), GR_E_WORKSPACE, "workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = colmean_grid();
  float* partials = static_cast<float*>(ws);
  colmean_partial_kernel<<<grid, 256, 0, st>>>(x, n, d, partials);
  GR_LAUNCH_CHECK();
  colmean_final_kernel<<<(d + 31) / 32, 1024, 0, st>>>(partials, (long long)grid * 8, n, d, center);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

extern "C" int gr_score_prep(const float* x, int64_t n, int32_t d, const float* center_or_null, int32_t d_pad,
                             int32_t parts, int32_t elem_type, uint16_t* out_q, float* stats4_or_null,
                             gr_stream_t stream) {
  GR_REQUIRE(n >= 0 && d > 0, GR_E_INVALID, "bad shape");
  GR_REQUIRE(d_pad == 64 || d_pad == 128 || d_pad == 192 || d_pad == 256, GR_E_INVALID, "d_pad must be 64, 128, 192 or 256");
  GR_REQUIRE(d <= d_pad, GR_E_INVALID, "d exceeds d_pad");
  GR_REQUIRE(parts == 1 || parts == 2, GR_E_INVALID, "parts must be 1 or 2");
  GR_REQUIRE(elem_type == GR_ELEM_BF16 || elem_type == GR_ELEM_FP16, GR_E_INVALID, "unknown element type");
  if (n == 0) return GR_OK;
  GR_REQUIRE(x && out_q, GR_E_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)gr::sm_count() * 16);
  score_prep_kernel<<<grid, 256, 0, st>>>(x, n, d, center_or_null, d_pad, parts, elem_type, out_q, stats4_or_null);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

extern "C" int gr_rescore_topk_f32(const float* h_user, const float* h_item, int64_t item_id_base, int32_t d,
                                   const float* center_or_null, const float* sl_score, const int32_t* sl_id,
                                   int32_t shortlist, int64_t n_users, const float* item_stats4, int32_t elem_type,
                                   int32_t parts_users, int32_t parts_items, float acc_err, const float* band_or_null,
                                   float tie_tol, int32_t k, float eps, const int32_t* user_map_or_null,
                                   int32_t* out_ids, float* out_scores, int32_t* overflow_users, int32_t* n_overflow,
                                   gr_stream_t stream) {
  GR_REQUIRE(n_users >= 0 && d > 0 && d <= 4096, GR_E_INVALID, "bad shape");
  GR_REQUIRE(shortlist >= 1 && shortlist <= 32, GR_E_INVALID, "shortlist must be in [1, 32]");
  GR_REQUIRE(k >= 1 && k <= 32, GR_E_INVALID, "k must be in [1, 32]");
  GR_REQUIRE(elem_type == GR_ELEM_BF16 || elem_type == GR_ELEM_FP16, GR_E_INVALID, "unknown element type");
  GR_REQUIRE((parts_users == 1 || parts_users == 2) && (parts_items == 1 || parts_items == 2), GR_E_INVALID, "parts must be 1 or 2");
  if (n_users == 0) return GR_OK;
  GR_REQUIRE(h_user && h_item && sl_score && sl_id && item_stats4 && out_ids && out_scores && overflow_users && n_overflow,
             GR_E_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<int64_t>((n_users + RS_WARPS - 1) / RS_WARPS, (int64_t)gr::sm_count() * 8);
  const size_t smem = sizeof(float) * RS_WARPS * d;
  if (smem > 48 * 1024) GR_CUDA(cudaFuncSetAttribute(rescore_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  RescoreErr em{elem_type, parts_users, parts_items, acc_err, band_or_null};
  rescore_kernel<<<grid, RS_WARPS * 32, smem, st>>>(h_user, h_item, item_id_base, d, center_or_null, sl_score, sl_id,
                                                    shortlist, n_users, item_stats4, em, tie_tol, k, eps,
                                                    user_map_or_null, out_ids, out_scores, overflow_users, n_overflow);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

extern "C" int gr_score_band(const float* item_stats4, const float* user_stats4, int32_t elem_type, int32_t parts_users,
                             int32_t parts_items, float acc_err, float* band, gr_stream_t stream) {
  GR_REQUIRE(item_stats4 && user_stats4 && band, GR_E_INVALID, "null pointer");
  GR_REQUIRE(elem_type == GR_ELEM_BF16 || elem_type == GR_ELEM_FP16, GR_E_INVALID, "unknown element type");
  RescoreErr em{elem_type, parts_users, parts_items, acc_err, nullptr};
  score_band_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(item_stats4, user_stats4, em, band);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

extern "C" int gr_score_topk_exact_f32(const float* h_user, const int32_t* user_list_or_null,
                                       const int32_t* n_list_or_null, int64_t n_users, const float* h_item,
                                       int64_t n_items, int64_t item_id_base, int32_t d,
                                       const int64_t* bought_indptr_or_null, const int32_t* bought_ids_or_null,
                                       int32_t k, float eps, const float* popularity_or_null, float weight_popularity,
                                       int32_t* out_ids, float* out_scores, gr_stream_t stream) {
  GR_REQUIRE(n_users >= 0 && n_items >= 0 && d > 0 && d <= 4096, GR_E_INVALID, "bad shape");
  GR_REQUIRE(k >= 1, GR_E_INVALID, "k must be positive");
  GR_REQUIRE(item_id_base >= 0 && item_id_base + n_items <= 0x7fffffffLL, GR_E_INVALID, "item ids must fit int32");
  if (n_users == 0) return GR_OK;
  GR_REQUIRE(h_user && out_ids && out_scores && (n_items == 0 || h_item), GR_E_INVALID, "null pointer");
  GR_REQUIRE(bought_indptr_or_null == nullptr || bought_ids_or_null != nullptr, GR_E_INVALID,
             "bought_indptr without bought_ids");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int k_first = std::min<int>(k, EX_KMAX);
  const size_t smem = sizeof(float) * ((d + 3) & ~3) + (size_t)k_first * EX_THREADS * 8;
  GR_CUDA(cudaFuncSetAttribute(exact_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<int64_t>(n_users, (int64_t)gr::sm_count() * 8);
  for (int k_off = 0; k_off < k; k_off += EX_KMAX) {  // one sweep over the items per EX_KMAX output entries
    const int k_pass = std::min<int>(EX_KMAX, k - k_off);
    exact_topk_kernel<<<grid, EX_THREADS, smem, st>>>(
        h_user, user_list_or_null, n_list_or_null, n_users, h_item, n_items, item_id_base, d,
        reinterpret_cast<const long long*>(bought_indptr_or_null), bought_ids_or_null, k, k_off, k_pass, eps,
        popularity_or_null, weight_popularity, out_ids, out_scores);
    GR_LAUNCH_CHECK();
  }
  return GR_OK;
}

extern "C" int gr_topk_merge(const float* scores, const int32_t* ids, int32_t parts, int64_t n_users, int32_t k_in,
                             int32_t k_out, float* out_scores, int32_t* out_ids, gr_stream_t stream) {
  GR_REQUIRE(parts >= 1 && parts <= 32, GR_E_INVALID, "parts must be in [1, 32]");
  GR_REQUIRE(n_users >= 0 && k_in >= 1 && k_out >= 1, GR_E_INVALID, "bad shape");
  if (n_users == 0) return GR_OK;
  GR_REQUIRE(scores && ids && out_scores && out_ids, GR_E_INVALID, "null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = (int)std::min<int64_t>((n_users + 7) / 8, (int64_t)gr::sm_count() * 16);
  topk_merge_kernel<<<grid, 256, 0, st>>>(scores, ids, parts, n_users, k_in, k_out, out_scores, out_ids);
  GR_LAUNCH_CHECK();
  return GR_OK;
}
