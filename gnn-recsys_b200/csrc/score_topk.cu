// gr_score_topk_tc: all-users x all-items scoring with a fused running shortlist (stage 1 of get_recs).
//
// Replaces the reference's per-user loop (src/metrics.py:52-77): torch.cat repeat of the user row, one
// nn.CosineSimilarity call over all items, a D2H copy of I scores, np.argsort and a Python already-bought filter --
// with ONE dense contraction users[U, D] x items[I, D]^T on the 5th-generation tensor cores:
//
//   * operands: 16-bit (fp16 or bf16), K-major, 128-byte-swizzled shared-memory tiles written by TMA
//     (cp.async.bulk.tensor). Each side carries 1 or 2 parts per row (hi, or hi + lo with x = hi + lo up to 2^-22 /
//     2^-18 relative); the product scheme follows from (parts_users, parts_items):
//       (1, 1)  hi.hi                    1 product   -- the default first pass (fp16, shortlist 32)
//       (2, 1)  hi.hi + lo.hi            2 products  -- user row exact, item row rounded once
//       (2, 2)  hi.hi + lo.hi + hi.lo    3 products  -- fp32-grade, the second pass for users the first cannot prove
//     all accumulated in the same TMEM tile. Whatever the scheme, the ANSWER comes from gr_rescore_topk_f32, which
//     re-scores the shortlist in exact fp32 and proves per user that nothing outside it can reach the top-k.
//   * math: tcgen05.mma.cta_group::2.kind::f16, M = 256 users (128 from each CTA of a pair) x N = 128 items x K = 16,
//     fp32 accumulators in TMEM. A CTA owns 256 users (two 128-row A tiles per part, resident in shared memory for the
//     whole sweep) and walks its item range once; of every 128-item B sub-tile (64 K-elements of one part) a CTA
//     TMA-loads and holds only ITS 64 rows (8 KB slots in a ring). The four 128-column accumulators ([user tile] x
//     [double buffer]) fill all 512 TMEM columns, so the epilogue of tile j overlaps the MMAs of tile j + 1.
//     A single-CTA variant (cta_group::1) is kept behind GR_SCORE_FLAG_SINGLE_CTA and for d_pad = 64.
//   * warp roles (352 threads): warp 0 = TMA producer; warps 1 and 10 = MMA issuers, one per user tile (+ TMEM
//     alloc); warps 2-9 = epilogue. An epilogue thread owns ONE user row (TMEM lane) for the whole sweep: four
//     tcgen05.ld.32x32b.x32 drain the row's 128 scores into registers, the TMEM slot is released immediately, then a
//     3-input max tree (FMNMX3) and one compare per 32 scores against the row's running threshold. Only scores that
//     beat it take the (compact, out-of-line) slow path: already-bought test against the user's sorted id list and
//     insertion into the row's sorted shortlist ([S][256] in shared memory).
//   * threshold: tau = max(S-th best so far, k-th best so far - band). The second term drops candidates that can
//     no longer reach the exact top-k (band >= 2 x the approximation error, a device scalar written by the host side
//     from the operand statistics), which halves the slow-path traffic of a 32-entry shortlist; the re-score kernel
//     accounts for both kinds of dropped items in its proof.
//   * scores are never materialised (10M x 1M would be 40 TB).
//   * small user counts: the item range is split over blockIdx.y so the grid still fills the chip; the per-split
//     shortlists are merged by gr_topk_merge (host side of this file).

#include "tc5.cuh"

namespace {

using namespace gr::tc5;

constexpr int TILE_M = 128;    // users per MMA = TMEM lanes
constexpr int UT = 2;          // user tiles per CTA
constexpr int TILE_N = 128;    // items per tile = accumulator columns
constexpr int SUB_BYTES = TILE_M * KBLK * 2;  // one [128 rows][64 x 16 bit] swizzled sub-tile = 16 KB
constexpr int ROWS_PER_CTA = TILE_M * UT;     // 256
constexpr int EPI_WARPS = 4 * UT;
constexpr int MMA_WARP1 = 2 + EPI_WARPS;              // second MMA issuer (user tile 1); warp 1 issues user tile 0
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS + 32;  // 352
constexpr int TMEM_COLS = 512;

// kind::f16 instruction descriptor: D = fp32, A/B = fp16 (0) or bf16 (1), both K-major, N = 128, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t ab_format, uint32_t m = TILE_M) {
  return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((uint32_t)(TILE_N >> 3) << 17) | ((m >> 4) << 24);
}

// v[i] for a run-time i over a register array: binary select tree (N - 1 selects), no local memory
template <int N>
__device__ __forceinline__ uint32_t select_n(const uint32_t* v, int i) {
  if constexpr (N == 1) {
    return v[0];
  } else {
    const uint32_t lo = select_n<N / 2>(v, i), hi = select_n<N / 2>(v + N / 2, i);
    return (i & (N / 2)) ? hi : lo;
  }
}
__device__ __forceinline__ uint32_t select32(const uint32_t* v, int i) { return select_n<32>(v, i); }

// ---- slow path -------------------------------------------------------------------------------------------------
// Shortlists live in shared memory ROW-major: row t of the CTA owns ls / li [t * lst .. t * lst + S), lst = S | 1 (odd
// pitch: conflict-free both for a thread walking its own row and for a warp reading one row with one entry per lane).
//
// bought_test: is `gid` one of this row's already-bought items (bought_ids[b0, b1), ascending)? `sig` is a 64-bit
// membership signature of the row's list built once per sweep (bit hash(id) set for every bought id): a clear bit
// answers "no" without touching memory, whatever order the ids arrive in (the item sweep may be permuted, see
// item_perm); only a set bit pays for the binary search.
__device__ __forceinline__ uint32_t bought_hash(int gid) { return ((uint32_t)gid * 0x9E3779B1u) >> 26; }
__device__ __noinline__ bool bought_test(int gid, uint64_t sig, long long b0, long long b1,
                                         const int* __restrict__ bought_ids) {
  if (((sig >> bought_hash(gid)) & 1ull) == 0) return false;
  long long lo = b0, hi = b1;  // first bought id >= gid
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (__ldg(bought_ids + mid) < gid) lo = mid + 1; else hi = mid;
  }
  return lo < b1 && __ldg(bought_ids + lo) == gid;
}

// Warp-cooperative sorted insert of candidate (s, gid) into ONE row's shortlist (rl / ri, scores descending): lane j
// holds entry j, a ballot finds the insert position, one shuffle shifts the tail, every lane stores its own slot -- a
// dozen instructions, no divergence and no dependent shared-memory chain (the per-thread shift loop this replaces cost
// ~1000 cycles per insert and stalled the whole warp). Equal scores keep the earlier (smaller) id first. Lane j only
// ever touches slot j, so back-to-back inserts need no barrier. Returns the row's new threshold
// max(S-th best, k-th best - band) in every lane.
__device__ __noinline__ float coop_insert(float s, int gid, float* __restrict__ rl, int* __restrict__ ri, int S, int kk,
                                          float band, int lane) {
  const float e = lane < S ? rl[lane] : -INFINITY;
  const int ei = lane < S ? ri[lane] : -1;
  const int pos = __popc(__ballot_sync(0xffffffffu, e >= s));  // sorted descending: the lanes with e >= s are a prefix
  const float up = __shfl_up_sync(0xffffffffu, e, 1);
  const int upi = __shfl_up_sync(0xffffffffu, ei, 1);
  const float ne = lane < pos ? e : (lane == pos ? s : up);
  const int nei = lane < pos ? ei : (lane == pos ? gid : upi);
  if (lane >= pos && lane < S) { rl[lane] = ne; ri[lane] = nei; }
  return fmaxf(__shfl_sync(0xffffffffu, ne, S - 1), __shfl_sync(0xffffffffu, ne, kk - 1) - band);
}

// Four inserts into four DIFFERENT rows at once: the same instruction sequence, interleaved, so that the ~100-cycle
// latency chain (LDS -> vote -> SHFL -> STS -> SHFL) of one insert hides behind the others. s.? == -inf marks an unused
// slot (every lane compares >= it: position 32, nothing is stored). row.? = offset of the row in ls / li.
__device__ __noinline__ float4 coop_insert4(float4 s4, int4 g4, int4 r4, float* __restrict__ ls, int* __restrict__ li,
                                            int S, int kk, float band, int lane) {
  const float s[4] = {s4.x, s4.y, s4.z, s4.w};
  const int g[4] = {g4.x, g4.y, g4.z, g4.w};
  const int r[4] = {r4.x, r4.y, r4.z, r4.w};
  float e[4], ne[4], nt[4];
  int ei[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    e[i] = lane < S ? ls[r[i] + lane] : -INFINITY;
    ei[i] = lane < S ? li[r[i] + lane] : -1;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pos = __popc(__ballot_sync(0xffffffffu, e[i] >= s[i]));
    const float up = __shfl_up_sync(0xffffffffu, e[i], 1);
    const int upi = __shfl_up_sync(0xffffffffu, ei[i], 1);
    ne[i] = lane < pos ? e[i] : (lane == pos ? s[i] : up);
    const int nei = lane < pos ? ei[i] : (lane == pos ? g[i] : upi);
    if (lane >= pos && lane < S) { ls[r[i] + lane] = ne[i]; li[r[i] + lane] = nei; }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
    nt[i] = fmaxf(__shfl_sync(0xffffffffu, ne[i], S - 1), __shfl_sync(0xffffffffu, ne[i], kk - 1) - band);
  return make_float4(nt[0], nt[1], nt[2], nt[3]);
}

// Shared-memory plan: [A: UT x PA x KB sub-tiles][B ring: `ring` slots][shortlists][barriers]
constexpr int MAX_RING = 8;
template <int KB, int PA, bool PAIR = false>
struct Cfg {
  static constexpr int A_BYTES = UT * PA * KB * SUB_BYTES;
  static constexpr int SLOT_BYTES = PAIR ? SUB_BYTES / 2 : SUB_BYTES;  // a CTA of a pair holds half of every B sub-tile
  static constexpr int TAIL_BYTES = 512;  // barriers + TMEM slot
  static int list_bytes(int S) { return (S | 1) * ROWS_PER_CTA * 8; }  // row pitch S | 1
  static int ring(int S) {
    const int r = (SMEM_LIMIT - 1024 - A_BYTES - list_bytes(S) - TAIL_BYTES) / SLOT_BYTES;
    return r > MAX_RING ? MAX_RING : r;
  }
  static size_t smem(int S) { return 1024 /*alignment slack*/ + A_BYTES + (size_t)ring(S) * SLOT_BYTES + list_bytes(S) + TAIL_BYTES; }
};

// PA / PB: parts per user / item row (1 = hi, 2 = hi + lo). Item part 0 (hi) pairs with every user part, item part 1
// (lo) with the user hi part only: (1,1) hi.hi; (2,1) hi.hi + lo.hi; (2,2) hi.hi + lo.hi + hi.lo.
// PAIR: two CTAs of a cluster work as one cta_group::2 unit -- M = 256 MMAs (128 user rows from each CTA), each CTA
// TMA-loads and holds only HALF of every item sub-tile (64 rows), so the shared-memory and L2 traffic per FLOP halve.
// The even CTA (leader) issues every MMA; TMA completions of both CTAs land on the leader's barriers, MMA commits
// are multicast to both CTAs, epilogues of both CTAs release the accumulators on the leader's barriers.
template <int KB, int PA, int PB, bool PAIR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap tm_users, const __grid_constant__ CUtensorMap tm_items,
                  long long n_users, long long n_items, long long item_id_base, const int* __restrict__ item_perm,
                  int tiles_per_split, uint32_t idesc, const long long* __restrict__ bought_indptr, const int* __restrict__ bought_ids, int S, int kk,
                  const float* __restrict__ band_ptr, const int* __restrict__ user_map, int ring,
                  float* __restrict__ sl_score, int* __restrict__ sl_id, int full_pairs, int tail_c,
                  float* __restrict__ tail_score, int* __restrict__ tail_id) {
  using L = Cfg<KB, PA, PAIR>;
  constexpr int D_PAD = KB * KBLK;
  constexpr int SLOT = L::SLOT_BYTES;
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;
  uint8_t* sB = sA + L::A_BYTES;
  const int lst = S | 1;                                          // shortlist row pitch
  float* ls = reinterpret_cast<float*>(sB + ring * SLOT);  // [256][lst]
  int* li = reinterpret_cast<int*>(ls + lst * ROWS_PER_CTA);     // [256][lst]
  uint64_t* bars = reinterpret_cast<uint64_t*>(
      (reinterpret_cast<uintptr_t>(li + lst * ROWS_PER_CTA) + 7) & ~(uintptr_t)7);  // lst is odd: re-align to 8 bytes
  uint64_t* full = bars;                  // [MAX_RING]  TMA -> MMA
  uint64_t* empty = full + MAX_RING;      // [MAX_RING]  MMA -> TMA
  uint64_t* a_full = empty + MAX_RING;    // [1]
  uint64_t* t_full = a_full + 1;          // [UT][2]     MMA -> epilogue
  uint64_t* t_empty = t_full + UT * 2;    // [UT][2]     epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + UT * 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler as well
  const int lane = threadIdx.x & 31;
  const int n_tiles_all = (int)((n_items + TILE_N - 1) / TILE_N);
  // Work unit of this CTA (pair): normally users [blockIdx.x * 256, +256) x item split blockIdx.y. TAIL SPLIT (pairs,
  // one split): the pairs of the last, partial wave -- index >= full_pairs -- are cut into tail_c item ranges each, so
  // that the wave is tail_c times shorter instead of leaving most SM pairs idle for a whole sweep; their shortlists go
  // to tail_score / tail_id ([range][tail row][S]) and are merged by the host side.
  int split = blockIdx.y, per_split = tiles_per_split;
  long long row_base = (long long)blockIdx.x * ROWS_PER_CTA;
  const bool tail = PAIR && tail_c > 1 && (int)(blockIdx.x >> 1) >= full_pairs;
  if (tail) {
    const int qq = (int)(blockIdx.x >> 1) - full_pairs;
    row_base = ((long long)(full_pairs + qq / tail_c) * 2 + (blockIdx.x & 1)) * ROWS_PER_CTA;
    split = qq % tail_c;
    per_split = (n_tiles_all + tail_c - 1) / tail_c;
  }
  const int tile0 = split * per_split;
  const int n_tiles = max(0, min(per_split, n_tiles_all - tile0));

  if (threadIdx.x == 0) {
    for (int i = 0; i < ring; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, UT); }  // one commit per issuer
    mbar_init(a_full, 1);
    for (int i = 0; i < UT * 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, PAIR ? 8 : 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM allocation: all 512 columns (this kernel runs one CTA per SM)
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / TMA completion
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: one sub-tile (64 K-elements of one part of one item tile) per ring slot
    if (lane == 0 && n_tiles > 0) {
      if (leader) mbar_expect_tx(a_full, (PAIR ? 2 : 1) * L::A_BYTES);
      for (int ut = 0; ut < UT; ++ut)
        for (int pa = 0; pa < PA; ++pa)
          for (int kb = 0; kb < KB; ++kb) {
            void* dst = sA + ((ut * PA + pa) * KB + kb) * SUB_BYTES;
            if (PAIR) tma_load_2d_pair(dst, &tm_users, pa * D_PAD + kb * KBLK, (int)(row_base + ut * TILE_M), a_full);
            else tma_load_2d(dst, &tm_users, pa * D_PAD + kb * KBLK, (int)(row_base + ut * TILE_M), a_full);
          }
      int buf = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        for (int pb = 0; pb < PB; ++pb) {
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(empty + buf, phase ^ 1u);
            if (PAIR) {  // this CTA's 64 rows of the 128-item sub-tile; the bytes of both halves count on the leader
              if (leader) mbar_expect_tx(full + buf, SUB_BYTES);
              tma_load_2d_pair(sB + buf * SLOT, &tm_items, pb * D_PAD + kb * KBLK,
                               (tile0 + j) * TILE_N + (int)crank * (TILE_N / 2), full + buf);
            } else {
              mbar_expect_tx(full + buf, SUB_BYTES);
              tma_load_2d(sB + buf * SLOT, &tm_items, pb * D_PAD + kb * KBLK, (tile0 + j) * TILE_N, full + buf);
            }
            if (++buf == ring) { buf = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1 || warp == MMA_WARP1) {
    // ================= MMA issuers (leader CTA), one warp per user tile. Why two warps: tcgen05.mma issue BLOCKS the
    // issuing warp on the tensor pipe's queue (tools/exp_mma_issue.cu: one warp sustains 1 MMA per 85 cycles, the pipe
    // wants 1 per 64; two warps reach 64). The WHOLE warp walks the loop (barrier waits included) and one elected lane
    // issues: everything the MMAs consume is warp-uniform (descriptors from uniform shared-memory offsets, the TMEM
    // base broadcast once), so UTCHMMA is fed from uniform registers -- four instructions for four MMAs. (With
    // `if (lane == 0)` every MMA went through an 18-instruction R2UR waterfall.) =================
    const int ut = warp == 1 ? 0 : 1;
    if (n_tiles > 0 && leader) {
      // The WHOLE warp walks the loop (barrier waits included); one elected lane issues. Everything the MMAs consume
      // is warp-uniform: descriptors are built from uniform shared-memory offsets, the TMEM base is broadcast once.
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
      mbar_wait(a_full, 0);
      tc_fence_after();
      int buf = 0;
      uint32_t phase = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const int slot = j & 1;
        const uint32_t aphase = (uint32_t)(j >> 1) & 1u;
        const uint32_t d_tmem = tmem_u + (uint32_t)((ut * 2 + slot) * TILE_N);
        for (int pb = 0; pb < PB; ++pb) {
#pragma unroll
          for (int kb = 0; kb < KB; ++kb) {
            mbar_wait(full + buf, phase);
            if (pb == 0 && kb == 0) mbar_wait(t_empty + ut * 2 + slot, aphase ^ 1u);  // drained by the epilogue of tile j - 2
            tc_fence_after();
            const uint64_t db = make_desc_sw128(sB_u + (uint32_t)(buf * SLOT));
            if (elect_one()) {
              const int n_pa = pb == 0 ? PA : 1;
              for (int pa = 0; pa < n_pa; ++pa) {
                const uint64_t da = make_desc_sw128(sA_u + (uint32_t)(((ut * PA + pa) * KB + kb) * SUB_BYTES));
#pragma unroll
                for (int k = 0; k < KBLK / UMMA_K; ++k) {  // +32 bytes (>>4 = 2) per K = 16 step inside the swizzle atom
                  if (PAIR) tc_mma_f16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                            (pb | kb | pa | k) != 0 ? 1u : 0u);
                  else tc_mma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                  (pb | kb | pa | k) != 0 ? 1u : 0u);
                }
              }
              if (pb == PB - 1 && kb == KB - 1) {
                if (PAIR) tc_commit_pair(t_full + ut * 2 + slot); else tc_commit(t_full + ut * 2 + slot);
              }
              if (PAIR) tc_commit_pair(empty + buf); else tc_commit(empty + buf);
            }
            __syncwarp();
            if (++buf == ring) { buf = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ================= epilogue: one thread = one user row =================
    const int ew = warp - 2;
    const int ut = ew >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int t = ut * TILE_M + q * 32 + lane;
    const long long row = row_base + t;
    const bool live = row < n_users;
    const float band = band_ptr != nullptr ? __ldg(band_ptr) : INFINITY;  // +inf: plain S-th-best threshold
    const int t0 = t - lane;  // first CTA row of this warp
    for (int s = 0; s < S; ++s) { ls[t * lst + s] = -INFINITY; li[t * lst + s] = -1; }
    long long b0 = 0, b1 = 0;
    if (live && bought_indptr != nullptr) {  // bought lists are indexed by the ORIGINAL user row
      const long long brow = user_map != nullptr ? (long long)user_map[row] : row;
      b0 = bought_indptr[brow]; b1 = bought_indptr[brow + 1];
    }
    uint64_t sig = 0;  // membership signature of this row's bought list (see bought_test)
    for (long long p = b0; p < b1; ++p) sig |= 1ull << bought_hash(__ldg(bought_ids + p));
    float tau = live ? -INFINITY : INFINITY;
    const int n_pos = (int)n_items;  // sweep positions [0, n_items): position p is item item_perm[p] (p itself without a permutation)
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int slot = j & 1;
      const uint32_t aphase = (uint32_t)(j >> 1) & 1u;
      const bool partial = (long long)(tile0 + j + 1) * TILE_N > n_items;  // last tile: columns past n_items are zero fill
      mbar_wait(t_full + ut * 2 + slot, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ut * 2 + slot) * TILE_N);
      // drain the whole accumulator row into registers, then hand the TMEM slot straight back to the MMA warp:
      // the scan below (and its data-dependent slow path) never holds up the tensor pipe
      uint32_t v[TILE_N];
#pragma unroll
      for (int c = 0; c < TILE_N / 32; ++c) tc_ld32(taddr + (uint32_t)(c * 32), v + c * 32);
      tc_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_leader(t_empty + ut * 2 + slot); else mbar_arrive(t_empty + ut * 2 + slot);
      }
#pragma unroll
      for (int c = 0; c < TILE_N / 32; ++c) {
        float m[11];
#pragma unroll
        for (int i = 0; i < 10; ++i)
          m[i] = max3(__uint_as_float(v[c * 32 + 3 * i]), __uint_as_float(v[c * 32 + 3 * i + 1]),
                      __uint_as_float(v[c * 32 + 3 * i + 2]));
        m[10] = fmaxf(__uint_as_float(v[c * 32 + 30]), __uint_as_float(v[c * 32 + 31]));
        const float mx = max3(max3(m[0], m[1], m[2]), max3(m[3], m[4], m[5]),
                              max3(max3(m[6], m[7], m[8]), m[9], m[10]));
        if (__any_sync(0xffffffffu, mx > tau)) {
          // Slow path, rare after the first tiles and deliberately COMPACT (a bit mask + a select tree instead of one
          // call site per column: it runs cold, so its cost is instruction-cache lines, not instructions). The whole
          // warp enters: every lane pops its own next candidate (bought test included), then the pending candidates
          // are inserted one row at a time by all 32 lanes together.
          const int pos0 = (tile0 + j) * TILE_N + c * 32;
          uint32_t cand = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (__uint_as_float(v[c * 32 + i]) > tau) cand |= 1u << i;
          if (partial) cand &= (n_pos - pos0 >= 32) ? 0xffffffffu : (n_pos > pos0 ? (1u << (n_pos - pos0)) - 1u : 0u);
          for (;;) {
            float s = 0.f;
            int gid = -1;
            while (cand != 0) {
              const int i = __ffs(cand) - 1;
              cand &= cand - 1;
              const float sv = __uint_as_float(select32(v + c * 32, i));
              if (sv > tau) {
                const int g = (int)item_id_base + (item_perm != nullptr ? __ldg(item_perm + pos0 + i) : pos0 + i);
                if (!bought_test(g, sig, b0, b1, bought_ids)) { s = sv; gid = g; break; }
              }
            }
            uint32_t pend = __ballot_sync(0xffffffffu, gid >= 0);
            if (pend == 0) break;
            while (pend != 0) {
              const int L0 = __ffs(pend) - 1;
              pend &= pend - 1;
              if (pend == 0) {  // a single pending row
                const float nt = coop_insert(__shfl_sync(0xffffffffu, s, L0), __shfl_sync(0xffffffffu, gid, L0),
                                             ls + (t0 + L0) * lst, li + (t0 + L0) * lst, S, kk, band, lane);
                if (lane == L0) tau = nt;
              } else {          // up to four rows per call, interleaved
                int L1 = __ffs(pend) - 1, L2 = -1, L3 = -1;
                pend &= pend - 1;
                if (pend != 0) { L2 = __ffs(pend) - 1; pend &= pend - 1; }
                if (pend != 0) { L3 = __ffs(pend) - 1; pend &= pend - 1; }
                float4 s4;
                int4 g4, r4;
                s4.x = __shfl_sync(0xffffffffu, s, L0); g4.x = __shfl_sync(0xffffffffu, gid, L0); r4.x = (t0 + L0) * lst;
                s4.y = __shfl_sync(0xffffffffu, s, L1); g4.y = __shfl_sync(0xffffffffu, gid, L1); r4.y = (t0 + L1) * lst;
                s4.z = L2 >= 0 ? __shfl_sync(0xffffffffu, s, L2 & 31) : -INFINITY;
                g4.z = __shfl_sync(0xffffffffu, gid, L2 & 31); r4.z = (t0 + (L2 & 31)) * lst;
                s4.w = L3 >= 0 ? __shfl_sync(0xffffffffu, s, L3 & 31) : -INFINITY;
                g4.w = __shfl_sync(0xffffffffu, gid, L3 & 31); r4.w = (t0 + (L3 & 31)) * lst;
                const float4 nt = coop_insert4(s4, g4, r4, ls, li, S, kk, band, lane);
                if (lane == L0) tau = nt.x;
                if (lane == L1) tau = nt.y;
                if (lane == L2) tau = nt.z;
                if (lane == L3) tau = nt.w;
              }
            }
          }
        }
      }
    }
    __syncwarp();
    for (int L = 0; L < 32; ++L) {  // [split][row][S]: one coalesced row per iteration
      const long long r = row_base + t0 + L;
      if (r < n_users && lane < S) {
        if (tail) {
          const long long tail_row0 = (long long)full_pairs * 2 * ROWS_PER_CTA;
          const long long o = ((long long)split * (n_users - tail_row0) + (r - tail_row0)) * S + lane;
          tail_score[o] = ls[(t0 + L) * lst + lane];
          tail_id[o] = li[(t0 + L) * lst + lane];
        } else {
          sl_score[((long long)split * n_users + r) * S + lane] = ls[(t0 + L) * lst + lane];
          sl_id[((long long)split * n_users + r) * S + lane] = li[(t0 + L) * lst + lane];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // the peer may still be reading its half of the accumulators / our barriers
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

__global__ void fill_empty_shortlist_kernel(float* sl_score, int* sl_id, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    sl_score[i] = -INFINITY;
    sl_id[i] = -1;
  }
}

struct ScoreArgs {
  CUtensorMap mu, mi, mi_half;  // mi_half: 64-row boxes, the half sub-tile a CTA of a pair loads
  long long n_users, n_items, item_id_base;
  const int* perm;          // sweep position -> item index (nullptr: identity)
  int splits, tiles_per_split;
  uint32_t ab_format;
  const long long* bptr;
  const int* bids;
  int S, k;
  const float* band;
  const int* user_map;
  float* sl_score;
  int* sl_id;
  int full_pairs, tail_c;   // tail split (see the kernel): 0 / 1 = off
  float* tail_score;
  int* tail_id;
};

// Tail split: with `pairs` CTA pairs on `slots` SM pairs the last wave holds pairs % slots of them. When that is at most
// half the slots, each of its pairs is cut into c item ranges (all of them still resident at once).
struct TailPlan { int full_pairs, tail_pairs, c; };
TailPlan plan_tail(long long n_users, long long n_items) {
  const long long pairs = ((n_users + ROWS_PER_CTA - 1) / ROWS_PER_CTA + 1) / 2;
  const int slots = gr::sm_count() / 2;
  const long long tiles = (n_items + TILE_N - 1) / TILE_N;
  TailPlan p{0, 0, 1};
  if (slots <= 0 || pairs <= slots) return p;
  const int tail = (int)(pairs % slots);
  if (tail == 0 || 2 * tail > slots) return p;
  const int c = (int)std::min<long long>(std::min<long long>(slots / tail, 8), tiles / 64);  // >= 64 tiles per range
  if (c < 2) return p;
  p.full_pairs = (int)(pairs - tail); p.tail_pairs = tail; p.c = c;
  return p;
}
size_t tail_half_bytes(const TailPlan& p, long long n_users, int S) {
  if (p.c < 2) return 0;
  const long long tail_rows = n_users - (long long)p.full_pairs * 2 * ROWS_PER_CTA;
  return gr::align_up((size_t)p.c * tail_rows * S * 4, 256);
}

template <int KB, int PA, int PB, bool PAIR>
int launch_score(const ScoreArgs& a, cudaStream_t st) {
  auto kern = score_topk_kernel<KB, PA, PB, PAIR>;
  using L = Cfg<KB, PA, PAIR>;
  const int ring = L::ring(a.S);
  GR_REQUIRE(ring >= 2, GR_E_INVALID, "shortlist too large for the shared-memory budget of this configuration");
  const size_t smem = L::smem(a.S);
  GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  unsigned gx = (unsigned)((a.n_users + ROWS_PER_CTA - 1) / ROWS_PER_CTA);
  const uint32_t idesc = PAIR ? make_idesc(a.ab_format, 2 * TILE_M) : make_idesc(a.ab_format);
  if (PAIR) {
    gx = (gx + 1) & ~1u;  // whole CTA pairs; the rows of a surplus CTA are out of range (TMA zero fill, not stored)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, (unsigned)a.splits);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (a.tail_c > 1) cfg.gridDim = dim3(2u * (unsigned)(a.full_pairs + (gx / 2 - a.full_pairs) * a.tail_c), 1u);
    GR_CUDA(cudaLaunchKernelEx(&cfg, kern, a.mu, a.mi_half, a.n_users, a.n_items, a.item_id_base, a.perm,
                               a.tiles_per_split, idesc, a.bptr, a.bids, a.S, a.k, a.band, a.user_map, ring, a.sl_score, a.sl_id,
                               a.full_pairs, a.tail_c, a.tail_score, a.tail_id));
  } else {
    kern<<<dim3(gx, (unsigned)a.splits), NUM_THREADS, smem, st>>>(a.mu, a.mi, a.n_users, a.n_items, a.item_id_base,
                                                                  a.perm, a.tiles_per_split, idesc, a.bptr, a.bids, a.S, a.k,
                                                                  a.band, a.user_map, ring, a.sl_score, a.sl_id, 0, 1,
                                                                  nullptr, nullptr);
  }
  GR_LAUNCH_CHECK();
  return GR_OK;
}

template <int KB, bool PAIR>
int launch_scheme(const ScoreArgs& a, int pa, int pb, cudaStream_t st) {
  if (pa == 1 && pb == 1) return launch_score<KB, 1, 1, PAIR>(a, st);
  if (pa == 2 && pb == 1) return launch_score<KB, 2, 1, PAIR>(a, st);
  return launch_score<KB, 2, 2, PAIR>(a, st);
}

// item-range splits: enough CTAs for two waves when the user count alone cannot fill the chip. (Splitting further
// to shorten a nearly empty last wave was measured and rejected: every split restarts its shortlists, and the
// warm-up insertions cost more than the wave round-up saves -- 250k x 200k: 29.6 ms unsplit, 39.5 ms in 3 splits.)
int choose_splits(long long n_users, long long n_items) {
  long long ctas = (n_users + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  ctas += ctas & 1;  // CTA pairs
  const long long tiles = (n_items + TILE_N - 1) / TILE_N;
  const long long want = 2LL * gr::sm_count();
  if (ctas >= want || tiles <= 1) return 1;
  long long sp = (want + ctas - 1) / ctas;
  sp = std::min<long long>(sp, std::min<long long>(tiles, GR_SCORE_MAX_SPLITS));
  return (int)std::max<long long>(sp, 1);
}

}  // namespace

extern "C" int gr_score_splits(int64_t n_users, int64_t n_items) {
  if (n_users <= 0 || n_items <= 0) return 1;
  return choose_splits(n_users, n_items);
}

extern "C" size_t gr_score_topk_workspace_bytes(int64_t n_users, int64_t n_items, int32_t shortlist) {
  if (n_users <= 0 || n_items <= 0) return 256;
  const int splits = choose_splits(n_users, n_items);
  if (splits == 1) return std::max<size_t>(256, 2 * tail_half_bytes(plan_tail(n_users, n_items), n_users, shortlist));
  return gr::align_up((size_t)splits * n_users * shortlist * 4, 256) * 2;
}

extern "C" int gr_score_topk_tc(const uint16_t* users_q, int64_t n_users, const uint16_t* items_q, int64_t n_items,
                                int64_t item_id_base, const int32_t* item_perm_or_null, int32_t d_pad,
                                int32_t parts_users, int32_t parts_items, int32_t elem_type,
                                const int64_t* bought_indptr_or_null,
                                const int32_t* bought_ids_or_null, int32_t shortlist, int32_t k,
                                const float* band_or_null, const int32_t* user_map_or_null, int32_t flags,
                                float* sl_score, int32_t* sl_id, void* ws, size_t ws_bytes, gr_stream_t stream) {
  GR_REQUIRE(n_users >= 0 && n_items >= 0, GR_E_INVALID, "negative size");
  GR_REQUIRE(d_pad == 64 || d_pad == 128 || d_pad == 192 || d_pad == 256, GR_E_INVALID,
             "d_pad must be 64, 128, 192 or 256 (pad the embeddings with gr_score_prep)");
  GR_REQUIRE(d_pad <= 128 || (parts_users == 1 && parts_items == 1), GR_E_INVALID,
             "d_pad 192 / 256 (the reference's Large / Very Large out_dim presets) support the single-product scheme only");
  GR_REQUIRE((parts_users == 1 || parts_users == 2) && (parts_items == 1 || parts_items == 2) && parts_items <= parts_users,
             GR_E_INVALID, "(parts_users, parts_items) must be (1,1), (2,1) or (2,2)");
  GR_REQUIRE(elem_type == GR_ELEM_BF16 || elem_type == GR_ELEM_FP16, GR_E_INVALID, "unknown element type");
  GR_REQUIRE(shortlist >= 1 && shortlist <= 32, GR_E_INVALID, "shortlist must be in [1, 32]");
  GR_REQUIRE(k >= 1 && k <= shortlist, GR_E_INVALID, "k must be in [1, shortlist]");
  GR_REQUIRE(item_id_base >= 0 && item_id_base + n_items <= 0x7fffffffLL, GR_E_INVALID, "item ids must fit int32");
  if (n_users == 0) return GR_OK;
  GR_REQUIRE(sl_score && sl_id, GR_E_INVALID, "null output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_items == 0) {
    fill_empty_shortlist_kernel<<<gr::sm_count() * 4, 256, 0, st>>>(sl_score, sl_id, (long long)n_users * shortlist);
    GR_LAUNCH_CHECK();
    return GR_OK;
  }
  GR_REQUIRE(users_q && items_q, GR_E_INVALID, "null input");
  GR_REQUIRE(bought_indptr_or_null == nullptr || bought_ids_or_null != nullptr, GR_E_INVALID,
             "bought_indptr without bought_ids");
  int major = 0, dev = 0;
  GR_CUDA(cudaGetDevice(&dev));
  GR_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  GR_REQUIRE(major == 10, GR_E_UNSUPPORTED, "tcgen05 scoring kernel needs an sm_100 device");
  ScoreArgs a;
  int rc = make_map(&a.mu, users_q, n_users, d_pad * parts_users, elem_type, TILE_M);
  if (rc != GR_OK) return rc;
  rc = make_map(&a.mi, items_q, n_items, d_pad * parts_items, elem_type, TILE_M);
  if (rc != GR_OK) return rc;
  rc = make_map(&a.mi_half, items_q, n_items, d_pad * parts_items, elem_type, TILE_N / 2);
  if (rc != GR_OK) return rc;
  a.n_users = n_users; a.n_items = n_items; a.item_id_base = item_id_base; a.perm = item_perm_or_null;
  a.splits = choose_splits(n_users, n_items);
  const int tiles = (int)((n_items + TILE_N - 1) / TILE_N);
  a.tiles_per_split = (tiles + a.splits - 1) / a.splits;
  a.ab_format = elem_type == GR_ELEM_FP16 ? 0u : 1u;
  a.bptr = reinterpret_cast<const long long*>(bought_indptr_or_null);
  a.bids = bought_ids_or_null;
  a.S = shortlist; a.k = k; a.band = band_or_null; a.user_map = user_map_or_null;
  float* part_score = sl_score;
  int* part_id = sl_id;
  if (a.splits > 1) {
    const size_t half = gr::align_up((size_t)a.splits * n_users * shortlist * 4, 256);
    GR_REQUIRE(ws != nullptr && ws_bytes >= 2 * half, GR_E_WORKSPACE, "workspace too small");
    part_score = static_cast<float*>(ws);
    part_id = reinterpret_cast<int*>(static_cast<char*>(ws) + half);
  }
  a.sl_score = part_score; a.sl_id = part_id;
  a.full_pairs = 0; a.tail_c = 1; a.tail_score = nullptr; a.tail_id = nullptr;
  TailPlan tp{0, 0, 1};
  const bool pair_kernel = d_pad >= 128 && !(d_pad == 128 && (flags & GR_SCORE_FLAG_SINGLE_CTA));
  if (a.splits == 1 && pair_kernel && !(flags & GR_SCORE_FLAG_NO_TAIL_SPLIT)) {
    tp = plan_tail(n_users, n_items);
    const size_t half = tail_half_bytes(tp, n_users, shortlist);
    if (tp.c > 1 && ws != nullptr && ws_bytes >= 2 * half) {
      a.full_pairs = tp.full_pairs; a.tail_c = tp.c;
      a.tail_score = static_cast<float*>(ws);
      a.tail_id = reinterpret_cast<int*>(static_cast<char*>(ws) + half);
    } else {
      tp.c = 1;
    }
  }
  if (d_pad == 64) rc = launch_scheme<1, false>(a, parts_users, parts_items, st);
  else if (d_pad == 192) rc = launch_score<3, 1, 1, true>(a, st);
  else if (d_pad == 256) rc = launch_score<4, 1, 1, true>(a, st);
  else if (flags & GR_SCORE_FLAG_SINGLE_CTA) rc = launch_scheme<2, false>(a, parts_users, parts_items, st);
  else rc = launch_scheme<2, true>(a, parts_users, parts_items, st);  // CTA pairs (cta_group::2): the default
  if (rc != GR_OK) return rc;
  if (a.splits > 1)
    return gr_topk_merge(part_score, part_id, a.splits, n_users, shortlist, shortlist, sl_score, sl_id, stream);
  if (a.tail_c > 1) {  // merge the item ranges of the tail pairs' rows into the final lists
    const long long tail_row0 = (long long)a.full_pairs * 2 * ROWS_PER_CTA;
    return gr_topk_merge(a.tail_score, a.tail_id, a.tail_c, n_users - tail_row0, shortlist, shortlist,
                         sl_score + tail_row0 * shortlist, sl_id + tail_row0 * shortlist, stream);
  }
  return GR_OK;
}
