// gr_score_topk_tc: all-users x all-items scoring with a fused running top-S shortlist (stage 1 of get_recs).
//
// Replaces the reference's per-user loop (src/metrics.py:52-77): torch.cat repeat of the user row, one
// nn.CosineSimilarity call over all items, a D2H copy of I scores, np.argsort and a Python already-bought filter --
// with ONE dense contraction users[U, D] x items[I, D]^T on the 5th-generation tensor cores:
//
//   * operands: 16-bit (bf16 or fp16), K-major, 128-byte-swizzled shared-memory tiles written by TMA
//     (cp.async.bulk.tensor). With parts == 2 every row carries a hi and a lo half (x = hi + lo up to 2^-18 / 2^-22
//     relative) and the score is the 3-product sum hi.hi + lo.hi + hi.lo, all accumulated in the same TMEM tile,
//     which brings the 16-bit rounding error down to the 1e-5 tie tolerance of the parity contract.
//   * math: tcgen05.mma.cta_group::1.kind::f16, M = 128 users x N = 128 items x K = 16, fp32 accumulators in TMEM
//   * a CTA owns 256 users (two 128-row A tiles, resident in shared memory for the whole sweep) and walks its item
//     range once. B arrives in 32 KB chunks (one part of one 128-item tile) through a TMA ring; every chunk feeds
//     both user tiles, halving the L2 traffic per FLOP. The four 128-column accumulators ([user tile] x [double
//     buffer]) fill all 512 TMEM columns, so the epilogue of tile j overlaps the MMAs of tile j+1.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2-9 = epilogue. An epilogue
//     thread owns ONE user row (TMEM lane) for the whole sweep: tcgen05.ld 32 scores at a time, a 3-input max tree
//     (FMNMX3) and one compare against the row's running threshold (~0.55 instructions per score); only scores that
//     beat the threshold take the slow path: already-bought test against the user's sorted id list (a cached "next
//     bought id" makes the common case one compare) and insertion into the row's sorted shortlist (global memory,
//     thread-private, L1-resident).
//   * scores are never materialised (10M x 1M would be 40 TB); the shortlist is re-scored exactly in fp32 by
//     gr_rescore_topk_f32, which also proves that it contains the exact top-k.
//   * small user counts: the item range is split over blockIdx.y so the grid still fills the chip; the per-split
//     shortlists are merged by gr_topk_merge (host side of this file).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace {

constexpr int TILE_M = 128;    // users per MMA = TMEM lanes
constexpr int UT = 2;          // user tiles per CTA
constexpr int TILE_N = 128;    // items per tile = accumulator columns
constexpr int KBLK = 64;       // 16-bit elements per 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int SUB_BYTES = TILE_M * KBLK * 2;  // one [128 rows][64 x 16 bit] swizzled sub-tile = 16 KB
constexpr int ROWS_PER_CTA = TILE_M * UT;     // 256
constexpr int EPI_WARPS = 4 * UT;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;  // 320
constexpr int TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 232448;  // 227 KB opt-in maximum per CTA

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (sm_100 format: version 1, layout type 2).
// start address >> 4 in [0,14); LBO unused for a single swizzle atom along K; SBO = 8 rows x 128 B = 1024 B.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = fp32, A/B = fp16 (0) or bf16 (1), both K-major, N = 128, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t ab_format) {
  return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((uint32_t)(TILE_N >> 3) << 17) |
         ((uint32_t)(TILE_M >> 4) << 24);
}

// ---- per-row state of the slow path; column t = row within the CTA ------------------------------------------
struct RowState {
  long long* bp;    // [256] cursor into bought_ids
  long long* bend;  // [256] end of the row's bought range
  int* nb;          // [256] bought id at the cursor (INT_MAX when exhausted)
  const int* bought_ids;
  int S;
};

// Slow path (rare): candidate `s` beat the row threshold. ls / li = the row's shortlist (S scores descending, S ids).
// Returns the new threshold (S-th best so far).
__device__ __noinline__ float shortlist_insert(float s, int gid, int t, const RowState* rs, float* __restrict__ ls,
                                               int* __restrict__ li) {
  const int S = rs->S;
  if (gid >= rs->nb[t]) {  // may be an already-bought item: advance the cursor to the first id >= gid
    long long lo = rs->bp[t], hi = rs->bend[t];
    const long long end = hi;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (rs->bought_ids[mid] < gid) lo = mid + 1; else hi = mid;
    }
    const bool is_bought = lo < end && rs->bought_ids[lo] == gid;
    long long nxt = lo;
    if (is_bought)  // skip duplicates of the same id (multi-edges)
      while (nxt < end && rs->bought_ids[nxt] == gid) ++nxt;
    rs->bp[t] = nxt;
    rs->nb[t] = nxt < end ? rs->bought_ids[nxt] : 0x7fffffff;
    if (is_bought) return ls[S - 1];
  }
  int j = S - 1;
  while (j > 0 && ls[j - 1] < s) {
    ls[j] = ls[j - 1];
    li[j] = li[j - 1];
    --j;
  }
  ls[j] = s;
  li[j] = gid;
  return ls[S - 1];
}

template <int KB, int PARTS>
struct Cfg {
  static constexpr int A_BYTES = UT * PARTS * KB * SUB_BYTES;
  static constexpr int CHUNK_BYTES = KB * SUB_BYTES;  // one part of one B tile
  static constexpr int TAIL_BYTES = ROWS_PER_CTA * 20 + 512;
  static constexpr int RING_RAW = (SMEM_LIMIT - 1024 - A_BYTES - TAIL_BYTES) / CHUNK_BYTES;
  static constexpr int RING = RING_RAW > 6 ? 6 : RING_RAW;
  static constexpr int B_BYTES = RING * CHUNK_BYTES;
  static constexpr size_t SMEM = 1024 /*alignment slack*/ + A_BYTES + B_BYTES + TAIL_BYTES;
  static_assert(RING >= 2, "B ring needs at least two chunks");
};

template <int KB, int PARTS>
__global__ void __launch_bounds__(NUM_THREADS, 1)
score_topk_kernel(const __grid_constant__ CUtensorMap tm_users, const __grid_constant__ CUtensorMap tm_items,
                  long long n_users, long long n_items, long long item_id_base, int tiles_per_split, uint32_t idesc,
                  const long long* __restrict__ bought_indptr, const int* __restrict__ bought_ids, int S,
                  float* __restrict__ sl_score, int* __restrict__ sl_id) {
  using L = Cfg<KB, PARTS>;
  constexpr int RING = L::RING;
  constexpr int D_PAD = KB * KBLK;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;
  uint8_t* sB = sA + L::A_BYTES;
  long long* s_bp = reinterpret_cast<long long*>(sB + L::B_BYTES);
  long long* s_bend = s_bp + ROWS_PER_CTA;
  int* s_nb = reinterpret_cast<int*>(s_bend + ROWS_PER_CTA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_nb + ROWS_PER_CTA);
  uint64_t* full = bars;                // [RING]    TMA -> MMA
  uint64_t* empty = full + RING;        // [RING]    MMA -> TMA
  uint64_t* a_full = empty + RING;      // [1]
  uint64_t* t_full = a_full + 1;        // [UT][2]   MMA -> epilogue
  uint64_t* t_empty = t_full + UT * 2;  // [UT][2]   epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + UT * 2);
  __shared__ RowState rs;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles_all = (int)((n_items + TILE_N - 1) / TILE_N);
  const int tile0 = blockIdx.y * tiles_per_split;
  const int n_tiles = max(0, min(tiles_per_split, n_tiles_all - tile0));
  const long long row_base = (long long)blockIdx.x * ROWS_PER_CTA;

  if (threadIdx.x == 0) {
    for (int i = 0; i < RING; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(a_full, 1);
    for (int i = 0; i < UT * 2; ++i) { mbar_init(t_full + i, 1); mbar_init(t_empty + i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    rs.bp = s_bp; rs.bend = s_bend; rs.nb = s_nb; rs.bought_ids = bought_ids; rs.S = S;
  }
  if (warp == 1) {  // TMEM allocation: all 512 columns (this kernel runs one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0 && n_tiles > 0) {
      mbar_expect_tx(a_full, L::A_BYTES);
      for (int ut = 0; ut < UT; ++ut)
        for (int pa = 0; pa < PARTS; ++pa)
          for (int kb = 0; kb < KB; ++kb)
            tma_load_2d(sA + ((ut * PARTS + pa) * KB + kb) * SUB_BYTES, &tm_users, pa * D_PAD + kb * KBLK,
                        (int)(row_base + ut * TILE_M), a_full);
      int g = 0;
      for (int j = 0; j < n_tiles; ++j) {
        for (int pb = 0; pb < PARTS; ++pb, ++g) {
          const int buf = g % RING;
          const uint32_t phase = (uint32_t)(g / RING) & 1u;
          mbar_wait(empty + buf, phase ^ 1u);
          mbar_expect_tx(full + buf, L::CHUNK_BYTES);
          for (int kb = 0; kb < KB; ++kb)
            tma_load_2d(sB + buf * L::CHUNK_BYTES + kb * SUB_BYTES, &tm_items, pb * D_PAD + kb * KBLK,
                        (tile0 + j) * TILE_N, full + buf);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0 && n_tiles > 0) {
      mbar_wait(a_full, 0);
      tc_fence_after();
      int g = 0;
      for (int j = 0; j < n_tiles; ++j) {
        const int slot = j & 1;
        const uint32_t aphase = (uint32_t)(j >> 1) & 1u;
        for (int pb = 0; pb < PARTS; ++pb, ++g) {
          const int buf = g % RING;
          const uint32_t phase = (uint32_t)(g / RING) & 1u;
          mbar_wait(full + buf, phase);
          tc_fence_after();
          for (int ut = 0; ut < UT; ++ut) {
            const uint32_t d_tmem = tmem_base + (uint32_t)((ut * 2 + slot) * TILE_N);
            if (pb == 0) {  // accumulator slot must have been drained by the epilogue of tile j - 2
              mbar_wait(t_empty + ut * 2 + slot, aphase ^ 1u);
              tc_fence_after();
            }
            // B part 0 (hi) pairs with every A part; B part 1 (lo) pairs with A hi only: hi.hi + lo.hi + hi.lo
            const int n_pa = pb == 0 ? PARTS : 1;
            for (int pa = 0; pa < n_pa; ++pa) {
#pragma unroll
              for (int kb = 0; kb < KB; ++kb) {
                const uint64_t da = make_desc_sw128(smem_u32(sA + ((ut * PARTS + pa) * KB + kb) * SUB_BYTES));
                const uint64_t db = make_desc_sw128(smem_u32(sB + buf * L::CHUNK_BYTES + kb * SUB_BYTES));
#pragma unroll
                for (int k = 0; k < KBLK / UMMA_K; ++k)  // +32 bytes (>>4 = 2) per K = 16 step inside the swizzle atom
                  tc_mma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                             (pb | pa | kb | k) != 0 ? 1u : 0u);
              }
            }
            if (pb == PARTS - 1) tc_commit(t_full + ut * 2 + slot);
          }
          tc_commit(empty + buf);
        }
      }
    }
  } else {
    // ================= epilogue: one thread = one user row =================
    const int ew = warp - 2;
    const int ut = ew >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int t = ut * TILE_M + q * 32 + lane;
    const long long row = row_base + t;
    const bool live = row < n_users;
    // the row's shortlist lives in the output arrays: [split][row][S]
    float* ls = sl_score + ((long long)blockIdx.y * n_users + (live ? row : 0)) * S;
    int* li = sl_id + ((long long)blockIdx.y * n_users + (live ? row : 0)) * S;
    if (live)
      for (int s = 0; s < S; ++s) { ls[s] = -INFINITY; li[s] = -1; }
    {
      long long lo = 0, end = 0;
      if (live && bought_indptr != nullptr) {
        lo = bought_indptr[row]; end = bought_indptr[row + 1];
        long long hi = end;  // first bought id >= first item id of this CTA's range
        const long long first_id = item_id_base + (long long)tile0 * TILE_N;
        while (lo < hi) {
          const long long mid = (lo + hi) >> 1;
          if ((long long)bought_ids[mid] < first_id) lo = mid + 1; else hi = mid;
        }
      }
      s_bp[t] = lo; s_bend[t] = end;
      s_nb[t] = lo < end ? bought_ids[lo] : 0x7fffffff;
    }
    float tau = live ? -INFINITY : INFINITY;
    const long long id_end = item_id_base + n_items;
    __syncwarp();
    for (int j = 0; j < n_tiles; ++j) {
      const int slot = j & 1;
      const uint32_t aphase = (uint32_t)(j >> 1) & 1u;
      mbar_wait(t_full + ut * 2 + slot, aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ut * 2 + slot) * TILE_N);
#pragma unroll 1
      for (int c = 0; c < TILE_N / 32; ++c) {
        uint32_t v[32];
        tc_ld32(taddr + (uint32_t)(c * 32), v);
        tc_ld_wait();
        if (c == TILE_N / 32 - 1) {  // accumulator fully read: hand the TMEM slot back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(t_empty + ut * 2 + slot);
        }
        float m[11];
#pragma unroll
        for (int i = 0; i < 10; ++i)
          m[i] = max3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
        m[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
        const float mx = max3(max3(m[0], m[1], m[2]), max3(m[3], m[4], m[5]),
                              max3(max3(m[6], m[7], m[8]), m[9], m[10]));
        if (mx > tau) {
          const long long id0 = item_id_base + (long long)(tile0 + j) * TILE_N + c * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float s = __uint_as_float(v[i]);
            if (s > tau && id0 + i < id_end) tau = shortlist_insert(s, (int)(id0 + i), t, &rs, ls, li);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
  }
}

__global__ void fill_empty_shortlist_kernel(float* sl_score, int* sl_id, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    sl_score[i] = -INFINITY;
    sl_id[i] = -1;
  }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// [rows][row_elems] 16-bit row-major -> 2-D tensor map with a {64 elements, 128 rows} box, 128-byte swizzle, zero OOB fill
int make_map(CUtensorMap* map, const uint16_t* ptr, long long rows, int row_elems, int elem_type) {
  auto enc = get_encode_fn();
  GR_REQUIRE(enc != nullptr, GR_E_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {(cuuint64_t)row_elems, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)KBLK, (cuuint32_t)TILE_M};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, elem_type == GR_ELEM_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                   2, const_cast<uint16_t*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  GR_REQUIRE(r == CUDA_SUCCESS, GR_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
  return GR_OK;
}

struct ScoreArgs {
  CUtensorMap mu, mi;
  long long n_users, n_items, item_id_base;
  int splits, tiles_per_split;
  uint32_t idesc;
  const long long* bptr;
  const int* bids;
  int S;
  float* sl_score;
  int* sl_id;
};

template <int KB, int PARTS>
int launch_score(const ScoreArgs& a, cudaStream_t st) {
  auto kern = score_topk_kernel<KB, PARTS>;
  const size_t smem = Cfg<KB, PARTS>::SMEM;
  static_assert(Cfg<KB, PARTS>::SMEM <= SMEM_LIMIT, "shared memory budget exceeded");
  GR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((a.n_users + ROWS_PER_CTA - 1) / ROWS_PER_CTA), (unsigned)a.splits);
  kern<<<grid, NUM_THREADS, smem, st>>>(a.mu, a.mi, a.n_users, a.n_items, a.item_id_base, a.tiles_per_split, a.idesc,
                                        a.bptr, a.bids, a.S, a.sl_score, a.sl_id);
  GR_LAUNCH_CHECK();
  return GR_OK;
}

// item-range splits: enough CTAs for two waves when the user count alone cannot fill the chip
int choose_splits(long long n_users, long long n_items) {
  const long long ctas = (n_users + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  const long long tiles = (n_items + TILE_N - 1) / TILE_N;
  const long long want = 2LL * gr::sm_count();
  if (ctas >= want || tiles <= 1) return 1;
  long long s = (want + ctas - 1) / ctas;
  s = std::min<long long>(s, std::min<long long>(tiles, GR_SCORE_MAX_SPLITS));
  return (int)std::max<long long>(s, 1);
}

}  // namespace

extern "C" int gr_score_splits(int64_t n_users, int64_t n_items) {
  if (n_users <= 0 || n_items <= 0) return 1;
  return choose_splits(n_users, n_items);
}

extern "C" size_t gr_score_topk_workspace_bytes(int64_t n_users, int64_t n_items, int32_t shortlist) {
  if (n_users <= 0 || n_items <= 0) return 256;
  const int splits = choose_splits(n_users, n_items);
  if (splits == 1) return 256;
  return gr::align_up((size_t)splits * n_users * shortlist * 4, 256) * 2;
}

extern "C" int gr_score_topk_tc(const uint16_t* users_q, int64_t n_users, const uint16_t* items_q, int64_t n_items,
                                int64_t item_id_base, int32_t d_pad, int32_t parts, int32_t elem_type,
                                const int64_t* bought_indptr_or_null, const int32_t* bought_ids_or_null,
                                int32_t shortlist, float* sl_score, int32_t* sl_id, void* ws, size_t ws_bytes,
                                gr_stream_t stream) {
  GR_REQUIRE(n_users >= 0 && n_items >= 0, GR_E_INVALID, "negative size");
  GR_REQUIRE(d_pad == 64 || d_pad == 128, GR_E_INVALID, "d_pad must be 64 or 128 (pad the embeddings with gr_score_prep)");
  GR_REQUIRE(parts == 1 || parts == 2, GR_E_INVALID, "parts must be 1 (single product) or 2 (hi/lo split, 3 products)");
  GR_REQUIRE(elem_type == GR_ELEM_BF16 || elem_type == GR_ELEM_FP16, GR_E_INVALID, "unknown element type");
  GR_REQUIRE(shortlist >= 1 && shortlist <= 32, GR_E_INVALID, "shortlist must be in [1, 32]");
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
  int rc = make_map(&a.mu, users_q, n_users, d_pad * parts, elem_type);
  if (rc != GR_OK) return rc;
  rc = make_map(&a.mi, items_q, n_items, d_pad * parts, elem_type);
  if (rc != GR_OK) return rc;
  a.n_users = n_users; a.n_items = n_items; a.item_id_base = item_id_base;
  a.splits = choose_splits(n_users, n_items);
  const int tiles = (int)((n_items + TILE_N - 1) / TILE_N);
  a.tiles_per_split = (tiles + a.splits - 1) / a.splits;
  a.idesc = make_idesc(elem_type == GR_ELEM_FP16 ? 0u : 1u);
  a.bptr = reinterpret_cast<const long long*>(bought_indptr_or_null);
  a.bids = bought_ids_or_null;
  a.S = shortlist;
  float* part_score = sl_score;
  int* part_id = sl_id;
  if (a.splits > 1) {
    const size_t half = gr::align_up((size_t)a.splits * n_users * shortlist * 4, 256);
    GR_REQUIRE(ws != nullptr && ws_bytes >= 2 * half, GR_E_WORKSPACE, "workspace too small");
    part_score = static_cast<float*>(ws);
    part_id = reinterpret_cast<int*>(static_cast<char*>(ws) + half);
  }
  a.sl_score = part_score; a.sl_id = part_id;
  if (d_pad == 64) rc = parts == 1 ? launch_score<1, 1>(a, st) : launch_score<1, 2>(a, st);
  else rc = parts == 1 ? launch_score<2, 1>(a, st) : launch_score<2, 2>(a, st);
  if (rc != GR_OK) return rc;
  if (a.splits > 1)
    return gr_topk_merge(part_score, part_id, a.splits, n_users, shortlist, shortlist, sl_score, sl_id, stream);
  return GR_OK;
}
