/*
 * gnn_recsys_b200.h -- C ABI of the B200 (sm_100a) embedding + recommendation hot path of hieucnm/GNN-RecSys.
 *
 * The reference has no FFI: its seam is the Python API of src/model.py, src/train/run.py::get_embeddings and
 * src/metrics.py::get_recs sitting on dgl==0.5.2 / torch==1.6 library calls. Each entry point below replaces
 * the library call(s) named in its comment (paths relative to the reference repository root); the Python
 * mirror in gnn-recsys_b200/ binds them with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name ends in _host
 *   - the caller owns all buffers, including scratch ("ws") sized by the matching *_workspace_bytes query
 *   - the library never allocates or frees device memory, never synchronises, never changes the device
 *   - work is enqueued on `stream` (a cudaStream_t); pass 0/NULL for the legacy default stream
 *   - return value: 0 = ok, negative = error (GR_E_*); gr_last_error() returns a thread-local message
 *   - row-major fp32 feature matrices with leading dimension == number of columns; int32 CSR
 */
#ifndef GNN_RECSYS_B200_H_
#define GNN_RECSYS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gr_stream_t; /* cudaStream_t */

enum {
  GR_OK = 0,
  GR_E_INVALID = -1,     /* bad argument (null pointer, unsupported dimension, ...) */
  GR_E_WORKSPACE = -2,   /* workspace too small */
  GR_E_CUDA = -3,        /* a CUDA call or kernel launch failed */
  GR_E_UNSUPPORTED = -4  /* valid request this build cannot serve (e.g. not an sm_100 device) */
};

enum { GR_REDUCE_MEAN = 0, GR_REDUCE_MAX = 1 };            /* fn.mean / fn.max, src/model.py:145-161 */
enum { GR_ACC_STORE = 0, GR_ACC_ADD = 1, GR_ACC_MAX = 2 }; /* HeteroGraphConv aggregate, src/model.py:384-406 */

const char* gr_last_error(void);
int gr_version(void);
/* number of kernels this library has launched in this process (bench.py reports it as gpu_launches) */
long long gr_launch_count(void);
/* sm count / compute capability of the current device; GR_E_UNSUPPORTED unless it is sm_100. */
int gr_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host);

/* ---- a1 NodeEmbedding.forward (src/model.py:19-24, nn.Linear with bias) and the pre-aggregation projection
 *      relu(fc_preagg(h)) of mean_nn / pool_nn (src/model.py:151,158; bias-free). Replaces torch addmm / cuBLAS sgemm.
 *      y[n, d_out] = x[n, d_in] . wt[d_in, d_out] (+ bias) (relu).  `wt` is the nn.Linear weight TRANSPOSED
 *      (k-major) so that output columns are contiguous. Paths (all fp32-accurate):
 *        d_in <= 8                                   fp32 FFMA stream (NodeEmbedding)
 *        d_in == d_out in {128, 256}, no bias,       tcgen05 GEMM: fp16 hi/lo split of both operands with exact power-of-
 *          n >= 256, workspace given                   two row / weight scaling, 3 products, fp32 accumulate in TMEM
 *        d_in % 8 == 0, d_out % 32 == 0, workspace   3xTF32 mma.sync GEMM (also what GR_LINEAR_FLAG_LEGACY selects)
 *        anything else                               fp32 FFMA SGEMM
 *      ws (gr_linear_workspace_bytes(n, d_in, d_out), 256-byte aligned) may be NULL (FFMA paths only). */
enum { GR_LINEAR_FLAG_LEGACY = 1 };
size_t gr_linear_workspace_bytes(int64_t n, int32_t d_in, int32_t d_out);
int gr_linear_f32(const float* x, int64_t n, int32_t d_in, const float* wt, const float* bias_or_null,
                  int32_t d_out, int relu, int32_t flags, float* y, void* ws, size_t ws_bytes, gr_stream_t stream);

/* ---- a3-a6 ConvLayer.forward for one relation, fused (src/model.py:123-237), replacing DGL update_all
 *      (libdgl SpMM copy_u/u_mul_e + mean/max), two torch sgemm, relu, norm/where/div and HeteroGraphConv's
 *      stack+reduce:
 *        n_v   = reduce_{e: dst(e)=v} (edge_w[e] *) h_src[indices[e]]       mean: sum / max(deg,1); max: 0 if deg==0
 *        z_v   = relu(h_dst[v] . w_self_t + n_v . w_neigh_t);  z_v /= (|z_v| or 1 if 0)   when l2norm
 *        out_v = z_scale * ( z_v (GR_ACC_STORE) | out_v + z_v (GR_ACC_ADD) | max(out_v, z_v) (GR_ACC_MAX) )
 *      (HeteroGraphConv 'mean' = ADD with z_scale = 1/n on the last relation, like stack(...).mean(0))
 *      for destination rows v in [row_begin, row_end) (a contiguous shard; out / h_dst are indexed by v).
 *      indptr has n_dst+1 entries and indexes `indices` / `edge_w` absolutely; nnz = indptr[n_dst].
 *      Rows with more than GR_SAGE_LONG_ROW in-edges are reduced by a deterministic multi-CTA split. */
#define GR_SAGE_LONG_ROW 2048
#define GR_SAGE_CHUNK 2048
size_t gr_sage_relation_workspace_bytes(int64_t nnz, int32_t d_neigh);
/* flags: the projection epilogue of the fused kernel is an fp16 hi/lo split on mma m16n8k16 with exact power-of-two row /
 * weight scaling; GR_SAGE_FLAG_TF32_EPILOGUE selects the tf32 hi/lo split on m16n8k8 instead. Both are fp32-accurate
 * (3 products). Per call: the library keeps no mutable global state (re-entrant, one stream per caller thread). */
enum { GR_SAGE_FLAG_TF32_EPILOGUE = 1 };
/* The fused kernel reads the two weight matrices pre-split and pre-packed in MMA fragment order. gr_sage_pack_weights
 * writes that form once per weight update (gr_sage_packed_weights_bytes(...) bytes, 256-byte aligned; 0 = these
 * dimensions take the generic kernel, nothing to pack) for the same `flags` the relation calls will use; pass it as
 * packed_or_null. With NULL every gr_sage_relation_f32 call packs into its workspace (two extra tiny launches). */
size_t gr_sage_packed_weights_bytes(int32_t d_neigh, int32_t d_self, int32_t d_out);
int gr_sage_pack_weights(const float* w_self_t, const float* w_neigh_t, int32_t d_neigh, int32_t d_self, int32_t d_out,
                         int32_t flags, void* packed, gr_stream_t stream);
int gr_sage_relation_f32(const int32_t* indptr, const int32_t* indices, const float* edge_w_or_null, int64_t nnz,
                         const float* h_src, const float* h_dst, int64_t row_begin, int64_t row_end,
                         int32_t d_neigh, int32_t d_self, const float* w_self_t, const float* w_neigh_t,
                         int32_t d_out, int reducer, int l2norm, int accumulate, float z_scale, int32_t flags,
                         const void* packed_or_null, float* out, void* ws, size_t ws_bytes, gr_stream_t stream);

/* The gather-reduce alone (no projection): agg[v - row_begin... indexed by v] = reduce of neighbour rows. Used to
 * report aggregation bandwidth in isolation and by tests; same kernels as above without the epilogue. */
int gr_gather_reduce_f32(const int32_t* indptr, const int32_t* indices, const float* edge_w_or_null, int64_t nnz,
                         const float* h_src, int64_t row_begin, int64_t row_end, int32_t d, int reducer,
                         float* agg, void* ws, size_t ws_bytes, gr_stream_t stream);

/* ---- a10 CosinePrediction.forward for one etype (src/model.py:317-327): replaces F.normalize x2 + DGL SDDMM u_dot_v.
 *      out[e] = <h_src[u[e]], h_dst[v[e]]> / (max(|h_src[u[e]]|, 1e-12) * max(|h_dst[v[e]]|, 1e-12)) */
int gr_edge_cosine_f32(const int32_t* u, const int32_t* v, int64_t n_edges, const float* h_src, const float* h_dst,
                       int32_t d, float* out, gr_stream_t stream);

/* ---- a13 get_recs (src/metrics.py:52-77): all users x all items cosine + top-k, replacing the per-user
 *      torch.cat / nn.CosineSimilarity / np.argsort / Python filter loop.
 *
 *  stage 0  gr_score_prep: rows are L2-normalised (x / max(|x|, 1e-12)), optionally shifted by `center`
 *           (the ranking per user is invariant to a common item shift; it shrinks the quantisation error bound),
 *           zero-padded to d_pad (64, 128, 192 or 256; above 128 only the (1,1) scheme) and rounded to 16 bit (GR_ELEM_BF16 | GR_ELEM_FP16). parts == 2 stores a
 *           hi and a lo half per row ([hi d_pad | lo d_pad], x = hi + lo up to 2^-18 / 2^-22 relative).
 *           stats (float[4], caller-initialised {0, +inf, 0, 0}; atomic max / min over the rows):
 *             [0] max_i |row_i - center|            [1] smallest non-zero |x_i| (tells stage 2 when the eps clamp of
 *             [2] max_i |w_i - (hi_i + lo_i)|           the cosine can bind)
 *             [3] max_i |w_i - hi_i|                 (w = the normalised, centred row; [2] == [3] for parts == 1)
 *           i.e. the MEASURED rounding residuals of the table, which bound the scoring error of stage 1.
 *           gr_colmean_normalized_f32 computes a deterministic `center` (mean of the normalised rows).
 *  stage 1  gr_score_topk_tc: TMA-fed tcgen05 GEMM users x items^T (16-bit in, fp32 accumulate in TMEM) with a
 *           fused per-user running shortlist epilogue that skips already-bought items. (parts_users, parts_items)
 *           selects the product scheme: (1,1) hi.hi; (2,1) hi.hi + lo.hi; (2,2) hi.hi + lo.hi + hi.lo. Bought lists
 *           are a CSR per user (int64 indptr, int32 ids sorted ascending, GLOBAL item ids, duplicates allowed).
 *           A candidate enters the list when its score beats max(S-th best so far, k-th best so far - *band)
 *           (band: device float, NULL = +inf = plain S-th-best rule). user_map (optional int32[n_users]): row u's
 *           bought list is that of user user_map[u] (second pass over a compacted user subset). Output: per user `shortlist` approximate
 *           (centred) scores, descending, and their global item ids (-1 = empty slot).
 *           item_perm (optional int32[n_items]): row p of items_q is item item_perm[p] (+ item_id_base) -- the item
 *           table may be swept in any order; gr_score_item_order + gr_permute_rows build the order that lets the
 *           thresholds settle after the first tiles (descending cosine to the mean user row). The shortlist holds
 *           real item ids either way, so stages 2 / 3 and the result do not depend on it.
 *           flags: GR_SCORE_FLAG_SINGLE_CTA = cta_group::1 kernel instead of CTA pairs (same results);
 *           GR_SCORE_FLAG_NO_TAIL_SPLIT = do not cut the user tiles of the last, partial wave into item ranges.
 *  stage 2  gr_rescore_topk_f32: exact fp32 cosine (torch formula x.y / sqrt(max(|x|^2 |y|^2, eps^2))) of the
 *           shortlisted items that can still reach the top-k, sorted by (score desc, id asc), first k. Proves per user
 *           that no item outside the shortlist can enter the top-k by more than tie_tol. With the user's own rounding
 *           residuals r_u (final) and r1_u (first level), recomputed here exactly as stage 0 rounds,
 *             err_u = r_u*Y + (1 + r_u)*R + [parts_users == parts_items == 2] r1_u*R1*(1 + u)^2 + acc_err*(1 + r_u)*(Y + R)
 *           (Y, R, R1 = item stats [0], [2], [3]; u = unit roundoff of elem_type) bounds |approx - exact| of every
 *           item, and every item stage 1 dropped has approx <= dropped = max(tau_S if the list is full, tau_k - *band).
 *           Users with kth_exact < dropped + err_u + x.center - tie_tol are appended to overflow_users / n_overflow
 *           (caller zero-initialises n_overflow) and must be recomputed by a more accurate scheme or by stage 3.
 *           user_map (optional int32[n_users]): overflow entries are user_map[u] instead of u (second pass over a
 *           compacted user subset).
 *  stage 3  gr_score_topk_exact_f32: exact fp32 scoring of all items for the listed users (the overflow list, or
 *           every user when user_list == NULL): the always-correct fallback and the brute-force checker. With
 *           popularity != NULL it ranks by softmax_i(cos) + weight * popularity_i instead (use_popularity branch of
 *           get_recs, src/metrics.py:69-72; popularity[i] belongs to local item i), two passes over the items.
 *  merge    gr_topk_merge: row-wise merge of `parts` partial (score desc, id) lists into the k_out best
 *           (scores[p][u][k_in]); ties by smaller id; ids < 0 are empty slots. */
enum { GR_ELEM_BF16 = 0, GR_ELEM_FP16 = 1 };
enum { GR_SCORE_FLAG_SINGLE_CTA = 1, GR_SCORE_FLAG_NO_TAIL_SPLIT = 2 };
#define GR_SCORE_MAX_SPLITS 32
size_t gr_colmean_workspace_bytes(int64_t n, int32_t d);
int gr_colmean_normalized_f32(const float* x, int64_t n, int32_t d, float* center, void* ws, size_t ws_bytes,
                              gr_stream_t stream);
int gr_score_prep(const float* x, int64_t n, int32_t d, const float* center_or_null, int32_t d_pad, int32_t parts,
                  int32_t elem_type, uint16_t* out_q, float* stats4_or_null, gr_stream_t stream);
int gr_score_splits(int64_t n_users, int64_t n_items);
size_t gr_score_topk_workspace_bytes(int64_t n_users, int64_t n_items, int32_t shortlist);
int gr_score_topk_tc(const uint16_t* users_q, int64_t n_users, const uint16_t* items_q, int64_t n_items,
                     int64_t item_id_base, const int32_t* item_perm_or_null, int32_t d_pad, int32_t parts_users,
                     int32_t parts_items, int32_t elem_type, const int64_t* bought_indptr_or_null, const int32_t* bought_ids_or_null, int32_t shortlist,
                     int32_t k, const float* band_or_null, const int32_t* user_map_or_null, int32_t flags,
                     float* sl_score, int32_t* sl_id, void* ws, size_t ws_bytes, gr_stream_t stream);
/* Item sweep order of stage 1: perm[p] = index of the item swept at position p, descending cos(h_item[i], dir)
 * (65536 equal-width buckets between the smallest and largest cosine, ascending index inside a bucket; dir = any
 * positive multiple of the mean normalised user row, e.g. gr_colmean_normalized_f32 of the user table).
 * gr_permute_rows: dst[p] = src[perm[p]] for rows of row_bytes (multiple of 16; not in place). */
size_t gr_score_item_order_workspace_bytes(int64_t n_items);
int gr_score_item_order(const float* h_item, int64_t n_items, int32_t d, const float* dir, int32_t* perm, void* ws,
                        size_t ws_bytes, gr_stream_t stream);
int gr_permute_rows(const void* src, int64_t n_rows, int64_t row_bytes, const int32_t* perm, void* dst,
                    gr_stream_t stream);
/* *band = 2 x the largest err_u of stage 2 over a user table whose gr_score_prep statistics are user_stats4 */
int gr_score_band(const float* item_stats4, const float* user_stats4, int32_t elem_type, int32_t parts_users,
                  int32_t parts_items, float acc_err, float* band, gr_stream_t stream);
int gr_rescore_topk_f32(const float* h_user, const float* h_item, int64_t item_id_base, int32_t d,
                        const float* center_or_null, const float* sl_score, const int32_t* sl_id, int32_t shortlist,
                        int64_t n_users, const float* item_stats4, int32_t elem_type, int32_t parts_users,
                        int32_t parts_items, float acc_err, const float* band_or_null, float tie_tol, int32_t k,
                        float eps, const int32_t* user_map_or_null, int32_t* out_ids, float* out_scores,
                        int32_t* overflow_users, int32_t* n_overflow, gr_stream_t stream);
int gr_score_topk_exact_f32(const float* h_user, const int32_t* user_list_or_null, const int32_t* n_list_or_null,
                            int64_t n_users, const float* h_item, int64_t n_items, int64_t item_id_base, int32_t d,
                            const int64_t* bought_indptr_or_null, const int32_t* bought_ids_or_null, int32_t k,
                            float eps, const float* popularity_or_null, float weight_popularity, int32_t* out_ids,
                            float* out_scores, gr_stream_t stream);
int gr_topk_merge(const float* scores, const int32_t* ids, int32_t parts, int64_t n_users, int32_t k_in,
                  int32_t k_out, float* out_scores, int32_t* out_ids, gr_stream_t stream);

/* ---- metrics@k next to the path: recs_to_metrics (src/metrics.py:81-107) as counters.
 *      recs[n_users][k] (-1 = empty), ground truth as a CSR per user (int64 indptr, int32 ids sorted ascending,
 *      duplicates kept). counters5 (device uint64[5]): [0] recommended ids, [1] of those in the ground truth
 *      (precision = [1]/[0]), [2] ground-truth entries, [3] of those recommended (recall = [3]/[2]), [4] distinct
 *      recommended items (coverage = [4]/n_items). */
size_t gr_metrics_workspace_bytes(int64_t n_items);
int gr_metrics_at_k(const int32_t* recs, int64_t n_users, int32_t k, const int64_t* truth_indptr,
                    const int32_t* truth_ids, int64_t n_items, uint64_t* counters5, void* ws, size_t ws_bytes,
                    gr_stream_t stream);

/* ---- graph ingest next to the path: stable COO -> int32 CSR over destination rows on the
 *      device, replacing what dgl.heterograph (src/builder.py:377-383) + DGL's lazy CSC build behind update_all do on
 *      the CPU. LSD radix sort of (dst, edge id): neighbours of a row stay in edge-id order (bit-exact against a
 *      stable argsort). eperm[j] = edge id stored in CSR slot j; indices[j] = src[eperm[j]]; indptr has n_dst + 1
 *      entries. src / dst are int32 device arrays with 0 <= dst < n_dst; *status (optional device int32) is set to 0,
 *      or to 1 when some dst lies outside that range (invalid input: the outputs are then unspecified, but the build
 *      never writes out of bounds). */
size_t gr_csr_build_workspace_bytes(int64_t nnz, int32_t n_dst);
int gr_csr_build_i32(const int32_t* src, const int32_t* dst, int64_t nnz, int32_t n_dst, int32_t* indptr,
                     int32_t* indices, int32_t* eperm, int32_t* status_or_null, void* ws, size_t ws_bytes,
                     gr_stream_t stream);

/* ---- ID remap next to the path: raw ids -> contiguous ids in order of FIRST APPEARANCE,
 *      replacing create_ids (src/builder.py:182-227: pandas unique() order + merge). new_ids[i] in [0, *n_unique);
 *      uniq_raw[new id] = raw id (optional reverse map, capacity n); n_unique is a device int32. Any int64 is a
 *      valid raw id. Deterministic (hash table keeps the smallest position per id) and bit-exact vs the CPU rule. */
size_t gr_remap_workspace_bytes(int64_t n);
int gr_remap_first_appearance_i64(const int64_t* raw, int64_t n, int32_t* new_ids, int64_t* uniq_raw_or_null,
                                  int32_t* n_unique, void* ws, size_t ws_bytes, gr_stream_t stream);

/* ---- sampled blocks next to the path: the frontier sampling and negative sampling that
 *      dgl.dataloading does on the CPU behind the reference's EdgeDataLoader / NodeDataLoader (src/sampling.py:153-207;
 *      consumed by the training-step forward, src/train/run.py:89-139). Counter-based randomness, restated bit for bit
 *      by oracle.straightline.sample_frontier / negative_uniform:
 *        hash64(key, ctr) = fin(key + (ctr + 1) * 0x9E3779B97F4A7C15), fin = the splitmix64 finaliser
 *        gr_sample_key(seed, stream_id) = hash64(fin(seed + 0x9E3779B97F4A7C15), stream_id)   (host helper, no GPU)
 *
 *  gr_sample_count_i32 / gr_sample_fill_i32: frontier of one relation (CSR over destination rows as built by
 *      gr_csr_build_i32; eperm NULL = edge id == CSR slot) for `seeds` (distinct destination ids, int64 like DGL's
 *      NID). fanout <= 0: every in-edge (MultiLayerFullNeighborSampler / in_subgraph); 1..32: at most `fanout`
 *      in-edges per seed, uniformly without replacement (MultiLayerNeighborSampler(fanouts, replace=False)): the
 *      candidates with the smallest (hash64(key, edge id) >> 32, CSR slot). Edge ids in `excl_sorted` (ascending int32;
 *      the batch's own edges and their reverse twins, exclude='reverse_types') are never candidates.
 *      count: out_indptr[n_seeds + 1] = exclusive scan of the per-seed counts (the block's indptr), *total =
 *      out_indptr[n_seeds] (device int32). fill: out_src[total] (GLOBAL source ids, int64 so that the buffer can feed
 *      gr_remap_first_appearance_i64 directly) and out_eid[total], per seed in CSR (= edge id) order.
 *  gr_negative_uniform_i64: negative_sampler.Uniform(k): for positive edge e (id eids[e], source edge_src[eids[e]]):
 *      out_src[e*k + j] = that source, out_dst[e*k + j] = hash64(key, eids[e]*k + j) mod n_dst_nodes, j < k. */
uint64_t gr_sample_key(uint64_t seed, uint64_t stream_id);
int gr_sample_count_i32(const int32_t* indptr, const int32_t* eperm_or_null, const int64_t* seeds, int64_t n_seeds,
                        int32_t fanout, const int32_t* excl_sorted_or_null, int32_t n_excl, int32_t* out_indptr,
                        int32_t* total, gr_stream_t stream);
int gr_sample_fill_i32(const int32_t* indptr, const int32_t* indices, const int32_t* eperm_or_null,
                       const int64_t* seeds, int64_t n_seeds, int32_t fanout, const int32_t* excl_sorted_or_null,
                       int32_t n_excl, uint64_t key, const int32_t* out_indptr, int64_t* out_src, int32_t* out_eid,
                       gr_stream_t stream);
int gr_negative_uniform_i64(const int32_t* edge_src, const int64_t* eids, int64_t n_pos, int32_t k,
                            int64_t n_dst_nodes, uint64_t key, int64_t* out_src, int64_t* out_dst, gr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GNN_RECSYS_B200_H_ */
