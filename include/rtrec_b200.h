/*
 * rtrec_b200.h -- C-ABI of the B200-native SLIM hot path (librtrec_b200.so).
 *
 * The reference (myui/rtrec 0.2.7, pure Python) has no FFI of its own; the seam this library
 * replaces is the pair of Python classes that `rtrec.models.SLIM` composes:
 *   - the store    `UserItemInteractions`  (/root/reference/rtrec/utils/interactions.py:14-353)
 *   - the operator `SLIMElastic`           (/root/reference/rtrec/models/internal/slim_elastic.py:156-857)
 * Each entry point below cites the reference lines it replaces.  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; rt_last_error() gives the message
 *     (thread-local).  No exceptions cross the boundary, no torch types in signatures.
 *   - pointers named d_* are DEVICE pointers (cudaMalloc / torch CUDA tensors), h_* are HOST
 *     pointers.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *     All device work is enqueued on that stream; functions do not synchronise unless stated.
 *   - matrices: float32 values, int32 indices/offsets (the reference builds float32/int32 scipy
 *     matrices, interactions.py:276,303).  CSR = (rptr[n_rows+1], ridx, rval), CSC likewise.
 *   - temporary device storage is owned by the library (grow-only arenas, rt_release_scratch());
 *     use one caller thread per process and keep calls stream-ordered, as the reference's
 *     single-threaded Python does (SURVEY.md 8b "Threading").
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef RTREC_B200_H
#define RTREC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define RT_OK 0
#define RT_ERR_ARG (-1)
#define RT_ERR_CUDA (-2)
#define RT_ERR_CAPACITY (-3) /* an output buffer was too small; sizes needed are reported */
#define RT_ERR_NO_DEVICE (-4)

int rt_version(void);
const char *rt_last_error(void);
/* number of SMs / bytes of opt-in shared memory per block of the current device (<0 on error) */
int rt_device_info(int *sm_count, int *smem_optin_bytes, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------------------------
 * Interaction store  (replaces interactions.py:81-119 add_interaction, :62-79 _apply_decay,
 * :259-303 to_csr/to_csc).  State = three device arrays sorted by key = (user << 32 | item):
 *   keys u64[n_pairs], vals f64[n_pairs] (stored rating), stamps f64[n_pairs] (last timestamp).
 * ---------------------------------------------------------------------------------------- */

/*
 * Fold one batch of events (arrival order) into the store.  Semantics per event k, exactly as
 * add_interaction (interactions.py:99-111):
 *   T_k = max(max_ts_in, max_{m<=k}(ts_m + 1));
 *   upsert: (val, stamp) <- (delta_k, ts_k)                      [no clipping]
 *   else  : cur = 0 if pair absent or stored val == 0 else val * rate^((T_k - stamp)/86400)
 *           (val, stamp) <- (clip(cur + delta_k, min_value, max_value), ts_k)
 * decay_rate <= 0 or NaN means "no decay" (decay_rate is None in the reference).
 * Inputs d_users/d_items int32, d_ts f64, d_delta f64.  The merged store is written to
 * d_out_* (capacity out_cap pairs, must be >= n_pairs + n_events); *h_n_out receives the new
 * pair count, *h_max_ts/-user/-item the running maxima (max_ts_in etc. are the carried-in
 * values; max ids start at 0 like interactions.py:40-41).  Synchronises the stream.
 */
int rt_store_fold(const int32_t *d_users, const int32_t *d_items, const double *d_ts,
                  const double *d_delta, int64_t n_events, int upsert, double min_value,
                  double max_value, double decay_rate, const uint64_t *d_keys, const double *d_vals,
                  const double *d_stamps, int64_t n_pairs, double max_ts_in, int32_t max_user_in,
                  int32_t max_item_in, uint64_t *d_out_keys, double *d_out_vals, double *d_out_stamps,
                  int64_t out_cap, int64_t *h_n_out, double *h_max_ts, int32_t *h_max_user,
                  int32_t *h_max_item, void *stream);

/*
 * Batch bookkeeping of add_interaction for events that already sit on the device as the int64 / f64
 * columns a DataFrame holds (interactions.py:92-99 future-timestamp check and max_timestamp,
 * :113-119 all_item_ids / hot_items / max ids).  rt_events_minmax returns the id ranges (validation,
 * max_user_id / max_item_id) and the largest timestamp; synchronises.  rt_events_item_stats fills,
 * per item id < n_items: the number of events with delta > 0 (hot_items frequency, :115-116), the
 * arrival index of the last such event (-1 if none; LRU order) and a seen flag (all_item_ids).
 */
int rt_events_minmax(const int64_t *d_users, const int64_t *d_items, const double *d_ts, int64_t n,
                     int64_t *h_min_user, int64_t *h_max_user, int64_t *h_min_item, int64_t *h_max_item,
                     double *h_max_ts, void *stream);
int rt_events_item_stats(const int64_t *d_items, const double *d_delta, int64_t n, int32_t n_items,
                         int32_t *d_count_pos, int32_t *d_last_pos, uint8_t *d_seen, void *stream);
/* same for int32 item ids (what rt_upload_events leaves on the device) */
int rt_events_item_stats32(const int32_t *d_items, const double *d_delta, int64_t n, int32_t n_items,
                           int32_t *d_count_pos, int32_t *d_last_pos, uint8_t *d_seen, void *stream);

/*
 * Host -> device upload of one batch of events held as four HOST columns in pageable memory (the
 * columns of the DataFrame that recommender.py:203-223 would otherwise walk row by row): int64 user
 * and item ids are narrowed to int32 on the way, timestamps / ratings stay float64.  n_threads host
 * threads (0 = choose) each convert a chunk into a pinned staging slot owned by the library and
 * issue the DMA on their own copy stream, so conversion, staging and PCIe transfers overlap.
 * Also returns the id ranges and the largest timestamp of the batch (interactions.py:92-99,
 * 118-119); ids outside [0, 2^31) are reported through these ranges and must be rejected by the
 * caller (the device copies are then meaningless).  Waits for `stream` first (the destinations may
 * come from a stream-ordered allocator) and returns when the data is resident.
 */
int rt_upload_events(const int64_t *h_users, const int64_t *h_items, const double *h_ts,
                     const double *h_delta, int64_t n, int32_t *d_users, int32_t *d_items, double *d_ts,
                     double *d_delta, int64_t *h_min_user, int64_t *h_max_user, int64_t *h_min_item,
                     int64_t *h_max_item, double *h_max_ts, int32_t n_threads, void *stream);

/*
 * Build the float32 CSR and CSC interaction matrices from the store (to_csr / to_csc,
 * interactions.py:259-303): x = f32(val * rate^((max_ts - stamp)/86400)), shape
 * (n_users, n_items) = (max_user_id+1, max_item_id+1); explicit zeros are kept.
 * d_item_mask (optional, uint8[n_items]) restricts the matrices to the flagged item columns
 * (to_csc(select_items), interactions.py:296).  Outputs have capacity n_pairs entries;
 * d_ccol (optional, int32[n_pairs]) receives the column id of every CSC entry (COO view used
 * by the Gram kernel).  *h_nnz = number of entries kept, *h_nonneg = 1 iff all kept values
 * are >= 0.  Synchronises the stream.
 */
int rt_store_build(const uint64_t *d_keys, const double *d_vals, const double *d_stamps,
                   int64_t n_pairs, double decay_rate, double max_ts, int32_t n_users, int32_t n_items,
                   const uint8_t *d_item_mask, int32_t *d_rptr, int32_t *d_ridx, float *d_rval,
                   int32_t *d_cptr, int32_t *d_cidx, float *d_cval, int32_t *d_ccol, int64_t *h_nnz,
                   int *h_nonneg, void *stream);

/* Point lookup of stored (val, stamp) for n (user,item) queries; found[q]=0 if absent.
 * (get_user_item_rating, interactions.py:134-149, without the decay which the host applies.) */
int rt_store_lookup(const uint64_t *d_keys, const double *d_vals, const double *d_stamps, int64_t n_pairs,
                    const uint64_t *d_query, int64_t n, double *d_val_out, double *d_stamp_out,
                    uint8_t *d_found, void *stream);

/* ------------------------------------------------------------------------------------------
 * Fit  (replaces SLIMElastic.fit / fit_in_parallel / partial_fit_items, slim_elastic.py:229-564,
 * FeatureSelectionWrapper.fit :139-154 and sklearn's sparse_enet_coordinate_descent)
 * ---------------------------------------------------------------------------------------- */

/*
 * Item-item Gram rows.  For every stored entry e of the CSC/COO view (column j = d_ccol[e],
 * user u = d_cidx[e], value y = d_cval[e]) adds y * X[u, :] into row j of G (ld = ldg floats).
 * Row j of G is then X^T x_j, which is both the feature-score vector of target j
 * (slim_elastic.py:141, X.T.dot(y)) and row j of the Gram matrix the solver replays on.
 * G must be zero-filled by the caller for the rows being produced.  Entries [e_begin, e_end)
 * only are processed (lets callers shard by item range across GPUs / build row subsets).
 */
int rt_gram_rows(const int32_t *d_ccol, const int32_t *d_cidx, const float *d_cval, int64_t e_begin,
                 int64_t e_end, const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                 float *d_G, int64_t ldg, void *stream);

/*
 * Same result as rt_gram_rows for the target columns [j_begin, j_end) of the CSC matrix, computed
 * with warp-private shared-memory accumulators instead of global atomics (gram2.cu): per (j, i)
 * the fp32 sum runs in ascending user order like scipy's csr_matvec (slim_elastic.py:141).
 * G rows being produced must be zero-filled by the caller.  Synchronises the stream.
 */
int rt_gram(int32_t n_users, int32_t n_items, const int32_t *d_cptr, const int32_t *d_cidx,
            const float *d_cval, const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
            int32_t j_begin, int32_t j_end, float *d_G, int64_t ldg, void *stream);

/*
 * Third-generation Gram (gram3.cu), the default fit path.  Items are relabelled by popularity rank
 * (count desc, id asc); rt_gram_lower computes the rows [h_cuts[part], h_cuts[part+1]) of the
 * LOWER triangle of G' = X'^T X' in rank space into d_Gp (zero-filled by the caller, ld = ldgp),
 * where the n_parts row ranges are balanced by exact multiply-add count; d_rank_of / d_orig_of
 * (int32[n_items]) receive the permutation.  With n_parts > 1 the caller exchanges the row slabs
 * (all-gather) before rt_gram_finish, which mirrors the triangle in place and writes the full
 * symmetric G in the caller's item ids to d_G (every entry is written; d_G != d_Gp).
 * rt_gram_lower synchronises only when n_parts > 1 (to return h_cuts[n_parts + 1]).
 */
int rt_gram_lower(int32_t n_users, int32_t n_items, const int32_t *d_cptr, const int32_t *d_cidx,
                  const float *d_cval, const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                  int64_t nnz, int32_t part, int32_t n_parts, float *d_Gp, int64_t ldgp, int32_t *d_rank_of,
                  int32_t *d_orig_of, int32_t *h_cuts, void *stream);
int rt_gram_finish(int32_t n_items, float *d_Gp, int64_t ldgp, const int32_t *d_rank_of,
                   const int32_t *d_orig_of, float *d_G, int64_t ldg, void *stream);

/* rt_gram_finish that also writes d_rowmax[i] = max_{c != i} G[i][c] (float[n_items], by item id) while the rows pass
 * through shared memory; *h_has_rowmax = 0 when the rows are too long to be staged (d_rowmax is then left untouched). */
int rt_gram_finish_rowmax(int32_t n_items, float *d_Gp, int64_t ldgp, const int32_t *d_rank_of,
                          const int32_t *d_orig_of, float *d_G, int64_t ldg, float *d_rowmax,
                          int32_t *h_has_rowmax, void *stream);

/*
 * Multi-GPU form of rt_gram_finish: the slab exchange is fused with the mirror step over peer memory.
 * h_slabs[p] (HOST array of n_parts DEVICE pointers, p = part index) addresses the d_Gp buffer of rank p
 * as seen from this process: the own buffer for p == part, a CUDA IPC mapping (rt_ipc_open) of the
 * peer's buffer otherwise; all have the same ldgp.  The call is made three times per fit:
 *   phase 0  tiles of this rank's stripe (and those touching its own slab) are read from the ranks that
 *            computed them (NVLink P2P loads), stored into the own buffer and mirrored into its upper
 *            triangle in one pass;
 *   phase 1  the remaining tiles are read from their stripe rank, which holds them since phase 0 -- so
 *            a large slab leaves its owner once, not once per peer, and every GPU sends about the same;
 *   phase 2  the symmetric matrix is written in the caller's item ids to d_G exactly like rt_gram_finish.
 * The caller orders the ranks with node barriers (e.g. a one-element NCCL all-reduce on the stream):
 * every rank has finished rt_gram_lower before any phase 0, every phase 0 before any phase 1, and no
 * rank overwrites its buffer until every rank has finished phase 1.
 */
#define RT_MAX_PEERS 8
int rt_gram_finish_p2p(int32_t n_items, const void *const *h_slabs, int32_t n_parts, int32_t part,
                       const int32_t *h_cuts, int64_t ldgp, const int32_t *d_rank_of,
                       const int32_t *d_orig_of, float *d_G, int64_t ldg, int32_t phase, void *stream);

/*
 * Multi-GPU fit without a full Gram exchange ("owner rows").  Replaces, like rt_gram_lower, the per-column
 * `X.T.dot(y)` of slim_elastic.py:141 -- for the target columns a GPU owns -- and hands the solver rows that may
 * live on a peer GPU.
 *
 * Ownership is block-cyclic in popularity-rank space: the 64-row block b of the rank-space Gram matrix belongs
 * to part b % n_parts and is stored at local block b / n_parts of that part's buffers (rt_gram_block_rows gives
 * the buffer height `rows_alloc`, identical on every part, and the number of real rows `rows_own`).  Per fit:
 *   1. rt_gram_lower_blocks   lower-triangle part (columns <= row) of the own rows into d_slab
 *                             [rows_alloc, ldgp] (zero-filled by the caller); also writes rank_of / orig_of;
 *   2. node barrier, then rt_gram_pull_cols: the rest of every own row is the transposed column below it in
 *                             the triangle; its 64 x 64 tiles are read out of the owners' slabs (h_slabs[p]: own
 *                             buffer or CUDA IPC mapping, NVLink P2P loads) and stored transposed.  Each GPU
 *                             pulls (N-1)/N^2 of the matrix; egress is balanced by construction;
 *   3. rt_gram_unpermute_rows columns back to item ids: d_rows[r][i] = d_slab[r][rank_of[i]] (rows keep
 *                             their local position), into a second peer-mapped buffer;
 *   4. rt_gram_row_slots      d_slots[i] = (owner part << 24) | local row of item i;
 *   5. node barrier, then rt_slim_solve_rows on the own targets (the items of the own blocks): rows of other
 *                             targets' candidates are gathered from the peers' d_rows through h_bases.
 * No rank may overwrite its buffers before every rank has finished step 5 (node barrier at the start of
 * the next fit).  With n_parts = 1 the result equals rt_gram_lower + rt_gram_finish + rt_slim_solve bit for
 * bit; with n_parts > 1 every row is assembled from the same lower-triangle values, so W is identical too.
 */
int rt_gram_block_rows(int32_t n_items, int32_t n_parts, int32_t part, int32_t *h_rows_alloc, int32_t *h_rows_own);
int rt_gram_lower_blocks(int32_t n_users, int32_t n_items, const int32_t *d_cptr, const int32_t *d_cidx,
                         const float *d_cval, const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                         int64_t nnz, int32_t part, int32_t n_parts, float *d_slab, int64_t ldgp,
                         int32_t *d_rank_of, int32_t *d_orig_of, void *stream);
int rt_gram_pull_cols(int32_t n_items, const void *const *h_slabs, int32_t n_parts, int32_t part, int64_t ldgp,
                      void *stream);
int rt_gram_unpermute_rows(int32_t n_rows, int32_t n_items, const float *d_slab, int64_t ldgp,
                           const int32_t *d_rank_of, float *d_rows, int64_t ldg, void *stream);
int rt_gram_row_slots(int32_t n_items, const int32_t *d_rank_of, int32_t n_parts, int32_t *d_slots, void *stream);

/*
 * Device buffers that other processes of the same node can map (CUDA IPC): rt_ipc_alloc returns a
 * cudaMalloc'ed pointer and its 64-byte handle; a peer process passes the handle to rt_ipc_open and
 * gets a pointer valid in its own address space (peer access over NVLink is enabled on demand).
 * rt_ipc_close unmaps, rt_ipc_free releases the owner's allocation.  rt_memset is cudaMemsetAsync for
 * such raw buffers.
 */
#define RT_IPC_HANDLE_BYTES 64
int rt_ipc_alloc(size_t bytes, void **d_ptr, uint8_t *h_handle);
int rt_ipc_open(const uint8_t *h_handle, void **d_ptr);
int rt_ipc_close(void *d_ptr);
int rt_ipc_free(void *d_ptr);
int rt_memset(void *d_ptr, int32_t value, size_t bytes, void *stream);

/*
 * Range-split row pointers of a CSR/CSC matrix with ascending minor indices:
 * d_seg[row * (n_ranges + 1) + g] = first position of `row` whose minor index is
 * >= base + g * range_width.  Helper of rt_gram; exported for tests.
 */
int rt_csr_split(int32_t n_rows, const int32_t *d_ptr, const int32_t *d_idx, int32_t base,
                 int32_t range_width, int32_t n_ranges, int32_t *d_seg, void *stream);

typedef struct {
    double alpha;        /* SLIMElastic.alpha      (slim_elastic.py:184) */
    double l1_ratio;     /* SLIMElastic.l1_ratio   (:185) */
    double tol;          /* :188 */
    int32_t max_iter;    /* :187 */
    int32_t positive;    /* positive_only :186 */
    uint32_t seed;       /* RandomState(random_state).randint(0, 2^31-1), _cd_fast.pyx:748 */
    int32_t nn;          /* nn_feature_selection (:190), 0 = all items are features */
    int32_t n_samples;   /* rows of X = max_user_id+1; scales the penalties (_coordinate_descent.py:781) */
    int32_t nonneg;      /* 1 iff every stored value of X is >= 0 (enables live-set pruning) */
    uint64_t rowmax_ptr;  /* optional DEVICE pointer (0 = none) to float[n_items]: the largest off-diagonal Gram entry of every
                            item's row, as rt_gram_finish_rowmax writes it.  With skip_trivial the warp solver then finishes a
                            target without a live coordinate without reading its Gram row at all. */
    int32_t skip_trivial; /* 1: with feature selection (nn > 0), a target whose Gram row holds no entry above
                            alpha*l1_ratio*n_samples -- all nn coefficients are 0 before the first sweep -- may be returned
                            with NO pairs (d_out_cnt = 0) instead of nn zeros, and its candidates are not selected.  Valid
                            when the result is assembled into an empty W (a bulk fit: zeros are never stored,
                            slim_elastic.py:273-274) and d_sel_out is not requested; a partial fit must pass 0, because a
                            returned zero deletes a stale entry of the old matrix (:533-538).  Ignored unless positive and
                            nonneg are set.  (All-features mode, nn == 0, returns only non-zero coefficients anyway.) */
} rt_fit_config;

/* xorshift32 draw table: out[t] = value of the (t+1)-th our_rand_r call from `seed`
 * (sklearn/utils/_random.pxd:20-34).  The sequence is identical for every column. */
int rt_rng_table(uint32_t seed, int64_t n, uint32_t *d_out, void *stream);

/*
 * Solve the ElasticNet problems of `n_targets` target columns on the Gram matrix.
 *   d_G          [n_items, ldg] float32, complete for every item with a non-empty column and symmetric
 *                (rt_gram_finish / rt_gram_finish_p2p output is, bit for bit: the warp-per-column kernel
 *                reads G[a][b] or G[b][a], whichever is convenient)
 *   d_targets    int32[n_targets] target item ids
 *   d_sel_in     optional int32[n_targets * nn]: candidate order to use instead of the built-in
 *                selection (score desc, ties -> larger item id first)
 *   d_rng        table from rt_rng_table with >= max_iter * min(nn or n_items, n_items) + 64 entries
 * Outputs
 *   d_sel_out    optional int32[n_targets * nn] candidate order used (-1 padded)
 *   d_out_off    int64[n_targets] offset of each target's result in d_out_rows/d_out_vals
 *   d_out_cnt    int32[n_targets] number of (row, value) pairs.  nn > 0: exactly min(nn,n_items)
 *                pairs in pick order, zeros included (slim_elastic.py:153); nn == 0: the non-zero
 *                coefficients in ascending row (sparse_coef_).
 *   d_out_rows/d_out_vals capacity out_cap pairs in total; on overflow returns RT_ERR_CAPACITY
 *                after the kernel has finished (*h_needed = pairs required).
 *   d_stats      optional int32[n_targets * 4] = n_iter, draws, gap evaluations, live-set size
 * Synchronises the stream.
 */
int rt_slim_solve(const float *d_G, int64_t ldg, int32_t n_items, const int32_t *d_targets,
                  int32_t n_targets, const rt_fit_config *cfg, const int32_t *d_sel_in,
                  const uint32_t *d_rng, int64_t rng_len, int32_t *d_sel_out, int64_t *d_out_off,
                  int32_t *d_out_cnt, int32_t *d_out_rows, float *d_out_vals, int64_t out_cap,
                  int64_t *h_needed, int32_t *d_stats, void *stream);

/* rt_slim_solve on the owner-rows layout (see rt_gram_lower_blocks): row i of G is read from buffer
 * h_bases[d_rowslot[i] >> 24] (HOST array of n_bases DEVICE pointers: own memory or CUDA IPC mappings of the
 * peers' row buffers, 16-byte aligned, common leading dimension ldg, a multiple of 4) at local row
 * d_rowslot[i] & 0xffffff.  Every other argument as rt_slim_solve. */
int rt_slim_solve_rows(const void *const *h_bases, int32_t n_bases, const int32_t *d_rowslot, int64_t ldg,
                       int32_t n_items, const int32_t *d_targets, int32_t n_targets, const rt_fit_config *cfg,
                       const int32_t *d_sel_in, const uint32_t *d_rng, int64_t rng_len, int32_t *d_sel_out,
                       int64_t *d_out_off, int32_t *d_out_cnt, int32_t *d_out_rows, float *d_out_vals,
                       int64_t out_cap, int64_t *h_needed, int32_t *d_stats, void *stream);

/*
 * Fit with positive coefficients on non-negative data WITHOUT the dense Gram matrix -- all features (nn == 0), or feature
 * selection in a bulk fit (cfg->skip_trivial) -- with the same result contract as rt_gram_lower + rt_gram_finish +
 * rt_slim_solve (replaces slim_elastic.py:229-281 for those configurations).
 * A coordinate c of target j can only leave 0 if G[j][c] > alpha*l1_ratio*n_samples, and G[j][c] <= sqrt(G[j][j] G[c][c]):
 * one pass over X (column sums of squares) names the items that can take part in any non-zero solution; only their Gram
 * rows are formed (library scratch, n_rows x n_items floats) and only they are solved, every other target is returned
 * as the zero column.  *h_n_rows = number of candidate items; *h_used = 1 if the fit was done this way, 0 if the path
 * does not apply (cfg) or the candidates exceed a quarter of the catalogue -- outputs are then untouched and the caller
 * takes the dense path.  Matrix arguments as rt_gram_lower (+ d_ccol, the COO column id of every CSC entry), solver
 * arguments and outputs as rt_slim_solve.  Synchronises.
 */
int rt_slim_fit_pruned(int32_t n_users, int32_t n_items, const int32_t *d_cptr, const int32_t *d_cidx,
                       const float *d_cval, const int32_t *d_ccol, const int32_t *d_rptr, const int32_t *d_ridx,
                       const float *d_rval, int64_t nnz, const int32_t *d_targets, int32_t n_targets,
                       const rt_fit_config *cfg, const uint32_t *d_rng, int64_t rng_len, int64_t *d_out_off,
                       int32_t *d_out_cnt, int32_t *d_out_rows, float *d_out_vals, int64_t out_cap,
                       int64_t *h_needed, int32_t *d_stats, int32_t *h_used, int32_t *h_n_rows, void *stream);

/*
 * Assemble / merge the item-similarity matrix W (CSC, n_items x n_items) from solver output,
 * with the LIL-assignment semantics of slim_elastic.py:273-274, 371-374, 556-557:
 * for every returned pair (i, v) of target j: v != 0 sets W[i,j] = v, v == 0 deletes W[i,j];
 * entries of column j that were not returned keep their old (stale) value; other columns are
 * copied.  d_old_* may be NULL (n_old_items = 0) for a fresh matrix; an old matrix smaller than
 * n_items is resized (slim_elastic.py:326-327).  rows_sorted != 0 promises that each target's
 * returned rows are ascending (the nn == 0 solver output).  Output capacity out_cap; *h_nnz
 * receives the size (RT_ERR_CAPACITY if it does not fit).  Rows within a column are ascending.
 * Synchronises.
 */
int rt_w_merge(int32_t n_items, const int32_t *d_old_ptr, const int32_t *d_old_idx,
               const float *d_old_val, int32_t n_old_items, const int32_t *d_targets,
               int32_t n_targets, const int64_t *d_off, const int32_t *d_cnt, const int32_t *d_rows,
               const float *d_vals, int32_t rows_sorted, int32_t *d_wptr, int32_t *d_widx, float *d_wval,
               int64_t out_cap, int64_t *h_nnz, void *stream);

/* CSC -> CSR (or back) of a float32 matrix, minor indices ascending; used to give the scoring
 * kernel W by source item. */
int rt_transpose(int32_t n_major_in, int32_t n_major_out, const int32_t *d_ptr, const int32_t *d_idx,
                 const float *d_val, int64_t nnz, int32_t *d_optr, int32_t *d_oidx, float *d_oval,
                 void *stream);

/* ------------------------------------------------------------------------------------------
 * Scoring  (replaces SLIMElastic.recommend / recommend_batch / _dense_topk_indicies /
 * _sparse_topk_indicies, slim_elastic.py:628-818, and similar_items :820-857)
 * ---------------------------------------------------------------------------------------- */

#define RT_TOPK_DENSE 0  /* every item eligible (incl. score 0 / negative), slim_elastic.py:743-779 */
#define RT_TOPK_SPARSE 1 /* only non-zero scores eligible, slim_elastic.py:781-818 */

/*
 * For each of n_query users: s = X[u, :] . W restricted to item columns [j_begin, j_end)
 * (the shard this GPU owns; pass 0, n_items for all), interacted items removed when
 * filter_interacted, top-k by (score desc, item id desc).  W is given by SOURCE item (CSR of W:
 * row i lists (j, W[i,j]) ascending j).  Outputs int32 ids (-1 padded) and float32 scores
 * [n_query, k]; d_out_cnt[q] = number of valid entries.  1 <= k <= 128.
 */
int rt_slim_recommend(const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                      const int32_t *d_users, int32_t n_query, const int32_t *d_wrptr,
                      const int32_t *d_wridx, const float *d_wrval, int32_t n_items, int32_t j_begin,
                      int32_t j_end, int32_t k, int32_t filter_interacted, int32_t mode,
                      int32_t *d_out_ids, float *d_out_scores, int32_t *d_out_cnt, void *stream);

/*
 * Scoring pack (score3.cu): the heavy rows of W (source items with >= min_row entries inside
 * [j_begin, j_end)) re-packed per score tile as bank-striped ELL groups so that the shared-memory
 * score update is conflict free and each entry is one 8-byte load.  Built once per W and item range:
 *   rt_w_pack_plan  fills d_heavy_of[n_items] (heavy index or -1), d_heavy_list[n_items] (first
 *                   *h_n_heavy entries valid) and d_ell_off[n_heavy * n_tiles + 1] (capacity
 *                   n_items * n_tiles + 1 with n_tiles from rt_score_tile is always enough), returns the tile geometry and the number of
 *                   32-slot groups; synchronises.
 *   rt_w_pack_fill  writes the groups: d_ell holds *h_n_groups * 32 (byte offset of the column inside its
 *                   score tile, float bits) int32 pairs; padding slots carry value 0 and point at a
 *                   dummy float behind the tile; 8-byte aligned.
 * rt_slim_recommend_packed is rt_slim_recommend with the pack; results are identical bit for bit.
 * In rt_slim_recommend_packed only, d_users[q] < 0 marks an unused query slot: its answer is empty (count 0, ids -1).
 * Callers that re-score a device-computed list of users in a fixed number of slots (the hand-backs of
 * rt_slim_recommend_tc) use this to avoid a device->host read of the list length.
 */
int rt_score_tile(int32_t n_items, int32_t j_begin, int32_t j_end, int32_t *h_tile, int32_t *h_n_tiles);
int rt_w_pack_plan(const int32_t *d_wrptr, const int32_t *d_wridx, int32_t n_items, int32_t j_begin,
                   int32_t j_end, int32_t min_row, int32_t *d_heavy_of, int32_t *d_heavy_list,
                   int32_t *d_ell_off, int32_t *h_n_heavy, int32_t *h_tile, int32_t *h_n_tiles,
                   int64_t *h_n_groups, void *stream);
int rt_w_pack_fill(const int32_t *d_wrptr, const int32_t *d_wridx, const float *d_wrval, int32_t n_items,
                   int32_t j_begin, int32_t j_end, const int32_t *d_heavy_list, int32_t n_heavy,
                   const int32_t *d_ell_off, int32_t *d_ell, int64_t ell_cap_groups, void *stream);
int rt_slim_recommend_packed(const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                             const int32_t *d_users, int32_t n_query, const int32_t *d_wrptr,
                             const int32_t *d_wridx, const float *d_wrval, const int32_t *d_heavy_of,
                             const int32_t *d_ell_off, const int32_t *d_ell, int32_t n_items,
                             int32_t j_begin, int32_t j_end, int32_t k, int32_t filter_interacted,
                             int32_t mode, int32_t *d_out_ids, float *d_out_scores, int32_t *d_out_cnt,
                             void *stream);

/* Candidate-restricted variant (slim_elastic.py:661-672, 722-739): dense scores of the
 * n_cand candidate items only, NO interacted filter, order (score desc, candidate position desc).
 * d_out_pos receives positions into the candidate list. */
int rt_slim_recommend_candidates(const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                                 const int32_t *d_users, int32_t n_query, const int32_t *d_wptr,
                                 const int32_t *d_widx, const float *d_wval, int32_t n_items,
                                 const int32_t *d_cand, int32_t n_cand, int32_t k, int32_t *d_out_pos,
                                 float *d_out_scores, int32_t *d_out_cnt, void *stream);

/* Merge per-shard top-k lists: d_ids/d_scores are [n_shards, n_query, k] (as produced by an
 * all-gather of rt_slim_recommend outputs); result [n_query, k].  New; no reference analogue. */
int rt_topk_merge(const int32_t *d_ids, const float *d_scores, int32_t n_shards, int32_t n_query,
                  int32_t k, int32_t *d_out_ids, float *d_out_scores, int32_t *d_out_cnt, void *stream);

/* similar_items (slim_elastic.py:820-857): column j of W (CSC) minus the diagonal, top-k by
 * score desc (ties: ascending row).  Outputs [n_query, k]. */
int rt_slim_similar(const int32_t *d_wptr, const int32_t *d_widx, const float *d_wval, int32_t n_items,
                    const int32_t *d_items, int32_t n_query, int32_t k, int32_t *d_out_ids,
                    float *d_out_scores, int32_t *d_out_cnt, void *stream);

/*
 * Scoring with the heavy rows of W on the tensor cores (score_tc.cu; same contract as rt_slim_recommend_packed --
 * slim_elastic.py:674-818 -- over the WHOLE item range, for k <= 16, but scores agree with the fp32 reference sums only to
 * ~1e-6 relative: top-k lists match up to ties within tolerance).  Preconditions, checked by the caller: every stored value
 * of W and of X is >= 0 (a light contribution can then only raise a score), at most 64 heavy rows.
 *   rt_tc_pack_size / rt_tc_pack_build: per W, the heavy rows (d_heavy_list[n_heavy], ascending item id = heavy slot, as
 *     rt_w_pack_plan numbers them) as three K-major bf16 planes d_bt[3][i_pad][64] (w = w0 + w1 + w2; TMA source, 128-byte
 *     aligned) and as dense fp32 rows d_wd[n_heavy + 1][n_items] (the last row = column maxima over the heavy rows, the bound
 *     that lets the merge kernel drop most light cells after one load); *h_w_nonneg = 1 iff no stored value of W is negative.
 *     Synchronises.
 *   rt_values_bf16_exact: are all n values >= 0 / exactly representable in bf16 (integer and half-integer ratings are:
 *     one operand plane instead of three).  Synchronises.
 *   rt_slim_recommend_tc: x_planes = 1 or 3 (see above); d_tc_* [n_query, 32] / [n_query] receive the heavy-only candidate
 *     lists of the tensor-core kernel (two lists of <= k per query, one per column half of the accumulators, -1 padded), d_out_* the final lists (as rt_slim_recommend_packed), d_fallback[n_query] != 0
 *     marks queries the fast path could not finish (the caller re-scores them with rt_slim_recommend_packed);
 *     d_dbg_scores (optional, [n_query, i_pad]) receives every heavy-only score (tests).  Needs an sm_100 device.
 */
int rt_tc_pack_size(int32_t n_items, int32_t n_heavy, int32_t *h_i_pad, int64_t *h_bt_bytes, int64_t *h_wd_bytes);
int rt_tc_pack_build(const int32_t *d_wrptr, const int32_t *d_wridx, const float *d_wrval, int64_t w_nnz,
                     int32_t n_items, const int32_t *d_heavy_list, int32_t n_heavy, void *d_bt, float *d_wd,
                     int32_t *h_w_nonneg, void *stream);
int rt_values_bf16_exact(const float *d_vals, int64_t n, int32_t *h_nonneg, int32_t *h_bf16_exact, void *stream);
int rt_slim_recommend_tc(const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval, const int32_t *d_users,
                         int32_t n_query, const int32_t *d_wrptr, const int32_t *d_wridx, const float *d_wrval,
                         const int32_t *d_heavy_of, int32_t n_heavy, const void *d_bt, const float *d_wd, int32_t n_items,
                         int32_t k, int32_t filter_interacted, int32_t mode, int32_t x_planes, int32_t *d_tc_ids,
                         float *d_tc_scores, int32_t *d_tc_cnt, int32_t *d_out_ids, float *d_out_scores,
                         int32_t *d_out_cnt, int32_t *d_fallback, float *d_dbg_scores, void *stream);

/*
 * Host-side replay of LRUFreqSet.add (lru.py:33-47; one call per event with delta > 0, interactions.py:115-116) for a
 * batch in which evictions can occur: h_values[n] non-negative integer keys < key_bound in arrival order, the current set
 * as h_in_keys / h_in_counts[n_in] in least- to most-recently-used order; the new set comes back the same way
 * (h_out_* hold up to `capacity` entries).  Pure host code (no device work, no stream).
 */
int rt_lru_replay(const int64_t *h_values, int64_t n, int64_t capacity, int64_t key_bound,
                  const int64_t *h_in_keys, const int64_t *h_in_counts, int64_t n_in,
                  int64_t *h_out_keys, int64_t *h_out_counts, int64_t *h_n_out);

/*
 * Ranking metrics of Recommender.evaluate on device-resident top-k lists (replaces the per-user loop of
 * recommender.py:163-200 over metrics.py:6-313: precision, recall, f1, ndcg, hit, reciprocal rank, average precision,
 * true positives, AUC).  d_ids [n_query, k_stride] / d_cnt [n_query] = the lists as rt_slim_recommend* returns them;
 * ground truth of query q = d_gidx[d_gptr[q] .. d_gptr[q+1]) sorted ascending (duplicates count towards its length,
 * like len(ground_truth)); d_discount[i] = 1/log2(i+2) for i < max(recommend_size, 1) as computed by the caller.
 * compensated_sum != 0: `sum()` over floats is CPython >= 3.12's Neumaier summation, else plain.  d_out [n_query, 9]
 * float64 per-user values in the key order of compute_scores (precision, recall, f1, ndcg, hit_rate, mrr, map, tp,
 * auc); each equals the Python function's result bit for bit, the caller adds them up in user order.
 */
int rt_eval_metrics(const int32_t *d_ids, const int32_t *d_cnt, int32_t n_query, int32_t k_stride,
                    int32_t recommend_size, const int64_t *d_gptr, const int32_t *d_gidx,
                    const double *d_discount, int32_t compensated_sum, double *d_out, void *stream);

/* Tuning switches: "score_impl" for rt_slim_recommend (1 = first-generation scoring kernel,
 * 2 = staged/pipelined kernel, default 2; the packed third generation has its own entry point);
 * "gram_impl", "gram_slice", "gram_ranges", "gram_adapt" (variants of gram_lower_kernel: 0 = four unconditional
 * batches per rater, 1 = segment-length guards, 2 = guards + packed (relative index, value) entries, the default) for
 * the Gram kernels; "gram_head" (1 = the 2,048 most popular items' corner of rt_gram_lower goes to the tensor cores when the
 * values are exact in bf16, the sums exact in fp32 and the corner is dense enough -- gram_tc.cu --, 0 = never; default 1);
 * "solve_impl" for rt_slim_solve (1 = one CTA per target column for every configuration, 2 = one warp
 * per target column when nn <= 64, default 2; 3 = like 2 but every 7th target is handed to the CTA
 * kernel, a test hook for the overflow fallback).  Returns RT_ERR_ARG for an unknown name. */
int rt_set_option(const char *name, int32_t value);

/* rt_gram_finish_rowmax for a Gram matrix that only ONE bulk fit with feature selection will read: when cfg says that
 * targets without a live coordinate are skipped (nn > 0, skip_trivial, positive, nonneg), a row whose off-diagonal maximum
 * is not above alpha * l1_ratio * n_samples gets only its diagonal entry written -- the solver skips such a target
 * (rt_fit_config.rowmax_ptr) and, G being symmetric, the item can be nobody's live coordinate.  The other entries of
 * those rows keep whatever d_G held.  *h_live_only = 1 when rows were left out.  rt_slim_solve on this matrix must be
 * called with the same cfg (with rowmax_ptr = d_rowmax), without candidate lists in or out. */
int rt_gram_finish_live(int32_t n_items, float *d_Gp, int64_t ldgp, const int32_t *d_rank_of,
                        const int32_t *d_orig_of, float *d_G, int64_t ldg, float *d_rowmax,
                        const rt_fit_config *cfg, int32_t *h_has_rowmax, int32_t *h_live_only, void *stream);

/* Rows of the Gram matrix the last rt_gram_lower call of this process computed on the tensor cores (0 or 2048). */
int32_t rt_gram_last_head(void);

/* Frees the library-owned device scratch (grow-only arenas reused across calls). */
void rt_release_scratch(void);

/* Launch counters (how many kernels this library has launched since load / since reset). */
int64_t rt_launch_count(void);
void rt_launch_count_reset(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* RTREC_B200_H */
