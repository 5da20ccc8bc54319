// score.cu -- K6/K7/K8: scoring X[u,:] . W fused with interacted-item filtering and top-k,
// shard merge, and similar_items.
//
// Replaces /root/reference/rtrec/models/internal/slim_elastic.py:
//   recommend / recommend_batch          :628-741  (safe_sparse_dot + per-user python top-k)
//   _dense_topk_indicies                 :743-779  (RT_TOPK_DENSE)
//   _sparse_topk_indicies                :781-818  (RT_TOPK_SPARSE)
//   similar_items                        :820-857
//
// K6: one CTA per query user.  The user's score vector for the item tile lives in shared memory;
// for each interacted item i (CSR row of X) the CTA streams row i of W (W stored by source item)
// with coalesced loads and adds x_ui * W[i, j] into the tile.  Rows are applied one after the
// other (one barrier per item), which makes the fp32 sums deterministic: ascending i per target
// j, the same order scipy's csr_matmat uses.  Filtering and the exact top-k (radix select + rank
// sort, block_select.cuh) run on the tile while it is still in shared memory, so no score
// matrix is ever written to HBM.
//
// Algorithmic bytes per user: e*nnz(row u) + e*sum_{i in row u} nnz(W[i,:]) + 8k.
#include "block_select.cuh"
#include "common.cuh"

namespace rt {

constexpr int SCORE_NT = 256;
constexpr int KMAX = 128;

struct ScoreShared {
    SelectScratch sel;
    uint32_t cand_key[KMAX];
    int cand_idx[KMAX];
    int tile_idx[KMAX];
    uint32_t tile_key[KMAX];
    // running best over tiles (2*KMAX so a tile list can be appended before re-ranking)
    uint32_t best_key[2 * KMAX];
    int best_idx[2 * KMAX];
    uint32_t tmp_key[KMAX];
    int tmp_idx[KMAX];
    int q;
};

__device__ __forceinline__ float key_to_float(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__device__ __forceinline__ int lower_bound_i32(const int *a, int lo, int hi, int v) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void __launch_bounds__(SCORE_NT)
recommend_kernel(const int *__restrict__ rptr, const int *__restrict__ ridx, const float *__restrict__ rval,
                 const int *__restrict__ users, int n_query, const int *__restrict__ wrptr,
                 const int *__restrict__ wridx, const float *__restrict__ wrval, int n_items, int j_begin,
                 int j_end, int k, int filter, int mode, int tile, int *__restrict__ out_ids,
                 float *__restrict__ out_scores, int *__restrict__ out_cnt, int *__restrict__ next_query) {
    extern __shared__ __align__(16) float acc[];
    __shared__ ScoreShared sh;
    const int tid = threadIdx.x;
    const uint32_t KEY_NEG_INF = float_key(-INFINITY);
    for (;;) {
        __syncthreads();
        if (tid == 0) sh.q = atomicAdd(next_query, 1);
        __syncthreads();
        const int q = sh.q;
        if (q >= n_query) break;
        const int u = users[q];
        const int r0 = rptr[u], r1 = rptr[u + 1];
        int nbest = 0;
        for (int t0 = j_begin; t0 < j_end; t0 += tile) {
            const int t1 = min(t0 + tile, j_end);
            const int width = t1 - t0;
            for (int x = tid; x < width; x += SCORE_NT) acc[x] = 0.0f;
            __syncthreads();
            const bool whole = (t0 == 0 && t1 == n_items);
            for (int p = r0; p < r1; ++p) {
                const int i = ridx[p];
                const float x = rval[p];
                int a = wrptr[i], b = wrptr[i + 1];
                if (!whole && b > a) {
                    a = lower_bound_i32(wridx, a, b, t0);
                    b = lower_bound_i32(wridx, a, b, t1);
                }
                for (int e = a + tid; e < b; e += SCORE_NT) { float *d = &acc[wridx[e] - t0]; *d = __fadd_rn(*d, __fmul_rn(x, wrval[e])); }
                __syncthreads();
            }
            if (filter) {
                const int fa = lower_bound_i32(ridx, r0, r1, t0), fb = lower_bound_i32(ridx, r0, r1, t1);
                for (int p = fa + tid; p < fb; p += SCORE_NT) acc[ridx[p] - t0] = -INFINITY;
                __syncthreads();
            }
            auto key_of = [&](int idx) -> uint32_t { return float_key(acc[idx]); };
            auto elig = [&](int idx, uint32_t key) -> bool {
                if (key == KEY_NEG_INF) return false;
                if (mode == RT_TOPK_SPARSE) return acc[idx] != 0.0f;
                return true;
            };
            const int c = block_top_n(width, k, key_of, elig, &sh.sel, sh.cand_key, sh.cand_idx, sh.tile_idx, sh.tile_key);
            // append the tile's list (global item ids) to the running best and re-rank
            for (int e = tid; e < c; e += SCORE_NT) { sh.best_key[nbest + e] = sh.tile_key[e]; sh.best_idx[nbest + e] = sh.tile_idx[e] + t0; }
            __syncthreads();
            const int tot = nbest + c;
            if (t0 != j_begin) {
                for (int e = tid; e < tot; e += SCORE_NT) {
                    const uint32_t ke = sh.best_key[e];
                    const int ie = sh.best_idx[e];
                    int rank = 0;
                    for (int f = 0; f < tot; ++f) rank += (sh.best_key[f] > ke) || (sh.best_key[f] == ke && sh.best_idx[f] > ie);
                    if (rank < k) { sh.tmp_key[rank] = ke; sh.tmp_idx[rank] = ie; }
                }
                __syncthreads();
                nbest = min(tot, k);
                for (int e = tid; e < nbest; e += SCORE_NT) { sh.best_key[e] = sh.tmp_key[e]; sh.best_idx[e] = sh.tmp_idx[e]; }
                __syncthreads();
            } else nbest = c;
        }
        for (int e = tid; e < k; e += SCORE_NT) {
            out_ids[(size_t)q * k + e] = e < nbest ? sh.best_idx[e] : -1;
            out_scores[(size_t)q * k + e] = e < nbest ? key_to_float(sh.best_key[e]) : 0.0f;
        }
        if (tid == 0) out_cnt[q] = nbest;
    }
}

// candidate-restricted scoring (W given by target column, CSC): one CTA per query user
__global__ void __launch_bounds__(SCORE_NT)
recommend_cand_kernel(const int *__restrict__ rptr, const int *__restrict__ ridx, const float *__restrict__ rval,
                      const int *__restrict__ users, int n_query, const int *__restrict__ wptr,
                      const int *__restrict__ widx, const float *__restrict__ wval, const int *__restrict__ cand,
                      int n_cand, int k, float *__restrict__ score_buf, int *__restrict__ out_pos,
                      float *__restrict__ out_scores, int *__restrict__ out_cnt) {
    __shared__ ScoreShared sh;
    const int tid = threadIdx.x;
    float *sc = score_buf + (size_t)blockIdx.x * n_cand;
    for (int q = blockIdx.x; q < n_query; q += gridDim.x) {
        __syncthreads();
        const int u = users[q];
        const int r0 = rptr[u], r1 = rptr[u + 1];
        for (int c = tid; c < n_cand; c += SCORE_NT) {
            const int j = cand[c];
            float s = 0.0f;
            for (int e = wptr[j]; e < wptr[j + 1]; ++e) {
                const int i = widx[e];
                const int p = lower_bound_i32(ridx, r0, r1, i);
                if (p < r1 && ridx[p] == i) s = __fadd_rn(s, __fmul_rn(rval[p], wval[e]));
            }
            sc[c] = s;
        }
        __syncthreads();
        auto key_of = [&](int idx) -> uint32_t { return float_key(sc[idx]); };
        auto elig = [&](int, uint32_t) -> bool { return true; };
        const int c = block_top_n(n_cand, k, key_of, elig, &sh.sel, sh.cand_key, sh.cand_idx, sh.tile_idx, sh.tile_key);
        for (int e = tid; e < k; e += SCORE_NT) {
            out_pos[(size_t)q * k + e] = e < c ? sh.tile_idx[e] : -1;
            out_scores[(size_t)q * k + e] = e < c ? key_to_float(sh.tile_key[e]) : 0.0f;
        }
        if (tid == 0) out_cnt[q] = c;
    }
}

// K7: merge per-shard lists; one warp per query
__global__ void topk_merge_kernel(const int *__restrict__ ids, const float *__restrict__ scores, int n_shards,
                                  int n_query, int k, int *__restrict__ out_ids, float *__restrict__ out_scores,
                                  int *__restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= n_query) return;
    const int tot = n_shards * k;
    int placed = 0;
    for (int e = lane; e < tot; e += 32) {
        const int s = e / k, r = e - s * k;
        const size_t pe = ((size_t)s * n_query + q) * k + r;
        const int ie = ids[pe];
        if (ie < 0) continue;
        const uint32_t ke = float_key(scores[pe]);
        int rank = 0;
        for (int f = 0; f < tot; ++f) {
            const int s2 = f / k, r2 = f - s2 * k;
            const size_t pf = ((size_t)s2 * n_query + q) * k + r2;
            const int jf = ids[pf];
            if (jf < 0) continue;
            const uint32_t kf = float_key(scores[pf]);
            rank += (kf > ke) || (kf == ke && jf > ie);
        }
        if (rank < k) { out_ids[(size_t)q * k + rank] = ie; out_scores[(size_t)q * k + rank] = scores[pe]; }
        ++placed;
    }
    placed = warp_sum_i(placed);
    const int cnt = min(placed, k);
    for (int e = cnt + lane; e < k; e += 32) { out_ids[(size_t)q * k + e] = -1; out_scores[(size_t)q * k + e] = 0.0f; }
    if (lane == 0) out_cnt[q] = cnt;
}

// K8: similar items; one warp per query item
__global__ void similar_kernel(const int *__restrict__ wptr, const int *__restrict__ widx,
                               const float *__restrict__ wval, int n_items, const int *__restrict__ items,
                               int n_query, int k, int *__restrict__ out_ids, float *__restrict__ out_scores,
                               int *__restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= n_query) return;
    const int j = items[q];
    int a = 0, b = 0;
    if (j >= 0 && j < n_items) { a = wptr[j]; b = wptr[j + 1]; }
    int valid = 0;
    for (int e = a + lane; e < b; e += 32) {
        const int ie = widx[e];
        if (ie == j) continue;
        ++valid;
        const float ve = wval[e];
        int rank = 0;
        for (int f = a; f < b; ++f) {
            if (widx[f] == j) continue;
            const float vf = wval[f];
            rank += (vf > ve) || (vf == ve && f < e);
        }
        if (rank < k) { out_ids[(size_t)q * k + rank] = ie; out_scores[(size_t)q * k + rank] = ve; }
    }
    valid = warp_sum_i(valid);
    const int cnt = min(valid, k);
    for (int e = cnt + lane; e < k; e += 32) { out_ids[(size_t)q * k + e] = -1; out_scores[(size_t)q * k + e] = 0.0f; }
    if (lane == 0) out_cnt[q] = cnt;
}

}  // namespace rt

using namespace rt;

extern "C" int rt_slim_recommend(const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                                 const int32_t *d_users, int32_t n_query, const int32_t *d_wrptr,
                                 const int32_t *d_wridx, const float *d_wrval, int32_t n_items, int32_t j_begin,
                                 int32_t j_end, int32_t k, int32_t filter_interacted, int32_t mode,
                                 int32_t *d_out_ids, float *d_out_scores, int32_t *d_out_cnt, void *stream) {
    RT_ARG(k >= 1 && k <= KMAX, "k must be in [1,128]");
    RT_ARG(n_items > 0 && j_begin >= 0 && j_end <= n_items && j_begin <= j_end, "item range");
    RT_ARG(mode == RT_TOPK_DENSE || mode == RT_TOPK_SPARSE, "mode");
    if (n_query <= 0) return RT_OK;
    RT_ARG(d_rptr && d_users && d_wrptr && d_out_ids && d_out_scores && d_out_cnt, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int *d_next = (int *)rt::scratch(SCR_MISC, 256);
    if (!d_next) return RT_ERR_CUDA;
    RT_CUDA(cudaMemsetAsync(d_next, 0, sizeof(int), st));
    const int width = j_end - j_begin;
    if (width == 0) {
        RT_CUDA(cudaMemsetAsync(d_out_cnt, 0, sizeof(int) * (size_t)n_query, st));
        RT_CUDA(cudaMemsetAsync(d_out_ids, 0xff, sizeof(int) * (size_t)n_query * k, st));
        RT_CUDA(cudaMemsetAsync(d_out_scores, 0, sizeof(float) * (size_t)n_query * k, st));
        return RT_OK;
    }
    if (rt::option(rt::OPT_SCORE_IMPL) != 1)  // 1 = first-generation kernel below (kept for A/B timing)
        return rt_launch_recommend2(d_rptr, d_ridx, d_rval, d_users, n_query, d_wrptr, d_wridx, d_wrval, n_items, j_begin,
                                    j_end, k, filter_interacted, mode, d_out_ids, d_out_scores, d_out_cnt, d_next, st);
    // tile: as much of the item range as fits next to the static shared memory; prefer two CTAs/SM
    const int optin = rt::smem_optin();
    const int static_bytes = (int)sizeof(ScoreShared) + 1024;
    int max_floats_1 = (optin - static_bytes) / 4;
    int max_floats_2 = (optin / 2 - static_bytes - 1024) / 4;
    int tile = width;
    if (width > max_floats_1) {
        const int ntiles = (width + max_floats_2 - 1) / max_floats_2;
        tile = (width + ntiles - 1) / ntiles;
    }
    tile = (tile + 3) & ~3;
    const size_t smem = (size_t)tile * sizeof(float);
    RT_CUDA(cudaFuncSetAttribute(recommend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((size_t)optin / (smem + static_bytes));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int grid = rt::sm_count() * per_sm;
    if (grid > n_query) grid = n_query;
    recommend_kernel<<<grid, SCORE_NT, smem, st>>>(d_rptr, d_ridx, d_rval, d_users, n_query, d_wrptr, d_wridx, d_wrval,
                                                  n_items, j_begin, j_end, k, filter_interacted, mode, tile, d_out_ids,
                                                  d_out_scores, d_out_cnt, d_next);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_slim_recommend_candidates(const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                                            const int32_t *d_users, int32_t n_query, const int32_t *d_wptr,
                                            const int32_t *d_widx, const float *d_wval, int32_t n_items,
                                            const int32_t *d_cand, int32_t n_cand, int32_t k, int32_t *d_out_pos,
                                            float *d_out_scores, int32_t *d_out_cnt, void *stream) {
    RT_ARG(k >= 1 && k <= KMAX, "k must be in [1,128]");
    RT_ARG(n_items > 0 && n_cand > 0, "shape");
    if (n_query <= 0) return RT_OK;
    RT_ARG(d_rptr && d_users && d_wptr && d_cand && d_out_pos && d_out_scores && d_out_cnt, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int grid = rt::sm_count() * 4;
    if (grid > n_query) grid = n_query;
    float *buf = (float *)rt::scratch(SCR_MISC, 256 + sizeof(float) * (size_t)grid * n_cand);
    if (!buf) return RT_ERR_CUDA;
    recommend_cand_kernel<<<grid, SCORE_NT, 0, st>>>(d_rptr, d_ridx, d_rval, d_users, n_query, d_wptr, d_widx, d_wval,
                                                    d_cand, n_cand, k, buf + 64, d_out_pos, d_out_scores, d_out_cnt);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_topk_merge(const int32_t *d_ids, const float *d_scores, int32_t n_shards, int32_t n_query,
                             int32_t k, int32_t *d_out_ids, float *d_out_scores, int32_t *d_out_cnt, void *stream) {
    RT_ARG(k >= 1 && k <= KMAX && n_shards >= 1, "k / n_shards");
    if (n_query <= 0) return RT_OK;
    RT_ARG(d_ids && d_scores && d_out_ids && d_out_scores && d_out_cnt, "null pointer");
    const int bs = 128;
    const unsigned grid = (unsigned)(((int64_t)n_query * 32 + bs - 1) / bs);
    topk_merge_kernel<<<grid, bs, 0, (cudaStream_t)stream>>>(d_ids, d_scores, n_shards, n_query, k, d_out_ids,
                                                             d_out_scores, d_out_cnt);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_slim_similar(const int32_t *d_wptr, const int32_t *d_widx, const float *d_wval, int32_t n_items,
                               const int32_t *d_items, int32_t n_query, int32_t k, int32_t *d_out_ids,
                               float *d_out_scores, int32_t *d_out_cnt, void *stream) {
    RT_ARG(k >= 1, "k");
    if (n_query <= 0) return RT_OK;
    RT_ARG(d_wptr && d_items && d_out_ids && d_out_scores && d_out_cnt, "null pointer");
    const int bs = 128;
    const unsigned grid = (unsigned)(((int64_t)n_query * 32 + bs - 1) / bs);
    similar_kernel<<<grid, bs, 0, (cudaStream_t)stream>>>(d_wptr, d_widx, d_wval, n_items, d_items, n_query, k,
                                                          d_out_ids, d_out_scores, d_out_cnt);
    RT_CHECK_LAUNCH();
    return RT_OK;
}
