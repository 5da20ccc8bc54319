// score2.cu -- K6 (v2): scoring X[u,:] . W fused with interacted-item filtering and top-k.
//
// Replaces /root/reference/rtrec/models/internal/slim_elastic.py:674-741 (recommend_batch:
// safe_sparse_dot + per-user python top-k), :743-779 (_dense_topk_indicies) and :781-818
// (_sparse_topk_indicies).  Same results as the v1 kernel in score.cu (same summation order, same
// (score desc, item id desc) order); what changed is where the time went (profiles/r1a_*):
//
//   * v1 issued three dependent global loads and one barrier per interacted item.  v2 first stages
//     the user's row -- (row bounds of W, rating) for every interacted item whose W row is not
//     empty in this tile -- into shared memory with independent coalesced loads, then walks the
//     staged rows in groups of 8: the first S2_NT entries of all 8 rows are fetched up front
//     (16 loads in flight per thread), then applied row after row with one barrier each.  Rows are
//     still applied strictly in ascending item order, so every score is the same fp32 sum, in the
//     order scipy's csr_matmat produces it.
//   * v1 ran a 4-pass radix select over all n_items scores.  v2 takes a lower bound T on the k-th
//     best score from 128 strided bucket maxima (the k-th largest bucket maximum: at least k
//     scores are >= T), collects the few scores >= T, and rank-sorts them.  Two light passes over
//     the tile instead of six heavy ones.  If the candidate list overflows (massive ties), the
//     exact radix select of block_select.cuh runs instead, so the result is always exact.
//
// Algorithmic bytes per user: e*nnz(row u) + e*sum_{i in row u} nnz(W[i,:]) + 8k.
#include "block_select.cuh"
#include "common.cuh"

namespace rt {

constexpr int S2_NT = 512;      // threads per CTA (2 CTAs/SM at the ML-20M tile size)
constexpr int S2_CH = 512;      // interacted items staged per chunk
constexpr int S2_GROUP = 8;     // rows whose head entries are fetched together
constexpr int KMAX2 = 128;

struct Score2Shared {
    union {
        struct { int a[S2_CH]; int b[S2_CH]; float x[S2_CH]; } st;    // staging
        FastSelScratch fs;                                             // top-k
        struct { uint32_t key[2 * KMAX2]; int idx[2 * KMAX2]; } tmp;   // tile merge
    } u;
    uint32_t best_key[2 * KMAX2];
    int best_idx[2 * KMAX2];
    int wsum[32];
    int n_rows, q;
};

__device__ __forceinline__ float key_to_float2(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__device__ __forceinline__ int lower_bound2(const int *__restrict__ a, int lo, int hi, int v) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// key of an eligible score, 0 for an ineligible one (float_key never returns 0 for a real number)
__device__ __forceinline__ uint32_t elig_key(float v, int mode) {
    if (v == -INFINITY) return 0u;
    if (mode == RT_TOPK_SPARSE && v == 0.0f) return 0u;
    return float_key(v);
}

__global__ void __launch_bounds__(S2_NT, 2)
recommend2_kernel(const int *__restrict__ rptr, const int *__restrict__ ridx, const float *__restrict__ rval,
                  const int *__restrict__ users, int n_query, const int *__restrict__ wrptr,
                  const int *__restrict__ wridx, const float *__restrict__ wrval, int n_items, int j_begin,
                  int j_end, int k, int filter, int mode, int tile, int *__restrict__ out_ids,
                  float *__restrict__ out_scores, int *__restrict__ out_cnt, int *__restrict__ next_query) {
    extern __shared__ __align__(16) float acc[];
    __shared__ Score2Shared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = S2_NT / 32;
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
            const bool whole = (t0 == 0 && t1 == n_items);
            for (int x = tid; x < width; x += S2_NT) acc[x] = 0.0f;
            // ---------------- accumulate: chunks of the user's row ----------------
            for (int c0 = r0; c0 < r1; c0 += S2_CH) {
                __syncthreads();  // previous chunk fully applied (and acc zeroed) before st.* is rewritten
                // stage (ordered compaction of the items whose W row is non-empty in this tile)
                {
                    const int p = c0 + tid;
                    int a = 0, b = 0;
                    float x = 0.f;
                    if (p < r1 && tid < S2_CH) {
                        const int i = ridx[p];
                        x = rval[p];
                        a = wrptr[i]; b = wrptr[i + 1];
                        if (!whole && b > a) {
                            a = lower_bound2(wridx, a, b, t0);
                            b = lower_bound2(wridx, a, b, t1);
                        }
                    }
                    const bool ne = b > a;
                    const unsigned bal = __ballot_sync(0xffffffffu, ne);
                    if (lane == 0) sh.wsum[warp] = __popc(bal);
                    __syncthreads();
                    int off = 0, tot = 0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) { const int c = sh.wsum[w]; if (w < warp) off += c; tot += c; }
                    if (ne) {
                        const int s = off + __popc(bal & ((1u << lane) - 1u));
                        sh.u.st.a[s] = a; sh.u.st.b[s] = b; sh.u.st.x[s] = x;
                    }
                    if (tid == 0) sh.n_rows = tot;
                    __syncthreads();
                }
                const int n = sh.n_rows;
                for (int g = 0; g < n; g += S2_GROUP) {
                    int hj[S2_GROUP];
                    float hv[S2_GROUP];
                    // heads of up to 8 rows: all loads are independent and issued back to back
#pragma unroll
                    for (int s = 0; s < S2_GROUP; ++s) {
                        hj[s] = -1; hv[s] = 0.f;
                        if (g + s < n) {
                            const int e = sh.u.st.a[g + s] + tid;
                            if (e < sh.u.st.b[g + s]) { hj[s] = wridx[e]; hv[s] = wrval[e]; }
                        }
                    }
#pragma unroll
                    for (int s = 0; s < S2_GROUP; ++s) {
                        if (g + s < n) {   // uniform across the CTA
                            const float x = sh.u.st.x[g + s];
                            if (hj[s] >= 0) { float *d = &acc[hj[s] - t0]; *d = __fadd_rn(*d, __fmul_rn(x, hv[s])); }
                            const int b = sh.u.st.b[g + s];
                            for (int e = sh.u.st.a[g + s] + tid + S2_NT; e < b; e += S2_NT) {
                                float *d = &acc[wridx[e] - t0]; *d = __fadd_rn(*d, __fmul_rn(x, wrval[e]));
                            }
                            __syncthreads();
                        }
                    }
                }
            }
            __syncthreads();
            if (filter) {
                int fa = r0, fb = r1;
                if (!whole) { fa = lower_bound2(ridx, r0, r1, t0); fb = lower_bound2(ridx, r0, r1, t1); }
                for (int p = fa + tid; p < fb; p += S2_NT) acc[ridx[p] - t0] = -INFINITY;
                __syncthreads();
            }
            // ---------------- top-k of the tile ----------------
            auto key_of = [&](int idx) -> uint32_t { return elig_key(acc[idx], mode); };
            // the tile's list is appended to best[] (capacity 2k); indices become global item ids
            const int c = block_top_n_fast(width, k, key_of, &sh.u.fs, sh.best_idx + nbest, sh.best_key + nbest);
            if (t0 != 0) {
                for (int e = tid; e < c; e += S2_NT) sh.best_idx[nbest + e] += t0;
                __syncthreads();
            }
            // merge with the running best of earlier tiles
            const int tot = nbest + c;
            if (t0 != j_begin && c > 0) {
                for (int e = tid; e < tot; e += S2_NT) {
                    const uint32_t ke = sh.best_key[e];
                    const int ie = sh.best_idx[e];
                    int rank = 0;
                    for (int f = 0; f < tot; ++f) rank += (sh.best_key[f] > ke) || (sh.best_key[f] == ke && sh.best_idx[f] > ie);
                    if (rank < k) { sh.u.tmp.key[rank] = ke; sh.u.tmp.idx[rank] = ie; }
                }
                __syncthreads();
                nbest = min(tot, k);
                for (int e = tid; e < nbest; e += S2_NT) { sh.best_key[e] = sh.u.tmp.key[e]; sh.best_idx[e] = sh.u.tmp.idx[e]; }
                __syncthreads();
            } else nbest = tot;
        }
        for (int e = tid; e < k; e += S2_NT) {
            out_ids[(size_t)q * k + e] = e < nbest ? sh.best_idx[e] : -1;
            out_scores[(size_t)q * k + e] = e < nbest ? key_to_float2(sh.best_key[e]) : 0.0f;
        }
        if (tid == 0) out_cnt[q] = nbest;
    }
}

}  // namespace rt

using namespace rt;

// launch helper used by rt_slim_recommend (score.cu)
int rt_launch_recommend2(const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval, const int32_t *d_users,
                         int32_t n_query, const int32_t *d_wrptr, const int32_t *d_wridx, const float *d_wrval,
                         int32_t n_items, int32_t j_begin, int32_t j_end, int32_t k, int32_t filter_interacted,
                         int32_t mode, int32_t *d_out_ids, float *d_out_scores, int32_t *d_out_cnt, int *d_next,
                         cudaStream_t st) {
    const int width = j_end - j_begin;
    const int optin = rt::smem_optin();
    const int static_bytes = (int)sizeof(Score2Shared) + 1024 + 64;
    const int max_floats_1 = (optin - static_bytes) / 4;
    // two CTAs per SM share 228 KB; each also pays 1 KB of reserved shared memory
    const int max_floats_2 = ((optin + 1024) / 2 - static_bytes) / 4;
    int tile = width;
    if (width > max_floats_2) {
        if (width <= max_floats_1 && n_query <= rt::sm_count()) tile = width;  // few users: one pass, 1 CTA/SM
        else {
            const int ntiles = (width + max_floats_2 - 1) / max_floats_2;
            tile = (width + ntiles - 1) / ntiles;
        }
    }
    tile = (tile + 3) & ~3;
    const size_t smem = (size_t)tile * sizeof(float);
    RT_CUDA(cudaFuncSetAttribute(recommend2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((size_t)(optin + 1024) / (smem + static_bytes));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    int grid = rt::sm_count() * per_sm;
    if (grid > n_query) grid = n_query;
    recommend2_kernel<<<grid, S2_NT, smem, st>>>(d_rptr, d_ridx, d_rval, d_users, n_query, d_wrptr, d_wridx, d_wrval,
                                                n_items, j_begin, j_end, k, filter_interacted, mode, tile, d_out_ids,
                                                d_out_scores, d_out_cnt, d_next);
    RT_CHECK_LAUNCH();
    return RT_OK;
}
