// score3.cu -- K6 (v3): scoring with a bank-striped ELL pack for the heavy rows of W.
//
// Replaces /root/reference/rtrec/models/internal/slim_elastic.py:674-741 (recommend_batch:
// safe_sparse_dot + per-user python top-k), :743-779 and :781-818, like score.cu / score2.cu, with
// the same results bit for bit (same fp32 summation order per score: ascending source item).
//
// Why a third generation (profiles/r1g_*, r1l_*): at the ML-20M shape 99 % of the multiply-adds of
// X.W come from the ~50 most popular source items, whose W rows hold thousands of entries (they are
// a nearest neighbour of a quarter of all targets).  v2 streams those rows as (column, value) CSR
// pairs and scatters into the shared-memory score tile: two global loads per entry and a
// read-modify-write whose 32 lanes hit random banks (~3.5-way conflicts measured: 1.0e9 conflicts per
// launch).  v3 re-packs every heavy row once per W, per score tile, as ELL groups of 32 slots where
// slot `lane` only ever holds a column with (column - tile start) mod 32 == lane:
//     * the tile update of a group is bank-conflict free by construction,
//     * one 8-byte load per entry ((column, value) interleaved) instead of two 4-byte loads,
//     * groups of a row are independent, so every warp keeps eight loads in flight and applies them as one
//       batch (read-all / add / write-all, no LDS -> STS chain).
// Rows are still applied one after the other in ascending item order (one barrier per row), light
// rows exactly as in v2, so each score is the same fp32 sum scipy's csr_matmat produces.
//
// Algorithmic bytes per user: e*nnz(row u) + e*sum_{i in row u} nnz(W[i,:]) + 8k.
#include <cub/cub.cuh>

#include "block_select.cuh"
#include "common.cuh"

namespace rt {

constexpr int S3_NT = 512;      // threads per CTA (2 CTAs/SM at the ML-20M tile size)
constexpr int S3_NW = S3_NT / 32;
constexpr int S3_CH = 512;      // interacted items staged per chunk
constexpr int S3_GROUP = 8;     // light rows whose head entries are fetched together
constexpr int KMAX3 = 128;

struct Score3Shared {
    union {
        struct { int a[S3_CH]; int b[S3_CH]; float x[S3_CH]; } st;    // staging (b < 0: heavy row, a = first group, ~b = end group)
        FastSelScratch fs;                                             // top-k
        struct { uint32_t key[2 * KMAX3]; int idx[2 * KMAX3]; } tmp;   // tile merge
    } u;
    uint32_t best_key[2 * KMAX3];
    int best_idx[2 * KMAX3];
    int wsum[32];
    int n_rows, q;
};

__device__ __forceinline__ float key_to_float3(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__device__ __forceinline__ int lower_bound3(const int *__restrict__ a, int lo, int hi, int v) {
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// ------------------------------------------------------------------------------------------------
// pack construction
// ------------------------------------------------------------------------------------------------

// heavy flag per source item: entries inside [j_begin, j_end) >= min_row
__global__ void pack_flag_kernel(const int *__restrict__ wrptr, const int *__restrict__ wridx, int n_items, int j_begin,
                                 int j_end, int min_row, int *__restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    int a = wrptr[i], b = wrptr[i + 1];
    if (b - a >= min_row && (j_begin > 0 || j_end < n_items)) {
        a = lower_bound3(wridx, a, b, j_begin);
        b = lower_bound3(wridx, a, b, j_end);
    }
    flag[i] = (b - a >= min_row) ? 1 : 0;
}

// heavy_of[i] = index among heavy rows or -1; heavy_list[h] = i
__global__ void pack_index_kernel(const int *__restrict__ flag, const int *__restrict__ pos, int n_items,
                                  int *__restrict__ heavy_of, int *__restrict__ heavy_list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    if (flag[i]) { heavy_of[i] = pos[i]; heavy_list[pos[i]] = i; }
    else heavy_of[i] = -1;
}

// one warp per (heavy row, tile): number of 32-slot groups = largest per-bank population
__global__ void pack_count_kernel(const int *__restrict__ wrptr, const int *__restrict__ wridx,
                                  const int *__restrict__ heavy_list, int n_heavy, int n_tiles, int j_begin, int j_end,
                                  int tile, int *__restrict__ n_groups) {
    const int lane = threadIdx.x & 31;
    const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (w >= n_heavy * n_tiles) return;
    const int h = w / n_tiles, tau = w - h * n_tiles;
    const int i = heavy_list[h];
    const int t0 = j_begin + tau * tile, t1 = min(t0 + tile, j_end);
    int a = lower_bound3(wridx, wrptr[i], wrptr[i + 1], t0);
    const int b = lower_bound3(wridx, a, wrptr[i + 1], t1);
    int mine = 0;  // population of bank `lane`
    for (int base = a; base < b; base += 32) {
        const int e = base + lane;
        const int bank = e < b ? ((wridx[e] - t0) & 31) : -1;
#pragma unroll
        for (int l = 0; l < 32; ++l) mine += (__shfl_sync(0xffffffffu, bank, l) == lane);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine = max(mine, __shfl_xor_sync(0xffffffffu, mine, o));
    // rounded up to one group per warp of the scoring CTA: every warp then runs the same number of batches (the
    // extra groups are padding slots)
    if (lane == 0) n_groups[w] = (mine + S3_NW - 1) / S3_NW * S3_NW;
}

// one warp per (heavy row, tile): slot (position within bank, bank) <- (byte offset of the column in the score
// tile, value); a padding slot points at the lane's own dummy float behind the tile with value 0, so the
// scoring loop needs no predicate
__global__ void pack_fill_kernel(const int *__restrict__ wrptr, const int *__restrict__ wridx, const float *__restrict__ wrval,
                                 const int *__restrict__ heavy_list, int n_heavy, int n_tiles, int j_begin, int j_end,
                                 int tile, const int *__restrict__ ell_off, int2 *__restrict__ ell) {
    const int lane = threadIdx.x & 31;
    const int w = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (w >= n_heavy * n_tiles) return;
    const int h = w / n_tiles, tau = w - h * n_tiles;
    const int i = heavy_list[h];
    const int t0 = j_begin + tau * tile, t1 = min(t0 + tile, j_end);
    int a = lower_bound3(wridx, wrptr[i], wrptr[i + 1], t0);
    const int b = lower_bound3(wridx, a, wrptr[i + 1], t1);
    const int g0 = ell_off[w], g1 = ell_off[w + 1];
    int2 *out = ell + (size_t)g0 * 32;
    for (int s = lane; s < (g1 - g0) * 32; s += 32) out[s] = make_int2((tile + lane) * 4, 0);
    __syncwarp();
    int filled = 0;  // entries already placed in bank `lane`
    for (int base = a; base < b; base += 32) {
        const int e = base + lane;
        int col = -1, bank = -1;
        float v = 0.f;
        if (e < b) { col = wridx[e] - t0; bank = col & 31; v = wrval[e]; }
        // position inside the bank: entries of the same bank earlier in this batch + earlier batches
        const unsigned same = __match_any_sync(0xffffffffu, bank);
        const int before = __popc(same & ((1u << lane) - 1u));
        const int prior = __shfl_sync(0xffffffffu, filled, bank < 0 ? 0 : bank);
        if (e < b) out[(size_t)(prior + before) * 32 + bank] = make_int2(col * 4, __float_as_int(v));
        // update the per-bank counters: bank `lane` gains the number of lanes whose bank == lane
        int gain = 0;
#pragma unroll
        for (int l = 0; l < 32; ++l) gain += (__shfl_sync(0xffffffffu, bank, l) == lane);
        filled += gain;
    }
}

// ------------------------------------------------------------------------------------------------
// scoring
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(S3_NT, 2)
recommend3_kernel(const int *__restrict__ rptr, const int *__restrict__ ridx, const float *__restrict__ rval,
                  const int *__restrict__ users, int n_query, const int *__restrict__ wrptr,
                  const int *__restrict__ wridx, const float *__restrict__ wrval, const int *__restrict__ heavy_of,
                  const int *__restrict__ ell_off, const int2 *__restrict__ ell, int n_tiles, int n_items, int j_begin,
                  int j_end, int k, int filter, int mode, int tile, int *__restrict__ out_ids,
                  float *__restrict__ out_scores, int *__restrict__ out_cnt, int *__restrict__ next_query,
                  const int *__restrict__ order, const int *__restrict__ n_active_queries) {
    extern __shared__ __align__(16) float acc[];
    __shared__ Score3Shared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // sparse mode: queries without a single W entry behind their items were answered by query_work_kernel and sit at the
    // end of order[]; only the first *n_active_queries are scored here
    if (n_active_queries) n_query = min(n_query, *n_active_queries);   // (the launcher passes the count of dense-tile queries)
    for (;;) {
        __syncthreads();
        if (tid == 0) sh.q = atomicAdd(next_query, 1);
        __syncthreads();
        if (sh.q >= n_query) break;
        // queries are taken longest row first (order[] from the launcher): the dynamic queue then ends with the
        // cheap users and no CTA is left alone with an expensive one
        const int q = order ? order[sh.q] : sh.q;
        const int u = users[q];
        if (u < 0) {   // an unused query slot (see query_work_kernel): empty answer
            for (int e = tid; e < k; e += S3_NT) { out_ids[(size_t)q * k + e] = -1; out_scores[(size_t)q * k + e] = 0.0f; }
            if (tid == 0) out_cnt[q] = 0;
            continue;
        }
        const int r0 = rptr[u], r1 = rptr[u + 1];
        int nbest = 0;
        int tau = 0;
        for (int t0 = j_begin; t0 < j_end; t0 += tile, ++tau) {
            const int t1 = min(t0 + tile, j_end);
            const int width = t1 - t0;
            const bool whole = (t0 == 0 && t1 == n_items);
            // the tile is zeroed when the first row with entries in it turns up: in sparse mode a tile nothing was added to
            // holds no candidate and is skipped without being touched (a W with few columns, or a user whose items have
            // no neighbours in this item range)
            bool zeroed = false;
            auto zero_tile = [&]() {   // tile starts are 16-byte aligned (tile is a multiple of 32 floats)
                const int w4 = width >> 2;
                for (int x = tid; x < w4; x += S3_NT) reinterpret_cast<float4 *>(acc)[x] = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int x = (w4 << 2) + tid; x < width; x += S3_NT) acc[x] = 0.0f;
                __syncthreads();
                zeroed = true;
            };
            // ---------------- accumulate: chunks of the user's row ----------------
            for (int c0 = r0; c0 < r1; c0 += S3_CH) {
                __syncthreads();  // previous chunk fully applied (and acc zeroed) before st.* is rewritten
                {
                    const int p = c0 + tid;
                    int a = 0, b = 0;
                    float x = 0.f;
                    bool ne = false;
                    if (p < r1 && tid < S3_CH) {
                        const int i = ridx[p];
                        x = rval[p];
                        const int h = heavy_of[i];
                        if (h >= 0) {
                            a = ell_off[h * n_tiles + tau];
                            const int ge = ell_off[h * n_tiles + tau + 1];
                            ne = ge > a;
                            b = ~ge;  // negative marks a heavy row
                        } else {
                            a = wrptr[i]; b = wrptr[i + 1];
                            if (!whole && b > a) {
                                a = lower_bound3(wridx, a, b, t0);
                                b = lower_bound3(wridx, a, b, t1);
                            }
                            ne = b > a;
                        }
                    }
                    const unsigned bal = __ballot_sync(0xffffffffu, ne);
                    if (lane == 0) sh.wsum[warp] = __popc(bal);
                    __syncthreads();
                    int off = 0, tot = 0;
#pragma unroll
                    for (int w = 0; w < S3_NW; ++w) { const int c = sh.wsum[w]; if (w < warp) off += c; tot += c; }
                    if (ne) {
                        const int s = off + __popc(bal & ((1u << lane) - 1u));
                        sh.u.st.a[s] = a; sh.u.st.b[s] = b; sh.u.st.x[s] = x;
                    }
                    if (tid == 0) sh.n_rows = tot;
                    __syncthreads();
                }
                const int n = sh.n_rows;
                if (n > 0 && !zeroed) zero_tile();
                int s = 0;
                while (s < n) {   // every branch below is uniform across the CTA (the staged list is shared)
                    if (sh.u.st.b[s] < 0) {
                        // ---- heavy row: warp w takes groups a+w, a+w+16, ... (every warp the same number: group counts
                        // are multiples of 16, pack_count_kernel), eight loads in flight, then 4 / 2 / 1 for the rest.
                        // All tile reads of a batch are issued before its first write: the slots of one lane within a
                        // row are distinct columns (or the lane's own dummy float, which only ever receives + x*0), so
                        // a batch carries no dependency; it is written read-all / add / write-all because the compiler
                        // must otherwise assume aliasing and chains LDS -> FADD -> STS per entry (the short-scoreboard
                        // stall of profiles/r1z_*: 16.5 -> 14.2 ms).  Prefetching the next row's first batch across
                        // the row barrier was measured as well (profiles/r2a_*: 14.35 ms) and dropped.
                        {
                            const int2 *base = ell + lane;
                            char *accb = (char *)acc;
                            const float x = sh.u.st.x[s];
                            const int ge = ~sh.u.st.b[s];
                            int gi = sh.u.st.a[s] + warp;
#define S3_BATCH(NB)                                                                                       \
                            {                                                                              \
                                int2 e[NB];                                                                \
                                float v[NB];                                                               \
                                _Pragma("unroll") for (int r = 0; r < NB; ++r) e[r] = base[(size_t)(gi + r * S3_NW) * 32]; \
                                _Pragma("unroll") for (int r = 0; r < NB; ++r) v[r] = *(const float *)(accb + e[r].x);     \
                                _Pragma("unroll") for (int r = 0; r < NB; ++r) v[r] = __fadd_rn(v[r], __fmul_rn(x, __int_as_float(e[r].y))); \
                                _Pragma("unroll") for (int r = 0; r < NB; ++r) *(float *)(accb + e[r].x) = v[r];           \
                            }
                            for (; gi + 7 * S3_NW < ge; gi += 8 * S3_NW) S3_BATCH(8)
                            if (gi + 3 * S3_NW < ge) { S3_BATCH(4) gi += 4 * S3_NW; }
                            if (gi + S3_NW < ge) { S3_BATCH(2) gi += 2 * S3_NW; }
                            if (gi < ge) S3_BATCH(1)
#undef S3_BATCH
                            __syncthreads();
                            ++s;
                        }
                    } else {
                        // ---- up to 8 consecutive light rows: the heads (first 512 entries) are fetched together, all
                        // loads independent and issued back to back, then applied row by row
                        int hj[S3_GROUP];
                        float hv[S3_GROUP];
                        int cnt = 0;
#pragma unroll
                        for (int t = 0; t < S3_GROUP; ++t) {
                            hj[t] = -1; hv[t] = 0.f;
                            if (cnt == t && s + t < n) {
                                const int b = sh.u.st.b[s + t];
                                if (b >= 0) {
                                    ++cnt;
                                    const int e = sh.u.st.a[s + t] + tid;
                                    if (e < b) { hj[t] = wridx[e]; hv[t] = wrval[e]; }
                                }
                            }
                        }
#pragma unroll
                        for (int t = 0; t < S3_GROUP; ++t) {
                            if (t < cnt) {
                                const float x = sh.u.st.x[s + t];
                                const int b = sh.u.st.b[s + t];
                                if (hj[t] >= 0) { float *d = &acc[hj[t] - t0]; *d = __fadd_rn(*d, __fmul_rn(x, hv[t])); }
                                for (int e = sh.u.st.a[s + t] + tid + S3_NT; e < b; e += S3_NT) {
                                    float *d = &acc[wridx[e] - t0]; *d = __fadd_rn(*d, __fmul_rn(x, wrval[e]));
                                }
                                __syncthreads();
                            }
                        }
                        s += cnt;
                    }
                }
            }
            __syncthreads();
            if (!zeroed) {
                if (mode == RT_TOPK_SPARSE) continue;   // nothing scored in this tile: no candidates (CTA-uniform)
                zero_tile();
            }
            if (filter) {
                int fa = r0, fb = r1;
                if (!whole) { fa = lower_bound3(ridx, r0, r1, t0); fb = lower_bound3(ridx, r0, r1, t1); }
                for (int p = fa + tid; p < fb; p += S3_NT) acc[ridx[p] - t0] = -INFINITY;
                __syncthreads();
            }
            // ---------------- top-k of the tile ----------------
            const int c = block_top_n_fast_f32(acc, width, k, mode == RT_TOPK_SPARSE, &sh.u.fs, sh.best_idx + nbest,
                                               sh.best_key + nbest);
            if (t0 != 0) {
                for (int e = tid; e < c; e += S3_NT) sh.best_idx[nbest + e] += t0;
                __syncthreads();
            }
            const int tot = nbest + c;
            if (t0 != j_begin && c > 0) {
                for (int e = tid; e < tot; e += S3_NT) {
                    const uint32_t ke = sh.best_key[e];
                    const int ie = sh.best_idx[e];
                    int rank = 0;
                    for (int f = 0; f < tot; ++f) rank += (sh.best_key[f] > ke) || (sh.best_key[f] == ke && sh.best_idx[f] > ie);
                    if (rank < k) { sh.u.tmp.key[rank] = ke; sh.u.tmp.idx[rank] = ie; }
                }
                __syncthreads();
                nbest = min(tot, k);
                for (int e = tid; e < nbest; e += S3_NT) { sh.best_key[e] = sh.u.tmp.key[e]; sh.best_idx[e] = sh.u.tmp.idx[e]; }
                __syncthreads();
            } else nbest = tot;
        }
        for (int e = tid; e < k; e += S3_NT) {
            out_ids[(size_t)q * k + e] = e < nbest ? sh.best_idx[e] : -1;
            out_scores[(size_t)q * k + e] = e < nbest ? key_to_float3(sh.best_key[e]) : 0.0f;
        }
        if (tid == 0) out_cnt[q] = nbest;
    }
}

// sort key of a query: its scoring work = stored W entries behind the items of its row (one warp per query).  In
// sparse mode a query with no work has an empty answer (no non-zero score): it is written here and never reaches the
// scoring kernel (n_active counts the others).
__global__ void __launch_bounds__(256) query_work_kernel(const int *__restrict__ rptr, const int *__restrict__ ridx,
                                                         const int *__restrict__ users, int n_query,
                                                         const int *__restrict__ wrptr, int j_begin, int j_end, int n_items,
                                                         const int *__restrict__ wridx, int k, int sparse,
                                                         unsigned *__restrict__ keys, int *__restrict__ idx,
                                                         int *__restrict__ n_active, int sparse_cap, int *__restrict__ out_ids,
                                                         float *__restrict__ out_scores, int *__restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const int q = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (q >= n_query) return;
    const int u = users[q];
    const bool whole = j_begin == 0 && j_end == n_items;
    unsigned long long work = 0;
    // a negative user id marks an unused query slot (fixed-size re-scoring batches of the tensor-core path, whose real
    // length only the device knows): answered here with an empty list, in both modes
    const int p_end = u < 0 ? 0 : rptr[u + 1];
    for (int p = (u < 0 ? 0 : rptr[u]) + lane; p < p_end; p += 32) {
        const int i = ridx[p];
        int a = wrptr[i], b = wrptr[i + 1];
        if (!whole && b > a) { a = lower_bound3(wridx, a, b, j_begin); b = lower_bound3(wridx, a, b, j_end); }
        work += (unsigned)(b - a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) work += __shfl_xor_sync(0xffffffffu, work, o);
    const bool skip = (sparse && work == 0) || u < 0;
    if (skip) {
        for (int e = lane; e < k; e += 32) { out_ids[(size_t)q * k + e] = -1; out_scores[(size_t)q * k + e] = 0.0f; }
    }
    if (lane == 0) {
        // key: work, +1 for every scored query so that only answered ones carry key 0 (they sort to the end)
        keys[q] = skip ? 0u : (unsigned)min(work + 1ull, 0xffffffffull);
        idx[q] = q;
        // n_active[0] = queries with work, n_active[1] = those among them that go to the dense-tile kernel (the rest, with
        // at most sparse_cap entries behind their items, are scored by recommend_sparse_kernel)
        if (skip) out_cnt[q] = 0;
        else {
            atomicAdd(n_active, 1);
            if (sparse_cap < 0 || work > (unsigned long long)sparse_cap) atomicAdd(n_active + 1, 1);   // < 0: no sparse kernel
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Queries with little work (sparse top-k semantics only): the scores of a user whose items have at most SP_CAP stored W
// entries behind them touch at most SP_CAP items, so a dense score tile (zeroing and scanning tens of thousands of floats
// per user and item range) is the wrong structure.  One warp per query accumulates into a small open-addressing table
// in shared memory.  Rows of W are applied one after the other in ascending item order with separate fp32 multiply and
// add -- the first touch of an item stores x*w (= 0 + x*w) -- so every score is bit-identical to the dense-tile kernels'
// and to scipy's csr_matmat sum.  Selection: k rounds of a warp arg-max with the usual order (score desc, item id desc).
// Serves W matrices with few columns (all-features fits at the H&M shape end with a handful of entries) and the long
// tail of users with a few ratings.
constexpr int SP_SLOTS = 2048;
constexpr int SP_CAP = 1024;
constexpr int SP_WARPS = 4;
constexpr int SP_KMAX = 32;

__device__ __forceinline__ int sp_hash(int j) { return (int)(((unsigned)j * 2654435761u) >> 21) & (SP_SLOTS - 1); }

__global__ void __launch_bounds__(SP_WARPS * 32)
recommend_sparse_kernel(const int *__restrict__ rptr, const int *__restrict__ ridx, const float *__restrict__ rval,
                        const int *__restrict__ users, const int *__restrict__ order, const int *__restrict__ n_active,
                        const int *__restrict__ wrptr, const int *__restrict__ wridx, const float *__restrict__ wrval,
                        int n_items, int j_begin, int j_end, int k, int filter, int *__restrict__ out_ids,
                        float *__restrict__ out_scores, int *__restrict__ out_cnt, int *__restrict__ next_query) {
    extern __shared__ __align__(16) int sp_smem[];   // per warp: SP_SLOTS keys, then SP_SLOTS values
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int *key = sp_smem + (size_t)warp * 2 * SP_SLOTS;
    float *val = reinterpret_cast<float *>(key + SP_SLOTS);
    const int q_end = n_active[0], q_begin = n_active[1];   // order[q_begin .. q_end): the queries of this kernel
    const bool whole = j_begin == 0 && j_end == n_items;
    for (;;) {
        int qi = 0;
        if (lane == 0) qi = q_begin + atomicAdd(next_query, 1);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= q_end) break;
        const int q = order[qi];
        const int u = users[q];
        const int r0 = rptr[u], r1 = rptr[u + 1];
        // ---- at most 32 entries behind a row of at most 32 items: everything in registers (no table to clear and scan)
        if (r1 - r0 <= 32) {
            const int p = r0 + lane;
            int a = 0, b = 0, item = -1;
            float x = 0.f;
            if (p < r1) {
                item = ridx[p];
                x = rval[p];
                a = wrptr[item]; b = wrptr[item + 1];
                if (!whole && b > a) { a = lower_bound3(wridx, a, b, j_begin); b = lower_bound3(wridx, a, b, j_end); }
            }
            int incl = b - a;                                   // inclusive prefix of the entry counts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t_ = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t_; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            if (total <= 32) {
                // entry e (ascending item, then ascending column) -> lane e, staged through the (unused) table memory
                for (int e = 0; e < b - a; ++e) { key[incl - (b - a) + e] = wridx[a + e]; val[incl - (b - a) + e] = __fmul_rn(x, wrval[a + e]); }
                __syncwarp();
                const int j = lane < total ? key[lane] : -1 - lane;      // distinct negative ids for the unused lanes
                const float v = lane < total ? val[lane] : 0.0f;
                // score of column j = its entries added in ascending source-item order, starting from 0 (scipy's order);
                // the first lane holding j keeps it
                float acc = 0.0f;
                bool first = true, interacted = false;
#pragma unroll 4
                for (int l = 0; l < 32; ++l) {
                    const int jl = __shfl_sync(0xffffffffu, j, l);
                    const float vl = __shfl_sync(0xffffffffu, v, l);
                    const int il = __shfl_sync(0xffffffffu, item, l);
                    if (jl == j) { acc = __fadd_rn(acc, vl); if (l < lane) first = false; }
                    if (filter && il == j) interacted = true;
                }
                __syncwarp();
                const bool cand = lane < total && first && !interacted && acc != 0.0f;
                const uint32_t fk = cand ? float_key(acc) : 0u;
                int rank = 0;
#pragma unroll 4
                for (int l = 0; l < 32; ++l) {
                    const uint32_t kl = __shfl_sync(0xffffffffu, fk, l);
                    const int jl = __shfl_sync(0xffffffffu, j, l);
                    const bool cl = __shfl_sync(0xffffffffu, (int)cand, l) != 0;
                    rank += cl && (kl > fk || (kl == fk && jl > j));
                }
                if (cand && rank < k) { out_ids[(size_t)q * k + rank] = j; out_scores[(size_t)q * k + rank] = acc; }
                const int cnt = min(k, __popc(__ballot_sync(0xffffffffu, cand)));
                for (int e = cnt + lane; e < k; e += 32) { out_ids[(size_t)q * k + e] = -1; out_scores[(size_t)q * k + e] = 0.0f; }
                if (lane == 0) out_cnt[q] = cnt;
                __syncwarp();
                continue;
            }
        }
        for (int s = lane; s < SP_SLOTS; s += 32) key[s] = -1;
        __syncwarp();
        // ---- accumulate, row by row in ascending item order
        for (int c0 = r0; c0 < r1; c0 += 32) {
            const int p = c0 + lane;
            int a = 0, b = 0;
            float x = 0.f;
            if (p < r1) {
                const int i = ridx[p];
                x = rval[p];
                a = wrptr[i]; b = wrptr[i + 1];
                if (!whole && b > a) { a = lower_bound3(wridx, a, b, j_begin); b = lower_bound3(wridx, a, b, j_end); }
            }
            unsigned todo = __ballot_sync(0xffffffffu, b > a);
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                const int al = __shfl_sync(0xffffffffu, a, l), bl = __shfl_sync(0xffffffffu, b, l);
                const float xl = __shfl_sync(0xffffffffu, x, l);
                for (int e = al + lane; e < bl; e += 32) {
                    const int j = wridx[e];
                    const float add = __fmul_rn(xl, wrval[e]);
                    int slot = sp_hash(j);
                    int prev;
                    for (;;) {
                        prev = atomicCAS(&key[slot], -1, j);
                        if (prev == -1 || prev == j) break;
                        slot = (slot + 1) & (SP_SLOTS - 1);
                    }
                    // items of one W row are distinct, so no two lanes update the same slot between two __syncwarp()s
                    val[slot] = prev == -1 ? __fadd_rn(0.0f, add) : __fadd_rn(val[slot], add);
                }
                __syncwarp();
            }
        }
        // ---- interacted items leave the table
        if (filter) {
            for (int p = r0 + lane; p < r1; p += 32) {
                const int i = ridx[p];
                if (i < j_begin || i >= j_end) continue;
                int slot = sp_hash(i);
                for (;;) {
                    const int kk = key[slot];
                    if (kk == -1) break;
                    if (kk == i) { val[slot] = 0.0f; break; }   // a zero score is not eligible in sparse mode
                    slot = (slot + 1) & (SP_SLOTS - 1);
                }
            }
            __syncwarp();
        }
        // ---- top-k: k rounds of arg-max over the table, order (score desc, item id desc); taken entries are cleared
        int cnt = 0;
        for (int round = 0; round < k; ++round) {
            uint32_t bk = 0u;
            int bi = -1, bs = -1;
            for (int s = lane; s < SP_SLOTS; s += 32) {
                const int kk = key[s];
                if (kk < 0) continue;
                const float v = val[s];
                if (v == 0.0f) continue;
                const uint32_t fk = float_key(v);
                if (fk > bk || (fk == bk && kk > bi)) { bk = fk; bi = kk; bs = s; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const uint32_t ok = __shfl_xor_sync(0xffffffffu, bk, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                const int os = __shfl_xor_sync(0xffffffffu, bs, o);
                if (ok > bk || (ok == bk && oi > bi)) { bk = ok; bi = oi; bs = os; }
            }
            if (bi < 0) break;
            if (lane == 0) {
                out_ids[(size_t)q * k + round] = bi;
                out_scores[(size_t)q * k + round] = key_to_float3(bk);
                val[bs] = 0.0f;
            }
            ++cnt;
            __syncwarp();
        }
        for (int e = cnt + lane; e < k; e += 32) { out_ids[(size_t)q * k + e] = -1; out_scores[(size_t)q * k + e] = 0.0f; }
        if (lane == 0) out_cnt[q] = cnt;
        __syncwarp();
    }
}

// tile geometry shared by the pack builder and the launcher: two CTAs per SM share the shared memory
static void score3_geometry(int width, int *tile, int *n_tiles) {
    const int optin = rt::smem_optin();
    const int static_bytes = (int)sizeof(Score3Shared) + 1024 + 64;
    const int max_floats_2 = ((optin + 1024) / 2 - static_bytes) / 4 - 32;  // 32 dummy floats behind the tile
    int t = width, nt = 1;
    if (width > max_floats_2) {
        nt = (width + max_floats_2 - 1) / max_floats_2;
        t = (width + nt - 1) / nt;
    }
    t = (t + 31) & ~31;  // bank alignment of every tile start
    nt = width > 0 ? (width + t - 1) / t : 1;
    *tile = t; *n_tiles = nt;
}

}  // namespace rt

using namespace rt;

#define S3_CUB(call_expr)                                                                          \
    do {                                                                                           \
        size_t tmp_bytes__ = 0;                                                                    \
        void *d_tmp__ = nullptr;                                                                   \
        RT_CUDA(call_expr);                                                                        \
        d_tmp__ = rt::scratch(SCR_CUB, tmp_bytes__);                                               \
        if (!d_tmp__) return RT_ERR_CUDA;                                                          \
        RT_CUDA(call_expr);                                                                        \
        rt::count_launch(2);                                                                       \
    } while (0)

extern "C" int rt_score_tile(int32_t n_items, int32_t j_begin, int32_t j_end, int32_t *h_tile, int32_t *h_n_tiles) {
    RT_ARG(n_items > 0 && j_begin >= 0 && j_end <= n_items && j_begin <= j_end && h_tile && h_n_tiles, "item range");
    int tile = 0, n_tiles = 1;
    score3_geometry(j_end - j_begin, &tile, &n_tiles);
    *h_tile = tile; *h_n_tiles = n_tiles;
    return RT_OK;
}

extern "C" int rt_w_pack_plan(const int32_t *d_wrptr, const int32_t *d_wridx, int32_t n_items, int32_t j_begin,
                              int32_t j_end, int32_t min_row, int32_t *d_heavy_of, int32_t *d_heavy_list,
                              int32_t *d_ell_off, int32_t *h_n_heavy, int32_t *h_tile, int32_t *h_n_tiles,
                              int64_t *h_n_groups, void *stream) {
    RT_ARG(n_items > 0 && j_begin >= 0 && j_end <= n_items && j_begin <= j_end, "item range");
    RT_ARG(d_wrptr && d_heavy_of && d_heavy_list && d_ell_off, "null pointer");
    RT_ARG(h_n_heavy && h_tile && h_n_tiles && h_n_groups, "host outputs");
    if (min_row < 32) min_row = 32;
    cudaStream_t st = (cudaStream_t)stream;
    int tile = 0, n_tiles = 1;
    score3_geometry(j_end - j_begin, &tile, &n_tiles);
    *h_tile = tile; *h_n_tiles = n_tiles; *h_n_heavy = 0; *h_n_groups = 0;
    const int bs = 256;
    int *flag = (int *)rt::scratch(SCR_WMAT_A, sizeof(int) * (2 * (size_t)n_items + 64));
    if (!flag) return RT_ERR_CUDA;
    int *pos = flag + n_items + 32;
    pack_flag_kernel<<<(n_items + bs - 1) / bs, bs, 0, st>>>(d_wrptr, d_wridx, n_items, j_begin, j_end, min_row, flag);
    RT_CHECK_LAUNCH();
    S3_CUB(cub::DeviceScan::ExclusiveSum(d_tmp__, tmp_bytes__, flag, pos, n_items, st));
    pack_index_kernel<<<(n_items + bs - 1) / bs, bs, 0, st>>>(flag, pos, n_items, d_heavy_of, d_heavy_list);
    RT_CHECK_LAUNCH();
    int last[2] = {0, 0};
    RT_CUDA(cudaMemcpyAsync(&last[0], flag + n_items - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaMemcpyAsync(&last[1], pos + n_items - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    const int n_heavy = last[0] + last[1];
    *h_n_heavy = n_heavy;
    RT_CUDA(cudaMemsetAsync(d_ell_off, 0, sizeof(int), st));
    if (n_heavy == 0 || j_end == j_begin) return RT_OK;
    const int n_w = n_heavy * n_tiles;
    int *cnt = (int *)rt::scratch(SCR_WMAT_B, sizeof(int) * ((size_t)n_w + 64));
    if (!cnt) return RT_ERR_CUDA;
    pack_count_kernel<<<(unsigned)(((int64_t)n_w * 32 + bs - 1) / bs), bs, 0, st>>>(d_wrptr, d_wridx, d_heavy_list, n_heavy, n_tiles,
                                                                                    j_begin, j_end, tile, cnt);
    RT_CHECK_LAUNCH();
    RT_CUDA(cudaMemsetAsync(cnt + n_w, 0, sizeof(int), st));
    S3_CUB(cub::DeviceScan::ExclusiveSum(d_tmp__, tmp_bytes__, cnt, d_ell_off, n_w + 1, st));
    int total = 0;
    RT_CUDA(cudaMemcpyAsync(&total, d_ell_off + n_w, sizeof(int), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    *h_n_groups = total;
    return RT_OK;
}

extern "C" int rt_w_pack_fill(const int32_t *d_wrptr, const int32_t *d_wridx, const float *d_wrval, int32_t n_items,
                              int32_t j_begin, int32_t j_end, const int32_t *d_heavy_list, int32_t n_heavy,
                              const int32_t *d_ell_off, int32_t *d_ell, int64_t ell_cap_groups, void *stream) {
    RT_ARG(n_items > 0 && j_begin >= 0 && j_end <= n_items && j_begin <= j_end, "item range");
    if (n_heavy <= 0) return RT_OK;
    RT_ARG(d_wrptr && d_wridx && d_wrval && d_heavy_list && d_ell_off && d_ell && ell_cap_groups > 0, "null pointer");
    RT_ARG((((uintptr_t)d_ell) & 7) == 0, "d_ell must be 8-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int tile = 0, n_tiles = 1;
    score3_geometry(j_end - j_begin, &tile, &n_tiles);
    const int n_w = n_heavy * n_tiles, bs = 256;
    pack_fill_kernel<<<(unsigned)(((int64_t)n_w * 32 + bs - 1) / bs), bs, 0, st>>>(d_wrptr, d_wridx, d_wrval, d_heavy_list, n_heavy,
                                                                                   n_tiles, j_begin, j_end, tile, d_ell_off,
                                                                                   (int2 *)d_ell);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_slim_recommend_packed(const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                                        const int32_t *d_users, int32_t n_query, const int32_t *d_wrptr,
                                        const int32_t *d_wridx, const float *d_wrval, const int32_t *d_heavy_of,
                                        const int32_t *d_ell_off, const int32_t *d_ell, int32_t n_items,
                                        int32_t j_begin, int32_t j_end, int32_t k, int32_t filter_interacted,
                                        int32_t mode, int32_t *d_out_ids, float *d_out_scores, int32_t *d_out_cnt,
                                        void *stream) {
    RT_ARG(k >= 1 && k <= KMAX3, "k must be in [1,128]");
    RT_ARG(n_items > 0 && j_begin >= 0 && j_end <= n_items && j_begin < j_end, "item range");
    RT_ARG(mode == RT_TOPK_DENSE || mode == RT_TOPK_SPARSE, "mode");
    if (n_query <= 0) return RT_OK;
    RT_ARG(d_rptr && d_users && d_wrptr && d_heavy_of && d_ell_off && d_out_ids && d_out_scores && d_out_cnt, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int *d_next = (int *)rt::scratch(SCR_MISC, 256);
    if (!d_next) return RT_ERR_CUDA;
    RT_CUDA(cudaMemsetAsync(d_next, 0, sizeof(int), st));
    int tile = 0, n_tiles = 1;
    score3_geometry(j_end - j_begin, &tile, &n_tiles);
    const int optin = rt::smem_optin();
    const int static_bytes = (int)sizeof(Score3Shared) + 1024 + 64;
    const size_t smem = (size_t)(tile + 32) * sizeof(float);
    RT_CUDA(cudaFuncSetAttribute(recommend3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((size_t)(optin + 1024) / (smem + static_bytes));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    int grid = rt::sm_count() * per_sm;
    if (grid > n_query) grid = n_query;
    const int *d_order = nullptr, *d_n_active = nullptr;
    bool use_sparse = false;
    if (n_query > 4 * grid) {
        // heaviest-query-first processing order; in sparse mode queries without work are answered on the spot
        const size_t nq = (size_t)n_query;
        unsigned *keys = (unsigned *)rt::scratch(SCR_SCORE, (4 * nq + 256) * sizeof(int));
        if (!keys) return RT_ERR_CUDA;
        unsigned *keys2 = keys + nq + 32;
        int *idx = (int *)(keys2 + nq + 32), *idx2 = idx + nq + 32;
        int *d_cnt_active = d_next + 8;            // [0] queries with work, [1] dense-tile queries among them
        RT_CUDA(cudaMemsetAsync(d_next, 0, 64, st));
        // the sparse-table kernel takes queries with little work when the semantics allow it (sparse top-k, k <= 32)
        use_sparse = mode == RT_TOPK_SPARSE && k <= SP_KMAX && rt::option(rt::OPT_SCORE_IMPL) != 1;
        query_work_kernel<<<(unsigned)(((int64_t)n_query * 32 + 255) / 256), 256, 0, st>>>(
            d_rptr, d_ridx, d_users, n_query, d_wrptr, j_begin, j_end, n_items, d_wridx, k, mode == RT_TOPK_SPARSE ? 1 : 0, keys, idx,
            d_cnt_active, use_sparse ? SP_CAP : -1, d_out_ids, d_out_scores, d_out_cnt);
        RT_CHECK_LAUNCH();
        S3_CUB(cub::DeviceRadixSort::SortPairsDescending(d_tmp__, tmp_bytes__, keys, keys2, idx, idx2, n_query, 0, 32, st));
        d_order = idx2;
        d_n_active = d_cnt_active + 1;
        if (use_sparse) {
            const int sgrid = rt::sm_count() * 3;
            const size_t ssmem = (size_t)SP_WARPS * 2 * SP_SLOTS * sizeof(int);
            RT_CUDA(cudaFuncSetAttribute(recommend_sparse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
            recommend_sparse_kernel<<<sgrid, SP_WARPS * 32, ssmem, st>>>(d_rptr, d_ridx, d_rval, d_users, d_order, d_cnt_active, d_wrptr,
                                                                   d_wridx, d_wrval, n_items, j_begin, j_end, k, filter_interacted,
                                                                   d_out_ids, d_out_scores, d_out_cnt, d_next + 4);
            RT_CHECK_LAUNCH();
        }
    }
    recommend3_kernel<<<grid, S3_NT, smem, st>>>(
        d_rptr, d_ridx, d_rval, d_users, n_query, d_wrptr, d_wridx, d_wrval, d_heavy_of, d_ell_off, (const int2 *)d_ell, n_tiles,
        n_items, j_begin, j_end, k, filter_interacted, mode, tile, d_out_ids, d_out_scores, d_out_cnt, d_next, d_order, d_n_active);
    RT_CHECK_LAUNCH();
    return RT_OK;
}
