// score_tc.cu -- K6 (v4): scoring with the heavy rows of W on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces /root/reference/rtrec/models/internal/slim_elastic.py:674-741 (recommend_batch: safe_sparse_dot + per-user
// top-k), :743-779 and :781-818 for large batches, like score3.cu, but NOT bit for bit: scores agree with the fp32
// reference sums to ~1e-6 relative, so top-k lists agree up to ties within tolerance (north_star section 6; the exact
// kernels of score3.cu remain the default for small batches, for item-sharded scoring and whenever a precondition fails).
//
// Why (profiles/r2c_ml20m_recommend3.md): at the ML-20M shape 99 % of the multiply-adds of X.W come from ~50
// source items whose W rows hold thousands of entries.  That part of the product is a dense contraction
//     S_h[U x I] = X_h[U x 64] . W_h[64 x I]
// which the exact kernels execute as 1.2e10 shared-memory read-modify-writes (L1/TEX bound, 14.3 ms).  Here:
//   * W_h is stored once per W as three bf16 planes (w = w0 + w1 + w2, 24 mantissa bits), K-major, and streamed by TMA
//     (cp.async.bulk.tensor, 128-byte swizzle) through a 3-stage mbarrier ring;
//   * X_h of 128 users is gathered from the CSR rows into a swizzled shared-memory operand (values that are exact in
//     bf16 -- integer and half-integer ratings -- need one plane, decayed values three);
//   * tcgen05.mma (M=128, N=128, K=16, fp32 accumulate) forms the split products x0w0 + x0w1 + x0w2 (+ x1w0 + x1w1 +
//     x2w0) into one of four TMEM accumulators;
//   * the epilogue warps read the accumulators with tcgen05.ld, drop the user's interacted items (a bit mask built from
//     the row cursor) and keep a threshold top-k per TMEM lane -- the score matrix never exists anywhere.
// The light rows of W (1 % of the multiply-adds, but they touch arbitrary cells) are added by recommend_tcfix_kernel on
// the CUDA cores: with W >= 0 and X >= 0 a light contribution can only raise a score, so the final top-k is contained in
// (heavy-only top-k) U (cells touched by light rows); those cells are accumulated in a shared-memory hash table (64-bit
// fixed point: deterministic whatever the order), completed with their heavy part from a dense fp32 copy of W_h, and
// merged with the tensor-core list.  Users whose table would overflow, or who end with fewer than k positive scores in
// dense mode, are flagged and re-scored by the exact kernel.
//
// Algorithmic bytes per user: e*nnz(row u) + e*sum_{i in row u} nnz(W[i,:]) + 8k.
#include "tc_common.cuh"

namespace rt {

constexpr int TC_M = 128;        // users per CTA tile (= TMEM lanes)
constexpr int TC_N = 128;        // items per accumulator tile (= TMEM columns per buffer)
constexpr int TC_KH = 64;        // heavy rows (K of one operand tile: 64 bf16 = one 128-byte swizzle row)
constexpr int TC_STAGES = 3;     // W_h ring
constexpr int TC_ACC = 4;        // TMEM accumulator buffers (4 x 128 = 512 columns)
constexpr int TC_KLIST = 16;     // longest list kept per user
constexpr int TC_EPI_WARPS = 8;  // two warps per TMEM lane quadrant: each takes half of the columns of every accumulator
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;   // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-9: operand gather + epilogue
constexpr int TC_OUT = 2 * TC_KLIST;   // candidates written per user (one list per column half)
constexpr int TC_TILE_BYTES = TC_N * TC_KH * 2;   // one [128 rows x 64 bf16] operand tile

struct TcParams {
    const int *rptr, *ridx;
    const float *rval;
    const int *users;
    int n_query;
    const int *heavy_of;     // item -> heavy slot (< TC_KH) or -1
    int n_tiles, i_pad;      // item tiles of TC_N, padded item count
    int k, filter;
    int *out_ids;            // [n_query, TC_OUT] heavy-only candidates (two unsorted lists of <= k, -1 padded)
    float *out_scores;
    int *out_cnt;
    float *dbg;              // optional [n_query, i_pad]: every heavy-only score (tests)
};

// byte offset of element (row, k) inside a swizzled operand tile (Swizzle<3,4,3>: 16-byte chunk index ^= row % 8)
__device__ __forceinline__ int tc_sw_off(int row, int k) {
    const int chunk = (k >> 3) ^ (row & 7);
    return row * 128 + chunk * 16 + (k & 7) * 2;
}

// threshold list of one user: unsorted, smallest member = thr once the list holds k entries.  Returns the new (n, thr)
// by value so that both stay in registers at the call sites.
struct TcListState { int n; float thr; };
__device__ __noinline__ TcListState tc_insert(float v, int j, float *ls, int *li, int n, float thr, int k) {
    if (n < k) {
        ls[n * TC_M] = v; li[n * TC_M] = j;
        ++n;
        if (n < k) return {n, thr};
    } else {
        int s = 0;
        while (s < k - 1 && ls[s * TC_M] != thr) ++s;
        ls[s * TC_M] = v; li[s * TC_M] = j;
    }
    float m = ls[0];
    for (int s = 1; s < k; ++s) m = fminf(m, ls[s * TC_M]);
    return {n, m};
}

template <int SX>
__global__ void __launch_bounds__(TC_THREADS, 1) recommend_tc_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams P) {
    extern __shared__ unsigned char tc_smem_raw[];
    // operand tiles need 1024-byte alignment (swizzle atom); everything else lives behind them
    unsigned char *smem = tc_smem_raw + ((1024u - (smem_u32(tc_smem_raw) & 1023u)) & 1023u);
    unsigned char *sA = smem;                                         // SX tiles
    unsigned char *sB = sA + SX * TC_TILE_BYTES;                      // TC_STAGES x 3 tiles
    float *list_s = reinterpret_cast<float *>(sB + TC_STAGES * 3 * TC_TILE_BYTES);   // [half][slot][row]
    int *list_i = reinterpret_cast<int *>(list_s + 2 * TC_KLIST * TC_M);
    uint64_t *bars = reinterpret_cast<uint64_t *>(list_i + 2 * TC_KLIST * TC_M);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);
    const uint32_t bar0 = smem_u32(bars);
    auto b_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto b_empty = [&](int s) { return bar0 + 8u * (uint32_t)(TC_STAGES + s); };
    auto acc_full = [&](int b) { return bar0 + 8u * (uint32_t)(2 * TC_STAGES + b); };
    auto acc_empty = [&](int b) { return bar0 + 8u * (uint32_t)(2 * TC_STAGES + TC_ACC + b); };
    const uint32_t a_full = bar0 + 8u * (uint32_t)(2 * TC_STAGES + 2 * TC_ACC);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int b = 0; b < TC_ACC; ++b) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), TC_EPI_WARPS); }
        mbar_init(a_full, TC_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int n_blocks = (P.n_query + TC_M - 1) / TC_M;
    const int NT = P.n_tiles;
    int it = 0;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++it) {
        if (warp == 0) {
            // ===== TMA producer: the three bf16 planes of one item tile per stage
            if (lane == 0) {
                for (int t = 0; t < NT; ++t) {
                    const int g = it * NT + t, s = g % TC_STAGES;
                    mbar_wait(b_empty(s), ((g / TC_STAGES) & 1) ^ 1);
                    mbar_arrive_expect_tx(b_full(s), 3 * TC_TILE_BYTES);
#pragma unroll
                    for (int pl = 0; pl < 3; ++pl)
                        tma_load_2d(smem_u32(sB + (s * 3 + pl) * TC_TILE_BYTES), &tmap, b_full(s), 0, pl * P.i_pad + t * TC_N);
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer (one thread)
            if (lane == 0) {
                mbar_wait(a_full, it & 1);
                tc_fence_after();
                for (int t = 0; t < NT; ++t) {
                    const int g = it * NT + t, s = g % TC_STAGES, b = g % TC_ACC;
                    mbar_wait(acc_empty(b), ((g / TC_ACC) & 1) ^ 1);
                    mbar_wait(b_full(s), (g / TC_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(b * TC_N);
                    uint32_t acc = 0;
                    // split products in decreasing magnitude: x0w0, x0w1, x1w0, x0w2, x1w1, x2w0 (the last three only with SX = 3)
#pragma unroll
                    for (int xa = 0; xa < SX; ++xa) {
#pragma unroll
                        for (int wb = 0; wb < 3 - xa; ++wb) {
                            const uint32_t a_addr = smem_u32(sA + xa * TC_TILE_BYTES);
                            const uint32_t b_addr = smem_u32(sB + (s * 3 + wb) * TC_TILE_BYTES);
#pragma unroll
                            for (int k4 = 0; k4 < TC_KH / 16; ++k4) {
                                umma_bf16(d, umma_desc_k128(a_addr + k4 * 32), umma_desc_k128(b_addr + k4 * 32), acc);
                                acc = 1;
                            }
                        }
                    }
                    umma_commit(b_empty(s));     // the stage may be refilled once these MMAs have read it
                    umma_commit(acc_full(b));    // ... and the accumulator is complete
                }
            }
        } else {
            // ===== operand gather + epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 = users of those rows; the two warps of a
            // quadrant split the columns of every accumulator (half 0: columns 0-63, half 1: 64-127)
            const int quad = warp & 3;
            const int half = (warp - 2) >> 2;
            const int row = quad * 32 + lane;
            // ---- X_h of the 128 users -> swizzled A tile(s).  Each warp fills the rows of 16 users: rows zeroed first, row
            // bounds fetched by 16 lanes at once, then one user at a time with four independent 32-entry batches in flight
            // (index -> heavy slot is a dependent load: without the batching this phase was 14 % of the kernel's samples)
            const int rbase = quad * 32 + half * 16;
#pragma unroll
            for (int xa = 0; xa < SX; ++xa)
                for (int z = lane; z < 16 * 8; z += 32)
                    *reinterpret_cast<uint4 *>(sA + xa * TC_TILE_BYTES + (rbase + (z >> 3)) * 128 + (z & 7) * 16) = make_uint4(0u, 0u, 0u, 0u);
            int my_a = 0, my_b = 0;
            if (lane < 16 && blk * TC_M + rbase + lane < P.n_query) {
                const int u = P.users[blk * TC_M + rbase + lane];
                my_a = P.rptr[u]; my_b = P.rptr[u + 1];
            }
            __syncwarp();
            auto put = [&](int r, int h, float x) {
                const int off = tc_sw_off(r, h);
                const __nv_bfloat16 x0 = __float2bfloat16_rn(x);
                *reinterpret_cast<__nv_bfloat16 *>(sA + off) = x0;
                if (SX == 3) {
                    const float r1 = x - __bfloat162float(x0);
                    const __nv_bfloat16 x1 = __float2bfloat16_rn(r1);
                    const __nv_bfloat16 x2 = __float2bfloat16_rn(r1 - __bfloat162float(x1));
                    *reinterpret_cast<__nv_bfloat16 *>(sA + TC_TILE_BYTES + off) = x1;
                    *reinterpret_cast<__nv_bfloat16 *>(sA + 2 * TC_TILE_BYTES + off) = x2;
                }
            };
            for (int uu = 0; uu < 16; ++uu) {
                const int r = rbase + uu;
                const int ra = __shfl_sync(0xffffffffu, my_a, uu), rb = __shfl_sync(0xffffffffu, my_b, uu);
                for (int p = ra + lane; p < rb; p += 128) {
                    int it_[4], h_[4];
                    float x_[4];
#pragma unroll
                    for (int z = 0; z < 4; ++z) {
                        const bool in = p + 32 * z < rb;
                        it_[z] = in ? P.ridx[p + 32 * z] : -1;
                        x_[z] = in ? P.rval[p + 32 * z] : 0.0f;
                    }
#pragma unroll
                    for (int z = 0; z < 4; ++z) h_[z] = it_[z] >= 0 ? P.heavy_of[it_[z]] : -1;
#pragma unroll
                    for (int z = 0; z < 4; ++z)
                        if (h_[z] >= 0) put(r, h_[z], x_[z]);
                }
            }
            __syncwarp();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full);

            // ---- epilogue.  One warp per scheduler has little latency hiding, so the common case is kept to a handful of
            // instructions per 32 columns: the maximum of the chunk (FMNMX3 tree) against the user's threshold; only chunks in
            // which some lane has a candidate build the per-lane pass mask, and the candidates themselves are re-read from TMEM
            // one column at a time (warp-uniform loop over the union of the pass masks), because registers cannot be indexed
            // by a run-time column.
            const int q = blk * TC_M + row;
            const bool live = q < P.n_query;
            int r_cur = 0, r_end = 0;
            if (live && P.filter) { const int u = P.users[q]; r_cur = P.rptr[u]; r_end = P.rptr[u + 1]; }
            // the next four interacted items of this user: the load that refills the window is issued three items before its
            // value is needed, so the cursor walk below does not wait for global memory in (nearly) every tile
            auto rd = [&](int p) { return p < r_end ? P.ridx[p] : 0x7fffffff; };
            int n0 = rd(r_cur), n1 = rd(r_cur + 1), n2 = rd(r_cur + 2), n3 = rd(r_cur + 3);
            float *ls = list_s + half * TC_KLIST * TC_M + row;
            int *li = list_i + half * TC_KLIST * TC_M + row;
            int n = 0;
            float thr = 0.0f;    // scores are >= 0 (W >= 0, X >= 0): only positive ones are candidates
            const int k = P.k;
            for (int t = 0; t < NT; ++t) {
                const int g = it * NT + t, b = g % TC_ACC;
                const int t0 = t * TC_N;
                // interacted items of this tile (own half) as two 32-bit masks; the row is ascending: a cursor walks it once
                uint32_t ma = 0, mb = 0;
                while (n0 < t0 + TC_N) {
                    const int c = n0 - t0 - half * 64;
                    const uint32_t bit = 1u << (c & 31);
                    ma |= (c >> 5) == 0 ? bit : 0u; mb |= (c >> 5) == 1 ? bit : 0u;
                    n0 = n1; n1 = n2; n2 = n3;
                    n3 = rd(r_cur + 4);
                    ++r_cur;
                }
                mbar_wait(acc_full(b), (g / TC_ACC) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b * TC_N + half * 64);
                // both 32-column chunks of this warp's half are requested before the wait: one TMEM round trip per tile
                uint32_t va[32], vb[32];
                tmem_ld32_nowait(taddr, va);
                tmem_ld32_nowait(taddr + 32u, vb);
                tmem_wait_ld();
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t (&v)[32] = cc == 0 ? va : vb;
                    const int col0 = t0 + half * 64 + cc * 32;
                    if (P.dbg && live) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) P.dbg[(size_t)q * P.i_pad + col0 + c] = __uint_as_float(v[c]);
                    }
                    float mx = __uint_as_float(v[0]);
#pragma unroll
                    for (int c = 1; c < 32; ++c) mx = fmaxf(mx, __uint_as_float(v[c]));
                    if (__any_sync(0xffffffffu, live && mx > thr)) {
                        uint32_t pm = 0;
#pragma unroll
                        for (int c = 0; c < 32; ++c) pm |= (__uint_as_float(v[c]) > thr) ? (1u << c) : 0u;
                        pm &= ~(cc == 0 ? ma : mb);
                        if (!live) pm = 0;
                        uint32_t un = __reduce_or_sync(0xffffffffu, pm);
                        while (un) {
                            const int c = __ffs(un) - 1;
                            un &= un - 1;
                            const float sv = __uint_as_float(tmem_ld1(taddr + (uint32_t)(cc * 32 + c)));
                            if (((pm >> c) & 1u) && sv > thr) {
                                const TcListState st = tc_insert(sv, col0 + c, ls, li, n, thr, k);
                                n = st.n; thr = st.thr;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty(b));
            }
            if (live) {
                for (int s = 0; s < TC_KLIST; ++s) {
                    P.out_ids[(size_t)q * TC_OUT + half * TC_KLIST + s] = s < n ? li[s * TC_M] : -1;
                    P.out_scores[(size_t)q * TC_OUT + half * TC_KLIST + s] = s < n ? ls[s * TC_M] : 0.0f;
                }
                if (half == 0) P.out_cnt[q] = 0;   // (the lists carry -1 padding; the count is kept for the interface)
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------------------------
// pack: W_h as three K-major bf16 planes [3][i_pad][64] (plane p, item j, heavy slot h) and as dense fp32 rows [n_heavy][n_items]
// ------------------------------------------------------------------------------------------------------------------
__global__ void tc_pack_kernel(const int *__restrict__ wrptr, const int *__restrict__ wridx, const float *__restrict__ wrval,
                               const int *__restrict__ heavy_list, int n_heavy, int n_items, int i_pad,
                               __nv_bfloat16 *__restrict__ bt, float *__restrict__ wd, float *__restrict__ colmax,
                               int *__restrict__ neg_flag) {
    const int h = blockIdx.y;
    if (h >= n_heavy) return;
    const int i = heavy_list[h];
    const int a = wrptr[i], b = wrptr[i + 1];
    for (int e = a + blockIdx.x * blockDim.x + threadIdx.x; e < b; e += gridDim.x * blockDim.x) {
        const int j = wridx[e];
        const float w = wrval[e];
        if (w < 0.0f) *neg_flag = 1;
        wd[(size_t)h * n_items + j] = w;
        if (w > 0.0f) atomicMax(reinterpret_cast<int *>(colmax + j), __float_as_int(w));   // non-negative floats order like ints
        const __nv_bfloat16 w0 = __float2bfloat16_rn(w);
        const float r1 = w - __bfloat162float(w0);
        const __nv_bfloat16 w1 = __float2bfloat16_rn(r1);
        const __nv_bfloat16 w2 = __float2bfloat16_rn(r1 - __bfloat162float(w1));
        bt[((size_t)0 * i_pad + j) * TC_KH + h] = w0;
        bt[((size_t)1 * i_pad + j) * TC_KH + h] = w1;
        bt[((size_t)2 * i_pad + j) * TC_KH + h] = w2;
    }
}

// any stored W value below zero (a light row)?  any X value below zero or not exactly a bf16?
__global__ void tc_scan_kernel(const float *__restrict__ vals, int64_t n, int *__restrict__ flags) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = vals[i];
    if (v < 0.0f) flags[0] = 1;
    if (__bfloat162float(__float2bfloat16_rn(v)) != v) flags[1] = 1;
}

// ------------------------------------------------------------------------------------------------------------------
// light rows + merge (CUDA cores), one CTA per query
// ------------------------------------------------------------------------------------------------------------------
constexpr int FX_NT = 256;
constexpr int FX_CAND = 64;           // candidates kept for the final selection (tensor-core lists + surviving cells)
constexpr double FX_SCALE = 4294967296.0;   // 2^32 fixed point
// Two configurations of the same kernel.  The kernel is a chain of dependent loads and barriers per user (~10 us), so what
// counts is how many users are in flight: the SMALL table (<= 1,280 light entries, <= 128 light items: 91 % of the users at
// the ML-20M shape) needs 27 KB per CTA = 8 CTAs per SM instead of 3; users beyond it are put on a list and taken by a second
// launch with the BIG table (<= 2,560 entries, <= 1,024 items); users beyond that go to the exact kernel.
constexpr int FX_SLOTS_BIG = 4096, FX_CAP_BIG = 2560, FX_ROWS_BIG = 1024;
constexpr int FX_SLOTS_SMALL = 2048, FX_CAP_SMALL = 1280, FX_ROWS_SMALL = 128;

template <int FX_SLOTS, int FX_ROWS>
struct FxShared {
    unsigned long long val[FX_SLOTS];
    int key[FX_SLOTS];
    int la[FX_ROWS];     // staged light rows: first entry in W's CSR, exclusive prefix of the lengths, rating
    int lpre[FX_ROWS + 1];
    float lx[FX_ROWS];
    unsigned long long staged;   // high word: staged rows, low word: their entries (one atomic claims a slot AND its offset)
    int n_rows;
    int hh[TC_KH];
    float hx[TC_KH];
    float cs[FX_CAND];
    int ci[FX_CAND];
    int n_heavy_u, n_cand, n_light, fallback;
    float thr, x1;
    int red_i[FX_NT / 32];
    int red_s[FX_NT / 32];
};

// the table of one user uses the first `mask + 1` slots (a power of two >= twice the user's light entries, at least 256): the
// median user has 250 light entries, and clearing and scanning all 4,096 slots for everybody was most of the kernel's time
__device__ __forceinline__ int fx_hash(int j, int mask) { return (int)(((unsigned)j * 2654435761u) >> 20) & mask; }

template <int FX_SLOTS, int FX_CAP, int FX_ROWS, bool DEFER>
__global__ void __launch_bounds__(FX_NT) recommend_tcfix_kernel(
    const int *__restrict__ rptr, const int *__restrict__ ridx, const float *__restrict__ rval, const int *__restrict__ users,
    int n_query, const int *__restrict__ wrptr, const int *__restrict__ wridx, const float *__restrict__ wrval,
    const int *__restrict__ heavy_of, const float *__restrict__ wd, const float *__restrict__ colmax, int n_items, int k,
    int filter, int dense_mode, const int *__restrict__ tc_ids, const float *__restrict__ tc_scores,
    int *__restrict__ out_ids, float *__restrict__ out_scores, int *__restrict__ out_cnt, int *__restrict__ fallback,
    int *__restrict__ next_query, int *__restrict__ work_list, int *__restrict__ work_count) {
    // DEFER: every query is visited, the ones that do not fit this configuration are appended to work_list (work_count).
    // !DEFER with work_list: only the queries of work_list[0 .. *work_count) are visited.
    extern __shared__ __align__(16) unsigned char fx_raw[];
    using Shared = FxShared<FX_SLOTS, FX_ROWS>;
    Shared &S = *reinterpret_cast<Shared *>(fx_raw);
    if (!DEFER && work_list) n_query = *work_count;
    __shared__ int s_q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_q = atomicAdd(next_query, 1);
        __syncthreads();
        if (s_q >= n_query) break;
        const int q = (!DEFER && work_list) ? work_list[s_q] : s_q;
        const int u = users[q];
        const int r0 = rptr[u], r1 = rptr[u + 1];
        // warp 0 merges the tensor-core lists in pass 3: their loads are issued here, a whole pass ahead of their use
        // (everybody else waited at the barrier behind pass 3 for this round trip: 29 % of the kernel's samples)
        int pre_j = -1;
        float pre_sc = -1.0f;
        if (warp == 0) {
            pre_j = tc_ids[(size_t)q * TC_OUT + lane];
            pre_sc = tc_scores[(size_t)q * TC_OUT + lane];
        }
        if (tid == 0) { S.n_heavy_u = 0; S.n_cand = 0; S.n_light = 0; S.fallback = 0; S.n_rows = 0; S.staged = 0ull; }
        __syncthreads();
        // ---- pass 1: one coalesced sweep over the row: heavy items (ascending, ordered compaction) and the light rows
        // (first entry, length, rating) staged in shared memory -- the scatter below then has no dependent global loads
        for (int base = r0; base < r1; base += FX_NT) {
            const int p = base + tid;
            int h = -1, a = 0, len = 0;
            float x = 0.f;
            if (p < r1) {
                const int i = ridx[p];
                x = rval[p];
                h = heavy_of[i];
                if (h < 0) { a = wrptr[i]; len = wrptr[i + 1] - a; }
            }
            if (len > 0) {
                // slot and exclusive entry offset from ONE atomic, so the offsets ascend with the slots (binary-searchable)
                const unsigned long long old = atomicAdd(&S.staged, (1ull << 32) | (unsigned long long)(unsigned)len);
                const int slot = (int)(old >> 32);
                if (slot < FX_ROWS) { S.la[slot] = a; S.lpre[slot] = (int)(old & 0xffffffffull); S.lx[slot] = x; }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, h >= 0);
            if (lane == 0) S.red_i[warp] = __popc(bal);
            __syncthreads();
            int off = S.n_heavy_u, tot = 0;
            for (int w = 0; w < FX_NT / 32; ++w) { if (w < warp) off += S.red_i[w]; tot += S.red_i[w]; }
            if (h >= 0) { const int s = off + __popc(bal & ((1u << lane) - 1u)); if (s < TC_KH) { S.hh[s] = h; S.hx[s] = x; } }
            __syncthreads();
            if (tid == 0) S.n_heavy_u += tot;
            __syncthreads();
        }
        const int n_rows = (int)(S.staged >> 32);
        if (n_rows > FX_ROWS) {         // more light items than can be staged: the big configuration / the exact kernel
            if (tid == 0) {
                if (DEFER) work_list[atomicAdd(work_count, 1)] = q;
                else { fallback[q] = 1; out_cnt[q] = 0; }
            }
            continue;
        }
        if (tid == 0) { S.lpre[n_rows] = (int)(S.staged & 0xffffffffull); S.n_light = (int)(S.staged & 0xffffffffull); }
        __syncthreads();
        const int n_light = S.n_light;
        if (n_light > FX_CAP) {      // the table would overflow: the big configuration / the exact kernel
            if (tid == 0) {
                if (DEFER) work_list[atomicAdd(work_count, 1)] = q;
                else { fallback[q] = 1; out_cnt[q] = 0; }
            }
            continue;
        }
        const int nh = min(S.n_heavy_u, TC_KH);
        int tsize = 256;
        while (tsize < 2 * n_light && tsize < FX_SLOTS) tsize <<= 1;   // (n_light <= FX_CAP: load <= 0.625 at the full size)
        const int mask = tsize - 1;
        {   // clear the table with 16-byte stores
            uint4 *kv = reinterpret_cast<uint4 *>(S.key);
            uint4 *vv = reinterpret_cast<uint4 *>(S.val);
            for (int s = tid; s < tsize / 4; s += FX_NT) kv[s] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            for (int s = tid; s < tsize / 2; s += FX_NT) vv[s] = make_uint4(0u, 0u, 0u, 0u);
        }
        __syncthreads();
        // ---- pass 2: every light entry -> table, one entry per thread and step (64-bit fixed-point sums: the result does
        // not depend on the order of the adds)
        for (int e = tid; e < n_light; e += FX_NT) {
            int lo = 0, hi = n_rows;            // staged row of entry e: last r with lpre[r] <= e
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (S.lpre[mid] <= e) lo = mid; else hi = mid; }
            const int ge = S.la[lo] + (e - S.lpre[lo]);
            const int j = wridx[ge];
            const float add = __fmul_rn(S.lx[lo], wrval[ge]);
            int slot = fx_hash(j, mask);
            for (;;) {
                const int prev = atomicCAS(&S.key[slot], -1, j);
                if (prev == -1 || prev == j) break;
                slot = (slot + 1) & mask;
            }
            atomicAdd(&S.val[slot], (unsigned long long)__double2ll_rn((double)add * FX_SCALE));
        }
        // ---- pass 3 (warp 0, meanwhile): the two heavy-only lists of the tensor-core kernel -> candidates 0..31; the
        // smallest score a cell must beat = k-th best of their union (0 while the union is short); sum of the heavy ratings
        if (warp == 0) {
            const int j = pre_j;
            const float sc = j >= 0 ? pre_sc : -1.0f;
            S.cs[lane] = sc; S.ci[lane] = j;
            int rank = 0;       // entries strictly better than mine (ties by lane)
            for (int l = 0; l < 32; ++l) {
                const float o = __shfl_sync(0xffffffffu, sc, l);
                rank += (o > sc) || (o == sc && l < lane);
            }
            const unsigned has = __ballot_sync(0xffffffffu, j >= 0);
            float thr = 0.0f;
            const unsigned kth = __ballot_sync(0xffffffffu, rank == k - 1 && j >= 0);
            if (__popc(has) >= k && kth) thr = __shfl_sync(0xffffffffu, sc, __ffs(kth) - 1);
            float x1 = 0.0f;
            for (int e = lane; e < nh; e += 32) x1 += S.hx[e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x1 += __shfl_xor_sync(0xffffffffu, x1, o);
            if (lane == 0) { S.thr = thr; S.x1 = x1; S.n_cand = 32; }
        }
        __syncthreads();
        const float thr = S.thr, x1 = S.x1;
        // tensor-core candidates that are also table cells are superseded by the cell (which carries the light part)
        if (tid < 32 && S.ci[tid] >= 0) {
            const int j = S.ci[tid];
            int slot = fx_hash(j, mask);
            for (;;) {
                const int kk = S.key[slot];
                if (kk == -1) break;
                if (kk == j) { S.cs[tid] = -1.0f; break; }
                slot = (slot + 1) & mask;
            }
        }
        // ---- pass 4: table cells.  Heavy part <= x1 * colmax[j] (every term is >= 0): most cells cannot reach the threshold
        // and are dropped after one load; the others get their heavy part (fp32, ascending item order) from the dense rows
        for (int s = tid; s < tsize; s += FX_NT) {
            const int j = S.key[s];
            if (j < 0) continue;
            const float lpart = (float)((double)(long long)S.val[s] / FX_SCALE);
            const float bound = __fadd_rn(lpart, __fmul_rn(__fmul_rn(x1, colmax[j]), 1.00001f));
            if (bound < thr) continue;
            if (filter) {
                int lo = r0, hi = r1;
                while (lo < hi) { const int mid = (lo + hi) >> 1; if (ridx[mid] < j) lo = mid + 1; else hi = mid; }
                if (lo < r1 && ridx[lo] == j) continue;
            }
            float sh = 0.0f;
            int e = 0;
            for (; e + 3 < nh; e += 4) {       // four gathers in flight, added in ascending item order
                const float g0 = wd[(size_t)S.hh[e] * n_items + j], g1 = wd[(size_t)S.hh[e + 1] * n_items + j];
                const float g2 = wd[(size_t)S.hh[e + 2] * n_items + j], g3 = wd[(size_t)S.hh[e + 3] * n_items + j];
                sh = __fadd_rn(sh, __fmul_rn(S.hx[e], g0)); sh = __fadd_rn(sh, __fmul_rn(S.hx[e + 1], g1));
                sh = __fadd_rn(sh, __fmul_rn(S.hx[e + 2], g2)); sh = __fadd_rn(sh, __fmul_rn(S.hx[e + 3], g3));
            }
            for (; e < nh; ++e) sh = __fadd_rn(sh, __fmul_rn(S.hx[e], wd[(size_t)S.hh[e] * n_items + j]));
            const float sc = __fadd_rn(sh, lpart);
            if (sc > thr || (sc == thr && sc > 0.0f)) {
                const int c = atomicAdd(&S.n_cand, 1);
                if (c < FX_CAND) { S.cs[c] = sc; S.ci[c] = j; }
                else S.fallback = 1;
            }
        }
        __syncthreads();
        if (S.fallback) {
            if (tid == 0) { fallback[q] = 1; out_cnt[q] = 0; }
            continue;
        }
        // ---- top-k of the (<= 64) candidates by one warp: rank by (score desc, item id desc), positive scores only
        if (warp == 0) {
            const int nc = min(S.n_cand, FX_CAND);
            float v0 = lane < nc ? S.cs[lane] : -1.0f, v1 = lane + 32 < nc ? S.cs[lane + 32] : -1.0f;
            int j0 = lane < nc ? S.ci[lane] : -1, j1 = lane + 32 < nc ? S.ci[lane + 32] : -1;
            int rk0 = 0, rk1 = 0;
            for (int l = 0; l < 32; ++l) {
                const float a0 = __shfl_sync(0xffffffffu, v0, l), a1 = __shfl_sync(0xffffffffu, v1, l);
                const int b0 = __shfl_sync(0xffffffffu, j0, l), b1 = __shfl_sync(0xffffffffu, j1, l);
                rk0 += (a0 > v0 || (a0 == v0 && b0 > j0)) + (a1 > v0 || (a1 == v0 && b1 > j0));
                rk1 += (a0 > v1 || (a0 == v1 && b0 > j1)) + (a1 > v1 || (a1 == v1 && b1 > j1));
            }
            const bool ok0 = v0 > 0.0f && rk0 < k, ok1 = v1 > 0.0f && rk1 < k;
            if (ok0) { out_ids[(size_t)q * k + rk0] = j0; out_scores[(size_t)q * k + rk0] = v0; }
            if (ok1) { out_ids[(size_t)q * k + rk1] = j1; out_scores[(size_t)q * k + rk1] = v1; }
            const int cnt = __popc(__ballot_sync(0xffffffffu, ok0)) + __popc(__ballot_sync(0xffffffffu, ok1));
            for (int e = cnt + lane; e < k; e += 32) { out_ids[(size_t)q * k + e] = -1; out_scores[(size_t)q * k + e] = 0.0f; }
            if (lane == 0) {
                out_cnt[q] = cnt;
                // dense semantics list zero-score items when fewer than k are positive: the exact kernel knows that order
                if (dense_mode && cnt < k) fallback[q] = 1;
            }
        }
    }
}

}  // namespace rt

using namespace rt;

extern "C" int rt_tc_pack_size(int32_t n_items, int32_t n_heavy, int32_t *h_i_pad, int64_t *h_bt_bytes, int64_t *h_wd_bytes) {
    RT_ARG(n_items > 0 && n_heavy >= 0 && h_i_pad && h_bt_bytes && h_wd_bytes, "arguments");
    const int i_pad = (n_items + TC_N - 1) / TC_N * TC_N;
    *h_i_pad = i_pad;
    *h_bt_bytes = (int64_t)3 * i_pad * TC_KH * 2;
    *h_wd_bytes = (int64_t)((n_heavy > 0 ? n_heavy : 1) + 1) * n_items * 4;   // dense rows + one row of column maxima
    return RT_OK;
}

extern "C" int rt_tc_pack_build(const int32_t *d_wrptr, const int32_t *d_wridx, const float *d_wrval, int64_t w_nnz,
                                int32_t n_items, const int32_t *d_heavy_list, int32_t n_heavy, void *d_bt, float *d_wd,
                                int32_t *h_w_nonneg, void *stream) {
    RT_ARG(n_items > 0 && n_heavy >= 0 && n_heavy <= TC_KH && d_wrptr && d_bt && d_wd && h_w_nonneg, "arguments (at most 64 heavy rows)");
    RT_ARG((((uintptr_t)d_bt) & 127) == 0, "d_bt must be 128-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t i_pad; int64_t bt_bytes, wd_bytes;
    rt_tc_pack_size(n_items, n_heavy, &i_pad, &bt_bytes, &wd_bytes);
    RT_CUDA(cudaMemsetAsync(d_bt, 0, (size_t)bt_bytes, st));
    RT_CUDA(cudaMemsetAsync(d_wd, 0, (size_t)wd_bytes, st));
    int *flags = (int *)rt::scratch(SCR_MISC, 256);
    if (!flags) return RT_ERR_CUDA;
    flags += 32;
    RT_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int), st));
    if (n_heavy > 0) {
        tc_pack_kernel<<<dim3(8, n_heavy), 256, 0, st>>>(d_wrptr, d_wridx, d_wrval, d_heavy_list, n_heavy, n_items, i_pad,
                                                       (__nv_bfloat16 *)d_bt, d_wd, d_wd + (size_t)n_heavy * n_items, flags);
        RT_CHECK_LAUNCH();
    }
    if (w_nnz > 0) {
        tc_scan_kernel<<<(unsigned)((w_nnz + 255) / 256), 256, 0, st>>>(d_wrval, w_nnz, flags);
        RT_CHECK_LAUNCH();
    }
    int h[2] = {0, 0};
    RT_CUDA(cudaMemcpyAsync(h, flags, sizeof(h), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    *h_w_nonneg = h[0] ? 0 : 1;
    return RT_OK;
}

extern "C" int rt_values_bf16_exact(const float *d_vals, int64_t n, int32_t *h_nonneg, int32_t *h_bf16_exact, void *stream) {
    RT_ARG(n >= 0 && h_nonneg && h_bf16_exact && (n == 0 || d_vals), "arguments");
    cudaStream_t st = (cudaStream_t)stream;
    int *flags = (int *)rt::scratch(SCR_MISC, 256);
    if (!flags) return RT_ERR_CUDA;
    flags += 40;
    RT_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int), st));
    if (n > 0) {
        tc_scan_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_vals, n, flags);
        RT_CHECK_LAUNCH();
    }
    int h[2] = {0, 0};
    RT_CUDA(cudaMemcpyAsync(h, flags, sizeof(h), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    *h_nonneg = h[0] ? 0 : 1;
    *h_bf16_exact = h[1] ? 0 : 1;
    return RT_OK;
}

extern "C" int rt_slim_recommend_tc(const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval, const int32_t *d_users,
                                    int32_t n_query, const int32_t *d_wrptr, const int32_t *d_wridx, const float *d_wrval,
                                    const int32_t *d_heavy_of, int32_t n_heavy, const void *d_bt, const float *d_wd, int32_t n_items, int32_t k,
                                    int32_t filter_interacted, int32_t mode, int32_t x_planes, int32_t *d_tc_ids,
                                    float *d_tc_scores, int32_t *d_tc_cnt, int32_t *d_out_ids, float *d_out_scores,
                                    int32_t *d_out_cnt, int32_t *d_fallback, float *d_dbg_scores, void *stream) {
    RT_ARG(k >= 1 && k <= TC_KLIST, "k must be in [1,16] for the tensor-core path");
    RT_ARG(n_items > 0 && (mode == RT_TOPK_DENSE || mode == RT_TOPK_SPARSE) && (x_planes == 1 || x_planes == 3) && n_heavy >= 1 && n_heavy <= TC_KH, "arguments");
    if (n_query <= 0) return RT_OK;
    RT_ARG(d_rptr && d_users && d_wrptr && d_heavy_of && d_bt && d_wd && d_tc_ids && d_tc_scores && d_tc_cnt && d_out_ids &&
               d_out_scores && d_out_cnt && d_fallback, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int cc_major = 0;
    rt_device_info(nullptr, nullptr, &cc_major, nullptr);
    if (cc_major != 10) { rt::set_error("rt_slim_recommend_tc needs an sm_100 device (tcgen05)"); return RT_ERR_NO_DEVICE; }
    tc_encode_fn enc = tc_encoder();
    if (!enc) { rt::set_error("cuTensorMapEncodeTiled is not available from the driver"); return RT_ERR_CUDA; }
    const int i_pad = (n_items + TC_N - 1) / TC_N * TC_N;
    CUtensorMap tmap;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)TC_KH, (cuuint64_t)3 * (cuuint64_t)i_pad};
        const cuuint64_t gstride[1] = {(cuuint64_t)TC_KH * 2};
        const cuuint32_t box[2] = {(cuuint32_t)TC_KH, (cuuint32_t)TC_N};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d_bt), gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { rt::set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return RT_ERR_CUDA; }
    }
    TcParams P;
    P.rptr = d_rptr; P.ridx = d_ridx; P.rval = d_rval; P.users = d_users; P.n_query = n_query; P.heavy_of = d_heavy_of;
    P.n_tiles = i_pad / TC_N; P.i_pad = i_pad; P.k = k; P.filter = filter_interacted;
    P.out_ids = d_tc_ids; P.out_scores = d_tc_scores; P.out_cnt = d_tc_cnt; P.dbg = d_dbg_scores;
    const int n_blocks = (n_query + TC_M - 1) / TC_M;
    int grid = rt::sm_count();
    if (grid > n_blocks) grid = n_blocks;
    const size_t smem = 1024 + (size_t)(x_planes + TC_STAGES * 3) * TC_TILE_BYTES + (size_t)2 * TC_KLIST * TC_M * 8 + 16 * 8 + 64;
    if (x_planes == 1) {
        RT_CUDA(cudaFuncSetAttribute(recommend_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        recommend_tc_kernel<1><<<grid, TC_THREADS, smem, st>>>(tmap, P);
    } else {
        RT_CUDA(cudaFuncSetAttribute(recommend_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        recommend_tc_kernel<3><<<grid, TC_THREADS, smem, st>>>(tmap, P);
    }
    RT_CHECK_LAUNCH();
    int *d_next = (int *)rt::scratch(SCR_MISC, 256);
    if (!d_next) return RT_ERR_CUDA;
    d_next += 48;                       // [0] queue of the small configuration, [1] of the big one, [2] length of the deferred list
    int *d_list = (int *)rt::scratch(SCR_SCORE, ((size_t)n_query + 64) * sizeof(int));
    if (!d_list) return RT_ERR_CUDA;
    RT_CUDA(cudaMemsetAsync(d_next, 0, 4 * sizeof(int), st));
    RT_CUDA(cudaMemsetAsync(d_fallback, 0, sizeof(int) * (size_t)n_query, st));
    const float *colmax = d_wd + (size_t)n_heavy * n_items;
    const int dense = mode == RT_TOPK_DENSE ? 1 : 0;
    {
        auto kern = recommend_tcfix_kernel<FX_SLOTS_SMALL, FX_CAP_SMALL, FX_ROWS_SMALL, true>;
        const size_t fsmem = sizeof(FxShared<FX_SLOTS_SMALL, FX_ROWS_SMALL>);
        RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        int fgrid = rt::sm_count() * 8;
        if (fgrid > n_query) fgrid = n_query;
        kern<<<fgrid, FX_NT, fsmem, st>>>(d_rptr, d_ridx, d_rval, d_users, n_query, d_wrptr, d_wridx, d_wrval, d_heavy_of, d_wd, colmax,
                                          n_items, k, filter_interacted, dense, d_tc_ids, d_tc_scores, d_out_ids, d_out_scores, d_out_cnt,
                                          d_fallback, d_next, d_list, d_next + 2);
        RT_CHECK_LAUNCH();
    }
    {
        auto kern = recommend_tcfix_kernel<FX_SLOTS_BIG, FX_CAP_BIG, FX_ROWS_BIG, false>;
        const size_t fsmem = sizeof(FxShared<FX_SLOTS_BIG, FX_ROWS_BIG>);
        RT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
        int fgrid = rt::sm_count() * 3;
        if (fgrid > n_query) fgrid = n_query;
        kern<<<fgrid, FX_NT, fsmem, st>>>(d_rptr, d_ridx, d_rval, d_users, n_query, d_wrptr, d_wridx, d_wrval, d_heavy_of, d_wd, colmax,
                                          n_items, k, filter_interacted, dense, d_tc_ids, d_tc_scores, d_out_ids, d_out_scores, d_out_cnt,
                                          d_fallback, d_next + 1, d_list, d_next + 2);
    }
    RT_CHECK_LAUNCH();
    return RT_OK;
}
