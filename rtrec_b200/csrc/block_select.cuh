// block_select.cuh -- CTA-wide exact top-n selection by radix select + rank sort.
//
// Order: (value descending, index descending) -- what `np.argsort(s)[-n:][::-1]` produces when
// the sort is stable (reference: slim_elastic.py:143, :769).  All threads of the CTA must call.
#pragma once
#include "common.cuh"

namespace rt {

struct SelectScratch {
    int hist[256];
    int warp_tot[32];
    int bcast[4];
    int count;
};

// Finds, over idx in [0, N) with eligible(idx), the n largest of (key(idx), idx) and writes them
// sorted (key desc, idx desc) to out_idx/out_key (capacity >= n, shared or global memory that
// every thread of the CTA can address).  cand_key/cand_idx are CTA-visible temporaries of
// capacity >= n.  Returns the number written (min(n, #eligible)), uniform across the CTA.
//
// KeyFn: uint32_t operator()(int idx) -- order-preserving key; called several times per idx.
// EligFn: bool operator()(int idx, uint32_t key).
template <typename KeyFn, typename EligFn>
__device__ int block_top_n(int N, int n, KeyFn key_of, EligFn eligible, SelectScratch *ss,
                           uint32_t *cand_key, int *cand_idx, int *out_idx, uint32_t *out_key) {
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31;
    if (n <= 0 || N <= 0) return 0;
    // ---- radix select on the 32-bit key, MSB first -------------------------------------------
    uint32_t prefix = 0, pmask = 0;
    int need = n;         // how many still to take among keys matching `prefix` under `pmask`
    int total_elig = 0;
    int eq_count = 0;     // elements equal to the final threshold key
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int b = tid; b < 256; b += NT) ss->hist[b] = 0;
        __syncthreads();
        for (int base = 0; base < N; base += NT) {
            const int i = base + tid;
            bool ok = false;
            uint32_t k = 0;
            if (i < N) {
                k = key_of(i);
                ok = eligible(i, k) && ((k & pmask) == prefix);
            }
            const int bin = (int)((k >> shift) & 0xffu);
            // warp-aggregate identical bins (rows are often dominated by one value, e.g. 0)
            const unsigned act = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const unsigned same = __match_any_sync(act, bin);
                if (lane == __ffs(same) - 1) atomicAdd(&ss->hist[bin], __popc(same));
            }
        }
        __syncthreads();
        if (tid < 32) {
            // warp 0: scan bins from the top; lane l owns bins [8l, 8l+8)
            int loc[8], s = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { loc[q] = ss->hist[lane * 8 + q]; s += loc[q]; }
            // suffix sums across lanes: above = sum over lanes > lane
            int incl = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_down_sync(0xffffffffu, incl, o);
                if (lane + o < 32) incl += v;
            }
            const int above = incl - s;  // elements in bins of higher lanes
            const int tot = __shfl_sync(0xffffffffu, incl, 0);
            if (pass == 0 && lane == 0) ss->bcast[3] = tot;
            // the threshold bin lives in the lane where above < need <= above + s
            const bool mine = (above < need) && (need <= above + s);
            if (mine) {
                int acc = above;
                for (int q = 7; q >= 0; --q) {
                    if (acc + loc[q] >= need) {
                        ss->bcast[0] = lane * 8 + q;   // bin
                        ss->bcast[1] = need - acc;     // still needed inside that bin
                        ss->bcast[2] = loc[q];         // population of that bin
                        break;
                    }
                    acc += loc[q];
                }
            }
            if (lane == 0 && tot < need) {  // fewer eligible than n: take everything
                ss->bcast[0] = -1;
            }
        }
        __syncthreads();
        if (pass == 0) total_elig = ss->bcast[3];
        if (ss->bcast[0] < 0) { need = -1; __syncthreads(); break; }
        prefix |= ((uint32_t)ss->bcast[0]) << shift;
        pmask |= 0xffu << shift;
        need = ss->bcast[1];
        eq_count = ss->bcast[2];
        __syncthreads();
    }
    // ---- tie at the threshold: choose the `need` largest indices among key == prefix --------
    int idx_thresh = 0;
    const bool take_all = (need < 0);
    if (!take_all && need < eq_count) {
        uint32_t ipref = 0, imask = 0;
        int ineed = need;
        int nbits = 32 - __clz(N > 1 ? N - 1 : 1);
        int npass = (nbits + 7) / 8;
        for (int pass = 0; pass < npass; ++pass) {
            const int shift = 8 * (npass - 1 - pass);
            for (int b = tid; b < 256; b += NT) ss->hist[b] = 0;
            __syncthreads();
            for (int base = 0; base < N; base += NT) {
                const int i = base + tid;
                if (i < N) {
                    const uint32_t k = key_of(i);
                    if (eligible(i, k) && k == prefix && (((uint32_t)i) & imask) == ipref)
                        atomicAdd(&ss->hist[(i >> shift) & 0xff], 1);
                }
            }
            __syncthreads();
            if (tid == 0) {
                int acc = 0;
                for (int b = 255; b >= 0; --b) {
                    if (acc + ss->hist[b] >= ineed) { ss->bcast[0] = b; ss->bcast[1] = ineed - acc; break; }
                    acc += ss->hist[b];
                }
            }
            __syncthreads();
            ipref |= ((uint32_t)ss->bcast[0]) << shift;
            imask |= 0xffu << shift;
            ineed = ss->bcast[1];
            __syncthreads();
        }
        idx_thresh = (int)ipref;  // unique indices => exactly `need` equal-key elements have idx >= ipref
    }
    // ---- collect (unordered) ---------------------------------------------------------------
    if (tid == 0) ss->count = 0;
    __syncthreads();
    for (int base = 0; base < N; base += NT) {
        const int i = base + tid;
        if (i < N) {
            const uint32_t k = key_of(i);
            if (eligible(i, k)) {
                const bool take = take_all || k > prefix || (k == prefix && i >= idx_thresh);
                if (take) {
                    const int p = atomicAdd(&ss->count, 1);
                    if (p < n) { cand_key[p] = k; cand_idx[p] = i; }
                }
            }
        }
    }
    __syncthreads();
    int cnt = ss->count;
    if (cnt > n) cnt = n;  // cannot happen; defensive
    (void)total_elig;
    // ---- rank sort -------------------------------------------------------------------------
    for (int e = tid; e < cnt; e += NT) {
        const uint32_t ke = cand_key[e];
        const int ie = cand_idx[e];
        int rank = 0;
        for (int f = 0; f < cnt; ++f) {
            const uint32_t kf = cand_key[f];
            const int jf = cand_idx[f];
            rank += (kf > ke) || (kf == ke && jf > ie);
        }
        out_idx[rank] = ie;
        if (out_key) out_key[rank] = ke;
    }
    __syncthreads();
    return cnt;
}


// ------------------------------------------------------------------------------------------------
// Fast path: threshold from bucket maxima.
//
// key_of(idx) returns an order-preserving key, 0 meaning "not eligible".  128 strided buckets
// (bucket = idx mod 128) each remember their largest key; the n-th largest bucket maximum T is a
// lower bound of the n-th largest key (n distinct elements are >= T).  Everything > T is
// collected (normally little more than n elements); if fewer than n are strictly greater, the
// remaining picks are the elements == T with the largest indices, found by scanning from the top
// index down and stopping early.  The list is rank-sorted by (key desc, idx desc), so the result is
// exactly what block_top_n returns; block_top_n itself is the fallback when the list overflows,
// n > 128 or the block is not a multiple of 128 threads.  Two light passes over the data.
constexpr int FS_BUCKETS = 128;
constexpr int FS_LIST = 512;

struct FastSelScratch {
    union {
        struct {
            uint32_t bmax[FS_BUCKETS];
            uint32_t lkey[FS_LIST];
            int lidx[FS_LIST];
        } f;
        struct {
            SelectScratch sel;
            uint32_t cand_key[FS_BUCKETS];
            int cand_idx[FS_BUCKETS];
        } x;  // exact fallback (n <= 128); larger n uses caller-provided global temporaries
    } u;
    int wtot[2][32];
    int cnt_gt, cnt_ge;
    uint32_t T;
};

// out_idx/out_key: capacity >= n, CTA-visible.  big_key/big_idx: temporaries of capacity >= n for the
// exact path when n > 128 (may be null if n <= 128).  Returns min(n, #eligible).
template <typename KeyFn>
__device__ int block_top_n_fast(int N, int n, KeyFn key_of, FastSelScratch *fs, int *out_idx, uint32_t *out_key,
                                uint32_t *big_key = nullptr, int *big_idx = nullptr) {
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = (NT + 31) >> 5;
    if (n <= 0 || N <= 0) return 0;
    auto exact = [&]() -> int {
        auto elig = [&](int, uint32_t key) -> bool { return key != 0u; };
        __syncthreads();
        const int c = block_top_n(N, n, key_of, elig, &fs->u.x.sel, n <= FS_BUCKETS ? fs->u.x.cand_key : big_key,
                                  n <= FS_BUCKETS ? fs->u.x.cand_idx : big_idx, out_idx, out_key);
        __syncthreads();
        return c;
    };
    if (n > FS_BUCKETS || (NT % FS_BUCKETS) != 0 || N < 4 * FS_BUCKETS) return exact();
    // ---- pass 1: bucket maxima and eligible count -------------------------------------------------
    __syncthreads();
    if (tid < FS_BUCKETS) fs->u.f.bmax[tid] = 0u;
    if (tid == 0) { fs->cnt_gt = 0; fs->cnt_ge = 0; }
    __syncthreads();
    uint32_t m = 0u;
    int ne = 0;
    for (int x = tid; x < N; x += NT) {
        const uint32_t key = key_of(x);
        m = max(m, key);
        ne += key != 0u;
    }
    if (m) atomicMax(&fs->u.f.bmax[tid & (FS_BUCKETS - 1)], m);
    ne = warp_sum_i(ne);
    if (lane == 0) fs->wtot[0][warp] = ne;
    __syncthreads();
    int n_elig = 0;
    for (int w = 0; w < nwarps; ++w) n_elig += fs->wtot[0][w];
    const int kk = min(n, n_elig);
    if (kk == 0) return 0;
    if (tid < FS_BUCKETS) {
        const uint32_t bm = fs->u.f.bmax[tid];
        int rank = 0;
        for (int f = 0; f < FS_BUCKETS; ++f) {
            const uint32_t kf = fs->u.f.bmax[f];
            rank += (kf > bm) || (kf == bm && f < tid);
        }
        if (rank == kk - 1) fs->T = bm == 0u ? 1u : bm;
    }
    __syncthreads();
    const uint32_t T = fs->T;
    // ---- pass 2: collect everything >= T (also count how many are strictly above) ------------------
    auto collect = [&](bool strict) {
        for (int base = 0; base < N; base += NT) {
            const int x = base + tid;
            uint32_t key = 0u;
            if (x < N) key = key_of(x);
            const bool gt = key > T;
            const bool take = strict ? gt : key >= T;
            const unsigned bal = __ballot_sync(0xffffffffu, take);
            if (bal) {
                const unsigned bgt = __ballot_sync(0xffffffffu, gt);
                int basepos = 0;
                if (lane == 0) {
                    basepos = atomicAdd(&fs->cnt_ge, __popc(bal));
                    if (bgt) atomicAdd(&fs->cnt_gt, __popc(bgt));
                }
                basepos = __shfl_sync(0xffffffffu, basepos, 0);
                const int pos = take ? basepos + __popc(bal & ((1u << lane) - 1u)) : FS_LIST;
                if (pos < FS_LIST) { fs->u.f.lkey[pos] = key; fs->u.f.lidx[pos] = x; }
            }
        }
        __syncthreads();
    };
    collect(false);
    int total = fs->cnt_ge;
    if (total > FS_LIST) {
        // too many ties at T (or T far below the n-th key): keep only the strictly larger ones ...
        const int c_gt = fs->cnt_gt;
        if (c_gt > FS_LIST) return exact();
        __syncthreads();
        if (tid == 0) { fs->cnt_ge = 0; fs->cnt_gt = 0; }
        __syncthreads();
        collect(true);
        total = c_gt;
        if (c_gt < kk) {
            // ... and take the (kk - c_gt) largest indices among the ties, scanning downwards
            const int need = kk - c_gt;
            int taken = 0, buf = 0;
            for (int top = ((N + NT - 1) / NT) * NT; top > 0 && taken < need; top -= NT, buf ^= 1) {
                const int x = top - 1 - tid;  // thread order = descending index
                const bool eq = x < N && key_of(x) == T;
                const unsigned bal = __ballot_sync(0xffffffffu, eq);
                if (lane == 0) fs->wtot[buf][warp] = __popc(bal);
                __syncthreads();
                int off = taken, tot = 0;
                for (int w = 0; w < nwarps; ++w) { const int c = fs->wtot[buf][w]; if (w < warp) off += c; tot += c; }
                if (eq) {
                    const int pos = off + __popc(bal & ((1u << lane) - 1u));
                    if (pos < need) { fs->u.f.lkey[c_gt + pos] = T; fs->u.f.lidx[c_gt + pos] = x; }
                }
                taken += tot;
            }
            total = kk;
            __syncthreads();
        }
    }
    // ---- rank sort ------------------------------------------------------------------------------
    for (int e = tid; e < total; e += NT) {
        const uint32_t ke = fs->u.f.lkey[e];
        const int ie = fs->u.f.lidx[e];
        int rank = 0;
        for (int f = 0; f < total; ++f) {
            const uint32_t kf = fs->u.f.lkey[f];
            const int jf = fs->u.f.lidx[f];
            rank += (kf > ke) || (kf == ke && jf > ie);
        }
        if (rank < kk) { out_idx[rank] = ie; if (out_key) out_key[rank] = ke; }
    }
    __syncthreads();
    return kk;
}

// ------------------------------------------------------------------------------------------------
// Same selection over a float array in shared memory (the score tile), with the two scans done on raw
// floats: an element is eligible iff v > -inf and, when skip_zero, v != 0.  Order-preserving keys are
// formed only for bucket maxima and collected elements, which takes the key transform out of the
// per-element loops (it was a third of the scoring kernel's instructions, profiles/r1l_*).  Results are
// identical to block_top_n_fast with key_of(x) = eligible ? float_key(v[x]) : 0; every unusual case
// (n > 128, list overflow through ties) is handed to that function.
__device__ __forceinline__ float fs_key_to_float(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

// (nth+1)-th largest of the 128 bucket maxima, multiplicity counted, by warp 0: every lane holds four maxima and
// each round removes one instance of the current maximum (REDUX.MAX + ballot).  `nth` rounds of ~12 instructions
// instead of the 128 x 128 compare matrix (5 % of the scoring kernel's instructions, profiles/r1z_*).  The result
// (0 when fewer than nth+1 buckets are non-empty) is left in fs->T; the caller synchronises.
__device__ __forceinline__ void fs_nth_bucket_max(FastSelScratch *fs, int nth) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    uint32_t k0 = fs->u.f.bmax[lane], k1 = fs->u.f.bmax[lane + 32], k2 = fs->u.f.bmax[lane + 64], k3 = fs->u.f.bmax[lane + 96];
    uint32_t M = 0u;
    for (int r = 0; r <= nth; ++r) {
        const uint32_t lm = max(max(k0, k1), max(k2, k3));
        M = __reduce_max_sync(0xffffffffu, lm);
        if (M == 0u) break;
        const unsigned has = __ballot_sync(0xffffffffu, lm == M);
        if (lane == __ffs(has) - 1) {
            if (k0 == M) k0 = 0u; else if (k1 == M) k1 = 0u; else if (k2 == M) k2 = 0u; else k3 = 0u;
        }
    }
    if (lane == 0) fs->T = M;
}

__device__ inline int block_top_n_fast_f32(const float *vals, int N, int n, bool skip_zero, FastSelScratch *fs, int *out_idx,
                                    uint32_t *out_key) {
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = (NT + 31) >> 5;
    if (n <= 0 || N <= 0) return 0;
    auto key_of = [&](int x) -> uint32_t {
        const float v = vals[x];
        if (!(v > -INFINITY) || (skip_zero && v == 0.0f)) return 0u;
        return float_key(v);
    };
    if (n > FS_BUCKETS || (NT % FS_BUCKETS) != 0 || N < 4 * FS_BUCKETS)
        return block_top_n_fast(N, n, key_of, fs, out_idx, out_key);
    __syncthreads();
    if (tid < FS_BUCKETS) fs->u.f.bmax[tid] = 0u;
    if (tid == 0) { fs->cnt_gt = 0; fs->cnt_ge = 0; }
    __syncthreads();
    // 16-byte shared-memory loads when the array allows it (a thread's elements all fall into its one bucket,
    // tid mod 128, whatever elements it visits)
    const bool v4 = ((((uintptr_t)vals) & 15) == 0);
    const int N4 = v4 ? (N >> 2) : 0;
    // ---- pass 1, common case: RAW bucket maxima, one FMNMX per element ----------------------------------
    // (the eligibility tests and the eligible count were 11 instructions per element = 15 % of the scoring
    // kernel, profiles/r1z_*).  The n-th largest raw maximum T is the threshold the exact pass would find
    // whenever it is itself eligible-and-decisive: dense mode -> T > -inf (n non-empty buckets, so n eligible
    // elements exist); sparse mode -> T > 0 (the n leading buckets have positive maxima, which are eligible, and a
    // bucket whose raw maximum is 0 has no eligible element above T).  Otherwise the exact pass below runs.
    {
        float m = -INFINITY;
#pragma unroll 4
        for (int q = tid; q < N4; q += NT) {
            const float4 v = reinterpret_cast<const float4 *>(vals)[q];
            m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
        for (int x = (N4 << 2) + tid; x < N; x += NT) m = fmaxf(m, vals[x]);
        if (m > -INFINITY) atomicMax(&fs->u.f.bmax[tid & (FS_BUCKETS - 1)], float_key(m));
    }
    __syncthreads();
    fs_nth_bucket_max(fs, n - 1);
    __syncthreads();
    uint32_t T = fs->T;
    int kk = n;
    const bool raw_ok = skip_zero ? (T > 0x80000000u) : (T != 0u);   // float_key(0.0f) == 0x80000000
    if (!raw_ok) {
        // ---- pass 1, exact: bucket maxima over eligible elements and the eligible count ------------------
        __syncthreads();
        if (tid < FS_BUCKETS) fs->u.f.bmax[tid] = 0u;
        __syncthreads();
        float m = -INFINITY;
        int ne = 0;
        auto see = [&](float v) {
            const bool ok = skip_zero ? ((v > -INFINITY) && (v != 0.0f)) : (v > -INFINITY);
            m = ok ? fmaxf(m, v) : m;
            ne += ok;
        };
        for (int q = tid; q < N4; q += NT) {
            const float4 v = reinterpret_cast<const float4 *>(vals)[q];
            see(v.x); see(v.y); see(v.z); see(v.w);
        }
        for (int x = (N4 << 2) + tid; x < N; x += NT) see(vals[x]);
        if (ne) atomicMax(&fs->u.f.bmax[tid & (FS_BUCKETS - 1)], float_key(m));
        ne = warp_sum_i(ne);
        if (lane == 0) fs->wtot[0][warp] = ne;
        __syncthreads();
        int n_elig = 0;
        for (int w = 0; w < nwarps; ++w) n_elig += fs->wtot[0][w];
        kk = min(n, n_elig);
        if (kk == 0) return 0;
        fs_nth_bucket_max(fs, kk - 1);
        __syncthreads();
        T = fs->T;
        if (T == 0u) T = 1u;   // fewer non-empty buckets than kk: below every eligible key
    }
    // T is the key of an eligible element here, or 1 (kk <= number of non-empty buckets is not guaranteed: with
    // fewer non-empty buckets than kk the kk-th maximum is 0 -> T = 1, below every eligible key)
    const float Tf = T > 1u ? fs_key_to_float(T) : -INFINITY;
    // ---- pass 2: collect everything >= T ---------------------------------------------------------------
    auto consider = [&](int x, float v, bool valid) {
        bool take = valid && (v >= Tf) && (v > -INFINITY);
        if (skip_zero) take = take && (v != 0.0f);
        const unsigned bal = __ballot_sync(0xffffffffu, take);
        if (bal) {
            const uint32_t key = take ? float_key(v) : 0u;
            const unsigned bgt = __ballot_sync(0xffffffffu, key > T);
            int basepos = 0;
            if (lane == 0) {
                basepos = atomicAdd(&fs->cnt_ge, __popc(bal));
                if (bgt) atomicAdd(&fs->cnt_gt, __popc(bgt));
            }
            basepos = __shfl_sync(0xffffffffu, basepos, 0);
            const int pos = take ? basepos + __popc(bal & ((1u << lane) - 1u)) : FS_LIST;
            if (pos < FS_LIST) { fs->u.f.lkey[pos] = key; fs->u.f.lidx[pos] = x; }
        }
    };
    {
        const int N4r = ((N4 + NT - 1) / NT) * NT;   // whole-CTA iterations: the ballots need converged warps
        for (int q = tid; q < N4r; q += NT) {
            float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            const bool valid = q < N4;
            if (valid) v = reinterpret_cast<const float4 *>(vals)[q];
            const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
            if (__any_sync(0xffffffffu, valid && m4 >= Tf)) {
                const int x0 = q << 2;
                consider(x0, v.x, valid); consider(x0 + 1, v.y, valid); consider(x0 + 2, v.z, valid); consider(x0 + 3, v.w, valid);
            }
        }
        for (int base = (N4 << 2); base < N; base += NT) {
            const int x = base + tid;
            consider(x, x < N ? vals[x] : -INFINITY, x < N);
        }
    }
    __syncthreads();
    const int total = fs->cnt_ge;
    if (total > FS_LIST) return block_top_n_fast(N, n, key_of, fs, out_idx, out_key);  // massive ties: generic path
    // ---- rank sort -------------------------------------------------------------------------------------
    for (int e = tid; e < total; e += NT) {
        const uint32_t ke = fs->u.f.lkey[e];
        const int ie = fs->u.f.lidx[e];
        int rank = 0;
        for (int f = 0; f < total; ++f) {
            const uint32_t kf = fs->u.f.lkey[f];
            const int jf = fs->u.f.lidx[f];
            rank += (kf > ke) || (kf == ke && jf > ie);
        }
        if (rank < kk) { out_idx[rank] = ie; if (out_key) out_key[rank] = ke; }
    }
    __syncthreads();
    return kk;
}

}  // namespace rt
