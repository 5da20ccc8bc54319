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

}  // namespace rt
