// solve.cu -- K4: batched ElasticNet coordinate descent, one CTA per target item column.
//
// Replaces, per target column j (reference: /root/reference/rtrec/models/internal/slim_elastic.py):
//   FeatureSelectionWrapper.fit  :139-154  (score = X^T y, top-n candidates, solve, scatter coef)
//   ElasticNet(...).fit          :195-208  -> sklearn 1.9.0 _cd_fast.pyx:653-1005
//
// The coordinate SEQUENCE of sklearn is replayed exactly (xorshift32 draws, gap-safe screening,
// stopping rule) but on the Gram matrix instead of the residual ("Gram-form replay"):
//     q_c = G[j][c], n2_c = G[c][c], h = G w (incremental), tmp = q_c - h_c + w_c n2_c,
//     R.R = yy - 2 w.q + w.h, R.y = yy - w.q, XtA_c = q_c - h_c - b w_c.
// The CPU model of this exact algorithm is oracle/gram_model.c; DESIGN.md section 3 has the
// derivation and the measured agreement with the residual-form reference.
//
// Thread organisation: warp 0 walks the draw sequence 32 draws at a time (a visit whose
// coefficient does not change is a no-op, so a whole batch of no-ops retires in one step; the
// first changing visit in the batch is applied and the walk resumes after it).  Vector work
// (h += d*G[c,:], gap, screening, selection) is spread over the whole CTA.
#include "block_select.cuh"
#include "common.cuh"

namespace rt {

struct SolveArgs {
    const float *G;
    // owner-rows layout of the multi-GPU fit (rt_slim_solve_rows): row i of G lives in buffer rowslot[i] >> 24 (own
    // memory or a CUDA IPC mapping of a peer GPU, read over NVLink) at local row rowslot[i] & 0xffffff.  rowslot ==
    // nullptr: one dense matrix G.
    const float *bases[RT_MAX_PEERS];
    const int *rowslot;
    const float *diag;  // G[i][i] gathered contiguously
    int64_t ldg;
    int n_items;
    const int *targets;
    int n_targets;
    int nn;  // 0 = all items
    int NU;  // universe size
    const int *sel_in;
    int *sel_out;
    double a, b, d_w_tol;
    int max_iter, positive, nonneg;
    const uint32_t *rng;
    int64_t *out_off;
    int *out_cnt;
    int *out_rows;
    float *out_vals;
    int64_t out_cap;
    unsigned long long *cursor;  // [0] = append cursor (all-items mode), [1] = next target
    int *stats;
    char *scratch;
    size_t scratch_per_cta;
    const int *only_flagged;  // when set, the block kernel solves only targets t with only_flagged[t] != 0
    int flag_mod;             // test hook (solve_impl = 3): the warp kernel hands every flag_mod-th target to the block kernel
    int skip_trivial;  // nn mode: targets without a live coordinate return no pairs (rt_fit_config.skip_trivial)
    const float *rowmax;   // optional: largest off-diagonal Gram entry per item (rt_fit_config.rowmax_ptr)
    const int *item_flag;  // pruned fit: items whose Gram row exists (others are trivial targets and are never dereferenced)
    int hit_mode;      // all-features mode: shared-memory bitmap of the live positions of active[] + per-sweep hit lists
    int bm_off;        // byte offset of the bitmap in dynamic shared memory (hit_mode)
    int hot_in_smem;  // per-visit arrays live in dynamic shared memory
    int use_gs;       // dense live x live Gram block cached in shared memory (nn mode)
};

__device__ __forceinline__ const float *g_row(const SolveArgs &A, int i) {
    if (A.rowslot) {
        const int s = __ldg(A.rowslot + i);
        return A.bases[s >> 24] + (size_t)(s & 0xffffff) * A.ldg;
    }
    return A.G + (size_t)i * A.ldg;
}

struct Misc {
    FastSelScratch fsel;
    double red[4][32];
    int wtot[2][2][32];
    int t;
    // event produced by the draw walker
    int ev_kind;  // 0 = sweep finished, 1 = coefficient update
    int ev_slot;
    double ev_delta, ev_wnew;
    double w_max, d_w_max;
    double gap, dual_norm;
};

__device__ __forceinline__ double block_max_d(double v, Misc *ms) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) ms->red[0][warp] = v;
    __syncthreads();
    double r = ms->red[0][0];
    for (int w = 1; w < nw; ++w) r = fmax(r, ms->red[0][w]);
    return r;
}

__device__ __forceinline__ void block_sum4(double &v0, double &v1, double &v2, double &v3, Misc *ms) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v0 = warp_sum(v0); v1 = warp_sum(v1); v2 = warp_sum(v2); v3 = warp_sum(v3);
    __syncthreads();
    if (lane == 0) { ms->red[0][warp] = v0; ms->red[1][warp] = v1; ms->red[2][warp] = v2; ms->red[3][warp] = v3; }
    __syncthreads();
    v0 = ms->red[0][0]; v1 = ms->red[1][0]; v2 = ms->red[2][0]; v3 = ms->red[3][0];
    for (int w = 1; w < nw; ++w) { v0 += ms->red[0][w]; v1 += ms->red[1][w]; v2 += ms->red[2][w]; v3 += ms->red[3][w]; }
}

// MAXT = 128 (feature selection: eight CTAs per SM, the candidate selection wants 128 threads) or 512 (all features: the
// universe loops over tens of thousands of coordinates are what a column costs)
template <int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 128 ? 8 : 2) slim_solve_kernel(SolveArgs A) {
    extern __shared__ __align__(16) char dyn_smem[];
    __shared__ Misc ms;
    const int tid = threadIdx.x, NT = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int nwarps = NT >> 5;
    const int NU = A.NU;
    const bool nnmode = A.nn > 0;
    const double a = A.a, b = A.b;

    // ---- carve state ----------------------------------------------------------------------
    // hot (touched per draw): active[NU] i32, live_slot[NU] i32, w[NU] f64, h[NU] f64, qv[NU] f32, n2[NU] f32
    // cold: feat[NU] i32, live[NU] i32, excl[NU] u8, xta[NU] f32, list[NU] i32, ckey[NU] u32, cidx[NU] i32
    char *cold = A.scratch + (size_t)blockIdx.x * A.scratch_per_cta;
    char *hot = A.hot_in_smem ? dyn_smem : cold;
    size_t ho = 0;
    auto take = [&](char *&base, size_t &off, size_t bytes) { char *p = base + off; off += (bytes + 15) & ~(size_t)15; return p; };
    double *w = (double *)take(hot, ho, sizeof(double) * NU);
    double *h = (double *)take(hot, ho, sizeof(double) * NU);
    int *active = (int *)take(hot, ho, sizeof(int) * NU);
    int *live_slot = (int *)take(hot, ho, sizeof(int) * NU);
    float *qv = (float *)take(hot, ho, sizeof(float) * NU);
    float *n2 = (float *)take(hot, ho, sizeof(float) * NU);
    float *Gs = nullptr;
    if (A.use_gs) Gs = (float *)take(hot, ho, sizeof(float) * (size_t)NU * NU);
    size_t co = A.hot_in_smem ? 0 : ho;
    int *feat = (int *)take(cold, co, sizeof(int) * NU);
    int *live = (int *)take(cold, co, sizeof(int) * NU);
    float *xta = (float *)take(cold, co, sizeof(float) * NU);
    int *list = (int *)take(cold, co, sizeof(int) * NU);
    uint32_t *ckey = (uint32_t *)take(cold, co, sizeof(uint32_t) * NU);
    int *cidx = (int *)take(cold, co, sizeof(int) * NU);
    unsigned char *excl = (unsigned char *)take(cold, co, NU);
    // hit mode (all features, few live coordinates): bit idx of bm = "active[idx] is a live coordinate"; a sweep then scans
    // its n_active draws against the bitmap with every warp (no dependent global loads) and the sequential walker only
    // sees the handful of draws that land on a live coordinate -- same visits in the same order as the plain walk
    constexpr int HITCAP = 256;
    uint32_t *bm = A.hit_mode ? (uint32_t *)(dyn_smem + A.bm_off) : nullptr;
    int *whits = A.hit_mode ? (int *)(dyn_smem + A.bm_off + (((size_t)(NU + 31) / 32 * 4 + 15) & ~(size_t)15)) : nullptr;
    bool bm_valid = false;


    for (;;) {
        __syncthreads();
        if (tid == 0) ms.t = (int)atomicAdd(&A.cursor[1], 1ull);
        __syncthreads();
        const int t = ms.t;
        if (t >= A.n_targets) break;
        if (A.only_flagged && !A.only_flagged[t]) continue;
        const int j = A.targets[t];
        if (nnmode && A.skip_trivial && ((A.rowmax && !((double)A.rowmax[j] > A.a)) || (A.item_flag && !A.item_flag[j]))) {
            // no live coordinate (row maximum from rt_gram_finish_rowmax / _live), or pruned fit: no row -- the zero column
            if (tid == 0) {
                A.out_off[t] = (int64_t)t * NU; A.out_cnt[t] = 0;
                if (A.stats) { A.stats[(size_t)t * 4 + 0] = 0; A.stats[(size_t)t * 4 + 1] = 0; A.stats[(size_t)t * 4 + 2] = 1; A.stats[(size_t)t * 4 + 3] = 0; }
            }
            continue;
        }
        const float *gj = g_row(A, j);

        // ---- universe ------------------------------------------------------------------------
        if (nnmode) {
            if (A.sel_in) {
                for (int k = tid; k < NU; k += NT) feat[k] = A.sel_in[(size_t)t * A.nn + k];
                __syncthreads();
            } else {
                // every item is eligible (float_key is never 0); the target itself scores 0 (its column is zeroed)
                auto key_of = [&](int i) -> uint32_t { return float_key(i == j ? 0.0f : gj[i]); };
                block_top_n_fast(A.n_items, NU, key_of, &ms.fsel, feat, (uint32_t *)nullptr, ckey, cidx);
            }
            if (A.sel_out) {
                for (int k = tid; k < A.nn; k += NT) A.sel_out[(size_t)t * A.nn + k] = k < NU ? feat[k] : -1;
            }
        }
        auto F = [&](int k) -> int { return nnmode ? feat[k] : k; };

        const double yy = (double)gj[j];
        const double tol_abs = A.d_w_tol * yy;

        // ---- live set (ordered compaction over the universe) ----------------------------------
        int m = 0;
        {
            int buf = 0;
            for (int base = 0; base < NU; base += NT, buf ^= 1) {
                const int k = base + tid;
                bool is_live = false;
                float qk = 0.f, nk = 0.f;
                int f = -1;
                if (k < NU) {
                    f = F(k);
                    if (f != j) { qk = gj[f]; nk = A.diag[f]; }
                    is_live = nk > 0.f && (!(A.positive && A.nonneg) || (double)qk > a);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, is_live);
                if (lane == 0) ms.wtot[buf][0][warp] = __popc(bal);
                __syncthreads();
                int off = m, tot = 0;
                for (int ww = 0; ww < nwarps; ++ww) { const int c = ms.wtot[buf][0][ww]; if (ww < warp) off += c; tot += c; }
                if (k < NU) {
                    if (is_live) {
                        const int s = off + __popc(bal & ((1u << lane) - 1u));
                        live_slot[k] = s; live[s] = k; w[s] = 0.0; h[s] = 0.0; qv[s] = qk; n2[s] = nk;
                    } else live_slot[k] = -1;
                }
                m += tot;
            }
            __syncthreads();
        }
        if (A.use_gs) {
            for (int e = tid; e < m * m; e += NT) {
                const int r = e / m, c = e - r * m;
                Gs[e] = g_row(A, F(live[r]))[F(live[c])];
            }
            __syncthreads();
        }
        // G entry between two live slots.  Without the shared-memory block the callers sweep sr for a fixed sc: by symmetry
        // (rt_gram_finish writes G[a][b] == G[b][a] bit for bit) the entry is read from ROW sc at column live[sr], so that the
        // threads of a warp gather ascending positions of ONE row instead of one cache line from each of m rows
        auto GL = [&](int sr, int sc) -> double {
            return A.use_gs ? (double)Gs[sr * m + sc] : (double)g_row(A, F(live[sc]))[F(live[sr])];
        };

        int n_active = 0, n_iter = 0, n_gap = 0, draws = 0;

        // ---- duality gap over the whole universe (fills xta) ---------------------------------
        auto eval_gap = [&]() {
            double wq = 0, wh = 0, l1 = 0, l2 = 0;
            for (int s = tid; s < m; s += NT) {
                const double ws = w[s];
                wq += ws * (double)qv[s]; wh += ws * h[s]; l1 += fabs(ws); l2 += ws * ws;
            }
            block_sum4(wq, wh, l1, l2, &ms);
            // support list (ordered) for the non-live rows
            int ns = 0;
            {
                int buf = 0;
                for (int base = 0; base < m; base += NT, buf ^= 1) {
                    const int s = base + tid;
                    const bool nz = s < m && w[s] != 0.0;
                    const unsigned bal = __ballot_sync(0xffffffffu, nz);
                    if (lane == 0) ms.wtot[buf][0][warp] = __popc(bal);
                    __syncthreads();
                    int off = ns, tot = 0;
                    for (int ww = 0; ww < nwarps; ++ww) { const int c = ms.wtot[buf][0][ww]; if (ww < warp) off += c; tot += c; }
                    if (nz) list[off + __popc(bal & ((1u << lane) - 1u))] = s;
                    ns += tot;
                }
                __syncthreads();
            }
            double dn = -INFINITY;
            for (int k = tid; k < NU; k += NT) {
                const int s = live_slot[k];
                double v;
                if (s >= 0) v = (double)qv[s] - h[s] - b * w[s];
                else {
                    const int f = F(k);
                    if (f == j) v = 0.0;
                    else {
                        double hk = 0.0;
                        for (int e = 0; e < ns; ++e) {
                            const int sc = list[e];
                            hk += (double)g_row(A, F(live[sc]))[f] * w[sc];
                        }
                        v = (double)gj[f] - hk;
                    }
                }
                xta[k] = (float)v;
                const double av = A.positive ? v : fabs(v);
                dn = fmax(dn, av);
            }
            dn = block_max_d(dn, &ms);
            const double Rn = yy - 2.0 * wq + wh, Ry = yy - wq;
            const double primal = 0.5 * (Rn + b * l2) + a * l1;
            const double scale = dn > a ? a / dn : 1.0;
            const double dualv = -0.5 * scale * scale * (Rn + b * l2) + scale * Ry;
            if (tid == 0) { ms.gap = primal - dualv; ms.dual_norm = dn; }
            __syncthreads();
            ++n_gap;
        };

        // ---- gap-safe screening; rebuilds active[] in ascending universe order ---------------
        auto screen = [&](bool first) {
            const double gap = ms.gap, dn = ms.dual_norm;
            const double denom = a > dn ? a : dn;
            const double thr = sqrt(2.0 * gap) / a;
            int na = 0, nd = 0, buf = 0;
            for (int base = 0; base < NU; base += NT, buf ^= 1) {
                const int k = base + tid;
                bool keep = false, drop = false;
                if (k < NU) {
                    const int s = live_slot[k];
                    const int f = F(k);
                    const double nk = s >= 0 ? (double)n2[s] : (f == j ? 0.0 : (double)A.diag[f]);
                    bool consider;
                    if (first) { consider = nk != 0.0; if (!consider) excl[k] = 1; }
                    else consider = !excl[k];
                    if (consider) {
                        const double theta = (double)xta[k] / denom;
                        const double dk = (1.0 - fabs(theta)) / sqrt(nk + b);
                        if (dk <= thr) { keep = true; excl[k] = 0; }
                        else { excl[k] = 1; drop = (s >= 0 && w[s] != 0.0); }
                    }
                }
                const unsigned bk = __ballot_sync(0xffffffffu, keep);
                const unsigned bd = __ballot_sync(0xffffffffu, drop);
                if (lane == 0) { ms.wtot[buf][0][warp] = __popc(bk); ms.wtot[buf][1][warp] = __popc(bd); }
                __syncthreads();
                int offk = na, totk = 0, offd = nd, totd = 0;
                for (int ww = 0; ww < nwarps; ++ww) {
                    const int c = ms.wtot[buf][0][ww], d = ms.wtot[buf][1][ww];
                    if (ww < warp) { offk += c; offd += d; }
                    totk += c; totd += d;
                }
                if (keep) active[offk + __popc(bk & ((1u << lane) - 1u))] = k;
                if (drop) list[offd + __popc(bd & ((1u << lane) - 1u))] = live_slot[k];
                na += totk; nd += totd;
            }
            __syncthreads();
            // excluded coordinates that still carry weight: remove their contribution from h
            for (int e = 0; e < nd; ++e) {
                const int sc = list[e];
                const double wsc = w[sc];
                __syncthreads();
                for (int r = tid; r < m; r += NT) h[r] -= wsc * GL(r, sc);
                if (tid == 0) w[sc] = 0.0;
                __syncthreads();
            }
            n_active = na;
            bm_valid = false;
            if (A.hit_mode && (long long)m * 16 < na) {
                for (int base = 0; base < na; base += NT) {   // NT = 128: base + 32 * warp is a multiple of 32
                    const int idx = base + tid;
                    const bool lv = idx < na && live_slot[active[idx]] >= 0;
                    const unsigned word = __ballot_sync(0xffffffffu, lv);
                    if (lane == 0) bm[(base >> 5) + warp] = word;
                }
                __syncthreads();
                bm_valid = true;
            }
        };

        eval_gap();
        bool done = ms.gap <= tol_abs;
        if (!done && m == 0) { done = true; n_iter = A.max_iter; }  // nothing can move: w stays 0
        if (!done) {
            screen(true);
            int64_t tdraw = 0;  // index of the next draw in the rng table
            for (int it = 0; it < A.max_iter; ++it) {
                // ---------------- one sweep: n_active draws ----------------
                int n_visit = n_active;
                bool hits = false;
                if (bm_valid) {
                    // phase A: warp w scans draws [w * seg, (w + 1) * seg) and lists, in order, those that hit a live position
                    const int seg = (n_active + nwarps - 1) / nwarps;
                    const int d0 = warp * seg, d1 = min(d0 + seg, n_active);
                    int cntw = 0;
                    for (int db = d0; db < d1; db += 32) {
                        const int d = db + lane;
                        bool hit = false;
                        int idx = 0;
                        if (d < d1) {
                            idx = (int)(A.rng[tdraw + d] % (uint32_t)n_active);
                            hit = (bm[idx >> 5] >> (idx & 31)) & 1u;
                        }
                        const unsigned bal = __ballot_sync(0xffffffffu, hit);
                        if (bal) {
                            const int pos = cntw + __popc(bal & ((1u << lane) - 1u));
                            if (hit && pos < HITCAP) whits[warp * HITCAP + pos] = idx;
                            cntw += __popc(bal);
                        }
                    }
                    if (lane == 0) ms.wtot[0][1][warp] = cntw;
                    __syncthreads();
                    int tot = 0;
                    bool fits = true;
                    for (int ww = 0; ww < nwarps; ++ww) { tot += ms.wtot[0][1][ww]; fits = fits && ms.wtot[0][1][ww] <= HITCAP; }
                    if (fits) { hits = true; n_visit = tot; }   // (a warp list overflowed: plain walk for this sweep)
                }
                int v = 0;
                double wmax_l = 0.0, dwmax_l = 0.0;  // per-lane running maxima (warp 0)
                for (;;) {
                    if (warp == 0) {
                        int kind = 0;
                        while (v < n_visit) {
                            const int nb = min(32, n_visit - v);
                            bool upd = false;
                            int s = -1;
                            double wc = 0.0, wn = 0.0;
                            if (lane < nb) {
                                int k;
                                if (hits) {
                                    // visit v + lane of the concatenated per-warp lists
                                    int i = v + lane, ww = 0;
                                    while (i >= ms.wtot[0][1][ww]) { i -= ms.wtot[0][1][ww]; ++ww; }
                                    k = active[whits[ww * HITCAP + i]];
                                } else {
                                    const uint32_t r = A.rng[tdraw + v + lane];
                                    k = active[r % (uint32_t)n_active];
                                }
                                s = live_slot[k];
                                if (s >= 0) {
                                    wc = w[s];
                                    const double nk = (double)n2[s];
                                    const double tmp = (double)qv[s] - h[s] + wc * nk;
                                    if (A.positive && tmp < 0.0) wn = 0.0;
                                    else {
                                        double mag = fabs(tmp) - a;
                                        if (!(mag > 0.0)) mag = 0.0;
                                        wn = (tmp > 0.0 ? mag : (tmp < 0.0 ? -mag : 0.0)) / (nk + b);
                                    }
                                    upd = wn != wc;
                                }
                            }
                            const unsigned mask = __ballot_sync(0xffffffffu, upd);
                            if (mask == 0u) {
                                wmax_l = fmax(wmax_l, fabs(wn));
                                v += nb;
                                continue;
                            }
                            const int L0 = __ffs(mask) - 1;
                            if (lane <= L0) wmax_l = fmax(wmax_l, fabs(wn));
                            if (lane == L0) {
                                dwmax_l = fmax(dwmax_l, fabs(wn - wc));
                                ms.ev_slot = s; ms.ev_delta = wn - wc; ms.ev_wnew = wn;
                            }
                            v += L0 + 1;
                            kind = 1;
                            break;
                        }
                        if (lane == 0) ms.ev_kind = kind;
                        if (kind == 0) {
                            const double wm = warp_max(wmax_l), dm = warp_max(dwmax_l);
                            if (lane == 0) { ms.w_max = wm; ms.d_w_max = dm; }
                        }
                    }
                    __syncthreads();
                    if (ms.ev_kind == 0) break;
                    {
                        const int sc = ms.ev_slot;
                        const double d = ms.ev_delta;
                        for (int r = tid; r < m; r += NT) h[r] += d * GL(r, sc);
                        if (tid == 0) w[sc] = ms.ev_wnew;
                    }
                    __syncthreads();
                }
                tdraw += n_active;
                draws += n_active;
                n_iter = it + 1;
                const double w_max = ms.w_max, d_w_max = ms.d_w_max;
                if (w_max == 0.0 || d_w_max / w_max <= A.d_w_tol || it == A.max_iter - 1) {
                    eval_gap();
                    if (ms.gap <= tol_abs) break;
                    screen(false);
                }
            }
        }

        // ---- output ---------------------------------------------------------------------------
        if (A.stats && tid == 0) {
            A.stats[(size_t)t * 4 + 0] = n_iter; A.stats[(size_t)t * 4 + 1] = draws;
            A.stats[(size_t)t * 4 + 2] = n_gap; A.stats[(size_t)t * 4 + 3] = m;
        }
        if (nnmode) {
            const int64_t off = (int64_t)t * NU;
            for (int k = tid; k < NU; k += NT) {
                const int s = live_slot[k];
                A.out_rows[off + k] = feat[k];
                A.out_vals[off + k] = s >= 0 ? (float)w[s] : 0.0f;
            }
            if (tid == 0) { A.out_off[t] = off; A.out_cnt[t] = NU; }
        } else {
            // non-zero coefficients in ascending row: live slots are already ascending in k
            int cnt = 0;
            {
                int buf = 0;
                for (int base = 0; base < m; base += NT, buf ^= 1) {
                    const int s = base + tid;
                    const bool nz = s < m && (float)w[s] != 0.0f;
                    const unsigned bal = __ballot_sync(0xffffffffu, nz);
                    if (lane == 0) ms.wtot[buf][0][warp] = __popc(bal);
                    __syncthreads();
                    int off = cnt, tot = 0;
                    for (int ww = 0; ww < nwarps; ++ww) { const int c = ms.wtot[buf][0][ww]; if (ww < warp) off += c; tot += c; }
                    if (nz) list[off + __popc(bal & ((1u << lane) - 1u))] = s;
                    cnt += tot;
                }
                __syncthreads();
            }
            if (tid == 0) {
                const unsigned long long o = atomicAdd(&A.cursor[0], (unsigned long long)cnt);
                A.out_off[t] = (int64_t)o; A.out_cnt[t] = cnt;
                ms.ev_delta = (double)o;  // broadcast
            }
            __syncthreads();
            const int64_t off = (int64_t)ms.ev_delta;
            if (off + cnt <= A.out_cap) {
                for (int e = tid; e < cnt; e += NT) {
                    const int s = list[e];
                    A.out_rows[off + e] = live[s];
                    A.out_vals[off + e] = (float)w[s];
                }
            }
        }
    }
}


// ================================================================================================
// Warp-per-column solver (nn mode, universe <= 64 candidates).
//
// Same algorithm and the same arithmetic per coordinate visit as slim_solve_kernel; what changes is who
// does it.  The block kernel spends 41 % of its time waiting on the two streaming passes of the target's
// Gram row (128 threads, one 4-byte load each per step) and 29 % at CTA barriers while warp 0 walks the
// draw sequence (profiles/r1l_*).  With nn <= 64 every per-column vector (universe, live set, w, h) has at
// most two elements per lane, so ONE warp can own a column end to end: no CTA barrier anywhere, sixteen
// columns in flight per SM instead of eight, and the row is streamed with 16-byte loads, eight in flight
// per lane.  Candidate selection keeps the threshold idea of block_top_n_fast (128 strided bucket maxima ->
// lower bound T of the n-th largest score -> collect everything >= T -> rank sort), with the bucket maxima
// in registers (four per lane) and comparisons in the float domain.  Columns whose candidate list would
// overflow (more than 512 scores above T) are flagged and solved by the block kernel afterwards.
constexpr int SW_MAXU = 64;
constexpr int SW_LIST = 512;

struct WarpScratch {           // carved per warp out of dynamic shared memory
    double w[SW_MAXU], h[SW_MAXU];
    float qv[SW_MAXU], n2[SW_MAXU], xta[SW_MAXU];
    int feat[SW_MAXU], live[SW_MAXU], live_slot[SW_MAXU], active[SW_MAXU], list[SW_MAXU];
    unsigned char excl[SW_MAXU];
    // followed by max(SW_LIST * 8, NU * NU * 4) bytes: candidate list during selection, then the live x live Gram block
};

__device__ __forceinline__ float sw_key_to_float(uint32_t k) {
    const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

__device__ __forceinline__ void sw_zero_at(float4 &v, int d) {   // v[d] = 0 without dynamic register indexing
    v.x = d == 0 ? 0.0f : v.x; v.y = d == 1 ? 0.0f : v.y; v.z = d == 2 ? 0.0f : v.z; v.w = d == 3 ? 0.0f : v.w;
}

// ordered compaction inside a warp: returns the slot of this lane's element (valid iff flag) and adds the total to *count
__device__ __forceinline__ int sw_compact(bool flag, int lane, int &count) {
    const unsigned bal = __ballot_sync(0xffffffffu, flag);
    const int pos = count + __popc(bal & ((1u << lane) - 1u));
    count += __popc(bal);
    return pos;
}

__global__ void __launch_bounds__(512, 1) slim_solve_warp_kernel(SolveArgs A, int *__restrict__ need_block, int per_warp_bytes) {
    extern __shared__ __align__(16) char sw_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    char *base = sw_smem + (size_t)warp * per_warp_bytes;
    WarpScratch *S = (WarpScratch *)base;
    char *tail = base + ((sizeof(WarpScratch) + 15) & ~(size_t)15);
    uint32_t *lkey = (uint32_t *)tail;
    int *lidx = (int *)(tail + sizeof(uint32_t) * SW_LIST);
    float *Gs = (float *)tail;
    const int NU = A.NU;
    const double a = A.a, b = A.b;
    const float *G = A.G;
    const int64_t ld = A.ldg;
    const int N = A.n_items;
    const bool vec4 = ((ld & 3) == 0) && (A.rowslot != nullptr || (((uintptr_t)G) & 15) == 0);  // row buffers are 16-byte aligned

    for (;;) {
        int t = 0;
        if (lane == 0) t = (int)atomicAdd(&A.cursor[1], 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= A.n_targets) break;
        const int j = A.targets[t];
        if (A.skip_trivial && ((A.rowmax && !((double)A.rowmax[j] > a)) || (A.item_flag && !A.item_flag[j]))) {
            // no entry of the Gram row is above the L1 threshold (known from rt_gram_finish_rowmax, or the item is not a
            // Cauchy-Schwarz candidate of the pruned fit and has no row): the zero column, and the row is not even read
            if (lane == 0) {
                A.out_off[t] = (int64_t)t * NU; A.out_cnt[t] = 0;
                if (A.stats) { A.stats[(size_t)t * 4 + 0] = 0; A.stats[(size_t)t * 4 + 1] = 0; A.stats[(size_t)t * 4 + 2] = 1; A.stats[(size_t)t * 4 + 3] = 0; }
            }
            continue;
        }
        const float *gj = g_row(A, j);
        __syncwarp();

        // ---- candidates ------------------------------------------------------------------------------
        bool overflow = A.flag_mod > 0 && (t % A.flag_mod) == 0;
        if (overflow) {
        } else if (A.sel_in) {
            for (int k = lane; k < NU; k += 32) S->feat[k] = A.sel_in[(size_t)t * A.nn + k];
        } else {
            // pass 1: 128 bucket maxima, bucket = (lane, index & 3); score of the target itself is 0
            float bm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            const int N4 = vec4 ? (N >> 2) : 0;
            {
                int q = lane;
                for (; q + 7 * 32 < N4; q += 8 * 32) {
                    float4 v[8];
#pragma unroll
                    for (int r = 0; r < 8; ++r) v[r] = __ldg(reinterpret_cast<const float4 *>(gj) + q + 32 * r);
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const int i0 = (q + 32 * r) << 2;
                        sw_zero_at(v[r], j - i0);
                        bm[0] = fmaxf(bm[0], v[r].x); bm[1] = fmaxf(bm[1], v[r].y);
                        bm[2] = fmaxf(bm[2], v[r].z); bm[3] = fmaxf(bm[3], v[r].w);
                    }
                }
                for (; q < N4; q += 32) {
                    float4 v = __ldg(reinterpret_cast<const float4 *>(gj) + q);
                    const int i0 = q << 2;
                    sw_zero_at(v, j - i0);
                    bm[0] = fmaxf(bm[0], v.x); bm[1] = fmaxf(bm[1], v.y); bm[2] = fmaxf(bm[2], v.z); bm[3] = fmaxf(bm[3], v.w);
                }
                for (int i = (N4 << 2) + lane; i < N; i += 32) {
                    const float v = i == j ? 0.0f : gj[i];
                    const int c = i & 3;
                    bm[0] = c == 0 ? fmaxf(bm[0], v) : bm[0]; bm[1] = c == 1 ? fmaxf(bm[1], v) : bm[1];
                    bm[2] = c == 2 ? fmaxf(bm[2], v) : bm[2]; bm[3] = c == 3 ? fmaxf(bm[3], v) : bm[3];
                }
            }
            if (A.skip_trivial) {
                // largest feature score of the row = largest bucket maximum: nothing above the L1 threshold -> w = 0 with a
                // duality gap of exactly 0 before the first sweep; the column is finished here (no candidate pass)
                float rm = fmaxf(fmaxf(bm[0], bm[1]), fmaxf(bm[2], bm[3]));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) rm = fmaxf(rm, __shfl_xor_sync(0xffffffffu, rm, o));
                if (!((double)rm > a)) {
                    if (lane == 0) {
                        A.out_off[t] = (int64_t)t * NU; A.out_cnt[t] = 0;
                        if (A.stats) { A.stats[(size_t)t * 4 + 0] = 0; A.stats[(size_t)t * 4 + 1] = 0; A.stats[(size_t)t * 4 + 2] = 1; A.stats[(size_t)t * 4 + 3] = 0; }
                    }
                    continue;
                }
            }
            const int kk = min(NU, N);
            // T = kk-th largest bucket maximum (as an order-preserving key), by bitwise search
            uint32_t bk[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) bk[c] = bm[c] == -INFINITY ? 0u : float_key(bm[c]);
            uint32_t T = 0u;
            for (int bit = 31; bit >= 0; --bit) {
                const uint32_t cand = T | (1u << bit);
                int cnt = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) cnt += __popc(__ballot_sync(0xffffffffu, bk[c] >= cand));
                if (cnt >= kk) T = cand;
            }
            if (T == 0u) T = 1u;
            const float Tf = T > 1u ? sw_key_to_float(T) : -INFINITY;
            // pass 2: collect every score >= T (list order is irrelevant: it is rank-sorted below)
            int total = 0, n_gt = 0;
            auto visit = [&](int i, float v, bool strict, bool valid) {
                const bool gt = valid && v > Tf;
                const bool take = valid && (strict ? gt : (v >= Tf));
                const unsigned bal = __ballot_sync(0xffffffffu, take);
                if (bal) {
                    const int pos = total + __popc(bal & ((1u << lane) - 1u));
                    if (take && pos < SW_LIST) { lkey[pos] = float_key(v); lidx[pos] = i; }
                    total += __popc(bal);
                    n_gt += __popc(__ballot_sync(0xffffffffu, gt));
                }
            };
            auto scan = [&](bool strict) {
                total = 0; n_gt = 0;
                int q = lane;
                const int N4r = ((N4 + 31) / 32) * 32;   // whole-warp iterations (ballots inside)
                for (; q < N4r; q += 32) {
                    float4 v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                    const int i0 = q << 2;
                    if (q < N4) {
                        v = __ldg(reinterpret_cast<const float4 *>(gj) + q);
                        sw_zero_at(v, j - i0);
                    }
                    const bool ok = q < N4;
                    const float m4 = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
                    if (__any_sync(0xffffffffu, ok && (strict ? (m4 > Tf) : (m4 >= Tf)))) {
                        visit(i0 + 0, v.x, strict, ok); visit(i0 + 1, v.y, strict, ok);
                        visit(i0 + 2, v.z, strict, ok); visit(i0 + 3, v.w, strict, ok);
                    }
                }
                for (int base_i = (N4 << 2); base_i < N; base_i += 32) {
                    const int i = base_i + lane;
                    const float v = i < N ? (i == j ? 0.0f : gj[i]) : -INFINITY;
                    visit(i, v, strict, i < N);
                }
            };
            scan(false);
            if (total > SW_LIST) {
                const int c_gt = n_gt;
                if (c_gt > SW_LIST) overflow = true;
                else {
                    __syncwarp();
                    scan(true);                      // strictly larger ones only: exactly c_gt entries
                    if (c_gt < kk) {
                        // the remaining picks are the scores == T with the largest indices: scan downwards
                        const int need = kk - c_gt;
                        int taken = 0;
                        for (int top = ((N + 31) / 32) * 32; top > 0 && taken < need; top -= 32) {
                            const int x = top - 1 - lane;
                            const float v = x < N ? (x == j ? 0.0f : gj[x]) : -INFINITY;
                            const bool eq = x < N && v == Tf;
                            const unsigned bal = __ballot_sync(0xffffffffu, eq);
                            const int pos = taken + __popc(bal & ((1u << lane) - 1u));
                            if (eq && pos < need) { lkey[c_gt + pos] = T; lidx[c_gt + pos] = x; }
                            taken += __popc(bal);
                        }
                        total = kk;
                    } else total = c_gt;
                }
            }
            if (!overflow) {
                __syncwarp();
                // rank sort by (key desc, index desc); the first kk ranks are the candidates in pick order
                for (int e = lane; e < total; e += 32) {
                    const uint32_t ke = lkey[e];
                    const int ie = lidx[e];
                    int rank = 0;
                    for (int f = 0; f < total; ++f) {
                        const uint32_t kf = lkey[f];
                        const int jf = lidx[f];
                        rank += (kf > ke) || (kf == ke && jf > ie);
                    }
                    if (rank < kk) S->feat[rank] = ie;
                }
            }
        }
        if (overflow) {
            if (lane == 0) need_block[t] = 1;
            continue;
        }
        __syncwarp();
        if (A.sel_out) {
            for (int k = lane; k < A.nn; k += 32) A.sel_out[(size_t)t * A.nn + k] = k < NU ? S->feat[k] : -1;
        }

        const double yy = (double)gj[j];
        const double tol_abs = A.d_w_tol * yy;

        // ---- live set (ordered compaction over the universe) ---------------------------------------------
        int m = 0;
        for (int base_k = 0; base_k < NU; base_k += 32) {
            const int k = base_k + lane;
            bool is_live = false;
            float qk = 0.f, nk = 0.f;
            if (k < NU) {
                const int f = S->feat[k];
                if (f != j) { qk = gj[f]; nk = A.diag[f]; }
                is_live = nk > 0.f && (!(A.positive && A.nonneg) || (double)qk > a);
            }
            const int s = sw_compact(is_live, lane, m);
            if (k < NU) {
                if (is_live) { S->live_slot[k] = s; S->live[s] = k; S->w[s] = 0.0; S->h[s] = 0.0; S->qv[s] = qk; S->n2[s] = nk; }
                else S->live_slot[k] = -1;
            }
        }
        __syncwarp();
        // live x live Gram block (symmetric: load the lower triangle, store both halves).  The loads are random
        // 4-byte reads of a matrix far larger than L2: four per lane are issued before any is consumed.
        {
            const int n_tri = m * (m + 1) / 2;
            for (int e0 = lane; e0 < n_tri; e0 += 4 * 32) {
                float v[4];
                int rr[4], cc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = e0 + 32 * q;
                    rr[q] = -1; cc[q] = 0; v[q] = 0.f;
                    if (e < n_tri) {
                        // e = r(r+1)/2 + c, c <= r
                        int r = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
                        while (r * (r + 1) / 2 > e) --r;
                        while ((r + 1) * (r + 2) / 2 <= e) ++r;
                        rr[q] = r; cc[q] = e - r * (r + 1) / 2;
                        v[q] = __ldg(g_row(A, S->feat[S->live[r]]) + S->feat[S->live[cc[q]]]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (rr[q] >= 0) { Gs[rr[q] * m + cc[q]] = v[q]; Gs[cc[q] * m + rr[q]] = v[q]; }
            }
        }
        __syncwarp();

        int n_active = 0, n_iter = 0, n_gap = 0, draws = 0;
        double gap = 0.0, dual_norm = 0.0;

        auto eval_gap = [&]() {
            double wq = 0, wh = 0, l1 = 0, l2 = 0;
            for (int s = lane; s < m; s += 32) {
                const double ws = S->w[s];
                wq += ws * (double)S->qv[s]; wh += ws * S->h[s]; l1 += fabs(ws); l2 += ws * ws;
            }
            wq = warp_sum(wq); wh = warp_sum(wh); l1 = warp_sum(l1); l2 = warp_sum(l2);
            int ns = 0;
            for (int base_s = 0; base_s < m; base_s += 32) {
                const int s = base_s + lane;
                const bool nz = s < m && S->w[s] != 0.0;
                const int pos = sw_compact(nz, lane, ns);
                if (nz) S->list[pos] = s;
            }
            __syncwarp();
            double dn = -INFINITY;
            for (int k = lane; k < NU; k += 32) {
                const int s = S->live_slot[k];
                double v;
                if (s >= 0) v = (double)S->qv[s] - S->h[s] - b * S->w[s];
                else {
                    const int f = S->feat[k];
                    if (f == j) v = 0.0;
                    else {
                        // G[support][f] for the support coordinates: ns independent loads, four in flight.  The entry is read
                        // from the SUPPORT item's row (G is symmetric): support items are live, so their rows exist in the
                        // pruned fit, where the row of a non-live candidate f may not
                        double hk = 0.0;
                        int e = 0;
                        for (; e + 3 < ns; e += 4) {
                            const int s0 = S->list[e], s1 = S->list[e + 1], s2 = S->list[e + 2], s3 = S->list[e + 3];
                            const float g0 = __ldg(g_row(A, S->feat[S->live[s0]]) + f), g1 = __ldg(g_row(A, S->feat[S->live[s1]]) + f);
                            const float g2 = __ldg(g_row(A, S->feat[S->live[s2]]) + f), g3 = __ldg(g_row(A, S->feat[S->live[s3]]) + f);
                            hk += (double)g0 * S->w[s0]; hk += (double)g1 * S->w[s1];
                            hk += (double)g2 * S->w[s2]; hk += (double)g3 * S->w[s3];
                        }
                        for (; e < ns; ++e) {
                            const int sc = S->list[e];
                            hk += (double)__ldg(g_row(A, S->feat[S->live[sc]]) + f) * S->w[sc];
                        }
                        v = (double)gj[f] - hk;
                    }
                }
                S->xta[k] = (float)v;
                dn = fmax(dn, A.positive ? v : fabs(v));
            }
            dn = warp_max(dn);
            const double Rn = yy - 2.0 * wq + wh, Ry = yy - wq;
            const double primal = 0.5 * (Rn + b * l2) + a * l1;
            const double scale = dn > a ? a / dn : 1.0;
            const double dualv = -0.5 * scale * scale * (Rn + b * l2) + scale * Ry;
            gap = primal - dualv; dual_norm = dn;
            ++n_gap;
            __syncwarp();
        };

        auto screen = [&](bool first) {
            const double denom = a > dual_norm ? a : dual_norm;
            const double thr = sqrt(2.0 * gap) / a;
            int na = 0, nd = 0;
            for (int base_k = 0; base_k < NU; base_k += 32) {
                const int k = base_k + lane;
                bool keep = false, drop = false;
                if (k < NU) {
                    const int s = S->live_slot[k];
                    const int f = S->feat[k];
                    const double nk = s >= 0 ? (double)S->n2[s] : (f == j ? 0.0 : (double)A.diag[f]);
                    bool consider;
                    if (first) { consider = nk != 0.0; if (!consider) S->excl[k] = 1; }
                    else consider = !S->excl[k];
                    if (consider) {
                        const double theta = (double)S->xta[k] / denom;
                        const double dk = (1.0 - fabs(theta)) / sqrt(nk + b);
                        if (dk <= thr) { keep = true; S->excl[k] = 0; }
                        else { S->excl[k] = 1; drop = (s >= 0 && S->w[s] != 0.0); }
                    }
                }
                const int pk = sw_compact(keep, lane, na);
                const int pd = sw_compact(drop, lane, nd);
                if (keep) S->active[pk] = k;
                if (drop) S->list[pd] = S->live_slot[k];
            }
            __syncwarp();
            for (int e = 0; e < nd; ++e) {
                const int sc = S->list[e];
                const double wsc = S->w[sc];
                for (int r = lane; r < m; r += 32) S->h[r] -= wsc * (double)Gs[r * m + sc];
                __syncwarp();
                if (lane == 0) S->w[sc] = 0.0;
                __syncwarp();
            }
            n_active = na;
        };

        eval_gap();
        bool done = gap <= tol_abs;
        if (!done && m == 0) { done = true; n_iter = A.max_iter; }
        if (!done) {
            screen(true);
            int64_t tdraw = 0;
            for (int it = 0; it < A.max_iter; ++it) {
                int v = 0;
                double wmax_l = 0.0, dwmax_l = 0.0;
                while (v < n_active) {
                    const int nb = min(32, n_active - v);
                    bool upd = false;
                    int s = -1;
                    double wc = 0.0, wn = 0.0;
                    if (lane < nb) {
                        const uint32_t r = A.rng[tdraw + lane];
                        const int k = S->active[r % (uint32_t)n_active];
                        s = S->live_slot[k];
                        if (s >= 0) {
                            wc = S->w[s];
                            const double nk = (double)S->n2[s];
                            const double tmp = (double)S->qv[s] - S->h[s] + wc * nk;
                            if (A.positive && tmp < 0.0) wn = 0.0;
                            else {
                                double mag = fabs(tmp) - a;
                                if (!(mag > 0.0)) mag = 0.0;
                                wn = (tmp > 0.0 ? mag : (tmp < 0.0 ? -mag : 0.0)) / (nk + b);
                            }
                            upd = wn != wc;
                        }
                    }
                    const unsigned mask = __ballot_sync(0xffffffffu, upd);
                    if (mask == 0u) {
                        wmax_l = fmax(wmax_l, fabs(wn));
                        v += nb; tdraw += nb;
                        continue;
                    }
                    const int L0 = __ffs(mask) - 1;
                    if (lane <= L0) wmax_l = fmax(wmax_l, fabs(wn));
                    if (lane == L0) dwmax_l = fmax(dwmax_l, fabs(wn - wc));
                    v += L0 + 1; tdraw += L0 + 1;
                    // apply the first changing visit of the batch
                    const int sc = __shfl_sync(0xffffffffu, s, L0);
                    const double d = __shfl_sync(0xffffffffu, wn - wc, L0);
                    const double wnew = __shfl_sync(0xffffffffu, wn, L0);
                    for (int r = lane; r < m; r += 32) S->h[r] += d * (double)Gs[r * m + sc];
                    if (lane == 0) S->w[sc] = wnew;
                    __syncwarp();
                }
                draws += n_active;
                n_iter = it + 1;
                const double w_max = warp_max(wmax_l), d_w_max = warp_max(dwmax_l);
                if (w_max == 0.0 || d_w_max / w_max <= A.d_w_tol || it == A.max_iter - 1) {
                    eval_gap();
                    if (gap <= tol_abs) break;
                    screen(false);
                }
            }
        }

        // ---- output ------------------------------------------------------------------------------------
        if (A.stats && lane == 0) {
            A.stats[(size_t)t * 4 + 0] = n_iter; A.stats[(size_t)t * 4 + 1] = draws;
            A.stats[(size_t)t * 4 + 2] = n_gap; A.stats[(size_t)t * 4 + 3] = m;
        }
        const int64_t off = (int64_t)t * NU;
        for (int k = lane; k < NU; k += 32) {
            const int s = S->live_slot[k];
            A.out_rows[off + k] = S->feat[k];
            A.out_vals[off + k] = s >= 0 ? (float)S->w[s] : 0.0f;
        }
        if (lane == 0) { A.out_off[t] = off; A.out_cnt[t] = NU; }
    }
}

// All-features mode with positive coefficients on non-negative data: a coordinate c can only leave 0 if G[j][c] > a
// ("live set").  A target whose Gram row holds no such entry has the solution w = 0 and a duality
// gap of exactly 0 at the start (primal = dual = yy/2), so sklearn's solver returns before its first sweep
// (_cd_fast.pyx gap check before the loop): the column is written here -- no coefficients, stats (0 sweeps, 0 draws,
// 1 gap evaluation, 0 live) -- and the CTA solver only sees the flagged rest.  One warp per target streams the row with
// 16-byte loads: the whole pass runs at HBM speed instead of three 128-thread passes per column over the universe.
__global__ void __launch_bounds__(256) live_prefilter_kernel(SolveArgs A, int *__restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int t = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (t >= A.n_targets) return;
    const int j = A.targets[t];
    const int N = A.n_items;
    const double a = A.a;
    const bool have_row = !A.item_flag || A.item_flag[j] != 0;
    const float *gj = have_row ? g_row(A, j) : nullptr;
    const bool vec4 = ((A.ldg & 3) == 0) && ((((uintptr_t)gj) & 15) == 0);
    bool any = false;
    if (have_row) {
    const int N4 = vec4 ? (N >> 2) : 0;
    for (int q = lane; q < N4 && !any; q += 32 * 4) {
        float4 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = (q + 32 * r < N4) ? __ldg(reinterpret_cast<const float4 *>(gj) + q + 32 * r) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i0 = (q + 32 * r) << 2;
            sw_zero_at(v[r], j - i0);
            any = any || (double)v[r].x > a || (double)v[r].y > a || (double)v[r].z > a || (double)v[r].w > a;
        }
    }
    for (int i = (N4 << 2) + lane; i < N; i += 32) any = any || (i != j && (double)gj[i] > a);
    }
    any = __any_sync(0xffffffffu, any);
    if (lane == 0) {
        flags[t] = any ? 1 : 0;
        if (!any) {
            A.out_off[t] = 0; A.out_cnt[t] = 0;
            if (A.stats) { A.stats[(size_t)t * 4 + 0] = 0; A.stats[(size_t)t * 4 + 1] = 0; A.stats[(size_t)t * 4 + 2] = 1; A.stats[(size_t)t * 4 + 3] = 0; }
        }
    }
}

__global__ void count_flags_kernel(const int *__restrict__ flags, int n, int *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int v = i < n ? (flags[i] != 0) : 0;
    v = warp_sum_i(v);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

__global__ void gather_diag_kernel(SolveArgs A, int n, float *diag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) diag[i] = g_row(A, i)[i];
}

// ---- xorshift32 table -------------------------------------------------------------------------
// GF(2) jump-ahead: state after t steps = M^t * state.  Each thread jumps to the start of its
// 64-draw segment with the precomputed powers M^(2^b) and then iterates.
__constant__ uint32_t c_xs_pow[40][32];  // column images of M^(2^b), b < 40

__global__ void rng_table_kernel(uint32_t seed, int64_t n, uint32_t *out) {
    const int64_t seg = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t start = seg * 64;
    if (start >= n) return;
    uint32_t s = seed == 0 ? 1u : seed;
    for (int bit = 0; bit < 40; ++bit) {
        if ((start >> bit) & 1) {
            uint32_t r = 0;
#pragma unroll
            for (int c = 0; c < 32; ++c) r ^= ((s >> c) & 1u) ? c_xs_pow[bit][c] : 0u;
            s = r;
        }
    }
    const int64_t end = start + 64 < n ? start + 64 : n;
    for (int64_t t = start; t < end; ++t) {
        s ^= s << 13; s ^= s >> 17; s ^= s << 5;
        out[t] = s & 0x7fffffffu;  // % (RAND_R_MAX + 1)
    }
}

static uint32_t xs_step(uint32_t s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

static int upload_xs_powers() {
    static bool done = false;
    if (done) return RT_OK;
    static uint32_t pw[40][32];
    for (int c = 0; c < 32; ++c) pw[0][c] = xs_step(1u << c);
    for (int bpow = 1; bpow < 40; ++bpow)
        for (int c = 0; c < 32; ++c) {
            // M^(2^b) e_c = M^(2^(b-1)) (M^(2^(b-1)) e_c)
            uint32_t v = pw[bpow - 1][c], r = 0;
            for (int k = 0; k < 32; ++k) if ((v >> k) & 1u) r ^= pw[bpow - 1][k];
            pw[bpow][c] = r;
        }
    RT_CUDA(cudaMemcpyToSymbol(c_xs_pow, pw, sizeof(pw)));
    done = true;
    return RT_OK;
}

}  // namespace rt

using namespace rt;

extern "C" int rt_rng_table(uint32_t seed, int64_t n, uint32_t *d_out, void *stream) {
    RT_ARG(n >= 0 && (n == 0 || d_out), "rng table output");
    if (n == 0) return RT_OK;
    int rc = upload_xs_powers();
    if (rc) return rc;
    const int64_t segs = (n + 63) / 64;
    const int bs = 128;
    rng_table_kernel<<<(unsigned)((segs + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(seed, n, d_out);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

namespace {
struct SolvePlan {
    int NU, NT, grid, hot_in_smem, use_gs, hit_mode, bm_off;
    size_t hot_bytes, cold_bytes, smem_bytes, scratch_per_cta;
};

SolvePlan make_plan(int n_items, int nn, int n_targets) {
    SolvePlan p;
    p.NU = nn > 0 ? (nn < n_items ? nn : n_items) : n_items;
    auto pad = [](size_t x) { return (x + 15) & ~(size_t)15; };
    const size_t NU = (size_t)p.NU;
    size_t hot = 2 * pad(8 * NU) + 4 * pad(4 * NU);
    p.use_gs = (nn > 0 && NU <= 160) ? 1 : 0;
    if (p.use_gs) hot += pad(4 * NU * NU);
    size_t cold = 6 * pad(4 * NU) + pad(NU);
    const int optin = rt::smem_optin();
    const size_t static_smem = sizeof(Misc) + 64;
    p.hot_in_smem = (hot + static_smem <= (size_t)optin - 1024) ? 1 : 0;
    p.hot_bytes = hot;
    p.cold_bytes = cold;
    p.smem_bytes = p.hot_in_smem ? hot : 0;
    // all-features mode: live-position bitmap + per-warp hit lists (16 warps x 256 draws) behind the hot arrays
    p.hit_mode = 0; p.bm_off = 0;
    if (nn == 0) {
        const size_t bm_bytes = pad((NU + 31) / 32 * 4) + 16 * 256 * sizeof(int);
        if (p.smem_bytes + bm_bytes + static_smem <= (size_t)optin - 1024) {
            p.hit_mode = 1; p.bm_off = (int)p.smem_bytes; p.smem_bytes += bm_bytes;
        }
    }
    p.scratch_per_cta = rt::align_up(cold + (p.hot_in_smem ? 0 : hot) + 256, 256);
    p.NT = nn > 0 ? 128 : 512;  // multiple of 128: the fast candidate selection uses 128 strided buckets
    // resident CTAs per SM, bounded by shared memory
    int per_sm = nn > 0 ? 8 : 2;
    if (p.hot_in_smem) {
        int fit = (int)(((size_t)optin) / (p.smem_bytes + static_smem + 1024));
        if (fit < 1) fit = 1;
        if (fit < per_sm) per_sm = fit;
    }
    int grid = rt::sm_count() * per_sm;
    if (grid > n_targets) grid = n_targets;
    if (grid < 1) grid = 1;
    p.grid = grid;
    return p;
}
}  // namespace

static int slim_solve_impl(const float *d_G, const void *const *h_bases, int32_t n_bases, const int32_t *d_rowslot,
                           const float *d_diag_in, const int32_t *d_item_flag,
                           int64_t ldg, int32_t n_items, const int32_t *d_targets,
                           int32_t n_targets, const rt_fit_config *cfg, const int32_t *d_sel_in,
                           const uint32_t *d_rng, int64_t rng_len, int32_t *d_sel_out, int64_t *d_out_off,
                           int32_t *d_out_cnt, int32_t *d_out_rows, float *d_out_vals, int64_t out_cap,
                           int64_t *h_needed, int32_t *d_stats, void *stream) {
    RT_ARG(cfg != nullptr, "cfg");
    RT_ARG(n_items > 0 && ldg >= n_items, "n_items/ldg");
    RT_ARG(n_targets >= 0, "n_targets");
    RT_ARG(cfg->nn >= 0, "nn");
    RT_ARG(cfg->alpha * cfg->l1_ratio > 0.0, "alpha*l1_ratio must be > 0 (L1 penalty; gap-safe screening path)");
    if (h_needed) *h_needed = 0;
    if (n_targets == 0) return RT_OK;
    RT_ARG((d_G || d_rowslot) && d_targets && d_rng && d_out_off && d_out_cnt && d_out_rows && d_out_vals, "null pointer");
    SolvePlan p = make_plan(n_items, cfg->nn, n_targets);
    RT_ARG(rng_len >= (int64_t)cfg->max_iter * p.NU + 64, "rng table too short");
    const size_t diag_bytes = rt::align_up((size_t)n_items * sizeof(float));
    void *d_workspace = rt::scratch(SCR_SOLVE, (size_t)p.grid * p.scratch_per_cta + 1024 + diag_bytes);
    if (!d_workspace) return RT_ERR_CUDA;
    if (cfg->nn > 0) RT_ARG(out_cap >= (int64_t)n_targets * p.NU, "out_cap < n_targets * nn");
    cudaStream_t st = (cudaStream_t)stream;

    SolveArgs A;
    A.G = d_G; A.rowslot = d_rowslot;
    for (int q = 0; q < RT_MAX_PEERS; ++q) A.bases[q] = d_rowslot ? (const float *)h_bases[q < n_bases ? q : 0] : d_G;
    A.ldg = ldg; A.n_items = n_items; A.targets = d_targets; A.n_targets = n_targets;
    A.nn = cfg->nn; A.NU = p.NU; A.sel_in = d_sel_in; A.sel_out = d_sel_out;
    A.a = (double)(float)(cfg->alpha * cfg->l1_ratio * (double)cfg->n_samples);
    A.b = (double)(float)(cfg->alpha * (1.0 - cfg->l1_ratio) * (double)cfg->n_samples);
    A.d_w_tol = (double)(float)cfg->tol;
    A.max_iter = cfg->max_iter; A.positive = cfg->positive; A.nonneg = cfg->nonneg;
    A.rng = d_rng; A.out_off = d_out_off; A.out_cnt = d_out_cnt; A.out_rows = d_out_rows;
    A.out_vals = d_out_vals; A.out_cap = out_cap; A.stats = d_stats;
    A.cursor = (unsigned long long *)d_workspace;
    float *d_diag = (float *)((char *)d_workspace + 1024);
    A.scratch = (char *)d_workspace + 1024 + diag_bytes;
    A.scratch_per_cta = p.scratch_per_cta;
    A.hot_in_smem = p.hot_in_smem; A.use_gs = p.use_gs;
    A.hit_mode = rt::option(rt::OPT_SOLVE_IMPL) == 1 ? 0 : p.hit_mode; A.bm_off = p.bm_off;
    A.only_flagged = nullptr;
    A.item_flag = d_item_flag;
    A.rowmax = (d_G && cfg->rowmax_ptr) ? reinterpret_cast<const float *>((uintptr_t)cfg->rowmax_ptr) : nullptr;
    A.skip_trivial = (cfg->nn > 0 && cfg->skip_trivial && cfg->positive && cfg->nonneg && !d_sel_out && !d_sel_in) ? 1 : 0;
    A.flag_mod = rt::option(rt::OPT_SOLVE_IMPL) == 3 ? 7 : 0;
    RT_CUDA(cudaMemsetAsync(d_workspace, 0, 1024, st));
    A.diag = nullptr;
    if (d_diag_in) A.diag = d_diag_in;     // pruned fit: the diagonal comes from X (most Gram rows do not exist)
    else {
        gather_diag_kernel<<<(n_items + 255) / 256, 256, 0, st>>>(A, n_items, d_diag);
        A.diag = d_diag;
        RT_CHECK_LAUNCH();
    }
    bool block_pass = true;
    if (cfg->nn > 0 && p.NU <= SW_MAXU && rt::option(rt::OPT_SOLVE_IMPL) != 1) {
        // warp-per-column kernel; columns it cannot take (candidate-list overflow) are flagged for the block kernel
        const size_t tail = (size_t)SW_LIST * 8 > (size_t)p.NU * p.NU * 4 ? (size_t)SW_LIST * 8 : (size_t)p.NU * p.NU * 4;
        const int per_warp = (int)(((sizeof(WarpScratch) + 15) & ~(size_t)15) + ((tail + 15) & ~(size_t)15));
        int warps = (rt::smem_optin() - 1024) / per_warp;
        if (warps > 16) warps = 16;
        if (warps >= 4) {
            int *d_flags = (int *)rt::scratch(SCR_SOLVE_FLAGS, sizeof(int) * ((size_t)n_targets + 64));
            if (!d_flags) return RT_ERR_CUDA;
            RT_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(int) * ((size_t)n_targets + 1), st));
            const size_t smem = (size_t)warps * per_warp;
            RT_CUDA(cudaFuncSetAttribute(slim_solve_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int grid = rt::sm_count();
            const int want = (n_targets + warps - 1) / warps;
            if (grid > want) grid = want;
            slim_solve_warp_kernel<<<grid, warps * 32, smem, st>>>(A, d_flags, per_warp);
            RT_CHECK_LAUNCH();
            // any flagged column?  (the flags are summed on the device; one int comes back)
            int *d_nflag = d_flags + n_targets;
            count_flags_kernel<<<(n_targets + 255) / 256, 256, 0, st>>>(d_flags, n_targets, d_nflag);
            RT_CHECK_LAUNCH();
            int n_flag = 0;
            RT_CUDA(cudaMemcpyAsync(&n_flag, d_nflag, sizeof(int), cudaMemcpyDeviceToHost, st));
            RT_CUDA(cudaStreamSynchronize(st));
            block_pass = n_flag > 0;
            if (block_pass) {
                A.only_flagged = d_flags;
                RT_CUDA(cudaMemsetAsync(d_workspace, 0, 1024, st));  // reset the target cursor
            }
        }
    }
    if (block_pass && cfg->nn == 0 && cfg->positive && cfg->nonneg && rt::option(rt::OPT_SOLVE_IMPL) != 1) {
        // all-features mode: columns without a single live coordinate are finished by the prefilter
        int *d_flags = (int *)rt::scratch(SCR_SOLVE_FLAGS, sizeof(int) * ((size_t)n_targets + 64));
        if (!d_flags) return RT_ERR_CUDA;
        live_prefilter_kernel<<<(unsigned)(((int64_t)n_targets * 32 + 255) / 256), 256, 0, st>>>(A, d_flags);
        RT_CHECK_LAUNCH();
        A.only_flagged = d_flags;
    }
    if (block_pass) {
        if (p.NT == 128) {
            RT_CUDA(cudaFuncSetAttribute(slim_solve_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
            slim_solve_kernel<128><<<p.grid, p.NT, p.smem_bytes, st>>>(A);
        } else {
            RT_CUDA(cudaFuncSetAttribute(slim_solve_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes));
            slim_solve_kernel<512><<<p.grid, p.NT, p.smem_bytes, st>>>(A);
        }
        RT_CHECK_LAUNCH();
    }
    unsigned long long cur[2] = {0, 0};
    RT_CUDA(cudaMemcpyAsync(cur, d_workspace, sizeof(cur), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    const int64_t needed = cfg->nn > 0 ? (int64_t)n_targets * p.NU : (int64_t)cur[0];
    if (h_needed) *h_needed = needed;
    if (needed > out_cap) {
        rt::set_error("rt_slim_solve: output capacity %lld < %lld pairs needed", (long long)out_cap, (long long)needed);
        return RT_ERR_CAPACITY;
    }
    return RT_OK;
}

extern "C" int rt_slim_solve(const float *d_G, int64_t ldg, int32_t n_items, const int32_t *d_targets,
                             int32_t n_targets, const rt_fit_config *cfg, const int32_t *d_sel_in,
                             const uint32_t *d_rng, int64_t rng_len, int32_t *d_sel_out, int64_t *d_out_off,
                             int32_t *d_out_cnt, int32_t *d_out_rows, float *d_out_vals, int64_t out_cap,
                             int64_t *h_needed, int32_t *d_stats, void *stream) {
    RT_ARG(d_G != nullptr || n_targets == 0, "null pointer");
    return slim_solve_impl(d_G, nullptr, 0, nullptr, nullptr, nullptr, ldg, n_items, d_targets, n_targets, cfg, d_sel_in, d_rng, rng_len, d_sel_out,
                           d_out_off, d_out_cnt, d_out_rows, d_out_vals, out_cap, h_needed, d_stats, stream);
}

extern "C" int rt_slim_solve_rows(const void *const *h_bases, int32_t n_bases, const int32_t *d_rowslot, int64_t ldg,
                                  int32_t n_items, const int32_t *d_targets, int32_t n_targets, const rt_fit_config *cfg,
                                  const int32_t *d_sel_in, const uint32_t *d_rng, int64_t rng_len, int32_t *d_sel_out,
                                  int64_t *d_out_off, int32_t *d_out_cnt, int32_t *d_out_rows, float *d_out_vals,
                                  int64_t out_cap, int64_t *h_needed, int32_t *d_stats, void *stream) {
    RT_ARG(h_bases && n_bases >= 1 && n_bases <= RT_MAX_PEERS && d_rowslot && (ldg % 4) == 0, "row buffers");
    for (int q = 0; q < n_bases; ++q) RT_ARG(h_bases[q] != nullptr && (((uintptr_t)h_bases[q]) & 15) == 0, "row buffer pointers");
    return slim_solve_impl(nullptr, h_bases, n_bases, d_rowslot, nullptr, nullptr, ldg, n_items, d_targets, n_targets, cfg, d_sel_in, d_rng, rng_len,
                           d_sel_out, d_out_off, d_out_cnt, d_out_rows, d_out_vals, out_cap, h_needed, d_stats, stream);
}

// ================================================================================================
// Pruned all-features fit: never forms the dense Gram matrix.
//
// With positive coefficients on non-negative data a coordinate c of target j can only leave 0 if G[j][c] > a
// (a = alpha*l1_ratio*n_samples), and by Cauchy-Schwarz G[j][c] <= sqrt(G[j][j] G[c][c]).  So a target can have a non-zero
// solution only if d_j * max_{c != j} d_c > a^2 with d = diag(G) = column sums of squares of X -- one pass over X.  At
// the H&M shape (1.37M users: a = 13,720) that leaves 300 of 105,542 items; every other column is written as the zero
// column at once, and only the Gram rows of the 300 candidates are formed (gram_rows_kernel with a row map) -- 127 MB
// instead of 2 x 44.6 GB, and the solver's prefilter scans 300 rows instead of 105,542.  The candidate set is closed under
// "is a live coordinate of", so the solver never dereferences a missing row; the diagonal for the screening rule comes
// from d.  When the candidates are more than a quarter of the catalogue (small a: ML-1M shape) *h_used = 0 is returned
// and the caller takes the dense path.
namespace rt {

__global__ void col_sumsq_kernel(const int *__restrict__ cptr, const float *__restrict__ cval, int n_items, float *__restrict__ d) {
    const int lane = threadIdx.x & 31;
    const int j = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (j >= n_items) return;
    float s = 0.0f;
    for (int e = cptr[j] + lane; e < cptr[j + 1]; e += 32) { const float v = cval[e]; s = fmaf(v, v, s); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) d[j] = s;
}

// two largest diagonal entries (one block)
__global__ void __launch_bounds__(1024) top2_kernel(const float *__restrict__ d, int n, float *__restrict__ out) {
    __shared__ float s1[32], s2[32];
    float m1 = 0.f, m2 = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = d[i];
        if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) m2 = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float a1 = __shfl_xor_sync(0xffffffffu, m1, o), a2 = __shfl_xor_sync(0xffffffffu, m2, o);
        const float n1 = fmaxf(m1, a1);
        m2 = fmaxf(fminf(m1, a1), fmaxf(m2, a2));
        m1 = n1;
    }
    if ((threadIdx.x & 31) == 0) { s1[threadIdx.x >> 5] = m1; s2[threadIdx.x >> 5] = m2; }
    __syncthreads();
    if (threadIdx.x < 32) {
        m1 = threadIdx.x < (blockDim.x >> 5) ? s1[threadIdx.x] : 0.f;
        m2 = threadIdx.x < (blockDim.x >> 5) ? s2[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float a1 = __shfl_xor_sync(0xffffffffu, m1, o), a2 = __shfl_xor_sync(0xffffffffu, m2, o);
            const float n1 = fmaxf(m1, a1);
            m2 = fmaxf(fminf(m1, a1), fmaxf(m2, a2));
            m1 = n1;
        }
        if (threadIdx.x == 0) { out[0] = m1; out[1] = m2; }
    }
}

// flag[j] = 1 iff item j can have a live coordinate (Cauchy-Schwarz with a 1e-3 safety margin on the fp32 sums)
__global__ void cs_flag_kernel(const float *__restrict__ d, int n, const float *__restrict__ top2, double a2, int *__restrict__ flag) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const float dj = d[j];
    const float other = dj == top2[0] ? top2[1] : top2[0];
    flag[j] = (dj > 0.0f && (double)dj * (double)other >= a2 * (1.0 - 1e-3)) ? 1 : 0;
}

// the solver takes y.y and the coordinate norms from the same numbers: diagonal of the rows that exist
__global__ void diag_from_rows_kernel(const int *__restrict__ slot, int n, const float *__restrict__ rows, int64_t ld, float *__restrict__ d) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n && slot[j] >= 0) d[j] = rows[(size_t)slot[j] * ld + j];
}

__global__ void slot_from_scan_kernel(const int *__restrict__ flag, const int *__restrict__ pos, int n, int *__restrict__ slot) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) slot[j] = flag[j] ? pos[j] : -1;      // single base buffer: (0 << 24) | row
}

}  // namespace rt

#include <cub/cub.cuh>

extern "C" int rt_slim_fit_pruned(int32_t n_users, int32_t n_items, const int32_t *d_cptr, const int32_t *d_cidx,
                                  const float *d_cval, const int32_t *d_ccol, const int32_t *d_rptr, const int32_t *d_ridx,
                                  const float *d_rval, int64_t nnz, const int32_t *d_targets, int32_t n_targets,
                                  const rt_fit_config *cfg, const uint32_t *d_rng, int64_t rng_len, int64_t *d_out_off,
                                  int32_t *d_out_cnt, int32_t *d_out_rows, float *d_out_vals, int64_t out_cap,
                                  int64_t *h_needed, int32_t *d_stats, int32_t *h_used, int32_t *h_n_rows, void *stream) {
    RT_ARG(cfg && h_used && h_n_rows, "cfg / outputs");
    *h_used = 0; *h_n_rows = 0;
    if (h_needed) *h_needed = 0;
    // all features, or feature selection in a bulk fit (skip_trivial: columns without a live coordinate may come back empty)
    if (!((cfg->nn == 0 || cfg->skip_trivial) && cfg->positive && cfg->nonneg) || nnz <= 0 || n_targets <= 0) return RT_OK;
    RT_ARG(n_users > 0 && n_items > 0 && n_items < (1 << 24) && d_cptr && d_cidx && d_cval && d_ccol && d_rptr && d_ridx && d_rval &&
               d_targets, "matrix arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int I = n_items;
    const int64_t ld = ((int64_t)I + 3) / 4 * 4;
    // workspace: diag, top2, flag, pos, slot
    const size_t wbytes = rt::align_up(sizeof(float) * (size_t)I) + 256 + 3 * rt::align_up(sizeof(int) * ((size_t)I + 1));
    char *ws = (char *)rt::scratch(SCR_GRAM_PACK, wbytes);
    if (!ws) return RT_ERR_CUDA;
    float *d_diag = (float *)ws;
    float *d_top2 = (float *)(ws + rt::align_up(sizeof(float) * (size_t)I));
    int *d_flag = (int *)((char *)d_top2 + 256);
    int *d_pos = (int *)((char *)d_flag + rt::align_up(sizeof(int) * ((size_t)I + 1)));
    int *d_slot = (int *)((char *)d_pos + rt::align_up(sizeof(int) * ((size_t)I + 1)));
    rt::col_sumsq_kernel<<<(unsigned)(((int64_t)I * 32 + 255) / 256), 256, 0, st>>>(d_cptr, d_cval, I, d_diag);
    RT_CHECK_LAUNCH();
    rt::top2_kernel<<<1, 1024, 0, st>>>(d_diag, I, d_top2);
    RT_CHECK_LAUNCH();
    const double a = (double)(float)(cfg->alpha * cfg->l1_ratio * (double)cfg->n_samples);
    rt::cs_flag_kernel<<<(I + 255) / 256, 256, 0, st>>>(d_diag, I, d_top2, a * a, d_flag);
    RT_CHECK_LAUNCH();
    {
        size_t tmp_bytes = 0;
        RT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_flag, d_pos, I, st));
        void *tmp = rt::scratch(SCR_CUB, tmp_bytes);
        if (!tmp) return RT_ERR_CUDA;
        RT_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, d_flag, d_pos, I, st));
        rt::count_launch(1);
    }
    int last[2] = {0, 0};
    RT_CUDA(cudaMemcpyAsync(&last[0], d_flag + I - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaMemcpyAsync(&last[1], d_pos + I - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    const int n_sel = last[0] + last[1];
    *h_n_rows = n_sel;
    if ((int64_t)n_sel * 4 > (int64_t)I) return RT_OK;    // too many candidate rows: the dense path is the better one
    rt::slot_from_scan_kernel<<<(I + 255) / 256, 256, 0, st>>>(d_flag, d_pos, I, d_slot);
    RT_CHECK_LAUNCH();
    const size_t gbytes = sizeof(float) * (size_t)(n_sel > 0 ? n_sel : 1) * (size_t)ld;
    float *d_rows = (float *)rt::scratch(SCR_GRAM_SEL, gbytes);
    if (!d_rows) return RT_ERR_CUDA;
    RT_CUDA(cudaMemsetAsync(d_rows, 0, gbytes, st));
    if (n_sel > 0) {
        int rc = rt_gram_rows_selected(d_ccol, d_cidx, d_cval, nnz, d_rptr, d_ridx, d_rval, d_slot, d_rows, ld, st);
        if (rc) return rc;
        rt::diag_from_rows_kernel<<<(I + 255) / 256, 256, 0, st>>>(d_slot, I, d_rows, ld, d_diag);
        RT_CHECK_LAUNCH();
    }
    const void *bases[1] = {d_rows};
    int rc = slim_solve_impl(nullptr, bases, 1, d_slot, d_diag, d_flag, ld, n_items, d_targets, n_targets, cfg, nullptr, d_rng, rng_len,
                             nullptr, d_out_off, d_out_cnt, d_out_rows, d_out_vals, out_cap, h_needed, d_stats, stream);
    if (rc == RT_OK || rc == RT_ERR_CAPACITY) *h_used = 1;
    return rc;
}
