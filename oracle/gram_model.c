/*
 * gram_model.c -- single-threaded CPU MODEL of the device algorithm.  TEST INFRASTRUCTURE ONLY.
 *
 * The CUDA fit path (rtrec_b200/csrc/solve.cu) does not replay sklearn's residual-form
 * coordinate descent literally; it replays the SAME coordinate sequence on the item-item Gram
 * matrix G = X^T X ("Gram-form replay").  This file states that algorithm in
 * plain C so that tests can check, without a GPU, that the reformulation agrees with the exact
 * port in slim_oracle.c (which is pinned to sklearn and to the reference).  It is never the
 * product path and never a fallback.
 *
 * Correspondence with sklearn/linear_model/_cd_fast.pyx:653-1005 for target column j with the
 * feature universe F (all items, or the nn selected items) :
 *     q_c   = x_c . y      = G[c][j]      (0 for c == j: the target column is zeroed)
 *     n2_c  = x_c . x_c    = G[c][c]      (0 for c == j)
 *     h_c   = x_c . (X w)  = sum_k G[c][k] w_k        (maintained incrementally)
 *     tmp   = x_c . R + w_c n2_c = q_c - h_c + w_c n2_c
 *     R.R   = y.y - 2 w.q + w.h ,  R.y = y.y - w.q ,  XtA_c = q_c - h_c - b w_c
 * Live set: with positive=True and non-negative data, tmp <= q_c, so a coordinate with
 * q_c <= a can never leave 0; visits to it only consume an RNG draw.  The solver keeps state
 * for the live coordinates only, but still draws/screens over the whole universe.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define RAND_R_MAX 2147483647u

static inline uint32_t rand_r32(uint32_t *s) {
    if (*s == 0) *s = 1u;
    *s ^= (uint32_t)(*s << 13);
    *s ^= (uint32_t)(*s >> 17);
    *s ^= (uint32_t)(*s << 5);
    return *s % (RAND_R_MAX + 1u);
}

/* G[j][i] = sum_u X[u][j] X[u][i], fp32 accumulate (the device uses fp32 atomics) */
int gm_gram(int n_users, int n_items, const float *csc_data, const int32_t *csc_idx,
            const int32_t *csc_ptr, const float *csr_data, const int32_t *csr_idx,
            const int32_t *csr_ptr, float *G) {
    (void)n_users;
    memset(G, 0, sizeof(float) * (size_t)n_items * (size_t)n_items);
    for (int j = 0; j < n_items; ++j) {
        float *g = G + (size_t)j * n_items;
        for (int32_t p = csc_ptr[j]; p < csc_ptr[j + 1]; ++p) {
            const int u = csc_idx[p];
            const float yu = csc_data[p];
            for (int32_t q = csr_ptr[u]; q < csr_ptr[u + 1]; ++q) g[csr_idx[q]] += csr_data[q] * yu;
        }
    }
    return 0;
}

typedef struct { float s; int32_t i; } cand_t;
static int cand_cmp(const void *pa, const void *pb) {
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    if (a->s > b->s) return -1;
    if (a->s < b->s) return 1;
    return (a->i > b->i) ? -1 : (a->i < b->i ? 1 : 0);
}

/*
 * Solve the targets.  nn > 0: universe = the nn picks (sel_in or built-in tie rule);
 * nn <= 0: universe = all items.  Output layout identical to so_fit_columns (slim_oracle.c).
 * stats[t*4..] = n_iter, draws, gap evaluations, live-set size.
 */
int gm_fit_columns(int n_users, int n_items, const float *G, int n_targets, const int32_t *targets,
                   int nn, const int32_t *sel_in, double alpha, double l1_ratio, int max_iter,
                   double tol, uint32_t seed, int positive, int nonneg, int32_t *sel_out, int out_cap,
                   int32_t *out_rows, float *out_vals, int32_t *out_cnt, int64_t *stats) {
    const double a = (double)(float)(alpha * l1_ratio * n_users);
    const double b = (double)(float)(alpha * (1.0 - l1_ratio) * n_users);
    const double d_w_tol = (double)(float)tol;
    const int NU = nn > 0 ? (nn < n_items ? nn : n_items) : n_items; /* universe size */
    int32_t *feat = (int32_t *)malloc(sizeof(int32_t) * (size_t)(NU + 1));
    cand_t *cand = (cand_t *)malloc(sizeof(cand_t) * (size_t)(n_items + 1));
    int32_t *live_slot = (int32_t *)malloc(sizeof(int32_t) * (size_t)(NU + 1));
    int32_t *live = (int32_t *)malloc(sizeof(int32_t) * (size_t)(NU + 1));
    int32_t *active = (int32_t *)malloc(sizeof(int32_t) * (size_t)(NU + 1));
    uint8_t *excluded = (uint8_t *)malloc((size_t)NU + 1);
    double *w = (double *)malloc(sizeof(double) * (size_t)(NU + 1));   /* per live slot */
    double *h = (double *)malloc(sizeof(double) * (size_t)(NU + 1));   /* per live slot */
    double *xta = (double *)malloc(sizeof(double) * (size_t)(NU + 1)); /* per universe index */
    int err = 0;

    for (int t = 0; t < n_targets; ++t) {
        const int j = targets[t];
        const float *gj = G + (size_t)j * n_items;
        /* ---- universe */
        if (nn > 0) {
            if (sel_in) for (int k = 0; k < NU; ++k) feat[k] = sel_in[(size_t)t * nn + k];
            else {
                for (int i = 0; i < n_items; ++i) { cand[i].s = (i == j) ? 0.0f : gj[i]; cand[i].i = i; }
                qsort(cand, (size_t)n_items, sizeof(cand_t), cand_cmp);
                for (int k = 0; k < NU; ++k) feat[k] = cand[k].i;
            }
        } else for (int k = 0; k < NU; ++k) feat[k] = k;
#define QV(k) ((feat[k] == j) ? 0.0 : (double)gj[feat[k]])
#define N2(k) ((feat[k] == j) ? 0.0 : (double)G[(size_t)feat[k] * n_items + feat[k]])
#define GG(ka, kb) ((feat[ka] == j || feat[kb] == j) ? 0.0 : (double)G[(size_t)feat[ka] * n_items + feat[kb]])
        const double yy = (double)gj[j];
        const double tol_abs = d_w_tol * yy;
        /* ---- live set */
        int m = 0;
        for (int k = 0; k < NU; ++k) {
            int is_live = N2(k) > 0.0 && (!(positive && nonneg) || QV(k) > a);
            live_slot[k] = is_live ? m : -1;
            if (is_live) { live[m] = k; w[m] = 0.0; h[m] = 0.0; ++m; }
        }
        int n_iter = 0, n_gap = 0;
        int64_t draws = 0;
        double gap = 0.0, dual_norm = 0.0;
        uint32_t rs = seed;
        int n_active = 0;
        int converged_at_start = 0;

        /* gap evaluation over the whole universe; fills xta[] */
#define EVAL_GAP()                                                                                  \
        do {                                                                                        \
            double wq = 0, wh = 0, l1 = 0, l2 = 0;                                                  \
            for (int s_ = 0; s_ < m; ++s_) {                                                        \
                wq += w[s_] * QV(live[s_]); wh += w[s_] * h[s_]; l1 += fabs(w[s_]); l2 += w[s_] * w[s_]; \
            }                                                                                       \
            double dn = -INFINITY;                                                                  \
            for (int k = 0; k < NU; ++k) {                                                          \
                double hk, wk;                                                                      \
                if (live_slot[k] >= 0) { hk = h[live_slot[k]]; wk = w[live_slot[k]]; }              \
                else {                                                                              \
                    hk = 0; wk = 0;                                                                 \
                    for (int s_ = 0; s_ < m; ++s_) if (w[s_] != 0.0) hk += GG(k, live[s_]) * w[s_]; \
                }                                                                                   \
                double v = QV(k) - hk - b * wk;                                                     \
                xta[k] = v;                                                                         \
                double av = positive ? v : fabs(v);                                                 \
                if (av > dn) dn = av;                                                               \
            }                                                                                       \
            double Rn = yy - 2.0 * wq + wh, Ry = yy - wq;                                           \
            double primal = 0.5 * (Rn + b * l2) + a * l1;                                           \
            double scale = dn > a ? a / dn : 1.0;                                                   \
            double dualv = -0.5 * scale * scale * (Rn + b * l2) + scale * Ry;                       \
            gap = primal - dualv; dual_norm = dn; ++n_gap;                                          \
        } while (0)

        /* screening; first!=0: consider every feature, else only not-yet-excluded ones */
#define SCREEN(first)                                                                               \
        do {                                                                                        \
            int na = 0;                                                                             \
            for (int k = 0; k < NU; ++k) {                                                          \
                if (first) { if (N2(k) == 0.0) { excluded[k] = 1; continue; } }                     \
                else if (excluded[k]) continue;                                                     \
                double theta = xta[k] / (a > dual_norm ? a : dual_norm);                            \
                double dk = (1.0 - fabs(theta)) / sqrt(N2(k) + b);                                  \
                if (dk <= sqrt(2.0 * gap) / a) { active[na++] = k; excluded[k] = 0; }               \
                else {                                                                              \
                    int s_ = live_slot[k];                                                          \
                    if (s_ >= 0 && w[s_] != 0.0) {                                                  \
                        for (int r_ = 0; r_ < m; ++r_) h[r_] -= w[s_] * GG(live[r_], k);            \
                        w[s_] = 0.0;                                                                \
                    }                                                                               \
                    excluded[k] = 1;                                                                \
                }                                                                                   \
            }                                                                                       \
            n_active = na;                                                                          \
        } while (0)

        EVAL_GAP();
        if (gap <= tol_abs) converged_at_start = 1;
        if (!converged_at_start) {
            SCREEN(1);
            int it;
            for (it = 0; it < max_iter; ++it) {
                double w_max = 0.0, d_w_max = 0.0;
                for (int v = 0; v < n_active; ++v) {
                    int k = active[rand_r32(&rs) % (uint32_t)n_active];
                    ++draws;
                    int s_ = live_slot[k];
                    if (s_ < 0) continue; /* norm2 == 0, or provably stays at 0 */
                    double wc = w[s_], n2 = N2(k);
                    double tmp = QV(k) - h[s_] + wc * n2;
                    double wn;
                    if (positive && tmp < 0.0) wn = 0.0;
                    else {
                        double mag = fabs(tmp) - a;
                        if (!(mag > 0)) mag = 0;
                        wn = (tmp > 0 ? 1.0 : (tmp < 0 ? -1.0 : 0.0)) * mag / (n2 + b);
                    }
                    if (wn != wc) {
                        double d = wn - wc;
                        for (int r_ = 0; r_ < m; ++r_) h[r_] += d * GG(live[r_], k);
                        w[s_] = wn;
                    }
                    double dw = fabs(wn - wc);
                    if (dw > d_w_max) d_w_max = dw;
                    if (fabs(wn) > w_max) w_max = fabs(wn);
                }
                if (w_max == 0.0 || d_w_max / w_max <= d_w_tol || it == max_iter - 1) {
                    EVAL_GAP();
                    if (gap <= tol_abs) break;
                    SCREEN(0);
                }
            }
            n_iter = it < max_iter ? it + 1 : max_iter;
        }
        stats[(size_t)t * 4 + 0] = n_iter; stats[(size_t)t * 4 + 1] = draws;
        stats[(size_t)t * 4 + 2] = n_gap; stats[(size_t)t * 4 + 3] = m;
        /* ---- output */
        int cnt = 0;
        int32_t *orow = out_rows + (size_t)t * out_cap;
        float *oval = out_vals + (size_t)t * out_cap;
        if (nn > 0) {
            for (int k = 0; k < NU; ++k) {
                if (sel_out) sel_out[(size_t)t * nn + k] = feat[k];
                float v = live_slot[k] >= 0 ? (float)w[live_slot[k]] : 0.0f;
                if (cnt < out_cap) { orow[cnt] = feat[k]; oval[cnt] = v; ++cnt; }
            }
            if (sel_out) for (int k = NU; k < nn; ++k) sel_out[(size_t)t * nn + k] = -1;
        } else {
            for (int k = 0; k < NU; ++k) {
                float v = live_slot[k] >= 0 ? (float)w[live_slot[k]] : 0.0f;
                if (v != 0.0f) { if (cnt < out_cap) { orow[cnt] = k; oval[cnt] = v; } ++cnt; }
            }
            if (cnt > out_cap) { err = 2; cnt = out_cap; }
        }
        out_cnt[t] = cnt;
#undef QV
#undef N2
#undef GG
#undef EVAL_GAP
#undef SCREEN
    }
    free(feat); free(cand); free(live_slot); free(live); free(active); free(excluded);
    free(w); free(h); free(xta);
    return err;
}
