// eval.cu -- ranking metrics of Recommender.evaluate on device-resident top-k lists.
//
// Replaces the per-user Python loop of
//   /root/reference/rtrec/recommender.py:163-200      (evaluate: recommend 100 users at a time, zip with the ground truth)
//   /root/reference/rtrec/utils/metrics.py:6-313       (ndcg, precision, recall, f1, hit, reciprocal rank, AP, AUC, TP)
// One thread per evaluated user: the recommended ids never leave the device as Python lists; what comes back is a
// [n_query, 9] float64 table of per-user values which the host sums sequentially (the reference's `+=` order).
//
// Bit-exactness: every per-user value is formed with the same IEEE-754 double operations in the same order as the
// Python code -- int/int true division = one correctly rounded division, no FMA contraction (explicit __d*_rn), the
// discount table 1/log2(i+2) is computed by the caller with Python's math.log2, and `sum(generator)` over floats is
// restated as CPython does it (>= 3.12: Neumaier compensated summation, `compensated` != 0; before: plain).
#include "common.cuh"

namespace rt {

struct PySum {          // CPython's float path of builtin sum(): Objects/bltinmodule.c
    double f, c;
    int compensated;
    __device__ explicit PySum(int comp) : f(0.0), c(0.0), compensated(comp) {}
    __device__ void add(double x) {
        if (compensated) {
            const double t = __dadd_rn(f, x);
            if (fabs(f) >= fabs(x)) c = __dadd_rn(c, __dadd_rn(__dsub_rn(f, t), x));
            else c = __dadd_rn(c, __dadd_rn(__dsub_rn(x, t), f));
            f = t;
        } else f = __dadd_rn(f, x);
    }
    __device__ double result() const {
        if (compensated && c != 0.0 && isfinite(c)) return __dadd_rn(f, c);
        return f;
    }
};

// out[q*9 + m], m = precision, recall, f1, ndcg, hit_rate, rr, ap, tp, auc (key order of compute_scores' dict)
__global__ void eval_metrics_kernel(const int *__restrict__ ids, const int *__restrict__ cnt, int n_query, int k_stride,
                                    int recommend_size, const int64_t *__restrict__ gptr, const int *__restrict__ gidx,
                                    const double *__restrict__ disc, int compensated, double *__restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_query) return;
    const int len_r = cnt[q];
    const int k = min(len_r, recommend_size);
    const int64_t g0 = gptr[q], g1 = gptr[q + 1];
    const int len_g = (int)(g1 - g0);
    const int *row = ids + (size_t)q * k_stride;
    double *o = out + (size_t)q * 9;

    int tp = 0, first = -1, ordered = 0;
    double ap_sum = 0.0;
    PySum dcg(compensated);
    for (int i = 0; i < k; ++i) {
        const int x = row[i];
        // membership in the sorted ground-truth segment (duplicates allowed)
        int64_t lo = g0, hi = g1;
        while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (gidx[mid] < x) lo = mid + 1; else hi = mid; }
        const bool rel = lo < g1 && gidx[lo] == x;
        if (rel) {
            ++tp;
            if (first < 0) first = i;
            dcg.add(disc[i]);
            ap_sum = __dadd_rn(ap_sum, __ddiv_rn((double)tp, (double)(i + 1)));
        } else ordered += tp;
    }
    double prec, rec, f1, ndcg, hitv, rr, ap, auc;
    if (len_g == 0) {
        // metrics.py: empty ground truth -> 1.0 when nothing was recommended either, else 0.0 (ndcg / hit / rr: 0.0)
        const double e = len_r == 0 ? 1.0 : 0.0;
        prec = rec = ap = auc = e;
        f1 = len_r == 0 ? 1.0 : 0.0;
        ndcg = 0.0; hitv = 0.0; rr = 0.0;
    } else {
        prec = k > 0 ? __ddiv_rn((double)tp, (double)k) : 0.0;
        rec = __ddiv_rn((double)tp, (double)len_g);
        const double s = __dadd_rn(prec, rec);
        f1 = s > 0.0 ? __ddiv_rn(__dmul_rn(2.0, __dmul_rn(prec, rec)), s) : 0.0;
        PySum ideal(compensated);
        const int ik = min(len_g, recommend_size);
        for (int i = 0; i < ik; ++i) ideal.add(disc[i]);
        const double idv = ideal.result();
        ndcg = idv > 0.0 ? __ddiv_rn(dcg.result(), idv) : 0.0;
        hitv = tp > 0 ? 1.0 : 0.0;
        rr = first >= 0 ? __ddiv_rn(1.0, (double)(first + 1)) : 0.0;
        ap = ik > 0 ? __ddiv_rn(ap_sum, (double)ik) : 0.0;
        const int fp = k - tp;
        if (len_r == 0 || tp == 0) auc = 0.0;
        else if (fp == 0) auc = 1.0;
        else auc = __ddiv_rn((double)ordered, (double)((long long)tp * fp));
    }
    o[0] = prec; o[1] = rec; o[2] = f1; o[3] = ndcg; o[4] = hitv; o[5] = rr; o[6] = ap; o[7] = (double)tp; o[8] = auc;
}

}  // namespace rt

extern "C" int rt_eval_metrics(const int32_t *d_ids, const int32_t *d_cnt, int32_t n_query, int32_t k_stride,
                               int32_t recommend_size, const int64_t *d_gptr, const int32_t *d_gidx,
                               const double *d_discount, int32_t compensated_sum, double *d_out, void *stream) {
    RT_ARG(n_query >= 0 && k_stride >= 1 && recommend_size >= 0, "shape");
    if (n_query == 0) return RT_OK;
    RT_ARG(d_ids && d_cnt && d_gptr && d_discount && d_out, "null pointer");
    rt::eval_metrics_kernel<<<(n_query + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        d_ids, d_cnt, n_query, k_stride, recommend_size, d_gptr, d_gidx, d_discount, compensated_sum, d_out);
    RT_CHECK_LAUNCH();
    return RT_OK;
}
