// store.cu -- K1/K2: device-resident interaction store.
//
// Replaces /root/reference/rtrec/utils/interactions.py:
//   add_interaction  :81-119   (per-event python dict update: decay-read, accumulate, clip / upsert)
//   _apply_decay     :62-79
//   to_csr / to_csc  :259-303  (python per-nnz walk -> scipy COO -> CSR/CSC float32)
//
// Store = arrays sorted by key (user << 32 | item): keys u64, vals f64, stamps f64.
// A batch of events is folded in arrival order:
//   T_k   = running max of (ts + 1) including the carried-in max_timestamp   (interactions.py:99)
//   stable sort of the batch by key, one thread per distinct key folds its events sequentially
//   against the pair's previous state, then the new pairs are merged into the sorted store.
// CUB supplies sort / scan / run-length primitives; fold, merge and the matrix builders are
// hand-written.
#include <cub/cub.cuh>

#include <limits.h>
#include <string.h>

#include "common.cuh"

namespace rt {

struct MaxOpD {
    __device__ __forceinline__ double operator()(const double &a, const double &b) const { return a > b ? a : b; }
};

__global__ void prep_events_kernel(const int *users, const int *items, const double *ts, int64_t n, double max_ts_in,
                                   unsigned long long *keys, unsigned *order, double *tplus, int *max_ui) {
    // grid-stride: the id maxima are reduced per thread, per warp and per block, so the whole launch issues two
    // atomics per block instead of two per warp on the same address (that serialisation was 1 ms at 20M events)
    __shared__ int s_mu[32], s_mi[32];
    int mu = 0, mi = 0;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const int u = users[k], i = items[k];
        keys[k] = (((unsigned long long)(unsigned)u) << 32) | (unsigned)i;
        order[k] = (unsigned)k;
        double t = ts[k] + 1.0;
        if (k == 0 && max_ts_in > t) t = max_ts_in;
        tplus[k] = t;
        mu = max(mu, u); mi = max(mi, i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mu = max(mu, __shfl_xor_sync(0xffffffffu, mu, o));
        mi = max(mi, __shfl_xor_sync(0xffffffffu, mi, o));
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) { s_mu[warp] = mu; s_mi[warp] = mi; }
    __syncthreads();
    if (warp == 0) {
        mu = lane < nw ? s_mu[lane] : 0; mi = lane < nw ? s_mi[lane] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mu = max(mu, __shfl_xor_sync(0xffffffffu, mu, o));
            mi = max(mi, __shfl_xor_sync(0xffffffffu, mi, o));
        }
        if (lane == 0) { atomicMax(&max_ui[0], mu); atomicMax(&max_ui[1], mi); }
    }
}

__device__ __forceinline__ int64_t lower_bound_u64(const unsigned long long *a, int64_t n, unsigned long long v) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; }
    return lo;
}

// one thread per distinct key of the batch
__global__ void fold_kernel(const unsigned long long *ukeys, const int *run_start, const int *run_len, int n_runs,
                            const unsigned *sorted_order, const double *ts, const double *delta, const double *T,
                            int upsert, double vmin, double vmax, double rate, int use_decay,
                            const unsigned long long *keys0, const double *vals0, const double *stamps0, int64_t n0,
                            double *new_val, double *new_stamp, int64_t *old_pos, unsigned char *is_new) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_runs) return;
    const unsigned long long key = ukeys[g];
    const int64_t p = lower_bound_u64(keys0, n0, key);
    const bool found = p < n0 && keys0[p] == key;
    double v = found ? vals0[p] : 0.0, s = found ? stamps0[p] : 0.0;
    const int a = run_start[g], len = run_len[g];
    for (int e = 0; e < len; ++e) {
        const unsigned k = sorted_order[a + e];
        if (upsert) v = delta[k];
        else {
            double cur = 0.0;
            if (v != 0.0) cur = use_decay ? v * pow(rate, (T[k] - s) / 86400.0) : v;
            double nv = cur + delta[k];
            nv = fmax(vmin, fmin(nv, vmax));
            v = nv;
        }
        s = ts[k];
    }
    new_val[g] = v; new_stamp[g] = s;
    old_pos[g] = found ? p : -1;
    is_new[g] = found ? 0 : 1;
}

// compacted list of brand-new keys (ascending) -> their rank; merged output positions
__global__ void merge_old_kernel(const unsigned long long *keys0, const double *vals0, const double *stamps0, int64_t n0,
                                 const unsigned long long *nkeys, int64_t n_new, unsigned long long *okeys,
                                 double *ovals, double *ostamps) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n0) return;
    const unsigned long long key = keys0[p];
    const int64_t pos = p + lower_bound_u64(nkeys, n_new, key);
    okeys[pos] = key; ovals[pos] = vals0[p]; ostamps[pos] = stamps0[p];
}

__global__ void merge_new_kernel(const unsigned long long *ukeys, const double *new_val, const double *new_stamp,
                                 const int64_t *old_pos, const int *new_rank, int n_runs,
                                 const unsigned long long *keys0, int64_t n0, const unsigned long long *nkeys,
                                 int64_t n_new, unsigned long long *okeys, double *ovals, double *ostamps) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_runs) return;
    const unsigned long long key = ukeys[g];
    int64_t pos;
    if (old_pos[g] >= 0) pos = old_pos[g] + lower_bound_u64(nkeys, n_new, key);  // overwrite the old pair
    else pos = (int64_t)new_rank[g] + lower_bound_u64(keys0, n0, key);
    okeys[pos] = key; ovals[pos] = new_val[g]; ostamps[pos] = new_stamp[g];
}

__global__ void gather_new_keys_kernel(const unsigned long long *ukeys, const unsigned char *is_new, const int *new_rank,
                                       int n_runs, unsigned long long *nkeys) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < n_runs && is_new[g]) nkeys[new_rank[g]] = ukeys[g];
}

__global__ void u8_to_i32_kernel(const unsigned char *in, int n, int *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

// ---- matrix build ------------------------------------------------------------------------------
__global__ void decay_values_kernel(const unsigned long long *keys, const double *vals, const double *stamps, int64_t n,
                                    double rate, int use_decay, double max_ts, const unsigned char *item_mask,
                                    float *x, unsigned char *keep, int *any_negative) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double v = vals[p];
    const double d = use_decay ? v * pow(rate, (max_ts - stamps[p]) / 86400.0) : v;
    const float f = (float)d;
    x[p] = f;
    const unsigned item = (unsigned)(keys[p] & 0xffffffffull);
    const bool k = item_mask ? item_mask[item] != 0 : true;
    if (keep) keep[p] = k ? 1 : 0;
    if (k && f < 0.0f) *any_negative = 1;
}

__global__ void split_keys_kernel(const unsigned long long *keys, int64_t n, int *hi, int *lo) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    if (hi) hi[p] = (int)(keys[p] >> 32);
    if (lo) lo[p] = (int)(keys[p] & 0xffffffffull);
}

__global__ void swap_halves_kernel(const unsigned long long *keys, int64_t n, unsigned long long *out) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) out[p] = (keys[p] << 32) | (keys[p] >> 32);
}

// ptr[r] = first position whose high half is >= r, for r in [0, n_major]
__global__ void ptr_from_keys_kernel(const unsigned long long *keys, int64_t n, int n_major, int *ptr) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_major) return;
    ptr[r] = (int)lower_bound_u64(keys, n, ((unsigned long long)(unsigned)r) << 32);
}

__global__ void lookup_kernel(const unsigned long long *keys, const double *vals, const double *stamps, int64_t n,
                              const unsigned long long *query, int64_t nq, double *val_out, double *stamp_out,
                              unsigned char *found) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const int64_t p = lower_bound_u64(keys, n, query[q]);
    const bool f = p < n && keys[p] == query[q];
    found[q] = f ? 1 : 0;
    val_out[q] = f ? vals[p] : 0.0;
    stamp_out[q] = f ? stamps[p] : 0.0;
}

// ---- ingest bookkeeping (interactions.py:92-99, 113-119 for a whole batch) ----------------------------
// out[0..5] = min user, max user, min item, max item (int64), out_ts[0] = max timestamp
__global__ void events_minmax_kernel(const long long *__restrict__ users, const long long *__restrict__ items,
                                     const double *__restrict__ ts, int64_t n, long long *__restrict__ out,
                                     double *__restrict__ out_ts) {
    long long umin = LLONG_MAX, umax = LLONG_MIN, imin = LLONG_MAX, imax = LLONG_MIN;
    double tmax = -INFINITY;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const long long u = users[k], i = items[k];
        umin = min(umin, u); umax = max(umax, u); imin = min(imin, i); imax = max(imax, i);
        tmax = fmax(tmax, ts[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        umin = min(umin, __shfl_xor_sync(0xffffffffu, umin, o)); umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        imin = min(imin, __shfl_xor_sync(0xffffffffu, imin, o)); imax = max(imax, __shfl_xor_sync(0xffffffffu, imax, o));
        tmax = fmax(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out[0], umin); atomicMax(&out[1], umax); atomicMin(&out[2], imin); atomicMax(&out[3], imax);
        // timestamps are non-negative in practice; order-preserving integer max works for any sign
        long long b = __double_as_longlong(tmax);
        b = b < 0 ? (b ^ 0x7fffffffffffffffll) : b;
        atomicMax((long long *)out_ts, b);
    }
}

// per item: number of events with delta > 0, arrival index of the last such event, seen flag
template <typename IdT>
__global__ void events_item_stats_kernel(const IdT *__restrict__ items, const double *__restrict__ delta, int64_t n,
                                         int *__restrict__ count_pos, int *__restrict__ last_pos,
                                         unsigned char *__restrict__ seen) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const long long i = (long long)items[k];
    seen[i] = 1;
    if (delta[k] > 0.0) {
        atomicAdd(&count_pos[i], 1);
        atomicMax(&last_pos[i], (int)k);
    }
}

static int bits_for64(long long n) { int b = 1; while ((1ll << b) < n && b < 32) ++b; return b; }

}  // namespace rt

using namespace rt;

#define CUB_CALL(call_expr)                                                                         \
    do {                                                                                            \
        size_t tmp_bytes__ = 0;                                                                     \
        void *d_tmp__ = nullptr;                                                                    \
        RT_CUDA(call_expr);                                                                         \
        d_tmp__ = rt::scratch(SCR_CUB, tmp_bytes__);                                                \
        if (!d_tmp__) return RT_ERR_CUDA;                                                           \
        RT_CUDA(call_expr);                                                                         \
        rt::count_launch(2);                                                                        \
    } while (0)

extern "C" int rt_store_fold(const int32_t *d_users, const int32_t *d_items, const double *d_ts,
                             const double *d_delta, int64_t n_events, int upsert, double min_value,
                             double max_value, double decay_rate, const uint64_t *d_keys, const double *d_vals,
                             const double *d_stamps, int64_t n_pairs, double max_ts_in, int32_t max_user_in,
                             int32_t max_item_in, uint64_t *d_out_keys, double *d_out_vals, double *d_out_stamps,
                             int64_t out_cap, int64_t *h_n_out, double *h_max_ts, int32_t *h_max_user,
                             int32_t *h_max_item, void *stream) {
    RT_ARG(n_events >= 0 && n_pairs >= 0, "sizes");
    RT_ARG(n_events < (1ll << 31) && n_pairs < (1ll << 31), "more than 2^31 events/pairs per call is not supported");
    RT_ARG(h_n_out && h_max_ts && h_max_user && h_max_item, "host outputs");
    cudaStream_t st = (cudaStream_t)stream;
    const int use_decay = (decay_rate > 0.0 && decay_rate == decay_rate) ? 1 : 0;
    if (n_events == 0) {
        RT_ARG(out_cap >= n_pairs, "out_cap");
        if (n_pairs) {
            RT_CUDA(cudaMemcpyAsync(d_out_keys, d_keys, sizeof(uint64_t) * n_pairs, cudaMemcpyDeviceToDevice, st));
            RT_CUDA(cudaMemcpyAsync(d_out_vals, d_vals, sizeof(double) * n_pairs, cudaMemcpyDeviceToDevice, st));
            RT_CUDA(cudaMemcpyAsync(d_out_stamps, d_stamps, sizeof(double) * n_pairs, cudaMemcpyDeviceToDevice, st));
        }
        *h_n_out = n_pairs; *h_max_ts = max_ts_in; *h_max_user = max_user_in; *h_max_item = max_item_in;
        RT_CUDA(cudaStreamSynchronize(st));
        return RT_OK;
    }
    RT_ARG(d_users && d_items && d_ts && d_delta, "event arrays");
    RT_ARG(out_cap >= n_pairs + n_events, "out_cap must be >= n_pairs + n_events");
    const int64_t n = n_events;
    // scratch layout
    auto plan = [&](Carver &c) {
        struct P { unsigned long long *keys, *skeys, *ukeys, *nkeys; unsigned *order, *sorder; double *tplus, *T, *nval, *nstamp;
                   int *run_len, *run_start, *n_runs, *isnew_i, *new_rank, *max_ui; int64_t *old_pos; unsigned char *is_new; } p;
        p.keys = c.take<unsigned long long>(n); p.skeys = c.take<unsigned long long>(n);
        p.ukeys = c.take<unsigned long long>(n); p.nkeys = c.take<unsigned long long>(n);
        p.order = c.take<unsigned>(n); p.sorder = c.take<unsigned>(n);
        p.tplus = c.take<double>(n); p.T = c.take<double>(n); p.nval = c.take<double>(n); p.nstamp = c.take<double>(n);
        p.run_len = c.take<int>(n + 1); p.run_start = c.take<int>(n + 1); p.n_runs = c.take<int>(4);
        p.isnew_i = c.take<int>(n + 1); p.new_rank = c.take<int>(n + 1); p.max_ui = c.take<int>(4);
        p.old_pos = c.take<int64_t>(n); p.is_new = c.take<unsigned char>(n);
        return p;
    };
    Carver sizing(nullptr, (size_t)-1);
    plan(sizing);
    void *base = rt::scratch(SCR_STORE_A, sizing.off + 1024);
    if (!base) return RT_ERR_CUDA;
    Carver real(base, sizing.off + 1024);
    auto P = plan(real);

    const int bs = 256;
    const unsigned gn = (unsigned)((n + bs - 1) / bs);
    int init_max[4] = {max_user_in, max_item_in, 0, 0};
    RT_CUDA(cudaMemcpyAsync(P.max_ui, init_max, sizeof(init_max), cudaMemcpyHostToDevice, st));
    {
        unsigned gp = (unsigned)rt::sm_count() * 8u;
        if (gp > gn) gp = gn;
        prep_events_kernel<<<gp, bs, 0, st>>>(d_users, d_items, d_ts, n, max_ts_in, (unsigned long long *)P.keys, P.order,
                                             P.tplus, P.max_ui);
    }
    RT_CHECK_LAUNCH();
    // T_k = inclusive running max of (ts+1)
    CUB_CALL(cub::DeviceScan::InclusiveScan(d_tmp__, tmp_bytes__, P.tplus, P.T, MaxOpD(), (int)n, st));
    // stable sort by key
    int host_max[4];
    RT_CUDA(cudaMemcpyAsync(host_max, P.max_ui, sizeof(host_max), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    // stable LSD sort over the two populated bit ranges only (item bits, then user bits): bits between the
    // largest item id and bit 32 are zero and would cost whole passes.  ukeys / isnew_i+new_rank serve as the
    // intermediate buffers (they are written later).
    {
        const int item_bits = bits_for64((long long)host_max[1] + 1), user_bits = bits_for64((long long)host_max[0] + 1);
        unsigned long long *mid_k = P.ukeys;
        unsigned *mid_o = (unsigned *)P.old_pos;  // int64[n] scratch, large enough for n uint32
        CUB_CALL(cub::DeviceRadixSort::SortPairs(d_tmp__, tmp_bytes__, P.keys, mid_k, P.order, mid_o, (int)n, 0, item_bits, st));
        CUB_CALL(cub::DeviceRadixSort::SortPairs(d_tmp__, tmp_bytes__, mid_k, P.skeys, mid_o, P.sorder, (int)n, 32, 32 + user_bits, st));
    }
    CUB_CALL(cub::DeviceRunLengthEncode::Encode(d_tmp__, tmp_bytes__, P.skeys, P.ukeys, P.run_len, P.n_runs, (int)n, st));
    int n_runs = 0;
    RT_CUDA(cudaMemcpyAsync(&n_runs, P.n_runs, sizeof(int), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp__, tmp_bytes__, P.run_len, P.run_start, n_runs, st));
    const unsigned gr = (unsigned)((n_runs + bs - 1) / bs);
    fold_kernel<<<gr, bs, 0, st>>>(P.ukeys, P.run_start, P.run_len, n_runs, P.sorder, d_ts, d_delta, P.T, upsert, min_value,
                                  max_value, decay_rate, use_decay, (const unsigned long long *)d_keys, d_vals, d_stamps,
                                  n_pairs, P.nval, P.nstamp, P.old_pos, P.is_new);
    RT_CHECK_LAUNCH();
    u8_to_i32_kernel<<<gr, bs, 0, st>>>(P.is_new, n_runs, P.isnew_i);
    RT_CHECK_LAUNCH();
    RT_CUDA(cudaMemsetAsync(P.isnew_i + n_runs, 0, sizeof(int), st));
    CUB_CALL(cub::DeviceScan::ExclusiveSum(d_tmp__, tmp_bytes__, P.isnew_i, P.new_rank, n_runs + 1, st));
    int n_new = 0;
    RT_CUDA(cudaMemcpyAsync(&n_new, P.new_rank + n_runs, sizeof(int), cudaMemcpyDeviceToHost, st));
    double last_T = 0.0;
    RT_CUDA(cudaMemcpyAsync(&last_T, P.T + (n - 1), sizeof(double), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    gather_new_keys_kernel<<<gr, bs, 0, st>>>(P.ukeys, P.is_new, P.new_rank, n_runs, P.nkeys);
    RT_CHECK_LAUNCH();
    if (n_pairs > 0) {
        merge_old_kernel<<<(unsigned)((n_pairs + bs - 1) / bs), bs, 0, st>>>((const unsigned long long *)d_keys, d_vals, d_stamps,
                                                                            n_pairs, P.nkeys, n_new,
                                                                            (unsigned long long *)d_out_keys, d_out_vals,
                                                                            d_out_stamps);
        RT_CHECK_LAUNCH();
    }
    merge_new_kernel<<<gr, bs, 0, st>>>(P.ukeys, P.nval, P.nstamp, P.old_pos, P.new_rank, n_runs,
                                       (const unsigned long long *)d_keys, n_pairs, P.nkeys, n_new,
                                       (unsigned long long *)d_out_keys, d_out_vals, d_out_stamps);
    RT_CHECK_LAUNCH();
    RT_CUDA(cudaStreamSynchronize(st));
    *h_n_out = n_pairs + n_new;
    *h_max_ts = last_T;
    *h_max_user = host_max[0];
    *h_max_item = host_max[1];
    return RT_OK;
}

extern "C" int rt_store_build(const uint64_t *d_keys, const double *d_vals, const double *d_stamps,
                              int64_t n_pairs, double decay_rate, double max_ts, int32_t n_users, int32_t n_items,
                              const uint8_t *d_item_mask, int32_t *d_rptr, int32_t *d_ridx, float *d_rval,
                              int32_t *d_cptr, int32_t *d_cidx, float *d_cval, int32_t *d_ccol, int64_t *h_nnz,
                              int *h_nonneg, void *stream) {
    RT_ARG(n_pairs >= 0 && n_pairs < (1ll << 31), "n_pairs");
    RT_ARG(n_users > 0 && n_items > 0, "shape");
    RT_ARG(d_rptr && d_cptr && h_nnz && h_nonneg, "pointers");
    cudaStream_t st = (cudaStream_t)stream;
    const int use_decay = (decay_rate > 0.0 && decay_rate == decay_rate) ? 1 : 0;
    const int bs = 256;
    if (n_pairs == 0) {
        RT_CUDA(cudaMemsetAsync(d_rptr, 0, sizeof(int) * ((size_t)n_users + 1), st));
        RT_CUDA(cudaMemsetAsync(d_cptr, 0, sizeof(int) * ((size_t)n_items + 1), st));
        *h_nnz = 0; *h_nonneg = 1;
        RT_CUDA(cudaStreamSynchronize(st));
        return RT_OK;
    }
    RT_ARG(d_keys && d_vals && d_stamps && d_ridx && d_rval && d_cidx && d_cval, "pointers");
    const int64_t n = n_pairs;
    Carver sizing(nullptr, (size_t)-1);
    auto plan = [&](Carver &c) {
        struct P { float *x, *xk; unsigned char *keep; unsigned long long *kk, *sw, *sws; int *flag, *n_sel; } p;
        p.x = c.take<float>(n); p.xk = c.take<float>(n); p.keep = c.take<unsigned char>(n);
        p.kk = c.take<unsigned long long>(n); p.sw = c.take<unsigned long long>(n); p.sws = c.take<unsigned long long>(n);
        p.flag = c.take<int>(4); p.n_sel = c.take<int>(4);
        return p;
    };
    plan(sizing);
    void *base = rt::scratch(SCR_STORE_B, sizing.off + 1024);
    if (!base) return RT_ERR_CUDA;
    Carver real(base, sizing.off + 1024);
    auto P = plan(real);
    RT_CUDA(cudaMemsetAsync(P.flag, 0, sizeof(int) * 4, st));
    const unsigned gn = (unsigned)((n + bs - 1) / bs);
    decay_values_kernel<<<gn, bs, 0, st>>>((const unsigned long long *)d_keys, d_vals, d_stamps, n, decay_rate, use_decay,
                                          max_ts, d_item_mask, P.x, d_item_mask ? P.keep : nullptr, P.flag);
    RT_CHECK_LAUNCH();
    const unsigned long long *keys = (const unsigned long long *)d_keys;
    const float *x = P.x;
    int64_t nnz = n;
    if (d_item_mask) {
        CUB_CALL(cub::DeviceSelect::Flagged(d_tmp__, tmp_bytes__, (const unsigned long long *)d_keys, P.keep, P.kk, P.n_sel, (int)n, st));
        CUB_CALL(cub::DeviceSelect::Flagged(d_tmp__, tmp_bytes__, P.x, P.keep, P.xk, P.n_sel, (int)n, st));
        int n_sel = 0;
        RT_CUDA(cudaMemcpyAsync(&n_sel, P.n_sel, sizeof(int), cudaMemcpyDeviceToHost, st));
        RT_CUDA(cudaStreamSynchronize(st));
        nnz = n_sel; keys = P.kk; x = P.xk;
    }
    *h_nnz = nnz;
    // CSR: keys are already (user, item)-sorted
    ptr_from_keys_kernel<<<(n_users + 1 + bs - 1) / bs, bs, 0, st>>>(keys, nnz, n_users, d_rptr);
    RT_CHECK_LAUNCH();
    if (nnz > 0) {
        const unsigned gz = (unsigned)((nnz + bs - 1) / bs);
        split_keys_kernel<<<gz, bs, 0, st>>>(keys, nnz, nullptr, d_ridx);
        RT_CHECK_LAUNCH();
        RT_CUDA(cudaMemcpyAsync(d_rval, x, sizeof(float) * nnz, cudaMemcpyDeviceToDevice, st));
        // CSC: sort by (item, user)
        swap_halves_kernel<<<gz, bs, 0, st>>>(keys, nnz, P.sw);
        RT_CHECK_LAUNCH();
        // the entries arrive (user, item)-sorted, so a STABLE sort on the item bits alone yields (item, user) order
        const int end_bit = 32 + bits_for64(n_items);
        CUB_CALL(cub::DeviceRadixSort::SortPairs(d_tmp__, tmp_bytes__, P.sw, P.sws, x, d_cval, (int)nnz, 32, end_bit, st));
        split_keys_kernel<<<gz, bs, 0, st>>>(P.sws, nnz, d_ccol, d_cidx);
        RT_CHECK_LAUNCH();
    }
    ptr_from_keys_kernel<<<(n_items + 1 + bs - 1) / bs, bs, 0, st>>>(nnz > 0 ? P.sws : keys, nnz, n_items, d_cptr);
    RT_CHECK_LAUNCH();
    int flag = 0;
    RT_CUDA(cudaMemcpyAsync(&flag, P.flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    *h_nonneg = flag ? 0 : 1;
    return RT_OK;
}

extern "C" int rt_store_lookup(const uint64_t *d_keys, const double *d_vals, const double *d_stamps, int64_t n_pairs,
                               const uint64_t *d_query, int64_t n, double *d_val_out, double *d_stamp_out,
                               uint8_t *d_found, void *stream) {
    if (n <= 0) return RT_OK;
    RT_ARG(d_query && d_val_out && d_stamp_out && d_found, "pointers");
    const int bs = 256;
    lookup_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(
        (const unsigned long long *)d_keys, d_vals, d_stamps, n_pairs, (const unsigned long long *)d_query, n, d_val_out,
        d_stamp_out, d_found);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_events_minmax(const int64_t *d_users, const int64_t *d_items, const double *d_ts, int64_t n,
                                int64_t *h_min_user, int64_t *h_max_user, int64_t *h_min_item, int64_t *h_max_item,
                                double *h_max_ts, void *stream) {
    RT_ARG(n > 0 && d_users && d_items && d_ts, "event arrays");
    RT_ARG(h_min_user && h_max_user && h_min_item && h_max_item && h_max_ts, "host outputs");
    cudaStream_t st = (cudaStream_t)stream;
    long long *d_out = (long long *)rt::scratch(SCR_MISC, 256);
    if (!d_out) return RT_ERR_CUDA;
    long long init[5] = {LLONG_MAX, LLONG_MIN, LLONG_MAX, LLONG_MIN, LLONG_MIN};
    RT_CUDA(cudaMemcpyAsync(d_out, init, sizeof(init), cudaMemcpyHostToDevice, st));
    int grid = rt::sm_count() * 8;
    const int64_t want = (n + 255) / 256;
    if (grid > want) grid = (int)want;
    events_minmax_kernel<<<grid, 256, 0, st>>>((const long long *)d_users, (const long long *)d_items, d_ts, n, d_out,
                                              (double *)(d_out + 4));
    RT_CHECK_LAUNCH();
    long long host[5];
    RT_CUDA(cudaMemcpyAsync(host, d_out, sizeof(host), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    *h_min_user = host[0]; *h_max_user = host[1]; *h_min_item = host[2]; *h_max_item = host[3];
    long long b = host[4];
    b = b < 0 ? (b ^ 0x7fffffffffffffffll) : b;
    double t;
    memcpy(&t, &b, sizeof(t));
    *h_max_ts = t;
    return RT_OK;
}

static int events_item_stats_impl(const void *d_items, int wide, const double *d_delta, int64_t n, int32_t n_items,
                                  int32_t *d_count_pos, int32_t *d_last_pos, uint8_t *d_seen, void *stream) {
    RT_ARG(n >= 0 && n < (1ll << 31) && n_items > 0, "sizes");
    RT_ARG(d_count_pos && d_last_pos && d_seen, "outputs");
    cudaStream_t st = (cudaStream_t)stream;
    RT_CUDA(cudaMemsetAsync(d_count_pos, 0, sizeof(int) * (size_t)n_items, st));
    RT_CUDA(cudaMemsetAsync(d_last_pos, 0xff, sizeof(int) * (size_t)n_items, st));  // -1
    RT_CUDA(cudaMemsetAsync(d_seen, 0, (size_t)n_items, st));
    if (n == 0) return RT_OK;
    RT_ARG(d_items && d_delta, "event arrays");
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (wide) events_item_stats_kernel<long long><<<grid, 256, 0, st>>>((const long long *)d_items, d_delta, n, d_count_pos, d_last_pos, d_seen);
    else events_item_stats_kernel<int><<<grid, 256, 0, st>>>((const int *)d_items, d_delta, n, d_count_pos, d_last_pos, d_seen);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_events_item_stats(const int64_t *d_items, const double *d_delta, int64_t n, int32_t n_items,
                                    int32_t *d_count_pos, int32_t *d_last_pos, uint8_t *d_seen, void *stream) {
    return events_item_stats_impl(d_items, 1, d_delta, n, n_items, d_count_pos, d_last_pos, d_seen, stream);
}

extern "C" int rt_events_item_stats32(const int32_t *d_items, const double *d_delta, int64_t n, int32_t n_items,
                                      int32_t *d_count_pos, int32_t *d_last_pos, uint8_t *d_seen, void *stream) {
    return events_item_stats_impl(d_items, 0, d_delta, n, n_items, d_count_pos, d_last_pos, d_seen, stream);
}
