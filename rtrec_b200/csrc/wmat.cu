// wmat.cu -- K5: assemble / merge the item-similarity matrix W on the device, and CSC<->CSR.
//
// Replaces the per-entry LIL writes `item_similarity[i, j] = value` and the tolil()/resize()/
// tocsc() round trip of /root/reference/rtrec/models/internal/slim_elastic.py:252,273-274,280
// (fit), :322-327,371-374,385 (fit_in_parallel) and :533-538,556-557,563 (partial_fit_items).
// LIL assignment semantics: a non-zero value inserts/overwrites, a zero value deletes, rows that
// the solver did not return keep their old ("stale") value.
//
// Method: one warp per column counts, then emits (col<<32|row, value) pairs unsorted inside the
// column's segment; one CUB radix sort over all pairs puts rows in ascending order.  CUB
// (sort/scan) is used as a library primitive; the count/emit kernels are the hand-written part.
#include <cub/cub.cuh>

#include "common.cuh"

namespace rt {

__global__ void fill_i32_kernel(int *p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void scatter_tmap_kernel(const int *targets, int n_targets, int n_items, int *tmap) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_targets) { const int j = targets[t]; if (j >= 0 && j < n_items) tmap[j] = t; }
}

// is `row` among the rows returned for this target?  rows_sorted selects binary search.
__device__ __forceinline__ bool returned_row(const int *rows, int c, int row, bool rows_sorted) {
    if (rows_sorted) {
        int lo = 0, hi = c;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (rows[mid] < row) lo = mid + 1; else hi = mid; }
        return lo < c && rows[lo] == row;
    }
    for (int e = 0; e < c; ++e) if (rows[e] == row) return true;
    return false;
}

// pass 0: count entries of every output column; pass 1: emit them at wptr[j]
template <int PASS>
__global__ void w_merge_kernel(int n_items, const int *old_ptr, const int *old_idx, const float *old_val,
                               int n_old_items, const int *tmap, const int64_t *off, const int *cnt,
                               const int *rows, const float *vals, int rows_sorted, int *counts,
                               const int *wptr, unsigned long long *keys, float *oval) {
    const int lane = threadIdx.x & 31;
    const int j = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (j >= n_items) return;
    const int o0 = j < n_old_items ? old_ptr[j] : 0, o1 = j < n_old_items ? old_ptr[j + 1] : 0;
    const int t = tmap[j];
    int total = 0;
    const int base = PASS ? wptr[j] : 0;
    const unsigned long long colkey = ((unsigned long long)(unsigned)j) << 32;
    if (t < 0) {
        if (PASS) for (int p = o0 + lane; p < o1; p += 32) { keys[base + p - o0] = colkey | (unsigned)old_idx[p]; oval[base + p - o0] = old_val[p]; }
        total = o1 - o0;
    } else {
        const int c = cnt[t];
        const int *r = rows + off[t];
        const float *v = vals + off[t];
        // stale entries: old rows the solver did not return (and that still fit the matrix)
        for (int p0 = o0; p0 < o1; p0 += 32) {
            const int p = p0 + lane;
            bool keep = false;
            int row = 0;
            if (p < o1) { row = old_idx[p]; keep = row < n_items && !returned_row(r, c, row, rows_sorted != 0); }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (PASS && keep) {
                const int pos = base + total + __popc(bal & ((1u << lane) - 1u));
                keys[pos] = colkey | (unsigned)row; oval[pos] = old_val[p];
            }
            total += __popc(bal);
        }
        // returned non-zeros
        for (int e0 = 0; e0 < c; e0 += 32) {
            const int e = e0 + lane;
            const bool nz = e < c && v[e] != 0.0f && r[e] >= 0 && r[e] < n_items;
            const unsigned bal = __ballot_sync(0xffffffffu, nz);
            if (PASS && nz) {
                const int pos = base + total + __popc(bal & ((1u << lane) - 1u));
                keys[pos] = colkey | (unsigned)r[e]; oval[pos] = v[e];
            }
            total += __popc(bal);
        }
    }
    if (!PASS && lane == 0) counts[j] = total;
}

__global__ void keys_low_kernel(const unsigned long long *keys, int64_t n, int *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)(keys[i] & 0xffffffffull);
}

__global__ void swap_keys_kernel(const int *ptr, const int *idx, int n_major, unsigned long long *keys) {
    // one warp per major index: key = (minor << 32 | major)
    const int lane = threadIdx.x & 31;
    const int j = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (j >= n_major) return;
    for (int p = ptr[j] + lane; p < ptr[j + 1]; p += 32) keys[p] = (((unsigned long long)(unsigned)idx[p]) << 32) | (unsigned)j;
}

__global__ void hist_high_kernel(const unsigned long long *keys, int64_t n, int *counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&counts[(int)(keys[i] >> 32)], 1);
}

static int bits_for(int n) { int b = 1; while ((1ll << b) < (long long)n && b < 32) ++b; return b; }

// sorts (keys, vals) by key using CUB, result in (keys_out, vals_out)
static int sort_pairs_u64_f32(const unsigned long long *k_in, unsigned long long *k_out, const float *v_in,
                              float *v_out, int64_t n, int end_bit, cudaStream_t st) {
    size_t tmp = 0;
    RT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, v_in, v_out, n, 0, end_bit, st));
    void *d_tmp = rt::scratch(SCR_CUB, tmp);
    if (!d_tmp) return RT_ERR_CUDA;
    RT_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp, k_in, k_out, v_in, v_out, n, 0, end_bit, st));
    rt::count_launch(4);
    return RT_OK;
}

static int exclusive_sum_i32(const int *in, int *out, int n, cudaStream_t st) {
    size_t tmp = 0;
    RT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, n, st));
    void *d_tmp = rt::scratch(SCR_CUB, tmp);
    if (!d_tmp) return RT_ERR_CUDA;
    RT_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp, in, out, n, st));
    rt::count_launch(1);
    return RT_OK;
}

}  // namespace rt

using namespace rt;

extern "C" int rt_w_merge(int32_t n_items, const int32_t *d_old_ptr, const int32_t *d_old_idx,
                          const float *d_old_val, int32_t n_old_items, const int32_t *d_targets,
                          int32_t n_targets, const int64_t *d_off, const int32_t *d_cnt, const int32_t *d_rows,
                          const float *d_vals, int32_t rows_sorted, int32_t *d_wptr, int32_t *d_widx,
                          float *d_wval, int64_t out_cap, int64_t *h_nnz, void *stream) {
    RT_ARG(n_items > 0, "n_items");
    RT_ARG(n_old_items >= 0 && n_old_items <= n_items, "old matrix larger than new one");
    RT_ARG(n_old_items == 0 || (d_old_ptr && (d_old_idx || true)), "old matrix pointers");
    RT_ARG(n_targets == 0 || (d_targets && d_off && d_cnt && d_rows && d_vals), "solver output pointers");
    RT_ARG(d_wptr != nullptr, "d_wptr");
    cudaStream_t st = (cudaStream_t)stream;
    // scratch A: tmap[n_items] | counts[n_items+1]
    const size_t a_bytes = align_up(sizeof(int) * (size_t)n_items) + align_up(sizeof(int) * ((size_t)n_items + 1));
    char *sa = (char *)rt::scratch(SCR_WMAT_A, a_bytes);
    if (!sa) return RT_ERR_CUDA;
    int *tmap = (int *)sa;
    int *counts = (int *)(sa + align_up(sizeof(int) * (size_t)n_items));
    const int bs = 256;
    fill_i32_kernel<<<(n_items + bs - 1) / bs, bs, 0, st>>>(tmap, n_items, -1);
    RT_CHECK_LAUNCH();
    if (n_targets > 0) {
        scatter_tmap_kernel<<<(n_targets + bs - 1) / bs, bs, 0, st>>>(d_targets, n_targets, n_items, tmap);
        RT_CHECK_LAUNCH();
    }
    RT_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * ((size_t)n_items + 1), st));
    const unsigned grid = (unsigned)(((int64_t)n_items * 32 + bs - 1) / bs);
    w_merge_kernel<0><<<grid, bs, 0, st>>>(n_items, d_old_ptr, d_old_idx, d_old_val, n_old_items, tmap, d_off, d_cnt,
                                          d_rows, d_vals, rows_sorted, counts, nullptr, nullptr, nullptr);
    RT_CHECK_LAUNCH();
    int rc = exclusive_sum_i32(counts, d_wptr, n_items + 1, st);
    if (rc) return rc;
    int nnz32 = 0;
    RT_CUDA(cudaMemcpyAsync(&nnz32, d_wptr + n_items, sizeof(int), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    const int64_t nnz = nnz32;
    if (h_nnz) *h_nnz = nnz;
    if (nnz > out_cap) {
        rt::set_error("rt_w_merge: output capacity %lld < nnz %lld", (long long)out_cap, (long long)nnz);
        return RT_ERR_CAPACITY;
    }
    if (nnz == 0) return RT_OK;
    RT_ARG(d_widx && d_wval, "output arrays");
    // scratch B: keys_in | keys_out | vals_in
    const size_t kb = align_up(sizeof(unsigned long long) * (size_t)nnz);
    char *sb = (char *)rt::scratch(SCR_WMAT_B, 2 * kb + align_up(sizeof(float) * (size_t)nnz));
    if (!sb) return RT_ERR_CUDA;
    unsigned long long *k_in = (unsigned long long *)sb, *k_out = (unsigned long long *)(sb + kb);
    float *v_in = (float *)(sb + 2 * kb);
    w_merge_kernel<1><<<grid, bs, 0, st>>>(n_items, d_old_ptr, d_old_idx, d_old_val, n_old_items, tmap, d_off, d_cnt,
                                          d_rows, d_vals, rows_sorted, nullptr, d_wptr, k_in, v_in);
    RT_CHECK_LAUNCH();
    rc = sort_pairs_u64_f32(k_in, k_out, v_in, d_wval, nnz, 32 + bits_for(n_items), st);
    if (rc) return rc;
    keys_low_kernel<<<(unsigned)((nnz + bs - 1) / bs), bs, 0, st>>>(k_out, nnz, d_widx);
    RT_CHECK_LAUNCH();
    RT_CUDA(cudaStreamSynchronize(st));
    return RT_OK;
}

extern "C" int rt_transpose(int32_t n_major_in, int32_t n_major_out, const int32_t *d_ptr, const int32_t *d_idx,
                            const float *d_val, int64_t nnz, int32_t *d_optr, int32_t *d_oidx, float *d_oval,
                            void *stream) {
    RT_ARG(n_major_in > 0 && n_major_out > 0 && nnz >= 0, "shape");
    RT_ARG(d_ptr && d_optr, "pointers");
    cudaStream_t st = (cudaStream_t)stream;
    const int bs = 256;
    const size_t c_bytes = align_up(sizeof(int) * ((size_t)n_major_out + 1));
    int *counts = (int *)rt::scratch(SCR_WMAT_A, c_bytes);
    if (!counts) return RT_ERR_CUDA;
    RT_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * ((size_t)n_major_out + 1), st));
    if (nnz == 0) {
        RT_CUDA(cudaMemsetAsync(d_optr, 0, sizeof(int) * ((size_t)n_major_out + 1), st));
        return RT_OK;
    }
    RT_ARG(d_idx && d_val && d_oidx && d_oval, "pointers");
    const size_t kb = align_up(sizeof(unsigned long long) * (size_t)nnz);
    char *sb = (char *)rt::scratch(SCR_WMAT_B, 2 * kb);
    if (!sb) return RT_ERR_CUDA;
    unsigned long long *k_in = (unsigned long long *)sb, *k_out = (unsigned long long *)(sb + kb);
    const unsigned grid = (unsigned)(((int64_t)n_major_in * 32 + bs - 1) / bs);
    swap_keys_kernel<<<grid, bs, 0, st>>>(d_ptr, d_idx, n_major_in, k_in);
    RT_CHECK_LAUNCH();
    hist_high_kernel<<<(unsigned)((nnz + bs - 1) / bs), bs, 0, st>>>(k_in, nnz, counts);
    RT_CHECK_LAUNCH();
    int rc = exclusive_sum_i32(counts, d_optr, n_major_out + 1, st);
    if (rc) return rc;
    rc = sort_pairs_u64_f32(k_in, k_out, d_val, d_oval, nnz, 32 + bits_for(n_major_out), st);
    if (rc) return rc;
    keys_low_kernel<<<(unsigned)((nnz + bs - 1) / bs), bs, 0, st>>>(k_out, nnz, d_oidx);
    RT_CHECK_LAUNCH();
    return RT_OK;
}
