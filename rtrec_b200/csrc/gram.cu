// gram.cu -- K3: item-item Gram rows G[j,:] = X^T x_j by row gather.
//
// Replaces the per-column `X.T.dot(y)` of FeatureSelectionWrapper.fit
// (/root/reference/rtrec/models/internal/slim_elastic.py:141) for ALL target columns at once and
// at the same time produces the Gram matrix the solver replays on (solve.cu).
//
// Work unit = one stored entry e of the column-sorted COO view of X: (j = ccol[e], u = cidx[e],
// y = cval[e]); it contributes y * X[u, :] to row j of G.  A warp takes 32 consecutive entries,
// broadcasts them one by one and streams the user's CSR row with coalesced 128 B loads; the
// accumulation is a fire-and-forget fp32 RED to the (L2-resident) G row.  Consecutive entries
// belong to the same column, so the 4*n_items-byte destination row stays hot in L2 while the
// column is being processed, and is written back to HBM once.
//
// Algorithmic bytes per entry (SURVEY.md 8d, K3 term): e * (1 + nnz(row u)), e = 8 B.
#include "common.cuh"

namespace rt {

constexpr int GRAM_WARPS = 8;

__global__ void __launch_bounds__(GRAM_WARPS * 32)
gram_rows_kernel(const int *__restrict__ ccol, const int *__restrict__ cidx, const float *__restrict__ cval,
                 int64_t e_begin, int64_t e_end, const int *__restrict__ rptr, const int *__restrict__ ridx,
                 const float *__restrict__ rval, float *__restrict__ G, int64_t ldg,
                 unsigned long long *__restrict__ counter, const int *__restrict__ row_slot) {
    // row_slot (optional): only the Gram rows of the items with row_slot[j] >= 0 are formed, row j at G + row_slot[j] * ldg
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned long long c = 0;
        if (lane == 0) c = atomicAdd(counter, 32ull);
        c = __shfl_sync(0xffffffffu, c, 0);
        const int64_t base = e_begin + (int64_t)c;
        if (base >= e_end) break;
        const int64_t e = base + lane;
        int j = 0, r0 = 0, r1 = 0;
        float y = 0.f;
        if (e < e_end) {
            j = ccol[e];
            if (row_slot) j = row_slot[j];
            if (j >= 0) {
                const int u = cidx[e];
                y = cval[e];
                r0 = rptr[u];
                r1 = rptr[u + 1];
            }
        }
        if (row_slot && !__any_sync(0xffffffffu, r1 > r0)) continue;   // none of these 32 entries belongs to a wanted row
        const int nb = (int)min((int64_t)32, e_end - base);
        for (int l = 0; l < nb; ++l) {
            const int jj = __shfl_sync(0xffffffffu, j, l);
            const float yy = __shfl_sync(0xffffffffu, y, l);
            const int a = __shfl_sync(0xffffffffu, r0, l);
            const int b = __shfl_sync(0xffffffffu, r1, l);
            float *g = G + (size_t)jj * ldg;
            int p = a + lane;
            // two loads in flight per lane
            for (; p + 32 < b; p += 64) {
                const int i0 = ridx[p], i1 = ridx[p + 32];
                const float x0 = rval[p], x1 = rval[p + 32];
                atomicAdd(g + i0, yy * x0);
                atomicAdd(g + i1, yy * x1);
            }
            if (p < b) atomicAdd(g + ridx[p], yy * rval[p]);
        }
    }
}

}  // namespace rt

using namespace rt;

extern "C" int rt_gram_rows(const int32_t *d_ccol, const int32_t *d_cidx, const float *d_cval, int64_t e_begin,
                            int64_t e_end, const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                            float *d_G, int64_t ldg, void *stream) {
    RT_ARG(e_end >= e_begin, "entry range");
    if (e_end == e_begin) return RT_OK;
    RT_ARG(d_ccol && d_cidx && d_cval && d_rptr && d_ridx && d_rval && d_G, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    static unsigned long long *d_counter = nullptr;
    if (!d_counter) RT_CUDA(cudaMalloc(&d_counter, sizeof(unsigned long long)));
    RT_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
    const int64_t n = e_end - e_begin;
    int64_t want = (n + 31) / 32;                         // warps of work
    int64_t grid = (want + GRAM_WARPS - 1) / GRAM_WARPS;
    const int64_t cap = (int64_t)rt::sm_count() * 8;       // persistent: 8 CTAs x 8 warps per SM
    if (grid > cap) grid = cap;
    gram_rows_kernel<<<(unsigned)grid, GRAM_WARPS * 32, 0, st>>>(d_ccol, d_cidx, d_cval, e_begin, e_end, d_rptr, d_ridx,
                                                                d_rval, d_G, ldg, d_counter, nullptr);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

// Gram rows of selected items only (internal; used by rt_slim_fit_pruned): d_row_slot[j] = row of item j in d_G or -1
int rt_gram_rows_selected(const int32_t *d_ccol, const int32_t *d_cidx, const float *d_cval, int64_t nnz, const int32_t *d_rptr,
                          const int32_t *d_ridx, const float *d_rval, const int32_t *d_row_slot, float *d_G, int64_t ldg,
                          cudaStream_t st) {
    if (nnz <= 0) return RT_OK;
    static unsigned long long *d_counter = nullptr;
    if (!d_counter) RT_CUDA(cudaMalloc(&d_counter, sizeof(unsigned long long)));
    RT_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
    int64_t grid = ((nnz + 31) / 32 + GRAM_WARPS - 1) / GRAM_WARPS;
    const int64_t cap = (int64_t)rt::sm_count() * 8;
    if (grid > cap) grid = cap;
    gram_rows_kernel<<<(unsigned)grid, GRAM_WARPS * 32, 0, st>>>(d_ccol, d_cidx, d_cval, 0, nnz, d_rptr, d_ridx, d_rval, d_G, ldg,
                                                                d_counter, d_row_slot);
    RT_CHECK_LAUNCH();
    return RT_OK;
}
