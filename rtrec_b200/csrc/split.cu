// split.cu -- range-split row pointers for a CSR/CSC matrix.
//
// seg[row * (n_ranges + 1) + g] = first position p in [ptr[row], ptr[row+1]) whose minor index is
// >= base + g * range_width (g = 0..n_ranges; the last one equals ptr[row+1] when the ranges cover
// the tail).  Minor indices must be ascending within a row.  With these pointers a warp that owns
// minor range g can stream exactly its part of every row, so CTAs can keep a dense accumulator
// tile in shared memory with each warp owning a disjoint slice: no atomics, no barriers
// (gram2.cu, score.cu).
#include "common.cuh"

namespace rt {

__global__ void csr_split_kernel(int n_rows, const int *__restrict__ ptr, const int *__restrict__ idx, int base,
                                 int range_width, int n_ranges, int *__restrict__ seg) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = n_ranges + 1;
    if (t >= (int64_t)n_rows * stride) return;
    const int row = (int)(t / stride), g = (int)(t - (int64_t)row * stride);
    int lo = ptr[row], hi = ptr[row + 1];
    const int64_t bound64 = (int64_t)base + (int64_t)g * range_width;
    if (g == n_ranges && bound64 > 0x7ffffff0ll) { seg[t] = hi; return; }
    const int bound = (int)bound64;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (idx[mid] < bound) lo = mid + 1; else hi = mid;
    }
    seg[t] = lo;
}

}  // namespace rt

using namespace rt;

extern "C" int rt_csr_split(int32_t n_rows, const int32_t *d_ptr, const int32_t *d_idx, int32_t base,
                            int32_t range_width, int32_t n_ranges, int32_t *d_seg, void *stream) {
    RT_ARG(n_rows >= 0 && range_width > 0 && n_ranges > 0, "shape");
    if (n_rows == 0) return RT_OK;
    RT_ARG(d_ptr && d_seg, "null pointer");
    const int64_t total = (int64_t)n_rows * (n_ranges + 1);
    const int bs = 256;
    csr_split_kernel<<<(unsigned)((total + bs - 1) / bs), bs, 0, (cudaStream_t)stream>>>(n_rows, d_ptr, d_idx, base,
                                                                                         range_width, n_ranges, d_seg);
    RT_CHECK_LAUNCH();
    return RT_OK;
}
