// common.cuh -- shared helpers for librtrec_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rtrec_b200.h"

namespace rt {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
int sm_count();
int smem_optin();
// tuning switches settable through rt_set_option(): which kernel generation serves an entry point
enum { OPT_SCORE_IMPL = 0, OPT_GRAM_IMPL, OPT_SOLVE_IMPL, OPT_GRAM_SLICE, OPT_GRAM_RANGES, OPT_GRAM_ADAPT, OPT_GRAM_HEAD, OPT_COUNT };
int option(int key);
// Grow-only device scratch owned by the library (one buffer per slot).  Reused across calls: the
// library assumes one caller thread and stream-ordered use (see include/rtrec_b200.h).
void *scratch(int slot, size_t bytes);
enum { SCR_SOLVE = 0, SCR_WMAT_A, SCR_WMAT_B, SCR_CUB, SCR_STORE_A, SCR_STORE_B, SCR_MISC, SCR_SCORE, SCR_SOLVE_FLAGS, SCR_GRAM_PACK, SCR_GRAM_SEL, SCR_GRAM_HEAD, SCR_SLOTS };

#define RT_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            rt::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return RT_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define RT_CHECK_LAUNCH()                                                                          \
    do {                                                                                           \
        rt::count_launch();                                                                        \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess) {                                                                  \
            rt::set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return RT_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define RT_ARG(cond, msg)                                                                          \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            rt::set_error("bad argument: %s (%s:%d)", msg, __FILE__, __LINE__);                    \
            return RT_ERR_ARG;                                                                     \
        }                                                                                          \
    } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// carve typed arrays out of one workspace allocation
struct Carver {
    char *base;
    size_t off;
    size_t cap;
    Carver(void *p, size_t c) : base((char *)p), off(0), cap(c) {}
    template <typename T> T *take(size_t n) {
        off = align_up(off);
        T *r = (T *)(base + off);
        off += n * sizeof(T);
        return r;
    }
    bool ok() const { return off <= cap; }
};

}  // namespace rt
// v2 scoring kernel launcher (score2.cu), called by rt_slim_recommend
int rt_launch_recommend2(const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval, const int32_t *d_users,
                         int32_t n_query, const int32_t *d_wrptr, const int32_t *d_wridx, const float *d_wrval,
                         int32_t n_items, int32_t j_begin, int32_t j_end, int32_t k, int32_t filter_interacted,
                         int32_t mode, int32_t *d_out_ids, float *d_out_scores, int32_t *d_out_cnt, int *d_next,
                         cudaStream_t st);
int rt_gram_rows_selected(const int32_t *d_ccol, const int32_t *d_cidx, const float *d_cval, int64_t nnz, const int32_t *d_rptr,
                          const int32_t *d_ridx, const float *d_rval, const int32_t *d_row_slot, float *d_G, int64_t ldg,
                          cudaStream_t st);
namespace rt {
// tensor-core head of the Gram matrix (gram_tc.cu): *h_head = rows taken (0 = not applicable / does not pay)
int gram_head_tc(int n_users, int n_items, const int *d_cptr, const int *d_cidx, const float *d_cval, int64_t nnz,
                 const int *d_rptr, const int *d_pidx, const int *d_orig_of, float *d_Gp, int64_t ldgp, int *h_head, cudaStream_t st);
}
extern "C" int rt_csr_split(int32_t n_rows, const int32_t *d_ptr, const int32_t *d_idx, int32_t base,
                            int32_t range_width, int32_t n_ranges, int32_t *d_seg, void *stream);
namespace rt {
// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_key(float f) {
    // order-preserving map float -> uint32 (larger float => larger key); -0.0 folded into +0.0
    f = f + 0.0f;
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace rt
