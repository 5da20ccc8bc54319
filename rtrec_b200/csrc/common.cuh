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
// Grow-only device scratch owned by the library (one buffer per slot).  Reused across calls: the
// library assumes one caller thread and stream-ordered use (see include/rtrec_b200.h).
void *scratch(int slot, size_t bytes);
enum { SCR_SOLVE = 0, SCR_WMAT_A, SCR_WMAT_B, SCR_CUB, SCR_STORE_A, SCR_STORE_B, SCR_MISC, SCR_SLOTS };

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
