// gram_tc.cu -- K3 (v4): the dense head of the item-item Gram matrix on the tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces, for the GH_HEAD most popular items, the same reference lines as gram3.cu:
// /root/reference/rtrec/models/internal/slim_elastic.py:141 (`X.T.dot(y)` per target column) and the implicit Gram products
// of the coordinate descent (:229-281).
//
// Why: in popularity-rank space the corner G[0:2048, 0:2048] holds more than half of the multiply-adds of the whole Gram
// matrix at the ML-20M shape (every pair of popular items a user rated), and the lower-triangle kernel of gram3.cu spends
// them as shared-memory atomics (~1e12 multiply-adds/s).  The corner is a SYRK
//     G_h = A^T A,   A = X[:, head]  (n_users x 2048)
// and dense enough for the tensor cores to win although most of A is zero: 136 lower-triangle tiles of 128 x 128, one per
// CTA (a single wave on 148 SMs), K = n_users streamed by TMA from a bf16 copy of A stored item-major (K-major for both
// operands, 128-byte swizzle), fp32 accumulation in TMEM.
//
// Exactness: the path is taken only when every stored value is exact in bf16 (integer and half-integer ratings are) and
//     max_j G[j][j] / unit^2 < 2^24,   unit = the largest power of two dividing all values,
// i.e. when every partial sum (bounded by sqrt(G[i][i] G[j][j]) whatever the signs and the order) is an integer multiple of
// unit^2 below 2^24 and therefore exact in fp32 in ANY order.  The
// tensor-core corner is then bit-identical to the exact sums, like the integer-rating results of the other kernels.
// Decayed (continuous) values never qualify and stay on the sparse kernel.
//
// Cost rule: dense multiply-adds (136 * 128^2 * n_users) against the sparse ones the corner would cost
// (sum over users of h(h+1)/2, h = the user's items inside the head); taken when the ratio is below GH_RATIO.
#include <math.h>
#include <string.h>

#include "tc_common.cuh"

namespace rt {

constexpr int GH_HEAD = 2048;             // head rows/columns (16 tiles of 128)
constexpr int GH_T = GH_HEAD / 128;
constexpr int GH_K = 64;                  // users per operand tile (one 128-byte swizzle row of bf16)
constexpr int GH_STAGES = 6;
constexpr int GH_TILE_BYTES = 128 * GH_K * 2;
constexpr int GH_THREADS = 192;           // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue (one TMEM lane quadrant each)
constexpr double GH_RATIO = 400.0;        // dense / sparse multiply-adds up to which the tensor cores are used

// stats[0] = sum over users of h(h+1)/2, [1] = values that are not exact in bf16, [2] = max |value| (float bits),
// [3] = min over non-zero values of (exponent of the lowest set bit) + 1024, [4] = largest diagonal entry of G (double bits)
__global__ void gh_value_stats_kernel(const float *__restrict__ vals, int64_t n, unsigned long long *__restrict__ stats) {
    unsigned inexact = 0, vmax = 0;
    int emin = 4096;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const float v = vals[e];
        const unsigned b = __float_as_uint(v) & 0x7fffffffu;
        if (b == 0) continue;
        inexact += (__bfloat162float(__float2bfloat16_rn(v)) != v) || b >= 0x7f800000u || b < 0x00800000u;   // (inf/nan/denormal: no)
        vmax = max(vmax, b);
        const int ex = (int)(b >> 23) - 127;
        const unsigned man = (b & 0x7fffffu) | 0x800000u;
        emin = min(emin, ex - 23 + (__ffs((int)man) - 1) + 1024);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        inexact += __shfl_xor_sync(0xffffffffu, inexact, o);
        vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        emin = min(emin, __shfl_xor_sync(0xffffffffu, emin, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (inexact) atomicAdd(&stats[1], (unsigned long long)inexact);
        if (vmax) atomicMax(&stats[2], (unsigned long long)vmax);
        if (emin < 4096) atomicMin(&stats[3], (unsigned long long)emin);
    }
}

// one thread per user: h = entries of the rank-sorted row below GH_HEAD (binary search), accumulates h(h+1)/2
__global__ void gh_head_work_kernel(int n_users, const int *__restrict__ rptr, const int *__restrict__ pidx,
                                    unsigned long long *__restrict__ stats) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long w = 0;
    if (u < n_users) {
        const int a = rptr[u];
        int lo = a, hi = rptr[u + 1];
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (pidx[mid] < GH_HEAD) lo = mid + 1; else hi = mid; }
        const unsigned long long h = (unsigned long long)(lo - a);
        w = h * (h + 1) / 2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&stats[0], w);
}

// one warp per item: stats[4] = max over items of sum_u x^2 (double bits; non-negative doubles order like their bit patterns).
// Every |G[i][j]| and every partial sum of it, in any order, is at most sqrt(G[i][i] G[j][j]) <= this maximum.
__global__ void gh_col_sumsq_kernel(int n_items, const int *__restrict__ cptr, const float *__restrict__ cval,
                                    unsigned long long *__restrict__ stats) {
    const int j = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (j >= n_items) return;
    double s = 0.0;
    for (int e = cptr[j] + lane; e < cptr[j + 1]; e += 32) { const double v = (double)cval[e]; s += v * v; }
    s = warp_sum(s);
    if (lane == 0 && s > 0.0) atomicMax(&stats[4], (unsigned long long)__double_as_longlong(s));
}

// A^T as bf16, item-major: At[r][u] = X[u, orig_of[r]] for the head ranks r (the buffer is zeroed before)
__global__ void __launch_bounds__(256) gh_densify_kernel(const int *__restrict__ orig_of, const int *__restrict__ cptr,
                                                         const int *__restrict__ cidx, const float *__restrict__ cval,
                                                         __nv_bfloat16 *__restrict__ At, int64_t kp) {
    const int r = blockIdx.x;
    const int j = orig_of[r];
    __nv_bfloat16 *row = At + (size_t)r * kp;
    const int a = cptr[j], b = cptr[j + 1];
    for (int e = a + blockIdx.y * blockDim.x + threadIdx.x; e < b; e += gridDim.y * blockDim.x) row[cidx[e]] = __float2bfloat16_rn(cval[e]);
}

__global__ void __launch_bounds__(GH_THREADS, 1) gram_head_tc_kernel(const __grid_constant__ CUtensorMap tmap, int n_k, float *__restrict__ Gp,
                                                                      int64_t ldgp) {
    extern __shared__ unsigned char gh_smem_raw[];
    unsigned char *smem = gh_smem_raw + ((1024u - (smem_u32(gh_smem_raw) & 1023u)) & 1023u);   // swizzle atoms: 1024-byte alignment
    unsigned char *sA = smem;                                   // GH_STAGES tiles of the row block
    unsigned char *sB = sA + GH_STAGES * GH_TILE_BYTES;         // GH_STAGES tiles of the column block (unused on the diagonal)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sB + GH_STAGES * GH_TILE_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * GH_STAGES + 2);
    const uint32_t bar0 = smem_u32(bars);
    auto b_full = [&](int s) { return bar0 + 8u * (uint32_t)s; };
    auto b_empty = [&](int s) { return bar0 + 8u * (uint32_t)(GH_STAGES + s); };
    const uint32_t acc_full = bar0 + 8u * (uint32_t)(2 * GH_STAGES);

    // lower-triangle tile of this CTA: linear index -> (ti, tj), tj <= ti
    const int t = blockIdx.x;
    int ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    const int tj = t - ti * (ti + 1) / 2;
    const bool diag = ti == tj;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < GH_STAGES; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: 64 users of the row block (and of the column block) per stage
        if (lane == 0) {
            for (int ks = 0; ks < n_k; ++ks) {
                const int s = ks % GH_STAGES;
                mbar_wait(b_empty(s), ((ks / GH_STAGES) & 1) ^ 1);
                mbar_arrive_expect_tx(b_full(s), diag ? GH_TILE_BYTES : 2 * GH_TILE_BYTES);
                tma_load_2d(smem_u32(sA + s * GH_TILE_BYTES), &tmap, b_full(s), ks * GH_K, ti * 128);
                if (!diag) tma_load_2d(smem_u32(sB + s * GH_TILE_BYTES), &tmap, b_full(s), ks * GH_K, tj * 128);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread): D[i][j] += sum_u At[ti*128 + i][u] * At[tj*128 + j][u]
        if (lane == 0) {
            uint32_t acc = 0;
            for (int ks = 0; ks < n_k; ++ks) {
                const int s = ks % GH_STAGES;
                mbar_wait(b_full(s), (ks / GH_STAGES) & 1);
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA + s * GH_TILE_BYTES);
                const uint32_t b_addr = diag ? a_addr : smem_u32(sB + s * GH_TILE_BYTES);
#pragma unroll
                for (int k4 = 0; k4 < GH_K / 16; ++k4) {
                    umma_bf16(tmem_base, umma_desc_k128(a_addr + k4 * 32), umma_desc_k128(b_addr + k4 * 32), acc);
                    acc = 1;
                }
                umma_commit(b_empty(s));
            }
            umma_commit(acc_full);
        }
    } else {
        // ===== epilogue: warp w reads TMEM lanes 32*(w%4)..+31 (rows of the tile), 32 columns at a time, and stores fp32
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        float *dst = Gp + (size_t)(ti * 128 + row) * ldgp + tj * 128;
        const bool vec = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(cc * 32), v);
            if (vec) {
#pragma unroll
                for (int c = 0; c < 32; c += 4)
                    *reinterpret_cast<float4 *>(dst + cc * 32 + c) = make_float4(__uint_as_float(v[c]), __uint_as_float(v[c + 1]),
                                                                                 __uint_as_float(v[c + 2]), __uint_as_float(v[c + 3]));
            } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) dst[cc * 32 + c] = __uint_as_float(v[c]);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

// Decides whether the head corner goes to the tensor cores and, if so, computes it into d_Gp[0:GH_HEAD, 0:GH_HEAD] (lower
// tiles; the mirror pass fills the rest).  *h_head = GH_HEAD when taken (the sparse kernel then starts at that row), else 0.
// Synchronises the stream once (a 40-byte read-back of the statistics).
int gram_head_tc(int n_users, int n_items, const int *d_cptr, const int *d_cidx, const float *d_cval, int64_t nnz,
                 const int *d_rptr, const int *d_pidx, const int *d_orig_of, float *d_Gp, int64_t ldgp, int *h_head, cudaStream_t st) {
    *h_head = 0;
    if (n_items < GH_HEAD || n_users < 16384 || nnz <= 0) return RT_OK;
    int cc_major = 0;
    rt_device_info(nullptr, nullptr, &cc_major, nullptr);
    if (cc_major != 10) return RT_OK;
    tc_encode_fn enc = tc_encoder();
    if (!enc) return RT_OK;
    static unsigned long long *d_stats = nullptr;
    if (!d_stats) RT_CUDA(cudaMalloc(&d_stats, 8 * sizeof(unsigned long long)));
    unsigned long long h[5] = {0, 0, 0, ~0ull, 0};
    RT_CUDA(cudaMemcpyAsync(d_stats, h, sizeof(h), cudaMemcpyHostToDevice, st));
    gh_value_stats_kernel<<<rt::sm_count() * 8, 256, 0, st>>>(d_cval, nnz, d_stats);
    RT_CHECK_LAUNCH();
    gh_head_work_kernel<<<(n_users + 255) / 256, 256, 0, st>>>(n_users, d_rptr, d_pidx, d_stats);
    RT_CHECK_LAUNCH();
    gh_col_sumsq_kernel<<<(unsigned)(((int64_t)n_items * 32 + 255) / 256), 256, 0, st>>>(n_items, d_cptr, d_cval, d_stats);
    RT_CHECK_LAUNCH();
    RT_CUDA(cudaMemcpyAsync(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    if (h[1] != 0 || h[2] == 0 || h[3] == ~0ull) return RT_OK;                         // inexact in bf16, or all zero
    const double dense = (double)(GH_T * (GH_T + 1) / 2) * 128.0 * 128.0 * (double)n_users;
    if (dense > GH_RATIO * (double)h[0]) return RT_OK;                                   // too sparse to pay
    {
        // every partial sum an integer multiple of unit^2 below 2^24
        double diag_max;
        const unsigned long long db = h[4];
        memcpy(&diag_max, &db, sizeof(diag_max));
        const double unit = ldexp(1.0, (int)h[3] - 1024);
        if (diag_max / (unit * unit) >= 16777216.0) return RT_OK;
    }
    const int64_t kp = ((int64_t)n_users + GH_K - 1) / GH_K * GH_K;
    __nv_bfloat16 *At = (__nv_bfloat16 *)rt::scratch(SCR_GRAM_HEAD, (size_t)GH_HEAD * (size_t)kp * 2 + 1024);
    if (!At) return RT_ERR_CUDA;
    RT_CUDA(cudaMemsetAsync(At, 0, (size_t)GH_HEAD * (size_t)kp * 2, st));
    gh_densify_kernel<<<dim3(GH_HEAD, 8), 256, 0, st>>>(d_orig_of, d_cptr, d_cidx, d_cval, At, kp);
    RT_CHECK_LAUNCH();
    CUtensorMap tmap;
    {
        const cuuint64_t gdim[2] = {(cuuint64_t)kp, (cuuint64_t)GH_HEAD};
        const cuuint64_t gstride[1] = {(cuuint64_t)kp * 2};
        const cuuint32_t box[2] = {(cuuint32_t)GH_K, 128u};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, At, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { rt::set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return RT_ERR_CUDA; }
    }
    const size_t smem = 1024 + (size_t)2 * GH_STAGES * GH_TILE_BYTES + (2 * GH_STAGES + 2) * 8 + 64;
    RT_CUDA(cudaFuncSetAttribute(gram_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gram_head_tc_kernel<<<GH_T * (GH_T + 1) / 2, GH_THREADS, smem, st>>>(tmap, (int)(kp / GH_K), d_Gp, ldgp);
    RT_CHECK_LAUNCH();
    *h_head = GH_HEAD;
    return RT_OK;
}

}  // namespace rt
