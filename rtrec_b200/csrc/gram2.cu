// gram2.cu -- K3 (v2): item-item Gram rows with warp-private shared-memory accumulators.
//
// Same result as gram.cu (G[j,:] = X^T x_j, replaces `X.T.dot(y)` of
// /root/reference/rtrec/models/internal/slim_elastic.py:141 for all targets at once) without the
// global fp32 atomics: the item axis is cut into R ranges of RW items; a warp-level task is
// (target column j, chunk of its raters, item range g).  The warp keeps the RW partial sums of
// G[j, g*RW : (g+1)*RW] in its private slice of shared memory, walks the chunk's raters and
// streams exactly the part of each rater's CSR row that falls into range g (range-split row
// pointers, split.cu), then writes the slice to the G row once (plain coalesced store when the
// column has a single chunk, fp32 RED otherwise).  Warps never share accumulator entries, so
// there are no atomics and no barriers in the hot loop, and per (j, i) the summation order is
// ascending user id -- the order scipy's csr_matvec uses.
//
// Task ids are (chunk, range)-major, so the warps of a CTA that fetch consecutive tasks work on
// the same raters and read neighbouring segments of the same CSR rows (L1/L2 locality).
//
// Algorithmic bytes per target column (SURVEY.md 8d, K3 term): e*(S_j + nnz_j) + 4*n_items written.
#include <vector>

#include "common.cuh"

namespace rt {

constexpr int G2_WARPS = 8;          // warps per CTA
constexpr int G2_SLICE = 1728;       // floats per warp slice (8 * 1728 * 4 B = 54 KB per CTA, 4 CTAs/SM)
constexpr int G2_CHUNK = 4096;       // raters per chunk

struct GramChunk {
    int j;       // target column
    int c0, c1;  // entry range in the CSC arrays
    int single;  // 1 = only chunk of this column (plain store), 0 = accumulate with RED
};

__global__ void __launch_bounds__(G2_WARPS * 32)
gram_slices_kernel(const GramChunk *__restrict__ chunks, int n_chunks, int R, int RW, int n_items,
                   const int *__restrict__ cidx, const float *__restrict__ cval, const int *__restrict__ xseg,
                   const int *__restrict__ ridx, const float *__restrict__ rval, float *__restrict__ G, int64_t ldg,
                   unsigned long long *__restrict__ counter) {
    extern __shared__ __align__(16) float g2_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *slice = g2_smem + (size_t)warp * G2_SLICE;
    const unsigned long long n_tasks = (unsigned long long)n_chunks * (unsigned)R;
    const int stride = R + 1;
    for (;;) {
        unsigned long long task = 0;
        if (lane == 0) task = atomicAdd(counter, 1ull);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= n_tasks) break;
        const int cc = (int)(task / (unsigned)R), g = (int)(task - (unsigned long long)cc * (unsigned)R);
        const GramChunk ch = chunks[cc];
        const int lo = g * RW;
        const int width = min(RW, n_items - lo);
        if (width <= 0) continue;
        for (int x = lane; x < width; x += 32) slice[x] = 0.0f;
        __syncwarp();
        bool touched = false;
        for (int base = ch.c0; base < ch.c1; base += 32) {
            const int e = base + lane;
            int a = 0, b = 0;
            float y = 0.f;
            if (e < ch.c1) {
                const int u = cidx[e];
                y = cval[e];
                a = xseg[(size_t)u * stride + g];
                b = xseg[(size_t)u * stride + g + 1];
            }
            unsigned mask = __ballot_sync(0xffffffffu, b > a);
            touched |= (mask != 0u);
            while (mask) {
                const int l = __ffs(mask) - 1;
                mask &= mask - 1;
                const float yy = __shfl_sync(0xffffffffu, y, l);
                const int aa = __shfl_sync(0xffffffffu, a, l);
                const int bb = __shfl_sync(0xffffffffu, b, l);
                for (int p = aa + lane; p < bb; p += 32) {
                    const int x = ridx[p] - lo;
                    slice[x] = __fadd_rn(slice[x], __fmul_rn(yy, rval[p]));
                }
                __syncwarp();
            }
        }
        if (touched) {
            float *g_row = G + (size_t)ch.j * ldg + lo;
            if (ch.single) {
                for (int x = lane; x < width; x += 32) g_row[x] = slice[x];
            } else {
                for (int x = lane; x < width; x += 32) {
                    const float v = slice[x];
                    if (v != 0.0f) atomicAdd(g_row + x, v);
                }
            }
        }
        __syncwarp();
    }
}

}  // namespace rt

using namespace rt;

extern "C" int rt_gram(int32_t n_users, int32_t n_items, const int32_t *d_cptr, const int32_t *d_cidx,
                       const float *d_cval, const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                       int32_t j_begin, int32_t j_end, float *d_G, int64_t ldg, void *stream) {
    RT_ARG(n_users > 0 && n_items > 0 && j_begin >= 0 && j_end <= n_items && j_begin <= j_end, "shape / target range");
    if (j_begin == j_end) return RT_OK;
    RT_ARG(d_cptr && d_rptr && d_G && ldg >= n_items, "null pointer / ldg");
    cudaStream_t st = (cudaStream_t)stream;
    const int nj = j_end - j_begin;
    std::vector<int> cptr((size_t)nj + 1);
    RT_CUDA(cudaMemcpyAsync(cptr.data(), d_cptr + j_begin, sizeof(int) * ((size_t)nj + 1), cudaMemcpyDeviceToHost, st));
    RT_CUDA(cudaStreamSynchronize(st));
    if (cptr[nj] == cptr[0]) return RT_OK;
    RT_ARG(d_cidx && d_cval && d_ridx && d_rval, "null pointer");
    // geometry: R ranges of RW items, RW <= slice capacity
    const int R = (n_items + G2_SLICE - 1) / G2_SLICE;
    int RW = (n_items + R - 1) / R;
    RW = (RW + 31) & ~31;
    if (RW > G2_SLICE) RW = G2_SLICE;
    // chunk list
    std::vector<GramChunk> chunks;
    chunks.reserve((size_t)nj + (size_t)(cptr[nj] - cptr[0]) / G2_CHUNK + 8);
    for (int t = 0; t < nj; ++t) {
        const int a = cptr[t], b = cptr[t + 1];
        if (b <= a) continue;
        const int n_ch = (b - a + G2_CHUNK - 1) / G2_CHUNK;
        for (int c = 0; c < n_ch; ++c) {
            GramChunk ch;
            ch.j = j_begin + t;
            ch.c0 = a + c * G2_CHUNK;
            ch.c1 = ch.c0 + G2_CHUNK < b ? ch.c0 + G2_CHUNK : b;
            ch.single = n_ch == 1 ? 1 : 0;
            chunks.push_back(ch);
        }
    }
    const int n_chunks = (int)chunks.size();
    // scratch: counter | chunks | xseg
    const size_t seg_bytes = align_up(sizeof(int) * (size_t)n_users * (R + 1));
    const size_t ch_bytes = align_up(sizeof(GramChunk) * (size_t)n_chunks);
    char *base = (char *)rt::scratch(SCR_STORE_B, 256 + ch_bytes + seg_bytes);
    if (!base) return RT_ERR_CUDA;
    unsigned long long *d_counter = (unsigned long long *)base;
    GramChunk *d_chunks = (GramChunk *)(base + 256);
    int *d_xseg = (int *)(base + 256 + ch_bytes);
    RT_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), st));
    RT_CUDA(cudaMemcpyAsync(d_chunks, chunks.data(), sizeof(GramChunk) * (size_t)n_chunks, cudaMemcpyHostToDevice, st));
    {
        int rc = rt_csr_split(n_users, d_rptr, d_ridx, 0, RW, R, d_xseg, stream);
        if (rc) return rc;
    }
    const size_t smem = sizeof(float) * (size_t)G2_WARPS * G2_SLICE;
    RT_CUDA(cudaFuncSetAttribute(gram_slices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = (int)((size_t)rt::smem_optin() / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t grid = (int64_t)rt::sm_count() * per_sm;
    const int64_t want = ((int64_t)n_chunks * R + G2_WARPS - 1) / G2_WARPS;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    gram_slices_kernel<<<(unsigned)grid, G2_WARPS * 32, smem, st>>>(d_chunks, n_chunks, R, RW, n_items, d_cidx, d_cval, d_xseg,
                                                                   d_ridx, d_rval, d_G, ldg, d_counter);
    RT_CHECK_LAUNCH();
    // the host vector `chunks` must outlive the async copy
    RT_CUDA(cudaStreamSynchronize(st));
    return RT_OK;
}
