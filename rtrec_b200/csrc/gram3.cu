// gram3.cu -- K3 (v3): item-item Gram matrix G = X^T X by popularity-ranked lower triangle.
//
// Replaces the per-column `X.T.dot(y)` of FeatureSelectionWrapper.fit
// (/root/reference/rtrec/models/internal/slim_elastic.py:141) for ALL target columns at once and
// produces the Gram matrix the solver replays on (solve.cu).  Same values as gram.cu / gram2.cu.
//
// Why a third generation (profiles/r1a_*, r1c_*): v1 issues one global fp32 RED per multiply-add
// (RED-issue bound, 28 GB of DRAM writes for a 2.9 GB matrix); v2 keeps warp-private accumulators
// in shared memory but, with item ids in arbitrary order, most (rater, item-range) segments hold a
// handful of entries, so 3 of 4 lanes idle.  v3 fixes the layout first:
//
//   1. items are relabelled by popularity rank (most rated first).  In rank space the matrix is
//      dense in the top-left corner and sparse elsewhere, whatever the id order of the caller;
//   2. G is symmetric, so only the lower triangle G'[j', i' <= j'] is computed: a rater u of item
//      j' contributes x_uj * (the prefix of its rank-sorted row up to j').  Half the work, and the
//      prefix of a row is one contiguous, coalesced stream;
//   3. the head of every row (ranks < R*RW) is accumulated in warp-private shared-memory slices
//      (no atomics, ascending-user summation order = scipy's csr_matvec order); the sparse tail
//      (ranks >= R*RW, short segments) goes to global memory with fp32 RED;
//   4. a tiled transpose mirrors the triangle, and a row-staged gather writes G back in the
//      caller's item ids, so nothing downstream knows about ranks.
//
// Multi-GPU: rows of G' are split between ranks by exact multiply-add count; each rank computes
// its slab (rt_gram_lower), slabs are all-gathered by the caller, every rank finishes locally
// (rt_gram_finish).
//
// Algorithmic bytes per target column (SURVEY.md 8d, K3 term): e*(S_j + nnz_j) + 4*n_items written;
// the symmetric formulation reads half of S_j.
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace rt {

constexpr int G3_WARPS = 8;       // warps per CTA
constexpr int G3_SLICE = 1728;    // floats per warp slice (8 * 1728 * 4 B = 54 KB per CTA, 4 CTAs/SM)
constexpr int G3_CHUNK = 2048;    // raters per task
constexpr int G3_MAX_R = 4;       // at most 4 shared-memory ranges (head = R * G3_SLICE ranks)

__global__ void rank_keys_kernel(const int *__restrict__ cptr, int n_items, unsigned long long *__restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const unsigned cnt = (unsigned)(cptr[i + 1] - cptr[i]);
    keys[i] = (((unsigned long long)(0xffffffffu - cnt)) << 32) | (unsigned)i;  // count desc, id asc
}

__global__ void rank_scatter_kernel(const unsigned long long *__restrict__ sorted, int n_items, int *__restrict__ orig_of,
                                    int *__restrict__ rank_of) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_items) return;
    const int id = (int)(sorted[r] & 0xffffffffull);
    orig_of[r] = id;
    rank_of[id] = r;
}

// one warp per user row: key = (user << 32 | rank of item)
__global__ void relabel_keys_kernel(int n_users, const int *__restrict__ rptr, const int *__restrict__ ridx,
                                    const int *__restrict__ rank_of, unsigned long long *__restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const int u = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (u >= n_users) return;
    const unsigned long long hi = ((unsigned long long)(unsigned)u) << 32;
    for (int p = rptr[u] + lane; p < rptr[u + 1]; p += 32) keys[p] = hi | (unsigned)rank_of[ridx[p]];
}

__global__ void low32_kernel(const unsigned long long *__restrict__ keys, int64_t n, int *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int)(keys[i] & 0xffffffffull);
}

// packed view of the rank-sorted rows for gram_lower_kernel<.., 2>: (rank - base of the range that holds it, value bits)
__global__ void pack_entries_kernel(const int *__restrict__ pidx, const float *__restrict__ pval, int64_t n, int RW, int R,
                                    int2 *__restrict__ pent) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int rank = pidx[i];
    const int g = min(rank / RW, R);
    pent[i] = make_int2(rank - g * RW, __float_as_int(pval[i]));
}

// ---- rank-sorted CSR by a segmented sort -----------------------------------------------------------------------------
// The Gram kernel wants every user's row ordered by the popularity rank of its items.  The rows are already contiguous
// (CSR), so this is a sort WITHIN each row -- not the two global radix sorts (rank bits, then user bits: five passes over
// 12-byte pairs) of the first version.  Rows of at most 64 entries are sorted by one warp in registers (two entries per
// lane, bitonic network over shuffles).  Longer rows are sorted by COUNTING: the ranks of one row are distinct numbers
// below n_items, so a team (a one-warp CTA for catalogues up to 64k items, a wider CTA above) marks them in a shared-memory
// bitmap of n_items bits, prefix-sums the population counts of the bitmap words, and every entry then reads its position:
// O(row + n_items / 32) per row, no compare-exchange network, any row length, and the host never needs the longest row
// (no device->host sync in front of the Gram kernel).  Catalogues whose bitmap does not fit one CTA's shared memory
// (> ~900k items) keep the radix path.

__device__ __forceinline__ void rs_cx(int &ka, float &va, int &kb, float &vb, bool up) {   // compare-exchange two entries
    if ((ka > kb) == up) { const int t = ka; ka = kb; kb = t; const float f = va; va = vb; vb = f; }
}

__global__ void __launch_bounds__(256) row_sort_small_kernel(int n_users, const int *__restrict__ rptr, const int *__restrict__ ridx,
                                                             const float *__restrict__ rval, const int *__restrict__ rank_of,
                                                             int *__restrict__ pidx, float *__restrict__ pval) {
    const int lane = threadIdx.x & 31;
    const int u = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (u >= n_users) return;
    const int a = rptr[u], n = rptr[u + 1] - a;
    if (n == 0 || n > 64) return;
    // element e of the network lives in lane e & 31, slot e >> 5; padding sorts to the end
    int k0 = 0x7fffffff, k1 = 0x7fffffff;
    float v0 = 0.f, v1 = 0.f;
    if (lane < n) { k0 = rank_of[ridx[a + lane]]; v0 = rval[a + lane]; }
    if (lane + 32 < n) { k1 = rank_of[ridx[a + lane + 32]]; v1 = rval[a + lane + 32]; }
#pragma unroll
    for (int size = 2; size <= 64; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride == 32) {
                // partner of element e is e ^ 32: the other slot of the same lane; direction from bit `size` of e (size = 64: up)
                rs_cx(k0, v0, k1, v1, true);
            } else {
                // both slots exchange with lane ^ stride
#pragma unroll
                for (int slot = 0; slot < 2; ++slot) {
                    int &k = slot ? k1 : k0;
                    float &v = slot ? v1 : v0;
                    const int e = lane + 32 * slot;
                    const int ok = __shfl_xor_sync(0xffffffffu, k, stride);
                    const float ov = __shfl_xor_sync(0xffffffffu, v, stride);
                    const bool up = (e & size) == 0;             // ascending block?
                    const bool lower = (e & stride) == 0;        // this element is the lower one of its pair
                    const bool take = lower ? ((k > ok) == up) : ((ok > k) == up);
                    if (take) { k = ok; v = ov; }
                }
            }
        }
    }
    if (lane < n) { pidx[a + lane] = k0; pval[a + lane] = v0; }
    if (lane + 32 < n) { pidx[a + lane + 32] = k1; pval[a + lane + 32] = v1; }
}

// rows of more than 64 entries: counting sort over a bitmap of ranks, one row per CTA at a time (blockDim.x = 32 .. 256)
__global__ void __launch_bounds__(256) row_sort_bitmap_kernel(int n_users, int n_items, const int *__restrict__ rptr,
                                                              const int *__restrict__ ridx, const float *__restrict__ rval,
                                                              const int *__restrict__ rank_of, int *__restrict__ pidx,
                                                              float *__restrict__ pval, int *__restrict__ next_row) {
    extern __shared__ __align__(16) unsigned rs_smem[];      // bits[nwords], then pre[nwords]
    __shared__ int s_claim, s_warp_tot[8];
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5;
    const int nwords = (n_items + 31) >> 5;
    unsigned *bits = rs_smem;
    int *pre = reinterpret_cast<int *>(rs_smem + nwords);
    const int seg = ((nwords + nt - 1) / nt) | 1;            // words per thread in the prefix pass (odd: no bank conflicts)
    for (;;) {
        __syncthreads();
        if (tid == 0) s_claim = atomicAdd(next_row, 32);
        __syncthreads();
        const int u0 = s_claim;
        if (u0 >= n_users) break;
        const int u1 = min(u0 + 32, n_users);
        for (int u = u0; u < u1; ++u) {
            const int a = rptr[u], n = rptr[u + 1] - a;       // (uniform: every thread reads the same rptr)
            if (n <= 64) continue;
            for (int w = tid; w < nwords; w += nt) bits[w] = 0u;
            __syncthreads();
            for (int e = tid; e < n; e += nt) {
                const int k = rank_of[ridx[a + e]];
                atomicOr(&bits[k >> 5], 1u << (k & 31));
            }
            __syncthreads();
            // exclusive prefix of the word populations: per-thread segment totals, scanned over the CTA
            const int w0 = min(tid * seg, nwords), w1 = min(w0 + seg, nwords);
            int c = 0;
            for (int w = w0; w < w1; ++w) c += __popc(bits[w]);
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_warp_tot[warp] = incl;
            __syncthreads();
            int run = incl - c;
            for (int x = 0; x < warp; ++x) run += s_warp_tot[x];
            for (int w = w0; w < w1; ++w) { pre[w] = run; run += __popc(bits[w]); }
            __syncthreads();
            for (int e = tid; e < n; e += nt) {
                const int k = rank_of[ridx[a + e]];
                const int pos = pre[k >> 5] + __popc(bits[k >> 5] & ((1u << (k & 31)) - 1u));
                pidx[a + pos] = k;
                pval[a + pos] = rval[a + e];
            }
            __syncthreads();
        }
    }
}

// one thread per stored entry e = (u, j): position of rank(j) inside the rank-sorted row of u; optionally
// the exact multiply-add count of every rank-column (sum of prefix lengths) for the multi-GPU partition
__global__ void entry_pos_kernel(int n_items, int64_t nnz, const int *__restrict__ rank_of, const int *__restrict__ cptr,
                                 const int *__restrict__ cidx, const int *__restrict__ rptr, const int *__restrict__ pidx,
                                 int *__restrict__ cpos, unsigned long long *__restrict__ cost, int blk_parts, int blk_me,
                                 int head) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = e < nnz;
    int jp = -1;
    unsigned long long work = 0;
    // columns of the CTA's first and last entry (two full searches per CTA); every thread then searches between them only
    __shared__ int s_col[2];
    if (threadIdx.x < 2) {
        const int64_t e0 = (int64_t)blockIdx.x * blockDim.x;
        const int ee = (int)(threadIdx.x == 0 ? e0 : min(e0 + (int64_t)blockDim.x, nnz) - 1);
        int lo = 0, hi = n_items;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cptr[mid] <= ee) lo = mid; else hi = mid; }
        s_col[threadIdx.x] = lo;
    }
    __syncthreads();
    if (valid) {
        // column of entry e: last j with cptr[j] <= e
        int lo = s_col[0], hi = s_col[1] + 1;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (cptr[mid] <= (int)e) lo = mid; else hi = mid; }
        jp = rank_of[lo];
        // block-cyclic mode: positions are only needed for the columns this part owns (cost == nullptr there)
        if (blk_parts > 0 && ((jp >> 6) % blk_parts) != blk_me) return;
        if (jp < head) return;     // column computed on the tensor cores (head > 0 only without a cost pass)
        const int u = cidx[e];
        const int a = rptr[u];
        int l2 = a, h2 = rptr[u + 1];
        while (l2 < h2) { const int mid = (l2 + h2) >> 1; if (pidx[mid] < jp) l2 = mid + 1; else h2 = mid; }
        cpos[e] = l2;
        work = (unsigned long long)(l2 - a + 1);
    }
    if (cost) {
        const int j0 = __shfl_sync(0xffffffffu, jp, 0);
        if (__all_sync(0xffffffffu, jp == j0)) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) work += __shfl_xor_sync(0xffffffffu, work, o);
            if ((threadIdx.x & 31) == 0 && j0 >= 0) atomicAdd(&cost[j0], work);
        } else if (valid) atomicAdd(&cost[jp], work);
    }
}

// blk_parts > 0: block-cyclic ownership (64-row block b of the rank-space matrix belongs to part b % blk_parts);
// rows of other parts get no chunks, so the task list of the lower-triangle kernel only holds own rows
__global__ void chunk_count_kernel(int n_items, const int *__restrict__ orig_of, const int *__restrict__ cptr,
                                   int *__restrict__ n_chunks, int blk_parts, int blk_me) {
    const int jp = blockIdx.x * blockDim.x + threadIdx.x;
    if (jp >= n_items) return;
    const int j = orig_of[jp];
    const bool mine = blk_parts <= 0 || ((jp >> 6) % blk_parts) == blk_me;
    n_chunks[jp] = mine ? (cptr[j + 1] - cptr[j] + G3_CHUNK - 1) / G3_CHUNK : 0;
}

// local row of rank-space row jp in the slab of its owner under block-cyclic ownership
__host__ __device__ __forceinline__ int blk_local_row(int jp, int blk_parts) { return (((jp >> 6) / blk_parts) << 6) | (jp & 63); }

// MODE (rt_set_option("gram_adapt", m); default 2: measured 12.03 -> 9.99 ms at the ML-20M shape, profiles/r3a_kbench_gram.log):
//   1  the four 32-entry batches of a rater's prefetch / update are guarded by warp-uniform tests on the segment length.
//      On the synthetic ML-20M shape 48 % of the issued batch slots hold an entry (ranges 1..3: 14-28 %, most segments
//      there are shorter than 32); with the guards it would be 83 % (CPU count over the real segment lengths, DESIGN.md
//      section 8);
//   2  as 1, and the rank-sorted rows are read as packed (index relative to the entry's own range, value) pairs
//      (`pent`, written by pack_entries_kernel): one 8-byte load and one address computation per entry instead of two,
//      no subtraction of the range base (~12 instead of ~16 instructions per batch slot).
// Every mode produces the same matrix bit for bit (same multiply-adds in the same order).
template <int SLICE, int MODE>
__global__ void __launch_bounds__(G3_WARPS * 32)
gram_lower_kernel(int row_begin, int row_end, const int *__restrict__ chunk_start, int R, int RW,
                  const int *__restrict__ orig_of, const int *__restrict__ cptr, const int *__restrict__ cidx,
                  const float *__restrict__ cval, const int *__restrict__ cpos, const int *__restrict__ hseg,
                  const int *__restrict__ pidx, const float *__restrict__ pval, float *__restrict__ Gp, int64_t ld,
                  unsigned long long *__restrict__ counter, int blk_parts, const int2 *__restrict__ pent) {
    constexpr bool ADAPT = MODE >= 1;
    constexpr bool PACKED = MODE == 2;
    extern __shared__ __align__(16) float g3_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *slice = g3_smem + (size_t)warp * SLICE;
    const int n_rows = row_end - row_begin;
    const int R1 = R + 1;
    const unsigned long long n_tasks = (unsigned long long)(chunk_start[n_rows] - chunk_start[0]) * (unsigned)R1;
    const int cs0 = chunk_start[0];
    for (;;) {
        unsigned long long task = 0;
        if (lane == 0) task = atomicAdd(counter, 1ull);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= n_tasks) break;
        const int cc = (int)(task / (unsigned)R1) + cs0, g = (int)(task % (unsigned)R1);
        // row of this chunk: last jr with chunk_start[jr] <= cc
        int lo_r = 0, hi_r = n_rows;
        while (hi_r - lo_r > 1) { const int mid = (lo_r + hi_r) >> 1; if (chunk_start[mid] <= cc) lo_r = mid; else hi_r = mid; }
        const int jp = row_begin + lo_r;
        const int gj = min(jp / RW, R);  // range that contains jp (R = tail)
        if (g > gj) continue;
        const int j = orig_of[jp];
        const int c0 = cptr[j], c1 = cptr[j + 1];
        const int e0 = c0 + (cc - chunk_start[lo_r]) * G3_CHUNK;
        const int e1 = min(e0 + G3_CHUNK, c1);
        // block-cyclic mode: the slab only holds the rows of this part, densely packed
        float *g_row = Gp + (size_t)(blk_parts > 0 ? blk_local_row(jp, blk_parts) : jp) * ld;
        if (g < R) {
            const int lo = g * RW;
            const int width = min(RW, jp + 1 - lo);
            for (int x = lane; x < width; x += 32) slice[x] = 0.0f;
            __syncwarp();
            const bool cut = (g == gj);
            for (int base = e0; base < e1; base += 32) {
                const int e = base + lane;
                int a = 0, b = 0;
                float y = 0.f;
                if (e < e1) {
                    const int u = cidx[e];
                    y = cval[e];
                    a = hseg[(size_t)u * R1 + g];
                    b = cut ? cpos[e] + 1 : hseg[(size_t)u * R1 + g + 1];
                }
                // Raters are applied one after the other (items of one rater are distinct, so the lanes never
                // collide in the slice).  Memory latency is hidden by fetching the first 128 entries of the
                // NEXT rater's segment (4 per lane) before the current one is applied.
                unsigned mask = __ballot_sync(0xffffffffu, b > a);
                int nx[4];
                float nv[4];
                int n_aa = 0, n_bb = 0;
                float n_yy = 0.f;
                auto fetch = [&](int l) {
                    n_yy = __shfl_sync(0xffffffffu, y, l);
                    n_aa = __shfl_sync(0xffffffffu, a, l);
                    n_bb = __shfl_sync(0xffffffffu, b, l);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int p = n_aa + lane + 32 * k;
                        nx[k] = -1; nv[k] = 0.f;
                        if (ADAPT && k > 0 && n_aa + 32 * k >= n_bb) continue;   // warp-uniform: nothing in this batch
                        if (p < n_bb) {
                            if constexpr (PACKED) { const int2 en = pent[p]; nx[k] = en.x; nv[k] = __int_as_float(en.y); }
                            else { nx[k] = pidx[p] - lo; nv[k] = pval[p]; }
                        }
                    }
                };
                if (mask) { fetch(__ffs(mask) - 1); mask &= mask - 1; }
                else continue;
                for (;;) {
                    int cx[4];
                    float cv[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) { cx[k] = nx[k]; cv[k] = nv[k]; }
                    const int aa = n_aa, bb = n_bb;
                    const float yy = n_yy;
                    const bool more = mask != 0u;
                    if (more) { fetch(__ffs(mask) - 1); mask &= mask - 1; }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (ADAPT && k > 0 && aa + 32 * k >= bb) continue;       // warp-uniform
                        if (cx[k] >= 0) slice[cx[k]] = __fadd_rn(slice[cx[k]], __fmul_rn(yy, cv[k]));
                    }
                    for (int p = aa + 128 + lane; p < bb; p += 32) {
                        if constexpr (PACKED) {
                            const int2 en = pent[p];
                            slice[en.x] = __fadd_rn(slice[en.x], __fmul_rn(yy, __int_as_float(en.y)));
                        } else {
                            const int x = pidx[p] - lo;
                            slice[x] = __fadd_rn(slice[x], __fmul_rn(yy, pval[p]));
                        }
                    }
                    __syncwarp();
                    if (!more) break;
                }
            }
            if (c1 - c0 <= G3_CHUNK) {
                for (int x = lane; x < width; x += 32) g_row[lo + x] = slice[x];
            } else {
                for (int x = lane; x < width; x += 32) { const float v = slice[x]; if (v != 0.0f) atomicAdd(g_row + lo + x, v); }
            }
            __syncwarp();
        } else {
            // tail: ranks in [R*RW, jp], short segments -> global fp32 RED
            for (int base = e0; base < e1; base += 32) {
                const int e = base + lane;
                int a = 0, b = 0;
                float y = 0.f;
                if (e < e1) {
                    const int u = cidx[e];
                    y = cval[e];
                    a = hseg[(size_t)u * R1 + R];
                    b = cpos[e] + 1;
                }
                // sub-warp groups of 8 lanes walk 4 entries at a time
                for (int l0 = 0; l0 < 32; l0 += 4) {
                    const int src = l0 + (lane >> 3);
                    const float yy = __shfl_sync(0xffffffffu, y, src);
                    const int aa = __shfl_sync(0xffffffffu, a, src);
                    const int bb = __shfl_sync(0xffffffffu, b, src);
                    for (int p = aa + (lane & 7); p < bb; p += 8) {
                        if constexpr (PACKED) {
                            const int2 en = pent[p];   // index relative to the tail range
                            atomicAdd(g_row + (size_t)R * RW + en.x, __fmul_rn(yy, __int_as_float(en.y)));
                        } else atomicAdd(g_row + pidx[p], __fmul_rn(yy, pval[p]));
                    }
                }
            }
        }
    }
}

// upper triangle <- transpose of the lower triangle: 64 x 64 tiles through shared memory, one CTA per tile of the
// triangle (linear tile index -> (by, bx), no empty CTAs), sixteen loads in flight per thread before the barrier
__global__ void __launch_bounds__(256) gram_mirror_kernel(float *__restrict__ G, int n, int64_t ld) {
    __shared__ float tile[64][65];
    const unsigned t = blockIdx.x;
    int by = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((unsigned)(by + 1) * (unsigned)(by + 2) / 2u <= t) ++by;
    while ((unsigned)by * (unsigned)(by + 1) / 2u > t) --by;
    const int bx = (int)(t - (unsigned)by * (unsigned)(by + 1) / 2u);      // bx <= by
    const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int row = by * 64 + ty + 4 * i, col = bx * 64 + tx;
        v[i] = (row < n && col < n) ? G[(size_t)row * ld + col] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) tile[ty + 4 * i][tx] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int r = ty + 4 * i;
        const int row = bx * 64 + r, col = by * 64 + tx;  // destination (upper)
        if (row < n && col < n && col > row) G[(size_t)row * ld + col] = tile[tx][r];
    }
}

// Multi-GPU: fused slab exchange + mirror over peer memory.  Rank `me` holds its own row slab of the lower
// triangle; the slabs of the other ranks are read straight out of their memory (CUDA IPC mappings, NVLink
// P2P loads) tile by tile while the tile is being mirrored, so every element of the triangle crosses
// NVLink once per destination and there is no separate all-gather pass: the lower tile is stored locally
// (if it came from a peer) together with its transpose in the upper triangle.
struct GramPeers {
    const float *src[RT_MAX_PEERS];
    int cuts[RT_MAX_PEERS + 1];
    int n_parts;
    int me;
};

__global__ void __launch_bounds__(256) gram_pull_mirror_kernel(GramPeers P, float *G, int n, int64_t ld) {
    __shared__ float tile[32][33];
    const int bx = blockIdx.x, by = blockIdx.y;
    if (bx > by) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int row = by * 32 + r, col = bx * 32 + tx;
        float v = 0.0f;
        if (row < n && col <= row) {
            int owner = 0;
#pragma unroll
            for (int p = 1; p < RT_MAX_PEERS; ++p) owner += (p < P.n_parts && row >= P.cuts[p]) ? 1 : 0;
            v = P.src[owner][(size_t)row * ld + col];
            if (owner != P.me) G[(size_t)row * ld + col] = v;
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int row = bx * 32 + r, col = by * 32 + tx;  // destination (upper)
        if (row < n && col < n && col > row) G[(size_t)row * ld + col] = tile[tx][r];
    }
}

// Same with 64 x 64 tiles and 16-byte accesses (needs ld % 4 == 0 and 16-byte aligned slabs): NVLink P2P
// reads reach their bandwidth only with wide loads and many bytes in flight (4 x LDG.128 per thread).
//
// Two phases keep every GPU's NVLink egress balanced although the slabs are not (slabs are balanced by
// multiply-adds, and the slab of the unpopular tail holds ~45 % of the triangle's bytes at N = 8: if every
// rank pulled it from its owner, that one GPU would have to send it seven times).  Every tile has a
// "stripe" rank (a hash of its coordinates).  Phase 0: a rank processes the tiles of its own stripe and the
// tiles touching its own slab, reading each row from its true owner -- the big slab leaves its owner only
// once, spread over all peers.  Phase 1 (after a node barrier): every remaining tile is read from its stripe
// rank, which has held it since phase 0.  Per-GPU egress: ~(N-1)/N of the triangle either way.
__device__ __forceinline__ int gram_tile_stripe(int by, int bx, int n_parts) { return (by * 5 + bx * 3) % n_parts; }

__global__ void __launch_bounds__(256) gram_pull_mirror_v4_kernel(GramPeers P, float *G, int n, int64_t ld, int phase) {
    __shared__ float tile[64][65];
    const int bx = blockIdx.x, by = blockIdx.y;
    if (bx > by) return;
    const int stripe = gram_tile_stripe(by, bx, P.n_parts);
    {
        // does this tile belong to phase 0 on this rank?  (own stripe, or rows of the own slab inside the tile)
        const int row_lo = by * 64, row_hi = min(by * 64 + 64, n);
        const bool mine = row_lo < P.cuts[P.me + 1] && row_hi > P.cuts[P.me];
        const bool first = (stripe == P.me) || mine;
        if ((phase == 0) != first) return;
    }
    const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;  // 16 threads cover the 64 columns of a row
    float4 v[4];
    int own[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int row = by * 64 + r0 + 16 * q, col = bx * 64 + c4;
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        own[q] = P.me;
        if (row < n && col <= row) {   // col <= row: the float4 starts inside the triangle (may run past the diagonal)
            int owner = 0;
#pragma unroll
            for (int p = 1; p < RT_MAX_PEERS; ++p) owner += (p < P.n_parts && row >= P.cuts[p]) ? 1 : 0;
            own[q] = owner;
            const int src = (phase == 0 || owner == P.me) ? owner : stripe;
            v[q] = *reinterpret_cast<const float4 *>(P.src[src] + (size_t)row * ld + col);
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + 16 * q;
        const int row = by * 64 + r, col = bx * 64 + c4;
        // entries right of the diagonal inside a float4 are not part of the lower triangle: the transposed store
        // below only uses entries with col < row, and only the in-triangle prefix is stored locally
        if (own[q] != P.me && row < n && col <= row) {
            float *dst = G + (size_t)row * ld + col;
            if (col + 3 <= row) *reinterpret_cast<float4 *>(dst) = v[q];
            else {
                dst[0] = v[q].x;
                if (col + 1 <= row) dst[1] = v[q].y;
                if (col + 2 <= row) dst[2] = v[q].z;
            }
        }
        tile[r][c4 + 0] = v[q].x; tile[r][c4 + 1] = v[q].y; tile[r][c4 + 2] = v[q].z; tile[r][c4 + 3] = v[q].w;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int row = bx * 64 + r0 + 16 * q;       // destination row (upper triangle)
        const int col = by * 64 + c4;                // destination columns col .. col+3
        if (row >= n) continue;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = tile[c4 + e][r0 + 16 * q];
        float *dst = G + (size_t)row * ld + col;
        if (col > row && col + 3 < n) *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (col + e > row && col + e < n) dst[e] = o[e];
        }
    }
}

// G[orig_of[jp]][x] = G'[jp][rank_of[x]]: one CTA per row; the row is staged in shared memory when it fits
// (orig_of == nullptr: rows keep their position -- the owner-rows layout of the multi-GPU fit, where only the columns go
// back to item ids and rows are addressed through rt_gram_row_slots)
__global__ void __launch_bounds__(1024) gram_unpermute_kernel(const float *__restrict__ Gp, int64_t ldp, int n_rows, int n,
                                                              const int *__restrict__ rank_of, const int *__restrict__ orig_of,
                                                              float *__restrict__ G, int64_t ld, int stage,
                                                              float *__restrict__ rowmax, int live_only, double live_above) {
    // rowmax (optional, staged mode with orig_of): largest off-diagonal entry of every row, by item id -- the solver uses it
    // to finish targets without a live coordinate without reading their Gram rows again.
    // live_only (with rowmax): a row whose off-diagonal maximum is not above live_above is a row the solver never reads
    // (its target is trivial, and it can be nobody's live coordinate: G is symmetric) -- only its diagonal entry is written.
    extern __shared__ __align__(16) float row_s[];
    __shared__ float s_red[32];
    __shared__ int s_skip;
    const int NT = blockDim.x;
    for (int jp = blockIdx.x; jp < n_rows; jp += gridDim.x) {
        const float *src = Gp + (size_t)jp * ldp;
        float *dst = G + (size_t)(orig_of ? orig_of[jp] : jp) * ld;
        if (stage) {
            __syncthreads();
            float mx = -INFINITY;
#pragma unroll 4
            for (int x = threadIdx.x; x < n; x += NT) { const float v = src[x]; row_s[x] = v; if (x != jp) mx = fmaxf(mx, v); }
            if (rowmax && orig_of) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
                if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mx;
            }
            __syncthreads();
            if (rowmax && orig_of && threadIdx.x < 32) {
                float m2 = threadIdx.x < (NT >> 5) ? s_red[threadIdx.x] : -INFINITY;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) m2 = fmaxf(m2, __shfl_xor_sync(0xffffffffu, m2, o));
                if (threadIdx.x == 0) {
                    rowmax[orig_of[jp]] = m2;
                    s_skip = live_only && !((double)m2 > live_above);
                }
            }
            if (live_only && rowmax && orig_of) {
                __syncthreads();
                if (s_skip) {
                    if (threadIdx.x == 0) dst[orig_of[jp]] = row_s[jp];
                    continue;
                }
            }
#pragma unroll 4
            for (int x = threadIdx.x; x < n; x += NT) dst[x] = row_s[rank_of[x]];
        } else {
#pragma unroll 4
            for (int x = threadIdx.x; x < n; x += NT) dst[x] = src[rank_of[x]];
        }
    }
}

// ---- multi-GPU, owner-rows layout -------------------------------------------------------------------------------
// Block-cyclic ownership: 64-row block b of the rank-space matrix belongs to part b % n_parts and sits at local block
// b / n_parts of that part's slab.  After rt_gram_lower_blocks a slab holds the lower-triangle part of its rows (columns
// <= row).  The rest of a row is the transposed column of the triangle below it, spread over every part: this kernel
// reads those 64 x 64 tiles straight out of their owners' slabs (CUDA IPC mappings, NVLink P2P loads for peers) and
// stores them transposed into the local rows.  A part only ever pulls the columns of ITS rows: (N-1)/N^2 of the matrix
// per GPU instead of the whole triangle, egress and ingress balanced by construction, one phase.  Writers touch
// columns > row of their own rows, readers columns <= row: no element is read and written in the same pass.
struct GramRowPeers {
    const float *src[RT_MAX_PEERS];
    int n_parts;
    int me;
};

__global__ void __launch_bounds__(256) gram_pull_cols_kernel(GramRowPeers P, float *local, int n, int64_t ld) {
    __shared__ float tile[64][65];
    const int bx = blockIdx.x;                  // source row block (global index) = destination column block
    const int lb = blockIdx.y;                  // local block of the destination rows
    const int bq = lb * P.n_parts + P.me;       // its global index
    if (bx < bq || bq * 64 >= n) return;
    const float *src = P.src[bx % P.n_parts] + (size_t)(bx / P.n_parts) * 64 * ld;
    const int c4 = (threadIdx.x & 15) * 4, r0 = threadIdx.x >> 4;  // 16 threads cover the 64 columns of a row
    float4 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + 16 * q;
        const int row = bx * 64 + r, col = bq * 64 + c4;
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        // col <= row: the float4 starts inside the triangle (on the diagonal tile it may run past the diagonal; those
        // entries are never used by the transposed store below)
        if (row < n && col <= row) v[q] = *reinterpret_cast<const float4 *>(src + (size_t)r * ld + col);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = r0 + 16 * q;
        tile[r][c4 + 0] = v[q].x; tile[r][c4 + 1] = v[q].y; tile[r][c4 + 2] = v[q].z; tile[r][c4 + 3] = v[q].w;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int lr = r0 + 16 * q;
        const int row = bq * 64 + lr;               // destination row (global); local row lb * 64 + lr
        const int col = bx * 64 + c4;               // destination columns col .. col + 3
        if (row >= n) continue;
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = tile[c4 + e][lr];
        float *dst = local + (size_t)(lb * 64 + lr) * ld + col;
        if (col > row && col + 3 < n) *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        else {
#pragma unroll
            for (int e = 0; e < 4; ++e) if (col + e > row && col + e < n) dst[e] = o[e];
        }
    }
}

// slot of item i = (owner part << 24) | local row in that part's buffer
__global__ void gram_row_slots_kernel(const int *__restrict__ rank_of, int n_items, int n_parts, int *__restrict__ slots) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const int jp = rank_of[i];
    slots[i] = (((jp >> 6) % n_parts) << 24) | blk_local_row(jp, n_parts);
}

static int bits_for_n(long long n) { int b = 1; while ((1ll << b) < n && b < 32) ++b; return b; }

}  // namespace rt

using namespace rt;

#define G3_CUB(call_expr)                                                                          \
    do {                                                                                           \
        size_t tmp_bytes__ = 0;                                                                    \
        void *d_tmp__ = nullptr;                                                                   \
        RT_CUDA(call_expr);                                                                        \
        d_tmp__ = rt::scratch(SCR_CUB, tmp_bytes__);                                               \
        if (!d_tmp__) return RT_ERR_CUDA;                                                          \
        RT_CUDA(call_expr);                                                                        \
        rt::count_launch(2);                                                                       \
    } while (0)

static int g_last_head = 0;
extern "C" int32_t rt_gram_last_head(void) { return g_last_head; }

// block_mode = 0: part owns a contiguous row slab [h_cuts[part], h_cuts[part + 1]) balanced by multiply-adds, written
// at its global rows of an I-row buffer.  block_mode = 1: block-cyclic ownership of 64-row blocks, the buffer holds only
// the rows of this part (blk_local_row); no cost pass, no host synchronisation, h_cuts unused.
static int gram_lower_impl(int32_t n_users, int32_t n_items, const int32_t *d_cptr, const int32_t *d_cidx,
                           const float *d_cval, const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                           int64_t nnz, int32_t part, int32_t n_parts, float *d_Gp, int64_t ldgp, int32_t *d_rank_of,
                           int32_t *d_orig_of, int32_t *h_cuts, int block_mode, void *stream) {
    g_last_head = 0;
    RT_ARG(n_users > 0 && n_items > 0 && nnz >= 0 && nnz < (1ll << 31), "shape");
    RT_ARG(n_parts >= 1 && part >= 0 && part < n_parts, "part / n_parts");
    RT_ARG(d_cptr && d_rptr && d_Gp && ldgp >= n_items && d_rank_of && d_orig_of && (h_cuts || block_mode), "null pointer / ldgp");
    int32_t cuts_dummy[RT_MAX_PEERS + 2];
    if (block_mode) { RT_ARG(n_parts <= RT_MAX_PEERS, "n_parts"); h_cuts = cuts_dummy; }
    const int blk_parts = block_mode ? n_parts : 0;
    cudaStream_t st = (cudaStream_t)stream;
    const int bs = 256;
    const int I = n_items;
    // ---- geometry --------------------------------------------------------------------------------
    // tuning switches (rt_set_option): "gram_slice" = floats per warp slice (1152 / 1728 / 2304),
    // "gram_ranges" = number of shared-memory ranges R (0 = derive from the item count)
    int RW = rt::option(rt::OPT_GRAM_SLICE);
    if (RW != 1152 && RW != 2304) RW = G3_SLICE;
    int R = rt::option(rt::OPT_GRAM_RANGES);
    if (R <= 0) R = (I + 4 * RW - 1) / (4 * RW);   // head = a quarter of the catalogue, at most 4 ranges (measured: kbench)
    if (R < 1) R = 1;
    if (R > G3_MAX_R) R = G3_MAX_R;
    // ---- workspace -------------------------------------------------------------------------------
    Carver sizing(nullptr, (size_t)-1);
    auto plan = [&](Carver &c) {
        struct P { unsigned long long *rk, *rks, *keys, *keys2, *cost, *cost_s, *counter; int *pidx, *cpos, *hseg, *n_chunks, *chunk_start; float *pval; } p;
        p.rk = c.take<unsigned long long>(I); p.rks = c.take<unsigned long long>(I);
        p.keys = c.take<unsigned long long>(nnz + 1); p.keys2 = c.take<unsigned long long>(nnz + 1);
        p.cost = c.take<unsigned long long>(I + 1); p.cost_s = c.take<unsigned long long>(I + 1);
        p.counter = c.take<unsigned long long>(4);
        p.pidx = c.take<int>(nnz + 1); p.cpos = c.take<int>(nnz + 1); p.hseg = c.take<int>((size_t)n_users * (R + 1));
        p.n_chunks = c.take<int>(I + 1); p.chunk_start = c.take<int>(I + 2);
        p.pval = c.take<float>(nnz + 1);
        return p;
    };
    plan(sizing);
    void *base = rt::scratch(SCR_STORE_B, sizing.off + 1024);
    if (!base) return RT_ERR_CUDA;
    Carver real(base, sizing.off + 1024);
    auto P = plan(real);
    // ---- popularity rank -------------------------------------------------------------------------
    rank_keys_kernel<<<(I + bs - 1) / bs, bs, 0, st>>>(d_cptr, I, P.rk);
    RT_CHECK_LAUNCH();
    G3_CUB(cub::DeviceRadixSort::SortKeys(d_tmp__, tmp_bytes__, P.rk, P.rks, I, 0, 64, st));
    rank_scatter_kernel<<<(I + bs - 1) / bs, bs, 0, st>>>(P.rks, I, d_orig_of, d_rank_of);
    RT_CHECK_LAUNCH();
    for (int p = 0; p <= n_parts; ++p) h_cuts[p] = p == 0 ? 0 : I;
    if (nnz == 0) { RT_CUDA(cudaStreamSynchronize(st)); return RT_OK; }
    RT_ARG(d_cidx && d_cval && d_ridx && d_rval, "null pointer");
    // ---- rank-sorted CSR ---------------------------------------------------------------------------
    const int nwords = (I + 31) >> 5;
    const size_t bm_smem = (size_t)nwords * 8;
    RT_CUDA(cudaMemsetAsync(P.counter, 0, 4 * sizeof(unsigned long long), st));   // ([0]: task cursor of the lower-triangle kernel)
    if (bm_smem <= (size_t)rt::smem_optin() - 1024 && rt::option(rt::OPT_GRAM_IMPL) != 1) {
        // segmented sort: every row by the rank of its items (registers for rows <= 64, bitmap counting sort for the rest)
        row_sort_small_kernel<<<(unsigned)(((int64_t)n_users * 32 + bs - 1) / bs), bs, 0, st>>>(n_users, d_rptr, d_ridx, d_rval,
                                                                                            d_rank_of, P.pidx, P.pval);
        RT_CHECK_LAUNCH();
        int *d_next = (int *)P.counter + 4;
        const int team = nwords <= 2048 ? 32 : nwords <= 16384 ? 128 : 256;
        RT_CUDA(cudaFuncSetAttribute(row_sort_bitmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bm_smem));
        int per_sm = 1;
        RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, row_sort_bitmap_kernel, team, bm_smem));
        if (per_sm < 1) per_sm = 1;
        row_sort_bitmap_kernel<<<rt::sm_count() * per_sm, team, bm_smem, st>>>(n_users, I, d_rptr, d_ridx, d_rval, d_rank_of, P.pidx,
                                                                             P.pval, d_next);
        RT_CHECK_LAUNCH();
    } else {
        const unsigned grid = (unsigned)(((int64_t)n_users * 32 + bs - 1) / bs);
        relabel_keys_kernel<<<grid, bs, 0, st>>>(n_users, d_rptr, d_ridx, d_rank_of, P.keys);
        RT_CHECK_LAUNCH();
        // two stable sorts over the populated bit ranges (rank bits, then user bits) instead of one over
        // [0, 32 + user bits): the zero bits in between would cost whole radix passes.  keys -> keys2 -> keys;
        // cpos (written later by entry_pos_kernel) holds the intermediate values.
        const int rank_bits = bits_for_n(n_items), user_bits = bits_for_n(n_users);
        float *mid_v = (float *)P.cpos;
        G3_CUB(cub::DeviceRadixSort::SortPairs(d_tmp__, tmp_bytes__, P.keys, P.keys2, d_rval, mid_v, (int)nnz, 0, rank_bits, st));
        G3_CUB(cub::DeviceRadixSort::SortPairs(d_tmp__, tmp_bytes__, P.keys2, P.keys, mid_v, P.pval, (int)nnz, 32, 32 + user_bits, st));
        { unsigned long long *t_ = P.keys; P.keys = P.keys2; P.keys2 = t_; }
        low32_kernel<<<(unsigned)((nnz + bs - 1) / bs), bs, 0, st>>>(P.keys2, nnz, P.pidx);
        RT_CHECK_LAUNCH();
    }
    {
        int rc = rt_csr_split(n_users, d_rptr, P.pidx, 0, RW, R, P.hseg, stream);
        if (rc) return rc;
    }
    // ---- dense head on the tensor cores (single-GPU form): the sparse kernel then starts below it, and the entries of the
    // head columns (most of the matrix: the popular items) need no position
    int head = 0;
    if (n_parts == 1 && !block_mode && rt::option(rt::OPT_GRAM_HEAD) != 0 && rt::option(rt::OPT_GRAM_IMPL) != 1) {
        const int rc = rt::gram_head_tc(n_users, I, d_cptr, d_cidx, d_cval, nnz, d_rptr, P.pidx, d_orig_of, d_Gp, ldgp, &head, st);
        if (rc) return rc;
        g_last_head = head;
    }
    const bool by_cost = n_parts > 1 && !block_mode;
    if (by_cost) RT_CUDA(cudaMemsetAsync(P.cost, 0, sizeof(unsigned long long) * ((size_t)I + 1), st));
    entry_pos_kernel<<<(unsigned)((nnz + bs - 1) / bs), bs, 0, st>>>(I, nnz, d_rank_of, d_cptr, d_cidx, d_rptr, P.pidx, P.cpos,
                                                                    by_cost ? P.cost : nullptr, blk_parts, part, head);
    RT_CHECK_LAUNCH();
    chunk_count_kernel<<<(I + bs - 1) / bs, bs, 0, st>>>(I, d_orig_of, d_cptr, P.n_chunks, blk_parts, part);
    RT_CHECK_LAUNCH();
    RT_CUDA(cudaMemsetAsync(P.n_chunks + I, 0, sizeof(int), st));
    G3_CUB(cub::DeviceScan::ExclusiveSum(d_tmp__, tmp_bytes__, P.n_chunks, P.chunk_start, I + 1, st));
    // ---- partition of the rows by exact work -------------------------------------------------------
    int row_begin = 0, row_end = I;
    if (by_cost) {
        G3_CUB(cub::DeviceScan::InclusiveSum(d_tmp__, tmp_bytes__, P.cost, P.cost_s, I, st));
        std::vector<unsigned long long> cs((size_t)I);
        RT_CUDA(cudaMemcpyAsync(cs.data(), P.cost_s, sizeof(unsigned long long) * (size_t)I, cudaMemcpyDeviceToHost, st));
        RT_CUDA(cudaStreamSynchronize(st));
        const unsigned long long total = cs[I - 1];
        int pos = 0;
        for (int p = 1; p < n_parts; ++p) {
            const unsigned long long want = total / (unsigned long long)n_parts * (unsigned long long)p;
            while (pos < I && cs[pos] < want) ++pos;
            h_cuts[p] = pos;
        }
        h_cuts[n_parts] = I;
        row_begin = h_cuts[part];
        row_end = h_cuts[part + 1];
    }
    if (head > row_begin) row_begin = head < row_end ? head : row_end;
    // ---- lower triangle ----------------------------------------------------------------------------
    if (row_end > row_begin) {
        RT_CUDA(cudaMemsetAsync(P.counter, 0, sizeof(unsigned long long), st));
        const size_t smem = sizeof(float) * (size_t)G3_WARPS * RW;
        int per_sm = (int)((size_t)(rt::smem_optin() + 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 8) per_sm = 8;
        const int grid = rt::sm_count() * per_sm;
        const int mode = rt::option(rt::OPT_GRAM_ADAPT);
        int2 *pent = nullptr;
        if (mode == 2) {
            pent = (int2 *)rt::scratch(SCR_GRAM_PACK, sizeof(int2) * ((size_t)nnz + 64));
            if (!pent) return RT_ERR_CUDA;
            pack_entries_kernel<<<(unsigned)((nnz + bs - 1) / bs), bs, 0, st>>>(P.pidx, P.pval, nnz, RW, R, pent);
            RT_CHECK_LAUNCH();
        }
#define G3_LAUNCH(SL, MD)                                                                                               \
        do {                                                                                                            \
            RT_CUDA(cudaFuncSetAttribute(gram_lower_kernel<SL, MD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            gram_lower_kernel<SL, MD><<<grid, G3_WARPS * 32, smem, st>>>(row_begin, row_end, P.chunk_start + row_begin, R, RW, \
                                                                        d_orig_of, d_cptr, d_cidx, d_cval, P.cpos, P.hseg, \
                                                                        P.pidx, P.pval, d_Gp, ldgp, P.counter, blk_parts, \
                                                                        pent);                                           \
        } while (0)
#define G3_PICK(SL)                                                                                                     \
        do {                                                                                                            \
            if (mode == 2) G3_LAUNCH(SL, 2); else if (mode == 1) G3_LAUNCH(SL, 1); else G3_LAUNCH(SL, 0);               \
        } while (0)
        if (RW == 1152) G3_PICK(1152);
        else if (RW == 2304) G3_PICK(2304);
        else G3_PICK(G3_SLICE);
#undef G3_PICK
#undef G3_LAUNCH
        RT_CHECK_LAUNCH();
    }
    return RT_OK;
}

extern "C" int rt_gram_lower(int32_t n_users, int32_t n_items, const int32_t *d_cptr, const int32_t *d_cidx,
                             const float *d_cval, const int32_t *d_rptr, const int32_t *d_ridx, const float *d_rval,
                             int64_t nnz, int32_t part, int32_t n_parts, float *d_Gp, int64_t ldgp, int32_t *d_rank_of,
                             int32_t *d_orig_of, int32_t *h_cuts, void *stream) {
    return gram_lower_impl(n_users, n_items, d_cptr, d_cidx, d_cval, d_rptr, d_ridx, d_rval, nnz, part, n_parts, d_Gp, ldgp,
                           d_rank_of, d_orig_of, h_cuts, 0, stream);
}

extern "C" int rt_gram_lower_blocks(int32_t n_users, int32_t n_items, const int32_t *d_cptr, const int32_t *d_cidx,
                                    const float *d_cval, const int32_t *d_rptr, const int32_t *d_ridx,
                                    const float *d_rval, int64_t nnz, int32_t part, int32_t n_parts, float *d_slab,
                                    int64_t ldgp, int32_t *d_rank_of, int32_t *d_orig_of, void *stream) {
    return gram_lower_impl(n_users, n_items, d_cptr, d_cidx, d_cval, d_rptr, d_ridx, d_rval, nnz, part, n_parts, d_slab, ldgp,
                           d_rank_of, d_orig_of, nullptr, 1, stream);
}

static int launch_unpermute(int32_t n_items, const float *d_Gp, int64_t ldgp, const int32_t *d_rank_of,
                            const int32_t *d_orig_of, float *d_G, int64_t ldg, cudaStream_t st, float *d_rowmax = nullptr,
                            int32_t *h_has_rowmax = nullptr, int live_only = 0, double live_above = 0.0) {
    const size_t row_bytes = sizeof(float) * (size_t)n_items;
    const int stage = row_bytes + 2048 <= (size_t)rt::smem_optin() ? 1 : 0;
    const size_t smem = stage ? row_bytes : 0;
    RT_CUDA(cudaFuncSetAttribute(gram_unpermute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = stage ? (int)((size_t)(rt::smem_optin() + 1024) / (smem + 1024)) : 2;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    int grid = rt::sm_count() * per_sm;
    if (grid > n_items) grid = n_items;
    gram_unpermute_kernel<<<grid, 1024, smem, st>>>(d_Gp, ldgp, n_items, n_items, d_rank_of, d_orig_of, d_G, ldg, stage,
                                                    stage ? d_rowmax : nullptr, (stage && d_rowmax) ? live_only : 0, live_above);
    RT_CHECK_LAUNCH();
    if (h_has_rowmax) *h_has_rowmax = (stage && d_rowmax) ? 1 : 0;
    return RT_OK;
}

extern "C" int rt_gram_finish_live(int32_t n_items, float *d_Gp, int64_t ldgp, const int32_t *d_rank_of,
                                   const int32_t *d_orig_of, float *d_G, int64_t ldg, float *d_rowmax,
                                   const rt_fit_config *cfg, int32_t *h_has_rowmax, int32_t *h_live_only, void *stream) {
    RT_ARG(n_items > 0 && d_Gp && d_G && d_rank_of && d_orig_of && ldgp >= n_items && ldg >= n_items && d_rowmax && h_has_rowmax && cfg && h_live_only, "arguments");
    RT_ARG(d_Gp != d_G, "rt_gram_finish is not in-place");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned nt = (unsigned)((n_items + 63) / 64);
    gram_mirror_kernel<<<nt * (nt + 1) / 2, 256, 0, st>>>(d_Gp, n_items, ldgp);
    RT_CHECK_LAUNCH();
    // the same conditions and the same threshold as the solver's trivial-target rule (solve.cu: A.skip_trivial, A.a)
    const int live_only = (cfg->nn > 0 && cfg->skip_trivial && cfg->positive && cfg->nonneg) ? 1 : 0;
    const double a = (double)(float)(cfg->alpha * cfg->l1_ratio * (double)cfg->n_samples);
    const int rc = launch_unpermute(n_items, d_Gp, ldgp, d_rank_of, d_orig_of, d_G, ldg, st, d_rowmax, h_has_rowmax, live_only, a);
    *h_live_only = (rc == RT_OK && live_only && *h_has_rowmax) ? 1 : 0;
    return rc;
}

extern "C" int rt_gram_block_rows(int32_t n_items, int32_t n_parts, int32_t part, int32_t *h_rows_alloc, int32_t *h_rows_own) {
    RT_ARG(n_items > 0 && n_parts >= 1 && n_parts <= RT_MAX_PEERS && part >= 0 && part < n_parts, "arguments");
    const int nt = (n_items + 63) / 64;
    if (h_rows_alloc) *h_rows_alloc = (nt + n_parts - 1) / n_parts * 64;
    if (h_rows_own) {
        int own = 0;
        for (int b = part; b < nt; b += n_parts) own += std::min(64, n_items - b * 64);
        *h_rows_own = own;
    }
    return RT_OK;
}

extern "C" int rt_gram_pull_cols(int32_t n_items, const void *const *h_slabs, int32_t n_parts, int32_t part, int64_t ldgp,
                                 void *stream) {
    RT_ARG(n_items > 0 && h_slabs && ldgp >= n_items && (ldgp % 4) == 0, "arguments (ldgp must be a multiple of 4)");
    RT_ARG(n_parts >= 1 && n_parts <= RT_MAX_PEERS && part >= 0 && part < n_parts, "part / n_parts");
    GramRowPeers P;
    for (int p = 0; p < RT_MAX_PEERS; ++p) {
        P.src[p] = (const float *)h_slabs[p < n_parts ? p : 0];
        RT_ARG(P.src[p] != nullptr && (((uintptr_t)P.src[p]) & 15) == 0, "slab pointers must be 16-byte aligned");
    }
    P.n_parts = n_parts; P.me = part;
    const int nt = (n_items + 63) / 64;
    const int n_local = (nt - part + n_parts - 1) / n_parts;
    if (n_local <= 0) return RT_OK;
    gram_pull_cols_kernel<<<dim3(nt, n_local), 256, 0, (cudaStream_t)stream>>>(P, (float *)h_slabs[part], n_items, ldgp);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_gram_unpermute_rows(int32_t n_rows, int32_t n_items, const float *d_slab, int64_t ldgp,
                                      const int32_t *d_rank_of, float *d_rows, int64_t ldg, void *stream) {
    RT_ARG(n_rows >= 0 && n_items > 0 && d_slab && d_rank_of && d_rows && ldgp >= n_items && ldg >= n_items && d_rows != d_slab,
           "arguments");
    if (n_rows == 0) return RT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t row_bytes = sizeof(float) * (size_t)n_items;
    const int stage = row_bytes + 2048 <= (size_t)rt::smem_optin() ? 1 : 0;
    const size_t smem = stage ? row_bytes : 0;
    RT_CUDA(cudaFuncSetAttribute(gram_unpermute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = stage ? (int)((size_t)(rt::smem_optin() + 1024) / (smem + 1024)) : 2;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    int grid = rt::sm_count() * per_sm;
    if (grid > n_rows) grid = n_rows;
    gram_unpermute_kernel<<<grid, 1024, smem, st>>>(d_slab, ldgp, n_rows, n_items, d_rank_of, nullptr, d_rows, ldg, stage, nullptr, 0, 0.0);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_gram_row_slots(int32_t n_items, const int32_t *d_rank_of, int32_t n_parts, int32_t *d_slots, void *stream) {
    RT_ARG(n_items > 0 && n_items < (1 << 24) && d_rank_of && d_slots && n_parts >= 1 && n_parts <= RT_MAX_PEERS, "arguments");
    gram_row_slots_kernel<<<(n_items + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_rank_of, n_items, n_parts, d_slots);
    RT_CHECK_LAUNCH();
    return RT_OK;
}

extern "C" int rt_gram_finish_p2p(int32_t n_items, const void *const *h_slabs, int32_t n_parts, int32_t part,
                                  const int32_t *h_cuts, int64_t ldgp, const int32_t *d_rank_of,
                                  const int32_t *d_orig_of, float *d_G, int64_t ldg, int32_t phase, void *stream) {
    RT_ARG(n_items > 0 && h_slabs && h_cuts && ldgp >= n_items, "arguments");
    RT_ARG(n_parts >= 1 && n_parts <= RT_MAX_PEERS && part >= 0 && part < n_parts, "part / n_parts");
    RT_ARG(phase >= 0 && phase <= 2, "phase");
    for (int p = 0; p < n_parts; ++p) RT_ARG(h_slabs[p] != nullptr && h_cuts[p] <= h_cuts[p + 1], "slab pointers / cuts");
    RT_ARG(h_cuts[0] == 0 && h_cuts[n_parts] == n_items, "cuts must cover [0, n_items]");
    cudaStream_t st = (cudaStream_t)stream;
    GramPeers P;
    for (int p = 0; p < RT_MAX_PEERS; ++p) { P.src[p] = (const float *)h_slabs[p < n_parts ? p : 0]; P.cuts[p] = p <= n_parts ? h_cuts[p] : n_items; }
    P.cuts[RT_MAX_PEERS] = n_items;
    P.n_parts = n_parts; P.me = part;
    float *local = (float *)h_slabs[part];
    bool wide = (ldgp % 4) == 0;
    for (int p = 0; p < n_parts; ++p) wide = wide && ((((uintptr_t)h_slabs[p]) & 15) == 0);
    if (phase <= 1) {
        if (wide) {
            const int nt = (n_items + 63) / 64;
            gram_pull_mirror_v4_kernel<<<dim3(nt, nt), 256, 0, st>>>(P, local, n_items, ldgp, phase);
            RT_CHECK_LAUNCH();
        } else if (phase == 0) {
            // narrow fallback: one phase, every row straight from its owner
            const int nt = (n_items + 31) / 32;
            gram_pull_mirror_kernel<<<dim3(nt, nt), 256, 0, st>>>(P, local, n_items, ldgp);
            RT_CHECK_LAUNCH();
        }
        return RT_OK;
    }
    RT_ARG(d_rank_of && d_orig_of && d_G && ldg >= n_items && d_G != local, "unpermute arguments");
    return launch_unpermute(n_items, local, ldgp, d_rank_of, d_orig_of, d_G, ldg, st);
}

extern "C" int rt_gram_finish(int32_t n_items, float *d_Gp, int64_t ldgp, const int32_t *d_rank_of,
                              const int32_t *d_orig_of, float *d_G, int64_t ldg, void *stream) {
    RT_ARG(n_items > 0 && d_Gp && d_G && d_rank_of && d_orig_of && ldgp >= n_items && ldg >= n_items, "arguments");
    RT_ARG(d_Gp != d_G, "rt_gram_finish is not in-place");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned nt = (unsigned)((n_items + 63) / 64);
    gram_mirror_kernel<<<nt * (nt + 1) / 2, 256, 0, st>>>(d_Gp, n_items, ldgp);
    RT_CHECK_LAUNCH();
    return launch_unpermute(n_items, d_Gp, ldgp, d_rank_of, d_orig_of, d_G, ldg, st);
}

extern "C" int rt_gram_finish_rowmax(int32_t n_items, float *d_Gp, int64_t ldgp, const int32_t *d_rank_of,
                                     const int32_t *d_orig_of, float *d_G, int64_t ldg, float *d_rowmax,
                                     int32_t *h_has_rowmax, void *stream) {
    RT_ARG(n_items > 0 && d_Gp && d_G && d_rank_of && d_orig_of && ldgp >= n_items && ldg >= n_items && d_rowmax && h_has_rowmax, "arguments");
    RT_ARG(d_Gp != d_G, "rt_gram_finish is not in-place");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned nt = (unsigned)((n_items + 63) / 64);
    gram_mirror_kernel<<<nt * (nt + 1) / 2, 256, 0, st>>>(d_Gp, n_items, ldgp);
    RT_CHECK_LAUNCH();
    return launch_unpermute(n_items, d_Gp, ldgp, d_rank_of, d_orig_of, d_G, ldg, st, d_rowmax, h_has_rowmax);
}
