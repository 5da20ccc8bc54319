// upload.cu -- host -> device event upload: multi-threaded narrow + stage + DMA pipeline.
//
// Replaces the per-event Python loop that feeds `add_interaction`
// (/root/reference/rtrec/recommender.py:203-223 generate_batches -> base.py:72-94 -> interactions.py:81-119)
// for callers that hold a batch as four host columns (what a pandas DataFrame gives: int64 user ids,
// int64 item ids, float64 timestamps, float64 ratings).  Those columns live in pageable memory; a
// plain cudaMemcpy of pageable memory is staged by the driver on one thread.  Here a small pool of
// host threads each converts a chunk (ids int64 -> int32, which also takes 8 bytes per event off the
// PCIe bus) into its own pinned staging slot, issues the DMA on its own copy stream and moves on to
// the next chunk while the DMA runs.  The id ranges and the largest timestamp (interactions.py:92-99,
// 118-119) are reduced on the way, so no extra device pass is needed for validation.
//
// Staging memory is owned by the library (allocated on first use, freed by rt_release_scratch()).
#include <atomic>
#include <limits>
#include <thread>
#include <vector>

#include <limits.h>
#include <string.h>

#include "common.cuh"

namespace rt {

constexpr int64_t UP_CHUNK = 1 << 18;   // events per chunk: 6 MB staged
constexpr int UP_MAX_THREADS = 16;
constexpr size_t UP_SLOT_BYTES = (size_t)UP_CHUNK * 24;

struct UploadPool {
    int device = -1;
    int n_threads = 0;
    char *pinned = nullptr;  // n_threads * 2 slots
    cudaStream_t streams[UP_MAX_THREADS] = {};
    cudaEvent_t events[UP_MAX_THREADS][2] = {};
};
static UploadPool g_up;

static void upload_pool_free() {
    if (g_up.pinned) cudaFreeHost(g_up.pinned);
    for (int t = 0; t < g_up.n_threads; ++t) {
        if (g_up.streams[t]) cudaStreamDestroy(g_up.streams[t]);
        for (int s = 0; s < 2; ++s) if (g_up.events[t][s]) cudaEventDestroy(g_up.events[t][s]);
    }
    g_up = UploadPool();
}

void upload_release() { upload_pool_free(); }

static int upload_pool_get(int n_threads) {
    int dev = 0;
    RT_CUDA(cudaGetDevice(&dev));
    if (g_up.device == dev && g_up.n_threads >= n_threads) return RT_OK;
    upload_pool_free();
    RT_CUDA(cudaHostAlloc((void **)&g_up.pinned, UP_SLOT_BYTES * 2 * (size_t)n_threads, cudaHostAllocDefault));
    for (int t = 0; t < n_threads; ++t) {
        RT_CUDA(cudaStreamCreateWithFlags(&g_up.streams[t], cudaStreamNonBlocking));
        for (int s = 0; s < 2; ++s) RT_CUDA(cudaEventCreateWithFlags(&g_up.events[t][s], cudaEventDisableTiming));
    }
    g_up.device = dev;
    g_up.n_threads = n_threads;
    return RT_OK;
}

struct UploadJob {
    const int64_t *h_users, *h_items;
    const double *h_ts, *h_delta;
    int64_t n;
    int32_t *d_users, *d_items;
    double *d_ts, *d_delta;
    std::atomic<int64_t> next{0};
    std::atomic<int> failed{0};
    int device;
};

struct UploadStats {
    int64_t umin = LLONG_MAX, umax = LLONG_MIN, imin = LLONG_MAX, imax = LLONG_MIN;
    double tmax = -std::numeric_limits<double>::infinity();
};

static void upload_worker(UploadJob *job, int tid, UploadStats *out) {
    if (cudaSetDevice(job->device) != cudaSuccess) { job->failed.store(1); return; }
    cudaStream_t st = g_up.streams[tid];
    UploadStats s;
    const int64_t n_chunks = (job->n + UP_CHUNK - 1) / UP_CHUNK;
    int turn = 0;
    bool used[2] = {false, false};
    for (;;) {
        const int64_t c = job->next.fetch_add(1);
        if (c >= n_chunks || job->failed.load()) break;
        const int64_t a = c * UP_CHUNK, m = (a + UP_CHUNK <= job->n ? UP_CHUNK : job->n - a);
        char *slot = g_up.pinned + UP_SLOT_BYTES * (size_t)(tid * 2 + turn);
        if (used[turn] && cudaEventSynchronize(g_up.events[tid][turn]) != cudaSuccess) { job->failed.store(1); break; }
        int32_t *su = (int32_t *)slot, *si = su + UP_CHUNK;
        double *sts = (double *)(slot + (size_t)UP_CHUNK * 8), *sd = sts + UP_CHUNK;
        const int64_t *hu = job->h_users + a, *hi = job->h_items + a;
        int64_t umin = s.umin, umax = s.umax, imin = s.imin, imax = s.imax;
        for (int64_t k = 0; k < m; ++k) {
            const int64_t u = hu[k], i = hi[k];
            umin = u < umin ? u : umin; umax = u > umax ? u : umax;
            imin = i < imin ? i : imin; imax = i > imax ? i : imax;
            su[k] = (int32_t)u; si[k] = (int32_t)i;
        }
        s.umin = umin; s.umax = umax; s.imin = imin; s.imax = imax;
        const double *hts = job->h_ts + a;
        double tmax = s.tmax;
        for (int64_t k = 0; k < m; ++k) { const double t = hts[k]; sts[k] = t; tmax = t > tmax ? t : tmax; }
        s.tmax = tmax;
        memcpy(sd, job->h_delta + a, sizeof(double) * (size_t)m);
        bool ok = cudaMemcpyAsync(job->d_users + a, su, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(job->d_items + a, si, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(job->d_ts + a, sts, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok = ok && cudaMemcpyAsync(job->d_delta + a, sd, sizeof(double) * (size_t)m, cudaMemcpyHostToDevice, st) == cudaSuccess;
        ok = ok && cudaEventRecord(g_up.events[tid][turn], st) == cudaSuccess;
        if (!ok) { job->failed.store(1); break; }
        used[turn] = true;
        turn ^= 1;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) job->failed.store(1);
    *out = s;
}

}  // namespace rt

using namespace rt;

extern "C" int rt_upload_events(const int64_t *h_users, const int64_t *h_items, const double *h_ts,
                                const double *h_delta, int64_t n, int32_t *d_users, int32_t *d_items, double *d_ts,
                                double *d_delta, int64_t *h_min_user, int64_t *h_max_user, int64_t *h_min_item,
                                int64_t *h_max_item, double *h_max_ts, int32_t n_threads, void *stream) {
    RT_ARG(n > 0, "n");
    RT_ARG(h_users && h_items && h_ts && h_delta && d_users && d_items && d_ts && d_delta, "event arrays");
    RT_ARG(h_min_user && h_max_user && h_min_item && h_max_item && h_max_ts, "host outputs");
    int T = n_threads;
    if (T <= 0) {
        const unsigned hc = std::thread::hardware_concurrency();
        T = hc >= 4 ? (int)(hc * 3 / 4) : 1;  // the pipeline is bound by host memory copies, not by PCIe
    }
    if (T > UP_MAX_THREADS) T = UP_MAX_THREADS;
    const int64_t n_chunks = (n + UP_CHUNK - 1) / UP_CHUNK;
    if ((int64_t)T > n_chunks) T = (int)n_chunks;
    int rc = upload_pool_get(T);
    if (rc) return rc;
    // the destination buffers may have been handed out by a stream-ordered allocator: everything queued on the
    // caller's stream must have finished before another stream writes into them
    RT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    UploadJob job;
    job.h_users = h_users; job.h_items = h_items; job.h_ts = h_ts; job.h_delta = h_delta; job.n = n;
    job.d_users = d_users; job.d_items = d_items; job.d_ts = d_ts; job.d_delta = d_delta;
    RT_CUDA(cudaGetDevice(&job.device));
    std::vector<UploadStats> stats((size_t)T);
    std::vector<std::thread> threads;
    for (int t = 1; t < T; ++t) threads.emplace_back(upload_worker, &job, t, &stats[(size_t)t]);
    upload_worker(&job, 0, &stats[0]);
    for (auto &th : threads) th.join();
    if (job.failed.load()) {
        const cudaError_t e = cudaGetLastError();
        rt::set_error("rt_upload_events: staged copy failed: %s", cudaGetErrorString(e));
        return RT_ERR_CUDA;
    }
    UploadStats s;
    for (const auto &q : stats) {
        s.umin = q.umin < s.umin ? q.umin : s.umin; s.umax = q.umax > s.umax ? q.umax : s.umax;
        s.imin = q.imin < s.imin ? q.imin : s.imin; s.imax = q.imax > s.imax ? q.imax : s.imax;
        s.tmax = q.tmax > s.tmax ? q.tmax : s.tmax;
    }
    *h_min_user = s.umin; *h_max_user = s.umax; *h_min_item = s.imin; *h_max_item = s.imax; *h_max_ts = s.tmax;
    return RT_OK;
}
