// api.cu -- library-wide plumbing of librtrec_b200: error string, device info, launch counter.
#include <stdarg.h>
#include <string.h>
#include <atomic>

#include "common.cuh"

namespace rt {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};
static int g_sm_count = -1, g_smem_optin = -1, g_cc_major = -1, g_cc_minor = -1;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static int query_device() {
    if (g_sm_count > 0) return RT_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        set_error("no CUDA device available (librtrec_b200 has no CPU fallback)");
        return RT_ERR_NO_DEVICE;
    }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        set_error("cudaGetDeviceProperties failed (librtrec_b200 has no CPU fallback)");
        return RT_ERR_NO_DEVICE;
    }
    g_sm_count = p.multiProcessorCount;
    g_smem_optin = (int)p.sharedMemPerBlockOptin;
    g_cc_major = p.major;
    g_cc_minor = p.minor;
    return RT_OK;
}

static void *g_scr[SCR_SLOTS] = {nullptr};
static size_t g_scr_cap[SCR_SLOTS] = {0};

void *scratch(int slot, size_t bytes) {
    if (slot < 0 || slot >= SCR_SLOTS) return nullptr;
    if (bytes <= g_scr_cap[slot] && g_scr[slot]) return g_scr[slot];
    if (g_scr[slot]) {
        cudaDeviceSynchronize();
        cudaFree(g_scr[slot]);
        g_scr[slot] = nullptr;
        g_scr_cap[slot] = 0;
    }
    size_t want = align_up(bytes + bytes / 4 + 4096, 1 << 20);
    void *p = nullptr;
    if (cudaMalloc(&p, want) != cudaSuccess) {
        cudaGetLastError();
        want = align_up(bytes + 256, 1 << 20);
        if (cudaMalloc(&p, want) != cudaSuccess) {
            set_error("out of device memory: scratch slot %d needs %zu bytes", slot, bytes);
            return nullptr;
        }
    }
    g_scr[slot] = p;
    g_scr_cap[slot] = want;
    return p;
}

static int g_opts[OPT_COUNT] = {2, 2, 2, 0, 0, 2, 1};
int option(int key) { return (key >= 0 && key < OPT_COUNT) ? g_opts[key] : 0; }

int sm_count() { return query_device() == RT_OK ? g_sm_count : 148; }
int smem_optin() { return query_device() == RT_OK ? g_smem_optin : 232448; }

}  // namespace rt

extern "C" int rt_version(void) { return 100; }
extern "C" const char *rt_last_error(void) { return rt::g_err; }
extern "C" int rt_device_info(int *sm_count, int *smem_optin_bytes, int *cc_major, int *cc_minor) {
    int rc = rt::query_device();
    if (rc) return rc;
    if (sm_count) *sm_count = rt::g_sm_count;
    if (smem_optin_bytes) *smem_optin_bytes = rt::g_smem_optin;
    if (cc_major) *cc_major = rt::g_cc_major;
    if (cc_minor) *cc_minor = rt::g_cc_minor;
    return RT_OK;
}
extern "C" int rt_set_option(const char *name, int32_t value) {
    if (!name) { rt::set_error("rt_set_option: null name"); return RT_ERR_ARG; }
    if (!strcmp(name, "score_impl")) { rt::g_opts[rt::OPT_SCORE_IMPL] = value; return RT_OK; }
    if (!strcmp(name, "gram_impl")) { rt::g_opts[rt::OPT_GRAM_IMPL] = value; return RT_OK; }
    if (!strcmp(name, "gram_slice")) { rt::g_opts[rt::OPT_GRAM_SLICE] = value; return RT_OK; }
    if (!strcmp(name, "gram_ranges")) { rt::g_opts[rt::OPT_GRAM_RANGES] = value; return RT_OK; }
    if (!strcmp(name, "gram_adapt")) { rt::g_opts[rt::OPT_GRAM_ADAPT] = value; return RT_OK; }
    if (!strcmp(name, "gram_head")) { rt::g_opts[rt::OPT_GRAM_HEAD] = value; return RT_OK; }
    if (!strcmp(name, "solve_impl")) { rt::g_opts[rt::OPT_SOLVE_IMPL] = value; return RT_OK; }
    rt::set_error("rt_set_option: unknown option '%s'", name);
    return RT_ERR_ARG;
}
namespace rt { void upload_release(); }
extern "C" void rt_release_scratch(void) {
    cudaDeviceSynchronize();
    rt::upload_release();
    for (int i = 0; i < rt::SCR_SLOTS; ++i) {
        if (rt::g_scr[i]) cudaFree(rt::g_scr[i]);
        rt::g_scr[i] = nullptr;
        rt::g_scr_cap[i] = 0;
    }
}
extern "C" int64_t rt_launch_count(void) { return (int64_t)rt::g_launches.load(); }
extern "C" void rt_launch_count_reset(void) { rt::g_launches.store(0); }

// ---- CUDA IPC buffers (multi-GPU slab exchange over peer memory, gram3.cu) ---------------------------
extern "C" int rt_ipc_alloc(size_t bytes, void **d_ptr, uint8_t *h_handle) {
    RT_ARG(bytes > 0 && d_ptr && h_handle, "arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == RT_IPC_HANDLE_BYTES, "IPC handle size");
    void *p = nullptr;
    RT_CUDA(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        rt::set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        return RT_ERR_CUDA;
    }
    memcpy(h_handle, &h, sizeof(h));
    *d_ptr = p;
    return RT_OK;
}
extern "C" int rt_ipc_open(const uint8_t *h_handle, void **d_ptr) {
    RT_ARG(h_handle && d_ptr, "arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle, sizeof(h));
    RT_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RT_OK;
}
extern "C" int rt_ipc_close(void *d_ptr) {
    if (d_ptr) RT_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return RT_OK;
}
extern "C" int rt_ipc_free(void *d_ptr) {
    if (d_ptr) RT_CUDA(cudaFree(d_ptr));
    return RT_OK;
}
extern "C" int rt_memset(void *d_ptr, int32_t value, size_t bytes, void *stream) {
    RT_ARG(d_ptr || bytes == 0, "null pointer");
    if (bytes) RT_CUDA(cudaMemsetAsync(d_ptr, value, bytes, (cudaStream_t)stream));
    return RT_OK;
}

// ---- host-side bookkeeping -------------------------------------------------------------------------------------
// Sequential replay of LRUFreqSet.add (reference: /root/reference/rtrec/utils/lru.py:33-47, called per event with
// delta > 0 from interactions.py:115-116) over an array of non-negative integer keys.  The vectorised host path of
// rtrec_b200/utils/lru.py covers batches in which no eviction can happen; when one can (more distinct items than the
// capacity, e.g. the 105k-item H&M catalogue against the default capacity of 100,000) the end state depends on the
// order of every event, and the Python loop over 31M events took 9 s.  Same state machine here over dense arrays
// (intrusive doubly linked recency list + hit counters): ~5 ns per event.  Pure host code, no device work.
#include <vector>

extern "C" int rt_lru_replay(const int64_t *h_values, int64_t n, int64_t capacity, int64_t key_bound,
                             const int64_t *h_in_keys, const int64_t *h_in_counts, int64_t n_in,
                             int64_t *h_out_keys, int64_t *h_out_counts, int64_t *h_n_out) {
    RT_ARG(n >= 0 && capacity > 0 && key_bound > 0 && key_bound < (1ll << 31) && n_in >= 0 && n_in <= capacity, "sizes");
    RT_ARG((n == 0 || h_values) && (n_in == 0 || (h_in_keys && h_in_counts)) && h_out_keys && h_out_counts && h_n_out, "null pointer");
    const int32_t NIL = -1;
    std::vector<int32_t> prev((size_t)key_bound, NIL), next((size_t)key_bound, NIL);
    std::vector<int64_t> hits((size_t)key_bound, 0);     // 0 = not in the set
    int32_t head = NIL, tail = NIL;                       // head = least recently used
    int64_t size = 0;
    auto push_back = [&](int32_t k) {
        prev[k] = tail; next[k] = NIL;
        if (tail != NIL) next[tail] = k; else head = k;
        tail = k;
    };
    auto unlink = [&](int32_t k) {
        const int32_t p = prev[k], q = next[k];
        if (p != NIL) next[p] = q; else head = q;
        if (q != NIL) prev[q] = p; else tail = p;
    };
    for (int64_t e = 0; e < n_in; ++e) {
        const int64_t k = h_in_keys[e];
        RT_ARG(k >= 0 && k < key_bound && hits[(size_t)k] == 0 && h_in_counts[e] > 0, "initial state");
        hits[(size_t)k] = h_in_counts[e];
        push_back((int32_t)k);
        ++size;
    }
    for (int64_t e = 0; e < n; ++e) {
        const int64_t v = h_values[e];
        RT_ARG(v >= 0 && v < key_bound, "key out of range");
        const int32_t k = (int32_t)v;
        if (hits[k] > 0) {
            if (tail != k) { unlink(k); push_back(k); }
            ++hits[k];
        } else {
            if (size >= capacity) {
                const int32_t lru = head;
                unlink(lru);
                hits[lru] = 0;
                --size;
            }
            hits[k] = 1;
            push_back(k);
            ++size;
        }
    }
    int64_t m = 0;
    for (int32_t k = head; k != NIL; k = next[k]) { h_out_keys[m] = k; h_out_counts[m] = hits[k]; ++m; }
    *h_n_out = m;
    return RT_OK;
}
