// fcx_host.cu -- HOST-pointer entry points (include/fcx.h, "host" family).
//
// This is the path a CPU-side caller takes: dolfinx hands `law.evaluate` host
// numpy arrays (reference solver/_lawonsubmesh.py:86-94).  The QP axis is cut
// into chunks; chunk c runs entirely on stream c % NSLOT:
//     H2D(inputs of c) -> kernel(c) -> D2H(outputs of c)
// so the upload of chunk c+1, the kernel of chunk c and the download of chunk
// c-1 overlap on the two copy engines and the SMs.  Stream order alone makes
// slot reuse safe (no events needed).  With page-locked arrays
// (fcx_host_register, or pinned allocations) the copies are true async DMA.
//
// PAGEABLE arrays -- what dolfinx actually hands over -- would make every
// cudaMemcpyAsync a synchronous, single-threaded driver staging copy (measured
// 26 MQP/s for Mises against 130 MQP/s pinned, profiles/r1k_e2e_host_memory.json).
// For them the library stages itself: a pool of host threads copies chunk c+1
// from the caller's array into a pinned ring slot while the DMA engines move
// chunk c, and a drain thread copies finished chunks from the pinned slot back
// into the caller's array.  Per-array decision (cudaPointerGetAttributes); the
// arithmetic still happens only on the GPU.
#include <cuda_runtime.h>

#include <atomic>
#include <climits>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/fcx.h"
#include "fcx_internal.h"

namespace fcx {

constexpr int NSLOT = 3;
constexpr int MAXARR = 8;

struct HostArr {
    const void *src;  // uploaded if non-null
    void *dst;        // downloaded if non-null
    size_t bpq;       // bytes per QP
};

struct HostCtx {
    cudaStream_t stream[NSLOT] = {nullptr, nullptr, nullptr};
    char *buf[NSLOT] = {nullptr, nullptr, nullptr};
    size_t cap = 0;  // bytes per slot
    char *pin[NSLOT] = {nullptr, nullptr, nullptr};  // pinned ring slots (pageable callers)
    size_t pin_cap = 0;
    cudaEvent_t done[NSLOT] = {nullptr, nullptr, nullptr};
    int *status = nullptr;
    bool ready = false;
};

static HostCtx g_ctx;
static std::mutex g_mu;
static size_t g_chunk = (size_t)1 << 18;
static size_t g_chunk_staged = (size_t)1 << 16;  // smaller chunks: the ring slots are pinned memory
static int g_staging = 1;                        // stage pageable arrays with the host-thread pool
static int g_threads = 0;                        // 0 = auto

// ---- host-thread pool (memcpy only) -----------------------------------------
class Pool {
public:
    static Pool &get()
    {
        static Pool *p = new Pool();  // leaked on purpose: no destructor-order games at exit
        return *p;
    }
    int size()
    {
        std::lock_guard<std::mutex> l(mu_);
        return (int)th_.size();
    }
    void ensure(int n)
    {
        std::lock_guard<std::mutex> l(mu_);
        while ((int)th_.size() < n)
            th_.emplace_back([this] { work(); });
    }
    void submit(std::function<void()> f)
    {
        {
            std::lock_guard<std::mutex> l(mu_);
            q_.push_back(std::move(f));
        }
        cv_.notify_one();
    }

private:
    void work()
    {
        for (;;) {
            std::function<void()> f;
            {
                std::unique_lock<std::mutex> l(mu_);
                cv_.wait(l, [this] { return !q_.empty(); });
                f = std::move(q_.front());
                q_.pop_front();
            }
            f();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    std::vector<std::thread> th_;
};

// Completion counter of one batch of pool tasks.
struct Group {
    std::mutex mu;
    std::condition_variable cv;
    int pending = 0;
    void add()
    {
        std::lock_guard<std::mutex> l(mu);
        ++pending;
    }
    void done()
    {
        std::lock_guard<std::mutex> l(mu);
        if (--pending == 0)
            cv.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [this] { return pending == 0; });
    }
};

static int pool_threads()
{
    if (g_threads > 0)
        return g_threads;
    const int hw = (int)std::thread::hardware_concurrency();
    int t = hw > 3 ? hw - 2 : 1;  // leave room for the caller and the drain thread
    return t > 12 ? 12 : t;
}

// memcpy split over the pool in pieces of >= 256 KiB
static void parallel_copy(void *dst, const void *src, size_t bytes, Group &g)
{
    Pool &pool = Pool::get();
    const int nt = pool_threads();
    pool.ensure(nt);
    size_t pieces = bytes / ((size_t)256 << 10);
    if (pieces < 1)
        pieces = 1;
    if (pieces > (size_t)nt)
        pieces = nt;
    const size_t per = ((bytes + pieces - 1) / pieces + 63) & ~(size_t)63;
    for (size_t off = 0; off < bytes; off += per) {
        const size_t len = bytes - off < per ? bytes - off : per;
        g.add();
        pool.submit([=, &g] {
            memcpy((char *)dst + off, (const char *)src + off, len);
            g.done();
        });
    }
}

static bool is_pageable(const void *p)
{
    if (p == nullptr)
        return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

static int ensure_ctx(size_t need)
{
    if (!g_ctx.ready) {
        for (int s = 0; s < NSLOT; ++s) {
            cudaError_t e = cudaStreamCreateWithFlags(&g_ctx.stream[s], cudaStreamNonBlocking);
            if (e != cudaSuccess)
                return note_cuda_error(e, "cudaStreamCreate");
        }
        for (int s = 0; s < NSLOT; ++s) {
            cudaError_t e = cudaEventCreateWithFlags(&g_ctx.done[s], cudaEventDisableTiming);
            if (e != cudaSuccess)
                return note_cuda_error(e, "cudaEventCreate");
        }
        cudaError_t e = cudaMalloc(&g_ctx.status, 2 * sizeof(int));
        if (e != cudaSuccess)
            return note_cuda_error(e, "cudaMalloc(status)");
        g_ctx.ready = true;
    }
    if (need > g_ctx.cap) {
        for (int s = 0; s < NSLOT; ++s) {
            if (g_ctx.buf[s])
                cudaFree(g_ctx.buf[s]);
            g_ctx.buf[s] = nullptr;
        }
        g_ctx.cap = 0;
        for (int s = 0; s < NSLOT; ++s) {
            cudaError_t e = cudaMalloc(&g_ctx.buf[s], need);
            if (e != cudaSuccess)
                return note_cuda_error(e, "cudaMalloc(chunk buffer)");
        }
        g_ctx.cap = need;
    }
    return FCX_OK;
}

static inline size_t round256(size_t x) { return (x + 255) & ~(size_t)255; }

static int ensure_pin(size_t need)
{
    if (need <= g_ctx.pin_cap)
        return FCX_OK;
    for (int s = 0; s < NSLOT; ++s) {
        if (g_ctx.pin[s])
            cudaFreeHost(g_ctx.pin[s]);
        g_ctx.pin[s] = nullptr;
    }
    g_ctx.pin_cap = 0;
    for (int s = 0; s < NSLOT; ++s) {
        cudaError_t e = cudaHostAlloc((void **)&g_ctx.pin[s], need, cudaHostAllocDefault);
        if (e != cudaSuccess)
            return note_cuda_error(e, "cudaHostAlloc(ring slot)");
    }
    g_ctx.pin_cap = need;
    return FCX_OK;
}

// Pipeline for callers with pageable arrays: host threads stage chunks through
// pinned ring slots on both sides of the DMA (see the header comment).
template <class Launch>
static int run_pipeline_staged(const HostArr *arr, int narr, const bool *pageable, size_t n,
                               Launch &&launch)
{
    size_t chunk = g_chunk_staged < g_chunk ? g_chunk_staged : g_chunk;  // fcx_host_chunk_qps caps both
    chunk = chunk < n ? chunk : n;
    chunk = (chunk + 127) & ~(size_t)127;
    size_t off[MAXARR], total = 0, pin_in[MAXARR], pin_out[MAXARR], pin_total = 0;
    for (int a = 0; a < narr; ++a) {
        off[a] = total;
        total += round256(arr[a].bpq * chunk);
        pin_in[a] = pin_out[a] = 0;
        if (pageable[a] && arr[a].src) {
            pin_in[a] = pin_total;
            pin_total += round256(arr[a].bpq * chunk);
        }
        if (pageable[a] && arr[a].dst) {
            pin_out[a] = pin_total;
            pin_total += round256(arr[a].bpq * chunk);
        }
    }
    int rc = ensure_ctx(total);
    if (rc != FCX_OK)
        return rc;
    rc = ensure_pin(pin_total);
    if (rc != FCX_OK)
        return rc;
    const int init[2] = {0, INT_MAX};
    cudaError_t e = cudaMemcpy(g_ctx.status, init, sizeof init, cudaMemcpyHostToDevice);
    if (e != cudaSuccess)
        return note_cuda_error(e, "cudaMemcpy(status init)");
    int device = 0;
    cudaGetDevice(&device);

    struct Item {
        size_t q0, cnt;
        int slot;
    };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Item> queue;     // chunks whose GPU work has been enqueued, in order
    bool slot_busy[NSLOT] = {false, false, false};
    bool finished = false;      // no more chunks will be queued
    int drain_rc = FCX_OK;

    std::thread drain([&] {
        cudaSetDevice(device);
        for (;;) {
            Item it;
            {
                std::unique_lock<std::mutex> l(mu);
                cv.wait(l, [&] { return !queue.empty() || finished; });
                if (queue.empty())
                    return;
                it = queue.front();
                queue.pop_front();
            }
            cudaError_t de = cudaEventSynchronize(g_ctx.done[it.slot]);
            if (de != cudaSuccess && drain_rc == FCX_OK)
                drain_rc = note_cuda_error(de, "cudaEventSynchronize(chunk)");
            if (de == cudaSuccess) {
                Group g;
                for (int a = 0; a < narr; ++a)
                    if (arr[a].dst && pageable[a])
                        parallel_copy((char *)arr[a].dst + it.q0 * arr[a].bpq,
                                      g_ctx.pin[it.slot] + pin_out[a], it.cnt * arr[a].bpq, g);
                g.wait();
            }
            {
                std::lock_guard<std::mutex> l(mu);
                slot_busy[it.slot] = false;
            }
            cv.notify_all();
        }
    });

    int slot = 0;
    for (size_t q0 = 0; q0 < n && rc == FCX_OK; q0 += chunk, slot = (slot + 1) % NSLOT) {
        const size_t cnt = (n - q0 < chunk) ? n - q0 : chunk;
        {
            std::unique_lock<std::mutex> l(mu);
            cv.wait(l, [&] { return !slot_busy[slot]; });
            slot_busy[slot] = true;
        }
        {
            Group g;  // stage-in: caller's pageable arrays -> pinned slot
            for (int a = 0; a < narr; ++a)
                if (arr[a].src && pageable[a])
                    parallel_copy(g_ctx.pin[slot] + pin_in[a], (const char *)arr[a].src + q0 * arr[a].bpq,
                                  cnt * arr[a].bpq, g);
            g.wait();
        }
        cudaStream_t st = g_ctx.stream[slot];
        void *dev[MAXARR];
        for (int a = 0; a < narr && rc == FCX_OK; ++a) {
            dev[a] = g_ctx.buf[slot] + off[a];
            if (arr[a].src) {
                const void *src = pageable[a] ? (const void *)(g_ctx.pin[slot] + pin_in[a])
                                              : (const void *)((const char *)arr[a].src + q0 * arr[a].bpq);
                e = cudaMemcpyAsync(dev[a], src, cnt * arr[a].bpq, cudaMemcpyHostToDevice, st);
                if (e != cudaSuccess)
                    rc = note_cuda_error(e, "cudaMemcpyAsync(H2D)");
            }
        }
        if (rc == FCX_OK)
            rc = launch(dev, cnt, st, g_ctx.status);
        for (int a = 0; a < narr && rc == FCX_OK; ++a) {
            if (arr[a].dst) {
                void *dst = pageable[a] ? (void *)(g_ctx.pin[slot] + pin_out[a])
                                        : (void *)((char *)arr[a].dst + q0 * arr[a].bpq);
                e = cudaMemcpyAsync(dst, dev[a], cnt * arr[a].bpq, cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess)
                    rc = note_cuda_error(e, "cudaMemcpyAsync(D2H)");
            }
        }
        if (rc == FCX_OK) {
            e = cudaEventRecord(g_ctx.done[slot], st);
            if (e != cudaSuccess)
                rc = note_cuda_error(e, "cudaEventRecord");
        }
        if (rc == FCX_OK) {
            {
                std::lock_guard<std::mutex> l(mu);
                queue.push_back(Item{q0, cnt, slot});
            }
            cv.notify_all();
        }
    }
    {
        std::lock_guard<std::mutex> l(mu);
        finished = true;
    }
    cv.notify_all();
    drain.join();
    for (int s = 0; s < NSLOT; ++s) {
        e = cudaStreamSynchronize(g_ctx.stream[s]);
        if (e != cudaSuccess && rc == FCX_OK)
            rc = note_cuda_error(e, "cudaStreamSynchronize");
    }
    if (rc != FCX_OK)
        return rc;
    if (drain_rc != FCX_OK)
        return drain_rc;
    int status[2] = {0, 0};
    e = cudaMemcpy(status, g_ctx.status, sizeof status, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess)
        return note_cuda_error(e, "cudaMemcpy(status)");
    return status[0] > 0 ? status[0] : FCX_OK;
}

// launch(dev_ptrs, q_count, stream, status_dev) enqueues the kernel for one chunk.
template <class Launch>
static int run_pipeline(const HostArr *arr, int narr, size_t n, Launch &&launch)
{
    if (n == 0)
        return FCX_OK;
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_staging && n >= 4096) {  // tiny calls: the driver's own staging is as good
        bool pageable[MAXARR], any = false;
        for (int a = 0; a < narr; ++a) {
            pageable[a] = is_pageable(arr[a].src ? arr[a].src : arr[a].dst);
            any = any || pageable[a];
        }
        if (any)
            return run_pipeline_staged(arr, narr, pageable, n, launch);
    }
    size_t chunk = g_chunk < n ? g_chunk : n;
    chunk = (chunk + 127) & ~(size_t)127;  // whole tiles
    size_t off[MAXARR], total = 0;
    for (int a = 0; a < narr; ++a) {
        off[a] = total;
        total += round256(arr[a].bpq * chunk);
    }
    int rc = ensure_ctx(total);
    if (rc != FCX_OK)
        return rc;
    const int init[2] = {0, INT_MAX};
    cudaError_t e = cudaMemcpy(g_ctx.status, init, sizeof init, cudaMemcpyHostToDevice);
    if (e != cudaSuccess)
        return note_cuda_error(e, "cudaMemcpy(status init)");

    int slot = 0;
    for (size_t q0 = 0; q0 < n; q0 += chunk, slot = (slot + 1) % NSLOT) {
        const size_t cnt = (n - q0 < chunk) ? n - q0 : chunk;
        cudaStream_t st = g_ctx.stream[slot];
        void *dev[MAXARR];
        for (int a = 0; a < narr; ++a) {
            dev[a] = g_ctx.buf[slot] + off[a];
            if (arr[a].src) {
                e = cudaMemcpyAsync(dev[a], (const char *)arr[a].src + q0 * arr[a].bpq,
                                    cnt * arr[a].bpq, cudaMemcpyHostToDevice, st);
                if (e != cudaSuccess)
                    return note_cuda_error(e, "cudaMemcpyAsync(H2D)");
            }
        }
        rc = launch(dev, cnt, st, g_ctx.status);
        if (rc != FCX_OK)
            return rc;
        for (int a = 0; a < narr; ++a) {
            if (arr[a].dst) {
                e = cudaMemcpyAsync((char *)arr[a].dst + q0 * arr[a].bpq, dev[a], cnt * arr[a].bpq,
                                    cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess)
                    return note_cuda_error(e, "cudaMemcpyAsync(D2H)");
            }
        }
    }
    for (int s = 0; s < NSLOT; ++s) {
        e = cudaStreamSynchronize(g_ctx.stream[s]);
        if (e != cudaSuccess)
            return note_cuda_error(e, "cudaStreamSynchronize");
    }
    int status[2] = {0, 0};
    e = cudaMemcpy(status, g_ctx.status, sizeof status, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess)
        return note_cuda_error(e, "cudaMemcpy(status)");
    return status[0] > 0 ? status[0] : FCX_OK;
}

}  // namespace fcx

using namespace fcx;

extern "C" {

size_t fcx_host_chunk_qps(size_t v)
{
    const size_t old = g_chunk;
    if (v > 0)
        g_chunk = v;
    return old;
}

int fcx_host_staging(int on)
{
    const int old = g_staging;
    if (on >= 0)
        g_staging = on ? 1 : 0;
    return old;
}

int fcx_host_threads(int n)
{
    const int old = pool_threads();
    if (n > 0)
        g_threads = n > 64 ? 64 : n;
    return old;
}

void fcx_host_release(void)
{
    std::lock_guard<std::mutex> lock(g_mu);
    for (int s = 0; s < NSLOT; ++s) {
        if (g_ctx.pin[s])
            cudaFreeHost(g_ctx.pin[s]);
        g_ctx.pin[s] = nullptr;
        if (g_ctx.done[s])
            cudaEventDestroy(g_ctx.done[s]);
        g_ctx.done[s] = nullptr;
        if (g_ctx.buf[s])
            cudaFree(g_ctx.buf[s]);
        g_ctx.buf[s] = nullptr;
        if (g_ctx.stream[s])
            cudaStreamDestroy(g_ctx.stream[s]);
        g_ctx.stream[s] = nullptr;
    }
    if (g_ctx.status)
        cudaFree(g_ctx.status);
    g_ctx = HostCtx{};
}

int fcx_host_register(void *ptr, size_t bytes)
{
    if (!ptr)
        return FCX_ERR_NULL;
    if (bytes == 0)
        return FCX_OK;
    return note_cuda_error(cudaHostRegister(ptr, bytes, cudaHostRegisterDefault), "cudaHostRegister");
}

int fcx_host_unregister(void *ptr)
{
    if (!ptr)
        return FCX_ERR_NULL;
    return note_cuda_error(cudaHostUnregister(ptr), "cudaHostUnregister");
}

int fcx_elastic_evaluate_host(int constraint, const double *D, size_t n, const double *grad,
                              double *stress, double *tangent)
{
    const int s = fcx_stress_strain_dim(constraint), g = fcx_geometric_dim(constraint);
    if (s < 0)
        return FCX_ERR_CONSTRAINT;
    if (n == 0)
        return FCX_OK;
    if (!D || !grad || !stress || !tangent)
        return FCX_ERR_NULL;
    const size_t d = sizeof(double);
    const HostArr arr[3] = {{grad, nullptr, d * g * g}, {stress, stress, d * s}, {nullptr, tangent, d * s * s}};
    return run_pipeline(arr, 3, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
        return fcx_elastic_evaluate(constraint, D, cnt, (const double *)dev[0], (double *)dev[1],
                                    (double *)dev[2], st);
    });
}

int fcx_mises_evaluate_host(const double *params, size_t n, const double *grad, double *stress,
                            double *tangent, double *eps_n, double *alpha,
                            unsigned char *plastic_flag)
{
    if (n == 0)
        return FCX_OK;
    if (!params || !grad || !stress || !tangent || !eps_n || !alpha)
        return FCX_ERR_NULL;
    const size_t d = sizeof(double);
    const HostArr arr[6] = {{grad, nullptr, d * 9}, {stress, stress, d * 6},
                            {nullptr, tangent, d * 36}, {eps_n, eps_n, d * 6},
                            {alpha, alpha, d}, {nullptr, plastic_flag, 1}};
    return run_pipeline(arr, 6, n, [&](void **dev, size_t cnt, cudaStream_t st, int *status) {
        return fcx_mises_evaluate(params, cnt, (const double *)dev[0], (double *)dev[1],
                                  (double *)dev[2], (double *)dev[3], (double *)dev[4],
                                  FCX_LAYOUT_AOS,
                                  plastic_flag ? (unsigned char *)dev[5] : nullptr, status, st);
    });
}

int fcx_mises_linear_hardening_evaluate_host(const double *params, size_t n, const double *grad,
                                             double *stress, double *tangent, double *history,
                                             unsigned char *plastic_flag)
{
    if (n == 0)
        return FCX_OK;
    if (!params || !grad || !stress || !tangent || !history)
        return FCX_ERR_NULL;
    const size_t d = sizeof(double);
    const HostArr arr[5] = {{grad, nullptr, d * 9}, {stress, stress, d * 6},
                            {nullptr, tangent, d * 36}, {history, history, d * 7},
                            {nullptr, plastic_flag, 1}};
    return run_pipeline(arr, 5, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
        return fcx_mises_linear_hardening_evaluate(
            params, cnt, (const double *)dev[0], (double *)dev[1], (double *)dev[2],
            (double *)dev[3], plastic_flag ? (unsigned char *)dev[4] : nullptr, st);
    });
}

int fcx_drucker_prager_evaluate_host(int hyperbolic, const double *params, size_t n,
                                     const double *grad, double *stress, double *tangent,
                                     double *history, unsigned char *plastic_flag)
{
    if (n == 0)
        return FCX_OK;
    if (!params || !grad || !stress || !tangent || !history)
        return FCX_ERR_NULL;
    const size_t d = sizeof(double);
    const HostArr arr[5] = {{grad, nullptr, d * 9}, {stress, stress, d * 6},
                            {nullptr, tangent, d * 36}, {history, history, d * 7},
                            {nullptr, plastic_flag, 1}};
    return run_pipeline(arr, 5, n, [&](void **dev, size_t cnt, cudaStream_t st, int *status) {
        return fcx_drucker_prager_evaluate(hyperbolic, params, cnt, (const double *)dev[0],
                                           (double *)dev[1], (double *)dev[2], (double *)dev[3],
                                           plastic_flag ? (unsigned char *)dev[4] : nullptr, status, st);
    });
}

int fcx_kelvin_evaluate_host(int constraint, const double *D0, const double *I2, double mu0,
                             double lam0, double mu1, double tau, double del_t, size_t n,
                             const double *grad, double *stress, double *tangent, double *ev,
                             double *et)
{
    const int s = fcx_stress_strain_dim(constraint), g = fcx_geometric_dim(constraint);
    if (s < 0)
        return FCX_ERR_CONSTRAINT;
    if (!(del_t > 0))
        return FCX_ERR_TIMESTEP;
    if (n == 0)
        return FCX_OK;
    if (!D0 || !I2 || !grad || !stress || !tangent || !ev || !et)
        return FCX_ERR_NULL;
    const size_t d = sizeof(double);
    const HostArr arr[5] = {{grad, nullptr, d * g * g}, {stress, stress, d * s},
                            {nullptr, tangent, d * s * s}, {ev, ev, d * s}, {et, et, d * s}};
    return run_pipeline(arr, 5, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
        return fcx_kelvin_evaluate(constraint, D0, I2, mu0, lam0, mu1, tau, del_t, cnt,
                                   (const double *)dev[0], (double *)dev[1], (double *)dev[2],
                                   (double *)dev[3], (double *)dev[4], st);
    });
}

int fcx_maxwell_evaluate_host(int constraint, const double *D0, const double *D1, double mu1,
                              double tau, double del_t, size_t n, const double *grad,
                              double *stress, double *tangent, double *ev, double *et)
{
    const int s = fcx_stress_strain_dim(constraint), g = fcx_geometric_dim(constraint);
    if (s < 0)
        return FCX_ERR_CONSTRAINT;
    if (!(del_t > 0))
        return FCX_ERR_TIMESTEP;
    if (n == 0)
        return FCX_OK;
    if (!D0 || !D1 || !grad || !stress || !tangent || !ev || !et)
        return FCX_ERR_NULL;
    const size_t d = sizeof(double);
    const HostArr arr[5] = {{grad, nullptr, d * g * g}, {stress, stress, d * s},
                            {nullptr, tangent, d * s * s}, {ev, ev, d * s}, {et, et, d * s}};
    return run_pipeline(arr, 5, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
        return fcx_maxwell_evaluate(constraint, D0, D1, mu1, tau, del_t, cnt,
                                    (const double *)dev[0], (double *)dev[1], (double *)dev[2],
                                    (double *)dev[3], (double *)dev[4], st);
    });
}

}  // extern "C"
