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
#include <emmintrin.h>  // SSE2 streaming stores (x86-64 baseline) for the wire expansion

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
    return t > 16 ? 16 : t;
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
        cudaError_t e = cudaHostAlloc((void **)&g_ctx.pin[s], need, cudaHostAllocMapped);
        if (e != cudaSuccess)
            return note_cuda_error(e, "cudaHostAlloc(ring slot)");
    }
    g_ctx.pin_cap = need;
    return FCX_OK;
}

// Optional "packed wire" for the download side of a model (see MisesWire below): the
// outputs it covers leave the GPU as a compact stream written by a pack kernel straight
// into the pinned ring slot (zero-copy stores over PCIe) and are expanded into the
// caller's arrays by the host-thread pool -- data movement only, no arithmetic.
struct Packer {
    size_t dev_bytes = 0;   // device scratch per slot
    size_t wire_bytes = 0;  // pinned bytes per slot
    // enqueue the pack kernels of one chunk after the model kernel
    std::function<int(void **dev, void *scratch, void *wire, size_t cnt, cudaStream_t st)> enqueue;
    // expand chunk [q0, q0 + cnt) from the wire into the caller's arrays (tasks on the pool)
    std::function<void(size_t q0, size_t cnt, const void *wire, Group &g)> expand;
};

// Pipeline for callers with pageable arrays and/or a packed wire: host threads stage
// chunks through pinned ring slots on both sides of the DMA (see the header comment).
template <class Launch>
static int run_pipeline_staged(const HostArr *arr, int narr, const bool *pageable, size_t n,
                               Launch &&launch, const Packer *packer = nullptr)
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
    size_t scratch_off = 0, wire_off = 0;
    if (packer) {
        scratch_off = total;
        total += round256(packer->dev_bytes);
        wire_off = pin_total;
        pin_total += round256(packer->wire_bytes);
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
                if (packer)
                    packer->expand(it.q0, it.cnt, g_ctx.pin[it.slot] + wire_off, g);
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
        if (rc == FCX_OK && packer)
            rc = packer->enqueue(dev, g_ctx.buf[slot] + scratch_off, g_ctx.pin[slot] + wire_off, cnt, st);
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


// ---------------------------------------------------------------------------
// Packed download wire of VonMises3D (fcx_mises_evaluate_host).
//
// 288 of the 392 bytes a Mises point sends back are its 6x6 tangent, and the
// host-array path is bound by the PCIe link (49 GB/s each way on this pool,
// profiles/r1j_pcie_probe.log).  But the tangent is bitwise symmetric
// (ka*xioi + cpp*xpp + cnn*outer(xn, xn): products commute) and an ELASTIC point
// has the same tangent as every other elastic point and leaves eps_n / alpha
// untouched.  So per chunk the GPU sends
//     stress (all points, plain DMA), one flag byte per point, and for the plastic
//     points only a 28-double record [21 upper-triangle entries, eps_n[6], alpha],
// compacted in point order (exclusive scan of the flags), written by the pack
// kernel straight into the pinned ring slot; host threads mirror the triangle /
// copy the constant elastic tangent (computed once on the GPU) into the caller's
// arrays.  104 + 1 + 224 p bytes per point instead of 392 (p = plastic fraction).
// Data movement only: every double the caller sees was computed on the GPU.
// ---------------------------------------------------------------------------
constexpr int WIRE_REC = 28;

// pos[q] = number of plastic points before q in the chunk; *count = total.  One CTA.
__global__ void __launch_bounds__(1024)
    wire_scan_kernel(const unsigned char *__restrict__ flag, unsigned cnt, unsigned *__restrict__ pos,
                     unsigned *__restrict__ count_out)
{
    __shared__ unsigned warp_sum[32];
    __shared__ unsigned carry;
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned per = (cnt + 1023) / 1024;
    const unsigned b = tid * per, e = (b + per < cnt) ? b + per : cnt;
    unsigned mine = 0;
    for (unsigned q = b; q < e; ++q)
        mine += flag[q];
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d)
            incl += t;
    }
    if (lane == 31)
        warp_sum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned w = warp_sum[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= (unsigned)d)
                wi += t;
        }
        warp_sum[lane] = wi - w;  // exclusive
        if (lane == 31)
            carry = wi;
    }
    __syncthreads();
    unsigned run = warp_sum[wid] + incl - mine;
    for (unsigned q = b; q < e; ++q) {
        pos[q] = run;
        run += flag[q];
    }
    if (tid == 0)
        *count_out = carry;
}

// rec[pos[q]] = [upper triangle of tangent[q] (row-major, i <= j), eps_n[q], alpha[q]] for plastic q
__global__ void wire_pack_kernel(const unsigned char *__restrict__ flag, const unsigned *__restrict__ pos,
                                 const double *__restrict__ tangent, const double *__restrict__ eps,
                                 const double *__restrict__ alpha, unsigned cnt, double *__restrict__ rec)
{
    // k -> offset of the k-th upper-triangle entry in the row-major 6x6
    const int tri[21] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 14, 15, 16, 17, 21, 22, 23, 28, 29, 35};
    const unsigned long long total = (unsigned long long)cnt * WIRE_REC;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const unsigned q = (unsigned)(i / WIRE_REC);
        const int k = (int)(i - (unsigned long long)q * WIRE_REC);
        if (!flag[q])
            continue;
        double v;
        if (k < 21)
            v = tangent[(size_t)q * 36 + tri[k]];
        else if (k < 27)
            v = eps[(size_t)q * 6 + (k - 21)];
        else
            v = alpha[q];
        rec[(size_t)pos[q] * WIRE_REC + k] = v;
    }
}

struct MisesWire {
    // caller arrays
    double *tangent, *eps, *alpha;
    unsigned char *user_flag;  // or nullptr
    double tmpl[36];           // elastic tangent, computed on the GPU
    size_t chunk;
    // wire layout inside the pinned slot
    size_t off_count() const { return 0; }
    size_t off_flag() const { return 256; }
    size_t off_rec() const { return 256 + round256(chunk); }
    size_t wire_bytes() const { return off_rec() + chunk * WIRE_REC * sizeof(double); }
    size_t dev_bytes() const { return chunk * sizeof(unsigned); }
};

static void mises_wire_expand(const MisesWire &W, size_t q0, size_t cnt, const void *wire, Group &g)
{
    const char *base = (const char *)wire;
    const unsigned char *flag = (const unsigned char *)(base + W.off_flag());
    const double *rec = (const double *)(base + W.off_rec());
    Pool &pool = Pool::get();
    const int nt = pool_threads();
    pool.ensure(nt);
    size_t parts = cnt / 2048;
    if (parts < 1)
        parts = 1;
    if (parts > (size_t)nt)
        parts = nt;
    const size_t per = (cnt + parts - 1) / parts;
    size_t r = 0;  // records before the part
    for (size_t a = 0; a < cnt; a += per) {
        const size_t b = a + per < cnt ? a + per : cnt;
        const size_t r0 = r;
        for (size_t q = a; q < b; ++q)
            r += flag[q];
        g.add();
        pool.submit([=, &W, &g] {
            // Streaming (non-temporal) stores: the caller's arrays are written once and not read
            // here, so skipping the read-for-ownership saves a third of the host-DRAM traffic.
            const bool nt = ((reinterpret_cast<uintptr_t>(W.tangent) | reinterpret_cast<uintptr_t>(W.eps)) & 15u) == 0;
            __m128d tm[18];
            for (int k = 0; k < 18; ++k)
                tm[k] = _mm_loadu_pd(W.tmpl + 2 * k);
            size_t rr = r0;
            for (size_t q = a; q < b; ++q) {
                double *T = W.tangent + (q0 + q) * 36;
                if (flag[q]) {
                    const double *R = rec + rr * WIRE_REC;
                    double full[36];
                    int k = 0;
                    for (int i = 0; i < 6; ++i)
                        for (int j = i; j < 6; ++j, ++k) {
                            full[i * 6 + j] = R[k];
                            full[j * 6 + i] = R[k];
                        }
                    if (nt) {
                        for (int m = 0; m < 18; ++m)
                            _mm_stream_pd(T + 2 * m, _mm_loadu_pd(full + 2 * m));
                        double *E = W.eps + (q0 + q) * 6;
                        for (int m = 0; m < 3; ++m)
                            _mm_stream_pd(E + 2 * m, _mm_loadu_pd(R + 21 + 2 * m));
                    } else {
                        memcpy(T, full, sizeof full);
                        memcpy(W.eps + (q0 + q) * 6, R + 21, 6 * sizeof(double));
                    }
                    W.alpha[q0 + q] = R[27];
                    ++rr;
                } else if (nt) {
                    for (int m = 0; m < 18; ++m)
                        _mm_stream_pd(T + 2 * m, tm[m]);
                } else {
                    memcpy(T, W.tmpl, sizeof W.tmpl);
                }
            }
            if (nt)
                _mm_sfence();
            if (W.user_flag)
                memcpy(W.user_flag + q0 + a, flag + a, b - a);
            g.done();
        });
    }
}

static int g_wire = 1;  // packed download wire for the Mises host path

// launch(dev_ptrs, q_count, stream, status_dev) enqueues the kernel for one chunk.
template <class Launch>
static int run_pipeline(const HostArr *arr, int narr, size_t n, Launch &&launch,
                        const Packer *packer = nullptr)
{
    if (n == 0)
        return FCX_OK;
    std::lock_guard<std::mutex> lock(g_mu);
    if ((g_staging || packer) && n >= 4096) {  // tiny calls: the driver's own staging is as good
        bool pageable[MAXARR], any = false;
        for (int a = 0; a < narr; ++a) {
            const void *p = arr[a].src ? arr[a].src : arr[a].dst;
            pageable[a] = g_staging && p != nullptr && is_pageable(p);
            any = any || pageable[a];
        }
        if (any || packer)
            return run_pipeline_staged(arr, narr, pageable, n, launch, packer);
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

// ---------------------------------------------------------------------------
// Constant-tangent models (LinearElasticityModel, SpringKelvinModel, SpringMaxwellModel): every
// point gets the same s*s matrix, so the tangent -- 63 % of what elastic FULL sends back -- does not
// cross PCIe at all: the matrix is taken from the GPU once (one point evaluated on a zero state)
// and the host threads replicate it into the caller's array with streaming stores.
// ---------------------------------------------------------------------------
template <class Launch>
static int run_pipeline_const_tangent(HostArr *arr, int narr, int tangent_idx, int ss, size_t n,
                                      Launch &&launch)
{
    double *tangent = (double *)arr[tangent_idx].dst;
    if (!g_wire || n < 4096 || ss > 36)
        return run_pipeline(arr, narr, n, launch);
    double tmpl[36];
    {
        std::lock_guard<std::mutex> lock(g_mu);
        size_t off[MAXARR], total = 0;
        for (int a = 0; a < narr; ++a) {
            off[a] = total;
            total += round256(arr[a].bpq * 128);
        }
        int rc = ensure_ctx(total);
        if (rc != FCX_OK)
            return rc;
        cudaStream_t st = g_ctx.stream[0];
        cudaError_t e = cudaMemsetAsync(g_ctx.buf[0], 0, total, st);
        if (e != cudaSuccess)
            return note_cuda_error(e, "cudaMemsetAsync(template)");
        void *dev[MAXARR];
        for (int a = 0; a < narr; ++a)
            dev[a] = g_ctx.buf[0] + off[a];
        rc = launch(dev, 1, st, g_ctx.status);
        if (rc != FCX_OK)
            return rc;
        e = cudaMemcpyAsync(tmpl, dev[tangent_idx], sizeof(double) * ss, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(st);
        if (e != cudaSuccess)
            return note_cuda_error(e, "template download");
    }
    arr[tangent_idx].dst = nullptr;  // stays on the device
    Packer P;
    P.enqueue = [](void **, void *, void *, size_t, cudaStream_t) { return (int)FCX_OK; };
    P.expand = [tangent, ss, &tmpl](size_t q0, size_t cnt, const void *, Group &g) {
        Pool &pool = Pool::get();
        const int nt = pool_threads();
        pool.ensure(nt);
        size_t parts = cnt / 4096;
        if (parts < 1)
            parts = 1;
        if (parts > (size_t)nt)
            parts = nt;
        const size_t per = (cnt + parts - 1) / parts;
        for (size_t a = 0; a < cnt; a += per) {
            const size_t b = a + per < cnt ? a + per : cnt;
            g.add();
            pool.submit([=, &tmpl, &g] {
                double *T = tangent + (q0 + a) * ss;
                const size_t m = b - a;
                if (ss % 2 == 0 && (reinterpret_cast<uintptr_t>(T) & 15u) == 0) {
                    __m128d tm[18];
                    for (int k = 0; k < ss / 2; ++k)
                        tm[k] = _mm_loadu_pd(tmpl + 2 * k);
                    for (size_t q = 0; q < m; ++q, T += ss)
                        for (int k = 0; k < ss / 2; ++k)
                            _mm_stream_pd(T + 2 * k, tm[k]);
                    _mm_sfence();
                } else {
                    for (size_t q = 0; q < m; ++q, T += ss)
                        for (int k = 0; k < ss; ++k)
                            T[k] = tmpl[k];
                }
                g.done();
            });
        }
    };
    return run_pipeline(arr, narr, n, launch, &P);
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

int fcx_host_wire(int on)
{
    const int old = g_wire;
    if (on >= 0)
        g_wire = on ? 1 : 0;
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
    HostArr arr[3] = {{grad, nullptr, d * g * g}, {stress, stress, d * s}, {nullptr, tangent, d * s * s}};
    return run_pipeline_const_tangent(arr, 3, 2, s * s, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
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
    if (g_wire && n >= 4096) {
        // elastic tangent as the kernel produces it: one virgin point with a zero increment
        // (elastic whenever y0 > 0; otherwise fall through to the plain path)
        MisesWire W{};
        {
            std::lock_guard<std::mutex> lock(g_mu);
            int rc = ensure_ctx(4096);
            if (rc != FCX_OK)
                return rc;
            double *z = (double *)g_ctx.buf[0];
            cudaStream_t st = g_ctx.stream[0];
            cudaError_t e = cudaMemsetAsync(z, 0, (9 + 6 + 36 + 6 + 1) * d + 8, st);
            if (e != cudaSuccess)
                return note_cuda_error(e, "cudaMemsetAsync(template)");
            unsigned char *fl = (unsigned char *)(z + 58);
            rc = fcx_mises_evaluate(params, 1, z, z + 9, z + 15, z + 51, z + 57, FCX_LAYOUT_AOS, fl, nullptr, st);
            if (rc != FCX_OK)
                return rc;
            double host[59];
            e = cudaMemcpyAsync(host, z, sizeof host, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess)
                e = cudaStreamSynchronize(st);
            if (e != cudaSuccess)
                return note_cuda_error(e, "template download");
            unsigned char f0;
            memcpy(&f0, &host[58], 1);
            memcpy(W.tmpl, host + 15, sizeof W.tmpl);
            if (f0 != 0)
                W.chunk = 0;  // degenerate parameters: the zero state yields
            else
                W.chunk = 1;
        }
        if (W.chunk != 0) {
            size_t chunk = g_chunk_staged < g_chunk ? g_chunk_staged : g_chunk;
            chunk = chunk < n ? chunk : n;
            chunk = (chunk + 127) & ~(size_t)127;
            W.chunk = chunk;
            W.tangent = tangent;
            W.eps = eps_n;
            W.alpha = alpha;
            W.user_flag = plastic_flag;
            Packer P;
            P.dev_bytes = W.dev_bytes();
            P.wire_bytes = W.wire_bytes();
            P.enqueue = [&W](void **dev, void *scratch, void *wire, size_t cnt, cudaStream_t st) {
                char *wb = (char *)wire;
                unsigned *pos = (unsigned *)scratch;
                const unsigned char *fl = (const unsigned char *)dev[5];
                wire_scan_kernel<<<1, 1024, 0, st>>>(fl, (unsigned)cnt, pos, (unsigned *)(wb + W.off_count()));
                unsigned long long work = (unsigned long long)cnt * WIRE_REC;
                unsigned grid = (unsigned)((work + 255) / 256);
                const unsigned cap = (unsigned)sm_count() * 8;
                if (grid > cap)
                    grid = cap;
                wire_pack_kernel<<<grid, 256, 0, st>>>(fl, pos, (const double *)dev[2], (const double *)dev[3],
                                                       (const double *)dev[4], (unsigned)cnt,
                                                       (double *)(wb + W.off_rec()));
                g_launches.fetch_add(2, std::memory_order_relaxed);
                cudaError_t e = cudaGetLastError();
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(wb + W.off_flag(), fl, cnt, cudaMemcpyDeviceToHost, st);
                return note_cuda_error(e, "mises wire pack");
            };
            P.expand = [&W](size_t q0, size_t cnt, const void *wire, Group &g) {
                mises_wire_expand(W, q0, cnt, wire, g);
            };
            // tangent, eps_n, alpha, flag: uploaded / kept on the device as before, downloaded by the wire
            const HostArr arr[6] = {{grad, nullptr, d * 9}, {stress, stress, d * 6}, {nullptr, nullptr, d * 36},
                                    {eps_n, nullptr, d * 6}, {alpha, nullptr, d},    {nullptr, nullptr, 1}};
            return run_pipeline(arr, 6, n, [&](void **dev, size_t cnt, cudaStream_t st, int *status) {
                return fcx_mises_evaluate(params, cnt, (const double *)dev[0], (double *)dev[1],
                                          (double *)dev[2], (double *)dev[3], (double *)dev[4],
                                          FCX_LAYOUT_AOS, (unsigned char *)dev[5], status, st);
            }, &P);
        }
    }
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
    HostArr arr[5] = {{grad, nullptr, d * g * g}, {stress, stress, d * s},
                      {nullptr, tangent, d * s * s}, {ev, ev, d * s}, {et, et, d * s}};
    return run_pipeline_const_tangent(arr, 5, 2, s * s, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
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
    HostArr arr[5] = {{grad, nullptr, d * g * g}, {stress, stress, d * s},
                      {nullptr, tangent, d * s * s}, {ev, ev, d * s}, {et, et, d * s}};
    return run_pipeline_const_tangent(arr, 5, 2, s * s, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
        return fcx_maxwell_evaluate(constraint, D0, D1, mu1, tau, del_t, cnt,
                                    (const double *)dev[0], (double *)dev[1], (double *)dev[2],
                                    (double *)dev[3], (double *)dev[4], st);
    });
}

}  // extern "C"
