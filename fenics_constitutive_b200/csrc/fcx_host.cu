// fcx_host.cu -- HOST-pointer entry points (include/fcx.h, "host" family).
//
// This is the path a CPU-side caller takes: dolfinx hands `law.evaluate` host
// numpy arrays (reference solver/_lawonsubmesh.py:86-94).  The QP axis is cut
// into chunks; chunk c runs entirely on stream c % nslot (fcx_host_slots, default 6):
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
#include <pthread.h>
#include <sched.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/fcx.h"
#include "fcx_internal.h"

namespace fcx {

constexpr int NSLOT = 8;  // maximum; g_nslot of them are used
constexpr int MAXARR = 8;

struct HostArr {
    const void *src;  // uploaded if non-null
    void *dst;        // downloaded if non-null
    size_t bpq;       // bytes per QP
};

struct HostCtx {
    cudaStream_t stream[NSLOT] = {};
    char *buf[NSLOT] = {};
    size_t cap = 0;  // bytes per slot
    int nbuf = 0;    // slots that have a device buffer of `cap` bytes
    char *pin[NSLOT] = {};  // pinned ring slots (pageable callers, wire records)
    size_t pin_cap = 0;
    int npin = 0;
    cudaEvent_t done[NSLOT] = {};
    int *status = nullptr;
    bool ready = false;
};

static HostCtx g_ctx;
static std::mutex g_mu;
static size_t g_chunk = (size_t)1 << 18;
static size_t g_chunk_staged = (size_t)1 << 16;  // smaller chunks: the ring slots are pinned memory
// Pageable callers: the pool also copies every input into the ring slots, so it wants more threads and
// finer chunks than the page-locked path (16-core host, 16 M QPs, profiles/r2b_e2e_sweep_*.jsonl:
// pageable 105 -> 118-121 M QP/s with 12-14 threads and 32 Ki-point chunks; page-locked arrays are
// best at 4-8 threads, 152-160 M QP/s, and lose 5-10 % with 12-16).
static size_t g_chunk_pageable = (size_t)1 << 15;
static bool g_chunk_user = false;     // fcx_host_chunk_qps was called: it rules both
static bool g_call_pageable = false;  // the call being served stages pageable arrays (set under g_mu)
static int g_staging = 1;                        // stage pageable arrays with the host-thread pool
static int g_threads = 0;                        // 0 = auto
// Chunks in flight.  The pipeline is a closed loop of stations (upload engine, kernels + download,
// host-thread expansion); with 3 slots it ran at ~60 % of its slowest station
// (profiles/r1za_host_wire_stats.jsonl), hence 6.
static int g_nslot = 6;

// Phase timings of the last staged / wire pipeline run (fcx_host_stats): where the wall time of a
// host-array call goes.  GPU phases are per-chunk event intervals summed over chunks (they overlap
// across the three streams, so their sum may exceed the wall time).
struct HostStats {
    double total_s = 0, main_wait_slot_s = 0, main_stage_in_s = 0, main_enqueue_s = 0;
    double drain_event_wait_s = 0, drain_expand_s = 0;
    double gpu_h2d_s = 0, gpu_kernel_s = 0, gpu_pack_s = 0, gpu_d2h_s = 0;
    double chunks = 0, chunk_qps = 0;
};
static HostStats g_stats;
// DIAGNOSTIC ONLY (fcx_host_debug_skip): leave phases of the staged pipeline out to see what the others
// cost -- results are garbage while any bit is set.  1 uploads, 2 wire kernels, 4 download DMAs,
// 8 host expansion, 16 model kernel.
static int g_skip = 0;
static int g_trace = 0;  // fcx_host_trace(1): also record per-chunk GPU phase events
static cudaEvent_t g_tev[NSLOT][5] = {};
static cudaEvent_t g_t0ev = nullptr;
// Per-chunk timeline of the last traced run (fcx_host_timeline), seconds since the call began:
// [chunk, slot, host: slot acquired, enqueued | gpu: chain start, h2d done, kernel done, pack done,
//  d2h done | host: drain woke, expansion done]
constexpr int TL_COLS = 11;
static std::vector<double> g_timeline;

static inline double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- NUMA placement ----------------------------------------------------------
// The pool threads touch every byte the DMA engines move, so they belong on the NUMA node the GPU
// hangs off (and the pinned ring slots in that node's memory): a copy that crosses the socket
// interconnect costs twice.  The node comes from sysfs (/sys/bus/pci/devices/<gpu>/numa_node), its
// CPUs from /sys/devices/system/node/nodeN/cpulist, intersected with the CPUs this process may run
// on (container cpuset).  A single-node host, a VM that hides the topology (numa_node = -1) or an
// empty intersection leave everything unpinned.  fcx_host_numa(0) switches it off.
static int g_numa = 1;
struct NumaInfo {
    int node = -1;       // NUMA node of the bound GPU, -1 = unknown / not applicable
    int ncpus = 0;       // CPUs of that node this process may use
    int allowed = 0;     // CPUs this process may use at all
    cpu_set_t set;
    bool valid = false;  // pin threads to `set`
};
static NumaInfo g_numa_info;
static std::once_flag g_numa_once;

static int allowed_cpus()
{
    cpu_set_t s;
    CPU_ZERO(&s);
    if (sched_getaffinity(0, sizeof s, &s) == 0) {
        const int c = CPU_COUNT(&s);
        if (c > 0)
            return c;
    }
    const int hw = (int)std::thread::hardware_concurrency();
    return hw > 0 ? hw : 1;
}

static void numa_probe()
{
    NumaInfo &N = g_numa_info;
    CPU_ZERO(&N.set);
    N.allowed = allowed_cpus();
    int dev = 0;
    char bus[32] = "";
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, sizeof bus, dev) != cudaSuccess) {
        cudaGetLastError();
        return;
    }
    for (char *p = bus; *p; ++p)
        if (*p >= 'A' && *p <= 'Z')
            *p += 'a' - 'A';
    char path[128];
    snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f)
        return;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1)
        node = -1;
    fclose(f);
    N.node = node;
    if (node < 0)
        return;
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    f = fopen(path, "r");
    if (!f)
        return;
    char list[4096] = "";
    if (!fgets(list, sizeof list, f))
        list[0] = 0;
    fclose(f);
    cpu_set_t may;
    CPU_ZERO(&may);
    if (sched_getaffinity(0, sizeof may, &may) != 0)
        return;
    for (char *tok = strtok(list, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1)
            b = a;
        if (k < 1)
            continue;
        for (int c = a; c <= b && c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &may))
                CPU_SET(c, &N.set);
    }
    N.ncpus = CPU_COUNT(&N.set);
    // pin only if the node is a proper, non-empty subset of what we may use (otherwise nothing to gain)
    N.valid = N.ncpus > 0 && N.ncpus < N.allowed;
}

static const NumaInfo &numa_info()
{
    std::call_once(g_numa_once, numa_probe);
    return g_numa_info;
}

static void pin_this_thread_to_gpu_node()
{
    const NumaInfo &N = numa_info();
    if (g_numa && N.valid)
        pthread_setaffinity_np(pthread_self(), sizeof N.set, &N.set);
}

// Prefer the GPU's node for the pages of the next allocations of this thread (the pinned ring
// slots); `restore` puts the default policy back.  Raw syscall: libnuma is not assumed.
static void prefer_gpu_node_memory(bool restore)
{
#ifdef SYS_set_mempolicy
    const NumaInfo &N = numa_info();
    if (!g_numa || !N.valid || N.node < 0 || N.node >= 1024)
        return;
    if (restore) {
        syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0);
        return;
    }
    unsigned long mask[16] = {};
    mask[N.node / (8 * sizeof(unsigned long))] |= 1UL << (N.node % (8 * sizeof(unsigned long)));
    syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8);
#else
    (void)restore;
#endif
}

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
        pin_this_thread_to_gpu_node();
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
    std::function<void()> on_zero;  // optional: runs (on the finishing thread) when the count reaches 0
    void add()
    {
        std::lock_guard<std::mutex> l(mu);
        ++pending;
    }
    void done()
    {
        std::function<void()> fire;  // a copy: the Group may be gone once on_zero has run
        {
            std::lock_guard<std::mutex> l(mu);
            if (--pending == 0) {
                cv.notify_all();
                fire = on_zero;
            }
        }
        if (fire)
            fire();
    }
    void wait()
    {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [this] { return pending == 0; });
    }
};

// Pool size.  wide = true for the work that is pure streaming fill (constant-tangent models: 288 B per
// point written by the host threads, 286 M QP/s with 14 threads, profiles/r1x); the plastic wire and
// the pageable staging ran best with 8 on the 16-core hosts (profiles/r1zf) -- more threads only add
// contention with the DMA traffic there.
static int pool_threads(bool wide = false)
{
    if (g_threads > 0)
        return g_threads;
    int hw = allowed_cpus();  // the container's cpuset, not the machine's core count
    // one process per GPU (torchrun): the ranks of a node share its cores
    static const int local_world = [] {
        const char *e = getenv("LOCAL_WORLD_SIZE");
        const int v = e ? atoi(e) : 1;
        return v > 0 ? v : 1;
    }();
    int t = hw > 3 ? hw - 2 : 1;  // leave room for the caller and the drain thread
    if (local_world > 1) {
        t = hw / local_world;
        t = t < 4 ? (hw > 3 ? 4 : 1) : t;
    }
    const int cap = wide ? 16 : (g_call_pageable ? 14 : 8);
    return t > cap ? cap : t;
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
    if (need > g_ctx.cap || g_ctx.nbuf < g_nslot) {
        if (need < g_ctx.cap)
            need = g_ctx.cap;
        for (int s = 0; s < NSLOT; ++s) {
            if (g_ctx.buf[s])
                cudaFree(g_ctx.buf[s]);
            g_ctx.buf[s] = nullptr;
        }
        g_ctx.cap = 0;
        g_ctx.nbuf = 0;
        for (int s = 0; s < g_nslot; ++s) {
            cudaError_t e = cudaMalloc(&g_ctx.buf[s], need);
            // zeroed once: the tangent part is only ever written by bulk async (TMA) stores, which
            // compute-sanitizer's initcheck does not track (100 % false positives in the pack kernel)
            if (e == cudaSuccess)
                e = cudaMemset(g_ctx.buf[s], 0, need);
            if (e != cudaSuccess)
                return note_cuda_error(e, "cudaMalloc(chunk buffer)");
        }
        g_ctx.cap = need;
        g_ctx.nbuf = g_nslot;
    }
    return FCX_OK;
}

static inline size_t round256(size_t x) { return (x + 255) & ~(size_t)255; }

static int ensure_pin(size_t need)
{
    if (need <= g_ctx.pin_cap && g_ctx.npin >= g_nslot)
        return FCX_OK;
    if (need < g_ctx.pin_cap)
        need = g_ctx.pin_cap;
    for (int s = 0; s < NSLOT; ++s) {
        if (g_ctx.pin[s])
            cudaFreeHost(g_ctx.pin[s]);
        g_ctx.pin[s] = nullptr;
    }
    g_ctx.pin_cap = 0;
    g_ctx.npin = 0;
    prefer_gpu_node_memory(false);
    for (int s = 0; s < g_nslot; ++s) {
        cudaError_t e = cudaHostAlloc((void **)&g_ctx.pin[s], need, cudaHostAllocMapped);
        if (e != cudaSuccess) {
            prefer_gpu_node_memory(true);
            return note_cuda_error(e, "cudaHostAlloc(ring slot)");
        }
    }
    prefer_gpu_node_memory(true);
    g_ctx.pin_cap = need;
    g_ctx.npin = g_nslot;
    return FCX_OK;
}

// Optional "packed wire" for the download side of a model (see PlasticWire below): the
// outputs it covers leave the GPU as a compact stream written by a pack kernel straight
// into the pinned ring slot (zero-copy stores over PCIe) and are expanded into the
// caller's arrays by the host-thread pool -- data movement only, no arithmetic.
struct Packer {
    size_t dev_bytes = 0;   // device scratch per slot
    size_t wire_bytes = 0;  // pinned bytes per slot
    // enqueue the pack kernels of one chunk after the model kernel
    std::function<int(void **dev, void *scratch, void *wire, size_t q0, size_t cnt, cudaStream_t st)> enqueue;
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
    g_call_pageable = false;
    for (int a = 0; a < narr; ++a)
        g_call_pageable = g_call_pageable || pageable[a];
    if (g_call_pageable && !g_chunk_user && g_chunk_pageable < chunk)
        chunk = g_chunk_pageable;
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
    HostStats S;
    S.chunk_qps = (double)chunk;
    const bool trace = g_trace != 0;
    if (trace)
        for (int s = 0; s < NSLOT; ++s)
            for (int k = 0; k < 5; ++k)
                if (g_tev[s][k] == nullptr)
                    cudaEventCreate(&g_tev[s][k]);
    std::vector<double> TL;
    if (trace) {
        if (g_t0ev == nullptr)
            cudaEventCreate(&g_t0ev);
        TL.assign(((n + chunk - 1) / chunk) * TL_COLS, 0.0);
        cudaEventRecord(g_t0ev, g_ctx.stream[0]);
    }
    const double t_begin = now_s();

    struct Item {
        size_t q0, cnt;
        int slot;
    };
    auto tl = [&](size_t q0, int col) -> double & { return TL[(q0 / chunk) * TL_COLS + col]; };
    const int nslot = g_nslot;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Item> queue;     // chunks whose GPU work has been enqueued, in order
    bool slot_busy[NSLOT] = {};
    bool finished = false;      // no more chunks will be queued
    int drain_rc = FCX_OK;
    // Expansion of a finished chunk runs on the pool while the drain thread already waits for the
    // next chunk; the last task of a chunk frees its slot.
    Group expand_group[NSLOT];
    double expand_t0[NSLOT] = {};
    size_t expand_q0[NSLOT] = {};
    for (int s = 0; s < nslot; ++s)
        expand_group[s].on_zero = [&, s] {
            std::lock_guard<std::mutex> l(mu);  // notify under the lock: nothing is touched after it
            S.drain_expand_s += now_s() - expand_t0[s];
            if (trace)
                tl(expand_q0[s], 10) = now_s() - t_begin;
            slot_busy[s] = false;
            cv.notify_all();
        };

    std::thread drain([&] {
        cudaSetDevice(device);
        pin_this_thread_to_gpu_node();
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
            const double t0 = now_s();
            cudaError_t de = cudaEventSynchronize(g_ctx.done[it.slot]);
            const double t1 = now_s();
            S.drain_event_wait_s += t1 - t0;
            if (de != cudaSuccess && drain_rc == FCX_OK)
                drain_rc = note_cuda_error(de, "cudaEventSynchronize(chunk)");
            if (de == cudaSuccess && trace) {
                float ms[4] = {0, 0, 0, 0};
                for (int k = 0; k < 4; ++k)
                    cudaEventElapsedTime(&ms[k], g_tev[it.slot][k], g_tev[it.slot][k + 1]);
                S.gpu_h2d_s += 1e-3 * ms[0];
                S.gpu_kernel_s += 1e-3 * ms[1];
                S.gpu_pack_s += 1e-3 * ms[2];
                S.gpu_d2h_s += 1e-3 * ms[3];
                for (int k = 0; k < 5; ++k) {
                    float abs_ms = 0;
                    cudaEventElapsedTime(&abs_ms, g_t0ev, g_tev[it.slot][k]);
                    tl(it.q0, 4 + k) = 1e-3 * abs_ms;
                }
                tl(it.q0, 0) = (double)(it.q0 / chunk);
                tl(it.q0, 1) = it.slot;
                tl(it.q0, 9) = t1 - t_begin;
            }
            Group &g = expand_group[it.slot];
            expand_t0[it.slot] = t1;
            expand_q0[it.slot] = it.q0;
            g.add();  // guard: the slot is not freed before every task has been submitted
            if (de == cudaSuccess && !(g_skip & 8)) {
                for (int a = 0; a < narr; ++a)
                    if (arr[a].dst && pageable[a])
                        parallel_copy((char *)arr[a].dst + it.q0 * arr[a].bpq,
                                      g_ctx.pin[it.slot] + pin_out[a], it.cnt * arr[a].bpq, g);
                if (packer)
                    packer->expand(it.q0, it.cnt, g_ctx.pin[it.slot] + wire_off, g);
            }
            g.done();
        }
    });

    int slot = 0;
    for (size_t q0 = 0; q0 < n && rc == FCX_OK; q0 += chunk, slot = (slot + 1) % nslot) {
        const size_t cnt = (n - q0 < chunk) ? n - q0 : chunk;
        const double tw0 = now_s();
        {
            std::unique_lock<std::mutex> l(mu);
            cv.wait(l, [&] { return !slot_busy[slot]; });
            slot_busy[slot] = true;
        }
        const double tw1 = now_s();
        S.main_wait_slot_s += tw1 - tw0;
        S.chunks += 1;
        if (trace)
            tl(q0, 2) = tw1 - t_begin;
        {
            Group g;  // stage-in: caller's pageable arrays -> pinned slot
            for (int a = 0; a < narr; ++a)
                if (arr[a].src && pageable[a])
                    parallel_copy(g_ctx.pin[slot] + pin_in[a], (const char *)arr[a].src + q0 * arr[a].bpq,
                                  cnt * arr[a].bpq, g);
            g.wait();
        }
        const double tw2 = now_s();
        S.main_stage_in_s += tw2 - tw1;
        cudaStream_t st = g_ctx.stream[slot];
        if (trace)
            cudaEventRecord(g_tev[slot][0], st);
        void *dev[MAXARR];
        for (int a = 0; a < narr && rc == FCX_OK; ++a) {
            dev[a] = g_ctx.buf[slot] + off[a];
            if (arr[a].src && !(g_skip & 1)) {
                const void *src = pageable[a] ? (const void *)(g_ctx.pin[slot] + pin_in[a])
                                              : (const void *)((const char *)arr[a].src + q0 * arr[a].bpq);
                e = cudaMemcpyAsync(dev[a], src, cnt * arr[a].bpq, cudaMemcpyHostToDevice, st);
                if (e != cudaSuccess)
                    rc = note_cuda_error(e, "cudaMemcpyAsync(H2D)");
            }
        }
        if (trace)
            cudaEventRecord(g_tev[slot][1], st);
        if (rc == FCX_OK && !(g_skip & 16))
            rc = launch(dev, cnt, st, g_ctx.status);
        if (trace)
            cudaEventRecord(g_tev[slot][2], st);
        if (rc == FCX_OK && packer && !(g_skip & 2))
            rc = packer->enqueue(dev, g_ctx.buf[slot] + scratch_off, g_ctx.pin[slot] + wire_off, q0, cnt, st);
        if (trace)
            cudaEventRecord(g_tev[slot][3], st);
        for (int a = 0; a < narr && rc == FCX_OK; ++a) {
            if (arr[a].dst && !(g_skip & 4)) {
                void *dst = pageable[a] ? (void *)(g_ctx.pin[slot] + pin_out[a])
                                        : (void *)((char *)arr[a].dst + q0 * arr[a].bpq);
                e = cudaMemcpyAsync(dst, dev[a], cnt * arr[a].bpq, cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess)
                    rc = note_cuda_error(e, "cudaMemcpyAsync(D2H)");
            }
        }
        if (trace)
            cudaEventRecord(g_tev[slot][4], st);
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
        } else {
            std::lock_guard<std::mutex> l(mu);
            slot_busy[slot] = false;  // never queued: nobody else will free it
        }
        S.main_enqueue_s += now_s() - tw2;
        if (trace)
            tl(q0, 3) = now_s() - t_begin;
    }
    {
        std::lock_guard<std::mutex> l(mu);
        finished = true;
    }
    cv.notify_all();
    drain.join();
    {
        std::unique_lock<std::mutex> l(mu);  // the expansions still running on the pool
        cv.wait(l, [&] {
            for (int s = 0; s < nslot; ++s)
                if (slot_busy[s])
                    return false;
            return true;
        });
    }
    for (int s = 0; s < NSLOT; ++s) {
        e = cudaStreamSynchronize(g_ctx.stream[s]);
        if (e != cudaSuccess && rc == FCX_OK)
            rc = note_cuda_error(e, "cudaStreamSynchronize");
    }
    S.total_s = now_s() - t_begin;
    g_stats = S;
    if (trace)
        g_timeline.swap(TL);
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
// Download wire of the plastic models (VonMises3D and the comfe-rs MisesPlasticityLinearHardening3D /
// DruckerPrager3D / DruckerPragerHyperbolic3D mirrors).
//
// 288 of the 392 bytes a 3-D point sends back are its 6x6 tangent, and the host-array path is
// bound by the PCIe link (49 GB/s each way on this pool, profiles/r1j_pcie_probe.log) and by what
// the host threads can copy.  But an ELASTIC point has the same tangent as every other elastic
// point and leaves its history untouched.  So per chunk the GPU sends
//     stress (all points, plain DMA), one flag byte per point, and for the PLASTIC points only
//     a record, compacted in point order (exclusive scan of the flags) and written by the pack
//     kernel straight into the pinned ring slot (zero-copy stores over PCIe);
// host threads scatter the records and copy the constant elastic tangent (computed once on the
// GPU) into the caller's arrays.  Data movement only: every double the caller sees was computed
// on the GPU.  What a record holds depends on the caller's tangent array:
//   * page-locked (pinned / fcx_host_register) and 16-byte aligned -- DIRECT mode: a kernel
//     stores the plastic points' 36 tangent entries straight into the caller's array through its
//     device alias (288-byte runs of 16-byte stores over PCIe); the record is the history only
//     (7 doubles).  PCIe: 48 + 1 + 344 p bytes per point; the host threads touch 56 p bytes of
//     records and stream 288 (1 - p) bytes of constant tangent -- instead of reading and
//     re-writing every tangent (mode below), which was what bound the pinned path.
//   * pageable -- SLOT mode: record = tangent + history, where a bitwise-symmetric tangent
//     (VonMises3D: ka*xioi + cpp*xpp + cnn*outer(xn, xn), products commute) travels as its 21
//     upper-triangle entries and is mirrored by the host threads; others travel as all 36.
// ---------------------------------------------------------------------------

// list[r] = index of the r-th plastic point of the chunk (point order); *count = their number.
// One CTA; every round takes 16 consecutive flags per thread with one coalesced 16-byte load
// (the first version let each thread walk 64 consecutive bytes: 32 sectors per warp instruction,
// 0.15 ms per 64 Ki-point chunk with the PCIe link idle behind it, profiles/r1ze).
__global__ void __launch_bounds__(1024)
    wire_scan_kernel(const unsigned char *__restrict__ flag, unsigned cnt, unsigned *__restrict__ list,
                     unsigned *__restrict__ count_out)
{
    __shared__ unsigned warp_sum[32];
    __shared__ unsigned round_total;
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    unsigned carry = 0;
    for (unsigned base = 0; base < cnt; base += 16384) {
        const unsigned q0 = base + tid * 16;
        unsigned w[4] = {0, 0, 0, 0};
        if (q0 + 16 <= cnt) {
            const uint4 v = *reinterpret_cast<const uint4 *>(flag + q0);  // chunk buffers are 256-byte aligned
            w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
        } else {
            for (unsigned k = 0; k < 16; ++k)
                if (q0 + k < cnt && flag[q0 + k])
                    w[k >> 2] |= 1u << (8 * (k & 3));
        }
        const unsigned mine = __popc(w[0] & 0x01010101u) + __popc(w[1] & 0x01010101u) +
                              __popc(w[2] & 0x01010101u) + __popc(w[3] & 0x01010101u);
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
            const unsigned ws = warp_sum[lane];
            unsigned wi = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= (unsigned)d)
                    wi += t;
            }
            warp_sum[lane] = wi - ws;  // exclusive
            if (lane == 31)
                round_total = wi;
        }
        __syncthreads();
        unsigned run = carry + warp_sum[wid] + incl - mine;
#pragma unroll
        for (unsigned k = 0; k < 16; ++k)
            if ((w[k >> 2] >> (8 * (k & 3))) & 1u)
                list[run++] = q0 + k;
        carry += round_total;
        __syncthreads();  // warp_sum / round_total are rewritten in the next round
    }
    if (tid == 0)
        *count_out = carry;
}

// The compact record stream, written front to back: thread o stores double o, so every warp
// writes 256 consecutive bytes of pinned memory whatever the flags are (52 GB/s over PCIe against
// 40 for a version whose lanes idled on elastic points, profiles/r1zb_pcie_probe2.log).
// rec[r] = [tangent part of point q = list[r] (nt = 0: none, 21: upper triangle row-major i <= j,
//           36: all), h0[q][0..w0), h1[q][0..w1)]
__global__ void wire_pack_kernel(const unsigned *__restrict__ list, const unsigned *__restrict__ count,
                                 const double *__restrict__ tangent, int nt, const double *__restrict__ h0,
                                 int w0, const double *__restrict__ h1, int w1, double *__restrict__ rec)
{
    // k -> offset of the k-th upper-triangle entry in the row-major 6x6
    const int tri[21] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 14, 15, 16, 17, 21, 22, 23, 28, 29, 35};
    const int R = nt + w0 + w1;
    const unsigned long long total = (unsigned long long)*count * R;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long o = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += stride) {
        const unsigned r = (unsigned)(o / R);
        const int k = (int)(o - (unsigned long long)r * R);
        const size_t q = list[r];
        double v;
        if (k < nt)
            v = tangent[q * 36 + (nt == 21 ? tri[k] : k)];
        else if (k < nt + w0)
            v = h0[q * w0 + (k - nt)];
        else
            v = h1[q * w1 + (k - nt - w0)];
        rec[o] = v;
    }
}

// DIRECT mode: the plastic points' tangents go from the chunk buffer straight into the caller's
// page-locked array (`dst` = its device alias at the chunk's first point), 16 bytes per thread,
// consecutive threads on consecutive addresses (18 per point).
__global__ void wire_direct_kernel(const unsigned char *__restrict__ flag, const double2 *__restrict__ tangent,
                                   unsigned cnt, double2 *__restrict__ dst)
{
    const unsigned long long total = (unsigned long long)cnt * 18;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const unsigned q = (unsigned)(i / 18);
        if (flag[q])
            dst[i] = tangent[i];
    }
}

struct PlasticWire {
    // caller arrays
    double *tangent = nullptr;
    double *tangent_dev = nullptr;  // device alias of `tangent` (DIRECT mode) or nullptr
    int nt = 36;                    // tangent doubles per record: 0 (direct), 21 (triangle) or 36
    int nh = 1;                     // history arrays scattered from the record
    double *hist[2] = {nullptr, nullptr};
    int hw[2] = {0, 0};
    unsigned char *user_flag = nullptr;
    double tmpl[36];  // elastic tangent, computed on the GPU
    size_t chunk = 0;
    // Mixed download: `mix` percent of the chunks leave by plain DMA straight into the caller's page-locked
    // arrays (392 B per point over the link, no host-thread byte), the others by records (161 B per point over
    // the link, 392 B per point written by the host threads): the link and the pool threads work side by side
    // on DIFFERENT chunks (an option, measured slower on this pool's hosts: see g_wire_mix).
    int mix = 0;
    bool plain(size_t q0) const
    {
        const size_t ci = chunk ? q0 / chunk : 0;
        return mix > 0 && ((ci + 1) * (size_t)mix) / 100 != (ci * (size_t)mix) / 100;
    }
    int rec() const { return nt + hw[0] + hw[1]; }
    // wire layout inside the pinned slot
    size_t off_count() const { return 0; }
    size_t off_flag() const { return 256; }
    size_t off_rec() const { return 256 + round256(chunk); }
    size_t wire_bytes() const { return off_rec() + chunk * rec() * sizeof(double); }
    size_t dev_bytes() const { return (chunk + 64) * sizeof(unsigned); }  // list + count
};

static inline void copy_doubles(double *dst, const double *src, int n, bool stream)
{
    if (stream && n % 2 == 0 && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
        for (int m = 0; m < n; m += 2)
            _mm_stream_pd(dst + m, _mm_loadu_pd(src + m));
    } else {
        for (int m = 0; m < n; ++m)
            dst[m] = src[m];
    }
}

static void plastic_wire_expand(const PlasticWire &W, size_t q0, size_t cnt, const void *wire, Group &g)
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
            const bool nts = (reinterpret_cast<uintptr_t>(W.tangent) & 15u) == 0;
            const int R = W.rec();
            __m128d tm[18];
            for (int k = 0; k < 18; ++k)
                tm[k] = _mm_loadu_pd(W.tmpl + 2 * k);
            size_t rr = r0;
            for (size_t q = a; q < b; ++q) {
                // W.tangent == nullptr: stress-only call, the records carry the history alone
                double *T = W.tangent ? W.tangent + (q0 + q) * 36 : nullptr;
                if (flag[q]) {
                    const double *P = rec + rr * R;
                    if (T == nullptr) {
                    } else if (W.nt == 21) {
                        double full[36];
                        int k = 0;
                        for (int i = 0; i < 6; ++i)
                            for (int j = i; j < 6; ++j, ++k) {
                                full[i * 6 + j] = P[k];
                                full[j * 6 + i] = P[k];
                            }
                        copy_doubles(T, full, 36, nts);
                    } else if (W.nt == 36) {
                        copy_doubles(T, P, 36, nts);
                    }  // nt == 0: the GPU wrote this tangent in place
                    const double *H = P + W.nt;
                    for (int h = 0; h < W.nh; ++h) {
                        copy_doubles(W.hist[h] + (q0 + q) * W.hw[h], H, W.hw[h], nts && W.hw[h] > 1);
                        H += W.hw[h];
                    }
                    ++rr;
                } else if (T == nullptr) {
                } else if (nts) {
                    for (int m = 0; m < 18; ++m)
                        _mm_stream_pd(T + 2 * m, tm[m]);
                } else {
                    memcpy(T, W.tmpl, sizeof W.tmpl);
                }
            }
            _mm_sfence();
            if (W.user_flag)
                memcpy(W.user_flag + q0 + a, flag + a, b - a);
            g.done();
        });
    }
}

// Download wire of the plastic host paths: 0 off, 1 slot records, 2 + direct tangents for page-locked
// arrays.  Default 1: on this pool's hosts the direct stores lose to the records (122-133 vs 150-177
// M QP/s, profiles/r1zf_host_wire_stats.jsonl): the whole path is bound by what host memory and the
// link move together (~51 GB/s of PCIe traffic in both wire modes while the host threads stream the
// tangents); most likely the GPU's 288-byte runs and the host threads' fills of the neighbouring
// elastic runs fight over the partial cache lines they share (not investigated further).
// Tried and reverted (profiles/r1zg_*): packing the records in device memory and letting a second
// drain stage DMA exactly `count` of them -- same link rate, one more host round trip per chunk.
// 3 = AUTO (the default): the record wire (1) unless several ranks share the host's memory system AND
// every result array of the call is page-locked -- then the direct wire (2): the GPU stores the plastic
// tangents in place, the few pool threads a rank has left only fill the elastic runs.  Measured on the
// 32-core hosts of this pool, 16 M QPs per rank, page-locked arrays (profiles/r2f_e2e_sweep_pinned_n*.jsonl):
//   4 ranks: wire 0 / 1 / 2 = 221 / 224 / 241 M QP/s,   8 ranks: 202 / 216 / 233
// (one rank on a 16-core host: 124 / 156 / 135, profiles/r2b_e2e_sweep_pinned.jsonl).  Plain DMA of every array
// (0) -- no host-thread byte at all -- is the SLOWEST with many ranks: all of them are bound by the host's
// DRAM (166-172 GB/s memcpy on these hosts), and 392 B/QP of DMA writes cost more of it than 161 B/QP of
// records plus streaming fills.  Threshold: FCX_WIRE_AUTO_RANKS (default 4).
static int g_wire = 3;
// Share (percent) of the chunks of a record-wire call (1) that leave by plain DMA instead, when every result
// array of the call is page-locked.  The idea: with records alone the link idles at 161 B per point while the
// pool threads write 392 B per point; a plain-DMA chunk costs the link 392 B per point and the threads nothing,
// so the two could work side by side on different chunks.  (The premise was wrong: four pool threads already
// keep up with the record expansion -- profiles/r2b_e2e_sweep_pinned.jsonl -- and no single phase bounds the
// call, the phases slow each other down through the host's memory system)  -1 = AUTO.
// MEASURED (16 M points, page-locked arrays, one GPU on a 16-core host, profiles/r2s_e2e_mix_pinned.jsonl):
// 0 / 15 / 25 / 35 / 50 / 100 % = 158 / 150 / 144 / 139 / 131 / 122 M QP/s, the same with 12 pool threads --
// the call time grows linearly with the share, i.e. the plain chunks and the record chunks do not overlap the
// way two independent resources would: AUTO therefore resolves to 0 (FCX_WIRE_MIX_AUTO overrides); the
// option stays for hosts with a faster link or fewer cores.
static int g_wire_mix = -1;

static int local_world_size()
{
    static const int v = [] {
        const char *e = getenv("LOCAL_WORLD_SIZE");
        const int w = e ? atoi(e) : 1;
        return w > 0 ? w : 1;
    }();
    return v;
}

static bool page_locked(const void *p)
{
    if (p == nullptr)
        return true;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// The wire mode a plastic-model host call runs with (g_wire resolved for this call's arrays).
static int effective_wire(const void *stress, const void *tangent, const void *h0, const void *h1)
{
    if (g_wire != 3)
        return g_wire;
    if (tangent == nullptr)  // stress-only call: 104 B per point come back, plain DMA beats the record wire
        return 0;            // (page-locked 250 vs 224 M QP/s, pageable 152 vs 148; profiles/r2b_e2e_sweep_so.jsonl)
    static const int ranks = [] {
        const char *e = getenv("FCX_WIRE_AUTO_RANKS");
        const int v = e ? atoi(e) : 4;
        return v > 0 ? v : 4;
    }();
    if (local_world_size() >= ranks && page_locked(stress) && page_locked(tangent) && page_locked(h0) &&
        page_locked(h1))
        return 2;
    return 1;
}
static int g_last_wire = -1;  // what the last plastic host call resolved to (fcx_host_wire_used)
static int g_last_wire_mix = 0;

// AUTO: FCX_WIRE_MIX_AUTO percent (default below) when one rank has the host's cores to itself; with several
// ranks per host the host's DRAM is the bound and plain DMA costs more of it than records do (g_wire above).
static int effective_wire_mix()
{
    if (g_wire_mix >= 0)
        return g_wire_mix;
    static const int pct = [] {
        const char *e = getenv("FCX_WIRE_MIX_AUTO");
        const int v = e ? atoi(e) : 0;
        return v < 0 ? 0 : (v > 100 ? 100 : v);
    }();
    return local_world_size() > 1 ? 0 : pct;
}

// Device alias of a page-locked host range (pinned allocation or cudaHostRegister), or nullptr.
static void *device_alias(const void *p, size_t bytes)
{
    if (p == nullptr || bytes == 0)
        return nullptr;
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess ||
        cudaPointerGetAttributes(&a1, (const char *)p + bytes - 1) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || a0.devicePointer == nullptr ||
        (const char *)a1.devicePointer - (const char *)a0.devicePointer != (ptrdiff_t)(bytes - 1))
        return nullptr;
    return a0.devicePointer;
}

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
    for (size_t q0 = 0; q0 < n; q0 += chunk, slot = (slot + 1) % g_nslot) {
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
    if (tangent == nullptr) {  // stress-only call: the tangent has no device slot either
        arr[tangent_idx].bpq = 0;
        return run_pipeline(arr, narr, n, launch);
    }
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
    P.enqueue = [](void **, void *, void *, size_t, size_t, cudaStream_t) { return (int)FCX_OK; };
    P.expand = [tangent, ss, &tmpl](size_t q0, size_t cnt, const void *, Group &g) {
        Pool &pool = Pool::get();
        const int nt = pool_threads(true);
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

// ---------------------------------------------------------------------------
// Host pipeline shared by the plastic models.  Chunk arrays (dev[] of `launch`):
//   0 grad [9]   1 stress [6]   2 tangent [36]   3 history 0   4 history 1 (or unused)   5 flag
// With the wire on, tangent / history / flag stay on the device and leave through the
// PlasticWire (see above); otherwise every array is downloaded by plain DMA.
// ---------------------------------------------------------------------------
struct PlasticHost {
    const double *grad;
    double *stress, *tangent;
    int nh;
    double *hist[2];
    int hw[2];
    unsigned char *flag;
    bool symmetric;  // tangent bitwise symmetric: slot records carry the upper triangle only
};

template <class Launch>
static int run_plastic_host(const PlasticHost &H, size_t n, Launch &&launch)
{
    const size_t d = sizeof(double);
    // H.tangent == nullptr (stress-only call): no tangent slot on the device, none on the wire
    const size_t bpq[6] = {d * 9, d * 6, H.tangent ? d * 36 : 0, d * H.hw[0], H.nh > 1 ? d * H.hw[1] : 0, 1};
    const int wire = n >= 4096 ? effective_wire(H.stress, H.tangent, H.hist[0], H.nh > 1 ? H.hist[1] : nullptr) : 0;
    g_last_wire = wire;
    if (wire) {
        // elastic tangent as the kernel produces it: one virgin point with a zero increment
        // (elastic for any sensible parameter set; otherwise fall through to the plain path)
        PlasticWire W;
        bool elastic0 = false;
        {
            std::lock_guard<std::mutex> lock(g_mu);
            size_t off[6], total = 0;
            for (int a = 0; a < 6; ++a) {
                off[a] = total;
                total += round256(bpq[a] * 128);
            }
            int rc = ensure_ctx(total);
            if (rc != FCX_OK)
                return rc;
            cudaStream_t st = g_ctx.stream[0];
            cudaError_t e = cudaMemsetAsync(g_ctx.buf[0], 0, total, st);
            if (e != cudaSuccess)
                return note_cuda_error(e, "cudaMemsetAsync(template)");
            void *dev[6];
            for (int a = 0; a < 6; ++a)
                dev[a] = g_ctx.buf[0] + off[a];
            rc = launch(dev, 1, st, nullptr);
            if (rc != FCX_OK)
                return rc;
            unsigned char f0 = 1;
            e = H.tangent ? cudaMemcpyAsync(W.tmpl, dev[2], sizeof W.tmpl, cudaMemcpyDeviceToHost, st) : cudaSuccess;
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(&f0, dev[5], 1, cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess)
                e = cudaStreamSynchronize(st);
            if (e != cudaSuccess)
                return note_cuda_error(e, "template download");
            elastic0 = (f0 == 0);
        }
        if (elastic0) {
            size_t chunk = g_chunk_staged < g_chunk ? g_chunk_staged : g_chunk;
            chunk = chunk < n ? chunk : n;
            chunk = (chunk + 127) & ~(size_t)127;
            W.chunk = chunk;
            W.tangent = H.tangent;
            W.nh = H.nh;
            for (int h = 0; h < 2; ++h) {
                W.hist[h] = h < H.nh ? H.hist[h] : nullptr;
                W.hw[h] = h < H.nh ? H.hw[h] : 0;
            }
            W.user_flag = H.flag;
            W.nt = H.tangent == nullptr ? 0 : (H.symmetric ? 21 : 36);
            if (wire == 1 && H.tangent != nullptr && page_locked(H.stress) && page_locked(H.tangent) &&
                page_locked(H.hist[0]) && page_locked(H.nh > 1 ? H.hist[1] : nullptr))
                W.mix = effective_wire_mix();
            g_last_wire_mix = W.mix;
            if (wire >= 2 && H.tangent != nullptr && (reinterpret_cast<uintptr_t>(H.tangent) & 15u) == 0) {
                W.tangent_dev = (double *)device_alias(H.tangent, n * 36 * d);
                if (W.tangent_dev != nullptr && (reinterpret_cast<uintptr_t>(W.tangent_dev) & 15u) == 0)
                    W.nt = 0;
                else
                    W.tangent_dev = nullptr;
            }
            Packer P;
            P.dev_bytes = W.dev_bytes();
            P.wire_bytes = W.wire_bytes();
            P.enqueue = [&W](void **dev, void *scratch, void *wire, size_t q0, size_t cnt, cudaStream_t st) {
                if (W.plain(q0)) {  // this chunk by plain DMA into the page-locked arrays of the caller
                    cudaError_t e = cudaMemcpyAsync(W.tangent + q0 * 36, dev[2], cnt * 36 * sizeof(double),
                                                    cudaMemcpyDeviceToHost, st);
                    for (int h = 0; h < W.nh && e == cudaSuccess; ++h)
                        e = cudaMemcpyAsync(W.hist[h] + q0 * W.hw[h], dev[3 + h], cnt * W.hw[h] * sizeof(double),
                                            cudaMemcpyDeviceToHost, st);
                    if (e == cudaSuccess && W.user_flag != nullptr)  // flags through the slot (the array may be pageable)
                        e = cudaMemcpyAsync((char *)wire + W.off_flag(), dev[5], cnt, cudaMemcpyDeviceToHost, st);
                    return note_cuda_error(e, "plastic wire: plain chunk");
                }
                char *wb = (char *)wire;
                unsigned *list = (unsigned *)scratch, *count = list + W.chunk;
                const unsigned char *fl = (const unsigned char *)dev[5];
                const unsigned cap = (unsigned)sm_count() * 8;
                wire_scan_kernel<<<1, 1024, 0, st>>>(fl, (unsigned)cnt, list, count);
                unsigned long long work = (unsigned long long)cnt * W.rec();
                unsigned grid = (unsigned)((work + 255) / 256);
                wire_pack_kernel<<<grid > cap ? cap : grid, 256, 0, st>>>(
                    list, count, (const double *)dev[2], W.nt, (const double *)dev[3], W.hw[0],
                    (const double *)dev[4], W.hw[1], (double *)(wb + W.off_rec()));
                g_launches.fetch_add(2, std::memory_order_relaxed);
                if (W.tangent_dev != nullptr) {
                    work = (unsigned long long)cnt * 18;
                    grid = (unsigned)((work + 255) / 256);
                    wire_direct_kernel<<<grid > cap ? cap : grid, 256, 0, st>>>(
                        fl, (const double2 *)dev[2], (unsigned)cnt, (double2 *)(W.tangent_dev + q0 * 36));
                    g_launches.fetch_add(1, std::memory_order_relaxed);
                }
                cudaError_t e = cudaGetLastError();
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(wb + W.off_flag(), fl, cnt, cudaMemcpyDeviceToHost, st);
                return note_cuda_error(e, "plastic wire pack");
            };
            P.expand = [&W](size_t q0, size_t cnt, const void *wire, Group &g) {
                if (!W.plain(q0))
                    plastic_wire_expand(W, q0, cnt, wire, g);
                else if (W.user_flag != nullptr)
                    memcpy(W.user_flag + q0, (const char *)wire + W.off_flag(), cnt);
            };
            // tangent, history, flag: uploaded / kept on the device as before, downloaded by the wire
            const HostArr arr[6] = {{H.grad, nullptr, bpq[0]},  {H.stress, H.stress, bpq[1]},
                                    {nullptr, nullptr, bpq[2]}, {H.hist[0], nullptr, bpq[3]},
                                    {H.nh > 1 ? H.hist[1] : nullptr, nullptr, bpq[4]}, {nullptr, nullptr, 1}};
            return run_pipeline(arr, 6, n, launch, &P);
        }
    }
    const HostArr arr[6] = {{H.grad, nullptr, bpq[0]},
                            {H.stress, H.stress, bpq[1]},
                            {nullptr, H.tangent, bpq[2]},
                            {H.hist[0], H.hist[0], bpq[3]},
                            {H.nh > 1 ? H.hist[1] : nullptr, H.nh > 1 ? H.hist[1] : nullptr, bpq[4]},
                            {nullptr, H.flag, 1}};
    return run_pipeline(arr, 6, n, launch);
}

}  // namespace fcx

using namespace fcx;

extern "C" {

size_t fcx_host_chunk_qps(size_t v)
{
    const size_t old = g_chunk;
    if (v > 0) {
        g_chunk = v;
        g_chunk_user = true;
    }
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
        g_wire = on > 3 ? 3 : on;
    return old;
}

int fcx_host_wire_used(void) { return g_last_wire; }

int fcx_host_wire_mix(int percent)
{
    const int old = g_wire_mix;
    if (percent >= -1)
        g_wire_mix = percent > 100 ? 100 : percent;
    return old;
}

int fcx_host_wire_mix_used(void) { return g_last_wire_mix; }

int fcx_host_numa(int on)
{
    const int old = g_numa;
    if (on >= 0)
        g_numa = on ? 1 : 0;
    return old;
}

int fcx_host_numa_info(int *out, int n)
{
    const NumaInfo &N = numa_info();
    const int v[4] = {N.node, N.ncpus, N.allowed, (g_numa && N.valid) ? 1 : 0};
    if (!out)
        return FCX_ERR_NULL;
    for (int i = 0; i < n && i < 4; ++i)
        out[i] = v[i];
    return 4;
}

int fcx_host_slots(int n)
{
    std::lock_guard<std::mutex> lock(g_mu);
    const int old = g_nslot;
    if (n > 0)
        g_nslot = n < 2 ? 2 : (n > NSLOT ? NSLOT : n);
    return old;
}

int fcx_host_debug_skip(int mask)
{
    const int old = g_skip;
    if (mask >= 0)
        g_skip = mask;
    return old;
}

int fcx_host_trace(int on)
{
    const int old = g_trace;
    if (on >= 0)
        g_trace = on ? 1 : 0;
    return old;
}

int fcx_host_timeline(double *out, int max_rows)
{
    const int rows = (int)(g_timeline.size() / TL_COLS);
    if (out)
        for (int r = 0; r < rows && r < max_rows; ++r)
            for (int c = 0; c < TL_COLS; ++c)
                out[r * TL_COLS + c] = g_timeline[(size_t)r * TL_COLS + c];
    return rows;
}

int fcx_host_stats(double *out, int n)
{
    const double v[12] = {g_stats.total_s, g_stats.main_wait_slot_s, g_stats.main_stage_in_s, g_stats.main_enqueue_s,
                          g_stats.drain_event_wait_s, g_stats.drain_expand_s, g_stats.gpu_h2d_s, g_stats.gpu_kernel_s,
                          g_stats.gpu_pack_s, g_stats.gpu_d2h_s, g_stats.chunks, g_stats.chunk_qps};
    if (!out)
        return FCX_ERR_NULL;
    for (int i = 0; i < n && i < 12; ++i)
        out[i] = v[i];
    return 12;
}

/* Host-side roofline probes (bench.py's e2e.roofline): what the host memory system and the PCIe
 * link of THIS box deliver, measured the way the host pipeline uses them. */
int fcx_diag_host_bandwidth(int threads, size_t bytes, double *out, int nout)
{
    if (!out || nout < 3)
        return FCX_ERR_NULL;
    if (threads <= 0)
        threads = pool_threads(true);
    bytes &= ~(size_t)4095;
    if (bytes < ((size_t)1 << 20))
        return FCX_ERR_ARG;
    char *a = (char *)aligned_alloc(4096, bytes), *b = (char *)aligned_alloc(4096, bytes);
    if (!a || !b) {
        free(a);
        free(b);
        return FCX_ERR_ARG;
    }
    Pool &pool = Pool::get();
    pool.ensure(threads);
    const size_t per = ((bytes / threads) + 4095) & ~(size_t)4095;
    auto run = [&](int what) {  // 0 first touch, 1 memcpy b <- a, 2 streaming fill of b, 3 read a
        Group g;
        std::atomic<long long> sink{0};
        const double t0 = now_s();
        for (size_t off = 0; off < bytes; off += per) {
            const size_t len = bytes - off < per ? bytes - off : per;
            g.add();
            pool.submit([=, &g, &sink] {
                if (what == 0) {
                    memset(a + off, 1, len);
                    memset(b + off, 2, len);
                } else if (what == 1) {
                    memcpy(b + off, a + off, len);
                } else if (what == 2) {
                    const __m128d v = _mm_set1_pd(1.5);
                    double *d = (double *)(b + off);
                    for (size_t i = 0; i < len / 8; i += 2)
                        _mm_stream_pd(d + i, v);
                    _mm_sfence();
                } else {
                    const long long *s = (const long long *)(a + off);
                    long long acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
                    for (size_t i = 0; i + 3 < len / 8; i += 4) {
                        acc0 += s[i];
                        acc1 += s[i + 1];
                        acc2 += s[i + 2];
                        acc3 += s[i + 3];
                    }
                    sink.fetch_add(acc0 + acc1 + acc2 + acc3, std::memory_order_relaxed);
                }
                g.done();
            });
        }
        g.wait();
        return now_s() - t0;
    };
    run(0);
    double best[3] = {1e30, 1e30, 1e30};
    for (int rep = 0; rep < 4; ++rep)
        for (int w = 0; w < 3; ++w) {
            const double t = run(w + 1);
            if (t < best[w])
                best[w] = t;
        }
    out[0] = 2.0 * bytes / best[0] / 1e9;  // memcpy: bytes read + bytes written per second
    out[1] = (double)bytes / best[1] / 1e9;  // streaming (non-temporal) fill
    out[2] = (double)bytes / best[2] / 1e9;  // read
    if (nout > 3)
        out[3] = threads;
    free(a);
    free(b);
    return FCX_OK;
}

int fcx_diag_pcie(size_t bytes, double *out, int nout)
{
    if (!out || nout < 4)
        return FCX_ERR_NULL;
    bytes &= ~(size_t)255;
    if (bytes < ((size_t)1 << 20))
        return FCX_ERR_ARG;
    char *h0 = nullptr, *h1 = nullptr, *d0 = nullptr, *d1 = nullptr;
    cudaStream_t s0 = nullptr, s1 = nullptr;
    cudaEvent_t ev[4] = {};
    cudaError_t e = cudaHostAlloc((void **)&h0, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess)
        e = cudaHostAlloc((void **)&h1, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&d0, bytes);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&d1, bytes);
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking);
    if (e == cudaSuccess)
        e = cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
        e = cudaEventCreate(&ev[i]);
    if (e == cudaSuccess) {
        memset(h0, 1, bytes);
        memset(h1, 2, bytes);
        auto timed = [&](bool up, bool down, float *ms_up, float *ms_down) {
            cudaDeviceSynchronize();
            if (up)
                cudaEventRecord(ev[0], s0);
            if (down)
                cudaEventRecord(ev[2], s1);
            for (int rep = 0; rep < 3; ++rep) {
                if (up)
                    cudaMemcpyAsync(d0, h0, bytes, cudaMemcpyHostToDevice, s0);
                if (down)
                    cudaMemcpyAsync(h1, d1, bytes, cudaMemcpyDeviceToHost, s1);
            }
            if (up)
                cudaEventRecord(ev[1], s0);
            if (down)
                cudaEventRecord(ev[3], s1);
            cudaDeviceSynchronize();
            if (up)
                cudaEventElapsedTime(ms_up, ev[0], ev[1]);
            if (down)
                cudaEventElapsedTime(ms_down, ev[2], ev[3]);
        };
        float a = 0, b = 0, c = 0, d = 0;
        timed(true, false, &a, &b);  // warm-up
        timed(true, false, &a, &b);
        timed(false, true, &c, &b);
        timed(true, true, &c, &d);
        const double gb = 3.0 * bytes / 1e9;
        out[0] = gb / (a * 1e-3);  // H2D alone
        out[1] = gb / (b * 1e-3);  // D2H alone
        out[2] = gb / (c * 1e-3);  // H2D while D2H runs
        out[3] = gb / (d * 1e-3);  // D2H while H2D runs
        e = cudaGetLastError();
    }
    for (int i = 0; i < 4; ++i)
        if (ev[i])
            cudaEventDestroy(ev[i]);
    if (s0)
        cudaStreamDestroy(s0);
    if (s1)
        cudaStreamDestroy(s1);
    cudaFree(d0);
    cudaFree(d1);
    if (h0)
        cudaFreeHost(h0);
    if (h1)
        cudaFreeHost(h1);
    return note_cuda_error(e, "fcx_diag_pcie");
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
    if (!D || !grad || !stress)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    const size_t d = sizeof(double);
    HostArr arr[3] = {{grad, nullptr, d * g * g}, {stress, stress, d * s}, {nullptr, tangent, d * s * s}};
    return run_pipeline_const_tangent(arr, 3, 2, s * s, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
        return fcx_elastic_evaluate(constraint, D, cnt, (const double *)dev[0], (double *)dev[1],
                                    tangent ? (double *)dev[2] : nullptr, st);
    });
}

int fcx_mises_evaluate_host(const double *params, size_t n, const double *grad, double *stress,
                            double *tangent, double *eps_n, double *alpha,
                            unsigned char *plastic_flag)
{
    if (n == 0)
        return FCX_OK;
    if (!params || !grad || !stress || !eps_n || !alpha)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    const PlasticHost H{grad, stress, tangent, 2, {eps_n, alpha}, {6, 1}, plastic_flag, true};
    return run_plastic_host(H, n, [&](void **dev, size_t cnt, cudaStream_t st, int *status) {
        return fcx_mises_evaluate(params, cnt, (const double *)dev[0], (double *)dev[1],
                                  tangent ? (double *)dev[2] : nullptr, (double *)dev[3], (double *)dev[4],
                                  FCX_LAYOUT_AOS, (unsigned char *)dev[5], status, st);
    });
}

int fcx_mises_linear_hardening_evaluate_host(const double *params, size_t n, const double *grad,
                                             double *stress, double *tangent, double *history,
                                             unsigned char *plastic_flag)
{
    if (n == 0)
        return FCX_OK;
    if (!params || !grad || !stress || !history)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    // full 36-entry records on the slot wire: kappa*1(x)1 + c*P_dev + c'*n n^T is symmetric too, but
    // nothing pins that bit for bit for this model, so the triangle shortcut is not taken
    const PlasticHost H{grad, stress, tangent, 1, {history, nullptr}, {7, 0}, plastic_flag, false};
    return run_plastic_host(H, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
        return fcx_mises_linear_hardening_evaluate(params, cnt, (const double *)dev[0], (double *)dev[1],
                                                   tangent ? (double *)dev[2] : nullptr, (double *)dev[3],
                                                   (unsigned char *)dev[5], st);
    });
}

int fcx_drucker_prager_evaluate_host(int hyperbolic, const double *params, size_t n,
                                     const double *grad, double *stress, double *tangent,
                                     double *history, unsigned char *plastic_flag)
{
    if (n == 0)
        return FCX_OK;
    if (!params || !grad || !stress || !history)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    // non-associated flow makes the tangent non-symmetric: full records
    const PlasticHost H{grad, stress, tangent, 1, {history, nullptr}, {7, 0}, plastic_flag, false};
    return run_plastic_host(H, n, [&](void **dev, size_t cnt, cudaStream_t st, int *status) {
        return fcx_drucker_prager_evaluate(hyperbolic, params, cnt, (const double *)dev[0],
                                           (double *)dev[1], tangent ? (double *)dev[2] : nullptr, (double *)dev[3],
                                           (unsigned char *)dev[5], status, st);
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
    if (!D0 || !I2 || !grad || !stress || !ev || !et)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    const size_t d = sizeof(double);
    HostArr arr[5] = {{grad, nullptr, d * g * g}, {stress, stress, d * s},
                      {nullptr, tangent, d * s * s}, {ev, ev, d * s}, {et, et, d * s}};
    return run_pipeline_const_tangent(arr, 5, 2, s * s, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
        return fcx_kelvin_evaluate(constraint, D0, I2, mu0, lam0, mu1, tau, del_t, cnt,
                                   (const double *)dev[0], (double *)dev[1], tangent ? (double *)dev[2] : nullptr,
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
    if (!D0 || !D1 || !grad || !stress || !ev || !et)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    const size_t d = sizeof(double);
    HostArr arr[5] = {{grad, nullptr, d * g * g}, {stress, stress, d * s},
                      {nullptr, tangent, d * s * s}, {ev, ev, d * s}, {et, et, d * s}};
    return run_pipeline_const_tangent(arr, 5, 2, s * s, n, [&](void **dev, size_t cnt, cudaStream_t st, int *) {
        return fcx_maxwell_evaluate(constraint, D0, D1, mu1, tau, del_t, cnt,
                                    (const double *)dev[0], (double *)dev[1], tangent ? (double *)dev[2] : nullptr,
                                    (double *)dev[3], (double *)dev[4], st);
    });
}

}  // extern "C"
