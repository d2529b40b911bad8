// fcx_host.cu -- HOST-pointer entry points (include/fcx.h, "host" family).
//
// This is the path a CPU-side caller takes: dolfinx hands `law.evaluate` host
// numpy arrays (reference solver/_lawonsubmesh.py:86-94).  The QP axis is cut
// into chunks; chunk c runs entirely on stream c % NSLOT:
//     H2D(inputs of c) -> kernel(c) -> D2H(outputs of c)
// so the upload of chunk c+1, the kernel of chunk c and the download of chunk
// c-1 overlap on the two copy engines and the SMs.  Stream order alone makes
// slot reuse safe (no events needed).  With page-locked arrays
// (fcx_host_register, or pinned allocations) the copies are true async DMA;
// pageable arrays still work, staged by the driver.
#include <cuda_runtime.h>

#include <climits>
#include <cstdint>
#include <mutex>

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
    int *status = nullptr;
    bool ready = false;
};

static HostCtx g_ctx;
static std::mutex g_mu;
static size_t g_chunk = (size_t)1 << 18;

static int ensure_ctx(size_t need)
{
    if (!g_ctx.ready) {
        for (int s = 0; s < NSLOT; ++s) {
            cudaError_t e = cudaStreamCreateWithFlags(&g_ctx.stream[s], cudaStreamNonBlocking);
            if (e != cudaSuccess)
                return note_cuda_error(e, "cudaStreamCreate");
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

// launch(dev_ptrs, q_count, stream, status_dev) enqueues the kernel for one chunk.
template <class Launch>
static int run_pipeline(const HostArr *arr, int narr, size_t n, Launch &&launch)
{
    if (n == 0)
        return FCX_OK;
    std::lock_guard<std::mutex> lock(g_mu);
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

void fcx_host_release(void)
{
    std::lock_guard<std::mutex> lock(g_mu);
    for (int s = 0; s < NSLOT; ++s) {
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
