// fcx_api.cu -- device-pointer entry points of the C ABI (include/fcx.h):
// parameter set-up on the host (same expressions as the reference's __init__ /
// evaluate preambles), kernel selection by constraint, persistent-grid launch.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/fcx.h"
#include "fcx_internal.h"
#include "fcx_mises_form.cuh"
#include "fcx_mises_ostage.cuh"
#include "fcx_models.cuh"

namespace fcx {

thread_local char g_cuda_error[256] = "";
std::atomic<unsigned long long> g_launches{0};

int note_cuda_error(cudaError_t e, const char *where)
{
    if (e == cudaSuccess)
        return FCX_OK;
    snprintf(g_cuda_error, sizeof g_cuda_error, "%s: %s", where, cudaGetErrorString(e));
    return FCX_ERR_CUDA;
}

static int g_sm_count[FCX_MAX_DEVICES] = {};  // per device index, 0 = not queried yet
int sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= FCX_MAX_DEVICES)
        return 148;  // B200
    if (g_sm_count[dev] == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_sm_count[dev] = n;
        else
            return 148;
    }
    return g_sm_count[dev];
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Tunables (fcx_tune): CTAs per SM for the persistent grid; 0 = occupancy query.
static int g_ctas_per_sm = 0;
// Newton iteration cap of the Mises return mapping (reference nmax = 100,
// mises_plasticity_isotropic_hardening.py:106); lowered only by tests that
// exercise the non-convergence reporting.
static int g_mises_nmax = 100;

// Tunables (fcx_tune): QPs per tile (64/128/256) and L2 cache-hint flags.
static int g_tile = 128;
// Mises AoS kernel: 1 = output-staged (tangent through shared memory + bulk
// store, fcx_mises_ostage.cuh), 0 = generic tile pipeline with the
// warp-cooperative tangent store.
static int g_mises_variant = 1;
static int g_mises_tile = 64;  // QPs per tile of the output-staged kernel (64 or 128)
// Stress-only Mises calls (tangent == NULL): 0 = double-buffered tile pipeline, stress-only instantiation
// (fcx_tile_kernel<MisesModel, 128, false>), 1 = single-stage output-staged kernel without the tangent
// region.  Measured on B200, 16 M QPs (profiles/r2a_tune_stress_only.jsonl): 0.75 vs 0.80 ms.
static int g_mises_so_variant = 0;
// Drucker-Prager kernels: 1 = slow fp64 operations spelled for latency (DruckerPragerModel VAR 1, the
// default), 0 = the reference's spelling (VAR 0)
static int g_dp_variant = 1;
// Hand tiles out through an atomic ticket counter (all tile kernels).
static int g_dynamic_tiles = 1;
static int g_fem_variant = 1;
// 1 = nodal values staged by cp.async one tile ahead (gather_staged_kernel, the default), 0 = register loads
// (gather_kernel), 2 = thread per cell with the basis-gradient table in the constant bank (gather_cell_kernel,
// 3-D 4-point rules): bit-identical, but measured SLOWER on B200 (0.139 vs 0.104 ms for 998 250 P2 tets,
// profiles/r2c_gather_ab.jsonl): its 1.1 KB of shared memory per cell caps an SM at 6 warps (ncu: 9 % occupancy,
// issue slots 34 %, every stall a fixed-latency wait with nothing else to issue); 3 = warp-uniform quadrature-point
// pair with the table in the constant bank at the staged kernel's occupancy (gather_wq_kernel): 0.127 ms, slower too
// (profiles/r2h_gather_ab.jsonl).
static int g_gather_variant = 1;
static int g_hints = 8;  // bit1: evict_first on bulk loads, bit2: on bulk stores,
                         // bit3: constant tangents written by bulk stores from shared memory

// Atomic tile-ticket counters: one per (device, stream) so that launches that
// may run concurrently (different streams) never share a counter; launches on
// one stream serialise, and the counter is zeroed on that stream before each.
static unsigned long long *ticket_for(cudaStream_t stream)
{
    struct Slot {
        int dev;
        cudaStream_t st;
        unsigned long long *ptr;
    };
    static Slot slots[256];
    static int nslots = 0;
    static std::mutex mu;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
        return nullptr;
    unsigned long long *ptr = nullptr;
    {
        std::lock_guard<std::mutex> lock(mu);
        for (int i = 0; i < nslots; ++i)
            if (slots[i].dev == dev && slots[i].st == stream)
                ptr = slots[i].ptr;
        if (ptr == nullptr && nslots < 256) {
            void *p = nullptr;
            if (cudaMalloc(&p, 128) != cudaSuccess)
                return nullptr;
            ptr = static_cast<unsigned long long *>(p);
            slots[nslots++] = Slot{dev, stream, ptr};
        }
    }
    if (ptr != nullptr && cudaMemsetAsync(ptr, 0, sizeof(unsigned long long), stream) != cudaSuccess)
        return nullptr;
    return ptr;  // nullptr -> static tile stride
}

unsigned long long *tile_ticket(cudaStream_t stream)
{
    return g_dynamic_tiles ? ticket_for(stream) : nullptr;
}
int tuned_ctas_per_sm() { return g_ctas_per_sm; }
int fem_variant() { return g_fem_variant; }
int gather_variant() { return g_gather_variant; }

template <class M, int TILE, bool WT = true>
static int launch_tile_t(const typename M::Params &prm, const SegPtrs<M::nseg()> &io,
                         double *tangent, size_t n, bool bulk_ok, unsigned char *flag,
                         int *status, cudaStream_t stream, unsigned long long qbase)
{
    auto kern = fcx_tile_kernel<M, TILE, WT>;
    constexpr size_t smem = tile_smem_bytes<M, TILE, WT>();
    static OccCache cache;  // per instantiation and device
    int occ = 1;
    if (int rc = kernel_occupancy(cache, kern, TILE, smem, "occupancy(fcx_tile_kernel)", &occ))
        return rc;
    const int per_sm = g_ctas_per_sm > 0 ? g_ctas_per_sm : occ;
    const unsigned long long ntiles = (n + TILE - 1) / TILE;
    unsigned long long grid = (unsigned long long)sm_count() * per_sm;
    if (grid > ntiles)
        grid = ntiles;
    const int flags = (bulk_ok ? 1 : 0) | (g_hints & 14);
    unsigned long long *ticket = (g_dynamic_tiles && ntiles > grid) ? ticket_for(stream) : nullptr;
    kern<<<(unsigned)grid, TILE, smem, stream>>>(prm, io, tangent, (unsigned long long)n, flags,
                                                 flag, status, qbase, ticket);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "fcx_tile_kernel launch");
}

template <class M>
static int launch_tile(const typename M::Params &prm, const SegPtrs<M::nseg()> &io,
                       double *tangent, size_t n, bool bulk_ok, unsigned char *flag, int *status,
                       cudaStream_t stream, unsigned long long qbase = 0)
{
    if (n == 0)
        return FCX_OK;
    switch (g_tile) {
    case 64: return launch_tile_t<M, 64>(prm, io, tangent, n, bulk_ok, flag, status, stream, qbase);
    case 256: return launch_tile_t<M, 256>(prm, io, tangent, n, bulk_ok, flag, status, stream, qbase);
    default:
        // stress-only calls (tangent == NULL) have their own instantiation at the default tile size;
        // the other tile sizes (tuning only) take the run-time branch of the full kernel
        if (tangent == nullptr)
            return launch_tile_t<M, 128, false>(prm, io, nullptr, n, bulk_ok, flag, status, stream, qbase);
        return launch_tile_t<M, 128>(prm, io, tangent, n, bulk_ok, flag, status, stream, qbase);
    }
}

// Output-staged Mises kernel (fcx_mises_ostage.cuh) over `ntiles` full tiles.
template <int TILE, int MINCTAS, bool WITH_TANGENT>
static int launch_mises_ostage(const MisesParams &P, const double *grad, double *stress,
                               double *tangent, double *eps_n, double *alpha,
                               unsigned long long ntiles, unsigned char *flag, int *status,
                               cudaStream_t stream)
{
    auto kern = fcx_mises_ostage_kernel<TILE, MINCTAS, WITH_TANGENT>;
    constexpr size_t smem = mises_ostage_smem_bytes<TILE, WITH_TANGENT>();
    static OccCache cache;  // per instantiation and device
    int occ = 1;
    if (int rc = kernel_occupancy(cache, kern, TILE, smem, "occupancy(mises_ostage)", &occ))
        return rc;
    const int per_sm = g_ctas_per_sm > 0 ? g_ctas_per_sm : occ;
    unsigned long long grid = (unsigned long long)sm_count() * per_sm;
    if (grid > ntiles)
        grid = ntiles;
    unsigned long long *ticket = (g_dynamic_tiles && ntiles > grid) ? ticket_for(stream) : nullptr;
    kern<<<(unsigned)grid, TILE, smem, stream>>>(P, grad, stress, tangent, eps_n, alpha, ntiles,
                                                 flag, status, ticket);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "fcx_mises_ostage_kernel launch");
}

// Fused form() pipeline for VonMises3D (fcx_mises_form.cuh).
template <int ND, int NQ>
static int launch_mises_form(const MisesParams &P, MisesFormArgs A, cudaStream_t stream)
{
    constexpr int TILE = 64;
    auto kern = fcx_mises_form_kernel<ND, NQ, TILE, 8>;
    constexpr size_t smem = mises_form_smem_bytes<ND, NQ, TILE>();
    static OccCache cache;  // per instantiation and device
    int occ = 1;
    if (int rc = kernel_occupancy(cache, kern, TILE, smem, "occupancy(mises_form)", &occ))
        return rc;
    constexpr int CPT = TILE / NQ;
    const unsigned long long ntiles = (A.ncells + CPT - 1) / CPT;
    const int per_sm = g_ctas_per_sm > 0 ? g_ctas_per_sm : occ;
    unsigned long long grid = (unsigned long long)sm_count() * per_sm;
    if (grid > ntiles)
        grid = ntiles;
    A.ticket = (g_dynamic_tiles && ntiles > grid) ? ticket_for(stream) : nullptr;
    kern<<<(unsigned)grid, TILE, smem, stream>>>(P, A);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "fcx_mises_form_kernel launch");
}

template <class M>
static int launch_uniaxial(const typename M::Params &prm, const SegPtrs<M::nseg()> &io,
                           double *tangent, size_t n, bool vec_ok, cudaStream_t stream)
{
    if (n == 0)
        return FCX_OK;
    const unsigned long long work = (n + 1) / 2;
    unsigned long long grid = (work + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (grid > cap)
        grid = cap;
    fcx_uniaxial_kernel<M><<<(unsigned)grid, 256, 0, stream>>>(prm, io, tangent,
                                                              (unsigned long long)n, vec_ok ? 1 : 0);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "fcx_uniaxial_kernel launch");
}

// ---- host-side parameter builders (shared with fcx_host.cu) ----------------

// fc/models/spring_kelvin_model.py:73-85
template <int S, int G>
static void kelvin_params(typename KelvinModel<S, G>::Params &P, const double *D0,
                          const double *I2, double mu0, double lam0, double mu1, double tau,
                          double del_t)
{
    const double factor = 1 / del_t + 1 / tau + mu0 / (tau * mu1);
    P.inv_factor = 1 / factor;
    P.c_sig = 1 / (tau * 2 * mu1);
    P.c_ev = 1 / tau;
    P.c_e = mu0 / (tau * mu1);
    P.c_tr = lam0 / (tau * 2 * mu1);
    P.two_mu0 = 2 * mu0;
    const double dscale = 1 - mu0 / (tau * mu1 * factor);
    for (int k = 0; k < S * S; ++k) {
        P.D0[k] = D0[k];
        P.Dt[k] = dscale * D0[k];
    }
    for (int k = 0; k < S; ++k)
        P.I2[k] = I2[k];
}

// fc/models/spring_maxwell_model.py:70-82
template <int S, int G>
static void maxwell_params(typename MaxwellModel<S, G>::Params &P, const double *D0,
                           const double *D1, double mu1, double tau, double del_t)
{
    const double factor = 1 / del_t + 1 / tau;
    P.inv_factor = 1 / factor;
    P.c_tot = 1 / (tau * 2 * mu1);
    P.c_ev = 1 / tau;
    P.two_mu1 = 2 * mu1;
    const double dscale = 1 - 1 / (tau * factor);
    for (int k = 0; k < S * S; ++k) {
        P.D1[k] = D1[k];
        P.D01[k] = D0[k] + D1[k];
        P.Dt[k] = D0[k] + dscale * D1[k];
    }
}

template <int S, int G>
static int elastic_dispatch(const double *D, size_t n, const double *grad, double *stress,
                            double *tangent, cudaStream_t st)
{
    using M = ElasticModel<S, G>;
    typename M::Params P;
    for (int k = 0; k < S * S; ++k)
        P.D[k] = D[k];
    SegPtrs<2> io{{const_cast<double *>(grad), stress}};
    const bool al = aligned16(grad) && aligned16(stress) && aligned16(tangent);
    if constexpr (S == 1)
        return launch_uniaxial<M>(P, io, tangent, n, al, st);
    else
        return launch_tile<M>(P, io, tangent, n, al, nullptr, nullptr, st);
}

template <int S, int G>
static int kelvin_dispatch(const double *D0, const double *I2, double mu0, double lam0, double mu1,
                           double tau, double del_t, size_t n, const double *grad, double *stress,
                           double *tangent, double *ev, double *et, cudaStream_t st)
{
    using M = KelvinModel<S, G>;
    typename M::Params P;
    kelvin_params<S, G>(P, D0, I2, mu0, lam0, mu1, tau, del_t);
    SegPtrs<4> io{{const_cast<double *>(grad), stress, ev, et}};
    const bool al = aligned16(grad) && aligned16(stress) && aligned16(tangent) && aligned16(ev) &&
                    aligned16(et);
    if constexpr (S == 1)
        return launch_uniaxial<M>(P, io, tangent, n, al, st);
    else
        return launch_tile<M>(P, io, tangent, n, al, nullptr, nullptr, st);
}

template <int S, int G>
static int maxwell_dispatch(const double *D0, const double *D1, double mu1, double tau,
                            double del_t, size_t n, const double *grad, double *stress,
                            double *tangent, double *ev, double *et, cudaStream_t st)
{
    using M = MaxwellModel<S, G>;
    typename M::Params P;
    maxwell_params<S, G>(P, D0, D1, mu1, tau, del_t);
    SegPtrs<4> io{{const_cast<double *>(grad), stress, ev, et}};
    const bool al = aligned16(grad) && aligned16(stress) && aligned16(tangent) && aligned16(ev) &&
                    aligned16(et);
    if constexpr (S == 1)
        return launch_uniaxial<M>(P, io, tangent, n, al, st);
    else
        return launch_tile<M>(P, io, tangent, n, al, nullptr, nullptr, st);
}

// strain_from_grad_u as a standalone op (fc/models/utils.py:132-208)
template <int S, int G>
__global__ void strain_kernel(const double *__restrict__ grad, double *__restrict__ strain,
                              unsigned long long n)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < n;
         q += stride) {
        double g[G * G], e[S];
#pragma unroll
        for (int i = 0; i < G * G; ++i)
            g[i] = grad[q * (G * G) + i];
        mandel_strain<S, G>(g, e);
#pragma unroll
        for (int i = 0; i < S; ++i)
            strain[q * S + i] = e[i];
    }
}


// 3D -> 1D/2D adapters (fc/models/utils.py:211-412: UniaxialStrainFrom3D, PlaneStrainFrom3D):
// embed the low-dimensional grad_del_u / stress into persistent 3D scratch arrays (only the
// mapped components are overwritten, :285-292 / :365-386), and extract the mapped components
// of the 3D stress / tangent afterwards (:293-297 / :388-412).
template <int G, int S>
__global__ void embed3d_kernel(const double *__restrict__ grad, const double *__restrict__ stress,
                               double *__restrict__ grad3, double *__restrict__ stress3,
                               unsigned long long n)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < n;
         q += stride) {
        if (G == 1) {
            grad3[q * 9] = grad[q];      // :285-287
            stress3[q * 6] = stress[q];  // :289-292
        } else {
            grad3[q * 9 + 0] = grad[q * 4 + 0];  // :375-376
            grad3[q * 9 + 1] = grad[q * 4 + 1];
            grad3[q * 9 + 3] = grad[q * 4 + 2];
            grad3[q * 9 + 4] = grad[q * 4 + 3];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                stress3[q * 6 + k] = stress[q * 4 + k];  // :384-386
        }
    }
}

template <int G, int S>
__global__ void extract3d_kernel(const double *__restrict__ stress3,
                                 const double *__restrict__ tangent3, double *__restrict__ stress,
                                 double *__restrict__ tangent, unsigned long long n)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long q = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; q < n;
         q += stride) {
#pragma unroll
        for (int i = 0; i < S; ++i) {
            stress[q * S + i] = stress3[q * 6 + i];  // :293-297 / :388-391
#pragma unroll
            for (int j = 0; j < S; ++j)
                tangent[q * S * S + i * S + j] = tangent3[q * 36 + i * 6 + j];  // :299-302 / :393-412
        }
    }
}

// Diagnostic: the simplest possible streaming kernel with the Mises kernel's
// read:write byte mix (per pair of QPs: 22 double2 read, 49 double2 written),
// perfectly coalesced, no shared memory, no arithmetic to speak of.  Its
// throughput is the practical DRAM ceiling for this mix on the box at hand.
template <int NR, int NW>
__global__ void __launch_bounds__(256)
    diag_stream_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst,
                       unsigned long long iters)
{
    const unsigned long long T = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long gt = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned long long it = 0; it < iters; ++it) {
        double acc = 0.0;
        double2 v[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k)
            v[k] = src[(it * NR + k) * T + gt];
#pragma unroll
        for (int k = 0; k < NR; ++k)
            acc += v[k].x + v[k].y;
#pragma unroll
        for (int k = 0; k < NW; ++k)
            st_stream_v2(reinterpret_cast<double *>(dst + (it * NW + k) * T + gt), acc, acc + k);
    }
}

// Diagnostic: fp64 FMA peak of the device -- 8 independent DFMA chains per thread, no memory traffic.
// BASELINE.md section 2 asks for this number next to any fp64-pipe utilisation that is quoted.
__global__ void __launch_bounds__(256) diag_dfma_kernel(double *out, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll 8
        for (int k = 0; k < 8; ++k) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456)  // never true: keeps the chains alive
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace fcx

using namespace fcx;

extern "C" {

/* DFMA-chain peak of the current device in TFLOP/s (2 flops per DFMA); < 0 on error. */
double fcx_diag_dfma_peak(void)
{
    const int iters = 4096, blocks = sm_count() * 8;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * 256 * blocks) != cudaSuccess)
        return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    diag_dfma_kernel<<<blocks, 256>>>(out, 64, 0.999999, 1e-9);  // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        diag_dfma_kernel<<<blocks, 256>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    g_launches.fetch_add(4, std::memory_order_relaxed);
    const cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    if (e != cudaSuccess)
        return -1.0;
    const double dfma = (double)blocks * 256 * iters * 64;
    return 2.0 * dfma / (best * 1e-3) / 1e12;
}

int fcx_version(void) { return FCX_VERSION; }

const char *fcx_strerror(int code)
{
    switch (code) {
    case FCX_OK: return "ok";
    case FCX_ERR_CONSTRAINT: return "unknown or unsupported stress-strain constraint";
    case FCX_ERR_TIMESTEP: return "Time step must be defined and positive.";
    case FCX_ERR_NULL: return "required pointer is NULL";
    case FCX_ERR_CUDA: return "CUDA runtime error (see fcx_last_cuda_error)";
    case FCX_ERR_ARG: return "invalid argument";
    default:
        return code > 0 ? "Newton-Raphson method did not converge for plastic multiplier."
                        : "unknown error";
    }
}

const char *fcx_last_cuda_error(void) { return g_cuda_error; }

int fcx_stress_strain_dim(int c) { return (c < 1 || c > 5) ? -1 : (c <= 2 ? 1 : (c <= 4 ? 4 : 6)); }
int fcx_geometric_dim(int c) { return (c < 1 || c > 5) ? -1 : (c <= 2 ? 1 : (c <= 4 ? 2 : 3)); }

unsigned long long fcx_launch_count(void) { return g_launches.load(); }

int fcx_tune(const char *key, int value)
{
    if (key && strcmp(key, "ctas_per_sm") == 0) {
        const int old = g_ctas_per_sm;
        g_ctas_per_sm = value;
        return old;
    }
    if (key && strcmp(key, "fem_variant") == 0) {
        const int old = g_fem_variant;
        if (value >= 0)  // negative = query
            g_fem_variant = value;
        return old;
    }
    if (key && strcmp(key, "gather_variant") == 0) {
        const int old = g_gather_variant;
        if (value >= 0)  // negative = query
            g_gather_variant = value;
        return old;
    }
    if (key && strcmp(key, "tile") == 0) {
        if (value != 64 && value != 128 && value != 256)
            return FCX_ERR_ARG;
        const int old = g_tile;
        g_tile = value;
        return old;
    }
    if (key && strcmp(key, "l2_hints") == 0) {
        const int old = g_hints;
        g_hints = value;
        return old;
    }
    if (key && strcmp(key, "mises_tile") == 0) {
        if (value != 64 && value != 128)
            return FCX_ERR_ARG;
        const int old = g_mises_tile;
        g_mises_tile = value;
        return old;
    }
    if (key && strcmp(key, "mises_variant") == 0) {
        if (value != 0 && value != 1)
            return FCX_ERR_ARG;
        const int old = g_mises_variant;
        g_mises_variant = value;
        return old;
    }
    if (key && strcmp(key, "dp_variant") == 0) {
        const int old = g_dp_variant;
        if (value == 0 || value == 1)
            g_dp_variant = value;
        return old;
    }
    if (key && strcmp(key, "mises_so_variant") == 0) {
        if (value != 0 && value != 1)
            return FCX_ERR_ARG;
        const int old = g_mises_so_variant;
        g_mises_so_variant = value;
        return old;
    }
    if (key && strcmp(key, "dynamic_tiles") == 0) {
        const int old = g_dynamic_tiles;
        g_dynamic_tiles = value ? 1 : 0;
        return old;
    }
    if (key && strcmp(key, "mises_nmax") == 0) {
        const int old = g_mises_nmax;
        g_mises_nmax = value;
        return old;
    }
    return FCX_ERR_ARG;
}

/* Diagnostic stream with the Mises byte mix; n_qps is rounded down to a whole
 * number of grid iterations.  Returns the number of QP-equivalents processed. */
long long fcx_diag_stream_mix(const double *src, double *dst, size_t n_qps, void *stream)
{
    const unsigned long long threads = (unsigned long long)sm_count() * 8 * 256;
    const unsigned long long iters = (n_qps / 2) / threads;
    if (iters == 0)
        return 0;
    diag_stream_kernel<22, 49><<<sm_count() * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const double2 *>(src), reinterpret_cast<double2 *>(dst), iters);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (note_cuda_error(cudaGetLastError(), "diag_stream_kernel launch") != FCX_OK)
        return FCX_ERR_CUDA;
    return (long long)(iters * threads * 2);
}

int fcx_set_device(int device)
{
    return note_cuda_error(cudaSetDevice(device), "cudaSetDevice");
}

int fcx_elastic_evaluate(int constraint, const double *D, size_t n, const double *grad,
                         double *stress, double *tangent, void *stream)
{
    if (n == 0)
        return FCX_OK;
    if (!D || !grad || !stress)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (constraint) {
    case FCX_UNIAXIAL_STRAIN:
    case FCX_UNIAXIAL_STRESS: return elastic_dispatch<1, 1>(D, n, grad, stress, tangent, st);
    case FCX_PLANE_STRAIN:
    case FCX_PLANE_STRESS: return elastic_dispatch<4, 2>(D, n, grad, stress, tangent, st);
    case FCX_FULL: return elastic_dispatch<6, 3>(D, n, grad, stress, tangent, st);
    default: return FCX_ERR_CONSTRAINT;
    }
}

int fcx_mises_evaluate(const double *params, size_t n, const double *grad, double *stress,
                       double *tangent, double *eps_n, double *alpha, int eps_layout,
                       unsigned char *plastic_flag, int *status, void *stream)
{
    if (n == 0)
        return FCX_OK;
    if (!params || !grad || !stress || !eps_n || !alpha)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MisesParams P{params[0], params[1], params[2], params[3], params[4], g_mises_nmax};
    SegPtrs<4> io{{const_cast<double *>(grad), stress, eps_n, alpha}};
    bool al = aligned16(grad) && aligned16(stress) && aligned16(tangent) && aligned16(eps_n) &&
              aligned16(alpha);
    if (eps_layout == FCX_LAYOUT_SOA) {
        al = al && (n % 2 == 0);  // plane starts must stay 16-byte aligned
        return launch_tile<MisesModel<true>>(P, io, tangent, n, al, plastic_flag, status, st);
    }
    if (eps_layout != FCX_LAYOUT_AOS)
        return FCX_ERR_ARG;
    if (tangent == nullptr && g_mises_so_variant == 0)
        return launch_tile<MisesModel<false>>(P, io, nullptr, n, al, plastic_flag, status, st);
    if (g_mises_variant == 1 && al) {
        // full tiles through the output-staged kernel, the tail (< one tile)
        // through the generic pipeline
        const size_t T = (size_t)g_mises_tile;
        const unsigned long long ntiles = n / T;
        const size_t nfull = (size_t)ntiles * T;
        int rc = FCX_OK;
        if (ntiles > 0 && tangent != nullptr)
            rc = (g_mises_tile == 64)
                     ? launch_mises_ostage<64, 8, true>(P, grad, stress, tangent, eps_n, alpha, ntiles,
                                                        plastic_flag, status, st)
                     : launch_mises_ostage<128, 4, true>(P, grad, stress, tangent, eps_n, alpha, ntiles,
                                                         plastic_flag, status, st);
        else if (ntiles > 0)  // stress-only: no tangent staging, no tangent store (280 B/QP)
            rc = (g_mises_tile == 64)
                     ? launch_mises_ostage<64, 8, false>(P, grad, stress, nullptr, eps_n, alpha, ntiles,
                                                         plastic_flag, status, st)
                     : launch_mises_ostage<128, 4, false>(P, grad, stress, nullptr, eps_n, alpha, ntiles,
                                                          plastic_flag, status, st);
        if (rc != FCX_OK || nfull == n)
            return rc;
        SegPtrs<4> tail{{const_cast<double *>(grad) + nfull * 9, stress + nfull * 6,
                         eps_n + nfull * 6, alpha + nfull}};
        return launch_tile<MisesModel<false>>(P, tail, tangent ? tangent + nfull * 36 : nullptr, n - nfull, al,
                                              plastic_flag ? plastic_flag + nfull : nullptr,
                                              status, st, nfull);
    }
    return launch_tile<MisesModel<false>>(P, io, tangent, n, al, plastic_flag, status, st);
}

int fcx_mises_linear_hardening_evaluate(const double *params, size_t n, const double *grad,
                                        double *stress, double *tangent, double *history,
                                        unsigned char *plastic_flag, void *stream)
{
    if (n == 0)
        return FCX_OK;
    if (!params || !grad || !stress || !history)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    MisesLinParams P{params[0], params[1], params[2], params[3]};
    SegPtrs<3> io{{const_cast<double *>(grad), stress, history}};
    const bool al = aligned16(grad) && aligned16(stress) && aligned16(tangent) && aligned16(history);
    return launch_tile<MisesLinModel>(P, io, tangent, n, al, plastic_flag, nullptr,
                                      static_cast<cudaStream_t>(stream));
}

int fcx_drucker_prager_evaluate(int hyperbolic, const double *params, size_t n, const double *grad,
                                double *stress, double *tangent, double *history,
                                unsigned char *plastic_flag, int *status, void *stream)
{
    if (n == 0)
        return FCX_OK;
    if (!params || !grad || !stress || !history)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    DruckerPragerParams P;
    P.mu = params[0];
    P.kappa = params[1];
    P.a = params[2];
    P.b = params[3];
    P.d2 = hyperbolic ? params[4] * params[4] : 0.0;  // d.powi(2)
    P.b_flow = hyperbolic ? params[5] : params[4];
    P.apex = P.a / P.b;
    SegPtrs<3> io{{const_cast<double *>(grad), stress, history}};
    const bool al = aligned16(grad) && aligned16(stress) && aligned16(tangent) && aligned16(history);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (g_dp_variant == 0) {  // the reference's spelling of every division / square root (A/B only)
        if (hyperbolic)
            return launch_tile<DruckerPragerModel<true, 0>>(P, io, tangent, n, al, plastic_flag, status, st);
        return launch_tile<DruckerPragerModel<false, 0>>(P, io, tangent, n, al, plastic_flag, status, st);
    }
    if (hyperbolic)
        return launch_tile<DruckerPragerModel<true, 1>>(P, io, tangent, n, al, plastic_flag, status, st);
    return launch_tile<DruckerPragerModel<false, 1>>(P, io, tangent, n, al, plastic_flag, status, st);
}

int fcx_mises_form(const double *params, size_t ncells, const int *cells, int nq, int nd, const int *dofmap,
                   const double *u, const double *u_prev, const double *dphi_ref,
                   const double *Jinv, const double *stress_prev, double *stress_cur,
                   double *tangent, const double *eps_n0, double *eps_n1, const double *alpha0,
                   double *alpha1, double *grad_out, double *tangent_rec,
                   unsigned char *plastic_flag, int *status, void *stream)
{
    if (ncells == 0)
        return FCX_OK;
    if (tangent_rec && !aligned16(tangent_rec))
        return FCX_ERR_ARG;
    if (!params || !dofmap || !u || !dphi_ref || !Jinv || !stress_prev || !stress_cur ||
        !eps_n0 || !eps_n1 || !alpha0 || !alpha1)  // tangent == NULL: stress-only
        return FCX_ERR_NULL;
    MisesParams P{params[0], params[1], params[2], params[3], params[4], g_mises_nmax};
    MisesFormArgs A;
    A.cells = cells;
    A.dofmap = dofmap;
    A.u = u;
    A.u_prev = u_prev;
    A.dphi_ref = dphi_ref;
    A.Jinv = Jinv;
    A.stress_prev = stress_prev;
    A.eps0 = eps_n0;
    A.alpha0 = alpha0;
    A.stress_cur = stress_cur;
    A.tangent = tangent;
    A.eps1 = eps_n1;
    A.alpha1 = alpha1;
    A.grad_out = grad_out;
    A.trec = tangent_rec;
    A.flag = plastic_flag;
    A.status = status;
    A.ticket = nullptr;
    A.ncells = ncells;
    A.bulk_ok = aligned16(stress_prev) && aligned16(stress_cur) && aligned16(tangent) &&
                aligned16(eps_n0) && aligned16(eps_n1) && aligned16(alpha0) && aligned16(alpha1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (nd == 10 && nq == 4)
        return launch_mises_form<10, 4>(P, A, st);
    if (nd == 4 && nq == 1)
        return launch_mises_form<4, 1>(P, A, st);
    if (nd == 4 && nq == 4)
        return launch_mises_form<4, 4>(P, A, st);
    return FCX_ERR_ARG;  // no fused specialisation: call fcx_gather_grad + fcx_mises_evaluate
}

int fcx_kelvin_evaluate(int constraint, const double *D0, const double *I2, double mu0,
                        double lam0, double mu1, double tau, double del_t, size_t n,
                        const double *grad, double *stress, double *tangent, double *ev,
                        double *et, void *stream)
{
    if (!(del_t > 0))
        return FCX_ERR_TIMESTEP;
    if (n == 0)
        return FCX_OK;
    if (!D0 || !I2 || !grad || !stress || !ev || !et)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (constraint) {
    case FCX_UNIAXIAL_STRAIN:
    case FCX_UNIAXIAL_STRESS:
        return kelvin_dispatch<1, 1>(D0, I2, mu0, lam0, mu1, tau, del_t, n, grad, stress, tangent, ev, et, st);
    case FCX_PLANE_STRAIN:
    case FCX_PLANE_STRESS:
        return kelvin_dispatch<4, 2>(D0, I2, mu0, lam0, mu1, tau, del_t, n, grad, stress, tangent, ev, et, st);
    case FCX_FULL:
        return kelvin_dispatch<6, 3>(D0, I2, mu0, lam0, mu1, tau, del_t, n, grad, stress, tangent, ev, et, st);
    default: return FCX_ERR_CONSTRAINT;
    }
}

int fcx_maxwell_evaluate(int constraint, const double *D0, const double *D1, double mu1,
                         double tau, double del_t, size_t n, const double *grad, double *stress,
                         double *tangent, double *ev, double *et, void *stream)
{
    if (!(del_t > 0))
        return FCX_ERR_TIMESTEP;
    if (n == 0)
        return FCX_OK;
    if (!D0 || !D1 || !grad || !stress || !ev || !et)  // tangent == NULL: stress-only evaluate
        return FCX_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (constraint) {
    case FCX_UNIAXIAL_STRAIN:
    case FCX_UNIAXIAL_STRESS:
        return maxwell_dispatch<1, 1>(D0, D1, mu1, tau, del_t, n, grad, stress, tangent, ev, et, st);
    case FCX_PLANE_STRAIN:
    case FCX_PLANE_STRESS:
        return maxwell_dispatch<4, 2>(D0, D1, mu1, tau, del_t, n, grad, stress, tangent, ev, et, st);
    case FCX_FULL:
        return maxwell_dispatch<6, 3>(D0, D1, mu1, tau, del_t, n, grad, stress, tangent, ev, et, st);
    default: return FCX_ERR_CONSTRAINT;
    }
}

int fcx_embed_3d(int constraint, size_t n, const double *grad, const double *stress, double *grad3d,
                 double *stress3d, void *stream)
{
    if (constraint != FCX_UNIAXIAL_STRAIN && constraint != FCX_PLANE_STRAIN)
        return FCX_ERR_CONSTRAINT;
    if (n == 0)
        return FCX_OK;
    if (!grad || !stress || !grad3d || !stress3d)
        return FCX_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long grid = (n + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (grid > cap)
        grid = cap;
    if (constraint == FCX_UNIAXIAL_STRAIN)
        embed3d_kernel<1, 1><<<(unsigned)grid, 256, 0, st>>>(grad, stress, grad3d, stress3d, n);
    else
        embed3d_kernel<2, 4><<<(unsigned)grid, 256, 0, st>>>(grad, stress, grad3d, stress3d, n);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "embed3d_kernel launch");
}

int fcx_extract_from_3d(int constraint, size_t n, const double *stress3d, const double *tangent3d,
                        double *stress, double *tangent, void *stream)
{
    if (constraint != FCX_UNIAXIAL_STRAIN && constraint != FCX_PLANE_STRAIN)
        return FCX_ERR_CONSTRAINT;
    if (n == 0)
        return FCX_OK;
    if (!stress3d || !tangent3d || !stress || !tangent)
        return FCX_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long grid = (n + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (grid > cap)
        grid = cap;
    if (constraint == FCX_UNIAXIAL_STRAIN)
        extract3d_kernel<1, 1><<<(unsigned)grid, 256, 0, st>>>(stress3d, tangent3d, stress, tangent, n);
    else
        extract3d_kernel<2, 4><<<(unsigned)grid, 256, 0, st>>>(stress3d, tangent3d, stress, tangent, n);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "extract3d_kernel launch");
}

int fcx_strain_from_grad_u(int constraint, size_t n, const double *grad, double *strain,
                           void *stream)
{
    if (n == 0)
        return FCX_OK;
    if (!grad || !strain)
        return FCX_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long grid = (n + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (grid > cap)
        grid = cap;
    switch (constraint) {
    case FCX_UNIAXIAL_STRAIN:
    case FCX_UNIAXIAL_STRESS:
        strain_kernel<1, 1><<<(unsigned)grid, 256, 0, st>>>(grad, strain, n);
        break;
    case FCX_PLANE_STRAIN:
    case FCX_PLANE_STRESS:
        strain_kernel<4, 2><<<(unsigned)grid, 256, 0, st>>>(grad, strain, n);
        break;
    case FCX_FULL:
        strain_kernel<6, 3><<<(unsigned)grid, 256, 0, st>>>(grad, strain, n);
        break;
    default: return FCX_ERR_CONSTRAINT;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "strain_kernel launch");
}

}  // extern "C"
