// fcx_pcg.cu -- fused vector kernels of the Jacobi-preconditioned conjugate-gradient
// iteration used by the stand-in NewtonSolver (solver/_newton.py; the reference
// delegates the linear solve to PETSc through dolfinx's NewtonSolver, third-party,
// SURVEY.md 3.2).  One Krylov iteration = Jacobian action (fcx_tangent_apply[_rec] +
// fcx_gather_sum) + the three kernels below: 13 vector passes instead of the 21
// passes and ~10 tiny launches of the same update written with torch ops.
//
// All scalars stay on the device.  Reductions are DETERMINISTIC: every CTA writes
// its partial sum to `partials`, the last CTA to finish (atomic ticket) adds them
// in index order -- no floating-point atomics, same result run to run (cf. the
// reference's partition-independence test, tests/solver/test_solver_mpi.py:93-121).
//
// `minv` is the inverse Jacobi diagonal with ZERO on constrained (Dirichlet) dofs;
// it doubles as the free-dof mask.
#include <cuda_runtime.h>

#include "../../include/fcx.h"
#include "fcx_internal.h"

namespace fcx {

constexpr int PCG_THREADS = 256;

__device__ __forceinline__ double block_sum(double v, double *sh)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0)
        sh[wid] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < PCG_THREADS / 32; ++w)
            t += sh[w];
    }
    __syncthreads();
    return t;  // valid in thread 0
}

// Last CTA adds partials[k * gridDim.x + b] over b in order, for k < NOUT.
template <int NOUT>
__device__ __forceinline__ void finish_reduction(const double *mine, double *partials, double *out,
                                                 unsigned *ticket)
{
    __shared__ bool last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NOUT; ++k)
            partials[(size_t)k * gridDim.x + blockIdx.x] = mine[k];
        __threadfence();
        const unsigned t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < NOUT) {
        __threadfence();
        double acc = 0.0;
        const volatile double *src = partials + (size_t)threadIdx.x * gridDim.x;
        for (unsigned b = 0; b < gridDim.x; ++b)
            acc += src[b];
        out[threadIdx.x] = acc;
        if (threadIdx.x == 0)
            *ticket = 0;  // ready for the next launch on this stream
    }
}

// out[0] = sum_i p_i * Ap_i over free dofs
__global__ void __launch_bounds__(PCG_THREADS)
    pcg_pAp_kernel(size_t n, const double *__restrict__ p, const double *__restrict__ Ap,
                   const double *__restrict__ minv, double *partials, double *out, unsigned *ticket)
{
    __shared__ double sh[PCG_THREADS / 32];
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * PCG_THREADS;
    for (size_t i = (size_t)blockIdx.x * PCG_THREADS + threadIdx.x; i < n; i += stride)
        if (minv[i] != 0.0)
            acc = fma(p[i], Ap[i], acc);
    double mine[1] = {block_sum(acc, sh)};
    finish_reduction<1>(mine, partials, out, ticket);
}

// alpha = rz / pAp (0 if pAp <= 0);  x += alpha p;  r = (r - alpha Ap) on free dofs;
// out[0] = sum r.(minv r) (= rz_new),  out[1] = sum r.r
__global__ void __launch_bounds__(PCG_THREADS)
    pcg_update_xr_kernel(size_t n, double *__restrict__ x, double *__restrict__ r,
                         const double *__restrict__ p, const double *__restrict__ Ap,
                         const double *__restrict__ minv, const double *__restrict__ rz,
                         const double *__restrict__ pAp, double *partials, double *out,
                         unsigned *ticket)
{
    __shared__ double sh[PCG_THREADS / 32];
    const double den = *pAp;
    const double alpha = den > 0.0 ? *rz / den : 0.0;
    double a0 = 0.0, a1 = 0.0;
    const size_t stride = (size_t)gridDim.x * PCG_THREADS;
    for (size_t i = (size_t)blockIdx.x * PCG_THREADS + threadIdx.x; i < n; i += stride) {
        const double m = minv[i];
        if (m != 0.0) {
            x[i] = fma(alpha, p[i], x[i]);
            const double ri = fma(-alpha, Ap[i], r[i]);
            r[i] = ri;
            a0 = fma(ri * m, ri, a0);
            a1 = fma(ri, ri, a1);
        }
    }
    double mine[2];
    mine[0] = block_sum(a0, sh);
    mine[1] = block_sum(a1, sh);
    finish_reduction<2>(mine, partials, out, ticket);
}

// p = minv r + (rz_new / rz) p   (beta = 0 if rz <= 0)
__global__ void __launch_bounds__(PCG_THREADS)
    pcg_update_p_kernel(size_t n, double *__restrict__ p, const double *__restrict__ r,
                        const double *__restrict__ minv, const double *__restrict__ rz_new,
                        const double *__restrict__ rz)
{
    const double den = *rz;
    const double beta = den > 0.0 ? *rz_new / den : 0.0;
    const size_t stride = (size_t)gridDim.x * PCG_THREADS;
    for (size_t i = (size_t)blockIdx.x * PCG_THREADS + threadIdx.x; i < n; i += stride)
        p[i] = fma(beta, p[i], minv[i] * r[i]);
}

static unsigned pcg_grid(size_t n)
{
    size_t g = (n + PCG_THREADS - 1) / PCG_THREADS;
    const size_t cap = (size_t)sm_count() * 8;
    return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace fcx

using namespace fcx;

extern "C" {

size_t fcx_pcg_scratch_doubles(void) { return 2 * (size_t)sm_count() * 8 + 8; }

int fcx_pcg_pap(size_t n, const double *p, const double *Ap, const double *minv, double *scratch,
                unsigned *ticket, double *out, void *stream)
{
    if (!p || !Ap || !minv || !scratch || !ticket || !out)
        return FCX_ERR_NULL;
    pcg_pAp_kernel<<<pcg_grid(n), PCG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(n, p, Ap, minv, scratch,
                                                                                      out, ticket);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "pcg_pAp_kernel launch");
}

int fcx_pcg_update_xr(size_t n, double *x, double *r, const double *p, const double *Ap,
                      const double *minv, const double *rz, const double *pAp, double *scratch,
                      unsigned *ticket, double *out2, void *stream)
{
    if (!x || !r || !p || !Ap || !minv || !rz || !pAp || !scratch || !ticket || !out2)
        return FCX_ERR_NULL;
    pcg_update_xr_kernel<<<pcg_grid(n), PCG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        n, x, r, p, Ap, minv, rz, pAp, scratch, out2, ticket);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "pcg_update_xr_kernel launch");
}

int fcx_pcg_update_p(size_t n, double *p, const double *r, const double *minv, const double *rz_new,
                     const double *rz, void *stream)
{
    if (!p || !r || !minv || !rz_new || !rz)
        return FCX_ERR_NULL;
    pcg_update_p_kernel<<<pcg_grid(n), PCG_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(n, p, r, minv,
                                                                                           rz_new, rz);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "pcg_update_p_kernel launch");
}

}  // extern "C"
