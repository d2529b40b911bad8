// fcx_pcg.cu -- fused vector kernels of the Jacobi-preconditioned conjugate-gradient
// iteration used by the stand-in NewtonSolver (solver/_newton.py; the reference
// delegates the linear solve to PETSc through dolfinx's NewtonSolver, third-party,
// SURVEY.md 3.2).  One Krylov iteration = Jacobian action (fcx_tangent_apply[_rec] +
// fcx_gather_sum) + the three kernels below: 13 vector passes instead of the 21
// passes and ~10 tiny launches of the same update written with torch ops.
//
// All scalars stay on the device.  Reductions are DETERMINISTIC: every CTA writes
// its partial sum to `partials`, the last CTA to finish (atomic ticket) adds them
// in a fixed order -- no floating-point atomics, same result run to run (cf. the
// reference's partition-independence test, tests/solver/test_solver_mpi.py:93-121).
//
// `minv` is the inverse Jacobi diagonal with ZERO on constrained (Dirichlet) dofs;
// it doubles as the free-dof mask.
#include <cuda_runtime.h>

#include <cstdint>

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

// The last CTA to finish adds partials[k * gridDim.x + b] over b, for k < NOUT: thread t takes b = t,
// t + 256, ... in that order, then the fixed tree of block_sum -- the same order every run.  (One thread
// per output walking all partials, one dependent L2 load after the other, cost tens of microseconds per
// launch with 8 x SMs CTAs.)
template <int NOUT>
__device__ __forceinline__ void finish_reduction(const double *mine, double *partials, double *out,
                                                 unsigned *ticket, double *sh)
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
    if (!last)
        return;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NOUT; ++k) {
        const volatile double *src = partials + (size_t)k * gridDim.x;
        double acc = 0.0;
        for (unsigned b = threadIdx.x; b < gridDim.x; b += PCG_THREADS)
            acc += src[b];
        const double tot = block_sum(acc, sh);
        if (threadIdx.x == 0)
            out[k] = tot;
    }
    if (threadIdx.x == 0)
        *ticket = 0;  // ready for the next launch on this stream
}

// The reduction kernels read 16 bytes per access, two pairs per trip with independent
// accumulators (the first, one-double-per-trip versions were latency-bound at 1.8-2.3 TB/s,
// profiles/r1w): branch-free, fixed summation order per thread.

// out[0] = sum_i p_i * Ap_i over free dofs
__global__ void __launch_bounds__(PCG_THREADS)
    pcg_pAp_kernel(size_t n, const double *__restrict__ p, const double *__restrict__ Ap,
                   const double *__restrict__ minv, double *partials, double *out, unsigned *ticket)
{
    __shared__ double sh[PCG_THREADS / 32];
    const size_t n2 = n / 2;
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    const double2 *a2 = reinterpret_cast<const double2 *>(Ap);
    const double2 *m2 = reinterpret_cast<const double2 *>(minv);
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    const size_t stride = (size_t)gridDim.x * PCG_THREADS;
    size_t i = (size_t)blockIdx.x * PCG_THREADS + threadIdx.x;
    for (; i + stride < n2; i += 2 * stride) {
        const double2 pa = p2[i], aa = a2[i], ma = m2[i];
        const double2 pb = p2[i + stride], ab = a2[i + stride], mb = m2[i + stride];
        acc0 = fma(ma.x != 0.0 ? pa.x : 0.0, aa.x, acc0);
        acc1 = fma(ma.y != 0.0 ? pa.y : 0.0, aa.y, acc1);
        acc2 = fma(mb.x != 0.0 ? pb.x : 0.0, ab.x, acc2);
        acc3 = fma(mb.y != 0.0 ? pb.y : 0.0, ab.y, acc3);
    }
    for (; i < n2; i += stride) {
        const double2 pa = p2[i], aa = a2[i], ma = m2[i];
        acc0 = fma(ma.x != 0.0 ? pa.x : 0.0, aa.x, acc0);
        acc1 = fma(ma.y != 0.0 ? pa.y : 0.0, aa.y, acc1);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
        acc0 = fma(minv[n - 1] != 0.0 ? p[n - 1] : 0.0, Ap[n - 1], acc0);
    double mine[1] = {block_sum((acc0 + acc1) + (acc2 + acc3), sh)};
    finish_reduction<1>(mine, partials, out, ticket, sh);
}

// alpha = rz / pAp (0 if pAp <= 0);  x += alpha p;  r -= alpha Ap  (free dofs; r is and stays 0
// on constrained dofs: alpha is masked there);
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
    const size_t n2 = n / 2;
    double2 *x2 = reinterpret_cast<double2 *>(x);
    double2 *r2 = reinterpret_cast<double2 *>(r);
    const double2 *p2 = reinterpret_cast<const double2 *>(p);
    const double2 *a2 = reinterpret_cast<const double2 *>(Ap);
    const double2 *m2 = reinterpret_cast<const double2 *>(minv);
    double s0 = 0.0, s1 = 0.0, q0 = 0.0, q1 = 0.0;
    auto one = [&](double m, double pv, double av, double &xv, double &rv, double &s, double &q) {
        const double al = m != 0.0 ? alpha : 0.0;
        xv = fma(al, pv, xv);
        rv = fma(-al, av, rv);
        s = fma(rv * m, rv, s);
        q = fma(rv, rv, q);
    };
    const size_t stride = (size_t)gridDim.x * PCG_THREADS;
    size_t i = (size_t)blockIdx.x * PCG_THREADS + threadIdx.x;
    for (; i + stride < n2; i += 2 * stride) {
        const double2 pa = p2[i], aa = a2[i], ma = m2[i];
        const double2 pb = p2[i + stride], ab = a2[i + stride], mb = m2[i + stride];
        double2 xa = x2[i], ra = r2[i], xb = x2[i + stride], rb = r2[i + stride];
        one(ma.x, pa.x, aa.x, xa.x, ra.x, s0, q0);
        one(ma.y, pa.y, aa.y, xa.y, ra.y, s1, q1);
        one(mb.x, pb.x, ab.x, xb.x, rb.x, s0, q0);
        one(mb.y, pb.y, ab.y, xb.y, rb.y, s1, q1);
        x2[i] = xa;
        r2[i] = ra;
        x2[i + stride] = xb;
        r2[i + stride] = rb;
    }
    for (; i < n2; i += stride) {
        const double2 pa = p2[i], aa = a2[i], ma = m2[i];
        double2 xa = x2[i], ra = r2[i];
        one(ma.x, pa.x, aa.x, xa.x, ra.x, s0, q0);
        one(ma.y, pa.y, aa.y, xa.y, ra.y, s1, q1);
        x2[i] = xa;
        r2[i] = ra;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0)
        one(minv[n - 1], p[n - 1], Ap[n - 1], x[n - 1], r[n - 1], s0, q0);
    double mine[2];
    mine[0] = block_sum(s0 + s1, sh);
    mine[1] = block_sum(q0 + q1, sh);
    finish_reduction<2>(mine, partials, out, ticket, sh);
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
    if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(Ap) | reinterpret_cast<uintptr_t>(minv)) & 15u)
        return FCX_ERR_ARG;  // vectors must be 16-byte aligned
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
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(p) |
         reinterpret_cast<uintptr_t>(Ap) | reinterpret_cast<uintptr_t>(minv)) & 15u)
        return FCX_ERR_ARG;  // vectors must be 16-byte aligned
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
