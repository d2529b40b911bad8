// fcx_gather.cu -- companion kernel: grad_del_u at the quadrature points from
// cell DOFs and precomputed basis-gradient tables.
//
// Reference semantics: IncrementalDisplacement.evaluate_local_incremental_gradient,
// src/fenics_constitutive/solver/_incrementalunknowns.py:19-27,40-49 -- a dolfinx
// Expression of ufl.nabla_grad(u - u_prev) interpolated at the basix quadrature
// points, i.e. grad[c][q][i][j] = d(u - u_prev)_j / dx_i, cell-major QP order.
//
// For an affine cell with reference gradients dphi_ref[q][a][k] = dphi_a/dX_k and
// Jinv[k][i] = dX_k/dx_i:
//     T[k][j]     = sum_a dphi_ref[q][a][k] * du[a][j]
//     grad[i][j]  = sum_k Jinv[k][i] * T[k][j]
// One thread owns one cell: it gathers the cell's nd nodal increments once
// (L2-resident, shared with neighbouring cells), keeps them in registers for
// all nq quadrature points, and stages its nq*g*g results in shared memory so
// the CTA writes the [cells][nq][g][g] block as one dense coalesced stream.
#include <cuda_runtime.h>

#include "../../include/fcx.h"
#include "fcx_internal.h"
#include "fcx_ptx.cuh"

namespace fcx {

constexpr int GATHER_CELLS = 128;  // cells (= threads) per CTA

template <int G, int ND, int NQ>
__global__ void __launch_bounds__(GATHER_CELLS)
    gather_kernel(const int *__restrict__ dofmap, const double *__restrict__ u,
                  const double *__restrict__ u_prev, const double *__restrict__ dphi_ref,
                  const double *__restrict__ Jinv, double *__restrict__ grad,
                  const unsigned long long ncells, const int vec_ok)
{
    constexpr int OUT = NQ * G * G;          // doubles per cell
    constexpr int OUTP = OUT | 1;            // odd stride: conflict-free staging
    extern __shared__ __align__(16) double sm[];
    double *tab = sm;                        // [NQ][ND][G]
    double *stage = sm + ((NQ * ND * G + 1) & ~1);  // [GATHER_CELLS][OUTP]

    const int tid = threadIdx.x;
    for (int i = tid; i < NQ * ND * G; i += GATHER_CELLS)
        tab[i] = dphi_ref[i];
    __syncthreads();

    const unsigned long long ngroups = (ncells + GATHER_CELLS - 1) / GATHER_CELLS;
    for (unsigned long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const unsigned long long c0 = grp * GATHER_CELLS;
        const unsigned long long c = c0 + tid;
        const int cnt = (ncells - c0 < (unsigned long long)GATHER_CELLS) ? (int)(ncells - c0)
                                                                         : GATHER_CELLS;
        if (tid < cnt) {
            double du[ND][G];
#pragma unroll
            for (int a = 0; a < ND; ++a) {
                const size_t node = (size_t)dofmap[c * ND + a];
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    double v = u[node * G + j];
                    if (u_prev != nullptr)
                        v -= u_prev[node * G + j];
                    du[a][j] = v;
                }
            }
            double K[G][G];
#pragma unroll
            for (int k = 0; k < G; ++k)
#pragma unroll
                for (int i = 0; i < G; ++i)
                    K[k][i] = Jinv[c * (G * G) + k * G + i];
            double *out = stage + tid * OUTP;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                double T[G][G];
#pragma unroll
                for (int k = 0; k < G; ++k)
#pragma unroll
                    for (int j = 0; j < G; ++j)
                        T[k][j] = 0.0;
#pragma unroll
                for (int a = 0; a < ND; ++a)
#pragma unroll
                    for (int k = 0; k < G; ++k) {
                        const double d = tab[(q * ND + a) * G + k];
#pragma unroll
                        for (int j = 0; j < G; ++j)
                            T[k][j] += d * du[a][j];
                    }
#pragma unroll
                for (int i = 0; i < G; ++i)
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        double acc = 0.0;
#pragma unroll
                        for (int k = 0; k < G; ++k)
                            acc += K[k][i] * T[k][j];
                        out[q * G * G + i * G + j] = acc;
                    }
            }
        }
        __syncthreads();
        double *dst = grad + c0 * OUT;
        const int total = cnt * OUT;
        if (vec_ok && (OUT % 2 == 0)) {
            for (int p = tid; p < total / 2; p += GATHER_CELLS) {
                const int e = 2 * p;
                const int cell = e / OUT, r = e - cell * OUT;
                st_stream_v2(dst + e, stage[cell * OUTP + r], stage[cell * OUTP + r + 1]);
            }
        } else {
            for (int e = tid; e < total; e += GATHER_CELLS) {
                const int cell = e / OUT, r = e - cell * OUT;
                dst[e] = stage[cell * OUTP + r];
            }
        }
        __syncthreads();
    }
}

// Generic fallback for element/quadrature combinations without a compiled
// specialisation: one thread per (cell, qp), runtime loops.
__global__ void gather_generic_kernel(int G, int nq, int nd, const int *__restrict__ dofmap,
                                      const double *__restrict__ u,
                                      const double *__restrict__ u_prev,
                                      const double *__restrict__ dphi_ref,
                                      const double *__restrict__ Jinv, double *__restrict__ grad,
                                      unsigned long long ncells)
{
    const unsigned long long total = ncells * (unsigned long long)nq;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
         t < total; t += stride) {
        const unsigned long long c = t / nq;
        const int q = (int)(t - c * nq);
        double T[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int a = 0; a < nd; ++a) {
            const size_t node = (size_t)dofmap[c * nd + a];
            for (int j = 0; j < G; ++j) {
                double v = u[node * G + j];
                if (u_prev != nullptr)
                    v -= u_prev[node * G + j];
                for (int k = 0; k < G; ++k)
                    T[k][j] += dphi_ref[((size_t)q * nd + a) * G + k] * v;
            }
        }
        for (int i = 0; i < G; ++i)
            for (int j = 0; j < G; ++j) {
                double acc = 0.0;
                for (int k = 0; k < G; ++k)
                    acc += Jinv[c * G * G + k * G + i] * T[k][j];
                grad[t * G * G + i * G + j] = acc;
            }
    }
}

template <int G, int ND, int NQ>
static int launch_gather(size_t ncells, const int *dofmap, const double *u, const double *u_prev,
                         const double *dphi, const double *Jinv, double *grad, cudaStream_t st)
{
    constexpr int OUT = NQ * G * G;
    constexpr int OUTP = OUT | 1;
    constexpr size_t smem = sizeof(double) * (((NQ * ND * G + 1) & ~1) + GATHER_CELLS * OUTP);
    auto kern = gather_kernel<G, ND, NQ>;
    static int occ = -1;
    if (occ < 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return note_cuda_error(e, "cudaFuncSetAttribute(gather)");
        int o = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, GATHER_CELLS, smem);
        if (e != cudaSuccess)
            return note_cuda_error(e, "cudaOccupancy(gather)");
        occ = o > 0 ? o : 1;
    }
    const unsigned long long ngroups = (ncells + GATHER_CELLS - 1) / GATHER_CELLS;
    unsigned long long grid = (unsigned long long)sm_count() * occ;
    if (grid > ngroups)
        grid = ngroups;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(grad) & 15u) == 0;
    kern<<<(unsigned)grid, GATHER_CELLS, smem, st>>>(dofmap, u, u_prev, dphi, Jinv, grad,
                                                     (unsigned long long)ncells, vec_ok ? 1 : 0);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "gather_kernel launch");
}

}  // namespace fcx

using namespace fcx;

extern "C" int fcx_gather_grad(int gdim, size_t ncells, int nq, int nd, const int *dofmap,
                               const double *u, const double *u_prev, const double *dphi_ref,
                               const double *Jinv, double *grad, void *stream)
{
    if (gdim < 1 || gdim > 3 || nq < 1 || nd < 1)
        return FCX_ERR_ARG;
    if (ncells == 0)
        return FCX_OK;
    if (!dofmap || !u || !dphi_ref || !Jinv || !grad)
        return FCX_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define FCX_GATHER_CASE(G, ND, NQ) \
    if (gdim == G && nd == ND && nq == NQ) \
        return launch_gather<G, ND, NQ>(ncells, dofmap, u, u_prev, dphi_ref, Jinv, grad, st);
    FCX_GATHER_CASE(3, 10, 4)  // P2 tetrahedron, q_degree 2 (BASELINE config 5)
    FCX_GATHER_CASE(3, 4, 1)   // P1 tetrahedron, q_degree 1 (reference tests/models/test_plasticity.py:16-17)
    FCX_GATHER_CASE(3, 10, 1)
    FCX_GATHER_CASE(3, 4, 4)   // P1 tetrahedron, q_degree 2
    FCX_GATHER_CASE(2, 3, 3)   // P1 triangle, q_degree 2
    FCX_GATHER_CASE(1, 2, 2)   // P1 interval, q_degree 2
    FCX_GATHER_CASE(2, 6, 3)   // P2 triangle, q_degree 2
    FCX_GATHER_CASE(2, 3, 1)   // P1 triangle, q_degree 1
    FCX_GATHER_CASE(1, 3, 2)   // P2 interval, q_degree 2/3
    FCX_GATHER_CASE(1, 2, 1)   // P1 interval
#undef FCX_GATHER_CASE
    const unsigned long long total = (unsigned long long)ncells * nq;
    unsigned long long grid = (total + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (grid > cap)
        grid = cap;
    gather_generic_kernel<<<(unsigned)grid, 256, 0, st>>>(gdim, nq, nd, dofmap, u, u_prev, dphi_ref,
                                                          Jinv, grad, (unsigned long long)ncells);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "gather_generic_kernel launch");
}
