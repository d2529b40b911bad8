// fcx_assemble.cu -- device kernels for the residual and the Jacobian action of
// IncrSmallStrainProblem, so that the 288 B/QP tangent
// never leaves HBM during a Newton solve.
//
// Reference semantics (src/fenics_constitutive/solver/_solver.py:87-101):
//     R(v)      = int  eps(v) . sigma                 dx      (R_form)
//     dR(du, v) = int  eps(du) . (C eps(v))           dx      (dR_form)
// with eps() the Mandel strain of ufl.nabla_grad (solver/utils.py:10-62, the
// symbolic twin of models/utils.py:132-208) and the quadrature rule of degree
// q_degree.  For the nodal basis function phi_a e_j the Mandel strain is the
// column b_aj(q) = mandel_strain(grad) with grad[i][j'] = dphi_a/dx_i delta_jj',
// so per cell
//     fe[a][j]      = sum_q w_q |detJ|  b_aj(q) . sigma_q
//     (K p)e[a][j]  = sum_q w_q |detJ|  b_aj(q) . (C_q^T eps_q(p))
//     diag(K)e[a][j]= sum_q w_q |detJ|  b_aj(q) . (C_q^T b_aj(q))
// One thread owns one affine simplex cell and writes its element vector
// [ND][G]; the global vector is then formed by a deterministic gather-sum over
// the node -> (cell, local index) adjacency (no atomics: results do not depend
// on scheduling, cf. reference tests/solver/test_solver_mpi.py:93-121).
#include <cuda_runtime.h>

#include "../../include/fcx.h"
#include "fcx_internal.h"
#include "fcx_models.cuh"

namespace fcx {

constexpr int ASM_THREADS = 128;

// physical basis gradients of local function a at QP q:  gphi[i] = sum_k Jinv[k][i] * dphi_ref[q][a][k]
template <int G>
__device__ __forceinline__ void phys_grad(const double *K, const double *dref, double *gphi)
{
#pragma unroll
    for (int i = 0; i < G; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < G; ++k)
            acc += K[k * G + i] * dref[k];
        gphi[i] = acc;
    }
}

// b_aj . t  for all j:  out[j] = sum_k mandel_strain(grad = gphi (x) e_j)[k] * t[k]
template <int S, int G>
__device__ __forceinline__ void bt_dot(const double *gphi, const double *t, double *out)
{
#pragma unroll
    for (int j = 0; j < G; ++j) {
        double g[G * G], e[S];
#pragma unroll
        for (int i = 0; i < G * G; ++i)
            g[i] = 0.0;
#pragma unroll
        for (int i = 0; i < G; ++i)
            g[i * G + j] = gphi[i];
        mandel_strain<S, G>(g, e);
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < S; ++k)
            acc += e[k] * t[k];
        out[j] = acc;
    }
}

// MODE 0: internal force from a QP vector field (sigma)           -> fe
// MODE 1: Jacobian action: p gathered, tau = C^T eps(p)           -> fe
// MODE 2: Jacobian diagonal                                        -> fe
template <int G, int S, int ND, int NQ, int MODE>
__global__ void __launch_bounds__(ASM_THREADS)
    cell_kernel(const int *__restrict__ dofmap, const double *__restrict__ p,
                const double *__restrict__ dphi_ref, const double *__restrict__ weights,
                const double *__restrict__ Jinv, const double *__restrict__ detJ,
                const double *__restrict__ qvec, const double *__restrict__ tangent,
                double *__restrict__ fe, const unsigned long long ncells)
{
    __shared__ double tab[NQ * ND * G];
    __shared__ double wq[NQ];
    for (int i = threadIdx.x; i < NQ * ND * G; i += ASM_THREADS)
        tab[i] = dphi_ref[i];
    if (threadIdx.x < NQ)
        wq[threadIdx.x] = weights[threadIdx.x];
    __syncthreads();
    const unsigned long long stride = (unsigned long long)gridDim.x * ASM_THREADS;
    for (unsigned long long c = (unsigned long long)blockIdx.x * ASM_THREADS + threadIdx.x;
         c < ncells; c += stride) {
        double K[G * G];
#pragma unroll
        for (int i = 0; i < G * G; ++i)
            K[i] = Jinv[c * (G * G) + i];
        const double dJ = detJ[c];
        double acc[ND][G];
#pragma unroll
        for (int a = 0; a < ND; ++a)
#pragma unroll
            for (int j = 0; j < G; ++j)
                acc[a][j] = 0.0;
        double pe[MODE == 1 ? ND : 1][G];
        if (MODE == 1) {
#pragma unroll
            for (int a = 0; a < ND; ++a) {
                const size_t node = (size_t)dofmap[c * ND + a];
#pragma unroll
                for (int j = 0; j < G; ++j)
                    pe[a][j] = p[node * G + j];
            }
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const double wdet = wq[q] * dJ;
            const unsigned long long qp = c * NQ + q;
            double t[S];
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < S; ++k)
                    t[k] = qvec[qp * S + k];
            } else if (MODE == 1) {
                // grad p at the QP (nabla_grad), its Mandel strain, tau = C^T eps
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
                            T[k][j] += d * pe[a][j];
                    }
                double g[G * G], e[S];
#pragma unroll
                for (int i = 0; i < G; ++i)
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k < G; ++k)
                            s += K[k * G + i] * T[k][j];
                        g[i * G + j] = s;
                    }
                mandel_strain<S, G>(g, e);
                const double *C = tangent + qp * (S * S);
#pragma unroll
                for (int k = 0; k < S; ++k)
                    t[k] = 0.0;
#pragma unroll
                for (int m = 0; m < S; ++m)
#pragma unroll
                    for (int k = 0; k < S; ++k)
                        t[k] += __ldg(C + m * S + k) * e[m];
            }
            double Cq[MODE == 2 ? S * S : 1];
            if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < S * S; ++i)
                    Cq[i] = __ldg(tangent + qp * (S * S) + i);
            }
#pragma unroll
            for (int a = 0; a < ND; ++a) {
                double gphi[G], out[G];
                phys_grad<G>(K, &tab[(q * ND + a) * G], gphi);
                if (MODE == 2) {
                    // b_aj . (C^T b_aj) for each j
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        double g[G * G], e[S];
#pragma unroll
                        for (int i = 0; i < G * G; ++i)
                            g[i] = 0.0;
#pragma unroll
                        for (int i = 0; i < G; ++i)
                            g[i * G + j] = gphi[i];
                        mandel_strain<S, G>(g, e);
                        double s = 0.0;
#pragma unroll
                        for (int m = 0; m < S; ++m)
#pragma unroll
                            for (int k = 0; k < S; ++k)
                                s += e[k] * (Cq[m * S + k] * e[m]);
                        out[j] = s;
                    }
                } else {
                    bt_dot<S, G>(gphi, t, out);
                }
#pragma unroll
                for (int j = 0; j < G; ++j)
                    acc[a][j] += wdet * out[j];
            }
        }
        double *dst = fe + c * (ND * G);
#pragma unroll
        for (int a = 0; a < ND; ++a)
#pragma unroll
            for (int j = 0; j < G; ++j)
                dst[a * G + j] = acc[a][j];
    }
}

// out[node][j] = beta * out[node][j] + alpha * sum_{e in adj(node)} fe[adj_idx[e]][j]
// adj_ptr [nnodes+1], adj_idx entries are (cell * ND + local index); fixed order.
template <int G>
__global__ void __launch_bounds__(256)
    gather_sum_kernel(const long long *__restrict__ adj_ptr, const int *__restrict__ adj_idx,
                      const double *__restrict__ fe, double *__restrict__ out, double alpha,
                      double beta, const unsigned long long nnodes)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
         v < nnodes; v += stride) {
        double acc[G];
#pragma unroll
        for (int j = 0; j < G; ++j)
            acc[j] = 0.0;
        const long long e0 = adj_ptr[v], e1 = adj_ptr[v + 1];
        for (long long e = e0; e < e1; ++e) {
            const double *src = fe + (size_t)adj_idx[e] * G;
#pragma unroll
            for (int j = 0; j < G; ++j)
                acc[j] += src[j];
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const double old = (beta == 0.0) ? 0.0 : beta * out[v * G + j];
            out[v * G + j] = old + alpha * acc[j];
        }
    }
}

template <int G, int S, int ND, int NQ>
static int launch_cell(int mode, size_t ncells, const int *dofmap, const double *p,
                       const double *dphi, const double *w, const double *Jinv,
                       const double *detJ, const double *qvec, const double *tangent, double *fe,
                       cudaStream_t st)
{
    unsigned long long grid = (ncells + ASM_THREADS - 1) / ASM_THREADS;
    const unsigned long long cap = (unsigned long long)sm_count() * 16;
    if (grid > cap)
        grid = cap;
    const unsigned long long nc = ncells;
    switch (mode) {
    case 0:
        cell_kernel<G, S, ND, NQ, 0><<<(unsigned)grid, ASM_THREADS, 0, st>>>(
            dofmap, p, dphi, w, Jinv, detJ, qvec, tangent, fe, nc);
        break;
    case 1:
        cell_kernel<G, S, ND, NQ, 1><<<(unsigned)grid, ASM_THREADS, 0, st>>>(
            dofmap, p, dphi, w, Jinv, detJ, qvec, tangent, fe, nc);
        break;
    default:
        cell_kernel<G, S, ND, NQ, 2><<<(unsigned)grid, ASM_THREADS, 0, st>>>(
            dofmap, p, dphi, w, Jinv, detJ, qvec, tangent, fe, nc);
        break;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "cell_kernel launch");
}

static int dispatch_cell(int mode, int gdim, int sdim, size_t ncells, int nq, int nd,
                         const int *dofmap, const double *p, const double *dphi, const double *w,
                         const double *Jinv, const double *detJ, const double *qvec,
                         const double *tangent, double *fe, cudaStream_t st)
{
#define FCX_CELL_CASE(G, S, ND, NQ) \
    if (gdim == G && sdim == S && nd == ND && nq == NQ) \
        return launch_cell<G, S, ND, NQ>(mode, ncells, dofmap, p, dphi, w, Jinv, detJ, qvec, tangent, fe, st);
    FCX_CELL_CASE(3, 6, 10, 4)  // P2 tetrahedron, q_degree 2
    FCX_CELL_CASE(3, 6, 4, 1)   // P1 tetrahedron, q_degree 1
    FCX_CELL_CASE(3, 6, 4, 4)   // P1 tetrahedron, q_degree 2
    FCX_CELL_CASE(2, 4, 6, 3)   // P2 triangle, q_degree 2
    FCX_CELL_CASE(2, 4, 3, 1)   // P1 triangle, q_degree 1
    FCX_CELL_CASE(2, 4, 3, 3)   // P1 triangle, q_degree 2
    FCX_CELL_CASE(1, 1, 3, 2)   // P2 interval, q_degree 2
    FCX_CELL_CASE(1, 1, 2, 1)   // P1 interval, q_degree 1
    FCX_CELL_CASE(1, 1, 2, 2)   // P1 interval, q_degree 2
#undef FCX_CELL_CASE
    return FCX_ERR_ARG;
}

}  // namespace fcx

using namespace fcx;

extern "C" {

int fcx_internal_force(int gdim, int sdim, size_t ncells, int nq, int nd, const double *dphi_ref,
                       const double *weights, const double *Jinv, const double *detJ,
                       const double *stress, double *fe, void *stream)
{
    if (ncells == 0)
        return FCX_OK;
    if (!dphi_ref || !weights || !Jinv || !detJ || !stress || !fe)
        return FCX_ERR_NULL;
    return dispatch_cell(0, gdim, sdim, ncells, nq, nd, nullptr, nullptr, dphi_ref, weights, Jinv,
                         detJ, stress, nullptr, fe, static_cast<cudaStream_t>(stream));
}

int fcx_tangent_apply(int gdim, int sdim, size_t ncells, int nq, int nd, const int *dofmap,
                      const double *p, const double *dphi_ref, const double *weights,
                      const double *Jinv, const double *detJ, const double *tangent, double *fe,
                      void *stream)
{
    if (ncells == 0)
        return FCX_OK;
    if (!dofmap || !p || !dphi_ref || !weights || !Jinv || !detJ || !tangent || !fe)
        return FCX_ERR_NULL;
    return dispatch_cell(1, gdim, sdim, ncells, nq, nd, dofmap, p, dphi_ref, weights, Jinv, detJ,
                         nullptr, tangent, fe, static_cast<cudaStream_t>(stream));
}

int fcx_tangent_diag(int gdim, int sdim, size_t ncells, int nq, int nd, const double *dphi_ref,
                     const double *weights, const double *Jinv, const double *detJ,
                     const double *tangent, double *fe, void *stream)
{
    if (ncells == 0)
        return FCX_OK;
    if (!dphi_ref || !weights || !Jinv || !detJ || !tangent || !fe)
        return FCX_ERR_NULL;
    return dispatch_cell(2, gdim, sdim, ncells, nq, nd, nullptr, nullptr, dphi_ref, weights, Jinv,
                         detJ, nullptr, tangent, fe, static_cast<cudaStream_t>(stream));
}

int fcx_gather_sum(int gdim, size_t nnodes, const long long *adj_ptr, const int *adj_idx,
                   const double *fe, double *out, double alpha, double beta, void *stream)
{
    if (gdim < 1 || gdim > 3)
        return FCX_ERR_ARG;
    if (nnodes == 0)
        return FCX_OK;
    if (!adj_ptr || !adj_idx || !fe || !out)
        return FCX_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long grid = (nnodes + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (grid > cap)
        grid = cap;
    switch (gdim) {
    case 1: gather_sum_kernel<1><<<(unsigned)grid, 256, 0, st>>>(adj_ptr, adj_idx, fe, out, alpha, beta, nnodes); break;
    case 2: gather_sum_kernel<2><<<(unsigned)grid, 256, 0, st>>>(adj_ptr, adj_idx, fe, out, alpha, beta, nnodes); break;
    default: gather_sum_kernel<3><<<(unsigned)grid, 256, 0, st>>>(adj_ptr, adj_idx, fe, out, alpha, beta, nnodes); break;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "gather_sum_kernel launch");
}

}  // extern "C"
