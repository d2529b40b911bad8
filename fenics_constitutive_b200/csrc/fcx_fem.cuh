// fcx_fem.cuh -- per-quadrature-point finite-element device functions shared by
// the gather kernel (fcx_gather.cu), the fused form() kernel (fcx_mises_form.cuh)
// and the residual / Jacobian-action kernels (fcx_assemble.cu).
//
// Affine simplex cells:  dphi_ref[q][a][k] = dphi_a/dX_k at quadrature point q,
// Jinv[k][i] = dX_k/dx_i.  All three users form grad_del_u with the SAME
// explicit fused-multiply-add chains below, so the fused form() kernel and the
// separate gather + evaluate launches agree bit for bit (libfcx.so is otherwise
// compiled with -fmad=false to track the reference's numpy arithmetic in the
// pointwise models; these kernels are not part of that contract and their
// fp64 pipe time matters: 117 chained multiply-adds per QP for a P2 tet).
#pragma once
#include "fcx_models.cuh"

namespace fcx {

// Doubles per (cell, local node) slot of an element vector fe [ncells][nd][FS]:
// 3-D slots are padded to 4 doubles = one 32-byte sector, so the node-wise
// gather-sum reads exactly one sector per contribution.
template <int G>
struct FeStride {
    static constexpr int v = (G == 3) ? 4 : G;
};

// QPs per CTA tile of the QP-parallel FEM kernels: whole cells, whole warps.
template <int NQ>
constexpr int fem_tile()
{
    return NQ == 3 ? 96 : 64;
}

// grad[i][j] = d(du)_j/dx_i (ufl.nabla_grad; reference solver/_incrementalunknowns.py:25-27)
//   T[k][j]    = sum_a tabq[a][k] * du(a, j)
//   grad[i][j] = sum_k K[k][i] * T[k][j]
// `du(a, v)` stores the G components of the nodal increment of local node a in v.
template <int G, int ND, class LoadDu>
__device__ __forceinline__ void grad_at_qp(const double *tabq, const double *K, LoadDu &&du,
                                           double *g)
{
    double T[G][G];
#pragma unroll
    for (int k = 0; k < G; ++k)
#pragma unroll
        for (int j = 0; j < G; ++j)
            T[k][j] = 0.0;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
        double v[G];
        du(a, v);
#pragma unroll
        for (int k = 0; k < G; ++k) {
            const double d = tabq[a * G + k];
#pragma unroll
            for (int j = 0; j < G; ++j)
                T[k][j] = fma(d, v[j], T[k][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < G; ++i)
#pragma unroll
        for (int j = 0; j < G; ++j) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < G; ++k)
                acc = fma(K[k * G + i], T[k][j], acc);
            g[i * G + j] = acc;
        }
}

// grad_at_qp for QPT consecutive quadrature points of ONE cell handled by one thread: each nodal
// increment is fetched once and used for all QPT points (per point the same operations in the
// same order as grad_at_qp, hence the same bits).  tab0 = table of the first point; the tables of
// consecutive points are ND*G doubles apart.
template <int G, int ND, int QPT, class LoadDu>
__device__ __forceinline__ void grad_at_qps(const double *tab0, const double *K, LoadDu &&du,
                                            double (*g)[G * G])
{
    double T[QPT][G][G];
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq)
#pragma unroll
        for (int k = 0; k < G; ++k)
#pragma unroll
            for (int j = 0; j < G; ++j)
                T[qq][k][j] = 0.0;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
        double v[G];
        du(a, v);
#pragma unroll
        for (int qq = 0; qq < QPT; ++qq)
#pragma unroll
            for (int k = 0; k < G; ++k) {
                const double d = tab0[(qq * ND + a) * G + k];
#pragma unroll
                for (int j = 0; j < G; ++j)
                    T[qq][k][j] = fma(d, v[j], T[qq][k][j]);
            }
    }
#pragma unroll
    for (int qq = 0; qq < QPT; ++qq)
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
            for (int j = 0; j < G; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < G; ++k)
                    acc = fma(K[k * G + i], T[qq][k][j], acc);
                g[qq][i * G + j] = acc;
            }
}

// One node's G components of a blocked nodal vector (node-major, block size G).
// G = 3: the 24-byte record is 8-byte aligned only, but one of its two halves is
// always 16-byte aligned -- even nodes load (x, y) as a pair and z alone, odd
// nodes x alone and (y, z) as a pair: two load instructions instead of three,
// branch-free.  Measured alternatives to every QP thread of a cell issuing its own
// (L1-hitting) loads were SLOWER on B200: one fetch per (cell, node) staged through
// shared memory (+27 % on the fused form() kernel, profiles/r1m) and a 4-lane
// split of the fetches exchanged with shuffles (+23 %, profiles/r1o).
template <int G>
__device__ __forceinline__ void load_node(const double *__restrict__ base, size_t node, double *v)
{
    if (G == 3) {
        const size_t odd = node & 1;
        const double *p = base + node * 3;
        const double2 pr = __ldg(reinterpret_cast<const double2 *>(p + odd));
        const double sc = __ldg(p + (odd ? 0 : 2));
        v[0] = odd ? sc : pr.x;
        v[1] = odd ? pr.x : pr.y;
        v[2 % G] = odd ? pr.y : sc;
    } else {
#pragma unroll
        for (int j = 0; j < G; ++j)
            v[j] = __ldg(base + node * G + j);
    }
}

// grad_at_qp for the increment u - u_prev (u_prev may be nullptr) of the cell whose
// dofmap row is `dm`.
template <int G, int ND>
__device__ __forceinline__ void grad_of_increment(const double *tabq, const double *K, const int *dm,
                                                  const double *__restrict__ u,
                                                  const double *__restrict__ u_prev, double *g)
{
    if (u_prev != nullptr)
        grad_at_qp<G, ND>(
            tabq, K,
            [&](int a, double *v) {
                double w[G];
                load_node<G>(u, (size_t)dm[a], v);
                load_node<G>(u_prev, (size_t)dm[a], w);
#pragma unroll
                for (int j = 0; j < G; ++j)
                    v[j] -= w[j];
            },
            g);
    else
        grad_at_qp<G, ND>(tabq, K, [&](int a, double *v) { load_node<G>(u, (size_t)dm[a], v); }, g);
}

// grad_at_qps for the increment u - u_prev (see grad_of_increment).
template <int G, int ND, int QPT>
__device__ __forceinline__ void grads_of_increment(const double *tab0, const double *K, const int *dm,
                                                   const double *__restrict__ u,
                                                   const double *__restrict__ u_prev, double (*g)[G * G])
{
    if (u_prev != nullptr)
        grad_at_qps<G, ND, QPT>(
            tab0, K,
            [&](int a, double *v) {
                double w[G];
                load_node<G>(u, (size_t)dm[a], v);
                load_node<G>(u_prev, (size_t)dm[a], w);
#pragma unroll
                for (int j = 0; j < G; ++j)
                    v[j] -= w[j];
            },
            g);
    else
        grad_at_qps<G, ND, QPT>(tab0, K, [&](int a, double *v) { load_node<G>(u, (size_t)dm[a], v); }, g);
}

// physical basis gradient of local function a at a QP:  gphi[i] = sum_k K[k][i] * dref[k]
template <int G>
__device__ __forceinline__ void phys_grad(const double *K, const double *dref, double *gphi)
{
#pragma unroll
    for (int i = 0; i < G; ++i) {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < G; ++k)
            acc = fma(K[k * G + i], dref[k], acc);
        gphi[i] = acc;
    }
}

// out[j] = b_aj . t  where b_aj = mandel_strain(grad = gphi (x) e_j) (the Mandel
// strain of nodal basis function phi_a e_j under nabla_grad, reference
// solver/utils.py:10-62 / models/utils.py:132-208), written out for its
// non-zero entries.  ts = t with the shear entries pre-multiplied by 1/sqrt(2).
template <int S, int G>
__device__ __forceinline__ void bt_dot(const double *gphi, const double *ts, double *out)
{
    if (G == 1) {
        out[0] = gphi[0] * ts[0];
    } else if (G == 2) {  // e = [g00, g11, 0, r (g01 + g10)]
        out[0] = fma(gphi[0], ts[0], gphi[1] * ts[3]);
        out[1] = fma(gphi[1], ts[1], gphi[0] * ts[3]);
    } else {  // e = [g00, g11, g22, r (g01 + g10), r (g02 + g20), r (g12 + g21)]
        out[0] = fma(gphi[0], ts[0], fma(gphi[1], ts[3], gphi[2] * ts[4]));
        out[1] = fma(gphi[1], ts[1], fma(gphi[0], ts[3], gphi[2] * ts[5]));
        out[2] = fma(gphi[2], ts[2], fma(gphi[0], ts[4], gphi[1] * ts[5]));
    }
}

template <int S, int G>
__device__ __forceinline__ void prescale_shear(const double *t, double *ts)
{
    constexpr int NNORMAL = (G == 1) ? 1 : 3;
#pragma unroll
    for (int k = 0; k < S; ++k)
        ts[k] = (k < NNORMAL) ? t[k] : shear_factor() * t[k];
}

}  // namespace fcx
