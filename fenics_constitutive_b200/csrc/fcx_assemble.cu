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
// Each cell's element vector goes to fe [ncells][ND][FS] (FS = 4 for G = 3: one
// 32-byte sector per (cell, node) slot, else G); the global vector is then
// formed by a deterministic gather-sum over the node -> (cell, local index)
// adjacency (no atomics: results do not depend on scheduling, cf. reference
// tests/solver/test_solver_mpi.py:93-121).
//
// Two implementations of the element kernels (fcx_tune "fem_variant"):
//   1 (default) qp_cell_kernel: one thread per QUADRATURE POINT of a tile of
//     whole cells; the tile's tangent (or stress) block, Jinv, detJ and dofmap
//     rows are contiguous ranges and arrive by 1-D bulk async copies (TMA) one
//     tile ahead; the NQ per-QP contributions of a cell are summed through
//     shared memory in a fixed order and leave as one coalesced stream.
//   0 cell_kernel: one thread per cell, plain strided loads (the first version;
//     kept for A/B measurements -- L1-bound at 12 % occupancy, profiles/r1k).
#include <cuda_runtime.h>

#include "../../include/fcx.h"
#include "fcx_fem.cuh"
#include "fcx_internal.h"
#include "fcx_ptx.cuh"

namespace fcx {

constexpr int ASM_THREADS = 128;

// Mandel strain of grad = gphi (x) e_j dotted with t, all j (generic form used by
// the per-cell kernel; the QP-parallel kernel uses bt_dot of fcx_fem.cuh).
template <int S, int G>
__device__ __forceinline__ void bt_dot_dense(const double *gphi, const double *t, double *out)
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
// MODE 3 (qp_cell_kernel only): Jacobian action with VonMises3D tangent RECORDS
//         [nqp][10] = {c0, c1, c2, c3, xn[6]}: tau = C eps(p) from 80 B instead of 288 B per QP
template <int G, int S, int ND, int NQ, int MODE>
__global__ void __launch_bounds__(ASM_THREADS)
    cell_kernel(const int *__restrict__ dofmap, const double *__restrict__ p,
                const double *__restrict__ dphi_ref, const double *__restrict__ weights,
                const double *__restrict__ Jinv, const double *__restrict__ detJ,
                const double *__restrict__ qvec, const double *__restrict__ tangent,
                double *__restrict__ fe, const unsigned long long ncells, const double *__restrict__ gate)
{
    if (gate != nullptr && *gate != 0.0)
        return;  // frozen Krylov loop (fcx_internal.h: fem_set_launch_gate)
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
                    bt_dot_dense<S, G>(gphi, t, out);
                }
#pragma unroll
                for (int j = 0; j < G; ++j)
                    acc[a][j] += wdet * out[j];
            }
        }
        constexpr int FS = FeStride<G>::v;
        double *dst = fe + c * (ND * FS);
#pragma unroll
        for (int a = 0; a < ND; ++a)
#pragma unroll
            for (int j = 0; j < FS; ++j)
                dst[a * FS + j] = (j < G) ? acc[a][j < G ? j : 0] : 0.0;
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
            // adj_idx == nullptr: fe is already in node-major slot order (contiguous per node)
            const double *src = fe + (size_t)(adj_idx != nullptr ? adj_idx[e] : e) * FeStride<G>::v;
            if (G == 3) {  // one 32-byte sector: 16-byte + 8-byte load
                const double2 xy = *reinterpret_cast<const double2 *>(src);
                acc[0] += xy.x;
                acc[1 % G] += xy.y;
                acc[2 % G] += src[2];
            } else {
#pragma unroll
                for (int j = 0; j < G; ++j)
                    acc[j] += src[j];
            }
        }
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const double old = (beta == 0.0) ? 0.0 : beta * out[v * G + j];
            out[v * G + j] = old + alpha * acc[j];
        }
    }
}


// Launch gate of the element kernels (fcx_internal.h), per host thread.
static thread_local const double *t_gate = nullptr;
void fem_set_launch_gate(const double *gate) { t_gate = gate; }
static thread_local unsigned long long *t_tickets = nullptr;
static thread_local unsigned *t_ticket_launches = nullptr;
void fem_set_launch_tickets(unsigned long long *pair, unsigned *launches)
{
    t_tickets = launches != nullptr ? pair : nullptr;
    t_ticket_launches = launches;
}

// ---------------------------------------------------------------------------
// QP-parallel element kernel (fem_variant 1)
// ---------------------------------------------------------------------------
struct CellArgs {
    const int *dofmap;       // [ncells][ND]            (MODE 1)
    const double *p;         // [nnodes][G]             (MODE 1)
    const double *dphi_ref;  // [NQ][ND][G]
    const double *weights;   // [NQ]
    const double *Jinv;      // [ncells][G][G]
    const double *detJ;      // [ncells]
    const double *qarr;      // MODE 0: stress [nqp][S]; MODE 1/2: tangent [nqp][S][S]; MODE 3: records [nqp][10]
    double *fe;              // [ncells][ND][FS], or node-major slots [pos[c][a]][FS] if pos != nullptr
    const int *pos;          // [ncells][ND] slot of (cell, local node) in the node-sorted adjacency, or nullptr
    unsigned long long ncells;
    unsigned long long *ticket;
    int bulk_ok;             // Jinv, detJ, dofmap, pos, qarr 16-byte aligned
    const double *gate;      // nullptr, or: return at once if *gate != 0 (fem_set_launch_gate)
    unsigned long long *ticket_reset;  // nullptr, or the NEXT launch's ticket counter, zeroed by this one
};

template <int G, int S, int ND, int NQ, int MODE>
struct CellCfg {
    static constexpr int TILE = fem_tile<NQ>();
    static constexpr int CPT = TILE / NQ;                    // cells per tile (a multiple of 16)
    static constexpr int SEGW = (MODE == 0) ? S : (MODE == 3 ? 10 : S * S);  // staged doubles per QP
    static constexpr bool GATHER = (MODE == 1 || MODE == 3);  // needs the nodal values of p
    static constexpr int NDG = ND * G;
    static constexpr int PST = NDG | 1;                      // odd stride: conflict-free partials
    static constexpr int DOF_DBL = (CPT * ND + 1) / 2;       // dofmap rows, in doubles
    // NQ = 4 in 3-D: the four QP threads of a cell are four consecutive lanes; their
    // contributions are reduce-scattered with shuffles and written straight to fe
    // (no partial-sum stage: 21 KB instead of 37 KB of shared memory per CTA).
    static constexpr bool SHFL = (NQ == 4 && G == 3);
    static constexpr size_t smem_doubles = (size_t)TILE * SEGW + CPT * G * G + CPT + 2 * DOF_DBL +
                                           (SHFL ? 0 : (size_t)TILE * PST) + NQ * ND * G + NQ + 2;
    static constexpr size_t smem_bytes = sizeof(double) * smem_doubles;
};

template <int G, int S, int ND, int NQ, int MODE>
__global__ void __launch_bounds__(fem_tile<NQ>())
    qp_cell_kernel(const __grid_constant__ CellArgs A)
{
    using Cfg = CellCfg<G, S, ND, NQ, MODE>;
    constexpr int TILE = Cfg::TILE, CPT = Cfg::CPT, SEGW = Cfg::SEGW, PST = Cfg::PST;
    constexpr int FS = FeStride<G>::v;
    static_assert(CPT % 16 == 0, "tile ranges must be 16-byte multiples");
    extern __shared__ __align__(128) double smem[];
    double *s_stage = smem;                              // [TILE][SEGW]   (bulk, barrier 1)
    double *s_jinv = s_stage + TILE * SEGW;              // [CPT][G*G]     (bulk, barrier 0)
    double *s_det = s_jinv + CPT * G * G;                // [CPT]          (bulk, barrier 0)
    int *s_dof = reinterpret_cast<int *>(s_det + CPT);   // [CPT][ND]      (bulk, barrier 0; MODE 1)
    int *s_pos = s_dof + 2 * Cfg::DOF_DBL;               // [CPT][ND]      (bulk, barrier 0; SHFL + A.pos)
    double *s_part = s_det + CPT + 2 * Cfg::DOF_DBL;     // [TILE][PST]    (not with SHFL)
    double *s_tab = s_part + (Cfg::SHFL ? 0 : TILE * PST);  // [NQ][ND][G]
    double *s_wq = s_tab + NQ * ND * G;                  // [NQ]
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_wq + NQ);  // [2]
    __shared__ unsigned long long s_next[2];  // slot = iteration parity (SHFL: one CTA barrier per tile)

    if (A.ticket_reset != nullptr && blockIdx.x == 0 && threadIdx.x == 0)
        *A.ticket_reset = 0ULL;  // also on the gated path: the counters keep alternating cleanly
    if (A.gate != nullptr && *A.gate != 0.0)
        return;  // frozen Krylov loop: uniform over the grid (the word changes only between launches)
    const int tid = threadIdx.x;
    for (int i = tid; i < NQ * ND * G; i += TILE)
        s_tab[i] = A.dphi_ref[i];
    if (tid < NQ)
        s_wq[tid] = A.weights[tid];
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const unsigned long long ntiles = (A.ncells + CPT - 1) / CPT;
    auto is_bulk = [&](unsigned long long t) { return A.bulk_ok && (t + 1) * CPT <= A.ncells; };
    auto issue = [&](unsigned long long t) {  // thread 0: all ranges of tile t, one tile ahead
        const unsigned long long c0 = t * CPT, q0 = t * TILE;
        const bool scatter = Cfg::SHFL && A.pos != nullptr;
        const uint32_t small_bytes =
            (uint32_t)(sizeof(double) * (CPT * G * G + CPT) +
                       ((Cfg::GATHER ? 1 : 0) + (scatter ? 1 : 0)) * sizeof(int) * CPT * ND);
        mbar_arrive_expect_tx(&bar[0], small_bytes);
        bulk_g2s(s_jinv, A.Jinv + c0 * (G * G), (uint32_t)(sizeof(double) * CPT * G * G), &bar[0]);
        bulk_g2s(s_det, A.detJ + c0, (uint32_t)(sizeof(double) * CPT), &bar[0]);
        if (Cfg::GATHER)
            bulk_g2s(s_dof, A.dofmap + c0 * ND, (uint32_t)(sizeof(int) * CPT * ND), &bar[0]);
        if (scatter)
            bulk_g2s(s_pos, A.pos + c0 * ND, (uint32_t)(sizeof(int) * CPT * ND), &bar[0]);
        mbar_arrive_expect_tx(&bar[1], (uint32_t)(sizeof(double) * TILE * SEGW));
        bulk_g2s(s_stage, A.qarr + q0 * SEGW, (uint32_t)(sizeof(double) * TILE * SEGW), &bar[1]);
    };

    uint32_t parity = 0;
    int it = 0;
    unsigned long long tile = blockIdx.x;
    bool bulk = tile < ntiles && is_bulk(tile);
    if (tid == 0 && bulk)
        issue(tile);
    while (tile < ntiles) {
        const unsigned long long c0 = tile * CPT, q0 = tile * TILE;
        const int ncell = (A.ncells - c0 < (unsigned long long)CPT) ? (int)(A.ncells - c0) : CPT;
        const int cnt = ncell * NQ;
        if (tid == 0)
            s_next[it] = (A.ticket != nullptr) ? gridDim.x + atomicAdd(A.ticket, 1ULL) : tile + gridDim.x;
        if (bulk) {
            mbar_wait(&bar[0], parity);
        } else {  // ragged last tile / unaligned views: cooperative plain loads into the same stage
            for (int i = tid; i < ncell * G * G; i += TILE)
                s_jinv[i] = A.Jinv[c0 * (G * G) + i];
            for (int i = tid; i < ncell; i += TILE)
                s_det[i] = A.detJ[c0 + i];
            if (Cfg::GATHER)
                for (int i = tid; i < ncell * ND; i += TILE)
                    s_dof[i] = A.dofmap[c0 * ND + i];
            if (Cfg::SHFL && A.pos != nullptr)
                for (int i = tid; i < ncell * ND; i += TILE)
                    s_pos[i] = A.pos[c0 * ND + i];
            for (int i = tid; i < cnt * SEGW; i += TILE)
                s_stage[i] = A.qarr[q0 * SEGW + i];
            __syncthreads();
        }

        const bool active = tid < cnt;
        const int lc = tid / NQ, q = tid - lc * NQ;
        double K[G * G], e[S];
        double wdet = 0.0;
#pragma unroll
        for (int i = 0; i < S; ++i)
            e[i] = 0.0;
        if (active || Cfg::SHFL) {
#pragma unroll
            for (int i = 0; i < G * G; ++i)
                K[i] = s_jinv[lc * (G * G) + i];
            wdet = s_wq[q] * s_det[lc];
            if (Cfg::GATHER && active) {  // Mandel strain of nabla_grad(p) at this QP
                double g[G * G];
                grad_of_increment<G, ND>(s_tab + q * ND * G, K, s_dof + lc * ND, A.p, nullptr, g);
                mandel_strain<S, G>(g, e);
            }
        }
        if (bulk) {
            mbar_wait(&bar[1], parity);
            parity ^= 1;
        }
        if (active || Cfg::SHFL) {  // SHFL: idle lanes of a ragged tile take part in the shuffles with zeros
            double row[SEGW];
            const double *src = s_stage + tid * SEGW;
            // (idle lanes read in-bounds stage garbage; a cell's 4 lanes are all idle or all
            //  active, so garbage is only ever exchanged among idle lanes and never stored)
            if (SEGW % 2 == 0) {
#pragma unroll
                for (int i = 0; i < SEGW; i += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(src + i);
                    row[i] = v.x;
                    row[i + 1 < SEGW ? i + 1 : i] = v.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < SEGW; ++i)
                    row[i] = src[i];
            }
            double t[S], ts[S];
            if (MODE == 0) {
#pragma unroll
                for (int k = 0; k < S; ++k)
                    t[k] = row[k];
            } else if (MODE == 1) {  // tau = C^T eps
#pragma unroll
                for (int k = 0; k < S; ++k)
                    t[k] = 0.0;
#pragma unroll
                for (int m = 0; m < S; ++m)
#pragma unroll
                    for (int k = 0; k < S; ++k)
                        t[k] = fma(row[(m * S + k) % SEGW], e[m], t[k]);
            } else if (MODE == 3) {
                // C = c0/c1 on the volumetric block, c2 on the shear diagonal, + c3 xn xn^T (symmetric)
                const double c0 = row[0], c1 = row[1 % SEGW], c2 = row[2 % SEGW], c3 = row[3 % SEGW];
                double ne = 0.0;
#pragma unroll
                for (int k = 0; k < S; ++k)
                    ne = fma(row[(4 + k) % SEGW], e[k], ne);
                const double tre = c1 * ((e[0] + e[1 % S]) + e[2 % S]), cn = c3 * ne, dd = c0 - c1;
#pragma unroll
                for (int k = 0; k < S; ++k)
                    t[k] = fma(cn, row[(4 + k) % SEGW], (k < 3) ? fma(dd, e[k], tre) : c2 * e[k]);
            }
            if (MODE != 2)
                prescale_shear<S, G>(t, ts);
            double *part = s_part + tid * PST;
            double *fe_cell = A.fe + (c0 + lc) * (ND * FS);
#pragma unroll
            for (int a = 0; a < ND; ++a) {
                double gphi[G], out[G];
                phys_grad<G>(K, s_tab + (q * ND + a) * G, gphi);
                if (MODE == 2) {  // b_aj . (C^T b_aj) for each j
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        double g[G * G], b[S];
#pragma unroll
                        for (int i = 0; i < G * G; ++i)
                            g[i] = 0.0;
#pragma unroll
                        for (int i = 0; i < G; ++i)
                            g[i * G + j] = gphi[i];
                        mandel_strain<S, G>(g, b);
                        double acc = 0.0;
#pragma unroll
                        for (int m = 0; m < S; ++m)
#pragma unroll
                            for (int k = 0; k < S; ++k)
                                acc += b[k] * (row[(m * S + k) % SEGW] * b[m]);
                        out[j] = acc;
                    }
                } else {
                    bt_dot<S, G>(gphi, ts, out);
                }
                if (Cfg::SHFL) {
                    // reduce-scatter over the cell's 4 lanes: lane q ends with component j = q
                    // of node a summed over the 4 QPs, (q0 + q2) + (q1 + q3); lane 3 = the pad.
                    const double v0 = wdet * out[0], v1 = wdet * out[1 % G], v2 = wdet * out[2 % G];
                    const bool hi = (q & 2) != 0, od = (q & 1) != 0;
                    const double ka = (hi ? v2 : v0) + __shfl_xor_sync(0xffffffffu, hi ? v0 : v2, 2);
                    const double kb = (hi ? 0.0 : v1) + __shfl_xor_sync(0xffffffffu, hi ? v1 : 0.0, 2);
                    const double r = (od ? kb : ka) + __shfl_xor_sync(0xffffffffu, od ? ka : kb, 1);
                    if (active) {  // 4 lanes = one 32-byte sector
                        if (A.pos != nullptr)
                            A.fe[(size_t)s_pos[lc * ND + a] * FS + q] = r;
                        else
                            fe_cell[a * FS + q] = r;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < G; ++j)
                        part[a * G + j] = wdet * out[j];
                }
            }
        }
        __syncthreads();  // partials complete; every staged range has been consumed
        const unsigned long long next = s_next[it];
        it ^= 1;
        const bool next_bulk = next < ntiles && is_bulk(next);
        if (tid == 0 && next_bulk)
            issue(next);  // overlaps the reduction below and the next tile's gathers

        // fe[c][a][j] = sum_q part[c*NQ + q][a*G + j], fixed order; dense coalesced stream
        double *dst = A.fe + c0 * (ND * FS);
        for (int idx = tid; !Cfg::SHFL && idx < ncell * ND * FS; idx += TILE) {
            const int cell = idx / (ND * FS), r = idx - cell * (ND * FS);
            const int a = r / FS, j = r - a * FS;
            double acc = 0.0;
            if (j < G) {
#pragma unroll
                for (int qq = 0; qq < NQ; ++qq)
                    acc += s_part[(cell * NQ + qq) * PST + a * G + j];
            }
            dst[idx] = acc;
        }
        if (!Cfg::SHFL)
            __syncthreads();  // partials consumed before the next tile overwrites them
        tile = next;
        bulk = next_bulk;
    }
}

template <int G, int S, int ND, int NQ, int MODE>
static int launch_qp_cell(const CellArgs &A0, cudaStream_t st)
{
    using Cfg = CellCfg<G, S, ND, NQ, MODE>;
    auto kern = qp_cell_kernel<G, S, ND, NQ, MODE>;
    static OccCache cache;  // per device
    int occ = 1;
    if (int rc = kernel_occupancy(cache, kern, Cfg::TILE, Cfg::smem_bytes, "occupancy(qp_cell)", &occ))
        return rc;
    CellArgs A = A0;
    const unsigned long long ntiles = (A.ncells + Cfg::CPT - 1) / Cfg::CPT;
    const int per_sm = tuned_ctas_per_sm() > 0 ? tuned_ctas_per_sm() : occ;
    unsigned long long grid = (unsigned long long)sm_count() * per_sm;
    if (grid > ntiles)
        grid = ntiles;
    if (t_tickets != nullptr) {  // alternating pair, no memset (fcx_internal.h)
        A.ticket = t_tickets + (*t_ticket_launches & 1u);
        A.ticket_reset = t_tickets + ((*t_ticket_launches + 1u) & 1u);
        ++*t_ticket_launches;
    } else {
        A.ticket = (ntiles > grid) ? tile_ticket(st) : nullptr;
        A.ticket_reset = nullptr;
    }
    A.gate = t_gate;
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    A.bulk_ok = al16(A.Jinv) && al16(A.detJ) && al16(A.qarr) && (!Cfg::GATHER || al16(A.dofmap)) &&
                (A.pos == nullptr || al16(A.pos));
    if (!Cfg::SHFL && A.pos != nullptr)
        return FCX_ERR_ARG;  // node-major slots are written by the shuffle path only (3-D, 4 QPs)
    kern<<<(unsigned)grid, Cfg::TILE, Cfg::smem_bytes, st>>>(A);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "qp_cell_kernel launch");
}

template <int G, int S, int ND, int NQ>
static int launch_cell(int mode, size_t ncells, const int *dofmap, const double *p,
                       const double *dphi, const double *w, const double *Jinv,
                       const double *detJ, const double *qvec, const double *tangent, double *fe,
                       const int *pos, cudaStream_t st)
{
    if (fem_variant() != 0) {
        CellArgs A{dofmap, p, dphi, w, Jinv, detJ, mode == 0 ? qvec : tangent, fe, pos,
                   (unsigned long long)ncells, nullptr, 0, nullptr, nullptr};
        switch (mode) {
        case 0: return launch_qp_cell<G, S, ND, NQ, 0>(A, st);
        case 1: return launch_qp_cell<G, S, ND, NQ, 1>(A, st);
        case 3:
            if constexpr (G == 3 && S == 6)
                return launch_qp_cell<G, S, ND, NQ, 3>(A, st);
            else
                return FCX_ERR_ARG;
        default: return launch_qp_cell<G, S, ND, NQ, 2>(A, st);
        }
    }
    if (mode == 3 || pos != nullptr)
        return FCX_ERR_ARG;  // records / node-major slots need the QP-parallel kernels
    unsigned long long grid = (ncells + ASM_THREADS - 1) / ASM_THREADS;
    const unsigned long long cap = (unsigned long long)sm_count() * 16;
    if (grid > cap)
        grid = cap;
    const unsigned long long nc = ncells;
    switch (mode) {
    case 0:
        cell_kernel<G, S, ND, NQ, 0><<<(unsigned)grid, ASM_THREADS, 0, st>>>(
            dofmap, p, dphi, w, Jinv, detJ, qvec, tangent, fe, nc, t_gate);
        break;
    case 1:
        cell_kernel<G, S, ND, NQ, 1><<<(unsigned)grid, ASM_THREADS, 0, st>>>(
            dofmap, p, dphi, w, Jinv, detJ, qvec, tangent, fe, nc, t_gate);
        break;
    default:
        cell_kernel<G, S, ND, NQ, 2><<<(unsigned)grid, ASM_THREADS, 0, st>>>(
            dofmap, p, dphi, w, Jinv, detJ, qvec, tangent, fe, nc, t_gate);
        break;
    }
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "cell_kernel launch");
}

static int dispatch_cell(int mode, int gdim, int sdim, size_t ncells, int nq, int nd,
                         const int *dofmap, const double *p, const double *dphi, const double *w,
                         const double *Jinv, const double *detJ, const double *qvec,
                         const double *tangent, double *fe, const int *pos, cudaStream_t st)
{
#define FCX_CELL_CASE(G, S, ND, NQ) \
    if (gdim == G && sdim == S && nd == ND && nq == NQ) \
        return launch_cell<G, S, ND, NQ>(mode, ncells, dofmap, p, dphi, w, Jinv, detJ, qvec, tangent, fe, pos, st);
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

int fcx_fe_stride(int gdim) { return gdim == 3 ? 4 : (gdim == 1 || gdim == 2 ? gdim : FCX_ERR_ARG); }

int fcx_internal_force(int gdim, int sdim, size_t ncells, int nq, int nd, const double *dphi_ref,
                       const double *weights, const double *Jinv, const double *detJ,
                       const double *stress, double *fe, const int *fe_pos, void *stream)
{
    if (ncells == 0)
        return FCX_OK;
    if (!dphi_ref || !weights || !Jinv || !detJ || !stress || !fe)
        return FCX_ERR_NULL;
    return dispatch_cell(0, gdim, sdim, ncells, nq, nd, nullptr, nullptr, dphi_ref, weights, Jinv,
                         detJ, stress, nullptr, fe, fe_pos, static_cast<cudaStream_t>(stream));
}

int fcx_tangent_apply(int gdim, int sdim, size_t ncells, int nq, int nd, const int *dofmap,
                      const double *p, const double *dphi_ref, const double *weights,
                      const double *Jinv, const double *detJ, const double *tangent, double *fe,
                      const int *fe_pos, void *stream)
{
    if (ncells == 0)
        return FCX_OK;
    if (!dofmap || !p || !dphi_ref || !weights || !Jinv || !detJ || !tangent || !fe)
        return FCX_ERR_NULL;
    return dispatch_cell(1, gdim, sdim, ncells, nq, nd, dofmap, p, dphi_ref, weights, Jinv, detJ,
                         nullptr, tangent, fe, fe_pos, static_cast<cudaStream_t>(stream));
}

int fcx_tangent_apply_rec(int gdim, int sdim, size_t ncells, int nq, int nd, const int *dofmap,
                          const double *p, const double *dphi_ref, const double *weights,
                          const double *Jinv, const double *detJ, const double *tangent_rec,
                          double *fe, const int *fe_pos, void *stream)
{
    if (ncells == 0)
        return FCX_OK;
    if (!dofmap || !p || !dphi_ref || !weights || !Jinv || !detJ || !tangent_rec || !fe)
        return FCX_ERR_NULL;
    return dispatch_cell(3, gdim, sdim, ncells, nq, nd, dofmap, p, dphi_ref, weights, Jinv, detJ,
                         nullptr, tangent_rec, fe, fe_pos, static_cast<cudaStream_t>(stream));
}

int fcx_tangent_diag(int gdim, int sdim, size_t ncells, int nq, int nd, const double *dphi_ref,
                     const double *weights, const double *Jinv, const double *detJ,
                     const double *tangent, double *fe, const int *fe_pos, void *stream)
{
    if (ncells == 0)
        return FCX_OK;
    if (!dphi_ref || !weights || !Jinv || !detJ || !tangent || !fe)
        return FCX_ERR_NULL;
    return dispatch_cell(2, gdim, sdim, ncells, nq, nd, nullptr, nullptr, dphi_ref, weights, Jinv,
                         detJ, nullptr, tangent, fe, fe_pos, static_cast<cudaStream_t>(stream));
}

int fcx_gather_sum(int gdim, size_t nnodes, const long long *adj_ptr, const int *adj_idx,
                   const double *fe, double *out, double alpha, double beta, void *stream)
{
    if (gdim < 1 || gdim > 3)
        return FCX_ERR_ARG;
    if (nnodes == 0)
        return FCX_OK;
    if (!adj_ptr || !fe || !out)
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
