// fcx_mises_form.cuh -- fused `form()` pipeline for VonMises3D on the device:
// one launch does what the reference does per Newton iteration in
// LawOnSubMesh.evaluate (solver/_lawonsubmesh.py:72-95) for a single law on the
// whole mesh (IdentityMap):
//   1. grad_del_u at the QPs from the cell dofs      (_incrementalunknowns.py:40-49)
//   2. history trial reset  history_1 <- history_0    (_history.py:64-79)
//   3. sigma_local <- stress.previous                 (_lawonsubmesh.py:58-61)
//   4. law.evaluate(...)                              (_lawonsubmesh.py:86-94)
//   5. stress.current / tangent <- local results      (_lawonsubmesh.py:63-70)
// Steps 2, 3 and 5 are full-array copies in the reference; here they cost
// nothing: the kernel READS the committed arrays (stress_prev, eps_n0, alpha0)
// and WRITES the trial arrays (stress_cur, eps_n1, alpha1, tangent), and
// grad_del_u never goes to memory unless the caller asks for it.
//
// A law that owns only part of the mesh (reference LawOnSubMesh with a SubSpaceMap,
// solver/maps.py:62-123, _lawonsubmesh.py:58-70) passes its cell list `cells`: dofmap / Jinv / the
// history arrays are then indexed by the law's LOCAL cell number (they live on the sub-mesh), while
// stress_prev / stress_cur / tangent / tangent_rec are rows cells[c] of the PARENT arrays -- the
// map_to_sub / map_to_parent passes (a read and a write of every stress and tangent byte) are
// folded into the bulk copies, one per cell row instead of one per tile.
//
// Same staging as fcx_mises_ostage.cuh: thread t owns QP t of a tile of TILE
// QPs (= TILE/NQ whole cells); committed state comes in with bulk async copies
// while the threads gather their cell's nodal increments (L1/L2-resident) and
// form grad_del_u in registers; sigma/eps_n/alpha/tangent leave through shared
// memory with bulk async stores.
#pragma once
#include "fcx_fem.cuh"
#include "fcx_models.cuh"

namespace fcx {

template <int ND, int NQ, int TILE>
constexpr size_t mises_form_smem_bytes()
{
    return sizeof(double) * (49 * TILE + ((NQ * ND * 3 + 1) & ~1)) + sizeof(uint64_t);
}

struct MisesFormArgs {
    const int *cells;           // [ncells] parent cell of local cell c, or nullptr (law on every cell)
    const int *dofmap;          // [ncells][ND]
    const double *u;            // [nnodes][3]
    const double *u_prev;       // [nnodes][3] or nullptr
    const double *dphi_ref;     // [NQ][ND][3]
    const double *Jinv;         // [ncells][3][3]   dX_k/dx_i
    const double *stress_prev;  // [n][6]
    const double *eps0;         // [n][6]
    const double *alpha0;       // [n]
    double *stress_cur;         // [n][6]
    double *tangent;            // [n][36] or nullptr (stress-only: residual evaluations, or J from the records)
    double *eps1;               // [n][6]
    double *alpha1;             // [n]
    double *grad_out;           // [n][9] or nullptr
    double *trec;               // [n][10] or nullptr: tangent records (4 coefficients + flow direction)
    unsigned char *flag;        // [n] or nullptr
    int *status;                // int[2] or nullptr
    unsigned long long *ticket;
    unsigned long long ncells;
    int bulk_ok;                // all state pointers 16-byte aligned
};

template <int ND, int NQ, int TILE, int MINCTAS>
__global__ void __launch_bounds__(TILE, MINCTAS)
    fcx_mises_form_kernel(const __grid_constant__ MisesParams P,
                          const __grid_constant__ MisesFormArgs A)
{
    static_assert(TILE % NQ == 0, "a tile holds whole cells");
    constexpr int CPT = TILE / NQ;  // cells per tile
    extern __shared__ __align__(128) double smem[];
    double *s_sig = smem;              // [TILE][6]
    double *s_eps = smem + 6 * TILE;   // [TILE][6]
    double *s_alp = smem + 12 * TILE;  // [TILE]
    double *s_tan = smem + 13 * TILE;  // [TILE][36]
    double *s_tab = smem + 49 * TILE;  // [NQ][ND][3]
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_tab + ((NQ * ND * 3 + 1) & ~1));
    __shared__ unsigned long long s_next;

    const int tid = threadIdx.x;
    for (int i = tid; i < NQ * ND * 3; i += TILE)
        s_tab[i] = A.dphi_ref[i];
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const unsigned long long ntiles = (A.ncells + CPT - 1) / CPT;
    const unsigned long long nqp = A.ncells * NQ;
    uint32_t parity = 0;
    unsigned long long tile = blockIdx.x;
    while (tile < ntiles) {
        const unsigned long long q0 = tile * TILE;
        const int cnt = (nqp - q0 < (unsigned long long)TILE) ? (int)(nqp - q0) : TILE;
        // full tiles move through the TMA engine; the ragged last tile (and
        // unaligned views) use plain per-thread loads/stores
        const bool bulk = A.bulk_ok && cnt == TILE;
        if (tid == 0) {
            s_next = (A.ticket != nullptr) ? gridDim.x + atomicAdd(A.ticket, 1ULL) : tile + gridDim.x;
            if (bulk) {
                bulk_wait_read_all();  // previous tile's stores have left shared memory
                mbar_arrive_expect_tx(bar, (uint32_t)(13 * TILE * sizeof(double)));
                if (A.cells == nullptr) {
                    bulk_g2s(s_sig, A.stress_prev + q0 * 6, TILE * 6 * sizeof(double), bar);
                } else {
                    const int *pc = A.cells + q0 / NQ;
#pragma unroll 4
                    for (int lc = 0; lc < CPT; ++lc)
                        bulk_g2s(s_sig + lc * NQ * 6, A.stress_prev + (size_t)pc[lc] * NQ * 6,
                                 NQ * 6 * sizeof(double), bar);
                }
                bulk_g2s(s_eps, A.eps0 + q0 * 6, TILE * 6 * sizeof(double), bar);
                bulk_g2s(s_alp, A.alpha0 + q0, TILE * sizeof(double), bar);
            }
        }

        // ---- grad_del_u of this thread's QP, in registers (overlaps the loads) ----
        double g[9];
        const bool active = tid < cnt;
        unsigned pcell;  // this thread's cell in the parent arrays
        {
            // idle lanes of the ragged last tile clamp to the last cell; nothing of theirs is stored
            unsigned long long c = q0 / NQ + tid / NQ;
            c = c < A.ncells ? c : A.ncells - 1;
            pcell = A.cells != nullptr ? (unsigned)A.cells[c] : (unsigned)c;
            const int q = tid % NQ;
            double K[9];
#pragma unroll
            for (int i = 0; i < 9; ++i)
                K[i] = A.Jinv[c * 9 + i];
            // same device functions as gather_kernel: fused == gather + evaluate, bit for bit
            if (active)
                grad_of_increment<3, ND>(s_tab + q * ND * 3, K, A.dofmap + c * ND, A.u, A.u_prev, g);
            if (active && A.grad_out != nullptr) {
#pragma unroll
                for (int i = 0; i < 9; ++i)
                    A.grad_out[(q0 + tid) * 9 + i] = g[i];
            }
        }

        if (bulk) {
            mbar_wait(bar, parity);
            parity ^= 1;
        } else {
            __syncthreads();
        }
        const unsigned long long next = s_next;  // rewritten by thread 0 only after the barrier below

        if (active) {
            const unsigned long long qg = q0 + tid;
            const unsigned long long qp_parent = (unsigned long long)pcell * NQ + tid % NQ;
            double sig[6], ep[6], al;
            if (bulk) {
#pragma unroll
                for (int i = 0; i < 6; i += 2) {
                    const double2 a = *reinterpret_cast<const double2 *>(s_sig + tid * 6 + i);
                    const double2 b = *reinterpret_cast<const double2 *>(s_eps + tid * 6 + i);
                    sig[i] = a.x;
                    sig[i + 1] = a.y;
                    ep[i] = b.x;
                    ep[i + 1] = b.y;
                }
                al = s_alp[tid];
            } else {
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    sig[i] = A.stress_prev[qp_parent * 6 + i];
                    ep[i] = A.eps0[qg * 6 + i];
                }
                al = A.alpha0[qg];
            }
            bool plastic = false, failed = false;
            double coef[4], xn[6];
            mises_point(P, g, sig, ep, al, coef, xn, plastic, failed);
            if (A.flag != nullptr)
                A.flag[qg] = plastic ? 1 : 0;
            if (failed && A.status != nullptr) {
                atomicAdd(&A.status[0], 1);
                const unsigned long long q = qg;
                atomicMin(&A.status[1], q > 0x7fffffffULL ? 0x7fffffff : (int)q);
            }
            if (bulk) {
#pragma unroll
                for (int i = 0; i < 6; i += 2) {
                    *reinterpret_cast<double2 *>(s_sig + tid * 6 + i) = make_double2(sig[i], sig[i + 1]);
                    *reinterpret_cast<double2 *>(s_eps + tid * 6 + i) = make_double2(ep[i], ep[i + 1]);
                }
                s_alp[tid] = al;
            } else {
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    A.stress_cur[qp_parent * 6 + i] = sig[i];
                    A.eps1[qg * 6 + i] = ep[i];
                }
                A.alpha1[qg] = al;
            }
            if (A.trec != nullptr) {  // 80 B instead of 288: what the matrix-free Jacobian action reads
                double *r = A.trec + qp_parent * 10;
                *reinterpret_cast<double2 *>(r + 0) = make_double2(coef[0], coef[1]);
                *reinterpret_cast<double2 *>(r + 2) = make_double2(coef[2], coef[3]);
                *reinterpret_cast<double2 *>(r + 4) = make_double2(xn[0], xn[1]);
                *reinterpret_cast<double2 *>(r + 6) = make_double2(xn[2], xn[3]);
                *reinterpret_cast<double2 *>(r + 8) = make_double2(xn[4], xn[5]);
            }
            double c[6][6];
#pragma unroll
            for (int i = 0; i < 6; ++i)
#pragma unroll
                for (int j = i; j < 6; ++j) {
                    const double base = (j < 3) ? ((i == j) ? coef[0] : coef[1])
                                                : ((i == j) ? coef[2] : 0.0);
                    c[i][j] = base + coef[3] * (xn[i] * xn[j]);  // (:170-175)
                }
            if (A.tangent == nullptr) {
                // stress-only call
            } else if (bulk) {
                double *row = s_tan + tid * 36;
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 6; j += 2) {
                        const double a = (i <= j) ? c[i][j] : c[j][i];
                        const double b = (i <= j + 1) ? c[i][j + 1] : c[j + 1][i];
                        *reinterpret_cast<double2 *>(row + i * 6 + j) = make_double2(a, b);
                    }
            } else {
                double *row = A.tangent + qp_parent * 36;
#pragma unroll
                for (int i = 0; i < 6; ++i)
#pragma unroll
                    for (int j = 0; j < 6; ++j)
                        row[i * 6 + j] = (i <= j) ? c[i][j] : c[j][i];
            }
        }
        if (bulk)
            fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0 && bulk) {
            if (A.cells == nullptr) {
                if (A.tangent != nullptr)
                    bulk_s2g(A.tangent + q0 * 36, s_tan, TILE * 36 * sizeof(double));
                bulk_s2g(A.stress_cur + q0 * 6, s_sig, TILE * 6 * sizeof(double));
            } else {
                const int *pc = A.cells + q0 / NQ;
#pragma unroll 4
                for (int lc = 0; lc < CPT; ++lc) {
                    const size_t pq = (size_t)pc[lc] * NQ;
                    if (A.tangent != nullptr)
                        bulk_s2g(A.tangent + pq * 36, s_tan + lc * NQ * 36, NQ * 36 * sizeof(double));
                    bulk_s2g(A.stress_cur + pq * 6, s_sig + lc * NQ * 6, NQ * 6 * sizeof(double));
                }
            }
            bulk_s2g(A.eps1 + q0 * 6, s_eps, TILE * 6 * sizeof(double));
            bulk_s2g(A.alpha1 + q0, s_alp, TILE * sizeof(double));
            bulk_commit();
        }
        tile = next;
    }
    if (tid == 0)
        bulk_wait_read_all();
}

}  // namespace fcx
