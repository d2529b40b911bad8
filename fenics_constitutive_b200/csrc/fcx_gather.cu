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
//     grad[i][j]  = sum_k Jinv[k][i] * T[k][j]              (grad_at_qp, fcx_fem.cuh)
// One thread owns one QUADRATURE POINT (two of a cell's four for 4-point rules) of a tile of
// whole cells (64, 96 or 128 QPs).
// The tile's dofmap rows and Jinv blocks are contiguous ranges and arrive by 1-D
// bulk async copies (TMA) one tile ahead; the nodal increments are gathered
// straight from L1/L2 (the NQ threads of a cell hit the same sectors; two load
// instructions per 3-D node, load_node) and accumulated on the fly (9
// accumulators, no per-cell register array, so the kernel runs at high
// occupancy); the tile's [TILE][g][g] result block is staged in shared memory
// (double-buffered) and leaves as one bulk async store.
#include <cuda_runtime.h>

#include <mutex>
#include <type_traits>

#include "../../include/fcx.h"
#include "fcx_fem.cuh"
#include "fcx_internal.h"
#include "fcx_ptx.cuh"

namespace fcx {

struct GatherArgs {
    const int *dofmap;
    const double *u;
    const double *u_prev;  // or nullptr
    const double *dphi_ref;
    const double *Jinv;
    double *grad;
    unsigned long long ncells;
    unsigned long long *ticket;
    int bulk_ok;  // dofmap, Jinv, grad 16-byte aligned
};

template <int G, int ND, int NQ>
struct GatherCfg {
    // QPs per thread: a thread that owns two of a cell's four points fetches every nodal increment
    // once for both -- the kernel is bound by instruction issue (loads, address arithmetic), not
    // by DRAM (profiles/r1n: 650 thread-instructions per QP with one point per thread).
    static constexpr int QPT = (NQ == 4) ? 2 : 1;
    static constexpr int THREADS = fem_tile<NQ>();
    static constexpr int TILE = THREADS * QPT;  // QPs per tile
    static constexpr int CPT = TILE / NQ;
    static constexpr int GG = G * G;
    static constexpr int DOF_DBL = (CPT * ND + 1) / 2;
    static constexpr size_t smem_bytes =
        sizeof(double) * (2 * (size_t)TILE * GG + CPT * GG + DOF_DBL + NQ * ND * G + 1);
};

template <int G, int ND, int NQ>
__global__ void __launch_bounds__(fem_tile<NQ>())
    gather_kernel(const __grid_constant__ GatherArgs A)
{
    using Cfg = GatherCfg<G, ND, NQ>;
    constexpr int TILE = Cfg::TILE, CPT = Cfg::CPT, GG = Cfg::GG, QPT = Cfg::QPT, NT = Cfg::THREADS;
    constexpr int TPC = NQ / QPT;  // threads per cell
    static_assert(CPT % 16 == 0, "tile ranges must be 16-byte multiples");
    extern __shared__ __align__(128) double smem[];
    double *s_out = smem;                                   // [2][TILE][GG]
    double *s_jinv = s_out + 2 * TILE * GG;                 // [CPT][GG]
    int *s_dof = reinterpret_cast<int *>(s_jinv + CPT * GG);  // [CPT][ND]
    double *s_tab = s_jinv + CPT * GG + Cfg::DOF_DBL;       // [NQ][ND][G]
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_tab + NQ * ND * G);
    __shared__ unsigned long long s_next[2];  // slot = iteration parity (one CTA barrier per tile)

    const int tid = threadIdx.x;
    for (int i = tid; i < NQ * ND * G; i += NT)
        s_tab[i] = A.dphi_ref[i];
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const unsigned long long ntiles = (A.ncells + CPT - 1) / CPT;
    auto is_bulk = [&](unsigned long long t) { return A.bulk_ok && (t + 1) * CPT <= A.ncells; };
    auto issue = [&](unsigned long long t) {
        const unsigned long long c0 = t * CPT;
        mbar_arrive_expect_tx(bar, (uint32_t)(sizeof(double) * CPT * GG + sizeof(int) * CPT * ND));
        bulk_g2s(s_jinv, A.Jinv + c0 * GG, (uint32_t)(sizeof(double) * CPT * GG), bar);
        bulk_g2s(s_dof, A.dofmap + c0 * ND, (uint32_t)(sizeof(int) * CPT * ND), bar);
    };

    uint32_t parity = 0;
    int buf = 0, it = 0;
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
            mbar_wait(bar, parity);
            parity ^= 1;
        } else {  // ragged last tile / unaligned views
            for (int i = tid; i < ncell * GG; i += NT)
                s_jinv[i] = A.Jinv[c0 * GG + i];
            for (int i = tid; i < ncell * ND; i += NT)
                s_dof[i] = A.dofmap[c0 * ND + i];
            __syncthreads();
        }
        {
            const int lc = tid / TPC, q = (tid - lc * TPC) * QPT;  // first of this thread's QPT points
            const int lq = lc * NQ + q;                            // tile-local QP index
            if (lq < cnt) {                                        // whole cells: all QPT points valid
                double K[GG], g[QPT][GG];
#pragma unroll
                for (int i = 0; i < GG; ++i)
                    K[i] = s_jinv[lc * GG + i];
                grads_of_increment<G, ND, QPT>(s_tab + q * ND * G, K, s_dof + lc * ND, A.u, A.u_prev, g);
                double *o = bulk ? s_out + (buf * TILE + lq) * GG : A.grad + (q0 + lq) * GG;
#pragma unroll
                for (int qq = 0; qq < QPT; ++qq)
#pragma unroll
                    for (int i = 0; i < GG; ++i)
                        o[qq * GG + i] = g[qq][i];
            }
        }
        if (bulk)
            fence_proxy_async_smem();
        __syncthreads();  // results staged; dofmap / Jinv stage consumed
        const unsigned long long next = s_next[it];
        it ^= 1;
        const bool next_bulk = next < ntiles && is_bulk(next);
        if (tid == 0) {
            if (bulk) {
                bulk_s2g(A.grad + q0 * GG, s_out + buf * TILE * GG, (uint32_t)(sizeof(double) * TILE * GG));
                bulk_commit();
                bulk_wait_read_1();  // the store issued one tile ago has left the other buffer
            }
            if (next_bulk)
                issue(next);
        }
        if (!bulk)
            __syncthreads();  // plain-path stage reuse
        buf ^= bulk ? 1 : 0;
        tile = next;
        bulk = next_bulk;
    }
    if (tid == 0)
        bulk_wait_read_all();
}

// ---------------------------------------------------------------------------
// gather_staged_kernel: the same map with the NODAL VALUES staged one tile ahead.
//
// gather_kernel above is bound by the latency of its nodal loads: 16 resident warps per SM
// (100 registers, 23 KB of shared memory per 64-thread CTA) each waiting on L2 most of the time
// (ncu profiles/r1y: 9 of 11 stall cycles per issue are long-scoreboard, DRAM at 33 %, issue
// slots at 30 %).  Here the gathers are per-thread 8-byte cp.async copies (LDGSTS) into shared
// memory -- they cost an issue slot but no register and no stall -- issued for tile i+1 while
// tile i is computed, so a CTA keeps a whole tile of nodal values (7.7 KB for 32 P2 tets) in
// flight all the time:
//     A(i): bulk copies of dofmap / Jinv rows of tile i            (thread 0, two tiles ahead)
//     B(i): cp.async of du[dofmap(i)] -> s_du[i & 1]               (all threads, one tile ahead)
//     C(i): grad_at_qps from shared memory -> s_out, one bulk store
// Same fma chains in the same order as gather_kernel (grad_at_qps), hence the same bits.
// Whole 16-byte-aligned tiles only; launch_gather sends the ragged tail (and unaligned views)
// to gather_kernel.
// Measured, 998 250 P2 tets: 0.104 ms against 0.127 (profiles/r1za_gather_ab.jsonl); ncu
// (profiles/r1zh_gather_ncu_full.json): long-scoreboard stalls gone, now bound by the
// shared-memory data pipe (l1tex wavefronts at 78 % of peak: 255 LDS + 78 STS + 144 LDGSTS + 180
// bulk-copy wavefronts per 32-cell tile), DRAM at 48 %.  Tried on top and reverted because slower
// (0.110 ms, profiles/r1zi_*): lane = cell / warp = quadrature point with the basis-gradient
// table as constant-bank DFMA operands, odd-stride nodal stage, rotated 16-byte result stores --
// the table loads it removes were broadcasts that cost one wavefront each, while the padded
// stage made the LDGSTS writes less regular.  Also measured slower (0.119 ms, profiles/r2p_*): the
// staging loop node by node (address arithmetic once per node instead of once per value: 33 % fewer
// instructions) -- a warp's copy then touches 32 different nodes instead of 11 and the shared-memory
// wavefronts go up by a quarter: this kernel is bound by L1 / shared-memory wavefronts, not by issue.
// ---------------------------------------------------------------------------
template <int G, int ND, int NQ, bool PREV>
struct StagedCfg {
    using Base = GatherCfg<G, ND, NQ>;
    static constexpr int NDU = Base::CPT * ND * G;  // nodal doubles per tile
    static constexpr int NVEC = PREV ? 2 : 1;
    static constexpr int DOF_DBL = Base::DOF_DBL;
    static constexpr size_t smem_bytes =
        sizeof(double) * ((size_t)Base::TILE * Base::GG + 2 * NVEC * NDU + 3 * Base::CPT * Base::GG +
                          3 * DOF_DBL + NQ * ND * G + 4);
};

template <int G, int ND, int NQ, bool PREV>
__global__ void __launch_bounds__(fem_tile<NQ>())
    gather_staged_kernel(const __grid_constant__ GatherArgs A)
{
    using Cfg = GatherCfg<G, ND, NQ>;
    using SC = StagedCfg<G, ND, NQ, PREV>;
    constexpr int TILE = Cfg::TILE, CPT = Cfg::CPT, GG = Cfg::GG, QPT = Cfg::QPT, NT = Cfg::THREADS;
    constexpr int TPC = NQ / QPT, NDU = SC::NDU, NVEC = SC::NVEC;
    static_assert(CPT % 16 == 0, "tile ranges must be 16-byte multiples");
    extern __shared__ __align__(128) double smem[];
    double *s_out = smem;                          // [TILE][GG]
    double *s_du = s_out + TILE * GG;              // [2][NVEC][NDU]
    double *s_jinv = s_du + 2 * NVEC * NDU;        // [3][CPT][GG]
    double *s_dofd = s_jinv + 3 * CPT * GG;        // [3][CPT][ND] ints
    double *s_tab = s_dofd + 3 * SC::DOF_DBL;      // [NQ][ND][G]
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_tab + NQ * ND * G);  // [3]
    __shared__ unsigned long long s_tile[4];
    auto dof_slot = [&](int s) { return reinterpret_cast<int *>(s_dofd + s * SC::DOF_DBL); };

    const int tid = threadIdx.x;
    const unsigned long long ntiles = A.ncells / CPT;  // whole tiles only
    for (int i = tid; i < NQ * ND * G; i += NT)
        s_tab[i] = A.dphi_ref[i];
    auto issue_A = [&](unsigned long long t, int s) {
        const unsigned long long c0 = t * CPT;
        mbar_arrive_expect_tx(bar + s, (uint32_t)(sizeof(double) * CPT * GG + sizeof(int) * CPT * ND));
        bulk_g2s(s_jinv + s * CPT * GG, A.Jinv + c0 * GG, (uint32_t)(sizeof(double) * CPT * GG), bar + s);
        bulk_g2s(dof_slot(s), A.dofmap + c0 * ND, (uint32_t)(sizeof(int) * CPT * ND), bar + s);
    };
    auto next_ticket = [&]() -> unsigned long long {
        return gridDim.x + atomicAdd(A.ticket, 1ULL);
    };
    auto issue_B = [&](int s, int b) {
        const int *dm = dof_slot(s);
        double *dst = s_du + b * NVEC * NDU;
        for (int e = tid; e < NDU; e += NT) {
            const int cn = e / G, j = e - cn * G;
            const size_t src = (size_t)dm[cn] * G + j;
            cp_async_8(dst + e, A.u + src);
            if (PREV)
                cp_async_8(dst + NDU + e, A.u_prev + src);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < 3; ++s)
            mbar_init(bar + s, 1);
        fence_mbar_init();
        const unsigned long long t0 = blockIdx.x, t1 = next_ticket();
        s_tile[0] = t0;
        s_tile[1] = t1;
        issue_A(t0, 0);  // grid <= ntiles: the first tile always exists
        if (t1 < ntiles)
            issue_A(t1, 1);
    }
    __syncthreads();
    uint32_t phase = 0;  // bit s = parity of the next completion of bar[s]
    mbar_wait(bar + 0, 0);
    phase ^= 1u;
    issue_B(0, 0);
    cp_async_commit();

    for (int i = 0;; ++i) {
        const unsigned long long tile = s_tile[i & 3];
        if (tile >= ntiles)
            break;
        const int s0 = i % 3, s1 = (i + 1) % 3, s2 = (i + 2) % 3;
        if (tid == 0) {
            const unsigned long long t2 = next_ticket();
            s_tile[(i + 2) & 3] = t2;
            if (t2 < ntiles)
                issue_A(t2, s2);
        }
        const unsigned long long tile1 = s_tile[(i + 1) & 3];
        if (tile1 < ntiles) {
            mbar_wait(bar + s1, (phase >> s1) & 1u);
            phase ^= 1u << s1;
            issue_B(s1, (i + 1) & 1);
        }
        cp_async_commit();
        cp_async_wait<1>();  // this thread's share of B(i) has landed
        if (tid == 0)
            bulk_wait_read_all();  // the previous tile's store has left s_out
        __syncthreads();           // B(i) complete for every thread; s_out free
        {
            const int lc = tid / TPC, q = (tid - lc * TPC) * QPT;
            const int lq = lc * NQ + q;
            double K[GG], g[QPT][GG];
            const double *kin = s_jinv + s0 * CPT * GG + lc * GG;
#pragma unroll
            for (int k = 0; k < GG; ++k)
                K[k] = kin[k];
            const double *du = s_du + (i & 1) * NVEC * NDU + lc * ND * G;
            grad_at_qps<G, ND, QPT>(
                s_tab + q * ND * G, K,
                [&](int a, double *v) {
#pragma unroll
                    for (int j = 0; j < G; ++j)
                        v[j] = PREV ? du[a * G + j] - du[NDU + a * G + j] : du[a * G + j];
                },
                g);
            double *o = s_out + lq * GG;
#pragma unroll
            for (int qq = 0; qq < QPT; ++qq)
#pragma unroll
                for (int k = 0; k < GG; ++k)
                    o[qq * GG + k] = g[qq][k];
        }
        fence_proxy_async_smem();
        __syncthreads();  // results staged; s_du[i & 1], Jinv slot s0 consumed
        if (tid == 0) {
            bulk_s2g(A.grad + tile * TILE * GG, s_out, (uint32_t)(sizeof(double) * TILE * GG));
            bulk_commit();
        }
    }
    cp_async_wait<0>();
    if (tid == 0)
        bulk_wait_read_all();
}

// ---------------------------------------------------------------------------
// gather_cell_kernel: ONE THREAD PER CELL, basis-gradient table in the constant bank.
//
// ncu on gather_staged_kernel (profiles/r1zh_gather_ncu_full.json) shows the shared-memory data
// pipe as its limiter: per 32-cell tile 255 LDS + 78 STS + 144 LDGSTS + 180 bulk-copy wavefronts,
// of which 120 LDS wavefronts only re-read the basis-gradient table and another 120 read every
// nodal value twice (two threads per cell).  Here
//   * a thread owns a whole cell (all NQ = 4 quadrature points): every nodal value is read from
//     shared memory once, as 16-byte loads whose 240-byte lane stride is bank-conflict free;
//   * the table dphi_ref sits in __constant__ memory (copied device-to-device on the launch
//     stream), so with the quadrature-point and node loops unrolled its entries are constant-bank
//     operands of the DFMAs: no load instruction, no register, no wavefront;
//   * each thread writes its cell's 36 results as one 288-byte row with a padded row stride
//     (38 doubles: conflict-free 16-byte stores) and ships it with its own bulk async store --
//     rows are thread-private, so the output needs no CTA barrier.
// Staging of dofmap / Jinv rows (bulk copies, two tiles ahead) and of the nodal values (8-byte
// cp.async, one tile ahead) is gather_staged_kernel's.  Same fma chains in the same order as
// grad_at_qp (a ascending, then the Jinv contraction), hence the same bits as the other kernels.
// 3-D cells with 4-point rules only (P2 / P1 tetrahedra, q_degree 2); whole tiles of 64 cells.
// ---------------------------------------------------------------------------
constexpr int GC_MAX_TAB = 4 * 10 * 3;
__constant__ double c_gather_tab[GC_MAX_TAB];  // dphi_ref [NQ][ND][3] of the launch in flight

template <int ND, bool PREV>
struct CellGatherCfg {
    static constexpr int NQ = 4, G = 3, GG = 9;
    static constexpr int CPT = 64;              // cells per tile = threads per CTA
    static constexpr int ROW = 38;              // doubles per staged output row (36 used)
    static constexpr int NDU = CPT * ND * G;    // nodal doubles per tile
    static constexpr int NVEC = PREV ? 2 : 1;
    static constexpr int DOF_DBL = (CPT * ND + 1) / 2;
    static constexpr size_t smem_bytes =
        sizeof(double) * ((size_t)CPT * ROW + 2 * NVEC * NDU + 3 * CPT * GG + 3 * DOF_DBL + 4);
};

template <int ND, bool PREV>
__global__ void __launch_bounds__(64)
    gather_cell_kernel(const __grid_constant__ GatherArgs A)
{
    using SC = CellGatherCfg<ND, PREV>;
    constexpr int NQ = 4, G = 3, GG = 9, CPT = SC::CPT, NT = SC::CPT, NDU = SC::NDU, NVEC = SC::NVEC, ROW = SC::ROW;
    static_assert((ND * G) % 2 == 0, "a cell's nodal block must be a whole number of 16-byte pairs");
    extern __shared__ __align__(128) double smem[];
    double *s_out = smem;                          // [CPT][ROW]
    double *s_du = s_out + CPT * ROW;              // [2][NVEC][NDU]
    double *s_jinv = s_du + 2 * NVEC * NDU;        // [3][CPT][GG]
    double *s_dofd = s_jinv + 3 * CPT * GG;        // [3][CPT][ND] ints
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_dofd + 3 * SC::DOF_DBL);  // [3]
    __shared__ unsigned long long s_tile[4];
    auto dof_slot = [&](int s) { return reinterpret_cast<int *>(s_dofd + s * SC::DOF_DBL); };

    const int tid = threadIdx.x;
    const unsigned long long ntiles = A.ncells / CPT;  // whole tiles only
    auto issue_A = [&](unsigned long long t, int s) {
        const unsigned long long c0 = t * CPT;
        mbar_arrive_expect_tx(bar + s, (uint32_t)(sizeof(double) * CPT * GG + sizeof(int) * CPT * ND));
        bulk_g2s(s_jinv + s * CPT * GG, A.Jinv + c0 * GG, (uint32_t)(sizeof(double) * CPT * GG), bar + s);
        bulk_g2s(dof_slot(s), A.dofmap + c0 * ND, (uint32_t)(sizeof(int) * CPT * ND), bar + s);
    };
    auto next_ticket = [&]() -> unsigned long long { return gridDim.x + atomicAdd(A.ticket, 1ULL); };
    auto issue_B = [&](int s, int b) {
        const int *dm = dof_slot(s);
        double *dst = s_du + b * NVEC * NDU;
        for (int e = tid; e < NDU; e += NT) {
            const int cn = e / G, j = e - cn * G;
            const size_t src = (size_t)dm[cn] * G + j;
            cp_async_8(dst + e, A.u + src);
            if (PREV)
                cp_async_8(dst + NDU + e, A.u_prev + src);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < 3; ++s)
            mbar_init(bar + s, 1);
        fence_mbar_init();
        const unsigned long long t0 = blockIdx.x, t1 = next_ticket();
        s_tile[0] = t0;
        s_tile[1] = t1;
        issue_A(t0, 0);  // grid <= ntiles: the first tile always exists
        if (t1 < ntiles)
            issue_A(t1, 1);
    }
    __syncthreads();
    uint32_t phase = 0;  // bit s = parity of the next completion of bar[s]
    mbar_wait(bar + 0, 0);
    phase ^= 1u;
    issue_B(0, 0);
    cp_async_commit();

    for (int i = 0;; ++i) {
        const unsigned long long tile = s_tile[i & 3];
        if (tile >= ntiles)
            break;
        const int s0 = i % 3, s1 = (i + 1) % 3, s2 = (i + 2) % 3;
        if (tid == 0) {
            const unsigned long long t2 = next_ticket();
            s_tile[(i + 2) & 3] = t2;
            if (t2 < ntiles)
                issue_A(t2, s2);
        }
        const unsigned long long tile1 = s_tile[(i + 1) & 3];
        if (tile1 < ntiles) {
            mbar_wait(bar + s1, (phase >> s1) & 1u);
            phase ^= 1u << s1;
            issue_B(s1, (i + 1) & 1);
        }
        cp_async_commit();
        cp_async_wait<1>();    // this thread's share of B(i) has landed
        bulk_wait_read_all();  // this thread's previous row has left shared memory
        __syncthreads();       // B(i) complete for every thread
        {
            double K[GG];
            const double *kin = s_jinv + s0 * CPT * GG + tid * GG;
#pragma unroll
            for (int k = 0; k < GG; ++k)
                K[k] = kin[k];
            const double2 *du2 = reinterpret_cast<const double2 *>(s_du + (i & 1) * NVEC * NDU + tid * ND * G);
            double T[NQ][G][G];
#pragma unroll
            for (int q = 0; q < NQ; ++q)
#pragma unroll
                for (int k = 0; k < G; ++k)
#pragma unroll
                    for (int j = 0; j < G; ++j)
                        T[q][k][j] = 0.0;
            // two nodes (six doubles, three 16-byte loads) at a time
#pragma unroll
            for (int a2 = 0; a2 < ND / 2; ++a2) {
                double v[6];
#pragma unroll
                for (int h = 0; h < 3; ++h) {
                    double2 x = du2[a2 * 3 + h];
                    if (PREV) {
                        const double2 y = du2[NDU / 2 + a2 * 3 + h];
                        x.x -= y.x;
                        x.y -= y.y;
                    }
                    v[2 * h] = x.x;
                    v[2 * h + 1] = x.y;
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int a = 2 * a2 + half;
#pragma unroll
                    for (int q = 0; q < NQ; ++q)
#pragma unroll
                        for (int k = 0; k < G; ++k) {
                            const double d = c_gather_tab[(q * ND + a) * G + k];
#pragma unroll
                            for (int j = 0; j < G; ++j)
                                T[q][k][j] = fma(d, v[3 * half + j], T[q][k][j]);
                        }
                }
            }
            double *o = s_out + tid * ROW;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                double g[GG];
#pragma unroll
                for (int ii = 0; ii < G; ++ii)
#pragma unroll
                    for (int j = 0; j < G; ++j) {
                        double acc = 0.0;
#pragma unroll
                        for (int k = 0; k < G; ++k)
                            acc = fma(K[k * G + ii], T[q][k][j], acc);
                        g[ii * G + j] = acc;
                    }
                // 9 doubles per point: written as pairs across the point boundaries below
#pragma unroll
                for (int k = 0; k < GG; ++k)
                    T[q][k / G][k % G] = g[k];  // reuse the accumulators as the result
            }
#pragma unroll
            for (int p = 0; p < NQ * GG / 2; ++p) {
                const int e0 = 2 * p, e1 = 2 * p + 1;
                *reinterpret_cast<double2 *>(o + e0) =
                    make_double2(T[e0 / GG][(e0 % GG) / G][e0 % G], T[e1 / GG][(e1 % GG) / G][e1 % G]);
            }
        }
        fence_proxy_async_smem();
        // the row is this thread's own: no barrier before its bulk store
        bulk_s2g(A.grad + (tile * CPT + tid) * (NQ * GG), s_out + tid * ROW, (uint32_t)(sizeof(double) * NQ * GG));
        bulk_commit();
        __syncthreads();  // s_du[i & 1] and Jinv / dofmap slot s0 consumed by every thread
    }
    cp_async_wait<0>();
    bulk_wait_read_all();
}

// ---------------------------------------------------------------------------
// gather_wq_kernel: gather_staged_kernel's staging and occupancy (two threads per cell, 6 CTAs per SM) with
// a WARP-UNIFORM quadrature-point pair: warp w of the CTA takes points {2w, 2w+1} of the tile's 32 cells
// (lane = cell), so the basis-gradient table entries are compile-time indices into __constant__ memory --
// operands of the DFMAs instead of 120 of the 255 LDS wavefronts per tile -- each thread reads its cell's
// nodal values with conflict-free 16-byte loads, and writes its 18 results into a padded row (38 doubles per
// cell: conflict-free 16-byte stores); warp 0 ships the 32 rows with one 288-byte bulk store each.
// Same fma chains in the same order (grad_at_qps), hence the same bits.  3-D, 4-point rules, whole 32-cell
// tiles; fcx_tune "gather_variant" 3.
// ---------------------------------------------------------------------------
template <int ND, bool PREV>
struct WqGatherCfg {
    static constexpr int NQ = 4, G = 3, GG = 9;
    static constexpr int CPT = 32, NT = 64;
    static constexpr int ROW = 38;
    static constexpr int NDU = CPT * ND * G;
    static constexpr int NVEC = PREV ? 2 : 1;
    static constexpr int DOF_DBL = (CPT * ND + 1) / 2;
    static constexpr size_t smem_bytes =
        sizeof(double) * ((size_t)CPT * ROW + 2 * NVEC * NDU + 3 * CPT * GG + 3 * DOF_DBL + 4);
};

template <int ND, bool PREV, int Q0>
__device__ __forceinline__ void wq_compute(const double *__restrict__ kin, const double2 *__restrict__ du2,
                                           int ndu_half, double *__restrict__ orow)
{
    constexpr int G = 3, GG = 9;
    double K[GG];
#pragma unroll
    for (int k = 0; k < GG; ++k)
        K[k] = kin[k];
    double T[2][G][G];
#pragma unroll
    for (int qq = 0; qq < 2; ++qq)
#pragma unroll
        for (int k = 0; k < G; ++k)
#pragma unroll
            for (int j = 0; j < G; ++j)
                T[qq][k][j] = 0.0;
#pragma unroll
    for (int a2 = 0; a2 < ND / 2; ++a2) {
        double v[6];
#pragma unroll
        for (int h = 0; h < 3; ++h) {
            double2 x = du2[a2 * 3 + h];
            if (PREV) {
                const double2 y = du2[ndu_half + a2 * 3 + h];
                x.x -= y.x;
                x.y -= y.y;
            }
            v[2 * h] = x.x;
            v[2 * h + 1] = x.y;
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int a = 2 * a2 + half;
#pragma unroll
            for (int qq = 0; qq < 2; ++qq)
#pragma unroll
                for (int k = 0; k < G; ++k) {
                    const double d = c_gather_tab[((Q0 + qq) * ND + a) * G + k];
#pragma unroll
                    for (int j = 0; j < G; ++j)
                        T[qq][k][j] = fma(d, v[3 * half + j], T[qq][k][j]);
                }
        }
    }
    double g[2 * GG];
#pragma unroll
    for (int qq = 0; qq < 2; ++qq)
#pragma unroll
        for (int ii = 0; ii < G; ++ii)
#pragma unroll
            for (int j = 0; j < G; ++j) {
                double acc = 0.0;
#pragma unroll
                for (int k = 0; k < G; ++k)
                    acc = fma(K[k * G + ii], T[qq][k][j], acc);
                g[qq * GG + ii * G + j] = acc;
            }
#pragma unroll
    for (int p = 0; p < GG; ++p)
        *reinterpret_cast<double2 *>(orow + Q0 * GG + 2 * p) = make_double2(g[2 * p], g[2 * p + 1]);
}

template <int ND, bool PREV>
__global__ void __launch_bounds__(64)
    gather_wq_kernel(const __grid_constant__ GatherArgs A)
{
    using SC = WqGatherCfg<ND, PREV>;
    constexpr int NQ = 4, G = 3, GG = 9, CPT = SC::CPT, NT = SC::NT, NDU = SC::NDU, NVEC = SC::NVEC, ROW = SC::ROW;
    static_assert((ND * G) % 2 == 0, "a cell's nodal block must be a whole number of 16-byte pairs");
    extern __shared__ __align__(128) double smem[];
    double *s_out = smem;                          // [CPT][ROW]
    double *s_du = s_out + CPT * ROW;              // [2][NVEC][NDU]
    double *s_jinv = s_du + 2 * NVEC * NDU;        // [3][CPT][GG]
    double *s_dofd = s_jinv + 3 * CPT * GG;        // [3][CPT][ND] ints
    uint64_t *bar = reinterpret_cast<uint64_t *>(s_dofd + 3 * SC::DOF_DBL);  // [3]
    __shared__ unsigned long long s_tile[4];
    auto dof_slot = [&](int s) { return reinterpret_cast<int *>(s_dofd + s * SC::DOF_DBL); };

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned long long ntiles = A.ncells / CPT;  // whole tiles only
    auto issue_A = [&](unsigned long long t, int s) {
        const unsigned long long c0 = t * CPT;
        mbar_arrive_expect_tx(bar + s, (uint32_t)(sizeof(double) * CPT * GG + sizeof(int) * CPT * ND));
        bulk_g2s(s_jinv + s * CPT * GG, A.Jinv + c0 * GG, (uint32_t)(sizeof(double) * CPT * GG), bar + s);
        bulk_g2s(dof_slot(s), A.dofmap + c0 * ND, (uint32_t)(sizeof(int) * CPT * ND), bar + s);
    };
    auto next_ticket = [&]() -> unsigned long long { return gridDim.x + atomicAdd(A.ticket, 1ULL); };
    auto issue_B = [&](int s, int b) {
        const int *dm = dof_slot(s);
        double *dst = s_du + b * NVEC * NDU;
        for (int e = tid; e < NDU; e += NT) {
            const int cn = e / G, j = e - cn * G;
            const size_t src = (size_t)dm[cn] * G + j;
            cp_async_8(dst + e, A.u + src);
            if (PREV)
                cp_async_8(dst + NDU + e, A.u_prev + src);
        }
    };
    if (tid == 0) {
        for (int s = 0; s < 3; ++s)
            mbar_init(bar + s, 1);
        fence_mbar_init();
        const unsigned long long t0 = blockIdx.x, t1 = next_ticket();
        s_tile[0] = t0;
        s_tile[1] = t1;
        issue_A(t0, 0);
        if (t1 < ntiles)
            issue_A(t1, 1);
    }
    __syncthreads();
    uint32_t phase = 0;
    mbar_wait(bar + 0, 0);
    phase ^= 1u;
    issue_B(0, 0);
    cp_async_commit();

    for (int i = 0;; ++i) {
        const unsigned long long tile = s_tile[i & 3];
        if (tile >= ntiles)
            break;
        const int s0 = i % 3, s1 = (i + 1) % 3, s2 = (i + 2) % 3;
        if (tid == 0) {
            const unsigned long long t2 = next_ticket();
            s_tile[(i + 2) & 3] = t2;
            if (t2 < ntiles)
                issue_A(t2, s2);
        }
        const unsigned long long tile1 = s_tile[(i + 1) & 3];
        if (tile1 < ntiles) {
            mbar_wait(bar + s1, (phase >> s1) & 1u);
            phase ^= 1u << s1;
            issue_B(s1, (i + 1) & 1);
        }
        cp_async_commit();
        cp_async_wait<1>();    // this thread's share of B(i) has landed
        if (wid == 0)
            bulk_wait_read_all();  // the row this lane shipped one tile ago has left shared memory
        __syncthreads();           // B(i) complete for every thread; s_out free
        {
            const double *kin = s_jinv + s0 * CPT * GG + lane * GG;
            const double2 *du2 = reinterpret_cast<const double2 *>(s_du + (i & 1) * NVEC * NDU + lane * ND * G);
            double *orow = s_out + lane * ROW;
            if (wid == 0)
                wq_compute<ND, PREV, 0>(kin, du2, NDU / 2, orow);
            else
                wq_compute<ND, PREV, 2>(kin, du2, NDU / 2, orow);
        }
        fence_proxy_async_smem();
        __syncthreads();  // both halves of every row staged; s_du[i & 1], Jinv slot s0 consumed
        if (wid == 0) {
            bulk_s2g(A.grad + (tile * CPT + lane) * (NQ * GG), s_out + lane * ROW, (uint32_t)(sizeof(double) * NQ * GG));
            bulk_commit();
        }
    }
    cp_async_wait<0>();
    if (wid == 0)
        bulk_wait_read_all();
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
static int launch_gather_plain(size_t ncells, const int *dofmap, const double *u, const double *u_prev,
                               const double *dphi, const double *Jinv, double *grad, cudaStream_t st)
{
    using Cfg = GatherCfg<G, ND, NQ>;
    auto kern = gather_kernel<G, ND, NQ>;
    static OccCache cache;  // per device
    int occ = 1;
    if (int rc = kernel_occupancy(cache, kern, Cfg::THREADS, Cfg::smem_bytes, "occupancy(gather)", &occ))
        return rc;
    const unsigned long long ntiles = (ncells + Cfg::CPT - 1) / Cfg::CPT;
    const int per_sm = tuned_ctas_per_sm() > 0 ? tuned_ctas_per_sm() : occ;
    unsigned long long grid = (unsigned long long)sm_count() * per_sm;
    if (grid > ntiles)
        grid = ntiles;
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    GatherArgs A{dofmap, u, u_prev, dphi, Jinv, grad, (unsigned long long)ncells,
                 (ntiles > grid) ? tile_ticket(st) : nullptr,
                 (al16(dofmap) && al16(Jinv) && al16(grad)) ? 1 : 0};
    kern<<<(unsigned)grid, Cfg::THREADS, Cfg::smem_bytes, st>>>(A);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "gather_kernel launch");
}

template <int G, int ND, int NQ, bool PREV>
static int launch_gather_staged(size_t nfull_cells, const int *dofmap, const double *u, const double *u_prev,
                                const double *dphi, const double *Jinv, double *grad, cudaStream_t st)
{
    using Cfg = GatherCfg<G, ND, NQ>;
    using SC = StagedCfg<G, ND, NQ, PREV>;
    auto kern = gather_staged_kernel<G, ND, NQ, PREV>;
    static OccCache cache;  // per device
    int occ = 1;
    if (int rc = kernel_occupancy(cache, kern, Cfg::THREADS, SC::smem_bytes, "occupancy(gather staged)", &occ))
        return rc;
    const unsigned long long ntiles = nfull_cells / Cfg::CPT;
    const int per_sm = tuned_ctas_per_sm() > 0 ? tuned_ctas_per_sm() : occ;
    unsigned long long grid = (unsigned long long)sm_count() * per_sm;
    if (grid > ntiles)
        grid = ntiles;
    unsigned long long *ticket = tile_ticket(st);
    if (ticket == nullptr)
        return -1;  // dynamic tiles switched off: caller falls back to gather_kernel
    GatherArgs A{dofmap, u, u_prev, dphi, Jinv, grad, (unsigned long long)nfull_cells, ticket, 1};
    kern<<<(unsigned)grid, Cfg::THREADS, SC::smem_bytes, st>>>(A);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "gather_staged_kernel launch");
}

// One launch of gather_cell_kernel over `nfull_cells` (a multiple of 64).  The table is copied
// device-to-device into c_gather_tab on the launch stream; launches on other streams are ordered
// behind the previous user of the symbol with an event (no host synchronisation).
template <int ND, bool PREV, bool WQ>
static int launch_gather_cell(size_t nfull_cells, const int *dofmap, const double *u, const double *u_prev,
                              const double *dphi, const double *Jinv, double *grad, cudaStream_t st)
{
    using SC = std::conditional_t<WQ, WqGatherCfg<ND, PREV>, CellGatherCfg<ND, PREV>>;
    auto kern = WQ ? gather_wq_kernel<ND, PREV> : gather_cell_kernel<ND, PREV>;
    constexpr int THREADS = 64;
    static OccCache cache;  // per device
    int occ = 1;
    if (int rc = kernel_occupancy(cache, kern, THREADS, SC::smem_bytes, "occupancy(gather cell)", &occ))
        return rc;
    unsigned long long *ticket = tile_ticket(st);
    if (ticket == nullptr)
        return -1;  // dynamic tiles switched off: caller falls back
    static std::mutex mu;
    static cudaEvent_t last_use[FCX_MAX_DEVICES] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= FCX_MAX_DEVICES)
        return -1;
    std::lock_guard<std::mutex> lock(mu);
    cudaError_t e = cudaSuccess;
    if (last_use[dev] == nullptr)
        e = cudaEventCreateWithFlags(&last_use[dev], cudaEventDisableTiming);
    else
        e = cudaStreamWaitEvent(st, last_use[dev], 0);
    if (e == cudaSuccess)
        e = cudaMemcpyToSymbolAsync(c_gather_tab, dphi, sizeof(double) * 4 * ND * 3, 0, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess)
        return note_cuda_error(e, "gather table -> constant bank");
    const unsigned long long ntiles = nfull_cells / SC::CPT;
    const int per_sm = tuned_ctas_per_sm() > 0 ? tuned_ctas_per_sm() : occ;
    unsigned long long grid = (unsigned long long)sm_count() * per_sm;
    if (grid > ntiles)
        grid = ntiles;
    GatherArgs A{dofmap, u, u_prev, dphi, Jinv, grad, (unsigned long long)nfull_cells, ticket, 1};
    kern<<<(unsigned)grid, THREADS, SC::smem_bytes, st>>>(A);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaEventRecord(last_use[dev], st);
    return note_cuda_error(e, "gather_cell_kernel launch");
}

template <int G, int ND, int NQ>
static int launch_gather(size_t ncells, const int *dofmap, const double *u, const double *u_prev,
                         const double *dphi, const double *Jinv, double *grad, cudaStream_t st)
{
    using Cfg = GatherCfg<G, ND, NQ>;
    auto al16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    // thread-per-cell kernel (fcx_tune "gather_variant" 2): 3-D cells with 4-point rules, whole 64-cell tiles
    if constexpr (G == 3 && NQ == 4) {
        // 2: thread per cell (64-cell tiles), 3: warp-uniform quadrature-point pair (32-cell tiles)
        const bool wq = gather_variant() >= 3;
        const size_t cpt = wq ? 32 : 64;
        const size_t nfull64 = ncells / cpt * cpt;
        if (gather_variant() >= 2 && nfull64 >= cpt * 4 && al16(dofmap) && al16(Jinv) && al16(grad) && al16(dphi)) {
            int rc = wq ? (u_prev ? launch_gather_cell<ND, true, true>(nfull64, dofmap, u, u_prev, dphi, Jinv, grad, st)
                                  : launch_gather_cell<ND, false, true>(nfull64, dofmap, u, u_prev, dphi, Jinv, grad, st))
                        : (u_prev ? launch_gather_cell<ND, true, false>(nfull64, dofmap, u, u_prev, dphi, Jinv, grad, st)
                                  : launch_gather_cell<ND, false, false>(nfull64, dofmap, u, u_prev, dphi, Jinv, grad, st));
            if (rc != -1) {
                if (rc != FCX_OK || nfull64 == ncells)
                    return rc;
                return launch_gather_plain<G, ND, NQ>(ncells - nfull64, dofmap + nfull64 * ND, u, u_prev, dphi,
                                                      Jinv + nfull64 * G * G, grad + nfull64 * NQ * G * G, st);
            }
        }
    }
    const size_t nfull = ncells / Cfg::CPT * Cfg::CPT;
    // staged kernel (fcx_tune "gather_variant" 1): whole aligned tiles
    if (gather_variant() >= 1 && nfull >= (size_t)Cfg::CPT * 4 && al16(dofmap) && al16(Jinv) && al16(grad)) {
        int rc = u_prev ? launch_gather_staged<G, ND, NQ, true>(nfull, dofmap, u, u_prev, dphi, Jinv, grad, st)
                        : launch_gather_staged<G, ND, NQ, false>(nfull, dofmap, u, u_prev, dphi, Jinv, grad, st);
        if (rc != -1) {
            if (rc != FCX_OK || nfull == ncells)
                return rc;
            return launch_gather_plain<G, ND, NQ>(ncells - nfull, dofmap + nfull * ND, u, u_prev, dphi,
                                                  Jinv + nfull * G * G, grad + nfull * NQ * G * G, st);
        }
    }
    return launch_gather_plain<G, ND, NQ>(ncells, dofmap, u, u_prev, dphi, Jinv, grad, st);
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
