// fcx_krylov.cu -- device-resident Krylov loop of the stand-in NewtonSolver: single-reduction
// (Chronopoulos-Gear) Jacobi-preconditioned conjugate gradients whose reduction and ghost exchange
// run over NVLink PEER MEMORY from inside the kernels -- no NCCL call, no host round trip and no
// Python in the iteration.
//
// Where this sits: the reference leaves the linear solve of every Newton step to PETSc through
// dolfinx.nls.petsc.NewtonSolver (third-party); with the mesh partitioned over MPI
// ranks (solver/_solver.py:64-68, tests/solver/test_solver_mpi.py:93-121) PETSc exchanges the ghost
// values of the Krylov vectors and all-reduces the dot products.  The stand-in did the same with
// NCCL issued from Python: two all-reduces + one send/recv per iteration, 0.43 ms per iteration on
// two GPUs against 0.18 ms of kernels.  Here one iteration is
//     K1  element kernel        fe = B^T C B u_e             (fcx_tangent_apply[_rec], unchanged)
//     K2  gsum_dots_kernel      w = sum fe (node-wise, fixed order);  partial (r.u, w.u, r.r) over
//                               owned free dofs; the last CTA adds the CTA partials (all its threads, in
//                               a fixed order), STORES the three sums into every rank's reduction slot,
//                               waits for the other ranks' sums and adds them in RANK order (every rank
//                               gets the same bits): alpha / beta -- and the RESIDUAL TEST: once
//                               r.r <= rtol^2 r0.r0 the solve freezes itself (control block `ctl`), every
//                               later kernel of the loop returns at once, on every rank in the same
//                               iteration
//     K3  cg_update_kernel      p = u + beta p;  s = w + beta s;  x += alpha p;  r -= alpha s;
//                               u = minv r
//     K4  halo_push_kernel      u of the nodes that are ghosts elsewhere is stored straight into the
//                               neighbours' vectors; a system-scope flag tells them.  The WAIT for the
//                               neighbours' flags is a one-thread kernel that the next iteration issues
//                               BEHIND its interior cells (K1 runs on the cells that touch no ghost node
//                               first, MeshPartition orders them first), so the exchange hides behind
//                               ~90 % of the element work; then K1 on the boundary cells
// with  gamma = r.u, delta = w.u:  beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old)
// -- ONE reduction per iteration instead of two, one fused vector pass instead of three kernels.
// Six launches per iteration (three on one rank); fcx_krylov_iterate enqueues a block of iterations from C, and
// because the stop rule runs on the device the host enqueues the NEXT block before it reads the snapshot of the
// previous one (fcx_krylov_snapshot / fcx_krylov_wait_snapshot): the stream is never drained inside a solve.
// One process per GPU: peers' buffers come from CUDA IPC handles.
//
// Ordering between ranks: a reduction slot / flag pair is double-buffered by iteration parity; a rank
// can only run ahead to iteration i+2's K2 after every rank has finished iteration i's K3 (it needs
// their i+1 sums), so a slot is never overwritten while someone still reads it.  Ghost values for
// iteration i+1 are pushed after K3 of iteration i, which every neighbour can only reach after its own
// K1 of iteration i has finished reading (the reduction is a barrier) -- no write-after-read hazard.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/fcx.h"
#include "fcx_fem.cuh"
#include "fcx_internal.h"

namespace fcx {

constexpr int KR_THREADS = 256;
constexpr int KR_MAX_WORLD = 16;
constexpr int KR_HIST = 1 << 16;  // residual history entries (ring)
constexpr int KR_CTL = 8;         // doubles in the control block (see Krylov::ctl)

// One rank's communication block (a single cudaMalloc, exported through an IPC handle):
//   [ red[2][W][4] doubles | redflag[2][W] u64 | haloflag[W] u64 | pad to 256 B | u vector: n doubles ]
// The control part comes FIRST: its offsets depend on the world size only, so they are the same in
// every rank's block although the ranks hold different numbers of dofs.
struct CommLayout {
    size_t off_red, off_redflag, off_haloflag, off_u, bytes;
    __host__ __device__ CommLayout() : off_red(0), off_redflag(0), off_haloflag(0), off_u(0), bytes(0) {}
    __host__ CommLayout(size_t n, int world)
    {
        off_red = 0;
        off_redflag = off_red + sizeof(double) * 2 * world * 4;
        off_haloflag = off_redflag + sizeof(unsigned long long) * 2 * world;
        off_u = (off_haloflag + sizeof(unsigned long long) * world + 255) & ~(size_t)255;
        bytes = (off_u + n * sizeof(double) + 255) & ~(size_t)255;
    }
};

struct PeerPtrs {
    char *base[KR_MAX_WORLD];  // comm block of every rank (own block included), device-visible
};

struct Krylov {
    int rank = 0, world = 1, gdim = 3;
    size_t n = 0, nnodes = 0, n_owned = 0;  // dofs, nodes, owned dofs (the first n_owned)
    CommLayout lay;
    char *comm = nullptr;  // own comm block
    PeerPtrs peers{};
    bool peer_open[KR_MAX_WORLD] = {};
    // vectors (own allocations); u lives in the comm block (peers store its ghost values)
    double *x = nullptr, *r = nullptr, *w = nullptr, *p = nullptr, *s = nullptr, *minv = nullptr;
    double *partials = nullptr;          // [3][grid]
    unsigned *ticket = nullptr;          // [2]: K2 reduction, K4 completion
    double *state = nullptr;             // [2][4]: gamma, alpha, rr, breakdown flag per parity
    double *scal = nullptr;              // [2]: alpha, beta of the iteration in flight (K2 -> K3)
    double *hist = nullptr;              // [KR_HIST]: r.r at the start of iteration it
    int *err = nullptr;                  // sticky: 1 = a peer's flag never arrived (bounded spin timed out)
    // Control block [KR_CTL] of the solve in flight, written by K2's finishing thread only:
    //   [0] frozen (0 / 1): the residual test passed or the operator broke down -- every later kernel of the
    //       iteration loop returns at once (same decision on every rank: it is taken from the all-gathered sums)
    //   [1] iteration at which it froze   [2] r.r at the start of the latest live iteration   [3] r.r of iteration 0
    //   [4] breakdown (p.Ap <= 0) seen     [5] live iterations so far
    double *ctl = nullptr;
    unsigned long long *tile_tickets = nullptr;  // [2]: alternating tile-ticket counters of the element kernels
    unsigned tile_launches = 0;                  // launches they have served (parity = whose turn it is)
    double rtol2 = 0.0;                  // freeze when r.r <= rtol2 * (r.r of iteration 0); 0 = never
    double *snap_host = nullptr;         // pinned [2][KR_CTL + 1]: snapshots of ctl (+ err) for the host
    cudaEvent_t snap_ev[2] = {nullptr, nullptr};
    // halo plan (device): flattened send entries over all neighbours
    int n_nbr = 0, n_send = 0;
    int nbr_rank[KR_MAX_WORLD] = {};
    int *send_src = nullptr, *send_dst = nullptr, *send_nbr = nullptr;  // [n_send]: my node, peer's node, slot
    // operator
    int op_mode = 0;  // 3 = tangent records, 1 = dense tangents
    int sdim = 6, nq = 0, nd = 0;
    size_t ncells = 0;
    const int *dofmap = nullptr, *fe_pos = nullptr, *adj_idx = nullptr;
    const long long *adj_ptr = nullptr;
    const double *dphi = nullptr, *weights = nullptr, *Jinv = nullptr, *detJ = nullptr, *tang = nullptr;
    double *fe = nullptr;
    // progress
    unsigned long long epoch = 0;  // reductions / pushes done since creation (same on every rank)
    unsigned long long push_epoch = 0;  // epoch of the last ghost push
    bool push_waited = true;            // ... and whether the neighbours' pushes of that epoch have been waited for
    bool push_gated = false;            // ... and whether that push was an iteration's (skipped once the solve is frozen)
    size_t ncells_interior = 0;         // local cells [0, ncells_interior) touch no ghost node
    unsigned long long it = 0;     // iterations of the current solve
    unsigned grid = 1;
};

__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Flag stores are RELAXED system-scope stores behind ONE system-scope fence of the signalling thread (a
// release per flag would repeat that fence for every peer: 8 of them per reduction at 8 ranks).
__device__ __forceinline__ void st_flag(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Bounded spin on a flag (until it is >= want, or == want if `exact`): a peer that never arrives must not
// hang the GPU.  ~2^24 polls of a system-scope load are several seconds; on time-out the sticky error word
// is set and the kernel carries on with whatever it has (the host raises after the block).
__device__ __forceinline__ void wait_flag(const unsigned long long *p, unsigned long long want, bool exact, int *err)
{
    for (unsigned spin = 0; spin < (1u << 24); ++spin) {
        const unsigned long long v = ld_flag(p);
        if (exact ? v == want : v >= want)
            return;
    }
    atomicExch(err, 1);
}
__device__ __forceinline__ double ld_volatile(const double *p)
{
    double v;
    asm volatile("ld.volatile.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ double kr_block_sum(double v, double *sh)
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
        for (int k = 0; k < KR_THREADS / 32; ++k)
            t += sh[k];
    }
    __syncthreads();
    return t;  // valid in thread 0
}

// x = 0, r = rhs on free owned dofs (minv != 0) else 0, u = minv r, p = s = 0.
// u is written on OWNED dofs only (the first n_owned): the ghost entries belong to the neighbours, whose
// peer stores may land before this kernel runs.
__global__ void __launch_bounds__(KR_THREADS)
    kr_begin_kernel(size_t n, size_t n_owned, const double *__restrict__ rhs, const double *__restrict__ minv_in,
                    double *__restrict__ minv, double *__restrict__ x, double *__restrict__ r,
                    double *__restrict__ u, double *__restrict__ p, double *__restrict__ s, double *__restrict__ ctl)
{
    if (blockIdx.x == 0 && threadIdx.x < KR_CTL)
        ctl[threadIdx.x] = 0.0;  // a new solve: not frozen (the kernels of the previous solve are behind us in stream order)
    const size_t stride = (size_t)gridDim.x * KR_THREADS;
    for (size_t i = (size_t)blockIdx.x * KR_THREADS + threadIdx.x; i < n; i += stride) {
        const double m = minv_in[i];
        const double rv = m != 0.0 ? rhs[i] : 0.0;
        minv[i] = m;
        x[i] = 0.0;
        r[i] = rv;
        if (i < n_owned)
            u[i] = m * rv;
        p[i] = 0.0;
        s[i] = 0.0;
    }
}

// K2: w = node-wise sum of the element vectors (fixed order: deterministic, no atomics) and the
// three dot products of the iteration over the free owned dofs.
template <int G>
__global__ void __launch_bounds__(KR_THREADS)
    gsum_dots_kernel(const long long *__restrict__ adj_ptr, const int *__restrict__ adj_idx,
                     const double *__restrict__ fe, const double *__restrict__ r, const double *__restrict__ u,
                     const double *__restrict__ minv, double *__restrict__ w, unsigned long long nnodes,
                     double *partials, unsigned *ticket, PeerPtrs peers, CommLayout lay, int rank, int world,
                     unsigned long long epoch, int first, double *state, double *scal, double *hist,
                     unsigned long long it, int *err, double *ctl, double rtol2)
{
    __shared__ double sh[KR_THREADS / 32];
    __shared__ bool last;
    // frozen solve: nothing to do.  Uniform over the grid and over the ranks: ctl[0] is written by the finishing
    // thread of an EARLIER launch only, from sums every rank holds bit for bit.
    if (ctl[0] != 0.0)
        return;
    double g = 0.0, d = 0.0, q = 0.0;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long v = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < nnodes; v += stride) {
        double acc[G];
#pragma unroll
        for (int j = 0; j < G; ++j)
            acc[j] = 0.0;
        const long long e0 = adj_ptr[v], e1 = adj_ptr[v + 1];
        for (long long e = e0; e < e1; ++e) {
            const double *src = fe + (size_t)(adj_idx != nullptr ? adj_idx[e] : e) * FeStride<G>::v;
            if (G == 3) {
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
            const size_t i = v * G + j;
            w[i] = acc[j];
            if (minv[i] != 0.0) {
                const double rv = r[i], uv = u[i];
                g = fma(rv, uv, g);
                d = fma(acc[j], uv, d);
                q = fma(rv, rv, q);
            }
        }
    }
    const double bg = kr_block_sum(g, sh), bd = kr_block_sum(d, sh), bq = kr_block_sum(q, sh);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = bg;
        partials[gridDim.x + blockIdx.x] = bd;
        partials[2 * gridDim.x + blockIdx.x] = bq;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last)
        return;
    // The last CTA adds the CTA partials: thread t takes partials t, t + 256, ... (in that order), then the
    // fixed shuffle / shared-memory tree of kr_block_sum -- the same order every run and on every rank, so the
    // sums are reproducible bit for bit.  (Three threads walking all 8 x SMs partials one dependent L2 load
    // after the other cost 0.04 ms of the 0.466 ms iteration on one GPU, profiles/r2k_newton55_device_n1.json.)
    __threadfence();
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += KR_THREADS) {
        a0 += ld_volatile(partials + b);
        a1 += ld_volatile(partials + gridDim.x + b);
        a2 += ld_volatile(partials + 2 * (size_t)gridDim.x + b);
    }
    const double s0 = kr_block_sum(a0, sh), s1 = kr_block_sum(a1, sh), s2 = kr_block_sum(a2, sh);
    if (threadIdx.x == 0) {
        // all-gather of the local sums: one slot per (parity, source rank) in EVERY rank's block
        const int par = (int)(epoch & 1ULL);
        for (int t = 0; t < world; ++t) {
            double *slot = reinterpret_cast<double *>(peers.base[t] + lay.off_red) + ((size_t)par * world + rank) * 4;
            slot[0] = s0;
            slot[1] = s1;
            slot[2] = s2;
        }
        __threadfence_system();
        for (int t = 0; t < world; ++t) {
            unsigned long long *flag =
                reinterpret_cast<unsigned long long *>(peers.base[t] + lay.off_redflag) + (size_t)par * world + rank;
            st_flag(flag, epoch);
        }
        // ... and finish the reduction right here, in this ONE thread: wait for every rank's sums, add them
        // in rank order (every rank gets the same bits), alpha / beta for the update kernel that follows in
        // stream order (1184 CTAs polling a system-scope flag there cost more than the exchange itself)
        const unsigned long long *flags =
            reinterpret_cast<const unsigned long long *>(peers.base[rank] + lay.off_redflag) + (size_t)par * world;
        const double *red = reinterpret_cast<const double *>(peers.base[rank] + lay.off_red) + (size_t)par * world * 4;
        double gs = 0.0, ds = 0.0, qs = 0.0;
        for (int t = 0; t < world; ++t) {
            wait_flag(flags + t, epoch, true, err);
            gs += ld_volatile(red + t * 4 + 0);
            ds += ld_volatile(red + t * 4 + 1);
            qs += ld_volatile(red + t * 4 + 2);
        }
        const double *prev = state + (size_t)(par ^ 1) * 4;  // written by the previous iteration's launch
        double beta = 0.0, den = ds;
        if (!first) {
            beta = prev[0] > 0.0 ? gs / prev[0] : 0.0;
            den = prev[1] != 0.0 ? ds - beta * gs / prev[1] : ds;
        }
        // den <= 0: the operator is not positive definite on the free dofs (or the solve has converged to
        // round-off): stop moving and flag it; the host decides
        double alpha = den > 0.0 ? gs / den : 0.0;
        // residual test on the device: r.r at the start of this iteration against rtol^2 times that of iteration 0
        const double rr0 = first ? qs : ctl[3];
        const bool passed = rtol2 > 0.0 && qs <= rtol2 * rr0;
        if (passed)
            alpha = 0.0;  // x is the answer as it stands (K3 of this iteration is skipped anyway)
        if (first)
            ctl[3] = qs;
        ctl[2] = qs;
        ctl[5] = (double)it;
        if (!(den > 0.0))
            ctl[4] = 1.0;
        if (passed || !(den > 0.0)) {
            ctl[1] = (double)it;
            __threadfence();
            ctl[0] = 1.0;  // freeze: every later kernel of the loop returns at once
        }
        scal[0] = alpha;
        scal[1] = beta;
        double *cur = state + (size_t)par * 4;
        cur[0] = gs;
        cur[1] = alpha;
        cur[2] = qs;
        cur[3] = den > 0.0 ? (first ? 0.0 : prev[3]) : 1.0;  // sticky breakdown flag of this solve
        hist[it % KR_HIST] = qs;
        *ticket = 0;
    }
}

// K3: fused vector update with the alpha / beta K2 left in `scal`.  u is written on OWNED dofs only: a faster
// neighbour may already have stored this iteration's ghost values.  (Folding the ghost push into this kernel's
// tail -- the last CTA to finish stores all send entries -- was measured: 24 k nodes x 3 peer stores from ONE
// CTA cost 0.23 ms per iteration on two GPUs, 0.55 vs 0.32 ms; the push has its own full-grid kernel.)
__global__ void __launch_bounds__(KR_THREADS)
    cg_update_kernel(size_t n, size_t n_owned, double *__restrict__ x, double *__restrict__ r, double *__restrict__ u,
                     const double *__restrict__ w, double *__restrict__ p, double *__restrict__ s,
                     const double *__restrict__ minv, const double *__restrict__ scal, const double *__restrict__ ctl)
{
    if (ctl[0] != 0.0)
        return;  // frozen
    const double alpha = scal[0], beta = scal[1];
    const size_t n2 = n / 2;
    double2 *x2 = reinterpret_cast<double2 *>(x), *r2 = reinterpret_cast<double2 *>(r);
    double2 *u2 = reinterpret_cast<double2 *>(u), *p2 = reinterpret_cast<double2 *>(p);
    double2 *s2 = reinterpret_cast<double2 *>(s);
    const double2 *w2 = reinterpret_cast<const double2 *>(w), *m2 = reinterpret_cast<const double2 *>(minv);
    auto one = [&](double m, double uv, double wv, double &pv, double &sv, double &xv, double &rv, double &un) {
        const double al = m != 0.0 ? alpha : 0.0;
        pv = fma(beta, pv, uv);
        sv = fma(beta, sv, wv);
        xv = fma(al, pv, xv);
        rv = fma(-al, sv, rv);
        un = m * rv;
    };
    const size_t stride = (size_t)gridDim.x * KR_THREADS;
    for (size_t i = (size_t)blockIdx.x * KR_THREADS + threadIdx.x; i < n2; i += stride) {
        const double2 m = m2[i], uv = u2[i], wv = w2[i];
        double2 pv = p2[i], sv = s2[i], xv = x2[i], rv = r2[i], un;
        one(m.x, uv.x, wv.x, pv.x, sv.x, xv.x, rv.x, un.x);
        one(m.y, uv.y, wv.y, pv.y, sv.y, xv.y, rv.y, un.y);
        p2[i] = pv;
        s2[i] = sv;
        x2[i] = xv;
        r2[i] = rv;
        if (2 * i + 1 < n_owned) {
            u2[i] = un;
        } else if (2 * i < n_owned) {
            u[2 * i] = un.x;
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const size_t i = n - 1;
        double un;
        one(minv[i], u[i], w[i], p[i], s[i], x[i], r[i], un);
        if (i < n_owned)
            u[i] = un;
    }
}

// K4: owned values that are ghosts elsewhere -> the neighbours' u vectors (peer stores), then the
// epoch flag; finally wait for every neighbour's flag of the same epoch.
template <int G>
__global__ void __launch_bounds__(KR_THREADS)
    halo_push_kernel(const double *__restrict__ u, const int *__restrict__ send_src, const int *__restrict__ send_dst,
                     const int *__restrict__ send_nbr, int n_send, PeerPtrs nbr_base, CommLayout lay, int n_nbr,
                     int rank, char *comm, const int *__restrict__ nbr_rank_dev, unsigned *ticket,
                     unsigned long long epoch, int do_wait, int *err, const double *__restrict__ gate)
{
    __shared__ bool last;
    if (gate != nullptr && *gate != 0.0)
        return;  // push of a frozen iteration: every rank skips it (and the wait that goes with it)
    const int stride = gridDim.x * blockDim.x;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_send; e += stride) {
        double *dst = reinterpret_cast<double *>(nbr_base.base[send_nbr[e]] + lay.off_u) + (size_t)send_dst[e] * G;
        const double *src = u + (size_t)send_src[e] * G;
#pragma unroll
        for (int j = 0; j < G; ++j)
            dst[j] = src[j];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence_system();
        for (int k = 0; k < n_nbr; ++k)
            st_flag(reinterpret_cast<unsigned long long *>(nbr_base.base[k] + lay.off_haloflag) + rank, epoch);
        const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(comm + lay.off_haloflag);
        if (do_wait)
            for (int k = 0; k < n_nbr; ++k)
                wait_flag(mine + nbr_rank_dev[k], epoch, false, err);
        *ticket = 0;
    }
}

// The wait half of a push issued with do_wait = 0: one thread, until every neighbour's flag has reached `epoch`.
__global__ void halo_wait_kernel(const char *comm, CommLayout lay, int n_nbr, const int *__restrict__ nbr_rank_dev,
                                 unsigned long long epoch, int *err, const double *__restrict__ gate)
{
    if (gate != nullptr && *gate != 0.0)
        return;  // the push this wait belongs to was skipped on every rank (frozen solve)
    if (threadIdx.x == 0) {
        const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(comm + lay.off_haloflag);
        for (int k = 0; k < n_nbr; ++k)
            wait_flag(mine + nbr_rank_dev[k], epoch, false, err);
    }
}

// Ghost refresh of an arbitrary nodal vector through the comm block: owned part -> exchange vector
__global__ void __launch_bounds__(KR_THREADS)
    kr_copy_kernel(size_t lo, size_t hi, const double *__restrict__ src, double *__restrict__ dst)
{
    const size_t stride = (size_t)gridDim.x * KR_THREADS;
    for (size_t i = lo + (size_t)blockIdx.x * KR_THREADS + threadIdx.x; i < hi; i += stride)
        dst[i] = src[i];
}

static unsigned kr_grid(size_t work)
{
    size_t g = (work + KR_THREADS - 1) / KR_THREADS;
    const size_t cap = (size_t)sm_count() * 8;
    return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

// gated = true: a push inside the iteration loop (skipped once the solve is frozen); false: the begin / halo-update
// pushes, which always run.
static int kr_push(Krylov *K, cudaStream_t st, int do_wait = 1, bool gated = false)
{
    if (K->world == 1 || K->n_nbr == 0)
        return FCX_OK;
    K->push_epoch = K->epoch;
    K->push_waited = do_wait != 0;
    K->push_gated = gated;
    PeerPtrs nb{};
    for (int k = 0; k < K->n_nbr; ++k)
        nb.base[k] = K->peers.base[K->nbr_rank[k]];
    const unsigned grid = kr_grid(K->n_send > 0 ? (size_t)K->n_send : 1);
    double *u = reinterpret_cast<double *>(K->comm + K->lay.off_u);
    // nbr ranks on the device: kept right behind the send plan
    const int *nbr_rank_dev = K->send_nbr + K->n_send;
#define FCX_PUSH(G) \
    halo_push_kernel<G><<<grid, KR_THREADS, 0, st>>>(u, K->send_src, K->send_dst, K->send_nbr, K->n_send, nb, K->lay, \
                                                     K->n_nbr, K->rank, K->comm, nbr_rank_dev, K->ticket + 1, K->epoch, \
                                                     do_wait, K->err, gated ? K->ctl : nullptr)
    if (K->gdim == 1)
        FCX_PUSH(1);
    else if (K->gdim == 2)
        FCX_PUSH(2);
    else
        FCX_PUSH(3);
#undef FCX_PUSH
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "halo_push_kernel launch");
}

// Wait for the neighbours' pushes of the last epoch this rank pushed without waiting.
static int kr_wait(Krylov *K, cudaStream_t st)
{
    if (K->world == 1 || K->n_nbr == 0 || K->push_waited)
        return FCX_OK;
    // gated by the control block when the owed push was an iteration's: a frozen iteration neither pushed nor waits
    halo_wait_kernel<<<1, 32, 0, st>>>(K->comm, K->lay, K->n_nbr, K->send_nbr + K->n_send, K->push_epoch, K->err,
                                       K->push_gated ? K->ctl : nullptr);
    K->push_waited = true;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "halo_wait_kernel launch");
}

}  // namespace fcx

using namespace fcx;

extern "C" {

int fcx_krylov_create(int rank, int world, int gdim, size_t nnodes, size_t nnodes_owned, void **handle_out,
                      void **comm_out, unsigned char *ipc_handle_out /* 64 bytes */)
{
    if (!handle_out || !comm_out || !ipc_handle_out)
        return FCX_ERR_NULL;
    if (world < 1 || world > KR_MAX_WORLD || rank < 0 || rank >= world || gdim < 1 || gdim > 3 || nnodes == 0 ||
        nnodes_owned > nnodes)
        return FCX_ERR_ARG;
    Krylov *K = new Krylov();
    K->rank = rank;
    K->world = world;
    K->gdim = gdim;
    K->nnodes = nnodes;
    K->n = nnodes * (size_t)gdim;
    K->n_owned = nnodes_owned * (size_t)gdim;
    K->lay = CommLayout(K->n, world);
    K->grid = kr_grid(K->n / 2 + 1);
    cudaError_t e = cudaMalloc((void **)&K->comm, K->lay.bytes);
    if (e == cudaSuccess)
        e = cudaMemset(K->comm, 0, K->lay.bytes);
    double **vecs[6] = {&K->x, &K->r, &K->w, &K->p, &K->s, &K->minv};
    for (int k = 0; k < 6 && e == cudaSuccess; ++k)
        e = cudaMalloc((void **)vecs[k], sizeof(double) * K->n);
    const unsigned gmax = (unsigned)sm_count() * 8;
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&K->partials, sizeof(double) * 3 * gmax);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&K->ticket, sizeof(unsigned) * 2);
    if (e == cudaSuccess)
        e = cudaMemset(K->ticket, 0, sizeof(unsigned) * 2);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&K->state, sizeof(double) * 8);
    if (e == cudaSuccess)
        e = cudaMemset(K->state, 0, sizeof(double) * 8);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&K->scal, sizeof(double) * 2);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&K->hist, sizeof(double) * KR_HIST);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&K->err, sizeof(int));
    if (e == cudaSuccess)
        e = cudaMemset(K->err, 0, sizeof(int));
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&K->ctl, sizeof(double) * KR_CTL);
    if (e == cudaSuccess)
        e = cudaMemset(K->ctl, 0, sizeof(double) * KR_CTL);
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&K->tile_tickets, sizeof(unsigned long long) * 2);
    if (e == cudaSuccess)
        e = cudaMemset(K->tile_tickets, 0, sizeof(unsigned long long) * 2);
    if (e == cudaSuccess)
        e = cudaMallocHost((void **)&K->snap_host, sizeof(double) * 2 * (KR_CTL + 1));
    for (int k = 0; k < 2 && e == cudaSuccess; ++k)
        e = cudaEventCreateWithFlags(&K->snap_ev[k], cudaEventDisableTiming);
    if (e == cudaSuccess) {
        cudaIpcMemHandle_t h;
        memset(&h, 0, sizeof h);
        if (world > 1)
            e = cudaIpcGetMemHandle(&h, K->comm);
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        memcpy(ipc_handle_out, &h, 64);
    }
    if (e != cudaSuccess) {
        delete K;
        return note_cuda_error(e, "fcx_krylov_create");
    }
    K->peers.base[rank] = K->comm;
    *handle_out = K;
    *comm_out = K->comm;
    return FCX_OK;
}

/* ipc_handles: world x 64 bytes (rank order, own entry ignored).  Opens the peers' comm blocks. */
int fcx_krylov_connect(void *handle, const unsigned char *ipc_handles)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K || (K->world > 1 && !ipc_handles))
        return FCX_ERR_NULL;
    for (int t = 0; t < K->world; ++t) {
        if (t == K->rank)
            continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, ipc_handles + (size_t)t * 64, 64);
        void *ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess)
            return note_cuda_error(e, "cudaIpcOpenMemHandle");
        K->peers.base[t] = static_cast<char *>(ptr);
        K->peer_open[t] = true;
    }
    return FCX_OK;
}

/* Halo plan: for neighbour slot k (rank nbr_rank[k]) the entries [send_ptr[k], send_ptr[k+1]) of
 * send_src (my local node) / send_dst (the neighbour's local node of the same mesh node).  HOST arrays. */
int fcx_krylov_set_halo(void *handle, int n_nbr, const int *nbr_rank, const int *send_ptr, const int *send_src,
                        const int *send_dst)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K)
        return FCX_ERR_NULL;
    if (n_nbr < 0 || n_nbr >= KR_MAX_WORLD)
        return FCX_ERR_ARG;
    K->n_nbr = n_nbr;
    if (n_nbr == 0)
        return FCX_OK;
    if (!nbr_rank || !send_ptr || !send_src || !send_dst)
        return FCX_ERR_NULL;
    const int total = send_ptr[n_nbr];
    std::vector<int> nbr((size_t)total + n_nbr);
    for (int k = 0; k < n_nbr; ++k) {
        K->nbr_rank[k] = nbr_rank[k];
        for (int e = send_ptr[k]; e < send_ptr[k + 1]; ++e)
            nbr[e] = k;
        nbr[(size_t)total + k] = nbr_rank[k];
    }
    K->n_send = total;
    cudaError_t e = cudaMalloc((void **)&K->send_src, sizeof(int) * (size_t)(total > 0 ? total : 1));
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&K->send_dst, sizeof(int) * (size_t)(total > 0 ? total : 1));
    if (e == cudaSuccess)
        e = cudaMalloc((void **)&K->send_nbr, sizeof(int) * ((size_t)total + n_nbr));
    if (e == cudaSuccess && total > 0)
        e = cudaMemcpy(K->send_src, send_src, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && total > 0)
        e = cudaMemcpy(K->send_dst, send_dst, sizeof(int) * (size_t)total, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cudaMemcpy(K->send_nbr, nbr.data(), sizeof(int) * nbr.size(), cudaMemcpyHostToDevice);
    return note_cuda_error(e, "fcx_krylov_set_halo");
}

/* The Jacobian action: the arguments of fcx_tangent_apply_rec (mode 3, `tangent` = the 10-double
 * records) or fcx_tangent_apply (mode 1, dense tangents) + the node adjacency of fcx_gather_sum. */
int fcx_krylov_set_operator(void *handle, int mode, int sdim, size_t ncells, int nq, int nd, const int *dofmap,
                            const double *dphi_ref, const double *weights, const double *Jinv, const double *detJ,
                            const double *tangent, double *fe, const int *fe_pos, const long long *adj_ptr,
                            const int *adj_idx, size_t ncells_interior)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K || !dofmap || !dphi_ref || !weights || !Jinv || !detJ || !tangent || !fe || !adj_ptr)
        return FCX_ERR_NULL;
    if (mode != 1 && mode != 3)
        return FCX_ERR_ARG;
    K->op_mode = mode;
    K->sdim = sdim;
    K->ncells = ncells;
    K->nq = nq;
    K->nd = nd;
    K->dofmap = dofmap;
    K->dphi = dphi_ref;
    K->weights = weights;
    K->Jinv = Jinv;
    K->detJ = detJ;
    K->tang = tangent;
    K->fe = fe;
    K->fe_pos = fe_pos;
    K->adj_ptr = adj_ptr;
    K->adj_idx = fe_pos != nullptr ? nullptr : adj_idx;
    K->ncells_interior = ncells_interior <= ncells ? ncells_interior : ncells;  // ncells = no split
    return FCX_OK;
}

/* Start a solve: x = 0, r = rhs on the dofs with minv != 0 (free AND owned), u = minv r; pushes the
 * ghost values of u.  rhs, minv: DEVICE vectors of n doubles. */
int fcx_krylov_begin(void *handle, const double *rhs, const double *minv, void *stream)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K || !rhs || !minv)
        return FCX_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *u = reinterpret_cast<double *>(K->comm + K->lay.off_u);
    if (int rc = kr_wait(K, st))  // a previous solve's last push is still owed its wait
        return rc;
    kr_begin_kernel<<<kr_grid(K->n), KR_THREADS, 0, st>>>(K->n, K->n_owned, rhs, minv, K->minv, K->x, K->r, u, K->p, K->s,
                                                          K->ctl);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return note_cuda_error(e, "kr_begin_kernel launch");
    K->it = 0;
    K->epoch += 1;
    return kr_push(K, st, 0);  // the first iteration waits, behind its interior cells
}

/* Enqueue `iters` iterations (K1..K4 each) on `stream`; never synchronises. */
int fcx_krylov_iterate(void *handle, int iters, void *stream)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K)
        return FCX_ERR_NULL;
    if (K->op_mode == 0)
        return FCX_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *u = reinterpret_cast<double *>(K->comm + K->lay.off_u);
    const unsigned ggrid = kr_grid(K->nnodes);
    // element kernel over the local cells [c0, c0 + nc)
    auto element_kernel = [&](size_t c0, size_t nc) -> int {
        if (nc == 0)
            return FCX_OK;
        const int fs = K->gdim == 3 ? 4 : K->gdim;
        const int rl = K->op_mode == 3 ? 10 : K->sdim * K->sdim;  // tangent doubles per quadrature point
        const int *pos = K->fe_pos ? K->fe_pos + c0 * K->nd : nullptr;
        double *fe = K->fe_pos ? K->fe : K->fe + c0 * K->nd * fs;
        return K->op_mode == 3
                   ? fcx_tangent_apply_rec(K->gdim, K->sdim, nc, K->nq, K->nd, K->dofmap + c0 * K->nd, u, K->dphi,
                                           K->weights, K->Jinv + c0 * K->gdim * K->gdim, K->detJ + c0,
                                           K->tang + c0 * K->nq * rl, fe, pos, stream)
                   : fcx_tangent_apply(K->gdim, K->sdim, nc, K->nq, K->nd, K->dofmap + c0 * K->nd, u, K->dphi,
                                       K->weights, K->Jinv + c0 * K->gdim * K->gdim, K->detJ + c0,
                                       K->tang + c0 * K->nq * rl, fe, pos, stream);
    };
    struct GateGuard {  // the element kernels of this loop return at once when the solve is frozen, and hand
                        // out their tiles from an alternating pair of counters (no memset per launch)
        GateGuard(const double *g, unsigned long long *t, unsigned *n)
        {
            fem_set_launch_gate(g);
            fem_set_launch_tickets(t, n);
        }
        ~GateGuard()
        {
            fem_set_launch_gate(nullptr);
            fem_set_launch_tickets(nullptr, nullptr);
        }
    } gate_guard(K->ctl, K->tile_tickets, &K->tile_launches);
    for (int k = 0; k < iters; ++k) {
        // K1: cells that touch no ghost node first -- the neighbours' ghost stores of the previous update arrive
        // behind them; then the wait (one thread), then the boundary cells
        const size_t n_int = (K->world > 1 && K->n_nbr > 0) ? K->ncells_interior : K->ncells;
        int rc = element_kernel(0, n_int);
        if (rc == FCX_OK)
            rc = kr_wait(K, st);
        if (rc == FCX_OK)
            rc = element_kernel(n_int, K->ncells - n_int);
        if (rc != FCX_OK)
            return rc;
        K->epoch += 1;
        const int first = K->it == 0 ? 1 : 0;
#define FCX_GSUM(G) \
    gsum_dots_kernel<G><<<ggrid, KR_THREADS, 0, st>>>(K->adj_ptr, K->adj_idx, K->fe, K->r, u, K->minv, K->w, K->nnodes, \
                                                      K->partials, K->ticket, K->peers, K->lay, K->rank, K->world, K->epoch, \
                                                      first, K->state, K->scal, K->hist, K->it, K->err, K->ctl, K->rtol2)
        if (K->gdim == 1)
            FCX_GSUM(1);
        else if (K->gdim == 2)
            FCX_GSUM(2);
        else
            FCX_GSUM(3);
#undef FCX_GSUM
        cg_update_kernel<<<K->grid, KR_THREADS, 0, st>>>(K->n, K->n_owned, K->x, K->r, u, K->w, K->p, K->s, K->minv, K->scal,
                                                         K->ctl);
        g_launches.fetch_add(2, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return note_cuda_error(e, "krylov iteration launch");
        K->it += 1;
        rc = kr_push(K, st, 0, true);  // push and go on: the wait sits behind the next iteration's interior cells
        if (rc != FCX_OK)
            return rc;
    }
    return FCX_OK;
}

/* Overwrite the ghost entries of the nodal vector x (DEVICE, n doubles, owned nodes first) with their owners'
 * values -- the same peer-memory push as inside the iteration, with the exchange vector of the (finished) solve
 * as the staging area: what the reference gets from PETSc's ghostUpdate / dolfinx's scatter_forward
 * (solver/_incrementalunknowns.py:36-38) after the Newton update.  Collective over the ranks; enqueue-only. */
int fcx_krylov_halo_update(void *handle, double *x, void *stream)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K || !x)
        return FCX_ERR_NULL;
    if (K->world == 1 || K->n_nbr == 0)
        return FCX_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *u = reinterpret_cast<double *>(K->comm + K->lay.off_u);
    int rc = kr_wait(K, st);  // the last iteration's push is still owed its wait
    if (rc != FCX_OK)
        return rc;
    kr_copy_kernel<<<kr_grid(K->n_owned), KR_THREADS, 0, st>>>(0, K->n_owned, x, u);
    K->epoch += 1;
    rc = kr_push(K, st, 1);
    if (rc != FCX_OK)
        return rc;
    kr_copy_kernel<<<kr_grid(K->n - K->n_owned + 1), KR_THREADS, 0, st>>>(K->n_owned, K->n, u, x);
    g_launches.fetch_add(2, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "fcx_krylov_halo_update");
}

/* After a synchronisation of the stream: out[0] = iterations done, out[1] = r.r at the start of the last
 * iteration (= after iterations-1 updates), out[2] = r.r of iteration 0 (the right-hand side), out[3] = 1 if a
 * breakdown (p.Ap <= 0) was seen.  Blocking copies. */
int fcx_krylov_status(void *handle, double *out4)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K || !out4)
        return FCX_ERR_NULL;
    out4[0] = (double)K->it;
    out4[1] = out4[2] = out4[3] = 0.0;
    if (K->it == 0)
        return FCX_OK;
    cudaError_t e = cudaMemcpy(out4 + 1, K->hist + (K->it - 1) % KR_HIST, sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
        e = cudaMemcpy(out4 + 2, K->hist, sizeof(double), cudaMemcpyDeviceToHost);
    double st[4];
    const int par = (int)(K->epoch & 1ULL);
    if (e == cudaSuccess)
        e = cudaMemcpy(st, K->state + (size_t)par * 4, sizeof st, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
        out4[3] = st[3];
    int err = 0;
    if (e == cudaSuccess)
        e = cudaMemcpy(&err, K->err, sizeof err, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && err != 0)
        out4[3] = 2.0;  // a peer never arrived
    return note_cuda_error(e, "fcx_krylov_status");
}

/* Residual test ON THE DEVICE: the solve freezes itself once r.r <= rtol^2 * (r.r of iteration 0) -- the
 * iteration that sees it leaves x untouched and every later kernel of the loop returns at once (also after a
 * breakdown, p.Ap <= 0).  The host may therefore enqueue blocks of iterations AHEAD of knowing the outcome
 * (fcx_krylov_snapshot / fcx_krylov_wait_snapshot) instead of draining the stream at every check: the answer and
 * the iteration count are those of the exact stopping iteration, not of the end of a block.  rtol <= 0: never
 * freeze (the host-driven stop rule of fcx_krylov_status).  Takes effect with the next fcx_krylov_begin. */
int fcx_krylov_set_tolerance(void *handle, double rtol)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K)
        return FCX_ERR_NULL;
    K->rtol2 = rtol > 0.0 ? rtol * rtol : 0.0;
    return FCX_OK;
}

/* Enqueue a snapshot of the control block into pinned host slot `slot` (0 / 1) on `stream`, with an event behind. */
int fcx_krylov_snapshot(void *handle, int slot, void *stream)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K)
        return FCX_ERR_NULL;
    if (slot < 0 || slot > 1)
        return FCX_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    double *dst = K->snap_host + (size_t)slot * (KR_CTL + 1);
    cudaError_t e = cudaMemcpyAsync(dst, K->ctl, sizeof(double) * KR_CTL, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)  // the sticky error word (an int) travels in the last double's storage
        e = cudaMemcpyAsync(dst + KR_CTL, K->err, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess)
        e = cudaEventRecord(K->snap_ev[slot], st);
    return note_cuda_error(e, "fcx_krylov_snapshot");
}

/* Wait for snapshot `slot`; out6 = {frozen, iteration at which it froze, r.r at the start of the latest live
 * iteration, r.r of iteration 0, breakdown seen (2 = a peer never arrived), live iterations so far}. */
int fcx_krylov_wait_snapshot(void *handle, int slot, double *out6)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K || !out6)
        return FCX_ERR_NULL;
    if (slot < 0 || slot > 1)
        return FCX_ERR_ARG;
    cudaError_t e = cudaEventSynchronize(K->snap_ev[slot]);
    if (e != cudaSuccess)
        return note_cuda_error(e, "fcx_krylov_wait_snapshot");
    const double *src = K->snap_host + (size_t)slot * (KR_CTL + 1);
    for (int k = 0; k < 6; ++k)
        out6[k] = src[k];
    int err = 0;
    memcpy(&err, src + KR_CTL, sizeof err);
    if (err != 0)
        out6[4] = 2.0;
    return FCX_OK;
}

/* x (DEVICE, n doubles) <- the current iterate, on `stream`. */
int fcx_krylov_solution(void *handle, double *x_out, void *stream)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K || !x_out)
        return FCX_ERR_NULL;
    return note_cuda_error(cudaMemcpyAsync(x_out, K->x, sizeof(double) * K->n, cudaMemcpyDeviceToDevice,
                                           static_cast<cudaStream_t>(stream)),
                           "fcx_krylov_solution");
}

void fcx_krylov_destroy(void *handle)
{
    Krylov *K = static_cast<Krylov *>(handle);
    if (!K)
        return;
    cudaDeviceSynchronize();
    for (int t = 0; t < K->world; ++t)
        if (K->peer_open[t])
            cudaIpcCloseMemHandle(K->peers.base[t]);
    void *ptrs[] = {K->comm, K->x, K->r, K->w, K->p, K->s, K->minv, K->partials, K->ticket, K->state, K->scal, K->hist, K->err,
                    K->ctl, K->tile_tickets, K->send_src, K->send_dst, K->send_nbr};
    for (void *q : ptrs)
        if (q)
            cudaFree(q);
    if (K->snap_host)
        cudaFreeHost(K->snap_host);
    for (cudaEvent_t ev : K->snap_ev)
        if (ev)
            cudaEventDestroy(ev);
    delete K;
}

}  // extern "C"
