// fcx_tile.cuh -- the persistent, double-buffered tile pipeline shared by every
// constitutive kernel with Mandel dim s in {4, 6}.
//
// Why a tile pipeline: the drop-in contract is array-of-structs (reference
// models/interfaces.py:82-101) -- a thread that owns one QP sees
// strides of 72/48/288 bytes, which would waste most of every 128-byte line.
// But a TILE of consecutive QPs is one contiguous byte range in every array.
// So each CTA
//   1. pulls the tile's ranges global->shared with 1-D bulk async copies (TMA
//      engine, completion on an mbarrier), double-buffered so the next tile is
//      in flight while the current one is computed;
//   2. lets thread t compute QP t entirely in registers, reading/writing its
//      slots of the shared stage (in place);
//   3. pushes the in-place segments (stress, history) back shared->global with
//      bulk async stores, and
//   4. writes the s*s tangent block of the tile -- >50 % of all traffic --
//      warp-cooperatively as a dense, fully coalesced stream of 16-byte stores
//      regenerated from a few per-QP scalars kept in shared memory
//      (Model::store_tangent), so the 288 B/QP never sit in shared memory.
// Grid = min(#tiles, SMs x resident CTAs); tiles are strided over CTAs.
//
// Tail tiles (n % TILE != 0) and pointers that are not 16-byte aligned take a
// generic-proxy path with plain coalesced loads/stores through the same stage.
#pragma once
#include <type_traits>

#include "fcx_ptx.cuh"

namespace fcx {

template <int N>
struct SegPtrs {
    double *p[N];
};

// Models whose plastic branch is an expensive loop (the Drucker-Prager return mapping) may opt into a
// TWO-PHASE update: Model::trial(...) classifies every point of the tile and finishes the elastic ones;
// the points that need the return mapping are appended to a CTA-wide list in shared memory and
// Model::qp(...) then runs on the list with consecutive threads -- whole warps iterate or sit out,
// instead of every warp iterating with half of its lanes idle at ~50 % plastic points.
template <class M, class = void>
struct TwoPhase : std::false_type {};
template <class M>
struct TwoPhase<M, std::void_t<decltype(M::two_phase())>> : std::bool_constant<M::two_phase()> {};

// Accessor for one QP's slot of segment K inside a shared stage.
template <class M, int TILE>
struct QpView {
    double *stage;
    int t;
    template <int K>
    __device__ __forceinline__ double ld(int i) const
    {
        if (M::soa(K))
            return stage[M::off(K) * TILE + i * TILE + t];
        return stage[M::off(K) * TILE + t * M::w(K) + i];
    }
    template <int K>
    __device__ __forceinline__ void st(int i, double v) const
    {
        if (M::soa(K))
            stage[M::off(K) * TILE + i * TILE + t] = v;
        else
            stage[M::off(K) * TILE + t * M::w(K) + i] = v;
    }
};

// WT = false: the stress-only instantiation (tangent == nullptr, the reference's
// `tangent: Option<..>`): no tangent record / constant block in shared memory, nothing of the
// tangent computed or stored -- the smaller footprint buys one more resident CTA per SM.
template <class M, int TILE, bool WT = true>
constexpr size_t tile_smem_bytes()
{
    return sizeof(double) * (2 * M::wsum() * TILE + (WT ? M::aux_doubles(TILE) : 0)) + 2 * sizeof(uint64_t);
}

template <class M, int TILE, bool WT>
constexpr int tile_min_ctas()
{
    if (WT || M::min_ctas(TILE) * TILE < 512)  // register-hungry models (Drucker-Prager) keep their budget
        return M::min_ctas(TILE);
    const int by_smem = (int)((227 * 1024) / (tile_smem_bytes<M, TILE, false>() + 1024));
    const int want = M::min_ctas(TILE) + 1;
    return want < by_smem ? want : (by_smem > M::min_ctas(TILE) ? by_smem : M::min_ctas(TILE));
}

template <class M, int TILE, bool WT = true>
__global__ void __launch_bounds__(TILE, tile_min_ctas<M, TILE, WT>())
    fcx_tile_kernel(const __grid_constant__ typename M::Params prm,
                    const __grid_constant__ SegPtrs<M::nseg()> io, double *__restrict__ tangent,
                    const unsigned long long n, const int flags,
                    unsigned char *__restrict__ flag, int *__restrict__ status,
                    const unsigned long long qbase, unsigned long long *__restrict__ ticket)
{
    // ticket: nullptr = tiles strided statically over the grid; otherwise an
    //        atomic counter (zero at launch) hands out tiles gridDim.x, gridDim.x+1, ...
    //        in request order, which keeps the tiles in flight a tight moving
    //        window in every array (DRAM page locality: +10 % on B200)
    // qbase: index of this launch's first QP in the caller's arrays (only used
    //        for the first-failing-point report in status[1])
    // flags: bit0 = bulk (TMA) path allowed (all pointers 16-byte aligned),
    //        bit1 = L2 evict_first hint on the bulk loads, bit2 = on the bulk stores
    const bool bulk_ok = (flags & 1) != 0;
    const bool hint_ld = (flags & 2) != 0;
    const bool hint_st = (flags & 4) != 0;
    // bit3: constant-tangent models stream the tile's tangent block with bulk
    //       stores from a constant shared-memory block instead of thread stores
    // tangent == nullptr: stress-only evaluate (the reference's `tangent: Option<..>`,
    //       comfe-rs/src/interfaces.rs:368): no tangent traffic at all
    const bool want_tan = WT && tangent != nullptr;
    const bool ct_bulk = (flags & 8) != 0 && M::const_tangent_qps() > 0 && want_tan;
    const uint64_t pol = policy_evict_first();
    constexpr int NSEG = M::nseg();
    constexpr int WSUM = M::wsum();
    constexpr int SS = M::sdim() * M::sdim();
    extern __shared__ __align__(128) double smem[];
    double *aux = WT ? smem + 2 * WSUM * TILE : nullptr;  // nullptr: Model::qp leaves no tangent record
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 2 * WSUM * TILE + (WT ? M::aux_doubles(TILE) : 0));

    const int tid = threadIdx.x;
    const unsigned long long ntiles = (n + TILE - 1) / TILE;

    __shared__ unsigned long long s_ticket;
    if (WT)
        M::init_aux(prm, aux, tid, TILE);
    if (tid == 0)
        s_ticket = (ticket != nullptr) ? gridDim.x + atomicAdd(ticket, 1ULL)
                                       : (unsigned long long)blockIdx.x + gridDim.x;
    if (WT && M::const_tangent_qps() > 0)
        fence_proxy_async_smem();  // aux is a bulk-store source (constant tangent block)
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // thread 0 only: start the bulk loads of a full tile into stage s
    auto issue_load = [&](unsigned long long tile, int s) {
        const unsigned long long q0 = tile * TILE;
        mbar_arrive_expect_tx(&bars[s], (uint32_t)(WSUM * TILE * sizeof(double)));
        double *dst = smem + s * WSUM * TILE;
#pragma unroll
        for (int k = 0; k < NSEG; ++k) {
            if (M::soa(k)) {
#pragma unroll
                for (int i = 0; i < M::w(k); ++i)
                    bulk_g2s(dst + (M::off(k) + i) * TILE, io.p[k] + (size_t)i * n + q0,
                             TILE * sizeof(double), &bars[s]);
            } else if (hint_ld) {
                bulk_g2s_hint(dst + M::off(k) * TILE, io.p[k] + q0 * M::w(k),
                              TILE * M::w(k) * sizeof(double), &bars[s], pol);
            } else {
                bulk_g2s(dst + M::off(k) * TILE, io.p[k] + q0 * M::w(k),
                         TILE * M::w(k) * sizeof(double), &bars[s]);
            }
        }
    };
    auto is_bulk = [&](unsigned long long tile) {
        return bulk_ok && (tile + 1) * TILE <= n;
    };

    unsigned long long tile = blockIdx.x;
    unsigned long long next = s_ticket;  // published before the barrier above
    if (tid == 0 && tile < ntiles && is_bulk(tile))
        issue_load(tile, 0);
    __syncthreads();  // everyone holds `next` before thread 0 overwrites s_ticket

    uint32_t parity0 = 0, parity1 = 0;
    for (int it = 0; tile < ntiles; ++it) {
        const int s = it & 1;
        const unsigned long long q0 = tile * TILE;
        const int cnt = (n - q0 < (unsigned long long)TILE) ? (int)(n - q0) : TILE;
        double *stage = smem + s * WSUM * TILE;
        const bool bulk = is_bulk(tile);

        // prefetch the next tile of this CTA into the other stage (free since
        // the closing barrier of the previous iteration)
        if (tid == 0) {
            if (next < ntiles && is_bulk(next))
                issue_load(next, s ^ 1);
            // tile after `next`; read by everyone after the mid-iteration barrier
            s_ticket = (ticket != nullptr) ? gridDim.x + atomicAdd(ticket, 1ULL) : next + gridDim.x;
        }

        if (bulk) {
            if (s == 0) {
                mbar_wait(&bars[0], parity0);
                parity0 ^= 1;
            } else {
                mbar_wait(&bars[1], parity1);
                parity1 ^= 1;
            }
        } else {
#pragma unroll
            for (int k = 0; k < NSEG; ++k) {
                if (M::soa(k)) {
                    for (int i = 0; i < M::w(k); ++i)
                        for (int j = tid; j < cnt; j += TILE)
                            stage[(M::off(k) + i) * TILE + j] = io.p[k][(size_t)i * n + q0 + j];
                } else {
                    const double *src = io.p[k] + q0 * M::w(k);
                    for (int j = tid; j < cnt * M::w(k); j += TILE)
                        stage[M::off(k) * TILE + j] = src[j];
                }
            }
            __syncthreads();
        }

        // ---- per-QP update, registers only; results go back in place ----
        auto report_failure = [&](int j) {
            if (status != nullptr) {
                atomicAdd(&status[0], 1);
                const unsigned long long q = qbase + q0 + j;
                atomicMin(&status[1], q > 0x7fffffffULL ? 0x7fffffff : (int)q);
            }
        };
        if constexpr (TwoPhase<M>::value) {
            __shared__ int s_np;
            __shared__ unsigned short s_list[TILE];
            if (tid == 0)
                s_np = 0;
            __syncthreads();
            bool plastic = false, failed = false, need = false;
            if (tid < cnt) {
                QpView<M, TILE> v{stage, tid};
                need = M::trial(prm, v, aux, tid, plastic, failed);
                if (flag != nullptr)
                    flag[q0 + tid] = plastic ? 1 : 0;
                if (failed)
                    report_failure(tid);
            }
            const unsigned m = __ballot_sync(0xffffffffu, need);
            if (m != 0) {
                const int lane = tid & 31, leader = __ffs(m) - 1;
                int base = 0;
                if (lane == leader)
                    base = atomicAdd(&s_np, __popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (need)
                    s_list[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)tid;
            }
            __syncthreads();
            const int np = s_np;
            for (int k = tid; k < np; k += TILE) {
                const int j = s_list[k];
                bool pl2 = false, f2 = false;
                QpView<M, TILE> v{stage, j};
                M::qp(prm, v, aux, j, pl2, f2);
                if (f2)
                    report_failure(j);
            }
        } else if (tid < cnt) {
            bool plastic = false, failed = false;
            QpView<M, TILE> v{stage, tid};
            M::qp(prm, v, aux, tid, plastic, failed);
            if (M::has_flag() && flag != nullptr)
                flag[q0 + tid] = plastic ? 1 : 0;
            if (M::has_flag() && failed)
                report_failure(tid);
        }
        if (bulk)
            fence_proxy_async_smem();
        __syncthreads();
        const unsigned long long after = s_ticket;

        // ---- write back the in-place segments ----
        if (bulk) {
            if (tid == 0) {
#pragma unroll
                for (int k = 0; k < NSEG; ++k) {
                    if (!M::wr(k))
                        continue;
                    if (M::soa(k)) {
#pragma unroll
                        for (int i = 0; i < M::w(k); ++i)
                            bulk_s2g(io.p[k] + (size_t)i * n + q0, stage + (M::off(k) + i) * TILE,
                                     TILE * sizeof(double));
                    } else if (hint_st) {
                        bulk_s2g_hint(io.p[k] + q0 * M::w(k), stage + M::off(k) * TILE,
                                      TILE * M::w(k) * sizeof(double), pol);
                    } else {
                        bulk_s2g(io.p[k] + q0 * M::w(k), stage + M::off(k) * TILE,
                                 TILE * M::w(k) * sizeof(double));
                    }
                }
                bulk_commit();
            }
        } else {
#pragma unroll
            for (int k = 0; k < NSEG; ++k) {
                if (!M::wr(k))
                    continue;
                if (M::soa(k)) {
                    for (int i = 0; i < M::w(k); ++i)
                        for (int j = tid; j < cnt; j += TILE)
                            io.p[k][(size_t)i * n + q0 + j] = stage[(M::off(k) + i) * TILE + j];
                } else {
                    double *dst = io.p[k] + q0 * M::w(k);
                    for (int j = tid; j < cnt * M::w(k); j += TILE)
                        dst[j] = stage[M::off(k) * TILE + j];
                }
            }
        }

        // ---- tangent block of the tile: dense coalesced stream ----
        if (!want_tan) {
            // stress-only call: nothing to write
        } else if (bulk && ct_bulk) {
            // every QP has the same s*s matrix: aux holds it repeated CQ times;
            // TILE/CQ bulk stores from that never-modified block cover the tile
            constexpr int CQ = M::const_tangent_qps() > 0 ? M::const_tangent_qps() : 1;
            if (tid == 0) {
#pragma unroll
                for (int b = 0; b < TILE / CQ; ++b)
                    bulk_s2g(tangent + (q0 + (unsigned long long)b * CQ) * SS, aux,
                             CQ * SS * sizeof(double));
                bulk_commit();
            }
        } else {
            M::store_tangent(prm, aux, tangent + q0 * SS, cnt, tid, TILE, bulk_ok);
        }

        if (bulk && tid == 0) {
            // stage may be refilled after the barrier; the tangent group of a
            // constant-tangent model reads only the constant block and may lag
            if (ct_bulk)
                bulk_wait_read_1();
            else
                bulk_wait_read_all();
        }
        __syncthreads();
        tile = next;
        next = after;
    }
    if (tid == 0)
        bulk_wait_read_all();  // shared memory must outlive every pending bulk store
}

// ---------------------------------------------------------------------------
// Elementwise kernel for the uniaxial constraints (s = g = 1): every array is
// [n] doubles, so plain vectorised grid-stride access is already coalesced.
// Each thread owns two consecutive QPs (16-byte loads/stores).
template <class M>
__global__ void __launch_bounds__(256)
    fcx_uniaxial_kernel(const __grid_constant__ typename M::Params prm,
                        const __grid_constant__ SegPtrs<M::nseg()> io,
                        double *__restrict__ tangent, const unsigned long long n, const int vec_ok)
{
    constexpr int NSEG = M::nseg();
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long gtid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec_ok) {
        const unsigned long long npair = n / 2;
        for (unsigned long long p = gtid; p < npair; p += stride) {
            double a[NSEG], b[NSEG], ta, tb;
#pragma unroll
            for (int k = 0; k < NSEG; ++k) {
                const double2 v = reinterpret_cast<const double2 *>(io.p[k])[p];
                a[k] = v.x;
                b[k] = v.y;
            }
            M::qp1(prm, a, ta);
            M::qp1(prm, b, tb);
#pragma unroll
            for (int k = 0; k < NSEG; ++k)
                if (M::wr(k))
                    reinterpret_cast<double2 *>(io.p[k])[p] = make_double2(a[k], b[k]);
            if (tangent != nullptr)
                reinterpret_cast<double2 *>(tangent)[p] = make_double2(ta, tb);
        }
        if ((n & 1ULL) && gtid == 0) {
            double a[NSEG], ta;
#pragma unroll
            for (int k = 0; k < NSEG; ++k)
                a[k] = io.p[k][n - 1];
            M::qp1(prm, a, ta);
#pragma unroll
            for (int k = 0; k < NSEG; ++k)
                if (M::wr(k))
                    io.p[k][n - 1] = a[k];
            if (tangent != nullptr)
                tangent[n - 1] = ta;
        }
    } else {
        for (unsigned long long q = gtid; q < n; q += stride) {
            double a[NSEG], ta;
#pragma unroll
            for (int k = 0; k < NSEG; ++k)
                a[k] = io.p[k][q];
            M::qp1(prm, a, ta);
#pragma unroll
            for (int k = 0; k < NSEG; ++k)
                if (M::wr(k))
                    io.p[k][q] = a[k];
            if (tangent != nullptr)
                tangent[q] = ta;
        }
    }
}

}  // namespace fcx
