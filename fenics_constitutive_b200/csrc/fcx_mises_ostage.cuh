// fcx_mises_ostage.cuh -- "output-staged" Mises kernel: ALL global traffic of a
// tile, the 6x6 tangent included, moves through the TMA engine.
//
// The generic tile pipeline (fcx_tile.cuh) regenerates the tangent block from a
// per-QP record with a warp-cooperative store loop whose per-element index
// arithmetic costs ~1000 issue slots per QP -- more than the constitutive
// update itself.  Here every thread writes the 36 entries of ITS OWN QP into a
// dense [TILE][36] shared-memory block with compile-time indexing (21 unique
// products, 18 16-byte shared stores) and ONE bulk async store streams the
// block's 36 864 contiguous bytes to HBM.
//
// Shared memory per tile is a single region of 49 doubles (392 B) per QP:
//     A: stress [T][6] | eps_n [T][6] | alpha [T]      loaded, updated in place, stored
//     B: tangent [T][36]                               stored;  grad_del_u [T][9]
//        is loaded into the head of B and consumed into registers before B
//        is overwritten (one __syncthreads in between).
// There is no second stage: a CTA's load latency is exposed and hidden by the
// other resident CTAs (4 x 128 threads per SM), which is what buys the smaller
// footprint (392 B/QP instead of 440 B/QP + tangent staging).
//
// Only full tiles with 16-byte aligned pointers come here; tails and unaligned
// views go through fcx_tile_kernel's generic path (see fcx_api.cu).
#pragma once
#include "fcx_models.cuh"

namespace fcx {

// WITH_TANGENT = false: stress-only evaluate (the reference's `tangent: Option<..>`,
// comfe-rs/src/interfaces.rs:368, bindings/src/lib.rs:83,109-113): region B shrinks to the
// grad_del_u slots [T][9], nothing of the tangent is formed or stored (280 instead of 568 B/QP).
template <int TILE, bool WITH_TANGENT = true>
constexpr size_t mises_ostage_smem_bytes()
{
    return sizeof(double) * (WITH_TANGENT ? 49 : 22) * TILE + sizeof(uint64_t);
}

template <int TILE, int MINCTAS, bool WITH_TANGENT = true>
__global__ void __launch_bounds__(TILE, MINCTAS)
    fcx_mises_ostage_kernel(const __grid_constant__ MisesParams P, const double *__restrict__ grad,
                            double *__restrict__ stress, double *__restrict__ tangent,
                            double *__restrict__ eps_n, double *__restrict__ alpha,
                            const unsigned long long ntiles, unsigned char *__restrict__ flag,
                            int *__restrict__ status, unsigned long long *__restrict__ ticket)
{
    // ticket != nullptr: tiles are handed out by an atomic counter (the first
    // gridDim.x tiles are implicit), which keeps the set of tiles in flight a
    // tight moving window in every array; nullptr: static stride gridDim.x.
    extern __shared__ __align__(128) double smem[];
    double *s_sig = smem;              // [TILE][6]
    double *s_eps = smem + 6 * TILE;   // [TILE][6]
    double *s_alp = smem + 12 * TILE;  // [TILE]
    double *s_tan = smem + 13 * TILE;  // [TILE][36]; grad [TILE][9] at its head on load
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + (WITH_TANGENT ? 49 : 22) * TILE);

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t parity = 0;
    __shared__ unsigned long long s_next;
    unsigned long long tile = blockIdx.x;
    while (tile < ntiles) {
        const unsigned long long q0 = tile * TILE;
        if (tid == 0) {
            s_next = (ticket != nullptr) ? gridDim.x + atomicAdd(ticket, 1ULL) : tile + gridDim.x;
            // the previous tile's bulk stores must have finished READING the region
            bulk_wait_read_all();
            mbar_arrive_expect_tx(bar, (uint32_t)(22 * TILE * sizeof(double)));
            bulk_g2s(s_tan, grad + q0 * 9, TILE * 9 * sizeof(double), bar);
            bulk_g2s(s_sig, stress + q0 * 6, TILE * 6 * sizeof(double), bar);
            bulk_g2s(s_eps, eps_n + q0 * 6, TILE * 6 * sizeof(double), bar);
            bulk_g2s(s_alp, alpha + q0, TILE * sizeof(double), bar);
        }
        mbar_wait(bar, parity);
        parity ^= 1;

        double g[9], sig[6], ep[6];
#pragma unroll
        for (int i = 0; i < 9; ++i)
            g[i] = s_tan[tid * 9 + i];
#pragma unroll
        for (int i = 0; i < 6; i += 2) {
            const double2 a = *reinterpret_cast<const double2 *>(s_sig + tid * 6 + i);
            const double2 b = *reinterpret_cast<const double2 *>(s_eps + tid * 6 + i);
            sig[i] = a.x;
            sig[i + 1] = a.y;
            ep[i] = b.x;
            ep[i + 1] = b.y;
        }
        double al = s_alp[tid];
        __syncthreads();  // every grad slot is in registers; B may be overwritten
        const unsigned long long next = s_next;  // thread 0 rewrites it only after the next barrier

        bool plastic = false, failed = false;
        double coef[4], xn[6];
        mises_point(P, g, sig, ep, al, coef, xn, plastic, failed);
        if (flag != nullptr)
            flag[q0 + tid] = plastic ? 1 : 0;
        if (failed && status != nullptr) {
            atomicAdd(&status[0], 1);
            const unsigned long long q = q0 + tid;
            atomicMin(&status[1], q > 0x7fffffffULL ? 0x7fffffff : (int)q);
        }

#pragma unroll
        for (int i = 0; i < 6; i += 2) {
            *reinterpret_cast<double2 *>(s_sig + tid * 6 + i) = make_double2(sig[i], sig[i + 1]);
            *reinterpret_cast<double2 *>(s_eps + tid * 6 + i) = make_double2(ep[i], ep[i + 1]);
        }
        s_alp[tid] = al;

        if constexpr (WITH_TANGENT) {
        // aah = ka*xioi + cpp*xpp + cnn*outer(xn, xn)   (:170-175), symmetric:
        // 21 unique entries, evaluated as base + cnn*(xn_i*xn_j) like the reference.
        double c[6][6];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = i; j < 6; ++j) {
                const double base = (j < 3) ? ((i == j) ? coef[0] : coef[1])
                                            : ((i == j) ? coef[2] : 0.0);
                c[i][j] = base + coef[3] * (xn[i] * xn[j]);
            }
        double *row = s_tan + tid * 36;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int j = 0; j < 6; j += 2) {
                const double a = (i <= j) ? c[i][j] : c[j][i];
                const double b = (i <= j + 1) ? c[i][j + 1] : c[j + 1][i];
                *reinterpret_cast<double2 *>(row + i * 6 + j) = make_double2(a, b);
            }
        }

        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            if constexpr (WITH_TANGENT)
                bulk_s2g(tangent + q0 * 36, s_tan, TILE * 36 * sizeof(double));
            bulk_s2g(stress + q0 * 6, s_sig, TILE * 6 * sizeof(double));
            bulk_s2g(eps_n + q0 * 6, s_eps, TILE * 6 * sizeof(double));
            bulk_s2g(alpha + q0, s_alp, TILE * sizeof(double));
            bulk_commit();
        }
        tile = next;
    }
    if (tid == 0)
        bulk_wait_read_all();
}

}  // namespace fcx
