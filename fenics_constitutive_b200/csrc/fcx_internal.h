// fcx_internal.h -- declarations shared between the translation units of libfcx.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace fcx {
extern thread_local char g_cuda_error[256];
extern std::atomic<unsigned long long> g_launches;
int note_cuda_error(cudaError_t e, const char *where);
int sm_count();
// Atomic tile-ticket counter of (current device, stream), zeroed on the stream;
// nullptr when dynamic tile hand-out is switched off (fcx_tune "dynamic_tiles").
unsigned long long *tile_ticket(cudaStream_t stream);
// fcx_tune "ctas_per_sm" (0 = occupancy query) and "fem_variant"
// (1 = QP-parallel bulk-staged FEM kernels, 0 = one-thread-per-cell kernels).
int tuned_ctas_per_sm();
int fem_variant();
// fcx_tune "gather_variant": 1 = gather_staged_kernel (cp.async-staged nodal values), 0 = gather_kernel.
int gather_variant();
// Launch gate of the FEM element kernels (fcx_assemble.cu): while set (per host thread), every element kernel
// launched by this thread returns at once if *gate != 0 when it starts -- the device-resident Krylov loop
// (fcx_krylov.cu) freezes its iterations on the device once the residual test has passed, so the host may
// enqueue blocks of iterations ahead of reading the test's outcome.  nullptr = no gate.
void fem_set_launch_gate(const double *gate);
// Tile tickets of the element kernels without a memset per launch: while set (per host thread), launch i takes
// pair[i & 1] as its ticket counter and zeroes pair[(i + 1) & 1] -- the counter of the NEXT launch on the same
// stream, which the previous launch has finished with.  Both counters must be zero when the pair is first used.
// `launches` (host memory of the pair's owner) counts the launches the pair has served, across calls.
// nullptr = the per-stream counter with its cudaMemsetAsync (tile_ticket).
void fem_set_launch_tickets(unsigned long long *pair, unsigned *launches);

// Resident CTAs per SM of `kern` at (threads, smem), with the opt-in to > 48 KB of dynamic shared
// memory set first.  Both are properties of (kernel, DEVICE): a process may drive several GPUs (the
// Python layer binds every call to the device that owns the tensors), so the cache is per device
// index, one OccCache per kernel instantiation (a function-local static of its launcher).
constexpr int FCX_MAX_DEVICES = 64;
struct OccCache {
    int v[FCX_MAX_DEVICES];
    OccCache()
    {
        for (int i = 0; i < FCX_MAX_DEVICES; ++i)
            v[i] = -1;
    }
};
template <class Kern>
static inline int kernel_occupancy(OccCache &cache, Kern kern, int threads, size_t smem, const char *what,
                                   int *occ_out)
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return note_cuda_error(e, what);
    const bool cached = dev >= 0 && dev < FCX_MAX_DEVICES;
    if (cached && cache.v[dev] > 0) {
        *occ_out = cache.v[dev];
        return 0;
    }
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess)
        return note_cuda_error(e, what);
    int o = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, threads, smem);
    if (e != cudaSuccess)
        return note_cuda_error(e, what);
    o = o > 0 ? o : 1;
    if (cached)
        cache.v[dev] = o;
    *occ_out = o;
    return 0;
}
}  // namespace fcx
