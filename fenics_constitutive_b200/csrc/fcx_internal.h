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
}  // namespace fcx
