// fcx_internal.h -- declarations shared between the translation units of libfcx.so.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace fcx {
extern thread_local char g_cuda_error[256];
extern std::atomic<unsigned long long> g_launches;
int note_cuda_error(cudaError_t e, const char *where);
int sm_count();
}  // namespace fcx
