// pcie_probe.cu -- sizes the host-array (e2e) path: PCIe copy rates, zero-copy
// store rate from a kernel into mapped pinned memory, and what the host threads
// can do (pageable<->pinned memcpy, symmetric 21->36 tangent expansion).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/pcie_probe scripts/pcie_probe.cu -lpthread
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

__global__ void zc_store(double2 *dst, size_t n2)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride)
        dst[i] = make_double2((double)i, 1.0);
}

__global__ void zc_load(const double2 *src, double2 *dst, size_t n2)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride)
        dst[i] = src[i];
}

template <class F>
static double par(int T, F &&f)
{
    std::vector<std::thread> th;
    const double t0 = now();
    for (int t = 0; t < T; ++t)
        th.emplace_back([&, t] { f(t); });
    for (auto &x : th)
        x.join();
    return now() - t0;
}

static const int IJ[21][2] = {{0, 0}, {0, 1}, {0, 2}, {0, 3}, {0, 4}, {0, 5}, {1, 1}, {1, 2}, {1, 3}, {1, 4}, {1, 5},
                              {2, 2}, {2, 3}, {2, 4}, {2, 5}, {3, 3}, {3, 4}, {3, 5}, {4, 4}, {4, 5}, {5, 5}};

static void expand(const double *w, double *c, size_t n)
{
    for (size_t q = 0; q < n; ++q) {
        const double *p = w + 21 * q;
        double *o = c + 36 * q;
        for (int k = 0; k < 21; ++k) {
            o[IJ[k][0] * 6 + IJ[k][1]] = p[k];
            o[IJ[k][1] * 6 + IJ[k][0]] = p[k];
        }
    }
}

int main(int argc, char **argv)
{
    const size_t GB = (size_t)1 << 30;
    const size_t bytes = 2 * GB;
    const int maxT = (int)std::thread::hardware_concurrency();
    printf("host threads: %d\n", maxT);
    char *pin_a, *pin_b, *dev_a, *dev_b;
    CK(cudaMallocHost(&pin_a, bytes));
    CK(cudaMallocHost(&pin_b, bytes));
    CK(cudaMalloc(&dev_a, bytes));
    CK(cudaMalloc(&dev_b, bytes));
    memset(pin_a, 1, bytes);
    memset(pin_b, 2, bytes);
    cudaStream_t s1, s2;
    CK(cudaStreamCreate(&s1));
    CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms;

    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0, s1));
        CK(cudaMemcpyAsync(dev_a, pin_a, bytes, cudaMemcpyHostToDevice, s1));
        CK(cudaEventRecord(e1, s1));
        CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("H2D alone            %.1f GB/s\n", bytes / ms / 1e6);
        CK(cudaEventRecord(e0, s1));
        CK(cudaMemcpyAsync(pin_b, dev_b, bytes, cudaMemcpyDeviceToHost, s1));
        CK(cudaEventRecord(e1, s1));
        CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("D2H alone            %.1f GB/s\n", bytes / ms / 1e6);
        double t0 = now();
        CK(cudaMemcpyAsync(dev_a, pin_a, bytes, cudaMemcpyHostToDevice, s1));
        CK(cudaMemcpyAsync(pin_b, dev_b, bytes, cudaMemcpyDeviceToHost, s2));
        CK(cudaStreamSynchronize(s1));
        double th = now() - t0;
        CK(cudaStreamSynchronize(s2));
        double td = now() - t0;
        printf("H2D || D2H           %.1f + %.1f GB/s\n", bytes / th / 1e9, bytes / td / 1e9);
    }
    // zero-copy: SM stores into mapped pinned memory
    for (int grid : {148, 592, 2368}) {
        CK(cudaEventRecord(e0, s1));
        zc_store<<<grid, 256, 0, s1>>>((double2 *)pin_b, bytes / 16);
        CK(cudaEventRecord(e1, s1));
        CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("zero-copy store grid=%d   %.1f GB/s\n", grid, bytes / ms / 1e6);
    }
    {
        CK(cudaEventRecord(e0, s1));
        zc_load<<<592, 256, 0, s1>>>((const double2 *)pin_a, (double2 *)dev_a, bytes / 16);
        CK(cudaEventRecord(e1, s1));
        CK(cudaStreamSynchronize(s1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("zero-copy load grid=592    %.1f GB/s\n", bytes / ms / 1e6);
        double t0 = now();
        zc_store<<<592, 256, 0, s1>>>((double2 *)pin_b, bytes / 16);
        CK(cudaMemcpyAsync(dev_a, pin_a, bytes, cudaMemcpyHostToDevice, s2));
        CK(cudaStreamSynchronize(s2));
        double th = now() - t0;
        CK(cudaStreamSynchronize(s1));
        double td = now() - t0;
        printf("H2D DMA || zero-copy store  %.1f + %.1f GB/s\n", bytes / th / 1e9, bytes / td / 1e9);
    }
    // host threads
    char *page_a = (char *)malloc(bytes), *page_b = (char *)malloc(bytes);
    memset(page_a, 3, bytes);
    memset(page_b, 4, bytes);
    for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
        if (T > maxT)
            break;
        const size_t per = bytes / T;
        double t = par(T, [&](int i) { memcpy(pin_b + i * per, page_a + i * per, per); });
        printf("T=%2d memcpy pageable->pinned  %.1f GB/s (payload)\n", T, bytes / t / 1e9);
        t = par(T, [&](int i) { memcpy(page_b + i * per, pin_a + i * per, per); });
        printf("T=%2d memcpy pinned->pageable  %.1f GB/s (payload)\n", T, bytes / t / 1e9);
        // expansion 21->36: wire in pin_a (168 B/QP), tangent into page_b (288 B/QP)
        const size_t nq = bytes / 288;
        const size_t qper = nq / T;
        t = par(T, [&](int i) { expand((const double *)pin_a + 21 * qper * i, (double *)page_b + 36 * qper * i, qper); });
        printf("T=%2d expand 21->36            %.1f MQP/s  (%.1f GB/s written)\n", T, qper * T / t / 1e6,
               qper * T * 288 / t / 1e9);
    }
    // expansion while DMA is running both ways (contention for host memory)
    {
        const int T = maxT > 2 ? maxT - 2 : maxT;
        const size_t nq = bytes / 288, qper = nq / T;
        std::thread dma([&] {
            for (int k = 0; k < 4; ++k) {
                CK(cudaMemcpyAsync(dev_a, pin_a, bytes, cudaMemcpyHostToDevice, s1));
                CK(cudaMemcpyAsync(pin_b, dev_b, bytes, cudaMemcpyDeviceToHost, s2));
            }
            CK(cudaStreamSynchronize(s1));
            CK(cudaStreamSynchronize(s2));
        });
        double t = par(T, [&](int i) { expand((const double *)pin_a + 21 * qper * i, (double *)page_b + 36 * qper * i, qper); });
        double t0 = now();
        dma.join();
        printf("T=%2d expand under DMA load     %.1f MQP/s (DMA finished %.2f s later)\n", T, qper * T / t / 1e6,
               now() - t0);
    }
    return 0;
}
