// pcie_probe2.cu -- how fast can SM stores into mapped pinned memory ("zero-copy") go for the
// access patterns of the download wire (csrc/fcx_host.cu), alone and while an H2D DMA runs?
//   a  contiguous 16-byte stores (baseline of pcie_probe.cu)
//   b  288-byte runs (one 6x6 tangent) for a random 50 % of the points, 16-byte stores
//   c  as b, but every 128-byte line touched by a plastic point is written whole
//   d  compact stream written as 8-byte stores by consecutive lanes, gaps of idle lanes
//      (the first pack kernel: lanes of elastic points idle)
//   e  compact stream, 8-byte stores, every warp writes 256 contiguous aligned bytes
//   f  as e with 16-byte stores (512 contiguous bytes per warp)
// and chunked (64 Ki points per launch, as the pipeline does) vs one big launch.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/pcie_probe2 scripts/pcie_probe2.cu
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e_ = (x);                                                          \
        if (e_ != cudaSuccess) {                                                       \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

__global__ void pat_a(double2 *dst, size_t n2)
{
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride)
        dst[i] = make_double2((double)i, 1.0);
}

// 18 double2 per point, only flagged points
__global__ void pat_b(const unsigned char *flag, const double2 *src, double2 *dst, size_t npts)
{
    const size_t total = npts * 18, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride)
        if (flag[i / 18])
            dst[i] = src[i];
}

// whole 128-byte lines (8 double2) if any point overlapping the line is flagged
__global__ void pat_c(const unsigned char *flag, const double2 *src, double2 *dst, size_t npts)
{
    const size_t total = npts * 18, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t l0 = i & ~(size_t)7, l1 = l0 + 7 < total ? l0 + 7 : total - 1;
        if (flag[l0 / 18] | flag[l1 / 18])
            dst[i] = src[i];
    }
}

// compact stream, lanes of unflagged points idle (R doubles per record)
__global__ void pat_d(const unsigned char *flag, const unsigned *pos, const double *src, double *dst, size_t npts, int R)
{
    const size_t total = npts * R, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += stride) {
        const size_t q = i / R;
        if (flag[q])
            dst[(size_t)pos[q] * R + (i - q * R)] = src[i];
    }
}

// compact stream, output-centric: thread o writes dst[o]
__global__ void pat_e(const unsigned *list, const double *src, double *dst, size_t nout, int R)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < nout; o += stride) {
        const size_t r = o / R;
        dst[o] = src[(size_t)list[r] * R + (o - r * R)];
    }
}

__global__ void pat_f(const unsigned *list, const double *src, double2 *dst, size_t nout2, int R)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < nout2; o += stride) {
        const size_t r = (2 * o) / R;  // R even
        const double *p = src + (size_t)list[r] * R + (2 * o - r * R);
        dst[o] = make_double2(p[0], p[1]);
    }
}

int main()
{
    const size_t npts = (size_t)4 << 20;  // 4 Mi points
    const size_t tbytes = npts * 288;     // 1.2 GB tangent array
    char *pin_t, *pin_in, *dev_t, *dev_in;
    CK(cudaMallocHost(&pin_t, tbytes));
    CK(cudaMallocHost(&pin_in, tbytes));
    CK(cudaMalloc(&dev_t, tbytes));
    CK(cudaMalloc(&dev_in, tbytes));
    memset(pin_t, 0, tbytes);
    memset(pin_in, 1, tbytes);
    CK(cudaMemset(dev_t, 3, tbytes));
    std::vector<unsigned char> flag(npts);
    std::vector<unsigned> pos(npts), list;
    srand(7);
    for (size_t q = 0; q < npts; ++q) {
        flag[q] = (rand() & 1);
        pos[q] = (unsigned)list.size();
        if (flag[q])
            list.push_back((unsigned)q);
    }
    const size_t nplastic = list.size();
    unsigned char *d_flag;
    unsigned *d_pos, *d_list;
    CK(cudaMalloc(&d_flag, npts));
    CK(cudaMalloc(&d_pos, npts * 4));
    CK(cudaMalloc(&d_list, npts * 4));
    CK(cudaMemcpy(d_flag, flag.data(), npts, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pos, pos.data(), npts * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_list, list.data(), nplastic * 4, cudaMemcpyHostToDevice));
    cudaStream_t s1, s2;
    CK(cudaStreamCreate(&s1));
    CK(cudaStreamCreate(&s2));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int grid = 1184;
    const size_t chunk = 65536;

    auto run = [&](const char *name, double payload_bytes, auto &&launch_all) {
        for (int with_h2d = 0; with_h2d < 2; ++with_h2d) {
            float ms = 0;
            double th2d = 0;
            const double t0 = now();
            CK(cudaEventRecord(e0, s1));
            launch_all();
            CK(cudaEventRecord(e1, s1));
            if (with_h2d) {
                CK(cudaMemcpyAsync(dev_in, pin_in, tbytes, cudaMemcpyHostToDevice, s2));
                CK(cudaStreamSynchronize(s2));
                th2d = now() - t0;
            }
            CK(cudaStreamSynchronize(s1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (with_h2d)
                printf("%-58s || H2D  %6.1f GB/s payload   (H2D %5.1f GB/s)\n", name, payload_bytes / ms / 1e6, tbytes / th2d / 1e9);
            else
                printf("%-58s alone   %6.1f GB/s payload\n", name, payload_bytes / ms / 1e6);
        }
    };

    run("a contiguous 16 B stores, one launch", (double)tbytes,
        [&] { pat_a<<<grid, 256, 0, s1>>>((double2 *)pin_t, tbytes / 16); });
    run("a contiguous 16 B stores, 64 Ki-point launches", (double)tbytes, [&] {
        for (size_t q0 = 0; q0 < npts; q0 += chunk)
            pat_a<<<grid, 256, 0, s1>>>((double2 *)pin_t + q0 * 18, chunk * 18);
    });
    run("b 288 B runs, 50 % of points, one launch", nplastic * 288.0,
        [&] { pat_b<<<grid, 256, 0, s1>>>(d_flag, (const double2 *)dev_t, (double2 *)pin_t, npts); });
    run("b 288 B runs, 50 % of points, 64 Ki-point launches", nplastic * 288.0, [&] {
        for (size_t q0 = 0; q0 < npts; q0 += chunk)
            pat_b<<<grid, 256, 0, s1>>>(d_flag + q0, (const double2 *)dev_t + q0 * 18, (double2 *)pin_t + q0 * 18, chunk);
    });
    run("c whole 128 B lines touched by plastic points, one launch", nplastic * 288.0,
        [&] { pat_c<<<grid, 256, 0, s1>>>(d_flag, (const double2 *)dev_t, (double2 *)pin_t, npts); });
    for (int R : {28, 7}) {
        char name[128];
        snprintf(name, sizeof name, "d compact R=%d, 8 B stores, idle lanes, one launch", R);
        run(name, nplastic * R * 8.0,
            [&] { pat_d<<<grid, 256, 0, s1>>>(d_flag, d_pos, (const double *)dev_t, (double *)pin_t, npts, R); });
        snprintf(name, sizeof name, "e compact R=%d, 8 B stores, output-centric, one launch", R);
        run(name, nplastic * R * 8.0,
            [&] { pat_e<<<grid, 256, 0, s1>>>(d_list, (const double *)dev_t, (double *)pin_t, nplastic * R, R); });
    }
    run("f compact R=28, 16 B stores, output-centric, one launch", nplastic * 28 * 8.0,
        [&] { pat_f<<<grid, 256, 0, s1>>>(d_list, (const double *)dev_t, (double2 *)pin_t, nplastic * 14, 28); });
    // DMA of the same compact volume for reference
    run("DMA D2H of the compact R=28 volume", nplastic * 28 * 8.0,
        [&] { CK(cudaMemcpyAsync(pin_t, dev_t, nplastic * 28 * 8, cudaMemcpyDeviceToHost, s1)); });
    run("DMA D2H in 64 Ki-point pieces (7.3 MB each)", nplastic * 28 * 8.0, [&] {
        const size_t piece = 32768 * 28 * 8;
        for (size_t off = 0; off + piece <= nplastic * 28 * 8; off += piece)
            CK(cudaMemcpyAsync(pin_t + off, dev_t + off, piece, cudaMemcpyDeviceToHost, s1));
    });
    return 0;
}
