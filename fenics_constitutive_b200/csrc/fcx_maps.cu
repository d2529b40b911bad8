// fcx_maps.cu -- parent <-> sub-mesh maps of quadrature arrays on the device
// (reference src/fenics_constitutive/solver/maps.py:62-123, SubSpaceMap.map_to_sub /
// map_to_parent; used by LawOnSubMesh, solver/_lawonsubmesh.py:58-70).
//
// A quadrature array is [cell][qp][...] flat (reference tests/solver/test_maps.py:119-121), so
// the map of a law that owns the cell list `cells` moves whole ROWS of `row` doubles:
//     to_sub:     sub[i][:]           = parent[cells[i]][:]
//     to_parent:  parent[cells[i]][:] = sub[i][:]
// One side of every copy is contiguous over the whole array, the other is contiguous per row
// (48 B .. 1152 B for s = 6: stress with 1 QP .. tangent with 4 QPs).  A group of G lanes
// (G = smallest power of two >= 16-byte chunks per row, at most 32) owns one row at a time, so a
// warp instruction touches 32/G rows x 16 G contiguous bytes: every sector it opens is used
// completely.  Rows with an odd number of doubles or unaligned views move 8 bytes per lane.
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/fcx.h"
#include "fcx_internal.h"

namespace fcx {

template <bool TO_PARENT, class V>
__global__ void __launch_bounds__(256)
    map_rows_kernel(const V *__restrict__ src, V *__restrict__ dst, const int *__restrict__ cells,
                    unsigned long long nrows, unsigned row_elems, unsigned group)
{
    // `group` lanes per row; rows are strided over all groups of the grid
    const unsigned lane_in_group = threadIdx.x & (group - 1);
    const unsigned long long gid = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) / group;
    const unsigned long long ngroups = ((unsigned long long)gridDim.x * blockDim.x) / group;
    for (unsigned long long i = gid; i < nrows; i += ngroups) {
        const unsigned long long c = (unsigned long long)cells[i];
        const V *s = src + (TO_PARENT ? i : c) * row_elems;
        V *d = dst + (TO_PARENT ? c : i) * row_elems;
        for (unsigned k = lane_in_group; k < row_elems; k += group)
            d[k] = s[k];
    }
}

template <bool TO_PARENT>
static int launch_map_rows(size_t nrows, size_t row, const int *cells, const double *src, double *dst,
                           cudaStream_t st)
{
    if (nrows == 0 || row == 0)
        return FCX_OK;
    if (!cells || !src || !dst)
        return FCX_ERR_NULL;
    if (row > 0x7fffffffULL)
        return FCX_ERR_ARG;
    const bool vec = row % 2 == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15u) == 0;
    const unsigned elems = (unsigned)(vec ? row / 2 : row);
    unsigned group = 1;
    while (group < elems && group < 32)
        group <<= 1;
    const unsigned long long threads = (unsigned long long)nrows * group;
    unsigned long long grid = (threads + 255) / 256;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (grid > cap)
        grid = cap;
    if (vec)
        map_rows_kernel<TO_PARENT, double2><<<(unsigned)grid, 256, 0, st>>>(
            reinterpret_cast<const double2 *>(src), reinterpret_cast<double2 *>(dst), cells, nrows, elems, group);
    else
        map_rows_kernel<TO_PARENT, double><<<(unsigned)grid, 256, 0, st>>>(src, dst, cells, nrows, elems, group);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "map_rows_kernel launch");
}

// du = u - u_prev, 16 bytes per access where the three views allow it
__global__ void __launch_bounds__(256)
    nodal_increment_kernel(const double *__restrict__ u, const double *__restrict__ u_prev, double *__restrict__ du,
                           unsigned long long n, int vec)
{
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        const unsigned long long n2 = n / 2;
        const double2 *a = reinterpret_cast<const double2 *>(u), *b = reinterpret_cast<const double2 *>(u_prev);
        double2 *d = reinterpret_cast<double2 *>(du);
        for (unsigned long long i = t; i < n2; i += stride) {
            const double2 x = a[i], y = b[i];
            d[i] = make_double2(x.x - y.x, x.y - y.y);
        }
        if ((n & 1ULL) && t == 0)
            du[n - 1] = u[n - 1] - u_prev[n - 1];
    } else {
        for (unsigned long long i = t; i < n; i += stride)
            du[i] = u[i] - u_prev[i];
    }
}

}  // namespace fcx

using namespace fcx;

extern "C" {

int fcx_nodal_increment(size_t n, const double *u, const double *u_prev, double *du, void *stream)
{
    if (n == 0)
        return FCX_OK;
    if (!u || !u_prev || !du)
        return FCX_ERR_NULL;
    const int vec = ((reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(u_prev) |
                      reinterpret_cast<uintptr_t>(du)) & 15u) == 0;
    unsigned long long grid = (n / 2 + 255) / 256 + 1;
    const unsigned long long cap = (unsigned long long)sm_count() * 8;
    if (grid > cap)
        grid = cap;
    nodal_increment_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(u, u_prev, du, n, vec);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return note_cuda_error(cudaGetLastError(), "nodal_increment_kernel launch");
}

int fcx_map_rows_to_sub(size_t nrows_sub, size_t row_doubles, const int *cells, const double *parent,
                        double *sub, void *stream)
{
    return launch_map_rows<false>(nrows_sub, row_doubles, cells, parent, sub, static_cast<cudaStream_t>(stream));
}

int fcx_map_rows_to_parent(size_t nrows_sub, size_t row_doubles, const int *cells, const double *sub,
                           double *parent, void *stream)
{
    return launch_map_rows<true>(nrows_sub, row_doubles, cells, sub, parent, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
