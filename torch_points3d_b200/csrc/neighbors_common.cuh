// neighbors_common.cuh -- cloud-range bookkeeping shared by the neighbour-search kernels (neighbors.cu, cellgrid.cu):
// the candidate range [x0, x1) of the cloud that owns a query, found on the device from the sorted batch vectors.
#pragma once
#include "common.cuh"

namespace p3d {

struct RangeCache {
    int64_t c[2], x0[2], x1[2];
};

// candidate range [x0, x1) of the cloud that owns query q (executed by a whole warp)
template <bool DENSE>
__device__ __forceinline__ void query_range(int64_t q, const int64_t *__restrict__ batch_x,
                                            const int64_t *__restrict__ batch_y, int64_t Nx, int64_t n_dense,
                                            int64_t s_dense, const RangeCache &rc, int lane, int64_t &x0,
                                            int64_t &x1) {
    if (DENSE) {
        const int64_t b = q / s_dense;
        x0 = b * n_dense;
        x1 = x0 + n_dense;
    } else if (batch_x == nullptr || batch_y == nullptr) {
        x0 = 0;
        x1 = Nx;
    } else {
        const int64_t c = batch_y[q];
        if (c == rc.c[0]) { x0 = rc.x0[0]; x1 = rc.x1[0]; }
        else if (c == rc.c[1]) { x0 = rc.x0[1]; x1 = rc.x1[1]; }
        else {
            x0 = warp_lower_bound(batch_x, Nx, c, lane);
            x1 = warp_lower_bound(batch_x, Nx, c + 1, lane);
        }
    }
}

template <bool DENSE>
__device__ __forceinline__ void cta_ranges(int64_t q0, int64_t q1, const int64_t *__restrict__ batch_x,
                                           const int64_t *__restrict__ batch_y, int64_t Nx, RangeCache &rc,
                                           int warp, int lane) {
    if (!DENSE && batch_x != nullptr && batch_y != nullptr) {
        if (warp < 2) {
            const int64_t c = batch_y[warp == 0 ? q0 : q1 - 1];
            const int64_t a = warp_lower_bound(batch_x, Nx, c, lane);
            const int64_t b = warp_lower_bound(batch_x, Nx, c + 1, lane);
            if (lane == 0) { rc.c[warp] = c; rc.x0[warp] = a; rc.x1[warp] = b; }
        }
    }
    __syncthreads();
}

}  // namespace p3d
