// kpconv_common.cuh -- the geometric half of KPConv shared by every KPConv kernel:
// kernel-point influence weights of one query point, computed by one warp.
//
// Follows torch_points3d/modules/KPConv/convolution_ops.py:49-87 in the reference's fp32
// operation order (centre the neighbour, subtract the kernel point, (dx*dx+dy*dy)+dz*dz,
// sqrt, divide by the extent, clamp) so the clamp-at-zero boundary falls on the same
// neighbours as in the reference.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace p3d {

constexpr int KP_K = 16;  // kernel points are padded to 16 (the reference uses 15)

struct KpGeom {
    float extent;
    float gauss_den;  // 2*(0.3*extent)^2 + 1e-9   (radius_gaussian, convolution_ops.py:7-16)
    int influence;
    int aggregation;
};

__device__ __forceinline__ float kp_influence(float d2, const KpGeom &g) {
    if (g.influence == P3D_INFLUENCE_CONSTANT) return 1.0f;
    if (g.influence == P3D_INFLUENCE_LINEAR) {
        const float w = __fsub_rn(1.0f, __fdiv_rn(__fsqrt_rn(d2), g.extent));
        return w > 0.0f ? w : 0.0f;
    }
    return expf(-__fdiv_rn(d2, g.gauss_den));
}

// Weights of point p for all (neighbour m, kernel point k): wsm[m*16 + k], m in [0, Mpad),
// Mpad = M rounded up to 32; rows m >= M and columns k >= K are written as 0.
// jsm[m] = feature row of neighbour m, or -1 for a shadow neighbour (index < 0 or >= N0:
// the point at 1e6 with a zero feature row, convolution_ops.py:49-50,92).
// lane owns kernel point (lane & 15): its coordinates are passed in registers.
__device__ __forceinline__ void kp_point_weights(const float *__restrict__ q, const float *__restrict__ s,
                                                 const int64_t *__restrict__ nbp, int64_t N0, int M, int K,
                                                 const KpGeom &g, float kx, float ky, float kz, float *wsm, int *jsm,
                                                 int lane) {
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    const int k = lane & 15;
    for (int m0 = 0; m0 < M; m0 += 32) {
        const int mm = m0 + lane;
        int64_t j = mm < M ? __ldg(nbp + mm) : -1;
        const bool shadow = (j < 0 || j >= N0);
        float cx = 1e6f, cy = 1e6f, cz = 1e6f;
        if (!shadow) {
            cx = __ldg(s + 3 * j);
            cy = __ldg(s + 3 * j + 1);
            cz = __ldg(s + 3 * j + 2);
        }
        cx = __fsub_rn(cx, qx);
        cy = __fsub_rn(cy, qy);
        cz = __fsub_rn(cz, qz);
        jsm[mm] = shadow ? -1 : (int)j;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int src = 2 * t + (lane >> 4);  // neighbour (inside this chunk) of this pair
            const float ax = __shfl_sync(0xffffffffu, cx, src);
            const float ay = __shfl_sync(0xffffffffu, cy, src);
            const float az = __shfl_sync(0xffffffffu, cz, src);
            const float d2 = sqdist3(ax, ay, az, kx, ky, kz);
            float w = kp_influence(d2, g);
            if (g.aggregation == P3D_AGGREGATION_CLOSEST) {
                // argmin over the 16 kernel points of this neighbour (first minimum wins)
                float bd = (k < K) ? d2 : CUDART_INF_F;
                int bk = k;
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) {
                    const float od = __shfl_xor_sync(0xffffffffu, bd, off);
                    const int ok = __shfl_xor_sync(0xffffffffu, bk, off);
                    if (od < bd || (od == bd && ok < bk)) { bd = od; bk = ok; }
                }
                if (bk != k) w = 0.0f;
            }
            if (k >= K || m0 + src >= M) w = 0.0f;
            wsm[(m0 + src) * KP_K + k] = w;  // == wsm[m0*16 + 32*t + lane]: conflict-free
        }
    }
    __syncwarp();
}

}  // namespace p3d
