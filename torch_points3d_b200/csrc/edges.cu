// edges.cu -- MESSAGE_PASSING-format operators: what runs either side of the
// per-edge MLP in the reference's PointConv (modules/pointnet2/message_passing.py:9-26, torch_geometric 1.7.2
// PointConv.message / aggr="max") and RandlaKernel (modules/RandLANet/modules.py:9-54, aggr="add").
//
// Edge lists come from `radius` / `knn` (core/spatial_ops/neighbour_finder.py:32-33,47): source j = support
// index, target i = query index, sorted by target.  All four kernels are bandwidth-bound row gathers / segment
// reductions: warps own edges (or targets), lanes run over channels so every row access is one coalesced line.
//   edge_message     out[e] = [ x[src[e]] , pos_src[src[e]] - pos_dst[dst[e]] ]         (PointConv.message input)
//   edge_encoding    out[e] = [ pos_i, pos_j, pos_i - pos_j, |pos_i - pos_j| ]          (RandlaKernel.message :34-40)
//   segment_max/sum  out[t] = reduce over the contiguous edges of target t             (MessagePassing.aggregate)
// Segment bounds are found on the device (32-ary warp search over the sorted target vector): no host sync.
#include "common.cuh"

namespace p3d {

constexpr int ED_WARPS = 8;

// out[e, 0:C] = x[src[e], :], out[e, C:C+3] = pos_src[src[e]] - pos_dst[dst[e]]
__global__ void __launch_bounds__(ED_WARPS * 32)
edge_message_kernel(const float *__restrict__ x, const float *__restrict__ pos_src, const float *__restrict__ pos_dst,
                    const int64_t *__restrict__ src, const int64_t *__restrict__ dst, int64_t E, int C,
                    float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int W = C + 3;
    for (int64_t e = (int64_t)blockIdx.x * ED_WARPS + (threadIdx.x >> 5); e < E; e += (int64_t)gridDim.x * ED_WARPS) {
        const int64_t j = __ldg(src + e), i = __ldg(dst + e);
        float *o = out + e * W;
        for (int c = lane; c < C; c += 32) o[c] = __ldg(x + j * C + c);
        if (lane < 3) o[C + lane] = __fsub_rn(__ldg(pos_src + 3 * j + lane), __ldg(pos_dst + 3 * i + lane));
    }
}

// dx[src[e], :] += dout[e, 0:C]   (edges are sorted by target, not by source: fp32 reductions at L2)
__global__ void __launch_bounds__(ED_WARPS * 32)
edge_message_bwd_kernel(const float *__restrict__ dout, const int64_t *__restrict__ src, int64_t E, int C,
                        float *__restrict__ dx) {
    const int lane = threadIdx.x & 31;
    const int W = C + 3;
    for (int64_t e = (int64_t)blockIdx.x * ED_WARPS + (threadIdx.x >> 5); e < E; e += (int64_t)gridDim.x * ED_WARPS) {
        const int64_t j = __ldg(src + e);
        for (int c = lane; c < C; c += 32) atomicAdd(dx + j * C + c, __ldg(dout + e * W + c));
    }
}

// out[e] = [pos_i (3), pos_j (3), pos_i - pos_j (3), |pos_i - pos_j| (1)], i = target, j = source
__global__ void __launch_bounds__(256)
edge_encoding_kernel(const float *__restrict__ pos_src, const float *__restrict__ pos_dst,
                     const int64_t *__restrict__ src, const int64_t *__restrict__ dst, int64_t E,
                     float *__restrict__ out) {
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < E; e += (int64_t)gridDim.x * 256) {
        const int64_t j = __ldg(src + e), i = __ldg(dst + e);
        const float ix = __ldg(pos_dst + 3 * i), iy = __ldg(pos_dst + 3 * i + 1), iz = __ldg(pos_dst + 3 * i + 2);
        const float jx = __ldg(pos_src + 3 * j), jy = __ldg(pos_src + 3 * j + 1), jz = __ldg(pos_src + 3 * j + 2);
        const float vx = __fsub_rn(ix, jx), vy = __fsub_rn(iy, jy), vz = __fsub_rn(iz, jz);
        float s = __fmul_rn(vx, vx);
        s = __fadd_rn(s, __fmul_rn(vy, vy));
        s = __fadd_rn(s, __fmul_rn(vz, vz));
        float *o = out + e * 10;
        o[0] = ix, o[1] = iy, o[2] = iz, o[3] = jx, o[4] = jy, o[5] = jz, o[6] = vx, o[7] = vy, o[8] = vz;
        o[9] = __fsqrt_rn(s);
    }
}

// MODE 0: max (argmax edge kept for the backward pass, empty segment -> 0 like torch_scatter), MODE 1: sum in edge order
template <int MODE>
__global__ void __launch_bounds__(ED_WARPS * 32)
segment_reduce_kernel(const float *__restrict__ msg, const int64_t *__restrict__ dst, int64_t E, int64_t T, int C,
                      float *__restrict__ out, int64_t *__restrict__ arg) {
    const int lane = threadIdx.x & 31;
    for (int64_t t = (int64_t)blockIdx.x * ED_WARPS + (threadIdx.x >> 5); t < T; t += (int64_t)gridDim.x * ED_WARPS) {
        const int64_t e0 = warp_lower_bound(dst, E, t, lane), e1 = warp_lower_bound(dst, E, t + 1, lane);
        for (int c0 = 0; c0 < C; c0 += 32) {
            const int c = c0 + lane;
            if (c >= C) continue;
            if (MODE == 0) {
                float best = 0.f;
                int64_t be = -1;
                for (int64_t e = e0; e < e1; ++e) {
                    const float v = __ldg(msg + e * C + c);
                    if (be < 0 || v > best) best = v, be = e;  // first maximum
                }
                out[t * C + c] = best;
                if (arg) arg[t * C + c] = be;
            } else {
                float acc = 0.f;
                for (int64_t e = e0; e < e1; ++e) acc = __fadd_rn(acc, __ldg(msg + e * C + c));
                out[t * C + c] = acc;
            }
        }
    }
}

// backward of the segment sum: dmsg[e, :] = dout[dst[e], :]
__global__ void __launch_bounds__(ED_WARPS * 32)
segment_sum_bwd_kernel(const float *__restrict__ dout, const int64_t *__restrict__ dst, int64_t E, int C,
                       float *__restrict__ dmsg) {
    const int lane = threadIdx.x & 31;
    for (int64_t e = (int64_t)blockIdx.x * ED_WARPS + (threadIdx.x >> 5); e < E; e += (int64_t)gridDim.x * ED_WARPS) {
        const int64_t t = __ldg(dst + e);
        for (int c = lane; c < C; c += 32) dmsg[e * C + c] = __ldg(dout + t * C + c);
    }
}

// backward of the segment max: dmsg (zeroed by the caller) [arg[t,c], c] = dout[t, c]
__global__ void __launch_bounds__(256)
segment_max_bwd_kernel(const float *__restrict__ dout, const int64_t *__restrict__ arg, int64_t total, int C,
                       float *__restrict__ dmsg) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        const int64_t e = __ldg(arg + i);
        if (e >= 0) dmsg[e * C + (i % C)] = __ldg(dout + i);
    }
}

static unsigned warp_grid(int64_t items) {
    int64_t g = (items + ED_WARPS - 1) / ED_WARPS;
    if (g > 32 * kNumSMs) g = 32 * kNumSMs;
    return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace p3d

using namespace p3d;

P3D_EXPORT int p3d_edge_message_f32(const float *x, const float *pos_src, const float *pos_dst, const int64_t *src,
                                    const int64_t *dst, int64_t E, int C, float *out, void *stream) {
    P3D_REQUIRE(E >= 0 && C >= 0, "edge_message: negative size");
    if (E == 0) return P3D_OK;
    P3D_REQUIRE(pos_src && pos_dst && src && dst && out && (C == 0 || x), "edge_message: null pointer");
    edge_message_kernel<<<warp_grid(E), ED_WARPS * 32, 0, (cudaStream_t)stream>>>(x, pos_src, pos_dst, src, dst, E, C, out);
    return finish_launch("edge_message");
}

P3D_EXPORT int p3d_edge_message_bwd_f32(const float *dout, const int64_t *src, int64_t E, int C, float *dx,
                                        void *stream) {
    P3D_REQUIRE(E >= 0 && C >= 0, "edge_message_bwd: negative size");
    if (E == 0 || C == 0) return P3D_OK;
    P3D_REQUIRE(dout && src && dx, "edge_message_bwd: null pointer");
    edge_message_bwd_kernel<<<warp_grid(E), ED_WARPS * 32, 0, (cudaStream_t)stream>>>(dout, src, E, C, dx);
    return finish_launch("edge_message_bwd");
}

P3D_EXPORT int p3d_edge_encoding_f32(const float *pos_src, const float *pos_dst, const int64_t *src, const int64_t *dst,
                                     int64_t E, float *out, void *stream) {
    P3D_REQUIRE(E >= 0, "edge_encoding: negative size");
    if (E == 0) return P3D_OK;
    P3D_REQUIRE(pos_src && pos_dst && src && dst && out, "edge_encoding: null pointer");
    int64_t g = (E + 255) / 256;
    if (g > 32 * kNumSMs) g = 32 * kNumSMs;
    edge_encoding_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(pos_src, pos_dst, src, dst, E, out);
    return finish_launch("edge_encoding");
}

P3D_EXPORT int p3d_segment_reduce_f32(const float *msg, const int64_t *dst, int64_t E, int64_t T, int C, int mode,
                                      float *out, int64_t *arg, void *stream) {
    P3D_REQUIRE(E >= 0 && T >= 0 && C >= 0, "segment_reduce: negative size");
    P3D_REQUIRE(mode == 0 || mode == 1, "segment_reduce: mode must be 0 (max) or 1 (sum)");
    if (T == 0 || C == 0) return P3D_OK;
    P3D_REQUIRE(out && (E == 0 || (msg && dst)), "segment_reduce: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) segment_reduce_kernel<0><<<warp_grid(T), ED_WARPS * 32, 0, st>>>(msg, dst, E, T, C, out, arg);
    else segment_reduce_kernel<1><<<warp_grid(T), ED_WARPS * 32, 0, st>>>(msg, dst, E, T, C, out, arg);
    return finish_launch("segment_reduce");
}

P3D_EXPORT int p3d_segment_reduce_bwd_f32(const float *dout, const int64_t *dst, const int64_t *arg, int64_t E,
                                          int64_t T, int C, int mode, float *dmsg, void *stream) {
    P3D_REQUIRE(E >= 0 && T >= 0 && C >= 0, "segment_reduce_bwd: negative size");
    P3D_REQUIRE(mode == 0 || mode == 1, "segment_reduce_bwd: mode must be 0 (max) or 1 (sum)");
    if (E == 0 || C == 0) return P3D_OK;
    P3D_REQUIRE(dout && dmsg, "segment_reduce_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (mode == 0) {
        P3D_REQUIRE(arg, "segment_reduce_bwd: max needs the argmax of the forward");
        cudaMemsetAsync(dmsg, 0, (size_t)E * C * sizeof(float), st);
        const int64_t total = T * C;
        int64_t g = (total + 255) / 256;
        if (g > 32 * kNumSMs) g = 32 * kNumSMs;
        if (total > 0) segment_max_bwd_kernel<<<(unsigned)g, 256, 0, st>>>(dout, arg, total, C, dmsg);
    } else {
        P3D_REQUIRE(dst, "segment_reduce_bwd: null pointer");
        segment_sum_bwd_kernel<<<warp_grid(E), ED_WARPS * 32, 0, st>>>(dout, dst, E, C, dmsg);
    }
    return finish_launch("segment_reduce_bwd");
}
