// gather.cu -- grouping_operation, three_interpolate and the KPConv strided-shortcut
// neighbour max-pool for sm_100a: bandwidth-bound gather / scatter kernels.
//
// Replaces tp.grouping_operation (modules/pointnet2/dense.py:38,45), tp.three_interpolate
// (core/base_conv/dense.py:140) and the [n,M,C] gather + max of modules/KPConv/blocks.py:206-210.
//
// Design of the gathers: the output is >100x larger than the gathered table (C2 SA2: 339 MB
// out of a 10 MB feature tensor), so the only HBM stream is the write.  A CTA stages a block
// of CC channel rows of one cloud in shared memory (coalesced float4 fill), reads each int64
// index once into registers (reused for all CC channels), gathers from shared memory and
// emits 16-byte streaming stores (st.global.cs) that are contiguous across the warp.  No
// [B,C,S*K] int64 index tensor is materialised (the upstream formulation's `.repeat`).
#include <math_constants.h>

#include "common.cuh"

namespace p3d {

constexpr int GR_THREADS = 256;
constexpr int GR_VEC = 4;                        // output positions per thread per step
constexpr int GR_SMEM_FLOATS = 24 * 1024;        // 96 KB of staged feature rows per CTA (2 CTAs / SM)

// out[b, c_off+c, e] = feat[b, c, idx[b, e]],   e in [0, E = S*K)
template <bool STAGED>
__global__ void __launch_bounds__(GR_THREADS)
group_fwd_kernel(const float *__restrict__ feat, const int64_t *__restrict__ idx, int C, int N, int64_t E,
                 int CC, int e_per_cta, float *__restrict__ out, int64_t out_batch_stride, int c_off) {
    extern __shared__ __align__(16) float rows[];  // [CC][N] when STAGED
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * CC;
    const int cc = (C - c0 < CC) ? (C - c0) : CC;
    const int64_t e0 = (int64_t)blockIdx.x * e_per_cta;
    const int64_t e1 = (e0 + e_per_cta < E) ? e0 + e_per_cta : E;
    const float *__restrict__ fb = feat + ((int64_t)b * C + c0) * N;
    const int64_t *__restrict__ ib = idx + (int64_t)b * E;
    float *__restrict__ ob = out + (int64_t)b * out_batch_stride + (int64_t)(c_off + c0) * E;

    if (STAGED) {
        const int total = cc * N;
        if ((((uintptr_t)fb) & 15) == 0 && (total & 3) == 0) {
            const float4 *__restrict__ s4 = reinterpret_cast<const float4 *>(fb);
            float4 *d4 = reinterpret_cast<float4 *>(rows);
            for (int i = threadIdx.x; i < total / 4; i += GR_THREADS) d4[i] = __ldg(s4 + i);
        } else {
            for (int i = threadIdx.x; i < total; i += GR_THREADS) rows[i] = __ldg(fb + i);
        }
        __syncthreads();
    }
    const bool vec_ok = ((E & 3) == 0) && ((((uintptr_t)ob) & 15) == 0) && ((((uintptr_t)ib) & 15) == 0);
    for (int64_t e = e0 + (int64_t)threadIdx.x * GR_VEC; e < e1; e += (int64_t)GR_THREADS * GR_VEC) {
        int j[GR_VEC];
        if (vec_ok && e + GR_VEC <= e1) {
            const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(ib + e));
            const longlong2 c = __ldg(reinterpret_cast<const longlong2 *>(ib + e + 2));
            j[0] = (int)a.x; j[1] = (int)a.y; j[2] = (int)c.x; j[3] = (int)c.y;
#pragma unroll 4
            for (int ch = 0; ch < cc; ++ch) {
                const float *__restrict__ r = STAGED ? rows + ch * N : fb + (int64_t)ch * N;
                float4 v;
                if (STAGED) { v.x = r[j[0]]; v.y = r[j[1]]; v.z = r[j[2]]; v.w = r[j[3]]; }
                else { v.x = __ldg(r + j[0]); v.y = __ldg(r + j[1]); v.z = __ldg(r + j[2]); v.w = __ldg(r + j[3]); }
                st_cs_f4(ob + (int64_t)ch * E + e, v);
            }
        } else {
            for (int t = 0; t < GR_VEC && e + t < e1; ++t) {
                const int jj = (int)__ldg(ib + e + t);
                for (int ch = 0; ch < cc; ++ch) {
                    const float *__restrict__ r = STAGED ? rows + ch * N : fb + (int64_t)ch * N;
                    ob[(int64_t)ch * E + e + t] = STAGED ? r[jj] : __ldg(r + jj);
                }
            }
        }
    }
}

// dfeat[b, c, idx[b,e]] += dout[b, c_off+c, e]   (fp32 RED to L2; dfeat zeroed by the caller)
__global__ void __launch_bounds__(GR_THREADS)
group_bwd_kernel(const float *__restrict__ dout, const int64_t *__restrict__ idx, int C, int N, int64_t E,
                 int CC, int e_per_cta, float *__restrict__ dfeat, int64_t dout_batch_stride, int c_off) {
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * CC;
    const int cc = (C - c0 < CC) ? (C - c0) : CC;
    const int64_t e0 = (int64_t)blockIdx.x * e_per_cta;
    const int64_t e1 = (e0 + e_per_cta < E) ? e0 + e_per_cta : E;
    float *__restrict__ db = dfeat + ((int64_t)b * C + c0) * N;
    const int64_t *__restrict__ ib = idx + (int64_t)b * E;
    const float *__restrict__ gb = dout + (int64_t)b * dout_batch_stride + (int64_t)(c_off + c0) * E;
    const bool vec_ok = ((E & 3) == 0) && ((((uintptr_t)gb) & 15) == 0) && ((((uintptr_t)ib) & 15) == 0);
    for (int64_t e = e0 + (int64_t)threadIdx.x * GR_VEC; e < e1; e += (int64_t)GR_THREADS * GR_VEC) {
        if (vec_ok && e + GR_VEC <= e1) {
            const longlong2 a = __ldg(reinterpret_cast<const longlong2 *>(ib + e));
            const longlong2 c = __ldg(reinterpret_cast<const longlong2 *>(ib + e + 2));
            const int j0 = (int)a.x, j1 = (int)a.y, j2 = (int)c.x, j3 = (int)c.y;
#pragma unroll 4
            for (int ch = 0; ch < cc; ++ch) {
                const float4 g = __ldcs(reinterpret_cast<const float4 *>(gb + (int64_t)ch * E + e));
                float *r = db + (int64_t)ch * N;
                atomicAdd(r + j0, g.x);
                atomicAdd(r + j1, g.y);
                atomicAdd(r + j2, g.z);
                atomicAdd(r + j3, g.w);
            }
        } else {
            for (int t = 0; t < GR_VEC && e + t < e1; ++t) {
                const int jj = (int)__ldg(ib + e + t);
                for (int ch = 0; ch < cc; ++ch) atomicAdd(db + (int64_t)ch * N + jj, gb[(int64_t)ch * E + e + t]);
            }
        }
    }
}

// out[b, c, s, k] = (pos[b, idx[b,s,k], c] - new_pos[b, s, c]) / norm,  c = 0..2
__global__ void __launch_bounds__(GR_THREADS)
group_xyz_kernel(const float *__restrict__ pos, const float *__restrict__ new_pos, const int64_t *__restrict__ idx,
                 int N, int S, int K, float norm, int do_norm, float *__restrict__ out, int64_t out_batch_stride) {
    const int b = blockIdx.y;
    const int64_t E = (int64_t)S * K;
    const float *__restrict__ pb = pos + (int64_t)b * N * 3;
    const float *__restrict__ nb = new_pos + (int64_t)b * S * 3;
    const int64_t *__restrict__ ib = idx + (int64_t)b * E;
    float *__restrict__ ob = out + (int64_t)b * out_batch_stride;
    for (int64_t e = (int64_t)blockIdx.x * GR_THREADS + threadIdx.x; e < E; e += (int64_t)gridDim.x * GR_THREADS) {
        const int j = (int)__ldg(ib + e);
        const int s = (int)(e / K);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __fsub_rn(__ldg(pb + 3 * j + c), __ldg(nb + 3 * s + c));
            if (do_norm) v = __fdiv_rn(v, norm);
            __stcs(ob + (int64_t)c * E + e, v);
        }
    }
}

// out[b,c,j] = ((f[i0]*w0) + f[i1]*w1) + f[i2]*w2   -- staged feature rows, 4 outputs / thread
template <bool STAGED>
__global__ void __launch_bounds__(GR_THREADS)
three_interp_fwd_kernel(const float *__restrict__ feat, const int64_t *__restrict__ idx,
                        const float *__restrict__ weight, int C, int m, int n, int CC, int n_per_cta,
                        float *__restrict__ out) {
    extern __shared__ __align__(16) float rows[];
    const int b = blockIdx.z;
    const int c0 = blockIdx.y * CC;
    const int cc = (C - c0 < CC) ? (C - c0) : CC;
    const int j0 = blockIdx.x * n_per_cta;
    const int j1 = (j0 + n_per_cta < n) ? j0 + n_per_cta : n;
    const float *__restrict__ fb = feat + ((int64_t)b * C + c0) * m;
    if (STAGED) {
        for (int i = threadIdx.x; i < cc * m; i += GR_THREADS) rows[i] = __ldg(fb + i);
        __syncthreads();
    }
    const int64_t *__restrict__ ib = idx + (int64_t)b * n * 3;
    const float *__restrict__ wb = weight + (int64_t)b * n * 3;
    float *__restrict__ ob = out + ((int64_t)b * C + c0) * n;
    for (int j = j0 + threadIdx.x; j < j1; j += GR_THREADS) {
        const int i0 = (int)__ldg(ib + 3 * j), i1 = (int)__ldg(ib + 3 * j + 1), i2 = (int)__ldg(ib + 3 * j + 2);
        const float w0 = __ldg(wb + 3 * j), w1 = __ldg(wb + 3 * j + 1), w2 = __ldg(wb + 3 * j + 2);
        for (int ch = 0; ch < cc; ++ch) {
            const float *__restrict__ r = STAGED ? rows + ch * m : fb + (int64_t)ch * m;
            const float f0 = STAGED ? r[i0] : __ldg(r + i0);
            const float f1 = STAGED ? r[i1] : __ldg(r + i1);
            const float f2 = STAGED ? r[i2] : __ldg(r + i2);
            float s = __fmul_rn(f0, w0);
            s = __fadd_rn(s, __fmul_rn(f1, w1));
            s = __fadd_rn(s, __fmul_rn(f2, w2));
            ob[(int64_t)ch * n + j] = s;
        }
    }
}

__global__ void __launch_bounds__(GR_THREADS)
three_interp_bwd_kernel(const float *__restrict__ dout, const int64_t *__restrict__ idx,
                        const float *__restrict__ weight, int C, int m, int n, float *__restrict__ dfeat) {
    const int b = blockIdx.z;
    const int c = blockIdx.y;
    const int64_t *__restrict__ ib = idx + (int64_t)b * n * 3;
    const float *__restrict__ wb = weight + (int64_t)b * n * 3;
    const float *__restrict__ gb = dout + ((int64_t)b * C + c) * n;
    float *__restrict__ db = dfeat + ((int64_t)b * C + c) * m;
    for (int j = blockIdx.x * GR_THREADS + threadIdx.x; j < n; j += gridDim.x * GR_THREADS) {
        const float g = gb[j];
#pragma unroll
        for (int t = 0; t < 3; ++t) atomicAdd(db + (int)__ldg(ib + 3 * j + t), g * __ldg(wb + 3 * j + t));
    }
}

// out[p,c] = max_m x'[nb[p,m], c];  one warp per query row chunk, lanes over channels
__global__ void __launch_bounds__(GR_THREADS)
neighbor_max_kernel(const float *__restrict__ x, const int64_t *__restrict__ nb, int64_t n, int64_t N0, int M,
                    int C, float *__restrict__ out, int64_t *__restrict__ arg) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * (GR_THREADS / 32) + (threadIdx.x >> 5);
    if (p >= n) return;
    for (int c = lane; c < C; c += 32) {
        float best = -CUDART_INF_F;
        int64_t bi = N0;
        for (int m = 0; m < M; ++m) {
            int64_t j = __ldg(nb + p * M + m);
            const bool shadow = (j < 0 || j >= N0);
            const float v = shadow ? 0.f : __ldg(x + j * C + c);
            if (v > best) { best = v; bi = shadow ? N0 : j; }
        }
        out[p * C + c] = best;
        if (arg) arg[p * C + c] = bi;
    }
}

static int pick_cc(int C, int N, bool &staged) {
    // channels staged per CTA: as many as fit in GR_SMEM_FLOATS, at most 32
    int cc = N > 0 ? GR_SMEM_FLOATS / N : 0;
    if (cc > 32) cc = 32;
    if (cc > C) cc = C;
    staged = cc >= 1;
    if (!staged) cc = C < 8 ? C : 8;
    return cc;
}

}  // namespace p3d

using namespace p3d;

P3D_EXPORT int p3d_group_fwd_f32(const float *feat, const int64_t *idx, int B, int C, int N, int S, int K,
                                 float *out, int64_t out_batch_stride, int c_off, void *stream) {
    P3D_REQUIRE(B >= 0 && C >= 0 && N >= 0 && S >= 0 && K >= 0, "group_fwd: negative dimension");
    const int64_t E = (int64_t)S * K;
    if (B == 0 || C == 0 || E == 0) return P3D_OK;
    P3D_REQUIRE(feat && idx && out, "group_fwd: null pointer");
    P3D_REQUIRE(N > 0, "group_fwd: empty feature table");
    bool staged;
    const int CC = pick_cc(C, N, staged);
    const int cchunks = (C + CC - 1) / CC;
    // positions per CTA: amortise the staged fill over >= 8x its size, but keep >= 2 waves of CTAs
    int64_t e_per_cta = (int64_t)GR_THREADS * GR_VEC * 8;
    while (e_per_cta > GR_THREADS * GR_VEC && ((E + e_per_cta - 1) / e_per_cta) * cchunks * B < 4 * kNumSMs)
        e_per_cta /= 2;
    dim3 grid((unsigned)((E + e_per_cta - 1) / e_per_cta), (unsigned)cchunks, (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    if (staged) {
        const size_t smem = (size_t)CC * N * sizeof(float);
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(group_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 GR_SMEM_FLOATS * (int)sizeof(float));
            attr_done = true;
        }
        group_fwd_kernel<true><<<grid, GR_THREADS, smem, st>>>(feat, idx, C, N, E, CC, (int)e_per_cta, out,
                                                               out_batch_stride, c_off);
    } else {
        group_fwd_kernel<false><<<grid, GR_THREADS, 0, st>>>(feat, idx, C, N, E, CC, (int)e_per_cta, out,
                                                             out_batch_stride, c_off);
    }
    return finish_launch("group_fwd");
}

P3D_EXPORT int p3d_group_bwd_f32(const float *dout, const int64_t *idx, int B, int C, int N, int S, int K,
                                 float *dfeat, int64_t dout_batch_stride, int c_off, void *stream) {
    P3D_REQUIRE(B >= 0 && C >= 0 && N >= 0 && S >= 0 && K >= 0, "group_bwd: negative dimension");
    const int64_t E = (int64_t)S * K;
    if (B == 0 || C == 0 || E == 0) return P3D_OK;
    P3D_REQUIRE(dout && idx && dfeat, "group_bwd: null pointer");
    const int CC = C < 16 ? C : 16;
    const int cchunks = (C + CC - 1) / CC;
    int64_t e_per_cta = (int64_t)GR_THREADS * GR_VEC * 4;
    while (e_per_cta > GR_THREADS * GR_VEC && ((E + e_per_cta - 1) / e_per_cta) * cchunks * B < 4 * kNumSMs)
        e_per_cta /= 2;
    dim3 grid((unsigned)((E + e_per_cta - 1) / e_per_cta), (unsigned)cchunks, (unsigned)B);
    group_bwd_kernel<<<grid, GR_THREADS, 0, (cudaStream_t)stream>>>(dout, idx, C, N, E, CC, (int)e_per_cta, dfeat,
                                                                   dout_batch_stride, c_off);
    return finish_launch("group_bwd");
}

P3D_EXPORT int p3d_group_xyz_f32(const float *pos, const float *new_pos, const int64_t *idx, int B, int N, int S,
                                 int K, float norm, float *out, int64_t out_batch_stride, void *stream) {
    P3D_REQUIRE(B >= 0 && N >= 0 && S >= 0 && K >= 0, "group_xyz: negative dimension");
    const int64_t E = (int64_t)S * K;
    if (B == 0 || E == 0) return P3D_OK;
    P3D_REQUIRE(pos && new_pos && idx && out, "group_xyz: null pointer");
    int64_t gx = (E + GR_THREADS - 1) / GR_THREADS;
    if (gx > 4 * kNumSMs) gx = 4 * kNumSMs;
    dim3 grid((unsigned)gx, (unsigned)B);
    group_xyz_kernel<<<grid, GR_THREADS, 0, (cudaStream_t)stream>>>(pos, new_pos, idx, N, S, K, norm,
                                                                   norm != 1.0f ? 1 : 0, out, out_batch_stride);
    return finish_launch("group_xyz");
}

P3D_EXPORT int p3d_three_interp_fwd_f32(const float *feat, const int64_t *idx, const float *weight, int B, int C,
                                        int m, int n, float *out, void *stream) {
    P3D_REQUIRE(B >= 0 && C >= 0 && m >= 0 && n >= 0, "three_interpolate: negative dimension");
    if (B == 0 || C == 0 || n == 0) return P3D_OK;
    P3D_REQUIRE(feat && idx && weight && out, "three_interpolate: null pointer");
    P3D_REQUIRE(m > 0, "three_interpolate: empty feature table");
    bool staged;
    const int CC = pick_cc(C, m, staged);
    const int cchunks = (C + CC - 1) / CC;
    int n_per_cta = GR_THREADS * 8;
    while (n_per_cta > GR_THREADS && (int64_t)((n + n_per_cta - 1) / n_per_cta) * cchunks * B < 2 * kNumSMs)
        n_per_cta /= 2;
    dim3 grid((unsigned)((n + n_per_cta - 1) / n_per_cta), (unsigned)cchunks, (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    if (staged) {
        static bool attr_done = false;
        if (!attr_done) {
            cudaFuncSetAttribute(three_interp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 GR_SMEM_FLOATS * (int)sizeof(float));
            attr_done = true;
        }
        three_interp_fwd_kernel<true><<<grid, GR_THREADS, (size_t)CC * m * sizeof(float), st>>>(
            feat, idx, weight, C, m, n, CC, n_per_cta, out);
    } else {
        three_interp_fwd_kernel<false><<<grid, GR_THREADS, 0, st>>>(feat, idx, weight, C, m, n, CC, n_per_cta, out);
    }
    return finish_launch("three_interp_fwd");
}

P3D_EXPORT int p3d_three_interp_bwd_f32(const float *dout, const int64_t *idx, const float *weight, int B, int C,
                                        int m, int n, float *dfeat, void *stream) {
    P3D_REQUIRE(B >= 0 && C >= 0 && m >= 0 && n >= 0, "three_interpolate_bwd: negative dimension");
    if (B == 0 || C == 0 || n == 0) return P3D_OK;
    P3D_REQUIRE(dout && idx && weight && dfeat, "three_interpolate_bwd: null pointer");
    P3D_REQUIRE(C <= 65535 && B <= 65535, "three_interpolate_bwd: C and B must be <= 65535");
    int gx = (n + GR_THREADS - 1) / GR_THREADS;
    dim3 grid((unsigned)gx, (unsigned)C, (unsigned)B);
    three_interp_bwd_kernel<<<grid, GR_THREADS, 0, (cudaStream_t)stream>>>(dout, idx, weight, C, m, n, dfeat);
    return finish_launch("three_interp_bwd");
}

P3D_EXPORT int p3d_neighbor_max_f32(const float *x, const int64_t *neighbors, int64_t n, int64_t N0, int M, int C,
                                    float *out, int64_t *arg, void *stream) {
    P3D_REQUIRE(n >= 0 && N0 >= 0 && M >= 1 && C >= 0, "neighbor_max: bad dimension");
    if (n == 0 || C == 0) return P3D_OK;
    P3D_REQUIRE(neighbors && out && (x || N0 == 0), "neighbor_max: null pointer");
    const int64_t grid = (n + GR_THREADS / 32 - 1) / (GR_THREADS / 32);
    neighbor_max_kernel<<<(unsigned)grid, GR_THREADS, 0, (cudaStream_t)stream>>>(x, neighbors, n, N0, M, C, out, arg);
    return finish_launch("neighbor_max");
}
