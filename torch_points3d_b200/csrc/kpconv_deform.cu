// kpconv_deform.cu -- deformable KPConv neighbour aggregation for sm_100a.
//
// Replaces KPConv_deform_ops (torch_points3d/modules/KPConv/convolution_ops.py:110-235), called from
// KPConvDeformableLayer.forward (modules/KPConv/kernels.py:228-241).  Same geometry as the rigid operator with
// three additions: every query point has its own kernel points (K_points + offsets[p]), neighbours that are in
// range of no kernel point are dropped (the reference's topk re-compaction, :158-185 -- it changes the result
// only for the gaussian influence, whose weights do not vanish at the extent), and the per-kernel-point weighted
// features are scaled by modulations[p,k] (:221-222).
//
// Forward: one warp per query point computes the [M x 16] influence table (fp32, the reference's operation
// order), optionally writes sq_distances [n,M,K] and/or the fused min over the neighbours that the fitting loss
// needs (losses.py:4-16: kpmin[p,k] = min_m d2, with its argmin for the backward pass), accumulates
// wf[p][k][:] = modulation * sum_m w[m][k] * x[nb[m]][:] and stores it; out = wf . K_values is a GEMM over
// that slab, and the slab is kept for the backward pass (dW = wf^T dout needs it).
// Backward: dwf = dout . K_values^T (GEMM), then one warp per query point recomputes the influence table and
// produces all four gradients: features (fp32 reductions at L2), offsets, modulations (and, through the
// GEMM, K_values).  d(weight)/d(d2) is nonzero only where the weight is: the inner products
// <dwf[k], x[nb[m]]> are evaluated for those (m, k) pairs only (~6 % of the table for the linear influence).
#include "kpconv_common.cuh"

namespace p3d {

int sgemm_strided(const float *A, int64_t sai, int64_t sak, const float *B, int64_t sbk, int64_t sbj, float *C,
                  int64_t ldc, int64_t Mr, int64_t Nc, int64_t Kd, cudaStream_t st);  // kpconv.cu
int dw_accumulate(const float *A, const float *B, int64_t n, int R, int Cn, float *dW, cudaStream_t st, float *part = nullptr,
                  size_t part_bytes = 0);  // kpconv_tc.cu
size_t dw_partial_bytes(int64_t n, int R, int Cn);
int sgemm_splitk_atomic(const float *A, const float *B, int64_t n, int64_t KC, int Cout, float *dW,
                        cudaStream_t st);  // kpconv.cu

struct DeformArgs {
    const float *q, *s, *x, *kp, *off, *mod;
    const int64_t *nb;
    int64_t n, N0;
    int M, Mpad, K, Cin;
    KpGeom g;
    float thr;  // extent^2: the in-range test of convolution_ops.py:156
};

// floats of shared memory per warp
__host__ __device__ inline int deform_warp_floats(int Mpad, bool backward) {
    return Mpad * KP_K * (backward ? 3 : 1) + 3 * Mpad + Mpad + (backward ? Mpad + KP_K * 32 : 0) + KP_K;
}

struct DeformWarp {
    float *wsm;   // [Mpad][16] influence weights (in-range mask applied, modulation not)
    float *d2sm;  // [Mpad][16] squared distances            (backward only)
    float *gsm;   // [Mpad][16] <dwf[k], x[nb[m]]>            (backward only)
    float *rel;   // [3][Mpad]  centred neighbour coordinates
    int *jsm;     // [Mpad]     feature row, -1 = shadow neighbour
    int *amask;   // [Mpad]     kernel points with a nonzero weight (backward only)
    float *dsm;   // [16][32]   dwf chunk                     (backward only)
    float *msm;   // [16]       modulations of this point
};

__device__ __forceinline__ DeformWarp deform_carve(float *base, int Mpad, bool backward) {
    DeformWarp w;
    w.wsm = base, base += Mpad * KP_K;
    w.d2sm = w.gsm = nullptr;
    if (backward) {
        w.d2sm = base, base += Mpad * KP_K;
        w.gsm = base, base += Mpad * KP_K;
    }
    w.rel = base, base += 3 * Mpad;
    w.jsm = reinterpret_cast<int *>(base), base += Mpad;
    w.amask = nullptr, w.dsm = nullptr;
    if (backward) {
        w.amask = reinterpret_cast<int *>(base), base += Mpad;
        w.dsm = base, base += KP_K * 32;
    }
    w.msm = base;
    return w;
}

// influence table of point p; lane owns kernel point (lane & 15) at (kx, ky, kz) = K_points[k] + offsets[p][k].
// best/bm: running (min d2, argmin m) of this lane's kernel point over the neighbours its half visits.
template <bool BACKWARD>
__device__ __forceinline__ void deform_point_table(const DeformArgs &a, int64_t p, const DeformWarp &sm, float kx,
                                                   float ky, float kz, int lane, float *__restrict__ sqd_out,
                                                   float &best, int &bm) {
    const float qx = __ldg(a.q + 3 * p), qy = __ldg(a.q + 3 * p + 1), qz = __ldg(a.q + 3 * p + 2);
    const int k = lane & 15, half = lane >> 4;
    const int64_t *nbp = a.nb + p * a.M;
    best = CUDART_INF_F, bm = 0x7fffffff;
    for (int m0 = 0; m0 < a.M; m0 += 32) {
        const int mm = m0 + lane;
        const int64_t j = mm < a.M ? __ldg(nbp + mm) : -1;
        const bool shadow = (j < 0 || j >= a.N0);
        float cx = 1e6f, cy = 1e6f, cz = 1e6f;
        if (!shadow) {
            cx = __ldg(a.s + 3 * j);
            cy = __ldg(a.s + 3 * j + 1);
            cz = __ldg(a.s + 3 * j + 2);
        }
        cx = __fsub_rn(cx, qx);
        cy = __fsub_rn(cy, qy);
        cz = __fsub_rn(cz, qz);
        sm.jsm[mm] = shadow ? -1 : (int)j;
        sm.rel[mm] = cx, sm.rel[a.Mpad + mm] = cy, sm.rel[2 * a.Mpad + mm] = cz;
#pragma unroll
        for (int t = 0; t < 16; ++t) {
            const int src = 2 * t + half;
            const int m = m0 + src;
            const float ax = __shfl_sync(0xffffffffu, cx, src);
            const float ay = __shfl_sync(0xffffffffu, cy, src);
            const float az = __shfl_sync(0xffffffffu, cz, src);
            const float d2 = sqdist3(ax, ay, az, kx, ky, kz);
            const bool real = k < a.K && m < a.M;
            const unsigned inr = __ballot_sync(0xffffffffu, real && d2 < a.thr);
            const bool any = ((inr >> (16 * half)) & 0xffffu) != 0u;  // neighbour m is in range of some kernel point
            float w = 0.f;
            if (real && any) {
                if (a.g.influence == P3D_INFLUENCE_CONSTANT) w = d2 < a.thr ? 1.f : 0.f;
                else w = kp_influence(d2, a.g);
            }
            sm.wsm[m * KP_K + k] = w;
            if (BACKWARD) sm.d2sm[m * KP_K + k] = d2;
            if (real) {
                if (sqd_out) sqd_out[(p * a.M + m) * a.K + k] = d2;
                if (d2 < best) best = d2, bm = m;  // m increases along t: the first minimum stays
            }
        }
    }
    // the two halves saw the even and the odd neighbours of the same kernel point
    const float ob = __shfl_xor_sync(0xffffffffu, best, 16);
    const int om = __shfl_xor_sync(0xffffffffu, bm, 16);
    if (ob < best || (ob == best && om < bm)) best = ob, bm = om;
    __syncwarp();
}

__global__ void __launch_bounds__(256)
kpconv_deform_wf_kernel(DeformArgs a, float *__restrict__ wf, float *__restrict__ sqd, float *__restrict__ kpmin,
                        int *__restrict__ argmin, int warps) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const DeformWarp sm = deform_carve(smem + (size_t)warp * deform_warp_floats(a.Mpad, false), a.Mpad, false);
    const int k = lane & 15;
    const float bx = k < a.K ? __ldg(a.kp + 3 * k) : 0.f;
    const float by = k < a.K ? __ldg(a.kp + 3 * k + 1) : 0.f;
    const float bz = k < a.K ? __ldg(a.kp + 3 * k + 2) : 0.f;
    for (int64_t p = (int64_t)blockIdx.x * warps + warp; p < a.n; p += (int64_t)gridDim.x * warps) {
        float kx = bx, ky = by, kz = bz;
        if (k < a.K) {
            const float *o = a.off + (p * a.K + k) * 3;
            kx = __fadd_rn(__ldg(o), bx), ky = __fadd_rn(__ldg(o + 1), by), kz = __fadd_rn(__ldg(o + 2), bz);
        }
        if (lane < KP_K) sm.msm[lane] = (a.mod && lane < a.K) ? __ldg(a.mod + p * a.K + lane) : 1.f;
        float best;
        int bm;
        deform_point_table<false>(a, p, sm, kx, ky, kz, lane, sqd, best, bm);
        if (lane < a.K) {
            if (kpmin) kpmin[p * a.K + lane] = best;
            if (argmin) argmin[p * a.K + lane] = bm;
        }
        for (int c0 = 0; c0 < a.Cin; c0 += 32) {
            const int c = c0 + lane;
            const bool cok = c < a.Cin;
            float acc[KP_K];
#pragma unroll
            for (int kk = 0; kk < KP_K; ++kk) acc[kk] = 0.f;
            for (int m = 0; m < a.M; ++m) {
                const int j = sm.jsm[m];
                if (j < 0) continue;  // warp-uniform
                const float xv = cok ? __ldg(a.x + (int64_t)j * a.Cin + c) : 0.f;
                const float4 *w4 = reinterpret_cast<const float4 *>(sm.wsm + m * KP_K);
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    const float4 w = w4[gq];
                    acc[4 * gq + 0] = fmaf(w.x, xv, acc[4 * gq + 0]);
                    acc[4 * gq + 1] = fmaf(w.y, xv, acc[4 * gq + 1]);
                    acc[4 * gq + 2] = fmaf(w.z, xv, acc[4 * gq + 2]);
                    acc[4 * gq + 3] = fmaf(w.w, xv, acc[4 * gq + 3]);
                }
            }
            if (cok) {
#pragma unroll
                for (int kk = 0; kk < KP_K; ++kk)
                    if (kk < a.K) wf[(p * a.K + kk) * a.Cin + c] = acc[kk] * sm.msm[kk];
            }
        }
        __syncwarp();
    }
}

// dwf = d(loss)/d(modulated wf) [n,K,Cin].  Produces dfeat (+=), doffsets [n,K,3], dmod [n,K].
// dsq (nullable) [n,M,K]: cotangent of sq_distances; dkpmin (nullable) [n,K] with argmin [n,K]: cotangent of the
// fused min over the neighbours.
__global__ void __launch_bounds__(256)
kpconv_deform_bwd_kernel(DeformArgs a, const float *__restrict__ dwf, const float *__restrict__ dsq,
                         const float *__restrict__ dkpmin, const int *__restrict__ argmin, float *__restrict__ dfeat,
                         float *__restrict__ doff, float *__restrict__ dmod, int warps) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const DeformWarp sm = deform_carve(smem + (size_t)warp * deform_warp_floats(a.Mpad, true), a.Mpad, true);
    const int k = lane & 15, half = lane >> 4;
    const float bx = k < a.K ? __ldg(a.kp + 3 * k) : 0.f;
    const float by = k < a.K ? __ldg(a.kp + 3 * k + 1) : 0.f;
    const float bz = k < a.K ? __ldg(a.kp + 3 * k + 2) : 0.f;
    for (int64_t p = (int64_t)blockIdx.x * warps + warp; p < a.n; p += (int64_t)gridDim.x * warps) {
        float kx = bx, ky = by, kz = bz;
        if (k < a.K) {
            const float *o = a.off + (p * a.K + k) * 3;
            kx = __fadd_rn(__ldg(o), bx), ky = __fadd_rn(__ldg(o + 1), by), kz = __fadd_rn(__ldg(o + 2), bz);
        }
        if (lane < KP_K) sm.msm[lane] = (a.mod && lane < a.K) ? __ldg(a.mod + p * a.K + lane) : 1.f;
        float best;
        int bm;
        deform_point_table<true>(a, p, sm, kx, ky, kz, lane, nullptr, best, bm);
        // kernel points with a nonzero weight per neighbour; clear the inner-product table
        for (int m = lane; m < a.Mpad; m += 32) {
            int bits = 0;
            if (m < a.M && sm.jsm[m] >= 0) {
#pragma unroll
                for (int kk = 0; kk < KP_K; ++kk) bits |= (sm.wsm[m * KP_K + kk] != 0.f) ? (1 << kk) : 0;
            }
            sm.amask[m] = bits;
        }
        for (int i = lane; i < a.Mpad * KP_K; i += 32) sm.gsm[i] = 0.f;
        __syncwarp();
        for (int c0 = 0; c0 < a.Cin; c0 += 32) {
            const int c = c0 + lane;
            const bool cok = c < a.Cin;
            float dm[KP_K];  // d(loss)/d(unmodulated wf[k][c])
#pragma unroll
            for (int kk = 0; kk < KP_K; ++kk) {
                const float d = (kk < a.K && cok) ? __ldg(dwf + (p * a.K + kk) * a.Cin + c) : 0.f;
                sm.dsm[kk * 32 + lane] = d;
                dm[kk] = d * sm.msm[kk];
            }
            __syncwarp();
            for (int m = 0; m < a.M; ++m) {
                const int j = sm.jsm[m];
                if (j < 0) continue;  // warp-uniform: shadow rows carry a zero feature and receive no gradient
                unsigned bits = (unsigned)sm.amask[m];
                if (bits == 0u) continue;
                const float4 *w4 = reinterpret_cast<const float4 *>(sm.wsm + m * KP_K);
                float v = 0.f;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    const float4 w = w4[gq];
                    v = fmaf(w.x, dm[4 * gq], v);
                    v = fmaf(w.y, dm[4 * gq + 1], v);
                    v = fmaf(w.z, dm[4 * gq + 2], v);
                    v = fmaf(w.w, dm[4 * gq + 3], v);
                }
                const float xv = cok ? __ldg(a.x + (int64_t)j * a.Cin + c) : 0.f;
                if (dfeat && cok) atomicAdd(dfeat + (int64_t)j * a.Cin + c, v);
                while (bits) {  // <dwf[kk], x[j]> for the kernel points that weigh this neighbour
                    const int kk = __ffs(bits) - 1;
                    bits &= bits - 1;
                    float part = sm.dsm[kk * 32 + lane] * xv;
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                    if (lane == 0) sm.gsm[m * KP_K + kk] += part;
                }
            }
            __syncwarp();
        }
        // gradients of the geometry: lanes = (half, kernel point) again
        float ox = 0.f, oy = 0.f, oz = 0.f, dmk = 0.f;
        const float modk = sm.msm[k];
        const float dkm = (dkpmin && k < a.K) ? __ldg(dkpmin + p * a.K + k) : 0.f;
        const int am = (argmin && k < a.K) ? __ldg(argmin + p * a.K + k) : -1;
        for (int m = half; m < a.M; m += 2) {
            if (k >= a.K) break;
            const float w = sm.wsm[m * KP_K + k];
            const float d2 = sm.d2sm[m * KP_K + k];
            const float G = sm.gsm[m * KP_K + k];
            float dd2 = 0.f;
            if (w != 0.f) {
                dmk = fmaf(w, G, dmk);
                float slope = 0.f;  // d(weight)/d(d2)
                if (a.g.influence == P3D_INFLUENCE_LINEAR) {
                    const float r = sqrtf(d2);
                    slope = r > 0.f ? -1.f / (2.f * a.g.extent * r) : 0.f;
                } else if (a.g.influence == P3D_INFLUENCE_GAUSSIAN) {
                    slope = -w / a.g.gauss_den;
                }
                dd2 = modk * G * slope;
            }
            if (dsq) dd2 += __ldg(dsq + (p * a.M + m) * a.K + k);
            if (m == am) dd2 += dkm;
            if (dd2 != 0.f) {  // d2 = |rel - kp|^2  ->  d(d2)/d(kp) = -2 (rel - kp)
                const float s2 = -2.f * dd2;
                ox = fmaf(s2, sm.rel[m] - kx, ox);
                oy = fmaf(s2, sm.rel[a.Mpad + m] - ky, oy);
                oz = fmaf(s2, sm.rel[2 * a.Mpad + m] - kz, oz);
            }
        }
        ox += __shfl_xor_sync(0xffffffffu, ox, 16);
        oy += __shfl_xor_sync(0xffffffffu, oy, 16);
        oz += __shfl_xor_sync(0xffffffffu, oz, 16);
        dmk += __shfl_xor_sync(0xffffffffu, dmk, 16);
        if (lane < a.K) {
            if (doff) {
                float *o = doff + (p * a.K + lane) * 3;
                o[0] = ox, o[1] = oy, o[2] = oz;
            }
            if (dmod) dmod[p * a.K + lane] = dmk;
        }
        __syncwarp();
    }
}

static KpGeom deform_geom(float extent, int influence) {
    KpGeom g;
    g.extent = extent;
    const float sig = extent * 0.3f;
    g.gauss_den = 2.0f * sig * sig + 1e-9f;
    g.influence = influence;
    g.aggregation = P3D_AGGREGATION_SUM;
    return g;
}

static int deform_check(const char *who, int64_t n, int64_t N0, int M, int K, int Cin, int Cout, int influence) {
    P3D_REQUIRE(n >= 0 && N0 >= 0, "%s: negative point count", who);
    P3D_REQUIRE(M >= 1 && M <= 128, "%s: neighbours per point must be in [1, 128], got %d", who, M);
    P3D_REQUIRE(K >= 1 && K <= KP_K, "%s: kernel points must be in [1, 16], got %d", who, K);
    P3D_REQUIRE(Cin >= 1 && Cout >= 1, "%s: channel counts must be positive", who);
    P3D_REQUIRE(N0 < (int64_t)1 << 31, "%s: more than 2^31 support points", who);
    P3D_REQUIRE(influence >= 0 && influence <= 2, "%s: unknown KP_influence %d", who, influence);
    return P3D_OK;
}

static DeformArgs deform_args(const float *q, const float *s, const int64_t *nb, const float *x, const float *kp,
                              const float *off, const float *mod, int64_t n, int64_t N0, int M, int K, int Cin,
                              float extent, int influence) {
    DeformArgs a;
    a.q = q, a.s = s, a.x = x, a.kp = kp, a.off = off, a.mod = mod, a.nb = nb;
    a.n = n, a.N0 = N0, a.M = M, a.Mpad = (M + 31) / 32 * 32, a.K = K, a.Cin = Cin;
    a.g = deform_geom(extent, influence);
    a.thr = (float)((double)extent * (double)extent);
    return a;
}

}  // namespace p3d

using namespace p3d;

P3D_EXPORT int p3d_kpconv_deform_fwd_f32(const float *query, const float *support, const int64_t *neighbors,
                                         const float *feat, const float *k_points, const float *offsets,
                                         const float *modulations, const float *k_values, int64_t n, int64_t N0,
                                         int M, int K, int Cin, int Cout, float extent, int influence, float *out,
                                         float *wf, float *sq_distances, float *kpmin, int32_t *argmin,
                                         void *stream) {
    if (int e = deform_check("kpconv_deform_fwd", n, N0, M, K, Cin, Cout, influence)) return e;
    if (n == 0) return P3D_OK;
    P3D_REQUIRE(query && neighbors && k_points && offsets && k_values && out && wf && (N0 == 0 || (support && feat)),
                "kpconv_deform_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const DeformArgs a = deform_args(query, support, neighbors, feat, k_points, offsets, modulations, n, N0, M, K, Cin,
                                     extent, influence);
    const int warps = 8;
    const size_t smem = (size_t)warps * deform_warp_floats(a.Mpad, false) * sizeof(float);
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(kpconv_deform_wf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(kpconv_deform_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done.mark();
    }
    int64_t grid = (n + warps - 1) / warps;
    if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
    kpconv_deform_wf_kernel<<<(unsigned)grid, warps * 32, smem, st>>>(a, wf, sq_distances, kpmin, argmin, warps);
    if (int e = finish_launch("kpconv_deform_wf")) return e;
    // out[p][o] = sum_{k,c} wf[p][k*Cin + c] * K_values[k*Cin + c][o]   (exact fp32)
    const int64_t KC = (int64_t)K * Cin;
    return sgemm_strided(wf, KC, 1, k_values, Cout, 1, out, Cout, n, Cout, KC, st);
}

P3D_EXPORT size_t p3d_kpconv_deform_bwd_ws_bytes(int64_t n, int K, int Cin) {
    return (size_t)n * (size_t)K * (size_t)Cin * sizeof(float);  // dwf
}

P3D_EXPORT int p3d_kpconv_deform_bwd_f32(const float *query, const float *support, const int64_t *neighbors,
                                         const float *feat, const float *k_points, const float *offsets,
                                         const float *modulations, const float *k_values, const float *wf,
                                         const float *dout, const float *d_sq_distances, const float *d_kpmin,
                                         const int32_t *argmin, int64_t n, int64_t N0, int M, int K, int Cin, int Cout,
                                         float extent, int influence, float *dfeat, float *dweights, float *doffsets,
                                         float *dmodulations, void *ws, size_t ws_bytes, void *stream) {
    if (int e = deform_check("kpconv_deform_bwd", n, N0, M, K, Cin, Cout, influence)) return e;
    if (n == 0) return P3D_OK;
    P3D_REQUIRE(query && neighbors && k_points && offsets && k_values && wf && dout && (N0 == 0 || (support && feat)),
                "kpconv_deform_bwd: null pointer");
    P3D_REQUIRE(d_kpmin == nullptr || argmin != nullptr, "kpconv_deform_bwd: d_kpmin needs the argmin of the forward");
    if (ws == nullptr || ws_bytes < p3d_kpconv_deform_bwd_ws_bytes(n, K, Cin)) {
        set_error("kpconv_deform_bwd: workspace of %zu bytes required", p3d_kpconv_deform_bwd_ws_bytes(n, K, Cin));
        return P3D_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t KC = (int64_t)K * Cin;
    if (dweights) {  // dW[kc][o] += sum_p wf[p][kc] * dout[p][o]
        if (dw_accumulate(wf, dout, n, (int)KC, Cout, dweights, st) != 0)
            if (int e = sgemm_splitk_atomic(wf, dout, n, KC, Cout, dweights, st)) return e;
    }
    float *dwf = (float *)ws;
    // dwf[p][kc] = sum_o dout[p][o] * K_values[kc][o]
    if (int e = sgemm_strided(dout, Cout, 1, k_values, 1, Cout, dwf, KC, n, KC, Cout, st)) return e;
    const DeformArgs a = deform_args(query, support, neighbors, feat, k_points, offsets, modulations, n, N0, M, K, Cin,
                                     extent, influence);
    const int warps = a.Mpad <= 64 ? 8 : 4;
    const size_t smem = (size_t)warps * deform_warp_floats(a.Mpad, true) * sizeof(float);
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(kpconv_deform_wf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(kpconv_deform_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done.mark();
    }
    int64_t grid = (n + warps - 1) / warps;
    if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
    kpconv_deform_bwd_kernel<<<(unsigned)grid, warps * 32, smem, st>>>(a, dwf, d_sq_distances, d_kpmin, argmin, dfeat,
                                                                     doffsets, dmodulations, warps);
    return finish_launch("kpconv_deform_bwd");
}
