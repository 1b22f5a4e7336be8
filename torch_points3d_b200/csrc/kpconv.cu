// kpconv.cu -- rigid KPConv neighbour aggregation (fp32 SIMT path) for sm_100a.
//
// Replaces KPConv_ops (torch_points3d/modules/KPConv/convolution_ops.py:19-107), called from
// KPConvLayer.forward (modules/KPConv/kernels.py:83-93).
//
// Forward, one fused kernel: a warp owns T consecutive query points.  Per point it computes the
// [M x 16] kernel-point influence table in shared memory (kpconv_common.cuh), gathers the
// neighbour feature rows straight from L2 (lanes over channels -> one 128-byte line per row
// and 32-channel slab) and accumulates wf[k][c] = sum_m w[k][m] * x[nb[m]][c] in 16
// registers; the [16 x 32 x T] slab of weighted features stays in shared memory and is
// contracted against the kernel weights W[k][c][:] with lanes over output channels.
// None of the reference's intermediates ([n,M,K,3] differences, [n,M,K] weights,
// [n,M,Cin] gathered features, [n,K,Cin] weighted features, [K,n,Cout] kernel outputs)
// exists in HBM.  This path is exact fp32 (1e-5 parity against the oracle); the tcgen05
// path for the kernel-point contraction lives in kpconv_tc.cu.
//
// Backward: wf and d(wf) go through a caller-provided workspace and two plain tiled
// SGEMMs (dW = wf^T dout, dwf = dout W^T), then one scatter kernel recomputes the influence
// weights and accumulates dfeat with fp32 reductions at L2.
#include "kpconv_common.cuh"

namespace p3d {

constexpr int KPF_WARPS = 8;
constexpr int KPF_T = 4;  // query points per warp per pass

// wf[k] (16 accumulators per lane, lane = channel c0+lane) of one point
__device__ __forceinline__ void kp_weighted_features(const float *__restrict__ x, int Cin, int c0, int M,
                                                     const float *wsm, const int *jsm, int lane, float (&acc)[KP_K]) {
#pragma unroll
    for (int k = 0; k < KP_K; ++k) acc[k] = 0.f;
    const int c = c0 + lane;
    const bool cok = c < Cin;
    for (int m = 0; m < M; ++m) {
        const int j = jsm[m];
        if (j < 0) continue;  // warp-uniform
        const float xv = cok ? __ldg(x + (int64_t)j * Cin + c) : 0.f;
        const float4 *w4 = reinterpret_cast<const float4 *>(wsm + m * KP_K);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 w = w4[g];
            acc[4 * g + 0] = fmaf(w.x, xv, acc[4 * g + 0]);
            acc[4 * g + 1] = fmaf(w.y, xv, acc[4 * g + 1]);
            acc[4 * g + 2] = fmaf(w.z, xv, acc[4 * g + 2]);
            acc[4 * g + 3] = fmaf(w.w, xv, acc[4 * g + 3]);
        }
    }
}

__global__ void __launch_bounds__(KPF_WARPS * 32)
kpconv_fwd_simt_kernel(const float *__restrict__ q, const float *__restrict__ s, const int64_t *__restrict__ nb,
                       const float *__restrict__ x, const float *__restrict__ kp, const float *__restrict__ W,
                       int64_t n, int64_t N0, int M, int Mpad, int K, int Cin, int Cout, KpGeom g,
                       float *__restrict__ out) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int per_warp = Mpad * KP_K + Mpad + KP_K * 32 * KPF_T;
    float *wsm = smem + (size_t)warp * per_warp;
    int *jsm = reinterpret_cast<int *>(wsm + Mpad * KP_K);
    float *wfs = wsm + Mpad * KP_K + Mpad;  // [16][T][32]

    const int kk = lane & 15;
    const float kx = kk < K ? __ldg(kp + 3 * kk) : 0.f;
    const float ky = kk < K ? __ldg(kp + 3 * kk + 1) : 0.f;
    const float kz = kk < K ? __ldg(kp + 3 * kk + 2) : 0.f;

    const int64_t tiles = (n + KPF_T - 1) / KPF_T;
    for (int64_t tile = (int64_t)blockIdx.x * KPF_WARPS + warp; tile < tiles; tile += (int64_t)gridDim.x * KPF_WARPS) {
        const int64_t p0 = tile * KPF_T;
        for (int c0 = 0; c0 < Cin; c0 += 32) {
            for (int t = 0; t < KPF_T; ++t) {
                const int64_t p = p0 + t;
                float acc[KP_K];
                if (p < n) {
                    kp_point_weights(q + 3 * p, s, nb + p * M, N0, M, K, g, kx, ky, kz, wsm, jsm, lane);
                    kp_weighted_features(x, Cin, c0, M, wsm, jsm, lane, acc);
                } else {
#pragma unroll
                    for (int k = 0; k < KP_K; ++k) acc[k] = 0.f;
                }
#pragma unroll
                for (int k = 0; k < KP_K; ++k) wfs[(k * KPF_T + t) * 32 + lane] = acc[k];
                __syncwarp();
            }
            const int cw = (Cin - c0 < 32) ? (Cin - c0) : 32;
            for (int o0 = 0; o0 < Cout; o0 += 32) {
                const int o = o0 + lane;
                const bool ook = o < Cout;
                float a[KPF_T];
#pragma unroll
                for (int t = 0; t < KPF_T; ++t) a[t] = 0.f;
                for (int k = 0; k < K; ++k) {
                    const float *__restrict__ wrow = W + ((int64_t)k * Cin + c0) * Cout + o;
                    const float *wk = wfs + k * KPF_T * 32;
#pragma unroll 4
                    for (int c = 0; c < cw; ++c) {
                        const float wv = ook ? __ldg(wrow + (int64_t)c * Cout) : 0.f;
#pragma unroll
                        for (int t = 0; t < KPF_T; ++t) a[t] = fmaf(wk[t * 32 + c], wv, a[t]);
                    }
                }
                if (ook) {
#pragma unroll
                    for (int t = 0; t < KPF_T; ++t) {
                        const int64_t p = p0 + t;
                        if (p < n) {
                            float *dst = out + p * Cout + o;
                            *dst = (c0 == 0) ? a[t] : (*dst + a[t]);  // rows are warp-private
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
}

// wf[p][k][c] (K real kernel points) -> workspace, for the backward pass
__global__ void __launch_bounds__(KPF_WARPS * 32)
kpconv_wf_kernel(const float *__restrict__ q, const float *__restrict__ s, const int64_t *__restrict__ nb,
                 const float *__restrict__ x, const float *__restrict__ kp, int64_t n, int64_t N0, int M, int Mpad,
                 int K, int Cin, KpGeom g, float *__restrict__ wf) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *wsm = smem + (size_t)warp * (Mpad * KP_K + Mpad);
    int *jsm = reinterpret_cast<int *>(wsm + Mpad * KP_K);
    const int kk = lane & 15;
    const float kx = kk < K ? __ldg(kp + 3 * kk) : 0.f;
    const float ky = kk < K ? __ldg(kp + 3 * kk + 1) : 0.f;
    const float kz = kk < K ? __ldg(kp + 3 * kk + 2) : 0.f;
    for (int64_t p = (int64_t)blockIdx.x * KPF_WARPS + warp; p < n; p += (int64_t)gridDim.x * KPF_WARPS) {
        kp_point_weights(q + 3 * p, s, nb + p * M, N0, M, K, g, kx, ky, kz, wsm, jsm, lane);
        for (int c0 = 0; c0 < Cin; c0 += 32) {
            float acc[KP_K];
            kp_weighted_features(x, Cin, c0, M, wsm, jsm, lane, acc);
            if (c0 + lane < Cin) {
#pragma unroll
                for (int k = 0; k < KP_K; ++k)
                    if (k < K) wf[(p * K + k) * Cin + c0 + lane] = acc[k];
            }
        }
        __syncwarp();
    }
}

// dfeat[nb[p,m]][c] += sum_k w[p,k,m] * dwf[p][k][c]
__global__ void __launch_bounds__(KPF_WARPS * 32)
kpconv_scatter_kernel(const float *__restrict__ q, const float *__restrict__ s, const int64_t *__restrict__ nb,
                      const float *__restrict__ kp, const float *__restrict__ dwf, int64_t n, int64_t N0, int M,
                      int Mpad, int K, int Cin, KpGeom g, float *__restrict__ dfeat) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *wsm = smem + (size_t)warp * (Mpad * KP_K + Mpad);
    int *jsm = reinterpret_cast<int *>(wsm + Mpad * KP_K);
    const int kk = lane & 15;
    const float kx = kk < K ? __ldg(kp + 3 * kk) : 0.f;
    const float ky = kk < K ? __ldg(kp + 3 * kk + 1) : 0.f;
    const float kz = kk < K ? __ldg(kp + 3 * kk + 2) : 0.f;
    for (int64_t p = (int64_t)blockIdx.x * KPF_WARPS + warp; p < n; p += (int64_t)gridDim.x * KPF_WARPS) {
        kp_point_weights(q + 3 * p, s, nb + p * M, N0, M, K, g, kx, ky, kz, wsm, jsm, lane);
        for (int c0 = 0; c0 < Cin; c0 += 32) {
            const int c = c0 + lane;
            float d[KP_K];
#pragma unroll
            for (int k = 0; k < KP_K; ++k) d[k] = (k < K && c < Cin) ? __ldg(dwf + (p * K + k) * Cin + c) : 0.f;
            for (int m = 0; m < M; ++m) {
                const int j = jsm[m];
                if (j < 0) continue;
                const float4 *w4 = reinterpret_cast<const float4 *>(wsm + m * KP_K);
                float v = 0.f;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    const float4 w = w4[gq];
                    v = fmaf(w.x, d[4 * gq], v);
                    v = fmaf(w.y, d[4 * gq + 1], v);
                    v = fmaf(w.z, d[4 * gq + 2], v);
                    v = fmaf(w.w, d[4 * gq + 3], v);
                }
                if (c < Cin) atomicAdd(dfeat + (int64_t)j * Cin + c, v);
            }
        }
        __syncwarp();
    }
}

// C[i][j] (+)= sum_k A(i,k) * B(k,j) with arbitrary element strides; 64x64x16 tiles, 4x4 per
// thread; gridDim.z splits k (partial sums are reduced with fp32 atomics into a zeroed C).
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float *__restrict__ A, int64_t sai, int64_t sak, const float *__restrict__ B, int64_t sbk,
                     int64_t sbj, float *__restrict__ C, int64_t ldc, int64_t Mr, int64_t Nc, int64_t Kd,
                     int64_t k_per_split, int atomic) {
    __shared__ float As[16][64 + 4], Bs[16][64 + 4];
    const int tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.y * 64, j0 = (int64_t)blockIdx.x * 64;
    const int64_t kb = (int64_t)blockIdx.z * k_per_split;
    const int64_t ke = (kb + k_per_split < Kd) ? kb + k_per_split : Kd;
    const int ty = tid / 16, tx = tid % 16;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int64_t k0 = kb; k0 < ke; k0 += 16) {
        for (int e = tid; e < 16 * 64; e += 256) {
            int kk_, ii;
            if (sak == 1) { kk_ = e % 16; ii = e / 16; } else { ii = e % 64; kk_ = e / 64; }
            const int64_t i = i0 + ii, k = k0 + kk_;
            As[kk_][ii] = (i < Mr && k < ke) ? __ldg(A + i * sai + k * sak) : 0.f;
            int kb_, jj;
            if (sbk == 1) { kb_ = e % 16; jj = e / 16; } else { jj = e % 64; kb_ = e / 64; }
            const int64_t j = j0 + jj, k2 = k0 + kb_;
            Bs[kb_][jj] = (j < Nc && k2 < ke) ? __ldg(B + k2 * sbk + j * sbj) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { a[u] = As[k][ty * 4 + u]; b[u] = Bs[k][tx * 4 + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int64_t i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
            if (i < Mr && j < Nc) {
                if (atomic) atomicAdd(C + i * ldc + j, acc[u][v]);
                else C[i * ldc + j] = acc[u][v];
            }
        }
}

static KpGeom make_geom(float extent, int influence, int aggregation) {
    KpGeom g;
    g.extent = extent;
    const float sig = extent * 0.3f;
    g.gauss_den = 2.0f * sig * sig + 1e-9f;
    g.influence = influence;
    g.aggregation = aggregation;
    return g;
}

static int check_common(const char *who, int64_t n, int64_t N0, int M, int K, int Cin, int Cout, int influence,
                        int aggregation) {
    P3D_REQUIRE(n >= 0 && N0 >= 0, "%s: negative point count", who);
    P3D_REQUIRE(M >= 1 && M <= 128, "%s: neighbours per point must be in [1, 128], got %d", who, M);
    P3D_REQUIRE(K >= 1 && K <= KP_K, "%s: kernel points must be in [1, 16], got %d", who, K);
    P3D_REQUIRE(Cin >= 1 && Cout >= 1, "%s: channel counts must be positive", who);
    P3D_REQUIRE(N0 < (int64_t)1 << 31, "%s: more than 2^31 support points", who);
    P3D_REQUIRE(influence >= 0 && influence <= 2, "%s: unknown KP_influence %d", who, influence);
    P3D_REQUIRE(aggregation >= 0 && aggregation <= 1, "%s: unknown aggregation_mode %d", who, aggregation);
    return P3D_OK;
}

int kpconv_fwd_simt(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                    const float *k_points, const float *k_values, int64_t n, int64_t N0, int M, int K, int Cin,
                    int Cout, float extent, int influence, int aggregation, float *out, cudaStream_t st) {
    const int Mpad = (M + 31) / 32 * 32;
    const size_t smem = (size_t)KPF_WARPS * (Mpad * KP_K + Mpad + KP_K * 32 * KPF_T) * sizeof(float);
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(kpconv_fwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done.mark();
    }
    const int64_t tiles = (n + KPF_T - 1) / KPF_T;
    int64_t grid = (tiles + KPF_WARPS - 1) / KPF_WARPS;
    if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
    kpconv_fwd_simt_kernel<<<(unsigned)grid, KPF_WARPS * 32, smem, st>>>(
        query, support, neighbors, feat, k_points, k_values, n, N0, M, Mpad, K, Cin, Cout,
        make_geom(extent, influence, aggregation), out);
    return finish_launch("kpconv_fwd_simt");
}

// C[R x N] = A[R x Kd] * B, B(k, j) = Bp[k * sbk + j * sbj], for SHORT reductions (Kd <= 64): the backward's
// dwf = dout . W^T at the shallow levels (Kd = Cout = 32 or 64, N = K * Cin = 480..960, R = 15-50 k points) is a
// write-bound problem -- 97 MB of output for 1.5 GFLOP -- that the tcgen05 GEMM spends on operand conversion and
// its epilogue (121 us).  Plain fp32: a CTA keeps a 256-column block of B in shared memory and walks row tiles of
// 64; thread tile 8 x 8 (two float4 column groups, so shared-memory reads are conflict-free and every row is
// stored as two coalesced 512-byte runs per warp).
constexpr int SK_ROWS = 64, SK_COLS = 256, SK_KD = 64, SK_APITCH = SK_ROWS + 4;

__global__ void __launch_bounds__(256, 2)
gemm_smallk_kernel(const float *__restrict__ A, int64_t lda, const float *__restrict__ Bp, int64_t sbk, int64_t sbj,
                   int64_t R, int Kd, int N, float *__restrict__ C, int64_t ldc) {
    extern __shared__ __align__(16) float sk_smem[];
    float *Bs = sk_smem;                  // [Kd][SK_COLS]
    float *As = sk_smem + Kd * SK_COLS;   // [Kd][SK_APITCH]
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int col0 = blockIdx.y * SK_COLS;
    for (int i = tid; i < Kd * SK_COLS; i += 256) {
        const int j = i / Kd, k = i - j * Kd;  // k fastest: contiguous when sbk == 1
        Bs[k * SK_COLS + j] = (col0 + j < N) ? __ldg(Bp + (int64_t)k * sbk + (int64_t)(col0 + j) * sbj) : 0.f;
    }
    const int64_t tiles = (R + SK_ROWS - 1) / SK_ROWS;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t row0 = tile * SK_ROWS;
        __syncthreads();  // previous tile's As reads are done (and, first time, Bs is complete)
        for (int i = tid; i < SK_ROWS * Kd; i += 256) {
            const int r = i / Kd, k = i - r * Kd;
            As[k * SK_APITCH + r] = (row0 + r < R) ? __ldg(A + (row0 + r) * lda + k) : 0.f;
        }
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int v = 0; v < 8; ++v) acc[u][v] = 0.f;
#pragma unroll 4
        for (int k = 0; k < Kd; ++k) {
            const float4 a0 = *reinterpret_cast<const float4 *>(As + k * SK_APITCH + ty * 8);
            const float4 a1 = *reinterpret_cast<const float4 *>(As + k * SK_APITCH + ty * 8 + 4);
            const float4 b0 = *reinterpret_cast<const float4 *>(Bs + k * SK_COLS + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4 *>(Bs + k * SK_COLS + 128 + tx * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int v = 0; v < 8; ++v) acc[u][v] = fmaf(a[u], b[v], acc[u][v]);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int64_t r = row0 + ty * 8 + u;
            if (r >= R) continue;
            float *crow = C + r * ldc + col0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = h * 128 + tx * 4;
                if (col0 + j + 3 < N && ((((uintptr_t)(crow + j)) & 15) == 0)) {
                    __stcs(reinterpret_cast<float4 *>(crow + j),
                           make_float4(acc[u][4 * h], acc[u][4 * h + 1], acc[u][4 * h + 2], acc[u][4 * h + 3]));
                } else {
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        if (col0 + j + v < N) crow[j + v] = acc[u][4 * h + v];
                }
            }
        }
    }
}

// returns 1 when the shape is outside this kernel (Kd > 64)
int gemm_smallk(const float *A, int64_t lda, const float *Bp, int64_t sbk, int64_t sbj, int64_t R, int Kd, int N,
                float *C, int64_t ldc, cudaStream_t st) {
    if (Kd > SK_KD || Kd < 1) return 1;
    const size_t smem = (size_t)Kd * (SK_COLS + SK_APITCH) * sizeof(float);
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(gemm_smallk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)(SK_KD * (SK_COLS + SK_APITCH) * sizeof(float)));
        attr_done.mark();
    }
    const int64_t tiles = (R + SK_ROWS - 1) / SK_ROWS;
    const int cblocks = (N + SK_COLS - 1) / SK_COLS;
    int64_t gx = (2 * kNumSMs + cblocks - 1) / cblocks;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    gemm_smallk_kernel<<<dim3((unsigned)gx, (unsigned)cblocks), 256, smem, st>>>(A, lda, Bp, sbk, sbj, R, Kd, N, C, ldc);
    return finish_launch("gemm_smallk");
}

// host-side launchers of the plain SGEMM, shared with kpconv_deform.cu
int sgemm_strided(const float *A, int64_t sai, int64_t sak, const float *B, int64_t sbk, int64_t sbj, float *C,
                  int64_t ldc, int64_t Mr, int64_t Nc, int64_t Kd, cudaStream_t st) {
    dim3 gd((unsigned)((Nc + 63) / 64), (unsigned)((Mr + 63) / 64), 1);
    sgemm_strided_kernel<<<gd, 256, 0, st>>>(A, sai, sak, B, sbk, sbj, C, ldc, Mr, Nc, Kd, Kd, 0);
    return finish_launch("sgemm_strided");
}

// dW[kc][o] += sum_p A[p][kc] * B[p][o]: split over p, fp32 reductions into the (zeroed) dW
int sgemm_splitk_atomic(const float *A, const float *B, int64_t n, int64_t KC, int Cout, float *dW, cudaStream_t st) {
    const int64_t tiles = ((KC + 63) / 64) * ((Cout + 63) / 64);
    int64_t split = (2 * kNumSMs + tiles - 1) / tiles;
    int64_t kps = (n + split - 1) / split;
    kps = (kps + 15) / 16 * 16;
    if (kps < 64) kps = 64;
    split = (n + kps - 1) / kps;
    dim3 gd((unsigned)((Cout + 63) / 64), (unsigned)((KC + 63) / 64), (unsigned)split);
    sgemm_strided_kernel<<<gd, 256, 0, st>>>(A, 1, KC, B, Cout, 1, dW, Cout, KC, Cout, n, kps, 1);
    return finish_launch("sgemm_splitk_atomic");
}

}  // namespace p3d

using namespace p3d;

namespace p3d {
// kpconv_tc.cu: tcgen05 path for the kernel-point contraction; returns 1 when it does not
// support the shape (the caller then uses the exact fp32 kernel above)
int kpconv_fwd_tc(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                  const float *k_points, const float *k_values, int64_t n, int64_t N0, int M, int K, int Cin,
                  int Cout, float extent, int influence, int aggregation, float *out, void *ws, size_t ws_bytes,
                  cudaStream_t st, float *wf_out = nullptr);
size_t kpconv_tc_ws_bytes(int64_t n, int K, int Cin, int Cout);
// tensor-core / register-tiled GEMMs of the backward (kpconv_tc.cu)
size_t gemm_tc_ws_bytes(int Kd, int N);
int gemm_tc(const float *A, int64_t lda, const float *Bp, int64_t sbk, int64_t sbj, int64_t R, int Kd, int N, float *C,
            int64_t ldc, void *ws, size_t ws_bytes, cudaStream_t st);
int dw_accumulate(const float *A, const float *B, int64_t n, int R, int Cn, float *dW, cudaStream_t st, float *part = nullptr,
                  size_t part_bytes = 0);
size_t dw_partial_bytes(int64_t n, int R, int Cn);
// rowgemm_tc.cu: tall tcgen05 GEMM (128-row tiles, wide hi|lo MMAs); returns 1 when the shape is not supported
size_t rowgemm_tc_ws_bytes(int Kd, int N);
int rowgemm_tc(const float *A, int64_t lda, const float *Bp, int64_t sbk, int64_t sbj, int64_t R, int Kd, int N, float *C,
               int64_t ldc, void *ws, size_t ws_bytes, cudaStream_t st);
// dW[R x N] += X^T . B on the same kernel (transposed operand, reduction cut over CTAs, deterministic partial sums)
size_t rowgemm_dw_ws_bytes(int64_t n, int R, int N);
int rowgemm_dw(const float *X, const float *Bp, int64_t n, int R, int N, float *dW, void *ws, size_t ws_bytes,
               cudaStream_t st);
// workspace of the weight gradient: the tensor-core kernel's image + partials, or the SIMT kernel's partial blocks
static size_t dw_ws_bytes(int64_t n, int KC, int Cout) {
    const size_t a = dw_partial_bytes(n, KC, Cout), b = rowgemm_dw_ws_bytes(n, KC, Cout);
    return a > b ? a : b;
}
// dW += wf^T . dout: tcgen05 row GEMM when the shape allows it (Cout % 16 == 0), else the register-tiled fp32 kernel;
// returns 1 when neither takes the shape
static int dw_product(const float *wf, const float *dout, int64_t n, int KC, int Cout, float *dW, cudaStream_t st,
                      void *ws, size_t ws_bytes) {
    // few points (deep levels: 114-464 points against a 7680 x 512 gradient) leave the tensor-core kernel with two or
    // three 64-point chunks per CTA and a weight image to build first: the fp32 kernel wins there (measured: 108 vs
    // 126 us at 256 -> 256 / 114 points; 176 vs 214 us the other way at 128 -> 128 / 2580 points)
    int r = n >= 1024 ? rowgemm_dw(wf, dout, n, KC, Cout, dW, ws, ws_bytes, st) : 1;
    if (r == 1) r = dw_accumulate(wf, dout, n, KC, Cout, dW, st, (float *)ws, ws_bytes);
    return r;
}
// bytes of the weight image of the backward's dwf GEMM (the larger of the two tensor-core kernels' needs)
static size_t dwf_image_bytes(int Cout, int KC) {
    const size_t a = gemm_tc_ws_bytes(Cout, KC), b = rowgemm_tc_ws_bytes(Cout, KC);
    return ((a > b ? a : b) + 255) / 256 * 256;
}
// sparse-influence backward helpers (kpconv_tc.cu)
bool kpconv_sparse_supported(int M, int K, int Cin);
int kpconv_wf_sparse(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                     const float *k_points, int64_t n, int64_t N0, int M, int K, int Cin, float extent, int influence,
                     int aggregation, float *wf, cudaStream_t st);
int kpconv_scatter_sparse(const float *query, const float *support, const int64_t *neighbors, const float *k_points,
                          const float *dwf, int64_t n, int64_t N0, int M, int K, int Cin, float extent, int influence,
                          int aggregation, float *dfeat, cudaStream_t st);
}  // namespace p3d

P3D_EXPORT size_t p3d_kpconv_fwd_ws_bytes(int64_t n, int K, int Cin, int Cout) { return kpconv_tc_ws_bytes(n, K, Cin, Cout); }

P3D_EXPORT int p3d_kpconv_fwd_f32(const float *query, const float *support, const int64_t *neighbors,
                                  const float *feat, const float *k_points, const float *k_values, int64_t n,
                                  int64_t N0, int M, int K, int Cin, int Cout, float extent, int influence,
                                  int aggregation, int precision, float *out, void *ws, size_t ws_bytes,
                                  void *stream) {
    if (int e = check_common("kpconv_fwd", n, N0, M, K, Cin, Cout, influence, aggregation)) return e;
    if (n == 0) return P3D_OK;
    P3D_REQUIRE(query && neighbors && k_points && k_values && out && (N0 == 0 || (support && feat)),
                "kpconv_fwd: null pointer");
    P3D_REQUIRE(precision == P3D_PRECISION_FP32 || precision == P3D_PRECISION_TENSOR,
                "kpconv_fwd: unknown precision %d", precision);
    if (precision == P3D_PRECISION_TENSOR) {
        const int r = kpconv_fwd_tc(query, support, neighbors, feat, k_points, k_values, n, N0, M, K, Cin, Cout,
                                    extent, influence, aggregation, out, ws, ws_bytes, (cudaStream_t)stream);
        if (r <= 0) return r;  // launched (0) or failed (<0); 1 = shape not covered by the tensor path
    }
    return kpconv_fwd_simt(query, support, neighbors, feat, k_points, k_values, n, N0, M, K, Cin, Cout, extent,
                           influence, aggregation, out, (cudaStream_t)stream);
}

// The two geometric halves of the operator on their own, for callers that run the dense contraction
// out = wf . K_values as a plain library GEMM: with a few hundred query points and 512 x 512 weights per kernel
// point (the deep levels of the KPConv encoder) the work is streaming 16 MB of weights, which a split-K GEMM
// spreads over the SMs and the fused kernel (2 point tiles x Cout/32 CTAs) does not.
P3D_EXPORT int p3d_kpconv_wf_f32(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                                 const float *k_points, int64_t n, int64_t N0, int M, int K, int Cin, float extent,
                                 int influence, int aggregation, float *wf, void *stream) {
    if (int e = check_common("kpconv_wf", n, N0, M, K, Cin, 1, influence, aggregation)) return e;
    if (n == 0) return P3D_OK;
    P3D_REQUIRE(query && neighbors && k_points && wf && (N0 == 0 || (support && feat)), "kpconv_wf: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (kpconv_sparse_supported(M, K, Cin))
        return kpconv_wf_sparse(query, support, neighbors, feat, k_points, n, N0, M, K, Cin, extent, influence,
                                aggregation, wf, st);
    const int Mpad = (M + 31) / 32 * 32;
    const size_t smem = (size_t)KPF_WARPS * (Mpad * KP_K + Mpad) * sizeof(float);
    cudaFuncSetAttribute(kpconv_wf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    int64_t grid = (n + KPF_WARPS - 1) / KPF_WARPS;
    if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
    kpconv_wf_kernel<<<(unsigned)grid, KPF_WARPS * 32, smem, st>>>(query, support, neighbors, feat, k_points, n, N0, M,
                                                                 Mpad, K, Cin, make_geom(extent, influence, aggregation),
                                                                 wf);
    return finish_launch("kpconv_wf");
}

P3D_EXPORT int p3d_kpconv_scatter_f32(const float *query, const float *support, const int64_t *neighbors,
                                      const float *k_points, const float *dwf, int64_t n, int64_t N0, int M, int K,
                                      int Cin, float extent, int influence, int aggregation, float *dfeat,
                                      void *stream) {
    if (int e = check_common("kpconv_scatter", n, N0, M, K, Cin, 1, influence, aggregation)) return e;
    if (n == 0) return P3D_OK;
    P3D_REQUIRE(query && neighbors && k_points && dwf && dfeat && (N0 == 0 || support), "kpconv_scatter: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (kpconv_sparse_supported(M, K, Cin))
        return kpconv_scatter_sparse(query, support, neighbors, k_points, dwf, n, N0, M, K, Cin, extent, influence,
                                     aggregation, dfeat, st);
    const int Mpad = (M + 31) / 32 * 32;
    const size_t smem = (size_t)KPF_WARPS * (Mpad * KP_K + Mpad) * sizeof(float);
    cudaFuncSetAttribute(kpconv_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    int64_t grid = (n + KPF_WARPS - 1) / KPF_WARPS;
    if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
    kpconv_scatter_kernel<<<(unsigned)grid, KPF_WARPS * 32, smem, st>>>(query, support, neighbors, k_points, dwf, n, N0,
                                                                      M, Mpad, K, Cin,
                                                                      make_geom(extent, influence, aggregation), dfeat);
    return finish_launch("kpconv_scatter");
}

// Training forward: the tensor path also stores the fp32 weighted features wf [n, K, Cin] it builds in registers, so
// the backward pass (p3d_kpconv_bwd_wf_f32) does not recompute them.  Only shapes the tensor path covers.
P3D_EXPORT int p3d_kpconv_fwd_wf_f32(const float *query, const float *support, const int64_t *neighbors,
                                     const float *feat, const float *k_points, const float *k_values, int64_t n,
                                     int64_t N0, int M, int K, int Cin, int Cout, float extent, int influence,
                                     int aggregation, float *out, float *wf, void *ws, size_t ws_bytes, void *stream) {
    if (int e = check_common("kpconv_fwd_wf", n, N0, M, K, Cin, Cout, influence, aggregation)) return e;
    if (n == 0) return P3D_OK;
    P3D_REQUIRE(query && neighbors && k_points && k_values && out && wf && (N0 == 0 || (support && feat)),
                "kpconv_fwd_wf: null pointer");
    const int r = kpconv_fwd_tc(query, support, neighbors, feat, k_points, k_values, n, N0, M, K, Cin, Cout, extent,
                                influence, aggregation, out, ws, ws_bytes, (cudaStream_t)stream, wf);
    if (r == 1) {
        set_error("kpconv_fwd_wf: shape outside the tensor path (Cin %% 32, Cout %% 32, Cout <= 512, M <= 40, workspace)");
        return P3D_ERR_INVALID;
    }
    return r;
}

static int kpconv_bwd_impl(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                           const float *k_points, const float *k_values, const float *dout, const float *wf_saved,
                           int64_t n, int64_t N0, int M, int K, int Cin, int Cout, float extent, int influence,
                           int aggregation, float *dfeat, float *dweights, void *ws, size_t ws_bytes, void *stream);

P3D_EXPORT size_t p3d_kpconv_bwd_ws_bytes(int64_t n, int M, int K, int Cin, int Cout) {
    (void)M;
    // wf + dwf  [n, K, Cin] each, then the bf16 hi/lo image of W^T for the tensor-core dwf GEMM (0 if unsupported),
    // then the per-chunk partial blocks of the deterministic dW reduction
    return (size_t)2 * (size_t)n * (size_t)K * (size_t)Cin * sizeof(float) + dwf_image_bytes(Cout, K * Cin) +
           dw_ws_bytes(n, K * Cin, Cout);
}

P3D_EXPORT int p3d_kpconv_bwd_f32(const float *query, const float *support, const int64_t *neighbors,
                                  const float *feat, const float *k_points, const float *k_values, const float *dout,
                                  int64_t n, int64_t N0, int M, int K, int Cin, int Cout, float extent,
                                  int influence, int aggregation, float *dfeat, float *dweights, void *ws,
                                  size_t ws_bytes, void *stream) {
    return kpconv_bwd_impl(query, support, neighbors, feat, k_points, k_values, dout, nullptr, n, N0, M, K, Cin, Cout,
                           extent, influence, aggregation, dfeat, dweights, ws, ws_bytes, stream);
}

// same, with the weighted features saved by p3d_kpconv_fwd_wf_f32 (wf [n, K, Cin]; nullptr = recompute)
P3D_EXPORT int p3d_kpconv_bwd_wf_f32(const float *query, const float *support, const int64_t *neighbors,
                                     const float *feat, const float *k_points, const float *k_values,
                                     const float *dout, const float *wf, int64_t n, int64_t N0, int M, int K, int Cin,
                                     int Cout, float extent, int influence, int aggregation, float *dfeat,
                                     float *dweights, void *ws, size_t ws_bytes, void *stream) {
    return kpconv_bwd_impl(query, support, neighbors, feat, k_points, k_values, dout, wf, n, N0, M, K, Cin, Cout, extent,
                           influence, aggregation, dfeat, dweights, ws, ws_bytes, stream);
}

static int kpconv_bwd_impl(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                           const float *k_points, const float *k_values, const float *dout, const float *wf_saved,
                           int64_t n, int64_t N0, int M, int K, int Cin, int Cout, float extent, int influence,
                           int aggregation, float *dfeat, float *dweights, void *ws, size_t ws_bytes, void *stream) {
    if (int e = check_common("kpconv_bwd", n, N0, M, K, Cin, Cout, influence, aggregation)) return e;
    if (n == 0 || (dfeat == nullptr && dweights == nullptr)) return P3D_OK;
    P3D_REQUIRE(query && neighbors && k_points && k_values && dout && (N0 == 0 || (support && feat)),
                "kpconv_bwd: null pointer");
    if (ws == nullptr || ws_bytes < p3d_kpconv_bwd_ws_bytes(n, M, K, Cin, Cout)) {
        set_error("kpconv_bwd: workspace of %zu bytes required", p3d_kpconv_bwd_ws_bytes(n, M, K, Cin, Cout));
        return P3D_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const KpGeom g = make_geom(extent, influence, aggregation);
    const int Mpad = (M + 31) / 32 * 32;
    const size_t smem = (size_t)KPF_WARPS * (Mpad * KP_K + Mpad) * sizeof(float);
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(kpconv_wf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(kpconv_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr_done.mark();
    }
    float *wf = (float *)ws;
    float *dwf = wf + (size_t)n * K * Cin;
    const int64_t KC = (int64_t)K * Cin;
    int64_t grid = (n + KPF_WARPS - 1) / KPF_WARPS;
    if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
    const bool sparse = kpconv_sparse_supported(M, K, Cin);
    const size_t part_bytes = dw_ws_bytes(n, (int)KC, Cout);
    float *dw_part = (float *)((char *)ws + (size_t)2 * n * K * Cin * sizeof(float) + dwf_image_bytes(Cout, K * Cin));
    if (dweights && wf_saved != nullptr) {
        // dW[kc][o] += sum_p wf[p][kc] * dout[p][o] straight from the forward's slab
        const int r = dw_product(wf_saved, dout, n, (int)KC, Cout, dweights, st, dw_part, part_bytes);
        if (r < 0) return r;
        if (r == 1)  // shape outside the register-tiled kernel
            if (int e = sgemm_splitk_atomic(wf_saved, dout, n, KC, Cout, dweights, st)) return e;
    } else if (dweights) {
        if (sparse) {
            if (int e = kpconv_wf_sparse(query, support, neighbors, feat, k_points, n, N0, M, K, Cin, extent, influence,
                                         aggregation, wf, st))
                return e;
        } else {
            kpconv_wf_kernel<<<(unsigned)grid, KPF_WARPS * 32, smem, st>>>(query, support, neighbors, feat, k_points,
                                                                         n, N0, M, Mpad, K, Cin, g, wf);
            if (int e = finish_launch("kpconv_wf")) return e;
        }
        // dW[kc][o] += sum_p wf[p][kc] * dout[p][o]
        {
        const int r = dw_product(wf, dout, n, (int)KC, Cout, dweights, st, dw_part, part_bytes);
        if (r < 0) return r;
        if (r == 0) goto dw_done;
        }
        {
        const int64_t tiles = ((KC + 63) / 64) * ((Cout + 63) / 64);
        int64_t split = (2 * kNumSMs + tiles - 1) / tiles;
        int64_t kps = (n + split - 1) / split;
        kps = (kps + 15) / 16 * 16;
        if (kps < 64) kps = 64;
        split = (n + kps - 1) / kps;
        dim3 gd((unsigned)((Cout + 63) / 64), (unsigned)((KC + 63) / 64), (unsigned)split);
        sgemm_strided_kernel<<<gd, 256, 0, st>>>(wf, 1, KC, dout, Cout, 1, dweights, Cout, KC, Cout, n, kps, 1);
        if (int e = finish_launch("kpconv_dW_gemm")) return e;
        }
    dw_done:;
    }
    if (dfeat) {
        // dwf[p][kc] = sum_o dout[p][o] * W[kc][o]: tcgen05 GEMM (bf16 hi+lo operands) when the shape allows it
        void *bimg = (char *)ws + (size_t)2 * n * K * Cin * sizeof(float);
        const size_t bimg_bytes = dwf_image_bytes(Cout, K * Cin);
        // the tall row GEMM takes every shape with K * Cin % 16 == 0 (measured against the two older kernels on the
        // encoder shapes: 61 vs 72 us at 32 -> 32 / 50 k points, 33 vs 70 us at 128 -> 128, 31 vs 90 us at 512 -> 512)
        int r = rowgemm_tc(dout, Cout, k_values, 1, Cout, n, Cout, (int)KC, dwf, KC, bimg, bimg_bytes, st);
        if (r == 1) r = gemm_smallk(dout, Cout, k_values, 1, Cout, n, Cout, (int)KC, dwf, KC, st);  // Cout <= 64, fp32
        if (r == 1) r = gemm_tc(dout, Cout, k_values, 1, Cout, n, Cout, (int)KC, dwf, KC, bimg, bimg_bytes, st);
        if (r < 0) return r;
        if (r == 1) {
            dim3 gd((unsigned)((KC + 63) / 64), (unsigned)((n + 63) / 64), 1);
            sgemm_strided_kernel<<<gd, 256, 0, st>>>(dout, Cout, 1, k_values, 1, Cout, dwf, KC, n, KC, Cout, Cout, 0);
            if (int e = finish_launch("kpconv_dwf_gemm")) return e;
        }
        if (sparse) {
            if (int e = kpconv_scatter_sparse(query, support, neighbors, k_points, dwf, n, N0, M, K, Cin, extent,
                                              influence, aggregation, dfeat, st))
                return e;
        } else {
            kpconv_scatter_kernel<<<(unsigned)grid, KPF_WARPS * 32, smem, st>>>(query, support, neighbors, k_points,
                                                                              dwf, n, N0, M, Mpad, K, Cin, g, dfeat);
            if (int e = finish_launch("kpconv_scatter")) return e;
        }
    }
    return P3D_OK;
}
