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

#include <cstdlib>

#include "async.cuh"
#include "common.cuh"

namespace p3d {

constexpr int GR_THREADS = 256;
constexpr int GR_VEC = 4;                        // output positions per thread per step
constexpr int GR_SMEM_FLOATS = 24 * 1024;        // 96 KB of staged feature rows per CTA (2 CTAs / SM)

// optional fused relative-coordinate channels of group_fwd_kernel (pos == nullptr: plain gather)
struct GroupXyz {
    const float *pos;      // [B,N,3]
    const float *new_pos;  // [B,S,3]
    int S, K;
    float norm;
    int do_norm;
};

// out[b, c_off+c, e] = feat[b, c, idx[b, e]],   e in [0, E = S*K)
// XYZ = false compiles the relative-coordinate block out of the plain gather
template <bool STAGED, bool XYZ = false>
__global__ void __launch_bounds__(GR_THREADS)
group_fwd_kernel(const float *__restrict__ feat, const int64_t *__restrict__ idx, int C, int N, int64_t E,
                 int CC, int e_per_cta, float *__restrict__ out, int64_t out_batch_stride, int c_off,
                 GroupXyz xyz) {
    extern __shared__ __align__(16) float rows[];  // [CC][N] when STAGED
    const int b = blockIdx.z;
    // fused _prepare_features: the first channel block of the grid (when xyz.pos is set) produces the three
    // relative-coordinate channels (pos[idx] - new_pos) * scale from the point-major position tensor
    const bool xyz_block = XYZ && blockIdx.y == 0;
    const int fy = blockIdx.y - (XYZ ? 1 : 0);
    const int c0 = xyz_block ? 0 : fy * CC;
    const int cc = xyz_block ? 3 : ((C - c0 < CC) ? (C - c0) : CC);
    const int64_t e0 = (int64_t)blockIdx.x * e_per_cta;
    const int64_t e1 = (e0 + e_per_cta < E) ? e0 + e_per_cta : E;
    const float *__restrict__ fb = feat + ((int64_t)b * C + c0) * N;
    const int64_t *__restrict__ ib = idx + (int64_t)b * E;
    float *__restrict__ ob = out + (int64_t)b * out_batch_stride + (int64_t)(xyz_block ? 0 : c_off + c0) * E;
    const float *__restrict__ nb = xyz_block ? xyz.new_pos + (int64_t)b * xyz.S * 3 : nullptr;

    if (STAGED) {
        const int total = cc * N;
        if (xyz_block) {
            const float *__restrict__ pb = xyz.pos + (int64_t)b * N * 3;
            for (int i = threadIdx.x; i < 3 * N; i += GR_THREADS) rows[(i % 3) * N + i / 3] = __ldg(pb + i);
        } else if ((((uintptr_t)fb) & 15) == 0 && (total & 3) == 0) {
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
                if (xyz_block) {
                    const int s0 = (int)(e / xyz.K), s3 = (int)((e + 3) / xyz.K);
                    const float n0 = __ldg(nb + 3 * s0 + ch);
                    if (s0 == s3) {
                        v.x = __fsub_rn(v.x, n0); v.y = __fsub_rn(v.y, n0); v.z = __fsub_rn(v.z, n0); v.w = __fsub_rn(v.w, n0);
                    } else {
                        v.x = __fsub_rn(v.x, n0);
                        v.y = __fsub_rn(v.y, __ldg(nb + 3 * (int)((e + 1) / xyz.K) + ch));
                        v.z = __fsub_rn(v.z, __ldg(nb + 3 * (int)((e + 2) / xyz.K) + ch));
                        v.w = __fsub_rn(v.w, __ldg(nb + 3 * s3 + ch));
                    }
                    if (xyz.do_norm) {  // norm = 1/r: ATen divides by a scalar as a multiply by 1/r
                        v.x = __fmul_rn(v.x, xyz.norm); v.y = __fmul_rn(v.y, xyz.norm);
                        v.z = __fmul_rn(v.z, xyz.norm); v.w = __fmul_rn(v.w, xyz.norm);
                    }
                }
                st_cs_f4(ob + (int64_t)ch * E + e, v);
            }
        } else {
            for (int t = 0; t < GR_VEC && e + t < e1; ++t) {
                const int jj = (int)__ldg(ib + e + t);
                for (int ch = 0; ch < cc; ++ch) {
                    const float *__restrict__ r = STAGED ? rows + ch * N : fb + (int64_t)ch * N;
                    float v = STAGED ? r[jj] : __ldg(r + jj);
                    if (xyz_block) {
                        v = __fsub_rn(v, __ldg(nb + 3 * (int)((e + t) / xyz.K) + ch));
                        if (xyz.do_norm) v = __fmul_rn(v, xyz.norm);
                    }
                    ob[(int64_t)ch * E + e + t] = v;
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
            if (do_norm) v = __fmul_rn(v, norm);  // norm = 1/r: ATen divides by a scalar as a multiply by 1/r
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
        const int total = cc * m;
        if ((((uintptr_t)fb) & 15) == 0 && (total & 3) == 0) {
            const float4 *__restrict__ s4 = reinterpret_cast<const float4 *>(fb);
            float4 *d4 = reinterpret_cast<float4 *>(rows);
            for (int i = threadIdx.x; i < total / 4; i += GR_THREADS) d4[i] = __ldg(s4 + i);
        } else {
            for (int i = threadIdx.x; i < total; i += GR_THREADS) rows[i] = __ldg(fb + i);
        }
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

// out[p,c] = max_m x'[nb[p,m], c];  one warp per query point, lanes over channels.  The M neighbour indices are
// read once per point (one coalesced load, broadcast by shuffle) and the row gathers are issued four at a time.
__global__ void __launch_bounds__(GR_THREADS)
neighbor_max_kernel(const float *__restrict__ x, const int64_t *__restrict__ nb, int64_t n, int64_t N0, int M,
                    int C, float *__restrict__ out, int64_t *__restrict__ arg) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * (GR_THREADS / 32) + (threadIdx.x >> 5);
    if (p >= n) return;
    for (int c0 = 0; c0 < C; c0 += 32) {
        const int c = c0 + lane;
        const bool cok = c < C;
        float best = -CUDART_INF_F;
        int64_t bi = N0;
        for (int m0 = 0; m0 < M; m0 += 32) {
            int64_t jl = (m0 + lane < M) ? __ldg(nb + p * M + m0 + lane) : -1;
            if (jl < 0 || jl >= N0) jl = -1;  // shadow row (zeros)
            const int cnt = (M - m0 < 32) ? (M - m0) : 32;
            for (int mm = 0; mm < cnt; mm += 4) {
                int64_t j[4];
                float v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    j[u] = __shfl_sync(0xffffffffu, jl, (mm + u) & 31);
                    v[u] = (mm + u < cnt && j[u] >= 0 && cok) ? __ldg(x + j[u] * C + c) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (mm + u < cnt && v[u] > best) { best = v[u]; bi = j[u] < 0 ? N0 : j[u]; }
            }
        }
        if (cok) {
            out[p * C + c] = best;
            if (arg) arg[p * C + c] = bi;
        }
    }
}

// C % 4 == 0: a warp owns (point, block of 128 channels), every lane four channels (one 16-byte load per neighbour
// row and lane, eight rows in flight); deep KPConv levels have ~100 query points x 1024 channels, so the channel
// blocks -- not the points -- provide the parallelism.
__global__ void __launch_bounds__(GR_THREADS)
neighbor_max_vec_kernel(const float *__restrict__ x, const int64_t *__restrict__ nb, int64_t n, int64_t N0, int M,
                        int C, int chunks, float *__restrict__ out, int64_t *__restrict__ arg) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (GR_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t p = w / chunks;
    if (p >= n) return;
    const int c = (int)(w - p * chunks) * 128 + lane * 4;
    const bool cok = c < C;
    float best[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    int64_t bi[4] = {N0, N0, N0, N0};
    for (int m0 = 0; m0 < M; m0 += 32) {
        int64_t jl = (m0 + lane < M) ? __ldg(nb + p * M + m0 + lane) : -1;
        if (jl < 0 || jl >= N0) jl = -1;  // shadow row (zeros)
        const int cnt = (M - m0 < 32) ? (M - m0) : 32;
        for (int mm = 0; mm < cnt; mm += 8) {
            int64_t j[8];
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                j[u] = __shfl_sync(0xffffffffu, jl, (mm + u) & 31);
                v[u] = (mm + u < cnt && j[u] >= 0 && cok) ? __ldg(reinterpret_cast<const float4 *>(x + j[u] * C + c))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (mm + u < cnt) {
                    const int64_t jj = j[u] < 0 ? N0 : j[u];
                    if (v[u].x > best[0]) { best[0] = v[u].x; bi[0] = jj; }
                    if (v[u].y > best[1]) { best[1] = v[u].y; bi[1] = jj; }
                    if (v[u].z > best[2]) { best[2] = v[u].z; bi[2] = jj; }
                    if (v[u].w > best[3]) { best[3] = v[u].w; bi[3] = jj; }
                }
            }
        }
    }
    if (cok) {
        *reinterpret_cast<float4 *>(out + p * C + c) = make_float4(best[0], best[1], best[2], best[3]);
        if (arg) {
#pragma unroll
            for (int u = 0; u < 4; ++u) arg[p * C + c + u] = bi[u];
        }
    }
}

// ------------------------------------------------- atomic-free backward (inverse index) ----
// Both scatter-type gradients (grouping, three_interpolate) are sums over "all output slots e
// that read source point j".  Instead of fp32 atomics at L2 (one RED per element: the bound of
// the first version of these kernels) the index tensor is inverted once per call:
//   invert_kernel   -- per cloud and per chunk of EC consecutive slots:
//                      (1) run-length compression in slot order: a run of equal consecutive
//                          indices (ball-query padding repeats the first hit up to K times) is
//                          represented by its first slot, and recorded in runs[] as (start, len);
//                      (2) a STABLE counting sort of the run heads by source point:
//                          order[b][chunk][EC] (uint16 slot offsets), seg[b][chunk][N+1].
//                      Per-warp histograms make the placement deterministic.
//   bwd_sorted      -- a CTA owns SB_CC channel rows of one cloud and streams the gradient rows chunk
//                      by chunk through a shared-memory ring filled by TMA bulk copies.  Per chunk,
//                      warps first fold every run into its head slot (coalesced shared-memory reads,
//                      fixed shuffle tree), then every thread sums the short lists of its own source
//                      points in increasing slot order.  No atomics, deterministic, each gradient
//                      element is read from HBM exactly once.
constexpr int INV_WARPS = 8;
constexpr int INV_MAX_N = 4096;   // per-warp histograms: N * 8 * 4 B of shared memory
constexpr int SB_THREADS = 512;
constexpr int SB_KPT = 4;         // source points per thread at most (one z-block covers 2048 points); the kernel is
                                  // instantiated for 1, 2 and 4 so that small tables do not pay registers for absent slots
__host__ inline int sb_kpt(int N) { return N <= 512 ? 1 : (N <= 1024 ? 2 : SB_KPT); }
constexpr int SB_CC = 4;          // channel rows per CTA
constexpr int SB_EC_GROUP = 2048; // slots per chunk (grouping)
constexpr int SB_EC_INTERP = 3072;  // slots per chunk (three_interpolate: 1024 points x 3 taps)
constexpr int SB_STAGES = 2;      // compile-time: the stage of a chunk is (chunk - first) & 1, no integer division

constexpr int INV_MAX_EC = 3072;  // SB_EC_INTERP
constexpr int INV_STEPS = INV_MAX_EC / (INV_WARPS * 32);  // 32-slot steps per warp: 12

// position of the first set bit at or after `pos` in a bit vector of `words` 32-bit words (or `limit`)
__device__ __forceinline__ int next_set_bit(const unsigned *bits, int words, int pos, int limit) {
    int w = pos >> 5;
    if (w >= words) return limit;
    unsigned m = bits[w] & (0xffffffffu << (pos & 31));
    while (m == 0u) {
        if (++w >= words) return limit;
        m = bits[w];
    }
    const int r = (w << 5) + __ffs(m) - 1;
    return r < limit ? r : limit;
}

__global__ void __launch_bounds__(INV_WARPS * 32)
invert_kernel(const int64_t *__restrict__ keys, int64_t E, int N, int EC, int nchunks, int compress,
              uint16_t *__restrict__ order, int *__restrict__ seg, uint32_t *__restrict__ runs,
              int *__restrict__ nruns) {
    extern __shared__ int ism[];
    int *hist = ism;                  // [INV_WARPS][N]  (warp-major: key-strided accesses are conflict-free)
    int *rowsum = ism + (size_t)N * INV_WARPS;  // [N]
    __shared__ int warp_tot[INV_WARPS];
    __shared__ int warp_runs[INV_WARPS];
    __shared__ unsigned headbits[INV_MAX_EC / 32];  // bit e = slot e is a run head
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int64_t e0 = (int64_t)chunk * EC;
    const int len = (int)((E - e0 < EC) ? (E - e0) : EC);
    const int64_t *__restrict__ kb = keys + (int64_t)b * E + e0;
    uint16_t *__restrict__ ob = order + ((int64_t)b * nchunks + chunk) * EC;
    int *__restrict__ sb = seg + ((int64_t)b * nchunks + chunk) * (N + 1);
    uint32_t *__restrict__ rb = runs + ((int64_t)b * nchunks + chunk) * (EC / 2);
    const int words = (EC + 31) / 32;

    for (int i = tid; i < N * INV_WARPS; i += INV_WARPS * 32) hist[i] = 0;
    const int sub = ((len + INV_WARPS * 32 - 1) / (INV_WARPS * 32)) * 32;  // slots per warp (multiple of 32)
    const int w0 = warp * sub;
    const int w1 = (w0 + sub < len) ? (w0 + sub) : len;
    // all keys of this warp's slots in registers (one coalesced load per step, issued back to back)
    int key[INV_STEPS];
#pragma unroll
    for (int s = 0; s < INV_STEPS; ++s) {
        const int e = w0 + s * 32 + lane;
        int64_t j = -1;
        if (s * 32 < sub && e < w1) j = __ldg(kb + e);
        key[s] = (j >= 0 && j < N) ? (int)j : -1;
    }
    int64_t jprev = -2;  // key of the slot just before this warp's range (lane 0 of step 0 needs it)
    if (compress && w0 > 0 && w0 < len) jprev = __ldg(kb + w0 - 1);
    __syncthreads();  // hist zeroed
    unsigned headm[INV_STEPS];
    int carry = (int)((jprev >= 0 && jprev < N) ? jprev : -2);
#pragma unroll
    for (int s = 0; s < INV_STEPS; ++s) {
        headm[s] = 0u;
        if (s * 32 < sub) {
            int prev = __shfl_up_sync(0xffffffffu, key[s], 1);
            if (lane == 0) prev = carry;
            const bool head = key[s] >= 0 && (!compress || key[s] != prev);
            headm[s] = __ballot_sync(0xffffffffu, head);
            // headbits marks every slot that BREAKS a run: heads and slots with an out-of-range index
            const unsigned breakm = headm[s] | __ballot_sync(0xffffffffu, key[s] < 0);
            if (head) atomicAdd(&hist[warp * N + key[s]], 1);
            carry = __shfl_sync(0xffffffffu, key[s], 31);
            if (lane == 0 && w0 + s * 32 < EC) headbits[(w0 >> 5) + s] = breakm;
        }
    }
    // words past the chunk length are empty
    for (int w = ((len + 31) >> 5) + tid; w < words; w += INV_WARPS * 32) headbits[w] = 0u;
    __syncthreads();
    // a RUN START is a head whose next slot exists, holds the same key and therefore is not a head
    int my_runs = 0;
    unsigned runm[INV_STEPS];
#pragma unroll
    for (int s = 0; s < INV_STEPS; ++s) {
        runm[s] = 0u;
        if (s * 32 < sub && compress) {
            const int e = w0 + s * 32 + lane;
            bool rs = false;
            if ((headm[s] >> lane) & 1u) {
                const int nx = e + 1;
                rs = nx < len && !((headbits[nx >> 5] >> (nx & 31)) & 1u);
            }
            runm[s] = __ballot_sync(0xffffffffu, rs);
            my_runs += __popc(runm[s]);
        }
    }
    if (lane == 0) warp_runs[warp] = my_runs;
    for (int j = tid; j < N; j += INV_WARPS * 32) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < INV_WARPS; ++w) t += hist[w * N + j];
        rowsum[j] = t;
    }
    __syncthreads();
    // exclusive scan of rowsum[N]: contiguous block per thread, warp scan, cross-warp offsets
    const int per = (N + INV_WARPS * 32 - 1) / (INV_WARPS * 32);
    const int j0 = tid * per;
    int local = 0;
    for (int i = 0; i < per; ++i)
        if (j0 + i < N) local += rowsum[j0 + i];
    int incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int base = incl - local;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    for (int i = 0; i < per; ++i) {  // rowsum[j] <- exclusive prefix (in place: each thread owns its block)
        const int j = j0 + i;
        if (j < N) {
            const int c = rowsum[j];
            rowsum[j] = base;
            base += c;
        }
    }
    if (tid == INV_WARPS * 32 - 1) sb[N] = base;  // the last thread ends on the total count
    __syncthreads();
    // per-(warp, key) start offsets and the segment table, key-strided (coalesced, conflict-free)
    for (int j = tid; j < N; j += INV_WARPS * 32) {
        int run = rowsum[j];
        sb[j] = run;
#pragma unroll
        for (int w = 0; w < INV_WARPS; ++w) {
            const int c = hist[w * N + j];
            hist[w * N + j] = run;
            run += c;
        }
    }
    int run_base = 0;
    for (int w = 0; w < warp; ++w) run_base += warp_runs[w];
    if (tid == INV_WARPS * 32 - 1) nruns[(int64_t)b * nchunks + chunk] = run_base + my_runs;
    __syncthreads();
    // stable placement of the heads + run table: every warp walks its own slots in order
#pragma unroll
    for (int s = 0; s < INV_STEPS; ++s) {
        if (s * 32 < sub) {
            const int e = w0 + s * 32 + lane;
            if ((runm[s] >> lane) & 1u) {
                const int end = next_set_bit(headbits, words, e + 1, len);  // next head (or the chunk end)
                rb[run_base + __popc(runm[s] & lanemask_lt())] = ((uint32_t)e << 16) | (uint32_t)(end - e - 1);
            }
            run_base += __popc(runm[s]);
            const bool valid = (headm[s] >> lane) & 1u;
            if (valid) {
                const unsigned peers = __match_any_sync(headm[s], key[s]);
                const int rank = __popc(peers & lanemask_lt());
                const int leader = __ffs(peers) - 1;
                int pos = 0;
                if (lane == leader) {
                    pos = hist[warp * N + key[s]];
                    hist[warp * N + key[s]] = pos + __popc(peers);
                }
                pos = __shfl_sync(peers, pos, leader);
                ob[pos + rank] = (uint16_t)e;
            }
            __syncwarp();
        }
    }
}

// INTERP = false: dfeat[b,c,j] = sum_{e: idx[b,e]=j} dout[b, c_off+c, e]
// INTERP = true : dfeat[b,c,j] = sum_{e: idx[b,e]=j} dout[b, c, e/3] * weight[b, e]   (no compression)
// One pipeline stage in shared memory: [SB_CC][Wd] gradient values | (INTERP: [EC] weights) |
// [EC] uint16 sorted head slots | (!INTERP: [EC/2] uint32 runs); the per-key list bounds (seg) are
// prefetched one chunk ahead in registers, so the inner loops touch shared memory only.
// gridDim.z = key blocks x splits: split s sums chunks [s*cps, (s+1)*cps) into its own partial
// gradient (reduced in split order by reduce_partials_kernel), which gives small-C problems
// (C = 3 at the first SA level) enough CTAs without atomics.
template <bool INTERP, int KPT>
__global__ void __launch_bounds__(SB_THREADS, 2)
bwd_sorted_kernel(const float *__restrict__ dout, const float *__restrict__ weight, const uint16_t *__restrict__ order,
                  const int *__restrict__ seg, const uint32_t *__restrict__ runs, const int *__restrict__ nruns, int C,
                  int N, int64_t E, int EC, int nchunks, int splits, int cps, float *__restrict__ dst_base,
                  int64_t split_stride, int64_t dout_batch_stride, int c_off, int use_tma) {
    extern __shared__ __align__(128) float stage[];
    __shared__ __align__(8) uint64_t full[SB_STAGES];
    const int Wd = INTERP ? EC / 3 : EC;             // gradient values per channel row and chunk
    const int buf_floats = SB_CC * Wd + (INTERP ? EC : 0) + EC / 2 + (INTERP ? 0 : EC / 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * SB_CC;
    const int split = blockIdx.z % splits;
    const int jbase = (blockIdx.z / splits) * (KPT * SB_THREADS);
    const int ch_lo = split * cps;
    const int ch_hi = (ch_lo + cps < nchunks) ? (ch_lo + cps) : nchunks;
    const int cc = (C - c0 < SB_CC) ? (C - c0) : SB_CC;
    const int64_t Ed = INTERP ? E / 3 : E;  // length of a gradient row
    const float *__restrict__ gb = dout + (int64_t)b * dout_batch_stride + (int64_t)(c_off + c0) * Ed;
    const float *__restrict__ wb = INTERP ? weight + (int64_t)b * E : nullptr;
    // TMA bulk copies need 16-byte aligned rows and lengths
    const bool tma = use_tma && ((Ed & 3) == 0) && ((((uintptr_t)gb) & 15) == 0) && ((Wd & 3) == 0) &&
                     (!INTERP || (((E & 3) == 0) && ((((uintptr_t)wb) & 15) == 0)));

    float acc[KPT][SB_CC];
#pragma unroll
    for (int k = 0; k < KPT; ++k)
#pragma unroll
        for (int c = 0; c < SB_CC; ++c) acc[k][c] = 0.f;

    if (tid == 0) {
        for (int s = 0; s < SB_STAGES; ++s) mbar_init(&full[s], 1);
        fence_barrier_init();
    }
    __syncthreads();

    // producer side (thread 0): all copies of one chunk land on one mbarrier
    auto issue = [&](int chunk) {
        float *dst = stage + (size_t)((chunk - ch_lo) % SB_STAGES) * buf_floats;
        uint64_t *bar = &full[(chunk - ch_lo) % SB_STAGES];
        const int64_t d0 = (int64_t)chunk * Wd;
        const int dl = (int)((Ed - d0 < Wd) ? (Ed - d0) : Wd);
        const int64_t e0 = (int64_t)chunk * EC;
        const int el = (int)((E - e0 < EC) ? (E - e0) : EC);
        const uint32_t bytes = (uint32_t)(cc * dl * 4 + EC * 2 + (INTERP ? el * 4 : EC * 2));
        fence_proxy_async();  // the run folding wrote this stage through the generic proxy
        mbar_arrive_expect_tx(bar, bytes);
        for (int ch = 0; ch < cc; ++ch) bulk_g2s(dst + ch * Wd, gb + (int64_t)ch * Ed + d0, (uint32_t)dl * 4, bar);
        if (INTERP) bulk_g2s(dst + SB_CC * Wd, wb + e0, (uint32_t)el * 4, bar);
        float *tail = dst + SB_CC * Wd + (INTERP ? EC : 0);
        bulk_g2s(tail, order + ((int64_t)b * nchunks + chunk) * EC, (uint32_t)EC * 2, bar);
        if (!INTERP) bulk_g2s(tail + EC / 2, runs + ((int64_t)b * nchunks + chunk) * (EC / 2), (uint32_t)EC * 2, bar);
    };
    // fallback for unaligned shapes: cooperative synchronous fill
    auto fill = [&](int chunk) {
        float *dst = stage + (size_t)((chunk - ch_lo) % SB_STAGES) * buf_floats;
        const int64_t d0 = (int64_t)chunk * Wd;
        const int dl = (int)((Ed - d0 < Wd) ? (Ed - d0) : Wd);
        for (int i = tid; i < cc * Wd; i += SB_THREADS) {
            const int ch = i / Wd, p = i - ch * Wd;
            dst[ch * Wd + p] = p < dl ? __ldcs(gb + (int64_t)ch * Ed + d0 + p) : 0.f;
        }
        if (INTERP) {
            const int64_t e0 = (int64_t)chunk * EC;
            const int el = (int)((E - e0 < EC) ? (E - e0) : EC);
            for (int i = tid; i < EC; i += SB_THREADS) dst[SB_CC * Wd + i] = i < el ? __ldg(wb + e0 + i) : 0.f;
        }
        const uint4 *__restrict__ osrc = reinterpret_cast<const uint4 *>(order + ((int64_t)b * nchunks + chunk) * EC);
        uint4 *odst = reinterpret_cast<uint4 *>(dst + SB_CC * Wd + (INTERP ? EC : 0));
        for (int i = tid; i < EC / 8; i += SB_THREADS) odst[i] = __ldg(osrc + i);
        if (!INTERP) {
            const uint4 *__restrict__ rsrc =
                reinterpret_cast<const uint4 *>(runs + ((int64_t)b * nchunks + chunk) * (EC / 2));
            for (int i = tid; i < EC / 8; i += SB_THREADS) odst[EC / 8 + i] = __ldg(rsrc + i);
        }
    };
    int s0[KPT], s1[KPT], nr = 0;
    // key slots of this thread that exist at all (N = 512 -> 1): the others cost no instructions
    const int kpt = (N - jbase + SB_THREADS - 1) / SB_THREADS < KPT ? (N - jbase + SB_THREADS - 1) / SB_THREADS : KPT;
#pragma unroll
    for (int k = 0; k < KPT; ++k) s0[k] = s1[k] = 0;
    auto bounds = [&](int chunk) {
        const int *__restrict__ sb = seg + ((int64_t)b * nchunks + chunk) * (N + 1);
#pragma unroll
        for (int k = 0; k < KPT; ++k) {
            if (k >= kpt) break;
            const int j = jbase + k * SB_THREADS + tid;
            s0[k] = j < N ? __ldg(sb + j) : 0;
            s1[k] = j < N ? __ldg(sb + j + 1) : 0;
        }
        if (!INTERP) nr = __ldg(nruns + (int64_t)b * nchunks + chunk);
    };

    if (tma) {
        if (tid == 0)
            for (int c = ch_lo; c < ch_hi && c < ch_lo + SB_STAGES - 1; ++c) issue(c);
    } else if (ch_lo < ch_hi) {
        fill(ch_lo);
        __syncthreads();
    }
    if (ch_lo < ch_hi) bounds(ch_lo);
    for (int chunk = ch_lo; chunk < ch_hi; ++chunk) {
        const int it = chunk - ch_lo;
        float *cur = stage + (size_t)(it % SB_STAGES) * buf_floats;
        int a0[KPT], a1[KPT];
#pragma unroll
        for (int k = 0; k < KPT; ++k) { a0[k] = s0[k]; a1[k] = s1[k]; }
        const int nr_cur = nr;
        if (tma) {
            if (tid == 0 && chunk + SB_STAGES - 1 < ch_hi) issue(chunk + SB_STAGES - 1);
        } else if (chunk + 1 < ch_hi) {
            fill(chunk + 1);
        }
        if (chunk + 1 < ch_hi) bounds(chunk + 1);
        if (tma) mbar_wait(&full[it % SB_STAGES], (uint32_t)((it / SB_STAGES) & 1));
        const float *wst = cur + SB_CC * Wd;
        const uint16_t *ob = reinterpret_cast<const uint16_t *>(cur + SB_CC * Wd + (INTERP ? EC : 0));
        if (!INTERP) {
            // fold every run of repeated indices into its head slot: one warp per run, lanes stride
            // the run, fixed xor tree -> deterministic
            const uint32_t *rt = reinterpret_cast<const uint32_t *>(ob + EC);
#pragma unroll 1
            for (int r = warp; r < nr_cur; r += SB_THREADS / 32) {
                const uint32_t rr = rt[r];
                const int st = (int)(rr >> 16), ln = (int)(rr & 0xffffu) + 1;
                float v[SB_CC];
#pragma unroll
                for (int c = 0; c < SB_CC; ++c) v[c] = 0.f;
#pragma unroll 1
                for (int i = lane; i < ln; i += 32) {
#pragma unroll
                    for (int c = 0; c < SB_CC; ++c) v[c] += cur[c * Wd + st + i];
                }
                // 4 values x 32 lanes -> 4 sums with 6 shuffles instead of 20: the upper half-warp takes over rows 2,3
                // (step 16), then the odd quarter the second row of each pair (step 8); fixed order -> deterministic
                static_assert(SB_CC == 4, "the butterfly below folds exactly 4 rows");
                {
                    const bool up = lane & 16;
                    const float s0 = up ? v[0] : v[2], s1 = up ? v[1] : v[3];   // what this lane gives away
                    const float k0 = up ? v[2] : v[0], k1 = up ? v[3] : v[1];   // what it keeps: rows (0,1) or (2,3)
                    const float p0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
                    const float p1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
                    const bool odd = lane & 8;
                    float m = (odd ? p1 : p0) + __shfl_xor_sync(0xffffffffu, odd ? p0 : p1, 8);
                    m += __shfl_xor_sync(0xffffffffu, m, 4);
                    m += __shfl_xor_sync(0xffffffffu, m, 2);
                    m += __shfl_xor_sync(0xffffffffu, m, 1);
                    // lanes 0..7 hold row 0, 8..15 row 1, 16..23 row 2, 24..31 row 3
                    if ((lane & 7) == 0) cur[(lane >> 3) * Wd + st] = m;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int k = 0; k < KPT; ++k) {
            if (k >= kpt) break;
            const int te = a1[k];
            int t = a0[k];
            if (!INTERP) {
                // two list entries per step: their index and value loads are independent (the walk is a chain of
                // shared-memory latencies otherwise); the additions keep the slot order
#pragma unroll 1
                if (KPT == 1) {
#pragma unroll 1
                    for (; t + 4 <= te; t += 4) {
                        const int e0 = ob[t], e1 = ob[t + 1], e2 = ob[t + 2], e3 = ob[t + 3];
                        float v0[SB_CC], v1[SB_CC], v2[SB_CC], v3[SB_CC];
#pragma unroll
                        for (int c = 0; c < SB_CC; ++c) {
                            v0[c] = cur[c * Wd + e0]; v1[c] = cur[c * Wd + e1];
                            v2[c] = cur[c * Wd + e2]; v3[c] = cur[c * Wd + e3];
                        }
#pragma unroll
                        for (int c = 0; c < SB_CC; ++c) acc[k][c] = (((acc[k][c] + v0[c]) + v1[c]) + v2[c]) + v3[c];
                    }
                }
#pragma unroll 1
                for (; KPT < 4 && t + 2 <= te; t += 2) {   // (the 4-slot instantiation has no registers to spare)
                    const int e0 = ob[t], e1 = ob[t + 1];
                    float v0[SB_CC], v1[SB_CC];
#pragma unroll
                    for (int c = 0; c < SB_CC; ++c) { v0[c] = cur[c * Wd + e0]; v1[c] = cur[c * Wd + e1]; }
#pragma unroll
                    for (int c = 0; c < SB_CC; ++c) acc[k][c] = (acc[k][c] + v0[c]) + v1[c];
                }
            }
#pragma unroll 1
            for (; t < te; ++t) {  // slot order: deterministic
                const int ea = ob[t];
                if (INTERP) {
                    const float wa = wst[ea];
                    const int pa = ea / 3;
#pragma unroll
                    for (int c = 0; c < SB_CC; ++c) acc[k][c] = fmaf(cur[c * Wd + pa], wa, acc[k][c]);
                } else {
#pragma unroll
                    for (int c = 0; c < SB_CC; ++c) acc[k][c] += cur[c * Wd + ea];
                }
            }
        }
        __syncthreads();  // stage (it % SB_STAGES) may be refilled from the next iteration on
    }
    float *__restrict__ dst = dst_base + (int64_t)split * split_stride;
#pragma unroll
    for (int k = 0; k < KPT; ++k) {
        const int j = jbase + k * SB_THREADS + tid;
        if (j < N) {
#pragma unroll
            for (int c = 0; c < SB_CC; ++c)
                if (c < cc) dst[((int64_t)b * C + c0 + c) * N + j] = acc[k][c];
        }
    }
}

// out[i] = ((p[0][i] + p[1][i]) + ...) in split order
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float *__restrict__ part, int splits, int64_t n, float *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
        float s = part[i];
        for (int k = 1; k < splits; ++k) s += part[(int64_t)k * n + i];
        out[i] = s;
    }
}

// ---------------------------------------------------------------------------------------------
// Narrow grouping backward (C * N <= 6144 floats, i.e. the first set-abstraction level: C = 3 normals / colours
// gathered from N = 2048 raw points into up to 65 536 slots per cloud).  The sorted kernel above spends most of its
// time on per-chunk bookkeeping there (three channel rows cannot amortise it: 91 us per C2 step for 22 MB of
// gradient, plus three inverse-index builds).  Here the WHOLE gradient table of a cloud, [C][N] floats, fits in
// shared memory once per warp: a CTA owns a contiguous slot range of one cloud, each of its 8 warps walks its own
// sub-range in slot order, 32 consecutive slots per step (coalesced gradient and index loads), and adds into its
// private table.  Duplicate indices inside a step (ball-query padding) are found with __match_any_sync and summed
// by a fixed shuffle tree, so no two lanes ever touch the same word; the 8 warp tables are added in warp order,
// the CTAs' partial tables in split order (reduce_partials_kernel): atomic-free and deterministic, and it needs
// no inverse index at all.
constexpr int DS_WARPS = 8;
constexpr int DS_MAX_FLOATS = 6144;  // C * N per warp table: 8 x 24 KB = 192 KB of shared memory
constexpr int DS_MAX_C = 8;

// CC = channel rows held in registers per step (>= C), U = 32-slot steps whose loads are in flight together
template <int CC, int U>
__global__ void __launch_bounds__(DS_WARPS * 32, 1)
bwd_direct_kernel(const float *__restrict__ dout, const int64_t *__restrict__ idx, int C, int N, int64_t E,
                  int64_t slots_per_cta, float *__restrict__ dst_base, int64_t split_stride, int64_t dout_batch_stride,
                  int c_off) {
    extern __shared__ __align__(16) float ds_acc[];  // [DS_WARPS][C*N]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, split = blockIdx.x;
    const int CN = C * N;
    for (int i = tid; i < DS_WARPS * CN; i += DS_WARPS * 32) ds_acc[i] = 0.f;
    __syncthreads();
    float *acc = ds_acc + (size_t)warp * CN;
    const int64_t r0 = (int64_t)split * slots_per_cta;
    const int64_t r1 = (r0 + slots_per_cta < E) ? r0 + slots_per_cta : E;
    const int64_t per = ((r1 - r0 + DS_WARPS - 1) / DS_WARPS + 31) / 32 * 32;  // slots per warp, multiple of 32
    const int64_t w0 = r0 + (int64_t)warp * per;
    const int64_t w1 = (w0 + per < r1) ? w0 + per : r1;
    const int64_t *__restrict__ ib = idx + (int64_t)b * E;
    const float *__restrict__ gb = dout + (int64_t)b * dout_batch_stride + (int64_t)c_off * E;
    // U steps (32 slots each) per iteration: all of their index and gradient loads are issued before the first
    // one is consumed -- a warp otherwise pays two dependent DRAM round trips per step with nothing else in flight
    for (int64_t e0 = w0; e0 < w1; e0 += 32 * U) {
        int64_t j[U];
        float v[U][CC];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = e0 + u * 32 + lane;
            j[u] = (e < w1) ? __ldg(ib + e) : -1;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t e = e0 + u * 32 + lane;
#pragma unroll
            for (int c = 0; c < CC; ++c) v[u][c] = (c < C && e < w1) ? __ldcs(gb + (int64_t)c * E + e) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (e0 + u * 32 >= w1) break;  // warp-uniform
            const bool valid = j[u] >= 0 && j[u] < N;
            const int jj = (int)j[u];
            const unsigned peers = __match_any_sync(0xffffffffu, valid ? jj : N + lane);
            const bool single = peers == (1u << lane);
            const unsigned leaders = __ballot_sync(0xffffffffu, valid && !single && (__ffs(peers) - 1) == lane);
#pragma unroll
            for (int c = 0; c < CC; ++c) {
                if (c >= C) break;
                const float val = v[u][c];
                if (valid && single) acc[c * N + jj] += val;
                unsigned m = leaders;
                while (m) {  // groups of equal indices inside this step: fixed shuffle tree over the members
                    const int L = __ffs(m) - 1;
                    m &= m - 1;
                    const unsigned members = __shfl_sync(0xffffffffu, peers, L);
                    float x = ((members >> lane) & 1u) ? val : 0.f;
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
                    if (lane == L) acc[c * N + jj] += x;
                }
            }
            __syncwarp();
        }
    }
    __syncthreads();
    float *__restrict__ dst = dst_base + (int64_t)split * split_stride + (int64_t)b * CN;
    for (int i = tid; i < CN; i += DS_WARPS * 32) {
        float sacc = ds_acc[i];
#pragma unroll
        for (int w = 1; w < DS_WARPS; ++w) sacc += ds_acc[(size_t)w * CN + i];
        dst[i] = sacc;
    }
}

static bool use_direct_path(int C, int N) {
    return C <= DS_MAX_C && (int64_t)C * N <= DS_MAX_FLOATS;
}

static int direct_splits(int B, int64_t E) {
    int splits = (kNumSMs + B - 1) / B;  // one CTA per SM
    const int64_t max_splits = (E + DS_WARPS * 32 - 1) / (DS_WARPS * 32);
    if (splits > max_splits) splits = (int)max_splits;
    if (B * splits > kNumSMs && splits > 1) --splits;
    return splits < 1 ? 1 : splits;
}

static int launch_bwd_direct(const float *dout, const int64_t *keys, int B, int C, int N, int64_t E, float *dfeat,
                             int64_t dout_batch_stride, int c_off, void *ws, cudaStream_t st) {
    const int splits = direct_splits(B, E);
    const int64_t total = (int64_t)B * C * N;
    int64_t spc = (E + splits - 1) / splits;
    spc = (spc + 31) / 32 * 32;
    const size_t smem = (size_t)DS_WARPS * C * N * sizeof(float);
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(bwd_direct_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             DS_WARPS * DS_MAX_FLOATS * (int)sizeof(float));
        cudaFuncSetAttribute(bwd_direct_kernel<DS_MAX_C, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             DS_WARPS * DS_MAX_FLOATS * (int)sizeof(float));
        attr_done.mark();
    }
    float *part = (float *)ws;
    const dim3 grid((unsigned)splits, (unsigned)B);
    if (C <= 4)  // the usual case (3 normals / colours): half the value registers, twice the steps in flight
        bwd_direct_kernel<4, 8><<<grid, DS_WARPS * 32, smem, st>>>(dout, keys, C, N, E, spc, splits > 1 ? part : dfeat,
                                                                  total, dout_batch_stride, c_off);
    else
        bwd_direct_kernel<DS_MAX_C, 4><<<grid, DS_WARPS * 32, smem, st>>>(dout, keys, C, N, E, spc,
                                                                         splits > 1 ? part : dfeat, total,
                                                                         dout_batch_stride, c_off);
    if (int e = finish_launch("group_bwd_direct")) return e;
    if (splits > 1) {
        int64_t gr = (total + 255) / 256;
        if (gr > 4 * kNumSMs) gr = 4 * kNumSMs;
        reduce_partials_kernel<<<(unsigned)gr, 256, 0, st>>>(part, splits, total, dfeat);
        return finish_launch("reduce_partials");
    }
    return P3D_OK;
}

static int pick_splits(int B, int C, int N, int nchunks) {
    const int64_t ctas = (int64_t)((C + SB_CC - 1) / SB_CC) * B * ((N + sb_kpt(N) * SB_THREADS - 1) / (sb_kpt(N) * SB_THREADS));
    int splits = 1;
    if (ctas < 2 * kNumSMs) splits = (int)((2 * kNumSMs + ctas - 1) / ctas);
    if (splits > nchunks) splits = nchunks;
    if (splits > 16) splits = 16;
    return splits < 1 ? 1 : splits;
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

struct InverseLayout {
    size_t order, seg, runs, nruns, total;
    int nchunks, splits;
};

static InverseLayout inverse_layout(int B, int C, int N, int64_t E, int EC) {
    InverseLayout L;
    L.nchunks = (int)((E + EC - 1) / EC);
    L.order = 0;
    L.seg = L.order + align256((size_t)B * L.nchunks * EC * sizeof(uint16_t));
    L.runs = L.seg + align256((size_t)B * L.nchunks * (N + 1) * sizeof(int));
    L.nruns = L.runs + align256((size_t)B * L.nchunks * (EC / 2) * sizeof(uint32_t));
    const size_t part = L.nruns + align256((size_t)B * L.nchunks * sizeof(int));
    L.splits = pick_splits(B, C, N, L.nchunks);
    L.total = part + (L.splits > 1 ? (size_t)L.splits * B * C * N * sizeof(float) : 0);
    return L;
}

enum { BWD_ALL = 0, BWD_PREPARE = 1, BWD_PREPARED = 2 };

template <bool INTERP>
static int launch_bwd_sorted(const float *dout, const float *weight, const int64_t *keys, int B, int C, int N,
                             int64_t E, float *dfeat, int64_t dout_batch_stride, int c_off, void *ws,
                             cudaStream_t st, int mode = BWD_ALL) {
    const int EC = INTERP ? SB_EC_INTERP : SB_EC_GROUP;
    const InverseLayout L = inverse_layout(B, C, N, E, EC);
    const int nchunks = L.nchunks;
    char *w8 = (char *)ws;
    uint16_t *order = (uint16_t *)(w8 + L.order);
    int *seg = (int *)(w8 + L.seg);
    uint32_t *runs = (uint32_t *)(w8 + L.runs);
    int *nruns = (int *)(w8 + L.nruns);
    float *part = (float *)(w8 + L.nruns + align256((size_t)B * nchunks * sizeof(int)));
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(bwd_sorted_kernel<INTERP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(bwd_sorted_kernel<INTERP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(bwd_sorted_kernel<INTERP, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        attr_done.mark();
    }
    if (mode != BWD_PREPARED) {
        const size_t inv_smem = ((size_t)N * INV_WARPS + N) * sizeof(int);
        invert_kernel<<<dim3((unsigned)nchunks, (unsigned)B), INV_WARPS * 32, inv_smem, st>>>(
            keys, E, N, EC, nchunks, INTERP ? 0 : 1, order, seg, runs, nruns);
        if (int e = finish_launch("invert")) return e;
        if (mode == BWD_PREPARE) return P3D_OK;
    }
    const int Wd = INTERP ? EC / 3 : EC;
    const int use_tma = 1;
    const size_t smem =
        (size_t)SB_STAGES * (SB_CC * Wd + (INTERP ? EC : 0) + EC / 2 + (INTERP ? 0 : EC / 2)) * sizeof(float);
    const int splits = L.splits;
    const int cps = (nchunks + splits - 1) / splits;
    const int kpt = sb_kpt(N);
    const int zblocks = (N + kpt * SB_THREADS - 1) / (kpt * SB_THREADS);
    dim3 grid((unsigned)((C + SB_CC - 1) / SB_CC), (unsigned)B, (unsigned)(zblocks * splits));
    const int64_t total = (int64_t)B * C * N;
    auto kern = kpt == 1 ? bwd_sorted_kernel<INTERP, 1> : (kpt == 2 ? bwd_sorted_kernel<INTERP, 2> : bwd_sorted_kernel<INTERP, 4>);
    kern<<<grid, SB_THREADS, smem, st>>>(dout, weight, order, seg, runs, nruns, C, N, E, EC,
                                                             nchunks, splits, cps, splits > 1 ? part : dfeat, total,
                                                             dout_batch_stride, c_off, use_tma);
    if (int e = finish_launch(INTERP ? "three_interp_bwd_sorted" : "group_bwd_sorted")) return e;
    if (splits > 1) {
        int64_t gr = (total + 255) / 256;
        if (gr > 4 * kNumSMs) gr = 4 * kNumSMs;
        reduce_partials_kernel<<<(unsigned)gr, 256, 0, st>>>(part, splits, total, dfeat);
        return finish_launch("reduce_partials");
    }
    return P3D_OK;
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

// backward of the neighbour max-pool: the gradient of out[p,c] goes to the winning row arg[p,c] (the shadow row N0
// receives nothing).  Several query points can have picked the same support row: fp32 reductions at L2, like the
// dfeat scatter of the KPConv backward.
__global__ void __launch_bounds__(GR_THREADS)
neighbor_max_bwd_kernel(const float *__restrict__ dout, const int64_t *__restrict__ arg, int64_t total, int C,
                        int64_t N0, float *__restrict__ dx) {
    for (int64_t i = (int64_t)blockIdx.x * GR_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * GR_THREADS) {
        const int64_t row = __ldg(arg + i);
        if (row >= 0 && row < N0) atomicAdd(dx + row * C + (i % C), __ldg(dout + i));
    }
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
        static DeviceOnce attr_done;
        if (!attr_done.done()) {
            cudaFuncSetAttribute(group_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 GR_SMEM_FLOATS * (int)sizeof(float));
            attr_done.mark();
        }
        group_fwd_kernel<true><<<grid, GR_THREADS, smem, st>>>(feat, idx, C, N, E, CC, (int)e_per_cta, out,
                                                               out_batch_stride, c_off, GroupXyz{});
    } else {
        group_fwd_kernel<false><<<grid, GR_THREADS, 0, st>>>(feat, idx, C, N, E, CC, (int)e_per_cta, out,
                                                             out_batch_stride, c_off, GroupXyz{});
    }
    return finish_launch("group_fwd");
}

P3D_EXPORT int p3d_group_xyz_feat_f32(const float *pos, const float *new_pos, const float *feat, const int64_t *idx,
                                      int B, int C, int N, int S, int K, float norm, float *out, void *stream) {
    P3D_REQUIRE(B >= 0 && C >= 1 && N >= 0 && S >= 0 && K >= 0, "group_xyz_feat: bad dimension");
    P3D_REQUIRE(norm > 0.f, "group_xyz_feat: norm must be positive");
    const int64_t E = (int64_t)S * K;
    if (B == 0 || E == 0) return P3D_OK;
    P3D_REQUIRE(pos && new_pos && feat && idx && out, "group_xyz_feat: null pointer");
    P3D_REQUIRE(N > 0, "group_xyz_feat: empty point table");
    const int64_t stride = (int64_t)(3 + C) * E;
    bool staged;
    const int CC = pick_cc(C, N, staged);
    if (!staged || 3 * N > GR_SMEM_FLOATS) {  // tables too large to stage: the two plain kernels
        if (int e = p3d_group_xyz_f32(pos, new_pos, idx, B, N, S, K, norm, out, stride, stream)) return e;
        return p3d_group_fwd_f32(feat, idx, B, C, N, S, K, out, stride, 3, stream);
    }
    const int cchunks = (C + CC - 1) / CC + 1;  // + the relative-coordinate block
    int64_t e_per_cta = (int64_t)GR_THREADS * GR_VEC * 8;
    while (e_per_cta > GR_THREADS * GR_VEC && ((E + e_per_cta - 1) / e_per_cta) * cchunks * B < 4 * kNumSMs)
        e_per_cta /= 2;
    dim3 grid((unsigned)((E + e_per_cta - 1) / e_per_cta), (unsigned)cchunks, (unsigned)B);
    const size_t smem = (size_t)((CC > 3 ? CC : 3)) * N * sizeof(float);
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(group_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             GR_SMEM_FLOATS * (int)sizeof(float));
        attr_done.mark();
    }
    GroupXyz xyz{pos, new_pos, S, K, 1.0f / norm, norm != 1.0f ? 1 : 0};
    group_fwd_kernel<true, true><<<grid, GR_THREADS, smem, (cudaStream_t)stream>>>(feat, idx, C, N, E, CC,
                                                                                  (int)e_per_cta, out, stride, 3, xyz);
    return finish_launch("group_xyz_feat");
}

P3D_EXPORT size_t p3d_group_bwd_ws_bytes(int B, int C, int N, int S, int K) {
    if (B > 65535) return 0;
    if (use_direct_path(C, N)) {  // partial tables of the narrow path: [splits][B*C*N]
        const int splits = direct_splits(B, (int64_t)S * K);
        return splits > 1 ? (size_t)splits * B * C * N * sizeof(float) : 256;
    }
    if (N > INV_MAX_N) return 0;  // atomic path needs no workspace
    return inverse_layout(B, C, N, (int64_t)S * K, SB_EC_GROUP).total;
}

P3D_EXPORT int p3d_group_bwd_prepare(const int64_t *idx, int B, int C, int N, int S, int K, void *ws, size_t ws_bytes,
                                     void *stream) {
    P3D_REQUIRE(B >= 0 && C >= 0 && N >= 0 && S >= 0 && K >= 0, "group_bwd_prepare: negative dimension");
    const size_t need = p3d_group_bwd_ws_bytes(B, C, N, S, K);
    if (need == 0 || B == 0 || C == 0 || N == 0 || (int64_t)S * K == 0) return P3D_OK;  // nothing to prepare
    P3D_REQUIRE(idx, "group_bwd_prepare: null pointer");
    if (ws == nullptr || ws_bytes < need) {
        set_error("group_bwd_prepare: workspace of %zu bytes required (p3d_group_bwd_ws_bytes)", need);
        return P3D_ERR_WORKSPACE;
    }
    if (use_direct_path(C, N)) return P3D_OK;  // the narrow path reads the index tensor directly
    return launch_bwd_sorted<false>(nullptr, nullptr, idx, B, C, N, (int64_t)S * K, nullptr, 0, 0, ws,
                                    (cudaStream_t)stream, BWD_PREPARE);
}

P3D_EXPORT int p3d_group_bwd_f32(const float *dout, const int64_t *idx, int B, int C, int N, int S, int K,
                                 float *dfeat, int64_t dout_batch_stride, int c_off, void *ws, size_t ws_bytes,
                                 int prepared, void *stream) {
    P3D_REQUIRE(B >= 0 && C >= 0 && N >= 0 && S >= 0 && K >= 0, "group_bwd: negative dimension");
    const int64_t E = (int64_t)S * K;
    if (B == 0 || C == 0 || N == 0) return P3D_OK;
    P3D_REQUIRE(dfeat, "group_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (E == 0) {
        cudaMemsetAsync(dfeat, 0, (size_t)B * C * N * sizeof(float), st);
        return P3D_OK;
    }
    P3D_REQUIRE(dout && idx, "group_bwd: null pointer");
    const size_t need = p3d_group_bwd_ws_bytes(B, C, N, S, K);
    if (need > 0 && ws != nullptr && ws_bytes >= need) {
        if (use_direct_path(C, N))
            return launch_bwd_direct(dout, idx, B, C, N, E, dfeat, dout_batch_stride, c_off, ws, st);
        return launch_bwd_sorted<false>(dout, nullptr, idx, B, C, N, E, dfeat, dout_batch_stride, c_off, ws, st,
                                        prepared ? BWD_PREPARED : BWD_ALL);
    }
    if (need > 0) {
        set_error("group_bwd: workspace of %zu bytes required (p3d_group_bwd_ws_bytes)", need);
        return P3D_ERR_WORKSPACE;
    }
    // very large source tables: fp32 reductions at L2 into a zeroed gradient
    cudaMemsetAsync(dfeat, 0, (size_t)B * C * N * sizeof(float), st);
    const int CC = C < 16 ? C : 16;
    const int cchunks = (C + CC - 1) / CC;
    int64_t e_per_cta = (int64_t)GR_THREADS * GR_VEC * 4;
    while (e_per_cta > GR_THREADS * GR_VEC && ((E + e_per_cta - 1) / e_per_cta) * cchunks * B < 4 * kNumSMs)
        e_per_cta /= 2;
    dim3 grid((unsigned)((E + e_per_cta - 1) / e_per_cta), (unsigned)cchunks, (unsigned)B);
    group_bwd_kernel<<<grid, GR_THREADS, 0, st>>>(dout, idx, C, N, E, CC, (int)e_per_cta, dfeat, dout_batch_stride,
                                                  c_off);
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
    group_xyz_kernel<<<grid, GR_THREADS, 0, (cudaStream_t)stream>>>(pos, new_pos, idx, N, S, K, 1.0f / norm,
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
    int CC = pick_cc(C, m, staged);
    // CTA tile = CC staged rows x n_per_cta output columns.  The fill (CC * m floats) is amortised over the columns,
    // so tiles are tall in n (>= 4 m columns when n allows) and the CTA count is made up with narrower channel
    // blocks -- 32 rows x 256 columns re-read the table 8 times on the C2 shapes.
    int n_per_cta = GR_THREADS * 8;
    if (staged) {
        while (n_per_cta > GR_THREADS && n_per_cta / 2 >= 4 * m) n_per_cta /= 2;
        while (n_per_cta / 2 >= n && n_per_cta > GR_THREADS) n_per_cta /= 2;
        const int64_t col_tiles = (n + n_per_cta - 1) / n_per_cta;
        while (CC > 4 && col_tiles * ((C + CC - 1) / CC) * B < 2 * kNumSMs) CC = (CC + 1) / 2;
    }
    const int cchunks = (C + CC - 1) / CC;
    while (n_per_cta > GR_THREADS && (int64_t)((n + n_per_cta - 1) / n_per_cta) * cchunks * B < 2 * kNumSMs)
        n_per_cta /= 2;
    dim3 grid((unsigned)((n + n_per_cta - 1) / n_per_cta), (unsigned)cchunks, (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    if (staged) {
        static DeviceOnce attr_done;
        if (!attr_done.done()) {
            cudaFuncSetAttribute(three_interp_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 GR_SMEM_FLOATS * (int)sizeof(float));
            attr_done.mark();
        }
        three_interp_fwd_kernel<true><<<grid, GR_THREADS, (size_t)CC * m * sizeof(float), st>>>(
            feat, idx, weight, C, m, n, CC, n_per_cta, out);
    } else {
        three_interp_fwd_kernel<false><<<grid, GR_THREADS, 0, st>>>(feat, idx, weight, C, m, n, CC, n_per_cta, out);
    }
    return finish_launch("three_interp_fwd");
}

P3D_EXPORT size_t p3d_three_interp_bwd_ws_bytes(int B, int C, int m, int n) {
    if (m > INV_MAX_N || B > 65535) return 0;
    return inverse_layout(B, C, m, (int64_t)n * 3, SB_EC_INTERP).total;
}

P3D_EXPORT int p3d_three_interp_bwd_prepare(const int64_t *idx, int B, int C, int m, int n, void *ws, size_t ws_bytes,
                                            void *stream) {
    P3D_REQUIRE(B >= 0 && C >= 0 && m >= 0 && n >= 0, "three_interpolate_bwd_prepare: negative dimension");
    const size_t need = p3d_three_interp_bwd_ws_bytes(B, C, m, n);
    if (need == 0 || B == 0 || C == 0 || m == 0 || n == 0) return P3D_OK;
    P3D_REQUIRE(idx, "three_interpolate_bwd_prepare: null pointer");
    if (ws == nullptr || ws_bytes < need) {
        set_error("three_interpolate_bwd_prepare: workspace of %zu bytes required", need);
        return P3D_ERR_WORKSPACE;
    }
    return launch_bwd_sorted<true>(nullptr, nullptr, idx, B, C, m, (int64_t)n * 3, nullptr, 0, 0, ws,
                                   (cudaStream_t)stream, BWD_PREPARE);
}

P3D_EXPORT int p3d_three_interp_bwd_f32(const float *dout, const int64_t *idx, const float *weight, int B, int C,
                                        int m, int n, float *dfeat, void *ws, size_t ws_bytes, int prepared,
                                        void *stream) {
    P3D_REQUIRE(B >= 0 && C >= 0 && m >= 0 && n >= 0, "three_interpolate_bwd: negative dimension");
    if (B == 0 || C == 0 || m == 0) return P3D_OK;
    P3D_REQUIRE(dfeat, "three_interpolate_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        cudaMemsetAsync(dfeat, 0, (size_t)B * C * m * sizeof(float), st);
        return P3D_OK;
    }
    P3D_REQUIRE(dout && idx && weight, "three_interpolate_bwd: null pointer");
    const size_t need = p3d_three_interp_bwd_ws_bytes(B, C, m, n);
    if (need > 0 && ws != nullptr && ws_bytes >= need)
        return launch_bwd_sorted<true>(dout, weight, idx, B, C, m, (int64_t)n * 3, dfeat, (int64_t)C * n, 0, ws, st,
                                       prepared ? BWD_PREPARED : BWD_ALL);
    if (need > 0) {
        set_error("three_interpolate_bwd: workspace of %zu bytes required (p3d_three_interp_bwd_ws_bytes)", need);
        return P3D_ERR_WORKSPACE;
    }
    P3D_REQUIRE(C <= 65535 && B <= 65535, "three_interpolate_bwd: C and B must be <= 65535");
    cudaMemsetAsync(dfeat, 0, (size_t)B * C * m * sizeof(float), st);
    int gx = (n + GR_THREADS - 1) / GR_THREADS;
    dim3 grid((unsigned)gx, (unsigned)C, (unsigned)B);
    three_interp_bwd_kernel<<<grid, GR_THREADS, 0, st>>>(dout, idx, weight, C, m, n, dfeat);
    return finish_launch("three_interp_bwd");
}

P3D_EXPORT int p3d_neighbor_max_f32(const float *x, const int64_t *neighbors, int64_t n, int64_t N0, int M, int C,
                                    float *out, int64_t *arg, void *stream) {
    P3D_REQUIRE(n >= 0 && N0 >= 0 && M >= 1 && C >= 0, "neighbor_max: bad dimension");
    if (n == 0 || C == 0) return P3D_OK;
    P3D_REQUIRE(neighbors && out && (x || N0 == 0), "neighbor_max: null pointer");
    if ((C & 3) == 0 && (((uintptr_t)x) & 15) == 0 && (((uintptr_t)out) & 15) == 0) {
        const int chunks = (C + 127) / 128;
        const int64_t warps = n * chunks;
        const int64_t grid = (warps + GR_THREADS / 32 - 1) / (GR_THREADS / 32);
        neighbor_max_vec_kernel<<<(unsigned)grid, GR_THREADS, 0, (cudaStream_t)stream>>>(x, neighbors, n, N0, M, C,
                                                                                        chunks, out, arg);
        return finish_launch("neighbor_max_vec");
    }
    const int64_t grid = (n + GR_THREADS / 32 - 1) / (GR_THREADS / 32);
    neighbor_max_kernel<<<(unsigned)grid, GR_THREADS, 0, (cudaStream_t)stream>>>(x, neighbors, n, N0, M, C, out, arg);
    return finish_launch("neighbor_max");
}

P3D_EXPORT int p3d_neighbor_max_bwd_f32(const float *dout, const int64_t *arg, int64_t n, int64_t N0, int C, float *dx,
                                        void *stream) {
    P3D_REQUIRE(n >= 0 && N0 >= 0 && C >= 0, "neighbor_max_bwd: bad dimension");
    cudaStream_t st = (cudaStream_t)stream;
    if (N0 > 0 && C > 0) {
        P3D_REQUIRE(dx, "neighbor_max_bwd: null pointer");
        cudaMemsetAsync(dx, 0, (size_t)N0 * C * sizeof(float), st);
    }
    if (n == 0 || C == 0 || N0 == 0) return P3D_OK;
    P3D_REQUIRE(dout && arg, "neighbor_max_bwd: null pointer");
    const int64_t total = n * C;
    int64_t grid = (total + GR_THREADS - 1) / GR_THREADS;
    if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
    neighbor_max_bwd_kernel<<<(unsigned)grid, GR_THREADS, 0, st>>>(dout, arg, total, C, N0, dx);
    return finish_launch("neighbor_max_bwd");
}
