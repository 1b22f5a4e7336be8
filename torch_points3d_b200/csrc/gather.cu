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


// ------------------------------------------------- atomic-free backward (inverse index) ----
// Both scatter-type gradients (grouping, three_interpolate) are sums over "all output slots e
// that read source point j".  Instead of fp32 atomics at L2 (one RED per element: the bound of
// the first version of these kernels) the index tensor is inverted once per call:
//   invert_kernel   -- per cloud and per chunk of EC consecutive slots, a STABLE counting sort of
//                      the slots by source point: order[b][chunk][EC] (uint16 slot offsets) and
//                      seg[b][chunk][N+1].  Per-warp histograms make the placement deterministic.
//   *_bwd_sorted    -- a CTA owns CC channel rows of one cloud, streams the gradient rows chunk by
//                      chunk through shared memory (coalesced float4, register double buffering)
//                      and every thread sums the lists of its own source points out of shared
//                      memory in increasing slot order.  No atomics, deterministic, each
//                      gradient element is read from HBM exactly once.
constexpr int INV_WARPS = 8;
constexpr int INV_MAX_N = 4096;   // per-warp histograms: N * 8 * 4 B of shared memory
constexpr int SB_THREADS = 512;
constexpr int SB_KPT = 4;         // source points per thread (one z-block covers 2048 points)
constexpr int SB_CC = 4;          // channel rows per CTA
constexpr int SB_EC_GROUP = 2048; // slots per chunk (grouping)
constexpr int SB_EC_INTERP = 3072;  // slots per chunk (three_interpolate: 1024 points x 3 taps)

__global__ void __launch_bounds__(INV_WARPS * 32)
invert_kernel(const int64_t *__restrict__ keys, int64_t E, int N, int EC, int nchunks,
              uint16_t *__restrict__ order, int *__restrict__ seg) {
    extern __shared__ int ism[];
    int *hist = ism;                  // [N][INV_WARPS]
    int *rowsum = ism + (size_t)N * INV_WARPS;  // [N]
    __shared__ int warp_tot[INV_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int64_t e0 = (int64_t)chunk * EC;
    const int len = (int)((E - e0 < EC) ? (E - e0) : EC);
    const int64_t *__restrict__ kb = keys + (int64_t)b * E + e0;
    uint16_t *__restrict__ ob = order + ((int64_t)b * nchunks + chunk) * EC;
    int *__restrict__ sb = seg + ((int64_t)b * nchunks + chunk) * (N + 1);

    for (int i = tid; i < N * INV_WARPS; i += INV_WARPS * 32) hist[i] = 0;
    __syncthreads();
    const int sub = ((len + INV_WARPS * 32 - 1) / (INV_WARPS * 32)) * 32;  // slots per warp (multiple of 32)
    const int w0 = warp * sub;
    const int w1 = (w0 + sub < len) ? (w0 + sub) : len;
    for (int e = w0 + lane; e < w1; e += 32) {
        const int64_t j = __ldg(kb + e);
        if (j >= 0 && j < N) atomicAdd(&hist[(int)j * INV_WARPS + warp], 1);
    }
    __syncthreads();
    for (int j = tid; j < N; j += INV_WARPS * 32) {
        int t = 0;
#pragma unroll
        for (int w = 0; w < INV_WARPS; ++w) t += hist[j * INV_WARPS + w];
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
    for (int i = 0; i < per; ++i) {
        const int j = j0 + i;
        if (j < N) {
            sb[j] = base;
            int run = base;
#pragma unroll
            for (int w = 0; w < INV_WARPS; ++w) {
                const int c = hist[j * INV_WARPS + w];
                hist[j * INV_WARPS + w] = run;
                run += c;
            }
            base = run;
        }
    }
    if (tid == INV_WARPS * 32 - 1) sb[N] = base;  // last thread ends on the total count
    __syncthreads();
    // stable placement: every warp walks its own slots in order
    for (int eb = w0; eb < w1; eb += 32) {
        const int e = eb + lane;
        int64_t j = -1;
        if (e < w1) j = __ldg(kb + e);
        const bool valid = (j >= 0 && j < N);
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            const unsigned peers = __match_any_sync(act, (int)j);
            const int rank = __popc(peers & lanemask_lt());
            const int leader = __ffs(peers) - 1;
            int pos = 0;
            if (lane == leader) {
                pos = hist[(int)j * INV_WARPS + warp];
                hist[(int)j * INV_WARPS + warp] = pos + __popc(peers);
            }
            pos = __shfl_sync(peers, pos, leader);
            ob[pos + rank] = (uint16_t)e;
        }
        __syncwarp();
    }
}

// INTERP = false: dfeat[b,c,j] = sum_{e: idx[b,e]=j} dout[b, c_off+c, e]
// INTERP = true : dfeat[b,c,j] = sum_{e: idx[b,e]=j} dout[b, c, e/3] * weight[b, e]
template <bool INTERP>
__global__ void __launch_bounds__(SB_THREADS)
bwd_sorted_kernel(const float *__restrict__ dout, const float *__restrict__ weight, const uint16_t *__restrict__ order,
                  const int *__restrict__ seg, int C, int N, int64_t E, int EC, int nchunks,
                  float *__restrict__ dfeat, int64_t dout_batch_stride, int c_off) {
    extern __shared__ __align__(16) float stage[];  // [2][SB_CC * W + (INTERP ? EC : 0)]
    const int Wd = INTERP ? EC / 3 : EC;             // gradient values per channel row and chunk
    const int buf_floats = SB_CC * Wd + (INTERP ? EC : 0);
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int c0 = blockIdx.x * SB_CC;
    const int cc = (C - c0 < SB_CC) ? (C - c0) : SB_CC;
    const int jbase = blockIdx.z * (SB_KPT * SB_THREADS);
    const int64_t Ed = INTERP ? E / 3 : E;  // length of a gradient row
    const float *__restrict__ gb = dout + (int64_t)b * dout_batch_stride + (int64_t)(c_off + c0) * Ed;
    const float *__restrict__ wb = INTERP ? weight + (int64_t)b * E : nullptr;
    const bool vec = ((Ed & 3) == 0) && ((((uintptr_t)gb) & 15) == 0) && ((Wd & 3) == 0);

    float acc[SB_KPT][SB_CC];
#pragma unroll
    for (int k = 0; k < SB_KPT; ++k)
#pragma unroll
        for (int c = 0; c < SB_CC; ++c) acc[k][c] = 0.f;

    auto fill = [&](int chunk, float *dst) {
        const int64_t d0 = (int64_t)chunk * Wd;
        const int dl = (int)((Ed - d0 < Wd) ? (Ed - d0) : Wd);
        if (vec) {
            const int q4 = Wd / 4;
            for (int i = tid; i < cc * q4; i += SB_THREADS) {
                const int ch = i / q4, p4 = i - ch * q4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p4 * 4 < dl) v = __ldcs(reinterpret_cast<const float4 *>(gb + (int64_t)ch * Ed + d0) + p4);
                reinterpret_cast<float4 *>(dst + ch * Wd)[p4] = v;
            }
        } else {
            for (int i = tid; i < cc * Wd; i += SB_THREADS) {
                const int ch = i / Wd, p = i - ch * Wd;
                dst[ch * Wd + p] = p < dl ? __ldcs(gb + (int64_t)ch * Ed + d0 + p) : 0.f;
            }
        }
        if (INTERP) {
            const int64_t e0 = (int64_t)chunk * EC;
            const int el = (int)((E - e0 < EC) ? (E - e0) : EC);
            for (int i = tid; i < EC; i += SB_THREADS) dst[SB_CC * Wd + i] = i < el ? __ldg(wb + e0 + i) : 0.f;
        }
    };

    fill(0, stage);
    __syncthreads();
    for (int chunk = 0; chunk < nchunks; ++chunk) {
        float *cur = stage + (size_t)(chunk & 1) * buf_floats;
        if (chunk + 1 < nchunks) fill(chunk + 1, stage + (size_t)((chunk + 1) & 1) * buf_floats);
        const uint16_t *__restrict__ ob = order + ((int64_t)b * nchunks + chunk) * EC;
        const int *__restrict__ sb = seg + ((int64_t)b * nchunks + chunk) * (N + 1);
        const float *wst = cur + SB_CC * Wd;
#pragma unroll
        for (int k = 0; k < SB_KPT; ++k) {
            const int j = jbase + k * SB_THREADS + tid;
            if (j < N) {
                const int s0 = __ldg(sb + j), s1 = __ldg(sb + j + 1);
                for (int t = s0; t < s1; ++t) {
                    const int e = ob[t];
                    if (INTERP) {
                        const float w = wst[e];
                        const int p = e / 3;
#pragma unroll
                        for (int c = 0; c < SB_CC; ++c) acc[k][c] = fmaf(cur[c * Wd + p], w, acc[k][c]);
                    } else {
#pragma unroll
                        for (int c = 0; c < SB_CC; ++c) acc[k][c] += cur[c * Wd + e];
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < SB_KPT; ++k) {
        const int j = jbase + k * SB_THREADS + tid;
        if (j < N) {
#pragma unroll
            for (int c = 0; c < SB_CC; ++c)
                if (c < cc) dfeat[((int64_t)b * C + c0 + c) * N + j] = acc[k][c];
        }
    }
}

static size_t inverse_ws_bytes(int B, int N, int64_t E, int EC) {
    const int64_t nchunks = (E + EC - 1) / EC;
    size_t order = (size_t)B * nchunks * EC * sizeof(uint16_t);
    order = (order + 255) / 256 * 256;
    const size_t seg = (size_t)B * nchunks * (N + 1) * sizeof(int);
    return order + seg;
}

template <bool INTERP>
static int launch_bwd_sorted(const float *dout, const float *weight, const int64_t *keys, int B, int C, int N,
                             int64_t E, float *dfeat, int64_t dout_batch_stride, int c_off, void *ws,
                             cudaStream_t st) {
    const int EC = INTERP ? SB_EC_INTERP : SB_EC_GROUP;
    const int nchunks = (int)((E + EC - 1) / EC);
    size_t order_bytes = (size_t)B * nchunks * EC * sizeof(uint16_t);
    order_bytes = (order_bytes + 255) / 256 * 256;
    uint16_t *order = (uint16_t *)ws;
    int *seg = (int *)((char *)ws + order_bytes);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(invert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(bwd_sorted_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        cudaFuncSetAttribute(bwd_sorted_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        attr_done = true;
    }
    const size_t inv_smem = ((size_t)N * INV_WARPS + N) * sizeof(int);
    invert_kernel<<<dim3((unsigned)nchunks, (unsigned)B), INV_WARPS * 32, inv_smem, st>>>(keys, E, N, EC, nchunks,
                                                                                       order, seg);
    if (int e = finish_launch("invert")) return e;
    const int Wd = INTERP ? EC / 3 : EC;
    const size_t smem = (size_t)2 * (SB_CC * Wd + (INTERP ? EC : 0)) * sizeof(float);
    dim3 grid((unsigned)((C + SB_CC - 1) / SB_CC), (unsigned)B,
              (unsigned)((N + SB_KPT * SB_THREADS - 1) / (SB_KPT * SB_THREADS)));
    bwd_sorted_kernel<INTERP><<<grid, SB_THREADS, smem, st>>>(dout, weight, order, seg, C, N, E, EC, nchunks, dfeat,
                                                             dout_batch_stride, c_off);
    return finish_launch(INTERP ? "three_interp_bwd_sorted" : "group_bwd_sorted");
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

P3D_EXPORT size_t p3d_group_bwd_ws_bytes(int B, int N, int S, int K) {
    if (N > INV_MAX_N || B > 65535) return 0;  // atomic path needs no workspace
    return inverse_ws_bytes(B, N, (int64_t)S * K, SB_EC_GROUP);
}

P3D_EXPORT int p3d_group_bwd_f32(const float *dout, const int64_t *idx, int B, int C, int N, int S, int K,
                                 float *dfeat, int64_t dout_batch_stride, int c_off, void *ws, size_t ws_bytes,
                                 void *stream) {
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
    const size_t need = p3d_group_bwd_ws_bytes(B, N, S, K);
    if (need > 0 && ws != nullptr && ws_bytes >= need)
        return launch_bwd_sorted<false>(dout, nullptr, idx, B, C, N, E, dfeat, dout_batch_stride, c_off, ws, st);
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

P3D_EXPORT size_t p3d_three_interp_bwd_ws_bytes(int B, int m, int n) {
    if (m > INV_MAX_N || B > 65535) return 0;
    return inverse_ws_bytes(B, m, (int64_t)n * 3, SB_EC_INTERP);
}

P3D_EXPORT int p3d_three_interp_bwd_f32(const float *dout, const int64_t *idx, const float *weight, int B, int C,
                                        int m, int n, float *dfeat, void *ws, size_t ws_bytes, void *stream) {
    P3D_REQUIRE(B >= 0 && C >= 0 && m >= 0 && n >= 0, "three_interpolate_bwd: negative dimension");
    if (B == 0 || C == 0 || m == 0) return P3D_OK;
    P3D_REQUIRE(dfeat, "three_interpolate_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        cudaMemsetAsync(dfeat, 0, (size_t)B * C * m * sizeof(float), st);
        return P3D_OK;
    }
    P3D_REQUIRE(dout && idx && weight, "three_interpolate_bwd: null pointer");
    const size_t need = p3d_three_interp_bwd_ws_bytes(B, m, n);
    if (need > 0 && ws != nullptr && ws_bytes >= need)
        return launch_bwd_sorted<true>(dout, weight, idx, B, C, m, (int64_t)n * 3, dfeat, (int64_t)C * n, 0, ws, st);
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
    const int64_t grid = (n + GR_THREADS / 32 - 1) / (GR_THREADS / 32);
    neighbor_max_kernel<<<(unsigned)grid, GR_THREADS, 0, (cudaStream_t)stream>>>(x, neighbors, n, N0, M, C, out, arg);
    return finish_launch("neighbor_max");
}
