// neighbors.cu -- ball query (dense + batch/CSR), knn and three_nn for sm_100a.
//
// Replaces tp.ball_query (core/spatial_ops/neighbour_finder.py:35-37,164), torch_cluster.knn via
// torch_geometric.nn.knn (neighbour_finder.py:47, interpolate.py:25,70) and tp.three_nn
// (core/base_conv/dense.py:136).  Semantics: SURVEY.md Appendix A.2/A.3/A.5/A.6.
//
// Design: candidates are swept in index order through a shared-memory tile (SoA x|y|z, one
// coalesced pass over the cloud per CTA); a WARP owns a query and tests 32 candidates per
// step (conflict-free LDS), `__ballot_sync` + popc gives every hit its output slot in
// candidate order ("first K in index order" survives the parallel scan), hits are written
// straight to their final slot and a warp stops as soon as its row is full.  A CTA carries
// 8*QPW queries per sweep so a tile loaded once is reused 8*QPW times from shared memory.
// Cloud ranges of the batch (partial-dense / CSR) form are found on the device with a 32-ary
// warp search over the sorted batch vector: no host round trip.
#include <math_constants.h>

#include <cstdlib>

#include "common.cuh"
#include "neighbors_common.cuh"

namespace p3d {

constexpr int NB_TILE = 2048;  // candidate points per shared-memory tile (24 KB)
constexpr int NB_WARPS = 8;

__device__ __forceinline__ void load_tile(const float *__restrict__ x, int64_t tb, int tn, float *sx, float *sy,
                                          float *sz, int tid, int nthreads) {
    const float *__restrict__ src = x + 3 * tb;
#pragma unroll 4
    for (int f = tid; f < 3 * tn; f += nthreads) {
        const float v = __ldg(src + f);
        const int p = f / 3, c = f - 3 * p;
        float *dst = c == 0 ? sx : (c == 1 ? sy : sz);
        dst[p] = v;
    }
}

// ------------------------------------------------------------------ ball query -----------
template <bool DENSE, int QPW>
__global__ void __launch_bounds__(NB_WARPS * 32)
ball_query_kernel(const float *__restrict__ x, const float *__restrict__ y, const int64_t *__restrict__ batch_x,
                  const int64_t *__restrict__ batch_y, int64_t Nx, int64_t Ny, int64_t n_dense, int64_t s_dense,
                  float r2, int K, int strict, int64_t shadow, int64_t *__restrict__ idx,
                  float *__restrict__ dist2) {
    constexpr int QPC = NB_WARPS * QPW;
    __shared__ float4 tile[NB_TILE];  // packed xyz_: one LDS.128 per candidate and lane
    __shared__ RangeCache rc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * QPC;
    const int64_t q1 = (q0 + QPC < Ny) ? q0 + QPC : Ny;
    // d <= r2  <=>  d < nextafter(r2): one compare for both membership rules
    const float r2x = strict ? r2 : __uint_as_float(__float_as_uint(r2) + 1u);

    cta_ranges<DENSE>(q0, q1, batch_x, batch_y, Nx, rc, warp, lane);

    float qx[QPW], qy[QPW], qz[QPW];
    int64_t x0[QPW], x1[QPW];
    int cnt[QPW], first[QPW];
    int64_t X0 = INT64_MAX, X1 = 0;
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        const int64_t q = q0 + warp * QPW + u;
        cnt[u] = 0;
        first[u] = 0;
        if (q < q1) {
            qx[u] = __ldg(y + 3 * q);
            qy[u] = __ldg(y + 3 * q + 1);
            qz[u] = __ldg(y + 3 * q + 2);
            query_range<DENSE>(q, batch_x, batch_y, Nx, n_dense, s_dense, rc, lane, x0[u], x1[u]);
        } else {
            qx[u] = qy[u] = qz[u] = 0.f;
            x0[u] = x1[u] = 0;
            cnt[u] = K;  // nothing to do
        }
    }
    // CTA sweep range = union of the ranges of its first and last query (clouds are contiguous)
    {
        int64_t a, b;
        query_range<DENSE>(q0, batch_x, batch_y, Nx, n_dense, s_dense, rc, lane, a, b);
        X0 = a;
        query_range<DENSE>(q1 - 1, batch_x, batch_y, Nx, n_dense, s_dense, rc, lane, a, b);
        X1 = b;
    }

    bool first_tile = true;
    for (int64_t tb = X0; tb < X1; tb += NB_TILE) {
        bool mydone = true;
#pragma unroll
        for (int u = 0; u < QPW; ++u) mydone = mydone && (cnt[u] >= K || tb >= x1[u]);
        if (!first_tile) {
            if (__syncthreads_and(mydone)) break;  // also: previous tile fully consumed
        }
        first_tile = false;
        const int tn = (int)((X1 - tb < NB_TILE) ? (X1 - tb) : NB_TILE);
        {
            const float *__restrict__ src = x + 3 * tb;
#pragma unroll 2
            for (int p = tid; p < tn; p += NB_WARPS * 32)
                tile[p] = make_float4(__ldg(src + 3 * p), __ldg(src + 3 * p + 1), __ldg(src + 3 * p + 2), 0.f);
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < QPW; ++u) {
            if (cnt[u] >= K) continue;
            const int64_t q = q0 + warp * QPW + u;
            const int lo = (int)((x0[u] > tb ? x0[u] : tb) - tb);
            const int hi = (int)((x1[u] < tb + tn ? x1[u] : tb + tn) - tb);
            // index reported for tile slot i (32-bit arithmetic inside the loop): dense mode is cloud-local
            const int ibase = (int)(tb - (DENSE ? x0[u] : 0));
            int64_t *__restrict__ row = idx + q * K;
            float *__restrict__ drow = dist2 ? dist2 + q * K : nullptr;
            int c = cnt[u];
            // 4 x 32 candidates per trip: the four distance chains and ballots are independent, only the
            // slot assignment is sequential (scanning a few candidates past the K-th hit is harmless)
            for (int i0 = lo; i0 < hi && c < K; i0 += 128) {
                float d[4];
                bool hit[4];
                unsigned m[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int i = i0 + 32 * v + lane;
                    const bool ok = i < hi;
                    const float4 cnd = tile[ok ? i : lo];
                    d[v] = sqdist3(cnd.x, cnd.y, cnd.z, qx[u], qy[u], qz[u]);
                    hit[v] = ok && d[v] < r2x;
                    m[v] = __ballot_sync(0xffffffffu, hit[v]);
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (m[v]) {
                        const int pos = c + __popc(m[v] & lanemask_lt());
                        if (hit[v] && pos < K) {
                            row[pos] = (int64_t)(ibase + i0 + 32 * v + lane);
                            if (drow) drow[pos] = d[v];
                        }
                        if (c == 0) first[u] = ibase + i0 + 32 * v + (__ffs(m[v]) - 1);
                        c += __popc(m[v]);
                    }
                }
            }
            cnt[u] = c;
        }
    }
    // padding: dense -> first hit (or 0 for an empty ball); batch -> shadow index; dist2 -> -1
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        const int64_t q = q0 + warp * QPW + u;
        if (q >= q1) continue;
        const int c = cnt[u] < K ? cnt[u] : K;
        const int64_t padv = DENSE ? (c > 0 ? (int64_t)first[u] : 0) : shadow;
        for (int e = c + lane; e < K; e += 32) {
            idx[q * K + e] = padv;
            if (dist2) dist2[q * K + e] = -1.f;
        }
    }
}

// ------------------------------------------------------------------------- knn -----------
// Sorted top-k list per query in shared memory (warp-private), ascending d2, ties keep the
// lower index (strict '<' admission against the current k-th distance, insertion after all
// entries <= the new distance; candidates arrive in index order).
__device__ __forceinline__ void knn_insert(float *dl, int *il, int k, int &len, float &tau, float nd, int ni,
                                           int lane) {
    int pos = 0;
    for (int c = 0; c < len; c += 32) {
        const int e = c + lane;
        const float v = e < len ? dl[e] : CUDART_INF_F;
        pos += __popc(__ballot_sync(0xffffffffu, v <= nd));
    }
    const int newlen = len < k ? len + 1 : k;
    for (int top = newlen - 1; top > pos; top -= 32) {  // elements [pos, newlen-1) move up by one
        const int e = top - 1 - lane;
        float v = 0.f;
        int iv = 0;
        if (e >= pos) { v = dl[e]; iv = il[e]; }
        __syncwarp();
        if (e >= pos) { dl[e + 1] = v; il[e + 1] = iv; }
        __syncwarp();
    }
    if (lane == 0) { dl[pos] = nd; il[pos] = ni; }
    __syncwarp();
    len = newlen;
    tau = len == k ? dl[k - 1] : CUDART_INF_F;
}

template <bool DENSE, int QPW>
__global__ void __launch_bounds__(NB_WARPS * 32)
knn_kernel(const float *__restrict__ x, const float *__restrict__ y, const int64_t *__restrict__ batch_x,
           const int64_t *__restrict__ batch_y, int64_t Nx, int64_t Ny, int64_t n_dense, int64_t s_dense, int k,
           int sqrt_out, int64_t *__restrict__ col, float *__restrict__ dout) {
    constexpr int QPC = NB_WARPS * QPW;
    __shared__ float sx[NB_TILE], sy[NB_TILE], sz[NB_TILE];
    __shared__ RangeCache rc;
    extern __shared__ __align__(16) unsigned char dyn[];
    float *dl_all = reinterpret_cast<float *>(dyn);            // [QPC][k]
    int *il_all = reinterpret_cast<int *>(dl_all + QPC * k);  // [QPC][k] index relative to X0
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * QPC;
    const int64_t q1 = (q0 + QPC < Ny) ? q0 + QPC : Ny;

    cta_ranges<DENSE>(q0, q1, batch_x, batch_y, Nx, rc, warp, lane);

    float qx[QPW], qy[QPW], qz[QPW], tau[QPW];
    int64_t x0[QPW], x1[QPW];
    int len[QPW];
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        const int64_t q = q0 + warp * QPW + u;
        len[u] = 0;
        tau[u] = CUDART_INF_F;
        if (q < q1) {
            qx[u] = __ldg(y + 3 * q);
            qy[u] = __ldg(y + 3 * q + 1);
            qz[u] = __ldg(y + 3 * q + 2);
            query_range<DENSE>(q, batch_x, batch_y, Nx, n_dense, s_dense, rc, lane, x0[u], x1[u]);
        } else {
            qx[u] = qy[u] = qz[u] = 0.f;
            x0[u] = x1[u] = 0;
        }
    }
    int64_t X0, X1;
    {
        int64_t a, b;
        query_range<DENSE>(q0, batch_x, batch_y, Nx, n_dense, s_dense, rc, lane, a, b);
        X0 = a;
        query_range<DENSE>(q1 - 1, batch_x, batch_y, Nx, n_dense, s_dense, rc, lane, a, b);
        X1 = b;
    }

    for (int64_t tb = X0; tb < X1; tb += NB_TILE) {
        const int tn = (int)((X1 - tb < NB_TILE) ? (X1 - tb) : NB_TILE);
        if (tb != X0) __syncthreads();
        load_tile(x, tb, tn, sx, sy, sz, tid, NB_WARPS * 32);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < QPW; ++u) {
            float *dl = dl_all + (warp * QPW + u) * k;
            int *il = il_all + (warp * QPW + u) * k;
            int lo = (int)((x0[u] > tb ? x0[u] : tb) - tb);
            int hi = (int)((x1[u] < tb + tn ? x1[u] : tb + tn) - tb);
            int ln = len[u];
            float t = tau[u];
            for (int i0 = lo; i0 < hi; i0 += 32) {
                const int i = i0 + lane;
                const bool ok = i < hi;
                const int ii = ok ? i : lo;
                const float d = sqdist3(sx[ii], sy[ii], sz[ii], qx[u], qy[u], qz[u]);
                unsigned m = __ballot_sync(0xffffffffu, ok && d < t);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const float nd = __shfl_sync(0xffffffffu, d, src);
                    if (nd < t) knn_insert(dl, il, k, ln, t, nd, (int)(tb - X0) + i0 + src, lane);
                }
            }
            len[u] = ln;
            tau[u] = t;
        }
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        const int64_t q = q0 + warp * QPW + u;
        if (q >= q1) continue;
        const float *dl = dl_all + (warp * QPW + u) * k;
        const int *il = il_all + (warp * QPW + u) * k;
        const int64_t off = DENSE ? x0[u] : 0;
        for (int e = lane; e < k; e += 32) {
            const bool have = e < len[u];
            col[q * k + e] = have ? (X0 + il[e] - off) : -1;
            if (dout) {
                float v = have ? dl[e] : -1.f;
                if (have && sqrt_out) v = sqrtf(v);
                dout[q * k + e] = v;
            }
        }
    }
}


// -------------------------------------------------------------------- three_nn -----------
// One THREAD per unknown point; the known points stream through a shared-memory tile of packed
// float4 (one broadcast LDS.128 per candidate for the whole warp); the best three live in
// registers with the strict '<' cascade of the reference kernel (ties keep the lower index).
constexpr int TNN_THREADS = 128;
constexpr int TNN_TILE = 1024;

// SPLIT consecutive lanes share one unknown point and scan interleaved candidates (lane s takes candidates
// s, s + SPLIT, ...); their sorted triples are merged by (distance, index) -- the order the sequential strict-<
// cascade produces -- with two shuffle rounds.  Small problems (C2: 8k-32k unknown points) would otherwise put
// 1-2 CTAs on an SM and run at ~10 % occupancy.
template <int SPLIT>
__global__ void __launch_bounds__(TNN_THREADS)
three_nn_kernel(const float *__restrict__ unknown, const float *__restrict__ known, int n, int m,
                float *__restrict__ dist, int64_t *__restrict__ idx, float *__restrict__ weight) {
    __shared__ float4 tile[TNN_TILE];
    const int b = blockIdx.y;
    const int q = (blockIdx.x * TNN_THREADS + threadIdx.x) / SPLIT;
    const int sub = threadIdx.x % SPLIT;
    const bool live = q < n;
    const float *__restrict__ u = unknown + ((int64_t)b * n + (live ? q : 0)) * 3;
    const float qx = __ldg(u), qy = __ldg(u + 1), qz = __ldg(u + 2);
    const float *__restrict__ kb = known + (int64_t)b * m * 3;
    float b1 = CUDART_INF_F, b2 = CUDART_INF_F, b3 = CUDART_INF_F;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int tb = 0; tb < m; tb += TNN_TILE) {
        const int tn = (m - tb < TNN_TILE) ? (m - tb) : TNN_TILE;
        __syncthreads();
        for (int i = threadIdx.x; i < tn; i += TNN_THREADS) {
            const float *__restrict__ c = kb + (int64_t)(tb + i) * 3;
            tile[i] = make_float4(__ldg(c), __ldg(c + 1), __ldg(c + 2), 0.f);
        }
        __syncthreads();
#pragma unroll 4
        for (int i = sub; i < tn; i += SPLIT) {
            const float4 c = tile[i];
            const float d = sqdist3(qx, qy, qz, c.x, c.y, c.z);
            if (d < b3) {  // within one lane candidates arrive in index order: strict < keeps the lower index on ties
                const int k = tb + i;
                if (d < b1) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
                else if (d < b2) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
                else { b3 = d; i3 = k; }
            }
        }
    }
    if (SPLIT > 1) {
        auto lt = [](float d, int k, float e, int l) { return d < e || (d == e && k < l); };
#pragma unroll
        for (int o = 1; o < SPLIT; o <<= 1) {
            float pd[3];
            int pi[3];
            pd[0] = __shfl_xor_sync(0xffffffffu, b1, o), pi[0] = __shfl_xor_sync(0xffffffffu, i1, o);
            pd[1] = __shfl_xor_sync(0xffffffffu, b2, o), pi[1] = __shfl_xor_sync(0xffffffffu, i2, o);
            pd[2] = __shfl_xor_sync(0xffffffffu, b3, o), pi[2] = __shfl_xor_sync(0xffffffffu, i3, o);
#pragma unroll
            for (int t = 0; t < 3; ++t) {
                const float d = pd[t];
                const int k = pi[t];
                if (d < CUDART_INF_F && lt(d, k, b3, i3)) {
                    if (lt(d, k, b1, i1)) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
                    else if (lt(d, k, b2, i2)) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
                    else { b3 = d; i3 = k; }
                }
            }
        }
    }
    if (live && sub == 0) {
        float *__restrict__ dd = dist + ((int64_t)b * n + q) * 3;
        int64_t *__restrict__ ii = idx + ((int64_t)b * n + q) * 3;
        const float d1 = __fsqrt_rn(b1), d2 = __fsqrt_rn(b2), d3 = __fsqrt_rn(b3);
        dd[0] = d1; dd[1] = d2; dd[2] = d3;
        ii[0] = i1; ii[1] = i2; ii[2] = i3;
        if (weight) {  // DenseFPModule.conv, core/base_conv/dense.py:137-139
            const float r1 = __fdiv_rn(1.0f, __fadd_rn(d1, 1e-8f)), r2 = __fdiv_rn(1.0f, __fadd_rn(d2, 1e-8f));
            const float r3 = __fdiv_rn(1.0f, __fadd_rn(d3, 1e-8f));
            const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
            float *__restrict__ ww = weight + ((int64_t)b * n + q) * 3;
            ww[0] = __fdiv_rn(r1, norm); ww[1] = __fdiv_rn(r2, norm); ww[2] = __fdiv_rn(r3, norm);
        }
    }
}


// ------------------------------------------------ knn, k <= 32: warp per query, list in lanes ----
// Candidate sweep through a shared-memory tile, 32 candidates per warp step.  The sorted best-k list of
// a query lives in REGISTERS ACROSS THE LANES of its warp (lane e = rank e), so an insertion is one
// ballot + one shfl_up + a select (~10 warp instructions) and never serialises other queries.
// The tile is stored TRANSPOSED (candidate j at [j % 64][j / 64], rows padded to 33 floats): step t
// tests candidates t, 64+t, 128+t, ... instead of 32 consecutive ones.  Real clouds are spatially ordered
// (LiDAR rings, voxel order), so consecutive candidates approach a query monotonically and EVERY one of
// them would improve the running k-th distance; the strided order sees them like a random permutation
// (k ln(N/k) insertions instead of O(N)).  Entries are ordered by (distance, index), so the result is
// still "ascending distance, ties to the lower index" regardless of the visiting order.
constexpr int KN_ROWS = 64, KN_PITCH = 33, KN_TILE = KN_ROWS * 32;

// one warp step: every lane offers candidate (d, base + lane * KN_ROWS + s) if `ok`; admitted ones are merged
// into the lane-resident list (d_l, i_l), kept sorted by (distance, index); (t, ti) = entry at rank k-1
__device__ __forceinline__ void knn_lane_step(float d, bool ok, int base, int s, int k, int lane, float &d_l,
                                              int &i_l, float &t, int &ti) {
    unsigned m = __ballot_sync(0xffffffffu, ok && d <= t);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float nd = __shfl_sync(0xffffffffu, d, src);
        const int ni = base + src * KN_ROWS + s;
        if (nd < t || (nd == t && ni < ti)) {  // warp-uniform: lands before rank k-1
            // ranks that move up by one: entries that sort after (nd, ni)
            const unsigned gt = __ballot_sync(0xffffffffu, d_l > nd || (d_l == nd && i_l > ni));
            const float up_d = __shfl_up_sync(0xffffffffu, d_l, 1);
            const int up_i = __shfl_up_sync(0xffffffffu, i_l, 1);
            const int first = __ffs(gt) - 1;
            if (lane == first) { d_l = nd; i_l = ni; }
            else if (lane > first) { d_l = up_d; i_l = up_i; }
            t = __shfl_sync(0xffffffffu, d_l, k - 1);
            ti = __shfl_sync(0xffffffffu, i_l, k - 1);
            m &= __ballot_sync(0xffffffffu, d <= t);  // drop candidates the tighter bound rejects
        }
    }
}

template <bool DENSE, int QPW>
__global__ void __launch_bounds__(NB_WARPS * 32)
knn_lane_kernel(const float *__restrict__ x, const float *__restrict__ y, const int64_t *__restrict__ batch_x,
                const int64_t *__restrict__ batch_y, int64_t Nx, int64_t Ny, int64_t n_dense, int64_t s_dense, int k,
                int64_t *__restrict__ col, float *__restrict__ dout) {
    constexpr int QPC = NB_WARPS * QPW;
    __shared__ float sx[KN_ROWS * KN_PITCH], sy[KN_ROWS * KN_PITCH], sz[KN_ROWS * KN_PITCH];
    __shared__ RangeCache rc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * QPC;
    const int64_t q1 = (q0 + QPC < Ny) ? q0 + QPC : Ny;
    cta_ranges<DENSE>(q0, q1, batch_x, batch_y, Nx, rc, warp, lane);

    float qx[QPW], qy[QPW], qz[QPW], tau[QPW], bd[QPW];
    int bi[QPW], taui[QPW];
    int64_t x0[QPW], x1[QPW];
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        const int64_t q = q0 + warp * QPW + u;
        tau[u] = CUDART_INF_F;
        taui[u] = 0x7fffffff;
        bd[u] = CUDART_INF_F;  // lanes >= k never matter: the admission test uses rank k-1
        bi[u] = 0x7fffffff;    // empty slot (sorts after every real entry)
        if (q < q1) {
            qx[u] = __ldg(y + 3 * q); qy[u] = __ldg(y + 3 * q + 1); qz[u] = __ldg(y + 3 * q + 2);
            query_range<DENSE>(q, batch_x, batch_y, Nx, n_dense, s_dense, rc, lane, x0[u], x1[u]);
        } else {
            qx[u] = qy[u] = qz[u] = 0.f;
            x0[u] = x1[u] = 0;
        }
    }
    int64_t X0, X1;
    {
        int64_t a, b;
        query_range<DENSE>(q0, batch_x, batch_y, Nx, n_dense, s_dense, rc, lane, a, b);
        X0 = a;
        query_range<DENSE>(q1 - 1, batch_x, batch_y, Nx, n_dense, s_dense, rc, lane, a, b);
        X1 = b;
    }
    // pass 0: step 0 of EVERY tile (candidates tb + 64 * lane, a 1/64 subsample of the cloud) straight from
    // global memory.  It costs 1/64 of the sweep and leaves a k-th distance bound that only ~64 k candidates
    // of the whole cloud can beat, so the main sweep (steps 1..63 of every tile) rarely inserts.
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        float t = tau[u], d_l = bd[u];
        int ti = taui[u], i_l = bi[u];
        for (int64_t tb = X0; tb < X1; tb += KN_TILE) {
            const int64_t cand = tb + (int64_t)lane * KN_ROWS;
            const bool ok = cand >= x0[u] && cand < x1[u];
            float d = 0.f;
            if (ok) d = sqdist3(__ldg(x + 3 * cand), __ldg(x + 3 * cand + 1), __ldg(x + 3 * cand + 2), qx[u], qy[u], qz[u]);
            knn_lane_step(d, ok, (int)(tb - X0), 0, k, lane, d_l, i_l, t, ti);
        }
        tau[u] = t; taui[u] = ti; bd[u] = d_l; bi[u] = i_l;
    }
    for (int64_t tb = X0; tb < X1; tb += KN_TILE) {
        const int tn = (int)((X1 - tb < KN_TILE) ? (X1 - tb) : KN_TILE);
        if (tb != X0) __syncthreads();
        {
            const float *__restrict__ src = x + 3 * tb;
#pragma unroll 4
            for (int f = tid; f < 3 * tn; f += NB_WARPS * 32) {
                const float v = __ldg(src + f);
                const int p = f / 3, c = f - 3 * p;
                float *dst = c == 0 ? sx : (c == 1 ? sy : sz);
                dst[(p & (KN_ROWS - 1)) * KN_PITCH + (p >> 6)] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < QPW; ++u) {
            const int lo = (int)((x0[u] > tb ? x0[u] : tb) - tb);
            const int hi = (int)((x1[u] < tb + tn ? x1[u] : tb + tn) - tb);
            if (lo >= hi) continue;  // warp-uniform
            float t = tau[u], d_l = bd[u];
            int ti = taui[u], i_l = bi[u];
            const int base = (int)(tb - X0);
#pragma unroll 2
            for (int s = 1; s < KN_ROWS; ++s) {  // step 0 was pass 0
                const int j = lane * KN_ROWS + s;  // candidate of this lane at this step
                const bool ok = j >= lo && j < hi;
                const int at = s * KN_PITCH + lane;
                const float d = sqdist3(sx[at], sy[at], sz[at], qx[u], qy[u], qz[u]);
                knn_lane_step(d, ok, base, s, k, lane, d_l, i_l, t, ti);
            }
            tau[u] = t; taui[u] = ti; bd[u] = d_l; bi[u] = i_l;
        }
    }
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        const int64_t q = q0 + warp * QPW + u;
        if (q >= q1) continue;
        const int64_t off = DENSE ? x0[u] : 0;
        if (lane < k) {
            const bool have = bi[u] != 0x7fffffff;
            col[q * k + lane] = have ? (X0 + bi[u] - off) : -1;
            if (dout) dout[q * k + lane] = have ? bd[u] : -1.f;
        }
    }
}

// ---------------------------------------------------- multi-scale dense ball query -----------
// PointNet++ MSG layers query the SAME (points, centroids) pair at several radii (pointnet2.yaml:95-130:
// r = [.1,.2,.4] with K = [32,64,128]).  One sweep serves all scales: every squared distance is computed once
// and tested against each radius; a scale stops taking hits when its row is full, the warp stops when all rows
// are.  Each scale's output is exactly what the single-scale kernel produces (same candidate order, same
// fp32 distance, same padding rule).
constexpr int BQ_MAX_SCALES = 4;

struct MultiScaleArgs {
    float r2x[BQ_MAX_SCALES];  // exclusive upper bound on d2 (strict: r2, otherwise nextafter(r2))
    int K[BQ_MAX_SCALES];
    int64_t *idx[BQ_MAX_SCALES];
    int ns;
};

template <int NS>
__global__ void __launch_bounds__(NB_WARPS * 32)
ball_query_multi_kernel(const float *__restrict__ x, const float *__restrict__ y, int64_t Ny, int64_t n_dense,
                        int64_t s_dense, MultiScaleArgs a) {
    __shared__ float4 tile[NB_TILE];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * NB_WARPS;
    const int64_t q1 = (q0 + NB_WARPS < Ny) ? q0 + NB_WARPS : Ny;
    const int64_t q = q0 + warp;
    const bool live = q < q1;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    int64_t x0 = 0, x1 = 0;
    int cnt[NS], first[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) cnt[s] = live ? 0 : a.K[s], first[s] = 0;
    if (live) {
        qx = __ldg(y + 3 * q), qy = __ldg(y + 3 * q + 1), qz = __ldg(y + 3 * q + 2);
        x0 = (q / s_dense) * n_dense, x1 = x0 + n_dense;
    }
    const int64_t X0 = (q0 / s_dense) * n_dense, X1 = ((q1 - 1) / s_dense) * n_dense + n_dense;
    bool first_tile = true;
    for (int64_t tb = X0; tb < X1; tb += NB_TILE) {
        bool mydone = tb >= x1;
        if (!mydone) {
            mydone = true;
#pragma unroll
            for (int s = 0; s < NS; ++s) mydone = mydone && cnt[s] >= a.K[s];
        }
        if (!first_tile) {
            if (__syncthreads_and(mydone)) break;
        }
        first_tile = false;
        const int tn = (int)((X1 - tb < NB_TILE) ? (X1 - tb) : NB_TILE);
        {
            const float *__restrict__ src = x + 3 * tb;
#pragma unroll 2
            for (int p = tid; p < tn; p += NB_WARPS * 32)
                tile[p] = make_float4(__ldg(src + 3 * p), __ldg(src + 3 * p + 1), __ldg(src + 3 * p + 2), 0.f);
        }
        __syncthreads();
        if (mydone) continue;
        const int lo = (int)((x0 > tb ? x0 : tb) - tb);
        const int hi = (int)((x1 < tb + tn ? x1 : tb + tn) - tb);
        const int ibase = (int)(tb - x0);
        for (int i0 = lo; i0 < hi; i0 += 128) {
            bool open = false;
#pragma unroll
            for (int s = 0; s < NS; ++s) open = open || cnt[s] < a.K[s];
            if (!open) break;
            float d[4];
            bool ok[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const int i = i0 + 32 * v + lane;
                ok[v] = i < hi;
                const float4 cnd = tile[ok[v] ? i : lo];
                d[v] = sqdist3(cnd.x, cnd.y, cnd.z, qx, qy, qz);
            }
#pragma unroll
            for (int s = 0; s < NS; ++s) {
                int c = cnt[s];
                if (c >= a.K[s]) continue;  // warp-uniform
                int64_t *__restrict__ row = a.idx[s] + q * a.K[s];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const bool hit = ok[v] && d[v] < a.r2x[s];
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (m) {
                        const int pos = c + __popc(m & lanemask_lt());
                        if (hit && pos < a.K[s]) row[pos] = (int64_t)(ibase + i0 + 32 * v + lane);
                        if (c == 0) first[s] = ibase + i0 + 32 * v + (__ffs(m) - 1);
                        c += __popc(m);
                    }
                }
                cnt[s] = c;
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int s = 0; s < NS; ++s) {  // padding: first hit, or 0 for an empty ball
        const int c = cnt[s] < a.K[s] ? cnt[s] : a.K[s];
        const int64_t padv = c > 0 ? (int64_t)first[s] : 0;
        for (int e = c + lane; e < a.K[s]; e += 32) a.idx[s][q * a.K[s] + e] = padv;
    }
}

template <bool DENSE>
static int launch_ball_query(const float *x, const float *y, const int64_t *bx, const int64_t *by, int64_t Nx,
                             int64_t Ny, int64_t n_dense, int64_t s_dense, float radius, int K, int strict,
                             int64_t shadow, int64_t *idx, float *dist2, cudaStream_t st) {
    if (Ny == 0 || K == 0) return P3D_OK;
    const float r2 = radius * radius;
    int qpw = Ny >= (int64_t)NB_WARPS * 4 * 2 * kNumSMs ? 4 : 1;
    if (qpw == 4) {
        const int64_t grid = (Ny + NB_WARPS * 4 - 1) / (NB_WARPS * 4);
        ball_query_kernel<DENSE, 4><<<(unsigned)grid, NB_WARPS * 32, 0, st>>>(
            x, y, bx, by, Nx, Ny, n_dense, s_dense, r2, K, strict, shadow, idx, dist2);
    } else if (qpw == 2) {
        const int64_t grid = (Ny + NB_WARPS * 2 - 1) / (NB_WARPS * 2);
        ball_query_kernel<DENSE, 2><<<(unsigned)grid, NB_WARPS * 32, 0, st>>>(
            x, y, bx, by, Nx, Ny, n_dense, s_dense, r2, K, strict, shadow, idx, dist2);
    } else {
        const int64_t grid = (Ny + NB_WARPS - 1) / NB_WARPS;
        ball_query_kernel<DENSE, 1><<<(unsigned)grid, NB_WARPS * 32, 0, st>>>(
            x, y, bx, by, Nx, Ny, n_dense, s_dense, r2, K, strict, shadow, idx, dist2);
    }
    return finish_launch("ball_query");
}

template <bool DENSE>
static int launch_knn(const float *x, const float *y, const int64_t *bx, const int64_t *by, int64_t Nx, int64_t Ny,
                      int64_t n_dense, int64_t s_dense, int k, int sqrt_out, int64_t *col, float *d,
                      cudaStream_t st) {
    if (Ny == 0 || k == 0) return P3D_OK;
    if (k <= 32 && !sqrt_out) {
        if (Ny >= (int64_t)NB_WARPS * 4 * 2 * kNumSMs) {
            const unsigned grid = (unsigned)((Ny + NB_WARPS * 4 - 1) / (NB_WARPS * 4));
            knn_lane_kernel<DENSE, 4><<<grid, NB_WARPS * 32, 0, st>>>(x, y, bx, by, Nx, Ny, n_dense, s_dense, k, col, d);
        } else {
            const unsigned grid = (unsigned)((Ny + NB_WARPS - 1) / NB_WARPS);
            knn_lane_kernel<DENSE, 1><<<grid, NB_WARPS * 32, 0, st>>>(x, y, bx, by, Nx, Ny, n_dense, s_dense, k, col, d);
        }
        return finish_launch("knn_lane");
    }
    if (Ny >= (int64_t)NB_WARPS * 4 * 2 * kNumSMs && k <= 32) {
        const int64_t grid = (Ny + NB_WARPS * 4 - 1) / (NB_WARPS * 4);
        const size_t dyn = (size_t)NB_WARPS * 4 * k * 8;
        knn_kernel<DENSE, 4><<<(unsigned)grid, NB_WARPS * 32, dyn, st>>>(x, y, bx, by, Nx, Ny, n_dense, s_dense, k,
                                                                        sqrt_out, col, d);
    } else {
        const int64_t grid = (Ny + NB_WARPS - 1) / NB_WARPS;
        const size_t dyn = (size_t)NB_WARPS * k * 8;
        knn_kernel<DENSE, 1><<<(unsigned)grid, NB_WARPS * 32, dyn, st>>>(x, y, bx, by, Nx, Ny, n_dense, s_dense, k,
                                                                        sqrt_out, col, d);
    }
    return finish_launch("knn");
}

}  // namespace p3d

using namespace p3d;

P3D_EXPORT int p3d_ball_query_dense_f32(const float *x, const float *y, int B, int N, int S, float radius,
                                        int nsample, int strict, int64_t *idx, float *dist2, void *stream) {
    P3D_REQUIRE(B >= 0 && N >= 0 && S >= 0 && nsample >= 0, "ball_query_dense: negative dimension");
    if (B == 0 || S == 0 || nsample == 0) return P3D_OK;
    P3D_REQUIRE(x && y && idx, "ball_query_dense: null pointer");
    return launch_ball_query<true>(x, y, nullptr, nullptr, (int64_t)B * N, (int64_t)B * S, N, S, radius, nsample,
                                   strict, 0, idx, dist2, (cudaStream_t)stream);
}

P3D_EXPORT int p3d_ball_query_dense_multi_f32(const float *x, const float *y, int B, int N, int S, int num_scales,
                                              const float *h_radii, const int *h_nsamples, int strict,
                                              int64_t *const *h_idx, void *stream) {
    P3D_REQUIRE(B >= 0 && N >= 0 && S >= 0, "ball_query_dense_multi: negative dimension");
    P3D_REQUIRE(num_scales >= 1 && num_scales <= BQ_MAX_SCALES, "ball_query_dense_multi: 1..%d scales", BQ_MAX_SCALES);
    P3D_REQUIRE(h_radii && h_nsamples && h_idx, "ball_query_dense_multi: null host array");
    if (B == 0 || S == 0) return P3D_OK;
    P3D_REQUIRE(x && y, "ball_query_dense_multi: null pointer");
    MultiScaleArgs a;
    a.ns = num_scales;
    for (int s = 0; s < BQ_MAX_SCALES; ++s) {
        const int t = s < num_scales ? s : 0;  // unused slots repeat scale 0 (never read)
        P3D_REQUIRE(h_nsamples[t] >= 1 && h_idx[t], "ball_query_dense_multi: scale %d: nsample >= 1 and idx required", t);
        const float r2 = h_radii[t] * h_radii[t];
        a.r2x[s] = strict ? r2 : __builtin_nextafterf(r2, __builtin_inff());
        a.K[s] = h_nsamples[t];
        a.idx[s] = h_idx[t];
    }
    const int64_t Ny = (int64_t)B * S;
    const unsigned grid = (unsigned)((Ny + NB_WARPS - 1) / NB_WARPS);
    cudaStream_t st = (cudaStream_t)stream;
    switch (num_scales) {
        case 1: ball_query_multi_kernel<1><<<grid, NB_WARPS * 32, 0, st>>>(x, y, Ny, N, S, a); break;
        case 2: ball_query_multi_kernel<2><<<grid, NB_WARPS * 32, 0, st>>>(x, y, Ny, N, S, a); break;
        case 3: ball_query_multi_kernel<3><<<grid, NB_WARPS * 32, 0, st>>>(x, y, Ny, N, S, a); break;
        default: ball_query_multi_kernel<4><<<grid, NB_WARPS * 32, 0, st>>>(x, y, Ny, N, S, a); break;
    }
    return finish_launch("ball_query_multi");
}

P3D_EXPORT int p3d_ball_query_batch_f32(const float *x, const float *y, const int64_t *batch_x,
                                        const int64_t *batch_y, int64_t Nx, int64_t Ny, float radius, int nsample,
                                        int strict, int64_t shadow, int64_t *idx, float *dist2, void *stream) {
    P3D_REQUIRE(Nx >= 0 && Ny >= 0 && nsample >= 0, "ball_query_batch: negative dimension");
    P3D_REQUIRE(Nx < ((int64_t)1 << 31), "ball_query_batch: more than 2^31 candidate points");
    if (Ny == 0 || nsample == 0) return P3D_OK;
    P3D_REQUIRE(y && idx && (x || Nx == 0), "ball_query_batch: null pointer");
    P3D_REQUIRE((batch_x == nullptr) == (batch_y == nullptr), "ball_query_batch: give both batch vectors or none");
    return launch_ball_query<false>(x, y, batch_x, batch_y, Nx, Ny, 0, 1, radius, nsample, strict, shadow, idx,
                                    dist2, (cudaStream_t)stream);
}

P3D_EXPORT int p3d_knn_batch_f32(const float *x, const float *y, const int64_t *batch_x, const int64_t *batch_y,
                                 int64_t Nx, int64_t Ny, int k, int64_t *col, float *d2, void *stream) {
    P3D_REQUIRE(Nx >= 0 && Ny >= 0, "knn: negative dimension");
    P3D_REQUIRE(k >= 1 && k <= 128, "knn: k must be in [1, 128], got %d", k);
    if (Ny == 0) return P3D_OK;
    P3D_REQUIRE(y && col && (x || Nx == 0), "knn: null pointer");
    P3D_REQUIRE((batch_x == nullptr) == (batch_y == nullptr), "knn: give both batch vectors or none");
    return launch_knn<false>(x, y, batch_x, batch_y, Nx, Ny, 0, 1, k, 0, col, d2, (cudaStream_t)stream);
}

P3D_EXPORT int p3d_knn_dense_f32(const float *x, const float *y, int B, int N, int M, int k, int64_t *idx, float *d2,
                                 void *stream) {
    P3D_REQUIRE(B >= 0 && N >= 0 && M >= 0, "knn_dense: negative dimension");
    P3D_REQUIRE(k >= 1 && k <= 128, "knn_dense: k must be in [1, 128], got %d", k);
    if (B == 0 || M == 0) return P3D_OK;
    P3D_REQUIRE(y && idx && (x || N == 0), "knn_dense: null pointer");
    return launch_knn<true>(x, y, nullptr, nullptr, (int64_t)B * N, (int64_t)B * M, N, M, k, 0, idx, d2,
                            (cudaStream_t)stream);
}

P3D_EXPORT int p3d_three_nn_f32(const float *unknown, const float *known, int B, int n, int m, float *dist,
                                int64_t *idx, float *weight, void *stream) {
    P3D_REQUIRE(B >= 0 && n >= 0, "three_nn: negative dimension");
    P3D_REQUIRE(m >= 3, "three_nn: needs at least 3 known points, got %d", m);
    if (B == 0 || n == 0) return P3D_OK;
    P3D_REQUIRE(unknown && known && dist && idx, "three_nn: null pointer");
    P3D_REQUIRE(B <= 65535, "three_nn: batch size must be <= 65535");
    if ((int64_t)B * n < (int64_t)kNumSMs * 512) {  // too few unknown points to fill the GPU one thread each
        dim3 grid((unsigned)(((int64_t)n * 4 + TNN_THREADS - 1) / TNN_THREADS), (unsigned)B);
        three_nn_kernel<4><<<grid, TNN_THREADS, 0, (cudaStream_t)stream>>>(unknown, known, n, m, dist, idx, weight);
    } else {
        dim3 grid((unsigned)((n + TNN_THREADS - 1) / TNN_THREADS), (unsigned)B);
        three_nn_kernel<1><<<grid, TNN_THREADS, 0, (cudaStream_t)stream>>>(unknown, known, n, m, dist, idx, weight);
    }
    return finish_launch("three_nn");
}
