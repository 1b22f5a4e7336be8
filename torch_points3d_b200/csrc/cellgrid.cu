// cellgrid.cu -- cell-list accelerated partial-dense ball query and knn for sm_100a.
//
// Same results as the sweep kernels of neighbors.cu (tp.ball_query partial_dense, core/spatial_ops/
// neighbour_finder.py:34-37; torch_cluster.knn, neighbour_finder.py:46-47; semantics SURVEY.md A.3 / A.6) -- bit
// for bit -- without testing every point of the cloud against every query:
//
//   build   every support point is hashed by (cloud, cell) -- cell = floor(p / size) on a global lattice, no bounding
//           boxes, no cloud count needed -- into an open-addressing table (64-bit keys, atomicCAS), counted, and
//           scattered into cell-contiguous order as packed float4 (x, y, z, index).  Four small launches.
//   ball    a warp owns a query: lanes 0..26 probe the 27 cells around it, the cells' ranges are walked as ONE
//           flattened candidate list (32 candidates per step, coalesced float4 loads), hits go to a per-warp buffer;
//           the reference order -- ascending candidate index, first K -- is restored at the end by rank counting
//           (indices are unique), so the visiting order never shows in the result.
//   knn     the same walk over growing cubes of cells; the best-k list lives in registers across the lanes, ordered
//           by (distance, index); the search stops when the k-th distance is inside the visited cube, and falls
//           back to a plain sweep of the cloud when it is not after a few rings (or when the cloud is small).
//
// Distances are the uncontracted fp32 (dx*dx + dy*dy) + dz*dz of every other kernel (sqdist3), so membership and
// ordering decisions are the oracle's.
#include <math_constants.h>

#include "common.cuh"
#include "neighbors_common.cuh"

namespace p3d {

constexpr int CG_WARPS = 8;
constexpr int CG_CAP = 256;            // hit buffer entries per warp (ball query)
constexpr int CG_MAX_K = 64;           // largest nsample served by the cell-list ball query
constexpr int CG_SMALL_CLOUD = 384;    // clouds up to this size are swept directly
constexpr int CG_COORD_MASK = 16383;   // cell coordinates live on a periodic lattice of 2^14 cells per axis
constexpr int CG_KNN_RMAX = 8;         // knn: rings of cells tried before the plain sweep (17^3 cells: still far cheaper)
constexpr uint64_t CG_EMPTY = 0ull;    // keys carry bit 63, so an all-zero table is an empty table

struct __align__(16) CgSlot {
    unsigned long long key;  // packed (cloud, cx, cy, cz) | bit 63, or CG_EMPTY
    int start;               // first row of the cell in `sorted`
    int cnt;                 // points in the cell
};

struct CellGrid {
    CgSlot *slots;    // [T] open-addressing hash table (one 16-byte load per probe)
    int *pslot;       // [Nx] table slot of every point
    int *prank;       // [Nx] arrival rank of the point inside its cell
    float4 *sorted;   // [Nx] (x, y, z, index bits), cell-contiguous
    int *cursor;      // [1] allocation cursor of `sorted`
    unsigned *bbox;   // [6] ordered-uint min xyz | max xyz (knn: cell size selection)
    uint32_t mask;    // T - 1
};

__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// lattice coordinate of v (saturating conversion: coordinates too large for the pitch all share one cell, which
// costs time, never correctness)
__device__ __forceinline__ int cg_coord(float v, float inv) { return __float2int_rd(v * inv); }
// the key wraps every coordinate modulo 2^14: cells a whole period apart share a slot (their points are rejected by
// the distance test), adjacent cells never do
__device__ __forceinline__ uint64_t cg_key(int64_t cloud, int cx, int cy, int cz) {
    return ((uint64_t)1 << 63) | ((uint64_t)(cloud & 0x1fffff) << 42) | ((uint64_t)(cx & CG_COORD_MASK) << 28) |
           ((uint64_t)(cy & CG_COORD_MASK) << 14) | (uint64_t)(cz & CG_COORD_MASK);
}
__device__ __forceinline__ uint32_t cg_hash(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return (uint32_t)k;
}

// knn: one lattice for the whole batch, sized from the batch's extent and the mean cloud size so that a cube of
// 27 cells holds a few times k points when the points lie on surfaces (two largest extents)
__device__ __forceinline__ float cg_knn_cell(const unsigned *bbox, const int64_t *batch_x, int64_t Nx, int k) {
    float e[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) e[a] = ord2f(bbox[3 + a]) - ord2f(bbox[a]);
    const float hi = fmaxf(e[0], fmaxf(e[1], e[2])), lo = fminf(e[0], fminf(e[1], e[2]));
    const float mid = e[0] + e[1] + e[2] - hi - lo;
    const float clouds = batch_x ? (float)(batch_x[Nx - 1] + 1) : 1.0f;
    const float per_cloud = fmaxf((float)Nx / fmaxf(clouds, 1.0f), 1.0f);
    float cell = 1.5f * sqrtf(fmaxf(hi * fmaxf(mid, 0.05f * hi), 0.f) * (float)k / per_cloud);
    if (!(cell > 0.f) || !isfinite(cell)) cell = 1.0f;
    return cell;
}

__global__ void __launch_bounds__(256)
cg_bbox_kernel(const float *__restrict__ x, int64_t Nx, unsigned *__restrict__ bbox) {
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < Nx; i += (int64_t)gridDim.x * 256) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const unsigned o = f2ord(__ldg(x + 3 * i + a));
            lo[a] = min(lo[a], o);
            hi[a] = max(hi[a], o);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const unsigned l = __reduce_min_sync(0xffffffffu, lo[a]), h = __reduce_max_sync(0xffffffffu, hi[a]);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(bbox + a, l);
            atomicMax(bbox + 3 + a, h);
        }
    }
}

// cell_size > 0: the lattice pitch (ball query); otherwise it is derived from the bounding box for k neighbours
__global__ void __launch_bounds__(256)
cg_insert_kernel(const float *__restrict__ x, const int64_t *__restrict__ batch_x, int64_t Nx, float cell_size, int k,
                 CellGrid g) {
    const float cell = cell_size > 0.f ? cell_size : cg_knn_cell(g.bbox, batch_x, Nx, k);
    const float inv = 1.0f / cell;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < Nx; i += (int64_t)gridDim.x * 256) {
        const float px = __ldg(x + 3 * i), py = __ldg(x + 3 * i + 1), pz = __ldg(x + 3 * i + 2);
        const uint64_t key = cg_key(batch_x ? batch_x[i] : 0, cg_coord(px, inv), cg_coord(py, inv), cg_coord(pz, inv));
        uint32_t h = cg_hash(key) & g.mask;
        while (true) {
            const unsigned long long prev = atomicCAS(&g.slots[h].key, (unsigned long long)CG_EMPTY, (unsigned long long)key);
            if (prev == CG_EMPTY || prev == key) break;
            h = (h + 1) & g.mask;
        }
        g.pslot[i] = (int)h;
        g.prank[i] = atomicAdd(&g.slots[h].cnt, 1);
    }
}

__global__ void __launch_bounds__(256)
cg_alloc_kernel(CellGrid g) {
    const uint32_t s = blockIdx.x * 256 + threadIdx.x;
    if (s > g.mask) return;
    const int c = g.slots[s].cnt;
    if (c > 0) g.slots[s].start = atomicAdd(g.cursor, c);
}

__global__ void __launch_bounds__(256)
cg_scatter_kernel(const float *__restrict__ x, int64_t Nx, CellGrid g) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < Nx; i += (int64_t)gridDim.x * 256) {
        const int row = g.slots[g.pslot[i]].start + g.prank[i];
        g.sorted[row] = make_float4(__ldg(x + 3 * i), __ldg(x + 3 * i + 1), __ldg(x + 3 * i + 2), __int_as_float((int)i));
    }
}

// (start, count) of the cell with this key, (0, 0) when no point fell into it
__device__ __forceinline__ void cg_lookup(const CellGrid &g, uint64_t key, int &st, int &n) {
    uint32_t h = cg_hash(key) & g.mask;
    st = 0;
    n = 0;
    while (true) {
        const uint4 v = __ldg(reinterpret_cast<const uint4 *>(g.slots + h));
        const uint64_t k = ((uint64_t)v.y << 32) | v.x;
        if (k == key) { st = (int)v.z; n = (int)v.w; return; }
        if (k == CG_EMPTY) return;
        h = (h + 1) & g.mask;
    }
}

// flattened position f of a candidate list made of up to 32 cell ranges: which range, and its row in `sorted`.
// pre[] (exclusive prefix sums of the counts, pre[32] = total) and st[] are this warp's shared-memory tables.
__device__ __forceinline__ int cg_flat_row(const int *pre, const int *st, int f) {
    int lo = 0;  // largest j with pre[j] <= f
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1)
        if (pre[lo + step] <= f) lo += step;
    return st[lo] + (f - pre[lo]);
}

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total) {
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
    }
    total = __shfl_sync(0xffffffffu, s, 31);
    return s - v;
}

// ------------------------------------------------------------------ ball query -----------
// per-warp hit buffer: candidate indices and squared distances in separate arrays (the index array is read four
// entries per LDS.128 by the rank counting below)
struct CgHits {
    int idx[CG_CAP + 4];
    float d2[CG_CAP];
};

// rank counting over NG register groups of 32 hits: rank of a hit = number of hits with a smaller index
template <int NG>
__device__ __forceinline__ int cg_select_groups(CgHits &hb, int H, int keep, int lane) {
    int mi[NG], rank[NG];
    float md[NG];
#pragma unroll
    for (int u = 0; u < NG; ++u) {
        const int e = u * 32 + lane;
        rank[u] = 0;
        mi[u] = e < H ? hb.idx[e] : 0x7fffffff;
        md[u] = e < H ? hb.d2[e] : 0.f;
    }
    if (lane < 4) hb.idx[H + lane] = 0x7fffffff;  // pad to a multiple of four: never smaller than a real index
    __syncwarp();
    for (int j = 0; j < H; j += 4) {
        const int4 v = *reinterpret_cast<const int4 *>(&hb.idx[j]);  // broadcast read (H only grows in steps the
                                                                       // buffer keeps 16-byte aligned: idx[] is)
#pragma unroll
        for (int u = 0; u < NG; ++u)
            rank[u] += (v.x < mi[u]) + (v.y < mi[u]) + (v.z < mi[u]) + (v.w < mi[u]);
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < NG; ++u)
        if (u * 32 + lane < H && rank[u] < keep) {
            hb.idx[rank[u]] = mi[u];
            hb.d2[rank[u]] = md[u];
        }
    __syncwarp();
    return H < keep ? H : keep;
}

// keep the `keep` lowest indices of the H hits, placed in ascending index order at [0, min(H, keep))
__device__ __forceinline__ int cg_select_lowest(CgHits &hb, int H, int keep, int lane) {
    if (H <= 32) return cg_select_groups<1>(hb, H, keep, lane);
    if (H <= 64) return cg_select_groups<2>(hb, H, keep, lane);
    if (H <= 128) return cg_select_groups<4>(hb, H, keep, lane);
    return cg_select_groups<CG_CAP / 32>(hb, H, keep, lane);
}

template <int QPW>
__global__ void __launch_bounds__(CG_WARPS * 32)
cg_ball_query_kernel(const float *__restrict__ x, const float *__restrict__ y, const int64_t *__restrict__ batch_x,
                     const int64_t *__restrict__ batch_y, int64_t Nx, int64_t Ny, float r2, float inv_cell, int K,
                     int strict, int64_t shadow, CellGrid g, int64_t *__restrict__ idx, float *__restrict__ dist2) {
    constexpr int QPC = CG_WARPS * QPW;
    __shared__ RangeCache rc;
    __shared__ __align__(16) CgHits hits[CG_WARPS];
    __shared__ int pre_s[CG_WARPS][33], st_s[CG_WARPS][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * QPC;
    const int64_t q1 = (q0 + QPC < Ny) ? q0 + QPC : Ny;
    const float r2x = strict ? r2 : __uint_as_float(__float_as_uint(r2) + 1u);  // d <= r2  <=>  d < nextafter(r2)
    cta_ranges<false>(q0, q1, batch_x, batch_y, Nx, rc, warp, lane);
    CgHits &hb = hits[warp];
    int *pre = pre_s[warp], *st = st_s[warp];

    for (int u = 0; u < QPW; ++u) {
        const int64_t q = q0 + warp * QPW + u;
        if (q >= q1) break;  // warp-uniform
        const float qx = __ldg(y + 3 * q), qy = __ldg(y + 3 * q + 1), qz = __ldg(y + 3 * q + 2);
        int64_t x0, x1;
        query_range<false>(q, batch_x, batch_y, Nx, 0, 1, rc, lane, x0, x1);
        int64_t *__restrict__ row = idx + q * K;
        float *__restrict__ drow = dist2 ? dist2 + q * K : nullptr;
        int H = 0;
        const int cx = cg_coord(qx, inv_cell), cy = cg_coord(qy, inv_cell), cz = cg_coord(qz, inv_cell);
        // beyond 2^16 pitches from the origin fp32 no longer resolves 1 % of a pitch (the margin the lattice is built
        // with): such queries are swept directly
        const bool far = max(abs(cx), max(abs(cy), abs(cz))) > 65536;
        if (x1 - x0 <= CG_SMALL_CLOUD || far) {
            // small cloud: plain sweep in index order straight from global memory, hits land in their final slot
            for (int64_t i0 = x0; i0 < x1 && H < K; i0 += 32) {
                const int64_t i = i0 + lane;
                const bool ok = i < x1;
                const int64_t ii = ok ? i : x0;
                const float d = sqdist3(__ldg(x + 3 * ii), __ldg(x + 3 * ii + 1), __ldg(x + 3 * ii + 2), qx, qy, qz);
                const bool hit = ok && d < r2x;
                const unsigned m = __ballot_sync(0xffffffffu, hit);
                const int pos = H + __popc(m & lanemask_lt());
                if (hit && pos < K) {
                    row[pos] = i;
                    if (drow) drow[pos] = d;
                }
                H += __popc(m);
            }
            if (H > K) H = K;
        } else {
            const int64_t cloud = batch_y ? batch_y[q] : 0;
            int s_ = 0, n_ = 0;
            if (lane < 27) {
                const int dx = lane % 3 - 1, dy = (lane / 3) % 3 - 1, dz = lane / 9 - 1;
                cg_lookup(g, cg_key(cloud, cx + dx, cy + dy, cz + dz), s_, n_);
            }
            int total;
            const int ex = warp_excl_scan(n_, lane, total);
            __syncwarp();
            pre[lane] = ex;
            st[lane] = s_;
            if (lane == 0) pre[32] = total;
            __syncwarp();
            const int ilo = (int)x0, ihi = (int)x1;
            // 4 x 32 candidates per trip: the four row look-ups and loads are independent and in flight together
            for (int base = 0; base < total; base += 128) {
                float4 c[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int f = base + 32 * v + lane;
                    c[v] = __ldg(g.sorted + cg_flat_row(pre, st, f < total ? f : 0));
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    if (base + 32 * v >= total) break;  // warp-uniform
                    const bool ok = base + 32 * v + lane < total;
                    const int ci = __float_as_int(c[v].w);
                    const float d = sqdist3(c[v].x, c[v].y, c[v].z, qx, qy, qz);
                    const bool hit = ok && d < r2x && ci >= ilo && ci < ihi;
                    const unsigned m = __ballot_sync(0xffffffffu, hit);
                    if (m) {
                        if (hit) {
                            const int slot = H + __popc(m & lanemask_lt());
                            hb.idx[slot] = ci;
                            hb.d2[slot] = d;
                        }
                        H += __popc(m);
                        if (H > CG_CAP - 32) {
                            __syncwarp();
                            H = cg_select_lowest(hb, H, K, lane);
                        }
                    }
                }
            }
            __syncwarp();
            H = cg_select_lowest(hb, H, K, lane);  // ascending index order, first K
            for (int e = lane; e < H; e += 32) {
                row[e] = hb.idx[e];
                if (drow) drow[e] = hb.d2[e];
            }
            __syncwarp();
        }
        for (int e = H + lane; e < K; e += 32) {  // padding: shadow index, dist2 = -1
            row[e] = shadow;
            if (drow) drow[e] = -1.f;
        }
    }
}

// ------------------------------------------------------------------------- knn -----------
// one warp step: every lane offers candidate (d, ci) if `ok`; admitted ones are merged into the lane-resident list
// (d_l, i_l) sorted by (distance, index); (t, ti) = entry at rank k-1.  Same rule as knn_lane_step of neighbors.cu.
__device__ __forceinline__ void cg_knn_offer(float d, int ci, bool ok, int k, int lane, float &d_l, int &i_l, float &t,
                                             int &ti) {
    unsigned m = __ballot_sync(0xffffffffu, ok && d <= t);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float nd = __shfl_sync(0xffffffffu, d, src);
        const int ni = __shfl_sync(0xffffffffu, ci, src);
        if (nd < t || (nd == t && ni < ti)) {
            const unsigned gt = __ballot_sync(0xffffffffu, d_l > nd || (d_l == nd && i_l > ni));
            const float up_d = __shfl_up_sync(0xffffffffu, d_l, 1);
            const int up_i = __shfl_up_sync(0xffffffffu, i_l, 1);
            const int first = __ffs(gt) - 1;
            if (lane == first) { d_l = nd; i_l = ni; }
            else if (lane > first) { d_l = up_d; i_l = up_i; }
            t = __shfl_sync(0xffffffffu, d_l, k - 1);
            ti = __shfl_sync(0xffffffffu, i_l, k - 1);
            m &= __ballot_sync(0xffffffffu, d <= t);
        }
    }
}

template <int QPW>
__global__ void __launch_bounds__(CG_WARPS * 32)
cg_knn_kernel(const float *__restrict__ x, const float *__restrict__ y, const int64_t *__restrict__ batch_x,
              const int64_t *__restrict__ batch_y, int64_t Nx, int64_t Ny, int k, CellGrid g,
              int64_t *__restrict__ col, float *__restrict__ dout) {
    constexpr int QPC = CG_WARPS * QPW;
    __shared__ RangeCache rc;
    __shared__ int pre_s[CG_WARPS][33], st_s[CG_WARPS][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t q0 = (int64_t)blockIdx.x * QPC;
    const int64_t q1 = (q0 + QPC < Ny) ? q0 + QPC : Ny;
    cta_ranges<false>(q0, q1, batch_x, batch_y, Nx, rc, warp, lane);
    int *pre = pre_s[warp], *st = st_s[warp];
    const float cell = cg_knn_cell(g.bbox, batch_x, Nx, k);  // the pitch cg_insert_kernel derived
    const float inv_cell = 1.0f / cell;

    for (int u = 0; u < QPW; ++u) {
        const int64_t q = q0 + warp * QPW + u;
        if (q >= q1) break;
        const float qx = __ldg(y + 3 * q), qy = __ldg(y + 3 * q + 1), qz = __ldg(y + 3 * q + 2);
        int64_t x0, x1;
        query_range<false>(q, batch_x, batch_y, Nx, 0, 1, rc, lane, x0, x1);
        float t = CUDART_INF_F, d_l = CUDART_INF_F;
        int ti = 0x7fffffff, i_l = 0x7fffffff;
        bool done = false;
        const int cx = cg_coord(qx, inv_cell), cy = cg_coord(qy, inv_cell), cz = cg_coord(qz, inv_cell);
        // (beyond 2^15 pitches from the origin fp32 no longer resolves the 1 % slack of the pruning tests: sweep)
        if (x1 - x0 > CG_SMALL_CLOUD && max(abs(cx), max(abs(cy), abs(cz))) <= 32768) {
            const int64_t cloud = batch_y ? batch_y[q] : 0;
            const int ilo = (int)x0, ihi = (int)x1;
            const float slack = 0.01f * cell;
            for (int R = 0; R <= CG_KNN_RMAX && !done; ++R) {
                // R = 0: the query's own cell alone -- it usually fills the list, and its k-th distance then prunes
                // most of the 26 cells around it (dense regions: thousands of points per cell)
                const int side = 2 * R + 1, ncell = side * side * side;
                for (int cb = 0; cb < ncell; cb += 32) {
                    const int cid = cb + lane;
                    int s_ = 0, n_ = 0;
                    if (cid < ncell) {
                        const int dx = cid % side - R, dy = (cid / side) % side - R, dz = cid / (side * side) - R;
                        const int cheb = max(abs(dx), max(abs(dy), abs(dz)));
                        if (cheb == R) {  // shell R only
                            // squared distance from the query to the cell's box, shrunk by 1 % of the pitch for the
                            // rounding of the cell assignment: a cell farther than the k-th distance holds nothing
                            const float lx = (float)(cx + dx) * cell, ly = (float)(cy + dy) * cell, lz = (float)(cz + dz) * cell;
                            const float ex = fmaxf(fmaxf(lx - qx, qx - (lx + cell)) - slack, 0.f);
                            const float ey = fmaxf(fmaxf(ly - qy, qy - (ly + cell)) - slack, 0.f);
                            const float ez = fmaxf(fmaxf(lz - qz, qz - (lz + cell)) - slack, 0.f);
                            if (ex * ex + ey * ey + ez * ez <= t) cg_lookup(g, cg_key(cloud, cx + dx, cy + dy, cz + dz), s_, n_);
                        }
                    }
                    int total;
                    const int ex_ = warp_excl_scan(n_, lane, total);
                    if (total == 0) continue;  // warp-uniform
                    __syncwarp();
                    pre[lane] = ex_;
                    st[lane] = s_;
                    if (lane == 0) pre[32] = total;
                    __syncwarp();
                    for (int base = 0; base < total; base += 128) {
                        float4 c[4];
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const int f = base + 32 * v + lane;
                            c[v] = __ldg(g.sorted + cg_flat_row(pre, st, f < total ? f : 0));
                        }
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            if (base + 32 * v >= total) break;  // warp-uniform
                            const int ci = __float_as_int(c[v].w);
                            const float d = sqdist3(c[v].x, c[v].y, c[v].z, qx, qy, qz);
                            cg_knn_offer(d, ci, base + 32 * v + lane < total && ci >= ilo && ci < ihi, k, lane, d_l, i_l, t, ti);
                        }
                    }
                    __syncwarp();
                }
                // every point not yet met is farther than (R - 0.01) cells from the query: it lies outside the cube of
                // 2R+1 cells around the query's cell (the 1 % absorbs the rounding of the cell assignment), or in a
                // cell that was pruned against a k-th distance that has only shrunk since
                const float reach = ((float)R - 0.01f) * cell;
                if (R >= 1 && ti != 0x7fffffff && t < reach * reach) done = true;
            }
        }
        if (!done) {
            // plain sweep of the whole cloud in index order (small clouds, sparse neighbourhoods, fewer than k points)
            t = CUDART_INF_F; d_l = CUDART_INF_F;
            ti = 0x7fffffff; i_l = 0x7fffffff;
            for (int64_t i0 = x0; i0 < x1; i0 += 128) {  // 4 x 32 candidates per trip, their loads in flight together
                float cxv[4], cyv[4], czv[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int64_t i = i0 + 32 * v + lane;
                    const int64_t ii = i < x1 ? i : x0;
                    cxv[v] = __ldg(x + 3 * ii);
                    cyv[v] = __ldg(x + 3 * ii + 1);
                    czv[v] = __ldg(x + 3 * ii + 2);
                }
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int64_t i = i0 + 32 * v + lane;
                    if (i0 + 32 * v >= x1) break;  // warp-uniform
                    const float d = sqdist3(cxv[v], cyv[v], czv[v], qx, qy, qz);
                    cg_knn_offer(d, (int)(i < x1 ? i : x0), i < x1, k, lane, d_l, i_l, t, ti);
                }
            }
        }
        if (lane < k) {
            const bool have = i_l != 0x7fffffff;
            col[q * k + lane] = have ? (int64_t)i_l : -1;
            if (dout) dout[q * k + lane] = have ? d_l : -1.f;
        }
    }
}

static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static uint32_t table_size(int64_t Nx) {
    uint32_t t = 1024;
    while ((int64_t)t < 2 * Nx) t <<= 1;
    return t;
}

static size_t cellgrid_bytes(int64_t Nx) {
    const size_t T = table_size(Nx);
    return align256(T * sizeof(CgSlot)) + 256 + 2 * align256((size_t)Nx * 4) + align256((size_t)Nx * 16);
}

static CellGrid carve(void *ws, int64_t Nx) {
    CellGrid g;
    const size_t T = table_size(Nx);
    char *p = (char *)ws;
    g.slots = (CgSlot *)p; p += align256(T * sizeof(CgSlot));
    g.cursor = (int *)p;            // cursor + bbox share one 256-byte block right behind the table
    g.bbox = (unsigned *)(p + 16);
    p += 256;
    g.pslot = (int *)p; p += align256((size_t)Nx * 4);
    g.prank = (int *)p; p += align256((size_t)Nx * 4);
    g.sorted = (float4 *)p;
    g.mask = (uint32_t)(T - 1);
    return g;
}

// builds the cell list of x; cell_size <= 0 -> derived on the device for k nearest neighbours
static int build_grid(const float *x, const int64_t *batch_x, int64_t Nx, float cell_size, int k, const CellGrid &g,
                      cudaStream_t st) {
    const size_t T = (size_t)g.mask + 1;
    cudaMemsetAsync(g.slots, 0, align256(T * sizeof(CgSlot)) + 16, st);   // empty table + cursor
    int64_t grid = (Nx + 255) / 256;
    if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
    if (cell_size <= 0.f) {
        cudaMemsetAsync(g.bbox, 0xff, 12, st);
        cudaMemsetAsync(g.bbox + 3, 0, 12, st);
        cg_bbox_kernel<<<(unsigned)(grid > kNumSMs ? kNumSMs : grid), 256, 0, st>>>(x, Nx, g.bbox);
        if (int e = finish_launch("cellgrid_bbox")) return e;
    }
    cg_insert_kernel<<<(unsigned)grid, 256, 0, st>>>(x, batch_x, Nx, cell_size, k, g);
    if (int e = finish_launch("cellgrid_insert")) return e;
    cg_alloc_kernel<<<(unsigned)((T + 255) / 256), 256, 0, st>>>(g);
    if (int e = finish_launch("cellgrid_alloc")) return e;
    cg_scatter_kernel<<<(unsigned)grid, 256, 0, st>>>(x, Nx, g);
    return finish_launch("cellgrid_scatter");
}

}  // namespace p3d

using namespace p3d;

P3D_EXPORT size_t p3d_cellgrid_ws_bytes(int64_t Nx) { return Nx > 0 ? cellgrid_bytes(Nx) : 0; }

P3D_EXPORT int p3d_ball_query_batch_grid_f32(const float *x, const float *y, const int64_t *batch_x,
                                             const int64_t *batch_y, int64_t Nx, int64_t Ny, float radius, int nsample,
                                             int strict, int64_t shadow, int64_t *idx, float *dist2, void *ws,
                                             size_t ws_bytes, void *stream) {
    P3D_REQUIRE(Nx >= 0 && Ny >= 0 && nsample >= 0, "ball_query_batch_grid: negative dimension");
    P3D_REQUIRE(Nx < ((int64_t)1 << 30), "ball_query_batch_grid: more than 2^30 candidate points");
    P3D_REQUIRE(nsample <= CG_MAX_K, "ball_query_batch_grid: nsample must be <= %d (use p3d_ball_query_batch_f32)", CG_MAX_K);
    P3D_REQUIRE(radius > 0.f, "ball_query_batch_grid: radius must be positive");
    if (Ny == 0 || nsample == 0) return P3D_OK;
    P3D_REQUIRE(y && idx && (x || Nx == 0), "ball_query_batch_grid: null pointer");
    P3D_REQUIRE((batch_x == nullptr) == (batch_y == nullptr), "ball_query_batch_grid: give both batch vectors or none");
    P3D_REQUIRE(Nx == 0 || (ws && ws_bytes >= cellgrid_bytes(Nx) && (((uintptr_t)ws) & 255) == 0),
                "ball_query_batch_grid: 256-byte aligned workspace of %zu bytes required", cellgrid_bytes(Nx));
    cudaStream_t st = (cudaStream_t)stream;
    CellGrid g{};
    // cell pitch = radius + 1 %: two points within `radius` of each other always fall into adjacent cells,
    // whatever the rounding of floor(p / pitch) (up to 2^16 pitches from the origin; farther queries are swept)
    const float cell = radius * 1.01f;
    if (Nx > 0) {
        g = carve(ws, Nx);
        if (int e = build_grid(x, batch_x, Nx, cell, 0, g, st)) return e;
    }
    const float r2 = radius * radius;
    const int64_t grid = (Ny + CG_WARPS * 4 - 1) / (CG_WARPS * 4);
    cg_ball_query_kernel<4><<<(unsigned)grid, CG_WARPS * 32, 0, st>>>(x, y, batch_x, batch_y, Nx, Ny, r2, 1.0f / cell,
                                                                     nsample, strict, shadow, g, idx, dist2);
    return finish_launch("cellgrid_ball_query");
}

P3D_EXPORT int p3d_knn_batch_grid_f32(const float *x, const float *y, const int64_t *batch_x, const int64_t *batch_y,
                                      int64_t Nx, int64_t Ny, int k, int64_t *col, float *d2, void *ws,
                                      size_t ws_bytes, void *stream) {
    P3D_REQUIRE(Nx >= 0 && Ny >= 0, "knn_batch_grid: negative dimension");
    P3D_REQUIRE(Nx < ((int64_t)1 << 30), "knn_batch_grid: more than 2^30 candidate points");
    P3D_REQUIRE(k >= 1 && k <= 32, "knn_batch_grid: k must be in [1, 32], got %d (use p3d_knn_batch_f32)", k);
    if (Ny == 0) return P3D_OK;
    P3D_REQUIRE(y && col && (x || Nx == 0), "knn_batch_grid: null pointer");
    P3D_REQUIRE((batch_x == nullptr) == (batch_y == nullptr), "knn_batch_grid: give both batch vectors or none");
    P3D_REQUIRE(Nx > 0, "knn_batch_grid: no support points (use p3d_knn_batch_f32)");
    P3D_REQUIRE(ws && ws_bytes >= cellgrid_bytes(Nx) && (((uintptr_t)ws) & 255) == 0,
                "knn_batch_grid: 256-byte aligned workspace of %zu bytes required", cellgrid_bytes(Nx));
    cudaStream_t st = (cudaStream_t)stream;
    CellGrid g = carve(ws, Nx);
    if (int e = build_grid(x, batch_x, Nx, 0.f, k, g, st)) return e;
    const int64_t grid = (Ny + CG_WARPS * 4 - 1) / (CG_WARPS * 4);
    cg_knn_kernel<4><<<(unsigned)grid, CG_WARPS * 32, 0, st>>>(x, y, batch_x, batch_y, Nx, Ny, k, g, col, d2);
    return finish_launch("cellgrid_knn");
}
