// grid.cu -- GridSampling3D on the device.
//
// Reference behaviour (core/data_transform/grid_transform.py:110-126):
//     coords  = torch.round(pos / size)
//     cluster = voxel_grid(coords, batch, 1)            # torch_cluster.grid_cluster over [x, y, z, batch]
//     cluster, unique_pos_indices = consecutive_cluster(cluster)
//     group_data(...)                                   # scatter_mean / item[unique_pos_indices] / label vote
//
// One persistent cooperative kernel (one CTA per SM, grid-wide barriers between phases) does the whole
// position side: voxel coordinates + extent reduction, voxel keys (x fastest, batch slowest: the
// grid_cluster stride order), a stable LSD radix sort of (key, point index) over the significant key bits
// only, run heads -> consecutive cluster ids, the last point of every voxel (`unique_pos_indices`) and the
// per-voxel mean position summed in point-index order (fp32, what a sequential scatter_add produces).
// Feature means and label votes are two small companion kernels over the (order, seg_start) segment list.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace p3d {

constexpr int kGridThreads = 512;
constexpr int kGridWarps = kGridThreads / 32;
constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;

// ---- workspace header ------------------------------------------------------------------------------
struct GridHeader {
    unsigned mn[4];  // ordered-uint encoding of the per-dimension minimum of [cx, cy, cz, batch]
    unsigned mx[4];
    unsigned pad[8];
};

__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

struct GridArgs {
    const float *pos;
    const int64_t *batch;  // nullable
    int64_t N;
    float size;
    int64_t *cluster;     // [N]
    int64_t *last_index;  // [N] (first n rows valid)
    int32_t *order;       // [N]
    int32_t *seg_start;   // [N+1]
    float *mean_pos;      // [N,3] nullable
    int64_t *out_batch;   // [N]   nullable
    int32_t *coords_out;  // [N,3] nullable
    int64_t *n_out;       // device scalar
    GridHeader *hdr;
    float4 *coords;          // [N]
    unsigned long long *k0;  // [N]
    unsigned long long *k1;  // [N]
    unsigned *v0;            // [N]
    unsigned *v1;            // [N]
    unsigned *hist;          // [kRadix * G]
    unsigned *block_heads;   // [G]
};

// voxel strides of torch_cluster.grid_cluster: num_voxels = ((end - start) / size).long() + 1, cumulative
// product, x fastest.  All arithmetic on the rounded coordinates is fp32 like the reference's.
struct GridGeom {
    float start[4];
    long long stride[4];
    int nbits;
};

__device__ __forceinline__ GridGeom grid_geometry(const GridHeader *h, bool has_batch) {
    GridGeom g;
    long long run = 1;
    const int nd = has_batch ? 4 : 3;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        g.start[d] = 0.f;
        g.stride[d] = 0;
        if (d < nd) {
            const float lo = ord2f(h->mn[d]), hi = ord2f(h->mx[d]);
            g.start[d] = lo;
            g.stride[d] = run;
            run *= (long long)__fsub_rn(hi, lo) + 1;
        }
    }
    unsigned long long top = (unsigned long long)(run > 0 ? run - 1 : 0);
    g.nbits = top ? 64 - __clzll((long long)top) : 0;
    if (run <= 0) g.nbits = 64;  // overflowed products: sort on everything
    return g;
}

__global__ void __launch_bounds__(kGridThreads, 1) grid_cluster_kernel(GridArgs a) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned s_cnt[kGridWarps][kRadix];  // per-warp digit counts of a tile (scatter) / scratch
    __shared__ unsigned s_base[kRadix];             // running global offset of every digit for this CTA
    __shared__ unsigned s_tile[kRadix];             // offset of the current tile per digit
    __shared__ unsigned s_red[kGridWarps * 8];
    __shared__ unsigned s_scan[kGridWarps];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = gridDim.x, b = blockIdx.x;
    const int64_t N = a.N;
    const int64_t chunk = ((N + G - 1) / G + kGridThreads - 1) / kGridThreads * kGridThreads;
    const int64_t lo = min(N, (int64_t)b * chunk), hi = min(N, lo + chunk);
    const bool has_batch = a.batch != nullptr;

    // ---- phase A: rounded voxel coordinates and their extent -------------------------------------
    {
        float mn[4] = {INFINITY, INFINITY, INFINITY, INFINITY};
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        const float size = a.size;
        for (int64_t i = lo + tid; i < hi; i += kGridThreads) {
            float4 c;
            c.x = rintf(__fdiv_rn(a.pos[3 * i + 0], size));
            c.y = rintf(__fdiv_rn(a.pos[3 * i + 1], size));
            c.z = rintf(__fdiv_rn(a.pos[3 * i + 2], size));
            c.w = has_batch ? (float)a.batch[i] : 0.f;
            a.coords[i] = c;
            mn[0] = fminf(mn[0], c.x), mx[0] = fmaxf(mx[0], c.x);
            mn[1] = fminf(mn[1], c.y), mx[1] = fmaxf(mx[1], c.y);
            mn[2] = fminf(mn[2], c.z), mx[2] = fmaxf(mx[2], c.z);
            mn[3] = fminf(mn[3], c.w), mx[3] = fmaxf(mx[3], c.w);
        }
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            unsigned omn = f2ord(mn[d]), omx = f2ord(mx[d]);
            omn = __reduce_min_sync(0xffffffffu, omn);
            omx = __reduce_max_sync(0xffffffffu, omx);
            if (lane == 0) s_red[warp * 8 + d] = omn, s_red[warp * 8 + 4 + d] = omx;
        }
        __syncthreads();
        if (tid < 8 && lo < hi) {
            unsigned v = s_red[tid];
            for (int w = 1; w < kGridWarps; ++w) {
                const unsigned o = s_red[w * 8 + tid];
                v = (tid < 4) ? min(v, o) : max(v, o);
            }
            if (tid < 4) atomicMin(&a.hdr->mn[tid], v);
            else atomicMax(&a.hdr->mx[tid - 4], v);
        }
    }
    grid.sync();

    // ---- phase B: voxel keys ---------------------------------------------------------------------
    const GridGeom geo = grid_geometry(a.hdr, has_batch);
    for (int64_t i = lo + tid; i < hi; i += kGridThreads) {
        const float4 c = a.coords[i];
        long long key = (long long)__fsub_rn(c.x, geo.start[0]) * geo.stride[0];
        key += (long long)__fsub_rn(c.y, geo.start[1]) * geo.stride[1];
        key += (long long)__fsub_rn(c.z, geo.start[2]) * geo.stride[2];
        if (has_batch) key += (long long)__fsub_rn(c.w, geo.start[3]) * geo.stride[3];
        a.k0[i] = (unsigned long long)key;
        a.v0[i] = (unsigned)i;
    }
    const int passes = (geo.nbits + kRadixBits - 1) / kRadixBits;

    // ---- phase C: stable LSD radix sort over the significant digits ------------------------------
    unsigned long long *kin = a.k0, *kout = a.k1;
    unsigned *vin = a.v0, *vout = a.v1;
    for (int p = 0; p < passes; ++p) {
        const int shift = p * kRadixBits;
        // C1: digit histogram of this CTA's range
        if (tid < kRadix) s_base[tid] = 0;
        __syncthreads();
        // pass 0 counts keys this CTA wrote itself in phase B; later passes read what every CTA scattered
        // before the grid barrier that closed the previous pass
        for (int64_t i = lo + tid; i < hi; i += kGridThreads) {
            const unsigned d = (unsigned)(kin[i] >> shift) & (kRadix - 1);
            atomicAdd(&s_base[d], 1u);
        }
        __syncthreads();
        if (tid < kRadix) a.hist[tid * G + b] = s_base[tid];
        grid.sync();
        // C2: every CTA derives its own digit offsets from the G x 256 table (x fastest = CTA order)
        {
            unsigned total = 0, before = 0;
            if (tid < kRadix) {
                const unsigned *row = a.hist + tid * G;
                for (int j = 0; j < G; ++j) {
                    const unsigned v = row[j];
                    before += (j < b) ? v : 0u;
                    total += v;
                }
            }
            // exclusive scan of `total` over the 256 digits (threads 0..255 = warps 0..7)
            unsigned incl = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            unsigned wbase = 0;
            for (int w = 0; w < warp; ++w) wbase += s_scan[w];
            if (tid < kRadix) s_base[tid] = wbase + incl - total + before;
            __syncthreads();
        }
        // C3: stable scatter, tile by tile in index order
        for (int64_t t0 = lo; t0 < hi; t0 += kGridThreads) {
            for (int j = tid; j < kGridWarps * kRadix; j += kGridThreads) (&s_cnt[0][0])[j] = 0;
            __syncthreads();
            const int64_t i = t0 + tid;
            const bool valid = i < hi;
            unsigned long long key = 0;
            unsigned val = 0, d = kRadix + lane;  // invalid lanes never match a digit or each other
            if (valid) {
                key = kin[i];
                val = vin[i];
                d = (unsigned)(key >> shift) & (kRadix - 1);
            }
            const unsigned peers = __match_any_sync(0xffffffffu, d);
            const unsigned r = __popc(peers & lanemask_lt());
            if (valid && r == 0) s_cnt[warp][d] = __popc(peers);
            __syncthreads();
            if (tid < kRadix) {
                unsigned run = 0;
#pragma unroll
                for (int w = 0; w < kGridWarps; ++w) {
                    const unsigned c = s_cnt[w][tid];
                    s_cnt[w][tid] = run;
                    run += c;
                }
                s_tile[tid] = s_base[tid];
                s_base[tid] += run;
            }
            __syncthreads();
            if (valid) {
                const unsigned dst = s_tile[d] + s_cnt[warp][d] + r;
                kout[dst] = key;
                vout[dst] = val;
            }
            __syncthreads();
        }
        grid.sync();
        unsigned long long *tk = kin;
        kin = kout, kout = tk;
        unsigned *tv = vin;
        vin = vout, vout = tv;
    }
    if (passes == 0) grid.sync();  // keys of the neighbouring CTA (kin[lo-1], kin[hi]) must be visible below

    // ---- phase D: run heads -> consecutive cluster ids --------------------------------------------
    // D1: heads in this CTA's range
    {
        unsigned cnt = 0;
        for (int64_t i = lo + tid; i < hi; i += kGridThreads) cnt += (i == 0 || kin[i] != kin[i - 1]) ? 1u : 0u;
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) s_scan[warp] = cnt;
        __syncthreads();
        if (tid == 0) {
            unsigned t = 0;
            for (int w = 0; w < kGridWarps; ++w) t += s_scan[w];
            a.block_heads[b] = t;
        }
    }
    grid.sync();
    // D2: ranks, cluster ids, segment starts, last point of every voxel
    {
        unsigned before = 0, total = 0;
        for (int j = lane; j < G; j += 32) {
            const unsigned v = a.block_heads[j];
            before += (j < b) ? v : 0u;
            total += v;
        }
        before = __reduce_add_sync(0xffffffffu, before);
        total = __reduce_add_sync(0xffffffffu, total);
        unsigned run = before;  // heads before the current tile
        for (int64_t t0 = lo; t0 < hi; t0 += kGridThreads) {
            const int64_t i = t0 + tid;
            const bool valid = i < hi;
            unsigned long long key = 0;
            bool head = false, tail = false;
            if (valid) {
                key = kin[i];
                head = (i == 0) || (kin[i - 1] != key);
                tail = (i == N - 1) || (kin[i + 1] != key);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, head);
            __syncthreads();  // s_scan reuse
            if (lane == 0) s_scan[warp] = __popc(bal);
            __syncthreads();
            unsigned wbase = 0, tile_total = 0;
#pragma unroll
            for (int w = 0; w < kGridWarps; ++w) {
                const unsigned c = s_scan[w];
                wbase += (w < warp) ? c : 0u;
                tile_total += c;
            }
            if (valid) {
                const unsigned incl = run + wbase + __popc(bal & (lanemask_lt() | (1u << lane)));
                const unsigned cid = incl - 1;
                const unsigned src = vin[i];
                a.cluster[src] = (int64_t)cid;
                a.order[i] = (int32_t)src;
                if (head) a.seg_start[cid] = (int32_t)i;
                if (tail) a.last_index[cid] = (int64_t)src;  // stable sort: highest point index of the voxel
            }
            run += tile_total;
        }
        if (b == 0 && tid == 0) {
            *a.n_out = (int64_t)total;
            a.seg_start[total] = (int32_t)N;
        }
    }
    if (a.mean_pos == nullptr && a.out_batch == nullptr && a.coords_out == nullptr) return;
    grid.sync();

    // ---- phase E: per-voxel outputs ------------------------------------------------------------
    {
        const int64_t n = *a.n_out;
        for (int64_t c = (int64_t)b * kGridThreads + tid; c < n; c += (int64_t)G * kGridThreads) {
            const int s = a.seg_start[c], e = a.seg_start[c + 1];
            const int64_t last = a.last_index[c];
            if (a.mean_pos) {
                float sx = 0.f, sy = 0.f, sz = 0.f;
                for (int j = s; j < e; ++j) {  // ascending point index: the order of a sequential scatter_add
                    const int64_t src = a.order[j];
                    sx = __fadd_rn(sx, a.pos[3 * src + 0]);
                    sy = __fadd_rn(sy, a.pos[3 * src + 1]);
                    sz = __fadd_rn(sz, a.pos[3 * src + 2]);
                }
                const float cntf = (float)(e - s);
                a.mean_pos[3 * c + 0] = __fdiv_rn(sx, cntf);
                a.mean_pos[3 * c + 1] = __fdiv_rn(sy, cntf);
                a.mean_pos[3 * c + 2] = __fdiv_rn(sz, cntf);
            }
            if (a.out_batch) a.out_batch[c] = has_batch ? a.batch[last] : 0;
            if (a.coords_out) {
                const float4 q = a.coords[last];
                a.coords_out[3 * c + 0] = (int32_t)q.x;
                a.coords_out[3 * c + 1] = (int32_t)q.y;
                a.coords_out[3 * c + 2] = (int32_t)q.z;
            }
        }
    }
}

// out[c, ch] = (sum over the voxel's points, ascending point index, of x[p, ch]) / count
__global__ void segment_mean_kernel(const float *__restrict__ x, const int32_t *__restrict__ order,
                                    const int32_t *__restrict__ seg_start, const int64_t *__restrict__ n_dev,
                                    int64_t n_rows, int C, float *__restrict__ out) {
    const int64_t n = n_dev ? min(n_rows, *n_dev) : n_rows;
    const int64_t total = n * C;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / C;
        const int ch = (int)(t - c * C);
        const int s = seg_start[c], e = seg_start[c + 1];
        float acc = 0.f;
        for (int j = s; j < e; ++j) acc = __fadd_rn(acc, x[(int64_t)order[j] * C + ch]);
        out[t] = __fdiv_rn(acc, (float)(e - s));
    }
}

__global__ void segment_vote_hist_kernel(const int64_t *__restrict__ labels, const int64_t *__restrict__ cluster,
                                         int64_t N, int64_t lmin, int L, int *__restrict__ hist) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = labels[i] - lmin;
        if (l >= 0 && l < L) atomicAdd(&hist[cluster[i] * L + l], 1);
    }
}

__global__ void segment_vote_argmax_kernel(const int *__restrict__ hist, int64_t n, int64_t lmin, int L,
                                           int64_t *__restrict__ out) {
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x) {
        int best = 0, bc = hist[c * L];
        for (int l = 1; l < L; ++l) {
            const int v = hist[c * L + l];
            if (v > bc) bc = v, best = l;  // first maximum, like argmax over the summed one-hot rows
        }
        out[c] = lmin + best;
    }
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static int grid_blocks(int64_t N) {
    int64_t g = (N + 2 * kGridThreads - 1) / (2 * kGridThreads);
    if (g < 1) g = 1;
    return (int)(g > kNumSMs ? kNumSMs : g);
}

}  // namespace p3d

using namespace p3d;

P3D_EXPORT size_t p3d_grid_cluster_ws_bytes(int64_t N) {
    if (N <= 0) return 256;
    const size_t n = (size_t)N;
    return align256(sizeof(GridHeader)) + align256(16 * n) + 2 * align256(8 * n) + 2 * align256(4 * n) +
           align256(sizeof(unsigned) * kRadix * kNumSMs) + align256(sizeof(unsigned) * kNumSMs);
}

P3D_EXPORT int p3d_grid_cluster_f32(const float *pos, const int64_t *batch, int64_t N, float size, int64_t *cluster,
                                    int64_t *last_index, int32_t *order, int32_t *seg_start, float *mean_pos,
                                    int64_t *out_batch, int32_t *coords_out, int64_t *n_out, void *ws,
                                    size_t ws_bytes, void *stream) {
    P3D_REQUIRE(N >= 0 && N < (int64_t)1 << 31, "grid_cluster: N=%lld out of range", (long long)N);
    P3D_REQUIRE(size > 0.f, "grid_cluster: voxel size must be positive (got %g)", (double)size);
    P3D_REQUIRE(n_out && seg_start, "grid_cluster: n_out and seg_start are required");
    cudaStream_t st = (cudaStream_t)stream;
    if (N == 0) {
        cudaMemsetAsync(n_out, 0, sizeof(int64_t), st);
        cudaMemsetAsync(seg_start, 0, sizeof(int32_t), st);
        return P3D_OK;
    }
    P3D_REQUIRE(pos && cluster && last_index && order, "grid_cluster: null pointer");
    if (!ws || ws_bytes < p3d_grid_cluster_ws_bytes(N)) {
        set_error("grid_cluster: workspace of %zu bytes required", p3d_grid_cluster_ws_bytes(N));
        return P3D_ERR_WORKSPACE;
    }
    const size_t n = (size_t)N;
    char *p = (char *)ws;
    GridArgs a;
    a.pos = pos, a.batch = batch, a.N = N, a.size = size;
    a.cluster = cluster, a.last_index = last_index, a.order = order, a.seg_start = seg_start;
    a.mean_pos = mean_pos, a.out_batch = out_batch, a.coords_out = coords_out, a.n_out = n_out;
    a.hdr = (GridHeader *)p, p += align256(sizeof(GridHeader));
    a.coords = (float4 *)p, p += align256(16 * n);
    a.k0 = (unsigned long long *)p, p += align256(8 * n);
    a.k1 = (unsigned long long *)p, p += align256(8 * n);
    a.v0 = (unsigned *)p, p += align256(4 * n);
    a.v1 = (unsigned *)p, p += align256(4 * n);
    a.hist = (unsigned *)p, p += align256(sizeof(unsigned) * kRadix * kNumSMs);
    a.block_heads = (unsigned *)p;
    cudaMemsetAsync(a.hdr->mn, 0xff, sizeof(unsigned) * 4, st);
    cudaMemsetAsync(a.hdr->mx, 0x00, sizeof(unsigned) * 4, st);
    void *args[] = {&a};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)grid_cluster_kernel, dim3(grid_blocks(N)),
                                                dim3(kGridThreads), args, 0, st);
    if (e != cudaSuccess) {
        set_error("grid_cluster: cooperative launch failed: %s", cudaGetErrorString(e));
        return P3D_ERR_CUDA;
    }
    return finish_launch("grid_cluster_kernel");
}

P3D_EXPORT int p3d_segment_mean_f32(const float *x, const int32_t *order, const int32_t *seg_start,
                                    const int64_t *n_dev, int64_t n_rows, int C, float *out, void *stream) {
    P3D_REQUIRE(n_rows >= 0 && C >= 0, "segment_mean: bad shape");
    if (n_rows == 0 || C == 0) return P3D_OK;
    P3D_REQUIRE(x && order && seg_start && out, "segment_mean: null pointer");
    const int64_t total = n_rows * C;
    int64_t blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    segment_mean_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, order, seg_start, n_dev, n_rows, C, out);
    return finish_launch("segment_mean_kernel");
}

P3D_EXPORT int p3d_segment_vote_i64(const int64_t *labels, const int64_t *cluster, int64_t N, int64_t n,
                                    int64_t label_min, int num_labels, int32_t *hist, int64_t *out, void *stream) {
    P3D_REQUIRE(N >= 0 && n >= 0 && num_labels >= 1, "segment_vote: bad shape");
    if (N == 0 || n == 0) return P3D_OK;
    P3D_REQUIRE(labels && cluster && hist && out, "segment_vote: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (N + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    segment_vote_hist_kernel<<<(unsigned)blocks, 256, 0, st>>>(labels, cluster, N, label_min, num_labels, hist);
    int rc = finish_launch("segment_vote_hist_kernel");
    if (rc) return rc;
    blocks = (n + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    segment_vote_argmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(hist, n, label_min, num_labels, out);
    return finish_launch("segment_vote_argmax_kernel");
}
