// kpconv_tc.cu -- rigid KPConv forward with the kernel-point contraction on tcgen05 tensor cores.
//
// Replaces KPConv_ops (torch_points3d/modules/KPConv/convolution_ops.py:19-107).
//
//   out[p, :] = sum_k wf[p, k, :] @ W[k]          wf[p, k, c] = sum_m w[p, k, m] * x[nb[p, m], c]
//
// A persistent CTA walks tiles of 16 or 32 query points (one or two per producer warp).  The contraction over (kernel
// point, input channel) is cut into PASSES of one 32-channel slab = 16 kernel points x 32 channels = 512 K-elements.
// Three roles, coupled by mbarriers only (no CTA-wide barrier after start-up):
//
//   producers (16 warps, SIMT) -- per point (its influence table is built once and serves every slab): the neighbour
//       rows of the slab go to shared memory with 16-byte cp.async (in flight while the influence weights are computed
//       with lanes = neighbours; one warp ballot per kernel point marks the ~6 % of (kernel point, neighbour) pairs
//       with a nonzero linear influence); then lanes = (kernel point, half slab): every lane walks the influencing
//       neighbours of ITS kernel point and accumulates 16 channels from shared memory (4 LDS.128 + 16 FMA per pair).
//       The fp32 result is split into bf16 hi + lo and stored with two 16-byte stores per operand into the A tile
//       (K-major, 128-byte swizzle: the canonical UMMA layout).  A warp needs the tile only at the very end of its
//       point (the store): it waits for the previous pass's MMAs there and nowhere earlier; with 32-point tiles the
//       first point's row waits in registers until the second point is done.  Shared memory goes to the gather scratch
//       of 16 producer warps and a 2..8-slot weight ring; layers whose ring allows it get a second A stage (fwd_split).
//   MMA warp -- W pieces (pre-split bf16 hi/lo rows, contiguous, pre-swizzled by kpconv_prep_w_kernel) stream through a
//       ring of TMA bulk copies that runs ahead of the MMAs.  Per 16-K step TWO instructions: A_hi x [B_hi | B_lo]
//       (N = 2 ng: the two partial products land in separate accumulator columns) and A_lo x B_hi (bf16x2 split:
//       ~2^-17 relative, fp32 accumulation in TMEM; rows beyond the tile of the M = 128 MMA are don't-care rows).
//       The whole warp runs the issue loop so that its counters and descriptors are uniform registers; one elected
//       lane executes the tcgen05 / TMA / commit instructions.  tcgen05.commit hands A stages and W slots back.
//       Accumulators are double-buffered in TMEM.
//   epilogue warp -- tcgen05.ld of both halves of the finished accumulator, their sum, 128-byte row stores, while
//       the next tile is already being accumulated in the other TMEM buffer.
//
// No [n,M,K], [n,M,Cin], [n,K,Cin] or [K,n,Cout] intermediate exists in HBM.
#include <cuda_bf16.h>

#include "async.cuh"
#include "kpconv_common.cuh"
#include "tc.cuh"

namespace p3d {

// (TC_TILE / TC_PWARPS / TC_THREADS describe the generic GEMM kernel below; the KPConv forward has its own F2_* set)
constexpr int TC_TILE = 64;           // rows per tile of gemm_tc_kernel (rows 0-63 of the M = 128 MMA)
constexpr int TC_PWARPS = 16;         // producer warps of gemm_tc_kernel
constexpr int TC_THREADS = (TC_PWARPS + 1) * 32;
constexpr int TC_MCAP = 40;           // neighbours per point supported by this path
constexpr int TC_NCH = 32;            // output channels per MMA (UMMA N)
constexpr int TC_A_CHUNK = TC_TILE * 128;            // one 64-K chunk of the A tile: 64 rows x 128 B = 8 KB
constexpr int TC_A_BYTES = 2 * 8 * TC_A_CHUNK;       // hi|lo x 8 chunks = 128 KB
constexpr int TC_WCH = 2;                            // 64-K chunks per weight piece
constexpr int TC_W_PIECE = 2 * TC_WCH * TC_NCH * 128;  // hi|lo x 2 chunks x 32 rows x 128 B = 16 KB (a quarter slab of K)

// ---- W image: bf16 hi/lo, K-major, 128-byte swizzled tiles ---------------------------------
// The output channels are cut into GROUPS of F2_NG = min(Cout, 128) columns (one UMMA N).  piece(s, g, kc) =
// [hi | lo] x [N_g rows = outputs of group g][128 B] for the 64 K-elements kc of slab s, always F2_PIECE(NG) bytes
// apart (a narrower last group leaves the tail of its piece unused).  K index inside a slab = k * 32 + c (k = kernel
// point, c = channel inside the 32-channel slab s; kernel points >= K are zero rows), so chunk kc holds kernel points
// 2 kc and 2 kc + 1.
__host__ __device__ constexpr int f2_piece_bytes(int ng) { return 2 * ng * 128; }

__global__ void __launch_bounds__(256)
kpconv_prep_w_kernel(const float *__restrict__ W, int K, int Cin, int Cout, int NG, uint8_t *__restrict__ img) {
    extern __shared__ __align__(16) uint8_t tile[];  // the piece in its final layout (f2_piece_bytes(NG))
    const int groups = (Cout + NG - 1) / NG;
    const int piece = blockIdx.x;  // (s * groups + g) * 8 + kc
    const int kc = piece & 7, g = (piece >> 3) % groups, s = (piece >> 3) / groups;
    const int ng = (Cout - g * NG < NG) ? (Cout - g * NG) : NG;
    // lanes run over output channels (contiguous in W): every warp load is one 128-byte line
    const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
    for (int r0 = 0; r0 < ng; r0 += 32) {
        const int r = r0 + lane;
#pragma unroll 4
        for (int ee = 0; ee < 8; ++ee) {
            const int e = ee * 8 + wq;               // K element inside the chunk: 0..63
            const int kidx = kc * 64 + e;
            const int k = kidx >> 5, c = kidx & 31;
            const float v = (k < K && r < ng) ? __ldg(W + ((int64_t)k * Cin + s * 32 + c) * Cout + g * NG + r) : 0.f;
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
            if (r < ng) {
                const int off = r * 128 + (((e >> 3) ^ (r & 7)) << 4) + ((e & 7) << 1);
                *reinterpret_cast<__nv_bfloat16 *>(tile + off) = hi;
                *reinterpret_cast<__nv_bfloat16 *>(tile + ng * 128 + off) = lo;
            }
        }
    }
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(img + (size_t)piece * f2_piece_bytes(NG));
    const uint4 *src = reinterpret_cast<const uint4 *>(tile);
    for (int i = threadIdx.x; i < f2_piece_bytes(ng) / 16; i += 256) dst[i] = src[i];
}

// ---- producer: wf[16 kernel points][32 channels] of one query point -----------------------
struct TcPoint {
    int j0, j1;           // feature row of neighbour lane / 32+lane (-1 = shadow or absent)
    float x0, y0, z0;     // centred neighbour lane
    float x1, y1, z1;     // centred neighbour 32+lane
};

__device__ __forceinline__ void tc_load_point(const float *__restrict__ q, const float *__restrict__ s,
                                              const int64_t *__restrict__ nbp, int64_t N0, int M, int lane,
                                              TcPoint &P) {
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    int64_t a = lane < M ? __ldg(nbp + lane) : -1;
    int64_t b = 32 + lane < M ? __ldg(nbp + 32 + lane) : -1;
    if (a >= N0) a = -1;
    if (b >= N0) b = -1;
    P.j0 = (int)a;
    P.j1 = (int)b;
    float ax = 1e6f, ay = 1e6f, az = 1e6f, bx = 1e6f, by = 1e6f, bz = 1e6f;
    if (a >= 0) { ax = __ldg(s + 3 * a); ay = __ldg(s + 3 * a + 1); az = __ldg(s + 3 * a + 2); }
    if (b >= 0) { bx = __ldg(s + 3 * b); by = __ldg(s + 3 * b + 1); bz = __ldg(s + 3 * b + 2); }
    P.x0 = __fsub_rn(ax, qx); P.y0 = __fsub_rn(ay, qy); P.z0 = __fsub_rn(az, qz);
    P.x1 = __fsub_rn(bx, qx); P.y1 = __fsub_rn(by, qy); P.z1 = __fsub_rn(bz, qz);
}

// influence of one kernel point, throughput version for the tensor path: sqrt.approx + multiply by
// 1/extent instead of the correctly rounded sqrt and division of the fp32 path (difference <= 3 ulp of
// the weight, far inside the bf16x2 operand error)
template <int INFL>
__device__ __forceinline__ float tc_influence(float cx, float cy, float cz, const float4 kp, const KpGeom &g,
                                              float inv_extent, float &d2) {
    const float dx = cx - kp.x, dy = cy - kp.y, dz = cz - kp.z;
    d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    if (INFL == P3D_INFLUENCE_LINEAR) {
        float r;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d2));
        return fmaxf(fmaf(-r, inv_extent, 1.0f), 0.f);
    }
    if (INFL == P3D_INFLUENCE_CONSTANT) return 1.0f;
    return __expf(-d2 / g.gauss_den);
}

// influence weights of all 16 kernel points on every neighbour, kept in REGISTERS: w0[k] = weight of kernel
// point k on neighbour `lane`; w1[it] = weight of kernel point (lane/8 + 4*it) on neighbour 32 + lane%8; plus
// one bit mask of influencing neighbours per kernel point (mask0: neighbours 0-31, mask1: 4 x 8 bits, kernel
// point k in byte k%4 of mask1[k/4], neighbours 32-39)
template <int INFL, bool CLOSEST>
__device__ __forceinline__ void tc_weights(const TcPoint &P, const float4 *kps, int K, int M, const KpGeom &g,
                                           float inv_extent, float (&w0)[KP_K], float (&w1)[4],
                                           unsigned (&mask0)[KP_K], unsigned (&mask1)[4], int lane) {
    constexpr bool closest = CLOSEST;
    // closest mode: the winning kernel point of every neighbour over ALL K kernel points (first minimum)
    int best0 = -1, best1 = -1;
    if (closest) {
        float bd0 = CUDART_INF_F, bd1 = CUDART_INF_F;
        for (int k = 0; k < K; ++k) {
            const float4 kp = kps[k];
            float d0, d1;
            tc_influence<INFL>(P.x0, P.y0, P.z0, kp, g, inv_extent, d0);
            tc_influence<INFL>(P.x1, P.y1, P.z1, kp, g, inv_extent, d1);
            if (d0 < bd0) { bd0 = d0; best0 = k; }
            if (d1 < bd1) { bd1 = d1; best1 = k; }
        }
    }
#pragma unroll
    for (int k = 0; k < KP_K; ++k) {
        float d2;
        float w = tc_influence<INFL>(P.x0, P.y0, P.z0, kps[k], g, inv_extent, d2);
        if (closest && best0 != k) w = 0.f;
        const bool nz = P.j0 >= 0 && k < K && w > 0.f;
        mask0[k] = __ballot_sync(0xffffffffu, nz);
        w0[k] = w;
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) { mask1[it] = 0u; w1[it] = 0.f; }
    if (M > 32) {
        // neighbours 32..39: lane = (mm = lane % 8, kq = lane / 8), four steps cover the 16 kernel points
        const int mm = lane & 7, kq = lane >> 3;
        const float cx = __shfl_sync(0xffffffffu, P.x1, mm), cy = __shfl_sync(0xffffffffu, P.y1, mm);
        const float cz = __shfl_sync(0xffffffffu, P.z1, mm);
        const int jj = __shfl_sync(0xffffffffu, P.j1, mm);
        const int bb = __shfl_sync(0xffffffffu, best1, mm);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int k = kq + 4 * it;
            float d2;
            float w = tc_influence<INFL>(cx, cy, cz, kps[k], g, inv_extent, d2);
            if (closest && bb != k) w = 0.f;
            const bool nz = jj >= 0 && k < K && w > 0.f;
            mask1[it] = __ballot_sync(0xffffffffu, nz);  // byte q = neighbours 32..39 of kernel point q + 4*it
            w1[it] = w;
        }
    }
}

// ---- forward kernel ---------------------------------------------------------------------------------------------
constexpr int F2_PWARPS = 16;                            // producer warps (warps 1..16)
constexpr int F2_THREADS = (F2_PWARPS + 2) * 32;         // warp 0 = epilogue, warps 1..16 = producers, warp 17 = MMA
// A tile: [hi|lo][8 chunks of 64 K][stages x 16 * ppw rows][128 B] = 32 KB x stages x ppw (with two stages the
// producers fill one while the tensor core reads the other); the weight ring takes the rest of a fixed 96 KB block.
// fwd_split() picks (stages, ppw) in {(1,1), (2,1), (1,2)} per layer shape
constexpr int F2_AW_BYTES = 98304;
constexpr int F2_MAXSLOTS = 8;                           // W ring slots = ring bytes / piece bytes, at most 8
constexpr int F2_XPITCH = 36;                            // floats per staged neighbour row (32 + 4: bank spread)
constexpr int F2_WT_PITCH = 40;                          // floats per kernel point in the weight table
constexpr int F2_SCRATCH_FLOATS = TC_MCAP * F2_XPITCH + KP_K * F2_WT_PITCH;   // per producer warp: 8320 B

// influence weights of one point with lanes = neighbours: the nonzero pattern of every kernel point as a warp ballot
// (mask0: neighbours 0-31; mask1[it]: byte q = neighbours 32-39 of kernel point q + 4 * it) and the weights themselves
// in the warp's shared-memory table wt[k][m].  Returns the masks of the kernel point THIS lane accumulates (lane >> 1).
template <int INFL, bool CLOSEST>
__device__ __forceinline__ void f2_weights(const TcPoint &P, const float4 *kps, int K, int M, const KpGeom &g,
                                           float inv_extent, float *wt, int lane, unsigned &my0, unsigned &my1) {
    int best0 = -1, best1 = -1;
    if (CLOSEST) {  // the winning kernel point of every neighbour over ALL K kernel points (first minimum)
        float bd0 = CUDART_INF_F, bd1 = CUDART_INF_F;
        for (int k = 0; k < K; ++k) {
            const float4 kp = kps[k];
            float d0, d1;
            tc_influence<INFL>(P.x0, P.y0, P.z0, kp, g, inv_extent, d0);
            tc_influence<INFL>(P.x1, P.y1, P.z1, kp, g, inv_extent, d1);
            if (d0 < bd0) { bd0 = d0; best0 = k; }
            if (d1 < bd1) { bd1 = d1; best1 = k; }
        }
    }
    const int myk = lane >> 1;
    my0 = 0u;
    my1 = 0u;
#pragma unroll
    for (int k = 0; k < KP_K; ++k) {
        if (k == KP_K - 1 && K < KP_K) break;  // the usual 15 kernel points: slot 15 is padding
        float d2;
        float w = tc_influence<INFL>(P.x0, P.y0, P.z0, kps[k], g, inv_extent, d2);
        if (CLOSEST && best0 != k) w = 0.f;
        const bool nz = P.j0 >= 0 && k < K && w > 0.f;
        const unsigned b = __ballot_sync(0xffffffffu, nz);
        wt[k * F2_WT_PITCH + lane] = w;
        if (myk == k) my0 = b;
    }
    if (M > 32) {
        // neighbours 32..39: lane = (mm = lane % 8, kq = lane / 8), four steps cover the 16 kernel points
        const int mm = lane & 7, kq = lane >> 3;
        const float cx = __shfl_sync(0xffffffffu, P.x1, mm), cy = __shfl_sync(0xffffffffu, P.y1, mm);
        const float cz = __shfl_sync(0xffffffffu, P.z1, mm);
        const int jj = __shfl_sync(0xffffffffu, P.j1, mm);
        const int bb = __shfl_sync(0xffffffffu, best1, mm);
#pragma unroll
        for (int it = 0; it < 4; ++it) {
            const int k = kq + 4 * it;
            float d2;
            float w = tc_influence<INFL>(cx, cy, cz, kps[k], g, inv_extent, d2);
            if (CLOSEST && bb != k) w = 0.f;
            const bool nz = jj >= 0 && k < K && w > 0.f;
            const unsigned b = __ballot_sync(0xffffffffu, nz);  // byte q = neighbours 32..39 of kernel point q + 4*it
            wt[k * F2_WT_PITCH + 32 + mm] = w;
            if ((myk >> 2) == it) my1 = (b >> (8 * (myk & 3))) & 0xffu;
        }
    }
}

// geometry of one query point held across the lanes of a producer warp, loaded in two dependent steps that the
// producer loop keeps two points ahead (indices) and one point ahead (coordinates) of the point it is working on
struct TcIdx {
    int64_t a, b;  // neighbour index of lane / 32 + lane (-1 = none)
};
__device__ __forceinline__ void f2_load_idx(const int64_t *__restrict__ nbp, int M, int lane, TcIdx &I) {
    I.a = lane < M ? __ldg(nbp + lane) : -1;
    I.b = 32 + lane < M ? __ldg(nbp + 32 + lane) : -1;
}
// raw loads only: the result (P = neighbour coordinates, Q = query point) is centred by f2_centre when the point's
// turn comes, one tile later, so no instruction waits on these loads here
struct TcQuery {
    float x, y, z;
};
__device__ __forceinline__ void f2_load_coords(const float *__restrict__ q, const float *__restrict__ s, int64_t N0,
                                               const TcIdx &I, TcPoint &P, TcQuery &Q) {
    Q.x = __ldg(q); Q.y = __ldg(q + 1); Q.z = __ldg(q + 2);
    const int64_t a = I.a >= N0 ? -1 : I.a, b = I.b >= N0 ? -1 : I.b;
    P.j0 = (int)a;
    P.j1 = (int)b;
    P.x0 = P.y0 = P.z0 = P.x1 = P.y1 = P.z1 = 1e6f;
    if (a >= 0) { P.x0 = __ldg(s + 3 * a); P.y0 = __ldg(s + 3 * a + 1); P.z0 = __ldg(s + 3 * a + 2); }
    if (b >= 0) { P.x1 = __ldg(s + 3 * b); P.y1 = __ldg(s + 3 * b + 1); P.z1 = __ldg(s + 3 * b + 2); }
}
__device__ __forceinline__ void f2_centre(TcPoint &P, const TcQuery &Q) {
    P.x0 = __fsub_rn(P.x0, Q.x); P.y0 = __fsub_rn(P.y0, Q.y); P.z0 = __fsub_rn(P.z0, Q.z);
    P.x1 = __fsub_rn(P.x1, Q.x); P.y1 = __fsub_rn(P.y1, Q.y); P.z1 = __fsub_rn(P.z1, Q.z);
}

// gridDim = (tile CTAs, output-channel groups, slab groups).  blockIdx.z owns the slabs [sl0, sl1) of the contraction
// and writes its partial sums to out + blockIdx.z * part_stride (part_stride = 0 when the slabs are not split).
template <int INFL, bool CLOSEST, bool SAVE_WF>
__global__ void __launch_bounds__(F2_THREADS, 1)
kpconv_fwd_tc_kernel(const float *__restrict__ q, const float *__restrict__ s, const int64_t *__restrict__ nb,
                     const float *__restrict__ x, const float *__restrict__ kp, const uint8_t *__restrict__ wimg,
                     int64_t n, int64_t N0, int M, int K, int Cin, int Cout, KpGeom g, uint32_t tmem_cols, int NG, int gpc,
                     int spz, int ppw, int stages, int64_t part_stride, float *__restrict__ out, float *__restrict__ wf_out) {
    // All shared memory of this kernel is dynamic (no static variables), so the buffer starts at the base of the CTA's
    // shared window and the 1024-byte alignment the UMMA 128-byte swizzle atoms need costs no slack: the kernel uses
    // 226.5 of the 227 KB a CTA can have.
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();            // (never observed; a misaligned tile would corrupt silently)
    const int a_rows = F2_PWARPS * ppw;                      // query points per tile
    const uint32_t a_stage = (uint32_t)a_rows * 128u;        // one stage of a chunk
    const uint32_t a_chunk = (uint32_t)stages * a_stage;     // one 64-K chunk of the A tile (all stages)
    const int ssh = stages - 1;   // stages is 1 or 2: pass % stages = pass & ssh, pass / stages = pass >> ssh
    const uint32_t a_bytes = 16u * a_chunk;                  // hi|lo x 8 chunks
    uint8_t *a_tile = smem;                                  // [hi|lo][8 chunks][stages x a_rows][128 B]
    uint8_t *w_ring = smem + a_bytes;                        // slots x [hi | lo][N rows][128 B]
    float *scratch = reinterpret_cast<float *>(smem + F2_AW_BYTES);
    uint8_t *tail = smem + F2_AW_BYTES + (size_t)F2_PWARPS * F2_SCRATCH_FLOATS * sizeof(float);
    float4 *kps = reinterpret_cast<float4 *>(tail);                          // [KP_K]
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail + KP_K * sizeof(float4));
    uint64_t *bar_a_full = bars, *bar_a_empty = bars + 2, *bar_d_full = bars + 4, *bar_d_empty = bars + 6,
             *bar_w_full = bars + 8, *bar_w_empty = bars + 8 + F2_MAXSLOTS;
    uint32_t &tmem_base_s = *reinterpret_cast<uint32_t *>(bars + 8 + 2 * F2_MAXSLOTS);
    const uint32_t piece_bytes = (uint32_t)f2_piece_bytes(NG);
    const uint32_t wslots = min((uint32_t)F2_MAXSLOTS, (F2_AW_BYTES - a_bytes) / piece_bytes);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int groups = (Cout + NG - 1) / NG;
    // blockIdx.y owns the output-channel groups [g0, g0 + gl); blockIdx.z the slabs [sl0, sl1): small point sets
    // (deep layers) are spread over the SMs by output channels and by slabs of the contraction, every CTA recomputing
    // the (cheap) producer side of its tile
    const int g0 = blockIdx.y * gpc;
    const int gl = (groups - g0 < gpc) ? (groups - g0) : gpc;
    const int col0 = g0 * NG;                                                      // first output channel of this CTA
    const int nc0 = g0;  // (the CTA with g0 == 0 stores wf)
    const int sl0 = blockIdx.z * spz;
    const int sl1 = (sl0 + spz < Cin / 32) ? (sl0 + spz) : Cin / 32;
    const int slabs = sl1 - sl0;
    float *__restrict__ dst_out = out + (int64_t)blockIdx.z * part_stride;
    const int64_t ntiles = (n + a_rows - 1) / a_rows;
    const int64_t my_tiles = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;  // persistent: tiles blockIdx.x + i * gridDim.x
    const float inv_extent = 1.0f / g.extent;
    // TMEM columns of one accumulator buffer: every output-channel group takes 2 * NG columns -- [A_hi B_hi + A_lo B_hi |
    // A_hi B_lo]: the hi and lo rows of a weight piece are contiguous, so ONE MMA with N = 2 * ng multiplies A_hi by
    // both (two issues per product instead of three; the epilogue adds the two halves)
    const uint32_t dcols = (uint32_t)(gl * 2 * NG);

    if (tid < KP_K) kps[tid] = tid < K ? make_float4(__ldg(kp + 3 * tid), __ldg(kp + 3 * tid + 1), __ldg(kp + 3 * tid + 2), 0.f)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bar_a_full[i], F2_PWARPS);
            mbar_init(&bar_a_empty[i], 1);
            mbar_init(&bar_d_full[i], 1);
            mbar_init(&bar_d_empty[i], 1);
        }
        for (int i = 0; i < F2_MAXSLOTS; ++i) {
            mbar_init(&bar_w_full[i], 1);
            mbar_init(&bar_w_empty[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == F2_PWARPS + 1) tmem_alloc(&tmem_base_s, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp >= 1 && warp <= F2_PWARPS) {
        // ------------------------------------------------------------ producers ----------------------------
        // a warp owns ONE row of every tile: its point's influence table is built once and serves every slab
        const int pw = warp - 1;
        float *xs = scratch + (size_t)pw * F2_SCRATCH_FLOATS;     // staged neighbour rows [40][36]
        float *wt = xs + TC_MCAP * F2_XPITCH;                      // weight table [16][40]
        for (int i = lane; i < F2_SCRATCH_FLOATS; i += 32) xs[i] = 0.f;
        __syncwarp();
        const int myk = lane >> 1, myh = lane & 1;
        const uint32_t xs_row = smem_u32(xs) + (uint32_t)((lane >> 3) * F2_XPITCH * 4 + (lane & 7) * 16);
        const float *wrow = wt + myk * F2_WT_PITCH;
        const float *xh = xs + myh * 16;
        // the points of this warp in processing order: unit u = (tile u / ppw, row pw * ppw + u % ppw)
        const int units = (int)my_tiles * ppw;
        const int psh = ppw - 1;  // ppw is 1 or 2: u / ppw = u >> psh, u % ppw = u & psh (no integer division in the loop)
        auto point_of = [&](int u) {
            return ((int64_t)blockIdx.x + (int64_t)(u >> psh) * gridDim.x) * a_rows + pw * ppw + (u & psh);
        };
        // geometry pipeline: coordinates of the current point, indices of the next one
        TcPoint P;
        TcQuery Q;
        Q.x = Q.y = Q.z = 0.f;
        TcIdx I1;
        I1.a = I1.b = -1;
        {
            TcIdx I0;
            I0.a = I0.b = -1;
            if (units > 0 && point_of(0) < n) f2_load_idx(nb + point_of(0) * M, M, lane, I0);
            if (units > 1 && point_of(1) < n) f2_load_idx(nb + point_of(1) * M, M, lane, I1);
            if (units > 0 && point_of(0) < n) f2_load_coords(q + 3 * point_of(0), s, N0, I0, P, Q);
        }
        uint32_t pass = 0;
        // 32-point tiles (ppw == 2): the first point's operand row is held in registers while the second point is
        // gathered, weighted and aggregated, and both rows are stored after ONE late wait for the tile -- a whole
        // point's work runs under the MMAs of the previous pass
        uint32_t held_hi[8], held_lo[8];
#pragma unroll
        for (int u8 = 0; u8 < 8; ++u8) held_hi[u8] = held_lo[u8] = 0u;
        for (int u = 0; u < units; ++u) {
            const int64_t p = point_of(u);
            const int pi = u & psh;
            TcPoint Pn;
            TcQuery Qn;
            Qn.x = Qn.y = Qn.z = 0.f;
            TcIdx I2;
            I2.a = I2.b = -1;
            if (u + 2 < units && point_of(u + 2) < n) f2_load_idx(nb + point_of(u + 2) * M, M, lane, I2);
            if (u + 1 < units && point_of(u + 1) < n) f2_load_coords(q + 3 * point_of(u + 1), s, N0, I1, Pn, Qn);
            if (p < n) f2_centre(P, Q);
            unsigned mask0 = 0u, mask1 = 0u;
            for (int sl = sl0; sl < sl1; ++sl) {   // (ppw == 2 is only launched with one slab per CTA)
                const int stage = (int)(pass & (uint32_t)ssh);
                const int row = stage * a_rows + pw * ppw + pi;
                float acc[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[c] = 0.f;
                if (p < n) {
                    // neighbour rows of this slab -> shared memory, 16 bytes per lane and copy: 4 rows per
                    // instruction; their latency hides behind the influence weights (first slab) / the wait below
                    const float *xcol = x + sl * 32 + (lane & 7) * 4;
#pragma unroll
                    for (int i = 0; i < TC_MCAP / 4; ++i) {
                        const int r = 4 * i + (lane >> 3);
                        const int j = __shfl_sync(0xffffffffu, r < 32 ? P.j0 : P.j1, r & 31);
                        if (r < M && j >= 0)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xs_row + (uint32_t)(i * 4 * F2_XPITCH * 4)),
                                         "l"(xcol + (int64_t)j * Cin)
                                         : "memory");
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                    if (sl == sl0) f2_weights<INFL, CLOSEST>(P, kps, K, M, g, inv_extent, wt, lane, mask0, mask1);
                }
                if (p < n) {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    __syncwarp();  // rows and weights written by other lanes are visible
                    // lane = (kernel point myk, half slab myh): walk the influencing neighbours of myk
                    unsigned m0 = mask0, m1 = mask1;
                    while (m0) {
                        const int m = __ffs(m0) - 1;
                        m0 &= m0 - 1;
                        const float w = wrow[m];
                        const float4 *xr = reinterpret_cast<const float4 *>(xh + m * F2_XPITCH);
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const float4 xv = xr[v];
                            acc[4 * v] = fmaf(w, xv.x, acc[4 * v]);
                            acc[4 * v + 1] = fmaf(w, xv.y, acc[4 * v + 1]);
                            acc[4 * v + 2] = fmaf(w, xv.z, acc[4 * v + 2]);
                            acc[4 * v + 3] = fmaf(w, xv.w, acc[4 * v + 3]);
                        }
                    }
                    while (m1) {
                        const int m = 32 + __ffs(m1) - 1;
                        m1 &= m1 - 1;
                        const float w = wrow[m];
                        const float4 *xr = reinterpret_cast<const float4 *>(xh + m * F2_XPITCH);
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const float4 xv = xr[v];
                            acc[4 * v] = fmaf(w, xv.x, acc[4 * v]);
                            acc[4 * v + 1] = fmaf(w, xv.y, acc[4 * v + 1]);
                            acc[4 * v + 2] = fmaf(w, xv.z, acc[4 * v + 2]);
                            acc[4 * v + 3] = fmaf(w, xv.w, acc[4 * v + 3]);
                        }
                    }
                    __syncwarp();  // the staged rows are rewritten for the next slab / point
                    if (SAVE_WF && nc0 == 0 && myk < K) {
                        // training: keep the fp32 weighted features (dW = wf^T dout needs them): 64 contiguous
                        // bytes per lane, the warp covers the K x 32 block of this point and slab
                        float4 *dst = reinterpret_cast<float4 *>(wf_out + (p * K + myk) * Cin + sl * 32 + myh * 16);
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            __stcs(dst + v, make_float4(acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]));
                    }
                }
                // bf16 hi/lo split; K index = myk * 32 + myh * 16 + c = lane * 16 + c: 32 contiguous bytes per lane
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float a0 = acc[2 * u], a1 = acc[2 * u + 1];
                    const float h0 = __bfloat162float(__float2bfloat16_rn(a0));
                    const float h1 = __bfloat162float(__float2bfloat16_rn(a1));
                    hi[u] = pack_bf16(h0, h1);
                    lo[u] = pack_bf16(a0 - h0, a1 - h1);
                }
                if (ppw == 2 && pi == 0) {
#pragma unroll
                    for (int u8 = 0; u8 < 8; ++u8) { held_hi[u8] = hi[u8]; held_lo[u8] = lo[u8]; }
                } else {
                    // the MMAs that read this stage `stages` passes ago must be done before a row is overwritten: waited
                    // for as late as possible -- the gather, the weights and the aggregation ran under them
                    if (pass >= (uint32_t)stages) mbar_wait(&bar_a_empty[stage], ((pass >> ssh) - 1) & 1);
                    const int g0 = (lane & 3) * 2;
                    if (ppw == 2) {
                        const int row0 = row - 1, sw0 = row0 & 7;
                        uint8_t *rp = a_tile + (lane >> 2) * a_chunk + row0 * 128;
                        *reinterpret_cast<uint4 *>(rp + ((g0 ^ sw0) << 4)) = make_uint4(held_hi[0], held_hi[1], held_hi[2], held_hi[3]);
                        *reinterpret_cast<uint4 *>(rp + (((g0 + 1) ^ sw0) << 4)) = make_uint4(held_hi[4], held_hi[5], held_hi[6], held_hi[7]);
                        rp += a_bytes / 2;
                        *reinterpret_cast<uint4 *>(rp + ((g0 ^ sw0) << 4)) = make_uint4(held_lo[0], held_lo[1], held_lo[2], held_lo[3]);
                        *reinterpret_cast<uint4 *>(rp + (((g0 + 1) ^ sw0) << 4)) = make_uint4(held_lo[4], held_lo[5], held_lo[6], held_lo[7]);
                    }
                    uint8_t *rowp = a_tile + (lane >> 2) * a_chunk + row * 128;
                    const int sw = row & 7;
                    *reinterpret_cast<uint4 *>(rowp + ((g0 ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4 *>(rowp + (((g0 + 1) ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                    rowp += a_bytes / 2;
                    *reinterpret_cast<uint4 *>(rowp + ((g0 ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    *reinterpret_cast<uint4 *>(rowp + (((g0 + 1) ^ sw) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                }
                if (pi == ppw - 1) {  // the last row of this warp for the pass
                    fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_a_full[stage]);
                    ++pass;
                }
            }
            P = Pn;
            Q = Qn;
            I1 = I2;
        }
    } else if (warp == F2_PWARPS + 1) {
        // ------------------------------------------------------------ MMA issuer ---------------------------
        // The whole warp runs the issue loop (so its counters, ring positions and descriptors are warp-uniform values
        // the compiler can keep in uniform registers); one elected lane executes the instructions with side effects.
        const bool leader = elect_one();
        {
            // every MMA re-reads its whole A operand (M = 128 rows, 16 or 32 of them real) from shared memory: the
            // count of MMAs per point is what bounds this warp, hence wide N (up to 128 output channels per
            // instruction) and, for single-slab layers, 32-point tiles
            // shared-memory matrix descriptors differ only in their 14-bit start-address field (bytes >> 4, low word):
            // one 32-bit add per operand instead of rebuilding the 64-bit descriptor for each MMA
            const uint64_t a_desc0 = tc_smem_desc(smem_u32(a_tile)), w_desc0 = tc_smem_desc(smem_u32(w_ring));
            const uint32_t desc_hi = (uint32_t)(a_desc0 >> 32);
            const uint32_t a_lo0 = (uint32_t)a_desc0, w_lo0 = (uint32_t)w_desc0;
            auto mkdesc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
            // W pieces stream through the ring in the order the MMAs use them: pass by pass (slab = sl0 + pass % slabs),
            // group by group, 64-K chunk by chunk.  All ring / stream positions are kept as running counters: this
            // thread's instruction stream is the pacing item of the CTA, an integer division per piece is not affordable
            uint32_t loads_left = (uint32_t)(my_tiles * slabs) * (uint32_t)(8 * gl);
            uint32_t ld_slot = 0, ld_round = 0, ld_sl = 0, ld_g = 0, ld_kc = 0;
            auto issue_load = [&]() {
                const uint32_t gg = (uint32_t)g0 + ld_g;
                const uint32_t ng = (uint32_t)((Cout - (int)gg * NG < NG) ? (Cout - (int)gg * NG) : NG);
                const size_t piece = ((size_t)(((uint32_t)sl0 + ld_sl) * groups + gg) * 8 + ld_kc);
                if (leader) {
                    mbar_arrive_expect_tx(&bar_w_full[ld_slot], 2 * ng * 128);
                    bulk_g2s(w_ring + ld_slot * piece_bytes, wimg + piece * piece_bytes, 2 * ng * 128, &bar_w_full[ld_slot]);
                }
                if (++ld_kc == 8) {
                    ld_kc = 0;
                    if (++ld_g == (uint32_t)gl) {
                        ld_g = 0;
                        if (++ld_sl == (uint32_t)slabs) ld_sl = 0;
                    }
                }
                if (++ld_slot == wslots) { ld_slot = 0; ++ld_round; }
                --loads_left;
            };
            for (uint32_t i = 0; i < wslots && loads_left; ++i) issue_load();
            uint32_t c_slot = 0, c_round = 0, pass = 0;
            bool first_piece = true;
            for (int64_t ti = 0; ti < my_tiles; ++ti) {
                const uint32_t dbuf = (uint32_t)(ti & 1);
                if (ti >= 2) mbar_wait(&bar_d_empty[dbuf], (uint32_t)(((ti >> 1) - 1) & 1));  // epilogue drained this buffer
                tc_fence_after();
                for (int sl = 0; sl < slabs; ++sl, ++pass) {
                    const uint32_t stage = pass & (uint32_t)ssh;
                    mbar_wait_parked(&bar_a_full[stage], (pass >> ssh) & 1, 2000);
                    tc_fence_after();
                    for (int gq = 0; gq < gl; ++gq) {
                        const int ng = (Cout - (g0 + gq) * NG < NG) ? (Cout - (g0 + gq) * NG) : NG;
                        const uint32_t idesc = tc_idesc(TC_MMA_M, ng), idesc_wide = tc_idesc(TC_MMA_M, 2 * ng);
                        const uint32_t d = tmem_base + dbuf * dcols + (uint32_t)(gq * 2 * NG);
                        for (int kc = 0; kc < 8; ++kc) {
                            mbar_wait(&bar_w_full[c_slot], c_round & 1);
                            tc_fence_after();
                            const uint32_t wl = w_lo0 + ((c_slot * piece_bytes) >> 4);
                            const uint32_t al = a_lo0 + ((kc * a_chunk + stage * a_stage) >> 4);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t ko = (uint32_t)((ks * 32) >> 4);
                                const uint64_t a_hi = mkdesc(al + ko);
                                const uint64_t a_lo = mkdesc(al + ko + ((a_bytes / 2) >> 4));
                                const uint64_t b_hi = mkdesc(wl + ko);   // hi rows, followed by the lo rows
                                const uint32_t first = (sl == 0 && kc == 0 && ks == 0) ? 0u : 1u;
                                if (leader) {
                                    tc_mma_bf16(d, a_hi, b_hi, idesc_wide, first);  // [A_hi B_hi | A_hi B_lo]
                                    tc_mma_bf16(d, a_lo, b_hi, idesc, 1u);          // + A_lo B_hi into the first half
                                }
                            }
                            if (leader) tc_commit(&bar_w_empty[c_slot]);
                            if (++c_slot == wslots) { c_slot = 0; ++c_round; }
                            // keep the ring full: the slot of the PREVIOUS piece (== the stream's next slot: loads run
                            // exactly one ring ahead) is free once its MMAs have retired -- which they do while this
                            // piece's MMAs, just issued, keep the tensor pipe busy
                            if (!first_piece && loads_left) {
                                mbar_wait(&bar_w_empty[ld_slot], (ld_round - 1) & 1);
                                issue_load();
                            }
                            first_piece = false;
                        }
                    }
                    if (leader) tc_commit(&bar_a_empty[stage]);
                }
                if (leader) tc_commit(&bar_d_full[dbuf]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------ epilogue (warp 0: TMEM lanes 0-31) ----
        for (int64_t ti = 0; ti < my_tiles; ++ti) {
            const uint32_t dbuf = (uint32_t)(ti & 1);
            mbar_wait_parked(&bar_d_full[dbuf], (uint32_t)((ti >> 1) & 1), 4000);
            tc_fence_after();
            const int64_t p = (blockIdx.x + ti * gridDim.x) * a_rows + lane;
            for (int gq = 0; gq < gl; ++gq) {
                const int ng = (Cout - (g0 + gq) * NG < NG) ? (Cout - (g0 + gq) * NG) : NG;
                const uint32_t dg = tmem_base + dbuf * dcols + (uint32_t)(gq * 2 * NG);
                for (int cq = 0; cq < ng / 32; ++cq) {
                    float v[32], v2[32];
                    tmem_ld32(dg + (uint32_t)(cq * 32), v);
                    tmem_ld32(dg + (uint32_t)(ng + cq * 32), v2);
                    if (lane < a_rows && p < n) {
                        float4 *dst = reinterpret_cast<float4 *>(dst_out + p * Cout + col0 + gq * NG + cq * 32);
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            dst[u] = make_float4(v[4 * u] + v2[4 * u], v[4 * u + 1] + v2[4 * u + 1],
                                                 v[4 * u + 2] + v2[4 * u + 2], v[4 * u + 3] + v2[4 * u + 3]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_d_empty[dbuf]);
        }
    }
    __syncthreads();
    if (warp == F2_PWARPS + 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// out[i] = ((part[0][i] + part[1][i]) + ...) in slab-group order (the split of the contraction over CTAs is
// deterministic: partial sums are stored, not accumulated atomically)
__global__ void __launch_bounds__(256)
kpconv_reduce_parts_kernel(const float4 *__restrict__ part, int splits, int64_t n4, float4 *__restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (int64_t)gridDim.x * 256) {
        float4 a = part[i];
        for (int z = 1; z < splits; ++z) {
            const float4 b = part[(int64_t)z * n4 + i];
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        out[i] = a;
    }
}

// ---- backward helpers on the same sparse producers --------------------------------------------------
// wf[p][k][c] = sum_m w[p,k,m] x[nb[p,m], c]  -> workspace (for dW = wf^T dout); one warp per point, the influence
// table is computed once per point and reused for every 32-channel slab
template <int INFL, bool CLOSEST>
__global__ void __launch_bounds__(256)
kpconv_wf_sparse_kernel(const float *__restrict__ q, const float *__restrict__ s, const int64_t *__restrict__ nb,
                        const float *__restrict__ x, const float *__restrict__ kp, int64_t n, int64_t N0, int M, int K,
                        int Cin, KpGeom g, float *__restrict__ wf) {
    __shared__ float xs_all[8 * TC_MCAP * 32];
    __shared__ float4 kps[KP_K];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < KP_K) kps[tid] = tid < K ? make_float4(__ldg(kp + 3 * tid), __ldg(kp + 3 * tid + 1), __ldg(kp + 3 * tid + 2), 0.f)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    float *xs = xs_all + warp * TC_MCAP * 32;
    for (int i = lane; i < TC_MCAP * 32; i += 32) xs[i] = 0.f;
    __syncthreads();
    const uint32_t xs_addr = smem_u32(xs) + (uint32_t)lane * 4u;
    const float inv_extent = 1.0f / g.extent;
    for (int64_t p = (int64_t)blockIdx.x * 8 + warp; p < n; p += (int64_t)gridDim.x * 8) {
        TcPoint P;
        tc_load_point(q + 3 * p, s, nb + p * M, N0, M, lane, P);
        float w0[KP_K], w1[4];
        unsigned mask0[KP_K], mask1[4];
        tc_weights<INFL, CLOSEST>(P, kps, K, M, g, inv_extent, w0, w1, mask0, mask1, lane);
        for (int sl = 0; sl < Cin / 32; ++sl) {
            const float *xcol = x + sl * 32 + lane;
#pragma unroll
            for (int m = 0; m < TC_MCAP; ++m) {
                const int j = __shfl_sync(0xffffffffu, m < 32 ? P.j0 : P.j1, m & 31);
                if (m < M && j >= 0)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(xs_addr + m * 128u),
                                 "l"(xcol + (int64_t)j * Cin)
                                 : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int k = 0; k < KP_K; ++k) {
                if (k == KP_K - 1 && K < KP_K) break;  // slot 15 is padding unless the layer has 16 kernel points
                unsigned mk = mask0[k];
                float a = 0.f;
                while (mk) {
                    const int m = __ffs(mk) - 1;
                    mk &= mk - 1;
                    a = fmaf(__shfl_sync(0xffffffffu, w0[k], m), xs[m * 32 + lane], a);
                }
                mk = (mask1[k >> 2] >> (8 * (k & 3))) & 0xffu;
                while (mk) {
                    const int mm = __ffs(mk) - 1;
                    mk &= mk - 1;
                    a = fmaf(__shfl_sync(0xffffffffu, w1[k >> 2], mm + 8 * (k & 3)), xs[(32 + mm) * 32 + lane], a);
                }
                if (k < K) wf[(p * K + k) * Cin + sl * 32 + lane] = a;
            }
        }
    }
}

// dfeat[nb[p,m]][c] += sum_k w[p,k,m] * dwf[p][k][c] over the influencing (k, m) pairs only
template <int INFL, bool CLOSEST>
__global__ void __launch_bounds__(256)
kpconv_scatter_sparse_kernel(const float *__restrict__ q, const float *__restrict__ s, const int64_t *__restrict__ nb,
                             const float *__restrict__ kp, const float *__restrict__ dwf, int64_t n, int64_t N0, int M,
                             int K, int Cin, KpGeom g, float *__restrict__ dfeat) {
    __shared__ float4 kps[KP_K];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < KP_K) kps[tid] = tid < K ? make_float4(__ldg(kp + 3 * tid), __ldg(kp + 3 * tid + 1), __ldg(kp + 3 * tid + 2), 0.f)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const float inv_extent = 1.0f / g.extent;
    for (int64_t p = (int64_t)blockIdx.x * 8 + warp; p < n; p += (int64_t)gridDim.x * 8) {
        TcPoint P;
        tc_load_point(q + 3 * p, s, nb + p * M, N0, M, lane, P);
        float w0[KP_K], w1[4];
        unsigned mask0[KP_K], mask1[4];
        tc_weights<INFL, CLOSEST>(P, kps, K, M, g, inv_extent, w0, w1, mask0, mask1, lane);
        for (int sl = 0; sl < Cin / 32; ++sl) {
            const int c = sl * 32 + lane;
            float d[KP_K];
#pragma unroll
            for (int k = 0; k < KP_K; ++k) d[k] = k < K ? __ldg(dwf + (p * K + k) * Cin + c) : 0.f;
#pragma unroll
            for (int k = 0; k < KP_K; ++k) {
                if (k == KP_K - 1 && K < KP_K) break;
                unsigned mk = mask0[k];
                while (mk) {
                    const int m = __ffs(mk) - 1;
                    mk &= mk - 1;
                    const int j = __shfl_sync(0xffffffffu, P.j0, m);
                    atomicAdd(dfeat + (int64_t)j * Cin + c, __shfl_sync(0xffffffffu, w0[k], m) * d[k]);
                }
                mk = (mask1[k >> 2] >> (8 * (k & 3))) & 0xffu;
                while (mk) {
                    const int mm = __ffs(mk) - 1;
                    mk &= mk - 1;
                    const int j = __shfl_sync(0xffffffffu, P.j1, mm);
                    atomicAdd(dfeat + (int64_t)j * Cin + c, __shfl_sync(0xffffffffu, w1[k >> 2], mm + 8 * (k & 3)) * d[k]);
                }
            }
        }
    }
}

static KpGeom tc_geom(float extent, int influence, int aggregation) {
    KpGeom g;
    g.extent = extent;
    const float sig = extent * 0.3f;
    g.gauss_den = 2.0f * sig * sig + 1e-9f;
    g.influence = influence;
    g.aggregation = aggregation;
    return g;
}

bool kpconv_sparse_supported(int M, int K, int Cin) { return M <= TC_MCAP && K <= KP_K && Cin % 32 == 0; }

#define P3D_KP_DISPATCH(KERNEL, ...)                                                          \
    do {                                                                                      \
        if (aggregation == P3D_AGGREGATION_CLOSEST) {                                         \
            if (influence == P3D_INFLUENCE_CONSTANT) KERNEL<P3D_INFLUENCE_CONSTANT, true> __VA_ARGS__; \
            else if (influence == P3D_INFLUENCE_LINEAR) KERNEL<P3D_INFLUENCE_LINEAR, true> __VA_ARGS__; \
            else KERNEL<P3D_INFLUENCE_GAUSSIAN, true> __VA_ARGS__;                            \
        } else {                                                                              \
            if (influence == P3D_INFLUENCE_CONSTANT) KERNEL<P3D_INFLUENCE_CONSTANT, false> __VA_ARGS__; \
            else if (influence == P3D_INFLUENCE_LINEAR) KERNEL<P3D_INFLUENCE_LINEAR, false> __VA_ARGS__; \
            else KERNEL<P3D_INFLUENCE_GAUSSIAN, false> __VA_ARGS__;                           \
        }                                                                                     \
    } while (0)

int kpconv_wf_sparse(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                     const float *k_points, int64_t n, int64_t N0, int M, int K, int Cin, float extent, int influence,
                     int aggregation, float *wf, cudaStream_t st) {
    const KpGeom g = tc_geom(extent, influence, aggregation);
    int64_t grid = (n + 7) / 8;
    if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
    P3D_KP_DISPATCH(kpconv_wf_sparse_kernel, <<<(unsigned)grid, 256, 0, st>>>(query, support, neighbors, feat, k_points,
                                                                             n, N0, M, K, Cin, g, wf));
    return finish_launch("kpconv_wf_sparse");
}

int kpconv_scatter_sparse(const float *query, const float *support, const int64_t *neighbors, const float *k_points,
                          const float *dwf, int64_t n, int64_t N0, int M, int K, int Cin, float extent, int influence,
                          int aggregation, float *dfeat, cudaStream_t st) {
    const KpGeom g = tc_geom(extent, influence, aggregation);
    int64_t grid = (n + 7) / 8;
    if (grid > 16 * kNumSMs) grid = 16 * kNumSMs;
    P3D_KP_DISPATCH(kpconv_scatter_sparse_kernel, <<<(unsigned)grid, 256, 0, st>>>(query, support, neighbors, k_points,
                                                                                  dwf, n, N0, M, K, Cin, g, dfeat));
    return finish_launch("kpconv_scatter_sparse");
}

// ---- generic tcgen05 GEMM  C[R x N] = A[R x Kd] * B[Kd x N]  (fp32 in HBM, bf16 hi+lo on the tensor cores) --------
// Used by the KPConv backward for dwf = dout * W^T.  Same machinery as the forward kernel: 64-row tiles in the
// canonical UMMA layout, weight pieces by TMA, TMEM accumulators, tcgen05.ld epilogue; the producers here are plain
// row loaders (16 consecutive K elements per lane, split into bf16 hi + lo).
// B image: piece(s, nc, qq) as in kpconv_prep_w_kernel but with the plain K index: kidx = k - 512 * s.
__global__ void __launch_bounds__(256)
gemm_prep_b_kernel(const float *__restrict__ B, int64_t sbk, int64_t sbj, int Kd, int N, uint8_t *__restrict__ img) {
    const int piece = blockIdx.x;  // ((s * ncs + nc) * 4 + q)
    const int ncs = N / TC_NCH;
    const int qq = piece & 3, nc = (piece >> 2) % ncs, s = (piece >> 2) / ncs;
    uint8_t *base = img + (size_t)piece * TC_W_PIECE;
    for (int i = threadIdx.x; i < TC_WCH * TC_NCH * 64; i += 256) {
        const int e = i & 63, r = (i >> 6) & 31, kc = i >> 11;
        const int k = s * 512 + (qq * TC_WCH + kc) * 64 + e;
        const float v = k < Kd ? __ldg(B + (int64_t)k * sbk + (int64_t)(nc * TC_NCH + r) * sbj) : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        const int off = (kc * TC_NCH + r) * 128 + (((e >> 3) ^ (r & 7)) << 4) + ((e & 7) << 1);
        *reinterpret_cast<__nv_bfloat16 *>(base + off) = hi;
        *reinterpret_cast<__nv_bfloat16 *>(base + TC_WCH * TC_NCH * 128 + off) = lo;
    }
}

// resident = 1: the whole B image of this CTA's output-channel chunks (ncl x chunks x 8 KB) is loaded into shared
// memory ONCE (small Kd: the backward's dwf GEMM has Kd = Cout), so a tile costs no TMA round trips at all;
// resident = 0: weight pieces are streamed per tile as in the forward kernel.
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const float *__restrict__ A, int64_t lda, const uint8_t *__restrict__ bimg, int64_t R, int Kd, int N,
               uint32_t tmem_cols, int ncpc, int resident, float *__restrict__ C, int64_t ldc) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *a_tile = smem;
    uint8_t *w_tile = smem + TC_A_BYTES;
    __shared__ __align__(8) uint64_t bar_a_empty, bar_w_full, bar_w_empty, bar_d_full;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slabs = (Kd + 511) / 512, ncs = N / TC_NCH;
    const int nc0 = blockIdx.y * ncpc;
    const int ncl = (ncs - nc0 < ncpc) ? (ncs - nc0) : ncpc;
    const int64_t ntiles = (R + TC_TILE - 1) / TC_TILE;
    if (tid == 0) {
        mbar_init(&bar_a_empty, 1);
        mbar_init(&bar_w_full, 1);
        mbar_init(&bar_w_empty, 1);
        mbar_init(&bar_d_full, 1);
        fence_barrier_init();
    }
    if (warp == TC_PWARPS) tmem_alloc(&tmem_base_s, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    uint32_t pass = 0, wloads = 0, tile_iter = 0;
    constexpr int PPW = TC_TILE / TC_PWARPS;
    const bool vec = ((lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0);
    const int kchunks = (Kd + 63) / 64;  // resident mode: Kd <= 512, one slab
    if (resident) {
        // B_res[ncq][chunk][hi 4 KB | lo 4 KB] right after the part of the A tile that is used
        if (tid == 0) {
            mbar_arrive_expect_tx(&bar_w_full, (uint32_t)(ncl * kchunks * 2 * TC_NCH * 128));
            for (int ncq = 0; ncq < ncl; ++ncq)
                for (int ch = 0; ch < kchunks; ++ch) {
                    const uint8_t *piece = bimg + ((size_t)(nc0 + ncq) * 4 + (ch >> 1)) * TC_W_PIECE;
                    uint8_t *dst = w_tile + (size_t)(ncq * kchunks + ch) * (2 * TC_NCH * 128);
                    bulk_g2s(dst, piece + (ch & 1) * (TC_NCH * 128), TC_NCH * 128, &bar_w_full);
                    bulk_g2s(dst + TC_NCH * 128, piece + TC_W_PIECE / 2 + (ch & 1) * (TC_NCH * 128), TC_NCH * 128,
                             &bar_w_full);
                }
        }
        mbar_wait(&bar_w_full, 0);
    }

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tile_iter) {
        const int64_t r_base = tile * TC_TILE;
        for (int sl = 0; sl < slabs; ++sl, ++pass) {
            const int kleft = Kd - sl * 512;                       // K elements of this slab (<= 512 used)
            const int pieces = kleft >= 512 ? 4 : (kleft + 127) / 128;  // 128-K weight pieces that hold data
            if (warp < TC_PWARPS) {
                if (pass > 0) mbar_wait(&bar_a_empty, (pass - 1) & 1);
                for (int pi = 0; pi < PPW; ++pi) {
                    const int row = warp * PPW + pi;
                    const int64_t r = r_base + row;
                    const int k0 = sl * 512 + lane * 16;
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = 0.f;
                    if (r < R && lane * 16 < pieces * 128) {
                        const float *src = A + r * lda + k0;
                        if (vec && k0 + 16 <= Kd) {
#pragma unroll
                            for (int g4 = 0; g4 < 4; ++g4) {
                                const float4 t4 = __ldg(reinterpret_cast<const float4 *>(src) + g4);
                                v[4 * g4] = t4.x; v[4 * g4 + 1] = t4.y; v[4 * g4 + 2] = t4.z; v[4 * g4 + 3] = t4.w;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e)
                                if (k0 + e < Kd) v[e] = __ldg(src + e);
                        }
                    }
                    if (lane * 16 < pieces * 128) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float h0 = __bfloat162float(__float2bfloat16_rn(v[2 * u]));
                            const float h1 = __bfloat162float(__float2bfloat16_rn(v[2 * u + 1]));
                            hi[u] = pack_bf16(h0, h1);
                            lo[u] = pack_bf16(v[2 * u] - h0, v[2 * u + 1] - h1);
                        }
                        uint8_t *rowp = a_tile + (lane >> 2) * TC_A_CHUNK + row * 128;
                        const int g0 = (lane & 3) * 2, sw = row & 7;
                        *reinterpret_cast<uint4 *>(rowp + ((g0 ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(rowp + (((g0 + 1) ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
                        rowp += TC_A_BYTES / 2;
                        *reinterpret_cast<uint4 *>(rowp + ((g0 ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        *reinterpret_cast<uint4 *>(rowp + (((g0 + 1) ^ sw) << 4)) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                    }
                }
                fence_proxy_async();
            }
            __syncthreads();
            if (warp == TC_PWARPS) {
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t idesc = tc_idesc(TC_MMA_M, TC_NCH);
                    const uint32_t a_addr = smem_u32(a_tile), w_addr = smem_u32(w_tile);
                    for (int ncq = 0; ncq < ncl; ++ncq) {
                        const int nc = nc0 + ncq;
                        const uint32_t d = tmem_base + (uint32_t)(ncq * TC_NCH);
                        if (resident) {
                            for (int ch = 0; ch < kchunks; ++ch) {
                                const uint32_t wb = w_addr + (uint32_t)((ncq * kchunks + ch) * (2 * TC_NCH * 128));
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    const uint32_t ko = (uint32_t)(ks * 32);
                                    const uint32_t ac = a_addr + (uint32_t)(ch * TC_A_CHUNK) + ko;
                                    const uint64_t a_hi = tc_smem_desc(ac);
                                    const uint64_t a_lo = tc_smem_desc(ac + TC_A_BYTES / 2);
                                    const uint64_t b_hi = tc_smem_desc(wb + ko);
                                    const uint64_t b_lo = tc_smem_desc(wb + TC_NCH * 128 + ko);
                                    const uint32_t first = (ch == 0 && ks == 0) ? 0u : 1u;
                                    tc_mma_bf16(d, a_hi, b_hi, idesc, first);
                                    tc_mma_bf16(d, a_lo, b_hi, idesc, 1u);
                                    tc_mma_bf16(d, a_hi, b_lo, idesc, 1u);
                                }
                            }
                            continue;
                        }
                        for (int qq = 0; qq < pieces; ++qq, ++wloads) {
                            if (wloads > 0) mbar_wait(&bar_w_empty, (wloads - 1) & 1);
                            mbar_arrive_expect_tx(&bar_w_full, TC_W_PIECE);
                            bulk_g2s(w_tile, bimg + ((size_t)(sl * ncs + nc) * 4 + qq) * TC_W_PIECE, TC_W_PIECE,
                                     &bar_w_full);
                            mbar_wait(&bar_w_full, wloads & 1);
                            tc_fence_after();
#pragma unroll
                            for (int kc = 0; kc < TC_WCH; ++kc) {
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    const uint32_t ko = (uint32_t)(ks * 32);
                                    const uint32_t ac = a_addr + (uint32_t)((qq * TC_WCH + kc) * TC_A_CHUNK) + ko;
                                    const uint64_t a_hi = tc_smem_desc(ac);
                                    const uint64_t a_lo = tc_smem_desc(ac + TC_A_BYTES / 2);
                                    const uint64_t b_hi = tc_smem_desc(w_addr + kc * (TC_NCH * 128) + ko);
                                    const uint64_t b_lo = tc_smem_desc(w_addr + TC_W_PIECE / 2 + kc * (TC_NCH * 128) + ko);
                                    const uint32_t first = (sl == 0 && qq == 0 && kc == 0 && ks == 0) ? 0u : 1u;
                                    tc_mma_bf16(d, a_hi, b_hi, idesc, first);
                                    tc_mma_bf16(d, a_lo, b_hi, idesc, 1u);
                                    tc_mma_bf16(d, a_hi, b_lo, idesc, 1u);
                                }
                            }
                            tc_commit(&bar_w_empty);
                        }
                    }
                    tc_commit(&bar_a_empty);
                    if (sl == slabs - 1) tc_commit(&bar_d_full);
                }
                __syncwarp();
            }
        }
        if (warp < TC_TILE / 32) {
            mbar_wait(&bar_d_full, tile_iter & 1);
            tc_fence_after();
            const int64_t r = r_base + warp * 32 + lane;
            for (int ncq = 0; ncq < ncl; ++ncq) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ncq * TC_NCH), v);
                if (r < R) {
                    float *dst = C + r * ldc + (nc0 + ncq) * TC_NCH;
                    if ((ldc & 3) == 0 && ((((uintptr_t)C) & 15) == 0)) {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            reinterpret_cast<float4 *>(dst)[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                    } else {
#pragma unroll
                        for (int u = 0; u < 32; ++u) dst[u] = v[u];
                    }
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == TC_PWARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

size_t gemm_tc_ws_bytes(int Kd, int N) {
    if (N % TC_NCH != 0 || N > 512) return 0;
    return (size_t)((Kd + 511) / 512) * (N / TC_NCH) * 4 * TC_W_PIECE;
}

// C[R x N] = A[R x Kd] * B, B(k, j) = Bp[k * sbk + j * sbj]; returns 1 when the shape is not supported
int gemm_tc(const float *A, int64_t lda, const float *Bp, int64_t sbk, int64_t sbj, int64_t R, int Kd, int N, float *C,
            int64_t ldc, void *ws, size_t ws_bytes, cudaStream_t st) {
    const size_t need = gemm_tc_ws_bytes(Kd, N);
    if (need == 0 || ws == nullptr || ws_bytes < need || (((uintptr_t)ws) & 15) != 0) return 1;
    const int ncs = N / TC_NCH;
    gemm_prep_b_kernel<<<((Kd + 511) / 512) * ncs * 4, 256, 0, st>>>(Bp, sbk, sbj, Kd, N, (uint8_t *)ws);
    if (int e = finish_launch("gemm_prep_b")) return e;
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(225 * 1024));
        attr_done.mark();
    }
    const int64_t ntiles = (R + TC_TILE - 1) / TC_TILE;
    int ygroups = 1;
    if (ntiles < kNumSMs) {
        ygroups = (int)((kNumSMs + ntiles - 1) / ntiles);
        if (ygroups > ncs) ygroups = ncs;
    }
    int ncpc = (ncs + ygroups - 1) / ygroups;
    // resident-B mode: one K slab and the B chunks of a CTA fit beside the A tile
    const int kchunks = (Kd + 63) / 64;
    int resident = 0;
    if (Kd <= 512) {
        const int fit = (int)((size_t)(92 * 1024) / ((size_t)kchunks * 2 * TC_NCH * 128));  // chunks of 8 KB in 92 KB
        if (fit >= 1) {
            resident = 1;
            if (ncpc > fit) ncpc = fit;
        }
    }
    ygroups = (ncs + ncpc - 1) / ncpc;
    uint32_t cols = 32;
    while ((int)cols < ncpc * TC_NCH) cols <<= 1;
    const size_t smem = TC_A_BYTES + (resident ? (size_t)ncpc * kchunks * 2 * TC_NCH * 128 : (size_t)TC_W_PIECE) + 1024;
    const dim3 grid((unsigned)(ntiles < kNumSMs ? ntiles : kNumSMs), (unsigned)ygroups);
    gemm_tc_kernel<<<grid, TC_THREADS, smem, st>>>(A, lda, (const uint8_t *)ws, R, Kd, N, cols, ncpc, resident, C, ldc);
    return finish_launch("gemm_tc");
}

// ---- dW[R x Cn] += A^T B over the points:  A = wf [n x R], B = dout [n x Cn]  (fp32 SIMT, split over points) -------
// CTA = 256 threads = 32 row groups x 8 column groups, thread tile 8 x 4 -> a 256 x 32 block of dW in registers;
// the CTA walks its share of the points in sub-chunks of 16 staged in shared memory and finishes with one fp32
// reduction per output element.
constexpr int DW_SUB = 32;

__global__ void __launch_bounds__(256)
dw_accumulate_kernel(const float *__restrict__ A, const float *__restrict__ B, int64_t n, int R, int Cn, int64_t p_per_cta,
                     float *__restrict__ dW, float *__restrict__ part) {
    extern __shared__ __align__(16) float dsm[];  // [2][DW_SUB][256 + 32]: double-buffered wf rows | dout rows
    constexpr int ROW = 256 + 32;
    const int tid = threadIdx.x, tr = tid >> 3, tc = tid & 7;  // row group (32), column group (8)
    const int r0 = blockIdx.y * 256, c0 = blockIdx.z * 32;
    const int64_t p_lo = (int64_t)blockIdx.x * p_per_cta;
    const int64_t p_hi = (p_lo + p_per_cta < n) ? p_lo + p_per_cta : n;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // stage DW_SUB points: thread t copies element t of every wf row (coalesced 4-byte cp.async), threads 0-31 the dout row
    auto stage_in = [&](int64_t pb, int buf) {
        float *base = dsm + (size_t)buf * DW_SUB * ROW;
        for (int pp = 0; pp < DW_SUB; ++pp) {
            const int64_t p = pb + pp;
            float *dst = base + pp * ROW;
            if (p < p_hi && r0 + tid < R)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + tid)), "l"(A + p * R + r0 + tid) : "memory");
            else
                dst[tid] = 0.f;
            if (tid < 32) {
                if (p < p_hi && c0 + tid < Cn)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + 256 + tid)), "l"(B + p * Cn + c0 + tid) : "memory");
                else
                    dst[256 + tid] = 0.f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int it = 0;
    if (p_lo < p_hi) stage_in(p_lo, 0);
    for (int64_t pb = p_lo; pb < p_hi; pb += DW_SUB, ++it) {
        if (pb + DW_SUB < p_hi) {
            stage_in(pb + DW_SUB, (it + 1) & 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const float *base = dsm + (size_t)(it & 1) * DW_SUB * ROW;
#pragma unroll 8
        for (int pp = 0; pp < DW_SUB; ++pp) {
            const float4 a0 = *reinterpret_cast<const float4 *>(base + pp * ROW + tr * 8);
            const float4 a1 = *reinterpret_cast<const float4 *>(base + pp * ROW + tr * 8 + 4);
            const float4 b = *reinterpret_cast<const float4 *>(base + pp * ROW + 256 + tc * 4);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();  // the buffer is refilled two iterations later
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + tr * 8 + i, c = c0 + tc * 4 + j;
            if (r < R && c < Cn) {
                // deterministic form: every point chunk stores its partial block, dw_reduce_kernel adds them in chunk
                // order; without a partial buffer the blocks are reduced with fp32 atomics at L2
                if (part) part[((int64_t)blockIdx.x * R + r) * Cn + c] = acc[i][j];
                else atomicAdd(dW + (int64_t)r * Cn + c, acc[i][j]);
            }
        }
}

// dW[i] += the sum of part[c][i] over the point chunks c, in a FIXED order: 8 thread groups each add a contiguous range
// of chunks in chunk order, the 8 range sums are then added in range order.  (A single thread walking all ~200 chunks
// was a chain of dependent loads: 34 us at C3 level 0 for 12 MB.)
constexpr int DWR_GROUPS = 8;
// few chunks (deep layers: few points, large dW): one thread per element, chunk order
__global__ void __launch_bounds__(256)
dw_reduce_flat_kernel(const float *__restrict__ part, int chunks, int64_t total, float *__restrict__ dW) {
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
        float a = part[i];
        for (int c = 1; c < chunks; ++c) a += part[(int64_t)c * total + i];
        dW[i] += a;
    }
}

__global__ void __launch_bounds__(256)
dw_reduce_kernel(const float *__restrict__ part, int chunks, int64_t total, float *__restrict__ dW) {
    __shared__ float sums[DWR_GROUPS][32];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int per = (chunks + DWR_GROUPS - 1) / DWR_GROUPS;
    const int c0 = g * per, c1 = (c0 + per < chunks) ? c0 + per : chunks;
    for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
        const int64_t i = base + lane;
        float a = 0.f;
        if (i < total) {
#pragma unroll 4
            for (int c = c0; c < c1; ++c) a += __ldg(part + (int64_t)c * total + i);
        }
        sums[g][lane] = a;
        __syncthreads();
        if (g == 0 && i < total) {
            float t = sums[0][lane];
#pragma unroll
            for (int q = 1; q < DWR_GROUPS; ++q) t += sums[q][lane];
            dW[i] += t;
        }
        __syncthreads();
    }
}

static void dw_plan(int64_t n, int R, int Cn, int &rb, int &cb, int64_t &chunks, int64_t &ppc) {
    rb = (R + 255) / 256;
    cb = (Cn + 31) / 32;
    chunks = (3 * kNumSMs + rb * cb - 1) / (rb * cb);  // 72 KB of staging per CTA: three CTAs per SM
    if (chunks < 1) chunks = 1;
    ppc = (n + chunks - 1) / chunks;
    ppc = (ppc + DW_SUB - 1) / DW_SUB * DW_SUB;
    if (ppc < 4 * DW_SUB) ppc = 4 * DW_SUB;
    chunks = (n + ppc - 1) / ppc;
}

// bytes of the partial buffer that makes dw_accumulate deterministic
size_t dw_partial_bytes(int64_t n, int R, int Cn) {
    int rb, cb;
    int64_t chunks, ppc;
    dw_plan(n, R, Cn, rb, cb, chunks, ppc);
    return (size_t)chunks * (size_t)R * (size_t)Cn * sizeof(float);
}

// dW += A^T B; `part` (dw_partial_bytes, nullable) selects the deterministic two-pass reduction
int dw_accumulate(const float *A, const float *B, int64_t n, int R, int Cn, float *dW, cudaStream_t st, float *part,
                  size_t part_bytes) {
    int rb, cb;
    int64_t chunks, ppc;
    dw_plan(n, R, Cn, rb, cb, chunks, ppc);
    if (rb > 65535 || cb > 65535) return 1;
    if (part != nullptr && part_bytes < (size_t)chunks * R * Cn * sizeof(float)) part = nullptr;
    static DeviceOnce attr_done;
    const size_t smem = (size_t)2 * DW_SUB * (256 + 32) * sizeof(float);
    if (!attr_done.done()) {
        cudaFuncSetAttribute(dw_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_done.mark();
    }
    dw_accumulate_kernel<<<dim3((unsigned)chunks, (unsigned)rb, (unsigned)cb), 256, smem, st>>>(A, B, n, R, Cn, ppc, dW, part);
    if (part == nullptr) return finish_launch("kpconv_dW");
    if (int e = finish_launch("kpconv_dW")) return e;
    const int64_t total = (int64_t)R * Cn;
    if (chunks <= 2 * DWR_GROUPS) {
        int64_t rg = (total + 255) / 256;
        if (rg > 8 * kNumSMs) rg = 8 * kNumSMs;
        dw_reduce_flat_kernel<<<(unsigned)rg, 256, 0, st>>>(part, (int)chunks, total, dW);
    } else {
        int64_t rg = (total + 31) / 32;
        if (rg > 8 * kNumSMs) rg = 8 * kNumSMs;
        dw_reduce_kernel<<<(unsigned)rg, 256, 0, st>>>(part, (int)chunks, total, dW);
    }
    return finish_launch("kpconv_dW_reduce");
}

// how the forward kernel spreads a problem over the SMs: output-channel groups of NG columns (grid.y takes gpc groups
// each) and slab groups (grid.z takes spz slabs each)
struct FwdSplit {
    int NG, groups, ygroups, gpc, zsplit, spz, ppw, stages;
};
constexpr int F2_NG_MAX = 128;   // output channels per accumulator group (one MMA's N)
static FwdSplit fwd_split(int64_t n, int Cin, int Cout) {
    const int slabs = Cin / 32;
    FwdSplit f;
    f.NG = Cout < F2_NG_MAX ? Cout : F2_NG_MAX;
    f.groups = (Cout + f.NG - 1) / f.NG;
    // The A tile holds stages x 16 * ppw rows and takes 32 KB x stages x ppw of the 96 KB block, the weight ring the
    // rest; a weight piece is NG / 4 KB.  Measured on the KPConvPaper encoder shapes (profiles/r2_kpconv_layers.md):
    //  - single-slab layers with narrow outputs (the big first levels): 32-point tiles, two points per producer warp
    //    -- half the MMAs per point (119 vs 145 us at level 0; 101 vs 108 us after the MMA issue stream was made
    //    cheaper: wide-N hi|lo products, warp-uniform issue loop);
    //  - other layers with NG <= 64: two 16-point stages, so the producers fill one while the tensor core reads the
    //    other (105 vs 109 us at 64 -> 64);
    //  - NG = 128: one 16-point stage and a 64 KB ring -- two 32 KB weight pieces in flight matter more than the
    //    second stage (85 vs 110 us at 128 -> 128).
    f.ppw = (slabs == 1 && f.NG <= 64) ? 2 : 1;
    f.stages = (f.ppw == 1 && f.NG <= 64) ? 2 : 1;
    const int rows = F2_PWARPS * f.ppw;
    const int64_t ntiles = (n + rows - 1) / rows;
    // two accumulator buffers of gpc * 2 * NG columns must fit the 512 TMEM columns
    int max_gpc = 256 / (2 * f.NG);   // (hi | lo halves: 2 * NG columns per group)
    if (max_gpc < 1) max_gpc = 1;
    f.ygroups = (f.groups + max_gpc - 1) / max_gpc;
    f.zsplit = 1;
    if (ntiles > 0 && ntiles * f.ygroups < kNumSMs) {
        // fewer tiles than SMs: first cut the sequential chain of slab passes, then the output channels
        int64_t want = kNumSMs / (ntiles * f.ygroups);
        f.zsplit = (int)(want < slabs ? want : slabs);
        if (f.zsplit < 1) f.zsplit = 1;
        want = kNumSMs / (ntiles * f.zsplit);
        if (want > f.groups) want = f.groups;
        if (want > f.ygroups) f.ygroups = (int)want;
    }
    f.gpc = (f.groups + f.ygroups - 1) / f.ygroups;
    f.ygroups = (f.groups + f.gpc - 1) / f.gpc;
    f.spz = (slabs + f.zsplit - 1) / f.zsplit;
    f.zsplit = (slabs + f.spz - 1) / f.spz;
    return f;
}

static size_t fwd_image_bytes(int Cin, int Cout) {
    const int NG = Cout < F2_NG_MAX ? Cout : F2_NG_MAX;
    return (size_t)(Cin / 32) * ((Cout + NG - 1) / NG) * 8 * f2_piece_bytes(NG);
}

// workspace: the bf16 hi/lo weight image, then (slab-split problems only) the partial outputs [zsplit][n][Cout]
size_t kpconv_tc_ws_bytes(int64_t n, int K, int Cin, int Cout) {
    if (K > KP_K || Cin % 32 != 0 || Cout % TC_NCH != 0 || Cout > 512) return 0;
    size_t bytes = fwd_image_bytes(Cin, Cout);
    const FwdSplit f = fwd_split(n, Cin, Cout);
    if (f.zsplit > 1) bytes += (size_t)f.zsplit * (size_t)n * Cout * sizeof(float);
    return bytes;
}

// returns 1 when the shape is not supported by this path (the caller then runs the fp32 SIMT kernel)
int kpconv_fwd_tc(const float *query, const float *support, const int64_t *neighbors, const float *feat,
                  const float *k_points, const float *k_values, int64_t n, int64_t N0, int M, int K, int Cin,
                  int Cout, float extent, int influence, int aggregation, float *out, void *ws, size_t ws_bytes,
                  cudaStream_t st, float *wf_out) {
    const size_t need = kpconv_tc_ws_bytes(n, K, Cin, Cout);
    if (need == 0 || M > TC_MCAP || ws == nullptr || ws_bytes < need) return 1;
    if ((((uintptr_t)out) & 15) != 0 || (((uintptr_t)ws) & 15) != 0) return 1;  // float4 row stores, TMA source
    if ((((uintptr_t)feat) & 15) != 0 || (wf_out && (((uintptr_t)wf_out) & 15) != 0)) return 1;  // 16-byte row copies
    KpGeom g;
    g.extent = extent;
    const float sig = extent * 0.3f;
    g.gauss_den = 2.0f * sig * sig + 1e-9f;
    g.influence = influence;
    g.aggregation = aggregation;
    const FwdSplit f = fwd_split(n, Cin, Cout);
    const int pieces_w = (Cin / 32) * f.groups * 8;
    static DeviceOnce prep_done;
    if (!prep_done.done()) {
        cudaFuncSetAttribute(kpconv_prep_w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, f2_piece_bytes(128));
        prep_done.mark();
    }
    kpconv_prep_w_kernel<<<pieces_w, 256, f2_piece_bytes(f.NG), st>>>(k_values, K, Cin, Cout, f.NG, (uint8_t *)ws);
    if (int e = finish_launch("kpconv_prep_w")) return e;
    const size_t smem = F2_AW_BYTES + (size_t)F2_PWARPS * F2_SCRATCH_FLOATS * sizeof(float) + KP_K * sizeof(float4) +
                        (8 + 2 * F2_MAXSLOTS) * sizeof(uint64_t) + 16;
    using Kern = void (*)(const float *, const float *, const int64_t *, const float *, const float *, const uint8_t *,
                          int64_t, int64_t, int, int, int, int, KpGeom, uint32_t, int, int, int, int, int, int64_t, float *,
                          float *);
    static const Kern kerns[2][3][2] = {
        {{kpconv_fwd_tc_kernel<P3D_INFLUENCE_CONSTANT, false, false>, kpconv_fwd_tc_kernel<P3D_INFLUENCE_CONSTANT, true, false>},
         {kpconv_fwd_tc_kernel<P3D_INFLUENCE_LINEAR, false, false>, kpconv_fwd_tc_kernel<P3D_INFLUENCE_LINEAR, true, false>},
         {kpconv_fwd_tc_kernel<P3D_INFLUENCE_GAUSSIAN, false, false>, kpconv_fwd_tc_kernel<P3D_INFLUENCE_GAUSSIAN, true, false>}},
        {{kpconv_fwd_tc_kernel<P3D_INFLUENCE_CONSTANT, false, true>, kpconv_fwd_tc_kernel<P3D_INFLUENCE_CONSTANT, true, true>},
         {kpconv_fwd_tc_kernel<P3D_INFLUENCE_LINEAR, false, true>, kpconv_fwd_tc_kernel<P3D_INFLUENCE_LINEAR, true, true>},
         {kpconv_fwd_tc_kernel<P3D_INFLUENCE_GAUSSIAN, false, true>, kpconv_fwd_tc_kernel<P3D_INFLUENCE_GAUSSIAN, true, true>}}};
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        for (int w = 0; w < 2; ++w)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 2; ++b)
                    cudaFuncSetAttribute(kerns[w][a][b], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // (+ static <= 227 KB)
        attr_done.mark();
    }
    const int64_t ntiles = (n + F2_PWARPS * f.ppw - 1) / (F2_PWARPS * f.ppw);
    uint32_t cols = 32;
    while ((int)cols < 2 * f.gpc * 2 * f.NG) cols <<= 1;
    float *parts = f.zsplit > 1 ? (float *)((char *)ws + fwd_image_bytes(Cin, Cout)) : out;
    const int64_t part_stride = f.zsplit > 1 ? n * (int64_t)Cout : 0;
    const dim3 grid((unsigned)(ntiles < kNumSMs ? ntiles : kNumSMs), (unsigned)f.ygroups, (unsigned)f.zsplit);
    kerns[wf_out != nullptr ? 1 : 0][influence][aggregation]<<<grid, F2_THREADS, smem, st>>>(
        query, support, neighbors, feat, k_points, (const uint8_t *)ws, n, N0, M, K, Cin, Cout, g, cols, f.NG, f.gpc, f.spz,
        f.ppw, f.stages, part_stride, parts, wf_out);
    if (f.zsplit > 1) {
        if (int e = finish_launch("kpconv_fwd_tc")) return e;
        const int64_t n4 = n * (int64_t)Cout / 4;
        int64_t rg = (n4 + 255) / 256;
        if (rg > 4 * kNumSMs) rg = 4 * kNumSMs;
        kpconv_reduce_parts_kernel<<<(unsigned)rg, 256, 0, st>>>((const float4 *)parts, f.zsplit, n4, (float4 *)out);
    }
    return finish_launch("kpconv_fwd_tc");
}

}  // namespace p3d
