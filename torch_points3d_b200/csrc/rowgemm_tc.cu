// rowgemm_tc.cu -- C[R x N] = A[R x Kd] . B on the tcgen05 tensor cores for TALL operands: R = all points of a level
// (hundreds to 50 k rows), Kd and N = channel counts (16..1536 and 16..7680).
//
// Used by the KPConv backward for dwf = dout . W^T (modules/KPConv/convolution_ops.py:95-105 differentiated: the
// gradient of the weighted features, [n, K * Cin]) and by the unary layers of the KPConv blocks
// (modules/KPConv/blocks.py:176-194: y = x . W^T and dx = dy . W).  fp32 in, fp32 out; operands are split into
// bf16 hi + lo and the three significant partial products are accumulated in fp32 in TMEM (error <= 2e-5 of max|C|,
// the contract of the KPConv forward).
//
// One CTA = one 128-row tile x one group of NG <= 128 output columns, persistent over the tiles of its group
// (grid = (tiles per group, column groups)); the reduction runs in 64-element chunks through two shared-memory rings:
//   producers (8 warps) -- a chunk of A: 128 rows x 64 floats.  Each warp owns 16 rows; all of its eight 16-byte
//       loads per lane are issued before the first conversion (32 KB in flight per SM), then hi/lo bf16 and 8-byte
//       stores into the K-major / 128-byte-swizzle operand tile [hi: 128 rows x 128 B | lo: same] of the stage.
//   MMA warp -- the matching chunk of the pre-split B image ([hi rows | lo rows] contiguous, built by
//       rowgemm_prep_b_kernel) arrives by TMA bulk copy; per 16-K step two instructions: A_hi x [B_hi | B_lo]
//       (N = 2 ng: the partial products land in separate accumulator columns) and A_lo x B_hi.  The warp runs the
//       loop uniformly, one elected lane issues; tcgen05.commit frees the A stage and the B slot.
//   epilogue (4 warps = the 128 TMEM lanes) -- adds the two accumulator halves and stores 128-byte row pieces, while
//       the next tile accumulates in the other TMEM buffer.
#include <cstdint>

#include "async.cuh"
#include "common.cuh"
#include "tc.cuh"

namespace p3d {

constexpr int RG_ROWS = 128;                 // rows per tile = M of the MMA (no idle rows)
constexpr int RG_NG_MAX = 128;               // output columns per CTA (N = 2 * 128 for the hi|lo instruction)
constexpr int RG_PWARPS = 8;                 // producer warps
constexpr int RG_EWARPS = 4;                 // epilogue warps (warp w reads TMEM lanes 32 w .. 32 w + 31)
constexpr int RG_THREADS = (RG_EWARPS + RG_PWARPS + 1) * 32;
constexpr int RG_A_STAGE = 2 * RG_ROWS * 128;  // hi | lo of one 64-K chunk: 32 KB
constexpr int RG_A_STAGES = 3;
constexpr int RG_B_STAGES = 3;

__host__ __device__ constexpr int rg_piece_bytes(int ng) { return 2 * ng * 128; }

// image piece (group g, chunk c) = [hi: ng rows x 128 B | lo: ng rows x 128 B], rows = output columns of the group,
// 64 reduction elements per row, 16-byte units XOR-swizzled by (row & 7); pieces are rg_piece_bytes(NG) apart
__global__ void __launch_bounds__(256)
rowgemm_prep_b_kernel(const float *__restrict__ Bp, int64_t sbk, int64_t sbj, int Kd, int N, int NG, int nk,
                      uint8_t *__restrict__ img) {
    extern __shared__ __align__(16) uint8_t rg_tile[];
    const int piece = blockIdx.x, c = piece % nk, g = piece / nk;
    const int ng = (N - g * NG < NG) ? (N - g * NG) : NG;
    // One work item = one 16-byte operand unit: 8 consecutive reduction elements of one row, written with ONE 16-byte
    // store per half (the XOR swizzle spreads the 8 rows of a quarter-warp over the 8 units of a 128-byte line:
    // conflict-free).  Items run over the units of a row first when the reduction index is the contiguous one
    // (sbk == 1: a warp reads 4 rows x 256 contiguous bytes), else over the rows (sbj == 1: a warp reads 32
    // consecutive columns of 8 reduction elements).
    for (int item = threadIdx.x; item < ng * 8; item += 256) {
        const int r = sbk == 1 ? item >> 3 : item % ng, u = sbk == 1 ? item & 7 : item / ng;
        const int k0 = c * 64 + u * 8;
        const float *src = Bp + (int64_t)k0 * sbk + (int64_t)(g * NG + r) * sbj;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = k0 + i < Kd ? __ldg(src + (int64_t)i * sbk) : 0.f;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float h0 = __bfloat162float(__float2bfloat16_rn(v[2 * q]));
            const float h1 = __bfloat162float(__float2bfloat16_rn(v[2 * q + 1]));
            hi[q] = pack_bf16(h0, h1);
            lo[q] = pack_bf16(v[2 * q] - h0, v[2 * q + 1] - h1);
        }
        const int off = r * 128 + ((u ^ (r & 7)) << 4);
        *reinterpret_cast<uint4 *>(rg_tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(rg_tile + ng * 128 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
    __syncthreads();
    uint4 *dst = reinterpret_cast<uint4 *>(img + (size_t)piece * rg_piece_bytes(NG));
    const uint4 *src = reinterpret_cast<const uint4 *>(rg_tile);
    for (int i = threadIdx.x; i < rg_piece_bytes(ng) / 16; i += 256) dst[i] = src[i];
}

// TRANS_A = false: A(r, k) = A[r * lda + k] (rows contiguous along the reduction).
// TRANS_A = true : A(r, k) = A[k * lda + r] (the reduction index is the slow one: C = X^T . B with X = A as stored) -- the
//     KPConv weight gradient dW[K * Cin, Cout] = wf^T . dout, a reduction over ALL points with a small output: the
//     reduction is also cut over CTAs (gridDim.z ranges of cps chunks, partial outputs part_stride apart, added in
//     range order by rowgemm_add_parts_kernel: deterministic).
template <bool TRANS_A>
__global__ void __launch_bounds__(RG_THREADS, 1)
rowgemm_tc_kernel(const float *__restrict__ A, int64_t lda, const uint8_t *__restrict__ bimg, int64_t R, int64_t Kd, int N,
                  int NG, uint32_t tmem_cols, float *__restrict__ C, int64_t ldc, int cps, int64_t part_stride) {
    extern __shared__ __align__(1024) uint8_t rg_smem[];
    if ((smem_u32(rg_smem) & 1023u) != 0u) __trap();
    const uint32_t piece_bytes = (uint32_t)rg_piece_bytes(NG);
    uint8_t *a_ring = rg_smem;                                   // RG_A_STAGES x 32 KB
    uint8_t *b_ring = rg_smem + RG_A_STAGES * RG_A_STAGE;        // RG_B_STAGES x piece_bytes
    uint64_t *bars = reinterpret_cast<uint64_t *>(b_ring + RG_B_STAGES * piece_bytes);
    uint64_t *a_full = bars, *a_empty = bars + RG_A_STAGES, *b_full = bars + 2 * RG_A_STAGES,
             *b_empty = b_full + RG_B_STAGES, *d_full = b_empty + RG_B_STAGES, *d_empty = d_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(d_empty + 2);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.y;
    const int ng = (N - g * NG < NG) ? (N - g * NG) : NG;
    const int nk_all = (int)((Kd + 63) / 64);
    const int c_lo = blockIdx.z * cps;                                   // this CTA's chunks of the reduction
    const int nk = (c_lo + cps < nk_all ? c_lo + cps : nk_all) - c_lo;   // (>= 1 by construction of the grid)
    C += (int64_t)blockIdx.z * part_stride;
    const int64_t ntiles = (R + RG_ROWS - 1) / RG_ROWS;
    const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (tid == 0) {
        for (int s = 0; s < RG_A_STAGES; ++s) { mbar_init(&a_full[s], RG_PWARPS); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < RG_B_STAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&d_full[s], 1); mbar_init(&d_empty[s], RG_EWARPS); }
        fence_barrier_init();
    }
    if (warp == RG_EWARPS + RG_PWARPS) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t dcols = (uint32_t)(2 * NG);  // one accumulator buffer: [hi.hi + lo.hi | hi.lo]

    if (warp >= RG_EWARPS && warp < RG_EWARPS + RG_PWARPS) {
        // ------------------------------------------------------------------ producers ------------------------
        const int pw = warp - RG_EWARPS;
        const bool vec = ((lda & 3) == 0) && ((((uintptr_t)A) & 15) == 0);
        const int sub = lane >> 4, q = lane & 15;  // two rows per load instruction, 16 float4 per row
        uint32_t it = 0;
        for (int64_t ti = 0; ti < my_tiles; ++ti) {
            const int64_t r_base = (blockIdx.x + ti * gridDim.x) * RG_ROWS + pw * 16;
            for (int c = 0; c < nk; ++c, ++it) {
                const uint32_t stage = it % RG_A_STAGES;
                if (TRANS_A) {
                    // warp pw = reduction elements 8 pw .. 8 pw + 7 of the chunk (one 16-byte unit of every operand
                    // row), lane = rows 4 lane .. 4 lane + 3: each load is one 512-byte line of a point's row
                    const int64_t r0 = (blockIdx.x + ti * gridDim.x) * RG_ROWS + lane * 4;
                    const int64_t kb = (int64_t)(c_lo + c) * 64 + pw * 8;
                    float4 v[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (kb + i < Kd) {
                            const float *src = A + (kb + i) * lda + r0;
                            if (vec && r0 + 4 <= R) {
                                v[i] = __ldg(reinterpret_cast<const float4 *>(src));
                            } else {
                                if (r0 < R) v[i].x = __ldg(src);
                                if (r0 + 1 < R) v[i].y = __ldg(src + 1);
                                if (r0 + 2 < R) v[i].z = __ldg(src + 2);
                                if (r0 + 3 < R) v[i].w = __ldg(src + 3);
                            }
                        }
                    }
                    if (it >= RG_A_STAGES) mbar_wait(&a_empty[stage], ((it / RG_A_STAGES) - 1) & 1);
                    uint8_t *st = a_ring + stage * RG_A_STAGE;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int row = lane * 4 + j;
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float a0 = j == 0 ? v[2 * u].x : (j == 1 ? v[2 * u].y : (j == 2 ? v[2 * u].z : v[2 * u].w));
                            const float a1 = j == 0 ? v[2 * u + 1].x
                                                    : (j == 1 ? v[2 * u + 1].y : (j == 2 ? v[2 * u + 1].z : v[2 * u + 1].w));
                            const float h0 = __bfloat162float(__float2bfloat16_rn(a0));
                            const float h1 = __bfloat162float(__float2bfloat16_rn(a1));
                            hi[u] = pack_bf16(h0, h1);
                            lo[u] = pack_bf16(a0 - h0, a1 - h1);
                        }
                        const int off = row * 128 + ((pw ^ (row & 7)) << 4);
                        *reinterpret_cast<uint4 *>(st + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4 *>(st + RG_ROWS * 128 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[stage]);
                    continue;
                }
                const int64_t k0 = (int64_t)(c_lo + c) * 64 + q * 4;
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int64_t r = r_base + 2 * i + sub;
                    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < R) {
                        const float *src = A + r * lda + k0;
                        if (vec && k0 + 4 <= Kd) {
                            v[i] = __ldg(reinterpret_cast<const float4 *>(src));
                        } else {
                            if (k0 < Kd) v[i].x = __ldg(src);
                            if (k0 + 1 < Kd) v[i].y = __ldg(src + 1);
                            if (k0 + 2 < Kd) v[i].z = __ldg(src + 2);
                            if (k0 + 3 < Kd) v[i].w = __ldg(src + 3);
                        }
                    }
                }
                if (it >= RG_A_STAGES) mbar_wait(&a_empty[stage], ((it / RG_A_STAGES) - 1) & 1);
                uint8_t *st = a_ring + stage * RG_A_STAGE;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = pw * 16 + 2 * i + sub;
                    const float h0 = __bfloat162float(__float2bfloat16_rn(v[i].x));
                    const float h1 = __bfloat162float(__float2bfloat16_rn(v[i].y));
                    const float h2 = __bfloat162float(__float2bfloat16_rn(v[i].z));
                    const float h3 = __bfloat162float(__float2bfloat16_rn(v[i].w));
                    // 64-K chunk row: 8 sixteen-byte units of 8 elements; this lane's 4 elements are half a unit
                    const int off = row * 128 + (((q >> 1) ^ (row & 7)) << 4) + ((q & 1) << 3);
                    *reinterpret_cast<uint2 *>(st + off) = make_uint2(pack_bf16(h0, h1), pack_bf16(h2, h3));
                    *reinterpret_cast<uint2 *>(st + RG_ROWS * 128 + off) =
                        make_uint2(pack_bf16(v[i].x - h0, v[i].y - h1), pack_bf16(v[i].z - h2, v[i].w - h3));
                }
                fence_proxy_async();  // generic-proxy stores -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_full[stage]);
            }
        }
    } else if (warp == RG_EWARPS + RG_PWARPS) {
        // ------------------------------------------------------------------ MMA / TMA issuer -----------------
        const bool leader = elect_one();
        const uint64_t a_desc0 = tc_smem_desc(smem_u32(a_ring)), b_desc0 = tc_smem_desc(smem_u32(b_ring));
        const uint32_t desc_hi = (uint32_t)(a_desc0 >> 32);
        const uint32_t a_lo0 = (uint32_t)a_desc0, b_lo0 = (uint32_t)b_desc0;
        auto mkdesc = [&](uint32_t lo) { return ((uint64_t)desc_hi << 32) | lo; };
        const uint8_t *gimg = bimg + ((size_t)g * nk_all + c_lo) * piece_bytes;
        uint32_t loads_left = (uint32_t)(my_tiles * nk), ld_slot = 0, ld_round = 0, ld_c = 0;
        auto issue_load = [&]() {
            if (leader) {
                mbar_arrive_expect_tx(&b_full[ld_slot], 2 * ng * 128);
                bulk_g2s(b_ring + ld_slot * piece_bytes, gimg + (size_t)ld_c * piece_bytes, 2 * ng * 128, &b_full[ld_slot]);
            }
            if (++ld_c == (uint32_t)nk) ld_c = 0;
            if (++ld_slot == RG_B_STAGES) { ld_slot = 0; ++ld_round; }
            --loads_left;
        };
        for (uint32_t i = 0; i < RG_B_STAGES && loads_left; ++i) issue_load();
        const uint32_t idesc = tc_idesc(TC_MMA_M, ng), idesc_wide = tc_idesc(TC_MMA_M, 2 * ng);
        uint32_t it = 0, b_slot = 0, b_round = 0;
        bool first_piece = true;
        for (int64_t ti = 0; ti < my_tiles; ++ti) {
            const uint32_t dbuf = (uint32_t)(ti & 1);
            if (ti >= 2) mbar_wait(&d_empty[dbuf], (uint32_t)(((ti >> 1) - 1) & 1));
            tc_fence_after();
            const uint32_t d = tmem_base + dbuf * dcols;
            for (int c = 0; c < nk; ++c, ++it) {
                const uint32_t stage = it % RG_A_STAGES;
                mbar_wait_parked(&a_full[stage], (it / RG_A_STAGES) & 1, 2000);
                mbar_wait(&b_full[b_slot], b_round & 1);
                tc_fence_after();
                const uint32_t al = a_lo0 + ((stage * RG_A_STAGE) >> 4);
                const uint32_t bl = b_lo0 + ((b_slot * piece_bytes) >> 4);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    const uint32_t ko = (uint32_t)((ks * 32) >> 4);
                    const uint64_t a_hi = mkdesc(al + ko), a_lo = mkdesc(al + ko + ((RG_ROWS * 128) >> 4));
                    const uint64_t b_hi = mkdesc(bl + ko);  // hi rows, followed by the lo rows
                    if (leader) {
                        tc_mma_bf16(d, a_hi, b_hi, idesc_wide, (c == 0 && ks == 0) ? 0u : 1u);  // [A_hi B_hi | A_hi B_lo]
                        tc_mma_bf16(d, a_lo, b_hi, idesc, 1u);                                   // + A_lo B_hi
                    }
                }
                if (leader) {
                    tc_commit(&a_empty[stage]);
                    tc_commit(&b_empty[b_slot]);
                }
                if (++b_slot == RG_B_STAGES) { b_slot = 0; ++b_round; }
                // refill the slot of the PREVIOUS chunk (its MMAs retire while this chunk's run)
                if (!first_piece && loads_left) {
                    mbar_wait(&b_empty[ld_slot], (ld_round - 1) & 1);
                    issue_load();
                }
                first_piece = false;
            }
            if (leader) tc_commit(&d_full[dbuf]);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue -------------------------
        const bool vec = ((ldc & 3) == 0) && ((((uintptr_t)C) & 15) == 0);
        for (int64_t ti = 0; ti < my_tiles; ++ti) {
            const uint32_t dbuf = (uint32_t)(ti & 1);
            mbar_wait_parked(&d_full[dbuf], (uint32_t)((ti >> 1) & 1), 2000);
            tc_fence_after();
            const int64_t r = (blockIdx.x + ti * gridDim.x) * RG_ROWS + warp * 32 + lane;
            const uint32_t dg = tmem_base + dbuf * dcols + ((uint32_t)(warp * 32) << 16);
            for (int cq = 0; cq * 32 < ng; ++cq) {
                float v[32], v2[32];
                tmem_ld32(dg + (uint32_t)(cq * 32), v);
                tmem_ld32(dg + (uint32_t)(ng + cq * 32), v2);
                if (r < R) {
                    float *dst = C + r * ldc + g * NG + cq * 32;
                    const int ncol = (ng - cq * 32 < 32) ? (ng - cq * 32) : 32;
                    if (vec && ncol == 32) {
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            reinterpret_cast<float4 *>(dst)[u] = make_float4(v[4 * u] + v2[4 * u], v[4 * u + 1] + v2[4 * u + 1],
                                                                           v[4 * u + 2] + v2[4 * u + 2], v[4 * u + 3] + v2[4 * u + 3]);
                    } else {
#pragma unroll
                        for (int u = 0; u < 32; ++u)
                            if (u < ncol) dst[u] = v[u] + v2[u];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d_empty[dbuf]);
        }
    }
    __syncthreads();
    if (warp == RG_EWARPS + RG_PWARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

static int rg_ng(int N) { return N < RG_NG_MAX ? N : RG_NG_MAX; }

static size_t rg_smem_bytes(int NG) {
    return (size_t)RG_A_STAGES * RG_A_STAGE + (size_t)RG_B_STAGES * rg_piece_bytes(NG) +
           (2 * RG_A_STAGES + 2 * RG_B_STAGES + 4) * sizeof(uint64_t) + 16;
}

static void rg_set_attributes() {
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(rowgemm_prep_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rg_piece_bytes(RG_NG_MAX));
        cudaFuncSetAttribute(rowgemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rg_smem_bytes(RG_NG_MAX));
        cudaFuncSetAttribute(rowgemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rg_smem_bytes(RG_NG_MAX));
        attr_done.mark();
    }
}

// bytes of the B image; 0 when the shape is outside the kernel (N must be a multiple of 16: the MMA's N granularity)
size_t rowgemm_tc_ws_bytes(int Kd, int N) {
    if (Kd < 1 || N < 16 || (N & 15) != 0) return 0;
    const int NG = rg_ng(N), groups = (N + NG - 1) / NG, nk = (Kd + 63) / 64;
    return (size_t)groups * nk * rg_piece_bytes(NG);
}

// C[R x N] = A[R x Kd] . B with B(k, j) = Bp[k * sbk + j * sbj]; returns 1 when the shape is not supported
int rowgemm_tc(const float *A, int64_t lda, const float *Bp, int64_t sbk, int64_t sbj, int64_t R, int Kd, int N, float *C,
               int64_t ldc, void *ws, size_t ws_bytes, cudaStream_t st) {
    const size_t need = rowgemm_tc_ws_bytes(Kd, N);
    if (need == 0 || ws == nullptr || ws_bytes < need || (((uintptr_t)ws) & 15) != 0 || R < 1) return 1;
    const int NG = rg_ng(N), groups = (N + NG - 1) / NG, nk = (Kd + 63) / 64;
    if (groups > 65535) return 1;
    rg_set_attributes();
    rowgemm_prep_b_kernel<<<groups * nk, 256, rg_piece_bytes(NG), st>>>(Bp, sbk, sbj, Kd, N, NG, nk, (uint8_t *)ws);
    if (int e = finish_launch("rowgemm_prep_b")) return e;
    const int64_t ntiles = (R + RG_ROWS - 1) / RG_ROWS;
    int64_t gx = kNumSMs / groups;
    if (gx < 1) gx = 1;
    if (gx > ntiles) gx = ntiles;
    uint32_t cols = 32;
    while ((int)cols < 4 * NG) cols <<= 1;  // two buffers of 2 * NG columns
    rowgemm_tc_kernel<false><<<dim3((unsigned)gx, (unsigned)groups), RG_THREADS, rg_smem_bytes(NG), st>>>(
        A, lda, (const uint8_t *)ws, R, Kd, N, NG, cols, C, ldc, nk, 0);
    return finish_launch("rowgemm_tc");
}

// ---- dW[R x N] += X^T . B over n points: X = [n x R] (the weighted features), B = [n x N] (the output gradient) -------
// dst[i] += part[0][i] + part[1][i] + ... : eight thread groups add contiguous ranges of the partials in order, the
// range sums are added in range order (fixed order: deterministic)
__global__ void __launch_bounds__(256)
rowgemm_add_parts_kernel(const float *__restrict__ part, int splits, int64_t total, float *__restrict__ dst) {
    __shared__ float sums[8][32];
    const int lane = threadIdx.x & 31, gq = threadIdx.x >> 5;
    const int per = (splits + 7) / 8;
    const int z0 = gq * per, z1 = (z0 + per < splits) ? z0 + per : splits;
    for (int64_t base = (int64_t)blockIdx.x * 32; base < total; base += (int64_t)gridDim.x * 32) {
        const int64_t i = base + lane;
        float a = 0.f;
        if (i < total)
            for (int z = z0; z < z1; ++z) a += __ldg(part + (int64_t)z * total + i);
        sums[gq][lane] = a;
        __syncthreads();
        if (gq == 0 && i < total) {
            float t = sums[0][lane];
#pragma unroll
            for (int u = 1; u < 8; ++u) t += sums[u][lane];
            dst[i] += t;
        }
        __syncthreads();
    }
}

struct RgDwPlan {
    int NG, groups, nk, tiles, splits, cps;
    size_t image, total;
};

static RgDwPlan rg_dw_plan(int64_t n, int R, int N) {
    RgDwPlan p;
    p.NG = rg_ng(N);
    p.groups = (N + p.NG - 1) / p.NG;
    p.nk = (int)((n + 63) / 64);
    p.tiles = (R + RG_ROWS - 1) / RG_ROWS;
    int want = kNumSMs / (p.tiles * p.groups);
    if (want < 1) want = 1;
    if (want > p.nk) want = p.nk;
    p.cps = (p.nk + want - 1) / want;
    p.splits = (p.nk + p.cps - 1) / p.cps;
    p.image = ((size_t)p.groups * p.nk * rg_piece_bytes(p.NG) + 255) / 256 * 256;
    p.total = p.image + (size_t)p.splits * R * N * sizeof(float);
    return p;
}

size_t rowgemm_dw_ws_bytes(int64_t n, int R, int N) {
    if (n < 1 || R < 1 || N < 16 || (N & 15) != 0 || n > (int64_t)1 << 30) return 0;
    return rg_dw_plan(n, R, N).total;
}

// returns 1 when the shape is not supported
int rowgemm_dw(const float *X, const float *Bp, int64_t n, int R, int N, float *dW, void *ws, size_t ws_bytes,
               cudaStream_t st) {
    const size_t need = rowgemm_dw_ws_bytes(n, R, N);
    if (need == 0 || ws == nullptr || ws_bytes < need || (((uintptr_t)ws) & 15) != 0) return 1;
    const RgDwPlan p = rg_dw_plan(n, R, N);
    if (p.groups > 65535 || p.splits > 65535) return 1;
    rg_set_attributes();
    float *part = (float *)((char *)ws + p.image);
    // image of B(k = point, j = column) = Bp[k * N + j]
    rowgemm_prep_b_kernel<<<p.groups * p.nk, 256, rg_piece_bytes(p.NG), st>>>(Bp, N, 1, (int)n, N, p.NG, p.nk, (uint8_t *)ws);
    if (int e = finish_launch("rowgemm_prep_b")) return e;
    uint32_t cols = 32;
    while ((int)cols < 4 * p.NG) cols <<= 1;
    rowgemm_tc_kernel<true><<<dim3((unsigned)p.tiles, (unsigned)p.groups, (unsigned)p.splits), RG_THREADS, rg_smem_bytes(p.NG), st>>>(
        X, R, (const uint8_t *)ws, R, n, N, p.NG, cols, part, N, p.cps, (int64_t)R * N);
    if (int e = finish_launch("rowgemm_dw")) return e;
    const int64_t total = (int64_t)R * N;
    int64_t gr = (total + 31) / 32;
    if (gr > 8 * kNumSMs) gr = 8 * kNumSMs;
    rowgemm_add_parts_kernel<<<(unsigned)gr, 256, 0, st>>>(part, p.splits, total, dW);
    return finish_launch("rowgemm_add_parts");
}

}  // namespace p3d

using namespace p3d;

P3D_EXPORT size_t p3d_rowgemm_ws_bytes(int Kd, int N) { return rowgemm_tc_ws_bytes(Kd, N); }

P3D_EXPORT int p3d_rowgemm_f32(const float *A, int64_t lda, const float *B, int64_t sbk, int64_t sbj, int64_t R, int Kd,
                               int N, float *C, int64_t ldc, void *ws, size_t ws_bytes, void *stream) {
    P3D_REQUIRE(R >= 0 && Kd >= 1 && N >= 1, "rowgemm: bad dimension");
    if (R == 0) return P3D_OK;
    P3D_REQUIRE(A && B && C, "rowgemm: null pointer");
    P3D_REQUIRE(lda >= Kd && ldc >= N, "rowgemm: leading dimension smaller than the row");
    const size_t need = rowgemm_tc_ws_bytes(Kd, N);
    P3D_REQUIRE(need > 0, "rowgemm: the output width must be a multiple of 16 (got %d)", N);
    if (ws == nullptr || ws_bytes < need) {
        set_error("rowgemm: workspace of %zu bytes required (p3d_rowgemm_ws_bytes)", need);
        return P3D_ERR_WORKSPACE;
    }
    const int r = rowgemm_tc(A, lda, B, sbk, sbj, R, Kd, N, C, ldc, ws, ws_bytes, (cudaStream_t)stream);
    if (r == 1) {
        set_error("rowgemm: shape not supported (R=%lld Kd=%d N=%d)", (long long)R, Kd, N);
        return P3D_ERR_INVALID;
    }
    return r;
}

P3D_EXPORT size_t p3d_rowgemm_dw_ws_bytes(int64_t n, int R, int N) { return rowgemm_dw_ws_bytes(n, R, N); }

P3D_EXPORT int p3d_rowgemm_dw_f32(const float *X, const float *B, int64_t n, int R, int N, float *dW, void *ws,
                                  size_t ws_bytes, void *stream) {
    P3D_REQUIRE(n >= 0 && R >= 1 && N >= 1, "rowgemm_dw: bad dimension");
    if (n == 0) return P3D_OK;
    P3D_REQUIRE(X && B && dW, "rowgemm_dw: null pointer");
    const size_t need = rowgemm_dw_ws_bytes(n, R, N);
    P3D_REQUIRE(need > 0, "rowgemm_dw: the output width must be a multiple of 16 (got %d)", N);
    if (ws == nullptr || ws_bytes < need) {
        set_error("rowgemm_dw: workspace of %zu bytes required (p3d_rowgemm_dw_ws_bytes)", need);
        return P3D_ERR_WORKSPACE;
    }
    const int r = rowgemm_dw(X, B, n, R, N, dW, ws, ws_bytes, (cudaStream_t)stream);
    if (r == 1) {
        set_error("rowgemm_dw: shape not supported (n=%lld R=%d N=%d)", (long long)n, R, N);
        return P3D_ERR_INVALID;
    }
    return r;
}
