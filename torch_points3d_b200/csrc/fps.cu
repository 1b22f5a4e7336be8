// fps.cu -- furthest point sampling for sm_100a.
//
// Replaces tp.furthest_point_sample (core/spatial_ops/sampling.py:100) and
// torch_geometric.nn.fps / torch_cluster.fps (sampling.py:63) -- semantics in SURVEY.md A.1:
// start index 0, mind = 1e10, per round mind = min(mind, d2(p, last)), next = first argmax.
//
// Design (latency-bound: S-1 dependent argmax rounds per cloud):
//   * one CTA -- or one thread-block CLUSTER of up to 16 CTAs -- per cloud;
//   * every thread keeps its PPT points AND their running min-distances in registers for
//     the whole kernel: HBM is touched once (12 B/point in, 4|8 B/sample out);
//   * per round: register update -> warp argmax with two redux.sync (max of the fp32 bit
//     pattern, then min index among the maxima == lowest-index tie-break) -> one 20-byte
//     record per warp in shared memory -> ONE __syncthreads -> every warp reduces the <=32
//     records redundantly, so the winner's coordinates are known everywhere without a
//     second barrier (records are double-buffered by round parity);
//   * clusters: each CTA's winner record is pushed into every peer's shared memory through
//     DSMEM, then one barrier.cluster per round.
#include <atomic>

#include <cooperative_groups.h>
#include <math_constants.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace p3d {

struct __align__(8) FpsRec {
    unsigned key;  // fp32 bit pattern of the candidate's min-distance (>= +0 -> order preserving)
    unsigned idx;  // local index inside the cloud, 0xffffffff = none
    float x, y, z;
    unsigned pad;
};

__device__ __forceinline__ void warp_argmax(unsigned key, unsigned idx, unsigned &wkey, unsigned &widx) {
    wkey = __reduce_max_sync(0xffffffffu, key);
    const unsigned cand = (key == wkey) ? idx : 0xffffffffu;
    widx = __reduce_min_sync(0xffffffffu, cand);
}

template <int T, int PPT, int CLUSTER, typename IdxT>
__global__ void __launch_bounds__(T, 1)
fps_reg_kernel(const float *__restrict__ xyz, const int64_t *__restrict__ ptr, int64_t n_dense,
               const int64_t *__restrict__ out_ptr, int64_t s_dense, IdxT *__restrict__ out,
               const int64_t *__restrict__ first, float *__restrict__ new_xyz) {
    constexpr int NW = T / 32;
    extern __shared__ float csm[];  // this CTA's coordinates: cx | cy | cz, T * PPT floats each (slot = j * T + tid)
    float *cx = csm, *cy = csm + T * PPT, *cz = csm + 2 * T * PPT;
    __shared__ FpsRec rec[2][32];
    __shared__ FpsRec crec[2][CLUSTER];

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int cloud = blockIdx.x / CLUSTER;
    const int rank = blockIdx.x % CLUSTER;

    int64_t start, n, s, obase, add;
    if (ptr != nullptr) {
        start = ptr[cloud];
        n = ptr[cloud + 1] - start;
        obase = out_ptr[cloud];
        s = out_ptr[cloud + 1] - obase;
        add = start;  // CSR form emits global indices
    } else {
        start = (int64_t)cloud * n_dense;
        n = n_dense;
        obase = (int64_t)cloud * s_dense;
        s = s_dense;
        add = 0;
    }
    if (s > n) s = n;  // host validates; keeps the kernel safe

    const float *__restrict__ p = xyz + 3 * start;
    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int64_t i = (int64_t)j * (CLUSTER * T) + rank * T + tid;
        if (i < n) {
            px[j] = p[3 * i];
            py[j] = p[3 * i + 1];
            pz[j] = p[3 * i + 2];
            md[j] = 1e10f;
        } else {
            px[j] = py[j] = pz[j] = 0.f;
            md[j] = -1.f;  // never selected: min(-1, d) stays -1 and -1 > best is false
        }
        cx[j * T + tid] = px[j];
        cy[j * T + tid] = py[j];
        cz[j * T + tid] = pz[j];
    }
    float lx = 0.f, ly = 0.f, lz = 0.f;
    int64_t f0 = 0;
    if (n > 0) {
        if (first != nullptr) {
            f0 = first[cloud];
            f0 = f0 < 0 ? 0 : (f0 >= n ? n - 1 : f0);
        }
        lx = p[3 * f0];
        ly = p[3 * f0 + 1];
        lz = p[3 * f0 + 2];
    }
    if (tid == 0 && rank == 0 && s > 0) {
        out[obase] = (IdxT)(add + f0);
        if (new_xyz) { new_xyz[3 * obase] = lx; new_xyz[3 * obase + 1] = ly; new_xyz[3 * obase + 2] = lz; }
    }

    for (int64_t r = 1; r < s; ++r) {
        const int buf = (int)(r & 1);
        float best = -1.f;
        int bj = 0;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const float d = sqdist3(px[j], py[j], pz[j], lx, ly, lz);
            const float m = fminf(md[j], d);
            md[j] = m;
            if (m > best) { best = m; bj = j; }  // strict: the lowest index wins inside a thread (indices ascend with j)
        }
        const unsigned bi = best >= 0.f ? (unsigned)(bj * (CLUSTER * T) + rank * T + tid) : 0xffffffffu;
        const unsigned key = best >= 0.f ? __float_as_uint(best) : 0u;
        unsigned wkey, widx;
        warp_argmax(key, bi, wkey, widx);
        if (key == wkey && bi == widx) {  // the winning lane fetches its point's coordinates from shared memory
            FpsRec rr;
            const int slot = bj * T + tid;
            rr.key = wkey; rr.idx = widx; rr.x = cx[slot]; rr.y = cy[slot]; rr.z = cz[slot]; rr.pad = 0;
            rec[buf][warp] = rr;
        }
        __syncthreads();

        unsigned k2 = 0u, i2 = 0xffffffffu;
        if (lane < NW) {
            k2 = rec[buf][lane].key;
            i2 = rec[buf][lane].idx;
        }
        unsigned ckey, cidx;
        warp_argmax(k2, i2, ckey, cidx);
        int src = __ffs(__ballot_sync(0xffffffffu, k2 == ckey && i2 == cidx)) - 1;

        if constexpr (CLUSTER == 1) {
            lx = rec[buf][src].x;
            ly = rec[buf][src].y;
            lz = rec[buf][src].z;
            if (tid == 0) {
                out[obase + r] = (IdxT)(add + (int64_t)cidx);
                if (new_xyz) { float *o = new_xyz + 3 * (obase + r); o[0] = lx; o[1] = ly; o[2] = lz; }
            }
        } else {
            cg::cluster_group cluster = cg::this_cluster();
            if (warp == 0 && lane < CLUSTER) {
                FpsRec rr = rec[buf][src];
                rr.key = ckey;
                rr.idx = cidx;
                FpsRec *dst = cluster.map_shared_rank(&crec[buf][rank], lane);
                *dst = rr;
            }
            cluster.sync();
            unsigned k3 = 0u, i3 = 0xffffffffu;
            if (lane < CLUSTER) {
                k3 = crec[buf][lane].key;
                i3 = crec[buf][lane].idx;
            }
            unsigned gkey, gidx;
            warp_argmax(k3, i3, gkey, gidx);
            src = __ffs(__ballot_sync(0xffffffffu, k3 == gkey && i3 == gidx)) - 1;
            lx = crec[buf][src].x;
            ly = crec[buf][src].y;
            lz = crec[buf][src].z;
            if (tid == 0 && rank == 0) {
                out[obase + r] = (IdxT)(add + (int64_t)gidx);
                if (new_xyz) { float *o = new_xyz + 3 * (obase + r); o[0] = lx; o[1] = ly; o[2] = lz; }
            }
        }
    }
    if constexpr (CLUSTER > 1) {
        // no CTA may exit while a peer can still write into its shared memory
        cg::this_cluster().sync();
    }
}

// ---- large clouds: cluster kernel with exact bounding-box pruning and a one-hop record exchange -------------------
// One cloud per cluster of CLUSTER CTAs.  Every thread owns PPT CONSECUTIVE points of the cloud: their running
// min-distances live in registers, their coordinates in the CTA's shared memory (slot = j * T + tid: conflict-free),
// plus the bounding box of the block and the (value, slot) of its current farthest point.
//   * pruning: a round's new sample can only lower a min-distance if it is closer to the point than that
//     min-distance.  With d_box = squared distance from the sample to the block's bounding box -- computed with the
//     same fp32 operations as sqdist3, so d_box <= d2(sample, p) holds bit-exactly for every p of the block (fp32
//     subtraction, multiplication and addition are monotone) -- a block with d_box >= max(min-distances) is untouched
//     by the round and keeps its cached farthest point.  After a few hundred samples almost every block is skipped;
//     how many depends on how compact 32 consecutive points are (scan order, voxel order: very; random order: the
//     kernel degrades to the unpruned update, results never change).
//   * exchange: the CTA's winner (warp argmax -> one record per warp in shared memory -> __syncthreads -> warp 0 reduces
//     them) is pushed as ONE 32-byte record (key, index, coordinates) into every CTA of the cluster with st.async (a
//     remote shared-memory store that signals the receiver's mbarrier with complete_tx); a CTA then waits on its OWN
//     mbarrier for the CLUSTER records and every warp reduces them redundantly.  One remote hop per round and no
//     barrier.cluster on the critical path.  (Pushing every WARP's record to every CTA instead -- no __syncthreads, no
//     CTA-level reduction -- was measured slower: 256 transaction updates per round serialise on each CTA's mbarrier.)
//   Records and barriers are double-buffered by round parity: a CTA can be at most one round ahead of a peer.
struct __align__(16) FpsRec32 {
    unsigned key, idx;
    float x, y, z;
    unsigned pad[3];
};

__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
                 "r"(a), "r"(b), "r"(c), "r"(d), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void fps_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fps_mbar_expect(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
// The default (.acquire.cta) form, as cluster-transaction barriers are waited on elsewhere (TMA multicast): the
// records land in THIS CTA's shared memory and their bytes are counted on THIS CTA's barrier, so a completed phase
// means they are here.  The .acquire.cluster form compiles to an L1 invalidate (CCTL.IVALL) after every successful
// wait, in every warp, every round (2 % of the round).
//
// What a round costs (clock64 instrumentation, one 120 000-point cloud, 8 CTAs): ~2000 cycles, of which the remote hop
// is ~150 (a run with every CTA exchanging with itself only is 8 % faster); the rest is the CTA-local chain -- the
// slowest warp's block update, then the barrier, warp 0's reduction and the second reduction of every warp.  Variants
// measured and NOT kept (no change within 2 %): 1024 threads x 16 points, tagged 16-byte records polled in shared
// memory instead of st.async + mbarrier, skipping the warp arg-max of untouched warps, batched coordinate loads.
// Freezing the min-distances (no update at all) gives 0.54 us per round: half of a round is the block updates of
// whichever warps the new sample touches.  Transposing the blocks (lane j = point j of 32 blocks, a touched block
// updated by the whole warp with a REDUX for its new maximum) was bit-exact but slower (1.26 us; 3.2 us on shuffled
// clouds): the per-block REDUX -> ballot -> select chains of one warp do not overlap.
__device__ __forceinline__ void fps_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FPS_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FPS_DONE_%=;\n"
        "bra FPS_WAIT_%=;\n"
        "FPS_DONE_%=:\n"
        "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
        "r"(parity)
        : "memory");
}

template <int T, int PPT, int CLUSTER, typename IdxT>
__global__ void __launch_bounds__(T, 1)
fps_pruned_kernel(const float *__restrict__ xyz, const int64_t *__restrict__ ptr, int64_t n_dense,
                  const int64_t *__restrict__ out_ptr, int64_t s_dense, IdxT *__restrict__ out,
                  const int64_t *__restrict__ first, float *__restrict__ new_xyz) {
    constexpr int NW = T / 32;
    static_assert(CLUSTER <= 32 && NW <= 32, "one record per lane");
    extern __shared__ float csm[];               // cx | cy | cz, T * PPT floats each (slot = j * T + tid)
    float *cx = csm, *cy = csm + T * PPT, *cz = csm + 2 * T * PPT;
    __shared__ FpsRec32 recs[2][CLUSTER];        // the CTA winners of a round, one per cluster rank
    __shared__ uint4 wrec[2][NW];                // this CTA's warp winners: key, index, slot
    __shared__ __align__(8) uint64_t bars[2];

    cg::cluster_group cluster = cg::this_cluster();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x / CLUSTER;
    const int rank = blockIdx.x % CLUSTER;

    int64_t start, n, s, obase, add;
    if (ptr != nullptr) {
        start = ptr[cloud];
        n = ptr[cloud + 1] - start;
        obase = out_ptr[cloud];
        s = out_ptr[cloud + 1] - obase;
        add = start;
    } else {
        start = (int64_t)cloud * n_dense;
        n = n_dense;
        obase = (int64_t)cloud * s_dense;
        s = s_dense;
        add = 0;
    }
    if (s > n) s = n;

    const float *__restrict__ p = xyz + 3 * start;
    // this thread's block of consecutive points: [blk, blk + PPT)
    const int64_t blk = ((int64_t)rank * T + tid) * PPT;
    float md[PPT];
    float bx0 = CUDART_INF_F, by0 = CUDART_INF_F, bz0 = CUDART_INF_F, bx1 = -CUDART_INF_F, by1 = -CUDART_INF_F,
          bz1 = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int64_t i = blk + j;
        float x = 0.f, y = 0.f, z = 0.f;
        if (i < n) {
            x = p[3 * i]; y = p[3 * i + 1]; z = p[3 * i + 2];
            md[j] = 1e10f;
            bx0 = fminf(bx0, x); by0 = fminf(by0, y); bz0 = fminf(bz0, z);
            bx1 = fmaxf(bx1, x); by1 = fmaxf(by1, y); bz1 = fmaxf(bz1, z);
        } else {
            md[j] = -1.f;  // never selected: min(-1, d) stays -1
        }
        cx[j * T + tid] = x; cy[j * T + tid] = y; cz[j * T + tid] = z;
    }
    float mybest = blk < n ? 1e10f : -1.f;   // max of md[] (exact), and the first slot that holds it
    int myj = 0;

    if (tid == 0) {
        fps_mbar_init(&bars[0], 1);
        fps_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float lx = 0.f, ly = 0.f, lz = 0.f;
    int64_t f0 = 0;
    if (n > 0) {
        if (first != nullptr) {
            f0 = first[cloud];
            f0 = f0 < 0 ? 0 : (f0 >= n ? n - 1 : f0);
        }
        lx = p[3 * f0]; ly = p[3 * f0 + 1]; lz = p[3 * f0 + 2];
    }
    if (tid == 0 && rank == 0 && s > 0) {
        out[obase] = (IdxT)(add + f0);
        if (new_xyz) { new_xyz[3 * obase] = lx; new_xyz[3 * obase + 1] = ly; new_xyz[3 * obase + 2] = lz; }
    }
    cluster.sync();  // every CTA's barriers are initialised before the first remote store can arrive

    // remote addresses of this CTA's record slot and of the barriers in CTA `lane` (warp 0, lanes < CLUSTER push)
    const uint32_t my_rec0 = (uint32_t)__cvta_generic_to_shared(&recs[0][rank]);
    const uint32_t my_rec1 = (uint32_t)__cvta_generic_to_shared(&recs[1][rank]);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&bars[0]);
    const uint32_t bar1 = (uint32_t)__cvta_generic_to_shared(&bars[1]);
    uint32_t r_rec[2] = {0u, 0u}, r_bar[2] = {0u, 0u};
    if (lane < CLUSTER) {
        r_rec[0] = mapa_shared(my_rec0, (uint32_t)lane);
        r_rec[1] = mapa_shared(my_rec1, (uint32_t)lane);
        r_bar[0] = mapa_shared(bar0, (uint32_t)lane);
        r_bar[1] = mapa_shared(bar1, (uint32_t)lane);
    }

    for (int64_t r = 1; r < s; ++r) {
        const int buf = (int)(r & 1);
        if (tid == 0) fps_mbar_expect(&bars[buf], CLUSTER * (uint32_t)sizeof(FpsRec32));
        // squared distance from the new sample to this block's bounding box, in sqdist3's operation order
        {
            const float ex = fmaxf(fmaxf(__fsub_rn(bx0, lx), __fsub_rn(lx, bx1)), 0.f);
            const float ey = fmaxf(fmaxf(__fsub_rn(by0, ly), __fsub_rn(ly, by1)), 0.f);
            const float ez = fmaxf(fmaxf(__fsub_rn(bz0, lz), __fsub_rn(lz, bz1)), 0.f);
            const float dbox = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
            if (dbox < mybest) {
                float best = -1.f;
                int bj = 0;
#pragma unroll
                for (int j = 0; j < PPT; ++j) {
                    const float d = sqdist3(cx[j * T + tid], cy[j * T + tid], cz[j * T + tid], lx, ly, lz);
                    const float m = fminf(md[j], d);
                    md[j] = m;
                    if (m > best) { best = m; bj = j; }  // strict: the lowest index wins inside a thread
                }
                mybest = best;
                myj = bj;
            }
        }
        // Point indices grow with the lane (blocks of consecutive points), with the warp and with the cluster rank, so
        // at all three levels "largest key, lowest index on ties" is "largest key, lowest lane on ties": one REDUX and
        // a ballot instead of two dependent REDUX -- the round is a latency chain, every link counts.
        // Keys: fp32 bit patterns of non-negative distances order like unsigned integers, + 1 so that 0 = "no point".
        const unsigned key = mybest >= 0.f ? __float_as_uint(mybest) + 1u : 0u;
        const unsigned wkey = __reduce_max_sync(0xffffffffu, key);
        const int wl = __ffs(__ballot_sync(0xffffffffu, key == wkey)) - 1;
        if (lane == wl) wrec[buf][warp] = make_uint4(wkey, (unsigned)(blk + myj), (unsigned)(myj * T + tid), 0u);
        __syncthreads();
        if (warp == 0) {
            // CTA winner, pushed to every CTA of the cluster (this one included) by lanes < CLUSTER
            uint4 w = make_uint4(0u, 0xffffffffu, 0u, 0u);
            if (lane < NW) w = wrec[buf][lane];
            const unsigned ckey = __reduce_max_sync(0xffffffffu, w.x);
            const int cl = __ffs(__ballot_sync(0xffffffffu, w.x == ckey)) - 1;
            const unsigned cidx = __shfl_sync(0xffffffffu, w.y, cl);
            const unsigned cslot = __shfl_sync(0xffffffffu, w.z, cl);
            const float wx = cx[cslot], wy = cy[cslot], wz = cz[cslot];
            if (lane < CLUSTER) {
                st_async_v4(r_rec[buf], ckey, cidx, __float_as_uint(wx), __float_as_uint(wy), r_bar[buf]);
                st_async_v4(r_rec[buf] + 16u, __float_as_uint(wz), 0u, 0u, 0u, r_bar[buf]);
            }
        }
        fps_mbar_wait(&bars[buf], (uint32_t)(((r - 1) >> 1) & 1));  // rounds 1, 2 are the first use of bars[1], bars[0]
        // every warp reduces the CLUSTER records
        unsigned k2 = 0u, i2 = 0xffffffffu;
        if (lane < CLUSTER) {
            const uint2 kv = *reinterpret_cast<const uint2 *>(&recs[buf][lane]);
            k2 = kv.x;
            i2 = kv.y;
        }
        const unsigned gkey = __reduce_max_sync(0xffffffffu, k2);
        const int src = __ffs(__ballot_sync(0xffffffffu, k2 == gkey)) - 1;
        const unsigned gidx = __shfl_sync(0xffffffffu, i2, src);
        lx = recs[buf][src].x;
        ly = recs[buf][src].y;
        lz = recs[buf][src].z;
        if (tid == 0 && rank == 0) {
            out[obase + r] = (IdxT)(add + (int64_t)gidx);
            if (new_xyz) { float *o = new_xyz + 3 * (obase + r); o[0] = lx; o[1] = ly; o[2] = lz; }
        }
    }
    cluster.sync();  // no CTA may exit while a peer's stores to it can still be in flight
}

// Single-CTA variant (clouds of up to 8192 points): besides the register copy used for the distance update, the
// coordinates sit in shared memory, so a round tracks only (best, slot) per point -- 12 instructions per point --
// the per-warp record is 8 bytes, and every thread fetches the winner's coordinates by index after the one
// barrier of the round.  A round is issue-bound inside its SM (warps x instructions), so every instruction shaved
// off the per-point update shortens all S-1 dependent rounds.
template <int T, int PPT, typename IdxT>
__global__ void __launch_bounds__(T, 1)
fps_cta_kernel(const float *__restrict__ xyz, const int64_t *__restrict__ ptr, int64_t n_dense,
               const int64_t *__restrict__ out_ptr, int64_t s_dense, IdxT *__restrict__ out,
               const int64_t *__restrict__ first, float *__restrict__ new_xyz) {
    constexpr int NW = T / 32;
    extern __shared__ float fsm[];  // sx | sy | sz, T * PPT floats each
    float *sx = fsm, *sy = fsm + T * PPT, *sz = fsm + 2 * T * PPT;
    __shared__ uint2 rec[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    int64_t start, n64, s64, obase, add;
    if (ptr != nullptr) {
        start = ptr[cloud];
        n64 = ptr[cloud + 1] - start;
        obase = out_ptr[cloud];
        s64 = out_ptr[cloud + 1] - obase;
        add = start;
    } else {
        start = (int64_t)cloud * n_dense;
        n64 = n_dense;
        obase = (int64_t)cloud * s_dense;
        s64 = s_dense;
        add = 0;
    }
    if (s64 > n64) s64 = n64;
    const int n = (int)n64, s = (int)s64;
    const float *__restrict__ p = xyz + 3 * start;
    float px[PPT], py[PPT], pz[PPT], md[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int i = j * T + tid;
        if (i < n) {
            px[j] = p[3 * i]; py[j] = p[3 * i + 1]; pz[j] = p[3 * i + 2];
            md[j] = 1e10f;
            sx[i] = px[j]; sy[i] = py[j]; sz[i] = pz[j];
        } else {
            px[j] = py[j] = pz[j] = 0.f;
            md[j] = -1.f;  // never selected
        }
    }
    int f0 = 0;
    if (n > 0 && first != nullptr) {
        const int64_t f = first[cloud];
        f0 = f < 0 ? 0 : (f >= n ? n - 1 : (int)f);
    }
    __syncthreads();
    float lx = 0.f, ly = 0.f, lz = 0.f;
    if (n > 0) { lx = sx[f0]; ly = sy[f0]; lz = sz[f0]; }
    IdxT *__restrict__ orow = out + obase;
    float *__restrict__ xrow = new_xyz ? new_xyz + 3 * obase : nullptr;
    if (tid == 0 && s > 0) {
        orow[0] = (IdxT)(add + f0);
        if (xrow) { xrow[0] = lx; xrow[1] = ly; xrow[2] = lz; }
    }
    for (int r = 1; r < s; ++r) {
        const int buf = r & 1;
        float best = -1.f;
        int bj = 0;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const float d = sqdist3(px[j], py[j], pz[j], lx, ly, lz);
            const float m = fminf(md[j], d);
            md[j] = m;
            if (m > best) { best = m; bj = j; }  // strict: the lowest index wins inside a thread
        }
        const unsigned bi = best >= 0.f ? (unsigned)(bj * T + tid) : 0xffffffffu;
        const unsigned key = best >= 0.f ? __float_as_uint(best) : 0u;
        unsigned wkey, widx;
        warp_argmax(key, bi, wkey, widx);
        if (lane == 0) rec[buf][warp] = make_uint2(wkey, widx);
        __syncthreads();
        unsigned k2 = 0u, i2 = 0xffffffffu;
        if (lane < NW) {
            const uint2 rr = rec[buf][lane];
            k2 = rr.x;
            i2 = rr.y;
        }
        unsigned ckey, cidx;
        warp_argmax(k2, i2, ckey, cidx);
        lx = sx[cidx]; ly = sy[cidx]; lz = sz[cidx];
        if (tid == 0) {
            orow[r] = (IdxT)(add + (int64_t)cidx);
            if (xrow) { xrow[3 * r] = lx; xrow[3 * r + 1] = ly; xrow[3 * r + 2] = lz; }
        }
    }
}

template <int T, int PPT, typename IdxT>
static int launch_cta(const float *xyz, const int64_t *ptr, int64_t n_dense, const int64_t *out_ptr, int64_t s_dense,
                      IdxT *out, const int64_t *first, float *new_xyz, int clouds, cudaStream_t st) {
    auto kern = fps_cta_kernel<T, PPT, IdxT>;
    const size_t smem = (size_t)3 * T * PPT * sizeof(float);
    if (smem + 2048 > 48 * 1024) {  // static shared memory counts against the 48 KB default as well
        static DeviceOnce done;
        if (!done.done()) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            done.mark();
        }
    }
    kern<<<clouds, T, smem, st>>>(xyz, ptr, n_dense, out_ptr, s_dense, out, first, new_xyz);
    return finish_launch("fps_cta");
}

// Fallback for clouds larger than 131072 points: one 1024-thread CTA per cloud, points read
// from global/L2 every round, min-distances in a caller-provided workspace.
template <typename IdxT>
__global__ void __launch_bounds__(1024, 1)
fps_big_kernel(const float *__restrict__ xyz, const int64_t *__restrict__ ptr, int64_t n_dense,
               const int64_t *__restrict__ out_ptr, int64_t s_dense, IdxT *__restrict__ out,
               float *__restrict__ mind, const int64_t *__restrict__ first, float *__restrict__ new_xyz) {
    __shared__ FpsRec rec[2][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    int64_t start, n, s, obase, add;
    if (ptr != nullptr) {
        start = ptr[cloud]; n = ptr[cloud + 1] - start;
        obase = out_ptr[cloud]; s = out_ptr[cloud + 1] - obase; add = start;
    } else {
        start = (int64_t)cloud * n_dense; n = n_dense;
        obase = (int64_t)cloud * s_dense; s = s_dense; add = 0;
    }
    if (s > n) s = n;
    const float *__restrict__ p = xyz + 3 * start;
    float *__restrict__ md = mind + start;
    for (int64_t i = tid; i < n; i += 1024) md[i] = 1e10f;
    float lx = 0.f, ly = 0.f, lz = 0.f;
    int64_t f0 = 0;
    if (n > 0) {
        if (first != nullptr) { f0 = first[cloud]; f0 = f0 < 0 ? 0 : (f0 >= n ? n - 1 : f0); }
        lx = p[3 * f0]; ly = p[3 * f0 + 1]; lz = p[3 * f0 + 2];
    }
    if (tid == 0 && s > 0) {
        out[obase] = (IdxT)(add + f0);
        if (new_xyz) { new_xyz[3 * obase] = lx; new_xyz[3 * obase + 1] = ly; new_xyz[3 * obase + 2] = lz; }
    }
    for (int64_t r = 1; r < s; ++r) {
        const int buf = (int)(r & 1);
        float best = -1.f, bx = 0.f, by = 0.f, bz = 0.f;
        unsigned bi = 0xffffffffu;
        for (int64_t i = tid; i < n; i += 1024) {
            const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
            const float m = fminf(md[i], sqdist3(x, y, z, lx, ly, lz));
            md[i] = m;
            if (m > best) { best = m; bi = (unsigned)i; bx = x; by = y; bz = z; }
        }
        const unsigned key = best >= 0.f ? __float_as_uint(best) : 0u;
        unsigned wkey, widx;
        warp_argmax(key, bi, wkey, widx);
        if (key == wkey && bi == widx) {
            FpsRec rr;
            rr.key = wkey; rr.idx = widx; rr.x = bx; rr.y = by; rr.z = bz; rr.pad = 0;
            rec[buf][warp] = rr;
        }
        __syncthreads();
        const unsigned k2 = rec[buf][lane].key, i2 = rec[buf][lane].idx;
        unsigned ckey, cidx;
        warp_argmax(k2, i2, ckey, cidx);
        const int src = __ffs(__ballot_sync(0xffffffffu, k2 == ckey && i2 == cidx)) - 1;
        lx = rec[buf][src].x; ly = rec[buf][src].y; lz = rec[buf][src].z;
        if (tid == 0) {
            out[obase + r] = (IdxT)(add + (int64_t)cidx);
            if (new_xyz) { float *o = new_xyz + 3 * (obase + r); o[0] = lx; o[1] = ly; o[2] = lz; }
        }
    }
}

template <int T, int PPT, int CLUSTER, typename IdxT>
static int launch_reg(const float *xyz, const int64_t *ptr, int64_t n_dense, const int64_t *out_ptr,
                      int64_t s_dense, IdxT *out, const int64_t *first, float *new_xyz, int clouds, cudaStream_t st) {
    auto kern = fps_reg_kernel<T, PPT, CLUSTER, IdxT>;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(clouds * CLUSTER));
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = (size_t)3 * T * PPT * sizeof(float);
    if (cfg.dynamicSmemBytes + 4096 > 48 * 1024) {
        static DeviceOnce smem_done;
        if (!smem_done.done()) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
            smem_done.mark();
        }
    }
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (CLUSTER > 8) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) {
            set_error("fps: non-portable cluster size refused: %s", cudaGetErrorString(e));
            return P3D_ERR_CUDA;
        }
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xyz, ptr, n_dense, out_ptr, s_dense, out, first, new_xyz);
    if (e != cudaSuccess) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        set_error("fps launch (T=%d PPT=%d CLUSTER=%d): %s", T, PPT, CLUSTER, cudaGetErrorString(e));
        return P3D_ERR_CUDA;
    }
    return finish_launch("fps");
}

// launch configuration of a pruned variant (function attributes are set once per device)
template <int T, int PPT, int CLUSTER, typename IdxT>
static void pruned_config(int clouds, cudaStream_t st, cudaLaunchConfig_t &cfg, cudaLaunchAttribute *attr) {
    auto kern = fps_pruned_kernel<T, PPT, CLUSTER, IdxT>;
    cfg.gridDim = dim3((unsigned)(clouds * CLUSTER));
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = (size_t)3 * T * PPT * sizeof(float);
    static DeviceOnce attr_done;
    if (!attr_done.done()) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
        if (CLUSTER > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        attr_done.mark();
    }
    cfg.stream = st;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
}

// how many clusters of this variant the device runs at once (0 when the cluster size cannot be placed); per device
template <int T, int PPT, int CLUSTER, typename IdxT>
static int pruned_active_clusters() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c != 0) return c > 0 ? c : 0;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    pruned_config<T, PPT, CLUSTER, IdxT>(1, nullptr, cfg, attr);
    int num = 0;
    if (cudaOccupancyMaxActiveClusters(&num, fps_pruned_kernel<T, PPT, CLUSTER, IdxT>, &cfg) != cudaSuccess) {
        cudaGetLastError();
        num = 0;
    }
    cache[dev].store(num > 0 ? num : -1, std::memory_order_relaxed);
    return num;
}

template <int T, int PPT, int CLUSTER, typename IdxT>
static int launch_pruned(const float *xyz, const int64_t *ptr, int64_t n_dense, const int64_t *out_ptr,
                         int64_t s_dense, IdxT *out, const int64_t *first, float *new_xyz, int clouds, cudaStream_t st) {
    auto kern = fps_pruned_kernel<T, PPT, CLUSTER, IdxT>;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    pruned_config<T, PPT, CLUSTER, IdxT>(clouds, st, cfg, attr);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, xyz, ptr, n_dense, out_ptr, s_dense, out, first, new_xyz);
    if (e != cudaSuccess) {
        g_launches.fetch_add(1, std::memory_order_relaxed);
        set_error("fps launch (pruned, T=%d PPT=%d CLUSTER=%d): %s", T, PPT, CLUSTER, cudaGetErrorString(e));
        return P3D_ERR_CUDA;
    }
    return finish_launch("fps_pruned");
}

template <typename IdxT>
static int fps_dispatch(const float *xyz, const int64_t *ptr, int64_t n_dense, const int64_t *out_ptr,
                        int64_t s_dense, IdxT *out, const int64_t *first, float *new_xyz, int clouds, int64_t max_n,
                        float *ws, size_t ws_bytes, int64_t total_rows, cudaStream_t st) {
    if (clouds == 0) return P3D_OK;
#define P3D_FPS_CASE(LIM, T, PPT, CL) \
    if (max_n <= (LIM)) return launch_reg<T, PPT, CL, IdxT>(xyz, ptr, n_dense, out_ptr, s_dense, out, first, new_xyz, clouds, st);
#define P3D_FPS_CTA(LIM, T, PPT) \
    if (max_n <= (LIM)) return launch_cta<T, PPT, IdxT>(xyz, ptr, n_dense, out_ptr, s_dense, out, first, new_xyz, clouds, st);
    // variants picked on B200 (profiles/r1_fps_tuning.md): few fat threads win, a round is bound by
    // warps x instructions issued inside the one SM
    P3D_FPS_CTA(256, 64, 4)
    P3D_FPS_CTA(512, 128, 4)
    P3D_FPS_CTA(1024, 128, 8)
    P3D_FPS_CTA(2048, 128, 16)
    P3D_FPS_CTA(4096, 256, 16)
    P3D_FPS_CTA(8192, 512, 16)
#undef P3D_FPS_CTA
    P3D_FPS_CASE(16384, 1024, 8, 2)
    P3D_FPS_CASE(32768, 1024, 8, 4)
    // 32 769..131 072 points: 8-CTA clusters with bounding-box pruning and the one-hop record exchange
    // A round is a latency chain whose longest link is the block update (PPT points by one lane): a 16-CTA cluster
    // halves the block (0.82 vs 1.02 us per round at 120 000 points), but a 16-CTA cluster needs 16 free SMs of ONE
    // GPC -- the device places only a handful at once (cudaOccupancyMaxActiveClusters), so it is used when that
    // takes fewer waves: few clouds per launch (one LiDAR scan per GPU), not the batch of 8.
    if (max_n <= 131072) {
        const bool small = max_n <= 65536;
        const int a8 = small ? pruned_active_clusters<512, 16, 8, IdxT>() : pruned_active_clusters<512, 32, 8, IdxT>();
        const int a16 = small ? pruned_active_clusters<512, 8, 16, IdxT>() : pruned_active_clusters<512, 16, 16, IdxT>();
        bool wide = false;
        if (a16 > 0) {
            const double t16 = 0.8 * ((clouds + a16 - 1) / a16);
            const double t8 = a8 > 0 ? 1.0 * ((clouds + a8 - 1) / a8) : 1e30;
            wide = t16 < t8;
        }
        if (small) {
            if (wide) return launch_pruned<512, 8, 16, IdxT>(xyz, ptr, n_dense, out_ptr, s_dense, out, first, new_xyz, clouds, st);
            return launch_pruned<512, 16, 8, IdxT>(xyz, ptr, n_dense, out_ptr, s_dense, out, first, new_xyz, clouds, st);
        }
        if (wide) return launch_pruned<512, 16, 16, IdxT>(xyz, ptr, n_dense, out_ptr, s_dense, out, first, new_xyz, clouds, st);
        return launch_pruned<512, 32, 8, IdxT>(xyz, ptr, n_dense, out_ptr, s_dense, out, first, new_xyz, clouds, st);
    }
#undef P3D_FPS_CASE
    if (ws == nullptr || ws_bytes < (size_t)total_rows * sizeof(float)) {
        set_error("fps: clouds of more than 131072 points need a workspace of %lld bytes",
                  (long long)(total_rows * (int64_t)sizeof(float)));
        return P3D_ERR_WORKSPACE;
    }
    fps_big_kernel<IdxT><<<clouds, 1024, 0, st>>>(xyz, ptr, n_dense, out_ptr, s_dense, out, ws, first, new_xyz);
    return finish_launch("fps_big");
}

}  // namespace p3d

using namespace p3d;

P3D_EXPORT size_t p3d_fps_dense_ws_bytes(int B, int N) {
    return N > 131072 ? (size_t)B * (size_t)N * sizeof(float) : 0;
}

P3D_EXPORT int p3d_fps_dense_f32(const float *xyz, int B, int N, int S, void *idx, int idx_is_i64, float *new_xyz,
                                 void *ws, size_t ws_bytes, void *stream) {
    P3D_REQUIRE(B >= 0 && N >= 0 && S >= 0, "fps_dense: negative dimension");
    P3D_REQUIRE(S <= N, "fps_dense: npoint (%d) > N (%d)", S, N);
    if (B == 0 || S == 0) return P3D_OK;
    P3D_REQUIRE(xyz && idx, "fps_dense: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (idx_is_i64)
        return fps_dispatch<int64_t>(xyz, nullptr, N, nullptr, S, (int64_t *)idx, nullptr, new_xyz, B, N, (float *)ws, ws_bytes,
                                     (int64_t)B * N, st);
    return fps_dispatch<int32_t>(xyz, nullptr, N, nullptr, S, (int32_t *)idx, nullptr, new_xyz, B, N, (float *)ws, ws_bytes,
                                 (int64_t)B * N, st);
}

P3D_EXPORT size_t p3d_fps_csr_ws_bytes(int64_t total_rows, int64_t max_n) {
    return max_n > 131072 ? (size_t)total_rows * sizeof(float) : 0;
}

P3D_EXPORT int p3d_fps_csr_f32(const float *xyz, const int64_t *ptr, int num_clouds, const int64_t *out_ptr,
                               int64_t *idx, const int64_t *start, int64_t max_n, void *ws, size_t ws_bytes,
                               void *stream) {
    P3D_REQUIRE(num_clouds >= 0 && max_n >= 0, "fps_csr: negative dimension");
    if (num_clouds == 0) return P3D_OK;
    P3D_REQUIRE(xyz && ptr && out_ptr && idx, "fps_csr: null pointer");
    return fps_dispatch<int64_t>(xyz, ptr, 0, out_ptr, 0, idx, start, nullptr, num_clouds, max_n, (float *)ws, ws_bytes,
                                 ws_bytes / sizeof(float), (cudaStream_t)stream);
}
