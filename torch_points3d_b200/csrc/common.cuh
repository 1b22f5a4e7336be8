// common.cuh -- shared helpers for libp3d_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/p3d_b200.h"

#define P3D_EXPORT extern "C" __attribute__((visibility("default")))

namespace p3d {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- error / bookkeeping (defined in cabi.cu) ------------------------------------------
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int finish_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return P3D_ERR_CUDA;
    }
    return P3D_OK;
}

// One-time kernel opt-ins (cudaFuncSetAttribute) are PER DEVICE: a process that drives several GPUs must repeat them
// on each.  `done()` / `mark()` keep one bit per device ordinal; a racing second thread at worst repeats the
// (idempotent) attribute call.
struct DeviceOnce {
    std::atomic<uint64_t> bits[4] = {};
    static int dev() {
        int d = 0;
        cudaGetDevice(&d);
        return d & 255;
    }
    bool done() const { const int d = dev(); return (bits[d >> 6].load(std::memory_order_acquire) >> (d & 63)) & 1u; }
    void mark() { const int d = dev(); bits[d >> 6].fetch_or((uint64_t)1 << (d & 63), std::memory_order_release); }
};

#define P3D_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            p3d::set_error(__VA_ARGS__); \
            return P3D_ERR_INVALID;     \
        }                               \
    } while (0)

// ---- the one fp32 squared distance every operator uses --------
// ((dx*dx + dy*dy) + dz*dz) with explicit round-to-nearest intrinsics: the compiler may
// not contract these into FMAs, so results are bit-identical to the -ffp-contract=off oracle.
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx);
    const float dy = __fsub_rn(ay, by);
    const float dz = __fsub_rn(az, bz);
    float s = __fmul_rn(dx, dx);
    s = __fadd_rn(s, __fmul_rn(dy, dy));
    s = __fadd_rn(s, __fmul_rn(dz, dz));
    return s;
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streaming (evict-first) vector stores for write-once outputs
__device__ __forceinline__ void st_cs_f4(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

// lower_bound / upper_bound over a sorted int64 vector, executed by a whole warp
// (32-ary search: each step probes 32 positions; ~log32(n) dependent loads).
__device__ __forceinline__ int64_t warp_lower_bound(const int64_t *__restrict__ v, int64_t n, int64_t key, int lane) {
    int64_t lo = 0, hi = n;  // invariant: answer in [lo, hi]
    while (lo < hi) {
        const int64_t step = (hi - lo + 31) / 32;
        const int64_t pos = lo + (int64_t)lane * step;  // probes lo, lo+step, ...
        const bool lt = (pos < hi) ? (v[pos] < key) : false;
        const int c = __popc(__ballot_sync(0xffffffffu, lt));  // sorted -> first c probes are < key
        if (c == 0) {
            hi = lo;  // v[lo] >= key
        } else {      // v[lo+(c-1)*step] < key <= v[lo+c*step] (when that probe exists)
            const int64_t nhi = lo + (int64_t)c * step;
            lo = lo + (int64_t)(c - 1) * step + 1;
            hi = nhi < hi ? nhi : hi;
        }
    }
    return lo;
}

}  // namespace p3d
