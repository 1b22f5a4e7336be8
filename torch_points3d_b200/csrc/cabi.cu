// cabi.cu -- process-wide bookkeeping behind the C ABI of libp3d_b200.so
// (error text, launch counter, version).  The operator entry points themselves live next to
// their kernels (fps.cu, neighbors.cu, gather.cu, kpconv.cu).
#include "common.cuh"

namespace p3d {

std::atomic<uint64_t> g_launches{0};

static thread_local char t_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

}  // namespace p3d

P3D_EXPORT int p3d_version(void) { return 100; }

P3D_EXPORT const char *p3d_last_error(void) { return p3d::t_error; }

P3D_EXPORT uint64_t p3d_launch_count(void) { return p3d::g_launches.load(std::memory_order_relaxed); }

P3D_EXPORT const char *p3d_arch(void) { return "sm_100a"; }
