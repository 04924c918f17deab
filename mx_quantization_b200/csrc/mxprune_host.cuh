// Host-side plumbing shared by the translation units of libmxprune: thread-local error text and
// launch counter, the per-device shared-memory opt-in, the path switches.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "mxprune.h"

namespace mxp {

inline thread_local char g_err[512] = "";
inline thread_local int g_launches = 0;

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(MXP_E_CUDA, "%s: %s", what, cudaGetErrorString(e));
    ++g_launches;
    return MXP_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: remember, per kernel, the set
// of devices it has been raised on (one bit per device ordinal; ordinals >= 64 simply set it on every launch).
template <typename Kernel>
inline int ensure_dyn_smem(Kernel kernel, int bytes, std::atomic<uint64_t>& done) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail(MXP_E_CUDA, "cudaGetDevice failed");
    const uint64_t bit = dev < 64 ? 1ull << dev : 0ull;
    if (bit && (done.load(std::memory_order_acquire) & bit)) return MXP_OK;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) return fail(MXP_E_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    if (bit) done.fetch_or(bit, std::memory_order_release);
    return MXP_OK;
}
// one static device set per call site (== per kernel instantiation when used inside a template)
#define MXP_ENSURE_DYN_SMEM(kernel, bytes)                                   \
    do {                                                                     \
        static std::atomic<uint64_t> done_{0};                               \
        const int rc_ = ::mxp::ensure_dyn_smem(kernel, bytes, done_);        \
        if (rc_) return rc_;                                                 \
    } while (0)

inline int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
        return n;
    return 148;
}

// process-wide path switches (test / A-B aids); atomics so that concurrent callers never see a torn value
inline std::atomic<int> g_attn_path{0};      // 0 = tcgen05 tensor-core path (default), 1 = CUDA-core dp4a path
inline std::atomic<int> g_predict_path{0};   // 0 = tensor-core scoring (default where it applies), 1 = CUDA-core XOR/POPC kernel
inline std::atomic<int> g_fused_path{1};     // 1 = fused / sparse kernels where they apply (default), 0 = the three-kernel path

constexpr size_t SMEM_PER_SM = 232448;       // 227 KiB usable per SM
constexpr size_t SMEM_2CTA = SMEM_PER_SM / 2 - 1024;   // dynamic bytes a CTA may use with two CTAs per SM

}  // namespace mxp
