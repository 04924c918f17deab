// Does a TMA tensor map over the STRIDED (B,H,N,hd) fp32 view of a fused qkv buffer do what K1
// needs?  dims {32 floats, N, nfull, H, B} (non-monotonic strides), box {32, 64 rows, nfull},
// SWIZZLE_128B, OOB rows zero-filled; plus an unswizzled tail map for hd % 32 != 0.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_test tools/tma_test.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#include "../mx_quantization_b200/csrc/mxprune_umma.cuh"
using namespace mxp;

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5,%6}], [%7];" ::
        "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4),
        "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2,%3,%4,%5}], [%6];" ::
        "r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3),
        "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(128) k(const __grid_constant__ CUtensorMap tm_main,
                                         const __grid_constant__ CUtensorMap tm_tail, int row0, int h, int b,
                                         int nfull, int tail, float* out_main, float* out_tail) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    float* s_main = reinterpret_cast<float*>(smem);
    float* s_tail = reinterpret_cast<float*>(smem + 64 * nfull * 128);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        const uint32_t bytes = 64 * nfull * 128 + 64 * tail * 4;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        tma_load_5d(s_main, &tm_main, 0, row0, 0, h, b, &bar);
        if (tail) tma_load_4d(s_tail, &tm_tail, 0, row0, h, b, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < 64 * nfull * 32; i += 128) out_main[i] = s_main[i];
    for (int i = threadIdx.x; i < 64 * tail; i += 128) out_tail[i] = s_tail[i];
}

int main() {
    const int B = 2, N = 197, H = 3, hd = 72, nfull = hd / 32, tail = hd % 32;
    const size_t total = (size_t)B * N * 3 * H * hd;
    std::vector<float> h(total);
    for (size_t i = 0; i < total; ++i) h[i] = (float)i;
    float* d; cudaMalloc(&d, total * 4); cudaMemcpy(d, h.data(), total * 4, cudaMemcpyHostToDevice);
    const int64_t sN = 3 * H * hd, sH = hd, sB = (int64_t)N * 3 * H * hd;
    float* kptr = d + 1 * H * hd;          // the k view
    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    printf("entry point: %s qres %d fn %p\n", cudaGetErrorString(ce), (int)qres, fn);
    EncodeTiled enc = (EncodeTiled)fn;
    CUtensorMap tm_main, tm_tail;
    {
        cuuint64_t dims[5] = {32, (cuuint64_t)N, (cuuint64_t)nfull, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[4] = {(cuuint64_t)sN * 4, 128, (cuuint64_t)sH * 4, (cuuint64_t)sB * 4};
        cuuint32_t box[5] = {32, 64, (cuuint32_t)nfull, 1, 1}, es[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&tm_main, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, kptr, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode main: %d\n", (int)r);
        if (r) return 1;
    }
    {
        cuuint64_t dims[4] = {(cuuint64_t)tail, (cuuint64_t)N, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)sN * 4, (cuuint64_t)sH * 4, (cuuint64_t)sB * 4};
        cuuint32_t box[4] = {(cuuint32_t)tail, 64, 1, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tm_tail, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, kptr + 32 * nfull, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode tail: %d\n", (int)r);
        if (r) return 1;
    }
    float *om, *ot; cudaMalloc(&om, 64 * nfull * 128); cudaMalloc(&ot, 64 * 32 * 4);
    const int smem = 64 * nfull * 128 + 64 * tail * 4 + 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int bad_total = 0;
    for (int row0 : {0, 128, 192}) {
        const int hh = 2, bb = 1;
        k<<<1, 128, smem>>>(tm_main, tm_tail, row0, hh, bb, nfull, tail, om, ot);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<float> rm(64 * nfull * 32), rt(64 * tail);
        cudaMemcpy(rm.data(), om, rm.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(rt.data(), ot, rt.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int bk = 0; bk < nfull; ++bk)
            for (int r = 0; r < 64; ++r)
                for (int c = 0; c < 8; ++c)
                    for (int t = 0; t < 4; ++t) {
                        const int rblk = bk * 64 + r;                       // [block][row] 128-byte rows
                        const float got = rm[rblk * 32 + ((c ^ (rblk & 7)) * 4) + t];
                        const int row = row0 + r;
                        const float want = row < N ? (float)((size_t)(kptr - d) + bb * sB + hh * sH + row * sN + bk * 32 + c * 4 + t) : 0.f;
                        if (got != want) { if (bad < 3) printf("  main mismatch row %d blk %d c %d t %d got %g want %g\n", row, bk, c, t, got, want); ++bad; }
                    }
        for (int r = 0; r < 64; ++r)
            for (int t = 0; t < tail; ++t) {
                const int row = row0 + r;
                const float want = row < N ? (float)((size_t)(kptr - d) + bb * sB + hh * sH + row * sN + 32 * nfull + t) : 0.f;
                if (rt[r * tail + t] != want) { if (bad < 6) printf("  tail mismatch row %d t %d got %g want %g\n", row, t, rt[r * tail + t], want); ++bad; }
            }
        printf("row0 %3d: %s, mismatches %d\n", row0, cudaGetErrorString(e), bad);
        bad_total += bad;
    }
    return bad_total != 0;
}
