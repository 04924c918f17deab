// Is tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM) EXACT for the predictor's operands?
// Operands are +-2^e with one exponent per 32-wide block; the exact score is an integer multiple
// of 2^g.  For exponent windows of increasing width this counts entries where the tensor-core
// result differs from the exact sum (double).  Tells how wide a window the accumulator keeps.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_exact_test tools/umma_exact_test.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../mx_quantization_b200/csrc/mxprune_umma.cuh"

using namespace mxp;

template <int N, int K>
__global__ void __launch_bounds__(128) umma_gemm(const float* __restrict__ A, const float* __restrict__ B,
                                                 float* __restrict__ D) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int KC = K / 8;
    unsigned char* sA = smem;
    unsigned char* sB = smem + KC * 128 * 16;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int kc = 0; kc < KC; ++kc) {
        uint32_t w[4];
        for (int h = 0; h < 4; ++h)
            w[h] = pack_bf16_trunc(A[tid * K + kc * 8 + 2 * h], A[tid * K + kc * 8 + 2 * h + 1]);
        *reinterpret_cast<uint4*>(sA + (kc * 128 + tid) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    for (int n = tid; n < N; n += 128)
        for (int kc = 0; kc < KC; ++kc) {
            uint32_t w[4];
            for (int h = 0; h < 4; ++h)
                w[h] = pack_bf16_trunc(B[n * K + kc * 8 + 2 * h], B[n * K + kc * 8 + 2 * h + 1]);
            *reinterpret_cast<uint4*>(sB + (kc * N + n) * 16) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    if (tid == 0) mbar_init(&bar, 1);
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    fence_proxy_async_smem();
    tcgen05_fence_before_sync();
    __syncthreads();
    tcgen05_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc_bf16_f32(128, N);
        for (int ks = 0; ks < K / 16; ++ks) {
            const uint64_t da = umma_smem_desc(smem_u32(sA + (2 * ks) * 128 * 16), 128 * 16, 128);
            const uint64_t db = umma_smem_desc(smem_u32(sB + (2 * ks) * N * 16), N * 16, 128);
            umma_bf16_ss(tmem, da, db, idesc, ks > 0);
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tcgen05_fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int c = 0; c < 16; ++c) D[tid * N + c0 + c] = __uint_as_float(r[c]);
    }
    tcgen05_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

template <int N, int K>
int run(int hd, int spread, int base, unsigned seed) {
    std::vector<float> A(128 * K, 0.f), B(N * K, 0.f), D(128 * N);
    std::vector<double> R(128 * N);
    srand(seed);
    const int nb = (hd + 31) / 32;
    auto fill = [&](std::vector<float>& X, int rows) {
        for (int r = 0; r < rows; ++r)
            for (int b = 0; b < nb; ++b) {
                const int e = base + (spread ? rand() % (spread + 1) : 0);
                for (int d = 32 * b; d < hd && d < 32 * b + 32; ++d)
                    X[r * K + d] = ((rand() & 1) ? -1.f : 1.f) * ldexpf(1.f, e);
            }
    };
    fill(A, 128); fill(B, N);
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
            R[m * N + n] = s;
        }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    const int smem = (K / 8) * (128 + N) * 16;
    cudaFuncSetAttribute(umma_gemm<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_gemm<N, K><<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0, bad_fp32 = 0;
    for (int i = 0; i < 128 * N; ++i) {
        if ((double)D[i] != R[i]) ++bad;
        if (D[i] != (float)R[i]) ++bad_fp32;       // differs even from the correctly rounded fp32
    }
    printf("hd %2d K %2d N %3d  per-operand spread %2d (pair window %2d bits) base %4d: %s  not-exact %6d  not-RN(fp32) %6d of %d\n",
           hd, K, N, spread, 2 * spread, base, cudaGetErrorString(e), bad, bad_fp32, 128 * N);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad;
}

int main() {
    for (int sp : {0, 2, 4, 6, 7, 8, 9, 10, 11, 12, 14, 16}) run<208, 64>(64, sp, 0, 1 + sp);
    for (int sp : {0, 2, 4, 6, 7, 8, 9, 10, 11, 12, 14, 16}) run<256, 80>(72, sp, 0, 100 + sp);
    for (int base : {-60, -30, 30, 55}) run<256, 80>(72, 6, base, 7);
    return 0;
}
