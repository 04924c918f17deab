// Standalone check of the hand-written tcgen05 path used by the exact-attention kernel:
// D[128 x N] (fp32, TMEM) = A[128 x K] * B[N x K]^T with bf16 operands in shared memory,
// K-major, no swizzle (8x16B core matrices, SBO = 128 B, LBO = rows*16 B).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_test tools/umma_test.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../mx_quantization_b200/csrc/mxprune_umma.cuh"

using namespace mxp;

template <int N, int K>
__global__ void __launch_bounds__(128) umma_gemm(const float* __restrict__ A, const float* __restrict__ B,
                                                 float* __restrict__ D) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int KC = K / 8;                         // 16-byte chunks along K
    unsigned char* sA = smem;                         // [KC][128][16 B]
    unsigned char* sB = smem + KC * 128 * 16;         // [KC][N][16 B]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    // operands: thread t writes row t of A; rows of B are spread over the threads
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

int main() {
    constexpr int N = 208, K = 64;
    std::vector<float> A(128 * K), B(N * K), D(128 * N), R(128 * N);
    srand(1);
    for (auto& x : A) x = (float)((rand() % 255) - 127) * (1.0f / 64) * (float)(1 << (rand() % 3));
    for (auto& x : B) x = (float)((rand() % 255) - 127) * (1.0f / 64) * (float)(1 << (rand() % 3));
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[m * K + k] * B[n * K + k];
            R[m * N + n] = (float)s;
        }
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const int smem = (K / 8) * (128 + N) * 16;
    cudaFuncSetAttribute(umma_gemm<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    umma_gemm<N, K><<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int i = 0; i < 128 * N; ++i) {
        double err = fabs((double)D[i] - R[i]);
        if (err > maxerr) maxerr = err;
        if (D[i] != R[i]) ++bad;
    }
    printf("max abs err %.6g, mismatching entries %d of %d; D[0]=%g R[0]=%g D[last]=%g R[last]=%g\n", maxerr, bad,
           128 * N, D[0], R[0], D[128 * N - 1], R[128 * N - 1]);
    return (e == cudaSuccess && bad == 0) ? 0 : 1;
}
