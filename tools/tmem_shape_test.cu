// Register layout of tcgen05.ld.16x256b: which (TMEM lane, column) does register k of thread T hold?
// TMEM is filled with lane*1000 + column through 32x32b stores, then read back with 16x256b.x2.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tmem_shape_test tools/tmem_shape_test.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include "../mx_quantization_b200/csrc/mxprune_umma.cuh"
using namespace mxp;

__global__ void __launch_bounds__(128) k(float* out) {
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    tcgen05_fence_before_sync();
    __syncthreads();
    tcgen05_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t mine = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t r[32];
    for (int c = 0; c < 32; ++c) r[c] = __float_as_uint((float)(tid * 1000 + c));
    tmem_st_32x32b_x32(mine, r);
    tmem_st_wait();
    tcgen05_fence_before_sync();
    __syncthreads();
    tcgen05_fence_after_sync();
    for (int half = 0; half < 2; ++half) {
        uint32_t q[8];
        const uint32_t addr = mine + ((uint32_t)(half * 16) << 16) + 8;      // lanes [16 half, +16), columns 8..23
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
                     : "r"(addr));
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) out[(half * 128 + tid) * 8 + i] = __uint_as_float(q[i]);
    }
    for (int half = 0; half < 2; ++half) {       // 16x32bx2: lanes [16 half, +16); threads 0-15 columns 4.., threads 16-31 columns 4+20..
        uint32_t q[4];
        const uint32_t addr = mine + ((uint32_t)(half * 16) << 16) + 4;
        asm volatile("tcgen05.ld.sync.aligned.16x32bx2.x4.b32 {%0,%1,%2,%3}, [%4], 20;"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]) : "r"(addr));
        tmem_ld_wait();
        for (int i = 0; i < 4; ++i) out[2048 + (half * 128 + tid) * 4 + i] = __uint_as_float(q[i]);
    }
    tcgen05_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
    float* d; cudaMalloc(&d, (2048 + 1024) * 4);
    k<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    static float h[2048 + 1024]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int half = 0; half < 2; ++half)
        for (int t : {0, 1, 2, 3, 4, 5, 31, 32, 33, 37, 127}) {
            printf("half %d thread %3d:", half, t);
            for (int i = 0; i < 8; ++i) { int v = (int)h[(half * 128 + t) * 8 + i]; printf(" (l%d,c%d)", v / 1000, v % 1000); }
            printf("\n");
        }
    // check the conjectured layout: reg i of thread T (lane base L0): lane L0 + T%32/4 + 8*((i>>1)&1), col 8 + 8*(i>>2) + 2*(T%4) + (i&1)
    int bad = 0;
    for (int half = 0; half < 2; ++half)
        for (int t = 0; t < 128; ++t)
            for (int i = 0; i < 8; ++i) {
                const int lane = (t / 32) * 32 + half * 16 + (t % 32) / 4 + 8 * ((i >> 1) & 1);
                const int col = 8 + 8 * (i >> 2) + 2 * (t % 4) + (i & 1);
                if ((int)h[(half * 128 + t) * 8 + i] != lane * 1000 + col) ++bad;
            }
    printf("conjectured layout mismatches: %d\n", bad);
    int bad2 = 0;
    for (int half = 0; half < 2; ++half)
        for (int t = 0; t < 128; ++t) {
            if (t % 32 < 3 || t % 32 == 16 || t % 32 == 31) {
                printf("16x32bx2 half %d thread %3d:", half, t);
                for (int i = 0; i < 4; ++i) { int v = (int)h[2048 + (half * 128 + t) * 4 + i]; printf(" (l%d,c%d)", v / 1000, v % 1000); }
                printf("\n");
            }
            for (int i = 0; i < 4; ++i) {
                const int lane = (t / 32) * 32 + half * 16 + (t % 16);
                const int col = 4 + i + ((t % 32) >= 16 ? 20 : 0);
                if ((int)h[2048 + (half * 128 + t) * 4 + i] != lane * 1000 + col) ++bad2;
            }
        }
    printf("16x32bx2 conjectured layout mismatches: %d\n", bad2);
    return 0;
}
