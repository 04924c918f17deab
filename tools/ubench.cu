// Instruction-throughput microbenchmarks for the integer ops the predictor leans on
// (POPC, LOP3, FFMA, I2F, IDP4A, REDUX, VOTE, SHFL, LDS).  Prints ops/clk/SM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench tools/ubench.cu && ./ubench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITERS = 2048;
constexpr int CHAINS = 8;

template <int OP>
__global__ void __launch_bounds__(256) bench(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t x[CHAINS];
    float f[CHAINS];
    __shared__ uint32_t sm[256 * 2];
    sm[threadIdx.x] = threadIdx.x; sm[threadIdx.x + 256] = seed;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) { x[c] = seed * (c + 1) + threadIdx.x; f[c] = (float)x[c]; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < CHAINS; ++c) {
            if (OP == 0) x[c] = __popc(x[c]) + seed;                    // POPC (+IADD)
            if (OP == 1) x[c] = (x[c] ^ seed) & (x[c] >> 1 | seed);     // LOP3/SHF
            if (OP == 2) f[c] = fmaf(f[c], 1.0001f, 0.5f);              // FFMA
            if (OP == 3) f[c] = (float)(int)(x[c] += seed) + f[c];      // I2F (+IADD+FADD)
            if (OP == 4) x[c] = __dp4a((int)x[c], (int)seed, (int)x[c]); // IDP4A
            if (OP == 5) x[c] = __reduce_add_sync(0xffffffffu, x[c]) + seed;   // REDUX
            if (OP == 6) x[c] = __ballot_sync(0xffffffffu, x[c] & 1) + x[c];   // VOTE
            if (OP == 7) x[c] = __shfl_xor_sync(0xffffffffu, x[c], 1) + seed;  // SHFL
            if (OP == 8) x[c] = sm[(x[c] & 255)] + seed;                // LDS (dependent)
            if (OP == 9) x[c] = (x[c] >= seed) ? x[c] + 1 : x[c] - 3;   // ISETP+SEL/IADD
            if (OP == 10) x[c] = __reduce_max_sync(0xffffffffu, x[c]) + seed;  // REDUX.MAX
            if (OP == 11) f[c] = (float)(int)x[c]; x[c] += (OP == 11) ? __float_as_uint(f[c]) : 0; // I2F only-ish
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc += x[c] + __float_as_uint(f[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int blocks_per_sm) {
    int sms = 148;
    int blocks = sms * blocks_per_sm;
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&cyc, blocks * 8);
    bench<OP><<<blocks, 256>>>(out, 12345u, cyc);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<OP><<<blocks, 256>>>(out, 12345u, cyc);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[148 * 8]; cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
    double ops_per_block = (double)ITERS * CHAINS * 256;
    // all blocks_per_sm blocks of an SM run concurrently: SM-level rate
    double per_clk_sm = ops_per_block * blocks_per_sm / avg;
    printf("%-10s blocks/SM=%d  %.1f lane-ops/clk/SM   (%.3f ms, %.0f cyc, %.2f GHz eff)\n", name, blocks_per_sm,
           per_clk_sm, ms, avg, avg / (ms * 1e6));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int bps : {2, 4}) {
        run<0>("POPC+IADD", bps);
        run<1>("LOP3", bps);
        run<2>("FFMA", bps);
        run<3>("I2F+2", bps);
        run<4>("IDP4A", bps);
        run<5>("REDUX.ADD", bps);
        run<10>("REDUX.MAX", bps);
        run<6>("VOTE", bps);
        run<7>("SHFL", bps);
        run<8>("LDS", bps);
        run<9>("ISETP+SEL", bps);
    }
    return 0;
}
