// MX Linear ("the step either side" of the attention core): the reference's
// mx.Linear forward (microxscaling/mx/linear.py:20-103) for MXINT8 activations and weights,
//     y = A1( A1( MXq(A1(x)) . MXq(A1(W))^T ) + A1(bias) )            A1 = bf16 half-away rounding or identity
// with both operands MX-quantized along in_features in blocks of 32.  Every MXINT8 value c * 2^(e-6)
// is exact in bf16, so the contraction is a bf16 x bf16 -> fp32 GEMM on the tcgen05 tensor cores with
// no product altered (only the fp32 summation order differs from the reference's BLAS).
//
//   k_quantize_gemm_operand  fp32 (rows, K) -> bf16 operand in the MMA-ready order the GEMM streams:
//                            [row tile][k step of 64][8 chunks][tile rows][16 B]  (K-major, no swizzle:
//                            one stage of one tile is ONE contiguous block -> one TMA bulk copy).
//                            One thread per 32-wide MX block, same F2I-free arithmetic as the predictor.
//   k_mx_linear_umma         128 x 256 output tile per CTA, 192 threads, warp-specialised:
//                              warp 0   TMA producer   (cp.async.bulk, ring of stages, full/empty mbarriers)
//                              warp 1   MMA issuer     (4 tcgen05.mma of K = 16 per stage, commit -> empty)
//                              warps 2-5 epilogue      (TMEM lane = output row: A1, + bias, A1, 128-bit stores)
//                            Two CTAs per SM (2 x 256 TMEM columns): one CTA's epilogue overlaps the
//                            other's main loop.
#pragma once
#include "mxprune_predict_tc.cuh"

namespace mxp {

constexpr int GL_BM = 128;            // output rows (tokens) per CTA == TMEM lanes
constexpr int GL_BN = 256;            // output features per CTA == TMEM columns
constexpr int GL_BK = 64;             // K per pipeline stage (8 chunks of 8 bf16)
constexpr int GL_T = 192;

struct GemmOpLayout {
    int ksteps;                       // K / 64
    size_t a_stage, b_stage;          // bytes of one stage of one tile
};
__host__ __device__ inline GemmOpLayout gemm_op_layout(int K) {
    GemmOpLayout L;
    L.ksteps = K / GL_BK;
    L.a_stage = (size_t)8 * GL_BM * 16;
    L.b_stage = (size_t)8 * GL_BN * 16;
    return L;
}

// rows_tile = 128 (activations) or 256 (weights); rows_pad = multiple of rows_tile (zero rows past `rows`)
__global__ void __launch_bounds__(256)
k_quantize_gemm_operand(const float* __restrict__ x, int64_t ld, int rows, int rows_pad, int K, int rows_tile,
                        int bf16, int flush, unsigned char* __restrict__ op) {
    const int nb = K >> 5, ksteps = K / GL_BK;
    const int64_t ntask = (int64_t)rows_pad * nb;
    const size_t stage_bytes = (size_t)8 * rows_tile * 16;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ntask; t += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / rows_pad), row = (int)(t - (int64_t)b * rows_pad);      // block-major: coalesced stores
        uint32_t xv[32];
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows) f = __ldg(reinterpret_cast<const float4*>(x + (int64_t)row * ld + 32 * b) + s);
            xv[4 * s] = __float_as_uint(f.x); xv[4 * s + 1] = __float_as_uint(f.y);
            xv[4 * s + 2] = __float_as_uint(f.z); xv[4 * s + 3] = __float_as_uint(f.w);
        }
        BlockQ r;
        quantize_block_thread<false, false>(xv, 32, bf16, flush, r);
        const int tile = row / rows_tile, rt = row - tile * rows_tile;
        const int kstep = b >> 1, kc0 = (b & 1) * 4;
        unsigned char* dst = op + ((size_t)tile * ksteps + kstep) * stage_bytes + ((size_t)kc0 * rows_tile + rt) * 16;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) *reinterpret_cast<uint4*>(dst + (size_t)ch * rows_tile * 16) = r.op[ch];
    }
}

struct LinearParams {
    const unsigned char *a_op, *w_op;
    const float* bias;                // already A1-rounded by the host wrapper kernel, or null
    float* out;
    int64_t ldo;
    int M, N, K, bf16, stages;
};

// bias -> A1(bias), N floats (tiny)
__global__ void k_round_bias(const float* __restrict__ b, float* __restrict__ o, int N, int bf16) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) o[i] = bf16 ? bf16_half_away(b[i]) : b[i];
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epilogue_barrier() {               // the 128 epilogue threads
    asm volatile("bar.sync 1, 128;" ::: "memory");
}

// Persistent: one CTA per SM walks the output tiles (feature tile fastest, so the CTAs that share an
// activation tile run together and the weight operand stays L2-resident).  The accumulator is
// double-buffered in TMEM (2 x 256 columns): the epilogue of tile i overlaps the main loop of tile i+1.
__global__ void __launch_bounds__(GL_T, 1)
k_mx_linear_umma(const LinearParams p) {
    extern __shared__ __align__(1024) unsigned char smem_gl[];
    const GemmOpLayout L = gemm_op_layout(p.K);
    const int stages = p.stages;
    const size_t stage_bytes = L.a_stage + L.b_stage;
    float* s_bias = reinterpret_cast<float*>(smem_gl + stages * stage_bytes);                 // [2][256]
    float* s_tr = s_bias + 2 * GL_BN;                                                          // [4 warps][32][36]
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem_gl + stages * stage_bytes + 2048 + 4 * 32 * 36 * 4);   // [8]
    uint64_t* bar_empty = bar_full + 8;                                                        // [8]
    uint64_t* bar_acc_full = bar_empty + 8;                                                    // [2]
    uint64_t* bar_acc_empty = bar_acc_full + 2;                                                // [2]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_n = (p.N + GL_BN - 1) / GL_BN, tiles_m = (p.M + GL_BM - 1) / GL_BM;
    const int ntiles = tiles_n * tiles_m;

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_empty[b], 1); }
    }
    if (warp == 1) tmem_alloc(s_tmem, 512u);
    tcgen05_fence_before_sync();
    __syncthreads();
    tcgen05_fence_after_sync();
    const uint32_t tmem = *s_tmem;

    if (warp == 0) {
        // ---------------- TMA producer
        if (lane == 0) {
            int it = 0;                                             // running stage counter across tiles
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int tile_m = t / tiles_n, tile_n = t - tile_m * tiles_n;
                const unsigned char* a_src = p.a_op + (size_t)tile_m * L.ksteps * L.a_stage;
                const unsigned char* b_src = p.w_op + (size_t)tile_n * L.ksteps * L.b_stage;
                for (int ks = 0; ks < L.ksteps; ++ks, ++it) {
                    const int s = it % stages;
                    mbar_wait(&bar_empty[s], (uint32_t)(((it / stages) & 1) ^ 1));      // first lap passes at once
                    unsigned char* dst = smem_gl + (size_t)s * stage_bytes;
                    mbar_expect_tx(&bar_full[s], (uint32_t)stage_bytes);
                    tma_bulk_g2s(dst, a_src + (size_t)ks * L.a_stage, (uint32_t)L.a_stage, &bar_full[s]);
                    tma_bulk_g2s(dst + L.a_stage, b_src + (size_t)ks * L.b_stage, (uint32_t)L.b_stage, &bar_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16_f32(GL_BM, GL_BN);
            int it = 0, i = 0;
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
                const int buf = i & 1;
                mbar_wait(&bar_acc_empty[buf], (uint32_t)(((i >> 1) & 1) ^ 1));         // epilogue has drained this buffer
                tcgen05_fence_after_sync();
                const uint32_t acc = tmem + (uint32_t)(buf * GL_BN);
                for (int ks = 0; ks < L.ksteps; ++ks, ++it) {
                    const int s = it % stages;
                    mbar_wait(&bar_full[s], (uint32_t)((it / stages) & 1));
                    tcgen05_fence_after_sync();
                    const unsigned char* sa = smem_gl + (size_t)s * stage_bytes;
                    const unsigned char* sb = sa + L.a_stage;
#pragma unroll
                    for (int k4 = 0; k4 < GL_BK / 16; ++k4) {
                        const uint64_t da = umma_smem_desc(smem_u32(sa + (size_t)(2 * k4) * GL_BM * 16), GL_BM * 16, 128);
                        const uint64_t db = umma_smem_desc(smem_u32(sb + (size_t)(2 * k4) * GL_BN * 16), GL_BN * 16, 128);
                        umma_bf16_ss(acc, da, db, idesc, ks > 0 || k4 > 0);
                    }
                    umma_commit(&bar_empty[s]);                     // stage free once these MMAs have read it
                }
                umma_commit(&bar_acc_full[buf]);                    // accumulator of this tile complete
            }
        }
    } else {
        // ---------------- epilogue: warps 2..5 -> TMEM lane quarter (warp & 3), thread = output row
        const int q = warp & 3;
        const int et = tid - 64;                                    // 0..127
        const bool bf16 = p.bf16 != 0;
        int i = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++i) {
            const int tile_m = t / tiles_n, tile_n = t - tile_m * tiles_n;
            const int n0 = tile_n * GL_BN, buf = i & 1;
            float* sb = s_bias + buf * GL_BN;
            for (int j = et; j < GL_BN; j += 128) sb[j] = (p.bias && n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
            epilogue_barrier();
            mbar_wait(&bar_acc_full[buf], (uint32_t)((i >> 1) & 1));
            tcgen05_fence_after_sync();
            const uint32_t t0 = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * GL_BN);
            // Coalesced stores: each warp transposes its 32 rows x 32 columns through shared memory
            // (row pitch 36 words: conflict-free 128-bit writes and reads), so that 8 lanes write 128
            // contiguous bytes of one output row instead of 32 lanes writing 16 bytes of 32 rows.
            float* tr = s_tr + (warp - 2) * (32 * 36);
            const int lr = lane >> 3, lc = (lane & 7) * 4;          // this lane's row (mod 4) and column group
            const int row_base = tile_m * GL_BM + q * 32;
#pragma unroll 1
            for (int c0 = 0; c0 < GL_BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(t0 + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int v = 0; v < 8; ++v)
                    *reinterpret_cast<uint4*>(tr + lane * 36 + 4 * v) = make_uint4(r[4 * v], r[4 * v + 1], r[4 * v + 2], r[4 * v + 3]);
                __syncwarp();
                const float4 b4 = *reinterpret_cast<const float4*>(sb + c0 + lc);
                const bool col_ok = n0 + c0 + lc < p.N;             // N is a multiple of 4 (checked on the host)
#pragma unroll
                for (int i8 = 0; i8 < 8; ++i8) {
                    const int rl = lr + 4 * i8, row = row_base + rl;
                    float4 y = *reinterpret_cast<const float4*>(tr + rl * 36 + lc);
                    if (bf16) { y.x = bf16_half_away(y.x); y.y = bf16_half_away(y.y); y.z = bf16_half_away(y.z); y.w = bf16_half_away(y.w); }   // A1, linear.py:85-87
                    if (p.bias) {
                        y.x = __fadd_rn(y.x, b4.x); y.y = __fadd_rn(y.y, b4.y); y.z = __fadd_rn(y.z, b4.z); y.w = __fadd_rn(y.w, b4.w);
                        if (bf16) { y.x = bf16_half_away(y.x); y.y = bf16_half_away(y.y); y.z = bf16_half_away(y.z); y.w = bf16_half_away(y.w); }   // :89-93
                    }
                    if (row < p.M && col_ok) *reinterpret_cast<float4*>(p.out + (int64_t)row * p.ldo + n0 + c0 + lc) = y;
                }
                __syncwarp();
            }
            tcgen05_fence_before_sync();
            epilogue_barrier();                                     // every epilogue thread has read its lanes
            if (et == 0) mbar_arrive(&bar_acc_empty[buf]);
        }
    }
    tcgen05_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512u);
}

}  // namespace mxp
