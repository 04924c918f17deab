// K2: exact MXINT8 attention over the kept keys on the Blackwell tensor cores (Nk <= 256).
//
// Key fact: every MXINT8 value c * 2^(e-6) (|c| <= 127) is EXACTLY representable in bf16 (8
// significant bits, fp32 exponent range).  So both contractions of the reference,
//     true_scores = mx.matmul(q, k^T)            microxscaling/mx/matmul.py:68-88
//     x           = mx.matmul(attn, v)           workloads/deit/scripts/main.py:152
// are products of exact bf16 operands accumulated in fp32 - which is precisely
// tcgen05.mma kind::f16 with bf16 inputs and fp32 accumulators in tensor memory.  No product is
// altered; only the fp32 summation order differs from MKL's (as between any two BLAS).
//
// One CTA (128 threads) per (head, row split).  Thread t owns query row t of the tile == TMEM
// lane t, so the softmax / P-quantisation epilogue needs no cross-thread communication at all.
//   stage (per head)  K codes/exps -> bf16 B-operand chunks;  V fp32 -> A1 -> MXINT8 along TOKENS
//                     (32-token windows per column, matmul.py:76-83) -> bf16 B-operand chunks (V^T)
//   per 128-row tile  Q codes -> bf16 A operand; S = Q.K^T  (hd/16 MMAs, M=128, N=Nk)  -> TMEM
//                     pass A: max over kept keys;  pass B: E = exp(s - max) kept / 0 pruned, sum,
//                     E written back to TMEM;  pass C: P = E/sum -> A1 -> MXINT8 per 32-key window
//                     of original positions -> bf16 A operand;  O += P_w . V_w (2 MMAs per window)
//                     O (fp32, TMEM) -> A1 -> global
// Operand layout in shared memory: K-major, no swizzle; a 16-byte chunk holds 8 consecutive K
// elements of one row; chunk (row r, k-chunk c) lives at c*(ROWS*16) + r*16, i.e. SBO = 128 B,
// LBO = ROWS*16 B.  Threads write one chunk each with consecutive r -> conflict-free 128-bit stores.
#pragma once
#include "mxprune_device.cuh"
#include "mxprune_umma.cuh"

namespace mxp {

constexpr int K2T = 128;
constexpr int K2_PW = 4;                       // P windows buffered per MMA group (4 x 32 keys)
constexpr int K2_P_BYTES = K2_PW * 4 * K2T * 16;   // 32 KiB; also holds the Q tile (<= 16 chunks)

struct AttnParams {
    const int8_t *q_codes, *q_exps, *k_codes, *k_exps;
    View v;
    const uint32_t* mask;
    int B, H, Nq, Nk, hd;
    float scale;
    int bf16, flush;
    float* out;
    int64_t o_sB, o_sH, o_sN;
};

struct K2Smem {
    int nkp, hdp, nw, tmem_cols;
    size_t off_v, off_p, total;
};

__host__ __device__ inline K2Smem k2_smem_layout(int Nk, int hd) {
    K2Smem L;
    L.nkp = (Nk + 15) & ~15;
    L.hdp = (hd + 15) & ~15;
    L.nw = (Nk + 31) >> 5;
    int need = L.nkp > L.hdp ? L.nkp : L.hdp;
    int c = 32;
    while (c < need) c <<= 1;
    L.tmem_cols = c;
    size_t o = (size_t)(L.hdp / 8) * L.nkp * 16;        // K operand
    L.off_v = o; o += (size_t)(L.nw * 4) * L.hdp * 16;  // V^T operand
    L.off_p = o; o += K2_P_BYTES;                       // P window group / Q tile
    L.total = o;
    return L;
}

// 8 int8 codes (two words) of one MX block with exponent weight w = 2^(e-6) -> 8 bf16 in a uint4
__device__ __forceinline__ uint4 dequant8_bf16(uint32_t lo, uint32_t hi, float w) {
    float f[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        f[t] = (float)(int)(signed char)((lo >> (8 * t)) & 0xff) * w;
        f[4 + t] = (float)(int)(signed char)((hi >> (8 * t)) & 0xff) * w;
    }
    return make_uint4(pack_bf16_trunc(f[0], f[1]), pack_bf16_trunc(f[2], f[3]), pack_bf16_trunc(f[4], f[5]),
                      pack_bf16_trunc(f[6], f[7]));
}

// r[w] for a runtime w without dynamic register indexing (a 3-level select tree)
__device__ __forceinline__ uint32_t pick8(const uint32_t (&r)[8], int w) {
    const uint32_t a = (w & 1) ? r[1] : r[0], b = (w & 1) ? r[3] : r[2];
    const uint32_t c = (w & 1) ? r[5] : r[4], d = (w & 1) ? r[7] : r[6];
    const uint32_t ab = (w & 2) ? b : a, cd = (w & 2) ? d : c;
    return (w & 4) ? cd : ab;
}

__global__ void __launch_bounds__(K2T)
k_attend_umma(const AttnParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_s, bar_o;
    __shared__ uint32_t tmem_base_s;
    const int Nk = p.Nk, Nq = p.Nq, hd = p.hd;
    const K2Smem L = k2_smem_layout(Nk, hd);
    const int nkp = L.nkp, hdp = L.hdp, NW = L.nw, NB = (hd + 31) >> 5;
    unsigned char* sK = smem;
    unsigned char* sV = smem + L.off_v;
    unsigned char* sP = smem + L.off_p;
    const int head = blockIdx.x, bb = head / p.H, hh = head % p.H;
    const int tid = threadIdx.x, warp = tid >> 5;
    const bool bf16 = p.bf16, flush = p.flush;

    if (tid == 0) { mbar_init(&bar_s, 1); mbar_init(&bar_o, 1); }
    if (warp == 0) tmem_alloc(&tmem_base_s, (uint32_t)L.tmem_cols);

    // ---------------- stage K: codes * 2^(e-6) -> bf16 chunks [kc][key]
    {
        const int kchunks = hdp >> 3;
#pragma unroll 4
        for (int t = tid; t < nkp * kchunks; t += K2T) {
            const int kc = t / nkp, j = t - kc * nkp;        // consecutive threads -> consecutive keys
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (j < Nk && kc * 8 < hd) {
                const int64_t krow = (int64_t)head * Nk + j;
                const uint2 c = __ldg(reinterpret_cast<const uint2*>(p.k_codes + krow * hd + kc * 8));
                const float w = exp2i((int)__ldg(p.k_exps + krow * NB + (kc >> 2)) - 6);
                v = dequant8_bf16(c.x, c.y, w);
            }
            *reinterpret_cast<uint4*>(sK + ((size_t)kc * nkp + j) * 16) = v;
        }
    }
    // ---------------- stage V: A1, MXINT8 along tokens (32-token windows per column), -> bf16 [tc][d]
    {
        const float* vb = p.v.p + bb * p.v.sB + hh * p.v.sH;
        for (int t = tid; t < NW * hdp; t += K2T) {
            const int w = t / hdp, d = t - w * hdp;          // consecutive threads -> consecutive columns
            uint32_t xb[32];
            uint32_t mx = 0u;
            const float* vp = vb + (int64_t)(w * 32) * p.v.sN + d;
            const int nvalid = (d < hd) ? min(32, Nk - w * 32) : 0;
#pragma unroll
            for (int tt = 0; tt < 32; ++tt) {
                uint32_t b = 0u;
                if (tt < nvalid) b = __float_as_uint(__ldg(vp));
                vp += p.v.sN;
                if (bf16) b = bf16_half_away(b);
                xb[tt] = b;
                mx = max(mx, b & 0x7fffffffu);
            }
            const int e = mx_shared_exp(mx);
            const bool dead = flush && e <= -127;
            const float wgt = exp2i(e - 6);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float f[8];
#pragma unroll
                for (int tt = 0; tt < 8; ++tt) f[tt] = (float)mx_code(xb[q * 8 + tt], e, dead) * wgt;
                *reinterpret_cast<uint4*>(sV + ((size_t)(w * 4 + q) * hdp + d) * 16) =
                    make_uint4(pack_bf16_trunc(f[0], f[1]), pack_bf16_trunc(f[2], f[3]),
                               pack_bf16_trunc(f[4], f[5]), pack_bf16_trunc(f[6], f[7]));
            }
        }
    }
    tcgen05_fence_before_sync();
    __syncthreads();
    tcgen05_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t my_tmem = tmem + ((uint32_t)(warp * 32) << 16);      // this warp's 32 lanes
    const uint32_t idesc_s = umma_idesc_bf16_f32(128, nkp);
    const uint32_t idesc_o = umma_idesc_bf16_f32(128, hdp);
    uint32_t ph_s = 0, ph_o = 0;

    for (int i0 = blockIdx.y * K2T; i0 < Nq; i0 += K2T * gridDim.y) {
        const int i = i0 + tid;
        const bool valid = i < Nq;
        const int64_t row = (int64_t)head * Nq + (valid ? i : 0);

        // ---- this row's kept-key bitmask (<= 8 words): requested at once, held in registers
        uint32_t mreg[8];
#pragma unroll
        for (int w = 0; w < 8; ++w) mreg[w] = (valid && w < NW) ? __ldg(p.mask + row * NW + w) : 0u;
        // ---- Q tile -> bf16 A operand (aliases the P buffer)
        for (int kc = 0; kc < (hdp >> 3); ++kc) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (valid && kc * 8 < hd) {
                const uint2 c = __ldg(reinterpret_cast<const uint2*>(p.q_codes + row * hd + kc * 8));
                const float w = exp2i((int)__ldg(p.q_exps + row * NB + (kc >> 2)) - 6);
                v = dequant8_bf16(c.x, c.y, w);
            }
            *reinterpret_cast<uint4*>(sP + ((size_t)kc * K2T + tid) * 16) = v;
        }
        fence_proxy_async_smem();
        tcgen05_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tcgen05_fence_after_sync();
            for (int ks = 0; ks < (hdp >> 4); ++ks) {
                const uint64_t da = umma_smem_desc(smem_u32(sP + (size_t)(2 * ks) * K2T * 16), K2T * 16, 128);
                const uint64_t db = umma_smem_desc(smem_u32(sK + (size_t)(2 * ks) * nkp * 16), nkp * 16, 128);
                umma_bf16_ss(tmem, da, db, idesc_s, ks > 0);
            }
            umma_commit(&bar_s);
        }
        mbar_wait(&bar_s, ph_s);
        ph_s ^= 1u;
        tcgen05_fence_after_sync();

        // ---- pass A: row max over the kept keys (A7: bf16 rounding of the matmul output, * scale)
        float m = -INFINITY;
        for (int w = 0; w < NW; ++w) {
            const uint32_t mw = pick8(mreg, w);
            uint32_t r[32];
            tmem_ld_32x32b_x32(my_tmem + w * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float s = __uint_as_float(r[c]);
                if (bf16) s = bf16_half_away(s);
                const float tv = __fmul_rn(s, p.scale);
                m = fmaxf(m, ((mw >> c) & 1u) ? tv : -INFINITY);
            }
        }
        if (m == -INFINITY) m = 0.f;           // row without kept keys (padding rows of the tile)
        // ---- pass B: E = exp(t - m) on kept keys, 0 elsewhere; written back over S; row sum
        float sum = 0.f;
        for (int w = 0; w < NW; ++w) {
            const uint32_t mw = pick8(mreg, w);
            uint32_t r[32];
            tmem_ld_32x32b_x32(my_tmem + w * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                float s = __uint_as_float(r[c]);
                if (bf16) s = bf16_half_away(s);
                const float ev = ((mw >> c) & 1u) ? exp_nonpos(__fsub_rn(__fmul_rn(s, p.scale), m)) : 0.f;
                sum += ev;
                r[c] = __float_as_uint(ev);
            }
            tmem_st_32x32b_x32(my_tmem + w * 32, r);
        }
        tmem_st_wait();
        const float inv = sum > 0.f ? 1.0f / sum : 0.f;

        // ---- pass C: P = E/sum -> A1 -> MXINT8 per window -> bf16 A operand; O += P_w . V_w
        for (int g0 = 0; g0 < NW; g0 += K2_PW) {
            if (g0 > 0) {                       // previous group's MMAs have finished reading sP
                mbar_wait(&bar_o, ph_o);
                ph_o ^= 1u;
            }
            const int g1 = min(g0 + K2_PW, NW);
            for (int w = g0; w < g1; ++w) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(my_tmem + w * 32, r);
                tmem_ld_wait();
                uint32_t mx = 0u;
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    uint32_t pb = __float_as_uint(__uint_as_float(r[c]) * inv);
                    if (bf16) pb = bf16_half_away(pb);
                    r[c] = pb;
                    mx = max(mx, pb);              // p >= 0: bit patterns order like the values
                }
                const int e = mx_shared_exp(mx);
                const bool dead = (flush && e <= -127) || mx == 0u;
                const float s1 = exp2i(-e), wgt = exp2i(e - 6);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float f[8];
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const float rr = __uint_as_float(r[q * 8 + t]) * s1 * 64.0f + 0.5f;
                        const int c = dead ? 0 : min(__float2int_rz(rr), 127);
                        f[t] = (float)c * wgt;
                    }
                    *reinterpret_cast<uint4*>(sP + ((size_t)((w - g0) * 4 + q) * K2T + tid) * 16) =
                        make_uint4(pack_bf16_trunc(f[0], f[1]), pack_bf16_trunc(f[2], f[3]),
                                   pack_bf16_trunc(f[4], f[5]), pack_bf16_trunc(f[6], f[7]));
                }
            }
            fence_proxy_async_smem();
            tcgen05_fence_before_sync();
            __syncthreads();
            if (tid == 0) {
                tcgen05_fence_after_sync();
                for (int w = g0; w < g1; ++w)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const uint64_t da = umma_smem_desc(
                            smem_u32(sP + (size_t)((w - g0) * 4 + 2 * h) * K2T * 16), K2T * 16, 128);
                        const uint64_t db = umma_smem_desc(
                            smem_u32(sV + (size_t)(w * 4 + 2 * h) * hdp * 16), hdp * 16, 128);
                        umma_bf16_ss(tmem, da, db, idesc_o, !(w == 0 && h == 0));
                    }
                umma_commit(&bar_o);
            }
        }
        mbar_wait(&bar_o, ph_o);
        ph_o ^= 1u;
        tcgen05_fence_after_sync();

        // ---- O -> A1 -> global (thread t writes row t)
        float* orow = p.out + bb * p.o_sB + hh * p.o_sH + (int64_t)(valid ? i : 0) * p.o_sN;
        for (int c0 = 0; c0 < hdp; c0 += 16) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(my_tmem + c0, r);
            tmem_ld_wait();
            if (valid) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (c0 + q * 4 < hd) {
                        float4 o = make_float4(__uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]),
                                               __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3]));
                        if (bf16) {
                            o.x = bf16_half_away(o.x); o.y = bf16_half_away(o.y);
                            o.z = bf16_half_away(o.z); o.w = bf16_half_away(o.w);
                        }
                        *reinterpret_cast<float4*>(orow + c0 + q * 4) = o;
                    }
                }
            }
        }
        tcgen05_fence_before_sync();
        __syncthreads();                        // every lane has read O before the next S MMA overwrites it
        tcgen05_fence_after_sync();
    }
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)L.tmem_cols);
}

}  // namespace mxp
