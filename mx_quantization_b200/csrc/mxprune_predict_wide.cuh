// K1-wide: the reference's OTHER top-k score sources for Nk <= 256, on the same
// staging / quantizer / tensor-core skeleton as K1-TC (mxprune_predict_tc.cuh):
//
//   pred_mode 1  partial_Q   Q = MXINT8 value c * 2^(e-6), K = +-2^e
//                            funcs/exponent_based_prediction.py:300-318, workloads/deit/scripts/main.py:111-112
//   pred_mode 2  partial_K   Q = +-2^e, K = MXINT8 value        funcs/exponent_based_prediction.py:274-298, main.py:113-114
//   pred_mode 3  exact       top-k of the TRUE scores, mx.matmul(q, k^T) * scale - the reference's
//                            `top_k and not approx_flag` branch  main.py:101-102,130
//   pred_mode 4  MXINT4      both sides re-quantized as MXINT4, c4 * 2^(e-2), |c4| <= 7 (Sanger)
//                            funcs/exponent_based_prediction.py:179-199, main.py:117-118
//   pred_mode 6  true_ex     sign * 2^floor(log2 |MX element|) per element (zero elements: +1, as the example's
//                            get_true_exponents leaves their exponent at 0): the leading one of every element  microxscaling/examples/deit/exponent_based_prediction.py:163-178 (the
//                            copy under funcs/ lacks the method), PixArt MX_transformer_block.py:663-664,811-812
//   pred_mode 7  ELSA        funcs/elsa_approximation.py:112-145, main.py:119-121: hash signs s = (MX . P^T >= 0 ? +1 : -1)
//                            of the MXINT8 rows under a d x d orthogonal matrix P, h = (d - s_q . s_k) / 2, ranked
//                            on  ||K_i|| * cos(max(pi/d * h - 0.127, 0))  - the key norm is broadcast over ROWS (:140),
//                            so inside a row the order is that of  min(s_q . s_k, cap)  (cap: the dot product of
//                            the largest h the fp32 clamp sends to angle 0, computed by the host) and an all-zero
//                            key row i ties the whole query row i.  The projections are fp32 FMAs on the CUDA cores
//                            over the exact operands the quantizer left in shared memory; the +-1 signs replace
//                            them as the MMA operands.  Needs Nq == Nk (as the reference's broadcast does), d <= 80.
//   pred_mode 5  two_step_leading_ones (EXION)   funcs/exponent_based_prediction.py:96-177, main.py:115-116
//                            value = sign(c) * e * (2^f1 + 2^f2) / 64 as the reference computes it: e is the shared
//                            exponent's VALUE, f1 the leading one of |c|, f2 the leading one of c - 2^f1 for
//                            c > 0 only (negative codes keep one term), zero codes give 0.  e * (2^f1 + 2^f2)
//                            can need more than bf16's 8 significant bits, so each side carries two exact parts
//                            A = sign * e * 2^(f1-6), B = e * 2^(f2-6) and the score is the sum of the four
//                            part products (TWO = true: operands twice as wide, one CTA per SM).
//
// Both operand kinds are exact in bf16 and come out of the same per-block quantizer
// (quantize_block_thread: r.op = c * 2^(e-6), r.pp = +-2^e), so a mode only chooses which of the two
// goes into the MMA's A and B operands.  The accumulator is the reference's fp32 matmul result bit for
// bit while one (query, key) pair's terms fit the tensor core's exact window (tools/umma_exact_test.cu:
// 22 bits; beyond it the reference's own BLAS summation order decides and parity is unpinned).
//
// These scores are not small integers (7 more bits per MXINT8 operand than the +-2^e predictor), so
// the selection works on the full 32-bit ORDERED fp32 key (written back over the accumulator once): three
// digit levels (14 + 14 + 4 bits), each re-reads the keys from TMEM, maps this row's keys to a digit
// relative to the prefix chosen so far (below prefix / inside / above) and runs K1-TC's register bisection
// (HSET2/HADD2 on fp16 bit patterns, two lanes per query row).  32 counting passes instead of ~10: these
// are the reference's comparison modes, not the headline path.  Ties: ascending key index, as everywhere.
#pragma once
#include "mxprune_predict_tc.cuh"

namespace mxp {

template <int SPLIT>
__device__ __forceinline__ void tmem_st_16x32bx2_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x32bx2.x16.b32 [%0], %1, {%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17};" ::
        "r"(taddr), "n"(SPLIT), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}

constexpr int PRED_EX = 0, PRED_PARTIAL_Q = 1, PRED_PARTIAL_K = 2, PRED_TRUE = 3, PRED_MXINT4 = 4, PRED_TWO_STEP = 5, PRED_TRUE_EX = 6, PRED_ELSA = 7;

// MXINT4 operand of one block: value = sign(x) * min(7, floor(|x| * 2^(2-e) + 0.5)) * 2^(e-2), the int4 element
// format of the reference (formats.py:86-88: mbits 4, emax 0 -> the MXINT8 shared exponent e; lshift by
// mbits - 2 = 2, round nearest, clamp to 7/4: elemwise_ops.py:155,64-65,163-164).  xv holds the block after
// A1; <= 3 significant bits, so the bf16 truncation is exact.
__device__ __forceinline__ void int4_operand(const uint32_t (&xv)[32], int e, bool dead, uint4 (&op4)[4]) {
    const bool tiny = e < -100;                                     // keep 2^(2-e) finite: pre-scale by 2^60
    const float up = tiny ? 1152921504606846976.0f : 1.0f;
    const float s1 = dead ? 0.f : exp2i(2 - e - (tiny ? 60 : 0));
    const float wgt = exp2i(e - 2);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t w[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float f[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const uint32_t x = xv[8 * c + 2 * h + t];
                const float a = fabsf(__uint_as_float(x)) * up;
                const float m = fminf(floorf(fmaf(a, s1, 0.5f)), 7.0f) * wgt;
                f[t] = __uint_as_float(__float_as_uint(m) | (x & 0x80000000u));
            }
            w[h] = pack_bf16_trunc(f[0], f[1]);
        }
        op4[c] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// true_ex operand from the exact bf16 operand c * 2^(e-6): its exponent field IS floor(log2 |value|), so the leading
// one is the operand with the mantissa cleared; a zero element (either sign: `MX < 0` is false for -0) is +1.0:
// get_true_exponents (:98-110) leaves the exponent of zeros at 0.  nd: elements of the block that exist.
__device__ __forceinline__ void true_ex_operand(const uint4 (&op)[4], int nd, uint4 (&out)[4]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t w[4] = {op[c].x, op[c].y, op[c].z, op[c].w};
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            const uint32_t m = w[h] & 0xff80ff80u;
            const uint32_t t = m & 0x7f807f80u;
            const uint32_t nz = ((t + 0x7f807f80u) | t) & 0x80008000u;        // bit 15 of a half: magnitude != 0
            const uint32_t z = ~nz & 0x80008000u;                            // zero halves
            w[h] = (m & ~z) | ((z >> 8) * 0x7fu);                           // -0 loses its sign; zero -> 0x3f80 = 1.0
        }
        out[c] = 8 * c < nd ? make_uint4(w[0], w[1], w[2], w[3]) : make_uint4(0u, 0u, 0u, 0u);
    }
}

// two_step_leading_ones parts of one block from its int8 codes (cw: element 8c + 4h + t in byte t of word 2c + h)
// and the predictor exponent ep (funcs/exponent_based_prediction.py:104-127):
//   sign = torch.sign(MX) (0 for a zero code); f1 = floor(log2 |c|); temp = max(c - 2^f1, 0) on the SIGNED code,
//   so only positive codes get a second term f2 = floor(log2 temp); value = sign * ep * (2^f1 + 2^f2) / 64.
__device__ __forceinline__ void two_step_operands(const uint32_t (&cw)[8], int ep, uint4 (&pa)[4], uint4 (&pb)[4]) {
    const float fe = (float)ep;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t wa[4], wb[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float fa[4], fb[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int code = (int)(signed char)((cw[2 * c + h] >> (8 * t)) & 0xffu);
                const int a = abs(code);
                const int f1 = 31 - __clz(a | 1);
                const int rest = code > 0 ? a - (1 << f1) : 0;
                const int f2 = 31 - __clz(rest | 1);
                fa[t] = code == 0 ? 0.f : (code < 0 ? -fe : fe) * exp2i(f1 - 6);
                fb[t] = rest == 0 ? 0.f : fe * exp2i(f2 - 6);
            }
            wa[2 * h] = pack_bf16_trunc(fa[0], fa[1]); wa[2 * h + 1] = pack_bf16_trunc(fa[2], fa[3]);
            wb[2 * h] = pack_bf16_trunc(fb[0], fb[1]); wb[2 * h + 1] = pack_bf16_trunc(fb[2], fb[3]);
        }
        pa[c] = make_uint4(wa[0], wa[1], wa[2], wa[3]);
        pb[c] = make_uint4(wb[0], wb[1], wb[2], wb[3]);
    }
}

constexpr int ELSA_MAX_HD = 80;

template <int NC, bool TWO, bool ELSA = false>
__global__ void __launch_bounds__(K1C_T, (TWO || ELSA) ? 1 : 2)
k_predict_topk_wide(const PredParams p, const __grid_constant__ K1cMaps maps, const int ring, const int G) {
    extern __shared__ __align__(1024) unsigned char smem_k1w[];
    unsigned char* const smem = smem_k1w;
    constexpr int NMMA = 32 * NC;
    constexpr int NCH = (NC + 1) / 2;                               // key chunks per thread
    const int Nk = p.Nk, Nq = p.Nq, hd = p.hd, kk = p.top_k;
    const int mode = p.pred_mode;
    const bool q_exact = mode == PRED_PARTIAL_Q || mode == PRED_TRUE || ELSA;
    const bool k_exact = mode == PRED_PARTIAL_K || mode == PRED_TRUE || ELSA;
    const bool true_mode = mode == PRED_TRUE;
    const bool int4_mode = mode == PRED_MXINT4 || mode == PRED_TRUE_EX;       // modes whose operand is formed in op4
    const bool true_ex_mode = mode == PRED_TRUE_EX;
    const K1cSmem L = k1c_smem_layout(hd, NC, ring, G, false, TWO ? 2 : 1);
    const int nfull = L.nfull, tail = L.tail, nb = L.nb;
    const int kch = L.hdp >> 3;
    const int tail_chunks_hbm = (((hd + 15) & ~15) >> 3) - 4 * nfull;
    const int tail_chunks = kch - 4 * nfull;
    unsigned char* s_kop = smem + L.off_kop;
    unsigned char* s_qop = smem + L.off_qop;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.off_misc + 32);
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + L.off_misc + 64);   // [K1C_MAXR]
    uint64_t* bar_mma = bar_full + K1C_MAXR;
    // ELSA: the projection matrix (fp32, hd x hd) behind the regular layout; key-row "all zero" flags in the
    // (otherwise unused) key-exponent bytes
    float* s_proj = reinterpret_cast<float*>(smem + ((L.total + 15) & ~(size_t)15));
    unsigned char* s_kzero = smem + L.off_kexp;

    const int head = blockIdx.x, bb = head / p.H, hh = head - bb * p.H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int lane_base = 32 * (warp & 3) + 16 * (warp >> 2);
    const int rr = lane_base + (lane & 15);
    const int part = lane >> 4;
    const bool bf16 = p.bf16, flush = p.flush;
    const bool write_kop = p.k_op != nullptr && blockIdx.y == 0;
    const OpsLayout OL = ops_layout(Nq, Nk, hd);
    unsigned char* k_op = p.k_op ? p.k_op + (size_t)head * OL.k_head_bytes : nullptr;
    unsigned char* q_op = p.q_op ? p.q_op + (size_t)head * OL.q_head_bytes : nullptr;
    const int kb_rows = OL.kb_rows;
    const float sscale = p.score_scale;
    // additive key bias (PixArt cross-attention text mask): added in fp32 to the ranked value of every mode, as
    // the reference does (MX_transformer_block.py:803 true_scores += attn_bias, :822 pred_scores + attn_bias)
    const float* kbias = p.key_bias ? p.key_bias + bb * p.kb_sB : nullptr;

    const int CR = K1C_ROWS * G, cr_shift = G == 2 ? 7 : 6;
    const int nks = (NMMA + CR - 1) / CR;
    const int qsteps = K1C_TILE / CR;
    const int tiles = (Nq + K1C_TILE - 1) / K1C_TILE;
    const int my_tiles = (tiles - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;
    const int nsteps = nks + qsteps * my_tiles;
    const uint32_t box_main = (uint32_t)L.box_main, box_tail = (uint32_t)L.box_tail;
    const uint32_t slot_tx = (uint32_t)G * (box_main + box_tail);
    const uint32_t slot_bytes = (uint32_t)L.slot_bytes;
    const uint32_t tail_base = (uint32_t)G * box_main;

    auto issue = [&](int c, int slot_i) {                           // one thread
        unsigned char* slot = smem + (size_t)slot_i * slot_bytes;
        uint64_t* bar = &bar_full[slot_i];
        const bool is_k = c < nks;
        int row0;
        if (is_k) row0 = c * CR;
        else {
            const int qc = c - nks;
            row0 = ((int)blockIdx.y + (qc / qsteps) * (int)gridDim.y) * K1C_TILE + (qc % qsteps) * CR;
        }
        mbar_expect_tx(bar, slot_tx);
        for (int g = 0; g < G; ++g) {
            if (nfull) tma_load_5d(slot + g * box_main, is_k ? &maps.k_main : &maps.q_main, 0, row0 + g * K1C_ROWS, 0, hh, bb, bar);
            if (tail) tma_load_4d(slot + tail_base + g * box_tail, is_k ? &maps.k_tail : &maps.q_tail, 0, row0 + g * K1C_ROWS, hh, bb, bar);
        }
    };

    if (tid == 0) {
        for (int r = 0; r < ring; ++r) mbar_init(&bar_full[r], 1);
        mbar_init(bar_mma, 1);
        prefetch_tmap(&maps.k_main); prefetch_tmap(&maps.q_main);
        if (tail) { prefetch_tmap(&maps.k_tail); prefetch_tmap(&maps.q_tail); }
    }
    if (warp == 0) tmem_alloc(s_tmem, (uint32_t)L.tmem_cols);
    if (ELSA) {
        for (int t = tid; t < hd * hd; t += K1C_T) s_proj[t] = __ldg(p.elsa_proj + t);
        s_kzero[tid] = 1;                                           // K1C_T == 256 == the most key rows
    }
    tcgen05_fence_before_sync();
    __syncthreads();
    tcgen05_fence_after_sync();
    const uint32_t tmem = *s_tmem;
    const uint32_t my_tmem = tmem + ((uint32_t)lane_base << 16);
    if (tid == 0) {
        const int pre = min(ring, nsteps);
        for (int c = 0; c < pre; ++c) issue(c, c);
    }
    const uint32_t idesc = umma_idesc_bf16_f32(128, NMMA);
    const int NW = (Nk + 31) >> 5;
    uint32_t ph_mma = 0;
    int slot_i = 0;
    uint32_t slot_par = 0;
    int tile = (int)blockIdx.y - (int)gridDim.y;
    int qstep = qsteps - 1;

    for (int c = 0; c < nsteps; ++c) {
        const bool is_k = c < nks;
        if (!is_k) {
            if (++qstep == qsteps) { qstep = 0; tile += (int)gridDim.y; }
        }
        const int row0 = is_k ? c * CR : tile * K1C_TILE + qstep * CR;
        const int nrows = is_k ? Nk : Nq;
        const unsigned char* slot = smem + (size_t)slot_i * slot_bytes;
        mbar_wait(&bar_full[slot_i], slot_par);

        // -------- quantize the step's rows (one thread per MX block), as K1-TC; the MMA operand of a
        // side is the exact value or the +-2^e predictor value according to the mode
        const int ntask = CR * nb;
        for (int t = tid; t < ntask; t += K1C_T) {
            const int b = t >> cr_shift, rl = t & (CR - 1);
            const int g = rl >> 6, rl6 = rl & 63;
            const int row = row0 + rl;
            const bool in_range = row < nrows;
            const bool full = b < nfull;
            uint32_t xv[32];
            if (full) {
                const unsigned char* src = slot + g * box_main + (b * 64 + rl6) * 128;
                const int sw7 = (rl6 & 7) << 4;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    const uint4 v = *reinterpret_cast<const uint4*>(src + ((s << 4) ^ sw7));
                    xv[4 * s] = v.x; xv[4 * s + 1] = v.y; xv[4 * s + 2] = v.z; xv[4 * s + 3] = v.w;
                }
            } else {
                const unsigned char* src = slot + tail_base + g * box_tail + rl6 * tail * 4;
#pragma unroll
                for (int s = 0; s < 8; ++s) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (4 * s < tail) v = *reinterpret_cast<const uint4*>(src + (s << 4));
                    xv[4 * s] = v.x; xv[4 * s + 1] = v.y; xv[4 * s + 2] = v.z; xv[4 * s + 3] = v.w;
                }
            }
            BlockQ r;
            quantize_block_thread<TWO>(xv, full ? 32 : tail, bf16, flush, r);        // xv: A1 applied in place
            uint4 op4[4], opb[4];                                   // first / second operand part of the mode
            if (true_ex_mode) true_ex_operand(r.op, full ? 32 : tail, op4);
            else if (int4_mode) int4_operand(xv, r.e, flush && r.e <= -127, op4);
            if (TWO) two_step_operands(r.cw, r.ep, op4, opb);
            const int nchunk = full ? 4 : tail_chunks, nchunk_hbm = full ? 4 : tail_chunks_hbm;
            const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
            if (is_k) {
                if (row < NMMA) {
                    unsigned char* dst = s_kop + ((4 * b) * NMMA + row) * 16;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch)
                        if (ch < nchunk)
                            *reinterpret_cast<uint4*>(dst + ch * (NMMA * 16)) =
                                !in_range ? zero4 : (int4_mode || TWO) ? op4[ch] : k_exact ? r.op[ch] : r.pp[ch];
                    if (TWO) {
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch)
                            if (ch < nchunk)
                                *reinterpret_cast<uint4*>(dst + (kch + ch) * (NMMA * 16)) = in_range ? opb[ch] : zero4;
                    }
                }
                if (write_kop && row < kb_rows) {
                    unsigned char* dst = k_op + ((size_t)(4 * b) * kb_rows + row) * 16;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch)
                        if (ch < nchunk_hbm) *reinterpret_cast<uint4*>(dst + (size_t)ch * kb_rows * 16) = r.op[ch];
                }
            } else {
                const int rt = qstep * CR + rl;
                unsigned char* dst = s_qop + ((4 * b) * K1C_TILE + rt) * 16;
#pragma unroll
                for (int ch = 0; ch < 4; ++ch)
                    if (ch < nchunk) *reinterpret_cast<uint4*>(dst + ch * (K1C_TILE * 16)) =
                        (int4_mode || TWO) ? op4[ch] : q_exact ? r.op[ch] : r.pp[ch];
                if (TWO) {
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch)
                        if (ch < nchunk) *reinterpret_cast<uint4*>(dst + (kch + ch) * (K1C_TILE * 16)) = opb[ch];
                }
                if (q_op) {
                    unsigned char* gdst = q_op + (size_t)tile * OL.q_tile_bytes + ((size_t)(4 * b) * K1C_TILE + rt) * 16;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch)
                        if (ch < nchunk_hbm) *reinterpret_cast<uint4*>(gdst + ch * (K1C_TILE * 16)) = r.op[ch];
                }
            }
        }
        if (ELSA) {
            // ---- hash signs of the step's rows: x = the row's exact operand (bf16 in shared memory), hash j =
            // (sum_d x[d] * P[j][d] >= 0); thread <-> (row, slice of the d hashes), a warp shares its slice of P
            __syncthreads();                                        // the step's operand rows are complete
            const int rl = tid & (CR - 1), jpart = tid >> cr_shift, nparts = K1C_T >> cr_shift;
            const int JT = hd / nparts, j0 = jpart * JT;            // hd % 4 == 0, nparts in {2, 4}
            const int ROWS = is_k ? NMMA : K1C_TILE;
            const int rindex = is_k ? row0 + rl : qstep * CR + rl;
            const bool present = !is_k || rindex < NMMA;
            unsigned char* base = (is_k ? s_kop : s_qop) + (size_t)rindex * 16;
            float xf[ELSA_MAX_HD];
            uint32_t any = 0u;
#pragma unroll
            for (int ch = 0; ch < ELSA_MAX_HD / 8; ++ch) {
                uint4 v = make_uint4(0u, 0u, 0u, 0u);
                if (present && ch < kch) v = *reinterpret_cast<const uint4*>(base + (size_t)ch * ROWS * 16);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    xf[8 * ch + 2 * h] = __uint_as_float(w[h] << 16);
                    xf[8 * ch + 2 * h + 1] = __uint_as_float(w[h] & 0xffff0000u);
                    any |= w[h] & 0x7fff7fffu;
                }
            }
            uint64_t bits = 0ull;                                   // bit jj: hash j0 + jj is negative
            for (int jj = 0; jj < JT; ++jj) {
                const float* pr = s_proj + (size_t)(j0 + jj) * hd;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int d = 0; d < ELSA_MAX_HD; d += 4) {
                    if (d < hd) {
                        const float4 pv = *reinterpret_cast<const float4*>(pr + d);
                        a0 = fmaf(xf[d], pv.x, a0); a1 = fmaf(xf[d + 1], pv.y, a1);
                        a2 = fmaf(xf[d + 2], pv.z, a2); a3 = fmaf(xf[d + 3], pv.w, a3);
                    }
                }
                const float acc = (a0 + a1) + (a2 + a3);
                bits |= (uint64_t)(acc >= 0.f ? 0u : 1u) << jj;
            }
            __syncthreads();                                        // every thread has read its row before any overwrite
            if (present) {
                for (int jj = 0; jj < JT; jj += 2) {                // JT and j0 are even: two hashes per 32-bit store
                    const int j = j0 + jj;
                    const uint32_t two = (((bits >> jj) & 1ull) ? 0xBF80u : 0x3F80u) |
                                         ((((bits >> (jj + 1)) & 1ull) ? 0xBF80u : 0x3F80u) << 16);
                    *reinterpret_cast<uint32_t*>(base + (size_t)(j >> 3) * ROWS * 16 + (j & 7) * 2) = two;
                }
                if (is_k && jpart == 0 && rindex < 256) s_kzero[rindex] = any == 0u ? 1 : 0;
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0 && c + ring < nsteps) issue(c + ring, slot_i);
        if (++slot_i == ring) { slot_i = 0; slot_par ^= 1u; }
        if (is_k || qstep != qsteps - 1) continue;

        // =============== a full query tile is quantized: score, select, emit
        const int i = tile * K1C_TILE + rr;
        const bool valid = i < Nq;
        const int64_t row = (int64_t)head * Nq + (valid ? i : 0);
        if (kk >= Nk) {                                             // every key kept
            if (valid) {
                for (int w = part; w < NW; w += 2) {
                    const int nv = Nk - 32 * w;
                    p.mask[row * NW + w] = nv >= 32 ? 0xffffffffu : (1u << nv) - 1u;
                }
                if (p.idx && part == 0)
                    for (int j = 0; j < Nk; ++j) p.idx[row * kk + j] = j;
            }
            continue;
        }
        if (tid == 0) {
            tcgen05_fence_after_sync();
            // TWO: (A_q + B_q) . (A_k + B_k) as the four part products into one accumulator
            for (int pq = 0; pq < (TWO ? 2 : 1); ++pq)
                for (int pk = 0; pk < (TWO ? 2 : 1); ++pk)
                    for (int ks = 0; ks < (L.hdp >> 4); ++ks) {
                        const uint64_t da = umma_smem_desc(smem_u32(s_qop + (size_t)(pq * kch + 2 * ks) * K1C_TILE * 16), K1C_TILE * 16, 128);
                        const uint64_t db = umma_smem_desc(smem_u32(s_kop + (size_t)(pk * kch + 2 * ks) * NMMA * 16), NMMA * 16, 128);
                        umma_bf16_ss(tmem, da, db, idesc, (pq | pk | ks) != 0);
                    }
            umma_commit(bar_mma);
        }
        mbar_wait(bar_mma, ph_mma);
        ph_mma ^= 1u;
        tcgen05_fence_after_sync();
        if (!__any_sync(FULL, valid)) continue;                     // a warp of rows past Nq has nothing to select

        const int col0 = part * NCH * 32;                           // this thread's first key column

        // ---- convert once: the SIGNED ordered key of the ranked value (int order == float order), written
        // back over the accumulator.  The exact mode ranks A1(s) * scale as the reference does
        // (matmul.py:89-91 then main.py:102); + 0.0f folds -0 into +0; columns past Nk sort below everything.
#pragma unroll
        for (int w = 0; w < NCH; ++w) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t r[16];
                tmem_ld_16x32bx2_x16<NCH * 32>(my_tmem + w * 32 + h * 16, r);
                tmem_ld_wait();
                const int nv = Nk - (col0 + w * 32 + h * 16);       // valid columns among these 16
                if (true_mode) {                                    // warp-uniform branches: the common modes skip them
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        float sc = __uint_as_float(r[t]);
                        if (bf16) sc = bf16_half_away(sc);
                        r[t] = __float_as_uint(__fmul_rn(sc, sscale));
                    }
                }
                if (ELSA) {
                    const bool zrow = i < 256 && s_kzero[valid ? i : 0] != 0;      // ||K_i|| = 0 scales query row i to 0
#pragma unroll
                    for (int t = 0; t < 16; ++t)
                        r[t] = zrow ? 0u : __float_as_uint(fminf(__uint_as_float(r[t]), p.elsa_cap));
                }
                if (kbias != nullptr) {
#pragma unroll
                    for (int t = 0; t < 16; ++t)
                        if (t < nv)
                            r[t] = __float_as_uint(__fadd_rn(__uint_as_float(r[t]), __ldg(kbias + col0 + w * 32 + h * 16 + t)));
                }
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const int b = __float_as_int(__uint_as_float(r[t]) + 0.0f);
                    const int o = b ^ ((b >> 31) & 0x7fffffff);
                    r[t] = t < nv ? (uint32_t)o : 0x80000000u;
                }
                tmem_st_16x32bx2_x16<NCH * 32>(my_tmem + w * 32 + h * 16, r);
            }
        }
        tmem_st_wait();

        // ---- select: three digit levels of the 32-bit key, most significant first (14 + 14 + 4 bits).  A level
        // maps this row's keys to fp16 bit patterns - 0x401 + digit inside the prefix chosen so far, anything
        // <= 0x400 below it, a larger pattern above it - and bisects the digit with K1-TC's register counting.
        uint32_t kw[NCH * 16];
        int base28 = 0;                                             // chosen prefix, in the (key >> 4) domain
        uint32_t Tpat = 0u;
#pragma unroll 1
        for (int level = 0; level < 3; ++level) {
            // one branch-free mapping for the three levels:
            //   y = clamp((key >> sh) - base + addc, 0, maxc);  pattern = y * mul + ((key & lowm) | orc)
            //   level 0  digit = (key >> 18) + 8192                      -> 0x401 + digit
            //   level 1  digit = (key >> 4) - prefix (28-bit operands)   -> 0x401 + digit, <= 0x400 below, 0x4401 above
            //   level 2  y = 0 below / 1 inside / 2 above the 28-bit prefix -> 0x400 + 32 y + (key & 15)
            const int sh = level == 0 ? 18 : 4;
            const int base = level == 0 ? -8192 : base28;
            const int addc = level == 2 ? 1 : 1 + (int)K1_KEY_BIAS;
            const int maxc = level == 2 ? 2 : 16384 + 1 + (int)K1_KEY_BIAS;
            const uint32_t mul = level == 2 ? 32u : 1u;
            const uint32_t lowm = level == 2 ? 15u : 0u, orc = level == 2 ? K1_KEY_BIAS : 0u;
#pragma unroll
            for (int w = 0; w < NCH; ++w) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t r[16];
                    tmem_ld_16x32bx2_x16<NCH * 32>(my_tmem + w * 32 + h * 16, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int t = 0; t < 16; ++t) {
                        const int y = __viaddmin_s32_relu(((int)r[t] >> sh) - base, addc, maxc);
                        const uint32_t f = (uint32_t)y * mul + ((r[t] & lowm) | orc);
                        if (h == 0) kw[16 * w + t] = f;
                        else kw[16 * w + t] = __byte_perm(kw[16 * w + t], f, 0x5410);
                    }
                }
            }
            const int D = level == 2 ? 4 : 14;
            const uint32_t pat0 = level == 2 ? K1_KEY_BIAS + 32u : K1_KEY_BIAS + 1u;        // pattern of digit 0
            uint32_t tsel = 0u;
#pragma unroll 1
            for (int bit = D - 1; bit >= 0; --bit) {
                const uint32_t cand = tsel | (1u << bit);
                const int mine = count_ge_regs<NCH * 16>(kw, pat0 + cand);
                const int theirs = __shfl_xor_sync(FULL, mine, 16);
                if (mine + theirs >= kk) tsel = cand;
            }
            if (level == 0) base28 = ((int)tsel - 8192) << 14;
            else if (level == 1) base28 += (int)tsel;
            else Tpat = pat0 + tsel;
        }
        // the tile's scores are consumed: TMEM and the Q-side shared memory may be reused
        tcgen05_fence_before_sync();

        // ---- emit from the last level's patterns: > Tpat kept, == Tpat kept in ascending key index until top_k
        uint32_t gtw[NCH], eqw[NCH];
        int ngt_m = 0, neq_m = 0;
        {
            const __half2 t2 = u32_as_h2(Tpat * 0x00010001u);
#pragma unroll
            for (int w = 0; w < NCH; ++w) {
                uint32_t gt = 0u, eq = 0u;
#pragma unroll
                for (int t = 0; t < 16; ++t) {
                    const __half2 kv = u32_as_h2(kw[16 * w + t]);
                    gt |= __hgt2_mask(kv, t2) & (0x00010001u << t);
                    eq |= __heq2_mask(kv, t2) & (0x00010001u << t);
                }
                const int nv = Nk - (col0 + w * 32);
                const uint32_t vm = nv >= 32 ? 0xffffffffu : (nv <= 0 ? 0u : (1u << nv) - 1u);
                gtw[w] = gt & vm;
                eqw[w] = eq & vm;
                ngt_m += __popc(gtw[w]);
                neq_m += __popc(eqw[w]);
            }
        }
        {
            const int ngt_o = __shfl_xor_sync(FULL, ngt_m, 16);
            const int neq_o = __shfl_xor_sync(FULL, neq_m, 16);
            const int rem_all = kk - (ngt_m + ngt_o);               // ties to keep in the whole row
            const int ties0 = part == 0 ? neq_m : neq_o;
            const int ngt0 = part == 0 ? ngt_m : ngt_o;
            int rem = part == 0 ? rem_all : rem_all - min(rem_all, ties0);
            int pos = part == 0 ? 0 : ngt0 + min(rem_all, ties0);
#pragma unroll
            for (int w = 0; w < NCH; ++w) {
                const int gw = part * NCH + w;
                const int cnt = __popc(eqw[w]);
                uint32_t take = eqw[w];
                if (cnt > rem) take = keep_lowest_bits_fast(eqw[w], rem);
                rem -= min(cnt, rem);
                const uint32_t word = gtw[w] | take;
                if (valid && gw < NW) {
                    p.mask[row * NW + gw] = word;
                    if (p.idx) {
                        uint32_t w2 = word;
                        while (w2) {
                            const int bpos = __ffs(w2) - 1;
                            w2 &= w2 - 1u;
                            p.idx[row * kk + pos++] = gw * 32 + bpos;
                        }
                    }
                }
            }
        }
    }
    tcgen05_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)L.tmem_cols);
}

}  // namespace mxp
