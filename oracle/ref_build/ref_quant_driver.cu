// oracle/_ref/libmxref.so: the REFERENCE's own C++ quantizer device functions
//   /root/reference/microxscaling/mx/cpp/{common,shared_exp,quantize}.cuh
// compiled from where they lie (never copied), driven by the loop below, which mirrors the
// structure of quantize_mx_cpp (microxscaling/mx/cpp/funcs.h:61-100): biased exponent of the
// block max -> mx_get_shared_scale -> quantize_mx_elem per element.
// TEST INFRASTRUCTURE ONLY: a second, independent oracle for stage A2.  Note the
// C++ path takes the exponent from the bit pattern, whereas the Python "golden" path evaluates
// floor(log2(.)) in fp32; they differ for block maxima a few ulps below a power of two, which is
// why ref_quantize takes the exponent policy as an argument.
#include <vector>
#include <cmath>

// The reference headers mention torch / ATen types in helper overloads we never call; these
// stand-ins let the headers parse without libtorch.
namespace torch { struct Tensor { int dim() const { return 0; } std::vector<long> sizes() const { return {}; } }; }
namespace at {
struct BFloat16 { float v; __host__ __device__ operator float() const { return v; } };
struct Half { float v; __host__ __device__ operator float() const { return v; } };
}

#include "common.cuh"
#include "shared_exp.cuh"
#include "quantize.cuh"

extern "C" {

// x: rows x hd fp32 (contiguous).  out: fake-quantised fp32 (rows x hd).  exps: rows x nb int
// (shared exponent the reference used, biased-127 removed).  block: MX block size (32).
// Returns 0.
int ref_quantize_mxint8(const float* x, long rows, int hd, int block, int flush, float* out, int* exps) {
    const int scale_bits = 8, ebits = 0, mbits = 8;
    const float max_norm = 1.984375f;   // formats.py:116-117 for int8: 127/64
    const int nb = (hd + block - 1) / block;
    for (long r = 0; r < rows; ++r) {
        for (int b = 0; b < nb; ++b) {
            const int lo = b * block, hi = (lo + block < hd) ? lo + block : hd;
            float mx = 0.f;
            for (int d = lo; d < hi; ++d) mx = fmaxf(mx, fabsf(x[r * hd + d]));
            int shared_exp = (int)get_biased_exponent(mx);
            const bool flush_tile = (shared_exp == 0 && flush);
            const float scale = mx_get_shared_scale(shared_exp, scale_bits, max_norm);
            if (exps) exps[r * nb + b] = get_biased_exponent(scale) - FLOAT32_EXP_BIAS;
            for (int d = lo; d < hi; ++d)
                out[r * hd + d] = quantize_mx_elem(x[r * hd + d], scale, flush_tile, ebits, mbits, max_norm, rd_away);
        }
    }
    return 0;
}

}  // extern "C"
