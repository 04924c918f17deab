// oracle/_ref/libmxref_cuda.so: the REFERENCE's own CUDA quantizer kernels
//   /root/reference/microxscaling/mx/cpp/mx.cuh:57-91   quantize_mx_innermost_cuda_kernel  (head_dim 64: tile 32 divides the axis)
//   /root/reference/microxscaling/mx/cpp/mx.cuh:98-158  quantize_mx_by_tile_cuda_kernel    (head_dim 72: 72 % 32 != 0)
// compiled for sm_100a from where they lie (never copied) and launched with the grid the reference's host code
// uses (mx.cu:156-180, common.cuh:92-118: one thread per element / per tile, blocks of 1024).  SURVEY.md 2a names
// these two kernels as the existing code a B200 quantizer has to beat; tools/bench_quant.py times them beside
// mxp_quantize_mxint8 on the same box.  TEST / MEASUREMENT INFRASTRUCTURE ONLY - nothing in the product loads it.
#include <vector>
#include <cmath>

// The reference headers mention torch / ATen types in helper overloads we never call; these stand-ins let the
// headers parse without libtorch (same as ref_quant_driver.cu).
namespace torch { struct Tensor { int dim() const { return 0; } std::vector<long> sizes() const { return {}; } }; }
namespace at {
struct BFloat16 { float v; __host__ __device__ operator float() const { return v; } };
struct Half { float v; __host__ __device__ operator float() const { return v; } };
}

#include "mx.cuh"

extern "C" {

// MXINT8 (ebits 0, mbits 8, max_norm 127/64, scale_bits 8, round half away) along the innermost axis.
// in / out: device pointers to `total` contiguous fp32; tile = 32.  Returns the CUDA error code of the launch.
int ref_cuda_quantize_innermost(const float* in, long total, int tile, int flush, float* out, void* stream) {
    const long blocks = get_blocks(total);
    const int threads = get_threads(total);
    quantize_mx_innermost_cuda_kernel<float><<<blocks, threads, 0, (cudaStream_t)stream>>>(
        in, 8, 0, 8, 1.984375f, total, tile, flush != 0, rd_away, out);
    return (int)cudaGetLastError();
}

// Any axis / tile (the path a head_dim of 72 takes): in viewed as (pre, axis_size, post), tiles of `tile` along the axis.
int ref_cuda_quantize_by_tile(const float* in, long pre, int axis_size, long post, int tile, int flush, float* out,
                              void* stream) {
    int num_tiles = axis_size / tile + (axis_size % tile ? 1 : 0);
    const long total_tiles = pre * num_tiles * post;
    const long blocks = get_blocks(total_tiles);
    const int threads = get_threads(total_tiles);
    quantize_mx_by_tile_cuda_kernel<float><<<blocks, threads, 0, (cudaStream_t)stream>>>(
        in, 8, 0, 8, 1.984375f, (int)total_tiles, tile, num_tiles, axis_size, (int)post, flush != 0, rd_away, out);
    return (int)cudaGetLastError();
}

}  // extern "C"
