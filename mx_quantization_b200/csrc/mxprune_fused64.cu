// libmxprune, third translation unit: the head_dim 64 instantiations of the fused kernel (DeiT / ViT heads), whose
// staging geometry and operand offsets are compile-time constants.  Its own unit so that it compiles in parallel.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mxprune.h"
#include "mxprune_host.cuh"
#include "mxprune_device.cuh"
#include "mxprune_attend.cuh"
#include "mxprune_attend_sparse.cuh"
#include "mxprune_fused.cuh"
#include "mxprune_fused_launch.cuh"

namespace mxp {
template int launch_fused_hd<8, 0, 64>(const FusedParams&, const FusedMaps&, int, cudaStream_t);
template int launch_fused_hd<7, 13, 64>(const FusedParams&, const FusedMaps&, int, cudaStream_t);
template int launch_fused_hd<7, 14, 64>(const FusedParams&, const FusedMaps&, int, cudaStream_t);
template int launch_fused_hd<7, 0, 64>(const FusedParams&, const FusedMaps&, int, cudaStream_t);
}  // namespace mxp
