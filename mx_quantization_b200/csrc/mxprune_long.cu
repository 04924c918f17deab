// libmxprune, fourth translation unit: the long-sequence (Nk > 256) exact-attention kernel with two lanes per query
// row and a pipelined key-block stream (mxprune_attend_long.cuh), and its launcher.
#include <cuda_runtime.h>
#include <stdint.h>

#include "mxprune.h"
#include "mxprune_host.cuh"
#include "mxprune_device.cuh"
#include "mxprune_attend.cuh"
#include "mxprune_attend_long.cuh"

namespace mxp {

int attend_long_pair_try(const AttnParams& p, cudaStream_t st, int* rc_out) {
    const OpsLayout O = ops_layout(p.Nq, p.Nk, p.hd);
    // domain: streamed key blocks, no additive bias; path switch 0 (mxp_set_fused_path) keeps the round-1 kernel for A/B
    if (g_fused_path.load() == 0 || O.single || p.key_bias) return 1;
    const KLPSmem L = klp_smem_layout(O);
    if (O.k_blk_bytes != O.v_blk_bytes || L.total > SMEM_PER_SM - 2048) return 1;
    auto run = [&]() -> int {
        MXP_ENSURE_DYN_SMEM((k_attend_long_pair<true>), (int)(SMEM_PER_SM - 2048));
        MXP_ENSURE_DYN_SMEM((k_attend_long_pair<false>), (int)(SMEM_PER_SM - 2048));
        // one CTA per query tile (a CTA holds nothing across tiles); 256 TMEM columns per CTA: at most two CTAs per SM,
        // which the dynamic shared-memory request enforces (see launch_attend_umma)
        dim3 grid((unsigned)((size_t)p.B * p.H * O.q_tiles));
        size_t dyn = L.total;
        const size_t floor_bytes = SMEM_PER_SM / 3 + 1024;
        if (dyn < floor_bytes) dyn = floor_bytes;
        if (p.bf16) k_attend_long_pair<true><<<grid, K2P_T, dyn, st>>>(p);
        else k_attend_long_pair<false><<<grid, K2P_T, dyn, st>>>(p);
        return check_launch("k_attend_long_pair");
    };
    *rc_out = run();
    return 0;
}

}  // namespace mxp
