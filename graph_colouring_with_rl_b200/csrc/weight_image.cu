// weight_image.cu -- the network's MMA weight windows as pre-converted bf16 hi | lo operand tiles (gn_common.cuh).
// Built by one launch at the start of gcrl_gn_forward / gcrl_gn_backward in the tensor-core math modes (52 tiles, 832 KB,
// in the caller's workspace): rebuilding per pass costs ~4 us and can never be stale, which a cached image keyed on the
// parameter pointer could be (an optimizer step rewrites the parameters in place).
#include "gn_common.cuh"
#include "tc.cuh"

namespace gcrl {

struct WimgSpec { const float* W; int ld, col0; };
struct WimgArgs { WimgSpec t[WIMG_TILES]; };

__global__ void __launch_bounds__(256) k_build_weight_image(const __grid_constant__ WimgArgs a, unsigned char* __restrict__ img) {
    pdl_prologue();
    const WimgSpec s = a.t[blockIdx.x];
    if (!s.W) return;                      // a window this network shape does not have (block 1: 2-wide vertex features)
    unsigned char* tile = img + (size_t)blockIdx.x * WIMG_TILE_BYTES;
    tc::load_weight_tile(s.W, s.ld, s.col0, reinterpret_cast<__nv_bfloat16*>(tile), reinterpret_cast<__nv_bfloat16*>(tile + 8192),
                         (int)threadIdx.x, 256);
}

int launch_build_weight_image(const NetW& net, unsigned char* img, cudaStream_t st) {
    WimgArgs a;
    for (int t = 0; t < WIMG_TILES; ++t) a.t[t] = WimgSpec{nullptr, 0, 0};
    for (int l = 0; l < GCRL_N_BLOCKS; ++l) {
        const BlockW& b = net.blk[l];
        WimgSpec* t = &a.t[l * WIMG_PER_BLOCK];
        const int ldW1 = 2 * b.Dv + b.De, ldU1 = 256 + b.Dv;
        t[WT_W2] = WimgSpec{b.W2, 64, 0};
        if (b.De == 64) t[WT_WC] = WimgSpec{b.W1, ldW1, 2 * b.Dv};
        for (int c = 0; c < 5; ++c)
            if (64 * (c + 1) <= ldU1) t[WT_U1 + c] = WimgSpec{b.U1, ldU1, 64 * c};
        t[WT_U2] = WimgSpec{b.U2, 64, 0};
        if (b.Dv == 64) {
            t[WT_WA] = WimgSpec{b.W1, ldW1, 0};
            t[WT_WB] = WimgSpec{b.W1, ldW1, 64};
        }
    }
    if (net.head.Dg == 0) {
        a.t[WIMG_F1] = WimgSpec{net.head.F1, 64, 0};
        a.t[WIMG_F2] = WimgSpec{net.head.F2, 64, 0};
    }
    launch_k(k_build_weight_image, dim3(WIMG_TILES), dim3(256), 0, st, a, img);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // namespace gcrl
