// gn_common.cuh -- pieces shared by the GN forward and backward kernels.
#pragma once
#include "segreduce.cuh"

namespace gcrl {

constexpr int TILE = 128;         // edge rows per tile
constexpr int LDT = 68;           // padded smem row stride (floats), keeps 16 B alignment
constexpr int ITEM_EDGES = 2048;  // target edges per work item

// acc[8][4] += A[8rg+i][:] . W[:][4cg+j], K = 64  (A: smem tile, W: smem [k][n])
__device__ __forceinline__ void gemm64(const float (*A)[LDT], const float (*W)[64], int rg, int cg,
                                       float acc[8][4]) {
#pragma unroll 2
    for (int k4 = 0; k4 < 16; ++k4) {
        float4 a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = *reinterpret_cast<const float4*>(&A[8 * rg + i][4 * k4]);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const float4 w = *reinterpret_cast<const float4*>(&W[4 * k4 + kk][4 * cg]);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float av = kk == 0 ? a[i].x : kk == 1 ? a[i].y : kk == 2 ? a[i].z : a[i].w;
                acc[i][0] = fmaf(av, w.x, acc[i][0]);
                acc[i][1] = fmaf(av, w.y, acc[i][1]);
                acc[i][2] = fmaf(av, w.z, acc[i][2]);
                acc[i][3] = fmaf(av, w.w, acc[i][3]);
            }
        }
    }
}

// ---- weight image (weight_image.cu) --------------------------------------------------------------------------------------
// Every 64 x 64 weight window the tensor-core kernels use as an MMA operand, pre-converted ONCE per network pass into the
// bf16 hi | lo SWIZZLE_128B tile format (8 KB + 8 KB) in the caller's workspace, so that a kernel's prologue is a flat copy
// (one L2 round trip) instead of strided fp32 loads + conversion per tile row: k_vertex_fwd_tc spent half of its ~23 us on
// eight such tiles when the batch is small.  Tile t of block l: l * WIMG_PER_BLOCK + kind; the head's fc1 / fc2 follow.
enum { WT_W2 = 0, WT_WC = 1, WT_U1 = 2 /* + K chunk c < 5 */, WT_U2 = 7, WT_WA = 8, WT_WB = 9, WIMG_PER_BLOCK = 10 };
constexpr int WIMG_F1 = GCRL_N_BLOCKS * WIMG_PER_BLOCK, WIMG_F2 = WIMG_F1 + 1, WIMG_TILES = WIMG_F1 + 2;
constexpr size_t WIMG_TILE_BYTES = 16384;
inline size_t wimg_bytes() { return (size_t)WIMG_TILES * WIMG_TILE_BYTES; }
inline const unsigned char* wimg_tile(const unsigned char* img, int l, int kind) {
    return img ? img + (size_t)(l * WIMG_PER_BLOCK + kind) * WIMG_TILE_BYTES : nullptr;
}
int launch_build_weight_image(const NetW& net, unsigned char* img, cudaStream_t st);

// Activations the backward pass needs (one carve of the caller's stash buffer).
struct Stash {
    float* e[5];                    // e_1..e_5 [E,64]  (e_5 only in the tensor-core modes: its backward then runs as a
                                    // regular pipelined block instead of recomputing it; the network itself discards e_5)
    float* h[GCRL_N_BLOCKS];        // h_1..h_5 [N,64]
    float* Pi[GCRL_N_BLOCKS];       // A_l h_{l-1} + b1_l [N,64]
    float* Pj[GCRL_N_BLOCKS];       // B_l h_{l-1}        [N,64]
    float* agg[GCRL_N_BLOCKS];      // [N,256]
    float* var[GCRL_N_BLOCKS];      // [N,64]
    int32_t* amin[GCRL_N_BLOCKS];   // [N,64]
    int32_t* amax[GCRL_N_BLOCKS];   // [N,64]
    float* u1[GCRL_N_BLOCKS];       // relu(update_mlp.0) [N,64]
    float* y1;                      // relu(fc1) [N,64]
    float* y2;                      // relu(fc2) [N,64]
};

inline size_t stash_bytes(int64_t N, int64_t E) {
    const size_t n64 = align_up((size_t)N * 64 * 4);
    size_t b = 5 * align_up((size_t)E * 64 * 4);
    b += GCRL_N_BLOCKS * (7 * n64 + align_up((size_t)N * 256 * 4));
    b += 2 * n64;
    return b + 256;
}

inline Stash carve_stash(void* p, int64_t N, int64_t E) {
    Arena ar(p, (size_t)-1);
    Stash s;
    for (int i = 0; i < 5; ++i) s.e[i] = ar.take<float>((size_t)E * 64);
    for (int i = 0; i < GCRL_N_BLOCKS; ++i) {
        s.h[i] = ar.take<float>((size_t)N * 64);
        s.Pi[i] = ar.take<float>((size_t)N * 64);
        s.Pj[i] = ar.take<float>((size_t)N * 64);
        s.agg[i] = ar.take<float>((size_t)N * 256);
        s.var[i] = ar.take<float>((size_t)N * 64);
        s.amin[i] = ar.take<int32_t>((size_t)N * 64);
        s.amax[i] = ar.take<int32_t>((size_t)N * 64);
        s.u1[i] = ar.take<float>((size_t)N * 64);
    }
    s.y1 = ar.take<float>((size_t)N * 64);
    s.y2 = ar.take<float>((size_t)N * 64);
    return s;
}

}  // namespace gcrl
