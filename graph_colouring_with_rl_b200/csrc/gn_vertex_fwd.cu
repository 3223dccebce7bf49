// gn_vertex_fwd.cu -- the whole vertex side of one GNNBlock's forward pass in ONE launch (math modes bf16 / bf16x3).
//
// Reference: GNNBlock.update (architecture.py:68-78: MLP([aggr | x]), Linear -> ReLU -> Linear), then either the next
// block's per-vertex projections Pi = A h + b1, Pj = B h (the algebraic split of message_mlp.0, gn_forward.cu) or, after
// block 5, the fc1/fc2/fc3 head (architecture.py:191-197).  gn_node_tc.cu runs these as a chain of row-GEMM launches
// (fix_empty_agg, U1, U2, [A|B] or fc1, fc2, fc3-dot: 4-6 launches per block); each launch is ~10 us of fixed latency
// whatever the row count, which is what a 64-transition learn step, a batch-1 act or a rollout shard of a thousand
// graphs consists of (profiles/r02_launches_learn_step_64x15-50.json).  Here a CTA walks 128-vertex tiles and keeps the
// chain on chip:
//   agg (4 K-chunks by TMA) | h (1 chunk, or the 2-wide x of block 1 as an epilogue tail)
//     -> U1 (5 accumulating tcgen05 products) -> +c1, ReLU -> u1 (operand tile in shared memory; fp32 copy to the stash)
//     -> U2 -> +c2 -> h_out (operand tile + TMA store)
//     -> [A | B] (one N = 128 product) -> Pi (+b1), Pj                                   (blocks 1-4)
//     -> fc1 -> ReLU -> y1 -> fc2 -> ReLU -> y2 -> q = y2 . F3 + f3  (registers + one shared-memory exchange)   (block 5)
// Two 32 KB tiles serve in turn as TMA landing tile, bf16 hi|lo operand (converted in place through registers) and fp32
// staging tile of the TMA stores; weights (bf16 hi|lo, 128 KB) stay resident for all tiles of the CTA.
// Rows without incoming edges (single-vertex graphs) take the scatter defaults (0, 0, 0, sqrt(eps)): patched in the landing
// tile and written back to the global agg array (the backward reads it), replacing the k_fix_empty_agg launch.
#include "gn_common.cuh"
#include "tc.cuh"

namespace gcrl {

using namespace tc;

struct VertexFwdArgs {
    const float* U1; int ldU1; const float* c1;
    const float* U2; const float* c2;
    const float* W3a; int ldW3a; int col3a;     // third product, output columns [0, 64):   Wn[:, 0:64]  (A)  or F1
    const float* W3b; int ldW3b; int col3b;     //                 output columns [64, 128): Wn[:, 64:128] (B) or F2 (LAST: 4th product)
    const float* b3a; const float* b3b;         // b1 of the next block / null   or f1 / f2
    const float* F3; const float* f3;           // LAST
    const float* xt; int Dv;                    // Dv <= 8: h_in = x [N, Dv] enters as an epilogue tail (U1 columns 256..)
    int nch;                                    // K chunks of U1 arriving by TMA: 4 (agg) + (Dv == 64)
    const int32_t* seg_off; float* agg;
    float* q;
    int store_u1, store_y;                      // write the activations the backward needs
    const unsigned char* img_u;                 // pre-converted tiles U1 chunk 0..4 | U2 (six consecutive tiles) or null
    const unsigned char* img_3a; const unsigned char* img_3b;   // tiles of W3a / W3b or null
    int64_t N;
};

// fp32 SW128-slab landing tile -> bf16 hi|lo operand tile IN PLACE (256 threads; every row is read before any is written)
template <bool X3>
__device__ __forceinline__ void convert_inplace128(unsigned char* T, int tid) {
    const float* Lf = reinterpret_cast<const float*>(T);
    float4 x[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int t = it * 256 + tid;
        x[it] = *reinterpret_cast<const float4*>(l_ptr(Lf, t >> 4, t & 15));
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int t = it * 256 + tid;
        const int r = t >> 4, c4 = t & 15;
        uint2 h, l;
        split4<X3>(x[it], h, l);
        const uint32_t off = bf_off(r, c4 >> 1) + (c4 & 1) * 8;
        *reinterpret_cast<uint2*>(T + off) = h;
        if (X3) *reinterpret_cast<uint2*>(T + 16384 + off) = l;
    }
}

constexpr int VF_W1_BYTES = 5 * 16384;          // U1: five K chunks, hi 8 KB | lo 8 KB each
constexpr int VF_W2_BYTES = 16384;
constexpr int VF_W3_BYTES = 32768;              // [A | B] as one 128-row tile (hi 16 KB | lo 16 KB), or F1 tile + F2 tile
constexpr int VF_SMEM = VF_W1_BYTES + VF_W2_BYTES + VF_W3_BYTES + 2 * 32768 + (64 + 64 + 128 + 64 + 4) * 4 + 64 * 8 * 4 +
                        2 * 128 * 4 + 64 + 1024;

template <bool LAST, bool X3>
__global__ void __launch_bounds__(256, 1) k_vertex_fwd_tc(const __grid_constant__ CUtensorMap t_agg,
                                                         const __grid_constant__ CUtensorMap t_h,
                                                         const __grid_constant__ CUtensorMap t_u1,
                                                         const __grid_constant__ CUtensorMap t_ho,
                                                         const __grid_constant__ CUtensorMap t_o1,
                                                         const __grid_constant__ CUtensorMap t_o2, const VertexFwdArgs a) {
    pdl_prologue();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* W1s = base;
    unsigned char* W2s = W1s + VF_W1_BYTES;
    unsigned char* W3s = W2s + VF_W2_BYTES;
    unsigned char* T0 = W3s + VF_W3_BYTES;
    unsigned char* T1 = T0 + 32768;
    float* c1_s = reinterpret_cast<float*>(T1 + 32768);      // [64]
    float* c2_s = c1_s + 64;                                 // [64]
    float* b3_s = c2_s + 64;                                 // [128]
    float* f3_s = b3_s + 128;                                // [64] F3, [64] = f3
    float* wt_s = f3_s + 68;                                 // [64][8] tail weights of U1 (block 1)
    float* qpart = wt_s + 64 * 8;                            // [2][128]
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(qpart + 2 * 128);   // [2]
    uint64_t* bar_mma = bar_tma + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sp = warp & 3, half = warp >> 2;
    const int row = 32 * sp + lane;
    if (tid == 0) {
        mbar_init(&bar_tma[0], 1);
        mbar_init(&bar_tma[1], 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
        tma_prefetch_desc(&t_agg);
        tma_prefetch_desc(&t_ho);
    }
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    if (a.img_u) {
        copy_image(W1s, a.img_u, VF_W1_BYTES + VF_W2_BYTES, tid, 256);      // U1 chunks 0..4 | U2: contiguous in the image and here
    } else {
        for (int c = 0; c < a.nch; ++c)
            load_weight_tile(a.U1, a.ldU1, 64 * c, reinterpret_cast<__nv_bfloat16*>(W1s + c * 16384),
                             reinterpret_cast<__nv_bfloat16*>(W1s + c * 16384 + 8192), tid, 256);
        load_weight_tile(a.U2, 64, 0, reinterpret_cast<__nv_bfloat16*>(W2s), reinterpret_cast<__nv_bfloat16*>(W2s + 8192), tid, 256);
    }
    if (LAST) {          // F1 tile | F2 tile, each hi 8 KB | lo 8 KB
        if (a.img_3a) {
            copy_image(W3s, a.img_3a, 16384, tid, 256);
            copy_image(W3s + 16384, a.img_3b, 16384, tid, 256);
        } else {
            load_weight_tile(a.W3a, a.ldW3a, a.col3a, reinterpret_cast<__nv_bfloat16*>(W3s), reinterpret_cast<__nv_bfloat16*>(W3s + 8192), tid, 256);
            load_weight_tile(a.W3b, a.ldW3b, a.col3b, reinterpret_cast<__nv_bfloat16*>(W3s + 16384),
                             reinterpret_cast<__nv_bfloat16*>(W3s + 16384 + 8192), tid, 256);
        }
    } else {             // one 128-row tile: rows 0-63 = A, rows 64-127 = B; hi 16 KB | lo 16 KB
        if (a.img_3a) {
            copy_image(W3s, a.img_3a, 8192, tid, 256);                      // A hi
            copy_image(W3s + 16384, a.img_3a + 8192, 8192, tid, 256);       // A lo
            copy_image(W3s + 8192, a.img_3b, 8192, tid, 256);               // B hi
            copy_image(W3s + 24576, a.img_3b + 8192, 8192, tid, 256);       // B lo
        } else {
            load_weight_tile(a.W3a, a.ldW3a, a.col3a, reinterpret_cast<__nv_bfloat16*>(W3s), reinterpret_cast<__nv_bfloat16*>(W3s + 16384), tid, 256);
            load_weight_tile(a.W3b, a.ldW3b, a.col3b, reinterpret_cast<__nv_bfloat16*>(W3s + 8192),
                             reinterpret_cast<__nv_bfloat16*>(W3s + 16384 + 8192), tid, 256);
        }
    }
    if (tid < 64) {
        c1_s[tid] = a.c1[tid];
        c2_s[tid] = a.c2[tid];
        b3_s[tid] = a.b3a ? a.b3a[tid] : 0.f;
        b3_s[64 + tid] = a.b3b ? a.b3b[tid] : 0.f;
        if (LAST) {
            f3_s[tid] = a.F3[tid];
            if (tid == 0) f3_s[64] = a.f3[0];
        }
    }
    for (int t = tid; t < 64 * 8; t += 256) {
        const int o = t >> 3, d = t & 7;
        wt_s[t] = (a.Dv <= 8 && d < a.Dv) ? a.U1[o * a.ldU1 + 256 + d] : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_accA = tmem, t_accB = tmem + 64, t_accC = tmem + 128;       // U1 | U2 (and fc2) | third product (128 columns)
    const uint32_t lane_addr = (uint32_t)(32 * sp) << 16;
    const uint32_t id64 = make_idesc_bf16(128, 64, false, false);
    const uint32_t id128 = make_idesc_bf16(128, 128, false, false);
    uint32_t ph_tma[2] = {0u, 0u}, ph_mma = 0u;
    const float sq_eps = sqrtf(PNA_EPS);

    auto issue_load = [&](int c, int32_t v0) {          // K chunk c of the tile starting at vertex v0 -> tile c & 1
        unsigned char* T = (c & 1) ? T1 : T0;
        const CUtensorMap* tm = c < 4 ? &t_agg : &t_h;
        const int col = c < 4 ? 64 * c : 0;
        mbar_arrive_expect_tx(&bar_tma[c & 1], 128 * 64 * 4);
        tma_load_2d(T, tm, col, v0, &bar_tma[c & 1]);
        tma_load_2d(T + 16384, tm, col + 32, v0, &bar_tma[c & 1]);
    };
    auto mma_wait = [&]() {
        mbar_wait_bounded(bar_mma, ph_mma);
        ph_mma ^= 1u;
        tc_fence_after();
    };
    // act(acc + bias) of this thread's row and 32 columns -> operand tile Top (bf16 hi|lo) and, if Tst, fp32 staging tile
    auto epilogue = [&](uint32_t t_acc, const float* bias_s, bool relu, bool tail, const float* xt, unsigned char* Top,
                        unsigned char* Tst, float* keep) {
        float acc[32];
        tmem_ld32(t_acc + lane_addr + 32 * half, acc);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float x = acc[j] + bias_s[32 * half + j];
            if (tail) {
#pragma unroll
                for (int d = 0; d < 8; ++d) x = fmaf(wt_s[(32 * half + j) * 8 + d], xt[d], x);
            }
            acc[j] = relu ? fmaxf(x, 0.f) : x;
        }
        if (Top) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                store_split8<X3>(&acc[8 * c], reinterpret_cast<__nv_bfloat16*>(Top), reinterpret_cast<__nv_bfloat16*>(Top + 16384), row,
                                 4 * half + c);
        }
        if (Tst) {
            float* Lf = reinterpret_cast<float*>(Tst);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(l_ptr(Lf, row, 8 * half + c)) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        }
        if (keep) {
#pragma unroll
            for (int j = 0; j < 32; ++j) keep[j] = acc[j];
        }
    };
    auto store_tile = [&](const CUtensorMap* tm, int32_t v0, const unsigned char* T) {     // one thread; rows past N are clipped
        tma_store_2d(tm, 0, v0, T);
        tma_store_2d(tm, 32, v0, T + 16384);
        bulk_commit();
    };

    const int64_t n_tiles = (a.N + 127) / 128;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int32_t v0 = (int32_t)(tile * 128);
        const bool valid = v0 + row < a.N;
        if (tid == 0) {
            issue_load(0, v0);
            issue_load(1, v0);
        }
        // rows without incoming edges (deg == 0): scatter defaults, here and in the global array the backward reads
        bool empty_row = false;
        if (tid < 128 && v0 + tid < a.N) empty_row = a.seg_off[v0 + tid + 1] == a.seg_off[v0 + tid];
        if (empty_row) {
            float4* g = reinterpret_cast<float4*>(a.agg + (int64_t)(v0 + tid) * 256);
            for (int c4 = 0; c4 < 64; ++c4) {
                const float s = c4 >= 48 ? sq_eps : 0.f;
                g[c4] = make_float4(s, s, s, s);
            }
        }
        float xt[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) xt[d] = (a.Dv <= 8 && valid && d < a.Dv) ? a.xt[(int64_t)(v0 + row) * a.Dv + d] : 0.f;
        // ---- U1: accumulate the K chunks ------------------------------------------------------------------------------------
        for (int c = 0; c < a.nch; ++c) {
            unsigned char* T = (c & 1) ? T1 : T0;
            mbar_wait_bounded(&bar_tma[c & 1], ph_tma[c & 1]);
            ph_tma[c & 1] ^= 1u;
            if (empty_row && c < 4) {
                float* Lf = reinterpret_cast<float*>(T);
                const float s = c == 3 ? sq_eps : 0.f;
                for (int k = 0; k < 16; ++k) *reinterpret_cast<float4*>(l_ptr(Lf, tid, k)) = make_float4(s, s, s, s);
            }
            __syncthreads();                                 // (the patch above precedes every thread's reads)
            convert_inplace128<X3>(T, tid);
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                const uint32_t w = smem_u32(W1s + c * 16384);
                issue_gemm_feat<X3>(t_accA, smem_u32(T), smem_u32(T) + 16384, w, w + 8192, id64, c > 0);
                umma_commit(bar_mma);
            }
            mma_wait();                                      // the product has read tile c & 1
            if (tid == 0 && c + 2 < a.nch) issue_load(c + 2, v0);
        }
        // ---- u1 = relu(acc + c1 [+ tail]) -> operand (T0), stash copy (T1) ----------------------------------------------------
        epilogue(t_accA, c1_s, true, a.Dv <= 8, xt, T0, a.store_u1 ? T1 : nullptr, nullptr);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0 && a.store_u1) store_tile(&t_u1, v0, T1);     // (bulk groups are per thread: thread 0 issues and waits them all)
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            issue_gemm_feat<X3>(t_accB, smem_u32(T0), smem_u32(T0) + 16384, smem_u32(W2s), smem_u32(W2s) + 8192, id64, false);
            umma_commit(bar_mma);
        }
        mma_wait();
        if (tid == 0 && a.store_u1) bulk_wait_read0();       // T1 is about to be rewritten
        __syncthreads();
        // ---- h_out = acc + c2 -> operand (T0), TMA store (T1) ------------------------------------------------------------------
        epilogue(t_accB, c2_s, false, false, xt, T0, T1, nullptr);
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) store_tile(&t_ho, v0, T1);
        if (warp == 0 && elect_one()) {
            tc_fence_after();
            if (LAST) issue_gemm_feat<X3>(t_accC, smem_u32(T0), smem_u32(T0) + 16384, smem_u32(W3s), smem_u32(W3s) + 8192, id64, false);
            else issue_gemm_feat<X3>(t_accC, smem_u32(T0), smem_u32(T0) + 16384, smem_u32(W3s), smem_u32(W3s) + 16384, id128, false);
            umma_commit(bar_mma);
        }
        mma_wait();
        if (tid == 0) bulk_wait_read0();                     // h_out has left T1
        __syncthreads();
        if (!LAST) {
            // ---- Pi = acc[:, 0:64] + b1, Pj = acc[:, 64:128]: warp half h owns output h (two 32-column passes) -----------------
            unsigned char* To = half ? T1 : T0;
            float* Lf = reinterpret_cast<float*>(To);
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                float acc[32];
                tmem_ld32(t_accC + lane_addr + 64 * half + 32 * p, acc);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 bb = *reinterpret_cast<const float4*>(&b3_s[64 * half + 32 * p + 4 * c]);
                    *reinterpret_cast<float4*>(l_ptr(Lf, row, 8 * p + c)) =
                        make_float4(acc[4 * c] + bb.x, acc[4 * c + 1] + bb.y, acc[4 * c + 2] + bb.z, acc[4 * c + 3] + bb.w);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                store_tile(&t_o1, v0, T0);
                store_tile(&t_o2, v0, T1);
                bulk_wait_read0();
            }
        } else {
            // ---- head: y1 = relu(fc1), y2 = relu(fc2), q = y2 . F3 + f3 ----------------------------------------------------------
            epilogue(t_accC, b3_s, true, false, xt, T0, a.store_y ? T1 : nullptr, nullptr);
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0 && a.store_y) store_tile(&t_o1, v0, T1);
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                issue_gemm_feat<X3>(t_accB, smem_u32(T0), smem_u32(T0) + 16384, smem_u32(W3s + 16384), smem_u32(W3s + 16384) + 8192, id64, false);
                umma_commit(bar_mma);
            }
            mma_wait();
            if (tid == 0 && a.store_y) bulk_wait_read0();
            __syncthreads();
            float y2[32];
            epilogue(t_accB, b3_s + 64, true, false, xt, nullptr, a.store_y ? T1 : nullptr, y2);
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) part = fmaf(y2[j], f3_s[32 * half + j], part);
            qpart[half * 128 + row] = part;
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (half == 0 && valid) a.q[v0 + row] = (qpart[row] + qpart[128 + row]) + f3_s[64];
            if (tid == 0 && a.store_y) {
                store_tile(&t_o2, v0, T1);
                bulk_wait_read0();
            }
        }
        tc_fence_before();
        __syncthreads();                                     // tiles and accumulators are free for the next vertex tile
        tc_fence_after();
    }
    if (tid == 0) bulk_wait0();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

// gn_node_tc.cu
int make_tmap_rows_pub(CUtensorMap* out, const float* basep, int64_t rows, int cols);

template <bool LAST, bool X3>
static int launch_vertex_variant(const CUtensorMap* tm, const VertexFwdArgs& a, cudaStream_t st) {
    static bool done = false;
    if (!done) {
        GCRL_CUDA(cudaFuncSetAttribute(k_vertex_fwd_tc<LAST, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, VF_SMEM));
        done = true;
    }
    int64_t tiles = (a.N + 127) / 128;
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    launch_k(k_vertex_fwd_tc<LAST, X3>, dim3((int)grid), dim3(256), VF_SMEM, st, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], a);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

// Fused vertex side of block l (0-based) of the forward pass; same contract as launch_node_update_tc with with_extra = true
// (gn_node_tc.cu).  stash: u1 / y1 / y2 are wanted by the backward (else the three buffers are ignored).
int launch_vertex_fwd_fused(const NetW& net, int l, bool x3, float* agg, const float* h_in, const int32_t* seg_off, int64_t N,
                            float* h_out, float* Pi, float* Pj, float* q, float* u1_buf, float* y1_buf, float* y2_buf, bool stash,
                            cudaStream_t st, const unsigned char* wimg) {
    if (N == 0) return GCRL_OK;
    const BlockW& b = net.blk[l];
    const bool last = l == GCRL_N_BLOCKS - 1;
    VertexFwdArgs a;
    a.U1 = b.U1; a.ldU1 = 256 + b.Dv; a.c1 = b.c1; a.U2 = b.U2; a.c2 = b.c2;
    a.Dv = b.Dv; a.xt = b.Dv <= 8 ? h_in : nullptr; a.nch = 4 + (b.Dv == 64 ? 1 : 0);
    a.seg_off = seg_off; a.agg = agg; a.q = q; a.N = N;
    a.store_u1 = stash ? 1 : 0; a.store_y = (stash && last) ? 1 : 0;
    a.F3 = nullptr; a.f3 = nullptr;
    a.img_u = wimg_tile(wimg, l, WT_U1);
    a.img_3a = wimg ? (last ? wimg + (size_t)WIMG_F1 * WIMG_TILE_BYTES : wimg_tile(wimg, l + 1, WT_WA)) : nullptr;
    a.img_3b = wimg ? (last ? wimg + (size_t)WIMG_F2 * WIMG_TILE_BYTES : wimg_tile(wimg, l + 1, WT_WB)) : nullptr;
    if (last) {
        a.W3a = net.head.F1; a.ldW3a = 64; a.col3a = 0; a.W3b = net.head.F2; a.ldW3b = 64; a.col3b = 0;
        a.b3a = net.head.f1; a.b3b = net.head.f2; a.F3 = net.head.F3; a.f3 = net.head.f3;
    } else {
        const BlockW& nb = net.blk[l + 1];
        const int ldn = 2 * nb.Dv + nb.De;
        a.W3a = nb.W1; a.ldW3a = ldn; a.col3a = 0; a.W3b = nb.W1; a.ldW3b = ldn; a.col3b = 64;
        a.b3a = nb.b1; a.b3b = nullptr;
    }
    CUtensorMap tm[6];
    int rc;
    if ((rc = make_tmap_rows_pub(&tm[0], agg, N, 256))) return rc;
    if ((rc = make_tmap_rows_pub(&tm[1], b.Dv == 64 ? h_in : h_out, N, 64))) return rc;
    if ((rc = make_tmap_rows_pub(&tm[2], stash ? u1_buf : h_out, N, 64))) return rc;
    if ((rc = make_tmap_rows_pub(&tm[3], h_out, N, 64))) return rc;
    if ((rc = make_tmap_rows_pub(&tm[4], last ? ((stash && y1_buf) ? y1_buf : h_out) : Pi, N, 64))) return rc;
    if ((rc = make_tmap_rows_pub(&tm[5], last ? ((stash && y2_buf) ? y2_buf : h_out) : Pj, N, 64))) return rc;
    if (last) return x3 ? launch_vertex_variant<true, true>(tm, a, st) : launch_vertex_variant<true, false>(tm, a, st);
    return x3 ? launch_vertex_variant<false, true>(tm, a, st) : launch_vertex_variant<false, false>(tm, a, st);
}

}  // namespace gcrl
