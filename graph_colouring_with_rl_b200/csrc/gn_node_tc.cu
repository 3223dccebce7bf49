// gn_node_tc.cu -- vertex side of the GN forward pass on the tensor cores (math modes bf16 / bf16x3).
//
// GNNBlock.update (architecture.py:68-78: MLP([aggr | x]) with Linear -> ReLU -> Linear), the next block's
// per-vertex projections Pi = A h + b1, Pj = B h (the algebraic split of message_mlp.0, gn_forward.cu) and the
// fc1/fc2/fc3 head (architecture.py:191-197) are chains of row-GEMMs  out[N, 64 or 128] = act([A1 | A2] W^T + b).
// k_rows_gemm_tc runs one such GEMM: 128-vertex tiles, the K = 64-column chunks of A1 / A2 arrive by TMA (fp32,
// SWIZZLE_128B), are split into bf16 hi/lo operand tiles and accumulated in TMEM by tcgen05.mma; the epilogue adds
// the bias (and the K < 64 tail of block 1, whose vertex features are 2 wide), applies ReLU and leaves through a
// TMA store.  The fp32 FFMA kernel k_node_update (gn_forward.cu) stays the parity path of math mode fp32; on the
// 10 000 x 50-vertex rollout batch it was 29 % of a step (2.05 ms per block, 16 TFLOP/s of FFMA behind 12 LDS per
// 32 FMA).
#include "gn_common.cuh"
#include "tc.cuh"

namespace gcrl {

using namespace tc;

struct RowsGemmArgs {
    const float* W0; int ldW0, col0;       // weights of output columns [0, 64): element (o, k) = W0[o * ldW0 + col0 + k]
    const float* W1; int ldW1, col1;       // output columns [64, 128) (NOUT = 128)
    const float* bias0; const float* bias1;   // [64] or null
    int relu;
    int n1, n2;                            // 64-column chunks taken from tensor map 1 / tensor map 2
    const float* xt; int Dt;               // tail: out0[:, o] += sum_{d < Dt} Wt[o * ldWt + colt + d] * xt[row * Dt + d]
    const float* Wt; int ldWt, colt;
    int wT;                                // weights read transposed: element (o, k) = W[k * ldW + col + o]   (dgrad: dX = dY W);
                                           // K chunk c then takes the window col0 + 64 c (dX = dY_0 W[:, w_0] + dY_1 W[:, w_1] ...)
    int ngroups;                           // wT, one K chunk, NOUT = 64: output group g = dY W[:, col0 + 64 g : +64] goes to output
                                           // columns ocol0 + 64 g (one operand conversion, ngroups products; <= NODE_MAX_CHUNKS)
    int ocol0, ocol1;                      // column offsets of the two outputs inside their [N, ld] matrices
    const float* mask; int ldmask;         // out0[r][o] *= (mask[r * ldmask + o] > 0)   (relu' of a stashed activation) or null
    int accumulate;                        // outputs are added to memory (TMA reduce-add) instead of stored
    int64_t N;
};

// transposed flavour of load_weight_tile: tile row o (output), column k (contraction) <- W[(k) * ld + col0 + o]
__device__ __forceinline__ void load_weight_tile_T(const float* W, int ld, int col0, __nv_bfloat16* hi, __nv_bfloat16* lo,
                                                   int tid, int nthreads) {
    for (int t = tid; t < 64 * 8; t += nthreads) {
        const int o = t & 63, cb = t >> 6;          // consecutive threads read consecutive addresses
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = W[(8 * cb + j) * ld + col0 + o];
        uint4 h, l;
        split8(x, h, l);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(hi) + bf_off(o, cb)) = h;
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(lo) + bf_off(o, cb)) = l;
    }
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, int c0, int c1, const void* smem_src) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(m), "r"(c0), "r"(c1), "r"(smem_u32(smem_src)) : "memory");
}

constexpr int NODE_MAX_CHUNKS = 5;

template <int NOUT, bool X3>
__global__ void __launch_bounds__(256, 1) k_rows_gemm_tc(const __grid_constant__ CUtensorMap tm1,
                                                        const __grid_constant__ CUtensorMap tm2,
                                                        const __grid_constant__ CUtensorMap tmo0,
                                                        const __grid_constant__ CUtensorMap tmo1, const RowsGemmArgs a) {
    pdl_prologue();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int nch = a.n1 + a.n2;
    const int ngr = a.ngroups > 1 ? a.ngroups : 1;            // weight tiles in shared memory: nch chunks or ngr groups
    const int nwt = nch > ngr ? nch : ngr;
    constexpr int WCH = (NOUT / 64) * 16384;                  // bytes of one weight chunk: hi [NOUT x 64] | lo [NOUT x 64]
    unsigned char* Wsm = base;                                // [nch][WCH]
    float* L0 = reinterpret_cast<float*>(base + nwt * WCH);   // landing / staging tiles (fp32 SW128 slabs)
    float* L1 = L0 + 128 * 64;
    unsigned char* Aop = reinterpret_cast<unsigned char*>(L1 + 128 * 64);   // operand tile hi | lo
    float* bias_s = reinterpret_cast<float*>(Aop + 32768);    // [128]
    float* wt_s = bias_s + 128;                               // [64][8] tail weights
    uint64_t* bars = reinterpret_cast<uint64_t*>(wt_s + 64 * 8);
    uint64_t* bar_tma = bars;                                 // [2]
    uint64_t* bar_mma = bars + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sp = warp & 3, half = warp >> 2;
    const int row = 32 * sp + lane;
    if (tid == 0) {
        mbar_init(&bar_tma[0], 1);
        mbar_init(&bar_tma[1], 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    for (int c = 0; c < nwt; ++c) {
        __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(Wsm + c * WCH);
        __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(Wsm + c * WCH + WCH / 2);
        if (a.wT) {          // the contraction runs over the 64 rows of W; tile c = window col0 + 64 c (K chunk or output group)
            load_weight_tile_T(a.W0, a.ldW0, a.col0 + 64 * c, hi, lo, tid, 256);
            if (NOUT == 128) load_weight_tile_T(a.W1, a.ldW1, a.col1 + 64 * c, hi + 64 * 64, lo + 64 * 64, tid, 256);
        } else {
            load_weight_tile(a.W0, a.ldW0, a.col0 + 64 * c, hi, lo, tid, 256);
            if (NOUT == 128) load_weight_tile(a.W1, a.ldW1, a.col1 + 64 * c, hi + 64 * 64, lo + 64 * 64, tid, 256);
        }
    }
    if (tid < 128) bias_s[tid] = tid < 64 ? (a.bias0 ? a.bias0[tid] : 0.f) : ((NOUT == 128 && a.bias1) ? a.bias1[tid - 64] : 0.f);
    for (int t = tid; t < 64 * 8; t += 256) {
        const int o = t >> 3, d = t & 7;
        wt_s[t] = (a.Dt > 0 && d < a.Dt) ? a.Wt[o * a.ldWt + a.colt + d] : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_addr = (uint32_t)(32 * sp) << 16;
    const uint32_t idesc = make_idesc_bf16(128, NOUT, false, false);
    uint32_t ph_tma[2] = {0u, 0u}, ph_mma = 0u;

    auto issue_load = [&](int c, int32_t v0) {          // chunk c of the tile starting at row v0 -> landing buffer c & 1
        float* L = (c & 1) ? L1 : L0;
        const CUtensorMap* tm = c < a.n1 ? &tm1 : &tm2;
        const int col = 64 * (c < a.n1 ? c : c - a.n1);
        mbar_arrive_expect_tx(&bar_tma[c & 1], 128 * 64 * 4);
        tma_load_2d(L, tm, col, v0, &bar_tma[c & 1]);
        tma_load_2d(L + 128 * 32, tm, col + 32, v0, &bar_tma[c & 1]);
    };

    const int64_t n_tiles = (a.N + 127) / 128;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int32_t v0 = (int32_t)(tile * 128);
        if (tid == 0) {
            issue_load(0, v0);
            if (nch > 1) issue_load(1, v0);
        }
        for (int c = 0; c < nch; ++c) {
            const float* L = (c & 1) ? L1 : L0;
            mbar_wait_bounded(&bar_tma[c & 1], ph_tma[c & 1]);
            ph_tma[c & 1] ^= 1u;
            if (c > 0) {                                 // the previous chunk's MMA has read the operand tile
                mbar_wait_bounded(bar_mma, ph_mma);
                ph_mma ^= 1u;
            }
            convert_tile128<X3>(L, reinterpret_cast<__nv_bfloat16*>(Aop), reinterpret_cast<__nv_bfloat16*>(Aop + 16384), tid);
            fence_proxy_async();
            __syncthreads();                             // operand tile complete; landing buffer c & 1 is free again
            if (warp == 0 && elect_one()) {
                if (c + 2 < nch) issue_load(c + 2, v0);
                tc_fence_after();
                if (ngr > 1) {          // one operand, ngr weight windows -> ngr accumulators of 64 columns
                    for (int g = 0; g < ngr; ++g) {
                        const uint32_t w_hi = smem_u32(Wsm + g * WCH);
                        issue_gemm_feat<X3>(tmem + 64 * g, smem_u32(Aop), smem_u32(Aop) + 16384, w_hi, w_hi + WCH / 2, idesc, false);
                    }
                } else {
                    const uint32_t w_hi = smem_u32(Wsm + c * WCH);
                    issue_gemm_feat<X3>(tmem, smem_u32(Aop), smem_u32(Aop) + 16384, w_hi, w_hi + WCH / 2, idesc, c > 0);
                }
                umma_commit(bar_mma);
            }
        }
        mbar_wait_bounded(bar_mma, ph_mma);
        ph_mma ^= 1u;
        tc_fence_after();
        // ---- epilogue: bias (+ tail) (+ relu) -> fp32 staging tiles (L0: outputs 0-63, L1: outputs 64-127) ----------------
        const bool valid = v0 + row < a.N;
        float xt[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) xt[d] = (valid && d < a.Dt) ? a.xt[(int64_t)(v0 + row) * a.Dt + d] : 0.f;
        constexpr int PASSES = NOUT / 64;               // 32-column passes per warp
        for (int g = 0; g < ngr; ++g) {                 // output groups (NOUT = 64 only): staging tiles alternate
        float* Lg = (NOUT == 64 && (g & 1)) ? L1 : L0;
        if (g >= 2) {                                   // the store of group g - 2 has read this staging tile
            if (tid == 0) bulk_wait_read1();
            __syncthreads();
        }
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            // NOUT = 64: warp half h owns columns [32h, 32h+32); NOUT = 128: half h owns output matrix h, two passes
            const int col0 = NOUT == 64 ? 32 * half : 64 * half + 32 * p;
            float acc[32];
            tmem_ld32(tmem + lane_addr + col0 + 64 * g, acc);
            float* Lout = (NOUT == 128 && half == 1) ? L1 : Lg;
            const int lc0 = col0 & 63;                  // column inside the 64-wide staging tile
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int o = col0 + 4 * c + j;
                    float x = acc[4 * c + j] + bias_s[o];
                    if (a.Dt > 0 && o < 64) {
#pragma unroll
                        for (int d = 0; d < 8; ++d) x = fmaf(wt_s[o * 8 + d], xt[d], x);
                    }
                    if (a.mask && o < 64) x = (valid && a.mask[(int64_t)(v0 + row) * a.ldmask + o] > 0.f) ? x : 0.f;
                    v[j] = a.relu ? fmaxf(x, 0.f) : x;
                }
                *reinterpret_cast<float4*>(l_ptr(Lout, row, (lc0 >> 2) + c)) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            const int oc = a.ocol0 + 64 * g;
            if (a.accumulate) {                          // rows past N are clipped by the tensor map
                tma_reduce_add_2d(&tmo0, oc, v0, Lg);
                tma_reduce_add_2d(&tmo0, oc + 32, v0, Lg + 128 * 32);
                if (NOUT == 128) {
                    tma_reduce_add_2d(&tmo1, a.ocol1, v0, L1);
                    tma_reduce_add_2d(&tmo1, a.ocol1 + 32, v0, L1 + 128 * 32);
                }
            } else {
                tma_store_2d(&tmo0, oc, v0, Lg);
                tma_store_2d(&tmo0, oc + 32, v0, Lg + 128 * 32);
                if (NOUT == 128) {
                    tma_store_2d(&tmo1, a.ocol1, v0, L1);
                    tma_store_2d(&tmo1, a.ocol1 + 32, v0, L1 + 128 * 32);
                }
            }
            bulk_commit();
            if (g + 1 == ngr) bulk_wait_read0();
        }
        }
        __syncthreads();                                 // staging tiles are free for the next tile's loads
    }
    if (tid == 0) bulk_wait0();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

// host side: tensor map of a row-major fp32 matrix [rows, cols] (cols a multiple of 32), box {32, 128}
static int make_tmap_rows(CUtensorMap* out, const float* basep, int64_t rows, int cols);

template <int NOUT, bool X3>
static int launch_rows_gemm(const CUtensorMap& t1, const CUtensorMap& t2, const CUtensorMap& o0, const CUtensorMap& o1,
                            const RowsGemmArgs& a, cudaStream_t st) {
    const int nch = a.n1 + a.n2;
    const int nwt = a.ngroups > nch ? a.ngroups : nch;
    GCRL_CHECK_ARG(nwt <= NODE_MAX_CHUNKS);
    GCRL_CHECK_ARG(a.ngroups <= 1 || (NOUT == 64 && a.wT && nch == 1 && !a.bias0 && !a.mask && a.Dt == 0 && !a.relu));
    const int smem = nwt * (NOUT / 64) * 16384 + 2 * 32768 + 32768 + 128 * 4 + 64 * 8 * 4 + 64 + 1024;
    static int max_set = 0;
    if (smem > max_set) {
        GCRL_CUDA(cudaFuncSetAttribute(k_rows_gemm_tc<NOUT, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        max_set = smem;
    }
    int64_t tiles = (a.N + 127) / 128;
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    launch_k(k_rows_gemm_tc<NOUT, X3>, dim3((int)grid), dim3(256), smem, st, t1, t2, o0, o1, a);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tmap_rows(CUtensorMap* out, const float* basep, int64_t rows, int cols) {
    static EncodeTiledFn2 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn2)p;
    }
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return GCRL_ERR_CUDA;
    }
    if (rows < 1) rows = 1;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(basep), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) base=%p rows=%lld cols=%d", (int)r, (const void*)basep, (long long)rows, cols);
        return GCRL_ERR_CUDA;
    }
    return GCRL_OK;
}

int make_tmap_rows_pub(CUtensorMap* out, const float* basep, int64_t rows, int cols) { return make_tmap_rows(out, basep, rows, cols); }

// G[o][gcol0 + 64 c + k] += sum_r Y[r][o] X_c[r][k]: weight gradients of the vertex-side Linear layers.  Y [N,64] and the
// 64-column chunks X_c of up to two [N, 64 n] matrices arrive by TMA; both become bf16 hi/lo operand tiles that are
// read MN-major (the contraction runs over the 128 rows of a tile); one TMEM accumulator per chunk lives for the
// whole kernel and is flushed with one atomic pass per CTA.
struct RowsWgradArgs {
    int n1, n2;
    float* G; int ldG, gcol0;
    float* gbias;                          // [64] += column sums of Y (the bias gradient of the same Linear layer) or null
    int64_t N;
};

template <bool X3>
__global__ void __launch_bounds__(256, 1) k_rows_wgrad_tc(const __grid_constant__ CUtensorMap tmY,
                                                         const __grid_constant__ CUtensorMap tm1,
                                                         const __grid_constant__ CUtensorMap tm2, const RowsWgradArgs a) {
    pdl_prologue();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* L0 = reinterpret_cast<float*>(base);
    float* L1 = L0 + 128 * 64;
    unsigned char* Yop = reinterpret_cast<unsigned char*>(L1 + 128 * 64);
    unsigned char* Xop = Yop + 32768;
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(Xop + 32768);      // [2]
    uint64_t* bar_mma = bar_tma + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 3);
    float* bsum_s = reinterpret_cast<float*>(bar_tma + 4);             // [64] bias-gradient partial sums
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nch = a.n1 + a.n2;
    if (tid < 64) bsum_s[tid] = 0.f;
    float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);                     // columns 4 (tid & 15) .. +3 over this thread's rows
    if (tid == 0) {
        mbar_init(&bar_tma[0], 1);
        mbar_init(&bar_tma[1], 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t id_rows = make_idesc_bf16(64, 64, true, true);
    uint32_t ph_tma[2] = {0u, 0u}, ph_mma = 0u;
    // loads are numbered per tile: 0 = Y, 1 + c = X chunk c; load j lands in buffer j & 1
    auto issue_load = [&](int j, int32_t v0) {
        float* L = (j & 1) ? L1 : L0;
        const int c = j - 1;
        const CUtensorMap* tm = j == 0 ? &tmY : (c < a.n1 ? &tm1 : &tm2);
        const int col = j == 0 ? 0 : 64 * (c < a.n1 ? c : c - a.n1);
        mbar_arrive_expect_tx(&bar_tma[j & 1], 128 * 64 * 4);
        tma_load_2d(L, tm, col, v0, &bar_tma[j & 1]);
        tma_load_2d(L + 128 * 32, tm, col + 32, v0, &bar_tma[j & 1]);
    };
    const int64_t n_tiles = (a.N + 127) / 128;
    bool first_tile = true;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int32_t v0 = (int32_t)(tile * 128);
        if (tid == 0) { issue_load(0, v0); issue_load(1, v0); }
        for (int j = 0; j <= nch; ++j) {
            const float* L = (j & 1) ? L1 : L0;
            mbar_wait_bounded(&bar_tma[j & 1], ph_tma[j & 1]);
            ph_tma[j & 1] ^= 1u;
            if (j >= 2) {                                 // the previous chunk's MMA has read Xop
                mbar_wait_bounded(bar_mma, ph_mma);
                ph_mma ^= 1u;
            }
            unsigned char* dstop = j == 0 ? Yop : Xop;
            if (j == 0 && a.gbias) {
                // rows past N arrive as zeros (TMA out-of-bounds fill): they add nothing
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int t = it * 256 + tid;
                    const float4 v = *reinterpret_cast<const float4*>(l_ptr(L, t >> 4, t & 15));
                    bacc.x += v.x; bacc.y += v.y; bacc.z += v.z; bacc.w += v.w;
                }
            }
            convert_tile128<X3>(L, reinterpret_cast<__nv_bfloat16*>(dstop), reinterpret_cast<__nv_bfloat16*>(dstop + 16384), tid);
            fence_proxy_async();
            __syncthreads();
            if (warp == 0 && elect_one()) {
                if (j + 2 <= nch) issue_load(j + 2, v0);
                if (j >= 1) {
                    tc_fence_after();
                    issue_gemm_rows<X3>(tmem + 64 * (j - 1), smem_u32(Yop), smem_u32(Yop) + 16384, smem_u32(Xop), smem_u32(Xop) + 16384,
                                        id_rows, 128, !first_tile);
                    umma_commit(bar_mma);
                }
            }
        }
        mbar_wait_bounded(bar_mma, ph_mma);                       // all products of this tile are done: Yop / Xop are free
        ph_mma ^= 1u;
        first_tile = false;
    }
    if (!first_tile) {
        tc_fence_after();
        if (warp < 4) {
            const int o = 16 * warp + lane;               // accumulator row of an M = 64 product
            const uint32_t lane_addr = (uint32_t)(32 * warp) << 16;
            for (int c = 0; c < nch; ++c) {
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    float v[32];
                    tmem_ld32(tmem + 64 * c + lane_addr + 32 * hh, v);
                    if (lane < 16)
#pragma unroll
                        for (int jj = 0; jj < 32; ++jj) atomicAdd(&a.G[o * a.ldG + a.gcol0 + 64 * c + 32 * hh + jj], v[jj]);
                }
            }
        }
    }
    if (a.gbias) {
        const int c4 = tid & 15;
        atomicAdd(&bsum_s[4 * c4 + 0], bacc.x); atomicAdd(&bsum_s[4 * c4 + 1], bacc.y);
        atomicAdd(&bsum_s[4 * c4 + 2], bacc.z); atomicAdd(&bsum_s[4 * c4 + 3], bacc.w);
    }
    tc_fence_before();
    __syncthreads();
    if (a.gbias && !first_tile && tid < 64) atomicAdd(&a.gbias[tid], bsum_s[tid]);
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// ---- backward of one Linear layer in ONE launch (round 2) ---------------------------------------------------------------------
// For Y = dL/d(out) [N,64], the layer's input X = [X1 (64 n1 columns) | X2 (64 n2 columns)] and its weights W [64, ld]:
//   G[o][gcol0 + 64 c + k] += sum_r Y[r][o] X_c[r][k]        (weight gradient, one TMEM accumulator per chunk, flushed once)
//   gbias[o]               += sum_r Y[r][o]                   (bias gradient)
//   dX_c[r][k]             (+)= sum_o Y[r][o] W[o][wcol0 + 64 c + k]  [* (X_c[r][k] > 0)]     for the chunks in dx_chunks
// k_rows_wgrad_tc computes the first two; the third used to be separate k_rows_gemm_tc launches that converted the same Y
// tile again.  Here the Y operand tile (one byte layout for the row-contracting and the feature-contracting product) is
// converted once per 128-vertex tile; per chunk the weight-gradient product and the data-gradient product are issued
// together, and the data gradient's epilogue runs under the NEXT chunk's landing and conversion: it masks with the fp32
// landing tile of X_c itself (relu' of the stashed activation, when the layer's input is that activation) and overwrites
// that tile in place, which then leaves through a TMA store / reduce.  A 64-transition learn step's backward goes from 61
// launches to ~41 (profiles/r02_launches_learn_step_64x15-50.json).
struct LinearBwdArgs {
    int n1, n2;
    float* G; int ldG, gcol0;
    float* gbias;
    const float* W; int ldW, wcol0;
    const unsigned char* wimg;             // pre-converted tile of window wcol0 (window c at + c * 16 KB) or null
    uint32_t dx_chunks, mask_chunks;       // bit c: chunk c produces dX / is masked by X_c > 0
    int acc1, acc2;                        // outputs of map o1 (chunks < n1) / o2 (chunks >= n1) are added to memory
    int ocol1, ocol2;                      // column of chunk 0 / chunk n1 inside their output matrices
    int64_t N;
};

template <bool X3>
__global__ void __launch_bounds__(256, 1) k_linear_bwd_tc(const __grid_constant__ CUtensorMap tmY,
                                                         const __grid_constant__ CUtensorMap tm1,
                                                         const __grid_constant__ CUtensorMap tm2,
                                                         const __grid_constant__ CUtensorMap to1,
                                                         const __grid_constant__ CUtensorMap to2, const LinearBwdArgs a) {
    pdl_prologue();
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int nch = a.n1 + a.n2;
    unsigned char* Wsm = base;                                          // [nch] weight tiles (hi 8 KB | lo 8 KB), dX chunks only
    float* L0 = reinterpret_cast<float*>(base + nch * 16384);
    float* L1 = L0 + 128 * 64;
    unsigned char* Yop = reinterpret_cast<unsigned char*>(L1 + 128 * 64);
    unsigned char* Xop = Yop + 32768;
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(Xop + 32768);      // [2]
    uint64_t* bar_mma = bar_tma + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tma + 3);
    float* bsum_s = reinterpret_cast<float*>(bar_tma + 4);             // [64]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sp = warp & 3, half = warp >> 2;
    const int row = 32 * sp + lane;
    if (tid < 64) bsum_s[tid] = 0.f;
    float4 bacc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid == 0) {
        mbar_init(&bar_tma[0], 1);
        mbar_init(&bar_tma[1], 1);
        mbar_init(bar_mma, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(tmem_slot);
    if (a.wimg) {
        copy_image(Wsm, a.wimg, nch * 16384, tid, 256);                 // (windows without a data gradient are copied too: unused)
    } else {
        for (int c = 0; c < nch; ++c)
            if ((a.dx_chunks >> c) & 1u)
                load_weight_tile(a.W, a.ldW, a.wcol0 + 64 * c, reinterpret_cast<__nv_bfloat16*>(Wsm + c * 16384),
                                 reinterpret_cast<__nv_bfloat16*>(Wsm + c * 16384 + 8192), tid, 256);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t t_accd = tmem + 64 * NODE_MAX_CHUNKS;               // the data gradient's accumulator (after the <= 5 wgrad ones)
    const uint32_t lane_addr = (uint32_t)(32 * sp) << 16;
    const uint32_t id_rows = make_idesc_bf16(64, 64, true, true);
    const uint32_t id_featT = make_idesc_bf16(128, 64, false, true);
    uint32_t ph_tma[2] = {0u, 0u}, ph_mma = 0u;
    auto issue_load = [&](int j, int32_t v0) {          // load j of a tile: 0 = Y, 1 + c = X chunk c; lands in buffer j & 1
        float* L = (j & 1) ? L1 : L0;
        const int c = j - 1;
        const CUtensorMap* tm = j == 0 ? &tmY : (c < a.n1 ? &tm1 : &tm2);
        const int col = j == 0 ? 0 : 64 * (c < a.n1 ? c : c - a.n1);
        mbar_arrive_expect_tx(&bar_tma[j & 1], 128 * 64 * 4);
        tma_load_2d(L, tm, col, v0, &bar_tma[j & 1]);
        tma_load_2d(L + 128 * 32, tm, col + 32, v0, &bar_tma[j & 1]);
    };
    // data gradient of chunk c: accumulator -> (mask with X_c) -> in place over X_c's landing tile (buffer (c + 1) & 1)
    auto dx_epilogue = [&](int c) {
        float* L = ((c + 1) & 1) ? L1 : L0;
        float acc[32];
        tmem_ld32(t_accd + lane_addr + 32 * half, acc);
        const bool masked = (a.mask_chunks >> c) & 1u;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float4* p = reinterpret_cast<float4*>(l_ptr(L, row, 8 * half + k));
            float4 r = make_float4(acc[4 * k], acc[4 * k + 1], acc[4 * k + 2], acc[4 * k + 3]);
            if (masked) {
                const float4 m = *p;
                r.x = m.x > 0.f ? r.x : 0.f; r.y = m.y > 0.f ? r.y : 0.f;
                r.z = m.z > 0.f ? r.z : 0.f; r.w = m.w > 0.f ? r.w : 0.f;
            }
            *p = r;
        }
    };
    auto dx_store = [&](int c, int32_t v0) {             // thread 0 only (bulk groups are per thread)
        const float* L = ((c + 1) & 1) ? L1 : L0;
        const bool second = c >= a.n1;
        const CUtensorMap* to = second ? &to2 : &to1;
        const int col = second ? a.ocol2 + 64 * (c - a.n1) : a.ocol1 + 64 * c;
        if (second ? a.acc2 : a.acc1) {
            tma_reduce_add_2d(to, col, v0, L);
            tma_reduce_add_2d(to, col + 32, v0, L + 128 * 32);
        } else {
            tma_store_2d(to, col, v0, L);
            tma_store_2d(to, col + 32, v0, L + 128 * 32);
        }
        bulk_commit();
        bulk_wait_read0();
    };
    const int64_t n_tiles = (a.N + 127) / 128;
    bool first_tile = true;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int32_t v0 = (int32_t)(tile * 128);
        if (tid == 0) { issue_load(0, v0); issue_load(1, v0); }
        // ---- Y: bias gradient, operand tile ----------------------------------------------------------------------------------
        mbar_wait_bounded(&bar_tma[0], ph_tma[0]);
        ph_tma[0] ^= 1u;
        if (a.gbias) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int t = it * 256 + tid;
                const float4 v = *reinterpret_cast<const float4*>(l_ptr(L0, t >> 4, t & 15));
                bacc.x += v.x; bacc.y += v.y; bacc.z += v.z; bacc.w += v.w;
            }
        }
        convert_tile128<X3>(L0, reinterpret_cast<__nv_bfloat16*>(Yop), reinterpret_cast<__nv_bfloat16*>(Yop + 16384), tid);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0 && nch >= 2) issue_load(2, v0);                    // X chunk 1 -> the buffer Y has left
        // ---- X chunks --------------------------------------------------------------------------------------------------------
        for (int c = 0; c < nch; ++c) {
            const int j = c + 1;
            const float* L = (j & 1) ? L1 : L0;
            mbar_wait_bounded(&bar_tma[j & 1], ph_tma[j & 1]);
            ph_tma[j & 1] ^= 1u;
            const bool prev_dx = c > 0 && ((a.dx_chunks >> (c - 1)) & 1u);
            if (c > 0) {                                                // chunk c - 1's products have read Xop / written acc_d
                mbar_wait_bounded(bar_mma, ph_mma);
                ph_mma ^= 1u;
                tc_fence_after();
            }
            convert_tile128<X3>(L, reinterpret_cast<__nv_bfloat16*>(Xop), reinterpret_cast<__nv_bfloat16*>(Xop + 16384), tid);
            if (prev_dx) dx_epilogue(c - 1);
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0 && c > 0) {
                if (prev_dx) dx_store(c - 1, v0);                       // (waits until the store has read the tile)
                if (c + 2 <= nch) issue_load(c + 2, v0);                // X chunk c + 1 -> the buffer chunk c - 1 has left
            }
            if (warp == 0 && elect_one()) {
                tc_fence_after();
                issue_gemm_rows<X3>(tmem + 64 * c, smem_u32(Yop), smem_u32(Yop) + 16384, smem_u32(Xop), smem_u32(Xop) + 16384,
                                    id_rows, 128, !first_tile);
                if ((a.dx_chunks >> c) & 1u) {
                    const uint32_t w = smem_u32(Wsm + c * 16384);
                    issue_gemm_featT<X3>(t_accd, smem_u32(Yop), smem_u32(Yop) + 16384, w, w + 8192, id_featT, false);
                }
                umma_commit(bar_mma);
            }
        }
        mbar_wait_bounded(bar_mma, ph_mma);                             // the last chunk's products
        ph_mma ^= 1u;
        tc_fence_after();
        if ((a.dx_chunks >> (nch - 1)) & 1u) {
            dx_epilogue(nch - 1);
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) dx_store(nch - 1, v0);
        }
        tc_fence_before();
        __syncthreads();                                                // tiles, operands and acc_d are free for the next vertex tile
        tc_fence_after();
        first_tile = false;
    }
    if (tid == 0) bulk_wait0();
    if (!first_tile) {
        tc_fence_after();
        if (warp < 4) {
            const int o = 16 * warp + lane;               // accumulator row of an M = 64 product
            const uint32_t la = (uint32_t)(32 * warp) << 16;
            // 16-byte vector reds when every row of the gradient window is 16-byte aligned (not block 1's 258-wide U1)
            const bool vec_flush = (((a.ldG | a.gcol0) & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.G) & 15) == 0);
            for (int c = 0; c < nch; ++c) {
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    float v[32];
                    tmem_ld32(tmem + 64 * c + la + 32 * hh, v);
                    float* gp = &a.G[o * a.ldG + a.gcol0 + 64 * c + 32 * hh];
                    if (lane < 16) {
                        if (vec_flush) {
#pragma unroll
                            for (int jj = 0; jj < 32; jj += 4)
                                atomicAdd(reinterpret_cast<float4*>(gp + jj), make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]));
                        } else {
#pragma unroll
                            for (int jj = 0; jj < 32; ++jj) atomicAdd(gp + jj, v[jj]);
                        }
                    }
                }
            }
        }
    }
    if (a.gbias) {
        const int c4 = tid & 15;
        atomicAdd(&bsum_s[4 * c4 + 0], bacc.x); atomicAdd(&bsum_s[4 * c4 + 1], bacc.y);
        atomicAdd(&bsum_s[4 * c4 + 2], bacc.z); atomicAdd(&bsum_s[4 * c4 + 3], bacc.w);
    }
    tc_fence_before();
    __syncthreads();
    if (a.gbias && !first_tile && tid < 64) atomicAdd(&a.gbias[tid], bsum_s[tid]);
    if (warp == 0) tmem_dealloc<512>(tmem);
}

// Backward of one Linear layer (see k_linear_bwd_tc).  X1 [N, ld1] contributes n1 64-column chunks, X2 [N, ld2] n2; the data
// gradients of chunks < n1 go to dX1 [N, ldo1] at column ocol1 + 64 c, those of the X2 chunks to dX2 [N, ldo2] at ocol2.
int linear_bwd_tc(bool x3, const float* Y, const float* X1, int ld1, int n1, const float* X2, int ld2, int n2, float* G, int ldG,
                  int gcol0, float* gbias, const float* W, int ldW, int wcol0, uint32_t dx_chunks, uint32_t mask_chunks,
                  float* dX1, int ldo1, int ocol1, int acc1, float* dX2, int ldo2, int ocol2, int acc2, int64_t N, cudaStream_t st,
                  const unsigned char* wimg) {
    if (N == 0) return GCRL_OK;
    const int nch = n1 + n2;
    GCRL_CHECK_ARG(nch >= 1 && nch <= NODE_MAX_CHUNKS);
    CUtensorMap ty, t1, t2, o1, o2;
    int rc;
    if ((rc = make_tmap_rows(&ty, Y, N, 64))) return rc;
    if ((rc = make_tmap_rows(&t1, X1, N, ld1))) return rc;
    if ((rc = make_tmap_rows(&t2, n2 ? X2 : X1, N, n2 ? ld2 : ld1))) return rc;
    if ((rc = make_tmap_rows(&o1, dX1 ? dX1 : Y, N, dX1 ? ldo1 : 64))) return rc;
    if ((rc = make_tmap_rows(&o2, dX2 ? dX2 : Y, N, dX2 ? ldo2 : 64))) return rc;
    LinearBwdArgs a;
    a.n1 = n1; a.n2 = n2; a.G = G; a.ldG = ldG; a.gcol0 = gcol0; a.gbias = gbias;
    a.W = W; a.ldW = ldW; a.wcol0 = wcol0; a.wimg = dx_chunks ? wimg : nullptr; a.dx_chunks = dx_chunks; a.mask_chunks = mask_chunks;
    a.acc1 = acc1; a.acc2 = acc2; a.ocol1 = ocol1; a.ocol2 = ocol2; a.N = N;
    const int smem = nch * 16384 + 2 * 32768 + 2 * 32768 + 64 + 256 + 1024;
    static int max_set[2] = {0, 0};
    if (smem > max_set[x3 ? 1 : 0]) {
        if (x3) GCRL_CUDA(cudaFuncSetAttribute(k_linear_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        else GCRL_CUDA(cudaFuncSetAttribute(k_linear_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        max_set[x3 ? 1 : 0] = smem;
    }
    int64_t tiles = (N + 127) / 128;
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    if (x3) launch_k(k_linear_bwd_tc<true>, dim3((int)grid), dim3(256), smem, st, ty, t1, t2, o1, o2, a);
    else launch_k(k_linear_bwd_tc<false>, dim3((int)grid), dim3(256), smem, st, ty, t1, t2, o1, o2, a);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

// agg rows of vertices without incoming edges take the scatter defaults (0, 0, 0, sqrt(eps)): the edge kernel never
// visits them (principal_neighbourhood_aggregation.py:45-64 on an empty segment)
__global__ void k_fix_empty_agg(float* __restrict__ agg, const int32_t* __restrict__ seg_off, int64_t N) {
    pdl_prologue();
    const int64_t total = N * 64;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = t >> 6;
        const int c4 = (int)(t & 63);
        if (seg_off[v + 1] == seg_off[v]) {
            const float s = c4 >= 48 ? sqrtf(PNA_EPS) : 0.f;
            reinterpret_cast<float4*>(agg + v * 256)[c4] = make_float4(s, s, s, s);
        }
    }
}

// q[v] = y2[v, :] . F3 + f3
__global__ void k_head_dot(const float* __restrict__ y2, const float* __restrict__ F3, const float* __restrict__ f3,
                           int64_t N, float* __restrict__ q) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float w0 = F3[lane], w1 = F3[32 + lane], b = f3[0];
    for (int64_t v = warp0; v < N; v += nwarps) {
        float s = y2[v * 64 + lane] * w0 + y2[v * 64 + 32 + lane] * w1;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) q[v] = s + b;
    }
}

// Tensor-core vertex update; same contract as launch_node_update (gn_forward.cu) for Dg == 0 and Dv == 64 or Dv <= 8.
// u1_buf / y1_buf / y2_buf: [N,64] outputs (stash) or scratch.
int launch_node_update_tc(const NetW& net, int l, bool with_extra, bool x3, float* agg, const float* h_in,
                          const int32_t* seg_off, int64_t N, float* h_out, float* Pi, float* Pj, float* q,
                          float* u1_buf, float* y1_buf, float* y2_buf, cudaStream_t st) {
    if (N == 0) return GCRL_OK;
    const BlockW& b = net.blk[l];
    const bool last = with_extra && (l == GCRL_N_BLOCKS - 1);
    int rc;
    {
        int64_t blocks = (N * 64 + 255) / 256;
        if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
        launch_k(k_fix_empty_agg, dim3((int)blocks), dim3(256), 0, st, agg, seg_off, N);
        GCRL_LAUNCH_CHECK();
    }
    CUtensorMap t_agg, t_h, t_u1, t_ho, t_pi, t_pj, t_y1, t_y2;
    if ((rc = make_tmap_rows(&t_agg, agg, N, 256))) return rc;
    if ((rc = make_tmap_rows(&t_u1, u1_buf, N, 64))) return rc;
    if ((rc = make_tmap_rows(&t_ho, h_out, N, 64))) return rc;
    const bool wide_h = b.Dv == 64;
    if ((rc = make_tmap_rows(&t_h, wide_h ? h_in : u1_buf, N, 64))) return rc;
    const int K1 = 256 + b.Dv;
    RowsGemmArgs a;
    // update_mlp.0 + ReLU
    a.W0 = b.U1; a.ldW0 = K1; a.col0 = 0; a.W1 = nullptr; a.ldW1 = 0; a.col1 = 0;
    a.bias0 = b.c1; a.bias1 = nullptr; a.relu = 1;
    a.n1 = 4; a.n2 = wide_h ? 1 : 0;
    a.xt = wide_h ? nullptr : h_in; a.Dt = wide_h ? 0 : b.Dv; a.Wt = b.U1; a.ldWt = K1; a.colt = 256;
    a.N = N; a.wT = 0; a.ngroups = 1; a.ocol0 = 0; a.ocol1 = 0; a.mask = nullptr; a.ldmask = 0; a.accumulate = 0;
    rc = x3 ? launch_rows_gemm<64, true>(t_agg, t_h, t_u1, t_u1, a, st) : launch_rows_gemm<64, false>(t_agg, t_h, t_u1, t_u1, a, st);
    if (rc) return rc;
    // update_mlp.2
    a.W0 = b.U2; a.ldW0 = 64; a.col0 = 0; a.bias0 = b.c2; a.relu = 0; a.n1 = 1; a.n2 = 0; a.xt = nullptr; a.Dt = 0;
    rc = x3 ? launch_rows_gemm<64, true>(t_u1, t_u1, t_ho, t_ho, a, st) : launch_rows_gemm<64, false>(t_u1, t_u1, t_ho, t_ho, a, st);
    if (rc) return rc;
    if (!with_extra) return GCRL_OK;
    if (!last) {
        // next block's projections: Pi = A h + b1, Pj = B h  (one N = 128 product)
        const BlockW& nb = net.blk[l + 1];
        const int ldn = 2 * nb.Dv + nb.De;
        if ((rc = make_tmap_rows(&t_pi, Pi, N, 64))) return rc;
        if ((rc = make_tmap_rows(&t_pj, Pj, N, 64))) return rc;
        a.W0 = nb.W1; a.ldW0 = ldn; a.col0 = 0; a.W1 = nb.W1; a.ldW1 = ldn; a.col1 = 64;
        a.bias0 = nb.b1; a.bias1 = nullptr; a.relu = 0; a.n1 = 1; a.n2 = 0;
        rc = x3 ? launch_rows_gemm<128, true>(t_ho, t_ho, t_pi, t_pj, a, st) : launch_rows_gemm<128, false>(t_ho, t_ho, t_pi, t_pj, a, st);
        return rc;
    }
    // head: fc1 + ReLU, fc2 + ReLU, fc3
    if ((rc = make_tmap_rows(&t_y1, y1_buf, N, 64))) return rc;
    if ((rc = make_tmap_rows(&t_y2, y2_buf, N, 64))) return rc;
    a.W0 = net.head.F1; a.ldW0 = 64; a.col0 = 0; a.W1 = nullptr; a.bias0 = net.head.f1; a.bias1 = nullptr; a.relu = 1; a.n1 = 1; a.n2 = 0;
    rc = x3 ? launch_rows_gemm<64, true>(t_ho, t_ho, t_y1, t_y1, a, st) : launch_rows_gemm<64, false>(t_ho, t_ho, t_y1, t_y1, a, st);
    if (rc) return rc;
    a.W0 = net.head.F2; a.bias0 = net.head.f2;
    rc = x3 ? launch_rows_gemm<64, true>(t_y1, t_y1, t_y2, t_y2, a, st) : launch_rows_gemm<64, false>(t_y1, t_y1, t_y2, t_y2, a, st);
    if (rc) return rc;
    {
        int64_t blocks = (N * 32 + 255) / 256;
        if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
        launch_k(k_head_dot, dim3((int)blocks), dim3(256), 0, st, y2_buf, net.head.F3, net.head.f3, N, q);
        GCRL_LAUNCH_CHECK();
    }
    return GCRL_OK;
}


// dX[N, 64] (+)= (dY[N,64] . W[64, wcol : wcol + 64]) * (mask > 0), on the tensor cores.  dX may be a 64-column window
// (column xcol) of a wider matrix [N, ldx].
//   dY2 != null: dX (+)= dY . W[:, wcol : +64] + dY2 . W[:, wcol + 64 : +64]   (two K chunks, one launch)
//   ngroups > 1: dX[:, xcol + 64 g : +64] = dY . W[:, wcol + 64 g : +64] for g < ngroups (one operand conversion)
int dgrad64_tc(bool x3, const float* dY, const float* W, int ldw, int wcol, float* dX, int ldx, int xcol, const float* mask,
               int accumulate, int64_t N, cudaStream_t st, const float* dY2 = nullptr, int ngroups = 1) {
    if (N == 0) return GCRL_OK;
    CUtensorMap ty, ty2, tx;
    int rc;
    if ((rc = make_tmap_rows(&ty, dY, N, 64))) return rc;
    if ((rc = make_tmap_rows(&ty2, dY2 ? dY2 : dY, N, 64))) return rc;
    if ((rc = make_tmap_rows(&tx, dX, N, ldx))) return rc;
    RowsGemmArgs a;
    a.W0 = W; a.ldW0 = ldw; a.col0 = wcol; a.W1 = nullptr; a.ldW1 = 0; a.col1 = 0;
    a.bias0 = nullptr; a.bias1 = nullptr; a.relu = 0; a.n1 = 1; a.n2 = dY2 ? 1 : 0;
    a.xt = nullptr; a.Dt = 0; a.Wt = nullptr; a.ldWt = 0; a.colt = 0;
    a.wT = 1; a.ngroups = ngroups; a.ocol0 = xcol; a.ocol1 = 0; a.mask = mask; a.ldmask = 64; a.accumulate = accumulate; a.N = N;
    return x3 ? launch_rows_gemm<64, true>(ty, ty2, tx, tx, a, st) : launch_rows_gemm<64, false>(ty, ty2, tx, tx, a, st);
}

// G[o][gcol0 + k] += sum_r Y[r][o] X[r][k] for X = [X1 (64 n1 columns) | X2 (64 n2 columns)]
int wgrad64_tc(bool x3, const float* Y, const float* X1, int ld1, int n1, const float* X2, int ld2, int n2, float* G, int ldG,
               int gcol0, float* gbias, int64_t N, cudaStream_t st) {
    if (N == 0) return GCRL_OK;
    CUtensorMap ty, t1, t2;
    int rc;
    if ((rc = make_tmap_rows(&ty, Y, N, 64))) return rc;
    if ((rc = make_tmap_rows(&t1, X1, N, ld1))) return rc;
    if ((rc = make_tmap_rows(&t2, n2 ? X2 : X1, N, n2 ? ld2 : ld1))) return rc;
    RowsWgradArgs a;
    a.n1 = n1; a.n2 = n2; a.G = G; a.ldG = ldG; a.gcol0 = gcol0; a.gbias = gbias; a.N = N;
    const int smem = 2 * 32768 + 2 * 32768 + 64 + 256 + 1024;
    static bool done = false;
    if (!done) {
        GCRL_CUDA(cudaFuncSetAttribute(k_rows_wgrad_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GCRL_CUDA(cudaFuncSetAttribute(k_rows_wgrad_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        done = true;
    }
    int64_t tiles = (N + 127) / 128;
    int64_t grid = sm_count();
    if (grid > tiles) grid = tiles;
    if (x3) launch_k(k_rows_wgrad_tc<true>, dim3((int)grid), dim3(256), smem, st, ty, t1, t2, a);
    else launch_k(k_rows_wgrad_tc<false>, dim3((int)grid), dim3(256), smem, st, ty, t1, t2, a);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // namespace gcrl
