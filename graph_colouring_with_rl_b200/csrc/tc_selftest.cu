// tc_selftest.cu -- tensor-map construction and a self-test of the tcgen05/TMA plumbing in tc.cuh:
// one CTA computes, through the same descriptors the edge kernels use,
//   D  [128,64] = A_tile * W^T            (K-major bf16x3 operands, M = 128, N = 64, K = 64)
//   DW [64,64]  = A_tile^T * A2_tile      (MN-major bf16x3 operands, M = 64, N = 64, K = 128)
// with the A tiles brought in either by plain loads or by TMA, and optionally TMA-stores a tile.
#include "tc.cuh"
#include <cudaTypedefs.h>

namespace gcrl {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_tmap_rows64(CUtensorMap* out, const float* base, int64_t rows, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return GCRL_ERR_CUDA;
    }
    if (rows < 1) rows = 1;
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) base=%p rows=%lld", (int)r, (const void*)base, (long long)rows);
        return GCRL_ERR_CUDA;
    }
    return GCRL_OK;
}

struct SelfSmem {
    float L[128 * 64];             // fp32 landing tile (SW128 slabs)
    float L2[128 * 64];
    __nv_bfloat16 A[2][128 * 64];  // hi / lo operand tiles
    __nv_bfloat16 A2[2][128 * 64];
    __nv_bfloat16 W[2][64 * 64];
    uint64_t bar_tma, bar_mma;
    uint32_t tmem_base;
};

// fp32 SW128-slab tile -> bf16 hi/lo operand tiles (column-per-lane mapping, 128 threads)
__device__ __forceinline__ void convert_tile(const float* L, __nv_bfloat16* hi, __nv_bfloat16* lo, int tid, int nthreads) {
    for (int t = tid; t < 128 * 8; t += nthreads) {
        const int r = t >> 3, cb = t & 7;
        float x[8];
        *reinterpret_cast<float4*>(&x[0]) = *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(L) + tc::sw_off<128>(r, 2 * cb));
        *reinterpret_cast<float4*>(&x[4]) = *reinterpret_cast<const float4*>(reinterpret_cast<const unsigned char*>(L) + tc::sw_off<128>(r, 2 * cb + 1));
        uint4 h, l;
        tc::split8(x, h, l);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(hi) + tc::bf_off(r, cb)) = h;
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(lo) + tc::bf_off(r, cb)) = l;
    }
}

__global__ void __launch_bounds__(128, 1) k_tc_selftest(const __grid_constant__ CUtensorMap tmA,
                                                       const __grid_constant__ CUtensorMap tmA2,
                                                       const __grid_constant__ CUtensorMap tmOut,
                                                       const float* __restrict__ A, const float* __restrict__ A2,
                                                       const float* __restrict__ W, int64_t rows_total, int row0,
                                                       float* __restrict__ D, float* __restrict__ DW, int use_tma,
                                                       int do_store) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    SelfSmem& S = *reinterpret_cast<SelfSmem*>(smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        tc::mbar_init(&S.bar_tma, 1);
        tc::mbar_init(&S.bar_mma, 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) tc::tmem_alloc<256>(&S.tmem_base);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tm = S.tmem_base;

    // weights: W[o][k] -> rows = o (N), columns = k (K)
    for (int t = tid; t < 64 * 8; t += 128) {
        const int o = t >> 3, cb = t & 7;
        float x[8];
        for (int j = 0; j < 8; ++j) x[j] = W[o * 64 + 8 * cb + j];
        uint4 h, l;
        tc::split8(x, h, l);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(S.W[0]) + tc::bf_off(o, cb)) = h;
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(S.W[1]) + tc::bf_off(o, cb)) = l;
    }
    if (use_tma & 1) {
        if (tid == 0) {
            tc::mbar_arrive_expect_tx(&S.bar_tma, 2 * 128 * 64 * 4);
            tc::tma_load_2d(S.L, &tmA, 0, row0, &S.bar_tma);
            tc::tma_load_2d(S.L + 128 * 32, &tmA, 32, row0, &S.bar_tma);
            tc::tma_load_2d(S.L2, &tmA2, 0, row0, &S.bar_tma);
            tc::tma_load_2d(S.L2 + 128 * 32, &tmA2, 32, row0, &S.bar_tma);
        }
        tc::mbar_wait(&S.bar_tma, 0);
    } else {
        for (int t = tid; t < 128 * 16; t += 128) {
            const int r = t >> 4, c4 = t & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), v2 = v;
            if (row0 + r < rows_total) {
                v = *reinterpret_cast<const float4*>(A + (int64_t)(row0 + r) * 64 + 4 * c4);
                v2 = *reinterpret_cast<const float4*>(A2 + (int64_t)(row0 + r) * 64 + 4 * c4);
            }
            *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(S.L) + tc::sw_off<128>(r, c4)) = v;
            *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(S.L2) + tc::sw_off<128>(r, c4)) = v2;
        }
        __syncthreads();
    }
    convert_tile(S.L, S.A[0], S.A[1], tid, 128);
    convert_tile(S.L2, S.A2[0], S.A2[1], tid, 128);
    const bool a_in_tmem = (use_tma & 2) != 0;      // D through the A-in-TMEM product (columns [128,192) hold the operand)
    if (a_in_tmem) {
        // thread = tile row: its 64 fp32 values -> bf16 hi | lo, 32 columns of K at a time
        for (int half = 0; half < 2; ++half) {
            float v[32];
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(&v[4 * c]) = *reinterpret_cast<const float4*>(tc::l_ptr(S.L, tid, 8 * half + c));
            tc::tmem_st_split32<true>(tm + ((uint32_t)(32 * warp) << 16) + 128 + 32 * half, v);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc::tc_fence_before();
    }
    tc::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        tc::tc_fence_after();
        // D = A * W^T into columns [0,64)
        if (a_in_tmem)
            tc::issue_gemm_feat_ts<true>(tm, tm + 128, tc::smem_u32(S.W[0]), tc::smem_u32(S.W[1]),
                                         tc::make_idesc_bf16(128, 64, false, false), false);
        else
        tc::issue_gemm_feat<true>(tm, tc::smem_u32(S.A[0]), tc::smem_u32(S.A[1]), tc::smem_u32(S.W[0]), tc::smem_u32(S.W[1]),
                                  tc::make_idesc_bf16(128, 64, false, false), false);
        // DW = A^T * A2 into columns [64,128): M = 64 (A columns), N = 64 (A2 columns), K = 128 rows
        tc::issue_gemm_rows<true>(tm + 64, tc::smem_u32(S.A[0]), tc::smem_u32(S.A[1]), tc::smem_u32(S.A2[0]),
                                  tc::smem_u32(S.A2[1]), tc::make_idesc_bf16(64, 64, true, true), 128, false);
        tc::umma_commit(&S.bar_mma);
    }
    tc::mbar_wait(&S.bar_mma, 0);
    tc::tc_fence_after();
    {
        float v[32];
        const int r = 32 * warp + lane;   // D: lane = row
        for (int half = 0; half < 2; ++half) {
            tc::tmem_ld32(tm + ((uint32_t)(32 * warp) << 16) + 32 * half, v);
            for (int j = 0; j < 32; ++j) D[r * 64 + 32 * half + j] = v[j];
        }
        // DW: M = 64 accumulator row m lives in lane (m % 16) + 32 * (m / 16)
        for (int half = 0; half < 2; ++half) {
            tc::tmem_ld32(tm + ((uint32_t)(32 * warp) << 16) + 64 + 32 * half, v);
            if (lane < 16) {
                const int m = 16 * warp + lane;
                for (int j = 0; j < 32; ++j) DW[m * 64 + 32 * half + j] = v[j];
            }
        }
    }
    if (do_store && (use_tma & 1)) {
        // TMA store of the (unchanged) fp32 landing tile to rows row0.. of tmOut
        __syncthreads();
        if (tid == 0) {
            tc::tma_store_2d(&tmOut, 0, row0, S.L);
            tc::tma_store_2d(&tmOut, 32, row0, S.L + 128 * 32);
            tc::bulk_commit();
            tc::bulk_wait0();
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tm);
}

}  // namespace gcrl

using namespace gcrl;

extern "C" int gcrl_selftest_tc(const float* A, const float* A2, const float* W, int64_t rows_total, int row0,
                                float* D, float* DW, float* copy_out, int use_tma, void* stream) {
    GCRL_CHECK_ARG(A && A2 && W && D && DW && rows_total >= 1 && row0 >= 0);
    CUtensorMap tmA, tmA2, tmOut;
    int rc = make_tmap_rows64(&tmA, A, rows_total, 128);
    if (rc) return rc;
    rc = make_tmap_rows64(&tmA2, A2, rows_total, 128);
    if (rc) return rc;
    rc = make_tmap_rows64(&tmOut, copy_out ? copy_out : A, rows_total, 128);
    if (rc) return rc;
    const int smem = (int)sizeof(SelfSmem) + 1024;
    GCRL_CUDA(cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_tc_selftest<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmA2, tmOut, A, A2, W, rows_total, row0, D, DW, use_tma,
                                                         copy_out ? 1 : 0);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}
