// gn_forward_tc.cu -- the fused edge kernel of the GN forward pass on the 5th-gen tensor cores.
//
// Same math and same fusion as k_edge_fwd in gn_forward.cu (message MLP of architecture.py:51-60
// with the first Linear split into per-vertex projections, PNA aggregation of :62-66 fused in),
// but the two 64x64 products per edge run as tcgen05.mma (kind::f16, bf16 hi/lo split operands,
// fp32 accumulators in TMEM) and the edge stream moves by TMA:
//
//   per 128-edge tile                                         who
//   1. TMA load e_{l-1} tile (fp32, 32 KB) -> L               async (issued one tile ahead of use)
//   2. split L into bf16 hi/lo operand tile B                 all threads
//   3. G1: acc1 = B . C^T   (12 or 4 UMMAs)                   one thread issues, tensor core runs
//   4. gather Pi[dst] + Pj[src] -> L  (overlaps G1)           all threads, coalesced rows
//   5. hid = relu(acc1 + L) -> bf16 hi/lo -> B                TMEM -> registers -> smem
//   6. G2: acc2 = B . W2^T
//   7. m = acc2 + b2 -> L (fp32)                              TMEM -> registers -> smem
//   8. TMA store L -> e_l ; segmented mean/min/max/std of L   async store + column scan
//
// Shared memory per CTA ~110 KB (weights 32 KB, L 32 KB, B 32 KB, reducer 12 KB) so two CTAs share
// an SM and hide each other's TMA / MMA latencies; TMEM: 128 columns per CTA.
// Block 1 (edge_dim-wide input) skips steps 1-3: acc1 = C e is a few FMAs per element.
#define GCRL_MBAR_TIMEOUT   // counted waits (with a trap on protocol bugs): measurably faster here than the tight loop
#include "gn_common.cuh"
#include "tc.cuh"

namespace gcrl {

using namespace tc;

constexpr int TC_MAX_DE_FIRST = 8;   // widest raw edge feature handled by the FIRST variant

struct FwdTcSmem {
    __nv_bfloat16 Wc[2][64 * 64];    // C  (message_mlp.0 edge columns) hi/lo, rows = o, cols = k
    __nv_bfloat16 W2[2][64 * 64];    // message_mlp.2 hi/lo
    float L[128 * 64];               // fp32 SW128-slab tile: landing / Pi+Pj / m
    __nv_bfloat16 B[2][128 * 64];    // operand tile hi/lo: e_{l-1}, then hid
    RedScratch red;
    float b2[64];
    int32_t dst[TILE];
    int32_t src[TILE];
    uint32_t start_bits[4];          // bit r of word q: tile row 32q + r starts a new destination
    uint64_t bar_tma, bar_mma;
    uint32_t tmem_base;
};

struct FwdTcArgs {
    const float* Pi; const float* Pj;
    const float* e_prev; int De;           // FIRST only reads e_prev directly
    const float* W1; int ldW1, coff;
    const float* W2; const float* b2;
    const int32_t* seg_off; const int32_t* src; const int32_t* dst;
    int32_t N, E;
    float* m_out;
    AggOut out;
};

template <bool FIRST, bool X3>
__global__ void __launch_bounds__(256, 2) k_edge_fwd_tc(const __grid_constant__ CUtensorMap tm_in,
                                                       const __grid_constant__ CUtensorMap tm_out, const FwdTcArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // round up inside the shared window (pointer arithmetic keeps the address space: LDS/STS, not generic)
    FwdTcSmem& S = *reinterpret_cast<FwdTcSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sp = warp & 3, half = warp >> 2;      // TMEM sub-partition (rows 32sp..) and column half
    const int row = 32 * sp + lane;                 // this thread's tile row in the epilogues

    if (tid == 0) {
        mbar_init(&S.bar_tma, 1);
        mbar_init(&S.bar_mma, 1);
        fence_barrier_init();
        if (!FIRST) tma_prefetch_desc(&tm_in);
        if (a.m_out) tma_prefetch_desc(&tm_out);
    }
    if (warp == 0) tmem_alloc<128>(&S.tmem_base);
    load_weight_tile(a.W2, 64, 0, S.W2[0], S.W2[1], tid, 256);
    if (!FIRST) load_weight_tile(a.W1, a.ldW1, a.coff, S.Wc[0], S.Wc[1], tid, 256);
    float (*Cf)[64] = reinterpret_cast<float (*)[64]>(&S.Wc[0][0]);   // FIRST: C^T in fp32 [d][o] (Wc is unused)
    if (FIRST)
        for (int t = tid; t < a.De * 64; t += 256) Cf[t >> 6][t & 63] = a.W1[(t & 63) * a.ldW1 + a.coff + (t >> 6)];
    if (tid < 64) S.b2[tid] = a.b2[tid];
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const uint32_t t_acc1 = tmem, t_acc2 = tmem + 64;
    const uint32_t lane_addr = (uint32_t)(32 * sp) << 16;
    const uint32_t idesc = make_idesc_bf16(128, 64, false, false);
    uint32_t ph_tma = 0, ph_mma = 0;

    const int64_t n_items = ((int64_t)a.E + ITEM_EDGES - 1) / ITEM_EDGES;
    for (int64_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int32_t v0 = seg_lower_bound(a.seg_off, a.N, w * ITEM_EDGES);
        const int32_t v1 = (w + 1 == n_items) ? a.N : seg_lower_bound(a.seg_off, a.N, (w + 1) * ITEM_EDGES);
        const int32_t e0 = a.seg_off[v0], e1 = a.seg_off[v1];
        Carry carry;
        carry.dst = -1;
        stat_init(carry.s);
        if (!FIRST && tid == 0 && e0 < e1) {
            mbar_arrive_expect_tx(&S.bar_tma, 128 * 64 * 4);
            tma_load_2d(S.L, &tm_in, 0, e0, &S.bar_tma);
            tma_load_2d(S.L + 128 * 32, &tm_in, 32, e0, &S.bar_tma);
        }
        // endpoint ids of the item's first tile; later tiles are prefetched one tile ahead
        int32_t nd = -1, ns = 0;
        if (tid < TILE && e0 + tid < e1) { nd = __ldg(a.dst + e0 + tid); ns = __ldg(a.src + e0 + tid); }
        for (int32_t t0 = e0; t0 < e1; t0 += TILE) {
            const int nrows = min(TILE, e1 - t0);
            if (tid < TILE) { S.dst[tid] = nd; S.src[tid] = ns; }
            __syncthreads();
            if (tid < TILE) {
                // segment-start bits for the reducer; next tile's ids (in flight for a whole tile)
                const bool st = tid > 0 && tid < nrows && S.dst[tid] != S.dst[tid - 1];
                const uint32_t bits = __ballot_sync(0xffffffffu, st);
                if (lane == 0) S.start_bits[warp] = bits;
                nd = -1; ns = 0;
                if (t0 + TILE + tid < e1) { nd = __ldg(a.dst + t0 + TILE + tid); ns = __ldg(a.src + t0 + TILE + tid); }
            }
            // ---- 4a. issue the Pi[dst] / Pj[src] row gathers (consumed after the operand split) ---------------
            float4 pg[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int t = it * 256 + tid;
                const int r = t >> 4, c4 = t & 15;
                pg[it] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nrows) {
                    const float4 pi = __ldg(reinterpret_cast<const float4*>(a.Pi + (int64_t)S.dst[r] * 64) + c4);
                    const float4 pj = __ldg(reinterpret_cast<const float4*>(a.Pj + (int64_t)S.src[r] * 64) + c4);
                    pg[it] = make_float4(pi.x + pj.x, pi.y + pj.y, pi.z + pj.z, pi.w + pj.w);
                }
            }
            if (!FIRST) {
                // ---- 2. e_{l-1}: fp32 landing tile -> bf16 hi/lo operand tile ---------------------------
                mbar_wait(&S.bar_tma, ph_tma);
                ph_tma ^= 1;
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int t = it * 256 + tid;
                    const int r = t >> 4, c4 = t & 15;      // half a warp per 256-byte row: conflict-free
                    const float4 v = *reinterpret_cast<const float4*>(l_ptr(S.L, r, c4));
                    uint2 h, l;
                    split4<X3>(v, h, l);
                    const uint32_t off = bf_off(r, c4 >> 1) + (c4 & 1) * 8;
                    *reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(S.B[0]) + off) = h;
                    if (X3) *reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(S.B[1]) + off) = l;
                }
                fence_proxy_async();
                __syncthreads();
                if (warp == 0 && elect_one()) {   // one elected lane of a converged warp: no per-thread election loop around each UMMA
                    // ---- 3. G1 -------------------------------------------------------------------------------
                    tc_fence_after();
                    issue_gemm_feat<X3>(t_acc1, smem_u32(S.B[0]), smem_u32(S.B[1]), smem_u32(S.Wc[0]), smem_u32(S.Wc[1]), idesc, false);
                    umma_commit(&S.bar_mma);
                }
            }
            // ---- 4b. (Pi + Pj) -> L (same association as the fp32 path: (Pi + Pj) + C e) -----------------
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int t = it * 256 + tid;
                *reinterpret_cast<float4*>(l_ptr(S.L, t >> 4, t & 15)) = pg[it];
            }
            __syncthreads();
            // ---- 5. hid = relu(acc1 + L) -> B ----------------------------------------------------------------
            float acc[32];
            if (!FIRST) {
                mbar_wait(&S.bar_mma, ph_mma);
                ph_mma ^= 1;
                tc_fence_after();
                tmem_ld32(t_acc1 + lane_addr + 32 * half, acc);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = 0.f;
                if (row < nrows) {
                    for (int d = 0; d < a.De; ++d) {
                        const float ev = a.e_prev[(int64_t)(t0 + row) * a.De + d];
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = fmaf(ev, Cf[d][32 * half + j], acc[j]);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float x[8];
                const float4 p0 = *reinterpret_cast<float4*>(l_ptr(S.L, row, 8 * half + 2 * c));
                const float4 p1 = *reinterpret_cast<float4*>(l_ptr(S.L, row, 8 * half + 2 * c + 1));
                x[0] = fmaxf(p0.x + acc[8 * c + 0], 0.f); x[1] = fmaxf(p0.y + acc[8 * c + 1], 0.f);
                x[2] = fmaxf(p0.z + acc[8 * c + 2], 0.f); x[3] = fmaxf(p0.w + acc[8 * c + 3], 0.f);
                x[4] = fmaxf(p1.x + acc[8 * c + 4], 0.f); x[5] = fmaxf(p1.y + acc[8 * c + 5], 0.f);
                x[6] = fmaxf(p1.z + acc[8 * c + 6], 0.f); x[7] = fmaxf(p1.w + acc[8 * c + 7], 0.f);
                uint4 h, l;
                if (X3) {
                    split8(x, h, l);
                } else {
                    h = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
                }
                *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(S.B[0]) + bf_off(row, 4 * half + c)) = h;
                if (X3) *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(S.B[1]) + bf_off(row, 4 * half + c)) = l;
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (warp == 0 && elect_one()) {   // one elected lane of a converged warp: no per-thread election loop around each UMMA
                // ---- 6. G2 -------------------------------------------------------------------------------
                tc_fence_after();
                issue_gemm_feat<X3>(t_acc2, smem_u32(S.B[0]), smem_u32(S.B[1]), smem_u32(S.W2[0]), smem_u32(S.W2[1]), idesc, false);
                umma_commit(&S.bar_mma);
            }
            mbar_wait(&S.bar_mma, ph_mma);
            ph_mma ^= 1;
            tc_fence_after();
            // ---- 7. m = acc2 + b2 -> L --------------------------------------------------------------------------
            tmem_ld32(t_acc2 + lane_addr + 32 * half, acc);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 bb = *reinterpret_cast<const float4*>(&S.b2[32 * half + 4 * c]);
                *reinterpret_cast<float4*>(l_ptr(S.L, row, 8 * half + c)) =
                    make_float4(acc[4 * c] + bb.x, acc[4 * c + 1] + bb.y, acc[4 * c + 2] + bb.z, acc[4 * c + 3] + bb.w);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            // ---- 8. store e_l + segmented PNA reduction ------------------------------------------------------------
            const bool tma_store = a.m_out && nrows == TILE;
            if (tma_store) {
                if (tid == 0) {
                    tma_store_2d(&tm_out, 0, t0, S.L);
                    tma_store_2d(&tm_out, 32, t0, S.L + 128 * 32);
                    bulk_commit();
                }
            } else if (a.m_out) {
                float4* g = reinterpret_cast<float4*>(a.m_out + (int64_t)t0 * 64);
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int idx = it * 256 + tid;
                    const int r = idx >> 4, c4 = idx & 15;
                    if (r < nrows) __stcs(g + idx, *reinterpret_cast<const float4*>(l_ptr(S.L, r, c4)));
                }
            }
            if (a.out.amin) tile_reduce_scan_sw<true>(S.L, S.dst, S.start_bits, nrows, t0, &S.red, a.out);
            else tile_reduce_scan_sw<false>(S.L, S.dst, S.start_bits, nrows, t0, &S.red, a.out);
            if (tma_store && tid == 0) bulk_wait_read0();
            __syncthreads();
            if (tid < 64) tile_reduce_chain(nrows, &S.red, carry, a.out);
            // L is free again: fetch the next tile of this item
            if (!FIRST && tid == 0 && t0 + TILE < e1) {
                mbar_arrive_expect_tx(&S.bar_tma, 128 * 64 * 4);
                tma_load_2d(S.L, &tm_in, 0, t0 + TILE, &S.bar_tma);
                tma_load_2d(S.L + 128 * 32, &tm_in, 32, t0 + TILE, &S.bar_tma);
            }
        }
        if (tid < 64) carry_flush(carry, a.out);
        __syncthreads();   // reducer scratch and L quiescent before the next item
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<128>(tmem);
}

template <bool FIRST, bool X3>
static int launch_variant(const CUtensorMap& tin, const CUtensorMap& tout, const FwdTcArgs& a, int grid, cudaStream_t st) {
    const int smem = (int)sizeof(FwdTcSmem) + 1024;
    static bool done = false;
    if (!done) {
        GCRL_CUDA(cudaFuncSetAttribute(k_edge_fwd_tc<FIRST, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        done = true;
    }
    k_edge_fwd_tc<FIRST, X3><<<grid, 256, smem, st>>>(tin, tout, a);
    return GCRL_OK;
}

// Tensor-core edge forward; same contract as launch_edge_fwd (gn_forward.cu).
int launch_edge_fwd_tc(const BlockW& b, bool first, bool x3, const float* Pi, const float* Pj, const float* e_prev,
                       const int32_t* seg_off, const int32_t* src, const int32_t* dst, int64_t N, int64_t E,
                       float* m_out, float* agg, float* var, int32_t* amin, int32_t* amax, cudaStream_t st) {
    if (E == 0) return GCRL_OK;
    FwdTcArgs a;
    a.Pi = Pi; a.Pj = Pj; a.e_prev = e_prev; a.De = b.De;
    a.W1 = b.W1; a.ldW1 = 2 * b.Dv + b.De; a.coff = 2 * b.Dv;
    a.W2 = b.W2; a.b2 = b.b2;
    a.seg_off = seg_off; a.src = src; a.dst = dst; a.N = (int32_t)N; a.E = (int32_t)E;
    a.m_out = m_out;
    a.out.agg = agg; a.out.var = var; a.out.amin = amin; a.out.amax = amax; a.out.seg_off = seg_off;
    CUtensorMap tin, tout;
    int rc = make_tmap_rows64(&tin, first ? Pi : e_prev, first ? N : E, 128);   // FIRST never dereferences tin
    if (rc) return rc;
    rc = make_tmap_rows64(&tout, m_out ? m_out : Pi, m_out ? E : N, 128);
    if (rc) return rc;
    int64_t items = (E + ITEM_EDGES - 1) / ITEM_EDGES;
    int64_t grid = (int64_t)sm_count() * 2;
    if (grid > items) grid = items;
    const int tok = prof_begin(first ? GCRL_PROF_EDGE_FWD_FIRST : (m_out ? GCRL_PROF_EDGE_FWD : GCRL_PROF_EDGE_FWD_LAST), st);
    if (first) rc = x3 ? launch_variant<true, true>(tin, tout, a, (int)grid, st) : launch_variant<true, false>(tin, tout, a, (int)grid, st);
    else rc = x3 ? launch_variant<false, true>(tin, tout, a, (int)grid, st) : launch_variant<false, false>(tin, tout, a, (int)grid, st);
    prof_end(tok, st);
    if (rc) return rc;
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // namespace gcrl
