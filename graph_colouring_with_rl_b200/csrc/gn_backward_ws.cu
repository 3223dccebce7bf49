// gn_backward_ws.cu -- warp-specialised, pipelined edge kernel of the GN backward pass.
//
// Same math as k_edge_bwd_tc (gn_backward_tc.cu; reference: loss.backward() through
// architecture.py:51-66, dqn_agent.py:219) but the per-tile phases no longer run one after the
// other in one group of threads: one persistent CTA per SM of 20 warps, each warp group owns a
// stage, and stages of consecutive 128-edge tiles overlap through mbarriers:
//
//   warp 0   lane 0   TMA producer, ring E : e_{l-1} boxes (32 rows x 64 fp32, 8 KB)
//   warp 2   lane 0   TMA producer, ring MD: (e_l, d e_l) box pairs (16 KB)
//   warp 1   lane 0   tcgen05.mma issuer (G1..G6 below), tcgen05.commit -> mbarriers
//   WG_A  warps 4-7   e_{l-1} box -> bf16 hi/lo operand tile B1[s];  Pi[dst]+Pj[src] gathered in
//                     the 16x256b fragment layout and written to TMEM (tcgen05.st) so that G1
//                     accumulates on top of it: no shared-memory staging of the projections
//   WG_B  warps 8-11  dm = d e_l + ca[v] + cb[v] * e_l + arg terms -> operand tile B3
//                     (block 5: e_5 recomputed by G2, dm from TMEM)
//   WG_C  warps 12-15 epilogue 1: hid = relu(acc1) -> B2, relu mask in registers
//                     epilogue 3: dz1 = acc4 * relu' -> B2; then, off the critical path,
//                     dPj[src] += dz1 (vector reds) and dPi[dst] += segment sums (warp
//                     transpose-reduce, one red per column per segment)
//   WG_D  warps 16-19 epilogue 4: d e_{l-1} = acc6 -> fp32 staging (the dead B1[s]) -> TMA store
//
//   G1  ACC_A[s] += e_{l-1} . C^T   (on top of Pi+Pj)       G4  ACC_A[s]  = dm . W2
//   G2  ACC_B[s]  = hid . W2^T      (block 5 only)          G5  accC    += dz1^T . e_{l-1}
//   G3  accW2    += dm^T . hid                              G6  ACC_B[s]  = dz1 . C
//
// s = tile parity: the e_{l-1} operand tile, both TMEM accumulators and the staging tile are
// double-buffered, so G1 / epilogue 1 of tile i+1 and epilogue 4 / the store of tile i-1 overlap
// the G3..G6 chain of tile i.  hid/dz1 (B2) and dm (B3) are single tiles: in bf16x3 mode the
// operand tiles are 32 KB each and shared memory (227 KB) is the binding resource:
//   weights 32 + B1 2x32 + B2 32 + B3 32 + ring E 4x8 + ring MD 2x16 = 224 KB.
// HBM traffic per edge is unchanged (blocks 2-4: read 768 B, write 256 B).
#include "gn_common.cuh"
#include "ws.cuh"
#include <stdlib.h>

namespace gcrl {

using namespace tc;

constexpr int WS_NE = 4;
constexpr int WS_NMD = 2;
constexpr int WS_THREADS = 768;

struct BwdWsSmem {
    __nv_bfloat16 Wc[2][64 * 64];        // C hi/lo (FIRST: C^T in fp32 [d][o])
    __nv_bfloat16 W2[2][64 * 64];
    unsigned char B1[2][32768];          // e_{l-1} operand (hi 16 KB | lo 16 KB); later d e_{l-1} fp32 staging
    unsigned char B2[32768];             // hid, then dz1
    unsigned char B3[32768];             // dm
    unsigned char ringE[WS_NE][8192];    // fp32 SW128 slabs, 32 rows
    unsigned char ringMD[WS_NMD][16384]; // e_l box (8 KB) | d e_l box (8 KB)
    float b2[64];
    float bsum[64];
    uint64_t bar_Efull[WS_NE], bar_Eempty[WS_NE], bar_MDfull[WS_NMD], bar_MDempty[WS_NMD];
    uint64_t bar_B1full[2], bar_B1free[2], bar_accBfree[2], bar_acc2full[2], bar_acc6full[2];
    uint64_t bar_accAfree[3], bar_acc1full[3], bar_acc4full[3];      // ACC_A (and the relu mask) is triple-buffered
    uint64_t bar_B2full, bar_B3full, bar_dz1full, bar_B2read, bar_done;
    uint64_t bar_OUTfull[2];
    uint32_t tmem_base;
};

struct BwdWsArgs {
    const float* Pi; const float* Pj;
    const float* e_prev; int De;
    const float* W1; int ldW1, coff;
    const float* W2; const float* b2;
    const int32_t* src; const int32_t* dst;
    int32_t N, E;
    const float* ca; const float* cb;
    const float* dagg;
    const int32_t* amin; const int32_t* amax;
    const float* m; const float* de;          // e_l, d e_l (L2 prefetch; the tiles themselves move by TMA)
    float* dPi; float* dPj;
    float* gW1; float* gW2; float* gb2;
    int no_de;                                // block 5: there is no d e_l (dm = ca + cb e_l + arg terms, applied here)
    int pf;                                   // L2 prefetch distance in tiles (0 = off)
    int dbg;                                  // timing experiments only (GCRL_WS_DBG): 1 no reds, 2 no gather, 4 no store, 8 single-pass bf16 weight-gradient products, 16 no dPj reds, 32 no dPi reds
    long long* trace;                         // debugging: clock64 per (role, tile, event) of CTA 0, or null
};

// pipeline timeline of CTA 0 (GCRL_WS_TRACE=<file>): roles 0 TMA-E, 1 MMA, 2 TMA-MD, 3 WG_A, 4 WG_B, 5 WG_C, 6 WG_D
#ifdef GCRL_WS_TRACE_BUILD
#define WS_TR(on, role, tile, ev)                                                           \
    do {                                                                                    \
        if (a.trace && (on) && blockIdx.x == 0 && (tile) < 64)                              \
            a.trace[(((role) * 64 + (tile)) * 8) + (ev)] = clock64();                       \
    } while (0)
#else
#define WS_TR(on, role, tile, ev) do { } while (0)
#endif

template <bool FIRST, bool LAST, bool X3>
__global__ void __launch_bounds__(WS_THREADS, 1) k_edge_bwd_ws(const __grid_constant__ CUtensorMap tm_e,
                                                              const __grid_constant__ CUtensorMap tm_m,
                                                              const __grid_constant__ CUtensorMap tm_de,
                                                              const __grid_constant__ CUtensorMap tm_out,
                                                              const BwdWsArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    BwdWsSmem& S = *reinterpret_cast<BwdWsSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = warp & 3;                     // TMEM sub-partition of this warp
    const long long t_start = clock64();
    const int wtid = tid & 127;                 // thread index inside its warp group

    if (tid == 0) {
        for (int i = 0; i < WS_NE; ++i) { mbar_init(&S.bar_Efull[i], 1); mbar_init(&S.bar_Eempty[i], 4); }
        for (int i = 0; i < WS_NMD; ++i) { mbar_init(&S.bar_MDfull[i], 1); mbar_init(&S.bar_MDempty[i], 4); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&S.bar_B1full[i], 4);
            mbar_init(&S.bar_B1free[i], 1);
            mbar_init(&S.bar_accBfree[i], LAST ? 4 : 8);      // epilogue 4: WG_B (block 5) or the two WG_C groups
            mbar_init(&S.bar_acc2full[i], 1);
            mbar_init(&S.bar_acc6full[i], 1);
            mbar_init(&S.bar_OUTfull[i], LAST ? 4 : 8);
        }
        for (int i = 0; i < 3; ++i) { mbar_init(&S.bar_accAfree[i], 12); mbar_init(&S.bar_acc1full[i], 1); mbar_init(&S.bar_acc4full[i], 1); }
        mbar_init(&S.bar_B2full, 8);
        mbar_init(&S.bar_B3full, 4);
        mbar_init(&S.bar_dz1full, 8);
        mbar_init(&S.bar_B2read, 4);
        mbar_init(&S.bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(&S.tmem_base);
    load_weight_tile(a.W2, 64, 0, S.W2[0], S.W2[1], tid, WS_THREADS);
    if (!FIRST) load_weight_tile(a.W1, a.ldW1, a.coff, S.Wc[0], S.Wc[1], tid, WS_THREADS);
    float (*Cf)[64] = reinterpret_cast<float (*)[64]>(&S.Wc[0][0]);   // FIRST: C^T fp32 [d][o]
    if (FIRST) {
        for (int t = tid; t < a.De * 64; t += WS_THREADS) Cf[t >> 6][t & 63] = a.W1[(t & 63) * a.ldW1 + a.coff + (t >> 6)];
        for (int t = tid; t < 2 * 32768 / 16; t += WS_THREADS)   // zero-padded raw edge-feature operand tiles
            reinterpret_cast<uint4*>(&S.B1[0][0])[t] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid < 64) { S.b2[tid] = a.b2[tid]; S.bsum[tid] = 0.f; }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem_base;
    const uint32_t t_accA0 = tmem, t_accB0 = tmem + 192, t_w2 = tmem + 320, t_c = tmem + 384, t_hm = tmem + 448;
    const uint32_t lane_addr = (uint32_t)(32 * q) << 16;

    const int64_t n_tiles = ((int64_t)a.E + TILE - 1) / TILE;
    // contiguous tile ranges: a CTA stays inside one or two graphs, so CTAs do not fight over the same dPi / dPj
    // rows (reds to one address serialise in L2) and its Pi / Pj gathers stay in L1 / L2
    const int64_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t tile_first = (int64_t)blockIdx.x * per_cta;
    const int n_cta = (int)min(per_cta, n_tiles - tile_first);                      // >= 1 (host sizes the grid)
    auto tile_t0 = [&](int i) { return (int32_t)((tile_first + i) * TILE); };

    if (warp == 0) {
        // ================= TMA producer: e_{l-1} boxes =================================================
        if (!FIRST && lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            auto prefetch = [&](int i) {
                if (i < n_cta) {
                    const int32_t t0 = tile_t0(i);
                    l2_prefetch_bulk(a.e_prev + (int64_t)t0 * 64, (uint32_t)min(TILE, a.E - t0) * 256u);
                }
            };
            for (int i = 1; i < a.pf; ++i) prefetch(i);
            int k = 0;
            for (int i = 0; i < n_cta; ++i) {
                const int32_t t0 = tile_t0(i);
                const int nrows = min(TILE, a.E - t0);
                if (a.pf > 0) prefetch(i + a.pf);
                for (int j = 0; 32 * j < nrows; ++j, ++k) {
                    const int slot = k % WS_NE, u = k / WS_NE;
                    mbar_wait(&S.bar_Eempty[slot], (u & 1) ^ 1);
                    WS_TR(true, 0, i, j);
                    mbar_arrive_expect_tx(&S.bar_Efull[slot], 8192);
                    tma_load_2d_hint(S.ringE[slot], &tm_e, 0, t0 + 32 * j, &S.bar_Efull[slot], pol);
                    tma_load_2d_hint(S.ringE[slot] + 4096, &tm_e, 32, t0 + 32 * j, &S.bar_Efull[slot], pol);
                }
            }
        }
    } else if (warp == 2) {
        // ================= TMA producer: (e_l, d e_l) box pairs =============================================
        if (!LAST && lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            auto prefetch = [&](int i) {
                if (i < n_cta) {
                    const int32_t t0 = tile_t0(i);
                    const uint32_t bytes = (uint32_t)min(TILE, a.E - t0) * 256u;
                    l2_prefetch_bulk(a.m + (int64_t)t0 * 64, bytes);
                    if (!a.no_de) l2_prefetch_bulk(a.de + (int64_t)t0 * 64, bytes);
                }
            };
            for (int i = 1; i < a.pf; ++i) prefetch(i);
            int k = 0;
            for (int i = 0; i < n_cta; ++i) {
                const int32_t t0 = tile_t0(i);
                const int nrows = min(TILE, a.E - t0);
                if (a.pf > 0) prefetch(i + a.pf);
                for (int j = 0; 32 * j < nrows; ++j, ++k) {
                    const int slot = k % WS_NMD, u = k / WS_NMD;
                    mbar_wait(&S.bar_MDempty[slot], (u & 1) ^ 1);
                    WS_TR(true, 2, i, j);
                    mbar_arrive_expect_tx(&S.bar_MDfull[slot], a.no_de ? 8192 : 16384);
                    tma_load_2d_hint(S.ringMD[slot], &tm_m, 0, t0 + 32 * j, &S.bar_MDfull[slot], pol);
                    tma_load_2d_hint(S.ringMD[slot] + 4096, &tm_m, 32, t0 + 32 * j, &S.bar_MDfull[slot], pol);
                    if (!a.no_de) {
                        tma_load_2d_hint(S.ringMD[slot] + 8192, &tm_de, 0, t0 + 32 * j, &S.bar_MDfull[slot], pol);
                        tma_load_2d_hint(S.ringMD[slot] + 12288, &tm_de, 32, t0 + 32 * j, &S.bar_MDfull[slot], pol);
                    }
                }
            }
        }
    } else if (warp == 3) {
        // ================= TMA store issuer: d e_{l-1} staging tiles -> global ========================================
        if (!FIRST && lane == 0) {
            const uint64_t pol = l2_policy_evict_first();
            for (int i = 0; i < n_cta; ++i) {
                const int s = i & 1, ph = (i >> 1) & 1;
                const float* out = reinterpret_cast<const float*>(S.B1[s]);
                mbar_wait(&S.bar_OUTfull[s], ph);
                if (!(a.dbg & 4)) {
                    tma_store_2d_hint(&tm_out, 0, tile_t0(i), out, pol);          // rows past E are clipped by the tensor map
                    tma_store_2d_hint(&tm_out, 32, tile_t0(i), out + 128 * 32, pol);
                }
                bulk_commit();
                bulk_wait_read0();
                WS_TR(true, 6, i, 2);
                mbar_arrive(&S.bar_B1free[s]);
            }
            bulk_wait0();
        }
    } else if (warp == 1) {
        // ================= MMA issuer: the whole warp walks the loop, one elected lane issues ================
        {
            const uint32_t id_feat = make_idesc_bf16(128, 64, false, false);
            const uint32_t id_featT = make_idesc_bf16(128, 64, false, true);
            const uint32_t id_rows = make_idesc_bf16(64, 64, true, true);
            const uint32_t wc_hi = smem_u32(S.Wc[0]), wc_lo = smem_u32(S.Wc[1]);
            const uint32_t w2_hi = smem_u32(S.W2[0]), w2_lo = smem_u32(S.W2[1]);
            const uint32_t b2_hi = smem_u32(S.B2), b2_lo = b2_hi + 16384;
            const uint32_t b3_hi = smem_u32(S.B3), b3_lo = b3_hi + 16384;
            auto g1_ready = [&](int j) {      // warp-uniform, non-blocking
                const uint32_t ok = mbar_test_wait(&S.bar_B1full[j & 1], (j >> 1) & 1) ? 1u : 0u;
                return __shfl_sync(0xffffffffu, ok, 0) != 0;
            };
            auto issue_g1 = [&](int j) {
                const int s = j & 1;
                mbar_wait(&S.bar_B1full[s], (j >> 1) & 1);
                tc_fence_after();
                WS_TR(lane == 0, 1, j, 0);
                if (!FIRST) {
                    const uint32_t b1_hi = smem_u32(S.B1[s]);
                    if (elect_one()) {
                        issue_gemm_feat<X3, true>(t_accA0 + 64 * (j % 3), b1_hi, b1_hi + 16384, wc_hi, wc_lo, id_feat, true);
                        umma_commit(&S.bar_acc1full[j % 3]);
                    }
                    __syncwarp();
                }
            };
            issue_g1(0);
            for (int i = 0; i < n_cta; ++i) {
                const int s = i & 1, ph = (i >> 1) & 1;
                const uint32_t b1_hi = smem_u32(S.B1[s]);
                mbar_wait(&S.bar_B2full, i & 1);
                if (LAST) {
                    mbar_wait(&S.bar_accBfree[s], ph ^ 1);
                    tc_fence_after();
                    if (elect_one()) {
                        issue_gemm_feat<X3, true>(t_accB0 + 64 * s, b2_hi, b2_lo, w2_hi, w2_lo, id_feat, false);
                        umma_commit(&S.bar_acc2full[s]);
                    }
                    __syncwarp();
                }
                mbar_wait(&S.bar_B3full, i & 1);
                tc_fence_after();
                WS_TR(lane == 0, 1, i, 1);
                if (elect_one()) {
                    if (a.dbg & 8) issue_gemm_rows<false, true>(t_w2, b3_hi, b3_lo, b2_hi, b2_lo, id_rows, 128, i > 0);
                    else issue_gemm_rows<X3, true>(t_w2, b3_hi, b3_lo, b2_hi, b2_lo, id_rows, 128, i > 0);
                    issue_gemm_featT<X3, true>(t_accA0 + 64 * (i % 3), b3_hi, b3_lo, w2_hi, w2_lo, id_featT, false);
                    umma_commit(&S.bar_acc4full[i % 3]);
                }
                __syncwarp();
                WS_TR(lane == 0, 1, i, 2);
                // G1 of the next tile goes in front of G5/G6 only if its operands are already there
                bool g1_next = (i + 1 >= n_cta);
                if (!g1_next && g1_ready(i + 1)) { issue_g1(i + 1); g1_next = true; }
                mbar_wait(&S.bar_dz1full, i & 1);
                if (!FIRST && !LAST) mbar_wait(&S.bar_accBfree[s], ph ^ 1);
                tc_fence_after();
                WS_TR(lane == 0, 1, i, 3);
                if (elect_one()) {
                    if (a.dbg & 8) issue_gemm_rows<false, true>(t_c, b2_hi, b2_lo, b1_hi, b1_hi + 16384, id_rows, 128, i > 0);
                    else issue_gemm_rows<X3, true>(t_c, b2_hi, b2_lo, b1_hi, b1_hi + 16384, id_rows, 128, i > 0);
                    if (!FIRST) issue_gemm_featT<X3, true>(t_accB0 + 64 * s, b2_hi, b2_lo, wc_hi, wc_lo, id_featT, false);
                    umma_commit(&S.bar_acc6full[s]);
                    if (FIRST) umma_commit(&S.bar_B1free[s]);
                }
                __syncwarp();
                WS_TR(lane == 0, 1, i, 4);
                if (!g1_next) issue_g1(i + 1);
            }
            if (elect_one()) umma_commit(&S.bar_done);
            __syncwarp();
        }
    } else if (warp >= 4 && warp < 8) {
        // ================= WG_A: e_{l-1} operand tile + Pi/Pj into TMEM =======================================
        int k = 0;
        const int cp = 2 * (lane & 3);
        // endpoint ids of this thread's 4 fragment rows (2 groups x rows a, b), fetched one tile ahead
        int32_t nd[4], ns[4];
        auto load_ids = [&](int i) {
#pragma unroll
            for (int t = 0; t < 4; ++t) { nd[t] = -1; ns[t] = 0; }
            if (i < n_cta) {
                const int32_t t0 = tile_t0(i);
                const int nrows = min(TILE, a.E - t0);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int r = 32 * q + 16 * (t >> 1) + 8 * (t & 1) + (lane >> 2);
                    if (r < nrows) { nd[t] = __ldg(a.dst + t0 + r); ns[t] = __ldg(a.src + t0 + r); }
                }
            }
        };
        load_ids(0);
        int32_t cur_d = -1;              // destination whose Pi row is cached in piv
        float2 piv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) piv[c] = make_float2(0.f, 0.f);
        for (int i = 0; i < n_cta; ++i) {
            const int s = i & 1, ph = (i >> 1) & 1;
            const int32_t t0 = tile_t0(i);
            const int nrows = min(TILE, a.E - t0);
            unsigned char* b1 = S.B1[s];
            int32_t d4[4], s4[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) { d4[t] = nd[t]; s4[t] = ns[t]; }
            load_ids(i + 1);
            mbar_wait(&S.bar_accAfree[i % 3], ((i / 3) & 1) ^ 1);
            tc_fence_after();
            WS_TR(wtid == 0, 3, i, 0);
#pragma unroll 1
            for (int g = 0; g < 2; ++g) {
                const int ra = 32 * q + 16 * g + (lane >> 2), rb = ra + 8;
                const int32_t da = g ? d4[2] : d4[0], db = g ? d4[3] : d4[1];
                const bool va = da >= 0 && !(a.dbg & 2), vb = db >= 0 && !(a.dbg & 2);
                const float2* pja = reinterpret_cast<const float2*>(a.Pj + (int64_t)(g ? s4[2] : s4[0]) * 64 + cp);
                const float2* pjb = reinterpret_cast<const float2*>(a.Pj + (int64_t)(g ? s4[3] : s4[1]) * 64 + cp);
                float2 ya[8], yb[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    ya[c] = va ? __ldg(pja + 4 * c) : make_float2(0.f, 0.f);
                    yb[c] = vb ? __ldg(pjb + 4 * c) : make_float2(0.f, 0.f);
                }
                if (va && da != cur_d) {          // new segment: refresh the cached Pi row
                    cur_d = da;
                    const float2* pia = reinterpret_cast<const float2*>(a.Pi + (int64_t)da * 64 + cp);
#pragma unroll
                    for (int c = 0; c < 8; ++c) piv[c] = __ldg(pia + 4 * c);
                }
                const bool b_other = vb && db != cur_d;
                const float2* pib = reinterpret_cast<const float2*>(a.Pi + (int64_t)(vb ? db : 0) * 64 + cp);
                float v[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float2 xb = b_other ? __ldg(pib + 4 * c) : piv[c];
                    v[4 * c + 0] = va ? piv[c].x + ya[c].x : 0.f;
                    v[4 * c + 1] = va ? piv[c].y + ya[c].y : 0.f;
                    v[4 * c + 2] = vb ? xb.x + yb[c].x : 0.f;
                    v[4 * c + 3] = vb ? xb.y + yb[c].y : 0.f;
                }
                if (FIRST) {
                    // block 1: C e is De FMAs per element, folded in here (same association as the
                    // other paths: (Pi + Pj) + C e)
                    float ca_[32];
#pragma unroll
                    for (int c = 0; c < 32; ++c) ca_[c] = 0.f;
                    for (int d = 0; d < a.De; ++d) {
                        const float ea = va ? a.e_prev[(int64_t)(t0 + ra) * a.De + d] : 0.f;
                        const float eb = vb ? a.e_prev[(int64_t)(t0 + rb) * a.De + d] : 0.f;
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float2 cc = *reinterpret_cast<const float2*>(&Cf[d][8 * c + cp]);
                            ca_[4 * c + 0] = fmaf(ea, cc.x, ca_[4 * c + 0]);
                            ca_[4 * c + 1] = fmaf(ea, cc.y, ca_[4 * c + 1]);
                            ca_[4 * c + 2] = fmaf(eb, cc.x, ca_[4 * c + 2]);
                            ca_[4 * c + 3] = fmaf(eb, cc.y, ca_[4 * c + 3]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) v[c] += ca_[c];
                }
                tmem_st_16x256b_x8(t_accA0 + 64 * (i % 3) + ((uint32_t)(32 * q + 16 * g) << 16), v);
            }
            tmem_st_wait();
            mbar_wait(&S.bar_B1free[s], ph ^ 1);        // the staging use of B1[s] two tiles ago has been stored
            WS_TR(wtid == 0, 3, i, 1);
            if (!FIRST) {
                for (int j = 0; j < 4; ++j) {
                    if (32 * j < nrows) {
                        const int slot = k % WS_NE, u = k / WS_NE;
                        ++k;
                        mbar_wait(&S.bar_Efull[slot], u & 1);
                        WS_TR(wtid == 0, 3, i, 3 + j);
                        const unsigned char* box = S.ringE[slot];
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int r = (wtid >> 4) + 8 * it, c4 = wtid & 15;
                            const float4 x = *reinterpret_cast<const float4*>(box + sw_off<32>(r, c4));
                            uint2 h, l;
                            split4<X3>(x, h, l);
                            const uint32_t off = bf_off(32 * j + r, c4 >> 1) + (c4 & 1) * 8;
                            *reinterpret_cast<uint2*>(b1 + off) = h;
                            if (X3) *reinterpret_cast<uint2*>(b1 + 16384 + off) = l;
                        }
                        warp_arrive(&S.bar_Eempty[slot], lane);
                    } else {
                        // rows past the end of the edge list: zero operand rows
                        for (int t = wtid; t < 32 * 8; t += 128) {
                            const uint32_t off = bf_off(32 * j + (t >> 3), t & 7);
                            *reinterpret_cast<uint4*>(b1 + off) = make_uint4(0u, 0u, 0u, 0u);
                            if (X3) *reinterpret_cast<uint4*>(b1 + 16384 + off) = make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                }
            } else if (wtid < TILE) {
                // zero-padded raw edge features: chunk 0 of the row holds e[0..De)
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = (wtid < nrows && j < a.De) ? a.e_prev[(int64_t)(t0 + wtid) * a.De + j] : 0.f;
                store_split8<X3>(x, reinterpret_cast<__nv_bfloat16*>(b1), reinterpret_cast<__nv_bfloat16*>(b1 + 16384), wtid, 0);
            }
            fence_proxy_async();
            tc_fence_before();
            WS_TR(wtid == 0, 3, i, 2);
            warp_arrive(&S.bar_B1full[s], lane);
        }
    } else if (warp >= 8 && warp < 12) {
        // ================= WG_B: dm operand tile; epilogue 4 (d e_{l-1} = acc6 -> fp32 staging in the dead B1[s]) ====
        unsigned char* b3 = S.B3;
        auto epilogue4 = [&](int i) {
            if (FIRST || !LAST) return;       // blocks 1-4: done by the WG_C groups
            const int s = i & 1, ph = (i >> 1) & 1;
            const int row = 32 * q + lane;
            float* out = reinterpret_cast<float*>(S.B1[s]);
            mbar_wait(&S.bar_acc6full[s], ph);      // G5 has finished with B1[s], G6 wrote ACC_B[s]
            tc_fence_after();
            WS_TR(wtid == 0, 6, i, 0);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float acc[32];
                tmem_ld32(t_accB0 + 64 * s + lane_addr + 32 * half, acc);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4*>(l_ptr(out, row, 8 * half + c)) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
            }
            tc_fence_before();
            fence_proxy_async();
            WS_TR(wtid == 0, 6, i, 1);
            warp_arrive(&S.bar_accBfree[s], lane);
            warp_arrive(&S.bar_OUTfull[s], lane);     // warp 3 issues the TMA store
        };
        if (!LAST) {
            const int c4 = wtid & 15;
            float bs0 = 0.f, bs1 = 0.f, bs2 = 0.f, bs3 = 0.f;
            int32_t cur_v = -1;
            float4 A4 = make_float4(0.f, 0.f, 0.f, 0.f), B4 = A4, Gmn = A4, Gmx = A4;
            int4 Imn = make_int4(-1, -1, -1, -1), Imx = Imn;       // no_de: arg slots / gradients of this thread's 4 columns
            int k = 0;
            for (int i = 0; i < n_cta; ++i) {
              {
                const int32_t t0 = tile_t0(i);
                const int nrows = min(TILE, a.E - t0);
                if (i > 0) mbar_wait(&S.bar_acc4full[(i - 1) % 3], ((i - 1) / 3) & 1);    // G3/G4 of the previous tile read B3
                WS_TR(wtid == 0, 4, i, 0);
#pragma unroll 1
                for (int j = 0; j < 4; ++j) {
                    if (32 * j < nrows) {
                        const int slot = k % WS_NMD, u = k / WS_NMD;
                        ++k;
                        const unsigned char* mbox = S.ringMD[slot];
                        const unsigned char* dbox = mbox + 8192;
                        mbar_wait(&S.bar_MDfull[slot], u & 1);
                        WS_TR(wtid == 0, 4, i, 1 + j);
#pragma unroll
                        for (int it = 0; it < 4; ++it) {
                            const int r = (wtid >> 4) + 8 * it;
                            const int R = 32 * j + r;
                            const int32_t v = R < nrows ? __ldg(a.dst + t0 + R) : -1;
                            float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (v >= 0) {
                                if (v != cur_v) {
                                    cur_v = v;
                                    A4 = __ldg(reinterpret_cast<const float4*>(a.ca + (int64_t)v * 64) + c4);
                                    B4 = __ldg(reinterpret_cast<const float4*>(a.cb + (int64_t)v * 64) + c4);
                                    if (a.no_de) {
                                        Imn = __ldg(reinterpret_cast<const int4*>(a.amin + (int64_t)v * 64) + c4);
                                        Imx = __ldg(reinterpret_cast<const int4*>(a.amax + (int64_t)v * 64) + c4);
                                        Gmn = __ldg(reinterpret_cast<const float4*>(a.dagg + (int64_t)v * 256 + 64) + c4);
                                        Gmx = __ldg(reinterpret_cast<const float4*>(a.dagg + (int64_t)v * 256 + 128) + c4);
                                    }
                                }
                                const float4 m4 = *reinterpret_cast<const float4*>(mbox + sw_off<32>(r, c4));
                                float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (!a.no_de) d4 = *reinterpret_cast<const float4*>(dbox + sw_off<32>(r, c4));
                                r4.x = fmaf(B4.x, m4.x, A4.x) + d4.x; r4.y = fmaf(B4.y, m4.y, A4.y) + d4.y;
                                r4.z = fmaf(B4.z, m4.z, A4.z) + d4.z; r4.w = fmaf(B4.w, m4.w, A4.w) + d4.w;
                                if (a.no_de) {
                                    const int32_t slot_e = t0 + R;
                                    if (Imn.x == slot_e) r4.x += Gmn.x;
                                    if (Imn.y == slot_e) r4.y += Gmn.y;
                                    if (Imn.z == slot_e) r4.z += Gmn.z;
                                    if (Imn.w == slot_e) r4.w += Gmn.w;
                                    if (Imx.x == slot_e) r4.x += Gmx.x;
                                    if (Imx.y == slot_e) r4.y += Gmx.y;
                                    if (Imx.z == slot_e) r4.z += Gmx.z;
                                    if (Imx.w == slot_e) r4.w += Gmx.w;
                                }
                            }
                            bs0 += r4.x; bs1 += r4.y; bs2 += r4.z; bs3 += r4.w;
                            uint2 h, l;
                            split4<X3>(r4, h, l);
                            const uint32_t off = bf_off(R, c4 >> 1) + (c4 & 1) * 8;
                            *reinterpret_cast<uint2*>(b3 + off) = h;
                            if (X3) *reinterpret_cast<uint2*>(b3 + 16384 + off) = l;
                        }
                        warp_arrive(&S.bar_MDempty[slot], lane);
                    } else {
                        for (int t = wtid; t < 32 * 8; t += 128) {
                            const uint32_t off = bf_off(32 * j + (t >> 3), t & 7);
                            *reinterpret_cast<uint4*>(b3 + off) = make_uint4(0u, 0u, 0u, 0u);
                            if (X3) *reinterpret_cast<uint4*>(b3 + 16384 + off) = make_uint4(0u, 0u, 0u, 0u);
                        }
                    }
                }
                fence_proxy_async();
                WS_TR(wtid == 0, 4, i, 5);
                warp_arrive(&S.bar_B3full, lane);
              }
            }
            atomicAdd(&S.bsum[4 * c4 + 0], bs0); atomicAdd(&S.bsum[4 * c4 + 1], bs1);
            atomicAdd(&S.bsum[4 * c4 + 2], bs2); atomicAdd(&S.bsum[4 * c4 + 3], bs3);
        } else {
            // block 5: m = e_5 recomputed (G2), no d e_5:  dm = ca + cb * m + arg terms, thread = tile row
            const int row = 32 * q + lane;
            for (int i = 0; i < n_cta; ++i) {
                const int s = i & 1, ph = (i >> 1) & 1;
                const int32_t t0 = tile_t0(i);
                const int nrows = min(TILE, a.E - t0);
                const bool valid = row < nrows;
                const int32_t v = valid ? __ldg(a.dst + t0 + row) : 0;
                const int32_t slot_e = t0 + row;
                if (i > 0) epilogue4(i - 1);
                mbar_wait(&S.bar_acc2full[s], ph);
                if (i > 0) mbar_wait(&S.bar_acc4full[(i - 1) % 3], ((i - 1) / 3) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    float acc[32];
                    tmem_ld32(t_accB0 + 64 * s + lane_addr + 32 * half, acc);
                    const float4* ca4 = reinterpret_cast<const float4*>(a.ca + (int64_t)v * 64 + 32 * half);
                    const float4* cb4 = reinterpret_cast<const float4*>(a.cb + (int64_t)v * 64 + 32 * half);
                    const float4* gmn4 = reinterpret_cast<const float4*>(a.dagg + (int64_t)v * 256 + 64 + 32 * half);
                    const float4* gmx4 = reinterpret_cast<const float4*>(a.dagg + (int64_t)v * 256 + 128 + 32 * half);
                    const int4* imn4 = reinterpret_cast<const int4*>(a.amin + (int64_t)v * 64 + 32 * half);
                    const int4* imx4 = reinterpret_cast<const int4*>(a.amax + (int64_t)v * 64 + 32 * half);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 bb = *reinterpret_cast<const float4*>(&S.b2[32 * half + 4 * c]);
                        float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid) {
                            const float4 A4 = __ldg(ca4 + c), B4 = __ldg(cb4 + c);
                            const float4 gmn = __ldg(gmn4 + c), gmx = __ldg(gmx4 + c);
                            const int4 imn = __ldg(imn4 + c), imx = __ldg(imx4 + c);
                            r4.x = fmaf(B4.x, acc[4 * c + 0] + bb.x, A4.x); r4.y = fmaf(B4.y, acc[4 * c + 1] + bb.y, A4.y);
                            r4.z = fmaf(B4.z, acc[4 * c + 2] + bb.z, A4.z); r4.w = fmaf(B4.w, acc[4 * c + 3] + bb.w, A4.w);
                            if (imn.x == slot_e) r4.x += gmn.x;
                            if (imn.y == slot_e) r4.y += gmn.y;
                            if (imn.z == slot_e) r4.z += gmn.z;
                            if (imn.w == slot_e) r4.w += gmn.w;
                            if (imx.x == slot_e) r4.x += gmx.x;
                            if (imx.y == slot_e) r4.y += gmx.y;
                            if (imx.z == slot_e) r4.z += gmx.z;
                            if (imx.w == slot_e) r4.w += gmx.w;
                        }
                        acc[4 * c + 0] = r4.x; acc[4 * c + 1] = r4.y; acc[4 * c + 2] = r4.z; acc[4 * c + 3] = r4.w;
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        store_split8<X3>(&acc[8 * c], reinterpret_cast<__nv_bfloat16*>(b3), reinterpret_cast<__nv_bfloat16*>(b3 + 16384), row, 4 * half + c);
                }
                fence_proxy_async();
                tc_fence_before();
                warp_arrive(&S.bar_B3full, lane);
            }
            epilogue4(n_cta - 1);
            // (db2 of block 5 = sum_v (g_mean + g_min + g_max): computed per vertex by k_pna_bwd_coef)
        }
        named_sync(1, 128);
        if (wtid < 64) atomicAdd(&a.gb2[wtid], S.bsum[wtid]);
    } else if (warp >= 12 && warp < 20) {
        // ================= WG_C0 / WG_C1: epilogues 1 and 3, one warp group per 32-column half ===================
        const int row = 32 * q + lane;
        const int half = warp >= 16 ? 1 : 0;
        const bool tr = (tid == 12 * 32);
        __nv_bfloat16* b2h = reinterpret_cast<__nv_bfloat16*>(S.B2);
        __nv_bfloat16* b2l = reinterpret_cast<__nv_bfloat16*>(S.B2 + 16384);
        // epilogue 4 of tile j, this group's 32 columns: d e_{l-1} = acc6 -> fp32 staging slab in the dead B1[j & 1]
        auto epilogue4_half = [&](int j) {
            if (FIRST || LAST) return;
            const int sj = j & 1;
            float* out = reinterpret_cast<float*>(S.B1[sj]);
            mbar_wait(&S.bar_acc6full[sj], (j >> 1) & 1);      // G5 has finished with B1[sj], G6 wrote ACC_B[sj]
            tc_fence_after();
            float acc[32];
            tmem_ld32(t_accB0 + 64 * sj + lane_addr + 32 * half, acc);
#pragma unroll
            for (int c = 0; c < 8; ++c)
                *reinterpret_cast<float4*>(l_ptr(out, row, 8 * half + c)) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
            tc_fence_before();
            fence_proxy_async();
            warp_arrive(&S.bar_accBfree[sj], lane);
            warp_arrive(&S.bar_OUTfull[sj], lane);     // warp 3 issues the TMA store
        };
        for (int i = 0; i < n_cta; ++i) {
            const int s = i & 1, ph = (i >> 1) & 1;
            const int a3 = i % 3, pa = (i / 3) & 1;
            const uint32_t t_acc = t_accA0 + 64 * a3 + lane_addr + 32 * half;
            uint32_t hm = 0u;
            // ---- epilogue 1: hid = relu(acc1) ---------------------------------------------------------------------
            if (FIRST) mbar_wait(&S.bar_B1full[s], ph);
            else mbar_wait(&S.bar_acc1full[a3], pa);
            tc_fence_after();
            WS_TR(tr, 5, i, 0);
            {
                float acc[32];
                tmem_ld32(t_acc, acc);
                if (i > 0) mbar_wait(&S.bar_acc6full[s ^ 1], ((i - 1) >> 1) & 1);   // G5/G6 of the previous tile read B2
                WS_TR(tr, 5, i, 1);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        x[j] = fmaxf(acc[8 * c + j], 0.f);
                        hm |= (acc[8 * c + j] > 0.f ? 1u : 0u) << (8 * c + j);
                    }
                    store_split8<X3>(x, b2h, b2l, row, 4 * half + c);
                }
            }
            tmem_st_x1(t_hm + 2 * a3 + half + lane_addr, hm);      // relu mask of this row / column half, for WG_D
            tmem_st_wait();
            fence_proxy_async();
            tc_fence_before();
            WS_TR(tr, 5, i, 2);
            warp_arrive(&S.bar_B2full, lane);
            if (i > 0) epilogue4_half(i - 1);          // while G3/G4 of this tile run
            // ---- epilogue 3: dz1 = d hid * relu' -> B2 (G3/G4 have finished with hid) -----------------------------------
            mbar_wait(&S.bar_acc4full[a3], pa);
            tc_fence_after();
            WS_TR(tr, 5, i, 3);
            {
                float acc[32];
                tmem_ld32(t_acc, acc);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float x[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) x[j] = ((hm >> (8 * c + j)) & 1u) ? acc[8 * c + j] : 0.f;
                    store_split8<X3>(x, b2h, b2l, row, 4 * half + c);
                }
            }
            fence_proxy_async();
            tc_fence_before();
            WS_TR(tr, 5, i, 4);
            warp_arrive(&S.bar_dz1full, lane);
            warp_arrive(&S.bar_accAfree[a3], lane);
        }
        epilogue4_half(n_cta - 1);
        // ---- weight gradients: TMEM -> global (one atomic pass per CTA) ----------------------------------------------------
        mbar_wait(&S.bar_done, 0);
        tc_fence_after();
        {
            const int o = 16 * q + lane;          // accumulator row of an M = 64 product (lanes 0..15 of each sub-partition)
            const int hh = half;
            {
                float v[32];
                tmem_ld32(t_w2 + lane_addr + 32 * hh, v);
                if (lane < 16)
#pragma unroll
                    for (int j = 0; j < 32; j += 4) red_add_v4(&a.gW2[o * 64 + 32 * hh + j], v[j], v[j + 1], v[j + 2], v[j + 3]);
                tmem_ld32(t_c + lane_addr + 32 * hh, v);
                if (lane < 16) {
                    if (!FIRST) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(&a.gW1[o * a.ldW1 + a.coff + 32 * hh + j], v[j]);
                    } else if (hh == 0) {
                        for (int j = 0; j < a.De; ++j) atomicAdd(&a.gW1[o * a.ldW1 + a.coff + j], v[j]);
                    }
                }
            }
        }
        tc_fence_before();
    } else if (warp >= 20) {
        // ================= WG_D: dPj / dPi reductions, then epilogue 4 (d e_{l-1} -> staging -> TMA store) =========
        {
            const int row = 32 * q + lane;
            for (int i = 0; i < n_cta; ++i) {
                const int s = i & 1, ph = (i >> 1) & 1;
                const int32_t t0 = tile_t0(i);
                const int nrows = min(TILE, a.E - t0);
                // ---- dPj[src] += dz1 (vector reds, one 128-byte line per lane), dPi[dst] += segment sums (warp
                // transpose-reduce): dz1 = acc4 * relu' re-read from TMEM, off the B2 critical path
                {
                    const bool valid = row < nrows;
                    const int32_t my_src = valid ? __ldg(a.src + t0 + row) : 0;
                    const int32_t my_dst = valid ? __ldg(a.dst + t0 + row) : -1;
                    const int32_t dprev = __shfl_up_sync(0xffffffffu, my_dst, 1);
                    const uint32_t sbits = __ballot_sync(0xffffffffu, valid && lane > 0 && my_dst != dprev);
                    const int nb = __popc(sbits);
                    const int32_t d0 = __shfl_sync(0xffffffffu, my_dst, 0);
                    const int bpos = nb ? (__ffs(sbits) - 1) : 0;
                    const int32_t d1 = __shfl_sync(0xffffffffu, my_dst, bpos);
                    const int a3 = i % 3;
                    const uint32_t t_acc = t_accA0 + 64 * a3 + lane_addr;
                    mbar_wait(&S.bar_acc4full[a3], (i / 3) & 1);
                    tc_fence_after();
                    WS_TR(wtid == 0, 6, i, 3);
                    uint32_t hm0, hm1;
                    tmem_ld_x2(t_hm + 2 * a3 + lane_addr, hm0, hm1);
#pragma unroll 1
                    for (int hf = 0; hf < 2; ++hf) {
                        const uint32_t hm = hf ? hm1 : hm0;
                        float acc[32];
                        tmem_ld32(t_acc + 32 * hf, acc);
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[j] = ((hm >> j) & 1u) ? acc[j] : 0.f;
                        if (a.dbg & 1) continue;
                        if (valid && !(a.dbg & 16)) {
                            float* pj = a.dPj + (int64_t)my_src * 64 + 32 * hf;
#pragma unroll
                            for (int c = 0; c < 8; ++c) red_add_v4(pj + 4 * c, acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                        }
                        if (a.dbg & 32) continue;
                        if (nb == 0) {
                            if (d0 >= 0) {
                                const float sum = warp_colsum32(acc, lane);
                                atomicAdd(a.dPi + (int64_t)d0 * 64 + 32 * hf + lane, sum);
                            }
                        } else if (nb == 1) {
                            // one boundary inside the warp's 32 rows: two masked column sums instead of 32 lanes
                            // hammering two dPi rows with reds (same-address atomics serialise in L2)
                            float t[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) t[j] = lane < bpos ? acc[j] : 0.f;
                            const float sa = warp_colsum32(t, lane);
                            atomicAdd(a.dPi + (int64_t)d0 * 64 + 32 * hf + lane, sa);
#pragma unroll
                            for (int j = 0; j < 32; ++j) acc[j] = (lane >= bpos && valid) ? acc[j] : 0.f;
                            const float sb = warp_colsum32(acc, lane);
                            atomicAdd(a.dPi + (int64_t)d1 * 64 + 32 * hf + lane, sb);
                        } else if (valid) {
                            float* pi = a.dPi + (int64_t)my_dst * 64 + 32 * hf;
#pragma unroll
                            for (int c = 0; c < 8; ++c) red_add_v4(pi + 4 * c, acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
                        }
                    }
                    tc_fence_before();
                    WS_TR(wtid == 0, 6, i, 4);
                    warp_arrive(&S.bar_accAfree[a3], lane);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (a.trace && tid == 0) {      // per-CTA busy time and SM id (slots after the 7 x 64 x 8 timeline)
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.trace[7 * 64 * 8 + blockIdx.x] = ((long long)smid << 48) | (clock64() - t_start);
    }
    if (warp == 2) tmem_dealloc<512>(tmem);
}

template <bool FIRST, bool LAST, bool X3>
static int launch_ws_variant(const CUtensorMap& te, const CUtensorMap& tm, const CUtensorMap& tde, const CUtensorMap& tout,
                             const BwdWsArgs& a, int grid, cudaStream_t st) {
    const int smem = (int)sizeof(BwdWsSmem) + 1024;
    static bool done = false;
    if (!done) {
        GCRL_CUDA(cudaFuncSetAttribute(k_edge_bwd_ws<FIRST, LAST, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        done = true;
    }
    static int n_launch = 0;
    const char* tr = getenv("GCRL_WS_TRACE");
    BwdWsArgs b = a;
    long long* d_tr = nullptr;
    const size_t tr_n = 7 * 64 * 8 + 160;
    if (tr && !FIRST && !LAST && ++n_launch == 3) {     // one warm launch of the blocks-2..4 variant
        GCRL_CUDA(cudaMalloc(&d_tr, tr_n * sizeof(long long)));
        GCRL_CUDA(cudaMemsetAsync(d_tr, 0, tr_n * sizeof(long long), st));
        b.trace = d_tr;
    }
    k_edge_bwd_ws<FIRST, LAST, X3><<<grid, WS_THREADS, smem, st>>>(te, tm, tde, tout, b);
    if (d_tr) {
        long long* h = (long long*)malloc(tr_n * sizeof(long long));
        GCRL_CUDA(cudaStreamSynchronize(st));
        GCRL_CUDA(cudaMemcpy(h, d_tr, tr_n * sizeof(long long), cudaMemcpyDeviceToHost));
        FILE* f = fopen(tr, "w");
        if (f) {
            for (size_t i = 0; i < tr_n; ++i)
                if (h[i] && i < 7 * 64 * 8) fprintf(f, "%d %d %d %lld\n", (int)(i / 512), (int)((i / 8) % 64), (int)(i % 8), h[i]);
            for (size_t i = 7 * 64 * 8; i < tr_n; ++i)
                if (h[i]) fprintf(f, "cta %d sm %d cycles %lld\n", (int)(i - 7 * 64 * 8), (int)(h[i] >> 48), h[i] & 0xffffffffffffll);
            fclose(f);
        }
        free(h);
        cudaFree(d_tr);
    }
    return GCRL_OK;
}

bool use_ws_kernels() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GCRL_EDGE_IMPL");
        v = (e && e[0] == 't' && e[1] == 'c') ? 0 : 1;
    }
    return v == 1;
}

// Warp-specialised tensor-core edge backward of one block; same contract as launch_edge_bwd_tc.
int launch_edge_bwd_ws(bool first, bool x3, const float* Pi, const float* Pj, const float* e_prev, int De,
                       const float* W1, int ldW1, int coff, const float* W2, const float* b2, const int32_t* seg_off,
                       const int32_t* src, const int32_t* dst, int64_t N, int64_t E, const float* m, const float* de,
                       const float* ca, const float* cb, const float* dagg, const int32_t* amin, const int32_t* amax,
                       float* de_prev, float* dPi, float* dPj, float* gW1, float* gW2, float* gb2, cudaStream_t st) {
    (void)seg_off;
    if (E == 0) return GCRL_OK;
    const bool last = (m == nullptr);          // recompute flavour (unused by the default dispatch)
    const bool no_de = (m != nullptr && de == nullptr);
    BwdWsArgs a;
    a.Pi = Pi; a.Pj = Pj; a.e_prev = e_prev; a.De = De;
    a.W1 = W1; a.ldW1 = ldW1; a.coff = coff; a.W2 = W2; a.b2 = b2;
    a.src = src; a.dst = dst; a.N = (int32_t)N; a.E = (int32_t)E;
    a.ca = ca; a.cb = cb; a.dagg = dagg; a.amin = amin; a.amax = amax;
    a.m = m; a.de = de; a.trace = nullptr; a.no_de = no_de ? 1 : 0;
    {
        static int pf = -1;
        if (pf < 0) { const char* e = getenv("GCRL_WS_PF"); pf = e ? atoi(e) : 1; }
        a.pf = pf;
        static int dbg = -1;
        if (dbg < 0) { const char* e = getenv("GCRL_WS_DBG"); dbg = e ? atoi(e) : 0; }
        a.dbg = dbg;
    }
    a.dPi = dPi; a.dPj = dPj; a.gW1 = gW1; a.gW2 = gW2; a.gb2 = gb2;
    CUtensorMap te, tm, tde, tout;
    int rc;
    // unused maps point at a valid [N,64] array so that encoding never fails
    if ((rc = make_tmap_rows64(&te, first ? Pi : e_prev, first ? N : E, 32))) return rc;
    if ((rc = make_tmap_rows64(&tm, last ? Pi : m, last ? N : E, 32))) return rc;
    if ((rc = make_tmap_rows64(&tde, (last || no_de) ? Pi : de, (last || no_de) ? N : E, 32))) return rc;
    if ((rc = make_tmap_rows64(&tout, first ? Pi : de_prev, first ? N : E, 128))) return rc;
    int64_t tiles = (E + TILE - 1) / TILE;
    const int64_t per_cta = (tiles + sm_count() - 1) / sm_count();
    int64_t grid = (tiles + per_cta - 1) / per_cta;        // every CTA owns >= 1 tile of a contiguous range
    const int tok = prof_begin(first ? GCRL_PROF_EDGE_BWD_FIRST : ((last || no_de) ? GCRL_PROF_EDGE_BWD_LAST : GCRL_PROF_EDGE_BWD), st);
    if (first) rc = x3 ? launch_ws_variant<true, false, true>(te, tm, tde, tout, a, (int)grid, st)
                       : launch_ws_variant<true, false, false>(te, tm, tde, tout, a, (int)grid, st);
    else if (last) rc = x3 ? launch_ws_variant<false, true, true>(te, tm, tde, tout, a, (int)grid, st)
                           : launch_ws_variant<false, true, false>(te, tm, tde, tout, a, (int)grid, st);
    else rc = x3 ? launch_ws_variant<false, false, true>(te, tm, tde, tout, a, (int)grid, st)
                 : launch_ws_variant<false, false, false>(te, tm, tde, tout, a, (int)grid, st);
    prof_end(tok, st);
    if (rc) return rc;
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // namespace gcrl
