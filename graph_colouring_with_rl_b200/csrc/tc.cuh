// tc.cuh -- sm_100a tensor-core plumbing used by the edge kernels: mbarriers, TMA (bulk tensor
// copies), TMEM allocation, tcgen05.mma (kind::f16, bf16 operands, fp32 accumulate in TMEM) with
// shared-memory matrix descriptors, tcgen05.ld.  Inline PTX only; no CUTLASS.
//
// Precision scheme ("bf16x3"): every fp32 operand x is split into hi = bf16(x) and
// lo = bf16(x - hi); a product A*B is issued as A_hi*B_hi + A_lo*B_hi + A_hi*B_lo with fp32
// accumulation, i.e. ~16 mantissa bits per operand (TF32 would give 10).  Mode "bf16" issues the
// hi*hi term only.
//
// Operand tile format: a [R rows x 64] bf16 tile is R lines of 128 bytes; line r sits at byte
// offset r*128 and its eight 16-byte chunks are XOR-swizzled with (r & 7) (SWIZZLE_128B).  For
// 16-bit types this one byte layout is BOTH the canonical K-major operand (rows = M/N, columns =
// K) and the canonical MN-major operand (rows = K, columns = M/N) of tcgen05 matrix descriptors,
// so a tile written once serves the products that contract over features (message MLP, its
// dgrads) and the ones that contract over edge rows (weight gradients).  (TF32 operands do not
// have this property: MN-major tf32 requires the 32-byte-atom swizzle.)
//
// fp32 staging tiles (TMA landing, epilogue exchange, TMA store source) use "SW128 slabs":
// [R x 64] fp32 as two slabs of [R x 32]; slab s holds columns [32s, 32s+32); row r of a slab is a
// 128-byte line at r*128 with chunks XOR-swizzled by (r & 7) -- what a TMA box {32, R} with
// CU_TENSOR_MAP_SWIZZLE_128B reads/writes; conflict-free both for row-per-lane and
// column-per-lane access.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace gcrl {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

// byte offset of 16-byte chunk c4 (0..15) of row r inside a tile whose slabs hold R rows
template <int R>
__device__ __forceinline__ uint32_t sw_off(int r, int c4) {
    return (uint32_t)((c4 >> 3) * (R * 128) + r * 128 + ((((c4 & 7) ^ (r & 7))) << 4));
}

// ---- mbarrier ------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");   // suspend-time hint: sleep, don't spin
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time when the phase is open)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps instead of hanging the GPU (out of line: the kernels have dozens of wait
// sites and their code must stay inside the instruction cache)
static __device__ __noinline__ void mbar_timeout_trap() {
    printf("gcrl: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
    __trap();
}
// Two wait flavours; each kernel names the one it was measured with.
//   mbar_wait_bounded: counted try_wait loop -- a protocol bug traps after ~2^22 suspended polls instead of hanging
//     the GPU.  k_edge_fwd_wg: 3.27 ms per 512-graph launch with it, 3.49 ms with the tight loop (the extra
//     instructions between polls keep four polling warp groups off each other's issue slots); the vertex-side GEMM
//     kernels use it too.
//   mbar_wait_tight: a three-instruction loop.  k_edge_bwd_wg: 5.11 ms vs 5.32 ms with the counted loop (its two
//     256-thread groups are instruction-fetch bound: every wait site counts).  -DGCRL_MBAR_TIMEOUT turns these into
//     bounded waits too (debugging build).
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 22)) mbar_timeout_trap();      // try_wait itself sleeps up to its time hint
    }
}
#ifdef GCRL_MBAR_TIMEOUT
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait_bounded(bar, parity); }
#else
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@!p bra WAIT_LOOP;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}
#endif

// ---- fences ----------------------------------------------------------------------------------
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMA -------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// L2 eviction policies for the streamed edge tensors: they are touched once per launch and must not push the
// per-vertex arrays (Pi, Pj, dPi, dPj, coefficients: re-used by every tile of a graph) out of L2
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, int c0, int c1, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* m, int c0, int c1, const void* smem_src, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
                 ::"l"(m), "r"(c0), "r"(c1), "r"(smem_u32(smem_src)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, int c0, int c1, const void* smem_src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(m), "r"(c0), "r"(c1), "r"(smem_u32(smem_src)) : "memory");
}
// contiguous global range -> L2 (no shared-memory destination): lets a kernel keep more bytes in
// flight than its landing buffers hold; addr 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- TMEM ------------------------------------------------------------------------------------
// one full warp; writes the base address (lane 0, column c) to *slot
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}

// one lane of a converged warp (elect.sync): code under `if (elect_one())` is single-threaded and the
// compiler knows it, so tcgen05.mma / commit are emitted once instead of inside a per-thread election loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// ---- tcgen05.mma -----------------------------------------------------------------------------
// matrix descriptor of a SWIZZLE_128B operand starting at smem byte address `addr`:
//   bits [0,14) start >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout = 2
// K-major: SBO = 1024 (stride between 8-row groups), LBO unused (16).
// MN-major: SBO = 1024 (stride between 8-K-row groups), LBO = stride between 64-element MN atoms
// (unused here: every operand is exactly 64 wide).
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor, kind::f16 with bf16 operands, fp32 accumulate
//   [4,6) D fmt = 1 (f32) | [7,10) A fmt = 1 (bf16) | [10,13) B fmt = 1 | [15] A MN-major | [16] B MN-major
//   [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand read from tensor memory (lane = row, one 32-bit column = two consecutive
// K elements, the even one in the low half; a K = 16 step is 8 columns)
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate ? 1u : 0u) : "memory");
}
// arrive on `bar` when all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// byte offset of 16-byte chunk cb (0..7, 8 bf16 each) of line r in a bf16 operand tile
__device__ __forceinline__ uint32_t bf_off(int r, int cb) { return (uint32_t)(r * 128 + ((cb ^ (r & 7)) << 4)); }

// two floats -> packed bf16x2 (round to nearest even); `lo` operand lands in the low half
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// split (x0, x1) into packed hi = bf16(x) and lo = bf16(x - hi)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(x0, x1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(x0 - h0, x1 - h1);
}
// split 8 consecutive floats into bf16 hi / lo chunks (16 bytes each)
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
    split2(x[0], x[1], hi.x, lo.x);
    split2(x[2], x[3], hi.y, lo.y);
    split2(x[4], x[5], hi.z, lo.z);
    split2(x[6], x[7], hi.w, lo.w);
}
template <bool X3>
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
    if (X3) {
        split2(v.x, v.y, hi.x, lo.x);
        split2(v.z, v.w, hi.y, lo.y);
    } else {
        hi.x = pack_bf16x2(v.x, v.y);
        hi.y = pack_bf16x2(v.z, v.w);
    }
}

// D[R x 64] (+)= A[R x 64] * B[64 x 64]^T contracting over the 64 columns (K-major operands):
// 4 K=16 steps; X3 adds the lo*hi and hi*lo correction terms.  a_lo/b_lo ignored when !X3.
template <bool X3, bool ROLL = false>
__device__ __forceinline__ void issue_gemm_feat(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                                uint32_t b_lo, uint32_t idesc, bool accumulate_first) {
    auto step = [&](int k) {
        const uint64_t ah = make_desc(a_hi + 32 * k, 16, 1024), bh = make_desc(b_hi + 32 * k, 16, 1024);
        umma_bf16(d_tmem, ah, bh, idesc, accumulate_first || k > 0);
        if (X3) {
            umma_bf16(d_tmem, make_desc(a_lo + 32 * k, 16, 1024), bh, idesc, true);
            umma_bf16(d_tmem, ah, make_desc(b_lo + 32 * k, 16, 1024), idesc, true);
        }
    };
    if (ROLL) {
#pragma unroll 1
        for (int k = 0; k < 4; ++k) step(k);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) step(k);
    }
}
// D[R x 64] (+)= A[R x 64] * W[64 x 64] (no transpose): A K-major, W read as an MN-major operand
// (rows of the weight tile = contraction index): the B descriptor walks 16 rows (2048 bytes) per step.
template <bool X3, bool ROLL = false>
__device__ __forceinline__ void issue_gemm_featT(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                                 uint32_t b_lo, uint32_t idesc, bool accumulate_first) {
    auto step = [&](int k) {
        const uint64_t ah = make_desc(a_hi + 32 * k, 16, 1024), bh = make_desc(b_hi + 2048 * k, 8192, 1024);
        umma_bf16(d_tmem, ah, bh, idesc, accumulate_first || k > 0);
        if (X3) {
            umma_bf16(d_tmem, make_desc(a_lo + 32 * k, 16, 1024), bh, idesc, true);
            umma_bf16(d_tmem, ah, make_desc(b_lo + 2048 * k, 8192, 1024), idesc, true);
        }
    };
    if (ROLL) {
#pragma unroll 1
        for (int k = 0; k < 4; ++k) step(k);
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) step(k);
    }
}
// D[64 x 64] (+)= A[rows x 64]^T * B[rows x 64] contracting over `rows` edge rows (MN-major
// operands, M = N = 64): rows/16 K=16 steps, each advancing two 8-row groups (2048 bytes).
template <bool X3, bool ROLL = false>
__device__ __forceinline__ void issue_gemm_rows(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi,
                                                uint32_t b_lo, uint32_t idesc, int rows, bool accumulate_first) {
    auto step = [&](int k) {
        const uint64_t ah = make_desc(a_hi + 2048 * k, 8192, 1024), bh = make_desc(b_hi + 2048 * k, 8192, 1024);
        umma_bf16(d_tmem, ah, bh, idesc, accumulate_first || k > 0);
        if (X3) {
            umma_bf16(d_tmem, make_desc(a_lo + 2048 * k, 8192, 1024), bh, idesc, true);
            umma_bf16(d_tmem, ah, make_desc(b_lo + 2048 * k, 8192, 1024), idesc, true);
        }
    };
    if (ROLL) {
#pragma unroll 1
        for (int k = 0; k < rows / 16; ++k) step(k);
    } else {
        for (int k = 0; k < rows / 16; ++k) step(k);
    }
}

// D[128 x 64] (+)= A * B^T with A [128 x 64] in tensor memory as two 32-column halves of 16 hi + 16 lo columns each
// (half h covers K in [32h, 32h + 32): hi at a_tmem + 32h, lo at a_tmem + 32h + 16 -- the layout an epilogue that converts
// 32 accumulator columns at a time writes IN PLACE over the columns it has just read), B [64 x 64] K-major in shared memory
template <bool X3>
__device__ __forceinline__ void issue_gemm_feat_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_hi, uint32_t b_lo, uint32_t idesc,
                                                   bool accumulate_first) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t ah = a_tmem + 32 * (k >> 1) + 8 * (k & 1), al = ah + 16;
        const uint64_t bh = make_desc(b_hi + 32 * k, 16, 1024);
        umma_bf16_ts(d_tmem, ah, bh, idesc, accumulate_first || k > 0);
        if (X3) {
            umma_bf16_ts(d_tmem, al, bh, idesc, true);
            umma_bf16_ts(d_tmem, ah, make_desc(b_lo + 32 * k, 16, 1024), idesc, true);
        }
    }
}
// 32 fp32 values of one tile row -> bf16 hi (16 packed columns) | lo (16 packed columns) at taddr (this thread's lane)
template <bool X3>
__device__ __forceinline__ void tmem_st_split32(uint32_t taddr, const float v[32]) {
    uint32_t h[16], l[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        if (X3) split2(v[2 * j], v[2 * j + 1], h[j], l[j]);
        else { h[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]); l[j] = 0u; }
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]),
          "r"(h[8]), "r"(h[9]), "r"(h[10]), "r"(h[11]), "r"(h[12]), "r"(h[13]), "r"(h[14]), "r"(h[15]) : "memory");
    if (X3)
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
            ::"r"(taddr + 16), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7]),
              "r"(l[8]), "r"(l[9]), "r"(l[10]), "r"(l[11]), "r"(l[12]), "r"(l[13]), "r"(l[14]), "r"(l[15]) : "memory");
}

// ---- tcgen05.ld: 32 lanes x 32 consecutive 32-bit columns; thread i of the warp gets lane
// (taddr.lane + i); the warp may only touch lanes [32*(warp%4), 32*(warp%4)+32) -------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
          "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
          "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// pointer to 16-byte chunk c4 of row r of a 128-row fp32 SW128-slab tile
__device__ __forceinline__ float* l_ptr(float* L, int r, int c4) {
    return reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(L) + sw_off<128>(r, c4));
}
__device__ __forceinline__ const float* l_ptr(const float* L, int r, int c4) {
    return reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(L) + sw_off<128>(r, c4));
}

// weights [64 x ld] row-major (columns col0..col0+63) -> bf16 hi/lo operand tiles (rows = out, cols = in)
__device__ __forceinline__ void load_weight_tile(const float* W, int ld, int col0, __nv_bfloat16* hi, __nv_bfloat16* lo,
                                                 int tid, int nthreads) {
    for (int t = tid; t < 64 * 8; t += nthreads) {
        const int o = t >> 3, cb = t & 7;
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = W[o * ld + col0 + 8 * cb + j];
        uint4 h, l;
        split8(x, h, l);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(hi) + bf_off(o, cb)) = h;
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(lo) + bf_off(o, cb)) = l;
    }
}

// flat global -> shared copy of a pre-converted weight image (multiples of 16 bytes), eight 16-byte loads in flight per thread:
// ONE L2 round trip per batch instead of one per tile row of in-kernel conversion (weight_image.cu)
__device__ __forceinline__ void copy_image(void* smem_dst, const unsigned char* gsrc, int nbytes, int tid, int nthreads) {
    const uint4* src = reinterpret_cast<const uint4*>(gsrc);
    uint4* dst = reinterpret_cast<uint4*>(smem_dst);
    const int n = nbytes >> 4;
    for (int t0 = tid; t0 < n; t0 += 8 * nthreads) {
        uint4 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) if (t0 + k * nthreads < n) v[k] = __ldg(src + t0 + k * nthreads);
#pragma unroll
        for (int k = 0; k < 8; ++k) if (t0 + k * nthreads < n) dst[t0 + k * nthreads] = v[k];
    }
}

// 8 floats -> bf16 chunk(s) of an operand tile at (row, chunk cb)
template <bool X3>
__device__ __forceinline__ void store_split8(const float* x, __nv_bfloat16* hi, __nv_bfloat16* lo, int row, int cb) {
    uint4 h, l;
    if (X3) {
        split8(x, h, l);
    } else {
        h = make_uint4(pack_bf16x2(x[0], x[1]), pack_bf16x2(x[2], x[3]), pack_bf16x2(x[4], x[5]), pack_bf16x2(x[6], x[7]));
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(hi) + bf_off(row, cb)) = h;
    if (X3) *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(lo) + bf_off(row, cb)) = l;
}

// fp32 SW128-slab tile (128 rows) -> bf16 hi/lo operand tile; half a warp per 256-byte row
template <bool X3>
__device__ __forceinline__ void convert_tile128(const float* L, __nv_bfloat16* hi, __nv_bfloat16* lo, int tid) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int t = it * 256 + tid;
        const int r = t >> 4, c4 = t & 15;
        const float4 v = *reinterpret_cast<const float4*>(l_ptr(L, r, c4));
        uint2 h, l;
        split4<X3>(v, h, l);
        const uint32_t off = bf_off(r, c4 >> 1) + (c4 & 1) * 8;
        *reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(hi) + off) = h;
        if (X3) *reinterpret_cast<uint2*>(reinterpret_cast<unsigned char*>(lo) + off) = l;
    }
}

}  // namespace tc

// host side: tensor map of a row-major fp32 matrix [rows, 64] with box {32, box_rows}, SWIZZLE_128B
int make_tmap_rows64(CUtensorMap* out, const float* base, int64_t rows, int box_rows);

}  // namespace gcrl
