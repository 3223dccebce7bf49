// ws.cuh -- helpers shared by the multi-group edge kernels (gn_forward_wg.cu, gn_backward_wg.cu):
// mbarrier arrivals per warp, named barriers, TMEM stores in the MMA fragment layout, warp transpose-reduce.
#pragma once
#include "tc.cuh"

namespace gcrl {

using namespace tc;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// every lane has finished its shared-memory / TMEM work -> one arrival per warp
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
// tcgen05.st 16 lanes x 64 columns in the 16x256b fragment layout: thread t holds, for i = 0..7,
// v[4i], v[4i+1] = (lane t/4, columns 8i + 2(t%4) + {0,1}) and v[4i+2], v[4i+3] = (lane t/4 + 8, same columns)
__device__ __forceinline__ void tmem_st_16x256b_x8(uint32_t taddr, const float v[32]) {
    const uint32_t* u = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x8.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]),
          "r"(u[8]), "r"(u[9]), "r"(u[10]), "r"(u[11]), "r"(u[12]), "r"(u[13]), "r"(u[14]), "r"(u[15]),
          "r"(u[16]), "r"(u[17]), "r"(u[18]), "r"(u[19]), "r"(u[20]), "r"(u[21]), "r"(u[22]), "r"(u[23]),
          "r"(u[24]), "r"(u[25]), "r"(u[26]), "r"(u[27]), "r"(u[28]), "r"(u[29]), "r"(u[30]), "r"(u[31])
        : "memory");
}
// tcgen05.ld counterpart: 16 lanes x 64 columns in the 16x256b fragment layout (same register <-> element map)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, float v[32]) {
    uint32_t* u = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
          "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
          "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// two 32-bit columns per lane (thread i <-> lane base + i): the relu mask of a tile row
__device__ __forceinline__ void tmem_st_x2(uint32_t taddr, uint32_t a0, uint32_t a1) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a0), "r"(a1) : "memory");
}
__device__ __forceinline__ void tmem_st_x1(uint32_t taddr, uint32_t a0) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(a0) : "memory");
}
__device__ __forceinline__ void tmem_ld_x2(uint32_t taddr, uint32_t& a0, uint32_t& a1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a0), "=r"(a1) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// column sums over the 32 lanes of a warp: on return lane l holds sum_lanes v[l] (v is destroyed)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const bool up = (lane & w) != 0;
#pragma unroll
        for (int k = 0; k < w; ++k) {
            const float send = up ? v[k] : v[k + w];
            const float keep = up ? v[k + w] : v[k];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
    }
    return v[0];
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(a, b));
}


}  // namespace gcrl
