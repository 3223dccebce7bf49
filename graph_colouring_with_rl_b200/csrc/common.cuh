// common.cuh -- shared helpers for libgcrl_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <math.h>
#include <stdlib.h>
#include <nvtx3/nvToolsExt.h>
#include "../../include/gcrl_b200.h"

#define HID GCRL_HID

namespace gcrl {

void set_error(const char* fmt, ...);
int sm_count();
void count_launch(int n);                       // prof.cu: kernels launched by the library
int prof_begin(int kind, cudaStream_t st);      // prof.cu: event timing when enabled, else -1
void prof_end(int token, cudaStream_t st);

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------------------------
// The small-batch paths (a 64-transition learn step, batch-1 act, rollouts of a few hundred graphs) are chains of
// ~10 us kernels, each depending on the one before: the launch gap is a third of their time.  Every kernel of the
// library starts with pdl_prologue(): `griddepcontrol.launch_dependents` lets the NEXT kernel of the stream be launched
// (its CTAs scheduled, parameters fetched) while this one still runs; `griddepcontrol.wait` then blocks until the
// PREVIOUS kernel has completed and its memory is visible -- before this kernel touches any global memory, so the
// stream's data dependences hold exactly as without PDL (transitively: no kernel completes before its own wait).
// launch_k() launches with the programmatic-stream-serialization attribute (GCRL_PDL=0 switches it off: the two PTX
// instructions are then no-ops).  Works under CUDA-graph capture (programmatic edges).
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
bool pdl_enabled();   // prof.cu
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// NVTX range for the profilers' timelines (nsys / ncu --nvtx): header-only nvtx3, a no-op pointer check unless a tool
// is attached.  One range per entry point and per block / direction inside the network passes.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    NvtxRange(const char* fmt, int a) { char b[64]; snprintf(b, sizeof b, fmt, a); nvtxRangePushA(b); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

#define GCRL_CHECK_ARG(cond)                                                          \
    do {                                                                              \
        if (!(cond)) {                                                                \
            gcrl::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, #cond); \
            return GCRL_ERR_INVALID;                                                  \
        }                                                                             \
    } while (0)

#define GCRL_CUDA(call)                                                               \
    do {                                                                              \
        cudaError_t e_ = (call);                                                      \
        if (e_ != cudaSuccess) {                                                      \
            gcrl::set_error("%s:%d: %s: %s", __FILE__, __LINE__, #call,               \
                            cudaGetErrorString(e_));                                  \
            return GCRL_ERR_CUDA;                                                     \
        }                                                                             \
    } while (0)

#define GCRL_LAUNCH_CHECK()              \
    do {                                 \
        gcrl::count_launch(1);           \
        GCRL_CUDA(cudaGetLastError());   \
    } while (0)

// ---- flat parameter layout (state_dict order, architecture.py:157-167) -----------------
struct BlockW {
    const float* W1;  // [64, 2*Dv + De]  message_mlp.0.weight
    const float* b1;  // [64]
    const float* W2;  // [64, 64]         message_mlp.2.weight
    const float* b2;  // [64]
    const float* U1;  // [64, 256 + Dv]   update_mlp.0.weight
    const float* c1;  // [64]
    const float* U2;  // [64, 64]         update_mlp.2.weight
    const float* c2;  // [64]
    int Dv, De;
};
struct HeadW {
    const float *F1, *f1, *F2, *f2, *F3, *f3;
    int Dg;
};
struct NetW {
    BlockW blk[GCRL_N_BLOCKS];
    HeadW head;
    int64_t total;
};

inline NetW make_net(const float* p, int vdim, int edim, int gdim) {
    NetW n;
    int64_t o = 0;
    int dv = vdim, de = edim;
    for (int l = 0; l < GCRL_N_BLOCKS; ++l) {
        BlockW& b = n.blk[l];
        b.Dv = dv; b.De = de;
        b.W1 = p + o; o += (int64_t)HID * (2 * dv + de);
        b.b1 = p + o; o += HID;
        b.W2 = p + o; o += HID * HID;
        b.b2 = p + o; o += HID;
        b.U1 = p + o; o += (int64_t)HID * (4 * HID + dv);
        b.c1 = p + o; o += HID;
        b.U2 = p + o; o += HID * HID;
        b.c2 = p + o; o += HID;
        dv = HID; de = HID;
    }
    n.head.Dg = gdim;
    n.head.F1 = p + o; o += (int64_t)HID * (HID + gdim);
    n.head.f1 = p + o; o += HID;
    n.head.F2 = p + o; o += HID * HID;
    n.head.f2 = p + o; o += HID;
    n.head.F3 = p + o; o += HID;
    n.head.f3 = p + o; o += 1;
    n.total = o;
    return n;
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// bump allocator over a caller-provided workspace
struct Arena {
    char* base; size_t size; size_t used;
    Arena(void* b, size_t s) : base((char*)b), size(s), used(0) {}
    template <typename T> T* take(size_t n) {
        size_t bytes = align_up(n * sizeof(T));
        T* p = (T*)(base + used);
        used += bytes;
        return p;
    }
    bool ok() const { return used <= size; }
};

#define PNA_EPS 1e-5f  // principal_neighbourhood_aggregation.py:64

}  // namespace gcrl
