// pack.cu -- graph batching: PyG-style (edge_index) or dense-adjacency batches -> the
// destination-major segment layout every other kernel consumes.
// Replaces: torch_geometric Batch.from_data_list (dqn_agent.py:55-56) and the gather/index
// plumbing of MessagePassing.__collect__ (architecture.py:84).  Integer-only; bit-exact.
#include "common.cuh"

namespace gcrl {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n = 148;
    }
    return n;
}

// ---- in-degree histogram ---------------------------------------------------------------
__global__ void k_count_deg(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                            int32_t* __restrict__ deg, int32_t* __restrict__ err) {
    pdl_prologue();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < E;
         k += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = ei[k], d = ei[E + k];
        if (s < 0 || s >= N || d < 0 || d >= N) {
            if (err) *err = 1;
            continue;
        }
        atomicAdd(&deg[d], 1);
    }
}

// ---- single-CTA exclusive scan: out[0..n] (n+1 entries) from in[0..n) --------------------
// n <= a few million vertices; one CTA of 1024 threads walks 4096-element chunks.
__global__ void __launch_bounds__(1024) k_exclusive_scan(const int32_t* __restrict__ in,
                                                         int32_t* __restrict__ out, int64_t n) {
    pdl_prologue();
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < n; base += 4096) {
        int64_t i0 = base + (int64_t)threadIdx.x * 4;
        int32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (i0 + j < n) ? in[i0 + j] : 0;
        int32_t tsum = v[0] + v[1] + v[2] + v[3];
        int32_t incl = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int32_t w = warp_tot[lane];
            int32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[lane] = wi - w;  // exclusive prefix of warp totals
        }
        __syncthreads();
        int32_t carry = carry_s;
        int32_t excl = carry + warp_tot[wid] + incl - tsum;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i0 + j < n) out[i0 + j] = excl;
            excl += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl;  // total through this chunk
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

// ---- scatter caller edge ids into their segment (arbitrary order inside the segment) ----
__global__ void k_scatter_ids(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                              const int32_t* __restrict__ seg_off, int32_t* __restrict__ cursor,
                              int32_t* __restrict__ tmp) {
    pdl_prologue();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < E;
         k += (int64_t)gridDim.x * blockDim.x) {
        int64_t s = ei[k], d = ei[E + k];
        if (s < 0 || s >= N || d < 0 || d >= N) continue;
        int32_t pos = atomicAdd(&cursor[d], 1);
        tmp[seg_off[d] + pos] = (int32_t)k;
    }
}

// ---- order each segment by caller edge id (rank by counting; segments are short) --------
__global__ void k_rank_in_segment(const int64_t* __restrict__ ei, int64_t E,
                                  const int32_t* __restrict__ seg_off,
                                  const int32_t* __restrict__ tmp, int64_t N,
                                  int32_t* __restrict__ src, int32_t* __restrict__ dst,
                                  int32_t* __restrict__ perm) {
    pdl_prologue();
    const int32_t total = seg_off[N];  // edges with valid ids (all of them unless err is set)
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < total;
         s += (int64_t)gridDim.x * blockDim.x) {
        int32_t k = tmp[s];
        int32_t d = (int32_t)ei[E + k];
        int32_t a = seg_off[d], b = seg_off[d + 1];
        int32_t rank = 0;
        for (int32_t t = a; t < b; ++t) rank += (tmp[t] < k);
        int32_t slot = a + rank;
        perm[slot] = k;
        src[slot] = (int32_t)ei[k];
        dst[slot] = d;
    }
}

// ---- dense complete-graph batches -------------------------------------------------------
__global__ void k_graph_edge_counts(const int32_t* __restrict__ node_off, int32_t G,
                                    int32_t* __restrict__ cnt) {
    pdl_prologue();
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < G) {
        int32_t n = node_off[g + 1] - node_off[g];
        cnt[g] = n * (n - 1);
    }
}

// one warp per destination vertex
__global__ void k_pack_dense(const uint8_t* __restrict__ adj, const int64_t* __restrict__ adj_off,
                             const int32_t* __restrict__ node_off,
                             const int32_t* __restrict__ edge_off, int32_t G, int64_t N,
                             int32_t* __restrict__ seg_off, int32_t* __restrict__ src,
                             int32_t* __restrict__ dst, int32_t* __restrict__ perm,
                             float* __restrict__ eattr) {
    pdl_prologue();
    const int lane = threadIdx.x & 31;
    int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp; v < N; v += nwarps) {
        // graph of v: last g with node_off[g] <= v
        int lo = 0, hi = G;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (node_off[mid] <= v) lo = mid; else hi = mid;
        }
        const int g = lo;
        const int32_t v0 = node_off[g];
        const int32_t n = node_off[g + 1] - v0;
        const int32_t i = (int32_t)(v - v0);
        const int32_t e0 = edge_off[g];
        const int32_t s0 = e0 + i * (n - 1);
        if (lane == 0) {
            seg_off[v] = s0;
            if (v == N - 1) seg_off[N] = s0 + (n - 1);
        }
        const uint8_t* A = adj + adj_off[g];
        for (int32_t r = lane; r < n - 1; r += 32) {
            int32_t j = r + (r >= i);
            int32_t a = min(i, j), b = max(i, j);
            // pair index of (a,b), a<b, in the canonical order of utils.py:304-310
            int32_t p = a * n - (a * (a + 1)) / 2 + (b - a - 1);
            int32_t caller = e0 + 2 * p + (j > i ? 1 : 0);
            src[s0 + r] = v0 + j;
            dst[s0 + r] = (int32_t)v;
            perm[s0 + r] = caller;
            eattr[s0 + r] = A[(int64_t)j * n + i] ? -1.0f : 0.0f;
        }
    }
}

__global__ void k_permute_rows(const float* __restrict__ in, float* __restrict__ out,
                               const int32_t* __restrict__ perm, int64_t rows, int32_t width,
                               int to_internal) {
    pdl_prologue();
    int64_t total = rows * width;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = t / width;
        int32_t c = (int32_t)(t - r * width);
        int64_t p = perm[r];
        if (to_internal) out[t] = in[p * width + c];
        else out[p * width + c] = in[t];
    }
}

}  // namespace gcrl

using namespace gcrl;

extern "C" {

const char* gcrl_last_error(void) { return g_err; }
int gcrl_version(void) { return 100; }
int gcrl_sm_count(void) { return sm_count(); }

int gcrl_param_layout(int vdim, int edim, int gdim, int64_t* off) {
    GCRL_CHECK_ARG(off != nullptr && vdim >= 1 && edim >= 1 && gdim >= 0);
    int64_t o = 0;
    int i = 0, dv = vdim, de = edim;
    for (int l = 0; l < GCRL_N_BLOCKS; ++l) {
        off[i++] = o; o += (int64_t)HID * (2 * dv + de);
        off[i++] = o; o += HID;
        off[i++] = o; o += HID * HID;
        off[i++] = o; o += HID;
        off[i++] = o; o += (int64_t)HID * (4 * HID + dv);
        off[i++] = o; o += HID;
        off[i++] = o; o += HID * HID;
        off[i++] = o; o += HID;
        dv = HID; de = HID;
    }
    off[i++] = o; o += (int64_t)HID * (HID + gdim);
    off[i++] = o; o += HID;
    off[i++] = o; o += HID * HID;
    off[i++] = o; o += HID;
    off[i++] = o; o += HID;
    off[i++] = o; o += 1;
    off[i] = o;
    return GCRL_OK;
}

size_t gcrl_pack_workspace_bytes(int64_t N, int64_t E) {
    return align_up((size_t)(N + 1) * 4) * 2 + align_up((size_t)(E > 0 ? E : 1) * 4) + 1024;
}

int gcrl_pack(const int64_t* ei, int64_t E, int64_t N, int32_t* seg_off, int32_t* src,
              int32_t* dst, int32_t* perm, int32_t* err, void* ws, size_t ws_bytes, void* stream) {
    GCRL_CHECK_ARG(N >= 0 && E >= 0 && E < (1ll << 31) && N < (1ll << 31));
    GCRL_CHECK_ARG(seg_off && (E == 0 || (ei && src && dst && perm)));
    if (ws_bytes < gcrl_pack_workspace_bytes(N, E) || !ws) {
        set_error("gcrl_pack: workspace too small");
        return GCRL_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(ws, ws_bytes);
    int32_t* deg = ar.take<int32_t>(N + 1);
    int32_t* cursor = ar.take<int32_t>(N + 1);
    int32_t* tmp = ar.take<int32_t>(E > 0 ? E : 1);
    GCRL_CUDA(cudaMemsetAsync(deg, 0, (N + 1) * 4, st));
    GCRL_CUDA(cudaMemsetAsync(cursor, 0, (N + 1) * 4, st));
    if (err) GCRL_CUDA(cudaMemsetAsync(err, 0, 4, st));
    const int threads = 256;
    int blocks = (int)((E + threads - 1) / threads);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    if (blocks < 1) blocks = 1;
    if (E > 0) launch_k(k_count_deg, dim3(blocks), dim3(threads), 0, st, ei, E, N, deg, err);
    launch_k(k_exclusive_scan, dim3(1), dim3(1024), 0, st, deg, seg_off, N);
    if (E > 0) {
        launch_k(k_scatter_ids, dim3(blocks), dim3(threads), 0, st, ei, E, N, seg_off, cursor, tmp);
        // edges with invalid ids are dropped (err flag set); ranking walks only the
        // seg_off[N] valid slots, a bound it reads on the device.
        launch_k(k_rank_in_segment, dim3(blocks), dim3(threads), 0, st, ei, E, seg_off, tmp, N, src, dst, perm);
        count_launch(3);
    }
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

size_t gcrl_pack_dense_workspace_bytes(int32_t G) { return align_up((size_t)(G + 2) * 4) * 2; }

int gcrl_pack_dense(const uint8_t* adj, const int64_t* adj_off, const int32_t* node_off,
                    int32_t G, int64_t N, int64_t E, int32_t* seg_off, int32_t* src,
                    int32_t* dst, int32_t* perm, float* eattr, void* ws, size_t ws_bytes,
                    void* stream) {
    GCRL_CHECK_ARG(G >= 0 && N >= 0 && E >= 0 && E < (1ll << 31));
    GCRL_CHECK_ARG(adj && adj_off && node_off && seg_off);
    if (!ws || ws_bytes < gcrl_pack_dense_workspace_bytes(G)) {
        set_error("gcrl_pack_dense: workspace too small");
        return GCRL_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(ws, ws_bytes);
    int32_t* cnt = ar.take<int32_t>(G + 2);
    int32_t* eoff = ar.take<int32_t>(G + 2);
    if (G > 0) launch_k(k_graph_edge_counts, dim3((G + 255) / 256), dim3(256), 0, st, node_off, G, cnt);
    launch_k(k_exclusive_scan, dim3(1), dim3(1024), 0, st, cnt, eoff, G);
    if (N == 0) GCRL_CUDA(cudaMemsetAsync(seg_off, 0, 4, st));
    if (N > 0) {
        int blocks = (int)((N * 32 + 255) / 256);
        if (blocks > sm_count() * 32) blocks = sm_count() * 32;
        launch_k(k_pack_dense, dim3(blocks), dim3(256), 0, st, adj, adj_off, node_off, eoff, G, N, seg_off, src, dst,
                                             perm, eattr);
        count_launch(G > 0 ? 2 : 1);
    }
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

int gcrl_permute_rows(const float* in, float* out, const int32_t* perm, int64_t rows,
                      int32_t width, int to_internal, void* stream) {
    GCRL_CHECK_ARG(rows >= 0 && width >= 1);
    if (rows == 0) return GCRL_OK;
    GCRL_CHECK_ARG(in && out && perm);
    int64_t total = rows * width;
    int blocks = (int)((total + 255) / 256);
    if (blocks > sm_count() * 32) blocks = sm_count() * 32;
    launch_k(k_permute_rows, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, in, out, perm, rows, width, to_internal);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // extern "C"
