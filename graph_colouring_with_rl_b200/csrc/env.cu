// env.cu -- the graph-colouring environment on the device (integer only, bit-exact).
// Replaces graph_colouring_env.py:24-61 (reset), :64-90 (step), :102-109 (reward),
// :144-191 (colour_vertex / colour_leaf_nodes), utils.py:285-293 (first-fit colour),
// :522-533,:559-565 (<=3-hop neighbour counts) and test_policy_functions.py:46-47
// (first-vertex rule).  State per graph: colour[v] (-1 = uncoloured)
// and the highest colour index used; adjacency is a dense byte matrix per graph.
//
// One CTA per graph.  The reference colours leaf vertices (no uncoloured neighbour) one after
// another; leaves are pairwise non-adjacent and colouring one changes no other vertex's leaf
// status, so evaluating them in parallel gives identical colours.
#include "common.cuh"

namespace gcrl {

constexpr int ENV_THREADS = 128;
constexpr int ENV_WARPS = ENV_THREADS / 32;
constexpr int ENV_MAX_N = 4096;                 // vertices per graph supported by the env kernels
constexpr int ENV_WORDS = ENV_MAX_N / 32;

__global__ void k_env_reset(int32_t* __restrict__ colour, int32_t* __restrict__ max_colour,
                            float* __restrict__ x, const int32_t* __restrict__ node_off, int32_t G,
                            int64_t N) {
    pdl_prologue();
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t v = t; v < N; v += stride) colour[v] = -1;
    for (int64_t g = t; g < G; g += stride) {
        max_colour[g] = -1;
        if (x) {
            const int32_t v0 = node_off[g], n = node_off[g + 1] - v0;
            for (int32_t i = 0; i < n; ++i) {   // GenerateData_v1: x = [vertex name, colour]
                x[(int64_t)(v0 + i) * 2] = (float)i;
                x[(int64_t)(v0 + i) * 2 + 1] = -1.0f;
            }
        }
    }
}

// first colour index not present in `used` (bitmask words), scanned by one warp
__device__ __forceinline__ int32_t first_free_colour(const uint32_t* used, int words, int lane) {
    for (int w0 = 0; w0 < words; w0 += 32) {
        const int w = w0 + lane;
        const uint32_t inv = (w < words) ? ~used[w] : 0u;
        const unsigned has = __ballot_sync(0xffffffffu, inv != 0u);
        if (has) {
            const int src_lane = __ffs(has) - 1;
            const uint32_t word = __shfl_sync(0xffffffffu, inv, src_lane);
            return (w0 + src_lane) * 32 + (__ffs(word) - 1);
        }
    }
    return words * 32;   // unreachable: a vertex has < n coloured neighbours
}

__global__ void __launch_bounds__(ENV_THREADS) k_env_step(
    const uint8_t* __restrict__ adj, const int64_t* __restrict__ adj_off,
    const int32_t* __restrict__ node_off, int32_t G, const int32_t* __restrict__ action,
    int32_t* __restrict__ colour, int32_t* __restrict__ max_colour, float* __restrict__ reward,
    int32_t* __restrict__ done, float* __restrict__ x, int32_t* __restrict__ err,
    const float* __restrict__ q, int32_t* __restrict__ action_out) {
    pdl_prologue();
    __shared__ int32_t col_s[ENV_MAX_N];
    __shared__ uint32_t used_s[ENV_WARPS][ENV_WORDS];
    __shared__ int32_t maxc_s, uncol_s;
    __shared__ float bq_s[ENV_WARPS];
    __shared__ int32_t ba_s[ENV_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int g = blockIdx.x; g < G; g += gridDim.x) {
        const int32_t v0 = node_off[g], n = node_off[g + 1] - v0;
        if (n > ENV_MAX_N) { if (tid == 0 && err) *err = 1; continue; }
        const uint8_t* A = adj + adj_off[g];
        const int words = (n + 31) >> 5;
        const int32_t prev_max = max_colour[g];
        for (int i = tid; i < n; i += ENV_THREADS) col_s[i] = colour[v0 + i];
        if (tid == 0) { maxc_s = prev_max; uncol_s = 0; }
        __syncthreads();
        int32_t a;
        if (q) {
            // greedy action scoring fused into the step (dqn_agent.py:131-134): first maximum of q over the uncoloured
            // vertices = np.argmax over the ascending unassigned list; -1 when the graph is finished
            float best = -INFINITY;
            int32_t arg = -1;
            for (int i = tid; i < n; i += ENV_THREADS) {
                if (col_s[i] < 0) {
                    const float v = q[v0 + i];
                    if (arg < 0 || v > best) { best = v; arg = i; }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int32_t oa = __shfl_xor_sync(0xffffffffu, arg, o);
                if (oa >= 0 && (arg < 0 || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
            }
            if (lane == 0) { bq_s[wid] = best; ba_s[wid] = arg; }
            __syncthreads();
            best = bq_s[0]; arg = ba_s[0];
#pragma unroll
            for (int w = 1; w < ENV_WARPS; ++w) {
                const float ob = bq_s[w];
                const int32_t oa = ba_s[w];
                if (oa >= 0 && (arg < 0 || ob > best || (ob == best && oa < arg))) { best = ob; arg = oa; }
            }
            a = arg;
            if (tid == 0 && action_out) action_out[g] = a;
            __syncthreads();            // bq_s / ba_s are reused by the next graph of this CTA
        } else {
            a = action[g];
        }
        const bool act = (a >= 0 && a < n && col_s[a] < 0);
        if (a >= n && tid == 0 && err) *err = 2;   // graph_colouring_env.py:72-73 raises
        if (act) {
            // ---- first-fit colour of the chosen vertex (utils.py:285-293) --------------------
            if (wid == 0) {
                for (int w = lane; w < words; w += 32) used_s[0][w] = 0u;
                __syncwarp();
                for (int u = lane; u < n; u += 32) {
                    const int32_t c = col_s[u];
                    if (u != a && c >= 0 && A[(int64_t)a * n + u]) atomicOr(&used_s[0][c >> 5], 1u << (c & 31));
                }
                __syncwarp();
                const int32_t c = first_free_colour(used_s[0], words, lane);
                if (lane == 0) { col_s[a] = c; maxc_s = max(maxc_s, c); }
            }
            __syncthreads();
            // ---- leaf pass (graph_colouring_env.py:185-191): warp per candidate vertex --------
            // Decisions are taken against the post-action state; new leaf colours go to the
            // global array only, so no candidate observes another leaf's colour.
            for (int u = wid; u < n; u += ENV_WARPS) {
                if (col_s[u] >= 0) continue;                     // warp-uniform
                bool has_uncol = false;
                for (int w0 = 0; w0 < n && !has_uncol; w0 += 32) {
                    const int w = w0 + lane;
                    const bool p = (w < n) && (w != u) && A[(int64_t)u * n + w] && (col_s[w] < 0);
                    has_uncol = __any_sync(0xffffffffu, p);
                }
                if (has_uncol) continue;
                for (int w = lane; w < words; w += 32) used_s[wid][w] = 0u;
                __syncwarp();
                for (int w = lane; w < n; w += 32) {
                    const int32_t c = col_s[w];
                    if (w != u && c >= 0 && A[(int64_t)u * n + w]) atomicOr(&used_s[wid][c >> 5], 1u << (c & 31));
                }
                __syncwarp();
                const int32_t c = first_free_colour(used_s[wid], words, lane);
                if (lane == 0) {
                    colour[v0 + u] = c;
                    if (x) x[(int64_t)(v0 + u) * 2 + 1] = (float)c;
                    atomicMax(&maxc_s, c);
                    atomicAdd(&uncol_s, -1);   // counted as uncoloured below (col_s[u] < 0)
                }
                __syncwarp();
            }
            if (tid == 0) {
                colour[v0 + a] = col_s[a];
                if (x) x[(int64_t)(v0 + a) * 2 + 1] = (float)col_s[a];
            }
        }
        int local = 0;
        for (int i = tid; i < n; i += ENV_THREADS) local += (col_s[i] < 0);
        if (local) atomicAdd(&uncol_s, local);
        __syncthreads();
        if (tid == 0) {
            max_colour[g] = maxc_s;
            if (reward) reward[g] = -(float)(maxc_s - prev_max);   // :107  -(new colours used)
            if (done) done[g] = (uncol_s == 0) ? 1 : 0;            // :82
        }
        __syncthreads();
    }
}

// ---- first-vertex rule -----------------------------------------------------------------------
__global__ void __launch_bounds__(ENV_THREADS) k_env_first_vertex(
    const uint8_t* __restrict__ adj, const int64_t* __restrict__ adj_off,
    const int32_t* __restrict__ node_off, int32_t G, int32_t* __restrict__ action,
    int32_t* __restrict__ dist_counts, int32_t* __restrict__ err) {
    pdl_prologue();
    __shared__ uint8_t dist_s[ENV_WARPS][ENV_MAX_N];
    __shared__ long long best_score[ENV_WARPS];
    __shared__ int32_t best_v[ENV_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (int g = blockIdx.x; g < G; g += gridDim.x) {
        const int32_t v0 = node_off[g], n = node_off[g + 1] - v0;
        if (n > ENV_MAX_N) { if (tid == 0 && err) *err = 1; continue; }
        const uint8_t* A = adj + adj_off[g];
        uint8_t* dist = dist_s[wid];
        long long bs = -1;
        int32_t bv = -1;
        for (int v = wid; v < n; v += ENV_WARPS) {
            for (int u = lane; u < n; u += 32) dist[u] = (u == v) ? 0 : (A[(int64_t)v * n + u] ? 1 : 255);
            __syncwarp();
            for (int d = 1; d < 3; ++d) {
                for (int u = 0; u < n; ++u) {
                    if (dist[u] != d) continue;                  // warp-uniform (smem broadcast)
                    for (int w = lane; w < n; w += 32)
                        if (dist[w] == 255 && A[(int64_t)u * n + w]) dist[w] = (uint8_t)(d + 1);
                }
                __syncwarp();
            }
            int c1 = 0, c2 = 0, c3 = 0;
            for (int u = lane; u < n; u += 32) {
                const int dd = dist[u];
                c1 += (dd == 1); c2 += (dd == 2); c3 += (dd == 3);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                c2 += __shfl_xor_sync(0xffffffffu, c2, o);
                c3 += __shfl_xor_sync(0xffffffffu, c3, o);
            }
            if (lane == 0 && dist_counts) {
                int32_t* dc = dist_counts + (int64_t)(v0 + v) * 3;
                dc[0] = c1; dc[1] = c2; dc[2] = c3;
            }
            const long long score = ((long long)c1 * n + c2) * n + c3;   // sum_i counts[i] * n^(2-i)
            if (score > bs) { bs = score; bv = v; }                      // ascending v: first max kept
            __syncwarp();
        }
        if (lane == 0) { best_score[wid] = bs; best_v[wid] = bv; }
        __syncthreads();
        if (tid == 0) {
            long long s = -1;
            int32_t bvv = -1;
            for (int w = 0; w < ENV_WARPS; ++w)
                if (best_v[w] >= 0 && (best_score[w] > s || (best_score[w] == s && best_v[w] < bvv))) {
                    s = best_score[w];
                    bvv = best_v[w];
                }
            if (action) action[g] = bvv;
        }
        __syncthreads();
    }
}

}  // namespace gcrl

using namespace gcrl;

extern "C" {

int gcrl_env_reset(int32_t* colour, int32_t* max_colour, float* x, const int32_t* node_off,
                   int32_t G, int64_t N, void* stream) {
    GCRL_CHECK_ARG(G >= 0 && N >= 0);
    if (G == 0) return GCRL_OK;
    GCRL_CHECK_ARG(colour && max_colour && node_off);
    int64_t work = N > G ? N : G;
    int blocks = (int)((work + 255) / 256);
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    launch_k(k_env_reset, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, colour, max_colour, x, node_off, G, N);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

int gcrl_env_step(const uint8_t* adj, const int64_t* adj_off, const int32_t* node_off, int32_t G,
                  const int32_t* action, int32_t* colour, int32_t* max_colour, float* reward,
                  int32_t* done, float* x, void* stream) {
    GCRL_CHECK_ARG(G >= 0);
    if (G == 0) return GCRL_OK;
    GCRL_CHECK_ARG(adj && adj_off && node_off && action && colour && max_colour);
    int blocks = G < sm_count() * 16 ? G : sm_count() * 16;
    launch_k(k_env_step, dim3(blocks), dim3(ENV_THREADS), 0, (cudaStream_t)stream, adj, adj_off, node_off, G, action, colour,
             max_colour, reward, done, x, nullptr, nullptr, nullptr);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

int gcrl_env_step_greedy(const uint8_t* adj, const int64_t* adj_off, const int32_t* node_off, int32_t G,
                         const float* q, int32_t* action_out, int32_t* colour, int32_t* max_colour, float* reward,
                         int32_t* done, float* x, void* stream) {
    GCRL_CHECK_ARG(G >= 0);
    if (G == 0) return GCRL_OK;
    GCRL_CHECK_ARG(adj && adj_off && node_off && q && colour && max_colour);
    int blocks = G < sm_count() * 16 ? G : sm_count() * 16;
    launch_k(k_env_step, dim3(blocks), dim3(ENV_THREADS), 0, (cudaStream_t)stream, adj, adj_off, node_off, G, nullptr, colour,
             max_colour, reward, done, x, nullptr, q, action_out);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

int gcrl_env_first_vertex(const uint8_t* adj, const int64_t* adj_off, const int32_t* node_off,
                          int32_t G, int32_t* action, int32_t* dist_counts, void* stream) {
    GCRL_CHECK_ARG(G >= 0);
    if (G == 0) return GCRL_OK;
    GCRL_CHECK_ARG(adj && adj_off && node_off && (action || dist_counts));
    int blocks = G < sm_count() * 16 ? G : sm_count() * 16;
    launch_k(k_env_first_vertex, dim3(blocks), dim3(ENV_THREADS), 0, (cudaStream_t)stream, adj, adj_off, node_off, G, action,
                                                                        dist_counts, nullptr);
    GCRL_LAUNCH_CHECK();
    return GCRL_OK;
}

}  // extern "C"
