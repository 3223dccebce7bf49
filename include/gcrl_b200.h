/*
 * gcrl_b200.h -- C ABI of the B200-native GN Q-network hot path (libgcrl_b200.so).
 *
 * The reference (gpdwatkins/graph_colouring_with_RL) has no FFI of its own: its boundary is
 * the Python API of architecture.py / principal_neighbourhood_aggregation.py / dqn_agent.py /
 * graph_colouring_env.py.  The drop-in Python classes in graph_colouring_with_rl_b200/ keep
 * that API and call the entry points below through ctypes; each entry point cites the
 * reference lines whose arithmetic it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *  - the caller (torch) owns all memory, including workspaces (query *_bytes first);
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs
 *    unless stated;
 *  - return 0 on success, negative gcrl_status otherwise; gcrl_last_error() gives the text;
 *    nothing throws across the ABI;
 *  - not thread-safe per stream; one process per GPU.
 *
 * Internal ("segment") edge layout: edges are permuted destination-major: slot s in
 * [seg_off[v], seg_off[v+1]) holds an edge into vertex v; within a segment slots are ordered
 * by caller edge id; perm[s] = caller edge id, src[s] / dst[s] = endpoint vertex ids.
 */
#ifndef GCRL_B200_H
#define GCRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

typedef enum {
    GCRL_OK = 0,
    GCRL_ERR_INVALID = -1,   /* bad argument (null pointer, negative size, unsupported dim) */
    GCRL_ERR_CUDA = -2,      /* a CUDA runtime call failed */
    GCRL_ERR_WORKSPACE = -3, /* workspace too small */
    GCRL_ERR_INDEX = -4      /* vertex index out of range in edge_index (validated pack only) */
} gcrl_status;

#define GCRL_HID 64          /* architecture.py:123-128: every hidden / output width */
#define GCRL_N_BLOCKS 5      /* architecture.py:157-161 */
#define GCRL_N_AGG 4         /* architecture.py:119 ['mean','min','max','std'] */

/* math mode of the per-edge / per-vertex MLPs */
#define GCRL_MATH_FP32 0     /* CUDA-core FFMA, fp32 operands and accumulate (parity mode) */
#define GCRL_MATH_BF16 1     /* tcgen05 kind::f16, bf16 operands, fp32 accumulate in TMEM */
#define GCRL_MATH_BF16X3 2   /* tcgen05, operands split hi+lo bf16, 3 MMAs per product (near-fp32) */

const char* gcrl_last_error(void);
int gcrl_version(void);
/* number of SMs of the current device (grid sizing) */
int gcrl_sm_count(void);

/* ---- instrumentation ---------------------------------------------------------------------
 * gcrl_launch_count: kernels this library has launched since it was loaded.
 * gcrl_prof_enable(1): bracket every launch of the dominant kernels with CUDA events on the
 * launch stream; gcrl_prof_read syncs on them and returns, per kind, total milliseconds and
 * launch count since the last read (then resets). */
#define GCRL_PROF_EDGE_FWD_FIRST 0  /* block 1: reads edge_attr, writes e_1 */
#define GCRL_PROF_EDGE_FWD 1        /* blocks 2-4: reads e_{l-1}, writes e_l */
#define GCRL_PROF_EDGE_FWD_LAST 2   /* block 5: reads e_4, writes nothing per edge */
#define GCRL_PROF_EDGE_BWD_FIRST 3  /* block 1: reads e_1, d e_1 */
#define GCRL_PROF_EDGE_BWD 4        /* blocks 2-4: reads e_{l-1}, e_l, d e_l; writes d e_{l-1} */
#define GCRL_PROF_EDGE_BWD_LAST 5   /* block 5: reads e_4; writes d e_4 */
#define GCRL_PROF_PNA_FWD 6
#define GCRL_PROF_PNA_BWD 7
#define GCRL_PROF_KINDS 8
int64_t gcrl_launch_count(void);
int gcrl_prof_enable(int on);
int gcrl_prof_read(double* ms, int64_t* launches, int n_kinds);

/* Self-test of the tcgen05 / TMA / TMEM plumbing the edge kernels use (one CTA, bf16x3):
 * D [128,64] = A[row0:row0+128] * W^T, DW [64,64] = A[rows]^T * A2[rows]; copy_out (may be NULL)
 * receives a TMA store of the A tile.  A, A2, copy_out: [rows_total,64] fp32; W [64,64]. */
int gcrl_selftest_tc(const float* A, const float* A2, const float* W, int64_t rows_total, int row0,
                     float* D, float* DW, float* copy_out, int use_tma, void* stream);

/* ---- parameter layout ------------------------------------------------------------------
 * One flat fp32 buffer in state_dict order (architecture.py:157-167): for block l=1..5
 * message_mlp.0.{weight,bias}, message_mlp.2.{weight,bias}, update_mlp.0.{weight,bias},
 * update_mlp.2.{weight,bias}; then fc1, fc2, fc3 {weight,bias}.  46 tensors; 198 529 floats
 * for (vertex_dim, edge_dim, global_dim) = (2, 1, 0).
 * offsets_host[47]: start of each tensor, last entry = total. */
int gcrl_param_layout(int vertex_dim, int edge_dim, int global_dim, int64_t* offsets_host);

/* ---- graph batching (replaces PyG Batch.from_data_list + MessagePassing.__collect__ index
 * plumbing, dqn_agent.py:55-56, architecture.py:84) ---------------------------------------*/
size_t gcrl_pack_workspace_bytes(int64_t n_vertices, int64_t n_edges);
/* edge_index: [2,E] int64 row-major (row 0 = source j, row 1 = target i), vertex ids already
 * batch-offset.  Outputs: seg_off[N+1], src[E], dst[E], perm[E] (all int32).
 * err_flag (device int32, may be NULL): set non-zero if an id is outside [0,N). */
int gcrl_pack(const int64_t* edge_index, int64_t n_edges, int64_t n_vertices,
              int32_t* seg_off, int32_t* src, int32_t* dst, int32_t* perm,
              int32_t* err_flag, void* workspace, size_t workspace_bytes, void* stream);
/* Complete-graph batches straight from dense adjacency (graph_colouring_env.py:31-44: every
 * ordered pair is an edge, attr -1 if present in the target graph, else 0).
 * adj: bytes, graph g's n_g x n_g matrix at adj_off[g]; node_off[G+1] int32.
 * Writes seg_off/src/dst, perm (= position in the canonical (0,1),(1,0),(0,2)... caller order
 * of utils.py:304-310) and eattr[E] (-1/0) in internal order. */
size_t gcrl_pack_dense_workspace_bytes(int32_t n_graphs);
int gcrl_pack_dense(const uint8_t* adj, const int64_t* adj_off, const int32_t* node_off,
                    int32_t n_graphs, int64_t n_vertices, int64_t n_edges,
                    int32_t* seg_off, int32_t* src, int32_t* dst, int32_t* perm, float* eattr,
                    void* workspace, size_t workspace_bytes, void* stream);
/* out[s, :] = in[perm[s], :] (to_internal != 0) or out[perm[s], :] = in[s, :] (== 0) */
int gcrl_permute_rows(const float* in, float* out, const int32_t* perm, int64_t n_rows,
                      int32_t width, int to_internal, void* stream);

/* ---- PNA aggregation, standalone (principal_neighbourhood_aggregation.py:41-64;
 * architecture.py:62-66) ------------------------------------------------------------------
 * msg: [E,F] fp32 in INTERNAL order.  out: [N, 4F] = [mean | min | max | std] with
 * std = sqrt(relu(mean(x^2) - mean(x)^2) + 1e-5); empty segment -> (0,0,0,sqrt(1e-5)).
 * arg_min / arg_max: [N,F] int32 internal slot of the first extremum (may be NULL).
 * var_out: [N,F] raw mean(x^2) - mean(x)^2 before the relu (may be NULL); the backward needs
 * its sign for relu'. */
int gcrl_pna_forward(const float* msg, const int32_t* seg_off, int64_t n_vertices,
                     int64_t n_edges, int32_t width, float* out, int32_t* arg_min,
                     int32_t* arg_max, float* var_out, void* stream);
/* d_msg[E,F] from d_out[N,4F] (gradient routing of torch_scatter: mean -> 1/deg to every edge,
 * min/max -> the arg edge only, std through mean and mean-of-squares where var > 0). */
int gcrl_pna_backward(const float* msg, const float* out, const float* var, const float* d_out,
                      const int32_t* seg_off, const int32_t* dst, const int32_t* arg_min,
                      const int32_t* arg_max, int64_t n_vertices, int64_t n_edges,
                      int32_t width, float* d_msg, void* stream);
/* single reductions for AGGREGATORS['sum'|'mean'|'min'|'max'|'var'|'std'] (pna.py:67-74):
 * which = 0 sum, 1 mean, 2 min, 3 max, 4 var, 5 std.  out [N,F]. */
int gcrl_segment_reduce(const float* msg, const int32_t* seg_off, int64_t n_vertices,
                        int64_t n_edges, int32_t width, int which, float* out, void* stream);

/* ---- GN Q-network forward (architecture.py:179-199; blocks :80-99) -----------------------
 * x [N,vertex_dim]; eattr [E,edge_dim] INTERNAL order; gfeat [G,global_dim] or NULL;
 * vgraph [N] int32 (vertex -> graph, only read when global_dim > 0).
 * stash: NULL for inference, else gcrl_gn_stash_bytes() bytes holding the activations the
 * backward pass needs.  h_out [N,64], q_out [N] (fc3 has one output, architecture.py:129). */
size_t gcrl_gn_stash_bytes(int64_t n_vertices, int64_t n_edges);
size_t gcrl_gn_forward_workspace_bytes(int64_t n_vertices, int64_t n_edges, int with_stash);
int gcrl_gn_forward(const float* params, int vertex_dim, int edge_dim, int global_dim,
                    int64_t n_vertices, int64_t n_edges, const int32_t* seg_off,
                    const int32_t* src, const int32_t* dst, const float* x, const float* eattr,
                    const float* gfeat, const int32_t* vgraph, void* stash, float* h_out,
                    float* q_out, void* workspace, size_t workspace_bytes, int math_mode,
                    void* stream);
/* One GNNBlock (architecture.py:98-99): msg_out [E,64] INTERNAL order, upd_out [N,64].
 * block = 1..5 selects the weights; x is [N, vertex_dim of that block]. */
int gcrl_gn_block_forward(const float* params, int vertex_dim, int edge_dim, int global_dim,
                          int block, int64_t n_vertices, int64_t n_edges,
                          const int32_t* seg_off, const int32_t* src, const int32_t* dst,
                          const float* x, const float* eattr, float* msg_out, float* upd_out,
                          void* workspace, size_t workspace_bytes, int math_mode, void* stream);

/* ---- GN backward (autograd of dqn_agent.py:219 through architecture.py:179-199) ----------
 * d_q [N]: gradient w.r.t. q_out.  d_h [N,64] or NULL: gradient w.r.t. h_out.
 * grad: flat buffer laid out like params; overwritten when accumulate == 0, else += . */
size_t gcrl_gn_backward_workspace_bytes(int64_t n_vertices, int64_t n_edges);
int gcrl_gn_backward(const float* params, int vertex_dim, int edge_dim, int global_dim,
                     int64_t n_vertices, int64_t n_edges, const int32_t* seg_off,
                     const int32_t* src, const int32_t* dst, const float* x,
                     const float* eattr, const float* gfeat, const int32_t* vgraph,
                     const void* stash, const float* d_q, const float* d_h, float* grad,
                     int accumulate, void* workspace, size_t workspace_bytes, int math_mode,
                     void* stream);

/* ---- action scoring (dqn_agent.py:131-134) -----------------------------------------------
 * per graph g: action[g] = first argmax of q over vertices with colour < 0 (uncoloured);
 * -1 if none.  q_best[g] (may be NULL) = that maximum (-inf if none). */
int gcrl_score_argmax(const float* q, const int32_t* colour, const int32_t* node_off,
                      int32_t n_graphs, int32_t* action, float* q_best, void* stream);

/* ---- DQN loss (dqn_agent.py:178-215) -----------------------------------------------------
 * q_cur [N] behaviour-net q on current states, q_next [N] target-net q on next states (same
 * vertex layout), next_colour [N] (>= 0 = assigned after the step), action[G] local vertex
 * index, reward[G], done[G].  Writes q_expected[G], q_target[G], d_q[N] (2/B_total *
 * (q_exp - q_tgt) at the action vertex, 0 elsewhere) and adds sum_g (q_exp-q_tgt)^2 / B_total
 * into loss[0] (caller zeroes it).  B_total = global minibatch size (all ranks).  A graph whose
 * action is outside [0, n_g) contributes nothing (q_expected / q_target = NaN for it) and sets
 * *err = g + 1 (err may be NULL; caller zeroes it): no out-of-bounds access on a sentinel action. */
int gcrl_td_target_loss(const float* q_cur, const float* q_next, const int32_t* next_colour,
                        const int32_t* node_off, int32_t n_graphs, const int32_t* action,
                        const float* reward, const int32_t* done, float gamma,
                        int32_t batch_total, float* q_expected, float* q_target, float* d_q,
                        float* loss, int32_t* err, void* stream);

/* ---- fused Adam + soft target update (architecture.py:173 torch.optim.Adam defaults;
 * dqn_agent.py:220,223,226-229) ------------------------------------------------------------
 * step is the 1-based Adam step count; target may be NULL (Adam only); tau as in :228. */
int gcrl_adam_soft_update(float* params, const float* grad, float* exp_avg, float* exp_avg_sq,
                          float* target, int64_t n, float lr, float beta1, float beta2,
                          float eps, int64_t step, float tau, void* stream);
/* t <- tau*b + (1-tau)*t (dqn_agent.py:226-229; tau = 1 copies, :100) */
int gcrl_soft_update(const float* behaviour, float* target, int64_t n, float tau, void* stream);

/* ---- colouring environment (graph_colouring_env.py:24-61,64-90,102-109,144-191;
 * utils.py:285-293) ------------------------------------------------------------------------
 * adj as in gcrl_pack_dense.  colour[N] int32 (-1 = uncoloured), max_colour[G] int32.
 * step: for every graph with action[g] >= 0: first-fit colour the vertex, then first-fit
 * every uncoloured vertex that has no uncoloured neighbour; reward[g] = -(new colours),
 * done[g] = no uncoloured vertex left; x (may be NULL) [N,2] gets x[:,1] = colour
 * (GenerateData_v1, :124-141).  Graphs with action[g] < 0 are untouched (reward 0). */
int gcrl_env_reset(int32_t* colour, int32_t* max_colour, float* x, const int32_t* node_off,
                   int32_t n_graphs, int64_t n_vertices, void* stream);
int gcrl_env_step(const uint8_t* adj, const int64_t* adj_off, const int32_t* node_off,
                  int32_t n_graphs, const int32_t* action, int32_t* colour,
                  int32_t* max_colour, float* reward, int32_t* done, float* x, void* stream);
/* first-vertex rule of the rollout harness (test_policy_functions.py:46-47; counts built at
 * utils.py:522-533,559-565): argmax_v deg*n^2 + #dist2*n + #dist3, first max. */
/* greedy rollout step: action scoring fused into the environment step (test_policy_functions.py:56 +
 * dqn_agent.py:131-134 + graph_colouring_env.py:64-90 in one launch).  Per graph the action is the first
 * maximum of q [N] over its uncoloured vertices (-1 = none left: the graph is left alone), written to
 * action_out[G] (may be NULL); then exactly gcrl_env_step. */
int gcrl_env_step_greedy(const uint8_t* adj, const int64_t* adj_off, const int32_t* node_off,
                         int32_t n_graphs, const float* q, int32_t* action_out, int32_t* colour,
                         int32_t* max_colour, float* reward, int32_t* done, float* x, void* stream);

int gcrl_env_first_vertex(const uint8_t* adj, const int64_t* adj_off, const int32_t* node_off,
                          int32_t n_graphs, int32_t* action, int32_t* dist_counts,
                          void* stream);

/* ---- device-resident replay ring (dqn_agent.py:8-65 ReplayBuffer.store_transition /
 * sample_buffer) ---------------------------------------------------------------
 * Edges are static per graph, so a transition is (graph id, colours before the step, colours
 * after, action, reward, done).  The ring arrays are caller-owned: r_graph/r_action/r_done
 * int32 [capacity], r_reward f32 [capacity], r_cur/r_next int16 [capacity, stride] (-1 padded,
 * stride >= the largest vertex count).  The host chooses slots (ring index = count % capacity,
 * dqn_agent.py:33) and samples them (np.random.choice with replacement, :48).
 * store: environment e of a batched env (node_off [n_envs+1]) with slot[e] >= 0 writes its
 *   transition to that slot; slot[e] < 0 skips it.
 * gather: minibatch position b takes ring slot slots[b]; the graph's adjacency bytes are copied
 *   from the pool (pool_adj [sum n^2], pool_adj_off int64 [P], pool_n int32 [P], n_max = largest
 *   pool_n) to out_adj + out_adj_off[b]; colours are widened to int32 at out_node_off[b]; the
 *   outputs are exactly the inputs of gcrl_pack_dense / gcrl_gn_forward / gcrl_td_target_loss.
 *   out_node_off [batch+1] / out_adj_off [batch+1] are computed by the caller from its host
 *   mirror of the slots' vertex counts; a mismatch with the ring sets *err = b+1 (err may be
 *   NULL) and copies nothing for that position. */
int gcrl_replay_store(int32_t n_envs, const int32_t* slot, const int32_t* graph_id,
                      const int32_t* node_off, const int32_t* cur_colour,
                      const int32_t* next_colour, const int32_t* action, const float* reward,
                      const int32_t* done, int32_t stride, int32_t* r_graph, int16_t* r_cur,
                      int16_t* r_next, int32_t* r_action, float* r_reward, int32_t* r_done,
                      void* stream);
int gcrl_replay_gather(int32_t batch, const int32_t* slots, const int32_t* r_graph,
                       const int16_t* r_cur, const int16_t* r_next, const int32_t* r_action,
                       const float* r_reward, const int32_t* r_done, int32_t stride,
                       const uint8_t* pool_adj, const int64_t* pool_adj_off,
                       const int32_t* pool_n, int32_t n_max, const int32_t* out_node_off,
                       const int64_t* out_adj_off, uint8_t* out_adj, int32_t* out_cur,
                       int32_t* out_next, int32_t* out_action, float* out_reward,
                       int32_t* out_done, int32_t* err, void* stream);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* GCRL_B200_H */
