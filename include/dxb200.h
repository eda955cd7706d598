/*
 * dxb200.h -- C ABI of libdxb200.so, the B200 (sm_100a) batched heuristic-search engine.
 *
 * The reference (forestagostinelli/deepxube) is 100% Python and has no FFI boundary; its boundary for
 * this path is a set of Python protocols.  Each entry point below names the reference interface it
 * stands behind (file:line relative to the reference root).  The Python layer in deepxube_b200/
 * binds these with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions: every function returns DX_OK (0) or a negative error code; dx_last_error() gives the
 * message for the calling thread's last failure.  All pointers are HOST pointers to caller-owned
 * memory unless the name ends in _dev.  The engine owns all device memory and one CUDA stream; calls
 * on one engine must not overlap (the reference's PathFind objects are single-threaded too).
 *
 * State rows are packed uint8:
 *   cube3      54 bytes, sticker colours 0..5          (deepxube/domains/cube3.py:14-29, 502-503)
 *   npuzzle.d  d*d bytes, tile numbers, 0 = blank      (deepxube/domains/npuzzle.py:18-33, 63-73)
 *   grid.d     2 bytes (robot_x, robot_y)              (deepxube/domains/grid.py:24-35)
 * Goal rows have the same layout as state rows (for npuzzle a goal byte equal to d*d is a wildcard,
 * npuzzle.py:154-158).
 */
#ifndef DXB200_H
#define DXB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DX_OK 0
#define DX_ERR_ARG (-1)        /* bad argument                                   */
#define DX_ERR_CUDA (-2)       /* CUDA runtime failure                           */
#define DX_ERR_CAPACITY (-3)   /* node store / OPEN / batch capacity exceeded    */
#define DX_ERR_STATE (-4)      /* call not valid in the engine's current state   */
#define DX_ERR_UNSUPPORTED (-5)

enum dx_domain { DX_DOMAIN_CUBE3 = 0, DX_DOMAIN_NPUZZLE = 1, DX_DOMAIN_GRID = 2,
                 DX_DOMAIN_LIGHTSOUT = 3 /* deepxube/domains/lightsout.py: dim*dim 0/1 tiles, dim*dim actions */ };

/* graph_v = batch weighted A* (deepxube/pathfinding/graph_search.py:182-194),
 * graph_q = batch weighted Q* (graph_search.py:197-209). */
enum dx_algo { DX_ALGO_GRAPH_V = 0, DX_ALGO_GRAPH_Q = 1,
               DX_ALGO_BEAM_V = 2 /* beam_v (deepxube/pathfinding/beam_search.py:106-125, 175-196): batch_size = beam size */,
               DX_ALGO_BEAM_Q = 3 /* beam_q (beam_search.py:117-125, 199-214, 217-229): edges ranked by their q-value */ };

/* Where heuristic values come from:
 *   ZERO      the all-zero default functions (deepxube/base/pathfind_fns.py:205-210, 235-243)
 *   HOST      any HeurVFn / HeurQFn callable on the host (base/pathfind_fns.py:20-31): the iteration is
 *             split in dx_step_begin / dx_step_end around the host call
 *   NET_FP32  resnet_fc on the device, fp32 CUDA-core GEMMs (parity mode)
 *   NET_BF16  resnet_fc on the device, tcgen05 bf16 tensor-core GEMMs with fp32 accumulation */
enum dx_heur_mode { DX_HEUR_ZERO = 0, DX_HEUR_HOST = 1, DX_HEUR_NET_FP32 = 2, DX_HEUR_NET_BF16 = 3 };

typedef struct dx_engine dx_engine;

typedef struct dx_config {
    int32_t device;          /* CUDA device ordinal                                                      */
    int32_t domain;          /* enum dx_domain                                                           */
    int32_t domain_dim;      /* npuzzle: side d (2..15); grid: side d (1..255); cube3: ignored           */
    int32_t algo;            /* enum dx_algo                                                             */
    int32_t heur_mode;       /* enum dx_heur_mode                                                        */
    int32_t max_instances;   /* capacity: instances alive in the engine at once                          */
    int32_t max_pop_total;   /* capacity: sum over instances of their batch sizes (nodes popped / iter)  */
    int32_t max_backups;     /* 0, or the capacity (records) of the Bellman-backup log: the engine then evaluates the
                                heuristic of EVERY generated child before the CLOSED filter, like the reference
                                (base/pathfinding.py:362-371), and logs the backup of every expanded node; graph_q: logs the label of every popped edge */
    int64_t max_nodes;       /* capacity: nodes stored (A*: kept children; Q*: all generated children)   */
    int64_t max_open;        /* capacity: OPEN entries (A*: <= max_nodes; Q*: edges); 0 = derive         */
    int32_t reserved[8];
} dx_config;

/* resnet_fc.<H>H_<B>B[_bn] (deepxube/nnets/resnet_fc.py:19-52).  The weight blob is float32, in this
 * order (all nn.Linear weights row-major [out, in] exactly as in the state_dict):
 *   heur.0.weight [H, in_count*one_hot_depth], heur.0.bias [H],
 *   for each block b, for each layer j in {0,1}:
 *       ...layers.j.0.weight [H,H], ...layers.j.0.bias [H],
 *       if batch_norm: ...layers.j.1.weight(gamma) [H], .bias(beta) [H], .running_mean [H], .running_var [H]
 *   heur.2.weight [out_dim, H], heur.2.bias [out_dim]
 * Eval-mode batch norm (eps 1e-5) is folded into the preceding Linear at load time. */
typedef struct dx_net_desc {
    int32_t in_count;        /* integer inputs per state (cube3 54, npuzzle d*d, grid 4)                 */
    int32_t one_hot_depth;   /* cube3 6, npuzzle d*d, grid d                                             */
    int32_t res_dim;         /* H                                                                        */
    int32_t num_blocks;      /* B                                                                        */
    int32_t out_dim;         /* 1 (heurv) or number of actions (heurq_fixout)                            */
    int32_t batch_norm;      /* 0 / 1                                                                    */
    int32_t act_depth;       /* 0, or the number of actions for an action-as-input Q network (heurq_in,
                                deepxube/pathfind_fns/fns_core.py:99-133): the action index is one more input with this
                                one-hot depth, out_dim is 1 and the network runs once per (state, action)         */
} dx_net_desc;

/* Per-instance view, the fields of Instance / InstanceGraph the callers read
 * (deepxube/base/pathfinding.py:132-184, pathfinding/graph_search.py:20-46). */
typedef struct dx_inst_status {
    double lb;                    /* InstanceGraph.lb                                                    */
    double ub;                    /* InstanceGraph.ub (inf while no goal node)                           */
    int64_t num_nodes_generated;  /* Instance.num_nodes_generated                                        */
    int64_t open_size;            /* frontier_size()                                                     */
    int64_t goal_node;            /* node id of goal_node, -1 if none                                    */
    int32_t itr;                  /* Instance.itr                                                        */
    int32_t finished;             /* Instance.finished()                                                 */
    int32_t num_curr;             /* len(_nodes_curr): nodes that the next step will pop/expand          */
    int32_t reserved;
} dx_inst_status;

/* Device-side counters of the last dx_run / step, for bench.py's accounting. */
typedef struct dx_counters {
    int64_t nodes_stored;         /* nodes in the node store                                             */
    int64_t heur_evals;           /* states evaluated by the heuristic (device net or host callback)     */
    int64_t children_generated;   /* sum of num_nodes_generated over all instances                       */
    int64_t iterations;           /* PathFind.itr                                                        */
    int64_t kernel_launches;      /* CUDA kernels launched by the engine since creation                  */
    int64_t backups_logged;       /* log slots in use (an upper bound of what dx_drain_backups returns)  */
    int64_t hidden_gemm_launches; /* H x H GEMM launches per network evaluation: 2 * num_blocks, or one less when the first
                                     hidden layer is computed by the input-layer launch (bf16 tensor-core path)             */
    int64_t reserved[1];
} dx_counters;

const char* dx_last_error(void);
int dx_version(void);

/* PathFind.__init__ (deepxube/base/pathfinding.py:219-228) + GraphSearch.__init__ (graph_search.py:87-92). */
int dx_engine_create(const dx_config* cfg, dx_engine** out);
int dx_engine_destroy(dx_engine* e);
/* Drops all instances and search state; keeps buffers and loaded weights (a fresh PathFind per
 * problem instance, deepxube/_solve.py:146-147, without re-allocating). */
int dx_engine_reset(dx_engine* e);

/* load_nnet + nnet.eval() (deepxube/pytorch/nnet_utils.py:44-62, factories/pathfind_fns_factory.py:98-104).
 * Calling it again on the same engine overwrites the parameters in place (same network shape required): what an RL
 * loop does between two batches when the heuristic / target network has been updated (updaters/updater_rl.py). */
int dx_load_weights(dx_engine* e, const dx_net_desc* desc, const float* blob, int64_t n_floats);

/* make_instances + add_instances (graph_search.py:94-109,142-145,161-166; base/pathfinding.py:242-243).
 * states [n, state_bytes], goals [n, state_bytes]; batch_sizes / weights per instance (may be NULL:
 * 1 / 1.0).  Root values are computed on the device for ZERO / NET modes; in HOST mode the caller
 * passes them in root_vals ([n] for graph_v, [n, A] for graph_q; may be NULL for graph_v). */
int dx_add_instances(dx_engine* e, const uint8_t* states, const uint8_t* goals, int32_t n,
                     const int32_t* batch_sizes, const double* weights, const float* root_vals);

/* PathFind.remove_instances / remove_finished_instances (base/pathfinding.py:300-314): the listed instance slots (the
 * order they were added in since the last reset) stop searching -- no further pops, expansions or OPEN entries -- and
 * no longer count against max_pop_total.  Their slots, nodes and CLOSED entries are kept until dx_engine_reset, so
 * dx_get_path / dx_get_status stay valid for them; max_instances bounds the instances added between two resets. */
int dx_remove_instances(dx_engine* e, const int32_t* slots, int32_t n);

/* PathFind.step() repeated (base/pathfinding.py:338-400, 471-525): runs until every instance is
 * finished() or max_iters steps were taken.  ZERO / NET modes only.  *iters_done = steps taken. */
int dx_run(dx_engine* e, int32_t max_iters, int32_t* iters_done);

/* HOST heuristic mode: one step() split around the host callable.
 *   dx_step_begin   pop/is_solved/record_goal/expand/closed filter (A*) or filter/edges/push/pop/next_state (Q*);
 *                   *n_pending = number of states that need values (graph_v: the children that passed the CLOSED filter;
 *                   with max_backups > 0 the filter runs in dx_step_end and EVERY generated child is pending)
 *   dx_get_pending  copies those states [n_pending, state_bytes] and their goals to the host
 *   dx_step_end     takes h [n_pending] (graph_v) or q [n_pending, A] (graph_q), float32, already clipped,
 *                   and finishes the step (costs, push, pop, lb, finished, itr) */
int dx_step_begin(dx_engine* e, int64_t* n_pending);
int dx_get_pending(dx_engine* e, uint8_t* states, uint8_t* goals, int64_t cap);
int dx_step_end(dx_engine* e, const float* vals, int64_t n);

int dx_num_instances(dx_engine* e);
int dx_all_finished(dx_engine* e, int32_t* out);
int dx_get_status(dx_engine* e, dx_inst_status* out, int32_t cap);
int dx_get_counters(dx_engine* e, dx_counters* out);

/* Training targets from the search (graph_v / graph_q engines created with max_backups > 0), one record per node
 * expanded (graph_v) or edge popped (graph_q) since the last drain, in the order of the `popped` lists that
 * PathFind.step() returns (instances in order, pop order within; base/pathfinding.py:399, 525):
 *   graph_v  Node.bellman_backup() (base/pathfinding.py:41-47, UpdateHeurVRLKeepGoalABC._get_labels_no_rb,
 *            updaters/updater_rl.py:143-158): 0 if the node is solved else min over its children of
 *            (1.0 + heuristic(child)); actions[] = -1
 *   graph_q  Node.backup_act(action) (base/pathfinding.py:66-77, UpdateHeurQRLKeepGoalABC._get_labels_no_rb,
 *            updaters/updater_rl.py:169-189): 0 if the edge's node is solved else 1.0 + min_a q(node_next, a);
 *            states[] = the edge's node, actions[] = the edge's action
 * float64 like the reference's Python floats.  states [cap, state_bytes], actions [cap], backups [cap], insts [cap] (any
 * may be NULL); *n = records written.  DX_ERR_ARG with *n = the count when cap is too small (nothing is drained).  The log is
 * emptied on success.  An iteration takes one log slot per popped node / edge (graph_v: slots of nodes popped by an
 * instance in the iteration it finished are dropped at the drain): size max_backups accordingly.
 *
 * mode selects the label of the updater's options (UpRLArgs, base/updater.py:669-672), computed over ALL records in the log at
 * the time of the drain -- drain when the instances' lives end, like the reference:
 *   DX_BACKUP_BELLMAN   as above
 *   DX_BACKUP_UB_SOLNS  ub_heur_solns: every solved expanded node (graph_q: solved node of a popped edge) bounds the value of
 *                       each ancestor by the path cost down to it (Node.upper_bound_parent_path, base/pathfinding.py:49-53;
 *                       updater_rl.py:148-152, 172-176); graph_q labels then use the bound of node_next where there is one
 *   DX_BACKUP_LHBL      lhbl: Node.tree_backup() from the roots (base/pathfinding.py:55-64; updater_rl.py:153-155, 177-179):
 *                       0 if solved else min over the children (graph_q: over the popped edges) of
 *                       (1.0 + (child was expanded ? its tree backup : max(heuristic(child), 0))); graph_q labels =
 *                       1.0 + tree backup of node_next */
enum dx_backup_mode { DX_BACKUP_BELLMAN = 0, DX_BACKUP_UB_SOLNS = 1, DX_BACKUP_LHBL = 2 };
int dx_drain_backups(dx_engine* e, int32_t mode, uint8_t* states, int32_t* actions, double* backups, int32_t* insts, int64_t cap,
                     int64_t* n);

/* Hindsight experience replay, goal relabelling (UpdateHER._get_her_goals, deepxube/base/updater.py:485-530): for every instance
 * slot, the state of the DEEPEST node that has children among the records currently in the backup log -- expanded nodes (graph_v) /
 * nodes of popped edges (graph_q); the reference keeps the first node of the largest depth in `root.get_all_descendants()` order
 * (breadth first, children in action order), i.e. the lexicographically smallest action path among the deepest -- or the root's state
 * when nothing deeper was expanded.  states [cap >= instances, state_bytes], depths [cap] (may be NULL).  The caller relabels with
 * domain.sample_goal_from_state(start, deepest) for the instances that did not finish solved.  Call before dx_drain_backups. */
int dx_backup_deepest(dx_engine* e, uint8_t* states, int32_t* depths, int32_t cap);

/* get_path (base/pathfinding.py:93-120): device-side parent walk from goal_node.  actions [cap],
 * states_on_path [(cap+1), state_bytes] (may be NULL).  *len = number of actions (path length). */
int dx_get_path(dx_engine* e, int32_t inst, int32_t* actions, uint8_t* states_on_path, int32_t cap, int32_t* len);

/* Instance.get_nodes() (base/pathfinding.py:147-148): the nodes the next step will expand, in pop order.
 * states [cap, state_bytes], g [cap], h [cap]; *n = count. */
int dx_get_curr_nodes(dx_engine* e, int32_t inst, uint8_t* states, float* g, float* h, int32_t cap, int32_t* n);

/* Stand-alone domain kernels on host buffers (the domain mixin API):
 *   ActsEnum.expand (base/domain.py:288-312): children [n*A, state_bytes], child i*A+a
 *   Domain.next_state (cube3.py:583-596, npuzzle.py:87-110, grid.py:74-86)
 *   Domain.is_solved (cube3.py:514-517, npuzzle.py:154-158, grid.py:68-69) */
int dx_domain_info(int32_t domain, int32_t dim, int32_t* state_bytes, int32_t* num_actions, int32_t* in_count,
                   int32_t* one_hot_depth);
int dx_expand(int32_t device, int32_t domain, int32_t dim, const uint8_t* states, int64_t n, uint8_t* children);
int dx_next_state(int32_t device, int32_t domain, int32_t dim, const uint8_t* states, const int32_t* actions, int64_t n,
                  uint8_t* next);
int dx_is_solved(int32_t device, int32_t domain, int32_t dim, const uint8_t* states, const uint8_t* goals, int64_t n,
                 uint8_t* solved);

/* HeurVFn / HeurQFn with the loaded network (base/pathfind_fns.py:196-227, pathfind_fns/fns_core.py:60-90,
 * pytorch/nnet_utils.py:73-111): out [n, out_dim] float32, clipped at 0.  Uses the engine's heur_mode
 * precision (NET_FP32 or NET_BF16). */
int dx_heur_eval(dx_engine* e, const uint8_t* states, const uint8_t* goals, int64_t n, float* out);

/* Value-iteration / Q-learning targets of replay-buffer samples, computed with the loaded (target) network:
 *   graph_v engines -- UpdateHeurVRL._get_labels_rb (deepxube/updaters/updater_rl.py:41-69): every state is expanded,
 *       out = min_a (1.0 + h(child_a)) * (not solved); solved == NULL: is_solved(state, goal) is computed on the device
 *   graph_q engines -- UpdateHeurQRL._get_labels_rb (:102-120): `states` are the NEXT states of the sampled edges,
 *       out = (1.0 + min_a q(next, a)) * (not solved); solved [n] = is_solved of each edge's node (NULL: none)
 * states / goals [n, state_bytes], out [n] float64 (the reference computes these in numpy float64). */
int dx_vi_targets(dx_engine* e, const uint8_t* states, const uint8_t* goals, const uint8_t* solved, int64_t n, double* out);

/* Bench / profiling helpers: same network evaluation with inputs and outputs resident on the device.
 * dev_states: [n, row_bytes] engine row layout (see dx_row_bytes); returns milliseconds of device time
 * measured with CUDA events on the engine's stream over `reps` back-to-back evaluations. */
int dx_row_bytes(dx_engine* e, int32_t* row_bytes);
int dx_bench_heur(dx_engine* e, int64_t n, int32_t reps, float* ms_total);
/* Device time (ms, CUDA events on the engine stream) spent per pipeline stage since the last reset:
 * [0] expand, [1] closed filter, [2] scan+scatter, [3] heuristic, [4] sort, [5] merge/pop, [6] bookkeeping.
 * Only collected while profiling is on (event records, no syncs): level 1 times every stage and the GEMM groups,
 * level 2 only the hidden-layer GEMM group ([7]; one interval per network evaluation, lowest overhead), 0 turns it off. */
int dx_set_profiling(dx_engine* e, int32_t level);
int dx_get_stage_ms(dx_engine* e, float* ms, int32_t cap);
/* Same indexing plus [7] hidden-layer GEMM launches and [8] input-layer launches (both nested inside [3]):
 * number of timed intervals per entry, so that ms / calls = average launch duration. */
int dx_get_stage_calls(dx_engine* e, int64_t* calls, int32_t cap);
/* Device time (ms, CUDA events on the engine stream) of the last dx_run call. */
int dx_last_run_ms(dx_engine* e, float* ms);

/* Debug: per-role barrier-wait cycle counters of the tensor-core GEMM, summed over CTAs since the last call
 * (all zero unless the library was built with -DTC_PROFILE): [0] MMA<-operands, [1] MMA<-accumulator free,
 * [2] TMA producer<-slot free, [3] epilogue<-accumulator ready, [4] epilogue<-residual, [5] epilogue<-store drain,
 * [6] epilogue warp total, [7] MMA thread total, [8] epilogue<-tcgen05.ld, [9] epilogue math + staging,
 * [10] epilogue proxy fence + warp sync, [11] epilogue TMA issue; out8 must hold 16 entries. */
int dx_debug_tc_profile(dx_engine* e, uint64_t* out8);

/* ---- HDA*-style hash-distributed single search (one instance over `world` ranks, graph_v, device heuristic) ----
 * Not present in the reference (its only multi-GPU mechanism is nn.DataParallel,
 * deepxube/factories/pathfind_fns_factory.py:99-104); semantics per rank are those of graph_search.py with
 * lb / ub / finished() evaluated on globally reduced quantities.  owner(state) = hash(state) mod world.
 * Per iteration:  dx_hda_expand -> all-to-all of the record buffers (host layer: NCCL) -> dx_hda_absorb ->
 * reduce {sum k_popped, min f_first, min ub, sum children_sent} over ranks -> dx_hda_finish.
 * Record = state row + g + parent global id ((rank << 28) | local id) + action; dx_hda_record_bytes gives its size. */
typedef struct dx_hda_local {
    int64_t k_popped;        /* nodes this rank popped for the next iteration                     */
    double f_first;          /* f of its first popped node (inf if none)                          */
    double ub;               /* g of the best solved node this rank has popped so far (inf if none) */
    int64_t goal_local;      /* local id of that node, -1 if none                                 */
    int64_t n_kept;          /* received children that passed the CLOSED filter                   */
    int64_t open_size;       /* local OPEN entries                                                */
    int64_t children_sent;   /* children this rank generated this iteration                       */
} dx_hda_local;
int dx_hda_config(dx_engine* e, int32_t rank, int32_t world);
int dx_hda_record_bytes(dx_engine* e, int32_t* rec_bytes);
int dx_hda_owner(dx_engine* e, const uint8_t* state, int32_t* owner);
int dx_hda_add_instance(dx_engine* e, const uint8_t* state, const uint8_t* goal, int32_t batch_local, double weight);
int dx_hda_expand(dx_engine* e, void* send_dev, int64_t send_cap_bytes, int64_t* counts);
int dx_hda_absorb(dx_engine* e, const void* recv_dev, int64_t n_records, dx_hda_local* out);
int dx_hda_finish(dx_engine* e, int64_t k_total, double f_first_min, double ub_min, int64_t children_total, int32_t* finished);
int dx_hda_node(dx_engine* e, int64_t local_id, uint8_t* state, uint32_t* parent_global, int32_t* action, float* g);
/* Device-resident form of the same iteration -- no host synchronisation inside, every launch on the engine's stream
 * (dx_engine_set_stream lets the caller's collectives share it):
 *   dx_hda_send(e, send_dev, seg_records)        expand the local popped list, group the children by owner and write W fixed-size
 *                                                segments [u32 count, 12 B pad, seg_records records] (segment d is for rank d)
 *   -- all-to-all of the segments with EQUAL splits (16 + seg_records * record_bytes each) --
 *   dx_hda_receive(e, recv_dev, seg_records, local_dev)   unpack the W received segments (source-major), CLOSED filter, heuristic,
 *                                                push / pop; local_dev[DX_HDA_LOCAL_DOUBLES] (device doubles) = {k_popped, f_first, ub,
 *                                                goal_local, children generated, device error flags, children received, children kept}
 *   -- all-gather of local_dev over the ranks --
 *   dx_hda_reduce(e, all_dev, summary_dev)       lb / ub / finished() from all_dev[DX_HDA_LOCAL_DOUBLES * W] on every rank; summary_dev[8]
 *                                                (device doubles) = {finished, iterations, nodes generated, ub, goal rank, goal local id,
 *                                                device error flags OR-ed over all ranks, max over ranks of children received}
 * A segment or receive-capacity overflow raises the device error flag on that rank; the flag travels in local_dev, so EVERY rank
 * sees it in the summary of the same iteration and stops (a failing rank keeps publishing empty segments; nobody is left waiting
 * in a collective). */
#define DX_HDA_LOCAL_DOUBLES 8
int dx_engine_set_stream(dx_engine* e, void* cuda_stream);   /* NULL: back to the engine's own stream */
int dx_hda_send(dx_engine* e, void* send_dev, int64_t seg_records);
/* Fused pack + exchange over peer memory: peer_recv[r] is rank r's receive buffer mapped into this process (symmetric /
 * IPC memory); the pack kernel stores the segment for rank r into slot <this rank> of that buffer over NVLink, so no
 * collective moves the records.  The caller puts a cross-rank barrier between this call and dx_hda_receive. */
int dx_hda_send_peers(dx_engine* e, const uint64_t* peer_recv, int64_t seg_records);
int dx_hda_receive(dx_engine* e, const void* recv_dev, int64_t seg_records, double* local_dev);
int dx_hda_reduce(dx_engine* e, const double* all_dev, double* summary_dev);

#ifdef __cplusplus
}
#endif
#endif /* DXB200_H */
