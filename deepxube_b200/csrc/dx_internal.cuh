// Internal definitions shared by the engine's translation units (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <string.h>

#include "dxb200.h"

#define DX_NIL 0xFFFFFFFFu
#define DX_PENDING 0x80000000u          // slot rep points into the child batch, not the node store
#define DX_EMPTY64 0xFFFFFFFFFFFFFFFFull
#define DX_SORT_TILE 2048               // elements per radix / merge tile
#define DX_NUM_DIGITS 12                // 8 bytes of f-key + 4 bytes of instance id
#define DX_MAX_OUT 256                  // max network outputs (actions) supported (lightsout.15: 225)

// error bits in Ctl::error (sticky, set on the device)
#define DX_DEVERR_NODES 1u
#define DX_DEVERR_OPEN 2u
#define DX_DEVERR_POP 4u
#define DX_DEVERR_HDA 8u      // HDA*: an exchange segment / the receive capacity overflowed
#define DX_DEVERR_BACKUPS 16u // the Bellman-backup log is full (dx_config.max_backups); drain it more often
#define DX_DEVERR_PEER 32u    // HDA*: another rank of the distributed search reported a device error
#define DX_DEVERR_NUMERIC 64u // split-fp16 network: an activation left the fp16 range (|x| > 65504)
// Not an error: every instance is finished().  Set by k_end_iter only inside dx_run's bursts (several iterations launched per
// host poll): every kernel returns on a non-zero error word, so the rest of a burst is a no-op and the search stops in the exact
// iteration the reference's loop would (deepxube/_solve.py:150-161).  dx_run clears it before it returns.
#define DX_DEVERR_DONE 0x80000000u
#define DX_END_HALT 8          // k_end_iter gen_mode flag: raise DX_DEVERR_DONE when no instance is left
#define DX_HDA_LOCALS 8       // doubles per rank in the all-gathered per-iteration scalars (k_hda_locals)

void dx_set_error(const std::string& msg);

// Programmatic dependent launch for the small-batch regime, where an iteration is a chain of tiny dependent kernels and the
// 2-3 us of launch / drain latency between two of them is a quarter of the iteration: a kernel launched through dx_launch()
// while g_dx_pdl is set may be scheduled as soon as its predecessor has STARTED (every kernel calls
// griddepcontrol.launch_dependents first) and parks in griddepcontrol.wait until the predecessor has completed and its writes
// are visible -- so nothing a predecessor produces is read before the wait, and nothing is written before it either.
// Dependencies stay transitive: a kernel's wait returns only after its predecessor's own wait has.  Without the launch
// attribute both instructions are no-ops, so the same kernels serve the plainly launched big-batch path.
extern thread_local bool g_dx_pdl;
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline void dx_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    if (!g_dx_pdl) {
        kernel<<<grid, block, smem, s>>>(args...);
        return;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}
__device__ __forceinline__ void dx_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void dx_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void dx_pdl_prologue() { dx_pdl_trigger(); dx_pdl_wait(); }
#endif

// optional per-kernel timing (CUDA events on the engine stream); no-ops unless profiling is on
#define DX_ST_GEMM_RES 7     // hidden-layer tensor-core / fp32 GEMM launches
#define DX_ST_GEMM_IN 8      // input-layer launches (one-hot + first GEMM, or the gather kernel)
void dx_prof_begin(int stage);
void dx_prof_end();

#define DX_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            dx_set_error(std::string(#call) + ": " + cudaGetErrorString(_e) + " at " + __FILE__ + ":" + \
                         std::to_string(__LINE__));                                            \
            return DX_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Domain description (device copy passed by value to kernels).
struct DomainDev {
    int kind;        // enum dx_domain
    int dim;
    int S;           // state bytes
    int SP;          // row bytes: S rounded up so that a 4-byte instance id fits at SP-4, multiple of 16
    int A;           // actions
    int in_count;    // NN integer inputs (S, or S + goal bytes for grid)
    int depth;       // one-hot depth
    const uint8_t* table;   // cube3: perm [12*54] (next[d] = cur[perm[a*54+d]]); npuzzle: lut [d*d*4]
};

// Device-resident control block: everything a kernel needs to size its own work, so that the host
// never has to synchronise inside an iteration.
struct Ctl {
    uint32_t n_inst;
    uint32_t n_pop;        // entries in the current popped list (all instances)
    uint32_t n_children;   // n_pop * A (A*) ; Q*: number of popped edges
    uint32_t n_kept;       // nodes appended to the node store this iteration
    uint32_t node_count;   // nodes in the node store
    uint32_t node_base;    // node_count at the start of this iteration
    uint32_t epoch;        // iteration stamp for the per-slot batch lists
    uint32_t error;
    uint32_t n_unfinished;
    uint32_t itr;          // PathFind.itr
    uint32_t open_total;   // entries in OPEN (all instances) after this iteration
    uint32_t n_tiles;      // merge tiles this iteration
    uint32_t nn_row_start; // heuristic job: rows [nn_row_start, nn_row_start + nn_n) of the job's row array
    uint32_t nn_n;
    uint32_t sort_final;   // which sort buffer holds the sorted batch
    uint32_t sort_n;       // elements in the sort batch
    unsigned long long heur_evals;
    unsigned long long gen_total;
    uint32_t pass_src[DX_NUM_DIGITS];
    uint32_t pass_skip[DX_NUM_DIGITS];
    uint32_t n_backups;    // records in the Bellman-backup log (evaluate-all engines)
    uint32_t backup_overflow;
    uint32_t backup_at;    // log position of this iteration's first record
    uint32_t young_total;  // entries in the young runs' current buffer (young-run OPEN, YoungBufs)
    uint32_t sort_slow;    // the sort's fix-up met a group with more than FIX_RUNS runs: the key cut does not fit these keys (host: cut 0)
};

// Per-instance state (structure of arrays, capacity max_instances (+1 for offset arrays)).
struct InstArrays {
    int32_t* batch;             // InstanceGraph.batch_size
    double* weight;             // InstanceGraph.weight
    double* lb;                 // InstanceGraph.lb
    unsigned long long* ub_key; // (float bits of g << 32) | pop sequence number of the best solved popped node
    uint32_t* goal_node;        // node id or DX_NIL
    uint32_t* itr;              // Instance.itr
    unsigned long long* gen;    // Instance.num_nodes_generated
    uint8_t* active;            // not finished() at the start of the current iteration
    uint32_t* pop_seq_base;     // nodes popped so far (tie-break for "first solved node wins")
    uint32_t* n_open;           // OPEN entries
    uint32_t* m_new;            // entries pushed this iteration
    uint32_t* k_pop;            // entries popped this iteration
    uint32_t* batch_off;        // [I+1] offsets into the sorted batch
    uint32_t* tile_off;         // [I+1] merge tile offsets
    unsigned long long* f_first;// key of the first popped entry this iteration
    uint8_t* goals;             // [I, SP] goal rows
    uint32_t* root;             // [I] node id of the instance's root
};

// is_solved (cube3.py:514-517, npuzzle.py:154-158: goal byte d*d = wildcard, grid.py:68-69, lightsout: all tiles off)
__device__ __forceinline__ bool row_solved(int kind, int dim, int S, const uint8_t* __restrict__ row,
                                           const uint8_t* __restrict__ goal) {
    bool ok = true;
    const int wild = dim * dim;
    for (int i = 0; i < S; ++i) {
        uint8_t g = goal[i];
        bool e = (row[i] == g);
        if (kind == DX_DOMAIN_NPUZZLE) e = e || (g == wild);
        ok = ok && e;
    }
    return ok;
}

// Training-label log (engines created with dx_config.max_backups > 0): one record per expanded node (A*) / popped edge (Q*)
struct BackupLog {
    uint8_t* state;      // [cap, SP] the record's state row
    uint8_t* pack;       // [cap, S] drain scratch: the rows without padding
    double* val;         // [cap] label (NaN: popped by an instance that had finished -- dropped at the drain)
    uint32_t* inst;      // [cap]
    uint8_t* act;        // [cap] Q*: the edge's action
    // what the drain-time passes (ub_heur_solns parent paths, LHBL tree backup) need
    uint32_t* node;      // [cap] A*: node id of the record; Q*: the node generated for the edge
    uint32_t* itr;       // [cap] PathFind.itr when it was expanded / popped
    uint8_t* solved;     // [cap] (A* only from here on)
    float* child_h;      // [cap, A] heuristic of every child
    uint32_t* child_id;  // [cap, A] node id of a kept child, DX_NIL for a child the CLOSED filter dropped
    uint32_t* node_rec;  // [max_nodes] record index + 1 of an expanded node, 0 = none
    double* node_aux;    // [max_nodes] Q* drain-time passes: per-node bound (ub_heur_solns) / min over popped edges (lhbl)
    uint32_t cap;
};

struct NetDev {
    int in_count, depth, H, Hp, num_blocks, out_dim, in_dim, in_dim_p;
    int act_depth;   // > 0: action-as-input Q network -- rows are (node, action) pairs, feature in_dim - act_depth + a is the action
    // fp32 parameters (BN folded)
    float* w0t;      // [in_dim, Hp]  first layer, transposed (row = one-hot feature)
    float* b0;       // [Hp]
    float* wres;     // [num_blocks*2, Hp, Hp]  row-major [out, in]
    float* bres;     // [num_blocks*2, Hp]
    float* wout;     // [out_dim, Hp]
    float* bout;     // [out_dim]
    // bf16 copies for the tensor-core path
    void* w0_bf16;   // [Hp, in_dim_p]  K-major
    void* wres_bf16; // [num_blocks*2, Hp, Hp] K-major
    // Bias folding of the tensor-core path: when H < Hp, activation columns H..H+2 hold 1.0 and weight columns
    // H..H+2 hold a 3-piece bf16 split of the fp32 bias (hi + mid + lo == bias exactly), so the GEMM adds the bias.
    int tc_fold;     // bit 0: hidden layers folded, bit 1: input layer folded
    float* b0_tc;    // [Hp] (fused first block: [2 * Hp]) first-layer bias for the unfolded input layer (carries the 1.0 of the ones
                     // columns), else nullptr
    // Fused first block (tc_l1 != 0): the network has no activation between the input Linear and the first hidden Linear
    // (resnet_fc.py:41-44, pytorch_models.py:149-160), so  relu(W1 (W0 x + b0) + b1) = relu((W1 W0) x + (W1 b0 + b1)):
    // w0_bf16 holds [W0 ; W1 W0] stacked ([2 * Hp, in_dim_p]) and ONE launch of the generated-input GEMM writes both the
    // skip stream x0 and the first hidden activation -- the first K = Hp hidden GEMM disappears.
    int tc_l1;
    void* wout_bf16; // [128, Hp] zero-padded bf16 output-layer weights (multi-output networks with out_dim <= 32), else nullptr
    float* bout_pad; // [128] output bias, zero-padded
    void* tc_plan;   // TMA tensor maps etc. of the tensor-core path (k_mlp_tc.cu), nullptr unless NET_BF16 / split NET_FP32
    void* tc_split;  // host TcSplitW*: fp16 (hi, lo) copies of the parameters -- NET_FP32 on the tensor cores (k_mlp_tc.cu), else nullptr
    void* small;     // host SmallNetHost*: packed weight stream of the one-launch small-evaluation kernel (k_mlp_small.cu), else nullptr
};
#define DX_SMALL_NET_ROWS 8192   // engines whose evaluations never exceed this many rows run the network with k_mlp_small

// Split-fp16 parameters of the fp32-accurate tensor-core path: every weight w is stored as the fp16 pair (hi, lo) of w * s
// (s = a power of two per layer that puts the largest weight in [2^13, 2^14), so that the lo halves are normal fp16 numbers):
// rows [hi(0..K) | lo(0..K)].
struct TcSplitW {
    void* w0;        // [Hp, 2 * in_dim_p]
    void* wres;      // [2 * num_blocks * Hp, 2 * Hp]
    void* wout;      // [128, 2 * Hp] zero-padded multi-output layer (1 < out_dim <= 32), else nullptr
    float s_in, s_out;
    float s_res[128];
};

static inline int dx_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// OPEN keys: order-preserving map of a float64 f onto unsigned 64-bit integers (negative values: all bits flipped,
// others: sign bit flipped), so that unsigned key order == numeric order for ANY f a host heuristic may return --
// the reference's heap (graph_search.py:48-51) orders negative costs correctly too.  NaN sorts after +inf.
__host__ __device__ __forceinline__ unsigned long long dx_f_to_key(double f) {
#ifdef __CUDA_ARCH__
    const unsigned long long b = (unsigned long long)__double_as_longlong(f);
#else
    unsigned long long b; memcpy(&b, &f, 8);
#endif
    return b ^ ((b >> 63) ? 0xFFFFFFFFFFFFFFFFull : 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dx_key_to_f(unsigned long long k) {
    const unsigned long long b = k ^ ((k >> 63) ? 0x8000000000000000ull : 0xFFFFFFFFFFFFFFFFull);
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double f; memcpy(&f, &b, 8); return f;
#endif
}

// ---------------------------------------------------------------------------------------------
// 64-bit hash of a row (state bytes + padding + instance id), 16 bytes at a time.
__host__ __device__ __forceinline__ uint64_t dx_mix64(uint64_t h) {
    h ^= h >> 33;
    h *= 0xff51afd7ed558ccdull;
    h ^= h >> 33;
    h *= 0xc4ceb9fe1a85ec53ull;
    h ^= h >> 33;
    return h;
}

__host__ __device__ __forceinline__ uint64_t dx_hash_row(const uint4* __restrict__ row, int chunks) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < chunks; ++i) {
        uint4 v = row[i];
        uint64_t a = ((uint64_t)v.y << 32) | v.x, b = ((uint64_t)v.w << 32) | v.z;
        h = dx_mix64(h ^ a) + 0x632BE59BD9B4E019ull;
        h = dx_mix64(h ^ b) + 0x9E3779B97F4A7C15ull;
    }
    return h;
}

__device__ __forceinline__ bool dx_rows_equal(const uint4* __restrict__ a, const uint4* __restrict__ b, int chunks) {
    bool eq = true;
    for (int i = 0; i < chunks; ++i) {
        uint4 x = a[i], y = b[i];
        eq = eq && (x.x == y.x) && (x.y == y.y) && (x.z == y.z) && (x.w == y.w);
    }
    return eq;
}

// ---------------------------------------------------------------------------------------------
// Launchers (each returns the number of kernels it launched).
struct ScanTemp {
    uint32_t* block_sums;   // [>= ceil(n_max / 4096) + 1]
};

// One CLOSED slot = one 32-byte DRAM sector: the key CAS, the per-iteration list head and the stored path cost of a state are
// touched by the same item within a few hundred cycles (link -> decide -> commit), so they share a sector instead of living in
// three arrays (three random sectors per item: round 1 measured 297 MB of DRAM traffic for 144 MB of algorithmic bytes).
struct __align__(32) ClosedSlot {
    unsigned long long key;    // (tag << 32) | rep ; DX_EMPTY64 when free
    unsigned long long head;   // (epoch << 32) | first item of this state in the current batch
    float g;                   // CLOSED[state] (inf when new)
    uint32_t pad_[3];
};
struct ClosedDev {
    ClosedSlot* slots;
    uint32_t mask;                  // capacity - 1
};

#define DX_KEY_CUT 23          // low f-key bits left to the sort's fix-up (k_open.cu, digit_of)
struct SortBufs {
    unsigned long long* key[2];
    uint32_t* inst[2];
    uint32_t* val[2];
    uint32_t* hist;       // [256 * max_tiles (+ slack)]
    uint32_t* ghist;      // [DX_NUM_DIGITS * 256] global digit histograms of the batch
    uint32_t* gbase;      // [DX_NUM_DIGITS * 256] exclusive bin offsets per digit position
    // one-sweep passes (decoupled look-back): per (tile, digit value) a 64-bit word (epoch << 32 | flag << 30 | count),
    // valid only while its epoch equals the pass's -- never needs clearing; tickets[d] hands out tile ids in launch order
    unsigned long long* status;   // [max_tiles * 256]
    uint32_t* tickets;            // [DX_NUM_DIGITS] + [DX_NUM_DIGITS]: the sort counter behind the epochs (survives engine resets:
                                  // the status words of earlier searches must never match again)
    uint32_t* part_a;             // [merge tiles + 1] merge-path partition of every merge tile (k_merge_partition)
    uint32_t* part_i;             // [merge tiles + 1] ... and its instance
    // the digit passes leave out the low key_cut bits of the f-key; groups of equal (instance, key >> key_cut) are put in full-key
    // order afterwards (k_sort_fixup lists the groups that are not, k_sort_fixup_groups repairs them)
    uint32_t* fix_list;           // [max_sort / 2 + 2] listed inversions; fix_list[0] = how many
    uint32_t* fix_claim;          // [max_sort] group claims, tagged with the sort counter
    int key_cut;                  // DX_KEY_CUT, or 0 (DXB200_SORT_CUT=0): every key bit goes through the digit passes
    int onesweep;                 // 1: one kernel per digit pass; 0 (DXB200_SORT_ONESWEEP=0): histogram / tile scan / scatter
    int max_tiles;
    int num_digits;       // digit positions that can vary: 8 key bytes + the bytes max_instances - 1 occupies
};

// Small-batch A* engines (the solve-time regime: <= DX_SMALL_SORT pushed entries per iteration) keep OPEN as TWO sorted runs per
// instance: the main run (open_key / open_val, consumed from the front through m_head) and a small YOUNG run of what was pushed
// since the last compaction.  One block then does a whole iteration's OPEN work in shared memory -- keys, sort of the batch, merge
// with the young run, pop of the batch_size smallest of (main run head, young run), lb / finished() bookkeeping -- instead of
// rewriting the whole 100 k-entry run through three grid-wide kernels (k_merge_partition, k_merge, k_end_iter) for 100 pops.
// Every DX_YOUNG_CAP / sort_bound iterations the host schedules a compaction (main <- merge(main, young); the existing merge
// kernels).  Order is unchanged: everything in the main run was pushed before everything in the young run, which was pushed
// before the batch, and every merge takes the older run first on ties -- the reference heap's (f, push count) order.
#define YT_MAX_INST_HOST 256                   // instance slots an engine with a two-run OPEN may have (= YT_MAX_INST, k_open.cu)
#define DX_YOUNG_CAP 6144                      // entries the young runs (all instances) may hold after an iteration
#define DX_YOUNG_BUF (DX_YOUNG_CAP + 2048)     // ... plus one pushed batch (DX_SMALL_SORT)
struct YoungBufs {
    unsigned long long* key[2];    // [DX_YOUNG_BUF] ping-pong by iteration parity
    uint32_t* val[2];
    uint32_t* off[2];              // [I + 1] per-instance offsets (packed)
    uint32_t* head[2];             // [I] entries of the run already popped
    uint32_t* m_head;              // [I] entries of the MAIN run already popped
    // big batches (dx_launch_young_iter: the same two runs handled by grid-wide merges, engines with <= 256 instance slots)
    uint32_t* ka;                  // [I] pops taken from the main run this iteration
    uint32_t* kp;                  // [I] ... and from the merged young run
    uint32_t* mh_old;              // [I] m_head before this iteration's pops
    uint32_t* tile_off1;           // [I + 1] merge tiles of (young remainder + batch)
    uint32_t* tile_off2;           // [I + 1] merge tiles of the popped lists
    uint32_t* zeros;               // [I] k = 0: the first merge pops nothing
    uint32_t* n_tiles;             // [2] tile counts of the two merges
    uint32_t cap;                  // entries a young buffer holds
};

struct MlpBufs {
    float* x[3];          // fp32 activations [cap, Hp]
    void* xb[3];          // 16-bit activations: bf16 [cap, Hp] slices of ONE [cap, 3 * Hp] array (xb[i] = base + i * Hp, so that a
                          // launch can write two buffers at once), or -- split-fp16 mode -- three [cap, 2 * Hp] arrays ([hi | lo])
    long long pitch;      // elements between consecutive rows of an xb buffer
    float* partial;       // [16, cap] fp32 partial output-layer dot products of the fused last GEMM (k_out_final_bf16)
    long long cap;        // rows
};

// k_domain.cu
int dx_launch_expand(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const InstArrays& ia, const uint8_t* node_state,
                     const float* node_g, const uint32_t* popped, const uint32_t* popped_inst, const uint32_t* pop_off,
                     uint8_t* child_state, float* child_g, uint8_t* pop_solved, int max_pop);
int dx_launch_goal_resolve(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const float* node_g, const uint32_t* popped,
                           const uint32_t* popped_inst, const uint32_t* pop_off, const uint8_t* pop_solved, int max_pop);
int dx_launch_check_solved(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const InstArrays& ia, const uint8_t* node_state,
                           const float* node_g, const uint32_t* popped, const uint32_t* popped_inst, const uint32_t* pop_off,
                           uint8_t* pop_solved, int max_pop);
int dx_launch_expand_edges(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const uint32_t* edges, uint32_t* popped_out,
                           uint8_t* node_state, float* node_g, uint32_t* node_parent, uint8_t* node_action, int max_pop);
int dx_launch_standalone_expand(cudaStream_t s, const DomainDev& d, const uint8_t* rows, long long n, const int32_t* actions,
                                uint8_t* out_rows);
int dx_launch_standalone_solved(cudaStream_t s, const DomainDev& d, const uint8_t* rows, const uint8_t* goal_rows, long long n,
                                uint8_t* out);

// k_closed.cu
int dx_launch_closed_init(cudaStream_t s, const ClosedDev& c);
int dx_launch_closed_seed(cudaStream_t s, const DomainDev& d, const ClosedDev& c, const uint8_t* node_state, uint32_t first_node,
                          uint32_t n);
int dx_launch_closed_filter(cudaStream_t s, const DomainDev& d, const ClosedDev& c, const Ctl* ctl, const InstArrays& ia,
                            const uint8_t* node_state, const float* node_g, const uint8_t* child_state, const float* child_g,
                            const uint32_t* item_node, const uint32_t* popped_inst, uint32_t* child_slot, uint32_t* child_next,
                            uint8_t* child_flags, int max_children, int group);
int dx_launch_scatter_kept(cudaStream_t s, const DomainDev& d, const ClosedDev& c, Ctl* ctl, const uint8_t* child_state,
                           const float* child_g, const uint32_t* child_slot, const uint8_t* child_flags,
                           const uint32_t* child_prefix, const uint32_t* popped, const uint32_t* popped_inst,
                           uint8_t* node_state, float* node_g, uint32_t* node_parent, uint8_t* node_action, uint32_t* new_inst,
                           int max_children, const float* child_h = nullptr, float* node_h = nullptr);
int dx_launch_small_push_v(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* child_prefix, const uint32_t* pop_off_cur,
                           uint32_t* pop_off_next, const uint32_t* open_off_cur, uint32_t* open_off_next, int A, uint32_t max_pop,
                           uint32_t max_open, const float* node_g, const float* node_h, const uint32_t* new_inst, const SortBufs& sb);
int dx_launch_small_filter_q(cudaStream_t s, const DomainDev& d, const ClosedDev& c, Ctl* ctl, const InstArrays& ia,
                             const uint8_t* node_state, const float* node_g, const uint32_t* popped, const uint32_t* popped_inst,
                             const uint32_t* pop_off, uint8_t* pop_solved, uint32_t* child_slot, uint32_t* child_next,
                             uint8_t* child_flags, uint32_t* child_prefix);
int dx_launch_small_push_q(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* child_prefix, const uint32_t* pop_off_cur,
                           uint32_t* pop_off_next, const uint32_t* open_off_cur, uint32_t* open_off_next, int A, uint32_t max_pop,
                           uint32_t max_open, const float* node_g, const float* node_q, const uint32_t* popped,
                           const uint32_t* popped_inst, const uint8_t* flags, const SortBufs& sb);
int dx_launch_small_filter_v(cudaStream_t s, const DomainDev& d, const ClosedDev& c, Ctl* ctl, const InstArrays& ia,
                             const uint8_t* child_state, const float* child_g, uint32_t* child_slot, uint32_t* child_next,
                             uint8_t* child_flags, uint32_t* child_prefix, const uint32_t* popped, const uint32_t* popped_inst,
                             uint8_t* node_state, float* node_g, uint32_t* node_parent, uint8_t* node_action, uint32_t* new_inst,
                             uint32_t max_nodes, const uint32_t* pop_off, const uint8_t* pop_solved);
int dx_launch_bellman_backup(cudaStream_t s, const DomainDev& d, Ctl* ctl, const InstArrays& ia, const uint8_t* node_state,
                             const uint32_t* popped, const uint32_t* popped_inst, const uint8_t* pop_solved, const float* child_h,
                             const BackupLog& bk, int max_pop);
int dx_launch_backup_children(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const uint8_t* child_flags,
                              const uint32_t* child_prefix, const BackupLog& bk, int max_children);
int dx_launch_backup_ub_paths(cudaStream_t s, const Ctl* ctl, const uint32_t* node_parent, const BackupLog& bk);
int dx_launch_her_deepest(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const BackupLog& bk, const uint32_t* node_parent,
                          const uint8_t* node_state, int is_q, const uint32_t* root, uint32_t* rec_cand,
                          uint32_t* rec_depth, uint32_t* best_depth, uint32_t* deep_node, uint8_t* out_rows, int max_inst);
int dx_launch_backup_tree(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const BackupLog& bk, uint32_t itr);
int dx_launch_backup_q_modes(cudaStream_t s, const Ctl* ctl, const uint32_t* node_parent, const uint8_t* node_solved,
                             const float* node_h, const BackupLog& bk, int mode, uint32_t it_first, uint32_t it_last);
int dx_launch_backup_pack_rows(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const BackupLog& bk);
int dx_launch_backup_clear_map(cudaStream_t s, const Ctl* ctl, const BackupLog& bk);
int dx_launch_mark_solved(cudaStream_t s, const Ctl* ctl, const uint32_t* popped, const uint32_t* pop_off,
                          const uint8_t* pop_solved, uint8_t* node_solved, int max_pop);
int dx_launch_bellman_backup_q(cudaStream_t s, const DomainDev& d, Ctl* ctl, const uint8_t* node_state, const uint32_t* node_parent,
                               const uint8_t* node_action, const float* node_h, const uint8_t* node_solved,
                               const uint32_t* edge_inst, const BackupLog& bk, int max_pop);
int dx_launch_closed_commit(cudaStream_t s, const ClosedDev& c, const Ctl* ctl, const float* node_g, const uint32_t* item_node,
                            const uint32_t* child_slot, const uint8_t* child_flags, int max_items);

// k_scan.cu
int dx_launch_scan_flags(cudaStream_t s, const uint8_t* flags, uint32_t* out_prefix, const uint32_t* n_ptr, uint32_t* total_out,
                         int n_max, const ScanTemp& t);
int dx_launch_scan_u32(cudaStream_t s, uint32_t* data, int n, const uint32_t* skip_flag, const ScanTemp& t);

// k_open.cu
int dx_launch_plan(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* child_prefix, const uint32_t* pop_off_cur,
                   uint32_t* pop_off_next, const uint32_t* open_off_cur, uint32_t* open_off_next, int child_mult, int push_mult,
                   uint32_t max_pop, uint32_t max_open, int kept_is_total = 0, int beam = 0);

// k_hda.cu (hash-distributed single search)
int dx_launch_hda_bucket(cudaStream_t s, const DomainDev& d, Ctl* ctl, const SortBufs& sb, const ScanTemp& t,
                         const uint8_t* child_state, const float* child_g, const uint32_t* popped, uint32_t rank, uint32_t world,
                         uint32_t* counts, uint8_t* send, int max_children);
int dx_launch_hda_unpack(cudaStream_t s, const DomainDev& d, Ctl* ctl, const uint8_t* recv, uint32_t n, uint8_t* child_state,
                         float* child_g, uint32_t* child_parent, uint8_t* child_action);
int dx_launch_scatter_kept_hda(cudaStream_t s, const DomainDev& d, const ClosedDev& c, Ctl* ctl, const uint8_t* child_state,
                               const float* child_g, const uint32_t* child_slot, const uint8_t* child_flags,
                               const uint32_t* child_prefix, const uint32_t* child_parent, const uint8_t* child_action,
                               uint8_t* node_state, float* node_g, uint32_t* node_parent, uint8_t* node_action,
                               uint32_t* new_inst, int max_children);
int dx_launch_hda_finish(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* pop_off_cur, const uint32_t* pop_off_next,
                         int A, unsigned long long k_total, unsigned long long f_first_min_bits, uint32_t ub_g_bits, int has_ub,
                         unsigned long long gen_add);
int dx_launch_hda_drop_root(cudaStream_t s, Ctl* ctl, uint32_t* pop_off);
struct HdaDst { uint8_t* seg[16]; };     // destination of the exchange segment for each rank (local send buffer or peer memory)
int dx_launch_hda_send(cudaStream_t s, const DomainDev& d, Ctl* ctl, const SortBufs& sb, const ScanTemp& t, const uint8_t* child_state,
                       const float* child_g, const uint32_t* popped, uint32_t rank, uint32_t world, uint32_t* counts, const HdaDst& dst,
                       uint32_t seg_records, uint32_t* aux, int max_children);
int dx_launch_hda_recv(cudaStream_t s, const DomainDev& d, Ctl* ctl, const uint8_t* recv, uint32_t world, uint32_t seg_records,
                       uint32_t max_children, uint32_t* src_off, uint8_t* child_state, float* child_g, uint32_t* child_parent,
                       uint8_t* child_action, uint32_t* aux);
int dx_launch_hda_locals(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const uint32_t* aux, double* loc);
int dx_launch_hda_reduce(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* pop_off_cur, const uint32_t* pop_off_next,
                         int A, int world, const double* all, double* summary);
uint32_t dx_hda_owner_host(const uint8_t* row, int SP, uint32_t world);
int dx_launch_make_keys(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const float* node_g, const float* node_h,
                        const uint32_t* new_inst, const SortBufs& sb, int max_n, int beam = 0);
int dx_launch_make_keys_q(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const float* node_g, const float* node_q,
                          const uint32_t* popped, const uint32_t* popped_inst, const uint8_t* flags, const uint32_t* prefix,
                          const SortBufs& sb, int A, int max_pop, int beam = 0);
#define DX_SMALL_SORT 2048     // batches up to this many entries are sorted by one block (k_sort_small)
int dx_launch_sort(cudaStream_t s, Ctl* ctl, const SortBufs& sb, const ScanTemp& t, int max_n, bool small = false);
int dx_launch_merge(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const SortBufs& sb,
                    const unsigned long long* open_key_cur, const uint32_t* open_val_cur, const uint32_t* open_off_cur,
                    unsigned long long* open_key_next, uint32_t* open_val_next, const uint32_t* open_off_next,
                    uint32_t* popped_next, uint32_t* popped_inst_next, const uint32_t* pop_off_next, int max_tiles, int keep_open = 1);
int dx_launch_small_tail_v(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* child_prefix, const uint32_t* pop_off_cur,
                           uint32_t* pop_off_next, const unsigned long long* m_key, const uint32_t* m_val, const uint32_t* m_off,
                           const YoungBufs& yb, int p, int A, uint32_t max_pop, const float* node_g, const float* node_h,
                           const uint32_t* new_inst, uint32_t* popped_next, uint32_t* popped_inst_next, int gen_mode);
// One iteration's OPEN side with two runs, grid-wide (after the batch is sorted): young <- merge(young remainder, batch); split of
// the k pops between main-run head and young run; popped list <- merge of the two prefixes.  push_mult / child_mult as dx_launch_plan.
int dx_launch_young_plan_b(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* child_prefix, const uint32_t* pop_off_cur,
                           uint32_t* pop_off_next, const uint32_t* m_off, const YoungBufs& yb, int p, int child_mult, int push_mult,
                           uint32_t max_pop);
int dx_launch_young_iter(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const SortBufs& sb, const unsigned long long* m_key,
                         const uint32_t* m_val, const uint32_t* m_off, const YoungBufs& yb, int p, uint32_t* popped_next,
                         uint32_t* popped_inst_next, const uint32_t* pop_off_next, int max_pop);
int dx_launch_young_compact(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const SortBufs& sb, const YoungBufs& yb, int p,
                            const unsigned long long* open_key_cur, const uint32_t* open_val_cur, const uint32_t* open_off_cur,
                            unsigned long long* open_key_next, uint32_t* open_val_next, uint32_t* open_off_next,
                            uint32_t* popped_scratch, uint32_t* popped_inst_scratch, const uint32_t* pop_off_any, int max_tiles,
                            uint32_t max_open);
// gen_mode: 0 graph_v (A children per popped node), 1 graph_q (one node per popped edge), 2 beam_v / 3 beam_q (as 0 / 1,
// finished on a goal)
int dx_launch_end_iter(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* pop_off_cur, const uint32_t* pop_off_next,
                       int A, int gen_mode);
int dx_launch_beam_keep(cudaStream_t s, const Ctl* ctl, const float* child_g, uint8_t* child_flags, int max_children);
int dx_launch_beam_keep_q(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const uint32_t* popped_inst, uint8_t* flags, int max_items);

// k_mlp_f32.cu / k_mlp_tc.cu
// rows: [*, SP] row array; the job is ctl->nn_row_start / nn_n.  out_h (nullable) <- max(out0,0) (V) or min_a q (Q);
// out_q (nullable) [*, out_dim] <- max(out,0).  Outputs are indexed by (nn_row_start + r) when absolute != 0, else r.
int dx_launch_out_layer(cudaStream_t s, const NetDev& net, const Ctl* ctl, const float* X, float* out_h, float* out_q,
                        int absolute, long long max_rows);
int dx_launch_mlp_f32(cudaStream_t s, const DomainDev& d, const NetDev& net, const MlpBufs& mb, const Ctl* ctl,
                      const uint8_t* rows, const uint8_t* goals, float* out_h, float* out_q, int absolute, long long max_rows);
int dx_launch_mlp_bf16(cudaStream_t s, const DomainDev& d, const NetDev& net, const MlpBufs& mb, const Ctl* ctl,
                       const uint8_t* rows, const uint8_t* goals, float* out_h, float* out_q, int absolute, long long max_rows);
int dx_launch_act_begin(cudaStream_t s, const NetDev& net, Ctl* ctl);
int dx_launch_act_end(cudaStream_t s, const NetDev& net, Ctl* ctl, const float* q, float* out_h, int absolute, long long max_nodes);
bool dx_mlp_bf16_gen_input(const NetDev& net, const DomainDev& d);
// fp32-accurate network on the tensor cores: whether this shape is supported, and the (re)upload of the split parameters from the
// host copies (w0p [Hp, in_dim_p], wres [2 * nb, Hp, Hp], wout [out_dim, Hp], all fp32 with BN folded)
bool dx_mlp_split_supported(const NetDev& net);
int dx_mlp_split_upload(cudaStream_t s, NetDev& net, const float* w0p, const float* wres, const float* wout);
int dx_mlp_bf16_prepare(const NetDev& net, const MlpBufs& mb, void* onehot_buf, bool gen_input, void** plan_out);
// k_mlp_small.cu: the whole network in one launch for engines with small evaluations (res_dim <= 256)
bool dx_mlp_small_supported(const NetDev& net, const DomainDev& d);
int dx_mlp_small_upload(cudaStream_t s, NetDev& net, bool split, const float* w0p, const float* wres);
void dx_mlp_small_free(NetDev& net);
int dx_launch_mlp_small(cudaStream_t s, const DomainDev& d, const NetDev& net, const Ctl* ctl, const uint8_t* rows,
                        const uint8_t* goals, float* out_h, float* out_q, int absolute, long long max_rows);
