// Engine: owns device memory + one stream, sequences the kernels of one PathFind.step() and implements
// the C ABI of include/dxb200.h.  No kernel launch depends on a host-side read of device state: every
// kernel sizes itself from the device-resident Ctl block, so an iteration is a fixed launch sequence.
//
// Iteration order (A*, deepxube/base/pathfinding.py:338-400 + :621-643):
//   expand + is_solved + record_goal -> CLOSED filter -> compaction (node ids = push order) ->
//   heuristic of the kept children -> f-keys -> sort -> merge/pop -> lb/finished/itr.
//   (The reference evaluates the heuristic before filtering; values of filtered children are never
//    read, so evaluating only the kept ones is observationally identical -- SURVEY.md App. B.6.)
// Iteration order (Q*, base/pathfinding.py:471-525 + :646-671):
//   is_solved + record_goal -> CLOSED filter on popped nodes -> edges of kept nodes -> f-keys -> sort ->
//   merge/pop edges -> next_state per popped edge -> heuristic (all-action q, h = min) -> lb/finished/itr.
#include <math.h>
#include <cuda_fp16.h>
#include <cmath>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "dx_internal.cuh"

static thread_local std::string g_last_error;
void dx_set_error(const std::string& msg) { g_last_error = msg; }
extern "C" const char* dx_last_error(void) { return g_last_error.c_str(); }
extern "C" int dx_version(void) { return 100; }

#define DX_FAIL(code, msg) do { dx_set_error(msg); return (code); } while (0)

// cube3 move table: next[d] = cur[perm[a][d]]; the 20 stickers each move displaces, as (dst, src) pairs.
// Same permutation as the reference's 24 (new, old) index pairs per move (deepxube/domains/cube3.py:598-671,
// action order :492); pinned against the reference by tests/golden/domains.npz::cube3_perm.
static const uint8_t kCube3Pairs[12][20][2] = {
    {{0,2},{1,5},{2,8},{3,1},{5,7},{6,0},{7,3},{8,6},{20,38},{23,41},{26,44},{29,47},{32,50},{35,53},{38,29},{41,32},{44,35},{47,20},{50,23},{53,26}},
    {{0,6},{1,3},{2,0},{3,7},{5,1},{6,8},{7,5},{8,2},{20,47},{23,50},{26,53},{29,38},{32,41},{35,44},{38,20},{41,23},{44,26},{47,29},{50,32},{53,35}},
    {{9,11},{10,14},{11,17},{12,10},{14,16},{15,9},{16,12},{17,15},{18,45},{21,48},{24,51},{27,36},{30,39},{33,42},{36,18},{39,21},{42,24},{45,27},{48,30},{51,33}},
    {{9,15},{10,12},{11,9},{12,16},{14,10},{15,17},{16,14},{17,11},{18,36},{21,39},{24,42},{27,45},{30,48},{33,51},{36,27},{39,30},{42,33},{45,18},{48,21},{51,24}},
    {{0,45},{1,46},{2,47},{9,44},{10,43},{11,42},{18,20},{19,23},{20,26},{21,19},{23,25},{24,18},{25,21},{26,24},{42,2},{43,1},{44,0},{45,9},{46,10},{47,11}},
    {{0,44},{1,43},{2,42},{9,45},{10,46},{11,47},{18,24},{19,21},{20,18},{21,25},{23,19},{24,26},{25,23},{26,20},{42,11},{43,10},{44,9},{45,0},{46,1},{47,2}},
    {{6,38},{7,37},{8,36},{15,51},{16,52},{17,53},{27,29},{28,32},{29,35},{30,28},{32,34},{33,27},{34,30},{35,33},{36,17},{37,16},{38,15},{51,6},{52,7},{53,8}},
    {{6,51},{7,52},{8,53},{15,38},{16,37},{17,36},{27,33},{28,30},{29,27},{30,34},{32,28},{33,35},{34,32},{35,29},{36,8},{37,7},{38,6},{51,15},{52,16},{53,17}},
    {{2,18},{5,19},{8,20},{9,33},{12,34},{15,35},{18,15},{19,12},{20,9},{33,8},{34,5},{35,2},{36,38},{37,41},{38,44},{39,37},{41,43},{42,36},{43,39},{44,42}},
    {{2,35},{5,34},{8,33},{9,20},{12,19},{15,18},{18,2},{19,5},{20,8},{33,9},{34,12},{35,15},{36,42},{37,39},{38,36},{39,43},{41,37},{42,44},{43,41},{44,38}},
    {{0,29},{3,28},{6,27},{11,26},{14,25},{17,24},{24,0},{25,3},{26,6},{27,11},{28,14},{29,17},{45,47},{46,50},{47,53},{48,46},{50,52},{51,45},{52,48},{53,51}},
    {{0,24},{3,25},{6,26},{11,27},{14,28},{17,29},{24,17},{25,14},{26,11},{27,6},{28,3},{29,0},{45,51},{46,48},{47,45},{48,52},{50,46},{51,53},{52,50},{53,47}}};

static int make_domain_host(int domain, int dim, DomainDev* d, std::vector<uint8_t>* table) {
    d->kind = domain;
    d->dim = dim;
    d->table = nullptr;
    table->clear();
    if (domain == DX_DOMAIN_CUBE3) {
        d->S = 54; d->A = 12; d->in_count = 54; d->depth = 6; d->dim = 0;
        table->resize(12 * 54);
        for (int a = 0; a < 12; ++a) {
            for (int i = 0; i < 54; ++i) (*table)[a * 54 + i] = (uint8_t)i;
            for (int p = 0; p < 20; ++p) (*table)[a * 54 + kCube3Pairs[a][p][0]] = kCube3Pairs[a][p][1];
        }
    } else if (domain == DX_DOMAIN_NPUZZLE) {
        if (dim < 2 || dim > 15) DX_FAIL(DX_ERR_ARG, "npuzzle dim must be in 2..15");
        d->S = dim * dim; d->A = 4; d->in_count = dim * dim; d->depth = dim * dim;
        table->resize(dim * dim * 4);
        for (int z = 0; z < dim * dim; ++z) {
            const int i = z / dim, j = z % dim;
            // blank moves: U -> row+1, D -> row-1, L -> col+1, R -> col-1; blocked move keeps the blank (npuzzle.py:264-308)
            (*table)[z * 4 + 0] = (uint8_t)(i < dim - 1 ? (i + 1) * dim + j : z);
            (*table)[z * 4 + 1] = (uint8_t)(i > 0 ? (i - 1) * dim + j : z);
            (*table)[z * 4 + 2] = (uint8_t)(j < dim - 1 ? i * dim + j + 1 : z);
            (*table)[z * 4 + 3] = (uint8_t)(j > 0 ? i * dim + j - 1 : z);
        }
    } else if (domain == DX_DOMAIN_GRID) {
        if (dim < 1 || dim > 255) DX_FAIL(DX_ERR_ARG, "grid dim must be in 1..255");
        d->S = 2; d->A = 4; d->in_count = 4; d->depth = dim;
    } else if (domain == DX_DOMAIN_LIGHTSOUT) {
        if (dim < 2 || dim > 15) DX_FAIL(DX_ERR_ARG, "lightsout dim must be in 2..15");
        // action a toggles [a, a+dim, a-dim, a+1, a-1], an off-board neighbour replaced by a (lightsout.py:69-78); one-hot
        // depth 1 = the 0 / 1 tile value is the network input itself (lightsout.py:107-108, pytorch_models.py:15-24)
        d->S = dim * dim; d->A = dim * dim; d->in_count = dim * dim; d->depth = 1;
        table->resize(dim * dim * 5);
        for (int a = 0; a < dim * dim; ++a) {
            const int x = a / dim, y = a % dim;
            (*table)[a * 5 + 0] = (uint8_t)a;
            (*table)[a * 5 + 1] = (uint8_t)(x < dim - 1 ? a + dim : a);
            (*table)[a * 5 + 2] = (uint8_t)(x > 0 ? a - dim : a);
            (*table)[a * 5 + 3] = (uint8_t)(y < dim - 1 ? a + 1 : a);
            (*table)[a * 5 + 4] = (uint8_t)(y > 0 ? a - 1 : a);
        }
    } else {
        DX_FAIL(DX_ERR_ARG, "unknown domain");
    }
    d->SP = ((d->S + 4 + 15) / 16) * 16;
    return DX_OK;
}

// ---------------------------------------------------------------------------------------------
__global__ void k_after_scan_v(Ctl* ctl, uint32_t max_nodes) {
    if (ctl->error) return;
    const uint32_t kept = ctl->n_kept;
    if ((unsigned long long)ctl->node_base + kept > max_nodes) { ctl->error |= DX_DEVERR_NODES; return; }
    ctl->node_count = ctl->node_base + kept;
    ctl->nn_row_start = ctl->node_base;
    ctl->nn_n = kept;
}

// Q*: after the merge the number of popped edges (= nodes to generate) is pop_off_next[I].
__global__ void k_after_merge_q(Ctl* ctl, const uint32_t* pop_off_next, uint32_t max_nodes) {
    dx_pdl_prologue();
    if (ctl->error) return;
    const uint32_t n = pop_off_next[ctl->n_inst];
    if ((unsigned long long)ctl->node_base + n > max_nodes) { ctl->error |= DX_DEVERR_NODES; return; }
    ctl->n_kept = n;
    ctl->node_count = ctl->node_base + n;
    ctl->nn_row_start = ctl->node_base;
    ctl->nn_n = n;
}

__global__ void k_fill_zero_vals(const Ctl* ctl, float* node_h, float* node_q, int A) {
    dx_pdl_prologue();
    if (ctl->error) return;
    const uint32_t n = ctl->nn_n, start = ctl->nn_row_start;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        node_h[start + r] = 0.f;
        if (node_q)
            for (int a = 0; a < A; ++a) node_q[(size_t)(start + r) * A + a] = 0.f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&ctl->heur_evals, (unsigned long long)n);
}

// HOST mode: values uploaded for rows [nn_row_start, +nn_n): V -> node_h ; Q -> node_q and node_h = min_a q
__global__ void k_set_host_vals(const Ctl* ctl, const float* vals, float* node_h, float* node_q, int A, int is_q) {
    if (ctl->error) return;
    const uint32_t n = ctl->nn_n, start = ctl->nn_row_start;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (!is_q) {
            node_h[start + r] = vals[r];
        } else {
            float m = __int_as_float(0x7f800000);
            for (int a = 0; a < A; ++a) {
                const float v = vals[(size_t)r * A + a];
                node_q[(size_t)(start + r) * A + a] = v;
                m = fminf(m, v);
            }
            node_h[start + r] = m;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&ctl->heur_evals, (unsigned long long)n);
}

__global__ void k_set_job(Ctl* ctl, uint32_t start, uint32_t n) { ctl->nn_row_start = start; ctl->nn_n = n; }
// PathFind.remove_instances (base/pathfinding.py:300-310): the instance stops searching -- no more pops, expansions or
// OPEN entries (k_plan drops the run of an inactive instance) -- and no longer counts as unfinished
// The popped list (nodes the next step expands) is compacted so that the removed instances' entries give their room back
// at once: one block per instance copies its segment to its new offset in the other buffer; the host copies it back.
__global__ void k_compact_popped(const uint32_t* __restrict__ off_old, const uint32_t* __restrict__ off_new, int n_inst,
                                 const uint32_t* __restrict__ popped, const uint32_t* __restrict__ popped_inst,
                                 uint32_t* __restrict__ popped_out, uint32_t* __restrict__ popped_inst_out) {
    for (int i = blockIdx.x; i < n_inst; i += gridDim.x) {
        const uint32_t src = off_old[i], dst = off_new[i], len = off_new[i + 1] - off_new[i];
        for (uint32_t t = threadIdx.x; t < len; t += blockDim.x) {
            popped_out[dst + t] = popped[src + t];
            popped_inst_out[dst + t] = popped_inst[src + t];
        }
    }
}
__global__ void k_after_remove(Ctl* ctl, uint32_t n_pop, uint32_t n_children, uint32_t n_unfinished) {
    ctl->n_pop = n_pop;
    ctl->n_children = n_children;
    ctl->n_unfinished = n_unfinished;
}
__global__ void k_reset_backups(Ctl* ctl) { ctl->n_backups = 0; }
__global__ void k_clear_done(Ctl* ctl) { ctl->error &= ~DX_DEVERR_DONE; }
__global__ void k_set_job_children(Ctl* ctl) { ctl->nn_row_start = 0; ctl->nn_n = ctl->error ? 0 : ctl->n_children; }

// device-side parent walk (base/pathfinding.py:93-120): actions written goal -> root, reversed on the host
__global__ void k_path_walk(const uint32_t* node_parent, const uint8_t* node_action, uint32_t goal, int32_t* actions,
                            uint32_t* ids, int cap, int* len_out) {
    int len = 0;
    uint32_t id = goal;
    while (node_parent[id] != DX_NIL) {
        if (len < cap) { actions[len] = node_action[id]; ids[len] = id; }
        ++len;
        id = node_parent[id];
    }
    if (len < cap + 1) ids[len <= cap ? len : cap] = id;
    *len_out = len;
}

__global__ void k_gather_rows(const uint8_t* node_state, const uint32_t* ids, int n, int SP, uint8_t* out) {
    const int chunks = SP / 16;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n * chunks; t += gridDim.x * blockDim.x) {
        const int r = t / chunks, c = t % chunks;
        reinterpret_cast<uint4*>(out + (size_t)r * SP)[c] = reinterpret_cast<const uint4*>(node_state + (size_t)ids[r] * SP)[c];
    }
}

__global__ void k_gather_f32(const float* src, const uint32_t* ids, int n, float* out) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) out[t] = src[ids[t]];
}

// ---------------------------------------------------------------------------------------------
enum { ST_EXPAND = 0, ST_CLOSED, ST_SCATTER, ST_HEUR, ST_SORT, ST_MERGE, ST_BOOK, ST_GEMM_RES = DX_ST_GEMM_RES, ST_GEMM_IN = DX_ST_GEMM_IN, ST_COUNT };

struct dx_engine {
    dx_config cfg;
    DomainDev dom;
    std::vector<uint8_t> table_h;
    cudaStream_t stream = nullptr;
    bool is_q = false;
    bool is_beam = false;              // beam_v: graph_v pipeline without CLOSED and without a persistent OPEN
    int A = 0, SP = 0, S = 0;
    uint32_t max_nodes = 0, max_open = 0, max_pop = 0, max_children = 0, max_sort = 0;
    int max_inst = 0;
    int n_inst = 0;
    int parity = 0;              // which popped / OPEN buffers are "current"
    bool pending = false;        // HOST mode: between step_begin and step_end
    bool weights_loaded = false;
    long long launches = 0;
    std::vector<uint8_t> removed;        // instance slots given up through dx_remove_instances (host copy)

    // device memory
    uint8_t* d_table = nullptr;
    Ctl* d_ctl = nullptr;
    Ctl* h_ctl = nullptr;        // pinned
    uint8_t* node_state = nullptr; float* node_g = nullptr; float* node_h = nullptr; uint32_t* node_parent = nullptr;
    uint8_t* node_action = nullptr; float* node_q = nullptr;
    ClosedDev closed{};
    uint8_t* child_state = nullptr; float* child_g = nullptr; uint32_t* child_slot = nullptr; uint32_t* child_next = nullptr;
    uint8_t* child_flags = nullptr; uint32_t* child_prefix = nullptr; uint32_t* new_inst = nullptr;
    // evaluate-all engines (dx_config.max_backups > 0): the heuristic of EVERY generated child, before the CLOSED filter,
    // and the Bellman-backup log of the expanded nodes
    bool fuse_small = true;
    bool eval_all = false;
    float* child_h = nullptr;
    BackupLog bk = {};
    uint32_t *her_cand = nullptr, *her_depth = nullptr, *her_best = nullptr, *her_node = nullptr;   // dx_backup_deepest scratch
    uint8_t* her_rows = nullptr;
    bool backups_q = false;          // graph_q with a backup log: labels of the popped edges
    uint8_t* node_solved = nullptr;
    uint32_t* popped[2] = {nullptr, nullptr}; uint32_t* popped_inst[2] = {nullptr, nullptr}; uint32_t* pop_off[2] = {nullptr, nullptr};
    uint32_t* edge_tmp = nullptr; uint8_t* pop_solved = nullptr;
    // HDA* mode (hash-distributed single search)
    int hda_rank = 0, hda_world = 0;
    uint32_t* hda_counts = nullptr; uint32_t* child_parent = nullptr; uint8_t* child_action = nullptr;
    uint32_t* hda_src_off = nullptr; uint32_t* hda_aux = nullptr;     // device-resident exchange: per-source offsets, children sent
    cudaStream_t own_stream = nullptr;                                 // the stream the engine created (e->stream may be a caller's)
    int64_t hda_children_sent = 0;
    unsigned long long* open_key[2] = {nullptr, nullptr}; uint32_t* open_val[2] = {nullptr, nullptr};
    uint32_t* open_off[2] = {nullptr, nullptr};
    SortBufs sort{};
    ScanTemp scan{};
    InstArrays ia{};
    NetDev net{};
    MlpBufs mlp{};
    uint8_t* d_stage = nullptr; size_t stage_bytes = 0;     // staging for host<->device transfers
    float* d_vals = nullptr;                                // HOST-mode value upload / heur_eval output
    std::vector<void*> allocs;

    // profiling
    int profiling = 0;                 // 0 off, 1 every stage + GEMM intervals, 2 only the hidden-layer GEMM interval
    int prof_graph_level = 0;          // level the profiling graphs were captured at
    std::vector<cudaEvent_t> ev;       // pairs
    std::vector<int> ev_stage;
    std::vector<int> ev_open;
    float stage_ms[ST_COUNT] = {0};
    long long stage_calls[ST_COUNT] = {0};
    cudaEvent_t run_ev[2] = {nullptr, nullptr};
    // one captured iteration per (identity of the OPEN buffer pair, buffer parity): slot = 2 * open_id + parity.  A compaction of the
    // two-run OPEN swaps the open_*[0] / [1] pointers instead of copying the merged runs back, so graphs exist for both identities
    int open_id = 0;
    cudaGraphExec_t graph_exec[4] = {nullptr, nullptr, nullptr, nullptr};
    // the same iteration captured with the stage timers as event-record nodes (used while profiling is on)
    cudaGraphExec_t graph_prof_exec[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> gev[4];   // begin/end pairs of the captured timers, per slot
    std::vector<int> gev_stage[4];
    int cap_prof = -1;                 // parity whose profiling graph is being captured, else -1
    int graph_launches = 0;
    long long sort_bound = 0;          // sum of batch_size * actions over the instances: no iteration pushes more entries
    bool graph_small = false;          // whether the captured graphs use the single-block sort
    bool use_graphs = true;
    int poll_every = 8;             // small-batch dx_run: iterations launched per host poll (DXB200_POLL_EVERY; 4 -> 8 -> 16: 67.2 / 65.7 / 65.2 us per KAT iteration)
    int poll_every_big = 4;         // ... for big batches (DXB200_POLL_EVERY_BIG)
    // young-run OPEN of small-batch A* engines (YoungBufs): the main run stays in open_*[0] while it is active
    YoungBufs young{};
    bool young_enabled = true;      // DXB200_YOUNG=0 disables
    bool young_active = false;      // decided when the first instances arrive after a reset; left (flushed) when the batch outgrows it
    bool young_big = false;         // ... handled by the grid-wide merges (dx_launch_young_iter) instead of the one-block tail
    int graph_young = 0;            // OPEN structure the captured graphs were built for: 0 one run, 1 young (one block), 2 young (grid-wide)
    long long y_bound = 0;          // upper bound of the entries in the young runs (host-side: compaction schedule)
    long long m_bound = 0;          // upper bound of the entries in the main runs (everything compacted so far)
    long long young_cap = 0;        // entries a young buffer holds (>= DX_YOUNG_BUF); big-batch two-run OPEN needs >= 4 * max_sort
    // DXB200_YOUNG_BIG: 0 never, 1 (default) graph_q only, 2 graph_v too.  Q* pushes A edges per kept node and pops one per batch slot, so
    // its runs grow by ~(A - 1) x batch entries per iteration and rewriting them dominates the OPEN side after a few iterations
    // (cube3, B = 10 000: 12 % of the step with 25 M entries).  A* pushes what survives the CLOSED filter: on cube3 / npuzzle the runs
    // stay within a few batches for tens of iterations and the five extra launches cost more than the shorter merge saves
    // (measured: npuzzle.5 B = 1000 x 64 -1.4 %, cube3 B = 10 000 x 8 +-0).
    int young_big_enabled = 1;
    bool cut_sticky = true;         // DXB200_SORT_CUT_STICKY=0: keep the key cut after a slow sort (tests of the fix-up's merge rounds)
    int graph_cut = -1;             // key cut the captured graphs were built with
    int young_every = 0;            // DXB200_YOUNG_EVERY=n: compact the big-batch young runs every n iterations (0: the cost model)
    bool pdl = true;                // small-batch graphs: programmatic dependent launches (DXB200_PDL=0 disables)
    bool halt_flag = false;         // the step being launched / captured belongs to dx_run: k_end_iter may raise DX_DEVERR_DONE
    float last_run_ms = 0.f;

    template <typename T> int alloc(T** p, size_t count) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
        if (e != cudaSuccess) {
            dx_set_error(std::string("cudaMalloc of ") + std::to_string(count * sizeof(T)) + " bytes failed: " + cudaGetErrorString(e));
            return DX_ERR_CUDA;
        }
        allocs.push_back(q);
        *p = (T*)q;
        return DX_OK;
    }
};

#define DX_TRY(x) do { int _r = (x); if (_r < 0) return _r; } while (0)

// Timed intervals may nest (kernel-level intervals inside the heuristic stage): begin pushes on a stack,
// end closes the innermost open interval.
static void stage_begin(dx_engine* e, int st) {
    if (!e->profiling) return;
    if (e->profiling == 2 && st != ST_GEMM_RES) { e->ev_open.push_back(-1); return; }   // not timed at this level
    cudaEvent_t a;
    cudaEventCreate(&a);
    if (e->cap_prof >= 0) {            // inside stream capture: a real event-record node, kept for every replay
        cudaEventRecordWithFlags(a, e->stream, cudaEventRecordExternal);
        auto& v = e->gev[e->cap_prof];
        e->ev_open.push_back((int)v.size());
        v.push_back(a);
        v.push_back(nullptr);
        e->gev_stage[e->cap_prof].push_back(st);
        return;
    }
    cudaEventRecord(a, e->stream);
    e->ev_open.push_back((int)e->ev.size());
    e->ev.push_back(a);
    e->ev.push_back(nullptr);
    e->ev_stage.push_back(st);
}
static void stage_end(dx_engine* e) {
    if (!e->profiling || e->ev_open.empty()) return;
    if (e->ev_open.back() < 0) { e->ev_open.pop_back(); return; }
    cudaEvent_t b;
    cudaEventCreate(&b);
    if (e->cap_prof >= 0) {
        cudaEventRecordWithFlags(b, e->stream, cudaEventRecordExternal);
        e->gev[e->cap_prof][e->ev_open.back() + 1] = b;
        e->ev_open.pop_back();
        return;
    }
    cudaEventRecord(b, e->stream);
    e->ev[e->ev_open.back() + 1] = b;
    e->ev_open.pop_back();
}
static void stage_collect(dx_engine* e) {
    if (e->ev.empty()) return;
    cudaStreamSynchronize(e->stream);
    for (size_t i = 0; i + 1 < e->ev.size(); i += 2) {
        if (e->ev[i + 1]) {
            float ms = 0;
            cudaEventElapsedTime(&ms, e->ev[i], e->ev[i + 1]);
            e->stage_ms[e->ev_stage[i / 2]] += ms;
            e->stage_calls[e->ev_stage[i / 2]] += 1;
            cudaEventDestroy(e->ev[i + 1]);
        }
        cudaEventDestroy(e->ev[i]);
    }
    e->ev.clear();
    e->ev_stage.clear();
    e->ev_open.clear();
}

// after a replay of the profiling graph of parity p has completed: add its captured intervals
static void stage_collect_graph(dx_engine* e, int p) {
    const auto& v = e->gev[p];
    for (size_t i = 0; i + 1 < v.size(); i += 2) {
        if (!v[i + 1]) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, v[i], v[i + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
        e->stage_ms[e->gev_stage[p][i / 2]] += ms;
        e->stage_calls[e->gev_stage[p][i / 2]] += 1;
    }
}

// kernel-level timing hooks used by the network launchers (nested inside the ST_HEUR stage)
static thread_local dx_engine* g_prof_engine = nullptr;
thread_local bool g_dx_pdl = false;      // dx_launch(): programmatic dependent launches (set while a small-batch step is captured)
void dx_prof_begin(int stage) { if (g_prof_engine) stage_begin(g_prof_engine, stage); }
void dx_prof_end() { if (g_prof_engine) stage_end(g_prof_engine); }

static int sync_ctl(dx_engine* e) {
    DX_CUDA(cudaMemcpyAsync(e->h_ctl, e->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, e->stream));
    DX_CUDA(cudaStreamSynchronize(e->stream));
    // the batch sort's key cut assumes float-like keys; a sort that had to merge > FIX_RUNS runs inside one group says these keys are
    // not (results are still exact): from the next launch on every key bit goes through the digit passes (graphs: dx_run re-captures)
    if (e->h_ctl->sort_slow && e->sort.key_cut > 0 && e->cut_sticky) e->sort.key_cut = 0;
    if (e->h_ctl->error & ~DX_DEVERR_DONE) {
        std::string m = "device capacity exceeded:";
        if (e->h_ctl->error & DX_DEVERR_NODES) m += " max_nodes";
        if (e->h_ctl->error & DX_DEVERR_OPEN) m += " max_open";
        if (e->h_ctl->error & DX_DEVERR_POP) m += " max_pop_total";
        if (e->h_ctl->error & DX_DEVERR_HDA) m += " HDA* exchange segment / receive capacity";
        if (e->h_ctl->error & DX_DEVERR_BACKUPS) m += " max_backups (drain the backup log with dx_drain_backups)";
        if (e->h_ctl->error & DX_DEVERR_PEER) m += " (reported by another rank of the HDA* search)";
        if (e->h_ctl->error & DX_DEVERR_NUMERIC) m += " network activation outside the fp16 range of the split tensor-core path (DXB200_FP32_TC=0 selects the CUDA-core fp32 network)";
        DX_FAIL(DX_ERR_CAPACITY, m);
    }
    return DX_OK;
}

// single-block sort instead of the radix passes: only when the host can bound the pushed batch (never in HDA* mode,
// where the received children are not bounded by the local batch)
static bool small_sort(const dx_engine* e) {
    return e->hda_world < 1 && e->sort_bound > 0 && e->sort_bound <= DX_SMALL_SORT;
}
// small A* iterations: CLOSED filter + scan + scatter in one block, plan + keys + sort in another (DXB200_SMALL_FUSE=0: the
// separate kernels, for A/B runs)
static bool small_fused_q(const dx_engine* e) {
    return small_sort(e) && e->fuse_small && e->is_q && !e->is_beam && !e->backups_q;
}
static bool small_fused_v(const dx_engine* e) {
    return small_sort(e) && e->fuse_small && !e->is_q && !e->is_beam && !e->eval_all;
}

static void drop_graphs(dx_engine* e) {
    for (int k = 0; k < 4; ++k) {
        if (e->graph_exec[k]) { cudaGraphExecDestroy(e->graph_exec[k]); e->graph_exec[k] = nullptr; }
        if (e->graph_prof_exec[k]) { cudaGraphExecDestroy(e->graph_prof_exec[k]); e->graph_prof_exec[k] = nullptr; }
        for (cudaEvent_t ev : e->gev[k]) if (ev) cudaEventDestroy(ev);
        e->gev[k].clear();
        e->gev_stage[k].clear();
    }
}

static int ensure_stage(dx_engine* e, size_t bytes) {
    if (bytes <= e->stage_bytes) return DX_OK;
    if (e->d_stage) {                        // grow: nothing in flight may still use the old buffer
        DX_CUDA(cudaStreamSynchronize(e->stream));
        e->allocs.erase(std::remove(e->allocs.begin(), e->allocs.end(), (void*)e->d_stage), e->allocs.end());
        cudaFree(e->d_stage);
        e->d_stage = nullptr;
        e->stage_bytes = 0;
    }
    void* q = nullptr;
    DX_CUDA(cudaMalloc(&q, bytes));
    e->allocs.push_back(q);
    e->d_stage = (uint8_t*)q;
    e->stage_bytes = bytes;
    return DX_OK;
}

static int reset_state(dx_engine* e) {
    DX_CUDA(cudaMemsetAsync(e->d_ctl, 0, sizeof(Ctl), e->stream));
    e->launches += dx_launch_closed_init(e->stream, e->closed);
    DX_CUDA(cudaMemsetAsync(e->sort.ghist, 0, DX_NUM_DIGITS * 256 * sizeof(uint32_t), e->stream));
    for (int p = 0; p < 2; ++p) {
        DX_CUDA(cudaMemsetAsync(e->pop_off[p], 0, (e->max_inst + 1) * sizeof(uint32_t), e->stream));
        DX_CUDA(cudaMemsetAsync(e->open_off[p], 0, (e->max_inst + 1) * sizeof(uint32_t), e->stream));
    }
    if (e->node_solved) DX_CUDA(cudaMemsetAsync(e->node_solved, 0, e->max_nodes, e->stream));
    if (e->bk.node_rec) DX_CUDA(cudaMemsetAsync(e->bk.node_rec, 0, (size_t)e->max_nodes * sizeof(uint32_t), e->stream));
    DX_CUDA(cudaStreamSynchronize(e->stream));
    if (e->young.m_head) {
        for (int p = 0; p < 2; ++p) {
            cudaMemsetAsync(e->young.off[p], 0, (e->max_inst + 1) * sizeof(uint32_t), e->stream);
            cudaMemsetAsync(e->young.head[p], 0, e->max_inst * sizeof(uint32_t), e->stream);
        }
        cudaMemsetAsync(e->young.m_head, 0, e->max_inst * sizeof(uint32_t), e->stream);
        DX_CUDA(cudaStreamSynchronize(e->stream));
    }
    e->young_active = false;
    e->young_big = false;
    e->y_bound = 0;
    e->m_bound = 0;
    e->n_inst = 0;
    e->parity = 0;
    e->pending = false;
    e->sort_bound = 0;
    e->removed.clear();
    return DX_OK;      // stage timers are cumulative across resets; dx_set_profiling() clears them
}

extern "C" int dx_domain_info(int32_t domain, int32_t dim, int32_t* state_bytes, int32_t* num_actions, int32_t* in_count,
                              int32_t* one_hot_depth) {
    DomainDev d;
    std::vector<uint8_t> t;
    DX_TRY(make_domain_host(domain, dim, &d, &t));
    if (state_bytes) *state_bytes = d.S;
    if (num_actions) *num_actions = d.A;
    if (in_count) *in_count = d.in_count;
    if (one_hot_depth) *one_hot_depth = d.depth;
    return DX_OK;
}

extern "C" int dx_engine_create(const dx_config* cfg, dx_engine** out) {
    if (!cfg || !out) DX_FAIL(DX_ERR_ARG, "null argument");
    if (cfg->max_instances < 1 || cfg->max_pop_total < 1 || cfg->max_nodes < 2) DX_FAIL(DX_ERR_ARG, "capacities must be positive");
    if (cfg->max_nodes >= 0x7FFFFFF0ll) DX_FAIL(DX_ERR_ARG, "max_nodes must be < 2^31");
    if (cfg->algo < DX_ALGO_GRAPH_V || cfg->algo > DX_ALGO_BEAM_Q) DX_FAIL(DX_ERR_ARG, "unknown algo");
    if (cfg->heur_mode < DX_HEUR_ZERO || cfg->heur_mode > DX_HEUR_NET_BF16) DX_FAIL(DX_ERR_ARG, "unknown heur_mode");
    if (cfg->max_backups < 0) DX_FAIL(DX_ERR_ARG, "max_backups must be >= 0");
    if (cfg->max_backups > 0 && cfg->algo != DX_ALGO_GRAPH_V && cfg->algo != DX_ALGO_GRAPH_Q)
        DX_FAIL(DX_ERR_UNSUPPORTED, "Bellman backups (max_backups > 0) are implemented for graph_v and graph_q only");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        DX_FAIL(DX_ERR_CUDA, "no CUDA device: libdxb200 has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) DX_FAIL(DX_ERR_ARG, "bad device ordinal");
    DX_CUDA(cudaSetDevice(cfg->device));
    dx_engine* e = new dx_engine();
    e->cfg = *cfg;
    int r = make_domain_host(cfg->domain, cfg->domain_dim, &e->dom, &e->table_h);
    if (r < 0) { delete e; return r; }
    e->is_q = cfg->algo == DX_ALGO_GRAPH_Q || cfg->algo == DX_ALGO_BEAM_Q;
    e->is_beam = cfg->algo == DX_ALGO_BEAM_V || cfg->algo == DX_ALGO_BEAM_Q;
    if (const char* ng = getenv("DXB200_NO_GRAPHS")) e->use_graphs = !(ng[0] == '1');
    if (const char* fs = getenv("DXB200_SMALL_FUSE")) e->fuse_small = !(fs[0] == '0');
    if (const char* pe = getenv("DXB200_POLL_EVERY")) e->poll_every = std::max(1, atoi(pe));
    if (const char* pe = getenv("DXB200_POLL_EVERY_BIG")) e->poll_every_big = std::max(1, atoi(pe));
    if (const char* pd = getenv("DXB200_PDL")) e->pdl = !(pd[0] == '0');
    if (const char* yg = getenv("DXB200_YOUNG")) e->young_enabled = !(yg[0] == '0');
    if (const char* yg = getenv("DXB200_YOUNG_BIG")) e->young_big_enabled = std::max(0, std::min(2, atoi(yg)));
    if (const char* yg = getenv("DXB200_YOUNG_EVERY")) e->young_every = std::max(0, atoi(yg));
    e->A = e->dom.A; e->SP = e->dom.SP; e->S = e->dom.S;
    e->max_inst = cfg->max_instances;
    e->max_nodes = (uint32_t)cfg->max_nodes;
    e->max_pop = (uint32_t)std::max(cfg->max_pop_total, cfg->max_instances);
    const unsigned long long mc = (unsigned long long)e->max_pop * e->A;
    if (mc >= 0x7FFFFFF0ull) { delete e; DX_FAIL(DX_ERR_ARG, "max_pop_total * actions must be < 2^31"); }
    e->max_children = e->is_q ? e->max_pop : (uint32_t)mc;   // items seen by the CLOSED filter
    e->max_sort = (uint32_t)mc;                               // entries pushed per iteration
    unsigned long long mo = cfg->max_open > 0 ? (unsigned long long)cfg->max_open
                                              : (unsigned long long)e->max_nodes * (e->is_q ? e->A : 1);
    if (mo >= 0xFFFFFFF0ull) mo = 0xFFFFFFF0ull;
    e->max_open = (uint32_t)mo;
    if (e->is_q && (unsigned long long)e->max_nodes * e->A >= 0xFFFFFFFFull) { delete e; DX_FAIL(DX_ERR_ARG, "max_nodes * actions must be < 2^32 for graph_q"); }

#define A_(ptr, count) do { int _r = e->alloc(&(ptr), (count)); if (_r < 0) { dx_engine_destroy(e); return _r; } } while (0)
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) { delete e; DX_FAIL(DX_ERR_CUDA, "stream create failed"); }
    e->own_stream = e->stream;
    if (cudaMallocHost((void**)&e->h_ctl, sizeof(Ctl)) != cudaSuccess) { delete e; DX_FAIL(DX_ERR_CUDA, "pinned alloc failed"); }
    A_(e->d_ctl, 1);
    A_(e->d_table, std::max<size_t>(e->table_h.size(), 16));
    if (!e->table_h.empty()) cudaMemcpy(e->d_table, e->table_h.data(), e->table_h.size(), cudaMemcpyHostToDevice);
    e->dom.table = e->d_table;
    const size_t N = e->max_nodes, C = e->max_children, P = e->max_pop, I = e->max_inst;
    A_(e->node_state, N * e->SP); A_(e->node_g, N); A_(e->node_h, N); A_(e->node_parent, N); A_(e->node_action, N);
    if (e->is_q) A_(e->node_q, N * e->A);
    size_t hcap = 1024;
    while (hcap < 2 * (N + C)) hcap <<= 1;   // load factor <= 0.5 even with a full node store and a full batch in flight
    e->closed.mask = (uint32_t)(hcap - 1);
    A_(e->closed.slots, hcap);
    if (!e->is_q) { A_(e->child_state, C * e->SP); A_(e->child_g, C); }
    e->eval_all = cfg->max_backups > 0 && !e->is_q;
    e->backups_q = cfg->max_backups > 0 && e->is_q;
    if (cfg->max_backups > 0) {
        const size_t K = (size_t)cfg->max_backups;
        e->bk.cap = (uint32_t)K;
        A_(e->bk.state, K * e->SP); A_(e->bk.pack, K * e->S); A_(e->bk.val, K); A_(e->bk.inst, K); A_(e->bk.act, K);
        if (e->eval_all) {
            A_(e->child_h, C + 1);
            A_(e->bk.solved, K); A_(e->bk.node, K); A_(e->bk.itr, K); A_(e->bk.child_h, K * e->A); A_(e->bk.child_id, K * e->A);
            A_(e->bk.node_rec, N);
        }
        if (e->backups_q) { A_(e->node_solved, N); A_(e->bk.node, K); A_(e->bk.itr, K); A_(e->bk.node_aux, N); }
        A_(e->her_cand, K); A_(e->her_depth, K);          // hindsight relabelling (dx_backup_deepest)
    }
    A_(e->her_best, I); A_(e->her_node, I); A_(e->her_rows, I * e->S);
    A_(e->child_slot, C + 1); A_(e->child_next, C + 1); A_(e->child_flags, C + 1); A_(e->child_prefix, C + 8); A_(e->new_inst, C + 1);
    for (int p = 0; p < 2; ++p) {
        A_(e->popped[p], P); A_(e->popped_inst[p], P); A_(e->pop_off[p], I + 1);
        A_(e->open_key[p], (size_t)e->max_open); A_(e->open_val[p], (size_t)e->max_open); A_(e->open_off[p], I + 1);
        A_(e->sort.key[p], (size_t)e->max_sort); A_(e->sort.inst[p], (size_t)e->max_sort); A_(e->sort.val[p], (size_t)e->max_sort);
    }
    A_(e->edge_tmp, P); A_(e->pop_solved, P);
    e->sort.max_tiles = dx_div_up(e->max_sort, DX_SORT_TILE);
    e->sort.num_digits = 8;
    for (uint32_t m = (uint32_t)(e->max_inst - 1); m && e->sort.num_digits < DX_NUM_DIGITS; m >>= 8) e->sort.num_digits += 1;
    A_(e->sort.hist, (size_t)256 * e->sort.max_tiles + 4096); A_(e->sort.ghist, DX_NUM_DIGITS * 256); A_(e->sort.gbase, DX_NUM_DIGITS * 256);
    A_(e->sort.status, (size_t)256 * e->sort.max_tiles + 256); A_(e->sort.tickets, DX_NUM_DIGITS + 4);
    cudaMemset(e->sort.status, 0, ((size_t)256 * e->sort.max_tiles + 256) * sizeof(unsigned long long));
    cudaMemset(e->sort.tickets, 0, (DX_NUM_DIGITS + 4) * sizeof(uint32_t));
    {
        // young-run capacity: a dozen pushed batches, never more than the main runs may hold
        e->young_cap = std::max<long long>(DX_YOUNG_BUF, std::min<long long>((long long)e->max_open, 12ll * e->max_sort));
        const size_t mt = (size_t)dx_div_up((long long)e->max_open + std::max<long long>(e->max_sort, e->young_cap), DX_SORT_TILE) + e->max_inst + 8 +
                          YT_MAX_INST_HOST;
        A_(e->sort.part_a, mt); A_(e->sort.part_i, mt);
    }
    A_(e->sort.fix_list, (size_t)e->max_sort / 2 + 2);       // (a listed inversion has a predecessor in its group: at most n / 2)
    A_(e->sort.fix_claim, (size_t)e->max_sort + 1);
    cudaMemset(e->sort.fix_claim, 0, ((size_t)e->max_sort + 1) * sizeof(uint32_t));
    e->sort.key_cut = DX_KEY_CUT;
    if (const char* c = getenv("DXB200_SORT_CUT")) e->sort.key_cut = std::max(0, std::min(40, atoi(c)));
    if (const char* c = getenv("DXB200_SORT_CUT_STICKY")) e->cut_sticky = !(c[0] == '0');
    e->sort.onesweep = 1;
    if (const char* c = getenv("DXB200_SORT_ONESWEEP")) e->sort.onesweep = (c[0] == '0') ? 0 : 1;
    A_(e->scan.block_sums, std::max<size_t>((C + 1) / 4096, ((size_t)256 * e->sort.max_tiles) / 4096) + 8);
    A_(e->ia.batch, I); A_(e->ia.weight, I); A_(e->ia.lb, I); A_(e->ia.ub_key, I); A_(e->ia.goal_node, I); A_(e->ia.itr, I);
    A_(e->ia.gen, I); A_(e->ia.active, I); A_(e->ia.pop_seq_base, I); A_(e->ia.n_open, I); A_(e->ia.m_new, I); A_(e->ia.k_pop, I);
    A_(e->ia.batch_off, I + 1); A_(e->ia.tile_off, I + 1); A_(e->ia.f_first, I); A_(e->ia.goals, I * e->SP); A_(e->ia.root, I);
    if (!e->is_beam && I <= YT_MAX_INST_HOST) {      // two-run OPEN: small-batch A* (one block, 200 KB) and big batches (A*, Q*)
        const bool big_capable = e->young_cap >= 4ll * e->max_sort && e->max_sort > DX_SMALL_SORT;
        if (!big_capable) e->young_cap = DX_YOUNG_BUF;
        if (big_capable || !e->is_q) {
            for (int p = 0; p < 2; ++p) {
                A_(e->young.key[p], (size_t)e->young_cap); A_(e->young.val[p], (size_t)e->young_cap);
                A_(e->young.off[p], I + 1); A_(e->young.head[p], I);
            }
            A_(e->young.m_head, I);
            A_(e->young.ka, I); A_(e->young.kp, I); A_(e->young.mh_old, I); A_(e->young.zeros, I);
            A_(e->young.tile_off1, I + 1); A_(e->young.tile_off2, I + 1); A_(e->young.n_tiles, 2);
            e->young.cap = (uint32_t)e->young_cap;
        }
    }
    A_(e->d_vals, std::max<size_t>((size_t)e->max_sort, P * e->A) + I * e->A);
#undef A_
    cudaDeviceSynchronize();                 // the legacy-stream table upload has landed before the engine stream runs
    r = reset_state(e);
    if (r < 0) { dx_engine_destroy(e); return r; }
    *out = e;
    return DX_OK;
}

extern "C" int dx_engine_destroy(dx_engine* e) {
    if (!e) return DX_OK;
    if (g_prof_engine == e) g_prof_engine = nullptr;
    cudaSetDevice(e->cfg.device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (int k = 0; k < 4; ++k) {
        if (e->graph_exec[k]) cudaGraphExecDestroy(e->graph_exec[k]);
        if (e->graph_prof_exec[k]) cudaGraphExecDestroy(e->graph_prof_exec[k]);
        for (cudaEvent_t ev : e->gev[k]) if (ev) cudaEventDestroy(ev);
    }
    for (void* p : e->allocs) cudaFree(p);
    if (e->net.tc_split) {
        TcSplitW* sw = (TcSplitW*)e->net.tc_split;
        cudaFree(sw->w0); cudaFree(sw->wres); cudaFree(sw->wout);
        delete sw;
    }
    dx_mlp_small_free(e->net);
    if (e->h_ctl) cudaFreeHost(e->h_ctl);
    if (e->own_stream) cudaStreamDestroy(e->own_stream);
    delete e;
    return DX_OK;
}

extern "C" int dx_engine_reset(dx_engine* e) {
    if (!e) DX_FAIL(DX_ERR_ARG, "null engine");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    return reset_state(e);
}

// ---------------------------------------------------------------------------------------------
static int need_net(dx_engine* e) {
    if ((e->cfg.heur_mode == DX_HEUR_NET_FP32 || e->cfg.heur_mode == DX_HEUR_NET_BF16) && !e->weights_loaded)
        DX_FAIL(DX_ERR_STATE, "heur_mode is NET_* but dx_load_weights has not been called");
    return DX_OK;
}

int dx_upload_net(dx_engine* e, const dx_net_desc* desc, const float* blob, int64_t n_floats);   // net_load.cu

// The loaded network over the current job: tcgen05 path (bf16, or split-fp16 for NET_FP32 when the shape is supported),
// else the CUDA-core fp32 path.  Returns the number of launches.
static int launch_net(dx_engine* e, const uint8_t* rows, const uint8_t* goals, float* out_h, float* out_q, int absolute, long long max_rows) {
    g_prof_engine = e->profiling ? e : nullptr;       // the launchers' dx_prof_begin / dx_prof_end hooks time into THIS engine
    if (e->net.small)
        return dx_launch_mlp_small(e->stream, e->dom, e->net, e->d_ctl, rows, goals, out_h, out_q, absolute, std::min<long long>(max_rows, e->mlp.cap));
    if (e->cfg.heur_mode == DX_HEUR_NET_BF16 || e->net.tc_split)
        return dx_launch_mlp_bf16(e->stream, e->dom, e->net, e->mlp, e->d_ctl, rows, goals, out_h, out_q, absolute, max_rows);
    return dx_launch_mlp_f32(e->stream, e->dom, e->net, e->mlp, e->d_ctl, rows, goals, out_h, out_q, absolute, max_rows);
}

// heuristic for the current job (ctl->nn_row_start / nn_n) over node-store rows
static int launch_heur_rows(dx_engine* e, const uint8_t* rows, float* out_h, float* out_q, long long max_rows) {
    const int mode = e->cfg.heur_mode;
    if (mode == DX_HEUR_ZERO) {
        dx_launch(k_fill_zero_vals, std::min(dx_div_up(max_rows, 256), 148 * 4), 256, 0, e->stream, (const Ctl*)e->d_ctl, out_h, out_q, e->A);
        e->launches += 1;
    } else if (mode == DX_HEUR_NET_FP32 || mode == DX_HEUR_NET_BF16) {
        e->launches += launch_net(e, rows, e->ia.goals, out_h, out_q, 1, max_rows);
    }
    return DX_OK;
}
static int launch_heur_nodes(dx_engine* e, long long max_rows) {
    return launch_heur_rows(e, e->node_state, e->node_h, e->node_q, max_rows);
}
// the heuristic job of an iteration: the stored kept children / generated nodes, or -- evaluate-all engines -- every generated
// child in flat child order (rows [0, n_children) of child_state; the job was set by step_v_a)
static int launch_heur_step(dx_engine* e, long long max_rows) {
    if (e->eval_all) return launch_heur_rows(e, e->child_state, e->child_h, nullptr, max_rows);
    return launch_heur_nodes(e, max_rows);
}

// One-hot inputs must be < depth: the reference's F.one_hot raises on a larger value (pytorch_models.py:15-24); the device
// kernels would otherwise have to clamp silently.  Checked on the host for every row that enters through the C ABI.
static int check_rows_in_range(const dx_engine* e, const uint8_t* states, int64_t n) {
    if (!e->weights_loaded || e->net.depth < 2 || !states) return DX_OK;
    const int S = e->S, lim = std::min(e->net.in_count, S), depth = e->net.depth;
    for (int64_t r = 0; r < n; ++r)
        for (int i = 0; i < lim; ++i)
            if (states[(size_t)r * S + i] >= depth)
                DX_FAIL(DX_ERR_ARG, "state " + std::to_string(r) + " has value " + std::to_string((int)states[(size_t)r * S + i]) +
                                        " at input " + std::to_string(i) + ": outside the network's one-hot depth " + std::to_string(depth));
    return DX_OK;
}

extern "C" int dx_remove_instances(dx_engine* e, const int32_t* slots, int32_t n) {
    if (!e || (n > 0 && !slots)) DX_FAIL(DX_ERR_ARG, "null argument");
    if (n <= 0) return DX_OK;
    if (e->pending) DX_FAIL(DX_ERR_STATE, "dx_remove_instances between dx_step_begin and dx_step_end");
    if (e->hda_world >= 1) DX_FAIL(DX_ERR_UNSUPPORTED, "HDA* mode holds one instance: use dx_engine_reset");
    for (int j = 0; j < n; ++j)
        if (slots[j] < 0 || slots[j] >= e->n_inst) DX_FAIL(DX_ERR_ARG, "instance slot out of range");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_TRY(sync_ctl(e));
    const int I = e->n_inst, p = e->parity, q = p ^ 1;
    cudaStream_t s = e->stream;
    std::vector<uint8_t> act(I);
    std::vector<uint32_t> off(I + 1), noff(I + 1);
    DX_CUDA(cudaMemcpy(act.data(), e->ia.active, I, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(off.data(), e->pop_off[p], (size_t)(I + 1) * 4, cudaMemcpyDeviceToHost));
    e->removed.resize(e->max_inst, 0);
    for (int j = 0; j < n; ++j) { act[slots[j]] = 0; e->removed[slots[j]] = 1; }
    uint32_t acc = 0, unfinished = 0;
    for (int i = 0; i < I; ++i) {                 // removed instances give up their popped entries (a finished one keeps its last
        noff[i] = acc;                            // `_nodes_curr` readable, Instance.get_nodes)
        if (!e->removed[i]) acc += off[i + 1] - off[i];
        if (act[i]) ++unfinished;
    }
    noff[I] = acc;
    DX_CUDA(cudaMemcpyAsync(e->ia.active, act.data(), I, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->pop_off[q], noff.data(), (size_t)(I + 1) * 4, cudaMemcpyHostToDevice, s));
    k_compact_popped<<<std::min(I, 148 * 8), 256, 0, s>>>(e->pop_off[p], e->pop_off[q], I, e->popped[p], e->popped_inst[p], e->popped[q],
                                                          e->popped_inst[q]);
    if (acc) {
        DX_CUDA(cudaMemcpyAsync(e->popped[p], e->popped[q], (size_t)acc * 4, cudaMemcpyDeviceToDevice, s));
        DX_CUDA(cudaMemcpyAsync(e->popped_inst[p], e->popped_inst[q], (size_t)acc * 4, cudaMemcpyDeviceToDevice, s));
    }
    DX_CUDA(cudaMemcpyAsync(e->pop_off[p], e->pop_off[q], (size_t)(I + 1) * 4, cudaMemcpyDeviceToDevice, s));
    k_after_remove<<<1, 1, 0, s>>>(e->d_ctl, acc, acc * (e->is_q ? 1u : (uint32_t)e->A), unfinished);
    e->launches += 2;
    DX_CUDA(cudaStreamSynchronize(s));            // (the host vectors above must outlive the copies)
    return DX_OK;
}

// ---- young-run OPEN (YoungBufs, dx_internal.cuh) ------------------------------------------------------------------------
__global__ void k_young_copyback(const Ctl* __restrict__ ctl, const unsigned long long* __restrict__ key_src, const uint32_t* __restrict__ val_src,
                                 const uint32_t* __restrict__ off_src, unsigned long long* __restrict__ key_dst, uint32_t* __restrict__ val_dst,
                                 uint32_t* __restrict__ off_dst) {
    if (ctl->error) return;
    const uint32_t n = ctl->open_total, I = ctl->n_inst;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    // 16-byte copies (the buffers are cudaMalloc-aligned): two keys / four values per access, then the odd tail
    const uint4* __restrict__ k4 = reinterpret_cast<const uint4*>(key_src);
    const uint4* __restrict__ v4 = reinterpret_cast<const uint4*>(val_src);
    uint4* __restrict__ dk4 = reinterpret_cast<uint4*>(key_dst);
    uint4* __restrict__ dv4 = reinterpret_cast<uint4*>(val_dst);
    for (uint32_t t = tid; t < n / 2; t += nth) dk4[t] = k4[t];
    for (uint32_t t = tid; t < n / 4; t += nth) dv4[t] = v4[t];
    if (tid == 0 && (n & 1)) key_dst[n - 1] = key_src[n - 1];
    if (tid < (n & 3)) val_dst[(n & ~3u) + tid] = val_src[(n & ~3u) + tid];
    for (uint32_t t = tid; t <= I; t += nth) off_dst[t] = off_src[t];
}

// main run <- merge(main run remainder, young runs); the main run's home is open_*[0]: merged into [1], then the pointers are swapped
static int young_compact(dx_engine* e) {
    cudaStream_t s = e->stream;
    const int p = e->parity;
    const int max_tiles = dx_div_up((long long)e->max_open + e->young_cap, DX_SORT_TILE) + e->max_inst;
    e->launches += dx_launch_young_compact(s, e->d_ctl, e->ia, e->sort, e->young, p, e->open_key[0], e->open_val[0], e->open_off[0],
                                           e->open_key[1], e->open_val[1], e->open_off[1], e->popped[p ^ 1], e->popped_inst[p ^ 1],
                                           e->pop_off[p], max_tiles, (uint32_t)e->max_open);
    std::swap(e->open_key[0], e->open_key[1]);        // the merged runs become the main runs: no copy back (graphs are kept per identity
    std::swap(e->open_val[0], e->open_val[1]);        // of the pair, dx_run picks / captures the set of the current one)
    std::swap(e->open_off[0], e->open_off[1]);
    e->open_id ^= 1;
    e->m_bound += e->y_bound;
    e->y_bound = 0;
    return DX_OK;
}
// before an iteration is launched: the merged young runs (young + pushed batch) must fit the block's buffer (one-block tail) /
// the young buffers (grid-wide merges), and the young runs are folded into the main runs when rewriting them every iteration
// starts to cost more than the compaction saves: per iteration the young merge moves ~ y entries and a compaction every n
// iterations (y = n * b) moves (M + y) / n, least at y = sqrt(2 M b) for main runs of M entries and pushes of b per iteration
static int young_before_step(dx_engine* e) {
    if (!e->young_active) return DX_OK;
    const long long b = std::max<long long>(e->sort_bound, 1);
    long long limit = DX_YOUNG_BUF;
    if (e->young_big) {
        const long long best = (long long)std::sqrt(2.0 * (double)e->m_bound * (double)b);
        limit = std::min<long long>(e->young_cap, (e->young_every > 0 ? (long long)e->young_every * b : std::max<long long>(4 * b, best)) + b);
    }
    if (e->y_bound + e->sort_bound > limit) DX_TRY(young_compact(e));
    e->y_bound += e->sort_bound;
    return DX_OK;
}
// leave young mode (the batch outgrew the small regime): everything back into ONE run, in the buffer the plain path reads next
static int young_flush(dx_engine* e) {
    if (!e->young_active) return DX_OK;
    DX_TRY(young_compact(e));
    if (e->parity != 0)
        k_young_copyback<<<148 * 8, 256, 0, e->stream>>>(e->d_ctl, e->open_key[0], e->open_val[0], e->open_off[0], e->open_key[1], e->open_val[1],
                                                          e->open_off[1]);
    DX_CUDA(cudaStreamSynchronize(e->stream));
    e->young_active = false;
    e->young_big = false;
    return DX_OK;
}

extern "C" int dx_add_instances(dx_engine* e, const uint8_t* states, const uint8_t* goals, int32_t n,
                                const int32_t* batch_sizes, const double* weights, const float* root_vals) {
    if (!e || (n > 0 && (!states || !goals))) DX_FAIL(DX_ERR_ARG, "null argument");
    if (n <= 0) return DX_OK;
    if (e->pending) DX_FAIL(DX_ERR_STATE, "dx_add_instances between dx_step_begin and dx_step_end");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_TRY(need_net(e));
    if (e->n_inst + n > e->max_inst)
        DX_FAIL(DX_ERR_CAPACITY, "max_instances exceeded: instance slots are numbered for the engine's life (finished / removed instances keep "
                                 "theirs until dx_engine_reset); size max_instances for every instance added between two resets");
    DX_TRY(check_rows_in_range(e, states, n));
    if (e->cfg.heur_mode == DX_HEUR_HOST && e->is_q && !root_vals) DX_FAIL(DX_ERR_ARG, "graph_q in HOST mode needs root_vals");
    DX_TRY(sync_ctl(e));
    Ctl c = *e->h_ctl;
    if ((unsigned long long)c.node_count + n > e->max_nodes) DX_FAIL(DX_ERR_CAPACITY, "max_nodes exceeded");
    if ((unsigned long long)c.n_pop + n > e->max_pop) DX_FAIL(DX_ERR_CAPACITY, "max_pop_total exceeded");
    long long bsum = 0;
    for (int i = 0; i < n; ++i) {
        const int b = batch_sizes ? batch_sizes[i] : 1;
        if (b < 1) DX_FAIL(DX_ERR_ARG, "batch size must be >= 1");
        bsum += b;
    }
    {   // total batch of LIVE instances (not finished, not removed) must fit max_pop_total
        std::vector<int32_t> hb(e->n_inst);
        std::vector<uint8_t> ha(e->n_inst);
        if (e->n_inst) {
            DX_CUDA(cudaMemcpy(hb.data(), e->ia.batch, e->n_inst * sizeof(int32_t), cudaMemcpyDeviceToHost));
            DX_CUDA(cudaMemcpy(ha.data(), e->ia.active, e->n_inst, cudaMemcpyDeviceToHost));
        }
        for (int i = 0; i < e->n_inst; ++i) if (ha[i]) bsum += hb[i];
        if (bsum > (long long)e->max_pop) DX_FAIL(DX_ERR_CAPACITY, "sum of the live instances' batch sizes exceeds max_pop_total");
    }
    // no iteration can push more entries than every live instance's full batch expanded -- but the CURRENT popped list may still hold
    // the last pops of instances that finished and were not removed (they stay readable, Instance.get_nodes): the next iteration's
    // kernels see those entries too, so the bound that selects the small-batch kernels must cover them
    e->sort_bound = std::max<long long>(bsum, (long long)c.n_pop + n) * e->A;
    {   // young-run OPEN: entered with the first instances after a reset, left for good when the batch outgrows the small regime
        const bool have = e->young_enabled && e->young.m_head != nullptr && e->hda_world < 1;
        const bool want_small = have && small_fused_v(e);
        const bool want_big = have && e->young_big_enabled >= (e->is_q ? 1 : 2) && !small_sort(e) && e->young_cap >= 4 * e->sort_bound &&
                              e->young_cap > DX_YOUNG_BUF;
        if (e->n_inst == 0) {
            e->young_active = want_small || want_big;
            e->young_big = want_big;
            e->y_bound = 0;
            e->m_bound = 0;
        } else if (e->young_active && !(want_small || want_big)) {
            DX_TRY(young_flush(e));
            DX_TRY(sync_ctl(e));
            c = *e->h_ctl;                   // (the compaction rewrote open_total / young_total)
        } else if (e->young_active) {
            e->young_big = want_big;         // same two runs, other kernels (young_before_step compacts first if the block's buffer is smaller)
        }
    }
    const int SP = e->SP, S = e->S, I0 = e->n_inst, p = e->parity;
    std::vector<uint8_t> rows((size_t)n * SP, 0), grow((size_t)n * SP, 0);
    std::vector<int32_t> hb(n);
    std::vector<double> hw(n), hlb(n, 0.0);
    std::vector<unsigned long long> hub(n, DX_EMPTY64), hgen(n, 0ull);
    std::vector<uint32_t> hgoal(n, DX_NIL), hz(n, 0u), hpar(n, DX_NIL), hpop(n), hpinst(n), hoff(n + 1), hooff(n + 1);
    std::vector<uint8_t> hact(n, 1), hacn(n, 0);
    std::vector<float> hg(n, 0.f);
    for (int i = 0; i < n; ++i) {
        memcpy(&rows[(size_t)i * SP], states + (size_t)i * S, S);
        const uint32_t inst = I0 + i;
        memcpy(&rows[(size_t)i * SP + SP - 4], &inst, 4);
        memcpy(&grow[(size_t)i * SP], goals + (size_t)i * S, S);
        hb[i] = batch_sizes ? batch_sizes[i] : 1;
        hw[i] = weights ? weights[i] : 1.0;
        hpop[i] = c.node_count + i;
        hpinst[i] = inst;
        hoff[i] = c.n_pop + i;
        hooff[i] = c.open_total;
    }
    hoff[n] = c.n_pop + n;
    hooff[n] = c.open_total;
    cudaStream_t s = e->stream;
    const uint32_t first = c.node_count;
    DX_CUDA(cudaMemcpyAsync(e->node_state + (size_t)first * SP, rows.data(), rows.size(), cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.goals + (size_t)I0 * SP, grow.data(), grow.size(), cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->node_g + first, hg.data(), n * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->node_h + first, hg.data(), n * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->node_parent + first, hpar.data(), n * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->node_action + first, hacn.data(), n, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.batch + I0, hb.data(), n * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.weight + I0, hw.data(), n * 8, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.lb + I0, hlb.data(), n * 8, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.ub_key + I0, hub.data(), n * 8, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.goal_node + I0, hgoal.data(), n * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.root + I0, hpop.data(), n * 4, cudaMemcpyHostToDevice, s));      // (hpop = the new roots' node ids)
    DX_CUDA(cudaMemcpyAsync(e->ia.itr + I0, hz.data(), n * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.gen + I0, hgen.data(), n * 8, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.active + I0, hact.data(), n, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.pop_seq_base + I0, hz.data(), n * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->ia.n_open + I0, hz.data(), n * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->popped[p] + c.n_pop, hpop.data(), n * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->popped_inst[p] + c.n_pop, hpinst.data(), n * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->pop_off[p] + I0, hoff.data(), (n + 1) * 4, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaMemcpyAsync(e->open_off[e->young_active ? 0 : p] + I0, hooff.data(), (n + 1) * 4, cudaMemcpyHostToDevice, s));
    std::vector<uint32_t> hyoff(n + 1, c.young_total);
    if (e->young_active) {
        DX_CUDA(cudaMemcpyAsync(e->young.off[p] + I0, hyoff.data(), (n + 1) * 4, cudaMemcpyHostToDevice, s));
        DX_CUDA(cudaMemcpyAsync(e->young.head[p] + I0, hz.data(), n * 4, cudaMemcpyHostToDevice, s));
        DX_CUDA(cudaMemcpyAsync(e->young.m_head + I0, hz.data(), n * 4, cudaMemcpyHostToDevice, s));
    }
    // control block
    Ctl nc = c;
    nc.n_inst = I0 + n;
    nc.n_pop = c.n_pop + n;
    nc.n_children = nc.n_pop * (e->is_q ? 1 : e->A);
    nc.node_count = c.node_count + n;
    nc.node_base = nc.node_count;
    nc.n_unfinished = c.n_unfinished + n;
    if (nc.epoch == 0) nc.epoch = 1;
    nc.nn_row_start = first;
    nc.nn_n = n;
    *e->h_ctl = nc;
    DX_CUDA(cudaMemcpyAsync(e->d_ctl, e->h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, s));
    if (!e->is_q) e->launches += dx_launch_closed_seed(s, e->dom, e->closed, e->node_state, first, n);   // graph_search.py:119-121
    // root values (graph_search.py:103-107): one heuristic call over the new roots
    if (e->cfg.heur_mode == DX_HEUR_HOST) {
        if (root_vals) {
            const size_t cnt = (size_t)n * (e->is_q ? e->A : 1);
            DX_CUDA(cudaMemcpyAsync(e->d_vals, root_vals, cnt * 4, cudaMemcpyHostToDevice, s));
            k_set_host_vals<<<dx_div_up(n, 128), 128, 0, s>>>(e->d_ctl, e->d_vals, e->node_h, e->node_q, e->A, e->is_q ? 1 : 0);
            e->launches += 1;
        }
    } else {
        DX_TRY(launch_heur_nodes(e, n));
    }
    DX_CUDA(cudaStreamSynchronize(s));
    e->n_inst += n;
    return DX_OK;
}

// ---------------------------------------------------------------------------------------------
static int step_v_filter(dx_engine* e);
// A* phase A: everything up to (and including) storing the kept children.
static int step_v_a(dx_engine* e) {
    cudaStream_t s = e->stream;
    const int p = e->parity;
    stage_begin(e, ST_EXPAND);
    e->launches += dx_launch_expand(s, e->dom, e->d_ctl, e->ia, e->node_state, e->node_g, e->popped[p], e->popped_inst[p],
                                    e->pop_off[p], e->child_state, e->child_g, e->pop_solved, e->max_pop);
    if (!small_fused_v(e))             // (the fused small-iteration kernel resolves the goal nodes itself)
        e->launches += dx_launch_goal_resolve(s, e->d_ctl, e->ia, e->node_g, e->popped[p], e->popped_inst[p], e->pop_off[p],
                                              e->pop_solved, e->max_pop);
    stage_end(e);
    if (e->eval_all) {                 // the reference order (base/pathfinding.py:362-371): evaluate all children, then filter
        k_set_job_children<<<1, 1, 0, s>>>(e->d_ctl);
        e->launches += 1;
        return DX_OK;
    }
    return step_v_filter(e);
}

// A*: CLOSED filter + storing the kept children (phase A; evaluate-all engines run it after the heuristic, in phase B)
static int step_v_filter(dx_engine* e) {
    cudaStream_t s = e->stream;
    const int p = e->parity;
    if (small_fused_v(e)) {
        stage_begin(e, ST_CLOSED);
        e->launches += dx_launch_small_filter_v(s, e->dom, e->closed, e->d_ctl, e->ia, e->child_state, e->child_g, e->child_slot,
                                                e->child_next, e->child_flags, e->child_prefix, e->popped[p], e->popped_inst[p],
                                                e->node_state, e->node_g, e->node_parent, e->node_action, e->new_inst,
                                                e->max_nodes, e->pop_off[p], e->pop_solved);
        stage_end(e);
        return DX_OK;
    }
    if (e->eval_all) {
        stage_begin(e, ST_BOOK);
        e->launches += dx_launch_bellman_backup(s, e->dom, e->d_ctl, e->ia, e->node_state, e->popped[p], e->popped_inst[p],
                                                e->pop_solved, e->child_h, e->bk, e->max_pop);
        stage_end(e);
    }
    stage_begin(e, ST_CLOSED);
    if (e->is_beam)
        e->launches += dx_launch_beam_keep(s, e->d_ctl, e->child_g, e->child_flags, e->max_children);
    else
        e->launches += dx_launch_closed_filter(s, e->dom, e->closed, e->d_ctl, e->ia, e->node_state, e->node_g, e->child_state,
                                               e->child_g, nullptr, e->popped_inst[p], e->child_slot, e->child_next, e->child_flags,
                                               e->max_children, e->A);
    stage_end(e);
    stage_begin(e, ST_SCATTER);
    e->launches += dx_launch_scan_flags(s, e->child_flags, e->child_prefix, &e->d_ctl->n_children, &e->d_ctl->n_kept,
                                        e->max_children, e->scan);
    k_after_scan_v<<<1, 1, 0, s>>>(e->d_ctl, e->max_nodes);
    e->launches += 1;
    e->launches += dx_launch_scatter_kept(s, e->dom, e->closed, e->d_ctl, e->child_state, e->child_g, e->child_slot,
                                          e->child_flags, e->child_prefix, e->popped[p], e->popped_inst[p], e->node_state,
                                          e->node_g, e->node_parent, e->node_action, e->new_inst, e->max_children,
                                          e->eval_all ? e->child_h : nullptr, e->node_h);
    if (e->eval_all) e->launches += dx_launch_backup_children(s, e->dom, e->d_ctl, e->child_flags, e->child_prefix, e->bk, e->max_children);
    stage_end(e);
    return DX_OK;
}

static int step_v_b(dx_engine* e) {
    cudaStream_t s = e->stream;
    const int p = e->parity, q = p ^ 1;
    if (e->eval_all) DX_TRY(step_v_filter(e));
    stage_begin(e, ST_SORT);
    const int beam = e->is_beam ? 1 : 0;
    if (e->young_active && !e->young_big) {                 // the whole OPEN side of the iteration in one block (k_small_tail_v)
        e->launches += dx_launch_small_tail_v(s, e->d_ctl, e->ia, e->child_prefix, e->pop_off[p], e->pop_off[q], e->open_key[0],
                                              e->open_val[0], e->open_off[0], e->young, p, e->A, (uint32_t)e->max_pop, e->node_g, e->node_h,
                                              e->new_inst, e->popped[q], e->popped_inst[q], e->halt_flag ? DX_END_HALT : 0);
        stage_end(e);
        e->parity = q;
        return DX_OK;
    }
    if (small_fused_v(e)) {
        e->launches += dx_launch_small_push_v(s, e->d_ctl, e->ia, e->child_prefix, e->pop_off[p], e->pop_off[q], e->open_off[p],
                                              e->open_off[q], e->A, e->max_pop, e->max_open, e->node_g, e->node_h, e->new_inst,
                                              e->sort);
    } else {
        if (e->young_big)
            e->launches += dx_launch_young_plan_b(s, e->d_ctl, e->ia, e->child_prefix, e->pop_off[p], e->pop_off[q], e->open_off[0],
                                                  e->young, p, e->A, 1, e->max_pop);
        else
            e->launches += dx_launch_plan(s, e->d_ctl, e->ia, e->child_prefix, e->pop_off[p], e->pop_off[q], e->open_off[p],
                                          e->open_off[q], e->A, 1, e->max_pop, e->max_open, 0, beam);
        e->launches += dx_launch_make_keys(s, e->d_ctl, e->ia, e->node_g, e->node_h, e->new_inst, e->sort, e->max_sort, beam);
        e->launches += dx_launch_sort(s, e->d_ctl, e->sort, e->scan, e->max_sort, small_sort(e));
    }
    stage_end(e);
    stage_begin(e, ST_MERGE);
    if (e->young_big) {
        e->launches += dx_launch_young_iter(s, e->d_ctl, e->ia, e->sort, e->open_key[0], e->open_val[0], e->open_off[0], e->young, p,
                                            e->popped[q], e->popped_inst[q], e->pop_off[q], (int)e->max_pop);
    } else {
        const int max_tiles = dx_div_up((long long)e->max_open + e->max_sort, DX_SORT_TILE) + e->max_inst;
        e->launches += dx_launch_merge(s, e->d_ctl, e->ia, e->sort, e->open_key[p], e->open_val[p], e->open_off[p], e->open_key[q],
                                       e->open_val[q], e->open_off[q], e->popped[q], e->popped_inst[q], e->pop_off[q], max_tiles, beam ? 0 : 1);
    }
    stage_end(e);
    stage_begin(e, ST_BOOK);
    e->launches += dx_launch_end_iter(s, e->d_ctl, e->ia, e->pop_off[p], e->pop_off[q], e->A, (beam ? 2 : 0) | (e->halt_flag ? DX_END_HALT : 0));
    stage_end(e);
    e->parity = q;
    return DX_OK;
}

static int step_q_pop(dx_engine* e);
// Q* phase A: up to (and including) generating the children of the popped edges.
static int step_q_a(dx_engine* e) {
    cudaStream_t s = e->stream;
    const int p = e->parity, q = p ^ 1;
    if (small_fused_q(e)) {
        stage_begin(e, ST_CLOSED);
        e->launches += dx_launch_small_filter_q(s, e->dom, e->closed, e->d_ctl, e->ia, e->node_state, e->node_g, e->popped[p],
                                                e->popped_inst[p], e->pop_off[p], e->pop_solved, e->child_slot, e->child_next,
                                                e->child_flags, e->child_prefix);
        stage_end(e);
        stage_begin(e, ST_SORT);
        e->launches += dx_launch_small_push_q(s, e->d_ctl, e->ia, e->child_prefix, e->pop_off[p], e->pop_off[q], e->open_off[p],
                                              e->open_off[q], e->A, e->max_pop, e->max_open, e->node_g, e->node_q, e->popped[p],
                                              e->popped_inst[p], e->child_flags, e->sort);
        stage_end(e);
        return step_q_pop(e);
    }
    stage_begin(e, ST_EXPAND);
    e->launches += dx_launch_check_solved(s, e->dom, e->d_ctl, e->ia, e->node_state, e->node_g, e->popped[p], e->popped_inst[p],
                                          e->pop_off[p], e->pop_solved, e->max_pop);
    e->launches += dx_launch_goal_resolve(s, e->d_ctl, e->ia, e->node_g, e->popped[p], e->popped_inst[p], e->pop_off[p],
                                          e->pop_solved, e->max_pop);
    if (e->backups_q)
        e->launches += dx_launch_mark_solved(s, e->d_ctl, e->popped[p], e->pop_off[p], e->pop_solved, e->node_solved, e->max_pop);
    stage_end(e);
    const int beam = e->is_beam ? 1 : 0;
    stage_begin(e, ST_CLOSED);
    if (beam)
        e->launches += dx_launch_beam_keep_q(s, e->d_ctl, e->ia, e->popped_inst[p], e->child_flags, e->max_children);
    else
        e->launches += dx_launch_closed_filter(s, e->dom, e->closed, e->d_ctl, e->ia, e->node_state, e->node_g, nullptr, nullptr,
                                               e->popped[p], e->popped_inst[p], e->child_slot, e->child_next, e->child_flags,
                                               e->max_children, 1);
    stage_end(e);
    stage_begin(e, ST_SCATTER);
    e->launches += dx_launch_scan_flags(s, e->child_flags, e->child_prefix, &e->d_ctl->n_children, &e->d_ctl->n_kept,
                                        e->max_children, e->scan);
    if (!beam)
        e->launches += dx_launch_closed_commit(s, e->closed, e->d_ctl, e->node_g, e->popped[p], e->child_slot, e->child_flags,
                                               e->max_children);
    stage_end(e);
    stage_begin(e, ST_SORT);
    if (e->young_big)
        e->launches += dx_launch_young_plan_b(s, e->d_ctl, e->ia, e->child_prefix, e->pop_off[p], e->pop_off[q], e->open_off[0],
                                              e->young, p, 1, e->A, e->max_pop);
    else
        e->launches += dx_launch_plan(s, e->d_ctl, e->ia, e->child_prefix, e->pop_off[p], e->pop_off[q], e->open_off[p],
                                      e->open_off[q], 1, e->A, e->max_pop, e->max_open, 0, beam);
    e->launches += dx_launch_make_keys_q(s, e->d_ctl, e->ia, e->node_g, e->node_q, e->popped[p], e->popped_inst[p],
                                         e->child_flags, e->child_prefix, e->sort, e->A, e->max_pop, beam);
    e->launches += dx_launch_sort(s, e->d_ctl, e->sort, e->scan, e->max_sort, small_sort(e));
    stage_end(e);
    return step_q_pop(e);
}

// Q*: merge the pushed edges into OPEN, pop, generate the popped edges' children
static int step_q_pop(dx_engine* e) {
    cudaStream_t s = e->stream;
    const int p = e->parity, q = p ^ 1;
    const int beam = e->is_beam ? 1 : 0;
    stage_begin(e, ST_MERGE);
    if (e->young_big) {
        e->launches += dx_launch_young_iter(s, e->d_ctl, e->ia, e->sort, e->open_key[0], e->open_val[0], e->open_off[0], e->young, p,
                                            e->edge_tmp, e->popped_inst[q], e->pop_off[q], (int)e->max_pop);
    } else {
        const int max_tiles = dx_div_up((long long)e->max_open + e->max_sort, DX_SORT_TILE) + e->max_inst;
        e->launches += dx_launch_merge(s, e->d_ctl, e->ia, e->sort, e->open_key[p], e->open_val[p], e->open_off[p], e->open_key[q],
                                       e->open_val[q], e->open_off[q], e->edge_tmp, e->popped_inst[q], e->pop_off[q], max_tiles, beam ? 0 : 1);
    }
    dx_launch(k_after_merge_q, 1, 1, 0, s, e->d_ctl, (const uint32_t*)e->pop_off[q], (uint32_t)e->max_nodes);
    e->launches += 1;
    e->launches += dx_launch_expand_edges(s, e->dom, e->d_ctl, e->edge_tmp, e->popped[q], e->node_state, e->node_g,
                                          e->node_parent, e->node_action, e->max_pop);
    stage_end(e);
    return DX_OK;
}

static int step_q_b(dx_engine* e) {
    const int p = e->parity, q = p ^ 1;
    stage_begin(e, ST_BOOK);
    if (e->backups_q)
        e->launches += dx_launch_bellman_backup_q(e->stream, e->dom, e->d_ctl, e->node_state, e->node_parent, e->node_action,
                                                  e->node_h, e->node_solved, e->popped_inst[q], e->bk, e->max_pop);
    e->launches += dx_launch_end_iter(e->stream, e->d_ctl, e->ia, e->pop_off[p], e->pop_off[q], e->A, (e->is_beam ? 3 : 1) | (e->halt_flag ? DX_END_HALT : 0));
    stage_end(e);
    e->parity = q;
    return DX_OK;
}

static int full_step(dx_engine* e) {
    if (e->is_q) {
        DX_TRY(step_q_a(e));
        stage_begin(e, ST_HEUR);
        DX_TRY(launch_heur_nodes(e, e->max_pop));
        stage_end(e);
        DX_TRY(step_q_b(e));
    } else {
        DX_TRY(step_v_a(e));
        stage_begin(e, ST_HEUR);
        DX_TRY(launch_heur_step(e, e->max_children));
        stage_end(e);
        DX_TRY(step_v_b(e));
    }
    return DX_OK;
}

static int build_step_graphs(dx_engine* e, bool prof) {
    e->graph_small = small_sort(e);
    e->graph_young = e->young_active ? (e->young_big ? 2 : 1) : 0;
    e->graph_cut = e->sort.key_cut;
    const int p0 = e->parity;
    const long long l0 = e->launches;
    for (int k = 0; k < 2; ++k) {
        const int p = e->parity;
        cudaGraph_t g = nullptr;
        DX_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
        const int slot = 2 * e->open_id + p;
        e->cap_prof = prof ? slot : -1;
        e->halt_flag = !prof;                        // (only the graphs dx_run launches in bursts)
        g_dx_pdl = !prof && e->graph_small && e->pdl && (small_fused_v(e) || small_fused_q(e));
        const int rc = full_step(e);                 // flips e->parity
        g_dx_pdl = false;
        e->halt_flag = false;
        e->cap_prof = -1;
        e->ev_open.clear();
        cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
        if (rc < 0) { if (g) cudaGraphDestroy(g); return rc; }
        if (ce != cudaSuccess) DX_FAIL(DX_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(ce));
        if (prof && e->graph_prof_exec[slot]) { cudaGraphExecDestroy(e->graph_prof_exec[slot]); e->graph_prof_exec[slot] = nullptr; }
        if (!prof && e->graph_exec[slot]) { cudaGraphExecDestroy(e->graph_exec[slot]); e->graph_exec[slot] = nullptr; }
        ce = cudaGraphInstantiate(prof ? &e->graph_prof_exec[slot] : &e->graph_exec[slot], g, 0);
        cudaGraphDestroy(g);
        if (ce != cudaSuccess) DX_FAIL(DX_ERR_CUDA, std::string("graph instantiate failed: ") + cudaGetErrorString(ce));
    }
    e->graph_launches = (int)((e->launches - l0) / 2);
    e->launches = l0;
    e->parity = p0;
    return DX_OK;
}

extern "C" int dx_run(dx_engine* e, int32_t max_iters, int32_t* iters_done) {
    if (!e) DX_FAIL(DX_ERR_ARG, "null engine");
    if (e->cfg.heur_mode == DX_HEUR_HOST) DX_FAIL(DX_ERR_STATE, "dx_run needs a device heuristic (use dx_step_begin/end in HOST mode)");
    if (e->pending) DX_FAIL(DX_ERR_STATE, "step pending");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_TRY(need_net(e));
    int done = 0;
    DX_TRY(sync_ctl(e));
    if (!e->run_ev[0]) { cudaEventCreate(&e->run_ev[0]); cudaEventCreate(&e->run_ev[1]); }
    // The iteration is a fixed launch sequence (all sizes live in the device control block), so it is captured
    // once per buffer parity into a CUDA graph and replayed; per-kernel profiling needs the plain launches.
    const bool use_graph = e->use_graphs, prof = e->profiling != 0;
    bool have_graphs = false;
    for (int k = 0; k < 4; ++k) have_graphs = have_graphs || e->graph_exec[k] || e->graph_prof_exec[k];
    if (have_graphs && (e->graph_small != small_sort(e) || e->graph_young != (e->young_active ? (e->young_big ? 2 : 1) : 0) ||
                                                        e->graph_cut != e->sort.key_cut))
        drop_graphs(e);                              // captured with the other sort / OPEN structure
    if (use_graph && !prof && !e->graph_exec[2 * e->open_id]) DX_TRY(build_step_graphs(e, false));
    if (use_graph && prof && e->prof_graph_level != e->profiling) {     // captured at another level
        for (int k = 0; k < 4; ++k) {
            if (e->graph_prof_exec[k]) cudaGraphExecDestroy(e->graph_prof_exec[k]);
            e->graph_prof_exec[k] = nullptr;
            for (cudaEvent_t ev : e->gev[k]) if (ev) cudaEventDestroy(ev);
            e->gev[k].clear();
            e->gev_stage[k].clear();
        }
    }
    if (use_graph && prof && !e->graph_prof_exec[2 * e->open_id]) {
        stage_collect(e);
        DX_TRY(build_step_graphs(e, true));
        e->prof_graph_level = e->profiling;
    }
    cudaEventRecord(e->run_ev[0], e->stream);
    // Small-batch searches (an iteration is ~100 us of kernels) launch a BURST of iterations per host poll: k_end_iter raises
    // DX_DEVERR_DONE in the iteration that finishes the last instance, which turns the rest of the burst into no-ops, so the
    // search still stops in exactly the iteration the reference's loop does; the host reads how many iterations really ran.
    // (big batches: an iteration is 1-10 ms, the host's poll costs ~30 us of idle GPU per iteration -- 2 % of a cube3 Q* iteration)
    const int burst_max = (use_graph && !prof) ? (e->graph_small ? e->poll_every : e->poll_every_big) : 1;
    int last_slot = 0;
    while (done < max_iters && e->h_ctl->n_unfinished > 0) {
        if (use_graph) {
            const int burst = std::min(burst_max, (int)max_iters - done);
            const uint32_t itr0 = e->h_ctl->itr;
            for (int b = 0; b < burst; ++b) {
                const int keep = e->parity;          // (the compaction reads the buffers of the iteration about to run)
                e->parity = keep ^ (b & 1);
                const int rcy = young_before_step(e);
                e->parity = keep;
                DX_TRY(rcy);
                last_slot = 2 * e->open_id + (e->parity ^ (b & 1));
                if (!(prof ? e->graph_prof_exec[last_slot] : e->graph_exec[last_slot])) {    // (first compaction: the OPEN pointers were swapped)
                    DX_TRY(build_step_graphs(e, prof));
                    if (prof) e->prof_graph_level = e->profiling;
                }
                DX_CUDA(cudaGraphLaunch(prof ? e->graph_prof_exec[last_slot] : e->graph_exec[last_slot], e->stream));
            }
            DX_TRY(sync_ctl(e));                     // synchronises the stream
            const int ran = burst == 1 ? 1 : (int)(e->h_ctl->itr - itr0);
            e->launches += (long long)ran * e->graph_launches;
            e->parity ^= (ran & 1);
            done += ran;
            if (prof) stage_collect_graph(e, last_slot);
            if (ran < burst) break;                  // (everything finished inside the burst)
            if (e->graph_cut != e->sort.key_cut) {   // (the sort asked for every key bit: capture the iteration again)
                drop_graphs(e);
                DX_TRY(build_step_graphs(e, prof));
                if (prof) e->prof_graph_level = e->profiling;
            }
        } else {
            DX_TRY(young_before_step(e));
            DX_TRY(full_step(e));
            ++done;
            DX_TRY(sync_ctl(e));
        }
    }
    if (e->h_ctl->error & DX_DEVERR_DONE) {         // internal to this loop: no other entry point ever sees it
        k_clear_done<<<1, 1, 0, e->stream>>>(e->d_ctl);
        e->h_ctl->error &= ~DX_DEVERR_DONE;
    }
    cudaEventRecord(e->run_ev[1], e->stream);
    DX_CUDA(cudaStreamSynchronize(e->stream));
    cudaEventElapsedTime(&e->last_run_ms, e->run_ev[0], e->run_ev[1]);
    stage_collect(e);
    DX_CUDA(cudaGetLastError());
    if (iters_done) *iters_done = done;
    return DX_OK;
}

extern "C" int dx_step_begin(dx_engine* e, int64_t* n_pending) {
    if (!e) DX_FAIL(DX_ERR_ARG, "null engine");
    if (e->cfg.heur_mode != DX_HEUR_HOST) DX_FAIL(DX_ERR_STATE, "dx_step_begin is for HOST heuristic mode");
    if (e->pending) DX_FAIL(DX_ERR_STATE, "step already pending");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_TRY(young_before_step(e));
    if (e->is_q) DX_TRY(step_q_a(e)); else DX_TRY(step_v_a(e));
    DX_TRY(sync_ctl(e));
    e->pending = true;
    if (n_pending) *n_pending = e->h_ctl->nn_n;
    return DX_OK;
}

extern "C" int dx_get_pending(dx_engine* e, uint8_t* states, uint8_t* goals, int64_t cap) {
    if (!e || !e->pending) DX_FAIL(DX_ERR_STATE, "no step pending");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    const uint32_t n = e->h_ctl->nn_n, start = e->h_ctl->nn_row_start;
    if ((int64_t)n > cap) DX_FAIL(DX_ERR_ARG, "buffer too small");
    if (n == 0) return DX_OK;
    const int SP = e->SP, S = e->S;
    std::vector<uint8_t> rows((size_t)n * SP), grows((size_t)e->n_inst * SP);
    const uint8_t* src = e->eval_all ? e->child_state : e->node_state;
    DX_CUDA(cudaMemcpyAsync(rows.data(), src + (size_t)start * SP, rows.size(), cudaMemcpyDeviceToHost, e->stream));
    DX_CUDA(cudaMemcpyAsync(grows.data(), e->ia.goals, grows.size(), cudaMemcpyDeviceToHost, e->stream));
    DX_CUDA(cudaStreamSynchronize(e->stream));
    for (uint32_t r = 0; r < n; ++r) {
        if (states) memcpy(states + (size_t)r * S, &rows[(size_t)r * SP], S);
        if (goals) {
            uint32_t inst;
            memcpy(&inst, &rows[(size_t)r * SP + SP - 4], 4);
            memcpy(goals + (size_t)r * S, &grows[(size_t)inst * SP], S);
        }
    }
    return DX_OK;
}

extern "C" int dx_step_end(dx_engine* e, const float* vals, int64_t n) {
    if (!e || !e->pending) DX_FAIL(DX_ERR_STATE, "no step pending");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    const uint32_t want = e->h_ctl->nn_n;
    if ((int64_t)want != n) DX_FAIL(DX_ERR_ARG, "dx_step_end: value count does not match pending states");
    if (n > 0) {
        if (!vals) DX_FAIL(DX_ERR_ARG, "null values");
        const size_t cnt = (size_t)n * (e->is_q ? e->A : 1);
        DX_CUDA(cudaMemcpyAsync(e->d_vals, vals, cnt * 4, cudaMemcpyHostToDevice, e->stream));
        k_set_host_vals<<<std::min(dx_div_up(n, 256), 148 * 4), 256, 0, e->stream>>>(e->d_ctl, e->d_vals,
                                                                                    e->eval_all ? e->child_h : e->node_h,
                                                                                    e->node_q, e->A, e->is_q ? 1 : 0);
        e->launches += 1;
    }
    if (e->is_q) DX_TRY(step_q_b(e)); else DX_TRY(step_v_b(e));
    e->pending = false;
    DX_TRY(sync_ctl(e));
    stage_collect(e);
    return DX_OK;
}

// ---------------------------------------------------------------------------------------------
extern "C" int dx_num_instances(dx_engine* e) { return e ? e->n_inst : 0; }

extern "C" int dx_all_finished(dx_engine* e, int32_t* out) {
    if (!e || !out) DX_FAIL(DX_ERR_ARG, "null argument");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_TRY(sync_ctl(e));
    *out = e->h_ctl->n_unfinished == 0 ? 1 : 0;
    return DX_OK;
}

extern "C" int dx_get_status(dx_engine* e, dx_inst_status* out, int32_t cap) {
    if (!e || !out) DX_FAIL(DX_ERR_ARG, "null argument");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    const int n = std::min(cap, e->n_inst);
    if (n <= 0) return DX_OK;
    std::vector<double> lb(n);
    std::vector<unsigned long long> ub(n), gen(n);
    std::vector<uint32_t> goal(n), itr(n), nopen(n), off(e->n_inst + 1);
    std::vector<uint8_t> act(n);
    DX_CUDA(cudaStreamSynchronize(e->stream));
    DX_CUDA(cudaMemcpy(lb.data(), e->ia.lb, n * 8, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(ub.data(), e->ia.ub_key, n * 8, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(gen.data(), e->ia.gen, n * 8, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(goal.data(), e->ia.goal_node, n * 4, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(itr.data(), e->ia.itr, n * 4, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(nopen.data(), e->ia.n_open, n * 4, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(act.data(), e->ia.active, n, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(off.data(), e->pop_off[e->parity], (e->n_inst + 1) * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        out[i].lb = lb[i];
        if (ub[i] == DX_EMPTY64) out[i].ub = INFINITY;
        else { uint32_t b = (uint32_t)(ub[i] >> 32); float f; memcpy(&f, &b, 4); out[i].ub = (double)f; }
        out[i].num_nodes_generated = (int64_t)gen[i];
        out[i].open_size = nopen[i];
        out[i].goal_node = goal[i] == DX_NIL ? -1 : (int64_t)goal[i];
        out[i].itr = (int32_t)itr[i];
        out[i].finished = act[i] ? 0 : 1;
        out[i].num_curr = (int32_t)(off[i + 1] - off[i]);
        out[i].reserved = 0;
    }
    return DX_OK;
}

// Bellman-backup log (evaluate-all engines): one record per node expanded since the last drain, in expansion order
// (iteration, then instance, then pop order -- the order of the `popped` lists PathFind.step() returns,
// base/pathfinding.py:349,399) -> the (state, target) pairs of UpdateHeurVRLKeepGoalABC (updaters/updater_rl.py:143-158).
extern "C" int dx_drain_backups(dx_engine* e, int32_t mode, uint8_t* states, int32_t* actions, double* backups, int32_t* insts,
                                int64_t cap, int64_t* n_out) {
    if (!e || !n_out) DX_FAIL(DX_ERR_ARG, "null argument");
    if (!e->eval_all && !e->backups_q) DX_FAIL(DX_ERR_STATE, "engine was created with max_backups == 0");
    if (mode != DX_BACKUP_BELLMAN && mode != DX_BACKUP_UB_SOLNS && mode != DX_BACKUP_LHBL) DX_FAIL(DX_ERR_ARG, "unknown backup mode");
    if (e->pending) DX_FAIL(DX_ERR_STATE, "step pending");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_TRY(sync_ctl(e));
    const size_t n = e->h_ctl->n_backups;
    const int SP = e->SP, S = e->S;
    cudaStream_t s = e->stream;
    if (n && mode != DX_BACKUP_BELLMAN) {           // drain-time passes over the whole log (labels are final only now)
        uint32_t it_first = 0, it_last = 0;         // the log is in iteration order
        if (mode == DX_BACKUP_LHBL) {
            DX_CUDA(cudaMemcpyAsync(&it_first, e->bk.itr, 4, cudaMemcpyDeviceToHost, s));
            DX_CUDA(cudaMemcpyAsync(&it_last, e->bk.itr + (n - 1), 4, cudaMemcpyDeviceToHost, s));
            DX_CUDA(cudaStreamSynchronize(s));
        }
        if (e->backups_q)
            e->launches += dx_launch_backup_q_modes(s, e->d_ctl, e->node_parent, e->node_solved, e->node_h, e->bk, mode, it_first, it_last);
        else if (mode == DX_BACKUP_UB_SOLNS)
            e->launches += dx_launch_backup_ub_paths(s, e->d_ctl, e->node_parent, e->bk);
        else
            for (uint32_t it = it_last + 1; it-- > it_first;) e->launches += dx_launch_backup_tree(s, e->dom, e->d_ctl, e->bk, it);
    }
    if ((int64_t)n <= cap && states && backups && insts && (actions || !e->backups_q)) {
        // direct path: the copies land in the caller's buffers (rows repacked SP -> S on the device first); the records of nodes
        // popped by finished instances (NaN, A* only, rare) are squeezed out in place afterwards
        std::vector<uint8_t> act(e->backups_q ? n : 0);
        if (n) {
            e->launches += dx_launch_backup_pack_rows(s, e->dom, e->d_ctl, e->bk);
            DX_CUDA(cudaMemcpyAsync(states, e->bk.pack, n * S, cudaMemcpyDeviceToHost, s));
            DX_CUDA(cudaMemcpyAsync(backups, e->bk.val, n * sizeof(double), cudaMemcpyDeviceToHost, s));
            DX_CUDA(cudaMemcpyAsync(insts, e->bk.inst, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            if (e->backups_q) DX_CUDA(cudaMemcpyAsync(act.data(), e->bk.act, n, cudaMemcpyDeviceToHost, s));
            DX_CUDA(cudaStreamSynchronize(s));
        }
        size_t k = 0;
        for (size_t r = 0; r < n; ++r) {
            if (backups[r] != backups[r]) continue;
            if (k != r) {
                memmove(states + k * S, states + r * S, S);
                backups[k] = backups[r];
                insts[k] = insts[r];
            }
            if (actions) actions[k] = e->backups_q ? (int32_t)act[r] : -1;
            ++k;
        }
        *n_out = (int64_t)k;
    } else {        // staging path: NULL outputs or cap below the raw slot count
        std::vector<uint8_t> rows(n * SP);
        std::vector<double> vals(n);
        std::vector<uint32_t> inst(n);
        std::vector<uint8_t> act(n);
        if (n) {
            if (e->backups_q) DX_CUDA(cudaMemcpyAsync(act.data(), e->bk.act, n, cudaMemcpyDeviceToHost, s));
            DX_CUDA(cudaMemcpyAsync(rows.data(), e->bk.state, rows.size(), cudaMemcpyDeviceToHost, s));
            DX_CUDA(cudaMemcpyAsync(vals.data(), e->bk.val, n * sizeof(double), cudaMemcpyDeviceToHost, s));
            DX_CUDA(cudaMemcpyAsync(inst.data(), e->bk.inst, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
            DX_CUDA(cudaStreamSynchronize(s));
        }
        int64_t m = 0;
        for (size_t r = 0; r < n; ++r) m += vals[r] == vals[r] ? 1 : 0;        // NaN: popped by an instance that had finished
        *n_out = m;
        if (m > cap) DX_FAIL(DX_ERR_ARG, "buffer too small (n_out holds the record count; the log was not drained)");
        int64_t k = 0;
        for (size_t r = 0; r < n; ++r) {
            if (vals[r] != vals[r]) continue;
            if (states) memcpy(states + (size_t)k * S, &rows[r * SP], S);
            if (backups) backups[k] = vals[r];
            if (insts) insts[k] = (int32_t)inst[r];
            if (actions) actions[k] = e->backups_q ? (int32_t)act[r] : -1;
            ++k;
        }
    }
    if (e->eval_all) e->launches += dx_launch_backup_clear_map(s, e->d_ctl, e->bk);
    k_reset_backups<<<1, 1, 0, s>>>(e->d_ctl);
    e->launches += 1;
    DX_CUDA(cudaStreamSynchronize(s));
    return DX_OK;
}

// UpdateHER._get_her_goals (deepxube/base/updater.py:485-530): per instance, the state of the deepest node that has children among
// the records currently in the backup log (the root's state when nothing deeper was expanded) and its depth.  Call before the drain.
extern "C" int dx_backup_deepest(dx_engine* e, uint8_t* states, int32_t* depths, int32_t cap) {
    if (!e || !states) DX_FAIL(DX_ERR_ARG, "null argument");
    if (!e->eval_all && !e->backups_q) DX_FAIL(DX_ERR_STATE, "engine was created with max_backups == 0");
    if (e->pending) DX_FAIL(DX_ERR_STATE, "step pending");
    if (cap < e->n_inst) DX_FAIL(DX_ERR_ARG, "buffer holds fewer rows than there are instances");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    if (e->n_inst == 0) return DX_OK;
    cudaStream_t s = e->stream;
    e->launches += dx_launch_her_deepest(s, e->dom, e->d_ctl, e->bk, e->node_parent, e->node_state, e->backups_q ? 1 : 0,
                                         e->ia.root, e->her_cand, e->her_depth, e->her_best, e->her_node, e->her_rows, e->n_inst);
    DX_CUDA(cudaMemcpyAsync(states, e->her_rows, (size_t)e->n_inst * e->S, cudaMemcpyDeviceToHost, s));
    if (depths) DX_CUDA(cudaMemcpyAsync(depths, e->her_best, (size_t)e->n_inst * 4, cudaMemcpyDeviceToHost, s));
    DX_CUDA(cudaStreamSynchronize(s));
    DX_CUDA(cudaGetLastError());
    return DX_OK;
}

extern "C" int dx_get_counters(dx_engine* e, dx_counters* out) {
    if (!e || !out) DX_FAIL(DX_ERR_ARG, "null argument");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_TRY(sync_ctl(e));
    memset(out, 0, sizeof(*out));
    out->nodes_stored = e->h_ctl->node_count;
    out->heur_evals = (int64_t)e->h_ctl->heur_evals;
    out->children_generated = (int64_t)e->h_ctl->gen_total;
    out->iterations = e->h_ctl->itr;
    out->kernel_launches = e->launches;
    out->backups_logged = e->h_ctl->n_backups;
    out->hidden_gemm_launches = e->weights_loaded ? 2 * e->net.num_blocks - (e->net.tc_l1 ? 1 : 0) : 0;
    return DX_OK;
}

extern "C" int dx_get_path(dx_engine* e, int32_t inst, int32_t* actions, uint8_t* states_on_path, int32_t cap, int32_t* len) {
    if (!e || !len || inst < 0 || inst >= e->n_inst) DX_FAIL(DX_ERR_ARG, "bad argument");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    uint32_t goal;
    DX_CUDA(cudaStreamSynchronize(e->stream));
    DX_CUDA(cudaMemcpy(&goal, e->ia.goal_node + inst, 4, cudaMemcpyDeviceToHost));
    if (goal == DX_NIL) { *len = -1; return DX_OK; }
    const size_t need = (size_t)cap * 4 + (size_t)(cap + 1) * 4 + 16 + (size_t)(cap + 1) * e->SP;
    DX_TRY(ensure_stage(e, need));
    int32_t* d_act = (int32_t*)e->d_stage;
    uint32_t* d_ids = (uint32_t*)(e->d_stage + (size_t)cap * 4);
    int* d_len = (int*)(e->d_stage + (size_t)cap * 4 + (size_t)(cap + 1) * 4);
    uint8_t* d_rows = e->d_stage + (((size_t)cap * 8 + 4 + 16 + 15) / 16) * 16;
    k_path_walk<<<1, 1, 0, e->stream>>>(e->node_parent, e->node_action, goal, d_act, d_ids, cap, d_len);
    e->launches += 1;
    int hlen = 0;
    DX_CUDA(cudaMemcpyAsync(&hlen, d_len, 4, cudaMemcpyDeviceToHost, e->stream));
    DX_CUDA(cudaStreamSynchronize(e->stream));
    *len = hlen;
    if (hlen > cap) DX_FAIL(DX_ERR_ARG, "path longer than cap");
    std::vector<int32_t> a(std::max(hlen, 1));
    if (hlen && actions) {
        DX_CUDA(cudaMemcpy(a.data(), d_act, (size_t)hlen * 4, cudaMemcpyDeviceToHost));
        for (int i = 0; i < hlen; ++i) actions[i] = a[hlen - 1 - i];
    }
    if (states_on_path) {
        k_gather_rows<<<dx_div_up((long long)(hlen + 1) * (e->SP / 16), 128), 128, 0, e->stream>>>(e->node_state, d_ids, hlen + 1, e->SP, d_rows);
        e->launches += 1;
        std::vector<uint8_t> rows((size_t)(hlen + 1) * e->SP);
        DX_CUDA(cudaMemcpyAsync(rows.data(), d_rows, rows.size(), cudaMemcpyDeviceToHost, e->stream));
        DX_CUDA(cudaStreamSynchronize(e->stream));
        for (int i = 0; i <= hlen; ++i) memcpy(states_on_path + (size_t)i * e->S, &rows[(size_t)(hlen - i) * e->SP], e->S);
    }
    return DX_OK;
}

extern "C" int dx_get_curr_nodes(dx_engine* e, int32_t inst, uint8_t* states, float* g, float* h, int32_t cap, int32_t* n) {
    if (!e || !n || inst < 0 || inst >= e->n_inst) DX_FAIL(DX_ERR_ARG, "bad argument");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_CUDA(cudaStreamSynchronize(e->stream));
    uint32_t off[2];
    DX_CUDA(cudaMemcpy(off, e->pop_off[e->parity] + inst, 8, cudaMemcpyDeviceToHost));
    const int cnt = (int)(off[1] - off[0]);
    *n = cnt;
    if (cnt > cap) DX_FAIL(DX_ERR_ARG, "buffer too small");
    if (cnt == 0) return DX_OK;
    const size_t need = (size_t)cnt * e->SP + (size_t)cnt * 8 + 64;
    DX_TRY(ensure_stage(e, need));
    uint8_t* d_rows = e->d_stage;
    float* d_g = (float*)(e->d_stage + (((size_t)cnt * e->SP + 15) / 16) * 16);
    float* d_h = d_g + cnt;
    const uint32_t* ids = e->popped[e->parity] + off[0];
    k_gather_rows<<<dx_div_up((long long)cnt * (e->SP / 16), 128), 128, 0, e->stream>>>(e->node_state, ids, cnt, e->SP, d_rows);
    k_gather_f32<<<dx_div_up(cnt, 128), 128, 0, e->stream>>>(e->node_g, ids, cnt, d_g);
    k_gather_f32<<<dx_div_up(cnt, 128), 128, 0, e->stream>>>(e->node_h, ids, cnt, d_h);
    e->launches += 3;
    std::vector<uint8_t> rows((size_t)cnt * e->SP);
    DX_CUDA(cudaMemcpyAsync(rows.data(), d_rows, rows.size(), cudaMemcpyDeviceToHost, e->stream));
    if (g) DX_CUDA(cudaMemcpyAsync(g, d_g, (size_t)cnt * 4, cudaMemcpyDeviceToHost, e->stream));
    if (h) DX_CUDA(cudaMemcpyAsync(h, d_h, (size_t)cnt * 4, cudaMemcpyDeviceToHost, e->stream));
    DX_CUDA(cudaStreamSynchronize(e->stream));
    if (states)
        for (int i = 0; i < cnt; ++i) memcpy(states + (size_t)i * e->S, &rows[(size_t)i * e->SP], e->S);
    return DX_OK;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone domain API on host buffers.
static int domain_call(int device, int domain, int dim, const uint8_t* states, const uint8_t* goals, const int32_t* actions,
                       long long n, uint8_t* out, int what) {   // what: 0 expand, 1 next_state, 2 is_solved
    if (n < 0 || (n > 0 && (!states || !out))) DX_FAIL(DX_ERR_ARG, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) DX_FAIL(DX_ERR_CUDA, "no CUDA device: libdxb200 has no CPU fallback");
    DX_CUDA(cudaSetDevice(device));
    DomainDev d;
    std::vector<uint8_t> table;
    DX_TRY(make_domain_host(domain, dim, &d, &table));
    if (n == 0) return DX_OK;
    const int SP = d.SP, S = d.S;
    const long long n_out = what == 0 ? n * d.A : n;
    std::vector<uint8_t> rows((size_t)n * SP, 0);
    for (long long i = 0; i < n; ++i) memcpy(&rows[(size_t)i * SP], states + (size_t)i * S, S);
    uint8_t *d_rows = nullptr, *d_out = nullptr, *d_table = nullptr, *d_goals = nullptr;
    int32_t* d_act = nullptr;
    int rc = DX_OK;
    auto cleanup = [&]() { cudaFree(d_rows); cudaFree(d_out); cudaFree(d_table); cudaFree(d_goals); cudaFree(d_act); };
#define C_(x) do { if ((x) != cudaSuccess) { dx_set_error(std::string(#x) + ": " + cudaGetErrorString(cudaGetLastError())); cleanup(); return DX_ERR_CUDA; } } while (0)
    C_(cudaMalloc(&d_rows, rows.size()));
    C_(cudaMalloc(&d_table, std::max<size_t>(table.size(), 16)));
    C_(cudaMemcpy(d_rows, rows.data(), rows.size(), cudaMemcpyHostToDevice));
    if (!table.empty()) C_(cudaMemcpy(d_table, table.data(), table.size(), cudaMemcpyHostToDevice));
    d.table = d_table;
    if (what == 2) {
        std::vector<uint8_t> grows((size_t)n * SP, 0);
        for (long long i = 0; i < n; ++i) memcpy(&grows[(size_t)i * SP], goals + (size_t)i * S, S);
        C_(cudaMalloc(&d_goals, grows.size()));
        C_(cudaMemcpy(d_goals, grows.data(), grows.size(), cudaMemcpyHostToDevice));
        C_(cudaMalloc(&d_out, (size_t)n));
        dx_launch_standalone_solved(0, d, d_rows, d_goals, n, d_out);
        C_(cudaMemcpy(out, d_out, (size_t)n, cudaMemcpyDeviceToHost));
    } else {
        if (what == 1) {
            if (!actions) { cleanup(); DX_FAIL(DX_ERR_ARG, "null actions"); }
            for (long long i = 0; i < n; ++i)
                if (actions[i] < 0 || actions[i] >= d.A) { cleanup(); DX_FAIL(DX_ERR_ARG, "action out of range"); }
            C_(cudaMalloc(&d_act, (size_t)n * 4));
            C_(cudaMemcpy(d_act, actions, (size_t)n * 4, cudaMemcpyHostToDevice));
        }
        C_(cudaMalloc(&d_out, (size_t)n_out * SP));
        dx_launch_standalone_expand(0, d, d_rows, n, what == 1 ? d_act : nullptr, d_out);
        std::vector<uint8_t> orows((size_t)n_out * SP);
        C_(cudaMemcpy(orows.data(), d_out, orows.size(), cudaMemcpyDeviceToHost));
        for (long long i = 0; i < n_out; ++i) memcpy(out + (size_t)i * S, &orows[(size_t)i * SP], S);
    }
    C_(cudaGetLastError());
#undef C_
    cleanup();
    return rc;
}

extern "C" int dx_expand(int32_t device, int32_t domain, int32_t dim, const uint8_t* states, int64_t n, uint8_t* children) {
    return domain_call(device, domain, dim, states, nullptr, nullptr, n, children, 0);
}
extern "C" int dx_next_state(int32_t device, int32_t domain, int32_t dim, const uint8_t* states, const int32_t* actions,
                             int64_t n, uint8_t* next) {
    return domain_call(device, domain, dim, states, nullptr, actions, n, next, 1);
}
extern "C" int dx_is_solved(int32_t device, int32_t domain, int32_t dim, const uint8_t* states, const uint8_t* goals, int64_t n,
                            uint8_t* solved) {
    if (n > 0 && !goals) DX_FAIL(DX_ERR_ARG, "null goals");
    return domain_call(device, domain, dim, states, goals, nullptr, n, solved, 2);
}

// ---------------------------------------------------------------------------------------------
extern "C" int dx_load_weights(dx_engine* e, const dx_net_desc* desc, const float* blob, int64_t n_floats) {
    if (!e || !desc || !blob) DX_FAIL(DX_ERR_ARG, "null argument");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    if (desc->in_count != e->dom.in_count || desc->one_hot_depth != e->dom.depth)
        DX_FAIL(DX_ERR_ARG, "network input shape does not match the domain's NN input");
    const int want_out = (e->is_q && desc->act_depth == 0) ? e->A : 1;       // an action-as-input Q network has one output
    if (desc->out_dim != want_out)
        DX_FAIL(DX_ERR_ARG, "network out_dim must be 1 for graph_v / action-as-input networks and num_actions for graph_q");
    DX_TRY(dx_upload_net(e, desc, blob, n_floats));
    e->weights_loaded = true;
    return DX_OK;
}

// Evaluate the loaded network on host-provided states: rows are staged in a scratch area with a private
// goal table (row r uses goal r), so it does not disturb the search state.
// min over actions of (1 + value) per row, zeroed where solved -- the reduction of dx_vi_targets
__global__ void k_vi_reduce(const float* __restrict__ vals, const uint8_t* __restrict__ solved, long long n, int A,
                            double* __restrict__ out) {
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        double v = __longlong_as_double(0x7ff0000000000000ll);
        for (int a = 0; a < A; ++a) v = fmin(v, 1.0 + (double)vals[r * A + a]);
        out[r] = (solved && solved[r]) ? 0.0 : v;
    }
}

// Value-iteration / Q-learning targets of replay-buffer samples with the loaded (target) network:
//   graph_v engines (UpdateHeurVRL._get_labels_rb, updaters/updater_rl.py:41-69):  states are expanded,
//       target = min_a (1.0 + h(child_a)) * (not solved);   solved == NULL: is_solved(state, goal) is computed
//   graph_q engines (UpdateHeurQRL._get_labels_rb, :102-120):  `states` are the NEXT states of the sampled edges,
//       target = (1.0 + min_a q(next, a)) * (not solved);   solved = is_solved of the edge's node (NULL: none solved)
// float64, like the reference's numpy arithmetic.
extern "C" int dx_vi_targets(dx_engine* e, const uint8_t* states, const uint8_t* goals, const uint8_t* solved, int64_t n,
                             double* out) {
    if (!e || n < 0 || (n > 0 && (!states || !out))) DX_FAIL(DX_ERR_ARG, "bad argument");
    if (!e->weights_loaded) DX_FAIL(DX_ERR_STATE, "no weights loaded");
    if (e->pending) DX_FAIL(DX_ERR_STATE, "step pending");
    if (e->is_beam) DX_FAIL(DX_ERR_UNSUPPORTED, "dx_vi_targets: graph_v / graph_q engines");
    DX_TRY(check_rows_in_range(e, states, n));
    DX_CUDA(cudaSetDevice(e->cfg.device));
    if (n == 0) return DX_OK;
    const int amul = e->net.act_depth > 0 ? e->net.act_depth : 1;       // action-as-input Q network: one row per (state, action)
    const int SP = e->SP, S = e->S, A = e->A;
    const int fan = e->is_q ? 1 : A;                                     // network rows per input state
    const int od = e->is_q ? e->net.out_dim * amul : 1;
    const long long chunk = std::max<long long>(1, std::min<long long>(e->mlp.cap / ((long long)fan * amul), 1 << 15));
    const size_t off_goals = (size_t)chunk * SP, off_child = off_goals + (size_t)chunk * SP,
                 off_vals = off_child + (size_t)chunk * fan * SP, off_h = off_vals + (size_t)chunk * fan * od * 4,
                 off_out = (off_h + (size_t)chunk * fan * 4 + 7) / 8 * 8, off_solved = off_out + (size_t)chunk * 8;
    DX_TRY(ensure_stage(e, off_solved + (size_t)chunk + 64));
    uint8_t* d_rows = e->d_stage;
    uint8_t* d_goals = e->d_stage + off_goals;
    uint8_t* d_child = e->d_stage + off_child;
    float* d_vals = (float*)(e->d_stage + off_vals);                     // Q: the q-values
    float* d_h = (float*)(e->d_stage + off_h);                           // V: h(child); Q: min_a q
    double* d_out = (double*)(e->d_stage + off_out);
    uint8_t* d_solved = e->d_stage + off_solved;
    std::vector<uint8_t> rows((size_t)chunk * SP), grows((size_t)chunk * SP);
    DX_TRY(sync_ctl(e));
    const Ctl saved = *e->h_ctl;
    cudaStream_t st = e->stream;
    for (long long s0 = 0; s0 < n; s0 += chunk) {
        const long long m = std::min(chunk, n - s0);
        std::fill(rows.begin(), rows.end(), 0);
        std::fill(grows.begin(), grows.end(), 0);
        for (long long r = 0; r < m; ++r) {
            memcpy(&rows[(size_t)r * SP], states + (size_t)(s0 + r) * S, S);
            const uint32_t gi = (uint32_t)r;                             // the row's goal index, inherited by its children
            memcpy(&rows[(size_t)r * SP + SP - 4], &gi, 4);
            if (goals) memcpy(&grows[(size_t)r * SP], goals + (size_t)(s0 + r) * S, S);
        }
        DX_CUDA(cudaMemcpyAsync(d_rows, rows.data(), (size_t)m * SP, cudaMemcpyHostToDevice, st));
        DX_CUDA(cudaMemcpyAsync(d_goals, grows.data(), (size_t)m * SP, cudaMemcpyHostToDevice, st));
        const uint8_t* sol = nullptr;
        if (solved) {
            DX_CUDA(cudaMemcpyAsync(d_solved, solved + s0, (size_t)m, cudaMemcpyHostToDevice, st));
            sol = d_solved;
        } else if (!e->is_q) {
            e->launches += dx_launch_standalone_solved(st, e->dom, d_rows, d_goals, m, d_solved);
            sol = d_solved;
        }
        const uint8_t* nn_rows = d_rows;
        if (!e->is_q) {
            e->launches += dx_launch_standalone_expand(st, e->dom, d_rows, m, nullptr, d_child);
            nn_rows = d_child;
        }
        k_set_job<<<1, 1, 0, st>>>(e->d_ctl, 0, (uint32_t)(m * fan));
        e->launches += 1;
        float* out_q = e->is_q ? d_vals : nullptr;
        e->launches += launch_net(e, nn_rows, d_goals, d_h, out_q, 0, m * fan);
        k_vi_reduce<<<std::min(dx_div_up(m, 256), 148 * 4), 256, 0, st>>>(d_h, sol, m, fan, d_out);
        e->launches += 1;
        DX_CUDA(cudaMemcpyAsync(out + s0, d_out, (size_t)m * 8, cudaMemcpyDeviceToHost, st));
        DX_CUDA(cudaStreamSynchronize(st));
    }
    k_set_job<<<1, 1, 0, st>>>(e->d_ctl, saved.nn_row_start, saved.nn_n);
    e->launches += 1;
    DX_CUDA(cudaStreamSynchronize(st));
    DX_CUDA(cudaGetLastError());
    return DX_OK;
}

extern "C" int dx_heur_eval(dx_engine* e, const uint8_t* states, const uint8_t* goals, int64_t n, float* out) {
    if (!e || n < 0 || (n > 0 && (!states || !out))) DX_FAIL(DX_ERR_ARG, "bad argument");
    if (!e->weights_loaded) DX_FAIL(DX_ERR_STATE, "no weights loaded");
    if (e->pending) DX_FAIL(DX_ERR_STATE, "step pending");
    DX_TRY(check_rows_in_range(e, states, n));
    DX_CUDA(cudaSetDevice(e->cfg.device));
    if (n == 0) return DX_OK;
    const int amul = e->net.act_depth > 0 ? e->net.act_depth : 1;
    const int SP = e->SP, S = e->S, od = e->net.out_dim * amul;        // action-as-input: one value per (state, action)
    const long long chunk = std::min<long long>(e->mlp.cap / amul, 1 << 16);
    DX_TRY(ensure_stage(e, (size_t)chunk * SP * 2 + (size_t)chunk * od * 4 + 64));
    uint8_t* d_rows = e->d_stage;
    uint8_t* d_goals = e->d_stage + (size_t)chunk * SP;
    float* d_out = (float*)(e->d_stage + (size_t)chunk * SP * 2);
    std::vector<uint8_t> rows((size_t)chunk * SP), grows((size_t)chunk * SP);
    DX_TRY(sync_ctl(e));
    const Ctl saved = *e->h_ctl;
    for (long long s0 = 0; s0 < n; s0 += chunk) {
        const long long m = std::min(chunk, n - s0);
        std::fill(rows.begin(), rows.end(), 0);
        std::fill(grows.begin(), grows.end(), 0);
        for (long long r = 0; r < m; ++r) {
            memcpy(&rows[(size_t)r * SP], states + (size_t)(s0 + r) * S, S);
            const uint32_t gi = (uint32_t)r;
            memcpy(&rows[(size_t)r * SP + SP - 4], &gi, 4);
            if (goals) memcpy(&grows[(size_t)r * SP], goals + (size_t)(s0 + r) * S, S);
        }
        DX_CUDA(cudaMemcpyAsync(d_rows, rows.data(), (size_t)m * SP, cudaMemcpyHostToDevice, e->stream));
        DX_CUDA(cudaMemcpyAsync(d_goals, grows.data(), (size_t)m * SP, cudaMemcpyHostToDevice, e->stream));
        k_set_job<<<1, 1, 0, e->stream>>>(e->d_ctl, 0, (uint32_t)m);
        e->launches += 1;
        e->launches += launch_net(e, d_rows, d_goals, nullptr, d_out, 0, m);
        DX_CUDA(cudaMemcpyAsync(out + (size_t)s0 * od, d_out, (size_t)m * od * 4, cudaMemcpyDeviceToHost, e->stream));
        DX_CUDA(cudaStreamSynchronize(e->stream));
    }
    // restore the job fields / counters the evaluation touched
    k_set_job<<<1, 1, 0, e->stream>>>(e->d_ctl, saved.nn_row_start, saved.nn_n);
    e->launches += 1;
    DX_CUDA(cudaStreamSynchronize(e->stream));
    DX_CUDA(cudaGetLastError());
    return DX_OK;
}

extern "C" int dx_row_bytes(dx_engine* e, int32_t* row_bytes) {
    if (!e || !row_bytes) DX_FAIL(DX_ERR_ARG, "null argument");
    *row_bytes = e->SP;
    return DX_OK;
}

// Times `reps` network evaluations over n node-store rows already resident on the device
// (rows [0, n) of the node store; whatever states are there -- the arithmetic does not depend on them).
extern "C" int dx_bench_heur(dx_engine* e, int64_t n, int32_t reps, float* ms_total) {
    if (!e || !ms_total || n < 1 || reps < 1) DX_FAIL(DX_ERR_ARG, "bad argument");
    if (!e->weights_loaded) DX_FAIL(DX_ERR_STATE, "no weights loaded");
    if (n > (int64_t)e->max_nodes || n * (e->net.act_depth > 0 ? e->net.act_depth : 1) > e->mlp.cap)
        DX_FAIL(DX_ERR_ARG, "n exceeds node store / activation capacity");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_TRY(sync_ctl(e));
    const Ctl saved = *e->h_ctl;
    DX_TRY(ensure_stage(e, (size_t)n * e->net.out_dim * (e->net.act_depth > 0 ? e->net.act_depth : 1) * 4 + 64));
    float* d_out = (float*)e->d_stage;
    k_set_job<<<1, 1, 0, e->stream>>>(e->d_ctl, 0, (uint32_t)n);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, e->stream);
    for (int r = 0; r < reps; ++r) {
        e->launches += launch_net(e, e->node_state, e->ia.goals, nullptr, d_out, 0, n);
    }
    cudaEventRecord(b, e->stream);
    k_set_job<<<1, 1, 0, e->stream>>>(e->d_ctl, saved.nn_row_start, saved.nn_n);
    DX_CUDA(cudaStreamSynchronize(e->stream));
    cudaEventElapsedTime(ms_total, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    DX_CUDA(cudaGetLastError());
    return DX_OK;
}

extern "C" int dx_set_profiling(dx_engine* e, int32_t on) {
    if (!e) DX_FAIL(DX_ERR_ARG, "null engine");
    stage_collect(e);
    e->profiling = on < 0 ? 0 : (on > 2 ? 2 : on);
    memset(e->stage_ms, 0, sizeof(e->stage_ms));
    memset(e->stage_calls, 0, sizeof(e->stage_calls));
    return DX_OK;
}

extern "C" int dx_get_stage_ms(dx_engine* e, float* ms, int32_t cap) {
    if (!e || !ms) DX_FAIL(DX_ERR_ARG, "null argument");
    stage_collect(e);
    for (int i = 0; i < std::min<int>(cap, ST_COUNT); ++i) ms[i] = e->stage_ms[i];
    return DX_OK;
}

extern "C" int dx_get_stage_calls(dx_engine* e, int64_t* calls, int32_t cap) {
    if (!e || !calls) DX_FAIL(DX_ERR_ARG, "null argument");
    stage_collect(e);
    for (int i = 0; i < std::min<int>(cap, ST_COUNT); ++i) calls[i] = e->stage_calls[i];
    return DX_OK;
}

extern "C" int dx_last_run_ms(dx_engine* e, float* ms) {
    if (!e || !ms) DX_FAIL(DX_ERR_ARG, "null argument");
    *ms = e->last_run_ms;
    return DX_OK;
}

// ---------------------------------------------------------------------------------------------
// Weight ingest: parse the blob (layout in dxb200.h), fold eval-mode BN into the preceding Linear
// (y = (Wx + b - mu) * gamma / sqrt(var + 1e-5) + beta, deepxube/pytorch/pytorch_models.py:188 ->
//  W' = s * W, b' = s * (b - mu) + beta with s computed in double), pad H to a multiple of 128 with zeros.
__global__ void k_f32_to_bf16(const float* __restrict__ src, uint16_t* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t u = __float_as_uint(src[i]);
        // round to nearest even
        const uint32_t r = u + 0x7FFFu + ((u >> 16) & 1u);
        dst[i] = (uint16_t)(r >> 16);
    }
}

int dx_upload_net(dx_engine* e, const dx_net_desc* desc, const float* blob, int64_t n_floats) {
    // A second load (the heuristic / target network changed between two batches of an RL loop, updaters/updater_rl.py)
    // overwrites the parameters IN PLACE: same shape required, the captured step graphs keep pointing at the same buffers.
    const bool reload = e->weights_loaded;
    DX_CUDA(cudaStreamSynchronize(e->stream));
    const int H = desc->res_dim, nb = desc->num_blocks, od = desc->out_dim, bn = desc->batch_norm;
    const int act_depth = desc->act_depth;
    if (act_depth < 0 || (act_depth > 0 && (od != 1 || !e->is_q || act_depth != e->A)))
        DX_FAIL(DX_ERR_ARG, "action-as-input network: needs a graph_q engine, out_dim 1 and act_depth == number of actions");
    const int in_dim = desc->in_count * desc->one_hot_depth + act_depth;
    if (H < 1 || nb < 0 || od < 1 || od > DX_MAX_OUT || desc->in_count > 256) DX_FAIL(DX_ERR_ARG, "unsupported network shape");
    const int64_t per_layer = (int64_t)H * H + H + (bn ? 4 * H : 0);
    const int64_t want = (int64_t)H * in_dim + H + 2 * nb * per_layer + (int64_t)od * H + od;
    if (n_floats != want) DX_FAIL(DX_ERR_ARG, "weight blob has " + std::to_string(n_floats) + " floats, expected " + std::to_string(want));
    const int Hp = ((H + 127) / 128) * 128;
    const int in_dim_p = ((in_dim + 63) / 64) * 64;
    NetDev& net = e->net;
    if (reload && (net.in_count != desc->in_count || net.depth != desc->one_hot_depth || net.H != H || net.num_blocks != nb ||
                   net.out_dim != od || net.act_depth != act_depth))
        DX_FAIL(DX_ERR_ARG, "dx_load_weights: a reload must have the shape of the network loaded first");
    net.in_count = desc->in_count; net.depth = desc->one_hot_depth; net.H = H; net.Hp = Hp; net.num_blocks = nb;
    net.out_dim = od; net.in_dim = in_dim; net.in_dim_p = in_dim_p; net.act_depth = act_depth;
    std::vector<float> w0t((size_t)in_dim * Hp, 0.f), b0(Hp, 0.f), w0p((size_t)Hp * in_dim_p, 0.f);
    std::vector<float> wres((size_t)2 * nb * Hp * Hp, 0.f), bres((size_t)2 * nb * Hp, 0.f), wout((size_t)od * Hp, 0.f), bout(od, 0.f);
    const float* p = blob;
    for (int o = 0; o < H; ++o)
        for (int i = 0; i < in_dim; ++i) {
            w0t[(size_t)i * Hp + o] = p[(size_t)o * in_dim + i];
            w0p[(size_t)o * in_dim_p + i] = p[(size_t)o * in_dim + i];
        }
    p += (size_t)H * in_dim;
    for (int o = 0; o < H; ++o) b0[o] = p[o];
    p += H;
    for (int l = 0; l < 2 * nb; ++l) {
        const float* W = p; p += (size_t)H * H;
        const float* b = p; p += H;
        const float *gam = nullptr, *bet = nullptr, *mu = nullptr, *var = nullptr;
        if (bn) { gam = p; bet = p + H; mu = p + 2 * H; var = p + 3 * H; p += 4 * H; }
        float* Wd = &wres[(size_t)l * Hp * Hp];
        float* bd = &bres[(size_t)l * Hp];
        for (int o = 0; o < H; ++o) {
            double s = 1.0, sh = 0.0, mo = 0.0;
            if (bn) { s = (double)gam[o] / sqrt((double)var[o] + 1e-5); sh = bet[o]; mo = mu[o]; }
            for (int i = 0; i < H; ++i) Wd[(size_t)o * Hp + i] = (float)((double)W[(size_t)o * H + i] * s);
            bd[o] = (float)(((double)b[o] - mo) * s + sh);
        }
    }
    for (int o = 0; o < od; ++o)
        for (int i = 0; i < H; ++i) wout[(size_t)o * Hp + i] = p[(size_t)o * H + i];
    p += (size_t)od * H;
    for (int o = 0; o < od; ++o) bout[o] = p[o];

    // rows one evaluation can hold: popped nodes (Q*) or children (A*); x actions for an action-as-input network
    const long long cap = std::max<long long>(e->is_q ? e->max_pop : e->max_children, e->max_inst) * (act_depth > 0 ? act_depth : 1);
#define UP_(dst, vec) do { if (!(dst)) DX_TRY(e->alloc(&(dst), (vec).size())); \
        DX_CUDA(cudaMemcpy((dst), (vec).data(), (vec).size() * sizeof(float), cudaMemcpyHostToDevice)); } while (0)
    UP_(net.w0t, w0t); UP_(net.b0, b0); UP_(net.wres, wres); UP_(net.bres, bres); UP_(net.wout, wout); UP_(net.bout, bout);
#undef UP_
    // Engines whose evaluations are small (solve-time regime: graph.100B) run the whole network in ONE launch (k_mlp_small.cu).
    // The choice is per ENGINE, not per call, so that every value an engine produces comes from the same arithmetic.
    if (e->cfg.heur_mode == DX_HEUR_NET_BF16 || e->cfg.heur_mode == DX_HEUR_NET_FP32) {
        const bool want = reload ? net.small != nullptr : (cap <= DX_SMALL_NET_ROWS && dx_mlp_small_supported(net, e->dom));
        if (want) DX_TRY(dx_mlp_small_upload(e->stream, net, e->cfg.heur_mode == DX_HEUR_NET_FP32, w0p.data(), wres.data()));
    }
    // bf16 copies for the tensor-core path (first layer K-major [Hp, in_dim_p]; residual layers [l, Hp, Hp]).
    // Bias folding (NetDev::tc_fold): spare padding columns carry the bias through the GEMM itself.
    net.tc_fold = 0;
    if (e->cfg.heur_mode == DX_HEUR_NET_BF16) {
        auto bf16r = [](float x) {                 // round to nearest even, as k_f32_to_bf16
            uint32_t u; memcpy(&u, &x, 4);
            u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;
            float y; memcpy(&y, &u, 4);
            return y;
        };
        auto split3 = [&](float b, float* out3) {
            out3[0] = bf16r(b);
            out3[1] = bf16r(b - out3[0]);
            out3[2] = bf16r(b - out3[0] - out3[1]);
        };
        auto f16r = [](float x) { return __half2float(__float2half_rn(x)); };
        auto split3h = [&](float b, float* out3) {      // the same 3-piece split in fp16 (fused first block with fp16 operands)
            out3[0] = f16r(b);
            out3[1] = f16r(b - out3[0]);
            out3[2] = f16r(b - out3[0] - out3[1]);
        };
        const char* fenv = getenv("DXB200_FOLD_BIAS");            // =0: keep the explicit bias adds (A/B measurements)
        const bool fold_h = Hp - H >= 3 && nb > 0 && !(fenv && fenv[0] == '0');
        const bool fold_in = fold_h && in_dim_p - in_dim >= 3;
        // Fused first block (NetDev::tc_l1): rows [Hp, 2 * Hp) of the input-layer weights hold W1 . W0 (and b10 = W1 . b0 + b1),
        // computed in double from the BN-folded fp32 parameters; needs the generated-input launch.
        const char* l1env = getenv("DXB200_FUSE_L1");
        const bool l1 = reload ? net.tc_l1 != 0 : (nb > 0 && dx_mlp_bf16_gen_input(net, e->dom) && !(l1env && l1env[0] == '0'));
        const char* l1s = getenv("DXB200_L1_SPLIT");
        const bool l1_pairs = l1 && (reload ? net.tc_l1 == 2 : (l1s && l1s[0] == '1'));   // bf16 (hi, lo) weight pairs instead of fp16 weights
        net.tc_l1 = l1 ? (l1_pairs ? 2 : 1) : 0;
        std::vector<float> b10(Hp, 0.f);
        if (l1) {
            w0p.resize((size_t)2 * Hp * in_dim_p, 0.f);
            std::vector<double> acc(in_dim);
            const float* W1 = &wres[0];                          // layer 0 of block 0, [Hp, Hp] row-major, BN folded
            for (int o = 0; o < H; ++o) {
                std::fill(acc.begin(), acc.end(), 0.0);
                double bb = (double)bres[o];
                for (int k = 0; k < H; ++k) {
                    const double w = (double)W1[(size_t)o * Hp + k];
                    if (w == 0.0) continue;
                    const float* w0row = &w0p[(size_t)k * in_dim_p];
                    for (int i = 0; i < in_dim; ++i) acc[i] += w * (double)w0row[i];
                    bb += w * (double)b0[k];
                }
                float* dst = &w0p[(size_t)(Hp + o) * in_dim_p];
                for (int i = 0; i < in_dim; ++i) dst[i] = (float)acc[i];
                b10[o] = (float)bb;
            }
        }
        std::vector<float> wres_tc(wres);
        if (fold_h) {
            net.tc_fold |= 1;
            for (int l = 0; l < 2 * nb; ++l) {
                float* Wd = &wres_tc[(size_t)l * Hp * Hp];
                for (int o = 0; o < H; ++o) split3(bres[(size_t)l * Hp + o], &Wd[(size_t)o * Hp + H]);
                // ones columns of the output: produced by the GEMM in the plain layer, carried by the residual otherwise
                if ((l & 1) == 0) for (int j = 0; j < 3; ++j) Wd[(size_t)(H + j) * Hp + H + j] = 1.f;
            }
            if (fold_in) {
                net.tc_fold |= 2;
                const bool h16 = net.tc_l1 == 1;          // the stacked first-block weights are fp16
                auto split_in = [&](float b, float* out3) { if (h16) split3h(b, out3); else split3(b, out3); };
                for (int o = 0; o < H; ++o) split_in(b0[o], &w0p[(size_t)o * in_dim_p + in_dim]);
                for (int j = 0; j < 3; ++j) w0p[(size_t)(H + j) * in_dim_p + in_dim + j] = 1.f;
                if (l1) {                   // the stacked half: bias b10 and its own ones columns (relu(1) = 1)
                    for (int o = 0; o < H; ++o) split_in(b10[o], &w0p[(size_t)(Hp + o) * in_dim_p + in_dim]);
                    for (int j = 0; j < 3; ++j) w0p[(size_t)(Hp + H + j) * in_dim_p + in_dim + j] = 1.f;
                }
            }
        }
        if (!(net.tc_fold & 2) && (fold_h || l1)) {       // explicit first-layer bias: [b0 (+ ones) | b10 (+ ones)]
            std::vector<float> b0tc(b0);
            if (fold_h) for (int j = 0; j < 3; ++j) b0tc[H + j] = 1.f;
            if (l1) {
                b0tc.insert(b0tc.end(), b10.begin(), b10.end());
                if (fold_h) for (int j = 0; j < 3; ++j) b0tc[Hp + H + j] = 1.f;
            }
            if (!net.b0_tc) DX_TRY(e->alloc(&net.b0_tc, b0tc.size()));
            DX_CUDA(cudaMemcpy(net.b0_tc, b0tc.data(), b0tc.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
        uint16_t *w0b = (uint16_t*)net.w0_bf16, *wrb = (uint16_t*)net.wres_bf16;
        if (!w0b) DX_TRY(e->alloc(&w0b, w0p.size() * (net.tc_l1 == 2 ? 2 : 1)));
        if (!wrb) DX_TRY(e->alloc(&wrb, wres_tc.size()));
        float* tmp = nullptr;
        const size_t tmp_n = std::max(std::max(w0p.size(), (size_t)128 * Hp), std::max<size_t>(wres_tc.size(), 1));
        DX_CUDA(cudaMalloc(&tmp, tmp_n * sizeof(float)));
        // copies go through the engine's (non-blocking) stream: a legacy-stream cudaMemcpy from pageable memory may
        // return before its DMA has landed, and nothing would order the conversion kernel after it
        std::vector<uint16_t> w0s;
        if (net.tc_l1 == 1) {        // stacked [W0 ; W1 W0] in fp16
            w0s.resize(w0p.size());
            for (size_t i = 0; i < w0p.size(); ++i) w0s[i] = __half_as_ushort(__float2half_rn(w0p[i]));
            cudaMemcpyAsync(w0b, w0s.data(), w0s.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, e->stream);
        } else if (net.tc_l1 == 2) { // ... or as bf16 (hi, lo) pairs, rows [hi(in_dim_p) | lo(in_dim_p)]
            w0s.resize(w0p.size() * 2);
            auto bits = [](float x) { uint32_t u; memcpy(&u, &x, 4); return (uint16_t)(u >> 16); };
            for (size_t r = 0; r < (size_t)2 * Hp; ++r)
                for (int i = 0; i < in_dim_p; ++i) {
                    const float w = w0p[r * in_dim_p + i], hi = bf16r(w), lo = bf16r(w - hi);
                    w0s[r * 2 * in_dim_p + i] = bits(hi);
                    w0s[r * 2 * in_dim_p + in_dim_p + i] = bits(lo);
                }
            cudaMemcpyAsync(w0b, w0s.data(), w0s.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, e->stream);
        } else {
            cudaMemcpyAsync(tmp, w0p.data(), w0p.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream);
            k_f32_to_bf16<<<148 * 4, 256, 0, e->stream>>>(tmp, w0b, w0p.size());
        }
        cudaStreamSynchronize(e->stream);
        if (!wres_tc.empty()) {
            cudaMemcpyAsync(tmp, wres_tc.data(), wres_tc.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream);
            k_f32_to_bf16<<<148 * 4, 256, 0, e->stream>>>(tmp, wrb, wres_tc.size());
        }
        cudaError_t ce = cudaStreamSynchronize(e->stream);
        if (ce == cudaSuccess && od > 1 && od <= 32 && tmp_n >= (size_t)128 * Hp) {
            // multi-output layer on the tensor cores: weight rows zero-padded to one 128-wide n-tile
            std::vector<float> wpad((size_t)128 * Hp, 0.f), bpad(128, 0.f);
            for (int o = 0; o < od; ++o) {
                memcpy(&wpad[(size_t)o * Hp], &wout[(size_t)o * Hp], (size_t)Hp * sizeof(float));
                bpad[o] = bout[o];
            }
            uint16_t* wob = (uint16_t*)net.wout_bf16;
            if (!wob) DX_TRY(e->alloc(&wob, wpad.size()));
            if (!net.bout_pad) DX_TRY(e->alloc(&net.bout_pad, bpad.size()));
            cudaMemcpyAsync(net.bout_pad, bpad.data(), bpad.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream);
            cudaMemcpyAsync(tmp, wpad.data(), wpad.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream);
            k_f32_to_bf16<<<148 * 4, 256, 0, e->stream>>>(tmp, wob, wpad.size());
            ce = cudaStreamSynchronize(e->stream);
            net.wout_bf16 = wob;
        }
        cudaFree(tmp);
        DX_CUDA(ce);
        net.w0_bf16 = w0b;
        net.wres_bf16 = wrb;
    }
    // NET_FP32 on the tensor cores (split-fp16 operands, k_mlp_tc.cu) where the shape allows; else the CUDA-core SGEMM path
    const bool split = e->cfg.heur_mode == DX_HEUR_NET_FP32 && (reload ? net.tc_split != nullptr : dx_mlp_split_supported(net));
    if (split) {
        DX_TRY(dx_mlp_split_upload(e->stream, net, w0p.data(), wres.data(), wout.data()));
        if (od > 1) {
            std::vector<float> bpad(128, 0.f);
            for (int o = 0; o < od; ++o) bpad[o] = bout[o];
            if (!net.bout_pad) DX_TRY(e->alloc(&net.bout_pad, bpad.size()));
            DX_CUDA(cudaMemcpy(net.bout_pad, bpad.data(), bpad.size() * sizeof(float), cudaMemcpyHostToDevice));
        }
    }
    if (reload) {
        DX_CUDA(cudaDeviceSynchronize());
        if (split) drop_graphs(e);            // the per-layer weight scales are kernel ARGUMENTS baked into the captured launches
        return DX_OK;                         // activation buffers and tensor maps stay (bf16 / SGEMM: the captured graphs too)
    }
    // activation buffers
    e->mlp.cap = cap;
    if (e->cfg.heur_mode == DX_HEUR_NET_BF16 || split) {
        int fill = 0;
        if (const char* f = getenv("DXB200_DEBUG_FILL")) fill = atoi(f);      // debug: poison value of never-written rows
        if (split) {                                                          // three [cap, 2 * Hp] arrays ([hi | lo] halves per row)
            e->mlp.pitch = 2 * (long long)Hp;
            for (int i = 0; i < 3; ++i) {
                uint16_t* q = nullptr;
                DX_TRY(e->alloc(&q, (size_t)cap * Hp * 2));
                DX_CUDA(cudaMemset(q, fill, (size_t)cap * Hp * 2 * 2));
                e->mlp.xb[i] = q;
            }
        } else {                                                              // column slices of ONE [cap, 3 * Hp] array
            e->mlp.pitch = 3 * (long long)Hp;
            uint16_t* q = nullptr;
            DX_TRY(e->alloc(&q, (size_t)cap * Hp * 3));
            DX_CUDA(cudaMemset(q, fill, (size_t)cap * Hp * 3 * 2));
            for (int i = 0; i < 3; ++i) e->mlp.xb[i] = q + (size_t)i * Hp;
        }
        DX_TRY(e->alloc(&e->mlp.partial, (size_t)cap * 16));
        uint16_t* onehot = nullptr;
        const bool gen_input = !split && dx_mlp_bf16_gen_input(net, e->dom);
        if (!gen_input) {                   // the one-hot matrix exists only for networks whose input layer is not fused
            DX_TRY(e->alloc(&onehot, (size_t)cap * in_dim_p));
            DX_CUDA(cudaMemset(onehot, fill, (size_t)cap * in_dim_p * 2));
        }
        DX_TRY(dx_mlp_bf16_prepare(net, e->mlp, onehot, gen_input, &net.tc_plan));
    } else {
        for (int i = 0; i < 3; ++i) DX_TRY(e->alloc(&e->mlp.x[i], (size_t)cap * Hp));
    }
    DX_CUDA(cudaDeviceSynchronize());        // every legacy-stream upload above has landed before the engine stream uses it
    return DX_OK;
}

// ---------------------------------------------------------------------------------------------
// HDA*-style hash-distributed single search (BASELINE config 5).  The exchange itself is done by
// the host layer (NCCL all-to-all through torch.distributed, or an in-process copy when several ranks are
// emulated on one GPU); the engine produces / consumes the per-destination record buffers.
extern "C" int dx_hda_config(dx_engine* e, int32_t rank, int32_t world) {
    if (!e) DX_FAIL(DX_ERR_ARG, "null engine");
    if (world < 1 || world > 16 || rank < 0 || rank >= world) DX_FAIL(DX_ERR_ARG, "bad rank / world (world <= 16)");
    if (e->is_q || e->is_beam) DX_FAIL(DX_ERR_UNSUPPORTED, "HDA* mode is implemented for graph_v");
    if (e->cfg.heur_mode == DX_HEUR_HOST) DX_FAIL(DX_ERR_UNSUPPORTED, "HDA* mode needs a device heuristic");
    if (e->eval_all) DX_FAIL(DX_ERR_UNSUPPORTED, "HDA* mode does not keep a Bellman-backup log (max_backups must be 0)");
    if (e->n_inst != 0) DX_FAIL(DX_ERR_STATE, "dx_hda_config must precede dx_hda_add_instance");
    if (e->max_nodes >= (1u << 28)) DX_FAIL(DX_ERR_ARG, "HDA* mode: max_nodes must be < 2^28 per rank (global node id = rank << 28 | id)");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    e->hda_rank = rank;
    e->hda_world = world;
    if (!e->hda_counts) {
        DX_TRY(e->alloc(&e->hda_counts, 16));
        DX_TRY(e->alloc(&e->hda_src_off, 32));
        DX_TRY(e->alloc(&e->hda_aux, 4));
        DX_CUDA(cudaMemset(e->hda_aux, 0, 4 * sizeof(uint32_t)));
        DX_TRY(e->alloc(&e->child_parent, (size_t)e->max_children + 1));
        DX_TRY(e->alloc(&e->child_action, (size_t)e->max_children + 1));
    }
    return DX_OK;
}

extern "C" int dx_hda_record_bytes(dx_engine* e, int32_t* rec_bytes) {
    if (!e || !rec_bytes) DX_FAIL(DX_ERR_ARG, "null argument");
    *rec_bytes = e->SP + 16;
    return DX_OK;
}

extern "C" int dx_hda_owner(dx_engine* e, const uint8_t* state, int32_t* owner) {
    if (!e || !state || !owner || e->hda_world < 1) DX_FAIL(DX_ERR_ARG, "bad argument / HDA* mode not configured");
    std::vector<uint8_t> row(e->SP, 0);            // instance id field = 0 (single instance)
    memcpy(row.data(), state, e->S);
    uint4 tmp[16];
    memcpy(tmp, row.data(), e->SP);
    *owner = (int32_t)dx_hda_owner_host(reinterpret_cast<const uint8_t*>(tmp), e->SP, (uint32_t)e->hda_world);
    return DX_OK;
}

extern "C" int dx_hda_add_instance(dx_engine* e, const uint8_t* state, const uint8_t* goal, int32_t batch_local, double weight) {
    if (!e || e->hda_world < 1) DX_FAIL(DX_ERR_STATE, "HDA* mode not configured");
    if (e->n_inst != 0) DX_FAIL(DX_ERR_STATE, "HDA* mode holds exactly one instance (dx_engine_reset first)");
    int32_t owner = 0;
    DX_TRY(dx_hda_owner(e, state, &owner));
    DX_TRY(dx_add_instances(e, state, goal, 1, &batch_local, &weight, nullptr));
    if (owner != e->hda_rank) {
        // this rank knows the instance (goal, weight, batch) but does not own the root
        e->launches += dx_launch_hda_drop_root(e->stream, e->d_ctl, e->pop_off[e->parity]);
        e->launches += dx_launch_closed_init(e->stream, e->closed);
        DX_CUDA(cudaStreamSynchronize(e->stream));
    }
    return DX_OK;
}

// Phase 1: expand the local popped list, group the children by owner rank, write the records into send_dev
// (device pointer, >= n_children * rec_bytes).  counts[world] = records per destination, in rank order.
extern "C" int dx_hda_expand(dx_engine* e, void* send_dev, int64_t send_cap_bytes, int64_t* counts) {
    if (!e || !send_dev || !counts || e->hda_world < 1) DX_FAIL(DX_ERR_ARG, "bad argument / HDA* mode not configured");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_TRY(sync_ctl(e));
    if ((int64_t)e->h_ctl->n_children * (e->SP + 16) > send_cap_bytes) DX_FAIL(DX_ERR_CAPACITY, "HDA* send buffer too small");
    cudaStream_t s = e->stream;
    const int p = e->parity;
    DX_CUDA(cudaMemsetAsync(e->hda_counts, 0, 16 * sizeof(uint32_t), s));
    e->launches += dx_launch_expand(s, e->dom, e->d_ctl, e->ia, e->node_state, e->node_g, e->popped[p], e->popped_inst[p],
                                    e->pop_off[p], e->child_state, e->child_g, e->pop_solved, e->max_pop);
    e->launches += dx_launch_goal_resolve(s, e->d_ctl, e->ia, e->node_g, e->popped[p], e->popped_inst[p], e->pop_off[p],
                                          e->pop_solved, e->max_pop);
    e->launches += dx_launch_hda_bucket(s, e->dom, e->d_ctl, e->sort, e->scan, e->child_state, e->child_g, e->popped[p],
                                        (uint32_t)e->hda_rank, (uint32_t)e->hda_world, e->hda_counts, (uint8_t*)send_dev,
                                        e->max_children);
    uint32_t hc[16];
    DX_CUDA(cudaMemcpyAsync(hc, e->hda_counts, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    DX_CUDA(cudaStreamSynchronize(s));
    for (int r = 0; r < e->hda_world; ++r) counts[r] = hc[r];
    e->hda_children_sent = e->h_ctl->n_children;
    return DX_OK;
}



// Phase 2: filter the n_records received children against the local CLOSED shard, store / evaluate / push the
// survivors, pop the local top batch.  Reports the local quantities the host reduces over ranks.
extern "C" int dx_hda_absorb(dx_engine* e, const void* recv_dev, int64_t n_records, dx_hda_local* out) {
    if (!e || !out || (n_records > 0 && !recv_dev) || e->hda_world < 1) DX_FAIL(DX_ERR_ARG, "bad argument / HDA* mode not configured");
    if (n_records > (int64_t)e->max_children) DX_FAIL(DX_ERR_CAPACITY, "HDA*: received more children than max_pop_total * actions (raise max_pop_total)");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    cudaStream_t s = e->stream;
    const int p = e->parity, q = p ^ 1;
    e->launches += dx_launch_hda_unpack(s, e->dom, e->d_ctl, (const uint8_t*)recv_dev, (uint32_t)n_records, e->child_state,
                                        e->child_g, e->child_parent, e->child_action);
    e->launches += dx_launch_closed_filter(s, e->dom, e->closed, e->d_ctl, e->ia, e->node_state, e->node_g, e->child_state,
                                           e->child_g, nullptr, e->popped_inst[p], e->child_slot, e->child_next, e->child_flags,
                                           e->max_children, e->A);
    e->launches += dx_launch_scan_flags(s, e->child_flags, e->child_prefix, &e->d_ctl->n_children, &e->d_ctl->n_kept,
                                        e->max_children, e->scan);
    k_after_scan_v<<<1, 1, 0, s>>>(e->d_ctl, e->max_nodes);
    e->launches += 1;
    e->launches += dx_launch_scatter_kept_hda(s, e->dom, e->closed, e->d_ctl, e->child_state, e->child_g, e->child_slot,
                                              e->child_flags, e->child_prefix, e->child_parent, e->child_action, e->node_state,
                                              e->node_g, e->node_parent, e->node_action, e->new_inst, e->max_children);
    DX_TRY(launch_heur_nodes(e, e->max_children));
    e->launches += dx_launch_plan(s, e->d_ctl, e->ia, e->child_prefix, e->pop_off[p], e->pop_off[q], e->open_off[p],
                                  e->open_off[q], e->A, 1, e->max_pop, e->max_open, 1);
    e->launches += dx_launch_make_keys(s, e->d_ctl, e->ia, e->node_g, e->node_h, e->new_inst, e->sort, e->max_sort);
    e->launches += dx_launch_sort(s, e->d_ctl, e->sort, e->scan, e->max_sort);
    const int max_tiles = dx_div_up((long long)e->max_open + e->max_sort, DX_SORT_TILE) + e->max_inst;
    e->launches += dx_launch_merge(s, e->d_ctl, e->ia, e->sort, e->open_key[p], e->open_val[p], e->open_off[p], e->open_key[q],
                                   e->open_val[q], e->open_off[q], e->popped[q], e->popped_inst[q], e->pop_off[q], max_tiles);
    DX_TRY(sync_ctl(e));
    uint32_t k = 0, goal = DX_NIL, nopen = 0;
    unsigned long long ff = 0, ubk = DX_EMPTY64;
    DX_CUDA(cudaMemcpy(&k, e->ia.k_pop, 4, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(&ff, e->ia.f_first, 8, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(&ubk, e->ia.ub_key, 8, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(&goal, e->ia.goal_node, 4, cudaMemcpyDeviceToHost));
    DX_CUDA(cudaMemcpy(&nopen, e->ia.n_open, 4, cudaMemcpyDeviceToHost));
    struct Local { int64_t k_popped; double f_first; double ub; int64_t goal_local; int64_t n_kept; int64_t open_size; int64_t children_sent; };
    Local* o = reinterpret_cast<Local*>(out);
    o->k_popped = k;
    o->f_first = dx_key_to_f(ff);
    if (k == 0) o->f_first = INFINITY;
    if (ubk == DX_EMPTY64) o->ub = INFINITY;
    else { uint32_t b = (uint32_t)(ubk >> 32); float f; memcpy(&f, &b, 4); o->ub = (double)f; }
    o->goal_local = goal == DX_NIL ? -1 : (int64_t)goal;
    o->n_kept = e->h_ctl->n_kept;
    o->open_size = nopen;
    o->children_sent = e->hda_children_sent;
    return DX_OK;
}

// Phase 3: identical bookkeeping on every rank from the reduced quantities.
extern "C" int dx_hda_finish(dx_engine* e, int64_t k_total, double f_first_min, double ub_min, int64_t children_total,
                             int32_t* finished) {
    if (!e || e->hda_world < 1) DX_FAIL(DX_ERR_ARG, "bad argument / HDA* mode not configured");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    const int p = e->parity, q = p ^ 1;
    unsigned long long fbits = 0;
    memcpy(&fbits, &f_first_min, 8);
    const int has_ub = std::isfinite(ub_min) ? 1 : 0;
    const float ubf = has_ub ? (float)ub_min : 0.f;
    uint32_t ubits;
    memcpy(&ubits, &ubf, 4);
    e->launches += dx_launch_hda_finish(e->stream, e->d_ctl, e->ia, e->pop_off[p], e->pop_off[q], e->A, (unsigned long long)k_total,
                                        fbits, ubits, has_ub, (unsigned long long)children_total);
    e->parity = q;
    DX_TRY(sync_ctl(e));
    if (finished) *finished = e->h_ctl->n_unfinished == 0 ? 1 : 0;
    return DX_OK;
}

// ---------------------------------------------------------------------------------------------
// Device-resident HDA* iteration: send -> (all-to-all, equal splits) -> receive -> (all-gather of 5 doubles) -> reduce.
// Nothing here synchronises with the host; the caller orders its collectives on the engine's stream
// (dx_engine_set_stream) and polls summary_dev when it wants to know whether the search has finished.
extern "C" int dx_engine_set_stream(dx_engine* e, void* cuda_stream) {
    if (!e) DX_FAIL(DX_ERR_ARG, "null engine");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_CUDA(cudaStreamSynchronize(e->stream));
    e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
    return DX_OK;
}

static int hda_send_impl(dx_engine* e, const HdaDst& dst, int64_t seg_records);

extern "C" int dx_hda_send(dx_engine* e, void* send_dev, int64_t seg_records) {
    if (!e || !send_dev || seg_records < 1 || e->hda_world < 1) DX_FAIL(DX_ERR_ARG, "bad argument / HDA* mode not configured");
    HdaDst dst = {};
    const size_t seg_bytes = 16 + (size_t)seg_records * (size_t)(e->SP + 16);
    for (int r = 0; r < e->hda_world; ++r) dst.seg[r] = (uint8_t*)send_dev + (size_t)r * seg_bytes;
    return hda_send_impl(e, dst, seg_records);
}

// Peer-memory exchange: peer_recv[r] = device address (valid in THIS process: symmetric / IPC-mapped memory) of rank r's
// receive buffer; the pack kernel stores this rank's segment for r directly into slot hda_rank of that buffer over NVLink.
// The caller places a cross-rank barrier between this call and dx_hda_receive.
extern "C" int dx_hda_send_peers(dx_engine* e, const uint64_t* peer_recv, int64_t seg_records) {
    if (!e || !peer_recv || seg_records < 1 || e->hda_world < 1) DX_FAIL(DX_ERR_ARG, "bad argument / HDA* mode not configured");
    HdaDst dst = {};
    const size_t seg_bytes = 16 + (size_t)seg_records * (size_t)(e->SP + 16);
    for (int r = 0; r < e->hda_world; ++r) {
        if (!peer_recv[r]) DX_FAIL(DX_ERR_ARG, "null peer receive buffer");
        dst.seg[r] = (uint8_t*)(uintptr_t)peer_recv[r] + (size_t)e->hda_rank * seg_bytes;
    }
    return hda_send_impl(e, dst, seg_records);
}

static int hda_send_impl(dx_engine* e, const HdaDst& dst, int64_t seg_records) {
    DX_CUDA(cudaSetDevice(e->cfg.device));
    cudaStream_t s = e->stream;
    const int p = e->parity;
    DX_CUDA(cudaMemsetAsync(e->hda_counts, 0, 16 * sizeof(uint32_t), s));
    e->launches += dx_launch_expand(s, e->dom, e->d_ctl, e->ia, e->node_state, e->node_g, e->popped[p], e->popped_inst[p],
                                    e->pop_off[p], e->child_state, e->child_g, e->pop_solved, e->max_pop);
    e->launches += dx_launch_goal_resolve(s, e->d_ctl, e->ia, e->node_g, e->popped[p], e->popped_inst[p], e->pop_off[p],
                                          e->pop_solved, e->max_pop);
    e->launches += dx_launch_hda_send(s, e->dom, e->d_ctl, e->sort, e->scan, e->child_state, e->child_g, e->popped[p],
                                      (uint32_t)e->hda_rank, (uint32_t)e->hda_world, e->hda_counts, dst,
                                      (uint32_t)seg_records, e->hda_aux, e->max_children);
    return DX_OK;
}

extern "C" int dx_hda_receive(dx_engine* e, const void* recv_dev, int64_t seg_records, double* local_dev) {
    if (!e || !recv_dev || !local_dev || seg_records < 1 || e->hda_world < 1) DX_FAIL(DX_ERR_ARG, "bad argument / HDA* mode not configured");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    cudaStream_t s = e->stream;
    const int p = e->parity, q = p ^ 1;
    e->launches += dx_launch_hda_recv(s, e->dom, e->d_ctl, (const uint8_t*)recv_dev, (uint32_t)e->hda_world, (uint32_t)seg_records,
                                      (uint32_t)e->max_children, e->hda_src_off, e->child_state, e->child_g, e->child_parent,
                                      e->child_action, e->hda_aux);
    e->launches += dx_launch_closed_filter(s, e->dom, e->closed, e->d_ctl, e->ia, e->node_state, e->node_g, e->child_state,
                                           e->child_g, nullptr, e->popped_inst[p], e->child_slot, e->child_next, e->child_flags,
                                           e->max_children, e->A);
    e->launches += dx_launch_scan_flags(s, e->child_flags, e->child_prefix, &e->d_ctl->n_children, &e->d_ctl->n_kept,
                                        e->max_children, e->scan);
    k_after_scan_v<<<1, 1, 0, s>>>(e->d_ctl, e->max_nodes);
    e->launches += 1;
    e->launches += dx_launch_scatter_kept_hda(s, e->dom, e->closed, e->d_ctl, e->child_state, e->child_g, e->child_slot,
                                              e->child_flags, e->child_prefix, e->child_parent, e->child_action, e->node_state,
                                              e->node_g, e->node_parent, e->node_action, e->new_inst, e->max_children);
    DX_TRY(launch_heur_nodes(e, e->max_children));
    e->launches += dx_launch_plan(s, e->d_ctl, e->ia, e->child_prefix, e->pop_off[p], e->pop_off[q], e->open_off[p],
                                  e->open_off[q], e->A, 1, e->max_pop, e->max_open, 1);
    e->launches += dx_launch_make_keys(s, e->d_ctl, e->ia, e->node_g, e->node_h, e->new_inst, e->sort, e->max_sort);
    e->launches += dx_launch_sort(s, e->d_ctl, e->sort, e->scan, e->max_sort);
    const int max_tiles = dx_div_up((long long)e->max_open + e->max_sort, DX_SORT_TILE) + e->max_inst;
    e->launches += dx_launch_merge(s, e->d_ctl, e->ia, e->sort, e->open_key[p], e->open_val[p], e->open_off[p], e->open_key[q],
                                   e->open_val[q], e->open_off[q], e->popped[q], e->popped_inst[q], e->pop_off[q], max_tiles);
    e->launches += dx_launch_hda_locals(s, e->d_ctl, e->ia, e->hda_aux, local_dev);
    return DX_OK;
}

extern "C" int dx_hda_reduce(dx_engine* e, const double* all_dev, double* summary_dev) {
    if (!e || !all_dev || !summary_dev || e->hda_world < 1) DX_FAIL(DX_ERR_ARG, "bad argument / HDA* mode not configured");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    const int p = e->parity, q = p ^ 1;
    e->launches += dx_launch_hda_reduce(e->stream, e->d_ctl, e->ia, e->pop_off[p], e->pop_off[q], e->A, e->hda_world, all_dev,
                                        summary_dev);
    e->parity = q;
    return DX_OK;
}

// One node of the local shard: state row, global parent id (0xFFFFFFFF for the root), action, g.
extern "C" int dx_hda_node(dx_engine* e, int64_t local_id, uint8_t* state, uint32_t* parent_global, int32_t* action, float* g) {
    if (!e || local_id < 0) DX_FAIL(DX_ERR_ARG, "bad argument");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_TRY(sync_ctl(e));
    if (local_id >= (int64_t)e->h_ctl->node_count) DX_FAIL(DX_ERR_ARG, "node id out of range");
    std::vector<uint8_t> row(e->SP);
    uint8_t act = 0;
    DX_CUDA(cudaMemcpy(row.data(), e->node_state + (size_t)local_id * e->SP, e->SP, cudaMemcpyDeviceToHost));
    if (state) memcpy(state, row.data(), e->S);
    if (parent_global) DX_CUDA(cudaMemcpy(parent_global, e->node_parent + local_id, 4, cudaMemcpyDeviceToHost));
    if (action) { DX_CUDA(cudaMemcpy(&act, e->node_action + local_id, 1, cudaMemcpyDeviceToHost)); *action = act; }
    if (g) DX_CUDA(cudaMemcpy(g, e->node_g + local_id, 4, cudaMemcpyDeviceToHost));
    return DX_OK;
}

int dx_tc_profile_read(unsigned long long* out16);   // k_mlp_tc.cu
int dx_small_profile_read(unsigned long long* out16);   // k_mlp_small.cu
int dx_tail_profile_read(unsigned long long* out16);    // k_open.cu
int dx_filter_profile_read(unsigned long long* out8);   // k_closed.cu
extern "C" int dx_debug_tc_profile(dx_engine* e, uint64_t* out8) {   /* out8: 16 entries */
    if (!e || !out8) DX_FAIL(DX_ERR_ARG, "null argument");
    DX_CUDA(cudaSetDevice(e->cfg.device));
    DX_CUDA(cudaStreamSynchronize(e->stream));
    if (getenv("DXB200_PROF_TAIL")) {                  // (-DYT_PROFILE timelines: tail kernel in slots 0..8, filter kernel in 9..14)
        unsigned long long f[8];
        DX_TRY(dx_tail_profile_read(reinterpret_cast<unsigned long long*>(out8)));
        DX_TRY(dx_filter_profile_read(f));
        for (int i = 1; i < 8; ++i) out8[8 + i] = f[i];   // (the filter runs once per tail launch: its launch counter is not reported)
        return DX_OK;
    }
    if (e->net.small) return dx_small_profile_read(reinterpret_cast<unsigned long long*>(out8));   // (-DSM_PROFILE timeline)
    return dx_tc_profile_read(reinterpret_cast<unsigned long long*>(out8));
}
