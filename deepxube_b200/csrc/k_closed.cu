// CLOSED: open-addressing hash set keyed on (instance, state) with a full-row compare, storing the best
// path cost g per state, plus the exact *sequential* filter semantics of the reference.
//
// Reference (deepxube/pathfinding/graph_search.py:73-80): children are visited one by one, in order;
//     keep child iff state not in CLOSED or CLOSED[state] > g;  then CLOSED[state] = g.
// So inside one batch a child is kept iff its g is strictly below CLOSED-before-the-batch AND strictly
// below the g of every EARLIER same-state child of the batch (a later, strictly cheaper duplicate is kept
// too).  A plain atomic-min insert is not equivalent.  Parallel formulation used here:
//   1. k_closed_link    every item finds / claims its slot (64-bit CAS on (tag | rep)), then pushes
//                       itself onto the slot's per-iteration list ((epoch | head) CAS, child_next links).
//   2. k_closed_decide  every item walks its slot's list (the same-state items of this batch, usually
//                       1-3) and evaluates the sequential rule directly from (g, batch position).
//                       The kept item with the lexicographically smallest (g, position) is "final": it
//                       is the value CLOSED holds after the batch.
//   3. the final item writes CLOSED[state] = g and, for a slot claimed in this batch, replaces the
//      pending batch reference by its node id (k_scatter_kept for A*, k_closed_commit for Q*).
// A slot's representative row is either a node-store row (rep = node id) or, while the batch that
// created it is in flight, a row of the child batch (rep = DX_PENDING | child index): a state that is new
// always has a kept child (its first occurrence beats CLOSED = +inf), so the pending reference is
// always resolved before the next iteration.
//
// Two uses: A* filters generated children (item rows = child batch, group = A items per popped node);
// Q* filters popped nodes (item rows = node-store rows of the popped list, group = 1; graph_search.py:132-133).
#include "dx_internal.cuh"

#define CL_THREADS 256

__global__ void k_closed_init(ClosedDev c) {
    const size_t cap = (size_t)c.mask + 1;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += (size_t)gridDim.x * blockDim.x) {
        ClosedSlot z;
        z.key = DX_EMPTY64; z.head = 0ull; z.g = __int_as_float(0x7f800000); z.pad_[0] = z.pad_[1] = z.pad_[2] = 0u;
        c.slots[i] = z;
    }
}

int dx_launch_closed_init(cudaStream_t s, const ClosedDev& c) {
    k_closed_init<<<148 * 8, 256, 0, s>>>(c);
    return 1;
}

__device__ __forceinline__ uint32_t closed_find_or_claim(const ClosedDev& c, const uint4* __restrict__ row, int chunks,
                                                         uint32_t claim_rep, const uint8_t* __restrict__ node_state,
                                                         const uint8_t* __restrict__ batch_rows, int SP) {
    const uint64_t h = dx_hash_row(row, chunks);
    const uint32_t tag = (uint32_t)(h >> 32);
    uint32_t i = (uint32_t)h & c.mask;
    const unsigned long long mine = ((unsigned long long)tag << 32) | claim_rep;
    // the table holds >= 2x (max_nodes + max items per batch) slots, so a probe sequence always ends;
    // the bound only guards against a corrupted table (never spin forever on the device)
    for (uint32_t probes = 0; probes <= c.mask; ++probes) {
        // CAS first: a new state (the common case) claims its slot in ONE round trip; an occupied slot returns its key
        const unsigned long long k = atomicCAS(&c.slots[i].key, DX_EMPTY64, mine);
        if (k == DX_EMPTY64) return i;
        if ((uint32_t)(k >> 32) == tag) {
            const uint32_t rep = (uint32_t)k;
            const uint8_t* other = (rep & DX_PENDING) ? (batch_rows + (size_t)(rep & ~DX_PENDING) * SP)
                                                      : (node_state + (size_t)rep * SP);
            if (dx_rows_equal(row, reinterpret_cast<const uint4*>(other), chunks)) return i;
        }
        i = (i + 1) & c.mask;
    }
    return DX_NIL;
}

// Root of an A* instance: CLOSED[root] = 0.0 (graph_search.py:119-121).
__global__ void k_closed_seed(DomainDev d, ClosedDev c, const uint8_t* __restrict__ node_state, uint32_t first, uint32_t n) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t id = first + t;
        const uint4* row = reinterpret_cast<const uint4*>(node_state + (size_t)id * d.SP);
        uint32_t i = closed_find_or_claim(c, row, d.SP / 16, id, node_state, node_state, d.SP);
        if (i != DX_NIL) c.slots[i].g = 0.0f;
    }
}

int dx_launch_closed_seed(cudaStream_t s, const DomainDev& d, const ClosedDev& c, const uint8_t* node_state, uint32_t first_node,
                          uint32_t n) {
    if (n == 0) return 0;
    k_closed_seed<<<dx_div_up(n, 128), 128, 0, s>>>(d, c, node_state, first_node, n);
    return 1;
}

// item -> (row pointer, g, owning instance active?)
struct ItemView {
    const uint8_t* batch_rows;    // A*: child rows ; Q*: unused
    const float* batch_g;         // A*: child g (NaN = dead)
    const uint32_t* item_node;    // Q*: popped node ids (nullptr for A*)
    const uint8_t* node_state;
    const float* node_g;
    const uint32_t* popped_inst;
    const uint8_t* active;
    int group;
    int SP;
};

__device__ __forceinline__ const uint8_t* item_row(const ItemView& v, uint32_t c) {
    return v.item_node ? v.node_state + (size_t)v.item_node[c] * v.SP : v.batch_rows + (size_t)c * v.SP;
}
__device__ __forceinline__ float item_g(const ItemView& v, uint32_t c) {
    if (v.item_node) return v.active[v.popped_inst[c]] ? v.node_g[v.item_node[c]] : __int_as_float(0x7fc00000);
    return v.batch_g[c];
}

__device__ __forceinline__ void closed_link_item(const ClosedDev& c, const Ctl* __restrict__ ctl, const ItemView& v,
                                                 uint32_t* __restrict__ child_slot, uint32_t* __restrict__ child_next, uint32_t ci,
                                                 uint32_t epoch, int chunks) {
    const float g = item_g(v, ci);
    if (g != g) { child_slot[ci] = DX_NIL; return; }      // item of a finished instance
    const uint4* row = reinterpret_cast<const uint4*>(item_row(v, ci));
    const uint32_t claim = v.item_node ? v.item_node[ci] : (DX_PENDING | ci);
    const uint32_t i = closed_find_or_claim(c, row, chunks, claim, v.node_state, v.batch_rows, v.SP);
    child_slot[ci] = i;
    if (i == DX_NIL) { atomicOr((uint32_t*)&ctl->error, DX_DEVERR_NODES); return; }
    // push on the slot's per-iteration list: nobody walks the lists before the link phase is over (kernel boundary or block
    // barrier), so one exchange is enough -- the old head (if it is of this iteration) becomes this item's successor
    const unsigned long long old = atomicExch(&c.slots[i].head, ((unsigned long long)epoch << 32) | ci);
    child_next[ci] = ((uint32_t)(old >> 32) == epoch) ? (uint32_t)old : DX_NIL;
}

__global__ void __launch_bounds__(CL_THREADS)
k_closed_link(ClosedDev c, const Ctl* __restrict__ ctl, ItemView v, uint32_t* __restrict__ child_slot,
              uint32_t* __restrict__ child_next) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children;
    const uint32_t epoch = ctl->epoch;
    const int chunks = v.SP / 16;
    for (uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n; ci += gridDim.x * blockDim.x)
        closed_link_item(c, ctl, v, child_slot, child_next, ci, epoch, chunks);
}

// flags: bit0 = keep, bit1 = final (this item's g becomes CLOSED[state])
__device__ __forceinline__ uint8_t closed_decide_item(const ClosedDev& c, const ItemView& v, const uint32_t* __restrict__ child_slot,
                                                      const uint32_t* __restrict__ child_next, uint32_t ci) {
    const uint32_t i = child_slot[ci];
    if (i == DX_NIL) return 0;
    const float g = item_g(v, ci);
    bool keep = g < c.slots[i].g;
    bool fin = keep;
    uint32_t c2 = (uint32_t)c.slots[i].head;
    while (c2 != DX_NIL) {
        if (c2 != ci) {
            const float g2 = item_g(v, c2);
            if (c2 < ci && g2 <= g) keep = false;                    // an earlier duplicate already set CLOSED <= g
            if (g2 < g || (g2 == g && c2 < ci)) fin = false;
        }
        c2 = child_next[c2];
    }
    return (keep ? 1 : 0) | ((keep && fin) ? 2 : 0);
}

__global__ void __launch_bounds__(CL_THREADS)
k_closed_decide(ClosedDev c, const Ctl* __restrict__ ctl, ItemView v, const uint32_t* __restrict__ child_slot,
                const uint32_t* __restrict__ child_next, uint8_t* __restrict__ child_flags) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children;
    for (uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n; ci += gridDim.x * blockDim.x)
        child_flags[ci] = closed_decide_item(c, v, child_slot, child_next, ci);
}

int dx_launch_closed_filter(cudaStream_t s, const DomainDev& d, const ClosedDev& c, const Ctl* ctl, const InstArrays& ia,
                            const uint8_t* node_state, const float* node_g, const uint8_t* child_state, const float* child_g,
                            const uint32_t* item_node, const uint32_t* popped_inst, uint32_t* child_slot, uint32_t* child_next,
                            uint8_t* child_flags, int max_children, int group) {
    ItemView v{child_state, child_g, item_node, node_state, node_g, popped_inst, ia.active, group, d.SP};
    int blocks = min(dx_div_up(max_children, CL_THREADS), 148 * 8);
    k_closed_link<<<blocks, CL_THREADS, 0, s>>>(c, ctl, v, child_slot, child_next);
    k_closed_decide<<<blocks, CL_THREADS, 0, s>>>(c, ctl, v, child_slot, child_next, child_flags);
    return 2;
}

// A*: kept children become nodes (ids in flat child order = push order); final items commit CLOSED.
struct ScatterArgs {
    const uint8_t* child_state; const float* child_g; const uint32_t* child_slot; const uint8_t* child_flags;
    const uint32_t* child_prefix; const uint32_t* popped; const uint32_t* popped_inst;
    uint8_t* node_state; float* node_g; uint32_t* node_parent; uint8_t* node_action; uint32_t* new_inst;
    const float* child_h; float* node_h;
};

__device__ __forceinline__ void scatter_kept_item(const DomainDev& d, const ClosedDev& c, const ScatterArgs& a, uint32_t ci,
                                                  uint32_t base, int chunks) {
    const uint8_t f = a.child_flags[ci];
    if (!(f & 1)) return;
    // every load first (the pointers may alias as far as the compiler knows: a load placed after a store waits for the data the
    // store needs, which turns a row copy into a chain of round trips), then the stores
    const uint32_t j = ci / d.A;
    const uint32_t k = a.child_prefix[ci];
    const float g = a.child_g[ci];
    const uint32_t par = a.popped[j], inst = a.popped_inst[j];
    const uint32_t slot = (f & 2) ? a.child_slot[ci] : 0u;
    const unsigned long long skey = (f & 2) ? c.slots[slot].key : 0ull;      // (read before the slot's sector is written below)
    const float h = a.child_h ? a.child_h[ci] : 0.f;
    const uint32_t id = base + k;
    const uint4* src = reinterpret_cast<const uint4*>(a.child_state + (size_t)ci * d.SP);
    uint4* dst = reinterpret_cast<uint4*>(a.node_state + (size_t)id * d.SP);
    for (int q = 0; q < chunks; q += 4) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (q + u < chunks) r[u] = src[q + u];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (q + u < chunks) dst[q + u] = r[u];
    }
    a.node_g[id] = g;
    a.node_parent[id] = par;
    a.node_action[id] = (uint8_t)(ci % d.A);
    a.new_inst[k] = inst;
    if (a.child_h) a.node_h[id] = h;                    // evaluate-all engines: the value was computed before the filter
    if (f & 2) {
        c.slots[slot].g = g;
        if ((uint32_t)skey & DX_PENDING) c.slots[slot].key = (skey & 0xFFFFFFFF00000000ull) | id;
    }
}

__global__ void __launch_bounds__(CL_THREADS)
k_scatter_kept(DomainDev d, ClosedDev c, Ctl* __restrict__ ctl, ScatterArgs a) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children;
    const uint32_t base = ctl->node_base;
    const int chunks = d.SP / 16;
    for (uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n; ci += gridDim.x * blockDim.x)
        scatter_kept_item(d, c, a, ci, base, chunks);
}

// Small batches (the host bounds the iteration's children by DX_SMALL_SORT, e.g. `graph.100B` on cube3 = 1200): CLOSED link +
// decide, the kept-children scan with its bookkeeping (k_after_scan_v) and the scatter in ONE block (plus k_goal_resolve, which
// only needs the expansion kernel finished) -- eight launches whose cost at this size is mostly their dependency latency.  Same device functions, same order of
// phases; __syncthreads() is the grid-wide barrier the separate launches provided.
#define SMALL_THREADS 1024
// -DYT_PROFILE: clock64 timeline of the block (thread 0): g_filter_prof[0] launches, [1] goal resolve, [2] link, [3] decide, [4] scan,
// [5] scatter (read and cleared with the tail kernel's timeline, slots 9..13 of dx_debug_tc_profile under DXB200_PROF_TAIL)
__device__ unsigned long long g_filter_prof[8];
#ifdef YT_PROFILE
#define SF_MARK(slot) do { if (threadIdx.x == 0) { const long long _t = clock64(); \
        atomicAdd(&g_filter_prof[(slot)], (unsigned long long)(_t - t_prev)); t_prev = _t; } } while (0)
#else
#define SF_MARK(slot) do { } while (0)
#endif
int dx_filter_profile_read(unsigned long long* out8) {
    if (cudaMemcpyFromSymbol(out8, g_filter_prof, sizeof(unsigned long long) * 8) != cudaSuccess) return DX_ERR_CUDA;
    unsigned long long z[8] = {0};
    if (cudaMemcpyToSymbol(g_filter_prof, z, sizeof(z)) != cudaSuccess) return DX_ERR_CUDA;
    return DX_OK;
}
__global__ void __launch_bounds__(SMALL_THREADS)
k_small_filter_v(DomainDev d, ClosedDev c, Ctl* __restrict__ ctl, ItemView v, uint32_t* __restrict__ child_slot,
                 uint32_t* __restrict__ child_next, uint8_t* __restrict__ child_flags, uint32_t* __restrict__ child_prefix,
                 ScatterArgs a, uint32_t max_nodes, InstArrays ia, const uint32_t* __restrict__ pop_off,
                 const uint8_t* __restrict__ pop_solved) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_err;
    dx_pdl_prologue();
#ifdef YT_PROFILE
    long long t_prev = clock64();
    if (threadIdx.x == 0) atomicAdd(&g_filter_prof[0], 1ull);
#endif
    if (ctl->error) return;
    // k_goal_resolve (k_domain.cu): the solved popped node that won the (g, pop order) atomicMin becomes goal_node
    for (uint32_t j = threadIdx.x; j < ctl->n_pop; j += SMALL_THREADS) {
        if (!pop_solved[j]) continue;
        const uint32_t inst = a.popped_inst[j], id = a.popped[j];
        const unsigned long long key = ((unsigned long long)__float_as_uint(a.node_g[id]) << 32) |
                                       (unsigned long long)(ia.pop_seq_base[inst] + j - pop_off[inst]);
        if (ia.ub_key[inst] == key) ia.goal_node[inst] = id;
    }
    const uint32_t n = ctl->n_children;
    if (n > 4u * SMALL_THREADS) {            // the host's bound was wrong: fail loudly
        if (threadIdx.x == 0) atomicOr(&ctl->error, DX_DEVERR_POP);
        return;
    }
    const uint32_t epoch = ctl->epoch;
    const int chunks = v.SP / 16;
    SF_MARK(1);
    for (uint32_t ci = threadIdx.x; ci < n; ci += SMALL_THREADS) closed_link_item(c, ctl, v, child_slot, child_next, ci, epoch, chunks);
    __syncthreads();
    if (threadIdx.x == 0) s_err = *((volatile uint32_t*)&ctl->error);
    __syncthreads();
    if (s_err) return;
    SF_MARK(2);
    for (uint32_t ci = threadIdx.x; ci < n; ci += SMALL_THREADS) child_flags[ci] = closed_decide_item(c, v, child_slot, child_next, ci);
    __syncthreads();
    SF_MARK(3);
    // exclusive scan of (flags & 1): 4 consecutive items per thread, prefix[n] = total (as dx_launch_scan_flags)
    const uint32_t base4 = threadIdx.x * 4;
    uint32_t f4[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { f4[i] = (base4 + i < n) ? (child_flags[base4 + i] & 1u) : 0u; sum += f4[i]; }
    uint32_t x = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        s_warp[lane] = w;
    }
    __syncthreads();
    const uint32_t total = s_warp[31];
    uint32_t ex = (warp ? s_warp[warp - 1] : 0) + x - sum;
#pragma unroll
    for (int i = 0; i < 4; ++i) { if (base4 + i < n) child_prefix[base4 + i] = ex; ex += f4[i]; }
    if (base4 <= n && n < base4 + 4) child_prefix[n] = ex;
    if (threadIdx.x == 0) {                  // k_after_scan_v
        ctl->n_kept = total;
        if ((unsigned long long)ctl->node_base + total > max_nodes) {
            ctl->error |= DX_DEVERR_NODES;
        } else {
            ctl->node_count = ctl->node_base + total;
            ctl->nn_row_start = ctl->node_base;
            ctl->nn_n = total;
        }
        s_err = ctl->error;
    }
    __syncthreads();
    if (s_err) return;
    SF_MARK(4);
    const uint32_t nbase = ctl->node_base;
    for (uint32_t ci = threadIdx.x; ci < n; ci += SMALL_THREADS) scatter_kept_item(d, c, a, ci, nbase, chunks);
    SF_MARK(5);
}

// The same block, restructured around what the timeline showed (-DYT_PROFILE: link 17 k, scatter 12 k, decide 7 k of 40 k cycles, all
// of them chains of dependent L2 / DRAM round trips on ONE SM, walked twice by the 176 threads that own a second item):
//   * a thread's (up to) two items advance TOGETHER through every phase -- both CAS probes, both row compares, both list pushes
//     and both slot reads are in flight at once, so a second item costs issue slots, not another latency chain;
//   * per-item state (g, slot, tag, "slot claimed by this batch", flags, prefix) stays in REGISTERS across the phases, the two
//     things other threads need (g and the same-state list links) live in shared memory -- no child_slot / child_next /
//     child_flags / child_prefix round trips between phases, and the commit no longer reads the slot key back;
//   * the kept-children scan is two ballot scans (items tid and tid + 1024 keep flat child order: round 0 precedes round 1).
// Same decisions as closed_link_item / closed_decide_item / scatter_kept_item; child_flags and child_prefix are still written
// (the OPEN side reads the prefix at the instance boundaries).
__global__ void __launch_bounds__(SMALL_THREADS)
k_small_filter_v2(DomainDev d, ClosedDev c, Ctl* __restrict__ ctl, ItemView v, uint8_t* __restrict__ child_flags,
                  uint32_t* __restrict__ child_prefix, ScatterArgs a, uint32_t max_nodes, InstArrays ia,
                  const uint32_t* __restrict__ pop_off, const uint8_t* __restrict__ pop_solved) {
    __shared__ float s_g[2 * SMALL_THREADS];
    __shared__ uint32_t s_next[2 * SMALL_THREADS];
    __shared__ uint32_t s_warp[2][32];
    __shared__ uint32_t s_err;
    dx_pdl_prologue();
#ifdef YT_PROFILE
    long long t_prev = clock64();
    if (threadIdx.x == 0) atomicAdd(&g_filter_prof[0], 1ull);
#endif
    if (ctl->error) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t j = tid; j < ctl->n_pop; j += SMALL_THREADS) {                    // k_goal_resolve
        if (!pop_solved[j]) continue;
        const uint32_t inst = a.popped_inst[j], id = a.popped[j];
        const unsigned long long key = ((unsigned long long)__float_as_uint(a.node_g[id]) << 32) |
                                       (unsigned long long)(ia.pop_seq_base[inst] + j - pop_off[inst]);
        if (ia.ub_key[inst] == key) ia.goal_node[inst] = id;
    }
    const uint32_t n = ctl->n_children;
    if (n > 2u * SMALL_THREADS) {            // the host's bound was wrong: fail loudly
        if (tid == 0) atomicOr(&ctl->error, DX_DEVERR_POP);
        return;
    }
    const uint32_t epoch = ctl->epoch, nbase = ctl->node_base;
    const int chunks = v.SP / 16;
    SF_MARK(1);
    // ---- link: find / claim the slot of each item, push it on the slot's list of this iteration ------------------------------
    uint32_t ci[2], slot[2], tag[2], idx[2], par[2], inst[2];
    float g[2];
    bool live[2], done[2], pend[2];
    uint4 rv[2][4];                                  // the items' rows (SP <= 64), loaded ONCE: every chunk load of both items is
#pragma unroll                                       // in flight together (a per-chunk loop serialises one L2 round trip per chunk)
    for (int r = 0; r < 2; ++r) {
        ci[r] = (uint32_t)tid + (uint32_t)r * SMALL_THREADS;
        const bool valid = ci[r] < n;
        g[r] = valid ? v.batch_g[ci[r]] : __int_as_float(0x7fc00000);
        par[r] = valid ? a.popped[ci[r] / d.A] : 0u;
        inst[r] = valid ? a.popped_inst[ci[r] / d.A] : 0u;
        const uint4* row = reinterpret_cast<const uint4*>(v.batch_rows + (size_t)ci[r] * v.SP);
#pragma unroll
        for (int q = 0; q < 4; ++q) rv[r][q] = (valid && q < chunks) ? row[q] : make_uint4(0, 0, 0, 0);
        slot[r] = DX_NIL; pend[r] = false; tag[r] = 0; idx[r] = 0;
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        live[r] = g[r] == g[r];                      // (NaN: beyond n, or the child of a finished instance)
        done[r] = !live[r];
        if (ci[r] < n) s_g[ci[r]] = g[r];
        if (live[r]) {
            uint64_t h = 0x9E3779B97F4A7C15ull;      // dx_hash_row over the register copy
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (q < chunks) {
                    const uint64_t x = ((uint64_t)rv[r][q].y << 32) | rv[r][q].x, y = ((uint64_t)rv[r][q].w << 32) | rv[r][q].z;
                    h = dx_mix64(h ^ x) + 0x632BE59BD9B4E019ull;
                    h = dx_mix64(h ^ y) + 0x9E3779B97F4A7C15ull;
                }
            tag[r] = (uint32_t)(h >> 32);
            idx[r] = (uint32_t)h & c.mask;
        }
    }
    SF_MARK(6);
    for (uint32_t probes = 0; probes <= c.mask && !(done[0] && done[1]); ++probes) {
        unsigned long long k[2] = {0ull, 0ull};
#pragma unroll
        for (int r = 0; r < 2; ++r)                  // CAS first: a new state claims its slot in ONE round trip
            if (!done[r]) k[r] = atomicCAS(&c.slots[idx[r]].key, DX_EMPTY64, ((unsigned long long)tag[r] << 32) | (DX_PENDING | ci[r]));
        const uint4* other[2] = {nullptr, nullptr};
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (done[r]) continue;
            if (k[r] == DX_EMPTY64) { done[r] = true; slot[r] = idx[r]; pend[r] = true; continue; }
            if ((uint32_t)(k[r] >> 32) == tag[r]) {
                const uint32_t rep = (uint32_t)k[r];
                other[r] = reinterpret_cast<const uint4*>((rep & DX_PENDING) ? (v.batch_rows + (size_t)(rep & ~DX_PENDING) * v.SP)
                                                                            : (v.node_state + (size_t)rep * v.SP));
            } else {
                idx[r] = (idx[r] + 1) & c.mask;
            }
        }
        uint4 ov[2][4];                               // the representatives' rows: both items' loads in flight together
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) ov[r][q] = (other[r] && q < chunks) ? other[r][q] : make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (!other[r]) continue;
            bool eq = true;
#pragma unroll
            for (int q = 0; q < 4; ++q)
                eq = eq && ov[r][q].x == rv[r][q].x && ov[r][q].y == rv[r][q].y && ov[r][q].z == rv[r][q].z && ov[r][q].w == rv[r][q].w;
            if (eq) { done[r] = true; slot[r] = idx[r]; pend[r] = ((uint32_t)k[r] & DX_PENDING) != 0; }
            else idx[r] = (idx[r] + 1) & c.mask;
        }
    }
    if ((live[0] && !done[0]) || (live[1] && !done[1])) atomicOr(&ctl->error, DX_DEVERR_NODES);
    SF_MARK(7);
    {
        unsigned long long old[2] = {0ull, 0ull};
#pragma unroll
        for (int r = 0; r < 2; ++r)                  // nobody walks the lists before the barrier below: one exchange is enough
            if (live[r] && slot[r] != DX_NIL) old[r] = atomicExch(&c.slots[slot[r]].head, ((unsigned long long)epoch << 32) | ci[r]);
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (ci[r] < n) s_next[ci[r]] = ((uint32_t)(old[r] >> 32) == epoch) ? (uint32_t)old[r] : DX_NIL;
    }
    __syncthreads();
    if (tid == 0) s_err = *((volatile uint32_t*)&ctl->error);
    __syncthreads();
    if (s_err) return;
    SF_MARK(2);
    // ---- decide: the sequential rule from (g, batch position) over the same-state items of the batch ---------------------------
    uint32_t flag[2] = {0u, 0u};
    {
        float sg[2] = {0.f, 0.f};
        uint32_t hd[2] = {DX_NIL, DX_NIL};
#pragma unroll
        for (int r = 0; r < 2; ++r)
            if (live[r]) { sg[r] = c.slots[slot[r]].g; hd[r] = (uint32_t)c.slots[slot[r]].head; }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (!live[r]) continue;
            bool keep = g[r] < sg[r];
            bool fin = keep;
            uint32_t c2 = hd[r];
            while (c2 != DX_NIL) {
                if (c2 != ci[r]) {
                    const float g2 = s_g[c2];
                    if (c2 < ci[r] && g2 <= g[r]) keep = false;                  // an earlier duplicate already set CLOSED <= g
                    if (g2 < g[r] || (g2 == g[r] && c2 < ci[r])) fin = false;
                }
                c2 = s_next[c2];
            }
            flag[r] = (keep ? 1u : 0u) | ((keep && fin) ? 2u : 0u);
        }
    }
    SF_MARK(3);
    // ---- kept-children scan: ids in flat child order = push order --------------------------------------------------------------
    uint32_t pre[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const uint32_t bal = __ballot_sync(0xffffffffu, (flag[r] & 1u) != 0);
        pre[r] = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) s_warp[r][warp] = __popc(bal);
    }
    __syncthreads();
    if (warp < 2) {
        uint32_t w = s_warp[warp][lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        s_warp[warp][lane] = w;                      // inclusive over the warps of round `warp`
    }
    __syncthreads();
    const uint32_t tot0 = s_warp[0][31], total = tot0 + s_warp[1][31];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        pre[r] += (r ? tot0 : 0u) + (warp ? s_warp[r][warp - 1] : 0u);
        if (ci[r] < n) { child_flags[ci[r]] = (uint8_t)flag[r]; child_prefix[ci[r]] = pre[r]; }
    }
    if (tid == 0) {                          // k_after_scan_v
        child_prefix[n] = total;
        ctl->n_kept = total;
        if ((unsigned long long)nbase + total > max_nodes) {
            ctl->error |= DX_DEVERR_NODES;
            s_err = 1;
        } else {
            ctl->node_count = nbase + total;
            ctl->nn_row_start = nbase;
            ctl->nn_n = total;
        }
    }
    __syncthreads();
    if (s_err) return;
    SF_MARK(4);
    // ---- kept children become nodes; final items commit CLOSED (no slot read: the link phase knows who claimed the slot) ------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (!(flag[r] & 1u)) continue;
        const uint32_t id = nbase + pre[r];
        uint4* dst = reinterpret_cast<uint4*>(a.node_state + (size_t)id * d.SP);
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q < chunks) dst[q] = rv[r][q];
        a.node_g[id] = g[r];
        a.node_parent[id] = par[r];
        a.node_action[id] = (uint8_t)(ci[r] % d.A);
        a.new_inst[pre[r]] = inst[r];
        if (flag[r] & 2u) {
            c.slots[slot[r]].g = g[r];
            if (pend[r]) c.slots[slot[r]].key = ((unsigned long long)tag[r] << 32) | id;
        }
    }
    SF_MARK(5);
}

// Q*, small batches: is_solved + record_goal of the popped nodes (k_check_solved, k_goal_resolve), their CLOSED test (link +
// decide), the scan of the kept ones and the CLOSED commit in ONE block (eight launches).
__global__ void __launch_bounds__(SMALL_THREADS)
k_small_filter_q(DomainDev d, ClosedDev c, Ctl* __restrict__ ctl, ItemView v, InstArrays ia, const uint32_t* __restrict__ pop_off,
                 uint8_t* __restrict__ pop_solved, uint32_t* __restrict__ child_slot, uint32_t* __restrict__ child_next,
                 uint8_t* __restrict__ child_flags, uint32_t* __restrict__ child_prefix) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_err;
    dx_pdl_prologue();
    if (ctl->error) return;
    const uint32_t n = ctl->n_children;          // = n_pop: the items are the popped nodes
    if (n > 4u * SMALL_THREADS || ctl->n_pop != n) {
        if (threadIdx.x == 0) atomicOr(&ctl->error, DX_DEVERR_POP);
        return;
    }
    for (uint32_t j = threadIdx.x; j < n; j += SMALL_THREADS) {                       // k_check_solved
        const uint32_t inst = v.popped_inst[j], id = v.item_node[j];
        bool solved = false;
        if (ia.active[inst]) {
            solved = row_solved(d.kind, d.dim, d.S, v.node_state + (size_t)id * d.SP, ia.goals + (size_t)inst * d.SP);
            if (solved) {
                const unsigned long long key = ((unsigned long long)__float_as_uint(v.node_g[id]) << 32) |
                                               (unsigned long long)(ia.pop_seq_base[inst] + j - pop_off[inst]);
                atomicMin(&ia.ub_key[inst], key);
            }
        }
        pop_solved[j] = solved ? 1 : 0;
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < n; j += SMALL_THREADS) {                       // k_goal_resolve
        if (!pop_solved[j]) continue;
        const uint32_t inst = v.popped_inst[j], id = v.item_node[j];
        const unsigned long long key = ((unsigned long long)__float_as_uint(v.node_g[id]) << 32) |
                                       (unsigned long long)(ia.pop_seq_base[inst] + j - pop_off[inst]);
        if (*((volatile unsigned long long*)&ia.ub_key[inst]) == key) ia.goal_node[inst] = id;
    }
    const uint32_t epoch = ctl->epoch;
    const int chunks = v.SP / 16;
    for (uint32_t ci = threadIdx.x; ci < n; ci += SMALL_THREADS) closed_link_item(c, ctl, v, child_slot, child_next, ci, epoch, chunks);
    __syncthreads();
    if (threadIdx.x == 0) s_err = *((volatile uint32_t*)&ctl->error);
    __syncthreads();
    if (s_err) return;
    for (uint32_t ci = threadIdx.x; ci < n; ci += SMALL_THREADS) child_flags[ci] = closed_decide_item(c, v, child_slot, child_next, ci);
    __syncthreads();
    const uint32_t base4 = threadIdx.x * 4;
    uint32_t f4[4], sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { f4[i] = (base4 + i < n) ? (child_flags[base4 + i] & 1u) : 0u; sum += f4[i]; }
    uint32_t x = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
        s_warp[lane] = w;
    }
    __syncthreads();
    uint32_t ex = (warp ? s_warp[warp - 1] : 0) + x - sum;
#pragma unroll
    for (int i = 0; i < 4; ++i) { if (base4 + i < n) child_prefix[base4 + i] = ex; ex += f4[i]; }
    if (base4 <= n && n < base4 + 4) child_prefix[n] = ex;
    if (threadIdx.x == 0) ctl->n_kept = s_warp[31];
    for (uint32_t ci = threadIdx.x; ci < n; ci += SMALL_THREADS)                       // k_closed_commit
        if (child_flags[ci] & 2) c.slots[child_slot[ci]].g = v.node_g[v.item_node[ci]];
}

int dx_launch_small_filter_q(cudaStream_t s, const DomainDev& d, const ClosedDev& c, Ctl* ctl, const InstArrays& ia,
                             const uint8_t* node_state, const float* node_g, const uint32_t* popped, const uint32_t* popped_inst,
                             const uint32_t* pop_off, uint8_t* pop_solved, uint32_t* child_slot, uint32_t* child_next,
                             uint8_t* child_flags, uint32_t* child_prefix) {
    ItemView v{nullptr, nullptr, popped, node_state, node_g, popped_inst, ia.active, 1, d.SP};
    dx_launch(k_small_filter_q, 1, SMALL_THREADS, 0, s, d, c, ctl, v, ia, pop_off, pop_solved, child_slot, child_next, child_flags,
              child_prefix);
    return 1;
}

int dx_launch_small_filter_v(cudaStream_t s, const DomainDev& d, const ClosedDev& c, Ctl* ctl, const InstArrays& ia,
                             const uint8_t* child_state, const float* child_g, uint32_t* child_slot, uint32_t* child_next,
                             uint8_t* child_flags, uint32_t* child_prefix, const uint32_t* popped, const uint32_t* popped_inst,
                             uint8_t* node_state, float* node_g, uint32_t* node_parent, uint8_t* node_action, uint32_t* new_inst,
                             uint32_t max_nodes, const uint32_t* pop_off, const uint8_t* pop_solved) {
    ItemView v{child_state, child_g, nullptr, node_state, node_g, popped_inst, ia.active, d.A, d.SP};
    ScatterArgs a{child_state, child_g, child_slot, child_flags, child_prefix, popped, popped_inst,
                  node_state, node_g, node_parent, node_action, new_inst, nullptr, nullptr};
    static int v2 = -1;
    if (v2 < 0) { const char* e = getenv("DXB200_FILTER_V2"); v2 = (e && e[0] == '0') ? 0 : 1; }
    if (v2 && d.SP <= 64)                         // (the register copy of the rows holds 4 x 16 bytes)
        dx_launch(k_small_filter_v2, 1, SMALL_THREADS, 0, s, d, c, ctl, v, child_flags, child_prefix, a, max_nodes, ia, pop_off, pop_solved);
    else
        dx_launch(k_small_filter_v, 1, SMALL_THREADS, 0, s, d, c, ctl, v, child_slot, child_next, child_flags, child_prefix, a, max_nodes, ia,
                  pop_off, pop_solved);
    return 1;
}

int dx_launch_scatter_kept(cudaStream_t s, const DomainDev& d, const ClosedDev& c, Ctl* ctl, const uint8_t* child_state,
                           const float* child_g, const uint32_t* child_slot, const uint8_t* child_flags,
                           const uint32_t* child_prefix, const uint32_t* popped, const uint32_t* popped_inst,
                           uint8_t* node_state, float* node_g, uint32_t* node_parent, uint8_t* node_action, uint32_t* new_inst,
                           int max_children, const float* child_h, float* node_h) {
    int blocks = min(dx_div_up(max_children, CL_THREADS), 148 * 8);
    ScatterArgs a{child_state, child_g, child_slot, child_flags, child_prefix, popped, popped_inst,
                  node_state, node_g, node_parent, node_action, new_inst, child_h, node_h};
    k_scatter_kept<<<blocks, CL_THREADS, 0, s>>>(d, c, ctl, a);
    return 1;
}

// Bellman backup of every node expanded this iteration (deepxube/base/pathfinding.py:41-47, called on the popped nodes by
// UpdateHeurVRLKeepGoalABC._get_labels_no_rb, updaters/updater_rl.py:143-158):
//     backup = 0 if the node is solved else min over its children of (transition cost 1.0 + heuristic(child)),
// in float64 like the reference's Python floats.  One record per popped node is appended, in pop order, to the engine's
// backup log (state row, instance id, value); the nodes an instance popped in the iteration it finished are never expanded by
// the reference (pathfinding.py:339) -- their records carry NaN and are dropped when the log is drained, and their children
// do not count as heuristic evaluations.
__global__ void __launch_bounds__(256)
k_bellman_backup(DomainDev d, Ctl* __restrict__ ctl, InstArrays ia, const uint8_t* __restrict__ node_state,
                 const uint32_t* __restrict__ popped, const uint32_t* __restrict__ popped_inst,
                 const uint8_t* __restrict__ pop_solved, const float* __restrict__ child_h, BackupLog bk) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children / (uint32_t)d.A, at = ctl->n_backups;
    if ((unsigned long long)at + n > bk.cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->backup_overflow = 1;     // turned into ctl->error by k_backup_commit
        return;
    }
    const int chunks = d.SP / 16;
    const uint32_t itr = ctl->itr;
    unsigned long long skipped = 0;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t inst = popped_inst[j], id = popped[j];
        const bool act = ia.active[inst] != 0, solved = act && pop_solved[j];
        double v;
        if (!act) {
            v = __longlong_as_double(0x7ff8000000000000ll);
            skipped += (unsigned long long)d.A;
        } else if (solved) {
            v = 0.0;
        } else {
            v = __longlong_as_double(0x7ff0000000000000ll);
            for (int a = 0; a < d.A; ++a) v = fmin(v, 1.0 + (double)child_h[(size_t)j * d.A + a]);
        }
        const size_t r = (size_t)at + j;
        bk.val[r] = v;
        bk.inst[r] = inst;
        bk.solved[r] = solved ? 1 : 0;
        bk.node[r] = id;
        bk.itr[r] = itr;
        if (act) bk.node_rec[id] = (uint32_t)r + 1;          // node -> its record (tree / parent-path passes at drain time)
        for (int a = 0; a < d.A; ++a) bk.child_h[r * d.A + a] = child_h[(size_t)j * d.A + a];
        const uint4* src = reinterpret_cast<const uint4*>(node_state + (size_t)id * d.SP);
        uint4* dst = reinterpret_cast<uint4*>(bk.state + r * d.SP);
        for (int q = 0; q < chunks; ++q) dst[q] = src[q];
    }
    if (skipped) atomicAdd(&ctl->heur_evals, 0ull - skipped);
}

__global__ void k_backup_commit(Ctl* ctl, int A) {
    if (ctl->error) return;
    if (ctl->backup_overflow) { ctl->error |= DX_DEVERR_BACKUPS; return; }
    ctl->backup_at = ctl->n_backups;
    ctl->n_backups += A > 0 ? ctl->n_children / (uint32_t)A : ctl->n_kept;
}

// after the CLOSED filter: which children of this iteration's records became nodes (kept children, ids in push order)
__global__ void __launch_bounds__(256)
k_backup_children(DomainDev d, const Ctl* __restrict__ ctl, const uint8_t* __restrict__ child_flags,
                  const uint32_t* __restrict__ child_prefix, BackupLog bk) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children, base = ctl->node_base;
    const size_t at = (size_t)ctl->backup_at * d.A;
    for (uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n; ci += gridDim.x * blockDim.x)
        bk.child_id[at + ci] = (child_flags[ci] & 1) ? base + child_prefix[ci] : DX_NIL;
}

__device__ __forceinline__ void atomic_min_f64(double* addr, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v) {
        const unsigned long long seen = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (seen == old) break;
        old = seen;
    }
}

// ub_heur_solns (updaters/updater_rl.py:148-152 over Node.upper_bound_parent_path, base/pathfinding.py:49-53): every solved
// expanded node bounds the backup of its ancestors by the path cost down to it (transition costs 1.0).
__global__ void __launch_bounds__(256)
k_backup_ub_paths(const Ctl* __restrict__ ctl, const uint32_t* __restrict__ node_parent, BackupLog bk) {
    const uint32_t n = ctl->n_backups;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (!bk.solved[r]) continue;
        uint32_t id = bk.node[r];
        double dist = 0.0;
        for (uint32_t par = node_parent[id]; par != DX_NIL; par = node_parent[id]) {
            dist += 1.0;
            const uint32_t rr = bk.node_rec[par];
            if (rr) atomic_min_f64(&bk.val[rr - 1], dist);
            id = par;
        }
    }
}

// LHBL (updaters/updater_rl.py:153-155 over Node.tree_backup, base/pathfinding.py:55-64), one search iteration per launch, last
// iteration first (a child is expanded in a later iteration than its parent, so its value is final when the parent reads it):
//     0 if solved, else min_a (1.0 + (child_a was expanded ? tree value of child_a : max(heuristic(child_a), 0)))
__global__ void __launch_bounds__(256)
k_backup_tree(DomainDev d, const Ctl* __restrict__ ctl, BackupLog bk, uint32_t itr) {
    const uint32_t n = ctl->n_backups;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (bk.itr[r] != itr || bk.val[r] != bk.val[r]) continue;
        double v = 0.0;
        if (!bk.solved[r]) {
            v = __longlong_as_double(0x7ff0000000000000ll);
            for (int a = 0; a < d.A; ++a) {
                const uint32_t cid = bk.child_id[(size_t)r * d.A + a];
                const uint32_t rr = cid != DX_NIL ? bk.node_rec[cid] : 0;
                const double cv = rr ? bk.val[rr - 1] : fmax((double)bk.child_h[(size_t)r * d.A + a], 0.0);
                v = fmin(v, 1.0 + cv);
            }
        }
        bk.val[r] = v;
    }
}

// ---- Q* label options, drain time (UpdateHeurQRLKeepGoalABC._get_labels_no_rb, updaters/updater_rl.py:169-189) ----
// node_aux[node]: ub_heur_solns -> the parent-path bound; lhbl -> min over the node's popped edges of (1 + tree value of the child)
__global__ void k_backup_q_fill(const Ctl* __restrict__ ctl, const uint32_t* __restrict__ node_parent, BackupLog bk) {
    const uint32_t n = ctl->n_backups;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        bk.node_aux[bk.node[r]] = inf;
        bk.node_aux[node_parent[bk.node[r]]] = inf;
    }
}

// ub_heur_solns: every popped edge whose node is solved bounds that node (0) and its ancestors (path cost down to it)
__global__ void __launch_bounds__(256)
k_backup_q_ub_paths(const Ctl* __restrict__ ctl, const uint32_t* __restrict__ node_parent, const uint8_t* __restrict__ node_solved,
                    BackupLog bk) {
    const uint32_t n = ctl->n_backups;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        uint32_t id = node_parent[bk.node[r]];
        if (!node_solved[id]) continue;
        double dist = 0.0;
        for (;;) {
            atomic_min_f64(&bk.node_aux[id], dist);
            id = node_parent[id];
            if (id == DX_NIL) break;
            dist += 1.0;
        }
    }
}

// ... then Node.backup_act: 0 if solved else 1.0 + (node_next.backup_val if it was bounded else heuristic(node_next))
__global__ void __launch_bounds__(256)
k_backup_q_ub_labels(const Ctl* __restrict__ ctl, const uint32_t* __restrict__ node_parent, const uint8_t* __restrict__ node_solved,
                     const float* __restrict__ node_h, BackupLog bk) {
    const uint32_t n = ctl->n_backups;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t c = bk.node[r];
        const double ub = bk.node_aux[c];
        bk.val[r] = node_solved[node_parent[c]] ? 0.0 : 1.0 + (ub < inf ? ub : (double)node_h[c]);
    }
}

// lhbl: Node.tree_backup over the popped edges, one search iteration per launch, last first:
//   T(c) = 0 if c is solved, else (c has popped edges ? min over them of (1 + T(child)) : max(heuristic(c), 0));
//   label of the edge (p -> c) = 0 if p is solved else 1 + T(c)
__global__ void __launch_bounds__(256)
k_backup_q_tree(const Ctl* __restrict__ ctl, const uint32_t* __restrict__ node_parent, const uint8_t* __restrict__ node_solved,
                const float* __restrict__ node_h, BackupLog bk, uint32_t itr) {
    const uint32_t n = ctl->n_backups;
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (bk.itr[r] != itr) continue;
        const uint32_t c = bk.node[r], p = node_parent[c];
        const double em = bk.node_aux[c];
        const double t = node_solved[c] ? 0.0 : (em < inf ? em : fmax((double)node_h[c], 0.0));
        bk.val[r] = node_solved[p] ? 0.0 : 1.0 + t;
        atomic_min_f64(&bk.node_aux[p], 1.0 + t);
    }
}

int dx_launch_backup_q_modes(cudaStream_t s, const Ctl* ctl, const uint32_t* node_parent, const uint8_t* node_solved,
                             const float* node_h, const BackupLog& bk, int mode, uint32_t it_first, uint32_t it_last) {
    int launches = 1;
    k_backup_q_fill<<<148 * 4, 256, 0, s>>>(ctl, node_parent, bk);
    if (mode == 1) {
        k_backup_q_ub_paths<<<148 * 4, 256, 0, s>>>(ctl, node_parent, node_solved, bk);
        k_backup_q_ub_labels<<<148 * 4, 256, 0, s>>>(ctl, node_parent, node_solved, node_h, bk);
        launches += 2;
    } else {
        for (uint32_t it = it_last + 1; it-- > it_first;) {
            k_backup_q_tree<<<148 * 4, 256, 0, s>>>(ctl, node_parent, node_solved, node_h, bk, it);
            launches += 1;
        }
    }
    return launches;
}

// drain: rows without their padding, contiguous, so that ONE copy brings them to the host
__global__ void k_backup_pack_rows(DomainDev d, const Ctl* __restrict__ ctl, BackupLog bk) {
    const size_t total = (size_t)ctl->n_backups * d.S;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / d.S;
        bk.pack[i] = bk.state[r * d.SP + (i - r * d.S)];
    }
}
int dx_launch_backup_pack_rows(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const BackupLog& bk) {
    k_backup_pack_rows<<<148 * 8, 256, 0, s>>>(d, ctl, bk);
    return 1;
}

// the drained records' nodes lose their record
__global__ void k_backup_clear_map(const Ctl* __restrict__ ctl, BackupLog bk) {
    const uint32_t n = ctl->n_backups;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t id = bk.node[r];
        if (bk.node_rec[id] == r + 1) bk.node_rec[id] = 0;
    }
}

int dx_launch_backup_children(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const uint8_t* child_flags,
                              const uint32_t* child_prefix, const BackupLog& bk, int max_children) {
    k_backup_children<<<min(dx_div_up(max_children, 256), 148 * 8), 256, 0, s>>>(d, ctl, child_flags, child_prefix, bk);
    return 1;
}
int dx_launch_backup_ub_paths(cudaStream_t s, const Ctl* ctl, const uint32_t* node_parent, const BackupLog& bk) {
    k_backup_ub_paths<<<148 * 4, 256, 0, s>>>(ctl, node_parent, bk);
    return 1;
}
int dx_launch_backup_tree(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const BackupLog& bk, uint32_t itr) {
    k_backup_tree<<<148 * 4, 256, 0, s>>>(d, ctl, bk, itr);
    return 1;
}
int dx_launch_backup_clear_map(cudaStream_t s, const Ctl* ctl, const BackupLog& bk) {
    k_backup_clear_map<<<148 * 4, 256, 0, s>>>(ctl, bk);
    return 1;
}

// Q*: remember is_solved of the nodes popped this iteration (PathFind.set_is_solved, base/pathfinding.py:353): the label of
// an edge popped in a LATER iteration needs it.
__global__ void k_mark_solved(const Ctl* __restrict__ ctl, const uint32_t* __restrict__ popped,
                              const uint32_t* __restrict__ pop_off, const uint8_t* __restrict__ pop_solved,
                              uint8_t* __restrict__ node_solved) {
    if (ctl->error) return;
    const uint32_t n = pop_off[ctl->n_inst];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        node_solved[popped[j]] = pop_solved[j];
}

// Q* labels (UpdateHeurQRLKeepGoalABC._get_labels_no_rb, updaters/updater_rl.py:169-189, over Node.backup_act,
// base/pathfinding.py:66-77): for every edge (node, action) popped this iteration,
//     0 if node is solved else transition cost 1.0 + heuristic(node_next), heuristic(node_next) = min_a q(node_next, a).
// node_next is the node generated for the edge this iteration (ids node_base + k, in pop order); its backup_val is still
// +inf when the reference reads it (it can only be set by edges popped later), hence always the heuristic branch.
__global__ void __launch_bounds__(256)
k_bellman_backup_q(DomainDev d, Ctl* __restrict__ ctl, const uint8_t* __restrict__ node_state,
                   const uint32_t* __restrict__ node_parent, const uint8_t* __restrict__ node_action,
                   const float* __restrict__ node_h, const uint8_t* __restrict__ node_solved,
                   const uint32_t* __restrict__ edge_inst, BackupLog bk) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_kept, base = ctl->node_base, at = ctl->n_backups;
    if ((unsigned long long)at + n > bk.cap) {
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->backup_overflow = 1;
        return;
    }
    const int chunks = d.SP / 16;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t id = base + k, par = node_parent[id];
        bk.val[at + k] = node_solved[par] ? 0.0 : 1.0 + (double)node_h[id];
        bk.inst[at + k] = edge_inst[k];
        bk.act[at + k] = node_action[id];
        bk.node[at + k] = id;                       // the node generated for the edge
        bk.itr[at + k] = ctl->itr;
        const uint4* src = reinterpret_cast<const uint4*>(node_state + (size_t)par * d.SP);
        uint4* dst = reinterpret_cast<uint4*>(bk.state + (size_t)(at + k) * d.SP);
        for (int q = 0; q < chunks; ++q) dst[q] = src[q];
    }
}

int dx_launch_mark_solved(cudaStream_t s, const Ctl* ctl, const uint32_t* popped, const uint32_t* pop_off,
                          const uint8_t* pop_solved, uint8_t* node_solved, int max_pop) {
    k_mark_solved<<<min(dx_div_up(max_pop, 256), 148 * 4), 256, 0, s>>>(ctl, popped, pop_off, pop_solved, node_solved);
    return 1;
}

int dx_launch_bellman_backup_q(cudaStream_t s, const DomainDev& d, Ctl* ctl, const uint8_t* node_state, const uint32_t* node_parent,
                               const uint8_t* node_action, const float* node_h, const uint8_t* node_solved,
                               const uint32_t* edge_inst, const BackupLog& bk, int max_pop) {
    k_bellman_backup_q<<<min(dx_div_up(max_pop, 256), 148 * 4), 256, 0, s>>>(d, ctl, node_state, node_parent, node_action, node_h,
                                                                            node_solved, edge_inst, bk);
    k_backup_commit<<<1, 1, 0, s>>>(ctl, 0);
    return 2;
}

int dx_launch_bellman_backup(cudaStream_t s, const DomainDev& d, Ctl* ctl, const InstArrays& ia, const uint8_t* node_state,
                             const uint32_t* popped, const uint32_t* popped_inst, const uint8_t* pop_solved, const float* child_h,
                             const BackupLog& bk, int max_pop) {
    k_bellman_backup<<<min(dx_div_up(max_pop, 256), 148 * 4), 256, 0, s>>>(d, ctl, ia, node_state, popped, popped_inst,
                                                                                pop_solved, child_h, bk);
    k_backup_commit<<<1, 1, 0, s>>>(ctl, d.A);
    return 2;
}

// Q*: popped nodes that pass the filter only need CLOSED committed (they already are nodes).
__global__ void k_closed_commit(ClosedDev c, const Ctl* __restrict__ ctl, const float* __restrict__ node_g,
                                const uint32_t* __restrict__ item_node, const uint32_t* __restrict__ child_slot,
                                const uint8_t* __restrict__ child_flags) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children;
    for (uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n; ci += gridDim.x * blockDim.x)
        if (child_flags[ci] & 2) c.slots[child_slot[ci]].g = node_g[item_node[ci]];
}

int dx_launch_closed_commit(cudaStream_t s, const ClosedDev& c, const Ctl* ctl, const float* node_g, const uint32_t* item_node,
                            const uint32_t* child_slot, const uint8_t* child_flags, int max_items) {
    int blocks = min(dx_div_up(max_items, 256), 148 * 8);
    k_closed_commit<<<blocks, 256, 0, s>>>(c, ctl, node_g, item_node, child_slot, child_flags);
    return 1;
}

// Beam search has no CLOSED list (InstanceNodeBeam.filter_expanded_nodes returns every child, beam_search.py:108-109):
// every child of an active instance is kept (flag 1, never the "commits CLOSED" flag 2).
__global__ void k_beam_keep(const Ctl* __restrict__ ctl, const float* __restrict__ child_g, uint8_t* __restrict__ child_flags) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children;
    for (uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n; ci += gridDim.x * blockDim.x) {
        const float g = child_g[ci];
        child_flags[ci] = (g == g) ? 1 : 0;          // NaN g = child of a finished instance
    }
}

int dx_launch_beam_keep(cudaStream_t s, const Ctl* ctl, const float* child_g, uint8_t* child_flags, int max_children) {
    k_beam_keep<<<min(dx_div_up(max_children, 256), 148 * 8), 256, 0, s>>>(ctl, child_g, child_flags);
    return 1;
}

// beam_q: every popped node of an active instance keeps its edges (InstanceEdgeBeam.filter_popped_nodes, beam_search.py:122-123)
__global__ void k_beam_keep_q(const Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ popped_inst,
                              uint8_t* __restrict__ flags) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x)
        flags[j] = ia.active[popped_inst[j]] ? 1 : 0;
}

int dx_launch_beam_keep_q(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const uint32_t* popped_inst, uint8_t* flags, int max_items) {
    k_beam_keep_q<<<min(dx_div_up(max_items, 256), 148 * 8), 256, 0, s>>>(ctl, ia, popped_inst, flags);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// Hindsight relabelling (UpdateHER._get_her_goals, deepxube/base/updater.py:485-530): for an instance that did not finish with a
// solution the relabelled goal is sampled from the DEEPEST node that has children (an expanded node; the root when nothing deeper
// was expanded); the reference walks `root.get_all_descendants()` (breadth first, children in action order) and keeps the first
// node of the largest depth -- among equally deep nodes, the one whose action path from the root is lexicographically smallest.
// Candidates come from the backup log: A* records are the expanded nodes; a Q* record is a popped edge, whose node has a child.
//   her_depth: depth of every record's candidate (parent hops to the root) and the per-instance maximum
//   her_pick : the first of the deepest in that walk's order (two equally deep nodes are compared below their lowest common ancestor)
__global__ void k_her_depth(const Ctl* __restrict__ ctl, BackupLog bk, const uint32_t* __restrict__ node_parent, int is_q,
                            uint32_t* __restrict__ rec_cand, uint32_t* __restrict__ rec_depth, uint32_t* __restrict__ best_depth) {
    const uint32_t n = ctl->n_backups;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        if (bk.val[r] != bk.val[r]) { rec_cand[r] = DX_NIL; continue; }          // popped by an instance that had finished
        uint32_t c = bk.node[r];
        if (is_q) c = node_parent[c];                                            // the edge's node (the generated node's parent)
        uint32_t depth = 0;
        for (uint32_t x = c; node_parent[x] != DX_NIL; x = node_parent[x]) ++depth;
        rec_cand[r] = c;
        rec_depth[r] = depth;
        atomicMax(&best_depth[bk.inst[r]], depth);
    }
}

// One thread per record of the largest depth: a compare-and-swap tournament on the instance's slot.  `earlier(c, cur)` orders two
// equally deep nodes the way the reference's breadth-first walk meets them: below their lowest common ancestor, by the order in
// which the LCA's two children entered its edge_dict -- action order for A* (children are generated together, ids in action
// order), pop order of the edges for Q* -- which is the order of the children's node ids in both cases.
__global__ void k_her_pick(const Ctl* __restrict__ ctl, BackupLog bk, const uint32_t* __restrict__ node_parent,
                           const uint32_t* __restrict__ rec_cand, const uint32_t* __restrict__ rec_depth,
                           const uint32_t* __restrict__ best_depth, uint32_t* __restrict__ deep_node) {
    const uint32_t n = ctl->n_backups;
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        const uint32_t c = rec_cand[r];
        if (c == DX_NIL) continue;
        const uint32_t i = bk.inst[r];
        if (rec_depth[r] != best_depth[i] || rec_depth[r] == 0) continue;
        uint32_t cur = *((volatile uint32_t*)&deep_node[i]);
        for (;;) {
            if (cur == c) break;
            if (cur != DX_NIL) {
                uint32_t a = c, b = cur, ap = c, bp = cur;                       // equal depths: the walks meet at the LCA
                while (a != b) { ap = a; bp = b; a = node_parent[a]; b = node_parent[b]; }
                if (ap > bp) break;                                              // the holder comes first
            }
            const uint32_t seen = atomicCAS(&deep_node[i], cur, c);
            if (seen == cur) break;
            cur = seen;
        }
    }
}

__global__ void k_her_rows(DomainDev d, const Ctl* __restrict__ ctl, const uint8_t* __restrict__ node_state,
                           const uint32_t* __restrict__ deep_node, const uint32_t* __restrict__ root, uint8_t* __restrict__ out_rows) {
    const uint32_t I = ctl->n_inst;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < I * (uint32_t)d.S; t += gridDim.x * blockDim.x) {
        const uint32_t i = t / d.S, b = t - i * d.S;
        const uint32_t nd = deep_node[i] == DX_NIL ? root[i] : deep_node[i];       // nothing deeper than the root was expanded
        out_rows[t] = node_state[(size_t)nd * d.SP + b];
    }
}

int dx_launch_her_deepest(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const BackupLog& bk, const uint32_t* node_parent,
                          const uint8_t* node_state, int is_q, const uint32_t* root, uint32_t* rec_cand,
                          uint32_t* rec_depth, uint32_t* best_depth, uint32_t* deep_node, uint8_t* out_rows, int max_inst) {
    cudaMemsetAsync(best_depth, 0, (size_t)max_inst * sizeof(uint32_t), s);
    k_her_depth<<<148 * 4, 256, 0, s>>>(ctl, bk, node_parent, is_q, rec_cand, rec_depth, best_depth);
    cudaMemsetAsync(deep_node, 0xff, (size_t)max_inst * sizeof(uint32_t), s);
    k_her_pick<<<148 * 4, 256, 0, s>>>(ctl, bk, node_parent, rec_cand, rec_depth, best_depth, deep_node);
    k_her_rows<<<min(dx_div_up((long long)max_inst * d.S, 256), 148 * 4), 256, 0, s>>>(d, ctl, node_state, deep_node, root, out_rows);
    return 3;
}
