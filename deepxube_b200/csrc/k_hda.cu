// HDA*-style hash-distributed batch weighted A* (BASELINE config 5): one search instance spread
// over W ranks.  owner(state) = hash(row) mod W owns the state's CLOSED entry, node and OPEN entry.  Per
// iteration every rank pops its local top-(B/W), expands them, and routes each child record to its owner
// (all-to-all, done by the host layer with NCCL / an in-process exchange); the owner filters the received
// children against its CLOSED shard, evaluates the heuristic for the survivors and pushes them.
// This file: bucketing of the children by owner, record pack / unpack, and the uniform end-of-iteration
// bookkeeping that replaces k_end_iter (lb / ub / finished() are global quantities here).
//
// Record layout (rec_bytes = SP + 16): [row SP bytes][g f32][parent global id u32][action u32][pad u32].
// Global node id = (rank << 28) | local id.
#include "dx_internal.cuh"

__device__ __forceinline__ uint32_t hda_owner(const uint4* row, int chunks, uint32_t world) {
    return (uint32_t)((dx_mix64(dx_hash_row(row, chunks)) >> 40) % world);
}

// sort key = owner rank of each child (stable radix sort on it groups the children by destination)
__global__ void k_hda_owner_keys(DomainDev d, Ctl* __restrict__ ctl, const uint8_t* __restrict__ child_state,
                                 const float* __restrict__ child_g, SortBufs sb, uint32_t world, uint32_t* __restrict__ counts) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->sort_n = n;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        const uint32_t o = hda_owner(reinterpret_cast<const uint4*>(child_state + (size_t)c * d.SP), d.SP / 16, world);
        sb.key[0][c] = o;
        sb.inst[0][c] = 0;
        sb.val[0][c] = c;
        atomicAdd(&counts[o], 1u);
    }
}

__global__ void k_hda_pack(DomainDev d, const Ctl* __restrict__ ctl, SortBufs sb, const uint8_t* __restrict__ child_state,
                           const float* __restrict__ child_g, const uint32_t* __restrict__ popped, uint32_t rank,
                           uint8_t* __restrict__ send) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children;
    const uint32_t* __restrict__ order = sb.val[ctl->sort_final];
    const int words = d.SP / 4, rw = words + 4;
    const size_t total = (size_t)n * rw;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / rw), w = (uint32_t)(t % rw);
        const uint32_t c = order[r];
        uint32_t v;
        if (w < (uint32_t)words) v = reinterpret_cast<const uint32_t*>(child_state + (size_t)c * d.SP)[w];
        else if (w == (uint32_t)words) v = __float_as_uint(child_g[c]);
        else if (w == (uint32_t)words + 1) v = (rank << 28) | popped[c / d.A];
        else if (w == (uint32_t)words + 2) v = c % d.A;
        else v = 0u;
        reinterpret_cast<uint32_t*>(send)[t] = v;
    }
}

__global__ void k_hda_unpack(DomainDev d, Ctl* __restrict__ ctl, const uint8_t* __restrict__ recv, uint32_t n,
                             uint8_t* __restrict__ child_state, float* __restrict__ child_g,
                             uint32_t* __restrict__ child_parent, uint8_t* __restrict__ child_action) {
    if (ctl->error) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) ctl->n_children = n;
    const int words = d.SP / 4, rw = words + 4;
    const size_t total = (size_t)n * rw;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / rw), w = (uint32_t)(t % rw);
        const uint32_t v = reinterpret_cast<const uint32_t*>(recv)[t];
        if (w < (uint32_t)words) reinterpret_cast<uint32_t*>(child_state + (size_t)r * d.SP)[w] = v;
        else if (w == (uint32_t)words) child_g[r] = __uint_as_float(v);
        else if (w == (uint32_t)words + 1) child_parent[r] = v;
        else if (w == (uint32_t)words + 2) child_action[r] = (uint8_t)v;
    }
}

// kept received children -> node store (ids in receive order = push order on this rank); CLOSED commit.
__global__ void k_scatter_kept_hda(DomainDev d, ClosedDev c, Ctl* __restrict__ ctl, const uint8_t* __restrict__ child_state,
                                   const float* __restrict__ child_g, const uint32_t* __restrict__ child_slot,
                                   const uint8_t* __restrict__ child_flags, const uint32_t* __restrict__ child_prefix,
                                   const uint32_t* __restrict__ child_parent, const uint8_t* __restrict__ child_action,
                                   uint8_t* __restrict__ node_state, float* __restrict__ node_g,
                                   uint32_t* __restrict__ node_parent, uint8_t* __restrict__ node_action,
                                   uint32_t* __restrict__ new_inst) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_children, base = ctl->node_base;
    const int chunks = d.SP / 16;
    for (uint32_t ci = blockIdx.x * blockDim.x + threadIdx.x; ci < n; ci += gridDim.x * blockDim.x) {
        const uint8_t f = child_flags[ci];
        if (!(f & 1)) continue;
        const uint32_t k = child_prefix[ci], id = base + k;
        const uint4* src = reinterpret_cast<const uint4*>(child_state + (size_t)ci * d.SP);
        uint4* dst = reinterpret_cast<uint4*>(node_state + (size_t)id * d.SP);
        for (int q = 0; q < chunks; ++q) dst[q] = src[q];
        const float g = child_g[ci];
        node_g[id] = g;
        node_parent[id] = child_parent[ci];
        node_action[id] = child_action[ci];
        new_inst[k] = 0;
        if (f & 2) {
            const uint32_t i = child_slot[ci];
            c.slot_g[i] = g;
            const unsigned long long key = c.slot_key[i];
            if ((uint32_t)key & DX_PENDING) c.slot_key[i] = (key & 0xFFFFFFFF00000000ull) | id;
        }
    }
}

// Uniform end-of-iteration bookkeeping on every rank from globally reduced quantities
// (deepxube/pathfinding/graph_search.py:43-46, 67-69 applied to the distributed instance).
__global__ void k_hda_finish(Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ pop_off_cur,
                             const uint32_t* __restrict__ pop_off_next, int A, unsigned long long k_total,
                             unsigned long long f_first_min_bits, uint32_t ub_g_bits, int has_ub,
                             unsigned long long gen_add) {
    if (ctl->error) return;
    double lb = ia.lb[0];
    if (k_total > 0) lb = fmax(__longlong_as_double((long long)f_first_min_bits), lb);
    ia.lb[0] = lb;
    ia.gen[0] += gen_add;
    ctl->gen_total += gen_add;
    ia.pop_seq_base[0] += pop_off_cur[1] - pop_off_cur[0];
    ia.itr[0] += 1;
    bool fin = false;
    if (has_ub) {
        const unsigned long long gk = ((unsigned long long)ub_g_bits << 32) | 0xFFFFFFFFull;
        if (gk < ia.ub_key[0]) ia.ub_key[0] = gk;
        const double ub = (double)__uint_as_float(ub_g_bits);
        fin = lb >= __dmul_rn(ia.weight[0], ub);
    }
    fin = fin || (k_total == 0);
    ia.active[0] = fin ? 0 : 1;
    ctl->n_unfinished = fin ? 0u : 1u;
    ctl->n_pop = pop_off_next[1];
    ctl->n_children = pop_off_next[1] * (uint32_t)A;
    ctl->itr += 1;
    ctl->epoch += 1;
    ctl->node_base = ctl->node_count;
}

// non-owner ranks hold the instance without a root in their popped list
__global__ void k_hda_drop_root(Ctl* ctl, uint32_t* pop_off) {
    ctl->n_pop = 0;
    ctl->n_children = 0;
    pop_off[0] = 0;
    pop_off[1] = 0;
}

int dx_launch_hda_bucket(cudaStream_t s, const DomainDev& d, Ctl* ctl, const SortBufs& sb, const ScanTemp& t,
                         const uint8_t* child_state, const float* child_g, const uint32_t* popped, uint32_t rank, uint32_t world,
                         uint32_t* counts, uint8_t* send, int max_children) {
    int launches = 0;
    const int blocks = min(dx_div_up(max_children, 256), 148 * 8);
    k_hda_owner_keys<<<blocks, 256, 0, s>>>(d, ctl, child_state, child_g, sb, world, counts);
    launches += 1 + dx_launch_sort(s, ctl, sb, t, max_children);
    const int pblocks = (int)min((long long)dx_div_up((long long)max_children * (d.SP / 4 + 4), 256), (long long)148 * 16);
    k_hda_pack<<<pblocks, 256, 0, s>>>(d, ctl, sb, child_state, child_g, popped, rank, send);
    return launches + 1;
}

int dx_launch_hda_unpack(cudaStream_t s, const DomainDev& d, Ctl* ctl, const uint8_t* recv, uint32_t n, uint8_t* child_state,
                         float* child_g, uint32_t* child_parent, uint8_t* child_action) {
    const int blocks = (int)min((long long)dx_div_up((long long)max(n, 1u) * (d.SP / 4 + 4), 256), (long long)148 * 16);
    k_hda_unpack<<<blocks, 256, 0, s>>>(d, ctl, recv, n, child_state, child_g, child_parent, child_action);
    return 1;
}

int dx_launch_scatter_kept_hda(cudaStream_t s, const DomainDev& d, const ClosedDev& c, Ctl* ctl, const uint8_t* child_state,
                               const float* child_g, const uint32_t* child_slot, const uint8_t* child_flags,
                               const uint32_t* child_prefix, const uint32_t* child_parent, const uint8_t* child_action,
                               uint8_t* node_state, float* node_g, uint32_t* node_parent, uint8_t* node_action,
                               uint32_t* new_inst, int max_children) {
    const int blocks = min(dx_div_up(max_children, 256), 148 * 8);
    k_scatter_kept_hda<<<blocks, 256, 0, s>>>(d, c, ctl, child_state, child_g, child_slot, child_flags, child_prefix, child_parent,
                                              child_action, node_state, node_g, node_parent, node_action, new_inst);
    return 1;
}

int dx_launch_hda_finish(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* pop_off_cur, const uint32_t* pop_off_next,
                         int A, unsigned long long k_total, unsigned long long f_first_min_bits, uint32_t ub_g_bits, int has_ub,
                         unsigned long long gen_add) {
    k_hda_finish<<<1, 1, 0, s>>>(ctl, ia, pop_off_cur, pop_off_next, A, k_total, f_first_min_bits, ub_g_bits, has_ub, gen_add);
    return 1;
}

int dx_launch_hda_drop_root(cudaStream_t s, Ctl* ctl, uint32_t* pop_off) {
    k_hda_drop_root<<<1, 1, 0, s>>>(ctl, pop_off);
    return 1;
}

uint32_t dx_hda_owner_host(const uint8_t* row, int SP, uint32_t world) {
    return (uint32_t)((dx_mix64(dx_hash_row(reinterpret_cast<const uint4*>(row), SP / 16)) >> 40) % world);
}
