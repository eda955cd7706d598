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
            const unsigned long long key = c.slots[i].key;       // (read before the slot's sector is written)
            c.slots[i].g = g;
            if ((uint32_t)key & DX_PENDING) c.slots[i].key = (key & 0xFFFFFFFF00000000ull) | id;
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

// ---------------------------------------------------------------------------------------------
// Device-resident exchange (no host synchronisation inside an iteration).  The send buffer holds W fixed-size
// segments, one per destination rank: [u32 count, 12 B pad][seg_records records]; an all-to-all with EQUAL splits
// moves them, so no counts have to reach the host first; the receiver reads the counts from the segment headers.
// Received children are numbered source-major, in the sender's order (the sender's stable sort by owner keeps its
// child order), exactly the order the variable-size all-to-all of the host-driven version produced.

// after the stable sort by owner: records -> segments; aux[0] = children this rank generated.
// dst.seg[o] is where the segment for rank o goes: a slice of the local send buffer (NCCL all-to-all follows), or --
// peer-memory exchange -- slot `rank` of rank o's receive buffer, written straight over NVLink by this kernel.
__global__ void k_hda_pack_seg(DomainDev d, Ctl* __restrict__ ctl, SortBufs sb, const uint8_t* __restrict__ child_state,
                               const float* __restrict__ child_g, const uint32_t* __restrict__ popped, uint32_t rank, uint32_t world,
                               const uint32_t* __restrict__ counts, HdaDst dst, uint32_t seg_records,
                               uint32_t* __restrict__ aux) {
    __shared__ uint32_t s_pref[17];
    if (ctl->error) {
        // A rank in the error state still publishes EMPTY segments: without fresh headers its peers would re-read the
        // counts and records this rank wrote in the previous iteration.  The error itself travels in the per-rank
        // scalars (k_hda_locals -> k_hda_reduce), so every rank raises in the same iteration.
        if (blockIdx.x == 0 && threadIdx.x < world) *reinterpret_cast<uint4*>(dst.seg[threadIdx.x]) = make_uint4(0u, 0u, 0u, 0u);
        if (blockIdx.x == 0 && threadIdx.x == 0) aux[0] = 0u;
        return;
    }
    const uint32_t n = ctl->n_children;
    const int words = d.SP / 4, rw = words + 4;
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t o = 0; o < world; ++o) { s_pref[o] = acc; acc += counts[o]; }
        s_pref[world] = acc;
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < world) {
        uint32_t c = counts[threadIdx.x];
        if (c > seg_records) { atomicOr(&ctl->error, DX_DEVERR_HDA); c = seg_records; }
        *reinterpret_cast<uint4*>(dst.seg[threadIdx.x]) = make_uint4(c, 0u, 0u, 0u);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) aux[0] = n;
    const unsigned long long* __restrict__ keys = sb.key[ctl->sort_final];
    const uint32_t* __restrict__ order = sb.val[ctl->sort_final];
    const size_t total = (size_t)n * rw;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / rw), w = (uint32_t)(t % rw);
        const uint32_t o = (uint32_t)keys[r];
        const uint32_t pos = r - s_pref[o];
        if (pos >= seg_records) continue;
        const uint32_t c = order[r];
        uint32_t v;
        if (w < (uint32_t)words) v = reinterpret_cast<const uint32_t*>(child_state + (size_t)c * d.SP)[w];
        else if (w == (uint32_t)words) v = __float_as_uint(child_g[c]);
        else if (w == (uint32_t)words + 1) v = (rank << 28) | popped[c / d.A];
        else if (w == (uint32_t)words + 2) v = c % d.A;
        else v = 0u;
        reinterpret_cast<uint32_t*>(dst.seg[o] + 16 + (size_t)pos * (rw * 4))[w] = v;
    }
}

// segment headers of the received buffer -> per-source offsets, ctl->n_children = records received
__global__ void k_hda_recv_plan(Ctl* __restrict__ ctl, const uint8_t* __restrict__ recv, uint32_t world, size_t seg_bytes,
                                uint32_t max_children, uint32_t* __restrict__ src_off, uint32_t* __restrict__ aux) {
    if (ctl->error) return;
    uint32_t acc = 0;
    for (uint32_t s = 0; s < world; ++s) {
        src_off[s] = acc;
        acc += *reinterpret_cast<const uint32_t*>(recv + (size_t)s * seg_bytes);
    }
    if (acc > max_children) { atomicOr(&ctl->error, DX_DEVERR_HDA); acc = 0; }
    src_off[world] = acc;
    ctl->n_children = acc;
    aux[1] = acc;
}

__global__ void k_hda_unpack_seg(DomainDev d, const Ctl* __restrict__ ctl, const uint8_t* __restrict__ recv, uint32_t world,
                                 size_t seg_bytes, const uint32_t* __restrict__ src_off, uint8_t* __restrict__ child_state,
                                 float* __restrict__ child_g, uint32_t* __restrict__ child_parent, uint8_t* __restrict__ child_action) {
    __shared__ uint32_t s_off[17];
    if (ctl->error) return;
    if (threadIdx.x <= world) s_off[threadIdx.x] = src_off[threadIdx.x];
    __syncthreads();
    const uint32_t n = ctl->n_children;
    const int words = d.SP / 4, rw = words + 4;
    const size_t total = (size_t)n * rw;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / rw), w = (uint32_t)(t % rw);
        uint32_t s = 0;
        while (r >= s_off[s + 1]) ++s;
        const uint32_t v = reinterpret_cast<const uint32_t*>(recv + (size_t)s * seg_bytes + 16 + (size_t)(r - s_off[s]) * (rw * 4))[w];
        if (w < (uint32_t)words) reinterpret_cast<uint32_t*>(child_state + (size_t)r * d.SP)[w] = v;
        else if (w == (uint32_t)words) child_g[r] = __uint_as_float(v);
        else if (w == (uint32_t)words + 1) child_parent[r] = v;
        else if (w == (uint32_t)words + 2) child_action[r] = (uint8_t)v;
    }
}

// this rank's per-iteration scalars, as DX_HDA_LOCALS doubles: {k_popped, f of the first popped, ub, goal node (local id or -1),
// children generated, device error flags, children received, received children that passed the CLOSED filter}
__global__ void k_hda_locals(const Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ aux, double* __restrict__ loc) {
    const uint32_t k = ctl->error ? 0u : ia.k_pop[0];
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    loc[0] = (double)k;
    loc[1] = k ? dx_key_to_f(ia.f_first[0]) : inf;
    const unsigned long long ubk = ia.ub_key[0];
    loc[2] = ubk == DX_EMPTY64 ? inf : (double)__uint_as_float((uint32_t)(ubk >> 32));
    const uint32_t goal = ia.goal_node[0];
    loc[3] = goal == DX_NIL ? -1.0 : (double)goal;
    loc[4] = ctl->error ? 0.0 : (double)aux[0];
    loc[5] = (double)ctl->error;
    loc[6] = ctl->error ? 0.0 : (double)aux[1];
    loc[7] = ctl->error ? 0.0 : (double)ctl->n_kept;
}

// every rank, from the all-gathered scalars: the bookkeeping of k_hda_finish plus a summary for the host:
// {finished, iterations, nodes generated, ub, goal rank, goal local id, device error flags OR-ed over ALL ranks,
//  max over ranks of the children received this iteration (load imbalance of hash mod W)}
// A device error on any rank freezes every rank in the same iteration (DX_DEVERR_PEER marks "a peer failed").
__global__ void k_hda_reduce(Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ pop_off_cur,
                             const uint32_t* __restrict__ pop_off_next, int A, int world, const double* __restrict__ all,
                             double* __restrict__ summary) {
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double k_total = 0.0, f_min = inf, children = 0.0, ub = inf, g_rank = -1.0, g_id = -1.0, recv_max = 0.0;
    uint32_t err_all = 0u;
    for (int r = 0; r < world; ++r) {
        const double* l = all + DX_HDA_LOCALS * r;
        err_all |= (uint32_t)l[5];
        recv_max = fmax(recv_max, l[6]);
        k_total += l[0];
        f_min = fmin(f_min, l[1]);
        children += l[4];
        if (l[3] >= 0.0 && l[2] < ub) { ub = l[2]; g_rank = (double)r; g_id = l[3]; }     // first rank wins ties, as min over (ub, rank)
    }
    if (err_all && !ctl->error) ctl->error = DX_DEVERR_PEER;      // a peer failed: stop here too (its segments are empty from now on)
    if (!ctl->error && ia.active[0]) {          // a finished search stays frozen
        double lb = ia.lb[0];
        if (k_total > 0.0) lb = fmax(f_min, lb);
        ia.lb[0] = lb;
        ia.gen[0] += (unsigned long long)children;
        ctl->gen_total += (unsigned long long)children;
        ia.pop_seq_base[0] += pop_off_cur[1] - pop_off_cur[0];
        ia.itr[0] += 1;
        bool fin = false;
        if (g_rank >= 0.0) {
            const uint32_t ub_bits = __float_as_uint((float)ub);
            const unsigned long long gk = ((unsigned long long)ub_bits << 32) | 0xFFFFFFFFull;
            if (gk < ia.ub_key[0]) ia.ub_key[0] = gk;
            fin = lb >= __dmul_rn(ia.weight[0], ub);
        }
        fin = fin || (k_total == 0.0);
        ia.active[0] = fin ? 0 : 1;
        ctl->n_unfinished = fin ? 0u : 1u;
        ctl->n_pop = fin ? 0u : pop_off_next[1];
        ctl->n_children = ctl->n_pop * (uint32_t)A;
        ctl->itr += 1;
        ctl->epoch += 1;
        ctl->node_base = ctl->node_count;
    }
    summary[0] = ia.active[0] ? 0.0 : 1.0;
    summary[1] = (double)ia.itr[0];
    summary[2] = (double)ia.gen[0];
    summary[3] = ub;
    summary[4] = g_rank;
    summary[5] = g_id;
    summary[6] = (double)(err_all | ctl->error);
    summary[7] = recv_max;
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
    SortBufs sb_all = sb;
    sb_all.key_cut = 0;                       // (owner keys are small integers: every bit goes through the digit passes)
    launches += 1 + dx_launch_sort(s, ctl, sb_all, t, max_children);
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

int dx_launch_hda_send(cudaStream_t s, const DomainDev& d, Ctl* ctl, const SortBufs& sb, const ScanTemp& t, const uint8_t* child_state,
                       const float* child_g, const uint32_t* popped, uint32_t rank, uint32_t world, uint32_t* counts, const HdaDst& dst,
                       uint32_t seg_records, uint32_t* aux, int max_children) {
    int launches = 0;
    const int blocks = min(dx_div_up(max_children, 256), 148 * 8);
    k_hda_owner_keys<<<blocks, 256, 0, s>>>(d, ctl, child_state, child_g, sb, world, counts);
    SortBufs sb_all = sb;
    sb_all.key_cut = 0;                       // (owner keys are small integers: every bit goes through the digit passes)
    launches += 1 + dx_launch_sort(s, ctl, sb_all, t, max_children);
    const int pblocks = (int)min((long long)dx_div_up((long long)max_children * (d.SP / 4 + 4), 256), (long long)148 * 16);
    k_hda_pack_seg<<<pblocks, 256, 0, s>>>(d, ctl, sb, child_state, child_g, popped, rank, world, counts, dst, seg_records, aux);
    return launches + 1;
}

int dx_launch_hda_recv(cudaStream_t s, const DomainDev& d, Ctl* ctl, const uint8_t* recv, uint32_t world, uint32_t seg_records,
                       uint32_t max_children, uint32_t* src_off, uint8_t* child_state, float* child_g, uint32_t* child_parent,
                       uint8_t* child_action, uint32_t* aux) {
    const size_t seg_bytes = 16 + (size_t)seg_records * (size_t)(d.SP + 16);
    k_hda_recv_plan<<<1, 1, 0, s>>>(ctl, recv, world, seg_bytes, max_children, src_off, aux);
    const int blocks = (int)min((long long)dx_div_up((long long)max_children * (d.SP / 4 + 4), 256), (long long)148 * 16);
    k_hda_unpack_seg<<<blocks, 256, 0, s>>>(d, ctl, recv, world, seg_bytes, src_off, child_state, child_g, child_parent, child_action);
    return 2;
}

int dx_launch_hda_locals(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const uint32_t* aux, double* loc) {
    k_hda_locals<<<1, 1, 0, s>>>(ctl, ia, aux, loc);
    return 1;
}

int dx_launch_hda_reduce(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* pop_off_cur, const uint32_t* pop_off_next,
                         int A, int world, const double* all, double* summary) {
    k_hda_reduce<<<1, 1, 0, s>>>(ctl, ia, pop_off_cur, pop_off_next, A, world, all, summary);
    return 1;
}

int dx_launch_hda_drop_root(cudaStream_t s, Ctl* ctl, uint32_t* pop_off) {
    k_hda_drop_root<<<1, 1, 0, s>>>(ctl, pop_off);
    return 1;
}

uint32_t dx_hda_owner_host(const uint8_t* row, int SP, uint32_t world) {
    return (uint32_t)((dx_mix64(dx_hash_row(reinterpret_cast<const uint4*>(row), SP / 16)) >> 40) % world);
}
