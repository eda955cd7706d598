// OPEN: device-resident batched priority structure.
//
// Reference semantics (deepxube/pathfinding/graph_search.py:48-71): a binary heap of
// (f, push_count, item); every iteration pushes the kept items in order and pops min(B, |OPEN|) items,
// i.e. the B smallest by (f, push order), returned in ascending order; lb = max(lb, f of first popped).
// f = float64(w) * float64(g) + float64(h), two separately rounded float64 operations
// (graph_search.py:153, :175).
//
// Device formulation: per instance, OPEN is one run sorted by (f-key, item), where f-key is the order-preserving
// unsigned image of the float64 f (dx_f_to_key: correct for negative f too) and items are numbered in push
// order (A*: node id; Q*: node id * A + action), so the pair order IS the heap's (f, push_count) order.
// One iteration =
//   k_plan        per-instance sizes and offsets of this iteration (pushed m, popped k, new |OPEN|),
//   k_make_keys   f-keys of the pushed batch (__dmul_rn/__dadd_rn: no FMA contraction),
//   radix sort    stable LSD sort of the batch by (instance, f-key) -- 8-bit digits, digit passes whose
//                 byte is constant over the batch are skipped (flags computed on the device),
//   k_merge       merge-path merge of each instance's run with its sorted batch; the first k outputs
//                 become the next popped list (already in pop order), the rest the new run (ping-pong),
//   k_end_iter    lb / ub / finished() / itr bookkeeping (graph_search.py:43-46, 67-69).
// Every kernel sizes itself from the device-side control block; the host never synchronises.
#include "dx_internal.cuh"

#define OPEN_THREADS 256
#define ITEMS 8   // DX_SORT_TILE / OPEN_THREADS

static_assert(OPEN_THREADS * ITEMS == DX_SORT_TILE, "tile size");

// ---------------------------------------------------------------------------------------------
// (all 1024 threads of ONE block call this; s_tot = 4 x 1024 words of shared memory)
__device__ __forceinline__ void plan_body(uint32_t (*s_tot)[1024], Ctl* __restrict__ ctl, const InstArrays& ia,
                                          const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ pop_off_cur,
                                          uint32_t* __restrict__ pop_off_next, const uint32_t* __restrict__ open_off_cur,
                                          uint32_t* __restrict__ open_off_next, int child_mult, int push_mult, uint32_t max_pop,
                                          uint32_t max_open, int kept_is_total, int beam) {
    const int I = ctl->n_inst;
    const int per = (I + 1023) / 1024;
    const int lo = min(I, (int)threadIdx.x * per), hi = min(I, lo + per);
    uint32_t sm = 0, sk = 0, sn = 0, st = 0;
    for (int i = lo; i < hi; ++i) {
        const bool act = ia.active[i] != 0;
        // HDA*: the filtered batch is what this rank RECEIVED, unrelated to its popped list (single instance)
        const uint32_t kept = kept_is_total ? ctl->n_kept
                                            : prefix[pop_off_cur[i + 1] * child_mult] - prefix[pop_off_cur[i] * child_mult];
        const uint32_t m = act ? kept * push_mult : 0;
        const uint32_t n = act ? (open_off_cur[i + 1] - open_off_cur[i]) : 0;   // OPEN of a finished instance is dropped
        const uint32_t t = n + m;
        const uint32_t k = act ? min((uint32_t)ia.batch[i], t) : 0;
        ia.m_new[i] = m;
        ia.k_pop[i] = k;
        // beam search keeps nothing but the selected nodes (beam_search.py:113-115): OPEN stays empty
        ia.n_open[i] = beam ? 0u : t - k;
        sm += m; sk += k; sn += ia.n_open[i]; st += (t + DX_SORT_TILE - 1) / DX_SORT_TILE;
    }
    s_tot[0][threadIdx.x] = sm; s_tot[1][threadIdx.x] = sk; s_tot[2][threadIdx.x] = sn; s_tot[3][threadIdx.x] = st;
    __syncthreads();
    if (threadIdx.x < 4) {        // tiny serial scan over the partials of the threads that own instances (the rest hold 0 and
        uint32_t acc = 0;         // never read their prefix): one step for a single instance instead of 1024
        const int used = (I + per - 1) / max(per, 1);
        for (int t = 0; t < used; ++t) { uint32_t v = s_tot[threadIdx.x][t]; s_tot[threadIdx.x][t] = acc; acc += v; }
        if (threadIdx.x == 0) { ia.batch_off[I] = acc; ctl->sort_n = acc; }
        if (threadIdx.x == 1) { pop_off_next[I] = acc; if (acc > max_pop) atomicOr(&ctl->error, DX_DEVERR_POP); }
        if (threadIdx.x == 2) { open_off_next[I] = acc; ctl->open_total = acc; if (acc > max_open) atomicOr(&ctl->error, DX_DEVERR_OPEN); }
        if (threadIdx.x == 3) { ia.tile_off[I] = acc; ctl->n_tiles = acc; }
    }
    __syncthreads();
    uint32_t om = s_tot[0][threadIdx.x], ok = s_tot[1][threadIdx.x], on = s_tot[2][threadIdx.x], ot = s_tot[3][threadIdx.x];
    for (int i = lo; i < hi; ++i) {
        ia.batch_off[i] = om; pop_off_next[i] = ok; open_off_next[i] = on; ia.tile_off[i] = ot;
        const uint32_t m = ia.m_new[i], k = ia.k_pop[i], n = ia.n_open[i];
        const uint32_t t = beam ? m : n + k;          // elements the merge of this instance walks
        om += m; ok += k; on += n; ot += (t + DX_SORT_TILE - 1) / DX_SORT_TILE;
    }
}

__global__ void __launch_bounds__(1024)
k_plan(Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ pop_off_cur,
       uint32_t* __restrict__ pop_off_next, const uint32_t* __restrict__ open_off_cur, uint32_t* __restrict__ open_off_next,
       int child_mult, int push_mult, uint32_t max_pop, uint32_t max_open, int kept_is_total, int beam) {
    __shared__ uint32_t s_tot[4][1024];
    if (ctl->error) return;
    plan_body(s_tot, ctl, ia, prefix, pop_off_cur, pop_off_next, open_off_cur, open_off_next, child_mult, push_mult, max_pop,
              max_open, kept_is_total, beam);
}

int dx_launch_plan(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* child_prefix, const uint32_t* pop_off_cur,
                   uint32_t* pop_off_next, const uint32_t* open_off_cur, uint32_t* open_off_next, int child_mult, int push_mult,
                   uint32_t max_pop, uint32_t max_open, int kept_is_total, int beam) {
    k_plan<<<1, 1024, 0, s>>>(ctl, ia, child_prefix, pop_off_cur, pop_off_next, open_off_cur, open_off_next, child_mult,
                              push_mult, max_pop, max_open, kept_is_total, beam);
    return 1;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f_key(double w, float g, float h) {
    // float64(w) * float64(g) + float64(h), each rounded once (no FMA) -- graph_search.py:153
    const double f = __dadd_rn(__dmul_rn(w, (double)g), (double)h);
    return dx_f_to_key(f);
}

// A*: batch element k is new node (node_base + k).
// beam: the selection cost of a child is parent_t_cost + heuristic = 1.0 + h (beam_search.py:183-187; the reference takes
// the beam_size LARGEST logits = -(cost), i.e. the smallest costs)
__global__ void k_make_keys_v(const Ctl* __restrict__ ctl, InstArrays ia, const float* __restrict__ node_g,
                              const float* __restrict__ node_h, const uint32_t* __restrict__ new_inst, SortBufs sb, int beam) {
    if (ctl->error) return;
    const uint32_t n = ctl->sort_n, base = ctl->node_base;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t id = base + k, inst = new_inst[k];
        sb.key[0][k] = beam ? dx_f_to_key(__dadd_rn(1.0, (double)node_h[id]))
                            : f_key(ia.weight[inst], node_g[id], node_h[id]);
        sb.inst[0][k] = inst;
        sb.val[0][k] = id;
    }
}

// Q*: batch element (r*A + a) is edge (kept popped node r, action a); f = w * g(node) + q(node, a)
// (graph_search.py:168-179; edge order = kept-node order x action id, base/pathfinding.py:568-588).
__global__ void k_make_keys_q(const Ctl* __restrict__ ctl, InstArrays ia, const float* __restrict__ node_g,
                              const float* __restrict__ node_q, const uint32_t* __restrict__ popped,
                              const uint32_t* __restrict__ popped_inst, const uint8_t* __restrict__ flags,
                              const uint32_t* __restrict__ prefix, SortBufs sb, int A, int beam) {
    if (ctl->error) return;
    const uint32_t n = ctl->n_pop * A;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t j = t / A, a = t % A;
        if (!(flags[j] & 1)) continue;
        const uint32_t node = popped[j], inst = popped_inst[j];
        const uint32_t k = prefix[j] * A + a;
        // beam_q ranks an edge by its q-value alone (logits = -q_val, beam_search.py:206-212)
        sb.key[0][k] = beam ? dx_f_to_key((double)node_q[(size_t)node * A + a])
                            : f_key(ia.weight[inst], node_g[node], node_q[(size_t)node * A + a]);
        sb.inst[0][k] = inst;
        sb.val[0][k] = node * A + a;
    }
}

int dx_launch_make_keys(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const float* node_g, const float* node_h,
                        const uint32_t* new_inst, const SortBufs& sb, int max_n, int beam) {
    int blocks = min(dx_div_up(max_n, 256), 148 * 8);
    k_make_keys_v<<<blocks, 256, 0, s>>>(ctl, ia, node_g, node_h, new_inst, sb, beam);
    return 1;
}

int dx_launch_make_keys_q(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const float* node_g, const float* node_q,
                          const uint32_t* popped, const uint32_t* popped_inst, const uint8_t* flags, const uint32_t* prefix,
                          const SortBufs& sb, int A, int max_pop, int beam) {
    int blocks = min(dx_div_up((long long)max_pop * A, 256), 148 * 8);
    k_make_keys_q<<<blocks, 256, 0, s>>>(ctl, ia, node_g, node_q, popped, popped_inst, flags, prefix, sb, A, beam);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// Radix sort (LSD, 8-bit digits, stable).  Digit d < 8: bits [cut + 8 d, cut + 8 d + 8) of the f-key; d >= 8: byte d-8 of the
// instance id.  cut = SortBufs::key_cut: the low 23 bits of the key do NOT go through the digit passes.  f = w g + h has 52
// varying mantissa bits (w g is a full double) but h is a float32: two different (g, h) pairs land within 2^-29 of each other
// about once per 10^8 pairs, so after sorting on (instance, key >> 23) almost every group of equal sort keys is a single entry or a
// set of exact ties; k_sort_fixup / k_sort_fixup_big put the few others into full-key order (stable, so the result is the order of
// the full 64-bit sort for ANY input -- only the time depends on the assumption).  That is 4-5 digit passes instead of 8-9
// (sign / high exponent digits are constant and skipped as before).
__device__ __forceinline__ uint32_t digit_of(unsigned long long key, uint32_t inst, int d, int cut) {
    if (d >= 8) return (inst >> (8 * (d - 8))) & 255u;
    const int sh = cut + 8 * d;
    return sh < 64 ? (uint32_t)(key >> sh) & 255u : 0u;
}

__global__ void __launch_bounds__(OPEN_THREADS)
k_radix_prepass(const Ctl* __restrict__ ctl, SortBufs sb) {
    __shared__ uint32_t s_h[DX_NUM_DIGITS * 256];
    if (ctl->error) return;
    const uint32_t n = ctl->sort_n;
    if (blockIdx.x * OPEN_THREADS >= n && blockIdx.x > 0) return;
    for (int i = threadIdx.x; i < DX_NUM_DIGITS * 256; i += OPEN_THREADS) s_h[i] = 0;
    __syncthreads();
    for (uint32_t k = blockIdx.x * OPEN_THREADS + threadIdx.x; k < n; k += gridDim.x * OPEN_THREADS) {
        const unsigned long long key = sb.key[0][k];
        const uint32_t inst = sb.inst[0][k];
#pragma unroll
        for (int d = 0; d < DX_NUM_DIGITS; ++d) atomicAdd(&s_h[d * 256 + digit_of(key, inst, d, sb.key_cut)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < DX_NUM_DIGITS * 256; i += OPEN_THREADS)
        if (s_h[i]) atomicAdd(&sb.ghist[i], s_h[i]);
}

// skip a pass when one bin holds every element; assign ping-pong sources to the remaining passes.
// Also turns each digit position's global histogram into exclusive bin offsets (gbase), the starting
// output position of every digit value in that pass.
__global__ void k_radix_flags(Ctl* __restrict__ ctl, SortBufs sb) {
    __shared__ uint32_t s_skip[DX_NUM_DIGITS];
    __shared__ uint32_t s_cnt[DX_NUM_DIGITS * 256];
    if (ctl->error) return;
    const uint32_t n = ctl->sort_n;
    if (threadIdx.x < DX_NUM_DIGITS) s_skip[threadIdx.x] = (n == 0) ? 1u : 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < DX_NUM_DIGITS * 256; i += blockDim.x) {
        const uint32_t c = sb.ghist[i];
        s_cnt[i] = c;
        if (n != 0 && c == n) s_skip[i / 256] = 1u;
        sb.ghist[i] = 0;
    }
    __syncthreads();
    {   // exclusive scan of every digit position's 256 counts: a warp per position, 8 consecutive bins per lane
        const int lane = threadIdx.x & 31;
        for (int d = threadIdx.x >> 5; d < DX_NUM_DIGITS; d += blockDim.x >> 5) {
            uint32_t c[8], sum = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) { c[k] = s_cnt[d * 256 + lane * 8 + k]; sum += c[k]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            uint32_t acc = incl - sum;
#pragma unroll
            for (int k = 0; k < 8; ++k) { sb.gbase[d * 256 + lane * 8 + k] = acc; acc += c[k]; }
        }
    }
    if (threadIdx.x < DX_NUM_DIGITS) sb.tickets[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        uint32_t src = 0;
        for (int d = 0; d < DX_NUM_DIGITS; ++d) {
            ctl->pass_skip[d] = s_skip[d];
            ctl->pass_src[d] = src;
            if (!s_skip[d]) src ^= 1u;
        }
        ctl->sort_final = src;
        sb.tickets[DX_NUM_DIGITS] += 1;        // new tag for the look-back words of this sort's passes
        sb.fix_list[0] = 0;
    }
}

__global__ void __launch_bounds__(OPEN_THREADS)
k_radix_hist(const Ctl* __restrict__ ctl, SortBufs sb, int d) {
    __shared__ uint32_t s_h[256];
    if (ctl->error || ctl->pass_skip[d]) return;
    const uint32_t n = ctl->sort_n;
    const uint32_t src = ctl->pass_src[d];
    const uint32_t tile = blockIdx.x;
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = tile * DX_SORT_TILE;
    if (base >= n) return;                 // tiles past the end are never read by the tile scan
    {
        const unsigned long long* __restrict__ keys = sb.key[src];
        const uint32_t* __restrict__ insts = sb.inst[src];
#pragma unroll
        for (int r = 0; r < ITEMS; ++r) {
            const uint32_t e = base + r * OPEN_THREADS + threadIdx.x;
            if (e < n) atomicAdd(&s_h[digit_of(keys[e], insts[e], d, sb.key_cut)], 1u);
        }
    }
    __syncthreads();
    sb.hist[(size_t)threadIdx.x * sb.max_tiles + tile] = s_h[threadIdx.x];
}

// One block per digit value: exclusive scan of that value's per-tile counts, offset by the value's global base.
__global__ void __launch_bounds__(OPEN_THREADS)
k_radix_tilescan(const Ctl* __restrict__ ctl, SortBufs sb, int d) {
    __shared__ uint32_t s_warp[OPEN_THREADS / 32];
    if (ctl->error || ctl->pass_skip[d]) return;
    const uint32_t n = ctl->sort_n;
    const uint32_t ntiles = (n + DX_SORT_TILE - 1) / DX_SORT_TILE;
    const uint32_t digit = blockIdx.x;
    uint32_t* __restrict__ h = sb.hist + (size_t)digit * sb.max_tiles;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t carry = sb.gbase[d * 256 + digit];
    for (uint32_t base = 0; base < ntiles; base += OPEN_THREADS) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < ntiles ? h[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        uint32_t woff = 0, total = 0;
#pragma unroll
        for (int w = 0; w < OPEN_THREADS / 32; ++w) {
            const uint32_t c = s_warp[w];
            if (w < warp) woff += c;
            total += c;
        }
        if (i < ntiles) h[i] = carry + woff + x - v;
        carry += total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(OPEN_THREADS)
k_radix_scatter(const Ctl* __restrict__ ctl, SortBufs sb, int d) {
    __shared__ uint32_t s_cnt[OPEN_THREADS / 32][256];
    if (ctl->error || ctl->pass_skip[d]) return;
    const uint32_t n = ctl->sort_n;
    const uint32_t src = ctl->pass_src[d], dst = src ^ 1u;
    const uint32_t tile = blockIdx.x;
    const uint32_t base = tile * DX_SORT_TILE;
    if (base >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (OPEN_THREADS / 32) * 256; i += OPEN_THREADS) (&s_cnt[0][0])[i] = 0;
    __syncthreads();

    unsigned long long key[ITEMS];
    uint32_t inst[ITEMS], val[ITEMS], rank[ITEMS], dig[ITEMS];
    const unsigned long long* __restrict__ keys = sb.key[src];
    const uint32_t* __restrict__ insts = sb.inst[src];
    const uint32_t* __restrict__ vals = sb.val[src];
    // warp w owns the contiguous elements [w*256, (w+1)*256) of the tile, 32 per round: stability = (warp, round, lane)
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        const uint32_t e = base + warp * (32 * ITEMS) + r * 32 + lane;
        const bool valid = e < n;
        key[r] = valid ? keys[e] : 0ull;
        inst[r] = valid ? insts[e] : 0u;
        val[r] = valid ? vals[e] : 0u;
        dig[r] = valid ? digit_of(key[r], inst[r], d, sb.key_cut) : (256u + lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, dig[r]);
        const int leader = __ffs(peers) - 1;
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = s_cnt[warp][dig[r]];
            s_cnt[warp][dig[r]] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rank[r] = old + before;
        __syncwarp();
    }
    __syncthreads();
    {   // per digit: exclusive offsets over warps, on top of the global (digit, tile) offset
        const uint32_t dg = threadIdx.x;
        uint32_t acc = sb.hist[(size_t)dg * sb.max_tiles + tile];
#pragma unroll
        for (int w = 0; w < OPEN_THREADS / 32; ++w) {
            const uint32_t c = s_cnt[w][dg];
            s_cnt[w][dg] = acc;
            acc += c;
        }
    }
    __syncthreads();
    unsigned long long* __restrict__ okeys = sb.key[dst];
    uint32_t* __restrict__ oinsts = sb.inst[dst];
    uint32_t* __restrict__ ovals = sb.val[dst];
#pragma unroll
    for (int r = 0; r < ITEMS; ++r) {
        if (dig[r] < 256u) {
            const uint32_t pos = s_cnt[warp][dig[r]] + rank[r];
            okeys[pos] = key[r];
            oinsts[pos] = inst[r];
            ovals[pos] = val[r];
        }
    }
}


// One digit pass in ONE kernel (decoupled look-back, "onesweep"): a block takes the next tile in ticket order, ranks its
// elements by digit (stable: warp w owns a contiguous eighth of the tile, rounds of 32, lanes in order), publishes the tile's
// per-digit counts, and thread `dg` walks back over the preceding tiles' words of digit dg until it meets an inclusive
// prefix -- replacing the per-tile histogram kernel, the 256-block tile scan and the scatter kernel (3 launches and two
// extra reads of the batch per pass).  Words are tagged with the pass's epoch, so the status array is never cleared;
// ticket order guarantees that every tile a block waits for has already been started by a resident (or finished) block.
#define OS_FLAG_AGG 1ull
#define OS_FLAG_PREFIX 2ull
#ifndef OS_ITEMS
#define OS_ITEMS 8                        // elements per thread: tiles of 2048.  16 (tiles of 4096) needed 181 registers = ONE block per SM
#endif                                    // (ncu: 12.5 % occupancy, 87 % of cycles without an eligible warp); 8 measured -18 % on the
#ifndef OS_MIN_BLOCKS                     // npuzzle.5 sort (64 instances x 4 k entries) and -2 % on the cube3 Q* sort (8 x 100 k)
#define OS_MIN_BLOCKS 2
#endif
#define OS_TILE (OPEN_THREADS * OS_ITEMS)
#ifndef OS_WINDOW
#define OS_WINDOW 8                       // predecessor words fetched per round trip of the look-back
#endif
// -DOS_PROFILE: clock64 timeline of thread 0 of every block, summed over tiles: [0] tiles, [1] ticket, [2] loads, [3] ranking,
// [4] count scan + publish, [5] look-back, [6] scatter (read and cleared through dx_tail_profile_read / dx_debug_tc_profile with
// DXB200_PROF_TAIL set).  cube3 Q*, 8 x 100 k entries (396 tiles): ticket 1.2 k, loads issued 1.1 k, ranking (absorbs the load latency) 4.6 k,
// scan + publish 0.3 k, look-back 6.8 k (3.2 k with one instance = 50 tiles), scatter 3.7 k cycles per tile -- latency end to end: tile sizes
// 1024-3072, 2-4 blocks per SM, look-back windows 8-32 and ballot-based ranking all measured within +-3 % of each other
#ifdef OS_PROFILE
__device__ unsigned long long g_os_prof[16];
#define OS_MARK(slot) do { if (threadIdx.x == 0) { const long long t_now = clock64(); atomicAdd(&g_os_prof[slot], (unsigned long long)(t_now - t_prev)); t_prev = t_now; } } while (0)
#else
#define OS_MARK(slot) do { } while (0)
#endif
__global__ void __launch_bounds__(OPEN_THREADS, OS_MIN_BLOCKS)
k_radix_onesweep(Ctl* __restrict__ ctl, SortBufs sb, int d) {
    __shared__ uint32_t s_cnt[OPEN_THREADS / 32][256];
    __shared__ uint32_t s_base[256];
    __shared__ uint32_t s_tile;
    if (ctl->error || ctl->pass_skip[d]) return;
    const uint32_t n = ctl->sort_n;
    const uint32_t ntiles = (n + OS_TILE - 1) / OS_TILE;
    const uint32_t src = ctl->pass_src[d], dst = src ^ 1u;
    const unsigned long long epoch = ((unsigned long long)(sb.tickets[DX_NUM_DIGITS] * (uint32_t)DX_NUM_DIGITS + (uint32_t)d + 1u)) << 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long* __restrict__ keys = sb.key[src];
    const uint32_t* __restrict__ insts = sb.inst[src];
    const uint32_t* __restrict__ vals = sb.val[src];
    unsigned long long* __restrict__ okeys = sb.key[dst];
    uint32_t* __restrict__ oinsts = sb.inst[dst];
    uint32_t* __restrict__ ovals = sb.val[dst];
#ifdef OS_PROFILE
    long long t_prev = clock64();
#endif
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(&sb.tickets[d], 1u);
        for (int i = threadIdx.x; i < (OPEN_THREADS / 32) * 256; i += OPEN_THREADS) (&s_cnt[0][0])[i] = 0;
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) return;
        OS_MARK(1);
#ifdef OS_PROFILE
        if (threadIdx.x == 0) atomicAdd(&g_os_prof[0], 1ull);
#endif
        const uint32_t base = tile * OS_TILE;
        unsigned long long key[OS_ITEMS];
        uint32_t inst[OS_ITEMS], val[OS_ITEMS];
        uint16_t rank[OS_ITEMS], dig[OS_ITEMS];
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) {
            const uint32_t e = base + warp * (32 * OS_ITEMS) + r * 32 + lane;
            const bool valid = e < n;
            key[r] = valid ? keys[e] : 0ull;
            inst[r] = valid ? insts[e] : 0u;
            val[r] = valid ? vals[e] : 0u;
        }
        if (key[0] == 1ull && inst[OS_ITEMS - 1] == 0xFFFFFFFFu && val[0] == 7u) __syncwarp();   // (keeps the loads ahead of the first mark)
        OS_MARK(2);
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) {
            const bool valid = base + warp * (32 * OS_ITEMS) + r * 32 + lane < n;
            const uint32_t dg = valid ? digit_of(key[r], inst[r], d, sb.key_cut) : (256u + lane);
            const uint32_t peers = __match_any_sync(0xffffffffu, dg);
            const int leader = __ffs(peers) - 1;
            const uint32_t before = __popc(peers & ((1u << lane) - 1u));
            uint32_t old = 0;
            if (valid && lane == leader) {
                old = s_cnt[warp][dg];
                s_cnt[warp][dg] = old + __popc(peers);
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            rank[r] = (uint16_t)(old + before);
            dig[r] = (uint16_t)dg;
            __syncwarp();
        }
        __syncthreads();
        OS_MARK(3);
        {   // thread dg: this tile's count of digit dg, exclusive offsets over the warps, then the look-back
            const uint32_t dg = threadIdx.x;
            uint32_t acc = 0;
#pragma unroll
            for (int w = 0; w < OPEN_THREADS / 32; ++w) {
                const uint32_t c = s_cnt[w][dg];
                s_cnt[w][dg] = acc;
                acc += c;
            }
            volatile unsigned long long* st = sb.status;
            if (tile == 0) {
                st[dg] = epoch | (OS_FLAG_PREFIX << 30) | acc;
                s_base[dg] = sb.gbase[d * 256 + dg];
            } else {
                st[(size_t)tile * 256 + dg] = epoch | (OS_FLAG_AGG << 30) | acc;
                OS_MARK(4);
                uint32_t excl = 0;
                int t = (int)tile - 1;
                bool done = false;
                uint32_t spins = 0;
                while (!done) {
                    // OS_WINDOW predecessor words per round trip (independent loads); consumed nearest first, stopping at the first
                    // inclusive prefix; a word whose pass has not published yet ends the round and is polled again
                    unsigned long long w[OS_WINDOW];
#pragma unroll
                    for (int j = 0; j < OS_WINDOW; ++j) w[j] = (t - j >= 0) ? st[(size_t)(t - j) * 256 + dg] : (epoch | (OS_FLAG_PREFIX << 30));
#pragma unroll
                    for (int j = 0; j < OS_WINDOW; ++j) {
                        if (done || (w[j] & 0xFFFFFFFF00000000ull) != epoch) break;
                        excl += (uint32_t)(w[j] & 0x3FFFFFFFull);
                        --t;
                        if (((w[j] >> 30) & 3ull) == OS_FLAG_PREFIX) done = true;
                    }
                    if (++spins > (1u << 24)) { atomicOr(&ctl->error, DX_DEVERR_POP); done = true; }   // protocol bug: fail, do not hang
                }
                st[(size_t)tile * 256 + dg] = epoch | (OS_FLAG_PREFIX << 30) | (excl + acc);
                s_base[dg] = sb.gbase[d * 256 + dg] + excl;
            }
        }
        __syncthreads();
        OS_MARK(5);
#pragma unroll
        for (int r = 0; r < OS_ITEMS; ++r) {
            if (dig[r] < 256u) {
                const uint32_t pos = s_base[dig[r]] + s_cnt[warp][dig[r]] + rank[r];
                okeys[pos] = key[r];
                oinsts[pos] = inst[r];
                ovals[pos] = val[r];
            }
        }
        OS_MARK(6);
    }
}

// Small batches (<= DX_SMALL_SORT entries, known on the host from the instances' batch sizes): the whole sort is ONE
// block -- a bitonic network in shared memory on (instance, f-key, original position), the position making it the same
// stable order as the radix sort -- instead of 2 + 3 launches per digit pass, which at these sizes are pure launch latency
// (26 of the 40 launches of an iteration).  Output goes to buffer 1, like an odd number of radix passes.
// (all 1024 threads of ONE block; s_key / s_inst / s_idx hold DX_SMALL_SORT entries each)
__device__ __forceinline__ void sort_small_body(Ctl* __restrict__ ctl, const SortBufs& sb, unsigned long long* s_key, uint32_t* s_inst,
                                                uint32_t* s_idx) {
    const uint32_t n = ctl->sort_n, nt = blockDim.x;
    if (n > DX_SMALL_SORT) {                 // the host's bound was wrong: fail loudly instead of sorting a prefix
        if (threadIdx.x == 0) atomicOr(&ctl->error, DX_DEVERR_POP);
        return;
    }
    if (threadIdx.x == 0) ctl->sort_final = 1;
    if (n == 0) return;
    uint32_t m = 1;
    while (m < n) m <<= 1;
    for (uint32_t i = threadIdx.x; i < m; i += nt) {
        const bool in = i < n;
        s_key[i] = in ? sb.key[0][i] : ~0ull;
        s_inst[i] = in ? sb.inst[0][i] : 0xFFFFFFFFu;
        s_idx[i] = i;
    }
    // one compare-exchange pair per thread and stage (m / 2 <= 1024 pairs).  Pair q of stage j is (i, i | j) with i = q with a
    // zero bit inserted at log2(j): for j <= 32 both elements lie in the 64-element block owned by q's warp, so those stages
    // (51 of the 66 for m = 2048) only need __syncwarp(); a block barrier is needed when the previous stage crossed warps.
    __syncthreads();
    const uint32_t half = m >> 1, q = threadIdx.x;
    for (uint32_t k = 2; k <= m; k <<= 1)
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) __syncthreads(); else __syncwarp();
            if (q < half) {
                const uint32_t i = ((q & ~(j - 1)) << 1) | (q & (j - 1)), p = i | j;
                const unsigned long long ka = s_key[i], kb = s_key[p];
                const uint32_t ia = s_inst[i], ib = s_inst[p], xa = s_idx[i], xb = s_idx[p];
                const bool a_gt_b = ia != ib ? ia > ib : (ka != kb ? ka > kb : xa > xb);
                if (a_gt_b == ((i & k) == 0)) {
                    s_key[i] = kb; s_key[p] = ka;
                    s_inst[i] = ib; s_inst[p] = ia;
                    s_idx[i] = xb; s_idx[p] = xa;
                }
            }
        }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += nt) {
        sb.key[1][i] = s_key[i];
        sb.inst[1][i] = s_inst[i];
        sb.val[1][i] = sb.val[0][s_idx[i]];
    }
}

__global__ void __launch_bounds__(1024)
k_sort_small(Ctl* __restrict__ ctl, SortBufs sb) {
    __shared__ unsigned long long s_key[DX_SMALL_SORT];
    __shared__ uint32_t s_inst[DX_SMALL_SORT];
    __shared__ uint32_t s_idx[DX_SMALL_SORT];
    if (ctl->error) return;
    sort_small_body(ctl, sb, s_key, s_inst, s_idx);
}

// Small batches, A*: k_plan + k_make_keys_v + k_sort_small in one block (three launches -> one); the plan's scratch and the
// sort's arrays share the block's shared memory.
__global__ void __launch_bounds__(1024)
k_small_push_v(Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ pop_off_cur,
               uint32_t* __restrict__ pop_off_next, const uint32_t* __restrict__ open_off_cur, uint32_t* __restrict__ open_off_next,
               int A, uint32_t max_pop, uint32_t max_open, const float* __restrict__ node_g, const float* __restrict__ node_h,
               const uint32_t* __restrict__ new_inst, SortBufs sb) {
    __shared__ unsigned long long s_key[DX_SMALL_SORT];      // 16 KB, first used as the plan's 4 x 1024 words
    __shared__ uint32_t s_inst[DX_SMALL_SORT];
    __shared__ uint32_t s_idx[DX_SMALL_SORT];
    __shared__ uint32_t s_err;
    static_assert(sizeof(unsigned long long) * DX_SMALL_SORT >= 4 * 1024 * sizeof(uint32_t), "plan scratch fits");
    dx_pdl_prologue();
    if (ctl->error) return;
    plan_body(reinterpret_cast<uint32_t (*)[1024]>(s_key), ctl, ia, prefix, pop_off_cur, pop_off_next, open_off_cur, open_off_next, A,
              1, max_pop, max_open, 0, 0);
    __syncthreads();
    if (threadIdx.x == 0) s_err = *((volatile uint32_t*)&ctl->error);
    __syncthreads();
    if (s_err) return;
    const uint32_t n = ctl->sort_n, base = ctl->node_base;
    for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) {          // k_make_keys_v
        const uint32_t id = base + k, inst = new_inst[k];
        sb.key[0][k] = f_key(ia.weight[inst], node_g[id], node_h[id]);
        sb.inst[0][k] = inst;
        sb.val[0][k] = id;
    }
    __syncthreads();
    sort_small_body(ctl, sb, s_key, s_inst, s_idx);
}

// Small batches, Q*: k_plan + k_make_keys_q + k_sort_small in one block.
__global__ void __launch_bounds__(1024)
k_small_push_q(Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ pop_off_cur,
               uint32_t* __restrict__ pop_off_next, const uint32_t* __restrict__ open_off_cur, uint32_t* __restrict__ open_off_next,
               int A, uint32_t max_pop, uint32_t max_open, const float* __restrict__ node_g, const float* __restrict__ node_q,
               const uint32_t* __restrict__ popped, const uint32_t* __restrict__ popped_inst, const uint8_t* __restrict__ flags,
               SortBufs sb) {
    __shared__ unsigned long long s_key[DX_SMALL_SORT];
    __shared__ uint32_t s_inst[DX_SMALL_SORT];
    __shared__ uint32_t s_idx[DX_SMALL_SORT];
    __shared__ uint32_t s_err;
    dx_pdl_prologue();
    if (ctl->error) return;
    plan_body(reinterpret_cast<uint32_t (*)[1024]>(s_key), ctl, ia, prefix, pop_off_cur, pop_off_next, open_off_cur, open_off_next, 1,
              A, max_pop, max_open, 0, 0);
    __syncthreads();
    if (threadIdx.x == 0) s_err = *((volatile uint32_t*)&ctl->error);
    __syncthreads();
    if (s_err) return;
    const uint32_t n = ctl->n_pop * A;
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {           // k_make_keys_q
        const uint32_t j = t / A, a = t % A;
        if (!(flags[j] & 1)) continue;
        const uint32_t node = popped[j], inst = popped_inst[j];
        const uint32_t k = prefix[j] * A + a;
        sb.key[0][k] = f_key(ia.weight[inst], node_g[node], node_q[(size_t)node * A + a]);
        sb.inst[0][k] = inst;
        sb.val[0][k] = node * A + a;
    }
    __syncthreads();
    sort_small_body(ctl, sb, s_key, s_inst, s_idx);
}

int dx_launch_small_push_q(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* child_prefix, const uint32_t* pop_off_cur,
                           uint32_t* pop_off_next, const uint32_t* open_off_cur, uint32_t* open_off_next, int A, uint32_t max_pop,
                           uint32_t max_open, const float* node_g, const float* node_q, const uint32_t* popped,
                           const uint32_t* popped_inst, const uint8_t* flags, const SortBufs& sb) {
    dx_launch(k_small_push_q, 1, 1024, 0, s, ctl, ia, child_prefix, pop_off_cur, pop_off_next, open_off_cur, open_off_next, A, max_pop,
              max_open, node_g, node_q, popped, popped_inst, flags, sb);
    return 1;
}

int dx_launch_small_push_v(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* child_prefix, const uint32_t* pop_off_cur,
                           uint32_t* pop_off_next, const uint32_t* open_off_cur, uint32_t* open_off_next, int A, uint32_t max_pop,
                           uint32_t max_open, const float* node_g, const float* node_h, const uint32_t* new_inst, const SortBufs& sb) {
    dx_launch(k_small_push_v, 1, 1024, 0, s, ctl, ia, child_prefix, pop_off_cur, pop_off_next, open_off_cur, open_off_next, A, max_pop,
              max_open, node_g, node_h, new_inst, sb);
    return 1;
}

// After the digit passes: entries are in (instance, key >> cut, original position) order; what is left to do is a stable sort by
// the low bits inside every group of equal (instance, key >> cut).  Almost every group is a single entry or a set of exact ties
// (a clipped heuristic gives f = w g for 10^5 entries at once): nothing to do, and that is found out with O(1) work per entry --
// k_sort_fixup (read-only) lists the entries that are SMALLER than their predecessor in the same group (inversions), at most one per
// FIX_BACK positions of a group; k_sort_fixup_groups repairs the listed groups, a block each.
#define FIX_BACK 64
__device__ __forceinline__ bool fix_same(const unsigned long long* K, const uint32_t* In, uint32_t j, uint32_t inst,
                                         unsigned long long head, int cut) {
    return In[j] == inst && (K[j] >> cut) == head;
}

__global__ void __launch_bounds__(256)
k_sort_fixup(const Ctl* __restrict__ ctl, SortBufs sb) {
    if (ctl->error) return;
    const uint32_t n = ctl->sort_n;
    const int cut = sb.key_cut;
    const unsigned long long* __restrict__ K = sb.key[ctl->sort_final];
    const uint32_t* __restrict__ In = sb.inst[ctl->sort_final];
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x + 1; e < n; e += gridDim.x * blockDim.x) {
        const unsigned long long key = K[e], pk = K[e - 1];
        const uint32_t inst = In[e];
        const unsigned long long head = key >> cut;
        if (pk <= key || (pk >> cut) != head || In[e - 1] != inst) continue;               // no inversion at e
        // an earlier inversion of the same group within FIX_BACK positions has this group listed already
        bool earlier = false;
        uint32_t lo = e - 1;
        for (int steps = 0; steps < FIX_BACK && lo > 0 && fix_same(K, In, lo - 1, inst, head, cut); ++steps, --lo)
            if (K[lo - 1] > K[lo]) { earlier = true; break; }
        if (!earlier) sb.fix_list[1 + atomicAdd(&sb.fix_list[0], 1u)] = e;
    }
}

// A block per listed inversion: the group [lo, hi) around it (256 positions per round trip), claimed through fix_claim[lo] (several
// listed inversions may belong to one long group; the word carries the sort's epoch, so it is never cleared), then its non-decreasing
// runs are merged by ranking -- an entry's output position is its offset in its own run plus, for every other run, the number of
// entries that precede it there (earlier runs win ties: stable).  O(len x runs x log len) loads: an outlier inside 10^5 exact ties is
// 3 runs.  With more than FIX_RUNS runs (keys whose information sits below the cut: hand-made heuristics with many distinct values
// inside one 2^cut window) the run starts go to global scratch and adjacent runs are merged pairwise, round by round (a merge sort
// by ranking, O(len x log runs x log len)), and the control block is flagged so that the host stops using the cut for this engine.
#define FIX_RUNS 1024
__global__ void __launch_bounds__(256)
k_sort_fixup_groups(Ctl* __restrict__ ctl, SortBufs sb) {
    __shared__ uint32_t s_run[FIX_RUNS + 1];
    __shared__ uint32_t s_w[8];
    __shared__ uint32_t s_a, s_b, s_cnt, s_mine;
    if (ctl->error) return;
    const uint32_t groups = sb.fix_list[0];
    if (groups == 0) return;
    const uint32_t n = ctl->sort_n, fin = ctl->sort_final;
    const uint32_t epoch = sb.tickets[DX_NUM_DIGITS];
    const int cut = sb.key_cut;
    unsigned long long* K = sb.key[fin];
    uint32_t* V = sb.val[fin];
    const uint32_t* __restrict__ In = sb.inst[fin];
    unsigned long long* TK = sb.key[fin ^ 1u];
    uint32_t* TV = sb.val[fin ^ 1u];
    uint32_t* scratch = sb.inst[fin ^ 1u];                 // (free after the last pass; a group uses the words of its own positions)
    const uint32_t tid = threadIdx.x;
    for (uint32_t gi = blockIdx.x; gi < groups; gi += gridDim.x) {
        const uint32_t x = sb.fix_list[1 + gi];
        const uint32_t inst = In[x];
        const unsigned long long head = K[x] >> cut;       // (the high bits of every entry of a group are the same: safe to read while
        __syncthreads();                                   //  another block rewrites the group)
        if (tid == 0) { s_a = 0; s_b = n; s_cnt = 0; s_mine = 0; }
        __syncthreads();
        for (uint32_t top = x;;) {                         // lo: positions top - 1 - tid; the nearest non-member below x decides
            const bool valid = tid < top;
            const uint32_t j = valid ? top - 1 - tid : 0;
            if (valid && !fix_same(K, In, j, inst, head, cut)) atomicMax(&s_a, j + 1);
            __syncthreads();
            const uint32_t found = s_a;
            __syncthreads();
            if (found > 0 || top <= 256) break;
            top -= 256;
        }
        const uint32_t lo = s_a;
        if (tid == 0) s_mine = atomicExch(&sb.fix_claim[lo], epoch) != epoch ? 1u : 0u;
        __syncthreads();
        if (!s_mine) continue;                             // (block-uniform) another listed inversion of this group got here first
        for (uint32_t base = x + 1; base < n; base += 256) {        // hi: the nearest non-member above x
            const uint32_t j = base + tid;
            if (j < n && !fix_same(K, In, j, inst, head, cut)) atomicMin(&s_b, j);
            __syncthreads();
            const uint32_t found = s_b;
            __syncthreads();
            if (found < n) break;
        }
        const uint32_t hi = s_b;
        uint32_t* R = scratch + lo;                                 // run starts (all of them; the first FIX_RUNS also in s_run)
        for (uint32_t base = lo; base < hi; base += 256) {          // run starts in position order (ordered compaction per 256)
            const uint32_t j = base + tid;
            const bool start = j < hi && (j == lo || K[j - 1] > K[j]);
            const uint32_t bal = __ballot_sync(0xffffffffu, start);
            if ((tid & 31) == 0) s_w[tid >> 5] = __popc(bal);
            __syncthreads();
            uint32_t off = s_cnt;
            for (uint32_t w = 0; w < (tid >> 5); ++w) off += s_w[w];
            off += __popc(bal & ((1u << (tid & 31)) - 1u));
            if (start && off < FIX_RUNS) s_run[off] = j;
            if (start) R[off] = j;
            __syncthreads();
            if (tid == 0) { uint32_t t = 0; for (int w = 0; w < 8; ++w) t += s_w[w]; s_cnt += t; }
            __syncthreads();
        }
        const uint32_t runs = s_cnt;
        if (runs <= FIX_RUNS) {
            if (tid == 0) s_run[runs] = hi;
            __syncthreads();
            for (uint32_t j = lo + tid; j < hi; j += 256) {
                const unsigned long long kj = K[j];
                uint32_t rl = 0, rh = runs - 1;                    // my run: the last one starting at or before j
                while (rl < rh) { const uint32_t m = (rl + rh + 1) >> 1; if (s_run[m] <= j) rl = m; else rh = m - 1; }
                uint32_t rank = j - s_run[rl];
                for (uint32_t r = 0; r < runs; ++r) {
                    if (r == rl) continue;
                    uint32_t a2 = s_run[r], b2 = s_run[r + 1];    // entries of run r that go before kj: <= for earlier runs, < for later ones
                    const uint32_t r0 = a2;
                    while (a2 < b2) {
                        const uint32_t m = (a2 + b2) >> 1;
                        const unsigned long long km = K[m];
                        if (r < rl ? km <= kj : km < kj) a2 = m + 1; else b2 = m;
                    }
                    rank += a2 - r0;
                }
                TK[lo + rank] = kj;
                TV[lo + rank] = V[j];
            }
        } else {
            if (tid == 0) ctl->sort_slow = 1;
            unsigned long long* sk = K; uint32_t* sv = V;             // ping-pong: the merged runs of a round land in the other pair
            unsigned long long* dk = TK; uint32_t* dv = TV;
            uint32_t nr = runs;
            __syncthreads();
            while (nr > 1) {
                for (uint32_t j = lo + tid; j < hi; j += 256) {
                    const unsigned long long kj = sk[j];
                    uint32_t rl = 0, rh = nr - 1;                  // my run
                    while (rl < rh) { const uint32_t m = (rl + rh + 1) >> 1; if (R[m] <= j) rl = m; else rh = m - 1; }
                    const uint32_t pa = rl & ~1u, pb = pa + 1;     // the pair (pa, pb); a last run without partner is copied
                    uint32_t pos = j;
                    if (pb < nr) {
                        const uint32_t a0 = R[pa], b0 = R[pb], b1 = pb + 1 < nr ? R[pb + 1] : hi;
                        uint32_t l2, h2;
                        if (rl == pa) { l2 = b0; h2 = b1; } else { l2 = a0; h2 = b0; }
                        const uint32_t base2 = l2;
                        while (l2 < h2) {                          // run A entries precede equal run B entries (stable)
                            const uint32_t m = (l2 + h2) >> 1;
                            const unsigned long long km = sk[m];
                            if (rl == pa ? km < kj : km <= kj) l2 = m + 1; else h2 = m;
                        }
                        pos = a0 + (j - (rl == pa ? a0 : b0)) + (l2 - base2);
                    }
                    dk[pos] = kj;
                    dv[pos] = sv[j];
                }
                __syncthreads();
                const uint32_t half = (nr + 1) >> 1;               // R[p] <- R[2 p], ascending chunks: read, barrier, write
                for (uint32_t c0 = 0; c0 < half; c0 += 256) {
                    const uint32_t pidx = c0 + tid;
                    const uint32_t v = pidx < half ? R[2 * pidx] : 0u;
                    __syncthreads();
                    if (pidx < half) R[pidx] = v;
                    __syncthreads();
                }
                nr = half;
                unsigned long long* tk = sk; sk = dk; dk = tk;
                uint32_t* tv = sv; sv = dv; dv = tv;
            }
            if (sk != K) {
                for (uint32_t j = lo + tid; j < hi; j += 256) { K[j] = sk[j]; V[j] = sv[j]; }
            }
            __syncthreads();
            continue;
        }
        __syncthreads();
        for (uint32_t j = lo + tid; j < hi; j += 256) { K[j] = TK[j]; V[j] = TV[j]; }
    }
}

static int launch_sort_fixup(cudaStream_t s, Ctl* ctl, const SortBufs& sb, int max_n) {
    if (sb.key_cut <= 0) return 0;
    k_sort_fixup<<<min(dx_div_up(max_n, 256), 148 * 8), 256, 0, s>>>(ctl, sb);
    k_sort_fixup_groups<<<148 * 4, 256, 0, s>>>(ctl, sb);
    return 2;
}

int dx_launch_sort(cudaStream_t s, Ctl* ctl, const SortBufs& sb, const ScanTemp& t, int max_n, bool small) {
    if (small) {
        k_sort_small<<<1, 1024, 0, s>>>(ctl, sb);
        return 1;
    }
    int launches = 0;
    const int pre_blocks = min(dx_div_up(max_n, OPEN_THREADS), 148 * 4);
    k_radix_prepass<<<pre_blocks, OPEN_THREADS, 0, s>>>(ctl, sb);
    k_radix_flags<<<1, 256, 0, s>>>(ctl, sb);
    launches += 2;
    if (sb.onesweep) {
        const int blocks = min(dx_div_up(max_n, OS_TILE), 148 * 4);           // persistent over tiles (ticket order)
        for (int d = 0; d < sb.num_digits; ++d) {
            if (d < 8 && sb.key_cut + 8 * d >= 64) continue;                   // (no key bits left in this digit)
            k_radix_onesweep<<<blocks, OPEN_THREADS, 0, s>>>(ctl, sb, d);
            ++launches;
        }
        return launches + launch_sort_fixup(s, ctl, sb, max_n);
    }
    for (int d = 0; d < sb.num_digits; ++d) {      // higher instance-id bytes are constant zero: never launched
        if (d < 8 && sb.key_cut + 8 * d >= 64) continue;
        k_radix_hist<<<sb.max_tiles, OPEN_THREADS, 0, s>>>(ctl, sb, d);
        k_radix_tilescan<<<256, OPEN_THREADS, 0, s>>>(ctl, sb, d);
        k_radix_scatter<<<sb.max_tiles, OPEN_THREADS, 0, s>>>(ctl, sb, d);
        launches += 3;
    }
    return launches + launch_sort_fixup(s, ctl, sb, max_n);
}

// ---------------------------------------------------------------------------------------------
// Merge-path merge of OPEN run (A) and sorted batch (B) per instance; ties take A first (older push).
__device__ __forceinline__ uint32_t merge_path(const unsigned long long* __restrict__ a, uint32_t na,
                                               const unsigned long long* __restrict__ b, uint32_t nb, uint32_t diag) {
    uint32_t lo = diag > nb ? diag - nb : 0, hi = min(diag, na);
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a[mid] <= b[diag - 1 - mid]) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// the same partition found by a whole warp: 32 probes per round trip (a 32-ary search, 3-4 dependent global accesses for any
// OPEN size instead of one per halving -- the partition search is a pure latency chain at the head of every merge tile)
__device__ __forceinline__ uint32_t merge_path_warp(const unsigned long long* __restrict__ a, uint32_t na,
                                                    const unsigned long long* __restrict__ b, uint32_t nb, uint32_t diag) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t lo = diag > nb ? diag - nb : 0, hi = min(diag, na);
    while (lo < hi) {                              // P(mid) = a[mid] <= b[diag - 1 - mid] is true ... true false ... false
        const uint32_t step = (hi - lo) / 32 + 1;
        const uint32_t idx = lo + lane * step;
        const bool p = idx < hi && a[idx] <= b[diag - 1 - idx];
        const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, p));
        if (cnt == 0) { hi = lo; break; }
        const uint32_t nlo = lo + (cnt - 1) * step + 1, nhi = min(hi, lo + cnt * step);
        lo = nlo; hi = nhi;
    }
    return lo;
}

// Partition points of every merge tile, found up front by one warp per tile (a 32-ary search: 3-4 dependent global accesses):
// part_a[T] = how many OPEN-run elements precede tile T's first output, part_i[T] = the tile's instance.  Inside k_merge
// this search was a serial latency chain at the head of every tile, with 6 of the block's 8 warps idle.
__global__ void __launch_bounds__(OPEN_THREADS)
k_merge_partition(const Ctl* __restrict__ ctl, InstArrays ia, SortBufs sb, const unsigned long long* __restrict__ open_key_cur,
                  const uint32_t* __restrict__ open_off_cur, uint32_t* __restrict__ part_a, uint32_t* __restrict__ part_i,
                  const uint32_t* __restrict__ a_head, const unsigned long long* __restrict__ b_keys,
                  const uint32_t* __restrict__ a_len, const uint32_t* __restrict__ n_tiles_ptr) {
    dx_pdl_prologue();
    if (ctl->error) return;
    const uint32_t n_tiles = n_tiles_ptr ? *n_tiles_ptr : ctl->n_tiles;
    const int I = ctl->n_inst;
    // (young-run compaction: the B run is the young run, and the A run starts behind its consumed prefix a_head[i])
    const unsigned long long* __restrict__ bkeys = b_keys ? b_keys : sb.key[ctl->sort_final];
    const uint32_t warps = gridDim.x * (OPEN_THREADS / 32);
    for (uint32_t T = blockIdx.x * (OPEN_THREADS / 32) + (threadIdx.x >> 5); T < n_tiles; T += warps) {
        int lo = 0, hi = I - 1;                 // instance of this tile: last i with tile_off[i] <= T
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (ia.tile_off[mid] <= T) lo = mid; else hi = mid - 1;
        }
        const int i = lo;
        const uint32_t ah = a_head ? a_head[i] : 0u;
        const uint32_t na = a_len ? a_len[i] : open_off_cur[i + 1] - open_off_cur[i] - ah, nb = ia.m_new[i];
        const uint32_t d0 = (T - ia.tile_off[i]) * DX_SORT_TILE;
        const uint32_t part = merge_path_warp(open_key_cur + open_off_cur[i] + ah, na, bkeys + ia.batch_off[i], nb, d0);
        if ((threadIdx.x & 31) == 0) { part_a[T] = part; part_i[T] = (uint32_t)i; }
    }
}

__global__ void __launch_bounds__(OPEN_THREADS)
k_merge(const Ctl* __restrict__ ctl, InstArrays ia, SortBufs sb, const unsigned long long* __restrict__ open_key_cur,
        const uint32_t* __restrict__ open_val_cur, const uint32_t* __restrict__ open_off_cur,
        unsigned long long* __restrict__ open_key_next, uint32_t* __restrict__ open_val_next,
        const uint32_t* __restrict__ open_off_next, uint32_t* __restrict__ popped_next, uint32_t* __restrict__ popped_inst_next,
        const uint32_t* __restrict__ pop_off_next, int keep_open, const uint32_t* __restrict__ part_a,
        const uint32_t* __restrict__ part_i, const uint32_t* __restrict__ a_head, const unsigned long long* __restrict__ b_keys,
        const uint32_t* __restrict__ b_vals, const uint32_t* __restrict__ a_len, const uint32_t* __restrict__ n_tiles_ptr) {
    __shared__ unsigned long long s_key[DX_SORT_TILE];
    __shared__ uint32_t s_val[DX_SORT_TILE];
    __shared__ uint32_t s_part[4];   // a0, a1, inst
    dx_pdl_prologue();
    if (ctl->error) return;
    const uint32_t n_tiles = n_tiles_ptr ? *n_tiles_ptr : ctl->n_tiles;
    const unsigned long long* __restrict__ bkeys = b_keys ? b_keys : sb.key[ctl->sort_final];
    const uint32_t* __restrict__ bvals = b_keys ? b_vals : sb.val[ctl->sort_final];
    for (uint32_t T = blockIdx.x; T < n_tiles; T += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const uint32_t i = part_i[T];
            const uint32_t na = a_len ? a_len[i] : open_off_cur[i + 1] - open_off_cur[i] - (a_head ? a_head[i] : 0u);
            s_part[0] = part_a[T];
            s_part[1] = (T + 1 < ia.tile_off[i + 1]) ? part_a[T + 1] : na;     // the instance's last tile ends at the end of its run
            s_part[2] = i;
        }
        __syncthreads();
        const int i = s_part[2];
        const uint32_t ah = a_head ? a_head[i] : 0u;
        const uint32_t na = a_len ? a_len[i] : open_off_cur[i + 1] - open_off_cur[i] - ah, nb = ia.m_new[i];
        const uint32_t d0 = (T - ia.tile_off[i]) * DX_SORT_TILE, d1 = min(d0 + DX_SORT_TILE, na + nb);
        const uint32_t a0 = s_part[0], a1 = s_part[1], b0 = d0 - a0, b1 = d1 - a1;
        const uint32_t ca = a1 - a0, cb = b1 - b0;
        const unsigned long long* __restrict__ ak = open_key_cur + open_off_cur[i] + ah + a0;
        const uint32_t* __restrict__ av = open_val_cur + open_off_cur[i] + ah + a0;
        const unsigned long long* __restrict__ bk = bkeys + ia.batch_off[i] + b0;
        const uint32_t* __restrict__ bv = bvals + ia.batch_off[i] + b0;
        for (uint32_t t = threadIdx.x; t < ca; t += OPEN_THREADS) { s_key[t] = ak[t]; s_val[t] = av[t]; }
        for (uint32_t t = threadIdx.x; t < cb; t += OPEN_THREADS) { s_key[ca + t] = bk[t]; s_val[ca + t] = bv[t]; }
        __syncthreads();
        const uint32_t cnt = ca + cb;
        const uint32_t my0 = min(threadIdx.x * ITEMS, cnt), my1 = min(my0 + ITEMS, cnt);
        uint32_t pa = merge_path(s_key, ca, s_key + ca, cb, my0);
        uint32_t pb = my0 - pa;
        const uint32_t k = ia.k_pop[i];
        const uint32_t pop_base = pop_off_next[i], open_base = open_off_next[i];
        // each thread merges its ITEMS consecutive outputs into registers ...
        unsigned long long okey[ITEMS];
        uint32_t oval[ITEMS];
#pragma unroll
        for (int o = 0; o < ITEMS; ++o) {
            bool take_a;
            if (pa >= ca) take_a = false;
            else if (pb >= cb) take_a = true;
            else take_a = s_key[pa] <= s_key[ca + pb];
            const uint32_t idx = min(take_a ? pa : ca + pb, (uint32_t)DX_SORT_TILE - 1);
            okey[o] = s_key[idx];
            oval[o] = s_val[idx];
            if (take_a) ++pa; else ++pb;
        }
        __syncthreads();                           // every thread has read its inputs: the tile buffers can take the outputs
#pragma unroll
        for (int o = 0; o < ITEMS; ++o)
            if (my0 + o < my1) { s_key[my0 + o] = okey[o]; s_val[my0 + o] = oval[o]; }
        __syncthreads();
        // ... and the tile is written out with consecutive threads on consecutive entries (a thread writing its own 8 outputs
        // touches a different 32-byte sector per store: 4x / 8x write amplification on the key / value arrays)
        for (uint32_t t = threadIdx.x; t < cnt; t += OPEN_THREADS) {
            const uint32_t r = d0 + t;
            const unsigned long long key = s_key[t];
            const uint32_t val = s_val[t];
            if (r < k) {
                popped_next[pop_base + r] = val;
                popped_inst_next[pop_base + r] = i;
                if (r == 0) ia.f_first[i] = key;
            } else if (keep_open) {            // (beam search drops what it does not select)
                open_key_next[open_base + r - k] = key;
                open_val_next[open_base + r - k] = val;
            }
        }
    }
}

int dx_launch_merge(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const SortBufs& sb,
                    const unsigned long long* open_key_cur, const uint32_t* open_val_cur, const uint32_t* open_off_cur,
                    unsigned long long* open_key_next, uint32_t* open_val_next, const uint32_t* open_off_next,
                    uint32_t* popped_next, uint32_t* popped_inst_next, const uint32_t* pop_off_next, int max_tiles, int keep_open) {
    int blocks = min(max_tiles, 148 * 8);      // 24.5 KB of shared memory and 256 threads per block: 8 blocks per SM
    dx_launch(k_merge_partition, min(dx_div_up(max_tiles, OPEN_THREADS / 32), 148 * 4), OPEN_THREADS, 0, s, ctl, ia, sb, open_key_cur,
              open_off_cur, sb.part_a, sb.part_i, (const uint32_t*)nullptr, (const unsigned long long*)nullptr, (const uint32_t*)nullptr,
              (const uint32_t*)nullptr);
    dx_launch(k_merge, blocks, OPEN_THREADS, 0, s, ctl, ia, sb, open_key_cur, open_val_cur, open_off_cur, open_key_next, open_val_next,
              open_off_next, popped_next, popped_inst_next, pop_off_next, keep_open, sb.part_a, sb.part_i, (const uint32_t*)nullptr,
              (const unsigned long long*)nullptr, (const uint32_t*)nullptr, (const uint32_t*)nullptr, (const uint32_t*)nullptr);
    return 2;
}

// ---------------------------------------------------------------------------------------------
// End of PathFind.step(): lb, counters, finished(), next iteration's sizes.
//   gen_mode 0 (A*): num_nodes_generated += popped nodes * A   (base/pathfinding.py:458-459)
//   gen_mode 1 (Q*): num_nodes_generated += popped edges       (base/pathfinding.py:562-563)
__global__ void __launch_bounds__(1024)
k_end_iter(Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ pop_off_cur,
           const uint32_t* __restrict__ pop_off_next, int A, int gen_mode) {
    __shared__ uint32_t s_unfinished;
    __shared__ unsigned long long s_gen;
    dx_pdl_prologue();
    if (ctl->error) return;
    if (threadIdx.x == 0) { s_unfinished = 0; s_gen = 0; }
    __syncthreads();
    const int I = ctl->n_inst;
    uint32_t unfin = 0;
    unsigned long long gen = 0;
    for (int i = threadIdx.x; i < I; i += blockDim.x) {
        if (!ia.active[i]) continue;
        // all loads, then the stores (the arrays may alias as far as the compiler knows; interleaved they form one long chain)
        const uint32_t n_pop = pop_off_cur[i + 1] - pop_off_cur[i];
        const uint32_t k = ia.k_pop[i];
        double lb = ia.lb[i];
        const unsigned long long ff = ia.f_first[i], gen_old = ia.gen[i], ubk = ia.ub_key[i];
        const uint32_t seq_old = ia.pop_seq_base[i], itr = ia.itr[i] + 1;
        const double w = ia.weight[i];
        if (k > 0) lb = fmax(dx_key_to_f(ff), lb);                                  // graph_search.py:67-69
        const bool edge_mode = (gen_mode & 7) == 1 || (gen_mode & 7) == 3;
        const unsigned long long g_add = !edge_mode ? (unsigned long long)n_pop * A : (unsigned long long)k;
        gen += g_add;
        bool fin = false;
        if (ubk != DX_EMPTY64) {
            const double ub = (double)__uint_as_float((uint32_t)(ubk >> 32));
            fin = (gen_mode & 7) >= 2 ? true                                               // beam search: finished() == has_soln() (beam_search.py:62-66)
                                : lb >= __dmul_rn(w, ub);                           // graph_search.py:44
        }
        fin = fin || (k == 0);                                                       // itr > 0 holds here (graph_search.py:45)
        ia.lb[i] = lb;
        ia.gen[i] = gen_old + g_add;
        ia.pop_seq_base[i] = seq_old + n_pop;
        ia.itr[i] = itr;
        ia.active[i] = fin ? 0 : 1;
        if (!fin) ++unfin;
    }
    if (unfin) atomicAdd(&s_unfinished, unfin);
    if (gen) atomicAdd(&s_gen, gen);
    __syncthreads();
    if (threadIdx.x == 0) {
        ctl->n_unfinished = s_unfinished;
        ctl->gen_total += s_gen;
        ctl->n_pop = pop_off_next[I];
        ctl->n_children = pop_off_next[I] * (((gen_mode & 7) == 0 || (gen_mode & 7) == 2) ? (uint32_t)A : 1u);   // items the next CLOSED filter sees
        if ((gen_mode & DX_END_HALT) && s_unfinished == 0) ctl->error |= DX_DEVERR_DONE;        // the rest of dx_run's burst is a no-op
        ctl->itr += 1;
        ctl->epoch += 1;
        ctl->node_base = ctl->node_count;
    }
}

int dx_launch_end_iter(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* pop_off_cur, const uint32_t* pop_off_next,
                       int A, int gen_mode) {
    dx_launch(k_end_iter, 1, 1024, 0, s, ctl, ia, pop_off_cur, pop_off_next, A, gen_mode);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// Young-run OPEN for small-batch A* engines (YoungBufs, dx_internal.cuh): the OPEN side of an iteration in ONE block.
//   plan   per instance: pushed m, main-run remainder nm, young-run remainder ny, k = min(batch, nm + ny + m)   (k_plan)
//   keys   f = w * g + h of the stored children                                                                (k_make_keys_v)
//   sort   bitonic network on (instance, key, position) in shared memory                                       (k_sort_small)
//   merge  young run + batch by RANKING: an element's output position is its own index plus the number of elements of the
//          other run that precede it (young before batch on ties: older push first) -- one binary search per element, no
//          serial merge; the result stays in shared memory and goes back to the other young buffer
//   pop    k smallest of (main run head, merged young run): a warp finds the merge-path split (ka, kp), the k outputs are
//          ranked into the popped list the same way (main run first on ties); the two runs' heads advance by ka / kp
//   book   lb / ub / finished() / counters                                                                    (k_end_iter)
#define YT_THREADS 1024
#define YT_MAX_INST 256
// -DYT_PROFILE: clock64 timeline of the block (thread 0), accumulated over launches: [0] launches, [1] plan, [2] staging + keys,
// [3] sort, [4] merge by ranking, [5] pop split, [6] pop ranking, [7] write-back, [8] bookkeeping (read and cleared by dx_debug_tc_profile
// when DXB200_PROF_TAIL is set)
__device__ unsigned long long g_tail_prof[16];
#ifdef YT_PROFILE
#define YT_MARK(slot) do { if (threadIdx.x == 0) { const long long _t = clock64(); \
        atomicAdd(&g_tail_prof[(slot)], (unsigned long long)(_t - t_prev)); t_prev = _t; } } while (0)
#else
#define YT_MARK(slot) do { } while (0)
#endif
int dx_tail_profile_read(unsigned long long* out16) {
    unsigned long long z[16] = {0};
#ifdef OS_PROFILE
    if (cudaMemcpyFromSymbol(out16, g_os_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return DX_ERR_CUDA;
    if (cudaMemcpyToSymbol(g_os_prof, z, sizeof(z)) != cudaSuccess) return DX_ERR_CUDA;
#else
    if (cudaMemcpyFromSymbol(out16, g_tail_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return DX_ERR_CUDA;
    if (cudaMemcpyToSymbol(g_tail_prof, z, sizeof(z)) != cudaSuccess) return DX_ERR_CUDA;
#endif
    return DX_OK;
}

__device__ __forceinline__ uint32_t yt_lower_bound(const unsigned long long* a, uint32_t n, unsigned long long key) {   // # a[j] < key
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ uint32_t yt_upper_bound(const unsigned long long* a, uint32_t n, unsigned long long key) {   // # a[j] <= key
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a[mid] <= key) lo = mid + 1; else hi = mid; }
    return lo;
}
// last i in [0, n) with off[i] <= x  (off ascending, off[0] = 0 <= x)
__device__ __forceinline__ int yt_find(const uint32_t* off, int n, uint32_t x) {
    int lo = 0, hi = n - 1;
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off[mid] <= x) lo = mid; else hi = mid - 1; }
    return lo;
}

// Sort of <= 2048 (instance, key, position) triples in shared memory without a 66-stage block-wide bitonic network (which cost
// 450 cycles per stage: 15 us of a 25 us kernel): every warp sorts a 64-element chunk in REGISTERS (two elements per lane, 21
// compare-exchange stages through shuffles), then two rounds of merging BY RANKING -- an element's position in the merged run is
// its own index plus, for every other run of its group, the number of that run's elements below it (binary searches in shared
// memory; the order is total because positions are unique, so no tie rule is needed): 64-runs -> 256-runs -> the whole batch.
struct YtElem { uint32_t inst; unsigned long long key; uint32_t idx; };
__device__ __forceinline__ bool yt_less(const YtElem& a, const YtElem& b) {
    return a.inst != b.inst ? a.inst < b.inst : (a.key != b.key ? a.key < b.key : a.idx < b.idx);
}
__device__ __forceinline__ YtElem yt_shfl_xor(const YtElem& a, int j) {
    YtElem o;
    o.inst = __shfl_xor_sync(0xffffffffu, a.inst, j);
    o.key = __shfl_xor_sync(0xffffffffu, a.key, j);
    o.idx = __shfl_xor_sync(0xffffffffu, a.idx, j);
    return o;
}
// # elements of the run [key, inst, idx][0 .. len) below x
__device__ __forceinline__ uint32_t yt_rank_in(const unsigned long long* key, const uint32_t* inst, const uint32_t* idx, uint32_t len,
                                               const YtElem& x) {
    uint32_t lo = 0, hi = len;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        YtElem m; m.inst = inst[mid]; m.key = key[mid]; m.idx = idx[mid];
        if (yt_less(m, x)) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// one round: runs of `len` elements, merged in groups of `grp` runs (total elements: multiple of 64, all threads of the block call)
__device__ __forceinline__ void yt_rank_merge(const unsigned long long* ikey, const uint32_t* iinst, const uint32_t* iidx,
                                              unsigned long long* okey, uint32_t* oinst, uint32_t* oidx, uint32_t total, uint32_t len,
                                              uint32_t grp) {
    for (uint32_t e = threadIdx.x; e < total; e += blockDim.x) {
        const uint32_t r = e / len, g = r / grp, l = e - r * len;
        YtElem x; x.inst = iinst[e]; x.key = ikey[e]; x.idx = iidx[e];
        uint32_t pos = g * grp * len + l;
        for (uint32_t r2 = g * grp; r2 < (g + 1) * grp; ++r2) {
            const uint32_t b2 = r2 * len;
            if (r2 == r || b2 >= total) continue;
            pos += yt_rank_in(ikey + b2, iinst + b2, iidx + b2, min(len, total - b2), x);
        }
        okey[pos] = x.key; oinst[pos] = x.inst; oidx[pos] = x.idx;
    }
}
// s_key / s_inst / s_idx [0 .. total) (total = n rounded up to 64, padding = (0xFFFFFFFF, ~0, position)) -> sorted in place;
// t_key / t_inst / t_idx: scratch of the same size
__device__ __forceinline__ void yt_sort(unsigned long long* s_key, uint32_t* s_inst, uint32_t* s_idx, unsigned long long* t_key,
                                        uint32_t* t_inst, uint32_t* t_idx, uint32_t total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cbase = (uint32_t)warp * 64;
    if (cbase < total) {                              // warp-uniform: this warp's chunk exists
        YtElem el[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) { const uint32_t e = cbase + lane + 32 * r; el[r].inst = s_inst[e]; el[r].key = s_key[e]; el[r].idx = s_idx[e]; }
#pragma unroll
        for (int k = 2; k <= 64; k <<= 1)
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                if (j == 32) {                        // partner = the lane's own second element; k = 64: ascending everywhere
                    if (yt_less(el[1], el[0])) { const YtElem t = el[0]; el[0] = el[1]; el[1] = t; }
                } else {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                        const int e = lane + 32 * r;
                        const YtElem o = yt_shfl_xor(el[r], j);
                        const bool up = (e & k) == 0, lower = (lane & j) == 0;
                        const bool take_min = up == lower;
                        const bool o_less = yt_less(o, el[r]);
                        if (take_min == o_less) el[r] = o;
                    }
                }
            }
#pragma unroll
        for (int r = 0; r < 2; ++r) { const uint32_t e = cbase + lane + 32 * r; s_inst[e] = el[r].inst; s_key[e] = el[r].key; s_idx[e] = el[r].idx; }
    }
    __syncthreads();
    if (total <= 64) return;
    yt_rank_merge(s_key, s_inst, s_idx, t_key, t_inst, t_idx, total, 64, 4);             // -> runs of 256
    __syncthreads();
    yt_rank_merge(t_key, t_inst, t_idx, s_key, s_inst, s_idx, total, 256, 8);            // -> one run (total <= 2048)
    __syncthreads();
}

// The same sort as a stable LSD radix sort in shared memory, taken when at most YT_RADIX_MAX_DIGITS digits vary (returns false
// otherwise and the caller runs the comparison sort above): both blocks of the small iteration are bound by instruction issue on
// ONE SM, and a radix pass over low-entropy digits costs ~160 instructions per thread.  A pass: per-warp digit counts through
// __match_any_sync (a warp owns 64 consecutive elements, two rows of 32: ranks stay in position order, so the sort is stable and
// the position never has to be compared), a block scan over (digit, warp), scatter into the other buffer.  Bytes that are the same
// in every key (the top bytes of a narrow f range) and the instance byte of single-instance batches are skipped.
// s_cnt: [256][32] bytes, s_base: [256][32] u16, s_red: [8 + 32] words of scratch.
#define YT_RADIX_MAX_DIGITS 3
__device__ __forceinline__ bool yt_radix_sort(unsigned long long* s_key, uint32_t* s_inst, uint32_t* s_idx, unsigned long long* t_key,
                                              uint32_t* t_inst, uint32_t* t_idx, uint8_t* s_cnt, uint16_t* s_base, uint32_t* s_red,
                                              uint32_t n) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t e0 = (uint32_t)warp * 64 + lane, e1 = e0 + 32;
    {   // which digits vary
        uint32_t o_lo = 0, o_hi = 0, a_lo = ~0u, a_hi = ~0u, o_in = 0, a_in = ~0u;
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t e = r ? e1 : e0;
            if (e < n) {
                const unsigned long long k = s_key[e];
                const uint32_t in = s_inst[e];
                o_lo |= (uint32_t)k; o_hi |= (uint32_t)(k >> 32); a_lo &= (uint32_t)k; a_hi &= (uint32_t)(k >> 32);
                o_in |= in; a_in &= in;
            }
        }
        o_lo = __reduce_or_sync(0xffffffffu, o_lo); o_hi = __reduce_or_sync(0xffffffffu, o_hi); o_in = __reduce_or_sync(0xffffffffu, o_in);
        a_lo = __reduce_and_sync(0xffffffffu, a_lo); a_hi = __reduce_and_sync(0xffffffffu, a_hi); a_in = __reduce_and_sync(0xffffffffu, a_in);
        if (tid < 6) s_red[tid] = tid < 3 ? 0u : ~0u;
        __syncthreads();
        if (lane == 0) {
            atomicOr(&s_red[0], o_lo); atomicOr(&s_red[1], o_hi); atomicOr(&s_red[2], o_in);
            atomicAnd(&s_red[3], a_lo); atomicAnd(&s_red[4], a_hi); atomicAnd(&s_red[5], a_in);
        }
        __syncthreads();
    }
    const uint32_t diff_lo = s_red[0] ^ s_red[3], diff_hi = s_red[1] ^ s_red[4], diff_in = s_red[2] ^ s_red[5];
    __syncthreads();                                  // (s_red[8..] is reused by the scans below)
    {   // __match_any_sync walks the distinct digit values of a warp one by one: with high-entropy bytes (the low mantissa bytes of
        // w * g + h with a network h) a pass costs ~5 k cycles and eight of them lose to the comparison sort (measured: KAT
        // iteration 67.5 -> 76.3 us), while keys with few varying bytes (zero / coarse heuristics) win (53.0 -> 45.5 us)
        int nvar = (diff_in & 0xFFu) ? 1 : 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) nvar += (((diff_lo >> (8 * b)) & 0xFFu) ? 1 : 0) + (((diff_hi >> (8 * b)) & 0xFFu) ? 1 : 0);
        if (nvar > YT_RADIX_MAX_DIGITS) return false;
    }
    unsigned long long *ksrc = s_key, *kdst = t_key;
    uint32_t *isrc = s_inst, *idst = t_inst, *xsrc = s_idx, *xdst = t_idx;
    for (int b = 0; b < 9; ++b) {                     // key bytes 0..7, then the instance byte (instances < 256 in this regime)
        const uint32_t varies = b < 4 ? (diff_lo >> (8 * b)) & 0xFFu : (b < 8 ? (diff_hi >> (8 * (b - 4))) & 0xFFu : (diff_in & 0xFFu));
        if (!varies) continue;                        // (block-uniform)
        for (int i = tid; i < 256 * 32 / 4; i += YT_THREADS) reinterpret_cast<uint32_t*>(s_cnt)[i] = 0u;
        __syncthreads();
        unsigned long long k[2];
        uint32_t in[2], ix[2], dg[2], lr[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const uint32_t e = r ? e1 : e0;
            const bool valid = e < n;
            k[r] = valid ? ksrc[e] : 0ull;
            in[r] = valid ? isrc[e] : 0u;
            ix[r] = valid ? xsrc[e] : 0u;
            const uint32_t d = b < 8 ? (uint32_t)(k[r] >> (8 * b)) & 0xFFu : (in[r] & 0xFFu);
            dg[r] = valid ? d : (256u + (uint32_t)lane);
            const uint32_t peers = __match_any_sync(0xffffffffu, dg[r]);
            const int leader = __ffs(peers) - 1;
            const uint32_t before = __popc(peers & ((1u << lane) - 1u));
            uint32_t old = 0;
            if (valid && lane == leader) {
                old = s_cnt[d * 32 + warp];
                s_cnt[d * 32 + warp] = (uint8_t)(old + __popc(peers));
            }
            old = __shfl_sync(0xffffffffu, old, leader);
            lr[r] = old + before;
            __syncwarp();
        }
        __syncthreads();
        {   // exclusive scan over (digit, warp): thread t owns digit t / 4, warps 8 * (t % 4) .. + 7
            const uint2 c8 = reinterpret_cast<const uint2*>(s_cnt)[tid];
            uint32_t c[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) { c[j] = (c8.x >> (8 * j)) & 0xFFu; c[4 + j] = (c8.y >> (8 * j)) & 0xFFu; }
            uint32_t sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += c[j];
            uint32_t x = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
            if (lane == 31) s_red[8 + warp] = x;
            __syncthreads();
            if (warp == 0) {
                uint32_t w = s_red[8 + lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
                s_red[8 + lane] = w;
            }
            __syncthreads();
            uint32_t run = (warp ? s_red[8 + warp - 1] : 0u) + x - sum;
#pragma unroll
            for (int j = 0; j < 8; ++j) { s_base[tid * 8 + j] = (uint16_t)run; run += c[j]; }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            if (dg[r] < 256u) {
                const uint32_t pos = (uint32_t)s_base[dg[r] * 32 + warp] + lr[r];
                kdst[pos] = k[r]; idst[pos] = in[r]; xdst[pos] = ix[r];
            }
        }
        __syncthreads();
        { unsigned long long* t = ksrc; ksrc = kdst; kdst = t; }
        { uint32_t* t = isrc; isrc = idst; idst = t; }
        { uint32_t* t = xsrc; xsrc = xdst; xdst = t; }
    }
    if (ksrc != s_key) {
        for (uint32_t e = tid; e < n; e += YT_THREADS) { s_key[e] = ksrc[e]; s_inst[e] = isrc[e]; s_idx[e] = xsrc[e]; }
        __syncthreads();
    }
    return true;
}

__global__ void __launch_bounds__(YT_THREADS)
k_small_tail_v(Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ pop_off_cur,
               uint32_t* __restrict__ pop_off_next, const unsigned long long* __restrict__ m_key, const uint32_t* __restrict__ m_val,
               const uint32_t* __restrict__ m_off, YoungBufs yb, int p, int A, uint32_t max_pop, const float* __restrict__ node_g,
               const float* __restrict__ node_h, const uint32_t* __restrict__ new_inst, uint32_t* __restrict__ popped_next,
               uint32_t* __restrict__ popped_inst_next, int gen_mode) {
    extern __shared__ __align__(16) uint8_t yt_raw[];
    unsigned long long* s_bkey = reinterpret_cast<unsigned long long*>(yt_raw);               // [2048] batch keys
    unsigned long long* s_ykey = s_bkey + DX_SMALL_SORT;                                      // [DX_YOUNG_BUF] merged young run
    uint32_t* s_binst = reinterpret_cast<uint32_t*>(s_ykey + DX_YOUNG_BUF);                   // [2048]
    uint32_t* s_bidx = s_binst + DX_SMALL_SORT;                                               // [2048]
    uint32_t* s_yval = s_bidx + DX_SMALL_SORT;                                                // [DX_YOUNG_BUF]
    unsigned long long* s_yin = reinterpret_cast<unsigned long long*>(s_yval + DX_YOUNG_BUF); // [DX_YOUNG_BUF] keys of the young remainders
    unsigned long long* s_mkey = s_yin + DX_YOUNG_BUF;                                        // [2048] keys of the main-run heads (first k)
    uint32_t* s_mval = reinterpret_cast<uint32_t*>(s_mkey + DX_SMALL_SORT);                   // [2048]
    uint32_t* s_boff = s_mval + DX_SMALL_SORT;       // [I + 1] batch offsets
    uint32_t* s_yoff = s_boff + YT_MAX_INST + 1;     // [I + 1] offsets of the young-run remainders (flat index space)
    uint32_t* s_ysrc = s_yoff + YT_MAX_INST + 1;     // [I]     where the remainder starts in the current young buffer
    uint32_t* s_ybo = s_ysrc + YT_MAX_INST + 1;      // [I + 1] offsets of the merged runs (next young buffer)
    uint32_t* s_nm = s_ybo + YT_MAX_INST + 1;        // [I]     main-run remainder
    uint32_t* s_msrc = s_nm + YT_MAX_INST + 1;       // [I]     where it starts
    uint32_t* s_poff = s_msrc + YT_MAX_INST + 1;     // [I + 1] popped-list offsets
    uint32_t* s_ka = s_poff + YT_MAX_INST + 1;       // [I]     pops from the main run
    __shared__ uint32_t s_err, s_unfinished;
    __shared__ unsigned long long s_gen;
    dx_pdl_prologue();
#ifdef YT_PROFILE
    long long t_prev = clock64();
    if (threadIdx.x == 0) atomicAdd(&g_tail_prof[0], 1ull);
#endif
    if (ctl->error) return;
    const int tid = threadIdx.x, q = p ^ 1;
    const int I = (int)ctl->n_inst;
    const uint32_t node_base = ctl->node_base;
    if (tid == 0) { s_err = 0; s_unfinished = 0; s_gen = 0; }
    if (I > YT_MAX_INST) { if (tid == 0) atomicOr(&ctl->error, DX_DEVERR_POP); return; }
    // ---- plan (serial over the instances: I is small in this regime) -------------------------------------------------
    if (tid == 0) {
        uint32_t sm = 0, sy = 0, syb = 0, sk = 0;
        for (int i = 0; i < I; ++i) {
            const bool act = ia.active[i] != 0;
            const uint32_t kept = prefix[pop_off_cur[i + 1] * A] - prefix[pop_off_cur[i] * A];
            const uint32_t m = act ? kept : 0u;
            const uint32_t mh = yb.m_head[i];
            const uint32_t nm = act ? (m_off[i + 1] - m_off[i]) - mh : 0u;        // (OPEN of a finished instance is dropped)
            const uint32_t yh = yb.head[p][i];
            const uint32_t ny = act ? (yb.off[p][i + 1] - yb.off[p][i]) - yh : 0u;
            const uint32_t t = nm + ny + m;
            const uint32_t k = act ? min((uint32_t)ia.batch[i], t) : 0u;
            s_boff[i] = sm; s_yoff[i] = sy; s_ybo[i] = syb; s_poff[i] = sk;
            s_ysrc[i] = yb.off[p][i] + yh; s_nm[i] = nm; s_msrc[i] = m_off[i] + mh;
            ia.m_new[i] = m; ia.k_pop[i] = k; ia.n_open[i] = t - k;
            sm += m; sy += ny; syb += ny + m; sk += k;
        }
        s_boff[I] = sm; s_yoff[I] = sy; s_ybo[I] = syb; s_poff[I] = sk;
        uint32_t err = 0;
        if (sm > DX_SMALL_SORT || sk > max_pop || sk > DX_SMALL_SORT) err |= DX_DEVERR_POP;
        if (syb > DX_YOUNG_BUF) err |= DX_DEVERR_OPEN;                           // (the host compacts before this can happen)
        if (err) { atomicOr(&ctl->error, err); s_err = err; }
        ctl->sort_n = sm;
    }
    __syncthreads();
    if (s_err) return;
    YT_MARK(1);
    const uint32_t n = s_boff[I], n_young = s_yoff[I], n_yb = s_ybo[I], n_popk = s_poff[I];
    // stage what the searches below read: the young remainders' keys and the first k entries of every main run (one round trip,
    // issued before the key computation; every binary search of the merge and the pop then stays in shared memory)
    for (uint32_t e = tid; e < n_young; e += YT_THREADS) {
        const int i = yt_find(s_yoff, I, e);
        s_yin[e] = yb.key[p][s_ysrc[i] + (e - s_yoff[i])];
    }
    for (uint32_t r = tid; r < n_popk; r += YT_THREADS) {
        const int i = yt_find(s_poff, I, r);
        const uint32_t rr = r - s_poff[i];
        const bool in = rr < s_nm[i];
        s_mkey[r] = in ? m_key[s_msrc[i] + rr] : ~0ull;
        s_mval[r] = in ? m_val[s_msrc[i] + rr] : 0u;
    }
    // ---- keys of the stored children, then the sort on (instance, key, position) ----------------------------------------
    const uint32_t total = (n + 63u) & ~63u;
    for (uint32_t k = tid; k < total; k += YT_THREADS) {
        if (k < n) {
            const uint32_t id = node_base + k, inst = new_inst[k];
            s_bkey[k] = f_key(ia.weight[inst], node_g[id], node_h[id]);
            s_binst[k] = inst;
        } else {
            s_bkey[k] = ~0ull; s_binst[k] = 0xFFFFFFFFu;
        }
        s_bidx[k] = k;
    }
    __syncthreads();
    YT_MARK(2);
    if (n > 1) {    // radix scratch: the merged-run buffers (not in use yet): keys [0, 2048), counters / bases / scan words behind them
        const bool done = yt_radix_sort(s_bkey, s_binst, s_bidx, s_ykey, s_yval, s_yval + DX_SMALL_SORT,
                                        reinterpret_cast<uint8_t*>(s_ykey + DX_SMALL_SORT), reinterpret_cast<uint16_t*>(s_ykey + DX_SMALL_SORT + 1024),
                                        reinterpret_cast<uint32_t*>(s_ykey + DX_SMALL_SORT + 1024 + 2048), n);
        if (!done) yt_sort(s_bkey, s_binst, s_bidx, s_ykey, s_yval, s_yval + DX_SMALL_SORT, total);      // (block-uniform decision)
    }
    YT_MARK(3);
    // ---- merged young run = young remainder + batch, by ranking --------------------------------------------------------
    const uint32_t* __restrict__ yval_cur = yb.val[p];
    for (uint32_t j = tid; j < n; j += YT_THREADS) {                             // batch elements: young entries <= key precede
        const int i = (int)s_binst[j];
        const unsigned long long key = s_bkey[j];
        const uint32_t b = j - s_boff[i], ny = s_yoff[i + 1] - s_yoff[i];
        const uint32_t pos = s_ybo[i] + b + yt_upper_bound(s_yin + s_yoff[i], ny, key);
        s_ykey[pos] = key;
        s_yval[pos] = node_base + s_bidx[j];
    }
    for (uint32_t e = tid; e < n_young; e += YT_THREADS) {                       // young entries: batch entries < key precede
        const int i = yt_find(s_yoff, I, e);                                     // (instances with ny = 0 share an offset: the LAST wins,
        const uint32_t a = e - s_yoff[i];                                        //  which is the one that owns e)
        const unsigned long long key = s_yin[e];
        const uint32_t pos = s_ybo[i] + a + yt_lower_bound(s_bkey + s_boff[i], s_boff[i + 1] - s_boff[i], key);
        s_ykey[pos] = key;
        s_yval[pos] = yval_cur[s_ysrc[i] + a];
    }
    __syncthreads();
    YT_MARK(4);
    // ---- pop split per instance (a warp each): ka entries of the main run among the k smallest ----------------------------
    for (int i = tid >> 5; i < I; i += YT_THREADS / 32) {
        const uint32_t k = s_poff[i + 1] - s_poff[i];
        const uint32_t nyb = s_ybo[i + 1] - s_ybo[i];
        const uint32_t ka = merge_path_warp(s_mkey + s_poff[i], min(s_nm[i], k), s_ykey + s_ybo[i], min(nyb, k), k);
        if ((tid & 31) == 0) s_ka[i] = ka;
    }
    __syncthreads();
    YT_MARK(5);
    // ---- the popped list (pop order), by ranking; main run first on ties ------------------------------------------------
    for (uint32_t r = tid; r < n_popk; r += YT_THREADS) {
        const int i = yt_find(s_poff, I, r);
        const uint32_t rr = r - s_poff[i], ka = s_ka[i], kp = (s_poff[i + 1] - s_poff[i]) - ka;
        unsigned long long key;
        uint32_t val, pos;
        if (rr < ka) {
            key = s_mkey[s_poff[i] + rr];
            val = s_mval[s_poff[i] + rr];
            pos = rr + yt_lower_bound(s_ykey + s_ybo[i], kp, key);
        } else {
            const uint32_t b = rr - ka;
            key = s_ykey[s_ybo[i] + b];
            val = s_yval[s_ybo[i] + b];
            pos = b + yt_upper_bound(s_mkey + s_poff[i], ka, key);
        }
        popped_next[s_poff[i] + pos] = val;
        popped_inst_next[s_poff[i] + pos] = (uint32_t)i;
        if (pos == 0) ia.f_first[i] = key;
    }
    YT_MARK(6);
    // ---- the merged young runs go to the other buffer; heads advance ------------------------------------------------------
    for (uint32_t e = tid; e < n_yb; e += YT_THREADS) { yb.key[q][e] = s_ykey[e]; yb.val[q][e] = s_yval[e]; }
    for (int i = tid; i <= I; i += YT_THREADS) {
        yb.off[q][i] = s_ybo[i];
        pop_off_next[i] = s_poff[i];
        if (i < I) {
            const uint32_t ka = s_ka[i];
            yb.head[q][i] = (s_poff[i + 1] - s_poff[i]) - ka;
            yb.m_head[i] += ka;
        }
    }
    __syncthreads();                                  // f_first of every instance is written (same block: visible)
    YT_MARK(7);
    // ---- lb / ub / finished() / counters (k_end_iter; graph_search.py:43-46, 67-69) ------------------------------------------
    uint32_t unfin = 0;
    unsigned long long gen = 0;
    for (int i = tid; i < I; i += YT_THREADS) {
        if (!ia.active[i]) continue;
        const uint32_t n_pop = pop_off_cur[i + 1] - pop_off_cur[i];
        const uint32_t k = s_poff[i + 1] - s_poff[i];
        double lb = ia.lb[i];
        const unsigned long long ff = ia.f_first[i], gen_old = ia.gen[i], ubk = ia.ub_key[i];
        const uint32_t seq_old = ia.pop_seq_base[i], itr = ia.itr[i] + 1;
        const double w = ia.weight[i];
        if (k > 0) lb = fmax(dx_key_to_f(ff), lb);
        const unsigned long long g_add = (unsigned long long)n_pop * A;
        gen += g_add;
        bool fin = false;
        if (ubk != DX_EMPTY64) {
            const double ub = (double)__uint_as_float((uint32_t)(ubk >> 32));
            fin = lb >= __dmul_rn(w, ub);
        }
        fin = fin || (k == 0);
        ia.lb[i] = lb;
        ia.gen[i] = gen_old + g_add;
        ia.pop_seq_base[i] = seq_old + n_pop;
        ia.itr[i] = itr;
        ia.active[i] = fin ? 0 : 1;
        if (!fin) ++unfin;
    }
    if (unfin) atomicAdd(&s_unfinished, unfin);
    if (gen) atomicAdd(&s_gen, gen);
    __syncthreads();
    if (tid == 0) {
        ctl->n_unfinished = s_unfinished;
        ctl->gen_total += s_gen;
        ctl->n_pop = n_popk;
        ctl->n_children = n_popk * (uint32_t)A;
        ctl->itr += 1;
        ctl->epoch += 1;
        ctl->node_base = ctl->node_count;
        ctl->young_total = n_yb;
        if ((gen_mode & DX_END_HALT) && s_unfinished == 0) ctl->error |= DX_DEVERR_DONE;
    }
    YT_MARK(8);
}

static size_t young_tail_smem() {
    return (size_t)DX_SMALL_SORT * (8 + 4 + 4) + (size_t)DX_YOUNG_BUF * (8 + 4) + (size_t)DX_YOUNG_BUF * 8 + (size_t)DX_SMALL_SORT * (8 + 4) +
           (size_t)8 * (YT_MAX_INST + 1) * 4;
}

int dx_launch_small_tail_v(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* child_prefix, const uint32_t* pop_off_cur,
                           uint32_t* pop_off_next, const unsigned long long* m_key, const uint32_t* m_val, const uint32_t* m_off,
                           const YoungBufs& yb, int p, int A, uint32_t max_pop, const float* node_g, const float* node_h,
                           const uint32_t* new_inst, uint32_t* popped_next, uint32_t* popped_inst_next, int gen_mode) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_small_tail_v, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)young_tail_smem()) != cudaSuccess) {
            dx_set_error("k_small_tail_v: cannot raise the dynamic shared-memory limit");
            return 0;
        }
        attr_set = true;
    }
    dx_launch(k_small_tail_v, 1, YT_THREADS, young_tail_smem(), s, ctl, ia, child_prefix, pop_off_cur, pop_off_next, m_key, m_val, m_off, yb,
              p, A, max_pop, node_g, node_h, new_inst, popped_next, popped_inst_next, gen_mode);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// Two-run OPEN for BIG batches (engines with <= YT_MAX_INST instance slots): the same state as above (main run consumed through
// m_head, young run of what was pushed since the last compaction), handled by the grid-wide merge kernels.  k_merge rewrote every
// instance's whole run each iteration -- O(|OPEN|) traffic for an O(batch) change, 12 % of a cube3 Q* step with a 25 M-entry OPEN;
// now an iteration moves O(|young| + batch) entries and the main run is only rewritten at the host-scheduled compactions.
//   k_young_plan_b   per instance: pushed m, remainders nm / ny, k = min(batch, nm + ny + m); offsets of the batch, the merged young
//                    runs and the popped lists; tile offsets of the two merges                                   (k_plan)
//   merge 1          young[q] <- merge(young[p] behind its head, sorted batch), nothing popped      (k_merge_partition + k_merge)
//   k_young_split    a warp per instance: ka = how many of the k smallest of (main-run head, merged young run) are main-run entries
//                    (main run first on ties: pushed earlier); heads advance by ka / k - ka
//   merge 2          popped list <- merge(main[m_head .. m_head + ka), young[q][0 .. k - ka))          (k_merge_partition + k_merge)
__global__ void __launch_bounds__(YT_MAX_INST)
k_young_plan_b(Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ prefix, const uint32_t* __restrict__ pop_off_cur,
               uint32_t* __restrict__ pop_off_next, const uint32_t* __restrict__ m_off, YoungBufs yb, int p, int child_mult,
               int push_mult, uint32_t max_pop) {
    __shared__ uint32_t s_scan[5][YT_MAX_INST + 1];
    dx_pdl_prologue();
    if (ctl->error) return;
    const int I = (int)ctl->n_inst, i = threadIdx.x, q = p ^ 1;
    if (I > YT_MAX_INST) { if (i == 0) atomicOr(&ctl->error, DX_DEVERR_POP); return; }
    uint32_t m = 0, ny = 0, k = 0, yh = 0;
    if (i < I) {
        const bool act = ia.active[i] != 0;
        const uint32_t kept = prefix[pop_off_cur[i + 1] * child_mult] - prefix[pop_off_cur[i] * child_mult];
        m = act ? kept * (uint32_t)push_mult : 0u;
        const uint32_t nm = act ? (m_off[i + 1] - m_off[i]) - yb.m_head[i] : 0u;          // (OPEN of a finished instance is dropped)
        yh = yb.head[p][i];
        ny = act ? (yb.off[p][i + 1] - yb.off[p][i]) - yh : 0u;
        const uint32_t t = nm + ny + m;
        k = act ? min((uint32_t)ia.batch[i], t) : 0u;
        ia.m_new[i] = m; ia.k_pop[i] = k; ia.n_open[i] = t - k;
        // merge 1 reads the young remainder through (off[p], head = total - remainder): a dropped run has remainder 0
        yb.head[p][i] = (yb.off[p][i + 1] - yb.off[p][i]) - ny;
        yb.zeros[i] = 0;
    }
    s_scan[0][i] = m; s_scan[1][i] = ny + m; s_scan[2][i] = k;
    s_scan[3][i] = (ny + m + DX_SORT_TILE - 1) / DX_SORT_TILE; s_scan[4][i] = (k + DX_SORT_TILE - 1) / DX_SORT_TILE;
    __syncthreads();
    if (i < 5) {                                     // five short serial scans (I <= 256), one thread each
        uint32_t acc = 0;
        for (int t = 0; t < I; ++t) { const uint32_t v = s_scan[i][t]; s_scan[i][t] = acc; acc += v; }
        s_scan[i][I] = acc;
    }
    __syncthreads();
    for (int j = i; j <= I; j += YT_MAX_INST) {       // (I + 1 entries: thread 0 also writes the last one when I = 256)
        ia.batch_off[j] = s_scan[0][j];
        yb.off[q][j] = s_scan[1][j];
        pop_off_next[j] = s_scan[2][j];
        yb.tile_off1[j] = s_scan[3][j];
        yb.tile_off2[j] = s_scan[4][j];
    }
    if (i == 0) {
        ctl->sort_n = s_scan[0][I];
        ctl->young_total = s_scan[1][I];
        yb.n_tiles[0] = s_scan[3][I];
        yb.n_tiles[1] = s_scan[4][I];
        uint32_t err = 0;
        if (s_scan[2][I] > max_pop) err |= DX_DEVERR_POP;
        if (s_scan[1][I] > yb.cap) err |= DX_DEVERR_OPEN;                        // (the host compacts before this can happen)
        if (err) atomicOr(&ctl->error, err);
    }
}

__global__ void __launch_bounds__(256)
k_young_split(const Ctl* __restrict__ ctl, InstArrays ia, const unsigned long long* __restrict__ m_key, const uint32_t* __restrict__ m_off,
              YoungBufs yb, int q) {
    dx_pdl_prologue();
    if (ctl->error) return;
    const int I = (int)ctl->n_inst;
    for (int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); i < I; i += gridDim.x * (blockDim.x / 32)) {
        const uint32_t k = ia.k_pop[i], mh = yb.m_head[i];
        const uint32_t nm = ia.active[i] ? (m_off[i + 1] - m_off[i]) - mh : 0u;
        const uint32_t nyb = yb.off[q][i + 1] - yb.off[q][i];
        const uint32_t ka = k ? merge_path_warp(m_key + m_off[i] + mh, min(nm, k), yb.key[q] + yb.off[q][i], min(nyb, k), k) : 0u;
        if ((threadIdx.x & 31) == 0) {
            yb.ka[i] = ka; yb.kp[i] = k - ka; yb.mh_old[i] = mh;
            yb.m_head[i] = mh + ka;
            yb.head[q][i] = k - ka;
        }
    }
}

int dx_launch_young_plan_b(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const uint32_t* child_prefix, const uint32_t* pop_off_cur,
                           uint32_t* pop_off_next, const uint32_t* m_off, const YoungBufs& yb, int p, int child_mult, int push_mult,
                           uint32_t max_pop) {
    dx_launch(k_young_plan_b, 1, YT_MAX_INST, 0, s, ctl, ia, child_prefix, pop_off_cur, pop_off_next, m_off, yb, p, child_mult, push_mult,
              max_pop);
    return 1;
}

int dx_launch_young_iter(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const SortBufs& sb, const unsigned long long* m_key,
                         const uint32_t* m_val, const uint32_t* m_off, const YoungBufs& yb, int p, uint32_t* popped_next,
                         uint32_t* popped_inst_next, const uint32_t* pop_off_next, int max_pop) {
    const int q = p ^ 1;
    const unsigned long long* no_keys = nullptr;
    const uint32_t* no_u32 = nullptr;
    // merge 1: young[q] <- young[p] remainder + sorted batch; k = 0, so everything lands in the "next run"
    InstArrays ia1 = ia;
    ia1.k_pop = yb.zeros;
    ia1.tile_off = yb.tile_off1;
    const int tiles1 = dx_div_up((long long)yb.cap, DX_SORT_TILE) + YT_MAX_INST;
    dx_launch(k_merge_partition, min(dx_div_up(tiles1, OPEN_THREADS / 32), 148 * 4), OPEN_THREADS, 0, s, (const Ctl*)ctl, ia1, sb,
              (const unsigned long long*)yb.key[p], (const uint32_t*)yb.off[p], sb.part_a, sb.part_i, (const uint32_t*)yb.head[p], no_keys,
              no_u32, (const uint32_t*)yb.n_tiles);
    dx_launch(k_merge, min(tiles1, 148 * 8), OPEN_THREADS, 0, s, (const Ctl*)ctl, ia1, sb, (const unsigned long long*)yb.key[p],
              (const uint32_t*)yb.val[p], (const uint32_t*)yb.off[p], yb.key[q], yb.val[q], (const uint32_t*)yb.off[q], popped_next,
              popped_inst_next, pop_off_next, 1, (const uint32_t*)sb.part_a, (const uint32_t*)sb.part_i, (const uint32_t*)yb.head[p], no_keys,
              no_u32, no_u32, (const uint32_t*)yb.n_tiles);
    dx_launch(k_young_split, min(dx_div_up(YT_MAX_INST, 8), 32), 256, 0, s, (const Ctl*)ctl, ia, m_key, m_off, yb, q);
    // merge 2: popped list <- main[mh_old .. mh_old + ka) + young[q][0 .. kp); every output is a pop (r < k)
    InstArrays ia2 = ia;
    ia2.m_new = yb.kp;
    ia2.batch_off = yb.off[q];
    ia2.tile_off = yb.tile_off2;
    const int tiles2 = dx_div_up(max_pop, DX_SORT_TILE) + YT_MAX_INST;
    dx_launch(k_merge_partition, min(dx_div_up(tiles2, OPEN_THREADS / 32), 148 * 4), OPEN_THREADS, 0, s, (const Ctl*)ctl, ia2, sb, m_key, m_off,
              sb.part_a, sb.part_i, (const uint32_t*)yb.mh_old, (const unsigned long long*)yb.key[q], (const uint32_t*)yb.ka,
              (const uint32_t*)(yb.n_tiles + 1));
    dx_launch(k_merge, min(tiles2, 148 * 8), OPEN_THREADS, 0, s, (const Ctl*)ctl, ia2, sb, m_key, m_val, m_off, yb.key[p], yb.val[p],
              (const uint32_t*)yb.off[p], popped_next, popped_inst_next, pop_off_next, 0, (const uint32_t*)sb.part_a,
              (const uint32_t*)sb.part_i, (const uint32_t*)yb.mh_old, (const unsigned long long*)yb.key[q], (const uint32_t*)yb.val[q],
              (const uint32_t*)yb.ka, (const uint32_t*)(yb.n_tiles + 1));
    return 5;
}

// Compaction: main run <- merge(main run remainder, young run remainder) per instance, young runs emptied.  The merge itself is
// k_merge_partition + k_merge with the young run as the B run (k = 0: nothing is popped).
__global__ void __launch_bounds__(1024)
k_young_plan(Ctl* __restrict__ ctl, InstArrays ia, const uint32_t* __restrict__ m_off, uint32_t* __restrict__ m_off_next, YoungBufs yb,
             int p, uint32_t max_open) {
    if (ctl->error) return;
    if (threadIdx.x != 0) return;
    const int I = (int)ctl->n_inst;
    uint32_t so = 0, st = 0;
    for (int i = 0; i < I; ++i) {
        const bool act = ia.active[i] != 0;
        const uint32_t nm = act ? (m_off[i + 1] - m_off[i]) - yb.m_head[i] : 0u;
        const uint32_t ny = act ? (yb.off[p][i + 1] - yb.off[p][i]) - yb.head[p][i] : 0u;
        ia.m_new[i] = ny;                                  // the B run of the merge = the young remainder ...
        ia.batch_off[i] = yb.off[p][i] + yb.head[p][i];    // ... which starts here in the young buffer
        ia.k_pop[i] = 0;
        m_off_next[i] = so;
        ia.tile_off[i] = st;
        so += nm + ny;
        st += (nm + ny + DX_SORT_TILE - 1) / DX_SORT_TILE;
        if (!act) yb.m_head[i] = m_off[i + 1] - m_off[i];  // (a finished instance's run is dropped: the merge sees na = 0)
    }
    m_off_next[I] = so;
    ia.tile_off[I] = st;
    ctl->n_tiles = st;
    ctl->open_total = so;
    if (so > max_open) atomicOr(&ctl->error, DX_DEVERR_OPEN);
}

__global__ void __launch_bounds__(1024)
k_young_finish(Ctl* __restrict__ ctl, YoungBufs yb) {
    if (ctl->error) return;
    const int I = (int)ctl->n_inst;
    for (int i = threadIdx.x; i <= I; i += blockDim.x) {
        yb.off[0][i] = 0; yb.off[1][i] = 0;
        if (i < I) { yb.head[0][i] = 0; yb.head[1][i] = 0; yb.m_head[i] = 0; }
    }
    if (threadIdx.x == 0) ctl->young_total = 0;
}

// launches the compaction: open_*_cur (with the consumed prefixes m_head) + young buffer p -> open_*_next
int dx_launch_young_compact(cudaStream_t s, Ctl* ctl, const InstArrays& ia, const SortBufs& sb, const YoungBufs& yb, int p,
                            const unsigned long long* open_key_cur, const uint32_t* open_val_cur, const uint32_t* open_off_cur,
                            unsigned long long* open_key_next, uint32_t* open_val_next, uint32_t* open_off_next,
                            uint32_t* popped_scratch, uint32_t* popped_inst_scratch, const uint32_t* pop_off_any, int max_tiles,
                            uint32_t max_open) {
    k_young_plan<<<1, 32, 0, s>>>(ctl, ia, open_off_cur, open_off_next, yb, p, max_open);
    const int blocks = min(max_tiles, 148 * 8);
    k_merge_partition<<<min(dx_div_up(max_tiles, OPEN_THREADS / 32), 148 * 4), OPEN_THREADS, 0, s>>>(
        ctl, ia, sb, open_key_cur, open_off_cur, sb.part_a, sb.part_i, yb.m_head, yb.key[p], nullptr, nullptr);
    k_merge<<<blocks, OPEN_THREADS, 0, s>>>(ctl, ia, sb, open_key_cur, open_val_cur, open_off_cur, open_key_next, open_val_next,
                                            open_off_next, popped_scratch, popped_inst_scratch, pop_off_any, 1, sb.part_a, sb.part_i,
                                            yb.m_head, yb.key[p], yb.val[p], nullptr, nullptr);
    k_young_finish<<<1, 256, 0, s>>>(ctl, yb);
    return 4;
}
