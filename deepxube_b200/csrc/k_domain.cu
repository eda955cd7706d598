// State expansion / next_state / is_solved for cube3, npuzzle and grid over packed uint8 rows.
//
// Reference behaviour reproduced (children are state-major / action-minor, child = j*A + a):
//   ActsEnum.expand                      deepxube/base/domain.py:288-312
//   Cube3._next_state_np                 deepxube/domains/cube3.py:583-596  (next[new] = cur[old], table :598-671)
//   NPuzzle._move_np / swap_zero_idxs    deepxube/domains/npuzzle.py:264-321 (blocked move = no-op, cost still 1.0)
//   Grid.next_state                      deepxube/domains/grid.py:74-86
//   is_solved                            cube3.py:514-517, npuzzle.py:154-158 (goal byte d*d = wildcard), grid.py:68-69
//   goal test on POPPED nodes + record_goal (ub = min g, first seen wins)
//                                        deepxube/base/pathfinding.py:353-359, pathfinding/graph_search.py:35-41
//
// Layout: a row is SP bytes = S state bytes, zero padding, and the owning instance id in the last 4
// bytes (so that hashing / comparing whole rows keys CLOSED by (instance, state)).  One block stages
// PB parent rows in shared memory and its threads emit the children one 32-bit word at a time, so
// global stores are fully coalesced (consecutive threads -> consecutive words of consecutive children).
#include "dx_internal.cuh"

#define EXP_THREADS 256
#define EXP_PB 16          // parents per block pass
#define MAX_SP 240         // npuzzle.15: 225 + 4 -> 240

__device__ __forceinline__ uint32_t child_byte(int kind, int dim, int S, const uint8_t* __restrict__ row,
                                               const uint8_t* __restrict__ table, int a, int i, int z) {
    if (i >= S) return 0u;
    if (kind == DX_DOMAIN_CUBE3) {
        return row[table[a * 54 + i]];
    } else if (kind == DX_DOMAIN_NPUZZLE) {
        int sw = table[z * 4 + a];
        if (i == z) return row[sw];          // sw == z (blocked): row[z] == 0, unchanged
        if (i == sw) return 0u;
        return row[i];
    } else if (kind == DX_DOMAIN_LIGHTSOUT) {
        // toggled once even when the border lists the pressed cell several times (lightsout.py:186)
        const uint8_t* m = table + a * 5;
        const bool hit = i == m[0] || i == m[1] || i == m[2] || i == m[3] || i == m[4];
        return (uint32_t)row[i] ^ (hit ? 1u : 0u);
    } else {
        int v = row[i];
        if (i == 0) {
            if (a == 0) v = max(v - 1, 0);
            else if (a == 1) v = min(v + 1, dim - 1);
        } else {
            if (a == 2) v = max(v - 1, 0);
            else if (a == 3) v = min(v + 1, dim - 1);
        }
        return (uint32_t)v;
    }
}

__device__ __forceinline__ int find_blank(const uint8_t* row, int S) {
    int z = 0;
    for (int i = 0; i < S; ++i)
        if (row[i] == 0) { z = i; break; }
    return z;
}

// ---------------------------------------------------------------------------------------------
// Search-time expansion of the popped list, one instantiation per domain.
//   * a block pass stages EXP_SP parents (16-byte loads); eight threads share a parent's goal test / blank search
//     (strided over the state bytes, combined with three shuffles), thread 0 of the group records it;
//   * the children are written 16 bytes per thread (consecutive threads -> consecutive chunks of consecutive children:
//     512 contiguous bytes per warp).  cube3: a shared-memory table holds, per (action, output word), the four source byte
//     indices packed in one word (indices >= S point at the row's zero padding), so a chunk costs 4 table loads + 16 byte
//     loads; npuzzle: a child is the parent with two bytes exchanged; grid: one clamped coordinate.
#define EXP_SP 32          // parents per block pass (8 threads each)

template <int KIND>
__global__ void __launch_bounds__(EXP_THREADS)
k_expand_t(DomainDev d, const Ctl* __restrict__ ctl, InstArrays ia, const uint8_t* __restrict__ node_state,
           const float* __restrict__ node_g, const uint32_t* __restrict__ popped, const uint32_t* __restrict__ popped_inst,
           const uint32_t* __restrict__ pop_off, uint8_t* __restrict__ child_state, float* __restrict__ child_g,
           uint8_t* __restrict__ pop_solved) {
    __shared__ __align__(16) uint8_t s_rows[EXP_SP * MAX_SP];
    __shared__ uint32_t s_src4[12 * 16];                 // cube3: source bytes of output word w under action a
    __shared__ uint8_t s_lut[225 * 5];                   // npuzzle: swap partner of blank z under action a; lightsout: toggled cells
    __shared__ int s_z[EXP_SP];
    __shared__ uint32_t s_inst[EXP_SP];
    __shared__ float s_g[EXP_SP];       // NaN marks a parent of an inactive (finished) instance

    dx_pdl_prologue();
    if (ctl->error) return;
    const int n_pop = ctl->n_pop;
    const int SP = d.SP, S = d.S, A = d.A, chunks = SP / 16;
    if (KIND == DX_DOMAIN_CUBE3) {
        for (int t = threadIdx.x; t < 12 * 16; t += EXP_THREADS) {
            const int a = t >> 4, w = t & 15;
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = 4 * w + j;
                v |= (uint32_t)(i < 54 ? d.table[a * 54 + i] : 54) << (8 * j);     // byte 54 of a row is zero padding
            }
            s_src4[t] = v;
        }
    } else if (KIND == DX_DOMAIN_NPUZZLE) {
        for (int i = threadIdx.x; i < d.dim * d.dim * 4; i += EXP_THREADS) s_lut[i] = d.table[i];
    } else if (KIND == DX_DOMAIN_LIGHTSOUT) {
        for (int i = threadIdx.x; i < d.dim * d.dim * 5; i += EXP_THREADS) s_lut[i] = d.table[i];
    }

    for (int base = blockIdx.x * EXP_SP; base < n_pop; base += gridDim.x * EXP_SP) {
        const int np = min(EXP_SP, n_pop - base);
        __syncthreads();
        for (int t = threadIdx.x; t < np * chunks; t += EXP_THREADS) {
            const int p = t / chunks, c = t - p * chunks;
            const uint32_t id = popped[base + p];
            reinterpret_cast<uint4*>(s_rows + p * SP)[c] = reinterpret_cast<const uint4*>(node_state + (size_t)id * SP)[c];
        }
        __syncthreads();
        {   // parent phase: 8 threads per parent (EXP_THREADS == 8 * EXP_SP)
            const int p = threadIdx.x >> 3, sub = threadIdx.x & 7;
            const bool live = p < np;
            const int j = base + (live ? p : 0);
            const uint8_t* row = s_rows + (live ? p : 0) * SP;
            const uint32_t inst = popped_inst[j];
            const bool act = ia.active[inst] != 0;
            const uint8_t* goal = ia.goals + (size_t)inst * SP;
            const int wild = d.dim * d.dim;
            bool ok = true;
            int z = 0x7fffffff;
            for (int i = sub; i < S; i += 8) {
                const uint8_t v = row[i], g = goal[i];
                bool e = v == g;
                if (KIND == DX_DOMAIN_NPUZZLE) {
                    e = e || (g == wild);
                    if (v == 0) z = min(z, i);
                }
                ok = ok && e;
            }
#pragma unroll
            for (int sh = 1; sh < 8; sh <<= 1) {
                const int ok_other = __shfl_xor_sync(0xffffffffu, (int)ok, sh);       // (never inside a short-circuit: every
                const int z_other = __shfl_xor_sync(0xffffffffu, z, sh);              //  lane must execute the shuffle)
                ok = ok && (ok_other != 0);
                z = min(z, z_other);
            }
            if (live && sub == 0) {
                const uint32_t id = popped[j];
                const float g = node_g[id];
                s_inst[p] = inst;
                s_g[p] = act ? g : __int_as_float(0x7fc00000);
                s_z[p] = (KIND == DX_DOMAIN_NPUZZLE && z != 0x7fffffff) ? z : 0;
                const bool solved = act && ok;
                if (solved) {
                    // lexicographic min over (g, pop order): strict '>' in record_goal == first seen wins
                    const unsigned long long key = ((unsigned long long)__float_as_uint(g) << 32) |
                                                   (unsigned long long)(ia.pop_seq_base[inst] + (uint32_t)j - pop_off[inst]);
                    atomicMin(&ia.ub_key[inst], key);
                }
                pop_solved[j] = solved ? 1 : 0;
            }
        }
        __syncthreads();
        // children: one 16-byte chunk per thread step
        const int total = np * A * chunks;
        for (int t = threadIdx.x; t < total; t += EXP_THREADS) {
            const int cl = t / chunks, c = t - cl * chunks;
            const int p = cl / A, a = cl - p * A;
            const uint8_t* row = s_rows + p * SP;
            uint32_t v[4];
            if (KIND == DX_DOMAIN_CUBE3) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t s4 = s_src4[a * 16 + 4 * c + q];
                    v[q] = (uint32_t)row[s4 & 0xFF] | ((uint32_t)row[(s4 >> 8) & 0xFF] << 8) |
                           ((uint32_t)row[(s4 >> 16) & 0xFF] << 16) | ((uint32_t)row[s4 >> 24] << 24);
                }
            } else {
                const uint4 pv = reinterpret_cast<const uint4*>(row)[c];
                v[0] = pv.x; v[1] = pv.y; v[2] = pv.z; v[3] = pv.w;
                if (KIND == DX_DOMAIN_NPUZZLE) {
                    const int z = s_z[p], sw = s_lut[z * 4 + a];
                    if (sw != z) {                        // a blocked move leaves the state unchanged
                        const uint32_t tile = row[sw];
                        const int lo = 16 * c;
                        if (z >= lo && z < lo + 16) { const int k = z - lo; v[k >> 2] = (v[k >> 2] & ~(0xFFu << (8 * (k & 3)))) | (tile << (8 * (k & 3))); }
                        if (sw >= lo && sw < lo + 16) { const int k = sw - lo; v[k >> 2] &= ~(0xFFu << (8 * (k & 3))); }
                    }
                } else if (KIND == DX_DOMAIN_LIGHTSOUT) {  // the pressed cell and its on-board neighbours flip, each once
                    const int lo = 16 * c;
#pragma unroll
                    for (int q = 0; q < 5; ++q) {
                        const int m = s_lut[a * 5 + q];
                        if ((q == 0 || m != a) && m >= lo && m < lo + 16) { const int k = m - lo; v[k >> 2] ^= 1u << (8 * (k & 3)); }
                    }
                } else if (c == 0) {                      // grid: byte 0 = x (actions 0 / 1), byte 1 = y (actions 2 / 3)
                    int x = (int)(v[0] & 0xFF), y = (int)((v[0] >> 8) & 0xFF);
                    if (a == 0) x = max(x - 1, 0);
                    else if (a == 1) x = min(x + 1, d.dim - 1);
                    else if (a == 2) y = max(y - 1, 0);
                    else y = min(y + 1, d.dim - 1);
                    v[0] = (v[0] & 0xFFFF0000u) | (uint32_t)x | ((uint32_t)y << 8);
                }
            }
            if (c == chunks - 1) v[3] = s_inst[p];        // the row's last word is the owning instance
            reinterpret_cast<uint4*>(child_state + ((size_t)(base * A + cl)) * SP)[c] = make_uint4(v[0], v[1], v[2], v[3]);
        }
        for (int t = threadIdx.x; t < np * A; t += EXP_THREADS) {
            const float g = s_g[t / A];
            child_g[(size_t)base * A + t] = g + 1.0f;     // transition cost 1.0 in all three domains; NaN stays NaN
        }
    }
}

int dx_launch_expand(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const InstArrays& ia, const uint8_t* node_state,
                     const float* node_g, const uint32_t* popped, const uint32_t* popped_inst, const uint32_t* pop_off,
                     uint8_t* child_state, float* child_g, uint8_t* pop_solved, int max_pop) {
    static_assert(EXP_THREADS == 8 * EXP_SP, "eight threads per staged parent");
    const int blocks = min(dx_div_up(max_pop, EXP_SP), 148 * 8);
#define DX_EXPAND_ARGS d, ctl, ia, node_state, node_g, popped, popped_inst, pop_off, child_state, child_g, pop_solved
    if (d.kind == DX_DOMAIN_CUBE3) dx_launch(k_expand_t<DX_DOMAIN_CUBE3>, blocks, EXP_THREADS, 0, s, DX_EXPAND_ARGS);
    else if (d.kind == DX_DOMAIN_NPUZZLE) dx_launch(k_expand_t<DX_DOMAIN_NPUZZLE>, blocks, EXP_THREADS, 0, s, DX_EXPAND_ARGS);
    else if (d.kind == DX_DOMAIN_LIGHTSOUT) dx_launch(k_expand_t<DX_DOMAIN_LIGHTSOUT>, blocks, EXP_THREADS, 0, s, DX_EXPAND_ARGS);
    else dx_launch(k_expand_t<DX_DOMAIN_GRID>, blocks, EXP_THREADS, 0, s, DX_EXPAND_ARGS);
#undef DX_EXPAND_ARGS
    return 1;
}

// goal_node = the popped node whose (g, sequence) key won the atomicMin.
__global__ void k_goal_resolve(const Ctl* __restrict__ ctl, InstArrays ia, const float* __restrict__ node_g,
                               const uint32_t* __restrict__ popped, const uint32_t* __restrict__ popped_inst,
                               const uint32_t* __restrict__ pop_off, const uint8_t* __restrict__ pop_solved) {
    if (ctl->error) return;
    const int n = ctl->n_pop;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        if (!pop_solved[j]) continue;
        uint32_t inst = popped_inst[j], id = popped[j];
        unsigned long long key = ((unsigned long long)__float_as_uint(node_g[id]) << 32) |
                                 (unsigned long long)(ia.pop_seq_base[inst] + (uint32_t)j - pop_off[inst]);
        if (ia.ub_key[inst] == key) ia.goal_node[inst] = id;
    }
}

int dx_launch_goal_resolve(cudaStream_t s, const Ctl* ctl, const InstArrays& ia, const float* node_g, const uint32_t* popped,
                           const uint32_t* popped_inst, const uint32_t* pop_off, const uint8_t* pop_solved, int max_pop) {
    int blocks = min(dx_div_up(max_pop, 256), 148 * 4);
    k_goal_resolve<<<blocks, 256, 0, s>>>(ctl, ia, node_g, popped, popped_inst, pop_off, pop_solved);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// Stand-alone domain API (expand / next_state / is_solved on caller-provided rows).
// actions == nullptr: all A children per row (out rows = n*A); else one child per row.
__global__ void __launch_bounds__(EXP_THREADS)
k_standalone_expand(DomainDev d, const uint8_t* __restrict__ rows, long long n, const int32_t* __restrict__ actions,
                    uint8_t* __restrict__ out_rows) {
    __shared__ __align__(16) uint8_t s_rows[EXP_PB * MAX_SP];
    __shared__ uint8_t s_table[225 * 5];          // cube3 12*54, npuzzle dim*dim*4, lightsout dim*dim*5
    __shared__ int s_z[EXP_PB];
    __shared__ int s_act[EXP_PB];
    const int SP = d.SP, S = d.S, words = SP / 4;
    const int A = actions ? 1 : d.A;
    const int tsz = d.kind == DX_DOMAIN_CUBE3 ? 12 * 54 : (d.kind == DX_DOMAIN_NPUZZLE ? d.dim * d.dim * 4
                    : (d.kind == DX_DOMAIN_LIGHTSOUT ? d.dim * d.dim * 5 : 0));
    for (int i = threadIdx.x; i < tsz; i += EXP_THREADS) s_table[i] = d.table[i];
    for (long long base = (long long)blockIdx.x * EXP_PB; base < n; base += (long long)gridDim.x * EXP_PB) {
        const int np = (int)min((long long)EXP_PB, n - base);
        __syncthreads();
        const int chunks = SP / 16;
        for (int t = threadIdx.x; t < np * chunks; t += EXP_THREADS) {
            int p = t / chunks, c = t % chunks;
            reinterpret_cast<uint4*>(s_rows + p * SP)[c] = reinterpret_cast<const uint4*>(rows + (size_t)(base + p) * SP)[c];
        }
        __syncthreads();
        if (threadIdx.x < np) {
            const uint8_t* row = s_rows + threadIdx.x * SP;
            s_z[threadIdx.x] = (d.kind == DX_DOMAIN_NPUZZLE) ? find_blank(row, S) : 0;
            s_act[threadIdx.x] = actions ? actions[base + threadIdx.x] : 0;
        }
        __syncthreads();
        const int total_words = np * A * words;
        for (int t = threadIdx.x; t < total_words; t += EXP_THREADS) {
            const int cl = t / words, w = t % words;
            const int p = cl / A, a = actions ? s_act[p] : (cl % A);
            const uint8_t* row = s_rows + p * SP;
            uint32_t v;
            if (w == words - 1) {
                v = reinterpret_cast<const uint32_t*>(row)[w];
            } else {
                const int i0 = w * 4;
                v = child_byte(d.kind, d.dim, S, row, s_table, a, i0, s_z[p]) |
                    (child_byte(d.kind, d.dim, S, row, s_table, a, i0 + 1, s_z[p]) << 8) |
                    (child_byte(d.kind, d.dim, S, row, s_table, a, i0 + 2, s_z[p]) << 16) |
                    (child_byte(d.kind, d.dim, S, row, s_table, a, i0 + 3, s_z[p]) << 24);
            }
            reinterpret_cast<uint32_t*>(out_rows + ((size_t)(base * A + cl)) * SP)[w] = v;
        }
    }
}

int dx_launch_standalone_expand(cudaStream_t s, const DomainDev& d, const uint8_t* rows, long long n, const int32_t* actions,
                                uint8_t* out_rows) {
    if (n <= 0) return 0;
    int blocks = (int)min((long long)dx_div_up(n, EXP_PB), (long long)148 * 8);
    k_standalone_expand<<<blocks, EXP_THREADS, 0, s>>>(d, rows, n, actions, out_rows);
    return 1;
}

__global__ void k_standalone_solved(DomainDev d, const uint8_t* __restrict__ rows, const uint8_t* __restrict__ goal_rows,
                                    long long n, uint8_t* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = row_solved(d.kind, d.dim, d.S, rows + (size_t)i * d.SP, goal_rows + (size_t)i * d.SP) ? 1 : 0;
}

int dx_launch_standalone_solved(cudaStream_t s, const DomainDev& d, const uint8_t* rows, const uint8_t* goal_rows, long long n,
                                uint8_t* out) {
    if (n <= 0) return 0;
    int blocks = (int)min((long long)dx_div_up(n, 256), (long long)148 * 8);
    k_standalone_solved<<<blocks, 256, 0, s>>>(d, rows, goal_rows, n, out);
    return 1;
}

// ---------------------------------------------------------------------------------------------
// Q*: goal test of the popped nodes only (no expansion at this point; base/pathfinding.py:486-492).
__global__ void k_check_solved(DomainDev d, const Ctl* __restrict__ ctl, InstArrays ia, const uint8_t* __restrict__ node_state,
                               const float* __restrict__ node_g, const uint32_t* __restrict__ popped,
                               const uint32_t* __restrict__ popped_inst, const uint32_t* __restrict__ pop_off,
                               uint8_t* __restrict__ pop_solved) {
    if (ctl->error) return;
    const int n = ctl->n_pop;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t inst = popped_inst[j], id = popped[j];
        bool solved = false;
        if (ia.active[inst]) {
            solved = row_solved(d.kind, d.dim, d.S, node_state + (size_t)id * d.SP, ia.goals + (size_t)inst * d.SP);
            if (solved) {
                unsigned long long key = ((unsigned long long)__float_as_uint(node_g[id]) << 32) |
                                         (unsigned long long)(ia.pop_seq_base[inst] + (uint32_t)j - pop_off[inst]);
                atomicMin(&ia.ub_key[inst], key);
            }
        }
        pop_solved[j] = solved ? 1 : 0;
    }
}

int dx_launch_check_solved(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const InstArrays& ia, const uint8_t* node_state,
                           const float* node_g, const uint32_t* popped, const uint32_t* popped_inst, const uint32_t* pop_off,
                           uint8_t* pop_solved, int max_pop) {
    int blocks = min(dx_div_up(max_pop, 128), 148 * 8);
    k_check_solved<<<blocks, 128, 0, s>>>(d, ctl, ia, node_state, node_g, popped, popped_inst, pop_off, pop_solved);
    return 1;
}

// Q*: one child per popped edge (base/pathfinding.py:527-566).  edges[r] = parent * A + action;
// popped_out[r] = id of the generated node (= node_base + r, creation order): the next popped-node list.
__global__ void __launch_bounds__(128)
k_expand_edges(DomainDev d, const Ctl* __restrict__ ctl, const uint32_t* __restrict__ edges, uint32_t* __restrict__ popped_out,
               uint8_t* __restrict__ node_state, float* __restrict__ node_g, uint32_t* __restrict__ node_parent,
               uint8_t* __restrict__ node_action) {
    dx_pdl_prologue();
    if (ctl->error) return;
    const uint32_t n = ctl->n_kept, base = ctl->node_base;
    const int SP = d.SP, S = d.S, words = SP / 4;
    // 4 lanes cooperate on one edge: coalesced 16-byte pieces of the parent row, word-wise stores
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t sub = gt & 3;
    for (uint32_t r = gt >> 2; r < n; r += (gridDim.x * blockDim.x) >> 2) {
        const uint32_t code = edges[r];
        const uint32_t parent = code / d.A, a = code % d.A;
        const uint8_t* prow = node_state + (size_t)parent * SP;
        const int z = (d.kind == DX_DOMAIN_NPUZZLE) ? find_blank(prow, S) : 0;
        const uint32_t id = base + r;
        uint32_t* out = reinterpret_cast<uint32_t*>(node_state + (size_t)id * SP);
        for (int w = sub; w < words; w += 4) {
            uint32_t v;
            if (w == words - 1) {
                v = reinterpret_cast<const uint32_t*>(prow)[w];
            } else {
                const int i0 = w * 4;
                v = child_byte(d.kind, d.dim, S, prow, d.table, a, i0, z) |
                    (child_byte(d.kind, d.dim, S, prow, d.table, a, i0 + 1, z) << 8) |
                    (child_byte(d.kind, d.dim, S, prow, d.table, a, i0 + 2, z) << 16) |
                    (child_byte(d.kind, d.dim, S, prow, d.table, a, i0 + 3, z) << 24);
            }
            out[w] = v;
        }
        if (sub == 0) {
            node_g[id] = node_g[parent] + 1.0f;
            node_parent[id] = parent;
            node_action[id] = (uint8_t)a;
            popped_out[r] = id;
        }
    }
}

int dx_launch_expand_edges(cudaStream_t s, const DomainDev& d, const Ctl* ctl, const uint32_t* edges, uint32_t* popped_out,
                           uint8_t* node_state, float* node_g, uint32_t* node_parent, uint8_t* node_action, int max_pop) {
    int blocks = min(dx_div_up((long long)max_pop * 4, 128), 148 * 8);
    dx_launch(k_expand_edges, blocks, 128, 0, s, d, ctl, edges, popped_out, node_state, node_g, node_parent, node_action);
    return 1;
}
