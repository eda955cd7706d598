// resnet_fc cost-to-go network for SMALL evaluations (a few thousand rows, res_dim <= 256): the WHOLE network in ONE launch.
//
// Reference network (eval mode), deepxube/nnets/resnet_fc.py:19-52 + deepxube/pytorch/pytorch_models.py:15-24, 149-209:
//   x0 = W0 . onehot(s) + b0 ; per block: y = relu(BN1(W1 x + b1)); x = relu(BN2(W2 y + b2) + x) ; out = Wo x + bo
//   h = max(out[0], 0) (base/pathfind_fns.py:218-221) ; q = max(out, 0), h = min_a q (fns_core.py:78-84, pathfinding.py:722-744)
//
// Why a second network kernel: the solve-time half of the metric (deepxube/_solve.py:150-161, one instance, graph.100B:
// ~1200 children per iteration, resnet_fc.200H_2B) is latency-bound -- six per-layer tcgen05 launches of 1200 rows cost
// ~8 us each although each has < 1 us of tensor work.  Here a CTA carries a tile of 16 / 32 rows through every layer:
//   * activations never leave the SM: bf16 (or fp16 hi/lo) A operands ping-pong between two shared-memory buffers, the
//     residual (skip) stream stays in fp32 REGISTERS (it has the accumulator's fragment layout), so the bf16 mode rounds
//     only the GEMM operands, never the skip connection;
//   * the one-hot encoding is written straight into the shared-memory A operand of the first layer from the state bytes
//     (no one-hot matrix in HBM);
//   * weights are pre-packed on the host in mma B-fragment order (16 consumer warps, a warp owns 16 output columns) in groups
//     of 4 (split: 2) k-steps; a 17th warp's lane 0 streams the groups L2 -> shared memory with ONE TMA bulk copy each through
//     a 3-stage full / empty mbarrier ring that runs across layer and tile boundaries; a consumer warp moves a group's
//     fragments into registers, hands the stage back, and issues the group's ldmatrix / MMAs with compile-time offsets;
//   * one consumer barrier per layer (the A operand of the next layer is complete);
//   * a scalar output layer is folded into the last layer's epilogue (fp32 dot with the accumulator fragments, fixed
//     reduction order); a multi-output layer runs on the CUDA cores from the fp32 activations.
// Tensor work is issued as warp-level mma.sync m16n8k16 (bf16, or fp16 for the split mode): with 16-32 rows per CTA a
// tcgen05 tile (M = 128) would be 75-88 % padding.  What bounds a k-step here is shared-memory bandwidth (measured by removing
// one ingredient at a time, profiles/r2_small_iteration.md): 13 warps x (512 B of B fragments + 512 B of A fragments) = 13 KB
// per step against 128 B / clk, ~105 of the ~130 cycles a step takes; the weight stream (L2 -> smem) and the MMAs hide under it.
// fp32-accurate mode (DX_HEUR_NET_FP32): operands split into fp16 (hi, lo) pairs, products A_hi.W_lo + A_lo.W_hi + A_hi.W_hi
// accumulated in fp32 (weights scaled by a power of two per layer so that the lo halves are normal numbers; the one-hot A of
// the first layer is exact, so it needs two products).
#include "dx_internal.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#define SM_CWARPS 16          // consumer (MMA) warps: a warp owns 16 output columns
#define SM_THREADS (32 * (SM_CWARPS + 1))   // + one producer warp (issues the bulk copies of the weight stream)
#define SM_NT 2               // n8 tiles per consumer warp
#define SM_MAX_LAYERS 32
#define SM_STAGES 3           // weight stages in flight

struct SmallNetDev {
    const uint4* wstream;     // [group][step in group][16 warps][plane][lane] 16-byte chunks (planes: bf16 1, split 2 = hi, lo);
                              // a group of G k-steps (bf16 4, split 2) is ONE contiguous bulk copy with compile-time strides
    const float* b0;          // [Hp]
    const float* bres;        // [L, Hp]
    const float* wout;        // [out_dim, Hp]
    const float* bout;        // [out_dim]
    int in_count, depth;
    int ks0, ksh;             // k16 steps of the first / a hidden layer, each rounded up to whole groups (zero weights)
    int L;                    // hidden layers (2 * num_blocks)
    int n_groups;             // groups per evaluation = (ks0 + L * ksh) / G
    int H, Hp, KP;            // KP = H rounded up to 16 (k extent of the hidden layers), Hp = row stride of the fp32 parameters
    int nwa;                  // active warps = ceil(H / 16)
    int out_dim;
    int wout_rows;            // out_dim > 1 and the output weights are staged in shared memory (out_dim * KP * 4 <= 16 KB), else 0
    float inv_scale[1 + SM_MAX_LAYERS];   // split mode: 1 / (power-of-two weight scale) of the first and the hidden layers
};

struct SmallNetHost {         // NetDev::small
    SmallNetDev dev;
    void* wstream;            // device allocation behind dev.wstream
    size_t stream_bytes;
    int split;
};

// -DSM_PROFILE: clock64 timeline of CTA 0 (thread 0), accumulated over launches: [0] launches, [1] job read + rows staged,
// [2] one-hot operand built, [3 + l] end of layer l, [14] outputs written (read and cleared by dx_debug_tc_profile)
__device__ unsigned long long g_small_prof[16];
#ifdef SM_PROFILE
#define SM_MARK(slot) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long _t = clock64(); \
        atomicAdd(&g_small_prof[(slot)], (unsigned long long)(_t - t_prev)); t_prev = _t; } } while (0)
#else
#define SM_MARK(slot) do { } while (0)
#endif

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sm_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Spinning wait (test_wait, no suspend): the waits of this kernel last tens to hundreds of cycles, and a suspended try_wait
// measured ~1 us of wake-up latency per blocked wait (the whole pipeline then ran at one group per microsecond).  Bounded: a
// barrier that does not complete within ~2^28 polls is a protocol bug -- trap instead of hanging the GPU.
__device__ __forceinline__ void sm_mbar_wait(uint32_t bar, uint32_t parity) {
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void sm_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sm_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA bulk copy global -> shared (1-D, bytes % 16 == 0), completion counted on an mbarrier
__device__ __forceinline__ void sm_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void sm_consumer_sync() {            // the SM_CWARPS consumer warps only (the producer runs ahead)
    asm volatile("bar.sync 1, %0;" ::"n"(32 * SM_CWARPS) : "memory");
}

__device__ __forceinline__ void sm_ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}

template <bool F16>
__device__ __forceinline__ void sm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if (F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// MT = m16 row tiles per CTA (rows per CTA = 16 * MT); SPLIT = fp32-accurate fp16 (hi, lo) mode.
// Weight pipeline: SM_STAGES stages of G k-steps each; the producer warp's lane 0 refills a stage with ONE bulk copy as soon as
// every active consumer warp has released it (full / empty mbarriers, the classic TMA ring), across layer and tile boundaries.
template <int MT, bool SPLIT>
__global__ void __launch_bounds__(SM_THREADS)
k_mlp_small(SmallNetDev sn, DomainDev d, const Ctl* __restrict__ ctl, const uint8_t* __restrict__ rows,
            const uint8_t* __restrict__ goals, float* __restrict__ out_h, float* __restrict__ out_q, int absolute) {
    constexpr int R = 16 * MT;
    constexpr int NP = SPLIT ? 2 : 1;                 // operand planes (hi, lo) = 16-byte chunks per lane and step
    constexpr int G = SPLIT ? 2 : 4;                  // k-steps per stage
    constexpr int D = SM_STAGES;
    extern __shared__ __align__(128) uint8_t sm_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int KP = sn.KP, astride = sn.ksh * 16 + 8;  // activation row stride in 16-bit elements (conflict-free ldmatrix)
    const int K0 = sn.ks0 * 16, a0stride = K0 + 8;
    const int od = sn.out_dim;
    const int SP = d.SP, gw = d.in_count > d.S ? 2 : 1;      // staged bytes per row: the row (+ its goal row when the goal is an input)
    constexpr int step_chunks = SM_CWARPS * NP * 32;  // 16-byte chunks per k-step (compile-time: immediate offsets in the k loop)
    const uint32_t stage_bytes = (uint32_t)G * step_chunks * 16;
    // shared-memory carve-up
    uint4* s_w = reinterpret_cast<uint4*>(sm_raw);                                    // [D stages][G][nwa][NP][32]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_w + D * G * step_chunks);         // full[D], empty[D]
    uint4* s_rows = reinterpret_cast<uint4*>(s_bar + 2 * D + 2);                      // [R][gw * SP] staged state (+ goal) rows
    uint16_t* s_a0 = reinterpret_cast<uint16_t*>(s_rows + R * gw * SP / 16);          // [R][a0stride] one-hot input operand
    uint16_t* s_act = s_a0 + R * a0stride;                                            // [2 buffers][NP planes][R][astride]
    float* s_fin = reinterpret_cast<float*>(s_act + 2 * NP * R * astride);            // [R][KP + 1] (out_dim > 1 only)
    float* s_part = s_fin + (od > 1 ? R * (KP + 1) : 0);                              // [SM_CWARPS][R] partial dots / [R][od] q-values
    float* s_wout = s_part + max(SM_CWARPS * R, od > 1 ? R * od : 0);                 // [wout_rows][KP] (0 rows: read from global)
    const int act_plane = R * astride;                // elements per plane
    const int act_buf = NP * act_plane;               // elements per buffer
    const uint32_t bar_full = sm_u32(s_bar), bar_empty = sm_u32(s_bar + D);
#ifdef SM_PROFILE
    long long t_prev = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&g_small_prof[0], 1ull);
#endif
    dx_pdl_trigger();
    if (tid == 0) {
        for (int i = 0; i < D; ++i) { sm_mbar_init(bar_full + 8 * i, 1); sm_mbar_init(bar_empty + 8 * i, sn.nwa); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    dx_pdl_wait();                                    // (everything above is independent of the kernels before this one)
    if (ctl->error) return;
    const uint32_t n = ctl->nn_n, start = ctl->nn_row_start;
    if (blockIdx.x == 0 && tid == 0) atomicAdd((unsigned long long*)&ctl->heur_evals, (unsigned long long)n);
    for (int i = tid; i < sn.wout_rows * KP; i += SM_THREADS) {
        const int o = i / KP, k = i - o * KP;
        s_wout[i] = k < sn.H ? sn.wout[(size_t)o * sn.Hp + k] : 0.f;
    }
    // columns KP .. ksh * 16 of the activation operands meet zero weights: they must be finite, and are never written again
    for (int i = tid; i < 2 * NP * R * (astride - KP); i += SM_THREADS) {
        const int r = i / (astride - KP), c = i - r * (astride - KP);
        s_act[r * astride + KP + c] = 0;
    }
    __syncthreads();                                  // barriers initialised

    const bool producer = warp == SM_CWARPS;
    const bool active = warp < sn.nwa;
    // producer warp, lane 0: streams the weight groups of a tile; `issue(k)` sends the next k groups of the current tile
    uint32_t p_count = 0;                             // groups issued so far (all tiles)
    int p_next = 0;                                   // next group of the current tile
    auto issue = [&](int k) {
        for (; k > 0 && p_next < sn.n_groups; --k, ++p_next, ++p_count) {
            const uint32_t slot = p_count % D, use = p_count / D;
            sm_mbar_wait(bar_empty + 8 * slot, (use & 1) ^ 1);             // (first use of a slot: passes at once)
            sm_mbar_expect_tx(bar_full + 8 * slot, stage_bytes);
            sm_bulk_g2s(sm_u32(s_w + (size_t)slot * G * step_chunks), reinterpret_cast<const uint8_t*>(sn.wstream) + (size_t)p_next * stage_bytes,
                        stage_bytes, bar_full + 8 * slot);
        }
    };
    const uint16_t one16 = SPLIT ? (uint16_t)0x3C00 : (uint16_t)0x3F80;   // 1.0 in fp16 / bf16
    const int dep = sn.depth, in_dim = sn.in_count * dep;
    const int rchunks = SP / 16;
    uint32_t gcount = 0;                              // groups consumed so far (all tiles)

    for (uint32_t base = blockIdx.x * R; base < n; base += gridDim.x * R) {
        const int nr = (int)min((uint32_t)R, n - base);
        __syncthreads();                              // previous tile done with every buffer (and with every weight stage)
        if (producer) {                               // the first stages fly while the rows are staged
            if (lane == 0) { p_next = 0; issue(D); }
            __syncwarp();
        }
        // stage the tile's rows with 16-byte loads (one round trip), then -- domains whose goal is a network input -- the goal
        // rows of their instances (the instance id is the row's last word)
        for (int idx = tid; idx < R * rchunks; idx += SM_THREADS) {
            const int r = idx / rchunks, c = idx - r * rchunks;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (r < nr) v = reinterpret_cast<const uint4*>(rows + (size_t)(start + base + r) * SP)[c];
            s_rows[r * gw * rchunks + c] = v;
        }
        if (gw == 2) {
            __syncthreads();
            for (int idx = tid; idx < R * rchunks; idx += SM_THREADS) {
                const int r = idx / rchunks, c = idx - r * rchunks;
                const uint32_t inst = reinterpret_cast<const uint32_t*>(s_rows + r * gw * rchunks)[SP / 4 - 1];
                uint4 v = make_uint4(0, 0, 0, 0);
                if (r < nr) v = reinterpret_cast<const uint4*>(goals + (size_t)inst * SP)[c];
                s_rows[r * gw * rchunks + rchunks + c] = v;
            }
        }
        __syncthreads();
        SM_MARK(1);
        // the one-hot A operand of the first layer (OneHot.forward, pytorch_models.py:15-24), built ONCE per tile: a warp per
        // row, lanes over the inputs: input i -> its `depth` columns; depth 1: the value itself (integers <= 255 are exact)
        for (int r = warp; r < R; r += SM_CWARPS + 1) {
            const uint8_t* rb = reinterpret_cast<const uint8_t*>(s_rows + r * gw * rchunks);
            uint16_t* arow0 = s_a0 + r * a0stride;
            for (int i = lane; i < sn.in_count; i += 32) {
                const int v = r < nr ? (int)(i < d.S ? rb[i] : rb[SP + (i - d.S)]) : -1;
                uint16_t* dst = arow0 + i * dep;
                if (dep > 1) {
                    for (int q = 0; q < dep; ++q) dst[q] = (q == v) ? one16 : (uint16_t)0;
                } else {
                    const float xv = v < 0 ? 0.f : (float)v;
                    dst[0] = SPLIT ? __half_as_ushort(__float2half_rn(xv)) : __bfloat16_as_ushort(__float2bfloat16_rn(xv));
                }
            }
            for (int c = in_dim + lane; c < a0stride; c += 32) arow0[c] = 0;          // padding columns
        }
        __syncthreads();
        SM_MARK(2);
        if (producer) {                               // (the consumers synchronise among themselves from here on)
            if (lane == 0) issue(sn.n_groups);
            __syncwarp();
            continue;
        }

        float skip[MT][SM_NT][4];
        int cur = 0;
        uint32_t slot = gcount % D;
        for (int layer = 0; layer <= sn.L; ++layer) {
            const bool first = layer == 0, last = layer == sn.L;
            const int ksteps = first ? sn.ks0 : sn.ksh;
            float acc[MT][SM_NT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int j = 0; j < SM_NT; ++j)
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
            if (active) {
                // epilogue parameters: requested now, needed after the k loop
                const float* bias = first ? sn.b0 : sn.bres + (size_t)(layer - 1) * sn.Hp;
                float bv[SM_NT][2], wv[SM_NT][2];
#pragma unroll
                for (int j = 0; j < SM_NT; ++j) {
                    const int col = warp * 16 + j * 8 + 2 * t;            // (col + 1 < Hp: 16 * nwa <= Hp)
                    bv[j][0] = bias[col]; bv[j][1] = bias[col + 1];
                    wv[j][0] = (last && od == 1) ? sn.wout[col] : 0.f;
                    wv[j][1] = (last && od == 1) ? sn.wout[col + 1] : 0.f;
                }
                const uint16_t* abase = first ? s_a0 : s_act + cur * act_buf;
                const int as = first ? a0stride : astride;
                const uint16_t* arow = abase + (lane & 15) * as + (lane >> 4) * 8;
                SM_MARK(3);
                for (int gi = 0; gi < ksteps / G; ++gi) {
                    sm_mbar_wait(bar_full + 8 * slot, (gcount / D) & 1);           // the group's bulk copy has landed
                    const uint4* wstage = s_w + (size_t)slot * G * step_chunks + (warp * NP) * 32 + lane;
                    uint4 w[G][NP];                   // the group's B fragments -> registers, then the stage goes back at once
#pragma unroll
                    for (int sg = 0; sg < G; ++sg)
#pragma unroll
                        for (int p = 0; p < NP; ++p) w[sg][p] = wstage[sg * step_chunks + p * 32];
                    __syncwarp();
                    if (lane == 0) sm_mbar_arrive(bar_empty + 8 * slot);
                    ++gcount;
                    slot = (slot + 1 == D) ? 0 : slot + 1;
                    const uint16_t* ag = arow + gi * (G * 16);
#pragma unroll
                    for (int sg = 0; sg < G; ++sg) {
                        uint32_t a[NP][MT][4];        // A fragments: [plane][m tile][4]
#pragma unroll
                        for (int p = 0; p < NP; ++p)
#pragma unroll
                            for (int m = 0; m < MT; ++m)
                                if (p == 0 || !first) // (the one-hot operand is exact: no lo plane)
                                    sm_ldmatrix_x4(a[p][m], ag + p * act_plane + m * 16 * as + sg * 16);
#pragma unroll
                        for (int j = 0; j < SM_NT; ++j) {
                            const uint32_t* wh = reinterpret_cast<const uint32_t*>(&w[sg][0]);
                            const uint32_t bh0 = wh[j * 2], bh1 = wh[j * 2 + 1];
                            if (SPLIT) {
                                const uint32_t* wl = reinterpret_cast<const uint32_t*>(&w[sg][NP - 1]);
                                const uint32_t bl0 = wl[j * 2], bl1 = wl[j * 2 + 1];
#pragma unroll
                                for (int m = 0; m < MT; ++m) {  // small terms first
                                    sm_mma<true>(acc[m][j], a[0][m], bl0, bl1);
                                    if (!first) sm_mma<true>(acc[m][j], a[NP - 1][m], bh0, bh1);
                                    sm_mma<true>(acc[m][j], a[0][m], bh0, bh1);
                                }
                            } else {
#pragma unroll
                                for (int m = 0; m < MT; ++m) sm_mma<false>(acc[m][j], a[0][m], bh0, bh1);
                            }
                        }
                    }
                }
                SM_MARK(11);
                // epilogue: bias (+ skip) (+ relu), then the next A operand into the other buffer -- or, after the last layer,
                // the output dot product (out_dim 1) / the fp32 rows of the output layer (out_dim > 1)
                const float inv = SPLIT ? sn.inv_scale[layer] : 1.f;
                const bool second = !first && ((layer - 1) & 1);          // second Linear of a block: + skip, relu
                uint16_t* nbase = s_act + (cur ^ 1) * act_buf;
                bool bad = false;
                float dot[MT][2];
#pragma unroll
                for (int m = 0; m < MT; ++m) dot[m][0] = dot[m][1] = 0.f;
#pragma unroll
                for (int j = 0; j < SM_NT; ++j) {
                    const int col = warp * 16 + j * 8 + 2 * t;
#pragma unroll
                    for (int m = 0; m < MT; ++m)
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            const int r = m * 16 + g + hh * 8;
                            float v0 = acc[m][j][hh * 2] * inv + bv[j][0], v1 = acc[m][j][hh * 2 + 1] * inv + bv[j][1];
                            if (first) {
                                skip[m][j][hh * 2] = v0; skip[m][j][hh * 2 + 1] = v1;
                            } else if (second) {
                                v0 = fmaxf(v0 + skip[m][j][hh * 2], 0.f); v1 = fmaxf(v1 + skip[m][j][hh * 2 + 1], 0.f);
                                skip[m][j][hh * 2] = v0; skip[m][j][hh * 2 + 1] = v1;
                            } else {
                                v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f);
                            }
                            if (last) {
                                if (od == 1) {
                                    dot[m][hh] = fmaf(v1, wv[j][1], fmaf(v0, wv[j][0], dot[m][hh]));
                                } else if (col < KP) {
                                    s_fin[r * (KP + 1) + col] = v0;
                                    s_fin[r * (KP + 1) + col + 1] = v1;
                                }
                            } else if (col < KP) {
                                if (SPLIT) {
                                    const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
                                    const __half l0 = __float2half_rn(v0 - __half2float(h0)), l1 = __float2half_rn(v1 - __half2float(h1));
                                    bad = bad || !(fabsf(v0) <= 65504.f) || !(fabsf(v1) <= 65504.f);
                                    *reinterpret_cast<uint32_t*>(nbase + r * astride + col) =
                                        (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
                                    *reinterpret_cast<uint32_t*>(nbase + act_plane + r * astride + col) =
                                        (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
                                } else {
                                    *reinterpret_cast<uint32_t*>(nbase + r * astride + col) =
                                        (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v0)) |
                                        ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v1)) << 16);
                                }
                            }
                        }
                }
                if (last && od == 1) {                 // this warp's 16 columns of every row: quad reduction in a fixed order
#pragma unroll
                    for (int m = 0; m < MT; ++m)
#pragma unroll
                        for (int hh = 0; hh < 2; ++hh) {
                            float x = dot[m][hh];
                            x += __shfl_xor_sync(0xffffffffu, x, 1);
                            x += __shfl_xor_sync(0xffffffffu, x, 2);
                            if (t == 0) s_part[warp * R + m * 16 + g + hh * 8] = x;
                        }
                }
                if (SPLIT && __any_sync(0xffffffffu, bad) && lane == 0)
                    atomicOr(const_cast<uint32_t*>(&ctl->error), DX_DEVERR_NUMERIC);
            } else if (last) {
                // idle consumer warps keep the group count of the others (they never touch the barriers)
                gcount += sn.n_groups;
                slot = gcount % D;
            }
            cur ^= 1;
            SM_MARK(12);
            sm_consumer_sync();                       // the next layer's A operand (or s_part / s_fin) is complete
            SM_MARK(13);
        }
        // output layer: h = max(out, 0) ; q = max(out, 0), h = min_a q   (consumer threads only)
        if (od == 1) {
            if (tid < nr) {
                float x = 0.f;
                for (int w = 0; w < sn.nwa; ++w) x += s_part[w * R + tid];             // fixed order: deterministic
                const float v = fmaxf(x + sn.bout[0], 0.f);
                const size_t oi = absolute ? (size_t)start + base + tid : (size_t)base + tid;
                if (out_q) out_q[oi] = v;
                if (out_h) out_h[oi] = v;
            }
        } else {
            for (int p = tid; p < nr * od; p += 32 * SM_CWARPS) {
                const int r = p / od, o = p - r * od;
                const float* x = s_fin + r * (KP + 1);
                const float* wrow = sn.wout_rows ? s_wout + o * KP : sn.wout + (size_t)o * sn.Hp;
                float a4[4] = {0.f, 0.f, 0.f, 0.f};
                int k = 0;
                for (; k + 4 <= sn.H; k += 4) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) a4[u] = fmaf(x[k + u], wrow[k + u], a4[u]);
                }
                for (; k < sn.H; ++k) a4[0] = fmaf(x[k], wrow[k], a4[0]);
                const float v = fmaxf((a4[0] + a4[1]) + (a4[2] + a4[3]) + sn.bout[o], 0.f);
                s_part[p] = v;
                const size_t oi = absolute ? (size_t)start + base + r : (size_t)base + r;
                if (out_q) out_q[oi * od + o] = v;
            }
            sm_consumer_sync();
            if (tid < nr && out_h) {
                float hmin = __int_as_float(0x7f800000);
                for (int o = 0; o < od; ++o) hmin = fminf(hmin, s_part[tid * od + o]);
                out_h[absolute ? (size_t)start + base + tid : (size_t)base + tid] = hmin;
            }
        }
        SM_MARK(14);
    }
}

// ---------------------------------------------------------------------------------------------
static size_t small_smem_bytes(const SmallNetDev& sn, const DomainDev& d, int MT, bool split) {
    const int R = 16 * MT, NP = split ? 2 : 1, G = split ? 2 : 4;
    size_t b = (size_t)SM_STAGES * G * SM_CWARPS * NP * 32 * 16 + (2 * SM_STAGES + 2) * 8;
    b += (size_t)R * (d.in_count > d.S ? 2 : 1) * d.SP;
    b += (size_t)R * (sn.ks0 * 16 + 8) * 2;
    b += (size_t)2 * NP * R * (sn.ksh * 16 + 8) * 2;
    if (sn.out_dim > 1) b += (size_t)R * (sn.KP + 1) * 4;
    b += (size_t)std::max(SM_CWARPS * R, sn.out_dim > 1 ? R * sn.out_dim : 0) * 4;
    b += (size_t)sn.wout_rows * sn.KP * 4;
    return (b + 127) & ~(size_t)127;
}

#define SM_SMEM_LIMIT (227 * 1024)


bool dx_mlp_small_supported(const NetDev& net, const DomainDev& d) {
    if (const char* c = getenv("DXB200_SMALL_NET")) if (c[0] == '0') return false;
    if (!(net.H <= 256 && net.act_depth == 0 && net.in_count <= 256 && 2 * net.num_blocks <= SM_MAX_LAYERS &&
          net.depth >= 1 && net.depth <= 255 && d.SP % 16 == 0 && net.in_count <= d.S + d.SP))
        return false;
    SmallNetDev sn = {};
    sn.KP = ((net.H + 15) / 16) * 16; sn.out_dim = net.out_dim; sn.nwa = (net.H + 15) / 16;
    sn.ks0 = ((net.in_dim_p / 16 + 3) / 4) * 4; sn.ksh = ((sn.KP / 16 + 3) / 4) * 4;      // (whole groups; G = 4 covers G = 2)
    sn.wout_rows = (net.out_dim > 1 && (size_t)net.out_dim * sn.KP * 4 <= 16 * 1024) ? net.out_dim : 0;
    return small_smem_bytes(sn, d, 2, true) <= SM_SMEM_LIMIT;      // (the largest configuration)
}

static float small_split_scale(const float* w, size_t rows, size_t cols, size_t ld) {
    float m = 0.f;
    for (size_t r = 0; r < rows; ++r)
        for (size_t c = 0; c < cols; ++c) m = fmaxf(m, fabsf(w[r * ld + c]));
    if (!(m > 0.f) || !std::isfinite(m)) return 1.f;
    int e = 0;
    frexpf(m, &e);                          // largest |w| * scale lands in [2^13, 2^14): every lo half is a normal fp16 number
    return ldexpf(1.f, 14 - e);
}

// W: [rows >= n_out, ld] fp32 ([out, in]); writes `ks` steps of B fragments for `nwa` warps at `dst` (step-major).
// mma.m16n8k16 B fragment (col-major B[k][n] = W[n][k]): lane (g = lane / 4, t = lane % 4) holds, per n8 tile, b0 =
// {k = 2t, 2t + 1}, b1 = {k = 2t + 8, 2t + 9} of column n = g; a lane's 16-byte chunk is {b0, b1} of the warp's two tiles.
static void small_pack_layer(const float* W, size_t ld, int n_out, int k_in, int ks, int nwa, bool split, float scale, uint16_t* dst) {
    const int NP = split ? 2 : 1;
    auto conv = [&](float x, int plane) -> uint16_t {
        if (!split) { __nv_bfloat16 b = __float2bfloat16_rn(x); uint16_t u; memcpy(&u, &b, 2); return u; }
        const float xs = x * scale;
        const __half hi = __float2half_rn(xs);
        if (plane == 0) return __half_as_ushort(hi);
        return __half_as_ushort(__float2half_rn(xs - __half2float(hi)));
    };
    for (int s = 0; s < ks; ++s)                      // (steps are contiguous: [step][16 warps][plane][lane], groups = runs of G steps)
        for (int w = 0; w < nwa; ++w)
            for (int p = 0; p < NP; ++p)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, t = lane & 3;
                    uint16_t* o = dst + ((((size_t)s * SM_CWARPS + w) * NP + p) * 32 + lane) * 8;
                    for (int j = 0; j < SM_NT; ++j) {
                        const int nidx = w * 16 + j * 8 + g;
                        for (int bi = 0; bi < 2; ++bi)
                            for (int kk = 0; kk < 2; ++kk) {
                                const int k = s * 16 + 2 * t + bi * 8 + kk;
                                const float x = (nidx < n_out && k < k_in) ? W[(size_t)nidx * ld + k] : 0.f;
                                o[(j * 2 + bi) * 2 + kk] = conv(x, p);
                            }
                    }
                }
}

// (Re)builds the packed weight stream from the host fp32 parameters (BN folded): w0p [>= H, in_dim_p], wres [L, Hp, Hp].
// The first call allocates, later calls (weight reload) overwrite in place.
int dx_mlp_small_upload(cudaStream_t s, NetDev& net, bool split, const float* w0p, const float* wres) {
    SmallNetHost* sh = (SmallNetHost*)net.small;
    const int L = 2 * net.num_blocks, H = net.H;
    const int NP = split ? 2 : 1, G = split ? 2 : 4;
    const int KP = ((H + 15) / 16) * 16, nwa = (H + 15) / 16;
    const int ks0 = ((net.in_dim_p / 16 + G - 1) / G) * G, ksh = ((KP / 16 + G - 1) / G) * G;   // every layer = whole groups
    const size_t steps = (size_t)ks0 + (size_t)L * ksh;
    const size_t n_groups = steps / G;
    const size_t step_elems = (size_t)SM_CWARPS * NP * 32 * 8;
    const size_t elems = steps * step_elems;
    if (!sh) {
        sh = new SmallNetHost();
        memset(sh, 0, sizeof(*sh));
        DX_CUDA(cudaMalloc(&sh->wstream, elems * 2));
        sh->stream_bytes = elems * 2;
        sh->split = split ? 1 : 0;
        net.small = sh;
    }
    SmallNetDev& sn = sh->dev;
    sn.wstream = (const uint4*)sh->wstream;
    sn.b0 = net.b0; sn.bres = net.bres; sn.wout = net.wout; sn.bout = net.bout;
    sn.in_count = net.in_count; sn.depth = net.depth;
    sn.ks0 = ks0; sn.ksh = ksh; sn.L = L; sn.H = H; sn.Hp = net.Hp; sn.KP = KP; sn.nwa = nwa; sn.out_dim = net.out_dim;
    sn.n_groups = (int)n_groups;
    sn.wout_rows = (net.out_dim > 1 && (size_t)net.out_dim * KP * 4 <= 16 * 1024) ? net.out_dim : 0;
    std::vector<uint16_t> buf(elems);
    float sc = split ? small_split_scale(w0p, H, net.in_dim, net.in_dim_p) : 1.f;
    sn.inv_scale[0] = 1.f / sc;
    small_pack_layer(w0p, net.in_dim_p, H, net.in_dim, ks0, nwa, split, sc, buf.data());
    for (int l = 0; l < L; ++l) {
        const float* W = wres + (size_t)l * net.Hp * net.Hp;
        sc = split ? small_split_scale(W, H, H, net.Hp) : 1.f;
        sn.inv_scale[1 + l] = 1.f / sc;
        small_pack_layer(W, net.Hp, H, H, ksh, nwa, split, sc, buf.data() + ((size_t)ks0 + (size_t)l * ksh) * step_elems);
    }
    DX_CUDA(cudaMemcpyAsync(sh->wstream, buf.data(), elems * 2, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaStreamSynchronize(s));
    return DX_OK;
}

void dx_mlp_small_free(NetDev& net) {
    SmallNetHost* sh = (SmallNetHost*)net.small;
    if (!sh) return;
    if (sh->wstream) cudaFree(sh->wstream);
    delete sh;
    net.small = nullptr;
}

template <int MT, bool SPLIT>
static int small_launch_t(cudaStream_t s, const SmallNetDev& sn, const DomainDev& d, const Ctl* ctl, const uint8_t* rows,
                          const uint8_t* goals, float* out_h, float* out_q, int absolute, long long max_rows) {
    const size_t smem = small_smem_bytes(sn, d, MT, SPLIT);
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_mlp_small<MT, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM_LIMIT) != cudaSuccess) {
            dx_set_error("k_mlp_small: cannot raise the dynamic shared-memory limit");
            return 0;
        }
        attr_set = true;
    }
    if (smem > SM_SMEM_LIMIT) { dx_set_error("k_mlp_small: network too large for the fused small-evaluation kernel"); return 0; }
    const int grid = (int)std::max<long long>(1, std::min<long long>(dx_div_up(max_rows, 16 * MT), 148 * 4));
    dx_launch(k_mlp_small<MT, SPLIT>, grid, SM_THREADS, smem, s, sn, d, ctl, rows, goals, out_h, out_q, absolute);
    return 1;
}

// The whole network over the current job (ctl->nn_row_start / nn_n) in one launch.  max_rows bounds nn_n (grid size).
int dx_launch_mlp_small(cudaStream_t s, const DomainDev& d, const NetDev& net, const Ctl* ctl, const uint8_t* rows,
                        const uint8_t* goals, float* out_h, float* out_q, int absolute, long long max_rows) {
    const SmallNetHost* sh = (const SmallNetHost*)net.small;
    if (!sh) { dx_set_error("small-evaluation network not prepared"); return 0; }
    static int mt_env = -1;
    if (mt_env < 0) { const char* c = getenv("DXB200_SMALL_NET_MT"); mt_env = c ? atoi(c) : 0; }
    // 16-row tiles while they still fit one wave of CTAs (latency), 32-row tiles beyond (half the weight traffic per row)
    const bool mt2 = mt_env == 2 || (mt_env != 1 && max_rows > 16 * 148);
    if (sh->split)
        return mt2 ? small_launch_t<2, true>(s, sh->dev, d, ctl, rows, goals, out_h, out_q, absolute, max_rows)
                   : small_launch_t<1, true>(s, sh->dev, d, ctl, rows, goals, out_h, out_q, absolute, max_rows);
    return mt2 ? small_launch_t<2, false>(s, sh->dev, d, ctl, rows, goals, out_h, out_q, absolute, max_rows)
               : small_launch_t<1, false>(s, sh->dev, d, ctl, rows, goals, out_h, out_q, absolute, max_rows);
}

// debug: read and clear the SM_PROFILE timeline (zeros unless built with -DSM_PROFILE)
int dx_small_profile_read(unsigned long long* out16) {
    if (cudaMemcpyFromSymbol(out16, g_small_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return DX_ERR_CUDA;
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(g_small_prof, z, sizeof(z)) != cudaSuccess) return DX_ERR_CUDA;
    return DX_OK;
}
