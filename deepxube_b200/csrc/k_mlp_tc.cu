// resnet_fc cost-to-go network on the 5th-generation tensor cores (sm_100a): bf16 operands, fp32
// accumulation in TMEM.  Same network as k_mlp_f32.cu (deepxube/nnets/resnet_fc.py:19-52, eval-mode BN
// folded at load time); only the arithmetic type of the H x H contractions changes.
//
// One GEMM kernel, launched once per layer:  C[M, N] = act(A[M, K] . W[N, K]^T + bias (+ R))
//   * CTA pairs (cluster of 2, `tcgen05.mma.cta_group::2`): a pair owns a 256 x BN output tile; each CTA holds its
//     128 A rows and HALF of the weight tile (BN/2 rows) in shared memory, so a k-block costs each SM 32 KB of TMA
//     writes and 8 KB of operand reads per MMA instead of 48 KB / 12 KB -- shared-memory bandwidth, not L2 or
//     HBM, is what capped the single-CTA version (measured: DESIGN.md section 5),
//   * A (activations, K-major) and W (nn.Linear weight [out, in] = K-major) are streamed by TMA (128-byte swizzle)
//     into a 5-stage shared-memory ring; both CTAs' loads complete on the LEADER's barrier,
//   * one thread of the leader CTA issues the MMAs (256 x BN x 16 per instruction) into one of two TMEM
//     accumulator stages; commits are multicast to both CTAs' barriers,
//   * eight epilogue warps per CTA (two per TMEM lane quarter, alternating 32-column chunks) drain their CTA's 128
//     accumulator rows with tcgen05.ld (pipelined one chunk ahead), add bias / residual, apply ReLU and write bf16 through 64-byte-swizzled staging buffers and TMA
//     (bulk tensor store; residual chunks arrive by TMA load TC_NBUF chunks ahead),
//   * persistent clusters walk the tiles; M is read from the device control block.
// CL = 1 keeps the single-CTA variant (128 x BN tiles, cta_group::1) for A/B measurements.
// The one-hot input layer is a K = in_dim (padded to 64) GEMM over a bf16 one-hot matrix written by
// k_onehot_bf16 (exact: the entries are 0 / 1).
#include <cuda.h>
#include <stdlib.h>
#include <math.h>
#include <cmath>
#include <algorithm>
#include <vector>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "dx_internal.cuh"

#define TC_BM 128
#define TC_BK 64
#ifndef TC_NBUF
#define TC_NBUF 2            // epilogue staging buffers per warp, for outputs and for residual chunks (power of 2)
#endif
#if TC_NBUF == 2
#define TC_NBUF_M1 "1"
#elif TC_NBUF == 4
#define TC_NBUF_M1 "3"
#else
#error "TC_NBUF must be 2 or 4"
#endif
#define TC_THREADS 384         // 4 control warps (TMA, MMA, TMEM alloc, spare) + 8 epilogue warps
#define TC_THREADS_GEN 448     // fused input layer: + 2 warps; warps 2, 3, 12, 13 generate the one-hot A tiles
#define TC_EPI_WARPS 8
#define TC_UMMA_K 16
#define TC_PEER_MASK 0xFEFFFFFFu     // clears the CTA-rank bit of a shared::cluster address -> the leader CTA's copy

struct TcPlan {
    CUtensorMap a_onehot;
    CUtensorMap a_x[3];
    CUtensorMap a_xh[3];         // the same buffers in 64-row boxes (clusters of two pairs: each CTA fetches half of an A tile)
    CUtensorMap b_w0;
    CUtensorMap b_res;
    CUtensorMap b_out;           // zero-padded output-layer weights [128, Hp] (multi-output networks), boxes of 128 / pair rows
    int q_out;                   // the multi-output layer runs on the tensor cores (DXB200_QOUT=0: warp-per-state kernel)
    CUtensorMap c_x[4];          // 32 x 32 boxes (64-byte swizzle) over the activation buffers: C stores / R loads; [3] = buffers 0 and 1
                                 // as one [rows, 2 * Hp] output (fused first block)
    int fuse_l1;                 // the generated-input launch also computes the first hidden layer (NetDev::tc_l1)
    __nv_bfloat16* onehot;
    long long max_rows;          // capacity of the activation buffers: no launch ever has more rows than this
    int BN;
    int CL;                      // CTAs per cluster of the hidden layers: 1 = single-CTA tiles, 2 = cta_group::2 pairs,
                                 // 4 = two pairs sharing their A tiles by TMA multicast (DXB200_GEMM_CLUSTER)
    int CL_in;                   // ... of the input layer (min(CL, 2))
    int grid4;                   // CTAs launched with clusters of 4 (4 x the co-resident clusters the device reports)
    int fuse_out;                // fold a scalar output layer into the last hidden GEMM (DXB200_FUSE_OUT=0 disables)
    int fuse_in;                 // the one-hot input is generated inside the first GEMM (dx_mlp_bf16_gen_input; DXB200_FUSE_IN=0 disables)
    int num_sms;
    int split;                   // fp32-accurate mode: fp16 (hi, lo) operands, activation buffers / weights are [rows, 2 * K]
};

// Fused input layer (GEN_A): the A operand is the one-hot encoding of the state rows
// (OneHot.forward, deepxube/pytorch/pytorch_models.py:15-24), written straight into the swizzled operand slots by
// two generator warps -- the one-hot matrix never exists in HBM.
struct TcGen {
    const uint8_t* rows;         // node-store rows (state + pad + instance id), row stride SP
    const uint8_t* goals;        // [instances, SP] goal rows (inputs S .. in_count - 1 come from the goal)
    int S, SP, in_count, depth;
    int ones_col;                // >= 0: columns ones_col .. +2 are 1.0 (folded bias), else -1
    int relu_col0;               // ReLU only for output columns >= relu_col0 (fused first block: [x0 | relu(hidden)]); 0: all
};

// fp32-accurate mode (DX_HEUR_NET_FP32 on the tensor cores): every operand is an fp16 (hi, lo) pair with hi + lo == the
// scaled fp32 value to ~22 bits, and a contraction is three MMA passes into the same fp32 TMEM accumulator,
//   A.W ~= A_hi.W_lo + A_lo.W_hi + A_hi.W_hi           (the dropped A_lo.W_lo term is 2^-22 relative)
// The kernel's main loop is unchanged: a "phase" is one pass over the k-blocks with its own column offsets into the
// [hi | lo] operand arrays (activations [rows, 2*Hp], weights [out, 2*K]).  Weights are scaled by a power of two per
// layer (fp16 lo halves stay normal numbers), activations by 2^6; the epilogue undoes the scales exactly.
struct TcSplit {
    int nph;                 // MMA passes over the k-blocks (1: plain bf16 GEMM)
    int a_col[3], w_col[3];  // column offset (elements) of the A / W operand of each pass
    float acc_scale;         // accumulator -> true pre-activation (1 / (weight scale * activation scale))
    float res_scale;         // stored residual -> true value (1 / activation scale)
    float out_scale;         // true activation -> stored value (activation scale)
    int lo_col;              // column offset of the lo half in the output / residual buffers
    int f16;                 // operands are fp16 (else bf16)
};
#define TC_ACT_SCALE 64.0f

// Optional in-kernel wait accounting (build with -DTC_PROFILE): cycles the role threads spend blocked on each
// barrier, summed over CTAs.  [0] MMA waits operands (full), [1] MMA waits accumulator (tmem empty),
// [2] producer waits slot (empty), [3] epilogue waits accumulator (tmem full), [4] epilogue waits residual,
// [5] epilogue waits TMA-store drain, [6] epilogue total, [7] MMA total.
__device__ unsigned long long g_tc_prof[16];   // [8] epi tmem-load wait, [9] epi math+smem, [10] epi fence+sync, [11] epi TMA issue
#ifdef TC_PROFILE
#define TC_T0() const long long _t0 = clock64()
#define TC_ACC(var) var += clock64() - _t0
#else
#define TC_T0()
#define TC_ACC(var)
#endif

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// Bounded wait: try_wait suspends up to ~10 ms per attempt; a barrier that does not complete within ~4 s is a
// protocol bug, and the kernel traps (CUDA error on the host) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    for (int attempt = 0; attempt < 400; ++attempt) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
        if (done) return;
    }
    __trap();
}
// same, cluster-scope acquire: the data the barrier guards was written by threads of the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    for (int attempt = 0; attempt < 400; ++attempt) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity), "r"(0x989680u)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
        "}" ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
        "}" ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
// 2-CTA load: data lands in THIS CTA's shared memory, the bytes are counted on the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & TC_PEER_MASK), "r"(c0), "r"(c1)
                 : "memory");
}
// 2-CTA load multicast to the CTAs in `mask` (same shared-memory offset in each); every destination counts its bytes on
// the barrier at this offset in the leader of ITS pair
__device__ __forceinline__ void tma_load_2d_2sm_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & TC_PEER_MASK), "r"(c0), "r"(c1), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int PAIR>
__device__ __forceinline__ void umma_commit(uint32_t bar, uint16_t mask) {
    if (PAIR == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else       // arrives on the barrier at this offset in every CTA of `mask`
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"(mask) : "memory");
}
template <int PAIR>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (PAIR == 1)
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

// K-major operand tile with 128-byte swizzle: rows of 64 bf16 (128 B), 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

template <int BN, int CL>
struct TcSmem {
    static constexpr int PAIR = CL >= 2 ? 2 : 1;          // CTAs per MMA (cta_group)
    static constexpr int STAGES = CL >= 2 ? 5 : 3;
    static constexpr int A_BYTES = TC_BM * TC_BK * 2;
    static constexpr int B_BYTES = (BN / PAIR) * TC_BK * 2;        // a pair CTA holds half of the weight tile
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_OFF = STAGES * STAGE_BYTES;
    static constexpr int EPI_BUF = 32 * 64;              // one 32-row x 32-column bf16 chunk
    // epilogue staging: TC_NBUF output buffers per warp, then (contiguous, 32 KB) TC_NBUF residual buffers per warp;
    // the fused input layer has no residual and keeps its tile's state rows in the residual region instead
    static constexpr int EPI_OUT_BYTES = TC_EPI_WARPS * TC_NBUF * EPI_BUF;
    static constexpr int RES_OFF = EPI_OFF + EPI_OUT_BYTES;
    static constexpr int RES_BYTES = TC_EPI_WARPS * TC_NBUF * EPI_BUF;
    static constexpr int BAR_OFF = RES_OFF + RES_BYTES;
    static constexpr int TOTAL = BAR_OFF + 512 + 1024;   // barriers + tmem ptr, + slack for 1024-byte alignment
};

// FUSE_OUT = true (last hidden layer of a scalar-output network): the epilogue does not store the activations; every
// thread dots its row's 128 activated columns with the output layer's weight row and writes one fp32 partial
// (partial[(2 * n_tile + h) * pstride + row]); k_out_final_bf16 adds the partials in a fixed order.
// GEN_A = true (input layer): no A tensor map is read; four generator warps write the one-hot A tiles (TcGen).
// OUT_MODE 0: store the bf16 activations; 1 (= FUSE_OUT, above); 2 (= Q_OUT): the launch IS the output layer of a
// multi-output network (Q*: out_dim = #actions <= 32, weight rows zero-padded to the n-tile): the epilogue adds the fp32
// bias, clips at 0 and writes the first qdim columns as fp32 q-values plus their minimum (node.heuristic) -- no bf16
// rounding of the outputs, no activation store.
template <int BN, int CL, int OUT_MODE, bool GEN_A, bool SPLIT = false>
__global__ void __launch_bounds__(GEN_A ? TC_THREADS_GEN : TC_THREADS, 1)
k_gemm_bf16_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const Ctl* __restrict__ ctl,
               int w_row0, int K, int N, const float* __restrict__ bias, int has_res, int relu_all,
               const float* __restrict__ wout, float* __restrict__ partial, long long pstride, const TcGen gen,
               float* __restrict__ out_h, int qdim, int absolute, const TcSplit sp, int relu_col0) {
    extern __shared__ uint8_t smem_raw[];
    static_assert(!(SPLIT && (GEN_A || CL != 2)), "the split-fp16 mode runs in CTA pairs with a TMA-fed A operand");
    const int nph = (SPLIT || GEN_A) ? sp.nph : 1;
    using SM = TcSmem<BN, CL>;
    constexpr int STAGES = SM::STAGES;
    constexpr bool FUSE_OUT = OUT_MODE == 1, Q_OUT = OUT_MODE == 2;
    constexpr int PAIR = SM::PAIR;           // CTAs per MMA
    constexpr int NP = CL / PAIR;            // pairs per cluster; NP = 2: two pairs work on adjacent n-tiles of the same rows
                                             // and fetch their common A tiles once, by TMA multicast
    static_assert(!(GEN_A && CL > 2), "the generated input layer runs in single CTAs or pairs");
    if (ctl->error) return;
    const int M = (int)ctl->nn_n;
    if (Q_OUT && blockIdx.x == 0 && threadIdx.x == 0)
        atomicAdd((unsigned long long*)&ctl->heur_evals, (unsigned long long)M);
    if (M == 0) return;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + SM::BAR_OFF;
    // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], residual[4][TC_NBUF]; then the TMEM base address
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    auto res_bar = [&](int ew, int b) { return bar_base + 8u * (2 * STAGES + 4 + TC_NBUF * ew + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4 + TC_EPI_WARPS * TC_NBUF);
    auto rows_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 5 + TC_EPI_WARPS * TC_NBUF + b); };   // GEN_A: state rows landed
    // GEN_A: the generated A tile of k-block kb (kb < 8) is complete / has been read by the MMAs of the row group's last n-tile
    auto afull_bar = [&](int kb) { return bar_base + 8u * (2 * STAGES + 7 + TC_EPI_WARPS * TC_NBUF + kb); };
    auto aempty_bar = [&](int kb) { return bar_base + 8u * (2 * STAGES + 15 + TC_EPI_WARPS * TC_NBUF + kb); };
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t cta_rank = 0;
    if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const uint32_t hr = cta_rank & (PAIR - 1);        // which 128-row half of the pair's tile
    const uint32_t pr = cta_rank / PAIR;              // which pair of the cluster
    const uint32_t leader = cta_rank - hr;            // the pair's MMA-issuing CTA
    const uint16_t pair_mask = (uint16_t)(((1u << PAIR) - 1u) << leader), all_mask = (uint16_t)((1u << CL) - 1u);
    if (threadIdx.x == 0) {
        // full: the leader's expect_tx arrival
        // empty: one MMA commit per pair of the cluster (a multicast A tile lands in the shared memory of both pairs)
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), NP); }
        // tmem_empty (leader's copy is the one the MMA thread waits on): one arrival per epilogue warp of the pair
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), TC_EPI_WARPS * PAIR); }
        for (int ew = 0; ew < TC_EPI_WARPS; ++ew)
            for (int b = 0; b < TC_NBUF; ++b) mbar_init(res_bar(ew, b), 1);
        if (GEN_A) {
            mbar_init(rows_bar(0), 1); mbar_init(rows_bar(1), 1);
            for (int kb = 0; kb < 8; ++kb) { mbar_init(afull_bar(kb), 4 * PAIR); mbar_init(aempty_bar(kb), 1); }   // one arrival per generator warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {      // the same warp in both CTAs of a pair
        if (PAIR == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * BN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * BN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // peer barriers / TMEM exist before anything is signalled across CTAs
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int m_tiles = (M + TC_BM - 1) / TC_BM, n_tiles = N / BN;
    const int nq = n_tiles / NP;                                   // work items per row group: NP adjacent n-tiles each
    const int total = ((m_tiles + PAIR - 1) / PAIR) * nq;          // row groups of PAIR vertically adjacent 128-row tiles
    auto tile_m0 = [&](int t) { return ((t / nq) * PAIR + (int)hr) * TC_BM; };
    auto tile_n0 = [&](int t) { return ((t % nq) * NP + (int)pr) * BN; };
    const int kblocks = K / TC_BK;
    const int cl_id = blockIdx.x / CL, n_cl = gridDim.x / CL;
    // GEN_A (input layer): a cluster takes whole row groups; the generated A tiles of a group (kblocks x 16 KB) stay resident
    // in the operand region for all n-tiles of the group, and what is left of the region is the ring of weight tiles.
    const int groups = (m_tiles + PAIR - 1) / PAIR;
    const int a_res_bytes = GEN_A ? kblocks * SM::A_BYTES : 0;
    const int nb_st = GEN_A ? min(STAGES, (STAGES * SM::STAGE_BYTES - a_res_bytes) / SM::B_BYTES) : STAGES;
    // k-th tile of this cluster, in the order every role walks them
    auto seq_tile = [&](int k, int& m0, int& n0) -> bool {
        if (GEN_A) {
            const int g = cl_id + (k / n_tiles) * n_cl;
            if (g >= groups) return false;
            m0 = (g * PAIR + (int)hr) * TC_BM;
            n0 = (k % n_tiles) * BN;
            return true;
        }
        const int t = cl_id + k * n_cl;
        if (t >= total) return false;
        m0 = tile_m0(t);
        n0 = tile_n0(t);
        return true;
    };

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer (every CTA: its A rows, its share of the weight tile) ----------------
            uint32_t it = 0;
            [[maybe_unused]] long long w_empty = 0;
            if (GEN_A) {       // weight tiles only: ring of nb_st slots behind the resident A tiles
                uint32_t s = 0, ph = 0;
                int m0, n0;
                for (int k = 0; seq_tile(k, m0, n0); ++k)
                    for (int wp = 0; wp < nph; ++wp)               // weight passes: [lo | hi] halves of hi/lo-split weights, else one
                    for (int kb = 0; kb < kblocks; ++kb) {
                        mbar_wait(empty_bar(s), ph ^ 1u);
                        const uint32_t b_dst = smem_base + a_res_bytes + s * SM::B_BYTES;
                        const int w_c = sp.w_col[wp] + kb * TC_BK;
                        if (PAIR == 1) {
                            mbar_arrive_expect_tx(full_bar(s), SM::B_BYTES);
                            tma_load_2d(b_dst, &tmB, full_bar(s), w_c, w_row0 + n0);
                        } else {
                            if (hr == 0) mbar_arrive_expect_tx(full_bar(s), 2 * SM::B_BYTES);
                            tma_load_2d_2sm(b_dst, &tmB, full_bar(s), w_c, w_row0 + n0 + (int)hr * (BN / PAIR));
                        }
                        if (++s == (uint32_t)nb_st) { s = 0; ph ^= 1u; }
                    }
            }
            for (int t = cl_id; !GEN_A && t < total; t += n_cl) {
                const int m0 = tile_m0(t), n0 = tile_n0(t);
                for (int ph = 0; ph < nph; ++ph)
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    { TC_T0(); mbar_wait(empty_bar(s), ((it / STAGES) & 1u) ^ 1u); TC_ACC(w_empty); }
                    const uint32_t a_dst = smem_base + s * SM::STAGE_BYTES, b_dst = a_dst + SM::A_BYTES;
                    constexpr int TX = SM::STAGE_BYTES;
                    const int a_c = (SPLIT ? sp.a_col[ph] : 0) + kb * TC_BK, w_c = (SPLIT ? sp.w_col[ph] : 0) + kb * TC_BK;
                    if (PAIR == 1) {
                        mbar_arrive_expect_tx(full_bar(s), TX);
                        tma_load_2d(a_dst, &tmA, full_bar(s), a_c, m0);
                        tma_load_2d(b_dst, &tmB, full_bar(s), w_c, w_row0 + n0);
                    } else {
                        // the pair leader's barrier counts the bytes of both CTAs (2 x STAGE_BYTES per phase)
                        if (hr == 0) mbar_arrive_expect_tx(full_bar(s), 2 * TX);
                        if (NP == 2) {
                            // CTAs hr and hr + 2 hold the same 128 A rows: each fetches 64 of them (tmA has 64-row boxes here)
                            // and multicasts them into both
                            tma_load_2d_2sm_mc(a_dst + pr * (SM::A_BYTES / 2), &tmA, full_bar(s), a_c, m0 + (int)pr * (TC_BM / 2),
                                               (uint16_t)((1u << hr) | (1u << (hr + 2))));
                        } else {
                            tma_load_2d_2sm(a_dst, &tmA, full_bar(s), a_c, m0);
                        }
                        tma_load_2d_2sm(b_dst, &tmB, full_bar(s), w_c, w_row0 + n0 + (int)hr * (BN / PAIR));
                    }
                }
            }
#ifdef TC_PROFILE
            atomicAdd(&g_tc_prof[2], (unsigned long long)w_empty);
#endif
        }
    } else if (warp == 1) {
        if (lane == 0 && hr == 0) {
            // ---------------- MMA issuer (the leader CTA of each pair) ----------------
            // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128 * PAIR
            // (A / B format field: 1 = bf16, 0 = fp16 -- the split-fp16 mode's operands)
            const uint32_t ab_fmt = ((SPLIT || GEN_A) && sp.f16) ? 0u : ((1u << 7) | (1u << 10));
            const uint32_t idesc = (1u << 4) | ab_fmt | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((TC_BM * PAIR) >> 4) << 24);
            uint32_t it = 0, lt = 0;
            [[maybe_unused]] long long w_full = 0, w_tempty = 0;
#ifdef TC_PROFILE
            const long long t_mma0 = clock64();
#endif
            if (GEN_A) {
                uint32_t s = 0, ph = 0;
                int m0, n0;
                for (int k = 0; seq_tile(k, m0, n0); ++k, ++lt) {
                    const int as = lt & 1, nt = k % n_tiles;
                    const uint32_t gph = (uint32_t)(k / n_tiles) & 1u;          // phase of the row group's A barriers
                    { TC_T0(); mbar_wait(tempty_bar(as), ((lt >> 1) & 1u) ^ 1u); TC_ACC(w_tempty); }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                    for (int wp = 0; wp < nph; ++wp)
                    for (int kb = 0; kb < kblocks; ++kb) {
                        { TC_T0(); mbar_wait(full_bar(s), ph); if (nt == 0 && wp == 0) mbar_wait(afull_bar(kb), gph); TC_ACC(w_full); }
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint64_t adesc = make_sw128_desc(smem_base + kb * SM::A_BYTES);
                        const uint64_t bdesc = make_sw128_desc(smem_base + a_res_bytes + s * SM::B_BYTES);
#pragma unroll
                        for (int q = 0; q < TC_BK / TC_UMMA_K; ++q)
                            umma_bf16<PAIR>(d_tmem, adesc + (uint64_t)(q * 2), bdesc + (uint64_t)(q * 2), idesc, (wp | kb | q) ? 1u : 0u);
                        umma_commit<PAIR>(empty_bar(s), all_mask);
                        // the group's A tile kb may be overwritten once its last use (last weight pass of the last n-tile) has been read
                        if (nt == n_tiles - 1 && wp == nph - 1) umma_commit<PAIR>(aempty_bar(kb), pair_mask);
                        if (++s == (uint32_t)nb_st) { s = 0; ph ^= 1u; }
                    }
                    umma_commit<PAIR>(tfull_bar(as), pair_mask);
                }
            }
            for (int t = cl_id; !GEN_A && t < total; t += n_cl, ++lt) {
                const int as = lt & 1;
                { TC_T0(); mbar_wait(tempty_bar(as), ((lt >> 1) & 1u) ^ 1u); TC_ACC(w_tempty); }
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int ph = 0; ph < nph; ++ph)
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % STAGES;
                    { TC_T0(); mbar_wait(full_bar(s), (it / STAGES) & 1u); TC_ACC(w_full); }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_src = smem_base + s * SM::STAGE_BYTES, b_src = a_src + SM::A_BYTES;
                    const uint64_t adesc = make_sw128_desc(a_src), bdesc = make_sw128_desc(b_src);
#pragma unroll
                    for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
                        umma_bf16<PAIR>(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (ph | kb | k) ? 1u : 0u);
                    umma_commit<PAIR>(empty_bar(s), all_mask);    // slot reusable (told to every CTA of the cluster) once these MMAs have read it
                }
                umma_commit<PAIR>(tfull_bar(as), pair_mask);      // accumulator complete (signalled to both CTAs of the pair)
            }
#ifdef TC_PROFILE
            atomicAdd(&g_tc_prof[0], (unsigned long long)w_full);
            atomicAdd(&g_tc_prof[1], (unsigned long long)w_tempty);
            atomicAdd(&g_tc_prof[7], (unsigned long long)(clock64() - t_mma0));
#endif
        }
    } else if (GEN_A && (warp == 2 || warp == 3 || warp >= 12)) {
        // ---------------- A generator (4 warps, 128 threads): thread gt writes the one-hot encoding of row gt ----------------
        // Residual region (unused by this layer): [256-entry LUT: byte -> 8 bf16 0/1 values][rows of tile t][rows of t+1][ranges].
        // A row's 64 features of a k-block are collected into a 64-bit mask, then expanded 8 features at a time through
        // the LUT: one 16-byte load + one 16-byte store per swizzled operand chunk, no zero-fill pass.
        const int gt = (warp >= 12 ? (warp - 10) : (warp - 2)) * 32 + lane;
        uint8_t* res_sm = smem_raw + (smem_base + SM::RES_OFF - smem_u32(smem_raw));
        uint4* lut = reinterpret_cast<uint4*>(res_sm);
        const uint32_t one16 = sp.f16 ? 0x3C00u : 0x3F80u;         // 1.0 in the operand format (fp16 / bf16)
        // LUT slot of byte e: low 3 bits permuted so that the few byte values a one-hot code produces spread over banks
        auto lut_slot = [](uint32_t e) { return e ^ (((e >> 3) ^ (e >> 6)) & 7u); };
        const int SP = gen.SP, depth = gen.depth;
        uint8_t* rows_buf0 = res_sm + 4096;                              // buffer b at rows_buf0 + b * TC_BM * SP
        const uint32_t row_start = ctl->nn_row_start;
        // input range [i_lo, i_hi] whose features fall into k-block kb (kblocks <= 64), computed once: no divisions in the loop
        int* irange = reinterpret_cast<int*>(rows_buf0 + 2 * TC_BM * SP);
        for (int kb = gt; kb < kblocks; kb += 128) {
            irange[2 * kb] = (kb * TC_BK) / depth;
            irange[2 * kb + 1] = min(gen.in_count - 1, (kb * TC_BK + TC_BK - 1) / depth);
        }
        for (uint32_t e = gt; e < 256; e += 128) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) w[k] = (((e >> (2 * k)) & 1) ? one16 : 0u) | (((e >> (2 * k + 1)) & 1) ? (one16 << 16) : 0u);
            lut[lut_slot(e)] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        // One bulk copy (TMA, completes on rows_bar) brings a tile's rows -- contiguous in the node store -- into buffer b.
        // It is tracked by the barrier, not by this thread's memory-operation scoreboard, so the proxy fences below do
        // not wait for it (a cp.async / ld.global prefetch would stall every fence until the rows arrived).
        auto prefetch_rows = [&](int g, int b) {       // rows of row group g (this CTA's half)
            const int m0 = (g * PAIR + (int)hr) * TC_BM;
            const int nrows = min(TC_BM, M - m0);                      // rows past M keep stale bytes: their outputs are unused
            if (gt == 0) {
                if (nrows > 0) {
                    const uint32_t bytes = (uint32_t)nrows * (uint32_t)SP;
                    mbar_arrive_expect_tx(rows_bar(b), bytes);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(rows_buf0 + b * TC_BM * SP)), "l"(gen.rows + (size_t)(row_start + m0) * SP), "r"(bytes),
                                   "r"(rows_bar(b)) : "memory");
                } else {
                    mbar_arrive(rows_bar(b));
                }
            }
        };
        if (cl_id < groups) prefetch_rows(cl_id, 0);
        uint32_t lt = 0;                                 // row groups done by this cluster
        [[maybe_unused]] long long w_gslot = 0, w_gmask = 0, w_gexp = 0, w_gsync = 0;
#ifdef TC_PROFILE
        const long long t_gen0 = clock64();
#endif
        const int r = gt;
        const uint32_t arow_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);      // SW128: chunk j of row r at j ^ (r & 7)
        for (int g = cl_id; g < groups; g += n_cl, ++lt) {
            asm volatile("bar.sync 1, 128;" ::: "memory");         // everyone is done with the other buffer (and the LUT is written)
            if (g + n_cl < groups) prefetch_rows(g + n_cl, (lt + 1) & 1);
            mbar_wait(rows_bar(lt & 1), (lt >> 1) & 1u);           // this group's rows have landed
            const uint8_t* srow = rows_buf0 + (lt & 1) * TC_BM * SP + r * SP;
            for (int kb = 0; kb < kblocks; ++kb) {
                const int f0 = kb * TC_BK;
                const int i_lo = irange[2 * kb], i_hi = irange[2 * kb + 1];
#ifdef TC_PROFILE
                const long long _g0 = clock64();
#endif
                // Feature mask of this row's 64 columns as two 32-bit halves.  Four inputs per shared-memory word; input i
                // sets bit i * depth + v - f0.  shl.b32 clamps: a shift amount outside [0, 32) (as unsigned) gives 0, so
                // inputs of the neighbouring k-blocks drop out without range tests.  Row bytes past in_count are zero padding /
                // the instance id and only reach columns >= in_dim, whose weights are zero (or the always-one bias columns).
                uint32_t mlo = 0u, mhi = 0u;
                auto shl1 = [](int sh) { uint32_t o; asm("shl.b32 %0, %1, %2;" : "=r"(o) : "r"(1u), "r"((uint32_t)sh)); return o; };
#pragma unroll 2
                for (int i4 = i_lo & ~3; i4 <= i_hi; i4 += 4) {
                    const uint32_t w = *reinterpret_cast<const uint32_t*>(srow + i4);
                    const int fb = i4 * depth - f0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int f = fb + j * depth + min((int)((w >> (8 * j)) & 0xFFu), depth - 1);
                        mlo |= shl1(f);
                        mhi |= shl1(f - 32);
                    }
                }
                if (gen.ones_col >= 0) {
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const int f = gen.ones_col + j - f0;
                        mlo |= shl1(f);
                        mhi |= shl1(f - 32);
                    }
                }
#ifdef TC_PROFILE
                const long long _g1 = clock64();
                w_gmask += _g1 - _g0;
#endif
                mbar_wait(aempty_bar(kb), (lt & 1u) ^ 1u);      // the previous group's MMAs have read this resident tile
#ifdef TC_PROFILE
                const long long _g2 = clock64();
                w_gslot += _g2 - _g1;
#endif
                uint8_t* arow = smem_raw + (smem_base + kb * SM::A_BYTES - smem_u32(smem_raw)) + arow_off;
                uint4 ch[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) ch[j] = lut[lut_slot(((j < 4 ? mlo : mhi) >> (8 * (j & 3))) & 0xFFu)];
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(arow + ((j ^ (r & 7)) << 4)) = ch[j];
#ifdef TC_PROFILE
                const long long _g3 = clock64();
                w_gexp += _g3 - _g2;
#endif
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(afull_bar(kb), leader);       // the pair leader's barrier
#ifdef TC_PROFILE
                w_gsync += clock64() - _g3;
#endif
            }
        }
#ifdef TC_PROFILE
        if (gt == 0) {
            atomicAdd(&g_tc_prof[12], (unsigned long long)w_gslot);
            atomicAdd(&g_tc_prof[13], (unsigned long long)w_gmask);
            atomicAdd(&g_tc_prof[14], (unsigned long long)w_gexp);
            atomicAdd(&g_tc_prof[15], (unsigned long long)(clock64() - t_gen0));
            atomicAdd(&g_tc_prof[2], (unsigned long long)w_gsync);     // (slot [2] doubles as generator fence+sync+arrive)
        }
#endif
    } else if (warp >= 4 && warp < 4 + TC_EPI_WARPS) {
        // ---------------- epilogue: 8 warps; warp 4 + ew drains TMEM lane quarter (ew & 3) -- a warp may only touch
        // lanes 32 * (warp % 4) .. + 31 -- and the 32-column chunks c with (c & 1) == ew >> 2 ----------------
        const int ew = warp - 4, q = ew & 3, h = ew >> 2;
        const uint32_t epi_base = smem_base + SM::EPI_OFF + ew * (TC_NBUF * SM::EPI_BUF);      // this warp's output buffers
        const uint32_t res_base = smem_base + SM::RES_OFF + ew * (TC_NBUF * SM::EPI_BUF);      // ... residual buffers
        uint8_t* epi_ptr = smem_raw + (epi_base - smem_u32(smem_raw));
        const uint8_t* res_ptr = smem_raw + (res_base - smem_u32(smem_raw));
        // 64-byte swizzle (Swizzle<2,4,3>): 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3)
        const uint32_t sw = (uint32_t)((lane >> 1) & 3);
        constexpr int NCH = BN / 32;
        uint32_t rphase = 0;                     // bit b = parity the next wait on residual buffer b expects
        uint32_t lt = 0;
        [[maybe_unused]] long long w_tfull = 0, w_res = 0, w_store = 0, w_ld = 0, w_math = 0, w_fence = 0, w_issue = 0;
#ifdef TC_PROFILE
        const long long t_epi0 = clock64();
#endif
        int m0, n0;
        for (int k = 0; seq_tile(k, m0, n0); ++k, ++lt) {
            const int as = lt & 1;
            const int row0 = m0 + q * 32;
            const int relu = relu_all && n0 >= relu_col0;       // (fused first block: only the hidden-layer half of the columns)
            if (SPLIT) {
                // split mode: buffer 0 holds the hi half, buffer 1 the lo half of ONE chunk (both on barrier 0); the epilogue
                // has three MMA passes' worth of time per tile, so chunks are not double-buffered here
                if (has_res && lane == 0 && h < NCH) {
                    mbar_arrive_expect_tx(res_bar(ew, 0), 2 * SM::EPI_BUF);
                    tma_load_2d(res_base, &tmR, res_bar(ew, 0), n0 + h * 32, row0);
                    tma_load_2d(res_base + SM::EPI_BUF, &tmR, res_bar(ew, 0), sp.lo_col + n0 + h * 32, row0);
                }
            } else if (has_res && lane == 0) {          // this warp's first TC_NBUF residual chunks of the tile
#pragma unroll
                for (int b = 0; b < TC_NBUF && h + 2 * b < NCH; ++b) {
                    mbar_arrive_expect_tx(res_bar(ew, b), SM::EPI_BUF);
                    tma_load_2d(res_base + b * SM::EPI_BUF, &tmR, res_bar(ew, b), n0 + (h + 2 * b) * 32, row0);
                }
            }
            { TC_T0(); mbar_wait(tfull_bar(as), (lt >> 1) & 1u); TC_ACC(w_tfull); }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);

            float oacc = 0.f;                     // FUSE_OUT: this row's partial output-layer dot product
            float hmin = __int_as_float(0x7f800000);      // Q_OUT: min over the row's q-values
            const size_t orow = (size_t)(absolute ? ctl->nn_row_start : 0u) + (size_t)(row0 + lane);   // Q_OUT: output row
            auto process = [&](const uint32_t* v, int c) {
                const int b = (c >> 1) & (TC_NBUF - 1);       // this warp's chunks are c = h, h + 2, ...
                const int col = n0 + c * 32;
                // bias first: with ~225 KB of shared memory in use the L1 keeps almost nothing, so these are L2 round
                // trips; issued here, their latency overlaps the barrier waits below instead of stalling the math
                // (bias == nullptr: the bias is folded into the GEMM through the ones columns, NetDev::tc_fold)
                float4 bv[8];
                if (bias) {
                    const float4* b4 = reinterpret_cast<const float4*>(bias + col);
#pragma unroll
                    for (int g = 0; g < 8; ++g) bv[g] = __ldg(b4 + g);
                } else {
#pragma unroll
                    for (int g = 0; g < 8; ++g) bv[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float4 wv[FUSE_OUT ? 8 : 1];
                if (FUSE_OUT) {
                    const float4* w4 = reinterpret_cast<const float4*>(wout + col);
#pragma unroll
                    for (int g = 0; g < 8; ++g) wv[g] = __ldg(w4 + g);
                }
                if (has_res) {
                    TC_T0();
                    mbar_wait(res_bar(ew, b), (rphase >> b) & 1u);
                    TC_ACC(w_res);
                    rphase ^= 1u << b;
                }
                if (OUT_MODE == 0) {   // the TMA store that last read output buffer b (TC_NBUF chunks ago) must be done with it
                    TC_T0();
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read " TC_NBUF_M1 ";" ::: "memory");
                    __syncwarp();
                    TC_ACC(w_store);
                }
#ifdef TC_PROFILE
                const long long _tm0 = clock64();
#endif
                const uint8_t* rrow = res_ptr + b * SM::EPI_BUF + lane * 64;
                uint8_t* crow = epi_ptr + b * SM::EPI_BUF + lane * 64;
#pragma unroll
                for (int g = 0; g < 4; ++g) {      // 8 columns = one 16-byte chunk per group
                    float f[8];
                    const float4 ba = bv[2 * g], bb = bv[2 * g + 1];
                    f[0] = __uint_as_float(v[8 * g + 0]) + ba.x; f[1] = __uint_as_float(v[8 * g + 1]) + ba.y;
                    f[2] = __uint_as_float(v[8 * g + 2]) + ba.z; f[3] = __uint_as_float(v[8 * g + 3]) + ba.w;
                    f[4] = __uint_as_float(v[8 * g + 4]) + bb.x; f[5] = __uint_as_float(v[8 * g + 5]) + bb.y;
                    f[6] = __uint_as_float(v[8 * g + 6]) + bb.z; f[7] = __uint_as_float(v[8 * g + 7]) + bb.w;
                    const uint32_t pos = ((uint32_t)g ^ sw) << 4;
                    if (has_res) {
                        const uint4 rr = *reinterpret_cast<const uint4*>(rrow + pos);
                        const __nv_bfloat162* rb = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 rf = __bfloat1622float2(rb[e]);
                            f[2 * e] += rf.x; f[2 * e + 1] += rf.y;
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
                    }
                    if (Q_OUT) {
                        if (row0 + lane < M) {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (8 * g + e < qdim) {
                                    hmin = fminf(hmin, f[e]);
                                    partial[orow * (size_t)qdim + (size_t)(8 * g + e)] = f[e];
                                }
                        }
                    } else if (FUSE_OUT) {
                        const float4 wa = wv[FUSE_OUT ? 2 * g : 0], wb = wv[FUSE_OUT ? 2 * g + 1 : 0];
                        oacc = fmaf(f[0], wa.x, oacc); oacc = fmaf(f[1], wa.y, oacc);
                        oacc = fmaf(f[2], wa.z, oacc); oacc = fmaf(f[3], wa.w, oacc);
                        oacc = fmaf(f[4], wb.x, oacc); oacc = fmaf(f[5], wb.y, oacc);
                        oacc = fmaf(f[6], wb.z, oacc); oacc = fmaf(f[7], wb.w, oacc);
                    } else {
                        uint4 o;
                        __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                        for (int e = 0; e < 4; ++e) ob[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                        *reinterpret_cast<uint4*>(crow + pos) = o;
                    }
                }
#ifdef TC_PROFILE
                const long long _tm1 = clock64();
                w_math += _tm1 - _tm0;
#endif
                if (OUT_MODE == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to TMA
                __syncwarp();
#ifdef TC_PROFILE
                const long long _tm2 = clock64();
                w_fence += _tm2 - _tm1;
#endif
                if (lane == 0) {
                    if (OUT_MODE == 0) {
                        tma_store_2d(&tmC, epi_base + b * SM::EPI_BUF, col, row0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (has_res && c + 2 * TC_NBUF < NCH) {      // all lanes are past their reads of residual buffer b
                        mbar_arrive_expect_tx(res_bar(ew, b), SM::EPI_BUF);
                        tma_load_2d(res_base + b * SM::EPI_BUF, &tmR, res_bar(ew, b), n0 + (c + 2 * TC_NBUF) * 32, row0);
                    }
                }
#ifdef TC_PROFILE
                w_issue += clock64() - _tm2;
#endif
            };

            // ---- split-fp16 epilogue: true value = acc * acc_scale + bias (+ (r_hi + r_lo) * res_scale), ReLU; stored as the fp16
            // pair (hi, lo) of value * out_scale into columns [col, col + 32) and [lo_col + col, ...) of the output buffer ----
            bool bad = false;                     // a stored value left the fp16 range
            auto process_split = [&](const uint32_t* v, int c) {
                const int col = n0 + c * 32;
                float4 bv[8];
                if (bias) {
                    const float4* b4 = reinterpret_cast<const float4*>(bias + col);
#pragma unroll
                    for (int g = 0; g < 8; ++g) bv[g] = __ldg(b4 + g);
                } else {
#pragma unroll
                    for (int g = 0; g < 8; ++g) bv[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float4 wv[FUSE_OUT ? 8 : 1];
                if (FUSE_OUT) {
                    const float4* w4 = reinterpret_cast<const float4*>(wout + col);
#pragma unroll
                    for (int g = 0; g < 8; ++g) wv[g] = __ldg(w4 + g);
                }
                if (has_res) {
                    mbar_wait(res_bar(ew, 0), rphase & 1u);
                    rphase ^= 1u;
                }
                if (OUT_MODE == 0) {              // the two TMA stores of the previous chunk have read the staging buffers
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
                }
                const uint8_t* rrow = res_ptr + lane * 64;
                uint8_t* crow = epi_ptr + lane * 64;
                const float as = sp.acc_scale, rs = sp.res_scale, os = sp.out_scale;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float f[8];
                    const float4 ba = bv[2 * g], bb = bv[2 * g + 1];
                    f[0] = fmaf(__uint_as_float(v[8 * g + 0]), as, ba.x); f[1] = fmaf(__uint_as_float(v[8 * g + 1]), as, ba.y);
                    f[2] = fmaf(__uint_as_float(v[8 * g + 2]), as, ba.z); f[3] = fmaf(__uint_as_float(v[8 * g + 3]), as, ba.w);
                    f[4] = fmaf(__uint_as_float(v[8 * g + 4]), as, bb.x); f[5] = fmaf(__uint_as_float(v[8 * g + 5]), as, bb.y);
                    f[6] = fmaf(__uint_as_float(v[8 * g + 6]), as, bb.z); f[7] = fmaf(__uint_as_float(v[8 * g + 7]), as, bb.w);
                    const uint32_t pos = ((uint32_t)g ^ sw) << 4;
                    if (has_res) {
                        const uint4 rh = *reinterpret_cast<const uint4*>(rrow + pos);
                        const uint4 rl = *reinterpret_cast<const uint4*>(rrow + SM::EPI_BUF + pos);
                        const __half2* hh = reinterpret_cast<const __half2*>(&rh);
                        const __half2* hl = reinterpret_cast<const __half2*>(&rl);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 a = __half22float2(hh[e]), b = __half22float2(hl[e]);
                            f[2 * e] += (a.x + b.x) * rs; f[2 * e + 1] += (a.y + b.y) * rs;
                        }
                    }
                    if (relu) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
                    }
                    if (Q_OUT) {
                        if (row0 + lane < M) {
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                if (8 * g + e < qdim) {
                                    hmin = fminf(hmin, f[e]);
                                    partial[orow * (size_t)qdim + (size_t)(8 * g + e)] = f[e];
                                }
                        }
                    } else if (FUSE_OUT) {
                        const float4 wa = wv[FUSE_OUT ? 2 * g : 0], wb = wv[FUSE_OUT ? 2 * g + 1 : 0];
                        oacc = fmaf(f[0], wa.x, oacc); oacc = fmaf(f[1], wa.y, oacc);
                        oacc = fmaf(f[2], wa.z, oacc); oacc = fmaf(f[3], wa.w, oacc);
                        oacc = fmaf(f[4], wb.x, oacc); oacc = fmaf(f[5], wb.y, oacc);
                        oacc = fmaf(f[6], wb.z, oacc); oacc = fmaf(f[7], wb.w, oacc);
                    } else {
                        uint4 oh, ol;
                        __half2* ph2 = reinterpret_cast<__half2*>(&oh);
                        __half2* pl2 = reinterpret_cast<__half2*>(&ol);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float x0 = f[2 * e] * os, x1 = f[2 * e + 1] * os;
                            bad = bad || fabsf(x0) > 65000.f || fabsf(x1) > 65000.f;
                            const __half2 hi = __floats2half2_rn(x0, x1);
                            const float2 hf = __half22float2(hi);
                            ph2[e] = hi;
                            pl2[e] = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
                        }
                        *reinterpret_cast<uint4*>(crow + pos) = oh;
                        *reinterpret_cast<uint4*>(crow + SM::EPI_BUF + pos) = ol;
                    }
                }
                if (OUT_MODE == 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    if (OUT_MODE == 0) {
                        tma_store_2d(&tmC, epi_base, col, row0);
                        tma_store_2d(&tmC, epi_base + SM::EPI_BUF, sp.lo_col + col, row0);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    if (has_res && c + 2 < NCH) {            // all lanes are past their reads of the residual buffers
                        mbar_arrive_expect_tx(res_bar(ew, 0), 2 * SM::EPI_BUF);
                        tma_load_2d(res_base, &tmR, res_bar(ew, 0), n0 + (c + 2) * 32, row0);
                        tma_load_2d(res_base + SM::EPI_BUF, &tmR, res_bar(ew, 0), sp.lo_col + n0 + (c + 2) * 32, row0);
                    }
                }
            };
#define TC_PROCESS(v, c) do { if (SPLIT) process_split(v, c); else process(v, c); } while (0)

            uint32_t va[32], vb[32];
            if (Q_OUT) {                                   // the q-values live in the first 32 columns: one chunk, by the h = 0 warps
                if (h == 0) {
                    tmem_ld32(t_row, va);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    TC_PROCESS(va, 0);
                    if (out_h && row0 + lane < M) out_h[orow] = hmin;
                }
            } else {
            tmem_ld32(t_row + (uint32_t)(h * 32), va);
#pragma unroll 1
            for (int c = h; c < NCH; c += 4) {
                { TC_T0(); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); TC_ACC(w_ld); }
                if (c + 2 < NCH) tmem_ld32(t_row + (uint32_t)((c + 2) * 32), vb);
                TC_PROCESS(va, c);
                if (c + 2 < NCH) {
                    { TC_T0(); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); TC_ACC(w_ld); }
                    if (c + 4 < NCH) tmem_ld32(t_row + (uint32_t)((c + 4) * 32), va);
                    TC_PROCESS(vb, c + 2);
                }
            }
            }
#undef TC_PROCESS
            if (SPLIT && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(const_cast<uint32_t*>(&ctl->error), DX_DEVERR_NUMERIC);
            if (FUSE_OUT && row0 + lane < M) partial[(size_t)((n0 / BN) * 2 + h) * (size_t)pstride + (size_t)(row0 + lane)] = oacc;
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) {                      // this warp's quarter of the accumulator stage is free again
                if (PAIR == 1) mbar_arrive(tempty_bar(as));
                else mbar_arrive_cluster(tempty_bar(as), leader);  // the MMA thread lives in the pair's leader CTA
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef TC_PROFILE
        if (lane == 0 && ew == 0) {
            atomicAdd(&g_tc_prof[3], (unsigned long long)w_tfull);
            atomicAdd(&g_tc_prof[4], (unsigned long long)w_res);
            atomicAdd(&g_tc_prof[5], (unsigned long long)w_store);
            atomicAdd(&g_tc_prof[6], (unsigned long long)(clock64() - t_epi0));
            atomicAdd(&g_tc_prof[8], (unsigned long long)w_ld);
            atomicAdd(&g_tc_prof[9], (unsigned long long)w_math);
            atomicAdd(&g_tc_prof[10], (unsigned long long)w_fence);
            atomicAdd(&g_tc_prof[11], (unsigned long long)w_issue);
        }
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // no CTA exits (or frees TMEM) while its peer may still use / signal it
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (PAIR == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
    }
}

// ---------------------------------------------------------------------------------------------
// bf16 one-hot input matrix [n, in_dim_p] (OneHot.forward, deepxube/pytorch/pytorch_models.py:15-24)
// ones_col >= 0: columns ones_col .. ones_col + 2 are set to 1.0 (they multiply the folded bias columns of W0)
// act_depth > 0 (heurq_in): row r is node r / act_depth with action r % act_depth; column act_col0 + action is 1.0
// f16 != 0: the matrix is fp16 (split-fp16 mode) -- 1.0 = 0x3C00 instead of bf16's 0x3F80
__global__ void k_onehot_bf16(DomainDev d, int in_count, int depth, int in_dim_p, int ones_col, int act_depth, int act_col0,
                              const Ctl* __restrict__ ctl, const uint8_t* __restrict__ rows, const uint8_t* __restrict__ goals,
                              __nv_bfloat16* __restrict__ out, int f16) {
    if (ctl->error) return;
    const uint32_t one_lo = f16 ? 0x00003C00u : 0x00003F80u, one_hi = one_lo << 16;
    const uint32_t n = ctl->nn_n, start = ctl->nn_row_start;
    const int chunks = in_dim_p / 8;
    const size_t total = (size_t)n * chunks;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / chunks);
        const int c0 = (int)(t % chunks) * 8;
        const uint32_t node = act_depth > 0 ? r / (uint32_t)act_depth : r;
        const uint8_t* row = rows + (size_t)(start + node) * d.SP;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        if (act_depth > 0) {
            const int f = act_col0 + (int)(r % (uint32_t)act_depth) - c0;
            if (f >= 0 && f < 8) w[f >> 1] |= (f & 1) ? one_hi : one_lo;
        }
        const int i_lo = c0 / depth, i_hi = min(in_count - 1, (c0 + 7) / depth);
        for (int i = i_lo; i <= i_hi; ++i) {
            int v;
            if (i < d.S) v = row[i];
            else {
                const uint32_t inst = *reinterpret_cast<const uint32_t*>(row + d.SP - 4);
                v = goals[(size_t)inst * d.SP + (i - d.S)];
            }
            if (depth == 1) {                       // OneHot with depth 1 is x.float(): feature i = the value (exact in bf16 up to 256)
                const int f = i - c0;
                const uint32_t bits = f16 ? (uint32_t)__half_as_ushort(__float2half_rn((float)v))
                                          : (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn((float)v));
                w[f >> 1] |= (f & 1) ? (bits << 16) : bits;
                continue;
            }
            v = min(v, depth - 1);
            const int f = i * depth + v - c0;
            if (f >= 0 && f < 8) w[f >> 1] |= (f & 1) ? one_hi : one_lo;   // bf16 1.0 = 0x3F80
        }
        if (ones_col >= 0) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int f = ones_col + j - c0;
                if (f >= 0 && f < 8) w[f >> 1] |= (f & 1) ? one_hi : one_lo;
            }
        }
        reinterpret_cast<uint4*>(out + (size_t)r * in_dim_p)[c0 / 8] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Final Linear(H, out_dim) + clipping on bf16 activations; one warp per row (see k_out_layer in k_mlp_f32.cu).
__global__ void __launch_bounds__(256)
k_out_layer_bf16(NetDev net, const Ctl* __restrict__ ctl, const __nv_bfloat16* __restrict__ X, long long pitch, int kmax,
                 float* __restrict__ out_h, float* __restrict__ out_q, int absolute) {      // kmax: columns of X the last GEMM wrote
    if (ctl->error) return;
    const uint32_t n = ctl->nn_n, start = ctl->nn_row_start;
    const int lane = threadIdx.x & 31;
    const int Hp = net.Hp, od = net.out_dim;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&ctl->heur_evals, (unsigned long long)n);
    for (uint32_t r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
        const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(X + (size_t)r * (size_t)pitch);
        const size_t oi = absolute ? (size_t)start * (net.act_depth > 0 ? net.act_depth : 1) + r : (size_t)r;
        float hmin = __int_as_float(0x7f800000);
        for (int o0 = 0; o0 < od; o0 += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k2 = lane; k2 < kmax / 2; k2 += 32) {
                const float2 xv = __bfloat1622float2(x2[k2]);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (o0 + j < od) {
                        const float2 wv = *reinterpret_cast<const float2*>(net.wout + (size_t)(o0 + j) * Hp + 2 * k2);
                        acc[j] = fmaf(xv.x, wv.x, acc[j]);
                        acc[j] = fmaf(xv.y, wv.y, acc[j]);
                    }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], s);
                if (o0 + j < od) {
                    const float v = fmaxf(acc[j] + net.bout[o0 + j], 0.f);
                    hmin = fminf(hmin, v);
                    if (out_q && lane == 0) out_q[oi * od + o0 + j] = v;
                }
            }
        }
        if (out_h && lane == 0) out_h[oi] = hmin;
    }
}

// Scalar-output networks: the last hidden GEMM wrote P = 2 * Hp / BN fp32 partial dot products per row
// (partial[p * pstride + row]); out = max(sum_p partial + bias, 0), partials added in index order.
__global__ void __launch_bounds__(256)
k_out_final_bf16(NetDev net, const Ctl* __restrict__ ctl, const float* __restrict__ partial, long long pstride, int P,
                 float* __restrict__ out_h, float* __restrict__ out_q, int absolute) {
    if (ctl->error) return;
    const uint32_t n = ctl->nn_n, start = ctl->nn_row_start;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&ctl->heur_evals, (unsigned long long)n);
    const float b = net.bout[0];
    for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int p = 0; p < P; ++p) acc += partial[(size_t)p * (size_t)pstride + r];
        const float v = fmaxf(acc + b, 0.f);
        const size_t oi = absolute ? (size_t)start * (net.act_depth > 0 ? net.act_depth : 1) + r : (size_t)r;
        if (out_q) out_q[oi] = v;
        if (out_h) out_h[oi] = v;
    }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiledFn enc, CUtensorMap* map, void* base, uint64_t cols, uint64_t rows, uint32_t box_rows,
                    uint32_t box_cols = TC_BK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B, uint64_t pitch = 0) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {(pitch ? pitch : cols) * 2};      // pitch (elements): the buffer is a column slice of a wider array
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        dx_set_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
        return DX_ERR_CUDA;
    }
    return DX_OK;
}

// Whether the input layer of this network / domain runs with the one-hot encoding generated inside the GEMM (then no one-hot
// matrix is needed at all).  Networks that also read the goal keep the separate one-hot kernel (the goal row is a dependent
// gather); the generated A tiles of a row group (in_dim_p / 64 tiles of 16 KB, at most 8) must fit next to a ring of at least
// 3 weight tiles, which needs CTA pairs (half-size weight tiles).
bool dx_mlp_bf16_gen_input(const NetDev& net, const DomainDev& d) {
    if (const char* c = getenv("DXB200_FUSE_IN")) if (c[0] == '0') return false;
    if (const char* c = getenv("DXB200_GEMM_CLUSTER")) if (c[0] == '1') return false;
    const int BN = (net.Hp % 256 == 0) ? 256 : 128;
    const int kb_in = net.in_dim_p / TC_BK;
    const int b_bytes = (BN / 2) * TC_BK * 2, a_bytes = TC_BM * TC_BK * 2;
    const int ring_in = (5 * (a_bytes + b_bytes) - kb_in * a_bytes) / b_bytes;
    if (net.depth < 2) return false;          // depth-1 inputs are raw values, not one-hot codes: separate input kernel
    if (net.act_depth > 0) return false;      // (node, action) rows: separate input kernel
    return kb_in <= 8 && ring_in >= 3 && net.in_count <= d.S && d.SP % 16 == 0 &&
           4096 + 2 * TC_BM * d.SP + 512 <= TC_EPI_WARPS * TC_NBUF * 32 * 64;
}

// Shapes the fp32-accurate tensor-core path covers: at least one residual block (the scalar output layer is folded into the
// last hidden GEMM) and, for multi-output networks, out_dim <= 32 (one 128-wide n-tile).  DXB200_FP32_TC=0 keeps NET_FP32 on
// the CUDA-core SGEMM path (A/B runs; also the fallback for every other shape).
bool dx_mlp_split_supported(const NetDev& net) {
    if (const char* c = getenv("DXB200_FP32_TC")) if (c[0] == '0') return false;
    if (net.num_blocks < 1 || 2 * net.num_blocks > 128 || net.in_dim_p > 4096) return false;
    return net.out_dim == 1 || net.out_dim <= 32;
}

static float split_scale(const float* w, size_t n) {
    float m = 0.f;
    for (size_t i = 0; i < n; ++i) m = fmaxf(m, fabsf(w[i]));
    if (!(m > 0.f) || !std::isfinite(m)) return 1.f;
    int e = 0;
    frexpf(m, &e);                          // m = f * 2^e, f in [0.5, 1)  ->  m * 2^(14 - e) in [2^13, 2^14)
    return ldexpf(1.f, 14 - e);
}

// rows x K fp32 (row stride ld) -> rows x [hi(K) | lo(K)] fp16 of w * scale
static void split_rows(const float* w, size_t rows, size_t K, size_t ld, float scale, uint16_t* out) {
    for (size_t r = 0; r < rows; ++r)
        for (size_t k = 0; k < K; ++k) {
            const float x = w[r * ld + k] * scale;
            const __half hi = __float2half_rn(x);
            const __half lo = __float2half_rn(x - __half2float(hi));
            out[r * 2 * K + k] = __half_as_ushort(hi);
            out[r * 2 * K + K + k] = __half_as_ushort(lo);
        }
}

// (Re)builds the split-fp16 parameter copies from the host fp32 parameters (BN folded, padded).  The first call allocates;
// later calls (dx_load_weights again) overwrite in place.
int dx_mlp_split_upload(cudaStream_t s, NetDev& net, const float* w0p, const float* wres, const float* wout) {
    TcSplitW* sw = (TcSplitW*)net.tc_split;
    const size_t Hp = net.Hp, Kin = net.in_dim_p, L = (size_t)2 * net.num_blocks;
    const bool multi = net.out_dim > 1;
    if (!sw) {
        sw = new TcSplitW();
        memset(sw, 0, sizeof(*sw));
        DX_CUDA(cudaMalloc(&sw->w0, Hp * 2 * Kin * 2));
        DX_CUDA(cudaMalloc(&sw->wres, L * Hp * 2 * Hp * 2));
        if (multi) DX_CUDA(cudaMalloc(&sw->wout, (size_t)128 * 2 * Hp * 2));
        net.tc_split = sw;
    }
    std::vector<uint16_t> buf(std::max(std::max(Hp * 2 * Kin, L * Hp * 2 * Hp), (size_t)128 * 2 * Hp));
    sw->s_in = split_scale(w0p, Hp * Kin);
    split_rows(w0p, Hp, Kin, Kin, sw->s_in, buf.data());
    DX_CUDA(cudaMemcpyAsync(sw->w0, buf.data(), Hp * 2 * Kin * 2, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaStreamSynchronize(s));
    for (size_t l = 0; l < L; ++l) {
        sw->s_res[l] = split_scale(wres + l * Hp * Hp, Hp * Hp);
        split_rows(wres + l * Hp * Hp, Hp, Hp, Hp, sw->s_res[l], buf.data() + l * Hp * 2 * Hp);
    }
    DX_CUDA(cudaMemcpyAsync(sw->wres, buf.data(), L * Hp * 2 * Hp * 2, cudaMemcpyHostToDevice, s));
    DX_CUDA(cudaStreamSynchronize(s));
    if (multi) {
        std::fill(buf.begin(), buf.begin() + (size_t)128 * 2 * Hp, (uint16_t)0);
        sw->s_out = split_scale(wout, (size_t)net.out_dim * Hp);
        split_rows(wout, (size_t)net.out_dim, Hp, Hp, sw->s_out, buf.data());
        DX_CUDA(cudaMemcpyAsync(sw->wout, buf.data(), (size_t)128 * 2 * Hp * 2, cudaMemcpyHostToDevice, s));
        DX_CUDA(cudaStreamSynchronize(s));
    }
    return DX_OK;
}

// Called once after the weights and activation buffers exist.  Returns a heap TcPlan in *plan_out.
// onehot_buf may be nullptr when dx_mlp_bf16_gen_input() holds (gen_input then tells the launcher).
int dx_mlp_bf16_prepare(const NetDev& net, const MlpBufs& mb, void* onehot_buf, bool gen_input, void** plan_out) {
    if (!gen_input && !onehot_buf) { dx_set_error("NET_BF16: one-hot buffer missing"); return DX_ERR_ARG; }
    if (net.in_dim_p > 4096) {
        dx_set_error("NET_BF16: one-hot input wider than 4096 features is not supported (use NET_FP32)");
        return DX_ERR_UNSUPPORTED;
    }
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        dx_set_error("cuTensorMapEncodeTiled not available from the driver");
        return DX_ERR_CUDA;
    }
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    TcPlan* p = new TcPlan();
    p->max_rows = mb.cap;
    p->BN = (net.Hp % 256 == 0) ? 256 : 128;      // n-tiles must divide the padded width (Hp is a multiple of 128)
    p->CL = 2;
    if (const char* c = getenv("DXB200_GEMM_CLUSTER")) p->CL = (c[0] == '1') ? 1 : (c[0] == '4' ? 4 : 2);
    if (p->CL == 4 && (p->BN != 256 || net.Hp % 512 != 0)) p->CL = 2;      // pairs of n-tiles must tile the width
    p->CL_in = p->CL == 4 ? 2 : p->CL;
    p->grid4 = 0;
    p->fuse_out = 1;
    if (const char* c = getenv("DXB200_FUSE_OUT")) p->fuse_out = (c[0] == '0') ? 0 : 1;
    p->fuse_in = gen_input ? 1 : 0;
    p->onehot = (__nv_bfloat16*)onehot_buf;
    const TcSplitW* sw = (const TcSplitW*)net.tc_split;
    p->split = sw ? 1 : 0;
    if (sw) { p->CL = 2; p->fuse_out = 1; p->fuse_in = 0; }
    const uint64_t xp = sw ? 2 : 1;               // pitch multiplier: split buffers hold [hi | lo]
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev);
    // (with a generated input the A map of the first launch is never read: any valid map will do)
    int rc = onehot_buf ? make_map(enc, &p->a_onehot, onehot_buf, net.in_dim_p, (uint64_t)mb.cap, TC_BM)
                        : make_map(enc, &p->a_onehot, mb.xb[0], net.Hp, (uint64_t)mb.cap, TC_BM);
    const uint64_t pitch = (uint64_t)mb.pitch;
    for (int i = 0; i < 3 && rc == DX_OK; ++i)
        rc = make_map(enc, &p->a_x[i], mb.xb[i], xp * net.Hp, (uint64_t)mb.cap, TC_BM, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B, pitch);
    for (int i = 0; i < 3 && rc == DX_OK; ++i)
        rc = make_map(enc, &p->a_xh[i], mb.xb[i], xp * net.Hp, (uint64_t)mb.cap, TC_BM / 2, TC_BK, CU_TENSOR_MAP_SWIZZLE_128B, pitch);
    for (int i = 0; i < 3 && rc == DX_OK; ++i)
        rc = make_map(enc, &p->c_x[i], mb.xb[i], xp * net.Hp, (uint64_t)mb.cap, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, pitch);
    const int pair = p->CL >= 2 ? 2 : 1;
    p->fuse_l1 = (!sw && net.tc_l1 && p->fuse_in && p->CL_in == 2) ? 1 : 0;
    if (net.tc_l1 && !p->fuse_l1) {
        dx_set_error("NET_BF16: the stacked first-block weights need the generated-input launch (internal error)");
        delete p;
        return DX_ERR_STATE;
    }
    if (rc == DX_OK && p->fuse_l1)
        rc = make_map(enc, &p->c_x[3], mb.xb[0], 2 * (uint64_t)net.Hp, (uint64_t)mb.cap, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B, pitch);
    if (rc == DX_OK)
        rc = make_map(enc, &p->b_w0, sw ? sw->w0 : net.w0_bf16, (sw || net.tc_l1 == 2 ? 2 : 1) * (uint64_t)net.in_dim_p,
                      (uint64_t)net.Hp * (p->fuse_l1 ? 2 : 1), p->BN / pair);
    if (rc == DX_OK && net.num_blocks > 0)
        rc = make_map(enc, &p->b_res, sw ? sw->wres : net.wres_bf16, xp * net.Hp, (uint64_t)2 * net.num_blocks * net.Hp, p->BN / pair);
    p->q_out = 0;
    void* wout_tc = sw ? sw->wout : net.wout_bf16;
    if (rc == DX_OK && wout_tc) {
        const char* c = getenv("DXB200_QOUT");
        p->q_out = (sw || !(c && c[0] == '0')) ? 1 : 0;
        rc = make_map(enc, &p->b_out, wout_tc, xp * net.Hp, 128, 128 / pair);
    }
    if (rc == DX_OK) {
        cudaError_t ce = cudaSuccess;
        auto set_attr = [&](const void* f, int bytes) {
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        };
        set_attr((const void*)k_gemm_bf16_tc<256, 1, 0, false>, TcSmem<256, 1>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<256, 2, 0, false>, TcSmem<256, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 1, 0, false>, TcSmem<128, 1>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 2, 0, false>, TcSmem<128, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<256, 1, 1, false>, TcSmem<256, 1>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<256, 2, 1, false>, TcSmem<256, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 1, 1, false>, TcSmem<128, 1>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 2, 1, false>, TcSmem<128, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<256, 1, 0, true>, TcSmem<256, 1>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<256, 2, 0, true>, TcSmem<256, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 1, 0, true>, TcSmem<128, 1>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 2, 0, true>, TcSmem<128, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 1, 2, false>, TcSmem<128, 1>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 2, 2, false>, TcSmem<128, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<256, 4, 0, false>, TcSmem<256, 4>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<256, 4, 1, false>, TcSmem<256, 4>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<256, 2, 0, false, true>, TcSmem<256, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<256, 2, 1, false, true>, TcSmem<256, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 2, 0, false, true>, TcSmem<128, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 2, 1, false, true>, TcSmem<128, 2>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 2, 2, false, true>, TcSmem<128, 2>::TOTAL);
        if (ce == cudaSuccess && p->CL == 4) {
            // clusters of 4 do not fill every GPC: launch exactly as many as can be co-resident (persistent tiles)
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(p->num_sms / 4 * 4));
            cfg.blockDim = dim3(TC_THREADS);
            cfg.dynamicSmemBytes = TcSmem<256, 4>::TOTAL;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            int ncl = 0;
            if (cudaOccupancyMaxActiveClusters(&ncl, (const void*)k_gemm_bf16_tc<256, 4, 0, false>, &cfg) != cudaSuccess || ncl < 1) {
                cudaGetLastError();
                p->CL = 2;
            } else {
                p->grid4 = ncl * 4;
            }
        }
        if (ce != cudaSuccess) {
            dx_set_error(std::string("cudaFuncSetAttribute(max dynamic smem) failed: ") + cudaGetErrorString(ce));
            rc = DX_ERR_CUDA;
        }
    }
    if (rc != DX_OK) { delete p; return rc; }
    *plan_out = p;
    return DX_OK;
}

struct TcQOut { float* out_h; int qdim; int absolute; };       // Q_OUT launches: where the fp32 q-values / minima go

template <int BN, int CL, int MODE, bool GEN, bool SPLIT = false>
static void launch_gemm_t(cudaLaunchConfig_t* cfg, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& tc,
                          const CUtensorMap& tr, const Ctl* ctl, int w_row0, int K, int N, const float* bias, int has_res,
                          int relu, const float* wout, float* partial, long long pstride, const TcGen& gen, const TcQOut& qo,
                          const TcSplit& sp) {
    cfg->dynamicSmemBytes = TcSmem<BN, CL>::TOTAL;
    cfg->blockDim = dim3(GEN ? TC_THREADS_GEN : TC_THREADS);
    cudaLaunchKernelEx(cfg, k_gemm_bf16_tc<BN, CL, MODE, GEN, SPLIT>, a, b, tc, tr, ctl, w_row0, K, N, bias, has_res, relu, wout, partial,
                       pstride, gen, qo.out_h, qo.qdim, qo.absolute, sp, gen.relu_col0);
}

// c: output buffer index, r: residual buffer index or -1; partial != nullptr selects the fused output-layer epilogue
// (no activation store; buffer c is not written); gen != nullptr selects the fused one-hot input (a is not read);
// qo != nullptr: this launch is the multi-output layer itself (128-wide n-tile, partial = q-value array);
// cl: CTAs per cluster of this launch
static void launch_gemm(cudaStream_t s, const TcPlan* p, int cl, const CUtensorMap& a, const CUtensorMap& b, const Ctl* ctl, int w_row0,
                        int K, int N, const float* bias, int r, int c, int relu, const float* wout = nullptr,
                        float* partial = nullptr, long long pstride = 0, const TcGen* gen = nullptr, const TcQOut* qo = nullptr,
                        const TcSplit* split = nullptr) {
    const CUtensorMap& tc = p->c_x[c];
    const CUtensorMap& tr = p->c_x[r >= 0 ? r : (c < 3 ? c : 0)];
    cudaLaunchConfig_t cfg = {};
    // persistent grid: one cluster per SM group, but never more clusters than the engine's capacity can give work to (a
    // `graph.100B` engine evaluates <= 1200 rows: 5 row groups, not 74 clusters' worth of barrier / TMEM set-up per launch)
    unsigned grid = (unsigned)(cl == 4 ? p->grid4 : p->num_sms / cl * cl);
    {
        const int pair = cl >= 2 ? 2 : 1, npairs = cl / pair, bn = qo ? 128 : p->BN;
        const long long groups = (p->max_rows + 128LL * pair - 1) / (128LL * pair);
        const long long items = gen ? groups : groups * std::max(1, N / bn / npairs);
        if (items * cl < (long long)grid) grid = (unsigned)(std::max(items, 1LL) * cl);
    }
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cl;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int has_res = r >= 0 ? 1 : 0;
    const TcGen g = gen ? *gen : TcGen{};
    const TcQOut q = qo ? *qo : TcQOut{nullptr, 0, 0};
    const bool big = p->BN == 256, fuse = partial != nullptr;
    TcSplit sp = {};
    sp.nph = 1;
    if (split) sp = *split;
#define DX_TC_ARGS &cfg, a, b, tc, tr, ctl, w_row0, K, N, bias, has_res, relu, wout, partial, pstride, g, q, sp
#define DX_TC_PICK(CLV, MODEV, GENV) do { if (big) launch_gemm_t<256, CLV, MODEV, GENV>(DX_TC_ARGS); \
                                          else launch_gemm_t<128, CLV, MODEV, GENV>(DX_TC_ARGS); } while (0)
    if (split && !gen) {               // fp32-accurate mode: CTA pairs, TMA-fed A, (hi, lo) epilogue
        if (qo) launch_gemm_t<128, 2, 2, false, true>(DX_TC_ARGS);
        else if (fuse) { if (big) launch_gemm_t<256, 2, 1, false, true>(DX_TC_ARGS); else launch_gemm_t<128, 2, 1, false, true>(DX_TC_ARGS); }
        else { if (big) launch_gemm_t<256, 2, 0, false, true>(DX_TC_ARGS); else launch_gemm_t<128, 2, 0, false, true>(DX_TC_ARGS); }
    } else if (qo) {
        if (cl == 2) launch_gemm_t<128, 2, 2, false>(DX_TC_ARGS); else launch_gemm_t<128, 1, 2, false>(DX_TC_ARGS);
    } else if (gen) {
        if (cl == 2) DX_TC_PICK(2, 0, true); else DX_TC_PICK(1, 0, true);
    } else if (cl == 4) {
        if (fuse) launch_gemm_t<256, 4, 1, false>(DX_TC_ARGS); else launch_gemm_t<256, 4, 0, false>(DX_TC_ARGS);
    } else if (cl == 2) {
        if (fuse) DX_TC_PICK(2, 1, false); else DX_TC_PICK(2, 0, false);
    } else {
        if (fuse) DX_TC_PICK(1, 1, false); else DX_TC_PICK(1, 0, false);
    }
#undef DX_TC_PICK
#undef DX_TC_ARGS
}

int dx_launch_mlp_bf16(cudaStream_t s, const DomainDev& d, const NetDev& net, const MlpBufs& mb, const Ctl* ctl,
                       const uint8_t* rows, const uint8_t* goals, float* out_h, float* out_q, int absolute, long long max_rows) {
    const TcPlan* p = (const TcPlan*)net.tc_plan;
    if (!p) { dx_set_error("NET_BF16 plan missing"); return 0; }
    // action-as-input Q networks (heurq_in): one row per (node, action); the scalar outputs are the q-values
    const bool act = net.act_depth > 0;
    const long long max_nodes = max_rows;
    if (act) max_rows *= net.act_depth;
    if (max_rows > mb.cap) max_rows = mb.cap;
    int launches = dx_launch_act_begin(s, net, const_cast<Ctl*>(ctl));
    float* const q_dst = out_q;
    float* const h_dst = out_h;
    if (act) { out_h = q_dst; out_q = nullptr; }
    const int oh_blocks = (int)min((long long)dx_div_up(max_rows * (net.in_dim_p / 8), 256), (long long)148 * 16);
    int ix = 0, iy = 1, iz = 2;
    const int Hp = net.Hp;
    const bool fold_h = (net.tc_fold & 1) != 0;
    const float* bias0 = (net.tc_fold & 2) ? nullptr : ((fold_h || net.tc_l1) ? net.b0_tc : net.b0);
    const int ones_col = (net.tc_fold & 2) ? net.in_dim : -1;
    const int kv_in = net.in_dim_p, kv_h = Hp, nv = Hp;      // padded shapes (trimming the padding measured no gain)
    if (p->split) {
        // ---- fp32-accurate mode: fp16 (hi, lo) operands, 2 (input layer: the one-hot A is exact) / 3 MMA passes per layer ----
        const TcSplitW* sw = (const TcSplitW*)net.tc_split;
        const float t = TC_ACT_SCALE;
        auto mk = [&](int nph, int a1, int w0c, float wscale, bool a_scaled) {
            TcSplit sp = {};
            sp.nph = nph;
            // small terms first: A_hi.W_lo, A_lo.W_hi, then A_hi.W_hi (input layer: A.W_lo, A.W_hi)
            sp.a_col[0] = 0;   sp.w_col[0] = w0c;
            sp.a_col[1] = a1;  sp.w_col[1] = 0;
            sp.a_col[2] = 0;   sp.w_col[2] = 0;
            sp.acc_scale = 1.0f / (wscale * (a_scaled ? t : 1.0f));
            sp.res_scale = 1.0f / t;
            sp.out_scale = t;
            sp.lo_col = Hp;
            sp.f16 = 1;
            return sp;
        };
        dx_prof_begin(DX_ST_GEMM_IN);
        k_onehot_bf16<<<oh_blocks, 256, 0, s>>>(d, net.in_count, net.depth, net.in_dim_p, -1, net.act_depth,
                                                net.in_dim - net.act_depth, ctl, rows, goals, p->onehot, 1);
        {
            const TcSplit sp = mk(2, 0, net.in_dim_p, sw->s_in, false);
            launch_gemm(s, p, 2, p->a_onehot, p->b_w0, ctl, 0, kv_in, nv, net.b0, -1, ix, 0, nullptr, nullptr, 0, nullptr, nullptr, &sp);
        }
        launches += 2;
        dx_prof_end();
        const bool fuse_s = net.out_dim == 1 && net.num_blocks > 0;
        dx_prof_begin(DX_ST_GEMM_RES);
        for (int b = 0; b < net.num_blocks; ++b) {
            const TcSplit s1 = mk(3, Hp, Hp, sw->s_res[2 * b], true), s2 = mk(3, Hp, Hp, sw->s_res[2 * b + 1], true);
            launch_gemm(s, p, 2, p->a_x[ix], p->b_res, ctl, (2 * b) * Hp, kv_h, nv, net.bres + (size_t)(2 * b) * Hp, -1, iy, 1,
                        nullptr, nullptr, 0, nullptr, nullptr, &s1);
            if (fuse_s && b == net.num_blocks - 1)
                launch_gemm(s, p, 2, p->a_x[iy], p->b_res, ctl, (2 * b + 1) * Hp, kv_h, nv, net.bres + (size_t)(2 * b + 1) * Hp, ix, iz, 1,
                            net.wout, mb.partial, mb.cap, nullptr, nullptr, &s2);
            else
                launch_gemm(s, p, 2, p->a_x[iy], p->b_res, ctl, (2 * b + 1) * Hp, kv_h, nv, net.bres + (size_t)(2 * b + 1) * Hp, ix, iz, 1,
                            nullptr, nullptr, 0, nullptr, nullptr, &s2);
            launches += 2;
            int ti = ix; ix = iz; iz = ti;
        }
        dx_prof_end();
        if (fuse_s) {
            const int blocks = (int)min((long long)dx_div_up(max_rows, 256), (long long)148 * 8);
            k_out_final_bf16<<<blocks, 256, 0, s>>>(net, ctl, mb.partial, mb.cap, 2 * (Hp / p->BN), out_h, out_q, absolute);
        } else {                          // multi-output layer (dx_mlp_split_supported guarantees out_dim <= 32 here)
            TcQOut qo{out_h, net.out_dim, absolute};
            const TcSplit so = mk(3, Hp, Hp, sw->s_out, true);
            launch_gemm(s, p, 2, p->a_x[ix], p->b_out, ctl, 0, Hp, 128, net.bout_pad, -1, iz, 1, nullptr, out_q, 0, nullptr, &qo, &so);
        }
        ++launches;
        launches += dx_launch_act_end(s, net, const_cast<Ctl*>(ctl), q_dst, h_dst, absolute, max_nodes);
        return launches;
    }
    dx_prof_begin(DX_ST_GEMM_IN);
    if (p->fuse_in && p->CL_in == 2) {
        // input layer with the one-hot encoding generated inside the GEMM (no one-hot matrix in HBM)
        TcGen g;
        g.rows = rows; g.goals = goals; g.S = d.S; g.SP = d.SP; g.in_count = net.in_count; g.depth = net.depth; g.ones_col = ones_col;
        g.relu_col0 = 0;
        if (p->fuse_l1) {      // [x0 | relu(W1 x0 + b1)] in one launch: 2 * Hp output columns into buffers 0 and 1 (NetDev::tc_l1)
            g.relu_col0 = Hp;
            // the stacked weights are bf16 (hi, lo) pairs ([hi | lo] per row): two weight passes over the resident one-hot tiles
            // make x0 and the first hidden pre-activation accurate to ~16 bits before their single bf16 rounding on store
            TcSplit wsp = {};
            if (net.tc_l1 == 2) {      // DXB200_L1_SPLIT=1: bf16 (hi, lo) weight pairs, two weight passes (A/B arm, +80 % launch time)
                wsp.nph = 2;
                wsp.w_col[0] = net.in_dim_p; wsp.w_col[1] = 0;
            } else {                   // default: fp16 operands (the one-hot A is exact in any format; 11-bit instead of 8-bit weights)
                wsp.nph = 1;
                wsp.f16 = 1;
            }
            launch_gemm(s, p, p->CL_in, p->a_onehot, p->b_w0, ctl, 0, kv_in, 2 * nv, bias0, -1, 3, 1, nullptr, nullptr, 0, &g, nullptr, &wsp);
        } else {
            launch_gemm(s, p, p->CL_in, p->a_onehot, p->b_w0, ctl, 0, kv_in, nv, bias0, -1, ix, 0, nullptr, nullptr, 0, &g);
        }
        ++launches;
    } else {
        k_onehot_bf16<<<oh_blocks, 256, 0, s>>>(d, net.in_count, net.depth, net.in_dim_p, ones_col, net.act_depth,
                                                net.in_dim - net.act_depth, ctl, rows, goals, p->onehot, 0);
        launch_gemm(s, p, p->CL_in, p->a_onehot, p->b_w0, ctl, 0, kv_in, nv, bias0, -1, ix, 0);
        launches += 2;
    }
    dx_prof_end();
    // one timed interval around the 2 * num_blocks back-to-back hidden-layer launches (an event between
    // every pair of launches costs ~3 % of the step; the average launch duration is interval / launches)
    // scalar output (heurv): the H -> 1 output layer is folded into the epilogue of the last hidden GEMM, which then
    // writes 2 * Hp / BN fp32 partials per row into the (otherwise unused) third activation buffer
    const bool fuse = net.out_dim == 1 && net.num_blocks > 0 && p->fuse_out;
    const int P = 2 * (Hp / p->BN);
    const CUtensorMap* ax = p->CL == 4 ? p->a_xh : p->a_x;
    dx_prof_begin(DX_ST_GEMM_RES);
    for (int b = 0; b < net.num_blocks; ++b) {
        if (!(b == 0 && p->fuse_l1)) {       // (fused first block: buffer iy already holds the first hidden activation)
            launch_gemm(s, p, p->CL, ax[ix], p->b_res, ctl, (2 * b) * Hp, kv_h, nv, fold_h ? nullptr : net.bres + (size_t)(2 * b) * Hp, -1, iy, 1);
            ++launches;
        }
        if (fuse && b == net.num_blocks - 1)
            launch_gemm(s, p, p->CL, ax[iy], p->b_res, ctl, (2 * b + 1) * Hp, kv_h, nv, fold_h ? nullptr : net.bres + (size_t)(2 * b + 1) * Hp, ix, iz, 1,
                        net.wout, mb.partial, mb.cap);
        else
            launch_gemm(s, p, p->CL, ax[iy], p->b_res, ctl, (2 * b + 1) * Hp, kv_h, nv, fold_h ? nullptr : net.bres + (size_t)(2 * b + 1) * Hp, ix, iz, 1);
        ++launches;
        int ti = ix; ix = iz; iz = ti;
    }
    dx_prof_end();
    if (fuse) {
        const int blocks = (int)min((long long)dx_div_up(max_rows, 256), (long long)148 * 8);
        k_out_final_bf16<<<blocks, 256, 0, s>>>(net, ctl, mb.partial, mb.cap, P, out_h, out_q, absolute);
    } else if (p->q_out && out_q) {
        // multi-output layer as one more GEMM (128-wide n-tile of zero-padded weight rows), fp32 q-values from the epilogue
        TcQOut qo{out_h, net.out_dim, absolute};
        launch_gemm(s, p, p->CL_in, p->a_x[ix], p->b_out, ctl, 0, Hp, 128, net.bout_pad, -1, iz, 1, nullptr, out_q, 0, nullptr, &qo);
    } else {
        const __nv_bfloat16* x = (const __nv_bfloat16*)mb.xb[ix];
        const int blocks = (int)min((long long)dx_div_up(max_rows, 8), (long long)148 * 8);
        k_out_layer_bf16<<<blocks, 256, 0, s>>>(net, ctl, x, mb.pitch, Hp, out_h, out_q, absolute);
    }
    ++launches;
    launches += dx_launch_act_end(s, net, const_cast<Ctl*>(ctl), q_dst, h_dst, absolute, max_nodes);
    return launches;
}

// debug: read and clear the TC_PROFILE counters (zeros unless built with -DTC_PROFILE)
int dx_tc_profile_read(unsigned long long* out16) {
    if (cudaMemcpyFromSymbol(out16, g_tc_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return DX_ERR_CUDA;
    unsigned long long z[16] = {0};
    if (cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z)) != cudaSuccess) return DX_ERR_CUDA;
    return DX_OK;
}
