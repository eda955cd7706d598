// resnet_fc cost-to-go network on the 5th-generation tensor cores (sm_100a): bf16 operands, fp32
// accumulation in TMEM.  Same network as k_mlp_f32.cu (deepxube/nnets/resnet_fc.py:19-52, eval-mode BN
// folded at load time); only the arithmetic type of the H x H contractions changes.
//
// One GEMM kernel, launched once per layer:  C[M, N] = act(A[M, K] . W[N, K]^T + bias (+ R))
//   * A (activations, row-major = K-major) and W (nn.Linear weight [out, in] = K-major) are streamed by
//     TMA (cp.async.bulk.tensor, 128-byte swizzle) into a 4-stage shared-memory ring,
//   * one elected thread issues tcgen05.mma.cta_group::1.kind::f16, 128 x BN x 16 per instruction, into one
//     of two TMEM accumulator stages (2 x BN columns),
//   * four epilogue warps drain the other accumulator stage with tcgen05.ld (32 lanes x 32 columns per
//     instruction, software-pipelined one chunk ahead), add bias / residual, apply ReLU, and write bf16.
//     Each warp stages its 32 x 32 chunk in a private, 64-byte-swizzled shared-memory buffer and moves it
//     with TMA (bulk tensor store; the residual chunk arrives by TMA load two chunks ahead), so every
//     global access of the epilogue is a full-line transfer instead of 32 strided row accesses,
//   * persistent CTAs (one per SM) walk the (m, n) tiles; M is read from the device control block.
// The one-hot input layer is a K = in_dim (padded to 64) GEMM over a bf16 one-hot matrix written by
// k_onehot_bf16 (exact: the entries are 0 / 1).
#include <cuda.h>
#include <stdlib.h>
#include <cuda_bf16.h>

#include "dx_internal.cuh"

#define TC_BM 128
#define TC_BK 64
#ifndef TC_STAGES
#define TC_STAGES 4          // operand ring depth (48 KB per stage at BN = 256)
#endif
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
#define TC_THREADS 256
#define TC_UMMA_K 16

struct TcPlan {
    CUtensorMap a_onehot;
    CUtensorMap a_x[3];
    CUtensorMap b_w0;
    CUtensorMap b_res;
    CUtensorMap c_x[3];          // 32 x 32 boxes (64-byte swizzle) over the activation buffers: C stores / R loads
    __nv_bfloat16* onehot;
    int BN;
    int CL;                      // CTAs per cluster (B tile multicast)
    int num_sms;
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(bar), "r"(parity), "r"(0x989680u)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
                 "[%0], [%1, {%3, %4}], [%2], %5;"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
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

template <int BN>
struct TcSmem {
    static constexpr int A_BYTES = TC_BM * TC_BK * 2;
    static constexpr int B_BYTES = BN * TC_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_OFF = TC_STAGES * STAGE_BYTES;
    static constexpr int EPI_BUF = 32 * 64;              // one 32-row x 32-column bf16 chunk
    static constexpr int EPI_WARP_BYTES = 2 * TC_NBUF * EPI_BUF;   // per epilogue warp: TC_NBUF output + TC_NBUF residual buffers
    static constexpr int BAR_OFF = EPI_OFF + 4 * EPI_WARP_BYTES;
    static constexpr int TOTAL = BAR_OFF + 256 + 1024;   // barriers + tmem ptr, + slack for 1024-byte alignment
};

// CL = CTAs per cluster (1 or 2).  With CL = 2 the two CTAs of a cluster work on vertically adjacent tiles
// (same n, rows m and m+1): each loads its own A tile and HALF of the shared B (weight) tile, multicast into
// both CTAs' shared memory, which cuts the L2 -> SM operand traffic per CTA from 48 KB to 32 KB per k-block.
// A stage is reusable only when BOTH CTAs' MMAs have read it (empty barrier count 2, multicast commits).
template <int BN, int CL>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_bf16_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const Ctl* __restrict__ ctl,
               int w_row0, int K, int N, const float* __restrict__ bias, int has_res, int relu) {
    extern __shared__ uint8_t smem_raw[];
    using SM = TcSmem<BN>;
    if (ctl->error) return;
    const int M = (int)ctl->nn_n;
    if (M == 0) return;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + SM::BAR_OFF;
    // barriers: full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2]; then the TMEM base address
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (TC_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * TC_STAGES + 2 + a); };
    auto res_bar = [&](int q, int b) { return bar_base + 8u * (2 * TC_STAGES + 4 + TC_NBUF * q + b); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * TC_STAGES + 4 + 4 * TC_NBUF);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t cta_rank = 0;
    if (CL > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1u);
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), CL); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 4); }
        for (int q = 0; q < 4; ++q)
            for (int b = 0; b < TC_NBUF; ++b) mbar_init(res_bar(q, b), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(2 * BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CL > 1) cluster_sync_all();          // peer barriers are initialised before any multicast lands
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int m_tiles = (M + TC_BM - 1) / TC_BM, n_tiles = N / BN;
    const int total = ((m_tiles + CL - 1) / CL) * n_tiles;       // tile groups of CL vertically adjacent tiles
    const int kblocks = K / TC_BK;
    const int cl_id = blockIdx.x / CL, n_cl = gridDim.x / CL;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            uint32_t it = 0;
            for (int t = cl_id; t < total; t += n_cl) {
                const int m0 = ((t / n_tiles) * CL + (int)cta_rank) * TC_BM, n0 = (t % n_tiles) * BN;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    mbar_wait(empty_bar(s), ((it / TC_STAGES) & 1u) ^ 1u);
                    const uint32_t a_dst = smem_base + s * SM::STAGE_BYTES, b_dst = a_dst + SM::A_BYTES;
                    mbar_arrive_expect_tx(full_bar(s), SM::STAGE_BYTES);
                    tma_load_2d(a_dst, &tmA, full_bar(s), kb * TC_BK, m0);
                    if (CL == 1)
                        tma_load_2d(b_dst, &tmB, full_bar(s), kb * TC_BK, w_row0 + n0);
                    else
                        tma_load_2d_mc(b_dst + cta_rank * (SM::B_BYTES / CL), &tmB, full_bar(s), kb * TC_BK,
                                       w_row0 + n0 + (int)cta_rank * (BN / CL), MC_MASK);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            // instruction descriptor: D = F32, A = B = BF16, both K-major, N = BN, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
            uint32_t it = 0, lt = 0;
            for (int t = cl_id; t < total; t += n_cl, ++lt) {
                const int as = lt & 1;
                mbar_wait(tempty_bar(as), ((lt >> 1) & 1u) ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    mbar_wait(full_bar(s), (it / TC_STAGES) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_src = smem_base + s * SM::STAGE_BYTES, b_src = a_src + SM::A_BYTES;
                    const uint64_t adesc = make_sw128_desc(a_src), bdesc = make_sw128_desc(b_src);
#pragma unroll
                    for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
                        umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
                    if (CL == 1) umma_commit(empty_bar(s));        // smem slot reusable once these MMAs have read it
                    else umma_commit_mc(empty_bar(s), MC_MASK);    // ... in both CTAs of the cluster
                }
                umma_commit(tfull_bar(as));             // accumulator complete
            }
        }
    } else if (warp >= 4) {
        // ---------------- epilogue (warps 4..7 <-> TMEM lane quarters 0..3) ----------------
        const int q = warp - 4;
        const uint32_t epi_base = smem_base + SM::EPI_OFF + q * SM::EPI_WARP_BYTES;
        uint8_t* epi_ptr = smem_raw + (epi_base - smem_u32(smem_raw));
        // 64-byte swizzle (Swizzle<2,4,3>): 16-byte chunk j of row r lives at chunk j ^ ((r >> 1) & 3)
        const uint32_t sw = (uint32_t)((lane >> 1) & 3);
        constexpr int NCH = BN / 32;
        uint32_t rphase = 0;                     // bit b = parity the next wait on residual buffer b expects
        uint32_t lt = 0;
        for (int t = cl_id; t < total; t += n_cl, ++lt) {
            const int as = lt & 1;
            const int m0 = ((t / n_tiles) * CL + (int)cta_rank) * TC_BM, n0 = (t % n_tiles) * BN;
            const int row0 = m0 + q * 32;
            if (has_res && lane == 0) {          // the first TC_NBUF residual chunks of this tile
#pragma unroll
                for (int b = 0; b < TC_NBUF && b < NCH; ++b) {
                    mbar_arrive_expect_tx(res_bar(q, b), SM::EPI_BUF);
                    tma_load_2d(epi_base + (TC_NBUF + b) * SM::EPI_BUF, &tmR, res_bar(q, b), n0 + b * 32, row0);
                }
            }
            mbar_wait(tfull_bar(as), (lt >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BN);

            auto process = [&](const uint32_t* v, int c) {
                const int b = c & (TC_NBUF - 1);
                const int col = n0 + c * 32;
                if (has_res) {
                    mbar_wait(res_bar(q, b), (rphase >> b) & 1u);
                    rphase ^= 1u << b;
                }
                // the TMA store that last read output buffer b (TC_NBUF chunks ago) must be done with it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read " TC_NBUF_M1 ";" ::: "memory");
                __syncwarp();
                const float4* b4 = reinterpret_cast<const float4*>(bias + col);
                const uint8_t* rrow = epi_ptr + (TC_NBUF + b) * SM::EPI_BUF + lane * 64;
                uint8_t* crow = epi_ptr + b * SM::EPI_BUF + lane * 64;
#pragma unroll
                for (int g = 0; g < 4; ++g) {      // 8 columns = one 16-byte chunk per group
                    float f[8];
                    const float4 ba = __ldg(b4 + 2 * g), bb = __ldg(b4 + 2 * g + 1);
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
                    uint4 o;
                    __nv_bfloat162* ob = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) ob[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                    *reinterpret_cast<uint4*>(crow + pos) = o;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to TMA
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmC, epi_base + b * SM::EPI_BUF, col, row0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    if (has_res && c + TC_NBUF < NCH) {      // all lanes are past their reads of residual buffer b
                        mbar_arrive_expect_tx(res_bar(q, b), SM::EPI_BUF);
                        tma_load_2d(epi_base + (TC_NBUF + b) * SM::EPI_BUF, &tmR, res_bar(q, b), n0 + (c + TC_NBUF) * 32, row0);
                    }
                }
            };

            uint32_t va[32], vb[32];
            tmem_ld32(t_row, va);
#pragma unroll 1
            for (int c = 0; c < NCH; c += 2) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tmem_ld32(t_row + (uint32_t)((c + 1) * 32), vb);
                process(va, c);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c + 2 < NCH) tmem_ld32(t_row + (uint32_t)((c + 2) * 32), va);
                process(vb, c + 1);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(as));
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (CL > 1) cluster_sync_all();          // no CTA exits while its peer may still multicast into it / signal it
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN));
    }
}

// ---------------------------------------------------------------------------------------------
// bf16 one-hot input matrix [n, in_dim_p] (OneHot.forward, deepxube/pytorch/pytorch_models.py:15-24)
__global__ void k_onehot_bf16(DomainDev d, int in_count, int depth, int in_dim_p, const Ctl* __restrict__ ctl,
                              const uint8_t* __restrict__ rows, const uint8_t* __restrict__ goals,
                              __nv_bfloat16* __restrict__ out) {
    if (ctl->error) return;
    const uint32_t n = ctl->nn_n, start = ctl->nn_row_start;
    const int chunks = in_dim_p / 8;
    const size_t total = (size_t)n * chunks;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(t / chunks);
        const int c0 = (int)(t % chunks) * 8;
        const uint8_t* row = rows + (size_t)(start + r) * d.SP;
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        const int i_lo = c0 / depth, i_hi = min(in_count - 1, (c0 + 7) / depth);
        for (int i = i_lo; i <= i_hi; ++i) {
            int v;
            if (i < d.S) v = row[i];
            else {
                const uint32_t inst = *reinterpret_cast<const uint32_t*>(row + d.SP - 4);
                v = goals[(size_t)inst * d.SP + (i - d.S)];
            }
            v = min(v, depth - 1);
            const int f = i * depth + v - c0;
            if (f >= 0 && f < 8) w[f >> 1] |= (f & 1) ? 0x3F800000u : 0x00003F80u;   // bf16 1.0 = 0x3F80
        }
        reinterpret_cast<uint4*>(out + (size_t)r * in_dim_p)[c0 / 8] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// Final Linear(H, out_dim) + clipping on bf16 activations; one warp per row (see k_out_layer in k_mlp_f32.cu).
__global__ void __launch_bounds__(256)
k_out_layer_bf16(NetDev net, const Ctl* __restrict__ ctl, const __nv_bfloat16* __restrict__ X, float* __restrict__ out_h,
                 float* __restrict__ out_q, int absolute) {
    if (ctl->error) return;
    const uint32_t n = ctl->nn_n, start = ctl->nn_row_start;
    const int lane = threadIdx.x & 31;
    const int Hp = net.Hp, od = net.out_dim;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&ctl->heur_evals, (unsigned long long)n);
    for (uint32_t r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
        const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(X + (size_t)r * Hp);
        const size_t oi = absolute ? (size_t)(start + r) : (size_t)r;
        float hmin = __int_as_float(0x7f800000);
        for (int o0 = 0; o0 < od; o0 += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k2 = lane; k2 < Hp / 2; k2 += 32) {
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

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(EncodeTiledFn enc, CUtensorMap* map, void* base, uint64_t cols, uint64_t rows, uint32_t box_rows,
                    uint32_t box_cols = TC_BK, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
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

// Called once after the weights and activation buffers exist.  Returns a heap TcPlan in *plan_out.
int dx_mlp_bf16_prepare(const NetDev& net, const MlpBufs& mb, void* onehot_buf, void** plan_out) {
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
    p->BN = net.Hp >= 256 ? 256 : 128;
    p->CL = 2;
    if (const char* c = getenv("DXB200_GEMM_CLUSTER")) p->CL = (c[0] == '1') ? 1 : 2;
    p->onehot = (__nv_bfloat16*)onehot_buf;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev);
    int rc = make_map(enc, &p->a_onehot, onehot_buf, net.in_dim_p, (uint64_t)mb.cap, TC_BM);
    for (int i = 0; i < 3 && rc == DX_OK; ++i) rc = make_map(enc, &p->a_x[i], mb.xb[i], net.Hp, (uint64_t)mb.cap, TC_BM);
    for (int i = 0; i < 3 && rc == DX_OK; ++i)
        rc = make_map(enc, &p->c_x[i], mb.xb[i], net.Hp, (uint64_t)mb.cap, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc == DX_OK) rc = make_map(enc, &p->b_w0, net.w0_bf16, net.in_dim_p, net.Hp, p->BN / p->CL);
    if (rc == DX_OK && net.num_blocks > 0)
        rc = make_map(enc, &p->b_res, net.wres_bf16, net.Hp, (uint64_t)2 * net.num_blocks * net.Hp, p->BN / p->CL);
    if (rc == DX_OK) {
        cudaError_t ce = cudaSuccess;
        auto set_attr = [&](const void* f, int bytes) {
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        };
        set_attr((const void*)k_gemm_bf16_tc<256, 1>, TcSmem<256>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<256, 2>, TcSmem<256>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 1>, TcSmem<128>::TOTAL);
        set_attr((const void*)k_gemm_bf16_tc<128, 2>, TcSmem<128>::TOTAL);
        if (ce != cudaSuccess) {
            dx_set_error(std::string("cudaFuncSetAttribute(max dynamic smem) failed: ") + cudaGetErrorString(ce));
            rc = DX_ERR_CUDA;
        }
    }
    if (rc != DX_OK) { delete p; return rc; }
    *plan_out = p;
    return DX_OK;
}

// c: output buffer index, r: residual buffer index or -1
static void launch_gemm(cudaStream_t s, const TcPlan* p, const CUtensorMap& a, const CUtensorMap& b, const Ctl* ctl, int w_row0,
                        int K, int N, const float* bias, int r, int c, int relu) {
    const CUtensorMap& tc = p->c_x[c];
    const CUtensorMap& tr = p->c_x[r >= 0 ? r : c];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(p->num_sms / p->CL * p->CL));
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = p->BN == 256 ? TcSmem<256>::TOTAL : TcSmem<128>::TOTAL;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p->CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int has_res = r >= 0 ? 1 : 0;
    if (p->BN == 256 && p->CL == 2) cudaLaunchKernelEx(&cfg, k_gemm_bf16_tc<256, 2>, a, b, tc, tr, ctl, w_row0, K, N, bias, has_res, relu);
    else if (p->BN == 256) cudaLaunchKernelEx(&cfg, k_gemm_bf16_tc<256, 1>, a, b, tc, tr, ctl, w_row0, K, N, bias, has_res, relu);
    else if (p->CL == 2) cudaLaunchKernelEx(&cfg, k_gemm_bf16_tc<128, 2>, a, b, tc, tr, ctl, w_row0, K, N, bias, has_res, relu);
    else cudaLaunchKernelEx(&cfg, k_gemm_bf16_tc<128, 1>, a, b, tc, tr, ctl, w_row0, K, N, bias, has_res, relu);
}

int dx_launch_mlp_bf16(cudaStream_t s, const DomainDev& d, const NetDev& net, const MlpBufs& mb, const Ctl* ctl,
                       const uint8_t* rows, const uint8_t* goals, float* out_h, float* out_q, int absolute, long long max_rows) {
    const TcPlan* p = (const TcPlan*)net.tc_plan;
    if (!p) { dx_set_error("NET_BF16 plan missing"); return 0; }
    if (max_rows > mb.cap) max_rows = mb.cap;
    int launches = 0;
    const int oh_blocks = (int)min((long long)dx_div_up(max_rows * (net.in_dim_p / 8), 256), (long long)148 * 16);
    dx_prof_begin(DX_ST_GEMM_IN);
    k_onehot_bf16<<<oh_blocks, 256, 0, s>>>(d, net.in_count, net.depth, net.in_dim_p, ctl, rows, goals, p->onehot);
    ++launches;
    int ix = 0, iy = 1, iz = 2;
    const int Hp = net.Hp;
    launch_gemm(s, p, p->a_onehot, p->b_w0, ctl, 0, net.in_dim_p, Hp, net.b0, -1, ix, 0);
    ++launches;
    dx_prof_end();
    for (int b = 0; b < net.num_blocks; ++b) {
        dx_prof_begin(DX_ST_GEMM_RES);
        launch_gemm(s, p, p->a_x[ix], p->b_res, ctl, (2 * b) * Hp, Hp, Hp, net.bres + (size_t)(2 * b) * Hp, -1, iy, 1);
        dx_prof_end();
        dx_prof_begin(DX_ST_GEMM_RES);
        launch_gemm(s, p, p->a_x[iy], p->b_res, ctl, (2 * b + 1) * Hp, Hp, Hp, net.bres + (size_t)(2 * b + 1) * Hp, ix, iz, 1);
        dx_prof_end();
        launches += 2;
        int ti = ix; ix = iz; iz = ti;
    }
    const __nv_bfloat16* x = (const __nv_bfloat16*)mb.xb[ix];
    const int blocks = (int)min((long long)dx_div_up(max_rows, 8), (long long)148 * 8);
    k_out_layer_bf16<<<blocks, 256, 0, s>>>(net, ctl, x, out_h, out_q, absolute);
    ++launches;
    return launches;
}
