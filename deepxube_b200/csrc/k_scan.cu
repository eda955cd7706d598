// Device-wide exclusive prefix sums with device-side element counts (no host round trip):
//   reduce (one block per 4096 elements) -> spine (one block scans the block sums) -> apply.
// Used for: kept-children compaction (node ids = push order, deepxube/pathfinding/graph_search.py:48-51)
// and the per-pass digit offsets of the radix sort.
#include "dx_internal.cuh"

#define SCAN_THREADS 1024
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp, uint32_t& block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < (int)(blockDim.x >> 5)) ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        s_warp[lane] = w;   // inclusive over warps
    }
    __syncthreads();
    block_total = s_warp[(blockDim.x >> 5) - 1];
    const uint32_t warp_off = warp ? s_warp[warp - 1] : 0;
    __syncthreads();
    return warp_off + x - v;
}

template <typename LoadT>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_reduce(const LoadT* __restrict__ in, const uint32_t* __restrict__ n_ptr, uint32_t n_fixed, uint32_t mask,
              uint32_t* __restrict__ block_sums, const uint32_t* __restrict__ skip) {
    if (skip && *skip) return;
    __shared__ uint32_t s_warp[32];
    const uint32_t n = n_ptr ? *n_ptr : n_fixed;
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) sum += (uint32_t)in[base + i] & mask;
    uint32_t total;
    block_exclusive_scan(sum, s_warp, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_spine(uint32_t* __restrict__ block_sums, int n_blocks, uint32_t* __restrict__ total_out,
             const uint32_t* __restrict__ skip) {
    if (skip && *skip) return;
    __shared__ uint32_t s_warp[32];
    uint32_t carry = 0;
    for (int base = 0; base < n_blocks; base += SCAN_THREADS) {
        const int i = base + threadIdx.x;
        const uint32_t v = (i < n_blocks) ? block_sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, s_warp, total);
        if (i < n_blocks) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename LoadT>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_apply(const LoadT* in, uint32_t* out, const uint32_t* __restrict__ n_ptr, uint32_t n_fixed,
             uint32_t mask, const uint32_t* __restrict__ block_sums, const uint32_t* __restrict__ skip) {
    if (skip && *skip) return;
    __shared__ uint32_t s_warp[32];
    const uint32_t n = n_ptr ? *n_ptr : n_fixed;
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? ((uint32_t)in[base + i] & mask) : 0;
        sum += v[i];
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(sum, s_warp, total) + block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    // one-past-the-end entry = grand total (lets consumers take differences at segment ends)
    if (base <= n && n < base + SCAN_ITEMS) out[n] = ex;
}

// prefix over (flags & 1); out has n_max + 1 entries; *total_out = number of set flags.
int dx_launch_scan_flags(cudaStream_t s, const uint8_t* flags, uint32_t* out_prefix, const uint32_t* n_ptr, uint32_t* total_out,
                         int n_max, const ScanTemp& t) {
    const int blocks = dx_div_up((long long)n_max + 1, SCAN_TILE);
    k_scan_reduce<uint8_t><<<blocks, SCAN_THREADS, 0, s>>>(flags, n_ptr, 0, 1u, t.block_sums, nullptr);
    k_scan_spine<<<1, SCAN_THREADS, 0, s>>>(t.block_sums, blocks, total_out, nullptr);
    k_scan_apply<uint8_t><<<blocks, SCAN_THREADS, 0, s>>>(flags, out_prefix, n_ptr, 0, 1u, t.block_sums, nullptr);
    return 3;
}

// in-place exclusive scan of n (host-known) uint32 counters.
int dx_launch_scan_u32(cudaStream_t s, uint32_t* data, int n, const uint32_t* skip_flag, const ScanTemp& t) {
    const int blocks = dx_div_up((long long)n + 1, SCAN_TILE);
    k_scan_reduce<uint32_t><<<blocks, SCAN_THREADS, 0, s>>>(data, nullptr, (uint32_t)n, 0xFFFFFFFFu, t.block_sums, skip_flag);
    k_scan_spine<<<1, SCAN_THREADS, 0, s>>>(t.block_sums, blocks, nullptr, skip_flag);
    k_scan_apply<uint32_t><<<blocks, SCAN_THREADS, 0, s>>>(data, data, nullptr, (uint32_t)n, 0xFFFFFFFFu, t.block_sums, skip_flag);
    return 3;
}
