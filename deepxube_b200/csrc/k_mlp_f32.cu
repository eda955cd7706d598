// resnet_fc cost-to-go network, fp32 parity path (CUDA-core GEMMs, fp32 accumulate in k-ascending order).
//
// Reference network (eval mode), deepxube/nnets/resnet_fc.py:19-52 + deepxube/pytorch/pytorch_models.py:
//   x0 = W0 . onehot(s) + b0                       (OneHot :15-24, first Linear: no activation / norm)
//   per block: y = relu(BN1(W1 x + b1)); z = BN2(W2 y + b2); x = relu(z + x)     (:149-160, :198-209)
//   out = Wo x + bo ; h = max(out[0], 0) (base/pathfind_fns.py:218-221) ; q = max(out, 0) (fns_core.py:78-84)
// BN (eval, eps 1e-5) is folded into W/b at load time (engine.cu).  The one-hot layer is never
// materialised: x0 is the sum of in_count rows of W0^T selected by the input integers.
#include "dx_internal.cuh"

#define FL_THREADS 256
#define FL_ROWS 8

__device__ __forceinline__ int nn_input_value(const DomainDev& d, const uint8_t* __restrict__ row,
                                              const uint8_t* __restrict__ goals, int i) {
    if (i < d.S) return row[i];
    const uint32_t inst = *reinterpret_cast<const uint32_t*>(row + d.SP - 4);
    return goals[(size_t)inst * d.SP + (i - d.S)];
}

__global__ void __launch_bounds__(FL_THREADS)
k_first_layer_f32(DomainDev d, NetDev net, const Ctl* __restrict__ ctl, const uint8_t* __restrict__ rows,
                  const uint8_t* __restrict__ goals, float* __restrict__ X) {
    __shared__ int s_feat[FL_ROWS][257];     // in_count <= 256 (npuzzle.15 has 225) + the action feature of heurq_in networks
    __shared__ float s_mul[FL_ROWS][257];    // 1 for one-hot inputs; the input value itself when one_hot_depth == 1
    if (ctl->error) return;
    const uint32_t n = ctl->nn_n, start = ctl->nn_row_start;
    const int Hp = net.Hp, ic0 = net.in_count, amul = net.act_depth > 0 ? net.act_depth : 1;
    const int ic = ic0 + (net.act_depth > 0 ? 1 : 0);          // inputs per row incl. the action of a heurq_in network
    for (uint32_t base = blockIdx.x * FL_ROWS; base < n; base += gridDim.x * FL_ROWS) {
        const int nr = min((uint32_t)FL_ROWS, n - base);
        __syncthreads();
        for (int t = threadIdx.x; t < nr * ic; t += FL_THREADS) {
            const int r = t / ic, i = t % ic;
            // action-as-input networks: row (base + r) is node (base + r) / A with action (base + r) % A
            const uint8_t* row = rows + (size_t)(start + (base + r) / amul) * d.SP;
            if (i == ic0) {
                s_feat[r][i] = (net.in_dim - amul + (int)((base + r) % amul)) * Hp;
                s_mul[r][i] = 1.f;
                continue;
            }
            int v = nn_input_value(d, row, goals, i);
            if (net.depth > 1) {
                v = min(v, net.depth - 1);                 // one_hot would raise on out-of-range input; clamp instead
                s_feat[r][i] = (i * net.depth + v) * Hp;
                s_mul[r][i] = 1.f;
            } else {                                       // OneHot with depth 1 is x.float() (pytorch_models.py:24)
                s_feat[r][i] = i * Hp;
                s_mul[r][i] = (float)v;
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < Hp; c += FL_THREADS) {
            const float b = net.b0[c];
            float acc[FL_ROWS];
#pragma unroll
            for (int r = 0; r < FL_ROWS; ++r) acc[r] = 0.f;
            for (int i = 0; i < ic; ++i) {
#pragma unroll
                for (int r = 0; r < FL_ROWS; ++r)
                    if (r < nr) acc[r] += s_mul[r][i] * net.w0t[s_feat[r][i] + c];     // (1.0f * w is exact: one-hot sums unchanged)
            }
#pragma unroll
            for (int r = 0; r < FL_ROWS; ++r)
                if (r < nr) X[(size_t)(base + r) * Hp + c] = acc[r] + b;
        }
    }
}

// C[M,N] = act(A[M,K] . W[N,K]^T + bias (+ R)), M = ctl->nn_n, N and K multiples of 128 / 16.
// BM x BN output tile per block of 256 threads (16 x 16 threads, each (BM/16) x (BN/16) outputs in groups of 4): 128 x 128 for
// full batches; 64 x 64 when the whole batch is a few thousand rows (small-batch searches: `graph.100B` evaluates ~1200 states
// per iteration, 20 big tiles would use 20 SMs and each runs 4x longer).  Every output element is one fmaf chain over k in
// ascending order from 0, whatever the tile: both variants give identical bits.
#define SG_BK 16
template <int BM, int BN>
__global__ void __launch_bounds__(256)
k_sgemm_bias_act(const float* __restrict__ A, const float* __restrict__ W, const float* __restrict__ bias,
                 const float* __restrict__ R, float* __restrict__ C, const Ctl* __restrict__ ctl, int N, int K, int relu) {
    constexpr int GM = BM / 64, GN = BN / 64;           // groups of 4 rows / columns per thread
    __shared__ __align__(16) float As[2][SG_BK][BM + 4];
    __shared__ __align__(16) float Bs[2][SG_BK][BN + 4];
    if (ctl->error) return;
    const int M = (int)ctl->nn_n;
    const int n0 = blockIdx.x * BN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int m0 = blockIdx.y * BM; m0 < M; m0 += gridDim.y * BM) {
        float acc[GM * 4][GN * 4];
#pragma unroll
        for (int i = 0; i < GM * 4; ++i)
#pragma unroll
            for (int j = 0; j < GN * 4; ++j) acc[i][j] = 0.f;
        float4 ra[GM], rb[GN];
        auto gload = [&](int k0) {
#pragma unroll
            for (int q = 0; q < GM; ++q) {
                const int idx = tid + q * 256, row = idx >> 2, kc = (idx & 3) * 4;
                ra[q] = (m0 + row < M) ? *reinterpret_cast<const float4*>(A + (size_t)(m0 + row) * K + k0 + kc)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int q = 0; q < GN; ++q) {
                const int idx = tid + q * 256, row = idx >> 2, kc = (idx & 3) * 4;
                rb[q] = *reinterpret_cast<const float4*>(W + (size_t)(n0 + row) * K + k0 + kc);
            }
        };
        auto sstore = [&](int buf) {
#pragma unroll
            for (int q = 0; q < GM; ++q) {
                const int idx = tid + q * 256, row = idx >> 2, kc = (idx & 3) * 4;
                As[buf][kc + 0][row] = ra[q].x; As[buf][kc + 1][row] = ra[q].y;
                As[buf][kc + 2][row] = ra[q].z; As[buf][kc + 3][row] = ra[q].w;
            }
#pragma unroll
            for (int q = 0; q < GN; ++q) {
                const int idx = tid + q * 256, row = idx >> 2, kc = (idx & 3) * 4;
                Bs[buf][kc + 0][row] = rb[q].x; Bs[buf][kc + 1][row] = rb[q].y;
                Bs[buf][kc + 2][row] = rb[q].z; Bs[buf][kc + 3][row] = rb[q].w;
            }
        };
        __syncthreads();
        gload(0);
        sstore(0);
        __syncthreads();
        const int nk = K / SG_BK;
        for (int kt = 0; kt < nk; ++kt) {
            const int buf = kt & 1;
            if (kt + 1 < nk) gload((kt + 1) * SG_BK);
#pragma unroll
            for (int k = 0; k < SG_BK; ++k) {
                float a[GM * 4], b[GN * 4];
#pragma unroll
                for (int g = 0; g < GM; ++g) {
                    const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][g * 64 + ty * 4]);
                    a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
                }
#pragma unroll
                for (int g = 0; g < GN; ++g) {
                    const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][g * 64 + tx * 4]);
                    b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
                }
#pragma unroll
                for (int i = 0; i < GM * 4; ++i)
#pragma unroll
                    for (int j = 0; j < GN * 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            if (kt + 1 < nk) sstore(buf ^ 1);
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < GM * 4; ++i) {
            const int row = m0 + (i / 4) * 64 + ty * 4 + (i & 3);
            if (row >= M) continue;
#pragma unroll
            for (int jh = 0; jh < GN; ++jh) {
                const int col = n0 + jh * 64 + tx * 4;
                const float4 bv = *reinterpret_cast<const float4*>(bias + col);
                float4 v = make_float4(acc[i][jh * 4 + 0] + bv.x, acc[i][jh * 4 + 1] + bv.y, acc[i][jh * 4 + 2] + bv.z,
                                       acc[i][jh * 4 + 3] + bv.w);
                if (R) {
                    const float4 r = *reinterpret_cast<const float4*>(R + (size_t)row * N + col);
                    v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
                }
                if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                *reinterpret_cast<float4*>(C + (size_t)row * N + col) = v;
            }
        }
    }
}

static void launch_sgemm(cudaStream_t s, bool small, long long max_rows, const float* A, const float* W, const float* bias,
                         const float* R, float* C, const Ctl* ctl, int N, int K, int relu) {
    if (small) {
        dim3 grid(N / 64, (unsigned)dx_div_up(max_rows, 64));
        k_sgemm_bias_act<64, 64><<<grid, 256, 0, s>>>(A, W, bias, R, C, ctl, N, K, relu);
    } else {
        dim3 grid(N / 128, (unsigned)min((long long)dx_div_up(max_rows, 128), (long long)(148 * 2 / max(1, N / 128) + 1)));
        k_sgemm_bias_act<128, 128><<<grid, 256, 0, s>>>(A, W, bias, R, C, ctl, N, K, relu);
    }
}

// Final Linear(H, out_dim) + clipping; one warp per row.  Shared by the fp32 and bf16 paths (X fp32).
__global__ void __launch_bounds__(256)
k_out_layer(NetDev net, const Ctl* __restrict__ ctl, const float* __restrict__ X, float* __restrict__ out_h,
            float* __restrict__ out_q, int absolute) {
    if (ctl->error) return;
    const uint32_t n = ctl->nn_n, start = ctl->nn_row_start;
    const int lane = threadIdx.x & 31;
    const int Hp = net.Hp, od = net.out_dim;
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd((unsigned long long*)&ctl->heur_evals, (unsigned long long)n);
    for (uint32_t r = blockIdx.x * 8 + (threadIdx.x >> 5); r < n; r += gridDim.x * 8) {
        const float* x = X + (size_t)r * Hp;
        const size_t oi = absolute ? (size_t)start * (net.act_depth > 0 ? net.act_depth : 1) + r : (size_t)r;   // (node, action) rows
        float hmin = __int_as_float(0x7f800000);
        for (int o0 = 0; o0 < od; o0 += 4) {
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
            for (int k = lane; k < Hp; k += 32) {
                const float xv = x[k];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (o0 + j < od) acc[j] = fmaf(xv, net.wout[(size_t)(o0 + j) * Hp + k], acc[j]);
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

int dx_launch_out_layer(cudaStream_t s, const NetDev& net, const Ctl* ctl, const float* X, float* out_h, float* out_q,
                        int absolute, long long max_rows) {
    int blocks = (int)min((long long)dx_div_up(max_rows, 8), (long long)148 * 8);
    k_out_layer<<<blocks, 256, 0, s>>>(net, ctl, X, out_h, out_q, absolute);
    return 1;
}

// action-as-input Q networks (heurq_in): the job's node count becomes a (node, action) row count for the duration of the
// evaluation (factor > 0: multiply, < 0: divide back)
__global__ void k_nn_scale_rows(Ctl* ctl, int factor) {
    if (factor > 0) ctl->nn_n *= (uint32_t)factor;
    else ctl->nn_n /= (uint32_t)(-factor);
}

// heuristic of a node = min over its per-action q-values (deepxube/base/pathfinding.py:722-744); ctl->nn_n holds ROWS here
__global__ void k_act_hmin(const Ctl* __restrict__ ctl, int A, const float* __restrict__ q, float* __restrict__ out_h, int absolute) {
    if (ctl->error) return;
    const uint32_t nodes = ctl->nn_n / (uint32_t)A, start = ctl->nn_row_start;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < nodes; j += gridDim.x * blockDim.x) {
        const size_t o = absolute ? (size_t)(start + j) : (size_t)j;
        float h = __int_as_float(0x7f800000);
        for (int a = 0; a < A; ++a) h = fminf(h, q[o * A + a]);
        out_h[o] = h;
    }
}

int dx_launch_act_begin(cudaStream_t s, const NetDev& net, Ctl* ctl) {
    if (net.act_depth <= 0) return 0;
    k_nn_scale_rows<<<1, 1, 0, s>>>(ctl, net.act_depth);
    return 1;
}

int dx_launch_act_end(cudaStream_t s, const NetDev& net, Ctl* ctl, const float* q, float* out_h, int absolute, long long max_nodes) {
    if (net.act_depth <= 0) return 0;
    int launches = 1;
    if (out_h) {
        k_act_hmin<<<(int)min((long long)dx_div_up(max_nodes, 256), (long long)148 * 4), 256, 0, s>>>(ctl, net.act_depth, q, out_h, absolute);
        ++launches;
    }
    k_nn_scale_rows<<<1, 1, 0, s>>>(ctl, -net.act_depth);
    return launches;
}

int dx_launch_mlp_f32(cudaStream_t s, const DomainDev& d, const NetDev& net, const MlpBufs& mb, const Ctl* ctl,
                      const uint8_t* rows, const uint8_t* goals, float* out_h, float* out_q, int absolute, long long max_rows) {
    const long long max_nodes = max_rows;
    const bool act = net.act_depth > 0;
    if (act) max_rows *= net.act_depth;
    if (max_rows > mb.cap) max_rows = mb.cap;
    int launches = dx_launch_act_begin(s, net, const_cast<Ctl*>(ctl));
    const int fl_blocks = (int)min((long long)dx_div_up(max_rows, FL_ROWS), (long long)148 * 8);
    k_first_layer_f32<<<fl_blocks, FL_THREADS, 0, s>>>(d, net, ctl, rows, goals, mb.x[0]);
    ++launches;
    const int Hp = net.Hp;
    // few big tiles: the 64 x 64 variant spreads a small batch over the SMs
    const bool small = dx_div_up(max_rows, 128) * (Hp / 128) < 148 && dx_div_up(max_rows, 64) * (Hp / 64) <= 148 * 8;
    float* x = mb.x[0];
    float* y = mb.x[1];
    float* z = mb.x[2];
    for (int b = 0; b < net.num_blocks; ++b) {
        const float* w1 = net.wres + (size_t)(2 * b) * Hp * Hp;
        const float* w2 = net.wres + (size_t)(2 * b + 1) * Hp * Hp;
        dx_prof_begin(DX_ST_GEMM_RES);
        launch_sgemm(s, small, max_rows, x, w1, net.bres + (size_t)(2 * b) * Hp, nullptr, y, ctl, Hp, Hp, 1);
        dx_prof_end();
        dx_prof_begin(DX_ST_GEMM_RES);
        launch_sgemm(s, small, max_rows, y, w2, net.bres + (size_t)(2 * b + 1) * Hp, x, z, ctl, Hp, Hp, 1);
        dx_prof_end();
        float* t = x; x = z; z = t;
        launches += 2;
    }
    // action-as-input networks: the scalar outputs ARE the q-values, written flat [(node, action)] into the q array
    launches += dx_launch_out_layer(s, net, ctl, x, act ? out_q : out_h, act ? nullptr : out_q, absolute, max_rows);
    launches += dx_launch_act_end(s, net, const_cast<Ctl*>(ctl), out_q, out_h, absolute, max_nodes);
    return launches;
}
