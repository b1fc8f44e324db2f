// K3, narrow products — exact fp32 FFMA kernels for contractions with one tiny dimension, where a 256 x 256
// tensor-core tile would be > 90 % padding.  In the PLA-GNN epoch these are the classifier head
// (code/model.py:17,28-29  liner2: 100 -> 12 classes, and its backward at code/train.py:204):
//
//   narrow_n_kernel     n <= 16            prob = sigmoid(h4 W2^T + b2)           [24041 x 12], k = 100
//   narrow_k_kernel     sum of k_p <= 32   dz4  = (dz5 W2) * leaky'(h4)           [24041 x 100], k = 12
//   narrow_wgrad_kernel m <= 16            dW2  = dz5^T h4, db2 = colsum(dz5)     [12 x 100],  k = 24041
//
// All three are HBM/L2-bound streaming kernels (the 24041 x 100 operand is read once, 9.6 MB); accumulation is plain
// fp32 FMA in k order (row-block partials of the weight gradient are folded in block order, in double), so results are
// deterministic and at FFMA accuracy.
#include "common.cuh"

namespace plagnn {

struct NarrowParams {
    int64_t m, n;
    int npairs;
    plagnn_gemm_pair pairs[PLAGNN_GEMM_MAX_PAIRS];
    const float* bias;
    int act;
    float slope;
    const float* gate;
    int64_t ldg;
    int gate_act;
    float* c;
    int64_t ldc;
};

constexpr int NW_THREADS = 256;
constexpr int NW_MAX = 16;      // the tiny dimension
constexpr int NW_ROWS = 64;     // narrow_n: output rows per block
constexpr int NW_KC = 128;      // narrow_n: k chunk staged in shared memory
constexpr int NK_MAX_K = 32;    // narrow_k: longest contraction
constexpr int NK_MAX_B = 8192;  // narrow_k: floats of B kept in shared memory
constexpr int NK_ROWS = 64;     // narrow_k: output rows per block step
constexpr int WG_ROWS = 128;    // narrow_wgrad: contraction rows per block
constexpr int WG_COLS = 128;    // narrow_wgrad: output columns per block

__device__ __forceinline__ float narrow_epilogue(float v, const NarrowParams& P, int64_t row, int64_t j) {
    if (P.bias) v += __ldg(P.bias + j);
    v = apply_act(v, P.act, P.slope);
    if (P.gate) v *= act_grad_from_output(__ldg(P.gate + row * P.ldg + j), P.gate_act, P.slope);
    return v;
}

// thread (r, g) of a 64-row block owns row r and columns g, g+4, ..., g+4(CG-1)
template <int CG>
__global__ void __launch_bounds__(NW_THREADS) narrow_n_kernel(const NarrowParams P) {
    pdl_enter();
    __shared__ float As[NW_ROWS][NW_KC + 1];
    __shared__ float Bs[NW_MAX][NW_KC + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = tid >> 2, g = tid & 3;
    const int64_t r0 = (int64_t)blockIdx.x * NW_ROWS;
    const int n = (int)P.n;
    float acc[CG];
#pragma unroll
    for (int i = 0; i < CG; ++i) acc[i] = 0.f;

    for (int p = 0; p < P.npairs; ++p) {
        const plagnn_gemm_pair q = P.pairs[p];
        for (int64_t k0 = 0; k0 < q.k; k0 += NW_KC) {
            const int kc = (int)(q.k - k0 < NW_KC ? q.k - k0 : NW_KC);
            // every thread issues 8 independent loads before the first shared-memory store (the tile load is pure latency)
            constexpr int NWARP = NW_THREADS / 32, RPW = NW_ROWS / NWARP;
            if (!q.a_trans) {
                for (int kk = lane; kk < kc; kk += 32) {
                    float t[RPW];
#pragma unroll
                    for (int i = 0; i < RPW; ++i) {
                        const int64_t row = r0 + warp + NWARP * i;
                        t[i] = row < P.m ? __ldg(q.a + row * q.lda + k0 + kk) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < RPW; ++i) As[warp + NWARP * i][kk] = t[i];
                }
            } else {
                for (int kb = warp; kb < kc; kb += NWARP * 4) {
                    float t[4][NW_ROWS / 32];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int i = 0; i < NW_ROWS / 32; ++i) {
                            const int kk = kb + NWARP * u;
                            const int64_t row = r0 + lane + 32 * i;
                            t[u][i] = (kk < kc && row < P.m) ? __ldg(q.a + (k0 + kk) * q.lda + row) : 0.f;
                        }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int i = 0; i < NW_ROWS / 32; ++i)
                            if (kb + NWARP * u < kc) As[lane + 32 * i][kb + NWARP * u] = t[u][i];
                }
            }
            if (!q.b_trans) {
                for (int kk = lane; kk < kc; kk += 32) {
                    float t[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const int j = warp + NWARP * i;
                        t[i] = j < n ? __ldg(q.b + (int64_t)j * q.ldb + k0 + kk) : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 2; ++i) Bs[warp + NWARP * i][kk] = t[i];
                }
            } else {
                for (int base = 0; base < kc * NW_MAX; base += NW_THREADS * 4) {
                    float t[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int idx = base + u * NW_THREADS + tid, kk = idx >> 4, j = idx & 15;
                        t[u] = (kk < kc && j < n) ? __ldg(q.b + (k0 + kk) * q.ldb + j) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int idx = base + u * NW_THREADS + tid, kk = idx >> 4, j = idx & 15;
                        if (kk < kc) Bs[j][kk] = t[u];
                    }
                }
            }
            __syncthreads();
#pragma unroll 4
            for (int kk = 0; kk < kc; ++kk) {
                const float a = As[r][kk];
#pragma unroll
                for (int i = 0; i < CG; ++i) acc[i] = fmaf(a, Bs[g + 4 * i][kk], acc[i]);
            }
            __syncthreads();
        }
    }
    const int64_t row = r0 + r;
    if (row >= P.m) return;
#pragma unroll
    for (int i = 0; i < CG; ++i) {
        const int j = g + 4 * i;
        if (j < n) P.c[row * P.ldc + j] = narrow_epilogue(acc[i], P, row, j);
    }
}

// one thread per (row, 4 output columns); the whole B operand ([sum k] x n, n contiguous) sits in shared memory
__global__ void __launch_bounds__(NW_THREADS) narrow_k_kernel(const NarrowParams P, int vec_io) {
    pdl_enter();
    extern __shared__ __align__(16) float Bsh[];
    const int n = (int)P.n, n4 = (n + 3) >> 2, npad = n4 * 4;
    int koff = 0;
    for (int p = 0; p < P.npairs; ++p) {
        const plagnn_gemm_pair q = P.pairs[p];
        const int k = (int)q.k;
        // row kk of the shared copy holds B[kk, 0..npad): a warp per kk, lanes over j, 4 loads in flight per thread
        for (int kk = threadIdx.x >> 5; kk < k; kk += NW_THREADS / 32) {
            for (int j0 = threadIdx.x & 31; j0 < npad; j0 += 128) {
                float t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + 32 * u;
                    t[u] = j < n ? (q.b_trans ? __ldg(q.b + (int64_t)kk * q.ldb + j) : __ldg(q.b + (int64_t)j * q.ldb + kk)) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j0 + 32 * u < npad) Bsh[(koff + kk) * npad + j0 + 32 * u] = t[u];
            }
        }
        koff += k;
    }
    __syncthreads();
    // blocks stride over groups of NK_ROWS rows; inside a group the (row, column quad) index is 32-bit (one cheap division)
    const unsigned per_group = (unsigned)(NK_ROWS * n4);
    for (int64_t r0 = (int64_t)blockIdx.x * NK_ROWS; r0 < P.m; r0 += (int64_t)gridDim.x * NK_ROWS)
    for (unsigned it = threadIdx.x; it < per_group; it += NW_THREADS) {
        const unsigned rr = it / (unsigned)n4;
        const int64_t row = r0 + rr;
        if (row >= P.m) break;
        const int j = (int)(it - rr * (unsigned)n4) * 4;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int ko = 0;
        for (int p = 0; p < P.npairs; ++p) {
            const plagnn_gemm_pair q = P.pairs[p];
            const int k = (int)q.k;
            const float* ap = q.a_trans ? q.a + row : q.a + row * q.lda;
            const int64_t as = q.a_trans ? q.lda : 1;
#pragma unroll 4
            for (int kk = 0; kk < k; ++kk) {
                const float a = __ldg(ap + kk * as);
                const float4 b = *reinterpret_cast<const float4*>(Bsh + (ko + kk) * npad + j);
                acc.x = fmaf(a, b.x, acc.x); acc.y = fmaf(a, b.y, acc.y);
                acc.z = fmaf(a, b.z, acc.z); acc.w = fmaf(a, b.w, acc.w);
            }
            ko += k;
        }
        if (vec_io && j + 3 < n) {
            float v[4] = {acc.x, acc.y, acc.z, acc.w};
            float gt[4] = {1.f, 1.f, 1.f, 1.f};
            if (P.gate) {
                const float4 g4 = ldg_f4(P.gate + row * P.ldg + j);
                gt[0] = g4.x; gt[1] = g4.y; gt[2] = g4.z; gt[3] = g4.w;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float t = v[i];
                if (P.bias) t += __ldg(P.bias + j + i);
                t = apply_act(t, P.act, P.slope);
                if (P.gate) t *= act_grad_from_output(gt[i], P.gate_act, P.slope);
                v[i] = t;
            }
            *reinterpret_cast<float4*>(P.c + row * P.ldc + j) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (j + i < n) P.c[row * P.ldc + j + i] = narrow_epilogue(v[i], P, row, j + i);
        }
    }
}

// dW[m x n] (+ db as column n) partials of one 128-row block of the contraction: thread (col, half) runs 64 rows
template <int MM>
__global__ void __launch_bounds__(NW_THREADS)
narrow_wgrad_kernel(const float* __restrict__ dz, int64_t lddz, const float* __restrict__ x, int64_t ldx, int64_t k, int m,
                    int n, float* __restrict__ part) {
    pdl_enter();
    __shared__ __align__(16) float dzs[WG_ROWS][NW_MAX];
    __shared__ float red[MM][WG_COLS];
    const int tid = threadIdx.x, cl = tid & (WG_COLS - 1), half = tid >> 7;
    const int64_t r0 = (int64_t)blockIdx.x * WG_ROWS;
    const int j = blockIdx.y * WG_COLS + cl;
    {
        float t[WG_ROWS * NW_MAX / NW_THREADS];
#pragma unroll
        for (int u = 0; u < WG_ROWS * NW_MAX / NW_THREADS; ++u) {
            const int idx = tid + u * NW_THREADS, rr = idx >> 4, i = idx & 15;
            t[u] = (i < m && r0 + rr < k) ? __ldg(dz + (r0 + rr) * lddz + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < WG_ROWS * NW_MAX / NW_THREADS; ++u) {
            const int idx = tid + u * NW_THREADS;
            dzs[idx >> 4][idx & 15] = t[u];
        }
    }
    __syncthreads();
    float acc[MM];
#pragma unroll
    for (int i = 0; i < MM; ++i) acc[i] = 0.f;
    const int rbeg = half * (WG_ROWS / 2);
    const float fill = j == n ? 1.f : 0.f;          // column n is the all-ones column: its product is the bias gradient
#pragma unroll 8
    for (int rr = rbeg; rr < rbeg + WG_ROWS / 2; ++rr) {
        const int64_t row = r0 + rr;
        const float xv = (j < n && row < k) ? __ldg(x + row * ldx + j) : fill;
#pragma unroll
        for (int i4 = 0; i4 < MM / 4; ++i4) {
            const float4 d = *reinterpret_cast<const float4*>(&dzs[rr][4 * i4]);
            acc[4 * i4 + 0] = fmaf(d.x, xv, acc[4 * i4 + 0]);
            acc[4 * i4 + 1] = fmaf(d.y, xv, acc[4 * i4 + 1]);
            acc[4 * i4 + 2] = fmaf(d.z, xv, acc[4 * i4 + 2]);
            acc[4 * i4 + 3] = fmaf(d.w, xv, acc[4 * i4 + 3]);
        }
    }
    if (half == 1) {
#pragma unroll
        for (int i = 0; i < MM; ++i) red[i][cl] = acc[i];
    }
    __syncthreads();
    if (half == 0 && j <= n) {
        float* out = part + (int64_t)blockIdx.x * m * (n + 1);
#pragma unroll
        for (int i = 0; i < MM; ++i)
            if (i < m) out[(int64_t)i * (n + 1) + j] = acc[i] + red[i][cl];
    }
}

// one warp per output element: row-block partials summed in double, lanes strided over the blocks, fixed shuffle tree
__global__ void __launch_bounds__(NW_THREADS)
narrow_wgrad_reduce_kernel(const float* __restrict__ part, int nblocks, int m, int n, float* __restrict__ dw, int64_t lddw,
                           float* __restrict__ db) {
    pdl_enter();
    const int lane = threadIdx.x & 31;
    const int e = blockIdx.x * (NW_THREADS / 32) + (threadIdx.x >> 5);
    if (e >= m * (n + 1)) return;
    const int64_t stride = (int64_t)m * (n + 1);
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s += (double)__ldg(part + b * stride + e);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) {
        const int i = e / (n + 1), j = e - i * (n + 1);
        if (j == n) db[i] = (float)s;
        else dw[(int64_t)i * lddw + j] = (float)s;
    }
}

// ---- host ------------------------------------------------------------------------------------
static inline bool narrow_n_shape(int64_t n, int64_t ktot) { return n <= NW_MAX && ktot <= 4096; }
static inline bool narrow_k_shape(int64_t n, int64_t ktot) { return ktot <= NK_MAX_K && ktot * ((n + 3) / 4 * 4) <= NK_MAX_B; }

bool gemm_narrow_eligible(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs) {
    int64_t ktot = 0;
    for (int p = 0; p < npairs; ++p) ktot += pairs[p].k;
    return m >= 1 && (narrow_n_shape(n, ktot) || narrow_k_shape(n, ktot));
}

int gemm_narrow_launch(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs, const float* bias, int act,
                       float slope, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc, cudaStream_t st) {
    NarrowParams P;
    P.m = m; P.n = n; P.npairs = npairs;
    int64_t ktot = 0;
    for (int p = 0; p < npairs; ++p) { P.pairs[p] = pairs[p]; ktot += pairs[p].k; }
    P.bias = bias; P.act = act; P.slope = slope; P.gate = gate; P.ldg = ldg; P.gate_act = gate_act; P.c = c; P.ldc = ldc;
    if (narrow_n_shape(n, ktot)) {
        const unsigned grid = (unsigned)ceil_div(m, NW_ROWS);
        switch ((n + 3) / 4) {
            case 1: launch_pdl(narrow_n_kernel<1>, dim3(grid), dim3(NW_THREADS), 0, st, P); break;
            case 2: launch_pdl(narrow_n_kernel<2>, dim3(grid), dim3(NW_THREADS), 0, st, P); break;
            case 3: launch_pdl(narrow_n_kernel<3>, dim3(grid), dim3(NW_THREADS), 0, st, P); break;
            default: launch_pdl(narrow_n_kernel<4>, dim3(grid), dim3(NW_THREADS), 0, st, P); break;
        }
        return check_launch("gemm(narrow n)");
    }
    if (!narrow_k_shape(n, ktot)) return fail(PLAGNN_ERR_UNSUPPORTED, "gemm", "narrow backend needs n <= 16 or a contraction <= 32");
    const int64_t n4 = (n + 3) / 4;
    const int vec_io = (ldc & 3) == 0 && aligned16(c) && (!gate || ((ldg & 3) == 0 && aligned16(gate)));
    const int64_t blocks = ceil_div(m, NK_ROWS);
    const int64_t cap = (int64_t)sm_count() * 8;
    const size_t smem = (size_t)ktot * n4 * 4 * sizeof(float);
    launch_pdl(narrow_k_kernel, dim3((unsigned)(blocks < cap ? blocks : cap)), dim3(NW_THREADS), smem, st, P, vec_io);
    return check_launch("gemm(narrow k)");
}

size_t gemm_narrow_wgrad_bytes(int64_t m, int64_t n, int64_t k) {
    if (m > NW_MAX) return 0;
    return align_up((size_t)ceil_div(k, WG_ROWS) * (size_t)m * (size_t)(n + 1) * sizeof(float), 256);
}

int gemm_narrow_wgrad_launch(int64_t m, int64_t n, const float* dz, int64_t lddz, const float* x, int64_t ldx, int64_t k,
                             float* dw, int64_t lddw, float* db, void* workspace, cudaStream_t st) {
    const int nblocks = (int)ceil_div(k, WG_ROWS);
    float* part = (float*)workspace;
    dim3 grid((unsigned)nblocks, (unsigned)ceil_div(n + 1, WG_COLS));
    switch ((m + 3) / 4) {
        case 1: launch_pdl(narrow_wgrad_kernel<4>, grid, dim3(NW_THREADS), 0, st, dz, lddz, x, ldx, k, (int)m, (int)n, part); break;
        case 2: launch_pdl(narrow_wgrad_kernel<8>, grid, dim3(NW_THREADS), 0, st, dz, lddz, x, ldx, k, (int)m, (int)n, part); break;
        case 3: launch_pdl(narrow_wgrad_kernel<12>, grid, dim3(NW_THREADS), 0, st, dz, lddz, x, ldx, k, (int)m, (int)n, part); break;
        default: launch_pdl(narrow_wgrad_kernel<16>, grid, dim3(NW_THREADS), 0, st, dz, lddz, x, ldx, k, (int)m, (int)n, part); break;
    }
    const int64_t elems = m * (n + 1);
    launch_pdl(narrow_wgrad_reduce_kernel, dim3((unsigned)ceil_div(elems, NW_THREADS / 32)), dim3(NW_THREADS), 0, st, part, nblocks,
               (int)m, (int)n, dw, lddw, db);
    return check_launch("gemm_wgrad_bias(narrow)", 2);
}

}  // namespace plagnn
