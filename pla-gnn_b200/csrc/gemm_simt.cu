// K3 (fallback backend) — fp32 FFMA GEMM with the same contract as the tcgen05 backend in gemm_tc.cu.
//
// Replaces the cuBLAS SGEMMs behind nn.Linear in code/model.py:16-17,20-28 and their backward.
// 128x128x16 block tile, 256 threads, 8x8 register tile, double-buffered shared memory, operands in
// either storage order, up to two (A,B) pairs accumulated into one tile, optional split-K with an
// ordered (deterministic) reduction, fused bias / activation / activation-gradient epilogue.
// This backend is exact fp32 FMA arithmetic and is what the tcgen05 path is validated against; it also
// serves shapes the tensor-core path does not take (n < 16, unaligned operands).
#include "common.cuh"

namespace plagnn {

constexpr int BM = 128, BN = 128, BK = 16, GT = 256;
constexpr int SPAD = 4;

struct GemmParams {
    int64_t m, n;
    int npairs;
    const float* a[PLAGNN_GEMM_MAX_PAIRS];
    const float* b[PLAGNN_GEMM_MAX_PAIRS];
    int64_t lda[PLAGNN_GEMM_MAX_PAIRS], ldb[PLAGNN_GEMM_MAX_PAIRS];
    int a_trans[PLAGNN_GEMM_MAX_PAIRS], b_trans[PLAGNN_GEMM_MAX_PAIRS];
    int a_vec[PLAGNN_GEMM_MAX_PAIRS], b_vec[PLAGNN_GEMM_MAX_PAIRS];
    int64_t k[PLAGNN_GEMM_MAX_PAIRS];
    int ktiles[PLAGNN_GEMM_MAX_PAIRS];
    int total_ktiles, tiles_per_split, splits;
    const float* bias;
    int act;
    float slope;
    const float* gate;
    int64_t ldg;
    int gate_act;
    float* c;
    int64_t ldc;
    float* partial;   // [splits][m][n] when splits > 1
};

// load a 128(rows) x 16(k) operand tile into registers: 8 floats per thread.
// trans == 0: storage [rows x k], k contiguous; thread t -> row r = t/4 (+64), kq = (t%4)*4
// trans == 1: storage [k x rows], rows contiguous; thread t -> kk = t/32 (+8), rq = (t%32)*4
__device__ __forceinline__ void load_tile(const float* __restrict__ p, int64_t ld, int trans, int vec, int64_t rows,
                                          int64_t kdim, int64_t r0, int64_t k0, float (&reg)[8]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!trans) {
            const int64_t r = r0 + (t >> 2) + 64 * i;
            const int64_t kk = k0 + (t & 3) * 4;
            if (r < rows && kk < kdim) {
                const float* q = p + r * ld + kk;
                if (vec && kk + 3 < kdim) {
                    v = ldg_f4(q);
                } else {
                    v.x = __ldg(q);
                    if (kk + 1 < kdim) v.y = __ldg(q + 1);
                    if (kk + 2 < kdim) v.z = __ldg(q + 2);
                    if (kk + 3 < kdim) v.w = __ldg(q + 3);
                }
            }
        } else {
            const int64_t kk = k0 + (t >> 5) + 8 * i;
            const int64_t r = r0 + (t & 31) * 4;
            if (kk < kdim && r < rows) {
                const float* q = p + kk * ld + r;
                if (vec && r + 3 < rows) {
                    v = ldg_f4(q);
                } else {
                    v.x = __ldg(q);
                    if (r + 1 < rows) v.y = __ldg(q + 1);
                    if (r + 2 < rows) v.z = __ldg(q + 2);
                    if (r + 3 < rows) v.w = __ldg(q + 3);
                }
            }
        }
        reg[4 * i + 0] = v.x; reg[4 * i + 1] = v.y; reg[4 * i + 2] = v.z; reg[4 * i + 3] = v.w;
    }
}

// shared tile layout S[k][row] with row stride BM+SPAD
__device__ __forceinline__ void store_tile(float* __restrict__ s, int trans, const float (&reg)[8]) {
    const int t = threadIdx.x;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        if (!trans) {
            const int r = (t >> 2) + 64 * i;
            const int kq = (t & 3) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) s[(kq + j) * (BM + SPAD) + r] = reg[4 * i + j];
        } else {
            const int kk = (t >> 5) + 8 * i;
            const int rq = (t & 31) * 4;
            *reinterpret_cast<float4*>(s + kk * (BM + SPAD) + rq) =
                make_float4(reg[4 * i], reg[4 * i + 1], reg[4 * i + 2], reg[4 * i + 3]);
        }
    }
}

__device__ __forceinline__ float epilogue_one(const GemmParams& P, float v, int64_t r, int64_t c) {
    if (P.bias) v += __ldg(P.bias + c);
    v = apply_act(v, P.act, P.slope);
    if (P.gate) v *= act_grad_from_output(__ldg(P.gate + r * P.ldg + c), P.gate_act, P.slope);
    return v;
}

__global__ void __launch_bounds__(GT, 2) gemm_simt_kernel(const GemmParams P) {
    __shared__ __align__(16) float As[2][BK * (BM + SPAD)];
    __shared__ __align__(16) float Bs[2][BK * (BN + SPAD)];
    const int t = threadIdx.x;
    const int ty = t >> 4, tx = t & 15;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int split = blockIdx.z;
    const int kt_beg = split * P.tiles_per_split;
    const int kt_end = min(P.total_ktiles, kt_beg + P.tiles_per_split);

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    auto fetch = [&](int kt) {
        int p = 0, local = kt;
        if (P.npairs > 1 && local >= P.ktiles[0]) { local -= P.ktiles[0]; p = 1; }
        const int64_t k0 = (int64_t)local * BK;
        load_tile(P.a[p], P.lda[p], P.a_trans[p], P.a_vec[p], P.m, P.k[p], m0, k0, ra);
        load_tile(P.b[p], P.ldb[p], P.b_trans[p], P.b_vec[p], P.n, P.k[p], n0, k0, rb);
        return p;
    };
    int buf = 0;
    if (kt_beg < kt_end) {
        const int p = fetch(kt_beg);
        store_tile(As[0], P.a_trans[p], ra);
        store_tile(Bs[0], P.b_trans[p], rb);
    }
    __syncthreads();
    for (int kt = kt_beg; kt < kt_end; ++kt) {
        int pn = 0;
        const bool more = kt + 1 < kt_end;
        if (more) pn = fetch(kt + 1);
        const float* as = As[buf];
        const float* bs = Bs[buf];
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(as + kk * (BM + SPAD) + ty * 4);
            const float4 a1 = *reinterpret_cast<const float4*>(as + kk * (BM + SPAD) + 64 + ty * 4);
            const float4 b0 = *reinterpret_cast<const float4*>(bs + kk * (BN + SPAD) + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(bs + kk * (BN + SPAD) + 64 + tx * 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            store_tile(As[buf ^ 1], P.a_trans[pn], ra);
            store_tile(Bs[buf ^ 1], P.b_trans[pn], rb);
        }
        __syncthreads();
        buf ^= 1;
    }

    const bool direct = P.splits == 1;
    float* dst = direct ? P.c : P.partial + (int64_t)split * P.m * P.n;
    const int64_t ldd = direct ? P.ldc : P.n;
    const bool vec_out = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (r >= P.m) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t c = n0 + h * 64 + tx * 4;
            if (c >= P.n) continue;
            float v[4] = {acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]};
            if (direct) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < P.n) v[j] = epilogue_one(P, v[j], r, c + j);
            }
            if (vec_out && c + 3 < P.n) {
                *reinterpret_cast<float4*>(dst + r * ldd + c) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (c + j < P.n) dst[r * ldd + c + j] = v[j];
            }
        }
    }
}

// ordered reduction of the split-K partials + epilogue
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(const GemmParams P) {
    const int64_t total = P.m * P.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int z = 0; z < P.splits; ++z) s += P.partial[(int64_t)z * total + i];
        const int64_t r = i / P.n, c = i - r * P.n;
        P.c[r * P.ldc + c] = epilogue_one(P, s, r, c);
    }
}

// fills GemmParams from the C-ABI arguments; shared with the tcgen05 backend for validation
int gemm_fill_params(GemmParams& P, int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs,
                     const float* bias, int act, float slope, const float* gate, int64_t ldg, int gate_act, float* c,
                     int64_t ldc) {
    if (m <= 0 || n <= 0 || npairs < 1 || npairs > PLAGNN_GEMM_MAX_PAIRS || !pairs || !c)
        return fail(PLAGNN_ERR_ARG, "gemm", "bad sizes or null pointers");
    if (act < PLAGNN_ACT_NONE || act > PLAGNN_ACT_SIGMOID || gate_act < PLAGNN_ACT_NONE || gate_act > PLAGNN_ACT_SIGMOID)
        return fail(PLAGNN_ERR_ARG, "gemm", "unknown activation");
    if (ldc < n || (gate && ldg < n)) return fail(PLAGNN_ERR_ARG, "gemm", "output/gate pitch smaller than n");
    P.m = m; P.n = n; P.npairs = npairs;
    P.total_ktiles = 0;
    for (int p = 0; p < npairs; ++p) {
        const plagnn_gemm_pair& q = pairs[p];
        if (!q.a || !q.b || q.k <= 0) return fail(PLAGNN_ERR_ARG, "gemm", "null operand or k <= 0");
        if (q.lda < (q.a_trans ? m : q.k) || q.ldb < (q.b_trans ? n : q.k)) return fail(PLAGNN_ERR_ARG, "gemm", "operand pitch too small");
        P.a[p] = q.a; P.b[p] = q.b; P.lda[p] = q.lda; P.ldb[p] = q.ldb;
        P.a_trans[p] = q.a_trans ? 1 : 0; P.b_trans[p] = q.b_trans ? 1 : 0;
        P.a_vec[p] = ((q.lda & 3) == 0 && aligned16(q.a)) ? 1 : 0;
        P.b_vec[p] = ((q.ldb & 3) == 0 && aligned16(q.b)) ? 1 : 0;
        P.k[p] = q.k;
        P.ktiles[p] = (int)ceil_div(q.k, BK);
        P.total_ktiles += P.ktiles[p];
    }
    for (int p = npairs; p < PLAGNN_GEMM_MAX_PAIRS; ++p) {
        P.a[p] = P.a[0]; P.b[p] = P.b[0]; P.lda[p] = P.lda[0]; P.ldb[p] = P.ldb[0];
        P.a_trans[p] = P.a_trans[0]; P.b_trans[p] = P.b_trans[0]; P.a_vec[p] = 0; P.b_vec[p] = 0; P.k[p] = 0; P.ktiles[p] = 0;
    }
    P.bias = bias; P.act = act; P.slope = slope; P.gate = gate; P.ldg = ldg; P.gate_act = gate_act;
    P.c = c; P.ldc = ldc; P.partial = nullptr; P.splits = 1; P.tiles_per_split = P.total_ktiles;
    return PLAGNN_OK;
}

static int choose_splits(int64_t m, int64_t n, int total_ktiles) {
    const int64_t tiles = ceil_div(m, BM) * ceil_div(n, BN);
    const int sms = sm_count();
    if (tiles >= sms || total_ktiles < 16) return 1;
    int64_t s = ceil_div(2 * (int64_t)sms, tiles);
    if (s > total_ktiles / 8) s = total_ktiles / 8;
    if (s > 64) s = 64;
    return s < 1 ? 1 : (int)s;
}

int gemm_simt_launch(GemmParams& P, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    int splits = choose_splits(P.m, P.n, P.total_ktiles);
    if (splits > 1) {
        const size_t need = (size_t)splits * P.m * P.n * sizeof(float);
        if (!workspace || workspace_bytes < need) splits = 1;   // still correct, just less parallel
    }
    P.splits = splits;
    P.tiles_per_split = (int)ceil_div(P.total_ktiles, splits);
    P.splits = (int)ceil_div(P.total_ktiles, P.tiles_per_split);
    P.partial = P.splits > 1 ? (float*)workspace : nullptr;
    dim3 grid((unsigned)ceil_div(P.n, BN), (unsigned)ceil_div(P.m, BM), (unsigned)P.splits);
    gemm_simt_kernel<<<grid, GT, 0, st>>>(P);
    if (P.splits > 1) {
        const int64_t total = P.m * P.n;
        const int g = (int)(ceil_div(total, 256) < (int64_t)sm_count() * 8 ? ceil_div(total, 256) : (int64_t)sm_count() * 8);
        gemm_splitk_reduce_kernel<<<g, 256, 0, st>>>(P);
    }
    return check_launch("gemm_simt", P.splits > 1 ? 2 : 1);
}

}  // namespace plagnn

namespace plagnn {
int gemm_simt_entry(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs, const float* bias, int act,
                    float slope, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc, void* workspace,
                    size_t workspace_bytes, cudaStream_t st) {
    GemmParams P;
    int rc = gemm_fill_params(P, m, n, npairs, pairs, bias, act, slope, gate, ldg, gate_act, c, ldc);
    if (rc) return rc;
    return gemm_simt_launch(P, workspace, workspace_bytes, st);
}
}  // namespace plagnn
