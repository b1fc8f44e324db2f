// K3 — C-ABI front end of the dense contraction: argument checks and backend choice.
// (reference call sites: code/model.py:16-17,20-28 nn.Linear / SAGEConv fc_* and their backward.)
#include "common.cuh"
#include <cstdlib>

namespace plagnn {
struct GemmParams;
// gemm_simt.cu
int gemm_simt_entry(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs, const float* bias, int act,
                    float slope, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc, void* workspace,
                    size_t workspace_bytes, cudaStream_t st);
// gemm_tc.cu
size_t gemm_tc_workspace_bytes(int64_t m, int64_t n, int64_t k_total);
bool gemm_tc_eligible(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs);
int gemm_tc_launch(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs, const float* bias, int act,
                   float slope, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc, void* workspace,
                   size_t workspace_bytes, cudaStream_t st);
// gemm_tma.cu
size_t gemm_tma_partial_bytes(int64_t m, int64_t n, int64_t k_total);
bool gemm_tma_eligible(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs);
int gemm_tma_launch(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs, const float* bias, int act,
                    float slope, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc, void* workspace,
                    size_t workspace_bytes, cudaStream_t st, float* ones_out = nullptr);
// gemm_narrow.cu
bool gemm_narrow_eligible(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs);
int gemm_narrow_launch(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs, const float* bias, int act,
                       float slope, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc, cudaStream_t st);
size_t gemm_narrow_wgrad_bytes(int64_t m, int64_t n, int64_t k);
int gemm_narrow_wgrad_launch(int64_t m, int64_t n, const float* dz, int64_t lddz, const float* x, int64_t ldx, int64_t k,
                             float* dw, int64_t lddw, float* db, void* workspace, cudaStream_t st);

static int forced_backend() {
    static const int forced = [] {
        const char* e = getenv("PLAGNN_GEMM");
        if (!e) return (int)PLAGNN_GEMM_AUTO;
        if (e[0] == 's' || e[0] == 'S') return (int)PLAGNN_GEMM_SIMT;
        if (e[0] == 'w' || e[0] == 'W') return -1;      // "wide": AUTO without the narrow kernels (A/B measurements)
        if (e[0] == 't' || e[0] == 'T') return (e[1] == 'm' || e[1] == 'M') ? (int)PLAGNN_GEMM_TMA : (int)PLAGNN_GEMM_TCGEN05;
        return (int)PLAGNN_GEMM_AUTO;
    }();
    return forced;
}
}  // namespace plagnn

using namespace plagnn;

extern "C" {

size_t plagnn_gemm_workspace_bytes(int64_t m, int64_t n, int64_t k_total) {
    // split-K partials: at most 64 splits, only used when the output has fewer tiles than SMs
    const int64_t tiles = ceil_div(m, 128) * ceil_div(n, 128);
    size_t simt = 0;
    if (tiles < 148 * 2 && k_total >= 128) {
        int64_t s = ceil_div((int64_t)2 * 148, tiles);
        if (s > 64) s = 64;
        simt = (size_t)s * (size_t)m * (size_t)n * sizeof(float);
    }
    const size_t tc = gemm_tc_workspace_bytes(m, n, k_total);
    const size_t tma = gemm_tma_partial_bytes(m, n, k_total);
    size_t w = simt > tc ? simt : tc;
    w = w > tma ? w : tma;
    return align_up(w, 256);
}

size_t plagnn_gemm_wgrad_bias_workspace_bytes(int64_t m, int64_t n, int64_t k) {
    const size_t a = plagnn_gemm_workspace_bytes(m, n + 1, k), b = plagnn_colsum_workspace_bytes(k, m);
    const size_t c = gemm_narrow_wgrad_bytes(m, n, k);
    return (a > b ? a : b) > c ? (a > b ? a : b) : c;
}

// dW[m x n] = dZ^T X and db[m] = column sums of dZ, in one pass when the TMA kernel can take the product (the bias gradient
// is the extra output column of a B operand whose column n reads as 1.0); otherwise plagnn_gemm + plagnn_colsum.
int plagnn_gemm_wgrad_bias(int64_t m, int64_t n, const float* dz, int64_t lddz, const float* x, int64_t ldx, int64_t k,
                           float* dw, int64_t lddw, float* db, void* workspace, size_t workspace_bytes,
                           plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (m <= 0 || n <= 0 || k <= 0 || !dz || !x || !dw || !db || lddz < m || ldx < n || lddw < n)
        return fail(PLAGNN_ERR_ARG, "gemm_wgrad_bias", "bad sizes or null pointers");
    plagnn_gemm_pair p{dz, lddz, 1, x, ldx, 1, k};
    const size_t part = gemm_tma_partial_bytes(m, n + 1, k);
    int forced = forced_backend();
    const size_t narrow = gemm_narrow_wgrad_bytes(m, n, k);
    if (forced == PLAGNN_GEMM_AUTO && narrow > 0 && workspace && narrow <= workspace_bytes) {
        ProfileScope prof("gemm", m, n, k, stream);
        return gemm_narrow_wgrad_launch(m, n, dz, lddz, x, ldx, k, dw, lddw, db, workspace, st);
    }
    if (forced < 0) forced = PLAGNN_GEMM_AUTO;
    if ((forced == PLAGNN_GEMM_AUTO || forced == PLAGNN_GEMM_TMA) && part > 0 && workspace && part <= workspace_bytes &&
        gemm_tma_eligible(m, n + 1, 1, &p)) {
        ProfileScope prof("gemm", m, n, k, stream);
        return gemm_tma_launch(m, n, 1, &p, nullptr, PLAGNN_ACT_NONE, 0.f, nullptr, 0, PLAGNN_ACT_NONE, dw, lddw, workspace,
                               workspace_bytes, st, db);
    }
    int rc = plagnn_gemm(m, n, 1, &p, nullptr, PLAGNN_ACT_NONE, 0.f, nullptr, 0, PLAGNN_ACT_NONE, dw, lddw, workspace,
                         workspace_bytes, PLAGNN_GEMM_AUTO, stream);
    if (rc) return rc;
    return plagnn_colsum(dz, k, m, lddz, db, workspace, workspace_bytes, stream);
}

int plagnn_gemm(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs, const float* bias, int act,
                float slope, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc, void* workspace,
                size_t workspace_bytes, int backend, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    long long ktot = 0;
    for (int p = 0; pairs && p < npairs && p < PLAGNN_GEMM_MAX_PAIRS; ++p) ktot += pairs[p].k;
    ProfileScope prof("gemm", m, n, ktot, stream);
    if (m <= 0 || n <= 0 || npairs < 1 || npairs > PLAGNN_GEMM_MAX_PAIRS || !pairs || !c)
        return fail(PLAGNN_ERR_ARG, "gemm", "bad sizes or null pointers");
    if (act < PLAGNN_ACT_NONE || act > PLAGNN_ACT_SIGMOID || gate_act < PLAGNN_ACT_NONE || gate_act > PLAGNN_ACT_SIGMOID)
        return fail(PLAGNN_ERR_ARG, "gemm", "unknown activation");
    if (ldc < n || (gate && ldg < n)) return fail(PLAGNN_ERR_ARG, "gemm", "output/gate pitch smaller than n");
    for (int p = 0; p < npairs; ++p) {
        const plagnn_gemm_pair& q = pairs[p];
        if (!q.a || !q.b || q.k <= 0) return fail(PLAGNN_ERR_ARG, "gemm", "null operand or k <= 0");
        if (q.lda < (q.a_trans ? m : q.k) || q.ldb < (q.b_trans ? n : q.k))
            return fail(PLAGNN_ERR_ARG, "gemm", "operand pitch too small");
    }
    bool narrow_ok = gemm_narrow_eligible(m, n, npairs, pairs);
    if (backend == PLAGNN_GEMM_AUTO) {
        backend = forced_backend();
        if (backend < 0) { backend = PLAGNN_GEMM_AUTO; narrow_ok = false; }
    }
    if (backend == PLAGNN_GEMM_NARROW && !narrow_ok)
        return fail(PLAGNN_ERR_UNSUPPORTED, "gemm", "narrow backend needs n <= 16 (k <= 4096) or a contraction of at most 32");
    if (backend == PLAGNN_GEMM_NARROW || (backend == PLAGNN_GEMM_AUTO && narrow_ok))
        return gemm_narrow_launch(m, n, npairs, pairs, bias, act, slope, gate, ldg, gate_act, c, ldc, st);
    const bool tc_ok = gemm_tc_eligible(m, n, npairs, pairs);
    const size_t part_bytes = gemm_tma_partial_bytes(m, n, ktot);
    const bool tma_ok = tc_ok && gemm_tma_eligible(m, n, npairs, pairs) &&
                        (part_bytes == 0 || (workspace && part_bytes <= workspace_bytes));
    if (backend == PLAGNN_GEMM_TMA && !tma_ok)
        return fail(PLAGNN_ERR_UNSUPPORTED, "gemm", "TMA backend needs n >= 16, k >= 8, 16-byte aligned rows and the split-K workspace");
    if (backend == PLAGNN_GEMM_TCGEN05 && !tc_ok)
        return fail(PLAGNN_ERR_UNSUPPORTED, "gemm", "tcgen05 backend needs n >= 16 and k >= 8");
    if (backend == PLAGNN_GEMM_AUTO) backend = tma_ok ? PLAGNN_GEMM_TMA : tc_ok ? PLAGNN_GEMM_TCGEN05 : PLAGNN_GEMM_SIMT;
    if (backend == PLAGNN_GEMM_TMA)
        return gemm_tma_launch(m, n, npairs, pairs, bias, act, slope, gate, ldg, gate_act, c, ldc, workspace, workspace_bytes, st);
    if (backend == PLAGNN_GEMM_TCGEN05)
        return gemm_tc_launch(m, n, npairs, pairs, bias, act, slope, gate, ldg, gate_act, c, ldc, workspace,
                              workspace_bytes, st);
    if (backend == PLAGNN_GEMM_SIMT)
        return gemm_simt_entry(m, n, npairs, pairs, bias, act, slope, gate, ldg, gate_act, c, ldc, workspace,
                               workspace_bytes, st);
    return fail(PLAGNN_ERR_ARG, "gemm", "unknown backend");
}

}  // extern "C"
