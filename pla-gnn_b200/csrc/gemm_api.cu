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
bool gemm_tma_eligible(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair_ex* pairs);
int gemm_tma_launch(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair_ex* pairs, const float* bias, int act,
                    float slope, const float* gate, int64_t ldg, int gate_act, float* c, float* c_lo, int64_t ldc,
                    void* workspace, size_t workspace_bytes, cudaStream_t st);
int tf32_lo_launch(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* lo, int64_t ldlo, cudaStream_t st);

static inline int64_t pitch32(int64_t f) { return (f + 31) / 32 * 32; }
// companions the plain plagnn_gemm() derives itself, in the caller's workspace behind the split-K partials
static size_t companion_bytes(int64_t m, int64_t n, int64_t k_total) {
    return (size_t)4 * (size_t)((m + 32) + (n + 32)) * (size_t)(k_total + 32 * PLAGNN_GEMM_MAX_PAIRS + 32) +
           (size_t)256 * 2 * PLAGNN_GEMM_MAX_PAIRS;
}

static int forced_backend() {
    static const int forced = [] {
        const char* e = getenv("PLAGNN_GEMM");
        if (!e) return (int)PLAGNN_GEMM_AUTO;
        if (e[0] == 's' || e[0] == 'S') return (int)PLAGNN_GEMM_SIMT;
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
    const size_t tma = gemm_tma_partial_bytes(m, n, k_total) + companion_bytes(m, n, k_total);
    size_t w = simt > tc ? simt : tc;
    w = w > tma ? w : tma;
    return align_up(w, 256);
}

/* split-K partials only: what plagnn_gemm_ex() needs when every operand comes with its companion */
size_t plagnn_gemm_ex_workspace_bytes(int64_t m, int64_t n, int64_t k_total) {
    return gemm_tma_partial_bytes(m, n, k_total);
}

int plagnn_tf32_lo(const float* x, int64_t ldx, int64_t rows, int64_t cols, float* lo, int64_t ldlo,
                   plagnn_stream_t stream) {
    ProfileScope prof("tf32_lo", rows, cols, 0, stream);
    if (!x || !lo || rows <= 0 || cols <= 0 || ldx < cols || ldlo < cols) return fail(PLAGNN_ERR_ARG, "tf32_lo", "bad sizes or null pointers");
    return tf32_lo_launch(x, ldx, rows, cols, lo, ldlo, (cudaStream_t)stream);
}

int plagnn_gemm_ex(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair_ex* pairs, const float* bias, int act,
                   float slope, const float* gate, int64_t ldg, int gate_act, float* c, float* c_lo, int64_t ldc,
                   void* workspace, size_t workspace_bytes, plagnn_stream_t stream) {
    long long ktot = 0;
    for (int p = 0; pairs && p < npairs && p < PLAGNN_GEMM_MAX_PAIRS; ++p) ktot += pairs[p].k;
    ProfileScope prof("gemm", m, n, ktot, stream);
    if (m <= 0 || n <= 0 || npairs < 1 || npairs > PLAGNN_GEMM_MAX_PAIRS || !pairs || !c)
        return fail(PLAGNN_ERR_ARG, "gemm_ex", "bad sizes or null pointers");
    if (act < PLAGNN_ACT_NONE || act > PLAGNN_ACT_SIGMOID || gate_act < PLAGNN_ACT_NONE || gate_act > PLAGNN_ACT_SIGMOID)
        return fail(PLAGNN_ERR_ARG, "gemm_ex", "unknown activation");
    if (ldc < n || (gate && ldg < n)) return fail(PLAGNN_ERR_ARG, "gemm_ex", "output/gate pitch smaller than n");
    for (int p = 0; p < npairs; ++p) {
        const plagnn_gemm_pair_ex& q = pairs[p];
        if (!q.a || !q.b || !q.a_lo || !q.b_lo || q.k <= 0) return fail(PLAGNN_ERR_ARG, "gemm_ex", "null operand / companion or k <= 0");
        if (q.lda < (q.a_trans ? m : q.k) || q.ldb < (q.b_trans ? n : q.k) || q.lda_lo < (q.a_trans ? m : q.k) ||
            q.ldb_lo < (q.b_trans ? n : q.k))
            return fail(PLAGNN_ERR_ARG, "gemm_ex", "operand pitch too small");
    }
    if (!gemm_tma_eligible(m, n, npairs, pairs))
        return fail(PLAGNN_ERR_UNSUPPORTED, "gemm_ex", "needs n >= 16, k >= 8, 16-byte aligned rows, one storage order per side");
    return gemm_tma_launch(m, n, npairs, pairs, bias, act, slope, gate, ldg, gate_act, c, c_lo, ldc, workspace,
                           workspace_bytes, (cudaStream_t)stream);
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
    if (backend == PLAGNN_GEMM_AUTO) backend = forced_backend();
    const bool tc_ok = gemm_tc_eligible(m, n, npairs, pairs);
    // TMA-fed kernel: the companions of all operands are derived here, into the workspace behind the split-K partials
    plagnn_gemm_pair_ex ex[PLAGNN_GEMM_MAX_PAIRS];
    bool tma_ok = tc_ok;
    const size_t part_bytes = gemm_tma_partial_bytes(m, n, ktot);
    // default: lo tiles derived inside the kernel; PLAGNN_TMA_COMPANION=1 derives companion matrices first (comparison)
    static const bool companion_mode = getenv("PLAGNN_TMA_COMPANION") != nullptr;
    if (tma_ok && !companion_mode) {
        for (int p = 0; p < npairs; ++p) {
            const plagnn_gemm_pair& q = pairs[p];
            ex[p].a = q.a; ex[p].lda = q.lda; ex[p].a_trans = q.a_trans; ex[p].a_lo = nullptr; ex[p].lda_lo = 0;
            ex[p].b = q.b; ex[p].ldb = q.ldb; ex[p].b_trans = q.b_trans; ex[p].b_lo = nullptr; ex[p].ldb_lo = 0;
            ex[p].k = q.k;
        }
        tma_ok = (part_bytes == 0 || (workspace && part_bytes <= workspace_bytes)) && gemm_tma_eligible(m, n, npairs, ex);
    } else if (tma_ok) {
        size_t off = part_bytes;
        for (int p = 0; p < npairs; ++p) {
            const plagnn_gemm_pair& q = pairs[p];
            ex[p].a = q.a; ex[p].lda = q.lda; ex[p].a_trans = q.a_trans;
            ex[p].b = q.b; ex[p].ldb = q.ldb; ex[p].b_trans = q.b_trans;
            ex[p].k = q.k;
            const int64_t a_outer = q.a_trans ? q.k : m, a_inner = q.a_trans ? m : q.k;
            const int64_t b_outer = q.b_trans ? q.k : n, b_inner = q.b_trans ? n : q.k;
            ex[p].lda_lo = pitch32(a_inner);
            ex[p].ldb_lo = pitch32(b_inner);
            off = align_up(off, 256);
            ex[p].a_lo = workspace ? reinterpret_cast<const float*>(static_cast<char*>(workspace) + off) : nullptr;
            off += (size_t)a_outer * ex[p].lda_lo * 4;
            off = align_up(off, 256);
            ex[p].b_lo = workspace ? reinterpret_cast<const float*>(static_cast<char*>(workspace) + off) : nullptr;
            off += (size_t)b_outer * ex[p].ldb_lo * 4;
        }
        tma_ok = workspace && off <= workspace_bytes && gemm_tma_eligible(m, n, npairs, ex);
    }
    if (backend == PLAGNN_GEMM_TMA && !tma_ok)
        return fail(PLAGNN_ERR_UNSUPPORTED, "gemm", "TMA backend needs n >= 16, k >= 8, aligned rows and the full workspace");
    if (backend == PLAGNN_GEMM_TCGEN05 && !tc_ok)
        return fail(PLAGNN_ERR_UNSUPPORTED, "gemm", "tcgen05 backend needs n >= 16 and k >= 8");
    if (backend == PLAGNN_GEMM_AUTO) backend = tma_ok ? PLAGNN_GEMM_TMA : tc_ok ? PLAGNN_GEMM_TCGEN05 : PLAGNN_GEMM_SIMT;
    if (backend == PLAGNN_GEMM_TMA) {
        for (int p = 0; companion_mode && p < npairs; ++p) {
            const plagnn_gemm_pair_ex& q = ex[p];
            int rc = tf32_lo_launch(q.a, q.lda, q.a_trans ? q.k : m, q.a_trans ? m : q.k, const_cast<float*>(q.a_lo), q.lda_lo, st);
            if (rc) return rc;
            rc = tf32_lo_launch(q.b, q.ldb, q.b_trans ? q.k : n, q.b_trans ? n : q.k, const_cast<float*>(q.b_lo), q.ldb_lo, st);
            if (rc) return rc;
        }
        return gemm_tma_launch(m, n, npairs, ex, bias, act, slope, gate, ldg, gate_act, c, nullptr, ldc, workspace, part_bytes, st);
    }
    if (backend == PLAGNN_GEMM_TCGEN05)
        return gemm_tc_launch(m, n, npairs, pairs, bias, act, slope, gate, ldg, gate_act, c, ldc, workspace,
                              workspace_bytes, st);
    if (backend == PLAGNN_GEMM_SIMT)
        return gemm_simt_entry(m, n, npairs, pairs, bias, act, slope, gate, ldg, gate_act, c, ldc, workspace,
                               workspace_bytes, st);
    return fail(PLAGNN_ERR_ARG, "gemm", "unknown backend");
}

}  // extern "C"
