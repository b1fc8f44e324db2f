// Whole-network entry points: one C call runs the forward pass of GNN32, one call its backward pass.
//
// Replaces, as a unit,  code/model.py:19-31 (GNN32.forward: 3x SAGEConv('pool') + leaky_relu, liner1 +
// leaky_relu, liner2 + sigmoid)  and the autograd backward that  code/train.py:204  triggers through it.
// Host-side orchestration only: every kernel launched here is one of the library's own (gemm, spmm, colsum,
// act_backward, pad_copy, transpose); the caller owns one arena that holds every intermediate, so an epoch
// performs no allocation and ~80 launches leave from C++ instead of from Python.
#include "common.cuh"

namespace plagnn {

static inline int64_t pitch32(int64_t f) { return (f + 31) / 32 * 32; }

struct Bump {
    char* base;
    size_t off;
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};

struct Layout {
    // dims: d[0] = in, d[1..3] = conv outputs, d[4] = liner1 out, d[5] = classes
    int64_t n, d[6];
    // persistent between forward and backward
    float* neigh[3];
    int32_t* arg[3];
    float* h[3];          // conv outputs after leaky_relu
    float* h4;
    float *wp[3], *ws[3], *wn[3], *w1, *w2;      // row-padded weight copies
    // temporaries
    float *m, *dz5, *dz4, *drst[2], *dneigh, *dm, *wt[2];
    void *gemm_ws, *colsum_ws, *spmm_ws;
    size_t gemm_ws_bytes, colsum_ws_bytes, spmm_ws_bytes;
    size_t total;
};

static Layout make_layout(const plagnn_gnn32_shape* s, void* arena) {
    Layout L{};
    L.n = s->num_nodes;
    L.d[0] = s->in_feats; L.d[1] = s->h1; L.d[2] = s->h2; L.d[3] = s->h3; L.d[4] = s->h4; L.d[5] = s->classes;
    Bump b{(char*)arena, 0};
    const int64_t n = L.n;
    int64_t fmax = 0, omax = 0;
    for (int l = 0; l < 3; ++l) {
        const int64_t f = L.d[l], o = L.d[l + 1];
        fmax = f > fmax ? f : fmax;
        omax = o > omax ? o : omax;
        L.neigh[l] = b.take<float>(n * pitch32(f));
        L.arg[l] = b.take<int32_t>(n * pitch32(f));
        L.h[l] = b.take<float>(n * pitch32(o));
        L.wp[l] = b.take<float>(f * pitch32(f));
        L.ws[l] = b.take<float>(o * pitch32(f));
        L.wn[l] = b.take<float>(o * pitch32(f));
    }
    L.h4 = b.take<float>(n * pitch32(L.d[4]));
    L.w1 = b.take<float>(L.d[4] * pitch32(L.d[3]));
    L.w2 = b.take<float>(L.d[5] * pitch32(L.d[4]));
    L.m = b.take<float>(n * pitch32(fmax));
    L.dz5 = b.take<float>(n * pitch32(L.d[5]));
    L.dz4 = b.take<float>(n * pitch32(L.d[4]));
    L.drst[0] = b.take<float>(n * pitch32(omax > fmax ? omax : fmax));
    L.drst[1] = b.take<float>(n * pitch32(omax > fmax ? omax : fmax));
    L.dneigh = b.take<float>(n * pitch32(fmax));
    L.dm = b.take<float>(n * pitch32(fmax));
    L.wt[0] = b.take<float>(fmax * pitch32(fmax));
    L.wt[1] = b.take<float>(fmax * pitch32(fmax));
    // workspaces: the largest request of any call made below
    size_t g = 0;
    auto upd = [&](int64_t mm, int64_t nn, int64_t kk) {
        const size_t w = plagnn_gemm_workspace_bytes(mm, nn, kk);
        g = w > g ? w : g;
    };
    auto updw = [&](int64_t mm, int64_t nn, int64_t kk) {
        const size_t w = plagnn_gemm_wgrad_bias_workspace_bytes(mm, nn, kk);
        g = w > g ? w : g;
    };
    for (int l = 0; l < 3; ++l) {
        const int64_t f = L.d[l], o = L.d[l + 1];
        upd(n, f, f); upd(n, o, 2 * f); upd(o, f, n); upd(f, f, n); upd(n, f, o); upd(n, f, o + f);
        updw(o, f, n); updw(f, f, n);
    }
    upd(n, L.d[4], L.d[3]); upd(n, L.d[5], L.d[4]); upd(L.d[5], L.d[4], n); upd(L.d[4], L.d[3], n);
    upd(n, L.d[4], L.d[5]); upd(n, L.d[3], L.d[4]);
    updw(L.d[5], L.d[4], n); updw(L.d[4], L.d[3], n);
    L.gemm_ws_bytes = g;
    L.gemm_ws = b.take<char>(g);
    L.colsum_ws_bytes = plagnn_colsum_workspace_bytes(n, pitch32(fmax));
    L.colsum_ws = b.take<char>(L.colsum_ws_bytes);
    L.spmm_ws_bytes = plagnn_spmm_partial_bytes(s->plan_counts[2], fmax, PLAGNN_REDUCE_MAX);
    L.spmm_ws = b.take<char>(L.spmm_ws_bytes);
    L.total = align_up(b.off, 256);
    return L;
}

// A weight matrix [rows x cols] as a GEMM operand: the parameter itself when its rows are 16-byte aligned (cols % 4 == 0),
// else the row-padded copy in the arena (only the 503-wide first-layer weights need one).
struct WeightRef {
    const float* p;
    int64_t ld;
};
static inline bool weight_in_place(const float* w, int64_t cols) { return (cols & 3) == 0 && aligned16(w); }
static int weight_operand(const float* w, int64_t rows, int64_t cols, float* padded, bool copy, plagnn_stream_t st, WeightRef* out) {
    if (weight_in_place(w, cols)) {
        out->p = w; out->ld = cols;
        return PLAGNN_OK;
    }
    out->p = padded; out->ld = pitch32(cols);
    return copy ? plagnn_pad_copy(w, rows, cols, cols, padded, pitch32(cols), st) : PLAGNN_OK;
}

static int check_shape(const plagnn_gnn32_shape* s, const char* who) {
    if (!s || s->num_nodes <= 0 || s->in_feats <= 0 || s->h1 <= 0 || s->h2 <= 0 || s->h3 <= 0 || s->h4 <= 0 ||
        s->classes <= 0 || !s->indptr || !s->indices || !s->plan)
        return fail(PLAGNN_ERR_ARG, who, "bad shape descriptor");
    return PLAGNN_OK;
}

#define TRY(expr)            \
    do {                     \
        int _rc = (expr);    \
        if (_rc) return _rc; \
    } while (0)

static int gemm1(int64_t m, int64_t n, const float* a, int64_t lda, int at, const float* b, int64_t ldb, int bt, int64_t k,
                 const float* bias, int act, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc,
                 const Layout& L, plagnn_stream_t st) {
    plagnn_gemm_pair p{a, lda, at, b, ldb, bt, k};
    return plagnn_gemm(m, n, 1, &p, bias, act, 0.01f, gate, ldg, gate_act, c, ldc, L.gemm_ws, L.gemm_ws_bytes,
                       PLAGNN_GEMM_AUTO, st);
}
static int gemm2(int64_t m, int64_t n, const float* a0, int64_t lda0, const float* b0, int64_t ldb0, int64_t k0,
                 const float* a1, int64_t lda1, const float* b1, int64_t ldb1, int64_t k1, int b_trans, const float* bias,
                 int act, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc, const Layout& L,
                 plagnn_stream_t st) {
    plagnn_gemm_pair p[2] = {{a0, lda0, 0, b0, ldb0, b_trans, k0}, {a1, lda1, 0, b1, ldb1, b_trans, k1}};
    return plagnn_gemm(m, n, 2, p, bias, act, 0.01f, gate, ldg, gate_act, c, ldc, L.gemm_ws, L.gemm_ws_bytes,
                       PLAGNN_GEMM_AUTO, st);
}

}  // namespace plagnn

using namespace plagnn;

extern "C" {

size_t plagnn_gnn32_arena_bytes(const plagnn_gnn32_shape* shape) {
    if (check_shape(shape, "gnn32_arena_bytes")) return 0;
    return make_layout(shape, nullptr).total;
}

// params (host array of 19 device pointers, contiguous tensors):
//   per conv l = 0..2: [5l+0] fc_pool.weight [F x F], [5l+1] fc_pool.bias [F], [5l+2] fc_self.weight [O x F],
//                      [5l+3] fc_neigh.weight [O x F], [5l+4] bias [O]
//   [15] liner1.weight [h4 x h3], [16] liner1.bias, [17] liner2.weight [C x h4], [18] liner2.bias
int plagnn_gnn32_forward(const plagnn_gnn32_shape* shape, const float* x, int64_t ldx, const float* const* params,
                         void* arena, size_t arena_bytes, float* prob, int64_t ldprob, plagnn_stream_t stream) {
    TRY(check_shape(shape, "gnn32_forward"));
    if (!x || !params || !arena || !prob) return fail(PLAGNN_ERR_ARG, "gnn32_forward", "null pointer");
    if ((ldx & 3) || !aligned16(x) || ldx < shape->in_feats) return fail(PLAGNN_ERR_ALIGN, "gnn32_forward", "x needs 16-byte aligned rows");
    const Layout L = make_layout(shape, arena);
    if (arena_bytes < L.total) return fail(PLAGNN_ERR_WORKSPACE, "gnn32_forward", "arena too small");
    const int64_t n = L.n;
    const float* h = x;
    int64_t ldh = ldx;
    for (int l = 0; l < 3; ++l) {
        const int64_t f = L.d[l], o = L.d[l + 1], pf = pitch32(f), po = pitch32(o);
        const float* const* P = params + 5 * l;
        // weight operands: in place when rows are 16-byte aligned, else a row-padded copy (refreshed here, after Adam)
        WeightRef wp, ws, wn;
        TRY(weight_operand(P[0], f, f, L.wp[l], true, stream, &wp));
        TRY(weight_operand(P[2], o, f, L.ws[l], true, stream, &ws));
        TRY(weight_operand(P[3], o, f, L.wn[l], true, stream, &wn));
        // m = relu(h Wp^T + bp);  neigh = max over in-neighbours;  h' = leaky(h Ws^T + neigh Wn^T + b)
        TRY(gemm1(n, f, h, ldh, 0, wp.p, wp.ld, 0, f, P[1], PLAGNN_ACT_RELU, nullptr, 0, 0, L.m, pf, L, stream));
        TRY(plagnn_spmm_max_fwd(shape->indptr, shape->indices, shape->plan, shape->plan_counts, n, L.m, pf, f,
                                L.neigh[l], L.arg[l], pf, L.spmm_ws, L.spmm_ws_bytes, stream));
        TRY(gemm2(n, o, h, ldh, ws.p, ws.ld, f, L.neigh[l], pf, wn.p, wn.ld, f, 0, P[4], PLAGNN_ACT_LEAKY, nullptr, 0, 0,
                  L.h[l], po, L, stream));
        h = L.h[l];
        ldh = po;
    }
    const int64_t d3 = L.d[3], d4 = L.d[4], c = L.d[5];
    WeightRef w1, w2;
    TRY(weight_operand(params[15], d4, d3, L.w1, true, stream, &w1));
    TRY(weight_operand(params[17], c, d4, L.w2, true, stream, &w2));
    TRY(gemm1(n, d4, h, ldh, 0, w1.p, w1.ld, 0, d3, params[16], PLAGNN_ACT_LEAKY, nullptr, 0, 0, L.h4, pitch32(d4),
              L, stream));
    TRY(gemm1(n, c, L.h4, pitch32(d4), 0, w2.p, w2.ld, 0, d4, params[18], PLAGNN_ACT_SIGMOID, nullptr, 0, 0, prob,
              ldprob, L, stream));
    return PLAGNN_OK;
}

// grads: host array of 19 device pointers laid out like params (contiguous); need_dx / dx optional.
int plagnn_gnn32_backward(const plagnn_gnn32_shape* shape, const float* x, int64_t ldx, const float* const* params,
                          void* arena, size_t arena_bytes, const float* prob, int64_t ldprob, const float* dprob,
                          int64_t lddprob, float* const* grads, float* dx, int64_t lddx, plagnn_stream_t stream) {
    TRY(check_shape(shape, "gnn32_backward"));
    if (!x || !params || !arena || !prob || !dprob || !grads) return fail(PLAGNN_ERR_ARG, "gnn32_backward", "null pointer");
    const Layout L = make_layout(shape, arena);
    if (arena_bytes < L.total) return fail(PLAGNN_ERR_WORKSPACE, "gnn32_backward", "arena too small");
    const int64_t n = L.n, d3 = L.d[3], d4 = L.d[4], c = L.d[5];
    const int64_t p3 = pitch32(d3), p4 = pitch32(d4), pc = pitch32(c);
    // same operands as the forward pass (padded copies, where needed, were refreshed there)
    WeightRef w1, w2;
    TRY(weight_operand(params[15], d4, d3, L.w1, false, stream, &w1));
    TRY(weight_operand(params[17], c, d4, L.w2, false, stream, &w2));
    // sigmoid, liner2, liner1
    TRY(plagnn_act_backward(dprob, lddprob, prob, ldprob, n, c, PLAGNN_ACT_SIGMOID, 0.01f, nullptr, L.dz5, pc, stream));
    // weight gradient + bias gradient of every linear map in one pass (the bias gradient is an extra output column)
    TRY(plagnn_gemm_wgrad_bias(c, d4, L.dz5, pc, L.h4, p4, n, grads[17], d4, grads[18], L.gemm_ws, L.gemm_ws_bytes, stream));
    // input gradients read the weights [out x in] as an MN-major B operand (b_trans = 1): no transposed copies
    TRY(gemm1(n, d4, L.dz5, pc, 0, w2.p, w2.ld, 1, c, nullptr, 0, L.h4, p4, PLAGNN_ACT_LEAKY, L.dz4, p4, L, stream));
    const float* h3 = L.h[2];
    TRY(plagnn_gemm_wgrad_bias(d4, d3, L.dz4, p4, h3, p3, n, grads[15], d3, grads[16], L.gemm_ws, L.gemm_ws_bytes, stream));
    float* drst = L.drst[0];
    TRY(gemm1(n, d3, L.dz4, p4, 0, w1.p, w1.ld, 1, d4, nullptr, 0, h3, p3, PLAGNN_ACT_LEAKY, drst, p3, L, stream));
    int cur = 0;
    for (int l = 2; l >= 0; --l) {
        const int64_t f = L.d[l], o = L.d[l + 1], pf = pitch32(f), po = pitch32(o);
        const float* hin = l == 0 ? x : L.h[l - 1];
        const int64_t ldin = l == 0 ? ldx : pf;
        float* const* G = grads + 5 * l;
        const float* const* Pl = params + 5 * l;
        WeightRef wp, ws, wn;
        TRY(weight_operand(Pl[0], f, f, L.wp[l], false, stream, &wp));
        TRY(weight_operand(Pl[2], o, f, L.ws[l], false, stream, &ws));
        TRY(weight_operand(Pl[3], o, f, L.wn[l], false, stream, &wn));
        TRY(plagnn_gemm_wgrad_bias(o, f, drst, po, hin, ldin, n, G[2], f, G[4], L.gemm_ws, L.gemm_ws_bytes, stream));
        TRY(gemm1(o, f, drst, po, 1, L.neigh[l], pf, 1, n, nullptr, 0, nullptr, 0, 0, G[3], f, L, stream));
        TRY(gemm1(n, f, drst, po, 0, wn.p, wn.ld, 1, o, nullptr, 0, nullptr, 0, 0, L.dneigh, pf, L, stream));
        TRY(plagnn_spmm_max_bwd(L.dneigh, pf, L.arg[l], pf, L.neigh[l], pf, n, f, L.dm, n, pf, stream));
        TRY(plagnn_gemm_wgrad_bias(f, f, L.dm, pf, hin, ldin, n, G[0], f, G[1], L.gemm_ws, L.gemm_ws_bytes, stream));
        if (l > 0 || dx) {
            float* out = l > 0 ? L.drst[cur ^ 1] : dx;
            const int64_t ldo = l > 0 ? pf : lddx;
            // d(input) = drst Ws + dm Wp, times leaky'(input) when the input is the previous layer's activation
            TRY(gemm2(n, f, drst, po, ws.p, ws.ld, o, L.dm, pf, wp.p, wp.ld, f, 1, nullptr, 0, l > 0 ? hin : nullptr, ldin,
                      l > 0 ? PLAGNN_ACT_LEAKY : PLAGNN_ACT_NONE, out, ldo, L, stream));
            drst = out;
            cur ^= 1;
        }
    }
    return PLAGNN_OK;
}

}  // extern "C"
