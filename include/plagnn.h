/*
 * plagnn.h — C ABI of libplagnn.so: the B200 (sm_100a) kernels behind the PLA-GNN
 * message-passing hot path.
 *
 * The reference (quinlanW/PLA-GNN) has no FFI of its own: its hot path is Python calling
 * third-party DGL / PyTorch binaries.  Each entry point below therefore names the reference
 * call site (path relative to the reference checkout, file:line) whose work it replaces.
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name says host;
 *   - matrices are row-major fp32 with an explicit row pitch ("ld", in elements);
 *   - graph structure is int32 (E' < 2^31 also for the 100M-edge synthetic graph);
 *   - the caller owns every buffer, chooses the device (cudaSetDevice) and the stream;
 *   - calls enqueue work on `stream` and return without synchronising it, except the two
 *     *_build functions, which are set-up calls and say so;
 *   - return value 0 = PLAGNN_OK, negative = error; plagnn_last_error() gives the text
 *     (thread-local).  Nothing throws, nothing allocates or frees caller memory.
 *   - re-entrant across streams and devices.
 */
#ifndef PLAGNN_H_
#define PLAGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* plagnn_stream_t; /* a cudaStream_t */

enum {
    PLAGNN_OK = 0,
    PLAGNN_ERR_ARG = -1,       /* bad argument (null pointer, negative size, unknown enum) */
    PLAGNN_ERR_ALIGN = -2,     /* pointer / pitch alignment requirement not met */
    PLAGNN_ERR_WORKSPACE = -3, /* workspace too small */
    PLAGNN_ERR_CUDA = -4,      /* CUDA runtime error at launch */
    PLAGNN_ERR_UNSUPPORTED = -5
};

enum { PLAGNN_ACT_NONE = 0, PLAGNN_ACT_RELU = 1, PLAGNN_ACT_LEAKY = 2, PLAGNN_ACT_SIGMOID = 3 };
enum { PLAGNN_GEMM_AUTO = 0, PLAGNN_GEMM_SIMT = 1, PLAGNN_GEMM_TCGEN05 = 2, PLAGNN_GEMM_TMA = 3, PLAGNN_GEMM_NARROW = 4 };
enum { PLAGNN_REDUCE_SUM = 0, PLAGNN_REDUCE_MAX = 1 };

int plagnn_version(void);
/* number of kernels this library has launched so far in this process (host-side counter) */
long long plagnn_launch_count(void);
/* Optional per-call timing with CUDA events on the launching stream (used by bench.py for the roofline):
 * enable(1) clears and starts recording every entry point, enable(2) only the aggregation (spmm_*) entry points
 * (a handful of events per epoch: cheap enough to stay on inside a timed region), enable(0) stops; report() synchronises the recorded events and
 * writes lines "name tag0 tag1 tag2 calls total_ms" (gemm: m n k; spmm_*: feat rows flag).  Returns the
 * number of bytes needed. */
int plagnn_profile_enable(int on);
size_t plagnn_profile_report(char* buf, size_t cap);
const char* plagnn_last_error(void);
/* 1 if the current device is compute capability 10.x (the only target), else 0. */
int plagnn_device_supported(void);

/* ------------------------------------------------------------------------------------------
 * K1  graph construction — replaces the lazy COO->CSR/CSC that DGL runs for
 *     code/utils.py:44-45  (dgl.graph((start,end)) + dgl.add_self_loop).
 *
 * Stable counting/radix sort of edge ids by `key` (key = destination gives the in-edge CSR
 * "CSC" that update_all(copy_u, reduce) walks; key = source gives the out-edge CSR used by the
 * transposed aggregation).  With add_self_loop != 0, N edges (i,i) with ids E..E+N-1 are
 * appended first, exactly like dgl.add_self_loop, so each row's self-loop is its last entry.
 *   indptr[N+1], indices[E'] = other endpoint, eids[E'] = original edge id, E' = E (+N).
 * Set-up call: synchronises `stream` before returning (host reads nothing back otherwise).
 * ---------------------------------------------------------------------------------------- */
size_t plagnn_csr_build_workspace_bytes(int64_t num_nodes, int64_t num_edges, int add_self_loop);
int plagnn_csr_build(const int32_t* key, const int32_t* other, int64_t num_edges, int64_t num_nodes,
                     int64_t num_other_nodes /* id range of `other`; 0 = num_nodes (row-partitioned slices differ) */,
                     int add_self_loop, int32_t* indptr, int32_t* indices, int32_t* eids,
                     void* workspace, size_t workspace_bytes, plagnn_stream_t stream);

/* Aggregation plan: splits every row's neighbour list into chunks of <= `chunk` edges so that one
 * warp owns one chunk (power-law hubs no longer serialise on one warp).
 *   plan layout (int32): see csrc/spmm.cu; sized by plagnn_spmm_plan_bytes.
 *   host_counts[0] = work items, [1] = rows split over several chunks, [2] = partial slots.
 * Set-up call: synchronises `stream`. */
size_t plagnn_spmm_plan_bytes(int64_t num_rows, int64_t num_edges, int32_t chunk);
int plagnn_spmm_plan_build(const int32_t* indptr, int64_t num_rows, int64_t num_edges, int32_t chunk,
                           void* plan, size_t plan_bytes, int64_t* host_counts, plagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2  aggregation — replaces DGL's gspmm for
 *     code/model.py:20,22,24  SAGEConv('pool'): update_all(copy_u, max)   and its autograd
 *     backward at code/train.py:204.
 *
 * max, forward:  out[v,f] = max_{u in in(v)} x[u,f]  (0 when in(v) is empty);
 *                arg[v,f] = source id of the FIRST maximum in in-edge order, -1 if none.
 * partial: scratch of plagnn_spmm_partial_bytes(plan counts, feat) bytes (may be NULL when the
 *          plan has no split rows).
 * ---------------------------------------------------------------------------------------- */
size_t plagnn_spmm_partial_bytes(int64_t partial_slots, int64_t feat, int reduce);
int plagnn_spmm_max_fwd(const int32_t* indptr, const int32_t* indices, const void* plan,
                        const int64_t* plan_counts /* host[3], from plagnn_spmm_plan_build */,
                        int64_t num_rows, const float* x, int64_t ldx, int64_t feat,
                        float* out, int32_t* arg, int64_t ldo,
                        void* partial, size_t partial_bytes, plagnn_stream_t stream);

/* max, backward: dx[arg[v,f], f] += dz[v,f] * (z ? (z[v,f] > 0) : 1).
 * `z` is the saved forward output; passing it folds the ReLU that precedes the aggregation in
 * SAGEConv-pool into the scatter (x[arg] == z, so relu'(x[arg]) == (z > 0)).
 * dx (n_src x feat, pitch lddx) is zeroed by the call.  Accumulation uses fp32 red.global.add,
 * like DGL's own backward (order not fixed; see plagnn_spmm_max_bwd_gather for the ordered twin). */
int plagnn_spmm_max_bwd(const float* dz, int64_t lddz, const int32_t* arg, int64_t ldarg,
                        const float* z, int64_t ldz, int64_t num_rows, int64_t feat,
                        float* dx, int64_t n_src, int64_t lddx, plagnn_stream_t stream);

/* Ordered (run-to-run bit-stable) twin of the above: per SOURCE row u, walk the out-edge CSR and
 * add dz[v,f] for the edges whose arg[v,f] == u, in out-edge order.  Costs a full pass over E'.
 * Needs a graph without duplicate (u,v) edges (a duplicate would be counted once per copy); the PPI
 * adjacency is simple (built from a set of pairs, code/data_preprocess.py:83-108). */
int plagnn_spmm_max_bwd_gather(const int32_t* out_indptr, const int32_t* out_indices, const void* out_plan,
                               const int64_t* out_plan_counts /* host[3] */, int64_t n_src,
                               const float* dz, int64_t lddz, const int32_t* arg, int64_t ldarg,
                               const float* z, int64_t ldz, int64_t feat, float* dx, int64_t lddx,
                               void* partial, size_t partial_bytes, plagnn_stream_t stream);

/* sum family (copy_u / u_mul_e + sum, optional per-destination scale = mean / right norm):
 *   out[v,:] = act( scale[v] * sum_{e in in(v)} w[eids[e]] * x[src(e),:] + bias ) [* dropout mask]
 * eids/w/scale/bias may be NULL; w without eids is read in CSR order (w[e]: weights permuted once at set-up, no
 * indirection in the kernel).  dropout_p == 0 disables dropout (the reference never applies
 * any: code/model.py:11 accepts and ignores `dropout`).  The same call on the out-edge CSR is the
 * exact transpose used by backward. */
int plagnn_spmm_sum(const int32_t* indptr, const int32_t* indices, const int32_t* eids, const void* plan,
                    const int64_t* plan_counts /* host[3] */,
                    int64_t num_rows, const float* w, const float* scale,
                    const float* x, int64_t ldx, int64_t feat,
                    const float* bias, int act, float slope, float dropout_p, uint64_t dropout_seed,
                    float* out, int64_t ldo, void* partial, size_t partial_bytes, plagnn_stream_t stream);
/* Row-range variant (multi-GPU pipelining: rows [row_begin,row_end) of the plan are aggregated while the exchange of
 * the previous chunk is in flight).  plagnn_spmm_plan_range (set-up call, synchronises) turns a row range into the
 * host[4] = {item_begin, item_end, hub_begin, hub_end} handle that plagnn_spmm_sum_rows takes. */
int plagnn_spmm_plan_range(const void* plan, int64_t num_rows, int64_t row_begin, int64_t row_end,
                           int64_t* host_range /* host[4] */, plagnn_stream_t stream);
int plagnn_spmm_sum_rows(const int32_t* indptr, const int32_t* indices, const int32_t* eids, const void* plan,
                         const int64_t* plan_counts /* host[3] */, const int64_t* row_range /* host[4] */,
                         int64_t num_rows, const float* w, const float* scale,
                         const float* x, int64_t ldx, int64_t feat, const float* bias, int act, float slope,
                         float* out, int64_t ldo, void* partial, size_t partial_bytes, plagnn_stream_t stream);
/* Source-slab passes of the same reducers for NARROW rows (feat <= 64: the column slice of one GPU in the feature partition of
 * BASELINE.json configs[3]).  The caller splits the in-edges by SOURCE range into several CSR structures (one per slab of the
 * feature matrix small enough to stay in L2) and calls once per slab, in slab order: each call continues from `prev*` (the
 * result so far; NULL for the first slab; may alias the outputs) and the call with last != 0 applies the epilogue (sum) or the
 * "row without in-edges = 0" rule (max).  Max reducer: on equal values the earlier slab wins, inside a slab the first in-edge. */
int plagnn_spmm_sum_slab(const int32_t* indptr, const int32_t* indices, const int32_t* eids, const void* plan,
                         const int64_t* plan_counts /* host[3] */, int64_t num_rows, const float* w, const float* scale,
                         const float* x, int64_t ldx, int64_t feat, const float* bias, int act, float slope,
                         const float* prev, int64_t ldprev, int last, float* out, int64_t ldo,
                         void* partial, size_t partial_bytes, plagnn_stream_t stream);
int plagnn_spmm_max_slab(const int32_t* indptr, const int32_t* indices, const void* plan, const int64_t* plan_counts /* host[3] */,
                         int64_t num_rows, const float* x, int64_t ldx, int64_t feat,
                         const float* prev_val, const int32_t* prev_arg, int64_t ldprev, int last,
                         float* out, int32_t* arg, int64_t ldo, void* partial, size_t partial_bytes, plagnn_stream_t stream);
int plagnn_spmm_max_fwd_rows(const int32_t* indptr, const int32_t* indices, const void* plan, const int64_t* plan_counts /* host[3] */,
                             const int64_t* row_range /* host[4] */, int64_t num_rows, const float* x, int64_t ldx, int64_t feat,
                             float* out, int32_t* arg, int64_t ldo, void* partial, size_t partial_bytes, plagnn_stream_t stream);
/* backward companion of the dropout epilogue: grad[r,c] *= keep(r,c)/(1-p) with the same counter-based mask */
int plagnn_dropout_scale(float* grad, int64_t rows, int64_t feat, int64_t ld, float dropout_p,
                         uint64_t dropout_seed, plagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3  dense contraction — replaces the cuBLAS SGEMMs behind nn.Linear in
 *     code/model.py:16-17,20-28 (fc_pool / fc_self / fc_neigh inside SAGEConv, liner1, liner2)
 *     and their autograd backward.
 *
 *   C[m x n] = epilogue( sum_p  op(A_p)[m x k_p] * op(B_p)[k_p x n] )
 *   a_trans = 0: A_p stored [m x k_p] (k contiguous);  1: stored [k_p x m] (m contiguous)
 *   b_trans = 0: B_p stored [n x k_p] (k contiguous, an nn.Linear weight: C = A*W^T);
 *             1: stored [k_p x n] (n contiguous)
 *   epilogue: v = acc (+ bias[n]);  v = act(v);  if gate: v *= act'(gate[m,n]) with the derivative
 *   written in terms of the saved forward OUTPUT (relu/leaky: sign test, sigmoid: y(1-y)).
 *   Up to PLAGNN_GEMM_MAX_PAIRS pairs accumulate into one tile (fc_self + fc_neigh in one pass).
 *   backend: AUTO picks PLAGNN_GEMM_TMA (TMA-fed tcgen05 on CTA pairs, 3xTF32 split, fp32-level accuracy) when the
 *   operands have 16-byte aligned rows, else TCGEN05 (first-generation kernel) / SIMT (exact fp32 FFMA); products with
 *   n <= 16 or a contraction of at most 32 (the 12-class head, code/model.py:17,28) go to PLAGNN_GEMM_NARROW, streaming
 *   FFMA kernels without tile padding.  Within the TMA backend a directly written product (no split-K) with more 256 x 128
 *   tiles than CTA pairs may run on the double-buffered kernel (accumulators of two tiles in tensor memory, read-out of one
 *   overlapping the MMAs of the next); the choice is made per product by a cost model and does not change the result (the
 *   two kernels are bit-identical).
 *   Parity mode (default; environment PLAGNN_GEMM_PARITY, read per call: 0 = off, 1 = only products with a bias / activation
 *   epilogue, 2 / unset = all): such tall products accumulate each tile in two chains (the two halves of the contraction, in the
 *   two halves of tensor memory) and the long-K weight-gradient products in chains of 24 k-blocks — the tensor core truncates
 *   when it adds into its accumulator, so shorter chains carry less bias; this is what holds the gradients of the full-size
 *   epoch (N = 24 041) within 1e-5 of the fp32 oracle.  Results of the two modes differ at the 1e-6 level.
 *   Output rows that are 16-byte aligned (ldc % 4 == 0, c 16-byte aligned) leave through TMA stores, which clip at the
 *   output's extent in 16-byte units: when n % 4 != 0 (and then necessarily ldc > n) the columns [n, roundup4(n)) of each row
 *   — padding inside the row pitch — may be overwritten with unspecified values.
 * ---------------------------------------------------------------------------------------- */
#define PLAGNN_GEMM_MAX_PAIRS 2
typedef struct {
    const float* a; int64_t lda; int32_t a_trans;
    const float* b; int64_t ldb; int32_t b_trans;
    int64_t k;
} plagnn_gemm_pair;

size_t plagnn_gemm_workspace_bytes(int64_t m, int64_t n, int64_t k_total);
int plagnn_gemm(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs /* host */,
                const float* bias, int act, float slope,
                const float* gate, int64_t ldg, int gate_act,
                float* c, int64_t ldc, void* workspace, size_t workspace_bytes,
                int backend, plagnn_stream_t stream);

/* Weight gradient and bias gradient of a linear layer in one pass (autograd of nn.Linear / SAGEConv fc_*,
 * code/train.py:204):  dw[m x n] = dz^T x,  db[m] = sum over rows of dz;  dz stored [k x m], x stored [k x n] (k = nodes).
 * On the TMA backend db is one more output column of the same split-K product (a B column that reads as 1.0);
 * m <= 16 (the class dimension) takes a streaming FFMA kernel with the same extra column; other backends run
 * plagnn_gemm + plagnn_colsum.  workspace >= plagnn_gemm_wgrad_bias_workspace_bytes(m, n, k). */
size_t plagnn_gemm_wgrad_bias_workspace_bytes(int64_t m, int64_t n, int64_t k);
int plagnn_gemm_wgrad_bias(int64_t m, int64_t n, const float* dz, int64_t lddz, const float* x, int64_t ldx, int64_t k,
                           float* dw, int64_t lddw, float* db, void* workspace, size_t workspace_bytes,
                           plagnn_stream_t stream);

/* column sums: out[j] = sum_i x[i,j]   (bias gradients). workspace >= plagnn_colsum_workspace_bytes. */
size_t plagnn_colsum_workspace_bytes(int64_t rows, int64_t cols);
int plagnn_colsum(const float* x, int64_t rows, int64_t cols, int64_t ldx, float* out,
                  void* workspace, size_t workspace_bytes, plagnn_stream_t stream);

/* dz = dy * act'(y) * row_scale[r], the derivative written through the saved forward output y
 * (torch.sigmoid after liner2, code/model.py:29, is the one activation not folded into a GEMM epilogue).
 * y == NULL skips the activation factor, row_scale == NULL the per-row factor. */
int plagnn_act_backward(const float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t rows, int64_t cols,
                        int act, float slope, const float* row_scale, float* dz, int64_t lddz,
                        plagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Alteration scoring (post-training) — replaces the numpy loops of
 *     code/main.py:15-29   scaling(logit_mat)
 *     code/main.py:32-48   mat_merge(): mean over the 100 runs of scaling(logits)
 *     code/main.py:80-84   diff = (inter - normal) / normal; argsort; reverse
 * scaling: out = scaling(x) in x's precision (is_f64 = 0: float, 1: double) and/or acc += (double)scaling(x);
 *   bit-identical to numpy (row sums in numpy's pairwise order; cols <= 128).
 * alteration_rank: diff[r,c] = (inter - normal) / normal and order[i] = flat index r*cols + c of the i-th largest
 *   score: NaN first, ties by descending index (= numpy stable argsort reversed).
 * workspace >= plagnn_scoring_workspace_bytes(rows, cols) for both.
 * ---------------------------------------------------------------------------------------- */
size_t plagnn_scoring_workspace_bytes(int64_t rows, int64_t cols);
int plagnn_scaling(const void* x, int is_f64, int64_t ldx, int64_t rows, int64_t cols, void* out, int64_t ldo,
                   double* acc, int64_t ldacc, void* workspace, size_t workspace_bytes, plagnn_stream_t stream);
int plagnn_divide_f64(double* x, int64_t ld, int64_t rows, int64_t cols, double divisor, plagnn_stream_t stream);
int plagnn_alteration_rank(const double* normal, int64_t ldn, const double* inter, int64_t ldi, int64_t rows,
                           int64_t cols, double* diff, int64_t ldd, int64_t* order /* rows*cols */,
                           void* workspace, size_t workspace_bytes, plagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4  loss and optimiser — replace
 *     code/train.py:89-108,203  multi_loss(logits[train_index], labels[train_index], i_weight)
 *     code/train.py:180,205     torch.optim.Adam(model.parameters(), lr).step()
 *
 * bce: p = probabilities (the model output is already sigmoid, code/model.py:29), rows selected by
 *   `index` (int64, NULL = all rows).  Computes the reference's class-weighted, clamp(1e-9,10)
 *   BCE exactly in fp32 (p, 1-p, clamp, log) and writes
 *     loss[0]            = sum_i -(1/R) sum_r [...]            (fp32)
 *     dprob[N x C]       = d loss / d p  on the selected rows, 0 elsewhere (if dprob != NULL)
 *   grad_scale multiplies dprob (upstream gradient of the scalar loss).  Negative indices count from the end, as in
 *   torch; an index outside [-N, N) (IndexError in the reference) makes loss[0] NaN.
 * ---------------------------------------------------------------------------------------- */
size_t plagnn_bce_workspace_bytes(int64_t num_index, int64_t classes);
int plagnn_bce_weighted(const float* prob, int64_t ldp, const float* target, int64_t ldt,
                        const int64_t* index, int64_t num_index, int64_t num_rows, int64_t classes,
                        const float* class_weight       /* device fp32[classes]: (float) w_i        */,
                        const float* class_weight_plus1 /* device fp32[classes]: (float)(w_i + 1.0) */,
                        float grad_scale, float* loss, float* dprob, int64_t lddp,
                        void* workspace, size_t workspace_bytes, plagnn_stream_t stream);

/* dst[rows x cols] = src * (*scalar), scalar on the device: the product of the saved loss gradient with autograd's incoming
 * gradient of the scalar loss (code/train.py:204 `train_loss.backward()`), without a host read of that scalar. */
int plagnn_scale_by_device_scalar(const float* src, int64_t lds, int64_t rows, int64_t cols, const float* scalar /* device */,
                                  float* dst, int64_t ldd, plagnn_stream_t stream);

/* multi-tensor Adam (betas, eps, no weight decay, no amsgrad — torch.optim.Adam defaults).
 * tensors: DEVICE array of `count` descriptors.  bias_correction1 = 1-beta1^t and
 * bias_correction2_sqrt = sqrt(1-beta2^t) are computed by the host in double, as torch does. */
typedef struct {
    float* param; const float* grad; float* exp_avg; float* exp_avg_sq; int64_t numel;
} plagnn_adam_tensor;
int plagnn_adam_multi(const plagnn_adam_tensor* tensors /* device */, int32_t count, int64_t max_numel,
                      double lr, double beta1, double beta2, double eps,
                      double bias_correction1, double bias_correction2_sqrt, plagnn_stream_t stream);
/* The same step with the step count held on the DEVICE, so that the call can sit inside a CUDA graph that is replayed once
 * per epoch (code/train.py:195-205 runs 200 epochs of fixed shape per model): `step_count` (device int64, starts at 0) is
 * incremented by the call; the bias corrections 1-beta1^t, sqrt(1-beta2^t) are formed from it on the device in double and
 * rounded once to fp32, like the host version.  `scalars`: device float[4] scratch owned by the caller. */
int plagnn_adam_multi_devstep(const plagnn_adam_tensor* tensors /* device */, int32_t count, int64_t max_numel,
                              double lr, double beta1, double beta2, double eps,
                              int64_t* step_count /* device */, float* scalars /* device float[4] */, plagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * next-1: label decision on device — replaces the Python row loop of
 *     code/train.py:19-40  protein_loc_correction(loc_proba, alpha)
 * pred[N x C] (fp32 0/1). */
size_t plagnn_loc_correction_workspace_bytes(int64_t classes);
int plagnn_loc_correction(const float* prob, int64_t ldp, int64_t num_rows, int64_t classes, float alpha,
                          float* pred, int64_t ldpred, void* workspace, size_t workspace_bytes,
                          plagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Whole-network calls — replace, as a unit,  code/model.py:19-31  GNN32.forward(g, in_feat)
 * (3x SAGEConv('pool') + leaky_relu, liner1 + leaky_relu, liner2 + sigmoid) and the autograd
 * backward that  code/train.py:204  runs through it.  One C call per direction; every
 * intermediate lives in a caller-owned arena (plagnn_gnn32_arena_bytes), so an epoch allocates
 * nothing.  The arena must be zero-filled once after allocation and must not be touched between
 * a forward and its backward.
 *   params / grads: HOST arrays of 19 DEVICE pointers to contiguous tensors, in the order
 *     conv{1,2,3}: fc_pool.weight [F x F], fc_pool.bias [F], fc_self.weight [O x F],
 *                  fc_neigh.weight [O x F], bias [O];   liner1.weight, liner1.bias, liner2.weight, liner2.bias
 *   x: N x in_feats, 16-byte aligned rows.  prob: N x classes probabilities (saved by the caller for backward).
 *   dx may be NULL (the reference's features do not require a gradient).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t num_nodes;
    int32_t in_feats, h1, h2, h3, h4, classes;
    const int32_t* indptr;    /* in-edge CSR of the graph (plagnn_csr_build, key = destination) */
    const int32_t* indices;
    const void* plan;         /* plagnn_spmm_plan_build */
    int64_t plan_counts[3];
} plagnn_gnn32_shape;

size_t plagnn_gnn32_arena_bytes(const plagnn_gnn32_shape* shape);
int plagnn_gnn32_forward(const plagnn_gnn32_shape* shape, const float* x, int64_t ldx,
                         const float* const* params /* host[19] */, void* arena, size_t arena_bytes,
                         float* prob, int64_t ldprob, plagnn_stream_t stream);
int plagnn_gnn32_backward(const plagnn_gnn32_shape* shape, const float* x, int64_t ldx,
                          const float* const* params /* host[19] */, void* arena, size_t arena_bytes,
                          const float* prob, int64_t ldprob, const float* dprob, int64_t lddprob,
                          float* const* grads /* host[19] */, float* dx, int64_t lddx, plagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Preprocessing (offline stage) — replaces the Python loops / dense numpy passes of
 *     code/data_preprocess.py:175-214   edge_clustering_coefficients(ppi_net, epsilon=0)
 *     code/data_preprocess.py:217-257   modify_network_topology(ppi_net, pcc_nor, pcc_inter, thr)
 * ecc: input = the PPI matrix as CSR with strictly ascending columns per row (what ppi_net.tocsr() gives for a simple
 *   0/1 matrix).  For every stored entry (i, j) with j > i, in row-major order (rank u): COO entries 2u = (i, j, v),
 *   2u+1 = (j, i, v) with v = |N(i) & N(j)| / (min(deg i, deg j) - 1), or epsilon when the denominator is 0 — the order
 *   and the float64 values of the reference's lists.  *n_out (device) = number of COO entries (2 x #upper entries);
 *   capacity = allocated entries (2*nnz always suffices).  *status (device, int32): 0 ok, bit 0 = a row is unsorted or
 *   holds a duplicate, bit 1 = column id out of range, bit 2 = capacity too small.
 * diff_moments: mean_std[0] = mean, mean_std[1] = population standard deviation of (inter - normal) over rows x cols
 *   (np.mean / np.std of code/data_preprocess.py:243-244; two passes, fixed summation tree: equal to numpy's pairwise
 *   sums to rounding, not bit for bit).
 * adj_bitmask: mask[r * words_per_row + c / 32] bit (c % 32) = 1 for every COO entry (r, c); the mask is cleared first.
 * rewire: new_mask = mask with (r, c) cleared where inter - normal < l_threshold and set where inter - normal >
 *   r_threshold (code/data_preprocess.py:250-253; the differences and comparisons are exact, so the result is
 *   bit-identical for given thresholds); rowptr[N+1] = CSR row pointer of new_mask, *n_out (device) = its entry count.
 * bitmask_to_coo: the set bits of mask in row-major order (= coo_matrix(dense), code/data_preprocess.py:255).
 * ---------------------------------------------------------------------------------------- */
/* pearson: out[rows x rows] (float64) = np.corrcoef(expr[rows x samples]) with the diagonal and every NaN (constant rows) set
 *   to 0 — what code/data_preprocess.py:165-170 (construct_gcn_matrix) leaves in memory before coo_matrix().  numpy's steps in
 *   numpy's order; the three-term products are FMA chains in sample order (numpy: a BLAS call), test bar 4 ulp. */
size_t plagnn_pearson_workspace_bytes(int64_t rows, int64_t samples);
int plagnn_pearson(const double* expr, int64_t ldx, int64_t rows, int64_t samples, double* out, int64_t ldo,
                   void* workspace, size_t workspace_bytes, plagnn_stream_t stream);
size_t plagnn_ecc_workspace_bytes(int64_t num_nodes);
int plagnn_ecc(const int32_t* indptr, const int32_t* indices, int64_t num_nodes, int64_t nnz, double epsilon,
               int32_t* ecc_row, int32_t* ecc_col, double* ecc_data, int64_t capacity, int64_t* n_out /* device */,
               int32_t* status /* device */, void* workspace, size_t workspace_bytes, plagnn_stream_t stream);
size_t plagnn_diff_moments_workspace_bytes(void);
int plagnn_diff_moments(const double* normal, int64_t ldn, const double* inter, int64_t ldi, int64_t rows, int64_t cols,
                        double* mean_std /* device[2] */, void* workspace, size_t workspace_bytes, plagnn_stream_t stream);
int plagnn_adj_bitmask(const int32_t* row, const int32_t* col, int64_t nnz, int64_t num_nodes, uint32_t* mask,
                       int64_t words_per_row, int32_t* status /* device */, plagnn_stream_t stream);
size_t plagnn_rewire_workspace_bytes(int64_t num_nodes);
int plagnn_rewire(const double* normal, int64_t ldn, const double* inter, int64_t ldi, int64_t num_nodes,
                  const uint32_t* mask, uint32_t* new_mask, int64_t words_per_row, double l_threshold, double r_threshold,
                  int32_t* rowptr /* N+1 */, int64_t* n_out /* device */, void* workspace, size_t workspace_bytes,
                  plagnn_stream_t stream);
int plagnn_bitmask_to_coo(const uint32_t* mask, int64_t words_per_row, int64_t num_nodes, const int32_t* rowptr,
                          int32_t* out_row, int32_t* out_col, plagnn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Exchange steps of the partitioned aggregation (BASELINE.json configs[3]).
 * The reference is single-process (code/main_normal.py:30,66: one `-d` device), so there is no reference
 * call site to replace: these wrap NCCL for the two ways the scaled synthetic graph is split over GPUs.
 *   row partition     : plagnn_nccl_allgather_rows (forward: projected rows of every rank),
 *                       plagnn_nccl_reducescatter_rows (backward: partial source gradients);
 *   feature partition : plagnn_cols_pack -> plagnn_nccl_alltoall_blocks -> (aggregate all rows x my columns)
 *                       -> plagnn_nccl_alltoall_blocks -> plagnn_cols_unpack;
 *   both              : plagnn_nccl_allreduce of the weight gradients, once per step.
 * `comm` is an ncclComm_t (the caller's own, or one made by plagnn_nccl_comm_init from an id that rank 0 obtained with
 * plagnn_nccl_get_unique_id and distributed by any means).  NCCL is resolved at run time (dlopen of the libnccl.so.2 already
 * in the process, else PLAGNN_NCCL_LIB); without it these return PLAGNN_ERR_UNSUPPORTED.  Calls enqueue on `stream`.
 * Matrices are exchanged with their row pitch (count = rows * pitch floats), so padded buffers go as they are.
 * ---------------------------------------------------------------------------------------- */
typedef void* plagnn_nccl_comm_t; /* an ncclComm_t */
#define PLAGNN_NCCL_UNIQUE_ID_BYTES 128
int plagnn_nccl_available(void);
int plagnn_nccl_get_unique_id(void* id_out /* host, 128 bytes */);
int plagnn_nccl_comm_init(const void* id /* host, 128 bytes */, int rank, int world,
                          int max_ctas /* > 0: cap NCCL's CTAs (ncclConfig_t.maxCTAs) so the exchange leaves the SMs to the aggregation */,
                          plagnn_nccl_comm_t* comm_out);
int plagnn_nccl_comm_destroy(plagnn_nccl_comm_t comm);
/* recv[world * rows x pitch] = concatenation over ranks of send[rows x pitch] */
int plagnn_nccl_allgather_rows(const float* send, float* recv, int64_t rows, int64_t pitch, plagnn_nccl_comm_t comm,
                               plagnn_stream_t stream);
/* recv[rows x pitch] = sum over ranks of their send[rank * rows ... (rank + 1) * rows) */
int plagnn_nccl_reducescatter_rows(const float* send, float* recv, int64_t rows, int64_t pitch, plagnn_nccl_comm_t comm,
                                   plagnn_stream_t stream);
int plagnn_nccl_allreduce(float* buf /* in place */, int64_t count, plagnn_nccl_comm_t comm, plagnn_stream_t stream);
/* block q of send goes to rank q, block q of recv comes from rank q (grouped send / recv) */
int plagnn_nccl_alltoall_blocks(const float* send, float* recv, int64_t block_elems, int world, plagnn_nccl_comm_t comm,
                                plagnn_stream_t stream);
/* x[rows x feat] (pitch ldx) <-> blocks[world][rows][feat / world] (block q = columns [q * feat/world, (q+1) * feat/world));
 * feat % (4 * world) == 0 */
int plagnn_cols_pack(const float* x, int64_t ldx, int64_t rows, int64_t feat, int world, float* blocks, plagnn_stream_t stream);
int plagnn_cols_unpack(const float* blocks, int64_t rows, int64_t feat, int world, float* x, int64_t ldx, plagnn_stream_t stream);

/* Peer-memory exchange for the feature partition (one process per GPU, NVLink / NVSwitch): the all-to-all around an aggregation
 * as one kernel that stores every column block straight into its owner's window, then a flag per peer; no NCCL, no pack /
 * unpack pass.  Set-up: every rank creates a window (device memory allocated by the library, because it has to be exported
 * with cudaIpcGetMemHandle), the 64-byte handles are exchanged by any means, plagnn_p2p_attach opens the peers' windows.
 *   plagnn_p2p_send mode 0: x[rows x feat] (my rows, all columns) -> every peer q gets my rows of ITS columns at
 *                           window[offset + ((rank * rows + r) * feat/world + j)]   (window = all rows x my columns)
 *                   mode 1: x_col[world * rows x feat/world] (all rows, my columns) -> peer q gets its rows of MY columns at
 *                           window[offset + (r * feat + rank * feat/world + j)]     (window = my rows x all columns)
 *   plagnn_p2p_wait: enqueues a one-warp kernel that returns when every peer has published `seq` (sequence numbers grow by
 *                    one per exchange, the same on every rank); it gives up after ~4 s and records `seq` (plagnn_p2p_error).
 * The caller alternates between two window offsets so that a peer's next exchange never lands on data still being read. */
typedef void* plagnn_p2p_t;
#define PLAGNN_P2P_HANDLE_BYTES 64
int plagnn_p2p_create(size_t bytes, int rank, int world, void* handle_out /* host, 64 bytes */, plagnn_p2p_t* out);
int plagnn_p2p_attach(plagnn_p2p_t p2p, const void* all_handles /* host, world x 64 bytes, rank order */);
void* plagnn_p2p_window(plagnn_p2p_t p2p);        /* device pointer of the local window's data */
long long plagnn_p2p_error(plagnn_p2p_t p2p);     /* 0, or the sequence number a wait gave up on (synchronises) */
int plagnn_p2p_destroy(plagnn_p2p_t p2p);
int plagnn_p2p_send(plagnn_p2p_t p2p, const float* src, int64_t lds, int64_t rows, int64_t feat, int mode,
                    size_t dst_offset_bytes, long long seq, plagnn_stream_t stream);
/* one part of a send: source rows [row_begin, row_end) only (mode 0: a chunk of my rows; mode 1: rows of the all-rows matrix),
 * so that the producer's next chunk is computed while this one crosses NVLink.  publish: 0 = more parts follow on this stream,
 * 1 = flag every peer, 2 = flag only the owner of these rows (mode 1, range inside one owner's block). */
int plagnn_p2p_send_part(plagnn_p2p_t p2p, const float* src, int64_t lds, int64_t rows, int64_t feat, int mode,
                         int64_t row_begin, int64_t row_end, int publish, size_t dst_offset_bytes, long long seq,
                         plagnn_stream_t stream);
int plagnn_p2p_wait(plagnn_p2p_t p2p, long long seq, plagnn_stream_t stream);

/* small utilities used by the host layer */
int plagnn_pad_copy(const float* src, int64_t rows, int64_t cols, int64_t lds, float* dst, int64_t ldd,
                    plagnn_stream_t stream); /* dst[:, :cols] = src; dst[:, cols:ldd] = 0 */
int plagnn_transpose(const float* src, int64_t rows, int64_t cols, int64_t lds, float* dst, int64_t ldd,
                     plagnn_stream_t stream); /* dst[cols x rows] = src^T */

#ifdef __cplusplus
}
#endif
#endif /* PLAGNN_H_ */
