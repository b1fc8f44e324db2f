// K4 — loss, optimiser and the small HBM-bound helpers around the aggregation / GEMM kernels.
//
//   plagnn_bce_weighted   <- code/train.py:89-108,203   multi_loss(logits[train_index], labels[train_index], i_weight)
//   plagnn_adam_multi     <- code/train.py:180,205      torch.optim.Adam(model.parameters(), lr).step()
//   plagnn_loc_correction <- code/train.py:19-40        protein_loc_correction (label decision)
//   plagnn_colsum         <- bias gradients of nn.Linear / SAGEConv.bias (autograd at code/train.py:204)
//
// The loss keeps the reference's fp32 operator order (p, 1-p, clamp(1e-9,10), log, *w, /(w+1), *2) and the
// closed-interval clamp gradient, because saturated sigmoid outputs make the "nicer" with-logits form differ.
#include "common.cuh"

namespace plagnn {

constexpr int BCE_THREADS = 256;
constexpr int BCE_MAX_CLASSES = 64;

__global__ void __launch_bounds__(BCE_THREADS)
bce_kernel(const float* __restrict__ prob, int64_t ldp, const float* __restrict__ target, int64_t ldt,
           const int64_t* __restrict__ index, int64_t num_index, int64_t num_rows, int classes,
           const float* __restrict__ cw, const float* __restrict__ cwp1, float grad_scale,
           double* __restrict__ block_part, float* __restrict__ dprob, int64_t lddp) {
    __shared__ double red[BCE_THREADS / 32][BCE_MAX_CLASSES];
    pdl_enter();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = (int64_t)blockIdx.x * BCE_THREADS + threadIdx.x;
    int64_t row = -1;
    int bad = 0;
    if (i < num_index) {
        row = index ? index[i] : i;
        if (row < 0) row += num_rows;            // python-style negative index
        if (row < 0 || row >= num_rows) { bad = 1; row = -1; }
    }
    const float inv_r = -1.0f / (float)num_index;   // grad of  -(sum)/R  w.r.t. sum
    constexpr int CU = 4;                            // classes whose loads are in flight together
    for (int c0 = 0; c0 < classes; c0 += CU) {
        float pv[CU], tv[CU];
#pragma unroll
        for (int u = 0; u < CU; ++u) {
            const bool on = row >= 0 && c0 + u < classes;
            pv[u] = on ? prob[row * ldp + c0 + u] : 0.5f;
            tv[u] = on ? target[row * ldt + c0 + u] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < CU; ++u) {
            const int c = c0 + u;
            if (c >= classes) break;
            double term = 0.0;
            if (row >= 0) {
                const float p = pv[u], t = tv[u];
                const float w = cw[c], wp1 = cwp1[c];
                const float q = 1.0f - p;
                const float pc = fminf(fmaxf(p, 1e-9f), 10.0f);
                const float qc = fminf(fmaxf(q, 1e-9f), 10.0f);
                const float pos = t * logf(pc) * w;
                const float neg = (1.0f - t) * logf(qc);
                term = (double)((pos + neg) / wp1 * 2.0f);
                if (dprob) {
                    const float g2 = (inv_r * grad_scale * 2.0f) / wp1;
                    float g = 0.f;
                    if (p >= 1e-9f && p <= 10.0f) g += ((g2 * w) * t) / pc;
                    if (q >= 1e-9f && q <= 10.0f) g -= (g2 * (1.0f - t)) / qc;
                    atomicAdd(dprob + row * lddp + c, g);
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) term += __shfl_xor_sync(0xffffffffu, term, d);
            if (lane == 0) red[warp][c] = term;
        }
    }
    // an index outside [-num_rows, num_rows) (torch raises IndexError there) poisons the loss: NaN, never a silent skip
    const int any_bad = __syncthreads_or(bad);
    if (threadIdx.x < classes) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < BCE_THREADS / 32; ++w) s += red[w][threadIdx.x];
        block_part[(int64_t)blockIdx.x * classes + threadIdx.x] = any_bad ? (double)NAN : s;
    }
}

// dst[r, c] = src ? src[r, c] * (*scalar) : 0 for c < cols: the clear of the loss gradient before bce_kernel's reductions and the
// product with autograd's incoming gradient (a device scalar), both as kernels so that they stay in the launch chain
__global__ void __launch_bounds__(256)
scale_or_zero_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int cols, const float* __restrict__ scalar,
                     float* __restrict__ dst, int64_t ldd) {
    pdl_enter();
    const float s = scalar ? *scalar : 1.f;
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        const int c = (int)(i - r * cols);
        dst[r * ldd + c] = src ? src[r * lds + c] * s : 0.f;
    }
}

__global__ void bce_finalize_kernel(const double* __restrict__ block_part, int num_blocks, int classes,
                                    int64_t num_index, float* __restrict__ loss) {
    __shared__ float per_class[BCE_MAX_CLASSES];
    pdl_enter();
    if (threadIdx.x < classes) {
        double s = 0.0;
        for (int b = 0; b < num_blocks; ++b) s += block_part[(int64_t)b * classes + threadIdx.x];
        // reference: scl_loss = -scl_loss.sum() / len(input)   (fp32 sum, fp32 divide)
        per_class[threadIdx.x] = -((float)s) / (float)num_index;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float l = 0.f;
        for (int c = 0; c < classes; ++c) l += per_class[c];   // loss += scl_loss, in class order
        loss[0] = l;
    }
}

// ---------------------------------------------------------------------------------------------
// step count on the device (graph-replayable Adam): t = ++(*step_count); scalars[0] = lr / (1 - beta1^t), scalars[1] =
// sqrt(1 - beta2^t), both formed in double and rounded once
__global__ void adam_step_scalars_kernel(int64_t* __restrict__ step_count, float* __restrict__ scalars, double lr, double beta1,
                                         double beta2) {
    pdl_enter();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        const int64_t t = *step_count + 1;
        *step_count = t;
        scalars[0] = (float)(lr / (1.0 - pow(beta1, (double)t)));
        scalars[1] = (float)sqrt(1.0 - pow(beta2, (double)t));
    }
}

__global__ void __launch_bounds__(256)
adam_kernel(const plagnn_adam_tensor* __restrict__ tensors, float lerp_w, float beta2, float one_minus_beta2,
            float eps, float step_size, float bc2_sqrt, const float* __restrict__ dev_scalars) {
    pdl_enter();
    if (dev_scalars) {
        step_size = dev_scalars[0];
        bc2_sqrt = dev_scalars[1];
    }
    const plagnn_adam_tensor T = tensors[blockIdx.y];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < T.numel; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = T.grad[i];
        float m = T.exp_avg[i];
        float v = T.exp_avg_sq[i];
        // The update of torch 2.x's single-tensor Adam (lerp form).  torch 1.10, which the reference pins, writes the same
        // mathematics as exp_avg.mul_(beta1).add_(grad, alpha=1-beta1) and multiplies by reciprocals of the scalar divisors:
        // the two differ at the ulp level per step; the tests hold this kernel to 1e-6 against torch.optim.Adam over 6 steps.
        m = m + lerp_w * (g - m);                     // exp_avg.lerp_(grad, 1 - beta1)
        v = v * beta2;                                // exp_avg_sq.mul_(beta2)
        v = v + one_minus_beta2 * (g * g);            //   .addcmul_(grad, grad, value=1 - beta2)
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        T.param[i] = T.param[i] - (step_size * m) / denom;   // param.addcdiv_(exp_avg, denom, value=-step_size)
        T.exp_avg[i] = m;
        T.exp_avg_sq[i] = v;
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int CS_ROWS_PER_CHUNK = 128;
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ldx, double* __restrict__ part) {
    __shared__ double red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS_PER_CHUNK;
    const int64_t r1 = r0 + CS_ROWS_PER_CHUNK < rows ? r0 + CS_ROWS_PER_CHUNK : rows;
    // 16 rows per thread are tree-summed in fp32 (error <= 4 ulp of the partial), everything above that level is
    // accumulated in double: column sums of gradients cancel heavily, and B200's fp64 rate is too low to spend one
    // DADD per element (the first version was fp64-throughput bound: 28 us for 38 MB).
    double s = 0.0;
    if (c < cols) {
        float v[CS_ROWS_PER_CHUNK / 8];
#pragma unroll
        for (int i = 0; i < CS_ROWS_PER_CHUNK / 8; ++i) {      // all loads first (16 in flight per thread)
            const int64_t r = r0 + ty + 8 * i;
            v[i] = r < r1 ? __ldg(x + r * ldx + c) : 0.f;
        }
#pragma unroll
        for (int w = CS_ROWS_PER_CHUNK / 16; w > 0; w >>= 1)
#pragma unroll
            for (int i = 0; i < w; ++i) v[i] += v[i + w];
        s = (double)v[0];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < cols) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        part[(int64_t)blockIdx.y * cols + c] = t;
    }
}
// Same sums (identical grouping, hence identical bits) with 16-byte loads: a lane owns 4 adjacent columns, a warp reads
// 512 contiguous bytes per row.  Used when the rows are 16-byte aligned (the 4-byte version moved 38 MB at 2 TB/s).
__global__ void __launch_bounds__(256)
colsum_partial4_kernel(const float* __restrict__ x, int64_t rows, int cols, int64_t ldx, double* __restrict__ part) {
    __shared__ double red[8][32][4];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + tx) * 4;
    const int64_t r0 = (int64_t)blockIdx.y * CS_ROWS_PER_CHUNK;
    const int64_t r1 = r0 + CS_ROWS_PER_CHUNK < rows ? r0 + CS_ROWS_PER_CHUNK : rows;
    float4 v[CS_ROWS_PER_CHUNK / 8];
    const bool on = c < cols;            // cols is rounded up to the row pitch by the caller's padding: c + 3 < ldx
#pragma unroll
    for (int i = 0; i < CS_ROWS_PER_CHUNK / 8; ++i) {
        const int64_t r = r0 + ty + 8 * i;
        v[i] = (on && r < r1) ? ldg_f4(x + r * ldx + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int w = CS_ROWS_PER_CHUNK / 16; w > 0; w >>= 1)
#pragma unroll
        for (int i = 0; i < w; ++i) {
            v[i].x += v[i + w].x; v[i].y += v[i + w].y; v[i].z += v[i + w].z; v[i].w += v[i + w].w;
        }
    red[ty][tx][0] = (double)v[0].x; red[ty][tx][1] = (double)v[0].y;
    red[ty][tx][2] = (double)v[0].z; red[ty][tx][3] = (double)v[0].w;
    __syncthreads();
    if (ty < 4 && on) {                  // warp ty finishes component ty of every lane's quad
        const int cc = c + ty;
        if (cc < cols) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) t += red[i][tx][ty];
            part[(int64_t)blockIdx.y * cols + cc] = t;
        }
    }
}
// 32 columns x 8 partial-row groups per block: the per-column chain over the row chunks is split 8 ways and
// unrolled, so 8 loads per thread are in flight (the first version walked 188 chunks with one dependent load at a
// time: 29 us of pure L2 latency).
__global__ void __launch_bounds__(256)
colsum_final_kernel(const double* __restrict__ part, int chunks, int cols, float* __restrict__ out) {
    __shared__ double red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    double s = 0.0;
    if (c < cols) {
        int k = ty;
        for (; k + 56 < chunks; k += 64) {
            double v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = part[(int64_t)(k + 8 * j) * cols + c];
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[j];
        }
        for (; k < chunks; k += 8) s += part[(int64_t)k * cols + c];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < cols) {
        double t = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][tx];
        out[c] = (float)t;
    }
}

// ---------------------------------------------------------------------------------------------
// label decision: column min/max, then per-row normalise + threshold
__global__ void __launch_bounds__(256)
colminmax_partial_kernel(const float* __restrict__ p, int64_t ldp, int64_t rows, int classes, float* __restrict__ part) {
    // part[block][2*classes] : mins then maxs
    __shared__ float smin[256], smax[256];
    for (int c = 0; c < classes; ++c) {
        float mn = INFINITY, mx = -INFINITY;
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
            const float v = p[r * ldp + c];
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
        smin[threadIdx.x] = mn;
        smax[threadIdx.x] = mx;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + s]);
                smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + s]);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            part[(int64_t)blockIdx.x * 2 * classes + c] = smin[0];
            part[(int64_t)blockIdx.x * 2 * classes + classes + c] = smax[0];
        }
        __syncthreads();
    }
}
__global__ void loc_decide_kernel(const float* __restrict__ p, int64_t ldp, int64_t rows, int classes, float alpha,
                                  const float* __restrict__ part, int nparts, float* __restrict__ pred, int64_t ldpred) {
    __shared__ float cmin[BCE_MAX_CLASSES], cmax[BCE_MAX_CLASSES];
    if (threadIdx.x < classes) {
        float mn = INFINITY, mx = -INFINITY;
        for (int b = 0; b < nparts; ++b) {
            mn = fminf(mn, part[(int64_t)b * 2 * classes + threadIdx.x]);
            mx = fmaxf(mx, part[(int64_t)b * 2 * classes + classes + threadIdx.x]);
        }
        cmin[threadIdx.x] = mn;
        cmax[threadIdx.x] = mx;
    }
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    float v[BCE_MAX_CLASSES];
    float s = 0.f;
    for (int c = 0; c < classes; ++c) {
        v[c] = (p[r * ldp + c] - cmin[c]) / (cmax[c] - cmin[c]);
        s += v[c];
    }
    float rmax = -INFINITY, rmin = INFINITY;
    for (int c = 0; c < classes; ++c) {
        v[c] = v[c] / s;
        rmax = fmaxf(rmax, v[c]);
        rmin = fminf(rmin, v[c]);
    }
    const float thr = rmax - (rmax - rmin) * alpha;
    for (int c = 0; c < classes; ++c) pred[r * ldpred + c] = v[c] > thr ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pad_copy_kernel(const float* __restrict__ src, int64_t rows, int cols, int64_t lds, float* __restrict__ dst, int64_t ldd) {
    pdl_enter();
    const int64_t total = rows * ldd;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ldd;
        const int c = (int)(i - r * ldd);
        dst[i] = c < cols ? __ldg(src + r * lds + c) : 0.f;
    }
}

__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ src, int64_t rows, int64_t cols, int64_t lds, float* __restrict__ dst, int64_t ldd) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int i = ty; i < 32; i += 8)
        tile[i][tx] = (r0 + i < rows && c0 + tx < cols) ? __ldg(src + (r0 + i) * lds + c0 + tx) : 0.f;
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < cols && r0 + tx < rows) dst[(c0 + i) * ldd + r0 + tx] = tile[tx][i];
}

static inline int capped_grid(int64_t work_items, int threads, int waves) {
    const int64_t want = ceil_div(work_items, threads);
    const int64_t cap = (int64_t)sm_count() * waves;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

}  // namespace plagnn

using namespace plagnn;

extern "C" {

size_t plagnn_bce_workspace_bytes(int64_t num_index, int64_t classes) {
    const int64_t blocks = ceil_div(num_index > 0 ? num_index : 1, BCE_THREADS);
    return align_up((size_t)blocks * classes * sizeof(double), 256) + 256;
}

int plagnn_bce_weighted(const float* prob, int64_t ldp, const float* target, int64_t ldt, const int64_t* index,
                        int64_t num_index, int64_t num_rows, int64_t classes, const float* class_weight,
                        const float* class_weight_plus1, float grad_scale, float* loss, float* dprob, int64_t lddp,
                        void* workspace, size_t workspace_bytes, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!prob || !target || !class_weight || !class_weight_plus1 || !loss || num_index <= 0 || num_rows <= 0 ||
        classes <= 0 || classes > BCE_MAX_CLASSES || ldp < classes || ldt < classes || (dprob && lddp < classes))
        return fail(PLAGNN_ERR_ARG, "bce_weighted", "bad arguments");
    const size_t need = plagnn_bce_workspace_bytes(num_index, classes);
    if (!workspace || workspace_bytes < need) return fail(PLAGNN_ERR_WORKSPACE, "bce_weighted", "workspace too small");
    ProfileScope prof("bce_weighted", num_index, classes, 0, stream);
    const int blocks = (int)ceil_div(num_index, BCE_THREADS);
    double* part = (double*)workspace;
    if (dprob)
        launch_pdl(scale_or_zero_kernel, dim3((unsigned)capped_grid(num_rows * classes, 256, 8)), dim3(256), 0, st, (const float*)nullptr,
                   (int64_t)0, num_rows, (int)classes, (const float*)nullptr, dprob, lddp);
    launch_pdl(bce_kernel, dim3(blocks), dim3(BCE_THREADS), 0, st, prob, ldp, target, ldt, index, num_index, num_rows, (int)classes,
                                                class_weight, class_weight_plus1, grad_scale, part, dprob, lddp);
    launch_pdl(bce_finalize_kernel, dim3(1), dim3(BCE_MAX_CLASSES), 0, st, part, blocks, (int)classes, num_index, loss);
    return check_launch("bce_weighted", dprob ? 3 : 2);
}

int plagnn_scale_by_device_scalar(const float* src, int64_t lds, int64_t rows, int64_t cols, const float* scalar, float* dst,
                                  int64_t ldd, plagnn_stream_t stream) {
    if (!src || !dst || !scalar || rows <= 0 || cols <= 0 || lds < cols || ldd < cols)
        return fail(PLAGNN_ERR_ARG, "scale_by_device_scalar", "bad arguments");
    launch_pdl(scale_or_zero_kernel, dim3((unsigned)capped_grid(rows * cols, 256, 8)), dim3(256), 0, (cudaStream_t)stream, src, lds, rows,
               (int)cols, scalar, dst, ldd);
    return check_launch("scale_by_device_scalar");
}

int plagnn_adam_multi(const plagnn_adam_tensor* tensors, int32_t count, int64_t max_numel, double lr, double beta1,
                      double beta2, double eps, double bias_correction1, double bias_correction2_sqrt,
                      plagnn_stream_t stream) {
    if (!tensors || count <= 0 || max_numel <= 0 || bias_correction1 <= 0.0 || bias_correction2_sqrt <= 0.0)
        return fail(PLAGNN_ERR_ARG, "adam_multi", "bad arguments");
    ProfileScope prof("adam_multi", count, max_numel, 0, stream);
    // scalars are formed in double on the host and rounded once to fp32, as torch does with Python floats
    const float step_size = (float)(lr / bias_correction1);
    const float lerp_w = (float)(1.0 - beta1);
    const float omb2 = (float)(1.0 - beta2);
    dim3 grid((unsigned)capped_grid(max_numel, 256, 4), (unsigned)count);
    launch_pdl(adam_kernel, grid, dim3(256), 0, (cudaStream_t)stream, tensors, lerp_w, (float)beta2, omb2, (float)eps, step_size,
                                                        (float)bias_correction2_sqrt, (const float*)nullptr);
    return check_launch("adam_multi");
}

int plagnn_adam_multi_devstep(const plagnn_adam_tensor* tensors, int32_t count, int64_t max_numel, double lr, double beta1,
                              double beta2, double eps, int64_t* step_count, float* scalars, plagnn_stream_t stream) {
    if (!tensors || count <= 0 || max_numel <= 0 || !step_count || !scalars)
        return fail(PLAGNN_ERR_ARG, "adam_multi_devstep", "bad arguments");
    ProfileScope prof("adam_multi", count, max_numel, 1, stream);
    launch_pdl(adam_step_scalars_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, step_count, scalars, lr, beta1, beta2);
    dim3 grid((unsigned)capped_grid(max_numel, 256, 4), (unsigned)count);
    launch_pdl(adam_kernel, grid, dim3(256), 0, (cudaStream_t)stream, tensors, (float)(1.0 - beta1), (float)beta2,
               (float)(1.0 - beta2), (float)eps, 0.f, 1.f, (const float*)scalars);
    return check_launch("adam_multi_devstep", 2);
}

size_t plagnn_colsum_workspace_bytes(int64_t rows, int64_t cols) {
    return align_up((size_t)ceil_div(rows > 0 ? rows : 1, CS_ROWS_PER_CHUNK) * (size_t)cols * sizeof(double), 256);
}

int plagnn_colsum(const float* x, int64_t rows, int64_t cols, int64_t ldx, float* out, void* workspace,
                  size_t workspace_bytes, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!x || !out || rows <= 0 || cols <= 0 || ldx < cols) return fail(PLAGNN_ERR_ARG, "colsum", "bad arguments");
    ProfileScope prof("colsum", rows, cols, 0, stream);
    if (!workspace || workspace_bytes < plagnn_colsum_workspace_bytes(rows, cols))
        return fail(PLAGNN_ERR_WORKSPACE, "colsum", "workspace too small");
    const int chunks = (int)ceil_div(rows, CS_ROWS_PER_CHUNK);
    if (aligned16(x) && (ldx & 3) == 0 && ceil_div(cols, 4) * 4 <= ldx) {
        dim3 grid4((unsigned)ceil_div(cols, 128), (unsigned)chunks);
        colsum_partial4_kernel<<<grid4, 256, 0, st>>>(x, rows, (int)cols, ldx, (double*)workspace);
    } else {
        dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)chunks);
        colsum_partial_kernel<<<grid, 256, 0, st>>>(x, rows, (int)cols, ldx, (double*)workspace);
    }
    colsum_final_kernel<<<(unsigned)ceil_div(cols, 32), 256, 0, st>>>((const double*)workspace, chunks, (int)cols, out);
    return check_launch("colsum", 2);
}

size_t plagnn_loc_correction_workspace_bytes(int64_t classes) {
    return align_up((size_t)sm_count() * 2 * (size_t)(classes > 0 ? classes : 1) * sizeof(float), 256);
}

int plagnn_loc_correction(const float* prob, int64_t ldp, int64_t num_rows, int64_t classes, float alpha, float* pred,
                          int64_t ldpred, void* workspace, size_t workspace_bytes, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!prob || !pred || num_rows <= 0 || classes <= 0 || classes > BCE_MAX_CLASSES || ldp < classes || ldpred < classes)
        return fail(PLAGNN_ERR_ARG, "loc_correction", "bad arguments");
    const int nparts = capped_grid(num_rows, 256, 1);
    if (!workspace || workspace_bytes < (size_t)nparts * 2 * classes * sizeof(float))
        return fail(PLAGNN_ERR_WORKSPACE, "loc_correction", "workspace too small (need 2*classes*sm_count floats)");
    colminmax_partial_kernel<<<nparts, 256, 0, st>>>(prob, ldp, num_rows, (int)classes, (float*)workspace);
    loc_decide_kernel<<<(unsigned)ceil_div(num_rows, 128), 128, 0, st>>>(prob, ldp, num_rows, (int)classes, alpha,
                                                                          (const float*)workspace, nparts, pred, ldpred);
    return check_launch("loc_correction", 2);
}

int plagnn_pad_copy(const float* src, int64_t rows, int64_t cols, int64_t lds, float* dst, int64_t ldd,
                    plagnn_stream_t stream) {
    if (!src || !dst || rows <= 0 || cols <= 0 || lds < cols || ldd < cols) return fail(PLAGNN_ERR_ARG, "pad_copy", "bad arguments");
    ProfileScope prof("pad_copy", rows, cols, 0, stream);
    launch_pdl(pad_copy_kernel, dim3(capped_grid(rows * ldd, 256, 8)), dim3(256), 0, (cudaStream_t)stream, src, rows, (int)cols, lds, dst, ldd);
    return check_launch("pad_copy");
}

int plagnn_transpose(const float* src, int64_t rows, int64_t cols, int64_t lds, float* dst, int64_t ldd,
                     plagnn_stream_t stream) {
    if (!src || !dst || rows <= 0 || cols <= 0 || lds < cols || ldd < rows) return fail(PLAGNN_ERR_ARG, "transpose", "bad arguments");
    ProfileScope prof("transpose", rows, cols, 0, stream);
    dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
    transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, lds, dst, ldd);
    return check_launch("transpose");
}

}  // extern "C"

// dz = dy * act'(y), derivative written through the saved forward output y (sigmoid after liner2,
// code/model.py:29, is the one activation whose gradient is not folded into a GEMM epilogue)
namespace plagnn {
__global__ void __launch_bounds__(256)
act_backward_kernel(const float* __restrict__ dy, int64_t lddy, const float* __restrict__ y, int64_t ldy, int64_t rows,
                    int cols, int act, float slope, const float* __restrict__ row_scale, float* __restrict__ dz,
                    int64_t lddz) {
    pdl_enter();
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        const int c = (int)(i - r * cols);
        float v = dy[r * lddy + c];
        if (y) v *= act_grad_from_output(y[r * ldy + c], act, slope);
        if (row_scale) v *= __ldg(row_scale + r);
        dz[r * lddz + c] = v;
    }
}
}  // namespace plagnn

extern "C" int plagnn_act_backward(const float* dy, int64_t lddy, const float* y, int64_t ldy, int64_t rows, int64_t cols,
                                   int act, float slope, const float* row_scale, float* dz, int64_t lddz,
                                   plagnn_stream_t stream) {
    using namespace plagnn;
    if (!dy || !dz || rows <= 0 || cols <= 0 || lddy < cols || (y && ldy < cols) || lddz < cols)
        return fail(PLAGNN_ERR_ARG, "act_backward", "bad arguments");
    ProfileScope prof("act_backward", rows, cols, 0, stream);
    launch_pdl(act_backward_kernel, dim3(capped_grid(rows * cols, 256, 8)), dim3(256), 0, (cudaStream_t)stream, dy, lddy, y, ldy, rows,
                                                                                          (int)cols, act, slope, row_scale, dz, lddz);
    return check_launch("act_backward");
}
