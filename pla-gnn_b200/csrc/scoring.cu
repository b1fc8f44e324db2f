// Alteration scoring — the post-training step that turns the 100 logit matrices of a condition
// into the ranked mislocalisation list:
//   code/main.py:15-29    scaling(): subtract column min, divide by column max, divide rows by their sum
//   code/main.py:32-48    mat_merge(): mean over the runs of scaling(logits)  (float32 runs accumulated in float64)
//   code/main.py:80-84    diff = (inter - normal) / normal ; argsort ; reverse
// All arithmetic is IEEE-exact per element; the only order-dependent step, the row sum, follows numpy's pairwise
// summation for rows of up to 128 elements (12 here), so results are bit-identical to the reference's numpy code.
// The ranking is a bitonic sort of (order-preserving key, flat index): descending score, NaN first (numpy sorts NaN
// last before the reverse), ties by descending index — what a stable argsort followed by reverse gives (numpy's default
// argsort is not stable, so the reference's own order inside a tie group is an artefact).
#include "common.cuh"
#include <math.h>

namespace plagnn {

constexpr int SC_MAX_COLS = 128;

template <typename T>
__global__ void __launch_bounds__(256)
sc_colminmax_kernel(const T* __restrict__ x, int64_t ldx, int64_t rows, int cols, T* __restrict__ part) {
    // part[block][2*cols]: mins then maxs (min / max are exact, any order)
    __shared__ T smin[256], smax[256];
    for (int c = 0; c < cols; ++c) {
        T mn = (T)INFINITY, mx = (T)-INFINITY;
        for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
            const T v = x[r * ldx + c];
            // numpy's min/max propagate NaN; logits are finite, keep the plain comparisons
            mn = v < mn ? v : mn;
            mx = v > mx ? v : mx;
        }
        smin[threadIdx.x] = mn;
        smax[threadIdx.x] = mx;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) {
                smin[threadIdx.x] = smin[threadIdx.x + s] < smin[threadIdx.x] ? smin[threadIdx.x + s] : smin[threadIdx.x];
                smax[threadIdx.x] = smax[threadIdx.x + s] > smax[threadIdx.x] ? smax[threadIdx.x + s] : smax[threadIdx.x];
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            part[(int64_t)blockIdx.x * 2 * cols + c] = smin[0];
            part[(int64_t)blockIdx.x * 2 * cols + cols + c] = smax[0];
        }
        __syncthreads();
    }
}

// numpy's pairwise_sum for n <= 128 (numpy/core/src/umath/loops_utils.h.src): 8 running sums, combined as a tree, tail added
template <typename T>
__device__ __forceinline__ T np_pairwise_sum(const T* a, int n) {
    if (n < 8) {
        T res = (T)0;
        for (int i = 0; i < n; ++i) res += a[i];
        return res;
    }
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] += a[i + j];
    }
    T res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
    for (; i < n; ++i) res += a[i];
    return res;
}

template <typename T>
__global__ void __launch_bounds__(128)
sc_scale_kernel(const T* __restrict__ x, int64_t ldx, int64_t rows, int cols, const T* __restrict__ part, int nparts,
                T* __restrict__ out, int64_t ldo, double* __restrict__ acc, int64_t ldacc) {
    __shared__ T cmin[SC_MAX_COLS], cmax[SC_MAX_COLS];
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        T mn = (T)INFINITY, mx = (T)-INFINITY;
        for (int b = 0; b < nparts; ++b) {
            const T a = part[(int64_t)b * 2 * cols + c], z = part[(int64_t)b * 2 * cols + cols + c];
            mn = a < mn ? a : mn;
            mx = z > mx ? z : mx;
        }
        cmin[c] = mn;
        cmax[c] = mx - mn;          // max over the column of (x - min): subtraction of a constant is monotone
    }
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    T v[SC_MAX_COLS];
    for (int c = 0; c < cols; ++c) v[c] = (x[r * ldx + c] - cmin[c]) / cmax[c];
    const T s = np_pairwise_sum(v, cols);
    for (int c = 0; c < cols; ++c) {
        const T o = v[c] / s;
        if (out) out[r * ldo + c] = o;
        if (acc) acc[r * ldacc + c] += (double)o;
    }
}

__global__ void __launch_bounds__(256) sc_divide_kernel(double* x, int64_t ld, int64_t rows, int cols, double d) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        const int c = (int)(i - r * cols);
        x[r * ld + c] /= d;
    }
}

// order-preserving key of a double: NaN largest (numpy sorts NaN last), -0 == +0
__device__ __forceinline__ unsigned long long sc_key(double v) {
    if (isnan(v)) return 0xFFFFFFFFFFFFFFFFull;
    if (v == 0.0) v = 0.0;
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

struct ScItem {
    unsigned long long key;
    long long idx;
};

__global__ void __launch_bounds__(256)
sc_diff_kernel(const double* __restrict__ normal, int64_t ldn, const double* __restrict__ inter, int64_t ldi, int64_t rows,
               int cols, double* __restrict__ diff, int64_t ldd, ScItem* __restrict__ items, int64_t padded) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < padded; i += (int64_t)gridDim.x * blockDim.x) {
        if (i < total) {
            const int64_t r = i / cols;
            const int c = (int)(i - r * cols);
            const double a = normal[r * ldn + c];
            const double d = (inter[r * ldi + c] - a) / a;
            diff[r * ldd + c] = d;
            items[i] = ScItem{sc_key(d), i};
        } else {
            items[i] = ScItem{0xFFFFFFFFFFFFFFFFull, i};      // padding sorts behind every real element (index >= total)
        }
    }
}

__device__ __forceinline__ bool sc_less(const ScItem& a, const ScItem& b) { return a.key < b.key || (a.key == b.key && a.idx < b.idx); }

__global__ void __launch_bounds__(256) sc_bitonic_step_kernel(ScItem* items, int64_t n, int64_t j, int64_t k) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t l = i ^ j;
    if (l > i) {
        const ScItem a = items[i], b = items[l];
        const bool up = (i & k) == 0;
        if (up ? sc_less(b, a) : sc_less(a, b)) {
            items[i] = b;
            items[l] = a;
        }
    }
}

__global__ void __launch_bounds__(256) sc_emit_order_kernel(const ScItem* __restrict__ items, int64_t total, int64_t* __restrict__ order) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) order[i] = items[total - 1 - i].idx;       // largest first
}

static int64_t next_pow2(int64_t n) {
    int64_t p = 1;
    while (p < n) p <<= 1;
    return p;
}
static int sc_parts(int64_t rows) {
    const int64_t b = ceil_div(rows, 256);
    return (int)(b < sm_count() ? b : sm_count());
}

}  // namespace plagnn

using namespace plagnn;

extern "C" {

size_t plagnn_scoring_workspace_bytes(int64_t rows, int64_t cols) {
    if (rows <= 0 || cols <= 0) return 0;
    const size_t minmax = (size_t)sm_count() * 2 * (size_t)cols * sizeof(double);
    const size_t sort = (size_t)next_pow2(rows * cols) * sizeof(ScItem);
    return align_up(minmax, 256) + align_up(sort, 256);
}

int plagnn_scaling(const void* x, int is_f64, int64_t ldx, int64_t rows, int64_t cols, void* out, int64_t ldo, double* acc,
                   int64_t ldacc, void* workspace, size_t workspace_bytes, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ProfileScope prof("scaling", rows, cols, is_f64, stream);
    if (!x || rows <= 0 || cols <= 0 || cols > SC_MAX_COLS || ldx < cols || (out && ldo < cols) || (acc && ldacc < cols) || (!out && !acc))
        return fail(PLAGNN_ERR_ARG, "scaling", "bad arguments (cols <= 128)");
    const int nparts = sc_parts(rows);
    if (!workspace || workspace_bytes < (size_t)nparts * 2 * cols * sizeof(double))
        return fail(PLAGNN_ERR_WORKSPACE, "scaling", "workspace too small (plagnn_scoring_workspace_bytes)");
    const unsigned grid = (unsigned)ceil_div(rows, 128);
    if (is_f64) {
        sc_colminmax_kernel<double><<<nparts, 256, 0, st>>>((const double*)x, ldx, rows, (int)cols, (double*)workspace);
        sc_scale_kernel<double><<<grid, 128, 0, st>>>((const double*)x, ldx, rows, (int)cols, (const double*)workspace, nparts,
                                                        (double*)out, ldo, acc, ldacc);
    } else {
        sc_colminmax_kernel<float><<<nparts, 256, 0, st>>>((const float*)x, ldx, rows, (int)cols, (float*)workspace);
        sc_scale_kernel<float><<<grid, 128, 0, st>>>((const float*)x, ldx, rows, (int)cols, (const float*)workspace, nparts,
                                                       (float*)out, ldo, acc, ldacc);
    }
    return check_launch("scaling", 2);
}

int plagnn_divide_f64(double* x, int64_t ld, int64_t rows, int64_t cols, double divisor, plagnn_stream_t stream) {
    if (!x || rows <= 0 || cols <= 0 || ld < cols) return fail(PLAGNN_ERR_ARG, "divide_f64", "bad arguments");
    const int64_t total = rows * cols;
    const int g = (int)(ceil_div(total, 256) < (int64_t)sm_count() * 8 ? ceil_div(total, 256) : (int64_t)sm_count() * 8);
    sc_divide_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(x, ld, rows, (int)cols, divisor);
    return check_launch("divide_f64");
}

int plagnn_alteration_rank(const double* normal, int64_t ldn, const double* inter, int64_t ldi, int64_t rows, int64_t cols,
                           double* diff, int64_t ldd, int64_t* order, void* workspace, size_t workspace_bytes,
                           plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ProfileScope prof("alteration_rank", rows, cols, 0, stream);
    if (!normal || !inter || !diff || !order || rows <= 0 || cols <= 0 || ldn < cols || ldi < cols || ldd < cols)
        return fail(PLAGNN_ERR_ARG, "alteration_rank", "bad arguments");
    const int64_t total = rows * cols, padded = next_pow2(total);
    const size_t off = align_up((size_t)sm_count() * 2 * (size_t)cols * sizeof(double), 256);
    if (!workspace || workspace_bytes < off + (size_t)padded * sizeof(ScItem))
        return fail(PLAGNN_ERR_WORKSPACE, "alteration_rank", "workspace too small (plagnn_scoring_workspace_bytes)");
    ScItem* items = reinterpret_cast<ScItem*>(static_cast<char*>(workspace) + off);
    const unsigned g = (unsigned)ceil_div(padded, 256);
    sc_diff_kernel<<<g, 256, 0, st>>>(normal, ldn, inter, ldi, rows, (int)cols, diff, ldd, items, padded);
    int launches = 1;
    for (int64_t k = 2; k <= padded; k <<= 1)
        for (int64_t j = k >> 1; j > 0; j >>= 1) {
            sc_bitonic_step_kernel<<<g, 256, 0, st>>>(items, padded, j, k);
            ++launches;
        }
    sc_emit_order_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, st>>>(items, total, order);
    return check_launch("alteration_rank", launches + 1);
}

}  // extern "C"
