// Preprocessing kernels — the two O(E·d) / O(N²) steps of the offline stage that produce the
// inputs of the hot path:
//   code/data_preprocess.py:175-214   edge_clustering_coefficients(ppi_net, epsilon)
//       for every edge i<j of the (symmetric, 0/1) PPI matrix: triangles = |N(i) ∩ N(j)|,
//       value = triangles / (min(deg_i, deg_j) - 1)  (epsilon when the denominator is 0); entries (i,j),(j,i) appended in
//       the order of the row loop (i ascending, then j ascending)
//   code/data_preprocess.py:217-257   modify_network_topology(ppi_net, pcc_nor, pcc_inter, thr)
//       diff = pcc_inter - pcc_nor (dense N×N float64); thresholds mean ∓ thr·std; an existing edge with diff < left is
//       removed, a missing edge with diff > right is added; result as row-major COO
// Integer work and IEEE-exact float64 operations only (count, one division, one subtraction, two comparisons), so the
// outputs are bit-identical to the reference's numpy/scipy code.  The only order-dependent quantity is the mean / standard
// deviation of diff (numpy: pairwise summation; here: a fixed two-level tree) — the caller forms the thresholds on the host
// exactly as the reference does and may pass any thresholds to plagnn_rewire.
// Bounds: ECC is latency/L2 bound (sorted-list intersections by binary search, one warp per edge); the moment and rewiring
// passes stream 16 bytes per matrix element once each (HBM bound, coalesced rows).
#include "common.cuh"
#include <math.h>

namespace plagnn {

constexpr int PP_THREADS = 256;
constexpr int PP_SCAN_THREADS = 1024;
constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------- shared: exclusive scan of per-row counts (one block)
// ptr[0..n] <- exclusive scan of cnt[0..n); *total_out (optional) <- scale * ptr[n].  n is a node count (≤ a few million): one
// 1024-thread block, every thread owns one contiguous slice.
__global__ void __launch_bounds__(PP_SCAN_THREADS)
pp_scan_kernel(const int32_t* __restrict__ cnt, int64_t n, int32_t* __restrict__ ptr, int64_t* __restrict__ total_out, int scale) {
    __shared__ long long part[PP_SCAN_THREADS];
    const int t = threadIdx.x;
    const int64_t per = (n + PP_SCAN_THREADS - 1) / PP_SCAN_THREADS;
    int64_t b = (int64_t)t * per, e = b + per;
    if (b > n) b = n;
    if (e > n) e = n;
    long long s = 0;
    for (int64_t k = b; k < e; ++k) s += cnt[k];
    part[t] = s;
    __syncthreads();
    for (int off = 1; off < PP_SCAN_THREADS; off <<= 1) {
        const long long v = t >= off ? part[t - off] : 0;
        __syncthreads();
        part[t] += v;
        __syncthreads();
    }
    long long run = part[t] - s;
    for (int64_t k = b; k < e; ++k) {
        const int32_t c = cnt[k];
        ptr[k] = (int32_t)run;
        run += c;
    }
    if (t == PP_SCAN_THREADS - 1) {
        ptr[n] = (int32_t)part[t];
        if (total_out) *total_out = (int64_t)scale * part[t];
    }
}

// ---------------------------------------------------------------- edge clustering coefficients
// per row: position of its first column > row id (rows have ascending columns), and how many such columns
__global__ void __launch_bounds__(PP_THREADS)
ecc_upper_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n,
                 int32_t* __restrict__ upper_start, int32_t* __restrict__ upper_cnt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int lo = indptr[i], hi = indptr[i + 1];
    const int end = hi;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (indices[mid] > (int)i) hi = mid; else lo = mid + 1;
    }
    upper_start[i] = lo;
    upper_cnt[i] = end - lo;
}

// one warp per stored entry e = (i, j); entries with j <= i return at once.  status bits: 1 = a row is not strictly
// ascending (unsorted or duplicate entry), 2 = column id out of range, 4 = output capacity too small.
__global__ void __launch_bounds__(PP_THREADS)
ecc_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int64_t n, int64_t nnz,
           const int32_t* __restrict__ upper_start, const int32_t* __restrict__ upper_ptr, double epsilon, int64_t capacity,
           int32_t* __restrict__ ecc_row, int32_t* __restrict__ ecc_col, double* __restrict__ ecc_data, int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < nnz; e += nwarps) {
        // the row of entry e: the smallest r with indptr[r + 1] > e (indptr[n] = nnz > e)
        int lo = 0, hi = (int)n - 1;
        while (lo < hi) {
            const int mid = lo + ((hi - lo) >> 1);
            if ((int64_t)indptr[mid + 1] > e) hi = mid; else lo = mid + 1;
        }
        const int i = lo;
        const int j = indices[e];
        if (j < 0 || j >= n) {
            if (lane == 0) atomicOr(status, 2);
            continue;
        }
        if (lane == 0 && e > indptr[i] && indices[e - 1] >= j) atomicOr(status, 1);
        if (j <= i) continue;
        const int ai = indptr[i], bi = indptr[i + 1], aj = indptr[j], bj = indptr[j + 1];
        const int di = bi - ai, dj = bj - aj;
        // lanes walk the shorter neighbour list and look each id up in the longer one
        const int s0 = di <= dj ? ai : aj, s1 = di <= dj ? bi : bj;
        const int l0 = di <= dj ? aj : ai, l1 = di <= dj ? bj : bi;
        int count = 0;
        for (int p = s0 + lane; p < s1; p += 32) {
            const int x = indices[p];
            int a = l0, b = l1;
            while (a < b) {
                const int mid = a + ((b - a) >> 1);
                if (indices[mid] < x) a = mid + 1; else b = mid;
            }
            count += (a < l1 && indices[a] == x) ? 1 : 0;
        }
        count = __reduce_add_sync(FULL, count);
        if (lane == 0) {
            const int64_t rank = (int64_t)upper_ptr[i] + (e - (int64_t)upper_start[i]);
            if (e < upper_start[i]) {
                atomicOr(status, 1);
            } else if (2 * rank + 1 >= capacity) {
                atomicOr(status, 4);
            } else {
                const int possible = (di < dj ? di : dj) - 1;
                const double value = possible == 0 ? epsilon : (double)count / (double)possible;
                ecc_row[2 * rank] = i;
                ecc_col[2 * rank] = j;
                ecc_data[2 * rank] = value;
                ecc_row[2 * rank + 1] = j;
                ecc_col[2 * rank + 1] = i;
                ecc_data[2 * rank + 1] = value;
            }
        }
    }
}

// ---------------------------------------------------------------- mean / standard deviation of (inter - normal)
// pass 0: partial sums of d; pass 1: partial sums of (d - mean)^2 with mean = out[0].  Block b owns rows b, b + grid, ...
__global__ void __launch_bounds__(PP_THREADS)
pp_moment_kernel(const double* __restrict__ nor, int64_t ldn, const double* __restrict__ inter, int64_t ldi, int64_t rows,
                 int64_t cols, int pass, const double* __restrict__ out, double* __restrict__ part) {
    __shared__ double sh[PP_THREADS];
    const double mean = pass ? out[0] : 0.0;
    double s = 0.0;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const double* a = inter + r * ldi;
        const double* b = nor + r * ldn;
        for (int64_t c = threadIdx.x; c < cols; c += PP_THREADS) {
            const double d = a[c] - b[c];
            if (pass) {
                const double t = d - mean;
                s += t * t;
            } else {
                s += d;
            }
        }
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = PP_THREADS / 2; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(PP_THREADS)
pp_moment_finish_kernel(const double* __restrict__ part, int nparts, double count, int pass, double* __restrict__ out) {
    __shared__ double sh[PP_THREADS];
    double s = 0.0;
    for (int k = threadIdx.x; k < nparts; k += PP_THREADS) s += part[k];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int k = PP_THREADS / 2; k > 0; k >>= 1) {
        if ((int)threadIdx.x < k) sh[threadIdx.x] += sh[threadIdx.x + k];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[pass] = pass ? sqrt(sh[0] / count) : sh[0] / count;
}

// ---------------------------------------------------------------- adjacency as a bit matrix, rewiring, COO emission
__global__ void __launch_bounds__(PP_THREADS)
pp_bitmask_kernel(const int32_t* __restrict__ row, const int32_t* __restrict__ col, int64_t nnz, int64_t n,
                  uint32_t* __restrict__ mask, int64_t wpr, int32_t* __restrict__ status) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const int r = row[e], c = col[e];
    if (r < 0 || r >= n || c < 0 || c >= n) {
        atomicOr(status, 2);
        return;
    }
    atomicOr(&mask[(int64_t)r * wpr + (c >> 5)], 1u << (c & 31));
}

// block b owns rows b, b + grid, ...; a warp decides 32 consecutive columns (one mask word) per step
__global__ void __launch_bounds__(PP_THREADS)
pp_rewire_kernel(const double* __restrict__ nor, int64_t ldn, const double* __restrict__ inter, int64_t ldi, int64_t n,
                 const uint32_t* __restrict__ mask, uint32_t* __restrict__ new_mask, int64_t wpr, double l_thr, double r_thr,
                 int32_t* __restrict__ row_cnt) {
    __shared__ int block_cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
        if (threadIdx.x == 0) block_cnt = 0;
        __syncthreads();
        int cnt = 0;
        for (int64_t w = warp; w < wpr; w += PP_THREADS / 32) {
            const int64_t c = w * 32 + lane;
            const uint32_t word = mask[r * wpr + w];
            const bool present = (word >> lane) & 1u;
            bool keep = false;
            if (c < n) {
                const double d = inter[r * ldi + c] - nor[r * ldn + c];
                keep = present;
                if (d < l_thr && present) keep = false;
                if (d > r_thr && !present) keep = true;
            }
            const uint32_t nw = __ballot_sync(FULL, keep);
            if (lane == 0) {
                new_mask[r * wpr + w] = nw;
                cnt += __popc(nw);
            }
        }
        if (lane == 0 && cnt) atomicAdd(&block_cnt, cnt);
        __syncthreads();
        if (threadIdx.x == 0) row_cnt[r] = block_cnt;
        __syncthreads();
    }
}

// one warp per row: set bits of the row in ascending column order, written from rowptr[row]
__global__ void __launch_bounds__(PP_THREADS)
pp_emit_kernel(const uint32_t* __restrict__ mask, int64_t wpr, int64_t n, const int32_t* __restrict__ rowptr,
               int32_t* __restrict__ out_row, int32_t* __restrict__ out_col) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nwarps) {
        int64_t base = rowptr[r];
        for (int64_t w0 = 0; w0 < wpr; w0 += 32) {
            const int64_t w = w0 + lane;
            uint32_t word = w < wpr ? mask[r * wpr + w] : 0u;
            const int pc = __popc(word);
            int incl = pc;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(FULL, incl, off);
                if (lane >= off) incl += v;
            }
            const int total = __shfl_sync(FULL, incl, 31);
            int64_t pos = base + (incl - pc);
            while (word) {
                const int b = __ffs(word) - 1;
                out_row[pos] = (int32_t)r;
                out_col[pos] = (int32_t)(w * 32 + b);
                ++pos;
                word &= word - 1;
            }
            base += total;
        }
    }
}

static int pp_stream_blocks(int64_t rows) {
    const int64_t cap = (int64_t)sm_count() * 8;      // 8 resident 256-thread blocks per SM
    return (int)(rows < cap ? (rows > 0 ? rows : 1) : cap);
}

}  // namespace plagnn

using namespace plagnn;

// ---------------------------------------------------------------- Pearson matrix of the expression rows (np.corrcoef)
// code/data_preprocess.py:165-170: expr_pcc = np.corrcoef(expr_gcn); fill_diagonal(0); NaN -> 0.  numpy's steps, restated:
// X = x - mean(x, axis 1); c = (X X^T) * (1 / (S - 1)); sd = sqrt(diag(c)); c /= sd[:, None]; c /= sd[None, :]; clip to
// [-1, 1].  The product X X^T is a BLAS call there (S = 3 samples: three terms per entry, whose rounding depends on the BLAS
// micro-kernel); here it is a fused-multiply-add chain in sample order, which agrees with numpy to <= 2 ulp on every entry
// probed, so the test bar is 4 ulp on values in [-1, 1], not bit-exactness.
__global__ void __launch_bounds__(PP_THREADS)
pp_center_kernel(const double* __restrict__ x, int64_t ldx, int64_t rows, int samples, double* __restrict__ xc /* rows x samples */,
                 double* __restrict__ sd /* rows */) {
    const int64_t r = (int64_t)blockIdx.x * PP_THREADS + threadIdx.x;
    if (r >= rows) return;
    double s = 0.0;
    for (int k = 0; k < samples; ++k) s += x[r * ldx + k];                 // np.add.reduce over < 8 items: left to right
    const double mean = s / (double)samples;
    double acc = 0.0;
    for (int k = 0; k < samples; ++k) {
        const double v = x[r * ldx + k] - mean;
        xc[r * samples + k] = v;
        acc = k ? fma(v, v, acc) : v * v;
    }
    sd[r] = sqrt(acc * (1.0 / (double)(samples - 1)));
}

__global__ void __launch_bounds__(PP_THREADS)
pp_pearson_kernel(const double* __restrict__ xc, const double* __restrict__ sd, int64_t rows, int samples, double* __restrict__ out,
                  int64_t ldo) {
    const int64_t j = (int64_t)blockIdx.x * PP_THREADS + threadIdx.x;      // consecutive threads: consecutive columns
    const double fact = 1.0 / (double)(samples - 1);
    for (int64_t i = blockIdx.y; i < rows; i += gridDim.y) {
        if (j >= rows) continue;
        double acc = 0.0;
        for (int k = 0; k < samples; ++k) {
            const double a = xc[i * samples + k], b = xc[j * samples + k];
            acc = k ? fma(a, b, acc) : a * b;
        }
        double c = ((acc * fact) / sd[i]) / sd[j];
        c = fmin(fmax(c, -1.0), 1.0);                                      // np.clip (NaN stays NaN: fmin / fmax would drop it)
        if (i == j || !(c == c) || !(sd[i] == sd[i]) || sd[i] == 0.0 || sd[j] == 0.0) c = 0.0;   // diagonal, NaN -> 0
        out[i * ldo + j] = c;
    }
}

extern "C" {

size_t plagnn_pearson_workspace_bytes(int64_t rows, int64_t samples) {
    if (rows <= 0 || samples <= 0) return 0;
    return align_up((size_t)rows * (size_t)(samples + 1) * sizeof(double), 256);
}

int plagnn_pearson(const double* expr, int64_t ldx, int64_t rows, int64_t samples, double* out, int64_t ldo, void* workspace,
                   size_t workspace_bytes, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ProfileScope prof("pearson", rows, samples, 0, stream);
    if (!expr || !out || rows <= 0 || samples < 2 || samples > 4096 || ldx < samples || ldo < rows)
        return fail(PLAGNN_ERR_ARG, "pearson", "bad arguments (at least two samples)");
    if (!workspace || workspace_bytes < plagnn_pearson_workspace_bytes(rows, samples))
        return fail(PLAGNN_ERR_WORKSPACE, "pearson", "workspace too small");
    double* xc = reinterpret_cast<double*>(workspace);
    double* sd = xc + rows * samples;
    pp_center_kernel<<<(unsigned)ceil_div(rows, PP_THREADS), PP_THREADS, 0, st>>>(expr, ldx, rows, (int)samples, xc, sd);
    const int64_t gy = rows < (int64_t)sm_count() * 8 ? rows : (int64_t)sm_count() * 8;
    pp_pearson_kernel<<<dim3((unsigned)ceil_div(rows, PP_THREADS), (unsigned)gy), PP_THREADS, 0, st>>>(xc, sd, rows, (int)samples, out, ldo);
    return check_launch("pearson", 2);
}

size_t plagnn_ecc_workspace_bytes(int64_t num_nodes) {
    if (num_nodes <= 0) return 0;
    return 3 * align_up((size_t)(num_nodes + 1) * sizeof(int32_t), 256);
}

int plagnn_ecc(const int32_t* indptr, const int32_t* indices, int64_t num_nodes, int64_t nnz, double epsilon, int32_t* ecc_row,
               int32_t* ecc_col, double* ecc_data, int64_t capacity, int64_t* n_out, int32_t* status, void* workspace,
               size_t workspace_bytes, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ProfileScope prof("ecc", num_nodes, nnz, 0, stream);
    if (!indptr || !n_out || !status || num_nodes <= 0 || nnz < 0 || num_nodes > 0x7fffffff || nnz > 0x7fffffff || capacity < 0 ||
        (nnz > 0 && (!indices || !ecc_row || !ecc_col || !ecc_data)))
        return fail(PLAGNN_ERR_ARG, "ecc", "bad arguments");
    if (!workspace || workspace_bytes < plagnn_ecc_workspace_bytes(num_nodes))
        return fail(PLAGNN_ERR_WORKSPACE, "ecc", "workspace too small (plagnn_ecc_workspace_bytes)");
    const size_t seg = align_up((size_t)(num_nodes + 1) * sizeof(int32_t), 256);
    int32_t* upper_start = reinterpret_cast<int32_t*>(workspace);
    int32_t* upper_cnt = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + seg);
    int32_t* upper_ptr = reinterpret_cast<int32_t*>(static_cast<char*>(workspace) + 2 * seg);
    PLAGNN_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    ecc_upper_kernel<<<(unsigned)ceil_div(num_nodes, PP_THREADS), PP_THREADS, 0, st>>>(indptr, indices, num_nodes, upper_start, upper_cnt);
    pp_scan_kernel<<<1, PP_SCAN_THREADS, 0, st>>>(upper_cnt, num_nodes, upper_ptr, n_out, 2);
    int launches = 2;
    if (nnz > 0) {
        const int64_t want = ceil_div(nnz, PP_THREADS / 32);
        const int64_t cap = (int64_t)sm_count() * 8;
        ecc_kernel<<<(unsigned)(want < cap ? want : cap), PP_THREADS, 0, st>>>(indptr, indices, num_nodes, nnz, upper_start, upper_ptr,
                                                                             epsilon, capacity, ecc_row, ecc_col, ecc_data, status);
        ++launches;
    }
    return check_launch("ecc", launches);
}

size_t plagnn_diff_moments_workspace_bytes(void) { return align_up((size_t)sm_count() * 8 * sizeof(double), 256); }

int plagnn_diff_moments(const double* normal, int64_t ldn, const double* inter, int64_t ldi, int64_t rows, int64_t cols,
                        double* mean_std, void* workspace, size_t workspace_bytes, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ProfileScope prof("diff_moments", rows, cols, 0, stream);
    if (!normal || !inter || !mean_std || rows <= 0 || cols <= 0 || ldn < cols || ldi < cols)
        return fail(PLAGNN_ERR_ARG, "diff_moments", "bad arguments");
    if (!workspace || workspace_bytes < plagnn_diff_moments_workspace_bytes())
        return fail(PLAGNN_ERR_WORKSPACE, "diff_moments", "workspace too small (plagnn_diff_moments_workspace_bytes)");
    const int blocks = pp_stream_blocks(rows);
    double* part = reinterpret_cast<double*>(workspace);
    const double count = (double)rows * (double)cols;
    for (int pass = 0; pass < 2; ++pass) {
        pp_moment_kernel<<<blocks, PP_THREADS, 0, st>>>(normal, ldn, inter, ldi, rows, cols, pass, mean_std, part);
        pp_moment_finish_kernel<<<1, PP_THREADS, 0, st>>>(part, blocks, count, pass, mean_std);
    }
    return check_launch("diff_moments", 4);
}

int plagnn_adj_bitmask(const int32_t* row, const int32_t* col, int64_t nnz, int64_t num_nodes, uint32_t* mask,
                       int64_t words_per_row, int32_t* status, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!mask || !status || num_nodes <= 0 || nnz < 0 || words_per_row * 32 < num_nodes || (nnz > 0 && (!row || !col)))
        return fail(PLAGNN_ERR_ARG, "adj_bitmask", "bad arguments (words_per_row >= ceil(num_nodes / 32))");
    PLAGNN_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
    PLAGNN_CUDA_TRY(cudaMemsetAsync(mask, 0, (size_t)num_nodes * (size_t)words_per_row * sizeof(uint32_t), st));
    if (nnz == 0) return PLAGNN_OK;
    pp_bitmask_kernel<<<(unsigned)ceil_div(nnz, PP_THREADS), PP_THREADS, 0, st>>>(row, col, nnz, num_nodes, mask, words_per_row, status);
    return check_launch("adj_bitmask");
}

size_t plagnn_rewire_workspace_bytes(int64_t num_nodes) {
    if (num_nodes <= 0) return 0;
    return align_up((size_t)num_nodes * sizeof(int32_t), 256);
}

int plagnn_rewire(const double* normal, int64_t ldn, const double* inter, int64_t ldi, int64_t num_nodes, const uint32_t* mask,
                  uint32_t* new_mask, int64_t words_per_row, double l_threshold, double r_threshold, int32_t* rowptr,
                  int64_t* n_out, void* workspace, size_t workspace_bytes, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ProfileScope prof("rewire", num_nodes, words_per_row, 0, stream);
    if (!normal || !inter || !mask || !new_mask || !rowptr || !n_out || num_nodes <= 0 || num_nodes > 0x7fffffff ||
        ldn < num_nodes || ldi < num_nodes || words_per_row * 32 < num_nodes || mask == new_mask)
        return fail(PLAGNN_ERR_ARG, "rewire", "bad arguments");
    if (!workspace || workspace_bytes < plagnn_rewire_workspace_bytes(num_nodes))
        return fail(PLAGNN_ERR_WORKSPACE, "rewire", "workspace too small (plagnn_rewire_workspace_bytes)");
    int32_t* row_cnt = reinterpret_cast<int32_t*>(workspace);
    pp_rewire_kernel<<<pp_stream_blocks(num_nodes), PP_THREADS, 0, st>>>(normal, ldn, inter, ldi, num_nodes, mask, new_mask,
                                                                         words_per_row, l_threshold, r_threshold, row_cnt);
    pp_scan_kernel<<<1, PP_SCAN_THREADS, 0, st>>>(row_cnt, num_nodes, rowptr, n_out, 1);
    return check_launch("rewire", 2);
}

int plagnn_bitmask_to_coo(const uint32_t* mask, int64_t words_per_row, int64_t num_nodes, const int32_t* rowptr, int32_t* out_row,
                          int32_t* out_col, plagnn_stream_t stream) {
    if (!mask || !rowptr || !out_row || !out_col || num_nodes <= 0 || words_per_row * 32 < num_nodes)
        return fail(PLAGNN_ERR_ARG, "bitmask_to_coo", "bad arguments");
    const int64_t want = ceil_div(num_nodes, PP_THREADS / 32);
    const int64_t cap = (int64_t)sm_count() * 8;
    pp_emit_kernel<<<(unsigned)(want < cap ? want : cap), PP_THREADS, 0, (cudaStream_t)stream>>>(mask, words_per_row, num_nodes, rowptr,
                                                                                              out_row, out_col);
    return check_launch("bitmask_to_coo");
}

}  // extern "C"
