// K1 — device CSR/CSC builder and aggregation plan.
//
// Replaces the lazy COO->CSR/CSC conversion DGL performs for the graph the reference builds in
// code/utils.py:44-45 (dgl.graph((start,end), num_nodes) ; dgl.add_self_loop(g)).  The result must be
// bit-exact with a stable sort of edge ids by key (the CPU oracle's coo_to_csc), so the sort
// is a hand-written *stable* LSD radix sort (8-bit digits) rather than an atomic scatter.
//
// All of this is HBM-bound integer work: coalesced streaming reads, shared-memory digit counters,
// grids sized by the input (set-up code, runs once per graph).
#include "common.cuh"
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace plagnn {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- optional event profile ---------------------------------------------------------------------
struct ProfRec {
    char name[24];
    long long tag[3];
    cudaEvent_t beg, end;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec> g_prof;
static std::atomic<int> g_prof_on{0};

ProfileScope::ProfileScope(const char* name, long long t0, long long t1, long long t2, plagnn_stream_t stream)
    : slot(-1), st((cudaStream_t)stream) {
    const int mode = g_prof_on.load(std::memory_order_relaxed);
    if (!mode) return;
    if (mode == 2 && strncmp(name, "spmm", 4) != 0) return;      // aggregation kernels only (low-overhead mode)
    ProfRec r{};
    snprintf(r.name, sizeof(r.name), "%s", name);
    r.tag[0] = t0; r.tag[1] = t1; r.tag[2] = t2;
    if (cudaEventCreate(&r.beg) != cudaSuccess || cudaEventCreate(&r.end) != cudaSuccess) return;
    cudaEventRecord(r.beg, st);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
    slot = (int)g_prof.size() - 1;
}
ProfileScope::~ProfileScope() {
    if (slot < 0) return;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (slot < (int)g_prof.size()) cudaEventRecord(g_prof[slot].end, st);
}

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("PLAGNN_PDL"); return !e || e[0] != '0'; }();
    return on;
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 148;
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached = v;
        cached_dev = dev;
    }
    return cached;
}

// ======================================================================================
// exclusive scan (int32), three-phase, recursive over block totals
// ======================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int warp_incl_scan(int v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// in-place exclusive scan of each SCAN_TILE-sized tile; tile totals to block_sums (may be null)
__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles_kernel(int32_t* data, int64_t n, int32_t* block_sums) {
    __shared__ int warp_tot[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? data[base + i] : 0;
        sum += v[i];
    }
    const int incl = warp_incl_scan(sum);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int t = (lane < SCAN_THREADS / 32) ? warp_tot[lane] : 0;
        int ti = warp_incl_scan(t);
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = ti - t;   // exclusive over warps
        if (lane == SCAN_THREADS / 32 - 1 && block_sums) block_sums[blockIdx.x] = ti;
    }
    __syncthreads();
    int run = warp_tot[warp] + incl - sum;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) data[base + i] = run;
        run += v[i];
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add_kernel(int32_t* data, int64_t n, const int32_t* block_offs) {
    const int off = block_offs[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE;
    for (int i = threadIdx.x; i < SCAN_TILE; i += SCAN_THREADS)
        if (base + i < n) data[base + i] += off;
}

static size_t scan_workspace_ints(int64_t n) {
    size_t tot = 0;
    while (n > SCAN_TILE) {
        n = ceil_div(n, SCAN_TILE);
        tot += align_up((size_t)n, 64);
    }
    return tot + 64;
}

// data[0..n) <- exclusive scan (in place). ws: scan_workspace_ints(n) int32s.
static void exclusive_scan_i32(int32_t* data, int64_t n, int32_t* ws, cudaStream_t st) {
    if (n <= 0) return;
    const int64_t blocks = ceil_div(n, SCAN_TILE);
    if (blocks == 1) {
        scan_tiles_kernel<<<1, SCAN_THREADS, 0, st>>>(data, n, nullptr);
        return;
    }
    scan_tiles_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(data, n, ws);
    exclusive_scan_i32(ws, blocks, ws + align_up((size_t)blocks, 64), st);
    scan_add_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(data, n, ws);
}

// ======================================================================================
// stable LSD radix sort of (key, edge id), 8-bit digits
// ======================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_ROUNDS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;
constexpr int RS_WARPS = RS_THREADS / 32;

// virtual edge list: e < E -> (key[e], e); e >= E -> self-loop (e-E, e)
__device__ __forceinline__ int32_t virt_key(const int32_t* key, int64_t e, int64_t E) {
    return e < E ? key[e] : (int32_t)(e - E);
}

template <bool FIRST>
__global__ void __launch_bounds__(RS_THREADS) radix_hist_kernel(const int32_t* keys_in, const int32_t* coo_key,
                                                                int64_t E, int64_t n, int shift,
                                                                int32_t* hist, int num_blocks) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int r = 0; r < RS_ROUNDS; ++r) {
        const int64_t i = base + (int64_t)r * RS_THREADS + threadIdx.x;
        if (i < n) {
            const int32_t k = FIRST ? virt_key(coo_key, i, E) : keys_in[i];
            atomicAdd(&h[(k >> shift) & 255], 1);
        }
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * num_blocks + blockIdx.x] = h[threadIdx.x];
}

template <bool FIRST>
__global__ void __launch_bounds__(RS_THREADS) radix_scatter_kernel(const int32_t* keys_in, const int32_t* vals_in,
                                                                   const int32_t* coo_key, int64_t E, int64_t n,
                                                                   int shift, const int32_t* hist_scanned,
                                                                   int num_blocks, int32_t* keys_out,
                                                                   int32_t* vals_out) {
    __shared__ int cnt[RS_WARPS][256];
    __shared__ int basep[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    basep[threadIdx.x] = hist_scanned[(int64_t)threadIdx.x * num_blocks + blockIdx.x];
    const int64_t tile = (int64_t)blockIdx.x * RS_TILE;
    for (int r = 0; r < RS_ROUNDS; ++r) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) cnt[w][threadIdx.x] = 0;
        __syncthreads();
        const int64_t i = tile + (int64_t)r * RS_THREADS + threadIdx.x;
        const bool valid = i < n;
        int32_t k = 0, v = 0;
        if (valid) {
            k = FIRST ? virt_key(coo_key, i, E) : keys_in[i];
            v = FIRST ? (int32_t)i : vals_in[i];
        }
        const int digit = valid ? ((k >> shift) & 255) : 256;   // 256 = "no element" class
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) cnt[warp][digit] = __popc(peers);
        __syncthreads();
        {
            int run = basep[threadIdx.x];
#pragma unroll
            for (int w = 0; w < RS_WARPS; ++w) {
                const int t = cnt[w][threadIdx.x];
                cnt[w][threadIdx.x] = run;
                run += t;
            }
            basep[threadIdx.x] = run;
        }
        __syncthreads();
        if (valid) {
            const int pos = cnt[warp][digit] + rank;
            keys_out[pos] = k;
            vals_out[pos] = v;
        }
        __syncthreads();
    }
}

__global__ void finalize_csr_kernel(const int32_t* sorted_keys, const int32_t* sorted_eids, const int32_t* other,
                                    int64_t E, int64_t n, int64_t num_nodes, int32_t* indptr, int32_t* indices,
                                    int32_t* eids) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t k = sorted_keys[i];
    const int32_t e = sorted_eids[i];
    eids[i] = e;
    indices[i] = e < E ? other[e] : (int32_t)(e - E);
    const int32_t prev = i == 0 ? -1 : sorted_keys[i - 1];
    // (range guards: out-of-range ids are reported by validate_keys_kernel, never written through)
    for (int64_t v = (int64_t)prev + 1; v <= k; ++v)
        if (v >= 0 && v <= num_nodes) indptr[v] = (int32_t)i;
    if (i == n - 1)
        for (int64_t v = (k < 0 ? 0 : (int64_t)k + 1); v <= num_nodes; ++v) indptr[v] = (int32_t)n;
}

__global__ void validate_keys_kernel(const int32_t* a, const int32_t* b, int64_t E, int32_t num_nodes,
                                     int32_t num_other, int* bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= E) return;
    const int32_t x = a[i], y = b[i];
    if (x < 0 || x >= num_nodes || y < 0 || y >= num_other) atomicOr(bad, 1);
}

static int key_bits(int64_t num_nodes) {
    int bits = 1;
    while (((int64_t)1 << bits) < num_nodes) ++bits;
    return bits;
}

struct CsrWs {
    size_t keys[2], vals[2], hist, scan, flag, total;
};
static CsrWs csr_ws_layout(int64_t num_nodes, int64_t n) {
    CsrWs w;
    const size_t arr = align_up((size_t)(n > 0 ? n : 1) * sizeof(int32_t), 256);
    const int64_t blocks = ceil_div(n > 0 ? n : 1, RS_TILE);
    size_t off = 0;
    w.keys[0] = off; off += arr;
    w.keys[1] = off; off += arr;
    w.vals[0] = off; off += arr;
    w.vals[1] = off; off += arr;
    w.hist = off; off += align_up((size_t)blocks * 256 * sizeof(int32_t), 256);
    w.scan = off; off += align_up(scan_workspace_ints(blocks * 256) * sizeof(int32_t), 256);
    w.flag = off; off += 256;
    w.total = off;
    (void)num_nodes;
    return w;
}

// ======================================================================================
// aggregation plan
//   plan (int32): [0..16) header {chunk, n_rows, max_items, max_hubs, ...}
//                 item_ptr[n_rows+1] | slot_ptr[n_rows+1] | hub_ptr[n_rows+1] | hub_rows[n_rows+1] | item_rec[max_items] (int4)
//   (every offset but the scan scratch depends on n_rows only, so consumers need no edge count)
//   item_rec[i] = {row, first in-edge, end in-edge, partial slot or -1 (row not split)}: everything a warp needs to start
//   gathering, in ONE 16-byte load.  (Until the last session of round 2 this was item_row[i] = row, and a warp walked
//   item_row -> item_ptr / indptr -> slot_ptr: three dependent global loads before its first gather, which on the 1 M-row /
//   100 M-edge graph at 32 columns is a fifth of the life of an average work item of 100 in-edges.)
// ======================================================================================
struct PlanLayout {
    size_t item_ptr, slot_ptr, hub_ptr, item_row, hub_rows, scan, total_ints;
    int64_t max_items, max_hubs;
};
static PlanLayout plan_layout(int64_t n_rows, int64_t n_edges, int32_t chunk) {
    PlanLayout p;
    p.max_items = n_rows + n_edges / chunk + 1;
    p.max_hubs = (n_edges / chunk + 1 < n_rows ? n_edges / chunk + 1 : n_rows) + 1;
    size_t off = 16;
    const size_t rp = align_up((size_t)n_rows + 1, 64);
    p.item_ptr = off; off += rp;
    p.slot_ptr = off; off += rp;
    p.hub_ptr = off; off += rp;
    p.hub_rows = off; off += rp;
    p.item_row = off; off += align_up((size_t)p.max_items * 4, 64);      // int4 records, 16-byte aligned (off is a multiple of 64)
    p.scan = off; off += scan_workspace_ints(n_rows + 1);
    p.total_ints = off;
    return p;
}

__global__ void plan_count_kernel(const int32_t* indptr, int64_t n_rows, int32_t chunk, int32_t* item_ptr,
                                  int32_t* slot_ptr, int32_t* hub_ptr) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v > n_rows) return;
    int nch = 0;
    if (v < n_rows) {
        const int deg = indptr[v + 1] - indptr[v];
        nch = deg <= chunk ? 1 : (deg + chunk - 1) / chunk;
    }
    item_ptr[v] = nch;
    slot_ptr[v] = nch > 1 ? nch : 0;
    hub_ptr[v] = nch > 1 ? 1 : 0;
}

__global__ void plan_fill_kernel(const int32_t* indptr, int64_t n_rows, int32_t chunk, const int32_t* item_ptr,
                                 const int32_t* slot_ptr, const int32_t* hub_ptr, int4* item_rec, int32_t* hub_rows) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_rows) return;
    const int b = item_ptr[v], e = item_ptr[v + 1];
    const int rbeg = indptr[v], rend = indptr[v + 1];
    const int slot0 = slot_ptr[v];
    for (int i = b; i < e; ++i) {
        const int beg = rbeg + (i - b) * chunk;
        item_rec[i] = make_int4((int32_t)v, beg, min(rend, beg + chunk), e - b > 1 ? slot0 + (i - b) : -1);
    }
    if (e - b > 1) hub_rows[hub_ptr[v]] = (int32_t)v;
}

}  // namespace plagnn

using namespace plagnn;

extern "C" {

int plagnn_version(void) { return 100; }
long long plagnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int plagnn_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (on) {
        for (auto& r : g_prof) { cudaEventDestroy(r.beg); cudaEventDestroy(r.end); }
        g_prof.clear();
    }
    g_prof_on.store(on == 2 ? 2 : on ? 1 : 0);
    return PLAGNN_OK;
}

// Synchronises the recorded events and writes one line per (name, tags): "name t0 t1 t2 calls total_ms\n".
size_t plagnn_profile_report(char* buf, size_t cap) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& r : g_prof) {
        float ms = 0.f;
        if (cudaEventSynchronize(r.end) != cudaSuccess || cudaEventElapsedTime(&ms, r.beg, r.end) != cudaSuccess) continue;
        char key[96];
        snprintf(key, sizeof(key), "%s %lld %lld %lld", r.name, r.tag[0], r.tag[1], r.tag[2]);
        auto& a = agg[key];
        a.first += 1;
        a.second += ms;
    }
    std::string out;
    for (auto& kv : agg) {
        char line[160];
        snprintf(line, sizeof(line), "%s %lld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        out += line;
    }
    if (buf && cap) {
        const size_t nb = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), nb);
        buf[nb] = 0;
    }
    return out.size() + 1;
}
const char* plagnn_last_error(void) { return g_err; }

int plagnn_device_supported(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

size_t plagnn_csr_build_workspace_bytes(int64_t num_nodes, int64_t num_edges, int add_self_loop) {
    const int64_t n = num_edges + (add_self_loop ? num_nodes : 0);
    return csr_ws_layout(num_nodes, n).total;
}

int plagnn_csr_build(const int32_t* key, const int32_t* other, int64_t num_edges, int64_t num_nodes,
                     int64_t num_other_nodes, int add_self_loop, int32_t* indptr, int32_t* indices, int32_t* eids, void* workspace,
                     size_t workspace_bytes, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (num_edges < 0 || num_nodes <= 0 || !indptr) return fail(PLAGNN_ERR_ARG, "csr_build", "bad sizes");
    if (num_other_nodes <= 0) num_other_nodes = num_nodes;
    if (add_self_loop && num_other_nodes != num_nodes)
        return fail(PLAGNN_ERR_ARG, "csr_build", "self-loops need a square adjacency");
    if (num_edges > 0 && (!key || !other)) return fail(PLAGNN_ERR_ARG, "csr_build", "null COO arrays");
    const int64_t n = num_edges + (add_self_loop ? num_nodes : 0);
    if (n >= ((int64_t)1 << 31) || num_nodes >= ((int64_t)1 << 31) || num_other_nodes >= ((int64_t)1 << 31))
        return fail(PLAGNN_ERR_UNSUPPORTED, "csr_build", "more than 2^31-1 edges or nodes");
    if (n == 0) {
        PLAGNN_CUDA_TRY(cudaMemsetAsync(indptr, 0, (num_nodes + 1) * sizeof(int32_t), st));
        PLAGNN_CUDA_TRY(cudaStreamSynchronize(st));
        return PLAGNN_OK;
    }
    if (!indices || !eids) return fail(PLAGNN_ERR_ARG, "csr_build", "null outputs");
    const CsrWs w = csr_ws_layout(num_nodes, n);
    if (!workspace || workspace_bytes < w.total) return fail(PLAGNN_ERR_WORKSPACE, "csr_build", "workspace too small");
    char* ws = (char*)workspace;
    int32_t* keys[2] = {(int32_t*)(ws + w.keys[0]), (int32_t*)(ws + w.keys[1])};
    int32_t* vals[2] = {(int32_t*)(ws + w.vals[0]), (int32_t*)(ws + w.vals[1])};
    int32_t* hist = (int32_t*)(ws + w.hist);
    int32_t* scanws = (int32_t*)(ws + w.scan);
    int* flag = (int*)(ws + w.flag);

    PLAGNN_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
    if (num_edges > 0)
        validate_keys_kernel<<<(unsigned)ceil_div(num_edges, 256), 256, 0, st>>>(key, other, num_edges,
                                                                                (int32_t)num_nodes,
                                                                                (int32_t)num_other_nodes, flag);
    const int blocks = (int)ceil_div(n, RS_TILE);
    const int passes = (key_bits(num_nodes) + 7) / 8;
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        if (p == 0) {
            radix_hist_kernel<true><<<blocks, RS_THREADS, 0, st>>>(nullptr, key, num_edges, n, shift, hist, blocks);
        } else {
            radix_hist_kernel<false><<<blocks, RS_THREADS, 0, st>>>(keys[cur], nullptr, num_edges, n, shift, hist,
                                                                     blocks);
        }
        exclusive_scan_i32(hist, (int64_t)blocks * 256, scanws, st);
        if (p == 0) {
            radix_scatter_kernel<true><<<blocks, RS_THREADS, 0, st>>>(nullptr, nullptr, key, num_edges, n, shift, hist,
                                                                       blocks, keys[0], vals[0]);
            cur = 0;
        } else {
            radix_scatter_kernel<false><<<blocks, RS_THREADS, 0, st>>>(keys[cur], vals[cur], nullptr, num_edges, n,
                                                                        shift, hist, blocks, keys[cur ^ 1],
                                                                        vals[cur ^ 1]);
            cur ^= 1;
        }
    }
    finalize_csr_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(keys[cur], vals[cur], other, num_edges, n,
                                                                   num_nodes, indptr, indices, eids);
    int rc = check_launch("csr_build", 2 + passes * 5);
    if (rc) return rc;
    int host_flag = 0;
    PLAGNN_CUDA_TRY(cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PLAGNN_CUDA_TRY(cudaStreamSynchronize(st));
    if (host_flag) return fail(PLAGNN_ERR_ARG, "csr_build", "node id out of range");
    return PLAGNN_OK;
}

size_t plagnn_spmm_plan_bytes(int64_t num_rows, int64_t num_edges, int32_t chunk) {
    if (chunk < 32) chunk = 32;
    return plan_layout(num_rows, num_edges, chunk).total_ints * sizeof(int32_t);
}

int plagnn_spmm_plan_build(const int32_t* indptr, int64_t num_rows, int64_t num_edges, int32_t chunk, void* plan,
                           size_t plan_bytes, int64_t* host_counts, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!indptr || !plan || !host_counts || num_rows <= 0 || chunk < 32 || (chunk % 32) != 0)
        return fail(PLAGNN_ERR_ARG, "spmm_plan_build", "bad arguments (chunk must be a positive multiple of 32)");
    const PlanLayout L = plan_layout(num_rows, num_edges, chunk);
    if (plan_bytes < L.total_ints * sizeof(int32_t)) return fail(PLAGNN_ERR_WORKSPACE, "spmm_plan_build", "plan buffer too small");
    int32_t* P = (int32_t*)plan;
    int32_t hdr[16] = {chunk, (int32_t)num_rows, (int32_t)L.max_items, (int32_t)L.max_hubs};
    PLAGNN_CUDA_TRY(cudaMemcpyAsync(P, hdr, sizeof(hdr), cudaMemcpyHostToDevice, st));
    const unsigned g = (unsigned)ceil_div(num_rows + 1, 256);
    plan_count_kernel<<<g, 256, 0, st>>>(indptr, num_rows, chunk, P + L.item_ptr, P + L.slot_ptr, P + L.hub_ptr);
    exclusive_scan_i32(P + L.item_ptr, num_rows + 1, P + L.scan, st);
    exclusive_scan_i32(P + L.slot_ptr, num_rows + 1, P + L.scan, st);
    exclusive_scan_i32(P + L.hub_ptr, num_rows + 1, P + L.scan, st);
    plan_fill_kernel<<<g, 256, 0, st>>>(indptr, num_rows, chunk, P + L.item_ptr, P + L.slot_ptr, P + L.hub_ptr,
                                        reinterpret_cast<int4*>(P + L.item_row), P + L.hub_rows);
    int rc = check_launch("spmm_plan_build", 5);
    if (rc) return rc;
    int32_t tot[3] = {0, 0, 0};
    PLAGNN_CUDA_TRY(cudaMemcpyAsync(&tot[0], P + L.item_ptr + num_rows, 4, cudaMemcpyDeviceToHost, st));
    PLAGNN_CUDA_TRY(cudaMemcpyAsync(&tot[1], P + L.hub_ptr + num_rows, 4, cudaMemcpyDeviceToHost, st));
    PLAGNN_CUDA_TRY(cudaMemcpyAsync(&tot[2], P + L.slot_ptr + num_rows, 4, cudaMemcpyDeviceToHost, st));
    PLAGNN_CUDA_TRY(cudaStreamSynchronize(st));
    host_counts[0] = tot[0];
    host_counts[1] = tot[1];
    host_counts[2] = tot[2];
    return PLAGNN_OK;
}

}  // extern "C"

// Accessors for the other translation units (spmm.cu) — plan field offsets.
namespace plagnn {
void plan_pointers(const void* plan, int64_t n_rows, const int32_t** item_ptr, const int32_t** slot_ptr,
                   const int32_t** item_row, const int32_t** hub_rows) {
    const PlanLayout L = plan_layout(n_rows, 0, 32);
    const int32_t* P = (const int32_t*)plan;
    *item_ptr = P + L.item_ptr;
    *slot_ptr = P + L.slot_ptr;
    *item_row = P + L.item_row;
    *hub_rows = P + L.hub_rows;
}
}  // namespace plagnn
