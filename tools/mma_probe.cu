// Hardware probe for the K3 redesign (standalone; not part of libplagnn.so):
//   1. tcgen05.mma kind::tf32 issue rate with operands resident (no loads): cta_group 1/2, N 128/256, A from smem or TMEM
//   2. does the tensor core truncate or round the low 13 bits of an fp32 word read as tf32?
//   3. shared-memory image written by TMA for SWIZZLE_128B and SWIZZLE_128B_ATOM_32B boxes of fp32
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/mma_probe tools/mma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 2000000000ll) __trap();
    }
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    if (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {   // K-major, SWIZZLE_128B, SBO = 1024
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint32_t make_idesc(uint32_t m, uint32_t n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------
// 1. issue-rate probe
// ---------------------------------------------------------------------------------------------------
template <int CG, int N, bool TS>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    constexpr int BROWS = N / CG;
    constexpr int A_BYTES = 128 * 128, B_BYTES = BROWS * 128;
    constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;
    constexpr int STAGES = 2;
    __shared__ uint64_t bar_store;
    __shared__ uint32_t tmem_slot;
    const uint32_t bar = smem_u32(&bar_store);
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;

    float* f = reinterpret_cast<float*>(smem_raw + (tiles - raw));
    for (int i = t; i < STAGES * STAGE / 4; i += blockDim.x) f[i] = 1.0f + 1e-3f * (float)(i & 255);
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<CG>(smem_u32(&tmem_slot), 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (t == 32 && rank == 0) {
        const uint32_t idesc = make_idesc(128 * CG, N);
        const bool two_acc = (2 * N + (TS ? 64 : 0)) <= 512;
        const uint32_t acc0 = tmem_base, acc1 = two_acc ? tmem_base + N : tmem_base;
        const uint32_t a_tm = tmem_base + 512 - 64;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t st = tiles + (it % STAGES) * STAGE;
            const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + A_BYTES);
            const uint64_t b_hi = make_smem_desc(st + 2 * A_BYTES), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t on = (it | kk) ? 1u : 0u;
                if (TS) {
                    umma_ts(acc1, a_tm + 8 * kk, b_hi + 2 * kk, idesc, on);
                    umma_ts(acc1, a_tm + 32 + 8 * kk, b_lo + 2 * kk, idesc, 1u);
                    umma_ts(acc0, a_tm + 32 + 8 * kk, b_hi + 2 * kk, idesc, two_acc ? on : 1u);
                } else {
                    umma_ss<CG>(acc1, a_lo + 2 * kk, b_hi + 2 * kk, idesc, on);
                    umma_ss<CG>(acc1, a_hi + 2 * kk, b_lo + 2 * kk, idesc, 1u);
                    umma_ss<CG>(acc0, a_hi + 2 * kk, b_hi + 2 * kk, idesc, two_acc ? on : 1u);
                }
            }
        }
        umma_commit<CG>(bar);
        mbar_wait(bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    } else if (t == 32 && CG == 2) {
        mbar_wait(bar, 0);    // the multicast commit also arrives on the follower's barrier
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<CG>(tmem_base, 512); }
}

template <int CG, int N, bool TS>
static void run_rate(const char* name) {
    constexpr int BROWS = N / CG;
    const int smem = 2 * (2 * 128 * 128 + 2 * BROWS * 128) + 1024;
    auto k = rate_kernel<CG, N, TS>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = 148 / CG * CG;
    long long* d;
    CK(cudaMalloc(&d, grid * sizeof(long long)));
    CK(cudaMemset(d, 0, grid * sizeof(long long)));
    const int iters = 400;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best_ms = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CK(cudaEventRecord(e0));
        CK(cudaLaunchKernelEx(&cfg, k, iters, d));
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
    }
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < grid; i += CG) mx = h[i] > mx ? h[i] : mx;
    const double mmas = (double)iters * 12;
    const double flops_per_mma = 2.0 * 128 * CG * N * 8;
    const double tf = flops_per_mma * mmas * (grid / CG) / (best_ms * 1e-3) / 1e12;
    printf("rate %-28s cycles/MMA %.1f  (%.0f cyc total)  kernel %.3f ms  -> %.0f TFLOP/s tf32 incl. launch (%.0f fp32-equiv)\n",
           name, (double)mx / mmas, (double)mx, best_ms, tf, tf / 3);
    CK(cudaFree(d));
}

// ---------------------------------------------------------------------------------------------------
// 2. truncation vs rounding of tf32 operands
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) trunc_kernel(const float* a /*128x32*/, const float* b /*128x32*/, float* d /*128x128*/) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bar_store;
    __shared__ uint32_t tmem_slot;
    const uint32_t bar = smem_u32(&bar_store);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    uint8_t* base = smem_raw + (tiles - raw);
    for (int i = t; i < 128 * 8; i += 128) {      // (row, 16-byte chunk)
        const int row = i >> 3, c = i & 7;
        const uint32_t off = row * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<float4*>(base + off) = *reinterpret_cast<const float4*>(a + row * 32 + c * 4);
        *reinterpret_cast<float4*>(base + 16384 + off) = *reinterpret_cast<const float4*>(b + row * 32 + c * 4);
    }
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<1>(smem_u32(&tmem_slot), 128);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (t == 32) {
        const uint32_t idesc = make_idesc(128, 128);
        const uint64_t da = make_smem_desc(tiles), db = make_smem_desc(tiles + 16384);
        for (int kk = 0; kk < 4; ++kk) umma_ss<1>(tmem_base, da + 2 * kk, db + 2 * kk, idesc, kk ? 1u : 0u);
        umma_commit<1>(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    for (int cq = 0; cq < 4; ++cq) {
        uint32_t acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + cq * 32, acc);
        for (int j = 0; j < 32; ++j) d[(warp * 32 + lane) * 128 + cq * 32 + j] = __uint_as_float(acc[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tmem_base, 128); }
}

static float tf32_trunc(float x) { uint32_t b; memcpy(&b, &x, 4); b &= 0xFFFFE000u; memcpy(&x, &b, 4); return x; }
static float tf32_rn(float x) { uint32_t b; memcpy(&b, &x, 4); b = (b + 0x1000u) & 0xFFFFE000u; memcpy(&x, &b, 4); return x; }

static void run_trunc() {
    std::vector<float> a(128 * 32), b(128 * 32), d(128 * 128);
    srand(7);
    for (auto& v : a) v = 1.0f + (float)(rand() & 0xFFFFFF) / 16777216.0f;      // full 24-bit mantissas in [1,2)
    for (int r = 0; r < 128; ++r) for (int k = 0; k < 32; ++k) b[r * 32 + k] = (k == (r & 31)) ? 1.0f : 0.0f;   // selects A[:, r%32]
    float *da, *db, *dd;
    CK(cudaMalloc(&da, a.size() * 4)); CK(cudaMalloc(&db, b.size() * 4)); CK(cudaMalloc(&dd, d.size() * 4));
    CK(cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(trunc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34816));
    trunc_kernel<<<1, 128, 34816>>>(da, db, dd);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
    int n_trunc = 0, n_rn = 0, n_exact = 0, n = 0;
    for (int r = 0; r < 128; ++r) for (int c = 0; c < 128; ++c) {
        const float x = a[r * 32 + (c & 31)], got = d[r * 128 + c];
        ++n;
        if (got == tf32_trunc(x)) ++n_trunc;
        if (got == tf32_rn(x)) ++n_rn;
        if (got == x) ++n_exact;
    }
    printf("tf32 operand handling: %d outputs; == trunc(x): %d, == rn(x): %d, == x exactly: %d\n", n, n_trunc, n_rn, n_exact);
    printf("  sample: x=%.9g got=%.9g trunc=%.9g rn=%.9g\n", a[5], d[5], tf32_trunc(a[5]), tf32_rn(a[5]));
}

// ---------------------------------------------------------------------------------------------------
// 3. TMA shared-memory images
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int c0, int c1, int bytes, float* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bar_store;
    const uint32_t bar = smem_u32(&bar_store);
    float* f = reinterpret_cast<float*>(smem_raw + (tiles - raw));
    for (int i = threadIdx.x; i < bytes / 4; i += 32) f[i] = -1.0f;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    fence_proxy_async_smem();
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(tiles), "l"(&map), "r"(c0), "r"(c1), "r"(bar) : "memory");
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < bytes / 4; i += 32) out[i] = f[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void run_tma() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return; }
    EncodeFn enc = (EncodeFn)fn;
    const int ROWS = 70, COLS = 200, PITCH = 256;    // ragged extents: out-of-bounds parts must read as zero
    std::vector<float> h(ROWS * PITCH);
    for (int r = 0; r < ROWS; ++r) for (int c = 0; c < PITCH; ++c) h[r * PITCH + c] = (float)(r * 1000 + c);
    float *dsrc, *dout;
    CK(cudaMalloc(&dsrc, h.size() * 4)); CK(cudaMalloc(&dout, 65536));
    CK(cudaMemcpy(dsrc, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000));
    struct Case { const char* name; CUtensorMapSwizzle sw; int box0, box1, c0, c1; } cases[] = {
        {"SWIZZLE_128B box{32,64} at (0,0)", CU_TENSOR_MAP_SWIZZLE_128B, 32, 64, 0, 0},
        {"SWIZZLE_128B box{32,64} at (192,32) ragged", CU_TENSOR_MAP_SWIZZLE_128B, 32, 64, 192, 32},
        {"SWIZZLE_128B_ATOM_32B box{32,32} at (0,0)", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 32, 32, 0, 0},
        {"SWIZZLE_128B_ATOM_32B box{32,32} at (192,64) ragged", CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, 32, 32, 192, 64},
    };
    for (auto& cs : cases) {
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)COLS, (cuuint64_t)ROWS};
        cuuint64_t strides[1] = {(cuuint64_t)PITCH * 4};
        cuuint32_t box[2] = {(cuuint32_t)cs.box0, (cuuint32_t)cs.box1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dsrc, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         cs.sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("tma %s: encode failed %d\n", cs.name, (int)r); continue; }
        const int bytes = cs.box0 * cs.box1 * 4;
        tma_kernel<<<1, 32, 40000>>>(map, cs.c0, cs.c1, bytes, dout);
        CK(cudaDeviceSynchronize());
        std::vector<float> o(bytes / 4);
        CK(cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost));
        // expected images
        int bad_k = 0, bad_mn = 0;
        for (int r1 = 0; r1 < cs.box1; ++r1) for (int c = 0; c < cs.box0; ++c) {
            const int gr = cs.c1 + r1, gc = cs.c0 + c;
            const float want = (gr < ROWS && gc < COLS) ? (float)(gr * 1000 + gc) : 0.0f;
            // K-major SW128: row r1 at r1*128, 16-byte chunk (c>>2) ^ (r1&7)
            const int off_k = r1 * 128 + ((((c >> 2) ^ (r1 & 7))) << 4) + (c & 3) * 4;
            // SW128 with 32-byte atoms: row r1 at r1*128, 32-byte chunk (c>>3) ^ (r1&3)
            const int off_mn = r1 * 128 + ((((c >> 3) ^ (r1 & 3))) << 5) + (c & 7) * 4;
            if (o[off_k / 4] != want) ++bad_k;
            if (o[off_mn / 4] != want) ++bad_mn;
        }
        printf("tma %-52s mismatches vs K-major-SW128 formula: %d, vs ATOM_32B formula: %d\n", cs.name, bad_k, bad_mn);
        if (bad_k && bad_mn) {
            printf("  first 4 rows of the image:\n");
            for (int r1 = 0; r1 < 4; ++r1) { printf("   "); for (int c = 0; c < 32; ++c) printf(" %g", o[r1 * 32 + c]); printf("\n"); }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// 4. shared-memory contention: MMA issue rate (cta_group::2, 256x256x8, SS) while HW other warps stream LDS.128 + STS.128
//    through a separate 64 KB region (what an in-SM hi/lo split of TMA-loaded tiles would do)
// ---------------------------------------------------------------------------------------------------
template <int HW>
__global__ void __launch_bounds__(64 + 32 * (HW > 0 ? HW : 1), 1) contend_kernel(int iters, long long* cycles, long long* moved) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int CG = 2, N = 256;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    constexpr int A_BYTES = 128 * 128, B_BYTES = 128 * 128, STAGE = 2 * A_BYTES + 2 * B_BYTES, STAGES = 2;
    __shared__ uint64_t bar_store;
    __shared__ uint32_t tmem_slot;
    __shared__ volatile int done;
    const uint32_t bar = smem_u32(&bar_store);
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t rank = cluster_ctarank();
    float* f = reinterpret_cast<float*>(smem_raw + (tiles - raw));
    for (int i = t; i < (STAGES * STAGE + 65536) / 4; i += blockDim.x) f[i] = 1.0f + 1e-3f * (float)(i & 255);
    if (t == 0) { mbar_init(bar, 1); fence_mbar_init(); done = 0; }
    if (warp == 0) tmem_alloc<CG>(smem_u32(&tmem_slot), 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (t == 32 && rank == 0) {
        const uint32_t idesc = make_idesc(256, N);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t st = tiles + (it % STAGES) * STAGE;
            const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + A_BYTES);
            const uint64_t b_hi = make_smem_desc(st + 2 * A_BYTES), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t on = (it | kk) ? 1u : 0u;
                umma_ss<CG>(tmem_base + N, a_lo + 2 * kk, b_hi + 2 * kk, idesc, on);
                umma_ss<CG>(tmem_base + N, a_hi + 2 * kk, b_lo + 2 * kk, idesc, 1u);
                umma_ss<CG>(tmem_base, a_hi + 2 * kk, b_hi + 2 * kk, idesc, on);
            }
        }
        umma_commit<CG>(bar);
        mbar_wait(bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
        done = 1;
    } else if (t == 32) {
        mbar_wait(bar, 0);
        done = 1;
    } else if (warp >= 2 && HW > 0) {
        // hammer: each warp streams 16-byte vectors through the 64 KB region behind the stages
        const uint32_t region = tiles + STAGES * STAGE;
        const int hw = warp - 2, lane = t & 31;
        long long n = 0;
        uint32_t off = (uint32_t)((hw * 32 + lane) * 16);
        while (!done) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float4 v;
                const uint32_t a = region + ((off + u * HW * 512) & 32767u);
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
                v.x += 1.f;
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a + 32768u), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
            }
            n += 8;
            off += 4096;
        }
        if (lane == 0) atomicAdd((unsigned long long*)&moved[blockIdx.x], (unsigned long long)(n * 1024));   // LDS + STS bytes per warp
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<CG>(tmem_base, 512); }
}

template <int HW>
static void run_contend() {
    const int smem = 2 * 65536 + 65536 + 1024;
    auto k = contend_kernel<HW>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = 148;
    long long *d, *mv;
    CK(cudaMalloc(&d, grid * sizeof(long long))); CK(cudaMalloc(&mv, grid * sizeof(long long)));
    CK(cudaMemset(d, 0, grid * sizeof(long long))); CK(cudaMemset(mv, 0, grid * sizeof(long long)));
    const int iters = 400;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64 + 32 * (HW > 0 ? HW : 1)); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k, iters, d, mv));
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(grid), m(grid);
    CK(cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(m.data(), mv, grid * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; double bytes = 0; int nb = 0;
    for (int i = 0; i < grid; i += 2) { mx = h[i] > mx ? h[i] : mx; }
    for (int i = 0; i < grid; ++i) { bytes += (double)m[i]; ++nb; }
    printf("contend: %d hammer warps/CTA: cycles/MMA %.1f, hammer LDS+STS %.1f B/clk per SM (%.1f KB per 1536 cycles)\n", HW,
           (double)mx / (iters * 12.0), bytes / nb / (double)mx, bytes / nb / (double)mx * 1536 / 1024);
    CK(cudaFree(d)); CK(cudaFree(mv));
}

// ---------------------------------------------------------------------------------------------------
// 5. ring probe: cta_group::2 256x256x8 MMAs issued like the GEMM main loop does: 12 MMAs per k-block, one commit per
//    k-block onto a ring of RING barriers, the issuer waits for the commit of k-block (it - RING + 1) before issuing
//    k-block it + 1 (i.e. at most RING k-blocks of MMAs queued).  RING = 0: a single commit at the end.
// ---------------------------------------------------------------------------------------------------
template <int RING>
__global__ void __launch_bounds__(128, 1) ring_kernel(int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    constexpr int CG = 2, N = 256;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    constexpr int A_BYTES = 128 * 128, B_BYTES = 128 * 128, STAGE = 2 * A_BYTES + 2 * B_BYTES, STAGES = 3;
    __shared__ uint64_t bar_store[8];
    __shared__ uint32_t tmem_slot;
    const int t = threadIdx.x, warp = t >> 5;
    const uint32_t rank = cluster_ctarank();
    float* f = reinterpret_cast<float*>(smem_raw + (tiles - raw));
    for (int i = t; i < STAGES * STAGE / 4; i += blockDim.x) f[i] = 1.0f + 1e-3f * (float)((i * 2654435761u) >> 20);
    if (t == 0) { for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar_store[i]), 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<CG>(smem_u32(&tmem_slot), 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (t == 32 && rank == 0) {
        const uint32_t idesc = make_idesc(256, N);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (RING > 0 && it >= RING) {
                const int w = it - RING;
                mbar_wait(smem_u32(&bar_store[w % RING]), (uint32_t)((w / RING) & 1));
                tc_fence_after();
            }
            const uint32_t st = tiles + (it % STAGES) * STAGE;
            const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + A_BYTES);
            const uint64_t b_hi = make_smem_desc(st + 2 * A_BYTES), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t on = (it | kk) ? 1u : 0u;
                umma_ss<CG>(tmem_base + N, a_lo + 2 * kk, b_hi + 2 * kk, idesc, on);
                umma_ss<CG>(tmem_base + N, a_hi + 2 * kk, b_lo + 2 * kk, idesc, 1u);
                umma_ss<CG>(tmem_base, a_hi + 2 * kk, b_hi + 2 * kk, idesc, on);
            }
            if (RING > 0) umma_commit<CG>(smem_u32(&bar_store[it % RING]));
        }
        umma_commit<CG>(smem_u32(&bar_store[7]));
        mbar_wait(smem_u32(&bar_store[7]), 0);
        cycles[blockIdx.x] = clock64() - t0;
    } else if (t == 32) {
        mbar_wait(smem_u32(&bar_store[7]), 0);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<CG>(tmem_base, 512); }
}

template <int RING>
static void run_ring(int iters) {
    const int smem = 3 * 65536 + 1024;
    auto k = ring_kernel<RING>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int grid = 148;
    long long* d;
    CK(cudaMalloc(&d, grid * sizeof(long long)));
    CK(cudaMemset(d, 0, grid * sizeof(long long)));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaLaunchKernelEx(&cfg, k, 10, d));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, k, iters, d));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), d, grid * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < grid; i += 2) mx = h[i] > mx ? h[i] : mx;
    printf("ring %d, %d k-blocks: cycles/MMA %.1f, %.3f ms -> SM clock ~%.0f MHz, %.0f TFLOP/s tf32\n", RING, iters,
           (double)mx / (iters * 12.0), ms, (double)mx / ms / 1e3, 2.0 * 256 * 256 * 8 * 12.0 * iters * 74 / (ms * 1e-3) / 1e12);
    CK(cudaFree(d));
}

int main(int argc, char** argv) {
    const char* which = argc > 1 ? argv[1] : "all";
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    if (!strcmp(which, "all") || !strcmp(which, "trunc")) run_trunc();
    if (!strcmp(which, "all") || !strcmp(which, "tma")) run_tma();
    if (!strcmp(which, "all") || !strcmp(which, "rate")) {
        run_rate<1, 128, false>("SS cta_group::1 128x128x8");
        run_rate<1, 256, false>("SS cta_group::1 128x256x8");
        run_rate<1, 128, true>("TS cta_group::1 128x128x8");
        run_rate<1, 256, true>("TS cta_group::1 128x256x8");
        run_rate<2, 128, false>("SS cta_group::2 256x128x8");
        run_rate<2, 256, false>("SS cta_group::2 256x256x8");
    }
    if (!strcmp(which, "all") || !strcmp(which, "ring")) {
        run_ring<0>(400); run_ring<1>(400); run_ring<2>(400); run_ring<3>(400); run_ring<0>(20000); run_ring<3>(20000); run_ring<3>(200000);
    }
    if (!strcmp(which, "all") || !strcmp(which, "contend")) {
        run_contend<0>(); run_contend<1>(); run_contend<2>(); run_contend<4>(); run_contend<8>(); run_contend<16>();
    }
    return 0;
}
