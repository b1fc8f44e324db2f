// K3 (second generation) — dense feature x weight contraction: TMA-fed tcgen05 with CTA pairs, sm_100a.
//
// Replaces the cuBLAS SGEMMs behind nn.Linear in code/model.py:16-17,20-28 (fc_pool / fc_self / fc_neigh inside
// SAGEConv, liner1, liner2) and their autograd backward.  gemm_tc.cu (first generation, operands split by SIMT loader
// warps) stays as backend PLAGNN_GEMM_TCGEN05 for cross-checks.
//
// fp32-level accuracy on TF32 tensor cores (3xTF32):  x = hi + lo,  x*y ~ hi*hi + lo*hi + hi*lo.
// Measured on B200 (tools/mma_probe.cu): the tensor core TRUNCATES an fp32 word to tf32 (drops the low 13 bits), so the
// raw fp32 tile IS the hi operand and the only derived data is  lo = rn_tf32(x - trunc_tf32(x)).
// Raw tiles are brought in by TMA straight into the swizzled layouts the UMMA descriptors read:
//     k-contiguous operand  : box {32 k, 128 rows}, SWIZZLE_128B                  (K-major descriptor)
//     mn-contiguous operand : 4 boxes {32 mn, 32 k}, SWIZZLE_128B_ATOM_32B         (MN-major descriptor)
// (out-of-bounds parts of a box read as zero, which handles every K / M / N tail) and eight warps derive the lo tiles
// in shared memory, elementwise, into a second ring.  Measured: tensor-core operand reads do not compete with LDS/STS
// for shared-memory bandwidth (the MMA rate is unchanged with 117 B/clk of LDS+STS traffic next to it).
// A variant that read precomputed lo matrices from global memory doubled L2 and HBM traffic and was dropped.
//
// CTA pair (cta_group::2, cluster 2x1x1): one 256 x 256 output tile per pair.  Each CTA stages its own 128 rows of A
// and its own half (128 rows) of B; the leader CTA issues 256x256x8 MMAs (128 cycles each, the full tf32 rate) that
// read both CTAs' shared memory; the accumulators (hi*hi in TMEM columns [0,256), lo*hi + hi*lo in [256,512)) live in
// each CTA's tensor memory for its 128 rows.  Per CTA and k-block of 32: 32 KB by TMA for 12 MMAs = 1536 cycles,
// i.e. 21 B/clk from L2 (a 128 x 128 single-CTA tile needs 43).  CG = 1 (128 x 128, one CTA) is kept as a cross-check.
//
// Warps: 0 = TMA producer, 1 = TMEM allocator + MMA issuer, 2..9 = hi/lo split, then epilogue
// (TMEM -> registers -> swizzled shared-memory image -> TMA store: whole 128-byte lines, tails clipped by the map).
#include "common.cuh"
#include <cuda.h>
// Diagnostics (per-CTA clock64 trace, PLAGNN_TMA_DEBUG timing modes, the single-accumulator experiment) are compiled in
// only with -DPLAGNN_TMA_DIAG=1 (`PLAGNN_TMA_DIAG=1 python pla-gnn_b200/csrc/build.py --force`): they cost registers and
// instruction-cache space in the production kernel.
#ifndef PLAGNN_TMA_DIAG
#define PLAGNN_TMA_DIAG 0
#endif
#if PLAGNN_TMA_DIAG
#define TM_DEBUG(P) ((P).debug)
#define TM_SINGLE(P) ((P).single_acc)
#else
#define TM_DEBUG(P) 0
#define TM_SINGLE(P) 0
#endif
#include <cstdlib>
#include <cstring>
#include <climits>
#include <unordered_map>
#include <string>

namespace plagnn {

constexpr int TM_BK = 32;
constexpr int TM_PART_BYTES = 128 * TM_BK * 4;           // 16 KB: 128 rows x 32 fp32
constexpr int TM_RAW_BYTES = 2 * TM_PART_BYTES;          // one ring stage: the A tile and the B tile of one k-block
constexpr int TM_RAW_STAGES = 4, TM_LO_STAGES = 3;       // raw fp32 tiles (TMA) / derived lo tiles
constexpr int TM_SMEM_BYTES = (TM_RAW_STAGES + TM_LO_STAGES) * TM_RAW_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TM_EPI_WARPS = 8;
constexpr int TM_THREADS = (2 + TM_EPI_WARPS) * 32;
constexpr int TM_MAX_CHAIN = 40;    // k-blocks per TMEM accumulation chain (see gemm_tc.cu: truncating accumulate)
// Long contractions (the weight gradients: K = number of nodes) can be cut into shorter chains (PLAGNN_TMA_LONG_CHAIN).  The
// tensor core truncates when it adds into the fp32 accumulator, so one chain's error grows with its length.  Measured at
// K = 24 041 on shared decisions (tests/test_gpu_model.py, full size, worst of the 19 gradient tensors against the fp32
// oracle): chains of 40 / 16 / 8 k-blocks -> 1.49e-5 / 1.23e-5 / 1.21e-5, and the PPI epoch 2.25 / 2.44 ms for 40 / 16 (more
// partials to write and fold).  What remains comes from the FORWARD products (K = 503 / 1006, one chain each), whose
// truncation bias the gradient sums amplify ~3.5x — so the default stays at TM_MAX_CHAIN.
constexpr int TM_LONG_K_BLOCKS = 64;   // contractions of at least this many k-blocks (K >= 2048) ...
constexpr int TM_LONG_CHAIN = 40;      // ... are cut into chains of at most this many

struct alignas(64) TmParams {
    CUtensorMap map[PLAGNN_GEMM_MAX_PAIRS][2];    // [pair][A, B]
    CUtensorMap map_out;                          // C (or the split-K partials) as {n, m, splits}, box {32, 32, 1}
    CUtensorMap map_gate;                         // gate as {n, m}, box {32, 32} (gate_tma = 1)
    CUtensorMap map_b64[PLAGNN_GEMM_MAX_PAIRS];   // k-contiguous B with a 64-row box (the 256 x 128 tiles of gemm_tma_db_kernel)
    int gate_tma;
    int tma_store;                                // 1: epilogue leaves through shared memory + TMA stores (aligned output)
    int64_t ldp;                                  // row pitch of the split-K partials
    int ones_col;                                 // >= 0: column of the (mn-contiguous) B operand that reads as 1.0: output column
                                                  // ones_col = sum over k of A, i.e. the bias gradient of a weight-gradient product
    float* ones_out;                              // where the reduction writes that column (length m)
    int tiles_m, tiles_n, total_tiles;            // tile t = (split * tiles_m + row block) * tiles_n + column block
    int persistent;                               // fewer CTA pairs than tiles: each pair walks several tiles
    int transpose_out;                            // split-K only: the reduction writes C^T (operands were swapped by the launcher)
    int64_t m, n;
    int npairs;
    int kblocks[PLAGNN_GEMM_MAX_PAIRS];
    int total_kblocks, kblocks_per_split, splits;
    const float* bias;
    int act;
    float slope;
    const float* gate;
    int64_t ldg;
    int gate_act;
    float* c;
    int64_t ldc;
    float* partial;
    int single_acc;     // PLAGNN_TMA_SINGLE_ACC=1 (experiment): corrections accumulate into the main accumulator
    long long* trace;   // PLAGNN_TMA_TRACE: per-CTA clock64 stamps (32 per CTA, first 64 CTAs), else null
    int debug;   // PLAGNN_TMA_DEBUG (timing experiments, results are garbage): 1 = TMA loads only for the first ring pass,
                 // 2 = no MMAs, 3 = no stores in the epilogue, 4 = shared-memory images written but not stored
};

// ---- PTX wrappers -----------------------------------------------------------------------------
namespace tm {
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// arrive on a barrier given by its shared::cluster address (own or peer CTA).  Default semantics (release at CTA scope),
// as CUTLASS' ClusterBarrier::arrive does: `.release.cluster` compiles to MEMBAR.ALL.GPU, which cost ~1 500 cycles per
// k-block on the critical path.  What the peer's tensor core reads was made visible to the async proxy by
// fence.proxy.async before this arrive.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// one lane of a converged warp.  The issuing warps run their loops with all 32 lanes (warp-uniform control flow and
// operands, which the compiler keeps in uniform registers) and predicate only the tcgen05 / TMA instructions with this:
// inside an `if (lane == 0)` region every UTCHMMA / UTMALDG cost an ELECT + 6 x R2UR.BROADCAST + branch loop (12-17
// instructions per MMA, i.e. the single issuing thread became a co-bottleneck of the main loop).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}

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
// arrive on `bar` (same shared offset in every CTA of the pair) when all MMAs issued so far have completed
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// 2-D tiled TMA load into this CTA's shared memory, completion bytes on this CTA's barrier
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// bulk tensor store of one {32 cols, 32 rows, 1} box from shared memory; out-of-bounds elements are not written
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
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
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr) {
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// MN-major operand as four TMA boxes {32 mn, 32 k} with SWIZZLE_128B_ATOM_32B: swizzle atom = 4 k-rows x 128 bytes;
// the atoms of one box (k-groups of 4) are 512 bytes apart (SBO), boxes (32 mn each) 4096 bytes apart (LBO).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t saddr) {
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((4096u >> 4) << 16);
    const uint32_t hi = (512u >> 4) | (1u << 14) | (1u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// kind::tf32, D = f32; bit 15 / 16 = A / B stored MN-major
__device__ __forceinline__ uint32_t make_idesc(uint32_t m, uint32_t n, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | ((n >> 3) << 17) |
           ((m >> 4) << 24);
}
}  // namespace tm

// lo = rn_tf32(x - trunc_tf32(x)) : what the tensor core does not see of x, on the tf32 grid (finite inputs)
__device__ __forceinline__ float tf32_lo(float x) {
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    return __uint_as_float((__float_as_uint(x - hi) + 0x1000u) & 0xFFFFE000u);
}

__device__ __forceinline__ float tm_epilogue_one(const TmParams& P, float v, int64_t r, int64_t c) {
    if (P.bias) v += __ldg(P.bias + c);
    v = apply_act(v, P.act, P.slope);
    if (P.gate) v *= act_grad_from_output(__ldg(P.gate + r * P.ldg + c), P.gate_act, P.slope);
    return v;
}

// One output tile of the launch: tile index t -> (split, row block, column block), column blocks fastest so that the
// column tiles of one row block run together and A streams from HBM once.
struct TmTile {
    int64_t m0, n0;
    int split, kb_beg, nkb;
    uint32_t n_eff, n_half;      // MMA width of this tile / rows of B one CTA stages
};
template <int CG>
__device__ __forceinline__ TmTile tm_tile(const TmParams& P, int t) {
    constexpr int TILE = 128 * CG;
    TmTile T;
    const int nt = t % P.tiles_n;
    const int rest = t / P.tiles_n;
    const int mt = rest % P.tiles_m;
    T.split = rest / P.tiles_m;
    T.m0 = (int64_t)mt * TILE;
    T.n0 = (int64_t)nt * TILE;
    T.kb_beg = T.split * P.kblocks_per_split;
    T.nkb = min(P.total_kblocks, T.kb_beg + P.kblocks_per_split) - T.kb_beg;
    // columns of this tile that exist, rounded to what one MMA can produce (CG = 2: both halves multiples of 32)
    const int64_t nrem = P.n - T.n0;
    T.n_eff = nrem >= TILE ? (uint32_t)TILE : (uint32_t)((nrem + 32 * CG - 1) / (32 * CG) * (32 * CG));
    T.n_half = T.n_eff / CG;
    return T;
}

// The kernel walks a contiguous range of tiles per CTA pair.  With one
// tile per pair that is the plain tiled launch; with fewer pairs than tiles (P.persistent) the rings simply run on across
// tile boundaries: while warps 2..9 read out the accumulators of tile i, the producer already fills the raw ring with
// tile i + 1, and set-up / tear-down are paid once per SM.  TMEM holds one tile (2 x 256 columns), so the MMAs of the next
// tile start when the read-out has released it (bar_tmem_empty).
template <int CG, bool AT, bool BT>
__global__ void __launch_bounds__(TM_THREADS, 1) gemm_tma_kernel(const __grid_constant__ TmParams P) {
    using namespace tm;
    pdl_trigger();      // the next grid may be placed as SMs free up; it holds at its own pdl_wait() until this one is done
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                  // 1024-byte aligned (swizzle atoms)
    const uint32_t lo_ring = tiles + TM_RAW_STAGES * TM_RAW_BYTES;
    const uint32_t bars = lo_ring + TM_LO_STAGES * TM_RAW_BYTES;
    const uint32_t bar_raw_full = bars, bar_raw_empty = bars + 8 * TM_RAW_STAGES;
    const uint32_t bar_lo_full = bars + 16 * TM_RAW_STAGES, bar_lo_empty = bar_lo_full + 8 * TM_LO_STAGES;
    const uint32_t bar_acc = bar_lo_empty + 8 * TM_LO_STAGES;
    const uint32_t bar_tmem_empty = bar_acc + 8;                    // accumulators read out (both CTAs): next tile may overwrite
    const uint32_t bar_gate = bar_tmem_empty + 8;                   // [TM_EPI_WARPS]: gate images of one warp landed
    const uint32_t tmem_slot = bar_gate + 8 * TM_EPI_WARPS;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

    constexpr int TILE_M = 128 * CG, TILE_N = 128 * CG;
    constexpr uint32_t TMEM_COLS = 2 * TILE_N;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    // this CTA pair's tiles: a contiguous, balanced range (2 or 3 tiles when 188 tiles meet 74 pairs).  Contiguous tiles
    // alternate over the column blocks, so every pair gets the same mix of full and ragged tiles and re-reads its rows of
    // A while they are hot in L2 (a strided assignment gave all full-width tiles to the even pairs).
    const int unit = blockIdx.x / CG, units = gridDim.x / CG;
    const int first_tile = (int)((int64_t)unit * P.total_tiles / units);
    const int end_tile = (int)((int64_t)(unit + 1) * P.total_tiles / units);
#if PLAGNN_TMA_DIAG
    long long* const tr = (P.trace && blockIdx.x < 64) ? P.trace + 32 * blockIdx.x : nullptr;   // stamps of the FIRST tile
#else
    long long* const tr = nullptr;
#endif
    if (tr && t == 0) tr[0] = clock64();

    // raw tiles of k-block `it` of tile T into ring stage g % TM_RAW_STAGES (g = k-blocks loaded so far; thread 0 only)
    auto load_kblock = [&](const TmTile& T, int it, int g) {
        const int s = g % TM_RAW_STAGES;
        int p = 0, local = T.kb_beg + it;
        if (P.npairs > 1 && local >= P.kblocks[0]) { local -= P.kblocks[0]; p = 1; }
        const int k0 = local * TM_BK;
        const int a_row = (int)(T.m0 + rank * 128), b_row = (int)(T.n0 + rank * T.n_half);
        const uint32_t st = tiles + s * TM_RAW_BYTES;
        const uint32_t rb = bar_raw_full + 8 * s;
        mbar_expect_tx(rb, TM_RAW_BYTES);
        if (!AT) {
            tma_load_2d(st, &P.map[p][0], k0, a_row, rb);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(st + j * 4096, &P.map[p][0], a_row + 32 * j, k0, rb);
        }
        if (!BT) {
            tma_load_2d(st + TM_PART_BYTES, &P.map[p][1], k0, b_row, rb);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) tma_load_2d(st + TM_PART_BYTES + j * 4096, &P.map[p][1], b_row + 32 * j, k0, rb);
        }
    };
    const TmTile T0 = tm_tile<CG>(P, first_tile);
    const int preloaded = T0.nkb < TM_RAW_STAGES ? T0.nkb : TM_RAW_STAGES;

    if (warp == 0) {
        if (elect_one()) {
            for (int s = 0; s < TM_RAW_STAGES; ++s) {
                mbar_init(bar_raw_full + 8 * s, 1);
                mbar_init(bar_raw_empty + 8 * s, 1);
            }
            for (int s = 0; s < TM_LO_STAGES; ++s) {
                mbar_init(bar_lo_full + 8 * s, (uint32_t)(CG * TM_EPI_WARPS));   // one arrival per splitting warp of the pair
                mbar_init(bar_lo_empty + 8 * s, 1);
            }
            mbar_init(bar_acc, 1);
            mbar_init(bar_tmem_empty, (uint32_t)(CG * TM_EPI_WARPS));
            for (int w = 0; w < TM_EPI_WARPS; ++w) mbar_init(bar_gate + 8 * w, 1);
            fence_mbar_init();
#pragma unroll
            for (int p = 0; p < PLAGNN_GEMM_MAX_PAIRS; ++p)
                if (p < P.npairs) { prefetch_map(&P.map[p][0]); prefetch_map(&P.map[p][1]); }
            if (P.tma_store) prefetch_map(&P.map_out);
            if (P.gate_tma) prefetch_map(&P.map_gate);
        }
        __syncwarp();
        pdl_wait();     // barrier init and descriptor prefetch above overlap the previous grid's tail; operands are read below
        // the first ring pass is requested right here, before the TMEM allocation and the CTA / cluster barriers: the loads
        // only need this CTA's own (just initialised) barriers, and their ~3 000-cycle latency overlaps the rest of the set-up
        for (int it = 0; it < preloaded; ++it) {
            if (elect_one()) load_kblock(T0, it, it);
            __syncwarp();
        }
    }
    if (warp == 1) tmem_alloc<CG>(tmem_slot, TMEM_COLS);
    pdl_wait();         // every thread, before bias / gate reads and output stores
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    if (tr && t == 0) tr[1] = clock64();

    if (warp == 0) {
        // ================= TMA producer (whole warp, loads issued by one elected lane): raw fp32 tiles of A and B =========
        {
            if (tr && lane == 0) tr[16] = clock64();
            int g = 0;                                   // k-blocks requested so far (all tiles)
            for (int tile = first_tile; tile < end_tile; ++tile) {
                const TmTile T = tm_tile<CG>(P, tile);
                for (int it = (tile == first_tile ? preloaded : 0); it < T.nkb; ++it) {
                    const int gg = g + it;
                    const int s = gg % TM_RAW_STAGES;
                    const uint32_t ph = (uint32_t)((gg / TM_RAW_STAGES) & 1);
                    const long long w0 = tr ? clock64() : 0;
                    mbar_wait(bar_raw_empty + 8 * s, ph ^ 1u);
                    if (tr && lane == 0 && tile == first_tile) tr[8] += clock64() - w0;
                    if (tr && lane == 0 && tile == first_tile && it == 8) tr[20] = clock64();
                    if (elect_one()) load_kblock(T, it, gg);
                    __syncwarp();
                }
                g += T.nkb;
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA; whole warp, instructions issued by one elected lane) =================
        if (rank == 0) {
            // descriptor start-address step per K = 8: 32 bytes inside a K-major row; two k-groups of 512 bytes if MN-major
            constexpr uint64_t a_step = AT ? (1024u >> 4) : 2u, b_step = BT ? (1024u >> 4) : 2u;
            const uint32_t acc_main = tmem_base, acc_corr = tmem_base + TILE_N;
            int g = 0, tile_iter = 0;
            for (int tile = first_tile; tile < end_tile; ++tile, ++tile_iter) {
                const TmTile T = tm_tile<CG>(P, tile);
                const uint32_t idesc = make_idesc(TILE_M, T.n_eff, AT, BT);
                if (tile_iter > 0) {                     // the previous tile's accumulators have been read out in both CTAs
                    mbar_wait(bar_tmem_empty, (uint32_t)((tile_iter - 1) & 1));
                    tc_fence_after();
                }
                for (int it = 0; it < T.nkb; ++it, ++g) {
                    const int s = g % TM_RAW_STAGES, l = g % TM_LO_STAGES;
                    const uint32_t phl = (uint32_t)((g / TM_LO_STAGES) & 1);
                    const long long w0 = tr ? clock64() : 0;
                    // the lo tiles of both CTAs are written (which also means the raw tiles have landed in both)
                    mbar_wait(bar_lo_full + 8 * l, phl);
                    tc_fence_after();
                    if (tr && lane == 0 && tile_iter == 0) {
                        const long long w1 = clock64();
                        tr[9] += w1 - w0;
                        if (it == 0) tr[2] = w1;
                        if (it == 0 || it == 8) tr[it ? 23 : 19] = w1;
                    }
                    if (tr && lane == 0 && tile_iter == 1 && it == 0) tr[30] = clock64() - tr[3];   // tile 0 committed -> tile 1's first k-block ready
                    const uint32_t st = tiles + s * TM_RAW_BYTES, sl = lo_ring + l * TM_RAW_BYTES;
                    const uint64_t a_hi = AT ? desc_mnmajor(st) : desc_kmajor(st);
                    const uint64_t a_lo = AT ? desc_mnmajor(sl) : desc_kmajor(sl);
                    const uint64_t b_hi = BT ? desc_mnmajor(st + TM_PART_BYTES) : desc_kmajor(st + TM_PART_BYTES);
                    const uint64_t b_lo = BT ? desc_mnmajor(sl + TM_PART_BYTES) : desc_kmajor(sl + TM_PART_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < TM_BK / 8; ++kk) {
                            if (TM_DEBUG(P) == 2) break;
                            const uint64_t adv_a = (uint64_t)kk * a_step, adv_b = (uint64_t)kk * b_step;
                            // corrections go to their own accumulator: the tensor core adds into TMEM with truncation, and the
                            // main accumulator then sees a third of the additions (gemm_tc.cu has the measurements)
                            const uint32_t acc_on = (it | kk) ? 1u : 0u;
                            if (TM_SINGLE(P)) {
                                umma_tf32<CG>(acc_main, a_hi + adv_a, b_hi + adv_b, idesc, acc_on);
                                umma_tf32<CG>(acc_main, a_lo + adv_a, b_hi + adv_b, idesc, 1u);
                                umma_tf32<CG>(acc_main, a_hi + adv_a, b_lo + adv_b, idesc, 1u);
                                continue;
                            }
                            umma_tf32<CG>(acc_corr, a_lo + adv_a, b_hi + adv_b, idesc, acc_on);
                            umma_tf32<CG>(acc_corr, a_hi + adv_a, b_lo + adv_b, idesc, 1u);
                            umma_tf32<CG>(acc_main, a_hi + adv_a, b_hi + adv_b, idesc, acc_on);
                        }
                        // both rings are released (in both CTAs) when the MMAs issued so far have read them
                        umma_commit<CG>(bar_lo_empty + 8 * l);
                        umma_commit<CG>(bar_raw_empty + 8 * s);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit<CG>(bar_acc);       // accumulators of this tile complete (both CTAs)
                __syncwarp();
                if (tr && lane == 0 && tile_iter == 0) tr[3] = clock64();
                if (tr && lane == 0 && tile_iter == 1) tr[31] = clock64() - tr[3] - tr[30];        // main loop of tile 1
            }
        }
        __syncwarp();
    } else {
        // ================= warps 2..9: hi/lo split in shared memory, then the read-out of the tile =================
        // The tensor core reads the raw fp32 tile as the hi operand (it drops the low 13 bits itself); the lo tile has
        // the same swizzled layout, so the pass is purely elementwise: lo[i] = rn_tf32(x[i] - trunc_tf32(x[i])).
        // Tensor-core operand reads do not compete with LDS/STS for bandwidth (tools/mma_probe.cu, "contend").
        // The raw ring is 4 deep (TMA latency ~2.5 us), the lo ring 3 deep: the chain "MMAs done -> commit -> split warps
        // -> STS -> proxy fence -> (remote) arrive -> issuer" is longer than one k-block of MMAs (1536 cycles).
        const int ct = t - 64;                                  // 0..255
        const uint32_t lo_full0 = CG == 2 ? mapa(bar_lo_full, 0) : bar_lo_full;   // the leader's barriers count both CTAs
        const uint32_t tmem_empty0 = CG == 2 ? mapa(bar_tmem_empty, 0) : bar_tmem_empty;
        const int lg = warp & 3;                        // TMEM lane group this warp may read (warp % 4)
        const int chalf = (warp - 2) >> 2;              // column half of the tile
        const bool direct = P.splits == 1;
        constexpr int CHUNKS = TILE_N / 32 / 2;
        // staging images of the read-out: in the raw ring for a single tile per CTA (everything is free by then); with
        // several tiles per CTA the raw ring is being refilled for the next tile, so they live in the lo ring, which is
        // idle until these same warps start splitting again (no gate images there: P.persistent excludes a gate)
        const uint32_t stg = (P.persistent ? lo_ring : tiles) + (uint32_t)(warp - 2) * 8192u;      // two 4 KB images per warp
        const uint32_t gimg = tiles + 65536u + (uint32_t)(warp - 2) * 16384u;                      // CHUNKS x 4 KB per warp
        const uint32_t gbar = bar_gate + 8 * (uint32_t)(warp - 2);
        const uint32_t row_off = (uint32_t)lane * 128u;
        int g = 0, tile_iter = 0, buf = 0;
        for (int tile = first_tile; tile < end_tile; ++tile, ++tile_iter) {
            const TmTile T = tm_tile<CG>(P, tile);
            const int64_t m0 = T.m0, n0 = T.n0;
            const uint32_t n_eff = T.n_eff;
            const int split = T.split;
            // does this thread own the piece of the B tile that holds the ones column?  MN-major box j = 32 columns,
            // k-row ct / 8, 16-byte piece ct % 8 of the row = logical 32-byte chunk ((ct % 8) / 2) ^ (row & 3), half (ct % 8) & 1
            int ones_piece = -1, ones_elem = 0;
            if (BT && P.ones_col >= 0) {
                const int64_t lc = (int64_t)P.ones_col - (n0 + rank * T.n_half);     // column inside this CTA's 128-wide B tile
                if (lc >= 0 && lc < 128) {
                    const int q = ct & 7, rr = (ct >> 3) & 3;
                    const int cbase = (((q >> 1) ^ rr) << 3) + ((q & 1) << 2);
                    const int c = (int)(lc & 31);
                    if (c >= cbase && c < cbase + 4) { ones_piece = (int)(lc >> 5); ones_elem = c - cbase; }
                }
            }
            for (int it = 0; it < T.nkb; ++it, ++g) {
                const int s = g % TM_RAW_STAGES, l = g % TM_LO_STAGES;
                const uint32_t phr = (uint32_t)((g / TM_RAW_STAGES) & 1), phl = (uint32_t)((g / TM_LO_STAGES) & 1);
                const long long w0 = (tr && t == 64) ? clock64() : 0;
                mbar_wait(bar_raw_full + 8 * s, phr);
                if (tr && t == 64 && tile_iter == 0) {
                    const long long w1 = clock64();
                    tr[10] += w1 - w0;
                    if (it == 0 || it == 8) tr[it ? 21 : 17] = w1;
                }
                const uint32_t src = tiles + s * TM_RAW_BYTES + (uint32_t)ct * 16u;
                const uint32_t dstl = lo_ring + l * TM_RAW_BYTES + (uint32_t)ct * 16u;
                float4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(src + (uint32_t)i * 4096u));
                if (BT && ones_piece >= 0) {
                    // bias-gradient column: this thread owns the 16-byte piece of B's box `ones_piece` that holds column
                    // ones_col of k-row ct / 8: it reads as 1.0 (rows beyond K meet zero-filled rows of A, so no k test)
#pragma unroll
                    for (int i = 4; i < 8; ++i)
                        if (i - 4 == ones_piece) {
                            if (ones_elem == 0) v[i].x = 1.0f;
                            else if (ones_elem == 1) v[i].y = 1.0f;
                            else if (ones_elem == 2) v[i].z = 1.0f;
                            else v[i].w = 1.0f;
                            sts_v4(src + (uint32_t)i * 4096u, v[i].x, v[i].y, v[i].z, v[i].w);
                        }
                }
                mbar_wait(bar_lo_empty + 8 * l, phl ^ 1u);      // the MMAs of k-block g - TM_LO_STAGES have read this lo stage
                if (tr && t == 64 && tile_iter == 0 && it == 8) tr[24] = clock64();
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    sts_v4(dstl + (uint32_t)i * 4096u, tf32_lo(v[i].x), tf32_lo(v[i].y), tf32_lo(v[i].z), tf32_lo(v[i].w));
                fence_proxy_async_smem();
                __syncwarp();
                if (tr && t == 64 && tile_iter == 0 && (it == 0 || it == 8)) tr[it ? 22 : 18] = clock64();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_cluster(lo_full0 + 8 * l);
                    else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_lo_full + 8 * l) : "memory");
                }
            }

            // ---------------- read-out: TMEM -> registers -> swizzled image -> TMA store ----------------
            mbar_wait(bar_acc, (uint32_t)(tile_iter & 1));
            tc_fence_after();
            if (tr && t == 64 && tile_iter == 0) tr[4] = clock64();
            const long long epi1 = (tr && t == 64 && tile_iter == 1) ? clock64() : 0;
            const int64_t r = m0 + rank * 128 + lg * 32 + lane;
            // Each warp turns its 32 rows x 32 columns into a 4 KB SWIZZLE_128B image and one lane stores it with a bulk
            // tensor copy: whole 128-byte lines leave the SM, rows >= m / columns >= n are clipped by the tensor map; two
            // images in flight per warp.  The activation / gate switches sit OUTSIDE the element loops (the first version
            // branched per element and spent 25 000 of its 28 000 epilogue cycles per tile fetching instructions), and the
            // loops are fully unrolled so that the 32 outputs of a thread stay in registers.
            const int row0 = (int)(m0 + rank * 128 + lg * 32);
            const int64_t r_ld = r < P.m ? r : P.m - 1;                  // clamped row for gate loads (clipped rows are never stored)
            // gate tiles (saved activations of the backward epilogues) come in by TMA as 32 x 32 SWIZZLE_128B images, all
            // chunks of this warp at once, while the first accumulator chunk is read
            const bool gate_img_on = direct && P.gate && P.gate_tma;
            if (gate_img_on && lane == 0) {
                int nch = 0;
                for (int ch = 0; ch < CHUNKS; ++ch) {
                    const int cb = (chalf * CHUNKS + ch) * 32;
                    if (cb < (int)n_eff && n0 + cb < P.n) ++nch;
                }
                if (nch) {
                    mbar_expect_tx(gbar, (uint32_t)nch * 4096u);
                    for (int ch = 0; ch < nch; ++ch)
                        tma_load_2d(gimg + (uint32_t)ch * 4096u, &P.map_gate, (int)(n0 + (chalf * CHUNKS + ch) * 32), row0, gbar);
                }
            }
            float* dst = direct ? P.c : P.partial + (int64_t)split * P.m * P.ldp;
            const int64_t ldd = direct ? P.ldc : P.ldp;
#pragma unroll 1
            for (int ch = 0; ch < CHUNKS; ++ch) {
                const int cbase = (chalf * CHUNKS + ch) * 32;
                if (cbase >= (int)n_eff || n0 + cbase >= P.n) break;     // warp-uniform
                const int64_t c0 = n0 + cbase;
                const int ncol = (int)((P.n - c0) < 32 ? (P.n - c0) : 32);   // valid columns of this chunk
                float v[32];
                {
                    uint32_t acc[32], acc_small[32];
                    const uint32_t ta = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)cbase;
                    const long long e0 = (tr && t == 64) ? clock64() : 0;
                    tmem_ld32(ta, acc);
                    tmem_ld32(ta + TILE_N, acc_small);
                    tmem_ld_wait();
                    if (tr && t == 64 && tile_iter == 0) tr[13] += clock64() - e0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) + (TM_SINGLE(P) ? 0.f : __uint_as_float(acc_small[j]));
                    if (tr && t == 64 && tile_iter == 0) {
                        // the loads are only guaranteed complete where their registers are first read: keep the sum alive
                        float keep = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) keep += v[j];
                        asm volatile("" ::"f"(keep));
                        tr[25] += clock64() - e0;
                    }
                }
                if (direct) {
                    if (P.bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += __ldg(P.bias + c0 + (j < ncol ? j : 0));
                    }
                    if (P.act == PLAGNN_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.f;
                    } else if (P.act == PLAGNN_ACT_LEAKY) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * P.slope;
                    } else if (P.act == PLAGNN_ACT_SIGMOID) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
                    }
                    if (gate_img_on) {
                        if (ch == 0) mbar_wait(gbar, (uint32_t)(tile_iter & 1));
                        float gt[32];
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                         : "=f"(gt[4 * q]), "=f"(gt[4 * q + 1]), "=f"(gt[4 * q + 2]), "=f"(gt[4 * q + 3])
                                         : "r"(gimg + (uint32_t)ch * 4096u + row_off + (uint32_t)((q ^ (lane & 7)) << 4)));
                        if (P.gate_act == PLAGNN_ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gt[j] > 0.f ? v[j] : 0.f;
                        } else if (P.gate_act == PLAGNN_ACT_LEAKY) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gt[j] > 0.f ? v[j] : v[j] * P.slope;
                        } else if (P.gate_act == PLAGNN_ACT_SIGMOID) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= gt[j] * (1.f - gt[j]);
                        }
                    } else if (P.gate) {
                        const float* gp = P.gate + r_ld * P.ldg + c0;
                        if (P.gate_act == PLAGNN_ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __ldg(gp + (j < ncol ? j : 0)) > 0.f ? v[j] : 0.f;
                        } else if (P.gate_act == PLAGNN_ACT_LEAKY) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __ldg(gp + (j < ncol ? j : 0)) > 0.f ? v[j] : v[j] * P.slope;
                        } else if (P.gate_act == PLAGNN_ACT_SIGMOID) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float y = __ldg(gp + (j < ncol ? j : 0));
                                v[j] *= y * (1.f - y);
                            }
                        }
                    }
                }
                if (TM_DEBUG(P) == 3) continue;
                if (P.tma_store) {
                    const long long e1 = (tr && t == 64) ? clock64() : 0;
                    if (lane == 0) bulk_wait_read<1>();      // the image written two chunks ago has been read
                    __syncwarp();
                    if (tr && t == 64 && tile_iter == 0) tr[14] += clock64() - e1;
                    const uint32_t img = stg + (uint32_t)buf * 4096u;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t off = row_off + (uint32_t)((q ^ (lane & 7)) << 4);
                        sts_v4(img + off, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    }
                    const long long e2 = (tr && t == 64) ? clock64() : 0;
                    fence_proxy_async_smem();
                    __syncwarp();
                    const long long e3 = (tr && t == 64) ? clock64() : 0;
                    if (TM_DEBUG(P) != 4 && lane == 0) {     // (bulk groups belong to the issuing thread: always lane 0)
                        tma_store_3d(&P.map_out, img, (int)c0, row0, direct ? 0 : split);
                        bulk_commit();
                    }
                    if (tr && t == 64 && tile_iter == 0) {
                        const long long e4 = clock64();
                        tr[15] += e4 - e1; tr[26] += e2 - e1; tr[27] += e3 - e2; tr[28] += e4 - e3;
                    }
                    buf ^= 1;
                } else if (r < P.m) {
                    // unaligned destination (row pitch % 4 != 0): plain scalar stores
                    float* drow = dst + r * ldd + c0;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < ncol) drow[j] = v[j];
                }
            }
            if (tr && t == 64 && tile_iter == 0) tr[5] = clock64();
            if (tr && t == 64 && tile_iter == 1) tr[29] = clock64() - epi1;
            // the accumulators of this tile are in registers / on their way out: the next tile may overwrite TMEM
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(tmem_empty0);
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tmem_empty) : "memory");
            }
            if (tile + 1 < end_tile) {
                // the staging images live in the lo ring, which these warps are about to refill: every warp's stores must
                // have been read out of shared memory first
                if (P.tma_store && lane == 0) bulk_wait_read<0>();
                asm volatile("bar.sync 1, %0;" ::"n"(TM_EPI_WARPS * 32) : "memory");
            }
            if (tr && t == 64 && tile_iter == 0) tr[6] = clock64();
        }
        if (P.tma_store) {
            if (lane == 0) bulk_wait_all();
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<CG>(tmem_base, TMEM_COLS);
    }
    if (tr && t == 0) { tr[7] = clock64(); tr[11] = T0.nkb; tr[12] = rank; }
}

// ordered reduction of split-K partials + epilogue
__global__ void __launch_bounds__(256) gemm_tma_reduce_kernel(const __grid_constant__ TmParams P) {
    pdl_enter();
    const int64_t total = P.m * P.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        // fp32, in split order (bit-stable).  Folding in double was tried in round 2: no measurable accuracy gain (the error is
        // in the chains, not in this sum) and 9.6 -> 15.5 us per launch, 3 % of the epoch over the 11 weight-gradient products.
        float s = 0.f;
        const int64_t r = i / P.n, c = i - r * P.n;
        const float* q = P.partial + r * P.ldp + c;
        const int64_t zs = P.m * P.ldp;
        int z = 0;
        for (; z + 8 <= P.splits; z += 8) {       // eight partials in flight, added in split order
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = __ldg(q + (int64_t)(z + j) * zs);
#pragma unroll
            for (int j = 0; j < 8; ++j) s += t[j];
        }
        for (; z < P.splits; ++z) s += __ldg(q + (int64_t)z * zs);
        if (c == P.ones_col) {                 // P.n counts the extra column; it goes to the bias gradient
            P.ones_out[r] = s;
            continue;
        }
        const float v = tm_epilogue_one(P, s, r, c);
        if (P.transpose_out) P.c[c * P.ldc + r] = v;
        else P.c[r * P.ldc + c] = v;
    }
}


// ---- third generation (round 2, last session): 256 x 128 tiles, accumulators double-buffered in tensor memory ----------------
// Per-CTA timelines of the kernel above on the forward / input-gradient products (m = 24 041, several tiles per CTA pair) show
// the tensor core idle between two tiles of a pair for 10 000 - 12 000 cycles against 20 000 - 24 000 cycles of MMAs: the last
// MMAs drain (1 750), the accumulators are read out (6 200, first tile 8 400), and only then can the next tile's first MMA be
// issued, because one 256 x 256 tile with its correction accumulator fills all 512 TMEM columns.  Here a tile is 256 x 128: main
// and correction accumulator take 256 columns, so TMEM holds TWO tiles; four dedicated warps read tile i out (bias / activation
// / gate, TMA stores) while the issuing thread is already accumulating tile i + 1 into the other half.  The price: A is fetched
// once per 128 instead of once per 256 output columns (31 instead of 21 B/clk per SM from L2), four instead of eight warps
// derive the lo tiles (24 KB per k-block of 768 MMA cycles).  Only for products written directly (no split-K), k-contiguous A,
// 16-byte aligned output, more tiles than CTA pairs; everything else runs on gemm_tma_kernel.
constexpr int DB_A_BYTES = 128 * TM_BK * 4, DB_B_BYTES = 64 * TM_BK * 4;
constexpr int DB_STAGE_BYTES = DB_A_BYTES + DB_B_BYTES;            // 24 KB
constexpr int DB_SPLIT_WARPS = 8, DB_EPI_WARPS = 4;
constexpr int DB_THREADS = (2 + DB_SPLIT_WARPS + DB_EPI_WARPS) * 32;
// ring depths (raw, lo) and the read-out warps' staging: two 4 KB store images per warp, plus one gate image when GATE
__host__ __device__ constexpr int db_epi_bytes(bool gate) { return (gate ? 3 : 2) * 4096; }
__host__ __device__ constexpr int db_smem_bytes(int raw, int lo, bool gate) {
    return (raw + lo) * DB_STAGE_BYTES + DB_EPI_WARPS * db_epi_bytes(gate) + 1024 + 256;
}

struct DbTile {
    int64_t m0, n0;
    uint32_t n_eff, n_half;
};
__device__ __forceinline__ DbTile db_tile(const TmParams& P, int t) {
    DbTile T;
    const int nt = t % P.tiles_n, mt = t / P.tiles_n;      // column blocks fastest: the N tiles of one row block run back to back
    T.m0 = (int64_t)mt * 256;
    T.n0 = (int64_t)nt * 128;
    const int64_t nrem = P.n - T.n0;
    T.n_eff = nrem >= 128 ? 128u : (uint32_t)((nrem + 63) / 64 * 64);
    T.n_half = T.n_eff / 2;
    return T;
}

// TWO (parity mode, PLAGNN_GEMM_PARITY=1): the two halves of TMEM are not two tiles but two accumulation chains of ONE tile —
// even k-blocks accumulate into the first (main + correction) pair, odd k-blocks into the second, the read-out adds the four
// accumulators in fp32 round-to-nearest.  The tensor core truncates when it adds into an accumulator, so a chain's bias grows
// with its length; halving the chains of the forward / input-gradient products halves the bias the gradient sums over 24 041
// rows amplify.  No overlap of read-out and MMAs in this mode (the tile owns all 512 columns again).
template <bool BT, int DB_RAW_STAGES, int DB_LO_STAGES, bool GATE, bool TWO = false>
__global__ void __launch_bounds__(DB_THREADS, 1) gemm_tma_db_kernel(const __grid_constant__ TmParams P) {
    using namespace tm;
    constexpr int DB_EPI_BYTES = db_epi_bytes(GATE);
    static_assert(db_smem_bytes(DB_RAW_STAGES, DB_LO_STAGES, GATE) <= 232448, "shared memory of one SM");
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t lo_ring = tiles + DB_RAW_STAGES * DB_STAGE_BYTES;
    const uint32_t epi_area = lo_ring + DB_LO_STAGES * DB_STAGE_BYTES;
    const uint32_t bars = epi_area + DB_EPI_WARPS * DB_EPI_BYTES;
    const uint32_t bar_raw_full = bars, bar_raw_empty = bars + 8 * DB_RAW_STAGES;
    const uint32_t bar_lo_full = bars + 16 * DB_RAW_STAGES, bar_lo_empty = bar_lo_full + 8 * DB_LO_STAGES;
    const uint32_t bar_acc_full = bar_lo_empty + 8 * DB_LO_STAGES;      // [2]: accumulators of buffer b complete
    const uint32_t bar_acc_empty = bar_acc_full + 16;                   // [2]: buffer b read out in both CTAs
    const uint32_t bar_gate = bar_acc_empty + 16;                       // [DB_EPI_WARPS]
    const uint32_t tmem_slot = bar_gate + 8 * DB_EPI_WARPS;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t rank = cluster_ctarank();
    const int unit = blockIdx.x / 2, units = gridDim.x / 2;
    const int first_tile = (int)((int64_t)unit * P.total_tiles / units);
    const int end_tile = (int)((int64_t)(unit + 1) * P.total_tiles / units);
#if PLAGNN_TMA_DIAG
    // diagnostics build: cycle sums per role of the first 64 CTAs (tools/gemm_trace.py prints them for this kernel as well)
    long long* const tr = (P.trace && blockIdx.x < 64) ? P.trace + 32 * blockIdx.x : nullptr;
#else
    long long* const tr = nullptr;
#endif
    if (tr && t == 0) { tr[0] = clock64(); tr[11] = P.total_kblocks; tr[12] = rank; tr[1] = end_tile - first_tile; }

    auto load_kblock = [&](const DbTile& T, int it, int g) {
        const int s = g % DB_RAW_STAGES;
        int p = 0, local = it;
        if (P.npairs > 1 && local >= P.kblocks[0]) { local -= P.kblocks[0]; p = 1; }
        const int k0 = local * TM_BK;
        const int a_row = (int)(T.m0 + rank * 128), b_row = (int)(T.n0 + rank * T.n_half);
        const uint32_t st = tiles + s * DB_STAGE_BYTES;
        const uint32_t rb = bar_raw_full + 8 * s;
        mbar_expect_tx(rb, DB_STAGE_BYTES);
        tma_load_2d(st, &P.map[p][0], k0, a_row, rb);
        if (!BT) {
            tma_load_2d(st + DB_A_BYTES, &P.map_b64[p], k0, b_row, rb);
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d(st + DB_A_BYTES + j * 4096, &P.map[p][1], b_row + 32 * j, k0, rb);
        }
    };
    const DbTile T0 = db_tile(P, first_tile);
    const int nkb = P.total_kblocks;                      // every tile runs the whole contraction
    const int preloaded = nkb < DB_RAW_STAGES ? nkb : DB_RAW_STAGES;

    if (warp == 0) {
        if (elect_one()) {
            for (int s = 0; s < DB_RAW_STAGES; ++s) {
                mbar_init(bar_raw_full + 8 * s, 1);
                mbar_init(bar_raw_empty + 8 * s, 1);
            }
            for (int s = 0; s < DB_LO_STAGES; ++s) {
                mbar_init(bar_lo_full + 8 * s, (uint32_t)(2 * DB_SPLIT_WARPS));
                mbar_init(bar_lo_empty + 8 * s, 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(bar_acc_full + 8 * b, 1);
                mbar_init(bar_acc_empty + 8 * b, (uint32_t)(2 * DB_EPI_WARPS));
            }
            for (int w = 0; w < DB_EPI_WARPS; ++w) mbar_init(bar_gate + 8 * w, 1);
            fence_mbar_init();
#pragma unroll
            for (int p = 0; p < PLAGNN_GEMM_MAX_PAIRS; ++p)
                if (p < P.npairs) { prefetch_map(&P.map[p][0]); prefetch_map(BT ? &P.map[p][1] : &P.map_b64[p]); }
            prefetch_map(&P.map_out);
            if (P.gate_tma) prefetch_map(&P.map_gate);
        }
        __syncwarp();
        pdl_wait();
        for (int it = 0; it < preloaded; ++it) {
            if (elect_one()) load_kblock(T0, it, it);
            __syncwarp();
        }
    }
    if (warp == 1) tmem_alloc<2>(tmem_slot, 512);
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        int g = 0;
        for (int tile = first_tile; tile < end_tile; ++tile) {
            const DbTile T = db_tile(P, tile);
            for (int it = (tile == first_tile ? preloaded : 0); it < nkb; ++it) {
                const int gg = g + it;
                const int s = gg % DB_RAW_STAGES;
                const uint32_t ph = (uint32_t)((gg / DB_RAW_STAGES) & 1);
                const long long w0 = tr ? clock64() : 0;
                mbar_wait(bar_raw_empty + 8 * s, ph ^ 1u);
                if (tr && lane == 0) tr[8] += clock64() - w0;
                if (elect_one()) load_kblock(T, it, gg);
                __syncwarp();
            }
            g += nkb;
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA) =================
        if (rank == 0) {
            constexpr uint64_t a_step = 2u, b_step = BT ? (1024u >> 4) : 2u;
            int g = 0, tile_iter = 0;
            for (int tile = first_tile; tile < end_tile; ++tile, ++tile_iter) {
                const DbTile T = db_tile(P, tile);
                const uint32_t idesc = make_idesc(256, T.n_eff, false, BT);
                const int b = TWO ? 0 : (tile_iter & 1), use = TWO ? tile_iter : (tile_iter >> 1);
                if (tr && lane == 0 && tile_iter == 0) tr[2] = clock64();
                if (use > 0) {                      // the previous tile in this half of TMEM has been read out in both CTAs
                    const long long w0 = tr ? clock64() : 0;
                    mbar_wait(bar_acc_empty + 8 * b, (uint32_t)((use - 1) & 1));
                    if (tr && lane == 0) tr[3] += clock64() - w0;
                    tc_fence_after();
                }
                for (int it = 0; it < nkb; ++it, ++g) {
                    // TWO: k-block `it` accumulates into chain it & 1; the first k-block of a chain starts it (acc_on = 0)
                    const uint32_t acc_main = tmem_base + (uint32_t)(TWO ? (it & 1) : b) * 256u, acc_corr = acc_main + 128u;
                    const int s = g % DB_RAW_STAGES, l = g % DB_LO_STAGES;
                    const uint32_t phl = (uint32_t)((g / DB_LO_STAGES) & 1);
                    const long long w0 = tr ? clock64() : 0;
                    mbar_wait(bar_lo_full + 8 * l, phl);
                    if (tr && lane == 0) tr[9] += clock64() - w0;
                    tc_fence_after();
                    const uint32_t st = tiles + s * DB_STAGE_BYTES, sl = lo_ring + l * DB_STAGE_BYTES;
                    const uint64_t a_hi = desc_kmajor(st), a_lo = desc_kmajor(sl);
                    const uint64_t b_hi = BT ? desc_mnmajor(st + DB_A_BYTES) : desc_kmajor(st + DB_A_BYTES);
                    const uint64_t b_lo = BT ? desc_mnmajor(sl + DB_A_BYTES) : desc_kmajor(sl + DB_A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < TM_BK / 8; ++kk) {
                            if (TM_DEBUG(P) == 2) break;
                            const uint64_t adv_a = (uint64_t)kk * a_step, adv_b = (uint64_t)kk * b_step;
                            const uint32_t acc_on = TWO ? ((it >= 2 || kk) ? 1u : 0u) : ((it | kk) ? 1u : 0u);
                            if (TM_DEBUG(P) == 7) {          // main product only (a third of the operand reads)
                                umma_tf32<2>(acc_main, a_hi + adv_a, b_hi + adv_b, idesc, acc_on);
                                continue;
                            }
                            umma_tf32<2>(acc_corr, a_lo + adv_a, b_hi + adv_b, idesc, acc_on);
                            umma_tf32<2>(acc_corr, a_hi + adv_a, b_lo + adv_b, idesc, 1u);
                            umma_tf32<2>(acc_main, a_hi + adv_a, b_hi + adv_b, idesc, acc_on);
                        }
                        umma_commit<2>(bar_lo_empty + 8 * l);
                        umma_commit<2>(bar_raw_empty + 8 * s);
                    }
                    __syncwarp();
                }
                if (elect_one()) umma_commit<2>(bar_acc_full + 8 * b);
                __syncwarp();
                if (tr && lane == 0) tr[4] = clock64();          // last tile's MMAs issued
            }
        }
        __syncwarp();
    } else if (warp < 2 + DB_SPLIT_WARPS) {
        // ================= splitting warps: lo = rn_tf32(x - trunc_tf32(x)), elementwise over the 24 KB stage =================
        const int ct = t - 64;                                  // 0 .. 32 DB_SPLIT_WARPS - 1
        const uint32_t lo_full0 = mapa(bar_lo_full, 0);
        constexpr int PIECES = DB_STAGE_BYTES / 16 / (DB_SPLIT_WARPS * 32);     // sixteen-byte pieces per thread
        constexpr uint32_t PSTRIDE = DB_SPLIT_WARPS * 32 * 16;                  // one piece of every thread
        const int total = (end_tile - first_tile) * nkb;
        for (int g = 0; g < total; ++g) {
            const int s = g % DB_RAW_STAGES, l = g % DB_LO_STAGES;
            const uint32_t phr = (uint32_t)((g / DB_RAW_STAGES) & 1), phl = (uint32_t)((g / DB_LO_STAGES) & 1);
            const long long w0 = (tr && t == 64) ? clock64() : 0;
            mbar_wait(bar_raw_full + 8 * s, phr);
            const long long w1 = (tr && t == 64) ? clock64() : 0;
            if (tr && t == 64) tr[10] += w1 - w0;
            const uint32_t src = tiles + s * DB_STAGE_BYTES + (uint32_t)ct * 16u;
            const uint32_t dstl = lo_ring + l * DB_STAGE_BYTES + (uint32_t)ct * 16u;
            float4 v[PIECES];
            if (TM_DEBUG(P) != 5) {
#pragma unroll
            for (int i = 0; i < PIECES; ++i)
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(src + (uint32_t)i * PSTRIDE));
            } else {
#pragma unroll
                for (int i = 0; i < PIECES; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const long long w2 = (tr && t == 64) ? clock64() : 0;
            mbar_wait(bar_lo_empty + 8 * l, phl ^ 1u);
            const long long w3 = (tr && t == 64) ? clock64() : 0;
            if (TM_DEBUG(P) != 5 && TM_DEBUG(P) != 6) {
#pragma unroll
            for (int i = 0; i < PIECES; ++i)
                sts_v4(dstl + (uint32_t)i * PSTRIDE, tf32_lo(v[i].x), tf32_lo(v[i].y), tf32_lo(v[i].z), tf32_lo(v[i].w));
            } else {
                float keep = 0.f;
#pragma unroll
                for (int i = 0; i < PIECES; ++i) keep += tf32_lo(v[i].x) + tf32_lo(v[i].y) + tf32_lo(v[i].z) + tf32_lo(v[i].w);
                asm volatile("" ::"f"(keep));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lo_full0 + 8 * l);
            if (tr && t == 64) { const long long w4 = clock64(); tr[13] += w3 - w2; tr[14] += (w2 - w1) + (w4 - w3); }
        }
    } else {
        // ================= last four warps: read-out of tile i while tile i + 1 accumulates in the other half of TMEM =================
        const int e = warp - (2 + DB_SPLIT_WARPS);              // 0..3
        const int lg = warp & 3;                                // TMEM lane quarter this warp may read (warp % 4)
        const uint32_t acc_empty0 = mapa(bar_acc_empty, 0);
        const uint32_t stg = epi_area + (uint32_t)e * DB_EPI_BYTES;      // two 4 KB store images, then one gate image
        const uint32_t gimg = stg + 8192u;
        const uint32_t gbar = bar_gate + 8 * (uint32_t)e;
        const uint32_t row_off = (uint32_t)lane * 128u;
        const bool gate_on = GATE && P.gate != nullptr;         // (always by TMA here: the launcher checked the alignment)
        int tile_iter = 0, buf = 0;
        uint32_t gph = 0;                                       // gate images received so far (barrier phase)
        for (int tile = first_tile; tile < end_tile; ++tile, ++tile_iter) {
            const DbTile T = db_tile(P, tile);
            const int b = TWO ? 0 : (tile_iter & 1), use = TWO ? tile_iter : (tile_iter >> 1);
            const int64_t n0 = T.n0;
            const int row0 = (int)(T.m0 + rank * 128 + lg * 32);
            // 32-column chunks of this tile that hold real columns
            int nch = (int)((T.n_eff + 31u) / 32u);
            {
                const int64_t real = (P.n - n0 + 31) / 32;
                nch = real < nch ? (int)real : nch;
            }
            if (gate_on && lane == 0) {                         // first gate image: requested before the accumulators are complete
                mbar_expect_tx(gbar, 4096u);
                tma_load_2d(gimg, &P.map_gate, (int)n0, row0, gbar);
            }
            const long long e0 = (tr && e == 0 && lane == 0) ? clock64() : 0;
            mbar_wait(bar_acc_full + 8 * b, (uint32_t)(use & 1));
            const long long e1 = (tr && e == 0 && lane == 0) ? clock64() : 0;
            if (tr && e == 0 && lane == 0) tr[15] += e1 - e0;
            tc_fence_after();
#pragma unroll 1
            for (int ch = 0; ch < nch; ++ch) {
                const int64_t c0 = n0 + ch * 32;
                const int ncol = (int)((P.n - c0) < 32 ? (P.n - c0) : 32);
                float v[32];
                {
                    uint32_t acc[32], acc_small[32];
                    const uint32_t ta = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)b * 256u + (uint32_t)(ch * 32);
                    if (!TWO) {
                        tmem_ld32(ta, acc);
                        tmem_ld32(ta + 128u, acc_small);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) + __uint_as_float(acc_small[j]);
                    } else {
                        // (main chain 0 + main chain 1) + (correction 0 + correction 1); a contraction of one k-block has no chain 1
                        const bool two = nkb >= 2;
                        tmem_ld32(ta, acc);
                        if (two) tmem_ld32(ta + 256u, acc_small);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) + (two ? __uint_as_float(acc_small[j]) : 0.f);
                        tmem_ld32(ta + 128u, acc);
                        if (two) tmem_ld32(ta + 384u, acc_small);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += __uint_as_float(acc[j]) + (two ? __uint_as_float(acc_small[j]) : 0.f);
                    }
                }
                if (P.bias) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += __ldg(P.bias + c0 + (j < ncol ? j : 0));
                }
                if (P.act == PLAGNN_ACT_RELU) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.f;
                } else if (P.act == PLAGNN_ACT_LEAKY) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * P.slope;
                } else if (P.act == PLAGNN_ACT_SIGMOID) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 1.f / (1.f + expf(-v[j]));
                }
                if (gate_on) {
                    mbar_wait(gbar, gph & 1u);
                    ++gph;
                    float gt[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(gt[4 * q]), "=f"(gt[4 * q + 1]), "=f"(gt[4 * q + 2]), "=f"(gt[4 * q + 3])
                                     : "r"(gimg + row_off + (uint32_t)((q ^ (lane & 7)) << 4)));
                    if (P.gate_act == PLAGNN_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gt[j] > 0.f ? v[j] : 0.f;
                    } else if (P.gate_act == PLAGNN_ACT_LEAKY) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = gt[j] > 0.f ? v[j] : v[j] * P.slope;
                    } else if (P.gate_act == PLAGNN_ACT_SIGMOID) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= gt[j] * (1.f - gt[j]);
                    }
                    // the image has been consumed by every lane (its values went into v above): fetch the next chunk's into it
                    __syncwarp();
                    if (ch + 1 < nch && lane == 0) {
                        mbar_expect_tx(gbar, 4096u);
                        tma_load_2d(gimg, &P.map_gate, (int)(c0 + 32), row0, gbar);
                    }
                }
                if (lane == 0) bulk_wait_read<1>();          // the image written two chunks ago has been read
                __syncwarp();
                const uint32_t img = stg + (uint32_t)buf * 4096u;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint32_t off = row_off + (uint32_t)((q ^ (lane & 7)) << 4);
                    sts_v4(img + off, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&P.map_out, img, (int)c0, row0, 0);
                    bulk_commit();
                }
                buf ^= 1;
            }
            // this half of TMEM is in registers / on its way out: the tile after next may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(acc_empty0 + 8 * b);
            if (tr && e == 0 && lane == 0) tr[16] += clock64() - e1;
        }
        if (lane == 0) bulk_wait_all();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2>(tmem_base, 512);
    }
    if (tr && t == 0) tr[7] = clock64();
}


// ---- parity mode with the read-out overlap back: half-chain ping-pong (last session of round 2) ------------------------------
// gemm_tma_db_kernel<..., TWO> gives a tile two accumulation chains but all 512 TMEM columns, so the tensor core idles during
// the read-out again.  Here the two chains of a tile are the two HALVES of its contraction in time: k-blocks [0, nkb/2) go to
// TMEM set h & 1, k-blocks [nkb/2, nkb) to the other set (h = running half-chain count), and the read-out warps fetch a half
// chain as soon as it is complete — the first half of tile i is read (into registers, main + correction summed) while its
// second half accumulates, the second half is read, added, finished (bias / activation / gate) and stored while the first
// half of tile i + 1 accumulates.  Same two chains per output as the TWO kernel (so the same halved truncation bias), summed as
// (main0 + corr0) + (main1 + corr1).  16 warps: TMA producer, MMA issuer, 6 warps for the lo pass, 8 read-out warps (two per
// TMEM lane quarter, 64 columns each: 64 partial sums per thread); one 4 KB staging buffer per read-out warp (a gate image
// and the store image take turns in it).
constexpr int DB2_SPLIT_WARPS = 6, DB2_EPI_WARPS = 8;
constexpr int DB2_THREADS = (2 + DB2_SPLIT_WARPS + DB2_EPI_WARPS) * 32;
constexpr int DB2_RAW = 5, DB2_LO = 3;
constexpr int DB2_SMEM_BYTES = (DB2_RAW + DB2_LO) * DB_STAGE_BYTES + DB2_EPI_WARPS * 4096 + 1024 + 256;
static_assert(DB2_SMEM_BYTES <= 232448, "shared memory of one SM");

namespace tm {
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
}  // namespace tm

template <bool BT, bool GATE>
__global__ void __launch_bounds__(DB2_THREADS, 1) gemm_tma_db2_kernel(const __grid_constant__ TmParams P) {
    using namespace tm;
    pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t lo_ring = tiles + DB2_RAW * DB_STAGE_BYTES;
    const uint32_t epi_area = lo_ring + DB2_LO * DB_STAGE_BYTES;
    const uint32_t bars = epi_area + DB2_EPI_WARPS * 4096;
    const uint32_t bar_raw_full = bars, bar_raw_empty = bars + 8 * DB2_RAW;
    const uint32_t bar_lo_full = bars + 16 * DB2_RAW, bar_lo_empty = bar_lo_full + 8 * DB2_LO;
    const uint32_t bar_acc_full = bar_lo_empty + 8 * DB2_LO;            // [2]: the half chain in TMEM set s is complete
    const uint32_t bar_acc_empty = bar_acc_full + 16;                   // [2]: set s has been read in both CTAs
    const uint32_t bar_gate = bar_acc_empty + 16;                       // [DB2_EPI_WARPS]
    const uint32_t tmem_slot = bar_gate + 8 * DB2_EPI_WARPS;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t rank = cluster_ctarank();
    const int unit = blockIdx.x / 2, units = gridDim.x / 2;
    const int first_tile = (int)((int64_t)unit * P.total_tiles / units);
    const int end_tile = (int)((int64_t)(unit + 1) * P.total_tiles / units);
    const int nkb = P.total_kblocks;
    const int nkb0 = (nkb + 1) >> 1;                      // k-blocks of the first half chain (the second has nkb - nkb0, maybe 0)

    auto load_kblock = [&](const DbTile& T, int it, int g) {
        const int s = g % DB2_RAW;
        int p = 0, local = it;
        if (P.npairs > 1 && local >= P.kblocks[0]) { local -= P.kblocks[0]; p = 1; }
        const int k0 = local * TM_BK;
        const int a_row = (int)(T.m0 + rank * 128), b_row = (int)(T.n0 + rank * T.n_half);
        const uint32_t st = tiles + s * DB_STAGE_BYTES;
        const uint32_t rb = bar_raw_full + 8 * s;
        mbar_expect_tx(rb, DB_STAGE_BYTES);
        tma_load_2d(st, &P.map[p][0], k0, a_row, rb);
        if (!BT) {
            tma_load_2d(st + DB_A_BYTES, &P.map_b64[p], k0, b_row, rb);
        } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_2d(st + DB_A_BYTES + j * 4096, &P.map[p][1], b_row + 32 * j, k0, rb);
        }
    };
    const DbTile T0 = db_tile(P, first_tile);
    const int preloaded = nkb < DB2_RAW ? nkb : DB2_RAW;

    if (warp == 0) {
        if (elect_one()) {
            for (int s = 0; s < DB2_RAW; ++s) {
                mbar_init(bar_raw_full + 8 * s, 1);
                mbar_init(bar_raw_empty + 8 * s, 1);
            }
            for (int s = 0; s < DB2_LO; ++s) {
                mbar_init(bar_lo_full + 8 * s, (uint32_t)(2 * DB2_SPLIT_WARPS));
                mbar_init(bar_lo_empty + 8 * s, 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(bar_acc_full + 8 * b, 1);
                mbar_init(bar_acc_empty + 8 * b, (uint32_t)(2 * DB2_EPI_WARPS));
            }
            for (int w = 0; w < DB2_EPI_WARPS; ++w) mbar_init(bar_gate + 8 * w, 1);
            fence_mbar_init();
#pragma unroll
            for (int p = 0; p < PLAGNN_GEMM_MAX_PAIRS; ++p)
                if (p < P.npairs) { prefetch_map(&P.map[p][0]); prefetch_map(BT ? &P.map[p][1] : &P.map_b64[p]); }
            prefetch_map(&P.map_out);
            if (GATE && P.gate_tma) prefetch_map(&P.map_gate);
        }
        __syncwarp();
        pdl_wait();
        for (int it = 0; it < preloaded; ++it) {
            if (elect_one()) load_kblock(T0, it, it);
            __syncwarp();
        }
    }
    if (warp == 1) tmem_alloc<2>(tmem_slot, 512);
    pdl_wait();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        int g = 0;
        for (int tile = first_tile; tile < end_tile; ++tile) {
            const DbTile T = db_tile(P, tile);
            for (int it = (tile == first_tile ? preloaded : 0); it < nkb; ++it) {
                const int gg = g + it;
                const int s = gg % DB2_RAW;
                const uint32_t ph = (uint32_t)((gg / DB2_RAW) & 1);
                mbar_wait(bar_raw_empty + 8 * s, ph ^ 1u);
                if (elect_one()) load_kblock(T, it, gg);
                __syncwarp();
            }
            g += nkb;
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA): one half chain after the other, TMEM sets alternating =================
        if (rank == 0) {
            constexpr uint64_t a_step = 2u, b_step = BT ? (1024u >> 4) : 2u;
            int g = 0, hc = 0;
            for (int tile = first_tile; tile < end_tile; ++tile) {
                const DbTile T = db_tile(P, tile);
                const uint32_t idesc = make_idesc(256, T.n_eff, false, BT);
                for (int part = 0; part < 2; ++part) {
                    const int kb_lo = part ? nkb0 : 0, kb_hi = part ? nkb : nkb0;
                    if (kb_lo >= kb_hi) continue;                    // a contraction of one k-block has no second half
                    const int set = hc & 1, use = hc >> 1;
                    if (use > 0) {                                   // the half chain that used this set before has been read out
                        mbar_wait(bar_acc_empty + 8 * set, (uint32_t)((use - 1) & 1));
                        tc_fence_after();
                    }
                    const uint32_t acc_main = tmem_base + (uint32_t)set * 256u, acc_corr = acc_main + 128u;
                    for (int it = kb_lo; it < kb_hi; ++it, ++g) {
                        const int s = g % DB2_RAW, l = g % DB2_LO;
                        const uint32_t phl = (uint32_t)((g / DB2_LO) & 1);
                        mbar_wait(bar_lo_full + 8 * l, phl);
                        tc_fence_after();
                        const uint32_t st = tiles + s * DB_STAGE_BYTES, sl = lo_ring + l * DB_STAGE_BYTES;
                        const uint64_t a_hi = desc_kmajor(st), a_lo = desc_kmajor(sl);
                        const uint64_t b_hi = BT ? desc_mnmajor(st + DB_A_BYTES) : desc_kmajor(st + DB_A_BYTES);
                        const uint64_t b_lo = BT ? desc_mnmajor(sl + DB_A_BYTES) : desc_kmajor(sl + DB_A_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < TM_BK / 8; ++kk) {
                                const uint64_t adv_a = (uint64_t)kk * a_step, adv_b = (uint64_t)kk * b_step;
                                const uint32_t acc_on = (it > kb_lo || kk) ? 1u : 0u;
                                umma_tf32<2>(acc_corr, a_lo + adv_a, b_hi + adv_b, idesc, acc_on);
                                umma_tf32<2>(acc_corr, a_hi + adv_a, b_lo + adv_b, idesc, 1u);
                                umma_tf32<2>(acc_main, a_hi + adv_a, b_hi + adv_b, idesc, acc_on);
                            }
                            umma_commit<2>(bar_lo_empty + 8 * l);
                            umma_commit<2>(bar_raw_empty + 8 * s);
                        }
                        __syncwarp();
                    }
                    if (elect_one()) umma_commit<2>(bar_acc_full + 8 * set);
                    __syncwarp();
                    ++hc;
                }
            }
        }
        __syncwarp();
    } else if (warp < 2 + DB2_SPLIT_WARPS) {
        // ================= lo pass: lo = rn_tf32(x - trunc_tf32(x)), elementwise over the 24 KB stage =================
        const int ct = t - 64;                                  // 0 .. 191
        const uint32_t lo_full0 = mapa(bar_lo_full, 0);
        constexpr int PIECES = DB_STAGE_BYTES / 16 / (DB2_SPLIT_WARPS * 32);     // 8 sixteen-byte pieces per thread
        constexpr uint32_t PSTRIDE = DB2_SPLIT_WARPS * 32 * 16;
        static_assert(PIECES * DB2_SPLIT_WARPS * 32 * 16 == DB_STAGE_BYTES, "the lo pass covers the stage exactly");
        const int total = (end_tile - first_tile) * nkb;
        for (int g = 0; g < total; ++g) {
            const int s = g % DB2_RAW, l = g % DB2_LO;
            const uint32_t phr = (uint32_t)((g / DB2_RAW) & 1), phl = (uint32_t)((g / DB2_LO) & 1);
            mbar_wait(bar_raw_full + 8 * s, phr);
            const uint32_t src = tiles + s * DB_STAGE_BYTES + (uint32_t)ct * 16u;
            const uint32_t dstl = lo_ring + l * DB_STAGE_BYTES + (uint32_t)ct * 16u;
            float4 v[PIECES];
#pragma unroll
            for (int i = 0; i < PIECES; ++i)
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(src + (uint32_t)i * PSTRIDE));
            mbar_wait(bar_lo_empty + 8 * l, phl ^ 1u);
#pragma unroll
            for (int i = 0; i < PIECES; ++i)
                sts_v4(dstl + (uint32_t)i * PSTRIDE, tf32_lo(v[i].x), tf32_lo(v[i].y), tf32_lo(v[i].z), tf32_lo(v[i].w));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lo_full0 + 8 * l);
        }
    } else {
        // ================= read-out warps: two per TMEM lane quarter, 64 columns (two chunks of 32) each =================
        const int e = warp - (2 + DB2_SPLIT_WARPS);             // 0..7
        const int lg = warp & 3;                                // TMEM lane quarter this warp may read (warp % 4)
        const int chalf = e >> 2;                               // columns [64 chalf, 64 chalf + 64) of the tile
        const uint32_t acc_empty0 = mapa(bar_acc_empty, 0);
        const uint32_t stg = epi_area + (uint32_t)e * 4096u;    // one 4 KB buffer: gate image, then store image
        const uint32_t gbar = bar_gate + 8 * (uint32_t)e;
        const uint32_t row_off = (uint32_t)lane * 128u;
        const bool gate_on = GATE && P.gate != nullptr;
        const bool two = nkb > nkb0;
        int hc = 0;
        uint32_t gph = 0;
        for (int tile = first_tile; tile < end_tile; ++tile) {
            const DbTile T = db_tile(P, tile);
            const int64_t n0 = T.n0;
            const int row0 = (int)(T.m0 + rank * 128 + lg * 32);
            int nch_tile = (int)((T.n_eff + 31u) / 32u);        // 32-column chunks of the tile that hold real columns
            {
                const int64_t real = (P.n - n0 + 31) / 32;
                nch_tile = real < nch_tile ? (int)real : nch_tile;
            }
            int mych = nch_tile - 2 * chalf;                    // of which this warp reads 0, 1 or 2
            mych = mych < 0 ? 0 : mych > 2 ? 2 : mych;
            float v[64];
            // ---- first half chain: partial sums into registers, then the set is free for the next tile's first half ----
            {
                const int set = hc & 1, use = hc >> 1;
                mbar_wait(bar_acc_full + 8 * set, (uint32_t)(use & 1));
                tc_fence_after();
                const uint32_t ta = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)set * 256u + (uint32_t)(chalf * 64);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q * 16 < mych * 32) {                   // warp-uniform
                        uint32_t a[16], c[16];
                        tmem_ld16(ta + (uint32_t)(q * 16), a);
                        tmem_ld16(ta + 128u + (uint32_t)(q * 16), c);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[q * 16 + j] = __uint_as_float(a[j]) + __uint_as_float(c[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty0 + 8 * set);
                ++hc;
            }
            // ---- second half chain: added to the partial sums ----
            if (two) {
                const int set = hc & 1, use = hc >> 1;
                mbar_wait(bar_acc_full + 8 * set, (uint32_t)(use & 1));
                tc_fence_after();
                const uint32_t ta = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)set * 256u + (uint32_t)(chalf * 64);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q * 16 < mych * 32) {
                        uint32_t a[16], c[16];
                        tmem_ld16(ta + (uint32_t)(q * 16), a);
                        tmem_ld16(ta + 128u + (uint32_t)(q * 16), c);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[q * 16 + j] += __uint_as_float(a[j]) + __uint_as_float(c[j]);
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc_empty0 + 8 * set);
                ++hc;
            }
            // ---- finish and store the warp's chunks (everything is in registers: TMEM is already released) ----
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                if (cc < mych) {                                // warp-uniform
                    const int64_t c0 = n0 + (chalf * 2 + cc) * 32;
                    const int ncol = (int)((P.n - c0) < 32 ? (P.n - c0) : 32);
                    float* w = v + cc * 32;
                    if (P.bias) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) w[j] += __ldg(P.bias + c0 + (j < ncol ? j : 0));
                    }
                    if (P.act == PLAGNN_ACT_RELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) w[j] = w[j] > 0.f ? w[j] : 0.f;
                    } else if (P.act == PLAGNN_ACT_LEAKY) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) w[j] = w[j] > 0.f ? w[j] : w[j] * P.slope;
                    } else if (P.act == PLAGNN_ACT_SIGMOID) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) w[j] = 1.f / (1.f + expf(-w[j]));
                    }
                    // the staging buffer is free when the previous store has read it
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    if (gate_on) {
                        if (lane == 0) {
                            mbar_expect_tx(gbar, 4096u);
                            tma_load_2d(stg, &P.map_gate, (int)c0, row0, gbar);
                        }
                        mbar_wait(gbar, gph & 1u);
                        ++gph;
                        float gt[32];
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                         : "=f"(gt[4 * q]), "=f"(gt[4 * q + 1]), "=f"(gt[4 * q + 2]), "=f"(gt[4 * q + 3])
                                         : "r"(stg + row_off + (uint32_t)((q ^ (lane & 7)) << 4)));
                        if (P.gate_act == PLAGNN_ACT_RELU) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) w[j] = gt[j] > 0.f ? w[j] : 0.f;
                        } else if (P.gate_act == PLAGNN_ACT_LEAKY) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) w[j] = gt[j] > 0.f ? w[j] : w[j] * P.slope;
                        } else if (P.gate_act == PLAGNN_ACT_SIGMOID) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) w[j] *= gt[j] * (1.f - gt[j]);
                        }
                        __syncwarp();                           // every lane has consumed the gate image: the buffer takes the output
                    }
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t off = row_off + (uint32_t)((q ^ (lane & 7)) << 4);
                        sts_v4(stg + off, w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_3d(&P.map_out, stg, (int)c0, row0, 0);
                        bulk_commit();
                    }
                }
            }
        }
        if (lane == 0) bulk_wait_all();
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2>(tmem_base, 512);
    }
}

// ---- host side: tensor maps ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

struct MapKey {
    const void* base;
    int64_t inner, outer, pitch;
    int mn_major;
    bool operator==(const MapKey& o) const {
        return base == o.base && inner == o.inner && outer == o.outer && pitch == o.pitch && mn_major == o.mn_major;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        size_t h = reinterpret_cast<size_t>(k.base);
        h = h * 1315423911u ^ (size_t)k.inner;
        h = h * 1315423911u ^ (size_t)k.outer;
        h = h * 1315423911u ^ (size_t)k.pitch;
        return h * 2 + (size_t)k.mn_major;
    }
};

// operand stored [outer][inner] fp32 with row pitch `pitch` elements.  k-contiguous operands: inner = k, box {32, 128}.
// mn-contiguous operands: inner = mn, box {32, 32}.  Encodings are cached per thread (the arena pointers of the
// whole-network engine repeat every epoch).
static int get_map(CUtensorMap* out, const float* base, int64_t inner, int64_t outer, int64_t pitch, int mn_major) {
    static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{base, inner, outer, pitch, mn_major};
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return PLAGNN_OK;
    }
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(PLAGNN_ERR_CUDA, "gemm_tma", "cuTensorMapEncodeTiled not available");
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 4};
    // mn_major == 2: a 32 x 32 box of a row-major matrix (gate tiles); == 3: k-contiguous operand with a 64-row box
    cuuint32_t box[2] = {32u, mn_major == 3 ? 64u : mn_major ? 32u : 128u};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    const CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           mn_major == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tma: cuTensorMapEncodeTiled failed (%d) base=%p inner=%lld outer=%lld pitch=%lld", (int)r, (const void*)base,
                  (long long)inner, (long long)outer, (long long)pitch);
        return PLAGNN_ERR_CUDA;
    }
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, m);
    *out = m;
    return PLAGNN_OK;
}

// output (or split-K partials) as a 3-D tensor {n, m, depth} with row pitch `pitch` and slice pitch m * pitch; box {32, 32, 1}
static int get_out_map(CUtensorMap* out, const float* base, int64_t n, int64_t m, int64_t depth, int64_t pitch) {
    static thread_local std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
    const MapKey key{base, n, m, pitch, (int)depth};
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return PLAGNN_OK;
    }
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(PLAGNN_ERR_CUDA, "gemm_tma", "cuTensorMapEncodeTiled not available");
    cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)m, (cuuint64_t)depth};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)m * (cuuint64_t)pitch * 4};
    cuuint32_t box[3] = {32u, 32u, 1u};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMap t;
    const CUresult r = enc(&t, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm_tma: cuTensorMapEncodeTiled (output) failed (%d) base=%p n=%lld m=%lld depth=%lld pitch=%lld", (int)r,
                  (const void*)base, (long long)n, (long long)m, (long long)depth, (long long)pitch);
        return PLAGNN_ERR_CUDA;
    }
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, t);
    *out = t;
    return PLAGNN_OK;
}

// PLAGNN_TMA_TRACE=1: clock64 stamps of the first 64 CTAs of the LAST launch (plagnn_tma_trace reads them back)
static long long* g_trace = nullptr;
static long long* trace_buffer(cudaStream_t st) {
    static const bool on = getenv("PLAGNN_TMA_TRACE") != nullptr;
    if (!on) return nullptr;
    if (!g_trace && cudaMalloc(&g_trace, 64 * 32 * sizeof(long long)) != cudaSuccess) return nullptr;
    cudaMemsetAsync(g_trace, 0, 64 * 32 * sizeof(long long), st);
    return g_trace;
}
extern "C" int plagnn_tma_trace(long long* host_out /* 64 x 32 */) {
    if (!g_trace) return -1;
    cudaDeviceSynchronize();
    return cudaMemcpy(host_out, g_trace, 64 * 32 * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}

// PLAGNN_TMA_CG=1 selects the single-CTA 128 x 128 tile (bring-up / cross-check); read per call so tests can flip it
static int tm_cg() {
    const char* e = getenv("PLAGNN_TMA_CG");
    return (e && e[0] == '1') ? 1 : 2;
}

// Parity mode (DEFAULT; PLAGNN_GEMM_PARITY, read per call: "0" off = the fastest kernels, "1" only the products with a bias /
// activation epilogue, i.e. the forward pass, "2" / unset all).  Tall directly written products run the 256 x 128 kernel with its
// two TMEM halves as two accumulation chains of one tile (half the truncation bias of the tensor core's accumulate), the long-K
// weight-gradient products are cut into chains of 24 k-blocks.  Measured at the full PPI size (N = 24 041), worst of the 19
// gradient tensors against the fp32 oracle on shared decisions / epoch time: off 1.49e-5 / 2.15 ms; with the two chains side by
// side in TMEM (gemm_tma_db_kernel<..., TWO>, no read-out overlap): level 1 9.6e-6 / 2.23 ms, level 2 8.6e-6 / 2.29 ms (chains
// of 40 in the weight gradients: 1.10e-5, of 16: 8.1e-6 / 2.37 ms); with the two chains as the two halves of the contraction in
// time (gemm_tma_db2_kernel, the default): level 2 8.8e-6 / 2.19 ms.  DESIGN.md 3.
static int tm_parity_level() {
    const char* e = getenv("PLAGNN_GEMM_PARITY");
    return e ? (e[0] == '0' ? 0 : e[0] == '1' ? 1 : 2) : 2;
}
static bool tm_parity() { return tm_parity_level() > 0; }

// PLAGNN_TMA_LONG_CHAIN overrides TM_LONG_CHAIN (accuracy / speed experiments); read once
static int tm_long_chain() {
    static const int v = [] { const char* e = getenv("PLAGNN_TMA_LONG_CHAIN"); const int x = e ? atoi(e) : 0; return x >= 4 && x <= 64 ? x : 0; }();
    return v ? v : tm_parity() ? 24 : TM_LONG_CHAIN;
}

static int tm_choose_splits(int64_t m, int64_t n, int total_kblocks, int cg) {
    const int64_t tile = 128 * cg;
    const int64_t tiles = ceil_div(m, tile) * ceil_div(n, tile);
    const int64_t units = sm_count() / cg;          // CTAs (CG = 1) or CTA pairs (CG = 2) resident at once
    int64_t s = 1;
    if (tiles < units && total_kblocks >= 8) {
        s = units / tiles;
        if (s > total_kblocks / 4) s = total_kblocks / 4;
        if (s > 64) s = 64;
        if (s < 1) s = 1;
    }
    const int64_t for_accuracy = ceil_div(total_kblocks, total_kblocks >= TM_LONG_K_BLOCKS ? tm_long_chain() : TM_MAX_CHAIN);
    int64_t s0 = s > for_accuracy ? s : for_accuracy;
    if (s0 > 1) {
        int64_t best = s0, best_cost = INT64_MAX;
        const int64_t lo = (s0 > 1 && ceil_div(total_kblocks, s0 - 1) <= 48) ? s0 - 1 : s0;
        for (int64_t c = lo; c <= s0 + 8 && c <= total_kblocks; ++c) {
            const int64_t cost = ceil_div(tiles * c, units) * (ceil_div(total_kblocks, c) + 4);   // +4: prologue/epilogue
            if (cost < best_cost) { best_cost = cost; best = c; }
        }
        s0 = best;
    }
    return (int)s0;
}

size_t gemm_tma_partial_bytes(int64_t m, int64_t n, int64_t k_total) {
    // either orientation (the launcher may compute the transpose), either tile configuration, pairs rounded up separately
    const int kb = (int)ceil_div(k_total, TM_BK);
    size_t best = 0;
    for (int o = 0; o < 2; ++o) {
        const int64_t mm = o ? n : m, nn = o ? m : n;
        int s = 1;
        for (int cg = 1; cg <= 2; ++cg)
            for (int extra = 0; extra <= PLAGNN_GEMM_MAX_PAIRS; ++extra) {
                const int c = tm_choose_splits(mm, nn, kb + extra, cg);
                s = c > s ? c : s;
            }
        const size_t b = s > 1 ? align_up((size_t)s * (size_t)mm * (size_t)((nn + 3) / 4 * 4) * sizeof(float), 256) : 0;
        best = b > best ? b : best;
    }
    return best;
}

bool gemm_tma_eligible(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs) {
    if (n < 16 || m < 1) return false;
    for (int p = 0; p < npairs; ++p) {
        const plagnn_gemm_pair& q = pairs[p];
        if (q.k < 8) return false;
        if ((q.lda & 3) || (q.ldb & 3) || !aligned16(q.a) || !aligned16(q.b)) return false;
        if ((q.a_trans != 0) != (pairs[0].a_trans != 0) || (q.b_trans != 0) != (pairs[0].b_trans != 0)) return false;
        if (q.k > INT_MAX - 64 || m > INT_MAX - 512 || n > INT_MAX - 512) return false;
    }
    return true;
}

// padded work of an m x n output in the CTA-pair tiling: 256-row tiles times the MMA widths of the column tiles
static int64_t tm_padded_area(int64_t m, int64_t n, int cg) {
    const int64_t tile = 128 * cg, gran = 32 * cg;
    const int64_t full = n / tile, rem = n - full * tile;
    return ceil_div(m, tile) * tile * (full * tile + ceil_div(rem, gran) * gran);
}

// Which kernel for a directly written product?  Both main loops run at the rate at which shared memory can feed them (TMA
// writes + the lo pass + the tensor core's operand reads: ~125 B/clk per SM in either kernel), so the 256 x 128 kernel pays
// 1 150 cycles per k-block of half a tile where the 256 x 256 kernel pays 1 600 for a whole one, and wins through what it does
// not pay: the 10 000 - 13 000 idle cycles between two tiles of a pair, and the padding of a ragged last column tile (n = 300 is
// 256 + 64 there, 128 + 128 + 64 here).  Constants fitted to the isolated timings of the epoch's products (tools/gemm_db_ab.py,
// profiles/r2_gemm_db_ab.json); the model picks the measured winner for every one of them except two ties.
static double tm_cost_tiles256(int64_t m, int64_t n, int kblocks, int64_t units) {
    const int64_t tiles_m = ceil_div(m, 256), full = n / 256, rem = n - full * 256;
    const int64_t tiles_n = full + (rem ? 1 : 0);
    const double wsum = (double)full + (rem ? 0.55 + 0.45 * (double)(ceil_div(rem, 64) * 64) / 256.0 : 0.0);
    const double rounds = (double)ceil_div(tiles_m * tiles_n, units);
    return rounds * ((double)kblocks * 1600.0 * (wsum / (double)tiles_n) + 13500.0) + 6000.0;
}
static double tm_cost_tiles128(int64_t m, int64_t n, int kblocks, int64_t units) {
    const int64_t tiles_m = ceil_div(m, 256), full = n / 128, rem = n - full * 128;
    const int64_t tiles_n = full + (rem ? 1 : 0);
    const double wsum = (double)full + (rem ? 0.7 + 0.3 * (double)(ceil_div(rem, 64) * 64) / 128.0 : 0.0);
    const double rounds = (double)ceil_div(tiles_m * tiles_n, units);
    return rounds * (double)kblocks * 1150.0 * (wsum / (double)tiles_n) + 10000.0;
}

// ones_out != nullptr: weight-gradient product with the bias gradient riding along — B (mn-contiguous, one pair) gets a
// virtual column n that reads as 1.0, C stays m x n and ones_out[m] receives sum_k A[:, k].  Always split-K.
int gemm_tma_launch(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs_in, const float* bias, int act,
                    float slope, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc, void* workspace,
                    size_t workspace_bytes, cudaStream_t st, float* ones_out) {
    TmParams P;
    memset(&P, 0, sizeof(P));
    P.ones_col = -1;
    const int64_t n_map = n;                 // extent of B's tensor map: the virtual column is out of bounds (zero fill)
    if (ones_out) {
        P.ones_col = (int)n;
        P.ones_out = ones_out;
        n += 1;
    }
    // Split-K products without an epilogue (the weight gradients) are written by the reduction kernel, which can just as
    // well write the transpose: compute C^T = B^T A^T when that orientation wastes less of the 256 x 256 tiles
    // (dW[300 x 400]: 512 x 448 padded vs 512 x 320 swapped).
    plagnn_gemm_pair swapped[PLAGNN_GEMM_MAX_PAIRS];
    const plagnn_gemm_pair* pairs = pairs_in;
    {
        const int cg0 = tm_cg();
        int tk = 0;
        for (int p = 0; p < npairs; ++p) tk += (int)ceil_div(pairs_in[p].k, TM_BK);
        static const bool no_swap = getenv("PLAGNN_TMA_NO_SWAP") != nullptr;
        if (!no_swap && !ones_out && !bias && !gate && act == PLAGNN_ACT_NONE && m >= 16 && tm_choose_splits(m, n, tk, cg0) > 1 &&
            tm_choose_splits(n, m, tk, cg0) > 1 && tm_padded_area(n, m, cg0) < tm_padded_area(m, n, cg0)) {
            for (int p = 0; p < npairs; ++p) {
                const plagnn_gemm_pair& q = pairs_in[p];
                swapped[p] = plagnn_gemm_pair{q.b, q.ldb, q.b_trans, q.a, q.lda, q.a_trans, q.k};
            }
            pairs = swapped;
            const int64_t t = m; m = n; n = t;
            P.transpose_out = 1;
        }
    }
    P.m = m; P.n = n; P.npairs = npairs;
    P.total_kblocks = 0;
    for (int p = 0; p < npairs; ++p) {
        const plagnn_gemm_pair& q = pairs[p];
        P.kblocks[p] = (int)ceil_div(q.k, TM_BK);
        P.total_kblocks += P.kblocks[p];
        int rc;
        if ((rc = q.a_trans ? get_map(&P.map[p][0], q.a, m, q.k, q.lda, 1) : get_map(&P.map[p][0], q.a, q.k, m, q.lda, 0))) return rc;
        const int64_t nb = P.transpose_out ? n : n_map;   // (no virtual column when the operands were swapped)
        if ((rc = q.b_trans ? get_map(&P.map[p][1], q.b, nb, q.k, q.ldb, 1) : get_map(&P.map[p][1], q.b, q.k, nb, q.ldb, 0))) return rc;
    }
    P.bias = bias; P.act = act; P.slope = slope; P.gate = gate; P.ldg = ldg; P.gate_act = gate_act;
    P.c = c; P.ldc = ldc;
    { const char* e = getenv("PLAGNN_TMA_DEBUG"); P.debug = e ? atoi(e) : 0; }
    P.trace = trace_buffer(st);
    { const char* e = getenv("PLAGNN_TMA_SINGLE_ACC"); P.single_acc = (e && e[0] == '1') ? 1 : 0; }
    const int cg = tm_cg();
    const int splits = tm_choose_splits(m, n, P.total_kblocks, cg);
    if (ones_out && splits < 2) return fail(PLAGNN_ERR_UNSUPPORTED, "gemm_tma", "the bias-gradient column needs a split-K product");
    P.ldp = (n + 3) / 4 * 4;
    if (splits > 1 && (!workspace || !aligned16(workspace) || workspace_bytes < (size_t)splits * m * P.ldp * sizeof(float)))
        return fail(PLAGNN_ERR_WORKSPACE, "gemm_tma", "split-K workspace too small (see plagnn_gemm_workspace_bytes)");
    P.kblocks_per_split = (int)ceil_div(P.total_kblocks, splits);
    P.splits = (int)ceil_div(P.total_kblocks, P.kblocks_per_split);
    P.partial = P.splits > 1 ? (float*)workspace : nullptr;
    // the epilogue leaves through TMA stores when the destination rows are 16-byte aligned (always true for the partials)
    {
        static const bool no_store = getenv("PLAGNN_TMA_NO_STORE") != nullptr;
        int rc;
        if (no_store) {
            P.tma_store = 0;
        } else if (P.splits > 1) {
            P.tma_store = 1;
            if ((rc = get_out_map(&P.map_out, P.partial, n, m, P.splits, P.ldp))) return rc;
        } else if ((ldc & 3) == 0 && aligned16(c)) {
            P.tma_store = 1;
            if ((rc = get_out_map(&P.map_out, c, n, m, 1, ldc))) return rc;
        } else {
            P.tma_store = 0;
        }
    }

    P.gate_tma = 0;
    if (gate && P.splits == 1 && (ldg & 3) == 0 && aligned16(gate)) {
        int rc;
        if ((rc = get_map(&P.map_gate, gate, n, m, ldg, 2))) return rc;
        P.gate_tma = 1;
    }

    // ---- products written directly, with more 256 x 128 tiles than CTA pairs: the double-buffered kernel ----
    {
        // PLAGNN_TMA_DB=0: never; PLAGNN_TMA_DB_NOW (read per launch, for A/B timing in one process): "0" never, "1" whenever
        // the product qualifies, unset / anything else: by the cost model
        static const bool db_allowed = [] { const char* e = getenv("PLAGNN_TMA_DB"); return !e || e[0] != '0'; }();
        const char* dyn = getenv("PLAGNN_TMA_DB_NOW");
        const int64_t db_tiles = ceil_div(m, 256) * ceil_div(n, 128);
        const int64_t pair_units = sm_count() / 2;
        const bool parity = tm_parity_level() == 2 || (tm_parity_level() == 1 && (bias != nullptr || act != PLAGNN_ACT_NONE));
        bool db_on = parity || (db_allowed && !(dyn && dyn[0] == '0'));
        if (db_on && !parity && !(dyn && dyn[0] == '1'))
            db_on = tm_cost_tiles128(m, n, P.total_kblocks, pair_units) < tm_cost_tiles256(m, n, P.total_kblocks, pair_units);
        bool k_major_a = true;
        for (int p = 0; p < npairs; ++p) k_major_a = k_major_a && !pairs[p].a_trans;
        if (db_on && tm_cg() == 2 && P.splits == 1 && !ones_out && P.tma_store && k_major_a && (!gate || P.gate_tma) &&
            db_tiles > sm_count() / 2 && db_tiles < ((int64_t)1 << 30)) {
            if (!pairs[0].b_trans) {
                for (int p = 0; p < npairs; ++p) {
                    int rc;
                    if ((rc = get_map(&P.map_b64[p], pairs[p].b, pairs[p].k, n_map, pairs[p].ldb, 3))) return rc;
                }
            }
            P.tiles_m = (int)ceil_div(m, 256);
            P.tiles_n = (int)ceil_div(n, 128);
            P.total_tiles = (int)db_tiles;
            P.persistent = 1;
            using DbFn = void (*)(const TmParams);
            // [ring variant][gate][b_trans]; ring variants (PLAGNN_TMA_DB_RING): 0 = raw 4 / lo 3, 1 = raw 5 / lo 3, 2 = raw 4 / lo 4
            // (the deeper rings only without a gate image)
            static const DbFn db_kernels[3][2][2] = {
                {{gemm_tma_db_kernel<false, 4, 3, false>, gemm_tma_db_kernel<true, 4, 3, false>},
                 {gemm_tma_db_kernel<false, 4, 3, true>, gemm_tma_db_kernel<true, 4, 3, true>}},
                {{gemm_tma_db_kernel<false, 5, 3, false>, gemm_tma_db_kernel<true, 5, 3, false>},
                 {gemm_tma_db_kernel<false, 4, 3, true>, gemm_tma_db_kernel<true, 4, 3, true>}},
                {{gemm_tma_db_kernel<false, 4, 4, false>, gemm_tma_db_kernel<true, 4, 4, false>},
                 {gemm_tma_db_kernel<false, 4, 3, true>, gemm_tma_db_kernel<true, 4, 3, true>}}};
            static const int db_smem[3][2] = {{db_smem_bytes(4, 3, false), db_smem_bytes(4, 3, true)},
                                              {db_smem_bytes(5, 3, false), db_smem_bytes(4, 3, true)},
                                              {db_smem_bytes(4, 4, false), db_smem_bytes(4, 3, true)}};
            // parity mode: the same kernel with its TMEM halves as two accumulation chains of one tile; [gate][b_trans]
            static const DbFn db_two[2][2] = {
                {gemm_tma_db_kernel<false, 5, 3, false, true>, gemm_tma_db_kernel<true, 5, 3, false, true>},
                {gemm_tma_db_kernel<false, 4, 3, true, true>, gemm_tma_db_kernel<true, 4, 3, true, true>}};
            // ... and with the read-out overlap back: the two chains as the two halves of the contraction in time
            // (gemm_tma_db2_kernel; PLAGNN_TMA_DB2=0, read per launch, falls back to the kernel above)
            static const DbFn db2[2][2] = {{gemm_tma_db2_kernel<false, false>, gemm_tma_db2_kernel<true, false>},
                                           {gemm_tma_db2_kernel<false, true>, gemm_tma_db2_kernel<true, true>}};
            bool use_db2 = parity;
            { const char* e = getenv("PLAGNN_TMA_DB2"); if (e && e[0] == '0') use_db2 = false; }
            int ring = 1;
            { const char* e = getenv("PLAGNN_TMA_DB_RING"); if (e && e[0] >= '0' && e[0] <= '2') ring = e[0] - '0'; }
            if (parity) ring = 1;
            const int gi = gate ? 1 : 0, bi = pairs[0].b_trans ? 1 : 0;
            static thread_local int db_attr_dev = -1;
            int dev = 0;
            cudaGetDevice(&dev);
            if (db_attr_dev != dev) {
                for (int r = 0; r < 3; ++r)
                    for (int g2 = 0; g2 < 2; ++g2)
                        for (int b2 = 0; b2 < 2; ++b2) {
                            cudaError_t e = cudaFuncSetAttribute(db_kernels[r][g2][b2], cudaFuncAttributeMaxDynamicSharedMemorySize, db_smem[r][g2]);
                            if (e != cudaSuccess) {
                                set_error("gemm_tma: cudaFuncSetAttribute (db): %s", cudaGetErrorString(e));
                                return PLAGNN_ERR_CUDA;
                            }
                        }
                for (int g2 = 0; g2 < 2; ++g2)
                    for (int b2 = 0; b2 < 2; ++b2) {
                        cudaError_t e2 = cudaFuncSetAttribute(db2[g2][b2], cudaFuncAttributeMaxDynamicSharedMemorySize, DB2_SMEM_BYTES);
                        if (e2 != cudaSuccess) {
                            set_error("gemm_tma: cudaFuncSetAttribute (db2): %s", cudaGetErrorString(e2));
                            return PLAGNN_ERR_CUDA;
                        }
                        cudaError_t e = cudaFuncSetAttribute(db_two[g2][b2], cudaFuncAttributeMaxDynamicSharedMemorySize, db_smem[1][g2]);
                        if (e != cudaSuccess) {
                            set_error("gemm_tma: cudaFuncSetAttribute (db, two chains): %s", cudaGetErrorString(e));
                            return PLAGNN_ERR_CUDA;
                        }
                    }
                db_attr_dev = dev;
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(sm_count() / 2 * 2), 1u, 1u);
            cfg.blockDim = dim3(use_db2 ? DB2_THREADS : DB_THREADS);
            cfg.dynamicSmemBytes = use_db2 ? DB2_SMEM_BYTES : db_smem[ring][gi];
            cfg.stream = st;
            cudaLaunchAttribute at[2];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
            cfg.attrs = at; cfg.numAttrs = 2;
            cudaError_t e = cudaLaunchKernelEx(&cfg, use_db2 ? db2[gi][bi] : parity ? db_two[gi][bi] : db_kernels[ring][gi][bi], P);
            if (e != cudaSuccess) {
                set_error("gemm_tma: launch (db): %s", cudaGetErrorString(e));
                return PLAGNN_ERR_CUDA;
            }
            return check_launch("gemm_tma", 1);
        }
    }

    using KernelFn = void (*)(const TmParams);
    static const KernelFn kernels[8] = {
        gemm_tma_kernel<1, false, false>, gemm_tma_kernel<1, false, true>, gemm_tma_kernel<1, true, false>, gemm_tma_kernel<1, true, true>,
        gemm_tma_kernel<2, false, false>, gemm_tma_kernel<2, false, true>, gemm_tma_kernel<2, true, false>, gemm_tma_kernel<2, true, true>};
    static thread_local int attr_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        for (int i = 0; i < 8; ++i) {
            cudaError_t e = cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, TM_SMEM_BYTES);
            if (e != cudaSuccess) {
                set_error("gemm_tma: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return PLAGNN_ERR_CUDA;
            }
        }
        attr_dev = dev;
    }
    const int kidx = (cg == 2 ? 4 : 0) + (pairs[0].a_trans ? 2 : 0) + (pairs[0].b_trans ? 1 : 0);
    const int64_t tile = 128 * cg;
    P.tiles_m = (int)ceil_div(m, tile);
    P.tiles_n = (int)ceil_div(n, tile);
    const int64_t total_tiles = (int64_t)P.tiles_m * P.tiles_n * P.splits;
    if (total_tiles >= ((int64_t)1 << 30)) return fail(PLAGNN_ERR_UNSUPPORTED, "gemm_tma", "too many tiles");
    P.total_tiles = (int)total_tiles;
    // more tiles than CTA pairs fit at once: each pair walks several tiles (set-up once, next tile's loads overlap the
    // read-out).  Not with a gate (its images need the shared memory the next tile's loads would use).
    static const bool allow_persistent = [] { const char* e = getenv("PLAGNN_TMA_PERSISTENT"); return !e || e[0] != '0'; }();
    const int64_t units = sm_count() / cg;
    P.persistent = (allow_persistent && !gate && total_tiles > units) ? 1 : 0;
    const int64_t launched = P.persistent ? units : total_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(launched * cg), 1u, 1u);
    cfg.blockDim = dim3(TM_THREADS);
    cfg.dynamicSmemBytes = TM_SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cg; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernels[kidx], P);
    if (e != cudaSuccess) {
        set_error("gemm_tma: launch: %s", cudaGetErrorString(e));
        return PLAGNN_ERR_CUDA;
    }
    if (P.splits > 1) {
        const int64_t total = m * n;
        const int g = (int)(ceil_div(total, 256) < (int64_t)sm_count() * 8 ? ceil_div(total, 256) : (int64_t)sm_count() * 8);
        launch_pdl(gemm_tma_reduce_kernel, dim3(g), dim3(256), 0, st, P);
    }
    return check_launch("gemm_tma", P.splits > 1 ? 2 : 1);
}

}  // namespace plagnn
