// K3 — dense feature x weight contraction on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Replaces the cuBLAS SGEMMs behind nn.Linear in code/model.py:16-17,20-28 (fc_pool / fc_self / fc_neigh
// inside SAGEConv, liner1, liner2) and their autograd backward.
//
// fp32-level accuracy on TF32 tensor cores: every fp32 operand x is split in registers into
//   hi = tf32(x)   and   lo = tf32(x - hi)
// and the product is accumulated in fp32 in tensor memory as  hi*hi + hi*lo + lo*hi  (3xTF32).
// Because the split has to pass through registers anyway, operands are NOT brought in by TMA: eight
// loader warps read the fp32 tiles with coalesced 16-byte loads, split them and write the four tf32
// tiles (A_hi, A_lo, B_hi, B_lo) straight into the 128-byte-swizzled K-major layout that the UMMA
// shared-memory descriptors expect, transposing on the way when the operand is stored MN-contiguous.
//
// CTA = 9 warps:  warps 0-7 loaders (then epilogue: TMEM -> registers -> global),  warp 8 = TMEM
// allocator + single-thread tcgen05.mma issuer.  3-stage shared-memory ring (64 KB per stage) handed
// over with mbarriers: loaders -> full[s] -> MMA -> tcgen05.commit -> empty[s]; accumulator hand-off
// to the epilogue through a third mbarrier.  One 128 x 128 output tile per CTA, optional split-K.
#include "common.cuh"
#include <cstdlib>
#include <cstdint>
#include <climits>
#ifndef PLAGNN_TC_DEBUG
#define PLAGNN_TC_DEBUG 0
#endif

namespace plagnn {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 32, TC_STAGES = 3;
constexpr int TC_LOAD_WARPS = 16;  // warps 0-7 stage the A tile, warps 8-15 the B tile (4 float4 per thread and k-block)
constexpr int TC_THREADS = (TC_LOAD_WARPS + 1) * 32;
constexpr int TC_PART_BYTES = TC_BM * TC_BK * 4;          // 16 KB: one tf32 tile (128 rows x 128 bytes)
constexpr int TC_STAGE_BYTES = 4 * TC_PART_BYTES;         // A_hi, A_lo, B_hi, B_lo
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TC_TMEM_COLS = 256;   // two fp32 accumulators: [0,128) hi*hi, [128,256) lo*hi + hi*lo

struct TcParams {
    int64_t m, n;
    int npairs;
    const float* a[PLAGNN_GEMM_MAX_PAIRS];
    const float* b[PLAGNN_GEMM_MAX_PAIRS];
    int64_t lda[PLAGNN_GEMM_MAX_PAIRS], ldb[PLAGNN_GEMM_MAX_PAIRS];
    int a_trans[PLAGNN_GEMM_MAX_PAIRS], b_trans[PLAGNN_GEMM_MAX_PAIRS];
    int a_vec[PLAGNN_GEMM_MAX_PAIRS], b_vec[PLAGNN_GEMM_MAX_PAIRS];
    int64_t k[PLAGNN_GEMM_MAX_PAIRS];
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
    int tiles_m, tiles_n, total_tiles;   // persistent kernel: tile = split * tiles_m * tiles_n + mt * tiles_n + nt
    int debug;   // PLAGNN_TC_DEBUG: 1 = loaders skip global loads / smem stores after the first k-block (MMA-bound
                 // timing), 2 = the issuer skips the MMAs (loader-bound timing).  Results are garbage; timing only.
};

// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
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

// K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), descriptor version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// MN-major operands (stored mn-contiguous, e.g. both operands of a weight-gradient GEMM).  For 32-bit elements the
// only MN-major layout the tensor core reads is "128-byte swizzle with 32-byte atomicity" (layout type 1): swizzle
// atom = 4 k-rows x 128 bytes (32 mn elements), the 32-byte chunk index of a row XORed with the row index
// (address bits [5,7) ^= bits [7,9)).  Atoms of one k-group (4 k) sit 512 bytes apart along mn (LBO), k-groups
// 2048 bytes apart (SBO); one tf32 MMA (K = 8) consumes two k-groups.
__device__ __forceinline__ uint64_t make_smem_desc_mn(uint32_t saddr) {
    const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | ((512u >> 4) << 16);
    const uint32_t hi = (2048u >> 4) | (1u << 14) | (1u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// kind::tf32, D = f32, M = 128, N = n; bit 15 / 16 = A / B stored MN-major
__device__ __forceinline__ uint32_t make_idesc(uint32_t n, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
           ((n >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}

// fp32 -> (hi, lo) with hi, lo on the tf32 grid (10 explicit mantissa bits) and hi + lo = x up to 2^-22 |x|.
// Rounding is done with integer arithmetic on the bit pattern (add half an ulp of tf32 to the magnitude, clear the
// low 13 bits): 2 instructions instead of the ~7-instruction sequence ptxas emits for cvt.rna.tf32.f32 on sm_100a
// (ncu: that sequence was 60 % of all issued instructions of the first versions).  Finite inputs only, which is what
// the layers produce; x - hi is exact in fp32.
__device__ __forceinline__ uint32_t round_tf32_bits(uint32_t b) { return (b + 0x1000u) & 0xFFFFE000u; }
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = round_tf32_bits(__float_as_uint(x));
    lo = round_tf32_bits(__float_as_uint(x - __uint_as_float(hi)));
}

// ---- operand tile: global fp32 -> registers ------------------------------------------------------
// Tile = 128 (mn) x 32 (k).  Each of the 256 loader threads owns 4 float4.
//   T == false (k contiguous):  i-th float4 = row (t>>3)+32i, k-chunk c = t&7         (elements k = 4c..4c+3)
//   T == true (mn contiguous):  i-th float4 = k-row 8*(w&3)+(lane&7), mn-quad 16*(w>>2)+4i+(lane>>3)
// Per-thread row pointers are set up once per operand pair with the row index clamped into range: a clamped
// row only feeds output rows / columns that are never stored, so no row predicate is needed afterwards.
// Only the K direction must contribute exact zeros beyond k = K (last k-block of a pair).
// The kernel is instantiated per (AT, BT) and the code is kept small on purpose: the first version spent a
// fifth of its issue slots in instruction-cache misses (ncu: stall_no_inst).
template <bool T>
__device__ __forceinline__ void tile_ptrs(const float* __restrict__ p, int64_t ld, int64_t rows, int64_t r0,
                                          const float* (&ptr)[4]) {
    const int t = threadIdx.x & 255, lane = t & 31, w = t >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (!T) {
            int64_t r = r0 + (t >> 3) + 32 * i;
            r = r < rows ? r : rows - 1;
            ptr[i] = p + r * ld + (t & 7) * 4;
        } else {
            int64_t r = r0 + 4 * (16 * (w >> 2) + 4 * i + (lane >> 3));
            const int64_t last = ((rows - 1) >> 2) << 2;
            r = r < last ? r : last;
            ptr[i] = p + (int64_t)(8 * (w & 3) + (lane & 7)) * ld + r;
        }
    }
}

template <bool T>
__device__ __forceinline__ void tile_load(const float* const (&ptr)[4], int64_t ld, int64_t k0, int64_t kdim,
                                          float4 (&v)[4]) {
    const int t = threadIdx.x & 255, lane = t & 31, w = t >> 5;
    if (!T) {
        const int64_t kk = k0 + (t & 7) * 4;
        if (kk + 3 < kdim) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = ldg_f4(ptr[i] + k0);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (kk < kdim) {
                    const float* q = ptr[i] + k0;
                    x.x = __ldg(q);
                    if (kk + 1 < kdim) x.y = __ldg(q + 1);
                    if (kk + 2 < kdim) x.z = __ldg(q + 2);
                }
                v[i] = x;
            }
        }
    } else {
        const int64_t kk = k0 + 8 * (w & 3) + (lane & 7);
        if (kk < kdim) {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = ldg_f4(ptr[i] + k0 * ld);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

// ---- registers -> split -> swizzled K-major shared tiles -----------------------------------------
// byte offset of (row, 16-byte chunk c) inside a tile: row*128 + ((c ^ (row & 7)) << 4)
template <bool T>
__device__ __forceinline__ void tile_store(uint32_t s_hi, uint32_t s_lo, const float4 (&v)[4]) {
    const int t = threadIdx.x & 255, lane = t & 31, w = t >> 5;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_tf32(x[e], hi[e], lo[e]);
        if (!T) {
            const int row = (t >> 3) + 32 * i;
            const int c = t & 7;
            const uint32_t off = (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4));
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(s_hi + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(s_lo + off), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        } else {
            // MN-major tile: atom (k-group kg4 of 4 rows, mn-group mg of 32 floats) at (kg4*4 + mg)*512;
            // inside an atom: k-row r at r*128, 32-byte chunk c32 of the row stored at chunk (c32 ^ r)
            const int k = 8 * (w & 3) + (lane & 7);
            const int kg4 = k >> 2, r = k & 3;
            const int mq = 16 * (w >> 2) + 4 * i + (lane >> 3);          // mn-quad 0..31 (4 floats = 16 bytes)
            const int c16 = mq & 7;
            const uint32_t off = (uint32_t)((kg4 * 4 + (mq >> 3)) * 512 + r * 128 + ((((c16 >> 1) ^ r) << 5) | ((c16 & 1) << 4)));
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(s_hi + off), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(s_lo + off), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        }
    }
}

__device__ __forceinline__ float tc_epilogue_one(const TcParams& P, float v, int64_t r, int64_t c) {
    if (P.bias) v += __ldg(P.bias + c);
    v = apply_act(v, P.act, P.slope);
    if (P.gate) v *= act_grad_from_output(__ldg(P.gate + r * P.ldg + c), P.gate_act, P.slope);
    return v;
}

template <bool AT, bool BT>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;                 // 1024-byte aligned (swizzle atom)
    const uint32_t bars = tiles + TC_STAGES * TC_STAGE_BYTES;       // full[3], empty[3], acc, tmem slot
    const uint32_t bar_full = bars, bar_empty = bars + 8 * TC_STAGES, bar_acc = bars + 16 * TC_STAGES;
    const uint32_t tmem_slot = bar_acc + 8;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int64_t m0 = (int64_t)blockIdx.y * TC_BM, n0 = (int64_t)blockIdx.x * TC_BN;
    const int split = blockIdx.z;
    const int kb_beg = split * P.kblocks_per_split;
    const int kb_end = min(P.total_kblocks, kb_beg + P.kblocks_per_split);
    const int nkb = kb_end - kb_beg;

    const bool is_b = warp >= 8;
    float4 ring[3][4];
    const float* ptr[4];
    int ptr_pair = -1;
    auto issue = [&](int kb, float4 (&v)[4]) {
        int p = 0, local = kb;
        if (P.npairs > 1 && local >= P.kblocks[0]) { local -= P.kblocks[0]; p = 1; }
        if (ptr_pair != p) {
            if (!is_b) tile_ptrs<AT>(P.a[p], P.lda[p], P.m, m0, ptr);
            else tile_ptrs<BT>(P.b[p], P.ldb[p], P.n, n0, ptr);
            ptr_pair = p;
        }
        const int64_t k0 = (int64_t)local * TC_BK;
#if PLAGNN_TC_DEBUG
        if (P.debug == 1 && kb > kb_beg) return;
#endif
        if (!is_b) tile_load<AT>(ptr, P.lda[p], k0, P.k[p], v);
        else tile_load<BT>(ptr, P.ldb[p], k0, P.k[p], v);
    };
    // the first two k-blocks are requested before the barrier / TMEM set-up below, hiding ~1 us of load latency
    if (warp < TC_LOAD_WARPS) {
        if (nkb > 0) issue(kb_beg, ring[0]);
        if (nkb > 1) issue(kb_beg + 1, ring[1]);
    }

    if (t == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, TC_LOAD_WARPS * 32);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        fence_mbar_init();
    }
    if (warp == TC_LOAD_WARPS) tmem_alloc(tmem_slot, TC_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < TC_LOAD_WARPS) {
        // ================= loaders =================
        // Register ring of depth 3: while k-block `it` is split and stored, the loads of `it+1` and `it+2` are in
        // flight (64 KB per SM), which is what the ~1.2 us loaded L2 latency needs to keep the tensor core fed.
        // 16 loader warps (4 per scheduler) so that the fixed-latency ALU chains of the split interleave.
        auto commit = [&](int it, const float4 (&v)[4]) {
            const int s = it % TC_STAGES;
            const uint32_t ph = (uint32_t)((it / TC_STAGES) & 1);
            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
            const uint32_t st = tiles + s * TC_STAGE_BYTES;
#if PLAGNN_TC_DEBUG
            if (!(P.debug == 1 && it > 0))
#endif
            {
                if (!is_b) tile_store<AT>(st, st + TC_PART_BYTES, v);
                else tile_store<BT>(st + 2 * TC_PART_BYTES, st + 3 * TC_PART_BYTES, v);
            }
            fence_proxy_async_smem();
            mbar_arrive(bar_full + 8 * s);
        };
#pragma unroll 1
        for (int it = 0; it < nkb; it += 3) {
            if (it + 2 < nkb) issue(kb_beg + it + 2, ring[2]);
            commit(it, ring[0]);
            if (it + 1 < nkb) {
                if (it + 3 < nkb) issue(kb_beg + it + 3, ring[0]);
                commit(it + 1, ring[1]);
            }
            if (it + 2 < nkb) {
                if (it + 4 < nkb) issue(kb_beg + it + 4, ring[1]);
                commit(it + 2, ring[2]);
            }
        }

        // ================= epilogue =================
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int lg = warp & 3, cq = warp >> 2;      // TMEM lane group (warp % 4) and 32-column quarter
        const int64_t r = m0 + lg * 32 + lane;
        const bool direct = P.splits == 1;
        float* dst = direct ? P.c : P.partial + (int64_t)split * P.m * P.n;
        const int64_t ldd = direct ? P.ldc : P.n;
        const bool vec_out = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
        {
            // both accumulators of this warp's 32 rows x 32 columns in one round trip to tensor memory
            const int cbase = 32 * cq;
            if (n0 + cbase < P.n) {   // warp-uniform
                uint32_t acc[32], acc_small[32];
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)cbase, acc);
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(TC_BN + cbase), acc_small);
                tmem_ld_wait();
                if (r < P.m) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int64_t c = n0 + cbase + 4 * q;
                        if (c >= P.n) break;
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            v[e] = __uint_as_float(acc[4 * q + e]) + __uint_as_float(acc_small[4 * q + e]);
                            if (direct && c + e < P.n) v[e] = tc_epilogue_one(P, v[e], r, c + e);
                        }
                        if (vec_out && c + 3 < P.n) {
                            *reinterpret_cast<float4*>(dst + r * ldd + c) = make_float4(v[0], v[1], v[2], v[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (c + e < P.n) dst[r * ldd + c + e] = v[e];
                        }
                    }
                }
            }
        }
        tc_fence_before();
    } else {
        // ================= MMA issuer (one elected lane) =================
        if (lane == 0) {
            const int64_t nrem = P.n - n0;
            const uint32_t n_eff = nrem >= TC_BN ? TC_BN : (uint32_t)((nrem + 15) / 16 * 16);
            const uint32_t idesc = make_idesc(n_eff, AT, BT);
            // descriptor start-address step per K = 8: 32 bytes inside a K-major row, one 4096-byte k-group if MN-major
            constexpr uint64_t a_step = AT ? (4096u >> 4) : 2u, b_step = BT ? (4096u >> 4) : 2u;
            for (int it = 0; it < nkb; ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (uint32_t)((it / TC_STAGES) & 1);
                mbar_wait(bar_full + 8 * s, ph);
                tc_fence_after();
                const uint32_t st = tiles + s * TC_STAGE_BYTES;
                const uint64_t a_hi = AT ? make_smem_desc_mn(st) : make_smem_desc(st);
                const uint64_t a_lo = AT ? make_smem_desc_mn(st + TC_PART_BYTES) : make_smem_desc(st + TC_PART_BYTES);
                const uint64_t b_hi = BT ? make_smem_desc_mn(st + 2 * TC_PART_BYTES) : make_smem_desc(st + 2 * TC_PART_BYTES);
                const uint64_t b_lo = BT ? make_smem_desc_mn(st + 3 * TC_PART_BYTES) : make_smem_desc(st + 3 * TC_PART_BYTES);
#pragma unroll
                for (int kk = 0; kk < TC_BK / 8; ++kk) {
#if PLAGNN_TC_DEBUG
                    if (P.debug == 2) break;
#endif
                    const uint64_t adv_a = (uint64_t)kk * a_step, adv_b = (uint64_t)kk * b_step;
                    // The tensor core adds into the fp32 accumulator with truncation, so the error of a chain grows
                    // with the number of accumulations.  The two correction products go to their own accumulator
                    // (2^-11 of the magnitude: its truncation is invisible) and the main one sees a third of the adds.
                    const uint32_t acc_on = (it | kk) ? 1u : 0u;
                    umma_tf32(tmem_base + TC_BN, a_lo + adv_a, b_hi + adv_b, idesc, acc_on);
                    umma_tf32(tmem_base + TC_BN, a_hi + adv_a, b_lo + adv_b, idesc, 1u);
                    umma_tf32(tmem_base, a_hi + adv_a, b_hi + adv_b, idesc, acc_on);
                }
                umma_commit(bar_empty + 8 * s);     // frees the stage when these MMAs have read it
            }
            umma_commit(bar_acc);                   // accumulator complete
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == TC_LOAD_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TC_TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks tiles (tile = blockIdx.x + i * gridDim.x); four extra epilogue warps
// drain accumulator stage a while the loaders / issuer already work on the next tile in stage a^1 (TMEM holds two
// accumulator stages of 2 x 128 columns).  The register ring of the loaders simply runs on across tile boundaries,
// so there is no pipeline drain between tiles and the barrier / TMEM set-up is paid once per SM instead of per tile.
// ------------------------------------------------------------------------------------------------------
constexpr int TCP_EPI_WARPS = 4;
constexpr int TCP_THREADS = (TC_LOAD_WARPS + 1 + TCP_EPI_WARPS) * 32;
constexpr int TCP_TMEM_COLS = 512;

struct TileCoord {
    int64_t m0, n0;
    int split, kb_beg, nkb;
};
__device__ __forceinline__ TileCoord tile_coord(const TcParams& P, int tile) {
    TileCoord c;
    const int per_split = P.tiles_m * P.tiles_n;
    c.split = tile / per_split;
    const int rem = tile - c.split * per_split;
    const int mt = rem / P.tiles_n;
    c.m0 = (int64_t)mt * TC_BM;
    c.n0 = (int64_t)(rem - mt * P.tiles_n) * TC_BN;
    c.kb_beg = c.split * P.kblocks_per_split;
    c.nkb = min(P.total_kblocks, c.kb_beg + P.kblocks_per_split) - c.kb_beg;
    return c;
}

template <bool AT, bool BT>
__global__ void __launch_bounds__(TCP_THREADS, 1) gemm_tc_persistent_kernel(const TcParams P) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t tiles = (raw + 1023u) & ~1023u;
    const uint32_t bars = tiles + TC_STAGES * TC_STAGE_BYTES;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * TC_STAGES;
    const uint32_t bar_tfull = bars + 16 * TC_STAGES, bar_tempty = bar_tfull + 16;   // 2 accumulator stages each
    const uint32_t tmem_slot = bar_tempty + 16;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - raw));

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int first_tile = blockIdx.x, tile_step = gridDim.x;

    // ---- loader state: the issue cursor runs two k-blocks ahead of the commit cursor, across tile boundaries
    const bool is_b = warp >= 8;
    float4 ring[2][4];   // depth 2 here: 21 warps leave 80 registers per thread
    const float* ptr[4];
    int ptr_pair = -1;
    int is_tile = first_tile, is_kb = 0;
    TileCoord ic = tile_coord(P, first_tile < P.total_tiles ? first_tile : 0);
    auto issue = [&](float4 (&v)[4]) {
        if (is_tile >= P.total_tiles) return;
        const int kb = ic.kb_beg + is_kb;
        int p = 0, local = kb;
        if (P.npairs > 1 && local >= P.kblocks[0]) { local -= P.kblocks[0]; p = 1; }
        if (ptr_pair != p) {
            if (!is_b) tile_ptrs<AT>(P.a[p], P.lda[p], P.m, ic.m0, ptr);
            else tile_ptrs<BT>(P.b[p], P.ldb[p], P.n, ic.n0, ptr);
            ptr_pair = p;
        }
        const int64_t k0 = (int64_t)local * TC_BK;
        if (!is_b) tile_load<AT>(ptr, P.lda[p], k0, P.k[p], v);
        else tile_load<BT>(ptr, P.ldb[p], k0, P.k[p], v);
        if (++is_kb == ic.nkb) {
            is_tile += tile_step;
            is_kb = 0;
            ptr_pair = -1;
            if (is_tile < P.total_tiles) ic = tile_coord(P, is_tile);
        }
    };
    if (warp < TC_LOAD_WARPS) issue(ring[0]);          // first request before the barrier / TMEM set-up

    if (t == 0) {
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, TC_LOAD_WARPS * 32);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull + 8 * a, 1);
            mbar_init(bar_tempty + 8 * a, TCP_EPI_WARPS * 32);
        }
        fence_mbar_init();
    }
    if (warp == TC_LOAD_WARPS) tmem_alloc(tmem_slot, TCP_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp < TC_LOAD_WARPS) {
        // ================= loaders =================
        int total_it = 0;
        for (int tl = first_tile; tl < P.total_tiles; tl += tile_step) total_it += tile_coord(P, tl).nkb;
        auto commit = [&](int it, const float4 (&v)[4]) {
            const int s = it % TC_STAGES;
            const uint32_t ph = (uint32_t)((it / TC_STAGES) & 1);
            mbar_wait(bar_empty + 8 * s, ph ^ 1u);
            const uint32_t st = tiles + s * TC_STAGE_BYTES;
            if (!is_b) tile_store<AT>(st, st + TC_PART_BYTES, v);
            else tile_store<BT>(st + 2 * TC_PART_BYTES, st + 3 * TC_PART_BYTES, v);
            fence_proxy_async_smem();
            mbar_arrive(bar_full + 8 * s);
        };
#pragma unroll 1
        for (int it = 0; it < total_it; it += 2) {
            issue(ring[1]);
            commit(it, ring[0]);
            if (it + 1 < total_it) {
                issue(ring[0]);
                commit(it + 1, ring[1]);
            }
        }
    } else if (warp == TC_LOAD_WARPS) {
        // ================= MMA issuer (one elected lane) =================
        if (lane == 0) {
            constexpr uint64_t a_step = AT ? (4096u >> 4) : 2u, b_step = BT ? (4096u >> 4) : 2u;
            int it = 0, tl_idx = 0;
            for (int tl = first_tile; tl < P.total_tiles; tl += tile_step, ++tl_idx) {
                const TileCoord tc = tile_coord(P, tl);
                const int a = tl_idx & 1;
                mbar_wait(bar_tempty + 8 * a, (uint32_t)(((tl_idx >> 1) & 1) ^ 1));   // epilogue has drained this stage
                tc_fence_after();
                const int64_t nrem = P.n - tc.n0;
                const uint32_t n_eff = nrem >= TC_BN ? TC_BN : (uint32_t)((nrem + 15) / 16 * 16);
                const uint32_t idesc = make_idesc(n_eff, AT, BT);
                const uint32_t acc_big = tmem_base + (uint32_t)(a * 2 * TC_BN), acc_small = acc_big + TC_BN;
                for (int kb = 0; kb < tc.nkb; ++kb, ++it) {
                    const int s = it % TC_STAGES;
                    const uint32_t ph = (uint32_t)((it / TC_STAGES) & 1);
                    mbar_wait(bar_full + 8 * s, ph);
                    tc_fence_after();
                    const uint32_t st = tiles + s * TC_STAGE_BYTES;
                    const uint64_t a_hi = AT ? make_smem_desc_mn(st) : make_smem_desc(st);
                    const uint64_t a_lo = AT ? make_smem_desc_mn(st + TC_PART_BYTES) : make_smem_desc(st + TC_PART_BYTES);
                    const uint64_t b_hi = BT ? make_smem_desc_mn(st + 2 * TC_PART_BYTES) : make_smem_desc(st + 2 * TC_PART_BYTES);
                    const uint64_t b_lo = BT ? make_smem_desc_mn(st + 3 * TC_PART_BYTES) : make_smem_desc(st + 3 * TC_PART_BYTES);
#pragma unroll
                    for (int kk = 0; kk < TC_BK / 8; ++kk) {
                        const uint64_t adv_a = (uint64_t)kk * a_step, adv_b = (uint64_t)kk * b_step;
                        const uint32_t acc_on = (kb | kk) ? 1u : 0u;
                        umma_tf32(acc_small, a_lo + adv_a, b_hi + adv_b, idesc, acc_on);
                        umma_tf32(acc_small, a_hi + adv_a, b_lo + adv_b, idesc, 1u);
                        umma_tf32(acc_big, a_hi + adv_a, b_hi + adv_b, idesc, acc_on);
                    }
                    umma_commit(bar_empty + 8 * s);
                }
                umma_commit(bar_tfull + 8 * a);      // this tile's accumulators are complete
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue warps =================
        const int lg = warp & 3;                       // TMEM lane group this warp may touch (warp % 4)
        int tl_idx = 0;
        for (int tl = first_tile; tl < P.total_tiles; tl += tile_step, ++tl_idx) {
            const TileCoord tc = tile_coord(P, tl);
            const int a = tl_idx & 1;
            mbar_wait(bar_tfull + 8 * a, (uint32_t)((tl_idx >> 1) & 1));
            tc_fence_after();
            const int64_t r = tc.m0 + lg * 32 + lane;
            const bool direct = P.splits == 1;
            float* dst = direct ? P.c : P.partial + (int64_t)tc.split * P.m * P.n;
            const int64_t ldd = direct ? P.ldc : P.n;
            const bool vec_out = ((ldd & 3) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
            const uint32_t acc_big = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(a * 2 * TC_BN);
#pragma unroll 1
            for (int cq = 0; cq < 8; ++cq) {           // 16 columns at a time: 21 warps leave 80 registers per thread
                const int cbase = 16 * cq;
                if (tc.n0 + cbase >= P.n) break;       // warp-uniform
                uint32_t acc[16], acc_small[16];
                tmem_ld16(acc_big + (uint32_t)cbase, acc);
                tmem_ld16(acc_big + (uint32_t)(TC_BN + cbase), acc_small);
                tmem_ld_wait();
                if (r < P.m) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int64_t c = tc.n0 + cbase + 4 * q;
                        if (c >= P.n) break;
                        float v[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            v[e] = __uint_as_float(acc[4 * q + e]) + __uint_as_float(acc_small[4 * q + e]);
                            if (direct && c + e < P.n) v[e] = tc_epilogue_one(P, v[e], r, c + e);
                        }
                        if (vec_out && c + 3 < P.n) {
                            *reinterpret_cast<float4*>(dst + r * ldd + c) = make_float4(v[0], v[1], v[2], v[3]);
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                if (c + e < P.n) dst[r * ldd + c + e] = v[e];
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_tempty + 8 * a);           // all TMEM reads of this stage are done
        }
    }
    __syncthreads();
    if (warp == TC_LOAD_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TCP_TMEM_COLS);
    }
}

// ordered reduction of split-K partials + epilogue
__global__ void __launch_bounds__(256) gemm_tc_reduce_kernel(const TcParams P) {
    const int64_t total = P.m * P.n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int z = 0; z < P.splits; ++z) s += P.partial[(int64_t)z * total + i];
        const int64_t r = i / P.n, c = i - r * P.n;
        P.c[r * P.ldc + c] = tc_epilogue_one(P, s, r, c);
    }
}

// The tensor core accumulates in fp32 with truncation, so the error of one TMEM accumulation chain grows
// (with a bias) with its length: measured 1.8e-5 relative at K = 24 041 in one chain vs ~1e-6 at K <= 1 000.
// Chains are therefore capped at TC_MAX_CHAIN k-blocks (K = 1 280); longer contractions are split and the
// partials are summed in fp32 with round-to-nearest by the ordered reduction kernel.
constexpr int TC_MAX_CHAIN = 40;

static int tc_choose_splits(int64_t m, int64_t n, int total_kblocks) {
    const int64_t tiles = ceil_div(m, TC_BM) * ceil_div(n, TC_BN);
    const int sms = sm_count();
    int64_t s = 1;
    if (tiles < sms && total_kblocks >= 8) {
        s = sms / tiles;   // one wave: tiles * splits <= SM count (1 CTA per SM)
        if (s > total_kblocks / 4) s = total_kblocks / 4;
        if (s > 64) s = 64;
        if (s < 1) s = 1;
    }
    const int64_t for_accuracy = ceil_div(total_kblocks, TC_MAX_CHAIN);
    int64_t s0 = s > for_accuracy ? s : for_accuracy;
    if (s0 > 1) {
        // wave quantisation: 1 CTA per SM, so time ~ ceil(tiles*s/SMs) * ceil(kblocks/s); look a little around s0
        // (one step below is allowed: chains of up to 48 k-blocks keep the measured error below 4e-6)
        int64_t best = s0, best_cost = INT64_MAX;
        const int64_t lo = (s0 > 1 && ceil_div(total_kblocks, s0 - 1) <= 48) ? s0 - 1 : s0;
        for (int64_t c = lo; c <= s0 + 8 && c <= total_kblocks; ++c) {
            const int64_t cost = ceil_div(tiles * c, (int64_t)sms) * (ceil_div(total_kblocks, c) + 6);   // +6: prologue/epilogue
            if (cost < best_cost) { best_cost = cost; best = c; }
        }
        s0 = best;
    }
    return (int)s0;
}

size_t gemm_tc_workspace_bytes(int64_t m, int64_t n, int64_t k_total) {
    // k_total may be split over up to PLAGNN_GEMM_MAX_PAIRS pairs, each rounded up to whole k-blocks
    const int kb = (int)ceil_div(k_total, TC_BK);
    int s = 1;
    for (int extra = 0; extra <= PLAGNN_GEMM_MAX_PAIRS; ++extra) {
        const int c = tc_choose_splits(m, n, kb + extra);
        s = c > s ? c : s;
    }
    return s > 1 ? (size_t)s * (size_t)m * (size_t)n * sizeof(float) : 0;
}

// The tensor-core path takes 16-byte aligned operands (row pitch % 4 == 0) whose pairs share one storage
// order; anything else goes to the FFMA backend.
bool gemm_tc_eligible(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs) {
    if (n < 16 || m < 1) return false;
    for (int p = 0; p < npairs; ++p) {
        const plagnn_gemm_pair& q = pairs[p];
        if (q.k < 8) return false;
        if ((q.lda & 3) || (q.ldb & 3) || !aligned16(q.a) || !aligned16(q.b)) return false;
        if ((q.a_trans != 0) != (pairs[0].a_trans != 0) || (q.b_trans != 0) != (pairs[0].b_trans != 0)) return false;
    }
    return true;
}

int gemm_tc_launch(int64_t m, int64_t n, int32_t npairs, const plagnn_gemm_pair* pairs, const float* bias, int act,
                   float slope, const float* gate, int64_t ldg, int gate_act, float* c, int64_t ldc, void* workspace,
                   size_t workspace_bytes, cudaStream_t st) {
    TcParams P;
    P.m = m; P.n = n; P.npairs = npairs;
    P.total_kblocks = 0;
    for (int p = 0; p < PLAGNN_GEMM_MAX_PAIRS; ++p) {
        const plagnn_gemm_pair& q = pairs[p < npairs ? p : 0];
        P.a[p] = q.a; P.b[p] = q.b; P.lda[p] = q.lda; P.ldb[p] = q.ldb;
        P.a_trans[p] = q.a_trans ? 1 : 0; P.b_trans[p] = q.b_trans ? 1 : 0;
        P.a_vec[p] = ((q.lda & 3) == 0 && aligned16(q.a)) ? 1 : 0;
        P.b_vec[p] = ((q.ldb & 3) == 0 && aligned16(q.b)) ? 1 : 0;
        P.k[p] = p < npairs ? q.k : 0;
        P.kblocks[p] = p < npairs ? (int)ceil_div(q.k, TC_BK) : 0;
        P.total_kblocks += P.kblocks[p];
    }
    P.bias = bias; P.act = act; P.slope = slope; P.gate = gate; P.ldg = ldg; P.gate_act = gate_act;
    P.c = c; P.ldc = ldc;
    static const int dbg = [] { const char* e = getenv("PLAGNN_TC_DEBUG"); return e ? atoi(e) : 0; }();
    P.debug = dbg;
    int splits = tc_choose_splits(m, n, P.total_kblocks);
    if (splits > 1 && (!workspace || workspace_bytes < (size_t)splits * m * n * sizeof(float)))
        return fail(PLAGNN_ERR_WORKSPACE, "gemm_tc", "split-K workspace too small (see plagnn_gemm_workspace_bytes)");
    P.kblocks_per_split = (int)ceil_div(P.total_kblocks, splits);
    P.splits = (int)ceil_div(P.total_kblocks, P.kblocks_per_split);
    P.partial = P.splits > 1 ? (float*)workspace : nullptr;

    using KernelFn = void (*)(const TcParams);
    static const KernelFn kernels[8] = {gemm_tc_kernel<false, false>, gemm_tc_kernel<false, true>,
                                        gemm_tc_kernel<true, false>, gemm_tc_kernel<true, true>,
                                        gemm_tc_persistent_kernel<false, false>, gemm_tc_persistent_kernel<false, true>,
                                        gemm_tc_persistent_kernel<true, false>, gemm_tc_persistent_kernel<true, true>};
    // The persistent variant (overlapped epilogue) measured FASTER only for long contractions (8192^3: 142 vs 140 TF,
    // weight gradients 54 vs 52 TF) and SLOWER on the K ~ 500 layer shapes (76 vs 94 TF): with 21 warps only 80
    // registers per thread are available (2-deep instead of 3-deep load ring) and its 4 epilogue warps need longer
    // than a 16-k-block main loop.  It is therefore opt-in (PLAGNN_TC_PERSISTENT=1).
    static const int persistent = [] { const char* e = getenv("PLAGNN_TC_PERSISTENT"); return e ? atoi(e) : 0; }();
    P.tiles_m = (int)ceil_div(m, TC_BM);
    P.tiles_n = (int)ceil_div(n, TC_BN);
    P.total_tiles = P.tiles_m * P.tiles_n * P.splits;
    static thread_local int attr_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_dev != dev) {
        for (int i = 0; i < 8; ++i) {
            cudaError_t e = cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
            if (e != cudaSuccess) {
                set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return PLAGNN_ERR_CUDA;
            }
        }
        attr_dev = dev;
    }
    const int kidx = (P.a_trans[0] ? 2 : 0) + (P.b_trans[0] ? 1 : 0);
    if (persistent) {
        const int ctas = P.total_tiles < sm_count() ? P.total_tiles : sm_count();
        kernels[4 + kidx]<<<ctas, TCP_THREADS, TC_SMEM_BYTES, st>>>(P);
    } else {
        dim3 grid((unsigned)ceil_div(n, TC_BN), (unsigned)ceil_div(m, TC_BM), (unsigned)P.splits);
        kernels[kidx]<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(P);
    }
    if (P.splits > 1) {
        const int64_t total = m * n;
        const int g = (int)(ceil_div(total, 256) < (int64_t)sm_count() * 8 ? ceil_div(total, 256) : (int64_t)sm_count() * 8);
        gemm_tc_reduce_kernel<<<g, 256, 0, st>>>(P);
    }
    return check_launch("gemm_tc", P.splits > 1 ? 2 : 1);
}

}  // namespace plagnn
