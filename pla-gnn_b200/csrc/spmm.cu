// K2 — neighbourhood aggregation (SpMM with a pluggable reducer) for sm_100a.
//
// Replaces DGL's gspmm behind  code/model.py:20,22,24  (SAGEConv 'pool': update_all(copy_u, max))
// and its autograd backward at  code/train.py:204;  the sum / u_mul_e / mean members of the same
// kernel family serve BASELINE.json configs[3] (weighted synthetic graph).
//
// Shape of the kernel (HBM / L2 bound, no tensor cores):
//   * one warp owns one work item = (destination row, chunk of <= `chunk` in-edges) from the plan
//     built in graph_build.cu, so power-law hubs are spread over many warps;
//   * a lane owns VEC float4 column groups (lane, lane+S, ...): every neighbour row is read with
//     fully coalesced LDG.128 requests; NB neighbour rows are in flight per lane;
//   * sum / match reducers read neighbour ids 32 at a time and broadcast them with shuffles and mask the
//     slots past the end; the max reducer (the PLA-GNN path) reads them as broadcast loads one step ahead
//     and needs no masks at all (see the comment in the kernel);
//   * rows split over several chunks write (value,arg) partials that a second small kernel folds
//     in chunk order, which keeps "first maximum wins" and makes sums order-stable.
#include "common.cuh"
#include <math.h>
#include <climits>
#include <cstdlib>

namespace plagnn {

void plan_pointers(const void* plan, int64_t n_rows, const int32_t** item_ptr, const int32_t** slot_ptr,
                   const int32_t** item_row, const int32_t** hub_rows);

enum { MODE_MAX = 0, MODE_SUM = 1, MODE_MATCH = 2 };
constexpr int SPMM_WARPS = 8;

struct SpmmEpilogue {
    const float* scale;   // per destination row, nullable
    const float* bias;    // per column, nullable
    int act;
    float slope;
    float dropout_p;
    unsigned long long seed;
};

// Source-slab passes (L2 blocking of the gather): the in-edges of every row are split by SOURCE range into several CSR
// structures, one pass per slab, so that the rows a pass gathers from (a slab of the feature matrix) stay L2-resident.
// A pass continues from the running result of the earlier slabs (`prev_val` / `prev_arg`, may alias the outputs) and only the
// last one applies the epilogue (sum) or turns "no in-edge at all" into 0 (max).  On value ties the earlier slab wins.
struct SpmmChain {
    const float* prev_val;     // nullable: first slab
    const int32_t* prev_arg;   // max reducer only
    int64_t ldprev;
    int last;                  // 1: finalise
};

__device__ __forceinline__ unsigned hash_u32(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (unsigned)((z ^ (z >> 31)) >> 32);
}
__device__ __forceinline__ float dropout_keep_scale(unsigned long long seed, long long row, int col, int64_t feat,
                                                    float p) {
    const float u = hash_u32(seed, (unsigned long long)row * (unsigned long long)feat + col) * (1.0f / 4294967296.0f);
    return u >= p ? 1.0f / (1.0f - p) : 0.0f;
}

__device__ __forceinline__ float4 sum_epilogue(float4 a, const SpmmEpilogue& ep, long long row, int col, int64_t feat) {
    float v[4] = {a.x, a.y, a.z, a.w};
    const float s = ep.scale ? __ldg(ep.scale + row) : 1.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float t = ep.scale ? v[i] * s : v[i];
        if (ep.bias && col + i < feat) t += __ldg(ep.bias + col + i);
        t = apply_act(t, ep.act, ep.slope);
        if (ep.dropout_p > 0.f) t *= dropout_keep_scale(ep.seed, row, col + i, feat, ep.dropout_p);
        v[i] = t;
    }
    return make_float4(v[0], v[1], v[2], v[3]);
}

// if (v > acc) { acc = v; arg = u; } with the work spread over two pipes.  Written in C the compiler emits FSETP + FSEL + SEL,
// three ALU-pipe instructions per gathered element, and the max reducer is bound by that pipe (ncu: ALU 76 %, FMA 16 %).
// Here the value moves with a predicated FFMA on the FMA pipe: acc = v * 1.0f + (-0.0f), exact for every v (the two
// constants arrive as kernel arguments so that ptxas cannot fold the multiply-add back into a select).
__device__ __forceinline__ void max_update(float& acc, int& arg, const float v, const int u, const float one, const float nzero) {
    float na;
    int ng;
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %2, %3;\n\tmov.f32 %0, %3;\n\t@p fma.rn.f32 %0, %2, %5, %6;\n\tselp.b32 %1, %4, %7, p;\n\t}"
        : "=&f"(na), "=r"(ng)
        : "f"(v), "f"(acc), "r"(u), "f"(one), "f"(nzero), "r"(arg));
    acc = na;
    arg = ng;
}

// base + u * pitch_bytes as one IMAD.WIDE.U32 (u: node id >= 0)
__device__ __forceinline__ const float* row_ptr(const float* base, int u, unsigned pitch_bytes) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"((unsigned)u), "r"(pitch_bytes), "l"((unsigned long long)base));
    return reinterpret_cast<const float*>(r);
}

template <int MODE>
__device__ __forceinline__ void reduce_one(float4& acc, int4& arg, const float4 v, const int u, const float one, const float nzero) {
    if (MODE == MODE_MAX) {
        max_update(acc.x, arg.x, v.x, u, one, nzero);
        max_update(acc.y, arg.y, v.y, u, one, nzero);
        max_update(acc.z, arg.z, v.z, u, one, nzero);
        max_update(acc.w, arg.w, v.w, u, one, nzero);
    } else {
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
}

// MODE_MAX  : x = features,  out/arg written
// MODE_SUM  : x = features,  optional edge weights, epilogue
// MODE_MATCH: x = dz (rows = destinations v of the out-edge u->v), argm = arg[v,:], zfwd = z[v,:] (nullable);
//             acc[u,f] += (argm[v,f]==u && z>0) ? dz[v,f] : 0
// The max reducer runs at the L2 -> SM bandwidth cap once enough warps are resident (measured: 24 warps/SM at 80
// registers 0.198 ms, 16 warps at 110 registers 0.322 ms for F = 503), so its default variants are held to 3 blocks per SM;
// the sum reducer (no arg registers) to 4 blocks = 32 warps (weighted 100 M-edge graph: 55.7 ms at 64 registers, 71.3 at 72).
template <int MODE, int VEC, int NB, bool SUM_LEAN = false, int WARPS = SPMM_WARPS>
__global__ void __launch_bounds__(WARPS * 32, (NB * VEC > 8 ? 1 : MODE == MODE_MAX ? 3 : MODE == MODE_SUM ? 4 : 1) * (SPMM_WARPS / WARPS))
spmm_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const int32_t* __restrict__ eids,
            const float* __restrict__ ew, const int32_t* __restrict__ plan_hdr, const int32_t* __restrict__ item_ptr,
            const int32_t* __restrict__ slot_ptr, const int32_t* __restrict__ item_row, int item_begin, int n_items,
            const float* __restrict__ x, int64_t ldx, int feat, const int32_t* __restrict__ argm, int64_t ldarg,
            const float* __restrict__ zfwd, int64_t ldzf, float* __restrict__ out, int32_t* __restrict__ arg_out, int64_t ldo,
            float* __restrict__ part_val, int32_t* __restrict__ part_arg, int part_ld, SpmmEpilogue ep, float one,
            float nzero) {
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int item = item_begin + blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    pdl_wait();
    // item_row holds one int4 record per work item, {row, first in-edge, end in-edge, partial slot or -1} (graph_build.cu): the
    // narrow kernel below starts from that one load; here, with rows of >= 512 bytes and items that live for thousands of
    // cycles, the three-step walk costs nothing measurable and keeps the register allocation the tuned loops were measured with
    const int chunk = __ldg(plan_hdr);
    const int row = __ldg(item_row + 4 * item);
    const int first = __ldg(item_ptr + row);
    const int nch = __ldg(item_ptr + row + 1) - first;
    const int k = item - first;
    const int rbeg = __ldg(indptr + row), rend = __ldg(indptr + row + 1);
    const int beg = rbeg + k * chunk;
    const int end = min(rend, beg + chunk);

    // Column groups (float4) of this block's slab are dealt to the lanes with stride S = ceil(groups / VEC): lane l owns
    // groups l, l + S, l + 2S, ...  For F = 400 (100 groups, VEC 4) that is 26 lanes x 3 full groups + 22 lanes; only the last
    // group of a lane can fall outside the slab.  Lanes >= S shadow lane 0 (same addresses, nothing stored).
    const int col0 = blockIdx.y * (128 * VEC);   // first column of this block's slab
    const int groups = min((feat + 3) / 4 - blockIdx.y * (32 * VEC), 32 * VEC);
    // max reducer: runtime stride, clamped columns (its loads are unconditional); the other reducers keep the fixed
    // stride of 32 with compile-time column offsets from one row pointer and guard their loads with cok[]
    const int S = MODE == MODE_MAX ? min(32, ((groups + VEC - 1) / VEC + 1) & ~1) : 32;   // even: whole 32-byte sectors
    const int lane_c = MODE == MODE_MAX ? ((lane < S && lane < groups) ? lane : 0) : lane;
    int col[VEC];
    bool cok[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
        if (MODE == MODE_MAX) {
            cok[q] = lane < S && lane_c + S * q < groups;
            col[q] = col0 + 4 * (lane_c + S * (lane_c + S * q < groups ? q : 0));   // clamped: always a readable column
        } else {
            col[q] = col0 + 4 * (lane + 32 * q);
            cok[q] = col[q] < feat;
        }
    }

    float4 acc[VEC];
    int4 arg[VEC];
    const float init = MODE == MODE_MAX ? -INFINITY : 0.f;
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
        acc[q] = make_float4(init, init, init, init);
        arg[q] = make_int4(-1, -1, -1, -1);
    }

    if (MODE == MODE_MAX) {
        // Unmasked inner loop: a slot past the end of the chunk repeats the chunk's last neighbour and a lane without
        // columns re-reads lane 0's — a repeated value never beats the running maximum (strict >), so neither needs
        // a select, and the loads carry no predicates.  One IMAD.WIDE per load forms the address.
        const float* xq[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) xq[q] = x + col[q];
        const unsigned ldx_bytes = (unsigned)ldx * 4u;      // host checks ldx < 2^30
        // neighbour ids: every lane reads the same word (one broadcast request, L1-resident after the first touch of a
        // sector) one step ahead of the rows it addresses — no shuffles, so no divergence side path in the loop
        if (beg < end) {
            const int last = end - 1;
            int u[NB];
#pragma unroll
            for (int t = 0; t < NB; ++t) u[t] = __ldg(indices + min(beg + t, last));
            for (int j = beg; j < end; j += NB) {
                int un[NB];
                float4 v[NB][VEC];
#pragma unroll
                for (int t = 0; t < NB; ++t) {
                    un[t] = __ldg(indices + min(j + NB + t, last));
#pragma unroll
                    for (int q = 0; q < VEC; ++q) v[t][q] = ldg_f4(row_ptr(xq[q], u[t], ldx_bytes));
                }
#pragma unroll
                for (int t = 0; t < NB; ++t)
#pragma unroll
                    for (int q = 0; q < VEC; ++q) reduce_one<MODE>(acc[q], arg[q], v[t][q], u[t], one, nzero);
#pragma unroll
                for (int t = 0; t < NB; ++t) u[t] = un[t];
            }
        }
    } else if (MODE == MODE_SUM && SUM_LEAN) {
        // The sum reducer with the max reducer's loop shape (round 2): neighbour ids and weights are broadcast loads issued one
        // step ahead, a slot past the end re-reads the chunk's last neighbour with weight 0, a lane whose column group lies
        // past the row reads column group 0 (never stored), and the row address is one 32-bit multiply-add — no shuffles, no
        // load predicates, no 64-bit address arithmetic in the loop.  An experiment that lost (see launch_main): opt-in only.
        const float* xq[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) xq[q] = x + (cok[q] ? col[q] : col0);
        const unsigned ldx_bytes = (unsigned)ldx * 4u;      // host checks ldx < 2^30
        if (beg < end) {
            const int last = end - 1;
            int u[NB];
            float w[NB];
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                const int idx = min(beg + t, last);
                u[t] = __ldg(indices + idx);
                w[t] = beg + t <= last ? (ew ? __ldg(ew + (eids ? __ldg(eids + idx) : idx)) : 1.f) : 0.f;
            }
            for (int j = beg; j < end; j += NB) {
                int un[NB];
                float wn[NB];
                float4 v[NB][VEC];
#pragma unroll
                for (int t = 0; t < NB; ++t) {
                    const int idx = min(j + NB + t, last);
                    un[t] = __ldg(indices + idx);
                    wn[t] = j + NB + t <= last ? (ew ? __ldg(ew + (eids ? __ldg(eids + idx) : idx)) : 1.f) : 0.f;
#pragma unroll
                    for (int q = 0; q < VEC; ++q) v[t][q] = ldg_f4(row_ptr(xq[q], u[t], ldx_bytes));
                }
#pragma unroll
                for (int t = 0; t < NB; ++t)
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        acc[q].x = fmaf(w[t], v[t][q].x, acc[q].x); acc[q].y = fmaf(w[t], v[t][q].y, acc[q].y);
                        acc[q].z = fmaf(w[t], v[t][q].z, acc[q].z); acc[q].w = fmaf(w[t], v[t][q].w, acc[q].w);
                    }
#pragma unroll
                for (int t = 0; t < NB; ++t) { u[t] = un[t]; w[t] = wn[t]; }
            }
        }
    } else {
    for (int base = beg; base < end; base += 32) {
        const int cnt = min(32, end - base);
        int my_u = 0;
        float my_w = 1.f;
        if (lane < cnt) {
            my_u = __ldg(indices + base + lane);
            if (MODE == MODE_SUM && ew) my_w = __ldg(ew + (eids ? __ldg(eids + base + lane) : base + lane));
        }
        for (int j = 0; j < cnt; j += NB) {
            int u[NB];
            float w[NB];
            float4 v[NB][VEC];
            int4 am[NB][VEC];
            float4 zz[NB][VEC];
#pragma unroll
            for (int t = 0; t < NB; ++t) {
                const int src_lane = min(j + t, 31);
                u[t] = __shfl_sync(0xffffffffu, my_u, src_lane);
                w[t] = __shfl_sync(0xffffffffu, my_w, src_lane);
                const bool live = (j + t) < cnt;
                const float* xr = x + (int64_t)u[t] * ldx;
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    if (live && cok[q]) {
                        v[t][q] = ldg_f4(xr + col[q]);
                        if (MODE == MODE_MATCH) {
                            am[t][q] = __ldg(reinterpret_cast<const int4*>(argm + (int64_t)u[t] * ldarg + col[q]));
                            if (zfwd) zz[t][q] = ldg_f4(zfwd + (int64_t)u[t] * ldzf + col[q]);
                        }
                    } else {
                        v[t][q] = make_float4(init, init, init, init);
                        if (MODE == MODE_MATCH) am[t][q] = make_int4(-2, -2, -2, -2);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < NB; ++t) {
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    float4 val = v[t][q];
                    if (MODE == MODE_SUM && ew) {
                        val.x *= w[t]; val.y *= w[t]; val.z *= w[t]; val.w *= w[t];
                    }
                    if (MODE == MODE_MATCH) {
                        const int4 a = am[t][q];
                        float4 g = val;
                        if (zfwd && (j + t) < cnt && cok[q]) {
                            const float4 z4 = zz[t][q];
                            g.x = z4.x > 0.f ? g.x : 0.f; g.y = z4.y > 0.f ? g.y : 0.f;
                            g.z = z4.z > 0.f ? g.z : 0.f; g.w = z4.w > 0.f ? g.w : 0.f;
                        }
                        val.x = a.x == row ? g.x : 0.f; val.y = a.y == row ? g.y : 0.f;
                        val.z = a.z == row ? g.z : 0.f; val.w = a.w == row ? g.w : 0.f;
                    }
                    reduce_one<MODE>(acc[q], arg[q], val, u[t], one, nzero);
                }
            }
        }
    }
    }

    // ---- write: final result for unsplit rows, ordered partial otherwise --------------------
    if (nch == 1) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            if (!cok[q]) continue;
            float4 r = acc[q];
            if (MODE == MODE_MAX) {
                const int4 a = arg[q];
                r.x = a.x < 0 ? 0.f : r.x; r.y = a.y < 0 ? 0.f : r.y;
                r.z = a.z < 0 ? 0.f : r.z; r.w = a.w < 0 ? 0.f : r.w;
                *reinterpret_cast<int4*>(arg_out + (int64_t)row * ldo + col[q]) = a;
            } else if (MODE == MODE_SUM) {
                r = sum_epilogue(r, ep, row, col[q], feat);
            }
            *reinterpret_cast<float4*>(out + (int64_t)row * ldo + col[q]) = r;
        }
    } else {
        const int64_t slot = (int64_t)__ldg(slot_ptr + row) + k;
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            if (!cok[q]) continue;
            *reinterpret_cast<float4*>(part_val + slot * part_ld + col[q]) = acc[q];
            if (MODE == MODE_MAX) *reinterpret_cast<int4*>(part_arg + slot * part_ld + col[q]) = arg[q];
        }
    }
}

// Narrow rows (feat <= 64: the column slice of one GPU in the feature partition, 128 or 256 bytes per row).  With one
// float4 per lane a row needs only G = 8 or 16 lanes, so the warp's S = 32 / G lane groups each take a different in-edge of
// the same work item: edge e of a batch of 32 goes to group e % S.  Every LDG.128 of the warp then fetches S neighbour rows
// (512 bytes per request, as in the wide kernel, instead of 128), and the S partial results are folded with shuffles at the
// end.  The max reducer keeps "first maximum in in-edge order wins" across the groups by carrying the winning edge position.
//
// Measured on the 1 M-node / 100 M-edge graph, 32 columns (tools/spmm_narrow_time.py): the wide kernel with 24 idle lanes
// 5.99 ms per aggregation, the first version of this kernel 2.55 ms, this one **1.99 ms** (6.9 TB/s algorithmic; 64 columns:
// 4.0 -> 2.94 ms).  What paid was fewer instructions per edge (ncu: 11 warp instructions per edge in the first version, mostly
// 64-bit address arithmetic and "live ? load : init" selects; the loop below has ~3), as in round 1's max reducer.  What did
// NOT pay, all measured against the 2.55 ms version, were three attempts to put more bytes in flight: (a) cutting the gather by source
// range into L2-sized slabs (plagnn_spmm_*_slab, kept as an option): 2.85 / 3.72 / 5.92 ms at 2 / 4 / 8 slabs — the items get
// shorter; (b) one warp walking 32 consecutive items with prefetched meta data and neighbour ids, rows still gathered into
// registers: 3.72 ms — in-order issue serialises the items of a warp; (c) the same walk with the rows of three 32-edge
// batches staged in shared memory by cp.async (96 rows in flight per warp, 16 warps per SM): 4.51 ms (max 5.39) — the extra
// LDS / STS / shuffle traffic and the three-phase step cost more than the deeper queue gains.  Staging pays when a staged row
// is reused; a gathered neighbour row is consumed once.
// (d) (last session of round 2) software pipelining in registers: two rounds of U gathers in flight per lane across batch
// boundaries, the ids / weights of batch b + 1 requested before batch b is gathered — bit-identical results, and slower in every
// configuration: U = 4 at 32 / 24 / 16 warps per SM 3.01 / 3.94 / 6.47 ms, U = 2 at 32 warps (the SAME number of loads in flight
// as this kernel, plus the id prefetch) 3.11 ms, against 1.99 ms here.  So the kernel is not bound by a chain of exposed
// latencies per warp but by how many row requests an SM keeps outstanding (~14 KB in flight per SM at both this and the PPI
// shape); more resident warps with short loops beat deeper loops.  L2 eviction priorities (ids / weights evict_first, rows
// evict_last through createpolicy + ld.global.L2::cache_hint) changed nothing either (tools/spmm_variants_time.py).
template <int MODE, int G, int U, int OCC, int WARPS = SPMM_WARPS>
__global__ void __launch_bounds__(WARPS * 32, OCC * SPMM_WARPS / WARPS)
spmm_narrow_kernel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const int32_t* __restrict__ eids,
                   const float* __restrict__ ew, const int32_t* __restrict__ plan_hdr, const int32_t* __restrict__ item_ptr,
                   const int32_t* __restrict__ slot_ptr, const int32_t* __restrict__ item_row, int item_begin, int n_items,
                   const float* __restrict__ x, int64_t ldx, int feat, float* __restrict__ out, int32_t* __restrict__ arg_out,
                   int64_t ldo, float* __restrict__ part_val, int32_t* __restrict__ part_arg, int part_ld, SpmmEpilogue ep,
                   SpmmChain ch) {
    static_assert(MODE == MODE_MAX || MODE == MODE_SUM, "narrow kernel: max and sum reducers");
    constexpr int S = 32 / G;          // lane groups = in-edges in flight per load instruction
    // U rounds unrolled = U rows in flight per lane (U = 4 at 32 warps per SM: 64 KB in flight per SM)
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int item = item_begin + blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    pdl_wait();
    const int4 rec = __ldg(reinterpret_cast<const int4*>(item_row) + item);       // {row, first in-edge, end in-edge, slot or -1}
    const int row = rec.x, beg = rec.y, end = rec.z;
    const int nch = rec.w < 0 ? 1 : 2;
    const int sub = lane / G, gl = lane % G;
    const int col = 4 * gl;
    const bool cok = col < feat;
    const float* xc = x + (cok ? col : 0);                  // lanes past the last column re-read column 0 (nothing stored)
    const float init = MODE == MODE_MAX ? -INFINITY : 0.f;
    float4 acc = make_float4(init, init, init, init);
    int4 arg = make_int4(-1, -1, -1, -1);
    int4 pos = make_int4(INT_MAX, INT_MAX, INT_MAX, INT_MAX);
    // continue from the earlier slabs' result (unsplit rows only: split rows take it in the combine kernel).  Plain loads:
    // prev may alias the output this kernel writes.  Position -1 makes the earlier slab win value ties.
    if (ch.prev_val && nch == 1 && cok && (MODE == MODE_MAX || sub == 0)) {
        acc = *reinterpret_cast<const float4*>(ch.prev_val + (int64_t)row * ch.ldprev + col);
        if (MODE == MODE_MAX) {
            arg = *reinterpret_cast<const int4*>(ch.prev_arg + (int64_t)row * ch.ldprev + col);
            pos = make_int4(-1, -1, -1, -1);
        }
    }

    // No predicates in the inner loop (the first version spent ~8 instructions per edge on 64-bit address arithmetic and on
    // "live ? load : init" selects): a slot past the end of the batch re-reads the batch's last neighbour with weight 0 (sum)
    // or as a repeated value that cannot beat a strict > (max), and the row address is one 32-bit multiply-add.
    const unsigned ldx_bytes = (unsigned)ldx * 4u;          // host checks ldx < 2^30
    for (int base = beg; base < end; base += 32) {
        const int cnt = min(32, end - base);
        const int my_u = __ldg(indices + min(base + lane, end - 1));
        float my_w = 0.f;
        if (MODE == MODE_SUM && lane < cnt) my_w = ew ? __ldg(ew + (eids ? __ldg(eids + base + lane) : base + lane)) : 1.f;
        for (int t = 0; t * S < cnt; t += U) {
            int u[U];
            float w[U];
            float4 v[U];
#pragma unroll
            for (int i = 0; i < U; ++i) {
                const int e = (t + i) * S + sub;            // <= 31: t advances in steps of U and 32 / S is a multiple of U
                u[i] = __shfl_sync(0xffffffffu, my_u, e);
                if (MODE == MODE_SUM) w[i] = __shfl_sync(0xffffffffu, my_w, e);
                v[i] = ldg_f4(row_ptr(xc, u[i], ldx_bytes));
            }
#pragma unroll
            for (int i = 0; i < U; ++i) {
                if (MODE == MODE_MAX) {
                    const int p = base + (t + i) * S + sub;
                    if (v[i].x > acc.x) { acc.x = v[i].x; arg.x = u[i]; pos.x = p; }
                    if (v[i].y > acc.y) { acc.y = v[i].y; arg.y = u[i]; pos.y = p; }
                    if (v[i].z > acc.z) { acc.z = v[i].z; arg.z = u[i]; pos.z = p; }
                    if (v[i].w > acc.w) { acc.w = v[i].w; arg.w = u[i]; pos.w = p; }
                } else {
                    acc.x = fmaf(w[i], v[i].x, acc.x); acc.y = fmaf(w[i], v[i].y, acc.y);
                    acc.z = fmaf(w[i], v[i].z, acc.z); acc.w = fmaf(w[i], v[i].w, acc.w);
                }
            }
        }
    }
    // fold the S lane groups (group 0 ends up with the result)
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
        const float ox = __shfl_xor_sync(0xffffffffu, acc.x, off), oy = __shfl_xor_sync(0xffffffffu, acc.y, off);
        const float oz = __shfl_xor_sync(0xffffffffu, acc.z, off), ow = __shfl_xor_sync(0xffffffffu, acc.w, off);
        if (MODE == MODE_MAX) {
            const int ax = __shfl_xor_sync(0xffffffffu, arg.x, off), ay = __shfl_xor_sync(0xffffffffu, arg.y, off);
            const int az = __shfl_xor_sync(0xffffffffu, arg.z, off), aw = __shfl_xor_sync(0xffffffffu, arg.w, off);
            const int px = __shfl_xor_sync(0xffffffffu, pos.x, off), py = __shfl_xor_sync(0xffffffffu, pos.y, off);
            const int pz = __shfl_xor_sync(0xffffffffu, pos.z, off), pw = __shfl_xor_sync(0xffffffffu, pos.w, off);
            if (ox > acc.x || (ox == acc.x && px < pos.x)) { acc.x = ox; arg.x = ax; pos.x = px; }
            if (oy > acc.y || (oy == acc.y && py < pos.y)) { acc.y = oy; arg.y = ay; pos.y = py; }
            if (oz > acc.z || (oz == acc.z && pz < pos.z)) { acc.z = oz; arg.z = az; pos.z = pz; }
            if (ow > acc.w || (ow == acc.w && pw < pos.w)) { acc.w = ow; arg.w = aw; pos.w = pw; }
        } else {
            acc.x += ox; acc.y += oy; acc.z += oz; acc.w += ow;
        }
    }
    if (sub != 0 || !cok) return;
    if (nch == 1) {
        float4 r = acc;
        if (MODE == MODE_MAX) {
            if (ch.last) {
                r.x = arg.x < 0 ? 0.f : r.x; r.y = arg.y < 0 ? 0.f : r.y;
                r.z = arg.z < 0 ? 0.f : r.z; r.w = arg.w < 0 ? 0.f : r.w;
            }
            *reinterpret_cast<int4*>(arg_out + (int64_t)row * ldo + col) = arg;
        } else if (ch.last) {
            r = sum_epilogue(r, ep, row, col, feat);
        }
        *reinterpret_cast<float4*>(out + (int64_t)row * ldo + col) = r;
    } else {
        const int64_t slot = rec.w;
        *reinterpret_cast<float4*>(part_val + slot * part_ld + col) = acc;
        if (MODE == MODE_MAX) *reinterpret_cast<int4*>(part_arg + slot * part_ld + col) = arg;
    }
}

// folds the partials of split rows in chunk order: one warp per (split row, 128-column group)
template <int MODE>
__global__ void __launch_bounds__(SPMM_WARPS * 32)
spmm_combine_kernel(const int32_t* __restrict__ item_ptr, const int32_t* __restrict__ slot_ptr,
                    const int32_t* __restrict__ hub_rows, int hub_begin, int n_hubs, int feat,
                    const float* __restrict__ part_val,
                    const int32_t* __restrict__ part_arg, int part_ld, float* __restrict__ out,
                    int32_t* __restrict__ arg_out, int64_t ldo, SpmmEpilogue ep, SpmmChain ch) {
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int h = hub_begin + blockIdx.x * SPMM_WARPS + (threadIdx.x >> 5);
    if (h >= n_hubs) return;
    pdl_wait();
    const int col = blockIdx.y * 128 + lane * 4;
    if (col >= feat) return;
    const int row = __ldg(hub_rows + h);
    const int nch = __ldg(item_ptr + row + 1) - __ldg(item_ptr + row);
    const int64_t slot0 = __ldg(slot_ptr + row);
    const float init = MODE == MODE_MAX ? -INFINITY : 0.f;
    float4 acc = make_float4(init, init, init, init);
    int4 arg = make_int4(-1, -1, -1, -1);
    if (ch.prev_val) {      // running result of the earlier source slabs (strict > below: it wins value ties)
        acc = *reinterpret_cast<const float4*>(ch.prev_val + (int64_t)row * ch.ldprev + col);
        if (MODE == MODE_MAX) arg = *reinterpret_cast<const int4*>(ch.prev_arg + (int64_t)row * ch.ldprev + col);
    }
    constexpr int CU = 8;   // partials in flight per lane (a hub of 16k edges has > 100 partials)
    for (int k0 = 0; k0 < nch; k0 += CU) {
        float4 v[CU];
        int4 a[CU];
#pragma unroll
        for (int j = 0; j < CU; ++j) {
            if (k0 + j < nch) {
                v[j] = __ldg(reinterpret_cast<const float4*>(part_val + (slot0 + k0 + j) * part_ld + col));
                if (MODE == MODE_MAX) a[j] = __ldg(reinterpret_cast<const int4*>(part_arg + (slot0 + k0 + j) * part_ld + col));
            } else {
                v[j] = make_float4(init, init, init, init);
                a[j] = make_int4(-1, -1, -1, -1);
            }
        }
#pragma unroll
        for (int j = 0; j < CU; ++j) {
            if (MODE == MODE_MAX) {
                if (v[j].x > acc.x) { acc.x = v[j].x; arg.x = a[j].x; }
                if (v[j].y > acc.y) { acc.y = v[j].y; arg.y = a[j].y; }
                if (v[j].z > acc.z) { acc.z = v[j].z; arg.z = a[j].z; }
                if (v[j].w > acc.w) { acc.w = v[j].w; arg.w = a[j].w; }
            } else {
                acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w;
            }
        }
    }
    if (MODE == MODE_MAX) {
        if (ch.last) {
            acc.x = arg.x < 0 ? 0.f : acc.x; acc.y = arg.y < 0 ? 0.f : acc.y;
            acc.z = arg.z < 0 ? 0.f : acc.z; acc.w = arg.w < 0 ? 0.f : acc.w;
        }
        *reinterpret_cast<int4*>(arg_out + (int64_t)row * ldo + col) = arg;
    } else if (MODE == MODE_SUM && ch.last) {
        acc = sum_epilogue(acc, ep, row, col, feat);
    }
    *reinterpret_cast<float4*>(out + (int64_t)row * ldo + col) = acc;
}

// reverse of the max reducer with fp32 reductions to global memory (RED.E.ADD.F32)
__global__ void __launch_bounds__(256)
spmm_max_scatter_kernel(const float* __restrict__ dz, int64_t lddz, const int32_t* __restrict__ arg, int64_t ldarg,
                        const float* __restrict__ z, int64_t ldz, int64_t n_rows, int feat, float* __restrict__ dx,
                        int64_t lddx) {
    pdl_enter();
    const int f4 = (feat + 3) >> 2;
    const int64_t total = n_rows * f4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = i / f4;
        const int c = (int)(i - v * f4) * 4;
        const float4 g = ldg_f4(dz + v * lddz + c);
        const int4 a = __ldg(reinterpret_cast<const int4*>(arg + v * ldarg + c));
        float4 zz = make_float4(1.f, 1.f, 1.f, 1.f);
        if (z) zz = ldg_f4(z + v * ldz + c);
        if (a.x >= 0 && g.x != 0.f && zz.x > 0.f && c + 0 < feat) atomicAdd(dx + (int64_t)a.x * lddx + c + 0, g.x);
        if (a.y >= 0 && g.y != 0.f && zz.y > 0.f && c + 1 < feat) atomicAdd(dx + (int64_t)a.y * lddx + c + 1, g.y);
        if (a.z >= 0 && g.z != 0.f && zz.z > 0.f && c + 2 < feat) atomicAdd(dx + (int64_t)a.z * lddx + c + 2, g.z);
        if (a.w >= 0 && g.w != 0.f && zz.w > 0.f && c + 3 < feat) atomicAdd(dx + (int64_t)a.w * lddx + c + 3, g.w);
    }
}

// dx[0..rows) x [0..4*w4) = 0 with 16-byte stores: the clear before the scatter, as a kernel so that it stays in the
// launch chain (a memset node would break it) — one block per row step, threads over the row's float4 groups
__global__ void __launch_bounds__(128)
zero_rows_kernel(float* __restrict__ dx, int64_t ld, int64_t rows, int w4) {
    pdl_enter();
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x)
        for (int c = threadIdx.x; c < w4; c += 128) *reinterpret_cast<float4*>(dx + r * ld + 4 * c) = z;
}

__global__ void __launch_bounds__(256)
dropout_scale_kernel(float* __restrict__ g, int64_t rows, int feat, int64_t ld, float p, unsigned long long seed) {
    pdl_enter();
    const int64_t total = rows * feat;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / feat;
        const int c = (int)(i - r * feat);
        g[r * ld + c] *= dropout_keep_scale(seed, r, c, feat, p);
    }
}

struct SpmmArgs {
    const int32_t *indptr, *indices, *eids;
    const float* ew;
    const void* plan;
    const int64_t* counts;
    int64_t n_rows;
    const float* x;
    int64_t ldx;
    int64_t feat;
    const int32_t* argm;
    int64_t ldarg;
    const float* zfwd;
    int64_t ldzf;
    float* out;
    int32_t* arg_out;
    int64_t ldo;
    void* partial;
    size_t partial_bytes;
    SpmmEpilogue ep;
    const int64_t* range = nullptr;   // optional host[4]: item_begin, item_end, hub_begin, hub_end (row-range launch)
    SpmmChain chain = SpmmChain{nullptr, nullptr, 0, 1};   // source-slab pass (narrow rows only)
};

static inline int part_ld_of(int64_t feat) { return (int)((feat + 3) / 4 * 4); }

template <int MODE, int VEC, int NB>
static void launch_main(const SpmmArgs& a, const int32_t* item_ptr, const int32_t* slot_ptr, const int32_t* item_row,
                        float* pv, int32_t* pa, cudaStream_t st) {
    const int item_begin = a.range ? (int)a.range[0] : 0;
    const int n_items = a.range ? (int)a.range[1] : (int)a.counts[0];
    if (n_items <= item_begin) return;
    // PLAGNN_SPMM_WARPS (read per launch): work items per block, 8 / 4 / 2 (default).  A block keeps its SM slot until its
    // longest item is done, and the rows of a power-law graph differ by orders of magnitude.  Measured, 8 / 4 / 2 warps per
    // block: 1 M / 100 M graph, 256 columns 12.85 / 11.96 / 11.87 ms (weighted sum), 11.29 / 10.45 / 10.32 (max); 128 columns
    // 7.40 / 6.66 / 6.47; PPI shape (24 041 rows, F = 503, max) 0.1916 / 0.1899 / 0.1903 ms.
    const char* we = getenv("PLAGNN_SPMM_WARPS");
    static const bool lean_env = [] { const char* e = getenv("PLAGNN_SPMM_SUM_LEAN"); return e && e[0] == '1'; }();
    const int want = we ? atoi(we) : 2;
    const int bw = ((want == 4 || want == 2) && MODE != MODE_MATCH && !(MODE == MODE_SUM && lean_env)) ? want : SPMM_WARPS;
    dim3 grid((unsigned)ceil_div(n_items - item_begin, bw), (unsigned)ceil_div((a.feat + 3) / 4, 32 * VEC));
    // PLAGNN_SPMM_SUM_LEAN=1: the max reducer's loop shape for the sum reducer.  Off by default — measured on the 1 M / 100 M
    // graph it LOSES to the shuffle form: F = 256 15.3 vs 12.8 ms, F = 128 9.7 vs 7.4 ms (two broadcast loads per neighbour, id
    // and weight, issued by every lane, against two coalesced loads and two shuffles per 32 neighbours).
    static const bool lean = [] { const char* e = getenv("PLAGNN_SPMM_SUM_LEAN"); return e && e[0] == '1'; }();
    constexpr int M2 = MODE == MODE_MATCH ? MODE_SUM : MODE;      // (MATCH never takes the small blocks; keeps the table instantiable)
    auto kernel = (MODE == MODE_SUM && lean) ? spmm_kernel<MODE, VEC, NB, true>
                  : bw == 4 ? spmm_kernel<M2, VEC, NB, false, 4>
                  : bw == 2 ? spmm_kernel<M2, VEC, NB, false, 2>
                            : spmm_kernel<MODE, VEC, NB, false>;
    launch_pdl(kernel, grid, dim3(bw * 32), 0, st,
        a.indptr, a.indices, a.eids, a.ew, (const int32_t*)a.plan, item_ptr, slot_ptr, item_row, item_begin, n_items, a.x, a.ldx,
        (int)a.feat, a.argm, a.ldarg, a.zfwd, a.ldzf, a.out, a.arg_out, a.ldo, pv, pa, part_ld_of(a.feat), a.ep, 1.0f, -0.0f);
}

template <int MODE, int G>
static void launch_narrow(const SpmmArgs& a, const int32_t* item_ptr, const int32_t* slot_ptr, const int32_t* item_row,
                          float* pv, int32_t* pa, cudaStream_t st) {
    const int item_begin = a.range ? (int)a.range[0] : 0;
    const int n_items = a.range ? (int)a.range[1] : (int)a.counts[0];
    if (n_items <= item_begin) return;
    // PLAGNN_SPMM_NARROW_U=8: eight rows in flight per lane at 24 warps per SM (A/B against the default 4 at 32 warps)
    static const int deep = [] { const char* e = getenv("PLAGNN_SPMM_NARROW_U"); return e && atoi(e) == 8; }();
    // PLAGNN_SPMM_NARROW_WARPS (read per launch): warps = work items per block.  A block keeps its SM slot until its longest
    // item is done; the rows of a power-law graph differ by orders of magnitude, so smaller blocks waste fewer warp slots.
    const char* we = getenv("PLAGNN_SPMM_NARROW_WARPS");
    const int warps = we ? atoi(we) : 2;      // measured at 32 columns, 8 / 4 / 2 / 1 warps per block: 1.95 / 1.78 / 1.75 / 1.83 ms (sum)
    auto kernel = deep ? spmm_narrow_kernel<MODE, G, 8, 3>
                  : warps == 4 ? spmm_narrow_kernel<MODE, G, 4, 4, 4>
                  : warps == 2 ? spmm_narrow_kernel<MODE, G, 4, 4, 2>
                  : warps == 1 ? spmm_narrow_kernel<MODE, G, 4, 4, 1>
                               : spmm_narrow_kernel<MODE, G, 4, 4>;
    const int bw = deep ? SPMM_WARPS : (warps == 4 || warps == 2 || warps == 1) ? warps : SPMM_WARPS;
    dim3 grid((unsigned)ceil_div(n_items - item_begin, bw));
    launch_pdl(kernel, grid, dim3(bw * 32), 0, st,
        a.indptr, a.indices, a.eids, a.ew, (const int32_t*)a.plan, item_ptr, slot_ptr, item_row, item_begin, n_items, a.x, a.ldx,
        (int)a.feat, a.out, a.arg_out, a.ldo, pv, pa, part_ld_of(a.feat), a.ep, a.chain);
}

template <int MODE>
static int spmm_dispatch(const SpmmArgs& a, const char* name, cudaStream_t st) {
    if (!a.indptr || !a.indices || !a.plan || !a.counts || !a.x || !a.out) return fail(PLAGNN_ERR_ARG, name, "null pointer");
    if (a.n_rows <= 0 || a.feat <= 0 || a.ldx >= ((int64_t)1 << 30)) return fail(PLAGNN_ERR_ARG, name, "bad sizes");
    const int64_t f4 = (a.feat + 3) / 4 * 4;
    if (a.ldx < f4 || a.ldo < f4 || (a.ldx & 3) || (a.ldo & 3) || !aligned16(a.x) || !aligned16(a.out) ||
        (a.arg_out && !aligned16(a.arg_out)))
        return fail(PLAGNN_ERR_ALIGN, name, "feature matrices need 16-byte aligned rows (pitch % 4 == 0, pitch >= roundup4(feat))");
    if (MODE == MODE_MATCH && (!a.argm || !aligned16(a.argm) || a.ldarg < f4 || (a.ldarg & 3) ||
                               (a.zfwd && (!aligned16(a.zfwd) || a.ldzf < f4 || (a.ldzf & 3)))))
        return fail(PLAGNN_ERR_ALIGN, name, "arg / z matrices need 16-byte aligned rows");
    const int64_t n_items = a.counts[0], n_hubs = a.counts[1], n_slots = a.counts[2];
    if (n_items <= 0 || n_items >= ((int64_t)1 << 31)) return fail(PLAGNN_ERR_ARG, name, "bad plan counts");
    const int pld = part_ld_of(a.feat);
    const size_t need_val = align_up((size_t)n_slots * pld * sizeof(float), 256);
    const size_t need = n_slots ? need_val * (MODE == MODE_MAX ? 2 : 1) : 0;
    if (need && (!a.partial || a.partial_bytes < need)) return fail(PLAGNN_ERR_WORKSPACE, name, "partial buffer too small");
    float* pv = (float*)a.partial;
    int32_t* pa = (int32_t*)((char*)a.partial + need_val);
    const int32_t *item_ptr, *slot_ptr, *item_row, *hub_rows;
    plan_pointers(a.plan, a.n_rows, &item_ptr, &slot_ptr, &item_row, &hub_rows);

    const int64_t groups = (a.feat + 3) / 4;   // float4 column groups
    // NB neighbour rows in flight per lane.  Measured on the PPI-shaped graph (F = 503, chunk 512): NB = 2 -> 0.2115 ms,
    // NB = 4 -> 0.2174 ms (128 registers, lower occupancy); the deeper variant only wins for chunks >= 1024 and stays
    // opt-in (PLAGNN_SPMM_DEEP=1).
    static const int deep = [] { const char* e = getenv("PLAGNN_SPMM_DEEP"); return e ? atoi(e) : 0; }();
    // rows of <= 64 floats: several in-edges per load instruction (PLAGNN_SPMM_NARROW=0: the wide kernel, for A/B runs)
    static const bool narrow_ok = [] { const char* e = getenv("PLAGNN_SPMM_NARROW"); return !e || e[0] != '0'; }();
    const bool chained = a.chain.prev_val != nullptr || !a.chain.last;
    if (chained && (MODE == MODE_MATCH || groups > 16))
        return fail(PLAGNN_ERR_UNSUPPORTED, name, "source-slab passes are implemented for rows of at most 64 columns");
    if (MODE != MODE_MATCH && (narrow_ok || chained) && groups <= 16) {
        constexpr int NM = MODE == MODE_MATCH ? MODE_SUM : MODE;     // (MATCH never gets here; keeps the template instantiable)
        if (groups <= 8) launch_narrow<NM, 8>(a, item_ptr, slot_ptr, item_row, pv, pa, st);
        else launch_narrow<NM, 16>(a, item_ptr, slot_ptr, item_row, pv, pa, st);
    } else if (MODE == MODE_MATCH || !deep) {
        if (groups <= 32) launch_main<MODE, 1, 8>(a, item_ptr, slot_ptr, item_row, pv, pa, st);
        else if (groups <= 64) launch_main<MODE, 2, 4>(a, item_ptr, slot_ptr, item_row, pv, pa, st);
        else if (groups <= 96) launch_main<MODE, 3, 2>(a, item_ptr, slot_ptr, item_row, pv, pa, st);
        else launch_main<MODE, 4, 2>(a, item_ptr, slot_ptr, item_row, pv, pa, st);
    } else {
        if (groups <= 32) launch_main<MODE, 1, 16>(a, item_ptr, slot_ptr, item_row, pv, pa, st);
        else if (groups <= 64) launch_main<MODE, 2, 8>(a, item_ptr, slot_ptr, item_row, pv, pa, st);
        else if (groups <= 96) launch_main<MODE, 3, 4>(a, item_ptr, slot_ptr, item_row, pv, pa, st);
        else launch_main<MODE, 4, 4>(a, item_ptr, slot_ptr, item_row, pv, pa, st);
    }
    const int64_t hub_begin = a.range ? a.range[2] : 0, hub_end = a.range ? a.range[3] : n_hubs;
    if (a.range && (a.range[0] < 0 || a.range[1] > n_items || a.range[2] < 0 || a.range[3] > n_hubs))
        return fail(PLAGNN_ERR_ARG, name, "row range outside the plan");
    if (hub_end > hub_begin) {
        dim3 grid((unsigned)ceil_div(hub_end - hub_begin, SPMM_WARPS), (unsigned)ceil_div(a.feat, 128));
        launch_pdl(spmm_combine_kernel<MODE>, grid, dim3(SPMM_WARPS * 32), 0, st, item_ptr, slot_ptr, hub_rows, (int)hub_begin, (int)hub_end, (int)a.feat,
                                                                     pv, pa, pld, a.out, a.arg_out, a.ldo, a.ep, a.chain);
    }
    return check_launch(name, hub_end > hub_begin ? 2 : 1);
}

}  // namespace plagnn

using namespace plagnn;

extern "C" {

size_t plagnn_spmm_partial_bytes(int64_t partial_slots, int64_t feat, int reduce) {
    if (partial_slots <= 0) return 0;
    const size_t one = align_up((size_t)partial_slots * part_ld_of(feat) * sizeof(float), 256);
    return one * (reduce == PLAGNN_REDUCE_MAX ? 2 : 1);
}

int plagnn_spmm_max_fwd(const int32_t* indptr, const int32_t* indices, const void* plan, const int64_t* plan_counts,
                        int64_t num_rows, const float* x, int64_t ldx, int64_t feat, float* out, int32_t* arg,
                        int64_t ldo, void* partial, size_t partial_bytes, plagnn_stream_t stream) {
    if (!arg) return fail(PLAGNN_ERR_ARG, "spmm_max_fwd", "arg output is required");
    ProfileScope prof("spmm_max_fwd", feat, num_rows, 0, stream);
    SpmmArgs a{indptr, indices, nullptr, nullptr, plan, plan_counts, num_rows, x, ldx, feat, nullptr, 0, nullptr, 0,
               out, arg, ldo, partial, partial_bytes, SpmmEpilogue{nullptr, nullptr, 0, 0.f, 0.f, 0ull}};
    return spmm_dispatch<MODE_MAX>(a, "spmm_max_fwd", (cudaStream_t)stream);
}

int plagnn_spmm_max_bwd(const float* dz, int64_t lddz, const int32_t* arg, int64_t ldarg, const float* z, int64_t ldz,
                        int64_t num_rows, int64_t feat, float* dx, int64_t n_src, int64_t lddx, plagnn_stream_t stream) {
    cudaStream_t st = (cudaStream_t)stream;
    ProfileScope prof("spmm_max_bwd", feat, num_rows, 0, stream);
    if (!dz || !arg || !dx || num_rows <= 0 || feat <= 0 || n_src <= 0) return fail(PLAGNN_ERR_ARG, "spmm_max_bwd", "bad arguments");
    const int64_t f4 = (feat + 3) / 4 * 4;
    if (lddz < f4 || (lddz & 3) || ldarg < f4 || (ldarg & 3) || lddx < feat || !aligned16(dz) || !aligned16(arg) ||
        (z && (!aligned16(z) || ldz < f4 || (ldz & 3))))
        return fail(PLAGNN_ERR_ALIGN, "spmm_max_bwd", "dz/arg/z need 16-byte aligned rows");
    int launched = 1;
    if (lddx >= f4 && (lddx & 3) == 0 && aligned16(dx)) {
        const int64_t zgrid = n_src < (int64_t)sm_count() * 16 ? n_src : (int64_t)sm_count() * 16;
        launch_pdl(zero_rows_kernel, dim3((unsigned)zgrid), dim3(128), 0, st, dx, lddx, n_src, (int)(f4 / 4));
        launched = 2;
    } else {
        PLAGNN_CUDA_TRY(cudaMemset2DAsync(dx, lddx * sizeof(float), 0, (size_t)feat * sizeof(float), n_src, st));
    }
    const int64_t total = num_rows * (f4 / 4);
    const int grid = (int)(ceil_div(total, 256) < (int64_t)sm_count() * 16 ? ceil_div(total, 256) : (int64_t)sm_count() * 16);
    launch_pdl(spmm_max_scatter_kernel, dim3(grid), dim3(256), 0, st, dz, lddz, arg, ldarg, z, ldz, num_rows, (int)feat, dx, lddx);
    return check_launch("spmm_max_bwd", launched);
}

int plagnn_spmm_max_bwd_gather(const int32_t* out_indptr, const int32_t* out_indices, const void* out_plan,
                               const int64_t* out_plan_counts, int64_t n_src, const float* dz, int64_t lddz,
                               const int32_t* arg, int64_t ldarg, const float* z, int64_t ldz, int64_t feat, float* dx,
                               int64_t lddx, void* partial, size_t partial_bytes, plagnn_stream_t stream) {
    ProfileScope prof("spmm_max_bwd_gather", feat, n_src, 0, stream);
    SpmmArgs a{out_indptr, out_indices, nullptr, nullptr, out_plan, out_plan_counts, n_src, dz, lddz, feat, arg, ldarg,
               z, ldz, dx, nullptr, lddx, partial, partial_bytes, SpmmEpilogue{nullptr, nullptr, 0, 0.f, 0.f, 0ull}};
    return spmm_dispatch<MODE_MATCH>(a, "spmm_max_bwd_gather", (cudaStream_t)stream);
}

int plagnn_spmm_sum(const int32_t* indptr, const int32_t* indices, const int32_t* eids, const void* plan,
                    const int64_t* plan_counts, int64_t num_rows, const float* w, const float* scale, const float* x,
                    int64_t ldx, int64_t feat, const float* bias, int act, float slope, float dropout_p,
                    uint64_t dropout_seed, float* out, int64_t ldo, void* partial, size_t partial_bytes,
                    plagnn_stream_t stream) {
    ProfileScope prof("spmm_sum", feat, num_rows, w ? 1 : 0, stream);
    if (act < PLAGNN_ACT_NONE || act > PLAGNN_ACT_SIGMOID) return fail(PLAGNN_ERR_ARG, "spmm_sum", "unknown activation");
    if (dropout_p < 0.f || dropout_p >= 1.f) return fail(PLAGNN_ERR_ARG, "spmm_sum", "dropout_p must be in [0,1)");
    SpmmArgs a{indptr, indices, eids, w, plan, plan_counts, num_rows, x, ldx, feat, nullptr, 0, nullptr, 0, out, nullptr, ldo,
               partial, partial_bytes, SpmmEpilogue{scale, bias, act, slope, dropout_p, (unsigned long long)dropout_seed}};
    return spmm_dispatch<MODE_SUM>(a, "spmm_sum", (cudaStream_t)stream);
}

int plagnn_spmm_sum_slab(const int32_t* indptr, const int32_t* indices, const int32_t* eids, const void* plan,
                         const int64_t* plan_counts, int64_t num_rows, const float* w, const float* scale, const float* x,
                         int64_t ldx, int64_t feat, const float* bias, int act, float slope, const float* prev, int64_t ldprev,
                         int last, float* out, int64_t ldo, void* partial, size_t partial_bytes, plagnn_stream_t stream) {
    ProfileScope prof("spmm_sum", feat, num_rows, w ? 3 : 2, stream);
    if (act < PLAGNN_ACT_NONE || act > PLAGNN_ACT_SIGMOID) return fail(PLAGNN_ERR_ARG, "spmm_sum_slab", "unknown activation");
    if (prev && (ldprev < (feat + 3) / 4 * 4 || (ldprev & 3) || !aligned16(prev)))
        return fail(PLAGNN_ERR_ALIGN, "spmm_sum_slab", "prev needs 16-byte aligned rows");
    SpmmArgs a{indptr, indices, eids, w, plan, plan_counts, num_rows, x, ldx, feat, nullptr, 0, nullptr, 0, out, nullptr, ldo,
               partial, partial_bytes, SpmmEpilogue{scale, bias, act, slope, 0.f, 0ull}};
    a.chain = SpmmChain{prev, nullptr, ldprev, last ? 1 : 0};
    return spmm_dispatch<MODE_SUM>(a, "spmm_sum_slab", (cudaStream_t)stream);
}

int plagnn_spmm_max_slab(const int32_t* indptr, const int32_t* indices, const void* plan, const int64_t* plan_counts,
                         int64_t num_rows, const float* x, int64_t ldx, int64_t feat, const float* prev_val,
                         const int32_t* prev_arg, int64_t ldprev, int last, float* out, int32_t* arg, int64_t ldo, void* partial,
                         size_t partial_bytes, plagnn_stream_t stream) {
    if (!arg) return fail(PLAGNN_ERR_ARG, "spmm_max_slab", "arg output is required");
    if ((prev_val == nullptr) != (prev_arg == nullptr)) return fail(PLAGNN_ERR_ARG, "spmm_max_slab", "prev_val and prev_arg go together");
    if (prev_val && (ldprev < (feat + 3) / 4 * 4 || (ldprev & 3) || !aligned16(prev_val) || !aligned16(prev_arg)))
        return fail(PLAGNN_ERR_ALIGN, "spmm_max_slab", "prev needs 16-byte aligned rows");
    ProfileScope prof("spmm_max_fwd", feat, num_rows, 2, stream);
    SpmmArgs a{indptr, indices, nullptr, nullptr, plan, plan_counts, num_rows, x, ldx, feat, nullptr, 0, nullptr, 0,
               out, arg, ldo, partial, partial_bytes, SpmmEpilogue{nullptr, nullptr, 0, 0.f, 0.f, 0ull}};
    a.chain = SpmmChain{prev_val, prev_arg, ldprev, last ? 1 : 0};
    return spmm_dispatch<MODE_MAX>(a, "spmm_max_slab", (cudaStream_t)stream);
}

int plagnn_spmm_plan_range(const void* plan, int64_t num_rows, int64_t row_begin, int64_t row_end, int64_t* host_range,
                           plagnn_stream_t stream) {
    if (!plan || !host_range || row_begin < 0 || row_end < row_begin || row_end > num_rows)
        return fail(PLAGNN_ERR_ARG, "spmm_plan_range", "bad arguments");
    const int32_t *item_ptr, *slot_ptr, *item_row, *hub_rows;
    plan_pointers(plan, num_rows, &item_ptr, &slot_ptr, &item_row, &hub_rows);
    const int32_t* hub_ptr = slot_ptr + (slot_ptr - item_ptr);     // item_ptr | slot_ptr | hub_ptr are equally spaced
    int32_t v[4];
    cudaStream_t st = (cudaStream_t)stream;
    PLAGNN_CUDA_TRY(cudaMemcpyAsync(&v[0], item_ptr + row_begin, 4, cudaMemcpyDeviceToHost, st));
    PLAGNN_CUDA_TRY(cudaMemcpyAsync(&v[1], item_ptr + row_end, 4, cudaMemcpyDeviceToHost, st));
    PLAGNN_CUDA_TRY(cudaMemcpyAsync(&v[2], hub_ptr + row_begin, 4, cudaMemcpyDeviceToHost, st));
    PLAGNN_CUDA_TRY(cudaMemcpyAsync(&v[3], hub_ptr + row_end, 4, cudaMemcpyDeviceToHost, st));
    PLAGNN_CUDA_TRY(cudaStreamSynchronize(st));
    for (int i = 0; i < 4; ++i) host_range[i] = v[i];
    return PLAGNN_OK;
}

int plagnn_spmm_sum_rows(const int32_t* indptr, const int32_t* indices, const int32_t* eids, const void* plan,
                         const int64_t* plan_counts, const int64_t* row_range, int64_t num_rows, const float* w,
                         const float* scale, const float* x, int64_t ldx, int64_t feat, const float* bias, int act,
                         float slope, float* out, int64_t ldo, void* partial, size_t partial_bytes,
                         plagnn_stream_t stream) {
    if (!row_range) return fail(PLAGNN_ERR_ARG, "spmm_sum_rows", "row_range is required");
    if (act < PLAGNN_ACT_NONE || act > PLAGNN_ACT_SIGMOID) return fail(PLAGNN_ERR_ARG, "spmm_sum_rows", "unknown activation");
    ProfileScope prof("spmm_sum", feat, row_range[1] - row_range[0], w ? 1 : 0, stream);
    SpmmArgs a{indptr, indices, eids, w, plan, plan_counts, num_rows, x, ldx, feat, nullptr, 0, nullptr, 0, out, nullptr, ldo,
               partial, partial_bytes, SpmmEpilogue{scale, bias, act, slope, 0.f, 0ull}};
    a.range = row_range;
    return spmm_dispatch<MODE_SUM>(a, "spmm_sum_rows", (cudaStream_t)stream);
}

int plagnn_spmm_max_fwd_rows(const int32_t* indptr, const int32_t* indices, const void* plan, const int64_t* plan_counts,
                             const int64_t* row_range, int64_t num_rows, const float* x, int64_t ldx, int64_t feat, float* out,
                             int32_t* arg, int64_t ldo, void* partial, size_t partial_bytes, plagnn_stream_t stream) {
    if (!row_range) return fail(PLAGNN_ERR_ARG, "spmm_max_fwd_rows", "row_range is required");
    if (!arg) return fail(PLAGNN_ERR_ARG, "spmm_max_fwd_rows", "arg output is required");
    ProfileScope prof("spmm_max_fwd", feat, row_range[1] - row_range[0], 0, stream);
    SpmmArgs a{indptr, indices, nullptr, nullptr, plan, plan_counts, num_rows, x, ldx, feat, nullptr, 0, nullptr, 0,
               out, arg, ldo, partial, partial_bytes, SpmmEpilogue{nullptr, nullptr, 0, 0.f, 0.f, 0ull}};
    a.range = row_range;
    return spmm_dispatch<MODE_MAX>(a, "spmm_max_fwd_rows", (cudaStream_t)stream);
}

int plagnn_dropout_scale(float* grad, int64_t rows, int64_t feat, int64_t ld, float dropout_p, uint64_t dropout_seed,
                         plagnn_stream_t stream) {
    if (!grad || rows <= 0 || feat <= 0 || dropout_p < 0.f || dropout_p >= 1.f) return fail(PLAGNN_ERR_ARG, "dropout_scale", "bad arguments");
    if (dropout_p == 0.f) return PLAGNN_OK;
    const int64_t total = rows * feat;
    const int grid = (int)(ceil_div(total, 256) < (int64_t)sm_count() * 16 ? ceil_div(total, 256) : (int64_t)sm_count() * 16);
    launch_pdl(dropout_scale_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, grad, rows, (int)feat, ld, dropout_p,
                                                                 (unsigned long long)dropout_seed);
    return check_launch("dropout_scale");
}

}  // extern "C"
