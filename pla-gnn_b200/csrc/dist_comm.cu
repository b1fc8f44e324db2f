// Exchange steps of the partitioned aggregation (SURVEY.md 8b "nccl_* wrappers taking an ncclComm_t", 8e).
//
// The reference has no distributed code (single process, one `-d` device: code/main_normal.py:30,66); these entry points
// exist for BASELINE.json configs[3] only (1 M nodes / 100 M weighted edges across 2 / 4 / 8 GPUs):
//   * row partition     : all-gather of projected rows (forward), reduce-scatter of source gradients (backward);
//   * feature partition : all-to-all between "my rows x all columns" and "all rows x my columns" around the aggregation,
//                         with the pack / unpack kernels below doing the column-block transposition on the device;
//   * both              : one all-reduce of the weight gradients per step.
// NCCL is resolved at run time from the libnccl the process already has (the one PyTorch loaded), so libplagnn.so carries
// no link-time dependency on it and still loads on a machine without NCCL (the ABI tests run there).
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <cstdlib>
#include <cstring>

namespace plagnn {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

static NcclApi& nccl_api() {
    static NcclApi api = [] {
        NcclApi a;
        const char* env = getenv("PLAGNN_NCCL_LIB");
        // the copy already mapped into the process first (two NCCL versions in one process do not mix)
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h && env && env[0]) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return a;
        a.handle = h;
#define PLAGNN_NCCL_SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(h, name))
        PLAGNN_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
        PLAGNN_NCCL_SYM(CommInitRankConfig, "ncclCommInitRankConfig");
        PLAGNN_NCCL_SYM(CommInitRank, "ncclCommInitRank");
        PLAGNN_NCCL_SYM(CommDestroy, "ncclCommDestroy");
        PLAGNN_NCCL_SYM(AllGather, "ncclAllGather");
        PLAGNN_NCCL_SYM(ReduceScatter, "ncclReduceScatter");
        PLAGNN_NCCL_SYM(AllReduce, "ncclAllReduce");
        PLAGNN_NCCL_SYM(Send, "ncclSend");
        PLAGNN_NCCL_SYM(Recv, "ncclRecv");
        PLAGNN_NCCL_SYM(GroupStart, "ncclGroupStart");
        PLAGNN_NCCL_SYM(GroupEnd, "ncclGroupEnd");
        PLAGNN_NCCL_SYM(GetErrorString, "ncclGetErrorString");
#undef PLAGNN_NCCL_SYM
        a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.ReduceScatter && a.AllReduce && a.Send &&
               a.Recv && a.GroupStart && a.GroupEnd;
        return a;
    }();
    return api;
}

static int nccl_fail(const char* who, ncclResult_t r) {
    NcclApi& a = nccl_api();
    set_error("%s: NCCL error %d (%s)", who, (int)r, a.GetErrorString ? a.GetErrorString(r) : "?");
    return PLAGNN_ERR_CUDA;
}

#define PLAGNN_NCCL_TRY(who, expr)                        \
    do {                                                  \
        ncclResult_t _r = (expr);                         \
        if (_r != ncclSuccess) return nccl_fail(who, _r); \
    } while (0)

static int need_nccl(const char* who) {
    if (!nccl_api().ok) return fail(PLAGNN_ERR_UNSUPPORTED, who, "libnccl.so.2 could not be resolved (set PLAGNN_NCCL_LIB)");
    return PLAGNN_OK;
}

// x[rows x feat] (row pitch ldx) -> out[world][rows][fc], fc = feat / world: block q holds columns [q*fc, (q+1)*fc).
// One float4 per thread, consecutive threads walk a row of the SOURCE (coalesced reads; writes are fc*4-byte runs).
__global__ void __launch_bounds__(256) cols_pack_kernel(const float* __restrict__ x, int64_t ldx, int64_t rows, int feat, int fc,
                                                        float* __restrict__ out) {
    pdl_enter();
    const int f4 = feat >> 2, fc4 = fc >> 2;
    const int64_t total = rows * f4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / f4;
        const int c4 = (int)(i - r * f4);
        const int q = c4 / fc4, j4 = c4 - q * fc4;
        const float4 v = ldg_f4(x + r * ldx + 4 * c4);
        *reinterpret_cast<float4*>(out + ((int64_t)q * rows + r) * fc + 4 * j4) = v;
    }
}

// in[world][rows][fc] -> x[rows x feat] (row pitch ldx): consecutive threads walk a row of the DESTINATION
__global__ void __launch_bounds__(256) cols_unpack_kernel(const float* __restrict__ in, int64_t rows, int feat, int fc,
                                                          float* __restrict__ x, int64_t ldx) {
    pdl_enter();
    const int f4 = feat >> 2, fc4 = fc >> 2;
    const int64_t total = rows * f4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / f4;
        const int c4 = (int)(i - r * f4);
        const int q = c4 / fc4, j4 = c4 - q * fc4;
        const float4 v = ldg_f4(in + ((int64_t)q * rows + r) * fc + 4 * j4);
        *reinterpret_cast<float4*>(x + r * ldx + 4 * c4) = v;
    }
}

static int check_cols(const char* who, const void* a, const void* b, int64_t ld, int64_t rows, int64_t feat, int world) {
    if (!a || !b || rows <= 0 || feat <= 0 || world <= 0) return fail(PLAGNN_ERR_ARG, who, "bad arguments");
    if (feat % (4 * (int64_t)world)) return fail(PLAGNN_ERR_ARG, who, "feat must be a multiple of 4 * world");
    if (ld < feat || (ld & 3) || !aligned16(a) || !aligned16(b)) return fail(PLAGNN_ERR_ALIGN, who, "16-byte aligned rows needed");
    return PLAGNN_OK;
}

}  // namespace plagnn

using namespace plagnn;

extern "C" {

int plagnn_nccl_available(void) { return nccl_api().ok ? 1 : 0; }

int plagnn_nccl_get_unique_id(void* id_out) {
    if (!id_out) return fail(PLAGNN_ERR_ARG, "nccl_get_unique_id", "null pointer");
    int rc = need_nccl("nccl_get_unique_id");
    if (rc) return rc;
    ncclUniqueId id;
    PLAGNN_NCCL_TRY("nccl_get_unique_id", nccl_api().GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == PLAGNN_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(id_out, &id, sizeof(id));
    return PLAGNN_OK;
}

int plagnn_nccl_comm_init(const void* id, int rank, int world, int max_ctas, plagnn_nccl_comm_t* comm_out) {
    if (!id || !comm_out || world <= 0 || rank < 0 || rank >= world) return fail(PLAGNN_ERR_ARG, "nccl_comm_init", "bad arguments");
    int rc = need_nccl("nccl_comm_init");
    if (rc) return rc;
    NcclApi& a = nccl_api();
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm = nullptr;
    if (max_ctas > 0 && a.CommInitRankConfig) {
        // fewer NCCL CTAs: the exchange runs beside the aggregation, which needs the SMs and the HBM bandwidth more
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.maxCTAs = max_ctas;
        cfg.minCTAs = 1;
        PLAGNN_NCCL_TRY("nccl_comm_init", a.CommInitRankConfig(&comm, world, uid, rank, &cfg));
    } else {
        PLAGNN_NCCL_TRY("nccl_comm_init", a.CommInitRank(&comm, world, uid, rank));
    }
    *comm_out = (plagnn_nccl_comm_t)comm;
    return PLAGNN_OK;
}

int plagnn_nccl_comm_destroy(plagnn_nccl_comm_t comm) {
    if (!comm) return PLAGNN_OK;
    int rc = need_nccl("nccl_comm_destroy");
    if (rc) return rc;
    PLAGNN_NCCL_TRY("nccl_comm_destroy", nccl_api().CommDestroy((ncclComm_t)comm));
    return PLAGNN_OK;
}

int plagnn_nccl_allgather_rows(const float* send, float* recv, int64_t rows, int64_t pitch, plagnn_nccl_comm_t comm,
                               plagnn_stream_t stream) {
    if (!send || !recv || !comm || rows <= 0 || pitch <= 0) return fail(PLAGNN_ERR_ARG, "nccl_allgather_rows", "bad arguments");
    int rc = need_nccl("nccl_allgather_rows");
    if (rc) return rc;
    ProfileScope prof("nccl_allgather", rows, pitch, 0, stream);
    PLAGNN_NCCL_TRY("nccl_allgather_rows", nccl_api().AllGather(send, recv, (size_t)(rows * pitch), ncclFloat, (ncclComm_t)comm,
                                                                (cudaStream_t)stream));
    return PLAGNN_OK;
}

int plagnn_nccl_reducescatter_rows(const float* send, float* recv, int64_t rows, int64_t pitch, plagnn_nccl_comm_t comm,
                                   plagnn_stream_t stream) {
    if (!send || !recv || !comm || rows <= 0 || pitch <= 0) return fail(PLAGNN_ERR_ARG, "nccl_reducescatter_rows", "bad arguments");
    int rc = need_nccl("nccl_reducescatter_rows");
    if (rc) return rc;
    ProfileScope prof("nccl_reducescatter", rows, pitch, 0, stream);
    PLAGNN_NCCL_TRY("nccl_reducescatter_rows", nccl_api().ReduceScatter(send, recv, (size_t)(rows * pitch), ncclFloat, ncclSum,
                                                                        (ncclComm_t)comm, (cudaStream_t)stream));
    return PLAGNN_OK;
}

int plagnn_nccl_allreduce(float* buf, int64_t count, plagnn_nccl_comm_t comm, plagnn_stream_t stream) {
    if (!buf || !comm || count <= 0) return fail(PLAGNN_ERR_ARG, "nccl_allreduce", "bad arguments");
    int rc = need_nccl("nccl_allreduce");
    if (rc) return rc;
    ProfileScope prof("nccl_allreduce", count, 0, 0, stream);
    PLAGNN_NCCL_TRY("nccl_allreduce", nccl_api().AllReduce(buf, buf, (size_t)count, ncclFloat, ncclSum, (ncclComm_t)comm,
                                                           (cudaStream_t)stream));
    return PLAGNN_OK;
}

int plagnn_nccl_alltoall_blocks(const float* send, float* recv, int64_t block_elems, int world, plagnn_nccl_comm_t comm,
                                plagnn_stream_t stream) {
    if (!send || !recv || !comm || block_elems <= 0 || world <= 0) return fail(PLAGNN_ERR_ARG, "nccl_alltoall_blocks", "bad arguments");
    int rc = need_nccl("nccl_alltoall_blocks");
    if (rc) return rc;
    NcclApi& a = nccl_api();
    ProfileScope prof("nccl_alltoall", block_elems, world, 0, stream);
    PLAGNN_NCCL_TRY("nccl_alltoall_blocks", a.GroupStart());
    for (int q = 0; q < world; ++q) {
        PLAGNN_NCCL_TRY("nccl_alltoall_blocks", a.Send(send + (int64_t)q * block_elems, (size_t)block_elems, ncclFloat, q,
                                                        (ncclComm_t)comm, (cudaStream_t)stream));
        PLAGNN_NCCL_TRY("nccl_alltoall_blocks", a.Recv(recv + (int64_t)q * block_elems, (size_t)block_elems, ncclFloat, q,
                                                        (ncclComm_t)comm, (cudaStream_t)stream));
    }
    PLAGNN_NCCL_TRY("nccl_alltoall_blocks", a.GroupEnd());
    return PLAGNN_OK;
}

int plagnn_cols_pack(const float* x, int64_t ldx, int64_t rows, int64_t feat, int world, float* out, plagnn_stream_t stream) {
    int rc = check_cols("cols_pack", x, out, ldx, rows, feat, world);
    if (rc) return rc;
    ProfileScope prof("cols_pack", rows, feat, world, stream);
    const int64_t total = rows * (feat / 4);
    const int grid = (int)(ceil_div(total, 256) < (int64_t)sm_count() * 16 ? ceil_div(total, 256) : (int64_t)sm_count() * 16);
    launch_pdl(cols_pack_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, x, ldx, rows, (int)feat, (int)(feat / world), out);
    return check_launch("cols_pack");
}

int plagnn_cols_unpack(const float* in, int64_t rows, int64_t feat, int world, float* x, int64_t ldx, plagnn_stream_t stream) {
    int rc = check_cols("cols_unpack", in, x, ldx, rows, feat, world);
    if (rc) return rc;
    ProfileScope prof("cols_unpack", rows, feat, world, stream);
    const int64_t total = rows * (feat / 4);
    const int grid = (int)(ceil_div(total, 256) < (int64_t)sm_count() * 16 ? ceil_div(total, 256) : (int64_t)sm_count() * 16);
    launch_pdl(cols_unpack_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, in, rows, (int)feat, (int)(feat / world), x, ldx);
    return check_launch("cols_unpack");
}

}  // extern "C"

// =====================================================================================================================
// Peer-memory exchange for the feature partition: the two all-to-all steps around an aggregation written as ONE kernel each
// that reads the local matrix and stores every column block straight into its owner's window over NVLink (no pack kernel, no
// staging buffer, no NCCL), followed by a system-scope flag per peer; the receiver's stream waits on those flags in a
// one-warp kernel.  Windows are device allocations of the library (they have to be exportable with cudaIpcGetMemHandle),
// opened by the peers at set-up.
// =====================================================================================================================
namespace plagnn {

constexpr int P2P_MAX_WORLD = 16;
constexpr size_t P2P_FLAG_BYTES = 4096;          // flags[world] (int64) + error word, ahead of the data

struct P2P {
    int rank, world;
    size_t bytes;
    char* local;                                  // flags | data
    char* peer[P2P_MAX_WORLD];                    // opened windows (peer[rank] == local)
    char** d_peer;                                // device copy of peer[]
    unsigned int* d_counter;                      // blocks of the running send kernel that have finished
};

__device__ __forceinline__ void st_release_sys(long long* p, long long v) {
    asm volatile("st.release.sys.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
    long long v;
    asm volatile("ld.acquire.sys.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// mode 0 (rows -> columns): src = x[rows x feat] (pitch lds); block q = columns [q fc, (q+1) fc) goes to peer q at
//                            data[off + ((rank * rows + r) * fc + j)]: the peer's window is the all-rows matrix [world * rows x fc].
// mode 1 (columns -> rows): src = x_col[world * rows x fc] (pitch lds); rows [q rows, (q+1) rows) go to peer q at
//                            data[off + (r * feat + rank * fc + j)]: the peer's window is its [rows x feat] matrix.
// A send may cover only the source rows [row_begin, row_end) (mode 0: a chunk of my rows, for every peer; mode 1: rows of the
// all-rows matrix, e.g. the block of one owner), so that the producer's next chunk is computed while this one crosses NVLink.
// publish: 0 = no flag (more parts follow on the same stream), 1 = flag every peer, 2 = flag the owner of row_begin only.
__global__ void __launch_bounds__(256)
p2p_send_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int feat, int fc, int mode, char* const* __restrict__ peer,
                size_t off_bytes, int rank, int world, long long seq, unsigned int* __restrict__ counter, int64_t row_begin,
                int64_t row_end, int publish) {
    pdl_enter();
    const int fc4 = fc >> 2;
    const int64_t total = (row_end - row_begin) * (mode == 0 ? (int64_t)world * fc4 : (int64_t)fc4);   // float4 elements of the part
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int q, j4;
        int64_t r;
        float4 v;
        float* dst;
        if (mode == 0) {
            const int f4 = world * fc4;
            r = row_begin + i / f4;
            const int c4 = (int)(i - (i / f4) * f4);
            q = c4 / fc4;
            j4 = c4 - q * fc4;
            v = ldg_f4(src + r * lds + 4 * c4);
            dst = reinterpret_cast<float*>(peer[q] + P2P_FLAG_BYTES + off_bytes) + ((int64_t)rank * rows + r) * fc + 4 * j4;
        } else {
            const int64_t gr = row_begin + i / fc4;
            j4 = (int)(i - (i / fc4) * fc4);
            q = (int)(gr / rows);
            r = gr - (int64_t)q * rows;
            v = ldg_f4(src + gr * lds + 4 * j4);
            dst = reinterpret_cast<float*>(peer[q] + P2P_FLAG_BYTES + off_bytes) + r * feat + (int64_t)rank * fc + 4 * j4;
        }
        *reinterpret_cast<float4*>(dst) = v;
    }
    // the last block to finish publishes the sequence number in every peer's flag array (after everybody's stores)
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(counter, 1u);
        if (done == gridDim.x - 1) {
            *counter = 0;
            __threadfence_system();
            if (publish == 1) {
                for (int q = 0; q < world; ++q) st_release_sys(reinterpret_cast<long long*>(peer[q]) + rank, seq);
            } else if (publish == 2) {
                st_release_sys(reinterpret_cast<long long*>(peer[(int)(row_begin / rows)]) + rank, seq);
            }
        }
    }
}

// one warp: lane p waits until peer p has published `seq`; gives up after ~4 s (error word set, later waits return at once)
__global__ void p2p_wait_kernel(char* local, int world, long long seq) {
    pdl_enter();
    long long* flags = reinterpret_cast<long long*>(local);
    long long* err = flags + P2P_MAX_WORLD;
    const int p = threadIdx.x;
    if (p < world && *reinterpret_cast<volatile long long*>(err) == 0) {
        const long long t0 = clock64();
        while (ld_acquire_sys(flags + p) < seq) {
            if (clock64() - t0 > 8000000000ll) {          // ~4 s: a peer is not coming
                *reinterpret_cast<volatile long long*>(err) = seq;
                break;
            }
        }
    }
    __syncwarp();
}

}  // namespace plagnn

extern "C" {

int plagnn_p2p_create(size_t bytes, int rank, int world, void* handle_out, plagnn_p2p_t* out) {
    if (!handle_out || !out || bytes == 0 || world <= 0 || world > P2P_MAX_WORLD || rank < 0 || rank >= world)
        return fail(PLAGNN_ERR_ARG, "p2p_create", "bad arguments");
    P2P* p = new P2P();
    p->rank = rank; p->world = world; p->bytes = bytes;
    p->local = nullptr; p->d_peer = nullptr; p->d_counter = nullptr;
    for (int i = 0; i < P2P_MAX_WORLD; ++i) p->peer[i] = nullptr;
    cudaError_t e = cudaMalloc(&p->local, bytes + P2P_FLAG_BYTES);
    if (e == cudaSuccess) e = cudaMemset(p->local, 0, P2P_FLAG_BYTES);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_peer, sizeof(char*) * P2P_MAX_WORLD);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_counter, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(p->d_counter, 0, sizeof(unsigned int));
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->local);
    if (e != cudaSuccess) {
        set_error("p2p_create: %s", cudaGetErrorString(e));
        cudaGetLastError();
        if (p->local) cudaFree(p->local);
        if (p->d_peer) cudaFree(p->d_peer);
        if (p->d_counter) cudaFree(p->d_counter);
        delete p;
        return PLAGNN_ERR_CUDA;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == PLAGNN_P2P_HANDLE_BYTES, "cudaIpcMemHandle_t size");
    memcpy(handle_out, &h, sizeof(h));
    p->peer[rank] = p->local;
    *out = (plagnn_p2p_t)p;
    return PLAGNN_OK;
}

int plagnn_p2p_attach(plagnn_p2p_t px, const void* all_handles) {
    P2P* p = (P2P*)px;
    if (!p || !all_handles) return fail(PLAGNN_ERR_ARG, "p2p_attach", "bad arguments");
    for (int q = 0; q < p->world; ++q) {
        if (q == p->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)all_handles + (size_t)q * PLAGNN_P2P_HANDLE_BYTES, sizeof(h));
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("p2p_attach: cudaIpcOpenMemHandle(peer %d): %s", q, cudaGetErrorString(e));
            cudaGetLastError();
            return PLAGNN_ERR_CUDA;
        }
        p->peer[q] = (char*)ptr;
    }
    PLAGNN_CUDA_TRY(cudaMemcpy(p->d_peer, p->peer, sizeof(char*) * P2P_MAX_WORLD, cudaMemcpyHostToDevice));
    return PLAGNN_OK;
}

void* plagnn_p2p_window(plagnn_p2p_t px) {
    P2P* p = (P2P*)px;
    return p ? (void*)(p->local + P2P_FLAG_BYTES) : nullptr;
}

long long plagnn_p2p_error(plagnn_p2p_t px) {
    P2P* p = (P2P*)px;
    if (!p) return -1;
    long long v = 0;
    if (cudaMemcpy(&v, p->local + sizeof(long long) * P2P_MAX_WORLD, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return v;
}

int plagnn_p2p_destroy(plagnn_p2p_t px) {
    P2P* p = (P2P*)px;
    if (!p) return PLAGNN_OK;
    cudaDeviceSynchronize();
    for (int q = 0; q < p->world; ++q)
        if (q != p->rank && p->peer[q]) cudaIpcCloseMemHandle(p->peer[q]);
    cudaFree(p->local);
    cudaFree(p->d_peer);
    cudaFree(p->d_counter);
    delete p;
    return PLAGNN_OK;
}

int plagnn_p2p_send_part(plagnn_p2p_t px, const float* src, int64_t lds, int64_t rows, int64_t feat, int mode, int64_t row_begin,
                         int64_t row_end, int publish, size_t dst_offset_bytes, long long seq, plagnn_stream_t stream);

int plagnn_p2p_send(plagnn_p2p_t px, const float* src, int64_t lds, int64_t rows, int64_t feat, int mode, size_t dst_offset_bytes,
                    long long seq, plagnn_stream_t stream) {
    P2P* p = (P2P*)px;
    if (!p) return fail(PLAGNN_ERR_ARG, "p2p_send", "bad arguments");
    return plagnn_p2p_send_part(px, src, lds, rows, feat, mode, 0, mode == 0 ? rows : rows * p->world, 1, dst_offset_bytes, seq, stream);
}

int plagnn_p2p_send_part(plagnn_p2p_t px, const float* src, int64_t lds, int64_t rows, int64_t feat, int mode, int64_t row_begin,
                         int64_t row_end, int publish, size_t dst_offset_bytes, long long seq, plagnn_stream_t stream) {
    P2P* p = (P2P*)px;
    if (!p || !src || rows <= 0 || feat <= 0 || (mode != 0 && mode != 1)) return fail(PLAGNN_ERR_ARG, "p2p_send", "bad arguments");
    const int64_t src_rows = mode == 0 ? rows : rows * p->world;
    if (row_begin < 0 || row_end <= row_begin || row_end > src_rows || publish < 0 || publish > 2 ||
        (publish == 2 && (mode != 1 || row_begin / rows != (row_end - 1) / rows)))
        return fail(PLAGNN_ERR_ARG, "p2p_send", "bad row range (publish = 2: mode 1, rows of one owner)");
    if (feat % (4 * (int64_t)p->world)) return fail(PLAGNN_ERR_ARG, "p2p_send", "feat must be a multiple of 4 * world");
    const int64_t fc = feat / p->world;
    if ((lds & 3) || lds < (mode == 0 ? feat : fc) || !aligned16(src) || (dst_offset_bytes & 15))
        return fail(PLAGNN_ERR_ALIGN, "p2p_send", "16-byte aligned rows needed");
    if (dst_offset_bytes + (size_t)rows * feat * sizeof(float) > p->bytes) return fail(PLAGNN_ERR_WORKSPACE, "p2p_send", "window too small");
    ProfileScope prof(mode == 0 ? "p2p_send_cols" : "p2p_send_rows", rows, feat, p->world, stream);
    const int64_t total = (row_end - row_begin) * (mode == 0 ? feat / 4 : fc / 4);
    const int grid = (int)(ceil_div(total, 256) < (int64_t)sm_count() * 8 ? ceil_div(total, 256) : (int64_t)sm_count() * 8);
    launch_pdl(p2p_send_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, src, lds, rows, (int)feat, (int)fc, mode,
               (char* const*)p->d_peer, dst_offset_bytes, p->rank, p->world, seq, p->d_counter, row_begin, row_end, publish);
    return check_launch("p2p_send");
}

int plagnn_p2p_wait(plagnn_p2p_t px, long long seq, plagnn_stream_t stream) {
    P2P* p = (P2P*)px;
    if (!p) return fail(PLAGNN_ERR_ARG, "p2p_wait", "bad arguments");
    ProfileScope prof("p2p_wait", seq, 0, 0, stream);
    launch_pdl(p2p_wait_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, p->local, p->world, seq);
    return check_launch("p2p_wait");
}

}  // extern "C"
