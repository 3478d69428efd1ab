// plf_comm.cu -- NCCL plumbing of the strip-partitioned solve and its C-ABI entry points (include/plfb200.h).
#include <dlfcn.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "mg.h"
#include "plf_comm.h"
#include "plf_model.h"
#include "plfb200.h"

// The few NCCL declarations this file needs, stated locally (NCCL 2.x ABI): the library is loaded with dlopen at run time,
// and with <nccl.h> the dev headers would be a hard build-time dependency of every single-GPU user of libplfb200.so.
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
typedef enum { ncclDouble = 8 } ncclDataType_t;   // ncclFloat64
}

namespace plf {

namespace {
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::string g_err;
void p2p_release(Comm &c);

int nccl_fail(const char *what, ncclResult_t r)
{
    g_err = std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error");
    return 1;
}
#define NC(call, what)                                   \
    do {                                                 \
        ncclResult_t r_ = (call);                        \
        if (r_ != ncclSuccess) return nccl_fail(what, r_); \
    } while (0)
}   // namespace

const char *comm_error() { return g_err.c_str(); }

int comm_load()
{
    if (g_nccl.lib) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_err = std::string("cannot load libnccl.so.2: ") + dlerror(); return 1; }
    NcclApi a;
    a.lib = h;
    bool ok = true;
    auto sym = [&](const char *name) { void *p = dlsym(h, name); if (!p) { ok = false; g_err = std::string("libnccl lacks ") + name; } return p; };
    a.GetUniqueId = (decltype(a.GetUniqueId))sym("ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))sym("ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))sym("ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))sym("ncclAllReduce");
    a.Send = (decltype(a.Send))sym("ncclSend");
    a.Recv = (decltype(a.Recv))sym("ncclRecv");
    a.Broadcast = (decltype(a.Broadcast))sym("ncclBroadcast");
    a.GroupStart = (decltype(a.GroupStart))sym("ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))sym("ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { dlclose(h); return 1; }
    g_nccl = a;
    return 0;
}

int comm_unique_id(unsigned char *id128)
{
    if (comm_load()) return 1;
    static_assert(sizeof(ncclUniqueId) == PLF_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id128, &id, sizeof id);
    return 0;
}

int comm_init(Comm &c, const unsigned char *id128, int rank, int world)
{
    if (comm_load()) return 1;
    comm_destroy(c);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t comm = nullptr;
    NC(g_nccl.CommInitRank(&comm, world, id, rank), "ncclCommInitRank");
    c.nccl = comm;
    c.rank = rank;
    c.world = world;
    c.strip_set = false;
    return 0;
}

void comm_destroy(Comm &c)
{
    p2p_release(c);
    if (c.nccl && g_nccl.CommDestroy) g_nccl.CommDestroy((ncclComm_t)c.nccl);
    c = Comm();
}


// ---- peer mailboxes ---------------------------------------------------------------------------------------------------------
// Layout of one rank's mailbox (device memory, mapped into every other rank's process through CUDA IPC):
//   [0, 16)   two arrival flags of the halo exchange (0: written by the left neighbour, 1: by the right neighbour)
//   [16, 24)  arrival counter of the all-reduce
//   [RED_OFF, HALO_OFF)  all-reduce slots, double[2 parities][64 ranks][64 values]
//   [HALO_OFF, ...)      halo slots, double[2 parities][2 sides][box_cols]
// An exchange is ONE kernel: push the payload into the peers' mailboxes (stores over NVLink), release-store the epoch into
// their flags, spin on the own flags, copy the own slots out.  Two parities are enough: a rank cannot run two exchanges ahead
// of a neighbour, because finishing exchange k + 1 needs that neighbour's push of k + 1, which comes after its copy-out of k.
// The exchanges of one communicator must not overlap in time (different streams of one rank): the callers order them with
// stream synchronisation, as they had to for NCCL.
namespace {
constexpr size_t RED_OFF = 64, RED_SLOT = 64, HALO_OFF = RED_OFF + 2 * 64 * RED_SLOT * sizeof(double) + 192;
static_assert(HALO_OFF % 256 == 0, "halo slots are 256-byte aligned");
constexpr long long SPIN_LIMIT = 120000000000ll;   // about a minute of SM clocks: a peer that never arrives is a lost rank

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void spin_until(const unsigned long long *flag, unsigned long long want)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < want)
        if (clock64() - t0 > SPIN_LIMIT) __trap();   // surfaces as a launch failure at the next synchronisation
}

__global__ void __launch_bounds__(1024) p2p_halo_kernel(double *vec, int ncol, long long sendL, long long recvL, long long sendR,
                                                        long long recvR, unsigned char *mine, unsigned char *left, unsigned char *right,
                                                        int box_cols, unsigned long long epoch)
{
    const int par = (int)(epoch & 1ull), tid = threadIdx.x;
    // the left boundary column lands in the left neighbour's "from the right" slot and vice versa
    if (left) {
        double *dst = reinterpret_cast<double *>(left + HALO_OFF) + (size_t)(par * 2 + 1) * box_cols;
        for (int i = tid; i < ncol; i += blockDim.x) dst[i] = vec[sendL + i];
    }
    if (right) {
        double *dst = reinterpret_cast<double *>(right + HALO_OFF) + (size_t)(par * 2 + 0) * box_cols;
        for (int i = tid; i < ncol; i += blockDim.x) dst[i] = vec[sendR + i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        if (left) st_release_sys(reinterpret_cast<unsigned long long *>(left) + 1, epoch);
        if (right) st_release_sys(reinterpret_cast<unsigned long long *>(right) + 0, epoch);
        if (left) spin_until(reinterpret_cast<unsigned long long *>(mine) + 0, epoch);
        if (right) spin_until(reinterpret_cast<unsigned long long *>(mine) + 1, epoch);
    }
    __syncthreads();
    if (left) {
        const double *src = reinterpret_cast<const double *>(mine + HALO_OFF) + (size_t)(par * 2 + 0) * box_cols;
        for (int i = tid; i < ncol; i += blockDim.x) vec[recvL + i] = __ldcg(src + i);
    }
    if (right) {
        const double *src = reinterpret_cast<const double *>(mine + HALO_OFF) + (size_t)(par * 2 + 1) * box_cols;
        for (int i = tid; i < ncol; i += blockDim.x) vec[recvR + i] = __ldcg(src + i);
    }
}

struct PeerBoxes { unsigned char *p[64]; };

// vals[0 .. n) combined over the ranks in rank order (the same bits on every rank): op 0 sum, 1 max, 2 min
__global__ void __launch_bounds__(512) p2p_allreduce_kernel(double *vals, int n, int op, PeerBoxes peers, int rank, int world,
                                                           unsigned long long epoch)
{
    const int par = (int)(epoch & 1ull), tid = threadIdx.x;
    for (int t = tid; t < n * world; t += blockDim.x) {
        const int r = t / n, i = t - r * n;
        reinterpret_cast<double *>(peers.p[r] + RED_OFF)[((size_t)par * 64 + rank) * RED_SLOT + i] = vals[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < world) atomicAdd_system(reinterpret_cast<unsigned long long *>(peers.p[tid] + 16), 1ull);
    if (tid == 0) spin_until(reinterpret_cast<unsigned long long *>(peers.p[rank] + 16), epoch * (unsigned long long)world);
    __syncthreads();
    for (int i = tid; i < n; i += blockDim.x) {
        const double *slot = reinterpret_cast<const double *>(peers.p[rank] + RED_OFF) + (size_t)par * 64 * RED_SLOT + i;
        double acc = __ldcg(slot);
        for (int r = 1; r < world; r++) {
            const double v = __ldcg(slot + (size_t)r * RED_SLOT);
            acc = op == 0 ? acc + v : (op == 1 ? fmax(acc, v) : fmin(acc, v));
        }
        vals[i] = acc;
    }
}

int cuda_fail(const char *what, cudaError_t e)
{
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
}

void p2p_release(Comm &c)
{
    for (int r = 0; r < 64; r++)
        if (c.peer_box[r] && c.peer_box[r] != c.box) cudaIpcCloseMemHandle(c.peer_box[r]);
    if (c.box) cudaFree(c.box);
    for (int r = 0; r < 64; r++) c.peer_box[r] = nullptr;
    c.box = nullptr;
    c.p2p = false;
    c.box_cols = 0;
    c.halo_epoch = c.red_epoch = 0;
}

int p2p_allreduce(Comm &c, double *dbuf, int count, int op, cudaStream_t s)
{
    PeerBoxes pb;
    for (int r = 0; r < 64; r++) pb.p[r] = c.peer_box[r];
    p2p_allreduce_kernel<<<1, 512, 0, s>>>(dbuf, count, op, pb, c.rank, c.world, ++c.red_epoch);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail("p2p all-reduce", e);
}

int p2p_halo(Comm &c, double *vec, size_t ncol, bool left, bool right, size_t sendL, size_t recvL, size_t sendR, size_t recvR,
             cudaStream_t s)
{
    p2p_halo_kernel<<<1, 1024, 0, s>>>(vec, (int)ncol, (long long)sendL, (long long)recvL, (long long)sendR, (long long)recvR, c.box,
                                       left ? c.peer_box[c.rank - 1] : nullptr, right ? c.peer_box[c.rank + 1] : nullptr, c.box_cols,
                                       ++c.halo_epoch);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : cuda_fail("p2p halo exchange", e);
}
}   // namespace

int comm_p2p_setup(Comm &c, int cols, cudaStream_t s)
{
    if (!c.active() || c.world < 2 || c.world > 64) return 0;
    const char *env = getenv("PLF_P2P");
    if (env && env[0] == '0') { p2p_release(c); return 0; }
    if (c.p2p && c.box_cols >= cols) return 0;   // every rank sees the same NY, hence takes the same branch
    double *d_h = nullptr;
    cudaError_t e = cudaMalloc(&d_h, (size_t)c.world * 64 + 8);
    if (e != cudaSuccess) return cuda_fail("comm_p2p_setup", e);
    if (c.box) {
        // A mesh with longer node columns than the mailbox holds (slots are sized for NY <= 8192 up front, so this is rare):
        // unmap the peers' mailboxes, wait until every rank has done the same (an exported allocation must not be freed while
        // another process still maps it), then free the own one.  The barrier runs on NCCL: p2p is off from here on.
        c.p2p = false;
        for (int r = 0; r < 64; r++) {
            if (c.peer_box[r] && c.peer_box[r] != c.box) cudaIpcCloseMemHandle(c.peer_box[r]);
            c.peer_box[r] = nullptr;
        }
        e = cudaMemsetAsync(d_h, 0, sizeof(double), s);
        if (e == cudaSuccess && comm_allreduce_op(c, d_h, 1, 0, s)) { cudaFree(d_h); return 1; }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { cudaFree(d_h); return cuda_fail("comm_p2p_setup", e); }
        p2p_release(c);
    }
    cols = std::max(cols, 2 * 8193);
    // every step below is attempted by all ranks, and the outcome is agreed on at the end: no rank may skip a collective
    bool ok = true;
    const size_t bytes = HALO_OFF + 4 * (size_t)cols * sizeof(double);
    cudaIpcMemHandle_t mine_h;
    memset(&mine_h, 0, sizeof mine_h);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    ok = ok && cudaMalloc(&c.box, bytes) == cudaSuccess;
    ok = ok && cudaMemset(c.box, 0, bytes) == cudaSuccess;
    ok = ok && cudaIpcGetMemHandle(&mine_h, c.box) == cudaSuccess;
    std::string all((size_t)c.world * 64, '\0');
    int64_t goff[64], cnt[64];
    for (int r = 0; r < c.world; r++) { goff[r] = 8 * r; cnt[r] = 8; }
    e = cudaMemcpyAsync(d_h + 8 * c.rank, &mine_h, 64, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && comm_gather_ranges(c, d_h + 8 * c.rank, d_h, goff, cnt, s)) { cudaFree(d_h); return 1; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&all[0], d_h, all.size(), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { cudaFree(d_h); return cuda_fail("comm_p2p_setup", e); }
    for (int r = 0; r < c.world && ok; r++) {
        if (r == c.rank) { c.peer_box[r] = c.box; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, &all[(size_t)r * 64], 64);
        void *ptr = nullptr;
        ok = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        c.peer_box[r] = (unsigned char *)ptr;
    }
    cudaGetLastError();   // a failed IPC call is not an error of the library: the exchanges stay on NCCL
    // agreement (this all-reduce is also the barrier behind every rank's cudaMemset)
    double flag = ok ? 1. : 0.;
    e = cudaMemcpyAsync(d_h + 8 * c.world, &flag, sizeof flag, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && comm_allreduce_op(c, d_h + 8 * c.world, 1, 2, s)) { cudaFree(d_h); return 1; }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, d_h + 8 * c.world, sizeof flag, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_h);
    if (e != cudaSuccess) return cuda_fail("comm_p2p_setup", e);
    if (flag > 0.5) { c.p2p = true; c.box_cols = cols; }
    else p2p_release(c);
    return 0;
}

int comm_allreduce_sum(Comm &c, double *dbuf, int count, cudaStream_t s)
{
    if (c.p2p && count <= (int)RED_SLOT) return p2p_allreduce(c, dbuf, count, 0, s);
    NC(g_nccl.AllReduce(dbuf, dbuf, (size_t)count, ncclDouble, ncclSum, (ncclComm_t)c.nccl, s), "ncclAllReduce");
    return 0;
}

int comm_allreduce_op(Comm &c, double *dbuf, int count, int op, cudaStream_t s)
{
    if (c.p2p && count <= (int)RED_SLOT) return p2p_allreduce(c, dbuf, count, op, s);
    const ncclRedOp_t o = op == 1 ? ncclMax : (op == 2 ? ncclMin : ncclSum);
    NC(g_nccl.AllReduce(dbuf, dbuf, (size_t)count, ncclDouble, o, (ncclComm_t)c.nccl, s), "ncclAllReduce");
    return 0;
}

int comm_halo_exchange_cols(Comm &c, double *vec, int NY, int a, int b, int NX, cudaStream_t s)
{
    const bool left = a > 0 && c.rank > 0, right = b <= NX && c.rank + 1 < c.world;
    if (!left && !right) return 0;
    const size_t ncol = 2 * (size_t)(NY + 1);
    if (c.p2p && (int)ncol <= c.box_cols)
        return p2p_halo(c, vec, ncol, left, right, ncol * (size_t)a, ncol * (size_t)(a - 1), ncol * (size_t)(b - 1), ncol * (size_t)b, s);
    ncclComm_t comm = (ncclComm_t)c.nccl;
    NC(g_nccl.GroupStart(), "ncclGroupStart");
    if (left) {
        NC(g_nccl.Send(vec + ncol * (size_t)a, ncol, ncclDouble, c.rank - 1, comm, s), "ncclSend");
        NC(g_nccl.Recv(vec + ncol * (size_t)(a - 1), ncol, ncclDouble, c.rank - 1, comm, s), "ncclRecv");
    }
    if (right) {
        NC(g_nccl.Send(vec + ncol * (size_t)(b - 1), ncol, ncclDouble, c.rank + 1, comm, s), "ncclSend");
        NC(g_nccl.Recv(vec + ncol * (size_t)b, ncol, ncclDouble, c.rank + 1, comm, s), "ncclRecv");
    }
    NC(g_nccl.GroupEnd(), "ncclGroupEnd");
    return 0;
}

// Every rank contributes one contiguous range of doubles to the same global array on all ranks: rank r's range is
// [goff[r], goff[r] + cnt[r]) of `global`; its own data sits at `mine`.  One broadcast per rank, grouped.
int comm_gather_ranges(Comm &c, const double *mine, double *global, const int64_t *goff, const int64_t *cnt, cudaStream_t s)
{
    ncclComm_t comm = (ncclComm_t)c.nccl;
    NC(g_nccl.GroupStart(), "ncclGroupStart");
    for (int r = 0; r < c.world; r++)
        NC(g_nccl.Broadcast(r == c.rank ? (const void *)mine : (const void *)(global + goff[r]), global + goff[r], (size_t)cnt[r],
                            ncclDouble, r, comm, s), "ncclBroadcast");
    NC(g_nccl.GroupEnd(), "ncclGroupEnd");
    return 0;
}

int comm_halo_exchange(Comm &c, double *vec, int NX, int NY, cudaStream_t s)
{
    const bool left = c.ghost_left() && c.rank > 0, right = c.ghost_right(NX) && c.rank + 1 < c.world;
    if (!left && !right) return 0;
    const size_t ncol = 2 * (size_t)(NY + 1);   // doubles per node column
    if (c.p2p && (int)ncol <= c.box_cols)
        return p2p_halo(c, vec, ncol, left, right, ncol * (size_t)c.own_first, ncol * (size_t)(c.own_first - 1),
                        ncol * (size_t)(c.own_end - 1), ncol * (size_t)c.own_end, s);
    ncclComm_t comm = (ncclComm_t)c.nccl;
    NC(g_nccl.GroupStart(), "ncclGroupStart");
    if (left) {
        NC(g_nccl.Send(vec + ncol * (size_t)c.own_first, ncol, ncclDouble, c.rank - 1, comm, s), "ncclSend");
        NC(g_nccl.Recv(vec + ncol * (size_t)(c.own_first - 1), ncol, ncclDouble, c.rank - 1, comm, s), "ncclRecv");
    }
    if (right) {
        NC(g_nccl.Send(vec + ncol * (size_t)(c.own_end - 1), ncol, ncclDouble, c.rank + 1, comm, s), "ncclSend");
        NC(g_nccl.Recv(vec + ncol * (size_t)c.own_end, ncol, ncclDouble, c.rank + 1, comm, s), "ncclRecv");
    }
    NC(g_nccl.GroupEnd(), "ncclGroupEnd");
    return 0;
}

}   // namespace plf

using namespace plf;

#define CKC(call)                                                                    \
    do {                                                                             \
        cudaError_t e_ = (call);                                                     \
        if (e_ != cudaSuccess) return fail(ctx, PLF_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" int plf_comm_unique_id(unsigned char *id)
{
    if (!id) return PLF_EINVAL;
    return comm_unique_id(id) ? PLF_EUNSUPPORTED : PLF_OK;
}

extern "C" int plf_comm_init(plf_ctx *ctx, const unsigned char *id, int rank, int world)
{
    if (!ctx) return PLF_EINVAL;
    if (!id || world < 1 || rank < 0 || rank >= world) return fail(ctx, PLF_EINVAL, "plf_comm_init: need 0 <= rank < world and an id");
    CKC(cudaSetDevice(ctx->device));
    if (!ctx->comm) ctx->comm = new Comm();
    if (comm_init(*ctx->comm, id, rank, world)) return fail(ctx, PLF_EUNSUPPORTED, std::string("plf_comm_init: ") + comm_error());
    return PLF_OK;
}

extern "C" int plf_comm_set_strip(plf_ctx *ctx, int own_first, int own_end)
{
    if (!ctx) return PLF_EINVAL;
    if (!ctx->comm || !ctx->comm->active()) return fail(ctx, PLF_ESTATE, "plf_comm_set_strip: call plf_comm_init first");
    if (!ctx->mesh_ready) return fail(ctx, PLF_ESTATE, "plf_comm_set_strip: call plf_mesh_structured first");
    const Comm &c = *ctx->comm;
    // one ghost node column at most on either side; the outer ranks own their outer boundary column
    const bool ok = (own_first == 0 || own_first == 1) && (own_end == ctx->NX || own_end == ctx->NX + 1) && own_first < own_end &&
                    (own_first == 1) == (c.rank > 0) && (own_end == ctx->NX) == (c.rank + 1 < c.world);
    if (!ok)
        return fail(ctx, PLF_EINVAL, "plf_comm_set_strip: owned node columns must be [halo, NX + last) with halo = (rank > 0), "
                                     "last = (rank == world - 1)");
    ctx->comm->own_first = own_first;
    ctx->comm->own_end = own_end;
    ctx->comm->strip_set = true;
    // The preconditioner of the distributed CG is one V-cycle per strip on the node columns the rank owns (+ its right
    // ghost column, masked): with a left halo column the hierarchy starts at node column 1, which keeps the number of
    // element columns under it at N / world (coarsenable) instead of N / world + 1.  All ranks must take the same code
    // path (the multigrid and the Jacobi variant issue different collectives): agree on it here.
    if (ctx->mg) {
        cudaError_t e = mg_build(*ctx->mg, ctx->NX, ctx->NY, own_first);
        if (e != cudaSuccess) return fail(ctx, PLF_ECUDA, std::string("plf_comm_set_strip: ") + cudaGetErrorString(e));
        model_apply_mg_options(ctx);
        ctx->mg_K_serial = -1;
    }
    if (comm_p2p_setup(*ctx->comm, 2 * (ctx->NY + 1), ctx->stream))
        return fail(ctx, PLF_ECUDA, std::string("plf_comm_set_strip: ") + comm_error());
    double have_mg = (ctx->mg && ctx->mg->levels() >= 2) ? 1. : 0.;
    int rc = comm_host_allreduce(ctx, &have_mg, 1, 2);
    if (rc != PLF_OK) return rc;
    ctx->comm->mg_all = have_mg > 0.5;
    return PLF_OK;
}

extern "C" int plf_comm_set_partition(plf_ctx *ctx, int own_first, int own_end, int first_col_global, int NX_global, double x_offset,
                                      double lenx_global)
{
    int rc = plf_comm_set_strip(ctx, own_first, own_end);
    if (rc != PLF_OK) return rc;
    if (first_col_global < 0 || first_col_global + ctx->NX > NX_global || !(lenx_global > 0.))
        return fail(ctx, PLF_EINVAL, "plf_comm_set_partition: the local mesh must lie inside the global one");
    Comm &c = *ctx->comm;
    c.first_col_global = first_col_global;
    c.NX_global = NX_global;
    c.x_offset = x_offset;
    c.lenx_global = lenx_global;
    c.global_set = true;
    return PLF_OK;
}

// host values combined over the ranks: op 0 sum, 1 max, 2 min
int comm_host_allreduce(plf_ctx *ctx, double *vals, int n, int op)
{
    if (!ctx->comm || !ctx->comm->active()) return PLF_OK;
    if (n < 1 || n > 64) return fail(ctx, PLF_EINVAL, "comm_host_allreduce: 1 <= n <= 64 values");
    double *d = reinterpret_cast<double *>(ctx->d_scalars + 1024);
    CKC(cudaMemcpyAsync(d, vals, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (comm_allreduce_op(*ctx->comm, d, n, op, ctx->stream)) return fail(ctx, PLF_ECUDA, std::string("allreduce: ") + comm_error());
    CKC(cudaMemcpyAsync(vals, d, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" int plf_comm_set_replica(plf_ctx *ctx, plf_ctx *global_ctx)
{
    if (!ctx) return PLF_EINVAL;
    if (!global_ctx) { ctx->replica = nullptr; return PLF_OK; }
    if (!ctx->comm || !ctx->comm->active() || !ctx->comm->global_set)
        return fail(ctx, PLF_ESTATE, "plf_comm_set_replica: call plf_comm_set_partition first");
    if (!global_ctx->mesh_ready || global_ctx->NX != ctx->comm->NX_global || global_ctx->NY != ctx->NY || global_ctx->device != ctx->device ||
        global_ctx->planestress != ctx->planestress)
        return fail(ctx, PLF_EINVAL, "plf_comm_set_replica: the second context must hold the GLOBAL mesh on the same device");
    ctx->replica = global_ctx;
    return PLF_OK;
}

extern "C" int plf_comm_p2p(plf_ctx *ctx) { return ctx && ctx->comm && ctx->comm->p2p ? 1 : 0; }

extern "C" int plf_comm_allreduce(plf_ctx *ctx, double *vals, int n)
{
    if (!ctx) return PLF_EINVAL;
    if (!ctx->comm || !ctx->comm->active()) return fail(ctx, PLF_ESTATE, "plf_comm_allreduce: call plf_comm_init first");
    if (!vals || n < 1 || n > 64) return fail(ctx, PLF_EINVAL, "plf_comm_allreduce: 1 <= n <= 64 values");
    CKC(cudaSetDevice(ctx->device));
    double *d = reinterpret_cast<double *>(ctx->d_scalars + 1024);   // scratch slot of the 4 KB scalar page
    CKC(cudaMemcpyAsync(d, vals, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (comm_allreduce_sum(*ctx->comm, d, n, ctx->stream)) return fail(ctx, PLF_ECUDA, std::string("plf_comm_allreduce: ") + comm_error());
    CKC(cudaMemcpyAsync(vals, d, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return PLF_OK;
}

extern "C" int plf_comm_destroy(plf_ctx *ctx)
{
    if (!ctx) return PLF_EINVAL;
    if (ctx->comm) { comm_destroy(*ctx->comm); delete ctx->comm; ctx->comm = nullptr; }
    return PLF_OK;
}
