// comm.cu — multi-GPU exchange of the per-bin statistics: one process per GPU, captures sharded by
// scan range, and only the Max / Min / pre-weighted Avg vectors (3 x fftSize float64, a few KB) cross NVLink:
// one grouped NCCL all-reduce (MAX, MIN, SUM).  The reference has no counterpart (single process, K:1139-1155).
//
// NCCL is resolved with dlopen at first use so that libkspec.so loads on hosts without it (single-GPU use).
#include "kspec_internal.h"
#include <dlfcn.h>
#include <nccl.h>
#include <stdio.h>
#include <string.h>
#include <new>
#include <vector>

namespace kspec {

namespace {

struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi& api() {
    static NcclApi a;
    if (a.h || a.ok) return a;
    // an already loaded libnccl (e.g. the one bundled with torch) is reused: same SONAME
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        a.h = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.h) break;
    }
    if (!a.h) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.h, "ncclAllReduce");
    a.AllGather = (decltype(a.AllGather))dlsym(a.h, "ncclAllGather");
    a.GroupStart = (decltype(a.GroupStart))dlsym(a.h, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.h, "ncclGroupEnd");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.AllReduce && a.AllGather && a.GroupStart && a.GroupEnd && a.CommDestroy && a.GetErrorString;
    return a;
}

int need_api() {
    if (!api().ok) {
        const char* e = dlerror();                 // one call: dlerror() clears the message it returns
        set_error("NCCL not available: %s", e ? e : "libnccl.so.2 could not be loaded");
        return KSPEC_ERR_NCCL;
    }
    return KSPEC_OK;
}

}  // namespace

}  // namespace kspec

using namespace kspec;

struct kspec_comm {
    ncclComm_t comm = nullptr;
    int device = 0, nRanks = 0, rank = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t evIn = nullptr, evCopied = nullptr, evDone = nullptr;
    double* buf = nullptr;
    size_t cap = 0;
    int64_t pendingN = 0;      // length of the vectors of the last asynchronous plan reduction held in buf (0: none)
    int64_t pendingSeq = -1;   // kspec_plan batch sequence number that reduction snapshotted
    // peer-memory exchange (kspec_comm_peer_setup): this rank's symmetric buffer and the peers' buffers mapped through CUDA IPC
    PeerExchange px;
    void* sym = nullptr;
    void* peerMapped[KSPEC_MAX_PEERS] = {};
};

namespace {
struct DevGuard {              // the caller's current device is restored on return (as the plan API does)
    int prev = -1;
    explicit DevGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
}  // namespace

#define NCK(call)                                                                         \
    do {                                                                                  \
        ncclResult_t r_ = (call);                                                         \
        if (r_ != ncclSuccess) {                                                          \
            set_error("%s failed: %s", #call, api().GetErrorString(r_));                  \
            return KSPEC_ERR_NCCL;                                                        \
        }                                                                                 \
    } while (0)
#define CCK(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            set_error("%s failed: %s", #call, cudaGetErrorString(e_));                    \
            return KSPEC_ERR_CUDA;                                                        \
        }                                                                                 \
    } while (0)

namespace kspec {
// grouped MAX / MIN / SUM on three consecutive device vectors of n float64 each
int comm_allreduce_device(kspec_comm* c, double* d3n, int64_t n, cudaStream_t st) {
    NCK(api().GroupStart());
    NCK(api().AllReduce(d3n, d3n, (size_t)n, ncclDouble, ncclMax, c->comm, st));
    NCK(api().AllReduce(d3n + n, d3n + n, (size_t)n, ncclDouble, ncclMin, c->comm, st));
    NCK(api().AllReduce(d3n + 2 * n, d3n + 2 * n, (size_t)n, ncclDouble, ncclSum, c->comm, st));
    NCK(api().GroupEnd());
    return KSPEC_OK;
}
}  // namespace kspec

extern "C" {

int kspec_comm_unique_id(char id[128]) {
    if (!id) { set_error("null argument"); return KSPEC_ERR_ARG; }
    int rc = need_api();
    if (rc) return rc;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId u;
    NCK(api().GetUniqueId(&u));
    memcpy(id, &u, 128);
    return KSPEC_OK;
}

int kspec_comm_init(kspec_comm** out, int nRanks, int rank, const char id[128], int device) {
    if (!out || !id || nRanks < 1 || rank < 0 || rank >= nRanks) { set_error("bad communicator arguments"); return KSPEC_ERR_ARG; }
    *out = nullptr;
    int rc = need_api();
    if (rc) return rc;
    DevGuard guard(device);
    kspec_comm* c = new (std::nothrow) kspec_comm();
    if (!c) { set_error("out of host memory"); return KSPEC_ERR_NOMEM; }
    c->device = device; c->nRanks = nRanks; c->rank = rank;
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclResult_t r = api().CommInitRank(&c->comm, nRanks, u, rank);
    if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", api().GetErrorString(r)); delete c; return KSPEC_ERR_NCCL; }
    if (cudaEventCreateWithFlags(&c->evIn, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&c->evCopied, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->evDone, cudaEventDisableTiming) != cudaSuccess || cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) { set_error("stream creation failed"); api().CommDestroy(c->comm); delete c; return KSPEC_ERR_CUDA; }
    *out = c;
    return KSPEC_OK;
}

int kspec_comm_allreduce_stats(kspec_comm* c, double* mx, double* mn, double* av, int64_t n) {
    if (!c || !mx || !mn || !av || n < 1) { set_error("bad all-reduce arguments"); return KSPEC_ERR_ARG; }
    DevGuard guard(c->device);
    const size_t bytes = (size_t)n * 8;
    c->pendingN = 0;                           // buf is reused: an unfetched plan reduction is gone
    if (c->cap < 3 * bytes) {
        CCK(cudaStreamSynchronize(c->st));
        if (c->buf) cudaFree(c->buf);
        c->buf = nullptr; c->cap = 0;
        CCK(cudaMalloc(&c->buf, 3 * bytes));
        c->cap = 3 * bytes;
    }
    CCK(cudaMemcpyAsync(c->buf, mx, bytes, cudaMemcpyHostToDevice, c->st));
    CCK(cudaMemcpyAsync(c->buf + n, mn, bytes, cudaMemcpyHostToDevice, c->st));
    CCK(cudaMemcpyAsync(c->buf + 2 * n, av, bytes, cudaMemcpyHostToDevice, c->st));
    int rc = comm_allreduce_device(c, c->buf, n, c->st);
    if (rc) return rc;
    CCK(cudaMemcpyAsync(mx, c->buf, bytes, cudaMemcpyDeviceToHost, c->st));
    CCK(cudaMemcpyAsync(mn, c->buf + n, bytes, cudaMemcpyDeviceToHost, c->st));
    CCK(cudaMemcpyAsync(av, c->buf + 2 * n, bytes, cudaMemcpyDeviceToHost, c->st));
    CCK(cudaStreamSynchronize(c->st));
    return KSPEC_OK;
}

int kspec_comm_allreduce_sum(kspec_comm* c, double* v, int64_t n) {
    if (!c || !v || n < 1) { set_error("bad all-reduce arguments"); return KSPEC_ERR_ARG; }
    DevGuard guard(c->device);
    const size_t bytes = (size_t)n * 8;
    if (c->cap < bytes) {
        CCK(cudaStreamSynchronize(c->st));
        if (c->buf) cudaFree(c->buf);
        c->buf = nullptr; c->cap = 0;
        CCK(cudaMalloc(&c->buf, bytes));
        c->cap = bytes;
    }
    c->pendingN = 0;
    CCK(cudaMemcpyAsync(c->buf, v, bytes, cudaMemcpyHostToDevice, c->st));
    NCK(api().AllReduce(c->buf, c->buf, (size_t)n, ncclDouble, ncclSum, c->comm, c->st));
    CCK(cudaMemcpyAsync(v, c->buf, bytes, cudaMemcpyDeviceToHost, c->st));
    CCK(cudaStreamSynchronize(c->st));
    return KSPEC_OK;
}

int kspec_comm_allreduce_plan(kspec_comm* c, kspec_plan* plan) {
    if (!c || !plan) { set_error("bad all-reduce arguments"); return KSPEC_ERR_ARG; }
    double* stats = nullptr;
    int F = 0;
    cudaStream_t st = nullptr;
    int64_t seq = 0;
    if (!plan_stats_view(plan, &stats, &F, &st, &seq)) { set_error("plan holds no batch statistics yet"); return KSPEC_ERR_STATE; }
    DevGuard guard(c->device);
    const size_t bytes = (size_t)3 * F * 8;
    if (c->cap < bytes) {
        CCK(cudaStreamSynchronize(c->st));
        if (c->buf) cudaFree(c->buf);
        c->buf = nullptr; c->cap = 0;
        CCK(cudaMalloc(&c->buf, bytes));
        c->cap = bytes;
    }
    // snapshot the plan's statistics on the communicator's stream, release the plan's stream at once, reduce in the
    // background: the next batch's kernels overlap the (latency-bound) NVLink exchange
    CCK(cudaEventRecord(c->evIn, st));
    CCK(cudaStreamWaitEvent(c->st, c->evIn, 0));
    CCK(cudaMemcpyAsync(c->buf, stats, bytes, cudaMemcpyDeviceToDevice, c->st));
    CCK(cudaEventRecord(c->evCopied, c->st));
    CCK(cudaStreamWaitEvent(st, c->evCopied, 0));
    int rc = comm_allreduce_device(c, c->buf, F, c->st);
    if (rc) return rc;
    CCK(cudaEventRecord(c->evDone, c->st));
    c->pendingN = F;
    c->pendingSeq = seq;
    return KSPEC_OK;
}

int kspec_comm_join(kspec_comm* c, kspec_plan* plan) {
    if (!c || !plan) { set_error("bad join arguments"); return KSPEC_ERR_ARG; }
    double* stats = nullptr;
    int F = 0;
    cudaStream_t st = nullptr;
    int64_t seq = 0;
    if (!plan_stats_view(plan, &stats, &F, &st, &seq)) { set_error("plan holds no batch statistics yet"); return KSPEC_ERR_STATE; }
    if (c->pendingN != F) { set_error("no reduction of this plan is pending"); return KSPEC_ERR_STATE; }
    if (c->pendingSeq != seq) {
        // the plan has run another batch since the snapshot: its statistics belong to that batch and stay untouched
        set_error("the plan has advanced since kspec_comm_allreduce_plan: read the reduced vectors with kspec_comm_fetch_reduced");
        return KSPEC_ERR_STATE;
    }
    DevGuard guard(c->device);
    // the plan's stream waits for the reduction and takes the reduced vectors back: kspec_zerospan_fetch then returns them
    CCK(cudaStreamWaitEvent(st, c->evDone, 0));
    CCK(cudaMemcpyAsync(stats, c->buf, (size_t)3 * F * 8, cudaMemcpyDeviceToDevice, st));
    CCK(cudaEventRecord(c->evIn, st));                    // the communicator's next snapshot must not overtake this copy
    CCK(cudaStreamWaitEvent(c->st, c->evIn, 0));
    c->pendingN = 0;
    return KSPEC_OK;
}

int kspec_comm_fetch_reduced(kspec_comm* c, double* mx, double* mn, double* av, int64_t n) {
    if (!c || !mx || !mn || !av || n < 1) { set_error("bad fetch arguments"); return KSPEC_ERR_ARG; }
    if (c->pendingN != n) { set_error("no reduction of %lld-element vectors is pending", (long long)n); return KSPEC_ERR_STATE; }
    DevGuard guard(c->device);
    const size_t bytes = (size_t)n * 8;
    CCK(cudaMemcpyAsync(mx, c->buf, bytes, cudaMemcpyDeviceToHost, c->st));
    CCK(cudaMemcpyAsync(mn, c->buf + n, bytes, cudaMemcpyDeviceToHost, c->st));
    CCK(cudaMemcpyAsync(av, c->buf + 2 * n, bytes, cudaMemcpyDeviceToHost, c->st));
    CCK(cudaStreamSynchronize(c->st));
    c->pendingN = 0;
    return KSPEC_OK;
}

// ---- the exchange as the tail of the compute kernel: peer-memory writes instead of a collective call -------------------------
int kspec_comm_peer_setup(kspec_comm* c, kspec_plan* plan) {
    if (!c || !plan) { set_error("bad peer setup arguments"); return KSPEC_ERR_ARG; }
    if (c->nRanks > KSPEC_MAX_PEERS) { set_error("peer exchange supports up to %d ranks", KSPEC_MAX_PEERS); return KSPEC_ERR_UNSUPPORTED; }
    DevGuard guard(c->device);
    if (c->sym) {
        // already set up: attach another plan of the same fftSize as well (every rank must do the same; the plans share the
        // buffers and the sequence, so their sharded batches must be issued in the same order on every rank)
        return plan_attach_peer(plan, &c->px);
    }
    kspec_plan_info_t info;
    if (kspec_plan_info(plan, &info) != KSPEC_OK) return KSPEC_ERR_ARG;
    const int F = info.fft_size;
    const int n = c->nRanks;
    const size_t dataBytes = (size_t)2 * n * 3 * F * 8, flagBytes = (size_t)2 * n * 8;
    const size_t total = dataBytes + flagBytes + 256;
    CCK(cudaMalloc(&c->sym, total));
    CCK(cudaMemset(c->sym, 0, total));
    cudaIpcMemHandle_t mine;
    CCK(cudaIpcGetMemHandle(&mine, c->sym));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    // all-gather of the 64-byte handles over the communicator that already exists
    char* dH = nullptr;
    CCK(cudaMalloc(&dH, (size_t)n * 64));
    CCK(cudaMemcpyAsync(dH + (size_t)c->rank * 64, &mine, 64, cudaMemcpyHostToDevice, c->st));
    NCK(api().AllGather(dH + (size_t)c->rank * 64, dH, 64, ncclChar, c->comm, c->st));
    std::vector<cudaIpcMemHandle_t> all(n);
    CCK(cudaMemcpyAsync(all.data(), dH, (size_t)n * 64, cudaMemcpyDeviceToHost, c->st));
    CCK(cudaStreamSynchronize(c->st));
    cudaFree(dH);
    c->px = PeerExchange();
    c->px.nRanks = n; c->px.rank = c->rank; c->px.F = F;
    bool failed = false;
    char why[160] = "";
    for (int r = 0; r < n; ++r) {
        void* base = c->sym;
        if (r != c->rank) {
            cudaError_t e = cudaIpcOpenMemHandle(&base, all[r], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                snprintf(why, sizeof(why), "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
                cudaGetLastError();
                failed = true;
                base = nullptr;
            }
            c->peerMapped[r] = base;
        }
        c->px.slots[r] = (double*)base;
        c->px.flags[r] = (unsigned long long*)((char*)base + dataBytes);
    }
    void* small = nullptr;
    CCK(cudaMalloc(&small, 64));
    CCK(cudaMemset(small, 0, 64));
    c->px.counter = (unsigned int*)small;
    c->px.status = (int*)((char*)small + 32);
    // nobody may start writing before every rank has mapped everything, and every rank must learn whether ANY rank failed (a
    // rank that left early would leave the others waiting in their first exchange): one tiny all-reduce as barrier + vote
    double* dz = nullptr;
    CCK(cudaMalloc(&dz, 8));
    const double vote = failed ? 1.0 : 0.0;
    CCK(cudaMemcpyAsync(dz, &vote, 8, cudaMemcpyHostToDevice, c->st));
    NCK(api().AllReduce(dz, dz, 1, ncclDouble, ncclSum, c->comm, c->st));
    double votes = 0.0;
    CCK(cudaMemcpyAsync(&votes, dz, 8, cudaMemcpyDeviceToHost, c->st));
    CCK(cudaStreamSynchronize(c->st));
    cudaFree(dz);
    if (votes > 0.0) {
        for (int r = 0; r < KSPEC_MAX_PEERS; ++r) if (c->peerMapped[r]) { cudaIpcCloseMemHandle(c->peerMapped[r]); c->peerMapped[r] = nullptr; }
        cudaFree(small);
        cudaFree(c->sym);
        c->sym = nullptr;
        c->px = PeerExchange();
        set_error("peer exchange unavailable (%s): peer access between the GPUs is required; use kspec_comm_allreduce_*", failed ? why : "a peer rank could not map the buffers");
        return KSPEC_ERR_UNSUPPORTED;
    }
    return plan_attach_peer(plan, &c->px);
}

int kspec_comm_peer_status(kspec_comm* c, int* timedOut) {
    if (!c || !timedOut) { set_error("bad argument"); return KSPEC_ERR_ARG; }
    *timedOut = 0;
    if (!c->sym) return KSPEC_OK;
    DevGuard guard(c->device);
    CCK(cudaMemcpy(timedOut, c->px.status, sizeof(int), cudaMemcpyDeviceToHost));
    return KSPEC_OK;
}

int kspec_comm_finalize(kspec_comm* c) {
    if (!c) return KSPEC_OK;
    DevGuard guard(c->device);
    if (c->sym) {
        cudaDeviceSynchronize();         // plans that were attached must not run sharded batches after this point
        if (c->comm && api().ok) {      // every rank has stopped using the peers' buffers before anything is unmapped or freed
            double* dz = nullptr;
            if (cudaMalloc(&dz, 8) == cudaSuccess) {
                cudaMemset(dz, 0, 8);
                api().AllReduce(dz, dz, 1, ncclDouble, ncclSum, c->comm, c->st);
                cudaStreamSynchronize(c->st);
                cudaFree(dz);
            }
        }
        for (int r = 0; r < KSPEC_MAX_PEERS; ++r) if (c->peerMapped[r]) cudaIpcCloseMemHandle(c->peerMapped[r]);
        if (c->px.counter) cudaFree(c->px.counter);
        cudaFree(c->sym);
        c->sym = nullptr;
    }
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->comm && api().ok) api().CommDestroy(c->comm);
    if (c->buf) cudaFree(c->buf);
    for (cudaEvent_t e : {c->evIn, c->evCopied, c->evDone}) if (e) cudaEventDestroy(e);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
    return KSPEC_OK;
}

}  // extern "C"
