// Multi-GPU extended window: the NCCL data plane inside the library.
//
// One store per GPU / rank holds a shard of the open window (record idx -> rank idx % world); a batch is given
// to every rank.  pixiu_mg_setitem_batch runs the three phases of encode.cu ("Multi-GPU extended window") with
// the two collectives between them issued HERE, on the store's own stream:
//     phase A (SA + LCP + LPF over shard + batch)  ->  ncclAllReduce(MAX, u32 M[batch positions])
//     phase B (flags, pair rule, runs, candidates) ->  ncclAllReduce(MIN, u32 (idx << 16 | to)[long runs])
//     phase C (emit; every rank stores the identical encoded batch)
// Kernels and collectives are ordered by the stream: no host synchronisation sits between a phase and its
// collective (the one host read-back left is the number of long runs, which sizes the second collective).
// This replaces the reference's single-process window (SuffixTree::setitem, SuffixTree.cpp:291-304, rotation
// PiXiuCtrl.cpp:13-25) for BASELINE config 5; the reference itself has no communication layer at all.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, or the path in PIXIU_NCCL_LIB): a process that already
// carries an NCCL (PyTorch bundles one) shares that copy instead of loading a second one, and the library still
// loads - and reports every symbol - on a machine without NCCL.  Without NCCL the pixiu_mg_comm_* calls fail
// loudly (PIXIU_EINVAL + pixiu_last_error); there is no fallback.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "store.h"

namespace pixiu {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    std::string err;
    bool load() {
        if (handle) return true;
        const char *path = getenv("PIXIU_NCCL_LIB");
        const char *cands[] = {path, "libnccl.so.2", "libnccl.so"};
        for (const char *c : cands) {
            if (!c || !*c) continue;
            handle = dlopen(c, RTLD_NOW | RTLD_LOCAL);
            if (handle) break;
            err = dlerror();
        }
        if (!handle) return false;
        auto sym = [&](const char *n) { return dlsym(handle, n); };
        GetUniqueId = reinterpret_cast<decltype(GetUniqueId)>(sym("ncclGetUniqueId"));
        CommInitRank = reinterpret_cast<decltype(CommInitRank)>(sym("ncclCommInitRank"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(sym("ncclCommDestroy"));
        AllReduce = reinterpret_cast<decltype(AllReduce)>(sym("ncclAllReduce"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(sym("ncclGetErrorString"));
        GetVersion = reinterpret_cast<decltype(GetVersion)>(sym("ncclGetVersion"));
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString) {
            err = "libnccl lacks a required symbol";
            dlclose(handle);
            handle = nullptr;
            return false;
        }
        return true;
    }
};

static NcclApi &nccl_api() {
    static NcclApi api;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    api.load();
    return api;
}

struct MgComm {
    ncclComm_t comm = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // accumulated since comm_init: collective time on the stream (CUDA events) and payload bytes
    double max_ms = 0, min_ms = 0;
    int64_t max_bytes = 0, min_bytes = 0, batches = 0;
    int version = 0;
    ~MgComm() {
        for (auto &e : ev)
            if (e) cudaEventDestroy(e);
        if (comm) nccl_api().CommDestroy(comm);
    }
};

void mg_comm_free(MgComm *c) { delete c; }

#define PX_NCCL(S, expr)                                                                              \
    do {                                                                                              \
        ncclResult_t _r = (expr);                                                                     \
        if (_r != ncclSuccess) {                                                                      \
            (S).err = std::string(#expr) + ": " + nccl_api().GetErrorString(_r);                       \
            return PIXIU_ECUDA;                                                                       \
        }                                                                                             \
    } while (0)

int mg_comm_init(Store &S, int rank, int world, const uint8_t *id) {
    if (world < 1 || rank < 0 || rank >= world || S.n_records() != 0 || S.mg_comm) return PIXIU_EINVAL;
    NcclApi &A = nccl_api();
    if (!A.handle) {
        S.err = "NCCL not available (" + A.err + "); set PIXIU_NCCL_LIB";
        return PIXIU_EINVAL;
    }
    ncclUniqueId uid;
    static_assert(sizeof(uid) == PIXIU_NCCL_UNIQUE_ID_BYTES, "ncclUniqueId size");
    memcpy(&uid, id, sizeof(uid));
    MgComm *c = new MgComm();
    ncclResult_t r = A.CommInitRank(&c->comm, world, uid, rank);
    if (r != ncclSuccess) {
        S.err = std::string("ncclCommInitRank: ") + A.GetErrorString(r);
        c->comm = nullptr;
        delete c;
        return PIXIU_ECUDA;
    }
    for (auto &e : c->ev) PX_CUDA(cudaEventCreate(&e));
    if (A.GetVersion) A.GetVersion(&c->version);
    // NCCL connects its channels lazily, on the first collective of a communicator (hundreds of milliseconds): do
    // that here, with one all-reduce of each kind and of a typical size, instead of inside the first batch
    if (world > 1) {
        DevBuf<uint32_t> tmp;
        const size_t n_max = 4u << 20, n_min = 64u << 10;
        tmp.reserve_discard(n_max);
        PX_CUDA(cudaMemsetAsync(tmp.p, 0, n_max * sizeof(uint32_t), S.st));
        ncclResult_t r1 = A.AllReduce(tmp.p, tmp.p, n_max, ncclUint32, ncclMax, c->comm, S.st);
        ncclResult_t r2 = A.AllReduce(tmp.p, tmp.p, n_min, ncclUint32, ncclMin, c->comm, S.st);
        PX_CUDA(cudaStreamSynchronize(S.st));
        if (r1 != ncclSuccess || r2 != ncclSuccess) {
            S.err = std::string("NCCL warm-up all-reduce: ") + A.GetErrorString(r1 != ncclSuccess ? r1 : r2);
            delete c;
            return PIXIU_ECUDA;
        }
    }
    S.mg_comm = c;
    S.mg_rank = rank;
    S.mg_world = world;
    return PIXIU_OK;
}

int mg_setitem_nccl(Store &S, int64_t n, const uint8_t *d_keys, const int64_t *d_koff, const uint8_t *d_vals,
                    const int64_t *d_voff, const uint8_t *h_keys, const int64_t *h_koff, const int64_t *h_voff,
                    int32_t *rc, int32_t *saved) {
    MgComm *C = S.mg_comm;
    if (!C) {
        S.err = "pixiu_mg_setitem_batch needs pixiu_mg_comm_init first";
        return PIXIU_EINVAL;
    }
    NcclApi &A = nccl_api();
    uint32_t *d_m = nullptr, *d_cand = nullptr;
    int64_t cm = 0, cc = 0;
    int r = S.mg_begin(n, d_keys, d_koff, d_vals, d_voff, h_keys, h_koff, h_voff, &d_m, &cm, false);
    if (r != PIXIU_OK) return r;
    PX_CUDA(cudaEventRecord(C->ev[0], S.st));
    if (cm && S.mg_world > 1) PX_NCCL(S, A.AllReduce(d_m, d_m, (size_t) cm, ncclUint32, ncclMax, C->comm, S.st));
    PX_CUDA(cudaEventRecord(C->ev[1], S.st));
    r = S.mg_mid(&d_cand, &cc, false);
    if (r != PIXIU_OK) return r;
    PX_CUDA(cudaEventRecord(C->ev[2], S.st));
    if (cc && S.mg_world > 1) PX_NCCL(S, A.AllReduce(d_cand, d_cand, (size_t) cc, ncclUint32, ncclMin, C->comm, S.st));
    PX_CUDA(cudaEventRecord(C->ev[3], S.st));
    r = S.mg_end(rc, saved);  // (synchronises the stream: the events above have completed)
    if (r != PIXIU_OK) return r;
    float a = 0, b = 0;
    PX_CUDA(cudaEventElapsedTime(&a, C->ev[0], C->ev[1]));
    PX_CUDA(cudaEventElapsedTime(&b, C->ev[2], C->ev[3]));
    C->max_ms += a;
    C->min_ms += b;
    C->max_bytes += cm * 4;
    C->min_bytes += cc * 4;
    C->batches++;
    return PIXIU_OK;
}

int mg_unique_id(uint8_t *id, std::string &err) {
    NcclApi &A = nccl_api();
    if (!A.handle) {
        err = "NCCL not available (" + A.err + "); set PIXIU_NCCL_LIB";
        return PIXIU_EINVAL;
    }
    ncclUniqueId uid;
    ncclResult_t r = A.GetUniqueId(&uid);
    if (r != ncclSuccess) {
        err = std::string("ncclGetUniqueId: ") + A.GetErrorString(r);
        return PIXIU_ECUDA;
    }
    memcpy(id, &uid, sizeof(uid));
    return PIXIU_OK;
}

void mg_comm_stats(const Store &S, pixiu_mg_stats *o) {
    memset(o, 0, sizeof(*o));
    o->rank = S.mg_rank;
    o->world = S.mg_world;
    if (const MgComm *C = S.mg_comm) {
        o->nccl_version = C->version;
        o->batches = C->batches;
        o->max_reduce_bytes = C->max_bytes;
        o->min_reduce_bytes = C->min_bytes;
        o->max_reduce_ms = C->max_ms;
        o->min_reduce_ms = C->min_ms;
    }
}

}  // namespace pixiu

