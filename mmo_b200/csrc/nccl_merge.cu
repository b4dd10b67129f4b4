// nccl_merge.cu -- K7 across GPUs: the path's only collective.
// Every rank (one process per GPU) contributes its local top-k list (k x {f64 score, i64 frame},
// 16 B each) to one ncclAllGather over NVLink/NVSwitch, then merges the n_ranks lists on the host with
// the reference's tie rule (first pose in loop order = smallest frame, lds.ml:1099-1105).
// NCCL is loaded with dlopen so that single-GPU users do not need it.
#include "common.cuh"
#include <dlfcn.h>
#include <string.h>
#include <algorithm>

namespace mmo {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.handle) return MMO_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) { set_error("NCCL: cannot dlopen libnccl.so.2 (%s)", dlerror()); return MMO_ENCCL; }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(g_nccl.handle, "ncclAllGather");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.CommDestroy) {
        set_error("NCCL: libnccl.so.2 lacks a required symbol");
        return MMO_ENCCL;
    }
    return MMO_OK;
}
static int nccl_fail(ncclResult_t r, const char *what) {
    set_error("NCCL error %d (%s) in %s", r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
    return MMO_ENCCL;
}

struct Entry { double s; long long f; };

}  // namespace mmo

using namespace mmo;

extern "C" {

int mmo_nccl_unique_id(uint8_t id[128]) try {
    MMO_REQUIRE(id != nullptr, "mmo_nccl_unique_id: null pointer");
    MMO_TRY(nccl_load());
    ncclUniqueId u;
    ncclResult_t r = g_nccl.GetUniqueId(&u);
    if (r != 0) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(id, u.internal, 128);
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_nccl_init(int32_t rank, int32_t nranks, const uint8_t id[128]) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(id != nullptr && nranks > 0 && rank >= 0 && rank < nranks, "mmo_nccl_init: bad arguments");
    MMO_TRY(nccl_load());
    if (g_nccl.comm) { g_nccl.CommDestroy(g_nccl.comm); g_nccl.comm = nullptr; }
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&g_nccl.comm, nranks, u, rank);
    if (r != 0) return nccl_fail(r, "ncclCommInitRank");
    g_nccl.rank = rank;
    g_nccl.nranks = nranks;
    return MMO_OK;
} MMO_CATCH_ALL

int mmo_nccl_finalize(void) try {
    if (g_nccl.comm) { g_nccl.CommDestroy(g_nccl.comm); g_nccl.comm = nullptr; }
    g_nccl.nranks = 1;
    g_nccl.rank = 0;
    return MMO_OK;
} MMO_CATCH_ALL

// collective: every rank calls it with its local list (n_local <= k entries); every rank gets the merged top-k
int mmo_topk_allgather_merge(int32_t k, int32_t n_local, const double *scores, const int64_t *frames,
                             double *out_scores, int64_t *out_frames, int32_t *out_n) try {
    MMO_TRY(require_ready());
    MMO_REQUIRE(k > 0 && n_local >= 0 && n_local <= k && out_scores && out_frames && out_n, "mmo_topk_allgather_merge: bad arguments");
    MMO_REQUIRE(n_local == 0 || (scores && frames), "mmo_topk_allgather_merge: null list");
    MMO_REQUIRE(g_nccl.comm != nullptr, "mmo_topk_allgather_merge: call mmo_nccl_init first");
    Runtime &R = rt();
    const int nr = g_nccl.nranks;
    // one entry more than k: slot k carries the count
    std::vector<Entry> mine((size_t)k + 1);
    for (int i = 0; i < k; i++) { mine[i].s = i < n_local ? scores[i] : INFINITY; mine[i].f = i < n_local ? (long long)frames[i] : -1; }
    mine[k].s = 0.0;
    mine[k].f = n_local;
    DevBuf<Entry> d_mine, d_all;
    MMO_TRY(d_mine.upload(mine));
    MMO_TRY(d_all.alloc((size_t)nr * (k + 1)));
    ncclResult_t r = g_nccl.AllGather(d_mine.p, d_all.p, (size_t)(k + 1) * sizeof(Entry), /*ncclChar*/ 0, g_nccl.comm, R.stream);
    if (r != 0) return nccl_fail(r, "ncclAllGather");
    std::vector<Entry> all((size_t)nr * (k + 1));
    MMO_CUDA(cudaMemcpyAsync(all.data(), d_all.p, all.size() * sizeof(Entry), cudaMemcpyDeviceToHost, R.stream));
    MMO_CUDA(cudaStreamSynchronize(R.stream));
    std::vector<double> S((size_t)nr * k);
    std::vector<int64_t> F((size_t)nr * k);
    std::vector<int32_t> cnt(nr);
    for (int q = 0; q < nr; q++) {
        cnt[q] = (int32_t)all[(size_t)q * (k + 1) + k].f;
        for (int i = 0; i < k; i++) { S[(size_t)q * k + i] = all[(size_t)q * (k + 1) + i].s; F[(size_t)q * k + i] = all[(size_t)q * (k + 1) + i].f; }
    }
    return mmo_topk_merge(nr, k, S.data(), F.data(), cnt.data(), out_scores, out_frames, out_n);
} MMO_CATCH_ALL

}  // extern "C"
