// xsq_comm.cpp -- NCCL plumbing for the domain-decomposed SSV2stab path: the
// set-up exchange of CUDA-IPC handles and, per step attempt, the all-gather of
// one scalar per rank (error norm).  The per-stage halo does NOT go through
// NCCL: the stage kernel loads the neighbour's boundary row in place over
// NVLink (xsq_rkc.cu).  The reference has no communication layer
// at all (single Python thread); this file is new.
#include "xsq_comm.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <string>

#include "xsq.h"
#include "xsq_user.h"

namespace xsq {

struct Comm {
    ncclComm_t nccl;
    int rank, world;
    CommWorkspace ws;
};

namespace {
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t,
                              cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
} g;
std::mutex g_mu;

template <class F>
bool bind(const char* name, F* out) {
    *out = reinterpret_cast<F>(dlsym(g.lib, name));
    return *out != nullptr;
}

bool load() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.lib) return true;
    // inside a torch process this resolves to the libnccl torch already loaded
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g.lib) break;
    }
    if (!g.lib) { set_detail("cannot dlopen libnccl.so.2"); return false; }
    bool ok = bind("ncclGetUniqueId", &g.GetUniqueId) &&
              bind("ncclCommInitRank", &g.CommInitRank) &&
              bind("ncclCommDestroy", &g.CommDestroy) &&
              bind("ncclAllGather", &g.AllGather) &&
              bind("ncclGetErrorString", &g.GetErrorString);
    if (!ok) set_detail("libnccl is missing a required symbol");
    return ok;
}

int check(ncclResult_t r, const char* what) {
    if (r == ncclSuccess) return 0;
    set_detail(std::string(what) + ": " + g.GetErrorString(r));
    return -1;
}
}  // namespace

int comm_unique_id(char id[128]) {
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    if (!load()) return XSQ_ERR_CUDA;
    ncclUniqueId u;
    if (check(g.GetUniqueId(&u), "ncclGetUniqueId")) return XSQ_ERR_CUDA;
    std::memcpy(id, &u, 128);
    return XSQ_OK;
}

int comm_create(int rank, int world, const char id[128], Comm** out) {
    if (!load()) return XSQ_ERR_CUDA;
    ncclUniqueId u;
    std::memcpy(&u, id, 128);
    Comm* c = new Comm{nullptr, rank, world, CommWorkspace()};
    if (check(g.CommInitRank(&c->nccl, world, u, rank), "ncclCommInitRank")) {
        delete c;
        return XSQ_ERR_CUDA;
    }
    *out = c;
    return XSQ_OK;
}

CommWorkspace* comm_workspace(Comm* c) { return c ? &c->ws : nullptr; }

void comm_workspace_release(CommWorkspace* w) {
    if (!w) return;
    if (w->peer_up) cudaIpcCloseMemHandle((void*)w->peer_up);
    if (w->peer_dn) cudaIpcCloseMemHandle((void*)w->peer_dn);
    if (w->buf) { cudaDeviceSynchronize(); cudaFree(w->buf); }
    if (w->xchg) cudaFree(w->xchg);
    const long long seq = w->seq;
    *w = CommWorkspace();
    w->seq = seq;
}

void comm_destroy(Comm* c) {
    if (!c) return;
    comm_workspace_release(&c->ws);
    if (g.lib) g.CommDestroy(c->nccl);
    delete c;
}
int comm_rank(const Comm* c) { return c->rank; }
int comm_world(const Comm* c) { return c->world; }

int comm_allgather_bytes(Comm* c, const void* send, void* recv, size_t nbytes, cudaStream_t st) {
    return check(g.AllGather(send, recv, nbytes, ncclChar, c->nccl, st), "ncclAllGather");
}

int comm_allgather1(Comm* c, const double* send, double* recv, cudaStream_t st) {
    return check(g.AllGather(send, recv, 1, ncclDouble, c->nccl, st), "ncclAllGather");
}

}  // namespace xsq
