// Internal: a thin NCCL communicator wrapper.  libnccl is dlopen'ed lazily so
// that libxsq.so loads on machines without NCCL / without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace xsq {

struct Comm;   // opaque

// SSV2stab's slab storage on a multi-GPU communicator: allocated, exported with
// CUDA IPC and mapped by the neighbours ONCE and kept for the next solve of the
// same shape (mapping a few GB of peer memory costs hundreds of milliseconds,
// more than the stages of a short solve).  Owned by the communicator, released
// by comm_destroy.
struct CommWorkspace {
    double* buf = nullptr;
    size_t doubles = 0;
    long long key[6] = {0, 0, 0, 0, 0, 0};   // nx, rows_global, rows_local, world, rank, extras
    const double* peer_up = nullptr;         // IPC mapping of the neighbours' storage
    const double* peer_dn = nullptr;
    int rows_up = 0;
    long long seq = 0;                       // handshake sequence number, never reset
    char* xchg = nullptr;                    // small device buffer for the handle exchange
};
CommWorkspace* comm_workspace(Comm* c);
void comm_workspace_release(CommWorkspace* w);

int comm_unique_id(char id[128]);
int comm_create(int rank, int world, const char id[128], Comm** out);
void comm_destroy(Comm* c);
int comm_rank(const Comm* c);
int comm_world(const Comm* c);

// all-gather of `nbytes` bytes per rank (device buffers)
int comm_allgather_bytes(Comm* c, const void* send, void* recv, size_t nbytes, cudaStream_t st);
// all-gather of one double per rank: recv[r] = *send of rank r
int comm_allgather1(Comm* c, const double* send, double* recv, cudaStream_t st);

}  // namespace xsq
