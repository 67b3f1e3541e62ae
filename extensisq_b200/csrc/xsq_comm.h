// Internal: a thin NCCL communicator wrapper.  libnccl is dlopen'ed lazily so
// that libxsq.so loads on machines without NCCL / without a GPU.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace xsq {

struct Comm;   // opaque

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
