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

// Nearest-neighbour halo exchange of one row each way (grouped send/recv):
// send `first_row` to `up` and receive its last row into `top_ghost`; send
// `last_row` to `down` and receive its first row into `bottom_ghost`.
// up/down < 0: no neighbour on that side.
int comm_halo(Comm* c, int up, int down, const double* first_row, double* top_ghost,
              const double* last_row, double* bottom_ghost, size_t n, cudaStream_t st);
// all-gather of one double per rank: recv[r] = *send of rank r
int comm_allgather1(Comm* c, const double* send, double* recv, cudaStream_t st);

}  // namespace xsq
