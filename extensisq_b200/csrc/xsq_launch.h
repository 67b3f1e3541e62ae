// Internal (not part of the C ABI): per-method launch entry points, one
// translation unit per tableau so the build parallelises.
#pragma once
#include <cuda_runtime.h>
#include "xsq_rk_core.cuh"

namespace xsq {

struct LaunchInfo {       // filled by the launcher for diagnostics/benchmarks
    int grid, block, blocks_per_sm, regs;
};

#define XSQ_DECL_LAUNCH(T) \
    int launch_##T(int rhs, const RkDev& P, cudaStream_t st, LaunchInfo* info);
XSQ_DECL_LAUNCH(Ts5)
XSQ_DECL_LAUNCH(BS5)
XSQ_DECL_LAUNCH(CK5)
XSQ_DECL_LAUNCH(Me4)
XSQ_DECL_LAUNCH(Pr7)
XSQ_DECL_LAUNCH(Pr8)
XSQ_DECL_LAUNCH(Pr9)
XSQ_DECL_LAUNCH(CFMR7osc)
XSQ_DECL_LAUNCH(CKdisc)
XSQ_DECL_LAUNCH(Fi4N)
XSQ_DECL_LAUNCH(Fi5N)
XSQ_DECL_LAUNCH(Mu5Nmb)
XSQ_DECL_LAUNCH(MR6NN)
#undef XSQ_DECL_LAUNCH

int launch_swag(int rhs, const RkDev& P, cudaStream_t st);
int launch_ens_init(int rhs, const RkDev& P, cudaStream_t st);
int launch_stiff_queue(int rhs, const RkDev& P, int cost, double stbrad, double tanang,
                       cudaStream_t st);

void count_launch();

}  // namespace xsq
