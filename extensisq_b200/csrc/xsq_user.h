// Internal: runtime-specialised kernels (NVRTC) for user right-hand sides and
// user tableaux.  See xsq_user.cpp.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "xsq_rk_core.cuh"

namespace xsq {

struct MethodInfo {
    int s, order, order2, fsal, npol;
    double sc[4];
    double stbrad, tanang;
};

void set_detail(const std::string& s);
void count_launch();
// xsq_profile_*: CUDA events around init / persistent kernel / queue kernels (xsq_api.cu)
void prof_mark(int i, cudaStream_t st);

bool user_tableau_info(MethodInfo* mi);
bool user_rhs_shape(int rhs, int* n_state, int* n_param);
int user_rk_launch(int method, int rhs, int events, const RkDev& P, int cost, double stbrad,
                   double tanang, cudaStream_t st);
// number of event functions behind an events handle, -1 if unknown
int user_events_count(int events);
// user PDE for SSV2stab: CUfunctions (as void*) of the eval / stage / final kernels
int user_pde_kernels(int pde, void* fn[3], int* n_param);
int user_pde_vector_size(int pde);   // > 0: a general system registered with xsq_pde_register_vector_source
int user_launch(void* fn, unsigned gx, unsigned gy, unsigned bx, unsigned by, void** args,
                cudaStream_t st);

}  // namespace xsq
