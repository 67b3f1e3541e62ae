// TEST INFRASTRUCTURE -- lets g++ compile the kernel sources of
// extensisq_b200/csrc for the HOST: one "thread" (lane 0 of a one-lane warp in
// a one-thread block) executes a kernel body as ordinary C++.  Arithmetic is
// the same IEEE add / mul / fma in the same order (-ffp-contract=off), the
// reciprocal seed comes from the B200 table of the oracle, so a kernel run here
// must equal the C oracle in device arithmetic bit for bit -- a check of the
// real kernel source that needs no GPU.  Never part of the product.
#pragma once
#include <cuda_runtime.h>      // vector types, empty __device__ / __global__ for g++
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#undef __shared__
#define __shared__ static       // one thread: static storage is the block's shared memory
#undef __constant__
#define __constant__
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __noinline__
#define __noinline__
#define __maxnreg__(n)

struct EmuIdx { unsigned x, y, z; };
static EmuIdx threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

static inline int __double2hiint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)(b >> 32); }
static inline int __double2loint(double x) { uint64_t b; std::memcpy(&b, &x, 8); return (int)(uint32_t)b; }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; std::memcpy(&x, &b, 8); return x;
}
static inline double __longlong_as_double(long long v) { double x; std::memcpy(&x, &v, 8); return x; }
static inline long long __double_as_longlong(double x) { long long v; std::memcpy(&v, &x, 8); return v; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
// a warp of ONE lane
static inline unsigned __ballot_sync(unsigned, bool p) { return p ? 1u : 0u; }
static inline bool __any_sync(unsigned, bool p) { return p; }
static inline bool __all_sync(unsigned, bool p) { return p; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
    const unsigned long long o = *p; *p = o + v; return o;
}
static inline int atomicAdd(int* p, int v) { const int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
static inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) {
    const unsigned long long o = *p; *p = v; return o;
}
static inline unsigned atomicExch(unsigned* p, unsigned v) { const unsigned o = *p; *p = v; return o; }
static inline unsigned atomicMax(unsigned* p, unsigned v) { const unsigned o = *p; if (v > o) *p = v; return o; }
static inline void __threadfence_block() {}
static inline void __nanosleep(unsigned) {}
static inline int atomicOr(int* p, int v) { const int o = *p; *p = o | v; return o; }
using std::fma;
