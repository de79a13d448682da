// Shared device helpers for libbpc_b200 (sm_100a).
//
// Floating-point policy: every operation whose rounding is part of the contract is written with an
// explicit round-to-nearest intrinsic (__dmul_rn, __dadd_rn, __fma_rn, __ddiv_rn, __fmul_rn,
// __fadd_rn) so the compiler can neither contract nor reassociate it; the translation units are
// additionally built with --fmad=false.  Where the reference's BLAS call fuses (OpenBLAS ddot /
// dgemm / dgemv kernels use FMA) the same fused order is spelled out with __fma_rn.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "bpc_b200.h"

namespace bpc {

extern std::atomic<unsigned long long> g_launches;   // host-side launch counter (api.cu)

#define BPC_LAUNCH_CHECK()                                  \
    do {                                                    \
        ::bpc::g_launches.fetch_add(1, std::memory_order_relaxed); \
        cudaError_t e__ = cudaGetLastError();               \
        if (e__ != cudaSuccess) return (int)e__;            \
    } while (0)

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }

// sum_k a_k*b_k, k ascending, fused: what np.dot (ddot) and 3x3 / 3x4 np.matmul (dgemm) do.
__device__ __forceinline__ double dot3_seq(double a0, double b0, double a1, double b1, double a2, double b2) {
    return dfma(a2, b2, dfma(a1, b1, dmul(a0, b0)));
}
// row-major (3x3) @ (3,) through dgemv: products accumulated in the order 1, 0, 2.
__device__ __forceinline__ double dot3_gemv(double a0, double b0, double a1, double b1, double a2, double b2) {
    return dfma(a2, b2, dfma(a0, b0, dmul(a1, b1)));
}

// exact cost value the reference stores: float32(((e12 + e13) + e23) / 3), epipolar_matching.py:81,96
__device__ __forceinline__ float cost_from_sum(double s) { return __double2float_rn(ddiv(s, 3.0)); }

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ double shfl_d(double v, int src) {
    return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ double shfl_xor_d(double v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}

}  // namespace bpc
