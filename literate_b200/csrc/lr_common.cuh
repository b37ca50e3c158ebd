// Shared device/host helpers for the literate_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/literate_b200.h"

#ifndef __CUDACC__
#error "literate_b200 is CUDA-only: there is no CPU path"
#endif

#define LR_WARP 32
#define LR_SLOTS 32            // register slots per side of one chain (one per lane)
#define LR_FIX_SHIFT 52        // fractions of a year are accumulated in 2^-52 fixed point
#define LR_FIX_SCALE 4503599627370496.0   // 2^52

// ---- error plumbing ------------------------------------------------------------------------
void lr_set_error(const char* fmt, ...);
#define LR_CUDA(call)                                                                        \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            lr_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return LR_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)
// same, running `cleanup` (free what the function has allocated so far) before returning
#define LR_CUDA_CLEAN(call, cleanup)                                                         \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            lr_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            cleanup;                                                                         \
            return LR_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)
#define LR_REQUIRE(cond, ...)                                                                \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            lr_set_error(__VA_ARGS__);                                                       \
            return LR_ERR_INVALID;                                                           \
        }                                                                                    \
    } while (0)

// staging buffers of the host-buffer entry points: enough in flight that the copy engine never waits for a K1 batch, even when
// another kernel (the chains of the previous table) leaves K1 only a few SMs
#define LR_NSTAGE 4

struct lr_handle_s {
    int device;
    int sm_count;
    int max_smem_optin;
    cudaStream_t stream;        // compute
    cudaStream_t copy_stream;   // host<->device staging
    cudaEvent_t ev[2 * LR_NSTAGE];   // [buf] copy of a batch done, [LR_NSTAGE + buf] its kernel done
    int64_t launches;           // kernels launched through this handle
    // grow-only workspace, handed from one entry point to the next IN STREAM ORDER (lr_ws_acquire)
    void* ws;
    size_t ws_bytes;
    cudaStream_t ws_stream;     // stream of the workspace's last user
    int ws_used;
    cudaEvent_t ev_order;       // scratch event of lr_order
    void* stage[LR_NSTAGE];
    size_t stage_bytes;
    // K1's memory of the last table's kind: one int in mapped pinned host memory that the kernel's CTA 0 writes (1 = the table
    // carries fractional times, 0 = integer years) and the NEXT lr_bin_accumulate reads without synchronising to choose between
    // its two bit-identical builds (k1_binstats.cu); a stale value costs speed only
    int* k1_hint;
    int k1_hint_used;
    // set by the host-buffer entry points around their per-batch passes: those are bound by the host link (K1 at 4.8 TB/s hides
    // behind 55 GB/s of copies) and run beside the chain kernels of the previous table, whose 64 KB shared-memory carve-out the
    // lane-private build (2 x 105 KB per SM) cannot share an SM with -- it would wait for whole SMs to drain
    int k1_general_only;
    int k1_last_build;          // what the last lr_bin_accumulate launched: 0 k1_bin_kernel, 1 k1_bin_lanes_kernel (lr_bin_last_build)
};

// Cross-stream ordering without host synchronisation.  Entry points take a caller stream or fall back to the handle's
// own; objects (datasets, chains, trend / dd samplers) and the workspace remember the stream that touched them last.
// lr_order makes `now` wait for everything submitted so far to `*last` (an event recorded there and waited for here) when
// the two differ, and moves `*last` to `now`.
struct lr_last_stream { cudaStream_t s; int valid; };
int lr_order(lr_handle_t h, lr_last_stream* last, cudaStream_t now);
// the workspace, at least `bytes` large, for work submitted to `stream` (ordered after its previous user)
int lr_ws_acquire(lr_handle_t h, size_t bytes, cudaStream_t stream);

// ---- small device helpers --------------------------------------------------------------------
__device__ __forceinline__ double2 ld_stream_f64x2(const double2* p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream_f64(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// warp_sum of a value that is zero in every lane >= K (the slots a side does not use), K warp-uniform: the butterfly steps that
// would only add zeros are skipped -- the bits are those of warp_sum -- and lane 0's total is broadcast.  One shuffle for a
// single-rate side instead of five.
__device__ __forceinline__ double warp_sum_first(double v, int K) {
    if (K > 8) return warp_sum(v);
    if (K > 1) {
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    }
    return __shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ double warp_prod(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- Philox-4x32-10 (Salmon et al., SC'11); counter-based, one call = 128 random bits --------
struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                          uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    Philox4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}
// uniform strictly inside (0,1) from 52 random bits: the bits become the mantissa of a double in [1,2) (no integer to
// floating-point conversion), shifted down by 1 and up by half a step: values (k + 1/2) 2^-52, k = 0 .. 2^52-1
__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
    const double x = __hiloint2double((int)(0x3ff00000u | (hi >> 12)), (int)((hi << 20) | (lo >> 12)));
    return (x - 1.0) + 1.1102230246251565e-16;
}

// exp(x) for |x| <= 0.1 (the log-multipliers of update_multiplier_proposal_vec: |2 log(1.1) (u - .5)| <= 0.0954): Taylor series to
// x^11, truncation error < 2e-21, i.e. rounding only.  The coefficients sit in constant memory: as literals every one costs
// two UMOV before its DFMA (17 extra instructions per call); from the constant bank two arrive per LDCU.128.
__constant__ double c_exp_small[12] = {2.505210838544172e-8 /* 1/11! */, 2.755731922398589e-7, 2.755731922398589e-6, 2.48015873015873e-5,
                                       1.984126984126984e-4, 1.388888888888889e-3, 8.333333333333333e-3, 4.166666666666666e-2,
                                       1.666666666666667e-1, 0.5, 1.0, 1.0};
__device__ __forceinline__ double exp_small(double x) {
    double p = c_exp_small[0];
#pragma unroll
    for (int k = 1; k < 12; ++k) p = fma(p, x, c_exp_small[k]);
    return p;
}

// Metropolis-Hastings test `x > log(u)` with the double-precision logarithm evaluated only when a single-precision bracket of
// log(u) cannot decide (error bound: __logf <= 2^-21.4 absolute on [.5, 2], 3 ulp elsewhere, plus the rounding of u to float);
// the decision is always the one the exact comparison gives.  NaN and -inf are rejected.
__device__ __forceinline__ bool mh_accept_gt(double x, double u) {
    const float lf = __logf((float)u);
    const float err = 2e-6f * (1.0f + fabsf(lf));
    if (x > (double)(lf + err)) return true;
    if (!(x >= (double)(lf - err))) return false;
    return x > log(u);
}

// standard normal from two uniforms in single precision (Box-Muller); a proposal step, not a likelihood term: 24 bits suffice
__device__ __forceinline__ double normal_f32(double u1, uint32_t bits) {
    const float r = sqrtf(-2.0f * __logf((float)u1));
    return (double)(r * cospif(((float)(bits >> 8) + 0.5f) * 1.1920929e-7f));     // angle 2 pi k / 2^24 as cospi(2 k / 2^24)
}
