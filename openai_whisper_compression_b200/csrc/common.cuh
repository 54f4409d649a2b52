// common.cuh -- shared helpers for libwhisperq (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <cstdlib>

#include "../../include/whisperq.h"

#ifndef WQ_SM_COUNT_FALLBACK
#define WQ_SM_COUNT_FALLBACK 148
#endif

// ----------------------------------------------------------------------------------------------
// error plumbing (thread-local message, no exceptions across the ABI)
// ----------------------------------------------------------------------------------------------
void wq_set_error(const char *fmt, ...);

#define WQ_REQUIRE(cond, ...)              \
    do {                                   \
        if (!(cond)) {                     \
            wq_set_error(__VA_ARGS__);     \
            return WQ_ERR_INVALID;         \
        }                                  \
    } while (0)

#define WQ_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t _e = (call);                                                        \
        if (_e != cudaSuccess) {                                                        \
            wq_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,           \
                         cudaGetErrorString(_e));                                       \
            return WQ_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define WQ_LAUNCH_CHECK()                                                               \
    do {                                                                                \
        cudaError_t _e = cudaGetLastError();                                            \
        if (_e != cudaSuccess) {                                                        \
            wq_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,       \
                         cudaGetErrorString(_e));                                       \
            return WQ_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

int wq_sm_count();          // cached SM count of the current device
int wq_check_device();      // WQ_OK when the current device is sm_100

static inline bool wq_aligned(const void *p, size_t a) {
    return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0;
}

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (PDL).  A decode step is ~80 short kernels back to back; with this launch
// attribute a kernel's CTAs are scheduled (launch latency, barrier/TMEM set-up, descriptor prefetch) while its
// stream predecessor is still draining.  Contract for every kernel launched through wq_launch_pdl: call pdl_wait()
// before the first access to global memory -- it blocks until the predecessor grid has completed and its writes
// are visible (a no-op without a programmatic predecessor) -- and pdl_trigger() once, AFTER the wait: the
// successor may then be scheduled.  Triggering only after the wait keeps the look-ahead at one kernel; triggering
// at kernel entry lets the whole downstream chain of a CUDA graph pile onto the SMs (each parked CTA holds
// registers, threads and shared memory) and cost the HBM-bound attention kernel a quarter of its bandwidth.
// Works under stream capture (the edge becomes a programmatic graph dependency).
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue_done() {
    pdl_wait();
    pdl_trigger();
}

template <typename... KArgs, typename... Args>
static inline cudaError_t wq_launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// the same launch with a thread-block cluster of `cluster_x` consecutive CTAs (tcgen05 cta_group::2 pairs)
template <typename... KArgs, typename... Args>
static inline cudaError_t wq_launch_pdl_cluster(unsigned cluster_x, void (*kern)(KArgs...), dim3 grid, dim3 block,
                                                size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = cluster_x;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

#define WQ_LAUNCH_PDL_CLUSTER(...)                                                      \
    do {                                                                                \
        cudaError_t _e = wq_launch_pdl_cluster(__VA_ARGS__);                            \
        if (_e != cudaSuccess) {                                                        \
            wq_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,       \
                         cudaGetErrorString(_e));                                       \
            return WQ_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define WQ_LAUNCH_PDL(...)                                                              \
    do {                                                                                \
        cudaError_t _e = wq_launch_pdl(__VA_ARGS__);                                    \
        if (_e != cudaSuccess) {                                                        \
            wq_set_error("kernel launch failed at %s:%d: %s", __FILE__, __LINE__,       \
                         cudaGetErrorString(_e));                                       \
            return WQ_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) {
    return __bfloat162float(v);
}

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
    return __float2bfloat16_rn(v);
}

// gelu(x) as torch computes it for approximate='none': x * 0.5 * (1 + erf(x / sqrt(2))) in fp32, rounded to T.
// One definition: the row kernels (rowops.cu), their lookup table, and the GEMM's outlier path that re-derives
// gelu(fc1 output) when the fp16 activation was never stored must agree bit for bit.
template <typename T> __device__ __forceinline__ T gelu_erf(float f) {
    return from_f32<T>(f * 0.5f * (1.0f + erff(f * 0.70710678118654752440f)));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i32(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Row scale of bitsandbytes' int8_vectorwise_quant.  Its kernel (csrc/kernels.cu, kInt8VectorQuant) forms 127 / absmax
// with the approximate-division intrinsic; over all fp16 (absmax, a) pairs that changes 8 734 of 503 856 639 codes
// against the IEEE quotient (a * scale landing on the other side of a .5 boundary; exhaustive sweep on B200,
// scripts/bnb_open_points.cu -> tests/golden/fdividef_127_fp16.npz, which lets the CPU oracle reproduce it), so
// the approximate form is what every LLM.int8 quantizer here uses.  absmax == 0 gives +inf -> NaN codes -> 0.
__device__ __forceinline__ float bnb_row_scale(float absmax) { return __fdividef(127.0f, absmax); }

// Order-preserving float <-> uint32 map (for atomicMin/atomicMax on floats).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// NF4 / FP4 code books (bitsandbytes csrc/kernels.cu value tables)
__device__ __constant__ const float kNF4Code[16] = {
    -1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
    -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
    0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f,
    0.33791524171829224f, 0.44070982933044434f, 0.5626170039176941f,
    0.7229568362236023f, 1.0f};
__device__ __constant__ const float kFP4Code[16] = {
    0.0f, 0.0052083333f, 0.6666667f, 1.0f, 0.33333334f, 0.5f, 0.16666667f, 0.25f,
    -0.0f, -0.0052083333f, -0.6666667f, -1.0f, -0.33333334f, -0.5f, -0.16666667f, -0.25f};
