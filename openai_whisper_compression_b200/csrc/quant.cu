// quant.cu -- weight packing and activation quantization kernels (HBM-bound, bit-exact).
//
// Every kernel here reads its input once from HBM (a second pass over the same warp-private
// row / block hits L1), so the algorithmic bytes are sizeof(T)*n read + packed bytes written.
// One warp owns one quantization group (a 64-element NF4 block or a matrix row), loads are
// 16-byte vectors where alignment allows, reductions are warp shuffles.
#include "common.cuh"

namespace {

constexpr int kWarpsPerCta = 8;

// ---------------------------------------------------------------------------------------------
// 4-bit encode (bitsandbytes csrc/kernels.cu dQuantizeNF4 / dQuantizeFP4): strict '>' tree.
// NaN (0 * inf for an all-zero block) fails every comparison and yields code 0.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t encode_nf4(float x) {
    if (x > 0.03979014977812767f) {
        if (x > 0.3893125355243683f) {
            if (x > 0.6427869200706482f) return (x > 0.8614784181118011f) ? 15u : 14u;
            return (x > 0.5016634166240692f) ? 13u : 12u;
        }
        if (x > 0.2035212516784668f) return (x > 0.2920137718319893f) ? 11u : 10u;
        return (x > 0.1202552504837513f) ? 9u : 8u;
    }
    if (x > -0.33967943489551544f) {
        if (x > -0.13791173323988914f) return (x > -0.045525018125772476f) ? 7u : 6u;
        return (x > -0.23460740596055984f) ? 5u : 4u;
    }
    if (x > -0.6106329262256622f) return (x > -0.4599952697753906f) ? 3u : 2u;
    return (x > -0.8480964004993439f) ? 1u : 0u;
}

__device__ __forceinline__ uint32_t encode_fp4(float x) {
    uint32_t sign = x < 0.0f ? 8u : 0u;
    x = fabsf(x);
    if (x > 0.29166667f) {
        if (x > 0.583333f) return (x > 0.8333333f ? 3u : 2u) + sign;
        return (x > 0.4166667f ? 5u : 4u) + sign;
    }
    if (x > 0.0859375f) return (x > 0.20833333f ? 7u : 6u) + sign;
    return (x > 0.00260417f ? 1u : 0u) + sign;
}

template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_quant_4bit(const T *__restrict__ w, int64_t n, int blocksize, int quant_type,
             uint8_t *__restrict__ packed, float *__restrict__ absmax, int64_t nblocks) {
    const int lane = threadIdx.x & 31;
    const int64_t blk = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (blk >= nblocks) return;
    const int64_t base = blk * blocksize;
    const int64_t end = min(base + (int64_t)blocksize, n);

    float am = 0.0f;
    for (int64_t i = base + lane * 2; i < end; i += 64) {
        float v0 = fabsf(to_f32(w[i]));
        float v1 = (i + 1 < end) ? fabsf(to_f32(w[i + 1])) : 0.0f;
        am = fmaxf(am, fmaxf(v0, v1));
    }
    am = warp_max(am);
    if (lane == 0) absmax[blk] = am;
    const float inv = __fdiv_rn(1.0f, am);

    for (int64_t i = base + lane * 2; i < end; i += 64) {
        float x0 = __fmul_rn(to_f32(w[i]), inv);
        float x1 = (i + 1 < n) ? __fmul_rn(to_f32(w[i + 1]), inv) : 0.0f;
        uint32_t c0 = quant_type ? encode_fp4(x0) : encode_nf4(x0);
        uint32_t c1 = quant_type ? encode_fp4(x1) : encode_nf4(x1);
        packed[i >> 1] = (uint8_t)((c0 << 4) | c1);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
k_dequant_4bit(const uint8_t *__restrict__ packed, const float *__restrict__ absmax, int64_t n,
               int blocksize_log2, int quant_type, T *__restrict__ out) {
    // one thread per packed byte pair group of 4 bytes (8 outputs)
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t e0 = g * 8;
    if (e0 >= n) return;
    const float *code = quant_type ? kFP4Code : kNF4Code;
    const float am = absmax[e0 >> blocksize_log2];
    if (e0 + 8 <= n) {
        const uint32_t p = *reinterpret_cast<const uint32_t *>(packed + (e0 >> 1));
        T v[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t byte = (p >> (8 * j)) & 0xffu;
            v[2 * j] = from_f32<T>(__fmul_rn(code[byte >> 4], am));
            v[2 * j + 1] = from_f32<T>(__fmul_rn(code[byte & 15u], am));
        }
        if (sizeof(T) == 2) {
            *reinterpret_cast<uint4 *>(out + e0) = *reinterpret_cast<uint4 *>(v);
        } else {
            reinterpret_cast<uint4 *>(out + e0)[0] = reinterpret_cast<uint4 *>(v)[0];
            reinterpret_cast<uint4 *>(out + e0)[1] = reinterpret_cast<uint4 *>(v)[1];
        }
    } else {
        for (int64_t i = e0; i < n; ++i) {
            const uint32_t byte = packed[i >> 1];
            const uint32_t c = (i & 1) ? (byte & 15u) : (byte >> 4);
            out[i] = from_f32<T>(__fmul_rn(code[c], am));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// bitsandbytes int8_vectorwise_quant: one warp per row of an fp16 matrix.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_quant_i8_rowwise_bnb(const __half *__restrict__ a, int64_t rows, int64_t cols, float threshold,
                       int8_t *__restrict__ out, float *__restrict__ row_stats,
                       int32_t *__restrict__ col_flags, int vec_ok) {
    pdl_prologue_done();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= rows) return;
    const __half *pr = a + row * cols;
    int8_t *po = out + row * cols;
    const bool sparse = threshold > 0.0f;

    float am = 0.0f;
    if (vec_ok) {
        for (int64_t c = lane * 8; c < cols; c += 256) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(pr + c);
            const __half2 *h = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                const float x = fabsf(f.x), y = fabsf(f.y);
                if (!sparse || x < threshold) am = fmaxf(am, x);
                if (!sparse || y < threshold) am = fmaxf(am, y);
            }
        }
    } else {
        for (int64_t c = lane; c < cols; c += 32) {
            const float x = fabsf(__half2float(pr[c]));
            if (!sparse || x < threshold) am = fmaxf(am, x);
        }
    }
    am = warp_max(am);
    if (lane == 0) row_stats[row] = am;
    const float scale = bnb_row_scale(am);      // __fdividef(127, absmax), as bitsandbytes ships it (common.cuh)

    if (vec_ok) {
        for (int64_t c = lane * 8; c < cols; c += 256) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(pr + c);
            const __half *h = reinterpret_cast<const __half *>(&raw);
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float v = __half2float(h[j]);
                int q;
                if (sparse && !(fabsf(v) < threshold)) {
                    q = 0;
                    col_flags[c + j] = 1;
                    col_flags[cols] = 1;   // "any outlier in this call"
                } else {
                    q = __float2int_rn(__fmul_rn(v, scale));  // NaN -> 0
                }
                const uint32_t b = (uint32_t)(q & 0xff);
                if (j < 4) lo |= b << (8 * j); else hi |= b << (8 * (j - 4));
            }
            *reinterpret_cast<uint2 *>(po + c) = make_uint2(lo, hi);
        }
    } else {
        for (int64_t c = lane; c < cols; c += 32) {
            const float v = __half2float(pr[c]);
            int q;
            if (sparse && !(fabsf(v) < threshold)) {
                q = 0;
                col_flags[c] = 1;
                col_flags[cols] = 1;
            } else {
                q = __float2int_rn(__fmul_rn(v, scale));
            }
            po[c] = (int8_t)q;
        }
    }
}

// compaction of the outlier column flags (single CTA) -------------------------------------------
__global__ void __launch_bounds__(1024)
k_outlier_compact(int32_t *__restrict__ col_flags, int cols, int32_t *__restrict__ outlier_cols,
                  int32_t *__restrict__ n_outliers) {
    __shared__ int s_warp[32];
    __shared__ int s_total;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int per = (cols + 1023) / 1024;
    const int c0 = tid * per, c1 = min(c0 + per, cols);
    int cnt = 0;
    for (int c = c0; c < c1; ++c) cnt += col_flags[c] != 0;
    // inclusive scan inside the warp, then across warps
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int v = s_warp[lane], w = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        s_warp[lane] = w - v;  // exclusive
        if (lane == 31) s_total = w;
    }
    __syncthreads();
    int pos = s_warp[wid] + inc - cnt;
    for (int c = c0; c < c1; ++c) {
        if (col_flags[c] != 0) {
            outlier_cols[pos++] = c;
            col_flags[c] = 0;
        }
    }
    if (tid == 0) {
        *n_outliers = s_total;
        col_flags[cols] = 0;       // the "any outlier" word
        col_flags[cols + 1] = 0;   // completion counter used by the fused GEMM
    }
}

__global__ void __launch_bounds__(256)
k_outlier_zero_cols(int8_t *__restrict__ ca, int64_t rows, int64_t cols,
                    const int32_t *__restrict__ outlier_cols,
                    const int32_t *__restrict__ n_outliers) {
    const int n = *n_outliers;
    if (n == 0) return;
    const int64_t total = rows * (int64_t)n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n;
        const int j = (int)(i - r * n);
        ca[r * cols + outlier_cols[j]] = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// quanto qint8 per-output-channel (warp per row)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_quant_i8_rowwise_quanto(const T *__restrict__ w, int64_t N, int64_t K, int8_t *__restrict__ q,
                          float *__restrict__ scale) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= N) return;
    const T *pr = w + row * K;
    float am = 0.0f;
    for (int64_t c = lane; c < K; c += 32) am = fmaxf(am, fabsf(to_f32(pr[c])));
    am = warp_max(am);
    // torch evaluates `absmax / 127` and `w / scale` on half / bfloat16 tensors in fp32 and rounds each result to the
    // tensor dtype (opmath); for fp32 weights the two roundings below are no-ops
    const float s = to_f32(from_f32<T>(__fdiv_rn(am, 127.0f)));
    if (lane == 0) scale[row] = s;
    for (int64_t c = lane; c < K; c += 32) {
        float r = rintf(to_f32(from_f32<T>(__fdiv_rn(to_f32(pr[c]), s))));
        if (r != r) r = 0.0f;
        r = fminf(fmaxf(r, -128.0f), 127.0f);
        q[row * K + c] = (int8_t)(int)r;
    }
}

// ---------------------------------------------------------------------------------------------
// bitsandbytes nested ("double") quantization of the 4-bit absmax statistics
// (quantize_4bit(compress_statistics=True)): offset = mean(absmax); absmax - offset is quantized
// block-wise (256) to 8 bits against the 256-entry "dynamic" code book with the library's binary
// search + nearest-neighbour rule (csrc/kernels.cu dQuantize<0>).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dquantize8(const float *code, float x) {
    int pivot = 127, upper_pivot = 255, lower_pivot = 0;
    float lower = -1.0f, upper = 1.0f;
    float val = code[pivot];
    for (int i = 64; i > 0; i >>= 1) {
        if (x > val) {
            lower_pivot = pivot;
            lower = val;
            pivot += i;
        } else {
            upper_pivot = pivot;
            upper = val;
            pivot -= i;
        }
        val = code[pivot];
    }
    if (upper_pivot == 255) upper = code[upper_pivot];
    if (lower_pivot == 0) lower = code[lower_pivot];
    if (x > val) {
        const float midpoint = (upper + val) * 0.5f;
        return x > midpoint ? upper_pivot : pivot;
    }
    const float midpoint = (lower + val) * 0.5f;
    return x < midpoint ? lower_pivot : pivot;
}

// single CTA: offset = mean (double accumulation, one rounding to fp32)
__global__ void __launch_bounds__(1024)
k_absmax_mean(const float *__restrict__ absmax, int64_t n, float *__restrict__ offset) {
    __shared__ double s_sum[32];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 1024) acc += (double)absmax[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = s_sum[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) *offset = (float)(acc / (double)n);
    }
}

// one warp per 256-element block of (absmax - offset)
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_quant_absmax_blockwise8(const float *__restrict__ absmax, int64_t n, const float *__restrict__ offset,
                          const float *__restrict__ code, uint8_t *__restrict__ q, float *__restrict__ absmax2,
                          float *__restrict__ absmax_deq) {
    __shared__ float s_code[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_code[i] = code[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t blk = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    const int64_t base = blk * 256;
    if (base >= n) return;
    const float off = *offset;
    float v[8];
    float am = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int64_t i = base + lane + 32 * j;
        v[j] = i < n ? __fsub_rn(absmax[i], off) : 0.0f;
        am = fmaxf(am, fabsf(v[j]));
    }
    am = warp_max(am);
    if (lane == 0) absmax2[blk] = am;
    const float inv = __fdiv_rn(1.0f, am);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int64_t i = base + lane + 32 * j;
        if (i < n) {
            const uint32_t c = dquantize8(s_code, __fmul_rn(v[j], inv));
            q[i] = (uint8_t)c;
            absmax_deq[i] = __fadd_rn(__fmul_rn(s_code[c], am), off);   // what dequantize_4bit will see
        }
    }
}

__global__ void __launch_bounds__(256)
k_dequant_absmax_blockwise8(const uint8_t *__restrict__ q, const float *__restrict__ absmax2,
                            const float *__restrict__ code, const float *__restrict__ offset, int64_t n,
                            float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = __fadd_rn(__fmul_rn(__ldg(code + q[i]), absmax2[i >> 8]), *offset);
}

// ---------------------------------------------------------------------------------------------
// quanto qint4: group-wise affine uint4 (MaxOptimizer + AffineQuantizer), one warp per group
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_quant_u4_group_quanto(const T *__restrict__ w, int64_t n_groups, int group, float qmax,
                        uint8_t *__restrict__ packed, float *__restrict__ scale, float *__restrict__ shift) {
    const int lane = threadIdx.x & 31;
    const int64_t g = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (g >= n_groups) return;
    const T *pw = w + g * group;
    float mn = INFINITY, mx = -INFINITY;
    for (int i = lane; i < group; i += 32) {
        const float v = to_f32(pw[i]);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    mn = warp_min(mn);
    mx = warp_max(mx);
    const float s = __fdiv_rn(__fsub_rn(mx, mn), qmax);    // (rmax - rmin) / (qmax - qmin): 15 (qint4) or 3 (qint2)
    const float sh = -mn;                                   // shift = -rmin
    if (lane == 0) {
        scale[g] = s;
        shift[g] = sh;
    }
    // q = clamp(round((w + shift) / scale), 0, qmax); two codes per byte, first in the high nibble
    for (int i = lane * 2; i < group; i += 64) {
        uint32_t c[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float r = rintf(__fdiv_rn(__fadd_rn(to_f32(pw[i + j]), sh), s));
            if (r != r) r = 0.0f;                           // constant group: 0 / 0
            r = fminf(fmaxf(r, 0.0f), qmax);
            c[j] = (uint32_t)(int)r;
        }
        packed[(g * group + i) >> 1] = (uint8_t)((c[0] << 4) | c[1]);
    }
}

// ---------------------------------------------------------------------------------------------
// torch dynamic int8: per-tensor symmetric weights
// ---------------------------------------------------------------------------------------------
__global__ void k_minmax_init(uint32_t *ws) {
    ws[0] = 0u;  // max of ordered(x)
    ws[1] = 0u;  // max of ordered(-x)
}

template <typename T>
__global__ void __launch_bounds__(256)
k_minmax(const T *__restrict__ x, int64_t n, uint32_t *__restrict__ ws) {
    float mx = -INFINITY, mn = INFINITY;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float v = to_f32(x[i]);
        mx = fmaxf(mx, v);
        mn = fminf(mn, v);
    }
    mx = warp_max(mx);
    mn = warp_min(mn);
    __shared__ float s_mx[8], s_mn[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_mx[wid] = mx; s_mn[wid] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { mx = fmaxf(mx, s_mx[i]); mn = fminf(mn, s_mn[i]); }
        if (mx >= mn) {  // at least one element seen by this CTA
            atomicMax(&ws[0], float_to_ordered(mx));
            atomicMax(&ws[1], float_to_ordered(-mn));
        }
    }
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_quant_i8_tensor_torch(const float *__restrict__ w, int64_t N, int64_t K,
                        const uint32_t *__restrict__ ws, int8_t *__restrict__ q,
                        float *__restrict__ scale_out, int32_t *__restrict__ wsum) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    // MinMaxObserver(per_tensor_symmetric, qint8): torch/ao/quantization/observer.py
    const float mx = ws[0] ? ordered_to_float(ws[0]) : 0.0f;
    const float mn = ws[1] ? -ordered_to_float(ws[1]) : 0.0f;
    const float min_neg = fminf(mn, 0.0f), max_pos = fmaxf(mx, 0.0f);
    float scale = __fdiv_rn(fmaxf(-min_neg, max_pos), 127.5f);
    scale = fmaxf(scale, 1.1920928955078125e-07f);
    if (blockIdx.x == 0 && threadIdx.x == 0) *scale_out = scale;
    if (row >= N) return;
    const float inv = __fdiv_rn(1.0f, scale);
    int sum = 0;
    for (int64_t c = lane; c < K; c += 32) {
        float r = nearbyintf(__fmul_rn(w[row * K + c], inv));
        r = fminf(fmaxf(r, -128.0f), 127.0f);
        const int v = (int)r;
        q[row * K + c] = (int8_t)v;
        sum += v;
    }
    sum = warp_sum_i32(sum);
    if (lane == 0) wsum[row] = sum;
}

// ---------------------------------------------------------------------------------------------
// torch dynamic int8: per-tensor affine uint8 activations (FBGEMM ChooseQuantizationParams)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void choose_qparams_u8(float mn, float mx, float &scale, int &zp) {
    const int qmin = 0, qmax = 127;  // reduce_range
    mn = fminf(mn, 0.0f);
    mx = fmaxf(mx, 0.0f);
    scale = (float)(((double)mx - (double)mn) / (double)(qmax - qmin));
    if (scale == 0.0f || isinf(__fdiv_rn(1.0f, scale))) scale = 0.1f;
    const double zmin = qmin - (double)mn / (double)scale;
    const double zmax = qmax - (double)mx / (double)scale;
    const double emin = fabs((double)qmin) + fabs((double)mn / (double)scale);
    const double emax = fabs((double)qmax) + fabs((double)mx / (double)scale);
    const double init = emin < emax ? zmin : zmax;
    if (init < qmin) zp = qmin;
    else if (init > qmax) zp = qmax;
    else zp = (int)nearbyint(init);
}

template <typename T>
__global__ void __launch_bounds__(256)
k_quant_act_u8(const T *__restrict__ x, int64_t n, const uint32_t *__restrict__ ws,
               uint8_t *__restrict__ q, float *__restrict__ qparams) {
    float scale;
    int zp;
    choose_qparams_u8(ws[1] ? -ordered_to_float(ws[1]) : 0.0f, ws[0] ? ordered_to_float(ws[0]) : 0.0f,
                      scale, zp);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        qparams[0] = scale;
        qparams[1] = (float)zp;
    }
    const float inv = __fdiv_rn(1.0f, scale);
    const float fzp = (float)zp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float r = __fadd_rn(nearbyintf(__fmul_rn(to_f32(x[i]), inv)), fzp);
        r = fminf(fmaxf(r, 0.0f), 255.0f);
        q[i] = (uint8_t)(int)r;
    }
}

int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return l;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" int wq_quant_4bit(const void *w, int w_dtype, int64_t n, int blocksize, int quant_type,
                             uint8_t *packed, float *absmax, wq_stream_t stream) {
    WQ_REQUIRE(n >= 0, "wq_quant_4bit: n < 0");
    WQ_REQUIRE(blocksize >= 64 && blocksize <= 4096 && (blocksize & (blocksize - 1)) == 0,
               "wq_quant_4bit: blocksize %d must be a power of two in [64, 4096]", blocksize);
    WQ_REQUIRE(quant_type == WQ_NF4 || quant_type == WQ_FP4, "wq_quant_4bit: bad quant_type");
    if (n == 0) return WQ_OK;
    WQ_REQUIRE(w && packed && absmax, "wq_quant_4bit: null pointer");
    const int64_t nblocks = (n + blocksize - 1) / blocksize;
    const unsigned grid = (unsigned)((nblocks + kWarpsPerCta - 1) / kWarpsPerCta);
    cudaStream_t s = (cudaStream_t)stream;
    switch (w_dtype) {
        case WQ_F32:
            k_quant_4bit<float><<<grid, kWarpsPerCta * 32, 0, s>>>((const float *)w, n, blocksize,
                                                                    quant_type, packed, absmax, nblocks);
            break;
        case WQ_F16:
            k_quant_4bit<__half><<<grid, kWarpsPerCta * 32, 0, s>>>((const __half *)w, n, blocksize,
                                                                     quant_type, packed, absmax, nblocks);
            break;
        case WQ_BF16:
            k_quant_4bit<__nv_bfloat16><<<grid, kWarpsPerCta * 32, 0, s>>>(
                (const __nv_bfloat16 *)w, n, blocksize, quant_type, packed, absmax, nblocks);
            break;
        default:
            WQ_REQUIRE(false, "wq_quant_4bit: bad dtype %d", w_dtype);
    }
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}

extern "C" int wq_dequant_4bit(const uint8_t *packed, const float *absmax, int64_t n, int blocksize,
                               int quant_type, void *out, int out_dtype, wq_stream_t stream) {
    WQ_REQUIRE(n >= 0, "wq_dequant_4bit: n < 0");
    WQ_REQUIRE(blocksize >= 64 && blocksize <= 4096 && (blocksize & (blocksize - 1)) == 0,
               "wq_dequant_4bit: blocksize %d must be a power of two in [64, 4096]", blocksize);
    WQ_REQUIRE(quant_type == WQ_NF4 || quant_type == WQ_FP4, "wq_dequant_4bit: bad quant_type");
    if (n == 0) return WQ_OK;
    WQ_REQUIRE(packed && absmax && out, "wq_dequant_4bit: null pointer");
    WQ_REQUIRE(wq_aligned(packed, 4) && wq_aligned(out, 16), "wq_dequant_4bit: misaligned buffer");
    const int64_t groups = (n + 7) / 8;
    const unsigned grid = (unsigned)((groups + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    const int lg = ilog2(blocksize);
    switch (out_dtype) {
        case WQ_F32:
            k_dequant_4bit<float><<<grid, 256, 0, s>>>(packed, absmax, n, lg, quant_type, (float *)out);
            break;
        case WQ_F16:
            k_dequant_4bit<__half><<<grid, 256, 0, s>>>(packed, absmax, n, lg, quant_type, (__half *)out);
            break;
        case WQ_BF16:
            k_dequant_4bit<__nv_bfloat16><<<grid, 256, 0, s>>>(packed, absmax, n, lg, quant_type,
                                                               (__nv_bfloat16 *)out);
            break;
        default:
            WQ_REQUIRE(false, "wq_dequant_4bit: bad dtype %d", out_dtype);
    }
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}

extern "C" int wq_quant_i8_rowwise_bnb(const void *a_f16, int64_t rows, int64_t cols, float threshold,
                                       int8_t *out, float *row_stats, int32_t *col_flags,
                                       wq_stream_t stream) {
    WQ_REQUIRE(rows >= 0 && cols >= 0, "wq_quant_i8_rowwise_bnb: negative shape");
    WQ_REQUIRE(threshold >= 0.0f, "wq_quant_i8_rowwise_bnb: negative threshold");
    if (rows == 0 || cols == 0) return WQ_OK;
    WQ_REQUIRE(a_f16 && out && row_stats, "wq_quant_i8_rowwise_bnb: null pointer");
    WQ_REQUIRE(threshold == 0.0f || col_flags, "wq_quant_i8_rowwise_bnb: threshold needs col_flags");
    const int vec_ok = (cols % 8 == 0) && wq_aligned(a_f16, 16) && wq_aligned(out, 8);
    const unsigned grid = (unsigned)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
    WQ_LAUNCH_PDL(k_quant_i8_rowwise_bnb, dim3(grid), dim3(kWarpsPerCta * 32), 0, (cudaStream_t)stream,
                  (const __half *)a_f16, rows, cols, threshold, out, row_stats, col_flags, vec_ok);
    return WQ_OK;
}

extern "C" int wq_outlier_columns(int32_t *col_flags, int64_t rows, int64_t cols, int8_t *ca,
                                  int32_t *outlier_cols, int32_t *n_outliers, wq_stream_t stream) {
    WQ_REQUIRE(rows >= 0 && cols >= 0 && cols < (1 << 30), "wq_outlier_columns: bad shape");
    WQ_REQUIRE(col_flags && outlier_cols && n_outliers, "wq_outlier_columns: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    k_outlier_compact<<<1, 1024, 0, s>>>(col_flags, (int)cols, outlier_cols, n_outliers);
    WQ_LAUNCH_CHECK();
    if (rows > 0 && ca) {
        const int grid = wq_sm_count() * 4;
        k_outlier_zero_cols<<<grid, 256, 0, s>>>(ca, rows, cols, outlier_cols, n_outliers);
        WQ_LAUNCH_CHECK();
    }
    return WQ_OK;
}

extern "C" int wq_quant_i8_rowwise_quanto(const void *w, int w_dtype, int64_t N, int64_t K, int8_t *q,
                                          float *scale, wq_stream_t stream) {
    WQ_REQUIRE(N >= 0 && K >= 0, "wq_quant_i8_rowwise_quanto: negative shape");
    if (N == 0) return WQ_OK;
    WQ_REQUIRE(w && q && scale, "wq_quant_i8_rowwise_quanto: null pointer");
    const unsigned grid = (unsigned)((N + kWarpsPerCta - 1) / kWarpsPerCta);
    cudaStream_t s = (cudaStream_t)stream;
    switch (w_dtype) {
        case WQ_F32:
            k_quant_i8_rowwise_quanto<float><<<grid, kWarpsPerCta * 32, 0, s>>>((const float *)w, N, K, q, scale);
            break;
        case WQ_F16:
            k_quant_i8_rowwise_quanto<__half><<<grid, kWarpsPerCta * 32, 0, s>>>((const __half *)w, N, K, q, scale);
            break;
        case WQ_BF16:
            k_quant_i8_rowwise_quanto<__nv_bfloat16><<<grid, kWarpsPerCta * 32, 0, s>>>(
                (const __nv_bfloat16 *)w, N, K, q, scale);
            break;
        default:
            WQ_REQUIRE(false, "wq_quant_i8_rowwise_quanto: bad dtype %d", w_dtype);
    }
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}

extern "C" int wq_quant_i8_tensor_torch(const float *w, int64_t N, int64_t K, int8_t *q, float *scale,
                                        int32_t *wsum, float *workspace, wq_stream_t stream) {
    WQ_REQUIRE(N >= 0 && K >= 0, "wq_quant_i8_tensor_torch: negative shape");
    WQ_REQUIRE(scale && workspace, "wq_quant_i8_tensor_torch: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t *ws = reinterpret_cast<uint32_t *>(workspace);
    const int64_t n = N * K;
    k_minmax_init<<<1, 1, 0, s>>>(ws);
    if (n > 0) {
        WQ_REQUIRE(w && q && wsum, "wq_quant_i8_tensor_torch: null pointer");
        const int grid = (int)min((int64_t)wq_sm_count() * 8, (n + 255) / 256);
        k_minmax<float><<<grid, 256, 0, s>>>(w, n, ws);
    }
    const unsigned grid2 = (unsigned)max((int64_t)1, (N + kWarpsPerCta - 1) / kWarpsPerCta);
    k_quant_i8_tensor_torch<<<grid2, kWarpsPerCta * 32, 0, s>>>(w, N, K, ws, q, scale, wsum);
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}

extern "C" int wq_quant_act_u8_tensor(const void *x, int x_dtype, int64_t n, uint8_t *q, float *qparams,
                                      uint32_t *workspace, wq_stream_t stream) {
    WQ_REQUIRE(n >= 0, "wq_quant_act_u8_tensor: n < 0");
    WQ_REQUIRE(qparams && workspace, "wq_quant_act_u8_tensor: null pointer");
    WQ_REQUIRE(x_dtype == WQ_F32 || x_dtype == WQ_F16, "wq_quant_act_u8_tensor: dtype must be F32/F16");
    cudaStream_t s = (cudaStream_t)stream;
    k_minmax_init<<<1, 1, 0, s>>>(workspace);
    const int grid = (int)max((int64_t)1, min((int64_t)wq_sm_count() * 8, (n + 255) / 256));
    if (x_dtype == WQ_F32) {
        if (n > 0) k_minmax<float><<<grid, 256, 0, s>>>((const float *)x, n, workspace);
        k_quant_act_u8<float><<<grid, 256, 0, s>>>((const float *)x, n, workspace, q, qparams);
    } else {
        if (n > 0) k_minmax<__half><<<grid, 256, 0, s>>>((const __half *)x, n, workspace);
        k_quant_act_u8<__half><<<grid, 256, 0, s>>>((const __half *)x, n, workspace, q, qparams);
    }
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}

extern "C" int wq_quant_u4_group_quanto(const void *w, int w_dtype, int64_t N, int64_t K, int group, uint8_t *packed,
                                        float *scale, float *shift, wq_stream_t stream) {
    return wq_quant_ubits_group_quanto(w, w_dtype, N, K, group, 4, packed, scale, shift, stream);
}

extern "C" int wq_quant_ubits_group_quanto(const void *w, int w_dtype, int64_t N, int64_t K, int group, int bits,
                                           uint8_t *packed, float *scale, float *shift, wq_stream_t stream) {
    WQ_REQUIRE(bits == 2 || bits == 4, "wq_quant_ubits_group_quanto: bits must be 2 or 4 (got %d)", bits);
    const float qmax = (float)((1 << bits) - 1);
    WQ_REQUIRE(N >= 0 && K >= 0, "wq_quant_u4_group_quanto: negative shape");
    WQ_REQUIRE(group >= 2 && group % 2 == 0 && (K == 0 || K % group == 0),
               "wq_quant_u4_group_quanto: group %d must be even and divide K=%lld", group, (long long)K);
    if (N == 0 || K == 0) return WQ_OK;
    WQ_REQUIRE(w && packed && scale && shift, "wq_quant_u4_group_quanto: null pointer");
    const int64_t n_groups = N * (K / group);
    const unsigned grid = (unsigned)((n_groups + kWarpsPerCta - 1) / kWarpsPerCta);
    cudaStream_t s = (cudaStream_t)stream;
    switch (w_dtype) {
        case WQ_F32:
            k_quant_u4_group_quanto<float><<<grid, kWarpsPerCta * 32, 0, s>>>((const float *)w, n_groups, group, qmax, packed, scale, shift);
            break;
        case WQ_F16:
            k_quant_u4_group_quanto<__half><<<grid, kWarpsPerCta * 32, 0, s>>>((const __half *)w, n_groups, group, qmax, packed, scale, shift);
            break;
        case WQ_BF16:
            k_quant_u4_group_quanto<__nv_bfloat16><<<grid, kWarpsPerCta * 32, 0, s>>>((const __nv_bfloat16 *)w, n_groups, group, qmax, packed, scale, shift);
            break;
        default:
            WQ_REQUIRE(false, "wq_quant_u4_group_quanto: bad dtype %d", w_dtype);
    }
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}

extern "C" int wq_quant_absmax_double(const float *absmax, int64_t n, const float *code256, uint8_t *q,
                                      float *absmax2, float *offset, float *absmax_deq, wq_stream_t stream) {
    WQ_REQUIRE(n >= 0, "wq_quant_absmax_double: n < 0");
    if (n == 0) return WQ_OK;
    WQ_REQUIRE(absmax && code256 && q && absmax2 && offset && absmax_deq, "wq_quant_absmax_double: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    k_absmax_mean<<<1, 1024, 0, s>>>(absmax, n, offset);
    const int64_t nblocks = (n + 255) / 256;
    k_quant_absmax_blockwise8<<<(unsigned)((nblocks + kWarpsPerCta - 1) / kWarpsPerCta), kWarpsPerCta * 32, 0, s>>>(
        absmax, n, offset, code256, q, absmax2, absmax_deq);
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}

extern "C" int wq_dequant_absmax_double(const uint8_t *q, const float *absmax2, const float *code256,
                                        const float *offset, int64_t n, float *absmax_out, wq_stream_t stream) {
    WQ_REQUIRE(n >= 0, "wq_dequant_absmax_double: n < 0");
    if (n == 0) return WQ_OK;
    WQ_REQUIRE(q && absmax2 && code256 && offset && absmax_out, "wq_dequant_absmax_double: null pointer");
    k_dequant_absmax_blockwise8<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(q, absmax2, code256,
                                                                                              offset, n, absmax_out);
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}

// ---------------------------------------------------------------------------------------------
// optimum-quanto float8 weights and static (calibrated) activation quantization -- round 2
// (model_utils.py:152-214 apply_static_quantization; quantization.py:53-86 "quanto_*_static*" / float8 configs).
// ---------------------------------------------------------------------------------------------
#include <cuda_fp8.h>

namespace {

// float -> e4m3fn code, round to nearest even, as torch's `.to(torch.float8_e4m3fn)` (quanto's SymmetricQuantizer for
// float qtypes: `(base / scale).to(qtype.dtype)`).  Values beyond +-448 do not occur (scale = absmax / 448) except
// through rounding of the quotient, where the conversion saturates to +-448 like the saturating cast.
__device__ __forceinline__ uint8_t f32_to_e4m3(float x) {
    return (uint8_t)__nv_cvt_float_to_fp8(x, __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ float e4m3_to_f32(uint8_t c) {
    const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)c, __NV_E4M3);
    return __half2float(*reinterpret_cast<const __half *>(&h));
}

// weights = qfloat8 (AbsmaxOptimizer, axis 0): scale[n] = max|W[n,:]| / 448, q = e4m3(W / scale)
template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_quant_f8_rowwise_quanto(const T *__restrict__ w, int64_t N, int64_t K, uint8_t *__restrict__ q,
                          float *__restrict__ scale) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
    if (row >= N) return;
    const T *pr = w + row * K;
    float am = 0.0f;
    for (int64_t c = lane; c < K; c += 32) am = fmaxf(am, fabsf(to_f32(pr[c])));
    am = warp_max(am);
    const float s = to_f32(from_f32<T>(__fdiv_rn(am, 448.0f)));       // rounded to the weight dtype (torch opmath)
    if (lane == 0) scale[row] = s;
    for (int64_t c = lane; c < K; c += 32) {
        float r = to_f32(from_f32<T>(__fdiv_rn(to_f32(pr[c]), s)));
        if (r != r) r = 0.0f;                       // all-zero row: 0 / 0
        q[row * K + c] = f32_to_e4m3(r);
    }
}

// Static activation quantization with a calibrated per-tensor scale (device scalar):
//   qtype 0 (qint8):   code = clamp(rint(x / s), -128, 127)
//   qtype 1 (qfloat8): code = e4m3(x / s)
// Outputs (each optional): codes_i8 (int8 codes; qint8 only -- the A operand of the int8 x int8 GEMM), grid_f16 (the
// code VALUE as fp16, exact for both types -- the A operand of the weight-expanding GEMMs), deq (code * s in the
// dtype of x: what quanto's ActivationQBytesTensor.dequantize() hands to the next float op).
template <typename T>
__global__ void __launch_bounds__(256)
k_quant_act_static(const T *__restrict__ x, int64_t n, const float *__restrict__ scale_ptr, int qtype,
                   int8_t *__restrict__ codes_i8, __half *__restrict__ grid_f16, T *__restrict__ deq) {
    pdl_prologue_done();
    const float s = *scale_ptr;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = to_f32(from_f32<T>(__fdiv_rn(to_f32(x[i]), s)));    // `x / scale` rounded to the dtype of x
        float g;
        if (qtype == 0) {
            g = rintf(v);
            if (g != g) g = 0.0f;
            g = fminf(fmaxf(g, -128.0f), 127.0f);
            if (codes_i8 != nullptr) codes_i8[i] = (int8_t)(int)g;
        } else {
            g = e4m3_to_f32(f32_to_e4m3(v != v ? 0.0f : v));
        }
        if (grid_f16 != nullptr) grid_f16[i] = __float2half_rn(g);
        if (deq != nullptr) deq[i] = from_f32<T>(__fmul_rn(g, s));
    }
}

}  // namespace

extern "C" int wq_quant_f8_rowwise_quanto(const void *w, int w_dtype, int64_t N, int64_t K, uint8_t *q, float *scale,
                                          wq_stream_t stream) {
    WQ_REQUIRE(N >= 0 && K >= 0, "wq_quant_f8_rowwise_quanto: negative shape");
    if (N == 0) return WQ_OK;
    WQ_REQUIRE(w && q && scale, "wq_quant_f8_rowwise_quanto: null pointer");
    const unsigned grid = (unsigned)((N + kWarpsPerCta - 1) / kWarpsPerCta);
    cudaStream_t s = (cudaStream_t)stream;
    switch (w_dtype) {
        case WQ_F32: k_quant_f8_rowwise_quanto<float><<<grid, kWarpsPerCta * 32, 0, s>>>((const float *)w, N, K, q, scale); break;
        case WQ_F16: k_quant_f8_rowwise_quanto<__half><<<grid, kWarpsPerCta * 32, 0, s>>>((const __half *)w, N, K, q, scale); break;
        case WQ_BF16:
            k_quant_f8_rowwise_quanto<__nv_bfloat16><<<grid, kWarpsPerCta * 32, 0, s>>>((const __nv_bfloat16 *)w, N, K, q, scale);
            break;
        default: WQ_REQUIRE(false, "wq_quant_f8_rowwise_quanto: bad dtype %d", w_dtype);
    }
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}

extern "C" int wq_quant_act_static(const void *x, int x_dtype, int64_t n, const float *scale, int qtype,
                                   int8_t *codes_i8, void *grid_f16, void *deq, wq_stream_t stream) {
    WQ_REQUIRE(n >= 0, "wq_quant_act_static: n < 0");
    WQ_REQUIRE(qtype == 0 || qtype == 1, "wq_quant_act_static: qtype must be 0 (qint8) or 1 (qfloat8 e4m3)");
    WQ_REQUIRE(qtype == 0 || codes_i8 == nullptr, "wq_quant_act_static: int8 codes exist for qint8 only");
    if (n == 0) return WQ_OK;
    WQ_REQUIRE(x && scale, "wq_quant_act_static: null pointer");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t want = (n + 255) / 256;
    const unsigned grid = (unsigned)(want < (int64_t)wq_sm_count() * 16 ? want : (int64_t)wq_sm_count() * 16);
    switch (x_dtype) {
        case WQ_F32:
            WQ_LAUNCH_PDL(k_quant_act_static<float>, dim3(grid), dim3(256), 0, s, (const float *)x, n, scale, qtype, codes_i8,
                          (__half *)grid_f16, (float *)deq);
            break;
        case WQ_F16:
            WQ_LAUNCH_PDL(k_quant_act_static<__half>, dim3(grid), dim3(256), 0, s, (const __half *)x, n, scale, qtype, codes_i8,
                          (__half *)grid_f16, (__half *)deq);
            break;
        case WQ_BF16:
            WQ_LAUNCH_PDL(k_quant_act_static<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, (const __nv_bfloat16 *)x, n, scale, qtype,
                          codes_i8, (__half *)grid_f16, (__nv_bfloat16 *)deq);
            break;
        default: WQ_REQUIRE(false, "wq_quant_act_static: bad dtype %d", x_dtype);
    }
    return WQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Sparse checkpoint -> dense tensor on the device.  The reference stores pruned models
// as per-tensor (indices, values) pairs (pruning/final_pruning_script/global_storing_as sparse.py:287-407: flat int64
// indices + values in a zip) or as torch COO tensors (pruning+quantization/bnb_implementation.py:386-486) and rebuilds
// the dense array on the host (`dense[indices] = values`, :468-471).  Here the indices / values are copied to the GPU
// as they are and scattered there: out (zero-filled by the caller's memset inside this call) [n_out] fp32,
// out[idx0[i] * cols + idx1[i]] = vals[i] (idx1 == NULL: flat indices).  Duplicate indices: last writer wins as in numpy
// is NOT guaranteed; the formats never contain duplicates.
// ---------------------------------------------------------------------------------------------
namespace {
template <typename I>
__global__ void __launch_bounds__(256)
k_scatter_f32(const I *__restrict__ idx0, const I *__restrict__ idx1, int64_t cols, const float *__restrict__ vals,
              int64_t nnz, float *__restrict__ out, int64_t n_out, int *__restrict__ err) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t at = (int64_t)idx0[i];
        if (idx1 != nullptr) at = at * cols + (int64_t)idx1[i];
        if (at < 0 || at >= n_out) {
            *err = 1;            // corrupt file: reported by the host wrapper, never written out of bounds
            continue;
        }
        out[at] = vals[i];
    }
}
}  // namespace

extern "C" int wq_scatter_dense_f32(const void *idx0, const void *idx1, int idx_bytes, int64_t cols, const float *vals,
                                    int64_t nnz, float *out, int64_t n_out, int *err_flag, wq_stream_t stream) {
    WQ_REQUIRE(nnz >= 0 && n_out >= 0 && cols >= 0, "wq_scatter_dense_f32: negative extent");
    WQ_REQUIRE(idx_bytes == 4 || idx_bytes == 8, "wq_scatter_dense_f32: indices must be int32 or int64");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_out > 0) {
        WQ_REQUIRE(out != nullptr, "wq_scatter_dense_f32: null output");
        WQ_CUDA(cudaMemsetAsync(out, 0, (size_t)n_out * sizeof(float), s));
    }
    if (nnz == 0) return WQ_OK;
    WQ_REQUIRE(idx0 && vals && err_flag, "wq_scatter_dense_f32: null pointer");
    const int64_t want = (nnz + 255) / 256;
    const unsigned grid = (unsigned)(want < (int64_t)wq_sm_count() * 16 ? want : (int64_t)wq_sm_count() * 16);
    if (idx_bytes == 8)
        k_scatter_f32<int64_t><<<grid, 256, 0, s>>>((const int64_t *)idx0, (const int64_t *)idx1, cols, vals, nnz, out, n_out, err_flag);
    else
        k_scatter_f32<int32_t><<<grid, 256, 0, s>>>((const int32_t *)idx0, (const int32_t *)idx1, cols, vals, nnz, out, n_out, err_flag);
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}
