// gemv_small.cu -- fused Linear8bitLt forward for decode-shaped calls (M <= 64 rows).
//
// A decode step multiplies a handful of activation rows by every weight matrix: the work is the
// N*K weight bytes (HBM / L2 bound) and, at Whisper sizes, launch latency.  One kernel does what
// bitsandbytes does in five launches: row-wise int8 quantization of the activations with the
// LLM.int8 outlier rule (each CTA quantizes the few rows itself into shared memory -- 2*M*K bytes
// from L2), the int8 x int8 -> int32 products with dp4a while streaming its slice of W once with
// 16-byte loads, the int8_mm_dequant formula, and the fp16 outlier side product.  Results are
// bit-identical to wq_quant_i8_rowwise_bnb + wq_gemm_llmint8 (same integer sums, same fp32
// operation order).
#include "common.cuh"

namespace {

constexpr int GS_THREADS = 256;
constexpr int GS_WARPS = GS_THREADS / 32;
constexpr int GS_MAXM = 64;

// sum 32 per-lane partials of 32 different quantities: afterwards lane l holds the total of v[l]
__device__ __forceinline__ int transpose_reduce32(int (&v)[32], int lane) {
#pragma unroll
    for (int d = 16, n = 32; d >= 1; d >>= 1, n >>= 1) {
        const bool up = (lane & d) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const int send = up ? v[i] : v[i + n / 2];
            const int keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
    }
    return v[0];
}

template <bool HI>   // HI: rows 32..63 exist
__global__ void __launch_bounds__(GS_THREADS)
k_llmint8_small(const __half *__restrict__ a, int M, int K, float threshold, const int8_t *__restrict__ cb,
                const float *__restrict__ scb, const float *__restrict__ bias, __half *__restrict__ y, int N,
                int cols_per_cta) {
    extern __shared__ __align__(16) uint8_t gs_smem[];
    int8_t *s_ca = reinterpret_cast<int8_t *>(gs_smem);                       // [M][K]
    float *s_sca = reinterpret_cast<float *>(gs_smem + (size_t)GS_MAXM * K);  // [64]
    uint8_t *s_flag = reinterpret_cast<uint8_t *>(s_sca + GS_MAXM);           // [K]
    __shared__ int s_any;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool sparse = threshold > 0.0f;

    if (tid == 0) s_any = 0;
    for (int c = tid; c < K; c += GS_THREADS) s_flag[c] = 0;
    __syncthreads();

    // ---- phase 1: int8_vectorwise_quant of the M rows into shared memory ----
    for (int m = warp; m < M; m += GS_WARPS) {
        const __half *pr = a + (size_t)m * K;
        float am = 0.0f;
        for (int c = lane * 8; c < K; c += 256) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(pr + c);
            const __half2 *h = reinterpret_cast<const __half2 *>(&raw);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h[j]);
                const float x0 = fabsf(f.x), x1 = fabsf(f.y);
                if (!sparse || x0 < threshold) am = fmaxf(am, x0);
                if (!sparse || x1 < threshold) am = fmaxf(am, x1);
            }
        }
        am = warp_max(am);
        if (lane == 0) s_sca[m] = am;
        const float scale = bnb_row_scale(am);
        for (int c = lane * 8; c < K; c += 256) {
            const uint4 raw = *reinterpret_cast<const uint4 *>(pr + c);
            const __half *h = reinterpret_cast<const __half *>(&raw);
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float v = __half2float(h[j]);
                int q;
                if (sparse && !(fabsf(v) < threshold)) {
                    q = 0;
                    s_flag[c + j] = 1;
                    s_any = 1;
                } else {
                    q = __float2int_rn(__fmul_rn(v, scale));
                }
                const uint32_t b = (uint32_t)(q & 0xff);
                if (j < 4) lo |= b << (8 * j); else hi |= b << (8 * (j - 4));
            }
            *reinterpret_cast<uint2 *>(s_ca + (size_t)m * K + c) = make_uint2(lo, hi);
        }
    }
    __syncthreads();
    const bool any = s_any != 0;
    if (any) {   // CA[:, outlier_cols] = 0
        for (int c = tid; c < K; c += GS_THREADS)
            if (s_flag[c])
                for (int m = 0; m < M; ++m) s_ca[(size_t)m * K + c] = 0;
        __syncthreads();
    }

    // ---- phase 2: one output column per warp iteration, W streamed once ----
    const int n_begin = blockIdx.x * cols_per_cta;
    const int n_end = min(n_begin + cols_per_cta, N);
    const int chunks = K / 16;
    for (int n = n_begin + warp; n < n_end; n += GS_WARPS) {
        int acc_lo[32], acc_hi[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { acc_lo[i] = 0; acc_hi[i] = 0; }
        const int8_t *wrow = cb + (size_t)n * K;
        for (int ch = lane; ch < chunks; ch += 32) {
            const int4 w4 = __ldg(reinterpret_cast<const int4 *>(wrow) + ch);
            const int8_t *abase = s_ca + ch * 16;
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                if (m < M) {
                    const int4 a4 = *reinterpret_cast<const int4 *>(abase + (size_t)m * K);
                    int s = acc_lo[m];
                    s = __dp4a(a4.x, w4.x, s); s = __dp4a(a4.y, w4.y, s);
                    s = __dp4a(a4.z, w4.z, s); s = __dp4a(a4.w, w4.w, s);
                    acc_lo[m] = s;
                }
            }
            if constexpr (HI) {
#pragma unroll
                for (int m = 0; m < 32; ++m) {
                    if (m + 32 < M) {
                        const int4 a4 = *reinterpret_cast<const int4 *>(abase + (size_t)(m + 32) * K);
                        int s = acc_hi[m];
                        s = __dp4a(a4.x, w4.x, s); s = __dp4a(a4.y, w4.y, s);
                        s = __dp4a(a4.z, w4.z, s); s = __dp4a(a4.w, w4.w, s);
                        acc_hi[m] = s;
                    }
                }
            }
        }
        const int c_lo = transpose_reduce32(acc_lo, lane);
        int c_hi = 0;
        if constexpr (HI) c_hi = transpose_reduce32(acc_hi, lane);
        const float cs = __ldg(scb + n);
        const float b = bias != nullptr ? __ldg(bias + n) : 0.0f;
#pragma unroll
        for (int half_i = 0; half_i < (HI ? 2 : 1); ++half_i) {
            const int m = lane + 32 * half_i;
            if (m >= M) continue;
            const int c32 = half_i ? c_hi : c_lo;
            const float x = __fmul_rn(__fmul_rn((float)c32, s_sca[m]), cs);
            float v = __fmaf_rn(x, 6.200012e-05f, b);
            if (any) {   // mixed-precision decomposition: fp16 side product over the outlier columns
                float o = 0.0f;
                for (int c = 0; c < K; ++c) {
                    if (!s_flag[c]) continue;
                    const float d = __fmul_rn(__fmul_rn((float)wrow[c], cs), 7.874015718698502e-3f);
                    o = fmaf(__half2float(a[(size_t)m * K + c]), __half2float(__float2half_rn(d)), o);
                }
                v = __half2float(__float2half_rn(v)) + o;
            }
            y[(size_t)m * N + n] = __float2half_rn(v);
        }
    }
}

}  // namespace

extern "C" int wq_linear_llmint8_small(const void *a_f16, int64_t M, int64_t K, float threshold, const int8_t *cb,
                                       const float *scb, const float *bias, void *y_f16, int64_t N,
                                       wq_stream_t stream) {
    WQ_REQUIRE(M >= 0 && N >= 0 && K > 0, "wq_linear_llmint8_small: bad shape");
    if (M == 0 || N == 0) return WQ_OK;
    WQ_REQUIRE(M <= GS_MAXM, "wq_linear_llmint8_small: M=%lld exceeds %d rows", (long long)M, GS_MAXM);
    WQ_REQUIRE(K % 16 == 0, "wq_linear_llmint8_small: K=%lld must be a multiple of 16", (long long)K);
    WQ_REQUIRE(a_f16 && cb && scb && y_f16, "wq_linear_llmint8_small: null pointer");
    WQ_REQUIRE(wq_aligned(a_f16, 16) && wq_aligned(cb, 16), "wq_linear_llmint8_small: misaligned buffer");
    const size_t smem = (size_t)GS_MAXM * K + GS_MAXM * sizeof(float) + (size_t)K + 16;
    WQ_REQUIRE(smem <= 200 * 1024, "wq_linear_llmint8_small: K=%lld too large for the shared-memory activation tile",
               (long long)K);
    static bool configured = false;
    if (!configured) {
        WQ_CUDA(cudaFuncSetAttribute(k_llmint8_small<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        WQ_CUDA(cudaFuncSetAttribute(k_llmint8_small<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    // one column per warp iteration; spread the columns over the SMs in multiples of 8
    const int sms = wq_sm_count();
    int cols_per_cta = (int)((N + sms - 1) / sms);
    cols_per_cta = ((cols_per_cta + GS_WARPS - 1) / GS_WARPS) * GS_WARPS;
    const unsigned grid = (unsigned)((N + cols_per_cta - 1) / cols_per_cta);
    cudaStream_t s = (cudaStream_t)stream;
    if (M > 32)
        k_llmint8_small<true><<<grid, GS_THREADS, smem, s>>>((const __half *)a_f16, (int)M, (int)K, threshold, cb, scb,
                                                              bias, (__half *)y_f16, (int)N, cols_per_cta);
    else
        k_llmint8_small<false><<<grid, GS_THREADS, smem, s>>>((const __half *)a_f16, (int)M, (int)K, threshold, cb, scb,
                                                               bias, (__half *)y_f16, (int)N, cols_per_cta);
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}
