// rowops.cu -- row-wise producers fused with the LLM.int8 activation quantizer.
//
// The quantized linears of a Whisper layer are fed by three kinds of producer: a LayerNorm (q/k/v, cross q,
// fc1), a GELU (fc2) and an attention output (out_proj).  bitsandbytes' Linear8bitLt quantizes its input row by
// row (int8_vectorwise_quant) in a launch of its own, and HF adds the residual in another
// (modeling_whisper.py: `hidden_states = residual + hidden_states`, `self_attn_layer_norm`, `activation_fn`).
// These kernels produce the fp16 tensor HF would have produced AND its int8 row quantization in one pass:
//
//   k_add_ln_quant : x' = x + delta (fp16 add, as torch), h = LayerNorm(x') -> x', h, [CA, SCA, outlier flags]
//   k_gelu_quant   : h = gelu(x) (erf form, fp32 math, as torch)          -> h,     [CA, SCA, outlier flags]
//
// The int8 codes are computed from the ROUNDED fp16 h, with exactly the arithmetic of k_quant_i8_rowwise_bnb
// (quant.cu), so `CA, SCA, flags` are bit-identical to running the stand-alone quantizer on the h written here.
// One warp per row; HBM-bound (decode: launch-bound, which is the point of fusing).
#include "common.cuh"

namespace {

constexpr int kWarps = 8;        // rows per CTA
constexpr int kMaxChunks = 8;    // 8 x 256 columns held in registers by the LayerNorm kernel

__device__ __forceinline__ float warp_sum_f32(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// |v| contribution to the row absmax under the LLM.int8 rule: outliers (|v| >= threshold) do not count.
__device__ __forceinline__ float absmax_step(float am, float v, bool sparse, float threshold) {
    const float x = fabsf(v);
    return (!sparse || x < threshold) ? fmaxf(am, x) : am;
}

// int8 code of one element (same arithmetic as k_quant_i8_rowwise_bnb); raises the column flag for outliers.
__device__ __forceinline__ uint32_t quant_elem(float v, float scale, bool sparse, float threshold,
                                               int32_t *col_flags, int64_t col, int64_t cols) {
    int q;
    if (sparse && !(fabsf(v) < threshold)) {
        q = 0;
        col_flags[col] = 1;
        col_flags[cols] = 1;
    } else {
        q = __float2int_rn(__fmul_rn(v, scale));
    }
    return (uint32_t)(q & 0xff);
}

template <typename T> struct Vec8 {
    uint4 raw;
    __device__ __forceinline__ T get(int j) const { return reinterpret_cast<const T *>(&raw)[j]; }
    __device__ __forceinline__ void set(int j, T v) { reinterpret_cast<T *>(&raw)[j] = v; }
};

// ---------------------------------------------------------------------------------------------
// x' = x + delta ; h = LayerNorm(x') ; optional int8 row quantization of h.  cols % 8 == 0, cols <= 2048.
// ---------------------------------------------------------------------------------------------
template <typename T, int CH>      // CH x 256 columns held in registers (sized to the row: 147 registers at CH = 8
                                   // left one CTA per SM and 0.25 of the HBM rate on d_model = 512 rows)
__global__ void __launch_bounds__(kWarps * 32)
k_add_ln_quant(const T *x, const T *__restrict__ delta, const T *__restrict__ gamma,
               const T *__restrict__ beta, float eps, int64_t rows, int cols, T *x_out /* may alias x */,
               T *__restrict__ h_out, float threshold, int8_t *__restrict__ ca, float *__restrict__ row_stats,
               int32_t *__restrict__ col_flags) {
    pdl_prologue_done();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int64_t base = row * cols;
    const bool sparse = threshold > 0.0f;

    float v[CH][8];
    Vec8<T> gam[CH], bet[CH];     // fetched up front: not a second L2 round trip after the statistics
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int c = i * 256 + lane * 8;
        if (c < cols) {
            gam[i].raw = *reinterpret_cast<const uint4 *>(gamma + c);
            bet[i].raw = *reinterpret_cast<const uint4 *>(beta + c);
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int c = i * 256 + lane * 8;
        if (c < cols) {
            Vec8<T> a;
            a.raw = *reinterpret_cast<const uint4 *>(x + base + c);
            if (delta != nullptr) {
                Vec8<T> d;
                d.raw = *reinterpret_cast<const uint4 *>(delta + base + c);
#pragma unroll
                for (int j = 0; j < 8; ++j) a.set(j, from_f32<T>(__fadd_rn(to_f32(a.get(j)), to_f32(d.get(j)))));
                *reinterpret_cast<uint4 *>(x_out + base + c) = a.raw;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[i][j] = to_f32(a.get(j));
                sum += v[i][j];
            }
        }
    }
    const float mean = warp_sum_f32(sum) / (float)cols;
    float m2 = 0.0f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        if (i * 256 + lane * 8 < cols) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[i][j] - mean;
                m2 += d * d;
            }
        }
    }
    const float rstd = rsqrtf(warp_sum_f32(m2) / (float)cols + eps);

    float am = 0.0f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int c = i * 256 + lane * 8;
        if (c < cols) {
            Vec8<T> h;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                // torch's layer_norm kernel: gamma * (rstd * (x - mean)) + beta in fp32, rounded once
                const T r = from_f32<T>(to_f32(gam[i].get(j)) * (rstd * (v[i][j] - mean)) + to_f32(bet[i].get(j)));
                h.set(j, r);
                v[i][j] = to_f32(r);
                am = absmax_step(am, v[i][j], sparse, threshold);
            }
            *reinterpret_cast<uint4 *>(h_out + base + c) = h.raw;
        }
    }
    if (ca == nullptr) return;
    am = warp_max(am);
    if (lane == 0) row_stats[row] = am;
    const float scale = bnb_row_scale(am);
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        const int c = i * 256 + lane * 8;
        if (c < cols) {
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t q = quant_elem(v[i][j], scale, sparse, threshold, col_flags, c + j, cols);
                if (j < 4) lo |= q << (8 * j); else hi |= q << (8 * (j - 4));
            }
            *reinterpret_cast<uint2 *>(ca + base + c) = make_uint2(lo, hi);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// h = gelu(x) (torch approximate='none': x * 0.5 * (1 + erf(x / sqrt(2))) in fp32); optional int8 row
// quantization of h.  Two passes over the row (the second re-reads the lane's own stores); cols % 8 == 0.
// ---------------------------------------------------------------------------------------------
template <typename T, int G>      // G warps share a row (G = 4 for decode-sized calls: rows are few, columns many)
__global__ void __launch_bounds__(kWarps * 32)
k_gelu_quant(const T *__restrict__ x, int64_t rows, int64_t cols, T *__restrict__ h_out, float threshold,
             int8_t *__restrict__ ca, float *__restrict__ row_stats, int32_t *__restrict__ col_flags) {
    __shared__ float s_am[kWarps];
    pdl_prologue_done();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * (kWarps / G) + warp / G;
    const int part = warp % G;
    const bool live = row < rows;
    const int64_t base = row * cols;
    const bool sparse = threshold > 0.0f;
    float am = 0.0f;
    if (live) {
        for (int64_t c = (part * 32 + lane) * 8; c < cols; c += 256 * G) {
            Vec8<T> a, h;
            a.raw = *reinterpret_cast<const uint4 *>(x + base + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const T r = gelu_erf<T>(to_f32(a.get(j)));
                h.set(j, r);
                am = absmax_step(am, to_f32(r), sparse, threshold);
            }
            if (h_out != nullptr) *reinterpret_cast<uint4 *>(h_out + base + c) = h.raw;
        }
    }
    if (ca == nullptr) return;
    am = warp_max(am);
    if (G > 1) {
        if (lane == 0) s_am[warp] = am;
        __syncthreads();
        am = 0.0f;
#pragma unroll
        for (int w = 0; w < G; ++w) am = fmaxf(am, s_am[(warp / G) * G + w]);
    }
    if (!live) return;
    if (lane == 0 && part == 0) row_stats[row] = am;
    const float scale = bnb_row_scale(am);
    for (int64_t c = (part * 32 + lane) * 8; c < cols; c += 256 * G) {
        Vec8<T> h;
        if (h_out != nullptr) {
            h.raw = *reinterpret_cast<const uint4 *>(h_out + base + c);     // this lane's own store
        } else {                                                            // h not kept: evaluate it again
            Vec8<T> a;
            a.raw = *reinterpret_cast<const uint4 *>(x + base + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) h.set(j, gelu_erf<T>(to_f32(a.get(j))));
        }
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t q = quant_elem(to_f32(h.get(j)), scale, sparse, threshold, col_flags, c + j, cols);
            if (j < 4) lo |= q << (8 * j); else hi |= q << (8 * (j - 4));
        }
        *reinterpret_cast<uint2 *>(ca + base + c) = make_uint2(lo, hi);
    }
}

// ---------------------------------------------------------------------------------------------
// The same for encoder-sized fp16 calls (rows >= 4096, cols <= 2048), through a table.  An fp16 GELU is a function
// of 65536 bit patterns: k_gelu_table_fill evaluates gelu_erf<__half> once per pattern into a 128 KB device table,
// every CTA of k_gelu_quant_lut copies it into shared memory and then replaces ~40 fp32 instructions per element
// (erff) by one 2-byte shared-memory load -- bit-identical by construction, and the kernel goes from
// instruction-bound (1.19 ms per 384000 x 2048 on B200) to the HBM stream it is (2 B read, 3 B written per element).
// Persistent CTAs of 16 warps, one warp per row, the row's NCH x 16-byte chunks per lane are loaded up front and
// converted in place (NCH = 8 / 12 / 16 / 20: rows up to 2048 / 3072 / 4096 / 5120 columns -- every Whisper ffn width).
// h_out may be NULL when only the int8 rows are wanted: the consumer GEMM re-derives the few fp16 values its outlier
// path needs from x (wq_gemm_llmint8_residual, a_pre_gelu), and 2 of the 5 bytes per element never reach HBM.
// ---------------------------------------------------------------------------------------------
constexpr int kLutWarps = 16;
constexpr int kLutEntries = 65536;
__device__ __half g_gelu_table[kLutEntries];

__global__ void k_gelu_table_fill() {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (unsigned)kLutEntries) g_gelu_table[i] = gelu_erf<__half>(__half2float(__ushort_as_half((unsigned short)i)));
}

template <int NCH>
__global__ void __launch_bounds__(kLutWarps * 32, 1)
k_gelu_quant_lut(const __half *__restrict__ x, int64_t rows, int cols, __half *__restrict__ h_out, float threshold,
                 int8_t *__restrict__ ca, float *__restrict__ row_stats, int32_t *__restrict__ col_flags) {
    extern __shared__ __align__(16) unsigned short s_lut[];
    {   // the table does not depend on the predecessor kernel: copied before the PDL wait
        const uint4 *src = reinterpret_cast<const uint4 *>(g_gelu_table);
        uint4 *dst = reinterpret_cast<uint4 *>(s_lut);
        for (int i = threadIdx.x; i < kLutEntries * 2 / 16; i += kLutWarps * 32) dst[i] = src[i];
    }
    __syncthreads();
    pdl_prologue_done();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool sparse = threshold > 0.0f;
    for (int64_t row = (int64_t)blockIdx.x * kLutWarps + warp; row < rows; row += (int64_t)gridDim.x * kLutWarps) {
        const int64_t base = row * cols;
        uint4 v[NCH];
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = i * 256 + lane * 8;
            if (c < cols) v[i] = __ldcs(reinterpret_cast<const uint4 *>(x + base + c));     // streamed once
        }
        float am = 0.0f;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = i * 256 + lane * 8;
            if (c < cols) {
                uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t lo = s_lut[w[j] & 0xffffu], hi = s_lut[w[j] >> 16];
                    am = absmax_step(am, __half2float(__ushort_as_half((unsigned short)lo)), sparse, threshold);
                    am = absmax_step(am, __half2float(__ushort_as_half((unsigned short)hi)), sparse, threshold);
                    w[j] = lo | (hi << 16);
                }
                v[i] = make_uint4(w[0], w[1], w[2], w[3]);
                if (h_out != nullptr) *reinterpret_cast<uint4 *>(h_out + base + c) = v[i];
            }
        }
        if (ca == nullptr) continue;
        am = warp_max(am);
        if (lane == 0) row_stats[row] = am;
        const float scale = bnb_row_scale(am);
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const int c = i * 256 + lane * 8;
            if (c < cols) {
                const __half *h8 = reinterpret_cast<const __half *>(&v[i]);
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t q = quant_elem(__half2float(h8[j]), scale, sparse, threshold, col_flags, c + j, cols);
                    if (j < 4) lo |= q << (8 * j); else hi |= q << (8 * (j - 4));
                }
                *reinterpret_cast<uint2 *>(ca + base + c) = make_uint2(lo, hi);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Greedy token choice: argmax over a row of logits with a per-call suppression mask (Whisper's logits
// processors only ever write -inf at positions that depend on the current length: SuppressTokens,
// SuppressTokensAtBegin, Min(New)Length -- transformers generation/logits_process.py).  torch.argmax semantics:
// first index among equal maxima, NaN counts as the maximum.  One CTA per row.
// ---------------------------------------------------------------------------------------------
struct ArgBest {
    float v;
    int i;      // INT_MAX = nothing seen yet
};
__device__ __forceinline__ void arg_merge(ArgBest &b, int &nan_i, float ov, int oi, int onan) {
    if (oi != 0x7fffffff && (b.i == 0x7fffffff || ov > b.v || (ov == b.v && oi < b.i))) b = ArgBest{ov, oi};
    nan_i = min(nan_i, onan);
}

template <typename T>
__global__ void __launch_bounds__(512)
k_masked_argmax(const T *__restrict__ logits, int64_t ld, int V, const uint8_t *__restrict__ mask,
                int64_t *__restrict__ out) {
    pdl_prologue_done();
    const T *row = logits + (int64_t)blockIdx.x * ld;
    // each thread scans its columns in increasing order, so a strict `>` keeps the first maximum; NaNs are
    // tracked on the side (torch: NaN is the maximum, first one wins)
    ArgBest best{-INFINITY, 0x7fffffff};
    int nan_i = 0x7fffffff;
    const int V8 = V & ~7;
    for (int c = threadIdx.x * 8; c < V8; c += 512 * 8) {
        Vec8<T> a;
        a.raw = *reinterpret_cast<const uint4 *>(row + c);
        uint2 m = make_uint2(0u, 0u);
        if (mask != nullptr) m = *reinterpret_cast<const uint2 *>(mask + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t mb = ((j < 4 ? m.x : m.y) >> (8 * (j & 3))) & 0xffu;
            const float v = mb ? -INFINITY : to_f32(a.get(j));
            if (v != v) nan_i = min(nan_i, c + j);
            else if (v > best.v || best.i == 0x7fffffff) best = ArgBest{v, c + j};
        }
    }
    for (int c = V8 + threadIdx.x; c < V; c += 512) {
        const float v = (mask != nullptr && mask[c]) ? -INFINITY : to_f32(row[c]);
        if (v != v) nan_i = min(nan_i, c);
        else if (v > best.v || best.i == 0x7fffffff) best = ArgBest{v, c};
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best.i, o);
        const int on = __shfl_xor_sync(0xffffffffu, nan_i, o);
        arg_merge(best, nan_i, ov, oi, on);
    }
    __shared__ float sv[16];
    __shared__ int si[16], sn[16];
    if ((threadIdx.x & 31) == 0) {
        sv[threadIdx.x >> 5] = best.v;
        si[threadIdx.x >> 5] = best.i;
        sn[threadIdx.x >> 5] = nan_i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 16; ++w) arg_merge(best, nan_i, sv[w], si[w], sn[w]);
        out[blockIdx.x] = nan_i != 0x7fffffff ? nan_i : best.i;
    }
}

}  // namespace

extern "C" int wq_add_layernorm_quant(const void *x, const void *delta, int dtype, const void *gamma,
                                      const void *beta, float eps, int64_t rows, int64_t cols, void *x_out,
                                      void *h_out, float threshold, int8_t *ca, float *row_stats,
                                      int32_t *col_flags, wq_stream_t stream) {
    WQ_REQUIRE(rows >= 0 && cols > 0, "wq_add_layernorm_quant: bad shape");
    WQ_REQUIRE(cols % 8 == 0 && cols <= 256 * kMaxChunks,
               "wq_add_layernorm_quant: cols must be a multiple of 8 and <= %d (got %lld)", 256 * kMaxChunks,
               (long long)cols);
    WQ_REQUIRE(dtype == WQ_F16 || dtype == WQ_BF16, "wq_add_layernorm_quant: dtype must be f16 or bf16");
    WQ_REQUIRE(threshold >= 0.0f, "wq_add_layernorm_quant: negative threshold");
    if (rows == 0) return WQ_OK;
    WQ_REQUIRE(x && gamma && beta && h_out, "wq_add_layernorm_quant: null pointer");
    WQ_REQUIRE(delta == nullptr || x_out != nullptr, "wq_add_layernorm_quant: delta needs x_out");
    WQ_REQUIRE(ca == nullptr || (dtype == WQ_F16 && row_stats != nullptr),
               "wq_add_layernorm_quant: the int8 outputs need fp16 rows and row_stats");
    WQ_REQUIRE(ca == nullptr || threshold == 0.0f || col_flags, "wq_add_layernorm_quant: threshold needs col_flags");
    WQ_REQUIRE(wq_aligned(x, 16) && wq_aligned(h_out, 16) && wq_aligned(gamma, 16) && wq_aligned(beta, 16) &&
                   (delta == nullptr || (wq_aligned(delta, 16) && wq_aligned(x_out, 16))) &&
                   (ca == nullptr || wq_aligned(ca, 8)),
               "wq_add_layernorm_quant: pointers must be 16-byte aligned");
    const unsigned grid = (unsigned)((rows + kWarps - 1) / kWarps);
    cudaStream_t s = (cudaStream_t)stream;
    const int chunks = (int)((cols + 255) / 256);
#define WQ_LN_CASE(CH)                                                                                                 \
    if (chunks <= CH) {                                                                                                \
        if (dtype == WQ_F16) {                                                                                         \
            WQ_LAUNCH_PDL((k_add_ln_quant<__half, CH>), dim3(grid), dim3(kWarps * 32), 0, s, (const __half *)x,        \
                          (const __half *)delta, (const __half *)gamma, (const __half *)beta, eps, rows, (int)cols,    \
                          (__half *)x_out, (__half *)h_out, threshold, ca, row_stats, col_flags);                      \
        } else {                                                                                                       \
            WQ_LAUNCH_PDL((k_add_ln_quant<__nv_bfloat16, CH>), dim3(grid), dim3(kWarps * 32), 0, s,                    \
                          (const __nv_bfloat16 *)x, (const __nv_bfloat16 *)delta, (const __nv_bfloat16 *)gamma,        \
                          (const __nv_bfloat16 *)beta, eps, rows, (int)cols, (__nv_bfloat16 *)x_out,                   \
                          (__nv_bfloat16 *)h_out, threshold, (int8_t *)nullptr, (float *)nullptr, (int32_t *)nullptr); \
        }                                                                                                              \
        return WQ_OK;                                                                                                  \
    }
    WQ_LN_CASE(2)
    WQ_LN_CASE(3)
    WQ_LN_CASE(4)
    WQ_LN_CASE(5)
    WQ_LN_CASE(6)
    WQ_LN_CASE(8)
#undef WQ_LN_CASE
    return WQ_OK;
}

extern "C" int wq_gelu_quant(const void *x, int dtype, int64_t rows, int64_t cols, void *h_out, float threshold,
                             int8_t *ca, float *row_stats, int32_t *col_flags, wq_stream_t stream) {
    WQ_REQUIRE(rows >= 0 && cols > 0, "wq_gelu_quant: bad shape");
    WQ_REQUIRE(cols % 8 == 0, "wq_gelu_quant: cols must be a multiple of 8 (got %lld)", (long long)cols);
    WQ_REQUIRE(dtype == WQ_F16 || dtype == WQ_BF16, "wq_gelu_quant: dtype must be f16 or bf16");
    WQ_REQUIRE(threshold >= 0.0f, "wq_gelu_quant: negative threshold");
    if (rows == 0) return WQ_OK;
    WQ_REQUIRE(x && (h_out || ca), "wq_gelu_quant: null pointer (h_out may be NULL only when the int8 rows are asked for)");
    WQ_REQUIRE(ca == nullptr || (dtype == WQ_F16 && row_stats != nullptr),
               "wq_gelu_quant: the int8 outputs need fp16 rows and row_stats");
    WQ_REQUIRE(ca == nullptr || threshold == 0.0f || col_flags, "wq_gelu_quant: threshold needs col_flags");
    WQ_REQUIRE(wq_aligned(x, 16) && (h_out == nullptr || wq_aligned(h_out, 16)) && (ca == nullptr || wq_aligned(ca, 8)),
               "wq_gelu_quant: pointers must be 16-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == WQ_F16 && rows >= 4096 && cols <= 256 * 20) {
        // table path (k_gelu_quant_lut).  The table of a device is filled on first use, synchronously; a first use
        // under stream capture (no synchronisation allowed) takes the erff kernel below instead.
        static bool ready[64] = {};
        static bool configured = false;
        int dev = 0;
        WQ_CUDA(cudaGetDevice(&dev));
        bool ok = dev >= 0 && dev < 64;
        if (ok && !ready[dev]) {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            WQ_CUDA(cudaStreamIsCapturing(s, &cap));
            if (cap == cudaStreamCaptureStatusNone) {
                k_gelu_table_fill<<<kLutEntries / 256, 256, 0, s>>>();
                WQ_LAUNCH_CHECK();
                WQ_CUDA(cudaStreamSynchronize(s));
                ready[dev] = true;
            } else {
                ok = false;
            }
        }
        if (ok) {
            if (!configured) {
                WQ_CUDA(cudaFuncSetAttribute(k_gelu_quant_lut<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLutEntries * 2));
                WQ_CUDA(cudaFuncSetAttribute(k_gelu_quant_lut<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLutEntries * 2));
                WQ_CUDA(cudaFuncSetAttribute(k_gelu_quant_lut<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLutEntries * 2));
                WQ_CUDA(cudaFuncSetAttribute(k_gelu_quant_lut<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLutEntries * 2));
                configured = true;
            }
            const int64_t want = (rows + kLutWarps - 1) / kLutWarps;
            const unsigned grid = (unsigned)(want < wq_sm_count() ? want : wq_sm_count());
#define WQ_LUT_LAUNCH(NCH)                                                                                           \
            WQ_LAUNCH_PDL(k_gelu_quant_lut<NCH>, dim3(grid), dim3(kLutWarps * 32), (size_t)kLutEntries * 2, s,           \
                          (const __half *)x, rows, (int)cols, (__half *)h_out, threshold, ca, row_stats, col_flags)
            if (cols <= 256 * 8) WQ_LUT_LAUNCH(8);
            else if (cols <= 256 * 12) WQ_LUT_LAUNCH(12);
            else if (cols <= 256 * 16) WQ_LUT_LAUNCH(16);
            else WQ_LUT_LAUNCH(20);
#undef WQ_LUT_LAUNCH
            return WQ_OK;
        }
    }
    const bool split = rows <= 4096 && cols >= 1024;     // few long rows: 4 warps per row
    const int rows_per_cta = split ? kWarps / 4 : kWarps;
    const unsigned grid = (unsigned)((rows + rows_per_cta - 1) / rows_per_cta);
    if (dtype == WQ_F16) {
        auto kern = split ? k_gelu_quant<__half, 4> : k_gelu_quant<__half, 1>;
        WQ_LAUNCH_PDL(kern, dim3(grid), dim3(kWarps * 32), 0, s, (const __half *)x, rows, cols, (__half *)h_out,
                      threshold, ca, row_stats, col_flags);
    } else {
        auto kern = split ? k_gelu_quant<__nv_bfloat16, 4> : k_gelu_quant<__nv_bfloat16, 1>;
        WQ_LAUNCH_PDL(kern, dim3(grid), dim3(kWarps * 32), 0, s, (const __nv_bfloat16 *)x, rows, cols,
                      (__nv_bfloat16 *)h_out, threshold, (int8_t *)nullptr, (float *)nullptr, (int32_t *)nullptr);
    }
    return WQ_OK;
}

extern "C" int wq_masked_argmax(const void *logits, int dtype, int64_t rows, int64_t cols, int64_t ld,
                                const uint8_t *mask, int64_t *out, wq_stream_t stream) {
    WQ_REQUIRE(rows >= 0 && cols > 0 && cols < (1ll << 31) && ld >= cols, "wq_masked_argmax: bad shape");
    WQ_REQUIRE(dtype == WQ_F16 || dtype == WQ_BF16, "wq_masked_argmax: dtype must be f16 or bf16");
    if (rows == 0) return WQ_OK;
    WQ_REQUIRE(logits && out, "wq_masked_argmax: null pointer");
    WQ_REQUIRE(ld % 8 == 0 && wq_aligned(logits, 16) && (mask == nullptr || wq_aligned(mask, 8)),
               "wq_masked_argmax: logits rows must be 16-byte aligned (ld %% 8 == 0), mask 8-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == WQ_F16)
        WQ_LAUNCH_PDL(k_masked_argmax<__half>, dim3((unsigned)rows), dim3(512), 0, s, (const __half *)logits, ld,
                      (int)cols, mask, out);
    else
        WQ_LAUNCH_PDL(k_masked_argmax<__nv_bfloat16>, dim3((unsigned)rows), dim3(512), 0, s,
                      (const __nv_bfloat16 *)logits, ld, (int)cols, mask, out);
    return WQ_OK;
}
