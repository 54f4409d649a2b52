// gemv_wq.cu -- decode-shaped forward of the weight-only quantized linears (M <= 32 rows): bnb Linear4bit (NF4 / FP4),
// quanto QLinear with qint8 / qint4 / qint2 / qfloat8 weights.
//
// At decode time a Whisper layer multiplies a handful of activation rows by every weight matrix.  The tcgen05 GEMM of
// gemm_tc.cu serves those calls with N / 64 CTAs that each walk all of K alone through a TMA -> expand -> MMA pipeline:
// 12 us per launch for whisper-small NF4 at 16-32 rows (78 GB/s on the packed bytes, profiles/r02_bench.json), all of it
// pipeline latency.  Here the packed weights are streamed ONCE with plain vector loads by N / 16 CTAs, dequantized in
// registers with exactly the values the GEMM's expansion warps produce (code * absmax, scale * q - shift: rounded
// once to the activation dtype; int8 / e4m3 codes are exact), multiplied into fp32 accumulators on the CUDA cores and
// reduced across the warp in a fixed order.  bitsandbytes has the same split (gemv_4bit for single rows, dequantize +
// GEMM otherwise); unlike its kernel this one keeps fp32 accumulation and serves up to 32 rows.
// At 16 rows the FMA count (K * N * 16) is what bounds it, not the bytes -- see DESIGN.md section 3.1.
#include "common.cuh"

#include <cuda_fp8.h>

namespace {

enum WMode { W_NF4 = 0, W_I8 = 1, W_U4 = 2, W_F8 = 3 };

constexpr int GV_THREADS = 256;
constexpr int GV_WARPS = 8;
constexpr int GV_COLS = 2;                       // output columns per warp
constexpr int GV_TILE = GV_WARPS * GV_COLS;      // 16 columns per CTA
constexpr int GV_KCH = 1024;                     // K elements of the activation rows staged in shared memory at a time

struct GvArgs {
    const void *x;
    int M, K, N;
    const uint8_t *w;        // packed weights, row-major per output feature
    const float *s0;         // NF4: absmax [N, K/64]; U4: scale [N, K/group]; I8 / F8: scale [N]
    const float *s1;         // U4: shift [N, K/group]
    int group;               // U4 group size
    int quant_type;          // NF4 (0) / FP4 (1)
    const float *bias;       // fp32 [N] or nullptr
    void *y;
};

// sum 32 per-lane partials of 32 different quantities in a fixed order: afterwards lane l holds the total of v[l]
__device__ __forceinline__ float transpose_reduce32f(float (&v)[32], int lane) {
#pragma unroll
    for (int d = 16, n = 32; d >= 1; d >>= 1, n >>= 1) {
        const bool up = (lane & d) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = up ? v[i] : v[i + n / 2];
            const float keep = up ? v[i + n / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, d);
        }
    }
    return v[0];
}

template <typename T> __device__ __forceinline__ float round_to(float f) { return to_f32(from_f32<T>(f)); }

// Raw packed weights (and their block statistics) of 8 consecutive K positions of one output feature: fetched for a
// whole K chunk BEFORE anything is consumed, so that the chunk costs one L2 / HBM round trip, not one per 256 columns.
struct WRaw {
    uint2 p;
    float s0, s1;
};

template <int MODE>
__device__ __forceinline__ WRaw fetch_w8(const GvArgs &a, int n, int k) {
    WRaw r;
    r.s0 = r.s1 = 0.0f;
    if constexpr (MODE == W_NF4) {
        r.p = make_uint2(__ldg(reinterpret_cast<const uint32_t *>(a.w + ((size_t)n * a.K + k) / 2)), 0u);
        r.s0 = __ldg(a.s0 + (size_t)n * (a.K / 64) + k / 64);
    } else if constexpr (MODE == W_U4) {
        r.p = make_uint2(__ldg(reinterpret_cast<const uint32_t *>(a.w + ((size_t)n * a.K + k) / 2)), 0u);
        const size_t gi = (size_t)n * (a.K / a.group) + k / a.group;
        r.s0 = __ldg(a.s0 + gi);
        r.s1 = __ldg(a.s1 + gi);
    } else {
        r.p = __ldg(reinterpret_cast<const uint2 *>(a.w + (size_t)n * a.K + k));
    }
    return r;
}

// ... decoded to the fp32 value of what the GEMM feeds the tensor core
template <typename T, int MODE>
__device__ __forceinline__ void decode_w8(const WRaw &r, const float *lut, float (&wv)[8]) {
    if constexpr (MODE == W_NF4) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t byte = (r.p.x >> (8 * b)) & 0xffu;
            wv[2 * b] = round_to<T>(__fmul_rn(lut[byte >> 4], r.s0));          // first element in the HIGH nibble
            wv[2 * b + 1] = round_to<T>(__fmul_rn(lut[byte & 15u], r.s0));
        }
    } else if constexpr (MODE == W_U4) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t byte = (r.p.x >> (8 * b)) & 0xffu;
            wv[2 * b] = round_to<T>(__fsub_rn(__fmul_rn(r.s0, (float)(byte >> 4)), r.s1));
            wv[2 * b + 1] = round_to<T>(__fsub_rn(__fmul_rn(r.s0, (float)(byte & 15u)), r.s1));
        }
    } else {
        const uint32_t w2[2] = {r.p.x, r.p.y};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t byte = (w2[j >> 2] >> (8 * (j & 3))) & 0xffu;
            if constexpr (MODE == W_I8) {
                wv[j] = (float)(int8_t)byte;
            } else {
                const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)byte, __NV_E4M3);
                wv[j] = __half2float(*reinterpret_cast<const __half *>(&h));
            }
        }
    }
}

// F32IO: the reference's fp32 flows (quanto on a model that was never .half()-ed, model_utils.py:139-142).  The rows
// arrive as fp32, are rounded to T while they are staged -- the operand the tensor-core GEMM of the same flow sees
// after its cast pass -- and the fp32 accumulators are stored as fp32.
template <typename T, int MODE, int MT, bool F32IO = false>     // MT: rows rounded up to 8 / 16 / 32
__global__ void __launch_bounds__(GV_THREADS)
k_gemv_wq(const GvArgs a) {
    extern __shared__ __align__(16) uint8_t gv_smem[];
    T *sx = reinterpret_cast<T *>(gv_smem);                 // [MT][GV_KCH]
    __shared__ float s_lut[16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (MODE == W_NF4 && tid < 16) s_lut[tid] = a.quant_type ? kFP4Code[tid] : kNF4Code[tid];
    const int n0 = (int)blockIdx.x * GV_TILE + warp * GV_COLS;
    constexpr int ITER = GV_KCH / 256;                      // 8-element slices per lane and chunk
    // the weights do not depend on the predecessor kernel: the first chunk's are requested before the PDL wait
    WRaw raw[ITER][GV_COLS];
    auto fetch_chunk = [&](int k0) {
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int k = k0 + it * 256 + lane * 8;
#pragma unroll
            for (int c = 0; c < GV_COLS; ++c) {
                raw[it][c].p = make_uint2(0u, 0u);
                raw[it][c].s0 = raw[it][c].s1 = 0.0f;
                if (k < a.K && n0 + c < a.N) raw[it][c] = fetch_w8<MODE>(a, n0 + c, k);
            }
        }
    };
    fetch_chunk(0);
    pdl_prologue_done();
    const T *x = reinterpret_cast<const T *>(a.x);
    float acc[MT][GV_COLS];
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int c = 0; c < GV_COLS; ++c) acc[m][c] = 0.0f;

    for (int k0 = 0; k0 < a.K; k0 += GV_KCH) {
        const int kc = min(GV_KCH, a.K - k0);               // multiple of 8
        if (k0 > 0) fetch_chunk(k0);
        __syncthreads();
        {   // stage the activation rows of this chunk: all loads in flight before the first store
            constexpr int VPR = GV_KCH / 8;                          // 16-byte vectors per row
            constexpr int XV = MT * VPR / GV_THREADS;
            uint4 xr[XV];
#pragma unroll
            for (int j = 0; j < XV; ++j) {
                const int i = tid + j * GV_THREADS, m = i / VPR, c = i - m * VPR;
                xr[j] = make_uint4(0u, 0u, 0u, 0u);
                if (m < a.M && c * 8 < kc) {
                    if constexpr (F32IO) {
                        const float4 *xf = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(a.x) +
                                                                            (size_t)m * a.K + k0 + c * 8);
                        const float4 lo = xf[0], hi = xf[1];
                        T *t8 = reinterpret_cast<T *>(&xr[j]);
                        t8[0] = from_f32<T>(lo.x); t8[1] = from_f32<T>(lo.y); t8[2] = from_f32<T>(lo.z); t8[3] = from_f32<T>(lo.w);
                        t8[4] = from_f32<T>(hi.x); t8[5] = from_f32<T>(hi.y); t8[6] = from_f32<T>(hi.z); t8[7] = from_f32<T>(hi.w);
                    } else {
                        xr[j] = *reinterpret_cast<const uint4 *>(x + (size_t)m * a.K + k0 + c * 8);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < XV; ++j) {
                const int i = tid + j * GV_THREADS, m = i / VPR, c = i - m * VPR;
                *reinterpret_cast<uint4 *>(sx + m * GV_KCH + c * 8) = xr[j];
            }
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < ITER; ++it) {
            const int kk = it * 256 + lane * 8;
            if (kk >= kc) continue;
            float wv[GV_COLS][8];
#pragma unroll
            for (int c = 0; c < GV_COLS; ++c) decode_w8<T, MODE>(raw[it][c], s_lut, wv[c]);
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const uint4 rx = *reinterpret_cast<const uint4 *>(sx + m * GV_KCH + kk);
                const T *xe = reinterpret_cast<const T *>(&rx);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xf = to_f32(xe[j]);
#pragma unroll
                    for (int c = 0; c < GV_COLS; ++c) acc[m][c] = fmaf(xf, wv[c][j], acc[m][c]);
                }
            }
        }
    }
    // reduce across the warp, 32 (row, column) quantities at a time; lane l ends up with quantity l of the pass
    T *y = reinterpret_cast<T *>(a.y);
#pragma unroll
    for (int pass = 0; pass < MT * GV_COLS / 32; ++pass) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = acc[(pass * 32 + i) / GV_COLS][(pass * 32 + i) % GV_COLS];
        const float tot = transpose_reduce32f(v, lane);
        const int q = pass * 32 + lane, m = q / GV_COLS, n = n0 + q % GV_COLS;
        if (m < a.M && n < a.N) {
            float r = tot;
            if constexpr (MODE == W_I8 || MODE == W_F8) r = __fmul_rn(r, __ldg(a.s0 + n));   // quanto: scale after the matmul
            if (a.bias != nullptr) r = __fadd_rn(r, __ldg(a.bias + n));
            if constexpr (F32IO) reinterpret_cast<float *>(a.y)[(size_t)m * a.N + n] = r;
            else y[(size_t)m * a.N + n] = from_f32<T>(r);
        }
    }
    if constexpr (MT * GV_COLS < 32) {      // MT = 8: a single pass over 16 quantities padded to 32
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = i < MT * GV_COLS ? acc[i / GV_COLS][i % GV_COLS] : 0.0f;
        const float tot = transpose_reduce32f(v, lane);
        const int m = lane / GV_COLS, n = n0 + lane % GV_COLS;
        if (lane < MT * GV_COLS && m < a.M && n < a.N) {
            float r = tot;
            if constexpr (MODE == W_I8 || MODE == W_F8) r = __fmul_rn(r, __ldg(a.s0 + n));
            if (a.bias != nullptr) r = __fadd_rn(r, __ldg(a.bias + n));
            if constexpr (F32IO) reinterpret_cast<float *>(a.y)[(size_t)m * a.N + n] = r;
            else y[(size_t)m * a.N + n] = from_f32<T>(r);
        }
    }
}

template <typename T, int MODE, bool F32IO = false>
int launch_gemv(const GvArgs &a, cudaStream_t s) {
    const unsigned grid = (unsigned)((a.N + GV_TILE - 1) / GV_TILE);
#define WQ_GV(MT)                                                                                              \
    {                                                                                                          \
        const size_t smem = (size_t)MT * GV_KCH * sizeof(T);                                                   \
        static bool configured = false;                                                                        \
        if (!configured && smem > 48 * 1024) {                                                                 \
            WQ_CUDA(cudaFuncSetAttribute(k_gemv_wq<T, MODE, MT, F32IO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            configured = true;                                                                                 \
        }                                                                                                      \
        WQ_LAUNCH_PDL((k_gemv_wq<T, MODE, MT, F32IO>), dim3(grid), dim3(GV_THREADS), smem, s, a);              \
        return WQ_OK;                                                                                          \
    }
    if (a.M <= 8) WQ_GV(8)
    if (a.M <= 16) WQ_GV(16)
    WQ_GV(32)
#undef WQ_GV
}

}  // namespace

extern "C" int wq_gemv_weightonly(const void *x, int x_dtype, int64_t M, int64_t K, int mode, const void *w,
                                  const float *s0, const float *s1, int group, int quant_type, const float *bias,
                                  void *y, int64_t N, wq_stream_t stream) {
    WQ_REQUIRE(M >= 0 && N >= 0 && K > 0 && N < (1ll << 31) && K < (1ll << 31), "wq_gemv_weightonly: bad shape");
    if (M == 0 || N == 0) return WQ_OK;
    WQ_REQUIRE(M <= 32, "wq_gemv_weightonly: at most 32 rows (got %lld)", (long long)M);
    WQ_REQUIRE(x_dtype == WQ_F16 || x_dtype == WQ_BF16 || x_dtype == WQ_F32, "wq_gemv_weightonly: activations must be f16, bf16 or f32");
    WQ_REQUIRE(mode >= W_NF4 && mode <= W_F8, "wq_gemv_weightonly: bad mode %d", mode);
    WQ_REQUIRE(x && w && s0 && y, "wq_gemv_weightonly: null pointer");
    WQ_REQUIRE(K % 8 == 0, "wq_gemv_weightonly: K=%lld must be a multiple of 8", (long long)K);
    WQ_REQUIRE(mode != W_NF4 || K % 64 == 0, "wq_gemv_weightonly: 4-bit blocks of 64 need K %% 64 == 0");
    WQ_REQUIRE(mode != W_U4 || (s1 != nullptr && group >= 8 && group % 8 == 0 && K % group == 0),
               "wq_gemv_weightonly: bad group size %d", group);
    WQ_REQUIRE(wq_aligned(x, 16) && wq_aligned(w, 8), "wq_gemv_weightonly: misaligned buffer");
    GvArgs a = {};
    a.x = x; a.M = (int)M; a.K = (int)K; a.N = (int)N;
    a.w = (const uint8_t *)w; a.s0 = s0; a.s1 = s1; a.group = group; a.quant_type = quant_type; a.bias = bias; a.y = y;
    cudaStream_t s = (cudaStream_t)stream;
#define WQ_GV_MODE(MODE)                                                            \
    if (mode == MODE) {                                                             \
        if (x_dtype == WQ_F32) return launch_gemv<__half, MODE, true>(a, s);        \
        if (x_dtype == WQ_F16) return launch_gemv<__half, MODE>(a, s);              \
        return launch_gemv<__nv_bfloat16, MODE>(a, s);                              \
    }
    WQ_GV_MODE(W_NF4)
    WQ_GV_MODE(W_I8)
    WQ_GV_MODE(W_U4)
    WQ_GV_MODE(W_F8)
#undef WQ_GV_MODE
    return WQ_ERR_INVALID;
}
