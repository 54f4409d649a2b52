// logmel.cu -- Whisper log-mel frontend (HF WhisperFeatureExtractor numerics) on the GPU.
//
//   pad/truncate to n_samples -> reflect-pad 200 -> periodic hann(400) -> 400-point real DFT
//   every 160 samples (last frame dropped) -> |.|^2 -> mel filterbank -> log10(clamp 1e-10) ->
//   max(., utterance max - 8) -> (. + 4) / 4
//
// One CTA owns 16 consecutive frames of one utterance: the 2800 samples they span are loaded
// once into shared memory (coalesced 16-byte loads), each real 400-point DFT is computed as a
// 200-point complex FFT (10 x 20 Cooley-Tukey, the 10- and 20-point DFTs as 2x5 and 4x5; warp-uniform twiddles in constant memory, per-thread
// tables in shared memory, padded strides against bank conflicts) plus the even/odd split, the
// mel projection only walks each triangle's non-zero support, and the per-utterance maximum is
// reduced with warp shuffles + one atomicMax per CTA.  A second elementwise pass applies the
// max-8 floor (the output of pass one is still L2-resident for typical batches).
// Algorithmic HBM bytes per utterance: 4*n_samples read + 4*n_mels*(n_samples/160) written.
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace {

constexpr int NFFT = 400, HOP = 160, NBIN = 201, HALF = 200;
constexpr int FR = 16;                          // frames per CTA
constexpr int SPAN = (FR - 1) * HOP + NFFT;     // 2800 samples
constexpr int LM_THREADS = 256;

// twiddles with compile-time (warp-uniform) indices stay in constant memory; tables indexed per
// thread (window, W200^(n2*k1), W400^k) would serialise on the constant cache, so they live in
// global memory and are staged into shared memory by every CTA
__constant__ float2 c_w10[10];
__constant__ float2 c_w20[20];
struct LmTables {
    float win[NFFT];
    float2 w200[HALF];
    float2 w400[NBIN + 1];
};
__device__ LmTables g_tables;
constexpr int S1 = 21;                 // padded k1 stride of the stage-1 output (bank-conflict free)
constexpr int SF = 10 * S1;            // per-frame stride of the stage-1 output

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ void cfma(float2 &acc, float2 a, float2 w) {
    acc.x = fmaf(a.x, w.x, acc.x);
    acc.x = fmaf(-a.y, w.y, acc.x);
    acc.y = fmaf(a.x, w.y, acc.y);
    acc.y = fmaf(a.y, w.x, acc.y);
}

// 5-point DFT, twiddles W5^j = W10^(2j) from constant memory (compile-time indices)
__device__ __forceinline__ void dft5(const float2 (&y)[5], float2 (&z)[5]) {
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        float2 acc = y[0];
#pragma unroll
        for (int b = 1; b < 5; ++b) cfma(acc, y[b], c_w10[(2 * b * d) % 10]);
        z[d] = acc;
    }
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_neg_i(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// 10-point DFT = 2 x 5 (index m = 5a + b, k = c + 2d)
__device__ __forceinline__ void dft10(const float2 (&in)[10], float2 (&out)[10]) {
    float2 y0[5], y1[5], z[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        y0[b] = cadd(in[b], in[5 + b]);
        y1[b] = csub(in[b], in[5 + b]);
        if (b) y1[b] = cmul(y1[b], c_w10[b]);
    }
    dft5(y0, z);
#pragma unroll
    for (int d = 0; d < 5; ++d) out[2 * d] = z[d];
    dft5(y1, z);
#pragma unroll
    for (int d = 0; d < 5; ++d) out[2 * d + 1] = z[d];
}

// 20-point DFT = 4 x 5 (index n = 5a + b, k = c + 4d); W4 = -i is free
__device__ __forceinline__ void dft20(const float2 (&in)[20], float2 (&out)[20]) {
    float2 y[4][5];
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const float2 t0 = cadd(in[b], in[10 + b]), t1 = csub(in[b], in[10 + b]);
        const float2 t2 = cadd(in[5 + b], in[15 + b]), t3 = mul_neg_i(csub(in[5 + b], in[15 + b]));
        y[0][b] = cadd(t0, t2);
        y[1][b] = cadd(t1, t3);
        y[2][b] = csub(t0, t2);
        y[3][b] = csub(t1, t3);
        if (b) {
            y[1][b] = cmul(y[1][b], c_w20[b]);
            y[2][b] = cmul(y[2][b], c_w20[2 * b]);
            y[3][b] = cmul(y[3][b], c_w20[(3 * b) % 20]);
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float2 z[5];
        dft5(y[c], z);
#pragma unroll
        for (int d = 0; d < 5; ++d) out[c + 4 * d] = z[d];
    }
}

// one warp per mel filter: first / last non-zero bin of its triangle
__global__ void k_mel_ranges(const float *__restrict__ filters, int n_mels, int *__restrict__ lo,
                             int *__restrict__ hi) {
    const int lane = threadIdx.x & 31;
    const int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (m >= n_mels) return;
    int l = NBIN, h = -1;
    for (int k = lane; k < NBIN; k += 32)
        if (filters[k * n_mels + m] != 0.0f) {
            l = min(l, k);
            h = max(h, k);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        l = min(l, __shfl_xor_sync(0xffffffffu, l, o));
        h = max(h, __shfl_xor_sync(0xffffffffu, h, o));
    }
    if (lane == 0) {
        lo[m] = l;
        hi[m] = h;
    }
}

__global__ void __launch_bounds__(LM_THREADS)
k_logmel_main(const float *__restrict__ audio, int64_t audio_stride, const int32_t *__restrict__ lengths,
              int n_samples, int n_frames, const float *__restrict__ filters, int n_mels,
              const int *__restrict__ mel_lo, const int *__restrict__ mel_hi, float *__restrict__ out,
              uint32_t *__restrict__ umax) {
    extern __shared__ __align__(16) float sm[];
    float *s_x = sm;                                        // [SPAN]
    float2 *s_a = reinterpret_cast<float2 *>(s_x + SPAN);   // [FR][SF] stage-1 output, later power
    float2 *s_b = s_a + FR * SF;                            // [FR][200] FFT output
    float *s_p = reinterpret_cast<float *>(s_a);            // [FR][201] power spectrum (aliases s_a)
    float *s_win = reinterpret_cast<float *>(s_b + FR * HALF);          // [400]
    float2 *s_w200 = reinterpret_cast<float2 *>(s_win + NFFT);          // [200]
    float2 *s_w400 = s_w200 + HALF;                                     // [201]
    __shared__ float s_red[LM_THREADS / 32];

    const int b = blockIdx.y, f0 = blockIdx.x * FR, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int len = n_samples;
    if (lengths != nullptr) len = min(len, lengths[b]);
    len = (int)min((int64_t)len, audio_stride);
    const float *pa = audio + (int64_t)b * audio_stride;

    // 0. twiddle / window tables -> shared memory
    for (int i = tid; i < NFFT; i += LM_THREADS) s_win[i] = g_tables.win[i];
    for (int i = tid; i < HALF; i += LM_THREADS) s_w200[i] = g_tables.w200[i];
    for (int i = tid; i < NBIN; i += LM_THREADS) s_w400[i] = g_tables.w400[i];

    // 1. samples f0*160-200 .. +2800 of the zero-padded, reflect-padded signal: 16-byte loads in
    //    the interior, scalar loads where reflection / padding / the utterance end interferes
    const int base = f0 * HOP - NFFT / 2;                   // multiple of 8
    const bool vec_base = ((reinterpret_cast<uintptr_t>(pa) & 15) == 0);
    for (int i4 = tid; i4 < SPAN / 4; i4 += LM_THREADS) {
        const int j = base + 4 * i4;
        float4 v;
        if (vec_base && j >= 0 && j + 4 <= len) {
            v = __ldg(reinterpret_cast<const float4 *>(pa + j));
        } else {
            float t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int jj = j + u;
                if (jj < 0) jj = -jj;
                if (jj >= n_samples) jj = 2 * (n_samples - 1) - jj;
                t[u] = (jj >= 0 && jj < len) ? __ldg(pa + jj) : 0.0f;
            }
            v = make_float4(t[0], t[1], t[2], t[3]);
        }
        reinterpret_cast<float4 *>(s_x)[i4] = v;
    }
    __syncthreads();

    // 2. 200-point complex FFT of z[n] = w[2n] x[2n] + i w[2n+1] x[2n+1], n = 20*n1 + n2:
    //    stage 1 = 10-point DFTs over n1 (for each n2), times W200^(n2*k1)
    for (int task = tid; task < FR * 20; task += LM_THREADS) {
        const int f = task / 20, n2 = task - f * 20;
        const float2 *xf = reinterpret_cast<const float2 *>(s_x + f * HOP);
        const float2 *wf = reinterpret_cast<const float2 *>(s_win);
        float2 in[10];
#pragma unroll
        for (int n1 = 0; n1 < 10; ++n1) {
            const float2 xv = xf[20 * n1 + n2], wv = wf[20 * n1 + n2];
            in[n1] = make_float2(xv.x * wv.x, xv.y * wv.y);
        }
        float2 y[10];
        dft10(in, y);
#pragma unroll
        for (int k1 = 0; k1 < 10; ++k1) s_a[f * SF + k1 * S1 + n2] = cmul(y[k1], s_w200[n2 * k1]);
    }
    __syncthreads();
    //    stage 2 = 20-point DFTs over n2 (for each k1): Z[k1 + 10*k2]
    for (int task = tid; task < FR * 10; task += LM_THREADS) {
        const int f = task / 10, k1 = task - f * 10;
        float2 in[20];
#pragma unroll
        for (int n2 = 0; n2 < 20; ++n2) in[n2] = s_a[f * SF + k1 * S1 + n2];
        float2 z[20];
        dft20(in, z);
#pragma unroll
        for (int k2 = 0; k2 < 20; ++k2) s_b[f * HALF + k1 + 10 * k2] = z[k2];
    }
    __syncthreads();

    // 3. even/odd split -> X[k], k = 0..200, power spectrum (one warp per frame)
    for (int f = warp; f < FR; f += LM_THREADS / 32) {
        const float2 *zb = s_b + f * HALF;
        for (int k = lane; k < NBIN; k += 32) {
            const float2 zk = zb[k == HALF ? 0 : k];
            const float2 zr = zb[k == 0 ? 0 : HALF - k];
            const float2 zc = make_float2(zr.x, -zr.y);
            const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
            const float2 d = make_float2(zk.x - zc.x, zk.y - zc.y);
            const float2 o = make_float2(0.5f * d.y, -0.5f * d.x);  // d / (2i)
            const float2 t = cmul(o, s_w400[k]);
            const float re = e.x + t.x, im = e.y + t.y;
            s_p[f * NBIN + k] = re * re + im * im;
        }
    }
    __syncthreads();

    // 4. mel projection over each triangle's support, log10, store, utterance max
    float mx = -INFINITY;
    for (int task = tid; task < n_mels * FR; task += LM_THREADS) {
        const int m = task / FR, f = task - m * FR;
        if (f0 + f >= n_frames) continue;
        const int lo = mel_lo[m], hi = mel_hi[m];
        float acc = 0.0f;
        for (int k = lo; k <= hi; ++k) acc = fmaf(__ldg(filters + k * n_mels + m), s_p[f * NBIN + k], acc);
        const float v = log10f(fmaxf(acc, 1e-10f));
        out[((int64_t)b * n_mels + m) * n_frames + f0 + f] = v;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    if (lane == 0) s_red[warp] = mx;
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < LM_THREADS / 32; ++i) mx = fmaxf(mx, s_red[i]);
        if (mx > -INFINITY) atomicMax(umax + b, float_to_ordered(mx));
    }
}

__global__ void __launch_bounds__(256)
k_logmel_finalize(float *__restrict__ out, int64_t per_utt, const uint32_t *__restrict__ umax) {
    const int b = blockIdx.y;
    const float floor_v = ordered_to_float(umax[b]) - 8.0f;
    float *p = out + (int64_t)b * per_utt;
    const int64_t n4 = (per_utt % 4 == 0) ? per_utt / 4 : 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = reinterpret_cast<float4 *>(p)[i];
        v.x = (fmaxf(v.x, floor_v) + 4.0f) * 0.25f;
        v.y = (fmaxf(v.y, floor_v) + 4.0f) * 0.25f;
        v.z = (fmaxf(v.z, floor_v) + 4.0f) * 0.25f;
        v.w = (fmaxf(v.w, floor_v) + 4.0f) * 0.25f;
        reinterpret_cast<float4 *>(p)[i] = v;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_utt;
         i += (int64_t)gridDim.x * blockDim.x)
        p[i] = (fmaxf(p[i], floor_v) + 4.0f) * 0.25f;
}

int upload_tables() {
    cudaError_t err = cudaSuccess;
    // tables are per-context constants; re-upload per device on first use there
    static bool done[64] = {false};
    int dev = 0;
    WQ_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && done[dev]) return WQ_OK;
    static LmTables host;
    float2 w10[10], w20[20];
    const double two_pi = 6.283185307179586476925286766559;
    for (int n = 0; n < NFFT; ++n) host.win[n] = (float)(0.5 - 0.5 * std::cos(two_pi * n / NFFT));
    for (int j = 0; j < 10; ++j) w10[j] = make_float2((float)std::cos(two_pi * j / 10), (float)-std::sin(two_pi * j / 10));
    for (int j = 0; j < 20; ++j) w20[j] = make_float2((float)std::cos(two_pi * j / 20), (float)-std::sin(two_pi * j / 20));
    for (int j = 0; j < HALF; ++j)
        host.w200[j] = make_float2((float)std::cos(two_pi * j / HALF), (float)-std::sin(two_pi * j / HALF));
    for (int j = 0; j <= NBIN; ++j)
        host.w400[j] = make_float2((float)std::cos(two_pi * j / NFFT), (float)-std::sin(two_pi * j / NFFT));
    err = cudaMemcpyToSymbol(c_w10, w10, sizeof(w10));
    if (err == cudaSuccess) err = cudaMemcpyToSymbol(c_w20, w20, sizeof(w20));
    if (err == cudaSuccess) err = cudaMemcpyToSymbol(g_tables, &host, sizeof(host));
    WQ_CUDA(err);
    if (dev >= 0 && dev < 64) done[dev] = true;
    return WQ_OK;
}

}  // namespace

extern "C" int wq_logmel(const float *audio, int64_t B, int64_t audio_stride, const int32_t *lengths,
                         int64_t n_samples, const float *filters, int n_mels, void *out, int out_dtype,
                         uint32_t *workspace, wq_stream_t stream) {
    WQ_REQUIRE(B >= 0 && audio_stride >= 0, "wq_logmel: negative shape");
    WQ_REQUIRE(n_samples >= NFFT && n_samples % HOP == 0 && n_samples < (1ll << 30),
               "wq_logmel: n_samples=%lld must be a multiple of 160 and >= 400", (long long)n_samples);
    WQ_REQUIRE(n_mels > 0 && n_mels <= 1024, "wq_logmel: bad n_mels %d", n_mels);
    WQ_REQUIRE(out_dtype == WQ_F32, "wq_logmel: only float32 output is implemented");
    if (B == 0) return WQ_OK;
    WQ_REQUIRE(B <= 65535, "wq_logmel: B=%lld exceeds 65535 utterances per call", (long long)B);
    WQ_REQUIRE(audio && filters && out && workspace, "wq_logmel: null pointer");
    WQ_REQUIRE(wq_aligned(out, 16), "wq_logmel: out must be 16-byte aligned");
    int rc = upload_tables();
    if (rc != WQ_OK) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const int n_frames = (int)(n_samples / HOP);
    uint32_t *umax = workspace;
    int *mel_lo = reinterpret_cast<int *>(workspace + B);
    int *mel_hi = mel_lo + n_mels;
    WQ_CUDA(cudaMemsetAsync(umax, 0, sizeof(uint32_t) * B, s));
    k_mel_ranges<<<(n_mels * 32 + 127) / 128, 128, 0, s>>>(filters, n_mels, mel_lo, mel_hi);
    WQ_LAUNCH_CHECK();
    const size_t smem = sizeof(float) * SPAN + sizeof(float2) * FR * (SF + HALF) + sizeof(float) * NFFT +
                        sizeof(float2) * (HALF + NBIN + 1);
    static bool configured = false;
    if (!configured) {
        WQ_CUDA(cudaFuncSetAttribute(k_logmel_main, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid((n_frames + FR - 1) / FR, (unsigned)B);
    k_logmel_main<<<grid, LM_THREADS, smem, s>>>(audio, audio_stride, lengths, (int)n_samples, n_frames, filters,
                                                 n_mels, mel_lo, mel_hi, (float *)out, umax);
    WQ_LAUNCH_CHECK();
    const int64_t per_utt = (int64_t)n_mels * n_frames;
    dim3 grid2((unsigned)std::min<int64_t>((per_utt / 4 + 255) / 256, 256), (unsigned)B);
    k_logmel_finalize<<<grid2, 256, 0, s>>>((float *)out, per_utt, umax);
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}
