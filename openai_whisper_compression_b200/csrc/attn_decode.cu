// attn_decode.cu -- decoder self-attention for ONE new token per utterance over a static KV cache
// (decoder attention at decode time).
//
// HF's WhisperAttention at decode time (modeling_whisper.py:310-355): q = q_proj(h) * scaling (rounded to the
// activation dtype), the new k/v rows are appended to the cache, softmax(q K^T) V over the positions seen so far,
// and the result feeds out_proj -- which, for bitsandbytes Linear8bitLt, quantizes it row-wise first.  Through
// torch that is: mul, 2 x index_copy, mask build, SDPA (34 us for a 17 MB cache), quantize.  Here it is one
// launch: CTA per utterance, warp per head (head_dim 64), cache kept in the projections' own [B, t_max, H*64]
// layout so the append is one 128-byte row per head; scores and softmax in fp32; optional LLM.int8 row
// quantization of the [H*64] output row (same arithmetic as k_quant_i8_rowwise_bnb, on the rounded fp16 values).
// L2-resident, latency-bound: what matters is launch count, not bandwidth.
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace {

constexpr int kHeadDim = 64;

// 8 consecutive elements of a head row: 16 bytes of fp16 / bf16, 32 bytes of fp32 (the reference's quanto flow keeps
// fp32 activations, model_utils.py:139-142: the fp32 caches go through the same kernels, two 16-byte loads per lane)
template <typename T> struct Chunk8 {
    uint4 r[sizeof(T) / 2];
};
template <typename T> __device__ __forceinline__ Chunk8<T> chunk_zero() {
    Chunk8<T> c;
#pragma unroll
    for (int i = 0; i < (int)sizeof(T) / 2; ++i) c.r[i] = make_uint4(0u, 0u, 0u, 0u);
    return c;
}
template <typename T> __device__ __forceinline__ Chunk8<T> chunk_ld(const T *p) {
    Chunk8<T> c;
#pragma unroll
    for (int i = 0; i < (int)sizeof(T) / 2; ++i) c.r[i] = reinterpret_cast<const uint4 *>(p)[i];
    return c;
}
template <typename T> __device__ __forceinline__ Chunk8<T> chunk_ldg(const T *p) {
    Chunk8<T> c;
#pragma unroll
    for (int i = 0; i < (int)sizeof(T) / 2; ++i) c.r[i] = __ldg(reinterpret_cast<const uint4 *>(p) + i);
    return c;
}
template <typename T> __device__ __forceinline__ void chunk_st(T *p, const Chunk8<T> &c) {
#pragma unroll
    for (int i = 0; i < (int)sizeof(T) / 2; ++i) reinterpret_cast<uint4 *>(p)[i] = c.r[i];
}
template <typename T> __device__ __forceinline__ float chunk_at(const Chunk8<T> &c, int j) {
    return to_f32(reinterpret_cast<const T *>(&c)[j]);
}

template <typename T>
__device__ __forceinline__ void load8(const T *p, float (&f)[8]) {
    const Chunk8<T> raw = chunk_ld(p);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = chunk_at(raw, j);
}

template <typename T>
__global__ void __launch_bounds__(1024)
k_self_attn_decode(const T *__restrict__ q, const T *__restrict__ k, const T *__restrict__ v, int64_t ld,
                   float scaling, T *__restrict__ kc, T *__restrict__ vc, int t_max,
                   const int64_t *__restrict__ pos_ptr, int H, T *__restrict__ out, float threshold,
                   int8_t *__restrict__ ca, float *__restrict__ row_stats, int32_t *__restrict__ col_flags) {
    extern __shared__ float smem[];          // [H][t_max] scores, then [H*64] output row (fp32 of rounded values)
    pdl_wait();
    const int b = blockIdx.x;
    const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 3, sub = lane & 7;  // 4 cache rows per warp step, 8 lanes x 8 dims per row
    const int d = H * kHeadDim;
    int pos = (int)*pos_ptr;
    pos = pos < t_max ? pos : t_max - 1;
    float *sc = smem + (size_t)h * t_max;
    float *orow = smem + (size_t)H * t_max;

    T *kc_b = kc + ((int64_t)b * t_max) * d + h * kHeadDim;
    T *vc_b = vc + ((int64_t)b * t_max) * d + h * kHeadDim;
    // the loops below are chains of L2 round trips (17 MB of cache, all L2-resident): pull this head's rows
    // (one 128-byte line per position) into L1 now, while q is fetched and scaled
    for (int t = lane; t < pos; t += 32) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(kc_b + (int64_t)t * d));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(vc_b + (int64_t)t * d));
    }

    // q (scaled, rounded as HF's `q_proj(h) * scaling`), this lane's 8 dims
    float q8[8];
    load8(q + (int64_t)b * ld + h * kHeadDim + sub * 8, q8);
#pragma unroll
    for (int j = 0; j < 8; ++j) q8[j] = to_f32(from_f32<T>(q8[j] * scaling));

    // append this step's k / v rows (lanes 0-7: k, lanes 8-15: v)
    if (lane < 8) {
        chunk_st(kc_b + (int64_t)pos * d + sub * 8, chunk_ld(k + (int64_t)b * ld + h * kHeadDim + sub * 8));
    } else if (lane < 16) {
        chunk_st(vc_b + (int64_t)pos * d + sub * 8, chunk_ld(v + (int64_t)b * ld + h * kHeadDim + sub * 8));
    }
    __syncwarp();

    // scores: 16 cache rows per step (4 independent 16-byte loads per lane in flight -- the loop is a chain of
    // L2 round trips, not bandwidth)
    float mx = -INFINITY;
    for (int t0 = 0; t0 <= pos; t0 += 16) {
        Chunk8<T> kr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * 4 + g;
            kr[u] = chunk_zero<T>();
            if (t <= pos) kr[u] = chunk_ld(kc_b + (int64_t)t * d + sub * 8);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * 4 + g;
            float s = 0.0f;
#pragma unroll
            for (int j = 0; j < 8; ++j) s = fmaf(q8[j], chunk_at(kr[u], j), s);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            if (t <= pos) {
                if (sub == 0) sc[t] = s;
                mx = fmaxf(mx, s);
            }
        }
    }
    mx = warp_max(mx);
    __syncwarp();
    float sum = 0.0f;
    for (int t = lane; t <= pos; t += 32) {
        const float p = expf(sc[t] - mx);
        sc[t] = p;
        sum += p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    __syncwarp();

    // P V
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t0 = 0; t0 <= pos; t0 += 16) {
        Chunk8<T> vr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * 4 + g;
            vr[u] = chunk_zero<T>();
            if (t <= pos) vr[u] = chunk_ld(vc_b + (int64_t)t * d + sub * 8);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * 4 + g;
            if (t <= pos) {
                const float p = sc[t];
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(p, chunk_at(vr[u], j), acc[j]);
            }
        }
    }
    pdl_trigger();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
        acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
    }
    if (g == 0) {
        Chunk8<T> raw;
        T *o8 = reinterpret_cast<T *>(&raw);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o8[j] = from_f32<T>(acc[j] * inv);
            if (ca != nullptr) orow[h * kHeadDim + sub * 8 + j] = to_f32(o8[j]);
        }
        chunk_st(out + (int64_t)b * d + h * kHeadDim + sub * 8, raw);
    }
    if (ca == nullptr) return;

    // LLM.int8 row quantization of the [d] output row (int8_vectorwise_quant)
    __shared__ float s_red[32];
    __syncthreads();
    const bool sparse = threshold > 0.0f;
    float am = 0.0f;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float x = fabsf(orow[c]);
        if (!sparse || x < threshold) am = fmaxf(am, x);
    }
    am = warp_max(am);
    if (lane == 0) s_red[h] = am;
    __syncthreads();
    am = 0.0f;
    for (int w = 0; w < H; ++w) am = fmaxf(am, s_red[w]);
    if (threadIdx.x == 0) row_stats[b] = am;
    const float scale = bnb_row_scale(am);
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float x = orow[c];
        int qv;
        if (sparse && !(fabsf(x) < threshold)) {
            qv = 0;
            col_flags[c] = 1;
            col_flags[d] = 1;
        } else {
            qv = __float2int_rn(__fmul_rn(x, scale));
        }
        ca[(int64_t)b * d + c] = (int8_t)qv;
    }
}

}  // namespace

extern "C" int wq_self_attn_decode(const void *q, const void *k, const void *v, int64_t ld, int dtype, float scaling,
                                   void *k_cache, void *v_cache, int64_t B, int H, int t_max, const int64_t *pos,
                                   void *out, float threshold, int8_t *ca, float *row_stats, int32_t *col_flags,
                                   wq_stream_t stream) {
    WQ_REQUIRE(B >= 0 && H >= 1 && H <= 32 && t_max >= 1, "wq_self_attn_decode: bad shape (H must be 1..32)");
    WQ_REQUIRE(dtype == WQ_F16 || dtype == WQ_BF16 || dtype == WQ_F32, "wq_self_attn_decode: dtype must be f16, bf16 or f32");
    WQ_REQUIRE(threshold >= 0.0f, "wq_self_attn_decode: negative threshold");
    if (B == 0) return WQ_OK;
    WQ_REQUIRE(q && k && v && k_cache && v_cache && pos && out, "wq_self_attn_decode: null pointer");
    WQ_REQUIRE(ld % 8 == 0 && ld >= (int64_t)H * kHeadDim, "wq_self_attn_decode: bad row stride %lld", (long long)ld);
    WQ_REQUIRE(wq_aligned(q, 16) && wq_aligned(k, 16) && wq_aligned(v, 16) && wq_aligned(k_cache, 16) &&
                   wq_aligned(v_cache, 16) && wq_aligned(out, 16),
               "wq_self_attn_decode: pointers must be 16-byte aligned");
    WQ_REQUIRE(ca == nullptr || (dtype == WQ_F16 && row_stats != nullptr),
               "wq_self_attn_decode: the int8 outputs need fp16 rows and row_stats");
    WQ_REQUIRE(ca == nullptr || threshold == 0.0f || col_flags, "wq_self_attn_decode: threshold needs col_flags");
    const size_t smem = ((size_t)H * t_max + (size_t)H * kHeadDim) * sizeof(float);
    WQ_REQUIRE(smem <= 200 * 1024, "wq_self_attn_decode: H * t_max too large for shared memory");
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == WQ_F32) {
        auto kern = k_self_attn_decode<float>;
        if (smem > 48 * 1024) WQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        WQ_LAUNCH_PDL(kern, dim3((unsigned)B), dim3(H * 32), smem, s, (const float *)q, (const float *)k, (const float *)v, ld,
                      scaling, (float *)k_cache, (float *)v_cache, t_max, pos, H, (float *)out, threshold, (int8_t *)nullptr,
                      (float *)nullptr, (int32_t *)nullptr);
    } else if (dtype == WQ_F16) {
        auto kern = k_self_attn_decode<__half>;
        if (smem > 48 * 1024) WQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        WQ_LAUNCH_PDL(kern, dim3((unsigned)B), dim3(H * 32), smem, s, (const __half *)q, (const __half *)k,
                      (const __half *)v, ld, scaling, (__half *)k_cache, (__half *)v_cache, t_max, pos, H, (__half *)out,
                      threshold, ca, row_stats, col_flags);
    } else {
        auto kern = k_self_attn_decode<__nv_bfloat16>;
        if (smem > 48 * 1024) WQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        WQ_LAUNCH_PDL(kern, dim3((unsigned)B), dim3(H * 32), smem, s, (const __nv_bfloat16 *)q,
                      (const __nv_bfloat16 *)k, (const __nv_bfloat16 *)v, ld, scaling, (__nv_bfloat16 *)k_cache,
                      (__nv_bfloat16 *)v_cache, t_max, pos, H, (__nv_bfloat16 *)out, threshold, (int8_t *)nullptr,
                      (float *)nullptr, (int32_t *)nullptr);
    }
    return WQ_OK;
}

// ---------------------------------------------------------------------------------------------
// Decoder cross-attention for one new token per utterance: softmax(q K^T) V over the S = 1500 encoder positions.
// This is the largest single cost of a decode step and pure HBM streaming: per utterance and layer the fp16 K and V
// projections of the encoder output (2 x S x d x 2 bytes; 3 MB at d = 512) are read once and nothing else moves.
// K and V stay where the k_proj / v_proj GEMM wrote them ([B, S, *] rows, `ld` elements apart; a layer's K and V
// may be column blocks of one buffer).
//
// Persistent kernel: a fixed grid (a small multiple of the SM count, at most one CTA per work item) walks the
// (utterance, head) work items, heads fastest.  The 8 warps of a CTA interleave over the positions of one item, 8
// lanes per 128-byte row, 4 rows per warp instruction, 4 instructions of K and of V in flight per lane and the next
// 4 already requested (256 B per lane outstanding); the first block of the NEXT item is requested before the current
// item's partial results are merged, so the stream never drains at an item boundary.  Flash-decoding style running
// max / sum per lane group, merged across groups (shuffles) and warps (shared memory, double-buffered by item
// parity).  Round 1 launched one 128-thread CTA per item: once heads x utterances exceeded ~1.4 waves (B = 256) the
// ragged last wave cost 5 % against cuDNN inside the step graph; the item walk has no waves.
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int kXUnroll = 4;

struct Partial {
    float m, l, acc[8];
};

__device__ __forceinline__ void partial_merge(Partial &a, float om, float ol, const float (&oacc)[8]) {
    const float m = fmaxf(a.m, om);
    const float ca = (a.m == -INFINITY) ? 0.0f : exp2f(a.m - m);
    const float cb = (om == -INFINITY) ? 0.0f : exp2f(om - m);
    a.l = a.l * ca + ol * cb;
#pragma unroll
    for (int j = 0; j < 8; ++j) a.acc[j] = a.acc[j] * ca + oacc[j] * cb;
    a.m = m;
}

// Fold kXUnroll landed (k, v) 16-byte row chunks into the running softmax of this lane group in ONE update: the
// four scores are independent chains (8 FMAs + 3 shuffles each, interleaved by the compiler), then a single
// max / rescale and four independent exponentials.  Row by row the update is a serial chain of ~250 cycles per row
// (max -> exp2 -> rescale), which left a warp latency-bound; block-wise it is ~200 cycles per FOUR rows.  valid[u]:
// the row exists (positions past S are masked).
template <typename T, int kXUnroll>
__device__ __forceinline__ void fold_rows(Partial &p, const float (&q8)[8], const Chunk8<T> (&kr)[kXUnroll],
                                          const Chunk8<T> (&vr)[kXUnroll], const bool (&valid)[kXUnroll]) {
    float sc[kXUnroll];
#pragma unroll
    for (int u = 0; u < kXUnroll; ++u) {
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(q8[j], chunk_at(kr[u], j), s);
        sc[u] = s;
    }
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) {
#pragma unroll
        for (int u = 0; u < kXUnroll; ++u) sc[u] += __shfl_xor_sync(0xffffffffu, sc[u], o);
    }
    float m = p.m;
#pragma unroll
    for (int u = 0; u < kXUnroll; ++u) {
        if (!valid[u]) sc[u] = -INFINITY;
        m = fmaxf(m, sc[u]);
    }
    if (m == -INFINITY) return;                       // nothing seen yet and nothing valid here
    const float corr = exp2f(p.m - m);                // exp2f(-inf) == 0 on the first rows
    float e[kXUnroll], l = p.l * corr;
#pragma unroll
    for (int u = 0; u < kXUnroll; ++u) {
        e[u] = exp2f(sc[u] - m);                      // masked rows: exp2f(-inf) == 0
        l += e[u];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float a = p.acc[j] * corr;
#pragma unroll
        for (int u = 0; u < kXUnroll; ++u) a = fmaf(e[u], chunk_at(vr[u], j), a);
        p.acc[j] = a;
    }
    p.l = l;
    p.m = m;
}

// LLM.int8 row quantization of out[b, :] by the last of this utterance's H head items to finish
// (int8_vectorwise_quant on the rounded values; the counter resets itself for the next launch).  Called by the ONE
// warp that wrote this item's output (lanes with g == 0 hold the stores), so no CTA-wide barrier is needed and the
// other warps are already streaming the next item (a __syncthreads + fence here cost the persistent kernel a third
// of its in-graph bandwidth: the fence waits for the CTA's prefetched loads).
template <typename T>
__device__ __forceinline__ void quantize_row_if_last(const T *out, int b, int d, int H, int lane, int g, float threshold,
                                                     int8_t *__restrict__ ca, float *__restrict__ row_stats,
                                                     int32_t *__restrict__ col_flags, int32_t *__restrict__ row_counters) {
    if (g == 0) __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) last = (atomicAdd(&row_counters[b], 1) == H - 1);
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();
    const bool sparse = threshold > 0.0f;
    const T *orow = out + (int64_t)b * d;
    // The row comes from L2 (other CTAs wrote it): every 16-byte chunk of it is requested before the first one is
    // used and stays in registers for the second pass.  (Element-wise loads in two dependent loops made this tail
    // 40 + 10 serial L2 round trips at d = 1280: ~25 us at the end of every launch.)
    constexpr int kRegChunks = 8;                    // 32 lanes x 8 chunks x 8 values: rows up to d = 2048
    const int n8 = d >> 3;                           // d is a multiple of 64
    if (n8 <= 32 * kRegChunks) {
        uint4 buf[kRegChunks];
#pragma unroll
        for (int i = 0; i < kRegChunks; ++i) {
            const int c8 = lane + 32 * i;
            buf[i] = c8 < n8 ? __ldcg(reinterpret_cast<const uint4 *>(orow) + c8) : make_uint4(0u, 0u, 0u, 0u);
        }
        float am = 0.0f;
#pragma unroll
        for (int i = 0; i < kRegChunks; ++i) {
            const T *x8 = reinterpret_cast<const T *>(&buf[i]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x = fabsf(to_f32(x8[j]));
                if (!sparse || x < threshold) am = fmaxf(am, x);
            }
        }
        am = warp_max(am);
        if (lane == 0) {
            row_stats[b] = am;
            row_counters[b] = 0;
        }
        const float scale = bnb_row_scale(am);
#pragma unroll
        for (int i = 0; i < kRegChunks; ++i) {
            const int c8 = lane + 32 * i;
            if (c8 >= n8) break;
            const T *x8 = reinterpret_cast<const T *>(&buf[i]);
            uint32_t pk[2] = {0u, 0u};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float x = to_f32(x8[j]);
                int qv;
                if (sparse && !(fabsf(x) < threshold)) {
                    qv = 0;
                    col_flags[c8 * 8 + j] = 1;
                    col_flags[d] = 1;
                } else {
                    qv = __float2int_rn(__fmul_rn(x, scale));
                }
                pk[j >> 2] |= (uint32_t)(qv & 0xff) << (8 * (j & 3));
            }
            *reinterpret_cast<uint2 *>(ca + (int64_t)b * d + c8 * 8) = make_uint2(pk[0], pk[1]);
        }
        return;
    }
    float am = 0.0f;
    for (int c = lane; c < d; c += 32) {
        const float x = fabsf(to_f32(__ldcg(orow + c)));
        if (!sparse || x < threshold) am = fmaxf(am, x);
    }
    am = warp_max(am);
    if (lane == 0) {
        row_stats[b] = am;
        row_counters[b] = 0;
    }
    const float scale = bnb_row_scale(am);
    for (int c = lane * 4; c < d; c += 128) {       // d is a multiple of 64: four codes per store
        uint32_t pk = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float x = to_f32(__ldcg(orow + c + j));
            int qv;
            if (sparse && !(fabsf(x) < threshold)) {
                qv = 0;
                col_flags[c + j] = 1;
                col_flags[d] = 1;
            } else {
                qv = __float2int_rn(__fmul_rn(x, scale));
            }
            pk |= (uint32_t)(qv & 0xff) << (8 * j);
        }
        *reinterpret_cast<uint32_t *>(ca + (int64_t)b * d + c) = pk;
    }
}

// fp32 rows are twice as wide: half the rows per block (kXU = 2) keep the same bytes in flight per lane and the same
// register budget
template <typename T, int kXWarps, int kXU = kXUnroll>
__global__ void __launch_bounds__(kXWarps * 32, 2)
k_cross_attn_decode(const T *__restrict__ q, int64_t ldq, float scaling, const T *__restrict__ kmat,
                    const T *__restrict__ vmat, int64_t ld, int S, int H, int n_items, T *out, float threshold,
                    int8_t *__restrict__ ca, float *__restrict__ row_stats, int32_t *__restrict__ col_flags,
                    int32_t *__restrict__ row_counters) {
    __shared__ float s_part[2][kXWarps][8][10];   // per item parity and warp: 8 dim-groups x (m, l, acc[8])
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 3, sub = lane & 7;
    const int d = H * kHeadDim;
    constexpr float kLog2e = 1.4426950408889634f;
    constexpr int ROWS_PER_IT = kXWarps * 4;
    constexpr int STEP = ROWS_PER_IT * kXU;
    const int t_first = warp * 4 + g;

    int item = (int)blockIdx.x;
    if (item >= n_items) return;

    Chunk8<T> kr[kXU], vr[kXU], kn[kXU], vn[kXU];
    auto fetch = [&](const T *kb, const T *vb, int t0, Chunk8<T> (&kk)[kXU], Chunk8<T> (&vv)[kXU]) {
#pragma unroll
        for (int u = 0; u < kXU; ++u) {
            const int t = t0 + u * ROWS_PER_IT;
            kk[u] = chunk_zero<T>();
            vv[u] = chunk_zero<T>();
            if (t < S) {
                kk[u] = chunk_ldg(kb + (int64_t)t * ld);
                vv[u] = chunk_ldg(vb + (int64_t)t * ld);
            }
        }
    };
    auto item_ptrs = [&](int it, const T *&kb, const T *&vb, const T *&qb) {
        const int b = it / H, h = it - b * H;
        kb = kmat + (int64_t)b * S * ld + h * kHeadDim + sub * 8;
        vb = vmat + (int64_t)b * S * ld + h * kHeadDim + sub * 8;
        qb = q + (int64_t)b * ldq + h * kHeadDim + sub * 8;
    };

    const T *kb, *vb, *qb;
    item_ptrs(item, kb, vb, qb);
    fetch(kb, vb, t_first, kr, vr);
    float q8[8];
    load8(qb, q8);
    int par = 0;
    for (;; item += (int)gridDim.x, par ^= 1) {
        const int next = item + (int)gridDim.x;
        const bool has_next = next < n_items;
        const T *kbn = kb, *vbn = vb, *qbn = qb;
        if (has_next) item_ptrs(next, kbn, vbn, qbn);
        const int b = item / H, h = item - b * H;
#pragma unroll
        for (int j = 0; j < 8; ++j) q8[j] = to_f32(from_f32<T>(q8[j] * scaling)) * kLog2e;   // scores in log2 units
        Partial p;
        p.m = -INFINITY;
        p.l = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) p.acc[j] = 0.0f;

        // the trip count is warp-uniform (full-mask shuffles inside): the loop is bounded on the block start, rows
        // past S are masked per lane
        int t0 = t_first;
        for (int tb = 0; tb < S; tb += STEP, t0 += STEP) {
            if (tb + STEP < S) fetch(kb, vb, t0 + STEP, kn, vn);          // next block of this item ...
            else if (has_next) fetch(kbn, vbn, t_first, kn, vn);          // ... or the first block of the next item
            bool valid[kXU];
#pragma unroll
            for (int u = 0; u < kXU; ++u) valid[u] = t0 + u * ROWS_PER_IT < S;
            fold_rows<T, kXU>(p, q8, kr, vr, valid);
#pragma unroll
            for (int u = 0; u < kXU; ++u) {
                kr[u] = kn[u];
                vr[u] = vn[u];
            }
        }
        if (has_next) load8(qbn, q8);      // next item's query row (L2) travels while this item is merged
        else pdl_trigger();                // this CTA's stream is over: the successor's CTAs may take their seats
        // merge the 4 lane groups of the warp (same dims, different rows)
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, p.m, o);
            const float ol = __shfl_xor_sync(0xffffffffu, p.l, o);
            float oacc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) oacc[j] = __shfl_xor_sync(0xffffffffu, p.acc[j], o);
            partial_merge(p, om, ol, oacc);
        }
        if (g == 0) {
            s_part[par][warp][sub][0] = p.m;
            s_part[par][warp][sub][1] = p.l;
#pragma unroll
            for (int j = 0; j < 8; ++j) s_part[par][warp][sub][2 + j] = p.acc[j];
        }
        __syncthreads();
        if (warp == 0 && g == 0) {
#pragma unroll
            for (int w = 1; w < kXWarps; ++w) {
                float oacc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) oacc[j] = s_part[par][w][sub][2 + j];
                partial_merge(p, s_part[par][w][sub][0], s_part[par][w][sub][1], oacc);
            }
            const float inv = 1.0f / p.l;
            Chunk8<T> raw;
            T *o8 = reinterpret_cast<T *>(&raw);
#pragma unroll
            for (int j = 0; j < 8; ++j) o8[j] = from_f32<T>(p.acc[j] * inv);
            chunk_st(out + (int64_t)b * d + h * kHeadDim + sub * 8, raw);
        }
        if constexpr (sizeof(T) == 2) {     // the int8 outputs exist for 16-bit rows only
            if (ca != nullptr && warp == 0)
                quantize_row_if_last(out, b, d, H, lane, g, threshold, ca, row_stats, col_flags, row_counters);
        }
        if (!has_next) break;
        kb = kbn;
        vb = vbn;
        qb = qbn;
    }
}

// ---------------------------------------------------------------------------------------------
// The same walk with the stream landing in SHARED memory (TMA) instead of registers.
//
// The register-fed kernel above needs ~128 KB of loads in flight per SM to run HBM at rate (measured: 64 KB/SM ->
// 4.6 TB/s, 128 KB/SM -> 6.5 TB/s at ~2 us loaded latency), and with 16-byte loads that is 60 K registers per SM: the
// SM is full, nothing of another stream can run beside it.  Here one elected thread issues cp.async.bulk.tensor loads
// of [64 positions x 128 B] boxes of K and of V (one head of one utterance: a dense 2-D box of the [B*S, d] matrix
// the projection GEMM wrote) into a 6-stage x 16 KB ring -- 96 KB in flight per SM in shared memory -- and four
// consumer warps fold the landed rows into the running softmax.  160 threads x ~64 registers: the SM keeps > 50 K
// registers and 130 KB of shared memory free, which is what lets the launch-bound rest of the decode step of the
// OTHER half-batch (fastgen's second stream: lean GEMM tiles, LayerNorm, self-attention) run under this kernel's
// HBM stream.  The ring is item-agnostic: the producer runs ahead across (utterance, head) boundaries.
// ---------------------------------------------------------------------------------------------
constexpr int kTMaxStages = 12;

// kTCW consumer warps; a stage holds 16 * kTCW positions (one TMA box of K, one of V).  Warp kTCW is the TMA producer,
// warp kTCW + 1 the finisher: the cross-CTA bookkeeping of the optional int8 row quantization (fence, counter, and
// for the utterance's last head the quantization itself) costs ~5 us of pure memory latency per item; on a streaming
// warp that stalls the whole ring (measured: 123 -> 215 us per launch), on its own warp it is free.
template <typename T, int kTCW>
__global__ void __launch_bounds__((kTCW + 2) * 32, 2)
k_cross_attn_decode_tma(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v,
                        int kTStages, const T *__restrict__ q, int64_t ldq, float scaling, int S, int H, int n_items, T *out,
                        float threshold, int8_t *__restrict__ ca, float *__restrict__ row_stats,
                        int32_t *__restrict__ col_flags, int32_t *__restrict__ row_counters) {
    using namespace wq;
    extern __shared__ __align__(1024) uint8_t xs_raw[];
    uint8_t *xs = xs_raw + ((1024u - (smem_u32(xs_raw) & 1023u)) & 1023u);
    constexpr int kTR = 16 * kTCW;                      // positions per stage
    constexpr int kOpBytes = kTR * 128;                 // bytes of K (or V) per stage
    uint8_t *sK = xs, *sV = xs + kTStages * kOpBytes;
    __shared__ uint64_t bar_full[kTMaxStages], bar_empty[kTMaxStages];
    __shared__ float s_part[2][kTCW][8][10];
    __shared__ volatile int s_done;                     // items whose output rows warp 0 has written
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 3, sub = lane & 7;
    const int d = H * kHeadDim;
    constexpr float kLog2e = 1.4426950408889634f;
    const int n_blk = (S + kTR - 1) / kTR;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        for (int s = 0; s < kTStages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], kTCW);
        }
        fence_mbar_init();
        s_done = 0;
    }
    __syncthreads();
    pdl_wait();

    if (warp == kTCW + 1) {
        // ---------------- finisher ----------------
        pdl_trigger();
        if (ca == nullptr) return;
        int handled = 0;
        for (int item = (int)blockIdx.x; item < n_items; item += (int)gridDim.x, ++handled) {
            while (s_done <= handled) __nanosleep(200);
            __syncwarp();
            quantize_row_if_last(out, item / H, d, H, lane, 0, threshold, ca, row_stats, col_flags, row_counters);
        }
        return;
    }
    if (warp == kTCW) {
        // ---------------- producer ----------------
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = (int)blockIdx.x; item < n_items; item += (int)gridDim.x) {
                const int b = item / H, h = item - b * H;
                for (int blk = 0; blk < n_blk; ++blk, ++it) {
                    const int s = it % kTStages;
                    mbar_wait(&bar_empty[s], ((it / kTStages) & 1) ^ 1);
                    mbar_arrive_expect_tx(&bar_full[s], 2 * kOpBytes);
                    tma_load_2d(sK + s * kOpBytes, &map_k, &bar_full[s], h * kHeadDim, b * S + blk * kTR);
                    tma_load_2d(sV + s * kOpBytes, &map_v, &bar_full[s], h * kHeadDim, b * S + blk * kTR);
                }
            }
        }
        __syncwarp();
        pdl_trigger();
        return;
    }

    // ---------------- consumers: warps 0..3, 16 positions of every stage each ----------------
    uint32_t it = 0;
    int par = 0;
    for (int item = (int)blockIdx.x; item < n_items; item += (int)gridDim.x, par ^= 1) {
        const int b = item / H, h = item - b * H;
        float q8[8];
        load8(q + (int64_t)b * ldq + h * kHeadDim + sub * 8, q8);
#pragma unroll
        for (int j = 0; j < 8; ++j) q8[j] = to_f32(from_f32<T>(q8[j] * scaling)) * kLog2e;   // scores in log2 units
        Partial p;
        p.m = -INFINITY;
        p.l = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) p.acc[j] = 0.0f;
        for (int blk = 0; blk < n_blk; ++blk, ++it) {
            const int s = it % kTStages;
            mbar_wait(&bar_full[s], (it / kTStages) & 1);
            const uint8_t *pk = sK + s * kOpBytes + (warp * 16 + g) * 128 + sub * 16;
            const uint8_t *pv = sV + s * kOpBytes + (warp * 16 + g) * 128 + sub * 16;
            static_assert(kXUnroll == 4, "a consumer warp takes 16 positions of a stage: 4 lane groups x 4 rows");
            Chunk8<T> kr[kXUnroll], vr[kXUnroll];
#pragma unroll
            for (int u = 0; u < kXUnroll; ++u) {
                kr[u].r[0] = *reinterpret_cast<const uint4 *>(pk + u * 4 * 128);
                vr[u].r[0] = *reinterpret_cast<const uint4 *>(pv + u * 4 * 128);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_empty[s]);      // this warp's rows of the stage are in registers
            const int t_base = blk * kTR + warp * 16 + g;
            bool valid[kXUnroll];
#pragma unroll
            for (int u = 0; u < kXUnroll; ++u) valid[u] = t_base + u * 4 < S;
            fold_rows<T, kXUnroll>(p, q8, kr, vr, valid);
        }
        if (item + (int)gridDim.x >= n_items) pdl_trigger();
        // merge the 4 lane groups of the warp (same dims, different rows)
#pragma unroll
        for (int o = 8; o <= 16; o <<= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, p.m, o);
            const float ol = __shfl_xor_sync(0xffffffffu, p.l, o);
            float oacc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) oacc[j] = __shfl_xor_sync(0xffffffffu, p.acc[j], o);
            partial_merge(p, om, ol, oacc);
        }
        if (g == 0) {
            s_part[par][warp][sub][0] = p.m;
            s_part[par][warp][sub][1] = p.l;
#pragma unroll
            for (int j = 0; j < 8; ++j) s_part[par][warp][sub][2 + j] = p.acc[j];
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTCW * 32) : "memory");      // the consumer warps only
        if (warp != 0) continue;
        if (g == 0) {
#pragma unroll
            for (int w = 1; w < kTCW; ++w) {
                float oacc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) oacc[j] = s_part[par][w][sub][2 + j];
                partial_merge(p, s_part[par][w][sub][0], s_part[par][w][sub][1], oacc);
            }
            const float inv = 1.0f / p.l;
            uint4 raw;
            T *o8 = reinterpret_cast<T *>(&raw);
#pragma unroll
            for (int j = 0; j < 8; ++j) o8[j] = from_f32<T>(p.acc[j] * inv);
            *reinterpret_cast<uint4 *>(out + (int64_t)b * d + h * kHeadDim + sub * 8) = raw;
        }
        if (ca != nullptr) {        // hand the item to the finisher warp: stores first, then the count
            __threadfence_block();
            __syncwarp();
            if (lane == 0) s_done = s_done + 1;
        }
    }
}

// CTAs per SM of the persistent grid (WQ_XATTN_CTAS = 1 | 2; 2 keeps 128 KB of loads in flight per SM)
int xattn_ctas_per_sm() {
    static const int n = [] {
        const char *e = getenv("WQ_XATTN_CTAS");
        const int v = e ? atoi(e) : 2;
        return v < 1 ? 1 : (v > 2 ? 2 : v);
    }();
    return n;
}

}  // namespace

extern "C" int wq_cross_attn_decode(const void *q, int64_t ldq, int dtype, float scaling, const void *k, const void *v,
                                    int64_t ld, int64_t B, int64_t S, int H, void *out, float threshold, int8_t *ca,
                                    float *row_stats, int32_t *col_flags, int32_t *row_counters,
                                    wq_stream_t stream) {
    WQ_REQUIRE(B >= 0 && S >= 1 && S < (1 << 30) && H >= 1 && H <= 65535 && B * H < (1ll << 30),
               "wq_cross_attn_decode: bad shape");
    WQ_REQUIRE(dtype == WQ_F16 || dtype == WQ_BF16 || dtype == WQ_F32, "wq_cross_attn_decode: dtype must be f16, bf16 or f32");
    if (B == 0) return WQ_OK;
    WQ_REQUIRE(q && k && v && out, "wq_cross_attn_decode: null pointer");
    WQ_REQUIRE(ld % 8 == 0 && ld >= (int64_t)H * kHeadDim && ldq % 8 == 0 && ldq >= (int64_t)H * kHeadDim,
               "wq_cross_attn_decode: bad row stride");
    WQ_REQUIRE(wq_aligned(q, 16) && wq_aligned(k, 16) && wq_aligned(v, 16) && wq_aligned(out, 16),
               "wq_cross_attn_decode: pointers must be 16-byte aligned");
    WQ_REQUIRE(threshold >= 0.0f, "wq_cross_attn_decode: negative threshold");
    WQ_REQUIRE(ca == nullptr || (dtype == WQ_F16 && row_stats != nullptr && row_counters != nullptr),
               "wq_cross_attn_decode: the int8 outputs need fp16 rows, row_stats and row_counters");
    WQ_REQUIRE(ca == nullptr || threshold == 0.0f || col_flags, "wq_cross_attn_decode: threshold needs col_flags");
    cudaStream_t s = (cudaStream_t)stream;
    const int n_items = (int)(B * H);
    static const bool use_tma = [] {       // WQ_XATTN=reg: the register-fed kernel (A/B measurements)
        const char *e = getenv("WQ_XATTN");
        return e == nullptr || e[0] != 'r';
    }();
    if (dtype == WQ_F32) {
        // the reference's fp32 flows (quanto, bnb *_32): register-fed walk, two 16-byte loads per lane and row
        const int cap = wq_sm_count() * 2;
        const dim3 grid((unsigned)(n_items < cap ? n_items : cap));
        WQ_LAUNCH_PDL((k_cross_attn_decode<float, 8, 2>), grid, dim3(256), 0, s, (const float *)q, ldq, scaling,
                      (const float *)k, (const float *)v, ld, (int)S, H, n_items, (float *)out, 0.0f, (int8_t *)nullptr,
                      (float *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr);
        return WQ_OK;
    }
    if (use_tma && B * S < (1ll << 31)) {
        // K and V as [B*S, H*64] matrices with a row pitch of ld elements; one box = 16 * CW positions of one head
        static const int cw = [] { const char *e = getenv("WQ_XATTN_CW"); return (e && atoi(e) == 4) ? 4 : 8; }();
        static const int stages = [] {
            const char *e = getenv("WQ_XATTN_STAGES");
            const int v = e ? atoi(e) : 3;
            return v < 2 ? 2 : (v > kTMaxStages ? kTMaxStages : v);
        }();
        static const CUtensorMapL2promotion promo = [] {
            const char *e = getenv("WQ_XATTN_PROMO");
            const int v = e ? atoi(e) : 256;
            return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B :
                   (v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
        }();
        const int box_rows = 16 * cw;
        int nst = stages;
        while (nst > 2 && 2 * nst * box_rows * 128 + 1024 > 200 * 1024) --nst;
        // WQ_XATTN_SMEM_KB pads the request (116 KB = at most one CTA of this kernel per SM, still room for a lean GEMM
        // tile of another row group; measured slower than letting two row groups' streams share an SM: 118 vs 111 ms)
        static const int pad_kb = [] { const char *e = getenv("WQ_XATTN_SMEM_KB"); return e ? atoi(e) : 0; }();
        size_t smem = (size_t)2 * nst * box_rows * 128 + 1024;
        if (smem < (size_t)pad_kb * 1024) smem = (size_t)pad_kb * 1024;
        if (smem > 201 * 1024) smem = 201 * 1024;
        const CUtensorMapDataType dt = dtype == WQ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
        CUtensorMap mk, mv;
        int rc = make_map_2d(&mk, k, dt, 2, (uint64_t)(B * S), (uint64_t)H * kHeadDim, box_rows, kHeadDim,
                             CU_TENSOR_MAP_SWIZZLE_NONE, (uint64_t)ld, promo);
        if (rc != WQ_OK) return rc;
        rc = make_map_2d(&mv, v, dt, 2, (uint64_t)(B * S), (uint64_t)H * kHeadDim, box_rows, kHeadDim,
                         CU_TENSOR_MAP_SWIZZLE_NONE, (uint64_t)ld, promo);
        if (rc != WQ_OK) return rc;
        // Grid: one CTA per SM walking the items.  With few items (<= 2.5 per SM) that walk ends in a ragged wave -- 320
        // items on 148 CTAs: 24 CTAs carry a third item while 124 SMs idle -- so up to two resident CTAs per SM share
        // them evenly instead (320 items as 160 x 2: 30.5 -> 27.9 us alone, large-v3 step 313 -> 309 ms; 192 items as
        // 192 x 1: 21.6 -> 18.1 us).  Beyond that the even split measured no better inside the multi-stream step
        // (whisper-base, 4 groups of 512 items: 256 x 2 109.3 ms, 128 x 4 107.1 ms, 148-CTA walk 106 ms per step).
        // WQ_XATTN_TMA_CTAS=1 forces the one-per-SM walk.
        static const bool even_split = [] { const char *e = getenv("WQ_XATTN_TMA_CTAS"); return !(e && atoi(e) == 1); }();
        int n_ctas = n_items < wq_sm_count() ? n_items : wq_sm_count();
        if (even_split && 2 * n_items <= 5 * wq_sm_count()) {
            const int per_cta = (n_items + 2 * wq_sm_count() - 1) / (2 * wq_sm_count());
            n_ctas = (n_items + per_cta - 1) / per_cta;
        }
        const dim3 grid((unsigned)n_ctas);
        static bool configured = false;
        if (!configured) {
            WQ_CUDA(cudaFuncSetAttribute(k_cross_attn_decode_tma<__half, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
            WQ_CUDA(cudaFuncSetAttribute(k_cross_attn_decode_tma<__half, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
            WQ_CUDA(cudaFuncSetAttribute(k_cross_attn_decode_tma<__nv_bfloat16, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
            WQ_CUDA(cudaFuncSetAttribute(k_cross_attn_decode_tma<__nv_bfloat16, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
            configured = true;
        }
#define WQ_XT_LAUNCH(T, CW, CA, RS, CF, RC, THR)                                                                        \
        WQ_LAUNCH_PDL((k_cross_attn_decode_tma<T, CW>), grid, dim3((CW + 2) * 32), smem, s, mk, mv, nst, (const T *)q, ldq,  \
                      scaling, (int)S, H, n_items, (T *)out, THR, CA, RS, CF, RC)
        if (dtype == WQ_F16) {
            if (cw == 8) WQ_XT_LAUNCH(__half, 8, ca, row_stats, col_flags, row_counters, threshold);
            else WQ_XT_LAUNCH(__half, 4, ca, row_stats, col_flags, row_counters, threshold);
        } else {
            if (cw == 8) WQ_XT_LAUNCH(__nv_bfloat16, 8, (int8_t *)nullptr, (float *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, 0.0f);
            else WQ_XT_LAUNCH(__nv_bfloat16, 4, (int8_t *)nullptr, (float *)nullptr, (int32_t *)nullptr, (int32_t *)nullptr, 0.0f);
        }
#undef WQ_XT_LAUNCH
        return WQ_OK;
    }
    const int cap = wq_sm_count() * xattn_ctas_per_sm();
    const dim3 grid((unsigned)(n_items < cap ? n_items : cap));
    if (dtype == WQ_F16) {
        WQ_LAUNCH_PDL((k_cross_attn_decode<__half, 8>), grid, dim3(256), 0, s, (const __half *)q, ldq, scaling,
                      (const __half *)k, (const __half *)v, ld, (int)S, H, n_items, (__half *)out, threshold, ca,
                      row_stats, col_flags, row_counters);
    } else {
        WQ_LAUNCH_PDL((k_cross_attn_decode<__nv_bfloat16, 8>), grid, dim3(256), 0, s, (const __nv_bfloat16 *)q, ldq,
                      scaling, (const __nv_bfloat16 *)k, (const __nv_bfloat16 *)v, ld, (int)S, H, n_items,
                      (__nv_bfloat16 *)out, 0.0f, (int8_t *)nullptr, (float *)nullptr, (int32_t *)nullptr,
                      (int32_t *)nullptr);
    }
    return WQ_OK;
}
