/*
 * whisperq_oracle.c -- CPU restatement of the compressed-Whisper hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ may be imported, linked or
 * executed by the product package; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg use it, and only as the
 * checker.
 *
 * Parity status
 *   - torch dynamic-int8 (orc_torch_*):     PINNED against live torch
 *     (tests/golden/make_golden.py -> tests/golden/torch_dynamic_*.npz).
 *   - log-mel: restated in numpy (oracle/logmel.py), PINNED against the live
 *     HF WhisperFeatureExtractor the reference calls (data_utils.py:56-58).
 *   - bitsandbytes NF4 / LLM.int8 and optimum-quanto qint8 (orc_nf4_*,
 *     orc_bnb_*, orc_quanto_*): "PARITY UNPINNED".  The reference holds no
 *     tests or golden vectors for them, bitsandbytes is not
 *     pinned by the reference (absent from pyproject.toml / uv.lock) and
 *     neither library is installed here, so these functions restate the
 *     libraries' published algorithms anchored on the
 *     reference call sites:
 *        model_utils.py:24-49,102-128          (bnb 4-bit config, quanto)
 *        pruning+quantization/bnb_implementation.py:1093-1118  (Linear4bit swap)
 *        pruning+quantization/quanto_implementation.py:648-670 (quanto qint8)
 *
 * All arithmetic is scalar fp32 / int32, compiled with -ffp-contract=off so
 * that no FMA is formed except where fmaf() is written explicitly.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ------------------------------------------------------------------------- */
/* NF4 / FP4 codebooks (QLoRA create_normal_map)               */
/* ------------------------------------------------------------------------- */
static const float NF4_CODE[16] = {
    -1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
    -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
    0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f,
    0.33791524171829224f, 0.44070982933044434f, 0.5626170039176941f,
    0.7229568362236023f, 1.0f};

/* bitsandbytes FP4 value table: sign bit 8, magnitudes {0,0.0625,8,12,4,6,2,3}/12 */
static const float FP4_CODE[16] = {
    0.0f, 0.0052083333f, 0.6666667f, 1.0f, 0.33333334f, 0.5f, 0.16666667f, 0.25f,
    -0.0f, -0.0052083333f, -0.6666667f, -1.0f, -0.33333334f, -0.5f, -0.16666667f, -0.25f};

const float *orc_nf4_codebook(void) { return NF4_CODE; }
const float *orc_fp4_codebook(void) { return FP4_CODE; }

/* Decision tree of bitsandbytes csrc/kernels.cu::dQuantizeNF4 -- strict '>'
 * against the midpoints of adjacent codebook entries.  NaN (0 * inf for an
 * all-zero block) falls through every comparison to code 0. */
static uint8_t nf4_encode(float x) {
    if (x > 0.03979014977812767f) {
        if (x > 0.3893125355243683f) {
            if (x > 0.6427869200706482f) return (x > 0.8614784181118011f) ? 15 : 14;
            return (x > 0.5016634166240692f) ? 13 : 12;
        }
        if (x > 0.2035212516784668f) return (x > 0.2920137718319893f) ? 11 : 10;
        return (x > 0.1202552504837513f) ? 9 : 8;
    }
    if (x > -0.33967943489551544f) {
        if (x > -0.13791173323988914f) return (x > -0.045525018125772476f) ? 7 : 6;
        return (x > -0.23460740596055984f) ? 5 : 4;
    }
    if (x > -0.6106329262256622f) return (x > -0.4599952697753906f) ? 3 : 2;
    return (x > -0.8480964004993439f) ? 1 : 0;
}

/* bitsandbytes csrc/kernels.cu::dQuantizeFP4 */
static uint8_t fp4_encode(float x) {
    int sign = x < 0 ? 8 : 0;
    x = fabsf(x);
    if (x > 0.29166667f) {
        if (x > 0.583333f) return (uint8_t)((x > 0.8333333f ? 3 : 2) + sign);
        return (uint8_t)((x > 0.4166667f ? 5 : 4) + sign);
    }
    if (x > 0.0859375f) return (uint8_t)((x > 0.20833333f ? 7 : 6) + sign);
    return (uint8_t)((x > 0.00260417f ? 1 : 0) + sign);
}

/* quantize_4bit(A, blocksize, quant_type): flatten row-major, per block absmax
 * (fp32), x * (1/absmax), two codes per byte with the FIRST element in the
 * HIGH nibble.  `w` is the source tensor converted (exactly) to fp32.
 * quant_type: 0 = nf4, 1 = fp4.  n need not be a multiple of blocksize; the
 * last block is ragged.  Odd n: the missing low nibble encodes 0.0f. */
void orc_quant_4bit(const float *w, int64_t n, int blocksize, int quant_type,
                    uint8_t *packed, float *absmax) {
    int64_t nblocks = (n + blocksize - 1) / blocksize;
    for (int64_t b = 0; b < nblocks; ++b) {
        int64_t lo = b * blocksize, hi = lo + blocksize;
        if (hi > n) hi = n;
        float am = 0.0f;
        for (int64_t i = lo; i < hi; ++i) {
            float a = fabsf(w[i]);
            if (a > am) am = a;
        }
        absmax[b] = am;
        float inv = 1.0f / am;
        for (int64_t i = lo; i < hi; i += 2) {
            float x0 = w[i] * inv;
            float x1 = (i + 1 < n) ? w[i + 1] * inv : 0.0f;
            uint8_t c0 = quant_type ? fp4_encode(x0) : nf4_encode(x0);
            uint8_t c1 = quant_type ? fp4_encode(x1) : nf4_encode(x1);
            packed[i >> 1] = (uint8_t)((c0 << 4) | c1);
        }
    }
}

/* dequantize_4bit: out[i] = code[nibble] * absmax[block] in fp32.  The caller
 * rounds to quant_state.dtype (fp16/bf16) with round-to-nearest-even. */
void orc_dequant_4bit(const uint8_t *packed, const float *absmax, int64_t n,
                      int blocksize, int quant_type, float *out) {
    const float *code = quant_type ? FP4_CODE : NF4_CODE;
    for (int64_t i = 0; i < n; ++i) {
        uint8_t byte = packed[i >> 1];
        uint8_t c = (i & 1) ? (byte & 15) : (byte >> 4);
        out[i] = code[c] * absmax[i / blocksize];
    }
}

/* bitsandbytes nested quantization of absmax (quantize_4bit(compress_statistics=True)):
 * offset = mean(absmax) [double accumulation -- documented deviation], blockwise-256 8-bit
 * quantization of absmax - offset against the 256-entry dynamic map (kernels.cu dQuantize<0>). */
static uint8_t dquantize8(const float *code, float x) {
    int pivot = 127, upper_pivot = 255, lower_pivot = 0;
    float lower = -1.0f, upper = 1.0f, val = code[pivot];
    for (int i = 64; i > 0; i >>= 1) {
        if (x > val) { lower_pivot = pivot; lower = val; pivot += i; }
        else { upper_pivot = pivot; upper = val; pivot -= i; }
        val = code[pivot];
    }
    if (upper_pivot == 255) upper = code[upper_pivot];
    if (lower_pivot == 0) lower = code[lower_pivot];
    if (x > val) {
        float midpoint = (upper + val) * 0.5f;
        return (uint8_t)(x > midpoint ? upper_pivot : pivot);
    }
    float midpoint = (lower + val) * 0.5f;
    return (uint8_t)(x < midpoint ? lower_pivot : pivot);
}

void orc_quant_absmax_double(const float *absmax, int64_t n, const float *code, uint8_t *q,
                             float *absmax2, float *offset_out, float *absmax_deq) {
    double acc = 0.0;
    for (int64_t i = 0; i < n; ++i) acc += (double)absmax[i];
    float off = (float)(acc / (double)n);
    *offset_out = off;
    for (int64_t b = 0; b * 256 < n; ++b) {
        int64_t lo = b * 256, hi = lo + 256 < n ? lo + 256 : n;
        float am = 0.0f;
        for (int64_t i = lo; i < hi; ++i) {
            float v = fabsf(absmax[i] - off);
            if (v > am) am = v;
        }
        absmax2[b] = am;
        float inv = 1.0f / am;
        for (int64_t i = lo; i < hi; ++i) {
            float v = absmax[i] - off;
            uint8_t c = dquantize8(code, v * inv);
            q[i] = c;
            float d = code[c] * am;
            absmax_deq[i] = d + off;
        }
    }
}

/* ------------------------------------------------------------------------- */
/* bitsandbytes LLM.int8                                      */
/* ------------------------------------------------------------------------- */

/* fp16 bit pattern of a float that holds an fp16 value exactly (incl. subnormals) */
static uint16_t half_bits_exact(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t sign = (u >> 16) & 0x8000u, e = (u >> 23) & 0xffu, m = u & 0x7fffffu;
    if (e == 0) return (uint16_t)sign;                 /* zero (fp32 subnormals are below fp16's range) */
    if (e == 0xff) return (uint16_t)(sign | 0x7c00u | (m ? 0x200u : 0u));
    const int he = (int)e - 127 + 15;
    if (he >= 1) return (uint16_t)(sign | ((uint32_t)he << 10) | (m >> 13));
    return (uint16_t)(sign | ((m | 0x800000u) >> (126 - (int)e)));
}

/* int8_vectorwise_quant(A, threshold): per row absmax over entries with
 * |a| < threshold (all entries when threshold == 0), q = rn(a * scale),
 * entries with |a| >= threshold stored as 0 and their column flagged.
 * `a` holds fp16 values converted to fp32.  col_flags (may be NULL) is OR-ed.
 * scale: bitsandbytes forms 127/absmax with __fdividef (approximate division,
 * csrc/kernels.cu kInt8VectorQuant), which a CPU cannot compute -- but absmax
 * is an fp16 value, so scale_table[fp16 bits of absmax] (65536 floats dumped
 * from a B200 by scripts/bnb_open_points.cu, tests/golden/fdividef_127_fp16.npz)
 * reproduces it exactly.  scale_table == NULL: the IEEE quotient (differs from
 * the approximate form in 8734 of 5.0e8 fp16 (absmax, a) pairs). */
void orc_bnb_int8_vectorwise_quant(const float *a, int64_t rows, int64_t cols,
                                   float threshold, int8_t *out, float *row_stats,
                                   uint8_t *col_flags, const float *scale_table) {
    for (int64_t r = 0; r < rows; ++r) {
        const float *row = a + r * cols;
        float am = 0.0f;
        for (int64_t c = 0; c < cols; ++c) {
            float v = fabsf(row[c]);
            if (threshold > 0.0f && !(v < threshold)) continue;
            if (v > am) am = v;
        }
        row_stats[r] = am;
        float scale = scale_table ? scale_table[half_bits_exact(am)] : 127.0f / am;
        for (int64_t c = 0; c < cols; ++c) {
            float v = row[c];
            if (threshold > 0.0f && !(fabsf(v) < threshold)) {
                out[r * cols + c] = 0;
                if (col_flags) col_flags[c] = 1;
            } else {
                float q = v * scale;                 /* NaN when am == 0 */
                out[r * cols + c] = (q != q) ? 0 : (int8_t)(int)rintf(q);
            }
        }
    }
}

/* "CA[:, outlier_cols] = 0" (bitsandbytes functional.int8_vectorwise_quant) */
void orc_bnb_zero_outlier_cols(int8_t *ca, int64_t rows, int64_t cols,
                               const uint8_t *col_flags) {
    for (int64_t r = 0; r < rows; ++r)
        for (int64_t c = 0; c < cols; ++c)
            if (col_flags[c]) ca[r * cols + c] = 0;
}

/* int8 x int8 -> int32, C[M,N] = A[M,K] . B[N,K]^T  (exact) */
void orc_igemm_nt(const int8_t *a, const int8_t *b, int64_t M, int64_t N, int64_t K,
                  int32_t *c) {
    for (int64_t m = 0; m < M; ++m)
        for (int64_t n = 0; n < N; ++n) {
            int32_t acc = 0;
            const int8_t *pa = a + m * K, *pb = b + n * K;
            for (int64_t k = 0; k < K; ++k) acc += (int32_t)pa[k] * (int32_t)pb[k];
            c[m * N + n] = acc;
        }
}

/* u8 x s8 -> int32 (torch dynamic): C[M,N] = (A[M,K] - zp) . B[N,K]^T */
void orc_igemm_u8s8_nt(const uint8_t *a, int32_t zp, const int8_t *b, int64_t M,
                       int64_t N, int64_t K, int32_t *c) {
    for (int64_t m = 0; m < M; ++m)
        for (int64_t n = 0; n < N; ++n) {
            int32_t acc = 0;
            const uint8_t *pa = a + m * K;
            const int8_t *pb = b + n * K;
            for (int64_t k = 0; k < K; ++k) acc += ((int32_t)pa[k] - zp) * (int32_t)pb[k];
            c[m * N + n] = acc;
        }
}

/* int8_mm_dequant: fmaf(c32 * rowStat * colStat, 1/(127*127), bias), fp32
 * result; the caller rounds to fp16.  bias may be NULL. */
void orc_bnb_mm_dequant(const int32_t *c32, const float *row_stats,
                        const float *col_stats, const float *bias, int64_t M,
                        int64_t N, float *out) {
    const float k = 6.200012e-05f; /* MM_DEQUANT_CONST */
    for (int64_t m = 0; m < M; ++m)
        for (int64_t n = 0; n < N; ++n) {
            float v = (float)c32[m * N + n] * row_stats[m];
            v = v * col_stats[n];
            out[m * N + n] = fmaf(v, k, bias ? bias[n] : 0.0f);
        }
}

/* int8_vectorwise_dequant(CB[:, cols], SCB): CB * SCB * (1/127) in fp32 */
void orc_bnb_vectorwise_dequant(const int8_t *cb, const float *scb, int64_t N,
                                int64_t K, float *out) {
    for (int64_t n = 0; n < N; ++n)
        for (int64_t k = 0; k < K; ++k) {
            float v = (float)cb[n * K + k] * scb[n];
            out[n * K + k] = v * 7.874015718698502e-3f;
        }
}

/* ------------------------------------------------------------------------- */
/* optimum-quanto 0.2.6 qint8 weights                          */
/* ------------------------------------------------------------------------- */
/* AbsmaxOptimizer axis 0: scale[n] = max_k|W[n,k]| / 127;
 * SymmetricQuantizer: q = clamp(round_half_even(W / scale), -128, 127).
 * An all-zero row (pruned) has scale 0 and W/scale = NaN -> code 0. */
void orc_quanto_qint8(const float *w, int64_t N, int64_t K, int8_t *q, float *scale) {
    for (int64_t n = 0; n < N; ++n) {
        float am = 0.0f;
        for (int64_t k = 0; k < K; ++k) {
            float v = fabsf(w[n * K + k]);
            if (v > am) am = v;
        }
        float s = am / 127.0f;
        scale[n] = s;
        for (int64_t k = 0; k < K; ++k) {
            float r = rintf(w[n * K + k] / s);
            if (r != r) r = 0.0f;
            if (r > 127.0f) r = 127.0f;
            if (r < -128.0f) r = -128.0f;
            q[n * K + k] = (int8_t)r;
        }
    }
}

/* quanto qint4 (MaxOptimizer + AffineQuantizer, float shift), groups of `group`
 * consecutive in-features of one output channel: scale = (max - min) / 15,
 * shift = -min, q = clamp(round_half_even((w + shift) / scale), 0, 15).
 * q holds one code per element (unpacked). */
void orc_quanto_qbits(const float *w, int64_t N, int64_t K, int group, int bits, uint8_t *q,
                      float *scale, float *shift);
void orc_quanto_qint4(const float *w, int64_t N, int64_t K, int group, uint8_t *q,
                      float *scale, float *shift) {
    orc_quanto_qbits(w, N, K, group, 4, q, scale, shift);
}

/* the same for qint2 / qint4: qmax = 2^bits - 1 levels above 0 */
void orc_quanto_qbits(const float *w, int64_t N, int64_t K, int group, int bits, uint8_t *q,
                      float *scale, float *shift) {
    const float qmax = (float)((1 << bits) - 1);
    int64_t ng = K / group;
    for (int64_t n = 0; n < N; ++n)
        for (int64_t g = 0; g < ng; ++g) {
            const float *p = w + n * K + g * group;
            float mn = p[0], mx = p[0];
            for (int i = 1; i < group; ++i) {
                if (p[i] < mn) mn = p[i];
                if (p[i] > mx) mx = p[i];
            }
            float s = (mx - mn) / qmax, sh = -mn;
            scale[n * ng + g] = s;
            shift[n * ng + g] = sh;
            for (int i = 0; i < group; ++i) {
                float r = rintf((p[i] + sh) / s);
                if (r != r) r = 0.0f;
                if (r < 0.0f) r = 0.0f;
                if (r > qmax) r = qmax;
                q[n * K + g * group + i] = (uint8_t)r;
            }
        }
}

/* dequantize: scale * q - shift (fp32, two operations) */
void orc_quanto_qint4_dequant(const uint8_t *q, const float *scale, const float *shift,
                              int64_t N, int64_t K, int group, float *out) {
    int64_t ng = K / group;
    for (int64_t n = 0; n < N; ++n)
        for (int64_t k = 0; k < K; ++k) {
            float v = scale[n * ng + k / group] * (float)q[n * K + k];
            out[n * K + k] = v - shift[n * ng + k / group];
        }
}

/* ------------------------------------------------------------------------- */
/* torch dynamic int8 (torch/ao/quantization/observer.py       */
/* MinMaxObserver per_tensor_symmetric; quantize_per_tensor; FBGEMM           */
/* ChooseQuantizationParams with reduce_range)                                */
/* ------------------------------------------------------------------------- */
void orc_torch_weight_qint8(const float *w, int64_t n, int8_t *q, float *scale_out) {
    float mn = 0.0f, mx = 0.0f;
    if (n > 0) { mn = w[0]; mx = w[0]; }
    for (int64_t i = 1; i < n; ++i) {
        if (w[i] < mn) mn = w[i];
        if (w[i] > mx) mx = w[i];
    }
    float min_neg = mn < 0.0f ? mn : 0.0f, max_pos = mx > 0.0f ? mx : 0.0f;
    float m = (-min_neg > max_pos) ? -min_neg : max_pos;
    float scale = m / 127.5f;
    if (scale < FLT_EPSILON) scale = FLT_EPSILON;
    *scale_out = scale;
    float inv = 1.0f / scale;
    for (int64_t i = 0; i < n; ++i) {
        float r = nearbyintf(w[i] * inv);
        if (r > 127.0f) r = 127.0f;
        if (r < -128.0f) r = -128.0f;
        q[i] = (int8_t)r;
    }
}

/* Dynamic activation parameters over the whole tensor, quint8, reduce_range
 * (qmin 0, qmax 127).  Follows FBGEMM ChooseQuantizationParams. */
void orc_torch_act_qparams(float mn, float mx, int reduce_range, float *scale_out,
                           int32_t *zp_out) {
    int32_t qmin = 0, qmax = reduce_range ? 127 : 255;
    if (mn > 0.0f) mn = 0.0f;
    if (mx < 0.0f) mx = 0.0f;
    float scale = (float)(((double)mx - (double)mn) / (double)(qmax - qmin));
    if (scale == 0.0f || isinf(1.0f / scale)) scale = 0.1f;
    double zp_from_min = qmin - (double)mn / (double)scale;
    double zp_from_max = qmax - (double)mx / (double)scale;
    double err_min = fabs((double)qmin) + fabs((double)mn / (double)scale);
    double err_max = fabs((double)qmax) + fabs((double)mx / (double)scale);
    double init = err_min < err_max ? zp_from_min : zp_from_max;
    int32_t zp;
    if (init < qmin) zp = qmin;
    else if (init > qmax) zp = qmax;
    else zp = (int32_t)nearbyint(init);
    *scale_out = scale;
    *zp_out = zp;
}

void orc_torch_act_quant(const float *x, int64_t n, int reduce_range, uint8_t *q,
                         float *scale_out, int32_t *zp_out) {
    float mn = 0.0f, mx = 0.0f;
    if (n > 0) { mn = x[0]; mx = x[0]; }
    for (int64_t i = 1; i < n; ++i) {
        if (x[i] < mn) mn = x[i];
        if (x[i] > mx) mx = x[i];
    }
    orc_torch_act_qparams(mn, mx, reduce_range, scale_out, zp_out);
    float inv = 1.0f / *scale_out;
    int32_t zp = *zp_out;
    for (int64_t i = 0; i < n; ++i) {
        float r = nearbyintf(x[i] * inv) + (float)zp;
        if (r < 0.0f) r = 0.0f;
        if (r > 255.0f) r = 255.0f;
        q[i] = (uint8_t)r;
    }
}

/* y = acc * (s_x * s_w) + bias (fp32) */
void orc_torch_requant(const int32_t *acc, float sx, float sw, const float *bias,
                       int64_t M, int64_t N, float *out) {
    float s = sx * sw;
    for (int64_t m = 0; m < M; ++m)
        for (int64_t n = 0; n < N; ++n) {
            float v = (float)acc[m * N + n] * s;
            out[m * N + n] = bias ? v + bias[n] : v;
        }
}

/* ------------------------------------------------------------------------- */
/* fp32 GEMM oracle: C[M,N] = A[M,K] . B[N,K]^T with double accumulation      */
/* ------------------------------------------------------------------------- */
void orc_gemm_nt_f64acc(const float *a, const float *b, int64_t M, int64_t N,
                        int64_t K, double *c) {
    for (int64_t m = 0; m < M; ++m)
        for (int64_t n = 0; n < N; ++n) {
            double acc = 0.0;
            const float *pa = a + m * K, *pb = b + n * K;
            for (int64_t k = 0; k < K; ++k) acc += (double)pa[k] * (double)pb[k];
            c[m * N + n] = acc;
        }
}

/* ------------------------------------------------------------------------- */
/* Levenshtein tallies: unit-cost edit distance on int ids    */
/* ------------------------------------------------------------------------- */
int64_t orc_edit_distance(const int32_t *ref, int64_t nr, const int32_t *hyp, int64_t nh) {
    int64_t *row = (int64_t *)malloc((size_t)(nh + 1) * sizeof(int64_t));
    for (int64_t j = 0; j <= nh; ++j) row[j] = j;
    for (int64_t i = 1; i <= nr; ++i) {
        int64_t diag = row[0];
        row[0] = i;
        for (int64_t j = 1; j <= nh; ++j) {
            int64_t up = row[j];
            int64_t best = diag + (ref[i - 1] != hyp[j - 1]);
            if (up + 1 < best) best = up + 1;
            if (row[j - 1] + 1 < best) best = row[j - 1] + 1;
            diag = up;
            row[j] = best;
        }
    }
    int64_t d = row[nh];
    free(row);
    return d;
}
