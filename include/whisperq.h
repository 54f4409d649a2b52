/*
 * whisperq.h -- C ABI of libwhisperq.so, the B200 (sm_100a) implementation of the
 * compressed-Whisper hot path (quantized / pruned linear layers + log-mel frontend).
 *
 * The reference (juligoat/openai-whisper-compression) is pure Python; its "FFI" for this path
 * is the set of third-party entry points its module swaps reach.  Each function below names the
 * interface it replaces (reference file:line of the call site, and the library routine behind
 * it).  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch)
 *     owns all buffers, the library never frees or retains them beyond the call;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous and ordered on it;
 *   - return value 0 = success, otherwise a wq_status code; wq_last_error() gives the text
 *     (thread-local).  No exceptions or aborts cross the ABI.  There is no CPU fallback.
 *   - matrices are row-major; "NT" GEMMs compute Y[M,N] = A[M,K] . W[N,K]^T.
 */
#ifndef WHISPERQ_H
#define WHISPERQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *wq_stream_t;

enum wq_status {
    WQ_OK = 0,
    WQ_ERR_INVALID = 1,     /* bad argument (shape, alignment, dtype enum)   */
    WQ_ERR_CUDA = 2,        /* a CUDA runtime / driver call failed           */
    WQ_ERR_UNSUPPORTED = 3  /* device is not sm_100 or feature not built     */
};

enum wq_dtype { WQ_F32 = 0, WQ_F16 = 1, WQ_BF16 = 2 };
enum wq_quant_type { WQ_NF4 = 0, WQ_FP4 = 1 };

const char *wq_last_error(void);
int wq_version(void);
/* sm count and compute capability of the current device */
int wq_device_info(int *sm_count, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------------------------
 * Weight packing / quantization (one-off at load; bit-exact against the oracle)
 * ---------------------------------------------------------------------------------------- */

/* bitsandbytes.functional.quantize_4bit -- reached from Params4bit.to(device)
 * (pruning+quantization/bnb_implementation.py:1106-1116,1221; model_utils.py:24-49,112-118).
 * w: n elements of w_dtype, flattened row-major.  packed: (n+1)/2 bytes, first element in the
 * high nibble.  absmax: ceil(n/blocksize) fp32.  blocksize: power of two in [64, 4096]. */
int wq_quant_4bit(const void *w, int w_dtype, int64_t n, int blocksize, int quant_type,
                  uint8_t *packed, float *absmax, wq_stream_t stream);

/* bitsandbytes.functional.dequantize_4bit (Linear4bit.forward; HF dequantize path
 * transformers/integrations/bitsandbytes.py:249): out[i] = code[nibble] * absmax[block],
 * rounded once to out_dtype. */
int wq_dequant_4bit(const uint8_t *packed, const float *absmax, int64_t n, int blocksize,
                    int quant_type, void *out, int out_dtype, wq_stream_t stream);

/* Nested ("double") quantization of the 4-bit statistics -- quantize_4bit(compress_statistics=
 * True), i.e. bnb_4bit_use_double_quant (model_utils.py:42,48 "bnb_*_double"):
 *   offset = mean(absmax); blockwise (256) 8-bit quantization of absmax - offset against the
 *   256-entry dynamic code book (binary search + nearest neighbour as csrc/kernels.cu dQuantize).
 * absmax: fp32 [n] (unchanged); code256: fp32 [256]; q: uint8 [n]; absmax2: fp32 [ceil(n/256)];
 * offset: fp32 [1]; absmax_deq: fp32 [n] = code[q] * absmax2 + offset, the statistics that
 * dequantize_4bit / the fused GEMM use from then on.  Deviation: the mean is accumulated in double
 * (bitsandbytes: torch.mean in fp32) so that the CPU oracle can reproduce it bit for bit. */
int wq_quant_absmax_double(const float *absmax, int64_t n, const float *code256, uint8_t *q,
                           float *absmax2, float *offset, float *absmax_deq, wq_stream_t stream);

/* dequantize_blockwise(absmax, state2) + offset: rebuilds the fp32 statistics from a state dict. */
int wq_dequant_absmax_double(const uint8_t *q, const float *absmax2, const float *code256,
                             const float *offset, int64_t n, float *absmax_out, wq_stream_t stream);

/* bitsandbytes.functional.int8_vectorwise_quant -- Int8Params.to(device) for weights
 * (threshold 0) and MatMul8bitLt.forward for activations (threshold 6.0); BASELINE.json
 * config 2.  a: fp16 [rows, cols].  out int8 [rows, cols], row_stats fp32 [rows].
 * threshold > 0: entries with |a| >= threshold are written as 0, excluded from the row absmax,
 * col_flags[c] is set to 1 and col_flags[cols] ("any outlier") is set to 1.  col_flags: int32
 * [cols + 2], all zero on entry (word cols + 1 is a completion counter owned by wq_gemm_llmint8);
 * may be NULL when threshold == 0. */
int wq_quant_i8_rowwise_bnb(const void *a_f16, int64_t rows, int64_t cols, float threshold,
                            int8_t *out, float *row_stats, int32_t *col_flags,
                            wq_stream_t stream);

/* Stand-alone outlier bookkeeping of int8_vectorwise_quant (returns the library's exact CA and
 * column list) without a host sync: compacts col_flags into outlier_cols[0..*n_outliers)
 * (ascending), clears col_flags (all cols + 2 words), then zeroes CA[:, outlier_cols].
 * n_outliers: device int32[1].  The module forward path does NOT need this call: wq_gemm_llmint8
 * consumes the flags directly. */
int wq_outlier_columns(int32_t *col_flags, int64_t rows, int64_t cols, int8_t *ca,
                       int32_t *outlier_cols, int32_t *n_outliers, wq_stream_t stream);

/* optimum.quanto quantize(model, weights=qint8); freeze(model) -- model_utils.py:126-128,
 * pruning+quantization/quanto_implementation.py:648-670.  Per output channel:
 * scale[n] = max|W[n,:]| / 127 (fp32), q = clamp(rint(W / scale), -128, 127).  w: [N, K]. */
int wq_quant_i8_rowwise_quanto(const void *w, int w_dtype, int64_t N, int64_t K, int8_t *q,
                               float *scale, wq_stream_t stream);

/* optimum.quanto quantize(model, weights=qint4); freeze(model) -- model_utils.py:126-128
 * ("quanto_int4", quantization.py:45-47).  MaxOptimizer + AffineQuantizer per group of `group`
 * consecutive in-features of one output channel (quanto: 128, reduced in steps of 32 until it
 * divides K; the whole row when K <= 128): scale = (max - min) / 15, shift = -min,
 * q = clamp(rint((w + shift) / scale), 0, 15).  packed: N*K/2 bytes, first code in the high nibble;
 * scale / shift: fp32 [N, K/group].  NB exact zeros are not preserved by this scheme. */
int wq_quant_u4_group_quanto(const void *w, int w_dtype, int64_t N, int64_t K, int group,
                             uint8_t *packed, float *scale, float *shift, wq_stream_t stream);

/* The same quantizer for weights=qint2 or qint4 (bits = 2 | 4) -- quantize(model, weights=qint2),
 * quantization/evaluation_scripts/dynamic_evaluation_int2.py:158-160: scale = (max - min) / (2^bits - 1),
 * q = clamp(rint((w + shift) / scale), 0, 2^bits - 1).  2-bit codes are stored in the SAME container as 4-bit
 * ones (one code per nibble, high nibble first), so wq_gemm_u4a16 consumes them unchanged; the arithmetic is
 * quanto's qint2, the storage is 4 bits per weight (quanto packs four codes per byte). */
int wq_quant_ubits_group_quanto(const void *w, int w_dtype, int64_t N, int64_t K, int group, int bits,
                                uint8_t *packed, float *scale, float *shift, wq_stream_t stream);

/* torch.quantization.quantize_dynamic weight observer + quantize_per_tensor --
 * model_utils.py:131-134, pruning+quantization/pytorch_implementation.py:657-665.
 * scale = max(-min, max) / 127.5 (>= FLT_EPSILON), q = clamp(nearbyint(w * (1/scale))).
 * w fp32 [N, K]; scale: device fp32[1]; wsum: int32 [N] = sum_k q[n, k] (for the u8 zero
 * point correction); workspace: device fp32[2] scratch. */
int wq_quant_i8_tensor_torch(const float *w, int64_t N, int64_t K, int8_t *q, float *scale,
                             int32_t *wsum, float *workspace, wq_stream_t stream);

/* torch.ao.nn.quantized.dynamic.Linear activation quantization (per call): per-tensor affine
 * uint8 over the whole tensor, reduce_range (0..127), FBGEMM ChooseQuantizationParams.
 * x: n elements (fp32 or fp16); q: uint8[n]; qparams: device float[2] = {scale, (float)zp};
 * workspace: device uint32[2] scratch. */
int wq_quant_act_u8_tensor(const void *x, int x_dtype, int64_t n, uint8_t *q, float *qparams,
                           uint32_t *workspace, wq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused GEMMs: TMA -> smem -> (in-register dequant) -> tcgen05.mma -> TMEM -> epilogue
 * Alignment: all matrix base pointers 16 B; K % 16 == 0 (int8 operands) / K % 64 == 0 (4-bit).
 * ---------------------------------------------------------------------------------------- */

/* bnb.matmul(x, Int8Params, state) = int8_linear_matmul + int8_mm_dequant (+ fp16 outlier
 * addmm) -- Linear8bitLt.forward, BASELINE.json config 2.
 *   y[m,n] = fp16( fmaf(int32(CA[m,:].CB[n,:]) * SCA[m] * SCB[n], 1/127^2, float(bias[n])) )
 *   outlier columns c (col_flags[c] != 0): their int8 products are removed from the accumulator
 *   (== bitsandbytes zeroing CA[:, c]) and y[m,n] = fp16( y[m,n] + sum_c A[m,c] *
 *   fp16(CB[n,c]*SCB[n]/127) ), all inside the GEMM epilogue (taken only when col_flags[K] != 0).
 * col_flags: the int32 [K + 2] array written by wq_quant_i8_rowwise_bnb for THIS activation; the
 * kernel clears it again when outliers were present.  a_f16 / col_flags may be NULL (CA already
 * has its outlier columns zeroed or threshold == 0).  bias: fp32 [N] holding the module's fp16
 * bias widened exactly (bitsandbytes converts it per element inside int8_mm_dequant), or NULL. */
int wq_gemm_llmint8(const int8_t *ca, const float *sca, const int8_t *cb, const float *scb,
                    const float *bias, void *y_f16, int64_t M, int64_t N, int64_t K,
                    const void *a_f16, int32_t *col_flags, wq_stream_t stream);

/* Same GEMM when several Linear8bitLt layers consume ONE quantized activation (HF Whisper projects the
 * encoder output through k_proj / v_proj of every decoder layer -- WhisperAttention.forward, cross-attention
 * branch -- and bitsandbytes would re-quantize it for each): keep_flags != 0 leaves col_flags set for the next
 * consumer; the last consumer passes 0 (== wq_gemm_llmint8) and clears them. */
int wq_gemm_llmint8_shared(const int8_t *ca, const float *sca, const int8_t *cb, const float *scb,
                           const float *bias, void *y_f16, int64_t M, int64_t N, int64_t K,
                           const void *a_f16, int32_t *col_flags, int keep_flags, wq_stream_t stream);

/* The same GEMM with the residual connection of the Whisper layer folded into its epilogue (SURVEY.md section 8f
 * rank 3): y = clamp(fp16(linear) + residual, -clamp_abs, clamp_abs) -- HF's `hidden_states = residual +
 * hidden_states` and the fp16 overflow clamp that close WhisperEncoderLayer.forward (modeling_whisper.py:408-414),
 * evaluated as torch does (the projection rounded to fp16, one fp16 addition, clamp).  residual_f16: [M, N] fp16,
 * contiguous (may be NULL); clamp_abs 0 = no clamp.
 * a_pre_gelu != 0: the int8 rows were produced by wq_gelu_quant with h_out == NULL (fc2 after the layer's GELU,
 * modeling_whisper.py:403-405) and a_f16 is the tensor BEFORE the activation (fc1's output); the outlier path
 * evaluates the same fp16 GELU on the entries it needs, so the result is the one the stored activation would give. */
int wq_gemm_llmint8_residual(const int8_t *ca, const float *sca, const int8_t *cb, const float *scb,
                             const float *bias, void *y_f16, int64_t M, int64_t N, int64_t K,
                             const void *a_f16, int32_t *col_flags, int keep_flags,
                             const void *residual_f16, float clamp_abs, int a_pre_gelu, wq_stream_t stream);

/* The same Linear8bitLt forward for decode-shaped calls (M <= 64 rows, 64*K + K + 272 <= 200 KiB):
 * activation quantization (threshold rule), int8 products (dp4a), int8_mm_dequant and the outlier
 * side product in ONE launch; bit-identical to wq_quant_i8_rowwise_bnb + wq_gemm_llmint8.
 * a_f16: fp16 [M, K]; bias: fp32 [N] (exact widening of the fp16 bias) or NULL; y: fp16 [M, N]. */
int wq_linear_llmint8_small(const void *a_f16, int64_t M, int64_t K, float threshold, const int8_t *cb,
                            const float *scb, const float *bias, void *y_f16, int64_t N,
                            wq_stream_t stream);

/* quanto QLinear.forward, weights-only qint8 (W8A16) -- model_utils.py:126-128 call sites:
 *   y = matmul(x, Wq.to(x.dtype).t()) * scale + bias, accumulated in fp32, rounded once.
 * x: [M, K] of x_dtype (F16/BF16); wq int8 [N, K]; scale fp32 [N]; bias fp32 [N] or NULL;
 * y: [M, N] of y_dtype. */
int wq_gemm_w8a16(const void *x, int x_dtype, const int8_t *wq, const float *scale,
                  const float *bias, void *y, int y_dtype, int64_t M, int64_t N, int64_t K,
                  wq_stream_t stream);

/* bnb Linear4bit.forward (NF4/FP4, W4A16) -- bnb_implementation.py:1093-1118:
 *   y = F.linear(x, dequantize_4bit(W).to(x.dtype), bias), fp32 accumulation.
 * packed: [N*K/2] bytes, absmax fp32 [N*K/blocksize] (blocksize 64), x/y dtype F16 or BF16,
 * bias fp32 [N] or NULL. */
int wq_gemm_w4a16(const void *x, int x_dtype, const uint8_t *packed, const float *absmax,
                  int quant_type, const float *bias, void *y, int y_dtype, int64_t M, int64_t N,
                  int64_t K, wq_stream_t stream);

/* quanto QLinear.forward with weights=qint4: y = x @ (scale * q - shift)^T + bias; the dequantised
 * weight is rounded once to the operand dtype (x_dtype F16/BF16), fp32 accumulation.
 * group: multiple of 32 dividing K; K % 64 == 0; bias fp32 [N] or NULL. */
int wq_gemm_u4a16(const void *x, int x_dtype, const uint8_t *packed, const float *scale,
                  const float *shift, int group, const float *bias, void *y, int y_dtype, int64_t M,
                  int64_t N, int64_t K, wq_stream_t stream);

/* optimum-quanto float8 weights and static activation quantization (quantize(model, weights=..., activations=...)
 * inside `with Calibration():`, model_utils.py:152-214 apply_static_quantization; quantization.py:53-86).
 *
 * wq_quant_f8_rowwise_quanto: weights=qfloat8 (e4m3fn), AbsmaxOptimizer per output channel:
 *   scale[n] = max|W[n,:]| / 448, q = e4m3(W / scale) (round to nearest even, as torch's .to(float8_e4m3fn)).
 * wq_gemm_wf8a16: QLinear.forward for those weights, y = matmul(x, Wq.to(x.dtype).t()) * scale + bias.
 * wq_quant_act_static: quantize_activation(x, qtype, scale) with a calibrated per-tensor scale (device fp32[1]):
 *   qtype 0 (qint8) code = clamp(rint(x / s), -128, 127); qtype 1 (qfloat8 e4m3) code = e4m3(x / s).  Optional
 *   outputs: codes_i8 (qint8 only), grid_f16 (the code value as fp16, exact), deq (code * s in the dtype of x =
 *   ActivationQBytesTensor.dequantize()).
 * wq_gemm_w8a8: qint8 weights x qint8 activations (quanto qbytes_int_mm):
 *   y = float(int32(xq . wq^T)) * out_scale[n] + bias[n] with out_scale[n] = input_scale * weight_scale[n]. */
int wq_quant_f8_rowwise_quanto(const void *w, int w_dtype, int64_t N, int64_t K, uint8_t *q, float *scale,
                               wq_stream_t stream);
int wq_gemm_wf8a16(const void *x, int x_dtype, const uint8_t *wq, const float *scale, const float *bias,
                   void *y, int y_dtype, int64_t M, int64_t N, int64_t K, wq_stream_t stream);
int wq_quant_act_static(const void *x, int x_dtype, int64_t n, const float *scale, int qtype,
                        int8_t *codes_i8, void *grid_f16, void *deq, wq_stream_t stream);
int wq_gemm_w8a8(const int8_t *xq, const int8_t *wq, const float *out_scale, const float *bias, void *y,
                 int y_dtype, int64_t M, int64_t N, int64_t K, wq_stream_t stream);

/* Decode-shaped forward (at most 32 rows) of the weight-only quantized linears -- bitsandbytes' own split: gemv_4bit
 * for decode, dequantize + GEMM otherwise (Linear4bit.forward) -- for every weight-only scheme here:
 *   mode 0  bnb NF4 / FP4 (quant_type), packed nibbles [N, K/2], s0 = absmax fp32 [N, K/64]
 *   mode 1  quanto qint8, int8 [N, K], s0 = scale fp32 [N] (applied after the accumulation, as QLinear does)
 *   mode 2  quanto qint4 / qint2, packed nibbles [N, K/2], s0 = scale, s1 = shift fp32 [N, K/group]
 *   mode 3  quanto qfloat8 (e4m3fn codes) [N, K], s0 = scale fp32 [N]
 * y[m, n] = sum_k x[m, k] * w[n, k] (+ scale) + bias[n]: the packed weights are streamed once, dequantized in registers
 * to the value wq_gemm_w4a16 / w8a16 / u4a16 / wf8a16 feed the tensor core (rounded once to the activation dtype where
 * the scheme rounds), fp32 accumulation on the CUDA cores, warp reduction in a fixed order.  x, y: F16 or BF16; or
 * x_dtype WQ_F32 (the reference's fp32 flows, model_utils.py:139-142): x and y are fp32, the rows are rounded to fp16 as
 * they are staged -- the operand the tensor-core GEMMs of the same flow multiply -- and the sums are stored unrounded. */
int wq_gemv_weightonly(const void *x, int x_dtype, int64_t M, int64_t K, int mode, const void *w,
                       const float *s0, const float *s1, int group, int quant_type, const float *bias,
                       void *y, int64_t N, wq_stream_t stream);

/* torch.ao.nn.quantized.dynamic.Linear.forward GPU twin (quantized::linear_dynamic):
 *   y[m,n] = float(acc[m,n] - zp * wsum[n]) * (s_x * s_w) + bias[n], fp32
 * xq uint8 [M, K]; qparams device float[2] = {s_x, zp}; wq int8 [N, K]; w_scale device fp32[1];
 * wsum int32 [N]; bias fp32 [N] or NULL; y fp32 [M, N]. */
int wq_gemm_dyn_i8(const uint8_t *xq, const float *qparams, const int8_t *wq,
                   const float *w_scale, const int32_t *wsum, const float *bias, float *y,
                   int64_t M, int64_t N, int64_t K, wq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Log-mel frontend -- WhisperFeatureExtractor.__call__ as used by data_utils.py:55-59
 * ---------------------------------------------------------------------------------------- */
/* audio: fp32 [B, audio_stride]; lengths: int32 [B] valid samples per utterance (NULL = all
 * n_samples); samples beyond the length are zero padding, utterances are truncated to
 * n_samples.  filters: fp32 [201, n_mels].  out: fp32 or fp16 [B, n_mels, n_samples/160].
 * Only WQ_F32 output is implemented.  workspace: device uint32[B + 2*n_mels] scratch. */
int wq_logmel(const float *audio, int64_t B, int64_t audio_stride, const int32_t *lengths,
              int64_t n_samples, const float *filters, int n_mels, void *out, int out_dtype,
              uint32_t *workspace, wq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Producers fused with the LLM.int8 activation quantizer.
 * HF Whisper feeds every quantized linear from a LayerNorm, a GELU or an attention output
 * (transformers modeling_whisper.py: WhisperEncoderLayer / WhisperDecoderLayer.forward,
 * WhisperAttention.forward) and bitsandbytes' Linear8bitLt.forward then quantizes that tensor in a
 * launch of its own (int8_vectorwise_quant).  These entry points write the fp16 tensor HF would have
 * produced and, when `ca` is not NULL, its int8 rows + row absmax + outlier column flags exactly as
 * wq_quant_i8_rowwise_bnb would from that tensor (same flag protocol: col_flags int32[cols + 2]).
 * ---------------------------------------------------------------------------------------- */
/* x_out = x + delta (skipped when delta is NULL; x_out may alias x); h_out = LayerNorm(x_out) * gamma +
 * beta, fp32 statistics, eps as nn.LayerNorm.  dtype WQ_F16 / WQ_BF16 for x, delta, gamma, beta, x_out,
 * h_out.  cols % 8 == 0, cols <= 2048.  int8 outputs only with WQ_F16. */
int wq_add_layernorm_quant(const void *x, const void *delta, int dtype, const void *gamma,
                           const void *beta, float eps, int64_t rows, int64_t cols, void *x_out,
                           void *h_out, float threshold, int8_t *ca, float *row_stats,
                           int32_t *col_flags, wq_stream_t stream);

/* h_out = gelu(x) (erf form, torch approximate="none"), fp32 math.  cols % 8 == 0.  h_out may be NULL when `ca` is
 * not: only the int8 rows (and row_stats / col_flags) are produced -- see wq_gemm_llmint8_residual, a_pre_gelu. */
int wq_gelu_quant(const void *x, int dtype, int64_t rows, int64_t cols, void *h_out, float threshold,
                  int8_t *ca, float *row_stats, int32_t *col_flags, wq_stream_t stream);

/* Decoder self-attention for one new token per utterance (WhisperAttention.forward with a KV cache,
 * q_len = 1).  q, k, v: this step's projections, rows of H*64 elements, row stride `ld` elements (so
 * the three can be column blocks of one fused [B, 3*H*64] projection).  q is multiplied by `scaling`
 * and rounded to `dtype` first, as HF does.  k_cache / v_cache: [B, t_max, H*64]; the new k/v rows are
 * written at position *pos (device scalar), then softmax(q K^T) V over positions 0..*pos in fp32.
 * out: [B, H*64].  Optional int8 row quantization of out as above.  head_dim is 64 (all Whisper sizes).
 * dtype: WQ_F16 / WQ_BF16, or WQ_F32 for the reference's fp32 flows (quanto / bnb *_32 on a model that was never
 * .half()-ed, model_utils.py:139-142, where HF runs torch's fp32 SDPA); the int8 outputs exist for WQ_F16 only. */
int wq_self_attn_decode(const void *q, const void *k, const void *v, int64_t ld, int dtype, float scaling,
                        void *k_cache, void *v_cache, int64_t B, int H, int t_max, const int64_t *pos,
                        void *out, float threshold, int8_t *ca, float *row_stats, int32_t *col_flags,
                        wq_stream_t stream);

/* Decoder cross-attention for one new token per utterance (WhisperAttention.forward, cross-attention branch
 * with cached K/V, q_len = 1): out[b] = softmax((q[b] * scaling) K[b]^T) V[b] per head over S encoder positions,
 * fp32 running softmax.  q: [B, H*64] rows `ldq` elements apart, multiplied by `scaling` and rounded to `dtype`
 * first (as HF does).  k, v: [B, S, H*64] with rows `ld` elements apart (a layer's K and V may be the two column
 * blocks of one [B, S, 2*H*64] projection).  out: [B, H*64] contiguous.  HBM-bound: 2*S*H*64*sizeof(dtype) bytes
 * per utterance.  head_dim is 64.  Optional int8 row quantization of out (ca != NULL, as above); it then needs
 * row_counters: device int32[B], zero before the first call (the kernel leaves it zero).
 * dtype: WQ_F16 / WQ_BF16 (TMA-fed walk), or WQ_F32 for the reference's fp32 flows (register-fed walk, 32-byte row
 * chunks; no int8 outputs). */
int wq_cross_attn_decode(const void *q, int64_t ldq, int dtype, float scaling, const void *k, const void *v,
                         int64_t ld, int64_t B, int64_t S, int H, void *out, float threshold, int8_t *ca,
                         float *row_stats, int32_t *col_flags, int32_t *row_counters, wq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * One persistent kernel for everything a WhisperDecoderLayer (bitsandbytes LLM.int8 linears, fp16) does to ONE new
 * token per utterance between two cross-attention passes (transformers modeling_whisper.py:458-520, reached from
 * data_utils.py:152 through generate): `run_after` = cross out_proj + residual, LayerNorm, fc1, GELU, fc2 + residual of
 * layer `after`; `run_before` = LayerNorm, q|k|v, self-attention over the KV cache (appending this token), out_proj +
 * residual, LayerNorm, cross q_proj of layer `before`; `run_final` = the decoder's closing LayerNorm.  Phases are
 * separated by a grid-wide barrier; arithmetic is that of wq_add_layernorm_quant, wq_linear_llmint8_small,
 * wq_self_attn_decode and wq_gelu_quant, operation by operation (bit-identical results).  At most 64 rows per call.
 * All pointers are device pointers; q|k|v and [k;v] weights are concatenated along N.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const int8_t *cb;        /* int8 [N, K] */
    const float *scb;        /* fp32 [N] row absmax of the weight */
    const float *bias;       /* fp32 [N] (exact widening of the fp16 bias) or NULL */
    int N, K;
} wq_decode_linear;

typedef struct {
    wq_decode_linear qkv, o, cq, co, fc1, fc2;
    const void *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;   /* fp16 [d]: self_attn / encoder_attn / final LayerNorm */
    float eps1, eps2, eps3;
    void *kcache, *vcache;   /* fp16 [rows, t_max, d] self-attention cache of the rows this call serves */
} wq_decode_layer;

typedef struct {
    int M, d, ffn, H, t_max;
    float threshold, scaling;
    const int64_t *pos;      /* device scalar: position of the token being decoded */
    void *x;                 /* fp16 [M, d] residual stream, updated in place */
    void *h, *att, *qkv, *f1, *g;          /* fp16 scratch: [M, d], [M, d], [M, 3d], [M, ffn], [M, ffn] */
    int8_t *ca_d, *ca_f;     /* int8 scratch: [3][M, d], [M, ffn] */
    float *sca;              /* fp32 scratch [4][M] */
    int32_t *flags;          /* int32 scratch [4][max(d, ffn) + 2] */
    void *q_out;             /* fp16 [M, d]: query rows for the next cross-attention (run_before) */
    const void *xa;          /* fp16 [M, d]: cross-attention output (run_after) ... */
    const int8_t *xa_ca;     /* ... its int8 rows, row absmax and outlier flags (wq_cross_attn_decode) */
    const float *xa_sca;
    int32_t *xa_flags;       /* [d + 2] or NULL; cleared after use */
    const void *lnf_g, *lnf_b;
    float epsf;
    void *hfinal;            /* fp16 [M, d] (run_final) */
    unsigned *bar;           /* uint32 [2], zero before the first call: grid barrier state */
} wq_decode_args;

/* max_ctas: upper bound of the grid (all CTAs must be resident at once; with g row groups decoding on g streams at the
 * same time pass at most 296 / g). */
int wq_decode_fused_llmint8(const wq_decode_layer *after, const wq_decode_layer *before,
                            const wq_decode_args *args, int run_after, int run_before, int run_final,
                            int max_ctas, wq_stream_t stream);

/* Greedy token choice for `rows` utterances: out[r] = argmax_c (mask[c] ? -inf : logits[r*ld + c]),
 * torch.argmax semantics (first index among equal maxima; NaN is the maximum).  mask: uint8/bool [cols] or
 * NULL.  Replaces masked_fill + argmax over the [B, vocab] logits between decode steps (the Whisper logits
 * processors the reference's generate() call installs only write -inf at length-dependent positions). */
int wq_masked_argmax(const void *logits, int dtype, int64_t rows, int64_t cols, int64_t ld,
                     const uint8_t *mask, int64_t *out, wq_stream_t stream);

/* Unquantized linear on the tcgen05 pipeline, for the vocabulary projection the HF bitsandbytes flows keep in
 * floating point (`proj_out`: transformers modeling_whisper.py:971,1081; HF get_modules_to_not_convert keeps the
 * output embedding out of replace_with_bnb_linear, so model_utils.py:112-118 leaves it fp16) with the greedy
 * choice of GenerationMixin._sample (torch.argmax over the processed logits, reached from data_utils.py:152)
 * folded into the epilogue.
 *   y[m, n] = x[m, :] . w[n, :] + bias[n]   fp32 accumulation, rounded once to y_dtype
 * x: [M, K] of x_dtype (F16 / BF16); w: [N, K] of x_dtype; bias fp32 [N] or NULL; y: [M, N] of y_dtype (= x_dtype
 * or F32) with a row pitch of ldy elements (0 = N), or NULL when only the arg-max is wanted (the logits then never
 * reach HBM).  argmax_keys: device uint64 [M], zero on entry, or NULL; every tile folds
 * key = (order-preserving bits of the rounded y << 32) | (0xFFFFFFFF - n) into keys[m] with atomicMax -- i.e.
 * torch.argmax's rules: first index among equal maxima, NaN is the maximum; columns with mask[n] != 0 count as -inf
 * (the Whisper logits processors only ever write -inf at length-dependent positions).  mask: uint8 / bool, 16-byte
 * aligned, mask_len >= N rounded up to whole 128-column tiles (64 for decode-shaped calls), or NULL.
 * wq_argmax_finalize turns keys into token ids and zeroes the keys for the next call. */
int wq_gemm_f16(const void *x, int x_dtype, const void *w, const float *bias, void *y, int y_dtype,
                int64_t ldy, int64_t M, int64_t N, int64_t K, const uint8_t *mask, int64_t mask_len,
                unsigned long long *argmax_keys, wq_stream_t stream);
int wq_argmax_finalize(unsigned long long *keys, int64_t M, int64_t *out, wq_stream_t stream);

/* Sparse checkpoint -> dense fp32 tensor on the device: the reference's loaders rebuild
 * every pruned tensor on the host with `dense[indices] = values` (pruning/final_pruning_script/
 * global_storing_as sparse.py:468-471; torch COO tensors in pruning+quantization/bnb_implementation.py:406-441).
 * out[0..n_out) is zero-filled, then out[idx0[i] * cols + idx1[i]] = vals[i] (idx1 == NULL: idx0 are flat indices).
 * idx_bytes: 4 (int32) or 8 (int64).  err_flag: device int, set to 1 when an index falls outside [0, n_out). */
int wq_scatter_dense_f32(const void *idx0, const void *idx1, int idx_bytes, int64_t cols, const float *vals,
                         int64_t nnz, float *out, int64_t n_out, int *err_flag, wq_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * WER / CER tallies -- evaluate.load("wer"/"cer").compute, evaluation.py:110-116
 * ---------------------------------------------------------------------------------------- */
/* Unit-cost Levenshtein distance for P pairs of int32 id sequences (word ids or code points).
 * ref/hyp: concatenated ids; *_off: int64 [P+1] offsets; dist: int64 [P].  Sequences up to
 * 4096 ids. */
int wq_edit_distance(const int32_t *ref, const int64_t *ref_off, const int32_t *hyp,
                     const int64_t *hyp_off, int64_t P, int64_t *dist, wq_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WHISPERQ_H */
