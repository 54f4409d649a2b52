// gemm_tc.cu -- dequant-fused NT GEMMs on the 5th-generation tensor cores.
//
//   Y[M,N] = epilogue( A[M,K] . W[N,K]^T )
//
// Persistent kernel: one CTA per SM walks output tiles.  Tile shape and schedule are template parameters picked per
// call (launch_i8 / launch_gemm; measurements in DESIGN.md section 3.1):
//   256 x BN, two row halves      two M = 128, N = BN MMAs per k-step share one W tile (BN = 128, or 64 for
//                                 decode-shaped calls); every scheme
//   128 x 256, two column halves  (COLS) one M = 128, N = 256 MMA per k-step; int8 x int8 schemes
//   round-robin                   tiles n-fastest over the grid: the CTAs sharing an A row-block run together (L2)
//   weight-stationary             (WS) K <= 512, int8 x int8: the CTA's W tile stays in shared memory, the ring streams
//                                 A alone and the producer prefetches the A row block of a later tile into L2
//   CTA pair                      (PAIR) 256 x 256 on a cluster of two CTAs: ONE tcgen05.mma.cta_group::2 per k-step
//                                 (M = 256, N = 256) issued by the leader; each CTA stages its 128 rows of A and HALF
//                                 of the W tile, and drains its own 128 x 256 accumulator; int8 x int8 schemes
// Warp roles:
//   warp 0      TMA producer: A tile (and the W tile, or the PACKED W tile) -> shared memory ring
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue into a 2-deep TMEM accumulator ring
//   warps 2..9  epilogue (2 halves x 4 TMEM lane quarters): tcgen05.ld the accumulator (one TMEM lane
//               = one output row per thread), apply the scheme's scale / bias formula, stage 32 x 128 B
//               boxes in swizzled shared memory and hand them to TMA stores -- overlaps the next tile
//   warps 10..  (W8A16: 4, W4A16: 8) expand packed weights smem -> registers -> fp16/bf16 in the
//               128-byte-swizzled UMMA operand layout; one expanded W tile feeds both 128-row halves
// Pipelines are mbarrier rings: full[s] (TMA bytes landed), bready[s] (dequantised operand
// written + proxy fence), empty[s] (tcgen05.commit: the MMAs that read stage s are done),
// tmem_full[a] / tmem_empty[a] (accumulator complete / drained).
//
// Schemes (reference call sites in include/whisperq.h):
//   EPI_LLMINT8  s8 x s8 -> s32, y = fp16(fmaf(acc*SCA[m]*SCB[n], 1/127^2, bias[n]))   (bnb)
//   EPI_DYN      u8 x s8 -> s32, y = (acc - zp*wsum[n]) * (s_x*s_w) + bias[n]          (torch)
//   EPI_W8A16    f16 x f16(int8) -> f32, y = acc*scale[n] + bias[n]                    (quanto)
//   EPI_W4A16    f16 x f16(code*absmax) -> f32, y = acc + bias[n]                       (bnb NF4)
//   EPI_PLAIN    f16 x f16 -> f32, y = acc + bias[n]; W taken straight from TMA (unquantized proj_out), optional
//                masked arg-max over the rounded outputs folded into the epilogue (greedy token choice)
#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace {
using namespace wq;

constexpr int BM = 256;            // rows of A per tile: two 128-row UMMA halves sharing one W tile
constexpr int BMH = 128;           // rows per UMMA (== TMEM lanes)
constexpr int ROW_BYTES = 128;     // bytes of K per smem row (one SW128 atom)
constexpr int UMMA_K_BYTES = 32;   // one tcgen05.mma consumes 32 bytes of K per row
constexpr int ACC_STAGES = 2;      // TMEM accumulator ring (epilogue of tile i overlaps mainloop of i+1)
// epilogue warps: 2 halves x 4 TMEM lane quarters.  The kernel also supports 16 (two warps split the
// columns of each 32-row slab); measured slower on B200 (56 vs 46 us at 96000x512x512), so 8 is used.
// LEAN tiles (decode-shaped calls with M <= 128: 128 rows, one half) run 4.
template <int BN, int BMODE, int LEAN = 0> constexpr int epi_warps() { return LEAN ? 4 : 8; }
constexpr int BOX_BYTES = 32 * 128;  // one TMA-store box: 32 rows x 128 bytes (SWIZZLE_128B)

enum AKind { A_F16 = 0, A_BF16 = 1, A_S8 = 2, A_U8 = 3 };
enum BMode { B_DIRECT = 0, B_I8 = 1, B_4BIT = 2, B_U4 = 3, B_F8 = 4 };   // B_U4: quanto group-wise affine uint4; B_F8: e4m3 codes
template <int BMODE> constexpr bool is_byte() { return BMODE == B_I8 || BMODE == B_F8; }
enum Epi { EPI_LLMINT8 = 0, EPI_W8A16 = 1, EPI_W4A16 = 2, EPI_DYN = 3, EPI_PLAIN = 4, EPI_W8A8 = 5 };

struct GemmArgs {
    int M, N, K;
    int num_kb;              // ceil(K / elements per 128-byte row)
    int tiles_m, tiles_n;
    int tma_store;           // 1: epilogue stores through smem + TMA; 0: direct global stores
    const float *row_scale;  // SCA[m]                         (LLMINT8)
    const float *col_scale;  // SCB[n] / quanto scale[n]       (LLMINT8, W8A16)
    const void *bias;        // fp32 [N] or nullptr (LLM.int8: the fp16 bias widened exactly)
    const float *absmax;     // [N, K/64] (W4A16 NF4/FP4) / group scale [N, K/group] (quanto qint4)
    int absmax_ld;           // K / 64 or K / group
    const float *shift;      // group shift [N, K/group]       (quanto qint4)
    int group;               // quanto group size (multiple of 32)
    const float *qparams;    // {s_x, zp}                      (DYN)
    const float *w_scale;    // s_w                            (DYN)
    const int32_t *wsum;     // sum_k wq[n,k]                  (DYN)
    void *out;
    int quant_type;
    // LLM.int8 mixed-precision decomposition (consumed only when flags != nullptr && flags[K] != 0)
    const int8_t *ca, *cb;
    const __half *a16;
    int a_pre_gelu;          // 1: a16 holds the values BEFORE the GELU whose output was quantized (the fp16 GELU output was
                             // never stored); the outlier path evaluates gelu_erf<__half> on the entries it needs
    int32_t *flags;          // [K + 2]: per-column flags, "any", completion counter
    int keep_flags;          // 1: leave the flags set (another GEMM consumes the same quantized rows next)
    int prefetch;            // > 0: the producer pulls the A row block it will load `prefetch` tiles later into L2
    int ldy;                 // row pitch of `out` in elements (>= N)
    // EPI_PLAIN: greedy choice folded into the epilogue.  keys[m] = max over columns of
    // (order-preserving bits of the ROUNDED output << 32) | (0xFFFFFFFF - n), columns with mask[n] != 0 count as -inf
    unsigned long long *argmax_keys;
    const uint8_t *mask;     // [>= tiles_n * tile columns] bytes, or nullptr
    int no_store;            // 1: the outputs themselves are not written (only the arg-max is wanted)
    // fused residual: out = clamp(round(y) + residual) -- HF's `hidden_states = residual + hidden_states` followed by
    // the fp16 clamp of WhisperEncoderLayer.forward (modeling_whisper.py:408-414), in the epilogue of fc2
    const void *residual;    // [M, N] of the output dtype, row pitch ldy, or nullptr
    float clamp_abs;         // > 0: clamp the sum to [-clamp_abs, clamp_abs]
};

// v = round_to_OutT(v) + residual (the addition HF performs on the rounded projection), optionally clamped.
template <typename OutT>
__device__ __forceinline__ void add_residual_chunk(float (&v)[32], const OutT *res_row, int nb, int N, bool row_ok,
                                                   float clamp_abs) {
    if (!row_ok) return;
    if (nb + 32 <= N) {
        uint4 raw[32 * sizeof(OutT) / 16];
#pragma unroll
        for (int j = 0; j < (int)(32 * sizeof(OutT) / 16); ++j) raw[j] = __ldg(reinterpret_cast<const uint4 *>(res_row + nb) + j);
        const OutT *r = reinterpret_cast<const OutT *>(raw);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __fadd_rn(to_f32(from_f32<OutT>(v[j])), to_f32(r[j]));
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (nb + j < N) v[j] = __fadd_rn(to_f32(from_f32<OutT>(v[j])), to_f32(res_row[nb + j]));
    }
    if (clamp_abs > 0.0f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fminf(fmaxf(v[j], -clamp_abs), clamp_abs);
    }
}

// Running arg-max of one output row over 32 adjacent columns starting at Y column `n_abs` (a multiple of 32).
// torch.argmax rules: first index among equal maxima (columns are visited in ascending order, strict >), NaN is the
// maximum.  The value compared is the output as it would be stored (rounded to OutT).
template <typename OutT>
__device__ __forceinline__ void argmax_chunk(const float (&v)[32], const uint8_t *mask, int n_abs, int N,
                                             uint32_t &best_ord, uint32_t &best_idx) {
    uint32_t mw[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
    if (mask != nullptr) {
        const uint4 m0 = __ldg(reinterpret_cast<const uint4 *>(mask + n_abs));
        const uint4 m1 = __ldg(reinterpret_cast<const uint4 *>(mask + n_abs) + 1);
        mw[0] = m0.x; mw[1] = m0.y; mw[2] = m0.z; mw[3] = m0.w;
        mw[4] = m1.x; mw[5] = m1.y; mw[6] = m1.z; mw[7] = m1.w;
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int n = n_abs + j;
        if (n >= N) break;
        float f = to_f32(from_f32<OutT>(v[j]));
        if ((mw[j >> 2] >> (8 * (j & 3))) & 0xffu) f = -INFINITY;
        const uint32_t ord = (f != f) ? 0xFFFFFFFFu : float_to_ordered(f);
        if (ord > best_ord) {
            best_ord = ord;
            best_idx = (uint32_t)n;
        }
    }
}

template <int BMODE> constexpr bool is_nibble() { return BMODE == B_4BIT || BMODE == B_U4; }
template <int BMODE> constexpr int dq_warps() { return BMODE == B_DIRECT ? 0 : (is_nibble<BMODE>() ? 8 : 4); }
template <int BN, int BMODE, int LEAN = 0> constexpr int num_threads() { return 32 * (2 + epi_warps<BN, BMODE, LEAN>() + dq_warps<BMODE>()); }

// WS > 0: weight-stationary schedule (B_DIRECT only).  The CTA keeps its W tile -- all num_kb <= WS k-blocks of BN
// rows -- resident in shared memory and walks M tiles of ONE n column block, so the ring streams A alone: per
// 256 x 128 tile with K = 512 the SM pulls 128 KB through L2 instead of 192 KB.
//
// COLS = 1: the tile is 128 rows x 2*BN columns and ONE tcgen05.mma per k-step covers it (M = 128, N = 256) instead
// of two M = 128, N = 128 instructions on two row halves.  An N = 128 instruction reads 4 KB of A and 4 KB of W from
// shared memory for 64 cycles of tensor work -- 128 B/clk, all the SM's shared-memory bandwidth, and the pipe ran
// ~45 % active; at N = 256 it is 12 KB per 128 cycles.  The two accumulator "halves" are then column halves.
//
// LEAN = 1: the tile is 128 rows (ONE half) x BN columns with 4 epilogue warps -- decode-shaped calls with at most 128
// rows.  Half the A bytes per stage, 192 threads and ~128 KB of shared memory instead of 320 / 226 KB, so that the CTA
// fits on an SM next to a resident attention CTA of the other half-batch's stream (fastgen two-stream decode).
//
// LEAN = 2 / 3: the same tile for calls with at most 32 / 64 rows (decode row groups).  The MMA still spans 128 TMEM
// lanes, but TMA fetches only a 32 / 64-row box of A per stage and the ring's A slots are packed at that pitch: the
// instruction reads the following slots as rows 32.. / 64.. -- arbitrary bytes, accumulated into TMEM lanes whose rows
// do not exist and are never stored.  A stage shrinks from 24 KB to 12 / 16 KB, so the same shared memory holds a ring
// twice as deep: these launches are bound by the round trips of a shallow ring over a deep K (fc2), not by bytes.
template <int LEAN> constexpr int lean_a_rows() { return LEAN == 2 ? 32 : (LEAN == 3 ? 64 : BMH); }

//
// PAIR = 1 (with COLS): a cluster of two CTAs computes 256 rows x 256 columns with one cta_group::2 instruction per
// k-step.  Per CTA and k-block the ring takes 16 KB of A (its 128 rows) + 16 KB of W (its 128 of the 256 W rows) for 512
// tensor clocks -- against 16 + 32 KB for the single-CTA 128 x 256 tile -- and the tensor core reads 8 KB instead of
// 12 KB of shared memory per instruction.  Weight-stationary pairs keep 64 KB of W per CTA and stream A alone
// (16 KB per stage: a ring of 7 stages = 3 500 tensor clocks of work in flight, above the L2 round trip that starved
// the 2-3-stage single-CTA rings).
template <int BN, int STAGES, int BMODE, int OUT_BUFS, int WS = 0, int COLS = 0, int LEAN = 0, int PAIR = 0>
struct SmemLayout {
    static_assert(PAIR == 0 || COLS == 1, "CTA pairs: column-split tiles (128 rows x 256 columns per CTA)");
    static_assert(LEAN == 0 || (WS == 0 && COLS == 0), "lean tiles: plain round-robin schedule");
    static_assert(WS == 0 || BMODE == B_DIRECT, "weight-stationary tiles take W straight from TMA");
    static_assert(COLS == 0 || ((BMODE == B_DIRECT || PAIR) && BN == 128),
                  "column-split tiles: W straight from TMA, or a CTA pair whose CTAs each expand BN = 128 rows of W");
    static constexpr int BMT = (COLS || LEAN) ? BMH : BM;      // tile rows
    static constexpr int BNT = COLS ? 2 * BN : BN;             // tile columns
    static constexpr int A_BYTES = (LEAN ? lean_a_rows<LEAN>() : BMT) * ROW_BYTES;
    static constexpr int B_BYTES = (PAIR ? BNT / 2 : BNT) * ROW_BYTES;   // pair: this CTA's half of the W rows
    static constexpr int TILE_M = PAIR ? 2 * BMT : BMT;                   // rows of one scheduled tile (pair: both CTAs)
    static_assert(LEAN < 2 || (BMODE == B_DIRECT && (STAGES - 1) * A_BYTES + BMH * ROW_BYTES <= STAGES * (A_BYTES + B_BYTES)),
                  "packed A slots: the last slot's 128-row read must stay inside the ring");
    static constexpr int B_SLOTS = WS > 0 ? WS : STAGES;       // W buffers: one per ring stage, or the resident k-blocks
    static constexpr int P_ROW = is_byte<BMODE>() ? 64 : (is_nibble<BMODE>() ? 32 : 0);
    static constexpr int P_BYTES = BN * P_ROW;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * A_BYTES;
    static constexpr int OFF_P = OFF_B + B_SLOTS * B_BYTES;
    static constexpr int EW = epi_warps<BN, BMODE, LEAN>();
    static constexpr int OFF_OUT = OFF_P + STAGES * P_BYTES;   // EW x OUT_BUFS boxes, 1024-byte aligned
    static constexpr int CONST_BUFS = WS > 0 ? 1 : 2;          // a weight-stationary CTA never changes columns
    static constexpr int OFF_CONST = OFF_OUT + EW * OUT_BUFS * BOX_BYTES;  // float [CONST_BUFS][3][BNT] per-tile constants
    static constexpr int OFF_LUT = OFF_CONST + CONST_BUFS * 3 * BNT * 4;          // float lut[16]
    static constexpr int OFF_BAR = OFF_LUT + 64;               // uint64 barriers
    static constexpr int NUM_BARS = 4 * STAGES + 2 * ACC_STAGES + 1;   // full | empty | bready | pfull (pairs: packed W landed)
    static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
    static constexpr int TOTAL = OFF_TMEM + 16 + 1024;         // + slack for manual 1024-B alignment
    static constexpr int TX_BYTES = A_BYTES + (WS > 0 ? 0 : (BMODE == B_DIRECT ? B_BYTES : P_BYTES));
    static_assert(TOTAL <= 232448, "shared memory budget exceeded");
};

// i-th tile of this CTA -> (m tile, n tile); false past the CTA's last tile.  Default: tiles round-robin over the
// grid, n fastest (the CTAs sharing an A row-block run together).  Weight-stationary: CTA c owns column block
// c % tiles_n and every (gridDim / tiles_n)-th row block; the tiles_n CTAs of one row block still run together.
// PAIR: the scheduling unit is the CTA pair (consecutive blockIdx.x), both CTAs walk the same tiles.
template <int WS, int PAIR = 0>
__device__ __forceinline__ bool tile_at(const GemmArgs &args, int i, int &mt, int &nt) {
    const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    if constexpr (WS > 0) {
        nt = unit % args.tiles_n;
        mt = unit / args.tiles_n + i * (units / args.tiles_n);
        return mt < args.tiles_m;
    } else {
        const int tile = unit + i * units;
        nt = tile % args.tiles_n;
        mt = tile / args.tiles_n;
        return tile < args.tiles_m * args.tiles_n;
    }
}

template <typename OutT> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// Epilogue arithmetic on a pair of adjacent columns (operation order fixed: the CPU oracle restates
// it step by step; the packed f32x2 instructions round exactly like their scalar forms).
template <int EPI>
__device__ __forceinline__ void epi_pair(uint32_t r0, uint32_t r1, float cs0, float cs1, float b0, float b1, int aux0,
                                         int aux1, float rs, float dyn_s, int dyn_zp, float &v0, float &v1) {
    if constexpr (EPI == EPI_LLMINT8) {
        v0 = (float)(int)r0;
        v1 = (float)(int)r1;
        mul2(v0, v1, rs, rs);
        mul2(v0, v1, cs0, cs1);
        fma2(v0, v1, 6.200012e-05f, 6.200012e-05f, b0, b1);
    } else if constexpr (EPI == EPI_DYN) {
        // scalar mul + add: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2, which
        // would round differently from the reference's two operations
        v0 = __fadd_rn(__fmul_rn((float)((int)r0 - dyn_zp * aux0), dyn_s), b0);
        v1 = __fadd_rn(__fmul_rn((float)((int)r1 - dyn_zp * aux1), dyn_s), b1);
    } else if constexpr (EPI == EPI_W8A16) {
        v0 = __fadd_rn(__fmul_rn(__uint_as_float(r0), cs0), b0);
        v1 = __fadd_rn(__fmul_rn(__uint_as_float(r1), cs1), b1);
    } else if constexpr (EPI == EPI_W8A8) {     // quanto qbytes_int_mm: int32 * (s_in * s_w[n]) in fp32, then + bias
        v0 = __fadd_rn(__fmul_rn((float)(int)r0, cs0), b0);
        v1 = __fadd_rn(__fmul_rn((float)(int)r1, cs1), b1);
    } else {
        v0 = __fadd_rn(__uint_as_float(r0), b0);
        v1 = __fadd_rn(__uint_as_float(r1), b1);
    }
}

// 32 adjacent columns starting at tile column `col0`; the per-column constants of the tile were
// staged in shared memory (sc: [3][BN] = col scale | bias | wsum) by the epilogue warps.
template <int BN, int EPI>
__device__ __forceinline__ void epi_chunk(const uint32_t (&r)[32], float (&v)[32], const float *sc, int col0, float rs,
                                          float dyn_s, int dyn_zp) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float4 cs = *reinterpret_cast<const float4 *>(sc + col0 + 4 * g);
        const float4 b = *reinterpret_cast<const float4 *>(sc + BN + col0 + 4 * g);
        int4 aux = make_int4(0, 0, 0, 0);
        if constexpr (EPI == EPI_DYN) aux = *reinterpret_cast<const int4 *>(sc + 2 * BN + col0 + 4 * g);
        epi_pair<EPI>(r[4 * g], r[4 * g + 1], cs.x, cs.y, b.x, b.y, aux.x, aux.y, rs, dyn_s, dyn_zp, v[4 * g],
                      v[4 * g + 1]);
        epi_pair<EPI>(r[4 * g + 2], r[4 * g + 3], cs.z, cs.w, b.z, b.w, aux.z, aux.w, rs, dyn_s, dyn_zp, v[4 * g + 2],
                      v[4 * g + 3]);
    }
}

// Write 32 finished columns of one row: into the swizzled staging box (TMA store path) or
// straight to global memory (row pitches that are not 16-byte multiples).
template <typename OutT, bool TMA_STORE>
__device__ __forceinline__ void emit_chunk(const float (&v)[32], uint8_t *box, int cc, int lane, OutT *row_ptr,
                                           int nb, int N, bool row_ok, bool vec_ok) {
    if constexpr (TMA_STORE) {
        constexpr int CH = 32 * (int)sizeof(OutT) / 16;   // 16-byte chunks per 32 columns
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int c16 = cc * CH + j;                   // chunk index inside the 128-byte row
            uint8_t *dst = box + lane * 128 + ((c16 ^ (lane & 7)) << 4);
            if constexpr (sizeof(OutT) == 4) {
                *reinterpret_cast<float4 *>(dst) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
                uint4 u;
                u.x = pack2<OutT>(v[8 * j + 0], v[8 * j + 1]);
                u.y = pack2<OutT>(v[8 * j + 2], v[8 * j + 3]);
                u.z = pack2<OutT>(v[8 * j + 4], v[8 * j + 5]);
                u.w = pack2<OutT>(v[8 * j + 6], v[8 * j + 7]);
                *reinterpret_cast<uint4 *>(dst) = u;
            }
        }
    } else if (row_ok) {
        if (vec_ok && nb + 32 <= N) {
            if constexpr (sizeof(OutT) == 4) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    reinterpret_cast<float4 *>(row_ptr + nb)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 u;
                    u.x = pack2<OutT>(v[8 * j + 0], v[8 * j + 1]);
                    u.y = pack2<OutT>(v[8 * j + 2], v[8 * j + 3]);
                    u.z = pack2<OutT>(v[8 * j + 4], v[8 * j + 5]);
                    u.w = pack2<OutT>(v[8 * j + 6], v[8 * j + 7]);
                    reinterpret_cast<uint4 *>(row_ptr + nb)[j] = u;
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (nb + j < N) row_ptr[nb + j] = from_f32<OutT>(v[j]);
        }
    }
}

// LLM.int8 mixed-precision decomposition for one 32-column chunk -- the RARE path (an activation
// entry with |a| >= threshold exists).  Kept out of line so that the hot epilogue loop stays small
// enough for the instruction cache.  Re-reads the accumulator itself, removes the outlier columns'
// int8 products from the exact int32 sums (== bitsandbytes zeroing CA[:, cols]) and adds the fp16
// side product sum_c A[m,c] * fp16(CB[n,c] * SCB[n] / 127) in ascending column order.
template <int BN, typename OutT, bool TMA_STORE>
__device__ __noinline__ void llmint8_outlier_chunk(const GemmArgs &args, const float *sc, uint32_t taddr, int n_abs,
                                                   int col0, int m, bool row_ok, float rs, uint8_t *box, int cc,
                                                   int lane, OutT *row_ptr, bool vec_ok) {
    uint32_t r[32];
    tmem_ld_32x32(taddr, r);
    tmem_ld_wait();
    float o[32], v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = 0.0f;
#pragma unroll 1
    for (int c = 0; c < args.K; ++c) {
        if (args.flags[c] == 0) continue;
        const int a8 = row_ok ? (int)args.ca[(size_t)m * args.K + c] : 0;
        float af = 0.0f;
        if (row_ok) {
            const __half raw = args.a16[(size_t)m * args.K + c];
            af = __half2float(args.a_pre_gelu ? gelu_erf<__half>(__half2float(raw)) : raw);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int n = n_abs + j;
            if (n < args.N) {
                const int b8 = (int)args.cb[(size_t)n * args.K + c];
                r[j] = (uint32_t)((int)r[j] - a8 * b8);
                const float d = __fmul_rn(__fmul_rn((float)b8, __ldg(args.col_scale + n)), 7.874015718698502e-3f);
                o[j] = fmaf(af, __half2float(__float2half_rn(d)), o[j]);
            }
        }
    }
    epi_chunk<BN, EPI_LLMINT8>(r, v, sc, col0, rs, 0.0f, 0);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __half2float(__float2half_rn(v[j])) + o[j];   // fp16 addmm
    if (args.residual != nullptr)
        add_residual_chunk<OutT>(v, reinterpret_cast<const OutT *>(args.residual) + (size_t)(row_ok ? m : 0) * args.ldy, n_abs,
                                 args.N, row_ok, args.clamp_abs);
    emit_chunk<OutT, TMA_STORE>(v, box, cc, lane, row_ptr, n_abs, args.N, row_ok, vec_ok);
}

// Body of one epilogue warp: drains its 32-row slab of every tile this CTA owns.
template <int BN, int EW, int EPI, typename OutT, int OUT_BUFS, bool TMA_STORE, int WS, int COLS, int LEAN, int PAIR>
__device__ __forceinline__ void epilogue_warp(const GemmArgs &args, const CUtensorMap *map_y, uint8_t *boxes,
                                              float *s_const, uint64_t *bar_tmem_full, uint64_t *bar_tmem_empty,
                                              uint32_t tmem_base, int warp, int lane, int rank) {
    constexpr int ACC_COLS = 2 * BN;
    constexpr int BMT = (COLS || LEAN) ? BMH : BM, BNT = COLS ? 2 * BN : BN;   // tile rows / columns
    constexpr int BOX_COLS = 128 / (int)sizeof(OutT);   // columns per TMA-store box (64 or 32)
    constexpr int CSPLIT = EW >= 8 ? EW / 8 : 1;        // warps sharing a 32-row slab split its columns
    constexpr int N_BOX = BN / BOX_COLS / CSPLIT;       // boxes per warp per tile
    constexpr int NCH = BOX_COLS / 32;                  // tcgen05.ld chunks per box (2 or 1)
    const int grp = (warp - 2) >> 2;
    const int h = grp & 1;                              // 128-row half of the tile (COLS: 128-column half)
    const int bx0 = (grp >> 1) * N_BOX;                 // first box (column range) of this warp
    const int q = warp & 3;                             // TMEM lane quarter this warp may access
    const bool vec_ok = ((size_t)args.ldy * sizeof(OutT)) % 16 == 0;
    const bool any_outlier = (EPI == EPI_LLMINT8) && args.flags != nullptr && args.flags[args.K] != 0;
    float dyn_s = 0.0f;
    int dyn_zp = 0;
    if constexpr (EPI == EPI_DYN) {
        dyn_s = __fmul_rn(args.qparams[0], args.w_scale[0]);
        dyn_zp = (int)args.qparams[1];
    }
    uint32_t t = 0, nstore = 0;
    int mt, nt;
    for (; tile_at<WS, PAIR>(args, (int)t, mt, nt); ++t) {
        const int n0 = nt * BNT, m0 = mt * (PAIR ? 2 * BMT : BMT) + rank * BMT;   // pair: this CTA's 128 of the 256 rows
        const uint32_t as = t % ACC_STAGES, aph = (t / ACC_STAGES) & 1;
        const int mrow0 = m0 + (COLS ? 0 : h * BMH) + q * 32;  // first row of this warp's 32-row slab
        const int ch0 = COLS ? h * BN : 0;        // first tile column of this warp's half
        const int nh0 = n0 + ch0;                 // ... as a column of Y
        const int m = mrow0 + lane;
        const bool row_ok = m < args.M;
        const bool st_ok = row_ok && !args.no_store;
        const bool slab_ok = mrow0 < args.M && nh0 < args.N;      // warp-uniform
        uint32_t best_ord = 0u, best_idx = 0xFFFFFFFFu;           // EPI_PLAIN arg-max of this row over the tile
        float rs = 1.0f;
        if constexpr (EPI == EPI_LLMINT8) rs = row_ok ? __ldg(args.row_scale + m) : 0.0f;
        // stage this tile's per-column constants once (the 8 epilogue warps share them); two
        // buffers, so a warp that runs ahead never overwrites constants still in use (weight-stationary
        // CTAs keep their column block: staged once)
        float *sc = s_const + (WS > 0 ? 0 : (t & 1) * 3 * BNT);
        if (WS == 0 || t == 0) {
            const float *bias = reinterpret_cast<const float *>(args.bias);
            for (int i = (warp - 2) * 32 + lane; i < 3 * BNT; i += EW * 32) {
                const int which = i / BNT;
                const int n = min(n0 + (i - which * BNT), args.N - 1);
                float val = 0.0f;
                if (which == 0) {
                    if constexpr (EPI == EPI_LLMINT8 || EPI == EPI_W8A16 || EPI == EPI_W8A8) val = __ldg(args.col_scale + n);
                } else if (which == 1) {
                    if (bias != nullptr) val = __ldg(bias + n);
                } else {
                    if constexpr (EPI == EPI_DYN) val = __int_as_float(__ldg(args.wsum + n));
                }
                sc[i] = val;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");
        }

        mbar_wait(&bar_tmem_full[as], aph);
        tc_fence_after();
        if (slab_ok) {
            OutT *row_ptr = reinterpret_cast<OutT *>(args.out) + (size_t)(row_ok ? m : 0) * args.ldy;
            const uint32_t tmem_row = tmem_base + ((uint32_t)(q * 32) << 16) + as * ACC_COLS + h * BN;
#pragma unroll 1
            for (int bx = bx0; bx < bx0 + N_BOX; ++bx) {
                uint8_t *box = boxes + (nstore % OUT_BUFS) * BOX_BYTES;
                if constexpr (TMA_STORE) {
                    // the store that last used this box must have finished reading it
                    if (lane == 0) tma_store_wait_read<OUT_BUFS - 1>();
                    __syncwarp();
                }
                bool fast = true;
                if constexpr (EPI == EPI_LLMINT8) {
                    if (any_outlier) {
                        fast = false;
#pragma unroll 1
                        for (int cc = 0; cc < NCH; ++cc) {
                            const int col0 = bx * BOX_COLS + cc * 32;
                            llmint8_outlier_chunk<BNT, OutT, TMA_STORE>(args, sc, tmem_row + col0, nh0 + col0, ch0 + col0, m,
                                                                        row_ok, rs, box, cc, lane, row_ptr, vec_ok);
                        }
                    }
                }
                if (fast) {
                    if constexpr (EW == 16) {   // 576 threads: 113 registers each -> one chunk in flight
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) {
                            const int col0 = bx * BOX_COLS + cc * 32;   // column inside the tile
                            uint32_t r[32];
                            tmem_ld_32x32(tmem_row + col0, r);
                            tmem_ld_wait();
                            float v[32];
                            epi_chunk<BNT, EPI>(r, v, sc, ch0 + col0, rs, dyn_s, dyn_zp);
                            emit_chunk<OutT, TMA_STORE>(v, box, cc, lane, row_ptr, nh0 + col0, args.N, st_ok, vec_ok);
                        }
                    } else {
                        uint32_t r[NCH][32];
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) tmem_ld_32x32(tmem_row + bx * BOX_COLS + cc * 32, r[cc]);
                        tmem_ld_wait();
#pragma unroll
                        for (int cc = 0; cc < NCH; ++cc) {
                            const int col0 = bx * BOX_COLS + cc * 32;   // column inside the tile
                            float v[32];
                            epi_chunk<BNT, EPI>(r[cc], v, sc, ch0 + col0, rs, dyn_s, dyn_zp);
                            if constexpr (EPI == EPI_PLAIN) {
                                if (args.argmax_keys != nullptr)
                                    argmax_chunk<OutT>(v, args.mask, nh0 + col0, args.N, best_ord, best_idx);
                            }
                            if (args.residual != nullptr)
                                add_residual_chunk<OutT>(v, reinterpret_cast<const OutT *>(args.residual) + (size_t)(row_ok ? m : 0) * args.ldy,
                                                         nh0 + col0, args.N, row_ok, args.clamp_abs);
                            emit_chunk<OutT, TMA_STORE>(v, box, cc, lane, row_ptr, nh0 + col0, args.N, st_ok, vec_ok);
                        }
                    }
                }
                if constexpr (TMA_STORE) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(map_y, box, nh0 + bx * BOX_COLS, mrow0);
                        tma_store_commit();
                    }
                    ++nstore;
                }
            }
        }
        if constexpr (EPI == EPI_PLAIN) {
            if (args.argmax_keys != nullptr && row_ok && best_idx != 0xFFFFFFFFu)
                atomicMax(args.argmax_keys + m, ((unsigned long long)best_ord << 32) | (0xFFFFFFFFu - best_idx));
        }
        // accumulator fully read: hand the TMEM stage back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_tmem_empty[as]), 0));   // the leader's MMA warp waits
            else mbar_arrive(&bar_tmem_empty[as]);
        }
    }
    if constexpr (TMA_STORE) {
        if (lane == 0) tma_store_wait<0>();
    }
    __syncwarp();
}

template <int BN, int STAGES, int AKIND, int BMODE, int EPI, typename OutT, int OUT_BUFS, int WS, int COLS, int LEAN, int PAIR>
__global__ void __launch_bounds__(num_threads<BN, BMODE, LEAN>(), (LEAN && BMODE == B_DIRECT) ? 2 : 1)   // lean int8: <= 168 registers
k_gemm_tc(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
          const __grid_constant__ CUtensorMap map_y, const GemmArgs args) {
    using L = SmemLayout<BN, STAGES, BMODE, OUT_BUFS, WS, COLS, LEAN, PAIR>;
    constexpr int BMT = L::BMT, BNT = L::BNT;
    static_assert(BN == 64 || BN == 128, "BN must be 64 or 128 (2 halves x 2 stages x BN <= 512 TMEM columns)");
    constexpr bool kIntKind = (AKIND == A_S8 || AKIND == A_U8);
    constexpr int A_ELEMS_PER_ROW = kIntKind ? 128 : 64;  // elements of K per 128-byte row
    constexpr int DQ_WARPS = dq_warps<BMODE>();
    constexpr int EW = epi_warps<BN, BMODE, LEAN>();
    constexpr int ACC_COLS = 2 * BN;                      // one accumulator stage: half 0 | half 1

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L::OFF_BAR);
    uint64_t *bar_full = bars;
    uint64_t *bar_empty = bars + STAGES;
    uint64_t *bar_bready = bars + 2 * STAGES;
    uint64_t *bar_pfull = bars + 3 * STAGES;                   // pairs with packed weights: this CTA's packed tile landed
    uint64_t *bar_tmem_full = bars + 4 * STAGES;
    uint64_t *bar_tmem_empty = bars + 4 * STAGES + ACC_STAGES;
    uint64_t *bar_w = bars + 4 * STAGES + 2 * ACC_STAGES;     // weight-stationary: the resident W tile has landed
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(smem + L::OFF_TMEM);
    float *s_lut = reinterpret_cast<float *>(smem + L::OFF_LUT);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = args.num_kb;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;      // 0: the pair's leader (issues the MMAs, owns the barriers the
                                                             // pair synchronises on: full[], tmem_empty[], bar_w)

    // ---------------- setup ----------------
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        if (args.tma_store) tma_prefetch_desc(&map_y);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], PAIR ? 2 : 1);      // pair: the leader's expect_tx arrive + the peer producer's arrive
            mbar_init(&bar_empty[s], 1);
            mbar_init(&bar_bready[s], DQ_WARPS > 0 ? (PAIR ? 2 * DQ_WARPS : DQ_WARPS) : 1);   // pair: both CTAs' expansion warps
            mbar_init(&bar_pfull[s], 1);
        }
        for (int a = 0; a < ACC_STAGES; ++a) {
            mbar_init(&bar_tmem_full[a], 1);
            mbar_init(&bar_tmem_empty[a], PAIR ? 2 * EW : EW);   // pair: both CTAs' epilogue warps
        }
        mbar_init(bar_w, PAIR ? 2 : 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) {
            tmem_alloc_pair(tmem_holder, ACC_STAGES * ACC_COLS);
            tmem_relinquish_pair();
        } else {
            tmem_alloc(tmem_holder, ACC_STAGES * ACC_COLS);
            tmem_relinquish();
        }
    }
    if constexpr (BMODE == B_4BIT) {
        if (threadIdx.x < 16) s_lut[threadIdx.x] = args.quant_type ? kFP4Code[threadIdx.x] : kNF4Code[threadIdx.x];
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();     // the peer's barriers are initialised before anything arrives on them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_prologue_done();

    if (warp == 0) {
        // ---------------- TMA producer ----------------
        // The whole warp walks the loop and waits; one elected lane issues.  (Under `if (lane == 0)` the compiler
        // cannot see that a single thread is active and wraps every UTMALDG / UTCIMMA / UTCBAR in an elect-broadcast-
        // branch loop: ~55 cycles per instruction, ~500 cycles per k-block of pure issue overhead in each of the two
        // warps -- more than the tensor time of an int8 k-block, and the whole cost of a decode-shaped call.)
        uint32_t s = 0, ph = 0;      // ring slot and its parity
        int mt, nt;
        // pair: every load lands in this CTA's shared memory and is counted on the LEADER's barrier, where the MMA
        // warp waits for both halves (the leader expects the bytes of both CTAs, the peer only arrives)
        const int b_row = PAIR ? rank * (BNT / 2) : 0;       // this CTA's rows inside the W tile
        if constexpr (WS > 0) {
            if (tile_at<WS, PAIR>(args, 0, mt, nt)) {      // the CTA's W tile, once: num_kb boxes of BN rows x 128 B
                if (elect_one()) {
                    if constexpr (PAIR) {
                        const uint32_t lw = mapa_u32(smem_u32(bar_w), 0);
                        if (rank == 0) mbar_arrive_expect_tx(bar_w, 2u * (uint32_t)num_kb * L::B_BYTES);
                        else mbar_arrive_cluster(lw);
                        for (int kb = 0; kb < num_kb; ++kb)
                            tma_load_2d_pair(smem + L::OFF_B + kb * L::B_BYTES, &map_b, lw, kb * 128, nt * BNT + b_row);
                    } else {
                        mbar_arrive_expect_tx(bar_w, (uint32_t)num_kb * L::B_BYTES);
                        for (int kb = 0; kb < num_kb; ++kb)
                            tma_load_2d(smem + L::OFF_B + kb * L::B_BYTES, &map_b, bar_w, kb * 128, nt * BNT);
                    }
                }
                __syncwarp();
            }
        }
        for (int i = 0; tile_at<WS, PAIR>(args, i, mt, nt); ++i) {
            const int n0 = nt * BNT, m0 = mt * L::TILE_M + rank * BMT;
            // A streams from HBM with ~1.5 us of loaded latency and the ring holds < 1 tile: ask L2 for the row
            // block this CTA reaches `prefetch` tiles from now (one of the tiles_n CTAs sharing it does)
            int pf_m0 = -1;
            if (args.prefetch > 0) {
                int pmt, pnt;
                if (tile_at<WS, PAIR>(args, i + args.prefetch, pmt, pnt) && pmt % args.tiles_n == pnt)
                    pf_m0 = pmt * L::TILE_M + rank * BMT;
            }
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&bar_empty[s], ph ^ 1);
                if constexpr (PAIR) {
                    if (elect_one()) {
                        const uint32_t lf = mapa_u32(smem_u32(&bar_full[s]), 0);
                        // bytes the leader's MMA warp waits for: A of both CTAs, and W when it comes straight from TMA
                        // (packed weights of the expanding schemes are counted on pfull[s] instead)
                        constexpr uint32_t kPairTx = L::TX_BYTES - (BMODE != B_DIRECT ? L::P_BYTES : 0);
                        if (rank == 0) mbar_arrive_expect_tx(&bar_full[s], 2u * kPairTx);
                        else mbar_arrive_cluster(lf);
                        tma_load_2d_pair(smem + L::OFF_A + s * L::A_BYTES, &map_a, lf, kb * A_ELEMS_PER_ROW, m0);
                        if (pf_m0 >= 0) tma_prefetch_2d(&map_a, kb * A_ELEMS_PER_ROW, pf_m0);
                        if constexpr (BMODE != B_DIRECT) {
                            // packed weights: this CTA's expansion warps wait for them on a barrier of their own
                            mbar_arrive_expect_tx(&bar_pfull[s], L::P_BYTES);
                            tma_load_2d(smem + L::OFF_P + s * L::P_BYTES, &map_b, &bar_pfull[s], kb * L::P_ROW, n0 + b_row);
                        } else if constexpr (WS == 0) {
                            tma_load_2d_pair(smem + L::OFF_B + s * L::B_BYTES, &map_b, lf, kb * A_ELEMS_PER_ROW, n0 + b_row);
                        }
                    }
                } else if (elect_one()) {
                    mbar_arrive_expect_tx(&bar_full[s], L::TX_BYTES);
                    tma_load_2d(smem + L::OFF_A + s * L::A_BYTES, &map_a, &bar_full[s], kb * A_ELEMS_PER_ROW, m0);
                    if (pf_m0 >= 0) tma_prefetch_2d(&map_a, kb * A_ELEMS_PER_ROW, pf_m0);
                    if constexpr (WS > 0)
                        (void)n0;
                    else if constexpr (BMODE == B_DIRECT)
                        tma_load_2d(smem + L::OFF_B + s * L::B_BYTES, &map_b, &bar_full[s], kb * A_ELEMS_PER_ROW, n0);
                    else
                        tma_load_2d(smem + L::OFF_P + s * L::P_BYTES, &map_b, &bar_full[s], kb * L::P_ROW, n0);
                }
                __syncwarp();
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ---------------- MMA issuer (whole warp in the loop, one elected lane issues; pair: the leader CTA's) ----------------
        constexpr uint32_t idesc =
            kIntKind ? make_idesc(kAccS32, AKIND == A_S8 ? kFmtS8 : kFmtU8, kFmtS8, PAIR ? 2 * BMH : BMH, BNT)
                     : make_idesc(kAccF32, AKIND == A_F16 ? kFmtF16 : kFmtBF16,
                                  AKIND == A_F16 ? kFmtF16 : kFmtBF16, PAIR ? 2 * BMH : BMH, PAIR ? BNT : BN);
        uint32_t s = 0, ph = 0, t = 0;
        int mt, nt;
        for (; tile_at<WS, PAIR>(args, (int)t, mt, nt); ++t) {
            const int m0 = mt * BMT;
            const bool two_halves = COLS == 0 && LEAN == 0 && m0 + BMH < args.M;   // second 128 rows hold real data
            if constexpr (WS > 0) {
                if (t == 0) mbar_wait(bar_w, 0);
            }
            const uint32_t as = t % ACC_STAGES, aph = (t / ACC_STAGES) & 1;
            mbar_wait(&bar_tmem_empty[as], aph ^ 1);      // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + as * ACC_COLS;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&bar_full[s], ph);
                if constexpr (BMODE != B_DIRECT) mbar_wait(&bar_bready[s], ph);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc0 = make_smem_desc_sw128(smem + L::OFF_A + s * L::A_BYTES);
                    const uint64_t adesc1 = make_smem_desc_sw128(smem + L::OFF_A + s * L::A_BYTES + BMH * ROW_BYTES);
                    const uint64_t bdesc = make_smem_desc_sw128(smem + L::OFF_B + (WS > 0 ? kb : s) * L::B_BYTES);
#pragma unroll
                    for (int k = 0; k < ROW_BYTES / UMMA_K_BYTES; ++k) {
                        const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                        if constexpr (PAIR) {
                            if constexpr (kIntKind) umma_i8_pair(tmem_acc, adesc0 + 2 * k, bdesc + 2 * k, idesc, acc);
                            else umma_f16_pair(tmem_acc, adesc0 + 2 * k, bdesc + 2 * k, idesc, acc);
                        } else if constexpr (kIntKind) {
                            umma_i8(tmem_acc, adesc0 + 2 * k, bdesc + 2 * k, idesc, acc);
                            if (two_halves) umma_i8(tmem_acc + BN, adesc1 + 2 * k, bdesc + 2 * k, idesc, acc);
                        } else {
                            umma_f16(tmem_acc, adesc0 + 2 * k, bdesc + 2 * k, idesc, acc);
                            if (two_halves) umma_f16(tmem_acc + BN, adesc1 + 2 * k, bdesc + 2 * k, idesc, acc);
                        }
                    }
                    if constexpr (PAIR) umma_commit_pair(&bar_empty[s]);   // frees the slot in both CTAs
                    else umma_commit(&bar_empty[s]);  // implicit tcgen05.fence::before_thread_sync
                }
                __syncwarp();
                if (++s == STAGES) {
                    s = 0;
                    ph ^= 1;
                }
            }
            if (elect_one()) {
                if constexpr (PAIR) umma_commit_pair(&bar_tmem_full[as]);
                else umma_commit(&bar_tmem_full[as]);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // the peer CTA's warp 1 only owns its half of the pair's TMEM allocation
    } else if (warp < 2 + EW) {
        // ---------------- epilogue: warps 2..9 ----------------
        uint8_t *boxes = smem + L::OFF_OUT + (warp - 2) * OUT_BUFS * BOX_BYTES;   // private staging of this warp
        float *s_const = reinterpret_cast<float *>(smem + L::OFF_CONST);
        if (args.tma_store)
            epilogue_warp<BN, EW, EPI, OutT, OUT_BUFS, true, WS, COLS, LEAN, PAIR>(args, &map_y, boxes, s_const, bar_tmem_full, bar_tmem_empty, tmem_base, warp, lane, rank);
        else
            epilogue_warp<BN, EW, EPI, OutT, OUT_BUFS, false, WS, COLS, LEAN, PAIR>(args, &map_y, boxes, s_const, bar_tmem_full, bar_tmem_empty, tmem_base, warp, lane, rank);
    } else {
        // ---------------- weight expansion (W8A16 / W4A16): warps 10.. ----------------
        const int t = threadIdx.x - 32 * (2 + EW);
        if constexpr (is_byte<BMODE>()) {
            const int qd = t & 3, row0 = t >> 2;      // 128 threads: 4 per row, 32 rows per pass
            uint32_t it = 0;
            int mt, nt;
            for (int ti = 0; tile_at<WS, PAIR>(args, ti, mt, nt); ++ti) {
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(PAIR ? &bar_pfull[s] : &bar_full[s], ph);
                    const uint8_t *P = smem + L::OFF_P + s * L::P_BYTES;
                    uint8_t *B = smem + L::OFF_B + s * L::B_BYTES;
#pragma unroll
                    for (int p = 0; p < BN / 32; ++p) {
                        const int r = row0 + 32 * p;
                        const uint4 v = *reinterpret_cast<const uint4 *>(P + r * 64 + qd * 16);
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                        uint32_t o[8];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if constexpr (BMODE == B_F8) {
                                // e4m3 -> fp16 is exact (3 mantissa bits, exponents inside fp16's range); bf16 likewise
                                uint32_t lo, hi;
                                asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(lo) : "h"((unsigned short)(w[j] & 0xffffu)));
                                asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(hi) : "h"((unsigned short)(w[j] >> 16)));
                                if constexpr (AKIND == A_F16) {
                                    o[2 * j] = lo;
                                    o[2 * j + 1] = hi;
                                } else {
                                    const float2 fl = __half22float2(*reinterpret_cast<__half2 *>(&lo));
                                    const float2 fh = __half22float2(*reinterpret_cast<__half2 *>(&hi));
                                    o[2 * j] = pack2<__nv_bfloat16>(fl.x, fl.y);
                                    o[2 * j + 1] = pack2<__nv_bfloat16>(fh.x, fh.y);
                                }
                            } else if constexpr (AKIND == A_F16) {
                                // s8 -> fp16 exactly: (1024 + (s ^ 0x80)) - 1152
                                const uint32_t x = w[j] ^ 0x80808080u;
                                uint32_t lo = __byte_perm(x, 0x64646464u, 0x4140);
                                uint32_t hi = __byte_perm(x, 0x64646464u, 0x4342);
                                const __half2 bias2 = __half2half2(__ushort_as_half((unsigned short)0x6480));
                                __half2 l2 = __hsub2(*reinterpret_cast<__half2 *>(&lo), bias2);
                                __half2 h2 = __hsub2(*reinterpret_cast<__half2 *>(&hi), bias2);
                                o[2 * j] = *reinterpret_cast<uint32_t *>(&l2);
                                o[2 * j + 1] = *reinterpret_cast<uint32_t *>(&h2);
                            } else {
                                const float f0 = (float)(int8_t)(w[j] & 0xff), f1 = (float)(int8_t)((w[j] >> 8) & 0xff);
                                const float f2 = (float)(int8_t)((w[j] >> 16) & 0xff), f3 = (float)(int8_t)(w[j] >> 24);
                                o[2 * j] = pack2<__nv_bfloat16>(f0, f1);
                                o[2 * j + 1] = pack2<__nv_bfloat16>(f2, f3);
                            }
                        }
                        *reinterpret_cast<uint4 *>(B + sw128_offset(r, 2 * qd)) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4 *>(B + sw128_offset(r, 2 * qd + 1)) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_bready[s]), 0));   // the leader's MMA warp waits
                        else mbar_arrive(&bar_bready[s]);
                    }
                }
            }
        } else if constexpr (is_nibble<BMODE>()) {
            const int hf = t & 1, r = t >> 1;         // 256 threads: 2 per row; rows >= BN idle
            const bool active = r < BN;
            uint32_t it = 0;
            int mt, nt;
            for (int ti = 0; tile_at<WS, PAIR>(args, ti, mt, nt); ++ti) {
                const int n = nt * BNT + (PAIR ? rank * BN : 0) + r;      // pair: this CTA expands its 128 of the 256 W rows
                const float *am_row = args.absmax + (size_t)(n < args.N ? n : 0) * args.absmax_ld;
                const float *sh_row = (BMODE == B_U4) ? args.shift + (size_t)(n < args.N ? n : 0) * args.absmax_ld : nullptr;
                const bool n_ok = active && n < args.N;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    // NF4/FP4: one absmax per 64-element block; quanto qint4: scale / shift of the
                    // group holding this thread's 32 weights
                    const int gi = (BMODE == B_U4) ? (kb * 64 + hf * 32) / args.group : kb;
                    const float am = n_ok ? __ldg(am_row + gi) : 0.0f;
                    float sh = 0.0f;
                    if constexpr (BMODE == B_U4) sh = n_ok ? __ldg(sh_row + gi) : 0.0f;
                    mbar_wait(PAIR ? &bar_pfull[s] : &bar_full[s], ph);
                    if (active) {
                        const uint8_t *P = smem + L::OFF_P + s * L::P_BYTES;
                        uint8_t *B = smem + L::OFF_B + s * L::B_BYTES;
                        const uint4 v = *reinterpret_cast<const uint4 *>(P + r * 32 + hf * 16);
                        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t o[4];
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const uint32_t byte = (w[j] >> (8 * b)) & 0xffu;
                                float f0, f1;
                                if constexpr (BMODE == B_U4) {   // quanto AffineQuantizer: scale * q - shift
                                    f0 = __fsub_rn(__fmul_rn(am, (float)(byte >> 4)), sh);
                                    f1 = __fsub_rn(__fmul_rn(am, (float)(byte & 15u)), sh);
                                } else {
                                    f0 = __fmul_rn(s_lut[byte >> 4], am);
                                    f1 = __fmul_rn(s_lut[byte & 15u], am);
                                }
                                if constexpr (AKIND == A_F16) o[b] = pack2<__half>(f0, f1);
                                else o[b] = pack2<__nv_bfloat16>(f0, f1);
                            }
                            *reinterpret_cast<uint4 *>(B + sw128_offset(r, 4 * hf + j)) = make_uint4(o[0], o[1], o[2], o[3]);
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (PAIR) mbar_arrive_cluster(mapa_u32(smem_u32(&bar_bready[s]), 0));
                        else mbar_arrive(&bar_bready[s]);
                    }
                }
            }
        }
    }

    // ---------------- teardown ----------------
    tc_fence_before();
    if constexpr (PAIR) {
        // neither CTA may leave while the other still reads its shared memory through the tensor core or arrives on
        // its barriers
        cluster_sync_all();
        if (warp == 1) tmem_dealloc_pair(tmem_base, ACC_STAGES * ACC_COLS);
    } else {
        __syncthreads();
        if (warp == 1) tmem_dealloc(tmem_base, ACC_STAGES * ACC_COLS);
    }
    if constexpr (EPI == EPI_LLMINT8) {
        // outlier flags are self-cleaning: the last CTA to finish clears them for the next call
        if (args.flags != nullptr && !args.keep_flags && args.flags[args.K] != 0) {
            int *s_last = reinterpret_cast<int *>(tmem_holder) + 1;   // spare word next to the TMEM handle
            if (threadIdx.x == 0) {
                __threadfence();
                *s_last = (atomicAdd(&args.flags[args.K + 1], 1) == (int)gridDim.x - 1);
            }
            __syncthreads();
            if (*s_last) {
                for (int c = threadIdx.x; c < args.K; c += blockDim.x) args.flags[c] = 0;
                __syncthreads();
                if (threadIdx.x == 0) {
                    args.flags[args.K + 1] = 0;
                    __threadfence();
                    args.flags[args.K] = 0;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor maps + dispatch
// ---------------------------------------------------------------------------------------------
// TMA-store boxes per epilogue warp (double-buffered when shared memory allows)
template <int BN, int BMODE> constexpr int pick_out_bufs() {
    return (is_byte<BMODE>() && BN == 128) ? 1 : 2;
}

// deepest smem ring that fits next to the epilogue staging boxes
template <int BN, int BMODE, int OUT_BUFS, int WS, int COLS, int LEAN, int PAIR = 0>
constexpr int pick_stages() {
    int best = 2;
    // lean tiles stop at 128 KB: they are meant to share an SM with another stream's resident CTAs
    const int budget = LEAN ? 131072 : 232448;
    for (int st = 2; st <= (LEAN >= 2 ? 12 : (PAIR ? 8 : 6)); ++st) {
        const int bmt = (COLS || LEAN) ? BMH : BM, bnt = COLS ? 2 * BN : BN;
        const int brows = PAIR ? bnt / 2 : bnt;     // W rows this CTA stages
        const int stage = (LEAN ? lean_a_rows<LEAN>() : bmt) * ROW_BYTES + (WS > 0 ? 0 : brows * ROW_BYTES) + BN * (is_byte<BMODE>() ? 64 : (is_nibble<BMODE>() ? 32 : 0));
        const int total = st * stage + WS * brows * ROW_BYTES + epi_warps<BN, BMODE, LEAN>() * OUT_BUFS * BOX_BYTES +
                          (WS > 0 ? 1 : 2) * 3 * bnt * 4 + 64 + (4 * st + 2 * ACC_STAGES + 1) * 8 + 16 + 1024;
        if (total <= budget) best = st;
    }
    return best;
}

template <typename OutT> CUtensorMapDataType out_dtype_enum();
template <> CUtensorMapDataType out_dtype_enum<float>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT32; }
template <> CUtensorMapDataType out_dtype_enum<__half>() { return CU_TENSOR_MAP_DATA_TYPE_FLOAT16; }
template <> CUtensorMapDataType out_dtype_enum<__nv_bfloat16>() { return CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; }

template <int BN, int AKIND, int BMODE, int EPI, typename OutT, int WS = 0, int COLS = 0, int LEAN = 0, int PAIR = 0>
int launch_gemm(const CUtensorMap &ma, const CUtensorMap &mb, GemmArgs args, cudaStream_t stream) {
    // a resident 256-row W tile (128 KB) leaves room for single store buffers only
#ifndef WQ_LEAN_OUT_BUFS
#define WQ_LEAN_OUT_BUFS 1     /* measured: 109.6 vs 110.7 ms per bench step with double buffers (smaller CTA beside the attention stream) */
#endif
#ifndef WQ_PAIR_OUT_BUFS
#define WQ_PAIR_OUT_BUFS 1
#endif
#ifndef WQ_PAIR_WS_OUT_BUFS
#define WQ_PAIR_WS_OUT_BUFS 2     /* weight-stationary pairs are bound by the epilogue's stores (K <= 512: 2 048 tensor clocks per tile) */
#endif
    constexpr int OUT_BUFS = PAIR ? ((WS > 0 && BMODE == B_DIRECT) ? WQ_PAIR_WS_OUT_BUFS : WQ_PAIR_OUT_BUFS) : ((WS > 0 && COLS) ? 1 : (LEAN ? WQ_LEAN_OUT_BUFS : pick_out_bufs<BN, BMODE>()));
    constexpr int STAGES = pick_stages<BN, BMODE, OUT_BUFS, WS, COLS, LEAN, PAIR>();
    using L = SmemLayout<BN, STAGES, BMODE, OUT_BUFS, WS, COLS, LEAN, PAIR>;
    auto kfn = k_gemm_tc<BN, STAGES, AKIND, BMODE, EPI, OutT, OUT_BUFS, WS, COLS, LEAN, PAIR>;
    static bool configured = false;
    static int max_pairs = 0;           // CTA pairs the device holds at once (persistent grid of the pair schedule)
    if (!configured) {
        WQ_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
        if (PAIR) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(2 * wq_sm_count());
            cfg.blockDim = dim3(num_threads<BN, BMODE, LEAN>());
            cfg.dynamicSmemBytes = (size_t)L::TOTAL;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, kfn, &cfg) != cudaSuccess || n < 1) {
                (void)cudaGetLastError();
                n = wq_sm_count() / 2;
            }
            max_pairs = n < wq_sm_count() / 2 ? n : wq_sm_count() / 2;
        }
        configured = true;
    }
    args.tiles_m = (args.M + L::TILE_M - 1) / L::TILE_M;
    args.tiles_n = (args.N + L::BNT - 1) / L::BNT;
    if (args.ldy == 0) args.ldy = args.N;
    // TMA store needs a 16-byte-aligned row pitch; otherwise the epilogue stores directly
    args.tma_store = (!args.no_store && ((size_t)args.ldy * sizeof(OutT)) % 16 == 0 && wq_aligned(args.out, 16)) ? 1 : 0;
    CUtensorMap my;
    if (args.tma_store) {
        int rc = make_map_2d(&my, args.out, out_dtype_enum<OutT>(), (int)sizeof(OutT), args.M, args.N, 32,
                             128 / (int)sizeof(OutT), CU_TENSOR_MAP_SWIZZLE_128B, (uint64_t)args.ldy);
        if (rc != WQ_OK) return rc;
    } else {
        my = ma;  // unused
    }
    const int total = args.tiles_m * args.tiles_n;
    if (PAIR) {
        int pairs = total < max_pairs ? total : max_pairs;
        // whole groups of tiles_n pairs (use_ws_pair()); if fewer pairs than one group can be resident at once (SMs
        // withheld from this context) the surplus pairs simply queue: no pair waits on another
        if (WS > 0) pairs = (max_pairs / args.tiles_n > 0 ? max_pairs / args.tiles_n : 1) * args.tiles_n;
        WQ_REQUIRE(pairs >= 1, "wq gemm: no CTA pair fits the device");
        WQ_LAUNCH_PDL_CLUSTER(2, kfn, dim3(2 * pairs), dim3(num_threads<BN, BMODE, LEAN>()), (size_t)L::TOTAL, stream, ma, mb,
                              my, args);
        return WQ_OK;
    }
    int grid = total < wq_sm_count() ? total : wq_sm_count();
    if (WS > 0) grid = (wq_sm_count() / args.tiles_n) * args.tiles_n;   // whole groups of tiles_n CTAs (use_ws())

    WQ_LAUNCH_PDL(kfn, dim3(grid), dim3(num_threads<BN, BMODE, LEAN>()), (size_t)L::TOTAL, stream, ma, mb, my, args);
    return WQ_OK;
}

// BN = 64 when a 128-wide tiling would leave most SMs idle (decode-shaped calls).
bool use_narrow_tile(int64_t M, int64_t N) {
    const int64_t tiles128 = ((M + BM - 1) / BM) * ((N + 127) / 128);  // BM = 256
    return tiles128 < wq_sm_count();
}

// Lean 128-row tile for decode-shaped calls with at most 128 rows (WQ_GEMM_LEAN=0 disables, A/B measurements).
bool use_lean_tile(int64_t M) {
    static const bool on = [] {
        const char *e = getenv("WQ_GEMM_LEAN");
        return e == nullptr || e[0] != '0';
    }();
    return on && M <= BMH;
}

// Weight-stationary schedule for the int8 x int8 schemes: K fits the resident W tile (<= kWS k-blocks of 128),
// every SM group has several row blocks to walk, and the column blocks divide the grid into whole groups.
constexpr int kWS = 4;
// L2 prefetch distance (tiles) of the weight-stationary schedule.  Measured on B200, M = 384000, K = 512
// (scripts/gemm_ws_bench.py): N = 512: 144 us round-robin -> 130 us stationary -> 118 us with the prefetch;
// N = 2048: 456 -> 448 -> 422 us.  A deeper ring instead (4 stages, single store buffers) gave 126 / 447 us.
constexpr int kWSPrefetch = 2;
bool ws_enabled() {     // WQ_GEMM_WS=0: round-robin tiles everywhere (A/B measurements, scripts/gemm_ws_bench.py)
    static const bool on = [] {
        const char *e = getenv("WQ_GEMM_WS");
        return e == nullptr || e[0] != '0';
    }();
    return on;
}
bool use_ws(int64_t M, int64_t N, int64_t K, bool cols) {
    const int64_t bmt = cols ? BMH : BM, bnt = cols ? 256 : 128;
    const int64_t tiles_n = (N + bnt - 1) / bnt, tiles_m = (M + bmt - 1) / bmt, kb = (K + 127) / 128;
    const int64_t sms = wq_sm_count();
    return ws_enabled() && kb <= kWS && tiles_n <= sms && tiles_m >= 4 * (sms / tiles_n);
}
// 128 x 256 tiles with one N = 256 MMA per k-step (COLS) for the int8 x int8 schemes: when the call is not
// decode-shaped and the 256-wide column blocks add no padding over 128-wide ones.  Measured on B200 at M = 384000
// (scripts/gemm_ws_bench.py, COLS vs 256 x 128 tiles): K = 2048, N = 512: 347 vs 368 us; K = 512 (both
// weight-stationary): N = 2048: 413 vs 421, N = 1536: 320 vs 331, N = 1024: 224 vs 220, N = 512: 132 vs 118 us -- the
// resident 256-row W tile leaves the A ring 48 KB, so narrow weight-stationary calls keep the 256 x 128 tile.
// WQ_GEMM_COLS=0 disables (A/B).
bool use_cols(int64_t M, int64_t N, int64_t K) {
    static const bool on = [] {
        const char *e = getenv("WQ_GEMM_COLS");
        return e == nullptr || e[0] != '0';
    }();
    if (!on || use_narrow_tile(M, N) || ((N + 127) / 128) % 2 != 0) return false;
    return !(use_ws(M, N, K, false) && N < 1536);
}

// CTA pairs (256 x 256 per cluster of two, one cta_group::2 instruction per k-step) for the int8 x int8 schemes: every
// call that is not decode-shaped and whose 256-wide column blocks add no padding over 128-wide ones.  WQ_GEMM_PAIR=0
// falls back to the single-CTA tiles (A/B measurements, scripts/gemm_ws_bench.py).
// WQ_GEMM_PAIR: 0 = no pairs, 1 = int8 x int8 schemes only, 2 = + the weight-expanding schemes (W8A16, W4A16, u4, e4m3),
// 3 (default) = + the unquantized fp16 / bf16 projection.
bool use_pair(int64_t M, int64_t N, int level = 1) {
    static const int max_level = [] {
        const char *e = getenv("WQ_GEMM_PAIR");
        return e == nullptr ? 3 : atoi(e);
    }();
    const bool on = level <= max_level;
    // (at most 128 rows: the peer CTA of every pair would hold no rows at all -- wide decode-shaped calls such as a
    // quantized vocabulary projection keep the single-CTA tiles)
    return on && M > BMH && !use_narrow_tile(M, N) && ((N + 127) / 128) % 2 == 0;
}
// weight-stationary pairs: K fits the resident half tiles (kWS k-blocks x 128 W rows = 64 KB per CTA), the column
// blocks divide the pairs into whole groups and every group has several row blocks to walk
bool use_ws_pair(int64_t M, int64_t N, int64_t K) {
    const int64_t tiles_n = (N + 255) / 256, tiles_m = (M + 255) / 256, kb = (K + 127) / 128;
    const int64_t pairs = wq_sm_count() / 2;
    return ws_enabled() && kb <= kWS && tiles_n <= pairs && tiles_m >= 4 * (pairs / tiles_n);
}

// tensor maps + launch of the int8 x int8 schemes (LLM.int8, torch dynamic): tile shape and schedule by shape
template <int AKIND, int EPI, typename OutT>
int launch_i8(const void *a, const void *b, GemmArgs args, cudaStream_t s) {
    const int64_t M = args.M, N = args.N, K = args.K;
    if (use_pair(M, N)) {
        CUtensorMap ma, mb;
        int rc = make_map_2d(&ma, a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, M, K, BMH, 128, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != WQ_OK) return rc;
        rc = make_map_2d(&mb, b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K, 128, 128, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != WQ_OK) return rc;
        if (use_ws_pair(M, N, K)) {
            args.prefetch = kWSPrefetch;
            return launch_gemm<128, AKIND, B_DIRECT, EPI, OutT, kWS, 1, 0, 1>(ma, mb, args, s);
        }
        return launch_gemm<128, AKIND, B_DIRECT, EPI, OutT, 0, 1, 0, 1>(ma, mb, args, s);
    }
    const bool narrow = use_narrow_tile(M, N), cols = use_cols(M, N, K);
    const bool lean = narrow && use_lean_tile(M);
    static const bool slim_on = [] {        // WQ_GEMM_SLIM=0: 128-row A boxes for every lean call (A/B measurements)
        const char *e = getenv("WQ_GEMM_SLIM");
        return e == nullptr || e[0] != '0';
    }();
    const int slim = (lean && slim_on) ? (M <= 32 ? 2 : (M <= 64 ? 3 : 0)) : 0;
    CUtensorMap ma, mb;
    int rc = make_map_2d(&ma, a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, M, K,
                         slim == 2 ? 32 : (slim == 3 ? 64 : ((cols || lean) ? BMH : BM)), 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != WQ_OK) return rc;
    rc = make_map_2d(&mb, b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K, narrow ? 64 : (cols ? 256 : 128), 128,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != WQ_OK) return rc;
    if (slim == 2) return launch_gemm<64, AKIND, B_DIRECT, EPI, OutT, 0, 0, 2>(ma, mb, args, s);
    if (slim == 3) return launch_gemm<64, AKIND, B_DIRECT, EPI, OutT, 0, 0, 3>(ma, mb, args, s);
    if (lean) return launch_gemm<64, AKIND, B_DIRECT, EPI, OutT, 0, 0, 1>(ma, mb, args, s);
    if (narrow) return launch_gemm<64, AKIND, B_DIRECT, EPI, OutT>(ma, mb, args, s);
    const bool ws = use_ws(M, N, K, cols);
    if (cols) {
        if (ws) {
            args.prefetch = 2 * kWSPrefetch;        // 128-row tiles: same look-ahead in rows
            return launch_gemm<128, AKIND, B_DIRECT, EPI, OutT, kWS, 1>(ma, mb, args, s);
        }
        return launch_gemm<128, AKIND, B_DIRECT, EPI, OutT, 0, 1>(ma, mb, args, s);
    }
    if (ws) {
        args.prefetch = kWSPrefetch;
        return launch_gemm<128, AKIND, B_DIRECT, EPI, OutT, kWS>(ma, mb, args, s);
    }
    return launch_gemm<128, AKIND, B_DIRECT, EPI, OutT>(ma, mb, args, s);
}

int check_common(const char *fn, int64_t M, int64_t N, int64_t K) {
    WQ_REQUIRE(M >= 0 && N >= 0 && K > 0, "%s: bad shape M=%lld N=%lld K=%lld", fn, (long long)M, (long long)N,
               (long long)K);
    WQ_REQUIRE(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "%s: shape exceeds int32", fn);
    return wq_check_device();
}

}  // namespace


extern "C" int wq_gemm_llmint8(const int8_t *ca, const float *sca, const int8_t *cb, const float *scb,
                               const float *bias, void *y_f16, int64_t M, int64_t N, int64_t K, const void *a_f16,
                               int32_t *col_flags, wq_stream_t stream) {
    return wq_gemm_llmint8_shared(ca, sca, cb, scb, bias, y_f16, M, N, K, a_f16, col_flags, 0, stream);
}

extern "C" int wq_gemm_llmint8_shared(const int8_t *ca, const float *sca, const int8_t *cb, const float *scb,
                                      const float *bias, void *y_f16, int64_t M, int64_t N, int64_t K,
                                      const void *a_f16, int32_t *col_flags, int keep_flags, wq_stream_t stream) {
    return wq_gemm_llmint8_residual(ca, sca, cb, scb, bias, y_f16, M, N, K, a_f16, col_flags, keep_flags, nullptr, 0.0f,
                                    0, stream);
}

extern "C" int wq_gemm_llmint8_residual(const int8_t *ca, const float *sca, const int8_t *cb, const float *scb,
                                        const float *bias, void *y_f16, int64_t M, int64_t N, int64_t K,
                                        const void *a_f16, int32_t *col_flags, int keep_flags,
                                        const void *residual_f16, float clamp_abs, int a_pre_gelu,
                                        wq_stream_t stream) {
    int rc = check_common("wq_gemm_llmint8", M, N, K);
    if (rc != WQ_OK) return rc;
    if (M == 0 || N == 0) return WQ_OK;
    WQ_REQUIRE(ca && sca && cb && scb && y_f16, "wq_gemm_llmint8: null pointer");
    WQ_REQUIRE(K % 16 == 0, "wq_gemm_llmint8: K=%lld must be a multiple of 16", (long long)K);
    WQ_REQUIRE(wq_aligned(ca, 16) && wq_aligned(cb, 16) && wq_aligned(y_f16, 16), "wq_gemm_llmint8: misaligned buffer");
    WQ_REQUIRE(col_flags == nullptr || a_f16 != nullptr, "wq_gemm_llmint8: the outlier path needs a_f16");
    cudaStream_t s = (cudaStream_t)stream;
    GemmArgs args = {};
    args.M = (int)M; args.N = (int)N; args.K = (int)K;
    args.num_kb = (int)((K + 127) / 128);
    args.row_scale = sca; args.col_scale = scb; args.bias = bias; args.out = y_f16;
    args.ca = ca; args.cb = cb; args.a16 = (const __half *)a_f16; args.flags = col_flags;
    args.keep_flags = keep_flags ? 1 : 0;
    args.a_pre_gelu = a_pre_gelu ? 1 : 0;
    WQ_REQUIRE(residual_f16 == nullptr || (wq_aligned(residual_f16, 16) && N % 8 == 0),
               "wq_gemm_llmint8: the residual needs 16-byte aligned rows (N %% 8 == 0)");
    WQ_REQUIRE(clamp_abs >= 0.0f, "wq_gemm_llmint8: negative clamp");
    args.residual = residual_f16; args.clamp_abs = clamp_abs;
    return launch_i8<A_S8, EPI_LLMINT8, __half>(ca, cb, args, s);
}

extern "C" int wq_gemm_dyn_i8(const uint8_t *xq, const float *qparams, const int8_t *wq, const float *w_scale,
                              const int32_t *wsum, const float *bias, float *y, int64_t M, int64_t N, int64_t K,
                              wq_stream_t stream) {
    int rc = check_common("wq_gemm_dyn_i8", M, N, K);
    if (rc != WQ_OK) return rc;
    if (M == 0 || N == 0) return WQ_OK;
    WQ_REQUIRE(xq && qparams && wq && w_scale && wsum && y, "wq_gemm_dyn_i8: null pointer");
    WQ_REQUIRE(K % 16 == 0, "wq_gemm_dyn_i8: K=%lld must be a multiple of 16", (long long)K);
    WQ_REQUIRE(wq_aligned(xq, 16) && wq_aligned(wq, 16) && wq_aligned(y, 16), "wq_gemm_dyn_i8: misaligned buffer");
    cudaStream_t s = (cudaStream_t)stream;
    GemmArgs args = {};
    args.M = (int)M; args.N = (int)N; args.K = (int)K;
    args.num_kb = (int)((K + 127) / 128);
    args.qparams = qparams; args.w_scale = w_scale; args.wsum = wsum; args.bias = bias; args.out = y;
    return launch_i8<A_U8, EPI_DYN, float>(xq, wq, args, s);
}

namespace {

// pair: the CTA-pair schedule (use_pair(); the A map then has 128-row boxes, the W map 128-row boxes as before)
template <int BMODE, int EPI>
int dispatch_a16(const CUtensorMap &ma, const CUtensorMap &mb, const GemmArgs &args, int x_dtype, int y_dtype,
                 bool narrow, cudaStream_t s, bool lean = false, bool pair = false) {
#define WQ_CASE(BN, AK, OT) return launch_gemm<BN, AK, BMODE, EPI, OT>(ma, mb, args, s)
#define WQ_LEAN(AK, OT) return launch_gemm<64, AK, BMODE, EPI, OT, 0, 0, 1>(ma, mb, args, s)
#define WQ_PAIR(AK, OT) return launch_gemm<128, AK, BMODE, EPI, OT, 0, 1, 0, 1>(ma, mb, args, s)
    if (x_dtype == WQ_F16) {
        if (y_dtype == WQ_F16) { if (pair) WQ_PAIR(A_F16, __half); if (lean) WQ_LEAN(A_F16, __half); if (narrow) WQ_CASE(64, A_F16, __half); else WQ_CASE(128, A_F16, __half); }
        if (y_dtype == WQ_F32) { if (pair) WQ_PAIR(A_F16, float); if (lean) WQ_LEAN(A_F16, float); if (narrow) WQ_CASE(64, A_F16, float); else WQ_CASE(128, A_F16, float); }
    } else if (x_dtype == WQ_BF16) {
        if (y_dtype == WQ_BF16) { if (pair) WQ_PAIR(A_BF16, __nv_bfloat16); if (lean) WQ_LEAN(A_BF16, __nv_bfloat16); if (narrow) WQ_CASE(64, A_BF16, __nv_bfloat16); else WQ_CASE(128, A_BF16, __nv_bfloat16); }
        if (y_dtype == WQ_F32) { if (pair) WQ_PAIR(A_BF16, float); if (narrow) WQ_CASE(64, A_BF16, float); else WQ_CASE(128, A_BF16, float); }
    }
#undef WQ_CASE
#undef WQ_LEAN
#undef WQ_PAIR
    wq_set_error("unsupported dtype combination x=%d y=%d (x: F16/BF16, y: same as x or F32)", x_dtype, y_dtype);
    return WQ_ERR_INVALID;
}

}  // namespace

extern "C" int wq_gemm_w8a16(const void *x, int x_dtype, const int8_t *wq, const float *scale, const float *bias,
                             void *y, int y_dtype, int64_t M, int64_t N, int64_t K, wq_stream_t stream) {
    int rc = check_common("wq_gemm_w8a16", M, N, K);
    if (rc != WQ_OK) return rc;
    if (M == 0 || N == 0) return WQ_OK;
    WQ_REQUIRE(x && wq && scale && y, "wq_gemm_w8a16: null pointer");
    WQ_REQUIRE(K % 16 == 0, "wq_gemm_w8a16: K=%lld must be a multiple of 16", (long long)K);
    WQ_REQUIRE(wq_aligned(x, 16) && wq_aligned(wq, 16) && wq_aligned(y, 16), "wq_gemm_w8a16: misaligned buffer");
    GemmArgs args = {};
    args.M = (int)M; args.N = (int)N; args.K = (int)K;
    args.num_kb = (int)((K + 63) / 64);
    args.col_scale = scale; args.bias = bias; args.out = y;
    const bool narrow = use_narrow_tile(M, N), pair = use_pair(M, N, 2);
    // decode-shaped calls, <= 128 rows (fp32 output of fp16 operands: the decode steps of the reference's fp32 flows)
    const bool lean = narrow && use_lean_tile(M) && (y_dtype == x_dtype || (x_dtype == WQ_F16 && y_dtype == WQ_F32));
    CUtensorMap ma, mb;
    rc = make_map_2d(&ma, x, x_dtype == WQ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     M, K, (lean || pair) ? BMH : BM, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != WQ_OK) return rc;
    rc = make_map_2d(&mb, wq, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K, narrow ? 64 : 128, 64,
                     CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != WQ_OK) return rc;
    return dispatch_a16<B_I8, EPI_W8A16>(ma, mb, args, x_dtype, y_dtype, narrow, (cudaStream_t)stream, lean, pair);
}

extern "C" int wq_gemm_w4a16(const void *x, int x_dtype, const uint8_t *packed, const float *absmax, int quant_type,
                             const float *bias, void *y, int y_dtype, int64_t M, int64_t N, int64_t K,
                             wq_stream_t stream) {
    int rc = check_common("wq_gemm_w4a16", M, N, K);
    if (rc != WQ_OK) return rc;
    if (M == 0 || N == 0) return WQ_OK;
    WQ_REQUIRE(x && packed && absmax && y, "wq_gemm_w4a16: null pointer");
    WQ_REQUIRE(K % 64 == 0, "wq_gemm_w4a16: K=%lld must be a multiple of the 64-element block", (long long)K);
    WQ_REQUIRE(quant_type == WQ_NF4 || quant_type == WQ_FP4, "wq_gemm_w4a16: bad quant_type");
    WQ_REQUIRE(wq_aligned(x, 16) && wq_aligned(packed, 16) && wq_aligned(y, 16), "wq_gemm_w4a16: misaligned buffer");
    GemmArgs args = {};
    args.M = (int)M; args.N = (int)N; args.K = (int)K;
    args.num_kb = (int)(K / 64);
    args.absmax = absmax; args.absmax_ld = (int)(K / 64);
    args.bias = bias; args.out = y; args.quant_type = quant_type;
    const bool narrow = use_narrow_tile(M, N), pair = use_pair(M, N, 2);
    // decode-shaped calls, <= 128 rows (fp32 output of fp16 operands: the decode steps of the reference's fp32 flows)
    const bool lean = narrow && use_lean_tile(M) && (y_dtype == x_dtype || (x_dtype == WQ_F16 && y_dtype == WQ_F32));
    CUtensorMap ma, mb;
    rc = make_map_2d(&ma, x, x_dtype == WQ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     M, K, (lean || pair) ? BMH : BM, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != WQ_OK) return rc;
    rc = make_map_2d(&mb, packed, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K / 2, narrow ? 64 : 128, 32,
                     CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != WQ_OK) return rc;
    return dispatch_a16<B_4BIT, EPI_W4A16>(ma, mb, args, x_dtype, y_dtype, narrow, (cudaStream_t)stream, lean, pair);
}

/* quanto QLinear.forward with weights=qint4 (group-wise affine uint4, MaxOptimizer):
 * y = x @ (scale * q - shift)^T + bias, dequantised weight rounded once to the operand dtype. */
extern "C" int wq_gemm_u4a16(const void *x, int x_dtype, const uint8_t *packed, const float *scale, const float *shift,
                             int group, const float *bias, void *y, int y_dtype, int64_t M, int64_t N, int64_t K,
                             wq_stream_t stream) {
    int rc = check_common("wq_gemm_u4a16", M, N, K);
    if (rc != WQ_OK) return rc;
    if (M == 0 || N == 0) return WQ_OK;
    WQ_REQUIRE(x && packed && scale && shift && y, "wq_gemm_u4a16: null pointer");
    WQ_REQUIRE(K % 64 == 0, "wq_gemm_u4a16: K=%lld must be a multiple of 64", (long long)K);
    WQ_REQUIRE(group >= 32 && group % 32 == 0 && K % group == 0, "wq_gemm_u4a16: group %d must be a multiple of 32 dividing K", group);
    WQ_REQUIRE(wq_aligned(x, 16) && wq_aligned(packed, 16) && wq_aligned(y, 16), "wq_gemm_u4a16: misaligned buffer");
    GemmArgs args = {};
    args.M = (int)M; args.N = (int)N; args.K = (int)K;
    args.num_kb = (int)(K / 64);
    args.absmax = scale; args.shift = shift; args.group = group; args.absmax_ld = (int)(K / group);
    args.bias = bias; args.out = y;
    const bool narrow = use_narrow_tile(M, N), pair = use_pair(M, N, 2);
    // decode-shaped calls, <= 128 rows (fp32 output of fp16 operands: the decode steps of the reference's fp32 flows)
    const bool lean = narrow && use_lean_tile(M) && (y_dtype == x_dtype || (x_dtype == WQ_F16 && y_dtype == WQ_F32));
    CUtensorMap ma, mb;
    rc = make_map_2d(&ma, x, x_dtype == WQ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     M, K, (lean || pair) ? BMH : BM, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != WQ_OK) return rc;
    rc = make_map_2d(&mb, packed, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K / 2, narrow ? 64 : 128, 32,
                     CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != WQ_OK) return rc;
    return dispatch_a16<B_U4, EPI_W4A16>(ma, mb, args, x_dtype, y_dtype, narrow, (cudaStream_t)stream, lean, pair);
}

/* Unquantized linear on the same tcgen05 pipeline: y = x @ W^T + bias with W [N, K] in the activation dtype, taken
 * straight from TMA (no expansion warps).  Serves the vocabulary projection the HF bitsandbytes flows keep in fp16
 * (proj_out, modeling_whisper.py:971,1081).  With argmax_keys the greedy choice (torch.argmax of the rounded logits
 * under a suppression mask) is folded into the epilogue; y may then be NULL: the logits never reach HBM. */
extern "C" int wq_gemm_f16(const void *x, int x_dtype, const void *w, const float *bias, void *y, int y_dtype,
                           int64_t ldy, int64_t M, int64_t N, int64_t K, const uint8_t *mask, int64_t mask_len,
                           unsigned long long *argmax_keys, wq_stream_t stream) {
    int rc = check_common("wq_gemm_f16", M, N, K);
    if (rc != WQ_OK) return rc;
    if (M == 0 || N == 0) return WQ_OK;
    WQ_REQUIRE(x && w, "wq_gemm_f16: null pointer");
    WQ_REQUIRE(y != nullptr || argmax_keys != nullptr, "wq_gemm_f16: neither an output nor arg-max keys requested");
    WQ_REQUIRE(K % 8 == 0, "wq_gemm_f16: K=%lld must be a multiple of 8", (long long)K);
    WQ_REQUIRE(wq_aligned(x, 16) && wq_aligned(w, 16) && (y == nullptr || wq_aligned(y, 16)), "wq_gemm_f16: misaligned buffer");
    if (ldy == 0) ldy = N;
    WQ_REQUIRE(ldy >= N && ldy < (1ll << 31), "wq_gemm_f16: bad output pitch %lld", (long long)ldy);
    GemmArgs args = {};
    args.M = (int)M; args.N = (int)N; args.K = (int)K;
    args.num_kb = (int)((K + 63) / 64);
    args.bias = bias; args.out = y; args.ldy = (int)ldy;
    args.argmax_keys = argmax_keys; args.mask = mask; args.no_store = y == nullptr ? 1 : 0;
    const bool narrow = use_narrow_tile(M, N);
    if (mask != nullptr) {
        const int64_t bnt = narrow ? 64 : 128;
        WQ_REQUIRE(argmax_keys != nullptr, "wq_gemm_f16: a mask needs argmax_keys");
        WQ_REQUIRE(wq_aligned(mask, 16) && mask_len >= ((N + bnt - 1) / bnt) * bnt,
                   "wq_gemm_f16: mask must be 16-byte aligned and padded to whole %lld-column tiles", (long long)bnt);
    }
    const CUtensorMapDataType dt = x_dtype == WQ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUtensorMap ma, mb;
    const bool pair = use_pair(M, N, 3);
    rc = make_map_2d(&ma, x, dt, 2, M, K, pair ? BMH : BM, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != WQ_OK) return rc;
    rc = make_map_2d(&mb, w, dt, 2, N, K, narrow ? 64 : 128, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != WQ_OK) return rc;
    return dispatch_a16<B_DIRECT, EPI_PLAIN>(ma, mb, args, x_dtype, y == nullptr ? x_dtype : y_dtype, narrow,
                                             (cudaStream_t)stream, false, pair);
}

namespace {
__global__ void k_argmax_finalize(unsigned long long *keys, int64_t M, int64_t *out) {
    pdl_prologue_done();
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < M) {
        out[m] = (int64_t)(0xFFFFFFFFu - (uint32_t)(keys[m] & 0xFFFFFFFFull));
        keys[m] = 0ull;      // ready for the next projection
    }
}
}  // namespace

/* Token ids from the arg-max keys of wq_gemm_f16 (and reset of the keys): out[m] = column of the maximum. */
extern "C" int wq_argmax_finalize(unsigned long long *keys, int64_t M, int64_t *out, wq_stream_t stream) {
    WQ_REQUIRE(M >= 0, "wq_argmax_finalize: bad shape");
    if (M == 0) return WQ_OK;
    WQ_REQUIRE(keys && out, "wq_argmax_finalize: null pointer");
    WQ_LAUNCH_PDL(k_argmax_finalize, dim3((unsigned)((M + 127) / 128)), dim3(128), 0, (cudaStream_t)stream, keys, M, out);
    return WQ_OK;
}


/* quanto QLinear.forward with weights=qfloat8 (e4m3fn codes, per-output-channel scale):
 * y = matmul(x, Wq.to(x.dtype).t()) * scale + bias -- the same pipeline as wq_gemm_w8a16, the packed byte tile is
 * expanded e4m3 -> fp16 / bf16 (exact) by the expansion warps. */
extern "C" int wq_gemm_wf8a16(const void *x, int x_dtype, const uint8_t *wq, const float *scale, const float *bias,
                              void *y, int y_dtype, int64_t M, int64_t N, int64_t K, wq_stream_t stream) {
    int rc = check_common("wq_gemm_wf8a16", M, N, K);
    if (rc != WQ_OK) return rc;
    if (M == 0 || N == 0) return WQ_OK;
    WQ_REQUIRE(x && wq && scale && y, "wq_gemm_wf8a16: null pointer");
    WQ_REQUIRE(K % 16 == 0, "wq_gemm_wf8a16: K=%lld must be a multiple of 16", (long long)K);
    WQ_REQUIRE(wq_aligned(x, 16) && wq_aligned(wq, 16) && wq_aligned(y, 16), "wq_gemm_wf8a16: misaligned buffer");
    GemmArgs args = {};
    args.M = (int)M; args.N = (int)N; args.K = (int)K;
    args.num_kb = (int)((K + 63) / 64);
    args.col_scale = scale; args.bias = bias; args.out = y;
    const bool narrow = use_narrow_tile(M, N), pair = use_pair(M, N, 2);
    CUtensorMap ma, mb;
    rc = make_map_2d(&ma, x, x_dtype == WQ_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     M, K, pair ? BMH : BM, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != WQ_OK) return rc;
    rc = make_map_2d(&mb, wq, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, N, K, narrow ? 64 : 128, 64, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != WQ_OK) return rc;
    return dispatch_a16<B_F8, EPI_W8A16>(ma, mb, args, x_dtype, y_dtype, narrow, (cudaStream_t)stream, false, pair);
}

/* quanto QLinear.forward with qint8 weights AND statically quantized qint8 activations (quantize(model, weights=qint8,
 * activations=qint8) + Calibration, model_utils.py:152-214): qbytes_int_mm, i.e.
 *   y = float(int32(xq . wq^T)) * out_scale[n] + bias[n],   out_scale[n] = input_scale * weight_scale[n]
 * on the kind::i8 tensor-core path.  xq int8 [M, K]; wq int8 [N, K]; out_scale, bias fp32 [N]; y of y_dtype. */
extern "C" int wq_gemm_w8a8(const int8_t *xq, const int8_t *wq, const float *out_scale, const float *bias, void *y,
                            int y_dtype, int64_t M, int64_t N, int64_t K, wq_stream_t stream) {
    int rc = check_common("wq_gemm_w8a8", M, N, K);
    if (rc != WQ_OK) return rc;
    if (M == 0 || N == 0) return WQ_OK;
    WQ_REQUIRE(xq && wq && out_scale && y, "wq_gemm_w8a8: null pointer");
    WQ_REQUIRE(K % 16 == 0, "wq_gemm_w8a8: K=%lld must be a multiple of 16", (long long)K);
    WQ_REQUIRE(wq_aligned(xq, 16) && wq_aligned(wq, 16) && wq_aligned(y, 16), "wq_gemm_w8a8: misaligned buffer");
    GemmArgs args = {};
    args.M = (int)M; args.N = (int)N; args.K = (int)K;
    args.num_kb = (int)((K + 127) / 128);
    args.col_scale = out_scale; args.bias = bias; args.out = y;
    cudaStream_t s = (cudaStream_t)stream;
    if (y_dtype == WQ_F32) return launch_i8<A_S8, EPI_W8A8, float>(xq, wq, args, s);
    if (y_dtype == WQ_F16) return launch_i8<A_S8, EPI_W8A8, __half>(xq, wq, args, s);
    if (y_dtype == WQ_BF16) return launch_i8<A_S8, EPI_W8A8, __nv_bfloat16>(xq, wq, args, s);
    wq_set_error("wq_gemm_w8a8: bad y_dtype %d", y_dtype);
    return WQ_ERR_INVALID;
}
