// bnb_open_points.cu -- closes the bitsandbytes open points of SURVEY.md Appendix A.2 / DESIGN.md section 4 on a real GPU.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_build/bnb_open_points scripts/bnb_open_points.cu
//   scripts/_build/bnb_open_points gpurun_out/bnb_open_points.json gpurun_out/fdividef_127_table.bin
//
// (1) int8_vectorwise_quant: bitsandbytes' kernel computes the row scale with __fdividef(127.0f, absmax) (approximate
//     division), this repo's kernels and CPU oracle use the IEEE quotient.  EXHAUSTIVE over fp16: for every finite
//     positive fp16 absmax and every fp16 a with 0 <= a <= absmax (the codes are odd-symmetric in a), compare
//     __float2int_rn(a * __fdividef(127, absmax)) with __float2int_rn(a * __fdiv_rn(127, absmax)); count and list the
//     pairs that differ, and dump the table of __fdividef(127, h) over all fp16 bit patterns h (so a CPU oracle can
//     reproduce the approximate form exactly for fp16 inputs).
// (2) int8_mm_dequant: "x * c + bias" written as mul-then-add is contracted to one FFMA by nvcc's default
//     -fmad=true, i.e. it IS fmaf(x, c, bias); compare the two source forms (and the uncontracted one) over the
//     int32 accumulator range reachable at K <= 5120 (|acc| <= 127*127*5120) with random scales / biases.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Diff { unsigned short absmax, a; int q_approx, q_ieee; };

__global__ void k_sweep(unsigned long long *n_pairs, unsigned long long *n_diff, unsigned long long *n_scale_diff,
                        Diff *list, int cap, float *table) {
    // one block per absmax bit pattern (positive finite non-zero: 0x0001 .. 0x7bff)
    const unsigned short hb = (unsigned short)(blockIdx.x + 1);
    const float am = __half2float(__ushort_as_half(hb));
    const float s_approx = __fdividef(127.0f, am);
    const float s_ieee = __fdiv_rn(127.0f, am);
    if (threadIdx.x == 0) {
        table[hb] = s_approx;
        if (s_approx != s_ieee) atomicAdd(n_scale_diff, 1ull);
    }
    unsigned long long pairs = 0, diffs = 0;
    for (unsigned a = threadIdx.x; a <= hb; a += blockDim.x) {      // fp16 bit patterns are monotone for positives
        const float x = __half2float(__ushort_as_half((unsigned short)a));
        const int q1 = __float2int_rn(__fmul_rn(x, s_approx));
        const int q2 = __float2int_rn(__fmul_rn(x, s_ieee));
        ++pairs;
        if (q1 != q2) {
            ++diffs;
            const unsigned long long slot = atomicAdd(n_diff, 1ull);
            if (slot < (unsigned long long)cap) list[slot] = Diff{hb, (unsigned short)a, q1, q2};
        }
    }
    atomicAdd(n_pairs, pairs);
}

__device__ __forceinline__ uint32_t rng(uint32_t &s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

__global__ void k_dequant_forms(unsigned long long *n, unsigned long long *d_plain_vs_fma, unsigned long long *d_nofma_vs_fma,
                                unsigned long long *d_half_nofma) {
    uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1);
    unsigned long long c0 = 0, c1 = 0, c2 = 0, cnt = 0;
    const float c = 6.200012e-05f;
    for (int i = 0; i < 4096; ++i) {
        const int acc = (int)(rng(s) % (2u * 82580480u + 1u)) - 82580480;            // |acc| <= 127*127*5120
        const float sca = __half2float(__ushort_as_half((unsigned short)(rng(s) % 0x7bffu + 1)));   // row absmax (fp16 value)
        const float scb = __half2float(__ushort_as_half((unsigned short)(rng(s) % 0x3c00u + 1)));   // weight absmax <= 1
        const float bias = __half2float(__ushort_as_half((unsigned short)(rng(s) & 0xbbffu)));      // |bias| < 1, either sign
        const float x = __fmul_rn(__fmul_rn((float)acc, sca), scb);
        const float f_fma = fmaf(x, c, bias);
        const float f_plain = x * c + bias;                       // contracted by -fmad=true (the nvcc default)
        const float f_nofma = __fadd_rn(__fmul_rn(x, c), bias);   // what -fmad=false would compute
        ++cnt;
        c0 += (f_plain != f_fma) && !(f_plain != f_plain);
        c1 += (f_nofma != f_fma) && !(f_nofma != f_nofma);
        c2 += __half_as_ushort(__float2half_rn(f_nofma)) != __half_as_ushort(__float2half_rn(f_fma)) && !(f_fma != f_fma);
    }
    atomicAdd(n, cnt); atomicAdd(d_plain_vs_fma, c0); atomicAdd(d_nofma_vs_fma, c1); atomicAdd(d_half_nofma, c2);
}

int main(int argc, char **argv) {
    const char *json = argc > 1 ? argv[1] : "bnb_open_points.json";
    const char *tbl = argc > 2 ? argv[2] : nullptr;
    const int cap = 4096;
    unsigned long long *d_cnt;
    Diff *d_list;
    float *d_table;
    CK(cudaMalloc(&d_cnt, 8 * sizeof(unsigned long long)));
    CK(cudaMemset(d_cnt, 0, 8 * sizeof(unsigned long long)));
    CK(cudaMalloc(&d_list, cap * sizeof(Diff)));
    CK(cudaMalloc(&d_table, 65536 * sizeof(float)));
    CK(cudaMemset(d_table, 0, 65536 * sizeof(float)));
    k_sweep<<<0x7bff, 256>>>(d_cnt, d_cnt + 1, d_cnt + 2, d_list, cap, d_table);
    CK(cudaGetLastError());
    k_dequant_forms<<<148 * 16, 256>>>(d_cnt + 3, d_cnt + 4, d_cnt + 5, d_cnt + 6);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    unsigned long long h[8];
    CK(cudaMemcpy(h, d_cnt, sizeof(h), cudaMemcpyDeviceToHost));
    std::vector<Diff> list(cap);
    CK(cudaMemcpy(list.data(), d_list, cap * sizeof(Diff), cudaMemcpyDeviceToHost));
    std::vector<float> table(65536);
    CK(cudaMemcpy(table.data(), d_table, 65536 * sizeof(float), cudaMemcpyDeviceToHost));
    FILE *f = fopen(json, "w");
    if (!f) { printf("cannot open %s\n", json); return 1; }
    fprintf(f, "{\"fp16_pairs\": %llu, \"code_mismatches\": %llu, \"absmax_values_with_different_scale\": %llu,\n"
               " \"dequant_samples\": %llu, \"plain_vs_fmaf_f32_mismatches\": %llu, \"mul_then_add_vs_fmaf_f32_mismatches\": %llu,"
               " \"mul_then_add_vs_fmaf_fp16_output_mismatches\": %llu,\n \"first_mismatches\": [",
            h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
    const unsigned long long nl = h[1] < (unsigned long long)cap ? h[1] : cap;
    for (unsigned long long i = 0; i < nl && i < 64; ++i)
        fprintf(f, "%s{\"absmax_bits\": %u, \"a_bits\": %u, \"q_fdividef\": %d, \"q_ieee\": %d}", i ? ", " : "", list[i].absmax,
                list[i].a, list[i].q_approx, list[i].q_ieee);
    fprintf(f, "]}\n");
    fclose(f);
    if (tbl) {
        FILE *t = fopen(tbl, "wb");
        if (t) { fwrite(table.data(), sizeof(float), 65536, t); fclose(t); }
    }
    printf("fp16 pairs %llu, code mismatches %llu, absmax values whose scale differs %llu\n", h[0], h[1], h[2]);
    printf("dequant samples %llu: plain-vs-fmaf %llu, mul-then-add-vs-fmaf f32 %llu, fp16 output %llu\n", h[3], h[4], h[5], h[6]);
    return 0;
}
