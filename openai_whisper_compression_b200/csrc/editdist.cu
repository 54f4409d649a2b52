// editdist.cu -- unit-cost Levenshtein distance per (reference, hypothesis) pair of id sequences,
// the integer core of the WER / CER tallies (evaluate.load("wer"/"cer"), evaluation.py:110-116).
// One CTA per pair; anti-diagonal wavefront over three uint16 diagonals in shared memory.
#include "common.cuh"

namespace {

constexpr int ED_MAX = 4096;
constexpr int ED_THREADS = 128;

__global__ void __launch_bounds__(ED_THREADS)
k_edit_distance(const int32_t *__restrict__ ref, const int64_t *__restrict__ ref_off, const int32_t *__restrict__ hyp,
                const int64_t *__restrict__ hyp_off, int64_t *__restrict__ dist) {
    __shared__ uint16_t diag[3][ED_MAX + 1];
    const int p = blockIdx.x;
    const int32_t *r = ref + ref_off[p];
    const int32_t *h = hyp + hyp_off[p];
    const int nr = (int)(ref_off[p + 1] - ref_off[p]);
    const int nh = (int)(hyp_off[p + 1] - hyp_off[p]);
    if (nr == 0 || nh == 0) {
        if (threadIdx.x == 0) dist[p] = nr + nh;
        return;
    }
    if (nr > ED_MAX || nh > ED_MAX || nr < 0 || nh < 0) {
        // longer than the shared-memory diagonals (and than uint16 distances): never index past them -- the pair
        // gets the sentinel -1, which every caller must treat as an error (tally.py refuses such input up front)
        if (threadIdx.x == 0) dist[p] = -1;
        return;
    }
    // diagonal d holds D[i][d - i] at index i
    for (int d = 0; d <= nr + nh; ++d) {
        uint16_t *cur = diag[d % 3];
        const uint16_t *p1 = diag[(d + 2) % 3];  // d - 1
        const uint16_t *p2 = diag[(d + 1) % 3];  // d - 2
        const int ilo = max(0, d - nh), ihi = min(nr, d);
        for (int i = ilo + threadIdx.x; i <= ihi; i += ED_THREADS) {
            const int j = d - i;
            int v;
            if (i == 0) v = j;
            else if (j == 0) v = i;
            else {
                const int sub = p2[i - 1] + (r[i - 1] != h[j - 1] ? 1 : 0);
                const int del = p1[i - 1] + 1;  // D[i-1][j]
                const int ins = p1[i] + 1;      // D[i][j-1]
                v = min(sub, min(del, ins));
            }
            cur[i] = (uint16_t)v;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) dist[p] = diag[(nr + nh) % 3][nr];
}

}  // namespace

extern "C" int wq_edit_distance(const int32_t *ref, const int64_t *ref_off, const int32_t *hyp,
                                const int64_t *hyp_off, int64_t P, int64_t *dist, wq_stream_t stream) {
    WQ_REQUIRE(P >= 0 && P < (1ll << 31), "wq_edit_distance: bad pair count");
    if (P == 0) return WQ_OK;
    WQ_REQUIRE(ref_off && hyp_off && dist, "wq_edit_distance: null pointer");
    k_edit_distance<<<(unsigned)P, ED_THREADS, 0, (cudaStream_t)stream>>>(ref, ref_off, hyp, hyp_off, dist);
    WQ_LAUNCH_CHECK();
    return WQ_OK;
}
