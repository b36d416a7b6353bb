// kernels_prealign.cuh — pre-alignment contig selection on the GPU (definition: prealign_core.h; replaces the per-read,
// per-target banded alignments of Aligners::align, fg-stitch-lib/src/align/aligners/mod.rs:246-295, 556-604).
//
// One persistent CTA per read in flight.  The k-mer index of all contig-strands lives in HBM (bucket offsets + positions,
// built once per context); a read's k-mers are looked up twice:
//   pass 1  every hit adds one to the counter of its contig-strand (per-CTA counters in global memory, L2-resident);
//           the strands with enough hits become candidates (at most PRE_MAX_CAND);
//   pass 2  every hit on a candidate adds one to the bin of its diagonal (width = band width);
//   score   one warp per candidate: the best pair of adjacent bins -> chain hits -> score; the strands reaching the
//           minimum score are written out in ascending order (at most MAX_STRANDS, the best ones).
// Memory traffic per read: n bucket look-ups (8 B) + hits x (4 B position + a 12-step binary search in the strand
// table, cached) twice: a few MB against the GBs of the alignment that follows; the kernel is latency-bound gather work.
#pragma once
#include <cuda_runtime.h>

#include "prealign_core.h"

namespace stitch {
namespace gpu {

struct PreKParams {
    const uint32_t *off, *pos;            // k-mer index
    const uint8_t *blob;                  // contig bases (k > 12: hits are verified)
    const uint32_t *seq_off, *strand_len; // per strand
    uint32_t n_strands, K, W, n_bins;     // n_bins per candidate (covers diagonals of the longest read x longest strand)
    int32_t match, min_score;
    const uint8_t *reads; const uint64_t *read_off; const uint32_t *read_len; uint32_t n_reads;
    uint32_t *counter;
    uint32_t *cnt;        // [grid][n_strands]   hits per strand, then candidate slot + 1
    uint32_t *bins;       // [grid][PRE_MAX_CAND][n_bins]
    PreHit *out;          // [n_reads][MAX_STRANDS]
    uint32_t *out_n;      // [n_reads]
};

constexpr int PRE_THREADS = 256;

// Calls fn(strand, diagonal + n) for every k-mer hit of the read.
template <typename F>
__device__ __forceinline__ void pre_for_each_hit(const PreKParams &P, const uint8_t *read, uint32_t n, F &&fn) {
    const uint32_t K = P.K;
    if (n < K) return;
    for (uint32_t j = threadIdx.x; j + K <= n; j += PRE_THREADS) {
        uint64_t code;
        if (!pre_kmer_code(read + j, K, code)) continue;
        const uint32_t b = pre_bucket(code, K);
        const uint32_t lo = __ldg(P.off + b), hi = __ldg(P.off + b + 1);
        for (uint32_t e = lo; e < hi; ++e) {
            const uint32_t p = __ldg(P.pos + e);
            if (K > PRE_DIRECT_K && !pre_same_kmer(P.blob + p, read + j, K)) continue;
            const uint32_t s = pre_strand_of(P.seq_off, P.n_strands, p);
            fn(s, (p - __ldg(P.seq_off + s)) + n - j);   // diagonal p_in_strand - j, shifted by n to stay non-negative
        }
    }
}

__global__ void __launch_bounds__(PRE_THREADS) prealign_kernel(const PreKParams P) {
    __shared__ uint32_t s_read, s_ncand, s_need;
    __shared__ uint32_t s_cand[PRE_MAX_CAND];
    __shared__ int32_t s_score[PRE_MAX_CAND];
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t *cnt = P.cnt + (size_t)blockIdx.x * P.n_strands;
    uint32_t *bins = P.bins + (size_t)blockIdx.x * PRE_MAX_CAND * P.n_bins;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_read = atomicAdd(P.counter, 1u);
        __syncthreads();
        const uint32_t r = s_read;
        if (r >= P.n_reads) break;
        const uint8_t *read = P.reads + P.read_off[r];
        const uint32_t n = P.read_len[r];
        // ---- pass 1: hits per strand ----
        for (uint32_t s = tid; s < P.n_strands; s += PRE_THREADS) cnt[s] = 0;
        if (tid == 0) s_need = pre_need(P.min_score, P.match, P.K);
        __syncthreads();
        pre_for_each_hit(P, read, n, [&](uint32_t s, uint32_t) { atomicAdd(cnt + s, 1u); });
        __syncthreads();
        // ---- candidates (the threshold doubles until they fit) ----
        for (;;) {
            if (tid == 0) s_ncand = 0;
            __syncthreads();
            const uint32_t need = s_need;
            for (uint32_t s = tid; s < P.n_strands; s += PRE_THREADS)
                if (cnt[s] >= need) { const uint32_t k = atomicAdd(&s_ncand, 1u); if (k < PRE_MAX_CAND) s_cand[k] = s; }
            __syncthreads();
            if (s_ncand <= PRE_MAX_CAND) break;
            __syncthreads();
            if (tid == 0) s_need = need * 2;
            __syncthreads();
        }
        const uint32_t ncand = s_ncand;
        if (ncand == 0) { if (tid == 0) P.out_n[r] = 0; continue; }
        // slot + 1 of every candidate replaces its counter (0 = not a candidate); bins cleared
        for (uint32_t s = tid; s < P.n_strands; s += PRE_THREADS) cnt[s] = 0;
        __syncthreads();
        for (uint32_t k = tid; k < ncand; k += PRE_THREADS) cnt[s_cand[k]] = k + 1;
        for (uint32_t x = tid; x < ncand * P.n_bins; x += PRE_THREADS) bins[x] = 0;
        __syncthreads();
        // ---- pass 2: hits per diagonal bin of the candidates ----
        pre_for_each_hit(P, read, n, [&](uint32_t s, uint32_t dshift) {
            const uint32_t slot = cnt[s];
            if (slot) {
                const uint32_t b = dshift / P.W;
                if (b < P.n_bins) atomicAdd(bins + (size_t)(slot - 1) * P.n_bins + b, 1u);
            }
        });
        __syncthreads();
        // ---- score: one warp per candidate ----
        for (uint32_t k = warp; k < ncand; k += PRE_THREADS / 32) {
            const uint32_t *bk = bins + (size_t)k * P.n_bins;
            uint32_t best = 0;
            for (uint32_t b = lane; b < P.n_bins; b += 32) {
                const uint32_t v = bk[b] + (b + 1 < P.n_bins ? bk[b + 1] : 0u);
                best = v > best ? v : best;
            }
            for (int d = 16; d >= 1; d >>= 1) { const uint32_t o = __shfl_xor_sync(0xffffffffu, best, d); best = o > best ? o : best; }
            if (lane == 0) s_score[k] = pre_score_of(P.match, best, P.K);
        }
        __syncthreads();
        // ---- selection, in ascending strand order (thread 0: a few candidates per read) ----
        if (tid == 0) {
            uint32_t nk = 0;
            for (uint32_t k = 0; k < ncand; ++k)
                if (s_score[k] >= P.min_score) { s_cand[nk] = s_cand[k]; s_score[nk] = s_score[k]; ++nk; }
            auto sort = [&](uint32_t cntk, bool by_score) {   // insertion sort of (s_cand, s_score): (score desc, strand asc) or strand asc
                for (uint32_t i = 1; i < cntk; ++i) {
                    const uint32_t c = s_cand[i]; const int32_t sc = s_score[i];
                    uint32_t j = i;
                    while (j > 0) {
                        const bool prev_first = (by_score && s_score[j - 1] != sc) ? s_score[j - 1] > sc : s_cand[j - 1] < c;
                        if (prev_first) break;
                        s_cand[j] = s_cand[j - 1]; s_score[j] = s_score[j - 1]; --j;
                    }
                    s_cand[j] = c; s_score[j] = sc;
                }
            };
            if (nk > MAX_STRANDS) { sort(nk, true); nk = MAX_STRANDS; }
            sort(nk, false);
            for (uint32_t k = 0; k < nk; ++k) { PreHit h; h.strand = s_cand[k]; h.score = s_score[k]; P.out[(size_t)r * MAX_STRANDS + k] = h; }
            P.out_n[r] = nk;
        }
    }
}

}  // namespace gpu
}  // namespace stitch
