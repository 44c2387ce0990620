// Hamming-histogram epilogue shared by the int8 MMD Gram kernels (mmd_tc.cu, mmd_tc2.cu).
//
// For +-1 rows every kernel value is a function of the integer Hamming distance h = (D - a.b) / 2,
// h in [0, D].  The Gram kernels therefore do not evaluate the RBF mixture at all: their epilogue
// COUNTS Gram entries per distance into three (D + 1)-bin histograms -- pairs inside x, pairs inside y,
// pairs across -- with integer atomics, and a tiny kernel (mmd_eval_hist_kernel) turns the histograms
// into  sum_ab t_ab  (the data-dependent bandwidth), and the three block sums  sum k(t_ab)  in float64.
// Consequences:
//   * the auto-bandwidth MMD needs ONE Gram pass instead of two ("MMD bandwidth
//     dependency");
//   * the result is exact up to the float64 evaluation of (D + 1) kernel values -- no fp32 partial sums;
//   * the histograms are integers: ranks that each contract a share of the tiles combine them with an
//     int64 all-reduce and obtain bit-identical block sums at any GPU count.
//
// Counting convention: hist[0] / hist[1] hold ORDERED pairs (a, b) of x x x / y x y, diagonal included
// (the kernels visit the upper triangle and weigh off-diagonal entries by 2); hist[2] holds x x y pairs.
#pragma once

#include "tc_common.cuh"

namespace b200grbm {

enum { HIST_XX = 0, HIST_YY = 1, HIST_XY = 2 };

// named barrier over the epilogue warps only (the producer / MMA warps never join it)
__device__ __forceinline__ void epi_bar_sync(int n_threads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

// Hamming distance of a Gram entry, clamped so that rows that are not +-1 (caller error) can never
// index outside the table
__device__ __forceinline__ int hamming_index(int two_d, int gram, int d)
{
    return min(max((two_d - 2 * gram) >> 2, 0), d);
}

// Per-CTA histogram in shared memory for the tiles whose entries all belong to one block (the vast
// majority); flushed to the global int64 histogram when the block type changes and at the end.
struct HistAccumulator {
    uint32_t *bins;                 // shared memory, d + 1 counters
    unsigned long long *global;     // [3][d + 1]
    int d, type;                    // type = -1: empty

    __device__ __forceinline__ void flush(int epi_tid, int epi_threads)
    {
        if (type < 0) return;
        epi_bar_sync(epi_threads);                                  // every warp's counts have landed
        const unsigned long long w = type == HIST_XY ? 1ull : 2ull; // strictly-upper tiles of xx / yy count twice
        unsigned long long *g = global + (size_t)type * (size_t)(d + 1);
        for (int k = epi_tid; k <= d; k += epi_threads) {
            const uint32_t v = bins[k];
            if (v != 0u) {
                atomicAdd(g + k, w * (unsigned long long)v);
                bins[k] = 0u;
            }
        }
        epi_bar_sync(epi_threads);                                  // zeroed before anyone counts again
        type = -1;
    }
};

// 32 Gram entries of one accumulator row (one tcgen05.ld chunk) into the histograms.
//   pure  : the whole tile lies strictly above the diagonal inside one block -> shared-memory counters
//   !pure : diagonal / boundary / edge tile -> guarded, weighted, straight to the global histogram
__device__ __forceinline__ void hist_count_chunk(const uint32_t (&v)[32], bool pure, HistAccumulator &acc, int two_d,
                                                 int row, int col_first, int m_x, int m)
{
    if (pure) {
#pragma unroll
        for (int c = 0; c < 32; ++c) atomicAdd(acc.bins + hamming_index(two_d, (int)v[c], acc.d), 1u);
    } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const int col = col_first + c;
            if (row < m && col < m && col >= row) {
                const bool rx = row < m_x, cx = col < m_x;
                const int type = rx && cx ? HIST_XX : (!rx && !cx ? HIST_YY : HIST_XY);
                const unsigned long long w = (type == HIST_XY || col == row) ? 1ull : 2ull;
                atomicAdd(acc.global + (size_t)type * (size_t)(acc.d + 1) + hamming_index(two_d, (int)v[c], acc.d), w);
            }
        }
    }
}

}  // namespace b200grbm
