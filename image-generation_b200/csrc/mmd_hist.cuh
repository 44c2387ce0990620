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

// The same count for fp32 accumulators (mmd_gram_fp4_kernel) of a pure tile: (D - g) / 2 is formed in fp32 (exact: an
// integer <= D), clamped there, and read out of the mantissa of  x + 2^23  -- no float-to-int conversion (F2I runs on the
// quarter-rate conversion pipe), no shifts.
__device__ __forceinline__ void hist_count_chunk_f32(const uint32_t (&v)[32], HistAccumulator &acc, float half_d, float d_f)
{
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        const float h = fminf(fmaxf(__fmaf_rn(u2f(v[c]), -0.5f, half_d), 0.0f), d_f);
        atomicAdd(acc.bins + (f2u(__fadd_rn(h, 8388608.0f)) & 0x7fffffu), 1u);
    }
}

// How a tile's entries are counted:
//   TILE_PURE  : the whole tile lies strictly above the diagonal inside one block -> shared-memory counters
//   TILE_DIAG  : a tile ON the diagonal of the x x x or y x y block, fully inside the matrix: entries right of the
//                diagonal -> the same shared-memory counters, the 128 diagonal entries -> global histogram (weight 1),
//                entries left of the diagonal are the mirror images and are skipped
//   TILE_MIXED : block boundary / matrix edge inside the tile -> guarded, weighted, straight to the global histogram
// (Every diagonal tile used to take the TILE_MIXED path: 32 768 global 64-bit atomics on a few hundred addresses.  With
// one or two such tiles per CTA and none on others that was the tail of the whole launch.)
enum { TILE_MIXED = 0, TILE_PURE = 1, TILE_DIAG = 2 };

__device__ __forceinline__ int tile_mode(bool strict_upper, bool in_bounds, bool rows_x, bool rows_y, bool cols_x, bool cols_y)
{
    if (!in_bounds || !(rows_x || rows_y) || !(cols_x || cols_y)) return TILE_MIXED;
    if (strict_upper) return TILE_PURE;
    return (rows_x && cols_x) || (rows_y && cols_y) ? TILE_DIAG : TILE_MIXED;
}

// 32 Gram entries of one accumulator row (one tcgen05.ld chunk) into the histograms.
__device__ __forceinline__ void hist_count_chunk(const uint32_t (&v)[32], int mode, HistAccumulator &acc, int two_d,
                                                 int row, int col_first, int m_x, int m)
{
    if (mode == TILE_PURE) {
#pragma unroll
        for (int c = 0; c < 32; ++c) atomicAdd(acc.bins + hamming_index(two_d, (int)v[c], acc.d), 1u);
    } else if (mode == TILE_DIAG) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const int col = col_first + c;
            const int h = hamming_index(two_d, (int)v[c], acc.d);
            if (col > row) atomicAdd(acc.bins + h, 1u);
            else if (col == row) atomicAdd(acc.global + (size_t)acc.type * (size_t)(acc.d + 1) + h, 1ull);
        }
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
