// Mixture-of-RBF MMD for +-1 rows: Gram contraction on tcgen05 tensor cores (sm_100a).
//
// Replaces the stock  cat -> cdist -> exp x 7 -> block means  behind
// maximum_mean_discrepancy_loss(x, y, GaussianKernel(7))  (third-party dwave-pytorch-plugin,
// call site src/model_wrapper.py:320; SURVEY.md section 8 rows a7/a8, config cfg3:
// 8192 encoder latents vs 8192 GRBM samples, D = 5640).
//
// For +-1 rows  ||a - b||^2 = 4 * Hamming(a, b) = 2 (D - a.b),  so every kernel value is a
// function of the integer Gram entry.  The kernel therefore
//   * stages 128 x 128-byte boxes of the int8 sample matrix Z = [x; y] with TMA
//     (cp.async.bulk.tensor, 128B swizzle) through a 4-stage mbarrier ring,
//   * contracts them with tcgen05.mma.kind::i8 (int8 x int8 -> int32, exact) into a
//     128 x 256 accumulator tile in TMEM (two accumulator stages = all 512 columns, so the
//     epilogue of tile t overlaps the MMAs of tile t+1),
//   * forward (TC_PASS_HIST): the epilogue counts the Gram entries of the tile per Hamming distance
//     (mmd_hist.cuh: shared-memory integer atomics, three (D+1)-bin histograms xx / yy / xy); ONE pass
//     yields the distance sum for the data-dependent bandwidth AND the three block sums, evaluated
//     afterwards in float64 by mmd_eval_hist_kernel.  Nothing of size m^2 touches HBM;
//   * backward (TC_PASS_COEF): the epilogue maps each entry through a (D+1)-entry look-up table of
//     dk/dt-coefficients and writes them as 2 or 3 int8 fixed-point digit planes for the int8 GEMM
//     of gemm_i8.cu, together with their exact integer row sums.
// Only the upper triangle of 128 x 256 tiles is visited in the forward pass (the kernel matrix is
// symmetric); `shard_rank / shard_world` deal the tiles of the enumeration round-robin over ranks.
//
// Warp roles (320 threads, one persistent CTA per SM): warp 0 lane 0 = TMA producer,
// warp 1 = TMEM allocator + (lane 0) MMA issuer, warps 2..9 = epilogue (two column halves x
// four TMEM lane quarters).
#include "mmd_hist.cuh"

#include <stdlib.h>

#include <algorithm>

namespace b200grbm {

constexpr int BM = 128;             // accumulator rows   (UMMA M)
constexpr int BN = 256;             // accumulator columns (UMMA N)
constexpr int BK = 128;             // bytes (= int8 elements) per k-block: one 128B swizzle row
constexpr int UMMA_K = 32;          // int8 elements per tcgen05.mma
constexpr int A_BYTES = BM * BK;    // 16 KB
constexpr int B_BYTES = BN * BK;    // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int TC_THREADS = 320;
constexpr int EPI_WARPS = 8;
constexpr uint32_t TMEM_COLS = 512;

enum { TC_PASS_HIST = 0, TC_PASS_COEF = 2 };
constexpr int TC_MAX_SB = 136;          // up to 16 super-blocks per side (m <= 65536)
constexpr int TC_SB_ROWS = 32, TC_SB_COLS = 16;   // tiles per super-block: 32 x 128 rows, 16 x 256 columns

struct TcParams {
    int m_x, m, d;
    int n_kblocks;
    int tiles_m, tiles_n, j0, p0, total_tiles;
    int stages;
    int pass;
    int shard_rank, shard_world;        // forward: this launch contracts tiles t = shard_rank (mod shard_world)
    unsigned long long *hist;           // TC_PASS_HIST: [3][d + 1] ordered-pair counts per Hamming distance
    // TC_PASS_COEF: backward coefficients A[a][b] = w * c(h_ab) as n_planes base-256 int8 digits (most significant
    // plane first), rows row0 .. row0 + n_rows - 1 of the stacked matrix against every column
    const float *lut;                   // [d + 1]  c(h) / max|c| * Q
    int row0, n_rows, tiles_rows;       // row tiles covering the requested rows
    int m_pad, rows_alloc, n_planes;    // plane p = planes + p * rows_alloc * m_pad, row pitch m_pad
    float r_xx, r_xy;                   // w_xx / max(|w_xx|, |w_xy|), w_xy / max(...)
    float q_max;                        // largest digit-plane magnitude (clamp for distances the scale did not see)
    int8_t *planes;
    long long *rowsum;                  // [n_rows] integer row sums of the quantised coefficients (accumulating)
    // L2-sized super-blocks of the tile triangle (forward pass): 4096 x 4096 entries = 32 x 16 tiles per
    // block pair, so the rows a wave of CTAs touches (2 x 23 MB at D = 5640) stay resident in one L2 partition
    int sb_count;                       // 0 -> plain column-major triangle
    int sb_start[TC_MAX_SB + 1];
    unsigned char sb_bi[TC_MAX_SB], sb_bj[TC_MAX_SB];
};

// upper-triangle tile enumeration: column tile j ascending, row tiles i = 0 .. min(tiles_m, 2j+2) - 1
__device__ __forceinline__ void tile_coords(const TcParams &p, int t, int &i, int &j)
{
    if (p.pass == TC_PASS_COEF) {          // full rectangle: the requested rows against every column
        i = t % p.tiles_rows;
        j = t / p.tiles_rows;
        return;
    }
    int tm = p.tiles_m, j0 = p.j0, p0 = p.p0, ibase = 0, jbase = 0;
    if (p.sb_count > 0) {
        int k = 0;
        while (k + 1 < p.sb_count && p.sb_start[k + 1] <= t) ++k;
        t -= p.sb_start[k];
        const int bi = p.sb_bi[k], bj = p.sb_bj[k];
        ibase = bi * TC_SB_ROWS;
        jbase = bj * TC_SB_COLS;
        const int nr = min(TC_SB_ROWS, p.tiles_m - ibase), nc = min(TC_SB_COLS, p.tiles_n - jbase);
        if (bi != bj) {                     // whole block lies right of the diagonal
            i = ibase + t % nr;
            j = jbase + t / nr;
            return;
        }
        tm = nr;                            // diagonal block: the same triangle rule in local coordinates
        j0 = nr / 2 < nc ? nr / 2 : nc;
        p0 = j0 * (j0 + 1);
    }
    if (t < p0) {
        j = (int)((sqrtf(4.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((j + 1) * (j + 2) <= t) ++j;
        while (j * (j + 1) > t) --j;
        i = t - j * (j + 1);
    } else {
        const int r = t - p0;
        j = j0 + r / tm;
        i = r % tm;
    }
    i += ibase;
    j += jbase;
}

// tiles of this launch: local index u -> global enumeration index
__device__ __forceinline__ int local_tiles(const TcParams &p)
{
    return p.total_tiles > p.shard_rank ? (p.total_tiles - p.shard_rank + p.shard_world - 1) / p.shard_world : 0;
}

__global__ void __launch_bounds__(TC_THREADS, 1) mmd_gram_i8_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                    const TcParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 128B-swizzled TMA boxes / UMMA descriptors need 1024-byte aligned stages: align by hand
    // (the launch reserves 1024 spare bytes)
    unsigned char *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    unsigned char *stage_base = smem;                                        // S x 48 KB, 1024-aligned
    // d + 1 words: histogram counters (forward) or the coefficient table (backward)
    uint32_t *table = reinterpret_cast<uint32_t *>(smem + (size_t)S * STAGE_BYTES);
    const size_t table_bytes = ((size_t)(p.d + 1) * 4 + 15) / 16 * 16;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * STAGE_BYTES + table_bytes);
    // bars: full[S], empty[S], tmem_full[2], tmem_empty[2]; then the TMEM base address
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 4);
    const uint32_t full0 = smem_addr(bars), empty0 = full0 + 8u * S, tfull0 = empty0 + 8u * S, tempty0 = tfull0 + 16u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { bar_init(full0 + 8u * s, 1); bar_init(empty0 + 8u * s, 1); }
        for (int a = 0; a < 2; ++a) { bar_init(tfull0 + 8u * a, 1); bar_init(tempty0 + 8u * a, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_addr(tmem_slot), TMEM_COLS);
    if (p.pass == TC_PASS_COEF) {
        for (int k = threadIdx.x; k <= p.d; k += blockDim.x) table[k] = f2u(p.lut[k]);
    } else {
        for (int k = threadIdx.x; k <= p.d; k += blockDim.x) table[k] = 0u;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int n_local = local_tiles(p);
    const int row_origin = p.pass == TC_PASS_COEF ? p.row0 : 0;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int u = blockIdx.x; u < n_local; u += gridDim.x) {
                int ti, tj;
                tile_coords(p, u * p.shard_world + p.shard_rank, ti, tj);
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    bar_wait(empty0 + 8u * s, ph ^ 1u);
                    const uint32_t fb = full0 + 8u * s;
                    const uint32_t a_dst = smem_addr(stage_base + (size_t)s * STAGE_BYTES);
                    bar_expect_tx(fb, STAGE_BYTES);
                    tma_load_2d(a_dst, &tmap, kb * BK, row_origin + ti * BM, fb);
                    tma_load_2d(a_dst + A_BYTES, &tmap, kb * BK, tj * BN, fb);
                    tma_load_2d(a_dst + A_BYTES + A_BYTES, &tmap, kb * BK, tj * BN + 128, fb);
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_i8(BM, BN);
            int s = 0;
            uint32_t ph = 0;
            int it = 0;
            for (int u = blockIdx.x; u < n_local; u += gridDim.x, ++it) {
                const uint32_t acc = (uint32_t)it & 1u, acc_ph = ((uint32_t)it >> 1) & 1u;
                bar_wait(tempty0 + 8u * acc, acc_ph ^ 1u);          // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    bar_wait(full0 + 8u * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_addr(stage_base + (size_t)s * STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(a_addr);
                    const uint64_t bdesc = umma_desc_sw128(a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)       // +32 bytes inside the swizzle row = +2 in the address field
                        umma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty0 + 8u * s);                    // frees the smem stage when the MMAs retire
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
                umma_commit(tfull0 + 8u * acc);                      // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue =====================
        const int ew = warp - 2;                 // 0..7
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int half = ew >> 2;                // column half of the accumulator
        const int epi_tid = threadIdx.x - 64;
        const int two_d = 2 * p.d;
        HistAccumulator hacc = {table, p.hist, p.d, -1};
        const float *lut = reinterpret_cast<const float *>(table);
        int it = 0;
        for (int u = blockIdx.x; u < n_local; u += gridDim.x, ++it) {
            int ti, tj;
            tile_coords(p, u * p.shard_world + p.shard_rank, ti, tj);
            const uint32_t acc = (uint32_t)it & 1u, acc_ph = ((uint32_t)it >> 1) & 1u;
            const int row0 = row_origin + ti * BM, col0 = tj * BN;
            const int row = row0 + quarter * 32 + lane;
            int mode = TILE_MIXED;
            if (p.pass == TC_PASS_HIST) {
                const bool strict_upper = ti < 2 * tj;          // every column of the tile is right of every row
                const bool rows_x = row0 + BM <= p.m_x, rows_y = row0 >= p.m_x;
                const bool cols_x = col0 + BN <= p.m_x, cols_y = col0 >= p.m_x;
                mode = tile_mode(strict_upper, (row0 + BM <= p.m) && (col0 + BN <= p.m), rows_x, rows_y, cols_x, cols_y);
                if (mode != TILE_MIXED) {
                    const int type = rows_x && cols_x ? HIST_XX : (rows_y && cols_y ? HIST_YY : HIST_XY);
                    if (type != hacc.type) {         // uniform over the epilogue threads: same tile sequence
                        hacc.flush(epi_tid, EPI_WARPS * 32);
                        hacc.type = type;
                    }
                }
            }
            bar_wait(tfull0 + 8u * acc, acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            long long row_q = 0;
#pragma unroll 1
            for (int chunk = 0; chunk < 4; ++chunk) {
                uint32_t v[32];
                const int cbase = half * 128 + chunk * 32;
                __syncwarp();                                   // tcgen05.ld is .sync.aligned
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + (uint32_t)cbase, v);
                if (p.pass == TC_PASS_HIST) {
                    hist_count_chunk(v, mode, hacc, two_d, row, col0 + cbase, p.m_x, p.m);
                } else if (row - p.row0 < p.n_rows) {
                    // 32 consecutive coefficients of this thread's row -> base-256 digits, 32 bytes per plane
                    uint32_t dig[3][8];
                    int part = 0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int col = col0 + cbase + c;
                        int q = 0;
                        if (col < p.m && col != row)
                            q = __float2int_rn(fminf(fmaxf(lut[hamming_index(two_d, (int)v[c], p.d)], -p.q_max), p.q_max) *
                                               (col < p.m_x ? p.r_xx : p.r_xy));
                        part += q;
                        // signed base-256 digits, least significant first: q = d0 + 256 d1 + 65536 d2
                        const int d0 = (q << 24) >> 24, q1 = (q - d0) >> 8;
                        const int d1 = (q1 << 24) >> 24, d2 = (q1 - d1) >> 8;
                        const int sh = 8 * (c & 3);
                        if ((c & 3) == 0) { dig[0][c >> 2] = 0u; dig[1][c >> 2] = 0u; dig[2][c >> 2] = 0u; }
                        dig[0][c >> 2] |= (uint32_t)(d0 & 0xff) << sh;
                        dig[1][c >> 2] |= (uint32_t)(d1 & 0xff) << sh;
                        dig[2][c >> 2] |= (uint32_t)(d2 & 0xff) << sh;
                    }
                    row_q += part;
                    if (col0 + cbase < p.m_pad) {              // m_pad is a multiple of 128: whole 32-byte groups
                        const size_t off = (size_t)(row - p.row0) * p.m_pad + (size_t)(col0 + cbase);
                        const size_t plane_stride = (size_t)p.rows_alloc * p.m_pad;
                        const auto store_plane = [&](int pl, const uint32_t (&src)[8]) {
                            uint4 *dst = reinterpret_cast<uint4 *>(p.planes + (size_t)pl * plane_stride + off);
                            dst[0] = make_uint4(src[0], src[1], src[2], src[3]);
                            dst[1] = make_uint4(src[4], src[5], src[6], src[7]);
                        };
                        if (p.n_planes == 3) {                     // plane 0 = most significant digit
                            store_plane(0, dig[2]);
                            store_plane(1, dig[1]);
                            store_plane(2, dig[0]);
                        } else {
                            store_plane(0, dig[1]);
                            store_plane(1, dig[0]);
                        }
                    }
                }
            }
            // accumulator drained: hand the TMEM stage back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(tempty0 + 8u * acc);
            if (p.pass == TC_PASS_COEF && row - p.row0 < p.n_rows && row_q != 0)
                atomicAdd(reinterpret_cast<unsigned long long *>(p.rowsum + (row - p.row0)), (unsigned long long)row_q);
        }
        if (p.pass == TC_PASS_HIST) hacc.flush(epi_tid, EPI_WARPS * 32);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// FP4 form of the forward pass.  +-1 and the zero padding are exact in e2m1, so the same integer Gram comes out of
// tcgen05.mma.kind::mxf4 on packed 4-bit operands: HALF the operand bytes per Gram entry (and twice the tensor rate) of
// int8.  The Gram kernels are bound by shared-memory bandwidth -- every k-block is written once by TMA and read once by
// the MMAs -- so the time follows the bytes.  Block scale factors are all 2^0: the TMEM columns behind the accumulators
// are filled with UE8M0 ones once, so whatever slot an instruction reads holds a one.  fp32 accumulation of +-1 products
// is exact (|sum| <= D < 2^24); the epilogue converts back to the integer Gram entry and counts as the int8 kernel does.
//
// Scale factors take TMEM columns, so two 256-column accumulators no longer fit: the shipped form has ONE 256-column
// accumulator, mainloop and epilogue take turns.  Measured at cfg3 (8192 + 8192 rows, D = 5640; times include the
// 0.04 ms packing pass and the histogram evaluation):
//   int8, CTA pair (mmd_tc2.cu) / single CTA                      0.47 / 0.45 ms
//   e2m1, one 256-column accumulator (this kernel)                0.34 ms
//   e2m1, two 128 x 128 sub-tiles through three 128-column accumulators (B200_F4_SPLIT=1: the epilogue of one sub-tile
//     under the MMAs of the next, but A is staged twice and B reuse halves)                                0.45 ms
//   per-lane conflict-free counters over a 256-distance window instead of ATOMS.POPC.INC on the CTA histogram, index
//     arithmetic in fp32 without F2I, TMEM load of the next chunk in flight while one is counted: no change each
//   the same pass as the Gram of the backward's coefficient pass: slower than int8 (1.53 vs 1.45 ms backward), because
//     that epilogue (245 MB of plane stores) is the longer half and here it is not overlapped -- not shipped
// What did matter (for the int8 kernels as much as for this one: 0.61 -> 0.47 ms) was the diagonal tiles, see
// mmd_hist.cuh TILE_DIAG.
#ifndef B200_F4_SPLIT
#define B200_F4_SPLIT 0
#endif
#if B200_F4_SPLIT                            // two 128-column sub-tiles through three accumulators (see above)
constexpr int F4_BN = 128, F4_ACCS = 3;
#else                                        // one 256-column accumulator, mainloop and epilogue take turns
constexpr int F4_BN = 256, F4_ACCS = 1;
#endif
constexpr int F4_NSUB = BN / F4_BN;
constexpr int F4_STAGE_BYTES = A_BYTES + F4_BN * BK;
constexpr uint32_t F4_SF_COL = F4_ACCS * F4_BN;          // scale-factor columns behind the accumulators

__global__ void __launch_bounds__(TC_THREADS, 1) mmd_gram_fp4_kernel(const __grid_constant__ CUtensorMap tmap, const TcParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    unsigned char *stage_base = smem;                                        // S x 32 KB, 1024-aligned
    uint32_t *table = reinterpret_cast<uint32_t *>(smem + (size_t)S * F4_STAGE_BYTES);       // d + 1 counters
    const size_t table_bytes = ((size_t)(p.d + 1) * 4 + 15) / 16 * 16;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * F4_STAGE_BYTES + table_bytes);
    // bars: full[S], empty[S], tmem_full[3], tmem_empty[3]; then the TMEM base address
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 2 * F4_ACCS);
    const uint32_t full0 = smem_addr(bars), empty0 = full0 + 8u * S, tfull0 = empty0 + 8u * S, tempty0 = tfull0 + 8u * F4_ACCS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { bar_init(full0 + 8u * s, 1); bar_init(empty0 + 8u * s, 1); }
        for (int a = 0; a < F4_ACCS; ++a) { bar_init(tfull0 + 8u * a, 1); bar_init(tempty0 + 8u * a, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_addr(tmem_slot), TMEM_COLS);
    for (int k = threadIdx.x; k <= p.d; k += blockDim.x) table[k] = 0u;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 2 && warp < 6) {                 // one warp per TMEM lane quarter: scale-factor columns <- UE8M0 1.0
        uint32_t ones[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) ones[c] = 0x7f7f7f7fu;
        for (uint32_t col = F4_SF_COL; col < TMEM_COLS; col += 32)
            tmem_st_32x32b_x32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + col, ones);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int n_local = local_tiles(p);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int u = blockIdx.x; u < n_local; u += gridDim.x) {
                int ti, tj;
                tile_coords(p, u * p.shard_world + p.shard_rank, ti, tj);
                for (int h = 0; h < F4_NSUB; ++h)
                    for (int kb = 0; kb < p.n_kblocks; ++kb) {
                        bar_wait(empty0 + 8u * s, ph ^ 1u);
                        const uint32_t fb = full0 + 8u * s;
                        const uint32_t a_dst = smem_addr(stage_base + (size_t)s * F4_STAGE_BYTES);
                        bar_expect_tx(fb, F4_STAGE_BYTES);
                        tma_load_2d(a_dst, &tmap, kb * BK, ti * BM, fb);
#pragma unroll
                        for (int b = 0; b < F4_BN / 128; ++b)
                            tma_load_2d(a_dst + A_BYTES + b * A_BYTES, &tmap, kb * BK, tj * BN + h * F4_BN + b * 128, fb);
                        if (++s == S) { s = 0; ph ^= 1u; }
                    }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_mxf4(BM, F4_BN);
            const uint32_t sfa = tmem_base + F4_SF_COL, sfb = tmem_base + F4_SF_COL + 64;
            int s = 0;
            uint32_t ph = 0;
            uint32_t acc = 0, acc_ph = 0;
            for (int u = blockIdx.x; u < n_local; u += gridDim.x) {
                for (int h = 0; h < F4_NSUB; ++h) {
                    bar_wait(tempty0 + 8u * acc, acc_ph ^ 1u);          // epilogue has drained this accumulator
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem_base + acc * F4_BN;
                    for (int kb = 0; kb < p.n_kblocks; ++kb) {
                        bar_wait(full0 + 8u * s, ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_addr = smem_addr(stage_base + (size_t)s * F4_STAGE_BYTES);
                        const uint64_t adesc = umma_desc_sw128(a_addr);
                        const uint64_t bdesc = umma_desc_sw128(a_addr + A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / 32; ++k)       // K = 64 e2m1 = 32 bytes = +2 in the address field
                            umma_mxf4(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, sfa, sfb,
                                      (kb | k) != 0 ? 1u : 0u);
                        umma_commit(empty0 + 8u * s);
                        if (++s == S) { s = 0; ph ^= 1u; }
                    }
                    umma_commit(tfull0 + 8u * acc);
                    if (++acc == F4_ACCS) { acc = 0; acc_ph ^= 1u; }
                }
            }
        }
    } else {
        // ===================== epilogue =====================
        const int ew = warp - 2;                 // 0..7
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int half = ew >> 2;                // column half of the sub-tile
        const int epi_tid = threadIdx.x - 64;
        const int two_d = 2 * p.d;
        HistAccumulator hacc = {table, p.hist, p.d, -1};
        uint32_t acc = 0, acc_ph = 0;
        for (int u = blockIdx.x; u < n_local; u += gridDim.x) {
            int ti, tj;
            tile_coords(p, u * p.shard_world + p.shard_rank, ti, tj);
            const int row0 = ti * BM, col0 = tj * BN;
            const int row = row0 + quarter * 32 + lane;
            const bool strict_upper = ti < 2 * tj;
            const bool rows_x = row0 + BM <= p.m_x, rows_y = row0 >= p.m_x;
            const bool cols_x = col0 + BN <= p.m_x, cols_y = col0 >= p.m_x;
            const int mode = tile_mode(strict_upper, (row0 + BM <= p.m) && (col0 + BN <= p.m), rows_x, rows_y, cols_x, cols_y);
            if (mode != TILE_MIXED) {
                const int type = rows_x && cols_x ? HIST_XX : (rows_y && cols_y ? HIST_YY : HIST_XY);
                if (type != hacc.type) {             // uniform over the epilogue threads: same tile sequence
                    hacc.flush(epi_tid, EPI_WARPS * 32);
                    hacc.type = type;
                }
            }
            for (int h = 0; h < F4_NSUB; ++h) {
                bar_wait(tfull0 + 8u * acc, acc_ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const float half_d = 0.5f * (float)p.d, d_f = (float)p.d;
#pragma unroll 1
                for (int chunk = 0; chunk < F4_BN / 64; ++chunk) {
                    uint32_t v[32];
                    const int cbase = half * (F4_BN / 2) + chunk * 32;
                    __syncwarp();                                   // tcgen05.ld is .sync.aligned
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * F4_BN + (uint32_t)cbase, v);
                    if (mode == TILE_PURE) {
                        hist_count_chunk_f32(v, hacc, half_d, d_f);
                    } else {
                        // fp32 -> int: the accumulator holds an integer |g| <= D < 2^22, so the low mantissa bits of
                        // g + 1.5 * 2^23 are g in two's complement
#pragma unroll
                        for (int c = 0; c < 32; ++c) v[c] = (uint32_t)(__float_as_int(__fadd_rn(u2f(v[c]), 12582912.0f)) - 0x4B400000);
                        hist_count_chunk(v, mode, hacc, two_d, row, col0 + h * F4_BN + cbase, p.m_x, p.m);
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) bar_arrive(tempty0 + 8u * acc);
                if (++acc == F4_ACCS) { acc = 0; acc_ph ^= 1u; }
            }
        }
        hacc.flush(epi_tid, EPI_WARPS * 32);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// int8 +-1 / 0 rows -> packed e2m1 (two per byte, element 2k in the low nibble): +1 = 0x2, -1 = 0xA, 0 = 0x0.
// One thread = 16 spins in (one 16-byte load), 8 bytes out.
__global__ void pack_fp4_i8_kernel(const int8_t *__restrict__ rows, int m, int d_pad, uint8_t *__restrict__ out, int row_bytes)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_row = row_bytes / 8;
    if (idx >= (size_t)m * per_row) return;
    const int r = (int)(idx / per_row), g = (int)(idx % per_row);
    uint32_t lo = 0u, hi = 0u;
    if (16 * g < d_pad) {                                   // d_pad is a multiple of 16: whole groups
        const uint4 v = *reinterpret_cast<const uint4 *>(rows + (size_t)r * d_pad + 16 * g);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t nib = 0u;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int8_t sp = (int8_t)(w[q] >> (8 * b));
                nib |= (sp > 0 ? 0x2u : (sp < 0 ? 0xAu : 0u)) << (4 * b);
            }
            if (q < 2) lo |= nib << (16 * q); else hi |= nib << (16 * (q - 2));
        }
    }
    *reinterpret_cast<uint2 *>(out + (size_t)r * row_bytes + 8 * g) = make_uint2(lo, hi);
}

// distance of two +-1 rows at Hamming distance h:  ||a - b|| = 2 sqrt(h)  (squared: 4 h)
__device__ __forceinline__ double hamming_to_t(int h, int squared) { return squared ? 4.0 * (double)h : 2.0 * sqrt((double)h); }

__device__ double block_reduce_sum(double v, double *scratch)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) t += scratch[w];    // same order in every thread
    return t;
}

// 1 / (bw * mul_factor^(u - n_kernels/2)) for the kernels of the mixture, computed once per block
__device__ void inverse_bandwidths(double bw, int n_kernels, float mul_factor, double *inv_b /* shared, [16] */)
{
    if ((int)threadIdx.x < n_kernels)
        inv_b[threadIdx.x] = 1.0 / (bw * pow((double)mul_factor, (double)((int)threadIdx.x - n_kernels / 2)));
    __syncthreads();
}

// Histograms -> sums[5] = { sum_{a,b in x} k, sum_{a,b in y} k, sum_{a in x, b in y} k, sum_ab t_ab, MMD^2 estimate }
// in float64.  One block; the reductions run in a fixed order, so the result is bit-reproducible.  The estimate
// (sums[4]) is  scale * (xx + yy - 2 xy)  with the block means of the unbiased (diagonal dropped) or biased form.
__global__ void __launch_bounds__(1024) mmd_eval_hist_kernel(const unsigned long long *__restrict__ hist, int d, int m_x, int m_y,
                                                             int n_kernels, float mul_factor, int squared, float bandwidth,
                                                             int unbiased, double scale, double *__restrict__ sums)
{
    __shared__ double scratch[32];
    __shared__ double inv_b[16];
    const size_t stride = (size_t)d + 1;
    double dist = 0.0;
    for (int h = threadIdx.x; h <= d; h += blockDim.x)
        dist += hamming_to_t(h, squared) * ((double)hist[h] + (double)hist[stride + h] + 2.0 * (double)hist[2 * stride + h]);
    dist = block_reduce_sum(dist, scratch);
    const double mm = (double)(m_x + m_y);
    const double bw = bandwidth > 0.f ? (double)bandwidth : dist / (mm * mm - mm);
    inverse_bandwidths(bw, n_kernels, mul_factor, inv_b);
    double s[3] = {0.0, 0.0, 0.0};
    for (int h = threadIdx.x; h <= d; h += blockDim.x) {
        const unsigned long long c0 = hist[h], c1 = hist[stride + h], c2 = hist[2 * stride + h];
        if ((c0 | c1 | c2) == 0ull) continue;
        const double t = hamming_to_t(h, squared);
        double k = 0.0;
        for (int u = 0; u < n_kernels; ++u) k += exp(-t * inv_b[u]);
        s[0] += k * (double)c0;
        s[1] += k * (double)c1;
        s[2] += k * (double)c2;
    }
    double tot[3];
    for (int b = 0; b < 3; ++b) tot[b] = block_reduce_sum(s[b], scratch);
    if (threadIdx.x == 0) {
        sums[0] = tot[0]; sums[1] = tot[1]; sums[2] = tot[2]; sums[3] = dist;
        const double nx = (double)m_x, ny = (double)m_y, diag = (double)n_kernels;       // k(a, a) = n_kernels * exp(0)
        const double xx = unbiased ? (tot[0] - diag * nx) / (nx * (nx - 1.0)) : tot[0] / (nx * nx);
        const double yy = unbiased ? (tot[1] - diag * ny) / (ny * (ny - 1.0)) : tot[1] / (ny * ny);
        sums[4] = scale * (xx + yy - 2.0 * (tot[2] / (nx * ny)));
    }
}

// Backward coefficient table over the Hamming distance:  c(h) = (dk/dt)(dt/d||.||)/||.||  (multiplies x_a - z_b;
// zero where the rows coincide), normalised to  lut[h] = c(h) / max|c| * Q  with Q the largest magnitude the
// n_planes base-256 digits represent.  scale_out[0] = max|c| * max(|w_xx|, |w_xy|) / Q  undoes it after the GEMM.
// One block; c(h) is evaluated once (float64) and parked in the output table's own storage as its float32 value
// until the maximum is known.
__global__ void __launch_bounds__(1024) mmd_coef_lut_kernel(int d, int m, int n_kernels, float mul_factor, int squared,
                                                            float bandwidth, const double *__restrict__ sums, float w_xx,
                                                            float w_xy, int n_planes, const unsigned long long *__restrict__ hist,
                                                            float *__restrict__ lut, double *__restrict__ scale_out)
{
    __shared__ double scratch[32];
    __shared__ double inv_b[16];
    const double mm = (double)m;
    const double bw = bandwidth > 0.f ? (double)bandwidth : sums[3] / (mm * mm - mm);
    inverse_bandwidths(bw, n_kernels, mul_factor, inv_b);
    constexpr int PER_THREAD = 8;                   // d + 1 <= 8192 bins are kept in registers, more are recomputed
    double cval[PER_THREAD];
    const auto coef_at = [&](int h) {
        if (h == 0) return 0.0;
        const double t = hamming_to_t(h, squared);
        double dk = 0.0;
        for (int u = 0; u < n_kernels; ++u) dk -= exp(-t * inv_b[u]) * inv_b[u];
        return squared ? 2.0 * dk : dk / t;
    };
    double cmax = 0.0;
    int j = 0;
    for (int h = threadIdx.x; h <= d; h += blockDim.x, ++j) {
        const double c = coef_at(h);
        if (j < PER_THREAD) cval[j] = c;
        // with the forward's histograms at hand the fixed-point range covers only distances that occur among the
        // x-x and x-y pairs (|c| grows steeply towards h = 1, which real data rarely reaches: ~7 bits regained)
        if (hist == nullptr || (hist[h] | hist[2 * ((size_t)d + 1) + h]) != 0ull) cmax = fmax(cmax, fabs(c));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmax = fmax(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = cmax;
    __syncthreads();
    cmax = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) cmax = fmax(cmax, scratch[w]);
    const double Q = n_planes >= 3 ? 8355711.0 : 32639.0;        // 127 * (65536 + 256 + 1) / 127 * (256 + 1)
    const double norm = cmax > 0.0 ? Q / cmax : 0.0;
    j = 0;
    for (int h = threadIdx.x; h <= d; h += blockDim.x, ++j) lut[h] = (float)((j < PER_THREAD ? cval[j] : coef_at(h)) * norm);
    if (threadIdx.x == 0) scale_out[0] = cmax * fmax(fabs((double)w_xx), fabs((double)w_xy)) / Q;
}

// sign-pack fp32 rows into the zero-padded int8 matrix the TMA descriptor reads
__global__ void mmd_pack_i8_kernel(const float *__restrict__ z, int m, int d, int d_pad, int8_t *__restrict__ out)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)m * d_pad;
    if (idx >= total) return;
    const int r = (int)(idx / d_pad), c = (int)(idx % d_pad);
    out[idx] = c < d ? (z[(size_t)r * d + c] > 0.f ? (int8_t)1 : (int8_t)-1) : (int8_t)0;
}

int32_t get_tensor_map_encoder(encode_tiled_fn *out)
{
    static encode_tiled_fn cached = nullptr;
    if (cached == nullptr) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        B200_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (fn == nullptr || qres != cudaDriverEntryPointSuccess)
            return fail(B200GRBM_EUNSUPPORTED, "cuTensorMapEncodeTiled is not available from this driver");
        cached = reinterpret_cast<encode_tiled_fn>(fn);
    }
    *out = cached;
    return 0;
}

int32_t make_tensor_map_2d(CUtensorMap *map, const void *base, CUtensorMapDataType dtype, uint64_t inner_elems,
                           uint64_t rows, uint64_t row_pitch_bytes, uint32_t box_inner_elems, uint32_t box_rows)
{
    encode_tiled_fn encode = nullptr;
    B200_TRY(get_tensor_map_encoder(&encode));
    // cuTensorMapEncodeTiled is a driver-API call: make sure the runtime's primary context is
    // current on THIS thread (autograd runs backward passes on its own threads)
    B200_CUDA(cudaFree(nullptr));
    const cuuint64_t gdim[2] = {(cuuint64_t)inner_elems, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)row_pitch_bytes};
    const cuuint32_t box[2] = {(cuuint32_t)box_inner_elems, (cuuint32_t)box_rows};
    const cuuint32_t estride[2] = {1u, 1u};
    const CUresult cr = encode(map, dtype, 2, const_cast<void *>(base), gdim, gstride, box, estride,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(B200GRBM_EINVAL, "cuTensorMapEncodeTiled failed (%d)", (int)cr);
    return 0;
}

// CTA-pair version of the forward pass (mmd_tc2.cu)
int32_t launch_gram_i8_2cta(const CUtensorMap &tmap, int m_x, int m, int d, int d_pad, unsigned long long *hist,
                            int shard_rank, int shard_world, cudaStream_t st);

// Forward tile shape (int8 operands; from 2048 rows up the host layer runs the e2m1 kernel above instead).  The
// single-CTA kernel reads 48 KB of operands per k-block from shared memory and receives as many from TMA: against the
// 128 B/clk of shared-memory bandwidth that alone caps it near 70 % of the tensor rate (742 clk per k-block measured,
// 512 ideal).  The CTA-pair kernel (mmd_tc2.cu) stages 32 KB per k-block and CTA and is the default from 74 pair-tiles
// up; since the diagonal tiles are counted in shared memory the two are within 4 % of each other at cfg3 (0.47 pair /
// 0.45 single).  B200GRBM_MMD_TILE=1|2 forces one of them (A/B measurements, parity tests of both).
//
// Counting experiments (cfg3, B200), all exact.  They were read as "the counting takes shared-memory bandwidth from the
// MMA operand stream" until the e2m1 kernel, whose mainloop and epilogue do not overlap, showed the same cost with the
// index arithmetic alone: the time was the TAIL -- the 128 diagonal tiles, one or two per CTA on some CTAs and none on
// others, each doing 32 768 global 64-bit atomics (mmd_hist.cuh, TILE_DIAG; ncu's "epilogue warps wait for accumulators a
// third of the time" was the other CTAs idling behind it):
//   ATOMS per entry on a CTA histogram, diagonal tiles through global atomics single 0.69 ms   pair 0.60 ms
//   ... diagonal tiles through the CTA histogram as well (this version)       single 0.45 ms   pair 0.47 ms
//   per-lane byte counters in shared memory (conflict-free, no atomics),
//     windows of 224 / 288 distances, flushed per tile into the CTA histogram single 0.72 ms   pair 0.62 ms  (old tail)
//   the same windows flushed straight to the global histogram                 single 1.43 ms   (8 M atomics on ~300
//                                                                             hot addresses serialise in L2)
//   epilogue that reads TMEM and counts nothing, 3 stages                     single 0.50 ms  (old tail)
static bool use_pair_kernel(int m, int d_pad)
{
    const char *env = getenv("B200GRBM_MMD_TILE");
    if (env != nullptr && env[0] == '1') return false;
    if (env != nullptr && env[0] == '2') return true;
    (void)d_pad;
    const int tiles = (m + 255) / 256;
    return tiles * (tiles + 1) / 2 >= 74;
}

static int32_t gram_smem(int d, int *stages_out, size_t *smem_out, const char *who)
{
    int dev = 0, smem_optin = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t table_bytes = ((size_t)(d + 1) * 4 + 15) / 16 * 16;
    int stages = 4;
    size_t smem = 0;
    for (; stages >= 2; --stages) {
        smem = (size_t)stages * STAGE_BYTES + table_bytes + (2 * stages + 4) * 8 + 16;
        if (smem + 1024 <= (size_t)smem_optin) break;
    }
    if (stages < 2)
        return fail(B200GRBM_EUNSUPPORTED, "%s: d=%d needs a %zu B table, too large for shared memory", who, d, table_bytes);
    *stages_out = stages;
    *smem_out = smem + 1024;
    return 0;
}

}  // namespace b200grbm

using namespace b200grbm;

extern "C" int32_t b200grbm_mmd_pack_i8(const float *z_dev, int32_t m, int32_t d, int32_t d_pad, int8_t *out_dev,
                                        void *stream)
{
    if (m <= 0 || d <= 0 || d_pad < d || d_pad % 16 != 0)
        return fail(B200GRBM_EINVAL, "mmd_pack_i8: m=%d d=%d d_pad=%d (d_pad must be a multiple of 16 >= d)", m, d, d_pad);
    if (!z_dev || !out_dev) return fail(B200GRBM_EINVAL, "mmd_pack_i8: NULL pointer argument");
    B200_TRY(require_device());
    const size_t total = (size_t)m * d_pad;
    mmd_pack_i8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(z_dev, m, d, d_pad, out_dev);
    B200_CUDA(cudaGetLastError());
    return 0;
}

// fp4 = false: z_dev int8 [m][d_pad]; fp4 = true: z_dev packed e2m1 [m][d_pad bytes] (two spins per byte)
static int32_t mmd_hist_impl(const void *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t d_pad, int32_t shard_rank,
                             int32_t shard_world, uint64_t *hist_dev, void *stream, bool fp4)
{
    if (m_x <= 0 || m_y <= 0 || d <= 0 || d_pad < (fp4 ? (d + 1) / 2 : d) || d_pad % 16 != 0 || (fp4 && d_pad % 128 != 0))
        return fail(B200GRBM_EINVAL, "mmd_hist_%s: m_x=%d m_y=%d d=%d row bytes=%d", fp4 ? "fp4" : "i8", m_x, m_y, d, d_pad);
    if (shard_world < 1 || shard_rank < 0 || shard_rank >= shard_world)
        return fail(B200GRBM_EINVAL, "mmd_hist_i8: shard %d of %d", shard_rank, shard_world);
    if (!z_dev || !hist_dev) return fail(B200GRBM_EINVAL, "mmd_hist_i8: NULL pointer argument");
    if ((reinterpret_cast<uintptr_t>(z_dev) & 15u) != 0)
        return fail(B200GRBM_EINVAL, "mmd_hist_i8: z_dev must be 16-byte aligned (TMA)");
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const int m = m_x + m_y;
    int stages = 0;
    size_t smem = 0;
    B200_TRY(gram_smem(d, &stages, &smem, "mmd_hist_i8"));

    CUtensorMap tmap;
    B200_TRY(make_tensor_map_2d(&tmap, z_dev, CU_TENSOR_MAP_DATA_TYPE_UINT8, (uint64_t)d_pad, (uint64_t)m, (uint64_t)d_pad, BK, 128));
    if (!fp4 && use_pair_kernel(m, d_pad))
        return launch_gram_i8_2cta(tmap, m_x, m, d, d_pad, reinterpret_cast<unsigned long long *>(hist_dev), shard_rank,
                                   shard_world, st);

    TcParams p = {};
    p.m_x = m_x; p.m = m; p.d = d;
    p.n_kblocks = (d_pad + BK - 1) / BK;
    p.tiles_m = (m + BM - 1) / BM;
    p.tiles_n = (m + BN - 1) / BN;
    p.j0 = p.tiles_m / 2 < p.tiles_n ? p.tiles_m / 2 : p.tiles_n;
    p.p0 = p.j0 * (p.j0 + 1);
    p.total_tiles = p.p0 + (p.tiles_n - p.j0) * p.tiles_m;
    p.stages = stages;
    p.pass = TC_PASS_HIST;
    p.shard_rank = shard_rank; p.shard_world = shard_world;
    p.hist = reinterpret_cast<unsigned long long *>(hist_dev);
    {   // super-block order (column block outer, row block inner); same tile set, L2-friendly order
        const int nb = (p.tiles_n + TC_SB_COLS - 1) / TC_SB_COLS;
        const char *env = getenv("B200GRBM_MMD_ORDER");
        if (nb >= 2 && nb * (nb + 1) / 2 <= TC_MAX_SB && !(env != nullptr && env[0] == '0')) {
            int k = 0, start = 0;
            for (int bj = 0; bj < nb; ++bj)
                for (int bi = 0; bi <= bj; ++bi) {
                    const int nr = std::min(TC_SB_ROWS, p.tiles_m - bi * TC_SB_ROWS), nc = std::min(TC_SB_COLS, p.tiles_n - bj * TC_SB_COLS);
                    if (nr <= 0 || nc <= 0) continue;
                    int cnt = nr * nc;
                    if (bi == bj) {
                        const int j0 = nr / 2 < nc ? nr / 2 : nc;
                        cnt = j0 * (j0 + 1) + (nc - j0) * nr;
                    }
                    p.sb_bi[k] = (unsigned char)bi;
                    p.sb_bj[k] = (unsigned char)bj;
                    p.sb_start[k] = start;
                    start += cnt;
                    ++k;
                }
            p.sb_start[k] = start;
            p.sb_count = k;
            if (start != p.total_tiles)
                return fail(B200GRBM_EINVAL, "mmd_hist_i8: internal tile enumeration mismatch (%d vs %d)", start, p.total_tiles);
        }
    }
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int n_local = p.total_tiles > shard_rank ? (p.total_tiles - shard_rank + shard_world - 1) / shard_world : 0;
    if (n_local == 0) return 0;
    const int grid = n_local < sms ? n_local : sms;
    if (fp4) {
        // as many stages as fit, at most 6
        int dev = 0, smem_optin = 0;
        B200_CUDA(cudaGetDevice(&dev));
        B200_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        const size_t table_bytes = ((size_t)(d + 1) * 4 + 15) / 16 * 16;
        int s4 = 6;
        size_t smem4 = 0;
        for (; s4 >= 2; --s4) {
            smem4 = (size_t)s4 * F4_STAGE_BYTES + table_bytes + (2 * s4 + 2 * F4_ACCS) * 8 + 16 + 1024;
            if (smem4 <= (size_t)smem_optin) break;
        }
        if (s4 < 2) return fail(B200GRBM_EUNSUPPORTED, "mmd_hist_fp4: d=%d needs a %zu B table, too large for shared memory", d, table_bytes);
        p.stages = s4;
        B200_CUDA(cudaFuncSetAttribute(mmd_gram_fp4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
        mmd_gram_fp4_kernel<<<grid, TC_THREADS, smem4, st>>>(tmap, p);
    } else {
        B200_CUDA(cudaFuncSetAttribute(mmd_gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mmd_gram_i8_kernel<<<grid, TC_THREADS, smem, st>>>(tmap, p);
    }
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_mmd_hist_i8(const int8_t *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t d_pad,
                                        int32_t shard_rank, int32_t shard_world, uint64_t *hist_dev, void *stream)
{
    return mmd_hist_impl(z_dev, m_x, m_y, d, d_pad, shard_rank, shard_world, hist_dev, stream, false);
}

extern "C" int32_t b200grbm_mmd_hist_fp4(const uint8_t *z4_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t row_bytes,
                                         int32_t shard_rank, int32_t shard_world, uint64_t *hist_dev, void *stream)
{
    return mmd_hist_impl(z4_dev, m_x, m_y, d, row_bytes, shard_rank, shard_world, hist_dev, stream, true);
}

extern "C" int32_t b200grbm_pack_fp4_i8(const int8_t *rows_dev, int32_t m, int32_t d_pad, uint8_t *out_dev, int32_t row_bytes,
                                        void *stream)
{
    if (m <= 0 || d_pad <= 0 || d_pad % 16 != 0 || row_bytes % 128 != 0 || (long long)row_bytes * 2 < d_pad)
        return fail(B200GRBM_EINVAL, "pack_fp4_i8: m=%d d_pad=%d row_bytes=%d (row_bytes a multiple of 128 >= d_pad / 2)", m, d_pad,
                    row_bytes);
    if (!rows_dev || !out_dev || ((reinterpret_cast<uintptr_t>(rows_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 15u) != 0)
        return fail(B200GRBM_EINVAL, "pack_fp4_i8: NULL or unaligned pointer");
    B200_TRY(require_device());
    const size_t total = (size_t)m * (row_bytes / 8);
    pack_fp4_i8_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rows_dev, m, d_pad, out_dev, row_bytes);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_mmd_eval_hist(const uint64_t *hist_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t n_kernels,
                                          float mul_factor, int32_t squared, float bandwidth, int32_t unbiased, double scale,
                                          double *sums_dev, void *stream)
{
    if (m_x <= 0 || m_y <= 0 || d <= 0) return fail(B200GRBM_EINVAL, "mmd_eval_hist: m_x=%d m_y=%d d=%d", m_x, m_y, d);
    if (n_kernels < 1 || n_kernels > 16 || !(mul_factor > 0.f))
        return fail(B200GRBM_EINVAL, "mmd_eval_hist: n_kernels=%d mul_factor=%g", n_kernels, mul_factor);
    if (!hist_dev || !sums_dev) return fail(B200GRBM_EINVAL, "mmd_eval_hist: NULL pointer argument");
    B200_TRY(require_device());
    if (unbiased && (m_x < 2 || m_y < 2))
        return fail(B200GRBM_EINVAL, "mmd_eval_hist: the unbiased estimator needs at least two rows in x and in y");
    mmd_eval_hist_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(reinterpret_cast<const unsigned long long *>(hist_dev), d, m_x,
                                                              m_y, n_kernels, mul_factor, squared, bandwidth, unbiased, scale,
                                                              sums_dev);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_mmd_forward_i8(const int8_t *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t d_pad,
                                           int32_t n_kernels, float mul_factor, int32_t squared, float bandwidth,
                                           int32_t unbiased, double scale, uint64_t *hist_dev, double *sums_dev, void *stream)
{
    if (n_kernels < 1 || n_kernels > 16 || !(mul_factor > 0.f))
        return fail(B200GRBM_EINVAL, "mmd_forward_i8: n_kernels=%d mul_factor=%g", n_kernels, mul_factor);
    if (!hist_dev || !sums_dev) return fail(B200GRBM_EINVAL, "mmd_forward_i8: NULL pointer argument");
    if (m_x <= 0 || m_y <= 0 || d <= 0) return fail(B200GRBM_EINVAL, "mmd_forward_i8: m_x=%d m_y=%d d=%d", m_x, m_y, d);
    B200_TRY(require_device());
    B200_CUDA(cudaMemsetAsync(hist_dev, 0, 3 * ((size_t)d + 1) * sizeof(uint64_t), (cudaStream_t)stream));
    B200_TRY(b200grbm_mmd_hist_i8(z_dev, m_x, m_y, d, d_pad, 0, 1, hist_dev, stream));
    return b200grbm_mmd_eval_hist(hist_dev, m_x, m_y, d, n_kernels, mul_factor, squared, bandwidth, unbiased, scale, sums_dev,
                                  stream);
}

extern "C" int32_t b200grbm_mmd_coef_i8(const int8_t *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t d_pad,
                                        int32_t row0, int32_t n_rows, int32_t n_kernels, float mul_factor, int32_t squared,
                                        float bandwidth, const double *sums_dev, const uint64_t *hist_dev, float w_xx,
                                        float w_xy, float *lut_dev, int8_t *planes_dev, int32_t n_planes, int32_t rows_alloc,
                                        int32_t m_pad, int64_t *rowsum_dev, double *scale_dev, void *stream)
{
    if (m_x <= 0 || m_y <= 0 || d <= 0 || d_pad < d || d_pad % 16 != 0)
        return fail(B200GRBM_EINVAL, "mmd_coef_i8: m_x=%d m_y=%d d=%d d_pad=%d", m_x, m_y, d, d_pad);
    const int m = m_x + m_y;
    if (row0 < 0 || n_rows <= 0 || row0 + n_rows > m_x)
        return fail(B200GRBM_EINVAL, "mmd_coef_i8: rows [%d, %d) must lie inside the x block (m_x=%d)", row0, row0 + n_rows, m_x);
    if (m_pad < m || m_pad % 128 != 0) return fail(B200GRBM_EINVAL, "mmd_coef_i8: m_pad=%d must be a multiple of 128 >= m=%d", m_pad, m);
    if (n_planes < 2 || n_planes > 3) return fail(B200GRBM_EINVAL, "mmd_coef_i8: n_planes=%d must be 2 or 3", n_planes);
    if (rows_alloc < n_rows) return fail(B200GRBM_EINVAL, "mmd_coef_i8: rows_alloc=%d < n_rows=%d", rows_alloc, n_rows);
    if (n_kernels < 1 || n_kernels > 16 || !(mul_factor > 0.f))
        return fail(B200GRBM_EINVAL, "mmd_coef_i8: n_kernels=%d mul_factor=%g", n_kernels, mul_factor);
    if (!z_dev || !lut_dev || !sums_dev || !planes_dev || !rowsum_dev || !scale_dev)
        return fail(B200GRBM_EINVAL, "mmd_coef_i8: NULL pointer argument");
    if (((reinterpret_cast<uintptr_t>(z_dev) | reinterpret_cast<uintptr_t>(planes_dev)) & 15u) != 0)
        return fail(B200GRBM_EINVAL, "mmd_coef_i8: z_dev / planes_dev must be 16-byte aligned");
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    int stages = 0;
    size_t smem = 0;
    B200_TRY(gram_smem(d, &stages, &smem, "mmd_coef_i8"));
    CUtensorMap tmap;
    B200_TRY(make_tensor_map_2d(&tmap, z_dev, CU_TENSOR_MAP_DATA_TYPE_UINT8, (uint64_t)d_pad, (uint64_t)m, (uint64_t)d_pad, BK, 128));
    const float wmax = fmaxf(fabsf(w_xx), fabsf(w_xy));
    TcParams p = {};
    p.m_x = m_x; p.m = m; p.d = d;
    p.n_kblocks = (d_pad + BK - 1) / BK;
    p.tiles_m = (m + BM - 1) / BM;
    p.tiles_n = (m_pad + BN - 1) / BN;        // cover the zero padding columns as well
    p.row0 = row0; p.n_rows = n_rows;
    p.tiles_rows = (n_rows + BM - 1) / BM;
    p.total_tiles = p.tiles_rows * p.tiles_n;
    p.stages = stages;
    p.pass = TC_PASS_COEF;
    p.shard_rank = 0; p.shard_world = 1;
    p.lut = lut_dev;
    p.m_pad = m_pad; p.rows_alloc = rows_alloc; p.n_planes = n_planes;
    p.r_xx = wmax > 0.f ? w_xx / wmax : 0.f;
    p.r_xy = wmax > 0.f ? w_xy / wmax : 0.f;
    p.q_max = n_planes >= 3 ? 8355711.0f : 32639.0f;
    p.planes = planes_dev;
    p.rowsum = reinterpret_cast<long long *>(rowsum_dev);
    B200_CUDA(cudaFuncSetAttribute(mmd_gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    B200_CUDA(cudaMemsetAsync(rowsum_dev, 0, (size_t)n_rows * sizeof(int64_t), st));
    mmd_coef_lut_kernel<<<1, 1024, 0, st>>>(d, m, n_kernels, mul_factor, squared, bandwidth, sums_dev, w_xx, w_xy, n_planes,
                                            reinterpret_cast<const unsigned long long *>(hist_dev), lut_dev, scale_dev);
    B200_CUDA(cudaGetLastError());
    mmd_gram_i8_kernel<<<grid, TC_THREADS, smem, st>>>(tmap, p);
    B200_CUDA(cudaGetLastError());
    return 0;
}
