// grad_x[a][i] = g * (rowsum_a * z_ai - sum_b A_ab z_bi)   on tcgen05 int8 tensor cores.
//
// Second half of the MMD backward pass for +-1 rows (reference: dvae_loss.backward() through
// maximum_mean_discrepancy_loss, src/model_wrapper.py:320-326):
//     d MMD / d x_a = sum_b A_ab (x_a - z_b),   A_ab = w * (dk/dt)(dt/d||.||)/||.||
// A (n_rows x m) is real-valued, but it is a function of the integer Hamming distance only, so
// mmd_gram_i8_kernel<COEF> (mmd_tc.cu) writes it as 2 or 3 signed base-256 digit planes of a fixed-point
// number (16 / 24 bits of the largest |A|) together with the exact integer row sums.  Z is +-1, so every
// product is exact in int8 x int8 -> int32 and the whole contraction is integer arithmetic at the int8
// tensor rate (2x the bf16 rate; the first version of this pass used a bf16 hi/lo pair: 4 bytes and 2 bf16
// MMAs per coefficient instead of 2 bytes and 2 int8 MMAs).  The result is deterministic -- no floating-point
// summation order anywhere before the final scale.
//
// Operands: planes [n_planes][rows_alloc][K] int8 (K = m_pad, K-major), ZT [d][K] int8 (Z transposed, written
// once by the forward's spin extraction).  Same skeleton as mmd_tc.cu: TMA 128B-swizzled boxes -> 4-stage mbarrier
// ring -> tcgen05.mma.kind::i8 (M128 x N256 x K32) -> two TMEM accumulator stages.  The digit planes of one output
// tile run back to back through alternating TMEM stages; the epilogue warps keep the running value
// acc = acc * 256 + plane in registers (128 fp32 per thread), so plane p+1's MMAs overlap plane p's read-out.
//
// Tile order: bands of row tiles whose planes fit in L2 together with one wave's worth of ZT; inside a band the
// column tile is the slow index, so the band's A planes are read from HBM once and ZT once per band (the first
// version walked row tiles fastest over the whole matrix and re-streamed A once per wave: 8.2x the operand bytes
// in dram__bytes_read).  (Tried and dropped: a cp.async.bulk.prefetch.tensor cursor 16 k-blocks ahead of the loads, to
// cover the HBM latency of the streamed planes -- the kernel got 15 % slower, 1.14 -> 1.31 ms at cfg3: the prefetched
// boxes evict Z^T lines the other CTAs of the wave are about to re-read.)
#include "tc_common.cuh"

#include <cstdlib>

namespace b200grbm {

constexpr int I_BM = 128, I_BN = 256, I_BK = 128;       // bytes = int8 elements per k-block: one 128-byte swizzle row
constexpr int I_UMMA_K = 32;
constexpr int I_A_BYTES = I_BM * I_BK, I_B_BYTES = I_BN * I_BK, I_STAGE_BYTES = I_A_BYTES + I_B_BYTES;
constexpr int I_STAGES = 4, I_THREADS = 320, I_EPI_WARPS = 8;

struct GemmI8Params {
    int n_rows, d, K;                   // output rows (x rows of this call), features, contraction length (m_pad)
    int rows_alloc, n_planes;
    int tiles_m, tiles_n, total_tiles, band_rows;
    int kblocks;
    const long long *rowsum;            // [n_rows]
    const double *scale;                // device scalar: fixed-point unit of the planes
    const float *grad_out;              // device scalar: incoming gradient of the loss value
    const int8_t *z;                    // [.. ][d_pad] the +-1 rows; row z_row0 + a is x_a
    int z_row0, d_pad;
    float *grad_x;                      // [n_rows][d]
};

__device__ __forceinline__ void gemm_i8_tile(const GemmI8Params &p, int t, int &ti, int &tj)
{
    const int per_band = p.band_rows * p.tiles_n;
    const int band = t / per_band, r = t - band * per_band;
    const int rows_here = min(p.band_rows, p.tiles_m - band * p.band_rows);
    tj = r / rows_here;
    ti = band * p.band_rows + r % rows_here;
}

// Output stage of both GEMM kernels: grad_x[row][col .. col + 128) = gscale * (rowsum * z - accum), four columns per step
// so that no second 32-float array is alive next to the 128 accumulators.
__device__ __forceinline__ void gemm_i8_store_row(const GemmI8Params &p, const float (&accum)[128], int row, int col0, float gscale)
{
    const float rs = (float)__ldg(p.rowsum + row);
    const int8_t *zrow = p.z + (size_t)(p.z_row0 + row) * p.d_pad;
    float *out = p.grad_x + (size_t)row * p.d;
    const bool vec_ok = (p.d & 3) == 0;
#pragma unroll
    for (int g = 0; g < 32; ++g) {              // 32 groups of 4 columns
        const int col = col0 + 4 * g;
        if (col >= p.d) break;
        const uint32_t zw = *reinterpret_cast<const uint32_t *>(zrow + col);    // d_pad is a multiple of 128; pad columns are zero
        float r[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) r[c] = gscale * (rs * (float)(int8_t)((zw >> (8 * c)) & 0xffu) - accum[4 * g + c]);
        if (vec_ok && col + 4 <= p.d) {
            *reinterpret_cast<float4 *>(out + col) = make_float4(r[0], r[1], r[2], r[3]);
        } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (col + c < p.d) out[col + c] = r[c];
        }
    }
}

__global__ void __launch_bounds__(I_THREADS, 1) gemm_i8_planes_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                      const __grid_constant__ CUtensorMap map_b,
                                                                      const GemmI8Params p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)I_STAGES * I_STAGE_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * I_STAGES + 4);
    const uint32_t full0 = smem_addr(bars), empty0 = full0 + 8u * I_STAGES, tfull0 = empty0 + 8u * I_STAGES,
                   tempty0 = tfull0 + 16u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < I_STAGES; ++s) { bar_init(full0 + 8u * s, 1); bar_init(empty0 + 8u * s, 1); }
        for (int a = 0; a < 2; ++a) { bar_init(tfull0 + 8u * a, 1); bar_init(tempty0 + 8u * a, I_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_addr(tmem_slot), 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                int ti, tj;
                gemm_i8_tile(p, t, ti, tj);
                for (int pl = 0; pl < p.n_planes; ++pl) {
                    const int a_row = pl * p.rows_alloc + ti * I_BM;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        bar_wait(empty0 + 8u * s, ph ^ 1u);
                        const uint32_t fb = full0 + 8u * s;
                        const uint32_t dst = smem_addr(smem + (size_t)s * I_STAGE_BYTES);
                        bar_expect_tx(fb, I_STAGE_BYTES);
                        tma_load_2d(dst, &map_a, kb * I_BK, a_row, fb);
                        tma_load_2d(dst + I_A_BYTES, &map_b, kb * I_BK, tj * I_BN, fb);
                        tma_load_2d(dst + I_A_BYTES + I_A_BYTES, &map_b, kb * I_BK, tj * I_BN + 128, fb);
                        if (++s == I_STAGES) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---- MMA issuer
            const uint32_t idesc = umma_idesc_i8(I_BM, I_BN);
            int s = 0, unit = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                for (int pl = 0; pl < p.n_planes; ++pl, ++unit) {
                    const uint32_t acc = (uint32_t)unit & 1u, acc_ph = ((uint32_t)unit >> 1) & 1u;
                    bar_wait(tempty0 + 8u * acc, acc_ph ^ 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem_base + acc * I_BN;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        bar_wait(full0 + 8u * s, ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_addr = smem_addr(smem + (size_t)s * I_STAGE_BYTES);
                        const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + I_A_BYTES);
#pragma unroll
                        for (int k = 0; k < I_BK / I_UMMA_K; ++k)
                            umma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit(empty0 + 8u * s);
                        if (++s == I_STAGES) { s = 0; ph ^= 1u; }
                    }
                    umma_commit(tfull0 + 8u * acc);
                }
            }
        }
    } else {                                               // ---- epilogue: TMEM planes -> fixed point -> gradient rows
        const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
        const float gscale = (float)((double)__ldg(p.grad_out) * __ldg(p.scale));
        int unit = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            int ti, tj;
            gemm_i8_tile(p, t, ti, tj);
            const int row = ti * I_BM + quarter * 32 + lane;
            float accum[128];
#pragma unroll
            for (int k = 0; k < 128; ++k) accum[k] = 0.f;
            for (int pl = 0; pl < p.n_planes; ++pl, ++unit) {
                const uint32_t acc = (uint32_t)unit & 1u, acc_ph = ((uint32_t)unit >> 1) & 1u;
                bar_wait(tfull0 + 8u * acc, acc_ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int chunk = 0; chunk < 4; ++chunk) {
                    uint32_t v[32];
                    __syncwarp();
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * I_BN + (uint32_t)(half * 128 + chunk * 32), v);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float x = (float)(int)v[c];           // |x| <= 128 K < 2^24: exact
                        accum[chunk * 32 + c] = fmaf(accum[chunk * 32 + c], 256.0f, x);      // in place: acc = 256 acc + digit
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) bar_arrive(tempty0 + 8u * acc);
            }
            if (row < p.n_rows) gemm_i8_store_row(p, accum, row, tj * I_BN + half * 128, gscale);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair version (tcgen05 cta_group::2): two SMs of a TPC share one 256 x 256 output tile.  Each CTA stages its own
// 128 rows of the digit plane and HALF of the Z^T rows per k-block -- 32 KB instead of 48 KB per 128 x 256 x 128 MACs.
// The single-CTA kernel above moves 48 KB per k-block out of shared memory into the tensor core and as many in from
// TMA; at 128 B/clk that alone caps it near 70 % of the tensor rate.  Same barrier scheme as mmd_tc2.cu: full[s] lives
// in the leader (2 arrivals + both CTAs' transaction bytes), empty[s] / tmem_full[a] are arrived in both CTAs by the
// multicast tcgen05.commit, tmem_empty[a] lives in the leader (2 x 8 epilogue warps).
constexpr int I2_STAGE_BYTES = 2 * I_A_BYTES;        // 16 KB of A + 16 KB of B per CTA
constexpr int I2_STAGES = 6;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(I_THREADS, 1)
    gemm_i8_planes_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                               const GemmI8Params p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)I2_STAGES * I2_STAGE_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * I2_STAGES + 4);
    const uint32_t full0 = smem_addr(bars), empty0 = full0 + 8u * I2_STAGES, tfull0 = empty0 + 8u * I2_STAGES,
                   tempty0 = tfull0 + 16u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < I2_STAGES; ++s) { bar_init(full0 + 8u * s, 2); bar_init(empty0 + 8u * s, 1); }
        for (int a = 0; a < 2; ++a) { bar_init(tfull0 + 8u * a, 1); bar_init(tempty0 + 8u * a, 2 * I_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) tmem_alloc_2cta(smem_addr(tmem_slot), 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                      // the peer's barriers are initialised before anyone arrives remotely
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer (both CTAs)
            int s = 0;
            uint32_t ph = 0;
            for (int t = pair; t < p.total_tiles; t += n_pairs) {
                int ti, tj;
                gemm_i8_tile(p, t, ti, tj);
                for (int pl = 0; pl < p.n_planes; ++pl) {
                    const int a_row = pl * p.rows_alloc + ti * 256 + (int)rank * 128;
                    const int b_row = tj * I_BN + (int)rank * 128;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        bar_wait(empty0 + 8u * s, ph ^ 1u);
                        const uint32_t lead_full = mapa_cluster(full0 + 8u * s, 0);
                        const uint32_t dst = smem_addr(smem + (size_t)s * I2_STAGE_BYTES);
                        if (leader) bar_expect_tx(full0 + 8u * s, 2 * I2_STAGE_BYTES);
                        else bar_arrive_cluster(lead_full);
                        tma_load_2d_2sm(dst, &map_a, kb * I_BK, a_row, lead_full);
                        tma_load_2d_2sm(dst + I_A_BYTES, &map_b, kb * I_BK, b_row, lead_full);
                        if (++s == I2_STAGES) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {                         // ---- MMA issuer (leader CTA only)
            const uint32_t idesc = umma_idesc_i8(256, 256);
            int s = 0, unit = 0;
            uint32_t ph = 0;
            for (int t = pair; t < p.total_tiles; t += n_pairs) {
                for (int pl = 0; pl < p.n_planes; ++pl, ++unit) {
                    const uint32_t acc = (uint32_t)unit & 1u, acc_ph = ((uint32_t)unit >> 1) & 1u;
                    bar_wait(tempty0 + 8u * acc, acc_ph ^ 1u);               // both epilogues drained this accumulator
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t d_tmem = tmem_base + acc * I_BN;
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        bar_wait(full0 + 8u * s, ph);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t a_addr = smem_addr(smem + (size_t)s * I2_STAGE_BYTES);
                        const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + I_A_BYTES);
#pragma unroll
                        for (int k = 0; k < I_BK / I_UMMA_K; ++k)
                            umma_i8_2cta(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                        umma_commit_2cta(empty0 + 8u * s);                   // frees the stage in both CTAs
                        if (++s == I2_STAGES) { s = 0; ph ^= 1u; }
                    }
                    umma_commit_2cta(tfull0 + 8u * acc);                     // accumulator ready in both CTAs
                }
            }
        }
    } else {                                               // ---- epilogue (both CTAs)
        const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
        const float gscale = (float)((double)__ldg(p.grad_out) * __ldg(p.scale));
        const uint32_t lead_tempty0 = mapa_cluster(tempty0, 0);
        int unit = 0;
        for (int t = pair; t < p.total_tiles; t += n_pairs) {
            int ti, tj;
            gemm_i8_tile(p, t, ti, tj);
            const int row = ti * 256 + (int)rank * 128 + quarter * 32 + lane;
            float accum[128];
#pragma unroll
            for (int k = 0; k < 128; ++k) accum[k] = 0.f;
            for (int pl = 0; pl < p.n_planes; ++pl, ++unit) {
                const uint32_t acc = (uint32_t)unit & 1u, acc_ph = ((uint32_t)unit >> 1) & 1u;
                bar_wait(tfull0 + 8u * acc, acc_ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int chunk = 0; chunk < 4; ++chunk) {
                    uint32_t v[32];
                    __syncwarp();
                    tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * I_BN + (uint32_t)(half * 128 + chunk * 32), v);
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const float x = (float)(int)v[c];
                        accum[chunk * 32 + c] = fmaf(accum[chunk * 32 + c], 256.0f, x);      // in place: acc = 256 acc + digit
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) bar_arrive_cluster(lead_tempty0 + 8u * acc);
            }
            if (row < p.n_rows) gemm_i8_store_row(p, accum, row, tj * I_BN + half * 128, gscale);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                      // neither CTA frees TMEM / exits while the peer may still touch it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

// int8 transpose with zero fill: out[c][r] = in[r][c] for r < rows, c < cols; out has `out_rows` = cols rows of pitch
// out_pitch >= rows; columns r in [rows, out_pitch) are zeroed.  (ZT for callers that hold only the row-major matrix.)
__global__ void transpose_i8_kernel(const int8_t *__restrict__ in, int rows, int cols, int in_pitch, int8_t *__restrict__ out,
                                    int out_pitch)
{
    __shared__ int8_t tile[64][65];
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    for (int k = threadIdx.x; k < 64 * 64; k += blockDim.x) {
        const int r = k >> 6, c = k & 63;
        tile[r][c] = (r0 + r < rows && c0 + c < cols) ? in[(size_t)(r0 + r) * in_pitch + c0 + c] : (int8_t)0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < 64 * 64; k += blockDim.x) {
        const int c = k >> 6, r = k & 63;
        if (c0 + c < cols && r0 + r < out_pitch) out[(size_t)(c0 + c) * out_pitch + r0 + r] = tile[r][c];
    }
}

}  // namespace b200grbm

using namespace b200grbm;

extern "C" int32_t b200grbm_transpose_i8(const int8_t *in_dev, int32_t rows, int32_t cols, int32_t in_pitch, int8_t *out_dev,
                                         int32_t out_pitch, void *stream)
{
    if (rows <= 0 || cols <= 0 || in_pitch < cols || out_pitch < rows)
        return fail(B200GRBM_EINVAL, "transpose_i8: rows=%d cols=%d in_pitch=%d out_pitch=%d", rows, cols, in_pitch, out_pitch);
    if (!in_dev || !out_dev) return fail(B200GRBM_EINVAL, "transpose_i8: NULL pointer argument");
    B200_TRY(require_device());
    dim3 grid((cols + 63) / 64, (out_pitch + 63) / 64);
    transpose_i8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in_dev, rows, cols, in_pitch, out_dev, out_pitch);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_mmd_grad_i8(const int8_t *planes_dev, int32_t n_planes, int32_t n_rows, int32_t rows_alloc,
                                        int32_t m_pad, const int8_t *zt_dev, int32_t d, const int64_t *rowsum_dev,
                                        const double *scale_dev, const float *grad_out_dev, const int8_t *z_dev,
                                        int32_t z_row0, int32_t d_pad, float *grad_x_dev, void *stream)
{
    if (n_planes < 2 || n_planes > 3 || n_rows <= 0 || rows_alloc < n_rows || rows_alloc % 128 != 0 || m_pad <= 0 ||
        m_pad % 128 != 0 || d <= 0 || d_pad < d || d_pad % 128 != 0 || z_row0 < 0)
        return fail(B200GRBM_EINVAL, "mmd_grad_i8: n_planes=%d n_rows=%d rows_alloc=%d (multiple of 128) m_pad=%d (multiple of 128) "
                                     "d=%d d_pad=%d (multiple of 128)", n_planes, n_rows, rows_alloc, m_pad, d, d_pad);
    if (!planes_dev || !zt_dev || !rowsum_dev || !scale_dev || !grad_out_dev || !z_dev || !grad_x_dev)
        return fail(B200GRBM_EINVAL, "mmd_grad_i8: NULL pointer argument");
    if (((reinterpret_cast<uintptr_t>(planes_dev) | reinterpret_cast<uintptr_t>(zt_dev) | reinterpret_cast<uintptr_t>(z_dev) |
          reinterpret_cast<uintptr_t>(grad_x_dev)) & 15u) != 0)
        return fail(B200GRBM_EINVAL, "mmd_grad_i8: operands must be 16-byte aligned");
    B200_TRY(require_device());
    CUtensorMap ma, mb;
    B200_TRY(make_tensor_map_2d(&ma, planes_dev, CU_TENSOR_MAP_DATA_TYPE_UINT8, (uint64_t)m_pad, (uint64_t)n_planes * rows_alloc,
                                (uint64_t)m_pad, I_BK, 128));
    B200_TRY(make_tensor_map_2d(&mb, zt_dev, CU_TENSOR_MAP_DATA_TYPE_UINT8, (uint64_t)m_pad, (uint64_t)d, (uint64_t)m_pad, I_BK, 128));
    GemmI8Params p = {};
    p.n_rows = n_rows; p.d = d; p.K = m_pad;
    p.rows_alloc = rows_alloc; p.n_planes = n_planes;
    p.tiles_m = (n_rows + I_BM - 1) / I_BM;
    p.tiles_n = (d + I_BN - 1) / I_BN;
    p.total_tiles = p.tiles_m * p.tiles_n;
    p.kblocks = m_pad / I_BK;
    // band of row tiles: its planes (band x 128 x K x n_planes bytes) stay L2-resident while the column tiles sweep
    const size_t band_budget = (size_t)48 << 20;
    const size_t per_row_tile = (size_t)I_BM * m_pad * n_planes;
    int band = (int)(band_budget / per_row_tile);
    if (band < 1) band = 1;
    if (band > p.tiles_m) band = p.tiles_m;
    p.band_rows = band;
    p.rowsum = reinterpret_cast<const long long *>(rowsum_dev);
    p.scale = scale_dev;
    p.grad_out = grad_out_dev;
    p.z = z_dev; p.z_row0 = z_row0; p.d_pad = d_pad;
    p.grad_x = grad_x_dev;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    // CTA pairs (256-row tiles) once they fill the machine; B200GRBM_GEMM_TILE=1|2 forces a shape (A/B measurements, tests)
    const int pair_tiles = ((n_rows + 255) / 256) * p.tiles_n;
    const char *env = getenv("B200GRBM_GEMM_TILE");
    const bool use_pair = env != nullptr && (env[0] == '1' || env[0] == '2') ? env[0] == '2' : pair_tiles >= sms / 2;
    if (use_pair) {
        p.tiles_m = (n_rows + 255) / 256;
        p.total_tiles = pair_tiles;
        int band2 = (int)(band_budget / (2 * per_row_tile));
        if (band2 < 1) band2 = 1;
        if (band2 > p.tiles_m) band2 = p.tiles_m;
        p.band_rows = band2;
        const size_t smem2 = (size_t)I2_STAGES * I2_STAGE_BYTES + (2 * I2_STAGES + 4) * 8 + 16 + 1024;
        B200_CUDA(cudaFuncSetAttribute(gemm_i8_planes_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        int pairs = sms / 2;
        if (pairs > p.total_tiles) pairs = p.total_tiles;
        gemm_i8_planes_2cta_kernel<<<2 * pairs, I_THREADS, smem2, (cudaStream_t)stream>>>(ma, mb, p);     // __cluster_dims__(2,1,1)
        B200_CUDA(cudaGetLastError());
        return 0;
    }
    const size_t smem = (size_t)I_STAGES * I_STAGE_BYTES + (2 * I_STAGES + 4) * 8 + 16 + 1024;
    B200_CUDA(cudaFuncSetAttribute(gemm_i8_planes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    gemm_i8_planes_kernel<<<grid, I_THREADS, smem, (cudaStream_t)stream>>>(ma, mb, p);
    B200_CUDA(cudaGetLastError());
    return 0;
}
