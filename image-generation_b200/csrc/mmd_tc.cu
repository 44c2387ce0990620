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
//   * and in the epilogue maps each Gram entry through a (D+1)-entry look-up table held in
//     shared memory -- LUT[h] = distance (pass 1, for the data-dependent bandwidth) or
//     sum_u exp(-t_h / (bw * mult_u)) (pass 2), computed once in float64 -- and reduces the
//     xx / yy / xy block sums in registers.  Nothing of size m^2 touches HBM.
// Only the upper triangle of 128 x 256 tiles is visited (the kernel matrix is symmetric).
//
// Warp roles (320 threads, one persistent CTA per SM): warp 0 lane 0 = TMA producer,
// warp 1 = TMEM allocator + (lane 0) MMA issuer, warps 2..9 = epilogue (two column halves x
// four TMEM lane quarters).
#include "tc_common.cuh"

#include <cuda_bf16.h>
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

enum { TC_PASS_DIST = 0, TC_PASS_KERNEL = 1, TC_PASS_COEF = 2 };
constexpr int TC_MAX_SB = 136;          // up to 16 super-blocks per side (m <= 65536)
constexpr int TC_SB_ROWS = 32, TC_SB_COLS = 16;   // tiles per super-block: 32 x 128 rows, 16 x 256 columns

struct TcParams {
    int m_x, m, d;
    int n_kblocks;
    int tiles_m, tiles_n, j0, p0, total_tiles;
    int stages;
    int pass;
    const float *lut;   // [d + 1]
    double *sums;       // [4]
    // TC_PASS_COEF: backward coefficients A[a][b] = w * c(h_ab), written as a bf16 (hi, lo) pair
    int tiles_mx;       // row tiles covering the x rows
    int m_pad;          // row pitch (elements) of coef_hi / coef_lo, multiple of 64
    float w_xx, w_xy;
    __nv_bfloat16 *coef_hi, *coef_lo;
    // L2-sized super-blocks of the tile triangle (forward passes): 4096 x 4096 entries = 32 x 16 tiles per
    // block pair, so the rows a wave of CTAs touches (2 x 23 MB at D = 5640) stay resident in one L2 partition
    int sb_count;                       // 0 -> plain column-major triangle
    int sb_start[TC_MAX_SB + 1];
    unsigned char sb_bi[TC_MAX_SB], sb_bj[TC_MAX_SB];
};

// upper-triangle tile enumeration: column tile j ascending, row tiles i = 0 .. min(tiles_m, 2j+2) - 1
// LUT index = Hamming distance (D - a.b) / 2; clamped so that rows that are not +-1 (caller error)
// can never read outside the table
__device__ __forceinline__ int lut_index(int two_d, int gram, int d)
{
    return min(max((two_d - 2 * gram) >> 2, 0), d);
}

__device__ __forceinline__ void tile_coords(const TcParams &p, int t, int &i, int &j)
{
    if (p.pass == TC_PASS_COEF) {          // full rectangle: x rows against every column
        i = t % p.tiles_mx;
        j = t / p.tiles_mx;
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
    float *lut = reinterpret_cast<float *>(smem + (size_t)S * STAGE_BYTES);  // d + 1 floats
    const size_t lut_bytes = ((size_t)(p.d + 1) * 4 + 15) / 16 * 16;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * STAGE_BYTES + lut_bytes);
    // bars: full[S], empty[S], tmem_full[2], tmem_empty[2]; then the TMEM base address
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 4);
    const uint32_t full0 = smem_addr(bars), empty0 = full0 + 8u * S, tfull0 = empty0 + 8u * S, tempty0 = tfull0 + 16u;
    __shared__ double red[3][EPI_WARPS];

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { bar_init(full0 + 8u * s, 1); bar_init(empty0 + 8u * s, 1); }
        for (int a = 0; a < 2; ++a) { bar_init(tfull0 + 8u * a, 1); bar_init(tempty0 + 8u * a, EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)),
                     "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int k = threadIdx.x; k <= p.d; k += blockDim.x) lut[k] = p.lut[k];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                int ti, tj;
                tile_coords(p, t, ti, tj);
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    bar_wait(empty0 + 8u * s, ph ^ 1u);
                    const uint32_t fb = full0 + 8u * s;
                    const uint32_t a_dst = smem_addr(stage_base + (size_t)s * STAGE_BYTES);
                    bar_expect_tx(fb, STAGE_BYTES);
                    tma_load_2d(a_dst, &tmap, kb * BK, ti * BM, fb);
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
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
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
        // ===================== epilogue: TMEM -> LUT -> block sums =====================
        const int ew = warp - 2;                 // 0..7
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int half = ew >> 2;                // column half of the accumulator
        double s_xx = 0.0, s_yy = 0.0, s_xy = 0.0;
        const int two_d = 2 * p.d;
        int it = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            int ti, tj;
            tile_coords(p, t, ti, tj);
            const uint32_t acc = (uint32_t)it & 1u, acc_ph = ((uint32_t)it >> 1) & 1u;
            const int row0 = ti * BM, col0 = tj * BN;
            const int row = row0 + quarter * 32 + lane;
            const bool strict_upper = ti < 2 * tj;          // every column of the tile is right of every row
            const bool rows_x = row0 + BM <= p.m_x, rows_y = row0 >= p.m_x;
            const bool cols_x = col0 + BN <= p.m_x, cols_y = col0 >= p.m_x;
            const bool pure = strict_upper && (row0 + BM <= p.m) && (col0 + BN <= p.m) && (rows_x || rows_y) &&
                              (cols_x || cols_y);
            bar_wait(tfull0 + 8u * acc, acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float a_xx = 0.f, a_yy = 0.f, a_xy = 0.f;
#pragma unroll 1
            for (int chunk = 0; chunk < 4; ++chunk) {
                uint32_t v[32];
                const int cbase = half * 128 + chunk * 32;
                __syncwarp();                                   // tcgen05.ld is .sync.aligned
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN + (uint32_t)cbase, v);
                if (p.pass == TC_PASS_COEF) {
                    // 32 consecutive coefficients of this thread's row -> bf16 hi / lo halves, 64 B each
                    if (row < p.m_x) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int c = 0; c < 32; c += 2) {
                            float cf[2];
#pragma unroll
                            for (int q = 0; q < 2; ++q) {
                                const int col = col0 + cbase + c + q;
                                const float raw = lut[lut_index(two_d, (int)v[c + q], p.d)];
                                cf[q] = (col < p.m && col != row) ? raw * (col < p.m_x ? p.w_xx : p.w_xy) : 0.f;
                            }
                            const __nv_bfloat162 h2 = __floats2bfloat162_rn(cf[0], cf[1]);
                            const __nv_bfloat162 l2 = __floats2bfloat162_rn(cf[0] - __low2float(h2), cf[1] - __high2float(h2));
                            hi[c >> 1] = *reinterpret_cast<const uint32_t *>(&h2);
                            lo[c >> 1] = *reinterpret_cast<const uint32_t *>(&l2);
                        }
                        const size_t off = (size_t)row * p.m_pad + (size_t)(col0 + cbase);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (col0 + cbase + 8 * q < p.m_pad) {
                                *reinterpret_cast<uint4 *>(p.coef_hi + off + 8 * q) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                                *reinterpret_cast<uint4 *>(p.coef_lo + off + 8 * q) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
                            }
                        }
                    }
                } else if (pure) {
                    float part = 0.f;
#pragma unroll
                    for (int c = 0; c < 32; ++c) part += lut[lut_index(two_d, (int)v[c], p.d)];
                    a_xx += part;            // sorted into its block after the loop
                } else {
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int col = col0 + cbase + c;
                        if (row < p.m && col < p.m && col >= row) {
                            const float kv = lut[lut_index(two_d, (int)v[c], p.d)];
                            const bool rx = row < p.m_x, cx = col < p.m_x;
                            const float w = col == row ? 1.f : 2.f;
                            if (p.pass == TC_PASS_DIST) a_xx += w * kv;
                            else if (rx && cx) a_xx += w * kv;
                            else if (!rx && !cx) a_yy += w * kv;
                            else a_xy += kv;
                        }
                    }
                }
            }
            // accumulator drained: hand the TMEM stage back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(tempty0 + 8u * acc);
            if (pure) {
                const double v2 = 2.0 * (double)a_xx;
                if (p.pass == TC_PASS_DIST) s_xx += v2;
                else if (rows_x && cols_x) s_xx += v2;
                else if (rows_y && cols_y) s_yy += v2;
                else s_xy += (double)a_xx;
            } else {
                s_xx += (double)a_xx; s_yy += (double)a_yy; s_xy += (double)a_xy;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s_xx += __shfl_xor_sync(0xffffffffu, s_xx, o);
            s_yy += __shfl_xor_sync(0xffffffffu, s_yy, o);
            s_xy += __shfl_xor_sync(0xffffffffu, s_xy, o);
        }
        if (lane == 0) { red[0][ew] = s_xx; red[1][ew] = s_yy; red[2][ew] = s_xy; }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 3) {
        double tot = 0.0;
        for (int w = 0; w < EPI_WARPS; ++w) tot += red[threadIdx.x][w];
        if (p.pass == TC_PASS_DIST) { if (threadIdx.x == 0) atomicAdd(p.sums + 3, tot); }
        else if (p.pass == TC_PASS_KERNEL && tot != 0.0) atomicAdd(p.sums + threadIdx.x, tot);
    }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// LUT over the Hamming distance h = 0..d:  t_h = 2 sqrt(h) (or 4h when squared)
__global__ void mmd_lut_kernel(int pass, int d, int m, int n_kernels, float mul_factor, int squared, float bandwidth,
                               const double *sums, float *lut)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h > d) return;
    const double t = squared ? 4.0 * (double)h : 2.0 * sqrt((double)h);
    if (pass == TC_PASS_DIST) { lut[h] = (float)t; return; }
    const double mm = (double)m;
    const double bw = bandwidth > 0.f ? (double)bandwidth : sums[3] / (mm * mm - mm);
    double k = 0.0, dk = 0.0;                       // k(t) and dk/dt
    for (int u = 0; u < n_kernels; ++u) {
        const double b = bw * pow((double)mul_factor, (double)(u - n_kernels / 2));
        const double e = exp(-t / b);
        k += e;
        dk -= e / b;
    }
    if (pass == TC_PASS_KERNEL) { lut[h] = (float)k; return; }
    // TC_PASS_COEF: (dk/dt) (dt/d||.||) / ||.||  -> multiplies (x_a - z_b);  zero where the rows coincide
    lut[h] = h == 0 ? 0.f : (float)(squared ? 2.0 * dk : dk / t);
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

// CTA-pair version of the forward passes (mmd_tc2.cu)
int32_t launch_gram_i8_2cta(const CUtensorMap &tmap, int m_x, int m, int d, int d_pad, int pass, const float *lut,
                            double *sums, cudaStream_t st);

// Forward tile shape.  Measured on B200 (8192 + 8192 rows): while the sample matrix fits in L2 the single-CTA
// kernel and the CTA-pair kernel run at the same k-block rate (742 vs 750 clk per 128-byte k-block at
// D = 5632); once it does not (D = 11264, 185 MB) the pair kernel's 33 % smaller operand traffic wins
// (1.04 vs 1.16 ms).  B200GRBM_MMD_TILE=1|2 forces one of them (A/B measurements, parity tests of both).
static bool use_pair_kernel(int m, int d_pad)
{
    const char *env = getenv("B200GRBM_MMD_TILE");
    if (env != nullptr && env[0] == '1') return false;
    if (env != nullptr && env[0] == '2') return true;
    const int tiles = (m + 255) / 256;
    return tiles * (tiles + 1) / 2 >= 74 && (size_t)m * (size_t)d_pad > (size_t)100 << 20;
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

extern "C" int32_t b200grbm_mmd_forward_i8(const int8_t *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t d_pad,
                                           int32_t n_kernels, float mul_factor, int32_t squared, float bandwidth,
                                           float *lut_dev, double *sums_dev, void *stream)
{
    if (m_x <= 0 || m_y <= 0 || d <= 0 || d_pad < d || d_pad % 16 != 0)
        return fail(B200GRBM_EINVAL, "mmd_forward_i8: m_x=%d m_y=%d d=%d d_pad=%d", m_x, m_y, d, d_pad);
    if (n_kernels < 1 || n_kernels > 16 || !(mul_factor > 0.f))
        return fail(B200GRBM_EINVAL, "mmd_forward_i8: n_kernels=%d mul_factor=%g", n_kernels, mul_factor);
    if (!z_dev || !lut_dev || !sums_dev) return fail(B200GRBM_EINVAL, "mmd_forward_i8: NULL pointer argument");
    if ((reinterpret_cast<uintptr_t>(z_dev) & 15u) != 0)
        return fail(B200GRBM_EINVAL, "mmd_forward_i8: z_dev must be 16-byte aligned (TMA)");
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const int m = m_x + m_y;

    int dev = 0, smem_optin = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t lut_bytes = ((size_t)(d + 1) * 4 + 15) / 16 * 16;
    int stages = 4;
    size_t smem = 0;
    for (; stages >= 2; --stages) {
        smem = (size_t)stages * STAGE_BYTES + lut_bytes + (2 * stages + 4) * 8 + 16;
        if (smem + 1024 <= (size_t)smem_optin) break;
    }
    if (stages < 2)
        return fail(B200GRBM_EUNSUPPORTED, "mmd_forward_i8: d=%d needs a %zu B look-up table, too large for shared memory", d,
                    lut_bytes);

    CUtensorMap tmap;
    B200_TRY(make_tensor_map_2d(&tmap, z_dev, CU_TENSOR_MAP_DATA_TYPE_UINT8, (uint64_t)d_pad, (uint64_t)m, (uint64_t)d_pad, BK, 128));

    TcParams p = {};
    p.m_x = m_x; p.m = m; p.d = d;
    p.n_kblocks = (d_pad + BK - 1) / BK;
    p.tiles_m = (m + BM - 1) / BM;
    p.tiles_n = (m + BN - 1) / BN;
    p.j0 = p.tiles_m / 2 < p.tiles_n ? p.tiles_m / 2 : p.tiles_n;
    p.p0 = p.j0 * (p.j0 + 1);
    p.total_tiles = p.p0 + (p.tiles_n - p.j0) * p.tiles_m;
    p.stages = stages;
    p.lut = lut_dev;
    p.sums = sums_dev;
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
                return fail(B200GRBM_EINVAL, "mmd_forward_i8: internal tile enumeration mismatch (%d vs %d)", start, p.total_tiles);
        }
    }

    B200_CUDA(cudaFuncSetAttribute(mmd_gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem + 1024)));
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    B200_CUDA(cudaMemsetAsync(sums_dev, 0, 4 * sizeof(double), st));
    const int lut_blocks = (d + 1 + 255) / 256;
    const bool pair = use_pair_kernel(m, d_pad);
    if (!(bandwidth > 0.f)) {
        mmd_lut_kernel<<<lut_blocks, 256, 0, st>>>(TC_PASS_DIST, d, m, n_kernels, mul_factor, squared, bandwidth, sums_dev,
                                                   lut_dev);
        B200_CUDA(cudaGetLastError());
        if (pair) {
            B200_TRY(launch_gram_i8_2cta(tmap, m_x, m, d, d_pad, TC_PASS_DIST, lut_dev, sums_dev, st));
        } else {
            p.pass = TC_PASS_DIST;
            mmd_gram_i8_kernel<<<grid, TC_THREADS, smem + 1024, st>>>(tmap, p);
            B200_CUDA(cudaGetLastError());
        }
    }
    mmd_lut_kernel<<<lut_blocks, 256, 0, st>>>(TC_PASS_KERNEL, d, m, n_kernels, mul_factor, squared, bandwidth, sums_dev,
                                               lut_dev);
    B200_CUDA(cudaGetLastError());
    if (pair) {
        B200_TRY(launch_gram_i8_2cta(tmap, m_x, m, d, d_pad, TC_PASS_KERNEL, lut_dev, sums_dev, st));
    } else {
        p.pass = TC_PASS_KERNEL;
        mmd_gram_i8_kernel<<<grid, TC_THREADS, smem + 1024, st>>>(tmap, p);
        B200_CUDA(cudaGetLastError());
    }
    return 0;
}

extern "C" int32_t b200grbm_mmd_coef_i8(const int8_t *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t d_pad,
                                        int32_t n_kernels, float mul_factor, int32_t squared, float bandwidth,
                                        const double *sums_dev, float w_xx, float w_xy, float *lut_dev,
                                        void *coef_hi_dev, void *coef_lo_dev, int32_t m_pad, void *stream)
{
    if (m_x <= 0 || m_y <= 0 || d <= 0 || d_pad < d || d_pad % 16 != 0)
        return fail(B200GRBM_EINVAL, "mmd_coef_i8: m_x=%d m_y=%d d=%d d_pad=%d", m_x, m_y, d, d_pad);
    const int m = m_x + m_y;
    if (m_pad < m || m_pad % 64 != 0) return fail(B200GRBM_EINVAL, "mmd_coef_i8: m_pad=%d must be a multiple of 64 >= m=%d", m_pad, m);
    if (n_kernels < 1 || n_kernels > 16 || !(mul_factor > 0.f))
        return fail(B200GRBM_EINVAL, "mmd_coef_i8: n_kernels=%d mul_factor=%g", n_kernels, mul_factor);
    if (!z_dev || !lut_dev || !sums_dev || !coef_hi_dev || !coef_lo_dev)
        return fail(B200GRBM_EINVAL, "mmd_coef_i8: NULL pointer argument");
    if (((reinterpret_cast<uintptr_t>(z_dev) | reinterpret_cast<uintptr_t>(coef_hi_dev) | reinterpret_cast<uintptr_t>(coef_lo_dev)) & 15u) != 0)
        return fail(B200GRBM_EINVAL, "mmd_coef_i8: z_dev / coef buffers must be 16-byte aligned");
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, smem_optin = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t lut_bytes = ((size_t)(d + 1) * 4 + 15) / 16 * 16;
    int stages = 4;
    size_t smem = 0;
    for (; stages >= 2; --stages) {
        smem = (size_t)stages * STAGE_BYTES + lut_bytes + (2 * stages + 4) * 8 + 16;
        if (smem + 1024 <= (size_t)smem_optin) break;
    }
    if (stages < 2) return fail(B200GRBM_EUNSUPPORTED, "mmd_coef_i8: d=%d look-up table does not fit shared memory", d);
    CUtensorMap tmap;
    B200_TRY(make_tensor_map_2d(&tmap, z_dev, CU_TENSOR_MAP_DATA_TYPE_UINT8, (uint64_t)d_pad, (uint64_t)m, (uint64_t)d_pad, BK, 128));
    TcParams p = {};
    p.m_x = m_x; p.m = m; p.d = d;
    p.n_kblocks = (d_pad + BK - 1) / BK;
    p.tiles_m = (m + BM - 1) / BM;
    p.tiles_n = (m_pad + BN - 1) / BN;        // cover the zero padding columns as well
    p.tiles_mx = (m_x + BM - 1) / BM;
    p.total_tiles = p.tiles_mx * p.tiles_n;
    p.stages = stages;
    p.pass = TC_PASS_COEF;
    p.lut = lut_dev;
    p.sums = const_cast<double *>(sums_dev);
    p.m_pad = m_pad; p.w_xx = w_xx; p.w_xy = w_xy;
    p.coef_hi = reinterpret_cast<__nv_bfloat16 *>(coef_hi_dev);
    p.coef_lo = reinterpret_cast<__nv_bfloat16 *>(coef_lo_dev);
    B200_CUDA(cudaFuncSetAttribute(mmd_gram_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem + 1024)));
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    mmd_lut_kernel<<<(d + 1 + 255) / 256, 256, 0, st>>>(TC_PASS_COEF, d, m, n_kernels, mul_factor, squared, bandwidth, sums_dev,
                                                        lut_dev);
    B200_CUDA(cudaGetLastError());
    mmd_gram_i8_kernel<<<grid, TC_THREADS, smem + 1024, st>>>(tmap, p);
    B200_CUDA(cudaGetLastError());
    return 0;
}
