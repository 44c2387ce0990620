// Mixture-of-RBF MMD for continuous (real-valued) rows on tcgen05 tensor cores: bf16 Gram
// with fp32 accumulation, optionally split-bf16 (hi.hi + hi.lo + lo.hi) for fp32-class accuracy.
//
// Same contract as b200grbm_mmd_forward_f32 / _i8 (reference: maximum_mean_discrepancy_loss(x, y,
// GaussianKernel(7)), call site src/model_wrapper.py:320; north star: "bf16/fp32-accumulate for
// continuous encoder latents, exp-sum over the bandwidths done in the epilogue").
//   ||a - b||^2 = |a|^2 + |b|^2 - 2 a.b   with the norms taken from the same rounded operands,
// so the distances are exactly those of the bf16 (or hi+lo) points; the epilogue turns each
// accumulator into t = ||.|| (or ||.||^2), evaluates sum_u exp2(t * c_u) with MUFU.EX2 and
// reduces the xx / yy / xy block sums in registers.  Nothing of size m^2 touches HBM.
// Skeleton shared with mmd_tc.cu / gemm_tc.cu (TMA ring -> tcgen05.mma.kind::f16 -> 2 TMEM stages).
#include "tc_common.cuh"

#include <cuda_bf16.h>

namespace b200grbm {

constexpr int F_BM = 128, F_BN = 256, F_BK = 64;
constexpr int F_UMMA_K = 16;
constexpr int F_A_BYTES = F_BM * F_BK * 2, F_B_BYTES = F_BN * F_BK * 2, F_STAGE_BYTES = F_A_BYTES + F_B_BYTES;
constexpr int F_STAGES = 4, F_THREADS = 320, F_EPI_WARPS = 8, F_MAX_KERNELS = 16;

enum { F_PASS_DIST = 0, F_PASS_KERNEL = 1, F_PASS_COEF = 2 };

struct BfParams {
    int m_x, m, k_pad;
    int kblocks_per_product, products;     // products = 1 (hi.hi) or 3 (hi.hi, hi.lo, lo.hi)
    int tiles_m, tiles_n, j0, p0, total_tiles;
    int pass, n_kernels, squared;
    float mul_factor, bandwidth;
    const float *norms;                    // [m] squared norms of the rounded rows
    double *sums;                          // [4]
    // F_PASS_COEF (backward): A[a][b] = w * (dk/dt) (dt/d||.||) / ||.|| as a bf16 (hi, lo) pair, x rows only
    int tiles_mx, m_pad;
    float w_xx, w_xy;
    __nv_bfloat16 *coef_hi, *coef_lo;
};

__device__ __forceinline__ void bf_tile_coords(const BfParams &p, int t, int &i, int &j)
{
    if (p.pass == F_PASS_COEF) {           // full rectangle: x rows against every column
        i = t % p.tiles_mx;
        j = t / p.tiles_mx;
        return;
    }
    if (t < p.p0) {
        j = (int)((sqrtf(4.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((j + 1) * (j + 2) <= t) ++j;
        while (j * (j + 1) > t) --j;
        i = t - j * (j + 1);
    } else {
        const int r = t - p.p0;
        j = p.j0 + r / p.tiles_m;
        i = r % p.tiles_m;
    }
}

__global__ void __launch_bounds__(F_THREADS, 1) mmd_gram_bf16_kernel(const __grid_constant__ CUtensorMap map_hi,
                                                                     const __grid_constant__ CUtensorMap map_lo,
                                                                     const BfParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *col_norms = reinterpret_cast<float *>(smem + (size_t)F_STAGES * F_STAGE_BYTES);     // [2][256]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)F_STAGES * F_STAGE_BYTES + 2 * F_BN * sizeof(float));
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * F_STAGES + 4);
    const uint32_t full0 = smem_addr(bars), empty0 = full0 + 8u * F_STAGES, tfull0 = empty0 + 8u * F_STAGES,
                   tempty0 = tfull0 + 16u;
    __shared__ double red[3][F_EPI_WARPS];

    if (threadIdx.x == 0) {
        for (int s = 0; s < F_STAGES; ++s) { bar_init(full0 + 8u * s, 1); bar_init(empty0 + 8u * s, 1); }
        for (int a = 0; a < 2; ++a) { bar_init(tfull0 + 8u * a, 1); bar_init(tempty0 + 8u * a, F_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hi) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_addr(tmem_slot), 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int kblocks = p.kblocks_per_product * p.products;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                int ti, tj;
                bf_tile_coords(p, t, ti, tj);
                for (int kb = 0; kb < kblocks; ++kb) {
                    const int prod = kb / p.kblocks_per_product;          // 0: hi.hi  1: hi.lo  2: lo.hi
                    const CUtensorMap *ma = prod == 2 ? &map_lo : &map_hi;
                    const CUtensorMap *mb = prod == 1 ? &map_lo : &map_hi;
                    const int kx = (kb % p.kblocks_per_product) * F_BK;
                    bar_wait(empty0 + 8u * s, ph ^ 1u);
                    const uint32_t fb = full0 + 8u * s;
                    const uint32_t dst = smem_addr(smem + (size_t)s * F_STAGE_BYTES);
                    bar_expect_tx(fb, F_STAGE_BYTES);
                    tma_load_2d(dst, ma, kx, ti * F_BM, fb);
                    tma_load_2d(dst + F_A_BYTES, mb, kx, tj * F_BN, fb);
                    tma_load_2d(dst + F_A_BYTES + F_A_BYTES, mb, kx, tj * F_BN + 128, fb);
                    if (++s == F_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---- MMA issuer
            const uint32_t idesc = umma_idesc_bf16(F_BM, F_BN);
            int s = 0, it = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
                const uint32_t acc = (uint32_t)it & 1u, acc_ph = ((uint32_t)it >> 1) & 1u;
                bar_wait(tempty0 + 8u * acc, acc_ph ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * F_BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    bar_wait(full0 + 8u * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_addr(smem + (size_t)s * F_STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + F_A_BYTES);
#pragma unroll
                    for (int k = 0; k < F_BK / F_UMMA_K; ++k)
                        umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty0 + 8u * s);
                    if (++s == F_STAGES) { s = 0; ph ^= 1u; }
                }
                umma_commit(tfull0 + 8u * acc);
            }
        }
    } else {
        // ---- epilogue: Gram -> distance -> sum_u exp2(t c_u) -> block sums
        const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
        const int et = threadIdx.x - 64;                   // 0..255 among the epilogue threads
        float c[F_MAX_KERNELS];
        {
            const double mm = (double)p.m;
            const double bw = p.bandwidth > 0.f ? (double)p.bandwidth : p.sums[3] / (mm * mm - mm);
#pragma unroll
            for (int u = 0; u < F_MAX_KERNELS; ++u)
                c[u] = u < p.n_kernels ? (float)(-1.4426950408889634 / (bw * pow((double)p.mul_factor, (double)(u - p.n_kernels / 2))))
                                       : 0.f;
        }
        double s_xx = 0.0, s_yy = 0.0, s_xy = 0.0;
        int it = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            int ti, tj;
            bf_tile_coords(p, t, ti, tj);
            const uint32_t acc = (uint32_t)it & 1u, acc_ph = ((uint32_t)it >> 1) & 1u;
            const int row0 = ti * F_BM, col0 = tj * F_BN;
            const int row = row0 + quarter * 32 + lane;
            const bool strict_upper = ti < 2 * tj;
            const bool rows_x = row0 + F_BM <= p.m_x, rows_y = row0 >= p.m_x;
            const bool cols_x = col0 + F_BN <= p.m_x, cols_y = col0 >= p.m_x;
            const bool pure = strict_upper && (row0 + F_BM <= p.m) && (col0 + F_BN <= p.m) && (rows_x || rows_y) &&
                              (cols_x || cols_y);
            // column norms of this tile, double-buffered with the accumulator stage
            float *cn = col_norms + acc * F_BN;
            cn[et] = (col0 + et < p.m) ? __ldg(p.norms + col0 + et) : 0.f;
            const float n_row = row < p.m ? __ldg(p.norms + row) : 0.f;
            asm volatile("bar.sync 2, 256;" ::: "memory");            // epilogue warps only
            bar_wait(tfull0 + 8u * acc, acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float a_xx = 0.f, a_yy = 0.f, a_xy = 0.f;
#pragma unroll 1
            for (int chunk = 0; chunk < 4; ++chunk) {
                uint32_t v[32];
                const int cbase = half * 128 + chunk * 32;
                __syncwarp();
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * F_BN + (uint32_t)cbase, v);
                if (p.pass == F_PASS_COEF) {
                    if (row < p.m_x) {
                        uint32_t hi[16], lo[16];
#pragma unroll
                        for (int q = 0; q < 32; q += 2) {
                            float cf[2];
#pragma unroll
                            for (int h2 = 0; h2 < 2; ++h2) {
                                const int col = col0 + cbase + q + h2;
                                const float d2 = fmaxf(fmaf(-2.f, u2f(v[q + h2]), n_row + cn[cbase + q + h2]), 0.f);
                                float tt;
                                if (p.squared) tt = d2;
                                else asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(tt) : "f"(d2));
                                float dk = 0.f;                       // dk/dt = ln2 * sum_u c_u exp2(t c_u)
#pragma unroll
                                for (int u = 0; u < F_MAX_KERNELS; ++u) {
                                    if (u < p.n_kernels) {
                                        float e;
                                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(tt * c[u]));
                                        dk = fmaf(c[u], e, dk);
                                    }
                                }
                                dk *= 0.6931471805599453f;
                                float val = p.squared ? 2.f * dk : (tt > 0.f ? __fdividef(dk, tt) : 0.f);
                                if (col >= p.m || col == row || d2 <= 0.f) val = 0.f;
                                cf[h2] = val * (col < p.m_x ? p.w_xx : p.w_xy);
                            }
                            const __nv_bfloat162 hh = __floats2bfloat162_rn(cf[0], cf[1]);
                            const __nv_bfloat162 ll = __floats2bfloat162_rn(cf[0] - __low2float(hh), cf[1] - __high2float(hh));
                            hi[q >> 1] = *reinterpret_cast<const uint32_t *>(&hh);
                            lo[q >> 1] = *reinterpret_cast<const uint32_t *>(&ll);
                        }
                        const size_t off = (size_t)row * p.m_pad + (size_t)(col0 + cbase);
#pragma unroll
                        for (int g4 = 0; g4 < 4; ++g4) {
                            if (col0 + cbase + 8 * g4 < p.m_pad) {
                                *reinterpret_cast<uint4 *>(p.coef_hi + off + 8 * g4) = make_uint4(hi[4 * g4], hi[4 * g4 + 1], hi[4 * g4 + 2], hi[4 * g4 + 3]);
                                *reinterpret_cast<uint4 *>(p.coef_lo + off + 8 * g4) = make_uint4(lo[4 * g4], lo[4 * g4 + 1], lo[4 * g4 + 2], lo[4 * g4 + 3]);
                            }
                        }
                    }
                    continue;
                }
#pragma unroll 4
                for (int q = 0; q < 32; ++q) {
                    const int col = col0 + cbase + q;
                    float d2 = fmaxf(fmaf(-2.f, u2f(v[q]), n_row + cn[cbase + q]), 0.f);
                    if (col == row) d2 = 0.f;
                    float tt;
                    if (p.squared) tt = d2;
                    else asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(tt) : "f"(d2));
                    float val;
                    if (p.pass == F_PASS_DIST) {
                        val = tt;
                    } else {
                        val = 0.f;
#pragma unroll
                        for (int u = 0; u < F_MAX_KERNELS; ++u) {
                            if (u < p.n_kernels) {
                                float e;
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(tt * c[u]));
                                val += e;
                            }
                        }
                    }
                    if (pure) {
                        a_xx += val;
                    } else if (row < p.m && col < p.m && col >= row) {
                        const bool rx = row < p.m_x, cx = col < p.m_x;
                        const float w = col == row ? 1.f : 2.f;
                        if (p.pass == F_PASS_DIST) a_xx += w * val;
                        else if (rx && cx) a_xx += w * val;
                        else if (!rx && !cx) a_yy += w * val;
                        else a_xy += val;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(tempty0 + 8u * acc);
            if (pure) {
                const double v2 = 2.0 * (double)a_xx;
                if (p.pass == F_PASS_DIST) s_xx += v2;
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
        for (int w = 0; w < F_EPI_WARPS; ++w) tot += red[threadIdx.x][w];
        if (p.pass == F_PASS_DIST) { if (threadIdx.x == 0) atomicAdd(p.sums + 3, tot); }
        else if (p.pass == F_PASS_KERNEL && tot != 0.0) atomicAdd(p.sums + threadIdx.x, tot);
    }
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200grbm

using namespace b200grbm;

extern "C" int32_t b200grbm_mmd_forward_bf16(const void *z_hi_dev, const void *z_lo_dev, const float *norms_dev, int32_t m_x,
                                             int32_t m_y, int32_t k_pad, int32_t n_kernels, float mul_factor, int32_t squared,
                                             float bandwidth, double *sums_dev, void *stream)
{
    if (m_x <= 0 || m_y <= 0 || k_pad <= 0 || k_pad % 64 != 0)
        return fail(B200GRBM_EINVAL, "mmd_forward_bf16: m_x=%d m_y=%d k_pad=%d (must be a multiple of 64)", m_x, m_y, k_pad);
    if (n_kernels < 1 || n_kernels > F_MAX_KERNELS || !(mul_factor > 0.f))
        return fail(B200GRBM_EINVAL, "mmd_forward_bf16: n_kernels=%d mul_factor=%g", n_kernels, mul_factor);
    if (!z_hi_dev || !norms_dev || !sums_dev) return fail(B200GRBM_EINVAL, "mmd_forward_bf16: NULL pointer argument");
    if (((reinterpret_cast<uintptr_t>(z_hi_dev) | reinterpret_cast<uintptr_t>(z_lo_dev)) & 15u) != 0)
        return fail(B200GRBM_EINVAL, "mmd_forward_bf16: operands must be 16-byte aligned (TMA)");
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const int m = m_x + m_y;
    CUtensorMap map_hi, map_lo;
    B200_TRY(make_tensor_map_2d(&map_hi, z_hi_dev, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)k_pad, (uint64_t)m,
                                (uint64_t)k_pad * 2, F_BK, 128));
    B200_TRY(make_tensor_map_2d(&map_lo, z_lo_dev ? z_lo_dev : z_hi_dev, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)k_pad,
                                (uint64_t)m, (uint64_t)k_pad * 2, F_BK, 128));
    BfParams p = {};
    p.m_x = m_x; p.m = m; p.k_pad = k_pad;
    p.kblocks_per_product = k_pad / F_BK;
    p.products = z_lo_dev ? 3 : 1;
    p.tiles_m = (m + F_BM - 1) / F_BM;
    p.tiles_n = (m + F_BN - 1) / F_BN;
    p.j0 = p.tiles_m / 2 < p.tiles_n ? p.tiles_m / 2 : p.tiles_n;
    p.p0 = p.j0 * (p.j0 + 1);
    p.total_tiles = p.p0 + (p.tiles_n - p.j0) * p.tiles_m;
    p.n_kernels = n_kernels; p.squared = squared; p.mul_factor = mul_factor; p.bandwidth = bandwidth;
    p.norms = norms_dev;
    p.sums = sums_dev;
    const size_t smem = (size_t)F_STAGES * F_STAGE_BYTES + 2 * F_BN * sizeof(float) + (2 * F_STAGES + 4) * 8 + 16 + 1024;
    B200_CUDA(cudaFuncSetAttribute(mmd_gram_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    B200_CUDA(cudaMemsetAsync(sums_dev, 0, 4 * sizeof(double), st));
    if (!(bandwidth > 0.f)) {
        p.pass = F_PASS_DIST;
        mmd_gram_bf16_kernel<<<grid, F_THREADS, smem, st>>>(map_hi, map_lo, p);
        B200_CUDA(cudaGetLastError());
    }
    p.pass = F_PASS_KERNEL;
    mmd_gram_bf16_kernel<<<grid, F_THREADS, smem, st>>>(map_hi, map_lo, p);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_mmd_coef_bf16(const void *z_hi_dev, const void *z_lo_dev, const float *norms_dev, int32_t m_x,
                                          int32_t m_y, int32_t k_pad, int32_t n_kernels, float mul_factor, int32_t squared,
                                          float bandwidth, const double *sums_dev, float w_xx, float w_xy, void *coef_hi_dev,
                                          void *coef_lo_dev, int32_t m_pad, void *stream)
{
    if (m_x <= 0 || m_y <= 0 || k_pad <= 0 || k_pad % 64 != 0)
        return fail(B200GRBM_EINVAL, "mmd_coef_bf16: m_x=%d m_y=%d k_pad=%d (must be a multiple of 64)", m_x, m_y, k_pad);
    const int m = m_x + m_y;
    if (m_pad < m || m_pad % 64 != 0) return fail(B200GRBM_EINVAL, "mmd_coef_bf16: m_pad=%d must be a multiple of 64 >= m=%d", m_pad, m);
    if (n_kernels < 1 || n_kernels > F_MAX_KERNELS || !(mul_factor > 0.f))
        return fail(B200GRBM_EINVAL, "mmd_coef_bf16: n_kernels=%d mul_factor=%g", n_kernels, mul_factor);
    if (!z_hi_dev || !norms_dev || !sums_dev || !coef_hi_dev || !coef_lo_dev)
        return fail(B200GRBM_EINVAL, "mmd_coef_bf16: NULL pointer argument");
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap map_hi, map_lo;
    B200_TRY(make_tensor_map_2d(&map_hi, z_hi_dev, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)k_pad, (uint64_t)m,
                                (uint64_t)k_pad * 2, F_BK, 128));
    B200_TRY(make_tensor_map_2d(&map_lo, z_lo_dev ? z_lo_dev : z_hi_dev, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)k_pad,
                                (uint64_t)m, (uint64_t)k_pad * 2, F_BK, 128));
    BfParams p = {};
    p.m_x = m_x; p.m = m; p.k_pad = k_pad;
    p.kblocks_per_product = k_pad / F_BK;
    p.products = z_lo_dev ? 3 : 1;
    p.tiles_m = (m + F_BM - 1) / F_BM;
    p.tiles_n = (m_pad + F_BN - 1) / F_BN;
    p.tiles_mx = (m_x + F_BM - 1) / F_BM;
    p.total_tiles = p.tiles_mx * p.tiles_n;
    p.pass = F_PASS_COEF;
    p.n_kernels = n_kernels; p.squared = squared; p.mul_factor = mul_factor; p.bandwidth = bandwidth;
    p.norms = norms_dev;
    p.sums = const_cast<double *>(sums_dev);
    p.m_pad = m_pad; p.w_xx = w_xx; p.w_xy = w_xy;
    p.coef_hi = reinterpret_cast<__nv_bfloat16 *>(coef_hi_dev);
    p.coef_lo = reinterpret_cast<__nv_bfloat16 *>(coef_lo_dev);
    const size_t smem = (size_t)F_STAGES * F_STAGE_BYTES + 2 * F_BN * sizeof(float) + (2 * F_STAGES + 4) * 8 + 16 + 1024;
    B200_CUDA(cudaFuncSetAttribute(mmd_gram_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    mmd_gram_bf16_kernel<<<grid, F_THREADS, smem, st>>>(map_hi, map_lo, p);
    B200_CUDA(cudaGetLastError());
    return 0;
}
