// C[M x N] (fp32) = (A_hi + A_lo)[M x K] * B[N x K]^T  in bf16 on tcgen05 tensor cores.
//
// Second half of the MMD backward pass for +-1 rows (reference: dvae_loss.backward() through
// maximum_mean_discrepancy_loss, src/model_wrapper.py:320-326):
//     grad_x[a] = sum_b A_ab (x_a - z_b) = rowsum_a x_a - (A Z)_a
// A (m_x x m) is the real-valued coefficient matrix written by mmd_gram_i8_kernel<COEF> as a
// bf16 (hi, lo) pair (relative error 2^-16), Z is +-1 (exact in bf16), B = [Z^T; 1] so the
// extra output column is rowsum_a.  Same skeleton as mmd_tc.cu: TMA 128B-swizzled boxes ->
// 4-stage mbarrier ring -> tcgen05.mma.kind::f16 (M128 x N256 x K16, fp32 accumulators in two
// TMEM stages) -> epilogue warps store fp32 rows.
#include "tc_common.cuh"

namespace b200grbm {

constexpr int G_BM = 128, G_BN = 256, G_BK = 64;     // bf16 elements: 64 * 2 B = one 128-byte swizzle row
constexpr int G_UMMA_K = 16;
constexpr int G_A_BYTES = G_BM * G_BK * 2, G_B_BYTES = G_BN * G_BK * 2, G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_STAGES = 4, G_THREADS = 320, G_EPI_WARPS = 8;

struct GemmParams {
    int M, N, K, ldc;
    int tiles_m, tiles_n, total_tiles;
    int kblocks_per_pass, passes;     // passes = 2 when a lo matrix is present
    float *c;
};

__global__ void __launch_bounds__(G_THREADS, 1) gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a_hi,
                                                                    const __grid_constant__ CUtensorMap map_a_lo,
                                                                    const __grid_constant__ CUtensorMap map_b,
                                                                    const GemmParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)G_STAGES * G_STAGE_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * G_STAGES + 4);
    const uint32_t full0 = smem_addr(bars), empty0 = full0 + 8u * G_STAGES, tfull0 = empty0 + 8u * G_STAGES,
                   tempty0 = tfull0 + 16u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G_STAGES; ++s) { bar_init(full0 + 8u * s, 1); bar_init(empty0 + 8u * s, 1); }
        for (int a = 0; a < 2; ++a) { bar_init(tfull0 + 8u * a, 1); bar_init(tempty0 + 8u * a, G_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    }
    if (warp == 1) tmem_alloc(smem_addr(tmem_slot), 512);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int kblocks = p.kblocks_per_pass * p.passes;

    if (warp == 0) {
        if (lane == 0) {                                   // ---- TMA producer
            int s = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const int ti = t % p.tiles_m, tj = t / p.tiles_m;     // row tiles fastest: B tile shared by neighbours in L2
                for (int kb = 0; kb < kblocks; ++kb) {
                    const CUtensorMap *ma = kb < p.kblocks_per_pass ? &map_a_hi : &map_a_lo;
                    const int kx = (kb % p.kblocks_per_pass) * G_BK;
                    bar_wait(empty0 + 8u * s, ph ^ 1u);
                    const uint32_t fb = full0 + 8u * s;
                    const uint32_t dst = smem_addr(smem + (size_t)s * G_STAGE_BYTES);
                    bar_expect_tx(fb, G_STAGE_BYTES);
                    tma_load_2d(dst, ma, kx, ti * G_BM, fb);
                    tma_load_2d(dst + G_A_BYTES, &map_b, kx, tj * G_BN, fb);
                    tma_load_2d(dst + G_A_BYTES + G_A_BYTES, &map_b, kx, tj * G_BN + 128, fb);
                    if (++s == G_STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                                   // ---- MMA issuer
            const uint32_t idesc = umma_idesc_bf16(G_BM, G_BN);
            int s = 0, it = 0;
            uint32_t ph = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
                const uint32_t acc = (uint32_t)it & 1u, acc_ph = ((uint32_t)it >> 1) & 1u;
                bar_wait(tempty0 + 8u * acc, acc_ph ^ 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * G_BN;
                for (int kb = 0; kb < kblocks; ++kb) {
                    bar_wait(full0 + 8u * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_addr(smem + (size_t)s * G_STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + G_A_BYTES);
#pragma unroll
                    for (int k = 0; k < G_BK / G_UMMA_K; ++k)      // 16 bf16 = 32 bytes = +2 in the address field
                        umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit(empty0 + 8u * s);
                    if (++s == G_STAGES) { s = 0; ph ^= 1u; }
                }
                umma_commit(tfull0 + 8u * acc);
            }
        }
    } else {                                               // ---- epilogue: TMEM -> global fp32
        const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
        int it = 0;
        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++it) {
            const int ti = t % p.tiles_m, tj = t / p.tiles_m;
            const uint32_t acc = (uint32_t)it & 1u, acc_ph = ((uint32_t)it >> 1) & 1u;
            const int row = ti * G_BM + quarter * 32 + lane;
            bar_wait(tfull0 + 8u * acc, acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int chunk = 0; chunk < 4; ++chunk) {
                uint32_t v[32];
                const int cbase = half * 128 + chunk * 32;
                __syncwarp();
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * G_BN + (uint32_t)cbase, v);
                const int col = tj * G_BN + cbase;
                if (row < p.M) {
                    float *dst = p.c + (size_t)row * p.ldc + col;
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (col + 4 * q < p.ldc)              // ldc is a multiple of 4: whole float4 groups only
                            *reinterpret_cast<uint4 *>(dst + 4 * q) = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(tempty0 + 8u * acc);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200grbm

using namespace b200grbm;

extern "C" int32_t b200grbm_gemm_bf16_tn(const void *a_hi_dev, const void *a_lo_dev, int32_t M, int32_t K, int32_t a_rows_alloc,
                                         const void *b_dev, int32_t N, float *c_dev, int32_t ldc, void *stream)
{
    if (M <= 0 || N <= 0 || K <= 0 || K % 64 != 0 || ldc < N || ldc % 4 != 0 || a_rows_alloc < M)
        return fail(B200GRBM_EINVAL, "gemm_bf16_tn: M=%d N=%d K=%d (multiple of 64) ldc=%d (multiple of 4, >= N) a_rows=%d", M, N,
                    K, ldc, a_rows_alloc);
    if (!a_hi_dev || !b_dev || !c_dev) return fail(B200GRBM_EINVAL, "gemm_bf16_tn: NULL pointer argument");
    if (((reinterpret_cast<uintptr_t>(a_hi_dev) | reinterpret_cast<uintptr_t>(a_lo_dev) | reinterpret_cast<uintptr_t>(b_dev) |
          reinterpret_cast<uintptr_t>(c_dev)) & 15u) != 0)
        return fail(B200GRBM_EINVAL, "gemm_bf16_tn: operands must be 16-byte aligned");
    B200_TRY(require_device());
    CUtensorMap ma_hi, ma_lo, mb;
    B200_TRY(make_tensor_map_2d(&ma_hi, a_hi_dev, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)a_rows_alloc,
                                (uint64_t)K * 2, G_BK, 128));
    B200_TRY(make_tensor_map_2d(&ma_lo, a_lo_dev ? a_lo_dev : a_hi_dev, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K,
                                (uint64_t)a_rows_alloc, (uint64_t)K * 2, G_BK, 128));
    B200_TRY(make_tensor_map_2d(&mb, b_dev, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (uint64_t)K, (uint64_t)N, (uint64_t)K * 2, G_BK, 128));
    GemmParams p = {};
    p.M = M; p.N = N; p.K = K; p.ldc = ldc;
    p.tiles_m = (M + G_BM - 1) / G_BM;
    p.tiles_n = (N + G_BN - 1) / G_BN;
    p.total_tiles = p.tiles_m * p.tiles_n;
    p.kblocks_per_pass = K / G_BK;
    p.passes = a_lo_dev ? 2 : 1;
    p.c = c_dev;
    const size_t smem = (size_t)G_STAGES * G_STAGE_BYTES + (2 * G_STAGES + 4) * 8 + 16 + 1024;
    B200_CUDA(cudaFuncSetAttribute(gemm_bf16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int grid = p.total_tiles < sms ? p.total_tiles : sms;
    gemm_bf16_tn_kernel<<<grid, G_THREADS, smem, (cudaStream_t)stream>>>(ma_hi, ma_lo, mb, p);
    B200_CUDA(cudaGetLastError());
    return 0;
}
