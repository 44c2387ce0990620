// CTA-pair (tcgen05 cta_group::2) version of the int8 MMD Gram forward (mmd_tc.cu): two SMs of
// a TPC share one 256 x 256 accumulator tile.  Each CTA stages only its own 128 rows of A and
// HALF of the B rows (128 of 256) per k-block -- 32 KB instead of 48 KB per 128 x 256 x 128 MACs --
// which is what the single-CTA kernel is short of (its tensor pipe sits at ~60 % waiting on L2).
//
//   CTA rank r of the pair, pair-tile (I, J), J >= I (upper triangle of 256 x 256 tiles):
//     A rows [256 I + 128 r, +128)   B rows [256 J + 128 r, +128)
//     accumulator: rows 256 I + 128 r + lane, all 256 columns, in the CTA's own TMEM
//   leader (rank 0) issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 32); both CTAs run
//   their own TMA producer and epilogue.  Barriers: full[s] lives in the leader (2 arrivals +
//   both CTAs' transaction bytes), empty[s] / tmem_full[a] are arrived in both CTAs by the
//   multicast tcgen05.commit, tmem_empty[a] lives in the leader (2 x 8 epilogue warps).
#include "mmd_hist.cuh"

namespace b200grbm {

constexpr int P_BK = 128;                 // bytes per k-block (one 128B swizzle row of int8)
constexpr int P_A_BYTES = 128 * P_BK;     // 16 KB: this CTA's 128 rows of A
constexpr int P_B_BYTES = 128 * P_BK;     // 16 KB: this CTA's half of B
constexpr int P_STAGE_BYTES = P_A_BYTES + P_B_BYTES;
constexpr int P_THREADS = 320, P_EPI_WARPS = 8;

struct PairParams {
    int m_x, m, d;
    int n_kblocks;
    int tiles, total_tiles;   // 256 x 256 pair-tiles per side; upper triangle count
    int stages;
    int shard_rank, shard_world;          // this launch contracts pair-tiles t = shard_rank (mod shard_world)
    unsigned long long *hist;             // [3][d + 1] ordered-pair counts per Hamming distance (mmd_hist.cuh)
};

// upper triangle enumerated row by row: pair-tile t -> (I, J) with J >= I
__device__ __forceinline__ void pair_tile_coords(int tiles, int t, int &I, int &J)
{
    // row I starts at offset I * tiles - I (I - 1) / 2
    int i = (int)((2.0f * tiles + 1.0f - sqrtf((2.0f * tiles + 1.0f) * (2.0f * tiles + 1.0f) - 8.0f * (float)t)) * 0.5f);
    if (i < 0) i = 0;
    if (i > tiles - 1) i = tiles - 1;
    while (i > 0 && i * tiles - i * (i - 1) / 2 > t) --i;
    while ((i + 1) * tiles - (i + 1) * i / 2 <= t) ++i;
    I = i;
    J = i + (t - (i * tiles - i * (i - 1) / 2));
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
    mmd_gram_i8_2cta_kernel(const __grid_constant__ CUtensorMap tmap, const PairParams p)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int S = p.stages;
    uint32_t *bins = reinterpret_cast<uint32_t *>(smem + (size_t)S * P_STAGE_BYTES);     // d + 1 counters
    const size_t lut_bytes = ((size_t)(p.d + 1) * 4 + 15) / 16 * 16;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)S * P_STAGE_BYTES + lut_bytes);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 4);
    const uint32_t full0 = smem_addr(bars), empty0 = full0 + 8u * S, tfull0 = empty0 + 8u * S, tempty0 = tfull0 + 16u;
    const int n_local = p.total_tiles > p.shard_rank ? (p.total_tiles - p.shard_rank + p.shard_world - 1) / p.shard_world : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { bar_init(full0 + 8u * s, 2); bar_init(empty0 + 8u * s, 1); }
        for (int a = 0; a < 2; ++a) { bar_init(tfull0 + 8u * a, 1); bar_init(tempty0 + 8u * a, 2 * P_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    if (warp == 1) tmem_alloc_2cta(smem_addr(tmem_slot), 512);
    for (int k = threadIdx.x; k <= p.d; k += blockDim.x) bins[k] = 0u;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                      // the peer's barriers are initialised before anyone arrives remotely
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int u = pair; u < n_local; u += n_pairs) {
                int I, J;
                pair_tile_coords(p.tiles, u * p.shard_world + p.shard_rank, I, J);
                const int a_row = I * 256 + (int)rank * 128, b_row = J * 256 + (int)rank * 128;
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    bar_wait(empty0 + 8u * s, ph ^ 1u);                       // own stage free (multicast commit)
                    const uint32_t lead_full = mapa_cluster(full0 + 8u * s, 0);
                    const uint32_t dst = smem_addr(smem + (size_t)s * P_STAGE_BYTES);
                    if (leader) bar_expect_tx(full0 + 8u * s, 2 * P_STAGE_BYTES);  // both CTAs' bytes land on this barrier
                    else bar_arrive_cluster(lead_full);
                    tma_load_2d_2sm(dst, &tmap, kb * P_BK, a_row, lead_full);
                    tma_load_2d_2sm(dst + P_A_BYTES, &tmap, kb * P_BK, b_row, lead_full);
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && lane == 0) {
            const uint32_t idesc = umma_idesc_i8(256, 256);
            int s = 0, it = 0;
            uint32_t ph = 0;
            for (int u = pair; u < n_local; u += n_pairs, ++it) {
                const uint32_t acc = (uint32_t)it & 1u, acc_ph = ((uint32_t)it >> 1) & 1u;
                bar_wait(tempty0 + 8u * acc, acc_ph ^ 1u);               // both epilogues drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    bar_wait(full0 + 8u * s, ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_addr(smem + (size_t)s * P_STAGE_BYTES);
                    const uint64_t adesc = umma_desc_sw128(a_addr), bdesc = umma_desc_sw128(a_addr + P_A_BYTES);
#pragma unroll
                    for (int k = 0; k < P_BK / 32; ++k)
                        umma_i8_2cta(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_2cta(empty0 + 8u * s);                   // frees the stage in both CTAs
                    if (++s == S) { s = 0; ph ^= 1u; }
                }
                umma_commit_2cta(tfull0 + 8u * acc);                     // accumulator ready in both CTAs
            }
        }
    } else {
        // ===================== epilogue (both CTAs): TMEM -> Hamming histograms =====================
        const int ew = warp - 2, quarter = warp & 3, half = ew >> 2;
        const int epi_tid = threadIdx.x - 64;
        const int two_d = 2 * p.d;
        const uint32_t lead_tempty0 = mapa_cluster(tempty0, 0);
        HistAccumulator hacc = {bins, p.hist, p.d, -1};
        int it = 0;
        for (int u = pair; u < n_local; u += n_pairs, ++it) {
            int I, J;
            pair_tile_coords(p.tiles, u * p.shard_world + p.shard_rank, I, J);
            const uint32_t acc = (uint32_t)it & 1u, acc_ph = ((uint32_t)it >> 1) & 1u;
            const int row0 = I * 256 + (int)rank * 128, col0 = J * 256;
            const int row = row0 + quarter * 32 + lane;
            const bool strict_upper = J > I;
            const bool rows_x = row0 + 128 <= p.m_x, rows_y = row0 >= p.m_x;
            const bool cols_x = col0 + 256 <= p.m_x, cols_y = col0 >= p.m_x;
            const int mode = tile_mode(strict_upper, (row0 + 128 <= p.m) && (col0 + 256 <= p.m), rows_x, rows_y, cols_x, cols_y);
            if (mode != TILE_MIXED) {
                const int type = rows_x && cols_x ? HIST_XX : (rows_y && cols_y ? HIST_YY : HIST_XY);
                if (type != hacc.type) {             // uniform over this CTA's epilogue threads
                    hacc.flush(epi_tid, P_EPI_WARPS * 32);
                    hacc.type = type;
                }
            }
            bar_wait(tfull0 + 8u * acc, acc_ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int chunk = 0; chunk < 4; ++chunk) {
                uint32_t v[32];
                const int cbase = half * 128 + chunk * 32;
                __syncwarp();
                tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * 256 + (uint32_t)cbase, v);
                hist_count_chunk(v, mode, hacc, two_d, row, col0 + cbase, p.m_x, p.m);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive_cluster(lead_tempty0 + 8u * acc);   // the leader's MMA thread waits for both CTAs
        }
        hacc.flush(epi_tid, P_EPI_WARPS * 32);
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                      // neither CTA frees TMEM / exits while the peer may still touch it
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

// launched from b200grbm_mmd_forward_i8 (mmd_tc.cu)
int32_t launch_gram_i8_2cta(const CUtensorMap &tmap, int m_x, int m, int d, int d_pad, unsigned long long *hist,
                            int shard_rank, int shard_world, cudaStream_t st)
{
    int dev = 0, smem_optin = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const size_t lut_bytes = ((size_t)(d + 1) * 4 + 15) / 16 * 16;
    int stages = 6;
    size_t smem = 0;
    for (; stages >= 2; --stages) {
        smem = (size_t)stages * P_STAGE_BYTES + lut_bytes + (2 * stages + 4) * 8 + 16 + 1024;
        if (smem <= (size_t)smem_optin) break;
    }
    if (stages < 2) return fail(B200GRBM_EUNSUPPORTED, "mmd (2-CTA): d=%d look-up table does not fit shared memory", d);
    PairParams p = {};
    p.m_x = m_x; p.m = m; p.d = d;
    p.n_kblocks = (d_pad + P_BK - 1) / P_BK;
    p.tiles = (m + 255) / 256;
    p.total_tiles = p.tiles * (p.tiles + 1) / 2;
    p.stages = stages;
    p.shard_rank = shard_rank; p.shard_world = shard_world;
    p.hist = hist;
    B200_CUDA(cudaFuncSetAttribute(mmd_gram_i8_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int sms = sm_count() > 0 ? sm_count() : 148;
    int pairs = sms / 2;
    const int n_local = p.total_tiles > shard_rank ? (p.total_tiles - shard_rank + shard_world - 1) / shard_world : 0;
    if (n_local == 0) return 0;
    if (pairs > n_local) pairs = n_local;
    mmd_gram_i8_2cta_kernel<<<2 * pairs, P_THREADS, smem, st>>>(tmap, p);     // __cluster_dims__(2,1,1)
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200grbm
