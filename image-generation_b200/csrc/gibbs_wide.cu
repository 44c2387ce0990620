// Throughput form of the bit-packed sweep kernel: 28 chains per lane, one CTA per SM, CTA size T and table width W
// known at compile time (Pegasus: W = 15, Zephyr: W = 20), tables through the 2-stage bulk-copy ring.
//
// Same arithmetic, same tables, same shared-memory layout and same results as gibbs_kernel<28, MODE> (gibbs.cu) -- this is
// the instantiation the headline configuration runs (BASELINE.json configs[1]: Pegasus P16, 4096 chains = 147 groups of
// 28 = one CTA per SM; reference call sites src/model_wrapper.py:309-316, src/utils/common.py:123-138), rebuilt around what the
// round-2 profile of the generic kernel showed (profiles/r2_gibbs_ncu_summary.txt, 40.7 warp instructions per 32
// updates at 76 % issue utilisation):
//   * 12 % of the executed instructions were bookkeeping of the run-time geometry -- the mode flags, the 64-bit round
//     counter, table-row address arithmetic and register moves of the rotating 7-slot loop.  With T and W as template
//     parameters the neighbour slots are fully unrolled, every table entry is one LDS.64 at an immediate offset from a
//     per-round base and the round loop carries a 32-bit counter.
//   * 0.57 warps per issue slot sat in the round's __syncthreads() while the Philox draw of the next round -- a quarter
//     of the kernel's time, and independent of the state -- waited behind it.  Here the round barrier is an mbarrier
//     used in two halves: a warp ARRIVES when its spins of round q are written, draws the uniforms of its round-q+1
//     lane-tasks into a per-thread shared-memory slot (4 calls x 16 bytes), and only then WAITS for the stragglers.
//
// Measured and dropped on top of this version (P16, 4096 chains x 1000 sweeps, 28.2-28.3 ms; source-level ncu shows the
// neighbour / decision / draw phases at 5.1-5.4 cycles per warp instruction with five warps per scheduler, i.e. at the
// issue limit, the remaining 20 % being the ~46-instruction round prologue at 9 cycles per instruction and the wait for
// the round's slowest warp):
//   * neighbour slots fetched 4, 5 or 14 at a time instead of 7: 28.15-28.19 ms, no difference;
//   * two rounds per loop iteration so that the tile stage (hence every table address and barrier parity) is a
//     compile-time constant: the executed loop body doubles to ~35 KB of SASS, past the instruction cache's comfortable
//     size: 28.56 ms;
//   * round geometry carried in registers, parities toggled instead of derived from the round counter (fewer prologue
//     instructions): 28.29 ms on P16, Z15 36.3 -> 36.6 ms -- the prologue's cost is latency after the barrier, not
//     instruction count;
//   * the bracket width -(K2 + K1 g) of two chains in one packed FFMA2 followed by one FADD per chain (1.5 issue slots
//     per decision instead of 2; 16 SASS instructions fewer per lane-task): P16 28.35 -> 28.58 ms, Z15 36.4 -> 36.1 ms,
//     i.e. inside the box-to-box spread on Zephyr and a loss on Pegasus, where the fma pipe (two passes per packed
//     op) is the busier resource.
#include "gibbs_packed.cuh"

namespace b200grbm {

constexpr int WIDE_CPL = 28;
#ifndef B200_WIDE_PACK2
#define B200_WIDE_PACK2 1
#endif
constexpr bool WIDE_PACK2 = B200_WIDE_PACK2 != 0;
constexpr int WIDE_CALLS = 4;           // Philox calls per lane-task: (28 + shift) / 8 rounded up, shift in {0, 4}

//
// SINGLE = the two-CTAs-per-SM form (Zephyr Z15: 2 x 384 threads): ONE tile stage per CTA, refilled at the top of each
// round while the SM's other CTA computes; that other CTA also fills this one's barrier waits, so the round barrier
// stays a __syncthreads() and the uniforms are drawn in place (no room for the slots next to two CTAs' tables anyway).
// (Tried: the tile copy split into three chunks with a barrier each -- rows of slot 0 and the first batch, then one
// chunk per further batch -- so that a round starts after a third of the copy: 36.3 -> 39.3 ms on the Z15 shard, the
// extra waits and smaller copies cost more than the earlier start saves.)
//
// TC = 0: the CTA size is a run-time value (any Pegasus- / Zephyr-width graph whose planned CTA size has no instantiation
// of its own); table rows are then one multiply apart instead of an immediate offset.
//
// PD = false on the ring (one CTA per SM): the form for CTAs whose pre-drawn slots do not fit behind two tile stages
// (Zephyr Z12, the Advantage2 fabric: 608 threads x 21 rows x 2 stages = 204 KB) -- __syncthreads() round barrier and
// uniforms drawn in place, but still the unrolled, immediate-offset neighbour loop and the packed decisions.
template <int MODE, int TC, int W, bool SINGLE, bool PD = !SINGLE>
__global__ void __launch_bounds__(TC > 0 ? TC : (SINGLE ? 384 : 768), SINGLE ? 2 : 1)
    gibbs_wide_kernel(const __grid_constant__ SweepParams p)
{
    static_assert(!(SINGLE && PD), "pre-drawn uniforms need the one-CTA-per-SM ring form");
    const int T = TC > 0 ? TC : (int)blockDim.x;
    constexpr int CPL = WIDE_CPL;
    const uint32_t TILE_BYTES = (uint32_t)(W + 1) * (uint32_t)T * 8u;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);                 // [0], [1]: tile stages; [2]: round barrier
    int2 *tinfo = reinterpret_cast<int2 *>(smem_raw + 128);
    uint32_t *Wst = reinterpret_cast<uint32_t *>(smem_raw + 128 + p.info_bytes);
    unsigned char *stage0 = smem_raw + 128 + p.info_bytes + p.state_bytes;

    const int tid = threadIdx.x;
    const int g = blockIdx.x;
    const int chain0 = g * CPL;
    const int nvalid = min(CPL, p.chains - chain0);
    const uint32_t dense_mask = (1u << nvalid) - 1u;
    const uint32_t blk0 = p.chain_block0 + (uint32_t)(chain0 >> 2);
    const uint32_t blk8 = blk0 >> 1;
    const bool shift4 = (blk0 & 1u) != 0;
    const uint32_t bar_addr = smem_u32(bars);
    const uint32_t stage_addr = smem_u32(stage0);
    const uint32_t n_tiles = (uint32_t)p.n_tiles;
    const uint32_t total = (uint32_t)p.num_sweeps * n_tiles;                  // launcher: < 2^31
    uint4 *drawn = reinterpret_cast<uint4 *>(smem_raw + p.drawn_offset) + tid;   // [call][T]

    if (tid == 0) {
        mbar_init(bar_addr, 1);
        mbar_init(bar_addr + 8, 1);
#ifdef B200_WIDE_ARRIVE_ALL     // sanitizer build: every thread arrives itself (racecheck does not follow __syncwarp + lane 0)
        if (PD) mbar_init(bar_addr + 16, T);
#else
        if (PD) mbar_init(bar_addr + 16, T / 32);         // round barrier: one arrival per warp
#endif
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (total > 0) {
            mbar_expect_tx(bar_addr, TILE_BYTES);
            bulk_g2s(stage_addr, p.tiles, TILE_BYTES, bar_addr);
        }
    }
    for (int k = tid; k < p.n_tiles; k += T) tinfo[k] = __ldg(p.tile_info + k);
    load_group_state<CPL>(p, Wst, tid, T, g, chain0, nvalid, dense_mask, blk0);
    __syncthreads();

    // uniforms of one round for this thread's lane-task: its own slot, nobody else reads it
    const auto draw_round = [&](uint32_t tile_next, uint32_t sweep_next) {
        const int2 inf = tinfo[tile_next];
        if (tid < inf.y) {
#pragma unroll
            for (int call = 0; call < WIDE_CALLS; ++call) {
                uint32_t r[4];
                philox4x32((uint32_t)(inf.x + tid), blk8 + call, sweep_next, B200GRBM_STREAM_SWEEP, p, r);
                drawn[call * T] = make_uint4(r[0], r[1], r[2], r[3]);
            }
        }
    };
    if (PD && total > 0) draw_round(0u, p.sweep_offset);

    Pair2Consts k2;
    k2.one2 = pack2(1.0f, 1.0f);
    k2.mone2 = pack2(-1.0f, -1.0f);
    k2.negc2 = pack2(-(128.0f - 0x1.0p-17f), -(128.0f - 0x1.0p-17f));
    k2.hi43 = p.hi43;                      // 0x43000000 from the parameter block: a register operand for PRMT

    uint32_t q = 0;
    for (int t = 0; t < p.num_sweeps; ++t) {
        const float coef = __ldg(p.coef + t);
        k2.coef2 = pack2(coef, coef);
        const uint32_t sweep = p.sweep_offset + (uint32_t)t;
#pragma unroll 1
        for (uint32_t tile = 0; tile < n_tiles; ++tile, ++q) {
            const uint32_t s = q & 1u;
            const bool more = q + 1 < total;
            const uint32_t tile_next = tile + 1 == n_tiles ? 0u : tile + 1;
            // stage s^1 was last read in round q-1; this thread has seen round q-1's barrier complete
            if (SINGLE) {
                if (tid == 0 && q > 0) {
                    mbar_expect_tx(bar_addr, TILE_BYTES);
                    bulk_g2s(stage_addr, reinterpret_cast<const unsigned char *>(p.tiles) + (size_t)tile * TILE_BYTES, TILE_BYTES,
                             bar_addr);
                }
            } else if (tid == 0 && more) {
                const uint32_t nb = bar_addr + 8u * (s ^ 1u);
                mbar_expect_tx(nb, TILE_BYTES);
                bulk_g2s(stage_addr + (s ^ 1u) * TILE_BYTES,
                         reinterpret_cast<const unsigned char *>(p.tiles) + (size_t)tile_next * TILE_BYTES, TILE_BYTES, nb);
            }
            const int2 info = tinfo[tile];
            if (SINGLE) mbar_wait(bar_addr, q & 1u);
            else mbar_wait(bar_addr + 8u * s, (q >> 1) & 1u);

            if (tid < info.y) {
                const int pp = info.x + tid;
                // tile rows: 0 = f0, 1 .. W = neighbour slots {2J bits, byte offset of the neighbour's state word}
                const uint2 *ep = reinterpret_cast<const uint2 *>(stage0 + (SINGLE ? 0u : s * TILE_BYTES)) + tid;
                float f[CPL];
                {
                    const float fz = u2f(ep[0].x);
                    const uint2 e0 = ep[T];
                    init_slot<CPL>(f, lds_word(smem_raw, e0.y), fz, u2f(e0.x));
                }
#pragma unroll
                for (int k0 = 1; k0 < W; k0 += 7) {
                    uint2 e[7];
                    uint32_t w[7];
#pragma unroll
                    for (int i = 0; i < 7; ++i)
                        if (k0 + i < W) e[i] = ep[(k0 + 1 + i) * T];
#pragma unroll
                    for (int i = 0; i < 7; ++i)
                        if (k0 + i < W) w[i] = lds_word(smem_raw, e[i].y);
#pragma unroll
                    for (int i = 0; i < 7; ++i)
                        if (k0 + i < W) add_slot<CPL>(f, w[i], u2f(e[i].x));
                }
                const uint32_t R0[2][4] = {};
                Wst[pp] = shift4 ? decide_word<CPL, MODE, 4, false, PD, WIDE_PACK2>(f, coef, (uint32_t)pp, sweep, blk8, p, t, chain0, R0, drawn, T, &k2)
                                 : decide_word<CPL, MODE, 0, false, PD, WIDE_PACK2>(f, coef, (uint32_t)pp, sweep, blk8, p, t, chain0, R0, drawn, T, &k2);
            }
            // split round barrier: arrive (release: this warp's words are visible), draw, wait (acquire)
            if constexpr (PD) {
#ifdef B200_WIDE_ARRIVE_ALL
                mbar_arrive(bar_addr + 16);
#else
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bar_addr + 16);
#endif
                if (more) draw_round(tile_next, tile + 1 == n_tiles ? sweep + 1u : sweep);
                mbar_wait(bar_addr + 16, q & 1u);
            } else {
                __syncthreads();
            }
        }
    }

    store_group_state<CPL>(p, Wst, tid, T, g, chain0, nvalid, dense_mask);
}

typedef void (*wide_fn)(const SweepParams);

template <int T, int W, bool SINGLE, bool PD = !SINGLE>      // T = 0: run-time CTA size
static wide_fn wide_pick_mode(int mode)
{
    return mode == MODE_PHILOX_FAST ? gibbs_wide_kernel<MODE_PHILOX_FAST, T, W, SINGLE, PD>
                                    : gibbs_wide_kernel<MODE_PHILOX_EXACT, T, W, SINGLE, PD>;
}

static wide_fn wide_pick(int mode, int threads, int width, bool single, bool predraw = true)
{
    if (mode != MODE_PHILOX_EXACT && mode != MODE_PHILOX_FAST) return nullptr;
    if (!single && !predraw)            // ring without room for the slots
        return width == 15 ? wide_pick_mode<0, 15, false, false>(mode) : width == 20 ? wide_pick_mode<0, 20, false, false>(mode) : nullptr;
    if (!single && threads == 640 && width == 15) return wide_pick_mode<640, 15, false>(mode);   // Pegasus P16: 9 rounds of 640 lanes
    if (single && threads == 384 && width == 20) return wide_pick_mode<384, 20, true>(mode);     // Zephyr Z15: 20 rounds, 2 CTAs per SM
    if (!single && threads == 480 && width == 20) return wide_pick_mode<480, 20, false>(mode);   // Zephyr Z15, < 296 groups: 16 rounds
    // any other CTA size on a Pegasus-width (15) or Zephyr-width (20) table: run-time T
    if (width == 15) return single ? (threads <= 384 ? wide_pick_mode<0, 15, true>(mode) : nullptr) : wide_pick_mode<0, 15, false>(mode);
    if (width == 20) return single ? (threads <= 384 ? wide_pick_mode<0, 20, true>(mode) : nullptr) : wide_pick_mode<0, 20, false>(mode);
    return nullptr;
}

// shared-memory bytes of the specialised kernel (0: no instantiation for this geometry): the generic layout, plus -- in
// the one-CTA-per-SM form -- the pre-drawn words behind it
size_t wide_kernel_smem(int cpl, int mode, int threads, int width, bool single, size_t ring_smem, size_t smem_limit,
                        uint32_t *drawn_offset)
{
    if (cpl != WIDE_CPL || wide_pick(mode, threads, width, single) == nullptr) return 0;
    *drawn_offset = 0;                      // 0 = no pre-drawn slots
    if (single) return ring_smem;           // the launcher's one-stage size; nothing behind it
    const size_t off = (ring_smem + 127) / 128 * 128;
    if (off + (size_t)WIDE_CALLS * 16 * threads > smem_limit) return ring_smem;      // ring, uniforms drawn in place
    *drawn_offset = (uint32_t)off;
    return off + (size_t)WIDE_CALLS * 16 * threads;
}

int32_t launch_gibbs_wide(const SweepParams &p, int mode, int threads, int groups, size_t smem, cudaStream_t st)
{
    wide_fn fn = wide_pick(mode, threads, p.width, p.single != 0, p.drawn_offset != 0);
    if (fn == nullptr) return fail(B200GRBM_EUNSUPPORTED, "gibbs_wide: no instantiation for threads=%d width=%d", threads, p.width);
    B200_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    fn<<<groups, threads, smem, st>>>(p);
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200grbm
