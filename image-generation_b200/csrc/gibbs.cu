// Colour-blocked heat-bath (Gibbs / annealing) sweeps for Ising graphs on sm_100a.
//
// Replaces the QPU / classical-annealer call  sampler.sample_ising(h, J, num_reads, ...)
// that GraphRestrictedBoltzmannMachine.sample makes (reference call sites
// src/model_wrapper.py:309-316, :369-376, src/utils/persistent_qpu_sampler.py:71-78;
// sampler built at src/utils/common.py:123-138).  Numerical contract:
// include/b200grbm_spec.h.
//
// Layout (why this is not "one int8 per spin"):
//   * One CTA owns a *group* of CPL <= 32 chains for the whole launch.  The group's state
//     is bit-packed: word W[p] in shared memory holds spin p of all CPL chains.  P16:
//     5640 words = 22.5 KB; Z15: 29.8 KB.
//   * Lanes are spins of the colour block being updated; each lane carries CPL fp32 local
//     fields in registers.  For neighbour slot k the lane reads ONE (2J, nbr) entry and ONE
//     state word, moves the word's bits into predicates (R2P, 7 per instruction) and does
//     CPL predicated adds  f[c] += 2J  (f starts at f0 = h - sum J).  Table and state
//     traffic is amortised over CPL chains; the kernel is bound by instruction issue
//     (one FADD per neighbour-chain pair, Philox + acceptance per update), not by
//     shared-memory or HBM bytes.  HBM sees one read and one write of the state per launch.
//     CPL = 28 keeps bit 7 of every byte free so four R2Ps cover the word exactly.
//   * The (f0; 2J, nbr) tables are tiled per colour round -- tile = (1 + width) x threads
//     8-byte entries, contiguous in global memory -- and streamed through a 2-stage
//     shared-memory ring by the bulk-copy engine (cp.async.bulk + mbarrier complete_tx,
//     SASS UBLKCP): the copy of round q+1 overlaps the arithmetic of round q, so no warp
//     ever waits on an L2 load inside the neighbour loop.  Two more modes, picked by the
//     launcher: all tiles resident (small graphs), and one stage with two CTAs per SM.
//   * Same-colour spins are never adjacent, so the parallel colour step equals the
//     sequential sweep in visit order; W[p] is written in place and __syncthreads()
//     separates rounds.
//   * Uniforms: Philox4x32-10 keyed by (seed; visit position, global chain / 8, sweep) --
//     one call yields the high 16 bits of the uniforms of 8 chains of the lane; the low 7
//     bits live in a second stream that is only evaluated when a decision depends on them --
//     so trajectories do not depend on CPL, CTA size, grid or GPU count.  The chain block
//     sits in counter word 1, which enters the first round only through an XOR: the calls
//     of one lane-task (same position and sweep, consecutive blocks) share the multiplies
//     of rounds 1-3 that do not depend on it.  Round keys are precomputed on the host into
//     the kernel parameter block (uniform-register operands).
//   * Lazy exact acceptance: the contract decision  fmaf(v, exp2_poly(x), v) < 1  is first
//     bracketed with MUFU.EX2 and the 16-bit midpoint uniform; only when the bracket
//     (2^-17 (1 + e) on v, 2^-20 relative on e) contains the threshold -- about 2e-5 of the
//     decisions -- is that decision redone with the polynomial and the full 23-bit uniform.
//     The result is bit-identical to evaluating the contract everywhere (oracle/oracle.c).
#include "gibbs_packed.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace b200grbm {

// MAXT = largest CTA this instantiation may be launched with.  768 caps ptxas at 80 registers per thread (and is the
// one two 384-thread CTAs per SM need); the 28-chain kernels also exist for MAXT = 640 -- 96 registers, a looser
// schedule of the acceptance phase: 32.8 -> 32.0 ms on P16 (9 rounds of 640 lanes).
//
// MG = several chain groups per CTA (resident tables only; see below).  A template parameter because the group barriers
// are named barriers with a run-time id: such a kernel reserves all 16 barriers of a CTA, which would cap the SM at two
// CTAs in the modes that want more.
template <int CPL, int MODE, int MAXT = 768, bool MG = false>
__global__ void __launch_bounds__(MAXT, 1) gibbs_kernel(const __grid_constant__ SweepParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);                 // 2 mbarriers (16 B, padded to 128)
    int2 *tinfo = reinterpret_cast<int2 *>(smem_raw + 128);                  // round table, copied once
    // Small graphs with many chain groups (the reference's 256-latent model scaled up in chains, BASELINE configs[4]):
    // gpc > 1 groups share one CTA and ONE resident copy of the tables.  A group is tpg threads (a "sub-CTA" with its
    // own state words and its own named barrier); with one group per CTA the 54 KB of tables per CTA cap the SM at
    // four two-warp CTAs -- two warps per scheduler.
    const int gi = MG ? (int)threadIdx.x / p.tpg : 0;                       // group within the CTA
    const unsigned char *sbase = smem_raw + (size_t)gi * p.state_bytes;      // the tiles' .nbr offsets are relative to this
    uint32_t *W = reinterpret_cast<uint32_t *>(smem_raw + 128 + p.info_bytes + (size_t)gi * p.state_bytes);
    unsigned char *stage0 = smem_raw + 128 + p.info_bytes + (size_t)(MG ? p.gpc : 1) * p.state_bytes;

    const int tid = MG ? (int)threadIdx.x - gi * p.tpg : (int)threadIdx.x;
    const int nthr = MG ? p.tpg : (int)blockDim.x;
    const int g = MG ? (int)blockIdx.x * p.gpc + gi : (int)blockIdx.x;
    const int chain0 = g * CPL;  // first chain of this group, local to the call
    const int nvalid = max(0, min(CPL, p.chains - chain0));                  // 0: no such group (last CTA)
    const uint32_t dense_mask = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
    const uint32_t blk0 = p.chain_block0 + (uint32_t)(chain0 >> 2);   // global chain / 4 (initial-state stream)
    const uint32_t blk8 = blk0 >> 1;                                   // global chain / 8 (sweep streams)
    const bool shift4 = (blk0 & 1u) != 0;                              // group starts in the middle of a block of 8
    const uint32_t bar_addr = smem_u32(bars);
    const uint32_t stage_addr = smem_u32(stage0);
    const long long total_tiles = (long long)p.num_sweeps * p.n_tiles;

    if (threadIdx.x == 0) {
        mbar_init(bar_addr, 1);
        mbar_init(bar_addr + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (total_tiles > 0 && p.resident) {   // small graph: every round's tile, once
            mbar_expect_tx(bar_addr, p.tile_bytes * (uint32_t)p.n_tiles);
            for (int k = 0; k < p.n_tiles; ++k)
                bulk_g2s(stage_addr + (uint32_t)k * p.tile_bytes,
                         reinterpret_cast<const unsigned char *>(p.tiles) + (size_t)k * p.tile_bytes, p.tile_bytes, bar_addr);
        } else if (total_tiles > 0) {          // prologue: round 0 into stage 0
            mbar_expect_tx(bar_addr, p.tile_bytes);
            bulk_g2s(stage_addr, p.tiles, p.tile_bytes, bar_addr);
        }
    }
    for (int k = threadIdx.x; k < p.n_tiles; k += blockDim.x) tinfo[k] = __ldg(p.tile_info + k);

    if (nvalid > 0) load_group_state<CPL>(p, W, tid, nthr, g, chain0, nvalid, dense_mask, blk0);
    __syncthreads();
    if (nvalid == 0) return;               // padding group of the last CTA: no CTA-wide barrier after this point

    // throughput groups (>= 16 chains per lane, Philox modes): decisions two chains at a time in packed fp32x2
    constexpr bool PACK2 = CPL >= 16 && MODE != MODE_SUPPLIED_EXACT;
    Pair2Consts k2;
    k2.one2 = pack2(1.0f, 1.0f);
    k2.mone2 = pack2(-1.0f, -1.0f);
    k2.negc2 = pack2(-(128.0f - 0x1.0p-17f), -(128.0f - 0x1.0p-17f));
    k2.hi43 = p.hi43;

    int tile = 0, t = 0;
    float coef = total_tiles > 0 ? __ldg(p.coef) : 0.f;
    float coef_next = p.num_sweeps > 1 ? __ldg(p.coef + 1) : 0.f;   // one sweep ahead: never waited on
    for (long long q = 0; q < total_tiles; ++q) {
        const uint32_t s = (uint32_t)q & 1u;
        // stage s^1 was last read in round q-1, which every thread left through the
        // __syncthreads() below -> safe to refill it now while round q computes
        if (tid == 0 && !p.resident) {
            if (p.single) {
                // one stage: the copy of round q can only start now (everyone has left round q-1); the SM's other
                // CTA computes meanwhile
                if (q > 0) {
                    mbar_expect_tx(bar_addr, p.tile_bytes);
                    bulk_g2s(stage_addr, reinterpret_cast<const unsigned char *>(p.tiles) + (size_t)tile * p.tile_bytes,
                             p.tile_bytes, bar_addr);
                }
            } else if (q + 1 < total_tiles) {
                const int next_tile = tile + 1 == p.n_tiles ? 0 : tile + 1;
                const uint32_t nb = bar_addr + 8u * (s ^ 1u);
                mbar_expect_tx(nb, p.tile_bytes);
                bulk_g2s(stage_addr + (s ^ 1u) * p.tile_bytes,
                         reinterpret_cast<const unsigned char *>(p.tiles) + (size_t)next_tile * p.tile_bytes, p.tile_bytes,
                         nb);
            }
        }
        const int2 info = tinfo[tile];
        // resident tables: one wait, before the first round (a copy round trip per round would otherwise be the
        // critical path of the short rounds of a small graph)
        if (p.resident) {
            if (q == 0) mbar_wait(bar_addr, 0u);
        } else if (p.single) {
            mbar_wait(bar_addr, (uint32_t)q & 1u);
        } else {
            mbar_wait(bar_addr + 8u * s, (uint32_t)(q >> 1) & 1u);
        }
        const uint32_t stage_idx = p.resident ? (uint32_t)tile : (p.single ? 0u : s);

        if (tid < info.y) {
            const int pp = info.x + tid;
            const uint32_t sweep = p.sweep_offset + (uint32_t)t;
            k2.coef2 = pack2(coef, coef);
            // tile rows: 0 = f0, 1 .. width = neighbour slots {2J bits, byte offset of the neighbour's state word}
            const uint2 *ep = reinterpret_cast<const uint2 *>(stage0 + stage_idx * p.tile_bytes) + tid;
            float f[CPL];
            const float fz = u2f(ep->x);
            ep += nthr;

            // Small groups (few chains spread over many CTAs, e.g. the reference's 256 reads) are latency-bound: a
            // round is one dependent chain  table entry -> state word -> adds -> Philox -> decision -> barrier.
            // The Philox words do not depend on the state, so they are drawn first; the slots are fetched eight
            // at a time (two shared-memory latencies per eight slots).
            constexpr bool PRE = SlotUnroll<CPL>::value == 4 && MODE != MODE_SUPPLIED_EXACT;
            uint32_t R[2][4];
            if constexpr (PRE) {
                philox4x32((uint32_t)pp, blk8, sweep, B200GRBM_STREAM_SWEEP, p, R[0]);
                if ((CPL + 4 + 7) / 8 > 1 && shift4) philox4x32((uint32_t)pp, blk8 + 1, sweep, B200GRBM_STREAM_SWEEP, p, R[1]);
            }
            if (SlotUnroll<CPL>::value == 4) {
                // width is a multiple of 4 here (padding slots hold 2J = 0, nbr = own position)
#pragma unroll
                for (int c = 0; c < CPL; ++c) f[c] = fz;
                int k = 0;
#pragma unroll 1
                for (; k + 8 <= p.width; k += 8) {
                    uint2 e[8];
                    uint32_t w[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) e[i] = ep[i * nthr];
                    ep += 8 * nthr;
#pragma unroll
                    for (int i = 0; i < 8; ++i) w[i] = lds_word(sbase, e[i].y);
#pragma unroll
                    for (int i = 0; i < 8; ++i) add_slot<CPL>(f, w[i], u2f(e[i].x));
                }
                if (k < p.width) {
                    const uint2 e0 = ep[0], e1 = ep[nthr], e2 = ep[2 * nthr], e3 = ep[3 * nthr];
                    const uint32_t w0 = lds_word(sbase, e0.y), w1 = lds_word(sbase, e1.y),
                                   w2 = lds_word(sbase, e2.y), w3 = lds_word(sbase, e3.y);
                    add_slot<CPL>(f, w0, u2f(e0.x));
                    add_slot<CPL>(f, w1, u2f(e1.x));
                    add_slot<CPL>(f, w2, u2f(e2.x));
                    add_slot<CPL>(f, w3, u2f(e3.x));
                }
            } else {
                // slot 0 initialises, then seven (or two) slots per iteration: the entries, then the state words,
                // then 7 x CPL predicated adds.  No software pipeline across iterations -- the other warps of the
                // scheduler cover the two shared-memory latencies, and a rotating pipeline costs register moves.
                {
                    const uint2 e0 = *ep;
                    ep += nthr;
                    init_slot<CPL>(f, lds_word(sbase, e0.y), fz, u2f(e0.x));
                }
                int k = 1;
#ifdef B200_EXP_NOSLOTS      // timing experiment: acceptance phase only
                k = p.width;
#endif
                const auto seven_slots = [&]() {
                    uint2 e[7];
                    uint32_t w[7];
#pragma unroll
                    for (int i = 0; i < 7; ++i) e[i] = ep[i * nthr];
                    ep += 7 * nthr;
#pragma unroll
                    for (int i = 0; i < 7; ++i) w[i] = lds_word(sbase, e[i].y);
#pragma unroll
                    for (int i = 0; i < 7; ++i) add_slot<CPL>(f, w[i], u2f(e[i].x));
                    k += 7;
                };
#pragma unroll 1
                while (k + 7 <= p.width) seven_slots();       // Pegasus: 1 + 7 + 7 slots, Zephyr: 1 + 7 + 7 + 5
#pragma unroll 1
                for (; k + 1 < p.width; k += 2) {
                    const uint2 ea = ep[0], eb = ep[nthr];
                    ep += 2 * nthr;
                    const uint32_t wa = lds_word(sbase, ea.y), wb = lds_word(sbase, eb.y);
                    add_slot<CPL>(f, wa, u2f(ea.x));
                    add_slot<CPL>(f, wb, u2f(eb.x));
                }
                if (k < p.width) {
                    const uint2 ea = *ep;
                    add_slot<CPL>(f, lds_word(sbase, ea.y), u2f(ea.x));
                }
            }

#ifdef B200_EXP_NOACCEPT      // timing experiment: neighbour loop only
            {
                uint32_t acc = 0;
#pragma unroll
                for (int c = 0; c < CPL; ++c) acc ^= f2u(f[c]);
                W[pp] = acc & 0x7f7f7f7fu;
            }
#else
            W[pp] = shift4 ? decide_word<CPL, MODE, 4, PRE, false, PACK2>(f, coef, (uint32_t)pp, sweep, blk8, p, t, chain0, R, nullptr, 0, &k2)
                           : decide_word<CPL, MODE, 0, PRE, false, PACK2>(f, coef, (uint32_t)pp, sweep, blk8, p, t, chain0, R, nullptr, 0, &k2);
#endif
            // bits of chains beyond nvalid are masked at write-back
        }
#ifndef B200_EXP_NOBARRIER   // timing experiment only (tools/build_variant.sh): results are wrong without it
        if constexpr (MG) group_bar_sync(1 + gi, nthr);  // the group's own named barrier (ids 1 .. 15)
        else __syncthreads();
#endif
        if (++tile == p.n_tiles) {
            tile = 0;
            ++t;
            coef = coef_next;
            if (t + 1 < p.num_sweeps) coef_next = __ldg(p.coef + t + 1);
        }
    }

    store_group_state<CPL>(p, W, tid, nthr, g, chain0, nvalid, dense_mask);
}

// Packed final state -> int8 rows in node order (dimod's SampleSet.record.sample layout).  A kernel of its own: the
// sweep kernel writes only its bit-packed words (coalesced), and this one turns the words of a chain group into the
// group's rows -- one contiguous range of the output -- by scattering bytes into SHARED memory (the permutation
// order[] costs nothing there) and streaming the rows out with aligned 16-byte stores.  Doing the same inside the
// sweep kernel's tail (either as nvalid single-byte global stores per spin, the round-1 form, or staged through the
// idle tile buffers) changes ptxas's allocation for the sweep loop: 92 -> 93 registers and 1.8 % of the headline rate.
__global__ void __launch_bounds__(256) unpack_state_kernel(const uint32_t *__restrict__ packed, int chains, int cpl, int n, int n_pad,
                                                           const int32_t *__restrict__ order, int8_t *__restrict__ state_out,
                                                           int rows_per_pass)
{
    extern __shared__ __align__(16) int8_t rows_buf[];
    const int g = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int chain0 = g * cpl;
    const int nvalid = min(cpl, chains - chain0);
    const uint32_t *words = packed + (size_t)g * n_pad;
    for (int c0 = 0; c0 < nvalid; c0 += rows_per_pass) {
        const int rows = min(rows_per_pass, nvalid - c0);
        int8_t *gdst = state_out + (size_t)(chain0 + c0) * n;
        const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(gdst) & 15u);   // same alignment on both sides
        int8_t *buf = rows_buf + mis;
        __syncthreads();                                                 // previous pass streamed out
        for (int pp = tid; pp < n; pp += nthr) {
            const uint32_t w = __ldg(words + pp) >> c0;
            const int node = __ldg(order + pp);
            for (int c = 0; c < rows; ++c) buf[c * n + node] = (w >> c) & 1u ? (int8_t)1 : (int8_t)-1;
        }
        __syncthreads();
        const int total = rows * n;
        const int head = min(total, (int)((16u - mis) & 15u));
        const int body = (total - head) >> 4;
        for (int k = tid; k < head; k += nthr) gdst[k] = buf[k];
        const uint4 *s4 = reinterpret_cast<const uint4 *>(buf + head);
        uint4 *g4 = reinterpret_cast<uint4 *>(gdst + head);
        for (int k = tid; k < body; k += nthr) g4[k] = s4[k];
        for (int k = head + (body << 4) + tid; k < total; k += nthr) gdst[k] = buf[k];
    }
}

// Measurement aid: largest relative error of ex2.approx.ftz.f32 (the MUFU.EX2 the lazy acceptance brackets
// with) against double-precision exp2 over n evenly spaced fp32 arguments of [x_lo, x_hi].
__global__ void ex2_probe_kernel(float x_lo, float x_hi, long long n, unsigned long long *max_bits)
{
    double worst = 0.0;
    const double step = n > 1 ? ((double)x_hi - (double)x_lo) / (double)(n - 1) : 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = (float)((double)x_lo + step * (double)i);
        float e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
        const double t = exp2((double)x);
        if (t >= 1.1754943508222875e-38 && t <= 3.0e38) worst = fmax(worst, fabs((double)e - t) / t);
    }
    atomicMax(max_bits, (unsigned long long)__double_as_longlong(worst));   // non-negative doubles order like integers
}

typedef void (*gibbs_fn)(const SweepParams);

template <int CPL>
static gibbs_fn pick_mode(int mode)
{
    switch (mode) {
        case MODE_PHILOX_EXACT: return gibbs_kernel<CPL, MODE_PHILOX_EXACT>;
        case MODE_PHILOX_FAST: return gibbs_kernel<CPL, MODE_PHILOX_FAST>;
        default: return gibbs_kernel<CPL, MODE_SUPPLIED_EXACT>;
    }
}

static gibbs_fn pick(int cpl, int mode, int threads, bool one_cta_per_sm, bool multi_group = false)
{
    if (multi_group)
        return cpl != 28 || mode == MODE_SUPPLIED_EXACT ? nullptr
               : (mode == MODE_PHILOX_FAST ? gibbs_kernel<28, MODE_PHILOX_FAST, 768, true> : gibbs_kernel<28, MODE_PHILOX_EXACT, 768, true>);
    if (cpl == 28 && threads <= 640 && one_cta_per_sm) {
        switch (mode) {
            case MODE_PHILOX_EXACT: return gibbs_kernel<28, MODE_PHILOX_EXACT, 640>;
            case MODE_PHILOX_FAST: return gibbs_kernel<28, MODE_PHILOX_FAST, 640>;
            default: return gibbs_kernel<28, MODE_SUPPLIED_EXACT, 640>;
        }
    }
    switch (cpl) {
        case 4: return pick_mode<4>(mode);
        case 8: return pick_mode<8>(mode);
        case 16: return pick_mode<16>(mode);
        case 24: return pick_mode<24>(mode);
        case 28: return pick_mode<28>(mode);
        case 32: return pick_mode<32>(mode);
        default: return nullptr;
    }
}

static thread_local int32_t g_last_launches = 0;
static thread_local int32_t g_last_kernel = 0;     // 0 = gibbs_kernel (chains bit-packed per lane), 1 = gibbs_small_kernel, 2 = gibbs_wide_kernel

// gibbs_wide.cu
size_t wide_kernel_smem(int cpl, int mode, int threads, int width, bool single, size_t ring_smem, size_t smem_limit,
                        uint32_t *drawn_offset);
int32_t launch_gibbs_wide(const SweepParams &p, int mode, int threads, int groups, size_t smem, cudaStream_t st);

// gibbs_small.cu
int small_kernel_cpc(int n, int width, int n_tiles, int threads_tile, int chains, int sms);
int32_t launch_gibbs_small(const SweepParams &p, int mode, int threads_tile, int cpc, cudaStream_t st);

}  // namespace b200grbm

using namespace b200grbm;

extern "C" int32_t b200grbm_last_launch_count(void) { return g_last_launches; }

extern "C" int32_t b200grbm_last_sweep_kernel(void) { return g_last_kernel; }

extern "C" int32_t b200grbm_ex2_probe(float x_lo, float x_hi, int64_t n, double *max_rel_err_out, void *stream)
{
    if (n <= 0 || !(x_hi >= x_lo) || max_rel_err_out == nullptr)
        return fail(B200GRBM_EINVAL, "ex2_probe: n=%lld range [%g, %g]", (long long)n, (double)x_lo, (double)x_hi);
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long *dev = nullptr, host = 0;
    B200_CUDA(cudaMallocAsync(&dev, sizeof(host), st));          // measurement aid: the one entry point that allocates
    B200_CUDA(cudaMemsetAsync(dev, 0, sizeof(host), st));
    ex2_probe_kernel<<<1184, 256, 0, st>>>(x_lo, x_hi, (long long)n, dev);
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaMemcpyAsync(&host, dev, sizeof(host), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaFreeAsync(dev, st));
    B200_CUDA(cudaStreamSynchronize(st));
    double v;
    memcpy(&v, &host, sizeof(v));
    *max_rel_err_out = v;
    return 0;
}

extern "C" int32_t b200grbm_sweep_state_offset(int32_t n_tiles)
{
    return 128 + (int32_t)(((int64_t)n_tiles * 8 + 127) / 128 * 128);
}

extern "C" int64_t b200grbm_sweep_smem_bytes(int32_t n, int32_t ell_width, int32_t threads, int32_t n_tiles)
{
    const int64_t state = ((int64_t)n * 4 + 127) / 128 * 128;
    const int64_t info = ((int64_t)n_tiles * 8 + 127) / 128 * 128;
    return 128 + info + state + 2 * (int64_t)(ell_width + 1) * threads * 8;
}

extern "C" int32_t b200grbm_gibbs_sweeps(const b200grbm_sweep_args *a, void *stream)
{
    g_last_launches = 0;
    if (a == nullptr || a->struct_size != sizeof(b200grbm_sweep_args))
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: args is NULL or struct_size mismatch (ABI %d)",
                    B200GRBM_ABI_VERSION);
    if (a->n <= 0 || a->chains <= 0 || a->num_sweeps < 0)
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: n=%d chains=%d num_sweeps=%d", a->n, a->chains, a->num_sweeps);
    if (a->n_pad < a->n || a->ell_width <= 0)
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: n_pad=%d < n=%d or ell_width=%d", a->n_pad, a->n, a->ell_width);
    if (a->n_tiles <= 0)
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: n_tiles=%d", a->n_tiles);
    if (a->tiles_dev == nullptr || a->tile_info_dev == nullptr || (a->num_sweeps > 0 && a->coef_dev == nullptr))
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: tiles_dev / tile_info_dev / coef_dev must not be NULL");
    if ((reinterpret_cast<uintptr_t>(a->tiles_dev) & 15u) != 0)
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: tiles_dev must be 16-byte aligned (bulk copy source)");
    if ((a->state_in_dev != nullptr || a->state_out_dev != nullptr) && a->order_dev == nullptr)
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: order_dev is required for int8 state I/O");
    if (a->threads < 64 || a->threads > 768 || a->threads % 32 != 0)
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: threads=%d must be a multiple of 32 in [64,768]", a->threads);
    if (a->chain_offset % 4 != 0)
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: chain_offset must be a multiple of 4");
    if (a->accept != B200GRBM_ACCEPT_EXACT && a->accept != B200GRBM_ACCEPT_FAST)
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: unknown accept rule %d", a->accept);
    if (a->uniforms_dev != nullptr && a->accept != B200GRBM_ACCEPT_EXACT)
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: supplied uniforms use the exact acceptance rule");
    const int mode = a->uniforms_dev != nullptr ? MODE_SUPPLIED_EXACT
                     : (a->accept == B200GRBM_ACCEPT_FAST ? MODE_PHILOX_FAST : MODE_PHILOX_EXACT);
    if (pick(a->chains_per_lane, mode, a->threads, false) == nullptr)
        return fail(B200GRBM_EUNSUPPORTED, "gibbs_sweeps: chains_per_lane=%d not in {4,8,16,24,28,32}",
                    a->chains_per_lane);
    B200_TRY(require_device());

    SweepParams p;
    p.tiles = reinterpret_cast<const uint2 *>(a->tiles_dev);
    p.tile_info = reinterpret_cast<const int2 *>(a->tile_info_dev);
    p.order = a->order_dev;
    p.coef = a->coef_dev;
    p.uniforms = a->uniforms_dev;
    p.state_in = a->state_in_dev;
    p.packed_in = a->packed_in_dev;
    p.state_out = a->state_out_dev;
    p.packed_out = a->packed_out_dev;
    p.n = a->n;
    p.n_pad = a->n_pad;
    p.width = a->ell_width;
    p.n_tiles = a->n_tiles;
    p.chains = a->chains;
    p.num_sweeps = a->num_sweeps;
    p.sweep_offset = a->sweep_offset;
    p.chain_block0 = (uint32_t)(a->chain_offset >> 2);
    p.drawn_offset = 0;
    p.gpc = 1;
    p.tpg = a->threads;
    p.hi43 = 0x43000000u;
    if (a->chains_per_lane <= 8 && a->ell_width % 4 != 0)
        return fail(B200GRBM_EINVAL, "gibbs_sweeps: chains_per_lane <= 8 consumes slots four at a time; ell_width=%d "
                                     "must be padded to a multiple of 4", a->ell_width);
    p.info_bytes = (uint32_t)(((size_t)a->n_tiles * 8 + 127) / 128 * 128);
    p.state_bytes = (uint32_t)(((size_t)a->n * 4 + 127) / 128 * 128);
    p.tile_bytes = (uint32_t)((size_t)(a->ell_width + 1) * a->threads * 8);
    uint32_t k0 = (uint32_t)a->seed, k1 = (uint32_t)(a->seed >> 32);
    for (int r = 0; r < B200GRBM_PHILOX_ROUNDS; ++r) {
        p.rk[2 * r] = k0;
        p.rk[2 * r + 1] = k1;
        k0 += B200GRBM_PHILOX_W0;
        k1 += B200GRBM_PHILOX_W1;
    }

    // small problems (few chains, small graph): one chain per lane column, every round's table in registers -- the launch
    // is a chain of dependent rounds and this kernel has the shortest round (gibbs_small.cu).  B200GRBM_SMALL=0 disables.
    g_last_kernel = 0;
    if (a->chains_per_lane == 4) {
        const char *env = getenv("B200GRBM_SMALL");
        const int cpc = (env != nullptr && env[0] == '0') ? 0 : small_kernel_cpc(a->n, a->ell_width, a->n_tiles, a->threads, a->chains, sm_count() > 0 ? sm_count() : 148);
        if (cpc > 0) {
            B200_TRY(launch_gibbs_small(p, mode, a->threads, cpc, (cudaStream_t)stream));
            g_last_launches = 1;
            g_last_kernel = 1;
            return 0;
        }
    }

    size_t smem = (size_t)b200grbm_sweep_smem_bytes(a->n, a->ell_width, a->threads, a->n_tiles);
    int dev = 0, smem_optin = 0;
    B200_CUDA(cudaGetDevice(&dev));
    B200_CUDA(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    // small graphs: keep every round's tile resident instead of streaming through the 2-stage ring
    const size_t smem_resident = smem + (size_t)(a->n_tiles > 2 ? a->n_tiles - 2 : 0) * p.tile_bytes;
    p.resident = (a->n_tiles <= 2 || (smem_resident <= (size_t)smem_optin &&
                                      (size_t)a->n_tiles * p.tile_bytes < (1u << 20))) ? 1u : 0u;
    if (p.resident) smem = smem_resident;
    // resident tables + many groups: several groups per CTA share the one copy of the tables (see the kernel)
    const int groups_all = (a->chains + a->chains_per_lane - 1) / a->chains_per_lane;
    p.gpc = 1;
    p.tpg = a->threads;
    if (p.resident && pick(a->chains_per_lane, mode, a->threads, false, true) != nullptr) {
        const char *gpc_env = getenv("B200GRBM_GPC");
        const int forced = gpc_env != nullptr ? atoi(gpc_env) : 0;
        const int sms = sm_count() > 0 ? sm_count() : 148;
        int smem_sm_all = 0;
        B200_CUDA(cudaDeviceGetAttribute(&smem_sm_all, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        // cost model (relative time; the same one as sampler.py::_plan_resident): whole waves of `per_sm` CTAs per SM
        // plus a partial last wave; an SM with w >= 8 resident warps runs at w / (w + 4.35) of its peak (fitted on the
        // 256-spin graph: 131072 chains x 100 sweeps take 7.95 ms with one group per CTA = 8 warps per SM, 6.20 ms with four
        // = 24 warps) and in proportion to w below that
        const double wpg = a->threads / 32.0;
        const auto sm_time = [&](int ctas_on_sm, int gpc) {
            const double w = ctas_on_sm * gpc * wpg;
            return ctas_on_sm * gpc / (w >= 8.0 ? w / (w + 4.35) : w / 12.35);
        };
        double best = 1e300;
        for (int gpc = 1; gpc <= 15 && gpc * a->threads <= 768; ++gpc) {
            if (forced > 0 && gpc != forced) continue;
            const size_t smem_cta = smem + (size_t)(gpc - 1) * p.state_bytes;
            if (smem_cta > (size_t)smem_optin) break;
            const int ctas = (groups_all + gpc - 1) / gpc;
            int per_sm = (int)std::min<size_t>({(size_t)(65536 / (gpc * a->threads * 80)), (size_t)smem_sm_all / (smem_cta + 1024),
                                                (size_t)(2048 / (gpc * a->threads)), (size_t)32});
            if (per_sm < 1) per_sm = 1;
            const int slots = sms * per_sm, full = ctas / slots, rem = ctas - full * slots;
            const double time = full * sm_time(per_sm, gpc) + (rem > 0 ? sm_time((rem + sms - 1) / sms, gpc) : 0.0);
            if (time < best * 0.98) {           // a larger CTA has to win by 2 %
                best = time;
                p.gpc = gpc;
            }
        }
        smem += (size_t)(p.gpc - 1) * p.state_bytes;
    }
    // many groups, narrow CTAs: ONE tile stage per CTA and two CTAs per SM -- twice the warps per scheduler, each
    // CTA's round barrier and copy latency covered by the other (Zephyr Z15: 2 x 384 threads instead of 1 x 480)
    const int groups = (a->chains + a->chains_per_lane - 1) / a->chains_per_lane;
    int smem_sm = 0;
    B200_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
    const size_t smem_single = smem - p.tile_bytes;
    p.single = (!p.resident && a->threads <= 384 && groups >= 2 * (sm_count() > 0 ? sm_count() : 148) &&
                2 * (smem_single + 1024) <= (size_t)smem_sm) ? 1u : 0u;
    if (p.single) smem = smem_single;
    if (smem > (size_t)smem_optin)
        return fail(B200GRBM_EUNSUPPORTED,
                    "gibbs_sweeps: n=%d width=%d threads=%d need %zu B of shared memory (> %d); use fewer threads",
                    a->n, a->ell_width, a->threads, smem, smem_optin);
    // int8 rows: with a packed output buffer at hand the sweep kernel writes only its words and a second kernel unpacks
    // them (coalesced); without one the sweep kernel stores the bytes itself
    const bool unpack = a->state_out_dev != nullptr && a->packed_out_dev != nullptr;
    if (unpack) p.state_out = nullptr;
    // throughput geometry (28 chains per lane, one CTA per SM, tile ring, known CTA size and width): the specialised
    // kernel of gibbs_wide.cu.  B200GRBM_WIDE=0 keeps the generic kernel (same results).
    const char *wide_env = getenv("B200GRBM_WIDE");
    const size_t smem_wide = (p.resident || (wide_env != nullptr && wide_env[0] == '0') ||
                              (long long)a->num_sweeps * a->n_tiles >= (1ll << 31))
                                 ? 0 : wide_kernel_smem(a->chains_per_lane, mode, a->threads, a->ell_width, p.single != 0, smem,
                                                        (size_t)smem_optin, &p.drawn_offset);
    g_last_kernel = 0;
    if (smem_wide > 0 && smem_wide <= (size_t)smem_optin) {
        B200_TRY(launch_gibbs_wide(p, mode, a->threads, groups, smem_wide, (cudaStream_t)stream));
        g_last_kernel = 2;
    } else {
        gibbs_fn fn = pick(a->chains_per_lane, mode, a->threads * p.gpc, !p.single && p.gpc == 1, p.gpc > 1);
        B200_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        B200_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        fn<<<(groups + p.gpc - 1) / p.gpc, a->threads * p.gpc, smem, (cudaStream_t)stream>>>(p);
        B200_CUDA(cudaGetLastError());
    }
    g_last_launches = 1;
    if (unpack) {
        int rows_per_pass = (int)((64u * 1024u - 16u) / (uint32_t)a->n);        // <= 64 KB of rows per CTA: three CTAs per SM
        if (rows_per_pass > a->chains_per_lane) rows_per_pass = a->chains_per_lane;
        if (rows_per_pass < 1) rows_per_pass = 1;
        const size_t usmem = (size_t)rows_per_pass * a->n + 16;
        if (usmem > (size_t)smem_optin)
            return fail(B200GRBM_EUNSUPPORTED, "gibbs_sweeps: n=%d too large for the state unpack kernel's shared-memory row", a->n);
        B200_CUDA(cudaFuncSetAttribute(unpack_state_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)usmem));
        unpack_state_kernel<<<groups, 256, usmem, (cudaStream_t)stream>>>(a->packed_out_dev, a->chains, a->chains_per_lane, a->n,
                                                                          a->n_pad, a->order_dev, a->state_out_dev, rows_per_pass);
        B200_CUDA(cudaGetLastError());
        g_last_launches = 2;
    }
    return 0;
}
