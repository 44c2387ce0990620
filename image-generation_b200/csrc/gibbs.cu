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
#include "gibbs_common.cuh"

#include <cstdlib>
#include <cstring>

namespace b200grbm {

// Bit of the in-kernel state word that holds chain c.  CPL = 28 leaves bit 7 of every byte
// unused so that R2P (7 predicates from one byte) covers the word with four instructions.
template <int CPL>
__device__ __forceinline__ constexpr int bitpos(int c) { return CPL == 28 ? c + c / 7 : c; }

template <int CPL>
__device__ __forceinline__ uint32_t dense_to_kernel(uint32_t w)
{
    if (CPL != 28) return w;
    return (w & 0x7fu) | ((w & 0x3f80u) << 1) | ((w & 0x1fc000u) << 2) | ((w & 0xfe00000u) << 3);
}

template <int CPL>
__device__ __forceinline__ uint32_t kernel_to_dense(uint32_t w)
{
    if (CPL != 28) return w;
    return (w & 0x7fu) | ((w >> 1) & 0x3f80u) | ((w >> 2) & 0x1fc000u) | ((w >> 3) & 0xfe00000u);
}

// How a lane-task walks its neighbour slots.  Large groups (CPL >= 16, value 1): slot 0 initialises the
// fields, then seven slots per loop iteration -- the other warps of the scheduler cover the LDS -> LDS
// dependency (entry, then the state word it points at).  Small groups (CPL <= 8, value 4: few chains spread
// over many CTAs, e.g. the reference's 256 reads) are latency-bound with one or two warps per scheduler: the
// Philox words are drawn first and the loads of eight (then four) slots are issued together, so a round costs
// two shared-memory latencies per batch instead of two per slot; the tables pad the width to a multiple of 4.
template <int CPL>
struct SlotUnroll { static constexpr int value = CPL <= 8 ? 4 : 1; };

// f[c] += j2 where bit bitpos(c) of w is set.  Written in PTX so that every bit -- bit 0 included, which the
// C++ front end would canonicalise into a different test -- has the same and/setp/predicated-add shape and
// ptxas folds seven tests into one R2P.
template <int CPL, int C>
__device__ __forceinline__ void add_slot_from(float (&f)[CPL], uint32_t w, float j2)
{
    if constexpr (C < CPL) {
        asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
            "and.b32 t, %2, %3;\n\t"
            "setp.ne.s32 p, t, 0;\n\t"
            "@p add.rn.f32 %0, %0, %1;\n\t}"
            : "+f"(f[C])
            : "f"(j2), "r"(w), "n"(1u << bitpos<CPL>(C)));
        add_slot_from<CPL, C + 1>(f, w, j2);
    }
}

template <int CPL>
__device__ __forceinline__ void add_slot(float (&f)[CPL], uint32_t w, float j2)
{
    add_slot_from<CPL, 0>(f, w, j2);
}

template <int CPL, int C>
__device__ __forceinline__ void init_slot_from(float (&f)[CPL], uint32_t w, float fz, float fa)
{
    if constexpr (C < CPL) {
        asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\t"
            "and.b32 t, %3, %4;\n\t"
            "setp.ne.s32 p, t, 0;\n\t"
            "selp.f32 %0, %1, %2, p;\n\t}"
            : "=f"(f[C])
            : "f"(fa), "f"(fz), "r"(w), "n"(1u << bitpos<CPL>(C)));
        init_slot_from<CPL, C + 1>(f, w, fz, fa);
    }
}

// slot 0 doubles as the initialisation: f = f0 or f0 + 2J (one select instead of a move and an add)
template <int CPL>
__device__ __forceinline__ void init_slot(float (&f)[CPL], uint32_t w, float fz, float j2)
{
    init_slot_from<CPL, 0>(f, w, fz, __fadd_rn(fz, j2));
}

// state word at a byte offset from the start of dynamic shared memory (what the tables' .nbr field holds)
__device__ __forceinline__ uint32_t lds_word(const unsigned char *smem, uint32_t byte_off)
{
    return *reinterpret_cast<const uint32_t *>(smem + byte_off);
}

// Cold paths: contract arithmetic for the decisions marked in `unsure` (usually one).  Out of line, fields
// through local memory, and only the marked chains are redone: the warp that lands here is the straggler the
// whole CTA waits for at the round barrier, so what matters is the latency of this path, not its size.
template <int CPL, int SHIFT>
__device__ __noinline__ uint32_t fix_word_philox(uint32_t neww, uint32_t unsure, const float *fl, float coef,
                                                 uint32_t pp, uint32_t sweep, uint32_t blk8, const SweepParams &p)
{
    do {
        const int c = 31 - __clz(unsure);
        unsure &= ~(1u << c);
        const int hidx = c + SHIFT;
        uint32_t r[4], q[4];
        philox4x32(pp, blk8 + (uint32_t)(hidx >> 3), sweep, B200GRBM_STREAM_SWEEP, p, r);
        philox4x32(pp, blk8 + (uint32_t)(hidx >> 3), sweep, B200GRBM_STREAM_SWEEP_LO, p, q);
        const int j = hidx & 7, sh = 16 * (j & 1);
        const uint32_t wsel = (uint32_t)(j >> 1);
        const uint32_t rw = wsel == 0 ? r[0] : wsel == 1 ? r[1] : wsel == 2 ? r[2] : r[3];
        const uint32_t qw = wsel == 0 ? q[0] : wsel == 1 ? q[1] : wsel == 2 ? q[2] : q[3];
        const float v = uniform_from_m23((((rw >> sh) & 0xffffu) << 7) | (((qw >> sh) & 0xffffu) >> 9));
        const uint32_t bit = 1u << (CPL == 28 ? c + c / 7 : c);
        neww = accept_exact(fl[c], coef, v) ? (neww | bit) : (neww & ~bit);
    } while (unsure != 0);
    return neww;
}

template <int CPL>
__device__ __noinline__ uint32_t fix_word_supplied(uint32_t neww, uint32_t unsure, const float *fl, float coef,
                                                   uint32_t pp, const SweepParams &p, int t, int chain0)
{
    do {
        const int c = 31 - __clz(unsure);
        unsure &= ~(1u << c);
        const int cc = min(chain0 + c, p.chains - 1);
        const float v = __ldg(p.uniforms + ((size_t)t * p.chains + cc) * p.n + pp);
        const uint32_t bit = 1u << (CPL == 28 ? c + c / 7 : c);
        neww = accept_exact(fl[c], coef, v) ? (neww | bit) : (neww & ~bit);
    } while (unsure != 0);
    return neww;
}

// Decisions of one lane-task: the new state word of visit position pp for the CPL chains of the group.
// Chains are taken in descending order so that one funnel shift per decision (sign bit of d into bit 0)
// assembles the word.  SHIFT = (first global chain of the group) mod 8, in {0, 4}: Philox blocks hold 8 chains.
//
// PRE (small groups, latency-bound): the Philox words were drawn by the caller before the neighbour loop and
// arrive in R[call][word].
template <int CPL, int MODE, int SHIFT, bool PRE>
__device__ __forceinline__ uint32_t decide_word(const float (&f)[CPL], float coef, uint32_t pp, uint32_t sweep,
                                                uint32_t blk8, const SweepParams &p, int t, int chain0,
                                                const uint32_t (&R)[2][4])
{
    constexpr int NC = (CPL + SHIFT + 7) / 8;
    static_assert(!PRE || NC <= 2, "pre-drawn Philox words: at most two calls per lane-task");
    constexpr bool CHECK = MODE != MODE_PHILOX_FAST;
    uint32_t neww = 0, unsure = 0;
    if constexpr (MODE == MODE_SUPPLIED_EXACT) {
#pragma unroll
        for (int c = CPL - 1; c >= 0; --c) {
            const int cc = min(chain0 + c, p.chains - 1);
            const float v = __ldg(p.uniforms + ((size_t)t * p.chains + cc) * p.n + pp);
            if (CPL == 28 && (c + 1) % 7 == 0) neww <<= 1;
            neww = __funnelshift_l(decide_quick<true, true>(f[c], coef, v, unsure), neww, 1);
        }
    } else {
#pragma unroll
        for (int call = NC - 1; call >= 0; --call) {
            uint32_t r[4];
            if constexpr (PRE) {
#pragma unroll
                for (int i = 0; i < 4; ++i) r[i] = R[call][i];
            } else {
                philox4x32(pp, blk8 + call, sweep, B200GRBM_STREAM_SWEEP, p, r);
            }
#pragma unroll
            for (int j = 7; j >= 0; --j) {
                const int c = 8 * call + j - SHIFT;
                if (c < 0 || c >= CPL) continue;
                if (CPL == 28 && (c + 1) % 7 == 0) neww <<= 1;
                neww = __funnelshift_l(decide_quick<CHECK>(f[c], coef, uniform_midpoint(r, j), unsure), neww, 1);
            }
        }
    }
    if (CHECK && unsure != 0) {
        // rare (about 6e-4 of the lane-tasks): a decision sits inside its bracket
        float fl[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) fl[c] = f[c];
        if constexpr (MODE == MODE_SUPPLIED_EXACT)
            neww = fix_word_supplied<CPL>(neww, unsure, fl, coef, pp, p, t, chain0);
        else
            neww = fix_word_philox<CPL, SHIFT>(neww, unsure, fl, coef, pp, sweep, blk8, p);
    }
    return neww;
}

// MAXT = largest CTA this instantiation may be launched with.  768 caps ptxas at 80 registers per thread (and is the
// one two 384-thread CTAs per SM need); the 28-chain kernels also exist for MAXT = 640 -- 96 registers, a looser
// schedule of the acceptance phase: 32.8 -> 32.0 ms on P16 (9 rounds of 640 lanes).
template <int CPL, int MODE, int MAXT = 768>
__global__ void __launch_bounds__(MAXT, 1) gibbs_kernel(const __grid_constant__ SweepParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw);                 // 2 mbarriers (16 B, padded to 128)
    int2 *tinfo = reinterpret_cast<int2 *>(smem_raw + 128);                  // round table, copied once
    uint32_t *W = reinterpret_cast<uint32_t *>(smem_raw + 128 + p.info_bytes);
    unsigned char *stage0 = smem_raw + 128 + p.info_bytes + p.state_bytes;

    const int tid = threadIdx.x;
    const int nthr = blockDim.x;
    const int g = blockIdx.x;
    const int chain0 = g * CPL;  // first chain of this group, local to the call
    const int nvalid = min(CPL, p.chains - chain0);
    const uint32_t dense_mask = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
    const uint32_t blk0 = p.chain_block0 + (uint32_t)(chain0 >> 2);   // global chain / 4 (initial-state stream)
    const uint32_t blk8 = blk0 >> 1;                                   // global chain / 8 (sweep streams)
    const bool shift4 = (blk0 & 1u) != 0;                              // group starts in the middle of a block of 8
    const uint32_t bar_addr = smem_u32(bars);
    const uint32_t stage_addr = smem_u32(stage0);
    const long long total_tiles = (long long)p.num_sweeps * p.n_tiles;

    if (tid == 0) {
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
    for (int k = tid; k < p.n_tiles; k += nthr) tinfo[k] = __ldg(p.tile_info + k);

    // ---- load / initialise the group's packed state
    for (int pp = tid; pp < p.n; pp += nthr) {
        uint32_t w = 0;
        if (p.state_in != nullptr) {
            const int node = p.order[pp];
            for (int c = 0; c < nvalid; ++c)
                w |= (p.state_in[(size_t)(chain0 + c) * p.n + node] > 0 ? 1u : 0u) << c;
        } else if (p.packed_in != nullptr) {
            w = p.packed_in[(size_t)g * p.n_pad + pp];
        } else {
#pragma unroll
            for (int c4 = 0; c4 < CPL / 4; ++c4) {
                uint32_t r[4];
                philox4x32((uint32_t)pp, blk0 + c4, 0u, B200GRBM_STREAM_INIT, p, r);
#pragma unroll
                for (int j = 0; j < 4; ++j) w |= (r[j] >> 31) << (4 * c4 + j);
            }
        }
        W[pp] = dense_to_kernel<CPL>(w & dense_mask);
    }
    __syncthreads();

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
                    for (int i = 0; i < 8; ++i) w[i] = lds_word(smem_raw, e[i].y);
#pragma unroll
                    for (int i = 0; i < 8; ++i) add_slot<CPL>(f, w[i], u2f(e[i].x));
                }
                if (k < p.width) {
                    const uint2 e0 = ep[0], e1 = ep[nthr], e2 = ep[2 * nthr], e3 = ep[3 * nthr];
                    const uint32_t w0 = lds_word(smem_raw, e0.y), w1 = lds_word(smem_raw, e1.y),
                                   w2 = lds_word(smem_raw, e2.y), w3 = lds_word(smem_raw, e3.y);
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
                    init_slot<CPL>(f, lds_word(smem_raw, e0.y), fz, u2f(e0.x));
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
                    for (int i = 0; i < 7; ++i) w[i] = lds_word(smem_raw, e[i].y);
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
                    const uint32_t wa = lds_word(smem_raw, ea.y), wb = lds_word(smem_raw, eb.y);
                    add_slot<CPL>(f, wa, u2f(ea.x));
                    add_slot<CPL>(f, wb, u2f(eb.x));
                }
                if (k < p.width) {
                    const uint2 ea = *ep;
                    add_slot<CPL>(f, lds_word(smem_raw, ea.y), u2f(ea.x));
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
            W[pp] = shift4 ? decide_word<CPL, MODE, 4, PRE>(f, coef, (uint32_t)pp, sweep, blk8, p, t, chain0, R)
                           : decide_word<CPL, MODE, 0, PRE>(f, coef, (uint32_t)pp, sweep, blk8, p, t, chain0, R);
#endif
            // bits of chains beyond nvalid are masked at write-back
        }
#ifndef B200_EXP_NOBARRIER   // timing experiment only (tools/build_variant.sh): results are wrong without it
        __syncthreads();
#endif
        if (++tile == p.n_tiles) {
            tile = 0;
            ++t;
            coef = coef_next;
            if (t + 1 < p.num_sweeps) coef_next = __ldg(p.coef + t + 1);
        }
    }

    // ---- write back
    for (int pp = tid; pp < p.n; pp += nthr) {
        const uint32_t w = kernel_to_dense<CPL>(W[pp]) & dense_mask;
        if (p.packed_out != nullptr) p.packed_out[(size_t)g * p.n_pad + pp] = w;
        if (p.state_out != nullptr) {
            const int node = p.order[pp];
            for (int c = 0; c < nvalid; ++c)
                p.state_out[(size_t)(chain0 + c) * p.n + node] = (w >> c) & 1u ? (int8_t)1 : (int8_t)-1;
        }
    }
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

static gibbs_fn pick(int cpl, int mode, int threads, bool one_cta_per_sm)
{
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
static thread_local int32_t g_last_kernel = 0;     // 0 = gibbs_kernel (chains bit-packed per lane), 1 = gibbs_small_kernel

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
    gibbs_fn fn = pick(a->chains_per_lane, mode, a->threads, !p.single);
    B200_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    // int8 rows: with a packed output buffer at hand the sweep kernel writes only its words and a second kernel unpacks
    // them (coalesced); without one the sweep kernel stores the bytes itself
    const bool unpack = a->state_out_dev != nullptr && a->packed_out_dev != nullptr;
    if (unpack) p.state_out = nullptr;
    fn<<<groups, a->threads, smem, (cudaStream_t)stream>>>(p);
    B200_CUDA(cudaGetLastError());
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
