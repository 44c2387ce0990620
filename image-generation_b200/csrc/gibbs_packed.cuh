// Pieces shared by the bit-packed sweep kernels (gibbs.cu: generic; gibbs_wide.cu: specialised 28-chain kernel):
// the state-word bit layout, the predicated-add neighbour slots and the decisions of one lane-task.
#pragma once

#include "gibbs_common.cuh"

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

// ---- packed fp32x2 arithmetic (Blackwell FMUL2 / FADD2 / FFMA2: one issue slot for two chains).  Each lane of a
// packed op is the same correctly rounded IEEE operation as its scalar form, so results are unchanged; what is saved
// is issue slots, the resource this kernel runs out of (the fma pipe takes two passes per packed op).
__device__ __forceinline__ uint64_t pack2(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// as_float(0x43000000 | halfword j of the call) = 128 + hw 2^-16.  `hi43` holds 0x43000000 in a REGISTER (PRMT takes one
// immediate: with the constant as the immediate ptxas materialises the selector per use, one extra instruction per decision).
__device__ __forceinline__ float uniform_big(const uint32_t (&r)[4], int j, uint32_t hi43)
{
    return u2f(__byte_perm(r[j >> 1], hi43, (j & 1) ? 0x7632u : 0x7610u));
}

struct Pair2Consts {
    uint64_t coef2, one2, mone2, negc2;
    uint32_t hi43;
};

// Two bracketed decisions (chains c_hi = c_lo + 1) with the shared steps in packed form; see decide_quick for the
// arithmetic.  Returns d (sign bit = decision) and, if CHECK, the marks m of both chains.
template <bool CHECK>
__device__ __forceinline__ void decide_quick2(float f_lo, float f_hi, float big_lo, float big_hi, const Pair2Consts &k,
                                              float &d_lo, float &d_hi, float &m_lo, float &m_hi)
{
    float x_lo, x_hi, e_lo, e_hi, g_lo, g_hi;
    unpack2(mul2(pack2(f_lo, f_hi), k.coef2), x_lo, x_hi);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e_lo) : "f"(x_lo));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e_hi) : "f"(x_hi));
    const uint64_t g2 = add2(pack2(e_lo, e_hi), k.one2);
    const uint64_t vm2 = add2(pack2(big_lo, big_hi), k.negc2);
    unpack2(fma2(vm2, g2, k.mone2), d_lo, d_hi);
    if (CHECK) {
        unpack2(g2, g_lo, g_hi);
        m_lo = __fmaf_rn(g_lo, -B200GRBM_LAZY_K1, __fadd_rn(fabsf(d_lo), -B200GRBM_LAZY_K2));
        m_hi = __fmaf_rn(g_hi, -B200GRBM_LAZY_K1, __fadd_rn(fabsf(d_hi), -B200GRBM_LAZY_K2));
    }
}

// Decisions of one lane-task: the new state word of visit position pp for the CPL chains of the group.
// Chains are taken in descending order so that one funnel shift per decision (sign bit of d into bit 0)
// assembles the word.  SHIFT = (first global chain of the group) mod 8, in {0, 4}: Philox blocks hold 8 chains.
//
// PRE (small groups, latency-bound): the Philox words were drawn by the caller before the neighbour loop and
// arrive in R[call][word].
//
// PD (throughput kernel, one CTA per SM): the words were drawn by this thread while it waited for the previous round's
// stragglers (see the round barrier in gibbs_kernel) and sit in shared memory, call-major: drawn[call * stride].
//
// PACK2: the decisions are taken two chains at a time with packed fp32x2 arithmetic (decide_quick2); same results.
template <int CPL, int MODE, int SHIFT, bool PRE, bool PD = false, bool PACK2 = false>
__device__ __forceinline__ uint32_t decide_word(const float (&f)[CPL], float coef, uint32_t pp, uint32_t sweep,
                                                uint32_t blk8, const SweepParams &p, int t, int chain0,
                                                const uint32_t (&R)[2][4], const uint4 *drawn = nullptr, int stride = 0,
                                                const Pair2Consts *k2 = nullptr)
{
    static_assert(!PACK2 || (MODE != MODE_SUPPLIED_EXACT && SHIFT % 2 == 0 && CPL % 2 == 0), "PACK2: Philox modes, even groups");
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
            if constexpr (PD) {
                const uint4 v = drawn[call * stride];
                r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
            } else if constexpr (PRE) {
#pragma unroll
                for (int i = 0; i < 4; ++i) r[i] = R[call][i];
            } else {
                philox4x32(pp, blk8 + call, sweep, B200GRBM_STREAM_SWEEP, p, r);
            }
            if constexpr (PACK2) {
#pragma unroll
                for (int j = 7; j >= 1; j -= 2) {
                    const int c = 8 * call + j - SHIFT;          // pair (c, c - 1): both inside the group or both outside
                    if (c - 1 < 0 || c >= CPL) continue;
                    float d_lo, d_hi, m_lo = 0.f, m_hi = 0.f;
                    decide_quick2<CHECK>(f[c - 1], f[c], uniform_big(r, j - 1, k2->hi43), uniform_big(r, j, k2->hi43), *k2,
                                         d_lo, d_hi, m_lo, m_hi);
                    if (CPL == 28 && (c + 1) % 7 == 0) neww <<= 1;
                    neww = __funnelshift_l(f2u(d_hi), neww, 1);
                    if (CPL == 28 && c % 7 == 0) neww <<= 1;
                    neww = __funnelshift_l(f2u(d_lo), neww, 1);
                    if (CHECK) {
                        unsure = __funnelshift_l(f2u(m_hi), unsure, 1);
                        unsure = __funnelshift_l(f2u(m_lo), unsure, 1);
                    }
                }
            } else {
#pragma unroll
                for (int j = 7; j >= 0; --j) {
                    const int c = 8 * call + j - SHIFT;
                    if (c < 0 || c >= CPL) continue;
                    if (CPL == 28 && (c + 1) % 7 == 0) neww <<= 1;
                    neww = __funnelshift_l(decide_quick<CHECK>(f[c], coef, uniform_midpoint(r, j), unsure), neww, 1);
                }
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

// The group's packed state, shared memory <- int8 rows / packed words / the initial-state Philox stream.
template <int CPL>
__device__ __forceinline__ void load_group_state(const SweepParams &p, uint32_t *W, int tid, int nthr, int g, int chain0,
                                                 int nvalid, uint32_t dense_mask, uint32_t blk0)
{
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
}

// ... and back: packed words (coalesced), and int8 rows when no unpack kernel follows (bits of chains beyond nvalid masked)
template <int CPL>
__device__ __forceinline__ void store_group_state(const SweepParams &p, const uint32_t *W, int tid, int nthr, int g, int chain0,
                                                  int nvalid, uint32_t dense_mask)
{
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

}  // namespace b200grbm
