// Pieces shared by the sweep kernels (gibbs.cu: chains bit-packed per lane; gibbs_small.cu: one chain per lane):
// launch parameters, bulk-copy / mbarrier wrappers and the contract arithmetic of include/b200grbm_spec.h
// (Philox4x32-10, uniforms, exact acceptance, the bracketed quick decision).
#pragma once

#include "common.cuh"

namespace b200grbm {

enum { MODE_PHILOX_EXACT = 0, MODE_PHILOX_FAST = 1, MODE_SUPPLIED_EXACT = 2 };

struct SweepParams {
    const uint2 *tiles;       // [n_tiles][1 + width][threads]: row 0 {f0 bits, -}, rows 1.. {2J bits, byte offset
                              //  of the neighbour's state word in dynamic shared memory}
    const int2 *tile_info;    // [n_tiles] {first visit position, spins in this round}
    const int32_t *order;
    const float *coef;
    const float *uniforms;
    const int8_t *state_in;
    const uint32_t *packed_in;
    int8_t *state_out;
    uint32_t *packed_out;
    int n, n_pad, width, n_tiles;
    int chains, num_sweeps;
    uint32_t sweep_offset;
    uint32_t chain_block0;  // (chain_offset >> 2)
    uint32_t info_bytes;    // shared-memory bytes reserved for the round table (multiple of 128)
    uint32_t state_bytes;   // shared-memory bytes reserved for W (multiple of 128)
    uint32_t tile_bytes;    // (width + 1) * threads * 8
    uint32_t resident;      // 1: all n_tiles tiles fit in shared memory and are copied once (small graphs)
    uint32_t single;        // 1: one tile stage per CTA, two CTAs per SM cover each other's copy latency
    int gpc, tpg;           // gibbs_kernel, resident tables only: chain groups per CTA and threads per group (gpc == 1: the CTA)
    uint32_t hi43;          // 0x43000000 (exponent of 128.0f): PRMT operand of the packed acceptance, kept out of the immediates
    uint32_t drawn_offset;  // byte offset of the pre-drawn Philox words [calls][threads] x 16 B (gibbs_kernel<.., PD = true>)
    uint32_t rk[2 * B200GRBM_PHILOX_ROUNDS];
};

// ---------------------------------------------------------------- bulk copy + mbarrier (PTX)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

// named barrier over one chain group of a multi-group CTA (id 1 .. 15, n_threads a multiple of 32)
__device__ __forceinline__ void group_bar_sync(int id, int n_threads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n_threads) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity)) return;            // the usual case costs two instructions
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();               // a lost copy / arrival must fail the launch, never hang the GPU
}

// ---------------------------------------------------------------- contract arithmetic

__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           const SweepParams &p, uint32_t (&out)[4])
{
#ifdef B200_EXP_NOPHILOX     // timing experiment: a few integer ops instead of the generator
    out[0] = c0 * 0x9E3779B9u ^ c1; out[1] = out[0] ^ (c2 << 3); out[2] = out[1] + c3; out[3] = out[2] ^ p.rk[0];
    return;
#endif
#pragma unroll
    for (int r = 0; r < B200GRBM_PHILOX_ROUNDS; ++r) {
        const uint64_t p0 = (uint64_t)B200GRBM_PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)B200GRBM_PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ p.rk[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ p.rk[2 * r + 1];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float uniform_from_m23(uint32_t m23)
{
    // as_float(m23 | 0x3f800000) - 1 + 2^-24, one exact add (spec header)
    return __fadd_rn(u2f(m23 | 0x3f800000u), -1.0f + B200GRBM_UNIFORM_HALF_ULP);
}

// halfword j (0..7) of a Philox call
__device__ __forceinline__ uint32_t halfword(const uint32_t (&r)[4], int j)
{
    return (j & 1) ? (r[j >> 1] >> 16) : (r[j >> 1] & 0xffffu);
}

// (hw + 0.5) * 2^-16 : midpoint of the 128 uniforms that share the 16 high bits hw.
// PRMT builds as_float(0x43000000 | hw) = 128 + hw 2^-16; one exact add maps it to the midpoint.
__device__ __forceinline__ float uniform_midpoint(const uint32_t (&r)[4], int j)
{
    const uint32_t big = __byte_perm(r[j >> 1], 0x43000000u, (j & 1) ? 0x7632u : 0x7610u);
    return __fadd_rn(u2f(big), -(128.0f - 0x1.0p-17f));
}

// The contract's decision (include/b200grbm_spec.h): +1 iff fmaf(v, exp2_poly(clamp(f coef)), v) < 1.
__device__ __forceinline__ bool accept_exact(float f, float coef, float v)
{
    float x = __fmul_rn(f, coef);
    x = fminf(fmaxf(x, -B200GRBM_EXP2_CLAMP), B200GRBM_EXP2_CLAMP);
    const float t = __fadd_rn(x, B200GRBM_EXP2_MAGIC);
    const float nn = __fadd_rn(t, -B200GRBM_EXP2_MAGIC);
    const float r = __fadd_rn(x, -nn);
    float q = B200GRBM_EXP2_C5;
    q = __fmaf_rn(q, r, B200GRBM_EXP2_C4);
    q = __fmaf_rn(q, r, B200GRBM_EXP2_C3);
    q = __fmaf_rn(q, r, B200GRBM_EXP2_C2);
    q = __fmaf_rn(q, r, B200GRBM_EXP2_C1);
    q = __fmaf_rn(q, r, B200GRBM_EXP2_C0);
    const float e = u2f(f2u(q) + (f2u(t) << 23));
    return __fmaf_rn(v, e, v) < 1.0f;
}

// Bracketed decision.  d = vm (1 + e~) - 1 with e~ = MUFU.EX2; the sign of d is the contract's
// decision whenever |d| exceeds the bracket  K1 (1 + e~) + K2:
//   |v - vm| <= 2^-17 (Philox mode: vm is the 16-bit midpoint; supplied mode: vm = v)
//   |e~ - e| <= 2^-20 e  (ex2.approx.ftz: 2^-22 relative; contract polynomial: 2.4e-7), roundings 2^-24
//   => |v (1 + e) - vm (1 + e~)| <= 2^-17 (1 + e)(1 + 2^-19) + (d + 1) 2^-19  <  K1 (1 + e~) + K2 - 2^-24
// with K1 = 2^-17 (1 + 2^-10), K2 = 2^-17.  `sure` accumulates over the lane-task.

// x is NOT clamped here (Philox modes: v >= 2^-24).  x > 128 gives e~ = g = d = +inf: sign clear = the contract's decision (its clamp at
// 120 leaves e >= 2^120 > 1/v for every v >= 2^-24), and the mark m = (-inf) + inf is the canonical NaN
// 0x7fffffff, sign clear = "sure".  x < -126 flushes e~ to 0, g = 1, d = vm - 1 < 0: also the contract's
// decision.  One instruction less per decision on the ALU pipe, the busiest one in this phase.
// Supplied uniforms (CLAMP) may be any float in (0,1), also below 2^-120 where the contract's clamp matters: that
// mode keeps the upper clamp, which makes g finite and the bracket argument unconditional.
template <bool CHECK, bool CLAMP = false>
__device__ __forceinline__ uint32_t decide_quick(float f, float coef, float vm, uint32_t &unsure)
{
    float x = __fmul_rn(f, coef);
    if (CLAMP) x = fminf(x, B200GRBM_EXP2_CLAMP);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
    const float g = __fadd_rn(e, 1.0f);
    const float d = __fmaf_rn(vm, g, -1.0f);
    if (CHECK) {
        // sign bit of  |d| - K2 - K1 g  marks the decision as inside its bracket; one funnel shift per
        // decision collects the marks (dense: chain c -> bit c)
        const float m = __fmaf_rn(g, -B200GRBM_LAZY_K1, __fadd_rn(fabsf(d), -B200GRBM_LAZY_K2));
        unsure = __funnelshift_l(f2u(m), unsure, 1);
    }
    return f2u(d);   // bit 31 set <=> +1
}

}  // namespace b200grbm
