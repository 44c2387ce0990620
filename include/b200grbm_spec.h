/*
 * b200grbm_spec.h -- the NUMERICAL CONTRACT of the GRBM heat-bath sampler.
 *
 * Everything in this header is arithmetic that must be reproduced bit-for-bit
 * by every implementation of the sampler (the sm_100a kernels in
 * image-generation_b200/csrc and the CPU oracle in oracle/).  It contains
 * constants and plain-C inline helpers only; it is included by both sides so
 * that there is exactly one statement of the rule.
 *
 * What is being restated: the classical stand-in for the QPU call
 *   sampler.sample_ising(h, J, **sample_params)
 * reached from the reference at src/model_wrapper.py:309-316 and
 * src/utils/persistent_qpu_sampler.py:71-78.
 * The reference delegates the arithmetic to dwave-samplers' simulated annealer
 * (un-vendored, not installed): single-spin heat-bath
 * ("Gibbs") updates visiting spins in index order, one inverse temperature per
 * sweep.  PARITY UNPINNED: the reference holds no test or golden vector for
 * this path, so this header *is* the pinned definition.
 *
 * Ising convention (static/eq6.png, README.md:140-142):
 *   E(s) = sum_i h_i s_i + sum_(i<j) J_ij s_i s_j,     P(s) ~ exp(-beta E(s))
 *   local field  f_i = h_i + sum_j J_ij s_j
 *   heat bath    P(s_i = +1 | rest) = 1 / (1 + exp(2 beta f_i))
 *
 * Reproducible evaluation order (fp32, no FMA contraction in the field sum):
 *   f0_i = h_i;  for k in row(i) ascending neighbour position: f0_i -= J_ik
 *   f    = f0_i; for k in row(i) ascending neighbour position:
 *                    if s_k == +1: f += (2 * J_ik)
 *   x    = clamp(f * coef, -120, 120)         coef = (float)(2 beta log2 e)
 *   e    = exp2_poly(x)                        (degree-5, below; ~2.4e-7 rel)
 *   s_i  = +1  iff  fmaf(v, e, v) < 1.0f       v in (0,1), 23-bit uniform
 *
 * Uniforms: either supplied by the caller (float in (0,1)) or drawn from
 * Philox4x32-10 with
 *   key      = (seed_lo, seed_hi)
 *   counter  = (visit position, global chain id >> 3, sweep index, stream)
 *   halfword = global chain id & 7   (halfword j = bits 16 (j & 1) .. +15 of output word j >> 1)
 *   m23      = (halfword of stream 0) << 7  |  (halfword of stream 2) >> 9
 *   v        = as_float(m23 | 0x3f800000) - 1.0f + 2^-24   (exact; v in [2^-24, 1 - 2^-24])
 * i.e. one Philox call carries the 16 high bits of the uniforms of 8 chains and a second
 * stream carries the 7 low bits.  An implementation may decide from the high bits alone
 * whenever the low bits cannot change the outcome (the sm_100a kernel does; the oracle
 * always forms the full v) -- the contract is the decision with the full 23-bit v.
 * stream 0 = sweep uniforms (high 16 bits), stream 2 = sweep uniforms (low 7 bits),
 * stream 1 = initial state: counter = (visit position, global chain id >> 2, 0, 1),
 *            word = global chain id & 3, bit 31 set -> +1.
 * The result therefore does not depend on launch geometry or GPU count.
 */
#ifndef B200GRBM_SPEC_H
#define B200GRBM_SPEC_H

#include <stdint.h>

#define B200GRBM_PHILOX_M0 0xD2511F53u
#define B200GRBM_PHILOX_M1 0xCD9E8D57u
#define B200GRBM_PHILOX_W0 0x9E3779B9u
#define B200GRBM_PHILOX_W1 0xBB67AE85u
#define B200GRBM_PHILOX_ROUNDS 10

#define B200GRBM_STREAM_SWEEP 0u
#define B200GRBM_STREAM_INIT 1u
#define B200GRBM_STREAM_SWEEP_LO 2u

/* exp2 on [-0.5, 0.5], Remez fit of relative error, coefficients rounded to fp32 */
#define B200GRBM_EXP2_C0 0x1.000002p+0f
#define B200GRBM_EXP2_C1 0x1.62e428p-1f
#define B200GRBM_EXP2_C2 0x1.ebf918p-3f
#define B200GRBM_EXP2_C3 0x1.c6b6e4p-5f
#define B200GRBM_EXP2_C4 0x1.3d0c52p-7f
#define B200GRBM_EXP2_C5 0x1.5c08e6p-10f
#define B200GRBM_EXP2_MAGIC 12582912.0f /* 1.5 * 2^23 */
#define B200GRBM_EXP2_CLAMP 120.0f
#define B200GRBM_UNIFORM_HALF_ULP 0x1.0p-24f

/*
 * NOT part of the contract: constants of the sm_100a kernel's evaluation strategy (csrc/gibbs.cu decide_quick),
 * here so that the oracle's self-test of the bracket argument uses the same numbers.  With the 16-bit midpoint
 * v_m and e~ = MUFU.EX2(x), g = 1 + e~, d = v_m g - 1, the sign of d is the contract's decision whenever
 * |d| > K1 g + K2  (K1 = 2^-17 (1 + 2^-10), K2 = 2^-17); otherwise the contract arithmetic is evaluated.
 */
#define B200GRBM_LAZY_K1 0x1.004p-17f
#define B200GRBM_LAZY_K2 0x1.0p-17f

#endif /* B200GRBM_SPEC_H */
