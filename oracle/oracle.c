/*
 * oracle.c -- CPU restatement of the GRBM negative-phase path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this file's shared object.  Nothing under image-generation_b200/
 * imports, links or executes it; the product path fails loudly without its CUDA
 * library instead of falling back to this code.
 *
 * PARITY UNPINNED.  The reference (dwave-examples/image-generation) keeps the
 * arithmetic of this path in two un-vendored, un-installed third-party packages
 * (requirements.txt:3-4: dwave-ocean-sdk~=9.0 -> dimod / dwave-samplers,
 * dwave-pytorch-plugin~=0.3) and its tests/ directory is empty, so there are no
 * reference golden vectors for the sampler.  What this file is pinned against:
 *   - the Random123 known-answer vectors for Philox4x32-10 (tests/test_oracle.py),
 *   - exact Boltzmann enumeration on small graphs (tests/test_oracle.py),
 *   - the energy formula static/eq6.png applied to the in-tree checkpoints
 *     models/<QPU>/grbm.pth (tests/golden/).
 *
 * Functions and the reference sites they restate:
 *   oracle_gibbs / oracle_gibbs_f64 : sampler.sample_ising(h, J, num_reads, ...)
 *        called through grbm.sample at src/model_wrapper.py:309-316, :369-376 and
 *        src/utils/persistent_qpu_sampler.py:71-78; classical stand-in semantics per
 *        dwave-samplers' SA with proposal_acceptance_criteria="Gibbs": for each sweep
 *        (one beta per sweep), visit i = 0..N-1 in order, resample s_i from its
 *        conditional.
 *   oracle_energies : GraphRestrictedBoltzmannMachine.forward, call site src/losses.py:61.
 *   oracle_edge_stats : the sufficient statistics behind the gradient of
 *        src/losses.py:61  (d/dh_i = <s_i>_data - <s_i>_model, d/dJ_ij = <s_i s_j>...).
 *
 * Build: oracle/build.sh (gcc -O2 -ffp-contract=off -mfma -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b200grbm_spec.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------------------------------------------------------- Philox4x32-10 */

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < B200GRBM_PHILOX_ROUNDS; ++r) {
        uint64_t p0 = (uint64_t)B200GRBM_PHILOX_M0 * c0;
        uint64_t p1 = (uint64_t)B200GRBM_PHILOX_M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += B200GRBM_PHILOX_W0;
        k1 += B200GRBM_PHILOX_W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline float as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static inline uint32_t as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* v from the 23 mantissa bits m23 (spec header): as_float(m23 | 0x3f800000) - 1 + 2^-24, exact */
float oracle_uniform_from_m23(uint32_t m23)
{
    float one_to_two = as_float((m23 & 0x7fffffu) | 0x3f800000u);
    return (one_to_two - 1.0f) + B200GRBM_UNIFORM_HALF_ULP;
}

/* exp2 by the contract polynomial: every operation is a single IEEE fp32 op */
float oracle_exp2_poly(float x)
{
    x = fminf(fmaxf(x, -B200GRBM_EXP2_CLAMP), B200GRBM_EXP2_CLAMP);
    volatile float t = x + B200GRBM_EXP2_MAGIC; /* volatile: forbid algebraic folding of t - magic */
    float n = t - B200GRBM_EXP2_MAGIC;
    float r = x - n;
    float p = B200GRBM_EXP2_C5;
    p = fmaf(p, r, B200GRBM_EXP2_C4);
    p = fmaf(p, r, B200GRBM_EXP2_C3);
    p = fmaf(p, r, B200GRBM_EXP2_C2);
    p = fmaf(p, r, B200GRBM_EXP2_C1);
    p = fmaf(p, r, B200GRBM_EXP2_C0);
    return as_float(as_uint(p) + (as_uint(t) << 23));
}

/* returns 1 when the heat-bath draw sets the spin to +1 */
int oracle_accept(float f, float coef, float v)
{
    float e = oracle_exp2_poly(f * coef);
    return fmaf(v, e, v) < 1.0f;
}

/*
 * Self-test of the bracket argument behind the sm_100a kernel's lazy acceptance (csrc/gibbs.cu decide_quick,
 * DESIGN.md section 3) -- CPU only, no GPU involved: the quick decision is formed from the 16 high bits of the
 * uniform and an ADVERSARIAL exp2, e~ = 2^x (1 + rel), and compared with the contract's decision for every value
 * of the 7 low bits.  Returns 0/1 when the mark says the decision is certain, -1 when it defers to the contract.
 */
int oracle_bracketed_decision(float f, float coef, uint32_t hw, double rel)
{
    const float x = f * coef;                              /* not clamped on the fast side (Philox modes) */
    double et = exp2((double)x) * (1.0 + rel);
    float e = et > 3.0e38 ? INFINITY : (float)et;
    if (e < 1.17549435e-38f) e = 0.0f;                     /* ex2.approx.ftz flushes subnormal results */
    const float g = e + 1.0f;
    const float vm = ((float)hw + 0.5f) * 0x1.0p-16f;      /* exact */
    const float d = fmaf(vm, g, -1.0f);
    const float m = fmaf(g, -B200GRBM_LAZY_K1, fabsf(d) - B200GRBM_LAZY_K2);
    if (!isnan(m) && signbit(m)) return -1;                /* the GPU's NaN (inf - inf) is 0x7fffffff: sign clear */
    return (!isnan(d) && signbit(d)) ? 1 : 0;
}

/* n random (f, beta, hw) triples, half of them with hw placed next to the acceptance threshold; every decision the
 * mark calls certain must equal the contract's for lo7 = 0, 127 and a random value.  Returns the number of
 * violations; *deferred receives the number of deferred decisions. */
int64_t oracle_bracket_selftest(int64_t n, uint64_t seed, double rel, int64_t *deferred)
{
    int64_t bad = 0, unsure = 0;
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (int64_t i = 0; i < n; ++i) {
        uint32_t ctr[4] = { (uint32_t)i, (uint32_t)(i >> 32), 7u, 11u }, r[4];
        oracle_philox4x32_10(ctr, key, r);
        const float u0 = (float)(r[0] >> 8) * 0x1.0p-24f, u1 = (float)(r[1] >> 8) * 0x1.0p-24f;
        const float f = (u0 - 0.5f) * 24.0f;                                   /* fields in [-12, 12] */
        const float beta = 0.02f * exp2f(u1 * 10.0f);                          /* beta in [0.02, 20]: |x| up to ~700 */
        const float coef = (float)(2.0 * (double)beta * 1.4426950408889634);
        uint32_t hw = r[2] & 0xffffu;
        if (i & 1) {                                                           /* next to the threshold 1 / (1 + 2^x) */
            const double p = 1.0 / (1.0 + exp2((double)(f * coef)));
            long t = (long)floor(p * 65536.0) + (long)(r[2] >> 16) % 5 - 2;
            hw = (uint32_t)(t < 0 ? 0 : t > 65535 ? 65535 : t);
        }
        const int q = oracle_bracketed_decision(f, coef, hw, rel);
        if (q < 0) { ++unsure; continue; }
        const uint32_t lows[3] = { 0u, 127u, r[3] & 127u };
        for (int k = 0; k < 3; ++k)
            if (oracle_accept(f, coef, oracle_uniform_from_m23((hw << 7) | lows[k])) != q) ++bad;
    }
    if (deferred) *deferred = unsure;
    return bad;
}

static inline uint32_t halfword(const uint32_t out[4], unsigned j) { return (out[j >> 1] >> (16u * (j & 1u))) & 0xffffu; }

/*
 * Sweep uniforms of the 8 chains of block `blk8` (global chain id >> 3) at (pos, sweep):
 * high 16 bits from stream 0, low 7 bits from stream 2 (spec header).
 */
void oracle_sweep_uniforms8(uint64_t seed, uint32_t pos, uint32_t sweep, uint32_t blk8, float v[8])
{
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t ctr[4] = { pos, blk8, sweep, B200GRBM_STREAM_SWEEP };
    uint32_t hi[4], lo[4];
    oracle_philox4x32_10(ctr, key, hi);
    ctr[3] = B200GRBM_STREAM_SWEEP_LO;
    oracle_philox4x32_10(ctr, key, lo);
    for (unsigned j = 0; j < 8; ++j) v[j] = oracle_uniform_from_m23((halfword(hi, j) << 7) | (halfword(lo, j) >> 9));
}

float oracle_sweep_uniform(uint64_t seed, uint32_t pos, uint32_t sweep, uint64_t chain)
{
    float v[8];
    oracle_sweep_uniforms8(seed, pos, sweep, (uint32_t)(chain >> 3), v);
    return v[chain & 7];
}

/* Initial state from Philox stream 1: bit 31 of the chain's word set -> +1. */
void oracle_init_state(int n, int chains, int8_t *state, uint64_t seed, uint64_t chain_offset)
{
#pragma omp parallel for schedule(static)
    for (int c = 0; c < chains; ++c) {
        for (int p = 0; p < n; ++p) {
            const uint64_t chain = chain_offset + (uint64_t)c;
            uint32_t ctr[4] = { (uint32_t)p, (uint32_t)(chain >> 2), 0u, B200GRBM_STREAM_INIT };
            uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
            uint32_t out[4];
            oracle_philox4x32_10(ctr, key, out);
            state[(size_t)c * n + p] = (out[chain & 3] >> 31) ? 1 : -1;
        }
    }
}

/*
 * Sequential heat-bath Gibbs, contract arithmetic (fp32).  Everything is in
 * "visit position" space: row p of the CSR is the p-th spin visited in a sweep and
 * col[] holds visit positions, sorted ascending within a row.  Jd[] is the coupling of
 * each directed CSR entry.  coef[t] = (float)(2 beta_t log2 e).
 * uniforms == NULL -> Philox; else uniforms[(t*chains + c)*n + p] in (0,1).
 */
void oracle_gibbs(int n, const int32_t *rowptr, const int32_t *col, const float *Jd, const float *h,
                  int chains, int8_t *state, int num_sweeps, const float *coef, const float *uniforms,
                  uint64_t seed, uint64_t chain_offset, uint32_t sweep_offset)
{
    float *f0 = (float *)malloc(sizeof(float) * (size_t)n);
    for (int p = 0; p < n; ++p) {
        float a = h[p];
        for (int k = rowptr[p]; k < rowptr[p + 1]; ++k) a = a - Jd[k];
        f0[p] = a;
    }
    /* chains are walked in Philox blocks of 8 (global chain id >> 3) so that the two calls behind the
     * uniforms of (position, sweep) are made once per block, as on the GPU */
    const int64_t first_blk = (int64_t)(chain_offset >> 3), last_blk = (int64_t)((chain_offset + chains - 1) >> 3);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = first_blk; b <= last_blk; ++b) {
        const int64_t g0 = b * 8 > (int64_t)chain_offset ? b * 8 : (int64_t)chain_offset;
        const int64_t g1 = b * 8 + 8 < (int64_t)chain_offset + chains ? b * 8 + 8 : (int64_t)chain_offset + chains;
        for (int t = 0; t < num_sweeps; ++t) {
            const float cf = coef[t];
            for (int p = 0; p < n; ++p) {
                float v8[8];
                if (!uniforms) oracle_sweep_uniforms8(seed, (uint32_t)p, sweep_offset + (uint32_t)t, (uint32_t)b, v8);
                for (int64_t gc = g0; gc < g1; ++gc) {
                    const int64_t c = gc - (int64_t)chain_offset;
                    int8_t *s = state + (size_t)c * n;
                    float f = f0[p];
                    for (int k = rowptr[p]; k < rowptr[p + 1]; ++k)
                        if (s[col[k]] > 0) f = f + 2.0f * Jd[k];
                    const float v = uniforms ? uniforms[((size_t)t * chains + c) * n + p] : v8[gc & 7];
                    s[p] = oracle_accept(f, cf, v) ? 1 : -1;
                }
            }
        }
    }
    free(f0);
}

/*
 * Textbook double-precision heat bath (what a reference-style C++ annealer computes,
 * SURVEY.md Appendix A.4 'Gibbs' rule): f = h + sum J s, P(+1) = 1/(1+exp(2 beta f)),
 * accept u < P.  Same uniforms as oracle_gibbs so trajectories can be compared
 * decision by decision; used to show that the fp32 contract does not change the law.
 */
void oracle_gibbs_f64(int n, const int32_t *rowptr, const int32_t *col, const float *Jd, const float *h,
                      int chains, int8_t *state, int num_sweeps, const double *beta, const float *uniforms,
                      uint64_t seed, uint64_t chain_offset, uint32_t sweep_offset)
{
    const int64_t first_blk = (int64_t)(chain_offset >> 3), last_blk = (int64_t)((chain_offset + chains - 1) >> 3);
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = first_blk; b <= last_blk; ++b) {
        const int64_t g0 = b * 8 > (int64_t)chain_offset ? b * 8 : (int64_t)chain_offset;
        const int64_t g1 = b * 8 + 8 < (int64_t)chain_offset + chains ? b * 8 + 8 : (int64_t)chain_offset + chains;
        for (int t = 0; t < num_sweeps; ++t) {
            for (int p = 0; p < n; ++p) {
                float v8[8];
                if (!uniforms) oracle_sweep_uniforms8(seed, (uint32_t)p, sweep_offset + (uint32_t)t, (uint32_t)b, v8);
                for (int64_t gc = g0; gc < g1; ++gc) {
                    const int64_t c = gc - (int64_t)chain_offset;
                    int8_t *s = state + (size_t)c * n;
                    double f = (double)h[p];
                    for (int k = rowptr[p]; k < rowptr[p + 1]; ++k) f += (double)Jd[k] * (double)s[col[k]];
                    const float v = uniforms ? uniforms[((size_t)t * chains + c) * n + p] : v8[gc & 7];
                    const double pplus = 1.0 / (1.0 + exp(2.0 * beta[t] * f));
                    s[p] = ((double)v < pplus) ? 1 : -1;
                }
            }
        }
    }
}

/* E(s) = sum_i h_i s_i + sum_e J_e s_i s_j, double accumulation, one value per row of `state`. */
void oracle_energies(int n, int n_edges, const int32_t *ei, const int32_t *ej, const float *J, const float *h,
                     int rows, const int8_t *state, double *out)
{
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; ++r) {
        const int8_t *s = state + (size_t)r * n;
        double e = 0.0;
        for (int i = 0; i < n; ++i) e += (double)h[i] * s[i];
        for (int k = 0; k < n_edges; ++k) e += (double)J[k] * (s[ei[k]] * s[ej[k]]);
        out[r] = e;
    }
}

/* same for real-valued rows (encoder spins are only approximately +-1) */
void oracle_energies_f32(int n, int n_edges, const int32_t *ei, const int32_t *ej, const float *J, const float *h,
                         int rows, const float *x, double *out)
{
#pragma omp parallel for schedule(static)
    for (int r = 0; r < rows; ++r) {
        const float *s = x + (size_t)r * n;
        double e = 0.0;
        for (int i = 0; i < n; ++i) e += (double)h[i] * (double)s[i];
        for (int k = 0; k < n_edges; ++k) e += (double)J[k] * ((double)s[ei[k]] * (double)s[ej[k]]);
        out[r] = e;
    }
}

/* integer sufficient statistics: sum_r s_i and sum_r s_i s_j */
void oracle_edge_stats(int n, int n_edges, const int32_t *ei, const int32_t *ej, int rows, const int8_t *state,
                       int64_t *sum_s, int64_t *sum_ss)
{
    memset(sum_s, 0, sizeof(int64_t) * (size_t)n);
    memset(sum_ss, 0, sizeof(int64_t) * (size_t)n_edges);
    for (int r = 0; r < rows; ++r) {
        const int8_t *s = state + (size_t)r * n;
        for (int i = 0; i < n; ++i) sum_s[i] += s[i];
        for (int k = 0; k < n_edges; ++k) sum_ss[k] += s[ei[k]] * s[ej[k]];
    }
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_num_threads(int t)
{
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}
