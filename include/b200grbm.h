/*
 * b200grbm.h -- C ABI of libb200grbm.so, the sm_100a implementation of the GRBM
 * negative-phase hot path of dwave-examples/image-generation.
 *
 * The reference is pure Python; the "FFI" a maintainer would bind is therefore ctypes
 * (INTEGRATION.md shows the stub).  Each entry point cites the reference interface whose
 * arithmetic it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, a negative B200GRBM_E* code for argument /
 *     shape errors, or a positive cudaError_t; b200grbm_last_error() returns a
 *     thread-local message for the last non-zero return.  No exceptions cross the ABI.
 *   - all *_dev pointers are device pointers owned by the caller (PyTorch in the host
 *     layer); the library never allocates or frees persistent device memory -- the one exception is the
 *     explicit exchange-buffer pair b200grbm_peer_alloc / b200grbm_peer_free (memory that must be exportable
 *     to the other ranks of the box), which the caller owns between the two calls.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing
 *     synchronises the device.
 *   - there is no CPU fallback: on a machine without an sm_100 device every compute entry
 *     point returns an error.
 */
#ifndef B200GRBM_H
#define B200GRBM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200GRBM_ABI_VERSION 5

#define B200GRBM_EINVAL (-1)   /* bad argument / shape */
#define B200GRBM_EUNSUPPORTED (-2) /* configuration not compiled in (e.g. chains_per_lane) */
#define B200GRBM_ENODEVICE (-3) /* no sm_100 device */

/* acceptance rule of the heat-bath draw (include/b200grbm_spec.h) */
#define B200GRBM_ACCEPT_EXACT 0 /* contract polynomial: bit-reproducible on the CPU oracle */
#define B200GRBM_ACCEPT_FAST 1  /* MUFU.EX2: same law, statistical parity only */

typedef struct b200grbm_ell_entry {
    uint32_t j2_bits; /* fp32 bits of 2*J_eff for this slot (0 for padding) */
    uint32_t nbr;     /* b200grbm_sweep_state_offset(n_tiles) + 4 * (neighbour visit position) */
} b200grbm_ell_entry;

const char *b200grbm_last_error(void);
int32_t b200grbm_abi_version(void);

/* sm count / compute capability / opt-in shared memory of the current device */
int32_t b200grbm_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor, int32_t *smem_optin);

/*
 * Sampler tables ("tiles").  The sweep kernel streams, per colour round, one contiguous tile
 * of (1 + ell_width) x threads 8-byte entries through shared memory with the bulk-copy engine:
 *   tile[0][lane]                        : { fp32 bits of f0 = h_eff - sum_k J_eff, unused }
 *   tile[1 + k][lane]      k < ell_width : { fp32 bits of 2*J_eff, byte offset of the neighbour's state word
 *                                            in the CTA's dynamic shared memory }
 * (ell_width = max degree, padded with zero slots to a multiple of 4 when chains_per_lane <= 8)
 * where lane = (visit position - first position of the round).  tile_info[t] = { first visit
 * position, number of spins } of round t; rounds never straddle a colour boundary.  The host
 * layer builds the .nbr fields and tile_info once per (graph, threads); the values are written
 * by b200grbm_set_weights.
 *
 * b200grbm_set_weights: effective Ising parameters from the GRBM parameters.  Replaces the
 * parameter preparation inside GraphRestrictedBoltzmannMachine.sample (third-party; call
 * sites src/model_wrapper.py:309-316, :369-376, src/utils/persistent_qpu_sampler.py:71-78):
 * h = clip(prefactor * linear, linear_range), J = clip(prefactor * quadratic, quadratic_range)
 * -- on device, so a training step never round-trips the parameters through Python dicts.
 *   slot_a/slot_b [n_edges]  flat entry index of the two tile slots carrying edge e
 *   row_base      [n]        flat entry index of the f0 entry of visit position p (slot k is
 *                            row_base[p] + (1 + k) * threads)
 *   h_eff_dev [n] node order (optional);  j_eff_dev [n_edges] (optional)
 */
int32_t b200grbm_set_weights(const float *linear_dev, const float *quadratic_dev, int32_t n, int32_t n_edges,
                             float prefactor, float h_lo, float h_hi, float j_lo, float j_hi,
                             const int32_t *order_dev, const int32_t *slot_a_dev, const int32_t *slot_b_dev,
                             const int32_t *row_base_dev, int32_t ell_width, int32_t threads,
                             b200grbm_ell_entry *tiles_dev, float *h_eff_dev, float *j_eff_dev, void *stream);

typedef struct b200grbm_sweep_args {
    uint32_t struct_size;      /* sizeof(b200grbm_sweep_args) */
    int32_t n;                 /* spins */
    int32_t n_pad;             /* row pitch of packed state */
    int32_t ell_width;         /* neighbour slots per spin (max degree) */
    int32_t n_tiles;           /* colour rounds per sweep */
    const b200grbm_ell_entry *tiles_dev; /* [n_tiles][1 + ell_width][threads], 16-byte aligned */
    const int32_t *tile_info_dev;        /* [n_tiles][2] */
    const int32_t *order_dev;  /* [n] node visited at position p (for int8 node-order I/O) */
    int32_t chains;            /* chains in this call */
    int32_t chains_per_lane;   /* 4, 8, 16, 24, 28 or 32: chains bit-packed per state word */
    int32_t threads;           /* CTA size the tiles were built for, multiple of 32 in [64, 768] */
    int32_t accept;            /* B200GRBM_ACCEPT_* */
    uint64_t chain_offset;     /* global id of chain 0 of this call (multiple of 4; sharded runs use multiples of 8) */
    uint64_t seed;
    uint32_t sweep_offset;     /* Philox sweep counter of the first sweep */
    int32_t num_sweeps;
    const float *coef_dev;     /* [num_sweeps] (float)(2 beta log2 e) */
    const float *uniforms_dev; /* NULL -> Philox; else [num_sweeps][chains][n] visit-position order */
    const int8_t *state_in_dev;   /* [chains][n] node order, +-1; NULL -> packed_in or random init */
    const uint32_t *packed_in_dev; /* [groups][n_pad] visit-position order, bit c = chain c; NULL -> random init */
    int8_t *state_out_dev;     /* optional */
    uint32_t *packed_out_dev;  /* optional */
} b200grbm_sweep_args;

/*
 * sampler.sample_ising(h, J, num_reads=chains, ...) -- the call GraphRestrictedBoltzmannMachine
 * .sample makes (boundary: src/utils/common.py:123-138 builds the sampler and its kwargs).
 * Runs num_sweeps colour-blocked heat-bath sweeps on `chains` independent chains.
 *
 * Kernel choice (results are identical whichever runs; b200grbm_last_sweep_kernel tells which did):
 *   chains_per_lane == 4 and a small graph (<= 5 rounds of <= 256 spins, degree <= 20)  -> gibbs_small_kernel
 *   chains_per_lane == 28, Philox uniforms, streamed tables of width 15 (Pegasus) or 20 (Zephyr) whose pre-drawn
 *     uniforms fit behind the tile stages: compile-time CTA size for (640, 15) [P16], (480, 20) and (384, 20) with
 *     >= 2 groups per SM [Z15]; any other CTA size as a run-time value                             -> gibbs_wide_kernel
 *   anything else                                                                                  -> gibbs_kernel
 *     (small graphs whose tables stay resident in shared memory, 28 chains per lane, many groups: several chain groups
 *      share one CTA and one copy of the tables; B200GRBM_GPC=n forces n groups per CTA)
 * Environment switches for A/B measurements and the parity tests of the variants: B200GRBM_SMALL=0, B200GRBM_WIDE=0
 * (fall back to gibbs_kernel), B200GRBM_MMD_TILE=1|2, B200GRBM_GEMM_TILE=1|2 (single-CTA / CTA-pair tensor-core kernels).
 * Read by the host layer: B200GRBM_MMD_FP4=0|1 (forward Gram on int8 / packed e2m1 operands; default: e2m1 from 2048 rows
 * up), B200GRBM_MMD_EXCHANGE=p2p|bits|int8 (row exchange of the sharded MMD, image_generation_b200/dist.py).
 */
int32_t b200grbm_gibbs_sweeps(const b200grbm_sweep_args *args, void *stream);

/* byte offset of state word 0 inside a sweep CTA's dynamic shared memory (2 mbarriers + round table before it);
 * the .nbr fields of the tiles are  this + 4 * position,  so a neighbour read is one LDS with no address arithmetic */
int32_t b200grbm_sweep_state_offset(int32_t n_tiles);

/* dynamic shared memory a sweep launch needs (round table + state + 2 tile stages); must fit the device's opt-in limit */
int64_t b200grbm_sweep_smem_bytes(int32_t n, int32_t ell_width, int32_t threads, int32_t n_tiles);

/*
 * Measurement aid for the sweep kernel's lazy acceptance: largest relative error of
 * ex2.approx.ftz.f32 against double-precision exp2 over n evenly spaced fp32 arguments of [x_lo, x_hi]
 * (normal results only).  max_rel_err_out is a HOST pointer; synchronises the stream.
 */
int32_t b200grbm_ex2_probe(float x_lo, float x_hi, int64_t n, double *max_rel_err_out, void *stream);

/* number of kernel launches the last b200grbm_gibbs_sweeps call on this thread enqueued */
int32_t b200grbm_last_launch_count(void);
/* which sweep kernel that call chose: 0 = chains bit-packed per lane (throughput), 1 = one chain per lane column with
 * every round's table in registers (small problems: chains_per_lane == 4, <= 5 colour rounds of <= 256 spins,
 * degree <= 20, few enough chains for all CTAs to be resident at once), 2 = the specialised throughput kernel (28 chains
 * per lane, compile-time CTA size and table width, pre-drawn uniforms behind a split round barrier; same results as 0) */
int32_t b200grbm_last_sweep_kernel(void);

/*
 * Sign-pack rows of real-valued spins (encoder output, src/model_wrapper.py:297,318) or int8
 * samples into bit-packed words: packed[g][pos[i]] bit c = (x[g*cpl + c][i] > 0), pos = inverse of order.
 * Also the int8 conversion GraphRestrictedBoltzmannMachine.sampleset_to_tensor undoes
 * (src/losses.py:59).
 */
int32_t b200grbm_pack_f32(const float *x_dev, int32_t rows, int32_t n, int32_t n_pad, const int32_t *pos_dev,
                          int32_t chains_per_lane, uint32_t *packed_dev, void *stream);
int32_t b200grbm_pack_i8(const int8_t *x_dev, int32_t rows, int32_t n, int32_t n_pad, const int32_t *pos_dev,
                         int32_t chains_per_lane, uint32_t *packed_dev, void *stream);

/*
 * Integer sufficient statistics of packed rows: sum_s[i] += sum_r s_ri (node order),
 * sum_ss[e] += sum_r s_ri s_rj.  These are the gradients of src/losses.py:61
 * (mean(grbm(spins)) - mean(grbm(samples))) wrt linear / quadratic up to the 1/rows factor.
 * Accumulates (+=) into int64 counters so ranks / batches can be combined exactly.
 */
int32_t b200grbm_edge_stats(const uint32_t *packed_dev, int32_t rows, int32_t chains_per_lane, int32_t n,
                            int32_t n_pad, int32_t n_edges, const int32_t *edge_pi_dev, const int32_t *edge_pj_dev,
                            const int32_t *order_dev, int64_t *sum_s_dev, int64_t *sum_ss_dev, void *stream);

/*
 * GraphRestrictedBoltzmannMachine.forward (call site src/losses.py:61):
 * energy[r] = sum_i linear_i x_ri + sum_e quadratic_e x_r,i(e) x_r,j(e)   (fp32 in, fp32 out,
 * fp32 pairwise accumulation per row).  backward: grad_linear[i] += sum_r g_r x_ri,
 * grad_quadratic[e] += sum_r g_r x_ri x_rj   (accumulating: the caller zeroes the outputs).
 */
int32_t b200grbm_energy_forward(const float *x_dev, int32_t rows, int32_t n, int32_t n_edges,
                                const int32_t *edge_i_dev, const int32_t *edge_j_dev, const float *linear_dev,
                                const float *quadratic_dev, float *energy_dev, void *stream);
int32_t b200grbm_energy_backward(const float *x_dev, const float *grad_energy_dev, int32_t rows, int32_t n,
                                 int32_t n_edges, const int32_t *edge_i_dev, const int32_t *edge_j_dev,
                                 float *grad_linear_dev, float *grad_quadratic_dev, void *stream);
/* gradient of the same energies wrt the input rows (the plugin's forward is an ordinary differentiable torch
 * expression in x): grad_x[r][i] = g_r * (linear_i + sum_{e ni i} quadratic_e x_r,other(e)) */
int32_t b200grbm_energy_grad_x(const float *x_dev, const float *grad_energy_dev, int32_t rows, int32_t n,
                               int32_t n_edges, const int32_t *edge_i_dev, const int32_t *edge_j_dev,
                               const float *linear_dev, const float *quadratic_dev, float *grad_x_dev, void *stream);
/* int8 +-1 rows -> fp64 energies (dimod SampleSet.record.energy, src/utils/persistent_qpu_sampler.py:84-88) */
int32_t b200grbm_energy_i8(const int8_t *s_dev, int32_t rows, int32_t n, int32_t n_edges,
                           const int32_t *edge_i_dev, const int32_t *edge_j_dev, const float *h_dev,
                           const float *j_dev, double *energy_dev, void *stream);
/*
 * The same energies from the sampler's bit-packed output (packed_out_dev of b200grbm_gibbs_sweeps:
 * [groups][n_pad] words in visit-position order, bit c = chain c of the group): one CTA per group, one pair
 * of state words per edge serves all chains_per_lane chains.  edge_pi/edge_pj = visit positions of the edge
 * ends, order[p] = node at position p (h_dev is in node order), fp64 accumulation.
 */
int32_t b200grbm_energy_packed(const uint32_t *packed_dev, int32_t rows, int32_t chains_per_lane, int32_t n,
                               int32_t n_pad, int32_t n_edges, const int32_t *edge_pi_dev,
                               const int32_t *edge_pj_dev, const int32_t *order_dev, const float *h_dev,
                               const float *j_dev, double *energy_dev, void *stream);

/*
 * maximum_mean_discrepancy_loss(x, y, GaussianKernel(n_kernels)) -- third-party
 * dwave-pytorch-plugin; call site src/model_wrapper.py:320 (kernel built :273); formulas
 * static/eq3.png / static/eq4.png (README.md:114-129); recollected code form SURVEY.md
 * Appendix A.3.  z_dev holds the stacked rows [x; y] (m = m_x + m_y rows, d features).
 *   t_ab  = ||z_a - z_b||  (squared != 0: ||.||^2)
 *   bw    = bandwidth > 0 ? bandwidth : sum_ab t_ab / (m^2 - m)
 *   k_ab  = sum_u exp(-t_ab / (bw * mul_factor^(u - n_kernels/2)))
 * forward writes sums_dev[4] = { sum_{a,b in x} k, sum_{a,b in y} k, sum_{a in x, b in y} k,
 * sum_ab t_ab } (diagonals included; the host layer forms the biased / unbiased estimate).
 * The fp32 entry points run on CUDA cores with directly accumulated differences (any real
 * input); the i8 entry point runs the Gram contraction of +-1 rows on tcgen05 tensor cores
 * (int8 x int8 -> int32, exact).
 */
int32_t b200grbm_mmd_forward_f32(const float *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t n_kernels,
                                 float mul_factor, int32_t squared, float bandwidth, double *sums_dev, void *stream);
/*
 * d(loss)/dx for loss = w_xx' * S_xx + ... : coef_dev [m_x][m] workspace receives
 * grad_out * w * (dk/dt) * (dt/d||.||) per pair (w = w_xx for x-x pairs, w_xy for x-y pairs,
 * both including the factor 2 of the symmetric sum), then
 * grad_x[a] = sum_b coef_ab (x_a - z_b).  sums_dev[3] must still hold the forward's distance sum.
 */
int32_t b200grbm_mmd_backward_f32(const float *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t n_kernels,
                                  float mul_factor, int32_t squared, float bandwidth, const double *sums_dev,
                                  float w_xx, float w_xy, const float *grad_out_dev, float *coef_dev,
                                  float *grad_x_dev, void *stream);

/*
 * tcgen05 path for +-1 rows (same contract as b200grbm_mmd_forward_f32; BASELINE.json cfg3).
 * z_dev: int8 [m][d_pad] row-major, 16-byte aligned, d_pad a multiple of 16 (128 for full TMA efficiency),
 * columns >= d zero (b200grbm_spin_extract_* / b200grbm_mmd_pack_i8 produce it from real-valued spins by sign).
 * ||a-b||^2 = 4 Hamming(a, b) = 2 (d - a.b) exactly, from the int32 Gram accumulators.
 *
 * b200grbm_mmd_hist_i8: ONE Gram pass over the upper triangle of 128 x 256 tiles; the epilogue counts the entries per
 * Hamming distance into hist_dev[3][d + 1] (uint64, ACCUMULATING; [0] ordered pairs inside x incl. the diagonal,
 * [1] inside y, [2] x-y pairs).  shard_rank / shard_world deal the tiles round-robin over ranks that each hold the
 * whole z: the per-rank histograms add up (int64 all-reduce) to the single-GPU histogram exactly -- this is the
 * "MMD partial sums combined by all-reduce" of the multi-GPU path.
 * b200grbm_mmd_eval_hist: sums_dev[5] from the histograms, float64, fixed reduction order: [0..3] as
 * b200grbm_mmd_forward_f32, [4] = scale * (E k(x,x') + E k(y,y') - 2 E k(x,y)) with the unbiased (diagonal dropped,
 * unbiased != 0) or biased block means -- the value maximum_mean_discrepancy_loss returns (scale = 1, or 1 / n_kernels
 * for the mean-of-kernels variant).  The data-dependent bandwidth comes from the same histograms, so the
 * auto-bandwidth forward needs one Gram pass, not two.
 * b200grbm_mmd_forward_i8 = memset + hist (all tiles) + eval; hist_dev is a workspace of 3 (d + 1) uint64.
 */
int32_t b200grbm_mmd_pack_i8(const float *z_dev, int32_t m, int32_t d, int32_t d_pad, int8_t *out_dev, void *stream);
int32_t b200grbm_mmd_hist_i8(const int8_t *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t d_pad, int32_t shard_rank,
                             int32_t shard_world, uint64_t *hist_dev, void *stream);
int32_t b200grbm_mmd_eval_hist(const uint64_t *hist_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t n_kernels,
                               float mul_factor, int32_t squared, float bandwidth, int32_t unbiased, double scale,
                               double *sums_dev, void *stream);
int32_t b200grbm_mmd_forward_i8(const int8_t *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t d_pad,
                                int32_t n_kernels, float mul_factor, int32_t squared, float bandwidth, int32_t unbiased,
                                double scale, uint64_t *hist_dev, double *sums_dev, void *stream);

/*
 * Fused spin extraction (src/model_wrapper.py:318: `spins.reshape(-1, n)` feeds the MMD at :320 and, detached, the
 * NLL at :332-342): one pass over real-valued (f32) or int8 rows writes, by sign, any of
 *   rows_dev   int8 [row_off + r][d_pad]     Gram operand (padding columns zeroed)
 *   zt_dev     int8 [i][row_off + r]         transposed copy, pitch zt_pitch (multiple of 16): B operand of the
 *                                            backward GEMM; columns >= the last row written are the caller's to zero
 *   packed_dev u32  [r / 32][pos[i]]         bit-packed words (32 rows per word, visit-position order) for
 *                                            b200grbm_edge_stats; positions >= d are the caller's to zero
 *   rows4_dev  e2m1 [row_off + r][row_bytes4] packed 4-bit Gram operand of b200grbm_mmd_hist_fp4 (two spins per byte,
 *                                            row_bytes4 a multiple of 16 -- 128 for the TMA boxes -- padding zeroed)
 * NULL skips an output.  nonspin_dev (optional, accumulating int32): number of 128 x 64 input tiles holding an entry
 * with | |x| - 1 | > tol -- lets the caller verify that sign-packing loses nothing before taking the int8 path.
 */
int32_t b200grbm_spin_extract_f32(const float *x_dev, int32_t rows, int32_t d, int8_t *rows_dev, int32_t d_pad,
                                  int32_t row_off, int8_t *zt_dev, int32_t zt_pitch, uint32_t *packed_dev,
                                  const int32_t *pos_dev, int32_t n_pad, int32_t *nonspin_dev, float tol,
                                  uint8_t *rows4_dev, int32_t row_bytes4, void *stream);
int32_t b200grbm_spin_extract_i8(const int8_t *x_dev, int32_t rows, int32_t d, int8_t *rows_dev, int32_t d_pad,
                                 int32_t row_off, int8_t *zt_dev, int32_t zt_pitch, uint32_t *packed_dev,
                                 const int32_t *pos_dev, int32_t n_pad, int32_t *nonspin_dev, float tol,
                                 uint8_t *rows4_dev, int32_t row_bytes4, void *stream);
/* out[c][r] = in[r][c] (int8), out pitch out_pitch >= rows, columns r >= rows zero-filled */
int32_t b200grbm_transpose_i8(const int8_t *in_dev, int32_t rows, int32_t cols, int32_t in_pitch, int8_t *out_dev,
                              int32_t out_pitch, void *stream);

/*
 * Backward of the +-1 MMD on int8 tensor cores (dvae_loss.backward() through the MMD term,
 * src/model_wrapper.py:320-326):  grad_x[a] = g (rowsum_a(A) x_a - (A Z)_a),  A_ab = w (dk/dt)(dt/d||.||)/||.||
 * (w = w_xx for x-x pairs, w_xy for x-y pairs, both including the factor 2 of the symmetric sum; 0 on the diagonal).
 * b200grbm_mmd_coef_i8 re-runs the int8 Gram over rows [row0, row0 + n_rows) of the x block against every column and
 * writes A as n_planes (2 or 3) signed base-256 digit planes of a fixed-point number -- planes_dev
 * [n_planes][rows_alloc][m_pad] int8, most significant plane first, m_pad a multiple of 128 >= m, padding columns
 * zero -- plus the exact integer row sums (rowsum_dev [n_rows]) and the value of one fixed-point unit (scale_dev,
 * device scalar).  lut_dev: workspace of d + 1 floats.  sums_dev[3] is the forward's distance sum.  hist_dev
 * (optional): the forward's histograms -- the fixed-point range then spans only the distances that occur (coefficients
 * of other distances, if any, saturate).
 * b200grbm_mmd_grad_i8 contracts the planes with zt_dev (Z transposed, [d][m_pad] int8) on tcgen05.mma.kind::i8 and
 * writes grad_x_dev [n_rows][d] = grad_out * scale * (rowsum_a z_ai - sum_b q_ab z_bi); z_dev row z_row0 + a is x_a.
 */
int32_t b200grbm_mmd_coef_i8(const int8_t *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t d_pad, int32_t row0,
                             int32_t n_rows, int32_t n_kernels, float mul_factor, int32_t squared, float bandwidth,
                             const double *sums_dev, const uint64_t *hist_dev, float w_xx, float w_xy, float *lut_dev,
                             int8_t *planes_dev, int32_t n_planes, int32_t rows_alloc, int32_t m_pad, int64_t *rowsum_dev,
                             double *scale_dev, void *stream);
int32_t b200grbm_mmd_grad_i8(const int8_t *planes_dev, int32_t n_planes, int32_t n_rows, int32_t rows_alloc, int32_t m_pad,
                             const int8_t *zt_dev, int32_t d, const int64_t *rowsum_dev, const double *scale_dev,
                             const float *grad_out_dev, const int8_t *z_dev, int32_t z_row0, int32_t d_pad,
                             float *grad_x_dev, void *stream);
/*
 * C[M][ldc] (fp32) = (A_hi + A_lo)[M][K] * B[N][K]^T with bf16 operands (tcgen05.mma.kind::f16); a_lo_dev may be
 * NULL; a_rows_alloc = rows allocated in the A buffers.  Used by the backward of the continuous (bf16) MMD path.
 */
int32_t b200grbm_gemm_bf16_tn(const void *a_hi_dev, const void *a_lo_dev, int32_t M, int32_t K, int32_t a_rows_alloc,
                              const void *b_dev, int32_t N, float *c_dev, int32_t ldc, void *stream);

/*
 * Continuous (real-valued) rows on tensor cores: bf16 Gram with fp32 accumulation (north star:
 * "bf16/fp32-accumulate for continuous encoder latents").  z_hi_dev: bf16 [m][k_pad] (k_pad a
 * multiple of 64, zero padded); z_lo_dev: optional bf16 residual z - hi for the split-bf16 form
 * (hi.hi + hi.lo + lo.hi, ~2^-16 relative error per product) or NULL; norms_dev: fp32 squared norms
 * of the rounded rows.  Same sums_dev contract as b200grbm_mmd_forward_f32.
 */
int32_t b200grbm_mmd_forward_bf16(const void *z_hi_dev, const void *z_lo_dev, const float *norms_dev, int32_t m_x,
                                  int32_t m_y, int32_t k_pad, int32_t n_kernels, float mul_factor, int32_t squared,
                                  float bandwidth, double *sums_dev, void *stream);

/*
 * Tensor-core peak probe (measurement aid, not on the product path): back-to-back tcgen05.mma
 * (kind 0 = int8, 1 = bf16; M128 x N256) from resident shared-memory operands on every SM.
 * Writes 2 * MAC / s to the HOST pointer ops_per_s_out and synchronises the stream.
 */
int32_t b200grbm_tensor_peak(int32_t kind, int32_t iters, double *ops_per_s_out, void *stream);

/*
 * Backward coefficients for real-valued rows on the tcgen05 bf16 kernel (counterpart of
 * b200grbm_mmd_coef_i8): A_ab = w * (dk/dt)(dt/d||.||)/||.|| as a bf16 (hi, lo) pair for the x rows,
 * then grad_x = rowsum(A) x - A Z through b200grbm_gemm_bf16_tn.  sums_dev[3] = the forward's distance sum.
 */
int32_t b200grbm_mmd_coef_bf16(const void *z_hi_dev, const void *z_lo_dev, const float *norms_dev, int32_t m_x,
                               int32_t m_y, int32_t k_pad, int32_t n_kernels, float mul_factor, int32_t squared,
                               float bandwidth, const double *sums_dev, float w_xx, float w_xy, void *coef_hi_dev,
                               void *coef_lo_dev, int32_t m_pad, void *stream);

/*
 * FP4 form of the forward Gram pass: +-1 (and the zero padding) are exact in e2m1, so the same Hamming histograms come
 * from tcgen05.mma.kind::mxf4 (block scale factors all 2^0, fp32 accumulation of +-1 products is exact) at half the
 * operand bytes and twice the tensor rate of int8.  b200grbm_pack_fp4_i8 converts the int8 rows [m][d_pad] of
 * b200grbm_spin_extract_* into packed e2m1 rows [m][row_bytes] (two spins per byte, row_bytes a multiple of 128 with
 * 2 * row_bytes >= d_pad, padding zero); b200grbm_mmd_hist_fp4 has the contract of b200grbm_mmd_hist_i8 on them.
 */
int32_t b200grbm_pack_fp4_i8(const int8_t *rows_dev, int32_t m, int32_t d_pad, uint8_t *out_dev, int32_t row_bytes, void *stream);
int32_t b200grbm_mmd_hist_fp4(const uint8_t *z4_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t row_bytes,
                              int32_t shard_rank, int32_t shard_world, uint64_t *hist_dev, void *stream);

/*
 * Cross-rank row exchange of the sharded MMD (the reference has no collective -- its
 * maximum_mean_discrepancy_loss call at src/model_wrapper.py:320 sees the whole batch in one process).  Rows are +-1,
 * so ranks exchange ONE BIT per spin and the expansion to the int8 Gram operand is fused with the transfer.
 *
 * b200grbm_spin_pack_bits_*: rows [rows][d] (f32 by sign, or int8) -> bits_dev[(row_off + r) * words_per_row + w],
 *   bit k of word w = (spin 32 w + k is +1), columns >= d zero; words_per_row >= ceil(d / 32).
 * b200grbm_peer_signal: stream-ordered st.release.sys of `value` into a flag word ("my bit rows of step `value` are
 *   complete"), to be polled by the other ranks' b200grbm_bits_to_rows.
 * b200grbm_bits_to_rows: bits_ptrs[r] (HOST array of `world` DEVICE pointers: rank r's bit rows, [mx_loc + my_loc]
 *   [words_per_row], local or peer-mapped) -> z_dev int8 [world * (mx_loc + my_loc)][32 * words_per_row], the stacked
 *   matrix [x_0 .. x_{W-1}; y_0 .. y_{W-1}] of b200grbm_mmd_hist_i8 (padding columns zero).  flag_ptrs (HOST array or
 *   NULL): where non-NULL the kernel waits (ld.acquire.sys) until *flag_ptrs[r] >= step before reading rank r; a peer
 *   that never signals traps the kernel after 20 s instead of hanging the device.  world <= 16.
 * b200grbm_peer_alloc / _free: cudaMalloc'd, zeroed buffer plus its 64-byte cudaIpcMemHandle_t (handle_out: HOST, 64
 *   bytes) for the other ranks; b200grbm_peer_open / _close: map / unmap a peer's buffer in this process
 *   (cudaIpcOpenMemHandle with lazy peer access).  These four synchronise the device; they are setup calls.
 */
int32_t b200grbm_spin_pack_bits_f32(const float *x_dev, int32_t rows, int32_t d, uint32_t *bits_dev, int32_t words_per_row,
                                    int32_t row_off, void *stream);
int32_t b200grbm_spin_pack_bits_i8(const int8_t *x_dev, int32_t rows, int32_t d, uint32_t *bits_dev, int32_t words_per_row,
                                   int32_t row_off, void *stream);
int32_t b200grbm_peer_signal(uint32_t *flag_dev, uint32_t value, void *stream);
int32_t b200grbm_bits_to_rows(const void *const *bits_ptrs, const void *const *flag_ptrs, int32_t world, int32_t mx_loc,
                              int32_t my_loc, int32_t d, int32_t words_per_row, int8_t *z_dev, uint32_t step, void *stream);
int32_t b200grbm_peer_alloc(int64_t bytes, void **ptr_out, void *handle_out);
int32_t b200grbm_peer_open(const void *handle, void **ptr_out);
int32_t b200grbm_peer_close(void *ptr);
int32_t b200grbm_peer_free(void *ptr);

#ifdef __cplusplus
}
#endif
#endif /* B200GRBM_H */
