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
 *     layer); the library never allocates or frees persistent device memory.
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

#define B200GRBM_ABI_VERSION 1

#define B200GRBM_EINVAL (-1)   /* bad argument / shape */
#define B200GRBM_EUNSUPPORTED (-2) /* configuration not compiled in (e.g. chains_per_lane) */
#define B200GRBM_ENODEVICE (-3) /* no sm_100 device */

#define B200GRBM_MAX_COLOURS 16

/* acceptance rule of the heat-bath draw (include/b200grbm_spec.h) */
#define B200GRBM_ACCEPT_EXACT 0 /* contract polynomial: bit-reproducible on the CPU oracle */
#define B200GRBM_ACCEPT_FAST 1  /* MUFU.EX2: same law, statistical parity only */

typedef struct b200grbm_ell_entry {
    uint32_t j2_bits; /* fp32 bits of 2*J_eff for this slot (0 for padding) */
    uint32_t nbr;     /* neighbour visit position */
} b200grbm_ell_entry;

const char *b200grbm_last_error(void);
int32_t b200grbm_abi_version(void);

/* sm count / compute capability / opt-in shared memory of the current device */
int32_t b200grbm_device_info(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor, int32_t *smem_optin);

/*
 * Effective Ising parameters and sampler tables from the GRBM parameters.
 * Replaces the parameter preparation inside GraphRestrictedBoltzmannMachine.sample
 * (third-party; call sites src/model_wrapper.py:309-316, :369-376,
 * src/utils/persistent_qpu_sampler.py:71-78): h = clip(prefactor * linear, linear_range),
 * J = clip(prefactor * quadratic, quadratic_range) -- done on device so the step never
 * round-trips the parameters through Python dicts.
 *   ell_dev    [ell_width][n_pad] entries, .nbr pre-filled by the host layer; .j2_bits written here
 *   f0_dev     [n]  h_eff - sum_k J_eff (contract order), visit-position order
 *   h_eff_dev  [n]  node order (optional, may be NULL);  j_eff_dev [n_edges] (optional)
 */
int32_t b200grbm_set_weights(const float *linear_dev, const float *quadratic_dev, int32_t n, int32_t n_edges,
                             float prefactor, float h_lo, float h_hi, float j_lo, float j_hi,
                             const int32_t *order_dev, const int32_t *slot_a_dev, const int32_t *slot_b_dev,
                             int32_t ell_width, int32_t n_pad, b200grbm_ell_entry *ell_dev, float *f0_dev,
                             float *h_eff_dev, float *j_eff_dev, void *stream);

typedef struct b200grbm_sweep_args {
    uint32_t struct_size;      /* sizeof(b200grbm_sweep_args) */
    int32_t n;                 /* spins */
    int32_t n_pad;             /* row pitch of the ELL tables and of packed state */
    int32_t ell_width;
    int32_t n_colours;
    int32_t colour_start[B200GRBM_MAX_COLOURS + 1]; /* visit positions of each colour block */
    const b200grbm_ell_entry *ell_dev;
    const float *f0_dev;
    const int32_t *order_dev;  /* [n] node visited at position p (for int8 node-order I/O) */
    int32_t chains;            /* chains in this call */
    int32_t chains_per_lane;   /* 16, 24, 28 or 32: chains bit-packed per state word */
    int32_t threads;           /* CTA size, multiple of 32 in [64, 768] */
    int32_t accept;            /* B200GRBM_ACCEPT_* */
    uint64_t chain_offset;     /* global id of chain 0 of this call (multiple of 4) */
    uint64_t seed;
    uint32_t sweep_offset;     /* Philox sweep counter of the first sweep */
    int32_t num_sweeps;
    const float *coef_dev;     /* [num_sweeps] (float)(2 beta log2 e) */
    const float *uniforms_dev; /* NULL -> Philox; else [num_sweeps][chains][n] visit-position order */
    const int8_t *state_in_dev;   /* [chains][n] node order, +-1; NULL -> packed_in or random init */
    const uint32_t *packed_in_dev; /* [groups][n_pad] visit-position order; NULL -> random init */
    int8_t *state_out_dev;     /* optional */
    uint32_t *packed_out_dev;  /* optional */
} b200grbm_sweep_args;

/*
 * sampler.sample_ising(h, J, num_reads=chains, ...) -- the call GraphRestrictedBoltzmannMachine
 * .sample makes (boundary: src/utils/common.py:123-138 builds the sampler and its kwargs).
 * Runs num_sweeps colour-blocked heat-bath sweeps on `chains` independent chains.
 */
int32_t b200grbm_gibbs_sweeps(const b200grbm_sweep_args *args, void *stream);

/* number of kernel launches the last b200grbm_gibbs_sweeps call on this thread enqueued */
int32_t b200grbm_last_launch_count(void);

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
/* int8 +-1 rows -> fp64 energies (dimod SampleSet.record.energy, src/utils/persistent_qpu_sampler.py:84-88) */
int32_t b200grbm_energy_i8(const int8_t *s_dev, int32_t rows, int32_t n, int32_t n_edges,
                           const int32_t *edge_i_dev, const int32_t *edge_j_dev, const float *h_dev,
                           const float *j_dev, double *energy_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200GRBM_H */
