// GRBM parameter preparation, sign packing, integer sufficient statistics and energies.
//
// Reference sites:
//   set_weights     : parameter scaling / clipping inside GraphRestrictedBoltzmannMachine.sample
//                     (call sites src/model_wrapper.py:309-316, src/utils/persistent_qpu_sampler.py:71-78)
//   pack_*          : the float<->int8 conversions around sampleset_to_tensor (src/losses.py:59)
//   edge_stats      : gradient of src/losses.py:61 wrt linear / quadratic (sufficient statistics)
//   energy_*        : GraphRestrictedBoltzmannMachine.forward, call site src/losses.py:61
//
// All of these are HBM-bound streaming kernels over the sample matrix: rows are read once,
// coalesced along the spin axis; statistics are integer (popcount of XORed bit-packed words)
// so that sums are exact and independent of reduction order / GPU count.
#include "common.cuh"

namespace b200grbm {

// ------------------------------------------------------------------ set_weights

__global__ void set_edge_weights_kernel(const float *__restrict__ quadratic, int n_edges, float prefactor, float lo,
                                        float hi, const int32_t *__restrict__ slot_a,
                                        const int32_t *__restrict__ slot_b, uint2 *__restrict__ tiles,
                                        float *__restrict__ j_eff)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const float j = fminf(fmaxf(__fmul_rn(prefactor, quadratic[e]), lo), hi);
    if (j_eff != nullptr) j_eff[e] = j;
    const uint32_t j2 = f2u(__fmul_rn(2.0f, j));
    tiles[slot_a[e]].x = j2;
    tiles[slot_b[e]].x = j2;
}

__global__ void set_node_weights_kernel(const float *__restrict__ linear, int n, float prefactor, float lo, float hi,
                                        const int32_t *__restrict__ order, const int32_t *__restrict__ row_base,
                                        int width, int stride, uint2 *__restrict__ tiles, float *__restrict__ h_eff)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int node = order[p];
    const float h = fminf(fmaxf(__fmul_rn(prefactor, linear[node]), lo), hi);
    if (h_eff != nullptr) h_eff[node] = h;
    float a = h;
    uint2 *row = tiles + row_base[p];   // row 0 = f0, rows 1 .. width = slots
    // contract order: subtract J_k for k ascending; padded slots hold 2J = 0
    for (int k = 0; k < width; ++k) a = __fsub_rn(a, __fmul_rn(0.5f, u2f(row[(size_t)(k + 1) * stride].x)));
    row[0].x = f2u(a);
}

// ------------------------------------------------------------------ sign packing

template <typename T>
__global__ void pack_kernel(const T *__restrict__ x, int rows, int n, int n_pad, const int32_t *__restrict__ pos,
                            int cpl, uint32_t *__restrict__ packed)
{
    const int g = blockIdx.y;
    const int row0 = g * cpl;
    const int nvalid = min(cpl, rows - row0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t w = 0;
        for (int c = 0; c < nvalid; ++c) w |= (x[(size_t)(row0 + c) * n + i] > (T)0 ? 1u : 0u) << c;
        packed[(size_t)g * n_pad + pos[i]] = w;
    }
}

// ------------------------------------------------------------------ integer statistics

__global__ void edge_stats_kernel(const uint32_t *__restrict__ packed, int rows, int cpl, int n_pad, int n_edges,
                                  const int32_t *__restrict__ edge_pi, const int32_t *__restrict__ edge_pj,
                                  int groups_per_block, unsigned long long *__restrict__ sum_ss)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int groups = (rows + cpl - 1) / cpl;
    const int g0 = blockIdx.y * groups_per_block;
    const int g1 = min(groups, g0 + groups_per_block);
    const int pi = edge_pi[e], pj = edge_pj[e];
    long long acc = 0;
    for (int g = g0; g < g1; ++g) {
        const int nvalid = min(cpl, rows - g * cpl);
        const uint32_t mask = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
        const uint32_t *row = packed + (size_t)g * n_pad;
        acc += nvalid - 2 * __popc((row[pi] ^ row[pj]) & mask);
    }
    atomicAdd(sum_ss + e, (unsigned long long)acc);
}

__global__ void node_stats_kernel(const uint32_t *__restrict__ packed, int rows, int cpl, int n, int n_pad,
                                  const int32_t *__restrict__ order, int groups_per_block,
                                  unsigned long long *__restrict__ sum_s)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int groups = (rows + cpl - 1) / cpl;
    const int g0 = blockIdx.y * groups_per_block;
    const int g1 = min(groups, g0 + groups_per_block);
    long long acc = 0;
    for (int g = g0; g < g1; ++g) {
        const int nvalid = min(cpl, rows - g * cpl);
        const uint32_t mask = nvalid >= 32 ? 0xffffffffu : ((1u << nvalid) - 1u);
        acc += 2 * __popc(packed[(size_t)g * n_pad + p] & mask) - nvalid;
    }
    atomicAdd(sum_s + order[p], (unsigned long long)acc);
}

// ------------------------------------------------------------------ energies

template <typename T>
__device__ __forceinline__ T block_sum(T v, T *scratch)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    T t = (threadIdx.x < nw) ? scratch[threadIdx.x] : (T)0;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;  // valid in thread 0
}

// one CTA per row; the row is staged in shared memory so edge gathers never leave the SM
__global__ void energy_forward_kernel(const float *__restrict__ x, int n, int n_edges,
                                      const int32_t *__restrict__ ei, const int32_t *__restrict__ ej,
                                      const float *__restrict__ linear, const float *__restrict__ quadratic,
                                      float *__restrict__ energy)
{
    extern __shared__ float row[];
    __shared__ float scratch[32];
    const float *xr = x + (size_t)blockIdx.x * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) row[i] = xr[i];
    __syncthreads();
    float acc = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fmaf(linear[i], row[i], acc);
    for (int e = threadIdx.x; e < n_edges; e += blockDim.x) acc = fmaf(quadratic[e], row[ei[e]] * row[ej[e]], acc);
    const float tot = block_sum<float>(acc, scratch);
    if (threadIdx.x == 0) energy[blockIdx.x] = tot;
}

__global__ void energy_i8_kernel(const int8_t *__restrict__ s, int n, int n_edges, const int32_t *__restrict__ ei,
                                 const int32_t *__restrict__ ej, const float *__restrict__ h,
                                 const float *__restrict__ j, double *__restrict__ energy)
{
    extern __shared__ int8_t srow[];
    __shared__ double dscratch[32];
    const int8_t *sr = s + (size_t)blockIdx.x * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) srow[i] = sr[i];
    __syncthreads();
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += (double)h[i] * (double)srow[i];
    for (int e = threadIdx.x; e < n_edges; e += blockDim.x)
        acc += (double)j[e] * (double)(srow[ei[e]] * srow[ej[e]]);
    const double tot = block_sum<double>(acc, dscratch);
    if (threadIdx.x == 0) energy[blockIdx.x] = tot;
}


// Energies of the (<= 32) chains of one bit-packed group per CTA: E_c = sum_i h_i s_i + sum_e J_e s_i s_j
// with s = +1 where the bit is set.  Per edge ONE pair of state words serves all chains of the group
// (x = W[pi] ^ W[pj]: bit set <=> the two spins differ), so shared-memory traffic and the edge-list reads
// are amortised 28-fold against the one-row-per-CTA int8 kernel above; accumulation in double:
//   E_c = (sum_e J_e - 2 sum_{e: x_c = 1} J_e) + (2 sum_{i: bit_c = 1} h_i - sum_i h_i).
__global__ void __launch_bounds__(512) energy_packed_kernel(const uint32_t *__restrict__ packed, int rows, int cpl, int n,
                                                           int n_pad, int n_edges, const int32_t *__restrict__ edge_pi,
                                                           const int32_t *__restrict__ edge_pj,
                                                           const int32_t *__restrict__ order,
                                                           const float *__restrict__ h, const float *__restrict__ j,
                                                           double *__restrict__ energy)
{
    extern __shared__ uint32_t wrow[];
    __shared__ double red[16][33];
    const int g = blockIdx.x;
    const uint32_t *row = packed + (size_t)g * n_pad;
    for (int p = threadIdx.x; p < n; p += blockDim.x) wrow[p] = row[p];
    __syncthreads();
    double acc[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[c] = 0.0;
    double base = 0.0;
    for (int e = threadIdx.x; e < n_edges; e += blockDim.x) {
        const double jd = (double)j[e];
        const uint32_t x = wrow[edge_pi[e]] ^ wrow[edge_pj[e]];
        base += jd;
        const double m2 = -2.0 * jd;
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (x & (1u << c)) acc[c] += m2;
    }
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        const double hd = (double)h[order[p]];
        const uint32_t x = wrow[p];
        base -= hd;
        const double p2 = 2.0 * hd;
#pragma unroll
        for (int c = 0; c < 32; ++c)
            if (x & (1u << c)) acc[c] += p2;
    }
    // block reduction: lanes first, then the warps through shared memory (column 32 = the common term)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        double v = acc[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][c] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(0xffffffffu, base, o);
    if (lane == 0) red[warp][32] = base;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    if (threadIdx.x < 32) {
        double v = 0.0, b = 0.0;
        for (int w = 0; w < nw; ++w) {
            v += red[w][threadIdx.x];
            b += red[w][32];
        }
        const int r = g * cpl + threadIdx.x;
        if (threadIdx.x < cpl && r < rows) energy[r] = v + b;
    }
}

// thread per parameter, rows split over blockIdx.y; fp32 partial sums, one atomic per thread
__global__ void energy_backward_kernel(const float *__restrict__ x, const float *__restrict__ g, int rows, int n,
                                       int n_edges, const int32_t *__restrict__ ei, const int32_t *__restrict__ ej,
                                       int rows_per_block, float *__restrict__ grad_linear,
                                       float *__restrict__ grad_quadratic)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n + n_edges) return;
    const int r0 = blockIdx.y * rows_per_block;
    const int r1 = min(rows, r0 + rows_per_block);
    float acc = 0.f;
    if (idx < n) {
        for (int r = r0; r < r1; ++r) acc = fmaf(g[r], x[(size_t)r * n + idx], acc);
        atomicAdd(grad_linear + idx, acc);
    } else {
        const int e = idx - n;
        const int a = ei[e], b = ej[e];
        for (int r = r0; r < r1; ++r) {
            const float *xr = x + (size_t)r * n;
            acc = fmaf(g[r], xr[a] * xr[b], acc);
        }
        atomicAdd(grad_quadratic + e, acc);
    }
}

// dE_r/dx_ri = linear_i + sum_{e ni i} quadratic_e x_r,other(e), times the incoming gradient g_r.  One CTA per row:
// the row and its gradient accumulators live in shared memory, the edge list is streamed once.
__global__ void energy_grad_x_kernel(const float *__restrict__ x, const float *__restrict__ g, int n, int n_edges,
                                     const int32_t *__restrict__ ei, const int32_t *__restrict__ ej,
                                     const float *__restrict__ linear, const float *__restrict__ quadratic,
                                     float *__restrict__ grad_x)
{
    extern __shared__ float rowbuf[];
    float *row = rowbuf, *acc = rowbuf + n;
    const float *xr = x + (size_t)blockIdx.x * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { row[i] = xr[i]; acc[i] = linear[i]; }
    __syncthreads();
    for (int e = threadIdx.x; e < n_edges; e += blockDim.x) {
        const int a = ei[e], b = ej[e];
        const float q = quadratic[e];
        atomicAdd(acc + a, q * row[b]);
        atomicAdd(acc + b, q * row[a]);
    }
    __syncthreads();
    const float gr = g[blockIdx.x];
    float *out = grad_x + (size_t)blockIdx.x * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = gr * acc[i];
}

}  // namespace b200grbm

using namespace b200grbm;

extern "C" int32_t b200grbm_energy_grad_x(const float *x_dev, const float *grad_energy_dev, int32_t rows, int32_t n,
                                          int32_t n_edges, const int32_t *edge_i_dev, const int32_t *edge_j_dev,
                                          const float *linear_dev, const float *quadratic_dev, float *grad_x_dev,
                                          void *stream)
{
    if (rows <= 0 || n <= 0 || n_edges < 0) return fail(B200GRBM_EINVAL, "energy_grad_x: rows=%d n=%d n_edges=%d", rows, n, n_edges);
    if (!x_dev || !grad_energy_dev || !linear_dev || !grad_x_dev || (n_edges > 0 && (!edge_i_dev || !edge_j_dev || !quadratic_dev)))
        return fail(B200GRBM_EINVAL, "energy_grad_x: NULL pointer argument");
    B200_TRY(require_device());
    const size_t smem = 2 * sizeof(float) * (size_t)n;
    if (smem > 200 * 1024) return fail(B200GRBM_EUNSUPPORTED, "energy_grad_x: n=%d too large for a shared-memory row", n);
    B200_CUDA(cudaFuncSetAttribute(energy_grad_x_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    energy_grad_x_kernel<<<rows, 256, smem, (cudaStream_t)stream>>>(x_dev, grad_energy_dev, n, n_edges, edge_i_dev, edge_j_dev,
                                                                    linear_dev, quadratic_dev, grad_x_dev);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_set_weights(const float *linear_dev, const float *quadratic_dev, int32_t n, int32_t n_edges,
                                        float prefactor, float h_lo, float h_hi, float j_lo, float j_hi,
                                        const int32_t *order_dev, const int32_t *slot_a_dev, const int32_t *slot_b_dev,
                                        const int32_t *row_base_dev, int32_t ell_width, int32_t threads,
                                        b200grbm_ell_entry *tiles_dev, float *h_eff_dev, float *j_eff_dev, void *stream)
{
    if (n <= 0 || n_edges < 0 || ell_width <= 0 || threads <= 0)
        return fail(B200GRBM_EINVAL, "set_weights: n=%d n_edges=%d ell_width=%d threads=%d", n, n_edges, ell_width, threads);
    if (!linear_dev || !order_dev || !row_base_dev || !tiles_dev ||
        (n_edges > 0 && (!quadratic_dev || !slot_a_dev || !slot_b_dev)))
        return fail(B200GRBM_EINVAL, "set_weights: NULL pointer argument");
    if (!(h_lo <= h_hi) || !(j_lo <= j_hi)) return fail(B200GRBM_EINVAL, "set_weights: empty clipping range");
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    uint2 *tiles = reinterpret_cast<uint2 *>(tiles_dev);
    if (n_edges > 0) {
        set_edge_weights_kernel<<<(n_edges + 255) / 256, 256, 0, st>>>(quadratic_dev, n_edges, prefactor, j_lo, j_hi,
                                                                      slot_a_dev, slot_b_dev, tiles, j_eff_dev);
        B200_CUDA(cudaGetLastError());
    }
    set_node_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>(linear_dev, n, prefactor, h_lo, h_hi, order_dev,
                                                            row_base_dev, ell_width, threads, tiles, h_eff_dev);
    B200_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int32_t pack_impl(const T *x_dev, int32_t rows, int32_t n, int32_t n_pad, const int32_t *pos_dev, int32_t cpl,
                         uint32_t *packed_dev, void *stream)
{
    if (rows <= 0 || n <= 0 || n_pad < n || cpl < 1 || cpl > 32)
        return fail(B200GRBM_EINVAL, "pack: rows=%d n=%d n_pad=%d chains_per_lane=%d", rows, n, n_pad, cpl);
    if (!x_dev || !pos_dev || !packed_dev) return fail(B200GRBM_EINVAL, "pack: NULL pointer argument");
    B200_TRY(require_device());
    const int groups = (rows + cpl - 1) / cpl;
    if (groups > 65535) return fail(B200GRBM_EUNSUPPORTED, "pack: %d groups exceed grid.y; split the call", groups);
    dim3 grid((n + 255) / 256, groups);
    pack_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(x_dev, rows, n, n_pad, pos_dev, cpl, packed_dev);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_pack_f32(const float *x_dev, int32_t rows, int32_t n, int32_t n_pad, const int32_t *pos_dev,
                                     int32_t cpl, uint32_t *packed_dev, void *stream)
{
    return pack_impl<float>(x_dev, rows, n, n_pad, pos_dev, cpl, packed_dev, stream);
}

extern "C" int32_t b200grbm_pack_i8(const int8_t *x_dev, int32_t rows, int32_t n, int32_t n_pad, const int32_t *pos_dev,
                                    int32_t cpl, uint32_t *packed_dev, void *stream)
{
    return pack_impl<int8_t>(x_dev, rows, n, n_pad, pos_dev, cpl, packed_dev, stream);
}

extern "C" int32_t b200grbm_edge_stats(const uint32_t *packed_dev, int32_t rows, int32_t cpl, int32_t n, int32_t n_pad,
                                       int32_t n_edges, const int32_t *edge_pi_dev, const int32_t *edge_pj_dev,
                                       const int32_t *order_dev, int64_t *sum_s_dev, int64_t *sum_ss_dev, void *stream)
{
    if (rows <= 0 || n <= 0 || n_pad < n || cpl < 1 || cpl > 32 || n_edges < 0)
        return fail(B200GRBM_EINVAL, "edge_stats: rows=%d n=%d n_pad=%d cpl=%d n_edges=%d", rows, n, n_pad, cpl, n_edges);
    if (!packed_dev || !order_dev || !sum_s_dev || (n_edges > 0 && (!edge_pi_dev || !edge_pj_dev || !sum_ss_dev)))
        return fail(B200GRBM_EINVAL, "edge_stats: NULL pointer argument");
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const int groups = (rows + cpl - 1) / cpl;
    // enough CTAs to fill the machine: split the group axis when the edge axis alone is short
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int eblocks = (n_edges + 255) / 256, nblocks = (n + 255) / 256;
    int ysplit = (4 * sms + (eblocks > 0 ? eblocks : 1) - 1) / (eblocks > 0 ? eblocks : 1);
    if (ysplit > groups) ysplit = groups;
    if (ysplit < 1) ysplit = 1;
    if (ysplit > 65535) ysplit = 65535;
    const int gpb = (groups + ysplit - 1) / ysplit;
    const int ygrid = (groups + gpb - 1) / gpb;
    if (n_edges > 0) {
        edge_stats_kernel<<<dim3(eblocks, ygrid), 256, 0, st>>>(packed_dev, rows, cpl, n_pad, n_edges, edge_pi_dev,
                                                                edge_pj_dev, gpb,
                                                                reinterpret_cast<unsigned long long *>(sum_ss_dev));
        B200_CUDA(cudaGetLastError());
    }
    // the node axis is ~8x shorter than the edge axis: its own split of the group axis
    int ysplit_n = (4 * sms + nblocks - 1) / nblocks;
    if (ysplit_n > groups) ysplit_n = groups;
    if (ysplit_n > 65535) ysplit_n = 65535;
    const int gpb_n = (groups + ysplit_n - 1) / ysplit_n;
    node_stats_kernel<<<dim3(nblocks, (groups + gpb_n - 1) / gpb_n), 256, 0, st>>>(
        packed_dev, rows, cpl, n, n_pad, order_dev, gpb_n, reinterpret_cast<unsigned long long *>(sum_s_dev));
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_energy_forward(const float *x_dev, int32_t rows, int32_t n, int32_t n_edges,
                                           const int32_t *edge_i_dev, const int32_t *edge_j_dev, const float *linear_dev,
                                           const float *quadratic_dev, float *energy_dev, void *stream)
{
    if (rows <= 0 || n <= 0 || n_edges < 0) return fail(B200GRBM_EINVAL, "energy_forward: rows=%d n=%d n_edges=%d", rows, n, n_edges);
    if (!x_dev || !linear_dev || !energy_dev || (n_edges > 0 && (!edge_i_dev || !edge_j_dev || !quadratic_dev)))
        return fail(B200GRBM_EINVAL, "energy_forward: NULL pointer argument");
    B200_TRY(require_device());
    const size_t smem = sizeof(float) * (size_t)n;
    if (smem > 200 * 1024) return fail(B200GRBM_EUNSUPPORTED, "energy_forward: n=%d too large for a shared-memory row", n);
    B200_CUDA(cudaFuncSetAttribute(energy_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    energy_forward_kernel<<<rows, 256, smem, (cudaStream_t)stream>>>(x_dev, n, n_edges, edge_i_dev, edge_j_dev, linear_dev,
                                                                     quadratic_dev, energy_dev);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_energy_i8(const int8_t *s_dev, int32_t rows, int32_t n, int32_t n_edges,
                                      const int32_t *edge_i_dev, const int32_t *edge_j_dev, const float *h_dev,
                                      const float *j_dev, double *energy_dev, void *stream)
{
    if (rows <= 0 || n <= 0 || n_edges < 0) return fail(B200GRBM_EINVAL, "energy_i8: rows=%d n=%d n_edges=%d", rows, n, n_edges);
    if (!s_dev || !h_dev || !energy_dev || (n_edges > 0 && (!edge_i_dev || !edge_j_dev || !j_dev)))
        return fail(B200GRBM_EINVAL, "energy_i8: NULL pointer argument");
    B200_TRY(require_device());
    const size_t smem = (size_t)n;
    if (smem > 200 * 1024) return fail(B200GRBM_EUNSUPPORTED, "energy_i8: n=%d too large for a shared-memory row", n);
    B200_CUDA(cudaFuncSetAttribute(energy_i8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    energy_i8_kernel<<<rows, 256, smem, (cudaStream_t)stream>>>(s_dev, n, n_edges, edge_i_dev, edge_j_dev, h_dev, j_dev,
                                                                energy_dev);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_energy_packed(const uint32_t *packed_dev, int32_t rows, int32_t chains_per_lane, int32_t n,
                                          int32_t n_pad, int32_t n_edges, const int32_t *edge_pi_dev,
                                          const int32_t *edge_pj_dev, const int32_t *order_dev, const float *h_dev,
                                          const float *j_dev, double *energy_dev, void *stream)
{
    if (rows <= 0 || n <= 0 || n_edges < 0 || n_pad < n || chains_per_lane <= 0 || chains_per_lane > 32)
        return fail(B200GRBM_EINVAL, "energy_packed: rows=%d n=%d n_pad=%d n_edges=%d chains_per_lane=%d", rows, n, n_pad,
                    n_edges, chains_per_lane);
    if (!packed_dev || !order_dev || !h_dev || !energy_dev || (n_edges > 0 && (!edge_pi_dev || !edge_pj_dev || !j_dev)))
        return fail(B200GRBM_EINVAL, "energy_packed: NULL pointer argument");
    B200_TRY(require_device());
    const size_t smem = (size_t)n * 4;
    if (smem > 160 * 1024) return fail(B200GRBM_EUNSUPPORTED, "energy_packed: n=%d too large for a shared-memory row", n);
    B200_CUDA(cudaFuncSetAttribute(energy_packed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int groups = (rows + chains_per_lane - 1) / chains_per_lane;
    energy_packed_kernel<<<groups, 512, smem, (cudaStream_t)stream>>>(packed_dev, rows, chains_per_lane, n, n_pad, n_edges,
                                                                      edge_pi_dev, edge_pj_dev, order_dev, h_dev, j_dev,
                                                                      energy_dev);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_energy_backward(const float *x_dev, const float *grad_energy_dev, int32_t rows, int32_t n,
                                            int32_t n_edges, const int32_t *edge_i_dev, const int32_t *edge_j_dev,
                                            float *grad_linear_dev, float *grad_quadratic_dev, void *stream)
{
    if (rows <= 0 || n <= 0 || n_edges < 0) return fail(B200GRBM_EINVAL, "energy_backward: rows=%d n=%d n_edges=%d", rows, n, n_edges);
    if (!x_dev || !grad_energy_dev || !grad_linear_dev || (n_edges > 0 && (!edge_i_dev || !edge_j_dev || !grad_quadratic_dev)))
        return fail(B200GRBM_EINVAL, "energy_backward: NULL pointer argument");
    B200_TRY(require_device());
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int xblocks = (n + n_edges + 255) / 256;
    int ysplit = (4 * sms + xblocks - 1) / xblocks;
    if (ysplit > rows) ysplit = rows;
    if (ysplit < 1) ysplit = 1;
    if (ysplit > 65535) ysplit = 65535;
    const int rpb = (rows + ysplit - 1) / ysplit;
    const int ygrid = (rows + rpb - 1) / rpb;
    energy_backward_kernel<<<dim3(xblocks, ygrid), 256, 0, (cudaStream_t)stream>>>(
        x_dev, grad_energy_dev, rows, n, n_edges, edge_i_dev, edge_j_dev, rpb, grad_linear_dev, grad_quadratic_dev);
    B200_CUDA(cudaGetLastError());
    return 0;
}
