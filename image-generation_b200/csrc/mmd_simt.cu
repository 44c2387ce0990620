// Mixture-of-RBF MMD, fp32 CUDA-core path (any real-valued inputs) -- forward, backward.
//
// Replaces  maximum_mean_discrepancy_loss(x, y, GaussianKernel(n_kernels=7))  (third-party
// dwave-pytorch-plugin; call site src/model_wrapper.py:320, kernel built :273) whose stock
// form is  cat -> cdist -> 7 x exp over a (7, m, m) temporary -> block means  (SURVEY.md
// Appendix A.3).  Here the pairwise distances are produced tile by tile in shared memory /
// registers and consumed in the epilogue, so nothing of size m^2 is ever written (the
// backward pass writes one m_x x m coefficient matrix, consumed by a tiled product).
//
// Distances are accumulated as sum_k (a_k - b_k)^2 (no ||a||^2 + ||b||^2 - 2ab cancellation),
// so this path is the precise one; the +-1 fast path on tcgen05 tensor cores is mmd_tc.cu.
#include "common.cuh"

namespace b200grbm {

constexpr int TILE = 64;   // pairs per CTA edge
constexpr int KC = 16;     // feature chunk
constexpr int MAX_KERNELS = 16;

struct MmdParams {
    const float *z;        // [m][d]
    int m_x, m_y, d;
    int n_kernels;
    int squared;
    float mul_factor;
    float bandwidth;       // <= 0: automatic (sum of distances / (m^2 - m))
    double *sums;          // [4] S_xx, S_yy, S_xy, dist_sum
    // backward only
    float w_xx, w_xy;
    const float *grad_out; // scalar upstream gradient (device)
    float *coef;           // [m_x][m]
    float *grad_x;         // [m_x][d]
};

enum { PASS_DIST = 0, PASS_KERNEL = 1, PASS_COEF = 2 };

__device__ __forceinline__ void load_scales(const MmdParams &p, float (&c)[MAX_KERNELS])
{
    const double m = (double)(p.m_x + p.m_y);
    const double bw = p.bandwidth > 0.f ? (double)p.bandwidth : p.sums[3] / (m * m - m);
    for (int u = 0; u < p.n_kernels; ++u) {
        const double mult = pow((double)p.mul_factor, (double)(u - p.n_kernels / 2));
        c[u] = (float)(-1.4426950408889634 / (bw * mult));   // exp(-t/(bw mult)) = exp2(t * c)
    }
}

template <int PASS>
__global__ void __launch_bounds__(256) mmd_pair_kernel(const MmdParams p)
{
    __shared__ float As[KC][TILE + 4];
    __shared__ float Bs[KC][TILE + 4];
    __shared__ double red[3][8];
    const int m = p.m_x + p.m_y;
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (PASS != PASS_COEF && tj < ti) return;          // symmetric: upper triangle of tiles only
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads, 4 x 4 pairs each
    const int row0 = ti * TILE, col0 = tj * TILE;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.d; k0 += KC) {
        // 64 rows x 16 features per operand, 4 elements per thread, transposed into smem
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + q * 256;
            const int r = idx >> 4, k = idx & 15;
            const int ga = row0 + r, gb = col0 + r, gk = k0 + k;
            As[k][r] = (ga < m && gk < p.d) ? p.z[(size_t)ga * p.d + gk] : 0.f;
            Bs[k][r] = (gb < m && gk < p.d) ? p.z[(size_t)gb * p.d + gk] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float df = a[i] - b[j];
                    acc[i][j] = fmaf(df, df, acc[i][j]);
                }
        }
        __syncthreads();
    }

    float c[MAX_KERNELS];
    if (PASS != PASS_DIST) load_scales(p, c);
    const float g_up = PASS == PASS_COEF ? p.grad_out[0] : 0.f;
    double s_xx = 0.0, s_yy = 0.0, s_xy = 0.0;   // PASS_DIST uses s_xx as the distance sum
    const float wsym = (tj > ti) ? 2.f : 1.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = row0 + ty * 4 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gj = col0 + tx * 4 + j;
            if (gi >= m || gj >= m) continue;
            const float d2 = acc[i][j];
            const float t = p.squared ? d2 : sqrtf(d2);
            if (PASS == PASS_DIST) {
                s_xx += (double)(wsym * t);
            } else if (PASS == PASS_KERNEL) {
                float kv = 0.f;
                for (int u = 0; u < p.n_kernels; ++u) kv += exp2f(t * c[u]);
                const bool ix = gi < p.m_x, jx = gj < p.m_x;
                if (ix && jx) s_xx += (double)(wsym * kv);
                else if (!ix && !jx) s_yy += (double)(wsym * kv);
                else if (ix && !jx) s_xy += (double)kv;
            } else {
                if (gi >= p.m_x) continue;
                float dk = 0.f;                     // d k / d t = sum_u c_u ln2 exp2(t c_u)
                for (int u = 0; u < p.n_kernels; ++u) dk = fmaf(c[u], exp2f(t * c[u]), dk);
                dk *= 0.6931471805599453f;
                float cf;
                if (p.squared) cf = 2.f * dk;
                else cf = t > 0.f ? dk / t : 0.f;
                if (gi == gj) cf = 0.f;
                const float w = gj < p.m_x ? p.w_xx : p.w_xy;
                p.coef[(size_t)gi * m + gj] = g_up * w * cf;
            }
        }
    }
    if (PASS == PASS_COEF) return;
    // block reduction of up to three double partial sums
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_xx += __shfl_xor_sync(0xffffffffu, s_xx, o);
        s_yy += __shfl_xor_sync(0xffffffffu, s_yy, o);
        s_xy += __shfl_xor_sync(0xffffffffu, s_xy, o);
    }
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) { red[0][warp] = s_xx; red[1][warp] = s_yy; red[2][warp] = s_xy; }
    __syncthreads();
    if (tid < 3) {
        double tot = 0.0;
        for (int w = 0; w < 8; ++w) tot += red[tid][w];
        if (PASS == PASS_DIST) { if (tid == 0) atomicAdd(p.sums + 3, tot); }
        else if (tot != 0.0) atomicAdd(p.sums + tid, tot);
    }
}

// grad_x[i][k] = rowsum_i(coef) * x[i][k] - sum_j coef[i][j] z[j][k]      (64 x 64 tiles)
__global__ void __launch_bounds__(256) mmd_grad_kernel(const MmdParams p)
{
    __shared__ float As[KC][TILE + 4];   // coef^T chunk: [j][i]
    __shared__ float Bs[KC][TILE + 4];   // z chunk:      [j][k]
    const int m = p.m_x + p.m_y;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int row0 = blockIdx.y * TILE, col0 = blockIdx.x * TILE;
    float acc[4][4], rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int j0 = 0; j0 < m; j0 += KC) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int idx = tid + q * 256;
            {   // coef tile: 64 rows(i) x 16 (j); consecutive threads along j
                const int r = idx >> 4, j = idx & 15;
                const int gi = row0 + r, gj = j0 + j;
                As[j][r] = (gi < p.m_x && gj < m) ? p.coef[(size_t)gi * m + gj] : 0.f;
            }
            {   // z tile: 16 rows(j) x 64 features; consecutive threads along k
                const int j = idx >> 6, k = idx & 63;
                const int gj = j0 + j, gk = col0 + k;
                Bs[j][k] = (gj < m && gk < p.d) ? p.z[(size_t)gj * p.d + gk] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < KC; ++j) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[j][ty * 4 + i]; rs[i] += a[i]; }
#pragma unroll
            for (int k = 0; k < 4; ++k) b[k] = Bs[j][tx * 4 + k];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(a[i], b[k], acc[i][k]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gi = row0 + ty * 4 + i;
        if (gi >= p.m_x) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int gk = col0 + tx * 4 + k;
            if (gk >= p.d) continue;
            const size_t o = (size_t)gi * p.d + gk;
            p.grad_x[o] = rs[i] * p.z[o] - acc[i][k];
        }
    }
}

static int32_t validate(const float *z, int m_x, int m_y, int d, int n_kernels, float mul_factor, const char *who)
{
    if (m_x <= 0 || m_y <= 0 || d <= 0) return fail(B200GRBM_EINVAL, "%s: m_x=%d m_y=%d d=%d", who, m_x, m_y, d);
    if (n_kernels < 1 || n_kernels > MAX_KERNELS || !(mul_factor > 0.f))
        return fail(B200GRBM_EINVAL, "%s: n_kernels=%d (1..%d) mul_factor=%g", who, n_kernels, MAX_KERNELS, mul_factor);
    if (z == nullptr) return fail(B200GRBM_EINVAL, "%s: NULL input", who);
    return 0;
}

}  // namespace b200grbm

using namespace b200grbm;

extern "C" int32_t b200grbm_mmd_forward_f32(const float *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t n_kernels,
                                            float mul_factor, int32_t squared, float bandwidth, double *sums_dev,
                                            void *stream)
{
    B200_TRY(validate(z_dev, m_x, m_y, d, n_kernels, mul_factor, "mmd_forward_f32"));
    if (sums_dev == nullptr) return fail(B200GRBM_EINVAL, "mmd_forward_f32: NULL sums");
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    MmdParams p = {};
    p.z = z_dev; p.m_x = m_x; p.m_y = m_y; p.d = d; p.n_kernels = n_kernels; p.squared = squared;
    p.mul_factor = mul_factor; p.bandwidth = bandwidth; p.sums = sums_dev;
    const int m = m_x + m_y, tiles = (m + TILE - 1) / TILE;
    B200_CUDA(cudaMemsetAsync(sums_dev, 0, 4 * sizeof(double), st));
    if (!(bandwidth > 0.f)) {
        mmd_pair_kernel<PASS_DIST><<<dim3(tiles, tiles), 256, 0, st>>>(p);
        B200_CUDA(cudaGetLastError());
    }
    mmd_pair_kernel<PASS_KERNEL><<<dim3(tiles, tiles), 256, 0, st>>>(p);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_mmd_backward_f32(const float *z_dev, int32_t m_x, int32_t m_y, int32_t d, int32_t n_kernels,
                                             float mul_factor, int32_t squared, float bandwidth, const double *sums_dev,
                                             float w_xx, float w_xy, const float *grad_out_dev, float *coef_dev,
                                             float *grad_x_dev, void *stream)
{
    B200_TRY(validate(z_dev, m_x, m_y, d, n_kernels, mul_factor, "mmd_backward_f32"));
    if (!sums_dev || !grad_out_dev || !coef_dev || !grad_x_dev)
        return fail(B200GRBM_EINVAL, "mmd_backward_f32: NULL pointer argument");
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    MmdParams p = {};
    p.z = z_dev; p.m_x = m_x; p.m_y = m_y; p.d = d; p.n_kernels = n_kernels; p.squared = squared;
    p.mul_factor = mul_factor; p.bandwidth = bandwidth; p.sums = const_cast<double *>(sums_dev);
    p.w_xx = w_xx; p.w_xy = w_xy; p.grad_out = grad_out_dev; p.coef = coef_dev; p.grad_x = grad_x_dev;
    const int m = m_x + m_y;
    mmd_pair_kernel<PASS_COEF><<<dim3((m + TILE - 1) / TILE, (m_x + TILE - 1) / TILE), 256, 0, st>>>(p);
    B200_CUDA(cudaGetLastError());
    mmd_grad_kernel<<<dim3((d + TILE - 1) / TILE, (m_x + TILE - 1) / TILE), 256, 0, st>>>(p);
    B200_CUDA(cudaGetLastError());
    return 0;
}
