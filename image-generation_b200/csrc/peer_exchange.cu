// Cross-rank row exchange of the sharded MMD ("MMD partial sums" -- the cross term of
// maximum_mean_discrepancy_loss, src/model_wrapper.py:320, needs every rank's encoder spins against every rank's samples).
//
// Rows are +-1, so what crosses NVLink is ONE BIT per spin (cfg3: 11.5 MB for the 16 384 x 5 640 matrix instead of the
// 92 MB of int8 rows), and the expansion to the int8 Gram operand is fused with the transfer:
//   spin_pack_bits_kernel   rows (fp32 by sign / int8) -> bit rows  [row][d_pad / 32] u32 in this rank's exchange buffer
//   peer_signal_kernel      publishes "step s is complete" in the buffer's flag word (st.release.sys)
//   bits_to_rows_kernel     ONE launch per rank: for every source rank r it waits for r's flag (ld.acquire.sys over NVLink),
//                           pulls r's bit rows straight out of r's memory (peer pointers from cudaIpcOpenMemHandle, plain
//                           loads through NVSwitch) and writes the int8 rows [x_0 .. x_{W-1}; y_0 .. y_{W-1}] the tcgen05
//                           Gram kernel reads -- no NCCL all-gather, no staging copy, no concatenation.
// The same kernel serves the fallback where the bit rows were all-gathered by NCCL into local memory (no flags).
//
// Re-use of a rank's bit buffer for the next step is ordered by the int64 histogram all-reduce that follows every
// exchange (dist._ShardedMMD.forward): a rank's all-reduce completes only after every rank has contributed, and a rank
// contributes only after its bits_to_rows launch -- its last read of peer memory -- has finished.
#include "common.cuh"

#include <string.h>

#include <algorithm>

namespace b200grbm {

constexpr int PEER_MAX_RANKS = 16;

template <typename T>
__global__ void __launch_bounds__(256) spin_pack_bits_kernel(const T *__restrict__ x, int rows, int d, uint32_t *__restrict__ bits,
                                                             int wpr, int row_off)
{
    // one warp task = 32 words (1024 spins) of one row: 32 coalesced 32-element loads, a ballot each; lane k keeps word k
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const int chunks = (wpr + 31) / 32;
    const long long tasks = (long long)rows * chunks;
    for (long long task = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); task < tasks; task += (long long)gridDim.x * wpb) {
        const int r = (int)(task / chunks), ch = (int)(task % chunks);
        const T *xr = x + (size_t)r * d;
        uint32_t mine = 0u;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
            const int col = (ch * 32 + k) * 32 + lane;
            const bool up = col < d && xr[col] > (T)0;
            const uint32_t b = __ballot_sync(0xffffffffu, up);
            if (lane == k) mine = b;
        }
        const int w = ch * 32 + lane;
        if (w < wpr) bits[(size_t)(row_off + r) * wpr + w] = mine;
    }
}

__global__ void peer_signal_kernel(uint32_t *flag, uint32_t value)
{
    // the bit rows were written by earlier kernels of this stream; release at system scope so that a peer's acquire
    // load of the flag (through NVLink, served by this GPU's L2) also sees them
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

struct PeerSources {
    const uint32_t *bits[PEER_MAX_RANKS];       // [mx_loc + my_loc][wpr] of rank r (peer-mapped or local)
    const uint32_t *flag[PEER_MAX_RANKS];       // NULL: no wait (local data / NCCL-gathered copy)
};

// four bits -> four bytes +1 / -1 (0x01 / 0xff)
__device__ __forceinline__ uint32_t nibble_to_spins(uint32_t n)
{
    const uint32_t spread = (n * 0x00204081u) & 0x01010101u;        // bit i -> byte i
    return ~(spread * 0xfeu);
}

__global__ void __launch_bounds__(256) bits_to_rows_kernel(const __grid_constant__ PeerSources src, int mx_loc, int my_loc, int m_x,
                                                           int d, int wpr, int8_t *__restrict__ z, uint32_t step)
{
    const int r = blockIdx.y;
    if (src.flag[r] != nullptr) {
        if (threadIdx.x == 0) {
            unsigned long long t0 = 0, now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            uint32_t seen;
            for (;;) {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(src.flag[r]) : "memory");
                if ((int32_t)(seen - step) >= 0) break;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (now - t0 > 20000000000ull) __trap();      // 20 s: a peer died -- fail loudly instead of hanging the GPU
                __nanosleep(200);
            }
        }
        __syncthreads();
    }
    const int rows_loc = mx_loc + my_loc;
    const size_t words = (size_t)rows_loc * wpr;
    const int d_pad = 32 * wpr;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < words; k += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(k / wpr), w = (int)(k % wpr);
        const uint32_t b = __ldcg(src.bits[r] + k);
        const int row = i < mx_loc ? r * mx_loc + i : m_x + r * my_loc + (i - mx_loc);
        uint32_t o[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            uint32_t v = nibble_to_spins((b >> (4 * q)) & 0xfu);
            const int c = 32 * w + 4 * q;                      // padding columns (>= d) are zero, not -1
            if (c + 4 > d) v = c >= d ? 0u : (v & (0xffffffffu >> (8 * (c + 4 - d))));
            o[q] = v;
        }
        uint4 *dst = reinterpret_cast<uint4 *>(z + (size_t)row * d_pad + 32 * w);
        dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
        dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
}

template <typename T>
static int32_t pack_bits_impl(const T *x_dev, int32_t rows, int32_t d, uint32_t *bits_dev, int32_t wpr, int32_t row_off, void *stream)
{
    if (rows <= 0 || d <= 0 || row_off < 0 || wpr <= 0 || (long long)wpr * 32 < d)
        return fail(B200GRBM_EINVAL, "spin_pack_bits: rows=%d d=%d row_off=%d words_per_row=%d", rows, d, row_off, wpr);
    if (!x_dev || !bits_dev) return fail(B200GRBM_EINVAL, "spin_pack_bits: NULL pointer");
    B200_TRY(require_device());
    const long long tasks = (long long)rows * ((wpr + 31) / 32);
    const int blocks = (int)std::min<long long>((tasks + 7) / 8, 148 * 16);
    spin_pack_bits_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(x_dev, rows, d, bits_dev, wpr, row_off);
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200grbm

using namespace b200grbm;

extern "C" int32_t b200grbm_spin_pack_bits_f32(const float *x_dev, int32_t rows, int32_t d, uint32_t *bits_dev, int32_t words_per_row,
                                               int32_t row_off, void *stream)
{
    return pack_bits_impl<float>(x_dev, rows, d, bits_dev, words_per_row, row_off, stream);
}

extern "C" int32_t b200grbm_spin_pack_bits_i8(const int8_t *x_dev, int32_t rows, int32_t d, uint32_t *bits_dev, int32_t words_per_row,
                                              int32_t row_off, void *stream)
{
    return pack_bits_impl<int8_t>(x_dev, rows, d, bits_dev, words_per_row, row_off, stream);
}

extern "C" int32_t b200grbm_peer_signal(uint32_t *flag_dev, uint32_t value, void *stream)
{
    if (!flag_dev) return fail(B200GRBM_EINVAL, "peer_signal: NULL flag");
    B200_TRY(require_device());
    peer_signal_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag_dev, value);
    B200_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int32_t b200grbm_bits_to_rows(const void *const *bits_ptrs, const void *const *flag_ptrs, int32_t world, int32_t mx_loc,
                                         int32_t my_loc, int32_t d, int32_t words_per_row, int8_t *z_dev, uint32_t step, void *stream)
{
    if (world < 1 || world > PEER_MAX_RANKS) return fail(B200GRBM_EUNSUPPORTED, "bits_to_rows: world=%d (1..%d)", world, PEER_MAX_RANKS);
    if (mx_loc < 0 || my_loc < 0 || mx_loc + my_loc <= 0 || d <= 0 || words_per_row <= 0 || (long long)words_per_row * 32 < d)
        return fail(B200GRBM_EINVAL, "bits_to_rows: mx_loc=%d my_loc=%d d=%d words_per_row=%d", mx_loc, my_loc, d, words_per_row);
    if (!bits_ptrs || !z_dev || (reinterpret_cast<uintptr_t>(z_dev) & 15u) != 0)
        return fail(B200GRBM_EINVAL, "bits_to_rows: NULL or unaligned pointer");
    PeerSources src;
    for (int r = 0; r < PEER_MAX_RANKS; ++r) {
        src.bits[r] = r < world ? static_cast<const uint32_t *>(bits_ptrs[r]) : nullptr;
        src.flag[r] = (r < world && flag_ptrs != nullptr) ? static_cast<const uint32_t *>(flag_ptrs[r]) : nullptr;
        if (r < world && src.bits[r] == nullptr) return fail(B200GRBM_EINVAL, "bits_to_rows: NULL source %d", r);
    }
    B200_TRY(require_device());
    const size_t words = (size_t)(mx_loc + my_loc) * words_per_row;
    // blocks that wait for a peer's flag must all be resident or be able to drain: cap the grid at what one wave holds
    // (8 CTAs of 256 threads per SM) so that no waiting block keeps a block with ready data off the machine
    const int per_src = std::max(1, (148 * 8) / world);
    dim3 grid((unsigned)std::min<size_t>((words + 255) / 256, (size_t)per_src), (unsigned)world);
    bits_to_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, mx_loc, my_loc, world * mx_loc, d, words_per_row, z_dev, step);
    B200_CUDA(cudaGetLastError());
    return 0;
}

// ---- exchange buffers: device memory this library allocates on explicit request so that it can be exported to the
// other ranks of the box (cudaIpc handles need the base of a cudaMalloc allocation, which a caching allocator's
// sub-blocks are not).  The caller owns the buffer until b200grbm_peer_free.
extern "C" int32_t b200grbm_peer_alloc(int64_t bytes, void **ptr_out, void *handle_out)
{
    if (bytes <= 0 || !ptr_out || !handle_out) return fail(B200GRBM_EINVAL, "peer_alloc: bytes=%lld", (long long)bytes);
    B200_TRY(require_device());
    void *p = nullptr;
    B200_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaFree(p);
        (void)cudaGetLastError();
        return check_cuda(e, "peer_alloc: cudaMemset / cudaIpcGetMemHandle");
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI documents a 64-byte handle");
    memcpy(handle_out, &h, sizeof(h));
    *ptr_out = p;
    return 0;
}

extern "C" int32_t b200grbm_peer_open(const void *handle, void **ptr_out)
{
    if (!handle || !ptr_out) return fail(B200GRBM_EINVAL, "peer_open: NULL argument");
    B200_TRY(require_device());
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();           // not sticky: the caller falls back to the NCCL exchange
        return check_cuda(e, "peer_open: cudaIpcOpenMemHandle");
    }
    *ptr_out = p;
    return 0;
}

extern "C" int32_t b200grbm_peer_close(void *ptr)
{
    if (!ptr) return 0;
    B200_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

extern "C" int32_t b200grbm_peer_free(void *ptr)
{
    if (!ptr) return 0;
    B200_CUDA(cudaFree(ptr));
    return 0;
}
