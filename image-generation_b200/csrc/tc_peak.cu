// Tensor-core peak probe: back-to-back tcgen05.mma from shared-memory operands that never change
// (no TMA, no epilogue), one CTA per SM.  Gives the denominator for the MMD kernels' tensor
// rooflines on THIS device: MEASURED_PEAKS.json has a bf16 cuBLAS figure but no int8 one
// ("measure an int8 tcgen05 peak on the box, or quote 2x bf16").
#include "tc_common.cuh"

namespace b200grbm {

template <int KIND>   // 0 = kind::i8 (K = 32 per MMA), 1 = kind::f16 with bf16 operands (K = 16 per MMA)
__global__ void __launch_bounds__(128, 1) tc_peak_kernel(int iters)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0u;
    if (threadIdx.x == 0) {
        bar_init(smem_addr(&bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(smem_addr(&tmem_slot), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy zero fill -> visible to the MMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        const uint64_t adesc = umma_desc_sw128(smem_addr(smem)), bdesc = umma_desc_sw128(smem_addr(smem) + 16384);
        const uint32_t idesc = KIND == 0 ? umma_idesc_i8(128, 256) : umma_idesc_bf16(128, 256);
        for (int it = 0; it < iters; ++it) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(it & 1) * 256;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (KIND == 0) umma_i8(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
                else umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, 1u);
            }
        }
        umma_commit(smem_addr(&bar));
        bar_wait(smem_addr(&bar), 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace b200grbm

using namespace b200grbm;

// ops_per_s_out (host pointer) receives 2 * MACs / s over the whole device; synchronises the stream.
extern "C" int32_t b200grbm_tensor_peak(int32_t kind, int32_t iters, double *ops_per_s_out, void *stream)
{
    if ((kind != 0 && kind != 1) || iters <= 0 || ops_per_s_out == nullptr)
        return fail(B200GRBM_EINVAL, "tensor_peak: kind=%d (0 = int8, 1 = bf16) iters=%d", kind, iters);
    B200_TRY(require_device());
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const size_t smem = 16384 + 32768 + 1024;
    if (kind == 0) B200_CUDA(cudaFuncSetAttribute(tc_peak_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else B200_CUDA(cudaFuncSetAttribute(tc_peak_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    B200_CUDA(cudaEventCreate(&e0));
    B200_CUDA(cudaEventCreate(&e1));
    for (int rep = 0; rep < 2; ++rep) {      // first launch warms up
        B200_CUDA(cudaEventRecord(e0, st));
        if (kind == 0) tc_peak_kernel<0><<<sms, 128, smem, st>>>(iters);
        else tc_peak_kernel<1><<<sms, 128, smem, st>>>(iters);
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaEventRecord(e1, st));
        B200_CUDA(cudaEventSynchronize(e1));
    }
    float ms = 0.f;
    B200_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double k_per_mma = kind == 0 ? 32.0 : 16.0;
    *ops_per_s_out = 2.0 * 128.0 * 256.0 * k_per_mma * 4.0 * (double)iters * (double)sms / ((double)ms * 1e-3);
    return 0;
}
