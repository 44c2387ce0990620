// Heat-bath sweeps for SMALL problems: one chain per lane-column, every round's neighbour table in registers.
//
// The reference's own default call is tiny: sampler.sample_ising(h, J, num_reads=256) on a 256-spin sub-graph of the
// QPU (src/training_parameters.yaml:1-13, call sites src/model_wrapper.py:309-316).  That is 6.5e4 spins in flight:
// nothing to keep 148 SMs busy with, so the launch is pure LATENCY -- (sweeps x colour rounds) dependent steps, each
// "read neighbours -> add fields in contract order -> draw -> decide -> write -> barrier".  The throughput kernel
// (gibbs.cu: 4..28 chains bit-packed per lane, tables streamed per round) spends ~1 260 cycles per round there,
// almost all of it one warp's dependent instruction stream.  This kernel minimises the round's critical path instead:
//   * a CTA is only as wide as one colour round (T lanes, e.g. 64 for the 256-spin graph) times CPC = 1 or 2 chains;
//     lane s handles spin (first position of round r) + s in EVERY round r, so no warp ever idles through a barrier
//     and the barrier joins two to four warps instead of sixteen;
//   * the lane's rows of the sampler tables for ALL rounds -- f0, and per neighbour slot 2J and the neighbour's state
//     address -- are loaded once into registers (R x (1 + W + W/2), R <= 5 rounds, W <= 20 slots: <= 155 registers);
//     the sweep loop is unrolled over the rounds;
//   * the state lives in shared memory as fp32 0.0 / 1.0, so a neighbour interaction is one LDS and one
//     fma(s, 2J, f): exact (s is 0 or 1: the result is f or the correctly rounded f + 2J, the contract's add) and
//     free of predicate set-up; the W loads are issued together, the W fmas follow in the contract's order;
//   * the Philox calls are made by PRODUCER warps of the same CTA one sweep ahead (a two-warp CTA leaves two of the SM's
//     four schedulers idle: the producers live there) and handed over through shared memory, so the generator's ten
//     dependent multiply rounds are not part of the round either;
//   * chains spread over CTAs (256 reads -> 256 two-warp CTAs), every CTA runs the same number of dependent rounds.
// Same contract, same Philox counters (position, chain / 8, sweep, stream), same packed / int8 I/O formats as
// gibbs_kernel<4, MODE>: trajectories are bit-identical to the throughput kernel and to the CPU oracle.
#include "gibbs_common.cuh"

#include <cstdlib>

namespace b200grbm {

constexpr int SMALL_W = 20;        // neighbour slots held in registers (Zephyr / Advantage2 sub-graphs: degree <= 20)
constexpr int SMALL_R = 5;         // colour rounds held in registers
constexpr int SMALL_T = 256;       // widest CTA

// contract arithmetic for a decision inside its bracket: out of line, it runs for ~2e-5 of the decisions
template <int MODE>
__device__ __noinline__ uint32_t small_fix(float f, float coef, float vm, uint32_t pp, uint32_t blk8, uint32_t sweep, int hw,
                                           const SweepParams &p)
{
    float v = vm;                       // supplied uniforms are already the full value
    if (MODE == MODE_PHILOX_EXACT) {
        uint32_t rr[4], q[4];
        philox4x32(pp, blk8, sweep, B200GRBM_STREAM_SWEEP, p, rr);
        philox4x32(pp, blk8, sweep, B200GRBM_STREAM_SWEEP_LO, p, q);
        const int w = hw >> 1, sh = 16 * (hw & 1);
        const uint32_t wr = w == 0 ? rr[0] : w == 1 ? rr[1] : w == 2 ? rr[2] : rr[3];
        const uint32_t wq = w == 0 ? q[0] : w == 1 ? q[1] : w == 2 ? q[2] : q[3];
        v = uniform_from_m23((((wr >> sh) & 0xffffu) << 7) | (((wq >> sh) & 0xffffu) >> 9));
    }
    return accept_exact(f, coef, v) ? 1u : 0u;
}

// explicit shared-space accesses: a 32-bit shared address in a register, no generic-address arithmetic in the round
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

__device__ __forceinline__ void cta_barrier() { asm volatile("bar.sync 0;" ::: "memory"); }

// Thread roles: threads [0, T x CPC) are CONSUMERS (one per lane slot and chain: they update spins); in the Philox modes
// the same number of PRODUCER threads follows -- warps on the other schedulers that draw, during round r of sweep t, the
// uniform consumer lane j will need in round r of sweep t + 1, and leave its 16 high bits in shared memory.  The
// generator's ten dependent multiply rounds (~110 cycles) are thereby off the round's critical path, which is
// LDS -> W fmas -> decision -> STS -> barrier.  The two roles run separate loops that meet at the same CTA barrier
// (bar.sync 0 counts arrivals, whichever instruction they come from).
template <int MODE, int CPC>
__global__ void __launch_bounds__(SMALL_T, 1) gibbs_small_kernel(const __grid_constant__ SweepParams p, int threads_tile)
{
    extern __shared__ __align__(16) float smem_f[];                // state [n][CPC] fp32 (1.0 = up), then uniforms [2][R][T x CPC] u32
    const int n_cons = threads_tile * CPC;
    const bool producer = (int)threadIdx.x >= n_cons;
    const int lt = producer ? (int)threadIdx.x - n_cons : (int)threadIdx.x;
    const int s = lt / CPC, c = lt % CPC;                          // lane slot inside a round, chain of this CTA
    const int chain = blockIdx.x * CPC + c;                        // local to the call
    const uint32_t gchain4 = p.chain_block0 + (uint32_t)(chain >> 2);      // global chain / 4
    const uint32_t gsub = (uint32_t)(chain & 3);                           // chain_offset is a multiple of 4
    const uint32_t blk8 = gchain4 >> 1;
    const int hw = (int)(((gchain4 & 1u) << 2) | gsub);                    // halfword of the Philox call: global chain & 7
    const int R = p.n_tiles;
    uint32_t state_addr;
    asm volatile("mov.u32 %0, %1;" : "=r"(state_addr) : "r"(smem_u32(smem_f)));      // held in a register, never re-derived
    const uint32_t ubuf_addr = state_addr + (uint32_t)(((size_t)p.n * CPC * 4 + 15) / 16 * 16);
    const uint32_t ubuf_stride = (uint32_t)(SMALL_R * n_cons * 4);         // one sweep's uniforms

    if (producer) {
        // ===================== producers: the uniforms of sweep t + 1 while the consumers run sweep t =====================
        int pos[SMALL_R];
#pragma unroll
        for (int r = 0; r < SMALL_R; ++r) {
            pos[r] = -1;
            if (r < R) {
                const int2 info = __ldg(p.tile_info + r);
                if (s < info.y) pos[r] = info.x + s;
            }
        }
        const auto draw16 = [&](uint32_t pp, uint32_t sweep) {      // 16 high bits of the uniform of (position, chain, sweep)
            uint32_t rr[4];
            philox4x32(pp, blk8, sweep, B200GRBM_STREAM_SWEEP, p, rr);
            const int w = hw >> 1;
            const uint32_t word = w == 0 ? rr[0] : w == 1 ? rr[1] : w == 2 ? rr[2] : rr[3];
            return (hw & 1) ? (word >> 16) : (word & 0xffffu);
        };
        if (p.num_sweeps > 0) {
#pragma unroll
            for (int r = 0; r < SMALL_R; ++r)
                if (pos[r] >= 0) sts_u32(ubuf_addr + (uint32_t)((r * n_cons + lt) * 4), draw16((uint32_t)pos[r], p.sweep_offset));
        }
        cta_barrier();
#pragma unroll 1
        for (int t = 0; t < p.num_sweeps; ++t) {
            const uint32_t u_next = ubuf_addr + (uint32_t)((t + 1) & 1) * ubuf_stride + (uint32_t)(lt * 4);
#pragma unroll
            for (int r = 0; r < SMALL_R; ++r) {
                if (r >= R) break;
                if (pos[r] >= 0 && t + 1 < p.num_sweeps)
                    sts_u32(u_next + (uint32_t)(r * n_cons * 4), draw16((uint32_t)pos[r], p.sweep_offset + (uint32_t)t + 1u));
                cta_barrier();
            }
        }
        return;
    }

    // ===================== consumers =====================
    // this lane's rows of the tables, every round, once
    float fz[SMALL_R], j2[SMALL_R][SMALL_W];
    uint32_t nb[SMALL_R][SMALL_W];              // shared addresses of the neighbours' state entries
    uint32_t own[SMALL_R];                      // shared address of the spin handled in round r, 0 = none
    const uint32_t base = 128u + p.info_bytes;  // b200grbm_sweep_state_offset(n_tiles): what the tables' .nbr fields count from
#pragma unroll
    for (int r = 0; r < SMALL_R; ++r) {
        own[r] = 0u;
        fz[r] = 0.f;
#pragma unroll
        for (int k = 0; k < SMALL_W; ++k) { j2[r][k] = 0.f; nb[r][k] = state_addr; }
        if (r < R) {
            const int2 info = __ldg(p.tile_info + r);
            if (s < info.y) {
                own[r] = state_addr + (uint32_t)(((info.x + s) * CPC + c) * 4);
                const uint2 *row = p.tiles + ((size_t)r * (p.width + 1)) * threads_tile + s;
                fz[r] = u2f(__ldg(&row->x));
#pragma unroll
                for (int k = 0; k < SMALL_W; ++k) {
                    nb[r][k] = own[r];                                                      // padding: own state, 2J = 0
                    if (k < p.width) {
                        const uint2 e = __ldg(row + (size_t)(k + 1) * threads_tile);
                        j2[r][k] = u2f(e.x);
                        nb[r][k] = state_addr + (((e.y - base) >> 2) * CPC + (uint32_t)c) * 4u;
                    }
                }
            }
        }
    }
    const auto position = [&](uint32_t own_addr) { return (int)(((own_addr - state_addr) >> 2) / CPC); };

    // initial state
#pragma unroll
    for (int r = 0; r < SMALL_R; ++r) {
        if (own[r] == 0u) continue;
        const int pp = position(own[r]);
        uint32_t bit;
        if (p.state_in != nullptr) {
            bit = chain < p.chains ? (p.state_in[(size_t)chain * p.n + p.order[pp]] > 0 ? 1u : 0u) : 0u;
        } else if (p.packed_in != nullptr) {
            bit = (p.packed_in[(size_t)(chain >> 2) * p.n_pad + pp] >> (chain & 3)) & 1u;
        } else {
            uint32_t rr[4];
            philox4x32((uint32_t)pp, gchain4, 0u, B200GRBM_STREAM_INIT, p, rr);
            bit = (gsub == 0 ? rr[0] : gsub == 1 ? rr[1] : gsub == 2 ? rr[2] : rr[3]) >> 31;
        }
        sts_f32(own[r], bit ? 1.0f : 0.0f);
    }
    cta_barrier();

    float coef = p.num_sweeps > 0 ? __ldg(p.coef) : 0.f;
    float coef_next = p.num_sweeps > 1 ? __ldg(p.coef + 1) : 0.f;       // one sweep ahead: never waited on
#pragma unroll 1
    for (int t = 0; t < p.num_sweeps; ++t) {
        const uint32_t u_this = ubuf_addr + (uint32_t)(t & 1) * ubuf_stride + (uint32_t)(lt * 4);
#pragma unroll
        for (int r = 0; r < SMALL_R; ++r) {
            if (r >= R) break;
            if (own[r] != 0u) {
                // neighbour spins first (independent loads), then the field in the contract's order
                float sv[SMALL_W];
#pragma unroll
                for (int k = 0; k < SMALL_W; ++k) sv[k] = lds_f32(nb[r][k]);
                float vm;
                if (MODE != MODE_SUPPLIED_EXACT)
                    vm = __fadd_rn(u2f(0x43000000u | lds_u32(u_this + (uint32_t)(r * n_cons * 4))), -(128.0f - 0x1.0p-17f));   // (h16 + 1/2) 2^-16
                else
                    vm = __ldg(p.uniforms + ((size_t)t * p.chains + min(chain, p.chains - 1)) * p.n + position(own[r]));
                float f = fz[r];
#pragma unroll
                for (int k = 0; k < SMALL_W; ++k) f = __fmaf_rn(sv[k], j2[r][k], f);      // s = 0 or 1: f, or round(f + 2J)
                uint32_t unsure = 0;
                const uint32_t dbits = MODE == MODE_SUPPLIED_EXACT ? decide_quick<true, true>(f, coef, vm, unsure)
                                       : MODE == MODE_PHILOX_FAST  ? decide_quick<false>(f, coef, vm, unsure)
                                                                   : decide_quick<true>(f, coef, vm, unsure);
                uint32_t bit = dbits >> 31;
                if (MODE != MODE_PHILOX_FAST && (unsure & 1u))
                    bit = small_fix<MODE>(f, coef, vm, (uint32_t)position(own[r]), blk8, p.sweep_offset + (uint32_t)t, hw, p);
                sts_f32(own[r], bit ? 1.0f : 0.0f);
            }
            cta_barrier();
        }
        coef = coef_next;
        if (t + 2 < p.num_sweeps) coef_next = __ldg(p.coef + t + 2);
    }

    // write back: int8 node order, and the bit-packed words of the 4-chain group layout (gibbs_kernel<4>)
#pragma unroll
    for (int r = 0; r < SMALL_R; ++r) {
        if (own[r] == 0u || chain >= p.chains) continue;
        const int pp = position(own[r]);
        const uint32_t bit = lds_f32(own[r]) != 0.0f ? 1u : 0u;
        if (p.state_out != nullptr) p.state_out[(size_t)chain * p.n + p.order[pp]] = bit ? (int8_t)1 : (int8_t)-1;
        if (p.packed_out != nullptr) {
            // chains 4g .. 4g+3 share a word and live in different CTAs.  Each chain rewrites only its own bit
            // (packed_out may alias packed_in: persistent chains); the group's first chain also clears the bits no
            // chain owns; all of these commute across CTAs.
            uint32_t *w = p.packed_out + (size_t)(chain >> 2) * p.n_pad + pp;
            const int in_group = min(4, p.chains - (chain & ~3));
            const uint32_t valid = (1u << in_group) - 1u;
            atomicAnd(w, (chain & 3) == 0 ? ((valid & 0xeu) | bit) : ~((bit ^ 1u) << (chain & 3)));
            if (bit) atomicOr(w, 1u << (chain & 3));
        }
    }
}

typedef void (*small_fn)(const SweepParams, int);

template <int CPC>
static small_fn pick_small_mode(int mode)
{
    switch (mode) {
        case MODE_PHILOX_EXACT: return gibbs_small_kernel<MODE_PHILOX_EXACT, CPC>;
        case MODE_PHILOX_FAST: return gibbs_small_kernel<MODE_PHILOX_FAST, CPC>;
        default: return gibbs_small_kernel<MODE_SUPPLIED_EXACT, CPC>;
    }
}

// Chains per CTA for the one-chain-per-lane kernel, or 0 when the problem does not fit it (more than SMALL_R colour
// rounds, degree above SMALL_W, rounds wider than half a CTA) or is large enough for the
// throughput kernel: a latency kernel only pays while all of its CTAs are resident at once.
int small_kernel_cpc(int n, int width, int n_tiles, int threads_tile, int chains, int sms)
{
    if (width > SMALL_W || n_tiles > SMALL_R || 2 * threads_tile > SMALL_T) return 0;
    const char *env = getenv("B200GRBM_SMALL_CPC");             // A/B measurements: force 1 or 2 chains per CTA
    for (int cpc = 1; cpc <= 2; ++cpc) {
        if (2 * threads_tile * cpc > SMALL_T) continue;             // consumers + as many producer threads
        if (env != nullptr && env[0] - '0' == cpc) return cpc;
        if (env == nullptr && (chains + cpc - 1) / cpc <= 2 * sms) return cpc;
    }
    return 0;
}

int32_t launch_gibbs_small(const SweepParams &p, int mode, int threads_tile, int cpc, cudaStream_t st)
{
    small_fn fn = cpc == 2 ? pick_small_mode<2>(mode) : pick_small_mode<1>(mode);
    const int n_cons = threads_tile * cpc;
    const bool producers = mode != MODE_SUPPLIED_EXACT;
    const size_t smem = ((size_t)p.n * cpc * sizeof(float) + 15) / 16 * 16 + (producers ? (size_t)2 * SMALL_R * n_cons * 4 : 0);
    fn<<<(p.chains + cpc - 1) / cpc, producers ? 2 * n_cons : n_cons, smem, st>>>(p, threads_tile);
    B200_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace b200grbm
