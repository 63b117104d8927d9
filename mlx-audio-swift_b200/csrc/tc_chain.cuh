// Decode-step GEMM CHAIN for sm_100a: up to four weights-as-A tcgen05 GEMMs (tc_gemm.cuh, BN = 16, hi/lo activations) and the
// add+RMSNorm phases between them run as ONE persistent kernel, one CTA per SM.
//
// Why: the decode step is a dependency chain of ~10 us kernels.  Measured (profiles/r01_decode_step_launches.md): each kernel
// boundary costs 3-5 us even with programmatic dependent launch, the two add_rmsnorm launches 5.3 us each, and the GEMMs
// themselves only reach 0.68 of the HBM peak because every launch ramps its weight stream up from an empty ring and drains it at
// the end.  Here the ring NEVER drains inside a chain: the TMA producer thread walks the concatenated (gemm, m_tile, k_block)
// unit list of its CTA and keeps issuing WEIGHT tiles (which depend on nothing) for the next GEMM while the current one is
// still being consumed; only the small activation tile of a stage waits for the phase before it (a device-wide barrier:
// release/acquire on a global counter).  Up to the whole ring (12 x 18 KB per SM = 31 MB across the chip, 4.8 us of HBM
// time) is in flight across a barrier or a norm phase.
//
//   phases:   [norm 0] gemm 0 | [norm 1] gemm 1 | ...      ('|' = grid barrier; a norm phase is another barrier)
//   warp 0    TMA producer (one thread)          warp 1    MMA issuer (one thread) + TMEM owner
//   warps 2-5 epilogue (TMEM -> registers -> global) and the norm phases (CTA b owns row b)
//
// STATUS (round 1): EXPERIMENTAL, off by default (B2A_CHAIN=1).  Parity-green, but measured slower than the PDL chain of separate
// kernels: 3.4-3.9 ms vs 1.93 ms per step.  Phase timestamps (tools/chain_timing.py): the GEMM phases run at HBM speed (GU: 648 KB per
// SM in 13.5 us) but every full-grid barrier drains the pipeline (arrive + fence + poll = 2-3 us, ramp 3-5 us per phase), the two
// norm phases take 25 us each (128 threads per row, loads not batched by the compiler), and the 217 KB footprint removes the overlap
// with the attention kernel.  The fix is tile-level dataflow (DOWN k-block kb only needs GU m-tile kb) instead of grid barriers;
// that is the next-round design.
//
// The barrier counters of a launch live in one of three sets; a launch zeroes the two sets it does not use (nobody polls
// them while it runs: the next chain only polls after griddepcontrol.wait), so CUDA-graph replays need no memset nodes.
#pragma once
#include "tc_gemm.cuh"

namespace b2a {
namespace chain {

using namespace b2a::tc;

constexpr int BN = 16, MAX_GEMM = 4, NSETS = 3, SET_STRIDE = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE = A_BYTES + B_BYTES;    // 18 KB
constexpr int NORM_MAXV = 32;                 // hidden size <= 128 * 32

struct Gemm {
    int M, K, m_tiles, k_blocks;
    int epi_full;             // EPI_STORE (fp32) | EPI_SWIGLU (bf16 hi/lo, rows interleaved gate/up)
    int epi_partial;          // EPI_ATOMIC (stream-K) or -1 (whole tiles per CTA)
    float* out_f32;
    __nv_bfloat16* out_bf16;
    int ldo;
    int ctas;                 // CTAs that own units of this GEMM (<= gridDim.x)
    int norm;                 // index of the norm phase that runs right before this GEMM, or -1
};
struct Norm {                 // x += delta (delta zeroed); xn = hi/lo(x * rsqrt(mean x^2 + eps) * w); zero zero_ptr[b, :zero_n]
    float* x; float* delta; const float* w; __nv_bfloat16* xn; float* trace; float* zero_ptr;
    int H, zero_n; float eps;
};
struct Args {
    Gemm g[MAX_GEMM];
    Norm n[MAX_GEMM];
    int n_gemm, N, stages;
    unsigned* bars;           // [NSETS][SET_STRIDE]
    int set;                  // barrier set of this launch
    unsigned long long* dbg;  // optional [gridDim.x][64] globaltimer stamps (B2A_CHAIN_DEBUG=1): see tools/chain_timing.py
};

static inline size_t smem_bytes(int stages) { return 1024 + (size_t)stages * STAGE + 256; }

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target) {
    for (unsigned spins = 0; ld_acquire(p) < target; ++spins) {
        __nanosleep(40);                           // ~150 pollers on one address: back off so the arrivals get through
        if (spins > (1u << 24)) __trap();          // a protocol bug traps instead of hanging the GPU
    }
}
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define CH_TS(slot) do { if (a.dbg) a.dbg[(size_t)blockIdx.x * 64 + (slot)] = gtimer(); } while (0)
// phase index of "gemm i done" and of "norm before gemm i done"
__device__ __forceinline__ int ph_gemm(int i) { return 2 * i + 1; }
__device__ __forceinline__ int ph_norm(int i) { return 2 * i; }

__device__ __forceinline__ void unit_range(const Gemm& g, long long& u0, long long& u1) {
    const long long units = (long long)g.m_tiles * g.k_blocks;
    const int c = blockIdx.x;
    if (c >= g.ctas) { u0 = u1 = 0; return; }
    if (g.epi_partial >= 0) {
        u0 = units * c / g.ctas; u1 = units * (c + 1) / g.ctas;
    } else {
        u0 = ((long long)g.m_tiles * c / g.ctas) * g.k_blocks;
        u1 = ((long long)g.m_tiles * (c + 1) / g.ctas) * g.k_blocks;
    }
}

static __global__ void __launch_bounds__(THREADS, 1)
chain_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
             const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
             const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
             const __grid_constant__ CUtensorMap tmA3, const __grid_constant__ CUtensorMap tmB3, Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * STAGE);
    uint64_t* empty = full + a.stages;
    uint64_t* tfull = empty + a.stages;   // [2]
    uint64_t* tempty = tfull + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    __shared__ float nred[4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const CUtensorMap* tmA[MAX_GEMM] = {&tmA0, &tmA1, &tmA2, &tmA3};
    const CUtensorMap* tmB[MAX_GEMM] = {&tmB0, &tmB1, &tmB2, &tmB3};
    unsigned* bar = a.bars + a.set * SET_STRIDE;
    const unsigned G = gridDim.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < a.n_gemm; ++i) { tma_prefetch_desc(tmA[i]); tma_prefetch_desc(tmB[i]); }
        for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<32>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            // pending[]: stages whose weight tile is in flight but whose activation tile still waits for its phase
            int stage = 0; uint32_t phase = 0;
            int pend_stage[16], pend_kb[16], pend_g[16], n_pend = 0;
            bool dep_waited = false;          // griddepcontrol.wait done (needed before the first dependent access)
            int ready_upto = -1;              // GEMMs <= ready_upto have their activations available
            CH_TS(40);
            auto flush = [&](int gi) {        // make GEMM gi's activations available and issue every pending activation tile of it
                CH_TS(32 + 2 * gi);
                if (!dep_waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); dep_waited = true; }
                if (gi > 0 || a.g[gi].norm >= 0) {
                    const int p = a.g[gi].norm >= 0 ? ph_norm(gi) : ph_gemm(gi - 1);
                    wait_counter(bar + p, G);
                    asm volatile("fence.proxy.async;" ::: "memory");     // other CTAs' generic-proxy stores -> our TMA loads
                }
                ready_upto = gi;
                int k = 0;
                for (int i = 0; i < n_pend; ++i) {
                    if (pend_g[i] == gi) {
                        tma_load_2d(smem + (size_t)pend_stage[i] * STAGE + A_BYTES, tmB[gi], &full[pend_stage[i]], pend_kb[i] * BK, 0);
                    } else {
                        pend_stage[k] = pend_stage[i]; pend_kb[k] = pend_kb[i]; pend_g[k] = pend_g[i]; ++k;
                    }
                }
                n_pend = k;
                CH_TS(33 + 2 * gi);
            };
            // non-blocking: release the oldest unresolved GEMM if the phase it waits for is already complete
            auto try_flush = [&]() {
                if (n_pend == 0 || !dep_waited) return;
                const int g0 = pend_g[0];
                if (g0 > 0 || a.g[g0].norm >= 0) {
                    const int p = a.g[g0].norm >= 0 ? ph_norm(g0) : ph_gemm(g0 - 1);
                    if (ld_acquire(bar + p) < G) return;
                }
                flush(g0);
            };
            for (int gi = 0; gi < a.n_gemm; ++gi) {
                const Gemm& g = a.g[gi];
                long long u0, u1;
                unit_range(g, u0, u1);
                for (long long u = u0; u < u1; ++u) {
                    const int mt = (int)(u / g.k_blocks), kb = (int)(u - (long long)mt * g.k_blocks);
                    // Unresolved stages are always the most recently issued ones (GEMMs become ready in order), so the slot we
                    // are about to reuse is either free, being consumed, or -- only when the whole ring is unresolved -- pending.
                    // (a poll is an L2 round trip on the producer's critical path: only poll when the slot is not free anyway)
                    if (n_pend == a.stages) flush(pend_g[0]);          // ring full of half-loaded stages: block on the oldest
                    else if (n_pend > 0 && !mbar_try_wait(&empty[stage], phase ^ 1)) try_flush();
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * STAGE;
                    mbar_arrive_expect_tx(&full[stage], STAGE);
                    tma_load_2d(sa, tmA[gi], &full[stage], kb * BK, mt * BM);
                    if (gi <= ready_upto) {
                        tma_load_2d(sa + A_BYTES, tmB[gi], &full[stage], kb * BK, 0);
                    } else {
                        pend_stage[n_pend] = stage; pend_kb[n_pend] = kb; pend_g[n_pend] = gi; ++n_pend;
                    }
                    if (++stage == a.stages) { stage = 0; phase ^= 1; }
                }
                if (u0 == u1 && gi > ready_upto && n_pend == 0) {
                    // no units of this GEMM here: nothing to load, but later GEMMs must still see it as "ready in order"
                }
            }
            while (n_pend > 0) flush(pend_g[0]);
            CH_TS(41);
            if (!dep_waited) asm volatile("griddepcontrol.wait;" ::: "memory");
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------------------------------------------------------- MMA issuer
            constexpr uint32_t idesc = make_idesc(BN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int gi = 0; gi < a.n_gemm; ++gi) {
                const Gemm& g = a.g[gi];
                long long u, u1;
                unit_range(g, u, u1);
                while (u < u1) {
                    const int mt = (int)(u / g.k_blocks);
                    const long long seg_end = min(u1, (long long)(mt + 1) * g.k_blocks);
                    mbar_wait(&tempty[acc], acc_phase ^ 1);
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)(acc * BN);
                    bool first = true;
                    for (; u < seg_end; ++u) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(smem + (size_t)stage * STAGE);
                        const uint64_t ad = make_smem_desc(sa), bd = make_smem_desc(sa + A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_bf16(d, ad + (uint64_t)(k * UMMA_K * 2 / 16), bd + (uint64_t)(k * UMMA_K * 2 / 16), idesc,
                                      (first && k == 0) ? 0u : 1u);
                        first = false;
                        umma_commit(&empty[stage]);
                        if (++stage == a.stages) { stage = 0; phase ^= 1; }
                    }
                    umma_commit(&tfull[acc]);
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // -------------------------------------------------------------------- epilogue warps (+ norm phases)
        const int q = warp & 3, et = threadIdx.x - 64;        // et: 0..127
        int acc = 0; uint32_t acc_phase = 0;
        // Nothing below may touch global memory before every earlier kernel of the stream is complete (the previous chain may
        // still be polling its barrier set while this CTA is already resident): wait first.  The epilogue warps have nothing
        // to do before the first accumulator is complete anyway.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        bool dep_waited = true;
        if (et == 0) CH_TS(0);
        if (blockIdx.x == 0 && et < 2 * SET_STRIDE) {         // clear the two barrier sets this launch does not use
            const int other = et / SET_STRIDE, idx = et % SET_STRIDE;
            const int s2 = (a.set + 1 + other) % NSETS;
            a.bars[s2 * SET_STRIDE + idx] = 0u;
        }
        auto arrive = [&](int p) {                            // this CTA is done with phase p
            __threadfence();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) atomicAdd(bar + p, 1u);
        };
        for (int gi = 0; gi < a.n_gemm; ++gi) {
            const Gemm& g = a.g[gi];
            // ---- norm phase before this GEMM
            if (g.norm >= 0) {
                const Norm& nm = a.n[g.norm];
                if ((int)blockIdx.x < a.N) {
                    if (!dep_waited) { asm volatile("griddepcontrol.wait;" ::: "memory"); dep_waited = true; }
                    if (gi > 0) {
                        if (et == 0) wait_counter(bar + ph_gemm(gi - 1), G);
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                    }
                    if (et == 0) CH_TS(1 + 6 * gi);
                    const int b = blockIdx.x, H = nm.H;
                    float* xr = nm.x + (long long)b * H;
                    float v[NORM_MAXV];
                    float ss = 0.f;
#pragma unroll
                    for (int j = 0; j < NORM_MAXV; ++j) {
                        const int i = et + j * 128;
                        float val = 0.f;
                        if (i < H) {
                            val = __ldcg(xr + i);
                            if (nm.delta) val += __ldcg(nm.delta + (long long)b * H + i);
                        }
                        v[j] = val;
                        ss = fmaf(val, val, ss);
                    }
#pragma unroll
                    for (int j = 0; j < NORM_MAXV; ++j) {
                        const int i = et + j * 128;
                        if (i < H) {
                            if (nm.delta) { xr[i] = v[j]; nm.delta[(long long)b * H + i] = 0.f; }
                            if (nm.trace) nm.trace[(long long)b * H + i] = v[j];
                        }
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
                    if (lane == 0) nred[warp - 2] = ss;
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const float tot = nred[0] + nred[1] + nred[2] + nred[3];
                    const float r = rsqrtf(tot / (float)H + nm.eps);
#pragma unroll
                    for (int j = 0; j < NORM_MAXV; ++j) {
                        const int i = et + j * 128;
                        if (i < H) {
                            const float gv = v[j] * r * nm.w[i];
                            const __nv_bfloat16 hi = __float2bfloat16_rn(gv);
                            nm.xn[(long long)b * H + i] = hi;
                            nm.xn[(long long)(b + 8) * H + i] = __float2bfloat16_rn(gv - __bfloat162float(hi));
                        }
                    }
                    if (nm.zero_ptr)
                        for (int i = et; i < nm.zero_n; i += 128) nm.zero_ptr[(long long)b * nm.zero_n + i] = 0.f;
                    if (et == 0) CH_TS(2 + 6 * gi);
                }
                arrive(ph_norm(gi));
                if (et == 0) CH_TS(3 + 6 * gi);
            }
            // ---- GEMM epilogue
            long long u, u1;
            unit_range(g, u, u1);
            bool first_acc = true;
            while (u < u1) {
                const int mt = (int)(u / g.k_blocks);
                const long long seg_begin = u;
                const long long seg_end = min(u1, (long long)(mt + 1) * g.k_blocks);
                u = seg_end;
                const bool whole = (seg_begin == (long long)mt * g.k_blocks) && (seg_end == (long long)(mt + 1) * g.k_blocks);
                const int epi = whole ? g.epi_full : g.epi_partial;
                mbar_wait(&tfull[acc], acc_phase);
                tc_fence_after();
                if (first_acc && et == 0) { CH_TS(4 + 6 * gi); first_acc = false; }
                const int m = mt * BM + q * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
                float v[16];
                tmem_ld16(taddr, v);                              // [0,8) hi columns, [8,16) lo columns
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += v[j + 8];
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[acc]);
                const bool m_ok = m < g.M;
                const int nvalid = min(8, a.N);
                if (epi == EPI_SWIGLU) {
                    // rows are (gate, up) pairs: even lane = gate, odd lane = up   (LlamaTTS.swift:282-284)
                    __nv_bfloat16* ph = g.out_bf16 + (m >> 1);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float other = __shfl_xor_sync(0xffffffffu, v[j], 1);
                        if (j < nvalid && m_ok && (lane & 1) == 0) {
                            const float r = v[j] / (1.0f + __expf(-v[j])) * other;
                            const __nv_bfloat16 hi = __float2bfloat16_rn(r);
                            ph[(long long)j * g.ldo] = hi;
                            ph[(long long)(j + 8) * g.ldo] = __float2bfloat16_rn(r - __bfloat162float(hi));
                        }
                    }
                } else {
                    float* pf = g.out_f32 + m;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < nvalid && m_ok) {
                            if (epi == EPI_STORE) pf[(long long)j * g.ldo] = v[j];
                            else atomicAdd(pf + (long long)j * g.ldo, v[j]);
                        }
                    }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (et == 0) CH_TS(5 + 6 * gi);
            if (gi + 1 < a.n_gemm) arrive(ph_gemm(gi));           // the last GEMM ends with the kernel
            if (et == 0) CH_TS(6 + 6 * gi);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<32>(tmem_base);
    }
}

}  // namespace chain
}  // namespace b2a
