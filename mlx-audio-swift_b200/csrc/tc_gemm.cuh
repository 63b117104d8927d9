// tcgen05 + TMA "weights-as-A" GEMM for sm_100a:  D[M, N] = W[M, K] * X[N, K]^T   (bf16 in, fp32 accumulate)
//   W : weights, [M, K] row-major (K-major)  -> A operand, 128 x 64 tiles by TMA (SWIZZLE_128B)
//   X : activations, [N, K] row-major        -> B operand, BN x 64 tiles by TMA
//   D : accumulated in TMEM (128 lanes x BN fp32 columns, double buffered), written TRANSPOSED as
//       out[n, m] so it is the next layer's [tokens, features] activation matrix.
// Swap-AB keeps the tensor-core M dimension (128) full with output features even when there are only
// 8 tokens (autoregressive decode), where the kernel is purely a weight-streaming engine: one elected
// thread issues TMA into a deep shared-memory ring (~200 KB in flight per SM), one elected thread issues
// tcgen05.mma, four warps drain TMEM.  Work is split stream-K style: the (m_tile, k_block) units are
// dealt evenly and contiguously to the CTAs; a CTA that owns only part of a tile's K range adds its
// partial into the (zeroed) fp32 output with red.global.add.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2a {
namespace tc {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
enum : int { EPI_STORE = 0, EPI_ATOMIC = 1, EPI_SWIGLU = 2, EPI_STORE_BF16 = 3, EPI_ADD = 4 };
enum : int { ACT_NONE = 0, ACT_GELU = 1 };

struct Args {
    float* out_f32;          // [N, ldo] fp32 (EPI_STORE / EPI_ATOMIC)
    __nv_bfloat16* out_bf16; // [N, ldo] bf16 (EPI_SWIGLU: ldo = M/2 ; EPI_STORE_BF16)
    int M, N, K;             // N = valid tokens (rows of X)
    int ldo;                 // leading dimension of the output (elements)
    int m_tiles, k_blocks;   // ceil(M/128), K/64
    int stages;              // smem ring depth
    int epi_full;            // epilogue when a CTA owns a tile's whole K range
    int epi_partial;         // epilogue for partial K ranges (EPI_ATOMIC); < 0 => CTAs own whole tiles only
    int hilo;                // X rows [0, BN/2) = hi(x), rows [BN/2, BN) = lo(x) = bf16(x - hi): columns j and
                             // j + BN/2 of the accumulator are summed, giving fp32-activation accuracy for free
    const float* bias;       // nullable [M]: added to every output column before the activation
    int act;                 // ACT_NONE | ACT_GELU (exact erf GELU, WhisperLayers.swift:101)
    int tile_rows;           // whole-tile mode (epi_partial < 0) only: weight rows per m-tile when != 0 (a multiple of 8, <= 128; tmA's box must
                             // have this many rows).  16384 gate/up rows are 128 tiles of 128 -- 20 of the 148 SMs idle and every CTA streams
                             // 128 rows; 147 tiles of 112 rows use all SMs with 12.5 % fewer bytes per CTA.  The MMA still multiplies 128
                             // shared-memory rows; rows past tile_rows are stale and their accumulator rows are never stored.
    const void* pf_ptr;      // optional L2 prefetch of a later GEMM's weights (issued by the epilogue warps at kernel start)
    long long pf_bytes;
    const float* rstd_ss;    // nullable [rstd_parts, 8]: the X rows are UN-normalised (h * gain); every accumulator column t is
    int rstd_parts;          // multiplied by rsqrt(sum_p rstd_ss[p, t] * rstd_inv_h + rstd_eps) first (fused RMSNorm, BN = 16 only)
    float rstd_inv_h, rstd_eps;
    int lo_rows;             // != 0: bf16 outputs are written as hi/lo pairs in the same tile-interleaved row layout the
                             // kernel reads X in: token t -> hi row (t / (BN/2)) * BN + t % (BN/2), lo row = hi row + BN/2
};

// ----------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // bounded spin: a protocol bug traps (cudaErrorLaunchFailure) instead of hanging the GPU
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 28)) __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 1024/16 | version [46,48) = 1 | layout [61,64) = 2
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
// a/b K-major (0), n_dim [17,23) = N>>3, m_dim [24,29) = M>>4
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// Epilogue warps per CTA.  The 128-column tile (prefill / Whisper encoder: 64 tokens as hi/lo pairs) gave each of 4 epilogue warps
// 32 rows x 128 columns to read from TMEM, add, activate and store -- ~5 us per tile against 1.1 us of MMA, so the whole kernel ran at
// the epilogue's pace (tensor pipe 22 % active).  16 warps split the columns four ways (the same fix conv_gemm.cuh needed).
template <int BN>
struct Cfg {
    static constexpr int EPI_WARPS = BN >= 128 ? 16 : 4;
    static constexpr int NTHREADS = 64 + 32 * EPI_WARPS;
};

template <int BN>
struct Smem {
    static constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE = A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = (2 * BN) <= 32 ? 32 : (2 * BN) <= 64 ? 64 : (2 * BN) <= 128 ? 128 : (2 * BN) <= 256 ? 256 : 512;
    static int max_stages() { return (227 * 1024 - 1024 - 256) / STAGE; }
    static size_t bytes(int stages) { return 1024 + (size_t)stages * STAGE + 256; }
};

template <int BN>
__global__ void __launch_bounds__(Cfg<BN>::NTHREADS, Cfg<BN>::EPI_WARPS > 4 ? 1 : 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Args a) {
    using S = Smem<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * S::STAGE);
    uint64_t* empty = full + a.stages;
    uint64_t* tfull = empty + a.stages;   // [2]
    uint64_t* tempty = tfull + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * BN;
    const int TR = a.tile_rows > 0 ? a.tile_rows : BM;                  // weight rows per m-tile
    const uint32_t stage_tx = (uint32_t)(TR * BK * 2 + S::B_BYTES);     // bytes the two TMA loads of a stage deliver
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: let the next kernel's prologue start

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], Cfg<BN>::EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<S::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // stream-K partition of the (m_tile, k_block) units, k fastest
    const long long units = (long long)a.m_tiles * a.k_blocks;
    long long u0, u1;
    if (a.epi_partial >= 0) {
        u0 = units * blockIdx.x / gridDim.x;
        u1 = units * (blockIdx.x + 1) / gridDim.x;
    } else {   // whole tiles per CTA (epilogues that cannot be split along K, e.g. SwiGLU)
        u0 = ((long long)a.m_tiles * blockIdx.x / gridDim.x) * a.k_blocks;
        u1 = ((long long)a.m_tiles * (blockIdx.x + 1) / gridDim.x) * a.k_blocks;
    }

    if (warp == 0) {
        if (lane == 0) {
            // Weights never depend on the previous kernel: fill the whole ring with A tiles right away, then wait
            // for the previous kernel (PDL) before loading the activation (B) tiles that complete those stages.
            const int npre = (int)min((long long)a.stages, u1 - u0);
            for (int i = 0; i < npre; ++i) {
                const long long u = u0 + i;
                const int mt = (int)(u / a.k_blocks), kb = (int)(u - (long long)mt * a.k_blocks);
                mbar_arrive_expect_tx(&full[i], stage_tx);
                tma_load_2d(smem + (size_t)i * S::STAGE, &tmA, &full[i], kb * BK, mt * TR);
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            for (int i = 0; i < npre; ++i) {
                const long long u = u0 + i;
                const int mt = (int)(u / a.k_blocks), kb = (int)(u - (long long)mt * a.k_blocks);
                tma_load_2d(smem + (size_t)i * S::STAGE + S::A_BYTES, &tmB, &full[i], kb * BK, n0);
            }
            int stage = npre == a.stages ? 0 : npre;
            uint32_t phase = npre == a.stages ? 1 : 0;
            for (long long u = u0 + npre; u < u1; ++u) {
                const int mt = (int)(u / a.k_blocks), kb = (int)(u - (long long)mt * a.k_blocks);
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + (size_t)stage * S::STAGE;
                mbar_arrive_expect_tx(&full[stage], stage_tx);
                tma_load_2d(sa, &tmA, &full[stage], kb * BK, mt * TR);
                tma_load_2d(sa + S::A_BYTES, &tmB, &full[stage], kb * BK, n0);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            long long u = u0;
            while (u < u1) {
                const int mt = (int)(u / a.k_blocks);
                const long long seg_end = min(u1, (long long)(mt + 1) * a.k_blocks);
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * BN);
                bool first = true;
                for (; u < seg_end; ++u) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * S::STAGE);
                    const uint64_t ad = make_smem_desc(sa), bd = make_smem_desc(sa + S::A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        umma_bf16(d, ad + (uint64_t)(k * UMMA_K * 2 / 16), bd + (uint64_t)(k * UMMA_K * 2 / 16), idesc,
                                  (first && k == 0) ? 0u : 1u);
                    }
                    first = false;
                    umma_commit(&empty[stage]);          // frees the smem slot when these MMAs retire
                    if (++stage == a.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);                // accumulator of this segment complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        if (a.pf_ptr) {
            constexpr long long CH = 8192;
            constexpr int NE = 32 * Cfg<BN>::EPI_WARPS;
            const long long w = (long long)blockIdx.x * NE + (threadIdx.x - 64), nw = (long long)gridDim.x * NE;
            for (long long off = w * CH; off < a.pf_bytes; off += nw * CH) {
                const unsigned n = (unsigned)(a.pf_bytes - off < CH ? ((a.pf_bytes - off) & ~15ll) : CH);
                if (n) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const char*)a.pf_ptr + off), "r"(n) : "memory");
            }
        }
        const int q = warp & 3;                           // TMEM lane quadrant this warp may access
        constexpr int NCG = Cfg<BN>::EPI_WARPS / 4;       // column groups: warps 2 + 4 g .. 5 + 4 g take every NCG-th column chunk
        const int cg = (warp - 2) >> 2;
        int acc = 0; uint32_t acc_phase = 0;
        long long u = u0;
        float rstd[8];
        bool have_rstd = false;
        if (BN == 16 && a.rstd_ss) {
            // fused RMSNorm: the producer of X left h * gain un-normalised plus per-m-tile sums of squares (SplitArgs::ss); the
            // accumulator column of token t is scaled by rstd[t].  Computed while the main loop streams weights: these warps are
            // idle until the first accumulator is ready, so they wait for the previous kernel themselves and reduce the partial
            // sums with one independent load per lane and step (lane = part * 8 + token).
            asm volatile("griddepcontrol.wait;" ::: "memory");
            float t = 0.f;
            for (int p0 = 0; p0 < a.rstd_parts; p0 += 4) {
                const int p = p0 + (lane >> 3);
                if (p < a.rstd_parts) t += a.rstd_ss[p * 8 + (lane & 7)];
            }
            t += __shfl_xor_sync(0xffffffffu, t, 8);
            t += __shfl_xor_sync(0xffffffffu, t, 16);
            const float rs = rsqrtf(t * a.rstd_inv_h + a.rstd_eps);
#pragma unroll
            for (int j = 0; j < 8; ++j) rstd[j] = __shfl_sync(0xffffffffu, rs, j);
            have_rstd = true;
        }
        while (u < u1) {
            const int mt = (int)(u / a.k_blocks);
            const long long seg_begin = u;
            const long long seg_end = min(u1, (long long)(mt + 1) * a.k_blocks);
            u = seg_end;
            const bool whole = (seg_begin == (long long)mt * a.k_blocks) && (seg_end == (long long)(mt + 1) * a.k_blocks);
            const int epi = whole ? a.epi_full : a.epi_partial;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int m = mt * TR + q * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            constexpr int HALF = BN / 2;
            constexpr int CH = (BN == 16) ? 8 : 16;            // token columns handled per iteration
            const int n_cols = a.hilo ? HALF : BN;              // hilo: column j of the hi half pairs with j + HALF
            const int ntok0 = a.hilo ? (int)blockIdx.y * HALF : n0;
            const bool m_ok = m < a.M && q * 32 + lane < TR;
            const int cstep = (BN == 16 && !a.hilo) ? 16 : CH;
            for (int c0 = cg * cstep; c0 < n_cols; c0 += cstep * NCG) {
                float v[16];
                if (BN == 16) {
                    tmem_ld16(taddr, v);                         // all 16 columns: [0,8) hi, [8,16) lo
                    if (a.hilo) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] += v[j + 8];
                    }
                    if (have_rstd) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] *= rstd[j];
                    }
                } else {
                    tmem_ld16(taddr + c0, v);
                    if (a.hilo) {
                        float w[16];
                        tmem_ld16(taddr + c0 + HALF, w);
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += w[j];
                    }
                }
                const int jn = (BN == 16 && !a.hilo) ? 16 : CH;
                if (c0 + cstep * NCG >= n_cols) {                // this warp's last TMEM read of the accumulator: hand it back early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty[acc]);
                }
                if (a.bias || a.act) {
                    // a partial K range (stream-K red.add) carries the bias only in the segment that starts the tile
                    const float bv = (a.bias && m_ok && (whole || seg_begin == (long long)mt * a.k_blocks)) ? a.bias[m] : 0.f;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float t = v[j] + bv;
                        if (a.act == ACT_GELU) t = 0.5f * t * (1.0f + erff(t * 0.70710678118654752f));
                        v[j] = t;
                    }
                }
                const int nt = ntok0 + c0;                       // first token of this chunk
                const int nvalid = min(jn, a.N - nt);            // tokens of the chunk that exist
                if (epi == EPI_SWIGLU) {
                    // rows are (gate, up) pairs: even lane = gate, odd lane = up   (LlamaTTS.swift:282-284)
                    // the chunk lies inside one HALF block, so hi rows are consecutive
                    const long long hrow = a.lo_rows ? (long long)(nt / HALF) * BN + (nt % HALF) : nt;
                    __nv_bfloat16* ph = a.out_bf16 + hrow * a.ldo + (m >> 1);
                    const long long lo_off = (long long)HALF * a.ldo;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (j >= jn) break;
                        const float other = __shfl_xor_sync(0xffffffffu, v[j], 1);
                        if (j < nvalid && m_ok && (lane & 1) == 0) {
                            const float r = v[j] / (1.0f + __expf(-v[j])) * other;
                            const __nv_bfloat16 hi = __float2bfloat16_rn(r);
                            ph[0] = hi;
                            if (a.lo_rows) ph[lo_off] = __float2bfloat16_rn(r - __bfloat162float(hi));
                        }
                        ph += a.ldo;
                    }
                } else if (epi == EPI_STORE_BF16) {
                    const long long hrow = a.lo_rows ? (long long)(nt / HALF) * BN + (nt % HALF) : nt;
                    __nv_bfloat16* ph = a.out_bf16 + hrow * a.ldo + m;
                    const long long lo_off = (long long)HALF * a.ldo;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (j >= jn) break;
                        if (j < nvalid && m_ok) {
                            const __nv_bfloat16 hi = __float2bfloat16_rn(v[j]);
                            ph[0] = hi;
                            if (a.lo_rows) ph[lo_off] = __float2bfloat16_rn(v[j] - __bfloat162float(hi));
                        }
                        ph += a.ldo;
                    }
                } else {
                    float* pf = a.out_f32 + (long long)nt * a.ldo + m;
                    if (epi == EPI_ADD) {     // residual accumulate (whole-tile CTAs only): all loads before any store
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j < jn && j < nvalid && m_ok) v[j] += pf[(long long)j * a.ldo];
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (j >= jn) break;
                        if (j < nvalid && m_ok) {
                            if (epi == EPI_STORE || epi == EPI_ADD) pf[0] = v[j];
                            else atomicAdd(pf, v[j]);
                        }
                        pf += a.ldo;
                    }
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<S::TMEM_COLS>(tmem_base);
    }
}

// ----------------------------------------------------------------------------------------------- cluster split-K
// D[M, 8 tokens] = W[M, K] X^T for the two GEMMs of a decoder layer whose output feeds the residual stream (o_proj, down_proj),
// with the residual add, the NEXT RMSNorm's gain, its hi/lo split and its sum of squares fused into the epilogue, so the step
// has no stand-alone norm kernel.  One thread-block CLUSTER per 128-row m-tile: CTA r of the cluster streams k-blocks
// [kb * r / C, kb * (r + 1) / C) of the tile (same TMA ring / single-thread tcgen05 issue as tc_gemm_kernel<16>), then every
// non-leader writes its 128 x 8 fp32 partial into the leader's shared memory (distributed shared memory, st.shared::cluster),
// one cluster barrier, and the leader adds the partials IN RANK ORDER (bit-reproducible, unlike stream-K's red.add) and runs
//     h[t, m] += acc          (fp32 residual stream, updated in place)
//     xn[t, m] = hi / lo of  h[t, m] * gain[m]        (UN-normalised: the consumer GEMM scales its accumulator by rstd[t])
//     ss[mt, t] = sum over the tile's rows of h[t, m]^2   (the consumer reduces the m-tiles: rstd = rsqrt(sum / H + eps))
// RMSNorm is linear in its per-token scale, so moving rstd behind the consumer GEMM is exact up to fp32 rounding.
struct SplitArgs {
    int M, N, K;                 // N = tokens (<= 8)
    int k_blocks, stages;
    float* h;                    // [8, M] residual stream (read-modify-write by the leader)
    const float* gain;           // [M] the next norm's weight
    __nv_bfloat16* xn;           // [16, M] hi rows 0..7 / lo rows 8..15
    float* ss;                   // [m_tiles, 8] partial sums of squares
    const float* rstd_ss;        // nullable: partial sums [rstd_parts, 8] of the norm this GEMM's INPUT went through
    int rstd_parts;
    float rstd_inv_h, rstd_eps;
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

constexpr int SPLIT_MAX_CLUSTER = 8;
struct SmemSplit {
    static constexpr int STAGE = Smem<16>::STAGE;
    static constexpr int PART_BYTES = BM * 8 * 4;                      // one CTA's 128 x 8 fp32 partial
    static size_t bytes(int stages, int cluster) { return 1024 + (size_t)stages * STAGE + (size_t)(cluster - 1) * PART_BYTES + 512; }
};

#ifdef B2A_TC_GEMM_IMPL   // the kernel body lives in tc_gemm.cu only (a non-template __global__ cannot be defined in every translation unit)
__global__ void __launch_bounds__(THREADS, 2)
tc_gemm_splitk_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, SplitArgs a) {
    using S = Smem<16>;
    constexpr int BN = 16;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t C = cluster_nctarank(), rank = cluster_ctarank();
    float* part = reinterpret_cast<float*>(smem + (size_t)a.stages * S::STAGE);                 // [C - 1][128][8]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * S::STAGE + (size_t)(C - 1) * SmemSplit::PART_BYTES);
    uint64_t* empty = full + a.stages;
    uint64_t* tfull = empty + a.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);
    float* s_red = reinterpret_cast<float*>(tmem_slot + 2);                                      // [4][8] + [8]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = blockIdx.x / C;
    const int kb0 = (int)((long long)a.k_blocks * rank / C), kb1 = (int)((long long)a.k_blocks * (rank + 1) / C);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < a.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(tfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<32>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // weights first (they never depend on the previous kernel), then wait for it, then the activation tiles
            const int n = kb1 - kb0, npre = min(a.stages, n);
            for (int i = 0; i < npre; ++i) {
                mbar_arrive_expect_tx(&full[i], S::STAGE);
                tma_load_2d(smem + (size_t)i * S::STAGE, &tmA, &full[i], (kb0 + i) * BK, mt * BM);
            }
            asm volatile("griddepcontrol.wait;" ::: "memory");
            for (int i = 0; i < npre; ++i) tma_load_2d(smem + (size_t)i * S::STAGE + S::A_BYTES, &tmB, &full[i], (kb0 + i) * BK, 0);
            int stage = npre == a.stages ? 0 : npre;
            uint32_t phase = npre == a.stages ? 1 : 0;
            for (int i = npre; i < n; ++i) {
                mbar_wait(&empty[stage], phase ^ 1);
                uint8_t* sa = smem + (size_t)stage * S::STAGE;
                mbar_arrive_expect_tx(&full[stage], S::STAGE);
                tma_load_2d(sa, &tmA, &full[stage], (kb0 + i) * BK, mt * BM);
                tma_load_2d(sa + S::A_BYTES, &tmB, &full[stage], (kb0 + i) * BK, 0);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(BN);
            int stage = 0; uint32_t phase = 0;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + (size_t)stage * S::STAGE);
                const uint64_t ad = make_smem_desc(sa), bd = make_smem_desc(sa + S::A_BYTES);
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k)
                    umma_bf16(tmem_base, ad + (uint64_t)(k * UMMA_K * 2 / 16), bd + (uint64_t)(k * UMMA_K * 2 / 16), idesc, (kb == kb0 && k == 0) ? 0u : 1u);
                umma_commit(&empty[stage]);
                if (++stage == a.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(tfull);
        }
    }
    // ---- epilogue part 1 (warps 2..5): this CTA's partial = hi + lo columns; non-leaders hand it to the leader
    float acc[8], hold[8];
    float g = 0.f;
    const int q = warp & 3, row = q * 32 + lane;
    const int m = mt * BM + row;
    const bool m_ok = m < a.M;
    if (warp >= 2) {
        if (rank == 0) {
            // the residual rows and the gain do not depend on this GEMM: fetch them while the weights stream
            asm volatile("griddepcontrol.wait;" ::: "memory");
            g = m_ok ? a.gain[m] : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) hold[j] = (j < a.N && m_ok) ? a.h[(long long)j * a.M + m] : 0.f;
        }
        mbar_wait(tfull, 0);
        tc_fence_after();
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = v[j] + v[j + 8];
        if (rank != 0) {
            const uint32_t dst = map_to_rank(smem_u32(part + ((size_t)(rank - 1) * BM + row) * 8), 0);
            st_cluster_f4(dst, acc[0], acc[1], acc[2], acc[3]);
            st_cluster_f4(dst + 16, acc[4], acc[5], acc[6], acc[7]);
        }
    }
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();                                          // partials are in the leader's shared memory
    if (rank == 0 && warp >= 2) {
        for (uint32_t r = 1; r < C; ++r) {                       // fixed order: deterministic sums
            const float4 p0 = *reinterpret_cast<const float4*>(part + ((size_t)(r - 1) * BM + row) * 8);
            const float4 p1 = *reinterpret_cast<const float4*>(part + ((size_t)(r - 1) * BM + row) * 8 + 4);
            acc[0] += p0.x; acc[1] += p0.y; acc[2] += p0.z; acc[3] += p0.w;
            acc[4] += p1.x; acc[5] += p1.y; acc[6] += p1.z; acc[7] += p1.w;
        }
        const int w4 = warp - 2;                                 // 0..3 (s_red rows); q = warp & 3 is the TMEM quadrant
        // rstd of the norm this GEMM's input went through (its producer wrote un-normalised hi/lo rows)
        if (a.rstd_ss) {
            if (w4 == 0 && lane < 8) {
                float t = 0.f;
                for (int p = 0; p < a.rstd_parts; ++p) t += a.rstd_ss[p * 8 + lane];
                s_red[32 + lane] = rsqrtf(t * a.rstd_inv_h + a.rstd_eps);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] *= s_red[32 + j];
        }
        float sq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float hv = 0.f;
            if (j < a.N && m_ok) {
                hv = hold[j] + acc[j];
                a.h[(long long)j * a.M + m] = hv;
                const float t = hv * g;
                const __nv_bfloat16 hi = __float2bfloat16_rn(t);
                a.xn[(long long)j * a.M + m] = hi;
                a.xn[(long long)(8 + j) * a.M + m] = __float2bfloat16_rn(t - __bfloat162float(hi));
            }
            sq[j] = hv * hv;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
            for (int o = 16; o; o >>= 1) sq[j] += __shfl_xor_sync(0xffffffffu, sq[j], o);
        }
        if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s_red[w4 * 8 + j] = sq[j];
        }
        asm volatile("bar.sync 2, 128;" ::: "memory");
        if (w4 == 0 && lane < 8) a.ss[mt * 8 + lane] = s_red[lane] + s_red[8 + lane] + s_red[16 + lane] + s_red[24 + lane];
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<32>(tmem_base);
    }
}
#endif  // B2A_TC_GEMM_IMPL

// ----------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D bf16 row-major [rows, cols] tensor map with a {64, box_rows} box and 128-byte swizzle
CUtensorMap make_tmap_bf16(const void* base, long long rows, long long cols, int box_rows);
CUtensorMap make_tmap_f16_3d(const void* base, long long d0, long long d1, long long d2, int b0, int b1);

template <int BN>
void launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const Args& a, int ctas, int n_tiles, cudaStream_t s);

void launch_splitk(const CUtensorMap& tmA, const CUtensorMap& tmB, const SplitArgs& a, int m_tiles, int cluster, cudaStream_t s);

void set_attributes();   // cudaFuncSetAttribute for every instantiation (call once, outside graph capture)

}  // namespace tc
}  // namespace b2a
