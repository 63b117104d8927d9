// Persistent tcgen05 + TMA GEMM for the codec path (SNAC in channels-last layout):
//     D[M, N] = (Wh + Wl)[M, K] * (Xh + Xl)[N, K]^T      fp32 weights and activations as bf16 hi/lo pairs
// A  = weights, TWO K-major operands (hi and lo halves of the fp32 weight, 16 mantissa bits together)
// B  = activations [tokens, channels] (NLC), 128-row tiles = 64 tokens as hi rows | lo rows (same layout as tc_gemm)
// per k-block:  D[:, 0:128] += Wh * [Xh; Xl]   and   D[:, 0:64] += Wl * Xh        (Wl*Xl, 2^-18 relative, is dropped)
// so the result equals the fp32 convolution to ~1e-5.  One CTA per SM loops over (token tile, m tile) work items;
// the TMA ring never drains between tiles.  The epilogue warps fuse what the reference runs as separate MLX ops:
// bias, Snake, residual add, NoiseBlock, the transposed-conv phase scatter, and the hi/lo re-split that feeds the
// next GEMM (optionally written twice, shifted by one token, which is the im2col the 2-tap transposed conv needs).
#pragma once
#include "tc_gemm.cuh"

namespace b2a {
namespace cg {

using namespace b2a::tc;

constexpr int BN = 128, HALF = 64;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE = 2 * A_BYTES + B_BYTES;   // 48 KB
constexpr int STAGES = 4;
// warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2..17: epilogue.  Sixteen epilogue warps = four per
// scheduler: the epilogue math (Snake, GELU, hi/lo split) is dependent-issue latency bound with one warp per scheduler
// (measured: 20 us per 64-token tile with 4 warps).  Warp w drains TMEM lane quadrant w % 4 (hardware rule) and the
// 16-token column group (w - 2) / 4.
constexpr int EPI_WARPS = 16;
constexpr int CG_THREADS = 64 + 32 * EPI_WARPS;
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE + 256;

enum : int { E_STORE_HILO = 0, E_CONVT = 1, E_NOISE = 2, E_ADD = 3, E_ADD_HILO = 4, E_STORE_F32 = 5 };

struct Args {
    int M, K, N;              // N = tokens (rows of X)
    int m_tiles, k_blocks, n_tiles;
    int epi;
    const float* bias;        // [channels] (E_CONVT: per output channel co) or null
    const float* alpha;       // Snake alpha applied to values written as hi/lo (null = identity)
    const float* gamma;       // nullable per-channel scale applied to (acc + bias) before any add (ConvNeXt layer scale)
    int gelu;                 // exact-erf GELU on (acc + bias) (Vocos pwconv1)
    float* x;                 // fp32 [tokens, ldx] read-modify-write target (E_NOISE / E_ADD / E_ADD_HILO) or E_CONVT output
    int ldx;
    __nv_bfloat16* hl;        // hi/lo output matrix (64-token tiles), leading dimension ldh
    int ldh;
    int dual;                 // hi/lo output is the 2-tap im2col of the next transposed conv:
                              //   token (b, t) -> row b*(T+1)+t cols [m], and row b*(T+1)+t+1 cols [M + m]
    int T;                    // tokens per utterance on the OUTPUT side of this GEMM (for dual / noise / convT)
    // E_CONVT: rows m = r*Cout + co; input token n = b*(Tin+1) + q  ->  t_out = q*stride + r - pad
    int Cout, stride, pad, Tin;
    // E_NOISE: x = x + noise[b, t] * acc      (NoiseBlock, Layers.swift:271-278)
    const float* noise;       // [B, T] or null => counter-based N(0,1) from seed
    unsigned long long seed;
};

// sin with an explicit two-term 2*pi range reduction + MUFU.SIN: |error| < 5e-7 for |x| < 1e4 (the libdevice sinf slow
// path costs ~40 dependent instructions per call and made every Snake epilogue issue bound)
__device__ __forceinline__ float fast_sin(float x) {
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(k, -6.28318548202514648f, x);
    r = fmaf(k, 1.7484555e-7f, r);
    return __sinf(r);
}
__device__ __forceinline__ float snake(float v, float al) {
    const float s = fast_sin(al * v);
    return v + (1.0f / (al + 1e-9f)) * s * s;
}
// same with the per-channel 1 / (alpha + 1e-9) computed once by the caller
__device__ __forceinline__ float snake_inv(float v, float al, float inv) {
    const float s = fast_sin(al * v);
    return fmaf(inv * s, s, v);
}
__device__ __forceinline__ float gauss(unsigned long long seed, unsigned long long idx) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u1 = ((unsigned)(z >> 40) + 1.0f) * (1.0f / 16777217.0f);
    const float u2 = (unsigned)((z >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}
__device__ __forceinline__ void put_hilo(__nv_bfloat16* base, long long ld, long long tok, long long col, float v) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const long long r = (tok / HALF) * BN + (tok % HALF);
    base[r * ld + col] = hi;
    base[(r + HALF) * ld + col] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

static __global__ void __launch_bounds__(CG_THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, Args a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * STAGE);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;     // [2]
    uint64_t* tempty = tfull + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const long long tiles = (long long)a.n_tiles * a.m_tiles;   // tile id = n_tile * m_tiles + m_tile, dealt round-robin

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int nt = (int)(t / a.m_tiles), mt = (int)(t - (long long)nt * a.m_tiles);
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* s0 = smem + (size_t)stage * STAGE;
                    mbar_arrive_expect_tx(&full[stage], STAGE);
                    tma_load_2d(s0, &tmA, &full[stage], kb * BK, mt * BM);
                    tma_load_2d(s0 + A_BYTES, &tmA2, &full[stage], kb * BK, mt * BM);
                    tma_load_2d(s0 + 2 * A_BYTES, &tmB, &full[stage], kb * BK, nt * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_full = make_idesc(BN), idesc_half = make_idesc(HALF);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < a.k_blocks; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t s0 = smem_u32(smem + (size_t)stage * STAGE);
                    const uint64_t ad = make_smem_desc(s0), a2d = make_smem_desc(s0 + A_BYTES), bd = make_smem_desc(s0 + 2 * A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t off = (uint64_t)(k * UMMA_K * 2 / 16);
                        umma_bf16(d, ad + off, bd + off, idesc_full, (kb == 0 && k == 0) ? 0u : 1u);   // Wh * [Xh; Xl]
                        umma_bf16(d, a2d + off, bd + off, idesc_half, 1u);                            // Wl * Xh -> columns [0, 64)
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        const int q = warp & 3, c0 = ((warp - 2) >> 2) * 16;
        int acc = 0; uint32_t acc_phase = 0;
        const bool rmw = a.epi == E_NOISE || a.epi == E_ADD || a.epi == E_ADD_HILO;
        for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
            const int nt = (int)(t / a.m_tiles), mt = (int)(t - (long long)nt * a.m_tiles);
            const int m = mt * BM + q * 32 + lane;
            const bool m_ok = m < a.M;
            float bias = 0.f, al = 0.f, gm = 1.f;
            int co = m, r = 0;
            if (a.epi == E_CONVT) { r = m / a.Cout; co = m - r * a.Cout; }
            if (m_ok) {
                if (a.bias) bias = a.bias[co];
                if (a.alpha) al = a.alpha[m];
                if (a.gamma) gm = a.gamma[m];
            }
            const float inv_al = 1.0f / (al + 1e-9f);
            const long long n_first = (long long)nt * HALF + c0;
            // everything that does not depend on the accumulator is fetched BEFORE waiting for the MMA: the residual /
            // read-modify-write operand (16 independent loads) and the NoiseBlock noise (one value per token: lane j computes
            // or loads token j, broadcast by shuffle below)
            float xv[16];
            if (rmw) {
#pragma unroll
                for (int j = 0; j < 16; ++j) xv[j] = (m_ok && n_first + j < a.N) ? a.x[(n_first + j) * a.ldx + m] : 0.f;
            }
            float nz_lane = 0.f;
            if (a.epi == E_NOISE) {
                const long long n = n_first + (lane & 15);
                if (n < a.N) nz_lane = a.noise ? a.noise[n] : gauss(a.seed, (unsigned long long)n);
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN);
            float v[16], w[16];
            tmem_ld16(taddr + c0, v);
            tmem_ld16(taddr + c0 + HALF, w);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const long long n = n_first + j;      // token (row of X)
                const float nz = a.epi == E_NOISE ? __shfl_sync(0xffffffffu, nz_lane, j) : 0.f;
                if (n >= a.N || !m_ok) continue;
                float val = v[j] + w[j] + bias;
                if (a.gelu) val = 0.5f * val * (1.0f + erff(val * 0.70710678118654752f));
                val *= gm;
                if (a.epi == E_STORE_F32) { a.x[n * a.ldx + m] = val; continue; }
                if (a.epi == E_CONVT) {
                    const int b = (int)(n / (a.Tin + 1)), qq = (int)(n - (long long)b * (a.Tin + 1));
                    const int to = qq * a.stride + r - a.pad;
                    if (to < 0 || to >= a.T) continue;
                    const long long tok = (long long)b * a.T + to;
                    a.x[tok * a.ldx + co] = val;
                    if (a.hl) put_hilo(a.hl, a.ldh, tok, co, val);
                    continue;
                }
                if (a.epi == E_NOISE) {
                    a.x[n * a.ldx + m] = xv[j] + nz * val;
                    continue;
                }
                if (rmw) {
                    val += xv[j];
                    a.x[n * a.ldx + m] = val;
                    if (a.epi == E_ADD) continue;
                }
                if (a.alpha) val = snake_inv(val, al, inv_al);
                if (a.dual) {
                    const long long b = n / a.T, tt = n - b * a.T;
                    const long long row = b * (a.T + 1) + tt;
                    put_hilo(a.hl, a.ldh, row, m, val);
                    put_hilo(a.hl, a.ldh, row + 1, a.M + m, val);
                } else {
                    put_hilo(a.hl, a.ldh, n, m, val);
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

}  // namespace cg
}  // namespace b2a
