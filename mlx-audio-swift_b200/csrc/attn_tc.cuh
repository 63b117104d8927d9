// Non-causal multi-head attention on tcgen05 for sm_100a (head_dim 64): the Whisper encoder's 1500 x 1500 maps
// (Sources/MLXAudioSTT/Models/Whisper/WhisperLayers.swift:11-73, MLXFast.scaledDotProductAttention there).
//
// Operands: plain fp16, ONE tensor-core product per GEMM (a CPU study of the oracle with this arithmetic emulated,
// profiles/r01_whisper_attention_precision_study.md, keeps the encoder output within 5e-5: scores and probabilities do not need the
// hi/lo pairs the weight GEMMs use); softmax in fp32.
//   Qh [B*nh][Tp][64] fp16, pre-scaled by head_dim^-1/2 (a power of two here: exact)      -> A operand of  S = Q K^T
//   Kh [B*nh][Tp][64] fp16                                                                 -> B operand (K-major: d contiguous)
//   Vt [B*nh][64][Tp] fp16 (transposed so that keys are contiguous)                        -> B operand of  O = P V
//   rows / keys >= T are zero padding up to Tp (a multiple of 128)
// One CTA per (128-query tile, head, clip); 320 threads: warp 0 = TMA producer (2-stage K / V ring), warp 1 = single-thread tcgen05
// issue + TMEM owner, warps 2-9 = softmax (a query row = a TMEM lane is shared by two threads, 64 key columns each).  TMEM: S = 128 fp32 columns, O = 64 (256 allocated, so two
// CTAs share an SM and one's softmax overlaps the other's MMAs).  Two passes over the keys instead of an online softmax: pass 1 only
// takes the row maxima of S (the QK^T product is cheap on this machine), pass 2 recomputes S, writes P = exp(S - max) as the fp16 A
// operand (K-major, 128-byte swizzle, written by hand) and accumulates O += P V in TMEM with no rescaling pass over O.
#pragma once
#include "tc_gemm.cuh"

#include <cuda_fp16.h>

namespace b2a {
namespace fa {

constexpr int BQ = 128, BKV = 128, HDIM = 64;
constexpr int FA_THREADS = 320;                   // producer warp, MMA warp, 8 softmax warps (two per TMEM lane quadrant: 64 key columns each)
constexpr int Q_BYTES = BQ * HDIM * 2, K_BYTES = BKV * HDIM * 2, V_BYTES = HDIM * BKV * 2, P_BYTES = BQ * BKV * 2;
constexpr int STAGE_BYTES = K_BYTES + V_BYTES, FA_STAGES = 2;
constexpr int SMEM_DATA = Q_BYTES + P_BYTES + FA_STAGES * STAGE_BYTES;      // 112 KB
// + barriers + alignment slack = 112.75 KB.  Two CTAs per SM need 2 x (dynamic + 1 KB reserved) <= 228 KB, i.e. <= 113 KB each: the first
// version carried its 1 KB max / sum exchange array next to the barriers (113.75 KB) and ran ONE CTA per SM (ncu: 15.6 % occupancy) --
// the exchange now aliases the head of the P tile, which is idle at both moments it is needed.
constexpr size_t FA_SMEM_BYTES = SMEM_DATA + 256 + 512;

struct Args {
    __nv_bfloat16* out;      // [2 * Tp_tokens, d_model] hi/lo tiles of `half` tokens (the out-projection GEMM's B operand)
    int T, Tp, nh, d_model, half;
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(tc::smem_u32(dst)), "l"(map), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// kind::f16 with fp16 A / B (format 0), fp32 accumulate, both K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_f16(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

__global__ void __launch_bounds__(FA_THREADS, 2)
mha_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, Args a) {
    extern __shared__ __align__(1024) uint8_t fa_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = smem;
    uint8_t* sP = smem + Q_BYTES;
    uint8_t* sKV = sP + P_BYTES;                                   // [stage][K 16 KB | V^T 16 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_DATA);
    uint64_t *qfull = bars, *full = bars + 1, *empty = bars + 3, *sfull = bars + 5, *sfree = bars + 6, *pready = bars + 7, *pfree = bars + 8,
             *ofull = bars + 9;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z, bh = b * a.nh + h;
    const int n_kv = a.Tp / BKV, n_it = 2 * n_kv;
    if (warp == 0 && lane == 0) {
        if (reinterpret_cast<uintptr_t>(smem) - reinterpret_cast<uintptr_t>(fa_raw) > 512) __trap();   // alignment slack exceeded
        tc::tma_prefetch_desc(&tmQ); tc::tma_prefetch_desc(&tmK); tc::tma_prefetch_desc(&tmV);
        tc::mbar_init(qfull, 1);
        for (int i = 0; i < FA_STAGES; ++i) { tc::mbar_init(&full[i], 1); tc::mbar_init(&empty[i], 1); }
        tc::mbar_init(sfull, 1); tc::mbar_init(sfree, 8); tc::mbar_init(pready, 8); tc::mbar_init(pfree, 1); tc::mbar_init(ofull, 1);
        tc::fence_barrier_init();
    }
    if (warp == 1) tc::tmem_alloc<256>(tmem_slot);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem_S = *tmem_slot, tmem_O = tmem_S + 128;

    if (warp == 0) {
        if (lane == 0) {
            tc::mbar_arrive_expect_tx(qfull, Q_BYTES);
            tma_load_3d(sQ, &tmQ, qfull, 0, qt * BQ, bh);
            for (int i = 0; i < n_it; ++i) {
                const int j = i % n_kv, pass = i / n_kv, st = i % FA_STAGES;
                const uint32_t ph = (uint32_t)((i / FA_STAGES) & 1);
                tc::mbar_wait(&empty[st], ph ^ 1);
                uint8_t* sk = sKV + (size_t)st * STAGE_BYTES;
                tc::mbar_arrive_expect_tx(&full[st], pass ? STAGE_BYTES : K_BYTES);
                tma_load_3d(sk, &tmK, &full[st], 0, j * BKV, bh);
                if (pass) {
                    tma_load_3d(sk + K_BYTES, &tmV, &full[st], j * BKV, 0, bh);                 // keys [0, 64) of the tile: [64 d][64 keys]
                    tma_load_3d(sk + K_BYTES + V_BYTES / 2, &tmV, &full[st], j * BKV + 64, 0, bh);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t id_qk = idesc_f16(BKV), id_pv = idesc_f16(HDIM);
            tc::mbar_wait(qfull, 0);
            const uint64_t dq = tc::make_smem_desc(tc::smem_u32(sQ));
            const uint64_t dp0 = tc::make_smem_desc(tc::smem_u32(sP)), dp1 = tc::make_smem_desc(tc::smem_u32(sP + P_BYTES / 2));
            for (int i = 0; i < n_it; ++i) {
                const int j = i % n_kv, pass = i / n_kv, st = i % FA_STAGES;
                tc::mbar_wait(&full[st], (uint32_t)((i / FA_STAGES) & 1));
                if (i > 0) tc::mbar_wait(sfree, (uint32_t)((i - 1) & 1));
                tc::tc_fence_after();
                const uint32_t sk = tc::smem_u32(sKV + (size_t)st * STAGE_BYTES);
                const uint64_t dk = tc::make_smem_desc(sk);
#pragma unroll
                for (int k = 0; k < HDIM / 16; ++k) tc::umma_bf16(tmem_S, dq + (uint64_t)(2 * k), dk + (uint64_t)(2 * k), id_qk, k ? 1u : 0u);
                tc::umma_commit(sfull);
                if (!pass) { tc::umma_commit(&empty[st]); continue; }
                tc::mbar_wait(pready, (uint32_t)(j & 1));
                tc::tc_fence_after();
                const uint64_t dv0 = tc::make_smem_desc(sk + K_BYTES), dv1 = tc::make_smem_desc(sk + K_BYTES + V_BYTES / 2);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_bf16(tmem_O, dp0 + (uint64_t)(2 * k), dv0 + (uint64_t)(2 * k), id_pv, (j == 0 && k == 0) ? 0u : 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) tc::umma_bf16(tmem_O, dp1 + (uint64_t)(2 * k), dv1 + (uint64_t)(2 * k), id_pv, 1u);
                tc::umma_commit(&empty[st]);
                tc::umma_commit(pfree);
                if (j == n_kv - 1) tc::umma_commit(ofull);
            }
        }
    } else {
        const int q = warp & 3, row = q * 32 + lane, ch = (warp - 2) >> 2;      // ch: which 64-column half of every S tile / which half of O
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        float* xch = reinterpret_cast<float*>(sP);                               // [2][128]: the two halves of a row exchange max / sum here (P tile idle)
        float m = -INFINITY, l = 0.f;
        for (int i = 0; i < n_it; ++i) {
            const int j = i % n_kv, pass = i / n_kv;
            if (i == n_kv) {                                                     // pass 1 done: combine the two half-row maxima
                xch[ch * BQ + row] = m;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                m = fmaxf(m, xch[(ch ^ 1) * BQ + row]);
                asm volatile("bar.sync 1, 256;" ::: "memory");                   // nobody writes P over the exchange before everybody has read it
            }
            tc::mbar_wait(sfull, (uint32_t)(i & 1));
            tc::tc_fence_after();
            if (!pass) {
#pragma unroll 1
                for (int c = 2 * ch; c < 2 * ch + 2; ++c) {
                    float v[32];
                    tmem_ld32(tmem_S + lane_off + (uint32_t)(c * 32), v);
                    const int k0 = j * BKV + c * 32;
#pragma unroll
                    for (int e = 0; e < 32; ++e) if (k0 + e < a.T) m = fmaxf(m, v[e]);
                }
                tc::tc_fence_before();
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(sfree);
                continue;
            }
            if (j > 0) tc::mbar_wait(pfree, (uint32_t)((j - 1) & 1));          // the previous P V product has read the P tile
            uint8_t* prow = sP + (size_t)ch * (P_BYTES / 2) + (size_t)row * 128; // this thread's 64 keys = one 128-byte row of panel `ch`
#pragma unroll 1
            for (int c2 = 0; c2 < 2; ++c2) {
                float v[32];
                tmem_ld32(tmem_S + lane_off + (uint32_t)((2 * ch + c2) * 32), v);
                if (c2 == 1) {                                                  // S is in registers: the next Q K^T may overwrite it
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(sfree);
                }
                const int k0 = j * BKV + (2 * ch + c2) * 32;
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 32; e += 2) {
                    const float p0 = (k0 + e < a.T) ? __expf(v[e] - m) : 0.f;
                    const float p1 = (k0 + e + 1 < a.T) ? __expf(v[e + 1] - m) : 0.f;
                    const __half2 hp = __floats2half2_rn(p0, p1);
                    l += __low2float(hp) + __high2float(hp);                   // the sum of what the tensor core will actually multiply
                    pk[e >> 1] = *reinterpret_cast<const uint32_t*>(&hp);
                }
                // K-major SWIZZLE_128B A tile: key kk of row r lives in panel kk / 64 at byte r * 128 + (((kk % 64) / 8) ^ (r & 7)) * 16 + (kk % 8) * 2
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const uint4 val = make_uint4(pk[4 * cc], pk[4 * cc + 1], pk[4 * cc + 2], pk[4 * cc + 3]);
                    *reinterpret_cast<uint4*>(prow + (size_t)(((4 * c2 + cc) ^ (row & 7)) * 16)) = val;
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(pready);
        }
        tc::mbar_wait(ofull, 0);                                                 // every P V product has completed: the P tile is free
        tc::tc_fence_after();
        xch[ch * BQ + row] = l;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        l += xch[(ch ^ 1) * BQ + row];
        const int qrow = qt * BQ + row;
        const float inv = 1.0f / l;
        {
            float v[32];
            tmem_ld32(tmem_O + lane_off + (uint32_t)(ch * 32), v);
            if (qrow < a.T) {
                const long long tok = (long long)b * a.T + qrow;
                const long long r = (tok / a.half) * 2 * a.half + (tok % a.half);
                __nv_bfloat16* ph = a.out + r * a.d_model + h * HDIM + ch * 32;
                __nv_bfloat16* pl = ph + (long long)a.half * a.d_model;
#pragma unroll
                for (int e = 0; e < 32; e += 8) {                               // 16-byte stores: 8 bf16 at a time
                    uint32_t hw[4], lw[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float o0 = v[e + 2 * u] * inv, o1 = v[e + 2 * u + 1] * inv;
                        const __nv_bfloat16 h0 = __float2bfloat16_rn(o0), h1 = __float2bfloat16_rn(o1);
                        const __nv_bfloat162 hh = __halves2bfloat162(h0, h1);
                        const __nv_bfloat162 ll = __halves2bfloat162(__float2bfloat16_rn(o0 - __bfloat162float(h0)), __float2bfloat16_rn(o1 - __bfloat162float(h1)));
                        hw[u] = *reinterpret_cast<const uint32_t*>(&hh);
                        lw[u] = *reinterpret_cast<const uint32_t*>(&ll);
                    }
                    *reinterpret_cast<uint4*>(ph + e) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    *reinterpret_cast<uint4*>(pl + e) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::tc_fence_after();
        tc::tmem_dealloc<256>(tmem_S);
    }
}

// qkv fp32 [B*T, 3 * d_model] (q | k | v) -> Qh (scaled), Kh, Vt.  One CTA per (64-token tile, clip, head); V goes through shared memory
// so that the transposed rows are written 128 bytes at a time.
__global__ void __launch_bounds__(256)
pack_qkv_f16_kernel(const float* __restrict__ qkv, __half* __restrict__ Qh, __half* __restrict__ Kh, __half* __restrict__ Vt, int T, int Tp,
                    int nh, float scale) {
    __shared__ __half sv[64][HDIM + 2];
    const int b = blockIdx.y, h = blockIdx.z, t0 = blockIdx.x * 64, dm = nh * HDIM;
    for (int i = threadIdx.x; i < 64 * HDIM; i += 256) {
        const int r = i >> 6, c = i & 63, t = t0 + r;
        float q = 0.f, k = 0.f, v = 0.f;
        if (t < T) {
            const float* src = qkv + ((long long)b * T + t) * 3 * dm + h * HDIM + c;
            q = src[0] * scale; k = src[dm]; v = src[2 * dm];
        }
        const long long o = (((long long)b * nh + h) * Tp + t) * HDIM + c;
        Qh[o] = __float2half_rn(q);
        Kh[o] = __float2half_rn(k);
        sv[r][c] = __float2half_rn(v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * HDIM; i += 256) {
        const int c = i >> 6, r = i & 63;                       // d = c, token = t0 + r: consecutive threads -> consecutive tokens
        Vt[(((long long)b * nh + h) * HDIM + c) * Tp + t0 + r] = sv[r][c];
    }
}

}  // namespace fa
}  // namespace b2a
