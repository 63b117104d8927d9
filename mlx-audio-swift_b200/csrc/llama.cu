// Orpheus / Llama-3 autoregressive step for sm_100a.  Replaces (reference paths):
//   Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:104-202  Llama3ScaledRoPE
//   Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:206-346  attention / MLP / block / inner model
//   Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:557-567  tied lm head
//   Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:658-765  generate loop (+ :383-434 parseOutput,
//                                                            :41-98 SNAC frame (de)interleave)
//   mlx-swift-lm 3.31.4 (un-vendored): TopPSampler / RepetitionContext (call sites :691-692)
//
// HBM layout: weights bf16 [out, in] row-major (q|k|v fused into one matrix, gate/up row-
// interleaved so SwiGLU is a GEMV epilogue); KV cache bf16 [layer][B][kv_head][ctx][128];
// residual stream fp32 [B, H]; GEMV inputs bf16 [B, K] (staged in shared memory per CTA).
// The whole decode step (embed -> 28 layers -> lm head -> logits processors -> sampler ->
// bookkeeping) is captured in one CUDA graph and replayed per token; nothing syncs with the
// host inside the loop except a poll of the "all rows finished" flag every few steps.
#include "common.cuh"
#include <cooperative_groups.h>
#include "tc_gemm.cuh"
#include "qwen3_sampler.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <cmath>

namespace b2a {

typedef __nv_bfloat16 bf16;

constexpr int TOK_START_OF_HUMAN = 128259, TOK_END_OF_HUMAN = 128260, TOK_END_OF_TEXT = 128009;
constexpr int TOK_START_OF_SPEECH = 128257, TOK_END_OF_SPEECH = 128258, TOK_PAD = 128263;
constexpr int TOK_AUDIO_OFFSET = 128266;

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ float bf_lo(unsigned u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(unsigned u) { return __uint_as_float(u & 0xffff0000u); }

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ------------------------------------------------------------------------------------------------
// embed + RMSNorm
// ------------------------------------------------------------------------------------------------
// x[b,:] = embed[token[b]]  (fp32 residual stream)
// also clears y[b,:] (the fp32 GEMM accumulation target) so a step never sees a previous step's leftovers
__global__ void embed_kernel(const int* __restrict__ tokens, const bf16* __restrict__ embed, float* __restrict__ x,
                             float* __restrict__ y, int H, int V) {
    const int b = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    int tok = tokens[b];
    tok = min(max(tok, 0), V - 1);
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        x[(long long)b * H + i] = __bfloat162float(embed[(long long)tok * H + i]);
        y[(long long)b * H + i] = 0.f;
    }
}

// inputs given as embeddings (row N1): x[b,:] = src[b,:], y cleared like embed_kernel does
__global__ void ext_embed_kernel(const float* __restrict__ src, float* __restrict__ x, float* __restrict__ y, int H) {
    const int b = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        x[(long long)b * H + i] = src[(long long)b * H + i];
        y[(long long)b * H + i] = 0.f;
    }
}

// fused-norm step, row N1 only: the fp32 normalised hidden state  out[b, :] = x[b, :] * rstd[b] * w  from the partial sums of squares
__global__ void finalize_norm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ ss, int parts,
                                     float* __restrict__ out, int H, float eps) {
    const int b = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    float t = 0.f;
    for (int p = 0; p < parts; ++p) t += ss[p * 8 + b];
    const float r = rsqrtf(t / (float)H + eps);
    for (int i = threadIdx.x; i < H; i += blockDim.x) out[(long long)b * H + i] = x[(long long)b * H + i] * r * w[i];
}

constexpr int LO_ROW = 8;   // activation matrices are [16, K] bf16: row b = hi(x_b), row 8 + b = lo(x_b) = bf16(x_b - hi)

// token t of a [tokens, K] activation matrix lives in rows  hi = (t / half) * 2 * half + t % half,  lo = hi + half
// (half = 8 for the decode step's [16, K] matrices, 64 for the prefill's 128-row TMA tiles)
__device__ __forceinline__ void store_hilo(bf16* base, long long ld, int t, long long i, float v, int half = LO_ROW) {
    const bf16 hi = __float2bfloat16_rn(v);
    const long long r = (long long)(t / half) * 2 * half + (t % half);
    base[r * ld + i] = hi;
    base[(r + half) * ld + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// x += delta (optional, delta is zeroed afterwards);  xn = hi/lo split of  x * rsqrt(mean(x^2) + eps) * w
// One 1024-thread CTA per row, the row lives in registers (H <= 8192).  Also zeroes `zero_ptr[b, :zero_n]`
// (the fp32 q|k|v row, so the next stream-K GEMM can accumulate into it with red.add).
constexpr int RN_THREADS = 1024, RN_MAXV = 8;
__global__ void __launch_bounds__(RN_THREADS)
add_rmsnorm_kernel(float* __restrict__ x, float* __restrict__ delta, const float* __restrict__ w,
                   bf16* __restrict__ xn, int H, float eps, float* __restrict__ trace, float* __restrict__ zero_ptr,
                   int zero_n, int half, L2Prefetch pf, float* __restrict__ normed = nullptr, float* __restrict__ ss_out = nullptr,
                   int ss_parts = 0) {
    __shared__ float red[RN_THREADS / 32];
    const int b = blockIdx.x, tid = threadIdx.x;
    pdl_trigger();
    l2_prefetch(pf, b * RN_THREADS + tid, gridDim.x * RN_THREADS);
    pdl_wait();
    float* xr = x + (long long)b * H;
    float v[RN_MAXV];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < RN_MAXV; ++j) {
        const int i = tid + j * RN_THREADS;
        float val = 0.f;
        if (i < H) {
            val = xr[i];
            if (delta) { val += delta[(long long)b * H + i]; xr[i] = val; delta[(long long)b * H + i] = 0.f; }
            if (trace) trace[(long long)b * H + i] = val;
        }
        v[j] = val;
        ss += val * val;
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < RN_THREADS / 32; ++i) tot += red[i];
    // ss_out != null ("raw" mode, fused-norm decode step): xn = hi/lo of x * w UN-normalised, the row's sum of squares goes to
    // ss_out[0, b] (parts 1.. are cleared): the consumer GEMM applies rstd in its epilogue (tc_gemm.cuh, Args::rstd_ss)
    const float r = ss_out ? 1.0f : rsqrtf(tot / (float)H + eps);
    if (ss_out && tid < ss_parts) ss_out[tid * 8 + b] = tid == 0 ? tot : 0.f;
#pragma unroll
    for (int j = 0; j < RN_MAXV; ++j) {
        const int i = tid + j * RN_THREADS;
        if (i < H) {
            const float o = v[j] * r * w[i];
            store_hilo(xn, H, b, i, o, half);
            if (normed) normed[(long long)b * H + i] = o;     // the fp32 normalised row (the talker's hidden state, row N1)
        }
    }
    if (zero_ptr)
        for (int i = tid; i < zero_n; i += RN_THREADS) zero_ptr[(long long)b * zero_n + i] = 0.f;
}

// ------------------------------------------------------------------------------------------------
// Weight-streaming GEMV: y[b, n] = sum_k W[n,k] * x[b,k],  NB rows of x in shared memory (bf16),
// each warp owns ROWS consecutive weight rows and streams them with 16-byte no-allocate loads.
// ------------------------------------------------------------------------------------------------
constexpr int GV_MAX_THREADS = 256;
enum : int { GV_F32 = 0, GV_SWIGLU = 1, GV_F32_ATOMIC = 2 };

// SIMT fallback (B2A_GEMM=simt, or K not a multiple of 64): same numerics as the tcgen05 path -- the
// activation matrix holds hi rows [0, NB) and lo rows [8, 8 + NB); both halves are accumulated and summed.
// grid = (row tiles, K splits); with gridDim.y == 2 the two K halves are combined by atomicAdd into a zeroed y.
template <int NB, int ROWS, int EPI>
__global__ void __launch_bounds__(GV_MAX_THREADS)
gemv_bf16_kernel(const bf16* __restrict__ W, const bf16* __restrict__ xin, float* __restrict__ y,
                 bf16* __restrict__ act, int N, int K) {
    extern __shared__ uint4 sx[];  // [2*NB][Kc/8]
    pdl_trigger();
    pdl_wait();
    const int K8 = K >> 3;
    const int Kc8 = K8 / gridDim.y, kbase = blockIdx.y * Kc8;
    for (int i = threadIdx.x; i < 2 * NB * Kc8; i += blockDim.x) {
        const int r = i / Kc8, k = i - r * Kc8;
        const int grow = r < NB ? r : LO_ROW + (r - NB);
        sx[i] = reinterpret_cast<const uint4*>(xin)[(long long)grow * K8 + kbase + k];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * (blockDim.x >> 5) + warp) * ROWS;
    if (row0 >= N) return;
    const uint4* Wr[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
        Wr[r] = reinterpret_cast<const uint4*>(W) + (long long)min(row0 + r, N - 1) * K8 + kbase;

    float acc[ROWS][2 * NB];
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int b = 0; b < 2 * NB; ++b) acc[r][b] = 0.f;

    for (int k8 = lane; k8 < Kc8; k8 += 32) {
        float wf[ROWS][8];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const uint4 wv = ldg_stream(Wr[r] + k8);
            wf[r][0] = bf_lo(wv.x); wf[r][1] = bf_hi(wv.x); wf[r][2] = bf_lo(wv.y); wf[r][3] = bf_hi(wv.y);
            wf[r][4] = bf_lo(wv.z); wf[r][5] = bf_hi(wv.z); wf[r][6] = bf_lo(wv.w); wf[r][7] = bf_hi(wv.w);
        }
#pragma unroll
        for (int b = 0; b < 2 * NB; ++b) {
            const uint4 xv = sx[b * Kc8 + k8];
            const float xf[8] = {bf_lo(xv.x), bf_hi(xv.x), bf_lo(xv.y), bf_hi(xv.y),
                                 bf_lo(xv.z), bf_hi(xv.z), bf_lo(xv.w), bf_hi(xv.w)};
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[r][b] = fmaf(wf[r][j], xf[j], acc[r][b]);
        }
    }
#pragma unroll
    for (int r = 0; r < ROWS; ++r)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[r][b] = warp_sum(acc[r][b] + acc[r][NB + b]);
    if (lane == 0) {
        if (EPI == GV_F32 || EPI == GV_F32_ATOMIC) {
#pragma unroll
            for (int r = 0; r < ROWS; ++r)
                if (row0 + r < N)
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        if (EPI == GV_F32) y[(long long)b * N + row0 + r] = acc[r][b];
                        else atomicAdd(&y[(long long)b * N + row0 + r], acc[r][b]);
                    }
        } else {  // rows are (gate, up) pairs: act[b, n/2] = silu(gate) * up   (LlamaTTS.swift:282-284)
#pragma unroll
            for (int r = 0; r < ROWS; r += 2)
                if (row0 + r + 1 < N)
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const float g = acc[r][b], u = acc[r + 1][b];
                        store_hilo(act, N / 2, b, (row0 + r) / 2, g / (1.0f + __expf(-g)) * u);
                    }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Decode attention (LlamaTTS.swift:235-266): RoPE on q/k, append k/v to the fp32 cache, softmax(qK^T)V.
// grid = (kv heads, rows, key splits): split s owns keys [s*AT_CAP, (s+1)*AT_CAP); its K and V rows are
// one contiguous block of the cache each, fetched with a single cp.async.bulk per matrix into shared memory
// (one HBM round trip), the partial (max, sum, out) goes to a workspace and the last CTA of a (row, head)
// to finish merges the partials (flash-decoding).
// ------------------------------------------------------------------------------------------------
constexpr int HD = 128, AT_THREADS = 256, MAXG = 8, AT_CAP = 64;   // AT_CAP * 4 == AT_THREADS

#ifdef B2A_ATTN_TIMING
__device__ long long g_attn_ts[8 * 4096];
#define ATS(i) do { if (threadIdx.x == 0) g_attn_ts[(((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 8 + (i)] = clock64(); } while (0)
#else
#define ATS(i) do {} while (0)
#endif

struct AttnArgs {
    const float* qkv;      // [B, (nq + 2 nkv) * 128] fp32
    const int* pos;        // [B]
    const float* freqs;    // [64] llama3 rope divisors
    float* kcache;         // this layer: [B][nkv][max_ctx][128] fp32
    float* vcache;
    bf16* out;             // [16, nq*128] hi/lo
    float* part_o;         // [B][nkv][S][G][128]
    float* part_ml;        // [B][nkv][S][G][2]
    int* counters;         // [B][nkv], zero between launches
    int nq, nkv, max_ctx, S;
    float scale;
    L2Prefetch pf;         // weights of a later GEMM, prefetched into L2 while attention (which reads little) runs
    const float* qnorm;    // nullable [128]: per-head RMSNorm gain applied to every q head BEFORE RoPE (Qwen3TTSTalker.swift:127-186)
    const float* knorm;    // nullable [128]: same for the k head  (cluster kernel only)
    float qk_eps;
    int zero_qkv;          // fused-norm step: clear this (row, kv head)'s q | k | v slices after reading them, so that the next layer's
                           // stream-K QKV GEMM can red.add into the row (the stand-alone norm kernel that used to do it is gone)
};

template <int G>
__global__ void __launch_bounds__(AT_THREADS)
attn_decode_kernel(AttnArgs a) {
    extern __shared__ __align__(16) uint8_t at_smem[];
    float* sK = reinterpret_cast<float*>(at_smem);              // [AT_CAP][128]
    float* sV = sK + AT_CAP * HD;                               // [AT_CAP][128]
    float* sq = sV + AT_CAP * HD;                               // [G][128]
    float* sc = sq + G * HD;                                    // [G][AT_CAP]
    float* wpo = sc + G * AT_CAP;                               // [8 warps][G][128] warp-partial outputs
    uint64_t* bar = reinterpret_cast<uint64_t*>(wpo + (AT_THREADS / 32) * G * HD);
    __shared__ float red[AT_THREADS / 32][MAXG];
    __shared__ int s_last;

    const int h = blockIdx.x, b = blockIdx.y, s = blockIdx.z, tid = threadIdx.x;
    ATS(0);
    pdl_trigger();
    l2_prefetch(a.pf, ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * AT_THREADS + tid,
                gridDim.x * gridDim.y * gridDim.z * AT_THREADS);
    // Everything that does not depend on this step's q|k|v runs BEFORE griddepcontrol.wait and overlaps with the
    // tail of the QKV GEMM: pos[] is only written by the sampler (last kernel of the previous step's graph), the
    // cache rows < pos by earlier steps.  So: position, RoPE angles, mbarrier, and the bulk K/V loads go first.
    const int p = a.pos[b];
    const bool row_ok = p >= 0 && p < a.max_ctx;
    const int S_eff = row_ok ? p / AT_CAP + 1 : 0;
    if (s >= S_eff) { pdl_wait(); return; }
    const int t0 = s * AT_CAP, t1 = min(t0 + AT_CAP, p + 1), nk = t1 - t0;
    const bool has_new = (s == S_eff - 1);                      // this split owns the new position p
    const int n_load = has_new ? nk - 1 : nk;
    const int qkv_ld = (a.nq + 2 * a.nkv) * HD;
    const float* row = a.qkv + (long long)b * qkv_ld;
    float* kc = a.kcache + (((long long)b * a.nkv + h) * a.max_ctx) * HD;
    float* vc = a.vcache + (((long long)b * a.nkv + h) * a.max_ctx) * HD;

    if (tid == 0) {
        tc::mbar_init(bar, 1);
        tc::fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (n_load > 0) {
            const uint32_t bytes = (uint32_t)n_load * HD * 4;
            tc::mbar_arrive_expect_tx(bar, 2 * bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tc::smem_u32(sK)), "l"(kc + (long long)t0 * HD), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tc::smem_u32(sV)), "l"(vc + (long long)t0 * HD), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
        } else {
            tc::mbar_arrive(bar);
        }
    }
    float sn = 0.f, cs = 1.f;
    if (tid < HD / 2) sincosf((float)p / a.freqs[tid], &sn, &cs);   // MLXFast.RoPE(freqs:): angle = pos / freqs[i]
    pdl_wait();
    ATS(1);
    if (tid < HD / 2) {  // non-traditional RoPE: pairs (i, i+64)
        const int d = tid;
        _Pragma("unroll") for (int g = 0; g < G; ++g) {
            const float* q = row + (h * G + g) * HD;
            const float x1 = q[d], x2 = q[d + HD / 2];
            sq[g * HD + d] = x1 * cs - x2 * sn;
            sq[g * HD + d + HD / 2] = x2 * cs + x1 * sn;
        }
        if (has_new) {
            const float* k = row + (a.nq + h) * HD;
            const float x1 = k[d], x2 = k[d + HD / 2];
            const float k1 = x1 * cs - x2 * sn, k2 = x2 * cs + x1 * sn;
            kc[(long long)p * HD + d] = k1;
            kc[(long long)p * HD + d + HD / 2] = k2;
            sK[(p - t0) * HD + d] = k1;
            sK[(p - t0) * HD + d + HD / 2] = k2;
        }
    } else if (tid >= 128 && has_new) {
        const int d = tid - 128;
        const float v = row[(a.nq + a.nkv + h) * HD + d];
        vc[(long long)p * HD + d] = v;
        sV[(p - t0) * HD + d] = v;
    }
    __syncthreads();            // barrier init + q / new-row staging visible
    ATS(2);
    tc::mbar_wait(bar, 0);      // bulk-copied K and V have landed
    ATS(3);

    // Each warp owns 8 keys end to end (4 lanes per key): scores, a warp-local softmax (max / sum by shuffles) and
    // its partial P*V; the 8 warp partials are merged through shared memory with ONE block barrier.
    const int lane = tid & 31, warp = tid >> 5;
    const int key = warp * 8 + (lane >> 2), part = lane & 3;
    float sacc[G];
    _Pragma("unroll") for (int g = 0; g < G; ++g) sacc[g] = 0.f;
    if (key < nk) {
        const float4* kr = reinterpret_cast<const float4*>(sK + key * HD);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int d4 = part + 4 * ((j + key) & 7);     // rotated columns: every quarter-warp hits 8 distinct bank groups
            const float4 kf = kr[d4];
            _Pragma("unroll") for (int g = 0; g < G; ++g) {
                const float4 qf = reinterpret_cast<const float4*>(sq + g * HD)[d4];
                sacc[g] = fmaf(qf.x, kf.x, sacc[g]); sacc[g] = fmaf(qf.y, kf.y, sacc[g]);
                sacc[g] = fmaf(qf.z, kf.z, sacc[g]); sacc[g] = fmaf(qf.w, kf.w, sacc[g]);
            }
        }
    }
    float pw[G], mw[G], lw[G];
    _Pragma("unroll") for (int g = 0; g < G; ++g) {
        float v = sacc[g];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        const float sv = key < nk ? v * a.scale : -INFINITY;
        float m = sv;
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
        const float e = (m == -INFINITY) ? 0.f : __expf(sv - m);    // all 4 lanes of a key hold the same value
        float l = part == 0 ? e : 0.f;
        l = warp_sum(l);
        pw[g] = e; mw[g] = m; lw[g] = l;
    }
    // warp-partial P*V: lane owns dims 4*lane .. 4*lane+3 (conflict-free float4 reads of a V row)
    float4 o4[G];
    _Pragma("unroll") for (int g = 0; g < G; ++g) o4[g] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const int t = warp * 8 + kk;
        if (t < nk) {
            const float4 v = reinterpret_cast<const float4*>(sV + t * HD)[lane];
            _Pragma("unroll") for (int g = 0; g < G; ++g) {
                const float pk = __shfl_sync(0xffffffffu, pw[g], kk * 4);
                o4[g].x = fmaf(pk, v.x, o4[g].x); o4[g].y = fmaf(pk, v.y, o4[g].y);
                o4[g].z = fmaf(pk, v.z, o4[g].z); o4[g].w = fmaf(pk, v.w, o4[g].w);
            }
        }
    }
    // spo doubles as the warp-partial buffer: [8 warps][G][128]; red / sc hold the warp (max, sum)
    _Pragma("unroll") for (int g = 0; g < G; ++g) {
        reinterpret_cast<float4*>(wpo + (warp * G + g) * HD)[lane] = o4[g];
        if (lane == 0) { red[warp][g] = mw[g]; sc[g * AT_CAP + warp] = lw[g]; }
    }
    __syncthreads();
    const long long pbase = (((long long)b * a.nkv + h) * a.S + s) * G;
    if (tid < HD) {
        _Pragma("unroll") for (int g = 0; g < G; ++g) {
            float M = red[0][g];
#pragma unroll
            for (int w = 1; w < AT_THREADS / 32; ++w) M = fmaxf(M, red[w][g]);
            float L = 0.f, O = 0.f;
#pragma unroll
            for (int w = 0; w < AT_THREADS / 32; ++w) {
                const float sc_w = red[w][g] == -INFINITY ? 0.f : __expf(red[w][g] - M);
                L = fmaf(sc[g * AT_CAP + w], sc_w, L);
                O = fmaf(wpo[(w * G + g) * HD + tid], sc_w, O);
            }
            a.part_o[(pbase + g) * HD + tid] = O;
            if (tid == 0) { a.part_ml[(pbase + g) * 2] = M; a.part_ml[(pbase + g) * 2 + 1] = L; }
        }
    }
    ATS(4);
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(&a.counters[b * a.nkv + h], 1) == S_eff - 1);
    __syncthreads();
    ATS(5);
    if (!s_last) return;
    __threadfence();
    // merge the S_eff partials
    if (tid < HD) {
        const long long mbase = (((long long)b * a.nkv + h) * a.S) * G;
        _Pragma("unroll") for (int g = 0; g < G; ++g) {
            float M = -INFINITY;
            for (int j = 0; j < S_eff; ++j) M = fmaxf(M, a.part_ml[(mbase + (long long)j * G + g) * 2]);
            float L = 0.f, O = 0.f;
            for (int j = 0; j < S_eff; ++j) {
                const float wj = __expf(a.part_ml[(mbase + (long long)j * G + g) * 2] - M);
                L = fmaf(a.part_ml[(mbase + (long long)j * G + g) * 2 + 1], wj, L);
                O = fmaf(a.part_o[(mbase + (long long)j * G + g) * HD + tid], wj, O);
            }
            store_hilo(a.out, (long long)a.nq * HD, b, (h * G + g) * HD + tid, O / L);
        }
    }
    if (tid == 0) a.counters[b * a.nkv + h] = 0;
    ATS(6);
}

// ------------------------------------------------------------------------------------------------
// Decode attention, one CTA per (kv head, row), looping over 64-key chunks with an online softmax (default).
// The split-K kernel above pays three dependent global round trips after its compute (partials -> fence -> atomic ->
// partial reads by the last CTA): measured 16.9 us per layer at context 320 for 21 MB of K/V.  Here nothing leaves the SM:
// chunks stream through an NB-deep ring of shared-memory buffers (one cp.async.bulk per matrix per chunk, up to NB chunks in
// flight, the first NB issued BEFORE griddepcontrol.wait so they overlap the tail of the QKV GEMM), every warp keeps a
// running (max, sum, P*V) for its 8 keys of each chunk, and the 8 warp states are merged once at the end.
// ------------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(AT_THREADS)
attn_decode_loop_kernel(AttnArgs a, int NB) {
    extern __shared__ __align__(16) uint8_t at_smem[];
    float* sKV = reinterpret_cast<float*>(at_smem);             // [NB][2][AT_CAP][128]  (K then V of each ring slot)
    float* sq = sKV + (size_t)NB * 2 * AT_CAP * HD;             // [G][128]
    float* snew = sq + G * HD;                                  // [2][128] the new k / v row of this step
    float* wpo = snew + 2 * HD;                                 // [8 warps][G][128] warp-partial outputs
    uint64_t* bars = reinterpret_cast<uint64_t*>(wpo + (AT_THREADS / 32) * G * HD);   // [NB]
    __shared__ float red_m[AT_THREADS / 32][MAXG], red_l[AT_THREADS / 32][MAXG];

    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    l2_prefetch(a.pf, (blockIdx.y * gridDim.x + blockIdx.x) * AT_THREADS + tid, gridDim.x * gridDim.y * AT_THREADS);
    const int p = a.pos[b];                                     // written by the previous step's sampler only
    if (!(p >= 0 && p < a.max_ctx)) { pdl_wait(); return; }
    const int nch = p / AT_CAP + 1;
    const int qkv_ld = (a.nq + 2 * a.nkv) * HD;
    const float* row = a.qkv + (long long)b * qkv_ld;
    float* kc = a.kcache + (((long long)b * a.nkv + h) * a.max_ctx) * HD;
    float* vc = a.vcache + (((long long)b * a.nkv + h) * a.max_ctx) * HD;
    // rows of chunk c already in the cache (the new position p is staged from this step's q|k|v instead)
    auto issue = [&](int c) {
        const int slot = c % NB;
        const int n_load = (c < nch - 1) ? AT_CAP : p - c * AT_CAP;
        float* dK = sKV + (size_t)slot * 2 * AT_CAP * HD;
        if (n_load > 0) {
            const uint32_t bytes = (uint32_t)n_load * HD * 4;
            tc::mbar_arrive_expect_tx(&bars[slot], 2 * bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tc::smem_u32(dK)), "l"(kc + (long long)c * AT_CAP * HD), "r"(bytes), "r"(tc::smem_u32(&bars[slot])) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tc::smem_u32(dK + AT_CAP * HD)), "l"(vc + (long long)c * AT_CAP * HD), "r"(bytes), "r"(tc::smem_u32(&bars[slot])) : "memory");
        } else {
            tc::mbar_arrive(&bars[slot]);
        }
    };
    if (tid == 0) {
        for (int i = 0; i < NB; ++i) tc::mbar_init(&bars[i], 1);
        tc::fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int c = 0; c < min(nch, NB); ++c) issue(c);
    }
    float sn = 0.f, cs = 1.f;
    if (tid < HD / 2) sincosf((float)p / a.freqs[tid], &sn, &cs);   // MLXFast.RoPE(freqs:): angle = pos / freqs[i]
    pdl_wait();
    if (tid < HD / 2) {  // non-traditional RoPE: pairs (i, i+64)
        const int d = tid;
        _Pragma("unroll") for (int g = 0; g < G; ++g) {
            const float* q = row + (h * G + g) * HD;
            const float x1 = q[d], x2 = q[d + HD / 2];
            sq[g * HD + d] = x1 * cs - x2 * sn;
            sq[g * HD + d + HD / 2] = x2 * cs + x1 * sn;
        }
        const float* k = row + (a.nq + h) * HD;
        const float x1 = k[d], x2 = k[d + HD / 2];
        const float k1 = x1 * cs - x2 * sn, k2 = x2 * cs + x1 * sn;
        kc[(long long)p * HD + d] = k1;
        kc[(long long)p * HD + d + HD / 2] = k2;
        snew[d] = k1;
        snew[d + HD / 2] = k2;
    } else if (tid >= 128) {
        const int d = tid - 128;
        const float v = row[(a.nq + a.nkv + h) * HD + d];
        vc[(long long)p * HD + d] = v;
        snew[HD + d] = v;
    }
    __syncthreads();            // barrier init + q / new-row staging visible

    const int lane = tid & 31, warp = tid >> 5;
    const int kslot = warp * 8 + (lane >> 2), part = lane & 3;
    float m_run[G], l_run[G];
    float4 o4[G];
    _Pragma("unroll") for (int g = 0; g < G; ++g) { m_run[g] = -INFINITY; l_run[g] = 0.f; o4[g] = make_float4(0.f, 0.f, 0.f, 0.f); }

    for (int c = 0; c < nch; ++c) {
        const int slot = c % NB;
        float* sK = sKV + (size_t)slot * 2 * AT_CAP * HD;
        float* sV = sK + AT_CAP * HD;
        const int t0 = c * AT_CAP, nk = min(AT_CAP, p + 1 - t0);
        tc::mbar_wait(&bars[slot], (uint32_t)((c / NB) & 1));    // bulk-copied K and V of this chunk have landed
        if (c == nch - 1) {                                      // splice in the new position (the bulk copy stopped before it)
            if (tid < HD) sK[(p - t0) * HD + tid] = snew[tid];
            else sV[(p - t0) * HD + tid - HD] = snew[tid];
            __syncthreads();
        }
        float sacc[G];
        _Pragma("unroll") for (int g = 0; g < G; ++g) sacc[g] = 0.f;
        if (kslot < nk) {
            const float4* kr = reinterpret_cast<const float4*>(sK + kslot * HD);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d4 = part + 4 * ((j + kslot) & 7);     // rotated columns: every quarter-warp hits 8 distinct bank groups
                const float4 kf = kr[d4];
                _Pragma("unroll") for (int g = 0; g < G; ++g) {
                    const float4 qf = reinterpret_cast<const float4*>(sq + g * HD)[d4];
                    sacc[g] = fmaf(qf.x, kf.x, sacc[g]); sacc[g] = fmaf(qf.y, kf.y, sacc[g]);
                    sacc[g] = fmaf(qf.z, kf.z, sacc[g]); sacc[g] = fmaf(qf.w, kf.w, sacc[g]);
                }
            }
        }
        float pw[G];
        _Pragma("unroll") for (int g = 0; g < G; ++g) {
            float v = sacc[g];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            const float sv = kslot < nk ? v * a.scale : -INFINITY;
            float m = sv;
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
            const float m_new = fmaxf(m_run[g], m);
            const float rescale = (m_run[g] == -INFINITY) ? 0.f : __expf(m_run[g] - m_new);
            const float e = (sv == -INFINITY) ? 0.f : __expf(sv - m_new);     // all 4 lanes of a key hold the same value
            float l = part == 0 ? e : 0.f;
            l = warp_sum(l);
            l_run[g] = l_run[g] * rescale + l;
            m_run[g] = m_new;
            o4[g].x *= rescale; o4[g].y *= rescale; o4[g].z *= rescale; o4[g].w *= rescale;
            pw[g] = e;
        }
        // warp-partial P*V: lane owns dims 4*lane .. 4*lane+3 (conflict-free float4 reads of a V row)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int t = warp * 8 + kk;
            if (t < nk) {
                const float4 v = reinterpret_cast<const float4*>(sV + t * HD)[lane];
                _Pragma("unroll") for (int g = 0; g < G; ++g) {
                    const float pk = __shfl_sync(0xffffffffu, pw[g], kk * 4);
                    o4[g].x = fmaf(pk, v.x, o4[g].x); o4[g].y = fmaf(pk, v.y, o4[g].y);
                    o4[g].z = fmaf(pk, v.z, o4[g].z); o4[g].w = fmaf(pk, v.w, o4[g].w);
                }
            }
        }
        if (c + NB < nch) {
            __syncthreads();                                     // every warp is done with this slot
            if (tid == 0) issue(c + NB);
        }
    }
    // merge the 8 warp states
    _Pragma("unroll") for (int g = 0; g < G; ++g) {
        reinterpret_cast<float4*>(wpo + (warp * G + g) * HD)[lane] = o4[g];
        if (lane == 0) { red_m[warp][g] = m_run[g]; red_l[warp][g] = l_run[g]; }
    }
    __syncthreads();
    if (tid < HD) {
        _Pragma("unroll") for (int g = 0; g < G; ++g) {
            float M = red_m[0][g];
#pragma unroll
            for (int w = 1; w < AT_THREADS / 32; ++w) M = fmaxf(M, red_m[w][g]);
            float L = 0.f, O = 0.f;
#pragma unroll
            for (int w = 0; w < AT_THREADS / 32; ++w) {
                const float sc_w = red_m[w][g] == -INFINITY ? 0.f : __expf(red_m[w][g] - M);
                L = fmaf(red_l[w][g], sc_w, L);
                O = fmaf(wpo[(w * G + g) * HD + tid], sc_w, O);
            }
            store_hilo(a.out, (long long)a.nq * HD, b, (h * G + g) * HD + tid, O / L);
        }
    }
}

// Same, with the chunks of one (kv head, row) dealt alternately to the TWO CTAs of a thread-block cluster: 128 CTAs instead of
// 64, so up to 2 x NB chunks (6 x 64 keys) are in flight before griddepcontrol.wait and the serial chunk count halves.  CTA 1
// hands its (max, sum, P*V) state to CTA 0 through distributed shared memory; one cluster barrier, nothing goes through HBM.
template <int G>
__global__ void __cluster_dims__(1, 1, 2) __launch_bounds__(AT_THREADS)
attn_decode_cluster_kernel(AttnArgs a, int NB) {
    namespace cgr = cooperative_groups;
    cgr::cluster_group cluster = cgr::this_cluster();
    const int rank = (int)cluster.block_rank();
    extern __shared__ __align__(16) uint8_t at_smem[];
    float* sKV = reinterpret_cast<float*>(at_smem);             // [NB][2][AT_CAP][128]  (K then V of each ring slot)
    float* sq = sKV + (size_t)NB * 2 * AT_CAP * HD;             // [G][128]
    float* snew = sq + G * HD;                                  // [2][128] the new k / v row of this step
    float* wpo = snew + 2 * HD;                                 // [8 warps][G][128] warp-partial outputs
    float* xo = wpo + (AT_THREADS / 32) * G * HD;               // [G][128] + [2][G]: the peer CTA's merged state (written remotely)
    float* xml = xo + G * HD;
    uint64_t* bars = reinterpret_cast<uint64_t*>(xml + 2 * MAXG);   // [NB]
    __shared__ float red_m[AT_THREADS / 32][MAXG], red_l[AT_THREADS / 32][MAXG];

    const int h = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    l2_prefetch(a.pf, ((blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * AT_THREADS + tid, gridDim.x * gridDim.y * gridDim.z * AT_THREADS);
    // pos[b] is read BEFORE griddepcontrol.wait so that the cached K / V chunks stream in under the QKV GEMM's tail.  Programmatic
    // launches chain (a kernel triggers its dependents at its first instruction, even while it is itself still waiting), so on a
    // small model -- every kernel of several layers resident at once -- this prologue can run before a kernel launched many
    // launches earlier has finished (measured: 1e-1 logits error on a 256-wide model).  The contract that makes the early read
    // safe: EVERY kernel that writes pos[] never calls pdl_trigger() (sample_kernel, prefill_advance_kernel, the q3_* kernels that
    // set a position), so nothing launched after it starts before it has completed.
    const int p = a.pos[b];
    if (!(p >= 0 && p < a.max_ctx)) { pdl_wait(); return; }
    const int nch = p / AT_CAP + 1;
    const int n_my = nch > rank ? (nch - rank + 1) / 2 : 0;     // this CTA's chunks: rank, rank + 2, ...
    const bool owner = ((nch - 1) & 1) == rank;                 // the CTA whose last chunk holds the new position
    const int qkv_ld = (a.nq + 2 * a.nkv) * HD;
    const float* row = a.qkv + (long long)b * qkv_ld;
    float* kc = a.kcache + (((long long)b * a.nkv + h) * a.max_ctx) * HD;
    float* vc = a.vcache + (((long long)b * a.nkv + h) * a.max_ctx) * HD;
    // rows of chunk c already in the cache (the new position p is staged from this step's q|k|v instead)
    auto issue = [&](int i) {
        const int slot = i % NB, c = rank + 2 * i;
        const int n_load = (c < nch - 1) ? AT_CAP : p - c * AT_CAP;
        float* dK = sKV + (size_t)slot * 2 * AT_CAP * HD;
        if (n_load > 0) {
            const uint32_t bytes = (uint32_t)n_load * HD * 4;
            tc::mbar_arrive_expect_tx(&bars[slot], 2 * bytes);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tc::smem_u32(dK)), "l"(kc + (long long)c * AT_CAP * HD), "r"(bytes), "r"(tc::smem_u32(&bars[slot])) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(tc::smem_u32(dK + AT_CAP * HD)), "l"(vc + (long long)c * AT_CAP * HD), "r"(bytes), "r"(tc::smem_u32(&bars[slot])) : "memory");
        } else {
            tc::mbar_arrive(&bars[slot]);
        }
    };
    if (tid == 0) {
        for (int i = 0; i < NB; ++i) tc::mbar_init(&bars[i], 1);
        tc::fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int i = 0; i < min(n_my, NB); ++i) issue(i);
    }
    float sn = 0.f, cs = 1.f;
    if (tid < HD / 2) sincosf((float)p / a.freqs[tid], &sn, &cs);   // MLXFast.RoPE(freqs:): angle = pos / freqs[i]
    pdl_wait();
    const float* qsrc = row + (long long)h * G * HD;            // this kv head's G query heads, contiguous
    const float* ksrc = row + (a.nq + h) * HD;
    if (a.qnorm) {
        // per-head RMSNorm of q and k before RoPE: x * rsqrt(mean(x^2) + eps) * w, one warp per 128-vector, staged in the
        // (not yet used) warp-partial output area
        for (int vec = tid >> 5; vec < G + 1; vec += AT_THREADS / 32) {
            const float* src = vec < G ? qsrc + vec * HD : ksrc;
            const float* w = vec < G ? a.qnorm : a.knorm;
            const float4 x = reinterpret_cast<const float4*>(src)[tid & 31];
            const float ss = warp_sum(x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w);
            const float r = rsqrtf(ss * (1.0f / HD) + a.qk_eps);
            const float4 g4 = reinterpret_cast<const float4*>(w)[tid & 31];
            reinterpret_cast<float4*>(wpo + vec * HD)[tid & 31] = make_float4(x.x * r * g4.x, x.y * r * g4.y, x.z * r * g4.z, x.w * r * g4.w);
        }
        __syncthreads();
        qsrc = wpo;
        ksrc = wpo + G * HD;
    }
    if (tid < HD / 2) {  // non-traditional RoPE: pairs (i, i+64)
        const int d = tid;
        _Pragma("unroll") for (int g = 0; g < G; ++g) {
            const float* q = qsrc + g * HD;
            const float x1 = q[d], x2 = q[d + HD / 2];
            sq[g * HD + d] = x1 * cs - x2 * sn;
            sq[g * HD + d + HD / 2] = x2 * cs + x1 * sn;
        }
        const float* k = ksrc;
        const float x1 = k[d], x2 = k[d + HD / 2];
        const float k1 = x1 * cs - x2 * sn, k2 = x2 * cs + x1 * sn;
        if (owner) { kc[(long long)p * HD + d] = k1; kc[(long long)p * HD + d + HD / 2] = k2; }
        snew[d] = k1;
        snew[d + HD / 2] = k2;
    } else if (tid >= 128) {
        const int d = tid - 128;
        const float v = row[(a.nq + a.nkv + h) * HD + d];
        if (owner) vc[(long long)p * HD + d] = v;
        snew[HD + d] = v;
    }
    __syncthreads();            // barrier init + q / new-row staging visible

    const int lane = tid & 31, warp = tid >> 5;
    const int kslot = warp * 8 + (lane >> 2), part = lane & 3;
    float m_run[G], l_run[G];
    float4 o4[G];
    _Pragma("unroll") for (int g = 0; g < G; ++g) { m_run[g] = -INFINITY; l_run[g] = 0.f; o4[g] = make_float4(0.f, 0.f, 0.f, 0.f); }

    for (int i = 0; i < n_my; ++i) {
        const int slot = i % NB, c = rank + 2 * i;
        float* sK = sKV + (size_t)slot * 2 * AT_CAP * HD;
        float* sV = sK + AT_CAP * HD;
        const int t0 = c * AT_CAP, nk = min(AT_CAP, p + 1 - t0);
        tc::mbar_wait(&bars[slot], (uint32_t)((i / NB) & 1));    // bulk-copied K and V of this chunk have landed
        if (c == nch - 1) {                                      // splice in the new position (the bulk copy stopped before it)
            if (tid < HD) sK[(p - t0) * HD + tid] = snew[tid];
            else sV[(p - t0) * HD + tid - HD] = snew[tid];
            __syncthreads();
        }
        float sacc[G];
        _Pragma("unroll") for (int g = 0; g < G; ++g) sacc[g] = 0.f;
        if (kslot < nk) {
            const float4* kr = reinterpret_cast<const float4*>(sK + kslot * HD);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int d4 = part + 4 * ((j + kslot) & 7);     // rotated columns: every quarter-warp hits 8 distinct bank groups
                const float4 kf = kr[d4];
                _Pragma("unroll") for (int g = 0; g < G; ++g) {
                    const float4 qf = reinterpret_cast<const float4*>(sq + g * HD)[d4];
                    sacc[g] = fmaf(qf.x, kf.x, sacc[g]); sacc[g] = fmaf(qf.y, kf.y, sacc[g]);
                    sacc[g] = fmaf(qf.z, kf.z, sacc[g]); sacc[g] = fmaf(qf.w, kf.w, sacc[g]);
                }
            }
        }
        float pw[G];
        _Pragma("unroll") for (int g = 0; g < G; ++g) {
            float v = sacc[g];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            const float sv = kslot < nk ? v * a.scale : -INFINITY;
            float m = sv;
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 8));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
            const float m_new = fmaxf(m_run[g], m);
            const float rescale = (m_run[g] == -INFINITY) ? 0.f : __expf(m_run[g] - m_new);
            const float e = (sv == -INFINITY) ? 0.f : __expf(sv - m_new);     // all 4 lanes of a key hold the same value
            float l = part == 0 ? e : 0.f;
            l = warp_sum(l);
            l_run[g] = l_run[g] * rescale + l;
            m_run[g] = m_new;
            o4[g].x *= rescale; o4[g].y *= rescale; o4[g].z *= rescale; o4[g].w *= rescale;
            pw[g] = e;
        }
        // warp-partial P*V: lane owns dims 4*lane .. 4*lane+3 (conflict-free float4 reads of a V row)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int t = warp * 8 + kk;
            if (t < nk) {
                const float4 v = reinterpret_cast<const float4*>(sV + t * HD)[lane];
                _Pragma("unroll") for (int g = 0; g < G; ++g) {
                    const float pk = __shfl_sync(0xffffffffu, pw[g], kk * 4);
                    o4[g].x = fmaf(pk, v.x, o4[g].x); o4[g].y = fmaf(pk, v.y, o4[g].y);
                    o4[g].z = fmaf(pk, v.z, o4[g].z); o4[g].w = fmaf(pk, v.w, o4[g].w);
                }
            }
        }
        if (i + NB < n_my) {
            __syncthreads();                                     // every warp is done with this slot
            if (tid == 0) issue(i + NB);
        }
    }
    // merge the 8 warp states
    _Pragma("unroll") for (int g = 0; g < G; ++g) {
        reinterpret_cast<float4*>(wpo + (warp * G + g) * HD)[lane] = o4[g];
        if (lane == 0) { red_m[warp][g] = m_run[g]; red_l[warp][g] = l_run[g]; }
    }
    __syncthreads();
    float Ms[G], Ls[G], Os[G];
    if (tid < HD) {
        _Pragma("unroll") for (int g = 0; g < G; ++g) {
            float M = red_m[0][g];
#pragma unroll
            for (int w = 1; w < AT_THREADS / 32; ++w) M = fmaxf(M, red_m[w][g]);
            float L = 0.f, O = 0.f;
#pragma unroll
            for (int w = 0; w < AT_THREADS / 32; ++w) {
                const float sc_w = red_m[w][g] == -INFINITY ? 0.f : __expf(red_m[w][g] - M);
                L = fmaf(red_l[w][g], sc_w, L);
                O = fmaf(wpo[(w * G + g) * HD + tid], sc_w, O);
            }
            Ms[g] = M; Ls[g] = L; Os[g] = O;
        }
        if (rank == 1) {                                         // hand the state to CTA 0 (distributed shared memory)
            float* rxo = cluster.map_shared_rank(xo, 0);
            float* rxml = cluster.map_shared_rank(xml, 0);
            _Pragma("unroll") for (int g = 0; g < G; ++g) {
                rxo[g * HD + tid] = Os[g];
                if (tid == 0) { rxml[g] = Ms[g]; rxml[MAXG + g] = Ls[g]; }
            }
        }
    }
    cluster.sync();
    if (a.zero_qkv && rank == 0) {
        float* wrow = const_cast<float*>(row);
        for (int i = tid; i < G * HD; i += AT_THREADS) wrow[(long long)h * G * HD + i] = 0.f;
        if (tid < HD) { wrow[(a.nq + h) * HD + tid] = 0.f; wrow[(a.nq + a.nkv + h) * HD + tid] = 0.f; }
    }
    if (rank == 0 && tid < HD) {
        _Pragma("unroll") for (int g = 0; g < G; ++g) {
            const float M1 = xml[g], L1 = xml[MAXG + g], O1 = xo[g * HD + tid];
            const float M = fmaxf(Ms[g], M1);
            const float w0 = Ms[g] == -INFINITY ? 0.f : __expf(Ms[g] - M), w1 = M1 == -INFINITY ? 0.f : __expf(M1 - M);
            const float L = Ls[g] * w0 + L1 * w1, O = Os[g] * w0 + O1 * w1;
            store_hilo(a.out, (long long)a.nq * HD, b, (h * G + g) * HD + tid, O / L);
        }
    }
}

#ifdef B2A_ATTN_TIMING
extern "C" int b2a_debug_attn_ts(long long* out, int n) {
    return (int)cudaMemcpyFromSymbol(out, g_attn_ts, sizeof(long long) * n);
}
#endif

// ------------------------------------------------------------------------------------------------
// Batched prefill (LlamaTTS.swift:711 `self(inputIds, cache)` on the whole prompt): all B*L prompt tokens go
// through each layer at once -- GEMMs on the tcgen05 kernel with 128-column tiles (64 tokens as hi/lo pairs),
// causal attention over the prompt per (row, kv head, 32-query tile), K/V written to the fp32 cache.
// ------------------------------------------------------------------------------------------------
constexpr int PF_HALF = 64;      // tokens per 128-row TMA tile
constexpr int PA_QT = 32, PA_THREADS = 256, PA_MAXL = 128;

__global__ void embed_rows_kernel(const int* __restrict__ ids, const bf16* __restrict__ embed, float* __restrict__ x,
                                  int H, int V) {
    const int t = blockIdx.x;
    int tok = ids[t];
    tok = min(max(tok, 0), V - 1);
    for (int i = threadIdx.x; i < H; i += blockDim.x) x[(long long)t * H + i] = __bfloat162float(embed[(long long)tok * H + i]);
}

// cos/sin of pos / freqs[d] for pos < L (MLXFast.RoPE with custom freqs, LlamaTTS.swift:192-200)
__global__ void rope_table_kernel(const float* __restrict__ freqs, float2* __restrict__ tab, int L) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L * (HD / 2)) return;
    const int p = i / (HD / 2), d = i - p * (HD / 2);
    float s, c;
    sincosf((float)p / freqs[d], &s, &c);
    tab[i] = make_float2(c, s);
}

struct PrefillAttnArgs {
    const float* qkv;     // [T, (nq + 2 nkv) * 128]
    const float2* rope;   // [L, 64] (cos, sin)
    float* kcache;        // this layer [B][nkv][max_ctx][128]
    float* vcache;
    bf16* out;            // [2 * T_pad, nq * 128] hi/lo, 64-token tiles
    int nq, nkv, max_ctx, L;
    float scale;
};

template <int G>
__global__ void __launch_bounds__(PA_THREADS)
prefill_attn_kernel(PrefillAttnArgs a) {
    extern __shared__ __align__(16) uint8_t pa_smem[];
    const int h = blockIdx.x, b = blockIdx.y, q0 = blockIdx.z * PA_QT;
    const int nqt = min(PA_QT, a.L - q0), kmax = q0 + nqt;     // causal: keys 0 .. q0 + nqt - 1
    float* sK = reinterpret_cast<float*>(pa_smem);              // [kmax][128]
    float* sV = sK + (size_t)a.L * HD;                          // [kmax][128]
    float* sQ = sV + (size_t)a.L * HD;                          // [G][PA_QT][128]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = (a.nq + 2 * a.nkv) * HD;
    const float* base = a.qkv + (long long)b * a.L * ld;
    float* kc = a.kcache + (((long long)b * a.nkv + h) * a.max_ctx) * HD;
    float* vc = a.vcache + (((long long)b * a.nkv + h) * a.max_ctx) * HD;

    // K (RoPE) and V of keys [0, kmax) -> shared memory; this tile's own keys also go to the cache
    for (int i = tid; i < kmax * (HD / 2); i += PA_THREADS) {
        const int t = i / (HD / 2), d = i - t * (HD / 2);
        const float2 cs = a.rope[t * (HD / 2) + d];
        const float* k = base + (long long)t * ld + (a.nq + h) * HD;
        const float x1 = k[d], x2 = k[d + HD / 2];
        const float k1 = x1 * cs.x - x2 * cs.y, k2 = x2 * cs.x + x1 * cs.y;
        sK[t * HD + d] = k1; sK[t * HD + d + HD / 2] = k2;
        if (t >= q0) { kc[(long long)t * HD + d] = k1; kc[(long long)t * HD + d + HD / 2] = k2; }
    }
    for (int i = tid; i < kmax * HD; i += PA_THREADS) {
        const int t = i / HD, d = i - t * HD;
        const float v = base[(long long)t * ld + (a.nq + a.nkv + h) * HD + d];
        sV[i] = v;
        if (t >= q0) vc[(long long)t * HD + d] = v;
    }
    for (int i = tid; i < G * nqt * (HD / 2); i += PA_THREADS) {
        const int g = i / (nqt * (HD / 2)), r = i - g * nqt * (HD / 2), qi = r / (HD / 2), d = r - qi * (HD / 2);
        const float2 cs = a.rope[(q0 + qi) * (HD / 2) + d];
        const float* q = base + (long long)(q0 + qi) * ld + (h * G + g) * HD;
        const float x1 = q[d], x2 = q[d + HD / 2];
        sQ[(g * PA_QT + qi) * HD + d] = x1 * cs.x - x2 * cs.y;
        sQ[(g * PA_QT + qi) * HD + d + HD / 2] = x2 * cs.x + x1 * cs.y;
    }
    __syncthreads();

    // one (head, query) row per warp iteration; lanes over keys for the scores, over dims for P*V
    constexpr int KJ = PA_MAXL / 32;
    for (int r = warp; r < G * nqt; r += PA_THREADS / 32) {
        const int g = r / nqt, qi = r - g * nqt, qpos = q0 + qi;
        const float4* q4 = reinterpret_cast<const float4*>(sQ + (g * PA_QT + qi) * HD);
        float sc[KJ];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < KJ; ++j) {
            const int t = j * 32 + lane;
            float acc = -INFINITY;
            if (t <= qpos) {
                const float4* k4 = reinterpret_cast<const float4*>(sK + t * HD);
                acc = 0.f;
#pragma unroll 8
                for (int i = 0; i < HD / 4; ++i) {
                    const int d4 = (i + lane) & (HD / 4 - 1);     // rotate: conflict-free rows 512 B apart
                    const float4 kf = k4[d4], qf = q4[d4];
                    acc = fmaf(qf.x, kf.x, acc); acc = fmaf(qf.y, kf.y, acc);
                    acc = fmaf(qf.z, kf.z, acc); acc = fmaf(qf.w, kf.w, acc);
                }
                acc *= a.scale;
            }
            sc[j] = acc;
            mx = fmaxf(mx, acc);
        }
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < KJ; ++j) { sc[j] = sc[j] == -INFINITY ? 0.f : __expf(sc[j] - mx); sum += sc[j]; }
        sum = warp_sum(sum);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < KJ; ++j) {
            if (j * 32 > qpos) break;
            for (int l = 0; l < 32; ++l) {
                const int t = j * 32 + l;
                if (t > qpos) break;
                const float p = __shfl_sync(0xffffffffu, sc[j], l);
                const float4 v = reinterpret_cast<const float4*>(sV + t * HD)[lane];
                o.x = fmaf(p, v.x, o.x); o.y = fmaf(p, v.y, o.y); o.z = fmaf(p, v.z, o.z); o.w = fmaf(p, v.w, o.w);
            }
        }
        const float inv = 1.0f / sum;
        const int tok = b * a.L + qpos;
        const long long col = (long long)(h * G + g) * HD + lane * 4;
        const long long ldo = (long long)a.nq * HD;
        store_hilo(a.out, ldo, tok, col + 0, o.x * inv, PF_HALF);
        store_hilo(a.out, ldo, tok, col + 1, o.y * inv, PF_HALF);
        store_hilo(a.out, ldo, tok, col + 2, o.z * inv, PF_HALF);
        store_hilo(a.out, ldo, tok, col + 3, o.w * inv, PF_HALF);
    }
}

// decode buffers <- last prompt position of every row:  x[b] = xp[b*L + L-1], y[b] = yp[...], pos[b] = L - 1
__global__ void gather_last_kernel(const float* __restrict__ xp, const float* __restrict__ yp, float* __restrict__ x,
                                   float* __restrict__ y, int* __restrict__ pos, int L, int H) {
    const int b = blockIdx.x;
    const long long src = ((long long)b * L + L - 1) * H;
    for (int i = threadIdx.x; i < H; i += blockDim.x) {
        x[(long long)b * H + i] = xp[src + i];
        y[(long long)b * H + i] = yp[src + i];
    }
    if (threadIdx.x == 0) pos[b] = L - 1;
}

// ------------------------------------------------------------------------------------------------
// Logits processors + sampler + bookkeeping: one CTA per row.
// ------------------------------------------------------------------------------------------------
constexpr int SM_THREADS = 1024, SM_WARPS = SM_THREADS / 32, SM_MAX_ATTEMPTS = 12;

struct SampleArgs {
    float* logits;        // [B, V]  (modified in place: penalty, EOS mask)
    float* probs;         // [B, V]  scratch
    int* tokens;          // [B] next input token (written)
    int* pos;             // [B] position of the NEXT input token (incremented)
    int* recent;          // [B, R] ring of the last R tokens (prompt + generated)
    int* recent_n;        // [B] total tokens pushed so far
    int* out_tokens;      // [B, max_tokens]
    int* n_gen;           // [B]
    int* done;            // [B]
    int* n_active;        // [1] rows not yet finished (decremented when a row emits EOS)
    const int* forced;    // nullable [B]: if set, ignore the sampler and emit forced[b] (prefill)
    int V, R, max_tokens;
    float temperature, top_p, rep_penalty;
    unsigned long long seed;
    int mask_eos;
};

__device__ __forceinline__ float block_sum_1024(float v, float* sred) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < SM_WARPS; ++i) t += sred[i];
    return t;
}

__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long a, unsigned long long b,
                                           unsigned long long c) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (a * 1000003ull + b * 131ull + c + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// Logits processors + sampler (deterministic for a given seed, no atomics):
//   RepetitionContext.process -> EOS mask (bench only) -> argmax | TopPSampler.
// Top-p = "sample from the smallest top set whose mass reaches top_p".  Implemented as rejection sampling: draw
// token ~ softmax(l/T) by the Gumbel-max trick (token = argmax_i l_i/T + g_i, g_i = -log(-log(u_i)), u_i a hash of (seed, row,
// step, attempt, i): one coalesced pass, no inverse-CDF walk), accept iff the mass of strictly more probable tokens is < top_p
// (exactly the nucleus membership test; ties are all kept).  The acceptance probability is >= top_p, so a handful of
// attempts suffice; after SM_MAX_ATTEMPTS the argmax (always in the nucleus) is returned.
// One thread-block CLUSTER of SM_CLUSTER CTAs per row: every CTA owns a contiguous 1/8 of the vocabulary (19 618 of Orpheus's
// 156 940 logits: 20 per thread) and the per-pass (max, index) / sums are exchanged through distributed shared memory -- every CTA
// writes its partial into every peer's slot, one cluster barrier, every CTA reduces the 8 partials in rank order.  (Round 1 ran one
// CTA per row: 8 of 148 SMs, and the sampling passes were bound by that one SM's issue rate: 104-148 us per launch against 17 us
// for the greedy pick, 6 % of the decode step.)
constexpr int SM_CLUSTER = 8;
struct SampleExchange { float v[2][SM_CLUSTER]; int i[2][SM_CLUSTER]; float s[2][SM_CLUSTER]; };

__global__ void __cluster_dims__(SM_CLUSTER, 1, 1) __launch_bounds__(SM_THREADS)
sample_kernel(SampleArgs a) {
    namespace cgr = cooperative_groups;
    cgr::cluster_group cluster = cgr::this_cluster();
    const int rank = (int)cluster.block_rank();
    __shared__ float sred[SM_WARPS];
    __shared__ float s_val[SM_WARPS];
    __shared__ int s_idx[SM_WARPS];
    __shared__ SampleExchange ex;
    const int b = blockIdx.x / SM_CLUSTER, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    pdl_wait();                  // NO pdl_trigger(): this kernel writes pos[] (see attn_decode_cluster_kernel)
    float* lg = a.logits + (long long)b * a.V;
    const int chunk = (a.V + SM_CLUSTER - 1) / SM_CLUSTER, i0 = rank * chunk, i1 = min(a.V, i0 + chunk);
    const int nrec = min(a.recent_n[b], a.R);
    int parity = 0;
    int tok_final = 0;

    // (value, index) and a sum: block reduce, then all-to-all through distributed shared memory; every CTA ends with the same result
    auto exchange = [&](float& v, int& i, float& sum) {
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, i, o);
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
        }
        sum = warp_sum(sum);
        __syncthreads();
        if (lane == 0) { s_val[warp] = v; s_idx[warp] = i; sred[warp] = sum; }
        __syncthreads();
        if (warp == 0) {
            float bv = s_val[lane], bs = sred[lane];
            int bi = s_idx[lane];
            for (int o = 16; o; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            bs = warp_sum(bs);
            if (lane < SM_CLUSTER) {             // lane p writes this CTA's partial into CTA p's slot [rank]
                SampleExchange* peer = cluster.map_shared_rank(&ex, lane);
                peer->v[parity][rank] = bv; peer->i[parity][rank] = bi; peer->s[parity][rank] = bs;
            }
        }
        cluster.sync();
        v = ex.v[parity][0]; i = ex.i[parity][0]; sum = ex.s[parity][0];
#pragma unroll
        for (int r = 1; r < SM_CLUSTER; ++r) {
            const float ov = ex.v[parity][r];
            const int oi = ex.i[parity][r];
            if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
            sum += ex.s[parity][r];
        }
        parity ^= 1;
    };

    if (a.forced == nullptr) {
        // RepetitionContext.process: once per unique token among the last R; every CTA handles the tokens of its own range
        if (a.rep_penalty != 1.0f && t < nrec) {
            const int tok = a.recent[b * a.R + t];
            bool dup = false;
            for (int j = 0; j < t; ++j) dup |= (a.recent[b * a.R + j] == tok);
            if (!dup && tok >= i0 && tok < i1) {
                const float l = lg[tok];
                lg[tok] = l < 0.f ? l * a.rep_penalty : l / a.rep_penalty;
            }
        }
        if (a.mask_eos && t == 0 && TOK_END_OF_SPEECH >= i0 && TOK_END_OF_SPEECH < i1) lg[TOK_END_OF_SPEECH] = -INFINITY;
        __syncthreads();

        // max (and argmax, lowest index wins ties)
        float best = -INFINITY, dummy = 0.f;
        int bi = 0x7fffffff;
        for (int i = i0 + t; i < i1; i += SM_THREADS) {
            const float v = lg[i];
            if (v > best) { best = v; bi = i; }
        }
        exchange(best, bi, dummy);
        const float mx = best;
        tok_final = bi;

        if (a.temperature > 0.f) {
            const float inv_t = 1.0f / a.temperature;
            const int step = a.n_gen[b];
            auto gumbel_key = [&](int i, float l, int att) {
                unsigned long long z = a.seed + 0x9E3779B97F4A7C15ull * ((unsigned long long)b * 1000003ull + (unsigned long long)step * 131ull + (unsigned long long)att + 1ull);
                z ^= (unsigned long long)(unsigned)i * 0xD6E8FEB86659FD93ull;
                z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
                z ^= z >> 31;
                const float u = ((float)(z >> 40) + 0.5f) * (1.0f / 16777216.0f);      // (0, 1)
                return l * inv_t - __logf(-__logf(u));
            };
            // pass A: Z = sum exp((l - max) / T) and the first draw
            float z = 0.f, gv = -INFINITY;
            int gi = 0x7fffffff;
            for (int i = i0 + t; i < i1; i += SM_THREADS) {
                const float l = lg[i];
                z += __expf((l - mx) * inv_t);
                const float k = l == -INFINITY ? -INFINITY : gumbel_key(i, l, 0);
                if (k > gv) { gv = k; gi = i; }
            }
            exchange(gv, gi, z);
            const float Z = z;
            int cand = gi;
            bool accepted = a.top_p >= 1.0f;
            for (int att = 0; !accepted; ++att) {
                // nucleus test: the mass of strictly more probable tokens must be below top_p
                const float lc = lg[cand];
                float gm = 0.f, dv = -INFINITY;
                int di = 0x7fffffff;
                for (int i = i0 + t; i < i1; i += SM_THREADS) { const float l = lg[i]; if (l > lc) gm += __expf((l - mx) * inv_t); }
                exchange(dv, di, gm);
                if (gm < a.top_p * Z) { accepted = true; break; }
                if (att + 1 >= SM_MAX_ATTEMPTS) { cand = tok_final; break; }      // the argmax is always in the nucleus
                gv = -INFINITY; gi = 0x7fffffff;
                float ds = 0.f;
                for (int i = i0 + t; i < i1; i += SM_THREADS) {
                    const float l = lg[i];
                    const float k = l == -INFINITY ? -INFINITY : gumbel_key(i, l, att + 1);
                    if (k > gv) { gv = k; gi = i; }
                }
                exchange(gv, gi, ds);
                cand = gi;
            }
            if (cand >= 0 && cand < a.V) tok_final = cand;
        }
    } else {
        tok_final = a.forced[b];
    }

    if (rank == 0 && t == 0) {
        const int tok = tok_final;
        a.tokens[b] = tok;
        a.pos[b] += 1;
        const int rn = a.recent_n[b];
        if (a.R > 0) {
            // keep `recent` as "last R tokens" in arrival order: shift when full
            if (rn < a.R) a.recent[b * a.R + rn] = tok;
            else {
                for (int j = 1; j < a.R; ++j) a.recent[b * a.R + j - 1] = a.recent[b * a.R + j];
                a.recent[b * a.R + a.R - 1] = tok;
            }
            a.recent_n[b] = rn + 1;
        }
        if (a.forced == nullptr && !a.done[b]) {
            if (tok == TOK_END_OF_SPEECH) {      // LlamaTTS.swift:734-736: stop, token not appended
                a.done[b] = 1;
                atomicSub(a.n_active, 1);
            } else {
                const int n = a.n_gen[b];
                if (n < a.max_tokens) a.out_tokens[b * a.max_tokens + n] = tok;
                a.n_gen[b] = n + 1;
                if (n + 1 >= a.max_tokens) { a.done[b] = 1; atomicSub(a.n_active, 1); }
            }
        }
    }
}

// prefill bookkeeping for positions that do not need logits: next token = ids[b, pos+1]
__global__ void prefill_advance_kernel(const int* __restrict__ ids, int L, int* tokens, int* pos, int B) {
    const int b = threadIdx.x;
    pdl_wait();                  // NO pdl_trigger(): writes pos[]
    if (b >= B) return;
    const int p = pos[b] + 1;
    pos[b] = p;
    if (p < L) tokens[b] = ids[b * L + p];
}

__global__ void init_rows_kernel(const int* __restrict__ ids, int L, int B, int R, int* tokens, int* pos, int* recent,
                                 int* recent_n, int* n_gen, int* done, int* n_active, int start_pos) {
    const int b = threadIdx.x;
    if (b == 0) *n_active = B;
    if (b >= B) return;
    tokens[b] = ids[b * L];
    pos[b] = start_pos;
    n_gen[b] = 0;
    done[b] = 0;
    // processor.prompt(promptTokens): the ring starts with the last R prompt tokens
    const int n = min(R, L);
    for (int j = 0; j < n; ++j) recent[b * R + j] = ids[b * L + L - n + j];
    recent_n[b] = n;
}

// device-side random init (benchmarks / full-size property tests): N(0, std^2) -> bf16
__global__ void random_bf16_kernel(bf16* __restrict__ w, long long n, float std, unsigned long long seed) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        const float u1 = ((unsigned)(z >> 40) + 1.0f) * (1.0f / 16777217.0f);
        const float u2 = (unsigned)((z >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
        w[i] = __float2bfloat16_rn(std * sqrtf(-2.0f * __logf(u1)) * cospif(2.0f * u2));
    }
}
__global__ void fill_f32_kernel(float* p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static std::vector<float> llama3_freqs(const b2a_llama_config& c) {
    // LlamaTTS.swift:121-156 in Float
    const int d = c.head_dim;
    std::vector<float> f(d / 2);
    const float low_wl = c.rope_old_context_len / c.rope_low_freq_factor;
    const float high_wl = c.rope_old_context_len / c.rope_high_freq_factor;
    for (int i = 0; i < d / 2; ++i) {
        float fr = powf(c.rope_theta, (float)(2 * i) / (float)d);
        const float wl = 2.0f * (float)M_PI * fr;
        const float base = fr;
        if (wl > low_wl) fr = fr * c.rope_factor;
        const bool med = wl > high_wl && wl < low_wl;
        if (med) {
            const float smooth = (c.rope_old_context_len / wl - c.rope_low_freq_factor) /
                                 (c.rope_high_freq_factor - c.rope_low_freq_factor);
            fr = base / ((1.0f - smooth) / c.rope_factor + smooth);
        }
        f[i] = fr;
    }
    return f;
}

struct LayerW {
    DBuf<bf16> wqkv, wo, wgu, wdown;
    DBuf<float> ln1, ln2;
    DBuf<float> qnorm, knorm;   // [128] each, only with StackSpec::qk_norm
};

// Which checkpoint keys a transformer stack is built from.  Orpheus: {"model.", embeddings + tied head}.  The Qwen3-TTS talker
// and code predictor (row N1) reuse the same engine with per-head q/k RMSNorm, inputs given as EMBEDDINGS and heads owned by the
// caller (Qwen3TTSTalker.swift:127-310, Qwen3TTSCodePredictor.swift:14-240).
struct StackSpec {
    std::string prefix = "model.";            // "<prefix>layers.N....", "<prefix>norm.weight"
    bool qk_norm = false;                      // self_attn.q_norm / k_norm
    bool has_embed = true;                     // "<prefix>embed_tokens.weight" (false: inputs are embeddings)
    std::string head;                          // "" = tied to the embedding / none; else an untied [vocab, hidden] matrix
    bool has_head = true;
};

}  // namespace b2a

using namespace b2a;

struct b2a_tts {
    int device;
    b2a_llama_config cfg;
    b2a_snac* snac;
    cudaStream_t stream = nullptr;
    std::vector<LayerW> layers;
    DBuf<bf16> embed, lm_head_w;
    const bf16* lm_head = nullptr;
    DBuf<float> final_ln, freqs;
    DBuf<float> kcache, vcache;   // [layer][B][nkv][ctx][hd] fp32
    // activations: fp32 residual stream; GEMM inputs as [16, K] bf16 hi/lo pairs
    DBuf<float> x, y, qkv, logits, probs;
    DBuf<bf16> xn, attn, act;
    // attention workspace (flash-decoding partials)
    DBuf<float> part_o, part_ml;
    DBuf<int> at_counters;
    int at_splits = 1;
    // tcgen05 / TMA path
    bool use_tc = true;
    int num_sms = 148;
    std::vector<CUtensorMap> tm_qkv, tm_o, tm_gu, tm_down;
    // weight rows per m-tile for a whole-tile GEMM of M rows (tc::Args::tile_rows): 128 unless that leaves > 1/4 of the SMs idle
    static int pick_tile_rows(int M, int sms) {
        if (cdiv(M, tc::BM) * 4 >= sms * 3) return tc::BM;
        return std::max(8, std::min(tc::BM, cdiv(cdiv(M, sms), 8) * 8));
    }
    int lm_tile_rows = 128, head_rows_now = 0;   // head_rows_now: tile rows of the map passed to the current OP_LM launch (0 = lm_tile_rows)
    std::vector<CUtensorMap> tm_gu_dec;     // gate/up with gu_tile_rows-row boxes for the decode step (tc::Args::tile_rows)
    int gu_tile_rows = 0;
    CUtensorMap tm_lm{}, tmx_xn{}, tmx_attn{}, tmx_act{};
    // batched prefill workspace (sized for the largest B*L seen so far)
    DBuf<float> xp, yp, qkvp;
    DBuf<bf16> xnp, attnp, actp;
    DBuf<float2> rope_tab;
    CUtensorMap tmp_xn{}, tmp_attn{}, tmp_act{};
    int pf_tokens_cap = 0;
    bool use_batched_prefill = true;
    DBuf<int> tokens, pos, recent, recent_n, out_tokens, n_gen, done, n_active, ids, forced;
    HBuf<int> h_flag;
    cudaEvent_t ev_poll[2] = {nullptr, nullptr};   // the generate loop's pipelined "rows still active" polls
    // fused-norm decode step (default on the tcgen05 path): o_proj / down_proj run as cluster split-K GEMMs whose leader CTA does the
    // residual add + the next norm's gain + hi/lo split + sum of squares; no stand-alone add_rmsnorm launches (tc_gemm.cuh)
    bool fused = false;
    int fused_cluster = 6, fused_parts = 0;   // 6 CTAs per 128-row tile: 144 of 148 SMs for hidden 3072 (measured best of 3..8)
    DBuf<float> ss_a, ss_b;          // [H / 128, 8] partial sums of squares: ss_a feeds the post-attention norm, ss_b the input norm
    StackSpec spec;                  // which keys / features this stack was built with
    const float* x_ext = nullptr;    // row N1: when set, a step starts from these embeddings [8, H] instead of embed(tokens)
    float* normed_out = nullptr;     // row N1: when set, the final RMSNorm also writes its fp32 output here [8, H]
    std::atomic<int> cancel{0};
    int bench_mask_eos = 0, bench_wrap_codes = 0;   // b2a_tts_set_bench_flags (include/b200audio_internal.h): fixed-work benchmark switches
    int nb_pad = 0;   // rows rounded up to 1/2/4/8
    bool trace_on = false;       // debug: residual stream at every RMSNorm input (eager forward only)
    DBuf<float> trace;           // [2*layers + 1][8][H]
    // CUDA graphs for the two step flavours (captured per (nb_pad, params) configuration)
    cudaGraphExec_t g_step = nullptr, g_prefill = nullptr;
    SampleArgs g_args{};
    int g_nb = 0;
    // codec workspaces
    DBuf<int> d_codes[3];
    DBuf<float> d_wave;

    ~b2a_tts() {
        if (g_step) cudaGraphExecDestroy(g_step);
        if (g_prefill) cudaGraphExecDestroy(g_prefill);
        for (auto& e : ev_poll) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }

    static void upload_bf16(const TensorTable& tt, const std::string& name, int64_t expect, DBuf<bf16>& dst, size_t offset_elems,
                            size_t total_elems) {
        const b2a_tensor& t = tt.get(name);
        B2A_CHECK(t.dtype == B2A_DTYPE_BF16 || t.dtype == B2A_DTYPE_F32, B2A_ERR_MODEL_NOT_INITIALIZED, "tensor must be bf16 or f32: " + name);
        B2A_CHECK(TensorTable::numel(t) == expect, B2A_ERR_MODEL_NOT_INITIALIZED, "bad shape for tensor: " + name);
        dst.alloc(total_elems);
        if (t.dtype == B2A_DTYPE_BF16) {
            B2A_CUDA(cudaMemcpy(dst.p + offset_elems, t.data, expect * sizeof(bf16), cudaMemcpyHostToDevice));
        } else {   // an fp32 checkpoint: the engine holds bf16 matrices, round to nearest even (stated deviation, as for Whisper)
            std::vector<bf16> tmp((size_t)expect);
            const float* src = (const float*)t.data;
            for (int64_t i = 0; i < expect; ++i) tmp[i] = __float2bfloat16_rn(src[i]);
            B2A_CUDA(cudaMemcpy(dst.p + offset_elems, tmp.data(), expect * sizeof(bf16), cudaMemcpyHostToDevice));
        }
    }

    template <int G>
    static void attn_attr() {
        B2A_CUDA(cudaFuncSetAttribute(attn_decode_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    }
    template <int G>
    static void attn_loop_attr() {
        B2A_CUDA(cudaFuncSetAttribute(attn_decode_loop_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));   // + 512 B static
    }
    int attn_loop_bufs() const {
        const int G = cfg.num_attention_heads / cfg.num_key_value_heads;
        const size_t extra = (size_t)(G * HD + 2 * HD + (AT_THREADS / 32) * G * HD + G * HD + 2 * MAXG) * sizeof(float) + 64;
        return (int)std::min<size_t>(3, (226 * 1024 - extra) / ((size_t)2 * AT_CAP * HD * sizeof(float)));
    }
    template <int G>
    static void attn_cluster_attr() {
        B2A_CUDA(cudaFuncSetAttribute(attn_decode_cluster_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    }
    bool attn_loop = true;     // B2A_ATTN=split selects the flash-decoding split kernel, =loop the single-CTA loop
    bool attn_cluster = true;  // default: two-CTA cluster per (kv head, row)
    void attn_launch(const AttnArgs& aa, int B, cudaStream_t s) {
        if (attn_loop && attn_cluster) {
            const int G = aa.nq / aa.nkv, NB = attn_loop_bufs();
            const size_t sm = (size_t)NB * 2 * AT_CAP * HD * sizeof(float) +
                              (size_t)(G * HD + 2 * HD + (AT_THREADS / 32) * G * HD + G * HD + 2 * MAXG) * sizeof(float) + 64;
            const dim3 g3(aa.nkv, B, 2);
            switch (G) {
                case 1: launch_pdl(attn_decode_cluster_kernel<1>, g3, dim3(AT_THREADS), sm, s, aa, NB); break;
                case 2: launch_pdl(attn_decode_cluster_kernel<2>, g3, dim3(AT_THREADS), sm, s, aa, NB); break;
                case 3: launch_pdl(attn_decode_cluster_kernel<3>, g3, dim3(AT_THREADS), sm, s, aa, NB); break;
                case 4: launch_pdl(attn_decode_cluster_kernel<4>, g3, dim3(AT_THREADS), sm, s, aa, NB); break;
                case 6: launch_pdl(attn_decode_cluster_kernel<6>, g3, dim3(AT_THREADS), sm, s, aa, NB); break;
                default: launch_pdl(attn_decode_cluster_kernel<8>, g3, dim3(AT_THREADS), sm, s, aa, NB); break;
            }
            return;
        }
        if (attn_loop) {
            const int G = aa.nq / aa.nkv, NB = attn_loop_bufs();
            const size_t sm = (size_t)NB * 2 * AT_CAP * HD * sizeof(float) + (size_t)(G * HD + 2 * HD + (AT_THREADS / 32) * G * HD) * sizeof(float) + 64;
            const dim3 g2(aa.nkv, B);
            switch (G) {
                case 1: launch_pdl(attn_decode_loop_kernel<1>, g2, dim3(AT_THREADS), sm, s, aa, NB); break;
                case 2: launch_pdl(attn_decode_loop_kernel<2>, g2, dim3(AT_THREADS), sm, s, aa, NB); break;
                case 3: launch_pdl(attn_decode_loop_kernel<3>, g2, dim3(AT_THREADS), sm, s, aa, NB); break;
                case 4: launch_pdl(attn_decode_loop_kernel<4>, g2, dim3(AT_THREADS), sm, s, aa, NB); break;
                case 6: launch_pdl(attn_decode_loop_kernel<6>, g2, dim3(AT_THREADS), sm, s, aa, NB); break;
                default: launch_pdl(attn_decode_loop_kernel<8>, g2, dim3(AT_THREADS), sm, s, aa, NB); break;
            }
            return;
        }
        const dim3 grid(aa.nkv, B, aa.S);
        const size_t sm = attn_smem_bytes();
        switch (aa.nq / aa.nkv) {
            case 1: launch_pdl(attn_decode_kernel<1>, grid, dim3(AT_THREADS), sm, s, aa); break;
            case 2: launch_pdl(attn_decode_kernel<2>, grid, dim3(AT_THREADS), sm, s, aa); break;
            case 3: launch_pdl(attn_decode_kernel<3>, grid, dim3(AT_THREADS), sm, s, aa); break;
            case 4: launch_pdl(attn_decode_kernel<4>, grid, dim3(AT_THREADS), sm, s, aa); break;
            case 6: launch_pdl(attn_decode_kernel<6>, grid, dim3(AT_THREADS), sm, s, aa); break;
            default: launch_pdl(attn_decode_kernel<8>, grid, dim3(AT_THREADS), sm, s, aa); break;
        }
    }
    size_t attn_smem_bytes() const {
        const int G = cfg.num_attention_heads / cfg.num_key_value_heads;
        return (size_t)(2 * AT_CAP * HD + G * HD + G * AT_CAP + (AT_THREADS / 32) * G * HD) * sizeof(float) + 16;
    }

    void check_config() {
        const b2a_llama_config& c = cfg;
        B2A_CHECK(c.head_dim == HD, B2A_ERR_INVALID_INPUT, "llama: head_dim must be 128");
        B2A_CHECK(c.hidden_size % 8 == 0 && c.intermediate_size % 8 == 0, B2A_ERR_INVALID_INPUT, "llama: sizes must be multiples of 8");
        {
            const int g = c.num_key_value_heads > 0 && c.num_attention_heads % c.num_key_value_heads == 0
                              ? c.num_attention_heads / c.num_key_value_heads : 0;
            B2A_CHECK(g == 1 || g == 2 || g == 3 || g == 4 || g == 6 || g == 8, B2A_ERR_INVALID_INPUT,
                      "llama: unsupported GQA ratio (q heads per kv head must be 1, 2, 3, 4, 6 or 8)");
        }
        B2A_CHECK(c.max_batch >= 1 && c.max_batch <= 8, B2A_ERR_INVALID_INPUT, "llama: max_batch must be in 1..8");
        B2A_CHECK(c.max_context >= 8, B2A_ERR_INVALID_INPUT, "llama: max_context too small");
        require_device(device);
        B2A_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    }

    void alloc_state() {
        const b2a_llama_config& c = cfg;
        const int H = c.hidden_size, I = c.intermediate_size, nq = c.num_attention_heads, nkv = c.num_key_value_heads;
        const int NQ = nq * HD, NKV = nkv * HD;
        B2A_CHECK(H <= RN_THREADS * RN_MAXV, B2A_ERR_INVALID_INPUT, "llama: hidden_size above 8192 is not supported");
        std::vector<float> fr = llama3_freqs(c);
        freqs.upload(fr.data(), fr.size());
        const size_t kv = (size_t)c.num_hidden_layers * c.max_batch * nkv * c.max_context * HD;
        kcache.alloc(kv);
        vcache.alloc(kv);
        B2A_CUDA(cudaMemset(kcache.p, 0, kv * sizeof(float)));
        B2A_CUDA(cudaMemset(vcache.p, 0, kv * sizeof(float)));
        const int B = 8, R16 = 16;
        x.alloc((size_t)B * H); y.alloc((size_t)B * H); qkv.alloc((size_t)B * (NQ + 2 * NKV));
        logits.alloc((size_t)B * c.vocab_size); probs.alloc((size_t)B * c.vocab_size);
        xn.alloc((size_t)R16 * H); attn.alloc((size_t)R16 * NQ); act.alloc((size_t)R16 * I);
        B2A_CUDA(cudaMemset(xn.p, 0, (size_t)R16 * H * sizeof(bf16)));
        B2A_CUDA(cudaMemset(attn.p, 0, (size_t)R16 * NQ * sizeof(bf16)));
        B2A_CUDA(cudaMemset(act.p, 0, (size_t)R16 * I * sizeof(bf16)));
        B2A_CUDA(cudaMemset(x.p, 0, (size_t)B * H * sizeof(float)));
        B2A_CUDA(cudaMemset(y.p, 0, (size_t)B * H * sizeof(float)));
        B2A_CUDA(cudaMemset(qkv.p, 0, (size_t)B * (NQ + 2 * NKV) * sizeof(float)));
        tokens.alloc(B); pos.alloc(B); recent.alloc(B * 64); recent_n.alloc(B); n_gen.alloc(B); done.alloc(B);
        n_active.alloc(1); forced.alloc(B);
        B2A_CUDA(cudaMemset(tokens.p, 0, B * sizeof(int)));
        B2A_CUDA(cudaMemset(pos.p, 0, B * sizeof(int)));
        h_flag.alloc(16);
        // attention workspace
        const int G = nq / nkv;
        at_splits = cdiv(c.max_context, AT_CAP);
        part_o.alloc((size_t)B * nkv * at_splits * G * HD);
        part_ml.alloc((size_t)B * nkv * at_splits * G * 2);
        at_counters.alloc((size_t)B * nkv);
        B2A_CUDA(cudaMemset(at_counters.p, 0, (size_t)B * nkv * sizeof(int)));
        // process-wide kernel attributes: always the same (largest) value, several handles may coexist
        gemv_attrs<1>(); gemv_attrs<2>(); gemv_attrs<4>(); gemv_attrs<8>();
        B2A_CHECK(attn_smem_bytes() <= 220 * 1024, B2A_ERR_INVALID_INPUT, "llama: GQA ratio too large for the attention tile");
        attn_attr<1>(); attn_attr<2>(); attn_attr<3>(); attn_attr<4>(); attn_attr<6>(); attn_attr<8>();
        attn_loop_attr<1>(); attn_loop_attr<2>(); attn_loop_attr<3>(); attn_loop_attr<4>(); attn_loop_attr<6>(); attn_loop_attr<8>();
        attn_cluster_attr<1>(); attn_cluster_attr<2>(); attn_cluster_attr<3>(); attn_cluster_attr<4>(); attn_cluster_attr<6>(); attn_cluster_attr<8>();
        { const char* e = getenv("B2A_ATTN"); attn_loop = !(e && std::string(e) == "split"); attn_cluster = !(e && std::string(e) == "loop"); }
        if (spec.qk_norm) attn_loop = attn_cluster = true;   // only the cluster kernel applies the per-head q/k RMSNorm
        {
            const char* e = getenv("B2A_FUSED");
            const char* cl = getenv("B2A_CLUSTER");
            if (cl) fused_cluster = std::max(1, std::min(tc::SPLIT_MAX_CLUSTER, atoi(cl)));
            fused_parts = H / tc::BM;
            const char* g = getenv("B2A_GEMM");
            const bool tc_ok = !(g && std::string(g) == "simt") && H % tc::BK == 0 && NQ % tc::BK == 0 && I % tc::BK == 0;
            fused = !(e && std::string(e) == "0") && tc_ok && attn_loop && attn_cluster && H % tc::BM == 0 && fused_parts <= 64;
            ss_a.alloc((size_t)64 * 8); ss_b.alloc((size_t)64 * 8);
            B2A_CUDA(cudaMemset(ss_a.p, 0, 64 * 8 * sizeof(float)));
            B2A_CUDA(cudaMemset(ss_b.p, 0, 64 * 8 * sizeof(float)));
        }
        B2A_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
        // tcgen05 / TMA path: needs every GEMM K to be a multiple of 64; B2A_GEMM=simt forces the SIMT fallback
        const char* env = getenv("B2A_GEMM");
        use_tc = !(env && std::string(env) == "simt") && H % tc::BK == 0 && NQ % tc::BK == 0 && I % tc::BK == 0;
        const char* envp = getenv("B2A_PREFILL");
        use_batched_prefill = !(envp && std::string(envp) == "step");
        if (use_tc) {
            tc::set_attributes();
            for (auto& L : layers) {
                tm_qkv.push_back(tc::make_tmap_bf16(L.wqkv.p, NQ + 2 * NKV, H, tc::BM));
                tm_o.push_back(tc::make_tmap_bf16(L.wo.p, H, NQ, tc::BM));
                tm_gu.push_back(tc::make_tmap_bf16(L.wgu.p, 2 * I, H, tc::BM));
                {   // decode step: when 128-row tiles would leave more than a quarter of the SMs idle (Qwen3-TTS: 6144 rows = 48 tiles), use
                    // as many m-tiles as SMs (rows per tile a multiple of 8): frame 3.47 -> 3.38 ms.  Orpheus (128 tiles on 148 SMs) measured
                    // no gain from 147 tiles of 112 rows (1.796 vs 1.812 ms per step) and keeps 128.  B2A_GU_ROWS overrides.
                    static const int e_rows = getenv("B2A_GU_ROWS") ? atoi(getenv("B2A_GU_ROWS")) : 0;
                    int rows = e_rows > 0 ? e_rows : (cdiv(2 * I, tc::BM) * 4 >= num_sms * 3 ? tc::BM : cdiv(cdiv(2 * I, num_sms), 8) * 8);
                    rows = std::max(8, std::min(tc::BM, rows / 8 * 8));
                    gu_tile_rows = rows;
                    tm_gu_dec.push_back(tc::make_tmap_bf16(L.wgu.p, 2 * I, H, rows));
                }
                tm_down.push_back(tc::make_tmap_bf16(L.wdown.p, H, I, tc::BM));
            }
            lm_tile_rows = pick_tile_rows(c.vocab_size, num_sms);
            if (lm_head) tm_lm = tc::make_tmap_bf16(lm_head, c.vocab_size, H, lm_tile_rows);
            tmx_xn = tc::make_tmap_bf16(xn.p, R16, H, 16);
            tmx_attn = tc::make_tmap_bf16(attn.p, R16, NQ, 16);
            tmx_act = tc::make_tmap_bf16(act.p, R16, I, 16);
        }
        B2A_CUDA(cudaDeviceSynchronize());
    }

    b2a_tts(int dev, const b2a_llama_config& c, const TensorTable& tt, b2a_snac* sn, const StackSpec& sp = StackSpec())
        : device(dev), cfg(c), snac(sn), spec(sp) {
        check_config();
        const int H = c.hidden_size, I = c.intermediate_size, nq = c.num_attention_heads, nkv = c.num_key_value_heads;
        const int NQ = nq * HD, NKV = nkv * HD;
        if (sp.has_embed) upload_bf16(tt, sp.prefix + "embed_tokens.weight", (int64_t)c.vocab_size * H, embed, 0, (size_t)c.vocab_size * H);
        if (sp.has_head) {
            if (sp.head.empty() && c.tie_word_embeddings) {
                B2A_CHECK(sp.has_embed, B2A_ERR_MODEL_NOT_INITIALIZED, "a tied head needs the embedding");
                lm_head = embed.p;   // embedTokens.asLinear (LlamaTTS.swift:563)
            } else {
                upload_bf16(tt, sp.head.empty() ? std::string("lm_head.weight") : sp.head, (int64_t)c.vocab_size * H, lm_head_w, 0, (size_t)c.vocab_size * H);
                lm_head = lm_head_w.p;
            }
        }
        layers.resize(c.num_hidden_layers);
        std::vector<bf16> tmp;
        for (int l = 0; l < c.num_hidden_layers; ++l) {
            const std::string p = sp.prefix + "layers." + std::to_string(l) + ".";
            if (sp.qk_norm) {
                std::vector<float> qn = tt.f32(p + "self_attn.q_norm.weight", HD), kn = tt.f32(p + "self_attn.k_norm.weight", HD);
                layers[l].qnorm.upload(qn.data(), HD);
                layers[l].knorm.upload(kn.data(), HD);
            }
            LayerW& L = layers[l];
            const size_t qkv_n = (size_t)(NQ + 2 * NKV) * H;
            upload_bf16(tt, p + "self_attn.q_proj.weight", (int64_t)NQ * H, L.wqkv, 0, qkv_n);
            upload_bf16(tt, p + "self_attn.k_proj.weight", (int64_t)NKV * H, L.wqkv, (size_t)NQ * H, qkv_n);
            upload_bf16(tt, p + "self_attn.v_proj.weight", (int64_t)NKV * H, L.wqkv, (size_t)(NQ + NKV) * H, qkv_n);
            upload_bf16(tt, p + "self_attn.o_proj.weight", (int64_t)H * NQ, L.wo, 0, (size_t)H * NQ);
            upload_bf16(tt, p + "mlp.down_proj.weight", (int64_t)H * I, L.wdown, 0, (size_t)H * I);
            // gate / up rows interleaved: row 2n = gate_n, row 2n+1 = up_n
            const b2a_tensor& tg = tt.get(p + "mlp.gate_proj.weight");
            const b2a_tensor& tu = tt.get(p + "mlp.up_proj.weight");
            B2A_CHECK(tg.dtype == tu.dtype && (tg.dtype == B2A_DTYPE_BF16 || tg.dtype == B2A_DTYPE_F32) &&
                          TensorTable::numel(tg) == (int64_t)I * H && TensorTable::numel(tu) == (int64_t)I * H,
                      B2A_ERR_MODEL_NOT_INITIALIZED, "bad gate/up projection: " + p);
            tmp.resize((size_t)2 * I * H);
            for (int n = 0; n < I; ++n) {
                if (tg.dtype == B2A_DTYPE_BF16) {
                    memcpy(&tmp[(size_t)(2 * n) * H], (const bf16*)tg.data + (size_t)n * H, H * sizeof(bf16));
                    memcpy(&tmp[(size_t)(2 * n + 1) * H], (const bf16*)tu.data + (size_t)n * H, H * sizeof(bf16));
                } else {
                    for (int k = 0; k < H; ++k) {
                        tmp[(size_t)(2 * n) * H + k] = __float2bfloat16_rn(((const float*)tg.data)[(size_t)n * H + k]);
                        tmp[(size_t)(2 * n + 1) * H + k] = __float2bfloat16_rn(((const float*)tu.data)[(size_t)n * H + k]);
                    }
                }
            }
            L.wgu.alloc(tmp.size());
            B2A_CUDA(cudaMemcpy(L.wgu.p, tmp.data(), tmp.size() * sizeof(bf16), cudaMemcpyHostToDevice));
            std::vector<float> g1 = tt.f32(p + "input_layernorm.weight", H), g2 = tt.f32(p + "post_attention_layernorm.weight", H);
            L.ln1.upload(g1.data(), H);
            L.ln2.upload(g2.data(), H);
        }
        std::vector<float> gf = tt.f32(sp.prefix + "norm.weight", H);
        final_ln.upload(gf.data(), H);
        alloc_state();
    }

    // random-init weights generated on the device (b2a_tts_create_random)
    b2a_tts(int dev, const b2a_llama_config& c, float std, unsigned long long seed, b2a_snac* sn, const StackSpec& sp = StackSpec())
        : device(dev), cfg(c), snac(sn), spec(sp) {
        check_config();
        const int H = c.hidden_size, I = c.intermediate_size, nq = c.num_attention_heads, nkv = c.num_key_value_heads;
        const int NQ = nq * HD, NKV = nkv * HD;
        unsigned long long sd = seed * 1000003ull + 17;
        auto rnd = [&](DBuf<bf16>& d, size_t n) {
            d.alloc(n);
            random_bf16_kernel<<<148 * 8, 256, 0, stream>>>(d.p, (long long)n, std, sd++);
            count_launch();
        };
        auto ones = [&](DBuf<float>& d, int n) {
            d.alloc(n);
            fill_f32_kernel<<<cdiv(n, 256), 256, 0, stream>>>(d.p, n, 1.0f);
            count_launch();
        };
        if (sp.has_embed) rnd(embed, (size_t)c.vocab_size * H);
        if (sp.has_head) {
            if (sp.head.empty() && c.tie_word_embeddings && sp.has_embed) lm_head = embed.p;
            else { rnd(lm_head_w, (size_t)c.vocab_size * H); lm_head = lm_head_w.p; }
        }
        layers.resize(c.num_hidden_layers);
        for (auto& L : layers) {
            if (sp.qk_norm) { ones(L.qnorm, HD); ones(L.knorm, HD); }
            rnd(L.wqkv, (size_t)(NQ + 2 * NKV) * H);
            rnd(L.wo, (size_t)H * NQ);
            rnd(L.wgu, (size_t)2 * I * H);
            rnd(L.wdown, (size_t)H * I);
            ones(L.ln1, H);
            ones(L.ln2, H);
        }
        ones(final_ln, H);
        B2A_CUDA(cudaStreamSynchronize(stream));
        B2A_CUDA(cudaGetLastError());
        alloc_state();
    }

    template <int NB, int ROWS, int EPI>
    void gemv_launch(const bf16* W, const bf16* xin, float* yout, bf16* actout, int N, int K, int warps, int ksplit,
                     cudaStream_t s) {
        const size_t sm = (size_t)2 * NB * (K / ksplit) * sizeof(bf16);
        dim3 grid(cdiv(N, warps * ROWS), ksplit);
        launch_pdl(gemv_bf16_kernel<NB, ROWS, EPI>, grid, dim3(warps * 32), sm, s, W, xin, yout, actout, N, K);
    }
    template <int NB, int ROWS, int EPI>
    static void gemv_attr() {
        B2A_CUDA(cudaFuncSetAttribute(gemv_bf16_kernel<NB, ROWS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    template <int NB>
    static void gemv_attrs() {
        gemv_attr<NB, 2, GV_F32>(); gemv_attr<NB, 2, GV_SWIGLU>(); gemv_attr<NB, 2, GV_F32_ATOMIC>();
    }
    enum { OP_QKV, OP_O, OP_GU, OP_DOWN, OP_LM };
    template <int NB>
    void gemv_op(int op, const bf16* W, const bf16* xin, float* yout, bf16* actout, int N, int K, cudaStream_t s) {
        const bool split = (K / 8) % 2 == 0 && (size_t)2 * NB * K * sizeof(bf16) > 96 * 1024;
        switch (op) {
            case OP_GU: gemv_launch<NB, 2, GV_SWIGLU>(W, xin, yout, actout, N, K, 8, 1, s); break;
            case OP_DOWN:
            case OP_O:
                if (split) gemv_launch<NB, 2, GV_F32_ATOMIC>(W, xin, yout, actout, N, K, 4, 2, s);
                else gemv_launch<NB, 2, GV_F32>(W, xin, yout, actout, N, K, 4, 1, s);
                break;
            default: gemv_launch<NB, 2, GV_F32>(W, xin, yout, actout, N, K, 8, 1, s); break;
        }
    }
    void gemv_nb(int op, const bf16* W, const bf16* xin, float* yout, bf16* actout, int N, int K, cudaStream_t s) {
        B2A_CHECK((size_t)2 * nb_pad * K * sizeof(bf16) <= 200 * 1024 * ((op == OP_DOWN || op == OP_O) ? 2 : 1),
                  B2A_ERR_INVALID_INPUT, "llama: layer too wide for the shared-memory activation tile");
        switch (nb_pad) {
            case 1: gemv_op<1>(op, W, xin, yout, actout, N, K, s); break;
            case 2: gemv_op<2>(op, W, xin, yout, actout, N, K, s); break;
            case 4: gemv_op<4>(op, W, xin, yout, actout, N, K, s); break;
            default: gemv_op<8>(op, W, xin, yout, actout, N, K, s); break;
        }
    }

    // D[tokens, M] = X[tokens, K] * W[M, K]^T on the tcgen05 path (hi/lo activations, BN = 16)
    // ---- L2 prefetch schedule (B2A_L2PF bit mask; see pf_of): which kernel prefetches which later GEMM's weights
    enum : int { L2_NORM1 = 0, L2_QKV, L2_ATTN, L2_O, L2_NORM2, L2_GU, L2_DOWN };
    int l2pf_mask = -1;
    L2Prefetch pf_of(int site, int l) {
        if (l2pf_mask < 0) {
            const char* e = getenv("B2A_L2PF");
            l2pf_mask = e ? (int)strtol(e, nullptr, 0) : L2PF_DEFAULT;
        }
        if (!use_tc || l < 0 || l >= cfg.num_hidden_layers) return L2Prefetch{nullptr, 0};
        const long long H = cfg.hidden_size, I = cfg.intermediate_size, NQ = (long long)cfg.num_attention_heads * HD,
                        NKV = (long long)cfg.num_key_value_heads * HD;
        const long long b_qkv = (NQ + 2 * NKV) * H * 2, b_o = H * NQ * 2, b_gu = 2 * I * H * 2, b_down = H * I * 2;
        LayerW& L = layers[l];
        const bool last = l + 1 == cfg.num_hidden_layers;
        switch (site) {
            case L2_NORM1: if (l2pf_mask & 1) return L2Prefetch{L.wqkv.p, b_qkv}; break;                 // norm1 -> QKV
            case L2_ATTN:
                if ((l2pf_mask & 2) && (l2pf_mask & 4)) return L2Prefetch{L.wo.p, b_o};                   // attn -> O (GU by the O GEMM)
                if (l2pf_mask & 2) {                                                                      // attn -> GU (first B2A_L2PF_GU_MB MB)
                    static const long long cap = getenv("B2A_L2PF_GU_MB") ? atoll(getenv("B2A_L2PF_GU_MB")) << 20 : (1ll << 40);
                    return L2Prefetch{L.wgu.p, std::min(b_gu, cap)};
                }
                if (l2pf_mask & 4) return L2Prefetch{L.wo.p, b_o};
                break;
            case L2_QKV: if (l2pf_mask & 8) return L2Prefetch{L.wo.p, b_o}; break;                        // QKV GEMM -> O
            case L2_O: if (l2pf_mask & 16) return L2Prefetch{L.wgu.p, b_gu}; break;                       // O GEMM -> GU
            case L2_NORM2: if (l2pf_mask & 32) return L2Prefetch{L.wdown.p, b_down}; break;               // norm2 -> DOWN
            case L2_GU: if (l2pf_mask & 64) return L2Prefetch{L.wdown.p, b_down}; break;                  // GU GEMM -> DOWN
            case L2_DOWN:
                if (l2pf_mask & 128) return last ? L2Prefetch{lm_head, 64ll << 20} : L2Prefetch{layers[l + 1].wqkv.p, b_qkv};   // DOWN -> next QKV
                break;
        }
        return L2Prefetch{nullptr, 0};
    }
    static constexpr int L2PF_DEFAULT = 0;

    void tc_gemm(const CUtensorMap& tmW, const CUtensorMap& tmX, int op, float* yout, bf16* actout, int B, int M, int K,
                 cudaStream_t s, L2Prefetch pf = L2Prefetch{nullptr, 0}, const float* rstd_ss = nullptr) {
        tc::Args a{};
        a.pf_ptr = pf.ptr; a.pf_bytes = pf.bytes;
        a.rstd_ss = rstd_ss; a.rstd_parts = fused_parts; a.rstd_inv_h = 1.0f / (float)cfg.hidden_size; a.rstd_eps = cfg.rms_norm_eps;
        a.out_f32 = yout; a.out_bf16 = actout; a.M = M; a.N = B; a.K = K;
        a.m_tiles = cdiv(M, tc::BM); a.k_blocks = K / tc::BK;
        a.stages = 6;   // 6 x 18 KB ring: two GEMM CTAs (this kernel's and the next kernel's prefetching one) fit per SM
        { static const int e_all = getenv("B2A_STAGES") ? atoi(getenv("B2A_STAGES")) : 0, e_gu = getenv("B2A_STAGES_GU") ? atoi(getenv("B2A_STAGES_GU")) : 0;
          if (e_all > 0) a.stages = e_all;
          if (op == OP_GU && e_gu > 0) a.stages = e_gu; }
        a.hilo = 1;
        int ctas = num_sms;
        if (op == OP_GU) {
            a.ldo = M / 2; a.epi_full = tc::EPI_SWIGLU; a.epi_partial = -1; a.lo_rows = LO_ROW;
            a.tile_rows = gu_tile_rows; a.m_tiles = cdiv(M, gu_tile_rows);      // tmW is tm_gu_dec[layer]
            ctas = std::min(num_sms, a.m_tiles);
        } else if (op == OP_LM) {
            a.ldo = M; a.epi_full = tc::EPI_STORE; a.epi_partial = -1; a.lo_rows = 0;
            a.tile_rows = head_rows_now > 0 ? head_rows_now : lm_tile_rows; a.m_tiles = cdiv(M, a.tile_rows);
            ctas = std::min(num_sms, a.m_tiles);
        } else {   // qkv / o / down: stream-K, partial tiles accumulate into the zeroed fp32 output
            a.ldo = M; a.epi_full = tc::EPI_STORE; a.epi_partial = tc::EPI_ATOMIC; a.lo_rows = 0;
            ctas = (int)std::min<long long>(num_sms, (long long)a.m_tiles * a.k_blocks);
        }
        tc::launch<16>(tmW, tmX, a, ctas, 1, s);
    }

    void gemm(int op, int layer, int B, cudaStream_t s) {
        const int H = cfg.hidden_size, I = cfg.intermediate_size, NQ = cfg.num_attention_heads * HD,
                  NKV = cfg.num_key_value_heads * HD;
        LayerW* L = layer >= 0 ? &layers[layer] : nullptr;
        switch (op) {
            case OP_QKV:
                if (use_tc) tc_gemm(tm_qkv[layer], tmx_xn, op, qkv.p, nullptr, B, NQ + 2 * NKV, H, s, pf_of(L2_QKV, layer));
                else gemv_nb(op, L->wqkv.p, xn.p, qkv.p, nullptr, NQ + 2 * NKV, H, s);
                break;
            case OP_O:
                if (use_tc) tc_gemm(tm_o[layer], tmx_attn, op, y.p, nullptr, B, H, NQ, s, pf_of(L2_O, layer));
                else gemv_nb(op, L->wo.p, attn.p, y.p, nullptr, H, NQ, s);
                break;
            case OP_GU:
                if (use_tc) tc_gemm(tm_gu_dec[layer], tmx_xn, op, nullptr, act.p, B, 2 * I, H, s, pf_of(L2_GU, layer));
                else gemv_nb(op, L->wgu.p, xn.p, nullptr, act.p, 2 * I, H, s);
                break;
            case OP_DOWN:
                if (use_tc) tc_gemm(tm_down[layer], tmx_act, op, y.p, nullptr, B, H, I, s, pf_of(L2_DOWN, layer));
                else gemv_nb(op, L->wdown.p, act.p, y.p, nullptr, H, I, s);
                break;
            default:
                if (use_tc) tc_gemm(tm_lm, tmx_xn, op, logits.p, nullptr, B, cfg.vocab_size, H, s);
                else gemv_nb(op, lm_head, xn.p, logits.p, nullptr, cfg.vocab_size, H, s);
                break;
        }
    }

    // timing ablation only (B2A_SKIP=norm|attn|gemm|qkv|o|gu|down, comma separated): skips launches, results invalid
    static bool skip(const char* what) {
        const char* e = getenv("B2A_SKIP");
        return e && strstr(e, what) != nullptr;
    }

    // o_proj / down_proj as a cluster split-K GEMM with the residual add and the next norm fused into the leader's epilogue
    void splitk_gemm(const CUtensorMap& tmW, const CUtensorMap& tmX, int M, int K, const float* gain, float* ss, int B, cudaStream_t s) {
        tc::SplitArgs a{};
        a.M = M; a.N = B; a.K = K; a.k_blocks = K / tc::BK; a.stages = 5;
        { static const int e_sk = getenv("B2A_STAGES_SK") ? atoi(getenv("B2A_STAGES_SK")) : 0; if (e_sk > 0) a.stages = e_sk; }
        a.h = x.p; a.gain = gain; a.xn = xn.p; a.ss = ss;
        a.rstd_ss = nullptr; a.rstd_parts = 0; a.rstd_inv_h = 0.f; a.rstd_eps = 0.f;
        tc::launch_splitk(tmW, tmX, a, cdiv(M, tc::BM), std::max(1, std::min(fused_cluster, a.k_blocks)), s);
    }
    // The fused-norm step: embed -> raw norm -> L x [qkv gemm (rstd in the epilogue) -> attention (clears q|k|v) -> o split-K (+ residual,
    // norm 2) -> gate/up gemm (rstd, SwiGLU) -> down split-K (+ residual, next layer's norm 1 / the final norm)].  Leaves the residual
    // stream in x, xn = hi/lo of x * final_norm_gain and its sums of squares in ss_b: the lm head GEMM applies rstd itself.
    void run_layers_fused(int B, cudaStream_t s) {
        const int H = cfg.hidden_size, I = cfg.intermediate_size, nq = cfg.num_attention_heads, nkv = cfg.num_key_value_heads;
        const int NQ = nq * HD, NKV = nkv * HD, L = cfg.num_hidden_layers;
        if (x_ext) launch_pdl(ext_embed_kernel, dim3(B), dim3(256), 0, s, x_ext, x.p, y.p, H);
        else launch_pdl(embed_kernel, dim3(B), dim3(256), 0, s, tokens.p, embed.p, x.p, y.p, H, cfg.vocab_size);
        launch_pdl(add_rmsnorm_kernel, dim3(B), dim3(RN_THREADS), 0, s, x.p, (float*)nullptr, layers[0].ln1.p, xn.p, H, cfg.rms_norm_eps,
                   (float*)nullptr, (float*)nullptr, 0, LO_ROW, L2Prefetch{nullptr, 0}, (float*)nullptr, ss_b.p, fused_parts);
        const size_t kv_layer = (size_t)cfg.max_batch * nkv * cfg.max_context * HD;
        for (int l = 0; l < L; ++l) {
            LayerW& Lw = layers[l];
            tc_gemm(tm_qkv[l], tmx_xn, OP_QKV, qkv.p, nullptr, B, NQ + 2 * NKV, H, s, pf_of(L2_QKV, l), ss_b.p);
            AttnArgs aa{qkv.p, pos.p, freqs.p, kcache.p + l * kv_layer, vcache.p + l * kv_layer, attn.p, part_o.p, part_ml.p,
                        at_counters.p, nq, nkv, cfg.max_context, at_splits, 1.0f / sqrtf((float)HD), pf_of(L2_ATTN, l),
                        spec.qk_norm ? Lw.qnorm.p : nullptr, spec.qk_norm ? Lw.knorm.p : nullptr, cfg.rms_norm_eps, 1};
            attn_launch(aa, B, s);
            splitk_gemm(tm_o[l], tmx_attn, H, NQ, Lw.ln2.p, ss_a.p, B, s);
            tc_gemm(tm_gu_dec[l], tmx_xn, OP_GU, nullptr, act.p, B, 2 * I, H, s, pf_of(L2_GU, l), ss_a.p);
            splitk_gemm(tm_down[l], tmx_act, H, I, l + 1 < L ? layers[l + 1].ln1.p : final_ln.p, ss_b.p, B, s);
        }
    }
    int launches_layers_fused() const { return 2 + cfg.num_hidden_layers * 5; }

    // embed(tokens) -> all layers; leaves the residual stream in x and the last MLP output in y
    void run_layers(int B, cudaStream_t s) {
        if (fused && !trace_on) { run_layers_fused(B, s); return; }
        const int H = cfg.hidden_size, nq = cfg.num_attention_heads, nkv = cfg.num_key_value_heads;
        const int QKV_N = (nq + 2 * nkv) * HD, G = nq / nkv;
        if (x_ext) launch_pdl(ext_embed_kernel, dim3(B), dim3(256), 0, s, x_ext, x.p, y.p, H);
        else launch_pdl(embed_kernel, dim3(B), dim3(256), 0, s, tokens.p, embed.p, x.p, y.p, H, cfg.vocab_size);
        const size_t kv_layer = (size_t)cfg.max_batch * nkv * cfg.max_context * HD;
        for (int l = 0; l < cfg.num_hidden_layers; ++l) {
            LayerW& L = layers[l];
            if (!skip("norm"))
            launch_pdl(add_rmsnorm_kernel, dim3(B), dim3(RN_THREADS), 0, s, x.p, l == 0 ? (float*)nullptr : y.p, L.ln1.p, xn.p, H,
                       cfg.rms_norm_eps, trace_on ? trace.p + (size_t)(2 * l) * 8 * H : (float*)nullptr, (float*)nullptr, 0, LO_ROW,
                       pf_of(L2_NORM1, l), (float*)nullptr, (float*)nullptr, 0);
            if (!skip("gemm") && !skip("qkv")) gemm(OP_QKV, l, B, s);
            AttnArgs aa{qkv.p, pos.p, freqs.p, kcache.p + l * kv_layer, vcache.p + l * kv_layer, attn.p, part_o.p, part_ml.p,
                        at_counters.p, nq, nkv, cfg.max_context, at_splits, 1.0f / sqrtf((float)HD), pf_of(L2_ATTN, l),
                        spec.qk_norm ? L.qnorm.p : nullptr, spec.qk_norm ? L.knorm.p : nullptr, cfg.rms_norm_eps, 0};
            if (!skip("attn")) attn_launch(aa, B, s);
            if (!skip("gemm") && !skip("o_proj")) gemm(OP_O, l, B, s);
            // also zeroes this row of q|k|v so the next layer's stream-K QKV GEMM can accumulate into it
            if (!skip("norm"))
            launch_pdl(add_rmsnorm_kernel, dim3(B), dim3(RN_THREADS), 0, s, x.p, y.p, L.ln2.p, xn.p, H, cfg.rms_norm_eps,
                       trace_on ? trace.p + (size_t)(2 * l + 1) * 8 * H : (float*)nullptr, qkv.p, QKV_N, LO_ROW, pf_of(L2_NORM2, l), (float*)nullptr, (float*)nullptr, 0);
            if (!skip("gemm") && !skip("gate")) gemm(OP_GU, l, B, s);
            if (!skip("gemm") && !skip("down")) gemm(OP_DOWN, l, B, s);
        }
        (void)G;
    }

    void run_final_norm(int B, cudaStream_t s) {
        if (fused && !trace_on) {   // x, xn (un-normalised) and ss_b are already final: only row N1 needs the fp32 normalised hidden
            if (normed_out)
                launch_pdl(finalize_norm_kernel, dim3(B), dim3(256), 0, s, (const float*)x.p, (const float*)final_ln.p, (const float*)ss_b.p, fused_parts,
                           normed_out, cfg.hidden_size, cfg.rms_norm_eps);
            return;
        }
        launch_pdl(add_rmsnorm_kernel, dim3(B), dim3(RN_THREADS), 0, s, x.p, y.p, final_ln.p, xn.p, cfg.hidden_size, cfg.rms_norm_eps,
                   trace_on ? trace.p + (size_t)(2 * cfg.num_hidden_layers) * 8 * cfg.hidden_size : (float*)nullptr, (float*)nullptr, 0, LO_ROW,
                   L2Prefetch{nullptr, 0}, normed_out, (float*)nullptr, 0);
    }
    void run_lm_head(int B, cudaStream_t s) {
        run_final_norm(B, s);
        if (fused && !trace_on) tc_gemm(tm_lm, tmx_xn, OP_LM, logits.p, nullptr, B, cfg.vocab_size, cfg.hidden_size, s, L2Prefetch{nullptr, 0}, ss_b.p);
        else gemm(OP_LM, -1, B, s);
    }
    // after prefill_batched's gather_last (x, y hold the last position un-added): always the stand-alone norm + plain GEMM
    void run_lm_head_after_prefill(int B, cudaStream_t s) {
        launch_pdl(add_rmsnorm_kernel, dim3(B), dim3(RN_THREADS), 0, s, x.p, y.p, final_ln.p, xn.p, cfg.hidden_size, cfg.rms_norm_eps,
                   (float*)nullptr, (float*)nullptr, 0, LO_ROW, L2Prefetch{nullptr, 0}, (float*)nullptr, (float*)nullptr, 0);
        gemm(OP_LM, -1, B, s);
    }
    // a head the caller owns (row N1: the code predictor's 15 lm heads): logits_out[b, :M] = W[M, H] * normed hidden
    void run_head(const CUtensorMap& tmW, const bf16* W, int M, float* logits_out, int B, cudaStream_t s, int tile_rows = 0) {
        head_rows_now = tile_rows;              // tmW's box rows (0: this stack's own lm_tile_rows)
        if (use_tc) tc_gemm(tmW, tmx_xn, OP_LM, logits_out, nullptr, B, M, cfg.hidden_size, s, L2Prefetch{nullptr, 0},
                            (fused && !trace_on) ? ss_b.p : nullptr);
        else gemv_nb(OP_LM, W, xn.p, logits_out, nullptr, M, cfg.hidden_size, s);
        head_rows_now = 0;
    }
    // logits are [8, V] row-major.

    template <int G>
    static void pattn_attr() {
        B2A_CUDA(cudaFuncSetAttribute(prefill_attn_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    }
    size_t pattn_smem(int L) const {
        const int G = cfg.num_attention_heads / cfg.num_key_value_heads;
        return ((size_t)2 * L * HD + (size_t)G * PA_QT * HD) * sizeof(float);
    }
    bool can_batch_prefill(int L) const {
        return use_tc && use_batched_prefill && spec.has_embed && !spec.qk_norm && L >= 2 && L <= PA_MAXL && pattn_smem(L) <= 220 * 1024;
    }

    // D[T, M] = X[T, K] W^T for all prompt tokens: 128-column tiles (64 tokens as hi/lo), CTAs own whole tiles
    void pf_gemm(const CUtensorMap& tmW, const CUtensorMap& tmX, int epi, float* yout, bf16* actout, int T, int M, int K,
                 cudaStream_t s) {
        tc::Args a{};
        a.out_f32 = yout; a.out_bf16 = actout; a.M = M; a.N = T; a.K = K;
        a.m_tiles = cdiv(M, tc::BM); a.k_blocks = K / tc::BK;
        a.stages = 6; a.hilo = 1; a.epi_full = epi; a.epi_partial = -1;
        a.ldo = epi == tc::EPI_SWIGLU ? M / 2 : M;
        a.lo_rows = epi == tc::EPI_SWIGLU ? PF_HALF : 0;
        const int n_tiles = cdiv(T, PF_HALF);
        const int ctas = std::max(1, std::min(a.m_tiles, num_sms / n_tiles));
        tc::launch<128>(tmW, tmX, a, ctas, n_tiles, s);
    }

    // Prompt pass over all B*L tokens; leaves K/V for positions 0..L-1 in the cache and the last position's
    // residual stream in the decode buffers (x, y), pos[b] = L-1: the caller then runs lm head + sampler.
    void prefill_batched(int B, int L, cudaStream_t s) {
        const int H = cfg.hidden_size, I = cfg.intermediate_size, nq = cfg.num_attention_heads, nkv = cfg.num_key_value_heads;
        const int NQ = nq * HD, QKV_N = (nq + 2 * nkv) * HD, T = B * L, G = nq / nkv;
        const int n_tiles = cdiv(T, PF_HALF), Tp = n_tiles * PF_HALF;
        if (Tp > pf_tokens_cap) {
            xp.alloc((size_t)Tp * H); yp.alloc((size_t)Tp * H); qkvp.alloc((size_t)Tp * QKV_N);
            xnp.alloc((size_t)2 * Tp * H); attnp.alloc((size_t)2 * Tp * NQ); actp.alloc((size_t)2 * Tp * I);
            pf_tokens_cap = Tp;
            tmp_xn = tc::make_tmap_bf16(xnp.p, 2 * Tp, H, 128);
            tmp_attn = tc::make_tmap_bf16(attnp.p, 2 * Tp, NQ, 128);
            tmp_act = tc::make_tmap_bf16(actp.p, 2 * Tp, I, 128);
            pattn_attr<1>(); pattn_attr<2>(); pattn_attr<3>(); pattn_attr<4>(); pattn_attr<6>(); pattn_attr<8>();
        }
        rope_tab.alloc((size_t)PA_MAXL * (HD / 2));
        // padding tokens of the last tile must read as zero
        B2A_CUDA(cudaMemsetAsync(xnp.p, 0, (size_t)2 * pf_tokens_cap * H * sizeof(bf16), s));
        B2A_CUDA(cudaMemsetAsync(attnp.p, 0, (size_t)2 * pf_tokens_cap * NQ * sizeof(bf16), s));
        B2A_CUDA(cudaMemsetAsync(actp.p, 0, (size_t)2 * pf_tokens_cap * I * sizeof(bf16), s));
        embed_rows_kernel<<<T, 256, 0, s>>>(ids.p, embed.p, xp.p, H, cfg.vocab_size);
        rope_table_kernel<<<cdiv(L * (HD / 2), 256), 256, 0, s>>>(freqs.p, rope_tab.p, L);
        count_launch(2);
        const size_t kv_layer = (size_t)cfg.max_batch * nkv * cfg.max_context * HD;
        for (int l = 0; l < cfg.num_hidden_layers; ++l) {
            LayerW& Lw = layers[l];
            launch_pdl(add_rmsnorm_kernel, dim3(T), dim3(RN_THREADS), 0, s, xp.p, l == 0 ? (float*)nullptr : yp.p, Lw.ln1.p, xnp.p, H,
                       cfg.rms_norm_eps, (float*)nullptr, (float*)nullptr, 0, PF_HALF, L2Prefetch{nullptr, 0}, (float*)nullptr, (float*)nullptr, 0);
            pf_gemm(tm_qkv[l], tmp_xn, tc::EPI_STORE, qkvp.p, nullptr, T, QKV_N, H, s);
            PrefillAttnArgs pa{qkvp.p, rope_tab.p, kcache.p + l * kv_layer, vcache.p + l * kv_layer, attnp.p, nq, nkv,
                               cfg.max_context, L, 1.0f / sqrtf((float)HD)};
            const dim3 grid(nkv, B, cdiv(L, PA_QT));
            const size_t sm = pattn_smem(L);
            switch (G) {
                case 1: prefill_attn_kernel<1><<<grid, PA_THREADS, sm, s>>>(pa); break;
                case 2: prefill_attn_kernel<2><<<grid, PA_THREADS, sm, s>>>(pa); break;
                case 3: prefill_attn_kernel<3><<<grid, PA_THREADS, sm, s>>>(pa); break;
                case 4: prefill_attn_kernel<4><<<grid, PA_THREADS, sm, s>>>(pa); break;
                case 6: prefill_attn_kernel<6><<<grid, PA_THREADS, sm, s>>>(pa); break;
                default: prefill_attn_kernel<8><<<grid, PA_THREADS, sm, s>>>(pa); break;
            }
            count_launch();
            pf_gemm(tm_o[l], tmp_attn, tc::EPI_STORE, yp.p, nullptr, T, H, NQ, s);
            launch_pdl(add_rmsnorm_kernel, dim3(T), dim3(RN_THREADS), 0, s, xp.p, yp.p, Lw.ln2.p, xnp.p, H, cfg.rms_norm_eps,
                       (float*)nullptr, (float*)nullptr, 0, PF_HALF, L2Prefetch{nullptr, 0}, (float*)nullptr, (float*)nullptr, 0);
            pf_gemm(tm_gu[l], tmp_xn, tc::EPI_SWIGLU, nullptr, actp.p, T, 2 * I, H, s);
            pf_gemm(tm_down[l], tmp_act, tc::EPI_STORE, yp.p, nullptr, T, H, I, s);
        }
        gather_last_kernel<<<B, 256, 0, s>>>(xp.p, yp.p, x.p, y.p, pos.p, L, H);
        count_launch();
        B2A_CUDA(cudaGetLastError());
    }

    void set_batch(int B) {
        nb_pad = B <= 1 ? 1 : B <= 2 ? 2 : B <= 4 ? 4 : 8;
    }

    void drop_graphs() {
        if (g_step) { cudaGraphExecDestroy(g_step); g_step = nullptr; }
        if (g_prefill) { cudaGraphExecDestroy(g_prefill); g_prefill = nullptr; }
    }

    static bool same_args(const SampleArgs& a, const SampleArgs& b) {
        return a.V == b.V && a.R == b.R && a.max_tokens == b.max_tokens && a.temperature == b.temperature &&
               a.top_p == b.top_p && a.rep_penalty == b.rep_penalty && a.seed == b.seed && a.mask_eos == b.mask_eos &&
               a.out_tokens == b.out_tokens;
    }

    void capture(int B, const SampleArgs& sa, int L) {
        if (g_step && g_nb == B && same_args(sa, g_args) && g_L == L) return;
        drop_graphs();
        cudaGraph_t g;
        B2A_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        run_layers(B, stream);
        run_lm_head(B, stream);
        launch_pdl(sample_kernel, dim3(B * SM_CLUSTER), dim3(SM_THREADS), 0, stream, sa);
        B2A_CUDA(cudaStreamEndCapture(stream, &g));
        B2A_CUDA(cudaGraphInstantiate(&g_step, g, 0));
        cudaGraphDestroy(g);
        B2A_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        run_layers(B, stream);
        launch_pdl(prefill_advance_kernel, dim3(1), dim3(32), 0, stream, ids.p, L, tokens.p, pos.p, B);
        B2A_CUDA(cudaStreamEndCapture(stream, &g));
        B2A_CUDA(cudaGraphInstantiate(&g_prefill, g, 0));
        cudaGraphDestroy(g);
        g_nb = B; g_args = sa; g_L = L;
        launches_step = fused ? launches_layers_fused() + 1 + 1 : 1 + cfg.num_hidden_layers * 7 + 2 + 1;
        launches_prefill = fused ? launches_layers_fused() + 1 : 1 + cfg.num_hidden_layers * 7 + 1;
    }
    int g_L = 0, launches_step = 0, launches_prefill = 0;
};

// ------------------------------------------------------------------------------------------------
// host-side token plumbing (ints; the reference does these on the host too)
// ------------------------------------------------------------------------------------------------
static std::vector<int> parse_row(const int* row, int n, int crop_after) {
    // LlamaTTS.swift:400-431 for one row: crop, drop 128258, trim to a multiple of 7, subtract 128266
    std::vector<int> r;
    for (int j = crop_after + 1; j < n; ++j)
        if (row[j] != TOK_END_OF_SPEECH) r.push_back(row[j]);
    r.resize((r.size() / 7) * 7);
    for (auto& t : r) t -= TOK_AUDIO_OFFSET;
    return r;
}

static void deinterleave(const int* cl, int n, std::vector<int>& l1, std::vector<int>& l2, std::vector<int>& l3) {
    // llamaDecodeAudioFromCodes, LlamaTTS.swift:46-58
    const int groups = (n + 1) / 7;
    for (int i = 0; i < groups; ++i) {
        const int b = 7 * i;
        l1.push_back(cl[b]);
        l2.push_back(cl[b + 1] - 4096);
        l3.push_back(cl[b + 2] - 2 * 4096);
        l3.push_back(cl[b + 3] - 3 * 4096);
        l2.push_back(cl[b + 4] - 4 * 4096);
        l3.push_back(cl[b + 5] - 5 * 4096);
        l3.push_back(cl[b + 6] - 6 * 4096);
    }
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// shared body of b2a_tts_generate / _dev.  ids_on_device: input ids pointer is a device pointer.
// chunked audio emission during generation (row N2): every `frames_per_chunk` new 7-token frames of a row are decoded with
// `left_context` already-emitted frames in front (the codec is convolutional: the context absorbs the left edge) and handed to
// on_audio; the samples of the context frames are dropped.
struct StreamSpec {
    int frames_per_chunk = 0;     // 0: no streaming
    int left_context = 0;
    b2a_audio_cb on_audio = nullptr;
};

static void tts_generate_impl(b2a_tts* h, const int32_t* input_ids, bool ids_on_device, int32_t B, int32_t L,
                              const b2a_gen_params* gp, int32_t* tokens_out, int32_t* n_tokens_out, float* wave_out,
                              bool wave_on_device, int64_t wave_cap, int64_t* wave_len, b2a_gen_info* info,
                              b2a_token_cb on_token, void* user, const StreamSpec& ss = StreamSpec()) {
    B2A_CHECK(h && input_ids && gp, B2A_ERR_INVALID_INPUT, "tts generate: null argument");
    B2A_CHECK(B >= 1 && B <= h->cfg.max_batch, B2A_ERR_INVALID_INPUT, "tts generate: batch exceeds max_batch");
    B2A_CHECK(L >= 1, B2A_ERR_INVALID_INPUT, "tts generate: empty prompt");
    B2A_CHECK(gp->max_tokens >= 1, B2A_ERR_INVALID_INPUT, "tts generate: max_tokens must be positive");
    B2A_CHECK(L + gp->max_tokens <= h->cfg.max_context, B2A_ERR_INVALID_INPUT, "tts generate: prompt + max_tokens exceeds max_context");
    B2A_CHECK(gp->repetition_context_size >= 0 && gp->repetition_context_size <= 64, B2A_ERR_INVALID_INPUT,
              "tts generate: repetition_context_size must be in 0..64");
    B2A_CHECK(gp->temperature >= 0.f && gp->top_p > 0.f, B2A_ERR_INVALID_INPUT, "tts generate: bad sampling parameters");
    if (wave_out) B2A_CHECK(h->snac, B2A_ERR_MODEL_NOT_INITIALIZED, "SNAC model not loaded");   // LlamaTTS.swift:672-674
    B2A_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    h->cancel.store(0);
    h->set_batch(B);
    if (h->trace_on) { h->trace_on = false; h->drop_graphs(); }
    const int MT = gp->max_tokens, R = std::max(1, gp->repetition_context_size);
    h->ids.alloc((size_t)B * L);
    h->out_tokens.alloc((size_t)B * MT);
    h->recent.alloc((size_t)8 * R);
    B2A_CUDA(cudaMemcpyAsync(h->ids.p, input_ids, (size_t)B * L * sizeof(int),
                             ids_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    SampleArgs sa{};
    sa.logits = h->logits.p; sa.probs = h->probs.p; sa.tokens = h->tokens.p; sa.pos = h->pos.p; sa.recent = h->recent.p;
    sa.recent_n = h->recent_n.p; sa.out_tokens = h->out_tokens.p; sa.n_gen = h->n_gen.p; sa.done = h->done.p;
    sa.n_active = h->n_active.p; sa.forced = nullptr; sa.V = h->cfg.vocab_size; sa.R = gp->repetition_context_size > 0 ? R : 0;
    sa.max_tokens = MT; sa.temperature = gp->temperature; sa.top_p = gp->top_p;
    sa.rep_penalty = gp->repetition_context_size > 0 ? gp->repetition_penalty : 1.0f;
    sa.seed = gp->seed; sa.mask_eos = h->bench_mask_eos;
    if (sa.R == 0) { sa.R = 1; }
    h->capture(B, sa, L);

    const double t0 = now_s();
    init_rows_kernel<<<1, 32, 0, s>>>(h->ids.p, L, B, gp->repetition_context_size > 0 ? R : 0, h->tokens.p, h->pos.p, h->recent.p,
                                      h->recent_n.p, h->n_gen.p, h->done.p, h->n_active.p, 0);
    count_launch();
    // prefill.  Batched: every prompt token through each layer at once (tcgen05 GEMMs, 64 tokens per tile), then
    // lm head + sampler on the last position.  Fallback (B2A_PREFILL=step, L > 128, SIMT mode): replay the decode
    // step per position -- positions 0..L-2 need no logits, position L-1 runs the full step.
    int steps = 0;
    if (h->can_batch_prefill(L)) {
        h->prefill_batched(B, L, s);
        h->run_lm_head_after_prefill(B, s);
        launch_pdl(sample_kernel, dim3(B * SM_CLUSTER), dim3(SM_THREADS), 0, s, sa);
        steps = 1;
    } else {
        for (int p = 0; p < L - 1; ++p) {
            B2A_CUDA(cudaGraphLaunch(h->g_prefill, s));
            count_launch(h->launches_prefill);
        }
    }
    B2A_CUDA(cudaStreamSynchronize(s));
    const double t1 = now_s();
    bool cancelled = false;
    int streamed = 0;
    std::vector<int> h_tok;
    // ---- streaming audio emission state
    const bool streaming = ss.frames_per_chunk > 0 && ss.on_audio;
    if (streaming) B2A_CHECK(h->snac && !ids_on_device, B2A_ERR_MODEL_NOT_INITIALIZED, "SNAC model not loaded");
    std::vector<int> emitted(B, 0), s_prompt, s_tok, s_ng(B);
    std::vector<float> s_wave;
    double codec_stream_t = 0;
    if (streaming) { s_prompt.assign(input_ids, input_ids + (size_t)B * L); s_tok.resize((size_t)B * MT); }
    auto emit_audio = [&](bool final_call) {
        const double c0 = now_s();
        B2A_CUDA(cudaMemcpy(s_ng.data(), h->n_gen.p, B * sizeof(int), cudaMemcpyDeviceToHost));
        B2A_CUDA(cudaMemcpy(s_tok.data(), h->out_tokens.p, (size_t)B * MT * sizeof(int), cudaMemcpyDeviceToHost));
        const int64_t hop = b2a_snac_hop_length(h->snac);
        for (int b = 0; b < B; ++b) {
            const int ng = std::min(s_ng[b], MT);
            std::vector<int> all(s_prompt.begin() + (size_t)b * L, s_prompt.begin() + (size_t)(b + 1) * L);
            all.insert(all.end(), s_tok.begin() + (size_t)b * MT, s_tok.begin() + (size_t)b * MT + ng);
            int last = -1;
            for (int j = 0; j < (int)all.size(); ++j) if (all[j] == TOK_START_OF_SPEECH) last = j;
            std::vector<int> cl = parse_row(all.data(), (int)all.size(), last);
            if (h->bench_wrap_codes)
                for (size_t i = 0; i < cl.size(); ++i) cl[i] = ((cl[i] % 4096) + 4096) % 4096 + 4096 * (int)(i % 7);
            const int total = (int)cl.size() / 7;
            while (total - emitted[b] >= ss.frames_per_chunk || (final_call && total > emitted[b])) {
                const int f1 = final_call ? total : emitted[b] + ss.frames_per_chunk;
                const int f0 = std::max(0, emitted[b] - ss.left_context);
                std::vector<int> l1, l2, l3;
                deinterleave(cl.data() + (size_t)f0 * 7, (f1 - f0) * 7, l1, l2, l3);
                const int F = f1 - f0;
                h->d_codes[0].upload(l1.data(), l1.size(), s);
                h->d_codes[1].upload(l2.data(), l2.size(), s);
                h->d_codes[2].upload(l3.data(), l3.size(), s);
                const int64_t T = 4ll * F, wl = T * hop;
                h->d_wave.alloc((size_t)wl);
                const int* dc[3] = {h->d_codes[0].p, h->d_codes[1].p, h->d_codes[2].p};
                B2A_CUDA(cudaStreamSynchronize(s));
                const int32_t st = b2a_snac_decode_dev(h->snac, dc, 1, T, nullptr, 0, gp->seed + (uint64_t)b, h->d_wave.p, s);
                B2A_CHECK(st == B2A_OK, B2A_ERR_AUDIO_DECODING_FAILED, std::string("SNAC decode failed: ") + b2a_last_error());
                const int64_t skip = (int64_t)(emitted[b] - f0) * 4 * hop, n_new = wl - skip;
                s_wave.resize((size_t)n_new);
                B2A_CUDA(cudaMemcpyAsync(s_wave.data(), h->d_wave.p + skip, (size_t)n_new * sizeof(float), cudaMemcpyDeviceToHost, s));
                B2A_CUDA(cudaStreamSynchronize(s));
                emitted[b] = f1;
                ss.on_audio(user, b, s_wave.data(), n_new, (final_call && f1 == total) ? 1 : 0);
            }
        }
        codec_stream_t += now_s() - c0;
    };
    auto stream_tokens = [&]() {   // .token events (LlamaTTS.swift:862), row-major per step
        h_tok.resize((size_t)B * MT);
        std::vector<int> ng(B);
        B2A_CUDA(cudaMemcpy(ng.data(), h->n_gen.p, B * sizeof(int), cudaMemcpyDeviceToHost));
        B2A_CUDA(cudaMemcpy(h_tok.data(), h->out_tokens.p, (size_t)B * MT * sizeof(int), cudaMemcpyDeviceToHost));
        for (int b = 0; b < B; ++b)
            if (ng[b] > streamed) on_token(user, b, streamed, h_tok[(size_t)b * MT + streamed]);
        ++streamed;
    };
    if (steps == 1) {
        if (on_token) stream_tokens();
        B2A_CUDA(cudaMemcpyAsync(h->h_flag.p, h->n_active.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        B2A_CUDA(cudaStreamSynchronize(s));
    }
    // Without per-token callbacks the host keeps ONE burst of graph launches queued ahead of the burst whose "rows still active"
    // flag it is waiting for, so the GPU never idles while the host polls (round 1 synchronised after every burst: 5 % of the
    // loop).  A burst launched after every row has finished only replays steps whose tokens are not recorded; steps never exceed
    // max_tokens, so the KV cache cannot overflow.
    const bool pipelined = !on_token && !streaming;
    if (pipelined && !h->ev_poll[0]) { B2A_CUDA(cudaEventCreateWithFlags(&h->ev_poll[0], cudaEventDisableTiming)); B2A_CUDA(cudaEventCreateWithFlags(&h->ev_poll[1], cudaEventDisableTiming)); }
    int slot = 0, pending = -1;
    while (steps < MT && !(steps == 1 && h->h_flag.p[0] <= 0)) {
        const int burst = on_token ? 1 : std::min(streaming ? 7 : 16, MT - steps);
        for (int i = 0; i < burst; ++i) {
            B2A_CUDA(cudaGraphLaunch(h->g_step, s));
            count_launch(h->launches_step);
        }
        steps += burst;
        if (pipelined) {
            B2A_CUDA(cudaMemcpyAsync(h->h_flag.p + 1 + slot, h->n_active.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            B2A_CUDA(cudaEventRecord(h->ev_poll[slot], s));
            if (pending >= 0) {
                B2A_CUDA(cudaEventSynchronize(h->ev_poll[pending]));
                if (h->cancel.load()) { cancelled = true; break; }
                if (h->h_flag.p[1 + pending] <= 0) break;
            }
            pending = slot; slot ^= 1;
            continue;
        }
        B2A_CUDA(cudaMemcpyAsync(h->h_flag.p, h->n_active.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        B2A_CUDA(cudaStreamSynchronize(s));
        if (on_token) stream_tokens();
        if (streaming) emit_audio(false);
        if (h->cancel.load()) { cancelled = true; break; }
        if (h->h_flag.p[0] <= 0) break;
    }
    if (pipelined) { B2A_CUDA(cudaStreamSynchronize(s)); if (h->cancel.load()) cancelled = true; }
    if (streaming && !cancelled) emit_audio(true);
    const double t2 = now_s();
    B2A_CHECK(!cancelled, B2A_ERR_CANCELLED, "generation cancelled");

    std::vector<int> ng(B), toks((size_t)B * MT), prompt((size_t)B * L);
    B2A_CUDA(cudaMemcpyAsync(ng.data(), h->n_gen.p, B * sizeof(int), cudaMemcpyDeviceToHost, s));
    B2A_CUDA(cudaMemcpyAsync(toks.data(), h->out_tokens.p, (size_t)B * MT * sizeof(int), cudaMemcpyDeviceToHost, s));
    B2A_CUDA(cudaMemcpyAsync(prompt.data(), h->ids.p, (size_t)B * L * sizeof(int), cudaMemcpyDeviceToHost, s));
    B2A_CUDA(cudaStreamSynchronize(s));
    int total_gen = 0;
    for (int b = 0; b < B; ++b) {
        ng[b] = std::min(ng[b], MT);
        total_gen += ng[b];
        if (n_tokens_out) n_tokens_out[b] = ng[b];
        if (tokens_out) memcpy(tokens_out + (size_t)b * MT, toks.data() + (size_t)b * MT, ng[b] * sizeof(int));
    }

    double codec_t = 0;
    if (wave_out) {
        const double c0 = now_s();
        // per row: generatedTokens = prompt + generated (LlamaTTS.swift:705-707,738) -> parseOutput -> frames
        std::vector<std::vector<int>> l1(B), l2(B), l3(B);
        std::vector<int> frames(B, 0);
        bool any = false;
        for (int b = 0; b < B; ++b) {
            std::vector<int> all(prompt.begin() + (size_t)b * L, prompt.begin() + (size_t)(b + 1) * L);
            all.insert(all.end(), toks.begin() + (size_t)b * MT, toks.begin() + (size_t)b * MT + ng[b]);
            int last = -1;
            for (int j = 0; j < (int)all.size(); ++j) if (all[j] == TOK_START_OF_SPEECH) last = j;
            std::vector<int> cl = parse_row(all.data(), (int)all.size(), last);
            if (h->bench_wrap_codes)   // benchmark only: fold random-init tokens into each slot's 4096-code range
                for (size_t i = 0; i < cl.size(); ++i) cl[i] = ((cl[i] % 4096) + 4096) % 4096 + 4096 * (int)(i % 7);
            deinterleave(cl.data(), (int)cl.size(), l1[b], l2[b], l3[b]);
            frames[b] = (int)l1[b].size();
            any |= frames[b] > 0;
            if (wave_len) wave_len[b] = 0;
        }
        B2A_CHECK(any, B2A_ERR_GENERATION_FAILED, "No audio codes generated");   // LlamaTTS.swift:752-754
        const int64_t hop = b2a_snac_hop_length(h->snac);
        // group rows with equal frame counts into one batched codec call
        std::vector<bool> used(B, false);
        for (int b = 0; b < B; ++b) {
            if (used[b] || frames[b] == 0) continue;
            std::vector<int> grp;
            for (int c = b; c < B; ++c) if (!used[c] && frames[c] == frames[b]) { grp.push_back(c); used[c] = true; }
            const int F = frames[b], nb = (int)grp.size();
            const int64_t T = 4ll * F, wl = T * hop;
            B2A_CHECK(wl <= wave_cap, B2A_ERR_INVALID_INPUT, "tts generate: wave buffer too small");
            std::vector<int> c0v, c1v, c2v;
            for (int r : grp) { c0v.insert(c0v.end(), l1[r].begin(), l1[r].end()); c1v.insert(c1v.end(), l2[r].begin(), l2[r].end());
                                c2v.insert(c2v.end(), l3[r].begin(), l3[r].end()); }
            h->d_codes[0].upload(c0v.data(), c0v.size(), s);
            h->d_codes[1].upload(c1v.data(), c1v.size(), s);
            h->d_codes[2].upload(c2v.data(), c2v.size(), s);
            h->d_wave.alloc((size_t)nb * wl);
            const int* dc[3] = {h->d_codes[0].p, h->d_codes[1].p, h->d_codes[2].p};
            B2A_CUDA(cudaStreamSynchronize(s));   // host vectors above go out of scope after the copy
            const int32_t st = b2a_snac_decode_dev(h->snac, dc, nb, T, nullptr, 0, gp->seed, h->d_wave.p, s);
            B2A_CHECK(st == B2A_OK, B2A_ERR_AUDIO_DECODING_FAILED, std::string("SNAC decode failed: ") + b2a_last_error());
            for (int i = 0; i < nb; ++i) {
                const int r = grp[i];
                B2A_CUDA(cudaMemcpyAsync(wave_out + (size_t)r * wave_cap, h->d_wave.p + (size_t)i * wl, wl * sizeof(float),
                                         wave_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, s));
                if (wave_len) wave_len[r] = wl;
            }
            B2A_CUDA(cudaStreamSynchronize(s));
        }
        codec_t = now_s() - c0;
    }
    if (info) {
        info->prompt_token_count = L;
        info->generation_token_count = total_gen;
        info->prefill_time = t1 - t0;
        info->generate_time = t2 - t1;
        info->tokens_per_second = total_gen / std::max(1e-9, t2 - t1);
        info->codec_time = codec_t + codec_stream_t;
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        info->peak_memory_gb = (double)(tot - fr) / 1e9;
    }
}

extern "C" {

int32_t b2a_tts_generate_stream(b2a_tts* h, const int32_t* input_ids, int32_t B, int32_t L, const b2a_gen_params* gp,
                                int32_t frames_per_chunk, int32_t left_context_frames, int32_t* tokens_out, int32_t* n_tokens_out,
                                b2a_gen_info* info, b2a_token_cb on_token, b2a_audio_cb on_audio, void* user) {
    return guarded([&] {
        B2A_CHECK(on_audio && frames_per_chunk >= 1 && left_context_frames >= 0, B2A_ERR_INVALID_INPUT,
                  "b2a_tts_generate_stream: needs on_audio, frames_per_chunk >= 1, left_context_frames >= 0");
        StreamSpec ss;
        ss.frames_per_chunk = frames_per_chunk; ss.left_context = left_context_frames; ss.on_audio = on_audio;
        tts_generate_impl(h, input_ids, false, B, L, gp, tokens_out, n_tokens_out, nullptr, false, 0, nullptr, info, on_token, user, ss);
    });
}

int32_t b2a_tts_create(int32_t device, const b2a_llama_config* cfg, const b2a_tensor* tensors, int32_t n, b2a_snac* snac,
                       b2a_tts** out) {
    return guarded([&] {
        B2A_CHECK(out, B2A_ERR_INVALID_INPUT, "b2a_tts_create: null out");
        *out = nullptr;
        B2A_CHECK(cfg && tensors && n > 0, B2A_ERR_MODEL_NOT_INITIALIZED, "b2a_tts_create: missing config or weights");
        TensorTable tt(tensors, n);
        *out = new b2a_tts(device, *cfg, tt, snac);
    });
}

int32_t b2a_tts_create_random(int32_t device, const b2a_llama_config* cfg, float std, uint64_t seed, b2a_snac* snac,
                              b2a_tts** out) {
    return guarded([&] {
        B2A_CHECK(out, B2A_ERR_INVALID_INPUT, "b2a_tts_create_random: null out");
        *out = nullptr;
        B2A_CHECK(cfg && std > 0.f, B2A_ERR_MODEL_NOT_INITIALIZED, "b2a_tts_create_random: missing config");
        *out = new b2a_tts(device, *cfg, std, seed, snac);
    });
}

void* b2a_tts_stream(b2a_tts* h) { return h ? (void*)h->stream : nullptr; }

int32_t b2a_tts_time_steps(b2a_tts* h, int32_t B, int32_t ctx, int32_t iters, float* ms_per_step) {
    return guarded([&] {
        B2A_CHECK(h && ms_per_step && iters > 0, B2A_ERR_INVALID_INPUT, "b2a_tts_time_steps: bad argument");
        B2A_CHECK(B >= 1 && B <= h->cfg.max_batch && ctx >= 0 && ctx + iters < h->cfg.max_context, B2A_ERR_INVALID_INPUT,
                  "b2a_tts_time_steps: batch / context out of range");
        B2A_CUDA(cudaSetDevice(h->device));
        cudaStream_t s = h->stream;
        h->set_batch(B);
        const int MT = iters + 8;
        h->ids.alloc((size_t)B * 1);
        h->out_tokens.alloc((size_t)B * MT);
        h->recent.alloc(8);
        SampleArgs sa{};
        sa.logits = h->logits.p; sa.probs = h->probs.p; sa.tokens = h->tokens.p; sa.pos = h->pos.p; sa.recent = h->recent.p;
        sa.recent_n = h->recent_n.p; sa.out_tokens = h->out_tokens.p; sa.n_gen = h->n_gen.p; sa.done = h->done.p;
        sa.n_active = h->n_active.p; sa.forced = nullptr; sa.V = h->cfg.vocab_size; sa.R = 1; sa.max_tokens = MT;
        sa.temperature = 0.f; sa.top_p = 1.f; sa.rep_penalty = 1.f; sa.seed = 0; sa.mask_eos = 1;
        h->capture(B, sa, 1);
        B2A_CUDA(cudaMemsetAsync(h->ids.p, 0, (size_t)B * sizeof(int), s));
        init_rows_kernel<<<1, 32, 0, s>>>(h->ids.p, 1, B, 0, h->tokens.p, h->pos.p, h->recent.p, h->recent_n.p, h->n_gen.p,
                                          h->done.p, h->n_active.p, ctx);
        count_launch();
        // K/V beyond what earlier calls wrote is whatever is in the cache: zero it so reads are defined
        // (timing only; values do not matter).
        cudaEvent_t e0, e1;
        B2A_CUDA(cudaEventCreate(&e0));
        B2A_CUDA(cudaEventCreate(&e1));
        B2A_CUDA(cudaGraphLaunch(h->g_step, s));   // warm-up
        B2A_CUDA(cudaEventRecord(e0, s));
        for (int i = 0; i < iters - 1; ++i) B2A_CUDA(cudaGraphLaunch(h->g_step, s));
        B2A_CUDA(cudaEventRecord(e1, s));
        B2A_CUDA(cudaStreamSynchronize(s));
        count_launch(h->launches_step * iters);
        float ms = 0.f;
        B2A_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *ms_per_step = ms / (float)std::max(1, iters - 1);
    });
}

int32_t b2a_tts_prepare_input_ids(const int32_t* const* prompt_ids, const int32_t* lens, int32_t batch, int32_t* out,
                                  int32_t* out_len) {
    return guarded([&] {
        B2A_CHECK(prompt_ids && lens && out_len && batch > 0, B2A_ERR_INVALID_INPUT, "b2a_tts_prepare_input_ids: null argument");
        int mx = 0;
        for (int b = 0; b < batch; ++b) mx = std::max(mx, lens[b]);
        *out_len = mx + 3;
        if (!out) return;
        for (int b = 0; b < batch; ++b) {   // LlamaTTS.swift:499-543
            int32_t* r = out + (size_t)b * (mx + 3);
            int j = 0;
            for (; j < mx - lens[b]; ++j) r[j] = TOK_PAD;
            r[j++] = TOK_START_OF_HUMAN;
            for (int i = 0; i < lens[b]; ++i) r[j++] = prompt_ids[b][i];
            r[j++] = TOK_END_OF_TEXT;
            r[j++] = TOK_END_OF_HUMAN;
        }
    });
}

// Debug / parity hook: residual stream seen by every RMSNorm (2 per layer + final) for the LAST position of
// the last b2a_tts_forward_logits call made while tracing was enabled; out is [2*layers+1, batch, hidden].
int32_t b2a_tts_debug_trace(b2a_tts* h, int32_t enable, int32_t batch, float* out) {
    return guarded([&] {
        B2A_CHECK(h, B2A_ERR_INVALID_INPUT, "b2a_tts_debug_trace: null handle");
        B2A_CUDA(cudaSetDevice(h->device));
        const int H = h->cfg.hidden_size, n = 2 * h->cfg.num_hidden_layers + 1;
        if (out) {
            B2A_CHECK(h->trace.p && batch >= 1 && batch <= 8, B2A_ERR_INVALID_INPUT, "b2a_tts_debug_trace: nothing traced");
            B2A_CUDA(cudaStreamSynchronize(h->stream));
            for (int i = 0; i < n; ++i)
                B2A_CUDA(cudaMemcpy(out + (size_t)i * batch * H, h->trace.p + (size_t)i * 8 * H, (size_t)batch * H * sizeof(float),
                                    cudaMemcpyDeviceToHost));
        }
        h->trace_on = enable != 0;
        if (h->trace_on) { h->trace.alloc((size_t)n * 8 * H); h->drop_graphs(); }
    });
}

int32_t b2a_tts_forward_logits(b2a_tts* h, const int32_t* ids, int32_t B, int32_t L, int32_t reset_cache, float* logits_out) {
    return guarded([&] {
        B2A_CHECK(h && ids && logits_out, B2A_ERR_INVALID_INPUT, "b2a_tts_forward_logits: null argument");
        B2A_CHECK(B >= 1 && B <= h->cfg.max_batch && L >= 1, B2A_ERR_INVALID_INPUT, "b2a_tts_forward_logits: bad batch / length");
        B2A_CUDA(cudaSetDevice(h->device));
        cudaStream_t s = h->stream;
        h->set_batch(B);
        h->drop_graphs();
        std::vector<int> hp(8, 0);
        if (!reset_cache) B2A_CUDA(cudaMemcpy(hp.data(), h->pos.p, 8 * sizeof(int), cudaMemcpyDeviceToHost));
        const int start = reset_cache ? 0 : hp[0];
        B2A_CHECK(start + L <= h->cfg.max_context, B2A_ERR_INVALID_INPUT, "b2a_tts_forward_logits: context overflow");
        h->ids.alloc((size_t)B * L);
        B2A_CUDA(cudaMemcpyAsync(h->ids.p, ids, (size_t)B * L * sizeof(int), cudaMemcpyHostToDevice, s));
        const int V = h->cfg.vocab_size;
        std::vector<int> tk(8, 0), ps(8, -1);
        for (int p = 0; p < L; ++p) {
            for (int b = 0; b < B; ++b) { tk[b] = ids[(size_t)b * L + p]; ps[b] = start + p; }
            B2A_CUDA(cudaMemcpyAsync(h->tokens.p, tk.data(), 8 * sizeof(int), cudaMemcpyHostToDevice, s));
            B2A_CUDA(cudaMemcpyAsync(h->pos.p, ps.data(), 8 * sizeof(int), cudaMemcpyHostToDevice, s));
            h->run_layers(B, s);
            h->run_lm_head(B, s);
            for (int b = 0; b < B; ++b)
                B2A_CUDA(cudaMemcpyAsync(logits_out + ((size_t)b * L + p) * V, h->logits.p + (size_t)b * V, V * sizeof(float),
                                         cudaMemcpyDeviceToHost, s));
            B2A_CUDA(cudaStreamSynchronize(s));
        }
        for (int b = 0; b < B; ++b) ps[b] = start + L;
        B2A_CUDA(cudaMemcpy(h->pos.p, ps.data(), 8 * sizeof(int), cudaMemcpyHostToDevice));
        B2A_CUDA(cudaGetLastError());
    });
}

int32_t b2a_tts_generate(b2a_tts* h, const int32_t* input_ids, int32_t B, int32_t L, const b2a_gen_params* gp,
                         int32_t* tokens_out, int32_t* n_tokens_out, float* wave_out, int64_t wave_cap, int64_t* wave_len,
                         b2a_gen_info* info, b2a_token_cb on_token, void* user) {
    return guarded([&] {
        tts_generate_impl(h, input_ids, false, B, L, gp, tokens_out, n_tokens_out, wave_out, false, wave_cap, wave_len, info,
                          on_token, user);
    });
}

int32_t b2a_tts_generate_dev(b2a_tts* h, const int32_t* d_input_ids, int32_t B, int32_t L, const b2a_gen_params* gp,
                             float* d_wave_out, int64_t wave_cap, int64_t* wave_len, b2a_gen_info* info) {
    return guarded([&] {
        tts_generate_impl(h, d_input_ids, true, B, L, gp, nullptr, nullptr, d_wave_out, true, wave_cap, wave_len, info, nullptr,
                          nullptr);
    });
}

int32_t b2a_tts_set_bench_flags(b2a_tts* h, int32_t mask_eos, int32_t wrap_codes) {
    if (!h) return B2A_ERR_INVALID_INPUT;
    h->bench_mask_eos = mask_eos != 0;
    h->bench_wrap_codes = wrap_codes != 0;
    return B2A_OK;
}

int32_t b2a_tts_cancel(b2a_tts* h) {
    if (!h) return B2A_ERR_INVALID_INPUT;
    h->cancel.store(1);
    return B2A_OK;
}

int32_t b2a_tts_parse_output(const int32_t* tokens, int32_t batch, int32_t n, int32_t* code_lists_out, int32_t* code_lens) {
    return guarded([&] {
        B2A_CHECK(tokens && code_lists_out && code_lens && batch > 0 && n >= 0, B2A_ERR_INVALID_INPUT, "b2a_tts_parse_output: bad argument");
        int last = -1;   // LlamaTTS.swift:391-398: last match in row-major visiting order
        for (int i = 0; i < batch; ++i)
            for (int j = 0; j < n; ++j)
                if (tokens[(size_t)i * n + j] == TOK_START_OF_SPEECH) last = j;
        for (int i = 0; i < batch; ++i) {
            std::vector<int> r = parse_row(tokens + (size_t)i * n, n, last);
            code_lens[i] = (int)r.size();
            memcpy(code_lists_out + (size_t)i * n, r.data(), r.size() * sizeof(int));
        }
    });
}

int32_t b2a_tts_deinterleave(const int32_t* code_list, int32_t n, int32_t* c0, int32_t* c1, int32_t* c2, int32_t* n_frames) {
    return guarded([&] {
        B2A_CHECK(code_list && c0 && c1 && c2 && n_frames && n >= 0, B2A_ERR_INVALID_INPUT, "b2a_tts_deinterleave: bad argument");
        const int groups = (n + 1) / 7;
        B2A_CHECK(groups * 7 <= n, B2A_ERR_INVALID_INPUT, "b2a_tts_deinterleave: code list must hold whole 7-token frames");
        std::vector<int> l1, l2, l3;
        deinterleave(code_list, n, l1, l2, l3);
        memcpy(c0, l1.data(), l1.size() * sizeof(int));
        memcpy(c1, l2.data(), l2.size() * sizeof(int));
        memcpy(c2, l3.data(), l3.size() * sizeof(int));
        *n_frames = groups;
    });
}

int32_t b2a_tts_interleave(const int32_t* c0, const int32_t* c1, const int32_t* c2, int32_t n_frames, int32_t* cl) {
    return guarded([&] {   // llamaEncodeAudioToCodes, LlamaTTS.swift:85-95
        B2A_CHECK(c0 && c1 && c2 && cl && n_frames >= 0, B2A_ERR_INVALID_INPUT, "b2a_tts_interleave: bad argument");
        for (int i = 0; i < n_frames; ++i) {
            int32_t* o = cl + 7 * i;
            o[0] = c0[i];
            o[1] = c1[2 * i] + 4096;
            o[2] = c2[4 * i] + 2 * 4096;
            o[3] = c2[4 * i + 1] + 3 * 4096;
            o[4] = c1[2 * i + 1] + 4 * 4096;
            o[5] = c2[4 * i + 2] + 5 * 4096;
            o[6] = c2[4 * i + 3] + 6 * 4096;
        }
    });
}

void b2a_tts_destroy(b2a_tts* h) { delete h; }

}  // extern "C"

// ================================================================================================
// Qwen3-TTS talker + code predictor on the same engine.
//   Qwen3TTSTalker.swift:127-366, Qwen3TTSCodePredictor.swift:14-243, Qwen3TTS.swift:380-495 (frame loop), :1003-1118 (sampleToken)
// Two b2a_tts stacks (talker: inputs are embeddings, untied codec_head; predictor: 5 layers, heads owned here) share one stream.
// One frame = ONE CUDA graph: talker step -> sampler -> [hidden, embed(c0)] + 14 more predictor steps (cache positions 0..16 are
// simply overwritten every frame = the reference's per-frame cache trim) -> summed-embedding feedback + bookkeeping.
// ================================================================================================
namespace b2a {

// dst[b, :] = float(table[ids[b * id_stride + id_col], :]);  optionally pos[b] = pos_value (the predictor's cache position)
__global__ void q3_gather_kernel(const bf16* __restrict__ table, int rows, const int* __restrict__ ids, int id_stride, int id_col,
                                 float* __restrict__ dst, int H, int* pos, int pos_value) {
    const int b = blockIdx.x;
    if (!pos) pdl_trigger();     // a kernel that writes pos[] must not trigger early (see attn_decode_cluster_kernel)
    pdl_wait();
    int t = ids[b * id_stride + id_col];
    t = min(max(t, 0), rows - 1);
    for (int i = threadIdx.x; i < H; i += blockDim.x) dst[(long long)b * H + i] = __bfloat162float(table[(long long)t * H + i]);
    if (pos && threadIdx.x == 0) pos[b] = pos_value;
}
// dst[b, :] = src[b * src_stride + :]; optionally pos[b] = pos_value (pos_value < 0: pos untouched)
__global__ void q3_copy_rows_kernel(const float* __restrict__ src, long long src_stride, float* __restrict__ dst, int H, int* pos, int pos_value) {
    const int b = blockIdx.x;
    if (!(pos && pos_value >= 0)) pdl_trigger();     // writers of pos[] never trigger early
    pdl_wait();
    for (int i = threadIdx.x; i < H; i += blockDim.x) dst[(long long)b * H + i] = src[(long long)b * src_stride + i];
    if (pos && pos_value >= 0 && threadIdx.x == 0) pos[b] = pos_value;
}
// y[t, o] = act(b[o] + sum_k W[o, k] x[t, k]); bf16 W, fp32 x / y; one warp per output, grid (ceil(O / 8), T).  The prompt's
// ResizeMLP only (text_projection, Qwen3TTSTalker.swift:209-221): a few dozen rows per utterance, off the frame loop.
__global__ void q3_linear_kernel(const bf16* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ x, float* __restrict__ y,
                                 int O, int K, int silu) {
    const int o = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), t = blockIdx.y, lane = threadIdx.x & 31;
    if (o >= O) return;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(__bfloat162float(W[(long long)o * K + k]), x[(long long)t * K + k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
        acc += bias ? bias[o] : 0.f;
        y[(long long)t * O + o] = silu ? acc / (1.0f + __expf(-acc)) : acc;
    }
}

struct Q3Feedback {
    const bf16* codec_emb; int codec_rows;
    const bf16* const* cp_emb;   // device array of G-1 tables [cp_vocab, H]
    int cp_rows;
    const int* codes;            // [8, G] this frame
    const float* trailing;       // [B, n_max, H]
    const int* n_trailing;       // [B]
    int n_max;
    const float* pad;            // [H]
    float* x_in;                 // [8, H] next talker input
    int* talker_pos;             // [8]
    int* row_frame;              // [8] frames stepped so far (= index of the trailing text row to consume)
    int* out_codes;              // [B, max_tokens, G]
    int* n_frames; int* done; int* n_active;
    int G, H, max_tokens, eos, mask_eos;
};
// x_next = text + codec_embed(c0) + sum_i predictor_embed_i(c_{i+1})  (Qwen3TTS.swift:470-487) + per-row bookkeeping (:424-428)
__global__ void q3_feedback_kernel(Q3Feedback a) {
    const int b = blockIdx.x;
    pdl_wait();                  // NO pdl_trigger(): writes the talker's pos[]
    const int f = a.row_frame[b];
    const float* text = f < a.n_trailing[b] ? a.trailing + ((long long)b * a.n_max + f) * a.H : a.pad;
    const int* c = a.codes + b * a.G;
    for (int i = threadIdx.x; i < a.H; i += blockDim.x) {
        float e = __bfloat162float(a.codec_emb[(long long)min(max(c[0], 0), a.codec_rows - 1) * a.H + i]);
        for (int g = 1; g < a.G; ++g) e += __bfloat162float(a.cp_emb[g - 1][(long long)min(max(c[g], 0), a.cp_rows - 1) * a.H + i]);
        a.x_in[(long long)b * a.H + i] = text[i] + e;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a.row_frame[b] = f + 1;
        a.talker_pos[b] += 1;
        if (!a.done[b]) {
            if (c[0] == a.eos && !a.mask_eos) { a.done[b] = 1; atomicSub(a.n_active, 1); }
            else {
                const int n = a.n_frames[b];
                if (n < a.max_tokens) for (int g = 0; g < a.G; ++g) a.out_codes[((long long)b * a.max_tokens + n) * a.G + g] = c[g];
                a.n_frames[b] = n + 1;
                if (n + 1 >= a.max_tokens) { a.done[b] = 1; atomicSub(a.n_active, 1); }
            }
        }
    }
}
__global__ void q3_init_rows_kernel(int B, int L, int* talker_pos, int* row_frame, int* n_frames, int* done, int* n_active, unsigned* seen, int words) {
    const int b = threadIdx.x;
    if (b == 0) *n_active = B;
    if (b < 8) { talker_pos[b] = L - 1; row_frame[b] = 0; n_frames[b] = 0; done[b] = b < B ? 0 : 1; }
    for (int i = threadIdx.x; i < 8 * words; i += blockDim.x) seen[i] = 0u;
}
// bench only: the talker never emits EOS -- its logit is dropped before the sampler sees the row
__global__ void q3_mask_eos_kernel(float* logits, int V, int eos) {
    if (threadIdx.x == 0 && eos >= 0 && eos < V) logits[(long long)blockIdx.x * V + eos] = -INFINITY;
}

}  // namespace b2a

struct b2a_qwen3_talker {
    int device;
    b2a_qwen3_talker_config cfg;
    b2a_tts* talker = nullptr;
    b2a_tts* pred = nullptr;
    cudaStream_t stream = nullptr;        // = talker->stream
    DBuf<bf16> codec_emb, text_emb, fc1_w, fc2_w;
    DBuf<float> fc1_b, fc2_b;
    std::vector<DBuf<bf16>> cp_emb, cp_head;
    std::vector<CUtensorMap> tm_cp_head;
    int cp_head_rows = 128;
    DBuf<const bf16*> cp_emb_ptrs;
    DBuf<float> x_in, hid, px, trailing, pad, embeds, tmp_a, tmp_b;
    DBuf<int> codes, out_codes, n_frames, done, n_active, row_frame, n_trailing, ids;
    DBuf<unsigned> seen;
    HBuf<int> h_flag;
    cudaGraphExec_t g_frame = nullptr;
    b2a_qwen3_gen_params g_params{};
    int g_B = 0, g_max_tokens = 0, g_nmax = 0, g_mask = -1;
    const float* g_trailing = nullptr;
    int launches_frame = 0;
    int bench_mask_eos = 0;
    std::atomic<int> cancel{0};

    ~b2a_qwen3_talker() {
        if (g_frame) cudaGraphExecDestroy(g_frame);
        delete pred;
        delete talker;
    }
    int G() const { return cfg.num_code_groups; }
    int H() const { return cfg.hidden_size; }

    static b2a_llama_config stack_cfg(const b2a_qwen3_talker_config& c, bool predictor) {
        b2a_llama_config l{};
        l.hidden_size = predictor ? c.cp_hidden_size : c.hidden_size;
        l.num_hidden_layers = predictor ? c.cp_num_hidden_layers : c.num_hidden_layers;
        l.intermediate_size = predictor ? c.cp_intermediate_size : c.intermediate_size;
        l.num_attention_heads = predictor ? c.cp_num_attention_heads : c.num_attention_heads;
        l.num_key_value_heads = predictor ? c.cp_num_key_value_heads : c.num_key_value_heads;
        l.head_dim = predictor ? c.cp_head_dim : c.head_dim;
        l.vocab_size = predictor ? c.cp_vocab_size : c.vocab_size;
        l.rms_norm_eps = predictor ? c.cp_rms_norm_eps : c.rms_norm_eps;
        l.rope_theta = predictor ? c.cp_rope_theta : c.rope_theta;
        // plain rotate-half RoPE: factor 1 makes Llama3ScaledRoPE's three wavelength bands collapse to base^(2i/d)
        l.rope_factor = 1.0f; l.rope_low_freq_factor = 1.0f; l.rope_high_freq_factor = 4.0f; l.rope_old_context_len = 8192.0f;
        l.tie_word_embeddings = 0;
        l.max_batch = c.max_batch;
        l.max_context = predictor ? std::max(32, c.num_code_groups + 8) : c.max_context;
        return l;
    }
    static StackSpec talker_spec() { StackSpec s; s.prefix = "model."; s.qk_norm = true; s.has_embed = false; s.head = "codec_head.weight"; s.has_head = true; return s; }
    static StackSpec pred_spec() { StackSpec s; s.prefix = "code_predictor.model."; s.qk_norm = true; s.has_embed = false; s.has_head = false; return s; }

    void check() {
        const b2a_qwen3_talker_config& c = cfg;
        B2A_CHECK(c.head_dim == HD && c.cp_head_dim == HD, B2A_ERR_INVALID_INPUT, "qwen3 talker: head_dim must be 128");
        B2A_CHECK(c.cp_hidden_size == c.hidden_size, B2A_ERR_INVALID_INPUT,
                  "qwen3 talker: code predictor hidden size must equal the talker's (small_to_mtp_projection is not implemented)");
        B2A_CHECK(c.vocab_size >= 1 && c.vocab_size <= q3s::SLOTS && c.cp_vocab_size >= 1 && c.cp_vocab_size <= q3s::SLOTS, B2A_ERR_INVALID_INPUT,
                  "qwen3 talker: codec vocabularies must be <= 4096");
        B2A_CHECK(c.num_code_groups >= 2 && c.num_code_groups <= 32, B2A_ERR_INVALID_INPUT, "qwen3 talker: num_code_groups must be in 2..32");
        B2A_CHECK(c.max_batch >= 1 && c.max_batch <= 8, B2A_ERR_INVALID_INPUT, "qwen3 talker: max_batch must be in 1..8");
    }
    void alloc_state() {
        stream = talker->stream;
        const int Hh = H();
        x_in.alloc((size_t)8 * Hh); hid.alloc((size_t)8 * Hh); px.alloc((size_t)8 * Hh); pad.alloc(Hh);
        B2A_CUDA(cudaMemset(x_in.p, 0, (size_t)8 * Hh * sizeof(float)));
        B2A_CUDA(cudaMemset(hid.p, 0, (size_t)8 * Hh * sizeof(float)));
        B2A_CUDA(cudaMemset(px.p, 0, (size_t)8 * Hh * sizeof(float)));
        codes.alloc((size_t)8 * G()); n_frames.alloc(8); done.alloc(8); n_active.alloc(1); row_frame.alloc(8); n_trailing.alloc(8);
        B2A_CUDA(cudaMemset(codes.p, 0, (size_t)8 * G() * sizeof(int)));
        seen.alloc((size_t)8 * cdiv(cfg.vocab_size, 32));
        h_flag.alloc(16);
        std::vector<const bf16*> ptrs;
        for (auto& e : cp_emb) ptrs.push_back(e.p);
        cp_emb_ptrs.upload(ptrs.data(), ptrs.size());
        if (pred->use_tc) {
            tm_cp_head.clear();
            cp_head_rows = b2a_tts::pick_tile_rows(cfg.cp_vocab_size, pred->num_sms);
            for (auto& hd : cp_head) tm_cp_head.push_back(tc::make_tmap_bf16(hd.p, cfg.cp_vocab_size, Hh, cp_head_rows));
        } else tm_cp_head.resize(cp_head.size());
        talker->x_ext = x_in.p; talker->normed_out = hid.p;
        pred->x_ext = px.p;
        talker->set_batch(8); pred->set_batch(8);
        B2A_CUDA(cudaDeviceSynchronize());
    }

    b2a_qwen3_talker(int dev, const b2a_qwen3_talker_config& c, const TensorTable& tt) : device(dev), cfg(c) {
        check();
        talker = new b2a_tts(dev, stack_cfg(c, false), tt, nullptr, talker_spec());
        pred = new b2a_tts(dev, stack_cfg(c, true), tt, nullptr, pred_spec());
        const int Hh = H(), TH = c.text_hidden_size;
        b2a_tts::upload_bf16(tt, "model.codec_embedding.weight", (int64_t)c.vocab_size * Hh, codec_emb, 0, (size_t)c.vocab_size * Hh);
        b2a_tts::upload_bf16(tt, "model.text_embedding.weight", (int64_t)c.text_vocab_size * TH, text_emb, 0, (size_t)c.text_vocab_size * TH);
        b2a_tts::upload_bf16(tt, "text_projection.linear_fc1.weight", (int64_t)TH * TH, fc1_w, 0, (size_t)TH * TH);
        b2a_tts::upload_bf16(tt, "text_projection.linear_fc2.weight", (int64_t)Hh * TH, fc2_w, 0, (size_t)Hh * TH);
        std::vector<float> b1 = tt.f32("text_projection.linear_fc1.bias", TH), b2 = tt.f32("text_projection.linear_fc2.bias", Hh);
        fc1_b.upload(b1.data(), TH); fc2_b.upload(b2.data(), Hh);
        cp_emb.resize(G() - 1); cp_head.resize(G() - 1);
        for (int i = 0; i < G() - 1; ++i) {
            b2a_tts::upload_bf16(tt, "code_predictor.model.codec_embedding." + std::to_string(i) + ".weight", (int64_t)c.cp_vocab_size * Hh, cp_emb[i], 0,
                                 (size_t)c.cp_vocab_size * Hh);
            b2a_tts::upload_bf16(tt, "code_predictor.lm_head." + std::to_string(i) + ".weight", (int64_t)c.cp_vocab_size * Hh, cp_head[i], 0,
                                 (size_t)c.cp_vocab_size * Hh);
        }
        alloc_state();
    }
    // device-drawn weights (bench): every matrix N(0, std^2) bf16, biases 0
    b2a_qwen3_talker(int dev, const b2a_qwen3_talker_config& c, float std, unsigned long long seed) : device(dev), cfg(c) {
        check();
        talker = new b2a_tts(dev, stack_cfg(c, false), std, seed, nullptr, talker_spec());
        pred = new b2a_tts(dev, stack_cfg(c, true), std, seed + 7777, nullptr, pred_spec());
        const int Hh = H(), TH = c.text_hidden_size;
        unsigned long long sd = seed * 7919ull + 3;
        auto rnd = [&](DBuf<bf16>& d, size_t n) {
            d.alloc(n);
            random_bf16_kernel<<<148 * 8, 256, 0, talker->stream>>>(d.p, (long long)n, std, sd++);
            count_launch();
        };
        rnd(codec_emb, (size_t)c.vocab_size * Hh); rnd(text_emb, (size_t)c.text_vocab_size * TH);
        rnd(fc1_w, (size_t)TH * TH); rnd(fc2_w, (size_t)Hh * TH);
        fc1_b.alloc(TH); fc2_b.alloc(Hh);
        B2A_CUDA(cudaMemsetAsync(fc1_b.p, 0, TH * sizeof(float), talker->stream));
        B2A_CUDA(cudaMemsetAsync(fc2_b.p, 0, Hh * sizeof(float), talker->stream));
        cp_emb.resize(G() - 1); cp_head.resize(G() - 1);
        for (int i = 0; i < G() - 1; ++i) { rnd(cp_emb[i], (size_t)c.cp_vocab_size * Hh); rnd(cp_head[i], (size_t)c.cp_vocab_size * Hh); }
        B2A_CUDA(cudaStreamSynchronize(talker->stream));
        alloc_state();
    }

    // ids [n] (device) -> out [n, H]: text_projection(text_embedding(ids)) = fc2(silu(fc1(e)))
    void embed_text_dev(const int* d_ids, int n, float* d_out, cudaStream_t s) {
        const int TH = cfg.text_hidden_size, Hh = H();
        tmp_a.alloc((size_t)n * TH); tmp_b.alloc((size_t)n * TH);
        q3_gather_kernel<<<n, 256, 0, s>>>(text_emb.p, cfg.text_vocab_size, d_ids, 1, 0, tmp_a.p, TH, nullptr, 0);
        q3_linear_kernel<<<dim3(cdiv(TH, 8), n), 256, 0, s>>>(fc1_w.p, fc1_b.p, tmp_a.p, tmp_b.p, TH, TH, 1);
        q3_linear_kernel<<<dim3(cdiv(Hh, 8), n), 256, 0, s>>>(fc2_w.p, fc2_b.p, tmp_b.p, d_out, Hh, TH, 0);
        count_launch(3);
        B2A_CUDA(cudaGetLastError());
    }

    q3s::Args sampler_args(const b2a_qwen3_gen_params& p, bool is_talker, int group) const {
        q3s::Args a{};
        a.logits = is_talker ? talker->logits.p : pred->logits.p;
        a.V = is_talker ? cfg.vocab_size : cfg.cp_vocab_size;
        a.temperature = p.temperature; a.top_p = p.top_p; a.top_k = p.top_k; a.min_p = p.min_p;
        a.rep_penalty = is_talker ? p.repetition_penalty : 1.0f;
        a.eos = is_talker ? cfg.codec_eos_token_id : -1;
        a.suppress_lo = is_talker ? std::max(cfg.vocab_size - 1024, 0) : 0;      // the special-token block except EOS (:383-385)
        a.suppress_hi = is_talker ? cfg.vocab_size : 0;
        a.seen = is_talker ? seen.p : nullptr;
        a.track = is_talker ? 1 : 0;
        a.seed = p.seed;
        a.step = group; a.step_ptr = row_frame.p; a.step_mul = G();
        a.tokens = nullptr; a.filtered = nullptr;
        return a;
    }
    void talker_step(int B, cudaStream_t s, bool with_head) {
        talker->run_layers(B, s);
        if (with_head) talker->run_lm_head(B, s);
    }

    void capture_frame(int B, const b2a_qwen3_gen_params& p, int max_tokens, int nmax) {
        if (g_frame && g_B == B && g_max_tokens == max_tokens && g_nmax == nmax && g_mask == bench_mask_eos && g_trailing == trailing.p &&
            memcmp(&g_params, &p, sizeof(p)) == 0) return;
        if (g_frame) { cudaGraphExecDestroy(g_frame); g_frame = nullptr; }
        cudaStream_t s = stream;
        const int n0 = (int)b2a_launch_count();
        cudaGraph_t g;
        B2A_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        // 1. talker step on x_in -> hid (final norm), codec logits -> c0
        talker_step(B, s, true);
        if (bench_mask_eos) { q3_mask_eos_kernel<<<B, 32, 0, s>>>(talker->logits.p, cfg.vocab_size, cfg.codec_eos_token_id); count_launch(); }
        {
            q3s::Args a = sampler_args(p, true, 0);
            a.tokens = codes.p; a.tokens_stride = G();
            q3s::sample_kernel<<<B, q3s::THREADS, 0, s>>>(a);
            count_launch();
        }
        // 2. code predictor: position 0 = the talker's hidden state, position 1 = codec_embed(c0) -> head 0 -> c1, then
        //    position k + 1 = predictor_embed_{k-1}(c_k) -> head k -> c_{k+1}
        launch_pdl(q3_copy_rows_kernel, dim3(B), dim3(256), 0, s, (const float*)hid.p, (long long)H(), px.p, H(), pred->pos.p, 0);
        pred->run_layers(B, s);
        for (int k = 0; k < G() - 1; ++k) {
            if (k == 0) launch_pdl(q3_gather_kernel, dim3(B), dim3(256), 0, s, (const bf16*)codec_emb.p, cfg.vocab_size, (const int*)codes.p, G(), 0, px.p, H(), pred->pos.p, 1);
            else launch_pdl(q3_gather_kernel, dim3(B), dim3(256), 0, s, (const bf16*)cp_emb[k - 1].p, cfg.cp_vocab_size, (const int*)codes.p, G(), k, px.p, H(), pred->pos.p, k + 1);
            pred->run_layers(B, s);
            pred->run_final_norm(B, s);
            pred->run_head(tm_cp_head[k], cp_head[k].p, cfg.cp_vocab_size, pred->logits.p, B, s, cp_head_rows);
            q3s::Args a = sampler_args(p, false, k + 1);
            a.tokens = codes.p + (k + 1); a.tokens_stride = G();
            q3s::sample_kernel<<<B, q3s::THREADS, 0, s>>>(a);
            count_launch();
        }
        // 3. feedback + bookkeeping
        Q3Feedback f{codec_emb.p, cfg.vocab_size, cp_emb_ptrs.p, cfg.cp_vocab_size, codes.p, trailing.p, n_trailing.p, nmax, pad.p, x_in.p,
                     talker->pos.p, row_frame.p, out_codes.p, n_frames.p, done.p, n_active.p, G(), H(), max_tokens, cfg.codec_eos_token_id, bench_mask_eos};
        launch_pdl(q3_feedback_kernel, dim3(B), dim3(256), 0, s, f);
        B2A_CUDA(cudaStreamEndCapture(s, &g));
        B2A_CUDA(cudaGraphInstantiate(&g_frame, g, 0));
        cudaGraphDestroy(g);
        launches_frame = (int)b2a_launch_count() - n0;
        g_B = B; g_params = p; g_max_tokens = max_tokens; g_nmax = nmax; g_mask = bench_mask_eos; g_trailing = trailing.p;
    }

    // positions 0 .. L-2 of the prompt through the talker (no logits needed); leaves x_in = embeds[:, L-1], talker pos = L-1
    void prefill(int B, int L, cudaStream_t s) {
        for (int p = 0; p < L; ++p) {
            launch_pdl(q3_copy_rows_kernel, dim3(B), dim3(256), 0, s, (const float*)(embeds.p + (size_t)p * H()), (long long)L * H(), x_in.p, H(),
                       talker->pos.p, p);
            if (p < L - 1) talker->run_layers(B, s);
        }
    }
};

extern "C" {

int32_t b2a_qwen3_talker_create(int32_t device, const b2a_qwen3_talker_config* cfg, const b2a_tensor* tensors, int32_t n, b2a_qwen3_talker** out) {
    return guarded([&] {
        B2A_CHECK(out, B2A_ERR_INVALID_INPUT, "b2a_qwen3_talker_create: null out");
        *out = nullptr;
        B2A_CHECK(cfg && tensors && n > 0, B2A_ERR_MODEL_NOT_INITIALIZED, "b2a_qwen3_talker_create: missing config or weights");
        TensorTable tt(tensors, n);
        *out = new b2a_qwen3_talker(device, *cfg, tt);
    });
}
int32_t b2a_qwen3_talker_create_random(int32_t device, const b2a_qwen3_talker_config* cfg, float std, uint64_t seed, b2a_qwen3_talker** out) {
    return guarded([&] {
        B2A_CHECK(out, B2A_ERR_INVALID_INPUT, "b2a_qwen3_talker_create_random: null out");
        *out = nullptr;
        B2A_CHECK(cfg && std > 0.f, B2A_ERR_MODEL_NOT_INITIALIZED, "b2a_qwen3_talker_create_random: missing config");
        *out = new b2a_qwen3_talker(device, *cfg, std, seed);
    });
}
void* b2a_qwen3_talker_stream(b2a_qwen3_talker* h) { return h ? (void*)h->stream : nullptr; }
int32_t b2a_qwen3_talker_set_bench_flags(b2a_qwen3_talker* h, int32_t mask_eos) {
    if (!h) return B2A_ERR_INVALID_INPUT;
    h->bench_mask_eos = mask_eos != 0;
    return B2A_OK;
}
int32_t b2a_qwen3_talker_cancel(b2a_qwen3_talker* h) {
    if (!h) return B2A_ERR_INVALID_INPUT;
    h->cancel.store(1);
    return B2A_OK;
}
void b2a_qwen3_talker_destroy(b2a_qwen3_talker* h) { delete h; }

int32_t b2a_qwen3_talker_embed_text(b2a_qwen3_talker* h, const int32_t* ids, int32_t n, float* out) {
    return guarded([&] {
        B2A_CHECK(h && ids && out && n >= 1, B2A_ERR_INVALID_INPUT, "b2a_qwen3_talker_embed_text: bad argument");
        B2A_CUDA(cudaSetDevice(h->device));
        for (int i = 0; i < n; ++i) B2A_CHECK(ids[i] >= 0 && ids[i] < h->cfg.text_vocab_size, B2A_ERR_INVALID_INPUT, "b2a_qwen3_talker_embed_text: id out of range");
        cudaStream_t s = h->stream;
        h->ids.upload(ids, n, s);
        h->embeds.alloc((size_t)n * h->H());
        h->embed_text_dev(h->ids.p, n, h->embeds.p, s);
        B2A_CUDA(cudaMemcpyAsync(out, h->embeds.p, (size_t)n * h->H() * sizeof(float), cudaMemcpyDeviceToHost, s));
        B2A_CUDA(cudaStreamSynchronize(s));
    });
}
int32_t b2a_qwen3_talker_embed_codec(b2a_qwen3_talker* h, const int32_t* ids, int32_t n, float* out) {
    return guarded([&] {
        B2A_CHECK(h && ids && out && n >= 1, B2A_ERR_INVALID_INPUT, "b2a_qwen3_talker_embed_codec: bad argument");
        B2A_CUDA(cudaSetDevice(h->device));
        for (int i = 0; i < n; ++i) B2A_CHECK(ids[i] >= 0 && ids[i] < h->cfg.vocab_size, B2A_ERR_INVALID_INPUT, "b2a_qwen3_talker_embed_codec: id out of range");
        cudaStream_t s = h->stream;
        h->ids.upload(ids, n, s);
        h->embeds.alloc((size_t)n * h->H());
        q3_gather_kernel<<<n, 256, 0, s>>>(h->codec_emb.p, h->cfg.vocab_size, h->ids.p, 1, 0, h->embeds.p, h->H(), nullptr, 0);
        count_launch();
        B2A_CUDA(cudaMemcpyAsync(out, h->embeds.p, (size_t)n * h->H() * sizeof(float), cudaMemcpyDeviceToHost, s));
        B2A_CUDA(cudaStreamSynchronize(s));
    });
}

int32_t b2a_qwen3_talker_forward(b2a_qwen3_talker* h, const float* input_embeds, int32_t B, int32_t L, float* logits_out, float* hidden_out) {
    return guarded([&] {
        B2A_CHECK(h && input_embeds && B >= 1 && B <= h->cfg.max_batch && L >= 1 && L <= h->cfg.max_context, B2A_ERR_INVALID_INPUT,
                  "b2a_qwen3_talker_forward: bad argument");
        B2A_CUDA(cudaSetDevice(h->device));
        cudaStream_t s = h->stream;
        h->talker->set_batch(B); h->talker->drop_graphs();
        h->embeds.upload(input_embeds, (size_t)B * L * h->H(), s);
        h->prefill(B, L, s);
        h->talker_step(B, s, true);
        if (logits_out)
            B2A_CUDA(cudaMemcpyAsync(logits_out, h->talker->logits.p, (size_t)B * h->cfg.vocab_size * sizeof(float), cudaMemcpyDeviceToHost, s));
        if (hidden_out) B2A_CUDA(cudaMemcpyAsync(hidden_out, h->hid.p, (size_t)B * h->H() * sizeof(float), cudaMemcpyDeviceToHost, s));
        B2A_CUDA(cudaStreamSynchronize(s));
        B2A_CUDA(cudaGetLastError());
    });
}

int32_t b2a_qwen3_talker_generate(b2a_qwen3_talker* h, const float* input_embeds, int32_t B, int32_t L, const float* trailing_text_hidden,
                                  const int32_t* n_trailing, int32_t n_trailing_max, const float* tts_pad_embed, const b2a_qwen3_gen_params* gp,
                                  int32_t* codes_out, int32_t* n_frames_out, b2a_gen_info* info, b2a_frame_cb on_frame, void* user) {
    return guarded([&] {
        B2A_CHECK(h && input_embeds && tts_pad_embed && gp && codes_out && n_frames_out, B2A_ERR_INVALID_INPUT, "qwen3 generate: null argument");
        B2A_CHECK(B >= 1 && B <= h->cfg.max_batch, B2A_ERR_INVALID_INPUT, "qwen3 generate: batch exceeds max_batch");
        B2A_CHECK(L >= 1 && gp->max_tokens >= 1 && L + gp->max_tokens <= h->cfg.max_context, B2A_ERR_INVALID_INPUT,
                  "qwen3 generate: prompt + max_tokens exceeds max_context");
        B2A_CHECK(n_trailing_max >= 0 && (n_trailing_max == 0 || (trailing_text_hidden && n_trailing)), B2A_ERR_INVALID_INPUT,
                  "qwen3 generate: bad trailing text");
        B2A_CHECK(gp->top_p > 0.f && gp->repetition_penalty > 0.f, B2A_ERR_INVALID_INPUT, "qwen3 generate: bad sampling parameters");
        B2A_CUDA(cudaSetDevice(h->device));
        cudaStream_t s = h->stream;
        h->cancel.store(0);
        const int Hh = h->H(), G = h->G(), MT = gp->max_tokens, nmax = std::max(1, n_trailing_max);
        h->talker->set_batch(B); h->pred->set_batch(B);
        h->talker->drop_graphs();
        h->embeds.upload(input_embeds, (size_t)B * L * Hh, s);
        h->trailing.alloc((size_t)B * nmax * Hh);
        std::vector<int> nt(8, 0);
        if (n_trailing_max > 0) {
            B2A_CUDA(cudaMemcpyAsync(h->trailing.p, trailing_text_hidden, (size_t)B * n_trailing_max * Hh * sizeof(float), cudaMemcpyHostToDevice, s));
            for (int b = 0; b < B; ++b) {
                B2A_CHECK(n_trailing[b] >= 0 && n_trailing[b] <= n_trailing_max, B2A_ERR_INVALID_INPUT, "qwen3 generate: n_trailing out of range");
                nt[b] = n_trailing[b];
            }
        }
        h->n_trailing.upload(nt.data(), 8, s);
        h->pad.upload(tts_pad_embed, Hh, s);
        h->out_codes.alloc((size_t)B * MT * G);
        B2A_CUDA(cudaStreamSynchronize(s));       // nt goes out of scope with the lambda only, but keep uploads ordered before capture
        h->capture_frame(B, *gp, MT, nmax);
        const double t0 = now_s();
        q3_init_rows_kernel<<<1, 256, 0, s>>>(B, L, h->talker->pos.p, h->row_frame.p, h->n_frames.p, h->done.p, h->n_active.p, h->seen.p,
                                              cdiv(h->cfg.vocab_size, 32));
        count_launch();
        h->prefill(B, L, s);
        B2A_CUDA(cudaStreamSynchronize(s));
        const double t1 = now_s();
        int steps = 0, emitted = 0;
        bool cancelled = false;
        std::vector<int> hn(8), hc;
        auto emit = [&]() {          // frames emitted since the last poll -> on_frame
            B2A_CUDA(cudaMemcpy(hn.data(), h->n_frames.p, 8 * sizeof(int), cudaMemcpyDeviceToHost));
            hc.resize((size_t)B * MT * G);
            B2A_CUDA(cudaMemcpy(hc.data(), h->out_codes.p, hc.size() * sizeof(int), cudaMemcpyDeviceToHost));
            for (int b = 0; b < B; ++b)
                if (std::min(hn[b], MT) > emitted) on_frame(user, b, emitted, hc.data() + ((size_t)b * MT + emitted) * G);
            ++emitted;
        };
        while (steps < MT) {
            const int burst = on_frame ? 1 : std::min(4, MT - steps);
            for (int i = 0; i < burst; ++i) { B2A_CUDA(cudaGraphLaunch(h->g_frame, s)); count_launch(h->launches_frame); }
            steps += burst;
            B2A_CUDA(cudaMemcpyAsync(h->h_flag.p, h->n_active.p, sizeof(int), cudaMemcpyDeviceToHost, s));
            B2A_CUDA(cudaStreamSynchronize(s));
            if (on_frame) emit();
            if (h->cancel.load()) { cancelled = true; break; }
            if (h->h_flag.p[0] <= 0) break;
        }
        const double t2 = now_s();
        B2A_CHECK(!cancelled, B2A_ERR_CANCELLED, "generation cancelled");
        B2A_CUDA(cudaMemcpy(hn.data(), h->n_frames.p, 8 * sizeof(int), cudaMemcpyDeviceToHost));
        B2A_CUDA(cudaMemcpy(codes_out, h->out_codes.p, (size_t)B * MT * G * sizeof(int), cudaMemcpyDeviceToHost));
        int total = 0;
        for (int b = 0; b < B; ++b) { n_frames_out[b] = std::min(hn[b], MT); total += n_frames_out[b]; }
        if (info) {
            info->prompt_token_count = L;
            info->generation_token_count = total;
            info->prefill_time = t1 - t0;
            info->generate_time = t2 - t1;
            info->tokens_per_second = total / std::max(1e-9, t2 - t1);
            info->codec_time = 0;
            size_t fr = 0, tot = 0;
            cudaMemGetInfo(&fr, &tot);
            info->peak_memory_gb = (double)(tot - fr) / 1e9;
        }
    });
}

}  // extern "C"
