// Qwen3-TTS speech-tokenizer DECODER for sm_100a (12.5 Hz codes -> 24 kHz waveform).
// Replaces (reference paths, file = Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSSpeechTokenizer.swift):
//   :9-121      split residual vector quantizer decode (usage-normalised Euclidean codebooks, k1 output projections)
//   :135-232    CausalConv1d (+ streaming step), :257-297 ConvNeXtBlock, :301-491 DecoderTransformer (KV cache)
//   :495-638    DecoderResidualUnit / DecoderBlockUpsample / DecoderBlock, :641-731 initial / output convs, SnakeBeta
//   :888-1025   Qwen3TTSSpeechTokenizerDecoder: callAsFunction, streamingStep, chunkedDecode
//   :1070-1092  Qwen3TTSSpeechTokenizer.streamingDecode
// EXPERIMENTAL -- written against oracle/qwen3_tts_codec.py but NOT yet run on a GPU (no GPU time was left in the
// round that added it); its parity tests are gated behind B2A_EXPERIMENTAL_N1=1.  Nothing else in the library calls it.
//
// Design.  Every chunk of code frames is decoded by the streaming step with carried state; the one-shot call is the
// same step after a reset (zero state == the reference's causal zero padding).  All dense layers -- RVQ projections,
// k3 / k7 / dilated causal convolutions, transposed convolutions (as phase-major causal convolutions), transformer
// and ConvNeXt linears -- run on the implicit-GEMM tcgen05 kernel of implicit_conv.cuh over planar bf16 hi/lo
// activations; its epilogue fuses bias, layer scale / gamma, exact GELU, the residual add, the next layer's SnakeBeta
// and the hi/lo split, and writes behind the H history frames the consumer's taps reach back over.  Small SIMT
// kernels cover the gathers, norms, RoPE + attention over the cache, the depthwise conv + LayerNorm and the 1-channel
// output conv.  The reference's double-counted transposed-conv bias at chunk boundaries (see the oracle's header) is
// reproduced (Args::bias_twice_t0).
#include "common.cuh"
#include "implicit_conv.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace b2a {
namespace st {

typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------------- tensor maps
static tc::EncodeTiledFn encode_fn() {
    static tc::EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        B2A_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
        B2A_CHECK(p && q == cudaDriverEntryPointSuccess, B2A_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (tc::EncodeTiledFn)p;
    }
    return fn;
}

// planar hi/lo activations [2][B][Ttot][C] bf16 -> rank-4 map {C, Ttot, B, 2}, box {64, 64, 1, 2}, 128-byte swizzle
static CUtensorMap make_tmap_planes(const bf16* base, int C, long long Ttot, int B, int f16) {
    B2A_CHECK(C % 8 == 0 && ((uintptr_t)base & 15) == 0 && Ttot >= 1 && B >= 1, B2A_ERR_INVALID_INPUT,
              "TMA: activation planes must be 16-byte aligned with channels % 8 == 0");
    CUtensorMap m;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Ttot, (cuuint64_t)B, 2};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)Ttot * C * 2, (cuuint64_t)B * Ttot * C * 2};
    const cuuint32_t box[4] = {(cuuint32_t)tc::BK, (cuuint32_t)ic::HALF, 1, 2};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = encode_fn()(&m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(base), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2A_CHECK(r == CUDA_SUCCESS, B2A_ERR_CUDA, "cuTensorMapEncodeTiled (rank 4) failed (" + std::to_string((int)r) + ")");
    return m;
}

// weights [rows, cols] 16-bit row-major, {64, 128} box, 128-byte swizzle (tc::make_tmap_bf16 with a selectable element type)
static CUtensorMap make_tmap_w16(const void* base, long long rows, long long cols, int f16) {
    B2A_CHECK(cols % 8 == 0 && ((uintptr_t)base & 15) == 0, B2A_ERR_INVALID_INPUT, "TMA: tensor must be 16-byte aligned with cols % 8 == 0");
    CUtensorMap m;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)tc::BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode_fn()(&m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B2A_CHECK(r == CUDA_SUCCESS, B2A_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return m;
}
// host-side hi / lo split in the operand format, as raw 16-bit words
static inline void split16(float v, int f16, uint16_t& hi, uint16_t& lo) {
    if (f16) {
        const __half h = __float2half_rn(v);
        hi = __half_as_ushort(h);
        lo = __half_as_ushort(__float2half_rn(v - __half2float(h)));
    } else {
        const bf16 h = __float2bfloat16_rn(v);
        hi = __bfloat16_as_ushort(h);
        lo = __bfloat16_as_ushort(__float2bfloat16_rn(v - __bfloat162float(h)));
    }
}
static inline float join16(uint16_t hi, uint16_t lo, int f16) {
    return f16 ? __half2float(__ushort_as_half(hi)) + __half2float(__ushort_as_half(lo))
               : __bfloat162float(__ushort_as_bfloat16(hi)) + __bfloat162float(__ushort_as_bfloat16(lo));
}

// ----------------------------------------------------------------------------------------------- SIMT kernels
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = wsum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < THREADS / 32; ++i) t += red[i];
    return t;
}
__device__ __forceinline__ void put_planes(bf16* base, long long plane, long long idx, float v, int f16) {
    ic::put_hilo16(reinterpret_cast<uint16_t*>(base), plane, idx, v, f16);
}

// codes [B, nq, T] -> planes [2][B*T][2*D2]: channels [0, D2) = sum of the semantic codebooks, [D2, 2*D2) = sum of the rest
__global__ void rvq_gather_kernel(const int* __restrict__ codes, const float* __restrict__ emb /*[nq][bins][D2]*/, bf16* __restrict__ out,
                                  int B, int T, int nq, int nq_model, int nsem, int bins, int D2, int f16) {
    const long long n = blockIdx.x;
    const int b = (int)(n / T), t = (int)(n - (long long)b * T);
    const long long plane = (long long)B * T * 2 * D2;
    for (int c = threadIdx.x; c < D2; c += blockDim.x) {
        float s0 = 0.f, s1 = 0.f;
        for (int qi = 0; qi < nq && qi < nq_model; ++qi) {
            int code = codes[((long long)b * nq + qi) * T + t];
            code = code < 0 ? 0 : (code >= bins ? bins - 1 : code);
            const float v = emb[((long long)qi * bins + code) * D2 + c];
            if (qi < nsem) s0 += v; else s1 += v;
        }
        put_planes(out, plane, n * 2 * D2 + c, s0, f16);
        put_planes(out, plane, n * 2 * D2 + D2 + c, s1, f16);
    }
}

// RMSNorm over channels -> planes [2][N][C]   (DecoderRMSNorm :301-314: w * (x * rsqrt(mean(x^2) + eps)))
constexpr int RN_THREADS = 128;
__global__ void __launch_bounds__(RN_THREADS)
rmsnorm_planes_kernel(const float* __restrict__ x, const float* __restrict__ w, bf16* __restrict__ out, long long N, int C, float eps, int f16) {
    __shared__ float red[RN_THREADS / 32];
    const long long n = blockIdx.x;
    float ss = 0.f;
    for (int c = threadIdx.x; c < C; c += RN_THREADS) { const float v = x[n * C + c]; ss += v * v; }
    const float r = rsqrtf(block_sum<RN_THREADS>(ss, red) / (float)C + eps);
    for (int c = threadIdx.x; c < C; c += RN_THREADS) put_planes(out, N * C, n * C + c, w[c] * (x[n * C + c] * r), f16);
}

// rotate-half RoPE on q (in place) and k (into the cache), v copied into the cache.  qkv [N, (nh + 2 nkv) * hd] fp32.
__global__ void rope_cache_kernel(float* __restrict__ qkv, float* __restrict__ Kc, float* __restrict__ Vc, const float* __restrict__ inv_freq,
                                  int T, int pos0, int nh, int nkv, int hd, int cap) {
    const long long n = blockIdx.x;
    const int b = (int)(n / T), t = (int)(n - (long long)b * T);
    const int pos = pos0 + t, half = hd / 2, ld = (nh + 2 * nkv) * hd;
    float* row = qkv + n * ld;
    for (int idx = threadIdx.x; idx < (nh + nkv) * half; idx += blockDim.x) {
        const int head = idx / half, i = idx - head * half;
        float sn, cs;
        sincosf((float)pos * inv_freq[i], &sn, &cs);
        const float x1 = row[head * hd + i], x2 = row[head * hd + i + half];
        const float o1 = x1 * cs - x2 * sn, o2 = x2 * cs + x1 * sn;
        if (head < nh) {
            row[head * hd + i] = o1;
            row[head * hd + i + half] = o2;
        } else {
            float* dst = Kc + (((long long)b * nkv + (head - nh)) * cap + pos) * hd;
            dst[i] = o1;
            dst[i + half] = o2;
        }
    }
    for (int idx = threadIdx.x; idx < nkv * hd; idx += blockDim.x) {
        const int kvh = idx / hd, d = idx - kvh * hd;
        Vc[(((long long)b * nkv + kvh) * cap + pos) * hd + d] = row[(nh + nkv) * hd + idx];
    }
}

// causal attention of the chunk's T queries over cache positions [0, pos0 + t]: one warp per (query, head), each lane
// owns hd / 32 consecutive dims, online softmax, four keys in flight.  Output -> planes [2][N][nh * hd].
constexpr int AT_WARPS = 4;
template <int DPL>
__global__ void __launch_bounds__(AT_WARPS * 32)
attn_kernel(const float* __restrict__ qkv, const float* __restrict__ Kc, const float* __restrict__ Vc, bf16* __restrict__ out,
            int B, int T, int pos0, int nh, int nkv, int cap, float scale, int f16) {
    constexpr int HD = DPL * 32;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * AT_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
    if (t >= T) return;
    const long long n = (long long)b * T + t;
    const int ld = (nh + 2 * nkv) * HD, kvh = h / (nh / nkv);
    float q[DPL], acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; ++d) { q[d] = qkv[n * ld + h * HD + lane * DPL + d] * scale; acc[d] = 0.f; }
    const float* Kb = Kc + ((long long)b * nkv + kvh) * cap * HD + lane * DPL;
    const float* Vb = Vc + ((long long)b * nkv + kvh) * cap * HD + lane * DPL;
    float m = -INFINITY, l = 0.f;
    const int nkeys = pos0 + t + 1;
    for (int p0 = 0; p0 < nkeys; p0 += 4) {
        float s[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float d0 = 0.f;
            if (p0 + u < nkeys) {
#pragma unroll
                for (int d = 0; d < DPL; ++d) d0 = fmaf(q[d], Kb[(long long)(p0 + u) * HD + d], d0);
            }
            s[u] = d0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) s[u] = wsum(s[u]);
        float mx = m;
#pragma unroll
        for (int u = 0; u < 4; ++u) if (p0 + u < nkeys) mx = fmaxf(mx, s[u]);
        const float corr = __expf(m - mx);      // m = -inf on the first pass: exp(-inf) = 0
        l *= corr;
#pragma unroll
        for (int d = 0; d < DPL; ++d) acc[d] *= corr;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (p0 + u < nkeys) {
                const float e = __expf(s[u] - mx);
                l += e;
#pragma unroll
                for (int d = 0; d < DPL; ++d) acc[d] = fmaf(e, Vb[(long long)(p0 + u) * HD + d], acc[d]);
            }
        }
        m = mx;
    }
    const float inv = 1.0f / l;
    const long long N = (long long)B * T;
#pragma unroll
    for (int d = 0; d < DPL; ++d) put_planes(out, N * nh * HD, n * nh * HD + h * HD + lane * DPL + d, acc[d] * inv, f16);
}

// gu [N, 2I] (gate | up) -> silu(gate) * up as planes [2][N][I]     (DecoderMLP :412-414)
__global__ void swiglu_planes_kernel(const float* __restrict__ gu, bf16* __restrict__ out, long long N, int I, int f16) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N * I) return;
    const long long n = i / I;
    const int c = (int)(i - n * I);
    const float g = gu[n * 2 * I + c], u = gu[n * 2 * I + I + c];
    put_planes(out, N * I, i, g / (1.0f + __expf(-g)) * u, f16);
}

// causal depthwise conv (k taps, history from `st` [B, k-1, C]) -> LayerNorm(eps) -> planes [2][B*T][C]
constexpr int DL_THREADS = 256, DL_MAXV = 4;    // channels <= 1024
__global__ void __launch_bounds__(DL_THREADS)
dw_ln_kernel(const float* __restrict__ x, const float* __restrict__ st, const float* __restrict__ dw_w /*[C, k]*/, const float* __restrict__ dw_b,
             const float* __restrict__ ln_w, const float* __restrict__ ln_b, bf16* __restrict__ out, int B, int T, int C, int k, float eps, int f16) {
    __shared__ float red[DL_THREADS / 32];
    const long long n = blockIdx.x;
    const int b = (int)(n / T), t = (int)(n - (long long)b * T), H = k - 1;
    float v[DL_MAXV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < DL_MAXV; ++j) {
        const int c = threadIdx.x + j * DL_THREADS;
        float val = 0.f;
        if (c < C) {
            val = dw_b[c];
            for (int kk = 0; kk < k; ++kk) {
                const int ti = t - H + kk;
                const float xin = ti >= 0 ? x[((long long)b * T + ti) * C + c] : st[((long long)b * H + (H + ti)) * C + c];
                val = fmaf(dw_w[c * k + kk], xin, val);
            }
        }
        v[j] = val;
        s += val;
    }
    const float mean = block_sum<DL_THREADS>(s, red) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < DL_MAXV; ++j) {
        const int c = threadIdx.x + j * DL_THREADS;
        if (c < C) { const float d = v[j] - mean; q += d * d; }
    }
    const float r = rsqrtf(block_sum<DL_THREADS>(q, red) / (float)C + eps);
    const long long N = (long long)B * T;
#pragma unroll
    for (int j = 0; j < DL_MAXV; ++j) {
        const int c = threadIdx.x + j * DL_THREADS;
        if (c < C) put_planes(out, N * C, n * C + c, (v[j] - mean) * r * ln_w[c] + ln_b[c], f16);
    }
}

// fp32 history: new[b][f] = last H frames of [old | x]
__global__ void state_update_f32_kernel(const float* __restrict__ x, const float* __restrict__ old, float* __restrict__ nw, int B, int T, int H, int C) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * H * C) return;
    const int c = (int)(i % C);
    const long long bf = i / C;
    const int f = (int)(bf % H), b = (int)(bf / H);
    const int src = T + f;        // frame index in [old (H) | x (T)]
    nw[i] = src < H ? old[((long long)b * H + src) * C + c] : x[((long long)b * T + (src - H)) * C + c];
}

// bf16 planes with a history prefix: X = [2][B][H + T][C].  Copies the old state into frames [0, H) and saves the last H
// frames of [old | new] as the new state (old and new are different buffers).  8 channels (16 bytes) per thread.
__global__ void carry_planes_kernel(bf16* __restrict__ X, const bf16* __restrict__ old, bf16* __restrict__ nw, int B, int T, int H, int C8) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_plane = (long long)B * H * C8;
    if (i >= 2 * per_plane) return;
    const int c = (int)(i % C8);
    long long r = i / C8;
    const int f = (int)(r % H); r /= H;
    const int b = (int)(r % B), p = (int)(r / B);
    const uint4* o4 = reinterpret_cast<const uint4*>(old);
    uint4* n4 = reinterpret_cast<uint4*>(nw);
    uint4* x4 = reinterpret_cast<uint4*>(X);
    const long long xrow = ((long long)p * B + b) * (H + T);
    const uint4 ov = o4[i];
    x4[(xrow + f) * C8 + c] = ov;
    const int src = T + f;
    n4[i] = src < H ? o4[(((long long)p * B + b) * H + src) * C8 + c] : x4[(xrow + src) * C8 + c];
}

// SnakeBeta -> causal k-tap conv to ONE channel -> clip(-1, 1)          (DecoderOutputSnake + DecoderOutputConv :693-731, :946)
// x [B, T, C] fp32 (raw, pre-activation), st [B, k-1, C] raw history.  64 outputs per CTA; the activated tile lives in smem.
constexpr int FC_TILE = 64, FC_THREADS = 128, FC_MAXK = 8;
__global__ void __launch_bounds__(FC_THREADS)
final_conv_kernel(const float* __restrict__ x, const float* __restrict__ st, const float* __restrict__ sa, const float* __restrict__ sb,
                  const float* __restrict__ w /*[k, C]*/, float bias, float* __restrict__ wave, int T, int C, int k) {
    extern __shared__ float fsm[];
    const int H = k - 1, rows = FC_TILE + H, ldc = C + 1;
    float* tile = fsm;                    // [rows][C + 1]
    float* wk = fsm + rows * ldc;         // [k][C]
    float* red = wk + k * C;              // [FC_THREADS]
    const int b = blockIdx.y, t0 = blockIdx.x * FC_TILE;
    for (int i = threadIdx.x; i < k * C; i += FC_THREADS) wk[i] = w[i];
    for (int i = threadIdx.x; i < rows * C; i += FC_THREADS) {
        const int rr = i / C, c = i - rr * C;
        const int ti = t0 - H + rr;
        float v = 0.f;
        if (ti >= 0) { if (ti < T) v = x[((long long)b * T + ti) * C + c]; }
        else v = st[((long long)b * H + (H + ti)) * C + c];
        tile[rr * ldc + c] = ic::snake_beta(v, sa[c], sb[c]);
    }
    __syncthreads();
    const int o = threadIdx.x & (FC_TILE - 1), part = threadIdx.x / FC_TILE;      // two threads per output, channels split in halves
    const int cbeg = part * ((C + 1) / 2), cend = min(C, cbeg + (C + 1) / 2);
    float acc = 0.f;
    for (int kk = 0; kk < k; ++kk)
        for (int c = cbeg; c < cend; ++c) acc = fmaf(wk[kk * C + c], tile[(o + kk) * ldc + c], acc);
    red[threadIdx.x] = acc;
    __syncthreads();
    if (part == 0 && t0 + o < T) {
        const float y = red[o] + red[o + FC_TILE] + bias;
        wave[(long long)b * T + t0 + o] = fminf(1.0f, fmaxf(-1.0f, y));
    }
}

// ----------------------------------------------------------------------------------------------- weights
// k-blocks per accumulation segment of the implicit convolution (ic::Args::seg_kb); B2A_ST_SEG=0 restores one accumulator per tile
static int seg_kb_default() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("B2A_ST_SEG"); v = e ? std::max(0, atoi(e)) : 4; }
    return v;
}

struct IW {                       // implicit-conv weight: [M][taps][cblocks * 64] as bf16 hi / lo, K-major
    DBuf<bf16> hi, lo;
    DBuf<float> bias, rscale;     // rscale[m] (fp16 operands only): 1 / the power of two row m is stored times
    CUtensorMap th{}, tl{};
    int M = 0, taps = 1, cblocks = 0, Cin = 0;
    bool has_bias = false;
    // [M][taps][Cin] -> [M][taps][cblocks * 64], every tap's channel run zero-padded to whole 64-channel k-blocks
    static std::vector<float> pad_k(const std::vector<float>& W, int M, int taps, int Cin) {
        const int cb = cdiv(Cin, tc::BK);
        const size_t K = (size_t)taps * cb * tc::BK;
        std::vector<float> g((size_t)M * K, 0.f);
        for (int m = 0; m < M; ++m)
            for (int j = 0; j < taps; ++j)
                memcpy(&g[(size_t)m * K + (size_t)j * cb * tc::BK], &W[((size_t)m * taps + j) * Cin], (size_t)Cin * sizeof(float));
        return g;
    }
    void build(const std::vector<float>& W /*[M][taps][Cin]*/, int M_, int taps_, int Cin_, int f16) {
        M = M_; taps = taps_; Cin = Cin_; cblocks = cdiv(Cin, tc::BK);
        const size_t K = (size_t)taps * cblocks * tc::BK;
        std::vector<float> g = pad_k(W, M, taps, Cin);
        if (f16) {
            // fp16 pairs: a weight of magnitude 0.01 has a SUBNORMAL lo half (|lo| < 2^-11 |w| < 6.1e-5), i.e. ~18 bits instead of 22.
            // Store row m times 2^e with max|w_m| * 2^e in [8192, 16384) and undo the (exact) scaling in the epilogue.
            std::vector<float> rs((size_t)M, 1.f);
            for (int m = 0; m < M; ++m) {
                float mx = 0.f;
                for (size_t k = 0; k < K; ++k) mx = std::max(mx, fabsf(g[(size_t)m * K + k]));
                if (mx > 0.f && std::isfinite(mx)) {
                    int e = 0;
                    frexpf(mx, &e);                              // mx = f * 2^e, f in [0.5, 1)
                    const float sc = ldexpf(1.0f, 14 - e);       // mx * sc in [8192, 16384)
                    for (size_t k = 0; k < K; ++k) g[(size_t)m * K + k] *= sc;
                    rs[m] = 1.0f / sc;
                }
            }
            rscale.upload(rs.data(), rs.size());
        }
        std::vector<uint16_t> h(g.size()), l(g.size());
        for (size_t i = 0; i < g.size(); ++i) split16(g[i], f16, h[i], l[i]);
        hi.upload(reinterpret_cast<const bf16*>(h.data()), h.size());
        lo.upload(reinterpret_cast<const bf16*>(l.data()), l.size());
        B2A_CUDA(cudaDeviceSynchronize());
        th = make_tmap_w16(hi.p, M, (long long)K, f16);
        tl = make_tmap_w16(lo.p, M, (long long)K, f16);
    }
    void set_bias(const std::vector<float>& b) { bias.upload(b.data(), b.size()); has_bias = true; B2A_CUDA(cudaDeviceSynchronize()); }
};

struct Snake { DBuf<float> a, ib; };                 // a = exp(alpha), ib = 1 / (exp(beta) + 1e-9)
struct PlaneState { DBuf<bf16> s[2]; int H = 0, C = 0; };      // [2][B][H][C]
struct F32State { DBuf<float> s[2]; int H = 0, C = 0; };       // [B][H][C]

struct TLayer { IW qkv, o, gu, down; DBuf<float> ln1, ln2, sc_attn, sc_mlp; DBuf<float> K, V; };
struct UpLayer { IW ct, pw1, pw2; DBuf<float> dw_w, dw_b, ln_w, ln_b, gamma; F32State st; int factor = 1; };
struct ResUnit { Snake a1, a2; IW c1, c2; PlaneState st; int dil = 1; };
struct DecBlock { Snake sn; IW ct; PlaneState st; ResUnit ru[3]; int rate = 1, cin = 0, cout = 0; };

}  // namespace st
}  // namespace b2a

using namespace b2a;
using namespace b2a::st;

struct b2a_speech_tokenizer {
    int device;
    b2a_speech_tokenizer_config cfg;
    cudaStream_t stream = nullptr;
    int num_sms = 148;
    int total_up = 1, D2 = 0, Mqkv = 0;
    int use_f16 = 1;              // fp16 hi/lo operand pairs (22 mantissa bits, saturating at 65504); B2A_ST_FP16=0: bf16 pairs (16 bits, fp32 range)
    // weights
    DBuf<float> emb;              // [nq][bins][D2] usage-normalised codebooks
    IW rvq_proj, pre_conv, in_proj, out_proj, dec0;
    PlaneState st_pre, st_dec0;
    DBuf<float> final_norm, inv_freq;
    std::vector<TLayer> layers;
    std::vector<UpLayer> ups;
    std::vector<DecBlock> blocks;
    Snake out_snake;
    DBuf<float> out_w;            // [k][C]
    float out_b = 0.f;
    int out_k = 7;
    F32State st_out;
    // streaming state
    int parity = 0, chunk_idx = 0, cache_len = 0, stream_B = 0;
    // workspace
    DBuf<int> d_codes;
    DBuf<bf16> P0, P1;
    DBuf<float> Xh, Xc, Q, wave;
    // diagnostics (b2a_speech_tokenizer_debug_stage): a copy of the fp32 activation tensor after stage `dbg_stage`
    int dbg_stage = -1;
    long long dbg_n = 0;
    DBuf<float> dbg;
    void dbg_tap(int stage, const float* src, long long n, cudaStream_t s) {
        if (stage != dbg_stage) return;
        dbg.alloc((size_t)n);
        B2A_CUDA(cudaMemcpyAsync(dbg.p, src, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, s));
        dbg_n = n;
    }

    ~b2a_speech_tokenizer() { if (stream) cudaStreamDestroy(stream); }

    static std::vector<float> conv_w(const TensorTable& tt, const std::string& name, int out, int k, int in) {
        return tt.f32(name, (int64_t)out * k * in);       // MLX [out, k, in] == [M][taps][Cin] with tap j <-> kernel index j
    }
    // transposed conv, MLX [out, k, in], k = n * r: rows m = rho * out + co, tap j <-> input frame q - (n - 1 - j) <-> kernel index rho + (n - 1 - j) * r
    static std::vector<float> convt_w(const std::vector<float>& w, int out, int k, int in, int r) {
        const int n = k / r;
        std::vector<float> g((size_t)r * out * n * in);
        for (int rho = 0; rho < r; ++rho)
            for (int co = 0; co < out; ++co)
                for (int j = 0; j < n; ++j)
                    memcpy(&g[(((size_t)rho * out + co) * n + j) * in], &w[((size_t)co * k + rho + (size_t)(n - 1 - j) * r) * in], (size_t)in * sizeof(float));
        return g;
    }
    static void up(DBuf<float>& d, const std::vector<float>& v) { d.upload(v.data(), v.size()); }
    static void load_snake(Snake& s, const TensorTable& tt, const std::string& p, int C) {
        std::vector<float> al = tt.f32(p + ".alpha", C), be = tt.f32(p + ".beta", C), a(C), ib(C);
        for (int c = 0; c < C; ++c) { a[c] = expf(al[c]); ib[c] = 1.0f / (expf(be[c]) + 1e-9f); }
        up(s.a, a); up(s.ib, ib);
    }
    void load_conv(IW& w, const TensorTable& tt, const std::string& p, int out, int k, int in, bool bias = true) {
        w.build(conv_w(tt, p + ".weight", out, k, in), out, k, in, use_f16);
        if (bias) w.set_bias(tt.f32(p + ".bias", out));
    }
    void load_linear(IW& w, const TensorTable& tt, const std::string& p, int out, int in, bool bias) {
        w.build(tt.f32(p + ".weight", (int64_t)out * in), out, 1, in, use_f16);
        if (bias) w.set_bias(tt.f32(p + ".bias", out));
    }

    b2a_speech_tokenizer(int dev, const b2a_speech_tokenizer_config& c, const TensorTable& tt) : device(dev), cfg(c) {
        B2A_CHECK(c.codebook_dim % 16 == 0 && c.latent_dim % 8 == 0 && c.hidden_size % 8 == 0 && c.intermediate_size % 8 == 0,
                  B2A_ERR_INVALID_INPUT, "speech tokenizer: channel counts must be multiples of 8 (codebook_dim of 16)");
        B2A_CHECK(c.head_dim == 32 || c.head_dim == 64 || c.head_dim == 128, B2A_ERR_INVALID_INPUT, "speech tokenizer: head_dim must be 32, 64 or 128");
        B2A_CHECK(c.num_attention_heads >= 1 && c.num_key_value_heads >= 1 && c.num_attention_heads % c.num_key_value_heads == 0,
                  B2A_ERR_INVALID_INPUT, "speech tokenizer: bad head counts");
        B2A_CHECK(c.num_upsample_rates >= 1 && c.num_upsample_rates <= 8 && c.num_upsampling_ratios >= 1 && c.num_upsampling_ratios <= 8,
                  B2A_ERR_INVALID_INPUT, "speech tokenizer: bad upsample lists");
        B2A_CHECK(c.num_quantizers >= 1 && c.num_semantic_quantizers >= 1 && c.num_semantic_quantizers <= c.num_quantizers && c.codebook_size >= 1,
                  B2A_ERR_INVALID_INPUT, "speech tokenizer: bad quantizer counts");
        B2A_CHECK(c.latent_dim <= DL_THREADS * DL_MAXV, B2A_ERR_INVALID_INPUT, "speech tokenizer: latent_dim must be <= 1024");
        B2A_CHECK((c.decoder_dim >> c.num_upsample_rates) >= 8 && (c.decoder_dim >> c.num_upsample_rates) % 8 == 0 &&
                      (c.decoder_dim >> c.num_upsample_rates) <= 128,
                  B2A_ERR_INVALID_INPUT, "speech tokenizer: decoder_dim / 2^blocks must be a multiple of 8 and <= 128");
        B2A_CHECK(c.max_batch >= 1 && c.max_cache_frames >= 1, B2A_ERR_INVALID_INPUT, "speech tokenizer: max_batch / max_cache_frames must be positive");
        require_device(dev);
        { const char* e = getenv("B2A_ST_FP16"); use_f16 = (e && e[0] == '0') ? 0 : 1; }
        B2A_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        B2A_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
        B2A_CUDA(cudaFuncSetAttribute(ic::implicit_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ic::SMEM_BYTES));
        const int nq = c.num_quantizers, ns = c.num_semantic_quantizers, bins = c.codebook_size, cbd = c.codebook_dim, L = c.latent_dim,
                  Hd = c.hidden_size, I = c.intermediate_size, nh = c.num_attention_heads, nkv = c.num_key_value_heads, hd = c.head_dim;
        D2 = cbd / 2;
        Mqkv = (nh + 2 * nkv) * hd;
        // codebooks: embedding = embedding_sum / max(cluster_usage, 1e-5)      (Quantization.swift:29-33)
        {
            std::vector<float> e((size_t)nq * bins * D2);
            for (int qi = 0; qi < nq; ++qi) {
                const std::string p = qi < ns ? "quantizer.rvq_first.vq.layers." + std::to_string(qi) : "quantizer.rvq_rest.vq.layers." + std::to_string(qi - ns);
                std::vector<float> sum = tt.f32(p + ".codebook.embedding_sum", (int64_t)bins * D2), use = tt.f32(p + ".codebook.cluster_usage", bins);
                for (int r = 0; r < bins; ++r) {
                    const float u = std::max(use[r], 1e-5f);
                    for (int d = 0; d < D2; ++d) e[((size_t)qi * bins + r) * D2 + d] = sum[(size_t)r * D2 + d] / u;
                }
            }
            up(emb, e);
            // both k1 output projections as one [cbd, 2 * D2] matrix over the concatenated (semantic | rest) sums
            std::vector<float> w1 = tt.f32("quantizer.rvq_first.output_proj.weight", (int64_t)cbd * D2), w((size_t)cbd * 2 * D2, 0.f);
            std::vector<float> w2 = nq > ns ? tt.f32("quantizer.rvq_rest.output_proj.weight", (int64_t)cbd * D2) : std::vector<float>((size_t)cbd * D2, 0.f);
            for (int o = 0; o < cbd; ++o) {
                memcpy(&w[(size_t)o * 2 * D2], &w1[(size_t)o * D2], (size_t)D2 * sizeof(float));
                memcpy(&w[(size_t)o * 2 * D2 + D2], &w2[(size_t)o * D2], (size_t)D2 * sizeof(float));
            }
            rvq_proj.build(w, cbd, 1, 2 * D2, use_f16);
        }
        load_conv(pre_conv, tt, "pre_conv.conv", L, 3, cbd);
        st_pre.H = 2; st_pre.C = cbd;
        load_linear(in_proj, tt, "pre_transformer.input_proj", Hd, L, true);
        load_linear(out_proj, tt, "pre_transformer.output_proj", L, Hd, true);
        up(final_norm, tt.f32("pre_transformer.norm.weight", Hd));
        {
            std::vector<float> f(hd / 2);
            for (int i = 0; i < hd / 2; ++i) f[i] = 1.0f / powf(c.rope_theta, (float)(2 * i) / (float)hd);    // :330-332
            up(inv_freq, f);
        }
        layers.resize(c.num_hidden_layers);
        for (int l = 0; l < c.num_hidden_layers; ++l) {
            const std::string p = "pre_transformer.layers." + std::to_string(l) + ".";
            TLayer& T = layers[l];
            const bool ab = c.attention_bias != 0;
            {
                std::vector<float> w((size_t)Mqkv * Hd), q = tt.f32(p + "self_attn.q_proj.weight", (int64_t)nh * hd * Hd),
                                   k = tt.f32(p + "self_attn.k_proj.weight", (int64_t)nkv * hd * Hd), v = tt.f32(p + "self_attn.v_proj.weight", (int64_t)nkv * hd * Hd);
                memcpy(w.data(), q.data(), q.size() * sizeof(float));
                memcpy(w.data() + q.size(), k.data(), k.size() * sizeof(float));
                memcpy(w.data() + q.size() + k.size(), v.data(), v.size() * sizeof(float));
                T.qkv.build(w, Mqkv, 1, Hd, use_f16);
                if (ab) {
                    std::vector<float> bq = tt.f32(p + "self_attn.q_proj.bias", nh * hd), bk = tt.f32(p + "self_attn.k_proj.bias", nkv * hd),
                                       bv = tt.f32(p + "self_attn.v_proj.bias", nkv * hd);
                    bq.insert(bq.end(), bk.begin(), bk.end());
                    bq.insert(bq.end(), bv.begin(), bv.end());
                    T.qkv.set_bias(bq);
                }
            }
            load_linear(T.o, tt, p + "self_attn.o_proj", Hd, nh * hd, ab);
            {
                std::vector<float> g = tt.f32(p + "mlp.gate_proj.weight", (int64_t)I * Hd), u = tt.f32(p + "mlp.up_proj.weight", (int64_t)I * Hd);
                g.insert(g.end(), u.begin(), u.end());
                T.gu.build(g, 2 * I, 1, Hd, use_f16);
            }
            load_linear(T.down, tt, p + "mlp.down_proj", Hd, I, false);
            up(T.ln1, tt.f32(p + "input_layernorm.weight", Hd));
            up(T.ln2, tt.f32(p + "post_attention_layernorm.weight", Hd));
            up(T.sc_attn, tt.f32(p + "self_attn_layer_scale.scale", Hd));
            up(T.sc_mlp, tt.f32(p + "mlp_layer_scale.scale", Hd));
        }
        total_up = 1;
        ups.resize(c.num_upsampling_ratios);
        for (int i = 0; i < c.num_upsampling_ratios; ++i) {
            const std::string p = "upsample." + std::to_string(i) + ".layers.";
            UpLayer& U = ups[i];
            U.factor = c.upsampling_ratios[i];
            B2A_CHECK(U.factor >= 1 && U.factor <= 16, B2A_ERR_INVALID_INPUT, "speech tokenizer: bad upsampling ratio");
            total_up *= U.factor;
            U.ct.build(convt_w(conv_w(tt, p + "0.conv.weight", L, U.factor, L), L, U.factor, L, U.factor), U.factor * L, 1, L, use_f16);
            U.ct.set_bias(tt.f32(p + "0.conv.bias", L));
            up(U.dw_w, tt.f32(p + "1.dwconv.conv.weight", (int64_t)L * 7));
            up(U.dw_b, tt.f32(p + "1.dwconv.conv.bias", L));
            up(U.ln_w, tt.f32(p + "1.norm.weight", L));
            up(U.ln_b, tt.f32(p + "1.norm.bias", L));
            load_linear(U.pw1, tt, p + "1.pwconv1", 4 * L, L, true);
            load_linear(U.pw2, tt, p + "1.pwconv2", L, 4 * L, true);
            up(U.gamma, tt.f32(p + "1.gamma", L));
            U.st.H = 6; U.st.C = L;
        }
        const int dd = c.decoder_dim, nb = c.num_upsample_rates;
        load_conv(dec0, tt, "decoder.0.conv", dd, 7, L);
        st_dec0.H = 6; st_dec0.C = L;
        blocks.resize(nb);
        for (int b = 0; b < nb; ++b) {
            const std::string p = "decoder." + std::to_string(1 + b) + ".block.";
            DecBlock& Bk = blocks[b];
            Bk.rate = c.upsample_rates[b];
            B2A_CHECK(Bk.rate >= 1 && Bk.rate <= 16, B2A_ERR_INVALID_INPUT, "speech tokenizer: bad upsample rate");
            total_up *= Bk.rate;
            Bk.cin = dd >> b; Bk.cout = dd >> (b + 1);
            B2A_CHECK(Bk.cin % 8 == 0 && Bk.cout % 8 == 0, B2A_ERR_INVALID_INPUT, "speech tokenizer: decoder channels must be multiples of 8");
            load_snake(Bk.sn, tt, p + "0", Bk.cin);
            Bk.ct.build(convt_w(conv_w(tt, p + "1.conv.weight", Bk.cout, 2 * Bk.rate, Bk.cin), Bk.cout, 2 * Bk.rate, Bk.cin, Bk.rate), Bk.rate * Bk.cout, 2, Bk.cin, use_f16);
            Bk.ct.set_bias(tt.f32(p + "1.conv.bias", Bk.cout));
            Bk.st.H = 1; Bk.st.C = Bk.cin;
            const int dil[3] = {1, 3, 9};
            for (int j = 0; j < 3; ++j) {
                const std::string q = p + std::to_string(2 + j) + ".";
                ResUnit& R = Bk.ru[j];
                R.dil = dil[j];
                load_snake(R.a1, tt, q + "act1", Bk.cout);
                load_snake(R.a2, tt, q + "act2", Bk.cout);
                load_conv(R.c1, tt, q + "conv1.conv", Bk.cout, 7, Bk.cout);
                load_conv(R.c2, tt, q + "conv2.conv", Bk.cout, 1, Bk.cout);
                R.st.H = 6 * R.dil; R.st.C = Bk.cout;
            }
        }
        const int Cf = dd >> nb;
        load_snake(out_snake, tt, "decoder." + std::to_string(nb + 1), Cf);
        {
            const b2a_tensor& t = tt.get("decoder." + std::to_string(nb + 2) + ".conv.weight");
            B2A_CHECK(t.ndim == 3 && t.shape[0] == 1 && t.shape[2] == Cf && t.shape[1] >= 1 && t.shape[1] <= FC_MAXK, B2A_ERR_MODEL_NOT_INITIALIZED,
                      "bad shape for tensor: decoder output conv weight");
            out_k = (int)t.shape[1];
            up(out_w, tt.f32("decoder." + std::to_string(nb + 2) + ".conv.weight", (int64_t)out_k * Cf));
            out_b = tt.f32("decoder." + std::to_string(nb + 2) + ".conv.bias", 1)[0];
            st_out.H = out_k - 1; st_out.C = Cf;
        }
        B2A_CUDA(cudaDeviceSynchronize());
        alloc_state();
        reset();
    }

    // ------------------------------------------------------------------------------------------- state
    void alloc_state() {
        const int B = cfg.max_batch;
        auto ps = [&](PlaneState& s) { for (int i = 0; i < 2; ++i) s.s[i].alloc((size_t)2 * B * std::max(s.H, 1) * s.C); };
        auto fs = [&](F32State& s) { for (int i = 0; i < 2; ++i) s.s[i].alloc((size_t)B * std::max(s.H, 1) * s.C); };
        ps(st_pre); ps(st_dec0); fs(st_out);
        for (auto& U : ups) fs(U.st);
        for (auto& Bk : blocks) { ps(Bk.st); for (auto& R : Bk.ru) ps(R.st); }
        const size_t kv = (size_t)B * cfg.num_key_value_heads * cfg.max_cache_frames * cfg.head_dim;
        for (auto& T : layers) { T.K.alloc(kv); T.V.alloc(kv); }
    }

    void reset() {                // resetStreamingState (:949-970): zero history == the reference's causal zero padding
        B2A_CUDA(cudaSetDevice(device));
        auto zp = [&](PlaneState& s) { for (int i = 0; i < 2; ++i) B2A_CUDA(cudaMemsetAsync(s.s[i].p, 0, s.s[i].n * sizeof(bf16), stream)); };
        auto zf = [&](F32State& s) { for (int i = 0; i < 2; ++i) B2A_CUDA(cudaMemsetAsync(s.s[i].p, 0, s.s[i].n * sizeof(float), stream)); };
        zp(st_pre); zp(st_dec0); zf(st_out);
        for (auto& U : ups) zf(U.st);
        for (auto& Bk : blocks) { zp(Bk.st); for (auto& R : Bk.ru) zp(R.st); }
        B2A_CUDA(cudaStreamSynchronize(stream));     // a following step may run on a caller's stream
        parity = 0; chunk_idx = 0; cache_len = 0; stream_B = 0;
    }

    // ------------------------------------------------------------------------------------------- launches
    bf16* planes(DBuf<bf16>& buf, int B, long long frames, int C) {
        const size_t need = (size_t)2 * B * frames * C;
        B2A_CHECK(need <= buf.n, B2A_ERR_GENERATION_FAILED, "speech tokenizer: internal workspace too small");
        return buf.p;
    }
    void conv(const IW& W, const bf16* in, long long in_frames, ic::Args a, cudaStream_t s) {
        a.M = W.M; a.m_tiles = cdiv(W.M, tc::BM);
        a.taps = W.taps; a.cblocks = W.cblocks;
        if (a.dil == 0) a.dil = 1;
        if (a.up == 0) a.up = 1;
        a.Cout = W.M / a.up;
        a.t_tiles = cdiv(a.T, ic::HALF);
        a.bias = W.has_bias ? W.bias.p : nullptr;
        a.f16 = use_f16;
        a.seg_kb = seg_kb_default();
        a.wscale = use_f16 ? W.rscale.p : nullptr;
        const CUtensorMap tb = make_tmap_planes(in, W.Cin, in_frames, a.B, use_f16);
        const long long tiles = (long long)a.B * a.t_tiles * a.m_tiles;
        launch_pdl(ic::implicit_conv_kernel, dim3((unsigned)std::min<long long>(num_sms, tiles)), dim3(ic::IC_THREADS), ic::SMEM_BYTES, s,
                   W.th, W.tl, tb, a);
    }
    void carry(bf16* X, PlaneState& st, int B, long long T, cudaStream_t s) {
        if (st.H == 0) return;
        const long long n = (long long)2 * B * st.H * (st.C / 8);
        carry_planes_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(X, st.s[parity].p, st.s[parity ^ 1].p, B, (int)T, st.H, st.C / 8);
        count_launch();
    }
    void update_f32(const float* x, F32State& st, int B, long long T, cudaStream_t s) {
        if (st.H == 0) return;
        const long long n = (long long)B * st.H * st.C;
        state_update_f32_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(x, st.s[parity].p, st.s[parity ^ 1].p, B, (int)T, st.H, st.C);
        count_launch();
    }
    void attention(TLayer& L, int B, int T, bf16* out, cudaStream_t s) {
        const dim3 grid(cdiv(T, AT_WARPS), cfg.num_attention_heads, B), block(AT_WARPS * 32);
        const float scale = 1.0f / sqrtf((float)cfg.head_dim);
        const int nh = cfg.num_attention_heads, nkv = cfg.num_key_value_heads, cap = cfg.max_cache_frames;
        if (cfg.head_dim == 32) attn_kernel<1><<<grid, block, 0, s>>>(Q.p, L.K.p, L.V.p, out, B, T, cache_len, nh, nkv, cap, scale, use_f16);
        else if (cfg.head_dim == 64) attn_kernel<2><<<grid, block, 0, s>>>(Q.p, L.K.p, L.V.p, out, B, T, cache_len, nh, nkv, cap, scale, use_f16);
        else attn_kernel<4><<<grid, block, 0, s>>>(Q.p, L.K.p, L.V.p, out, B, T, cache_len, nh, nkv, cap, scale, use_f16);
        count_launch();
    }

    long long out_len(int T) const { return (long long)T * total_up; }

    // streamingStep (:973-1008): d_codes [B, nq, T] int32 (device) -> d_wave [B, T * total_up]
    void step_dev(const int* dcodes, int B, int nq, int T, float* d_wave, cudaStream_t s) {
        const auto& c = cfg;
        B2A_CHECK(B >= 1 && B <= c.max_batch, B2A_ERR_INVALID_INPUT, "speech tokenizer: batch must be in [1, max_batch]");
        B2A_CHECK(T >= 1 && nq >= 1 && nq <= c.num_quantizers, B2A_ERR_INVALID_INPUT, "speech tokenizer: need >= 1 frame and 1..num_quantizers code groups");
        B2A_CHECK(chunk_idx == 0 || B == stream_B, B2A_ERR_INVALID_INPUT, "speech tokenizer: batch size changed inside a stream (reset first)");
        B2A_CHECK(cache_len + T <= c.max_cache_frames, B2A_ERR_INVALID_INPUT, "speech tokenizer: stream longer than max_cache_frames");
        B2A_CHECK((long long)T * total_up + 64 < (1ll << 31) && (long long)B * T * total_up / 64 < (1ll << 31), B2A_ERR_INVALID_INPUT,
                  "speech tokenizer: chunk too large");      // per-row frame indices and grid sizes are 32-bit
        B2A_CUDA(cudaSetDevice(device));
        stream_B = B;
        const int cbd = c.codebook_dim, L = c.latent_dim, Hd = c.hidden_size, I = c.intermediate_size, nh = c.num_attention_heads, hd = c.head_dim;
        const long long N = (long long)B * T;
        // workspace: the largest planar activation and fp32 tensors of the chunk
        {
            size_t pmax = 0, xc = 0;
            auto pl = [&](long long frames, int C) { pmax = std::max(pmax, (size_t)(2ll * B * frames * C)); };
            pl(T + 2, cbd); pl(T, L); pl(T, std::max(std::max(Hd, I), nh * hd));
            long long Tc = T;
            for (auto& U : ups) { Tc *= U.factor; pl(Tc + 6, L); pl(Tc, 4 * L); xc = std::max(xc, (size_t)((long long)B * Tc * L)); }
            pl(Tc + 1, c.decoder_dim);
            for (auto& Bk : blocks) { Tc *= Bk.rate; pl(Tc + 54, Bk.cout); xc = std::max(xc, (size_t)((long long)B * Tc * Bk.cout)); }
            P0.alloc(pmax); P1.alloc(pmax);
            Xh.alloc((size_t)N * Hd); Xc.alloc(xc);
            Q.alloc((size_t)N * std::max(Mqkv, 2 * I));
        }
        // 1. codebook gathers + both output projections                                    (:112-119)
        rvq_gather_kernel<<<(unsigned)N, 128, 0, s>>>(dcodes, emb.p, planes(P0, B, T, 2 * D2), B, T, nq, c.num_quantizers, c.num_semantic_quantizers,
                                                     c.codebook_size, D2, use_f16);
        count_launch();
        { ic::Args a{}; a.B = B; a.T = T; a.hl = planes(P1, B, T + 2, cbd); a.Hout = 2; conv(rvq_proj, P0.p, T, a, s); }
        // 2. pre_conv (k3 causal) -> input_proj                                             (:929-931, :454)
        carry(P1.p, st_pre, B, T, s);
        { ic::Args a{}; a.B = B; a.T = T; a.hl = planes(P0, B, T, L); conv(pre_conv, P1.p, T + 2, a, s); }
        { ic::Args a{}; a.B = B; a.T = T; a.xo = Xh.p; conv(in_proj, P0.p, T, a, s); }
        // 3. transformer layers over the KV cache                                           (:419-428, :473-476)
        for (auto& Ly : layers) {
            rmsnorm_planes_kernel<<<(unsigned)N, RN_THREADS, 0, s>>>(Xh.p, Ly.ln1.p, planes(P0, B, T, Hd), N, Hd, c.rms_norm_eps, use_f16);
            count_launch();
            { ic::Args a{}; a.B = B; a.T = T; a.xo = Q.p; conv(Ly.qkv, P0.p, T, a, s); }
            rope_cache_kernel<<<(unsigned)N, 256, 0, s>>>(Q.p, Ly.K.p, Ly.V.p, inv_freq.p, T, cache_len, nh, c.num_key_value_heads, hd, c.max_cache_frames);
            count_launch();
            attention(Ly, B, T, planes(P1, B, T, nh * hd), s);
            { ic::Args a{}; a.B = B; a.T = T; a.xo = Xh.p; a.add = 1; a.gamma = Ly.sc_attn.p; conv(Ly.o, P1.p, T, a, s); }
            rmsnorm_planes_kernel<<<(unsigned)N, RN_THREADS, 0, s>>>(Xh.p, Ly.ln2.p, planes(P0, B, T, Hd), N, Hd, c.rms_norm_eps, use_f16);
            count_launch();
            { ic::Args a{}; a.B = B; a.T = T; a.xo = Q.p; conv(Ly.gu, P0.p, T, a, s); }
            swiglu_planes_kernel<<<(unsigned)cdiv(N * I, 256), 256, 0, s>>>(Q.p, planes(P1, B, T, I), N, I, use_f16);
            count_launch();
            { ic::Args a{}; a.B = B; a.T = T; a.xo = Xh.p; a.add = 1; a.gamma = Ly.sc_mlp.p; conv(Ly.down, P1.p, T, a, s); }
        }
        dbg_tap(0, Xh.p, N * Hd, s);
        rmsnorm_planes_kernel<<<(unsigned)N, RN_THREADS, 0, s>>>(Xh.p, final_norm.p, planes(P0, B, T, Hd), N, Hd, c.rms_norm_eps, use_f16);
        count_launch();
        { ic::Args a{}; a.B = B; a.T = T; a.hl = planes(P1, B, T, L); conv(out_proj, P0.p, T, a, s); }
        // 4. upsample layers: transposed conv (k = stride) + ConvNeXt                       (:758-775)
        bf16* cur = P1.p;          // planes [2][B][Hcur + Tc][L]
        bf16* other = P0.p;
        long long Tc = T;
        int Hcur = 0;
        for (size_t i = 0; i < ups.size(); ++i) {
            UpLayer& U = ups[i];
            { ic::Args a{}; a.B = B; a.T = (int)Tc; a.up = U.factor; a.xo = Xc.p; conv(U.ct, cur, Hcur + Tc, a, s); }
            Tc *= U.factor;
            dw_ln_kernel<<<(unsigned)(B * Tc), DL_THREADS, 0, s>>>(Xc.p, U.st.s[parity].p, U.dw_w.p, U.dw_b.p, U.ln_w.p, U.ln_b.p, cur, B, (int)Tc, L, 7, 1e-6f, use_f16);
            count_launch();
            update_f32(Xc.p, U.st, B, Tc, s);
            { ic::Args a{}; a.B = B; a.T = (int)Tc; a.gelu = 1; a.hl = other; conv(U.pw1, cur, Tc, a, s); }
            Hcur = i + 1 < ups.size() ? 0 : st_dec0.H;
            { ic::Args a{}; a.B = B; a.T = (int)Tc; a.xo = Xc.p; a.add = 1; a.gamma = U.gamma.p; a.hl = cur; a.Hout = Hcur; conv(U.pw2, other, Tc, a, s); }
            dbg_tap(1 + (int)i, Xc.p, (long long)B * Tc * L, s);
        }
        // 5. decoder.0 (k7 causal) with block 0's SnakeBeta fused                            (:641-656)
        carry(cur, st_dec0, B, Tc, s);
        { ic::Args a{}; a.B = B; a.T = (int)Tc; a.hl = other; a.Hout = 1; a.sa = blocks[0].sn.a.p; a.sb = blocks[0].sn.ib.p; conv(dec0, cur, Hcur + Tc, a, s); }
        std::swap(cur, other);     // cur = planes [2][B][1 + Tc][decoder_dim], already activated
        // 6. decoder blocks                                                                   (:584-601)
        for (size_t b = 0; b < blocks.size(); ++b) {
            DecBlock& Bk = blocks[b];
            carry(cur, Bk.st, B, Tc, s);
            {
                ic::Args a{}; a.B = B; a.T = (int)Tc; a.up = Bk.rate; a.xo = Xc.p; a.hl = other; a.Hout = Bk.ru[0].st.H;
                a.sa = Bk.ru[0].a1.a.p; a.sb = Bk.ru[0].a1.ib.p; a.bias_twice_t0 = chunk_idx > 0 ? 1 : 0;
                conv(Bk.ct, cur, 1 + Tc, a, s);
            }
            Tc *= Bk.rate;
            dbg_tap(10 + 4 * (int)b, Xc.p, (long long)B * Tc * Bk.cout, s);
            std::swap(cur, other);
            for (int j = 0; j < 3; ++j) {
                ResUnit& R = Bk.ru[j];
                carry(cur, R.st, B, Tc, s);
                { ic::Args a{}; a.B = B; a.T = (int)Tc; a.dil = R.dil; a.hl = other; a.sa = R.a2.a.p; a.sb = R.a2.ib.p; conv(R.c1, cur, R.st.H + Tc, a, s); }
                ic::Args a{}; a.B = B; a.T = (int)Tc; a.xo = Xc.p; a.add = 1;
                if (j < 2) { a.hl = cur; a.Hout = Bk.ru[j + 1].st.H; a.sa = Bk.ru[j + 1].a1.a.p; a.sb = Bk.ru[j + 1].a1.ib.p; }
                else if (b + 1 < blocks.size()) { a.hl = cur; a.Hout = 1; a.sa = blocks[b + 1].sn.a.p; a.sb = blocks[b + 1].sn.ib.p; }
                conv(R.c2, other, Tc, a, s);
                dbg_tap(11 + 4 * (int)b + j, Xc.p, (long long)B * Tc * Bk.cout, s);
            }
        }
        // 7. output SnakeBeta + k7 conv to one channel + clip                                 (:693-731, :946)
        {
            const int Cf = st_out.C;
            const size_t smem = ((size_t)(FC_TILE + out_k - 1) * (Cf + 1) + (size_t)out_k * Cf + FC_THREADS) * sizeof(float);
            final_conv_kernel<<<dim3(cdiv(Tc, FC_TILE), B), FC_THREADS, smem, s>>>(Xc.p, st_out.s[parity].p, out_snake.a.p, out_snake.ib.p, out_w.p, out_b,
                                                                                  d_wave, (int)Tc, Cf, out_k);
            count_launch();
            update_f32(Xc.p, st_out, B, Tc, s);
        }
        B2A_CUDA(cudaGetLastError());
        parity ^= 1;
        chunk_idx += 1;
        cache_len += T;
    }

    // host codes [B, nq, T] -> host wave [B, T * total_up]
    void step_host(const int32_t* codes, int B, int nq, int T, float* out, long long out_stride, long long drop) {
        cudaStream_t s = stream;
        const size_t nin = (size_t)B * nq * T;
        const long long nout = out_len(T);
        d_codes.alloc(nin); wave.alloc((size_t)B * nout);
        B2A_CUDA(cudaMemcpyAsync(d_codes.p, codes, nin * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        step_dev(d_codes.p, B, nq, T, wave.p, s);
        B2A_CUDA(cudaMemcpy2DAsync(out, (size_t)out_stride * sizeof(float), wave.p + drop, (size_t)nout * sizeof(float), (size_t)(nout - drop) * sizeof(float),
                                   (size_t)B, cudaMemcpyDeviceToHost, s));
        B2A_CUDA(cudaStreamSynchronize(s));
    }
};

extern "C" {

int32_t b2a_speech_tokenizer_create(int32_t device, const b2a_speech_tokenizer_config* cfg, const b2a_tensor* tensors, int32_t n,
                                    b2a_speech_tokenizer** out) {
    return guarded([&] {
        B2A_CHECK(out, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_create: null out");
        *out = nullptr;
        B2A_CHECK(cfg && tensors && n > 0, B2A_ERR_MODEL_NOT_INITIALIZED, "b2a_speech_tokenizer_create: missing config or weights");
        TensorTable tt(tensors, n);
        *out = new b2a_speech_tokenizer(device, *cfg, tt);
    });
}

int32_t b2a_speech_tokenizer_total_upsample(const b2a_speech_tokenizer* h) { return h ? h->total_up : 0; }
void* b2a_speech_tokenizer_stream(b2a_speech_tokenizer* h) { return h ? (void*)h->stream : nullptr; }

int32_t b2a_speech_tokenizer_reset(b2a_speech_tokenizer* h) {
    return guarded([&] {
        B2A_CHECK(h, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_reset: null handle");
        h->reset();
    });
}

int32_t b2a_speech_tokenizer_streaming_step_dev(b2a_speech_tokenizer* h, const int32_t* d_codes, int32_t B, int32_t nq, int32_t T, float* d_wave,
                                                void* stream) {
    return guarded([&] {
        B2A_CHECK(h && d_codes && d_wave, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_streaming_step_dev: null argument");
        h->step_dev(d_codes, B, nq, T, d_wave, stream ? (cudaStream_t)stream : h->stream);
    });
}

int32_t b2a_speech_tokenizer_streaming_step(b2a_speech_tokenizer* h, const int32_t* codes, int32_t B, int32_t nq, int32_t T, float* wave) {
    return guarded([&] {
        B2A_CHECK(h && codes && wave, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_streaming_step: null argument");
        B2A_CHECK(B >= 1 && T >= 1 && nq >= 1, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_streaming_step: empty input");
        h->step_host(codes, B, nq, T, wave, h->out_len(T), 0);
    });
}

// streamingDecode (:1070-1092): reset, step over chunk_tokens-sized pieces, reset.  codes [B, nq, T] -> wave [B, T * total_up]
int32_t b2a_speech_tokenizer_streaming_decode(b2a_speech_tokenizer* h, const int32_t* codes, int32_t B, int32_t nq, int32_t T, int32_t chunk_tokens,
                                              float* wave) {
    return guarded([&] {
        B2A_CHECK(h && codes && wave, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_streaming_decode: null argument");
        B2A_CHECK(B >= 1 && T >= 1 && nq >= 1 && chunk_tokens >= 1, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_streaming_decode: empty input");
        h->reset();
        std::vector<int32_t> piece;
        for (int start = 0; start < T; start += chunk_tokens) {
            const int n = std::min(chunk_tokens, T - start);
            piece.resize((size_t)B * nq * n);
            for (int r = 0; r < B * nq; ++r) memcpy(&piece[(size_t)r * n], &codes[(size_t)r * T + start], (size_t)n * sizeof(int32_t));
            h->step_host(piece.data(), B, nq, n, wave + h->out_len(start), h->out_len(T), 0);
        }
        h->reset();
    });
}

// chunkedDecode (:1010-1024): every chunk is decoded from a clean state with up to left_context frames of context, whose
// audio is dropped.  codes [B, nq, T] -> wave [B, T * total_up]
int32_t b2a_speech_tokenizer_chunked_decode(b2a_speech_tokenizer* h, const int32_t* codes, int32_t B, int32_t nq, int32_t T, int32_t chunk_size,
                                            int32_t left_context, float* wave) {
    return guarded([&] {
        B2A_CHECK(h && codes && wave, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_chunked_decode: null argument");
        B2A_CHECK(B >= 1 && T >= 1 && nq >= 1 && chunk_size >= 1 && left_context >= 0, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_chunked_decode: empty input");
        std::vector<int32_t> piece;
        for (int start = 0; start < T; start += chunk_size) {
            const int end = std::min(start + chunk_size, T);
            const int ctx = start - left_context > 0 ? left_context : start;
            const int n = end - (start - ctx);
            piece.resize((size_t)B * nq * n);
            for (int r = 0; r < B * nq; ++r) memcpy(&piece[(size_t)r * n], &codes[(size_t)r * T + start - ctx], (size_t)n * sizeof(int32_t));
            h->reset();
            h->step_host(piece.data(), B, nq, n, wave + h->out_len(start), h->out_len(T), h->out_len(ctx));
        }
        h->reset();
    });
}

void b2a_speech_tokenizer_destroy(b2a_speech_tokenizer* h) { delete h; }

// Diagnostics: stage >= 0 makes later decodes keep a copy of the fp32 activation tensor after that stage (0 = transformer output
// before the final norm, 1 + i = upsample layer i, 10 + 4 b = decoder block b after its transposed conv, 11 + 4 b + j = after residual
// unit j); out != null copies the last kept tensor to the host (capacity in floats) and returns its length in *n.
int32_t b2a_speech_tokenizer_debug_stage(b2a_speech_tokenizer* h, int32_t stage, float* out, int64_t capacity, int64_t* n) {
    return guarded([&] {
        B2A_CHECK(h, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_debug_stage: null handle");
        B2A_CUDA(cudaSetDevice(h->device));
        if (out) {
            B2A_CHECK(n && h->dbg_n <= capacity, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_debug_stage: output buffer too small");
            B2A_CUDA(cudaStreamSynchronize(h->stream));
            B2A_CUDA(cudaMemcpy(out, h->dbg.p, (size_t)h->dbg_n * sizeof(float), cudaMemcpyDeviceToHost));
            *n = h->dbg_n;
        }
        h->dbg_stage = stage;
    });
}

// Host-only: the GEMM weight matrix the implicit convolution reads, from an MLX-layout [out, k, in] weight.
// stride == 0: plain causal conv (rows = out, taps = k); stride > 0: transposed conv with k = n * stride (rows = stride * out
// phase-major, taps = n).  layout_out receives [rows][taps][ceil(in / 64) * 64] fp32 (capacity in floats).
int32_t b2a_speech_tokenizer_debug_layout(const float* w, int32_t out, int32_t k, int32_t in, int32_t stride, float* layout_out, int64_t capacity,
                                          int32_t* rows, int32_t* taps, int32_t* kpad) {
    return guarded([&] {
        B2A_CHECK(w && layout_out && rows && taps && kpad && out >= 1 && k >= 1 && in >= 1 && stride >= 0, B2A_ERR_INVALID_INPUT,
                  "b2a_speech_tokenizer_debug_layout: bad argument");
        B2A_CHECK(stride == 0 || k % stride == 0, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_debug_layout: kernel must be a multiple of the stride");
        std::vector<float> src(w, w + (size_t)out * k * in);
        const int M = stride ? stride * out : out, T = stride ? k / stride : k;
        const std::vector<float> g = IW::pad_k(stride ? b2a_speech_tokenizer::convt_w(src, out, k, in, stride) : src, M, T, in);
        B2A_CHECK((int64_t)g.size() <= capacity, B2A_ERR_INVALID_INPUT, "b2a_speech_tokenizer_debug_layout: output buffer too small");
        memcpy(layout_out, g.data(), g.size() * sizeof(float));
        *rows = M; *taps = T; *kpad = cdiv(in, tc::BK) * tc::BK;
    });
}

// Standalone entry for tests/test_gpu_implicit_conv.py: one launch of ic::implicit_conv_kernel on host data.
//   w [M][taps][Cin], x [B][Ttot][Cin] (split into hi/lo planes here), bias / gamma / sa / sb [M / up] or null,
//   xo [B][T*up][M/up] in/out or null, hl_out [B][Hout + T*up][M/up] (hi + lo recombined; frames below Hout come back 0) or null.
int32_t b2a_implicit_conv_test(const float* w, int32_t M, int32_t taps, int32_t Cin, const float* x, int32_t B, int32_t Ttot, int32_t T,
                               int32_t dil, int32_t shift0, int32_t up, const float* bias, const float* gamma, int32_t gelu, int32_t add,
                               int32_t bias_twice_t0, const float* sa, const float* sb, int32_t Hout, int32_t fp16, float* xo, float* hl_out) {
    return guarded([&] {
        B2A_CHECK(w && x && M >= 1 && taps >= 1 && Cin >= 8 && Cin % 8 == 0 && B >= 1 && Ttot >= 1 && T >= 1 && dil >= 1 && shift0 >= 0 && up >= 1 &&
                      M % up == 0 && (M / up) % 8 == 0 && Hout >= 0 && (xo || hl_out) && (!add || xo) && ((sa == nullptr) == (sb == nullptr)),
                  B2A_ERR_INVALID_INPUT, "b2a_implicit_conv_test: bad argument");
        require_device(0);
        B2A_CUDA(cudaSetDevice(0));
        B2A_CUDA(cudaFuncSetAttribute(ic::implicit_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ic::SMEM_BYTES));
        int num_sms = 148;
        B2A_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, 0));
        const int Cout = M / up;
        const long long To = (long long)T * up;
        IW W;
        W.build(std::vector<float>(w, w + (size_t)M * taps * Cin), M, taps, Cin, fp16);
        const size_t nx = (size_t)B * Ttot * Cin, no = (size_t)B * To * Cout, nh = (size_t)B * (Hout + To) * Cout;
        std::vector<uint16_t> xp(2 * nx);
        for (size_t i = 0; i < nx; ++i) split16(x[i], fp16, xp[i], xp[nx + i]);
        DBuf<bf16> dx, dh;
        DBuf<float> dxo, dbias, dgamma, dsa, dsb;
        dx.upload(reinterpret_cast<const bf16*>(xp.data()), xp.size());
        ic::Args a{};
        a.M = M; a.m_tiles = cdiv(M, tc::BM); a.taps = taps; a.cblocks = W.cblocks; a.dil = dil; a.shift0 = shift0;
        a.B = B; a.T = T; a.t_tiles = cdiv(T, ic::HALF); a.Cout = Cout; a.up = up; a.gelu = gelu; a.add = add; a.bias_twice_t0 = bias_twice_t0; a.Hout = Hout; a.f16 = fp16;
        a.seg_kb = seg_kb_default();
        a.wscale = fp16 ? W.rscale.p : nullptr;
        if (bias) { dbias.upload(bias, Cout); a.bias = dbias.p; }
        if (gamma) { dgamma.upload(gamma, Cout); a.gamma = dgamma.p; }
        if (sa) { dsa.upload(sa, Cout); dsb.upload(sb, Cout); a.sa = dsa.p; a.sb = dsb.p; }
        if (xo) { dxo.upload(xo, no); a.xo = dxo.p; }
        if (hl_out) { dh.alloc(2 * nh); B2A_CUDA(cudaMemset(dh.p, 0, 2 * nh * sizeof(bf16))); a.hl = dh.p; }
        B2A_CUDA(cudaDeviceSynchronize());
        const CUtensorMap tb = make_tmap_planes(dx.p, Cin, Ttot, B, fp16);
        const long long tiles = (long long)B * a.t_tiles * a.m_tiles;
        launch_pdl(ic::implicit_conv_kernel, dim3((unsigned)std::min<long long>(num_sms, tiles)), dim3(ic::IC_THREADS), ic::SMEM_BYTES, (cudaStream_t)0,
                   W.th, W.tl, tb, a);
        B2A_CUDA(cudaGetLastError());
        B2A_CUDA(cudaDeviceSynchronize());
        if (xo) B2A_CUDA(cudaMemcpy(xo, dxo.p, no * sizeof(float), cudaMemcpyDeviceToHost));
        if (hl_out) {
            std::vector<uint16_t> hp(2 * nh);
            B2A_CUDA(cudaMemcpy(hp.data(), dh.p, 2 * nh * sizeof(uint16_t), cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < nh; ++i) hl_out[i] = join16(hp[i], hp[nh + i], fp16);
        }
    });
}

}  // extern "C"
