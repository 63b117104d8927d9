// SNAC codec decode + RVQ code search for sm_100a.  Replaces (reference paths):
//   Sources/MLXAudioCodecs/SNAC/VQ.swift:14-20,47-120,150-191    (RVQ lookup / code search)
//   Sources/MLXAudioCodecs/SNAC/Layers.swift:44-50,54-183,202-232,263-315,364-421 (decoder)
//   Sources/MLXAudioCodecs/SNAC/SNACDecoder.swift:127-131          (SNAC.decode)
//
// HBM layout: activations float32 [B, C, T] (time contiguous -> every load/store is coalesced
// along T); two ping-pong buffers sized for the widest stage.  Weight-norm (g*v/||v||) is folded
// once at load time instead of on every call (Layers.swift:102-103,166).  1x1 convolutions and the
// transposed convolutions are GEMMs C[M,N] = A[M,K] * B[K,N] with N = time:
//   1x1 conv : A = W[co,ci]                         B[k,n] = x[ci=k, t=n]
//   convT    : A[(co*s+r),(tap*Cin+ci)] = W[ci, r+tap*s, co]   B[k,n] = x[ci, q=n-tap]
//              (kernel 2s, stride s: exactly two taps per output phase r), scattered to
//              t_out = q*s + r - pad.
// Snake activations are fused into the epilogue of the producing kernel (or the prologue of the
// depthwise conv), the residual add / noise injection into the GEMM epilogue.
#include "common.cuh"
#include "conv_gemm.cuh"
#include "snac_fused.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace b2a {

__device__ __forceinline__ float snake_fi(float x, float alpha, float inv) {      // 1 / (alpha + 1e-9) hoisted by the caller
    const float ax = alpha * x;
    const float k = rintf(ax * 0.15915494309189535f);
    float r = fmaf(k, -6.28318548202514648f, ax);
    r = fmaf(k, 1.7484555e-7f, r);
    const float s = __sinf(r);
    return fmaf(inv * s, s, x);
}
__device__ __forceinline__ float snake_f(float x, float alpha) {
    // Layers.swift:44-50: x + 1/(alpha + 1e-9) * sin(alpha*x)^2
    // explicit 2*pi range reduction + MUFU.SIN (|error| < 5e-7): libdevice sinf is ~40 dependent instructions per call
    const float ax = alpha * x;
    const float k = rintf(ax * 0.15915494309189535f);
    float r = fmaf(k, -6.28318548202514648f, ax);
    r = fmaf(k, 1.7484555e-7f, r);
    const float s = __sinf(r);
    return x + (1.0f / (alpha + 1e-9f)) * s * s;
}

// ------------------------------------------------------------------------------------------------
// RVQ lookup: z[b,c,t] (+)= bias_i[c] + sum_d Wout_i[c,d] * codebook_i[codes_i[b, t/s_i], d]
// (VectorQuantize.decodeCode + outProj + repeat-interleave, VQ.swift:88-94,165-191)
// ------------------------------------------------------------------------------------------------
struct RvqLevel {
    const int* codes;      // [B, T/stride]
    const float* codebook; // [N, D]
    const float* wout;     // [C, D]
    const float* bias;     // [C]
    int stride;
};
struct RvqArgs { RvqLevel lv[4]; int n_levels; int D; int C; int T; int codebook_size; };

// sign: +1 accumulate into out (beta = 1) or write (beta = 0); used with sign=-1 for the residual
__global__ void rvq_lookup_kernel(RvqArgs a, float* __restrict__ out, float beta, float sign) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = blockIdx.y, b = blockIdx.z;
    if (t >= a.T) return;
    float acc = 0.f;
    for (int i = 0; i < a.n_levels; ++i) {
        const RvqLevel& L = a.lv[i];
        const int Ti = a.T / L.stride;
        int code = L.codes[(long long)b * Ti + t / L.stride];
        code = min(max(code, 0), a.codebook_size - 1);
        const float* e = L.codebook + (long long)code * a.D;
        const float* w = L.wout + (long long)c * a.D;
        float v = L.bias[c];
        for (int d = 0; d < a.D; ++d) v = fmaf(w[d], e[d], v);
        acc += v;
    }
    float* o = out + ((long long)b * a.C + c) * a.T + t;
    *o = (beta != 0.f ? beta * (*o) : 0.f) + sign * acc;
}

// ------------------------------------------------------------------------------------------------
// Depthwise conv k7 (dilation d, "same" padding) with optional Snake before and after.
// One CTA = one (b, c) row segment; the activated input tile (+halo) is staged in shared memory
// so sin() is evaluated once per element.
// ------------------------------------------------------------------------------------------------
constexpr int DW_TT = 1024, DW_THREADS = 256, DW_MAXHALO = 27;

__global__ void __launch_bounds__(DW_THREADS)
dwconv7_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w /*[C,7]*/,
               const float* __restrict__ bias, const float* __restrict__ alpha_in,
               const float* __restrict__ alpha_out, int C, int T, int dil) {
    __shared__ float s[DW_TT + 2 * DW_MAXHALO];
    const int c = blockIdx.y, b = blockIdx.z;
    const int t0 = blockIdx.x * DW_TT;
    const int halo = 3 * dil;
    const float* x = in + ((long long)b * C + c) * T;
    const float ai = alpha_in ? alpha_in[c] : 0.f;
    for (int i = threadIdx.x; i < DW_TT + 2 * halo; i += DW_THREADS) {
        const int t = t0 + i - halo;
        float v = 0.f;
        if (t >= 0 && t < T) { v = x[t]; if (alpha_in) v = snake_f(v, ai); }
        s[i] = v;
    }
    __syncthreads();
    float wk[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) wk[k] = w[c * 7 + k];
    const float bv = bias ? bias[c] : 0.f;
    const float ao = alpha_out ? alpha_out[c] : 0.f;
    float* y = out + ((long long)b * C + c) * T;
    for (int i = threadIdx.x; i < DW_TT; i += DW_THREADS) {
        const int t = t0 + i;
        if (t >= T) break;
        float acc = bv;
#pragma unroll
        for (int k = 0; k < 7; ++k) acc = fmaf(wk[k], s[i + k * dil], acc);
        y[t] = alpha_out ? snake_f(acc, ao) : acc;
    }
}

// ------------------------------------------------------------------------------------------------
// fp32 GEMM core: C[M,N] = A[M,K] * B[K,N], 64x128 CTA tile, 8x8 per thread, BK = 16.
// ------------------------------------------------------------------------------------------------
constexpr int GM = 64, GN = 128, GK = 16, G_THREADS = 128;

enum : int { EPI_PLAIN = 0, EPI_RESIDUAL = 1, EPI_NOISE = 2, EPI_CONVT = 3 };

struct GemmArgs {
    const float* A;      // [M, K] row-major (folded weights)
    const float* X;      // input activations [B, Cin, Tin]
    float* Y;            // output [B, Cout, Tout]
    const float* bias;   // [Cout] or null
    const float* res;    // EPI_RESIDUAL / EPI_NOISE: [B, Cout, Tout] added to the result
    const float* noise;  // EPI_NOISE: [B, Tout] (null => generated from seed)
    const float* alpha_out;  // optional Snake on the result, [Cout]
    int M, N, K;
    int Cin, Tin, Cout, Tout;
    int stride, pad;     // EPI_CONVT
    unsigned long long seed;
    int noise_layer;
};

__device__ __forceinline__ float gauss_from_counter(unsigned long long seed, unsigned long long idx) {
    // counter-based N(0,1): splitmix64 -> two uniforms -> Box-Muller (MLXRandom.normal stand-in)
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u1 = ((unsigned)(z >> 40) + 1.0f) * (1.0f / 16777217.0f);
    const float u2 = (unsigned)((z >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

template <int EPI>
__global__ void __launch_bounds__(G_THREADS)
gemm_f32_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[GK][GM];
    __shared__ __align__(16) float Bs[GK][GN];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN, b = blockIdx.z;
    const int ty = tid / 16, tx = tid % 16;
    const float* Xb = g.X + (long long)b * g.Cin * g.Tin;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < g.K; k0 += GK) {
        // A tile: 64 rows x 16 k ; thread loads 8 consecutive k of one row
        {
            const int r = tid >> 1, kk = (tid & 1) * 8;
            const int m = m0 + r;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + kk + j;
                As[kk + j][r] = (m < g.M && k < g.K) ? g.A[(long long)m * g.K + k] : 0.f;
            }
        }
        // B tile: 16 k x 128 n ; thread loads column n = tid of every row (coalesced per warp)
        {
            const int n = n0 + tid;
#pragma unroll
            for (int kk = 0; kk < GK; ++kk) {
                const int k = k0 + kk;
                float v = 0.f;
                if (k < g.K && n < g.N) {
                    if (EPI == EPI_CONVT) {
                        const int tap = k / g.Cin, ci = k - tap * g.Cin, q = n - tap;
                        if (q >= 0 && q < g.Tin) v = Xb[(long long)ci * g.Tin + q];
                    } else {
                        v = Xb[(long long)k * g.Tin + n];
                    }
                }
                Bs[kk][tid] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][32 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 32 + ty * 4 + (i - 4));
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (EPI == EPI_CONVT) {
                const int co = m / g.stride, r = m - co * g.stride;
                const int t = n * g.stride + r - g.pad;
                if (t < 0 || t >= g.Tout) continue;
                if (g.bias) v += g.bias[co];
                if (g.alpha_out) v = snake_f(v, g.alpha_out[co]);
                g.Y[((long long)b * g.Cout + co) * g.Tout + t] = v;
            } else {
                const long long o = ((long long)b * g.Cout + m) * g.Tout + n;
                if (g.bias) v += g.bias[m];
                if (EPI == EPI_RESIDUAL) v += g.res[o];
                if (EPI == EPI_NOISE) {
                    // NoiseBlock (Layers.swift:271-278): x + noise[b,0,t] * (W x)
                    const float nz = g.noise ? g.noise[(long long)b * g.Tout + n]
                                             : gauss_from_counter(g.seed + 0x1000193ull * (g.noise_layer + 1),
                                                                  (unsigned long long)b * g.Tout + n);
                    v = g.res[o] + nz * v;
                }
                if (g.alpha_out) v = snake_f(v, g.alpha_out[m]);
                g.Y[o] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Final conv k7 (C -> 1) + tanh on an already Snake-activated input (Layers.swift:411-415)
// ------------------------------------------------------------------------------------------------
constexpr int FC_TT = 512, FC_THREADS = 128, FC_CH = 16;

__global__ void __launch_bounds__(FC_THREADS)
final_conv7_tanh_kernel(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ w /*[C,7]*/,
                        float bias, int C, int T) {
    __shared__ float s[FC_CH][FC_TT + 8];
    __shared__ float sw[FC_CH * 7];
    const int b = blockIdx.y, t0 = blockIdx.x * FC_TT;
    float acc[4] = {bias, bias, bias, bias};
    for (int c0 = 0; c0 < C; c0 += FC_CH) {
        const int nc = min(FC_CH, C - c0);
        for (int i = threadIdx.x; i < nc * (FC_TT + 6); i += FC_THREADS) {
            const int cc = i / (FC_TT + 6), tt = i - cc * (FC_TT + 6);
            const int t = t0 + tt - 3;
            s[cc][tt] = (t >= 0 && t < T) ? in[((long long)b * C + c0 + cc) * T + t] : 0.f;
        }
        for (int i = threadIdx.x; i < nc * 7; i += FC_THREADS) sw[i] = w[c0 * 7 + i];
        __syncthreads();
        for (int cc = 0; cc < nc; ++cc)
#pragma unroll
            for (int k = 0; k < 7; ++k) {
                const float wv = sw[cc * 7 + k];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = fmaf(wv, s[cc][threadIdx.x + j * FC_THREADS + k], acc[j]);
            }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int t = t0 + threadIdx.x + j * FC_THREADS;
        if (t < T) out[(long long)b * T + t] = tanhf(acc[j]);
    }
}

// ------------------------------------------------------------------------------------------------
// RVQ encode side (VQ.swift:47-120): avg-pool + in_proj, then nearest-code search in explicitly
// ordered float32 (no FMA contraction) so indices are reproducible bit for bit.
// ------------------------------------------------------------------------------------------------
// ze[n = b*Ts + t', d] = bias[d] + sum_c Win[d,c] * mean_j res[b, c, t'*s + j]
__global__ void vq_inproj_kernel(const float* __restrict__ res, const float* __restrict__ win /*[D,C]*/,
                                 const float* __restrict__ bias, float* __restrict__ ze, int C, int T, int stride,
                                 int D) {
    const int Ts = T / stride;
    const int tp = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (tp >= Ts) return;
    float acc[16];
    for (int d = 0; d < D; ++d) acc[d] = bias[d];
    const float inv = 1.0f / (float)stride;
    for (int c = 0; c < C; ++c) {
        const float* x = res + ((long long)b * C + c) * T + (long long)tp * stride;
        float m = 0.f;
        for (int j = 0; j < stride; ++j) m += x[j];
        m *= inv;
        for (int d = 0; d < D; ++d) acc[d] = fmaf(win[d * C + c], m, acc[d]);
    }
    for (int d = 0; d < D; ++d) ze[((long long)b * Ts + tp) * D + d] = acc[d];
}

// rows of x [N, D] -> L2-normalised rows + squared norm of the normalised row (VQ.swift:14-20)
__global__ void l2_normalize_rows_kernel(const float* __restrict__ x, float* __restrict__ xn, float* __restrict__ sq,
                                         int N, int D) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = __fadd_rn(acc, __fmul_rn(x[(long long)n * D + d], x[(long long)n * D + d]));
    const float nrm = fmaxf(__fsqrt_rn(acc), 1e-12f);
    float s2 = 0.f;
    for (int d = 0; d < D; ++d) {
        const float v = __fdiv_rn(x[(long long)n * D + d], nrm);
        xn[(long long)n * D + d] = v;
        s2 = __fadd_rn(s2, __fmul_rn(v, v));
    }
    sq[n] = s2;
}

constexpr int NC_TILE = 1024, NC_THREADS = 128, NC_MAXD = 16;
// idx[n] = first argmin_j ((|e_n|^2 - 2 e_n.c_j) + |c_j|^2)   (== argMax(-dist), VQ.swift:111-115)
__global__ void __launch_bounds__(NC_THREADS)
nearest_code_kernel(const float* __restrict__ en, const float* __restrict__ e2, const float* __restrict__ cn,
                    const float* __restrict__ c2, int* __restrict__ idx, int N, int n_codes, int D) {
    extern __shared__ float sh[];  // [NC_TILE * D] codes + [NC_TILE] norms
    float* sc = sh;
    float* sc2 = sh + NC_TILE * D;
    const int n = blockIdx.x * NC_THREADS + threadIdx.x;
    float e[NC_MAXD];
    float my_e2 = 0.f;
    if (n < N) {
        for (int d = 0; d < D; ++d) e[d] = en[(long long)n * D + d];
        my_e2 = e2[n];
    }
    float best = INFINITY;
    int best_i = 0;
    for (int j0 = 0; j0 < n_codes; j0 += NC_TILE) {
        const int nj = min(NC_TILE, n_codes - j0);
        __syncthreads();
        for (int i = threadIdx.x; i < nj * D; i += NC_THREADS) sc[i] = cn[(long long)j0 * D + i];
        for (int i = threadIdx.x; i < nj; i += NC_THREADS) sc2[i] = c2[j0 + i];
        __syncthreads();
        if (n < N)
            for (int j = 0; j < nj; ++j) {
                float dot = 0.f;
                for (int d = 0; d < D; ++d) dot = __fadd_rn(dot, __fmul_rn(e[d], sc[j * D + d]));
                const float dist = __fadd_rn(__fsub_rn(my_e2, __fmul_rn(2.0f, dot)), sc2[j]);
                if (dist < best) { best = dist; best_i = j0 + j; }
            }
    }
    if (n < N) idx[n] = best_i;
}

// ------------------------------------------------------------------------------------------------
// Channels-last (NLC) tensor-core path: activations fp32 [B*T, C] plus bf16 hi/lo copies in 64-token tiles that
// the conv GEMM (conv_gemm.cuh) reads through TMA.  The kernels below are the non-GEMM pieces.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void put_hilo_nlc(__nv_bfloat16* base, long long ld, long long tok, long long col, float v) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const long long r = (tok / 64) * 128 + (tok % 64);
    base[r * ld + col] = hi;
    base[(r + 64) * ld + col] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// z[b*T + t, c] = sum_i (bias_i[c] + Wout_i[c,:] . codebook_i[codes_i[b, t/s_i], :])      (VQ.swift:165-191)
__global__ void rvq_lookup_nlc_kernel(RvqArgs a, float* __restrict__ out) {
    const long long tok = blockIdx.x;                  // b*T + t
    const int b = (int)(tok / a.T), t = (int)(tok - (long long)b * a.T);
    __shared__ float se[4][16];
    if (threadIdx.x < a.n_levels * a.D) {
        const int i = threadIdx.x / a.D, d = threadIdx.x - i * a.D;
        const RvqLevel& L = a.lv[i];
        int code = L.codes[(long long)b * (a.T / L.stride) + t / L.stride];
        code = min(max(code, 0), a.codebook_size - 1);
        se[i][d] = L.codebook[(long long)code * a.D + d];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
        float acc = 0.f;
        for (int i = 0; i < a.n_levels; ++i) {
            const float* w = a.lv[i].wout + (long long)c * a.D;
            float v = a.lv[i].bias[c];
            for (int d = 0; d < a.D; ++d) v = fmaf(w[d], se[i][d], v);
            acc += v;
        }
        out[tok * a.C + c] = acc;
    }
}

// Depthwise conv k7 (dilation d) along time in NLC, optional Snake before / after, output as bf16 hi/lo tiles.
// CTA = 128 tokens (+ halo) x CT <= 64 channels; the Snake'd input tile lives in shared memory (sin once per element,
// halo overhead 1.42x at dilation 9); float4 global loads, one channel PAIR per thread so hi and lo leave as bf16x2.
// C must be even (every SNAC width is a multiple of 64); C % 4 == 0 takes the vector load path.
constexpr int DWN_TT = 128, DWN_THREADS = 256;
__global__ void __launch_bounds__(DWN_THREADS)
dw7_nlc_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, const float* __restrict__ w /*[C,7]*/,
               const float* __restrict__ bias, const float* __restrict__ alpha_in, const float* __restrict__ alpha_out,
               int T, int C, int CT, int dil) {
    extern __shared__ __align__(16) float dsm[];     // [(128 + 6*dil)][CT]
    const int t0 = blockIdx.x * DWN_TT, c0 = blockIdx.y * CT, b = blockIdx.z;
    const int halo = 3 * dil, rows = DWN_TT + 2 * halo;
    const float* xb = x + (long long)b * T * C;
    if ((C & 3) == 0 && (CT & 3) == 0) {
        const int cq = CT >> 2;
        for (int i = threadIdx.x; i < rows * cq; i += DWN_THREADS) {
            const int r = i / cq, c = (i - r * cq) * 4;
            const int t = t0 + r - halo;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && t < T && c0 + c < C) {
                v = *reinterpret_cast<const float4*>(xb + (long long)t * C + c0 + c);
                if (alpha_in) {
                    const float4 al = *reinterpret_cast<const float4*>(alpha_in + c0 + c);
                    v.x = snake_fi(v.x, al.x, 1.0f / (al.x + 1e-9f)); v.y = snake_fi(v.y, al.y, 1.0f / (al.y + 1e-9f));
                    v.z = snake_fi(v.z, al.z, 1.0f / (al.z + 1e-9f)); v.w = snake_fi(v.w, al.w, 1.0f / (al.w + 1e-9f));
                }
            }
            *reinterpret_cast<float4*>(dsm + (size_t)r * CT + c) = v;
        }
    } else {
        for (int i = threadIdx.x; i < rows * CT; i += DWN_THREADS) {
            const int r = i / CT, c = i - r * CT;
            const int t = t0 + r - halo;
            float v = 0.f;
            if (t >= 0 && t < T && c0 + c < C) {
                v = xb[(long long)t * C + c0 + c];
                if (alpha_in) v = snake_f(v, alpha_in[c0 + c]);
            }
            dsm[i] = v;
        }
    }
    __syncthreads();
    const int hp = CT >> 1;                                    // channel pairs per row
    const int c = (threadIdx.x % hp) * 2, grp = threadIdx.x / hp, ngrp = DWN_THREADS / hp;
    if (c0 + c >= C) return;
    float wa[7], wb[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) { wa[k] = w[(c0 + c) * 7 + k]; wb[k] = w[(c0 + c + 1) * 7 + k]; }
    const float ba = bias ? bias[c0 + c] : 0.f, bb = bias ? bias[c0 + c + 1] : 0.f;
    const float aoa = alpha_out ? alpha_out[c0 + c] : 0.f, aob = alpha_out ? alpha_out[c0 + c + 1] : 0.f;
    const float ioa = 1.0f / (aoa + 1e-9f), iob = 1.0f / (aob + 1e-9f);
    for (int tt = grp; tt < DWN_TT; tt += ngrp) {
        const int t = t0 + tt;
        if (t >= T) break;
        float va = ba, vb = bb;
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const float2 xv = *reinterpret_cast<const float2*>(dsm + (size_t)(tt + k * dil) * CT + c);
            va = fmaf(wa[k], xv.x, va); vb = fmaf(wb[k], xv.y, vb);
        }
        if (alpha_out) { va = snake_fi(va, aoa, ioa); vb = snake_fi(vb, aob, iob); }
        const long long tok = (long long)b * T + t;
        const long long r = (tok / 64) * 128 + (tok % 64);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(va, vb);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(va - __low2float(hi), vb - __high2float(hi));
        *reinterpret_cast<__nv_bfloat162*>(out + r * C + c0 + c) = hi;
        *reinterpret_cast<__nv_bfloat162*>(out + (r + 64) * C + c0 + c) = lo;
    }
}

// zero the two half-rows per utterance of the 2-tap im2col matrix that no producer writes:
// row b*(T+1) cols [C, 2C) (tap 1 of q = 0) and row b*(T+1)+T cols [0, C) (tap 0 of q = T)
__global__ void x2_zero_edges_kernel(__nv_bfloat16* __restrict__ x2, int T, int C) {
    const int b = blockIdx.x;
    const long long r0 = (long long)b * (T + 1), r1 = r0 + T;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const long long h0 = (r0 / 64) * 128 + (r0 % 64), h1 = (r1 / 64) * 128 + (r1 % 64);
        const __nv_bfloat16 z = __float2bfloat16_rn(0.f);
        x2[h0 * 2 * C + C + c] = z; x2[(h0 + 64) * 2 * C + C + c] = z;
        x2[h1 * 2 * C + c] = z; x2[(h1 + 64) * 2 * C + c] = z;
    }
}

// final Snake -> conv k7 (C -> 1) -> tanh in NLC (Layers.swift:411-415), C == 64.
// 256 tokens per CTA; the Snake'd tile (+3 halo rows each side) is staged in shared memory.  Thread (tg, cg) owns 8 consecutive
// tokens x 8 channels (two float4 column groups {4cg..4cg+3} and {32+4cg..}, so a quarter-warp's LDS.128 covers 128 contiguous
// bytes): 14 row reads feed 8 x 7 x 8 FMAs against weights held in registers; the 8 channel groups are then reduced by shuffles.
constexpr int FN_TT = 256, FN_THREADS = 256, FN_MAXC = 64;
__global__ void __launch_bounds__(FN_THREADS)
final_nlc_kernel(const float* __restrict__ x, float* __restrict__ wave, const float* __restrict__ w /*[C,7]*/,
                 const float* __restrict__ alpha, float bias, int T, int C) {
    extern __shared__ __align__(16) float fsm[];     // [(256 + 6)][64]
    const int t0 = blockIdx.x * FN_TT, b = blockIdx.y;
    const float* xb = x + (long long)b * T * FN_MAXC;
    {   // FN_THREADS is a multiple of 16: a thread always stages the same 4 channels
        const int c = (threadIdx.x & 15) * 4;
        const float4 al = *reinterpret_cast<const float4*>(alpha + c);
        const float4 iv = make_float4(1.0f / (al.x + 1e-9f), 1.0f / (al.y + 1e-9f), 1.0f / (al.z + 1e-9f), 1.0f / (al.w + 1e-9f));
        for (int i = threadIdx.x; i < (FN_TT + 6) * 16; i += FN_THREADS) {
            const int r = i >> 4;
            const int t = t0 + r - 3;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t >= 0 && t < T) {
                v = *reinterpret_cast<const float4*>(xb + (long long)t * FN_MAXC + c);
                v.x = snake_fi(v.x, al.x, iv.x); v.y = snake_fi(v.y, al.y, iv.y); v.z = snake_fi(v.z, al.z, iv.z); v.w = snake_fi(v.w, al.w, iv.w);
            }
            *reinterpret_cast<float4*>(fsm + r * FN_MAXC + c) = v;
        }
    }
    const int cgp = threadIdx.x & 7, tg = threadIdx.x >> 3;     // 8 channel groups x 32 token groups
    float wk[8][7];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = (i < 4 ? 0 : 28) + cgp * 4 + i;            // i >= 4 -> 32 + 4*cgp + (i - 4)
#pragma unroll
        for (int k = 0; k < 7; ++k) wk[i][k] = w[c * 7 + k];
    }
    __syncthreads();
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const float* base = fsm + (tg * 8) * FN_MAXC + cgp * 4;
#pragma unroll
    for (int r = 0; r < 14; ++r) {
        const float4 lo = *reinterpret_cast<const float4*>(base + r * FN_MAXC);
        const float4 hi = *reinterpret_cast<const float4*>(base + r * FN_MAXC + 32);
        const float xv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int k = r - j;
            if (k >= 0 && k < 7) {
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[j] = fmaf(wk[i][k], xv[i], acc[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float v = acc[j];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        acc[j] = v;
    }
    const int t = t0 + tg * 8 + cgp;                            // lane cgp of the group writes token cgp
    float out = acc[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) out = cgp == j ? acc[j] : out;
    if (t < T) wave[(long long)b * T + t] = tanhf(out + bias);
}

// ------------------------------------------------------------------------------------------------
// Host-side model
// ------------------------------------------------------------------------------------------------
struct ConvW {       // folded weights on the device
    DBuf<float> w, bias;
    bool has_bias = false;
};

// fp32 weight matrix [M, K] as two bf16 K-major operands (hi + lo) with their TMA maps
struct TcW {
    DBuf<__nv_bfloat16> hi, lo;
    CUtensorMap th{}, tl{};
    int M = 0, K = 0;
    void build(const std::vector<float>& W, int M_, int K_) {
        M = M_; K = K_;
        std::vector<__nv_bfloat16> h((size_t)M * K), l((size_t)M * K);
        for (size_t i = 0; i < h.size(); ++i) {
            h[i] = __float2bfloat16_rn(W[i]);
            l[i] = __float2bfloat16_rn(W[i] - __bfloat162float(h[i]));
        }
        hi.upload(h.data(), h.size());
        lo.upload(l.data(), l.size());
        B2A_CUDA(cudaDeviceSynchronize());
        th = tc::make_tmap_bf16(hi.p, M, K, tc::BM);
        tl = tc::make_tmap_bf16(lo.p, M, K, tc::BM);
    }
};

struct ResUnit { DBuf<float> a0, a2; ConvW dw, pw; TcW pw_tc, pw_bd; int dil; };   // pw_bd: [W 0; 0 W] for the fused C = 64 kernel
struct DecBlock {
    int cin, cout, stride, pad;
    DBuf<float> alpha;   // Snake before the transposed conv
    ConvW ct;            // A[(co*s+r), (tap*Cin+ci)]
    ConvW noise;         // [cout, cout], no bias
    TcW ct_tc, noise_tc, noise_bd; // tensor-core operands: ct_tc rows are m = r*cout + co (phase-major); *_bd block-diagonal (C = 64)
    bool has_noise;
    ResUnit ru[3];
};

static std::vector<float> fold_wn(const TensorTable& tt, const std::string& prefix, int d0, int d1, int d2, bool eps) {
    // weight = g * v / (||v||_{dims 1,2} (+ 1e-12))   Layers.swift:102-103 (eps) / :166 (no eps)
    std::vector<float> v = tt.f32(prefix + ".weight_v", (int64_t)d0 * d1 * d2);
    std::vector<float> g = tt.f32(prefix + ".weight_g", d0);
    for (int i = 0; i < d0; ++i) {
        double s = 0;
        for (int j = 0; j < d1 * d2; ++j) s += (double)v[(size_t)i * d1 * d2 + j] * v[(size_t)i * d1 * d2 + j];
        const double nrm = std::sqrt(s) + (eps ? 1e-12 : 0.0);
        for (int j = 0; j < d1 * d2; ++j)
            v[(size_t)i * d1 * d2 + j] = (float)((double)g[i] * v[(size_t)i * d1 * d2 + j] / nrm);
    }
    return v;
}

static void load_bias(const TensorTable& tt, const std::string& prefix, int n, ConvW& c) {
    if (tt.find(prefix + ".bias")) {
        std::vector<float> b = tt.f32(prefix + ".bias", n);
        c.bias.upload(b.data(), n);
        c.has_bias = true;
    }
}

}  // namespace b2a

using namespace b2a;

struct b2a_snac {
    int device;
    b2a_snac_config cfg;
    int latent, hop;
    cudaStream_t stream = nullptr;
    // quantizer
    struct Level { DBuf<float> codebook, wout, bout, win, bin, cb_n, cb_n2; int stride; };
    std::vector<Level> levels;
    // decoder
    ConvW dw0, pw0, conv0;   // depthwise: dw0 + pw0 ; otherwise conv0 (k7 dense, unsupported on device)
    TcW pw0_tc;
    bool use_tc = true;      // tcgen05 / NLC path (B2A_SNAC=simt selects the fp32 CUDA-core path)
    int num_sms = 148;
    DBuf<float> xs, xs2;                 // NLC fp32 activations of the current stage (ping-pong for the fused units)
    bool use_fused = true;               // fused ResidualUnit / NoiseBlock kernel for C <= 128 (B2A_SNAC_FUSED=0 disables)
    DBuf<__nv_bfloat16> hA, x2;          // hi/lo tiles: GEMM input of the stage / 2-tap im2col of the next transposed conv
    std::vector<DecBlock> blocks;
    DBuf<float> alpha_final;
    ConvW final_conv;
    int final_c = 0;
    float final_bias = 0.f;
    // workspaces
    DBuf<float> bufX, bufY, d_wave, d_noise[8], d_ze, d_en, d_e2, d_zq;
    DBuf<int> d_codes[8], d_idx;

    ~b2a_snac() { if (stream) cudaStreamDestroy(stream); }

    b2a_snac(int dev, const b2a_snac_config& c, const TensorTable& tt) : device(dev), cfg(c) {
        B2A_CHECK(c.attn_window_size == 0, B2A_ERR_INVALID_INPUT,
                  "SNAC: attn_window_size (LocalMHA) is only used by the 32/44 kHz models and is not implemented");
        B2A_CHECK(c.depthwise != 0, B2A_ERR_INVALID_INPUT, "SNAC: only depthwise=true decoders are implemented");
        B2A_CHECK(c.n_vq_strides >= 1 && c.n_vq_strides <= 4 && c.n_decoder_rates >= 1 && c.n_decoder_rates <= 8,
                  B2A_ERR_INVALID_INPUT, "SNAC: unsupported number of codebooks / decoder stages");
        B2A_CHECK(c.codebook_dim >= 1 && c.codebook_dim <= NC_MAXD, B2A_ERR_INVALID_INPUT, "SNAC: codebook_dim > 16");
        require_device(dev);
        B2A_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        latent = c.latent_dim > 0 ? c.latent_dim : c.encoder_dim << c.n_encoder_rates;
        hop = 1;
        for (int i = 0; i < c.n_decoder_rates; ++i) hop *= c.decoder_rates[i];
        const int D = c.codebook_dim, N = c.codebook_size;
        levels.resize(c.n_vq_strides);
        for (int i = 0; i < c.n_vq_strides; ++i) {
            const std::string q = "quantizer.quantizers." + std::to_string(i);
            Level& L = levels[i];
            L.stride = c.vq_strides[i];
            std::vector<float> cb = tt.f32(q + ".codebook.weight", (int64_t)N * D);
            L.codebook.upload(cb.data(), cb.size());
            std::vector<float> wo = fold_wn(tt, q + ".out_proj", latent, 1, D, true);
            L.wout.upload(wo.data(), wo.size());
            std::vector<float> bo = tt.f32(q + ".out_proj.bias", latent);
            L.bout.upload(bo.data(), bo.size());
            if (tt.find(q + ".in_proj.weight_v")) {
                std::vector<float> wi = fold_wn(tt, q + ".in_proj", D, 1, latent, true);
                L.win.upload(wi.data(), wi.size());
                std::vector<float> bi = tt.f32(q + ".in_proj.bias", D);
                L.bin.upload(bi.data(), bi.size());
            }
            L.cb_n.alloc((size_t)N * D);
            L.cb_n2.alloc(N);
            l2_normalize_rows_kernel<<<cdiv(N, 128), 128, 0, stream>>>(L.codebook.p, L.cb_n.p, L.cb_n2.p, N, D);
            count_launch();
        }
        const std::string p = "decoder.model.layers.";
        const int C = c.decoder_dim;
        {
            std::vector<float> w = fold_wn(tt, p + "0", latent, 7, 1, true);
            dw0.w.upload(w.data(), w.size());
            load_bias(tt, p + "0", latent, dw0);
            std::vector<float> w1 = fold_wn(tt, p + "1", C, 1, latent, true);
            pw0.w.upload(w1.data(), w1.size());
            host_pw0 = w1;
            load_bias(tt, p + "1", C, pw0);
        }
        int li = 2;
        blocks.resize(c.n_decoder_rates);
        for (int i = 0; i < c.n_decoder_rates; ++i, ++li) {
            DecBlock& B = blocks[i];
            B.cin = C >> i; B.cout = C >> (i + 1); B.stride = c.decoder_rates[i];
            B.pad = (B.stride + 1) / 2;  // Int(ceil(stride/2)), Layers.swift:295
            const std::string b = p + std::to_string(li) + ".block.layers.";
            std::vector<float> a = tt.f32(b + "0.alpha", B.cin);
            B.alpha.upload(a.data(), a.size());
            const int s = B.stride, k = 2 * s;
            std::vector<float> wt = fold_wn(tt, b + "1", B.cin, k, B.cout, false);  // [ci, k, co]
            std::vector<float> A((size_t)B.cout * s * 2 * B.cin);
            for (int co = 0; co < B.cout; ++co)
                for (int r = 0; r < s; ++r)
                    for (int tap = 0; tap < 2; ++tap)
                        for (int ci = 0; ci < B.cin; ++ci)
                            A[((size_t)(co * s + r)) * (2 * B.cin) + tap * B.cin + ci] =
                                wt[((size_t)ci * k + (r + tap * s)) * B.cout + co];
            B.ct.w.upload(A.data(), A.size());
            {   // tensor-core operand: rows phase-major (m = r*cout + co) so a warp's 32 lanes write 32 consecutive channels
                std::vector<float> At((size_t)B.cout * s * 2 * B.cin);
                for (int co = 0; co < B.cout; ++co)
                    for (int r = 0; r < s; ++r)
                        memcpy(&At[((size_t)r * B.cout + co) * 2 * B.cin], &A[((size_t)co * s + r) * 2 * B.cin], (size_t)2 * B.cin * sizeof(float));
                host_ct.push_back(At);
            }
            load_bias(tt, b + "1", B.cout, B.ct);
            int j = 2;
            B.has_noise = c.noise != 0;
            if (B.has_noise) {
                std::vector<float> wn_ = fold_wn(tt, b + "2.linear", B.cout, 1, B.cout, true);
                B.noise.w.upload(wn_.data(), wn_.size());
                host_noise.push_back(wn_);
                j = 3;
            }
            const int dils[3] = {1, 3, 9};
            for (int u = 0; u < 3; ++u, ++j) {
                ResUnit& R = B.ru[u];
                R.dil = dils[u];
                const std::string r = b + std::to_string(j) + ".block.layers.";
                std::vector<float> a0 = tt.f32(r + "0.alpha", B.cout), a2 = tt.f32(r + "2.alpha", B.cout);
                R.a0.upload(a0.data(), a0.size());
                R.a2.upload(a2.data(), a2.size());
                std::vector<float> wd = fold_wn(tt, r + "1", B.cout, 7, 1, true);
                R.dw.w.upload(wd.data(), wd.size());
                load_bias(tt, r + "1", B.cout, R.dw);
                std::vector<float> wp = fold_wn(tt, r + "3", B.cout, 1, B.cout, true);
                R.pw.w.upload(wp.data(), wp.size());
                host_pw.push_back(wp);
                load_bias(tt, r + "3", B.cout, R.pw);
            }
        }
        final_c = C >> c.n_decoder_rates;
        {
            std::vector<float> a = tt.f32(p + std::to_string(li) + ".alpha", final_c);
            alpha_final.upload(a.data(), a.size());
            std::vector<float> w = fold_wn(tt, p + std::to_string(li + 1), 1, 7, final_c, true);  // [1,7,C]
            std::vector<float> wt((size_t)final_c * 7);
            for (int k = 0; k < 7; ++k)
                for (int ci = 0; ci < final_c; ++ci) wt[(size_t)ci * 7 + k] = w[(size_t)k * final_c + ci];
            final_conv.w.upload(wt.data(), wt.size());
            if (tt.find(p + std::to_string(li + 1) + ".bias")) final_bias = tt.f32(p + std::to_string(li + 1) + ".bias", 1)[0];
        }
        B2A_CUDA(cudaStreamSynchronize(stream));
        B2A_CUDA(cudaGetLastError());
        B2A_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
        const char* env = getenv("B2A_SNAC");
        bool shapes_ok = latent % 64 == 0 && C % 64 == 0 && final_c == FN_MAXC && c.noise != 0;
        for (auto& B : blocks) shapes_ok = shapes_ok && B.cin % 64 == 0 && B.cout % 64 == 0;
        use_tc = shapes_ok && !(env && std::string(env) == "simt");
        if (use_tc) {
            B2A_CUDA(cudaFuncSetAttribute(cg::conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cg::SMEM_BYTES));
            B2A_CUDA(cudaFuncSetAttribute(dw7_nlc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            fused_attrs<64>(); fused_attrs<128>();
            B2A_CUDA(cudaFuncSetAttribute(rf::convt_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rf::convt_smem_bytes()));
            B2A_CUDA(cudaFuncSetAttribute(final_nlc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (FN_TT + 6) * FN_MAXC * (int)sizeof(float)));
            const char* ef = getenv("B2A_SNAC_FUSED");
            use_fused = !(ef && std::string(ef) == "0");
            pw0_tc.build(host_pw0, C, latent);
            size_t ip = 0;
            for (size_t i = 0; i < blocks.size(); ++i) {
                DecBlock& B = blocks[i];
                B.ct_tc.build(host_ct[i], B.cout * B.stride, 2 * B.cin);
                B.noise_tc.build(host_noise[i], B.cout, B.cout);
                auto blockdiag = [&](const std::vector<float>& w) {       // [64, 64] -> [128, 128] = [W 0; 0 W]
                    std::vector<float> d((size_t)128 * 128, 0.f);
                    for (int r = 0; r < 64; ++r)
                        for (int c2 = 0; c2 < 64; ++c2) { d[(size_t)r * 128 + c2] = w[(size_t)r * 64 + c2]; d[(size_t)(r + 64) * 128 + 64 + c2] = w[(size_t)r * 64 + c2]; }
                    return d;
                };
                if (B.cout == 64) B.noise_bd.build(blockdiag(host_noise[i]), 128, 128);
                for (int u = 0; u < 3; ++u) {
                    if (B.cout == 64) B.ru[u].pw_bd.build(blockdiag(host_pw[ip]), 128, 128);
                    B.ru[u].pw_tc.build(host_pw[ip++], B.cout, B.cout);
                }
            }
        }
        host_pw0.clear(); host_ct.clear(); host_noise.clear(); host_pw.clear();
    }
    std::vector<float> host_pw0;
    std::vector<std::vector<float>> host_ct, host_noise, host_pw;

    // ---- tensor-core / NLC decode --------------------------------------------------------------------------
    void cgemm(const TcW& W, const __nv_bfloat16* X, long long x_rows, cg::Args a, cudaStream_t s) {
        a.M = W.M; a.K = W.K;
        a.m_tiles = cdiv(W.M, tc::BM); a.k_blocks = W.K / tc::BK; a.n_tiles = cdiv(a.N, cg::HALF);
        const CUtensorMap tb = tc::make_tmap_bf16(X, x_rows, W.K, 128);
        const long long tiles = (long long)a.n_tiles * a.m_tiles;
        launch_pdl(cg::conv_gemm_kernel, dim3((unsigned)std::min<long long>(num_sms, tiles)), dim3(cg::CG_THREADS), cg::SMEM_BYTES, s,
                   W.th, W.tl, tb, a);
    }
    void dw_nlc(const ConvW& W, const float* xin, __nv_bfloat16* out, const float* a_in, const float* a_out, int batch, int T, int C,
                int dil, cudaStream_t s) {
        const int CT = std::min(C, 64);
        B2A_CHECK(C % 2 == 0 && (C <= 64 || C % 64 == 0), B2A_ERR_INVALID_INPUT, "snac: channel widths must be even (multiples of 64 above 64)");
        const size_t sm = (size_t)(DWN_TT + 6 * dil) * CT * sizeof(float);
        dw7_nlc_kernel<<<dim3(cdiv(T, DWN_TT), cdiv(C, CT), batch), DWN_THREADS, sm, s>>>(xin, out, W.w.p, W.has_bias ? W.bias.p : nullptr,
                                                                                       a_in, a_out, T, C, CT, dil);
        count_launch();
    }
    static long long pad64(long long n) { return (n + 63) / 64 * 64; }
    template <int CC>
    static void fused_attrs() {
        B2A_CUDA(cudaFuncSetAttribute(rf::ru_fused_kernel<rf::MODE_RU, 1, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B2A_CUDA(cudaFuncSetAttribute(rf::ru_fused_kernel<rf::MODE_RU, 3, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B2A_CUDA(cudaFuncSetAttribute(rf::ru_fused_kernel<rf::MODE_RU, 9, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        B2A_CUDA(cudaFuncSetAttribute(rf::ru_fused_kernel<rf::MODE_NOISE, 0, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    }
    bool block_fused(const DecBlock& B) const {
        bool ok = use_fused && (B.cout == 64 || B.cout == 128);
        for (int u = 0; u < 3; ++u) ok = ok && (B.ru[u].dil == 1 || B.ru[u].dil == 3 || B.ru[u].dil == 9);
        return ok;
    }
    // Snake + transposed conv fused (snac_fused.cuh convt_fused_kernel): 128 input channels -> stride * C_out = 128 phase rows,
    // fed by a fused block (its fp32 output is the only copy)
    bool convt_fused_ok(size_t i) const {
        if (!use_fused || i == 0 || i >= blocks.size()) return false;
        const DecBlock& B = blocks[i];
        return B.cin == 128 && B.stride * B.cout == 128 && block_fused(blocks[i - 1]) && block_fused(B);
    }
    template <int CC>
    void fused_c(const TcW& W, const rf::Args& a, dim3 g, size_t sm, cudaStream_t s) {
        const dim3 bl(rf::THREADS);
        if (a.mode == rf::MODE_NOISE) launch_pdl(rf::ru_fused_kernel<rf::MODE_NOISE, 0, CC>, g, bl, sm, s, W.th, W.tl, a);
        else if (a.dil == 1) launch_pdl(rf::ru_fused_kernel<rf::MODE_RU, 1, CC>, g, bl, sm, s, W.th, W.tl, a);
        else if (a.dil == 3) launch_pdl(rf::ru_fused_kernel<rf::MODE_RU, 3, CC>, g, bl, sm, s, W.th, W.tl, a);
        else launch_pdl(rf::ru_fused_kernel<rf::MODE_RU, 9, CC>, g, bl, sm, s, W.th, W.tl, a);
    }
    // W: the [128, 128] operand (the layer's own weights for C = 128, the block-diagonal copy for C = 64)
    void fused(const TcW& W, rf::Args a, int batch, long long T, cudaStream_t s) {
        a.B = batch; a.T = (int)T;
        const int tile_tokens = a.C == 64 ? 2 * rf::TOK : rf::TOK;
        a.tiles_per_utt = cdiv(T, tile_tokens); a.n_tiles = (long long)batch * a.tiles_per_utt;
        const long long ctas = std::min<long long>(num_sms, (a.n_tiles + rf::TEAMS - 1) / rf::TEAMS);
        const size_t sm = rf::smem_bytes(a.dil, a.mode);
        if (a.C == 64) fused_c<64>(W, a, dim3((unsigned)ctas), sm, s);
        else fused_c<128>(W, a, dim3((unsigned)ctas), sm, s);
    }

    void decode_dev_tc(const int* const* d_codes_in, int batch, long long T, const float* const* d_noise_in, int noise_mode,
                       unsigned long long seed, float* d_wave_out, cudaStream_t s) {
        const int C = cfg.decoder_dim;
        // buffer sizes over all stages
        size_t max_x = (size_t)batch * T * std::max(latent, C), max_h = (size_t)(2 * pad64((long long)batch * T)) * (size_t)std::max(latent, C), max_x2 = 0;
        {
            long long t = T;
            for (auto& B : blocks) {
                max_x2 = std::max<size_t>(max_x2, (size_t)(2 * pad64((long long)batch * (t + 1))) * 2 * B.cin);
                t *= B.stride;
                max_x = std::max<size_t>(max_x, (size_t)batch * t * B.cout);
                max_h = std::max<size_t>(max_h, (size_t)(2 * pad64((long long)batch * t)) * B.cout);
            }
        }
        xs.alloc(max_x); xs2.alloc(max_x); hA.alloc(max_h); x2.alloc(max_x2);
        // RVQ lookup -> z (NLC) ; depthwise k7 -> hi/lo ; 1x1 (768 -> 1024) + Snake(block 0) -> 2-tap im2col of block 0
        RvqArgs ra{};
        ra.n_levels = (int)levels.size(); ra.D = cfg.codebook_dim; ra.C = latent; ra.T = (int)T; ra.codebook_size = cfg.codebook_size;
        for (size_t i = 0; i < levels.size(); ++i)
            ra.lv[i] = RvqLevel{d_codes_in[i], levels[i].codebook.p, levels[i].wout.p, levels[i].bout.p, levels[i].stride};
        rvq_lookup_nlc_kernel<<<(unsigned)(batch * T), 256, 0, s>>>(ra, xs.p);
        count_launch();
        dw_nlc(dw0, xs.p, hA.p, nullptr, nullptr, batch, (int)T, latent, 1, s);
        {
            x2_zero_edges_kernel<<<batch, 256, 0, s>>>(x2.p, (int)T, C);
            count_launch();
            cg::Args a{};
            a.N = (int)(batch * T); a.epi = cg::E_STORE_HILO; a.bias = pw0.has_bias ? pw0.bias.p : nullptr; a.alpha = blocks[0].alpha.p;
            a.hl = x2.p; a.ldh = 2 * C; a.dual = 1; a.T = (int)T;
            cgemm(pw0_tc, hA.p, 2 * pad64((long long)batch * T), a, s);
        }
        long long t = T;
        for (size_t i = 0; i < blocks.size(); ++i) {
            DecBlock& B = blocks[i];
            const long long tout = t * B.stride, ntok = (long long)batch * tout;
            if (convt_fused_ok(i)) {
                // last block: Snake + transposed conv straight from the previous block's fp32 output (no 2-tap im2col round trip)
                rf::ConvtArgs a{};
                a.x = xs.p; a.y = xs2.p; a.alpha = B.alpha.p; a.bias = B.ct.has_bias ? B.ct.bias.p : nullptr;
                a.Tin = (int)t; a.T = (int)tout; a.B = batch; a.stride = B.stride; a.cout = B.cout; a.pad = B.pad;
                a.tiles_per_utt = cdiv(t + 1, rf::TOK); a.n_tiles = (long long)batch * a.tiles_per_utt;
                const long long ctas = std::min<long long>(num_sms, (a.n_tiles + rf::TEAMS - 1) / rf::TEAMS);
                launch_pdl(rf::convt_fused_kernel, dim3((unsigned)ctas), dim3(rf::THREADS), rf::convt_smem_bytes(), s, B.ct_tc.th, B.ct_tc.tl, a);
                std::swap(xs.p, xs2.p); std::swap(xs.n, xs2.n);
            } else {   // transposed conv: tokens (b, q), q = 0..t ; rows m = r*cout + co ; scatter to t_out = q*s + r - pad
                cg::Args a{};
                a.N = (int)(batch * (t + 1)); a.epi = cg::E_CONVT; a.bias = B.ct.has_bias ? B.ct.bias.p : nullptr;
                a.x = xs.p; a.ldx = B.cout; a.hl = block_fused(B) ? nullptr : hA.p;     // the fused units read fp32 only
                a.ldh = B.cout; a.T = (int)tout; a.Cout = B.cout; a.stride = B.stride;
                a.pad = B.pad; a.Tin = (int)t;
                cgemm(B.ct_tc, x2.p, 2 * pad64((long long)batch * (t + 1)), a, s);
            }
            const float* nz = d_noise_in ? d_noise_in[i] : nullptr;
            const bool fz = block_fused(B);
            if (fz) {
                // narrow stages: one fused kernel per NoiseBlock / ResidualUnit, fp32 in -> fp32 out (ping-pong xs <-> xs2)
                float* cur = xs.p; float* oth = xs2.p;
                if (B.has_noise && (nz || noise_mode == 0)) {
                    rf::Args a{};
                    a.x = cur; a.y = oth; a.C = B.cout; a.mode = rf::MODE_NOISE; a.dil = 0; a.noise = nz;
                    a.seed = seed + 0x1000193ull * (i + 1);
                    fused(B.cout == 64 ? B.noise_bd : B.noise_tc, a, batch, tout, s);
                    std::swap(cur, oth);
                }
                for (int u = 0; u < 3; ++u) {
                    ResUnit& R = B.ru[u];
                    rf::Args a{};
                    a.x = cur; a.y = oth; a.C = B.cout; a.mode = rf::MODE_RU; a.dil = R.dil;
                    a.dw_w = R.dw.w.p; a.dw_b = R.dw.has_bias ? R.dw.bias.p : nullptr; a.a_in = R.a0.p; a.a_mid = R.a2.p;
                    a.pw_bias = R.pw.has_bias ? R.pw.bias.p : nullptr;
                    if (u == 2 && i + 1 < blocks.size() && !convt_fused_ok(i + 1)) {
                        x2_zero_edges_kernel<<<batch, 256, 0, s>>>(x2.p, (int)tout, B.cout);
                        count_launch();
                        a.hl = x2.p; a.a_next = blocks[i + 1].alpha.p;
                    }
                    fused(B.cout == 64 ? R.pw_bd : R.pw_tc, a, batch, tout, s);
                    std::swap(cur, oth);
                }
                if (cur != xs.p) { std::swap(xs.p, xs2.p); std::swap(xs.n, xs2.n); }       // the live activation is always xs
                t = tout;
                continue;
            }
            if (B.has_noise && (nz || noise_mode == 0)) {
                cg::Args a{};
                a.N = (int)ntok; a.epi = cg::E_NOISE; a.x = xs.p; a.ldx = B.cout; a.noise = nz;
                a.seed = seed + 0x1000193ull * (i + 1); a.T = (int)tout;
                cgemm(B.noise_tc, hA.p, 2 * pad64(ntok), a, s);
            }
            for (int u = 0; u < 3; ++u) {
                ResUnit& R = B.ru[u];
                dw_nlc(R.dw, xs.p, hA.p, R.a0.p, R.a2.p, batch, (int)tout, B.cout, R.dil, s);
                cg::Args a{};
                a.N = (int)ntok; a.bias = R.pw.has_bias ? R.pw.bias.p : nullptr; a.x = xs.p; a.ldx = B.cout; a.T = (int)tout;
                if (u == 2 && i + 1 < blocks.size()) {
                    x2_zero_edges_kernel<<<batch, 256, 0, s>>>(x2.p, (int)tout, B.cout);
                    count_launch();
                    a.epi = cg::E_ADD_HILO; a.alpha = blocks[i + 1].alpha.p; a.hl = x2.p; a.ldh = 2 * B.cout; a.dual = 1;
                } else {
                    a.epi = cg::E_ADD;
                }
                cgemm(R.pw_tc, hA.p, 2 * pad64(ntok), a, s);
            }
            t = tout;
        }
        final_nlc_kernel<<<dim3(cdiv(t, FN_TT), batch), FN_THREADS, (FN_TT + 6) * FN_MAXC * sizeof(float), s>>>(xs.p, d_wave_out, final_conv.w.p, alpha_final.p, final_bias,
                                                                            (int)t, final_c);
        count_launch();
        B2A_CUDA(cudaGetLastError());
    }

    size_t max_act(int batch, long long T) const {
        size_t m = (size_t)batch * std::max(latent, cfg.decoder_dim) * T;
        long long t = T;
        for (auto& B : blocks) { t *= B.stride; m = std::max<size_t>(m, (size_t)batch * B.cout * t); }
        return m;
    }

    void gemm(int epi, const ConvW& W, const float* X, float* Y, const float* res, const float* noise,
              const float* alpha_out, int batch, int M, int N, int K, int Cin, int Tin, int Cout, int Tout, int stride,
              int pad, unsigned long long seed, int layer, cudaStream_t s) {
        GemmArgs g{W.w.p, X, Y, W.has_bias ? W.bias.p : nullptr, res, noise, alpha_out, M, N, K, Cin, Tin, Cout, Tout,
                   stride, pad, seed, layer};
        dim3 grid(cdiv(N, GN), cdiv(M, GM), batch);
        switch (epi) {
            case EPI_PLAIN: gemm_f32_kernel<EPI_PLAIN><<<grid, G_THREADS, 0, s>>>(g); break;
            case EPI_RESIDUAL: gemm_f32_kernel<EPI_RESIDUAL><<<grid, G_THREADS, 0, s>>>(g); break;
            case EPI_NOISE: gemm_f32_kernel<EPI_NOISE><<<grid, G_THREADS, 0, s>>>(g); break;
            default: gemm_f32_kernel<EPI_CONVT><<<grid, G_THREADS, 0, s>>>(g); break;
        }
        count_launch();
    }

    void dwconv(const ConvW& W, const float* in, float* out, const float* a_in, const float* a_out, int batch, int C,
                int T, int dil, cudaStream_t s) {
        dim3 grid(cdiv(T, DW_TT), C, batch);
        dwconv7_kernel<<<grid, DW_THREADS, 0, s>>>(in, out, W.w.p, W.has_bias ? W.bias.p : nullptr, a_in, a_out, C, T,
                                                   dil);
        count_launch();
    }

    // codes/noise/wave are DEVICE pointers
    void decode_dev(const int* const* d_codes_in, int batch, long long T, const float* const* d_noise_in, int noise_mode,
                    unsigned long long seed, float* d_wave_out, cudaStream_t s) {
        B2A_CHECK(batch > 0 && T > 0, B2A_ERR_INVALID_INPUT, "snac decode: empty input");
        for (auto& L : levels)
            B2A_CHECK(T % L.stride == 0, B2A_ERR_INVALID_INPUT, "snac decode: t_latent must be a multiple of every vq stride");
        B2A_CHECK(T * hop < (1ll << 31) / 2, B2A_ERR_INVALID_INPUT, "snac decode: sequence too long");
        B2A_CUDA(cudaSetDevice(device));
        if (use_tc && (long long)batch * T * hop < (1ll << 31) - 64) {
            decode_dev_tc(d_codes_in, batch, T, d_noise_in, noise_mode, seed, d_wave_out, s);
            return;
        }
        const size_t need = max_act(batch, T);
        bufX.alloc(need);
        bufY.alloc(need);
        float *X = bufX.p, *Y = bufY.p;
        // RVQ lookup -> X [B, latent, T]
        RvqArgs ra{};
        ra.n_levels = (int)levels.size(); ra.D = cfg.codebook_dim; ra.C = latent; ra.T = (int)T;
        ra.codebook_size = cfg.codebook_size;
        for (size_t i = 0; i < levels.size(); ++i)
            ra.lv[i] = RvqLevel{d_codes_in[i], levels[i].codebook.p, levels[i].wout.p, levels[i].bout.p, levels[i].stride};
        rvq_lookup_kernel<<<dim3(cdiv(T, 256), latent, batch), 256, 0, s>>>(ra, X, 0.f, 1.f);
        count_launch();
        // depthwise k7 + 1x1 (Layers.swift:378-389); Snake of block 0 fused into the 1x1 epilogue
        dwconv(dw0, X, Y, nullptr, nullptr, batch, latent, (int)T, 1, s);
        const int C = cfg.decoder_dim;
        gemm(EPI_PLAIN, pw0, Y, X, nullptr, nullptr, blocks[0].alpha.p, batch, C, (int)T, latent, latent, (int)T, C, (int)T,
             1, 0, 0, 0, s);
        long long t = T;
        for (size_t i = 0; i < blocks.size(); ++i) {
            DecBlock& B = blocks[i];
            const long long tout = t * B.stride;
            // transposed conv: X [cin, t] (already Snake-activated) -> Y [cout, tout]
            gemm(EPI_CONVT, B.ct, X, Y, nullptr, nullptr, nullptr, batch, B.cout * B.stride, (int)t + 1, 2 * B.cin, B.cin,
                 (int)t, B.cout, (int)tout, B.stride, B.pad, 0, 0, s);
            float* cur = Y;
            float* other = X;
            const float* nz = d_noise_in ? d_noise_in[i] : nullptr;
            if (B.has_noise && (nz || noise_mode == 0)) {
                gemm(EPI_NOISE, B.noise, cur, other, cur, nz, nullptr, batch, B.cout, (int)tout, B.cout, B.cout, (int)tout,
                     B.cout, (int)tout, 1, 0, seed, (int)i, s);
                std::swap(cur, other);
            }
            for (int u = 0; u < 3; ++u) {
                ResUnit& R = B.ru[u];
                dwconv(R.dw, cur, other, R.a0.p, R.a2.p, batch, B.cout, (int)tout, R.dil, s);
                const float* a_next = nullptr;
                if (u == 2) a_next = (i + 1 < blocks.size()) ? blocks[i + 1].alpha.p : alpha_final.p;
                gemm(EPI_RESIDUAL, R.pw, other, cur, cur, nullptr, a_next, batch, B.cout, (int)tout, B.cout, B.cout, (int)tout,
                     B.cout, (int)tout, 1, 0, 0, 0, s);
            }
            if (cur != X) std::swap(X, Y);   // keep "X" = current activations
            t = tout;
        }
        final_conv7_tanh_kernel<<<dim3(cdiv(t, FC_TT), batch), FC_THREADS, 0, s>>>(X, d_wave_out, final_conv.w.p, final_bias,
                                                                                   final_c, (int)t);
        count_launch();
        B2A_CUDA(cudaGetLastError());
    }

    void nearest(const float* d_enc, int N, int* d_out_idx, const Level& L, cudaStream_t s) {
        const int D = cfg.codebook_dim;
        d_en.alloc((size_t)N * D);
        d_e2.alloc(N);
        l2_normalize_rows_kernel<<<cdiv(N, 128), 128, 0, s>>>(d_enc, d_en.p, d_e2.p, N, D);
        const size_t sm = (size_t)NC_TILE * (D + 1) * sizeof(float);
        B2A_CUDA(cudaFuncSetAttribute(nearest_code_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        nearest_code_kernel<<<cdiv(N, NC_THREADS), NC_THREADS, sm, s>>>(d_en.p, d_e2.p, L.cb_n.p, L.cb_n2.p, d_out_idx, N,
                                                                        cfg.codebook_size, D);
        count_launch(2);
    }
};

extern "C" {

int32_t b2a_snac_create(int32_t device, const b2a_snac_config* cfg, const b2a_tensor* tensors, int32_t n,
                        b2a_snac** out) {
    return guarded([&] {
        B2A_CHECK(out, B2A_ERR_INVALID_INPUT, "b2a_snac_create: null out");
        *out = nullptr;
        B2A_CHECK(cfg && tensors && n > 0, B2A_ERR_MODEL_NOT_INITIALIZED, "b2a_snac_create: missing config or weights");
        TensorTable tt(tensors, n);
        *out = new b2a_snac(device, *cfg, tt);
    });
}

int64_t b2a_snac_hop_length(const b2a_snac* h) { return h ? h->hop : 0; }
void* b2a_snac_stream(b2a_snac* h) { return h ? (void*)h->stream : nullptr; }

int32_t b2a_snac_decode_dev(b2a_snac* h, const int32_t* const* d_codes, int32_t batch, int64_t T,
                            const float* const* d_noise, int32_t noise_mode, uint64_t seed, float* d_wave, void* stream) {
    return guarded([&] {
        B2A_CHECK(h && d_codes && d_wave, B2A_ERR_INVALID_INPUT, "b2a_snac_decode_dev: null argument");
        h->decode_dev(d_codes, batch, T, d_noise, noise_mode, seed, d_wave, (cudaStream_t)stream);
    });
}

int32_t b2a_snac_decode(b2a_snac* h, const int32_t* const* codes, int32_t batch, int64_t T, const float* const* noise,
                        int32_t noise_mode, uint64_t seed, float* wave) {
    return guarded([&] {
        B2A_CHECK(h && codes && wave, B2A_ERR_INVALID_INPUT, "b2a_snac_decode: null argument");
        B2A_CHECK(batch > 0 && T > 0, B2A_ERR_INVALID_INPUT, "b2a_snac_decode: empty input");
        B2A_CUDA(cudaSetDevice(h->device));
        cudaStream_t s = h->stream;
        const int nl = (int)h->levels.size();
        const int* dc[8];
        for (int i = 0; i < nl; ++i) {
            B2A_CHECK(codes[i], B2A_ERR_INVALID_INPUT, "b2a_snac_decode: null code layer");
            const size_t n = (size_t)batch * (T / h->levels[i].stride);
            h->d_codes[i].alloc(n);
            B2A_CUDA(cudaMemcpyAsync(h->d_codes[i].p, codes[i], n * sizeof(int), cudaMemcpyHostToDevice, s));
            dc[i] = h->d_codes[i].p;
        }
        const float* dn[8] = {nullptr};
        bool any_noise = false;
        long long t = T;
        for (size_t i = 0; i < h->blocks.size(); ++i) {
            t *= h->blocks[i].stride;
            if (noise && noise[i]) {
                h->d_noise[i].alloc((size_t)batch * t);
                B2A_CUDA(cudaMemcpyAsync(h->d_noise[i].p, noise[i], (size_t)batch * t * sizeof(float), cudaMemcpyHostToDevice, s));
                dn[i] = h->d_noise[i].p;
                any_noise = true;
            }
        }
        h->d_wave.alloc((size_t)batch * t);
        h->decode_dev(dc, batch, T, any_noise ? dn : nullptr, noise_mode, seed, h->d_wave.p, s);
        B2A_CUDA(cudaMemcpyAsync(wave, h->d_wave.p, (size_t)batch * t * sizeof(float), cudaMemcpyDeviceToHost, s));
        B2A_CUDA(cudaStreamSynchronize(s));
    });
}

int32_t b2a_snac_quantize(b2a_snac* h, const float* z, int32_t batch, int64_t T, int32_t* const* codes, float* z_q) {
    return guarded([&] {
        B2A_CHECK(h && z && codes, B2A_ERR_INVALID_INPUT, "b2a_snac_quantize: null argument");
        B2A_CHECK(batch > 0 && T > 0, B2A_ERR_AUDIO_ENCODING_FAILED, "b2a_snac_quantize: empty input");
        for (auto& L : h->levels) {
            B2A_CHECK(T % L.stride == 0, B2A_ERR_AUDIO_ENCODING_FAILED, "b2a_snac_quantize: T must be a multiple of every vq stride");
            B2A_CHECK(L.win.p, B2A_ERR_MODEL_NOT_INITIALIZED, "b2a_snac_quantize: in_proj weights were not provided");
        }
        B2A_CUDA(cudaSetDevice(h->device));
        cudaStream_t s = h->stream;
        const int C = h->latent, D = h->cfg.codebook_dim;
        const size_t n = (size_t)batch * C * T;
        h->bufX.alloc(n);   // residual
        h->d_zq.alloc(n);
        B2A_CUDA(cudaMemcpyAsync(h->bufX.p, z, n * sizeof(float), cudaMemcpyHostToDevice, s));
        B2A_CUDA(cudaMemsetAsync(h->d_zq.p, 0, n * sizeof(float), s));
        for (size_t i = 0; i < h->levels.size(); ++i) {
            auto& L = h->levels[i];
            const int Ts = (int)(T / L.stride), N = batch * Ts;
            h->d_ze.alloc((size_t)N * D);
            h->d_idx.alloc(N);
            vq_inproj_kernel<<<dim3(cdiv(Ts, 64), batch), 64, 0, s>>>(h->bufX.p, L.win.p, L.bin.p, h->d_ze.p, C, (int)T, L.stride, D);
            count_launch();
            h->nearest(h->d_ze.p, N, h->d_idx.p, L, s);
            RvqArgs ra{};
            ra.n_levels = 1; ra.D = D; ra.C = C; ra.T = (int)T; ra.codebook_size = h->cfg.codebook_size;
            ra.lv[0] = RvqLevel{h->d_idx.p, L.codebook.p, L.wout.p, L.bout.p, L.stride};
            rvq_lookup_kernel<<<dim3(cdiv(T, 256), C, batch), 256, 0, s>>>(ra, h->d_zq.p, 1.f, 1.f);    // zQ += zQ_i
            rvq_lookup_kernel<<<dim3(cdiv(T, 256), C, batch), 256, 0, s>>>(ra, h->bufX.p, 1.f, -1.f);   // residual -= zQ_i
            count_launch(2);
            B2A_CHECK(codes[i], B2A_ERR_INVALID_INPUT, "b2a_snac_quantize: null code output");
            B2A_CUDA(cudaMemcpyAsync(codes[i], h->d_idx.p, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, s));
        }
        if (z_q) B2A_CUDA(cudaMemcpyAsync(z_q, h->d_zq.p, n * sizeof(float), cudaMemcpyDeviceToHost, s));
        B2A_CUDA(cudaStreamSynchronize(s));
        B2A_CUDA(cudaGetLastError());
    });
}

void b2a_snac_destroy(b2a_snac* h) { delete h; }

}  // extern "C"
