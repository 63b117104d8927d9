// Encodec decode for sm_100a.  Replaces (reference paths):
//   Sources/MLXAudioCodecs/Encodec/EncodecQuantization.swift:117-133  EncodecResidualVectorQuantizer.decode
//   Sources/MLXAudioCodecs/Encodec/Encodec.swift:94-167               EncodecDecoder
//   Sources/MLXAudioCodecs/Encodec/EncodecLayers.swift:15-88          EncodecLSTM / EncodecLSTMBlock (T sequential tiny matmuls)
//   Sources/MLXAudioCodecs/Encodec/EncodecLayers.swift:92-211         EncodecConv1d (causal / asymmetric, clamped reflect padding)
//   Sources/MLXAudioCodecs/Encodec/EncodecLayers.swift:216-273,371-450 transposed conv (a 5-deep scalar host loop in the reference)
//   Sources/MLXAudioCodecs/Encodec/EncodecLayers.swift:278-337        EncodecResnetBlock
//   Sources/MLXAudioCodecs/Encodec/Encodec.swift:294-402              decodeFrame / linearOverlapAdd / decode
// fp32, channels-last [N, T, C] (the reference's own layout), N = chunks x batch.
//   * every dense conv is ONE kernel, ec_conv_kernel: an implicit-GEMM over (token tile x output tile) whose K axis gathers
//     the taps straight from the activation (no im2col buffer), with the ELU of the *input*, the bias and the residual /
//     shortcut fused.  A transposed conv (k = J*s) is the same kernel: output phases are stacked on the M axis
//     (row = r*C_out + co), taps run backwards in time, and because ((q*s + r)*C_out + co) == q*s*C_out + m the phase
//     scatter is a plain contiguous store.  A resnet block is two launches: k3 conv, then [shortcut | k1 conv] as one GEMM
//     over two sources (x raw, hidden through ELU).
//   * the LSTM stack is ONE persistent cooperative kernel: each CTA owns 4 hidden units of every layer, keeps its slices of
//     Wh (and Wx of the upper layers) in shared memory for the whole sequence, layers run as a wavefront (layer l works on
//     time s - l in step s), so the whole block costs T + L - 1 grid barriers instead of L*T dependent launches.
#include "common.cuh"

#include <algorithm>
#include <cmath>

namespace b2a {
namespace ec {

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }

// ------------------------------------------------------------------ RVQ decode: sum of codebook gathers
// codes [N, n_q, T] int32, books [n_q][size, dim] contiguous -> out [N, T, dim]
__global__ void rvq_sum_kernel(const int* __restrict__ codes, const float* __restrict__ books, float* __restrict__ out,
                               int n_q, int T, int size, int dim) {
    const long long tok = blockIdx.x;
    const int n = (int)(tok / T), t = (int)(tok - (long long)n * T);
    for (int c = threadIdx.x; c < dim; c += blockDim.x) {
        float acc = 0.f;
        for (int q = 0; q < n_q; ++q) {
            int idx = codes[((long long)n * n_q + q) * T + t];
            idx = min(max(idx, 0), size - 1);
            acc += books[((long long)q * size + idx) * dim + c];
        }
        out[tok * dim + c] = acc;
    }
}

__global__ void scale_kernel(float* __restrict__ x, long long n, float f) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] *= f;
}

// ------------------------------------------------------------------ implicit-GEMM conv / transposed conv
struct ConvArgs {
    // source A: taps over xa [N, La, Ca]
    const float* xa; int La, Ca, taps, padL, reflect, elu_a, backward;   // forward: src = q + tap - padL; backward: src = q - tap
    // source B (optional): one tap at src = q over xb [N, Lq, Cb]
    const float* xb; int Cb, elu_b;
    const float* A;          // [M, K] row-major, K = taps*Ca + Cb
    const float* bias;       // [M] or null
    const float* res;        // optional residual, same addressing as out
    float* out;              // [N, Tout, Cout]; element (q, m) lives at q*M + m - shift, valid inside [0, Tout*Cout)
    int M, K, Lq, N;
    long long out_per_n;     // Tout * Cout
    long long shift;         // pl * Cout (left trim of a transposed conv)
};

constexpr int BK = 16;

__device__ __forceinline__ int src_index(int q, int tap, const ConvArgs& a) {
    if (a.backward) {
        const int s = q - tap;
        return (s >= 0 && s < a.La) ? s : -1;
    }
    int s = q + tap - a.padL;
    if (s < 0) return a.reflect ? min(-s, a.La - 1) : -1;
    if (s >= a.La) return a.reflect ? max(a.La - 2 - (s - a.La), 0) : -1;
    return s;
}

// BM outputs x BT tokens per CTA, 256 threads, thread (tx = tid % 16 -> m, ty = tid / 16 -> token).
template <int BM, int BT>
__global__ void __launch_bounds__(256) ec_conv_kernel(ConvArgs a) {
    constexpr int RM = BM / 16, RT = BT / 16;
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Xs[BK][BT + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int n = blockIdx.z;
    const int q0 = blockIdx.x * BT, m0 = blockIdx.y * BM;
    const int Ka = a.taps * a.Ca;
    float acc[RM][RT];
#pragma unroll
    for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RT; ++j) acc[i][j] = 0.f;

    // register double buffering: the global loads of k-tile i+1 are in flight while tile i is multiplied out of shared memory
    constexpr int NA = (BM * 4 + 255) / 256, NX = (BT * 4 + 255) / 256;
    float4 ra[NA], rx[NX];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int e = tid + i * 256;
            const int m = e >> 2, k4 = (e & 3) * 4;
            ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < BM * 4 && m0 + m < a.M && k0 + k4 < a.K) ra[i] = *reinterpret_cast<const float4*>(a.A + (long long)(m0 + m) * a.K + k0 + k4);
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const int e = tid + i * 256;
            const int t = e >> 2, k4 = (e & 3) * 4;
            const int q = q0 + t, kk = k0 + k4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < BT * 4 && q < a.Lq && kk < a.K) {
                if (kk < Ka) {
                    const int tap = kk / a.Ca, ci = kk - tap * a.Ca;
                    const int s = src_index(q, tap, a);
                    if (s >= 0) {
                        v = *reinterpret_cast<const float4*>(a.xa + ((long long)n * a.La + s) * a.Ca + ci);
                        if (a.elu_a) { v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w); }
                    }
                } else {
                    v = *reinterpret_cast<const float4*>(a.xb + ((long long)n * a.Lq + q) * a.Cb + (kk - Ka));
                    if (a.elu_b) { v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w); }
                }
            }
            rx[i] = v;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < NA; ++i) {
            const int e = tid + i * 256;
            if (e < BM * 4) {
                const int m = e >> 2, k4 = (e & 3) * 4;
                As[k4 + 0][m] = ra[i].x; As[k4 + 1][m] = ra[i].y; As[k4 + 2][m] = ra[i].z; As[k4 + 3][m] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < NX; ++i) {
            const int e = tid + i * 256;
            if (e < BT * 4) {
                const int t = e >> 2, k4 = (e & 3) * 4;
                Xs[k4 + 0][t] = rx[i].x; Xs[k4 + 1][t] = rx[i].y; Xs[k4 + 2][t] = rx[i].z; Xs[k4 + 3][t] = rx[i].w;
            }
        }
    };
    load_tiles(0);
    store_tiles();
    __syncthreads();
    for (int k0 = 0; k0 < a.K; k0 += BK) {
        const bool more = k0 + BK < a.K;
        if (more) load_tiles(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float av[RM], xv[RT];
#pragma unroll
            for (int i = 0; i < RM; ++i) av[i] = As[k][tx + 16 * i];
#pragma unroll
            for (int j = 0; j < RT; ++j) xv[j] = Xs[k][ty + 16 * j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RT; ++j) acc[i][j] = fmaf(av[i], xv[j], acc[i][j]);
        }
        __syncthreads();
        if (more) {
            store_tiles();
            __syncthreads();
        }
    }
    float* outn = a.out + (long long)n * a.out_per_n;
    const float* resn = a.res ? a.res + (long long)n * a.out_per_n : nullptr;
#pragma unroll
    for (int j = 0; j < RT; ++j) {
        const int q = q0 + ty + 16 * j;
        if (q >= a.Lq) continue;
#pragma unroll
        for (int i = 0; i < RM; ++i) {
            const int m = m0 + tx + 16 * i;
            if (m >= a.M) continue;
            const long long o = (long long)q * a.M + m - a.shift;
            if (o < 0 || o >= a.out_per_n) continue;
            float v = acc[i][j] + (a.bias ? a.bias[m] : 0.f);
            if (resn) v += resn[o];
            outn[o] = v;
        }
    }
}

// ------------------------------------------------------------------ last conv: ELU -> k-tap conv C -> audio channels (1 or 2)
// x [N, L, C] -> per-chunk wave [N, L, CH] scaled by scale[n] (decodeFrame, Encodec.swift:294-301).  256 samples per CTA; the
// tile (+ left halo) goes through ELU once into shared memory; weights [CH, k, C] in shared memory.
__global__ void __launch_bounds__(256) final_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                         const float* __restrict__ scale, float* __restrict__ out, int L, int C, int k,
                                                         int CH, int padL, int reflect) {
    extern __shared__ float sm[];
    float* xs = sm;                          // [(256 + k - 1)][C + 1]
    float* ws = sm + (size_t)(256 + k - 1) * (C + 1);   // [CH][k][C]
    const int n = blockIdx.y, t0 = blockIdx.x * 256;
    const int rows = 256 + k - 1;
    for (int e = threadIdx.x; e < CH * k * C; e += 256) ws[e] = w[e];
    for (int e = threadIdx.x; e < rows * C; e += 256) {
        const int r = e / C, c = e - r * C;
        int s = t0 + r - padL;
        if (s < 0) s = reflect ? min(-s, L - 1) : -1;
        else if (s >= L) s = reflect ? max(L - 2 - (s - L), 0) : -1;
        xs[r * (C + 1) + c] = s >= 0 ? elu1(x[((long long)n * L + s) * C + c]) : 0.f;
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= L) return;
    const float sc = scale ? scale[n] : 1.f;
    for (int ch = 0; ch < CH; ++ch) {
        float acc = 0.f;
        for (int kk = 0; kk < k; ++kk) {
            const float* xr = xs + (threadIdx.x + kk) * (C + 1);
            const float* wr = ws + (ch * k + kk) * C;
            for (int c = 0; c < C; ++c) acc = fmaf(wr[c], xr[c], acc);
        }
        out[((long long)n * L + t) * CH + ch] = (acc + bias[ch]) * sc;
    }
}

// ------------------------------------------------------------------ linearOverlapAdd (Encodec.swift:304-356) as a gather
// frames [n_chunks, B, L, CH] -> out [B, total, CH]; weight w[t] = 0.5 - |(t+1)/(L+1) - 0.5|, normalised by the weight sum.
__global__ void overlap_add_kernel(const float* __restrict__ frames, float* __restrict__ out, int n_chunks, int B, int L, int CH,
                                   int hop, long long total) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (t >= total) return;
    int c_hi = (int)min((long long)n_chunks - 1, t / hop);
    int c_lo = (int)max(0ll, (t - L + hop) / hop);        // smallest c with c*hop + L > t
    for (int ch = 0; ch < CH; ++ch) {
        float acc = 0.f, sw = 0.f;
        for (int c = c_lo; c <= c_hi; ++c) {
            const int tt = (int)(t - (long long)c * hop);
            if (tt < 0 || tt >= L) continue;
            const float wv = 0.5f - fabsf((float)(tt + 1) / (float)(L + 1) - 0.5f);
            acc += wv * frames[(((long long)c * B + b) * L + tt) * CH + ch];
            sw += wv;
        }
        out[((long long)b * total + t) * CH + ch] = sw != 0.f ? acc / sw : acc;
    }
}

// ------------------------------------------------------------------ LSTM stack, persistent wavefront kernel
constexpr int LSTM_MAX_LAYERS = 4;
constexpr int LSTM_BC = 8;          // batch rows per launch
constexpr int LSTM_UNITS = 4;       // hidden units per CTA -> 16 gate rows per layer

struct LstmArgs {
    const float* xproj;                     // [N, T, 4H]: layer-0 input projection + bias
    const float* Wh[LSTM_MAX_LAYERS];       // [4H, H]
    const float* Wx[LSTM_MAX_LAYERS];       // [4H, H] for l >= 1
    const float* bias[LSTM_MAX_LAYERS];     // [4H] for l >= 1
    float* hseq[LSTM_MAX_LAYERS];           // [N, T, H] hidden sequence of every layer (exchange buffer between CTAs)
    const float* skip;                      // [N, T, H] block input
    float* out;                             // [N, T, H] = h_last + skip
    unsigned* bar;                          // grid barrier counter (zeroed by the host)
    int n0, nb, T, H, NL;                   // this launch handles rows n0 .. n0+nb-1 (nb <= LSTM_BC)
};

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// smem: per layer Wh slice [16][H]; per layer l>=1 Wx slice [16][H]; hs [NL][BC][H]; gates [NL][16][BC]; cst [NL][4][BC]
__global__ void __launch_bounds__(256) lstm_kernel(LstmArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int H = a.H, NL = a.NL, T = a.T;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int u0 = blockIdx.x * LSTM_UNITS;
    float* whs = sm;                                        // [NL][16][H]
    float* wxs = whs + (size_t)NL * 16 * H;                 // [NL-1][16][H]
    float* hs = wxs + (size_t)(NL - 1) * 16 * H;            // [NL][BC][H]
    float* gates = hs + (size_t)NL * LSTM_BC * H;           // [NL][16][BC]
    float* cst = gates + NL * 16 * LSTM_BC;                 // [NL][UNITS][BC]
    // row r of a slice = gate (r / UNITS), unit u0 + r % UNITS  ->  global row gate*H + u0 + r%UNITS
    for (int l = 0; l < NL; ++l)
        for (int e = tid; e < 16 * H; e += 256) {
            const int r = e / H, k = e - r * H;
            const long long grow = (long long)(r / LSTM_UNITS) * H + u0 + (r % LSTM_UNITS);
            whs[((size_t)l * 16 + r) * H + k] = a.Wh[l][grow * H + k];
            if (l >= 1) wxs[((size_t)(l - 1) * 16 + r) * H + k] = a.Wx[l][grow * H + k];
        }
    for (int e = tid; e < NL * LSTM_UNITS * LSTM_BC; e += 256) cst[e] = 0.f;
    __syncthreads();

    const int r0 = warp * 2;            // this warp's two rows of every slice
    for (int s = 0; s < T + NL - 1; ++s) {
        // stage hs[l] = h_l[s - l - 1] (zeros before the sequence starts).  All loads of a thread are issued before the first
        // store: a plain load/store loop with a run-time trip count is not software-pipelined by the compiler and paid one L2
        // round trip per iteration (32 x 0.6 us = the 18 us per step measured by the first version).
        {
            const int hq = H >> 2, per_layer = LSTM_BC * hq;          // float4 items per layer (H % 4 == 0)
            for (int base = 0; base < NL * per_layer; base += 256 * 8) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + u * 256 + tid;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (e < NL * per_layer) {
                        const int l = e / per_layer, r = e - l * per_layer, b = r / hq, k4 = r - b * hq;
                        const int tp = s - l - 1;
                        if (tp >= 0 && tp < T && b < a.nb)
                            v[u] = __ldcg(reinterpret_cast<const float4*>(a.hseq[l] + ((long long)(a.n0 + b) * T + tp) * H) + k4);
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = base + u * 256 + tid;
                    if (e < NL * per_layer) reinterpret_cast<float4*>(hs)[e] = v[u];
                }
            }
        }
        __syncthreads();
        for (int l = 0; l < NL; ++l) {
            const int t = s - l;
            if (t < 0 || t >= T) continue;
            float acc[2][LSTM_BC];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int b = 0; b < LSTM_BC; ++b) acc[i][b] = 0.f;
            const float* w0 = whs + ((size_t)l * 16 + r0) * H;
            const float* hl = hs + (size_t)l * LSTM_BC * H;
            for (int k = lane; k < H; k += 32) {
                const float wa = w0[k], wb = w0[H + k];
#pragma unroll
                for (int b = 0; b < LSTM_BC; ++b) {
                    const float hv = hl[b * H + k];
                    acc[0][b] = fmaf(wa, hv, acc[0][b]);
                    acc[1][b] = fmaf(wb, hv, acc[1][b]);
                }
            }
            if (l >= 1) {
                const float* x0 = wxs + ((size_t)(l - 1) * 16 + r0) * H;
                const float* hp = hs + (size_t)(l - 1) * LSTM_BC * H;
                for (int k = lane; k < H; k += 32) {
                    const float wa = x0[k], wb = x0[H + k];
#pragma unroll
                    for (int b = 0; b < LSTM_BC; ++b) {
                        const float hv = hp[b * H + k];
                        acc[0][b] = fmaf(wa, hv, acc[0][b]);
                        acc[1][b] = fmaf(wb, hv, acc[1][b]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int b = 0; b < LSTM_BC; ++b) {
                    float v = acc[i][b];
#pragma unroll
                    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    if (lane == 0) gates[(l * 16 + r0 + i) * LSTM_BC + b] = v;
                }
        }
        __syncthreads();
        // pointwise: thread -> (layer, unit, batch row)
        if (tid < NL * LSTM_UNITS * LSTM_BC) {
            const int l = tid / (LSTM_UNITS * LSTM_BC), u = (tid / LSTM_BC) % LSTM_UNITS, b = tid % LSTM_BC;
            const int t = s - l;
            if (t >= 0 && t < T && b < a.nb) {
                const long long row = (long long)(a.n0 + b) * T + t;
                float g[4];
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    float v = gates[(l * 16 + gi * LSTM_UNITS + u) * LSTM_BC + b];
                    const long long col = (long long)gi * H + u0 + u;
                    v += (l == 0) ? a.xproj[row * 4 * H + col] : a.bias[l][col];
                    g[gi] = v;
                }
                const float ig = sigmoidf_(g[0]), fg = sigmoidf_(g[1]), gg = tanhf(g[2]), og = sigmoidf_(g[3]);
                const float c = fg * cst[tid] + ig * gg;
                cst[tid] = c;
                const float h = og * tanhf(c);
                a.hseq[l][row * H + u0 + u] = h;
                if (l == NL - 1) a.out[row * H + u0 + u] = h + a.skip[row * H + u0 + u];
            }
        }
        grid_barrier(a.bar, (unsigned)(s + 1) * gridDim.x);
    }
}

}  // namespace ec
}  // namespace b2a

using namespace b2a;

struct EcConv {
    DBuf<float> A, bias;
    int M = 0, K = 0;
};

struct b2a_encodec {
    int device = 0, num_sms = 148;
    b2a_encodec_config cfg{};
    cudaStream_t stream = nullptr;
    int n_q = 0;
    DBuf<float> books;                       // [n_q][size][dim]
    EcConv conv0, xproj;
    DBuf<float> lWh[ec::LSTM_MAX_LAYERS], lWx[ec::LSTM_MAX_LAYERS], lb[ec::LSTM_MAX_LAYERS];
    struct Stage { int ratio, cin, cout, taps; EcConv up, r1, r2; };
    std::vector<Stage> stages;
    DBuf<float> wlast, blast;
    // workspaces
    DBuf<float> bufA, bufB, bufC, xp, hseq[ec::LSTM_MAX_LAYERS], chunks, scales, wave;
    DBuf<int> codes;
    DBuf<unsigned> bar;
    int dim0 = 0;

    static void up(DBuf<float>& d, const std::vector<float>& v) { d.upload(v.data(), v.size()); }

    b2a_encodec(int dev, const b2a_encodec_config& c, const TensorTable& tt) : device(dev), cfg(c) {
        require_device(dev);
        B2A_CUDA(cudaSetDevice(dev));
        cudaDeviceProp prop{};
        B2A_CUDA(cudaGetDeviceProperties(&prop, dev));
        num_sms = prop.multiProcessorCount;
        B2A_CHECK(c.norm_type == 0, B2A_ERR_INVALID_INPUT, "encodec: only norm_type weight_norm (folded weights) is implemented");
        B2A_CHECK(c.n_upsampling_ratios >= 1 && c.n_upsampling_ratios <= 8, B2A_ERR_INVALID_INPUT, "encodec: bad upsampling_ratios");
        B2A_CHECK(c.num_residual_layers == 1 || c.dilation_growth_rate == 1, B2A_ERR_INVALID_INPUT,
                  "encodec: dilated residual layers change the frame count in the reference (EncodecLayers.swift:117); not implemented");
        B2A_CHECK(c.num_lstm_layers >= 0 && c.num_lstm_layers <= ec::LSTM_MAX_LAYERS, B2A_ERR_INVALID_INPUT, "encodec: too many LSTM layers");
        B2A_CHECK(c.audio_channels >= 1 && c.audio_channels <= 2, B2A_ERR_INVALID_INPUT, "encodec: audio_channels must be 1 or 2");
        B2A_CHECK(c.compress >= 1 && c.kernel_size >= 1 && c.last_kernel_size >= 1 && c.residual_kernel_size >= 1, B2A_ERR_INVALID_INPUT, "encodec: bad config");
        B2A_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        // codebooks (as many as the checkpoint holds)
        while (tt.find("quantizer.layers." + std::to_string(n_q) + ".codebook.embed")) ++n_q;
        B2A_CHECK(n_q >= 1, B2A_ERR_MODEL_NOT_INITIALIZED, "missing tensor: quantizer.layers.0.codebook.embed");
        {
            std::vector<float> all((size_t)n_q * c.codebook_size * c.codebook_dim);
            for (int q = 0; q < n_q; ++q) {
                std::vector<float> e = tt.f32("quantizer.layers." + std::to_string(q) + ".codebook.embed", (int64_t)c.codebook_size * c.codebook_dim);
                memcpy(&all[(size_t)q * e.size()], e.data(), e.size() * sizeof(float));
            }
            up(books, all);
        }
        B2A_CHECK(c.codebook_dim == c.hidden_size, B2A_ERR_INVALID_INPUT, "encodec: codebook_dim must equal hidden_size");
        int scaling = 1 << c.n_upsampling_ratios;
        int idx = 0;
        auto key = [&](int i, const char* rest) { return "decoder.layers." + std::to_string(i) + "." + rest; };
        auto chan_ok = [&](int ch) { B2A_CHECK(ch >= 4 && ch % 4 == 0, B2A_ERR_INVALID_INPUT, "encodec: channel counts must be multiples of 4"); };
        auto plain = [&](EcConv& cv, const std::string& p, int cout, int k, int cin) {
            chan_ok(cin);
            cv.M = cout; cv.K = k * cin;
            up(cv.A, tt.f32(p + "conv.weight", (int64_t)cout * k * cin));      // [out, k, in] == [M, tap*Cin + ci]
            up(cv.bias, tt.f32(p + "conv.bias", cout));
        };
        dim0 = scaling * c.num_filters;
        plain(conv0, key(idx, ""), dim0, c.kernel_size, c.hidden_size); ++idx;
        if (c.num_lstm_layers > 0) {
            B2A_CHECK(dim0 % ec::LSTM_UNITS == 0, B2A_ERR_INVALID_INPUT, "encodec: LSTM width must be a multiple of 4");
            for (int l = 0; l < c.num_lstm_layers; ++l) {
                const std::string p = key(idx, "lstm.") + std::to_string(l) + ".";
                std::vector<float> wx = tt.f32(p + "Wx", (int64_t)4 * dim0 * dim0), wh = tt.f32(p + "Wh", (int64_t)4 * dim0 * dim0);
                std::vector<float> b = tt.find(p + "bias") ? tt.f32(p + "bias", 4 * dim0) : std::vector<float>(4 * dim0, 0.f);
                up(lWh[l], wh);
                if (l == 0) { xproj.M = 4 * dim0; xproj.K = dim0; up(xproj.A, wx); up(xproj.bias, b); }
                else { up(lWx[l], wx); up(lb[l], b); }
            }
        }
        ++idx;   // the LSTM block occupies a slot even when it has no layers
        for (int i = 0; i < c.n_upsampling_ratios; ++i) {
            Stage st{};
            st.ratio = c.upsampling_ratios[i];
            B2A_CHECK(st.ratio >= 1, B2A_ERR_INVALID_INPUT, "encodec: bad upsampling ratio");
            st.cin = scaling * c.num_filters; st.cout = st.cin / 2;
            chan_ok(st.cin); chan_ok(st.cout);
            ++idx;                                    // ELU slot
            {   // transposed conv k = 2*ratio: phase-major rows m = r*Cout + co, K = taps*Cin, tap j <-> kernel index r + j*s
                const int k = 2 * st.ratio, s = st.ratio;
                st.taps = (k + s - 1) / s;
                std::vector<float> w = tt.f32(key(idx, "conv.weight"), (int64_t)st.cout * k * st.cin), bsrc = tt.f32(key(idx, "conv.bias"), st.cout);
                std::vector<float> A((size_t)s * st.cout * st.taps * st.cin, 0.f), bb((size_t)s * st.cout);
                for (int r = 0; r < s; ++r)
                    for (int co = 0; co < st.cout; ++co) {
                        bb[(size_t)r * st.cout + co] = bsrc[co];
                        for (int j = 0; j < st.taps; ++j) {
                            const int kk = r + j * s;
                            if (kk >= k) continue;
                            memcpy(&A[(((size_t)r * st.cout + co) * st.taps + j) * st.cin], &w[((size_t)co * k + kk) * st.cin], st.cin * sizeof(float));
                        }
                    }
                st.up.M = s * st.cout; st.up.K = st.taps * st.cin;
                up(st.up.A, A); up(st.up.bias, bb);
                ++idx;
            }
            for (int j = 0; j < c.num_residual_layers; ++j) {
                B2A_CHECK(j == 0, B2A_ERR_INVALID_INPUT, "encodec: one residual layer per stage is implemented");
                const int dim = st.cout, hid = dim / c.compress;
                chan_ok(hid);
                plain(st.r1, key(idx, "block.1."), hid, c.residual_kernel_size, dim);
                // second launch: [shortcut | block.3] over K = dim + hid (x raw, hidden through ELU)
                std::vector<float> w1 = tt.f32(key(idx, "block.3.conv.weight"), (int64_t)dim * hid), b1 = tt.f32(key(idx, "block.3.conv.bias"), dim);
                if (c.use_conv_shortcut) {
                    std::vector<float> ws = tt.f32(key(idx, "shortcut.conv.weight"), (int64_t)dim * dim), bs = tt.f32(key(idx, "shortcut.conv.bias"), dim);
                    std::vector<float> A((size_t)dim * (dim + hid));
                    for (int m = 0; m < dim; ++m) {
                        memcpy(&A[(size_t)m * (dim + hid)], &ws[(size_t)m * dim], dim * sizeof(float));
                        memcpy(&A[(size_t)m * (dim + hid) + dim], &w1[(size_t)m * hid], hid * sizeof(float));
                        b1[m] += bs[m];
                    }
                    st.r2.M = dim; st.r2.K = dim + hid; up(st.r2.A, A);
                } else {
                    st.r2.M = dim; st.r2.K = hid; up(st.r2.A, w1);
                }
                up(st.r2.bias, b1);
                ++idx;
            }
            stages.push_back(std::move(st));
            scaling /= 2;
        }
        ++idx;   // ELU slot
        chan_ok(c.num_filters);
        up(wlast, tt.f32(key(idx, "conv.weight"), (int64_t)c.audio_channels * c.last_kernel_size * c.num_filters));
        up(blast, tt.f32(key(idx, "conv.bias"), c.audio_channels));
        bar.alloc(1);
        B2A_CUDA(cudaDeviceSynchronize());
    }
    ~b2a_encodec() {
        if (stream) cudaStreamDestroy(stream);
    }

    int hop() const {
        int h = 1;
        for (int i = 0; i < cfg.n_upsampling_ratios; ++i) h *= cfg.upsampling_ratios[i];
        return h;
    }
    int chunk_length() const { return cfg.chunk_length_s > 0.f ? (int)(cfg.chunk_length_s * (float)cfg.sampling_rate) : 0; }
    int chunk_stride() const {
        if (cfg.chunk_length_s <= 0.f || cfg.overlap < 0.f) return 0;
        return std::max(1, (int)((1.0f - cfg.overlap) * (float)chunk_length()));
    }
    long long out_len(int n_chunks, int T) const {
        const long long per = (long long)T * hop();
        if (chunk_length() == 0) return per;
        const int st = chunk_stride() > 0 ? chunk_stride() : 1;
        return (long long)st * (n_chunks - 1) + per;
    }

    void pads(int k, int& padL) const {
        const int total = k - 1;
        padL = cfg.use_causal_conv ? total : total - total / 2;
    }

    void run_conv(ec::ConvArgs a, cudaStream_t s) {
        const int M = a.M;
        if (M >= 64) {
            dim3 g(cdiv(a.Lq, 64), cdiv(M, 64), a.N);
            ec::ec_conv_kernel<64, 64><<<g, 256, 0, s>>>(a);
        } else if (M >= 32) {
            dim3 g(cdiv(a.Lq, 128), cdiv(M, 32), a.N);
            ec::ec_conv_kernel<32, 128><<<g, 256, 0, s>>>(a);
        } else {
            dim3 g(cdiv(a.Lq, 256), cdiv(M, 16), a.N);
            ec::ec_conv_kernel<16, 256><<<g, 256, 0, s>>>(a);
        }
        count_launch();
    }

    // d_codes [n_chunks, B, n_q_used, T] int32 (device), d_scales [n_chunks, B] or null -> d_wave [B, out_len, channels]
    void decode_dev(const int* d_codes, int n_chunks, int B, int nq, int T, const float* d_scales, float* d_wave, cudaStream_t s) {
        B2A_CHECK(n_chunks >= 1 && B >= 1 && T >= 1, B2A_ERR_INVALID_INPUT, "encodec decode: empty input");
        B2A_CHECK(nq >= 1 && nq <= n_q, B2A_ERR_INVALID_INPUT, "encodec decode: more codebooks than the checkpoint holds");
        B2A_CHECK(chunk_length() != 0 || n_chunks == 1, B2A_ERR_AUDIO_DECODING_FAILED, "Expected one frame");   // Encodec.swift:375-377
        B2A_CUDA(cudaSetDevice(device));
        const int N = n_chunks * B, CH = cfg.audio_channels;
        const long long Lfin = (long long)T * hop();
        B2A_CHECK((long long)N * Lfin * 64 < (1ll << 40) && Lfin < (1ll << 30), B2A_ERR_INVALID_INPUT, "encodec decode: too long");
        // the widest activation: max over stages of N * L * C
        size_t big = (size_t)N * T * std::max(dim0, 4 * dim0);
        {
            long long L = T;
            for (auto& st : stages) { L *= st.ratio; big = std::max(big, (size_t)((long long)N * L * st.cout)); }
        }
        bufA.alloc(big); bufB.alloc(big); bufC.alloc(big);
        float* x = bufA.p; float* y = bufB.p; float* z = bufC.p;
        // 1. RVQ decode
        ec::rvq_sum_kernel<<<(unsigned)((long long)N * T), 128, 0, s>>>(d_codes, books.p, x, nq, T, cfg.codebook_size, cfg.codebook_dim);
        count_launch();
        // 2. first conv
        {
            ec::ConvArgs a{};
            a.xa = x; a.La = T; a.Ca = cfg.hidden_size; a.taps = cfg.kernel_size; pads(cfg.kernel_size, a.padL); a.reflect = cfg.pad_mode_reflect;
            a.A = conv0.A.p; a.bias = conv0.bias.p; a.M = conv0.M; a.K = conv0.K; a.Lq = T; a.N = N; a.out = y; a.out_per_n = (long long)T * dim0;
            run_conv(a, s);
            std::swap(x, y);
        }
        // 3. LSTM block
        if (cfg.num_lstm_layers > 0) {
            const int H = dim0, NL = cfg.num_lstm_layers;
            xp.alloc((size_t)N * T * 4 * H);
            for (int l = 0; l < NL; ++l) hseq[l].alloc((size_t)N * T * H);
            {
                ec::ConvArgs a{};
                a.xa = x; a.La = T; a.Ca = H; a.taps = 1; a.A = xproj.A.p; a.bias = xproj.bias.p; a.M = 4 * H; a.K = H; a.Lq = T; a.N = N;
                a.out = xp.p; a.out_per_n = (long long)T * 4 * H;
                run_conv(a, s);
            }
            const size_t smem = ((size_t)(2 * NL - 1) * 16 * H + (size_t)NL * ec::LSTM_BC * H + NL * 16 * ec::LSTM_BC + NL * ec::LSTM_UNITS * ec::LSTM_BC) * sizeof(float);
            B2A_CHECK(smem <= 220 * 1024, B2A_ERR_INVALID_INPUT, "encodec: LSTM slice does not fit shared memory");
            B2A_CUDA(cudaFuncSetAttribute(ec::lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            const int grid = H / ec::LSTM_UNITS;
            int per_sm = 0;
            B2A_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ec::lstm_kernel, 256, smem));
            B2A_CHECK((long long)per_sm * num_sms >= grid, B2A_ERR_INVALID_INPUT, "encodec: LSTM too wide for a co-resident grid");
            for (int n0 = 0; n0 < N; n0 += ec::LSTM_BC) {
                ec::LstmArgs la{};
                la.xproj = xp.p; la.skip = x; la.out = y; la.bar = bar.p; la.n0 = n0; la.nb = std::min(ec::LSTM_BC, N - n0); la.T = T; la.H = H; la.NL = NL;
                for (int l = 0; l < NL; ++l) { la.Wh[l] = lWh[l].p; la.Wx[l] = lWx[l].p; la.bias[l] = lb[l].p; la.hseq[l] = hseq[l].p; }
                B2A_CUDA(cudaMemsetAsync(bar.p, 0, sizeof(unsigned), s));
                void* params[] = {&la};
                B2A_CUDA(cudaLaunchCooperativeKernel((void*)ec::lstm_kernel, dim3(grid), dim3(256), params, smem, s));
                count_launch();
            }
            std::swap(x, y);
        } else {
            // an EncodecLSTMBlock without layers still adds its skip: h + hiddenStates = 2x (EncodecLayers.swift:82-88)
            const long long cnt = (long long)N * T * dim0;
            ec::scale_kernel<<<cdiv(cnt, 256), 256, 0, s>>>(x, cnt, 2.f);
            count_launch();
        }
        // 4. upsampling stages
        long long L = T;
        for (auto& st : stages) {
            const int s_ = st.ratio, k = 2 * s_;
            const long long Lo = L * s_;                      // after the trim: (L-1)*s + k - (k - s)
            {
                const int padding_total = k - s_;
                const int pr = cfg.use_causal_conv ? (int)std::ceil((float)padding_total * cfg.trim_right_ratio) : padding_total / 2;
                const int pl = padding_total - pr;
                ec::ConvArgs a{};
                a.xa = x; a.La = (int)L; a.Ca = st.cin; a.taps = st.taps; a.backward = 1; a.elu_a = 1;
                a.A = st.up.A.p; a.bias = st.up.bias.p; a.M = st.up.M; a.K = st.up.K; a.Lq = (int)L + st.taps - 1; a.N = N;
                a.out = y; a.out_per_n = Lo * st.cout; a.shift = (long long)pl * st.cout;
                run_conv(a, s);
                std::swap(x, y);
            }
            L = Lo;
            {
                const int hid = st.r1.M;
                ec::ConvArgs a{};
                a.xa = x; a.La = (int)L; a.Ca = st.cout; a.taps = cfg.residual_kernel_size; pads(cfg.residual_kernel_size, a.padL); a.reflect = cfg.pad_mode_reflect; a.elu_a = 1;
                a.A = st.r1.A.p; a.bias = st.r1.bias.p; a.M = hid; a.K = st.r1.K; a.Lq = (int)L; a.N = N; a.out = z; a.out_per_n = L * hid;
                run_conv(a, s);
                ec::ConvArgs b{};
                b.N = N; b.Lq = (int)L; b.M = st.cout; b.K = st.r2.K; b.A = st.r2.A.p; b.bias = st.r2.bias.p; b.out = y; b.out_per_n = L * st.cout;
                if (cfg.use_conv_shortcut) {
                    b.xa = x; b.La = (int)L; b.Ca = st.cout; b.taps = 1; b.xb = z; b.Cb = hid; b.elu_b = 1;
                } else {
                    b.xa = z; b.La = (int)L; b.Ca = hid; b.taps = 1; b.elu_a = 1; b.res = x;
                }
                run_conv(b, s);
                std::swap(x, y);
            }
        }
        // 5. ELU -> last conv (+ per-chunk scale), 6. overlap-add when chunked
        const bool chunked = chunk_length() != 0;
        float* dst = d_wave;
        if (chunked) { chunks.alloc((size_t)N * L * CH); dst = chunks.p; }
        {
            int padL; pads(cfg.last_kernel_size, padL);
            const int C = cfg.num_filters, k = cfg.last_kernel_size;
            const size_t smem = ((size_t)(256 + k - 1) * (C + 1) + (size_t)CH * k * C) * sizeof(float);
            B2A_CHECK(smem <= 200 * 1024, B2A_ERR_INVALID_INPUT, "encodec: last conv too wide");
            B2A_CUDA(cudaFuncSetAttribute(ec::final_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ec::final_conv_kernel<<<dim3(cdiv(L, 256), N), 256, smem, s>>>(x, wlast.p, blast.p, d_scales, dst, (int)L, C, k, CH, padL, cfg.pad_mode_reflect);
            count_launch();
        }
        if (chunked) {
            const long long total = out_len(n_chunks, T);
            const int hopc = chunk_stride() > 0 ? chunk_stride() : 1;
            ec::overlap_add_kernel<<<dim3(cdiv(total, 256), B), 256, 0, s>>>(chunks.p, d_wave, n_chunks, B, (int)L, CH, hopc, total);
            count_launch();
        }
        B2A_CUDA(cudaGetLastError());
    }
};

extern "C" {

int32_t b2a_encodec_create(int32_t device, const b2a_encodec_config* cfg, const b2a_tensor* tensors, int32_t n, b2a_encodec** out) {
    return guarded([&] {
        B2A_CHECK(out, B2A_ERR_INVALID_INPUT, "b2a_encodec_create: null out");
        *out = nullptr;
        B2A_CHECK(cfg && tensors && n > 0, B2A_ERR_MODEL_NOT_INITIALIZED, "b2a_encodec_create: missing config or weights");
        TensorTable tt(tensors, n);
        *out = new b2a_encodec(device, *cfg, tt);
    });
}

int64_t b2a_encodec_output_length(const b2a_encodec* h, int32_t n_chunks, int32_t frames) {
    return h && n_chunks >= 1 && frames >= 1 ? h->out_len(n_chunks, frames) : 0;
}
int32_t b2a_encodec_num_codebooks(const b2a_encodec* h) { return h ? h->n_q : 0; }
void* b2a_encodec_stream(b2a_encodec* h) { return h ? (void*)h->stream : nullptr; }

int32_t b2a_encodec_decode_dev(b2a_encodec* h, const int32_t* d_codes, int32_t n_chunks, int32_t B, int32_t nq, int32_t T,
                               const float* d_scales, float* d_wave, void* stream) {
    return guarded([&] {
        B2A_CHECK(h && d_codes && d_wave, B2A_ERR_INVALID_INPUT, "b2a_encodec_decode_dev: null argument");
        h->decode_dev(d_codes, n_chunks, B, nq, T, d_scales, d_wave, stream ? (cudaStream_t)stream : h->stream);
    });
}

int32_t b2a_encodec_decode(b2a_encodec* h, const int32_t* codes, int32_t n_chunks, int32_t B, int32_t nq, int32_t T,
                           const float* scales, float* wave) {
    return guarded([&] {
        B2A_CHECK(h && codes && wave, B2A_ERR_INVALID_INPUT, "b2a_encodec_decode: null argument");
        B2A_CHECK(n_chunks >= 1 && B >= 1 && T >= 1 && nq >= 1, B2A_ERR_AUDIO_DECODING_FAILED, "b2a_encodec_decode: empty codes");
        B2A_CUDA(cudaSetDevice(h->device));
        cudaStream_t s = h->stream;
        const size_t nin = (size_t)n_chunks * B * nq * T, nout = (size_t)B * h->out_len(n_chunks, T) * h->cfg.audio_channels;
        h->codes.alloc(nin); h->wave.alloc(nout);
        B2A_CUDA(cudaMemcpyAsync(h->codes.p, codes, nin * sizeof(int), cudaMemcpyHostToDevice, s));
        const float* dsc = nullptr;
        if (scales) {
            h->scales.alloc((size_t)n_chunks * B);
            B2A_CUDA(cudaMemcpyAsync(h->scales.p, scales, (size_t)n_chunks * B * sizeof(float), cudaMemcpyHostToDevice, s));
            dsc = h->scales.p;
        }
        h->decode_dev(h->codes.p, n_chunks, B, nq, T, dsc, h->wave.p, s);
        B2A_CUDA(cudaMemcpyAsync(wave, h->wave.p, nout * sizeof(float), cudaMemcpyDeviceToHost, s));
        B2A_CUDA(cudaStreamSynchronize(s));
    });
}

void b2a_encodec_destroy(b2a_encodec* h) { delete h; }

}  // extern "C"
