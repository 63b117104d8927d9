/*
 * b200audio_internal.h -- NOT part of the drop-in boundary.  Test and benchmark hooks of libb200audio.so that have no
 * counterpart in the reference: device-side random-init constructors for BASELINE.json's full-size configurations (there are no
 * checkpoints here and a 3B-parameter host copy is pointless), fixed-work switches for bench.py, parity trace hooks and the
 * single-kernel test entries.  A Swift wrapper binds include/b200audio.h only; tests/, bench.py and tools/ may
 * also bind these.
 */
#ifndef B200AUDIO_INTERNAL_H
#define B200AUDIO_INTERNAL_H

#include "b200audio.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ benchmark constructors / switches
 *   b2a_tts_create_random / b2a_stt_create_random : same as b2a_tts_create / b2a_stt_create but the weights are drawn ON THE
 *       DEVICE (N(0, std^2) bf16 from a counter-based generator, norm gains 1).
 *   b2a_tts_set_bench_flags : mask_eos != 0 -> never stop on 128258 (fixed work per call); wrap_codes != 0 -> audio codes are
 *       taken mod 4096 per slot so that random-init tokens index the SNAC codebooks.  Both default to 0 (reference behaviour).
 *   b2a_stt_set_bench_flags : mask_eot != 0 -> a clip never stops on end-of-text (fixed work).
 *   b2a_tts_time_steps : runs `iters` captured decode steps for `batch` rows at context `ctx` (greedy, no host sync inside)
 *       between two CUDA events on the handle's stream; *ms_per_step = average device time of one step.                      */
int32_t b2a_tts_create_random(int32_t device, const b2a_llama_config* cfg, float std, uint64_t seed,
                              b2a_snac* snac, b2a_tts** out);
int32_t b2a_stt_create_random(int32_t device, const b2a_whisper_config* cfg, float std, uint64_t seed, b2a_stt** out);
int32_t b2a_tts_set_bench_flags(b2a_tts* h, int32_t mask_eos, int32_t wrap_codes);
int32_t b2a_stt_set_bench_flags(b2a_stt* h, int32_t mask_eot);
int32_t b2a_tts_time_steps(b2a_tts* h, int32_t batch, int32_t ctx, int32_t iters, float* ms_per_step);
/*   b2a_qwen3_talker_create_random : the Qwen3-TTS talker + code predictor with device-drawn weights (BASELINE config 5 bench).
 *   b2a_qwen3_talker_set_bench_flags : mask_eos != 0 -> the talker never emits codec_eos_token_id (fixed work per call).        */
int32_t b2a_qwen3_talker_create_random(int32_t device, const b2a_qwen3_talker_config* cfg, float std, uint64_t seed, b2a_qwen3_talker** out);
int32_t b2a_qwen3_talker_set_bench_flags(b2a_qwen3_talker* h, int32_t mask_eos);

/* ------------------------------------------------------------------ parity hooks
 *   b2a_tts_debug_trace : enable != 0 makes later b2a_tts_forward_logits calls record the residual stream at every RMSNorm
 *       input; out (nullable) receives the record of the last traced position as [2*layers+1, batch, hidden] float32.        */
int32_t b2a_tts_debug_trace(b2a_tts* h, int32_t enable, int32_t batch, float* out);

/* Host-only (no device needed): the GEMM weight matrix the implicit convolution reads for an MLX-layout [out, k, in] weight --
 * stride 0: causal conv, rows = out, taps = k; stride > 0: transposed conv with k = n * stride, rows = stride * out
 * (phase-major), taps = n.  layout_out: [rows][taps][kpad] float32, kpad = ceil(in / 64) * 64.                              */
int32_t b2a_speech_tokenizer_debug_layout(const float* w, int32_t out, int32_t k, int32_t in, int32_t stride, float* layout_out,
                                          int64_t capacity, int32_t* rows, int32_t* taps, int32_t* kpad);
/* tools/diag_n1_stages.py: stage >= 0 makes later decodes keep a copy of the fp32 activation tensor after that stage (0 = transformer
 * output before the final norm, 1 + i = upsample layer i, 10 + 4 b = decoder block b after its transposed conv, 11 + 4 b + j = after
 * its residual unit j); out != null first copies the last kept tensor to the host (capacity in floats, length in *n).            */
int32_t b2a_speech_tokenizer_debug_stage(b2a_speech_tokenizer* h, int32_t stage, float* out, int64_t capacity, int64_t* n);
/* tests/test_gpu_qwen3_sampler.py: the Qwen3-TTS in-graph sampler kernel (csrc/qwen3_sampler.cu = sampleToken,
 * Qwen3TTS.swift:1003-1118) on HOST logits [B, V <= 4096]; suppress [lo, hi) except eos; seen = bitmap of the tokens generated
 * so far [B, ceil(V/32)] (nullable; updated when track != 0); tokens_out [B]; filtered_out [B, V] (nullable) = the logits handed
 * to categorical, -inf where removed.                                                                                     */
int32_t b2a_qwen3_sample_test(const float* logits, int32_t batch, int32_t vocab, float temperature, float top_p, int32_t top_k,
                              float min_p, float repetition_penalty, int32_t eos, int32_t suppress_lo, int32_t suppress_hi,
                              uint32_t* seen, int32_t track, uint64_t seed, int32_t step, int32_t* tokens_out, float* filtered_out);
/* tests/test_gpu_implicit_conv.py: one launch of the implicit-GEMM causal convolution kernel (csrc/implicit_conv.cuh) on HOST
 * data: w [M][taps][Cin], x [B][Ttot][Cin]; out[b, t*up + rho, co] for m = rho * (M/up) + co is
 * sum_j sum_c w[m, j, c] * x[b, t + shift0 + j*dil, c] through the fused epilogue (bias, bias twice at t = 0, GELU, gamma, add,
 * SnakeBeta on the hi/lo copy).  xo [B][T*up][M/up] in/out or null; hl_out [B][Hout + T*up][M/up] or null.  fp16 != 0: operands
 * as fp16 hi/lo pairs instead of bf16 ones.                                                                               */
int32_t b2a_implicit_conv_test(const float* w, int32_t M, int32_t taps, int32_t Cin, const float* x, int32_t B, int32_t Ttot, int32_t T,
                               int32_t dil, int32_t shift0, int32_t up, const float* bias, const float* gamma, int32_t gelu, int32_t add,
                               int32_t bias_twice_t0, const float* sa, const float* sb, int32_t Hout, int32_t fp16, float* xo, float* hl_out);

#ifdef __cplusplus
}
#endif
#endif /* B200AUDIO_INTERNAL_H */
