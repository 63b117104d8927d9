/*
 * b200audio.h -- C ABI of libb200audio.so: the B200-native (sm_100a) hot path behind
 * MLXAudio's Swift protocols.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the
 * Blaizzy/mlx-audio-swift checkout).  INTEGRATION.md shows the Swift-side binding.
 *
 * Conventions
 *   - every function returns an int32 status (B2A_OK == 0).  Codes 1..5 map 1:1 onto the cases
 *     of AudioGenerationError (Sources/MLXAudioCore/Generation/GenerationTypes.swift:66-87);
 *     the library never aborts the process.  b2a_last_error() returns the message of the last
 *     failure on the calling thread.
 *   - handles are opaque, own their device memory and CUDA stream, and are NOT thread-safe:
 *     one in-flight call per handle ("Threading").
 *   - the caller owns every host buffer.  `_dev` variants take DEVICE pointers (already resident
 *     in HBM) plus a cudaStream_t passed as void*; all other variants take HOST pointers and do
 *     the host<->device copies themselves.
 *   - there is NO CPU fallback: with no usable CUDA device every create/compute call fails with
 *     B2A_ERR_CUDA.
 */
#ifndef B200AUDIO_H
#define B200AUDIO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2A_OK 0
#define B2A_ERR_MODEL_NOT_INITIALIZED 1 /* AudioGenerationError.modelNotInitialized */
#define B2A_ERR_GENERATION_FAILED 2     /* .generationFailed   */
#define B2A_ERR_INVALID_INPUT 3         /* .invalidInput       */
#define B2A_ERR_AUDIO_DECODING_FAILED 4 /* .audioDecodingFailed */
#define B2A_ERR_AUDIO_ENCODING_FAILED 5 /* .audioEncodingFailed */
#define B2A_ERR_CANCELLED 6             /* Task.checkCancellation (LlamaTTS.swift:715) */
#define B2A_ERR_CUDA 7                  /* no device / CUDA runtime failure */

#define B2A_DTYPE_F32 0
#define B2A_DTYPE_BF16 1
#define B2A_DTYPE_I32 2

/* One named host tensor (row-major).  Names follow the reference's safetensors keys. */
typedef struct b2a_tensor {
    const char* name;
    int32_t dtype;
    int32_t ndim;
    int64_t shape[4];
    const void* data;
} b2a_tensor;

const char* b2a_last_error(void);
const char* b2a_version(void);
/* number of visible CUDA devices (0 when none); never fails */
int32_t b2a_device_count(void);
/* kernels launched by this library on the calling process so far (for bench.py's gpu_launches) */
int64_t b2a_launch_count(void);

/* ------------------------------------------------------------------ DSP tables (host math)
 * hanningWindow  Sources/MLXAudioCore/DSP.swift:15-22  (symmetric; periodic!=0 gives the
 *                WhisperAudio.swift:42-43 window)
 * melFilters     Sources/MLXAudioCore/DSP.swift:76-168 ; out is [n_fft/2+1, n_mels] row-major;
 *                f_max < 0 means sample_rate/2; mel_scale 0 = htk, 1 = slaney; norm_slaney 0/1. */
int32_t b2a_hanning_window(int32_t size, int32_t periodic, float* out);
int32_t b2a_mel_filters(int32_t sample_rate, int32_t n_fft, int32_t n_mels, float f_min, float f_max,
                        int32_t norm_slaney, int32_t mel_scale, float* out);
/* The other two host helpers of DSP.swift, off the mel path but with known answers in the reference's tests
 * (Tests/MLXAudioCodecsTests.swift:117-140): hammingWindow (:25-42; periodic != 0 is the default) and powerToDB (:61-73;
 * top_db < 0 means no dynamic-range clipping).                                                                    */
int32_t b2a_hamming_window(int32_t size, int32_t periodic, float* out);
int32_t b2a_power_to_db(const float* spectrogram, int64_t n, float amin, float top_db, float* out);

/* ------------------------------------------------------------------ streaming log-mel
 * Replaces class IncrementalMelSpectrogram
 *   (Sources/MLXAudioSTT/Streaming/IncrementalMelSpectrogram.swift:18-208):
 *   init(sampleRate:nFft:hopLength:nMels:) :43-62 -> b2a_mel_create
 *   process(samples:) :68-147 -> b2a_mel_process   (*n_frames == 0  <=>  reference returns nil)
 *   flush() :151-200          -> b2a_mel_flush
 *   reset() :203-208          -> b2a_mel_reset
 *   totalFrames :41           -> b2a_mel_total_frames
 * `out` receives [n_frames, n_mels] float32; out_cap_frames is its capacity in frames
 * (B2A_ERR_INVALID_INPUT if too small; b2a_mel_max_frames() gives a safe bound).           */
typedef struct b2a_mel b2a_mel;
int32_t b2a_mel_create(int32_t device, int32_t sample_rate, int32_t n_fft, int32_t hop_length,
                       int32_t n_mels, b2a_mel** out);
int64_t b2a_mel_max_frames(const b2a_mel* h, int64_t n_samples);
int32_t b2a_mel_process(b2a_mel* h, const float* samples, int64_t n_samples, float* out,
                        int64_t out_cap_frames, int64_t* n_frames);
int32_t b2a_mel_flush(b2a_mel* h, float* out, int64_t out_cap_frames, int64_t* n_frames);
int32_t b2a_mel_reset(b2a_mel* h);
int64_t b2a_mel_total_frames(const b2a_mel* h);
void b2a_mel_destroy(b2a_mel* h);

/* ------------------------------------------------------------------ offline / batched log-mel
 * kind 0: computeMelSpectrogram (Sources/MLXAudioCore/DSP.swift:230-273): symmetric Hann, HTK
 *         scale, reflect pad both sides, global max-8 clamp, out [B, 1+n/hop, n_mels].
 * kind 1: WhisperAudio.encoderFeatures (Sources/MLXAudioSTT/Models/Whisper/WhisperAudio.swift:
 *         7-13,38-87): pad/trim each clip to 480000, periodic Hann, Slaney scale, last frame
 *         dropped, per-clip max-8 clamp, out [B, 3000, n_mels].
 * pcm is [B, n_samples] (every clip the same length); b2a_logmel_frames gives frames per clip. */
typedef struct b2a_logmel b2a_logmel;
int32_t b2a_logmel_create(int32_t device, int32_t kind, int32_t sample_rate, int32_t n_fft,
                          int32_t hop_length, int32_t n_mels, b2a_logmel** out);
int64_t b2a_logmel_frames(const b2a_logmel* h, int64_t n_samples);
int32_t b2a_logmel_compute(b2a_logmel* h, const float* pcm, int32_t batch, int64_t n_samples, float* out);
int32_t b2a_logmel_compute_dev(b2a_logmel* h, const float* d_pcm, int32_t batch, int64_t n_samples,
                               float* d_out, void* stream);
void b2a_logmel_destroy(b2a_logmel* h);

/* ------------------------------------------------------------------ SNAC codec
 * Replaces class SNAC (Sources/MLXAudioCodecs/SNAC/SNACDecoder.swift:12-131) behind the
 * AudioCodecModel protocol (Sources/MLXAudioCodecs/AudioCodecModel.swift:4-27):
 *   SNAC.fromConfig/fromModelDirectory :135-189 -> b2a_snac_create (config + named tensors)
 *   decode(_ codes:) :127-131 / decodeAudio :199 -> b2a_snac_decode
 *   quantizer(z) (ResidualVectorQuantize.callAsFunction, SNAC/VQ.swift:150-163), the
 *   encode-side code search                      -> b2a_snac_quantize
 * codes[i] is [B, T_i] int32 with T_i = t_latent / vq_strides[i]; wave is [B, 1, t_latent*hop].
 * noise[i] (nullable array of nullable pointers) is the [B, 1, T] Gaussian draw of decoder
 * block i's NoiseBlock (Layers.swift:263-279); NULL = draw on device from `seed`
 * (noise_mode 0) or use zero noise (noise_mode 1).                                          */
typedef struct b2a_snac_config {
    int32_t sampling_rate;
    int32_t encoder_dim;
    int32_t n_encoder_rates;
    int32_t encoder_rates[8];
    int32_t latent_dim; /* 0 => encoder_dim * 2^n_encoder_rates */
    int32_t decoder_dim;
    int32_t n_decoder_rates;
    int32_t decoder_rates[8];
    int32_t attn_window_size; /* 0 => none (only value supported) */
    int32_t codebook_size;
    int32_t codebook_dim;
    int32_t n_vq_strides;
    int32_t vq_strides[8];
    int32_t noise;
    int32_t depthwise;
} b2a_snac_config;

typedef struct b2a_snac b2a_snac;
int32_t b2a_snac_create(int32_t device, const b2a_snac_config* cfg, const b2a_tensor* tensors,
                        int32_t n_tensors, b2a_snac** out);
int64_t b2a_snac_hop_length(const b2a_snac* h);
void* b2a_snac_stream(b2a_snac* h); /* the handle's cudaStream_t, for event timing */
int32_t b2a_snac_decode(b2a_snac* h, const int32_t* const* codes, int32_t batch, int64_t t_latent,
                        const float* const* noise, int32_t noise_mode, uint64_t seed, float* wave);
int32_t b2a_snac_decode_dev(b2a_snac* h, const int32_t* const* d_codes, int32_t batch, int64_t t_latent,
                            const float* const* d_noise, int32_t noise_mode, uint64_t seed,
                            float* d_wave, void* stream);
/* z [B, latent_dim, T] float32 -> codes[i] [B, T/stride_i] int32 (+ optional z_q [B, latent, T]) */
int32_t b2a_snac_quantize(b2a_snac* h, const float* z, int32_t batch, int64_t t_latent,
                          int32_t* const* codes, float* z_q);
void b2a_snac_destroy(b2a_snac* h);

/* ------------------------------------------------------------------ Orpheus / Llama TTS
 * Replaces class LlamaTTSModel (Sources/MLXAudioTTS/Models/Llama/LlamaTTS.swift:354-977) behind
 * SpeechGenerationModel (Sources/MLXAudioTTS/Generation.swift:8-39):
 *   fromModelDirectory :942-977 (weights + config)   -> b2a_tts_create
 *   callAsFunction(_:cache:) :557-567                -> b2a_tts_forward_logits (parity hook)
 *   generate(text:voice:...) :658-765                -> b2a_tts_generate
 *   generateStream :777-913 (.token/.info/.audio)    -> b2a_tts_generate + on_token callback
 *   Task cancellation :715,911                       -> b2a_tts_cancel
 * Tokenisation stays host-side: the ABI takes token ids already framed by prepareInputIds
 * (:446-553; b2a_tts_prepare_input_ids does the framing for raw text-token ids).
 * A batch is B independent utterances ("batched == serial"; the reference itself is batch-1). */
typedef struct b2a_llama_config {
    int32_t hidden_size;
    int32_t num_hidden_layers;
    int32_t intermediate_size;
    int32_t num_attention_heads;
    int32_t num_key_value_heads;
    int32_t head_dim;
    int32_t vocab_size;
    float rms_norm_eps;
    float rope_theta;
    float rope_factor; /* llama3 rope_scaling (LlamaTTS.swift:114-118) */
    float rope_low_freq_factor;
    float rope_high_freq_factor;
    float rope_old_context_len;
    int32_t tie_word_embeddings;
    int32_t max_batch;   /* KV-cache rows */
    int32_t max_context; /* KV-cache positions per row */
} b2a_llama_config;

/* GenerateParameters as used at LlamaTTS.swift:573-581 (defaults 1200 / 0.6 / 0.8 / 1.3 / 20) */
typedef struct b2a_gen_params {
    int32_t max_tokens;
    float temperature; /* 0 => greedy argmax, lowest index wins ties */
    float top_p;
    float repetition_penalty;
    int32_t repetition_context_size;
    uint64_t seed;
} b2a_gen_params;

/* AudioGenerationInfo (GenerationTypes.swift:14-45) */
typedef struct b2a_gen_info {
    int32_t prompt_token_count;
    int32_t generation_token_count;
    double prefill_time;
    double generate_time;
    double tokens_per_second;
    double codec_time;
    double peak_memory_gb;
} b2a_gen_info;

typedef struct b2a_tts b2a_tts;
/* on_token(user, utterance, step, token): the .token(Int) events of generateStream (:862) */
typedef void (*b2a_token_cb)(void* user, int32_t utterance, int32_t step, int32_t token);

int32_t b2a_tts_create(int32_t device, const b2a_llama_config* cfg, const b2a_tensor* tensors,
                       int32_t n_tensors, b2a_snac* snac /* borrowed, may be NULL */, b2a_tts** out);
/* [SOH] ids [EOT, EOH], left-padded with 128263 to the longest prompt; out is [B, max_len+3] */
int32_t b2a_tts_prepare_input_ids(const int32_t* const* prompt_ids, const int32_t* lens, int32_t batch,
                                  int32_t* out, int32_t* out_len);
/* One forward over ids [B, L] appended at the cache's current offset (reset_cache != 0 clears
 * it first); logits_out [B, L, vocab] float32 (host). */
int32_t b2a_tts_forward_logits(b2a_tts* h, const int32_t* ids, int32_t batch, int32_t len,
                               int32_t reset_cache, float* logits_out);
/* Full text-token-ids -> tokens -> (parseOutput, 7-token de-interleave, SNAC decode) -> waveform.
 * input_ids [B, L] (all rows length L, as produced by b2a_tts_prepare_input_ids or the host).
 * tokens_out [B, max_tokens] (nullable) receives generated ids, n_tokens_out[B] their counts.
 * wave_out [B, wave_cap] (nullable => skip the codec) receives each waveform, wave_len[B] its
 * sample count.  Rows that yield no audio codes report wave_len 0; if every row does the call
 * fails with B2A_ERR_GENERATION_FAILED ("No audio codes generated", LlamaTTS.swift:752-754).  */
int32_t b2a_tts_generate(b2a_tts* h, const int32_t* input_ids, int32_t batch, int32_t len,
                         const b2a_gen_params* params, int32_t* tokens_out, int32_t* n_tokens_out,
                         float* wave_out, int64_t wave_cap, int64_t* wave_len, b2a_gen_info* info,
                         b2a_token_cb on_token, void* user);
/* generateStream with audio DURING generation.  The reference's Orpheus emits one .audio event at the end
 * (LlamaTTS.swift:901-904); its streaming models decode their codes in chunks while generating (decodeAudioFromCodes' chunk loop,
 * Sources/MLXAudioTTS/Models/Qwen3/Qwen3.swift:47-83; the CLI's --benchmark TTFB is the latency of the first .audio event,
 * Sources/Tools/mlx-audio-swift-tts/App.swift:155-211).  Here every `frames_per_chunk` new 7-token frames of a row are decoded by
 * SNAC with `left_context_frames` already-emitted frames in front of them (their samples are dropped) and handed to on_audio as
 * 2048 * frames float32 samples; is_final marks a row's last chunk.  Concatenated chunks equal the one-shot waveform except near
 * chunk boundaries (the codec's receptive field), exactly like the reference's chunked decode.                                 */
typedef void (*b2a_audio_cb)(void* user, int32_t utterance, const float* samples, int64_t n_samples, int32_t is_final);
int32_t b2a_tts_generate_stream(b2a_tts* h, const int32_t* input_ids, int32_t batch, int32_t len, const b2a_gen_params* params,
                                int32_t frames_per_chunk, int32_t left_context_frames, int32_t* tokens_out, int32_t* n_tokens_out,
                                b2a_gen_info* info, b2a_token_cb on_token, b2a_audio_cb on_audio, void* user);
/* Device-resident variant for bench.py's `value`: ids already in HBM, waveform left in HBM. */
int32_t b2a_tts_generate_dev(b2a_tts* h, const int32_t* d_input_ids, int32_t batch, int32_t len,
                             const b2a_gen_params* params, float* d_wave_out, int64_t wave_cap,
                             int64_t* wave_len, b2a_gen_info* info);
int32_t b2a_tts_cancel(b2a_tts* h);
/* the handle's cudaStream_t (as void*), so a caller can order its own work / record events on the stream the kernels run on */
void* b2a_tts_stream(b2a_tts* h);
/* parseOutput (:383-434) and llamaDecodeAudioFromCodes' de-interleave (:41-63), host-side ints.
 * tokens [B, n]; code_lists_out [B, n] / code_lens[B]; then per row codes0/1/2 sized n/7, 2n/7, 4n/7 */
int32_t b2a_tts_parse_output(const int32_t* tokens, int32_t batch, int32_t n, int32_t* code_lists_out,
                             int32_t* code_lens);
int32_t b2a_tts_deinterleave(const int32_t* code_list, int32_t n, int32_t* codes0, int32_t* codes1,
                             int32_t* codes2, int32_t* n_frames);
int32_t b2a_tts_interleave(const int32_t* codes0, const int32_t* codes1, const int32_t* codes2,
                           int32_t n_frames, int32_t* code_list);
void b2a_tts_destroy(b2a_tts* h);

/* ------------------------------------------------------------------ weight / format plumbing
 * Host-only.  Replaces MLX.loadArrays on *.safetensors (llamaTTSLoadWeights, LlamaTTS.swift:982-994: every file of a directory, later
 * files win), WhisperModel.detectFormat / sanitize / remapMlxWhisperKey / whisperSinusoids (WhisperModel.swift:315-480),
 * LlamaTTSModel.sanitize (LlamaTTS.swift:583-593) and the MLX affine de-quantisation behind quantize(model:) (:955-966).
 * A b2a_weights handle keeps the files mapped; b2a_weights_get returns borrowed views valid until b2a_weights_free.
 * F16 tensors are widened to F32, I64 narrowed to I32, U32 (packed quantised words) is reported as B2A_DTYPE_I32.
 *   b2a_weights_sanitize_whisper: -> HF (`transformers`) names with the "model." prefix, conv weights in the PyTorch [out, in, k]
 *     layout (what b2a_stt_create takes), missing encoder positions synthesised; *format = 0 huggingFace, 1 mlxWhisper.
 *   b2a_weights_sanitize_llama: drops rotary inv_freq (and lm_head.weight when tied); bits > 0: every layer with a ".scales"
 *     tensor is expanded to bf16 (w = scales * q + biases, value j of a uint32 word at bits [j*bits, (j+1)*bits); bits 2 / 4 / 8).
 *   b2a_tts_config_from_json: config.json -> the b2a_llama_config struct, LlamaTTSConfig.swift:100-166, + the "quantization" block.
 *   b2a_tts_create_from_directory = LlamaTTSModel.fromModelDirectory (LlamaTTS.swift:942-977) without tokenizer / SNAC download. */
typedef struct b2a_weights b2a_weights;
int32_t b2a_weights_load(const char* file_or_directory, b2a_weights** out);
int32_t b2a_weights_count(const b2a_weights* w);
int32_t b2a_weights_get(const b2a_weights* w, int32_t index, b2a_tensor* out);
int32_t b2a_weights_sanitize_whisper(b2a_weights* w, int32_t* format);
int32_t b2a_weights_sanitize_llama(b2a_weights* w, int32_t tie_word_embeddings, int32_t group_size, int32_t bits);
/* Same, with the quantisation read from config.json the way the reference's loader does (LlamaTTS.swift:955-966 through
 * mlx-swift-lm's PerLayerQuantization): "quantization": {"group_size", "bits", "<layer path>": false | {"group_size", "bits"}} --
 * per-layer settings override the default, a layer marked false must not carry .scales.  b2a_tts_create_from_directory uses this. */
int32_t b2a_weights_sanitize_llama_config(b2a_weights* w, const char* config_path);
/* MLX affine de-quantisation (to bf16) of every layer that carries "<path>.scales", with one group_size / bits -- what
 * WhisperModel.fromDirectory's quantize(model:groupSize:bits:) implies for a quantised checkpoint (WhisperModel.swift:499-511:
 * every Linear and decoder.embed_tokens; the tied projection then multiplies by the de-quantised embedding,
 * Tests/WhisperQuantizedTiedEmbeddingTests.swift).  Call after b2a_weights_sanitize_whisper.                              */
int32_t b2a_weights_dequantize(b2a_weights* w, int32_t group_size, int32_t bits);
void b2a_weights_free(b2a_weights* w);
int32_t b2a_tts_config_from_json(const char* config_path, int32_t max_batch, int32_t max_context, b2a_llama_config* cfg,
                                 int32_t* quant_group_size, int32_t* quant_bits);
int32_t b2a_tts_create_from_directory(const char* model_dir, int32_t device, int32_t max_batch, int32_t max_context,
                                      b2a_snac* snac, b2a_tts** out);

/* ------------------------------------------------------------------ Vocos vocoder
 * Replaces class Vocos (Sources/MLXAudioCodecs/Vocos/Vocos.swift:284-322) behind AudioDecoderModel
 * (Sources/MLXAudioCodecs/AudioCodecModel.swift:4-13):
 *   Vocos(backbone:head:) + weights (keys backbone.* / head.*, MLX layouts) -> b2a_vocos_create
 *   decode(_ features:) / decodeAudio (:302-306,318-320)                   -> b2a_vocos_decode
 * features are [B, L, input_channels] float32 (the layout VocosBackbone expects, VocosBackbone.swift:170-175);
 * the waveform is [B, (L-1)*hop_length] (ISTFTHead centre trim, Vocos.swift:150-158).  AdaLayerNorm models
 * (Vocos.swift:17-47; weights backbone.norm.{scale,shift}.{weight,bias}, backbone.convnext.N.norm.{scale,shift}.*) take their
 * `bandwidthId` conditioning through b2a_vocos_decode_cond: cond [B, adanorm_num_embeddings] float32, the rows the scale / shift
 * Linears are applied to (decode(_:bandwidthId:) :302-306); decoding such a model without it fails with invalidInput where the
 * reference fatalErrors (VocosBackbone.swift:66-68,181-183).                                        */
typedef struct b2a_vocos_config {
    int32_t input_channels;
    int32_t dim;
    int32_t intermediate_dim;
    int32_t num_layers;
    int32_t n_fft;
    int32_t hop_length;
    int32_t input_kernel_size;
    int32_t dw_kernel_size;
    int32_t adanorm_num_embeddings; /* 0: LayerNorm; > 0: AdaLayerNorm(numEmbeddings, dim) for backbone.norm and every block's norm */
} b2a_vocos_config;

typedef struct b2a_vocos b2a_vocos;
int32_t b2a_vocos_create(int32_t device, const b2a_vocos_config* cfg, const b2a_tensor* tensors, int32_t n_tensors,
                         b2a_vocos** out);
int64_t b2a_vocos_output_length(const b2a_vocos* h, int32_t frames);
void* b2a_vocos_stream(b2a_vocos* h);
int32_t b2a_vocos_decode(b2a_vocos* h, const float* features, int32_t batch, int32_t frames, float* wave);
int32_t b2a_vocos_decode_dev(b2a_vocos* h, const float* d_features, int32_t batch, int32_t frames, float* d_wave, void* stream);
int32_t b2a_vocos_decode_cond(b2a_vocos* h, const float* features, const float* cond, int32_t batch, int32_t frames, float* wave);
void b2a_vocos_destroy(b2a_vocos* h);

/* ------------------------------------------------------------------ Encodec decode
 * Replaces class Encodec's decode side (Sources/MLXAudioCodecs/Encodec/Encodec.swift:170-402) behind AudioCodecModel /
 * AudioDecoderModel (Sources/MLXAudioCodecs/AudioCodecModel.swift:4-27, conformance at Encodec.swift:447-461):
 *   Encodec(config:) + fromModelDirectory weights (:405-431; keys quantizer.layers.N.codebook.embed,
 *     decoder.layers.N.{conv,lstm.L.{Wx,Wh,bias},block.{1,3}.conv,shortcut.conv}.*, MLX layouts:
 *     Conv1d / ConvTranspose1d [out, k, in], LSTM [4H, in])                    -> b2a_encodec_create
 *   decode(_ audioCodes:_ audioScales:paddingMask:) / decodeAudio (:366-402,458-460) -> b2a_encodec_decode
 * audio_codes are [n_chunks, B, n_q, T] int32 (n_q <= the codebooks the checkpoint holds: the bandwidth chosen at encode
 * time), audio_scales [n_chunks, B] float32 or NULL (nil scales); the waveform is [B, samples, audio_channels] with
 * samples = b2a_encodec_output_length(n_chunks, T) (T*hop un-chunked; stride*(n_chunks-1) + T*hop with linearOverlapAdd).
 * The padding-mask truncation (:397-399) is a host-side slice of the result.  The reference's fatalError on
 * "Expected one frame" (:375-377) is B2A_ERR_AUDIO_DECODING_FAILED here.  Only norm_type "weight_norm" (plain folded
 * conv weights, no norm layer: EncodecLayers.swift:133-137) is implemented; "time_group_norm" -> invalidInput. */
typedef struct b2a_encodec_config {
    int32_t audio_channels;
    int32_t num_filters;
    int32_t kernel_size;
    int32_t num_residual_layers;
    int32_t dilation_growth_rate;
    int32_t codebook_size;
    int32_t codebook_dim;
    int32_t hidden_size;
    int32_t num_lstm_layers;
    int32_t residual_kernel_size;
    int32_t use_causal_conv;
    int32_t pad_mode_reflect;       /* 1 = "reflect" (clamped indices, EncodecLayers.swift:160-186), 0 = zero padding */
    int32_t norm_type;              /* 0 = weight_norm; anything else -> invalidInput */
    int32_t last_kernel_size;
    int32_t compress;
    int32_t n_upsampling_ratios;
    int32_t upsampling_ratios[8];
    int32_t sampling_rate;
    int32_t use_conv_shortcut;
    float trim_right_ratio;
    float chunk_length_s;           /* <= 0: nil (one frame) */
    float overlap;                  /* < 0: nil */
} b2a_encodec_config;

typedef struct b2a_encodec b2a_encodec;
int32_t b2a_encodec_create(int32_t device, const b2a_encodec_config* cfg, const b2a_tensor* tensors, int32_t n_tensors,
                           b2a_encodec** out);
int64_t b2a_encodec_output_length(const b2a_encodec* h, int32_t n_chunks, int32_t frames);
int32_t b2a_encodec_num_codebooks(const b2a_encodec* h);
void* b2a_encodec_stream(b2a_encodec* h);
int32_t b2a_encodec_decode(b2a_encodec* h, const int32_t* audio_codes, int32_t n_chunks, int32_t batch, int32_t n_q,
                           int32_t frames, const float* audio_scales, float* wave);
int32_t b2a_encodec_decode_dev(b2a_encodec* h, const int32_t* d_audio_codes, int32_t n_chunks, int32_t batch, int32_t n_q,
                               int32_t frames, const float* d_audio_scales, float* d_wave, void* stream);
void b2a_encodec_destroy(b2a_encodec* h);

/* ------------------------------------------------------------------ Whisper STT
 * Replaces class WhisperModel (Sources/MLXAudioSTT/Models/Whisper/WhisperModel.swift:7-309) behind
 * STTGenerationModel (Sources/MLXAudioSTT/Generation.swift:52-64):
 *   fromDirectory / sanitize (:321-382)           -> b2a_stt_create (config + HF-named tensors: conv weights in the
 *                                                    PyTorch [out,in,k] layout, matrices bf16 or fp32 (rounded to bf16))
 *   model.encoder(features) (WhisperLayers.swift:146-155) -> b2a_stt_encode          (parity hook)
 *   model.decoder(tokens:...) + projectToVocab     -> b2a_stt_decoder_logits          (parity hook, after an encode)
 *   generate(audio:) / transcribeChunk (:36,186-282) -> b2a_stt_transcribe: B independent <=30 s clips
 *     ("batched == serial"; the reference transcribes one chunk at a time), greedy decode with the reference's
 *     suppress masks; returns token ids (detokenisation stays with the host tokenizer).  The 30 s chunking of
 *     longer audio (:165-182) is the caller's loop.
 * pcm is [B, n_samples] float32 16 kHz mono, every clip padded / trimmed to 30 s like WhisperAudio.padOrTrimToWindow. */
typedef struct b2a_whisper_config {
    int32_t vocab_size;
    int32_t num_mel_bins;
    int32_t d_model;
    int32_t encoder_layers;
    int32_t encoder_attention_heads;
    int32_t encoder_ffn_dim;
    int32_t max_source_positions;
    int32_t decoder_layers;
    int32_t decoder_attention_heads;
    int32_t decoder_ffn_dim;
    int32_t max_target_positions;
    int32_t max_batch; /* <= 16 clips per call */
} b2a_whisper_config;

/* STTGenerateParameters as used by transcribeChunk + WhisperGenerationConfig's suppress lists */
typedef struct b2a_stt_params {
    int32_t max_tokens;           /* defaultGenerationParameters: max_target_positions - 16 */
    float temperature;            /* 0: greedy argmax (lowest index wins ties); > 0: categorical(logits / T), WhisperModel.swift:284-291 */
    const int32_t* prompt_ids;    /* decoder prefix from buildPromptTokens (WhisperTokenizer.swift:98-113) */
    int32_t n_prompt;
    const int32_t* begin_suppress; /* suppressed at step 0 only (default [endOfText]) */
    int32_t n_begin_suppress;
    const int32_t* suppress;      /* suppressed at every step */
    int32_t n_suppress;
    int32_t timestamp_begin;      /* ids >= this are always suppressed (WhisperModel.swift:236) */
    int32_t eot;                  /* end-of-text id: stops a clip */
    uint64_t seed;                /* temperature > 0: the draw of (clip b, step s) is a pure function of (seed, b, s) */
} b2a_stt_params;

typedef struct b2a_stt_info {
    int32_t prompt_tokens;
    int32_t generation_tokens;
    int32_t decode_steps;
    double encode_time; /* log-mel + encoder + cross K/V */
    double decode_time;
    double total_time;
} b2a_stt_info;

typedef struct b2a_stt b2a_stt;
int32_t b2a_stt_create(int32_t device, const b2a_whisper_config* cfg, const b2a_tensor* tensors,
                       int32_t n_tensors, b2a_stt** out);
void* b2a_stt_stream(b2a_stt* h);
/* enc_out [B, 1500, d_model] float32 (host) */
int32_t b2a_stt_encode(b2a_stt* h, const float* pcm, int32_t batch, int64_t n_samples, float* enc_out);
/* tokens [B, T] teacher-forced from position 0 against the last encode; logits_out [B, T, vocab] (host) */
int32_t b2a_stt_decoder_logits(b2a_stt* h, const int32_t* tokens, int32_t batch, int32_t len, float* logits_out);
/* tokens_out [B, params->max_tokens], n_tokens_out [B] (host) */
int32_t b2a_stt_transcribe(b2a_stt* h, const float* pcm, int32_t batch, int64_t n_samples, const b2a_stt_params* params,
                           int32_t* tokens_out, int32_t* n_tokens_out, b2a_stt_info* info);
int32_t b2a_stt_transcribe_dev(b2a_stt* h, const float* d_pcm, int32_t batch, int64_t n_samples,
                               const b2a_stt_params* params, int32_t* tokens_out, int32_t* n_tokens_out, b2a_stt_info* info);
/* generate(audio:) beyond one window (WhisperModel.swift:95-182, chunkAudioFor30sWindows :165-182): consecutive 30 s windows of one
 * mono 16 kHz signal, transcribed as a BATCH (the reference loops over them); tokens_out [n_chunks, params->max_tokens],
 * n_tokens_out / offsets_s [n_chunks] (offsets_s nullable), *n_chunks_out = ceil(n / 480000) <= max_chunks. */
int32_t b2a_stt_transcribe_long(b2a_stt* h, const float* pcm, int64_t n_samples, const b2a_stt_params* params, int32_t max_chunks,
                                int32_t* tokens_out, int32_t* n_tokens_out, float* offsets_s, int32_t* n_chunks_out,
                                b2a_stt_info* info);
int32_t b2a_stt_cancel(b2a_stt* h);
void b2a_stt_destroy(b2a_stt* h);

/* ---- the streaming STT session around the model ------------------------------------------------------
 * Sources/MLXAudioSTT/Streaming/StreamingInferenceSession.swift:589-950 (the core that drives `any STTGenerationModel` through
 * streamingDecodeTokenIds(audio:config:confirmedTokenIds:)) and StreamingTypes.swift:36-92 (StreamingConfig, DelayPreset), at the token
 * level: the host keeps the tokenizer, the text de-duplication of the window overlap and the AsyncStream of TranscriptionEvents; it calls
 * feed() from feedAudio(samples:) and stop() from stop(), passing its own clock (Date().timeIntervalSinceReferenceDate).
 *   feed : samples join the pending buffer.  A whole window (window_s) pending -> it is frozen (the buffer keeps its last
 *          window_overlap_s), decoded once without a prefix, and its tokens become completed window n (kind 2); confirmed / provisional
 *          tokens are cleared.  Otherwise, with >= 0.5 s pending and max(0.2, decode_interval_s) since the last pass, the whole pending
 *          buffer is decoded with the confirmed tokens as a forced decoder prefix (kind 1) and promoteTokens runs: a provisional
 *          position keeps its first-seen time and gains an agreement while it repeats the previous pass; the longest prefix older than
 *          delay_ms AND agreed on by min_agreement_passes passes moves to the confirmed list.  At most one decode pass per call.
 *   stop : what is pending is decoded as a last window; left-over provisional tokens are confirmed (kind 3).
 * Decode passes run synchronously on the model's stream (the reference detaches a Task and drops feeds that arrive while one runs).  */
typedef struct b2a_stt_stream_config {
    double decode_interval_s;     /* StreamingConfig.decodeIntervalSeconds, 1.0 */
    double window_s;              /* 8.0: the reference freezes 8 s windows */
    double window_overlap_s;      /* encoderWindowOverlapSeconds, 1.0 */
    int32_t delay_ms;             /* DelayPreset.delayMs: realtime 200, agent 480 (default), subtitle 2400 */
    int32_t min_agreement_passes; /* minAgreementPasses, 2 */
    int32_t max_tokens_per_pass;  /* maxTokensPerPass, 512 (clamped to what the decoder context leaves after the prefix) */
    int32_t sample_rate;          /* 16000 */
} b2a_stt_stream_config;

typedef struct b2a_stt_stream_update {
    int32_t kind;                 /* 0 nothing ran, 1 partial pass, 2 window finalised, 3 ended */
    int32_t promoted;             /* tokens moved provisional -> confirmed by this pass */
    int32_t completed_windows;    /* finalised windows so far (their tokens: b2a_stt_session_tokens(which = 0, window)) */
    int32_t n_confirmed;          /* confirmed tokens of the current (pending) window */
    int32_t n_provisional;
    double total_audio_s;         /* StreamingStats.totalAudioSeconds */
    double pass_encode_time;      /* of the pass this call ran (0 when none) */
    double pass_decode_time;
} b2a_stt_stream_update;

/* A host-side decoder (the reference accepts `any STTGenerationModel`, :162): writes the continuation AFTER `prefix` for `pcm`. */
typedef int32_t (*b2a_stt_decode_cb)(void* user, const float* pcm, int64_t n_samples, const int32_t* prefix, int32_t n_prefix,
                                     int32_t* tokens_out, int32_t capacity, int32_t* n_tokens_out);

typedef struct b2a_stt_session b2a_stt_session;
/* params: the decode parameters of every pass (prompt / suppress lists are copied; max_tokens is replaced by max_tokens_per_pass) */
int32_t b2a_stt_session_create(b2a_stt* model, const b2a_stt_params* params, const b2a_stt_stream_config* config, b2a_stt_session** out);
int32_t b2a_stt_session_create_with_decoder(b2a_stt_decode_cb decode, void* user, const b2a_stt_stream_config* config, b2a_stt_session** out);
int32_t b2a_stt_session_feed(b2a_stt_session* s, const float* pcm, int64_t n_samples, double now_s, b2a_stt_stream_update* update);
int32_t b2a_stt_session_stop(b2a_stt_session* s, double now_s, b2a_stt_stream_update* update);
/* which: 0 = completed window `window`, 1 = confirmed, 2 = provisional; *n_out = the list's length (written even when capacity is short) */
int32_t b2a_stt_session_tokens(b2a_stt_session* s, int32_t which, int32_t window, int32_t* tokens_out, int32_t capacity, int32_t* n_out);
void b2a_stt_session_destroy(b2a_stt_session* s);

/* ------------------------------------------------------------------ Qwen3-TTS speech-tokenizer decoder
 * Replaces Qwen3TTSSpeechTokenizerDecoder and the decode entry points of Qwen3TTSSpeechTokenizer
 * (Sources/MLXAudioTTS/Models/Qwen3TTS/Qwen3TTSSpeechTokenizer.swift):
 *   init(config:) + sanitized weights (keys below the "decoder." module, MLX layouts)  -> create
 *   resetStreamingState (:949-970)                                                       -> reset
 *   streamingStep (:973-1008)                                                            -> streaming_step[_dev]
 *   streamingDecode(chunkTokens:) (:1070-1092, what decodeChunk uses, Qwen3TTS.swift:214-231) -> streaming_decode
 *   chunkedDecode(chunkSize:leftContextSize:) (:1010-1024, used by decode :1059-1068)   -> chunked_decode
 * codes are [B, num_quantizers given, T] int32 (the decoder's own layout, i.e. audioCodes transposed (0, 2, 1)); the
 * waveform is [B, T * total_upsample] float32 in [-1, 1].  Config defaults: Qwen3TTSConfig.swift:358-385.            */
typedef struct b2a_speech_tokenizer_config {
    int32_t codebook_size;
    int32_t codebook_dim;
    int32_t latent_dim;
    int32_t decoder_dim;
    int32_t hidden_size;
    int32_t intermediate_size;
    int32_t head_dim;
    int32_t num_attention_heads;
    int32_t num_key_value_heads;
    int32_t num_hidden_layers;
    int32_t num_quantizers;
    int32_t num_semantic_quantizers;
    float rms_norm_eps;
    float rope_theta;
    int32_t attention_bias;
    int32_t num_upsample_rates;
    int32_t upsample_rates[8];
    int32_t num_upsampling_ratios;
    int32_t upsampling_ratios[8];
    int32_t max_batch;        /* rows decoded together */
    int32_t max_cache_frames; /* code frames one stream may span (the reference's cache is unbounded) */
} b2a_speech_tokenizer_config;

typedef struct b2a_speech_tokenizer b2a_speech_tokenizer;
int32_t b2a_speech_tokenizer_create(int32_t device, const b2a_speech_tokenizer_config* cfg, const b2a_tensor* tensors,
                                    int32_t n_tensors, b2a_speech_tokenizer** out);
int32_t b2a_speech_tokenizer_total_upsample(const b2a_speech_tokenizer* h);
void* b2a_speech_tokenizer_stream(b2a_speech_tokenizer* h);
int32_t b2a_speech_tokenizer_reset(b2a_speech_tokenizer* h);
int32_t b2a_speech_tokenizer_streaming_step(b2a_speech_tokenizer* h, const int32_t* codes, int32_t batch, int32_t num_groups,
                                            int32_t frames, float* wave);
int32_t b2a_speech_tokenizer_streaming_step_dev(b2a_speech_tokenizer* h, const int32_t* d_codes, int32_t batch, int32_t num_groups,
                                                int32_t frames, float* d_wave, void* stream);
int32_t b2a_speech_tokenizer_streaming_decode(b2a_speech_tokenizer* h, const int32_t* codes, int32_t batch, int32_t num_groups,
                                              int32_t frames, int32_t chunk_tokens, float* wave);
int32_t b2a_speech_tokenizer_chunked_decode(b2a_speech_tokenizer* h, const int32_t* codes, int32_t batch, int32_t num_groups,
                                            int32_t frames, int32_t chunk_size, int32_t left_context, float* wave);
void b2a_speech_tokenizer_destroy(b2a_speech_tokenizer* h);
/* Loading (host-only except the final create): the decoder half of Qwen3TTSSpeechTokenizer.sanitize (:1094-1440) on an open
 * checkpoint -- prefixes stripped, PyTorch conv / transposed-conv layouts moved to MLX's with the reference's shape heuristic
 * (checkArrayShapeQwen3 :1445-1455), upsample.X.Y -> upsample.X.layers.Y, codebook statistics kept, encoder.* and speaker-encoder
 * keys dropped, the leading "decoder." removed; Qwen3TTSTokenizerConfig decoding (Qwen3TTSConfig.swift:358-385,518-527; a missing
 * config.json means defaults, Qwen3TTS.swift:1246-1255); loadSpeechTokenizer (Qwen3TTS.swift:1244-1275).                       */
int32_t b2a_weights_sanitize_speech_tokenizer(b2a_weights* w);
int32_t b2a_speech_tokenizer_config_from_json(const char* config_path, int32_t max_batch, int32_t max_cache_frames,
                                              b2a_speech_tokenizer_config* cfg, int32_t* decode_upsample_rate);
int32_t b2a_speech_tokenizer_create_from_directory(const char* dir, int32_t device, int32_t max_batch, int32_t max_cache_frames,
                                                   b2a_speech_tokenizer** out, int32_t* decode_upsample_rate);
/* ------------------------------------------------------------------ Qwen3-TTS talker + code predictor
 * Replaces the autoregressive half of class Qwen3TTSModel (Sources/MLXAudioTTS/Models/Qwen3TTS/):
 *   Qwen3TTSTalkerForConditionalGeneration (Qwen3TTSTalker.swift:127-366: per-head q/k RMSNorm before RoPE, interleaved 3-section
 *     MRoPE -- all three position rows are equal for the text-only prompts the reference builds, where it is plain rotate-half
 *     RoPE --, inputs are EMBEDDINGS, codec_head)                                               -> the 28-layer stack
 *   Qwen3TTSCodePredictor (Qwen3TTSCodePredictor.swift:14-243: 5 layers, 15 lm heads / embeddings, cache reset every frame) -> the
 *     inner 15-step loop
 *   the frame loop of generate (Qwen3TTS.swift:380-495: talker step -> sampleToken -> 15 predictor steps -> summed-embedding
 *     feedback `text + codec_embed(c0) + sum_i predictor_embed_i(c_i+1)`) and sampleToken (:1003-1118) -> b2a_qwen3_talker_generate:
 *     ONE CUDA graph per 12.5 Hz frame, nothing syncs with the host inside it
 *   text_projection(text_embedding(ids)) / codec_embedding(ids) (:898-999, prepareGenerationInputs) -> b2a_qwen3_talker_embed_text /
 *     _embed_codec: the host composes the prompt from these rows exactly as prepareGenerationInputs does (tokenisation, the chat
 *     template and the special-token ids stay with the host)
 * Weights: the reference's keys after sanitize strips "talker." (model.layers.N.*, model.codec_embedding.weight,
 * model.text_embedding.weight, text_projection.linear_fc{1,2}.{weight,bias}, codec_head.weight, code_predictor.model.layers.N.*,
 * code_predictor.model.codec_embedding.I.weight, code_predictor.lm_head.I.weight), bf16 or f32 (rounded to bf16: the engine holds
 * bf16 matrices; an MLX affine-quantised (8-bit) checkpoint is expanded by b2a_weights_dequantize first).  head_dim must be 128
 * and the predictor's hidden size must equal the talker's (no small_to_mtp_projection) -- true of the shipped 0.6B geometry.   */
typedef struct b2a_qwen3_talker_config {
    int32_t vocab_size;            /* codec vocabulary (3072) */
    int32_t hidden_size;
    int32_t intermediate_size;
    int32_t num_hidden_layers;
    int32_t num_attention_heads;
    int32_t num_key_value_heads;
    int32_t head_dim;
    float rms_norm_eps;
    float rope_theta;
    int32_t num_code_groups;       /* 16: one talker code + 15 predictor codes per frame */
    int32_t text_hidden_size;
    int32_t text_vocab_size;
    int32_t codec_eos_token_id;
    int32_t cp_vocab_size;         /* code predictor (Qwen3TTSConfig.swift:45-63) */
    int32_t cp_hidden_size;
    int32_t cp_intermediate_size;
    int32_t cp_num_hidden_layers;
    int32_t cp_num_attention_heads;
    int32_t cp_num_key_value_heads;
    int32_t cp_head_dim;
    float cp_rms_norm_eps;
    float cp_rope_theta;
    int32_t max_batch;             /* <= 8 utterances per call */
    int32_t max_context;           /* prompt embeddings + frames per utterance */
} b2a_qwen3_talker_config;

/* Qwen3TTS generate's sampling parameters (Qwen3TTS.swift:360-385; sampleToken :1003-1118) */
typedef struct b2a_qwen3_gen_params {
    int32_t max_tokens;            /* frames; the caller applies min(maxTokens, max(75, 6 * text tokens)) (:380) */
    float temperature;             /* <= 0: greedy argmax (lowest index wins ties) for the talker code AND the predictor codes */
    float top_p;
    int32_t top_k;
    float min_p;
    float repetition_penalty;      /* talker code only, over the unique codes generated so far */
    uint64_t seed;
} b2a_qwen3_gen_params;

typedef struct b2a_qwen3_talker b2a_qwen3_talker;
int32_t b2a_qwen3_talker_create(int32_t device, const b2a_qwen3_talker_config* cfg, const b2a_tensor* tensors, int32_t n_tensors,
                                b2a_qwen3_talker** out);
void* b2a_qwen3_talker_stream(b2a_qwen3_talker* h);
/* out [n, hidden] float32 (host): text_projection(text_embedding(ids)) resp. codec_embedding(ids) */
int32_t b2a_qwen3_talker_embed_text(b2a_qwen3_talker* h, const int32_t* ids, int32_t n, float* out);
int32_t b2a_qwen3_talker_embed_codec(b2a_qwen3_talker* h, const int32_t* ids, int32_t n, float* out);
/* Parity hook: the talker over input_embeds [B, L, hidden] from an empty cache -> codec logits of the LAST position
 * [B, vocab] and its final-norm hidden state [B, hidden] (Qwen3TTSTalker.swift:340-350).                                   */
int32_t b2a_qwen3_talker_forward(b2a_qwen3_talker* h, const float* input_embeds, int32_t batch, int32_t len, float* logits_out,
                                 float* hidden_out);
/* The frame loop.  input_embeds [B, L, hidden] (every row the same L), trailing_text_hidden [B, n_trailing_max, hidden] with
 * n_trailing[B] valid rows each (one is consumed per frame, then tts_pad_embed [hidden] is used), codes_out [B, max_tokens,
 * num_code_groups] int32, n_frames_out[B].  A row stops after the frame whose talker code is codec_eos_token_id (that frame is not
 * emitted, :424-428) or at max_tokens.  on_frame (nullable) is called on the calling thread for every emitted frame with the
 * frame's num_code_groups codes -- the hook a streaming caller decodes audio chunks from (generateStream, Qwen3TTS.swift:500-569). */
typedef void (*b2a_frame_cb)(void* user, int32_t utterance, int32_t frame, const int32_t* codes);
int32_t b2a_qwen3_talker_generate(b2a_qwen3_talker* h, const float* input_embeds, int32_t batch, int32_t len,
                                  const float* trailing_text_hidden, const int32_t* n_trailing, int32_t n_trailing_max,
                                  const float* tts_pad_embed, const b2a_qwen3_gen_params* params, int32_t* codes_out,
                                  int32_t* n_frames_out, b2a_gen_info* info, b2a_frame_cb on_frame, void* user);
/* Loading (Qwen3TTSModel.fromModelDirectory, Qwen3TTS.swift:1136-1175, talker half): config.json's "talker_config" (+ nested
 * "code_predictor_config", defaults of Qwen3TTSConfig.swift:45-63,268-292) -> the config struct; every *.safetensors of the directory
 * -> keep "talker.*" and strip the prefix (Qwen3TTSTalker.swift:356-365) -> MLX affine de-quantisation of every layer that carries
 * ".scales" as config.json's "quantization" block says (:1156-1171; 8-bit group-64 for the shipped 8-bit checkpoints) -> create.     */
int32_t b2a_qwen3_talker_config_from_json(const char* config_path, int32_t max_batch, int32_t max_context, b2a_qwen3_talker_config* cfg);
int32_t b2a_weights_sanitize_qwen3_talker(b2a_weights* w, const char* config_path /* nullable: no quantisation */);
int32_t b2a_qwen3_talker_create_from_directory(const char* model_dir, int32_t device, int32_t max_batch, int32_t max_context,
                                               b2a_qwen3_talker** out);
int32_t b2a_qwen3_talker_cancel(b2a_qwen3_talker* h);
void b2a_qwen3_talker_destroy(b2a_qwen3_talker* h);

#ifdef __cplusplus
}
#endif
#endif /* B200AUDIO_H */
