/* infinisst_b200 C-ABI: the drop-in boundary of the B200-native per-chunk streaming step.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes (no torch types) and returns an
 * int status: 0 = ok, non-zero = error with the message available from isst_last_error().  There is
 * no CPU fallback: without a CUDA device isst_create fails.
 *
 * Each function names the reference interface (file:line under /root/reference) it replaces; the
 * Python shims in infinisst_b200/model.py and infinisst_b200/agent.py rebuild the reference's own
 * signatures on top (INTEGRATION.md shows the binding a reference maintainer would add).
 */
#ifndef INFINISST_B200_H_
#define INFINISST_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ISST_MAX_CONV 8
#define ISST_DTYPE_F32 0
#define ISST_DTYPE_BF16 1

typedef struct isst_ctx isst_ctx;

/* Model + capacity description.  Field names follow agents/options.py:1-41 (block_size,
 * max_cache_size), scripts/infer/infinisst.sh:68-71 and the wav2vec2 / Llama hyper-parameters
 * (SURVEY App. A.1/A.3). */
typedef struct isst_config {
  /* wav2vec2 conv feature extractor, layer_norm mode: (dim, kernel, stride) per block */
  int n_conv;
  int conv_dim[ISST_MAX_CONV], conv_k[ISST_MAX_CONV], conv_s[ISST_MAX_CONV];
  /* streaming transformer encoder */
  int enc_dim, enc_ffn, enc_heads, enc_layers;
  int block_size;      /* frames per block at multiplier 1 (--block-size) */
  int max_cache_size;  /* encoder KV window in frames (--max-cache-size) */
  /* length adapter (--length-shrink-cfg): bias-free Conv1d(k, s) + LayerNorm + GELU blocks */
  int n_adapter;
  int adapter_dim[ISST_MAX_CONV], adapter_k[ISST_MAX_CONV], adapter_s[ISST_MAX_CONV];
  /* Llama */
  int hidden, layers, heads, kv_heads, head_dim, ffn, vocab;
  float rms_eps;
  /* capacities */
  int max_streams;     /* concurrently open streams (state slots) */
  int max_batch;       /* streams advanced by one call */
  int max_multiplier;  /* largest latency multiplier (chunk = 48*m frames) */
  int kv_pages;        /* LLM KV pool size in 16-token pages (shared by all streams) */
  int max_kv_len;      /* largest logical KV length of one stream (RoPE table / page table size) */
  int max_prompt;      /* longest turn prompt in tokens */
  int max_new_tokens;  /* longest generation per call */
  /* encoder position variants (agents/options.py:32-41; zero = production, scripts/infer/infinisst.sh:70-71):
   *   enc_xpos    --xpos 1: xPos scaling of rotated q / k, RotaryEmbedding(use_xpos) at patch_speech_encoder.py:631,823-824
   *   enc_no_rope --rope 0: no rotation; the bf16 sinusoidal table of patch_speech_encoder.py:448-461 is added to
   *               the frames at their absolute index (:488-493) */
  int enc_xpos;
  int enc_no_rope;
} isst_config;

/* Greedy generation parameters: the kwargs of model.generate at agents/infinisst.py:307-332 that
 * affect the greedy path. */
typedef struct isst_gen_params {
  int max_new_tokens;
  int no_repeat_ngram_size;       /* also used as encoder_no_repeat_ngram_size (:319-320) */
  float repetition_penalty;
  int n_eos;
  int eos_token_ids[8];
  int n_suppress;
  const int32_t* suppress_tokens; /* host */
  int pin_prefix;                 /* system_prompt_size when --always-cache-system-prompt, else 0;
                                     only read on the first prefill of a stream */
} isst_gen_params;

const char* isst_last_error(void);

/* Lifetime.  Replaces model construction at agents/infinisst.py:150-181. */
int isst_create(const isst_config* cfg, int device, isst_ctx** out);
void isst_destroy(isst_ctx* ctx);

/* Weight ingestion, keyed by the reference's state-dict names (`pytorch_model.bin`,
 * agents/infinisst.py:179-180; key layout SURVEY §8b).  `data` may be a host or a device pointer.
 * Extra keys: "rope.enc.inv_freq" [enc_head_dim/2] f32 (rotary_embedding_torch `freqs`,
 * patch_speech_encoder.py:631,823-824) and "rope.llm.inv_freq" [head_dim/2] f32
 * (HF LlamaRotaryEmbedding, patch_llm.py:287-299).  Unknown keys (e.g. encoder.pos_conv.*, mask_emb)
 * are accepted and ignored, like load_state_dict on unused modules. */
int isst_load_weight(isst_ctx* ctx, const char* name, const void* data, const int64_t* shape, int ndim,
                     int dtype);
int isst_finalize_weights(isst_ctx* ctx);

/* Per-stream state = S2TAgentStates.speech_cache + .past_key_values (agents/infinisst.py:50-67),
 * W2V2RoPECache (model/speech_encoder.py:80-97).  Opaque slot ids. */
int isst_stream_open(isst_ctx* ctx, int* stream_id);
int isst_stream_close(isst_ctx* ctx, int stream_id);

/* SpeechEncoderW2V2RoPE.encode_speech (model/speech_encoder.py:219-236) for n streams in lock-step:
 * pcm [n][n_samples] f32 (host or device) -> speech features [n][n_samples/320/4][hidden] bf16, kept
 * in the context for the following isst_generate and optionally copied to out_feats (device or host
 * pointer, may be NULL).  A fresh stream takes 79+320 leading samples more (agents/infinisst.py:216-218).
 * n_samples = j * block_size * 320 with 1 <= j <= multiplier: a short final chunk is padded to whole
 * segments of ONE block only (agents/infinisst.py:211-213) while the encoder mask keeps the block size of
 * the multiplier (set_blocksize, model/speech_encoder.py:143-145); j segments give 12 * j speech rows. */
int isst_encode_chunk(isst_ctx* ctx, int n, const int* stream_ids, const float* pcm, int n_samples,
                      int multiplier, void* out_feats, void* cuda_stream);

/* model.generate(num_beams=1, do_sample=False) as called at agents/infinisst.py:307-332: prefill the
 * turn prompt (speech features spliced into the <sp_patch> slots, model/llm.py:86-113) on the
 * stream's KV cache, then greedy decode with the HF logits processors.  Token ids and lengths are
 * host arrays: ids packed [sum(lens)], speech_slot packed (index into the stream's speech features
 * or -1), enc_ids packed [sum(enc_lens)] = last `lookback` target ids (:298-301).  forced (optional,
 * [n][max_new_tokens]) teacher-forces the chosen tokens (parity tests).  Outputs: out_tokens
 * [n][max_new_tokens] (all chosen tokens including the one that is never forwarded),
 * out_counts [n]. */
int isst_generate(isst_ctx* ctx, int n, const int* stream_ids, const int32_t* ids, const int* lens,
                  const int32_t* speech_slot, const int32_t* enc_ids, const int* enc_lens,
                  const isst_gen_params* gen, const int32_t* forced, int32_t* out_tokens, int* out_counts,
                  void* cuda_stream);

/* Beam search with KV hand-back: the reference's shipped decoding (scripts/infer/infinisst.sh:48 `--beam 4`).
 * Replaces model.generate(num_beams=k, ...) as agents/infinisst.py:307-336 calls it: patch_hf.py:626-655 (dispatch),
 * :305-342 (k-fold input expansion - here the prompt is prefilled ONCE per stream and the beams share its KV pages),
 * :687-967 (loop: log-softmax, logits processors on the log-probs, top max(2, 1 + n_eos) * k candidates, KV reorder),
 * :43-157 / :278-302 (scorer: EOS candidates close hypotheses, `is_done` bound), :159-275 (finalize: best hypothesis
 * and ITS KV cache).  out_tokens [n][max_new_tokens + 1]: generated part of `sequences` (the closing EOS is appended
 * when it fits, as :262-264 does), -1 padded; out_counts [n]; out_scores [n] (may be null) = sequence score.
 * After the call each stream's KV holds its prompt and the forwarded tokens of the winning hypothesis.
 * n * num_beams must not exceed max_batch; the page pool needs 3 * num_beams spare tail page sets
 * per stream of the call.  `follow` (may be null) teacher-forces the discrete choices for tolerance-aware parity
 * tests; `trace` (may be null) returns every step's candidates and choices.  All pointers are HOST pointers. */
typedef struct isst_beam_follow {
  const int32_t* closed;   /* [n][max_new][k][2] (parent beam, eos token) closed at each step, -1 terminated per step */
  const int32_t* next;     /* [n][max_new][k][2] (parent beam, token) continued at each step */
  const int32_t* steps;    /* [n] selection steps to run */
  const int32_t* done;     /* [n] 1: the scorer declared the sentence done at the last of those steps */
} isst_beam_follow;
typedef struct isst_beam_trace {
  float* cand_scores;      /* [n][max_new][n_keep] best first */
  int32_t* cand_index;     /* [n][max_new][n_keep] parent beam * vocab + token */
  int32_t* next;           /* [n][max_new][k][2] (parent beam, token) */
  float* next_scores;      /* [n][max_new][k] */
  int32_t* steps;          /* [n] selection steps run */
} isst_beam_trace;
int isst_generate_beam(isst_ctx* ctx, int n, const int* stream_ids, const int32_t* ids, const int* lens,
                       const int32_t* speech_slot, const int32_t* enc_ids, const int* enc_lens,
                       const isst_gen_params* gen, int num_beams, float length_penalty,
                       const isst_beam_follow* follow, int32_t* out_tokens, int* out_counts, float* out_scores,
                       isst_beam_trace* trace, void* cuda_stream);

/* model.forward (model/llm.py:192-270) on the stream caches: append T_b tokens per stream and return
 * the logits of every stream's last position ([n][vocab] f32, device or host pointer).  embeds_override
 * (optional device pointer, bf16 [sum(lens)][hidden]) replaces the embedding/splice step. */
int isst_forward(isst_ctx* ctx, int n, const int* stream_ids, const int32_t* ids, const int* lens,
                 const int32_t* speech_slot, const void* embeds_override, int pin_prefix, float* out_logits,
                 void* cuda_stream);

/* The same forward with the reference's own output shape: logits of EVERY position (model/llm.py:236-237 applies
 * lm_head to all T rows; CausalLMOutputWithPast.logits is [B, T, vocab]), packed [sum(lens)][vocab] f32 in stream
 * order.  For scoring / teacher-forced evaluation; generation only ever reads the last row (isst_forward). */
int isst_forward_all(isst_ctx* ctx, int n, const int* stream_ids, const int32_t* ids, const int* lens,
                     const int32_t* speech_slot, const void* embeds_override, int pin_prefix, float* out_logits,
                     void* cuda_stream);

/* KV length (past_key_values[0][0].size(2), agents/infinisst.py:337) and sliding-window eviction
 * (agents/infinisst.py:354-361): drop the logical tokens [keep_prefix, drop_upto).  Page-table edit,
 * no data movement.  keep_prefix must be 0 or the stream's pinned prefix. */
int isst_kv_len(isst_ctx* ctx, int stream_id, int* len);
int isst_kv_evict(isst_ctx* ctx, int stream_id, int keep_prefix, int drop_upto);
int isst_enc_steps(isst_ctx* ctx, int stream_id, int* n_steps);   /* W2V2RoPECache.n_steps */

/* Introspection for tests / bench. */
/* bit 0: keep taps - encoder stages, "step_logits" (raw last-position logits per step, f32 [max_new][n][vocab]),
 * "step_scores" (the same rows after the logits processors = what the device's arg-max saw) and "step_picked"
 * (int32 [n][max_new]: the device's own arg-max at every step, also when `forced` overrides the token that is fed). */
int isst_debug_enable(isst_ctx* ctx, int on);
int isst_debug_read(isst_ctx* ctx, const char* name, void* dst_host, int64_t max_bytes, int64_t* n_bytes);
/* Test hook: pretend `delta` more ring tokens were appended and evicted before now, i.e. move the stream's
 * absolute RoPE positions (a one-hour stream reaches ~1e5).  Scores only depend on position differences
 * (patch_llm.py:287-299), so results must not change. */
int isst_debug_shift_positions(isst_ctx* ctx, int stream_id, int64_t delta);
int64_t isst_launch_count(isst_ctx* ctx);       /* kernels launched so far */
/* Launches so far of one kernel variant, by name ("prefill_attention_tc_unsplit", "decode_attention_direct",
 * "gemm_sk_swap64_deferred", "decode_chain64", ...): parity tests assert that the variant they mean to cover ran. */
int isst_path_count(isst_ctx* ctx, const char* name, int64_t* count);
/* Per-context test / tuning options (no environment variables on the hot path): "pdl" 0/1 programmatic dependent
 * launch, "decode_splits" fixed key-split count of decode attention (0 = automatic), "decode_chain" 0/1 the fused
 * layer kernel (0 = one kernel per operator), "chain_fold" 0/1 RMSNorms folded into the decode GEMMs (default 0: the
 * reference's rounding points), "gemm_pair" 0/1 CTA-pair GEMM above 128 token rows (0 = one CTA per tile),
 * "prefill_l2_ahead" K/V tiles the prefill attention asks into L2 ahead of its ring, "defer_splits_as_chain" and
 * "tap_llm_layers" (parity tests: operator path with the chain's k-ranges; per-layer residual taps). */
int isst_debug_option(isst_ctx* ctx, const char* key, int value);
/* Per-kernel-class device timing for the roofline leg of bench.py: while enabled every launch is
 * bracketed by CUDA events on the caller's stream.  isst_profile_read walks the classes by index
 * (returns 1 past the last one): launches, summed event time, and the ALGORITHMIC flops / bytes
 * of those launches (DESIGN.md states the per-unit figures; the reference's own probe is the
 * whole-step synchronized_timer, agents/infinisst.py:37-48). */
int isst_profile_enable(isst_ctx* ctx, int on);
int isst_profile_reset(isst_ctx* ctx);
int isst_profile_read(isst_ctx* ctx, int index, char* name, int name_cap, int64_t* launches, double* ms,
                      double* flops, double* bytes);
int isst_pages_free(isst_ctx* ctx);

/* Stand-alone operator entry points (device pointers) used by the parity tests and bench.py:
 * out[M,N] = act[M,K] . w[N,K]^T (+bias, gelu, +resid) through the tcgen05 GEMM. */
int isst_op_gemm(isst_ctx* ctx, const void* act_bf16, const void* w_bf16, int M, int N, int K,
                 const float* bias, int act_gelu, const void* resid_bf16, int dual, void* out, int out_f32,
                 int force_swap, int force_splits, void* cuda_stream);
/* Decode attention over synthetic paged KV for `n` streams of length L (bench roofline leg). */
int isst_op_decode_attention_bench(isst_ctx* ctx, int n, int L, int iters, float* ms_per_iter,
                                   void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* INFINISST_B200_H_ */
