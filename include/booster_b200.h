/* include/booster_b200.h — ADDITIVE token-level and operator-level C-ABI of libbooster_b200.so.
 *
 * include/bridge.h is the boundary Booster's Go server binds. The entry points below sit one seam
 * lower — the llama.h calls cpp/bridge.cpp makes on the hot path — so that the CUDA path can be
 * parity-tested against the reference's CPU path on token ids and logits, without tokenizer or
 * sampler in between (SURVEY.md §8b "inner seams"). Every function cites the reference interface it
 * replaces. Plain C, host pointers and sizes only; all device memory is owned by the library.
 *
 * Error convention: functions returning int return 0 on success, non-zero on failure;
 * b200_last_error() returns a thread-local message. There is NO CPU fallback anywhere: without a CUDA
 * device every compute entry point fails loudly.
 */
#ifndef BOOSTER_B200_H
#define BOOSTER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b200_model b200_model;   /* replaces struct llama_model   (cpp/include/llama.h) */
typedef struct b200_ctx   b200_ctx;     /* replaces struct llama_context (cpp/include/llama.h) */

/* ggml tensor type ids of the block formats on the path (cpp/ggml/include/ggml.h:360-375) */
enum b200_type {
    B200_TYPE_F32  = 0,
    B200_TYPE_F16  = 1,
    B200_TYPE_Q8_0 = 8,
    B200_TYPE_Q4_K = 12,
    B200_TYPE_Q5_K = 13,
    B200_TYPE_Q6_K = 14,
};

const char * b200_last_error(void);
int          b200_device_count(void);                 /* 0 when no CUDA device is visible */
const char * b200_version(void);

/* ---- model + context -------------------------------------------------------------------------------------
 * b200_model_load     replaces llama_load_model_from_file (cpp/bridge.cpp:131; loader cpp/src/llama.cpp:16539,
 *                     tensors :6062-6110). Parses the GGUF, keeps every matrix in its block format
 *                     (re-tiled into 16-byte-aligned planes, same bytes) in HBM of CUDA device `device`.
 *                     A pipeline stage holds layers [layer_begin, layer_end); pass 0,-1 for the whole model.
 *                     The stage with layer_begin == 0 also holds token_embd; the stage with
 *                     layer_end == n_layer also holds output_norm + output (cf. cpp/src/llama.cpp:5932-5968).
 * b200_ctx_new        replaces llama_new_context_with_model (cpp/bridge.cpp:162; cpp/src/llama.cpp:16592-16993):
 *                     f16 KV cache for the stage's layers (:2926-3022), n_ctx padded to 32 (:16655).            */
b200_model * b200_model_load(const char * gguf_path, int device, int layer_begin, int layer_end);
void         b200_model_free(b200_model * m);

enum { B200_INFO_N_VOCAB = 0, B200_INFO_N_EMBD, B200_INFO_N_LAYER, B200_INFO_N_HEAD, B200_INFO_N_HEAD_KV,
       B200_INFO_N_FF, B200_INFO_HEAD_DIM, B200_INFO_N_CTX_TRAIN, B200_INFO_LAYER_BEGIN, B200_INFO_LAYER_END,
       B200_INFO_FTYPE, B200_INFO_COUNT = 16 };
int     b200_model_info(const b200_model * m, int32_t info[B200_INFO_COUNT]);
/* algorithmic HBM bytes one decoded token reads from this stage's matrices + norm vectors
 * (token_embd excluded: one row) — SURVEY.md §8(d) / BASELINE.md §3 */
int64_t b200_model_weight_bytes(const b200_model * m);

b200_ctx * b200_ctx_new(b200_model * m, int n_ctx);
void       b200_ctx_free(b200_ctx * c);
int        b200_n_ctx(const b200_ctx * c);
void       b200_kv_clear(b200_ctx * c);               /* llama_kv_cache_clear (cpp/bridge.cpp:459) */

/* Rows [pos0, pos0 + n) of one layer's K (post-RoPE) and V cache as f16 bits [n][n_head_kv * head_dim]: the cells of
 * struct llama_kv_cache (cpp/src/llama.cpp:2495-2539), position-major on both sides. Either pointer may be NULL.
 * For parity tests (both sides start from one cache at n_kv in the thousands; K-shift checks), not for serving. */
int b200_kv_write(b200_ctx * c, int layer, int pos0, int n, const uint16_t * k_rows, const uint16_t * v_rows);
int b200_kv_read(b200_ctx * c, int layer, int pos0, int n, uint16_t * k_rows, uint16_t * v_rows);

/* Context shift: llama_kv_cache_seq_rm(ctx, 0, p0, p1) and llama_kv_cache_seq_add(ctx, 0, p0, p1, delta)
 * (cpp/bridge.cpp:500-501; cpp/src/llama.cpp:3154-3206, 3268-3314). Cells keep their places, freed cells are re-used by the
 * next tokens (llama_kv_cache_find_slot, :3028-3125), attention masks by the cells' positions, and the cached K rows of the
 * moved cells are re-rotated by the accumulated delta at the next decode (build_k_shift, :8482-8510) — the same cell order
 * and arithmetic as the reference, so logits stay bit-identical after a shift. Afterwards decode with b200_decode /
 * b200_step_greedy / b200_stage_forward (the device-resident greedy loop and b200_kv_write need an unshifted cache). */
int b200_kv_seq_rm(b200_ctx * c, int p0, int p1);
int b200_kv_seq_add(b200_ctx * c, int p0, int p1, int delta);
/* Self-Extend (cpp/bridge.cpp:509-524): llama_kv_cache_seq_div(ctx, 0, p0, p1, d) (cpp/src/llama.cpp:3316-3349) — the positions
 * of the cells in [p0, p1) are divided by d; K rows re-rotated lazily like after b200_kv_seq_add. */
int b200_kv_seq_div(b200_ctx * c, int p0, int p1, int d);

/* ---- the hot path ------------------------------------------------------------------------------------------
 * b200_decode == llama_decode(ctx, llama_batch_get_one(tokens, n, pos0, 0))  (cpp/bridge.cpp:549-560,
 *                cpp/src/llama.cpp:18517 → llama_decode_internal :14537-14840) followed by
 *                llama_get_logits (cpp/janus.cpp:224): logits_out[n_vocab] receives the LAST token's logits
 *                (the only row llama_batch_get_one keeps). logits_out may be NULL (prompt chunks).
 *                Single-stage contexts only; pipeline stages use b200_stage_step.                            */
int b200_decode(b200_ctx * c, const int32_t * tokens, int n, int pos0, float * logits_out);

/* Device-resident greedy loop: token → decode → argmax → next token, n_steps times, one CUDA-graph replay
 * per token, no host round trip between tokens. Equivalent to the reference loop cpp/bridge.cpp:467-646 with
 * argmax sampling (cpp/bridge.cpp:962-981 sample_top_token). out_tokens[n_steps] receives the sampled ids. */
int b200_generate_greedy(b200_ctx * c, int32_t first_token, int pos0, int n_steps, int32_t * out_tokens);

/* One generated token of the bridge's loop on a single-stage context: llama_decode of `token` at `pos` followed by the
 * arg-max of its logits (cpp/bridge.cpp:549-560, 962-981), as one CUDA-graph replay and one synchronisation. */
int b200_step_greedy(b200_ctx * c, int32_t token, int pos, int32_t * next_token);

/* Per-node taps for layer-wise parity (the counterpart of llama_context_params.cb_eval,
 * cpp/include/llama.h:324-325; node names cpp/src/llama.cpp:13812-13817). When enabled, b200_decode runs
 * un-graphed and keeps host copies of: "Qcur" (post-RoPE), "kqv_merged_cont", "ffn_inp", "ffn_gate_par",
 * "l_out" per layer and "result_output". Returns the element count, copies min(count, cap) floats. */
void    b200_set_taps(b200_ctx * c, int enable);
int64_t b200_get_tap(b200_ctx * c, const char * name, int layer, float * out, int64_t cap);

/* µs-resolution counters (the reference's are ms-truncated at the bridge: cpp/bridge.cpp:650-655) */
void b200_timings(b200_ctx * c, double * t_prompt_us, int64_t * n_prompt, double * t_gen_us, int64_t * n_gen);
void b200_reset_timings(b200_ctx * c);
/* number of kernels launched by this library's own code since the context was created */
int64_t b200_kernel_launches(const b200_ctx * c);
/* device time (CUDA events on the engine's stream) of the last b200_generate_greedy /
 * b200_pipeline_generate_greedy call on this rank, milliseconds */
float   b200_last_device_ms(const b200_ctx * c);
/* one token, un-graphed, with an event pair around every launch: summed device ms and launch count per kernel
 * kind (0 embed, 1 qkv, 2 attention scores+softmax, 3 wo, 4 gate/up, 5 down, 6 head, 7 attention P.V) — the live
 * per-kernel roofline of bench.py */
int     b200_profile_token(b200_ctx * c, int32_t token, int pos, float ms_by_kind[8], int32_t n_by_kind[8]);
/* the launches of ONE kernel kind (numbering as above) of every layer, back to back, `reps` times, between one pair of
 * CUDA events on the engine's stream: ms_total / n_launches = that kernel's steady-state time outside the token's
 * dependency chain (each layer has its own weights: nothing is served from L2) */
int     b200_profile_kind(b200_ctx * c, int kind, int pos, int reps, float * ms_total, int32_t * n_launches);
/* one token through a graph whose kernels stamp %globaltimer (ns) at their phase boundaries (thread 0 of every CTA):
 * out = [n_launches][512 CTAs][12 phases] u64 (0 = not stamped), meta = [n_launches][2] = (kind, CTAs). Returns the
 * number of launches traced (-1 on error). Diagnostic companion of b200_profile_token: shows launch gaps, PDL overlap
 * and prologue/main-loop split inside the replayed graph, which ncu's serialised replays cannot. */
int64_t b200_trace_token(b200_ctx * c, int32_t token, int pos, int reps, uint64_t * out, int64_t cap_words,
                         int32_t * meta, int64_t cap_meta);

/* Prompt batches (b200_decode with n >= 8 on a single-stage context): 1 (default) = the batched kernels of prefill.cuh (every
 * weight tile fetched once per 64 tokens), 0 = token by token through the decode kernels. Same arithmetic, bit-identical
 * logits; the switch exists for A/B measurements and for the test that compares the two. Env BOOSTER_B200_PREFILL_BATCH=0. */
void b200_set_prefill_batch(int on);
/* K-quant prompt batches: 1 (default, the fastest measured) = mma.sync HMMA (k_mma_batch), 2 = tcgen05.mma with TMEM
   accumulators (k_umma_batch), 0 = the dp4a batch kernel; all three are bit-identical */
void b200_set_prefill_mma(int mode);
/* 1 (default): prompt-batch attention through k_attn_softmax_rows + k_attn_pv_batch; 0: the per-token kernels over blockIdx.z */
void b200_set_prefill_attn_batch(int on);
/* Decode path of contexts created AFTER the call: 0 (default) = one kernel per operator joined by programmatic dependent
 * launch, 1 = ONE persistent kernel per token (token_kernel.cuh: phase list + grid barriers; same arithmetic, bit-identical
 * results; measured slower on B200, kept selectable: DESIGN.md §4). Env BOOSTER_B200_TOKEN_KERNEL=1 sets the initial value. */
void b200_set_token_kernel(int on);
/* one token through the persistent per-token kernel with %globaltimer stamps (ns) at its phase boundaries:
 * out = [n_phases][n_ctas][4]: 0 the phase's dependent half starts | 1 it is done | 2 arrived at the grid barrier, next
 * phase's independent half issued | 3 barrier passed. kinds[i] = 1 attention scores, 2 soft-max + P.V, 10 + EPI for a mat-vec
 * (10 head/store, 11 +residual (wo, down), 12 QKV, 13 gate|up). Returns n_phases; 0 if the context runs one kernel per
 * operator; -1 on error. */
int64_t b200_trace_phases(b200_ctx * c, int32_t token, int pos, int reps, uint64_t * out, int64_t cap_words, int32_t * kinds,
                          int64_t cap_kinds, int32_t * n_ctas);

/* µs-resolution per-token timings of a finished bridge job (additive companion of promptEval()/timing(),
 * whose integer-millisecond values read 0 on a B200) */
int b200_job_timing_us(const char * jobID, double * prompt_us_per_token, double * gen_us_per_token);

/* ---- layer-split pipeline over NCCL (replaces ggml_backend_cuda_cpy_tensor_async,
 * cpp/ggml/src/ggml-cuda.cu:2386-2407, invoked from cpp/ggml/src/ggml-backend.c:1782) ---------------------
 * One process per GPU. Rank r owns the stage its model was loaded with. Per token each boundary moves the
 * residual stream f32[n_embd] with ONE ncclSend/ncclRecv; the last stage returns the argmax token to
 * stage 0 with one more 4-byte send/recv (greedy) or the logits stay on the last rank.                      */
int b200_comm_unique_id(uint8_t id[128]);
int b200_comm_init(b200_ctx * c, int rank, int world, const uint8_t id[128]);
/* Direct NVLink hand-off instead of ncclSend / ncclRecv at the stage boundaries (same protocol: one message of f32[n_embd] per
 * boundary per token + the 4-byte token hand-back, cf. cpp/ggml/src/ggml-cuda.cu:2386-2407): the producer's last step stores
 * the vector into the consumer's inbox over NVLink and publishes a sequence number, the consumer's first step spins on it.
 * b200_p2p_handle returns this rank's inbox as a CUDA IPC handle (64 bytes) to be passed to its neighbours through the
 * host-side process group; b200_p2p_connect maps rank + 1's inbox (next_handle; unused on the last rank) and rank 0's
 * (first_handle; used by the last rank). b200_comm_init is still required (the ids of a burst are broadcast with NCCL). */
int b200_p2p_handle(b200_ctx * c, uint8_t handle[64]);
int b200_p2p_connect(b200_ctx * c, int rank, int world, const uint8_t next_handle[64], const uint8_t first_handle[64]);
/* back to ncclSend / ncclRecv (collective decision of the group when a rank could not map a peer's inbox) */
void b200_p2p_disable(b200_ctx * c);
/* All ranks call this collectively: run n_steps greedy tokens through the pipeline, starting from
 * first_token at position pos0. Every rank receives the token ids in out_tokens. */
int b200_pipeline_generate_greedy(b200_ctx * c, int32_t first_token, int pos0, int n_steps, int32_t * out_tokens);
/* Collective: feed n prompt tokens (positions pos0..) through the pipeline; logits_out (may be NULL)
 * is filled on the LAST rank only. */
int b200_pipeline_decode(b200_ctx * c, const int32_t * tokens, int n, int pos0, float * logits_out);

/* ---- in-process layer split (one process, several GPUs — what the Go server uses when `gpus: [..]` names more
 * than one device; replaces ggml_backend_sched_compute_splits + cudaMemcpyPeerAsync + event,
 * cpp/ggml/src/ggml-backend.c:1751-1844, cpp/ggml/src/ggml-cuda.cu:2386-2407) -----------------------------
 * b200_stage_forward runs ONE token through the layers of stage `c`. If `prev` is non-NULL the residual
 * stream f32[n_embd] is first copied from prev's device (peer copy ordered by an event, no host sync).
 * batch_gt1 != 0 selects the reference's batch>1 arithmetic (q rounded to f16 before K.q).
 * b200_stage_logits / b200_stage_argmax synchronise and read the LAST stage's result.                       */
int b200_stage_forward(b200_ctx * c, int32_t token, int pos, int batch_gt1, b200_ctx * prev);
/* one prompt chunk of n <= 512 tokens through THIS stage's layers with the batched prompt kernels (first stage: embeds `tokens`;
   others: take the predecessor's residual streams [n][n_embd] with one peer copy; last stage: logits of the chunk's last token).
   b200_stage_batch_usable(c, n) == 1 says whether the batched kernels can run the chunk on this stage (same answer on every
   stage of a pod); b200_stage_forward_batch returns 2 without doing anything when they cannot. */
int b200_stage_batch_usable(b200_ctx * c, int n);
int b200_stage_forward_batch(b200_ctx * c, const int32_t * tokens, int n, int pos0, b200_ctx * prev);
int b200_stage_sync(b200_ctx * c);    /* wait for everything enqueued on this stage (llama_synchronize, cpp/src/llama.cpp:18536) */
int b200_stage_logits(b200_ctx * c, float * logits_out);
int b200_stage_argmax(b200_ctx * c, int32_t * token_out);
/* llama_get_logits without a second copy: the logits in the context's own pinned host buffer, valid (and writable — the
 * sampler modifies them in place like the reference's, cpp/janus.cpp:224-283) until the next call on the context.
 * b200_decode_view = llama_decode of one token on a single-stage context (one CUDA-graph replay) + that view. NULL on error. */
float * b200_stage_logits_view(b200_ctx * c);
float * b200_decode_view(b200_ctx * c, int32_t token, int pos);

/* launch shape (warps per CTA, warps sharing a 32-row unit, ring stages per warp) the engine picks for a mat-vec whose
 * segments have the given block types: pure host arithmetic, exposed so that tests pin the shapes */
int b200_op_launch_shape(const int32_t * types, int n_types, int64_t n_units, int64_t k, int norm, int sm_count, int32_t wgs[3]);

/* ---- tokenizer (host side) ----------------------------------------------------------------------
 * The vocabulary of a GGUF (tokenizer.ggml.model = "llama" (SPM), "gpt2" (byte-level BPE, LLaMA-3 pre-tokenizer) or
 * "no_vocab" (prompts are decimal ids)), loaded without the weights.
 * b200_tokenize        == llama_tokenize(model, text, text_len, tokens, n_max, add_special, parse_special)
 *                         (cpp/include/llama.h; cpp/src/llama-vocab.cpp:1243-1392,1497-1516): number of tokens, or minus
 *                         that number when n_max is too small; INT32_MIN when the text cannot be tokenized.
 * b200_token_to_piece  == llama_token_to_piece(model, token, buf, length, 0, special) (cpp/src/llama-vocab.cpp:1539-1608):
 *                         bytes written (no terminator), or minus the size needed.
 * b200_token_is_eog    == llama_token_is_eog (cpp/src/llama-vocab.cpp:1433-1439).
 * b200_token_nl        == llama_token_nl (cpp/src/llama-vocab.cpp:1461-1463; the id is found at load, cpp/src/llama.cpp:5585-5597):
 *                         the newline token whose logit the standard sampling chain restores after its penalties; -1 if none.
 * The bridge (doInference / status) uses exactly these with add_special = 0, parse_special = 1, special = 1
 * (cpp/bridge.cpp:275-278, 630, 640). */
typedef struct b200_tokenizer b200_tokenizer;
b200_tokenizer * b200_tokenizer_load(const char * gguf_path);
void    b200_tokenizer_free(b200_tokenizer * t);
int32_t b200_tokenizer_n_vocab(const b200_tokenizer * t);
int32_t b200_tokenize(const b200_tokenizer * t, const char * text, int32_t text_len, int32_t * tokens, int32_t n_max,
                      int add_special, int parse_special);
int32_t b200_token_to_piece(const b200_tokenizer * t, int32_t token, char * buf, int32_t length, int special);
int     b200_token_is_eog(const b200_tokenizer * t, int32_t token);
int32_t b200_token_nl(const b200_tokenizer * t);
/* ---- samplers (host side) ----------------------------------------------------------------------------
 * What doInference does with the logits of every token, exposed on its own so that it can be pinned against the reference's
 * sampler on the reference's logits (no GPU involved): janus != 0 -> sample_janus_token (cpp/janus.cpp:191-331) with the
 * tables of initJanus (:405-700) built from the GGUF's vocabulary; janus == 0 -> llama_sampling_sample's default chain
 * (llama_sampling_sample, cpp/common/sampling.cpp:271-340 over cpp/src/llama-sampling.cpp: repetition penalty over the last
 * accepted tokens, then top-k, tail-free, typical, top-p, min-p 0.05, temperature and a draw — or mirostat 1 / 2), which the
 * reference links but its bridge never calls (cpp/bridge.cpp:598). b200_sampler_set_standard carries the chain's parameters that
 * are not in b200_sampler_new's initContext-shaped signature (mirostat, mirostat_tau, mirostat_eta, typical_p of initContext;
 * tfs_z and min_p, which initContext cannot set). b200_sampler_reset starts a job (prompt ids, rng seed = llama_set_rng_seed);
 * b200_sampler_sample draws the next token from logits[n_vocab] (Janus modifies them in place) at position pos = number of
 * tokens decoded so far. */
typedef struct b200_sampler b200_sampler;
b200_sampler * b200_sampler_new(const char * gguf_path, int n_ctx, int32_t janus, int32_t depth, float scale, float hi, float lo,
                                float temperature, int top_k, float top_p, float repetition_penalty, int penalty_last_n);
void    b200_sampler_set_standard(b200_sampler * s, int32_t mirostat, float mirostat_tau, float mirostat_eta, float typical_p,
                                  float tfs_z, float min_p);
void    b200_sampler_free(b200_sampler * s);
void    b200_sampler_reset(b200_sampler * s, const int32_t * prompt, int32_t n_prompt, uint32_t seed);
int32_t b200_sampler_sample(b200_sampler * s, float * logits, int32_t pos, int32_t n_predict);

/* codepoint classes of the LLaMA-3 pre-tokenizer regex: bit 0 \p{L}, bit 1 \p{N}, bit 2 \s (cpp/src/unicode.h:8-46,
 * cpp/src/unicode-data.cpp) */
int     b200_cpt_class(uint32_t cp);

/* ---- operator-level entry points (parity tests; host pointers, compute on the GPU with the SAME kernels
 * the engine launches) ----------------------------------------------------------------------------------- */
/* quantize_row_q8_K (cpp/ggml/src/ggml-quants.c:3593-3630): out = block_q8_K[k/256] in ggml layout (292 B each) */
int b200_op_quantize_q8_K(const float * x, int64_t k, void * out);
/* quantize_row_q8_0, AVX path (cpp/ggml/src/ggml-quants.c:866-1000): out = block_q8_0[k/32] (34 B each) */
int b200_op_quantize_q8_0(const float * x, int64_t k, void * out);
/* dequantize_row_q4_K/q5_K/q6_K/q8_0 (cpp/ggml/src/ggml-quants.c:2548,2756,2970,1609) = the embedding
 * get_rows path (cpp/ggml/src/ggml.c:13186): w = one row of blocks in ggml layout */
int b200_op_dequantize_row(int type, const void * w, int64_t k, float * y);
/* ggml_compute_forward_mul_mat at batch 1 (cpp/ggml/src/ggml.c:12277-12490) with vec_dot_{q4_K,q5_K,q6_K}_q8_K /
 * q8_0_q8_0 (cpp/ggml/src/ggml-quants.c:6832,7400,8037,5227): y[n_rows] = W[n_rows x k] . x[k];
 * w = n_rows rows of blocks in ggml row-major layout */
int b200_op_mul_mat_vec(int type, const void * w, int64_t n_rows, int64_t k, const float * x, float * y);
/* batch > 1 (ggml_compute_forward_mul_mat with ne11 = T, cpp/ggml/src/ggml.c:12277): y[T][n_rows] for x[T][k], T <= 512, through the
   prompt-batch kernels */
int b200_op_mul_mat(int type, const void * w, int64_t n_rows, int64_t k, const float * x, int64_t T, float * y);
/* ggml_compute_forward_rms_norm_f32 then ggml_mul by the weight (cpp/ggml/src/ggml.c:11850-11896,
 * cpp/src/llama.cpp:7928-7958). w may be NULL. */
int b200_op_rms_norm(const float * x, const float * w, int64_t k, float eps, float * y);
/* ggml_compute_forward_rope_f32, NORM mode (cpp/ggml/src/ggml.c:14043-14166): x[n_heads][head_dim] in place
 * for position pos. freq_factors may be NULL. */
int b200_op_rope(float * x, int n_heads, int head_dim, int pos, float freq_base, float freq_scale,
                 const float * freq_factors);
/* the default (non-flash) attention route (cpp/src/llama.cpp:8188-8299): q[n_head][hd] f32,
 * k_cache/v_cache = f16 bits [n_kv][n_head_kv*hd] (position-major), out[n_head*hd].
 * round_q = 0: batch-1 arithmetic (tinyBLAS F16xF32); 1: batch>1 (q rounded to f16, ggml_vec_dot_f16). */
int b200_op_attention(const float * q, const uint16_t * k_cache, const uint16_t * v_cache, int n_kv,
                      int n_head, int n_head_kv, int head_dim, float scale, int round_q, float * out);
/* which kernels run the attention: 0 = automatic (scores + fused softmax/P.V when the GQA score rows of the context fit
 * one CTA's shared memory, else the three-launch long-context route), 1 = always the long-context route. Process-wide;
 * exists so that tests cover both routes at every size. */
void b200_set_attention_route(int route);

#ifdef __cplusplus
}
#endif
#endif /* BOOSTER_B200_H */
