// oracle/ref_shim.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.
//
// A thin C-ABI veneer compiled INTO oracle/_ref/libbooster_cpu_ref*.so next to the
// UNMODIFIED reference sources (see oracle/Makefile). It exists because the reference's
// llama.h API passes large structs by value (llama_model_params, llama_context_params,
// llama_batch), which ctypes cannot do portably. Every function below only calls the
// reference's public API (cpp/include/llama.h, cpp/ggml/include/ggml.h); it contains no
// arithmetic of its own.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
// load the resulting library.
//
// Reference call sites mirrored:
//   model load / context creation: cpp/bridge.cpp:118-171 (init_context)
//   decode loop:                   cpp/bridge.cpp:549-560 (llama_decode + llama_batch_get_one)
//   logits:                        cpp/janus.cpp:224 (llama_get_logits)
//   node tap:                      cpp/include/llama.h:324-325 (cb_eval), cpp/src/llama.cpp:14707
//   tokenizer (vocab-only load):   cpp/bridge.cpp:275-278 (llama_tokenize, add_special=false, parse_special=true),
//                                  cpp/bridge.cpp:630 (llama_token_to_piece), :640 (llama_token_is_eog)
//   sampler:                       cpp/bridge.cpp:196 (initJanus), :437-438, :586-603 (sample_janus_token + last_tokens)
//   context shift:                 cpp/bridge.cpp:500-503 (llama_kv_cache_seq_rm / llama_kv_cache_seq_add)
//   standard sampling chain:       cpp/bridge.cpp:456, 598, 605, 618 (llama_sampling_init / _sample / _accept), :763-776

#include "llama.h"
#include "ggml.h"
#include "ggml-backend.h"
#include "common.h"
#include "sampling.h"
#include "janus.h"

#include <cstdint>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

namespace {

struct ref_handle {
    llama_model   * model = nullptr;
    llama_context * ctx   = nullptr;
    std::set<std::string> tap_names;                 // exact tensor names to capture
    std::map<std::string, std::vector<float>> taps;  // captured values (as f32)
    std::map<std::string, std::vector<int64_t>> tap_shapes;
};

bool tap_cb(struct ggml_tensor * t, bool ask, void * ud) {
    auto * h = static_cast<ref_handle *>(ud);
    const bool want = h->tap_names.count(t->name) > 0;
    if (ask) return want;
    if (!want) return true;
    const int64_t n = ggml_nelements(t);
    std::vector<float> out((size_t) n);
    // tensors on the path are f32 except KV views (f16); convert through ggml's own helpers
    if (t->type == GGML_TYPE_F32 && ggml_is_contiguous(t)) {
        ggml_backend_tensor_get(t, out.data(), 0, (size_t) n * sizeof(float));
    } else if (t->type == GGML_TYPE_F16 && ggml_is_contiguous(t)) {
        std::vector<ggml_fp16_t> tmp((size_t) n);
        ggml_backend_tensor_get(t, tmp.data(), 0, (size_t) n * sizeof(ggml_fp16_t));
        ggml_fp16_to_fp32_row(tmp.data(), out.data(), n);
    } else if (t->type == GGML_TYPE_F32) {
        // strided f32 view: walk the 4-d index space
        const char * base = (const char *) t->data;
        int64_t k = 0;
        for (int64_t i3 = 0; i3 < t->ne[3]; i3++)
        for (int64_t i2 = 0; i2 < t->ne[2]; i2++)
        for (int64_t i1 = 0; i1 < t->ne[1]; i1++)
        for (int64_t i0 = 0; i0 < t->ne[0]; i0++)
            out[(size_t) k++] = *(const float *)(base + i0*t->nb[0] + i1*t->nb[1] + i2*t->nb[2] + i3*t->nb[3]);
    } else {
        return true;  // not a type we tap
    }
    h->taps[t->name] = std::move(out);
    h->tap_shapes[t->name] = { t->ne[0], t->ne[1], t->ne[2], t->ne[3] };
    return true;
}

void log_nothing(ggml_log_level, const char *, void *) {}

}  // namespace

extern "C" {

void refshim_init(int quiet) {
    if (quiet) llama_log_set(log_nothing, nullptr);
    llama_backend_init();
}

// ftype: 7 = Q8_0, 15 = Q4_K_M, 17 = Q5_K_M, 18 = Q6_K (enum llama_ftype, cpp/include/llama.h)
int refshim_quantize(const char * fin, const char * fout, int ftype, int nthread) {
    llama_model_quantize_params qp = llama_model_quantize_default_params();
    qp.ftype   = (llama_ftype) ftype;
    qp.nthread = nthread;
    return (int) llama_model_quantize(fin, fout, &qp);
}

void * refshim_load(const char * path, int n_ctx, int n_batch, int n_threads, int flash_attn) {
    auto * h = new ref_handle();
    llama_model_params mp = llama_model_default_params();
    mp.n_gpu_layers = 0;
    mp.use_mmap     = true;
    h->model = llama_load_model_from_file(path, mp);
    if (!h->model) { delete h; return nullptr; }
    llama_context_params cp = llama_context_default_params();
    cp.n_ctx           = (uint32_t) n_ctx;
    cp.n_batch         = (uint32_t) n_batch;
    cp.n_ubatch        = (uint32_t) (n_batch < 512 ? n_batch : 512);   // common.h:81 default n_ubatch = 512
    cp.n_threads       = (uint32_t) n_threads;
    cp.n_threads_batch = (uint32_t) n_threads;
    cp.flash_attn      = flash_attn != 0;
    cp.cb_eval           = tap_cb;
    cp.cb_eval_user_data = h;
    h->ctx = llama_new_context_with_model(h->model, cp);
    if (!h->ctx) { llama_free_model(h->model); delete h; return nullptr; }
    return h;
}

void refshim_free(void * hv) {
    auto * h = static_cast<ref_handle *>(hv);
    if (!h) return;
    if (h->ctx)   llama_free(h->ctx);
    if (h->model) llama_free_model(h->model);
    delete h;
}

int refshim_n_vocab(void * hv) { return llama_n_vocab(static_cast<ref_handle *>(hv)->model); }
int refshim_n_ctx(void * hv)   { return (int) llama_n_ctx(static_cast<ref_handle *>(hv)->ctx); }

void refshim_kv_clear(void * hv) { llama_kv_cache_clear(static_cast<ref_handle *>(hv)->ctx); }

// One llama_decode over tokens[0..n) at positions pos0.. (sequence 0); copies the LAST token's
// logits (the only row llama_batch_get_one keeps) into logits_out[n_vocab]. Returns llama_decode's code.
int refshim_decode(void * hv, const int32_t * tokens, int n, int pos0, float * logits_out) {
    auto * h = static_cast<ref_handle *>(hv);
    std::vector<llama_token> toks(tokens, tokens + n);
    const int rc = llama_decode(h->ctx, llama_batch_get_one(toks.data(), n, pos0, 0));
    if (rc != 0) return rc;
    if (logits_out) {
        const float * lg = llama_get_logits(h->ctx);
        std::memcpy(logits_out, lg, sizeof(float) * (size_t) llama_n_vocab(h->model));
    }
    return 0;
}

// comma-separated exact node names, e.g. "l_out-0,Qcur-0,result_output"; empty string clears
void refshim_set_taps(void * hv, const char * csv) {
    auto * h = static_cast<ref_handle *>(hv);
    h->tap_names.clear();
    h->taps.clear();
    h->tap_shapes.clear();
    std::string s = csv ? csv : "";
    size_t p = 0;
    while (p < s.size()) {
        size_t q = s.find(',', p);
        if (q == std::string::npos) q = s.size();
        if (q > p) h->tap_names.insert(s.substr(p, q - p));
        p = q + 1;
    }
}

// returns element count of the captured node (0 if not captured); copies min(count, cap) floats.
// NB: nodes with the same name are emitted several times per layer (e.g. "Qcur-0" for MUL_MAT,
// RESHAPE and ROPE); the LAST one evaluated wins, which for Qcur/Kcur is the post-RoPE tensor.
int64_t refshim_get_tap(void * hv, const char * name, float * out, int64_t cap, int64_t * shape4) {
    auto * h = static_cast<ref_handle *>(hv);
    auto it = h->taps.find(name);
    if (it == h->taps.end()) return 0;
    const int64_t n = (int64_t) it->second.size();
    if (out) std::memcpy(out, it->second.data(), sizeof(float) * (size_t) (n < cap ? n : cap));
    if (shape4) for (int i = 0; i < 4; i++) shape4[i] = h->tap_shapes[name][(size_t) i];
    return n;
}

void refshim_reset_timings(void * hv) { llama_reset_timings(static_cast<ref_handle *>(hv)->ctx); }

// µs-resolution counters kept inside the reference (cpp/src/llama.cpp:18528-18552, 19199-19214)
void refshim_timings(void * hv, double * t_p_eval_ms, int * n_p_eval, double * t_eval_ms, int * n_eval) {
    const llama_timings t = llama_get_timings(static_cast<ref_handle *>(hv)->ctx);
    *t_p_eval_ms = t.t_p_eval_ms; *n_p_eval = t.n_p_eval;
    *t_eval_ms   = t.t_eval_ms;   *n_eval   = t.n_eval;
}

// ---- context shift: the two calls of cpp/bridge.cpp:500-503; the K-shift itself runs inside the next llama_decode
// (llama_kv_cache_update -> build_k_shift, cpp/src/llama.cpp:8482-8510)
void refshim_kv_seq_rm(void * hv, int p0, int p1)             { llama_kv_cache_seq_rm(static_cast<ref_handle *>(hv)->ctx, 0, p0, p1); }
void refshim_kv_seq_div(void * hv, int p0, int p1, int d)     { llama_kv_cache_seq_div(static_cast<ref_handle *>(hv)->ctx, 0, p0, p1, d); }   // cpp/bridge.cpp:518
void refshim_kv_seq_add(void * hv, int p0, int p1, int delta) { llama_kv_cache_seq_add(static_cast<ref_handle *>(hv)->ctx, 0, p0, p1, delta); }

// ---- sampler oracle: the generation loop of cpp/bridge.cpp with the reference's own initJanus / sample_janus_token
// (cpp/janus.cpp, compiled unmodified): clear the cache, decode the prompt as one batch, then n_gen times
// { id = sample_janus_token(...); shift last_tokens; decode id } — last_tokens is n_ctx zeros that only GENERATED tokens
// enter (cpp/bridge.cpp:437-438, 602-603); the rng is seeded like cpp/bridge.cpp:216-217 but with the caller's seed.
// Stops after an end-of-generation token (cpp/bridge.cpp:640). Returns the number of ids written.
int refshim_janus_generate(void * hv, const int32_t * prompt, int n_prompt, int n_gen, int depth, float scale, float hi, float lo,
                           uint32_t seed, int n_predict, int32_t * out_ids) {
    auto * h = static_cast<ref_handle *>(hv);
    janus_params jp;
    jp.janus = 1; jp.depth = depth; jp.scale = scale; jp.hi = hi; jp.lo = lo;
    llama_sampling_params sp;
    initJanus(h->ctx, jp, nullptr);
    llama_set_rng_seed(h->ctx, seed);
    llama_kv_cache_clear(h->ctx);
    std::vector<llama_token> toks(prompt, prompt + n_prompt);
    if (llama_decode(h->ctx, llama_batch_get_one(toks.data(), n_prompt, 0, 0))) return -1;
    std::vector<llama_token> last_tokens((size_t) llama_n_ctx(h->ctx), 0);
    int n_past = n_prompt, n = 0;
    for (int i = 0; i < n_gen; i++) {
        llama_token id = sample_janus_token(h->ctx, sp, jp, last_tokens, (size_t) n_prompt, (size_t) n_past, (size_t) n_predict);
        last_tokens.erase(last_tokens.begin());
        last_tokens.push_back(id);
        out_ids[n++] = id;
        if (llama_token_is_eog(h->model, id)) break;
        if (llama_decode(h->ctx, llama_batch_get_one(&id, 1, n_past, 0))) return -1;
        n_past += 1;
    }
    return n;
}

// ---- standard-chain oracle: the generation loop of cpp/bridge.cpp with the branch the reference keeps commented out
// (cpp/bridge.cpp:586-599 "FIXME: Allow standard samplings": id = llama_sampling_sample(ctx_sampling, ctx, ctx_guidance)) taken
// instead of sample_janus_token. Everything is the reference's own code (common/sampling.cpp, src/llama-sampling.cpp):
// llama_sampling_init on the parameters initContext stores (cpp/bridge.cpp:763-776), llama_set_rng_seed (cpp/bridge.cpp:216-217; the
// mirostat draws use the context's rng), prompt tokens accepted without grammar (:618), every sampled token accepted (:605).
// The sampling context's own rng is seeded with the caller's seed too (the reference leaves it to std::random_device).
int refshim_standard_generate(void * hv, const int32_t * prompt, int n_prompt, int n_gen,
                              int mirostat, float mirostat_tau, float mirostat_eta,
                              float temp, int top_k, float top_p, float typical_p, float tfs_z, float min_p,
                              float penalty_repeat, int penalty_last_n, uint32_t seed, int32_t * out_ids) {
    auto * h = static_cast<ref_handle *>(hv);
    llama_sampling_params sp;
    sp.mirostat = mirostat; sp.mirostat_tau = mirostat_tau; sp.mirostat_eta = mirostat_eta;
    sp.temp = temp; sp.top_k = top_k; sp.top_p = top_p;
    sp.typical_p = typical_p > 0 ? typical_p : 1.0f;          // cpp/bridge.cpp:773
    sp.tfs_z = tfs_z; sp.min_p = min_p;
    sp.penalty_repeat = penalty_repeat; sp.penalty_last_n = penalty_last_n;
    sp.seed = seed;
    llama_sampling_context * cs = llama_sampling_init(sp);
    if (!cs) return -1;
    llama_set_rng_seed(h->ctx, seed);
    llama_kv_cache_clear(h->ctx);
    std::vector<llama_token> toks(prompt, prompt + n_prompt);
    if (llama_decode(h->ctx, llama_batch_get_one(toks.data(), n_prompt, 0, 0))) { llama_sampling_free(cs); return -1; }
    for (int i = 0; i < n_prompt; i++) llama_sampling_accept(cs, h->ctx, toks[(size_t) i], false);
    int n_past = n_prompt, n = 0;
    for (int i = 0; i < n_gen; i++) {
        llama_token id = llama_sampling_sample(cs, h->ctx, nullptr);
        llama_sampling_accept(cs, h->ctx, id, true);
        out_ids[n++] = id;
        if (llama_token_is_eog(h->model, id)) break;
        if (llama_decode(h->ctx, llama_batch_get_one(&id, 1, n_past, 0))) { llama_sampling_free(cs); return -1; }
        n_past += 1;
    }
    llama_sampling_free(cs);
    return n;
}
int refshim_token_nl(void * hv) { return llama_token_nl(static_cast<ref_handle *>(hv)->model); }

// ---- tokenizer oracle: the reference's own llama_tokenize / llama_token_to_piece / llama_token_is_eog on a
// vocab-only load of a GGUF (llama_model_params.vocab_only, cpp/include/llama.h)
void * refshim_vocab_load(const char * path) {
    llama_model_params mp = llama_model_default_params();
    mp.vocab_only = true;
    mp.n_gpu_layers = 0;
    return llama_load_model_from_file(path, mp);
}
void refshim_vocab_free(void * m) { if (m) llama_free_model(static_cast<llama_model *>(m)); }
int refshim_tokenize(void * m, const char * text, int text_len, int32_t * out, int cap, int add_special, int parse_special) {
    return llama_tokenize(static_cast<llama_model *>(m), text, text_len, out, cap, add_special != 0, parse_special != 0);
}
int refshim_token_to_piece(void * m, int32_t token, char * buf, int cap, int special) {
    return llama_token_to_piece(static_cast<llama_model *>(m), token, buf, cap, 0, special != 0);
}
int refshim_vocab_token_nl(void * m) { return llama_token_nl(static_cast<llama_model *>(m)); }
int refshim_token_is_eog(void * m, int32_t token) { return llama_token_is_eog(static_cast<llama_model *>(m), token) ? 1 : 0; }

}  // extern "C"

// codepoint category flags of the reference's tables (cpp/src/unicode.h:8-46, unicode-data.cpp), for the test that pins
// booster_b200/csrc/unicode_tables.hpp
#include "unicode.h"
extern "C" uint16_t refshim_cpt_flags(uint32_t cp) { return unicode_cpt_flags(cp).as_uint(); }
