// bridge.cpp — the nine cgo entry points of include/bridge.h on top of the B200 engine.
//
// Host-side mirror of cpp/bridge.cpp's operator interface for the hot path: same symbols, same argument
// meaning, same return codes. The generation loop follows the reference's control flow
// (cpp/bridge.cpp:175-658: tokenize -> reject if > n_ctx-4 -> clear KV -> prompt in n_batch chunks ->
// sample/decode one token at a time until EOG / n_predict / n_ctx-4 / stop flag -> per-token timings), but every
// llama_decode is the CUDA engine. Differences, all documented in DESIGN.md / INTEGRATION.md:
//   * sampler: the reference's Janus sampler restated on the host (janus.hpp; logits come back once per token), or — with
//     janus = 0, a setting the reference ignores — the standard chain; temperature <= 0 / top_k = 1 there is greedy
//     arg-max on the device;
//   * tokenizer: models with tokenizer.ggml.model == "no_vocab" take prompts that are white-space separated
//     token ids and render pieces as "<id> "; text tokenizers (BPE/SPM) are row f-1, see tokenizer.hpp;
//   * context shift (cpp/bridge.cpp:487-507): same branch, same unreachability (the loop ends at n_ctx - 4 first); the KV
//     operations behind it are b200_kv_seq_rm / b200_kv_seq_add (engine.cu), bit-exact against the reference. Self-Extend
//     (ga_n > 1, :509-524) cannot be switched on through initContext in the reference either and is not built;
//   * status() returns a per-thread snapshot instead of a pointer into a string another thread appends to
//     (the reference race).
#include "../../include/bridge.h"
#include "../../include/booster_b200.h"
#include "tokenizer.hpp"
#include "janus.hpp"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace {

constexpr int MAX_PODS = 8;   // cpp/bridge.cpp:102-110: all per-pod arrays are sized 8

struct Pod {
    std::string model_path;
    int n_ctx = 0, n_batch = 512, n_predict = -1;
    int ga_n = 1, ga_w = 512;           // Self-Extend: gpt_params.grp_attn_n / grp_attn_w (cpp/common/common.h defaults; cpp/bridge.cpp:430-431)
    uint32_t seed = 0;
    std::vector<b200_model *> models;   // one per stage (in-process layer split)
    std::vector<b200_ctx *>   stages;
    std::unique_ptr<b200::Tokenizer> tok;
    int n_vocab = 0;
    // samplers (cpp/bridge.cpp:746-761 stores the same parameters per pod)
    b200::JanusParams jparams;
    b200::StandardParams sparams;
    b200::JanusSampler janus;           // scales / types tables: initJanus, once per pod (the reference redoes it per job)
    // contexts first, then the models they point into
    void release() {
        for (b200_ctx * c : stages) b200_ctx_free(c);
        for (b200_model * m : models) b200_model_free(m);
        stages.clear(); models.clear(); tok.reset();
    }
};

Pod               g_pods[MAX_PODS];
std::atomic<bool> g_stop[MAX_PODS];
bool              g_debug_cuda = false;

struct Job {
    std::string text;           // prompt pieces + generated pieces (cpp/bridge.cpp:628-632)
    int64_t prompt_eval_ms = 0; // ms per prompt token, integer-truncated (cpp/bridge.cpp:653)
    int64_t timing_ms = 0;      // ms per generated token (cpp/bridge.cpp:654)
    int64_t prompt_tokens = 0;
    uint32_t seed = 0;
    double  prompt_us_per_tok = 0, gen_us_per_tok = 0;   // µs-resolution copies (additive, see b200_job_timing_us)
};
std::mutex                 g_mu;
std::map<std::string, Job> g_jobs;

double now_us() {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// layer -> device assignment by cumulative proportions, as llm_load_tensors does
// (cpp/src/llama.cpp:5932-5968: normalised cumulative splits, std::upper_bound on (i / act_gpu_layers))
std::vector<int> split_layers(int n_layer, const std::vector<float> & prop) {
    std::vector<float> cum(prop.size());
    float sum = 0.f;
    for (size_t i = 0; i < prop.size(); i++) { sum += prop[i]; cum[i] = sum; }
    for (auto & c : cum) c /= sum;
    const int act = n_layer + 1;   // all layers + output are offloaded (cpp/bridge.cpp:746-750: n_gpu_layers = sum)
    std::vector<int> dev((size_t) n_layer);
    for (int il = 0; il < n_layer; il++) {
        const float f = (float) il / (float) act;
        size_t d = 0;
        while (d + 1 < cum.size() && !(f < cum[d])) d++;   // upper_bound
        dev[(size_t) il] = (int) d;
    }
    return dev;
}

int run_token(Pod & p, int32_t token, int pos, int batch_gt1) {
    b200_ctx * prev = nullptr;
    for (b200_ctx * c : p.stages) {
        if (b200_stage_forward(c, token, pos, batch_gt1, prev) != 0) return 1;
        prev = c;
    }
    return 0;
}

}  // namespace

extern "C" {

void init(char * swap, char * debug) {
    (void) swap;   // session directory: accepted and ignored, as in the reference (sessions are commented out)
    g_debug_cuda = debug && std::strstr(debug, "cuda") != nullptr;
    for (auto & s : g_stop) s.store(false);
}

static void * init_context_impl(
    int idx, char * modelName, int threads, int batch_size,
    int gpu1, int gpu2, int gpu3, int gpu4,
    int context, int predict,
    int32_t mirostat, float mirostat_tau, float mirostat_eta,
    float temperature, int top_k, float top_p, float typical_p,
    float repetition_penalty, int penalty_last_n,
    int32_t janus, int32_t depth, float scale, float hi, float lo,
    uint32_t seed, char * debug) {
    (void) threads; (void) debug;
    // mirostat and typical-p belong to the standard chain, which the reference's bridge never calls (cpp/bridge.cpp:586-599):
    // with janus != 0 they have no effect here either; with janus = 0 the chain runs with them
    if (janus != 0 && (mirostat != 0 || (typical_p > 0.f && typical_p < 1.f)))
        std::fprintf(stderr, "initContext: mirostat / typical_p have no effect with janus != 0 (as in the reference: cpp/bridge.cpp:586-596); janus = 0 selects the standard chain\n");
    if (idx < 0 || idx >= MAX_PODS || !modelName) return nullptr;
    Pod & p = g_pods[idx];
    p.release();                          // re-initialising a pod frees its previous weights and KV cache
    p = Pod();
    p.model_path = modelName;
    p.seed = seed;
    p.n_predict = predict;
    const int n_dev = b200_device_count();
    if (n_dev <= 0) {
        std::fprintf(stderr, "initContext: no CUDA device: %s\n", "booster_b200 has no CPU path");
        return nullptr;
    }
    // proportions: gpu1..gpu4 (server.go:514-530), or env BOOSTER_B200_SPLIT="p0,p1,..,p7" for up to 8 devices
    std::vector<float> prop;
    if (const char * env = std::getenv("BOOSTER_B200_SPLIT")) {
        std::string s = env; size_t q = 0;
        while (q < s.size()) { size_t e = s.find(',', q); if (e == std::string::npos) e = s.size(); prop.push_back((float) std::atof(s.substr(q, e - q).c_str())); q = e + 1; }
    } else {
        prop = { (float) gpu1, (float) gpu2, (float) gpu3, (float) gpu4 };
    }
    if ((int) prop.size() > n_dev) {
        for (size_t i = (size_t) n_dev; i < prop.size(); i++) if (prop[i] > 0) { std::fprintf(stderr, "initContext: split names device %zu but only %d visible\n", i, n_dev); return nullptr; }
        prop.resize((size_t) n_dev);
    }
    for (float & f : prop) if (!(f > 0.f)) f = 0.f;   // negative or NaN shares name no device
    float sum = 0; for (float f : prop) sum += f;
    if (sum <= 0) { prop.assign(1, 1.f); }   // the reference would run on the CPU; this library is the GPU path

    // peek hyper-parameters with a first-stage load of the first device that has a share
    // (cheap way: load stage ranges after reading n_layer from a probe load of layer 0 only)
    int first_dev = 0; while (first_dev < (int) prop.size() && prop[(size_t) first_dev] <= 0) first_dev++;
    b200_model * probe = b200_model_load(modelName, first_dev, 0, 1);
    if (!probe) { std::fprintf(stderr, "initContext: %s\n", b200_last_error()); return nullptr; }
    int32_t info[B200_INFO_COUNT];
    b200_model_info(probe, info);
    b200_model_free(probe);
    const int n_layer = info[B200_INFO_N_LAYER];
    const std::vector<int> dev = split_layers(n_layer, prop);
    int n_ctx = context > 0 ? context : info[B200_INFO_N_CTX_TRAIN];
    for (int il = 0; il < n_layer;) {
        int e = il; while (e < n_layer && dev[(size_t) e] == dev[(size_t) il]) e++;
        b200_model * m = b200_model_load(modelName, dev[(size_t) il], il, e);
        if (!m) { std::fprintf(stderr, "initContext: %s\n", b200_last_error()); p.release(); return nullptr; }
        p.models.push_back(m);
        b200_ctx * c = b200_ctx_new(m, n_ctx);
        if (!c) { std::fprintf(stderr, "initContext: %s\n", b200_last_error()); p.release(); return nullptr; }
        p.stages.push_back(c);
        il = e;
    }
    p.n_ctx = b200_n_ctx(p.stages[0]);
    p.n_vocab = info[B200_INFO_N_VOCAB];
    // batch (cpp/bridge.cpp:154-160): 0 -> 512 on GPU
    p.n_batch = (batch_size > 0 && batch_size <= p.n_ctx) ? batch_size : 512;
    std::string terr;
    p.tok = b200::make_tokenizer(modelName, terr);
    if (!p.tok) { std::fprintf(stderr, "initContext: %s\n", terr.c_str()); p.release(); return nullptr; }
    // Self-Extend: the reference reads gpt_params.grp_attn_n / grp_attn_w (cpp/bridge.cpp:430-431), which initContext cannot
    // set (always 1 / 512 there); here the two come from the environment so that the branch can be switched on
    if (const char * e = std::getenv("BOOSTER_B200_GRP_ATTN_N")) p.ga_n = std::atoi(e);
    if (const char * e = std::getenv("BOOSTER_B200_GRP_ATTN_W")) p.ga_w = std::atoi(e);
    // sampler parameters (cpp/bridge.cpp:746-761)
    p.jparams.janus = janus; p.jparams.depth = depth; p.jparams.scale = scale; p.jparams.hi = hi; p.jparams.lo = lo;
    p.sparams.temp = temperature; p.sparams.top_k = top_k; p.sparams.top_p = top_p;
    p.sparams.penalty_repeat = repetition_penalty; p.sparams.penalty_last_n = penalty_last_n;
    p.sparams.mirostat = mirostat; p.sparams.mirostat_tau = mirostat_tau; p.sparams.mirostat_eta = mirostat_eta;
    p.sparams.typical_p = typical_p > 0 ? typical_p : 1.0f;       // cpp/bridge.cpp:773
    p.sparams.nl_token = p.tok->linefeed();
    if (janus != 0) p.janus.init(*p.tok, p.jparams, 0);
    return (void *) &p;
}

// No C++ exception may cross the cgo boundary (it would terminate the Go server): allocation failures and anything else
// unexpected end the call with the reference's failure value (nullptr / 0) and a line on stderr.
void * initContext(
    int idx, char * modelName, int threads, int batch_size,
    int gpu1, int gpu2, int gpu3, int gpu4,
    int context, int predict,
    int32_t mirostat, float mirostat_tau, float mirostat_eta,
    float temperature, int top_k, float top_p, float typical_p,
    float repetition_penalty, int penalty_last_n,
    int32_t janus, int32_t depth, float scale, float hi, float lo,
    uint32_t seed, char * debug) {
    try {
        return init_context_impl(idx, modelName, threads, batch_size, gpu1, gpu2, gpu3, gpu4, context, predict, mirostat, mirostat_tau,
                                 mirostat_eta, temperature, top_k, top_p, typical_p, repetition_penalty, penalty_last_n, janus, depth,
                                 scale, hi, lo, seed, debug);
    } catch (const std::exception & e) {
        std::fprintf(stderr, "initContext: %s\n", e.what());
    } catch (...) {
        std::fprintf(stderr, "initContext: unknown failure\n");
    }
    if (idx >= 0 && idx < MAX_PODS) g_pods[idx].release();
    return nullptr;
}

static int64_t do_inference_impl(int idx, void * ctx, char * jobID, char * sessionID, char * prompt) {
    (void) sessionID;
    if (idx < 0 || idx >= MAX_PODS || !ctx || !jobID || !prompt) return 0;
    Pod & p = g_pods[idx];
    if ((void *) &p != ctx || p.stages.empty()) return 0;
    const std::string job = jobID;         // callee copies: the Go side never frees its C strings (server.go:73)
    const std::string text = prompt;
    g_stop[idx].store(false);

    const uint32_t seed = p.seed ? p.seed : (uint32_t) std::time(nullptr);   // cpp/bridge.cpp:216-221
    { std::lock_guard<std::mutex> lk(g_mu); Job & j = g_jobs[job]; j = Job(); j.seed = seed; }

    std::vector<int32_t> inp;
    // llama_tokenize(model, prompt, add_special = false, parse_special = true): cpp/bridge.cpp:275-278
    if (!p.tok->tokenize(text, false, true, inp)) return 0;
    { std::lock_guard<std::mutex> lk(g_mu); g_jobs[job].prompt_tokens = (int64_t) inp.size(); }
    const int max_embd = p.n_ctx - 4;
    if ((int) inp.size() > max_embd || inp.empty()) {     // cpp/bridge.cpp:382-386
        std::fprintf(stderr, "doInference: prompt is too long (%d tokens, max %d) or empty\n", (int) inp.size(), max_embd);
        return 0;
    }
    if (p.ga_n != 1 && p.ga_n <= 0) return 0;               // cpp/bridge.cpp:433: grp_attn_n must be positive
    if (p.ga_n != 1 && (p.ga_w % p.ga_n != 0)) return 0;    // cpp/bridge.cpp:434: grp_attn_w must be a multiple of grp_attn_n
    for (b200_ctx * c : p.stages) b200_kv_clear(c);        // cpp/bridge.cpp:459

    int n_past = 0;
    int64_t n_p_eval = 0, n_eval = 0;
    double t_p_us = 0, t_e_us = 0;
    int n_remain = p.n_predict;                            // -1 = until EOG / context
    b200_ctx * last = p.stages.back();
    // context extension via Self-Extend (cpp/bridge.cpp:509-524), before every llama_decode when grp_attn_n != 1: whole windows of
    // ga_w positions are compressed by ga_n and n_past falls back; the K rows are re-rotated lazily inside the next decode
    int ga_i = 0;
    const auto self_extend = [&]() -> bool {
        while (p.ga_n != 1 && n_past >= ga_i + p.ga_w) {
            const int ib = (p.ga_n * ga_i) / p.ga_w;
            const int bd = (p.ga_w / p.ga_n) * (p.ga_n - 1);
            const int dd = (p.ga_w / p.ga_n) - ib * bd - p.ga_w;
            for (b200_ctx * c : p.stages) {
                if (b200_kv_seq_add(c, ga_i, n_past, ib * bd) != 0 ||
                    b200_kv_seq_div(c, ga_i + ib * bd, ga_i + ib * bd + p.ga_w, p.ga_n) != 0 ||
                    b200_kv_seq_add(c, ga_i + ib * bd + p.ga_w, n_past + ib * bd, dd) != 0) return false;
            }
            n_past -= bd;
            ga_i += p.ga_w / p.ga_n;
        }
        return true;
    };

    // ---- prompt, in chunks of n_batch (cpp/bridge.cpp:549-560, 613-624); pieces are published per chunk
    static const bool pipeline_chunks = [] { const char * e = std::getenv("BOOSTER_B200_PIPELINE_CHUNKS"); return !(e && e[0] == '0'); }();   // A/B switch
    size_t consumed = 0;
    while (consumed < inp.size() && !g_stop[idx].load()) {
        const size_t n = std::min((size_t) p.n_batch, inp.size() - consumed);
        if (!self_extend()) return 1;
        const double t0 = now_us();
        if (p.stages.size() == 1) {
            // one llama_decode of the chunk (cpp/bridge.cpp:549-560): the batched prompt kernels on a pod that sits on one GPU
            if (b200_decode(last, inp.data() + consumed, (int) n, n_past, nullptr) != 0) return 1;
        } else {
            // a pod split over several GPUs: the chunk goes through every stage's batched kernels in pieces of <= 512 tokens
            // (one peer copy of the residual streams per stage boundary); token by token where they cannot run
            for (size_t i0 = 0; i0 < n; ) {
                const size_t nn = std::min((size_t) 512, n - i0);
                bool batch = true;
                for (b200_ctx * c : p.stages) batch = batch && b200_stage_batch_usable(c, (int) nn) == 1;
                if (batch) {
                    b200_ctx * prev = nullptr;
                    for (b200_ctx * c : p.stages) {
                        if (b200_stage_forward_batch(c, inp.data() + consumed + i0, (int) nn, n_past + (int) i0, prev) != 0) return 1;
                        prev = c;
                    }
                } else {
                    for (size_t i = i0; i < i0 + nn; i++)
                        if (run_token(p, inp[consumed + i], n_past + (int) i, n > 1 ? 1 : 0) != 0) return 1;   // llama_decode failed: bridge.cpp:556-558
                }
                i0 += nn;
            }
        }
        // every chunk is a blocking llama_decode in the reference (cpp/bridge.cpp:549-560): synchronise the chain end so
        // that the chunk's time is the prompt's and not the first generated token's. A pod split over several GPUs keeps its
        // prompt chunks IN FLIGHT instead: chunk i+1 runs on the first stage while chunk i runs on the next one (the events of
        // the hand-off order them; what ggml's scheduler does with GGML_SCHED_MAX_COPIES micro-batches,
        // cpp/ggml/src/ggml-backend.c:1029-1030) and only the last chunk is waited for
        const bool last_chunk = consumed + n >= inp.size();
        if (p.stages.size() == 1 || last_chunk || !pipeline_chunks) { if (b200_stage_sync(last) != 0) return 1; }
        const double dt = now_us() - t0;
        if (n > 1) { t_p_us += dt; n_p_eval += (int64_t) n; } else { t_e_us += dt; n_eval += 1; }
        n_past += (int) n;
        std::string pieces;
        for (size_t i = 0; i < n; i++) pieces += p.tok->piece(inp[consumed + i], true);
        { std::lock_guard<std::mutex> lk(g_mu); g_jobs[job].text += pieces; }
        consumed += n;
    }
    if (p.stages.size() > 1 && b200_stage_sync(last) != 0) return 1;   // (a stopped job may have left chunks in flight)

    // ---- generation: sample, then decode the sampled token (cpp/bridge.cpp:586-646).
    //   janus != 0 (what the reference always does): the logits come to the host once per token and the Janus sampler picks;
    //   janus == 0: the standard chain; when that is greedy, the arg-max stays on the device and a pod on one GPU replays
    //   one CUDA graph per token (decode + arg-max + 4-byte read-back).
    // A layer-split pod chains its stages with plain launches and peer copies.
    static const bool graph_ok = [] { const char * e = std::getenv("BOOSTER_B200_BRIDGE_GRAPH"); return !(e && e[0] == '0'); }();   // A/B switch
    static const bool force_greedy = [] { const char * e = std::getenv("BOOSTER_B200_SAMPLER"); return e && std::strcmp(e, "greedy") == 0; }();
    const bool single = p.stages.size() == 1 && graph_ok;
    const bool use_janus = p.jparams.janus != 0 && !force_greedy;
    b200::StandardSampler std_sampler;
    std_sampler.init(p.sparams, seed);
    std_sampler.n_vocab_model = p.n_vocab;
    for (int32_t t : inp) std_sampler.accept(t);           // cpp/bridge.cpp:618: prompt tokens enter the penalty window
    const bool device_argmax = !use_janus && (force_greedy || std_sampler.greedy());
    if (use_janus) p.janus.rng.seed(seed);                 // llama_set_rng_seed(ctx, seed): cpp/bridge.cpp:216-217
    std::vector<int32_t> last_tokens((size_t) p.n_ctx, 0); // cpp/bridge.cpp:437-438: only generated tokens enter
    int32_t id = 0;
    bool have_next = false;                                // device arg-max path: the next id came back with the decode
    float * lg = nullptr;                                  // host path: the logits of the last decoded token (the engine's pinned buffer)
    while (n_remain != 0 && n_past < max_embd && !g_stop[idx].load()) {
        const double t0 = now_us();
        if (device_argmax) {
            if (!have_next) { if (b200_stage_argmax(last, &id) != 0) return 1; }
            have_next = false;
        } else {
            if (!lg) { lg = b200_stage_logits_view(last); if (!lg) return 1; }
            if (use_janus) {
                id = p.janus.sample(lg, last_tokens, inp.size(), (size_t) n_past, (size_t) p.n_predict);
                last_tokens.erase(last_tokens.begin());    // cpp/bridge.cpp:602-603
                last_tokens.push_back(id);
            } else {
                id = std_sampler.sample(lg, p.n_vocab);
            }
            lg = nullptr;
        }
        std_sampler.accept(id);                            // cpp/bridge.cpp:605
        --n_remain;
        { std::lock_guard<std::mutex> lk(g_mu); g_jobs[job].text += p.tok->piece(id, true); }
        if (p.tok->is_eog(id)) break;                      // cpp/bridge.cpp:640
        if (n_remain == 0 || n_past >= max_embd) break;
        // infinite text generation via context shifting (cpp/bridge.cpp:487-507, n_keep = 0: gpt_params' default): keep the
        // first n_keep tokens, drop half of the rest, move the tail down. As in the reference the loop condition above
        // (n_past < n_ctx - 4) ends the job before n_past + 1 can exceed n_ctx, so the branch only documents the behaviour —
        // the operation itself is b200_kv_seq_rm / b200_kv_seq_add, tested against the reference on their own.
        if (p.ga_n != 1) {
            if (!self_extend()) return 1;
        } else if (n_past + 1 > p.n_ctx) {
            const int n_keep = 0, n_left = n_past - n_keep, n_discard = n_left / 2;
            for (b200_ctx * c : p.stages) {
                if (b200_kv_seq_rm(c, n_keep, n_keep + n_discard) != 0 || b200_kv_seq_add(c, n_keep + n_discard, n_past, -n_discard) != 0) return 1;
            }
            n_past -= n_discard;
        }
        if (single && device_argmax) {
            int32_t next = 0;
            if (b200_step_greedy(last, id, n_past, &next) != 0) return 1;
            id = next; have_next = true;
        } else if (single) {
            lg = b200_decode_view(last, id, n_past);       // one graph: state in, forward, logits out (pinned, no second copy)
            if (!lg) return 1;
        } else {
            if (run_token(p, id, n_past, 0) != 0) return 1;
        }
        n_past += 1;
        // single stage: the step is synchronous; layer split: launches are asynchronous and the interval that ends at
        // the NEXT read-back's sync is what one token costs
        t_e_us += now_us() - t0; n_eval += 1;
    }
    // per-token timings, integer milliseconds like the reference (cpp/bridge.cpp:650-655)
    {
        std::lock_guard<std::mutex> lk(g_mu);
        Job & j = g_jobs[job];
        j.prompt_us_per_tok = n_p_eval ? t_p_us / (double) n_p_eval : 0;
        j.gen_us_per_tok    = n_eval ? t_e_us / (double) n_eval : 0;
        j.prompt_eval_ms = (int64_t) (j.prompt_us_per_tok / 1000.0);
        j.timing_ms      = (int64_t) (j.gen_us_per_tok / 1000.0);
    }
    return n_p_eval + n_eval;
}

int64_t doInference(int idx, void * ctx, char * jobID, char * sessionID, char * prompt) {
    try {
        return do_inference_impl(idx, ctx, jobID, sessionID, prompt);
    } catch (const std::exception & e) {
        std::fprintf(stderr, "doInference: %s\n", e.what());
    } catch (...) {
        std::fprintf(stderr, "doInference: unknown failure\n");
    }
    return 0;
}

void stopInference(int idx) {
    if (idx >= 0 && idx < MAX_PODS) g_stop[idx].store(true);
}

const char * status(char * jobID) {
    static thread_local std::string snap;
    if (!jobID) return "";
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_jobs.find(jobID);
    snap = it == g_jobs.end() ? std::string() : it->second.text;
    return snap.c_str();
}

int64_t promptEval(char * jobID) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_jobs.find(jobID ? jobID : "");
    return it == g_jobs.end() ? 0 : it->second.prompt_eval_ms;
}
int64_t getPromptTokenCount(char * jobID) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_jobs.find(jobID ? jobID : "");
    return it == g_jobs.end() ? 0 : it->second.prompt_tokens;
}
int64_t timing(char * jobID) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_jobs.find(jobID ? jobID : "");
    return it == g_jobs.end() ? 0 : it->second.timing_ms;
}
uint32_t getSeed(char * jobID) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_jobs.find(jobID ? jobID : "");
    return it == g_jobs.end() ? 0 : it->second.seed;
}

// additive: µs-resolution per-token timings (the bridge's ms-truncated values read 0 on a B200)
int b200_job_timing_us(const char * jobID, double * prompt_us_per_token, double * gen_us_per_token) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_jobs.find(jobID ? jobID : "");
    if (it == g_jobs.end() || !prompt_us_per_token || !gen_us_per_token) return 1;
    *prompt_us_per_token = it->second.prompt_us_per_tok;
    *gen_us_per_token    = it->second.gen_us_per_tok;
    return 0;
}

}  // extern "C"
