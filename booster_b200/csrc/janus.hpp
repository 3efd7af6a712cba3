// janus.hpp — host-side samplers behind doInference.
//
// The reference's bridge samples EVERY token with its "Janus" sampler (cpp/bridge.cpp:586-596; the standard sampling
// chain is commented out there): per-token repetition scales derived from the token TEXT (cpp/janus.cpp initJanus
// :405-700), an <EOS> boost, language heuristics, a cut of the sorted candidates at a ratio to the top logit and a
// softmax draw from the short list (sample_janus_token :191-331). JanusSampler restates that arithmetic on the host —
// the logits come back from the GPU once per token — so that with the same seed the drop-in library generates the
// same token ids as the reference. Pinned against the reference's own janus.cpp (compiled unmodified into oracle/_ref):
// tests/golden/janus_*.json, tests/test_sampler.py.
//
// Where the reference's arithmetic is undefined, this file is defined and says so:
//   * initJanus indexes a 20-entry table with the token's byte length (cpp/janus.cpp:474-492): tokens of 20 bytes or more
//     (40 for Cyrillic) read past the table. Here the index is clamped to the last entry; warned once per process.
//   * scales[llama_token_eot(model)] is written even when the model has no EOT token (index -1), and the LLaMA-2 branch
//     writes fixed token ids regardless of the vocabulary size (:616-693). Here both are bounds-checked.
//   * candidates whose logits tie exactly may be ordered differently by the reference's std::sort of the whole vocabulary.
//
// StandardSampler is the chain the reference's bridge keeps commented out (cpp/bridge.cpp:598 llama_sampling_sample:
// repetition penalty, top-k, tail-free, typical, top-p, min-p, temperature, or mirostat 1 / 2; cpp/common/sampling.cpp),
// selected with janus = 0: the reference's own code for it is compiled and linked, its bridge just never calls it.
#pragma once
#include <cstdint>
#include <random>
#include <string>
#include <vector>

#include "tokenizer.hpp"

namespace b200 {

struct JanusParams {       // cpp/janus.h:13-19
    int32_t janus = 1;
    int32_t depth = 200;
    float   scale = 0.96f;
    float   hi    = 0.99f;
    float   lo    = 0.96f;
};

struct JanusSampler {
    JanusParams p;
    std::vector<float> scales, types;      // per token id (cpp/janus.cpp ::scales, ::types)
    std::vector<uint8_t> pedantic;         // isPedantic(id), cached (the reference re-derives it from the piece per call)
    std::vector<float> block_max;          // scratch of sample(): the maximum of every block of 64 logits
    std::mt19937 rng;                      // llama_context's sampling rng (cpp/src/llama.cpp:18610 llama_set_rng_seed)
    int32_t n_vocab = 0;

    // initJanus(ctx, params, debug) (cpp/bridge.cpp:196) + llama_set_rng_seed(ctx, seed) (cpp/bridge.cpp:216-217)
    void init(const Tokenizer & tok, const JanusParams & params, uint32_t seed);
    // sample_janus_token (cpp/janus.cpp:191-331). logits[n_vocab] is modified in place, as the reference modifies the
    // context's logits. last_tokens: n_ctx entries, zeros, then the generated tokens (cpp/bridge.cpp:437-438, 602-603).
    int32_t sample(float * logits, const std::vector<int32_t> & last_tokens, size_t prompt_len, size_t pos, size_t max);
};

struct StandardParams {    // llama_sampling_params (cpp/common/sampling.h:23-48) — what initContext stores (cpp/bridge.cpp:763-776)
    float   temp = 0.8f;
    int32_t top_k = 40;
    float   top_p = 0.95f;
    float   min_p = 0.05f;            // not in initContext's signature: the reference's default
    float   tfs_z = 1.0f;             // likewise (1 = off)
    float   typical_p = 1.0f;         // initContext: typical_p > 0 ? typical_p : 1 (cpp/bridge.cpp:773)
    float   penalty_repeat = 1.0f;
    int32_t penalty_last_n = 64;      // 0 = off, < 0 = the whole window
    int32_t n_prev = 64;              // window of accepted tokens the penalties look at
    int32_t min_keep = 0;
    int32_t mirostat = 0;             // 0 off, 1 mirostat, 2 mirostat 2.0
    float   mirostat_tau = 5.0f;
    float   mirostat_eta = 0.1f;
    bool    penalize_nl = false;
    int32_t nl_token = -1;            // llama_token_nl(model): its logit is restored after the penalties unless penalize_nl
};

// llama_sampling_context + llama_sampling_sample / llama_sampling_accept (cpp/common/sampling.cpp:5-47, 271-340, 342-427,
// 429-440) over the samplers of cpp/src/llama-sampling.cpp, restated: same candidate order (the same std::sort /
// std::partial_sort calls at the same places, the bucket pre-sort of top-k > 128 included), the same float arithmetic
// (incl. the three places where the reference's -march build contracts a*b+c into one fma), the same std::mt19937 /
// std::discrete_distribution draws. Pinned token for token against the reference's own chain (oracle/_ref):
// tests/test_sampler.py::test_standard_chain_equals_reference_token_for_token.
struct StandardSampler {
    StandardParams p;
    std::mt19937 rng;                 // llama_sampling_context::rng — the temperature chain's draw
    std::mt19937 ctx_rng;             // llama_context's sampling rng (llama_set_rng_seed) — the mirostat draws
    float mirostat_mu = 0.f;          // value-initialised by llama_sampling_init and never set (upstream uses 2 * tau)
    std::vector<int32_t> prev;        // n_prev zeros, then every accepted token (prompt tokens too: cpp/bridge.cpp:605, 618)
    int32_t n_vocab_model = 0;        // llama_sampling::n_vocab (mirostat's k estimate)
    void init(const StandardParams & params, uint32_t seed) {
        p = params; rng.seed(seed); ctx_rng.seed(seed); mirostat_mu = 0.f;
        prev.assign((size_t) (p.n_prev > 0 ? p.n_prev : 0), 0);
    }
    void accept(int32_t id) { if (!prev.empty()) { prev.erase(prev.begin()); prev.push_back(id); } }
    bool penalties_active() const { return p.penalty_last_n != 0 && p.penalty_repeat != 1.0f && !prev.empty(); }
    // true when the chain reduces to the arg-max of the RAW logits (the bridge then keeps the arg-max on the device)
    bool greedy() const { return !penalties_active() && (p.temp <= 0.f || (p.mirostat == 0 && p.top_k == 1)); }
    int32_t sample(float * logits, int32_t n_vocab);      // the penalties are written into logits
};

}  // namespace b200
