// janus.cpp — see janus.hpp. Host-only, no CUDA.
#include "janus.hpp"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <limits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

namespace b200 {

namespace {

// token types (cpp/janus.h:27-37)
constexpr int LANG_ZERO = 0, LANG_EN = 2, SPACE_EN = 20, LANG_RU = 3, SPACE_RU = 30, LANG_OTHER = 4, SPACE_OTHER = 40;
constexpr int EOS = 2, NL = 13;    // cpp/janus.h:23-24: fixed ids, whatever the model

// tokType (cpp/janus.cpp:724-829): byte-level classification of the piece
int tok_type(const std::string & in) {
    int en = 0, ru = 0, other = 0;
    const size_t n = in.size();
    const bool space = n > 0 && (unsigned char) in[0] == 0x20;
    for (size_t i = 0; i < n; i++) {
        const unsigned char b = (unsigned char) in[i];
        if ((b >= 0x41 && b <= 0x5A) || (b >= 0x61 && b <= 0x7A)) { en++; continue; }
        if (b < 0x80) continue;
        if (b == 0xD0 && i + 1 < n) {
            i++;
            const unsigned char c = (unsigned char) in[i];
            if ((c >= 0x90 && c <= 0xBF) || c == 0x81) ru++; else other++;
            continue;
        }
        if (b == 0xD1 && i + 1 < n) {
            i++;
            const unsigned char c = (unsigned char) in[i];
            if ((c >= 0x80 && c <= 0x8F) || c == 0x91) ru++; else other++;
            continue;
        }
        if (b >= 0xC3 && b < 0xE3) { i++; other++; continue; }
        if (b >= 0xE3 && b < 0xF0) { i += 2; other++; continue; }
        if (b >= 0xF0) { i += 3; continue; }
    }
    if (space) {
        if (other) return SPACE_OTHER;
        if (en) return SPACE_EN;
        if (ru) return SPACE_RU;
    }
    if (other) return LANG_OTHER;
    if (en) return LANG_EN;
    if (ru) return LANG_RU;
    return LANG_ZERO;
}

// isLower (cpp/janus.cpp:832-865)
bool is_lower(const std::string & in) {
    if (in.empty()) return false;
    const unsigned char b0 = (unsigned char) in[0];
    if (b0 >= 0x61 && b0 <= 0x7A) return true;
    if (in.size() >= 2) {
        const unsigned char b1 = (unsigned char) in[1];
        if (b0 == 0xD0 && b1 >= 0xB0 && b1 <= 0xBF) return true;
        if (b0 == 0xD1 && ((b1 >= 0x80 && b1 <= 0x8F) || b1 == 0x91)) return true;
    }
    return false;
}

// isPedantic (cpp/janus.cpp:378-401)
bool is_pedantic(const std::string & token) {
    char * end = nullptr;
    strtol(token.c_str(), &end, 10);
    if (*end == 0) return true;            // numbers (and the empty piece)
    if (token == " *" || token == " =" || token == " -" || token == " +") return true;
    if (token == "{" || token == "}" || token == "[" || token == "]") return true;
    if (token == " {" || token == " }" || token == " [" || token == " ]") return true;
    if (token == "<|end_of_text|>" || token == "```") return true;
    return false;
}

// The two scans below touch all 128 k logits of a token; the library is built for baseline x86-64, so they are compiled a second
// time for AVX2 and picked at load time (GNU function multi-versioning). Same comparisons, same results: only wider vectors.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define B200_WIDE __attribute__((target_clones("avx2", "default")))
#else
#define B200_WIDE
#endif
// The maximum of x[0..n) (n >= 1) — order-independent, so exact — in 8-float vectors (GNU vector extensions: one vmaxps per 8
// logits in the AVX2 clone, two maxps in the baseline one; a scalar running maximum is a 4-cycle dependency per element and
// costs more than the GPU spends on the token), and on the way the maximum of every block of 64 logits (bm[(n + 63) / 64]): the
// scan for the short list then only opens the blocks that can hold a candidate instead of reading all 128 k logits a second time
typedef float vf8 __attribute__((vector_size(32)));
constexpr int32_t BM_BLOCK = 64;
B200_WIDE float block_maxima(const float * x, int32_t n, float * bm) {
    const int32_t nb = n / BM_BLOCK;
    const float ninf = -std::numeric_limits<float>::infinity();   // every lane starts below everything: a NaN logit never wins and
    const vf8 vinf = { ninf, ninf, ninf, ninf, ninf, ninf, ninf, ninf };   // never sticks (x > m is false for it), as in a scalar scan
    float r = ninf;
    for (int32_t b = 0; b < nb; b++) {
        const float * q = x + (size_t) b * BM_BLOCK;
        vf8 m0 = vinf, m1 = vinf, a;
        for (int j = 0; j < BM_BLOCK; j += 16) {
            std::memcpy(&a, q + j, 32);     m0 = a > m0 ? a : m0;
            std::memcpy(&a, q + j + 8, 32); m1 = a > m1 ? a : m1;
        }
        m0 = m1 > m0 ? m1 : m0;
        float t[8];
        std::memcpy(t, &m0, 32);
        float v = t[0];
        for (int j = 1; j < 8; j++) v = t[j] > v ? t[j] : v;
        bm[b] = v;
        r = v > r ? v : r;
    }
    if (nb * BM_BLOCK < n) {
        float v = ninf;
        for (int32_t i = nb * BM_BLOCK; i < n; i++) v = x[i] > v ? x[i] : v;
        bm[nb] = v;
        r = v > r ? v : r;
    }
    return r;
}
// first index >= from with x[index] >= thr, or n (blocks of 64 tested branch-free, then the hit located)
B200_WIDE int32_t first_at_least(const float * x, int32_t n, float thr, int32_t from) {
    int32_t i = from;
    for (; i < n && (i & 63); i++) if (x[i] >= thr) return i;
    for (; i + 64 <= n; i += 64) {
        int any = 0;
        for (int j = 0; j < 64; j++) any |= x[i + j] >= thr;
        if (any) { for (int j = 0; j < 64; j++) if (x[i + j] >= thr) return i + j; }
    }
    for (; i < n; i++) if (x[i] >= thr) return i;
    return n;
}

// first index >= from with x[index] > thr (strict), or n
B200_WIDE int32_t first_above(const float * x, int32_t n, float thr, int32_t from) {
    int32_t i = from;
    for (; i < n && (i & 63); i++) if (x[i] > thr) return i;
    for (; i + 64 <= n; i += 64) {
        int any = 0;
        for (int j = 0; j < 64; j++) any |= x[i + j] > thr;
        if (any) { for (int j = 0; j < 64; j++) if (x[i + j] > thr) return i + j; }
    }
    for (; i < n; i++) if (x[i] > thr) return i;
    return n;
}

}  // namespace

void JanusSampler::init(const Tokenizer & tok, const JanusParams & params, uint32_t seed) {
    p = params;
    rng.seed(seed);
    n_vocab = tok.n_vocab();
    scales.assign((size_t) n_vocab, 0.f);
    types.assign((size_t) n_vocab, 0.f);
    pedantic.assign((size_t) n_vocab, 0);
    // safe defaults (cpp/janus.cpp:437-441) — they change the caller's parameters, as the reference does
    if (p.depth <= 0) p.depth = 200;
    if (p.scale <= 0.0 || p.scale > 1.0) p.scale = 0.97f;
    if (p.hi <= 0.0 || p.hi > 1.0) p.hi = 0.99f;
    if (p.lo <= 0.0 || p.lo > 1.0) p.lo = 0.96f;
    const float scale = p.scale;
    static const float probes[20] = { 0.20f, 0.22f, 0.25f, 0.28f, 0.30f, 0.32f, 0.33f, 0.35f, 0.36f, 0.38f,
                                      0.40f, 0.42f, 0.44f, 0.45f, 0.46f, 0.48f, 0.50f, 0.52f, 0.53f, 0.55f };
    static bool warned = false;
    auto probe = [&](size_t i) {
        if (i >= 20) {
            if (!warned) { warned = true; std::fprintf(stderr, "booster_b200: Janus: token longer than the reference's 20-entry length table; clamped (the reference reads out of bounds there)\n"); }
            i = 19;
        }
        return probes[i];
    };
    std::vector<std::string> piece((size_t) n_vocab);
    for (int32_t id = 0; id < n_vocab; id++) piece[(size_t) id] = tok.piece(id, true);   // llama_token_to_piece(ctx, id): special = true
    for (int32_t id = 0; id < n_vocab; id++) {
        const std::string & s = piece[(size_t) id];
        const int type = tok_type(s);
        const bool lower = is_lower(s);
        const size_t len = s.size();
        types[(size_t) id] = (float) type;
        pedantic[(size_t) id] = is_pedantic(s) ? 1 : 0;
        // every expression below is evaluated in double and stored as float, as in the reference (1.0 is a double literal)
        if (pedantic[(size_t) id]) { scales[(size_t) id] = (float) (1.0 - (1.0 - scale) * 0.20); continue; }
        if (type == LANG_RU && lower) { scales[(size_t) id] = (float) (1.0 - (1.0 - scale) * probe(len / 2)); continue; }
        if (type == LANG_EN && lower) { scales[(size_t) id] = (float) (1.0 - (1.0 - scale) * probe(len)); continue; }
        scales[(size_t) id] = scale;
    }
    auto set = [&](int64_t id, double v) { if (id >= 0 && id < n_vocab) scales[(size_t) id] = (float) v; };
    set(0, 1.0);
    set(tok.eos(), scale);
    set(tok.eot(), scale);
    // "LLaMA v2/v3 and Mistral": llama_model_desc starts with the architecture name, "llama" for every model on this path
    if (n_vocab > 128000) {                 // LLaMA-3 (cpp/janus.cpp:536-624): by piece text and id range
        for (int32_t id = 0; id < n_vocab; id++) {
            const std::string & t = piece[(size_t) id];
            const int type = (int) types[(size_t) id];
            if (t == "\n" || t == "\n\n") { set(id, 1.0 - (1.0 - scale) * 0.10); continue; }
            if (t == "  " || t == "    ") { set(id, 1.0 - (1.0 - scale) * 0.20); continue; }
            if (t == " " || t == "," || t == ".") { set(id, 1.0 - (1.0 - scale) * 0.10); continue; }
            if (t == " \xE2\x80\x94" || t == "-" || t == ":" || t == ";") { set(id, 1.0 - (1.0 - scale) * 0.30); continue; }
            if (t == " (" || t == ")." || t == " )" || t == ")" || t == "(") { set(id, 1.0 - (1.0 - scale) * 0.30); continue; }
            if (id < 20000 && type == SPACE_RU) { set(id, 1.0 - (1.0 - scale) * 0.30); continue; }
            if (id >= 20000 && id < 35000 && type == SPACE_RU) { set(id, 1.0 - (1.0 - scale) * 0.40); continue; }
            if (id >= 35000 && id < 50000 && type == SPACE_RU) { set(id, 1.0 - (1.0 - scale) * 0.50); continue; }
            if (id < 500 && type == SPACE_EN) { set(id, 1.0 - (1.0 - scale) * 0.30); continue; }
            if (id >= 500 && id < 800 && type == SPACE_EN) { set(id, 1.0 - (1.0 - scale) * 0.40); continue; }
            if (id >= 800 && id < 1100 && type == SPACE_EN) { set(id, 1.0 - (1.0 - scale) * 0.50); continue; }
        }
    } else {                                // LLaMA-2 (cpp/janus.cpp:626-693): fixed token ids
        set(0, 1.0);
        set(EOS, scale);
        set(NL, 1.0 - (1.0 - scale) * 0.10);
        static const struct { int id; double f; } fixed[] = {
            {259, 0.20}, {268, 0.20}, {29871, 0.10}, {29892, 0.10}, {29889, 0.20}, {813, 0.30}, {29899, 0.30}, {29901, 0.30}, {29936, 0.30},
            {313, 0.30}, {467, 0.30}, {1723, 0.30}, {29897, 0.30}, {29898, 0.30},
            {490, 0.30}, {531, 0.30}, {606, 0.30}, {614, 0.30}, {665, 0.35}, {733, 0.35}, {863, 0.35}, {1077, 0.40}, {1097, 0.40}, {1186, 0.40},
            {1447, 0.45}, {1538, 0.45}, {1604, 0.45}, {1685, 0.45}, {4281, 0.50}, {857, 0.50}, {939, 0.50}, {1651, 0.50},
            {263, 0.30}, {278, 0.30}, {297, 0.30}, {304, 0.30}, {310, 0.30}, {322, 0.30}, {363, 0.35}, {372, 0.35}, {373, 0.35}, {385, 0.35},
            {393, 0.35}, {408, 0.35}, {411, 0.35}, {470, 0.40}, {472, 0.40}, {526, 0.40}, {319, 0.50},
        };
        for (const auto & f : fixed) set(f.id, 1.0 - (1.0 - scale) * f.f);
    }
}

int32_t JanusSampler::sample(float * logits, const std::vector<int32_t> & last_tokens, size_t prompt_len, size_t pos, size_t max) {
    const size_t ctx_size = last_tokens.size();
    const int32_t last_token = last_tokens[ctx_size - 1];
    const float last_type = types[(size_t) last_token];
    // boost <EOS> when we are closer to the limit (cpp/janus.cpp:235): double arithmetic, stored back as float
    if (EOS < n_vocab) logits[EOS] = (float) ((double) logits[EOS] * (1.0 + std::log(1.0 + (double) ((float) (pos - prompt_len) / (float) max)) * 0.05));
    // pessimization of repeated tokens among the generated ones (cpp/janus.cpp:240-267)
    const size_t depth = std::min((size_t) p.depth, pos - prompt_len);
    for (size_t i = 0; i < depth; i++) {
        const int32_t id = last_tokens[ctx_size - 1 - i];
        const float cur_type = types[(size_t) id];
        if ((last_type == SPACE_RU || last_type == LANG_RU) && cur_type == LANG_RU) {
            logits[id] = (float) ((double) logits[id] * (1.0 - (1.0 - (double) scales[(size_t) id]) * 0.20));
            continue;
        }
        logits[id] *= scales[(size_t) id];
    }
    // double down incompatible tokens (cpp/janus.cpp:271-285)
    if (last_type == SPACE_RU || last_type == LANG_RU) {
        for (int32_t id = 0; id < n_vocab; id++) {
            const float t = types[(size_t) id];
            if (t == LANG_EN || t == LANG_OTHER) logits[id] = (float) ((double) logits[id] * 0.5);
        }
    }
    // The reference sorts all candidates by logit (descending) and cuts the list at the first one whose ratio to the top
    // logit is below the cutoff (cpp/janus.cpp:289-324). For a positive top logit the ratio falls along the sorted order,
    // so the short list is exactly the candidates whose ratio is not below the cutoff — found in one pass, no full sort.
    block_max.resize((size_t) (n_vocab + BM_BLOCK - 1) / BM_BLOCK);
    float top_logit = block_maxima(logits, n_vocab, block_max.data());
    struct Cand { int32_t id; float logit; float p; };
    std::vector<Cand> cand;
    const auto by_logit = [](const Cand & a, const Cand & b) { return a.logit > b.logit; };
    int32_t top;
    float cutoff;
    const auto cutoff_for = [&](int32_t t) {                // cpp/janus.cpp:306-312: the pedantic / single-language top token takes `hi`
        const float top_type = types[(size_t) t];
        return (pedantic[(size_t) t] || top_type == LANG_RU || top_type == LANG_EN) ? p.hi : p.lo;
    };
    if (top_logit > 0.f && p.lo > 0.f && p.hi > 0.f) {
        // ONE more scan finds both the top token (the first logit equal to the maximum) and the short list. The cutoff depends on
        // the top token, so the scan uses the smaller of the two possible cutoffs and the exact test follows: the reference's
        // division x / top < cutoff; the multiplication with a safety margin only keeps 128 k divisions per token out of the
        // loop (x / top < cutoff certainly holds when x < top * cutoff * (1 - 2^-10)).
        const float guard = top_logit * std::min(p.lo, p.hi) * 0.999f;
        top = -1;
        for (size_t b = 0; b < block_max.size(); b++) {
            if (!(block_max[b] >= guard)) continue;
            const int32_t end = std::min(n_vocab, (int32_t) (b + 1) * BM_BLOCK);
            for (int32_t id = (int32_t) b * BM_BLOCK; id < end; id++) {
                if (!(logits[id] >= guard)) continue;
                if (top < 0 && logits[id] == top_logit) top = id;
                cand.push_back({id, logits[id], 0.f});
            }
        }
        cutoff = cutoff_for(top);
        size_t kept = 0;
        for (const Cand & c : cand) if (!(c.logit / top_logit < cutoff)) cand[kept++] = c;
        cand.resize(kept);
        std::sort(cand.begin(), cand.end(), by_logit);
    } else {
        // (init() clamps lo and hi into (0, 1], so this is the top-logit <= 0 / NaN case) the two-step form: top token first
        top = first_at_least(logits, n_vocab, top_logit, 0);
        cutoff = top < n_vocab ? cutoff_for(top) : p.lo;
        if (top_logit > 0.f && cutoff > 0.f) {
            const float guard = top_logit * cutoff * 0.999f;
            for (int32_t id = first_at_least(logits, n_vocab, guard, 0); id < n_vocab; id = first_at_least(logits, n_vocab, guard, id + 1))
                if (!(logits[id] / top_logit < cutoff)) cand.push_back({id, logits[id], 0.f});
            std::sort(cand.begin(), cand.end(), by_logit);
        } else {
            // the ratio does not fall along the order; walk the fully sorted list as the reference does (the whole vocabulary
            // survives when every logit is negative)
            cand.reserve((size_t) n_vocab);
            for (int32_t id = 0; id < n_vocab; id++) cand.push_back({id, logits[id], 0.f});
            std::sort(cand.begin(), cand.end(), by_logit);
            for (size_t i = 1; i < cand.size(); i++) if (cand[i].logit / cand[0].logit < cutoff) { cand.resize(i); break; }
        }
    }
    // llama_sample_token: softmax over the short list, then one draw (cpp/src/llama-sampling.cpp:32-59, 610-631)
    const float max_l = cand[0].logit;
    float cum = 0.f;
    for (auto & c : cand) { c.p = expf(c.logit - max_l); cum += c.p; }
    std::vector<float> probs;
    probs.reserve(cand.size());
    for (auto & c : cand) { c.p /= cum; probs.push_back(c.p); }
    std::discrete_distribution<> dist(probs.begin(), probs.end());
    return cand[(size_t) dist(rng)].id;
}

// ------------------------------------------------------------------------------------------------------------
// the standard chain: llama_sampling_sample (cpp/common/sampling.cpp:271-340) over cpp/src/llama-sampling.cpp
// ------------------------------------------------------------------------------------------------------------
namespace {

struct Cand { int32_t id; float logit; float p; };
struct Cands {                          // llama_token_data_array: the live prefix of `v` and whether it is sorted by logit
    std::vector<Cand> v;
    bool sorted = false;
};
const auto by_logit_desc = [](const Cand & a, const Cand & b) { return a.logit > b.logit; };

// llama_sample_softmax_impl (:32-59)
void softmax(Cands & c) {
    if (!c.sorted) { std::sort(c.v.begin(), c.v.end(), by_logit_desc); c.sorted = true; }
    const float top = c.v[0].logit;
    float total = 0.0f;
    for (Cand & x : c.v) { x.p = expf(x.logit - top); total += x.p; }
    for (Cand & x : c.v) x.p /= total;
}

// llama_sample_top_k_impl (:61-140). k <= 128: partial sort. Larger k: 128 buckets over logits -10..10, the buckets above the
// one that holds the k-th candidate fully sorted, that one partially — restated because the order of tied logits (and with it
// a later draw) depends on it. The bucket index is one fused multiply-add in the reference's build.
void top_k(Cands & c, int32_t k, size_t min_keep) {
    const int n = (int) c.v.size();
    if (k <= 0) k = n;
    k = std::max(k, (int) min_keep);
    k = std::min(k, n);
    if (!c.sorted) {
        if (k <= 128) {
            std::partial_sort(c.v.begin(), c.v.begin() + k, c.v.end(), by_logit_desc);
        } else {
            constexpr int NB = 128;
            constexpr float lo = -10.0f, hi = 10.0f, scale = NB / (hi - lo), inter = -lo * scale;
            std::vector<int> which((size_t) n), count(NB, 0);
            for (int i = 0; i < n; i++) {
                int b = (int) fmaf(scale, c.v[(size_t) i].logit, inter);
                b = std::max(0, std::min(NB - 1, b));
                which[(size_t) i] = b; count[(size_t) b]++;
            }
            int have = 0, cut = NB - 1;
            for (; cut >= 0; cut--) { have += count[(size_t) cut]; if (have >= k) break; }
            std::vector<Cand> kept((size_t) have);
            std::vector<size_t> fill((size_t) NB, 0);         // write cursor of every kept bucket, highest bucket first
            { size_t at = 0; for (int b = NB - 1; b >= cut; b--) { fill[(size_t) b] = at; at += (size_t) count[(size_t) b]; } }
            for (int i = 0; i < n; i++) { const int b = which[(size_t) i]; if (b >= cut) kept[fill[(size_t) b]++] = c.v[(size_t) i]; }
            size_t at = 0; int done = 0;
            for (int b = NB - 1; b > cut; b--) {
                std::sort(kept.begin() + (long) at, kept.begin() + (long) (at + (size_t) count[(size_t) b]), by_logit_desc);
                at += (size_t) count[(size_t) b]; done += count[(size_t) b];
            }
            std::partial_sort(kept.begin() + (long) at, kept.begin() + (long) at + (k - done), kept.begin() + (long) (at + (size_t) count[(size_t) cut]), by_logit_desc);
            std::copy(kept.begin(), kept.begin() + k, c.v.begin());
        }
        c.sorted = true;
    }
    c.v.resize((size_t) k);
}

// llama_sample_top_p_impl (:142-172)
void top_p(Cands & c, float p, size_t min_keep) {
    if (p >= 1.0f) return;
    softmax(c);
    float cum = 0.0f;
    size_t last = c.v.size();
    for (size_t i = 0; i < c.v.size(); i++) {
        cum += c.v[i].p;
        if (cum >= p && i + 1 >= min_keep) { last = i + 1; break; }
    }
    c.v.resize(last);
}

// llama_sample_min_p_impl (:174-233): on LOGITS (logit >= top + logf(p)), unsorted filter first when it keeps enough
void min_p(Cands & c, float p, size_t min_keep) {
    if (p <= 0.0f || c.v.empty()) return;
    if (!c.sorted) {
        float top = -FLT_MAX;
        for (const Cand & x : c.v) top = std::max(top, x.logit);
        const float floor_logit = top + logf(p);
        std::vector<Cand> keep;
        for (const Cand & x : c.v) if (x.logit >= floor_logit) keep.push_back(x);
        if (keep.size() >= min_keep) { c.v.swap(keep); return; }
        std::sort(c.v.begin(), c.v.end(), by_logit_desc);
        c.sorted = true;
    }
    const float floor_logit = c.v[0].logit + logf(p);
    size_t i = 1;
    for (; i < c.v.size(); i++) if (c.v[i].logit < floor_logit && i >= min_keep) break;
    c.v.resize(i);
}

// llama_sample_tail_free_impl (:235-292)
void tail_free(Cands & c, float z, size_t min_keep) {
    if (z >= 1.0f || c.v.size() <= 2) return;
    softmax(c);
    const size_t n = c.v.size();
    std::vector<float> d1(n - 1), d2(n - 2);
    for (size_t i = 0; i + 1 < n; i++) d1[i] = c.v[i].p - c.v[i + 1].p;
    for (size_t i = 0; i + 2 < n; i++) d2[i] = std::abs(d1[i] - d1[i + 1]);
    float total = 0.0f;
    for (float x : d2) total += x;
    if (total > 1e-6f) { for (float & x : d2) x /= total; }
    else               { for (float & x : d2) x = 1.0f / d2.size(); }
    float cum = 0.0f;
    size_t last = n;
    for (size_t i = 0; i < d2.size(); i++) {
        cum += d2[i];
        if (cum > z && i >= min_keep) { last = i; break; }
    }
    c.v.resize(last);
}

// llama_sample_typical_impl (:294-356): candidates ordered by |surprise - entropy|; leaves the array unsorted
void typical(Cands & c, float p, size_t min_keep) {
    if (p >= 1.0f) return;
    softmax(c);
    const size_t n = c.v.size();
    float entropy = 0.0f;
    for (const Cand & x : c.v) entropy += -x.p * logf(x.p);
    std::vector<float> dist(n);
    for (size_t i = 0; i < n; i++) dist[i] = fabsf(-logf(c.v[i].p) - entropy);
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = i;
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return dist[a] < dist[b]; });
    float cum = 0.0f;
    size_t last = n;
    for (size_t i = 0; i < n; i++) {
        cum += c.v[order[i]].p;
        if (cum > p && i >= min_keep - 1) { last = i + 1; break; }      // (min_keep - 1 in size_t, as in the reference)
    }
    std::vector<Cand> keep;
    keep.reserve(last);
    for (size_t i = 0; i < last; i++) keep.push_back(c.v[order[i]]);
    c.v.swap(keep);
    c.sorted = false;
}

// llama_sample_token_with_rng_impl (:610-631)
int32_t draw(Cands & c, std::mt19937 & rng) {
    softmax(c);
    std::vector<float> probs;
    probs.reserve(c.v.size());
    for (const Cand & x : c.v) probs.push_back(x.p);
    std::discrete_distribution<> dist(probs.begin(), probs.end());
    return c.v[(size_t) dist(rng)].id;
}

float surprise_of(const Cands & c, int32_t id) {
    size_t i = 0;
    while (i < c.v.size() && c.v[i].id != id) i++;
    return -log2f(c.v[i].p);
}

}  // namespace

int32_t StandardSampler::sample(float * logits, int32_t n_vocab) {
    // llama_sampling_prepare_impl (cpp/common/sampling.cpp:342-409): every token a candidate, penalties over the tail of prev.
    // The penalised values are written into `logits` (the caller's buffer, like the Janus sampler does): the candidates are
    // then simply (id, logits[id]).
    const int window = p.penalty_last_n < 0 ? p.n_prev : p.penalty_last_n;
    const int used = std::min((int) prev.size(), window);
    if (used) {
        const bool have_nl = p.nl_token >= 0 && p.nl_token < n_vocab;
        const float nl_logit = have_nl ? logits[p.nl_token] : 0.0f;
        // llama_sample_repetition_penalties_impl (cpp/src/llama-sampling.cpp:437-482) with frequency / presence penalties 0:
        // every DISTINCT token of the window once
        if (p.penalty_repeat != 1.0f) {
            std::vector<int32_t> distinct;
            for (size_t i = prev.size() - (size_t) used; i < prev.size(); i++) {
                const int32_t id = prev[i];
                if (id < 0 || id >= n_vocab || std::find(distinct.begin(), distinct.end(), id) != distinct.end()) continue;
                distinct.push_back(id);
                if (logits[id] <= 0) logits[id] *= p.penalty_repeat; else logits[id] /= p.penalty_repeat;
            }
        }
        if (!p.penalize_nl && have_nl) logits[p.nl_token] = nl_logit;
    }
    Cands c;
    // top-k of at most 128 candidates out of the vocabulary is std::partial_sort in the reference (llama_sample_top_k_impl,
    // cpp/src/llama-sampling.cpp:61-140), i.e. libstdc++'s heap select: a heap of the first k candidates, then every later
    // candidate that beats the heap's smallest replaces it (__pop_heap into the heap), then sort_heap. The same heap operations
    // in the same order give the same k candidates in the same order (ties included) without materialising 128 k candidates:
    // the scan for "beats the smallest" runs over the logits themselves, and std::pop_heap over k + 1 slots with the newcomer in
    // the last one IS that replacement step.
    const size_t min_keep_q = (size_t) std::max(1, p.min_keep);
    int k_eff = p.top_k <= 0 ? n_vocab : p.top_k;
    k_eff = std::min(std::max(k_eff, (int) min_keep_q), n_vocab);
    if (p.temp > 0.0f && p.mirostat == 0 && k_eff <= 128 && k_eff < n_vocab) {
        std::vector<Cand> & h = c.v;
        h.resize((size_t) k_eff + 1);
        for (int32_t id = 0; id < k_eff; id++) h[(size_t) id] = {id, logits[id], 0.0f};
        std::make_heap(h.begin(), h.begin() + k_eff, by_logit_desc);
        for (int32_t id = first_above(logits, n_vocab, h[0].logit, k_eff); id < n_vocab; id = first_above(logits, n_vocab, h[0].logit, id + 1)) {
            h[(size_t) k_eff] = {id, logits[id], 0.0f};
            std::pop_heap(h.begin(), h.end(), by_logit_desc);
        }
        std::sort_heap(h.begin(), h.begin() + k_eff, by_logit_desc);
        h.resize((size_t) k_eff);
        c.sorted = true;
        tail_free(c, p.tfs_z, min_keep_q);
        typical(c, p.typical_p, min_keep_q);
        top_p(c, p.top_p, min_keep_q);
        min_p(c, p.min_p, min_keep_q);
        for (Cand & x : c.v) x.logit /= p.temp;
        return draw(c, rng);
    }
    c.v.resize((size_t) n_vocab);
    for (int32_t id = 0; id < n_vocab; id++) c.v[(size_t) id] = {id, logits[id], 0.0f};
    // llama_sampling_sample_impl (cpp/common/sampling.cpp:271-340)
    if (p.temp < 0.0f) { softmax(c); return c.v[0].id; }
    if (p.temp == 0.0f) {
        // llama_sample_token_greedy_impl: std::max_element, the FIRST of equal maxima
        return std::max_element(c.v.begin(), c.v.end(), [](const Cand & a, const Cand & b) { return a.logit < b.logit; })->id;
    }
    if (p.mirostat == 1) {
        // llama_sample_token_mirostat_impl (:507-550), m = 100
        for (Cand & x : c.v) x.logit /= p.temp;
        softmax(c);
        const int m = 100;
        float sum_tb = 0.0f, sum_tt = 0.0f;
        for (size_t i = 0; i < (size_t) (m - 1) && i < c.v.size() - 1; i++) {
            const float t_i = logf((float) (i + 2) / (float) (i + 1));
            const float b_i = logf(c.v[i].p / c.v[i + 1].p);
            sum_tb = fmaf(t_i, b_i, sum_tb);
            sum_tt = fmaf(t_i, t_i, sum_tt);
        }
        const float s_hat = sum_tb / sum_tt;
        const float eps_hat = s_hat - 1;
        const float k = powf((eps_hat * powf(2, mirostat_mu)) / (1 - powf((float) n_vocab_model, -eps_hat)), 1 / s_hat);
        top_k(c, (int) k, 1);
        const int32_t id = draw(c, ctx_rng);
        const float e = surprise_of(c, id) - p.mirostat_tau;
        mirostat_mu = fmaf(-p.mirostat_eta, e, mirostat_mu);
        return id;
    }
    if (p.mirostat == 2) {
        // llama_sample_token_mirostat_v2_impl (:552-592)
        for (Cand & x : c.v) x.logit /= p.temp;
        softmax(c);
        size_t keep = 0;
        while (keep < c.v.size() && !(-log2f(c.v[keep].p) > mirostat_mu)) keep++;
        c.v.resize(std::max<size_t>(1, keep));
        softmax(c);
        const int32_t id = draw(c, ctx_rng);
        const float e = surprise_of(c, id) - p.mirostat_tau;
        mirostat_mu = fmaf(-p.mirostat_eta, e, mirostat_mu);
        return id;
    }
    // sampler_queue (cpp/common/sampling.cpp:231-269), default order k f y p m t
    const size_t min_keep = (size_t) std::max(1, p.min_keep);
    top_k(c, p.top_k, min_keep);
    tail_free(c, p.tfs_z, min_keep);
    typical(c, p.typical_p, min_keep);
    top_p(c, p.top_p, min_keep);
    min_p(c, p.min_p, min_keep);
    for (Cand & x : c.v) x.logit /= p.temp;
    return draw(c, rng);
}

}  // namespace b200

// ------------------------------------------------------------------------------------------------------------
// C-ABI of the samplers alone (include/booster_b200.h): the parity tests feed them the REFERENCE's logits step by step
// and compare the ids with the reference's own sample_janus_token — no GPU involved
// ------------------------------------------------------------------------------------------------------------
#include "../../include/booster_b200.h"

struct b200_sampler {
    std::unique_ptr<b200::Tokenizer> tok;
    b200::JanusSampler janus;
    b200::StandardSampler standard;
    b200::StandardParams sp;
    bool use_janus = true;
    std::vector<int32_t> last_tokens;
    size_t n_prompt = 0;
};

extern "C" b200_sampler * b200_sampler_new(const char * gguf_path, int n_ctx, int32_t janus, int32_t depth, float scale, float hi, float lo,
                                           float temperature, int top_k, float top_p, float repetition_penalty, int penalty_last_n) {
    try {
        if (!gguf_path || n_ctx <= 0) return nullptr;
        std::string err;
        auto s = std::make_unique<b200_sampler>();
        s->tok = b200::make_tokenizer(gguf_path, err);
        if (!s->tok) return nullptr;
        b200::JanusParams jp; jp.janus = janus; jp.depth = depth; jp.scale = scale; jp.hi = hi; jp.lo = lo;
        s->sp.temp = temperature; s->sp.top_k = top_k; s->sp.top_p = top_p; s->sp.penalty_repeat = repetition_penalty; s->sp.penalty_last_n = penalty_last_n;
        s->sp.nl_token = s->tok->linefeed();
        s->use_janus = janus != 0;
        if (s->use_janus) s->janus.init(*s->tok, jp, 0);
        s->standard.init(s->sp, 0);
        s->standard.n_vocab_model = s->tok->n_vocab();
        s->last_tokens.assign((size_t) n_ctx, 0);
        return s.release();
    } catch (const std::exception &) { return nullptr; }
}
// the parameters of the standard chain that b200_sampler_new's (initContext-shaped) signature does not carry; takes effect at
// the next b200_sampler_reset. typical_p <= 0 means 1 (off), as in initContext (cpp/bridge.cpp:773).
extern "C" void b200_sampler_set_standard(b200_sampler * s, int32_t mirostat, float mirostat_tau, float mirostat_eta, float typical_p,
                                          float tfs_z, float min_p) {
    if (!s) return;
    s->sp.mirostat = mirostat; s->sp.mirostat_tau = mirostat_tau; s->sp.mirostat_eta = mirostat_eta;
    s->sp.typical_p = typical_p > 0 ? typical_p : 1.0f; s->sp.tfs_z = tfs_z; s->sp.min_p = min_p;
}
extern "C" void b200_sampler_free(b200_sampler * s) { delete s; }
// a new job: the prompt's ids (accepted into the standard chain's penalty window like cpp/bridge.cpp:618; Janus only needs their
// count) and the rng seed
extern "C" void b200_sampler_reset(b200_sampler * s, const int32_t * prompt, int32_t n_prompt, uint32_t seed) {
    if (!s || (n_prompt > 0 && !prompt)) return;
    std::fill(s->last_tokens.begin(), s->last_tokens.end(), 0);
    s->n_prompt = (size_t) (n_prompt > 0 ? n_prompt : 0);
    s->janus.rng.seed(seed);
    s->standard.init(s->sp, seed);
    for (int32_t i = 0; i < n_prompt; i++) s->standard.accept(prompt[i]);
}
// one token from logits[n_vocab] (modified in place by Janus, as the reference modifies the context's logits) at position
// pos = tokens decoded so far; n_predict as passed to initContext
extern "C" int32_t b200_sampler_sample(b200_sampler * s, float * logits, int32_t pos, int32_t n_predict) {
    if (!s || !logits) return -1;
    int32_t id;
    if (s->use_janus) {
        id = s->janus.sample(logits, s->last_tokens, s->n_prompt, (size_t) pos, (size_t) n_predict);
        s->last_tokens.erase(s->last_tokens.begin());
        s->last_tokens.push_back(id);
    } else {
        id = s->standard.sample(logits, s->tok->n_vocab());
    }
    s->standard.accept(id);      // cpp/bridge.cpp:605: accepted whichever sampler drew it
    return id;
}
