// tokenizer.cpp — see tokenizer.hpp.
#include "tokenizer.hpp"
#include "gguf.hpp"

#include <cstdlib>

namespace b200 {

namespace {

struct IdTokenizer final : Tokenizer {
    int32_t n_vocab = 0;
    int32_t eos = -1, eot = -1;
    bool tokenize(const std::string & text, std::vector<int32_t> & out) const override {
        out.clear();
        size_t i = 0;
        while (i < text.size()) {
            while (i < text.size() && (text[i] == ' ' || text[i] == '\n' || text[i] == '\t' || text[i] == ',')) i++;
            if (i >= text.size()) break;
            char * end = nullptr;
            const long v = std::strtol(text.c_str() + i, &end, 10);
            if (end == text.c_str() + i) return false;            // not a number
            if (v < 0 || v >= n_vocab) return false;
            out.push_back((int32_t) v);
            i = (size_t) (end - text.c_str());
        }
        return true;
    }
    std::string piece(int32_t id) const override { return std::to_string(id) + " "; }
    bool is_eog(int32_t id) const override { return id >= 0 && (id == eos || id == eot); }
};

}  // namespace

std::unique_ptr<Tokenizer> make_tokenizer(const std::string & gguf_path, std::string & err) {
    gguf_file g;
    err = g.open(gguf_path);
    if (!err.empty()) return nullptr;
    const std::string model = g.get_s("tokenizer.ggml.model", "no_vocab");
    if (model == "no_vocab") {
        auto t = std::make_unique<IdTokenizer>();
        const gguf_tensor * te = g.find("token_embd.weight");
        t->n_vocab = te ? (int32_t) te->ne[1] : (int32_t) g.get_u("llama.vocab_size", 0);
        t->eos = (int32_t) (int64_t) g.get_u("tokenizer.ggml.eos_token_id", (uint64_t) -1);
        t->eot = (int32_t) (int64_t) g.get_u("tokenizer.ggml.eot_token_id", (uint64_t) -1);
        return t;
    }
    err = "tokenizer.ggml.model = '" + model + "' is not implemented yet (SURVEY.md §8 row f-1); "
          "use the token-level API (b200_decode) or a no_vocab model";
    return nullptr;
}

}  // namespace b200
