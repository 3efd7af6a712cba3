// tokenizer.hpp — host-side text <-> token-id interface used by bridge.cpp.
// Replaces the reference's llama_tokenize / llama_token_to_piece / llama_token_is_eog calls
// (cpp/bridge.cpp:275-278, 630, 640; implementation cpp/src/llama-vocab.cpp).
//
// Round 1 ships the "no_vocab" tokenizer only (tokenizer.ggml.model == "no_vocab",
// cpp/src/llama.cpp:5267-5268): the reference cannot tokenize text for such models at all, so the
// bridge defines the prompt as white-space separated decimal token ids and a piece as "<id> ".
// BPE / SPM are SURVEY.md §8 row f-1 ("next"); make_tokenizer fails loudly for them.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace b200 {

struct Tokenizer {
    virtual ~Tokenizer() = default;
    virtual bool tokenize(const std::string & text, std::vector<int32_t> & out) const = 0;
    virtual std::string piece(int32_t id) const = 0;
    virtual bool is_eog(int32_t id) const = 0;
};

// reads tokenizer.* metadata from the GGUF; returns nullptr and sets err when the model's tokenizer
// is not implemented
std::unique_ptr<Tokenizer> make_tokenizer(const std::string & gguf_path, std::string & err);

}  // namespace b200
