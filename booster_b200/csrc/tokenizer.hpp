// tokenizer.hpp — host-side text <-> token-id interface used by bridge.cpp.
// Replaces the reference's llama_tokenize / llama_token_to_piece / llama_token_is_eog calls
// (cpp/bridge.cpp:275-278, 630, 640; implementation cpp/src/llama-vocab.cpp, cpp/src/unicode.cpp; vocabulary
// loading cpp/src/llama.cpp:5250-5760).
//
// Three vocabularies, chosen by tokenizer.ggml.model:
//   "no_vocab" — the reference cannot tokenize text for such models at all (cpp/src/llama.cpp:5267-5268); the
//                bridge defines the prompt as white-space separated decimal token ids and a piece as "<id> ";
//   "llama"    — SentencePiece-style score-driven merges with byte fallback (LLaMA-2, Mistral);
//   "gpt2"     — byte-level BPE with merge ranks and the LLaMA-3 pre-tokenizer (tokenizer.ggml.pre = llama3 |
//                llama-v3 | llama-bpe). Other pre-tokenizers fail loudly at load.
// Results are pinned token for token against the reference's own tokenizer (oracle/_ref) on synthetic
// vocabularies: tests/test_tokenizer.py, tests/golden/tokenizer_*.json.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace b200 {

struct Tokenizer {
    virtual ~Tokenizer() = default;
    // llama_tokenize(model, text, add_special, parse_special): false when the text cannot be tokenized
    // (ids out of range for no_vocab, invalid UTF-8 or a byte without a token for the text vocabularies)
    virtual bool tokenize(const std::string & text, bool add_special, bool parse_special, std::vector<int32_t> & out) const = 0;
    // llama_token_to_piece(ctx, id, special)
    virtual std::string piece(int32_t id, bool special) const = 0;
    virtual bool is_eog(int32_t id) const = 0;
    virtual int32_t n_vocab() const = 0;
    virtual int32_t bos() const { return -1; }
    virtual int32_t eos() const { return -1; }
    virtual int32_t eot() const { return -1; }   // llama_token_eot (cpp/src/llama-vocab.cpp:1489-1491)
    virtual int32_t linefeed() const { return -1; }   // llama_token_nl (cpp/src/llama-vocab.cpp:1461-1463; found at cpp/src/llama.cpp:5585-5597)
};

// bit 0 \p{L}, bit 1 \p{N}, bit 2 \s of a codepoint (unicode_tables.hpp)
int codepoint_class(uint32_t cp);

// reads tokenizer.* metadata from the GGUF; returns nullptr and sets err when the model's tokenizer
// is not implemented
std::unique_ptr<Tokenizer> make_tokenizer(const std::string & gguf_path, std::string & err);

}  // namespace b200
