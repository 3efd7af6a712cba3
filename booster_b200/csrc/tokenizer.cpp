// tokenizer.cpp — see tokenizer.hpp. Host-side only; every rule below cites the reference code it restates.
#include "tokenizer.hpp"
#include "gguf.hpp"
#include "unicode_tables.hpp"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <stdexcept>
#include <unordered_map>

namespace b200 {

namespace {

// ------------------------------------------------------------------------------------------------------------
// "no_vocab": prompts are decimal token ids
// ------------------------------------------------------------------------------------------------------------
struct IdTokenizer final : Tokenizer {
    int32_t nv = 0;
    int32_t eos_id = -1, eot_id = -1;
    bool tokenize(const std::string & text, bool, bool, std::vector<int32_t> & out) const override {
        out.clear();
        size_t i = 0;
        while (i < text.size()) {
            while (i < text.size() && (text[i] == ' ' || text[i] == '\n' || text[i] == '\t' || text[i] == ',')) i++;
            if (i >= text.size()) break;
            char * end = nullptr;
            const long v = std::strtol(text.c_str() + i, &end, 10);
            if (end == text.c_str() + i) return false;            // not a number
            if (v < 0 || v >= nv) return false;
            out.push_back((int32_t) v);
            i = (size_t) (end - text.c_str());
        }
        return true;
    }
    std::string piece(int32_t id, bool) const override { return std::to_string(id) + " "; }
    bool is_eog(int32_t id) const override { return id >= 0 && (id == eos_id || id == eot_id); }
    int32_t n_vocab() const override { return nv; }
    int32_t eos() const override { return eos_id; }
    int32_t eot() const override { return eot_id; }
};

// ------------------------------------------------------------------------------------------------------------
// UTF-8 (cpp/src/unicode.cpp:22-120, 563-590): the reference decodes without overlong / surrogate checks and
// throws on a malformed sequence; here a malformed sequence makes tokenize() fail
// ------------------------------------------------------------------------------------------------------------
size_t utf8_len(char c) {
    static const uint8_t len_of[16] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 2, 3, 4};
    return len_of[(uint8_t) c >> 4];
}
bool utf8_decode(const std::string & s, std::vector<uint32_t> & out) {
    out.clear();
    out.reserve(s.size());
    const auto cont = [&](size_t i) { return i < s.size() && ((uint8_t) s[i] & 0xc0) == 0x80; };
    size_t i = 0;
    while (i < s.size()) {
        const uint8_t b = (uint8_t) s[i];
        if (!(b & 0x80)) { out.push_back(b); i += 1; continue; }
        if (!(b & 0x40)) return false;
        if (!(b & 0x20)) {
            if (!cont(i + 1)) return false;
            out.push_back(((uint32_t) (b & 0x1f) << 6) | ((uint8_t) s[i + 1] & 0x3f));
            i += 2; continue;
        }
        if (!(b & 0x10)) {
            if (!cont(i + 1) || !cont(i + 2)) return false;
            out.push_back(((uint32_t) (b & 0x0f) << 12) | ((uint32_t) ((uint8_t) s[i + 1] & 0x3f) << 6) | ((uint8_t) s[i + 2] & 0x3f));
            i += 3; continue;
        }
        if (!(b & 0x08)) {
            if (!cont(i + 1) || !cont(i + 2) || !cont(i + 3)) return false;
            out.push_back(((uint32_t) (b & 0x07) << 18) | ((uint32_t) ((uint8_t) s[i + 1] & 0x3f) << 12) |
                          ((uint32_t) ((uint8_t) s[i + 2] & 0x3f) << 6) | ((uint8_t) s[i + 3] & 0x3f));
            i += 4; continue;
        }
        return false;
    }
    return true;
}
void utf8_append(std::string & out, uint32_t cp) {
    if (cp <= 0x7f) { out.push_back((char) cp); }
    else if (cp <= 0x7ff) { out.push_back((char) (0xc0 | ((cp >> 6) & 0x1f))); out.push_back((char) (0x80 | (cp & 0x3f))); }
    else if (cp <= 0xffff) {
        out.push_back((char) (0xe0 | ((cp >> 12) & 0x0f))); out.push_back((char) (0x80 | ((cp >> 6) & 0x3f)));
        out.push_back((char) (0x80 | (cp & 0x3f)));
    } else {
        out.push_back((char) (0xf0 | ((cp >> 18) & 0x07))); out.push_back((char) (0x80 | ((cp >> 12) & 0x3f)));
        out.push_back((char) (0x80 | ((cp >> 6) & 0x3f))); out.push_back((char) (0x80 | (cp & 0x3f)));
    }
}

// codepoint classes of the pre-tokenizer regex (unicode_tables.hpp, generated)
enum : uint8_t { CL_LETTER = 1, CL_NUMBER = 2, CL_SPACE = 4 };
template <size_t N> bool in_ranges(const CptRange (&r)[N], uint32_t cp) {
    size_t lo = 0, hi = N;
    while (lo < hi) {
        const size_t mid = (lo + hi) / 2;
        if (cp < r[mid].first) hi = mid; else if (cp > r[mid].last) lo = mid + 1; else return true;
    }
    return false;
}
uint8_t cpt_class(uint32_t cp) {
    uint8_t c = 0;
    if (in_ranges(k_cpt_letter, cp)) c |= CL_LETTER;
    else if (in_ranges(k_cpt_number, cp)) c |= CL_NUMBER;
    for (uint32_t w : k_cpt_whitespace) if (w == cp) { c |= CL_SPACE; break; }
    return c;
}

// ------------------------------------------------------------------------------------------------------------
// the byte <-> printable-codepoint alphabet of byte-level BPE (cpp/src/unicode.cpp:154-200): bytes '!'..'~',
// 0xA1..0xAC, 0xAE..0xFF stand for themselves, the other 68 bytes map to U+0100.. in byte order
// ------------------------------------------------------------------------------------------------------------
struct ByteAlphabet {
    std::string enc[256];
    std::unordered_map<std::string, uint8_t> dec;
    ByteAlphabet() {
        bool own[256] = {};
        for (int b = 0x21; b <= 0x7e; b++) own[b] = true;
        for (int b = 0xa1; b <= 0xac; b++) own[b] = true;
        for (int b = 0xae; b <= 0xff; b++) own[b] = true;
        uint32_t next = 256;
        for (int b = 0; b < 256; b++) {
            utf8_append(enc[b], own[b] ? (uint32_t) b : next++);
            dec[enc[b]] = (uint8_t) b;
        }
    }
};
const ByteAlphabet & byte_alphabet() { static const ByteAlphabet a; return a; }

// ------------------------------------------------------------------------------------------------------------
// vocabulary (cpp/src/llama.cpp:5250-5760)
// ------------------------------------------------------------------------------------------------------------
// token kinds of tokenizer.ggml.token_type (cpp/include/llama.h llama_token_type)
enum : int { TT_UNDEFINED = 0, TT_NORMAL = 1, TT_UNKNOWN = 2, TT_CONTROL = 3, TT_USER_DEFINED = 4, TT_UNUSED = 5, TT_BYTE = 6 };

struct PairHash {
    size_t operator()(const std::pair<std::string, std::string> & p) const {
        return std::hash<std::string>()(p.first) * 1000003u ^ std::hash<std::string>()(p.second);
    }
};

struct TextTokenizer final : Tokenizer {
    bool bpe = false;                                  // false: SPM ("llama"), true: byte-level BPE ("gpt2")
    std::vector<std::string> text;                     // id -> token text
    std::vector<float> score;
    std::vector<int> kind;                             // TT_*
    std::unordered_map<std::string, int32_t> id_of;    // a text listed twice keeps its LAST id (:5519)
    std::unordered_map<std::pair<std::string, std::string>, int, PairHash> rank;   // first listing wins (:5327 emplace)
    std::vector<int32_t> specials;                     // control | user-defined | unknown ids, longest text first (:5680-5694)
    std::vector<std::string> piece_cache;              // llama_token_to_piece(id, special = true) of every id (:5698-5709)
    int32_t bos_id = -1, eos_id = -1, unk_id = -1, eot_id = -1, eom_id = -1, lf_id = -1;
    bool add_bos = false, add_eos = false, add_space_prefix = false, ignore_merges = false;

    int32_t n_vocab() const override { return (int32_t) text.size(); }
    int32_t bos() const override { return bos_id; }
    int32_t eos() const override { return eos_id; }
    int32_t eot() const override { return eot_id; }
    int32_t linefeed() const override { return lf_id; }
    bool is_eog(int32_t id) const override { return id != -1 && (id == eos_id || id == eot_id || id == eom_id); }   // llama-vocab.cpp:1433-1439

    bool is_special_kind(int32_t id) const { return kind[(size_t) id] == TT_UNKNOWN || kind[(size_t) id] == TT_CONTROL; }

    // ---- llama_token_to_piece_impl (cpp/src/llama-vocab.cpp:1539-1608)
    std::string render(int32_t id) const {
        const std::string & t = text[(size_t) id];
        const int k = kind[(size_t) id];
        if (k == TT_UNKNOWN || k == TT_CONTROL || k == TT_USER_DEFINED) return t;
        if (!bpe) {
            if (k == TT_NORMAL) {                      // llama_unescape_whitespace: U+2581 -> ' '
                std::string r;
                for (size_t i = 0; i < t.size();) {
                    if (t.compare(i, 3, "\xe2\x96\x81") == 0) { r.push_back(' '); i += 3; } else { r.push_back(t[i]); i++; }
                }
                return r;
            }
            if (k == TT_BYTE) {                        // "<0xXX>" (llama_token_to_byte, :130-139)
                const std::string hex = t.size() >= 5 ? t.substr(3, 2) : std::string();
                return std::string(1, (char) std::strtol(hex.c_str(), nullptr, 16));
            }
            return std::string();
        }
        if (k == TT_NORMAL) {                          // llama_decode_text (:1518-1536)
            std::vector<uint32_t> cps;
            std::string r;
            if (!utf8_decode(t, cps)) return r;
            const ByteAlphabet & A = byte_alphabet();
            for (uint32_t cp : cps) {
                std::string u; utf8_append(u, cp);
                const auto it = A.dec.find(u);
                if (it != A.dec.end()) { r.push_back((char) it->second); continue; }
                r += "[UNK_BYTE_0x";
                char hx[4];
                for (unsigned char ch : u) { std::snprintf(hx, sizeof hx, "%02x", ch); r += hx; }
                r += t + "]";
            }
            return r;
        }
        return std::string();
    }
    std::string piece(int32_t id, bool special) const override {
        if (id < 0 || id >= n_vocab()) return std::string();
        if (!special && is_special_kind(id)) return std::string();
        return piece_cache[(size_t) id];
    }

    // ---- tokenizer_st_partition (cpp/src/llama-vocab.cpp:1123-1241): cut the text at special tokens, longest first
    struct Fragment { bool is_token; int32_t token; size_t off, len; };
    void partition(const std::string & raw, bool parse_special, std::vector<Fragment> & frags) const {
        frags.clear();
        if (raw.empty()) return;
        frags.push_back({false, -1, 0, raw.size()});
        for (int32_t sid : specials) {
            if (!parse_special && is_special_kind(sid)) continue;   // user-defined tokens are always cut out (:1128-1135)
            const std::string & st = text[(size_t) sid];
            if (st.empty()) continue;
            std::vector<Fragment> next;
            next.reserve(frags.size() + 2);
            for (const Fragment & f : frags) {
                if (f.is_token) { next.push_back(f); continue; }
                size_t off = f.off, len = f.len;
                for (;;) {
                    const size_t m = raw.find(st, off);
                    if (m == std::string::npos || m + st.size() > off + len) { next.push_back({false, -1, off, len}); break; }
                    if (m > off) next.push_back({false, -1, off, m - off});
                    next.push_back({true, sid, 0, 0});
                    const size_t roff = m + st.size();
                    if (roff >= off + len) break;
                    len = off + len - roff; off = roff;
                }
            }
            frags.swap(next);
        }
    }

    // ---- SPM (llm_tokenizer_spm, cpp/src/llama-vocab.cpp:190-310): symbols = UTF-8 characters; repeatedly merge the
    // adjacent pair whose concatenation is the token with the highest score (ties: leftmost); what is left and is not
    // a token goes out byte by byte (<0xXX>)
    struct Sym { int prev, next; size_t off, n; };
    bool spm_piece(const std::string & s, std::vector<int32_t> & out) const {
        std::vector<Sym> sym;
        for (size_t off = 0; off < s.size();) {
            const size_t n = std::min(utf8_len(s[off]), s.size() - off);
            sym.push_back({(int) sym.size() - 1, 0, off, n});
            off += n;
            sym.back().next = off == s.size() ? -1 : (int) sym.size();
        }
        if (sym.empty()) return true;
        struct Cand { int l, r; float score; size_t size; };
        const auto worse = [](const Cand & a, const Cand & b) { return a.score < b.score || (a.score == b.score && a.l > b.l); };
        std::priority_queue<Cand, std::vector<Cand>, decltype(worse)> q(worse);
        const auto consider = [&](int l, int r) {
            if (l < 0 || r < 0) return;
            const auto it = id_of.find(s.substr(sym[(size_t) l].off, sym[(size_t) l].n + sym[(size_t) r].n));
            if (it == id_of.end()) return;
            q.push({l, r, score[(size_t) it->second], sym[(size_t) l].n + sym[(size_t) r].n});
        };
        for (size_t i = 1; i < sym.size(); i++) consider((int) i - 1, (int) i);
        while (!q.empty()) {
            const Cand c = q.top(); q.pop();
            Sym & L = sym[(size_t) c.l]; Sym & R = sym[(size_t) c.r];
            if (L.n == 0 || R.n == 0 || L.n + R.n != c.size) continue;      // stale candidate
            L.n += R.n; R.n = 0;
            L.next = R.next;
            if (R.next >= 0) sym[(size_t) R.next].prev = c.l;
            consider(L.prev, c.l);
            consider(c.l, L.next);
        }
        for (int i = 0; i != -1; i = sym[(size_t) i].next) {
            const Sym & y = sym[(size_t) i];
            const auto it = id_of.find(s.substr(y.off, y.n));
            if (it != id_of.end()) { out.push_back(it->second); continue; }
            for (size_t j = 0; j < y.n; j++) {                               // llama_byte_to_token_impl (:1394-1409)
                const uint8_t b = (uint8_t) s[y.off + j];
                char name[8];
                std::snprintf(name, sizeof name, "<0x%02X>", b);
                auto bt = id_of.find(name);
                if (bt == id_of.end()) bt = id_of.find(std::string(1, (char) b));
                if (bt == id_of.end()) return false;                         // the reference throws here
                out.push_back(bt->second);
            }
        }
        return true;
    }

    // ---- LLaMA-3 pre-tokenizer: unicode_regex_split_custom_llama3 (cpp/src/unicode.cpp:344-483), the hand-written
    // matcher of
    //   (?i:'s|'t|'re|'ve|'m|'ll|'d) | [^\r\n\p{L}\p{N}]?\p{L}+ | \p{N}{1,3} | ?[^\s\p{L}\p{N}]+[\r\n]* | \s*[\r\n]+ | \s+(?!\S) | \s+
    // over codepoints. Returns word boundaries as (begin, end) codepoint index pairs.
    static void split_llama3(const std::vector<uint32_t> & cp, std::vector<std::pair<size_t, size_t>> & words) {
        const size_t n = cp.size();
        const auto in = [&](size_t i) { return i < n; };
        const auto cls = [&](size_t i) -> uint8_t { return i < n ? cpt_class(cp[i]) : (uint8_t) 0; };
        const auto low = [&](size_t i) -> uint32_t { const uint32_t c = cp[i]; return c >= 'A' && c <= 'Z' ? c + 32 : c; };
        size_t begin = 0;
        const auto emit = [&](size_t end) { if (end > begin) words.emplace_back(begin, end); begin = end; };
        size_t pos = 0;
        while (pos < n) {
            const uint32_t c = cp[pos];
            const uint8_t k = cls(pos);
            if (c == '\'' && pos + 1 < n) {                                   // contractions, ASCII case-insensitive
                const uint32_t a = low(pos + 1);
                if (a == 's' || a == 't' || a == 'm' || a == 'd') { pos += 2; emit(pos); continue; }
                if (pos + 2 < n) {
                    const uint32_t b = low(pos + 2);
                    if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e') || (a == 'l' && b == 'l')) { pos += 3; emit(pos); continue; }
                }
            }
            if (!(c == '\r' || c == '\n' || (k & CL_NUMBER))) {               // one optional non-letter, then letters
                if ((k & CL_LETTER) || (cls(pos + 1) & CL_LETTER)) {
                    pos++;
                    while (cls(pos) & CL_LETTER) pos++;
                    emit(pos);
                    continue;
                }
            }
            if (k & CL_NUMBER) {                                              // digits in groups of at most three
                size_t ini = pos;
                while (cls(pos) & CL_NUMBER) {
                    if (++pos - ini >= 3) { emit(pos); ini = pos; }
                }
                emit(pos);
                continue;
            }
            {                                                                  // optional space, punctuation run, newlines
                const bool sp = c == ' ';
                const bool other = sp ? !(cls(pos + 1) & (CL_SPACE | CL_LETTER | CL_NUMBER))   // (also true past the end)
                                      : !(k & (CL_SPACE | CL_LETTER | CL_NUMBER));
                if (other) {
                    pos += sp ? 1 : 0;
                    while (in(pos) && !(cls(pos) & (CL_SPACE | CL_LETTER | CL_NUMBER))) pos++;
                    while (in(pos) && (cp[pos] == '\r' || cp[pos] == '\n')) pos++;
                    emit(pos);
                    continue;
                }
            }
            size_t n_ws = 0, after_last_nl = 0;
            while (cls(pos + n_ws) & CL_SPACE) {
                const uint32_t w = cp[pos + n_ws];
                if (w == '\r' || w == '\n') after_last_nl = pos + n_ws + 1;
                n_ws++;
            }
            if (after_last_nl > 0) { pos = after_last_nl; emit(pos); continue; }          // \s*[\r\n]+
            if (n_ws > 1 && in(pos + n_ws)) { pos += n_ws - 1; emit(pos); continue; }     // \s+(?!\S)
            if (n_ws > 0) { pos += n_ws; emit(pos); continue; }                            // \s+
            emit(++pos);
        }
    }

    // ---- byte-level BPE (llm_tokenizer_bpe, cpp/src/llama-vocab.cpp:340-629): per pre-tokenizer word (byte-encoded),
    // merge the adjacent pair with the lowest merge rank (ties: leftmost) until none is listed
    bool bpe_piece(const std::string & raw, std::vector<int32_t> & out) const {
        std::vector<uint32_t> cps;
        if (!utf8_decode(raw, cps)) return false;
        std::vector<std::pair<size_t, size_t>> words;
        split_llama3(cps, words);
        const ByteAlphabet & A = byte_alphabet();
        for (const auto & wd : words) {
            std::string plain, w;
            for (size_t i = wd.first; i < wd.second; i++) utf8_append(plain, cps[i]);
            for (unsigned char b : plain) w += A.enc[b];
            std::vector<Sym> sym;
            if (ignore_merges && id_of.count(w)) {
                sym.push_back({-1, -1, 0, w.size()});
            } else {
                for (size_t off = 0; off < w.size();) {
                    const size_t nn = std::min(w.size() - off, utf8_len(w[off]));
                    sym.push_back({(int) sym.size() - 1, 0, off, nn});
                    off += nn;
                    sym.back().next = off == w.size() ? -1 : (int) sym.size();
                }
            }
            struct Cand { int l, r, rank; std::string joined; };
            const auto worse = [](const Cand & a, const Cand & b) { return a.rank > b.rank || (a.rank == b.rank && a.l > b.l); };
            std::priority_queue<Cand, std::vector<Cand>, decltype(worse)> q(worse);
            const auto consider = [&](int l, int r) {
                if (l < 0 || r < 0) return;
                const std::string a = w.substr(sym[(size_t) l].off, sym[(size_t) l].n), b = w.substr(sym[(size_t) r].off, sym[(size_t) r].n);
                const auto it = rank.find(std::make_pair(a, b));
                if (it == rank.end()) return;
                q.push({l, r, it->second, a + b});
            };
            for (size_t i = 1; i < sym.size(); i++) consider((int) i - 1, (int) i);
            while (!q.empty()) {
                const Cand c = q.top(); q.pop();
                Sym & L = sym[(size_t) c.l]; Sym & R = sym[(size_t) c.r];
                if (L.n == 0 || R.n == 0) continue;
                if (w.compare(L.off, L.n, c.joined, 0, L.n) != 0 || L.n + R.n != c.joined.size() ||
                    w.compare(R.off, R.n, c.joined, L.n, R.n) != 0) continue;                  // stale candidate
                L.n += R.n; R.n = 0;
                L.next = R.next;
                if (R.next >= 0) sym[(size_t) R.next].prev = c.l;
                consider(L.prev, c.l);
                consider(c.l, L.next);
            }
            for (const Sym & y : sym) {
                if (y.n == 0) continue;
                const std::string t = w.substr(y.off, y.n);
                const auto it = id_of.find(t);
                if (it != id_of.end()) { out.push_back(it->second); continue; }
                for (char ch : t) {                                            // (:585-591) bytes that happen to be tokens
                    const auto b1 = id_of.find(std::string(1, ch));
                    if (b1 != id_of.end()) out.push_back(b1->second);
                }
            }
        }
        return true;
    }

    // ---- llama_tokenize_internal (cpp/src/llama-vocab.cpp:1243-1330)
    bool tokenize(const std::string & raw, bool add_special, bool parse_special, std::vector<int32_t> & out) const override {
        out.clear();
        std::vector<Fragment> frags;
        partition(raw, parse_special, frags);
        if (!bpe) {
            bool prev_special = true;                                          // prefix with a space if first
            if (add_special && add_bos) { if (bos_id < 0) return false; out.push_back(bos_id); }
            for (const Fragment & f : frags) {
                if (f.is_token) { out.push_back(f.token); prev_special = true; continue; }
                std::string s = raw.substr(f.off, f.len);
                if (add_space_prefix && prev_special) s = " " + s;
                std::string esc;                                               // llama_escape_whitespace: ' ' -> U+2581
                for (char ch : s) { if (ch == ' ') esc += "\xe2\x96\x81"; else esc.push_back(ch); }
                if (!spm_piece(esc, out)) return false;
                prev_special = false;
            }
            if (add_special && add_eos) { if (eos_id < 0) return false; out.push_back(eos_id); }
            return true;
        }
        if (add_special && add_bos) { if (bos_id < 0) return false; out.push_back(bos_id); }
        for (const Fragment & f : frags) {
            if (f.is_token) { out.push_back(f.token); continue; }
            if (!bpe_piece(raw.substr(f.off, f.len), out)) return false;
        }
        if (add_special && add_eos) { if (eos_id < 0) return false; out.push_back(eos_id); }
        return true;
    }
};

std::unique_ptr<Tokenizer> load_text_tokenizer(const gguf_file & g, const std::string & model, std::string & err) {
    auto t = std::make_unique<TextTokenizer>();
    t->bpe = model == "gpt2";
    const auto tk = g.kv.find("tokenizer.ggml.tokens");
    if (tk == g.kv.end() || tk->second.arr_s.empty()) { err = "cannot find tokenizer vocab in model file"; return nullptr; }
    t->text = tk->second.arr_s;
    const size_t n = t->text.size();
    t->score.assign(n, 0.f);
    t->kind.assign(n, TT_NORMAL);
    const auto sc = g.kv.find("tokenizer.ggml.scores");
    if (sc != g.kv.end() && sc->second.arr_f.size() == n) for (size_t i = 0; i < n; i++) t->score[i] = (float) sc->second.arr_f[i];
    const auto tt = g.kv.find("tokenizer.ggml.token_type");
    if (tt != g.kv.end() && tt->second.arr_f.size() == n)
        for (size_t i = 0; i < n; i++) { const int v = (int) tt->second.arr_f[i]; t->kind[i] = v >= TT_NORMAL && v <= TT_BYTE ? v : TT_UNDEFINED; }
    for (size_t i = 0; i < n; i++) t->id_of[t->text[i]] = (int32_t) i;

    if (t->bpe) {
        const std::string pre = g.get_s("tokenizer.ggml.pre", "");
        if (!(pre == "llama3" || pre == "llama-v3" || pre == "llama-bpe")) {
            err = "tokenizer.ggml.pre = '" + pre + "' is not implemented (LLaMA-3 family pre-tokenizer only: llama3 | llama-v3 | llama-bpe)";
            return nullptr;
        }
        const auto mg = g.kv.find("tokenizer.ggml.merges");
        if (mg == g.kv.end()) { err = "cannot find tokenizer merges in model file"; return nullptr; }
        for (size_t i = 0; i < mg->second.arr_s.size(); i++) {                 // "first second", split at the first space after byte 0
            const std::string & word = mg->second.arr_s[i];
            std::string a, b;
            const size_t p = word.find(' ', 1);
            if (p != std::string::npos) { a = word.substr(0, p); b = word.substr(p + 1); }
            t->rank.emplace(std::make_pair(a, b), (int) i);
        }
        t->bos_id = 11; t->eos_id = 11;                                        // BPE defaults (:5330-5337)
        t->ignore_merges = true; t->add_bos = true;                            // llama3 family (:5389-5395)
        t->add_space_prefix = false;
    } else {
        t->bos_id = 1; t->eos_id = 2; t->unk_id = 0;                           // SPM defaults (:5283-5292, 5473-5478)
        t->add_space_prefix = true; t->add_bos = true; t->add_eos = false;
    }
    if (g.has("tokenizer.ggml.add_space_prefix")) t->add_space_prefix = g.get_u("tokenizer.ggml.add_space_prefix", 1) != 0;
    const auto special_id = [&](const char * key, int32_t & id) {              // (:5614-5645): out-of-range ids keep the default
        if (!g.has(key)) return;
        const uint64_t v = g.get_u(key, 0);
        if (v < n) id = (int32_t) v;
    };
    special_id("tokenizer.ggml.bos_token_id", t->bos_id);
    special_id("tokenizer.ggml.eos_token_id", t->eos_id);
    special_id("tokenizer.ggml.unknown_token_id", t->unk_id);
    special_id("tokenizer.ggml.eot_token_id", t->eot_id);
    special_id("tokenizer.ggml.eom_token_id", t->eom_id);
    if (g.has("tokenizer.ggml.add_bos_token")) t->add_bos = g.get_u("tokenizer.ggml.add_bos_token", 1) != 0;
    if (g.has("tokenizer.ggml.add_eos_token")) t->add_eos = g.get_u("tokenizer.ggml.add_eos_token", 0) != 0;
    if (t->eot_id == -1) {
        // end-of-turn token found by its text (:5660-5680). The reference takes whichever match its unordered_map yields
        // first; with more than one candidate in a vocabulary the lowest id is taken here.
        for (const char * cand : {"<|eot_id|>", "<|im_end|>", "<|end|>", "<end_of_turn>", "<|endoftext|>"}) {
            const auto it = t->id_of.find(cand);
            if (it != t->id_of.end() && (t->eot_id == -1 || it->second < t->eot_id)) t->eot_id = it->second;
        }
    }
    if (t->eom_id == -1) { const auto it = t->id_of.find("<|eom_id|>"); if (it != t->id_of.end()) t->eom_id = it->second; }
    for (size_t i = 0; i < n; i++)
        if (t->kind[i] == TT_CONTROL || t->kind[i] == TT_USER_DEFINED || t->kind[i] == TT_UNKNOWN) t->specials.push_back((int32_t) i);
    // the same std::sort call on the same input (ids ascending) as the reference: equal lengths end up in the same order
    std::sort(t->specials.begin(), t->specials.end(),
              [&](const int32_t a, const int32_t b) { return t->text[(size_t) a].size() > t->text[(size_t) b].size(); });
    t->piece_cache.resize(n);
    for (size_t i = 0; i < n; i++) t->piece_cache[i] = t->render((int32_t) i);
    // the newline token (cpp/src/llama.cpp:5585-5597): SPM = the byte token of '\n' (else the padding id, -1 by default);
    // BPE = the first token of the TEXT U+010A pushed through the byte-level tokenizer (for a LLaMA-3 vocabulary that is the
    // token of byte 0xC4, 'Ä', not a newline — the reference's load log says so itself; reproduced, the sampler depends on it)
    if (!t->bpe) {
        auto it = t->id_of.find("<0x0A>");
        if (it == t->id_of.end()) it = t->id_of.find("\n");
        if (it != t->id_of.end()) t->lf_id = it->second;
    } else {
        std::vector<int32_t> ids;
        if (t->tokenize("\xC4\x8A", false, false, ids) && !ids.empty()) t->lf_id = ids[0];
    }
    return t;
}

}  // namespace

int codepoint_class(uint32_t cp) { return (int) cpt_class(cp); }

static std::unique_ptr<Tokenizer> make_tokenizer_impl(const std::string & gguf_path, std::string & err);
// never throws: the callers are extern "C" entry points reached from cgo (initContext)
std::unique_ptr<Tokenizer> make_tokenizer(const std::string & gguf_path, std::string & err) {
    try {
        return make_tokenizer_impl(gguf_path, err);
    } catch (const std::exception & e) {
        err = std::string("tokenizer load failed: ") + e.what();
        return nullptr;
    }
}
static std::unique_ptr<Tokenizer> make_tokenizer_impl(const std::string & gguf_path, std::string & err) {
    gguf_file g;
    err = g.open(gguf_path);
    if (!err.empty()) return nullptr;
    const std::string model = g.get_s("tokenizer.ggml.model", "no_vocab");
    if (model == "no_vocab") {
        auto t = std::make_unique<IdTokenizer>();
        const gguf_tensor * te = g.find("token_embd.weight");
        t->nv = te ? (int32_t) te->ne[1] : (int32_t) g.get_u("llama.vocab_size", 0);
        t->eos_id = (int32_t) (int64_t) g.get_u("tokenizer.ggml.eos_token_id", (uint64_t) -1);
        t->eot_id = (int32_t) (int64_t) g.get_u("tokenizer.ggml.eot_token_id", (uint64_t) -1);
        return t;
    }
    if (model == "llama" || model == "gpt2") return load_text_tokenizer(g, model, err);
    err = "tokenizer.ggml.model = '" + model + "' is not implemented (supported: no_vocab, llama (SPM), gpt2 (BPE, LLaMA-3 pre-tokenizer))";
    return nullptr;
}

}  // namespace b200

// codepoint class used by the pre-tokenizer (bit 0 \p{L}, bit 1 \p{N}, bit 2 \s): include/booster_b200.h
extern "C" int b200_cpt_class(uint32_t cp) { return b200::codepoint_class(cp); }
