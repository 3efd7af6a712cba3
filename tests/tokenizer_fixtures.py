"""Synthetic vocabularies and test strings for the tokenizer parity tests (no tokenizer files exist offline).

Two vocab-only GGUFs are built deterministically:
  * SPM  ("llama"): <unk>/<s>/</s>, the 256 byte tokens <0x00>..<0xFF>, single characters and multi-character pieces
                    cut from a small corpus with SentencePiece-like scores (ties included), user-defined and control tokens;
  * BPE  ("gpt2", pre = llama-bpe): the 256-symbol byte alphabet, merges learned by a tiny BPE trainer on the same
                    corpus, LLaMA-3 style control tokens and one user-defined token.
The GGUFs carry the LLaMA hyper-parameter keys the reference's loader insists on, and no tensors.
"""
from __future__ import annotations

import collections
import os
import re
import sys
from typing import Dict, List, Tuple

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from booster_b200 import gguf_io as G  # noqa: E402

CORPUS = (
    "The quick brown fox jumps over the lazy dog. I'm sure you're right, they've said it'll work and he'd agree; "
    "DON'T shout, it's 'fine'. Numbers: 7 42 365 1024 65536 3.14159 2024-10-17 1,000,000. "
    "Tokenizers split text into pieces: the tokenizer, tokenization, tokens, token. "
    "Whitespace   runs\tand\ttabs\n\nnew paragraphs\r\nwindows lines   \n   indented. "
    "Unicode: naïve café résumé Ångström über straße, Ελληνικά, русский язык, עברית, العربية, 日本語のテキスト, 中文文本, 한국어. "
    "Emoji 😀🎉 and symbols ±×÷ €£¥ → ∑∫ ©®™ — “quotes” … #hash @user http://example.com/path?q=1&r=2 "
    "for (int i = 0; i < n; i++) { x[i] += y[i] * 2; } // code\n"
)

TT_NORMAL, TT_UNKNOWN, TT_CONTROL, TT_USER, TT_UNUSED, TT_BYTE = 1, 2, 3, 4, 5, 6


def _byte_alphabet() -> List[str]:
    own = set(range(0x21, 0x7F)) | set(range(0xA1, 0xAD)) | set(range(0xAE, 0x100))
    out, nxt = [], 256
    for b in range(256):
        if b in own:
            out.append(chr(b))
        else:
            out.append(chr(nxt)); nxt += 1
    return out


def spm_vocab() -> Tuple[List[str], List[float], List[int]]:
    rng = np.random.default_rng(7)
    toks = ["<unk>", "<s>", "</s>"] + [f"<0x{b:02X}>" for b in range(256)]
    kinds = [TT_UNKNOWN, TT_CONTROL, TT_CONTROL] + [TT_BYTE] * 256
    scores = [0.0, 0.0, 0.0] + [0.0] * 256
    text = CORPUS.replace(" ", "▁")
    seen = set(toks)
    pieces: List[str] = []
    # characters (a few corpus characters are left out on purpose: they must fall back to bytes)
    chars = sorted(set(text))
    left_out = set("ž€Å😀ע")
    for ch in chars:
        if ch not in left_out and ch not in seen:
            pieces.append(ch); seen.add(ch)
    # multi-character pieces: random substrings of the corpus words (with and without the leading U+2581)
    words = re.findall(r"▁?[^▁]+", text)
    for _ in range(900):
        w = words[int(rng.integers(len(words)))]
        a = int(rng.integers(0, max(1, len(w) - 1)))
        b = int(rng.integers(a + 2, a + 8))
        p = w[a:b]
        if len(p) >= 2 and p not in seen:
            pieces.append(p); seen.add(p)
    for p in ["▁▁", "▁▁▁", "▁the", "▁The", "ing", "tion", "▁token", "izer", "\n\n", "▁\n"]:
        if p not in seen:
            pieces.append(p); seen.add(p)
    order = rng.permutation(len(pieces))
    for rank, i in enumerate(order):
        toks.append(pieces[i]); kinds.append(TT_NORMAL)
        scores.append(-float(rank // 3))          # every score shared by three pieces: exercises the leftmost-first tie rule
    for t, k in [("<|user|>", TT_USER), ("<|assistant|>", TT_USER), ("<|im_end|>", TT_CONTROL), ("<pad>", TT_UNUSED), ("<|im_start|>", TT_CONTROL)]:
        toks.append(t); kinds.append(k); scores.append(0.0)
    return toks, scores, kinds


def _pretokenize_simple(text: str) -> List[str]:
    return re.findall(r"'[a-zA-Z]{1,2}|[^\r\n\w]?[^\W\d_]+|\d{1,3}| ?[^\s\w]+[\r\n]*|\s+", text)


def bpe_vocab(n_merges: int = 600) -> Tuple[List[str], List[str], List[int]]:
    alpha = _byte_alphabet()
    words = collections.Counter(_pretokenize_simple(CORPUS * 3))
    seqs: Dict[Tuple[str, ...], int] = {}
    for w, c in words.items():
        key = tuple(alpha[b] for b in w.encode("utf-8"))
        seqs[key] = seqs.get(key, 0) + c
    merges: List[str] = []
    toks = list(alpha)
    have = set(toks)
    for _ in range(n_merges):
        pairs = collections.Counter()
        for s, c in seqs.items():
            for a, b in zip(s, s[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        (a, b), cnt = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        if cnt < 2:
            break
        merges.append(f"{a} {b}")
        if a + b not in have:
            toks.append(a + b); have.add(a + b)
        new = {}
        for s, c in seqs.items():
            out, i = [], 0
            while i < len(s):
                if i + 1 < len(s) and s[i] == a and s[i + 1] == b:
                    out.append(a + b); i += 2
                else:
                    out.append(s[i]); i += 1
            new[tuple(out)] = new.get(tuple(out), 0) + c
        seqs = new
    kinds = [TT_NORMAL] * len(toks)
    # a whole-word token that the merges cannot build: tokenizer_ignore_merges must still pick it
    for w in ["Ġtokenization", "ĠWhitespace"]:
        if w not in have:
            toks.append(w); kinds.append(TT_NORMAL); have.add(w)
    for t, k in [("<|begin_of_text|>", TT_CONTROL), ("<|end_of_text|>", TT_CONTROL), ("<|start_header_id|>", TT_CONTROL),
                 ("<|end_header_id|>", TT_CONTROL), ("<|eot_id|>", TT_CONTROL), ("<|eom_id|>", TT_CONTROL), ("<tool>", TT_USER)]:
        toks.append(t); kinds.append(k)
    return toks, merges, kinds


def vocab_kv(kind: str, pad_to: int = 1) -> Dict[str, tuple]:
    """tokenizer.* keys of the synthetic vocabulary; pad_to > 1 appends UNUSED filler tokens up to a multiple (a model
    file needs n_vocab % 32 == 0 for the output matrix tiles)"""
    if kind == "spm":
        toks, scores, kinds = spm_vocab()
        while len(toks) % pad_to:
            toks.append(f"<fill_{len(toks)}>"); scores.append(0.0); kinds.append(TT_UNUSED)
        extra = {"tokenizer.ggml.model": ("str", "llama"),
                 "tokenizer.ggml.tokens": ("arr", ("str", toks)),
                 "tokenizer.ggml.scores": ("arr", ("f32", scores)),
                 "tokenizer.ggml.token_type": ("arr", ("i32", kinds)),
                 "tokenizer.ggml.bos_token_id": ("u32", 1), "tokenizer.ggml.eos_token_id": ("u32", 2),
                 "tokenizer.ggml.unknown_token_id": ("u32", 0)}
    else:
        toks, merges, kinds = bpe_vocab()
        while len(toks) % pad_to:
            toks.append(f"<fill_{len(toks)}>"); kinds.append(TT_UNUSED)
        extra = {"tokenizer.ggml.model": ("str", "gpt2"), "tokenizer.ggml.pre": ("str", "llama-bpe"),
                 "tokenizer.ggml.tokens": ("arr", ("str", toks)),
                 "tokenizer.ggml.token_type": ("arr", ("i32", kinds)),
                 "tokenizer.ggml.merges": ("arr", ("str", merges)),
                 "tokenizer.ggml.bos_token_id": ("u32", toks.index("<|begin_of_text|>")),
                 "tokenizer.ggml.eos_token_id": ("u32", toks.index("<|end_of_text|>"))}
    extra["llama.vocab_size"] = ("u32", len(toks))
    return extra


def write_vocab_gguf(path: str, kind: str) -> None:
    G.write_gguf(path, G.llama_kv(G.CONFIGS["tiny"], "F32", vocab_kv(kind)), [])


def test_strings(seed: int = 11, n_random: int = 70) -> List[str]:
    fixed = [
        "", " ", "  ", "\n", " \n", "a", "Hello world", " Hello world", "Hello  world ", "Hello world  ",
        "I'm sure you're right, they've said it'll work and he'd agree; DON'T I'M YOU'RE WE'VE 'tis 'Sup x'y",
        "1 12 123 1234 12345 1234567890 3.14159 1,000,000 ٣٤٥ ①②③④ x1y22z333",
        "tabs\tand\t\tmore   spaces \n\n\n  indented\r\n\r\nwindows \r mac\n", "trailing space ", "trailing newline\n", "   ",
        "naïve café résumé Ångström über straße Ελληνικά русский язык עברית العربية 日本語のテキスト 中文文本 한국어",
        "Emoji 😀🎉👍🏽 and symbols ±×÷ €£¥ → ∑∫ ©®™ — “quotes” … ž",
        "<|user|>hi<|assistant|> there<|im_end|>", "<|im_start|>user\nHello<|im_end|>\n<|im_start|>assistant\n",
        "<|begin_of_text|><|start_header_id|>user<|end_header_id|>\n\nWhat's 2+2?<|eot_id|><|start_header_id|>assistant<|end_header_id|>\n\n",
        "text <tool> call <tool><tool> end<|eom_id|>", "<|eot_id|>", "<s>not special in bpe</s>", "a<unk>b<pad>c", "<|user|", "|user|>",
        "for (int i = 0; i < n; i++) { x[i] += y[i] * 2; } // code\n#include <stdio.h>\n",
        "http://example.com/path?q=1&r=2 user@mail.org #hash-tag __init__ snake_case CamelCase",
        "The tokenizer's tokenization of tokens: Whitespace tokenization.", " nbsp em　ideographic ls",
        "é combining, ​zero width, ﻿bom, \U0002ebf0 ext-I, \U0001f1fa\U0001f1f8 flag", "!!!???...,,,;;;:::", "' '' ''' 's 't",
    ]
    rng = np.random.default_rng(seed)
    pools = [list("abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"), list("0123456789"), list(" \t\n\r"), list(" "),
             list(".,;:!?'\"()[]{}<>|/\\-_=+*&^%$#@~`"), list("äöüßéèêñçøåæœ"), list("αβγδεζηθικλμ"), list("абвгдежзийкл"),
             list("日本語中文한국어のテキスト"), list("😀🎉👍❤✨"), list("  　 "), ["<|user|>", "<|eot_id|>", "<tool>", "<s>", "'s", "'re", "'LL"]]
    weights = np.array([8, 3, 2, 5, 3, 1, 1, 1, 1, 0.5, 0.3, 0.6])
    weights = weights / weights.sum()
    out = list(fixed)
    for _ in range(n_random):
        n = int(rng.integers(1, 48))
        s = []
        for _ in range(n):
            pool = pools[int(rng.choice(len(pools), p=weights))]
            run = int(rng.integers(1, 5))
            s += [pool[int(rng.integers(len(pool)))] for _ in range(run)]
        out.append("".join(s))
    return out
