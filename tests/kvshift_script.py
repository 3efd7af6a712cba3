"""The context-shift scenario of SURVEY §8 f-3, written once and run against any runner with the llama.h-shaped interface
(kv_clear / decode(tokens, pos0) -> logits / kv_seq_rm / kv_seq_add): the reference (oracle/ref.py RefModel, which makes
tests/golden/kvshift_*.npz) and the CUDA engine (booster_b200.engine.Context).

It follows cpp/bridge.cpp:487-507: when the context is (nearly) full, keep n_keep tokens, drop half of the rest
(llama_kv_cache_seq_rm), move the tail down (llama_kv_cache_seq_add) and carry on at the reduced n_past. New tokens then fill
the freed cells IN THE MIDDLE of the cache, so cell order != position order; a second shift moves cells that were already
moved once and cells written after the first shift; a batch of several tokens is placed by one find_slot."""
import numpy as np


def run(r, prompt, n_ctx=64, n_keep=4):
    logits = []
    r.kv_clear()
    lg = r.decode(prompt, 0)                       # one batch
    logits.append(lg)
    n_past = len(prompt)

    def greedy(n):
        nonlocal lg, n_past
        for _ in range(n):
            t = int(np.argmax(lg))
            lg = r.decode([t], n_past)
            logits.append(lg)
            n_past += 1

    def shift():
        nonlocal n_past
        n_left = n_past - n_keep                   # cpp/bridge.cpp:495-503
        n_discard = n_left // 2
        r.kv_seq_rm(n_keep, n_keep + n_discard)
        r.kv_seq_add(n_keep + n_discard, n_past, -n_discard)
        n_past -= n_discard

    greedy(n_ctx - 6 - len(prompt))                # up to the bridge's n_ctx - 4 limit, roughly
    shift()
    greedy(12)
    lg = r.decode([5, 9, 200, 17, 3], n_past)      # a batch after the shift: one find_slot for 5 cells
    logits.append(lg)
    n_past += 5
    greedy(n_ctx - 6 - n_past)
    shift()                                        # second shift: cells moved once, and cells written in between
    greedy(10)
    return np.stack(logits)


def run_self_extend(r, prompt, ga_n=2, ga_w=16, n_batch=8, n_gen=28):
    """Self-Extend (cpp/bridge.cpp:509-524, grp_attn_n = ga_n, grp_attn_w = ga_w): before EVERY llama_decode — prompt chunks of
    n_batch tokens and generated tokens alike — whole windows of ga_w positions are compressed by ga_n
    (llama_kv_cache_seq_add / seq_div / seq_add) and n_past falls back accordingly. Cells never move or free; several cells end
    up with the same position; every window is re-rotated by a different per-cell delta at the next decode."""
    logits = []
    r.kv_clear()
    n_past, ga_i = 0, 0

    def extend():
        nonlocal n_past, ga_i
        while n_past >= ga_i + ga_w:
            ib = (ga_n * ga_i) // ga_w
            bd = (ga_w // ga_n) * (ga_n - 1)
            dd = (ga_w // ga_n) - ib * bd - ga_w
            r.kv_seq_add(ga_i, n_past, ib * bd)
            r.kv_seq_div(ga_i + ib * bd, ga_i + ib * bd + ga_w, ga_n)
            r.kv_seq_add(ga_i + ib * bd + ga_w, n_past + ib * bd, dd)
            n_past -= bd
            ga_i += ga_w // ga_n

    lg = None
    for i in range(0, len(prompt), n_batch):
        chunk = prompt[i:i + n_batch]
        extend()
        lg = r.decode(chunk, n_past)
        logits.append(lg)
        n_past += len(chunk)
    for _ in range(n_gen):
        t = int(np.argmax(lg))
        extend()
        lg = r.decode([t], n_past)
        logits.append(lg)
        n_past += 1
    return np.stack(logits)
