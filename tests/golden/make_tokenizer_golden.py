"""Regenerate the tokenizer golden fixtures from the REAL reference tokenizer (oracle/_ref, built from /root/reference):
    python tests/golden/make_tokenizer_golden.py
writes tests/golden/vocab_{spm,bpe}.gguf (vocab-only synthetic GGUFs, tests/tokenizer_fixtures.py) and
tests/golden/tokenizer_{spm,bpe}.json = for every test string the reference's llama_tokenize output for
(add_special, parse_special) in {0,1}^2, plus llama_token_to_piece (special = 1 and 0) and llama_token_is_eog of every id."""
import base64
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import tokenizer_fixtures as F  # noqa: E402
from oracle import ref  # noqa: E402

for kind in ("spm", "bpe"):
    path = os.path.join(HERE, f"vocab_{kind}.gguf")
    F.write_vocab_gguf(path, kind)
    rv = ref.RefVocab(path)
    cases = []
    for s in F.test_strings():
        b = s.encode("utf-8")
        cases.append({"text_b64": base64.b64encode(b).decode(),
                      "ids": {f"{int(a)}{int(p)}": rv.tokenize(b, a, p) for a in (False, True) for p in (False, True)}})
    n_vocab = len(F.spm_vocab()[0] if kind == "spm" else F.bpe_vocab()[0])
    pieces = [[base64.b64encode(rv.piece(i, True)).decode(), base64.b64encode(rv.piece(i, False)).decode(), int(rv.is_eog(i))]
              for i in range(n_vocab)]
    json.dump({"source": "oracle/_ref (" + str(ref.variant()) + "): llama_tokenize / llama_token_to_piece / llama_token_is_eog",
               "cases": cases, "pieces": pieces}, open(os.path.join(HERE, f"tokenizer_{kind}.json"), "w"))
    rv.close()
    print(kind, len(cases), "cases,", n_vocab, "tokens")
