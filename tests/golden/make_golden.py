"""Regenerates tests/golden/literal_cases.json from the string-level literal
restatement of core.ts (oracle/ref_literal.py).  The reference itself cannot run
here (no Node in the image), so these vectors are outputs of the restatement,
which is pinned by the core.spec.ts known answers in tests/kat_suite.py.

    python tests/golden/make_golden.py
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import LiteralTokenizer  # noqa: E402
from bpe_tokenizer_b200.synth import synth_corpus  # noqa: E402

EOF_ = chr(4)


def case(name, docs, options, texts):
    t = LiteralTokenizer()
    for d in docs:
        t.addToCorpus(d)
    t.mergeUntil(options)
    enc = []
    for text in texts:
        try:
            enc.append([text, t.encodeToVector(text)])
        except ValueError as e:
            enc.append([text, "throws: " + str(e)])
    return {
        "name": name,
        "docs": docs,
        "options": options,
        "json": t.toJSON(),
        "merges": [[a.index, b.index, c.weight] for a, b, c in t.merge_tokens],
        "encode": enc,
    }


def main():
    rng = random.Random(2026)
    cases = [
        case("readme", ["aaabdaaabac"], {"min_weight": 2}, ["aaabdaaabac", "abc", "daab"]),
        case("readme_wrapped", [EOF_ + "aaabdaaabac" + EOF_], {"min_weight": 2}, ["aaabdaaabac", "db"]),
        case("runs_9x", [EOF_ + "x" * 9 + EOF_], {"min_weight": 2}, ["x" * 9, "x" * 4, "x" * 7]),
        case("runs_mixed", ["aaaa", "aaa", "aaaaa", "baaab", "aabaa"], {}, ["aaaaaaa", "baab", "abab"]),
        case("tie_index_sum", ["abcd", "badc", "cdab", "dcba"], {}, ["abcd", "dcba", "acbd"]),
        case("tie_last_position", ["ab", "ba", "ab", "ba", "cd", "dc"], {}, ["abba", "cddc"]),
        case("multi_doc_isolation", ["ab", "ab", "a", "b", "", "ba"], {}, ["abab", "ba"]),
        case("chains", ["abababab", "ababab", "bababa"], {}, ["ababababab", "bab"]),
        case("max_length_3", ["the cat and the hat and the bat"] * 3, {"max_length": 3}, ["the cat", "that"]),
        case("min_weight_4", ["the cat and the hat and the bat"] * 3, {"min_weight": 4}, ["the cat", "that"]),
        case("astral", ["\U0001F600\U0001F600ab\U0001F600\U0001F600ab"], {"max_length": 4}, ["\U0001F600ab"]),
        case("max_iterations_5", ["she sells sea shells by the sea shore"] * 2, {"max_iterations": 5}, ["sea shells"]),
    ]
    for i in range(6):
        alphabet = "abc"[: 1 + i % 3]
        docs = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 30))) for _ in range(1 + i % 4)]
        cases.append(case(f"fuzz_{i}", docs, {"max_length": [0, 4, 8][i % 3]}, docs[:2] + ["".join(alphabet) * 3]))
    text, off = synth_corpus(6000)
    docs = [bytes(text[off[d]:off[d + 1]]).decode() for d in range(len(off) - 1)]
    cases.append(case("zipf_6k", docs, {"max_iterations": 300}, docs[:3] + ["\rzzz qqq\n"]))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "literal_cases.json")
    with open(out, "w", encoding="utf-8") as f:
        json.dump(cases, f, ensure_ascii=True, indent=0)
    print("wrote", out, len(cases), "cases")


if __name__ == "__main__":
    main()
