"""Known-answer tests of the reference's own spec (``core.spec.ts``), written once
and run against the oracle (CPU) and the product (GPU, through the C-ABI).

Each function cites the ``core.spec.ts`` lines whose golden values it carries.
``make`` is a zero-argument factory returning a tokenizer with the reference's
method names; ``compact_merge`` is the module-level ``compactMerge``.
"""
from __future__ import annotations

import json
import os

EOF = chr(4)
CONTENT_ABC = "aaabdaaabac"  # core.spec.ts:17
CONTENT_X = "xxxxxxxxx"  # core.spec.ts:91


def wrap(content: str) -> str:  # core.spec.ts:13-15
    return EOF + content + EOF


def token_rows(t):
    return [(x.chars, x.weight, x.original_weight, x.code, x.index) for x in t.token_table]


def kat_abc_segments(make):  # core.spec.ts:19-33
    t = make()
    t.addToCorpus(CONTENT_ABC)
    t.mergeUntil({"min_weight": 2})
    assert " ".join(tok.chars for tok in t.encodeToTokens(CONTENT_ABC)) == "aaab d aaab a c"
    # merge order stated in the comment at core.spec.ts:20-23
    assert [c.chars for _, _, c in t.merge_tokens] == ["aa", "ab", "aaab"]


def kat_abc_vector(make):  # core.spec.ts:35-88
    t = make()
    t.addToCorpus(wrap(CONTENT_ABC))
    t.mergeUntil({"min_weight": 2})
    t.compactVectorIndex()
    assert list(t.encodeToVector(CONTENT_ABC)) == [4, 2, 4, 1, 3]
    # weight trace, core.spec.ts:63-71
    assert [(x.chars, x.weight) for x in t.token_table] == [
        (EOF, 2), ("a", 1), ("b", 0), ("d", 1), ("c", 1), ("aa", 0), ("ab", 0), ("aaab", 2),
    ]


def kat_x_segments(make):  # core.spec.ts:93-106
    t = make()
    t.addToCorpus(wrap(CONTENT_X))
    t.mergeUntil({"min_weight": 2})
    assert " ".join(tok.chars for tok in t.encodeToTokens(CONTENT_X)) == "xxxx xxxx x"


def kat_x_vector(make):  # core.spec.ts:108-139
    t = make()
    t.addToCorpus(wrap(CONTENT_X))
    t.mergeUntil({"min_weight": 2})
    t.compactVectorIndex()
    assert list(t.encodeToVector(CONTENT_X)) == [2, 2, 1]
    assert [(x.chars, x.weight) for x in t.token_table] == [(EOF, 2), ("x", 1), ("xx", 0), ("xxxx", 2)]


def kat_json_round_trip(make):  # core.spec.ts:142-165 (the spec trains on its own source text)
    with open(os.path.abspath(__file__), encoding="utf-8") as f:
        text = f.read()
    t = make()
    t.addToCorpus(wrap(text))
    t.mergeUntil({"min_weight": 2})
    s = json.dumps(t.toJSON())
    assert len(s) > 0 and len(t.token_table) > 1
    u = make()
    u.fromJSON(json.loads(s))
    assert token_rows(u) == token_rows(t)
    assert u.toJSON() == t.toJSON()
    return t


def kat_merge_log_resume(make, compact_merge):  # core.spec.ts:167-199
    t = make()
    t.addToCorpus(wrap(CONTENT_ABC))
    merges = []
    while True:
        merge = t.findNextMerge()
        if not merge:
            break
        if merge[2].weight < 2:
            break
        merges.append(compact_merge(merge))
        t.applyMerge(merge)
    vector = list(t.encodeToVector(CONTENT_ABC))
    assert len(merges) > 0 and len(vector) > 0
    u = make()
    u.addToCorpus(wrap(CONTENT_ABC))
    for m in merges:
        u.restoreMerge(m)
    assert token_rows(u) == token_rows(t)
    assert list(u.encodeToVector(CONTENT_ABC)) == vector
    assert u.decodeVector(vector) == CONTENT_ABC


def kat_vector_invalidation(make):  # core.spec.ts:201-224
    content = "x" * 10
    t = make()
    t.addToCorpus(wrap(content))
    assert list(t.encodeToVector(content)) == [1] * 10
    t.applyMerge(t.findNextMerge({"max_length": 5}))
    assert list(t.encodeToVector(content)) == [1] * 5
    t.applyMerge(t.findNextMerge({"max_length": 5}))
    assert list(t.encodeToVector(content)) == [2, 2, 1]


def _expect_merge(merge, a, b):  # core.spec.ts:239-244
    assert merge is not None
    assert merge[0].chars == a and merge[1].chars == b and merge[2].chars == a + b


def kat_max_length(make):  # core.spec.ts:226-264
    content = "x" * 10
    t = make()
    t.addToCorpus(wrap(content))
    m = t.findNextMerge()
    _expect_merge(m, "x", "x")
    t.applyMerge(m)
    m = t.findNextMerge({"max_length": 4})
    _expect_merge(m, "xx", "xx")
    t = make()
    t.addToCorpus(wrap(content))
    m = t.findNextMerge()
    _expect_merge(m, "x", "x")
    t.applyMerge(m)
    assert t.findNextMerge({"max_length": 3}) is None


def kat_min_weight(make):  # core.spec.ts:266-313
    content = "x" * 10
    t = make()
    t.addToCorpus(wrap(content))
    _expect_merge(t.findNextMerge({"min_weight": 5}), "x", "x")
    assert t.findNextMerge({"min_weight": 6}) is None
    m = t.findNextMerge()
    _expect_merge(m, "x", "x")
    t.applyMerge(m)
    m = t.findNextMerge()
    _expect_merge(m, "xx", "xx")
    t.applyMerge(m)
    assert t.findNextMerge() is None


def kat_merge_until(make):  # core.spec.ts:315-417
    content = "x" * 10
    full = [(EOF, 2), ("x", 0), ("xx", 1), ("xxxx", 2)]
    short = [(EOF, 2), ("x", 0), ("xx", 5)]
    for options, want in [
        ({"min_weight": 2}, full),
        ({"min_weight": 3}, short),
        ({"max_length": 4}, full),
        ({"max_length": 3}, short),
        ({"min_weight": 3, "max_length": 3}, short),
    ]:
        t = make()
        t.addToCorpus(wrap(content))
        t.mergeUntil(options)
        assert [(x.chars, x.weight) for x in t.token_table] == want, options


def kat_falsy_options(make):  # core.ts:255-256, :272, :373-376 (falsy means default)
    content = "x" * 10
    t = make()
    t.addToCorpus(wrap(content))
    t.mergeUntil({"min_weight": 0, "max_length": 0, "max_iterations": 0})
    assert [(x.chars, x.weight) for x in t.token_table] == [(EOF, 2), ("x", 0), ("xx", 1), ("xxxx", 2)]
    t = make()
    t.addToCorpus(wrap(content))
    t.mergeUntil({"max_iterations": 1})
    assert [(x.chars, x.weight) for x in t.token_table] == [(EOF, 2), ("x", 0), ("xx", 5)]


def kat_error_messages(make):  # core.ts:136, :226-228, :399, :440, :467, :481-483
    import pytest

    t = make()
    with pytest.raises(Exception, match="token table is empty, have you called tokenizer.addToCorpus\\(\\)\\?"):
        t.compactVectorIndex()
    with pytest.raises(Exception, match="invalid format"):
        t.fromJSON({"version": 1, "token_table": [], "merge_codes": []})
    t = make()
    t.addToCorpus(wrap(CONTENT_ABC))
    with pytest.raises(Exception, match='unknown token, char: "z"'):
        t.encodeToVector("az")
    with pytest.raises(Exception, match='unknown token, char: "\\\\n"'):
        t.encodeToCode("a\nz")
    t.mergeUntil({"min_weight": 2})
    # 'b' (index 2) is fully absorbed into 'ab' -> zero weight -> hole in to_vector_index
    with pytest.raises(Exception, match="unknown token index: 2"):
        t.encodeToVector("db")
    with pytest.raises(Exception, match="unknown vector index: 99"):
        t.decodeVector([0, 99])
    with pytest.raises(Exception, match='unknown token, a_code: "\\\\u0063"|unknown token, a_code: "c"'):
        t.restoreMerge(["c", "\x01", 3])
    with pytest.raises(Exception, match="unknown token, b_code"):
        t.restoreMerge(["\x01", "ꯍ", 3])


ALL = [
    kat_abc_segments,
    kat_abc_vector,
    kat_x_segments,
    kat_x_vector,
    kat_json_round_trip,
    kat_vector_invalidation,
    kat_max_length,
    kat_min_weight,
    kat_merge_until,
    kat_falsy_options,
    kat_error_messages,
]
