"""Snapshot compatibility with the reference's sqlite twin: ``toJSON`` output must survive
``BPETokenizerDB.fromJSON`` -> ``toJSON`` unchanged (the reference's own test: db/core.spec.ts:32-41)."""
import json

import pytest

from oracle import LiteralTokenizer
from oracle.db_snapshot import DBSnapshot

EOF_ = chr(4)


def _as_lists(j):
    return json.loads(json.dumps(j))


def test_db_round_trip_of_the_reference_example_oracle():  # db/core.spec.ts:32-41, in-memory side = the literal oracle
    t = LiteralTokenizer()
    t.addToCorpus(EOF_ + "aaabdaaabac" + EOF_)
    t.mergeUntil({"min_weight": 2})
    db = DBSnapshot()
    db.from_json(t.toJSON())
    assert db.to_json() == _as_lists(t.toJSON())


def test_db_rejects_what_the_reference_rejects():
    db = DBSnapshot()
    for bad in ({}, {"version": 1, "token_table": [], "merge_codes": []}, {"version": 2, "token_table": {}, "merge_codes": []}):
        with pytest.raises(ValueError, match="invalid format"):
            db.from_json(bad)


@pytest.mark.gpu
def test_db_imports_gpu_snapshots_unchanged():
    from bpe_tokenizer_b200 import BPETokenizer

    for docs, opts in (([EOF_ + "aaabdaaabac" + EOF_], {"min_weight": 2}),
                       (["the cat sat on the mat\n" * 20, "café \U0001F600 naïve " * 15, ""], {"max_length": 6}),
                       (["x" * 300, "xyxyxy" * 50], {})):
        t = BPETokenizer()
        for d in docs:
            t.addToCorpus(d)
        t.mergeUntil(opts)
        snap = t.toJSON()
        db = DBSnapshot()
        db.from_json(snap)
        assert db.to_json() == _as_lists(snap)
        back = BPETokenizer()
        back.fromJSON(db.to_json())  # and the database's export loads back into the engine
        assert back.toJSON() == snap
        for d in docs:
            assert back.encodeToCode(d) == t.encodeToCode(d)
