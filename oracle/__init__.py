"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the BPE train/encode hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or
as the timed CPU baseline), never as the thing shipped.

Parity status: PINNED against every known-answer test of the reference's own
``core.spec.ts`` (see ``tests/test_oracle_golden.py``).  Behaviours the
reference's tests do not pin (tie-break among pairs of equal weight and
equal index sum, multi-document corpora, astral input, the throw paths) are
"parity unpinned": for those the oracle is the literal restatement below,
cross-checked by fuzzing between the string-level and the int-level forms.
"""
from .ref_literal import (  # noqa: F401
    LiteralTokenizer,
    Token,
    compact_merge,
    file_content_to_corpus,
    lines_to_corpus,
    lines_trimmed_to_corpus,
    js_stringify,
    utf16_len,
    FS,
    EOF,
    LF,
    CR,
)
