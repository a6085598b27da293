"""TEST INFRASTRUCTURE ONLY -- literal CPU restatement of the reference's
in-memory tokenizer (``/root/reference/core.ts``), string level.

Every method cites the ``core.ts`` lines it follows.  The data shapes are kept
(one "code" character per token, ``code = chr(index + 1)``; corpus held as one
code string per document; a fresh two-level count map per ``find_next_merge``;
the *running* arg-max inside the counting loop; ``str.replace`` standing in
for ``String.prototype.replaceAll`` -- both are left-to-right, non-overlapping).

JS strings are UTF-16, Python strings are code points.  The reference iterates
strings by code point (``for (x of s)``), which is what iterating a Python
``str`` does; the only place UTF-16 shows through is ``chars.length`` in the
``max_length`` filter (core.ts:272), restated by :func:`utf16_len`.

Pure-Python loops: use for small cases only (see ``int_oracle.py`` for the
compiled int-level form used on MB-scale inputs).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

# core.ts:36-45
FS = chr(28)
EOF = chr(4)
LF = "\n"
CR = "\r"


def utf16_len(s: str) -> int:
    """``s.length`` as JS sees it (UTF-16 code units)."""
    return sum(2 if ord(ch) > 0xFFFF else 1 for ch in s)


def js_stringify(s: str) -> str:
    """``JSON.stringify`` of a JS string (well-formed variant, ES2019)."""
    out = ['"']
    short = {'"': '\\"', "\\": "\\\\", "\b": "\\b", "\f": "\\f", "\n": "\\n", "\r": "\\r", "\t": "\\t"}
    for ch in s:
        o = ord(ch)
        if ch in short:
            out.append(short[ch])
        elif o < 0x20 or 0xD800 <= o <= 0xDFFF:
            out.append("\\u%04x" % o)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


# core.ts:55-58
def file_content_to_corpus(content) -> str:
    text = content.decode("utf-8") if isinstance(content, (bytes, bytearray)) else str(content)
    return FS + text + EOF


_JS_WS = (
    "\t\n\x0b\x0c\r \xa0\u1680\u2000\u2001\u2002\u2003\u2004\u2005\u2006\u2007\u2008\u2009\u200a"
    "\u2028\u2029\u202f\u205f\u3000\ufeff"
)


# core.ts:61-64  (String.prototype.trim strips the ECMAScript WhiteSpace+LineTerminator set)
def lines_to_corpus(text: str) -> List[str]:
    return ["\r" + line.strip(_JS_WS) + "\n" for line in text.split("\n")]


# core.ts:67-75
def lines_trimmed_to_corpus(text: str) -> List[str]:
    out = []
    for line in text.split("\n"):
        if line.endswith("\r"):
            line = line[: len(line) - 1]
        out.append("\r" + line + "\n")
    return out


@dataclass
class Token:  # core.ts:1-10
    chars: str
    weight: int
    original_weight: int
    code: str
    index: int


MergeToken = Tuple[Token, Token, Token]


# core.ts:500-503
def compact_merge(merge: MergeToken):
    a, b, c = merge
    return [a.code, b.code, c.weight]


class LiteralTokenizer:
    """core.ts:77-495, restated line for line."""

    def __init__(self):
        self.char_to_token: Dict[str, Token] = {}
        self.code_to_token: Dict[str, Token] = {}
        self.token_table: List[Token] = []
        self.merge_tokens: List[MergeToken] = []
        self.merge_codes: List[Tuple[str, str]] = []
        self.to_vector_index: Optional[Dict[int, int]] = None  # sparse array
        self.from_vector_index: Optional[Dict[int, int]] = None
        self.corpus_in_code: List[str] = []

    # core.ts:112-127
    def to_json(self) -> dict:
        return {
            "version": 2,
            "char_count": len(self.char_to_token),
            "token_table": [[t.chars, t.weight, t.original_weight] for t in self.token_table],
            "merge_codes": [[a.code, b.code, c.code] for (a, b, c) in self.merge_tokens],
        }

    # core.ts:130-171
    def from_json(self, json: dict) -> None:
        if (
            not isinstance(json, dict)
            or json.get("version") != 2
            or not isinstance(json.get("token_table"), list)
            or not isinstance(json.get("merge_codes"), list)
        ):
            raise ValueError("invalid format")
        char_count = json.get("char_count")
        self.__init__()
        for chars, weight, original_weight in json["token_table"]:
            index = len(self.token_table)
            code = chr(index + 1)
            token = Token(chars, weight, original_weight, code, index)
            if char_count is not None and index < char_count:
                self.char_to_token[chars] = token
            self.code_to_token[code] = token
            self.token_table.append(token)
        for a_code, b_code, c_code in json["merge_codes"]:
            a = self.code_to_token[a_code]
            b = self.code_to_token[b_code]
            c = self.code_to_token[c_code]
            self.merge_tokens.append((a, b, c))
            self.merge_codes.append((a.code + b.code, c.code))
        self.compact_vector_index()

    # core.ts:173-176
    def _invalidate_vector_index(self) -> None:
        self.to_vector_index = None
        self.from_vector_index = None

    # core.ts:182-207
    def add_to_corpus(self, content: str) -> None:
        sample_in_code = []
        for char in content:
            token = self.char_to_token.get(char)
            if token is None:
                index = len(self.token_table)
                code = chr(index + 1)
                token = Token(char, 1, 1, code, index)
                self.char_to_token[char] = token
                self.code_to_token[code] = token
                self.token_table.append(token)
            else:
                token.weight += 1
                token.original_weight += 1
            sample_in_code.append(token.code)
        self.corpus_in_code.append("".join(sample_in_code))

    # core.ts:213-216
    def restore_to_corpus(self, content: str) -> None:
        self.corpus_in_code.append(self.encode_to_code(content))

    # core.ts:222-241
    def compact_vector_index(self) -> None:
        if len(self.token_table) == 0:
            raise ValueError("token table is empty, have you called tokenizer.addToCorpus()?")
        self.to_vector_index = {}
        self.from_vector_index = {}
        vector_index = 0
        for index, token in enumerate(self.token_table):
            if token.weight > 0:
                self.to_vector_index[index] = vector_index
                self.from_vector_index[vector_index] = index
                vector_index += 1

    # core.ts:247-326
    def find_next_merge(self, min_weight=None, max_length=None) -> Optional[MergeToken]:
        min_weight = min_weight or 2  # core.ts:256 (falsy -> 2)
        a_b_c_weights: Dict[int, Dict[int, int]] = {}
        max_a = max_b = None
        max_c_index = None
        max_c_weight = None
        code_to_token = self.code_to_token
        for sample_in_code in self.corpus_in_code:  # core.ts:265
            last_a = None
            a = None
            for code in sample_in_code:  # core.ts:268
                b = code_to_token[code]
                # core.ts:270-273
                if a is not None and (not max_length or utf16_len(a.chars) + utf16_len(b.chars) <= max_length):
                    b_c_weights = a_b_c_weights.get(a.index)
                    if b_c_weights is None:
                        b_c_weights = {}
                        a_b_c_weights[a.index] = b_c_weights
                    c_weight = b_c_weights.get(b.index)
                    if not c_weight:  # core.ts:281-283
                        b_c_weights[b.index] = 1
                        c_weight = 1
                    else:
                        # core.ts:285-290: "X X X" is one occurrence of "X X"
                        if a is b and last_a is a:
                            last_a = None
                            a = b
                            continue
                        c_weight += 1
                        b_c_weights[b.index] = c_weight
                    c_index = a.index + b.index
                    # core.ts:296-305 running arg-max
                    if (not max_c_weight) or c_weight > max_c_weight or (
                        c_weight == max_c_weight and c_index < max_c_index
                    ):
                        max_a = a
                        max_b = b
                        max_c_weight = c_weight
                        max_c_index = c_index
                last_a = a  # core.ts:307-308
                a = b
        if not max_c_weight:  # core.ts:312
            return None
        if min_weight and max_c_weight < min_weight:  # core.ts:313
            return None
        new_index = len(self.token_table)  # core.ts:315-325
        max_c = Token(max_a.chars + max_b.chars, max_c_weight, max_c_weight, chr(new_index + 1), new_index)
        return (max_a, max_b, max_c)

    # core.ts:332-360
    def apply_merge(self, merge: MergeToken) -> None:
        a, b, c = merge
        from_code = a.code + b.code
        to_code = c.code
        a.weight -= c.weight
        b.weight -= c.weight
        self._invalidate_vector_index()
        self.code_to_token[c.code] = c
        self.token_table.append(c)
        self.merge_tokens.append(merge)
        self.merge_codes.append((from_code, to_code))
        corpus = self.corpus_in_code
        for i in range(len(corpus)):
            corpus[i] = corpus[i].replace(from_code, to_code)

    # core.ts:365-383
    def merge_until(self, min_weight=None, max_length=None, max_iterations=None) -> int:
        iteration = 1
        done = 0
        while (not max_iterations) or iteration <= max_iterations:
            merge = self.find_next_merge(min_weight=min_weight, max_length=max_length)
            if not merge:
                break
            self.apply_merge(merge)
            done += 1
            iteration += 1
        return done

    # core.ts:392-409
    def encode_to_code(self, content: str) -> str:
        parts = []
        for char in content:
            token = self.char_to_token.get(char)
            if token is None:
                raise ValueError("unknown token, char: " + js_stringify(char))
            parts.append(token.code)
        content_in_code = "".join(parts)
        for from_code, to_code in self.merge_codes:
            content_in_code = content_in_code.replace(from_code, to_code)
        return content_in_code

    # core.ts:411-422
    def encode_to_tokens(self, content: str) -> List[Token]:
        return [self.code_to_token[code] for code in self.encode_to_code(content)]

    # core.ts:424-445
    def encode_to_vector(self, content: str) -> List[int]:
        if self.to_vector_index is None:
            self.compact_vector_index()
        to_vector_index = self.to_vector_index
        vector = []
        for code in self.encode_to_code(content):
            index = self.code_to_token[code].index
            if index in to_vector_index:
                vector.append(to_vector_index[index])
            else:
                raise ValueError(f"unknown token index: {index}")
        return vector

    # core.ts:447-453
    def decode_tokens(self, tokens: List[Token]) -> str:
        return "".join(t.chars for t in tokens)

    # core.ts:455-471
    def decode_vector(self, vector: List[int]) -> str:
        if self.from_vector_index is None:
            self.compact_vector_index()
        out = []
        for vector_index in vector:
            if vector_index in self.from_vector_index:
                out.append(self.token_table[self.from_vector_index[vector_index]].chars)
            else:
                raise ValueError(f"unknown vector index: {vector_index}")
        return "".join(out)

    # core.ts:477-494
    def restore_merge(self, compact) -> None:
        a_code, b_code, c_weight = compact
        a = self.code_to_token.get(a_code)
        if a is None:
            raise ValueError(f"unknown token, a_code: {js_stringify(a_code)}")
        b = self.code_to_token.get(b_code)
        if b is None:
            raise ValueError(f"unknown token, b_code: {js_stringify(b_code)}")
        index = len(self.token_table)
        c = Token(a.chars + b.chars, c_weight, c_weight, chr(index + 1), index)
        self.apply_merge((a, b, c))


def _opts(options, kw):
    o = dict(options or {})
    o.update(kw)
    return o


def _install_reference_names(cls):
    """camelCase aliases with the reference's option-object calling convention,
    so one known-answer suite (tests/kat_suite.py) drives oracle and product alike."""
    cls.toJSON = lambda self: self.to_json()
    cls.fromJSON = lambda self, json: self.from_json(json)
    cls.addToCorpus = lambda self, content: self.add_to_corpus(content)
    cls.restoreToCorpus = lambda self, content: self.restore_to_corpus(content)
    cls.compactVectorIndex = lambda self: self.compact_vector_index()
    cls.findNextMerge = lambda self, options=None, **kw: self.find_next_merge(
        **{k: v for k, v in _opts(options, kw).items() if k in ("min_weight", "max_length")}
    )
    cls.applyMerge = lambda self, merge: self.apply_merge(merge)
    cls.mergeUntil = lambda self, options=None, **kw: self.merge_until(**_opts(options, kw))
    cls.encodeToCode = lambda self, content: self.encode_to_code(content)
    cls.encodeToTokens = lambda self, content: self.encode_to_tokens(content)
    cls.encodeToVector = lambda self, content: self.encode_to_vector(content)
    cls.decodeTokens = lambda self, tokens: self.decode_tokens(tokens)
    cls.decodeVector = lambda self, vector: self.decode_vector(vector)
    cls.restoreMerge = lambda self, compact: self.restore_merge(compact)
    return cls


_install_reference_names(LiteralTokenizer)
