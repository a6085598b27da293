"""CPU: the pure host-side helpers of the product (SURVEY.md section 8(a) row a14 and the option / message conventions of 8(b)) against
the reference's definitions (core.ts:36-75, :255-256, :399, :500-503) and against the literal oracle's restatement of the same lines.
Nothing here needs the CUDA library: the helpers are module-level functions of bpe_tokenizer_b200/tokenizer.py."""
import json
import unicodedata

import pytest

from bpe_tokenizer_b200 import tokenizer as T
from oracle import ref_literal as R


def test_marker_constants():  # core.ts:36-45
    assert (T.FS, T.EOF, T.LF, T.CR) == (chr(28), chr(4), "\n", "\r")
    assert (T.FS, T.EOF) == (R.FS, R.EOF)


def test_file_content_to_corpus():  # core.ts:55-58: `content.toString()` wrapped in FS .. EOF
    assert T.fileContentToCorpus("abc") == "\x1cabc\x04"
    assert T.fileContentToCorpus("") == "\x1c\x04"
    assert T.fileContentToCorpus("héllo 中".encode("utf-8")) == "\x1chéllo 中\x04"  # a Buffer decodes as UTF-8
    assert T.fileContentToCorpus("a\nb") == R.file_content_to_corpus("a\nb")


JS_WHITESPACE = "\t\n\x0b\x0c\r \xa0                　﻿"


def test_lines_to_corpus_trims_like_string_prototype_trim():  # core.ts:61-64
    assert T.linesToCorpus(" a \r\nb\t\n\n") == ["\ra\n", "\rb\n", "\r\n", "\r\n"]
    assert T.linesToCorpus("") == ["\r\n"]  # ''.split('\n') is ['']
    # ECMA-262 WhiteSpace + LineTerminator: every one of them is trimmed, nothing else is (U+0085, U+180E, U+200B stay)
    for ws in JS_WHITESPACE:
        if ws == "\n":
            continue
        assert T.linesToCorpus(ws + "x" + ws) == ["\rx\n"], hex(ord(ws))
    for keep in "\x85᠎​\x00\x1c":
        assert T.linesToCorpus(keep + "x" + keep) == ["\r" + keep + "x" + keep + "\n"], hex(ord(keep))
    # the set is exactly Unicode Zs + the six controls + BOM + the two separators
    zs = {chr(c) for c in range(0x3100) if unicodedata.category(chr(c)) == "Zs"}
    assert zs | set("\t\n\x0b\x0c\r﻿  ") == set(JS_WHITESPACE)
    text = "  first line \r\n\tsecond \n\nlast﻿"
    assert T.linesToCorpus(text) == R.lines_to_corpus(text)


def test_lines_trimmed_to_corpus_only_strips_one_carriage_return():  # core.ts:67-75 (the names are swapped relative to behaviour)
    assert T.linesTrimmedToCorpus(" a \r\nb\r\r\n c") == ["\r a \n", "\rb\r\n", "\r c\n"]
    assert T.linesTrimmedToCorpus("") == ["\r\n"]
    text = "  first line \r\n\tsecond \n\nlast\r"
    assert T.linesTrimmedToCorpus(text) == R.lines_trimmed_to_corpus(text)


@pytest.mark.parametrize("s", ["a", '"', "\\", "\n\r\t\b\f", "\x00\x01\x1f", "\x7f", "é中\U0001F600", "  ", "a\"b\\c"])
def test_js_stringify_matches_json_stringify(s):  # the text of `unknown token, char: ${JSON.stringify(char)}` (core.ts:399)
    # JSON.stringify escapes only ", \\, the C0 controls (short forms for \b \f \n \r \t) and lone surrogates
    assert T._js_stringify(s) == json.dumps(s, ensure_ascii=False)
    assert T._js_stringify(s) == R.js_stringify(s)


def test_js_stringify_lone_surrogates():  # well-formed JSON.stringify (ES2019): \\udXXX in lower-case hex
    assert T._js_stringify("\ud800") == '"\\ud800"'
    assert T._js_stringify("a\udfffb") == '"a\\udfffb"'
    assert T._js_stringify("\ud800") == R.js_stringify("\ud800")


def test_utf16_length():  # `chars.length` of core.ts:272 counts UTF-16 units
    assert T._utf16_len("") == 0 and T._utf16_len("abc") == 3 and T._utf16_len("é中") == 2
    assert T._utf16_len("\U0001F600") == 2 and T._utf16_len("a\U0001F600b\U00010000") == 6
    assert T._utf16_len("￿") == 1


def test_compact_merge_and_token_rows():  # core.ts:500-503, :1-10
    a, b = T.Token("a", 5, 7, chr(1), 0), T.Token("b", 4, 4, chr(2), 1)
    c = T.Token("ab", 3, 3, chr(3), 2)
    assert T.compactMerge((a, b, c)) == [chr(1), chr(2), 3]
    assert T.compactMerge((a, b, c)) == R.compact_merge((a, b, c))
    assert a == T.Token("a", 5, 7, chr(1), 0) and a != b  # value equality, identity hash (tokens are mutable records)
    assert len({a, T.Token("a", 5, 7, chr(1), 0)}) == 2


@pytest.mark.parametrize("given,want", [(None, 2), (0, 2), (False, 2), (1, 1), (2, 2), (5, 5), (2.5, 3), (0.5, 1)])
def test_min_weight_option_falsy_means_two(given, want):  # core.ts:256 `options.min_weight || 2`; counts are integers, so a
    assert T._min_weight({"min_weight": given}) == want   # fractional bound w is `count >= ceil(w)`
    assert T._min_weight({}) == 2


def test_options_merge_keyword_arguments():
    assert T._options({"min_weight": 3}, {"max_length": 5}) == {"min_weight": 3, "max_length": 5}
    assert T._options(None, {}) == {}
    assert T._options({"min_weight": 3}, {"min_weight": 4}) == {"min_weight": 4}
