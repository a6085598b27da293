"""The Node binding of INTEGRATION.md exists as source (bindings/node/): Node, node_api.h and tsc are absent from this image,
so the addon is type-checked against a stand-in header and both files are checked against the C ABI and against the
reference's public surface.  CPU only; nothing is executed."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADDON = os.path.join(ROOT, "bindings", "node", "bpe_b200_napi.c")
TS = os.path.join(ROOT, "bindings", "node", "bpe_tokenizer.ts")
HEADER = os.path.join(ROOT, "include", "bpe_b200.h")


def _read(path):
    with open(path, encoding="utf-8") as f:
        return f.read()


def _strip_c_comments(src):
    return re.sub(r"/\*.*?\*/", " ", src, flags=re.S)


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_addon_type_checks_against_the_abi_header():
    """-Wall -Wextra -Werror -fsyntax-only: every bpe_* call of the addon has the argument types include/bpe_b200.h declares."""
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "tests", "mock_node_api"),
                        "-I", os.path.join(ROOT, "include"), ADDON], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_addon_calls_only_declared_entry_points_and_covers_the_path():
    declared = set(re.findall(r"\b(bpe_[a-z0-9_]+)\s*\(", _strip_c_comments(_read(HEADER))))
    called = set(re.findall(r"\b(bpe_[a-z0-9_]+)\s*\(", _strip_c_comments(_read(ADDON))))
    assert called <= declared, called - declared
    # the entry points behind the reference's members
    needed = {"bpe_create", "bpe_destroy", "bpe_last_error", "bpe_set_tokens", "bpe_num_tokens", "bpe_load_merges", "bpe_add_documents",
              "bpe_restore_documents", "bpe_clear_corpus", "bpe_corpus_size", "bpe_get_corpus", "bpe_find_next_merge", "bpe_apply_merge",
              "bpe_apply_merges", "bpe_merge_until", "bpe_encode_batch", "bpe_encode_text_batch", "bpe_decode_batch", "bpe_set_chars",
              "bpe_add_text"}
    assert needed <= called, needed - called


def test_typescript_class_calls_only_what_the_addon_exports():
    addon = _read(ADDON)
    exported = set(re.findall(r'\{"([A-Za-z]+)",\s*[A-Za-z]+\}', addon)) | {"MAX_TOKENS", "ABI_VERSION"}
    ts = _read(TS)
    used = set(re.findall(r"\bnative\.([A-Za-z_]+)", ts))
    assert used <= exported, used - exported
    declared = set(re.findall(r"^  ([A-Za-z_]+)[(:]", ts[ts.index("interface Native"):ts.index("/** core.ts:1-10 */")], flags=re.M))
    assert declared == exported, declared ^ exported


def test_typescript_class_has_the_reference_surface():
    """Every public member of core.ts's class and every module export a caller can name."""
    ts = _read(TS)
    cls = ts[ts.index("export class BPETokenizer"):]
    for member in ["char_to_token", "code_to_token", "token_table", "merge_tokens", "merge_codes", "to_vector_index", "from_vector_index",
                   "get corpus_in_code", "set corpus_in_code", "toJSON", "fromJSON", "addToCorpus", "restoreToCorpus", "compactVectorIndex",
                   "findNextMerge", "applyMerge", "mergeUntil", "encodeToCode", "encodeToTokens", "encodeToVector", "decodeTokens",
                   "decodeVector", "restoreMerge"]:
        assert re.search(r"^  (/\*\*.*?\*/ )?(private )?%s\b" % re.escape(member), cls, flags=re.M), member
    for export in ["type Token", "type MergeToken", "type CompactMerge", "type MergeCode", "type BPETokenizerJSON", "const FS", "const EOF",
                   "const LF", "const CR", "function fileContentToCorpus", "function linesToCorpus", "function linesTrimmedToCorpus",
                   "function compactMerge", "class BPETokenizer"]:
        assert "export " + export in ts, export
    # the reference's fixed error texts (error convention)
    for text in ["'invalid format'", "'token table is empty, have you called tokenizer.addToCorpus()?'", "'unknown token, char: ' + JSON.stringify(",
                 "`unknown token index: ${", "`unknown vector index: ${", "`unknown token, a_code: ${JSON.stringify(a_code)}`",
                 "`unknown token, b_code: ${JSON.stringify(b_code)}`"]:
        assert text in ts, text
