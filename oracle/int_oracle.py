"""TEST INFRASTRUCTURE ONLY -- ctypes binding of ``int_oracle.cpp``.

:class:`IntOracleTokenizer` keeps the dictionary / Token / JSON logic of
:class:`oracle.ref_literal.LiteralTokenizer` and swaps the three hot loops
(core.ts:265-310, :356-359, :404-406) for their compiled int-level
restatement, so MB-scale parity inputs finish in seconds.  The two forms are
checked against each other by ``tests/test_oracle_golden.py``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np

from .ref_literal import LiteralTokenizer, Token, js_stringify, utf16_len

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbpe_oracle.so")


def build_oracle(force: bool = False) -> str:
    src = os.path.join(_HERE, "int_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libbpe_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(_LIB_PATH)
        i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.orc_create.restype = C.c_void_p
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_add_document.argtypes = [C.c_void_p, i32p, C.c_int64]
        L.orc_clear_corpus.argtypes = [C.c_void_p]
        L.orc_set_len16.argtypes = [C.c_void_p, i32p, C.c_int32]
        L.orc_num_documents.argtypes = [C.c_void_p]
        L.orc_num_documents.restype = C.c_int64
        L.orc_document_length.argtypes = [C.c_void_p, C.c_int64]
        L.orc_document_length.restype = C.c_int64
        L.orc_get_document.argtypes = [C.c_void_p, C.c_int64, i32p]
        L.orc_total_tokens.argtypes = [C.c_void_p]
        L.orc_total_tokens.restype = C.c_int64
        L.orc_find_next_merge.argtypes = [C.c_void_p, C.c_int64, C.c_int32, i32p, i32p, i64p]
        L.orc_find_next_merge.restype = C.c_int
        L.orc_apply_merge.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.orc_merge_until.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, i32p, i32p, i64p, C.c_int64]
        L.orc_merge_until.restype = C.c_int64
        L.orc_load_merges.argtypes = [C.c_void_p, i32p, C.c_int64]
        L.orc_encode.argtypes = [C.c_void_p, i32p, C.c_int64, i32p]
        L.orc_encode.restype = C.c_int64
        L.orc_encode_fast.argtypes = [C.c_void_p, i32p, C.c_int64, i32p]
        L.orc_encode_fast.restype = C.c_int64
        _lib = L
    return _lib


def _p32(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _p64(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


class IntOracle:
    """Thin handle over the C functions (ids in, ids out)."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())

    def __del__(self):
        try:
            self.L.orc_destroy(self.h)
        except Exception:
            pass

    def add_document(self, ids) -> None:
        a = np.ascontiguousarray(ids, dtype=np.int32)
        self.L.orc_add_document(self.h, _p32(a), a.size)

    def add_documents(self, ids, offsets) -> None:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        for d in range(len(offsets) - 1):
            self.add_document(ids[offsets[d]:offsets[d + 1]])

    def clear_corpus(self) -> None:
        self.L.orc_clear_corpus(self.h)

    def set_len16(self, len16) -> None:
        a = np.ascontiguousarray(len16, dtype=np.int32)
        self.L.orc_set_len16(self.h, _p32(a), a.size)

    def num_documents(self) -> int:
        return self.L.orc_num_documents(self.h)

    def total_tokens(self) -> int:
        return self.L.orc_total_tokens(self.h)

    def document(self, d: int) -> np.ndarray:
        n = self.L.orc_document_length(self.h, d)
        out = np.empty(n, dtype=np.int32)
        if n:
            self.L.orc_get_document(self.h, d, _p32(out))
        return out

    def find_next_merge(self, min_weight: int, max_length: int):
        a, b, w = C.c_int32(), C.c_int32(), C.c_int64()
        ok = self.L.orc_find_next_merge(self.h, min_weight, max_length, C.byref(a), C.byref(b), C.byref(w))
        return (a.value, b.value, w.value) if ok else None

    def apply_merge(self, a: int, b: int, c: int) -> None:
        self.L.orc_apply_merge(self.h, a, b, c)

    def merge_until(self, min_weight: int, max_length: int, max_iterations: int, first_new_index: int, cap: int):
        la = np.empty(cap, dtype=np.int32)
        lb = np.empty(cap, dtype=np.int32)
        lw = np.empty(cap, dtype=np.int64)
        n = self.L.orc_merge_until(self.h, min_weight, max_length, max_iterations, first_new_index, _p32(la), _p32(lb), _p64(lw), cap)
        return la[:n], lb[:n], lw[:n]

    def load_merges(self, abc) -> None:
        a = np.ascontiguousarray(abc, dtype=np.int32).reshape(-1)
        self.L.orc_load_merges(self.h, _p32(a), a.size // 3)

    def encode(self, ids, fast: bool = False) -> np.ndarray:
        a = np.ascontiguousarray(ids, dtype=np.int32)
        out = np.empty(max(a.size, 1), dtype=np.int32)
        fn = self.L.orc_encode_fast if fast else self.L.orc_encode
        n = fn(self.h, _p32(a), a.size, _p32(out))
        return out[:n].copy()


class IntOracleTokenizer(LiteralTokenizer):
    """LiteralTokenizer with the hot loops running in ``int_oracle.cpp``."""

    def __init__(self):
        super().__init__()
        self._o = IntOracle()
        self._len16_synced = 0

    # -- corpus lives in C; expose the reference's public field on demand ----
    @property
    def corpus_in_code(self) -> List[str]:
        o = self.__dict__.get("_o")
        if o is None:
            return []
        return ["".join(chr(int(i) + 1) for i in o.document(d)) for d in range(o.num_documents())]

    @corpus_in_code.setter
    def corpus_in_code(self, value) -> None:
        o = self.__dict__.get("_o")
        if o is None:
            return
        o.clear_corpus()
        for s in value:
            o.add_document([ord(ch) - 1 for ch in s])

    def _sync_len16(self) -> None:
        self._o.set_len16([utf16_len(t.chars) for t in self.token_table])

    def from_json(self, json: dict) -> None:
        super().from_json(json)
        self._o = IntOracle()
        self._o.load_merges([[a.index, b.index, c.index] for a, b, c in self.merge_tokens])

    # core.ts:182-207
    def add_to_corpus(self, content: str) -> None:
        ids = []
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
            ids.append(token.index)
        self._o.add_document(ids)

    def add_ids(self, ids, offsets, weights_already_counted: bool = False) -> None:
        """Bulk form of add_to_corpus for pre-mapped single-character ids."""
        ids = np.asarray(ids, dtype=np.int32)
        if not weights_already_counted:
            cnt = np.bincount(ids, minlength=len(self.token_table))
            for i, c in enumerate(cnt):
                self.token_table[i].weight += int(c)
                self.token_table[i].original_weight += int(c)
        self._o.add_documents(ids, offsets)

    # core.ts:213-216
    def restore_to_corpus(self, content: str) -> None:
        self._o.add_document(self._encode_ids(content))

    # core.ts:247-326
    def find_next_merge(self, min_weight=None, max_length=None):
        self._sync_len16()
        r = self._o.find_next_merge(int(min_weight or 2), int(max_length or 0))
        if r is None:
            return None
        a, b, w = r
        ta, tb = self.token_table[a], self.token_table[b]
        new_index = len(self.token_table)
        return (ta, tb, Token(ta.chars + tb.chars, w, w, chr(new_index + 1), new_index))

    # core.ts:332-360
    def apply_merge(self, merge) -> None:
        a, b, c = merge
        a.weight -= c.weight
        b.weight -= c.weight
        self._invalidate_vector_index()
        self.code_to_token[c.code] = c
        self.token_table.append(c)
        self.merge_tokens.append(merge)
        self.merge_codes.append((a.code + b.code, c.code))
        self._sync_len16()
        self._o.apply_merge(a.index, b.index, c.index)

    # core.ts:365-383 (loop kept in C so that 10 MB x 4k merges is practical)
    def merge_until(self, min_weight=None, max_length=None, max_iterations=None) -> int:
        self._sync_len16()
        first = len(self.token_table)
        cap = int(max_iterations) if max_iterations else 1 << 20
        la, lb, lw = self._o.merge_until(int(min_weight or 2), int(max_length or 0), int(max_iterations or 0), first, cap)
        for a, b, w in zip(la.tolist(), lb.tolist(), lw.tolist()):
            ta, tb = self.token_table[a], self.token_table[b]
            index = len(self.token_table)
            c = Token(ta.chars + tb.chars, w, w, chr(index + 1), index)
            ta.weight -= w
            tb.weight -= w
            self.code_to_token[c.code] = c
            self.token_table.append(c)
            self.merge_tokens.append((ta, tb, c))
            self.merge_codes.append((ta.code + tb.code, c.code))
        self._invalidate_vector_index()
        return len(la)

    def _encode_ids(self, content: str, fast: bool = False) -> np.ndarray:
        ids = []
        for char in content:
            token = self.char_to_token.get(char)
            if token is None:
                raise ValueError("unknown token, char: " + js_stringify(char))
            ids.append(token.index)
        return self._o.encode(ids, fast=fast)

    # core.ts:392-409
    def encode_to_code(self, content: str) -> str:
        return "".join(chr(int(i) + 1) for i in self._encode_ids(content))

    def encode_ids(self, ids, fast: bool = False) -> np.ndarray:
        """Encode pre-mapped single-character ids (bulk parity helper)."""
        return self._o.encode(ids, fast=fast)
