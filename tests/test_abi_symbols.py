"""CPU: the C-ABI library builds, loads and exports every symbol include/bpe_b200.h declares
(no compute calls: there is no GPU here), and the compiled synthetic-corpus generator agrees with
its numpy specification."""
import ctypes as C
import os
import re

import numpy as np

from bpe_tokenizer_b200 import _abi
from bpe_tokenizer_b200.synth import synth_corpus, first_appearance_ids

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "bpe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bpe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load_library()
    declared = _header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_abi.SYMBOLS) == declared
    assert lib.bpe_abi_version() == 2


def test_struct_layouts_match_header():
    assert C.sizeof(_abi.bpe_merge) == 24
    assert _abi.MERGE_DTYPE.itemsize == 24
    assert C.sizeof(_abi.bpe_stats) == 28 * 8


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        return
    lib = _abi.load_library()
    h = C.c_void_p()
    assert lib.bpe_create(0, C.byref(h)) == _abi.BPE_E_CUDA
    from bpe_tokenizer_b200 import BPETokenizer, BpeError
    import pytest

    with pytest.raises(BpeError):
        BPETokenizer()


def test_compiled_synth_matches_numpy_spec():
    lib = _abi.load_library()
    nb, nd = C.c_int64(), C.c_int64()
    for target, seed in ((1, 43), (5000, 43), (200000, 44)):
        assert lib.bpe_synth_corpus(target, seed, 50000, 42, None, 0, None, 0, C.byref(nb), C.byref(nd)) == 0
        text = np.empty(nb.value, dtype=np.uint8)
        off = np.empty(nd.value + 1, dtype=np.int64)
        assert lib.bpe_synth_corpus(target, seed, 50000, 42, text.ctypes.data_as(_abi.u8p), text.size, _abi.p64(off), off.size, C.byref(nb), C.byref(nd)) == 0
        t2, o2 = synth_corpus(target, seed=seed)
        assert np.array_equal(text, t2) and np.array_equal(off, o2)
        assert text[0] == ord("\r") and text[-1] == ord("\n")
    ids, alphabet = first_appearance_ids(text)
    assert alphabet[0] == ord("\r") and len(alphabet) <= 29 and ids.max() == len(alphabet) - 1
