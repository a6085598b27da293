"""bpe_tokenizer_b200 -- B200-native BPE train/encode engine behind the API of
beenotung/bpe-tokenizer's in-memory ``BPETokenizer`` (reference core.ts).

The hot path lives in ``csrc/`` (hand-written sm_100a CUDA kernels + the C ABI of
``include/bpe_b200.h``); ``tokenizer.py`` is the host-side mirror of the reference class.
"""
from .tokenizer import (  # noqa: F401
    BPETokenizer,
    Token,
    compactMerge,
    fileContentToCorpus,
    linesToCorpus,
    linesTrimmedToCorpus,
    FS,
    EOF,
    LF,
    CR,
)
from ._abi import BpeError, load_library  # noqa: F401

__all__ = [
    "BPETokenizer", "Token", "compactMerge", "fileContentToCorpus", "linesToCorpus", "linesTrimmedToCorpus",
    "FS", "EOF", "LF", "CR", "BpeError", "load_library",
]
