"""Seeded, integer-only synthetic corpus: Zipf-distributed
"words" over a-z, documents wrapped the way ``linesToCorpus`` wraps lines
(reference core.ts:61-64: ``'\\r' + line + '\\n'``).

Definition (bit-reproducible; the C++ twin in csrc/synth.cpp must agree byte for byte):

* PRNG = splitmix64 used counter-style: ``stream(seed)[i] = mix(seed + (i+1)*GAMMA)``.
* Word list: ``V`` words from ``stream(word_seed)`` consumed sequentially: one draw
  ``r`` -> length ``2 + r % 9``, then one draw per letter -> ``'a' + r % 26``.
* Zipf (s = 1): integer weights ``w_k = 2**40 // (k+1)``; a draw ``r`` picks the first
  ``k`` with ``cum[k] > r % total``.
* Document ``d`` of a text with seed ``s``: its own sub-stream seeded by
  ``stream(s)[d]``; draw 0 -> ``8 + r % 57`` words, draw ``i+1`` -> word ``i``.
  Text = ``'\\r' + ' '.join(words) + '\\n'``.  Documents are emitted until the total
  byte count reaches ``target_bytes`` (last document kept whole).

This numpy form is the specification and is used for small inputs; large inputs
go through the compiled generator (``bpe_synth_*`` in libbpe_b200.so).
"""
from __future__ import annotations

import numpy as np

GAMMA = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
V_DEFAULT = 50000
WORD_SEED = 42
TRAIN_SEED = 43
ENCODE_SEED = 44


def _mix(z: np.ndarray) -> np.ndarray:
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def stream(seed, n: int, start: int = 0) -> np.ndarray:
    """draws ``start .. start+n-1`` of the splitmix64 stream seeded with ``seed``
    (``seed`` may be an array: one stream per element, result shape seed.shape+(n,))."""
    seed = np.asarray(seed, dtype=np.uint64)
    idx = np.arange(start + 1, start + n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        return _mix(seed[..., None] + idx * GAMMA)


class WordList:
    def __init__(self, vocab: int = V_DEFAULT, seed: int = WORD_SEED):
        self.vocab = vocab
        # sequential consumption: at most 1 + 10 draws per word
        draws = stream(np.uint64(seed), vocab * 11)
        words = []
        p = 0
        for _ in range(vocab):
            ln = 2 + int(draws[p] % np.uint64(9))
            p += 1
            letters = (draws[p:p + ln] % np.uint64(26)).astype(np.uint8) + ord("a")
            p += ln
            words.append(letters.tobytes())
        self.words = words
        self.lengths = np.array([len(w) for w in words], dtype=np.int64)
        self.offsets = np.zeros(vocab + 1, dtype=np.int64)
        np.cumsum(self.lengths, out=self.offsets[1:])
        self.blob = np.frombuffer(b"".join(words), dtype=np.uint8)
        w = (np.uint64(1) << np.uint64(40)) // np.arange(1, vocab + 1, dtype=np.uint64)
        self.cum = np.cumsum(w, dtype=np.uint64)
        self.total = self.cum[-1]

    def pick(self, r: np.ndarray) -> np.ndarray:
        return np.searchsorted(self.cum, r % self.total, side="right")


_wordlists = {}


def wordlist(vocab: int = V_DEFAULT, seed: int = WORD_SEED) -> WordList:
    key = (vocab, seed)
    if key not in _wordlists:
        _wordlists[key] = WordList(vocab, seed)
    return _wordlists[key]


def synth_corpus(target_bytes: int, seed: int = TRAIN_SEED, vocab: int = V_DEFAULT, word_seed: int = WORD_SEED):
    """-> (text_bytes uint8[n], doc_offsets int64[D+1]); pure numpy, fine up to ~100 MB."""
    wl = wordlist(vocab, word_seed)
    chunks, lens = [], []
    total, d0 = 0, 0
    batch = max(16, min(1 << 16, target_bytes // 200 + 16))
    while total < target_bytes:
        dseeds = stream(np.uint64(seed), batch, start=d0)  # (batch,)
        draws = stream(dseeds, 65)  # (batch, 65)
        nwords = 8 + (draws[:, 0] % np.uint64(57)).astype(np.int64)
        wid = wl.pick(draws[:, 1:])  # (batch, 64)
        mask = np.arange(64)[None, :] < nwords[:, None]
        wlen = np.where(mask, wl.lengths[wid], 0)
        doclen = wlen.sum(axis=1) + nwords + 1  # '\r' + words + (nwords-1) spaces + '\n'
        cum = np.cumsum(doclen)
        need = target_bytes - total
        k = int(np.searchsorted(cum, need, side="left")) + 1  # docs until total >= target
        k = min(k, batch)
        for i in range(k):
            parts = [b"\r", b" ".join(wl.words[j] for j in wid[i, : nwords[i]]), b"\n"]
            chunks.append(b"".join(parts))
        lens.extend(doclen[:k].tolist())
        total += int(cum[k - 1])
        d0 += k
    text = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    offsets = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(np.array(lens, dtype=np.int64), out=offsets[1:])
    assert offsets[-1] == text.size
    return text, offsets


_native = None


def native_corpus(target_bytes: int, seed: int = TRAIN_SEED, vocab: int = V_DEFAULT, word_seed: int = WORD_SEED):
    """Same text as synth_corpus, from the compiled generator (csrc/synth.cpp built on its own as libbpe_synth.so --
    host code only, no CUDA): -> (text_bytes uint8[n], doc_offsets int64[D+1]).  For the 10 MB .. 1 GB corpora."""
    import ctypes as C

    global _native
    if _native is None:
        from . import build as _build

        lib = C.CDLL(_build.build_synth())
        lib.bpe_synth_corpus.restype = C.c_int
        lib.bpe_synth_corpus.argtypes = [C.c_int64, C.c_uint64, C.c_int32, C.c_uint64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                         C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _native = lib
    nb, nd = C.c_int64(), C.c_int64()
    if _native.bpe_synth_corpus(target_bytes, seed, vocab, word_seed, None, 0, None, 0, C.byref(nb), C.byref(nd)) != 0:
        raise RuntimeError("bpe_synth_corpus (sizing) failed")
    text = np.empty(nb.value, dtype=np.uint8)
    off = np.empty(nd.value + 1, dtype=np.int64)
    if _native.bpe_synth_corpus(target_bytes, seed, vocab, word_seed, text.ctypes.data, text.size, off.ctypes.data, off.size,
                                C.byref(nb), C.byref(nd)) != 0:
        raise RuntimeError("bpe_synth_corpus failed")
    return text, off


def first_appearance_ids(text: np.ndarray):
    """Map bytes to token indices in first-appearance order, as addToCorpus does
    (reference core.ts:186-199).  -> (ids int32[n], alphabet list[int] by index)."""
    text = np.asarray(text, dtype=np.uint8)
    first = np.full(256, text.size, dtype=np.int64)
    # position of first occurrence per byte value
    vals, idx = np.unique(text, return_index=True)
    first[vals] = idx
    order = vals[np.argsort(idx, kind="stable")]
    lut = np.full(256, -1, dtype=np.int32)
    lut[order] = np.arange(order.size, dtype=np.int32)
    return lut[text], [int(x) for x in order]
