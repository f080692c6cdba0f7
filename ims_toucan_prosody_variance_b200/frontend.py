"""Phoneme string -> articulatory feature tensor as a vectorised lookup.

The reference does this one character at a time in Python (`ArticulatoryCombinedTextFrontend.string_to_tensor`,
Preprocessing/TextFrontend.py:213-288): a dict lookup and a list copy per phoneme, a chain of 14 string comparisons
per character.  At the engine's throughput (thousands of sentences per second per GPU) that loop is the host
bottleneck, so the same mapping is done here with array operations over the code points of a sentence, and batches
are staged in pinned memory for an asynchronous host-to-device copy.

Grapheme-to-phoneme conversion (espeak, `get_phone_string`) stays with the reference's frontend: this module starts
from the phoneme string.  The phone -> vector table and the feature index are the reference's own
(`generate_feature_table()`, `get_feature_to_index_lookup()`, Preprocessing/articulatory_features.py:817,904) and are
passed in, not duplicated.
"""
import numpy as np
import torch

# characters that modify the PREVIOUS phoneme vector (TextFrontend.py:236-274): character -> feature name
_PREVIOUS = {
    "ː": "lengthened", "ˑ": "half-length", "̆": "shortened", "̃": "nasal",
    "˥": "very-high-tone", "˦": "high-tone", "˧": "mid-tone", "˨": "low-tone", "˩": "very-low-tone",
    "⭧": "rising-tone", "⭨": "falling-tone", "⮁": "peaking-tone", "⮃": "dipping-tone",
}
_STRESS = "ˈ"          # primary stress: marks the NEXT phoneme (TextFrontend.py:232-234,282-284)
_UNKNOWN, _BASE, _STRESSED, _PREV = 0, 1, 2, 3


class PhoneTensoriser:
    """`encode(phones)` == `string_to_tensor(phones, input_phonemes=True)` of the reference, `batch(list)` pads a list of
    sentences into one pinned (B, Tmax, D) tensor plus lengths."""

    def __init__(self, phone_to_vector, feature_to_index):
        self.dim = len(next(iter(phone_to_vector.values())))
        keys = list(phone_to_vector.keys())
        if any(len(k) != 1 for k in keys):
            raise ValueError("phone_to_vector keys must be single characters")
        codes = [ord(k) for k in keys] + [ord(c) for c in _PREVIOUS] + [ord(_STRESS)]
        self._size = max(codes) + 2                                   # last slot: every code point beyond the table
        self._kind = np.zeros(self._size, dtype=np.uint8)
        self._row = np.zeros(self._size, dtype=np.int32)
        self._feat = np.zeros(self._size, dtype=np.int32)
        self._table = np.asarray([phone_to_vector[k] for k in keys], dtype=np.float32).reshape(len(keys), self.dim)
        for c, name in _PREVIOUS.items():                             # the if / elif chain of the reference comes first:
            self._kind[ord(c)] = _PREV                                # a modifier character is never looked up in the table
            self._feat[ord(c)] = feature_to_index[name]
        self._kind[ord(_STRESS)] = _STRESSED
        for i, k in enumerate(keys):
            if self._kind[ord(k)] == _UNKNOWN:
                self._kind[ord(k)] = _BASE
                self._row[ord(k)] = i
        self._stressed = feature_to_index["stressed"]

    @classmethod
    def from_frontend(cls, frontend, feature_to_index=None):
        """Build from the reference's frontend object (its `phone_to_vector`); the feature index comes from the
        argument, the frontend's `feature_to_index` attribute, or the reference module when it is importable."""
        if feature_to_index is None:
            feature_to_index = getattr(frontend, "feature_to_index", None)
        if feature_to_index is None:
            from Preprocessing.articulatory_features import get_feature_to_index_lookup   # the reference, on sys.path
            feature_to_index = get_feature_to_index_lookup()
        return cls(frontend.phone_to_vector, feature_to_index)

    def encode(self, phones, handle_missing=True):
        """(T, D) float32 array.  Same results and the same failure modes as the reference loop: a modifier (or a stress
        mark resolved by an unknown character) before the first phoneme raises IndexError; an unknown character raises
        KeyError unless handle_missing, in which case it is skipped -- and, like in the reference, still consumes a
        pending stress mark, which then lands on the previous phoneme."""
        phones = phones.replace("ɚ", "ə").replace("ᵻ", "ɨ")          # TextFrontend.py:223
        codes = np.frombuffer(phones.encode("utf-32-le"), dtype=np.uint32).astype(np.int64)
        idx = np.minimum(codes, self._size - 1)
        kind = self._kind[idx]
        base = kind == _BASE
        if not handle_missing and np.any(kind == _UNKNOWN):
            raise KeyError(chr(int(codes[np.argmax(kind == _UNKNOWN)])))
        out = self._table[self._row[idx[base]]].copy()
        n_incl = np.cumsum(base)                                      # phonemes appended up to and including position i
        prev = np.nonzero(kind == _PREV)[0]
        if prev.size:
            rows = n_incl[prev] - 1
            if rows.min() < 0:
                raise IndexError("modifier before the first phoneme")
            out[rows, self._feat[idx[prev]]] = 1.0
        stress = np.nonzero(kind == _STRESSED)[0]
        if stress.size:
            other = np.nonzero((kind == _BASE) | (kind == _UNKNOWN))[0]   # characters that reach the final else branch
            nxt = np.searchsorted(other, stress, side="right")
            nxt = nxt[nxt < other.size]                                   # a trailing stress mark is never consumed
            if nxt.size:
                rows = n_incl[other[nxt]] - 1
                if rows.min() < 0:
                    raise IndexError("stress mark resolved before the first phoneme")
                out[rows, self._stressed] = 1.0
        return out

    def tensor(self, phones, handle_missing=True):
        return torch.from_numpy(self.encode(phones, handle_missing))

    def batch(self, phone_strings, pin=True, handle_missing=True):
        """-> (features (B, Tmax, D) float32 [pinned when CUDA is available], lengths (B) int32, list of (T_i, D) views).
        One pass over the code points of ALL sentences (no per-sentence Python work beyond the two replacements)."""
        strs = [p.replace("ɚ", "ə").replace("ᵻ", "ɨ") for p in phone_strings]
        n_sent = len(strs)
        nchar = np.fromiter((len(x) for x in strs), dtype=np.int64, count=n_sent)
        codes = np.frombuffer("".join(strs).encode("utf-32-le"), dtype=np.uint32).astype(np.int64)
        sent = np.repeat(np.arange(n_sent), nchar)
        idx = np.minimum(codes, self._size - 1)
        kind = self._kind[idx]
        base = kind == _BASE
        if not handle_missing and np.any(kind == _UNKNOWN):
            raise KeyError(chr(int(codes[np.argmax(kind == _UNKNOWN)])))
        n_incl = np.cumsum(base)                                      # global row count up to and including position i
        lens = np.bincount(sent[base], minlength=n_sent).astype(np.int64)
        first = np.concatenate([[0], np.cumsum(lens)[:-1]]) if n_sent else np.zeros(0, dtype=np.int64)
        flat = self._table[self._row[idx[base]]]                      # (total phonemes, D), a fresh array
        prev = np.nonzero(kind == _PREV)[0]
        if prev.size:
            rows = n_incl[prev] - 1
            if np.any(rows < first[sent[prev]]):
                raise IndexError("modifier before the first phoneme of a sentence")
            flat[rows, self._feat[idx[prev]]] = 1.0
        stress = np.nonzero(kind == _STRESSED)[0]
        if stress.size:
            other = np.nonzero((kind == _BASE) | (kind == _UNKNOWN))[0]
            nxt = np.searchsorted(other, stress, side="right")
            ok = nxt < other.size
            stress, nxt = stress[ok], nxt[ok]
            same = sent[other[nxt]] == sent[stress]                   # a stress mark never crosses into the next sentence
            stress, nxt = stress[same], nxt[same]
            if stress.size:
                rows = n_incl[other[nxt]] - 1
                if np.any(rows < first[sent[stress]]):
                    raise IndexError("stress mark resolved before the first phoneme of a sentence")
                flat[rows, self._stressed] = 1.0
        t_max = int(lens.max()) if n_sent else 0
        use_pin = bool(pin and torch.cuda.is_available())
        feats = torch.zeros((n_sent, t_max, self.dim), dtype=torch.float32, pin_memory=use_pin)
        if flat.shape[0]:
            row_sent = np.repeat(np.arange(n_sent), lens)
            dest = row_sent * t_max + (np.arange(flat.shape[0]) - first[row_sent])
            feats.view(-1, self.dim)[torch.from_numpy(dest)] = torch.from_numpy(flat)
        lens_l = [int(v) for v in lens]
        return feats, torch.tensor(lens_l, dtype=torch.int32), [feats[i, :n] for i, n in enumerate(lens_l)]
