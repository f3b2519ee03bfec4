"""Score matrices for tests and the bench harness (caller-side helper, not the hot path).

The library itself only ever sees ``int* scoreMatrix`` + ``alphabetLength``
(reference src/opal.h:115-116).  This module provides what the reference's CLI
gets from ``ScoreMatrix`` (reference src/ScoreMatrix.cpp:17-35, 57-84): the
standard NCBI BLOSUM62 / BLOSUM50 tables in the reference's letter order, a
reader for its ``.mat`` text format (alphabet line, then rows) and simple
match/mismatch matrices (reference src/test.cpp:190-197).
"""
from __future__ import annotations

import numpy as np

# Standard BLOSUM tables, stored as lower triangles (they are symmetric).
_BLOSUM62_ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX"
_BLOSUM62_TRI = [
    "4", "-1 5", "-2 0 6", "-2 -2 1 6", "0 -3 -3 -3 9", "-1 1 0 0 -3 5", "-1 0 0 2 -4 2 5",
    "0 -2 0 -1 -3 -2 -2 6", "-2 0 1 -1 -3 0 0 -2 8", "-1 -3 -3 -3 -1 -3 -3 -4 -3 4",
    "-1 -2 -3 -4 -1 -2 -3 -4 -3 2 4", "-1 2 0 -1 -3 1 1 -2 -1 -3 -2 5",
    "-1 -1 -2 -3 -1 0 -2 -3 -2 1 2 -1 5", "-2 -3 -3 -3 -2 -3 -3 -3 -1 0 0 -3 0 6",
    "-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4 7", "1 -1 1 0 -1 0 0 0 -1 -2 -2 0 -1 -2 -1 4",
    "0 -1 0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1 1 5",
    "-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1 1 -4 -3 -2 11",
    "-2 -2 -2 -3 -2 -1 -2 -3 2 -1 -1 -2 -1 3 -3 -2 -2 2 7",
    "0 -3 -3 -3 -1 -2 -2 -3 -3 3 1 -2 1 -1 -2 -2 0 -3 -1 4",
    "-2 -1 3 4 -3 0 1 -1 0 -3 -4 0 -3 -3 -2 0 -1 -4 -3 -3 4",
    "-1 0 0 1 -3 3 4 -2 0 -3 -3 1 -1 -3 -1 0 -1 -3 -2 -2 1 4",
    "0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2 0 0 -2 -1 -1 -1 -1 -1",
]
_BLOSUM50_ALPHABET = "ARNDCQEGHILKMFPSTWYVBZX*"
_BLOSUM50_TRI = [
    "5", "-2 7", "-1 -1 7", "-2 -2 2 8", "-1 -4 -2 -4 13", "-1 1 0 0 -3 7", "-1 0 0 2 -3 2 6",
    "0 -3 0 -1 -3 -2 -3 8", "-2 0 1 -1 -3 1 0 -2 10", "-1 -4 -3 -4 -2 -3 -4 -4 -4 5",
    "-2 -3 -4 -4 -2 -2 -3 -4 -3 2 5", "-1 3 0 -1 -3 2 1 -2 0 -3 -3 6",
    "-1 -2 -2 -4 -2 0 -2 -3 -1 2 3 -2 7", "-3 -3 -4 -5 -2 -4 -3 -4 -1 0 1 -4 0 8",
    "-1 -3 -2 -1 -4 -1 -1 -2 -2 -3 -4 -1 -3 -4 10", "1 -1 1 0 -1 0 -1 0 -1 -3 -3 0 -2 -3 -1 5",
    "0 -1 0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1 2 5",
    "-3 -3 -4 -5 -5 -1 -3 -3 -3 -3 -2 -3 -1 1 -4 -4 -3 15",
    "-2 -1 -2 -3 -3 -1 -2 -3 2 -1 -1 -2 0 4 -3 -2 -2 2 8",
    "0 -3 -3 -4 -1 -3 -3 -4 -4 4 1 -3 1 -1 -3 -2 0 -3 -1 5",
    "-2 -1 5 6 -3 0 1 -1 0 -4 -4 0 -3 -4 -2 0 0 -5 -3 -3 6",
    "-1 0 0 1 -3 4 5 -2 0 -3 -3 1 -1 -4 -1 0 -1 -2 -2 -3 1 5",
    "-1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1 -1",
    "-5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 -5 1",
]


class ScoreMatrix:
    """Alphabet + row-major int32 matrix, as the C API wants them."""

    def __init__(self, alphabet: str, matrix):
        self.alphabet = alphabet
        self.matrix = np.ascontiguousarray(matrix, dtype=np.int32).reshape(len(alphabet), len(alphabet))

    @property
    def alphabet_length(self):
        return len(self.alphabet)

    def flat(self):
        return self.matrix.ravel()

    def letter_index(self):
        """letter -> alphabet index; '*' (if present) catches every unknown letter, as in
        reference src/opal_aligner.cpp:250-258.  Without '*' unknown letters are an error here
        (the reference reads uninitialised memory)."""
        idx = np.full(256, -1, dtype=np.int16)
        if "*" in self.alphabet:
            idx[:] = self.alphabet.index("*")
        for i, ch in enumerate(self.alphabet):
            idx[ord(ch)] = i
        return idx

    def encode(self, text: str):
        codes = self.letter_index()[np.frombuffer(text.encode("ascii"), dtype=np.uint8)]
        if (codes < 0).any():
            raise ValueError("sequence holds letters outside the matrix alphabet")
        return codes.astype(np.uint8)

    @classmethod
    def from_file(cls, path):
        """Reader for the reference's .mat format (reference src/ScoreMatrix.cpp:17-35)."""
        with open(path) as f:
            lines = [l for l in f.read().split("\n") if l.strip()]
        alphabet = "".join(tok[0] for tok in lines[0].split())
        vals = [int(x) for l in lines[1:] for x in l.split()]
        return cls(alphabet, vals)


def _from_triangle(alphabet, tri):
    n = len(alphabet)
    m = np.zeros((n, n), dtype=np.int32)
    for i, row in enumerate(tri):
        vals = [int(x) for x in row.split()]
        assert len(vals) == i + 1
        m[i, : i + 1] = vals
        m[: i + 1, i] = vals
    return ScoreMatrix(alphabet, m)


def blosum62():
    return _from_triangle(_BLOSUM62_ALPHABET, _BLOSUM62_TRI)


def blosum50():
    return _from_triangle(_BLOSUM50_ALPHABET, _BLOSUM50_TRI)


def simple(alphabet_length, match, mismatch, alphabet=None):
    """match on the diagonal, mismatch elsewhere (reference src/test.cpp:190-197)."""
    m = np.full((alphabet_length, alphabet_length), mismatch, dtype=np.int32)
    np.fill_diagonal(m, match)
    alphabet = alphabet or "ACGT"[:alphabet_length] if alphabet_length <= 4 else alphabet
    if alphabet is None:
        alphabet = "".join(chr(ord("A") + i) for i in range(alphabet_length))
    return ScoreMatrix(alphabet, m)
