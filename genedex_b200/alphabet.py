"""Alphabets: mirror of genedex's `alphabet` module (src/alphabet.rs).

Symbols have an IO representation (usually ASCII) and a dense representation 0..k-1 used by the
index; dense 0 is always the sentinel / text delimiter (alphabet.rs:14-16).
"""
from __future__ import annotations

from typing import Iterable, Sequence


class Alphabet:
    """src/alphabet.rs:24-28."""

    def __init__(self, io_to_dense: bytes, dense_to_io: bytes, num_io_symbols_not_searchable: int):
        size = len(dense_to_io) + 1
        assert size > 1, "Alphabet size must be at least 2 (including sentinel)"
        assert size <= 256, "Alphabet size can be at most 256 (including sentinel)"
        assert len(io_to_dense) == 256
        dense = {s for s in io_to_dense if s != 0}
        assert len(dense) + 1 == size, "The alphabet translation tables are invalid."
        assert num_io_symbols_not_searchable + 2 <= size, \
            "Invalid alphabet. there must be at least one searchable symbol."
        self._io_to_dense = bytes(io_to_dense)
        self._dense_to_io = bytes(dense_to_io)
        self._not_searchable = num_io_symbols_not_searchable

    # alphabet.rs:43-75
    @classmethod
    def from_io_symbols(cls, symbols: Iterable[int], num_io_symbols_not_searchable: int = 0) -> "Alphabet":
        dense_to_io = bytes(symbols)
        assert len(set(dense_to_io)) == len(dense_to_io), "Symbols of the alphabet must be unique."
        assert len(dense_to_io) <= 255, "Alphabet size can be at most 255 (to leave space for the sentinel)."
        table = bytearray(256)
        for i, s in enumerate(dense_to_io):
            table[s] = i + 1
        return cls(bytes(table), dense_to_io, num_io_symbols_not_searchable)

    # alphabet.rs:99-149
    @classmethod
    def from_ambiguous_io_symbols(cls, groups: Sequence[bytes], num_io_symbols_not_searchable: int = 0) -> "Alphabet":
        groups = [bytes(g) for g in groups]
        assert all(len(g) > 0 for g in groups), "Every group of symbols must contain at least one symbol"
        flat = [s for g in groups for s in g]
        assert len(set(flat)) == len(flat), "Symbols of the alphabet must be unique."
        assert len(groups) <= 255, "Alphabet size can be at most 255 (to leave space for the sentinel)."
        table = bytearray(256)
        for i, g in enumerate(groups):
            for s in g:
                table[s] = i + 1
        return cls(bytes(table), bytes(g[0] for g in groups), num_io_symbols_not_searchable)

    # alphabet.rs:195-249
    def io_to_dense_representation(self, symbol: int) -> int:
        d = self._io_to_dense[symbol]
        if d == 0:
            raise ValueError("symbol in io representation should be valid")
        return d

    def try_io_to_dense_representation(self, symbol: int):
        d = self._io_to_dense[symbol]
        return None if d == 0 else d

    def dense_to_io_representation(self, symbol: int) -> int:
        r = self.try_dense_to_io_representation(symbol)
        if r is None:
            raise ValueError("symbol in dense representation should be valid")
        return r

    def try_dense_to_io_representation(self, symbol: int):
        if symbol == 0 or symbol - 1 >= len(self._dense_to_io):
            return None
        return self._dense_to_io[symbol - 1]

    def iter_io_symbols(self):
        return (i for i, d in enumerate(self._io_to_dense) if d != 0)

    def num_dense_symbols(self) -> int:
        return len(self._dense_to_io) + 1

    def num_searchable_dense_symbols(self) -> int:
        return self.num_dense_symbols() - self._not_searchable - 1

    def contains_sentinel_in_dense_representation(self) -> bool:
        return True

    @property
    def io_to_dense_table(self) -> bytes:
        return self._io_to_dense

    def __eq__(self, other):
        return isinstance(other, Alphabet) and (self._io_to_dense, self._dense_to_io, self._not_searchable) == (
            other._io_to_dense, other._dense_to_io, other._not_searchable)


def _pairs(letters: bytes):
    return [bytes([c, c + 32]) for c in letters]


def ascii_dna() -> Alphabet:  # alphabet.rs:251-253
    return Alphabet.from_ambiguous_io_symbols(_pairs(b"ACGT"), 0)


def ascii_dna_with_n() -> Alphabet:  # alphabet.rs:255-258: N is not allowed to be searched
    return Alphabet.from_ambiguous_io_symbols(_pairs(b"ACGTN"), 1)


def ascii_dna_iupac() -> Alphabet:  # alphabet.rs:264-273
    return Alphabet.from_ambiguous_io_symbols(_pairs(b"ACGTNRYKMSWBDHV"), 0)


def ascii_dna_iupac_as_dna_with_n() -> Alphabet:  # alphabet.rs:277-288
    return Alphabet.from_ambiguous_io_symbols(_pairs(b"ACGT") + [b"NnRrYyKkMmSsWwBbDdHhVv"], 1)


def ascii_amino_acid() -> Alphabet:  # alphabet.rs:291-299
    return Alphabet.from_ambiguous_io_symbols(_pairs(b"ACDEFGHIKLMNOPQRSTUVWY"), 0)


def ascii_amino_acid_iupac() -> Alphabet:  # alphabet.rs:303-336
    return Alphabet.from_ambiguous_io_symbols(_pairs(b"ABCDEFGHIJKLMNOPQRSTUVWXYZ") + [b"*"], 0)


def u8_until(max_symbol: int) -> Alphabet:  # alphabet.rs:339-341
    return Alphabet.from_io_symbols(range(0, max_symbol + 1), 0)


def ascii_printable() -> Alphabet:  # alphabet.rs:344-346
    return Alphabet.from_io_symbols(range(32, 127), 0)
