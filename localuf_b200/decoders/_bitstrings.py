"""Bit-string wire format of Snowflake (``localuf/decoders/snowflake/main.py:729-883``).

Input: one *syndrome vector* per decoding cycle -- the detectors of the newly
exposed layer as a string of '0' / '1' (repetition code: west to east; surface
code: west to east along a row, rows from south to north).  Output: one *floor
vector* per cycle -- the correction bits of the bottom layer's edges just before
the window drops, ordered by edge type, then north to south, then west to east.
"""
from __future__ import annotations


class Bitstrings:
    """Node and edge orders of the wire format for one code distance / window height."""

    def __init__(self, d: int, window_height: int, repetition: bool):
        self.D = d
        self.TOP = window_height - 1
        self.REPETITION = repetition

    @property
    def LENGTH(self) -> int:
        """Bits of a syndrome vector = detectors of one layer."""
        return self.D - 1 if self.REPETITION else self.D * (self.D - 1)

    def _kind(self) -> str:
        return 'Repetition' if self.REPETITION else 'Surface'

    # -- detectors ---------------------------------------------------------
    def detector(self, position: int):
        """Coordinates (in the top layer) of the detector at bit ``position``."""
        if self.REPETITION:
            return (position, self.TOP)
        cols = self.D - 1
        row_from_south, j = divmod(position, cols)
        return (self.D - 1 - row_from_south, j, self.TOP)

    def position(self, defect) -> int:
        if self.REPETITION:
            j, t = defect
            ok = 0 <= j < self.D - 1
            pos = j
        else:
            i, j, t = defect
            ok = 0 <= i < self.D and 0 <= j < self.D - 1
            pos = (self.D - 1 - i) * (self.D - 1) + j
        if not ok or t != self.TOP:
            raise ValueError(f'Defect {defect} out of range for {self._kind().lower()} code of distance {self.D} '
                             f'with window height {self.TOP + 1}.')
        return pos

    def vector_to_set(self, vector: str) -> set:
        if len(vector) != self.LENGTH:
            raise ValueError(f'{self._kind()} code of distance {self.D} expects syndrome vector of length '
                             f'{self.LENGTH} but received length {len(vector)}.')
        return {self.detector(k) for k, bit in enumerate(vector) if bit == '1'}

    def set_to_vector(self, syndrome) -> str:
        bits = ['0'] * self.LENGTH
        for defect in syndrome:
            bits[self.position(defect)] = '1'
        return ''.join(bits)

    def syndrome_vectors_to_sets(self, vectors) -> list:
        return [self.vector_to_set(v) for v in vectors]

    def sets_to_syndrome_vectors(self, syndromes) -> list:
        return [self.set_to_vector(s) for s in syndromes]

    # -- floor edges ---------------------------------------------------------
    def lowest_edges(self, noise_model: str, _include_timelike_lowest_edges: bool = True) -> tuple:
        """Edges whose lower end is in layer 0, in output order."""
        d = self.D
        if self.REPETITION:
            if noise_model != 'phenomenological':
                raise NotImplementedError(f"Noise model {noise_model} not implemented for repetition code.")
            groups = [[((j, 0), (j, 1)) for j in range(d - 1)] if _include_timelike_lowest_edges else [],
                      [((j, 0), (j + 1, 0)) for j in range(-1, d - 1)]]
            return tuple(e for g in groups for e in g)
        rows, inner = range(d), range(d - 1)
        timelike = [((i, j, 0), (i, j, 1)) for i in rows for j in inner] if _include_timelike_lowest_edges else []
        south = [((i, j, 0), (i + 1, j, 0)) for i in range(d - 1) for j in inner]
        east = [((i, j, 0), (i, j + 1, 0)) for i in rows for j in range(-1, d - 1)]
        if noise_model == 'phenomenological':
            groups = [timelike, south, east]
        elif noise_model == 'circuit-level':
            south_down = [((i, j, 1), (i + 1, j, 0)) for i in range(d - 1) for j in inner]
            east_up = [((i, j, 0), (i, j + 1, 1)) for i in rows for j in range(d - 2)]
            south_east_up = [((i, j, 0), (i + 1, j + 1, 1)) for i in range(d - 1) for j in range(d - 2)]
            groups = [timelike, south_down, east_up, south_east_up, south, east]
        else:
            raise NotImplementedError(f"Noise model {noise_model} not implemented for surface code.")
        return tuple(e for g in groups for e in g)
