"""Enumerations shared with the reference (``localuf/constants.py:17-31``)."""
from enum import IntEnum


class Growth(IntEnum):
    """Growth value of an edge: UNGROWN -> HALF -> FULL; BURNT only in dynamic UF."""
    BURNT = -1
    UNGROWN = 0
    HALF = 1
    FULL = 2


# device encoding (2-bit fields) -> Growth
GROWTH_FROM_DEVICE = (Growth.UNGROWN, Growth.HALF, Growth.FULL, Growth.BURNT)
