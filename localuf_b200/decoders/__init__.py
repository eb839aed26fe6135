"""Decoders of the B200 engine (reference: ``localuf/decoders/__init__.py``)."""
from localuf_b200.decoders.uf import UF, BUF, NodeUF, NodeBUF  # noqa: F401
from localuf_b200.decoders.luf import Macar, Actis  # noqa: F401
from localuf_b200.decoders.snowflake import Snowflake  # noqa: F401

__all__ = ['UF', 'BUF', 'NodeUF', 'NodeBUF', 'Macar', 'Actis', 'Snowflake']
