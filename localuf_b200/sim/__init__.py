"""Experiment drivers (reference: ``localuf/sim``)."""
from localuf_b200.sim import accuracy, runtime  # noqa: F401
