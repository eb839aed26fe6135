"""Experiment drivers (reference: ``localuf/sim``)."""
from localuf_b200.sim import accuracy, latency, runtime  # noqa: F401
