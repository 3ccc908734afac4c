"""pybpl_b200 -- B200-native render+score path for pyBPL.

Importing the package is cheap; anything that computes loads libpybpl_b200.so and raises if it is
missing (build it with `python -m pybpl_b200.build`).  There is no CPU fallback."""
from . import _lib  # noqa: F401
from .parameters import Parameters  # noqa: F401

__all__ = ["Parameters", "ops", "rendering", "splines", "model", "objects", "dropin", "workload"]
