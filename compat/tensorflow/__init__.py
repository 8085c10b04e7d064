"""Minimal stand-in for the four TensorFlow names the reference driver touches (example/00_quick_start/sequential.py:
24-25, 440, 481, 499).  Nothing here computes anything."""
from . import compat  # noqa: F401

__version__ = "none (pamrec_b200 shim)"
