"""Importable alias of the `composable-sdr_b200/` package directory (a hyphen is not a valid module name).

All code lives in ../composable-sdr_b200/; this shim only points the package search path there.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "..", "composable-sdr_b200"))

from .blocks import *  # noqa: F401,F403,E402
from . import blocks, build, synth  # noqa: F401,E402
