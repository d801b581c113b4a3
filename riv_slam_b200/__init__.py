"""Importable alias for the hyphenated package directory ``riv-slam_b200/``.

``riv-slam_b200`` is not a valid Python identifier, so the sources live there and this shim
extends the package search path to it: ``import riv_slam_b200.datagen`` resolves to
``riv-slam_b200/datagen.py``.
"""
import os as _os

_SRC = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "riv-slam_b200")
__path__.append(_SRC)
REPO_ROOT = _os.path.dirname(_SRC)
PKG_DIR = _SRC
