"""Importable alias for the package directory ``normal-clustering-nerf_b200/``.

The directory name (mandated by the repo layout) contains hyphens, so it cannot be
imported by name.  ``import ncn_b200`` loads it as a regular package under the module
name ``ncn_b200`` (sub-modules: ``ncn_b200.vren``, ``ncn_b200.tinycudann`` ...).
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "normal-clustering-nerf_b200")
_spec = importlib.util.spec_from_file_location(
    "ncn_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ncn_b200"] = _mod
_spec.loader.exec_module(_mod)
