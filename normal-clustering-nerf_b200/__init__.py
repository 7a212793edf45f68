"""ncn_b200 - B200-native (sm_100a) hot path of normal-clustering-nerf.

Layout
  csrc/               hand-written CUDA kernels + the C-ABI (include/ncn.h) -> libncn.so
  _lib.py             ctypes binding of libncn.so (fails loudly when it is missing)
  vren.py             drop-in for the reference's `vren` pybind module (models/csrc/binding.cpp:330-350)
  custom_functions.py drop-in for models/custom_functions.py (same autograd classes)
  tinycudann.py       drop-in for the tcnn `Encoding` / `Network` modules used by models/ngp_mt.py
  shims/              directory to put on sys.path so that the reference's unchanged files resolve
                      `import vren`, `import tinycudann`, `import torch_scatter`, `import faiss`
"""
__version__ = "0.1.0"

import os as _os

PACKAGE_DIR = _os.path.dirname(_os.path.abspath(__file__))
SHIMS_DIR = _os.path.join(PACKAGE_DIR, "shims")


def install_shims():
    """Put the drop-in module names (vren, tinycudann, torch_scatter, faiss) on sys.path."""
    import sys
    if SHIMS_DIR not in sys.path:
        sys.path.insert(0, SHIMS_DIR)
