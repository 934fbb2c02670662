"""B200-native IPDM domain-progressive inference path (host side).

Layout of this directory
  csrc/                   hand-written sm_100a CUDA + the C ABI (include/ipdm_b200.h) -> libipdm_b200.so
  _lib.py, engine.py      ctypes binding and thin tensor-level wrappers
  Config/ Utils/ Model/ Recon/ Dataset/ main.py
                          mirror of the reference's module paths for this path, so that with this
                          directory on sys.path `main.py` and notebook cells 0-2 run unchanged
  synthetic.py            seeded synthetic sinograms / phantoms (no Mayo data offline)

There is no CPU fallback: importing `engine` without a built libipdm_b200.so raises.
"""
import os
import sys

PACKAGE_DIR = os.path.dirname(os.path.abspath(__file__))


def add_reference_paths():
    """Put this directory first on sys.path so `from Utils.train_test_utils import ...` resolves here."""
    if PACKAGE_DIR not in sys.path:
        sys.path.insert(0, PACKAGE_DIR)
    return PACKAGE_DIR
