"""Makes `ipdm_pytorch_b200` importable when only this directory is on sys.path (main.py, notebook)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.append(_root)
import ipdm_pytorch_b200  # noqa: E402,F401
import ipdm_pytorch_b200.engine as engine  # noqa: E402,F401
