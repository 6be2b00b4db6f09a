"""Import helper: the product directory is called ``intel-qs_b200`` (not a Python identifier)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import __graft_entry__ as _g  # noqa: E402

_p = _g.load_package()
capi = _p.capi
circuits = _p.circuits
