"""intel-qs_b200: B200-native (sm_100a) state-vector engine behind the Intel-QS API.

The product is native code: ``csrc/`` (CUDA kernels + the C ABI of include/iqsb.h, built into
``lib/libiqs_b200.so``) and ``include/`` + ``src/`` (the re-authored C++ ``iqs::QubitRegister``
host API, built into ``lib/libiqs.so``, and the ``intelqs_py`` pybind11 module).  The Python in
this package is harness plumbing only: a ctypes binding (``capi``), circuit descriptions
(``circuits``) and the build recipe (``build``).
"""
from . import capi, circuits  # noqa: F401
from .build import build_all  # noqa: F401
