"""CPU: the C-ABI library loads and exports every symbol include/iqsb.h declares; without a GPU
it fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

from pkg import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "iqsb.h")).read()
    return sorted(set(re.findall(r"\b(iqsb_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 60
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in include/iqsb.h but not exported"


def test_binding_covers_every_declared_symbol():
    lib = capi.load()
    src = open(os.path.join(ROOT, "intel-qs_b200", "capi.py")).read()
    for name in declared_symbols():
        assert name in src, f"capi.py does not bind {name}"
        getattr(lib, name)


def test_header_cites_the_reference_for_every_gate_entry_point():
    hdr = open(os.path.join(ROOT, "include", "iqsb.h")).read()
    for key in ["highperfkernels.cpp", "qureg_applyctrl1qubitgate.cpp", "qureg_applyswap.cpp", "qureg_applydiag.cpp", "qureg_fusion.cpp",
                "qureg_measure.cpp", "qureg_expectval.cpp", "qureg_utils.cpp", "qureg_permute.cpp", "qureg_init.cpp", "mpi_env"]:
        assert key in hdr


def test_version_and_error_string():
    lib = capi.load()
    assert lib.iqsb_version() >= 100
    assert isinstance(lib.iqsb_last_error(), bytes)


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to start (run on the CPU box only)."""
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(capi.IqsbError, match="no CUDA device"):
        capi.Context()


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the product directory may reference it."""
    prod = os.path.join(ROOT, "intel-qs_b200")
    for dirpath, _, files in os.walk(prod):
        for f in files:
            if f.endswith((".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "iqs_oracle" not in txt and "oracle_run" not in txt, f
    for f in ("capi.py", "circuits.py", "__init__.py"):
        txt = open(os.path.join(prod, f)).read()
        assert "import oracle" not in txt and "load_oracle" not in txt and "liboracle" not in txt, f
