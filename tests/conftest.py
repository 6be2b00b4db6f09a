import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (HERE, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # built artefacts are not in git: a fresh checkout compiles them once (nvcc cross-compiles without a GPU)
    lib = os.path.join(ROOT, "intel-qs_b200", "lib")
    need = ["libiqs_b200.so", "libiqs.so"]
    if not all(os.path.exists(os.path.join(lib, n)) for n in need) or not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        import __graft_entry__ as g

        g.build()


@pytest.fixture(scope="session")
def oracle():
    import __graft_entry__ as g

    return g.load_oracle()


@pytest.fixture(scope="session")
def gpu_ctx():
    """One engine context for the whole session.  Fails loudly when the CUDA library or the
    device is missing -- there is no fallback to hide behind."""
    from pkg import capi

    ctx = capi.Context()
    yield ctx
    ctx.close()
