import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): plain-C restatement of the reference hot path."""
    import oracle as orc
    orc.build()
    return orc


@pytest.fixture(scope="session")
def ctx():
    """One libofps_b200 context on cuda:0 — creation fails loudly without a B200."""
    if os.environ.get("OFPSB_EMU_CTX") == "1":   # dry run of the cv-front GPU tests on the CPU emulation (tests/emu)
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        import fake_ctx
        yield fake_ctx.EmuContext()
        return
    from ofps_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()
