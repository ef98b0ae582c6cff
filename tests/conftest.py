import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
FIX_IDS = os.path.join(GOLDEN, "test.ids")
FIX_CLM = os.path.join(GOLDEN, "test.clm")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _have_gpu():
    try:
        import allhic_b200._ffi as f
        return f.lib().ahx_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly (never skip silently) when selected with -m gpu on a
    # box without a device; they are skipped only when not explicitly selected.
    if config.getoption("-m") and "gpu" in config.getoption("-m") and "not gpu" not in config.getoption("-m"):
        return
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    import __graft_entry__ as g
    g.build()
    import allhic_b200._ffi as f
    return f.lib()


@pytest.fixture(scope="session")
def oracle_fixture():
    from oracle_lib import OracleCLM
    return OracleCLM(FIX_CLM, FIX_IDS)


@pytest.fixture(scope="session")
def synth_small(tmp_path_factory):
    """A small synthetic group (N=300) in the reference's wire formats."""
    from allhic_b200 import synth
    d = tmp_path_factory.mktemp("synth")
    g = synth.generate(300, 777)
    ids, clm = synth.write(g, str(d / "s300"))
    return ids, clm, g


@pytest.fixture(scope="session")
def engine(built_lib):
    import allhic_b200 as A
    return A.Engine(0)
