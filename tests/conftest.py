import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# tests/_ref holds the reference's own unittest files (staged, git-ignored); they run in their own interpreters from
# test_gpu_reference_suite.py and must not be collected here
collect_ignore = ["_ref"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_assets():
    import oracle_util
    return oracle_util.OracleAssets(lightmap_sizes=(64, 16, 32, 64, 64))


@pytest.fixture(scope="session")
def gpu_ctx():
    from stillleben_b200 import lib
    ctx = lib.Context(0)
    ctx.lightmap_sizes = (64, 16, 32, 64, 64)
    yield ctx
    ctx.close()
