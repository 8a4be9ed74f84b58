import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu through gpurun)")
    config.addinivalue_line("markers", "long_cpu: minutes on the CPU (whole steps through the block emulator); run with "
                                       "TB_LONG_CPU_TESTS=1.  The default CPU suite keeps one whole-step case of each "
                                       "kind and stays within a few minutes.")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("TB_LONG_CPU_TESTS"):
        return
    skip = pytest.mark.skip(reason="long CPU test: set TB_LONG_CPU_TESTS=1")
    for item in items:
        if "long_cpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def built_lib():
    from textboost_b200 import build
    return build.build()
