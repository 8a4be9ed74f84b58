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


def act_dtype():
    """torch dtype of the product's 16-bit tensors under the process's precision policy (fp16, or bf16 when the suite
    is re-run with TEXTBOOST_B200_PRECISION=bf16 by tests/test_gpu_bf16_policy.py)."""
    from textboost_b200.precision import POLICY
    return POLICY.act


def tol_scale() -> float:
    """Every 16-bit tolerance in the GPU tests is written for fp16 (unit roundoff 2^-11); the bf16 build rounds at 2^-8,
    so its errors against the same fp32 statement are allowed 8x those bounds."""
    from textboost_b200.precision import POLICY
    return 8.0 if POLICY.name == "bf16" else 1.0
