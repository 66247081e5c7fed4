import os
import sys

import pytest

# The sharded-proof tests run up to 8 ranks as contexts on ONE GPU, each with a compute and a copy
# stream: with the default 8 hardware work queues two ranks' streams share a queue, and a copy queued
# behind another rank's spinning barrier kernel never starts (a false dependency that cannot occur with
# one rank per GPU).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu under gpurun)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (checker only)."""
    from oracle import stark_oracle as so

    so.lib()
    return so


@pytest.fixture(scope="session")
def ctx():
    """A canonical-form aero_b200 context on cuda:0; GPU tests fail loudly if the library or the
    device is unusable (no silent fallback)."""
    import aero_b200

    c = aero_b200.Context(0, form=aero_b200.AERO_FORM_CANONICAL)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx_mont():
    import aero_b200

    c = aero_b200.Context(0, form=aero_b200.AERO_FORM_MONTGOMERY)
    yield c
    c.close()


GOLDEN = os.path.join(ROOT, "tests", "golden")
