import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _chdir_repo_root():
    # scene files reference textures / sub-scenes relative to the CWD, like the reference CLI
    old = os.getcwd()
    os.chdir(ROOT)
    yield
    os.chdir(old)


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    from nr_ray_tracer_b200 import build as B
    from oracle import oracle as O
    B.build()
    O.build()


@pytest.fixture(scope="session")
def gpu_ctx():
    from nr_ray_tracer_b200 import api
    ctx = api.Context(0)  # raises loudly without a GPU: there is no fallback
    yield ctx
    ctx.close()
