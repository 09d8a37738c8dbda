import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _build_oracle():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build as oracle_build
    return oracle_build.build()


@pytest.fixture(scope="session")
def oracle():
    _build_oracle()
    import oracle_binding
    return oracle_binding.Oracle()


@pytest.fixture(scope="session")
def engine_lib():
    import synthesis_b200
    return synthesis_b200


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def engine(engine_lib):
    """One engine for the GPU tests.  Fails (does not skip) when the CUDA library cannot run: a GPU
    test that silently passes without the native code would void the parity claim."""
    eng = engine_lib.Engine(device=0, max_games_in_flight=4736, max_explores=1600)
    # the parity suite plays tens to hundreds of games per call: left to itself the engine would run all of them on the lane-group
    # kernels it prefers for small batches.  The suite pins the thread-per-game kernels (the large-batch product path);
    # tests/test_gpu_parity.py::test_mapping_chosen_per_launch_does_not_change_results covers the automatic choice.
    eng.set_group_lanes(1)
    yield eng
    eng.close()
