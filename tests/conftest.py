import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def pytest_sessionfinish(session, exitstatus):
    """Write every gradient comparison of the session (key, errors, which bar applied) to
    gpurun_out/parity_log.json (BBD_PARITY_LOG overrides the path); profiles/r02_parity.json is a committed copy
    of a B200 run."""
    try:
        from helpers import PARITY_LOG
    except Exception:
        return
    if not PARITY_LOG:
        return
    import json
    path = os.environ.get("BBD_PARITY_LOG", os.path.join(ROOT, "gpurun_out", "parity_log.json"))
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(PARITY_LOG, f, indent=1)
    except OSError:
        pass
