import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(autouse=True)
def _reset_library_options():
    """Tests change library switches with T.set_option (the library never reads the environment); every test starts
    from the defaults."""
    yield
    try:
        from thaler_study_b200 import reset_options

        reset_options()
    except Exception:
        pass
