import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _resources():
    """The bundled scenes reference resources/*.png; generate them when missing."""
    res = os.path.join(ROOT, "resources")
    if not os.path.exists(os.path.join(res, "celestial.png")):
        sys.path.insert(0, os.path.join(ROOT, "scenes"))
        import make_textures
        make_textures.main(res)
    return res
