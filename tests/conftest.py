import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


@pytest.fixture
def switches(monkeypatch):
    """Set / clear the library's ZS_* A/B switches for one test.  They are read once when a context is created (never on a
    launch path), so a live context is told to re-read them; everything is restored at teardown."""
    touched = []

    class Switches:
        def set(self, ctx, name, value="1"):
            monkeypatch.setenv(name, str(value)); ctx.reload_switches(); touched.append(ctx)

        def clear(self, ctx, name):
            monkeypatch.delenv(name, raising=False); ctx.reload_switches(); touched.append(ctx)

    yield Switches()
    monkeypatch.undo()
    for c in touched:
        c.reload_switches()


def has_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False
