import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The shared library is a build artefact (git-ignored): build it when a fresh checkout runs the tests
    before __graft_entry__.build().  nvcc cross-compiles without a GPU."""
    lib = os.path.join(ROOT, "audio_video_textures_b200", "libavtex.so")
    if not os.path.exists(lib):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "audio_video_textures_b200", "csrc"), "-j8"], check=True,
                       stdout=subprocess.DEVNULL)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden


CLASSIC_GOLDEN = ["classic_small_m1", "classic_ragged_m2", "classic_stride_m3", "classic_c1"]
