import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """Builds (or reuses) the in-tree native libraries once per session."""
    from jsplayer_b200 import build
    from oracle import pyoracle
    build.build_all()
    import synth
    synth.build()
    pyoracle.build()


def load_golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name))
    ln = z["frame_len"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum(ln)]).astype(np.int64)
    frames = [z["data"][int(off[i]):int(off[i + 1])].tobytes() for i in range(len(ln))]
    return z, frames


GOLDEN_MSV1 = ["msv1_ffmpeg_rgb555_64x48.npz", "msv1_ffmpeg_rgb555_320x240.npz",
               "msv1_ffmpeg_pal8_64x48.npz", "msv1_ffmpeg_pal8_320x240.npz"]
