import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_path(name):
    return os.path.join(GOLDEN, name + ".npz")


SHIFT_CASES = [
    "p1_c64_h16_centre_b1",
    "p1_c64_h16_irr_b3_tw2p5",
    "p1_c32_h8_empty_b2",
    "p1_c32_h8_full_b1",
    "p2_c64_h16_irr_b2",
    "p3_c32_h8_centre_b1",
    "p1_c512_h16_centre_b1",
    "p1_c256_h32_centre_b1",
]
