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

# shift_sz != 1 / stride != 1: forward outputs of the reference (oracle/make_golden.py run_patch_case)
PATCH_CASES = [
    "k3_c32_h8_irr_b1",
    "k3_c16_h16_irr_b2_t5",
    "k2s2_c16_h16_irr_b1",
    "k4s2_c16_h12_irr_b1_t3",
    "k3_c64_h16_centre_b1",
]

# full-size cases stored compactly (seed + fingerprints of the reference's outputs): oracle/make_golden.py run_case_compact
COMPACT_CASES = [
    "p1_c256_h64_centre_b1_compact",
    "p2_c256_h64_centre_b1_compact",
    "p1_c512_h32_centre_b1_compact",
    "p2_c512_h32_irr_b2_compact",
]
