"""The committed golden vectors must be reproduced by the oracle (guards oracle edits)."""
import os

import numpy as np
import pytest

from tests.cases import GOLDEN_CASES, oracle_pair

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("case", GOLDEN_CASES, ids=[c["name"] for c in GOLDEN_CASES])
def test_oracle_reproduces_golden(case):
    g = np.load(os.path.join(GOLD, case["name"] + ".npz"))
    ref = oracle_pair(case)
    for k in ("mixed", "speech", "noise"):
        assert g[k].shape == ref[k].shape
        assert np.max(np.abs(g[k] - ref[k])) < 2e-5  # float32 storage of values up to ~1e2 dB
    scale = np.max(np.abs(ref["mixed_pcm"]))
    assert np.max(np.abs(g["mixed_pcm"] - ref["mixed_pcm"])) < 1e-6 * scale
    assert np.max(np.abs(g["recon"] - ref["recon"])) < 1e-6 * scale
