"""Regenerates tests/golden/*.npz from the float64 oracle (oracle/avse_oracle.py).

The reference holds no golden vectors and cannot be imported here (librosa / mediaio absent), so
these fixtures pin the ORACLE's outputs on seeded inputs: they guard the oracle against silent
edits and give the GPU box a /root/reference-free, oracle-independent target.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.cases import GOLDEN_CASES, oracle_pair  # noqa: E402

if __name__ == "__main__":
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for case in GOLDEN_CASES:
        ref = oracle_pair(case)
        path = os.path.join(out_dir, case["name"] + ".npz")
        np.savez_compressed(path, **{k: np.asarray(v, dtype=np.float32) for k, v in ref.items()})
        print(path, {k: v.shape for k, v in ref.items()}, os.path.getsize(path))
