"""Generates tests/golden/*.npz by running the fixture simulators on the
*reference* CPU backend (oracle/_ref, built from /root/reference by
oracle/Makefile).  Run where /root/reference exists:

    make -C oracle && python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.runner import run_reference  # noqa: E402
from sims import SIMS  # noqa: E402
from trace_utils import make_inputs, golden_path  # noqa: E402

CASES = [
    # (file, sim, worlds, steps, cfg)
    ("cartpole_w64_s300", "cartpole", 64, 300, {"max_steps": 200, "seed": 0}),
    ("cartpole_w3_s50", "cartpole", 3, 50, {"max_steps": 10, "seed": 7}),
]

if __name__ == "__main__":
    only = sys.argv[1:]
    for name, sim, W, steps, cfg in CASES:
        if only and name not in only:
            continue
        inputs = make_inputs(sim, W, steps, seed=1234)
        outs, _ = run_reference(SIMS[sim], W, steps, inputs, cfg, workers=1)
        payload = {"in_" + k: v for k, v in inputs.items()}
        payload.update({"out_" + k: v for k, v in outs.items()})
        payload["meta"] = np.array([W, steps], dtype=np.int64)
        np.savez_compressed(golden_path(name), **payload)
        print(name, {k: v.shape for k, v in outs.items()})
