"""The C++ facade (madrona_b200/host/madrona/mw_gpu.hpp) compiles a
reference-style Manager snippet unchanged; failures abort like the reference's
FATAL()."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "facade_cartpole")
    cmd = ["g++", "-std=c++20", "-O1", "-I" + os.path.join(ROOT, "madrona_b200", "host"),
           os.path.join(ROOT, "tests", "cpp", "facade_cartpole.cpp"), "-o", exe,
           "-L" + os.path.join(ROOT, "madrona_b200"), "-lmadrona_b200",
           "-Wl,-rpath," + os.path.join(ROOT, "madrona_b200"),
           "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.run(cmd, check=True, capture_output=True)
    return exe


def test_facade_compiles_and_fails_loudly_without_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    res = subprocess.run([exe, os.path.join(ROOT, "sims", "cartpole", "sim.cpp")],
                         capture_output=True, text=True)
    assert res.returncode != 0
    assert "madrona_b200" in res.stderr


@pytest.mark.gpu
def test_facade_runs_cartpole(tmp_path):
    from trace_utils import rollout_gpu
    exe = _build(tmp_path)
    res = subprocess.run([exe, os.path.join(ROOT, "sims", "cartpole", "sim.cpp")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    got = np.array([float(v) for v in res.stdout.split()[1:5]], dtype=np.float32)
    ins = {"reset": np.zeros((10, 64, 1), np.int32), "action": np.zeros((10, 64, 1), np.int32)}
    ref, _ = rollout_gpu("cartpole", 64, 10, ins, {"max_steps": 200, "seed": 0})
    assert np.allclose(got, ref["state"][10, 0], rtol=0, atol=0)
