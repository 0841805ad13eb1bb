"""Code paths kept behind switches for A/B measurements must stay correct: each runs the
parity tests of the kernels it touches in a child process (the switches are read once per
process / select a different JIT build).

  MADRONA_B200_SORT_FUSE_COPYBACK=1   copy-back of exported columns as work items of the
                                      rearrange kernel (default: separate launch)
  MADRONA_B200_JIT_DEFINES=-DMB2_TRACE_MASK=0
                                      BVH::traceRay as the one-phase ordered scan (default:
                                      candidate mask + ordered re-test, device/madrona/physics.hpp)
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    "fused_copy_back": ({"MADRONA_B200_SORT_FUSE_COPYBACK": "1"},
                        ["tests/test_sort_custom_key.py", "tests/test_gridworld.py"], None),
    "trace_ray_one_phase_scan": ({"MADRONA_B200_JIT_DEFINES": "-DMB2_TRACE_MASK=0"},
                                 ["tests/test_room.py"], "gpu_matches_golden"),
}


@pytest.mark.gpu
@pytest.mark.parametrize("case", sorted(CASES))
def test_parity_holds_with_switch(case):
    env, targets, select = CASES[case]
    cmd = [sys.executable, "-m", "pytest", *targets, "-m", "gpu", "-q", "-x", "--tb=short", "-p", "no:cacheprovider"]
    if select:
        cmd += ["-k", select]
    res = subprocess.run(cmd, cwd=ROOT, env={**os.environ, **env}, capture_output=True, text=True, timeout=900)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail      # 0 = every selected test passed
    assert " passed" in res.stdout, tail    # ... and something was selected
