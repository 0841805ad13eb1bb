"""smoke(): one small rigid-body invocation on cuda:0 checked against the
reference CPU backend's golden trace."""
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def smoke_extra():
    from trace_utils import assert_traces_equal, load_golden, rollout_gpu
    W, steps, ins, outs = load_golden("room_w4_s210")
    steps = 110   # one full episode + reset
    ins = {k: v[:steps] for k, v in ins.items()}
    outs = {k: (v[:steps + 1]) for k, v in outs.items()}
    got, n_kernels = rollout_gpu("room", W, steps, ins, {"episode_len": 100, "seed": 21})
    assert n_kernels > 10
    assert_traces_equal(got, outs)
