"""NVLink peer-store gather (include/madrona_b200.h mb2_peer_gather_*): single-rank
loop-back on one GPU (the N > 1 path is exercised by `bench.py --gpus N`; the
host-side sharding logic by tests/test_sharding.py with gloo)."""
import numpy as np
import pytest


@pytest.mark.gpu
def test_loopback_gather_matches_exported_columns():
    import torch
    from sims import SIMS, make_executor
    from trace_utils import make_inputs

    W, steps = 257, 9
    ex = make_executor("cartpole", W, max_steps=50, seed=4)
    graph = ex.buildLaunchGraphAllTaskGraphs()
    desc = SIMS["cartpole"]
    outs = [s for s in desc.outputs]
    pg = ex.peerGather([s.slot for s in outs], [(W,) + s.per_world for s in outs], [s.dtype for s in outs], 1, 0)
    pg.connect([pg.local_handle()])
    ins = make_inputs("cartpole", W, steps, seed=2)
    act = ex.tensor(1, "int32", (W, 1))
    stream = torch.cuda.current_stream()
    cols = [ex.tensor(s.slot, s.dtype, (W,) + s.per_world) for s in outs]
    for t in range(steps):
        act.copy_(torch.from_numpy(np.ascontiguousarray(ins["action"][t])))
        ex.runAsync(graph, stream)
        pg.push(stream)
        pg.wait(stream)
        torch.cuda.synchronize()
        for i, c in enumerate(cols):
            got = pg.tensor(t & 1, i)
            assert got.shape == c.shape and torch.equal(got, c), (t, outs[i].name)
        pg.release(stream)
    pg.close()
    ex.close()
