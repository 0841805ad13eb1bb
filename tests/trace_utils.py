"""Shared helpers for parity tests: seeded synthetic inputs, GPU roll-outs
through the C ABI, golden fixture I/O."""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

from sims import SIMS, make_executor

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def make_inputs(sim: str, num_worlds: int, num_steps: int, seed: int = 0) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    if sim == "cartpole":
        return {
            "reset": (rng.random((num_steps, num_worlds, 1)) < 0.01).astype(np.int32),
            "action": rng.integers(0, 2, size=(num_steps, num_worlds, 1), dtype=np.int32),
        }
    if sim == "gridworld":
        return {
            "reset": (rng.random((num_steps, num_worlds, 1)) < 0.02).astype(np.int32),
            "action": rng.integers(0, 5, size=(num_steps, num_worlds, 2), dtype=np.int32),
        }
    if sim in ("room", "room_tgs"):
        # mostly "full speed ahead" so agents reach cubes, walls and each other
        amount = np.where(rng.random((num_steps, num_worlds, 2)) < 0.7, 3,
                          rng.integers(0, 4, size=(num_steps, num_worlds, 2)))
        angle = np.where(rng.random((num_steps, num_worlds, 2)) < 0.7, 0,
                         rng.integers(0, 8, size=(num_steps, num_worlds, 2)))
        act = np.stack([amount, angle,
                        rng.integers(0, 5, size=(num_steps, num_worlds, 2))], axis=-1)
        return {
            "reset": (rng.random((num_steps, num_worlds, 1)) < 0.005).astype(np.int32),
            "action": act.astype(np.int32),
        }
    if sim == "arena":
        shape = (num_steps, num_worlds, 6)
        amount = np.where(rng.random(shape) < 0.6, 3, rng.integers(0, 4, size=shape))
        angle = np.where(rng.random(shape) < 0.6, 0, rng.integers(0, 8, size=shape))
        act = np.stack([amount, angle, rng.integers(0, 5, size=shape),
                        (rng.random(shape) < 0.15).astype(np.int64)], axis=-1)
        return {
            "reset": (rng.random((num_steps, num_worlds, 1)) < 0.004).astype(np.int32),
            "action": act.astype(np.int32),
        }
    if sim in ("balls", "balls_many"):
        return {}          # physics only: no inputs
    raise KeyError(sim)


def rollout_gpu(sim: str, num_worlds: int, num_steps: int, inputs, cfg=None, gpu_id: int = 0):
    """Roll a fixture sim on the B200 engine; returns outputs[name][steps+1, W, ...]."""
    import torch

    desc = SIMS[sim]
    ex = make_executor(sim, num_worlds, gpu_id=gpu_id, **(cfg or {}))
    graph = ex.buildLaunchGraphAllTaskGraphs()
    in_t = {s.name: ex.tensor(s.slot, s.dtype, (num_worlds,) + s.per_world) for s in desc.inputs}
    out_t = {s.name: ex.tensor(s.slot, s.dtype, (num_worlds,) + s.per_world)
             for s in desc.outputs if not s.dynamic}

    def grab(frames):
        for s in desc.outputs:
            if s.dynamic:
                rows = ex.exportedNumRows(s.slot)
                t = ex.tensor(s.slot, s.dtype, (max(rows, 1),) + s.per_world)
                frames[s.name].append(t.cpu().numpy()[:rows].copy())
            else:
                frames[s.name].append(out_t[s.name].cpu().numpy().copy())

    frames = {s.name: [] for s in desc.outputs}
    grab(frames)
    for step in range(num_steps):
        if inputs is not None:
            for s in desc.inputs:
                in_t[s.name].copy_(torch.from_numpy(np.ascontiguousarray(inputs[s.name][step])))
            torch.cuda.synchronize()
        ex.run(graph)
        grab(frames)
    n_kernels = graph.num_kernels
    del graph
    ex.close()
    dyn = {s.name for s in desc.outputs if s.dynamic}
    return {k: (v if k in dyn else np.stack(v)) for k, v in frames.items()}, n_kernels


def save_golden(name, inputs, outs, W, steps):
    payload = {"in_" + k: v for k, v in inputs.items()}
    for k, v in outs.items():
        if isinstance(v, list):
            payload["dyn_" + k] = np.concatenate(v) if v else np.zeros((0,))
            payload["dynlen_" + k] = np.array([len(f) for f in v], dtype=np.int64)
        else:
            payload["out_" + k] = v
    payload["meta"] = np.array([W, steps], dtype=np.int64)
    np.savez_compressed(golden_path(name), **payload)


def load_golden(name):
    z = np.load(golden_path(name))
    ins = {k[3:]: z[k] for k in z.files if k.startswith("in_")}
    outs = {k[4:]: z[k] for k in z.files if k.startswith("out_")}
    for k in z.files:
        if k.startswith("dyn_"):
            lens = z["dynlen_" + k[4:]]
            offs = np.concatenate([[0], np.cumsum(lens)])
            outs[k[4:]] = [z[k][offs[i]:offs[i + 1]] for i in range(len(lens))]
    W, steps = (int(v) for v in z["meta"])
    return W, steps, ins, outs


def assert_traces_equal(got, want, exact=True, rtol=1e-4, atol=1e-6):
    for k, w in want.items():
        g = got[k]
        if isinstance(w, list):
            assert len(g) == len(w), k
            for t, (gf, wf) in enumerate(zip(g, w)):
                assert gf.shape == wf.shape, f"{k} frame {t}: rows {gf.shape} vs {wf.shape}"
                assert np.array_equal(gf, wf), f"{k} frame {t} differs"
        elif exact or not np.issubdtype(w.dtype, np.floating):
            same = g.view(np.uint8) == w.view(np.uint8) if g.dtype.itemsize == 1 else g == w
            if np.issubdtype(w.dtype, np.floating):
                same = g.view(np.uint32) == w.view(np.uint32)
            assert same.all(), f"{k}: first mismatch at {np.argwhere(~same)[0]}"
        else:
            np.testing.assert_allclose(g, w, rtol=rtol, atol=atol, err_msg=k)


def golden_path(name: str) -> str:
    return os.path.join(GOLDEN_DIR, name + ".npz")
