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
    raise KeyError(sim)


def rollout_gpu(sim: str, num_worlds: int, num_steps: int, inputs, cfg=None, gpu_id: int = 0):
    """Roll a fixture sim on the B200 engine; returns outputs[name][steps+1, W, ...]."""
    import torch

    desc = SIMS[sim]
    ex = make_executor(sim, num_worlds, gpu_id=gpu_id, **(cfg or {}))
    graph = ex.buildLaunchGraphAllTaskGraphs()
    in_t = {s.name: ex.tensor(s.slot, s.dtype, (num_worlds,) + s.per_world) for s in desc.inputs}
    out_t = {s.name: ex.tensor(s.slot, s.dtype, (num_worlds,) + s.per_world) for s in desc.outputs}
    frames = {s.name: [out_t[s.name].cpu().numpy().copy()] for s in desc.outputs}
    for step in range(num_steps):
        if inputs is not None:
            for s in desc.inputs:
                in_t[s.name].copy_(torch.from_numpy(np.ascontiguousarray(inputs[s.name][step])))
            torch.cuda.synchronize()
        ex.run(graph)
        for s in desc.outputs:
            frames[s.name].append(out_t[s.name].cpu().numpy().copy())
    n_kernels = graph.num_kernels
    del graph
    ex.close()
    return {k: np.stack(v) for k, v in frames.items()}, n_kernels


def golden_path(name: str) -> str:
    return os.path.join(GOLDEN_DIR, name + ".npz")
