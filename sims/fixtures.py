from __future__ import annotations

import os
import struct
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Sequence, Tuple

import numpy as np

_ROOT = os.path.dirname(os.path.abspath(__file__))


@dataclass
class Slot:
    slot: int
    name: str
    dtype: str                 # numpy dtype name
    per_world: Tuple[int, ...]  # shape of one world's rows (fixed tables) / of one row (dynamic)
    dynamic: bool = False      # rows = live rows of a dynamic table (varies per step)


@dataclass
class SimDesc:
    name: str
    sources: List[str]
    num_exports: int
    num_taskgraphs: int
    inputs: List[Slot]
    outputs: List[Slot]
    # (cfg dict) -> bytes of the simulator's Config struct
    pack_config: Callable[[Dict], bytes]
    # (world index, cfg dict) -> bytes of one WorldInit struct
    pack_init: Callable[[int, Dict], bytes]
    # extra args for the oracle harness (--x0 .. --x3)
    oracle_extra: Callable[[Dict], List[int]]
    defaults: Dict = field(default_factory=dict)


def _cartpole_cfg(cfg):
    return struct.pack("<I", int(cfg.get("max_steps", 200)))


def _cartpole_init(w, cfg):
    return struct.pack("<I", int(cfg.get("seed", 0)) + w)


def _grid_cfg(cfg):
    return struct.pack("<iii", int(cfg["grid_size"]), int(cfg["episode_len"]), int(cfg["init_items"]))


def _grid_init(w, cfg):
    return struct.pack("<I", int(cfg.get("seed", 0)) + w)


SIMS: Dict[str, SimDesc] = {
    "gridworld": SimDesc(
        name="gridworld",
        sources=[os.path.join(_ROOT, "gridworld", "sim.cpp")],
        num_exports=10,
        num_taskgraphs=1,
        inputs=[Slot(0, "reset", "int32", (1,)), Slot(1, "action", "int32", (2,))],
        outputs=[Slot(2, "agent_pos", "int32", (2, 2)), Slot(3, "reward", "float32", (2,)),
                 Slot(4, "obs", "int32", (2, 4)), Slot(5, "item_count", "int32", (1,)),
                 Slot(9, "done", "int32", (1,)),
                 Slot(6, "item_pos", "int32", (2,), dynamic=True),
                 Slot(7, "item_entity", "int32", (2,), dynamic=True),
                 Slot(8, "item_kind", "int32", (1,), dynamic=True)],
        pack_config=_grid_cfg,
        pack_init=_grid_init,
        oracle_extra=lambda cfg: [int(cfg["grid_size"]), int(cfg["episode_len"]),
                                  int(cfg["init_items"]), int(cfg.get("seed", 0))],
        defaults={"grid_size": 8, "episode_len": 50, "init_items": 6, "seed": 0},
    ),
    "cartpole": SimDesc(
        name="cartpole",
        sources=[os.path.join(_ROOT, "cartpole", "sim.cpp")],
        num_exports=5,
        num_taskgraphs=1,
        inputs=[Slot(0, "reset", "int32", (1,)), Slot(1, "action", "int32", (1,))],
        outputs=[Slot(2, "state", "float32", (4,)), Slot(3, "reward", "float32", (1,)),
                 Slot(4, "done", "int32", (1,))],
        pack_config=_cartpole_cfg,
        pack_init=_cartpole_init,
        oracle_extra=lambda cfg: [int(cfg.get("max_steps", 200)), int(cfg.get("seed", 0))],
        defaults={"max_steps": 200, "seed": 0},
    ),
}


def pack_world_inits(desc: SimDesc, num_worlds: int, cfg: Dict) -> bytes:
    return b"".join(desc.pack_init(w, cfg) for w in range(num_worlds))


def make_executor(name: str, num_worlds: int, gpu_id: int = 0, **cfg):
    """Build a madrona_b200.MWCudaExecutor for a fixture sim (needs a B200)."""
    import madrona_b200 as mb

    desc = SIMS[name]
    full = dict(desc.defaults)
    full.update(cfg)
    inits = pack_world_inits(desc, num_worlds, full)
    state = mb.StateConfig(
        worldInit=inits,
        numWorldInitBytes=len(inits) // num_worlds,
        userConfig=desc.pack_config(full),
        numWorldDataBytes=0,          # 0 => sizeof(Sim) as the device compiler sees it
        worldDataAlignment=16,
        numWorlds=num_worlds,
        numTaskGraphs=desc.num_taskgraphs,
        numExportedBuffers=desc.num_exports,
    )
    compile_cfg = mb.CompileConfig(userSources=desc.sources,
                                   userCompileFlags=["-I" + os.path.dirname(desc.sources[0])])
    return mb.MWCudaExecutor(state, compile_cfg, gpu_id=gpu_id)
