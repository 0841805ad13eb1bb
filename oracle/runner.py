"""TEST INFRASTRUCTURE ONLY -- runs a fixture simulator on the reference CPU
backend (oracle/_ref/ref_<sim>, built by oracle/Makefile from the reference's
own sources) and parses the dumped trace.  Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import json
import os
import subprocess
import tempfile
from typing import Dict, List, Optional

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))


def binary(sim: str) -> str:
    return os.path.join(_DIR, "_ref", f"ref_{sim}")


def available(sim: str) -> bool:
    return os.path.exists(binary(sim))


def run_reference(desc, num_worlds: int, num_steps: int, inputs: Optional[Dict[str, np.ndarray]],
                  cfg: Dict, workers: int = 1, want_outputs: bool = True, timeout: float = 600.0):
    """inputs[name]: array [steps, worlds, *per_world].  Returns (outputs, timing)
    where outputs[name] has shape [steps + 1, worlds, *per_world] (index 0 =
    state right after construction); dynamic-table slots are lists of
    [rows_t, *per_row] arrays instead."""
    exe = binary(desc.name)
    if not os.path.exists(exe):
        raise FileNotFoundError(f"{exe} missing: run `make -C oracle` where /root/reference exists")
    full = dict(desc.defaults)
    full.update(cfg)
    with tempfile.TemporaryDirectory() as tmp:
        args = [exe, "--worlds", str(num_worlds), "--steps", str(num_steps),
                "--workers", str(workers)]
        for i, v in enumerate(desc.oracle_extra(full)):
            args += [f"--x{i}", str(v)]
        if getattr(desc, "objects", None) is not None:
            from sims.objects import write_blob_file
            blob, relocs = desc.objects()
            obj_path = os.path.join(tmp, "objects.bin")
            write_blob_file(obj_path, blob, relocs)
            args += ["--objects", obj_path]
        if inputs is not None:
            in_path = os.path.join(tmp, "in.bin")
            with open(in_path, "wb") as f:
                for step in range(num_steps):
                    for slot in desc.inputs:
                        arr = np.ascontiguousarray(inputs[slot.name][step], dtype=slot.dtype)
                        assert arr.shape == (num_worlds,) + slot.per_world, (slot.name, arr.shape)
                        f.write(arr.tobytes())
            args += ["--in", in_path]
        out_path = os.path.join(tmp, "out.bin")
        if want_outputs:
            args += ["--out", out_path]
        # The reference's thread pool has a rare lost-wakeup race between
        # ThreadPoolExecutor::run and workerThread (src/mw/cpu_exec.cpp:145-237:
        # workerWakeup is reset by a worker after the main thread may already
        # have re-armed it), which leaves the process asleep forever.  It is not
        # ours to fix: run with a bounded timeout and retry.
        attempts = 4
        per_try = max(30.0, timeout / attempts)
        res = None
        for attempt in range(attempts):
            try:
                res = subprocess.run(args, capture_output=True, text=True, timeout=per_try)
                break
            except subprocess.TimeoutExpired:
                if attempt == attempts - 1:
                    raise
        if res.returncode != 0:
            raise RuntimeError(f"reference run failed: {res.stderr[-2000:]}")
        timing = json.loads(res.stdout.strip().splitlines()[-1])
        outputs = {}
        if want_outputs:
            raw = open(out_path, "rb").read()
            off = 0
            frames: List[Dict[str, np.ndarray]] = []
            for _ in range(num_steps + 1):
                frame = {}
                for slot in desc.outputs:
                    n = int(np.frombuffer(raw, dtype="<u8", count=1, offset=off)[0])
                    off += 8
                    arr = np.frombuffer(raw, dtype=slot.dtype, count=n // np.dtype(slot.dtype).itemsize,
                                        offset=off)
                    off += n
                    frame[slot.name] = arr.reshape((-1,) + slot.per_world)
                frames.append(frame)
            for slot in desc.outputs:
                if slot.dynamic:
                    outputs[slot.name] = [f[slot.name] for f in frames]
                else:
                    outputs[slot.name] = np.stack([f[slot.name] for f in frames])
    return outputs, timing
