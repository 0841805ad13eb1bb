"""Python mirror of madrona::py::Tensor (include/madrona/py/utils.hpp:58-141,
python binding src/python/bindings.cpp:332-361): a typed, shaped view of a
buffer a simulator exports -- what a sim's `Manager.*_tensor()` returns -- with
`to_torch()` handing it to PyTorch without a copy.

    t = Tensor(ex.getExported(slot), TensorElementType.Float32, [num_worlds, 2, 3], gpu_id=0)
    actions = t.to_torch()          # torch.float32 [num_worlds, 2, 3] on cuda:0, aliasing the column

A CPU buffer (gpu_id None, e.g. a pinned staging area) is viewed through numpy.
"""
from __future__ import annotations

import ctypes
import enum
from typing import Optional, Sequence

import numpy as np


class TensorElementType(enum.Enum):
    """Same members, same order as madrona::py::TensorElementType."""
    UInt8 = 0
    Int8 = 1
    Int16 = 2
    Int32 = 3
    Int64 = 4
    Float16 = 5
    Float32 = 6


_NUMPY = {
    TensorElementType.UInt8: np.uint8, TensorElementType.Int8: np.int8,
    TensorElementType.Int16: np.int16, TensorElementType.Int32: np.int32,
    TensorElementType.Int64: np.int64, TensorElementType.Float16: np.float16,
    TensorElementType.Float32: np.float32,
}
_FROM_NAME = {np.dtype(v).name: k for k, v in _NUMPY.items()}


class Tensor:
    maxDimensions = 16

    def __init__(self, dev_ptr: int, type: TensorElementType, dimensions: Sequence[int],
                 gpu_id: Optional[int] = None):
        if len(dimensions) > Tensor.maxDimensions:
            raise ValueError(f"Cannot construct Tensor with more than {Tensor.maxDimensions} dimensions")
        self._ptr = int(dev_ptr)
        self._type = TensorElementType(type)
        self._dims = tuple(int(d) for d in dimensions)
        self._gpu_id = -1 if gpu_id is None else int(gpu_id)

    @classmethod
    def from_torch(cls, tensor) -> "Tensor":
        """The reference constructs Tensor from any dlpack-capable array (bindings.cpp:333-357)."""
        name = str(tensor.dtype).replace("torch.", "")
        if name not in _FROM_NAME:
            raise TypeError(f"Tensor: Invalid tensor dtype {tensor.dtype}")
        if not tensor.is_contiguous():
            raise ValueError("Tensor: only dense row-major tensors can be wrapped")
        gpu_id = tensor.device.index if tensor.device.type == "cuda" else None
        return cls(tensor.data_ptr(), _FROM_NAME[name], tensor.shape, gpu_id)

    # --- accessors named like the C++ class -----------------------------------------------
    def devicePtr(self) -> int:
        return self._ptr

    def type(self) -> TensorElementType:
        return self._type

    def isOnGPU(self) -> bool:
        return self._gpu_id != -1

    def gpuID(self) -> int:
        return self._gpu_id

    def numDims(self) -> int:
        return len(self._dims)

    def dims(self):
        return self._dims

    def numBytesPerItem(self) -> int:
        return np.dtype(_NUMPY[self._type]).itemsize

    def numBytes(self) -> int:
        n = self.numBytesPerItem()
        for d in self._dims:
            n *= d
        return n

    # --- zero-copy views ----------------------------------------------------------------------
    @property
    def __cuda_array_interface__(self):
        if not self.isOnGPU():
            raise AttributeError("host tensor")
        return {"shape": self._dims, "typestr": np.dtype(_NUMPY[self._type]).str,
                "data": (self._ptr, False), "version": 2, "strides": None}

    def to_numpy(self) -> np.ndarray:
        """Host buffers only: a numpy array aliasing the memory."""
        if self.isOnGPU():
            raise ValueError("to_numpy() needs a host tensor; use to_torch().cpu() for GPU tensors")
        buf = (ctypes.c_char * self.numBytes()).from_address(self._ptr)
        return np.frombuffer(buf, dtype=_NUMPY[self._type]).reshape(self._dims)

    def to_torch(self):
        import torch
        if self.isOnGPU():
            return torch.as_tensor(self, device=f"cuda:{self._gpu_id}")
        return torch.from_numpy(self.to_numpy())


# ---- madrona::py::TrainInterface (include/madrona/py/utils.hpp:143-201) ----------------------
# What a simulator's Manager hands to a learner: named action tensors, resets and
# simCtrl going in; observations, rewards, dones (+ stats / pbt) coming out.  The
# reference's JAX bridge (src/python/bindings.cpp:80-283) turns these into XLA custom
# calls; here the same structure is plain data plus the copy helpers, with torch
# doing the device copies.

from dataclasses import dataclass, field
from typing import Dict, List


@dataclass
class NamedTensor:
    name: str
    tensor: Tensor


@dataclass
class TrainStepInputInterface:
    actions: List[NamedTensor]
    resets: Tensor
    simCtrl: Optional[Tensor] = None
    pbt: List[NamedTensor] = field(default_factory=list)


@dataclass
class TrainStepOutputInterface:
    observations: List[NamedTensor]
    rewards: Tensor
    dones: Tensor
    stats: List[NamedTensor] = field(default_factory=list)
    pbt: List[NamedTensor] = field(default_factory=list)


@dataclass
class TrainCheckpointingInterface:
    checkpointData: Tensor


class TrainInterface:
    """Mirror of madrona::py::TrainInterface: stepInputs() / stepOutputs() /
    checkpointing(), and the copy helpers in the reference's buffer order
    (src/python/utils.cpp: inputs = actions..., resets, simCtrl, pbt...; outputs =
    observations..., rewards, dones, stats..., pbt...)."""

    def __init__(self, step_inputs: TrainStepInputInterface, step_outputs: TrainStepOutputInterface,
                 checkpointing: Optional[TrainCheckpointingInterface] = None):
        self._in, self._out, self._ckpt = step_inputs, step_outputs, checkpointing

    def stepInputs(self) -> TrainStepInputInterface:
        return self._in

    def stepOutputs(self) -> TrainStepOutputInterface:
        return self._out

    def checkpointing(self) -> Optional[TrainCheckpointingInterface]:
        return self._ckpt

    # -- ordered views (the order the reference's copy functions walk the buffers in)
    def _input_tensors(self) -> List[Tensor]:
        out = [nt.tensor for nt in self._in.actions] + [self._in.resets]
        if self._in.simCtrl is not None:
            out.append(self._in.simCtrl)
        return out + [nt.tensor for nt in self._in.pbt]

    def _observation_tensors(self) -> List[Tensor]:
        return [nt.tensor for nt in self._out.observations]

    def _output_tensors(self) -> List[Tensor]:
        return (self._observation_tensors() + [self._out.rewards, self._out.dones] +
                [nt.tensor for nt in self._out.stats] + [nt.tensor for nt in self._out.pbt])

    def copyStepInputs(self, buffers) -> None:
        """cpuCopyStepInputs / cudaCopyStepInputs: caller buffers -> the simulator's input tensors."""
        import torch
        for dst, src in zip(self._input_tensors(), buffers):
            dst.to_torch().copy_(torch.as_tensor(src).reshape(dst.dims()), non_blocking=True)

    def copyObservations(self, buffers) -> None:
        for src, dst in zip(self._observation_tensors(), buffers):
            dst.copy_(src.to_torch().reshape(dst.shape), non_blocking=True)

    def copyStepOutputs(self, buffers) -> None:
        """cpuCopyStepOutputs / cudaCopyStepOutputs: the simulator's outputs -> caller buffers."""
        for src, dst in zip(self._output_tensors(), buffers):
            dst.copy_(src.to_torch().reshape(dst.shape), non_blocking=True)

    # -- the pytree view the reference's Python side works with (bindings.cpp:30-78)
    def step_inputs(self) -> Dict[str, object]:
        d = {"actions": {nt.name: nt.tensor.to_torch() for nt in self._in.actions},
             "resets": self._in.resets.to_torch()}
        if self._in.simCtrl is not None:
            d["sim_ctrl"] = self._in.simCtrl.to_torch()
        if self._in.pbt:
            d["pbt"] = {nt.name: nt.tensor.to_torch() for nt in self._in.pbt}
        return d

    def step_outputs(self) -> Dict[str, object]:
        d = {"obs": {nt.name: nt.tensor.to_torch() for nt in self._out.observations},
             "rewards": self._out.rewards.to_torch(), "dones": self._out.dones.to_torch()}
        if self._out.stats:
            d["stats"] = {nt.name: nt.tensor.to_torch() for nt in self._out.stats}
        if self._out.pbt:
            d["pbt"] = {nt.name: nt.tensor.to_torch() for nt in self._out.pbt}
        return d
