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
