"""madrona_b200.Tensor mirrors madrona::py::Tensor (include/madrona/py/utils.hpp:58-141):
element types, accessors, zero-copy torch views.  CPU part here; the CUDA view is
covered by the gpu test below."""
import ctypes

import numpy as np
import pytest
import torch

import madrona_b200 as mb
from madrona_b200 import Tensor, TensorElementType


def test_element_types_match_reference_order():
    names = [e.name for e in TensorElementType]
    assert names == ["UInt8", "Int8", "Int16", "Int32", "Int64", "Float16", "Float32"]
    sizes = [Tensor(0, e, [1]).numBytesPerItem() for e in TensorElementType]
    assert sizes == [1, 1, 2, 4, 8, 2, 4]


def test_host_tensor_aliases_memory():
    backing = (ctypes.c_float * 24)()
    t = Tensor(ctypes.addressof(backing), TensorElementType.Float32, [2, 3, 4])
    assert not t.isOnGPU() and t.gpuID() == -1
    assert t.numDims() == 3 and tuple(t.dims()) == (2, 3, 4) and t.numBytes() == 96
    view = t.to_torch()
    assert view.dtype == torch.float32 and tuple(view.shape) == (2, 3, 4)
    view[1, 2, 3] = 7.5
    assert backing[23] == 7.5                 # zero copy
    backing[0] = -1.0
    assert t.to_numpy()[0, 0, 0] == -1.0


def test_round_trip_through_torch():
    src = torch.arange(12, dtype=torch.int32).reshape(3, 4)
    t = Tensor.from_torch(src)
    assert t.type() == TensorElementType.Int32 and tuple(t.dims()) == (3, 4)
    assert t.devicePtr() == src.data_ptr()
    assert torch.equal(t.to_torch(), src)


def test_too_many_dimensions_is_rejected():
    with pytest.raises(ValueError):
        Tensor(0, TensorElementType.UInt8, [1] * 17)


@pytest.mark.gpu
def test_exported_column_as_tensor():
    from sims import make_executor
    ex = make_executor("cartpole", 8, max_steps=10, seed=1)
    t = ex.exportedTensor(0, TensorElementType.Int32, [8, 1])
    assert t.isOnGPU() and t.gpuID() == 0
    view = t.to_torch()
    assert view.device.type == "cuda" and tuple(view.shape) == (8, 1)
    assert view.data_ptr() == ex.getExported(0)
    ex.close()


def test_train_interface_mirrors_reference_structure_and_buffer_order():
    from madrona_b200 import (NamedTensor, TrainInterface, TrainStepInputInterface, TrainStepOutputInterface)
    act = torch.zeros(4, 2, dtype=torch.int32)
    resets = torch.zeros(4, 1, dtype=torch.int32)
    obs_a = torch.arange(8, dtype=torch.float32).reshape(4, 2)
    obs_b = torch.ones(4, 3)
    rew, done = torch.full((4, 1), 0.5), torch.zeros(4, 1, dtype=torch.int32)
    iface = TrainInterface(
        TrainStepInputInterface(actions=[NamedTensor("move", Tensor.from_torch(act))], resets=Tensor.from_torch(resets)),
        TrainStepOutputInterface(observations=[NamedTensor("self", Tensor.from_torch(obs_a)),
                                               NamedTensor("lidar", Tensor.from_torch(obs_b))],
                                 rewards=Tensor.from_torch(rew), dones=Tensor.from_torch(done)))
    assert [nt.name for nt in iface.stepInputs().actions] == ["move"]
    assert iface.checkpointing() is None
    # inputs: actions..., resets (utils.hpp:149-154 order); zero copy into the sim's tensors
    iface.copyStepInputs([torch.full((4, 2), 3, dtype=torch.int32), torch.ones(4, 1, dtype=torch.int32)])
    assert (act == 3).all() and (resets == 1).all()
    # outputs: observations..., rewards, dones (utils.hpp:156-162 order)
    bufs = [torch.empty(4, 2), torch.empty(4, 3), torch.empty(4, 1), torch.empty(4, 1, dtype=torch.int32)]
    iface.copyStepOutputs(bufs)
    assert torch.equal(bufs[0], obs_a) and torch.equal(bufs[1], obs_b) and (bufs[2] == 0.5).all()
    tree = iface.step_outputs()
    assert set(tree) == {"obs", "rewards", "dones"} and set(tree["obs"]) == {"self", "lidar"}
    assert tree["obs"]["self"].data_ptr() == obs_a.data_ptr()
    assert iface.step_inputs()["actions"]["move"].data_ptr() == act.data_ptr()
