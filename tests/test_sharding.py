"""Multi-GPU path on CPU: world_size-2 gloo processes exercise the shard /
gather / slice logic bench.py uses with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from madrona_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world_size, port, total_worlds, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    first, count = sharding.shard_range(total_worlds, world_size, rank)
    # "exported reward column" of this rank: value encodes the global world index
    local = (torch.arange(count, dtype=torch.float32) + first).reshape(count, 1).repeat(1, 2)
    gathered = sharding.gather_exported(local)
    # global actions -> local slice
    actions = torch.arange(total_worlds * 3, dtype=torch.int32).reshape(total_worlds, 3)
    mine = sharding.local_slice(actions)
    seeds = [sharding.world_seed(100, first, w) for w in range(count)]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), gathered=gathered.numpy(), mine=mine.numpy(),
             seeds=np.array(seeds), first=first)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_gather_slice_world_size_2(tmp_path):
    total = 12
    mp.spawn(_worker, args=(2, _free_port(), total, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    want = np.repeat(np.arange(total, dtype=np.float32)[:, None], 2, axis=1)
    assert np.array_equal(r0["gathered"], want) and np.array_equal(r1["gathered"], want)
    acts = np.arange(total * 3, dtype=np.int32).reshape(total, 3)
    assert np.array_equal(r0["mine"], acts[:6]) and np.array_equal(r1["mine"], acts[6:])
    # the union of per-rank seeds == the seeds of one 12-world executor
    assert np.concatenate([r0["seeds"], r1["seeds"]]).tolist() == list(range(100, 112))


def test_shard_range_rejects_uneven_split():
    with pytest.raises(ValueError):
        sharding.shard_range(10, 4, 0)
    assert sharding.shard_range(8192 * 8, 8, 3) == (3 * 8192, 8192)
