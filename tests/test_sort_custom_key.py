"""SortArchetypeNode<A, C> with a custom key: stable LSD sort on all 32 key bits
(4 passes), rows are NOT regrouped by world, nothing is truncated
(src/mw/device/sort_archetype.cpp:1431-1440).  Oracle = the numpy
restatement (oracle/restate.py); the reference CPU backend cannot serve here
because its sortArchetype applies the inverse permutation."""
import numpy as np
import pytest

from oracle import restate
from sims import SIMS, make_executor


def _expected_keys(seed, items, t, mask):
    return np.array([restate.bits32(restate.split_i(restate.init_key(seed), i, t)) & mask
                     for i in range(items)], dtype=np.uint32)


def test_restatement_custom_key_sort_is_stable():
    keys = np.array([5, 1, 5, 0, 1, 0xFFFFFFFF, 3], dtype=np.uint32)
    perm, n, off, cnt = restate.sort_archetype(keys, num_worlds=4, world_sort=False)
    assert n == 7 and off is None
    assert perm.tolist() == [3, 1, 4, 6, 0, 2, 5]


@pytest.mark.gpu
@pytest.mark.parametrize("W,items,mask", [(7, 9, 0xFFFFFFFF), (300, 11, 0xFF), (1000, 5, 0x3)])
def test_gpu_custom_key_sort_matches_restatement(W, items, mask):
    ex = make_executor("sortcheck", W, items_per_world=items, key_mask=mask, seed=40)
    graph = ex.buildLaunchGraphAllTaskGraphs()
    n = W * items
    # arrangement after construction: world-major, creation order
    world = np.repeat(np.arange(W, dtype=np.uint32), items)
    item = np.tile(np.arange(items, dtype=np.uint32), W)
    for t in range(1, 4):
        ex.run(graph)
        assert ex.exportedNumRows(0) == n
        key = ex.tensor(0, "uint32", (n, 1)).cpu().numpy()[:, 0]
        payload = ex.tensor(1, "uint32", (n, 4)).cpu().numpy()
        tag = ex.tensor(2, "uint8", (n, 6)).cpu().numpy()
        # keys every row got this step, in the PREVIOUS arrangement
        new_keys = np.empty(n, dtype=np.uint32)
        for w in range(W):
            k = _expected_keys(40 + w, items, t, mask)
            sel = world == w
            new_keys[sel] = k[item[sel]]
        perm, _, _, _ = restate.sort_archetype(new_keys, W, world_sort=False)
        world, item = world[perm], item[perm]
        assert np.array_equal(key, new_keys[perm])
        assert np.array_equal(payload[:, 0], world) and np.array_equal(payload[:, 1], item)
        want_tag = ((item[:, None] * 7 + np.arange(6)[None, :] + 40 + world[:, None]) & 0xFF).astype(np.uint8)
        assert np.array_equal(tag, want_tag)
    ex.close()
