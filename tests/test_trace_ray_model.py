"""BVH::traceRay (madrona_b200/device/madrona/physics.hpp): the shipped mechanism -- per chunk
of 64 boxes a uniform pass that keeps the boxes entered with the CURRENT t_max as a mask, then
an ordered pass over the mask bits that repeats the slab test with the t_max of that moment --
must enter exactly the leaves, in exactly the order, of the one-phase ordered scan it replaced
(which in turn equals the reference's stack walk, src/physics/broadphase.cpp:658-724).

This is a CPU model of the two control flows (float32, NaN-skipping fmin / fmax like the
device's fminf / fmaxf, math.inl:1670-1696) over random worlds, including > 64 boxes (several
chunks, which no GPU fixture reaches), unbounded plane boxes, axis-parallel rays (infinite
1/d) and ties.  The GPU parity tests pin the real kernels; this pins the equivalence argument."""
import numpy as np

F = np.float32


def _slab(o, inv_d, box, t_max):
    with np.errstate(invalid="ignore", over="ignore"):
        lo = inv_d * (box[:3] - o)
        hi = inv_d * (box[3:] - o)
    entry = np.fmax(np.fmin(lo[0], hi[0]), np.fmax(np.fmin(lo[1], hi[1]), np.fmax(np.fmin(lo[2], hi[2]), F(0))))
    exit_ = np.fmin(np.fmax(lo[0], hi[0]), np.fmin(np.fmax(lo[1], hi[1]), np.fmin(np.fmax(lo[2], hi[2]), F(t_max))))
    return bool(entry <= exit_)


def _leaf_test(hit_t, t_max):
    """traceRayIntoLeaf: a hit only inside [0, t_max]."""
    return hit_t is not None and hit_t <= t_max


def one_phase_scan(o, inv_d, boxes, hits, t_max):
    entered, closest = [], -1
    for j in range(len(boxes)):
        if _slab(o, inv_d, boxes[j], t_max):
            entered.append(j)
            if _leaf_test(hits[j], t_max):
                t_max, closest = hits[j], j
    return entered, closest, t_max


def mask_then_ordered_retest(o, inv_d, boxes, hits, t_max):
    entered, closest = [], -1
    for base in range(0, len(boxes), 64):
        chunk = range(base, min(base + 64, len(boxes)))
        cand = [j for j in chunk if _slab(o, inv_d, boxes[j], t_max)]       # uniform pass
        for j in cand:                                                       # ordered pass
            if _slab(o, inv_d, boxes[j], t_max):
                entered.append(j)
                if _leaf_test(hits[j], t_max):
                    t_max, closest = hits[j], j
    return entered, closest, t_max


def _world(rng, n):
    c = rng.uniform(-20, 20, size=(n, 3)).astype(F)
    h = rng.uniform(0.2, 6, size=(n, 3)).astype(F)
    boxes = np.concatenate([c - h, c + h], axis=1)
    # a ground-plane style box: unbounded in x / y
    boxes[rng.integers(0, n)] = np.array([-np.inf, -np.inf, -1, np.inf, np.inf, 0], dtype=F)
    return boxes


def test_mask_and_retest_enters_the_same_leaves_in_the_same_order():
    rng = np.random.default_rng(7)
    total_entered = 0
    for case in range(600):
        n = int(rng.choice([5, 33, 49, 64, 65, 130, 200]))
        boxes = _world(rng, n)
        o = rng.uniform(-15, 15, size=3).astype(F)
        d = rng.normal(size=3).astype(F)
        if case % 4 == 0:
            d[rng.integers(0, 3)] = F(0)          # axis-parallel: 1/d = inf, 0 * inf = NaN in the slabs
        d /= np.linalg.norm(d)
        with np.errstate(divide="ignore"):
            inv_d = (F(1) / d).astype(F)
        # a leaf's hit lies behind the entry of its box (or it has none); some exact ties
        hits = []
        for j in range(n):
            if rng.random() < 0.5:
                hits.append(None)
            else:
                hits.append(F(rng.choice([3.0, 7.5, 12.0])) if rng.random() < 0.2 else F(rng.uniform(0.5, 60)))
        a = one_phase_scan(o, inv_d, boxes, hits, F(200))
        b = mask_then_ordered_retest(o, inv_d, boxes, hits, F(200))
        assert a == b, (case, n)
        total_entered += len(a[0])
    assert total_entered > 600       # the walks are not vacuous
