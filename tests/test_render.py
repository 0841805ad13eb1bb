"""Batch ray caster (BASELINE configs[3] class).
PARITY UNPINNED against the reference (it can only ray cast on its GPU backend);
the oracle is oracle/restate_render.py, a numpy restatement of the reference's
image-formation formulas, fed with the transforms the same GPU run exported.
Tolerance 1e-4 relative on depth (fmaf is emulated through float64 in numpy,
tanf differs by an ulp between libm and CUDA)."""
import numpy as np
import pytest

from oracle import restate_render as rr
from sims.render_assets import room_meshes, unit_box


def _aabb_trs(mn, mx, pos, rot, scale):
    # AABB::applyTRS restated (include/madrona/math.inl:1737-1769)
    w, x, y, z = (np.float32(v) for v in rot)
    s = np.asarray(scale, np.float32)
    x2, y2, z2 = x * x, y * y, z * z
    xz, xy, yz, wx, wy, wz = x * z, x * y, y * z, w * x, w * y, w * z
    ds = np.float32(2) * s
    cols = np.array([[s[0] - ds[0] * (y2 + z2), ds[0] * (xy + wz), ds[0] * (xz - wy)],
                     [ds[1] * (xy - wz), s[1] - ds[1] * (x2 + z2), ds[1] * (yz + wx)],
                     [ds[2] * (xz + wy), ds[2] * (yz - wx), s[2] - ds[2] * (x2 + y2)]], dtype=np.float32)
    lo = np.array(pos, np.float32).copy()
    hi = np.array(pos, np.float32).copy()
    for i in range(3):
        for j in range(3):
            e = cols[j][i] * np.float32(mn[j])
            f = cols[j][i] * np.float32(mx[j])
            if e < f:
                lo[i] += e
                hi[i] += f
            else:
                lo[i] += f
                hi[i] += e
    return lo, hi


def test_restatement_sees_a_box_where_expected():
    # camera at origin looking down +y at a unit cube 5 m away: centre depth 4.5
    descs, verts, indices = room_meshes()
    inst = dict(position=[0, 5, 0], rotation=[1, 0, 0, 0], scale=[1, 1, 1], object_id=0)
    inst["aabb_min"], inst["aabb_max"] = _aabb_trs(descs[0]["mn"], descs[0]["mx"], inst["position"],
                                                   inst["rotation"], inst["scale"])
    depth = rr.render_depth([inst], descs, verts, indices, [0, 0, 0], [1, 0, 0, 0], -1.0, 16)
    assert abs(depth[8, 8] - 4.5) < 0.05          # slightly off-axis pixel centre
    assert depth[0, 0] == 0.0 and depth[15, 15] == 0.0
    assert (depth > 0).sum() in range(4, 40)


def test_box_mesh_is_closed():
    v, t = unit_box()
    edges = {}
    for tri in t:
        for a, b in ((tri[0], tri[1]), (tri[1], tri[2]), (tri[2], tri[0])):
            edges[(a, b)] = edges.get((a, b), 0) + 1
    assert all(edges.get((b, a), 0) == 1 for (a, b) in edges)      # every edge has its twin


@pytest.mark.gpu
@pytest.mark.parametrize("rgbd", [False, True])
def test_gpu_depth_matches_restatement(rgbd):
    import torch
    from sims import make_executor
    from trace_utils import make_inputs

    W, res, steps = 6, 32, 25
    ex = make_executor("room_render", W, episode_len=20, seed=3, resolution=res, rgbd=rgbd)
    step = ex.buildLaunchGraphAllTaskGraphs()
    render = ex.buildRenderGraph()
    ins = make_inputs("room", W, steps, seed=8)
    act = ex.tensor(1, "int32", (W, 2, 3))
    descs, verts, indices = room_meshes()
    checked = 0
    for t in range(steps):
        act.copy_(torch.from_numpy(np.ascontiguousarray(ins["action"][t])))
        torch.cuda.synchronize()
        ex.run(step)
        ex.run(render)
        if t not in (0, 11, 24):      # 24: after an episode reset (new cubes, new walls)
            continue
        n_views = ex.exportedNumRows(14)
        assert n_views == 2 * W
        depth = ex.tensor(14, "float32", (n_views, res, res)).cpu().numpy()
        rgb = ex.tensor(13, "uint8", (n_views, res, res, 4)).cpu().numpy()
        agent_pos = ex.tensor(6, "float32", (W, 2, 3)).cpu().numpy()
        agent_rot = ex.tensor(7, "float32", (W, 2, 4)).cpu().numpy()
        n_body = ex.exportedNumRows(9)
        assert n_body == 31 * W
        body_pos = ex.tensor(9, "float32", (n_body, 3)).cpu().numpy().reshape(W, 31, 3)
        body_rot = ex.tensor(10, "float32", (n_body, 4)).cpu().numpy().reshape(W, 31, 4)
        body_scale = ex.tensor(15, "float32", (n_body, 3)).cpu().numpy().reshape(W, 31, 3)
        body_obj = ex.tensor(16, "int32", (n_body, 1)).cpu().numpy().reshape(W, 31)
        for w in (0, W - 1):
            # engine instance order: archetype id ascending (Agent was registered before
            # PhysicsEntity), rows in world order
            insts = [dict(position=agent_pos[w, a], rotation=agent_rot[w, a], scale=[1.0, 1.0, 1.5],
                          object_id=2) for a in range(2)]
            insts += [dict(position=body_pos[w, b], rotation=body_rot[w, b], scale=body_scale[w, b],
                           object_id=int(body_obj[w, b])) for b in range(31)]
            for i in insts:
                d = descs[i["object_id"]]
                i["aabb_min"], i["aabb_max"] = _aabb_trs(d["mn"], d["mx"], i["position"], i["rotation"],
                                                         i["scale"])
            for a in range(2):
                cam_pos = (agent_pos[w, a] + np.array([0, 0, 0.5], np.float32)).astype(np.float32)
                q = agent_rot[w, a]
                cam_rot_inv = np.array([q[0], -q[1], -q[2], -q[3]], np.float32)
                want = rr.render_depth(insts, descs, verts, indices, cam_pos, cam_rot_inv, -1.0, res)
                got = depth[2 * w + a]
                assert ((got > 0) == (want > 0)).mean() > 0.999
                both = (got > 0) & (want > 0)
                np.testing.assert_allclose(got[both], want[both], rtol=1e-4, atol=1e-5)
                checked += int(both.sum())
                if rgbd:
                    hit = got > 0
                    assert (rgb[2 * w + a][..., 3] == 255).all()
                    assert (rgb[2 * w + a][hit][:, 0] == 51).all()       # 0.2 * white * 255
                    assert (rgb[2 * w + a][~hit][:, :3] == 0).all()
    assert checked > 5000
    ex.close()
