"""Batch ray caster with real acceleration structures:
reference-format BLAS (quantised 4-wide MeshBVH over de-indexed triangles, built
by mb2_build_mesh_bvhs), per-world TLAS rebuilt every step on the GPU (Morton
sort + LBVH + 4-wide collapse), materials and lights with shadow rays.

PARITY UNPINNED against the reference (its ray caster exists only inside its GPU
backend).  The pin used instead is independent of any acceleration structure:
every pixel's (instance, triangle, depth) must equal the closest hit of a
BRUTE-FORCE float64 scan over all triangles of the world, and the colour must
follow bvh_raycast.cpp:756-938 evaluated on that brute-force hit."""
import os

import numpy as np
import pytest

from sims.render_assets import GALLERY_MATERIALS, gallery_meshes


# ---- CPU: the BLAS builder ---------------------------------------------------------------------

def _decode_nodes(raw):
    n = len(raw)
    out = dict(
        min_point=raw[:, 0:12].copy().view(np.float32).reshape(n, 3),
        exp=raw[:, 12:15].copy().view(np.int8).reshape(n, 3),
        num_children=raw[:, 15],
        tri_size=raw[:, 16:20],
        qmin=np.stack([raw[:, 20:24], raw[:, 24:28], raw[:, 28:32]], axis=2),     # [n, child, axis]
        qmax=np.stack([raw[:, 32:36], raw[:, 36:40], raw[:, 40:44]], axis=2),
        children=raw[:, 44:60].copy().view(np.uint32).reshape(n, 4),
    )
    return out


def test_blas_builder_produces_valid_reference_format_trees():
    import madrona_b200 as mb
    meshes = gallery_meshes()
    bvh = mb.MeshBVHData(meshes, gpu_id=-1)
    raw, verts = bvh.host_arrays()
    view = bvh.view(device=False)
    assert view.num_bvhs == len(meshes) and view.num_verts == 3 * sum(len(m[1]) for m in meshes)
    nodes = _decode_nodes(raw)
    src = bvh.triangle_sources()
    node_base, tri_base = 0, 0
    for pos, tris, _mat in meshes:
        nt = len(tris)
        # de-indexed vertices of this mesh reproduce the source triangles (permuted)
        mine = verts[tri_base * 3:(tri_base + nt) * 3, :3].reshape(nt, 3, 3)
        perm = src[tri_base:tri_base + nt]
        assert sorted(perm.tolist()) == list(range(nt))
        assert np.array_equal(mine, pos[tris[perm]])
        # walk the tree: every triangle in exactly one leaf, inside every box on its path
        seen = np.zeros(nt, dtype=np.int32)
        stack = [(0, np.full(3, -np.inf), np.full(3, np.inf))]
        count = 0
        while stack:
            ni, lo_p, hi_p = stack.pop()
            g = node_base + ni
            count += 1
            scale = np.ldexp(1.0, nodes["exp"][g].astype(np.int32))
            for c in range(4):
                child = int(nodes["children"][g, c])
                if child == 0xFFFFFFFF:
                    continue
                lo = nodes["min_point"][g] + scale * nodes["qmin"][g, c]
                hi = nodes["min_point"][g] + scale * nodes["qmax"][g, c]
                if child & 0x80000000:
                    first = child & 0x7FFFFFFF
                    k = int(nodes["tri_size"][g, c])
                    assert 1 <= k <= 2
                    for t in range(first, first + k):
                        seen[t] += 1
                        p = mine[t]
                        assert (p >= lo - 1e-6).all() and (p <= hi + 1e-6).all()
                        assert (p >= lo_p - 1e-6).all() and (p <= hi_p + 1e-6).all()
                else:
                    stack.append((child, np.maximum(lo, lo_p), np.minimum(hi, hi_p)))
        assert (seen == 1).all()
        node_base += count
        tri_base += nt
    assert node_base == view.num_nodes


# ---- brute-force oracle ------------------------------------------------------------------------

def _quat_rotate(q, v):
    w, x, y, z = q
    u = np.array([x, y, z], dtype=np.float64)
    return v + 2.0 * np.cross(u, np.cross(u, v) + w * v)


def _rot_matrix(q):
    return np.stack([_quat_rotate(q, e) for e in np.eye(3)], axis=1)


def _world_triangles(meshes, pos, rot, scale, obj):
    """All triangles of a world: [T, 3, 3] float64, owner instance [T], source triangle [T]."""
    tris, inst, src = [], [], []
    for i in range(len(pos)):
        v, f, _ = meshes[int(obj[i])]
        M = _rot_matrix(rot[i].astype(np.float64)) * scale[i].astype(np.float64)[None, :]
        wv = v.astype(np.float64) @ M.T + pos[i].astype(np.float64)
        tris.append(wv[f])
        inst.append(np.full(len(f), i))
        src.append(np.arange(len(f)))
    return np.concatenate(tris), np.concatenate(inst), np.concatenate(src)


def _closest_hits(o, d, tris, t_min=0.0, chunk=256):
    """Moeller-Trumbore in float64, rays [R,3] x tris [T,3,3] -> (t [R], tri [R], second-best t [R])."""
    R = len(d)
    best = np.full(R, np.inf)
    second = np.full(R, np.inf)
    best_tri = np.full(R, -1)
    e1 = tris[:, 1] - tris[:, 0]
    e2 = tris[:, 2] - tris[:, 0]
    for s in range(0, len(tris), chunk):
        a, b, c = tris[s:s + chunk, 0], e1[s:s + chunk], e2[s:s + chunk]
        p = np.cross(d[:, None, :], c[None, :, :])
        det = (b[None] * p).sum(-1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            tv = o[:, None, :] - a[None]
            u = (tv * p).sum(-1) * inv
            q = np.cross(tv, b[None])
            v = (d[:, None, :] * q).sum(-1) * inv
            t = (c[None] * q).sum(-1) * inv
        ok = (np.abs(det) > 1e-14) & (u >= 0) & (v >= 0) & (u + v <= 1) & (t > t_min)
        t = np.where(ok, t, np.inf)
        order = np.sort(t, axis=1)
        cand = t.argmin(axis=1)
        cand_t = t[np.arange(R), cand]
        second_here = order[:, 1] if t.shape[1] > 1 else np.full(R, np.inf)
        new_best = cand_t < best
        second = np.where(new_best, np.minimum(best, second_here), np.minimum(second, cand_t))
        best_tri = np.where(new_best, s + cand, best_tri)
        best = np.where(new_best, cand_t, best)
    return best, best_tri, second


def _camera_rays(pos, rot_inv, fov_scale, res):
    """bvh_raycast.cpp:58-88 in float64."""
    q_inv = rot_inv.astype(np.float64)
    q = np.array([q_inv[0], -q_inv[1], -q_inv[2], -q_inv[3]])
    fwd = _quat_rotate(q, np.array([0.0, 1.0, 0.0]))
    fwd /= np.linalg.norm(fwd)
    u = _quat_rotate(q, np.array([1.0, 0.0, 0.0]))
    h = 1.0 / fov_scale
    vv = np.cross(fwd, u)
    vv /= np.linalg.norm(vv)
    horizontal, vertical = u * 2 * h, vv * 2 * h
    ll = pos - horizontal / 2 - vertical / 2 + fwd
    px = (np.arange(res) + 0.5) / res
    d = ll[None, None] + px[None, :, None] * horizontal + px[:, None, None] * vertical - pos
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return d.reshape(-1, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("P", [100, 40], ids=["tlas_100_instances", "flat_list_40_instances"])
def test_gpu_hits_match_brute_force_closest_hit(monkeypatch, P):
    # 100 props/world: per-world TLAS traversal; 40: the shared-memory staged instance list
    # (worlds with <= 64 instances) -- same BLAS traversal, same oracle
    import torch
    from sims import make_executor

    monkeypatch.setenv("MADRONA_B200_RENDER_DEBUG", "1")
    W, res, steps = 3, 40, 3
    ex = make_executor("gallery", W, num_props=P, seed=5, resolution=res, rgbd=True)
    step, render = ex.buildLaunchGraphAllTaskGraphs(), ex.buildRenderGraph()
    for _ in range(steps):
        ex.run(step)
    ex.run(render)
    assert ex.exportedNumRows(0) == W * P and ex.exportedNumRows(8) == 2 * W

    def col(slot, dtype, shape):
        return ex.tensor(slot, dtype, shape).cpu().numpy()
    pos, rot = col(0, "float32", (W, P, 3)), col(1, "float32", (W, P, 4))
    scale, obj = col(2, "float32", (W, P, 3)), col(3, "int32", (W, P))
    mat, color = col(4, "int32", (W, P)), col(5, "uint32", (W, P))
    vpos, vrot = col(6, "float32", (W, 2, 3)), col(7, "float32", (W, 2, 4))
    rgb = col(8, "uint8", (2 * W, res, res, 4))
    depth = col(9, "float32", (2 * W, res, res))
    hits = ex.renderDebugHits(2 * W, res).cpu().numpy()
    ex.close()

    import madrona_b200 as mb
    meshes = gallery_meshes()
    bvh = mb.MeshBVHData(meshes, gpu_id=-1)
    src_of = bvh.triangle_sources()
    first_tri = np.cumsum([0] + [len(m[1]) for m in meshes])
    fov_scale = 1.0 / np.tan(np.radians(70.0 * 0.5))

    total, id_checked, rgb_checked = 0, 0, 0
    for w in range(W):
        visible = np.array([i == 0 or i % 17 != 0 for i in range(P)])
        idx = np.nonzero(visible)[0]              # instance k of the engine = k-th visible prop
        tris, owner, src = _world_triangles(meshes, pos[w, idx], rot[w, idx], scale[w, idx], obj[w, idx])
        for v in range(2):
            view = 2 * w + v
            cam = vpos[w, v].astype(np.float64) + np.array([0.0, 0.0, 0.25])
            q = vrot[w, v].astype(np.float64)
            rays = _camera_rays(cam, np.array([q[0], -q[1], -q[2], -q[3]]), fov_scale, res)
            o = np.broadcast_to(cam, rays.shape)
            t, tri, t2 = _closest_hits(o, rays, tris)
            want_hit = np.isfinite(t)
            got_d = depth[view].reshape(-1)
            got_hit = got_d > 0
            total += len(t)
            # silhouettes may differ between the watertight fp32 test and float64 Moeller-Trumbore
            assert (want_hit == got_hit).mean() > 0.995
            both = want_hit & got_hit
            np.testing.assert_allclose(got_d[both], t[both], rtol=1e-4, atol=1e-4)
            # ids: exact wherever the closest hit is not a near tie
            with np.errstate(invalid="ignore"):
                clear = both & ((t2 - t) > 1e-3 * np.maximum(t, 1.0))
            g_inst, g_tri = hits[view].reshape(-1, 2)[:, 0], hits[view].reshape(-1, 2)[:, 1]
            want_inst = owner[np.maximum(tri, 0)]
            assert np.array_equal(g_inst[clear], want_inst[clear])
            mesh_of = obj[w, idx][want_inst]
            got_src = src_of[first_tri[mesh_of] + np.maximum(g_tri, 0)]
            assert np.array_equal(got_src[clear], src[np.maximum(tri, 0)][clear])
            id_checked += int(clear.sum())

            # ---- colour: bvh_raycast.cpp:756-938 on the brute-force hit
            sel = np.nonzero(clear)[0]
            hit_tri = tris[tri[sel]]
            n = np.cross(hit_tri[:, 1] - hit_tri[:, 0], hit_tri[:, 2] - hit_tri[:, 0])
            # the engine's normal: object-space geometric normal rotated by the instance
            # rotation (scale ignored, as the reference does): recompute it that way
            inst_k = want_inst[sel]
            n_obj = []
            for k, s_tri in zip(inst_k, src[tri[sel]]):
                vtx, f, _ = meshes[int(obj[w, idx][k])]
                a, b, c = vtx[f[s_tri]].astype(np.float64)
                nn = np.cross(b - a, c - a)
                n_obj.append(_quat_rotate(rot[w, idx][k].astype(np.float64), nn / np.linalg.norm(nn)))
            n = np.array(n_obj)
            hit_pos = o[sel] + t[sel, None] * rays[sel]
            contrib = np.zeros(len(sel))
            # light 0: directional, shadow rays from 1 mm above the surface (see the kernel)
            ldir = -np.array([0.3, 0.2, -0.9327379])
            facing = (n @ ldir) > 0
            sh_o = hit_pos + 1e-3 * n
            st, _, _ = _closest_hits(sh_o, np.broadcast_to(ldir, hit_pos.shape), tris, t_min=1e-6)
            lit = facing & ~np.isfinite(st)
            # pixels whose shadow ray only grazes an occluder (or starts inside one) are skipped
            st_a, _, _ = _closest_hits(hit_pos + 3e-3 * n, np.broadcast_to(ldir, hit_pos.shape), tris, t_min=1e-6)
            st_b, _, _ = _closest_hits(hit_pos + 3e-4 * n, np.broadcast_to(ldir, hit_pos.shape), tris, t_min=1e-6)
            graze = (np.isfinite(st) != np.isfinite(st_a)) | (np.isfinite(st) != np.isfinite(st_b))
            contrib += np.where(lit, np.clip(n @ ldir, 0, 1), 0.0)
            # light 1: spotlight at (0, 0, 9) pointing down, cutoff 0.9 rad, no shadows
            to_l = np.array([0.0, 0.0, 9.0]) - hit_pos
            to_l /= np.linalg.norm(to_l, axis=1, keepdims=True)
            ang = np.arccos(np.clip((-to_l) @ np.array([0.0, 0.0, -1.0]), -1, 1))
            inside = np.abs(ang) <= 0.9
            contrib += np.where(inside, np.clip((n * to_l).sum(1), 0, 1), 0.0)
            base = np.ones((len(sel), 3))
            for j, k in enumerate(inst_k):
                m = int(mat[w, idx][k])
                if m == -2:
                    hx = int(color[w, idx][k])
                    base[j] = [((hx >> 16) & 255) / 255.0, ((hx >> 8) & 255) / 255.0, (hx & 255) / 255.0]
                else:
                    if m == -1:
                        m = meshes[int(obj[w, idx][k])][2]
                    if m >= 0:
                        base[j] = GALLERY_MATERIALS[m, :3]
            want_rgb = np.clip(np.maximum(0.2, contrib)[:, None] * base, 0, 1) * 255.0
            got_rgb = rgb[view].reshape(-1, 4)[sel, :3].astype(np.float64)
            edge = graze | (np.abs(np.abs(ang) - 0.9) < 5e-3) | (np.abs(n @ ldir) < 5e-3)
            ok = np.abs(got_rgb - want_rgb).max(axis=1) <= 2.0
            assert ok[~edge].mean() > 0.99, (w, v, float(ok[~edge].mean()), float(edge.mean()))
            assert (rgb[view][..., 3] == 255).all()
            rgb_checked += int((~edge).sum())
    assert id_checked > 0.8 * total * 0.5 and rgb_checked > 1000
