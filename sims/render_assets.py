"""Render meshes of the rigid-body fixtures for the batch ray caster: one flat
triangle range per object ID (cube, wall, agent = unit box; plane = big quad).
The reference bakes MeshBVHs with embree (src/common/mesh_bvh_builder.cpp, not
available: SURVEY.md 8f N4); fixtures need a dozen triangles."""
from __future__ import annotations

import ctypes

import numpy as np


def unit_box():
    v = np.array([[x, y, z] for z in (-0.5, 0.5) for y in (-0.5, 0.5) for x in (-0.5, 0.5)],
                 dtype=np.float32)
    quads = [[0, 2, 3, 1], [4, 5, 7, 6], [0, 1, 5, 4], [2, 6, 7, 3], [0, 4, 6, 2], [1, 3, 7, 5]]
    tris = []
    for q in quads:
        tris += [[q[0], q[1], q[2]], [q[0], q[2], q[3]]]
    return v, np.array(tris, dtype=np.uint32)


def room_meshes():
    """Returns (mesh_descs [n,8] float/uint view, vertices [nv,3] f32, indices [nt,3] u32)."""
    bv, bt = unit_box()
    e = 500.0
    pv = np.array([[-e, -e, 0], [e, -e, 0], [e, e, 0], [-e, e, 0]], dtype=np.float32)
    pt = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    verts = np.concatenate([bv, pv])
    plane_tris = pt + len(bv)
    indices = np.concatenate([bt, plane_tris])
    # objects: 0 cube, 1 wall, 2 agent share the box triangles; 3 = plane
    descs = np.zeros(4, dtype=[("first", "<u4"), ("count", "<u4"), ("mn", "<f4", 3), ("mx", "<f4", 3)])
    for o in range(3):
        descs[o] = (0, len(bt), (-0.5, -0.5, -0.5), (0.5, 0.5, 0.5))
    descs[3] = (len(bt), len(pt), (-e, -e, 0), (e, e, 0))
    return descs, verts, indices


def mesh_list(descs, verts, indices):
    """[(positions, triangles, material)] per object from the flat description."""
    out = []
    for d in descs:
        tris = indices[d["first"]:d["first"] + d["count"]]
        used = np.unique(tris)
        remap = {int(v): i for i, v in enumerate(used)}
        out.append((verts[used], np.vectorize(remap.get)(tris).astype(np.uint32), -1))
    return out


def make_render_config_from_meshes(meshes, resolution: int, rgbd: bool, gpu_id: int = 0, materials=None):
    """mb2_render_config (== CudaBatchRenderConfig) whose geoBVHData was built by the
    engine's BLAS builder; returns (config, keep_alive)."""
    import madrona_b200 as mb
    from madrona_b200.executor import _MaterialViewC, _RenderConfigC

    bvh = mb.MeshBVHData(meshes, gpu_id=gpu_id)
    mat_view = _MaterialViewC(None, 0, None, None)
    keep = [bvh]
    if materials is not None:
        import torch
        m = torch.from_numpy(np.ascontiguousarray(materials, dtype=np.float32)).to(f"cuda:{gpu_id}")
        keep.append(m)
        mat_view = _MaterialViewC(None, 0, None, m.data_ptr())
    rc = _RenderConfigC(0 if rgbd else 1, bvh.view(device=True), mat_view, resolution, 0.001, 1000.0)
    return rc, keep


def make_render_config(resolution: int, rgbd: bool = False, gpu_id: int = 0):
    descs, verts, indices = room_meshes()
    return make_render_config_from_meshes(mesh_list(descs, verts, indices), resolution, rgbd, gpu_id)


# ---- gallery fixture: four prop meshes + ground ------------------------------------------------

def icosphere(subdiv: int = 2):
    t = (1.0 + 5 ** 0.5) / 2.0
    v = [[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
         [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]]
    f = [[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
         [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
         [6, 2, 10], [8, 6, 7], [9, 8, 1]]
    v = [np.array(p, dtype=np.float64) / np.linalg.norm(p) for p in v]
    for _ in range(subdiv):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = (v[a] + v[b]) * 0.5
                v.append(m / np.linalg.norm(m))
                cache[key] = len(v) - 1
            return cache[key]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f = nf
    return (np.array(v) * 0.5).astype(np.float32), np.array(f, dtype=np.uint32)


def prism(n: int):
    """n-gon prism with fans on the caps: 4n triangles."""
    ang = 2 * np.pi * np.arange(n) / n
    ring = np.stack([0.5 * np.cos(ang), 0.5 * np.sin(ang)], axis=1)
    v = [[x, y, -0.5] for x, y in ring] + [[x, y, 0.5] for x, y in ring] + [[0, 0, -0.5], [0, 0, 0.5]]
    f = []
    for i in range(n):
        j = (i + 1) % n
        f += [[i, j, j + n], [i, j + n, i + n], [2 * n, j, i], [2 * n + 1, i + n, j + n]]
    return np.array(v, dtype=np.float32), np.array(f, dtype=np.uint32)


def gallery_meshes():
    """[(positions, triangles, default material)]: box, icosphere (320 tris), 12-gon prism (48),
    tetrahedron (4), ground quad (2)."""
    bv, bt = unit_box()
    sv, st = icosphere(2)
    pv, pt = prism(12)
    tv = np.array([[0.5, 0.5, 0.5], [-0.5, -0.5, 0.5], [-0.5, 0.5, -0.5], [0.5, -0.5, -0.5]], dtype=np.float32)
    tt = np.array([[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]], dtype=np.uint32)
    e = 60.0
    gv = np.array([[-e, -e, 0], [e, -e, 0], [e, e, 0], [-e, e, 0]], dtype=np.float32)
    gt = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.uint32)
    return [(bv, bt, 1), (sv, st, 2), (pv, pt, -1), (tv, tt, 3), (gv, gt, 0)]


GALLERY_MATERIALS = np.array([
    # color rgba, textureIdx (int bits), roughness, metalness  == madrona::Material (28 B)
    [0.55, 0.55, 0.5, 1.0, 0, 0.8, 0.0],
    [0.9, 0.3, 0.2, 1.0, 0, 0.5, 0.0],
    [0.2, 0.6, 0.9, 1.0, 0, 0.5, 0.0],
    [0.3, 0.8, 0.3, 1.0, 0, 0.5, 0.0],
], dtype=np.float32)
GALLERY_MATERIALS[:, 4] = np.array([-1], dtype=np.int32).view(np.float32)[0]     # textureIdx = -1


def make_gallery_render_config(resolution: int, rgbd: bool, gpu_id: int = 0):
    return make_render_config_from_meshes(gallery_meshes(), resolution, rgbd, gpu_id, materials=GALLERY_MATERIALS)
