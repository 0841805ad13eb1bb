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


def make_render_config(resolution: int, rgbd: bool = False):
    """ctypes mb2_render_config + the arrays it points into (keep them alive)."""
    from madrona_b200.executor import _RenderConfigC

    descs, verts, indices = room_meshes()
    descs_b = np.ascontiguousarray(descs)
    verts = np.ascontiguousarray(verts)
    indices = np.ascontiguousarray(indices)
    rc = _RenderConfigC(
        0 if rgbd else 1, resolution, 0.001, 1000.0,
        descs_b.ctypes.data_as(ctypes.c_void_p), len(descs_b),
        verts.ctypes.data_as(ctypes.c_void_p), len(verts),
        indices.ctypes.data_as(ctypes.c_void_p), len(indices))
    return rc, (descs_b, verts, indices)
