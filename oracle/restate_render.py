"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the batch ray caster's image
formation (reference: src/mw/device/bvh_raycast.cpp :58-88 ray generation,
:620-645 / :744-751 object-space ray + t rescaling, :317-448 watertight
ray-triangle test, :820-838 output).  PARITY UNPINNED: the reference can only
ray cast on its GPU backend (src/render/ecs_system.cpp:684-689 disables it on
CPU), so nothing in /root/reference can produce a known-answer image here; this
restatement pins the engine's kernel to the published formulas instead.
float32 arithmetic throughout; fmaf is emulated through float64 (exact product,
one extra rounding in rare cases -> compare with a small tolerance)."""
from __future__ import annotations

import numpy as np

F = np.float32


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def quat_rotate(q, v):
    """q = (w,x,y,z) [4], v [...,3]: v + 2 (w (q x v) + q x (q x v))"""
    w = F(q[0])
    p = np.asarray(q[1:], dtype=F)
    qv = np.cross(p, v).astype(F)
    qqv = np.cross(p, qv).astype(F)
    return (v + F(2) * ((qv * w) + qqv)).astype(F)


def quat_inv(q):
    return np.array([q[0], -q[1], -q[2], -q[3]], dtype=F)


def normalize(v):
    l2 = (v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1] + v[..., 2] * v[..., 2]).astype(F)
    inv = (F(1) / np.sqrt(l2).astype(F)).astype(F)
    return (v * inv[..., None]).astype(F)


def primary_rays(cam_pos, cam_rot_inv, y_scale, res):
    """cam_rot_inv = inverse of the viewing entity's rotation (as stored in the view)."""
    rot = np.asarray(cam_rot_inv, dtype=F)
    start = np.asarray(cam_pos, dtype=F)
    look_at = quat_rotate(quat_inv(rot), np.array([0, 1, 0], dtype=F))
    h = F(1) / F(-y_scale)
    viewport = F(2) * h
    forward = normalize(look_at)
    u = quat_rotate(quat_inv(rot), np.array([1, 0, 0], dtype=F))
    v = normalize(np.cross(forward, u).astype(F))
    horizontal = (u * viewport).astype(F)
    vertical = (v * viewport).astype(F)
    half = F(1) / F(2)
    lower_left = (start - horizontal * half - vertical * half + forward).astype(F)
    px = (np.arange(res, dtype=F) + F(0.5)) / F(res)
    pu, pv = np.meshgrid(px, px)       # [py, px]
    d = (lower_left + pu[..., None] * horizontal + pv[..., None] * vertical - start).astype(F)
    return start, normalize(d)


def _slab(box_min, box_max, o, inv_d, t_max):
    t_min = np.zeros_like(t_max)
    t_hi = t_max.copy()
    ok = np.ones(t_max.shape, dtype=bool)
    for i in range(3):
        t0 = ((box_min[i] - o[..., i]) * inv_d[..., i]).astype(F)
        t1 = ((box_max[i] - o[..., i]) * inv_d[..., i]).astype(F)
        neg = inv_d[..., i] < 0
        t0, t1 = np.where(neg, t1, t0), np.where(neg, t0, t1)
        t_min = np.where(ok & (t0 > t_min), t0, t_min)
        t_hi = np.where(ok & (t1 < t_hi), t1, t_hi)
        ok = ok & ~(t_hi <= t_min)
    return ok


def _ray_triangle(a, b, c, kx, ky, kz, Sx, Sy, Sz, org, t_max):
    A, B, C = (a - org).astype(F), (b - org).astype(F), (c - org).astype(F)

    def comp(V, k):
        return np.take_along_axis(V, k[..., None], axis=-1)[..., 0]

    a_kz, a_kx, a_ky = comp(A, kz), comp(A, kx), comp(A, ky)
    b_kz, b_kx, b_ky = comp(B, kz), comp(B, kx), comp(B, ky)
    c_kz, c_kx, c_ky = comp(C, kz), comp(C, kx), comp(C, ky)
    Ax, Ay = _fma(-Sx, a_kz, a_kx), _fma(-Sy, a_kz, a_ky)
    Bx, By = _fma(-Sx, b_kz, b_kx), _fma(-Sy, b_kz, b_ky)
    Cx, Cy = _fma(-Sx, c_kz, c_kx), _fma(-Sy, c_kz, c_ky)
    U = _fma(Cx, By, -(Cy * Bx).astype(F))
    V = _fma(Ax, Cy, -(Ay * Cx).astype(F))
    W = _fma(Bx, Ay, -(By * Ax).astype(F))
    eps = F(1e-7)
    U = np.where((U > -eps) & (U < eps), F(0), U)
    V = np.where((V > -eps) & (V < eps), F(0), V)
    W = np.where((W > -eps) & (W < eps), F(0), W)
    mixed = ((U < 0) | (V < 0) | (W < 0)) & ((U > 0) | (V > 0) | (W > 0))
    edge = ~mixed & ((U == 0) | (V == 0) | (W == 0))
    Ud = (Cx.astype(np.float64) * By - Cy.astype(np.float64) * Bx).astype(F)
    Vd = (Ax.astype(np.float64) * Cy - Ay.astype(np.float64) * Cx).astype(F)
    Wd = (Bx.astype(np.float64) * Ay - By.astype(np.float64) * Ax).astype(F)
    U, V, W = np.where(edge, Ud, U), np.where(edge, Vd, V), np.where(edge, Wd, W)
    mixed2 = ((U < 0) | (V < 0) | (W < 0)) & ((U > 0) | (V > 0) | (W > 0))
    det = (U + V + W).astype(F)
    Az, Bz, Cz = (Sz * a_kz).astype(F), (Sz * b_kz).astype(F), (Sz * c_kz).astype(F)
    T = _fma(U, Az, _fma(V, Bz, (W * Cz).astype(F)))
    sign = det.view(np.uint32) & np.uint32(0x80000000)
    xor_T = (T.view(np.uint32) ^ sign).view(F)
    abs_det = np.abs(det)
    ok = ~mixed & ~mixed2 & (det != 0) & ~(xor_T < 0) & ~(xor_T > (t_max * abs_det).astype(F))
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (T * (F(1) / det)).astype(F)
    return ok, t


def render_depth(instances, meshes, verts, indices, cam_pos, cam_rot_inv, y_scale, res):
    """instances: list of dict(position[3], rotation[4] (w,x,y,z), scale[3], object_id,
    aabb_min[3], aabb_max[3]) in the engine's per-world order.  Returns depth [res,res]."""
    o_w, d_w = primary_rays(cam_pos, cam_rot_inv, y_scale, res)
    o_w = np.ascontiguousarray(np.broadcast_to(o_w, d_w.shape)).astype(F)
    with np.errstate(divide="ignore"):
        inv_d_w = (F(1) / d_w).astype(F)
    t_max = np.full((res, res), F(10000), dtype=F)
    hit = np.zeros((res, res), dtype=bool)

    # the engine visits a view's instances near-to-far: squared distance from the
    # eye to the instance's world box, ties in engine order (kernels_render.cu)
    def eye_key(inst):
        c = np.asarray(cam_pos, dtype=F)
        lo, hi = np.asarray(inst["aabb_min"], F), np.asarray(inst["aabb_max"], F)
        d = np.maximum(np.maximum(lo - c, F(0)), c - hi).astype(F)
        return float((d[0] * d[0] + d[1] * d[1] + d[2] * d[2]).astype(F))

    order = sorted(range(len(instances)), key=lambda i: (eye_key(instances[i]), i))
    for inst in (instances[i] for i in order):
        ok = _slab(np.asarray(inst["aabb_min"], F), np.asarray(inst["aabb_max"], F), o_w, inv_d_w, t_max)
        if not ok.any():
            continue
        q = np.asarray(inst["rotation"], dtype=F)
        inv_s = (F(1) / np.asarray(inst["scale"], dtype=F)).astype(F)
        o = (inv_s * quat_rotate(quat_inv(q), (o_w - np.asarray(inst["position"], F)).astype(F))).astype(F)
        d = (inv_s * quat_rotate(quat_inv(q), d_w)).astype(F)
        t_scale = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]).astype(F)).astype(F)
        t_obj = (t_max * t_scale).astype(F)
        d = (d * (F(1) / t_scale)[..., None]).astype(F)
        with np.errstate(divide="ignore"):
            inv_d = (F(1) / d).astype(F)
        ad = np.abs(d)
        kz = np.where((ad[..., 0] > ad[..., 1]) & (ad[..., 0] > ad[..., 2]), 0,
                      np.where(ad[..., 1] > ad[..., 2], 1, 2))
        kx = (kz + 1) % 3
        ky = (kx + 1) % 3
        dkz = np.take_along_axis(d, kz[..., None], axis=-1)[..., 0]
        kx, ky = np.where(dkz < 0, ky, kx), np.where(dkz < 0, kx, ky)
        ikz = np.take_along_axis(inv_d, kz[..., None], axis=-1)[..., 0]
        Sx = (np.take_along_axis(d, kx[..., None], axis=-1)[..., 0] * ikz).astype(F)
        Sy = (np.take_along_axis(d, ky[..., None], axis=-1)[..., 0] * ikz).astype(F)
        Sz = ikz
        m = meshes[inst["object_id"]]
        hit_here = np.zeros((res, res), dtype=bool)
        for tri in range(int(m["first"]), int(m["first"]) + int(m["count"])):
            a, b, c = (verts[indices[tri][k]].astype(F) for k in range(3))
            okt, t = _ray_triangle(a, b, c, kx, ky, kz, Sx, Sy, Sz, o, t_obj)
            okt = okt & ok
            t_obj = np.where(okt, t, t_obj)
            hit_here |= okt
        t_new = (t_obj / t_scale).astype(F)
        t_max = np.where(ok, t_new, t_max)
        hit |= hit_here
    return np.where(hit, t_max, F(0)).astype(F)
