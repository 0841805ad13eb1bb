"""Ray caster fast path (madrona_b200/csrc/kernels_render.cu, worlds with <= 64 instances):
the per-view frustum cull and the per-warp 8 x 4-pixel tile cull must be CONSERVATIVE -- a
box that any pixel ray of the tile enters may never be dropped.  CPU model of the kernel's
plane construction (`boxOutside`, the tile edges a quarter pixel wider, the view planes with
their pad) in float32 against exact float64 rays, over random cameras, fields of view,
resolutions that are not multiples of the tile size, and boxes behind / around the camera.
The GPU oracle test (tests/test_render_bvh.py) pins the pixels of one scene; this pins the
geometry argument everywhere else."""
import numpy as np

F = np.float32


def _box_outside(lo, hi, n):
    c = F(0.5) * (lo + hi)
    h = F(0.5) * (hi - lo)
    reach = n[0] * c[0] + n[1] * c[1] + n[2] * c[2] + abs(n[0]) * h[0] + abs(n[1]) * h[1] + abs(n[2]) * h[2]
    return bool(reach < 0)


def _ray_enters(lo, hi, d):
    """exact slab test from the origin, t in [0, 1e4]"""
    t0, t1 = 0.0, 1e4
    for a in range(3):
        if d[a] == 0:
            if not (lo[a] <= 0 <= hi[a]):
                return False
            continue
        x, y = lo[a] / d[a], hi[a] / d[a]
        t0, t1 = max(t0, min(x, y)), min(t1, max(x, y))
    return t0 <= t1


def _random_rotation(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def test_view_and_tile_culls_never_drop_a_box_a_pixel_ray_enters():
    rng = np.random.default_rng(11)
    kept_and_hit, culled = 0, 0
    for case in range(50):
        R = _random_rotation(rng)
        u, forward = R[:, 0], R[:, 1]
        vv = np.cross(forward, u)
        vv /= np.linalg.norm(vv)
        res = int(rng.choice([16, 20, 40, 64]))
        h = 1.0 / (1.0 / np.tan(np.radians(rng.uniform(40, 110) * 0.5)))
        viewport = 2 * h
        uf, ff, vf = u.astype(F), forward.astype(F), vv.astype(F)
        hf, vpf = F(h), F(viewport)
        # boxes relative to the camera: all around it, some containing it, some huge
        n = 24
        c = rng.uniform(-12, 12, size=(n, 3))
        half = rng.uniform(0.1, 4, size=(n, 3))
        half[0] = [1e4, 1e4, 0.5]                     # a ground-plane style slab
        lo64, hi64 = c - half, c + half
        lo32, hi32 = lo64.astype(F), hi64.astype(F)
        pad = abs(hf) * (F(1) + F(2) / F(res))
        in_view = [not (_box_outside(lo32[k], hi32[k], ff) or
                        _box_outside(lo32[k], hi32[k], uf + pad * ff) or _box_outside(lo32[k], hi32[k], pad * ff - uf) or
                        _box_outside(lo32[k], hi32[k], vf + pad * ff) or _box_outside(lo32[k], hi32[k], pad * ff - vf))
                   for k in range(n)]
        tiles_x = (res + 7) // 8
        for tile in range(tiles_x * ((res + 3) // 4)):
            tx0, ty0 = (tile % tiles_x) * 8, (tile // tiles_x) * 4
            inv_res = F(1) / F(res)
            a0 = (F(tx0) * inv_res - F(0.25) * inv_res - F(0.5)) * vpf
            a1 = (F(tx0 + 8) * inv_res + F(0.25) * inv_res - F(0.5)) * vpf
            b0 = (F(ty0) * inv_res - F(0.25) * inv_res - F(0.5)) * vpf
            b1 = (F(ty0 + 4) * inv_res + F(0.25) * inv_res - F(0.5)) * vpf
            al, ar, bl, br = min(a0, a1), max(a0, a1), min(b0, b1), max(b0, b1)
            planes = [uf - al * ff, ar * ff - uf, vf - bl * ff, br * ff - vf]
            keep = [in_view[k] and not any(_box_outside(lo32[k], hi32[k], p) for p in planes) for k in range(n)]
            for py in range(ty0, min(ty0 + 4, res)):
                for px in range(tx0, min(tx0 + 8, res)):
                    pu, pv = (px + 0.5) / res, (py + 0.5) / res
                    d = forward + (pu - 0.5) * viewport * u + (pv - 0.5) * viewport * vv
                    d /= np.linalg.norm(d)
                    for k in range(n):
                        # boxes shrunk by 1e-4 so that float32 rounding of a grazing ray does not count
                        shrink = 1e-4 * (1 + np.abs(c[k]))
                        if _ray_enters(lo64[k] + shrink, hi64[k] - shrink, d):
                            assert keep[k], (case, tile, px, py, k)
                            kept_and_hit += 1
            culled += keep.count(False)
    assert kept_and_hit > 4000 and culled > 4000     # both sides of the cull are exercised
