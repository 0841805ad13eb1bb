"""Small identities the ray caster's fast path relies on (madrona_b200/csrc/kernels_render.cu),
restated on the CPU:
  * `byteAsFloat`: byte q of a packed word, moved under the exponent 0x4B000000 (= 2^23), minus
    2^23 is exactly float(q) -- the quantised child boxes are decoded without int->float
    conversions;
  * `instanceXform`: the world -> object matrix built from the instance quaternion is the
    matrix of Quat::rotateVec (device/madrona/math.hpp: v + 2 (w (u x v) + u x (u x v))),
    transposed and divided by the scale, also for a quaternion that is not exactly unit
    length -- so object-space rays equal the previous per-ray quaternion algebra;
  * the object-space ray keeps the world ray's parametrisation (no renormalisation)."""
import numpy as np


def test_byte_under_the_exponent_is_the_integer_value():
    for word in (0x00000000, 0xFFFFFFFF, 0x80FF017F, 0x12345678):
        for i in range(4):
            q = (word >> (8 * i)) & 0xFF
            # __byte_perm(word, 0x4B000000, 0x7650 + i): byte 0 = byte i of word, bytes 1-2 = 0, byte 3 = 0x4B
            packed = np.array(0x4B000000 | q, dtype=np.uint32)
            assert packed.view(np.float32) - np.float32(8388608.0) == np.float32(q)


def _rotate_vec(q, v):
    w, u = q[0], q[1:]
    return v + 2.0 * (w * np.cross(u, v) + np.cross(u, np.cross(u, v)))


def _instance_xform(q, scale):
    w, x, y, z = q
    diag = 1.0 - 2.0 * (x * x + y * y + z * z)
    R = np.array([[diag + 2 * x * x, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                  [2 * (x * y + w * z), diag + 2 * y * y, 2 * (y * z - w * x)],
                  [2 * (x * z - w * y), 2 * (y * z + w * x), diag + 2 * z * z]])
    return (R.T / scale[:, None]), R


def test_instance_matrix_equals_the_quaternion_algebra_it_replaces():
    rng = np.random.default_rng(5)
    for _ in range(200):
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        q *= 1.0 + rng.uniform(-1e-3, 1e-3)                  # physics leaves rotations slightly off unit length
        scale = rng.uniform(0.2, 5.0, size=3)
        pos = rng.uniform(-10, 10, size=3)
        M, R = _instance_xform(q, scale)
        v = rng.normal(size=3)
        assert np.allclose(R @ v, _rotate_vec(q, v), rtol=1e-12, atol=1e-12)
        q_inv = np.array([q[0], -q[1], -q[2], -q[3]])
        o, d = rng.uniform(-10, 10, size=3), rng.normal(size=3)
        assert np.allclose(M @ (o - pos), _rotate_vec(q_inv, o - pos) / scale, rtol=1e-10, atol=1e-10)
        # same parametrisation: object-space point at t == object-space image of the world point at t
        t = rng.uniform(0.1, 50)
        assert np.allclose(M @ (o - pos) + t * (M @ d), M @ ((o + t * d) - pos), rtol=1e-9, atol=1e-9)
