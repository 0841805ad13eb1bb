"""GJK closest point (sphere - hull contacts): pins madrona_b200/device/madrona/gjk.hpp
to the reference's src/physics/gjk.hpp + geo::hullClosestPointToOriginGJK.

* oracle/gjk_probe.cpp is compiled against the reference (private header +
  libmadrona_ref.a) and against the engine's header; the two binaries must print
  identical bits: 900 sub-simplex solves, 1200 hull distance queries (a third of
  them with the origin inside the hull);
* the assertions of the reference's own tests (tests/gjk.cpp:18-47) are applied
  to both outputs.
"""
import hashlib
import os
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "gjk_probe_ref")
MINE = os.path.join(ROOT, "oracle", "_ref", "gjk_probe_mine")
GOLDEN = os.path.join(ROOT, "tests", "golden", "gjk_probe.sha256")

needs_probes = pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(MINE)),
                                  reason="oracle/_ref not built")


def _run(path):
    return subprocess.run([path], capture_output=True, text=True, check=True).stdout


def _f(hexbits):
    return struct.unpack("<f", struct.pack("<I", int(hexbits, 16)))[0]


def _tagged(text):
    out = {}
    for ln in text.splitlines():
        tag, val = ln.split()
        out.setdefault(tag, []).append(_f(val))
    return out


@needs_probes
def test_engine_header_matches_reference_bit_for_bit():
    ref, mine = _run(REF), _run(MINE)
    assert len(ref.splitlines()) > 10000
    assert ref == mine
    # the committed digest was produced by the reference build (tests/golden/make_golden.py)
    if os.path.exists(GOLDEN):
        assert hashlib.sha256(ref.encode()).hexdigest() == open(GOLDEN).read().strip()


@needs_probes
@pytest.mark.parametrize("binary", [REF, MINE], ids=["reference", "engine"])
def test_reference_gjk_known_answers(binary):
    vals = _tagged(_run(binary))
    # tests/gjk.cpp:18-32 Solve4SimplexDuplicatePoint
    assert vals["kat_dup_diff"][0] <= 1e-5
    # tests/gjk.cpp:34-47 Solve4SimplexAroundOrigin
    vx, vy, vz, len2 = vals["kat_origin_s4"][:4]
    assert abs(vx) < 1e-5 and abs(vy) < 1e-5 and abs(vz) < 1e-5
    assert len2 < 1e-5


@pytest.mark.skipif(not os.path.exists(MINE), reason="oracle/_ref not built")
def test_engine_header_against_committed_digest():
    if not os.path.exists(GOLDEN):
        pytest.skip("no committed digest")
    assert hashlib.sha256(_run(MINE).encode()).hexdigest() == open(GOLDEN).read().strip()
