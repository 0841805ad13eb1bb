// TEST INFRASTRUCTURE ONLY.  Known-answer / equivalence probe for the GJK
// closest-point routine behind sphere - hull contacts.  Compiled twice by
// oracle/Makefile:
//   _ref/gjk_probe_ref   the REFERENCE: src/physics/gjk.hpp (private header) and
//                        geo::hullClosestPointToOriginGJK from libmadrona_ref.a
//   _ref/gjk_probe_mine  this engine's madrona_b200/device/madrona/gjk.hpp
// tests/test_gjk.py requires identical output bits and applies the assertions of
// the reference's own tests (tests/gjk.cpp:18-47) to both.
#include <cstdint>
#include <cstdio>
#include <cstring>

#ifdef PROBE_REF
#include <madrona/geo.hpp>
#include "gjk.hpp"
#else
#include <madrona/gjk.hpp>
#endif

using namespace madrona;
using namespace madrona::math;

struct Solved {
    Vector3 v;
    float len2;
    float w[4];
};

#ifdef PROBE_REF
static Solved pack(const geo::GJKSimplexSolveState &s)
{
    return Solved { s.v, s.vLen2, { s.lambdas[0], s.lambdas[1], s.lambdas[2], s.lambdas[3] } };
}
static Solved solve2(Vector3 a, Vector3 b) { return pack(geo::gjkSolve2Simplex(a, b)); }
static Solved solve3(Vector3 a, Vector3 b, Vector3 c) { return pack(geo::gjkSolve3Simplex(a, b, c)); }
static Solved solve4(Vector3 a, Vector3 b, Vector3 c, Vector3 d) { return pack(geo::gjkSolve4Simplex(a, b, c, d)); }
static float hullClosest(Vector3 *verts, uint32_t n, float tol2, Vector3 *closest)
{
    geo::HalfEdgeMesh mesh {};
    mesh.vertices = verts;
    mesh.numVertices = n;
    return geo::hullClosestPointToOriginGJK(mesh, tol2, closest);
}
#else
static Solved pack(const geo::SimplexClosest &s)
{
    return Solved { s.v, s.len2, { s.w0, s.w1, s.w2, s.w3 } };
}
static Solved solve2(Vector3 a, Vector3 b) { return pack(geo::gjk_detail::closestOnSegment(a, b)); }
static Solved solve3(Vector3 a, Vector3 b, Vector3 c) { return pack(geo::gjk_detail::closestOnTriangle(a, b, c)); }
static Solved solve4(Vector3 a, Vector3 b, Vector3 c, Vector3 d)
{
    return pack(geo::gjk_detail::closestOnTetrahedron(a, b, c, d));
}
static float hullClosest(Vector3 *verts, uint32_t n, float tol2, Vector3 *closest)
{
    return geo::hullVerticesClosestPointToOriginGJK(verts, n, tol2, closest);
}
#endif

static void pf(const char *tag, float v)
{
    uint32_t b;
    memcpy(&b, &v, 4);
    printf("%s %08x\n", tag, b);
}
static void ps(const char *tag, const Solved &s)
{
    pf(tag, s.v.x); pf(tag, s.v.y); pf(tag, s.v.z); pf(tag, s.len2);
    for (int i = 0; i < 4; i++) pf(tag, s.w[i]);
}

static uint64_t lcg_state = 0x2545F4914F6CDD1Dull;
static float unit()          // [0, 1)
{
    lcg_state = lcg_state * 6364136223846793005ull + 1442695040888963407ull;
    return (float)((lcg_state >> 40) & 0xFFFFFF) / 16777216.f;
}
static float range(float lo, float hi) { return lo + (hi - lo) * unit(); }
static Vector3 point(float r) { return Vector3 { range(-r, r), range(-r, r), range(-r, r) }; }

int main()
{
    // ---- the reference's own test vectors (tests/gjk.cpp:20-47)
    {
        Vector3 Y[4] { { 0.814353108f, 0.195752025f, -0.698764443f },
                       { -0.784147143f, 0.126484752f, 0.701235533f },
                       { -0.784147143f, 0.126484752f, -0.698764443f },
                       { -0.784147143f, 0.126484752f, 0.701235533f } };
        Solved s3 = solve3(Y[0], Y[1], Y[2]);
        Solved s4 = solve4(Y[0], Y[1], Y[2], Y[3]);
        pf("kat_dup_diff", s4.len2 - s3.len2);
        ps("kat_dup_s3", s3);
        ps("kat_dup_s4", s4);
    }
    {
        Vector3 Y[4] { { 0.793287277f, 2.86326122f, -0.700307727f },
                       { -0.794485092f, -0.542466521f, 0.699692249f },
                       { 0.80550468f, -0.536717057f, -0.700307727f },
                       { -0.794485092f, -0.542466521f, -0.700307727f } };
        ps("kat_origin_s4", solve4(Y[0], Y[1], Y[2], Y[3]));
    }

    // ---- sub-simplex solves on random points (offset so the origin is outside / inside)
    for (int i = 0; i < 300; i++) {
        const Vector3 off = (i % 3 == 0) ? Vector3 { 0, 0, 0 } : point(2.f);
        const Vector3 a = point(1.f) + off, b = point(1.f) + off, c = point(1.f) + off, d = point(1.f) + off;
        ps("s2", solve2(a, b));
        ps("s3", solve3(a, b, c));
        ps("s4", solve4(a, b, c, d));
    }

    // ---- full distance queries: scaled / rotated / translated boxes and point clouds
    for (int i = 0; i < 1200; i++) {
        Vector3 verts[16];
        uint32_t n;
        const Quat rot = Quat::angleAxis(range(-3.f, 3.f), point(1.f).normalize());
        const Vector3 scale { range(0.2f, 3.f), range(0.2f, 3.f), range(0.2f, 3.f) };
        // every third case keeps the hull over the origin (distance 0 paths)
        const Vector3 pos = (i % 3 == 0) ? point(0.3f) : point(4.f);
        if (i % 2 == 0) {
            n = 8;
            for (uint32_t k = 0; k < 8; k++) {
                const Vector3 corner { (k & 1) ? 0.5f : -0.5f, (k & 2) ? 0.5f : -0.5f, (k & 4) ? 0.5f : -0.5f };
                verts[k] = rot.rotateVec(Vector3 { corner.x * scale.x, corner.y * scale.y, corner.z * scale.z }) + pos;
            }
        } else {
            // sheared octahedron (random point clouds trip the reference's own
            // monotonicity assert inside geo.cpp, so only true hull vertex sets)
            n = 6;
            const float shear = range(-0.4f, 0.4f);
            for (uint32_t k = 0; k < 6; k++) {
                Vector3 p { 0, 0, 0 };
                const float sgn = (k & 1) ? 1.f : -1.f;
                if (k / 2 == 0) p = Vector3 { sgn * scale.x, sgn * shear, 0 };
                else if (k / 2 == 1) p = Vector3 { 0, sgn * scale.y, sgn * shear };
                else p = Vector3 { sgn * shear, 0, sgn * scale.z };
                verts[k] = rot.rotateVec(p) + pos;
            }
        }
        Vector3 closest { 0, 0, 0 };
        const float dist2 = hullClosest(verts, n, 1e-10f, &closest);
        pf("hull_dist2", dist2);
        if (dist2 != 0.f) {
            pf("hull_pt", closest.x); pf("hull_pt", closest.y); pf("hull_pt", closest.z);
        }
    }
    return 0;
}
