// TEST INFRASTRUCTURE ONLY.  Prints a deterministic battery of madrona::rand /
// madrona::math results as hex.  Compiled twice by oracle/Makefile:
//   _ref/kat_probe_ref   against the REFERENCE headers (/root/reference/include)
//   _ref/kat_probe_mine  against this engine's headers (madrona_b200/device)
// tests/test_oracle_kat.py requires the two outputs to be identical, which pins
// the engine's math / RNG headers to the reference bit for bit (same compiler,
// same flags, -ffp-contract=off), and checks the reference's own known-answer
// values (tests/rand.cpp:131-141, tests/math.cpp:23-48) on both.
#include <madrona/math.hpp>
#include <madrona/rand.hpp>

#include <cstdio>
#include <cstring>

using namespace madrona;
using namespace madrona::math;

static void pf(const char *tag, float v)
{
    uint32_t b;
    memcpy(&b, &v, 4);
    printf("%s %08x\n", tag, b);
}
static uint32_t bits(float v)
{
    uint32_t b;
    memcpy(&b, &v, 4);
    return b;
}
static void pv(const char *tag, Vector3 v) { pf(tag, v.x); pf(tag, v.y); pf(tag, v.z); }
static void pq(const char *tag, Quat q) { pf(tag, q.w); pf(tag, q.x); pf(tag, q.y); pf(tag, q.z); }

int main()
{
    // ---- reference known-answer values
    RandKey k { 0xFFFFFFFFu, 0u };
    printf("kat_bits32 %08x\n", rand::bits32(k));
    printf("kat_sampleI32 %d\n", rand::sampleI32(k, 0, 64));
    printf("kat_sampleI32Biased %d\n", rand::sampleI32Biased(k, 0, 64));
    pq("kat_q1", Quat::angleAxis(0, { 0, 1, 0 }));
    Quat q2 = Quat::angleAxis(toRadians(45), { 0, 1, 0 });
    Quat q3 = Quat::angleAxis(toRadians(45), { 1, 0, 0 });
    pq("kat_q2", q2);
    pq("kat_q3", q3);
    pq("kat_m1", q2 * q3);

    // ---- RNG streams
    for (uint32_t seed = 0; seed < 4; seed++) {
        RandKey key = rand::initKey(seed * 7919u + 1u, seed);
        printf("key %08x %08x\n", key.a, key.b);
        for (uint32_t i = 0; i < 4; i++) {
            RandKey s = rand::split_i(key, i, i * 3u);
            printf("split %08x %08x\n", s.a, s.b);
            printf("i32 %d %d %d\n", rand::sampleI32(s, -20, 2), rand::sampleI32(s, 0, 1000003),
                   rand::sampleI32Biased(s, 3, 77));
            pf("uniform", rand::sampleUniform(s));
            printf("bool %d\n", (int)rand::sampleBool(s));
            Vector2 u2 = rand::sample2xUniform(s);
            pf("u2", u2.x); pf("u2", u2.y);
        }
        RNG rng(seed + 11u);
        for (int i = 0; i < 6; i++) {
            pf("rng_uniform", rng.sampleUniform());
            printf("rng_i32 %d\n", rng.sampleI32(0, 97));
        }
    }

    // ---- math on pseudo-random inputs
    RNG r(12345u);
    auto rf = [&]() { return r.sampleUniform() * 4.f - 2.f; };
    for (int it = 0; it < 24; it++) {
        Vector3 a { rf(), rf(), rf() }, b { rf(), rf(), rf() }, c { rf(), rf(), rf() };
        Quat q = Quat { rf(), rf(), rf(), rf() }.normalize();
        Quat p = Quat { rf(), rf(), rf(), rf() }.normalize();
        Diag3x3 s { 0.5f + r.sampleUniform(), 0.5f + r.sampleUniform(), 0.5f + r.sampleUniform() };
        pf("dot", dot(a, b));
        pv("cross", cross(a, b));
        pv("norm", a.normalize());
        pf("len", a.length());
        pf("invlen", a.invLength());
        pf("dist", a.distance(b));
        Vector3 fa, fb;
        Vector3 n = a.normalize();
        n.frame(&fa, &fb);
        pv("frame_a", fa); pv("frame_b", fb);
        pq("qmul", q * p);
        pq("qnorm", (q + p).normalize());
        pv("rot", q.rotateVec(a));
        pv("rotinv", q.inv().rotateVec(a));
        pq("basis", Quat::fromBasis(q.rotateVec({ 1, 0, 0 }), q.rotateVec({ 0, 1, 0 }), q.rotateVec({ 0, 0, 1 })));
        Mat3x3 m = Mat3x3::fromQuat(q);
        pv("m0", m[0]); pv("m1", m[1]); pv("m2", m[2]);
        Mat3x3 rs = Mat3x3::fromRS(q, s);
        pv("rs0", rs[0]); pv("rs2", rs[2]);
        pv("mv", rs * a);
        Mat3x3 prod = m * rs;
        pv("mm1", prod[1]);
        pf("det", rs.determinant());
        Mat3x3 ms = m * s.inv();
        pv("minv", (ms * a).normalize());
        Symmetric3x3 x = Symmetric3x3::vvT(b);
        Symmetric3x3 axat = Symmetric3x3::AXAT(m, x);
        pv("axat_d", axat.diag); pv("axat_o", axat.off);
        Symmetric3x3 aat = Symmetric3x3::AAT(rs);
        pv("aat_d", aat.diag); pv("aat_o", aat.off);
        AABB box { Vector3::min(a, b), Vector3::max(a, b) };
        AABB t = box.applyTRS(c, q, s);
        pv("trs_min", t.pMin); pv("trs_max", t.pMax);
        pf("area", box.surfaceArea());
        pv("cent", box.centroid());
        printf("ovl %d %d\n", (int)box.overlaps(t), (int)box.intersects(t));
        Mat3x4 m34 = Mat3x4::fromTRS(c, q, s);
        pv("txfm", m34.txfmPoint(a));
        pv("txdir", m34.txfmDir(b));
        AABB e = AABB::invalid();
        e.expand(a); e.expand(b); e.expand(c);
        pv("exp_min", e.pMin); pv("exp_max", e.pMax);
        pv("div", a / 3.f); pv("div2", 2.f / b);

        // slab test incl. its edge cases: axis-parallel rays (infinite inverse
        // direction), origins exactly on a box plane (0 * inf = NaN), touching
        // intervals (t_box_min == t_box_max)
        for (int variant = 0; variant < 6; variant++) {
            Vector3 o = c, d = (b - c).normalize();
            if (variant == 1) d = Vector3 { d.x, d.y, 0.f };
            if (variant == 2) { d = Vector3 { 0.f, 1.f, 0.f }; o.x = box.pMin.x; }
            if (variant == 3) { d = Vector3 { 1.f, 0.f, 0.f }; o.z = box.pMax.z; o.x = box.pMin.x - 1.f; }
            if (variant == 4) { o = box.pMax; d = Vector3 { 1.f, 1.f, 1.f }; }
            if (variant == 5) { o = box.pMin - Vector3 { 1.f, 0.f, 0.f }; d = Vector3 { 1.f, 0.f, 0.f }; }
            Diag3x3 inv_d = Diag3x3::fromVec(d).inv();
            float t_hit = -1.f, t_far = -1.f;
            float t_max = variant == 5 ? 1.f : 100.f;
            bool hit = box.rayIntersects(o, inv_d, 0.f, t_max, t_hit, t_far);
            bool hit2 = box.rayIntersects(o, inv_d, 0.f, t_max);
            printf("ray %d %d %08x %08x\n", (int)hit, (int)hit2, hit ? bits(t_hit) : 0u, hit ? bits(t_far) : 0u);
        }
    }
    return 0;
}
