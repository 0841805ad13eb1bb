// Distance from the origin to a convex point set (GJK with the signed-volumes
// sub-algorithm of Montanari, Petrinic & Barbieri, "Improving the GJK Algorithm
// for Faster and More Reliable Distance Queries Between Two Convex Objects",
// ToG 2017), as used by the reference for sphere - hull contacts
// (src/physics/gjk.hpp, src/physics/geo.cpp:38-59, narrowphase.cpp:1326-1402).
//
// Written for registers: the simplex is four named slots and every sub-simplex
// solve is a straight-line function, so nothing is indexed dynamically on the
// GPU.  The arithmetic follows the reference expression by expression (operand
// order included: float addition is not associative), including its departures
// from the paper (negated sign tests so degenerate simplices still examine
// their faces, the extra fourth face / third edge); oracle/gjk_probe.cpp pins
// this header to the reference bit for bit.
#pragma once

#include <madrona/math.hpp>

#include <cfloat>

namespace madrona::geo {

struct SimplexClosest {
    math::Vector3 v;       // point of the simplex closest to the origin
    float len2;            // |v|^2
    float w0, w1, w2, w3;  // barycentric weight of each simplex slot (0: slot not needed)
};

namespace gjk_detail {

MB2_HD inline bool sameStrictSign(float a, float b)
{
    return (a > 0 && b > 0) || (a < 0 && b < 0);
}

MB2_HD inline SimplexClosest closestOnPoint(math::Vector3 p0)
{
    return SimplexClosest { p0, p0.length2(), 1.f, 0.f, 0.f, 0.f };
}

// segment: weights come back in (w0, w1) for (p0, p1)
MB2_HD inline SimplexClosest closestOnSegment(math::Vector3 p0, math::Vector3 p1)
{
    using math::Vector3;
    // the paper's s1 is the newest point
    const Vector3 a = p1, b = p0;
    const Vector3 ab = b - a;
    const float ab_len2 = ab.length2();

    // project onto the coordinate axis along which the segment is longest
    float extent = a.x - b.x;
    float a_c = a.x, b_c = b.x;
    {
        const float ey = a.y - b.y;
        if (fabsf(ey) > fabsf(extent)) {
            extent = ey;
            a_c = a.y;
            b_c = b.y;
        }
        const float ez = a.z - b.z;
        if (fabsf(ez) > fabsf(extent)) {
            extent = ez;
            a_c = a.z;
            b_c = b.z;
        }
    }
    // that coordinate of the origin's projection onto the line
    const float proj_c = (math::dot(b, ab) / ab_len2) * (a_c - b_c) + b_c;
    const float part_b = proj_c - b_c;
    const float part_a = a_c - proj_c;

    if (sameStrictSign(extent, part_b) && sameStrictSign(extent, part_a)) {
        const float wb = part_a / extent;
        const Vector3 v = a + ab * wb;
        const float wa = 1.f - wb;
        return SimplexClosest { v, v.length2(), wb, wa, 0.f, 0.f };
    }
    return SimplexClosest { a, a.length2(), 0.f, 1.f, 0.f, 0.f };
}

// triangle: weights in (w0, w1, w2) for (p0, p1, p2)
MB2_HD inline SimplexClosest closestOnTriangle(math::Vector3 p0, math::Vector3 p1, math::Vector3 p2)
{
    using math::Vector3;
    const Vector3 a = p2, b = p1, c = p0;

    const Vector3 n = math::cross(b - a, c - a);
    const float n_len2 = n.length2();
    const Vector3 proj = math::dot(a, n) * n / n_len2;     // origin projected onto the plane

    // signed areas of the triangle's projections onto the three coordinate planes
    const float area_yz = b.y * c.z - c.y * b.z
                        - a.y * c.z + c.y * a.z
                        + a.y * b.z - b.y * a.z;
    const float area_xz = b.x * c.z - c.x * b.z
                        - a.x * c.z + c.x * a.z
                        + a.x * b.z - b.x * a.z;
    const float area_xy = b.x * c.y - c.x * b.y
                        - a.x * c.y + c.x * a.y
                        + a.x * b.y - b.x * a.y;
    const float abs_yz = fabsf(area_yz), abs_xz = fabsf(area_xz), abs_xy = fabsf(area_xy);

    // work in the plane where the triangle is largest
    float area;
    float au, av, bu, bv, cu, cv, pu, pv;
    if (abs_yz >= abs_xz && abs_yz >= abs_xy) {
        area = area_yz;
        au = a.y; av = a.z; bu = b.y; bv = b.z; cu = c.y; cv = c.z; pu = proj.y; pv = proj.z;
    } else if (abs_xz >= abs_xy) {
        area = area_xz;
        au = a.x; av = a.z; bu = b.x; bv = b.z; cu = c.x; cv = c.z; pu = proj.x; pv = proj.z;
    } else {
        area = area_xy;
        au = a.x; av = a.y; bu = b.x; bv = b.y; cu = c.x; cv = c.y; pu = proj.x; pv = proj.y;
    }

    // sub-areas with the projected origin in place of a, b, c
    const float sub_a = pu * bv + pv * cu + bu * cv
                      - pu * cv - pv * bu - cu * bv;
    const float sub_b = pu * cv + pv * au + cu * av
                      - pu * av - pv * cu - au * cv;
    const float sub_c = pu * av + pv * bu + au * bv
                      - pu * bv - pv * au - bu * av;

    const bool in_a = sameStrictSign(area, sub_a);
    const bool in_b = sameStrictSign(area, sub_b);
    const bool in_c = sameStrictSign(area, sub_c);

    if (in_a && in_b && in_c) {
        const float wb = sub_b / area;
        const float wc = sub_c / area;
        const float wa = 1.f - wb - wc;
        const Vector3 v = a * wa + b * wb + c * wc;
        return SimplexClosest { v, v.length2(), wc, wb, wa, 0.f };
    }

    // otherwise the closest point is on an edge: examine every edge whose
    // opposite sub-area failed the sign test, keep the nearest
    SimplexClosest best;
    best.len2 = FLT_MAX;
    if (!in_b) {
        const SimplexClosest e = closestOnSegment(p0, p2);
        best = SimplexClosest { e.v, e.len2, e.w0, 0.f, e.w1, 0.f };
    }
    if (!in_c) {
        const SimplexClosest e = closestOnSegment(p1, p2);
        if (e.len2 < best.len2) best = SimplexClosest { e.v, e.len2, 0.f, e.w0, e.w1, 0.f };
    }
    if (!in_a) {
        const SimplexClosest e = closestOnSegment(p0, p1);
        if (e.len2 < best.len2) best = SimplexClosest { e.v, e.len2, e.w0, e.w1, 0.f, 0.f };
    }
    return best;
}

MB2_HD inline float tripleProduct(math::Vector3 a, math::Vector3 b, math::Vector3 c)
{
    return math::dot(a, math::cross(b, c));
}

// tetrahedron: weights in (w0 .. w3) for (p0 .. p3)
MB2_HD inline SimplexClosest closestOnTetrahedron(math::Vector3 p0, math::Vector3 p1, math::Vector3 p2,
                                                  math::Vector3 p3)
{
    using math::Vector3;
    const Vector3 a = p3, b = p2, c = p1, d = p0;

    // cofactors of the bottom row of [a b c d; 1 1 1 1]: signed volumes of the
    // tetrahedra the origin spans with each face
    const float vol_a = -tripleProduct(b, c, d);
    const float vol_b = tripleProduct(a, c, d);
    const float vol_c = -tripleProduct(a, b, d);
    const float vol_d = tripleProduct(a, b, c);
    const float vol = vol_a + vol_b + vol_c + vol_d;

    const bool in_a = sameStrictSign(vol, vol_a);
    const bool in_b = sameStrictSign(vol, vol_b);
    const bool in_c = sameStrictSign(vol, vol_c);
    const bool in_d = sameStrictSign(vol, vol_d);

    if (in_a && in_b && in_c && in_d) {
        const float wa = vol_a / vol;
        const float wb = vol_b / vol;
        const float wc = vol_c / vol;
        const float wd = 1.f - wa - wb - wc;
        const Vector3 v = a * wa + b * wb + c * wc + d * wd;
        return SimplexClosest { v, v.length2(), wd, wc, wb, wa };
    }

    // the origin is outside: examine every face whose opposite volume failed
    // the sign test (a zero volume fails too, so flat tetrahedra still look at
    // their faces), keep the nearest
    SimplexClosest best;
    best.len2 = FLT_MAX;
    if (!in_b) {
        const SimplexClosest f = closestOnTriangle(p0, p1, p3);
        best = SimplexClosest { f.v, f.len2, f.w0, f.w1, 0.f, f.w2 };
    }
    if (!in_c) {
        const SimplexClosest f = closestOnTriangle(p0, p2, p3);
        if (f.len2 < best.len2) best = SimplexClosest { f.v, f.len2, f.w0, 0.f, f.w1, f.w2 };
    }
    if (!in_d) {
        const SimplexClosest f = closestOnTriangle(p1, p2, p3);
        if (f.len2 < best.len2) best = SimplexClosest { f.v, f.len2, 0.f, f.w0, f.w1, f.w2 };
    }
    if (!in_a) {
        const SimplexClosest f = closestOnTriangle(p0, p1, p2);
        if (f.len2 < best.len2) best = SimplexClosest { f.v, f.len2, f.w0, f.w1, f.w2, 0.f };
    }
    return best;
}

}

// Squared distance from the origin to the convex hull of the points `support`
// can return; *closest receives the closest point (when the result is 0 -- the
// origin is inside / touching -- it is the previous iterate, as in the reference).  support(dir) must return the point
// with the largest dot(point, dir).
template <typename SupportFn>
MB2_HD inline float gjkDistance2ToOrigin(SupportFn &&support, math::Vector3 first_dir, float tolerance2,
                                         math::Vector3 *closest)
{
    using math::Vector3;
    using namespace gjk_detail;

    Vector3 dir = first_dir;
    Vector3 s0 = Vector3::zero(), s1 = Vector3::zero(), s2 = Vector3::zero(), s3 = Vector3::zero();
    int count = 0;
    float dist2 = 0.f;
    float prev_dist2 = FLT_MAX;

    while (true) {
        const Vector3 w = support(dir);

        SimplexClosest sol;
        if (count == 0) {
            s0 = w;
            sol = closestOnPoint(s0);
        } else if (count == 1) {
            s1 = w;
            sol = closestOnSegment(s0, s1);
        } else if (count == 2) {
            s2 = w;
            sol = closestOnTriangle(s0, s1, s2);
        } else {
            s3 = w;
            sol = closestOnTetrahedron(s0, s1, s2, s3);
        }

        // keep only the slots that carry weight, in order
        {
            const Vector3 o0 = s0, o1 = s1, o2 = s2, o3 = s3;
            count = 0;
            auto keep = [&](Vector3 p, float weight) {
                if (weight == 0.f) return;
                if (count == 0) s0 = p;
                else if (count == 1) s1 = p;
                else if (count == 2) s2 = p;
                else s3 = p;
                count++;
            };
            keep(o0, sol.w0);
            keep(o1, sol.w1);
            keep(o2, sol.w2);
            keep(o3, sol.w3);
        }

        if (count == 4) {
            *closest = -dir;
            return 0.f;                       // the origin is enclosed
        }
        if (sol.len2 <= tolerance2) {
            *closest = -dir;
            return 0.f;
        }
        {
            float largest = s0.length2();
            const float l1 = s1.length2(), l2 = s2.length2(), l3 = s3.length2();
            if (1 < count && l1 > largest) largest = l1;
            if (2 < count && l2 > largest) largest = l2;
            if (3 < count && l3 > largest) largest = l3;
            // too close relative to the simplex for the direction to mean anything
            if (sol.len2 <= FLT_EPSILON * largest) {
                *closest = -dir;
                return 0.f;
            }
        }

        dist2 = sol.len2;
        dir = -sol.v;
        if (prev_dist2 - dist2 <= FLT_EPSILON * prev_dist2) break;   // no longer improving
        prev_dist2 = dist2;
    }

    *closest = -dir;
    return dist2;
}

// == geo::hullClosestPointToOriginGJK (include/madrona/geo.hpp:71) on an
// explicit vertex array (the caller has already placed the hull relative to the
// query point).
MB2_HD inline float hullVerticesClosestPointToOriginGJK(const math::Vector3 *vertices, uint32_t num_vertices,
                                                        float tolerance2, math::Vector3 *closest)
{
    using math::Vector3;
    auto support = [vertices, num_vertices](Vector3 dir) {
        float best_dot = -FLT_MAX;
        Vector3 best = Vector3::zero();
        for (uint32_t i = 0; i < num_vertices; i++) {
            const Vector3 p = vertices[i];
            const float along = math::dot(p, dir);
            if (along > best_dot) {
                best_dot = along;
                best = p;
            }
        }
        return best;
    };
    return gjkDistance2ToOrigin(support, -vertices[0], tolerance2, closest);
}

}
